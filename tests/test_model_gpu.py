"""Parity of the CUDA engine (through the C ABI / MSAUWrapper) with the oracle and the golden fixtures.

Tolerances (fp32 engine vs fp32 reference, different summation order): |dlogit| <= 5e-4 absolute,
arg-max agreement >= 99.9 % (BASELINE.json north_star); gradients 2e-3 relative to the tensor's max."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import msau_b200
from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input

pytestmark = pytest.mark.gpu
LOGIT_ATOL = 5e-4


@pytest.fixture(params=[1, 0], ids=["tcgen05", "fp32core"])
def tc(request):
    """Run every model test on both convolution paths: the tcgen05 implicit-GEMM kernels (bf16 hi/lo split, the
    default) and the fp32 CUDA-core kernels."""
    _lib.set_option("tensor_core_conv", request.param)
    yield request.param
    _lib.set_option("tensor_core_conv", 1)
FIXTURES = ["model_s3r2_c12", "model_s4r2_c96", "model_s2r3_c8", "model_s6r3_c16", "model_s3r2_c12_k17"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta, om.MsauConfig(**meta["cfg"])


def build(cfg, sd):
    m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=cfg.feat_root,
                                                              scale_space_num=cfg.scale_space_num, res_depth=cfg.res_depth))
    m.load_state_dict(sd)
    return m.cuda()


def plan_tensors(m, pl):
    """Every internal activation (and its gradient) of the last training forward, as NCHW CPU tensors."""
    L = _lib.lib()
    pk, act, n = C.c_longlong(), C.c_longlong(), C.c_int()
    _lib.check(L.msau_debug_layout(pl.handle, C.byref(pk), C.byref(act), C.byref(n)))
    base, _ = pl.ws_ptr(True)
    ws = pl.workspace(True)
    shift = (base - ws.data_ptr())
    fl = ws[shift:shift + (ws.numel() - shift) // 4 * 4].view(torch.float32)
    out = []
    for i in range(n.value):
        off, c, h, w = C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(L.msau_debug_tensor(pl.handle, i, C.byref(off), C.byref(c), C.byref(h), C.byref(w)))
        cnt = pl.B * h.value * w.value * c.value
        a = fl[pk.value + off.value: pk.value + off.value + cnt].view(pl.B, h.value, w.value, c.value).permute(0, 3, 1, 2).cpu()
        g = fl[pk.value + act.value + off.value: pk.value + act.value + off.value + cnt].view(pl.B, h.value, w.value, c.value).permute(0, 3, 1, 2).cpu()
        out.append((a, g))
    return out


@pytest.mark.parametrize("name", FIXTURES)
def test_forward_matches_golden(golden_dir, name, tc):
    z, meta, cfg = load(golden_dir, name)
    sd = om.init_state_dict(cfg, meta["seed"])
    x, labels = synth_input(cfg.channels, cfg.n_class, meta["B"], meta["H"], meta["W"], meta["seed"] + 1)
    m = build(cfg, sd).eval()
    with torch.no_grad():
        probs, logits, aux = m(x.cuda())
    lg, ax, pr = logits.cpu().numpy(), aux.cpu().numpy(), probs.cpu().numpy()
    assert np.abs(lg - z["logits"]).max() <= LOGIT_ATOL
    assert np.abs(ax - z["aux"]).max() <= LOGIT_ATOL
    assert np.abs(pr - z["probs"]).max() <= 1e-4
    agree = (lg.argmax(1) == z["logits"].argmax(1)).mean()
    assert agree >= 0.999
    am = m.predict_classes(x.cuda()).cpu().numpy()
    assert (am == lg.argmax(1)).all()
    # channels-last input (what the device rasteriser produces) gives the identical result
    cp = (cfg.channels + 3) // 4 * 4
    xn = torch.zeros(meta["B"], meta["H"], meta["W"], cp)
    xn[..., :cfg.channels] = x.permute(0, 2, 3, 1)
    am2 = m.predict_classes(xn.cuda(), layout=1).cpu().numpy()
    assert (am2 == am).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_every_activation_and_gradient_matches_oracle(golden_dir, name, tc):
    """Walks the engine's workspace tensor by tensor against the oracle's traced forward/backward: the first
    mismatch names the kernel at fault."""
    z, meta, cfg = load(golden_dir, name)
    sd = om.init_state_dict(cfg, meta["seed"])
    x, labels = synth_input(cfg.channels, cfg.n_class, meta["B"], meta["H"], meta["W"], meta["seed"] + 1)
    m = build(cfg, sd).train()
    _, logits, aux = m(x.cuda())
    loss = m.loss(logits, aux, labels.cuda())
    pl = m._last[0]
    torch.cuda.synchronize()
    got = plan_tensors(m, pl)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out, axo, trace = om.msau_forward_trace(leaves, cfg, x, retain_grad=True)
    ref_loss = om.batch_loss(out, axo, labels)
    ref_loss.backward()
    assert len(trace) == len(got)
    bad = []
    for (nm, t), (a, g) in zip(trace, got):
        if t is None:
            continue
        c = t.shape[1]
        err = (a[:, :c] - t.detach()).abs().max().item()
        scale = max(1.0, t.detach().abs().max().item())
        if not err <= 2e-4 * scale:
            bad.append(("act", nm, err))
        if t.grad is not None:
            # tensors that are the output of a ReLU (inner residual activations a_r, residual-block outputs rr,
            # coupling outputs cc / uc): the engine keeps the gradient w.r.t. the pre-activation (multiplied by
            # the ReLU mask, materialised once for all consumers), the oracle w.r.t. the ReLU output
            leaf = nm.rsplit(".", 1)[-1]
            post_relu = (leaf.startswith("a") and leaf != "att") or leaf in ("rr", "cc", "uc")
            want = t.grad * (t.detach() > 0) if post_relu else t.grad
            gs = max(want.abs().max().item(), 1e-12)
            diff = (g[:, :c] - want).abs()
            if tc:
                # bf16 hi/lo-split convs perturb pre-activations by ~1e-5: a handful of ReLU masks flip w.r.t. the
                # oracle; each flip changes one pixel's gradient completely and diffuses through the convs upstream.
                # Demand agreement in L2 (<= 3 %) and that ≥ 97 % of the elements are within 2 % of the max (the small
                # fixtures have few pixels per level, so a handful of flips already moves ~1 % of the elements).
                frac = (diff <= 2e-2 * gs).float().mean().item()
                l2 = (diff.double().norm() / max(want.double().norm().item(), 1e-30)).item()
                if not (frac >= 0.97 and l2 <= 0.03):
                    bad.append(("grad", nm, frac, l2))
            elif name == "model_s6r3_c16":
                # 6 levels x 3 blocks of pooling: one 2x2 max-pool window of this page holds two values that differ by less than
                # the fp32 summation-order noise (~1e-7), so the arg-max (and with it one gradient element) goes to the other
                # row than in the oracle and the difference diffuses upstream (scripts/trace_check.py shows the pair; the two
                # elements' SUM matches).  Same criterion as above with a 10x tighter element threshold.
                frac = (diff <= 2e-3 * gs).float().mean().item()
                l2 = (diff.double().norm() / max(want.double().norm().item(), 1e-30)).item()
                if not (frac >= 0.95 and l2 <= 0.02):
                    bad.append(("grad", nm, frac, l2))
            elif not diff.max().item() <= 2e-3 * gs:
                bad.append(("grad", nm, diff.max().item() / gs))
    assert not bad, bad[:12]
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * max(1.0, abs(float(ref_loss)))
    assert abs(float(loss) - z["page_losses"].mean()) <= 1e-4 * z["page_losses"].mean()


@pytest.mark.parametrize("name", FIXTURES)
def test_param_grads_and_train_step_match_golden(golden_dir, name, tc):
    z, meta, cfg = load(golden_dir, name)
    sd = om.init_state_dict(cfg, meta["seed"])
    x, labels = synth_input(cfg.channels, cfg.n_class, meta["B"], meta["H"], meta["W"], meta["seed"] + 1)
    m = build(cfg, sd).train()
    # ---- the reference's own loop shape: forward -> loss -> backward -> clip -> Adam (train...py:46-59)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=0.0001)
    m.zero_grad()
    _, ypred, ypred_aux = m(x.cuda())
    loss = m.loss(ypred, ypred_aux, labels.cuda())
    loss.backward()
    keys = [k for k, _ in om.param_schema(cfg)]
    named = dict(m.named_parameters())
    none = np.array([named[k].grad is None for k in keys])
    assert (none == z["grad_is_none"]).all()
    norms = np.array([0.0 if named[k].grad is None else float(named[k].grad.double().norm()) for k in keys])
    # atol: attention f.conv.bias has an analytically zero gradient (softmax shift invariance); with bf16 operands the
    # cancellation leaves ~5e-6 of noise next to gradient norms of 0.1 .. 1
    np.testing.assert_allclose(norms, z["grad_norms"], rtol=3e-2 if tc else 2e-3, atol=2e-5 if tc else 1e-6)
    for k in z.files:
        if k.startswith("grad::"):
            want = z[k]
            np.testing.assert_allclose(named[k[6:]].grad.cpu().numpy(), want, rtol=0, atol=(3e-2 if tc else 2e-3) * np.abs(want).max())
    total = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    assert abs(float(total) - float(z["total_norm"])) <= (2e-2 if tc else 2e-3) * float(z["total_norm"])
    opt.step()
    sums = np.array([float(named[k].detach().double().sum()) for k in keys])
    # Adam's first step is lr * sign-like: an element whose gradient is ~0 moves by +-1e-4 with the sign of the rounding noise,
    # so the tolerated drift of a tensor's SUM grows with its size (the 256x256x3x3 convs of the S6 model hold 590k weights)
    numel = np.array([named[k].numel() for k in keys], np.float64)
    tol = np.maximum(1e-2, 2e-7 * numel) if tc else np.full_like(numel, 2e-4)
    assert (np.abs(sums - z["param_sums_after_step"]) <= tol + 1e-4 * np.abs(z["param_sums_after_step"])).all()
    for k in z.files:
        if k.startswith("param_after::"):
            # Adam's first step moves every weight by ~lr*sign(g): elements with |g| ~ 0 may flip, so bound by lr
            assert np.abs(named[k[13:]].detach().cpu().numpy() - z[k]).max() <= 2.1e-4

    # ---- fused path: same numbers from train_step (forward+loss+backward+clip+Adam in the engine)
    m2 = build(cfg, sd).train()
    l2 = m2.train_step(x.cuda(), labels.cuda())
    assert abs(float(l2) - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))
    torch.cuda.synchronize()
    live = torch.tensor(m2._live_mask())
    for (k, p2), p1 in zip(m2.named_parameters(), m.parameters()):
        # attention f.conv.bias: analytically zero gradient, so its Adam update (+-lr) follows the sign of rounding noise,
        # which depends on the order of the weight-gradient atomics
        # (and with the 13 M weights of the S6 model a few more elements have |g| ~ eps: bounded by 2 * lr)
        tol = 2.5e-4 if k.endswith("attention_block.f.conv.bias") or name == "model_s6r3_c16" else 2e-6
        assert (p2.detach() - p1.detach()).abs().max().item() <= tol, k
        if not k.endswith("attention_block.f.conv.bias"):
            assert (p2.detach() - p1.detach()).abs().mean().item() <= (4e-7 if name == "model_s6r3_c16" else 2e-7), k
    assert abs(float(m2._adam["total"]) - float(z["total_norm"])) <= (2e-2 if tc else 2e-3) * float(z["total_norm"])
    # dead attention params untouched
    dead = [k for k, lv in zip(keys, m2._live_mask()) if not lv]
    for k in dead:
        assert torch.equal(dict(m2.named_parameters())[k].detach().cpu(), sd[k])


@pytest.mark.parametrize("option,value", [("conv3_fold", 0), ("structured_first_layer", 0), ("lrn_coop", 0), ("lrn_coop", 2), ("fuse_relu_mask", 0),
                                          ("wgrad_multi_plane", 0)])
def test_alternative_kernel_paths_match_oracle(golden_dir, option, value):
    """The default tensor-core path uses the kx-folded 3x3 kernel (conv3_tc.cu) and, for one-hot inputs, the id-gather first
    layer (first_layer.cu).  With either switched off the generic implicit-GEMM kernels do the same work; with a dense
    (not one-hot) input the structured first layer must step aside on its own (device flag).  LRN: thread-per-pixel kernels
    only (0) / lane-cooperative kernels at every width (2) instead of the measured per-width choice.  wgrad_multi_plane = 0: the
    per-plane weight-gradient kernels (wgrad_tc3 / wgrad_tc) instead of wgrad_tc4 for the >= 16-channel layers."""
    z, meta, cfg = load(golden_dir, "model_s4r2_c96")
    sd = om.init_state_dict(cfg, meta["seed"])
    x, labels = synth_input(cfg.channels, cfg.n_class, meta["B"], meta["H"], meta["W"], meta["seed"] + 1)
    _lib.set_option(option, value)
    try:
        m = build(cfg, sd).train()
        _, logits, aux = m(x.cuda())
        assert (logits.cpu() - torch.from_numpy(z["logits"])).abs().max().item() <= LOGIT_ATOL
        loss = m.loss(logits, aux, labels.cuda())
        loss.backward()
        keys = [k for k, _ in om.param_schema(cfg)]
        named = dict(m.named_parameters())
        norms = np.array([0.0 if named[k].grad is None else float(named[k].grad.double().norm()) for k in keys])
        np.testing.assert_allclose(norms, z["grad_norms"], rtol=3e-2, atol=2e-5)
    finally:
        _lib.set_option(option, 1)
    # dense input (0.5 * one-hot is not one-hot): same network, generic first-layer kernels, checked against the oracle
    xd = 0.5 * x
    m = build(cfg, sd).train()
    _, logits, aux = m(xd.cuda())
    ref_loss, ref_logits, _, ref_grads = om.loss_and_grads(sd, cfg, xd, labels)
    assert (logits.cpu() - ref_logits).abs().max().item() <= LOGIT_ATOL
    loss = m.loss(logits, aux, labels.cuda())
    loss.backward()
    k = "msau_net.blocks.0.downsamplingblock.conv1s.0.conv.weight"
    g = dict(m.named_parameters())[k].grad.cpu()
    assert (g - ref_grads[k]).double().norm().item() <= 3e-2 * ref_grads[k].double().norm().item()


def test_full_size_page_properties():
    """512x512 chargrid page at the train-script config: finite outputs, soft-max rows sum to 1, loss decreases
    over a few fused steps (size-independent sanity at BASELINE.json's full page size)."""
    cfg = om.MsauConfig()
    sd = om.init_state_dict(cfg, 0)
    x, labels = synth_input(cfg.channels, cfg.n_class, 2, 512, 512, 5)
    m = build(cfg, sd).train()
    xc, lc = x.cuda(), labels.cuda()
    losses = [float(m.train_step(xc, lc, lr=1e-3)) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    m.eval()
    with torch.no_grad():
        probs, logits, aux = m(xc)
    assert torch.isfinite(logits).all() and torch.isfinite(aux).all()
    assert (probs.sum(1) - 1).abs().max().item() < 1e-5


def test_bert_grid_config():
    """BASELINE.json config 3: 768-channel BERT-grid input (dense, not one-hot): parity with the oracle on a small page,
    then one full-size batch (8 x 768 x 512 x 512) through two fused train steps."""
    cfg = om.MsauConfig(channels=768)
    sd = om.init_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(4)
    x = 0.3 * torch.randn(1, 768, 16, 24, generator=g)
    x = x * (torch.rand(1, 1, 16, 24, generator=g) < 0.4)          # boxes cover part of the page
    labels = torch.randint(0, cfg.n_class, (1, 16, 24), generator=g)
    labels[:, 0, 0] = 1
    m = build(cfg, sd).train()
    _, logits, aux = m(x.cuda())
    ref_loss, ref_logits, _, ref_grads = om.loss_and_grads(sd, cfg, x, labels)
    assert (logits.cpu() - ref_logits).abs().max().item() <= LOGIT_ATOL
    loss = m.loss(logits, aux, labels.cuda())
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-4 * max(1.0, float(ref_loss))
    loss.backward()
    k = "msau_net.blocks.0.downsamplingblock.conv1s.0.conv.weight"
    gk = dict(m.named_parameters())[k].grad.cpu()
    assert (gk - ref_grads[k]).double().norm().item() <= 3e-2 * ref_grads[k].double().norm().item()
    del m
    # full size
    m = build(cfg, sd).train()
    xg = 0.3 * torch.randn(8, 768, 512, 512, device="cuda")
    xg *= (torch.rand(8, 1, 512, 512, device="cuda") < 0.3)
    lg = torch.randint(0, cfg.n_class, (8, 512, 512), device="cuda")
    losses = [float(m.train_step(xg, lg, lr=1e-3)) for _ in range(3)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_high_resolution_inference():
    """BASELINE.json config 5: 1024 x 768 chargrid pages, inference only.  The class map of a batch equals the class maps of
    its chunks, agrees with the arg-max of the full forward, and (one page) with the oracle on >= 99.9 % of the pixels."""
    cfg = om.MsauConfig()
    sd = om.init_state_dict(cfg, 0)
    m = build(cfg, sd).eval()
    B, H, W = 6, 1024, 768
    g = torch.Generator(device="cuda").manual_seed(9)
    ids = torch.randint(0, cfg.channels, (B, 1, H, W), device="cuda", generator=g)
    occ = (torch.rand((B, 1, H, W), device="cuda", generator=g) < 0.1).float()
    x = torch.zeros(B, cfg.channels, H, W, device="cuda").scatter_(1, ids, occ)
    with torch.no_grad():
        cm = m.predict_classes(x)
        cm2 = m.predict_classes(x, pages_per_call=4)
        probs, logits, _ = m(x[:2])
    assert cm.dtype == torch.uint8 and cm.shape == (B, H, W) and int(cm.max()) < cfg.n_class
    assert torch.equal(cm, cm2)
    assert torch.equal(cm[:2].long(), logits.argmax(1))
    out, _ = om.msau_forward(sd, cfg, x[:1].cpu())
    agree = (out.argmax(1) == cm[:1].cpu().long()).float().mean().item()
    assert agree >= 0.999, agree


def test_inference_graph_replay_matches_eager(golden_dir):
    """predict_classes_graph (one CUDA-graph launch per call) returns the eager class map, follows new inputs through its
    static buffer and sees parameter updates made between replays."""
    z, meta, cfg = load(golden_dir, "model_s4r2_c96")
    sd = om.init_state_dict(cfg, meta["seed"])
    m = build(cfg, sd).eval()
    xs = [synth_input(cfg.channels, cfg.n_class, 1, 64, 48, s)[0].cuda() for s in (3, 4)]
    with torch.no_grad():
        for x in xs + xs[:1]:
            assert torch.equal(m.predict_classes_graph(x), m.predict_classes(x))
        assert len(m._graphs) == 1
        sd2 = om.init_state_dict(cfg, meta["seed"] + 1)
        m.load_state_dict(sd2)
        got = m.predict_classes_graph(xs[0])
        assert torch.equal(got, m.predict_classes(xs[0]))
        assert torch.equal(got, build(cfg, sd2).eval().predict_classes(xs[0]))


def test_graph_train_step_matches_eager(golden_dir):
    """train_step(use_graph=...) -- forward + loss + backward replayed from a CUDA graph (two streams inside) -- gives the eager
    step's losses and parameters, with private input copies (True) and bound to the caller's tensors ("static")."""
    z, meta, cfg = load(golden_dir, "model_s4r2_c96")
    sd = om.init_state_dict(cfg, meta["seed"])
    batches = [synth_input(cfg.channels, cfg.n_class, 2, 64, 48, s) for s in (5, 6, 7)]
    batches = [(x.cuda(), l.cuda()) for x, l in batches]
    me, mg, ms = (build(cfg, sd).train() for _ in range(3))
    xs, ls = batches[0][0].clone(), batches[0][1].clone()
    def compare_params():
        # atomics make the weight gradient's summation order (and the Adam sign of ~zero gradients) run-dependent: same
        # bounds as the eager-vs-eager comparison in test_param_grads_and_train_step_match_golden
        for (k, pe), pg, ps in zip(me.named_parameters(), mg.parameters(), ms.parameters()):
            tol = 2.5e-4 if k.endswith("attention_block.f.conv.bias") else 2e-6
            assert (pe.detach() - pg.detach()).abs().max().item() <= tol, k
            assert (pe.detach() - ps.detach()).abs().max().item() <= tol, k

    le, lg, lst = [], [], []
    n_eager = n_graph = 0
    for i, (x, l) in enumerate(batches):
        n0 = _lib.launch_count()
        le.append(float(me.train_step(x, l)))
        n1 = _lib.launch_count()
        lg.append(float(mg.train_step(x, l, use_graph=True)))
        n2 = _lib.launch_count()
        n_eager += n1 - n0
        n_graph += n2 - n1
        xs.copy_(x); ls.copy_(l)
        lst.append(float(ms.train_step(xs, ls, use_graph="static")))
        if i == 0:
            compare_params()
    assert len(mg._train_graphs) == 1 and len(ms._train_graphs) == 1
    # warm-up + capture enqueue the forward/backward kernels twice more than the eager loop; every replay counts like an eager step
    # (the graph holds the 2 optimiser launches as well: graph_tail)
    assert n_graph == n_eager + 2 * ((n_eager - 6) // 3) + 4
    for a, b, c in zip(le, lg, lst):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)) and abs(a - c) <= 1e-4 * max(1.0, abs(a))
    assert le[-1] != le[0]
    with pytest.raises(_lib.MsauError):
        ms.train_step(batches[1][0], batches[1][1], use_graph="static")


def test_s6r3_deep_levels(tc):
    """Wrapper-default depth (S=6, R=3) on a 256x256 page, where the 128- and 256-channel levels are 16x16 and 8x8 maps: wide
    enough for the tensor-core kernels, so the 256-column convs run as two 128-column launches (forward, data gradient and
    weight gradient).  Logits, loss and every parameter-gradient norm against the torch-fp32 oracle."""
    cfg = om.MsauConfig(channels=16, n_class=5, scale_space_num=6, res_depth=3, feat_root=8)
    sd = om.init_state_dict(cfg, 31)
    x, labels = synth_input(cfg.channels, cfg.n_class, 1, 256, 256, 32)
    m = build(cfg, sd).train()
    _, logits, aux = m(x.cuda())
    loss = m.loss(logits, aux, labels.cuda())
    loss.backward()
    ref_loss, ref_logits, ref_aux, ref_grads = om.loss_and_grads(sd, cfg, x, labels)
    assert (logits.cpu() - ref_logits).abs().max().item() <= LOGIT_ATOL
    assert (aux.cpu() - ref_aux).abs().max().item() <= LOGIT_ATOL
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-4 * max(1.0, float(ref_loss))
    named = dict(m.named_parameters())
    keys = [k for k, _ in om.param_schema(cfg)]
    got = np.array([0.0 if named[k].grad is None else float(named[k].grad.double().norm()) for k in keys])
    want = np.array([0.0 if ref_grads[k] is None else float(ref_grads[k].double().norm()) for k in keys])
    np.testing.assert_allclose(got, want, rtol=3e-2 if tc else 2e-3, atol=2e-5 if tc else 2e-6)
    for k in ("msau_net.blocks.0.downsamplingblock.conv_res_list.5.conv_res_list.1.custom_conv.weight",
              "msau_net.blocks.1.downsamplingblock.conv1s.5.conv.weight",
              "msau_net.blocks.1.upsamplingblock.deconvs.4.conv.weight"):
        d = (named[k].grad.cpu() - ref_grads[k]).double().norm().item() / ref_grads[k].double().norm().item()
        assert d <= (3e-2 if tc else 2e-3), (k, d)
