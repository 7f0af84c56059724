"""Bit-exact parity of the rasterisers / morphology / CCL kernels (through the C ABI) with the numpy oracle
and the golden fixtures generated from the reference."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from msau_b200 import morph, raster
from oracle import morph as omo
from oracle import raster as orr
from oracle.synth import class_map

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def page(mp):
    words, lines = orr.synth_page(mp["seed"], mp["gh"], mp["gw"], mp["n_words"])
    if mp["tag"] == "odd":
        words["chars"][3] = np.zeros(0, np.int32)
        words["chars"][7] = np.zeros(0, np.int32)
    return words, lines


@pytest.fixture(scope="module")
def rgold(golden_dir):
    z = np.load(os.path.join(golden_dir, "raster.npz"))
    return z, json.loads(str(z["meta"]))


def kv_inputs(words, charset):
    boxes = np.stack([words["x"], words["y"], words["x"] + words["w"], words["y"] + words["h"]], 1)
    tok = {t: i for i, t in enumerate(" $" + charset)}
    ids = []
    for ch in words["chars"]:
        text = "".join(charset[c - 2] for c in ch)
        text = "".join(c if not c.isdigit() else "0" for c in text)
        ids.append(np.array([tok.get(c, 1) for c in text], np.int32))
    return boxes, ids


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_rasterisers_match_golden(rgold, idx):
    z, meta = rgold
    mp = meta["pages"][idx]
    tag = mp["tag"]
    words, lines = page(mp)
    grid, label, geom = raster.rasterize_word_chargrid([words], [lines], np.eye(96))
    g = geom.cpu().numpy()[0]
    assert (int(g[5]), int(g[6])) == (mp["gh"], mp["gw"])
    assert sha(grid[0].cpu().numpy()) == str(z[f"{tag}::r1_sha"])
    assert sha(label[0].cpu().numpy()) == str(z[f"{tag}::r1_label_sha"])
    # channels-last output holds the same values
    grid_l, _, _ = raster.rasterize_word_chargrid([words], [lines], np.eye(96), layout="nhwc")
    assert torch.equal(grid_l[0].permute(2, 0, 1)[:96], grid[0])
    feats = np.random.RandomState(mp["seed"] + 100).randn(len(lines["x"]), mp["feat_dim"])
    grid2, label2, _ = raster.rasterize_box_grid([lines], [feats])
    assert sha(grid2[0].cpu().numpy()) == str(z[f"{tag}::r2_sha"])
    assert sha(label2[0].cpu().numpy()) == str(z[f"{tag}::r2_label_sha"])
    boxes, ids = kv_inputs(words, meta["charset"])
    r3 = raster.rasterize_kv([boxes], [ids])
    assert tuple(z[f"{tag}::r3_shape"]) == tuple(r3["input_mask"].shape[1:])
    for nm, key in (("input", "input_mask"), ("line", "line_id_mask"), ("char", "character_id_mask")):
        assert sha(r3[key][0].cpu().numpy().view(np.uint16)) == str(z[f"{tag}::r3_{nm}_sha"]), nm
    assert (r3["scaled_boxes"].cpu().numpy() == z[f"{tag}::r3_boxes"]).all()
    g3 = r3["geom3"].cpu().numpy()[0]
    assert g3[2] == z[f"{tag}::r3_scale_pad"][0] and g3[3] == z[f"{tag}::r3_scale_pad"][1]
    if tag == "small":
        oh = raster.one_hot(r3["input_mask"], int(z["small::r3_n_token"]))
        assert sha(oh.cpu().numpy()) == str(z["small::r3_onehot_sha"])


def test_raster_batch_matches_oracle():
    """A ragged batch (different page sizes, empty words) against the numpy oracle, page by page."""
    specs = [(3, 48, 40, 24), (9, 37, 53, 30), (12, 64, 64, 60), (13, 20, 33, 5)]
    wp, lp, fe = [], [], []
    for seed, gh, gw, n in specs:
        w, l = orr.synth_page(seed, gh, gw, n)
        if seed == 9:
            w["chars"][2] = np.zeros(0, np.int32)
        wp.append(w); lp.append(l)
        fe.append(np.random.RandomState(seed).randn(len(l["x"]), 20))
    grid, label, geom = raster.rasterize_word_chargrid(wp, lp, np.eye(96), out_hw=(64, 64))
    grid2, label2, _ = raster.rasterize_box_grid(lp, fe, out_hw=(64, 64))
    boxes_ids = [kv_inputs(w, "".join(chr(c) for c in range(33, 127) if chr(c) != "$") + chr(161)) for w in wp]
    r3 = raster.rasterize_kv([b for b, _ in boxes_ids], [i for _, i in boxes_ids])
    for p, (seed, gh, gw, n) in enumerate(specs):
        og, ol = orr.raster_word_chargrid(wp[p], lp[p], np.eye(96))
        want = np.zeros((96, 64, 64), np.float32); want[:, :gh, :gw] = og
        assert np.array_equal(grid[p].cpu().numpy(), want)
        wl = np.zeros((64, 64), np.uint8); wl[:gh, :gw] = ol
        assert np.array_equal(label[p].cpu().numpy(), wl)
        og2, ol2 = orr.raster_box_grid(lp[p], fe[p])
        want2 = np.zeros((20, 64, 64), np.float32); want2[:, :gh, :gw] = torch.Tensor(og2).numpy()
        assert np.array_equal(grid2[p].cpu().numpy(), want2)
        o3 = orr.raster_kv_chargrid(*boxes_ids[p])
        h3, w3 = o3["input_mask"].shape
        for key in ("input_mask", "line_id_mask", "character_id_mask"):
            got = r3[key][p].cpu().numpy().view(np.uint16)
            assert np.array_equal(got[:h3, :w3], o3[key]), (p, key)
            assert got[h3:].sum() == 0 and got[:, w3:].sum() == 0


def test_id_map_path_equals_dense_path():
    """The rasteriser's id map (layout "ids") is the arg-max of its dense one-hot grid, and a train step fed with the id map
    (x_layout 2: structured first layer, no dense tensor) gives the same loss and parameters as the dense-grid step."""
    import msau_b200
    from oracle import model as om
    wp, lp = [], []
    for seed, n in ((3, 60), (9, 45)):
        w, l = orr.synth_page(seed, 64, 64, n)
        wp.append(w); lp.append(l)
    table = torch.eye(96, dtype=torch.float64, device="cuda")
    words = raster.BoxBatch(wp, "cuda", with_chars=True)
    lines = raster.BoxBatch(lp, "cuda", with_chars=False, with_labels=True)
    geom = words.geometry()
    dense = raster.raster_features(words, geom, table, (64, 64), True, "nchw")
    ids = raster.raster_features(words, geom, table, (64, 64), True, "ids")
    label = raster.raster_labels(lines, geom, (64, 64))
    want = torch.where(dense.sum(1) > 0, dense.argmax(1), torch.full_like(dense.argmax(1), -1))
    assert ids.dtype == torch.int16 and torch.equal(ids.long(), want)
    cfg = om.MsauConfig()
    sd = om.init_state_dict(cfg, 2)
    kw = dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)
    ma = msau_b200.MSAUWrapper(96, 5, kw); ma.load_state_dict(sd); ma = ma.cuda().train()
    mb = msau_b200.MSAUWrapper(96, 5, kw); mb.load_state_dict(sd); mb = mb.cuda().train()
    la = float(ma.train_step(dense, label.long()))
    lb_ = float(mb.train_step(ids, label.long(), layout=2))
    assert abs(la - lb_) <= 1e-6 * max(1.0, abs(la))
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        tol = 2.5e-4 if k.endswith("attention_block.f.conv.bias") else 2e-6
        assert (pa.detach() - pb.detach()).abs().max().item() <= tol, k


def test_bert_row_id_path_equals_dense_path():
    """BERT grid (R2, data_generator_funsd_bert.py:64-93) as (row-id map, feature table), x_layout 3: the id map is the owner
    map of the dense 768-channel grid, forward / loss / gradients equal the dense-grid run of the fp32 kernels and the oracle,
    and a fused train step moves the parameters the same way."""
    import msau_b200
    from msau_b200 import _lib
    from oracle import model as om
    D = 768
    lp, fe = [], []
    for seed, n in ((5, 40), (11, 25)):
        _, l = orr.synth_page(seed, 64, 64, n)
        lp.append(l)
        fe.append(0.3 * np.random.RandomState(seed).randn(len(l["x"]), D))
    cells = raster.BoxBatch(lp, "cuda", with_chars=False, with_labels=True)
    geom = cells.geometry()
    table = torch.from_numpy(np.concatenate(fe)).cuda()
    dense = raster.raster_features(cells, geom, table, (64, 64), False, "nchw")
    ids = raster.raster_features(cells, geom, table, (64, 64), False, "ids")
    label = raster.raster_labels(cells, geom, (64, 64)).long()
    t32 = table.float()
    want = torch.zeros_like(dense)
    sel = ids.long().clamp(min=0)
    want = (t32[sel] * (ids >= 0).unsqueeze(-1)).permute(0, 3, 1, 2)
    assert ids.dtype == torch.int16 and (ids >= 0).any() and torch.equal(want.contiguous(), dense)
    cfg = om.MsauConfig(channels=D)
    sd = om.init_state_dict(cfg, 6)
    kw = dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)
    ref_loss, ref_logits, _, ref_grads = om.loss_and_grads(sd, cfg, dense.cpu(), label.cpu())
    mb = msau_b200.MSAUWrapper(D, 5, kw); mb.load_state_dict(sd); mb = mb.cuda().train()
    with pytest.raises(_lib.MsauError):
        mb.train_step(ids, label, layout=3)            # no table yet
    mb.set_feature_table(table)
    pl, xc, logits, aux, _, _ = mb._run_forward(ids, 3, True, False)
    assert (logits.cpu() - ref_logits).abs().max().item() <= 5e-4
    mb._last = (pl, xc, 3, 0, 0)
    loss = mb._backward_from_last(label)
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * max(1.0, float(ref_loss))
    gb = mb.flat_grads.clone()
    ma = msau_b200.MSAUWrapper(D, 5, kw); ma.load_state_dict(sd); ma = ma.cuda().train()
    _, la_logits, la_aux = ma(dense)
    ma.loss(la_logits, la_aux, label).backward()
    ga = ma.flat_grads
    assert (la_logits - logits).abs().max().item() <= 5e-4
    assert (ga - gb).double().norm().item() <= 3e-2 * ga.double().norm().item()
    off = 0
    for k, p in mb.named_parameters():
        n = p.numel()
        if k.startswith("msau_net.blocks.0.downsamplingblock.conv1s.0.conv."):
            g = gb[off:off + n].view(p.shape).cpu()
            assert (g - ref_grads[k]).double().norm().item() <= 3e-2 * ref_grads[k].double().norm().item(), k
        off += n
    l0 = float(mb.train_step(ids, label, layout=3, lr=1e-3))
    l1 = float(mb.train_step(ids, label, layout=3, lr=1e-3))
    l2 = float(mb.train_step(ids, label, layout=3, lr=1e-3))
    assert abs(l0 - float(ref_loss)) <= 1e-4 * max(1.0, float(ref_loss)) and l2 < l0 and np.isfinite(l1)


def test_bert_row_id_path_edge_cases():
    """An empty page (every id -1) next to a page covered by ONE box: the empty page contributes the bias-only response, the
    first-layer weight gradient equals that of the dense run, and out-of-range table sizes are refused."""
    import msau_b200
    from oracle import model as om
    D, H, W = 768, 32, 32
    table = 0.3 * torch.randn(3, D, generator=torch.Generator().manual_seed(1)).cuda()
    ids = torch.full((2, H, W), -1, dtype=torch.int16, device="cuda")
    ids[1, 4:20, 3:29] = 2
    dense = torch.zeros(2, D, H, W, device="cuda")
    dense[1, :, 4:20, 3:29] = table[2][:, None, None]
    labels = torch.randint(0, 5, (2, H, W), generator=torch.Generator().manual_seed(2)).cuda()
    labels[:, 0, 0] = 1
    cfg = om.MsauConfig(channels=D)
    sd = om.init_state_dict(cfg, 8)
    kw = dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)
    ma = msau_b200.MSAUWrapper(D, 5, kw); ma.load_state_dict(sd); ma = ma.cuda().train()
    mb = msau_b200.MSAUWrapper(D, 5, kw); mb.load_state_dict(sd); mb = mb.cuda().train()
    mb.set_feature_table(table)
    pl, xc, lg_b, _, _, _ = mb._run_forward(ids, 3, True, False)
    mb._last = (pl, xc, 3, 0, 0)
    loss_b = mb._backward_from_last(labels)
    _, lg_a, aux_a = ma(dense)
    loss_a = ma.loss(lg_a, aux_a, labels)
    loss_a.backward()
    assert (lg_a - lg_b).abs().max().item() <= 5e-4 and abs(float(loss_a.detach()) - float(loss_b)) <= 1e-4
    k = "msau_net.blocks.0.downsamplingblock.conv1s.0.conv.weight"
    ga = dict(ma.named_parameters())[k].grad
    n = ga.numel()
    gb = mb.flat_grads[:0]
    off = 0
    for kk, p in mb.named_parameters():
        if kk == k:
            gb = mb.flat_grads[off:off + p.numel()].view(p.shape)
        off += p.numel()
    assert gb.numel() == n and (ga - gb).double().norm().item() <= 3e-2 * ga.double().norm().item()
    with pytest.raises(ValueError):
        mb.set_feature_table(torch.zeros(40000, D))
    with pytest.raises(ValueError):
        mb.set_feature_table(torch.zeros(3, D + 1))


@pytest.fixture(scope="module")
def mgold(golden_dir):
    z = np.load(os.path.join(golden_dir, "morph.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_closing_ccl_match_golden(mgold, idx):
    z, meta = mgold
    mp = meta[idx]
    tag, H, W = mp["tag"], mp["H"], mp["W"]
    m = class_map(mp["seed"], H, W)
    for c in range(2, 5):
        closed = morph.r_closing(m == c, (1, 3))
        want = np.unpackbits(z[f"{tag}::{c}::closed"])[:H * W].reshape(H, W).astype(bool)
        assert closed.dtype == np.bool_ and (closed == want).all()
        labels, objs = morph.connected_components(closed)
        assert labels.dtype == np.int32
        assert sha(labels) == str(z[f"{tag}::{c}::labels_sha"])
        bb = np.array([[o[0].start, o[0].stop, o[1].start, o[1].stop] for o in objs], np.int32).reshape(-1, 4)
        assert (bb == z[f"{tag}::{c}::bboxes"]).all()
    fns = dict(dil=morph.r_dilation, ero=morph.r_erosion, open=morph.r_opening)
    n = 0
    for key in z.files:
        parts = key.split("::")
        if parts[0] != tag or parts[1] not in fns:
            continue
        sh_, sw_ = (int(v) for v in parts[2].split("x"))
        got = fns[parts[1]](m > 2, (sh_, sw_), eval(parts[3]))
        want = np.unpackbits(z[key])[:H * W].reshape(H, W).astype(bool)
        assert (got == want).all(), key
        n += 1
    assert n == 15


def test_ccl_batch_against_scipy_full_size():
    """256 x 3 class maps at 512x512 (BASELINE.json config 4): every label map equals scipy.ndimage.label's,
    plus the size-independent properties (labels are 1..n, bbox tight)."""
    from scipy import ndimage
    maps = np.stack([class_map(100 + i, 512, 512) for i in range(256)])
    t = torch.from_numpy(maps).cuda()
    for c in (2, 3, 4):
        closed = morph.closing_batch(morph.class_equals(t, c), (1, 3))
        # the one-pass (class == c) + closing kernel the inference driver uses gives the same map, also for even windows
        assert torch.equal(closed, morph.class_closing_batch(t, c, (1, 3)))
        if c == 2:
            for sw in (1, 2, 4, 5, 8):
                assert torch.equal(morph.closing_batch(morph.class_equals(t[:8], c), (1, sw)), morph.class_closing_batch(t[:8], c, (1, sw)))
        labels, n_labels, bboxes = morph.ccl_batch(closed)
        lab = labels.cpu().numpy(); nl = n_labels.cpu().numpy(); bb = bboxes.cpu().numpy(); cl = closed.cpu().numpy()
        for i in range(maps.shape[0]):
            want_closed = ndimage.minimum_filter(ndimage.maximum_filter(maps[i] == c, (1, 3), mode="constant"), (1, 3), mode="constant")
            assert (cl[i].astype(bool) == want_closed).all()
            want, n = ndimage.label(want_closed)
            assert nl[i] == n and (lab[i] == want).all()
            objs = ndimage.find_objects(want)
            got = morph.objects_from_bboxes(n, bb[i])
            assert got == objs


def test_degenerate_maps():
    for img in (np.zeros((7, 9), bool), np.ones((7, 9), bool), np.eye(8, dtype=bool)):
        labels, objs = morph.connected_components(img)
        wl, wo = omo.connected_components(img)
        assert (labels == wl).all() and objs == wo
        assert (morph.r_closing(img, (1, 3)) == omo.r_closing(img, (1, 3))).all()
        assert (morph.r_erosion(img, (3, 3)) == omo.r_erosion(img, (3, 3))).all()
