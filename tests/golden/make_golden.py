#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(datvo06/MSAU mounted read-only at /root/reference) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Nothing is copied from the reference; its modules are imported and called.  A 3-line stub for
``skimage.morphology.skeletonize`` (absent here, used only by the off-path ``skelet``) lets
``inference.morph_util`` import.

Inputs are regenerated at test time from the seeds recorded in each fixture (weights through
``oracle.model.init_state_dict`` -- a deterministic CPU torch.Generator stream -- and pages through
``oracle.raster.synth_page``), so the fixtures hold only the reference's OUTPUTS.
"""
import hashlib
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

_sk = types.ModuleType("skimage"); _skm = types.ModuleType("skimage.morphology")
_skm.skeletonize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
sys.modules["skimage"] = _sk; sys.modules["skimage.morphology"] = _skm

import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    from model.model import MSAUWrapper                      # noqa: E402
    import data_generator_funsd_bert as dgfb                 # noqa: E402
    from inference import morph_util                         # noqa: E402
    from inference.kv_model import KVModel                   # noqa: E402
    from inference.generic_util import to_categorical        # noqa: E402

from oracle import model as om                               # noqa: E402
from oracle import raster as orr                             # noqa: E402
from oracle.synth import synth_input as _synth_input, class_map, kv_pred_mask  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ----------------------------------------------------------------------------- model
def ref_model(cfg: om.MsauConfig, sd):
    with contextlib.redirect_stdout(io.StringIO()):
        m = MSAUWrapper(cfg.channels, cfg.n_class, dict(
            model="msau", final_act="softmax", featRoot=cfg.feat_root,
            scale_space_num=cfg.scale_space_num, res_depth=cfg.res_depth))
    assert [k for k in m.state_dict()] == [k for k, _ in om.param_schema(cfg)]
    assert [tuple(v.shape) for v in m.state_dict().values()] == [s for _, s in om.param_schema(cfg)]
    m.load_state_dict(sd)
    return m


def synth_input(cfg, B, H, W, seed):
    return _synth_input(cfg.channels, cfg.n_class, B, H, W, seed)


def golden_model(name, cfg, B, H, W, seed, full_grads):
    sd = om.init_state_dict(cfg, seed)
    x, labels = synth_input(cfg, B, H, W, seed + 1)
    m = ref_model(cfg, sd)
    m.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=0.0001)
    out = {}
    # --- per-page reference train-step semantics, accumulated the way DP averaging would (D6)
    m.zero_grad()
    losses = []
    logits_all, aux_all, probs_all = [], [], []
    for b in range(B):
        probs, logits, aux = m(x[b:b + 1])
        loss = m.loss(logits, aux, labels[b:b + 1])
        (loss / B).backward()
        losses.append(float(loss))
        logits_all.append(logits.detach()); aux_all.append(aux.detach()); probs_all.append(probs.detach())
    out["logits"] = torch.cat(logits_all).numpy()
    out["aux"] = torch.cat(aux_all).numpy()
    out["probs"] = torch.cat(probs_all).numpy()
    out["page_losses"] = np.array(losses, np.float64)
    keys = [k for k, _ in om.param_schema(cfg)]
    named = dict(m.named_parameters())
    out["grad_is_none"] = np.array([named[k].grad is None for k in keys])
    out["grad_norms"] = np.array([0.0 if named[k].grad is None else float(named[k].grad.double().norm()) for k in keys])
    for k in full_grads:
        out["grad::" + k] = named[k].grad.numpy().copy()
    total = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    out["total_norm"] = np.float64(total)
    opt.step()
    out["param_sums_after_step"] = np.array([float(named[k].detach().double().sum()) for k in keys])
    for k in full_grads:
        out["param_after::" + k] = named[k].detach().numpy().copy()
    meta = dict(cfg=cfg.__dict__, B=B, H=H, W=W, seed=seed)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", losses, "total_norm", float(total))


def synth_onehot_targets(n_class, B, H, W, seed):
    """one-hot main / aux target maps [B, n_class, H, W] int64 (what the reference's generators hand to Trainer.train)"""
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, n_class, (B, H, W), generator=g)
    t = t * (torch.rand((B, H, W), generator=g) < 0.5)          # half of the pixels are background (class 0)
    ta = torch.where(torch.rand((B, H, W), generator=g) < 0.9, t, torch.randint(0, n_class, (B, H, W), generator=g))
    return torch.nn.functional.one_hot(t, n_class).permute(0, 3, 1, 2).contiguous(), \
        torch.nn.functional.one_hot(ta, n_class).permute(0, 3, 1, 2).contiguous()


def golden_trainer(name, cfg, B, H, W, seed, optimizer, class_weights, steps=2):
    """The alternative trainer's step (model/training/trainer.py:122-137) with the reference's own UNetLoss and get_optimizer."""
    from model.training.cost import UNetLoss
    from model.training.optimizer import get_optimizer
    sd = om.init_state_dict(cfg, seed)
    x, _ = synth_input(cfg, B, H, W, seed + 1)
    tgt, tgt_aux = synth_onehot_targets(cfg.n_class, B, H, W, seed + 2)
    m = ref_model(cfg, sd)
    m.train()
    ck = {"aux_logits": None, "aux_tgt": None}
    if class_weights is not None:
        ck["class_weights"] = class_weights
    crit = UNetLoss(ck)
    with contextlib.redirect_stdout(io.StringIO()):
        opt = get_optimizer(m, {"optimizer": optimizer} if optimizer != "rmsprop" else {})
    out = {}
    accs, losses, finals = [], [], []
    for s in range(steps):
        opt.zero_grad()
        _, logits, aux = m(x)
        ck["aux_logits"] = aux
        ck["aux_tgt"] = tgt_aux
        acc, loss, final = crit(logits, tgt, ck)
        loss.backward()
        if s == 0:
            out["logits"] = logits.detach().numpy()
            keys = [k for k, _ in om.param_schema(cfg)]
            named = dict(m.named_parameters())
            out["grad_norms"] = np.array([0.0 if named[k].grad is None else float(named[k].grad.double().norm()) for k in keys])
        opt.step()
        accs.append(float(acc)); losses.append(float(loss)); finals.append(float(final))
    keys = [k for k, _ in om.param_schema(cfg)]
    named = dict(m.named_parameters())
    out["acc"] = np.array(accs); out["loss"] = np.array(losses); out["final_loss"] = np.array(finals)
    out["param_sums_after"] = np.array([float(named[k].detach().double().sum()) for k in keys])
    k0 = "msau_net.blocks.1.upsamplingblock.conv1s.0.custom_conv.weight"
    out["param_after::" + k0] = named[k0].detach().numpy().copy()
    meta = dict(cfg=cfg.__dict__, B=B, H=H, W=W, seed=seed, optimizer=optimizer, class_weights=class_weights, steps=steps)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "acc", accs, "loss", losses, "final", finals)


# ----------------------------------------------------------------------------- rasterisers
class _Cell:
    def __init__(self, x, y, w, h, ocr_value):
        self.x, self.y, self.w, self.h, self.ocr_value = x, y, w, h, ocr_value


class _DS:
    pass


def _ref_r1_r2(words, lines, D, feats768):
    eye = np.eye(D)
    cells_word = [_Cell(float(words["x"][i]), float(words["y"][i]), float(words["w"][i]), float(words["h"][i]),
                        "a" * len(words["chars"][i])) for i in range(len(words["x"]))]
    cells = [_Cell(float(lines["x"][i]), float(lines["y"][i]), float(lines["w"][i]), float(lines["h"][i]), "x")
             for i in range(len(lines["x"]))]
    ds = _DS()
    ds.inp_list = [dict(cells_word=cells_word, cells=cells,
                        charset_feature=[eye[c] for c in words["chars"]],
                        labels=np.asarray(lines["label"]), transformer_feature=feats768)]
    ds.getitem_box = dgfb.getitem_box_bert
    r1 = dgfb.get_box_mask_box_label_word(ds, 0)
    r2 = dgfb.get_box_mask_box_label(ds, 0)
    return r1, r2


def _ref_r3(boxes, texts, charset):
    kv = KVModel()
    kv.charset = " " + "$" + charset
    kv.tok_to_id = {t: i for i, t in enumerate(kv.charset)}
    kv.blank_idx = 1
    kv.n_token = len(kv.tok_to_id)
    lines = [dict(box=[int(v) for v in b], text=t, type=0, value=0) for b, t in zip(boxes, texts)]
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(dict(lines=lines), f)
    try:
        return kv, kv._generate_masks_from_label(f.name)
    finally:
        os.unlink(f.name)


CHARSET = "".join(chr(c) for c in range(33, 127) if chr(c) != "$") + chr(161)   # 94 chars, SURVEY 8


def texts_for(words):
    return ["".join(CHARSET[c - 2] for c in ch) for ch in words["chars"]]


def golden_raster():
    out = {}
    meta = []
    for tag, (gh, gw, n, seed) in dict(small=(48, 40, 24, 3), odd=(37, 53, 30, 4), full0=(512, 512, 198, 0),
                                       full1=(512, 512, 198, 1)).items():
        words, lines = orr.synth_page(seed, gh, gw, n)
        if tag == "odd":   # some empty-text words + a degenerate tiny box exercise the fallbacks
            words["chars"][3] = np.zeros(0, np.int32)
            words["chars"][7] = np.zeros(0, np.int32)
        D = 96
        feats = np.random.RandomState(seed + 100).randn(len(lines["x"]), 24 if tag != "small" else 768)
        r1, r2 = _ref_r1_r2(words, lines, D, feats)
        g1 = torch.Tensor(r1["mask"]).numpy()      # the fp32 the network sees (dgfb.py:220)
        g2 = torch.Tensor(r2["mask"]).numpy()
        assert g1.shape[1:] == (gh, gw), (g1.shape, gh, gw)
        out[f"{tag}::r1_sha"] = np.array(sha(g1)); out[f"{tag}::r1_label_sha"] = np.array(sha(r1["label"]))
        out[f"{tag}::r2_sha"] = np.array(sha(g2)); out[f"{tag}::r2_label_sha"] = np.array(sha(r2["label"]))
        out[f"{tag}::r1_ids"] = (g1.argmax(0) * (g1.max(0) > 0)).astype(np.uint8)
        out[f"{tag}::r1_label"] = r1["label"]; out[f"{tag}::r2_label"] = r2["label"]
        if tag in ("small", "odd"):
            out[f"{tag}::r2_grid"] = g2.astype(np.float32)
        # R3 on the same boxes: [x, y, x+w, y+h], digits folded by the reference itself
        boxes = np.stack([words["x"], words["y"], words["x"] + words["w"], words["y"] + words["h"]], 1)
        texts = texts_for(words)
        kv, (im, lm, cm, label_lines, scale, bg_pad, bbox) = _ref_r3(boxes, texts, CHARSET)
        for nm, a in (("input", im), ("line", lm), ("char", cm)):
            out[f"{tag}::r3_{nm}_sha"] = np.array(sha(a))
            if tag in ("small", "odd"):
                out[f"{tag}::r3_{nm}"] = a
        out[f"{tag}::r3_shape"] = np.array(im.shape)
        out[f"{tag}::r3_boxes"] = np.array([l["box"] for l in label_lines], np.int64)
        out[f"{tag}::r3_scale_pad"] = np.array([scale, bg_pad], np.float64)
        if tag == "small":
            oh = to_categorical(im, kv.n_token)
            bx = torch.from_numpy(np.expand_dims(oh, 0)).transpose(1, -1).transpose(2, 3).float().numpy()
            out["small::r3_onehot_sha"] = np.array(sha(bx)); out["small::r3_n_token"] = np.array(kv.n_token)
        meta.append(dict(tag=tag, gh=gh, gw=gw, n_words=n, seed=seed, feat_dim=int(feats.shape[1])))
    out["meta"] = np.array(json.dumps(dict(pages=meta, charset=CHARSET)))
    np.savez_compressed(os.path.join(HERE, "raster.npz"), **out)
    print("raster ok")


# ----------------------------------------------------------------------------- morphology
def golden_morph():
    out = {}
    meta = []
    for tag, (H, W, seed) in dict(small=(40, 56, 5), odd=(33, 47, 6), full=(512, 512, 7)).items():
        m = class_map(seed, H, W)
        for c in range(2, 5):
            closed = morph_util.r_closing(m == c, (1, 3))
            labels, objs = morph_util.connected_components(closed)
            assert labels.dtype == np.int32
            bb = np.array([[o[0].start, o[0].stop, o[1].start, o[1].stop] for o in objs], np.int32).reshape(-1, 4)
            out[f"{tag}::{c}::closed"] = np.packbits(closed)
            out[f"{tag}::{c}::labels_sha"] = np.array(sha(labels))
            out[f"{tag}::{c}::bboxes"] = bb
            if tag != "full":
                out[f"{tag}::{c}::labels"] = labels
        for nm, fn in (("dil", morph_util.r_dilation), ("ero", morph_util.r_erosion), ("open", morph_util.r_opening)):
            for size, origin in (((3, 3), 0), ((2, 5), 0), ((1, 4), 0), ((3, 3), (1, -1)), ((4, 2), (-2, 0))):
                r = fn((m > 2), size, origin=origin)
                out[f"{tag}::{nm}::{size[0]}x{size[1]}::{origin}"] = np.packbits(r)
        meta.append(dict(tag=tag, H=H, W=W, seed=seed))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "morph.npz"), **out)
    print("morph ok")


# ----------------------------------------------------------------------------- _extract_value
def _plain(v):
    """values tuple -> JSON-able (numpy ints from slices / np.unique -> int)."""
    if v is None:
        return None
    if isinstance(v, str):
        return v
    if isinstance(v, (list, tuple)):
        return [_plain(e) for e in v]
    return int(v)


def golden_kv():
    """KVModel._extract_value (inference/kv_model.py:151-261) of the unmodified reference on the R3 masks of two synthetic
    pages and a seeded synthetic soft-max map (oracle.synth.kv_pred_mask), with 5 classes and with 7 (class 5 is one of the
    reference's multiple_lines_fields)."""
    out, meta = {}, []
    for tag, (gh, gw, n, seed) in dict(small=(48, 40, 24, 3), odd=(37, 53, 30, 4)).items():
        words, _ = orr.synth_page(seed, gh, gw, n)
        boxes = np.stack([words["x"], words["y"], words["x"] + words["w"], words["y"] + words["h"]], 1)
        texts = texts_for(words)
        for n_class, noise in ((5, 0.01), (7, 0.01), (7, 0.0)):
            kv, (im, lm, cm, label_lines, scale, bg_pad, bbox) = _ref_r3(boxes, texts, CHARSET)
            pm = kv_pred_mask(seed * 10 + n_class, im.shape, [l["box"] for l in label_lines], n_class, noise)
            values, new_mask = KVModel._extract_value(lm, cm, label_lines, pm, n_class)
            key = f"{tag}::{n_class}::{noise}"
            out[key + "::values"] = np.array(json.dumps([_plain(v) for v in values]))
            out[key + "::new_mask_sha"] = np.array(sha(new_mask))
            out[key + "::new_mask_fg"] = np.packbits(new_mask[:, :, 1:] > 0)
            out[key + "::shape"] = np.array(new_mask.shape)
            meta.append(dict(key=key, gh=gh, gw=gw, n_words=n, seed=seed, n_class=n_class, noise=noise, pred_seed=seed * 10 + n_class))
            print(key, [v[0] for v in values])
    out["meta"] = np.array(json.dumps(dict(cases=meta, charset=CHARSET)))
    np.savez_compressed(os.path.join(HERE, "kv_extract.npz"), **out)
    print("kv ok")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])          # e.g. `make_golden.py model_s6r3_c16` regenerates one fixture

    def want(name):
        return not only or name in only

    if want("raster"):
        golden_raster()
    if want("morph"):
        golden_morph()
    if want("kv"):
        golden_kv()
    # alternative trainer (UNetLoss + get_optimizer); 17 classes = the KV checkpoints' class count (inference/postprocess.py:2-5)
    kv17 = om.MsauConfig(channels=12, n_class=17, scale_space_num=3, res_depth=2, feat_root=8)
    if want("trainer_rmsprop_c17"):
        golden_trainer("trainer_rmsprop_c17", kv17, B=2, H=24, W=40, seed=31, optimizer="rmsprop", class_weights=None)
    if want("trainer_momentum_w_c17"):
        golden_trainer("trainer_momentum_w_c17", kv17, B=2, H=24, W=40, seed=33, optimizer="momentum",
                       class_weights=[0.5] + [1.0 + 0.1 * i for i in range(16)])
    if want("model_s3r2_c12_k17"):
        golden_model("model_s3r2_c12_k17", kv17, B=2, H=37, W=43, seed=13,
                     full_grads=["msau_net.end_convs.2.custom_conv.weight", "msau_net.blocks.1.downsamplingblock.conv1s.0.conv.weight"])
    small = om.MsauConfig(channels=12, n_class=5, scale_space_num=3, res_depth=2, feat_root=8)
    if want("model_s3r2_c12"):
        golden_model("model_s3r2_c12", small, B=2, H=37, W=43, seed=11,
                     full_grads=["msau_net.blocks.0.downsamplingblock.conv1s.0.conv.weight",
                                 "msau_net.blocks.1.downsamplingblock.layer_attentions.attention_block.g.conv.weight",
                                 "msau_net.blocks.1.upsamplingblock.deconvs.0.conv.weight",
                                 "msau_net.end_convs.2.custom_conv.bias"])
    train_cfg = om.MsauConfig(channels=96, n_class=5, scale_space_num=4, res_depth=2, feat_root=8)
    if want("model_s4r2_c96"):
        golden_model("model_s4r2_c96", train_cfg, B=1, H=64, W=48, seed=0,
                     full_grads=["msau_net.blocks.2.upsamplingblock.conv1_1s.0.custom_conv.weight"])
    deep = om.MsauConfig(channels=8, n_class=3, scale_space_num=2, res_depth=3, feat_root=16)
    if want("model_s2r3_c8"):
        golden_model("model_s2r3_c8", deep, B=3, H=16, W=24, seed=5, full_grads=[])
    # the wrapper's own defaults (model/model.py:406-408): S=6, R=3, featRoot=8 -> levels of 8..256 channels, attention at 256
    dflt = om.MsauConfig(channels=16, n_class=5, scale_space_num=6, res_depth=3, feat_root=8)
    if want("model_s6r3_c16"):
        golden_model("model_s6r3_c16", dflt, B=1, H=64, W=160, seed=21,
                     full_grads=["msau_net.blocks.0.downsamplingblock.layer_attentions.attention_block.f.conv.weight"])
