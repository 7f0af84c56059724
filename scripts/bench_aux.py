"""Secondary measurements (BASELINE.json configs 0, 2, 3, 4): one JSON object on stdout.

  c1  chargrid inference, 1 page 512x512                      -> ms / page (latency), pages/s at batch 16
      + KVModel.predict end to end on one page (reference inference API)
  c3  BERT-grid (768-channel dense input) train step, batch 8  -> pages/s
  s6r3  the wrapper-default model (S=6, R=3: levels of 8..256 channels), chargrid train step, batch 16 -> pages/s
  c4  R1 rasterisation of 256 pages (~200 boxes) + closing(1,3) + 4-connected labelling of 3 class maps per page
      -> pages/s and achieved HBM GB/s of the dense-grid write
  c5  1024x768 chargrid inference, 64 pages (chunked)          -> pages/s
CUDA events, 3 warm-up + 5 timed repetitions each, inputs resident in HBM (c4: page records on the host, H2D inside)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import msau_b200
from msau_b200 import morph, raster
from oracle import model as om
from oracle import raster as orr


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def onehot_pages(B, C, H, W, seed, occ=0.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ids = torch.randint(0, C, (B, 1, H, W), device="cuda", generator=g)
    o = (torch.rand((B, 1, H, W), device="cuda", generator=g) < occ).float()
    return torch.zeros(B, C, H, W, device="cuda").scatter_(1, ids, o)


def main():
    out = {}
    cfg = om.MsauConfig()
    kw = dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)
    m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, kw)
    m.load_state_dict(om.init_state_dict(cfg, 0))
    m = m.cuda().eval()
    with torch.no_grad():
        x1 = onehot_pages(1, 96, 512, 512, 1)
        ms1 = timed(lambda: m.predict_classes(x1))
        ms1g = timed(lambda: m.predict_classes_graph(x1), reps=20)
        x16 = onehot_pages(16, 96, 512, 512, 2)
        ms16 = timed(lambda: m.predict_classes(x16))
        out["c1_inference_512"] = dict(ms_per_page_batch1=ms1, ms_per_page_batch1_cuda_graph=ms1g, pages_per_s_batch16=16 / (ms16 * 1e-3))
        x64 = onehot_pages(64, 96, 1024, 768, 3)
        ms64 = timed(lambda: m.predict_classes(x64), reps=3, warm=1)
        out["c5_inference_1024x768_b64"] = dict(ms=ms64, pages_per_s=64 / (ms64 * 1e-3))
        del x64, x16, x1
    torch.cuda.empty_cache()
    # c1 through the reference's inference API: KVModel.predict on one synthetic FUNSD-shaped page (198 text lines): JSON lines ->
    # R3 masks -> one-hot -> network -> arg-max -> closing + labelling per class -> _extract_value -> field strings
    from msau_b200 import kv_model
    charset = "".join(chr(c) for c in range(33, 127) if chr(c) != "$") + chr(161)
    words, _ = orr.synth_page(11, 512, 512, 198)
    texts = ["".join(charset[c - 2] for c in ch) for ch in words["chars"]]
    page = dict(lines=[dict(box=[int(words["x"][i]), int(words["y"][i]), int(words["x"][i] + words["w"][i]),
                                 int(words["y"][i] + words["h"][i])], text=texts[i], type=0, value=0) for i in range(len(texts))])
    kv = kv_model.KVModel()
    kv.net = m
    kv.load(model_weight=None, charset=None, n_class=cfg.n_class)
    kv.set_charset(charset)
    import copy

    def kv_predict():
        return kv.predict((copy.deepcopy(page), None))

    with torch.no_grad():
        ms_kv = timed(kv_predict, reps=10)
        im = kv._generate_masks_from_label(copy.deepcopy(page), as_numpy=False)[0]
    out["c1_kv_predict_one_page"] = dict(ms=ms_kv, grid=list(im.shape), lines=len(texts),
                                         note="KVModel.predict: JSON lines -> R3 masks -> network -> closing / labelling -> field values, host syncs included")
    # c3
    cfg3 = om.MsauConfig(channels=768)
    m3 = msau_b200.MSAUWrapper(768, 5, kw)
    m3.load_state_dict(om.init_state_dict(cfg3, 3))
    m3 = m3.cuda().train()
    xg = 0.3 * torch.randn(8, 768, 512, 512, device="cuda")
    xg *= (torch.rand(8, 1, 512, 512, device="cuda") < 0.3)
    lg = torch.randint(0, 5, (8, 512, 512), device="cuda")
    ms3 = timed(lambda: m3.train_step(xg, lg))
    out["c3_bert_grid_train_b8"] = dict(ms_per_step=ms3, pages_per_s=8 / (ms3 * 1e-3))
    del xg, lg
    torch.cuda.empty_cache()
    # c3 as (row-id map, feature table): R2 page records -> id map on the device -> structured first layer (x_layout 3)
    cp, fe = [], []
    for i in range(8):
        _, l = orr.synth_page(500 + i, 512, 512, 198)
        cp.append(l)
        fe.append(0.3 * np.random.RandomState(500 + i).randn(len(l["x"]), 768))
    host = raster.HostBatch(cp, with_chars=False, with_labels=True)
    feats = torch.from_numpy(np.concatenate(fe)).pin_memory()

    def bert_step():
        cells = raster.BoxBatch.from_host(host, "cuda")
        geom = cells.geometry()
        table = feats.to("cuda", non_blocking=True)
        ids = raster.raster_features(cells, geom, table, (512, 512), False, "ids")
        lab = raster.raster_labels(cells, geom, (512, 512))
        m3.set_feature_table(table)
        return m3.train_step(ids, lab, layout=3)

    ms3t = timed(bert_step)
    out["c3_bert_grid_train_b8_row_ids"] = dict(ms_per_step=ms3t, pages_per_s=8 / (ms3t * 1e-3), h2d_bytes=host.nbytes + feats.numel() * 8,
                                                note="host R2 page records + fp64 feature table -> H2D -> int16 row-id map -> train step")
    del m3
    torch.cuda.empty_cache()
    # wrapper-default model (model/model.py:406-408: S=6, R=3, featRoot=8; 13 M parameters), chargrid train step, batch 16
    cfg6 = om.MsauConfig(channels=96, n_class=5, scale_space_num=6, res_depth=3, feat_root=8)
    m6 = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax"))          # the reference's own defaults
    m6.load_state_dict(om.init_state_dict(cfg6, 6))
    m6 = m6.cuda().train()
    x6 = onehot_pages(16, 96, 512, 512, 7)
    l6 = torch.randint(0, 5, (16, 512, 512), device="cuda")
    ms6 = timed(lambda: m6.train_step(x6, l6))
    out["s6r3_default_model_train_b16"] = dict(ms_per_step=ms6, pages_per_s=16 / (ms6 * 1e-3), params=int(m6.flat_params.numel()))
    del m6, x6, l6
    torch.cuda.empty_cache()
    # c4
    wp, lp = [], []
    for i in range(256):
        w, l = orr.synth_page(i, 512, 512, 198)
        wp.append(w); lp.append(l)
    table = torch.eye(96, dtype=torch.float64, device="cuda")

    def raster_dense():
        return raster.rasterize_word_chargrid(wp, lp, table, out_hw=(512, 512), layout="nhwc")

    ms_r = timed(raster_dense, reps=3, warm=1)
    grid_bytes = 256 * 512 * 512 * 96 * 4
    out["c4_raster_R1_256_pages"] = dict(ms=ms_r, pages_per_s=256 / (ms_r * 1e-3), grid_write_gbs=grid_bytes / (ms_r * 1e-3) / 1e9,
                                         note="host page records -> H2D -> geometry + owner + dense fp32 NHWC grid + label map")

    def raster_ids():
        words = raster.BoxBatch(wp, "cuda", with_chars=True)
        lines = raster.BoxBatch(lp, "cuda", with_chars=False, with_labels=True)
        geom = words.geometry()
        return raster.raster_features(words, geom, table, (512, 512), True, "ids"), raster.raster_labels(lines, geom, (512, 512))

    ms_i = timed(raster_ids, reps=3, warm=1)
    out["c4_raster_R1_ids_256_pages"] = dict(ms=ms_i, pages_per_s=256 / (ms_i * 1e-3))
    from oracle.synth import class_map
    maps = torch.from_numpy(np.stack([class_map(100 + i, 512, 512) for i in range(64)])).cuda()
    maps = maps.repeat(4, 1, 1)

    def post():
        for c in (2, 3, 4):
            closed = morph.closing_batch(morph.class_equals(maps, c), (1, 3))
            morph.ccl_batch(closed)

    ms_p = timed(post, reps=3, warm=1)
    out["c4_closing_ccl_256_pages_x3_classes"] = dict(ms=ms_p, pages_per_s=256 / (ms_p * 1e-3), maps_per_s=768 / (ms_p * 1e-3))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
