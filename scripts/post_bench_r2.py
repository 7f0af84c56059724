"""Rasterisation + post-process kernels timed alone (BASELINE.json configs[3]: 256 pages x ~200 boxes, 3 class maps per page).

    python scripts/post_bench_r2.py [n_pages] [reps]      -> one JSON object: per-kernel ms / launches / algorithmic GB/s
    (under ncu: `python scripts/post_bench_r2.py 16 1`)

Inputs are resident on the device before the timed region (page records already uploaded), so the numbers are the kernels'
own: CUDA events around every launch (msau_profile_enable) plus an end-to-end event pair around each stage."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench_inputs import synth_page
from msau_b200 import _lib, morph, raster


def stage(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    H = W = 512
    wp, lp = [], []
    for i in range(n):
        w, l = synth_page(i, H, W, 198)
        wp.append(w); lp.append(l)
    words = raster.BoxBatch(wp, "cuda", with_chars=True)
    lines = raster.BoxBatch(lp, "cuda", with_chars=False, with_labels=True)
    table = torch.eye(96, dtype=torch.float64, device="cuda")
    geom = words.geometry()
    owner = torch.empty((n * H * W,), dtype=torch.int32, device="cuda")
    out = dict(n_pages=n, reps=reps, stages={})

    def r1_dense():
        return raster.raster_features(words, geom, table, (H, W), True, "nhwc", owner=owner)

    def r1_ids():
        return raster.raster_features(words, geom, table, (H, W), True, "ids", owner=owner)

    def labels():
        return raster.raster_labels(lines, geom, (H, W), owner=owner)

    from bench_inputs import class_map_rects
    maps = torch.from_numpy(np.stack([class_map_rects(100 + i, H, W) for i in range(min(n, 64))])).cuda()
    maps = maps.repeat((n + maps.shape[0] - 1) // maps.shape[0], 1, 1)[:n].contiguous()

    def post():
        for c in (2, 3, 4):
            closed = morph.class_closing_batch(maps, c, (1, 3))
            morph.ccl_batch(closed)

    _lib.profile_enable(False)
    for name, fn in (("R1_dense_nhwc", r1_dense), ("R1_ids", r1_ids), ("labels", labels), ("closing_ccl_3_classes", post)):
        ms = stage(fn, reps)
        out["stages"][name] = dict(ms=ms, pages_per_s=n / (ms * 1e-3))
    _lib.profile_enable(True)
    for fn in (r1_dense, r1_ids, labels, post):
        for _ in range(reps):
            fn()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    out["kernels"] = {k: dict(ms_per_call=v["ms"] / v["launches"], launches=v["launches"], gbs=v["bytes"] / v["ms"] / 1e6 if v["ms"] else None)
                      for k, v in rep.items()}
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        pk = json.load(open(peaks))["hbm_gbs"]
        for v in out["kernels"].values():
            if v["gbs"]:
                v["frac_of_measured_hbm_peak"] = v["gbs"] / pk
    print(json.dumps(out))


if __name__ == "__main__":
    main()
