"""Per-kernel timing of the BERT-grid (row-id, x_layout 3) train step at batch 8, plus the wall time of its host-side parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import msau_b200
from msau_b200 import _lib, raster
from bench_inputs import synth_page

m = msau_b200.MSAUWrapper(768, 5, dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)).cuda().train()
m.reset_parameters(seed=3)
cp, fe = [], []
for i in range(8):
    _, l = synth_page(500 + i, 512, 512, 198)
    cp.append(l); fe.append(0.3 * np.random.RandomState(500 + i).randn(len(l["x"]), 768))
host = raster.HostBatch(cp, with_chars=False, with_labels=True)
feats = torch.from_numpy(np.concatenate(fe)).pin_memory()

def prep():
    cells = raster.BoxBatch.from_host(host, "cuda")
    geom = cells.geometry()
    table = feats.to("cuda", non_blocking=True)
    ids = raster.raster_features(cells, geom, table, (512, 512), False, "ids")
    lab = raster.raster_labels(cells, geom, (512, 512))
    m.set_feature_table(table)
    return ids, lab

def ev(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

ids, lab = prep()
print("prep (H2D + raster + table)   %.3f ms" % ev(prep))
print("train_step layout 3           %.3f ms" % ev(lambda: m.train_step(ids, lab, layout=3)))
print("prep + step                   %.3f ms" % ev(lambda: m.train_step(*prep(), layout=3)))
_lib.set_option("wgrad_side_stream", 0)
_lib.profile_enable(True)
for _ in range(3): m.train_step(ids, lab, layout=3)
rep = _lib.profile_report(); _lib.profile_enable(False)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    print(f"{k:28s} n={v['launches']//3:3d} {v['ms']/3:7.3f} ms")
print("sum %.3f ms" % (sum(v["ms"] for v in rep.values()) / 3))
