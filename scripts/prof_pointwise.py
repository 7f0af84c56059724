"""Driver for an ncu capture of the HBM-bound kernels: dense R1 rasterisation of 16 pages (grid fill), then one fused train step
(pool, 1x1 convs, LRN, loss head) -- see profiles/ncu_r1_streaming_kernels_summary.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import msau_b200
from msau_b200 import raster
from bench_inputs import synth_page

pages = [synth_page(i, 512, 512, 198) for i in range(16)]
wp, lp = [p[0] for p in pages], [p[1] for p in pages]
table = torch.eye(96, dtype=torch.float64, device="cuda")
m = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2)).cuda().train()
m.reset_parameters(seed=0)
for _ in range(2):
    grid, label, _ = raster.rasterize_word_chargrid(wp, lp, table, out_hw=(512, 512), layout="nchw")
    loss = m.train_step(grid, label.long())
torch.cuda.synchronize()
print("loss", float(loss))
