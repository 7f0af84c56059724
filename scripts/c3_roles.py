"""Where the three warp roles of conv3_tc spend their cycles (MSAU_TC_DEBUG=32 role timers), one conv shape at a time.

    MSAU_TC_DEBUG=32 python scripts/c3_roles.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msau_b200
from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input

cfg = om.MsauConfig()
m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2))
m.load_state_dict(om.init_state_dict(cfg, 0))
m = m.cuda().train()
m.set_option("wgrad_side_stream", 0)
x, labels = synth_input(cfg.channels, cfg.n_class, 16, 512, 512, 3)
x, labels = x.cuda(), labels.cuda()
buf = (C.c_ulonglong * 16)()
for _ in range(2):
    m.train_step(x, labels)
_lib.check(_lib.lib().msau_debug_c3_prof(buf))
m.train_step(x, labels)
_lib.check(_lib.lib().msau_debug_c3_prof(buf))
v = list(buf)
ctas = max(v[12], 1)
names = {0: "producer: wait stage free (MMA retired)", 1: "producer: wait raw plane (TMA)", 2: "producer: named barrier", 3: "producer: warp total",
         4: "MMA: wait operands (producers)", 5: "MMA: wait accumulator (epilogue)", 6: "MMA: warp total",
         8: "epilogue: wait accumulator (MMA)", 9: "epilogue: wait extras (TMA)", 10: "epilogue: warp total"}
nw = {0: 6, 1: 6, 2: 6, 3: 6, 4: 2, 5: 2, 6: 2, 8: 8, 9: 8, 10: 8}
tot = {0: v[3], 1: v[3], 2: v[3], 3: v[3], 4: v[6], 5: v[6], 6: v[6], 8: v[10], 9: v[10], 10: v[10]}
print(f"one train step, all conv3_tc launches: {ctas} CTAs")
for k in sorted(names):
    print(f"  {names[k]:45s} {v[k] / ctas / nw[k]:12.0f} cycles per warp and CTA   {100.0 * v[k] / max(tot[k], 1):5.1f} % of the role's time")
