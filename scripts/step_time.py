"""Wall (CUDA-event) time of the fused train step under engine options: python scripts/step_time.py [B] [graph|eager] [opt=val ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msau_b200
from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = sys.argv[2] if len(sys.argv) > 2 else "eager"
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
cfg = om.MsauConfig()
m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2))
m.load_state_dict(om.init_state_dict(cfg, 0)); m = m.cuda().train()
x, labels = synth_input(cfg.channels, cfg.n_class, B, 512, 512, 3); x, labels = x.cuda(), labels.cuda()
ug = "static" if mode == "graph" else False
for _ in range(4): m.train_step(x, labels, use_graph=ug)
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): m.train_step(x, labels, use_graph=ug)
e1.record(); torch.cuda.synchronize()
print(f"{mode} {' '.join(sys.argv[3:]):40s} {e0.elapsed_time(e1) / K:8.3f} ms/step  {B / (e0.elapsed_time(e1) / K) * 1e3:7.1f} pages/s")
