"""Per-layer-shape kernel timing of one train step (MSAU_PROF_DETAIL=1)."""
import os, sys
os.environ.setdefault("MSAU_PROF_DETAIL", "1")      # MSAU_PROF_DETAIL=0: per-kernel-family totals (cheaper host side)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msau_b200
from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for kv in sys.argv[3:]:                      # engine options, e.g. wgrad_side_stream=0 lrn_coop=0
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
S, R = int(os.environ.get("MSAU_S", 4)), int(os.environ.get("MSAU_R", 2))     # MSAU_S=6 MSAU_R=3: the wrapper-default model
cfg = om.MsauConfig(scale_space_num=S, res_depth=R)
m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=8, scale_space_num=S, res_depth=R))
m.load_state_dict(om.init_state_dict(cfg, 0)); m = m.cuda().train()
x, labels = synth_input(cfg.channels, cfg.n_class, B, 512, 512, 3); x, labels = x.cuda(), labels.cuda()
for _ in range(3): m.train_step(x, labels)
torch.cuda.synchronize()
_lib.profile_enable(True)
K = 3
for _ in range(K): m.train_step(x, labels)
rep = _lib.profile_report(); _lib.profile_enable(False)
tot = sum(v["ms"] for v in rep.values())
print("total ms/step", tot / K)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    n = v["launches"] // K
    print(f"{k:52s} n={n:3d} {v['ms']/K:7.3f} ms  {1e3*v['ms']/v['launches']:7.1f} us/launch  {v['bytes']/v['ms']/1e6:7.0f} GB/s {v['flops']/v['ms']/1e9:6.1f} TF")
