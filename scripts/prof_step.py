"""Small driver for ncu: a couple of fused train steps (B pages of 512x512, train-script config)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msau_b200
from oracle import model as om
from oracle.synth import synth_input

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = om.MsauConfig()
m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=8, scale_space_num=4, res_depth=2))
m.load_state_dict(om.init_state_dict(cfg, 0))
m = m.cuda().train()
x, labels = synth_input(cfg.channels, cfg.n_class, B, 512, 512, 3)
x, labels = x.cuda(), labels.cuda()
for _ in range(steps):
    loss = m.train_step(x, labels)
torch.cuda.synchronize()
print("loss", float(loss))
