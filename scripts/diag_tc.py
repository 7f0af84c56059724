import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import msau_b200
from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input
from test_model_gpu import build, plan_tensors

cfg = om.MsauConfig(channels=16, n_class=5, scale_space_num=3, res_depth=2, feat_root=8)
sd = om.init_state_dict(cfg, 11)
x, labels = synth_input(cfg.channels, cfg.n_class, 1, 40, 48, 12)
leaves = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
out, axo, trace = om.msau_forward_trace(leaves, cfg, x.double(), retain_grad=True)
om.batch_loss(out, axo, labels).backward()
for tc in (0, 1):
    _lib.set_option("tensor_core_conv", tc)
    m = build(cfg, sd).train()
    _, logits, aux = m(x.cuda())
    loss = m.loss(logits, aux, labels.cuda())
    torch.cuda.synchronize()
    got = plan_tensors(m, m._last[0])
    print("tc =", tc)
    for (nm, t), (a, g) in list(zip(trace, got)):
        if t is None: continue
        c = t.shape[1]
        td = t.detach().float()
        e = (a[:, :c] - td).abs().max().item() / td.abs().max().item()
        ge = float('nan')
        if t.grad is not None:
            want = t.grad * (t.detach() > 0) if nm.rsplit(".", 1)[-1].startswith("a") and not nm.endswith("att") else t.grad
            ge = (g[:, :c] - want.float()).abs().max().item() / max(want.abs().max().item(), 1e-30)
        print(f"  {nm:14s} act_rel {e:9.2e}  grad_rel {ge:9.2e}")
