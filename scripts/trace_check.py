"""Per-tensor comparison of the engine's workspace with the oracle's traced forward/backward for one golden fixture.
usage: python scripts/trace_check.py model_s6r3_c16 [tc=0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from msau_b200 import _lib
from oracle import model as om
from oracle.synth import synth_input
import test_model_gpu as T

name = sys.argv[1]
tc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.set_option("tensor_core_conv", tc)
gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z, meta, cfg = T.load(gd, name)
sd = om.init_state_dict(cfg, meta["seed"])
x, labels = synth_input(cfg.channels, cfg.n_class, meta["B"], meta["H"], meta["W"], meta["seed"] + 1)
m = T.build(cfg, sd).train()
_, logits, aux = m(x.cuda())
loss = m.loss(logits, aux, labels.cuda())
pl = m._last[0]
torch.cuda.synchronize()
got = T.plan_tensors(m, pl)
leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
out, axo, trace = om.msau_forward_trace(leaves, cfg, x, retain_grad=True)
ref_loss = om.batch_loss(out, axo, labels)
ref_loss.backward()
print("loss", float(loss.detach()), float(ref_loss))
for (nm, t), (a, g) in zip(trace, got):
    if t is None:
        continue
    c = t.shape[1]
    err = (a[:, :c] - t.detach()).abs().max().item() / max(1.0, t.detach().abs().max().item())
    line = f"{nm:28s} {tuple(t.shape)!s:22s} act {err:9.2e}"
    if t.grad is not None:
        leaf = nm.rsplit(".", 1)[-1]
        post_relu = (leaf.startswith("a") and leaf != "att") or leaf in ("rr", "cc", "uc")
        want = t.grad * (t.detach() > 0) if post_relu else t.grad
        gs = max(want.abs().max().item(), 1e-12)
        diff = (g[:, :c] - want).abs()
        line += f"  grad max {diff.max().item() / gs:9.2e}  l2 {(diff.double().norm() / max(want.double().norm().item(), 1e-30)).item():9.2e}"
    print(line)
# parameter gradients
m._assign_grads()
named = dict(m.named_parameters())
for k, _ in om.param_schema(cfg):
    gr = leaves[k].grad
    if gr is None or named[k].grad is None:
        continue
    d = (named[k].grad.cpu() - gr).double().norm().item() / max(gr.double().norm().item(), 1e-30)
    if d > 2e-3:
        print("param grad", k, f"{d:.3e}")
# where are the errors of a named tensor's gradient?
if len(sys.argv) > 3:
    for (nm, t), (a, g) in zip(trace, got):
        if nm == sys.argv[3]:
            leaf = nm.rsplit(".", 1)[-1]
            post_relu = (leaf.startswith("a") and leaf != "att") or leaf in ("rr", "cc", "uc")
            want = t.grad * (t.detach() > 0) if post_relu else t.grad
            diff = (g[:, :t.shape[1]] - want).abs()
            gs = want.abs().max().item()
            idx = (diff > 2e-3 * gs).nonzero()
            print(nm, "bad elements", idx.shape[0], "of", diff.numel())
            print("channels", sorted(set(idx[:, 1].tolist())))
            print("rows", sorted(set(idx[:, 2].tolist())))
            print("cols", sorted(set(idx[:, 3].tolist())))
            for r in idx[:20].tolist():
                print(r, float(g[tuple(r)]), float(want[tuple(r)]))
