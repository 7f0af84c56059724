"""Stand-alone check + timing of the tensor-core attention operator against a torch fp64 reference."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msau_b200 import _lib


def ref(fg, hh, x, do, d):
    fg = fg.double().requires_grad_(True); hh = hh.double().requires_grad_(True)
    f, g = fg[..., :d], fg[..., d:]
    s = torch.einsum('bic,bjc->bij', g, f)
    beta = torch.softmax(s, dim=-1)
    out = x.double() + torch.einsum('bij,bic->bjc', beta, hh)
    out.backward(do.double())
    lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    return out.detach(), lse2.detach(), fg.grad, hh.grad


def run(B, N, Cc, scale, seed=0, timing=False):
    d = Cc // 8
    g = torch.Generator(device="cuda").manual_seed(seed)
    fg = torch.randn(B, N, 2 * d, device="cuda", generator=g) * scale
    hh = torch.randn(B, N, Cc, device="cuda", generator=g)
    x = torch.randn(B, N, Cc, device="cuda", generator=g)
    do = torch.randn(B, N, Cc, device="cuda", generator=g)
    L = _lib.lib()
    nb = L.msau_attention_scratch_bytes(B, N, Cc)
    scratch = torch.empty(nb + 256, dtype=torch.uint8, device="cuda")
    sp = (scratch.data_ptr() + 255) // 256 * 256
    lse = torch.empty(B, N, device="cuda"); out = torch.empty_like(x)
    dfg = torch.zeros_like(fg); dhh = torch.zeros_like(hh)
    st = _lib.current_stream()
    _lib.check(L.msau_attention_forward(fg.data_ptr(), hh.data_ptr(), x.data_ptr(), B, N, Cc, lse.data_ptr(), out.data_ptr(), sp, nb, st))
    _lib.check(L.msau_attention_backward(fg.data_ptr(), hh.data_ptr(), do.data_ptr(), lse.data_ptr(), B, N, Cc, dfg.data_ptr(), dhh.data_ptr(), sp, nb, st))
    torch.cuda.synchronize()
    res = {}
    if B * N * N <= 2 * 4096 * 4096:
        ro, rl, rfg, rhh = ref(fg, hh, x, do, d)
        o_only = (out.double() - x.double()); ro_only = ro - x.double()
        res = dict(out=float((out.double() - ro).abs().max()), o_rel=float((o_only - ro_only).abs().max() / ro_only.abs().max()),
                   lse=float((lse.double() - rl).abs().max()),
                   dhh=float((dhh.double() - rhh).abs().max() / rhh.abs().max()),
                   dhh_l2=float((dhh.double() - rhh).norm() / rhh.norm()),
                   dg=float((dfg.double() - rfg)[..., d:].abs().max() / rfg[..., d:].abs().max()),
                   df=float((dfg.double() - rfg)[..., :d].abs().max() / rfg[..., :d].abs().max()),
                   dfg_l2=float((dfg.double() - rfg).norm() / rfg.norm()))
    if timing:
        for name, fn in (("fwd", lambda: L.msau_attention_forward(fg.data_ptr(), hh.data_ptr(), x.data_ptr(), B, N, Cc, lse.data_ptr(), out.data_ptr(), sp, nb, st)),
                         ("bwd", lambda: L.msau_attention_backward(fg.data_ptr(), hh.data_ptr(), do.data_ptr(), lse.data_ptr(), B, N, Cc, dfg.data_ptr(), dhh.data_ptr(), sp, nb, st))):
            for _ in range(2): fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            res[name + "_ms"] = e0.elapsed_time(e1) / 5
    print(f"B={B} N={N} C={Cc} scale={scale}: " + " ".join(f"{k}={v:.3g}" for k, v in res.items()), flush=True)
    return res


if __name__ == "__main__":
    run(1, 128, 64, 0.5)
    run(2, 256, 64, 1.0)
    run(2, 1008, 64, 1.0)
    run(3, 96, 32, 1.0)
    run(1, 600, 32, 2.0)
    run(2, 4096, 64, 0.3)
    run(2, 4096, 64, 2.0)
    run(16, 4096, 64, 0.3, timing=True)
    run(4, 12288, 64, 0.3, timing=True)
