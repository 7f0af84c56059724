"""Tensor-core self-attention operator (msau_attention_forward / _backward, C ABI) against a torch fp64 restatement
of SelfAttentionBlock.forward (model/layers/attention.py:152-162: s = g^T f, soft-max over the LAST axis, contraction
over the FIRST) and its autograd gradient.

Tolerances: forward |d out| <= 3e-5 * max|o| (bf16 hi/lo split, fp32 accumulate); the backward uses single bf16 terms:
<= 1 % in L2, <= 3 % of the tensor's max element-wise."""
import pytest
import torch

from msau_b200 import _lib

pytestmark = pytest.mark.gpu
LOG2E = 1.4426950408889634


def reference(fg, hh, x, do, d):
    fg = fg.double().requires_grad_(True)
    hh = hh.double().requires_grad_(True)
    f, g = fg[..., :d], fg[..., d:]
    s = torch.einsum("bic,bjc->bij", g, f)
    beta = torch.softmax(s, dim=-1)
    out = x.double() + torch.einsum("bij,bic->bjc", beta, hh)
    out.backward(do.double())
    return out.detach(), torch.logsumexp(s, dim=-1).detach() * LOG2E, fg.grad, hh.grad


def run_operator(fg, hh, x, do):
    B, N, Cc = hh.shape
    L = _lib.lib()
    nb = L.msau_attention_scratch_bytes(B, N, Cc)
    scratch = torch.empty(nb + 256, dtype=torch.uint8, device="cuda")
    sp = (scratch.data_ptr() + 255) // 256 * 256
    lse = torch.empty(B, N, device="cuda")
    out = torch.empty_like(x)
    dfg = torch.full_like(fg, float("nan"))
    dhh = torch.full_like(hh, float("nan"))
    st = _lib.current_stream()
    _lib.check(L.msau_attention_forward(fg.data_ptr(), hh.data_ptr(), x.data_ptr(), B, N, Cc, lse.data_ptr(), out.data_ptr(), sp, nb, st))
    _lib.check(L.msau_attention_backward(fg.data_ptr(), hh.data_ptr(), do.data_ptr(), lse.data_ptr(), B, N, Cc, dfg.data_ptr(),
                                         dhh.data_ptr(), sp, nb, st))
    torch.cuda.synchronize()
    return out, lse, dfg, dhh


@pytest.mark.parametrize("B,N,Cc,scale", [(1, 128, 64, 0.5), (2, 256, 64, 1.0), (2, 1008, 64, 1.0), (3, 96, 32, 1.0),
                                          (1, 600, 32, 2.0), (1, 1, 64, 1.0), (2, 4096, 64, 0.3), (1, 4096, 64, 2.0)])
def test_attention_operator_matches_reference(B, N, Cc, scale):
    d = Cc // 8
    g = torch.Generator(device="cuda").manual_seed(N + Cc)
    fg = torch.randn(B, N, 2 * d, device="cuda", generator=g) * scale
    hh = torch.randn(B, N, Cc, device="cuda", generator=g)
    x = torch.randn(B, N, Cc, device="cuda", generator=g)
    do = torch.randn(B, N, Cc, device="cuda", generator=g)
    out, lse, dfg, dhh = run_operator(fg, hh, x, do)
    ro, rl, rfg, rhh = reference(fg, hh, x, do, d)
    o_scale = max((ro - x.double()).abs().max().item(), 1.0)      # column sums of P are unbounded: |o| can exceed |h|
    errs = dict(out=(out.double() - ro).abs().max().item() / o_scale,
                lse=(lse.double() - rl).abs().max().item() / max(rl.abs().max().item(), 1.0))
    assert errs["out"] <= 5e-5 and errs["lse"] <= 5e-5, errs
    for name, got, want in (("dhh", dhh, rhh), ("df", dfg[..., :d], rfg[..., :d]), ("dg", dfg[..., d:], rfg[..., d:])):
        assert torch.isfinite(got).all(), name
        errs[name + "_max"] = (got.double() - want).abs().max().item() / (want.abs().max().item() + 1e-5)
        errs[name + "_l2"] = (got.double() - want).norm().item() / (want.norm().item() + 1e-5)
        # (N = 1: the true gradient of f, g is exactly 0, hence the absolute floor)
        assert errs[name + "_max"] <= 3e-2 and errs[name + "_l2"] <= 1.5e-2, errs


def test_attention_scratch_too_small_is_an_error():
    L = _lib.lib()
    t = torch.zeros(64 * 128, device="cuda")
    rc = L.msau_attention_forward(t.data_ptr(), t.data_ptr(), t.data_ptr(), 1, 128, 64, t.data_ptr(), t.data_ptr(), t.data_ptr(), 16,
                                  _lib.current_stream())
    assert rc == -4
    rc = L.msau_attention_forward(t.data_ptr(), t.data_ptr(), t.data_ptr(), 1, 128, 48, t.data_ptr(), t.data_ptr(), t.data_ptr(), 1 << 30,
                                  _lib.current_stream())
    assert rc == -1
