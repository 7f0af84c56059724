"""Torch-fp32 (or fp64) CPU restatement of the MSAU network, loss and train step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Functional, stateless: every
function takes the reference ``state_dict`` (same key schema as the reference's
``MSAUWrapper.state_dict()``) and plain tensors.

What it follows, by reference file:line (paths relative to /root/reference):

* SAME padding arithmetic ............ model/layers/utils.py:5-28
* dilated conv + LRN ................. model/layers/layers.py:105-164 (use_lrn=True, activation=None
                                        as constructed at model/model.py:100-104)
* conv (+ReLU) ....................... model/layers/layers.py:10-102
* transposed conv w/ output_size ..... model/layers/layers.py:207-260, model/model.py:230
* residual multi-conv block .......... model/model.py:37-50
* down tower ......................... model/model.py:129-164
* up tower ........................... model/model.py:224-259
* SAGAN-style self attention ......... model/layers/attention.py:138-162
* 3 coupled blocks + 4x4 heads ....... model/model.py:378-396
* softmax predictor, masked CE loss .. model/model.py:426-437, 446-459
* train step (clip 1.0 + Adam 1e-4) .. train_chargrid_funsd_msau.py:24-26,45-59

Pinned by tests/golden/model_*.npz (generated from the unmodified reference by
tests/golden/make_golden.py) -- see tests/test_oracle_model.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class MsauConfig:
    channels: int = 96
    n_class: int = 5
    scale_space_num: int = 4
    res_depth: int = 2
    feat_root: int = 8
    filter_size: int = 3
    pool_size: int = 2
    num_blocks: int = 3

    def feat(self, level: int) -> int:
        return self.feat_root * self.pool_size ** level


# ----------------------------------------------------------------------------- schema
def param_schema(cfg: MsauConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) in the reference's ``state_dict`` iteration order.

    Order = torch module registration order in model/model.py:
    down tower: conv_res_list, conv1s, conv1_1s, layer_attentions (:79-127);
    up tower:   conv_res_list, conv1s, conv1_1s, deconvs (:180-222);
    MSAUNet:    blocks.{0,1,2} then end_convs.{0,1,2} (:355-376).
    """
    S, R, k = cfg.scale_space_num, cfg.res_depth, cfg.filter_size
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(prefix: str, co: int, ci: int, kh: int, kw: int) -> None:
        out.append((prefix + ".weight", (co, ci, kh, kw)))
        out.append((prefix + ".bias", (co,)))

    for b in range(cfg.num_blocks):
        cin0 = cfg.channels if b == 0 else cfg.n_class
        dn = f"msau_net.blocks.{b}.downsamplingblock."
        for l in range(S):
            f = cfg.feat(l)
            for r in range(R):
                conv(dn + f"conv_res_list.{l}.conv_res_list.{r}.custom_conv", f, f, k, k)
        for l in range(S):
            f = cfg.feat(l)
            conv(dn + f"conv1s.{l}.conv", f, cin0 if l == 0 else cfg.feat(l - 1), k, k)
        if b > 0:
            for l in range(S):
                f = cfg.feat(l)
                conv(dn + f"conv1_1s.{l}.custom_conv", f, 2 * f, 1, 1)
        fa = cfg.feat(S - 1)
        att = dn + "layer_attentions.attention_block."
        conv(att + "f.conv", fa // 8, fa, 1, 1)
        conv(att + "g.conv", fa // 8, fa, 1, 1)
        conv(att + "h.conv", fa, fa, 1, 1)
        up = f"msau_net.blocks.{b}.upsamplingblock."
        for l in range(S - 1):
            f = cfg.feat(l)
            for r in range(R):
                conv(up + f"conv_res_list.{l}.conv_res_list.{r}.custom_conv", f, f, k, k)
        for l in range(S - 1):
            f = cfg.feat(l)
            conv(up + f"conv1s.{l}.custom_conv", f, 2 * f, k, k)
        if b > 0:
            for l in range(S - 1):
                f = cfg.feat(l)
                conv(up + f"conv1_1s.{l}.custom_conv", f, 2 * f, 1, 1)
        for l in range(S - 1):
            f = cfg.feat(l)
            # ConvTranspose2d(in=2f, out=f): weight [in, out, kh, kw]  (layers.py:221-226)
            out.append((up + f"deconvs.{l}.conv.weight", (2 * f, f, k, k)))
            out.append((up + f"deconvs.{l}.conv.bias", (f,)))
    for b in range(cfg.num_blocks):
        conv(f"msau_net.end_convs.{b}.custom_conv", cfg.n_class, cfg.feat_root, 4, 4)
    return out


def init_state_dict(cfg: MsauConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random weights with the reference's *distributions* (not its RNG stream):
    conv/deconv N(0, sqrt(2/(kh*kw*Cin+Cout))), bias N(0.1, 1e-5) (layers.py:33-36,59-60);
    attention 1x1 convs: torch Conv2d default, U(+-1/sqrt(fan_in)) (attention.py:19-21)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape in param_schema(cfg):
        attn = ".attention_block." in key
        if key.endswith(".weight"):
            if attn:
                bound = 1.0 / math.sqrt(shape[1])
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
            else:
                if ".deconvs." in key:
                    cin_k, cout_k = shape[1], shape[0]  # kernel_shape[2]=out(f), [3]=in(2f)
                else:
                    cin_k, cout_k = shape[1], shape[0]
                std = math.sqrt(2.0 / (shape[2] * shape[3] * cin_k + cout_k))
                t = torch.randn(shape, generator=g, dtype=torch.float64) * std
        else:
            if attn:
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
            else:
                t = 0.1 + 1e-5 * torch.randn(shape, generator=g, dtype=torch.float64)
        sd[key] = t.to(dtype)
    return sd


# ----------------------------------------------------------------------------- layers
def same_pad(size: int, k: int, stride: int = 1, dilation: int = 1) -> Tuple[int, int]:
    """TF 'SAME' (before, after) padding for one axis -- model/layers/utils.py:5-28."""
    k_eff = k + (k - 1) * (dilation - 1)
    out = -(-size // stride)
    total = max((out - 1) * stride + k_eff - size, 0)
    before = total // 2
    return before, total - before


def _pad_same(x: Tensor, kh: int, kw: int, stride: int = 1, dilation: int = 1) -> Tensor:
    pt, pb = same_pad(x.shape[2], kh, stride, dilation)
    pl, pr = same_pad(x.shape[3], kw, stride, dilation)
    return F.pad(x, (pl, pr, pt, pb))


def conv_same(x: Tensor, w: Tensor, b: Tensor, dilation: int = 1) -> Tensor:
    return F.conv2d(_pad_same(x, w.shape[2], w.shape[3], 1, dilation), w, b, dilation=dilation)


def lrn_full(z: Tensor) -> Tensor:
    """LocalResponseNorm(size=C_out) with torch defaults alpha=1e-4, beta=0.75, k=1
    (layers.py:145,161-162)."""
    return F.local_response_norm(z, z.shape[1], alpha=1e-4, beta=0.75, k=1.0)


def res_block(sd, prefix: str, x: Tensor, R: int) -> Tensor:
    """model/model.py:37-50 -- relu(x) -> R convs (ReLU on all but the last) -> +x -> relu."""
    o = x
    t = F.relu(x)
    for r in range(R):
        p = f"{prefix}.conv_res_list.{r}.custom_conv"
        t = conv_same(t, sd[p + ".weight"], sd[p + ".bias"])
        if r < R - 1:
            t = F.relu(t)
    return F.relu(t + o)


def self_attention(sd, prefix: str, x: Tensor) -> Tensor:
    """model/layers/attention.py:152-162.  s = g^T f, softmax over the LAST axis (j),
    o = h @ beta (contraction over the FIRST axis i), out = x + o."""
    B, C, H, W = x.shape
    f = F.conv2d(x, sd[prefix + ".f.conv.weight"], sd[prefix + ".f.conv.bias"]).reshape(B, -1, H * W)
    g = F.conv2d(x, sd[prefix + ".g.conv.weight"], sd[prefix + ".g.conv.bias"]).reshape(B, -1, H * W)
    h = F.conv2d(x, sd[prefix + ".h.conv.weight"], sd[prefix + ".h.conv.bias"]).reshape(B, C, H * W)
    s = torch.matmul(g.transpose(1, 2), f)
    beta = torch.softmax(s, dim=-1)
    o = torch.matmul(h, beta).reshape(B, C, H, W)
    return o + x


def down_tower(sd, cfg: MsauConfig, b: int, x: Tensor, prev_dw, with_attention: bool = True):
    """model/model.py:129-164.  Returns (dw dict, pre-attention deepest feature)."""
    S = cfg.scale_space_num
    pre = f"msau_net.blocks.{b}.downsamplingblock"
    dw = {}
    inp = x
    for l in range(S):
        p = f"{pre}.conv1s.{l}.conv"
        z = conv_same(inp, sd[p + ".weight"], sd[p + ".bias"], dilation=2 ** l)
        t = lrn_full(z)
        t = res_block(sd, f"{pre}.conv_res_list.{l}", t, cfg.res_depth)
        if b > 0:
            p = f"{pre}.conv1_1s.{l}.custom_conv"
            t = F.relu(F.conv2d(torch.cat([prev_dw[l], t], 1), sd[p + ".weight"], sd[p + ".bias"]))
        if l == S - 1:
            # the attention output is only ever read by the NEXT block's coupling concat
            dw[l] = self_attention(sd, f"{pre}.layer_attentions.attention_block", t) if with_attention else t
        else:
            dw[l] = t
            inp = F.max_pool2d(_pad_same(t, cfg.pool_size, cfg.pool_size, cfg.pool_size), cfg.pool_size)
    return dw, t


def up_tower(sd, cfg: MsauConfig, b: int, dw, x: Tensor, prev_up):
    """model/model.py:224-259."""
    S = cfg.scale_space_num
    pre = f"msau_net.blocks.{b}.upsamplingblock"
    up = {}
    for l in range(S - 2, -1, -1):
        skip = dw[l]
        p = f"{pre}.deconvs.{l}.conv"
        hin, win = x.shape[2:]
        hout, wout = skip.shape[2:]
        k = cfg.filter_size
        opad = (hout - ((hin - 1) * 2 - 2 * (k // 2) + k), wout - ((win - 1) * 2 - 2 * (k // 2) + k))
        d = F.conv_transpose2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=k // 2,
                               output_padding=opad)
        p = f"{pre}.conv1s.{l}.custom_conv"
        t = conv_same(torch.cat([skip, d], 1), sd[p + ".weight"], sd[p + ".bias"])
        t = res_block(sd, f"{pre}.conv_res_list.{l}", t, cfg.res_depth)
        if b > 0:
            p = f"{pre}.conv1_1s.{l}.custom_conv"
            t = F.relu(F.conv2d(torch.cat([prev_up[l], t], 1), sd[p + ".weight"], sd[p + ".bias"]))
        up[l] = t
        x = t
    return x, up


def msau_forward(sd, cfg: MsauConfig, x: Tensor) -> Tuple[Tensor, Tensor]:
    """model/model.py:378-396 -> (logits, logits_aux).  The last block's attention output is
    dead (never read) and is skipped; results are unaffected."""
    prev_dw = prev_up = None
    aux = None
    inp = x
    for b in range(cfg.num_blocks):
        dw, deep = down_tower(sd, cfg, b, inp, prev_dw, with_attention=(b < cfg.num_blocks - 1))
        out, up = up_tower(sd, cfg, b, dw, deep, prev_up)
        p = f"msau_net.end_convs.{b}.custom_conv"
        out = conv_same(out, sd[p + ".weight"], sd[p + ".bias"])
        prev_dw, prev_up = dw, up
        inp = out
        if b == cfg.num_blocks - 2:
            aux = out
    return out, aux


def page_loss(logits: Tensor, aux: Tensor, label: Tensor) -> Tensor:
    """model/model.py:446-459 for ONE page: CE over pixels with label != 0 (mean), main + aux.
    logits [1,C,H,W], label [1,H,W] integer."""
    keep = label[0] != 0
    tgt = label[0][keep].long()
    lg = logits[0][:, keep].t()
    la = aux[0][:, keep].t()
    return F.cross_entropy(lg, tgt) + F.cross_entropy(la, tgt)


def batch_loss(logits: Tensor, aux: Tensor, labels: Tensor) -> Tensor:
    """Batched definition: mean over pages of the reference per-page loss."""
    B = logits.shape[0]
    tot = logits.new_zeros(())
    for b in range(B):
        tot = tot + page_loss(logits[b:b + 1], aux[b:b + 1], labels[b:b + 1])
    return tot / B


def loss_and_grads(sd, cfg: MsauConfig, x: Tensor, labels: Tensor):
    """forward + batched loss + backward.  Returns (loss, logits, aux, {key: grad or None})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, aux = msau_forward(leaves, cfg, x)
    loss = batch_loss(logits, aux, labels)
    loss.backward()
    return loss.detach(), logits.detach(), aux.detach(), {k: v.grad for k, v in leaves.items()}


def clip_adam_step(sd, grads, m, v, step: int, lr: float = 1e-4, b1: float = 0.9, b2: float = 0.999,
                   eps: float = 1e-8, max_norm: float = 1.0):
    """train_chargrid_funsd_msau.py:58-59: clip_grad_norm(params, 1.0) then Adam(lr=1e-4).step().
    Parameters whose grad is None (the dead last-block attention) are skipped by both, exactly as
    torch does.  Updates sd/m/v in place; returns the pre-clip total norm."""
    live = [k for k in sd if grads[k] is not None]
    total = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in live)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k in live:
        g = grads[k] * coef
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        sd[k].addcdiv_(m[k], denom, value=-(lr / bc1))
    return total


# ----------------------------------------------------------------------------- alternative trainer (UNetLoss + RMSprop / SGD)
def unet_loss(logits: Tensor, tgt_onehot: Tensor, aux_logits: Tensor = None, aux_tgt_onehot: Tensor = None, class_weights=None):
    """UNetLoss.forward, model/training/cost.py:35-65: targets = argmax of the one-hot maps (:41,:52), CrossEntropyLoss over ALL
    pixels (optionally class-weighted, :27-31), loss = 0.5 final + 0.5 aux (:61), masked accuracy over tgt != 0 (:44-50).
    Returns (acc float, loss, final_loss or None)."""
    tgt = torch.argmax(tgt_onehot, dim=1)
    pred = torch.argmax(logits, dim=1)
    nz = tgt != 0
    acc = float((tgt[nz] == pred[nz]).sum().double() / nz.sum().double())
    w = None if class_weights is None else torch.tensor(class_weights, dtype=logits.dtype, device=logits.device)
    final = F.cross_entropy(logits, tgt, weight=w)
    if aux_logits is None:
        return acc, final, None
    aux = F.cross_entropy(aux_logits, torch.argmax(aux_tgt_onehot, dim=1), weight=w)
    return acc, 0.5 * final + 0.5 * aux, final


def rmsprop_step(sd, grads, sq, lr: float = 0.001, alpha: float = 0.99, eps: float = 1e-8, weight_decay: float = 0.0):
    """torch.optim.RMSprop as constructed by get_optimizer (model/training/optimizer.py:14-16): momentum 0, not centred.
    Parameters whose grad is None are skipped."""
    for k in sd:
        if grads[k] is None:
            continue
        g = grads[k] if weight_decay == 0.0 else grads[k] + weight_decay * sd[k]
        sq[k].mul_(alpha).addcmul_(g, g, value=1 - alpha)
        sd[k].addcdiv_(g, sq[k].sqrt().add_(eps), value=-lr)


def sgd_momentum_step(sd, grads, buf, step: int, lr: float = 0.001, momentum: float = 0.9, weight_decay: float = 0.0):
    """torch.optim.SGD(params, lr, momentum) (optimizer.py:8-13): buf = g at the first step, then momentum * buf + g."""
    for k in sd:
        if grads[k] is None:
            continue
        g = grads[k] if weight_decay == 0.0 else grads[k] + weight_decay * sd[k]
        if step == 1:
            buf[k].copy_(g)
        else:
            buf[k].mul_(momentum).add_(g)
        sd[k].add_(buf[k], alpha=-lr)


def unet_loss_and_grads(sd, cfg: MsauConfig, x: Tensor, tgt_onehot: Tensor, aux_tgt_onehot: Tensor, class_weights=None):
    """forward + UNetLoss + backward (model/training/trainer.py:122-136).  Returns (acc, loss, final_loss, logits, aux, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, aux = msau_forward(leaves, cfg, x)
    acc, loss, final = unet_loss(logits, tgt_onehot, aux, aux_tgt_onehot, class_weights)
    loss.backward()
    grads = {k: leaves[k].grad for k in leaves}
    return acc, loss.detach(), final.detach(), logits.detach(), aux.detach(), grads


# ----------------------------------------------------------------------------- traced forward (debugging aid)
def msau_forward_trace(sd, cfg: MsauConfig, x: Tensor, retain_grad: bool = False):
    """Same arithmetic as ``msau_forward`` but also returns every intermediate activation, in the order the
    CUDA plan allocates them (msau_b200/csrc/plan.cu), as a list of (name, tensor NCHW).  Entries that the
    engine allocates but that have no oracle counterpart (soft-max row statistics) are (name, None)."""
    S, R, NB = cfg.scale_space_num, cfg.res_depth, cfg.num_blocks
    trace = []

    def rec(name, t):
        if retain_grad and t is not None and t.requires_grad:
            t.retain_grad()
        trace.append((name, t))
        return t

    def res(prefix, name, t):
        o = t
        cur = F.relu(t)
        for r in range(R):
            p = f"{prefix}.conv_res_list.{r}.custom_conv"
            cur = conv_same(cur, sd[p + ".weight"], sd[p + ".bias"])
            if r < R - 1:
                cur = rec(f"{name}.a{r}", F.relu(cur))
        return rec(f"{name}.rr", F.relu(cur + o))

    prev_dw = prev_up = None
    inp = x
    aux = None
    for b in range(NB):
        dn = f"msau_net.blocks.{b}.downsamplingblock"
        dw = {}
        cur_in = inp
        for l in range(S):
            p = f"{dn}.conv1s.{l}.conv"
            z = rec(f"b{b}.d{l}.z1", conv_same(cur_in, sd[p + ".weight"], sd[p + ".bias"], dilation=2 ** l))
            t = rec(f"b{b}.d{l}.y1", lrn_full(z))
            t = res(f"{dn}.conv_res_list.{l}", f"b{b}.d{l}", t)
            if b > 0:
                p = f"{dn}.conv1_1s.{l}.custom_conv"
                t = rec(f"b{b}.d{l}.cc", F.relu(F.conv2d(torch.cat([prev_dw[l], t], 1), sd[p + ".weight"], sd[p + ".bias"])))
            dw[l] = t
            if l < S - 1:
                cur_in = rec(f"b{b}.d{l}.pooled",
                             F.max_pool2d(_pad_same(t, cfg.pool_size, cfg.pool_size, cfg.pool_size), cfg.pool_size))
        deep = t
        if b < NB - 1:
            ap = f"{dn}.layer_attentions.attention_block"
            f_ = F.conv2d(deep, sd[ap + ".f.conv.weight"], sd[ap + ".f.conv.bias"])
            g_ = F.conv2d(deep, sd[ap + ".g.conv.weight"], sd[ap + ".g.conv.bias"])
            rec(f"b{b}.fg", torch.cat([f_, g_], 1))
            h_ = rec(f"b{b}.hh", F.conv2d(deep, sd[ap + ".h.conv.weight"], sd[ap + ".h.conv.bias"]))
            B_, C_, H_, W_ = deep.shape
            s = torch.matmul(g_.reshape(B_, -1, H_ * W_).transpose(1, 2), f_.reshape(B_, -1, H_ * W_))
            beta = torch.softmax(s, dim=-1)
            o = torch.matmul(h_.reshape(B_, C_, H_ * W_), beta).reshape(B_, C_, H_, W_)
            dw[S - 1] = rec(f"b{b}.att", o + deep)
            for nm in ("mrow", "zinv", "dvec"):
                trace.append((f"b{b}.{nm}", None))
        up = {}
        xx = deep
        un = f"msau_net.blocks.{b}.upsamplingblock"
        for l in range(S - 2, -1, -1):
            skip = dw[l]
            p = f"{un}.deconvs.{l}.conv"
            hin, win = xx.shape[2:]
            hout, wout = skip.shape[2:]
            opad = (hout - (2 * hin - 1), wout - (2 * win - 1))
            d = rec(f"b{b}.u{l}.d", F.conv_transpose2d(xx, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1,
                                                      output_padding=opad))
            p = f"{un}.conv1s.{l}.custom_conv"
            t = rec(f"b{b}.u{l}.u", conv_same(torch.cat([skip, d], 1), sd[p + ".weight"], sd[p + ".bias"]))
            t = res(f"{un}.conv_res_list.{l}", f"b{b}.u{l}", t)
            if b > 0:
                p = f"{un}.conv1_1s.{l}.custom_conv"
                t = rec(f"b{b}.u{l}.uc", F.relu(F.conv2d(torch.cat([prev_up[l], t], 1), sd[p + ".weight"], sd[p + ".bias"])))
            up[l] = t
            xx = t
        p = f"msau_net.end_convs.{b}.custom_conv"
        out = rec(f"b{b}.logits", conv_same(xx, sd[p + ".weight"], sd[p + ".bias"]))
        prev_dw, prev_up = dw, up
        inp = out
        if b == NB - 2:
            aux = out
    return out, aux, trace
