"""Drop-in ``MSAUWrapper`` backed by the CUDA engine (C ABI in ``include/msau_b200.h``).

Mirrors the reference's public model interface, model/model.py:399-459:

    MSAUWrapper(channels=1, n_class=2, model_kwargs={})
    .forward(inp[B,C,H,W] fp32) -> (predictor(logits), logits, aux_logits)      :435-437
    .loss(out_grid, out_grid_aux, label_mask) -> scalar                          :446-459
    .save(path) / .load_weights(path)                                            :439-444
    .parameters() / .state_dict() / .train() / .eval() / .to() / .zero_grad()    (torch.nn.Module)

``state_dict()`` has exactly the reference's keys, shapes and order, so
checkpoints move both ways.  Every parameter is a view into ONE flat fp32 buffer (what the kernels, the
NCCL all-reduce and the fused clip+Adam consume).

The reference training loop (train_chargrid_funsd_msau.py:45-59) works unchanged:
``model(V)`` -> ``model.loss(...)`` -> ``loss.backward()`` -> ``clip_grad_norm`` -> ``optimizer.step()``;
``loss.backward()`` runs the hand-written backward kernels and fills ``p.grad`` (views of a flat gradient
buffer).  ``train_step`` is the fused fast path (forward + loss + backward [+ all-reduce] + clip + Adam).

Differences from the reference, on purpose:
  * batched loss: the reference's ``loss`` only works for B=1 (:452-457); here B>1 means "mean over pages of
    the per-page loss" (SURVEY.md D6);
  * ``final_act="sigmoid"`` raises the same TypeError the reference raises at construction (:428-429);
  * there is no CPU path: tensors must live on a CUDA device when ``forward`` is called.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


def param_schema(channels: int, n_class: int, S: int, R: int, feat_root: int, k: int = 3, num_blocks: int = 3):
    """(key, shape) list in the reference's state_dict order (module registration order of
    model/model.py:79-127, 180-222, 355-376)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(prefix, co, ci, kh, kw):
        out.append((prefix + ".weight", (co, ci, kh, kw)))
        out.append((prefix + ".bias", (co,)))

    feat = lambda l: feat_root * 2 ** l  # noqa: E731
    for b in range(num_blocks):
        cin0 = channels if b == 0 else n_class
        dn = f"msau_net.blocks.{b}.downsamplingblock."
        for l in range(S):
            for r in range(R):
                conv(dn + f"conv_res_list.{l}.conv_res_list.{r}.custom_conv", feat(l), feat(l), k, k)
        for l in range(S):
            conv(dn + f"conv1s.{l}.conv", feat(l), cin0 if l == 0 else feat(l - 1), k, k)
        if b > 0:
            for l in range(S):
                conv(dn + f"conv1_1s.{l}.custom_conv", feat(l), 2 * feat(l), 1, 1)
        fa = feat(S - 1)
        att = dn + "layer_attentions.attention_block."
        conv(att + "f.conv", fa // 8, fa, 1, 1)
        conv(att + "g.conv", fa // 8, fa, 1, 1)
        conv(att + "h.conv", fa, fa, 1, 1)
        up = f"msau_net.blocks.{b}.upsamplingblock."
        for l in range(S - 1):
            for r in range(R):
                conv(up + f"conv_res_list.{l}.conv_res_list.{r}.custom_conv", feat(l), feat(l), k, k)
        for l in range(S - 1):
            conv(up + f"conv1s.{l}.custom_conv", feat(l), 2 * feat(l), k, k)
        if b > 0:
            for l in range(S - 1):
                conv(up + f"conv1_1s.{l}.custom_conv", feat(l), 2 * feat(l), 1, 1)
        for l in range(S - 1):
            out.append((up + f"deconvs.{l}.conv.weight", (2 * feat(l), feat(l), k, k)))
            out.append((up + f"deconvs.{l}.conv.bias", (feat(l),)))
    for b in range(num_blocks):
        conv(f"msau_net.end_convs.{b}.custom_conv", n_class, feat_root, 4, 4)
    return out


class _Node(torch.nn.Module):
    """Anonymous container; the module tree only exists to reproduce the reference's state_dict keys."""


class _Plan:
    """One C-side launch plan + its workspace for a fixed (B, H, W)."""

    def __init__(self, cfg: _lib.MsauConfig, B: int, H: int, W: int, device: torch.device):
        self.handle = C.c_void_p()
        self.B, self.H, self.W = B, H, W
        self.device = device
        with torch.cuda.device(device):
            _lib.check(_lib.lib().msau_plan_create(C.byref(cfg), B, H, W, C.byref(self.handle)))
        self.ws: Optional[torch.Tensor] = None
        self.ws_training = False

    def workspace(self, training: bool) -> torch.Tensor:
        if self.ws is None or (training and not self.ws_training):
            n = C.c_size_t()
            _lib.check(_lib.lib().msau_workspace_bytes(self.handle, int(training), C.byref(n)))
            self.ws = None
            self.ws = torch.empty(n.value + 256, dtype=torch.uint8, device=self.device)
            self.ws_training = training
        return self.ws

    def ws_ptr(self, training: bool) -> Tuple[int, int]:
        ws = self.workspace(training)
        base = ws.data_ptr()
        aligned = (base + 255) // 256 * 256
        return aligned, ws.numel() - (aligned - base)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().msau_plan_destroy(self.handle)
        except Exception:
            pass


class _LossBackward(torch.autograd.Function):
    """Graph node so that ``loss.backward()`` (train_chargrid_funsd_msau.py:57) drives the CUDA backward."""

    @staticmethod
    def forward(ctx, anchor, model, labels):
        ctx.model = model
        ctx.labels = labels
        return model._loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.model
        m._assign_grads(scale=grad_out)
        return None, None, None


class MSAUWrapper(torch.nn.Module):
    def __init__(self, channels=1, n_class=2, model_kwargs={}):
        super().__init__()
        self.n_class = n_class
        self.channels = channels
        # hyper-parameters: same keys and defaults as model/model.py:406-419
        self.scale_space_num = model_kwargs.get("scale_space_num", 6)
        self.res_depth = model_kwargs.get("res_depth", 3)
        self.featRoot = model_kwargs.get("featRoot", 8)
        self.filter_size = model_kwargs.get("filter_size", 3)
        self.pool_size = model_kwargs.get("pool_size", 2)
        self.activation_name = model_kwargs.get("activation_name", "relu")
        if self.activation_name != "relu":
            raise NotImplementedError("msau_b200 implements activation_name='relu' only")
        self.model = model_kwargs.get("model", "msau")
        self.num_scales = model_kwargs.get("num_scales", 3)
        self.final_act = model_kwargs.get("final_act", "sigmoid")
        if self.final_act == "sigmoid":
            # the reference constructs torch.nn.Sigmoid(dim=1) here, which raises (model/model.py:428-429)
            torch.nn.Sigmoid(dim=1)
        if self.final_act not in ("softmax", "identity"):
            raise ValueError(f"final_act={self.final_act!r}: the reference supports 'softmax' and 'identity'")
        self.num_blocks = 3

        self._schema = param_schema(channels, n_class, self.scale_space_num, self.res_depth, self.featRoot,
                                    self.filter_size, self.num_blocks)
        self._numel = sum(int(torch.Size(s).numel()) for _, s in self._schema)
        self._flat = torch.zeros(self._numel, dtype=torch.float32)
        self._flat_grad: Optional[torch.Tensor] = None
        self._build_tree()
        self.reset_parameters()
        self._cfg = _lib.MsauConfig(channels, n_class, self.scale_space_num, self.res_depth, self.featRoot,
                                    self.filter_size, self.pool_size, self.num_blocks)
        self._plans: Dict[Tuple[int, int, int, int], _Plan] = {}
        self._last = None            # (plan, x, layout) of the last training forward
        self._loss_value = None
        self._anchor = None
        self._adam = None            # (exp_avg, exp_avg_sq, step, scratch)
        self._table = None           # fp32 [rows, channels] feature table of a box-constant (BERT-grid) batch, layout 3
        self._graphs = {}            # (B, H, W, layout, dtype) -> (CUDAGraph, static input, static class map)
        self._train_graphs = {}      # ... -> (CUDAGraph of forward + loss + backward, static x, static labels, loss, launches)

    # ------------------------------------------------------------------ parameters
    def _build_tree(self):
        self._param_list: List[torch.nn.Parameter] = []
        off = 0
        for key, shape in self._schema:
            n = int(torch.Size(shape).numel())
            parts = key.split(".")
            node = self
            for name in parts[:-1]:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            p = torch.nn.Parameter(self._flat[off:off + n].view(shape))
            node.register_parameter(parts[-1], p)
            self._param_list.append(p)
            off += n

    def _rebind(self):
        """Point every Parameter at its slice of the flat buffer (after device moves / loads)."""
        off = 0
        for p, (_, shape) in zip(self._param_list, self._schema):
            n = p.numel()
            p.data = self._flat[off:off + n].view(shape)
            off += n
        self._flat_grad = None
        self._adam = None
        for p in self._param_list:
            p.grad = None

    def reset_parameters(self, seed: Optional[int] = None):
        """The reference's init distributions: conv/deconv N(0, sqrt(2/(kh*kw*Cin+Cout))), bias N(0.1, 1e-5)
        (model/layers/layers.py:33-36,59-60,130-131,216,227-228); attention 1x1 convs keep torch's Conv2d
        default (attention.py:19-21)."""
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(seed)
        else:
            g.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        with torch.no_grad():
            for p, (key, shape) in zip(self._param_list, self._schema):
                attn = ".attention_block." in key
                if key.endswith(".weight"):
                    if attn:
                        bound = 1.0 / (shape[1] * shape[2] * shape[3]) ** 0.5
                        v = (torch.rand(shape, generator=g) * 2 - 1) * bound
                    else:
                        std = (2.0 / (shape[2] * shape[3] * shape[1] + shape[0])) ** 0.5
                        v = torch.randn(shape, generator=g) * std
                else:
                    if attn:
                        v = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
                    else:
                        v = 0.1 + 1e-5 * torch.randn(shape, generator=g)
                p.copy_(v.to(p.device))

    def _apply(self, fn, recurse=True):
        # keep the flat-buffer invariant across .to()/.cuda()/.float(): move the buffer, re-create the views
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise TypeError("msau_b200 parameters are fp32 only")
        self._flat = new_flat.contiguous()
        self._rebind()
        self._plans = {}
        self._graphs = {}
        self._train_graphs = {}
        self._last = None
        self._table = None
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        keys = [k for k, _ in self._schema]
        missing = [k for k in keys if k not in state_dict]
        unexpected = [k for k in state_dict if k not in set(keys)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:3]}..., unexpected {unexpected[:3]}...")
        with torch.no_grad():
            for p, (k, shape) in zip(self._param_list, self._schema):
                if k in state_dict:
                    src = state_dict[k]
                    if tuple(src.shape) != tuple(shape):
                        raise RuntimeError(f"size mismatch for {k}: {tuple(src.shape)} vs {tuple(shape)}")
                    p.copy_(src.to(device=p.device, dtype=torch.float32))
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def save(self, path):
        torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in self.state_dict().items()), path)

    def load_weights(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu"))

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_grads(self) -> torch.Tensor:
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
        return self._flat_grad

    # ------------------------------------------------------------------ engine plumbing
    def _plan(self, B: int, H: int, W: int) -> _Plan:
        if not self._flat.is_cuda:
            raise _lib.MsauError("msau_b200 has no CPU path: move the model to a CUDA device first (.cuda() / .to('cuda'))")
        key = (B, H, W, self._flat.device.index or 0)
        pl = self._plans.get(key)
        if pl is None:
            pl = _Plan(self._cfg, B, H, W, self._flat.device)
            n = int(_lib.lib().msau_param_count(pl.handle))
            assert n == self._numel, (n, self._numel)
            self._plans[key] = pl
        return pl

    def _check_input(self, x: torch.Tensor, layout: int):
        if not x.is_cuda or x.device != self._flat.device:
            raise _lib.MsauError("input must be a CUDA tensor on the model's device (no CPU fallback)")
        if layout in (2, 3):
            if x.dtype != torch.int16 or x.dim() != 3:
                raise TypeError(f"layout {layout} expects an int16 [B, H, W] id map (-1 = empty pixel)")
            if layout == 3 and self._table is None:
                raise _lib.MsauError("layout 3 (row-id map of a box-constant grid) needs set_feature_table() first")
            return tuple(x.shape)
        if x.dtype != torch.float32:
            raise TypeError("input must be float32")
        if layout == 0:
            B, Cc, H, W = x.shape
            ok = Cc == self.channels
        else:
            B, H, W, Cc = x.shape
            ok = Cc == (self.channels + 3) // 4 * 4
        if not ok:
            raise ValueError(f"input has {Cc} channels, model expects {self.channels}")
        return B, H, W

    def _run_forward(self, x: torch.Tensor, layout: int, training: bool, want_probs: bool, want_argmax: bool = False,
                     want_logits: bool = True):
        x = x.contiguous()
        B, H, W = self._check_input(x, layout)
        pl = self._plan(B, H, W)
        ws, ws_bytes = pl.ws_ptr(training)
        dev = x.device
        if layout == 3:
            _lib.check(_lib.lib().msau_plan_set_feature_table(pl.handle, self._table.data_ptr(), self._table.shape[0]))
        logits = torch.empty((B, self.n_class, H, W), dtype=torch.float32, device=dev) if want_logits else None
        aux = torch.empty((B, self.n_class, H, W), dtype=torch.float32, device=dev) if want_logits else None
        probs = torch.empty((B, self.n_class, H, W), dtype=torch.float32, device=dev) if want_probs else None
        amax = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if want_argmax else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().msau_forward(pl.handle, x.data_ptr(), layout, self._flat.data_ptr(), ws, ws_bytes, int(training),
                                               _lib.ptr(logits), _lib.ptr(aux), _lib.ptr(probs), _lib.ptr(amax),
                                               _lib.current_stream()))
        return pl, x, logits, aux, probs, amax

    def set_feature_table(self, table: torch.Tensor):
        """Feature vectors of the boxes of the next batch(es) for ``layout=3`` inputs (data_generator_funsd_bert.py:64-93
        paints ``feats[i]`` over the whole rectangle of cell i, so a BERT-grid page is an int16 map of table rows -- what
        ``raster.raster_features(..., layout="ids")`` writes -- plus this [rows, channels] table; the dense 768-channel grid is
        never built).  ``table`` is converted to fp32 exactly like ``torch.Tensor(float64 ndarray)`` (:82-84)."""
        if not self._flat.is_cuda:
            raise _lib.MsauError("msau_b200 has no CPU path: move the model to a CUDA device first")
        t = torch.as_tensor(table).to(device=self._flat.device, dtype=torch.float32).contiguous()
        if t.dim() != 2 or t.shape[1] != self.channels or not 1 <= t.shape[0] <= 32767:
            raise ValueError(f"feature table must be [1..32767, {self.channels}], got {tuple(t.shape)}")
        self._table = t

    # ------------------------------------------------------------------ reference API
    def forward(self, inp):
        """model/model.py:435-437 -> (predictor(logits), logits, aux_logits), each [B, n_class, H, W]."""
        track = self.training and torch.is_grad_enabled()
        want_probs = self.final_act == "softmax"
        pl, x, logits, aux, probs, _ = self._run_forward(inp, 0, track, want_probs)
        self._last = (pl, x, 0, logits.data_ptr(), aux.data_ptr()) if track else None
        return (probs if want_probs else logits), logits, aux

    def _eval_workspace_bytes(self, B: int, H: int, W: int) -> int:
        n = C.c_size_t()
        _lib.check(_lib.lib().msau_workspace_bytes(self._plan(B, H, W).handle, 0, C.byref(n)))
        return n.value

    def predict_classes(self, inp, layout: int = 0, pages_per_call: Optional[int] = None):
        """argmax over classes, uint8 [B,H,W] (train_chargrid_funsd_msau.py:133-136, kv_model.py:162) without
        materialising logits on the host.  Large batches (BASELINE.json config 5: 64 pages of 1024x768 per GPU) run in
        chunks of ``pages_per_call`` pages; by default the largest power-of-two chunk whose activation workspace fits in
        80 % of the free device memory."""
        B, H, W = self._check_input(inp, layout)
        if pages_per_call is None:
            free = torch.cuda.mem_get_info(inp.device)[0]
            pages_per_call = B
            while pages_per_call > 1 and self._eval_workspace_bytes(pages_per_call, H, W) > 0.8 * free:
                pages_per_call = (pages_per_call + 1) // 2
        if pages_per_call >= B:
            return self._run_forward(inp, layout, False, False, want_argmax=True, want_logits=False)[5]
        out = torch.empty((B, H, W), dtype=torch.uint8, device=inp.device)
        for b0 in range(0, B, pages_per_call):
            chunk = inp[b0:b0 + pages_per_call]
            out[b0:b0 + chunk.shape[0]] = self._run_forward(chunk, layout, False, False, want_argmax=True, want_logits=False)[5]
        return out

    def predict_classes_graph(self, inp, layout: int = 0):
        """``predict_classes`` replayed from a CUDA graph: the ~150 kernel launches of one inference forward are captured once
        per (B, H, W, layout) and replayed with a single launch, which takes the host out of the latency of small batches
        (BASELINE.json config 1: one 512x512 page).  The input is copied into the graph's static buffer, the class map comes
        back as a fresh tensor; parameters are read from the flat buffer at every replay, so ``load_state_dict`` / training
        steps between calls are seen."""
        B, H, W = self._check_input(inp, layout)
        key = (B, H, W, layout, inp.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            static_in = inp.contiguous().clone()
            side = torch.cuda.Stream(device=inp.device)
            side.wait_stream(torch.cuda.current_stream(inp.device))
            with torch.cuda.stream(side):            # warm-up outside capture: plan, workspace, kernel attributes
                for _ in range(2):
                    self._run_forward(static_in, layout, False, False, want_argmax=True, want_logits=False)
            torch.cuda.current_stream(inp.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may call CUDA
                static_out = self._run_forward(static_in, layout, False, False, want_argmax=True, want_logits=False)[5]
            ent = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = ent
        static_in.copy_(inp)
        graph.replay()
        return static_out.clone()

    def _backward_from_last(self, labels: torch.Tensor, loss_scale: float = 1.0) -> torch.Tensor:
        if self._last is None:
            raise _lib.MsauError("loss/backward needs a preceding forward() in training mode with grad enabled")
        pl, x, layout = self._last[:3]
        if labels.dim() == 2:
            labels = labels.unsqueeze(0)
        if tuple(labels.shape) != (pl.B, pl.H, pl.W):
            raise ValueError(f"label_mask shape {tuple(labels.shape)} != {(pl.B, pl.H, pl.W)}")
        if labels.dtype == torch.uint8:
            ld = 0
        else:
            labels = labels.to(torch.int64)
            ld = 1
        labels = labels.to(x.device).contiguous()
        ws, ws_bytes = pl.ws_ptr(True)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().msau_loss_backward(pl.handle, x.data_ptr(), layout, labels.data_ptr(), ld, float(loss_scale), ws,
                                                     ws_bytes, loss.data_ptr(), self.flat_grads.data_ptr(),
                                                     _lib.current_stream()))
        return loss

    def loss(self, out_grid, out_grid_aux, label_mask):
        """model/model.py:446-459 (mean over pages for B>1).  ``out_grid`` / ``out_grid_aux`` must be the logits
        returned by the immediately preceding ``forward``; the value and the gradients come from one fused
        CUDA pass, the returned scalar carries a grad_fn so ``loss.backward()`` fills ``p.grad``."""
        if self._last is None or out_grid.data_ptr() != self._last[3] or out_grid_aux.data_ptr() != self._last[4]:
            raise _lib.MsauError("loss(): pass the logits returned by the preceding forward() (training mode, grad enabled)")
        self._loss_value = self._backward_from_last(label_mask)
        if self._anchor is None or self._anchor.device != self._flat.device:
            self._anchor = torch.zeros((), device=self._flat.device, requires_grad=True)
        return _LossBackward.apply(self._anchor, self, label_mask)

    def _live_mask(self) -> List[bool]:
        dead = f"msau_net.blocks.{self.num_blocks - 1}.downsamplingblock.layer_attentions."
        return [not k.startswith(dead) for k, _ in self._schema]

    def _assign_grads(self, scale=None):
        """p.grad <- views of the flat gradient buffer.  The last block's attention parameters keep grad=None
        exactly like the reference (their output is never read)."""
        g = self.flat_grads
        if scale is not None:
            g.mul_(scale.to(g.dtype))
        off = 0
        for p, (_, shape), live in zip(self._param_list, self._schema, self._live_mask()):
            n = p.numel()
            if live:
                view = g[off:off + n].view(shape)
                if p.grad is None:
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    p.grad.add_(view)
            off += n

    # ------------------------------------------------------------------ fused training step
    def train_step(self, x: torch.Tensor, labels: torch.Tensor, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                   max_norm: float = 1.0, layout: int = 0, process_group=None, world_size: int = 1, use_graph: bool = False):
        """One step of train_chargrid_funsd_msau.py:45-59 on a batch of pages: forward, masked CE (main + aux),
        backward, [NCCL all-reduce of the flat gradient over ``process_group``], clip_grad_norm(max_norm),
        Adam.  Returns the (local-batch) loss as a 0-d CUDA tensor; nothing synchronises with the host.
        ``use_graph``: forward + loss + backward (~420 kernel launches on two streams) are captured once per input shape into
        a CUDA graph and replayed with one launch per step, so the step no longer depends on how fast the host can enqueue
        kernels; the all-reduce and the clip + Adam kernels (which take the step count as an argument) stay eager.
        ``use_graph="static"``: the caller keeps feeding the same (refilled) input tensors, so the graph reads them in place."""
        if use_graph and layout != 3:
            loss = self._fwd_bwd_graph(x, labels, layout, world_size, use_graph == "static")
        else:
            pl, xc, _, _, _, _ = self._run_forward(x, layout, True, False, want_logits=False)
            self._last = (pl, xc, layout, 0, 0)
            loss = self._backward_from_last(labels, loss_scale=1.0 / world_size)
        if world_size > 1:
            torch.distributed.all_reduce(self.flat_grads, group=process_group)
        self.adam_step(lr, betas, eps, max_norm)
        return loss

    def _fwd_bwd_graph(self, x: torch.Tensor, labels: torch.Tensor, layout: int, world_size: int, static: bool) -> torch.Tensor:
        """forward(training) + loss + backward as one CUDA-graph launch.  The graph reads private copies of the inputs (every
        call copies x / labels into them: 12 MB for an id-map batch of 16 pages); with ``static`` it is bound to the caller's
        own tensors instead, for loops that refill the same buffers in place (no copy; a dense 1.6 GB batch stays put)."""
        B, H, W = self._check_input(x, layout)
        if labels.dim() == 2:
            labels = labels.unsqueeze(0)
        if labels.dtype != torch.uint8:
            labels = labels.to(torch.int64)
        labels = labels.to(x.device)
        key = (B, H, W, layout, x.dtype, labels.dtype, world_size, static)
        ent = self._train_graphs.get(key)
        if ent is None:
            if static:
                if not (x.is_contiguous() and labels.is_contiguous()):
                    raise ValueError("use_graph='static' needs contiguous input tensors")
                sx, sl = x, labels                        # the caller promises to keep feeding these very tensors
            else:
                sx, sl = x.contiguous().clone(), labels.contiguous().clone()
            dev = x.device

            def fwd_bwd():
                pl, xc, _, _, _, _ = self._run_forward(sx, layout, True, False, want_logits=False)
                self._last = (pl, xc, layout, 0, 0)
                return self._backward_from_last(sl, loss_scale=1.0 / world_size)

            self.flat_grads                               # allocate outside the capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                 # warm-up: plan, workspace, side stream / events, kernel attributes
                fwd_bwd()
            torch.cuda.current_stream(dev).wait_stream(side)
            n0 = _lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may call CUDA
                sloss = fwd_bwd()
            ent = self._train_graphs[key] = (graph, sx, sl, sloss, _lib.launch_count() - n0)
        graph, sx, sl, sloss, n_launch = ent
        if static:
            if x.data_ptr() != sx.data_ptr() or labels.data_ptr() != sl.data_ptr():
                raise _lib.MsauError("use_graph='static': the step must be fed the tensors the graph was captured with")
        else:
            sx.copy_(x)
            sl.copy_(labels)
        graph.replay()
        _lib.lib().msau_launch_count_add(n_launch)
        return sloss.clone()

    def adam_step(self, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0):
        if self._adam is None:
            self._adam = [torch.zeros_like(self._flat), torch.zeros_like(self._flat), 0,
                          torch.empty(2048, dtype=torch.float32, device=self._flat.device),
                          torch.zeros((), dtype=torch.float32, device=self._flat.device)]
        m, v, step, scratch, total = self._adam
        step += 1
        self._adam[2] = step
        with torch.cuda.device(self._flat.device):
            _lib.check(_lib.lib().msau_clip_adam_step(self._flat.data_ptr(), self.flat_grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                      self._numel, step, lr, betas[0], betas[1], eps, max_norm, scratch.data_ptr(),
                                                      total.data_ptr(), _lib.current_stream()))
        return total


MSAU = MSAUWrapper
