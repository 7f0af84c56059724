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
        self.generation = 0          # bumped whenever the workspace is (re)allocated: CUDA graphs captured before are stale

    def workspace(self, training: bool) -> torch.Tensor:
        if self.ws is None or (training and not self.ws_training):
            # an inference-sized workspace grows when the plan is first used for training.  Captured CUDA graphs have the old
            # pointers baked in, so every graph entry records the generation it was captured at and is re-captured (never
            # replayed) once it differs -- see MSAUWrapper._graph_entry
            n = C.c_size_t()
            _lib.check(_lib.lib().msau_workspace_bytes(self.handle, int(training), C.byref(n)))
            self.ws = None
            self.ws = torch.empty(n.value + 256, dtype=torch.uint8, device=self.device)
            self.ws_training = training
            self.generation += 1
        return self.ws

    def set_option(self, name: str, value: int) -> None:
        _lib.check(_lib.lib().msau_plan_set_option(self.handle, name.encode(), int(value)))

    def error_flags(self) -> int:
        v = C.c_int(0)
        _lib.check(_lib.lib().msau_plan_error_flags(self.handle, C.byref(v)))
        return v.value

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
        self._adam = None            # fused optimiser state: dict(kind, s1, s2, step_dev, scratch, total)
        self._options: Dict[str, int] = {}   # engine options of this model's plans (msau_plan_set_option)
        self._grad_new: Optional[torch.Tensor] = None   # what the backward kernels of loss() wrote (accumulated into flat_grads)
        self._acc = None             # device int32[2]: {correct, kept} of the last loss (masked accuracy)
        self._loss_main = None
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
        self._grad_new = None
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
        if new_flat.device == self._flat.device and new_flat.data_ptr() == self._flat.data_ptr():
            return self                                  # no-op move (.cuda() on a CUDA model): plans, graphs, optimiser state stay
        self._flat = new_flat.contiguous()
        self._rebind()
        if self._adam is not None:                       # the optimiser state follows the parameters
            for k in ("s1", "s2", "step_dev", "scratch", "total"):
                self._adam[k] = self._adam[k].to(self._flat.device)
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
    def set_option(self, name: str, value: int) -> None:
        """Engine option of THIS model's plans (msau_plan_set_option; include/msau_b200.h lists the names).  Options are
        per plan, so two models with different options can run side by side; captured CUDA graphs are dropped."""
        self._options[name] = int(value)
        for pl in self._plans.values():
            pl.set_option(name, value)
        self._graphs = {}
        self._train_graphs = {}

    def _plan(self, B: int, H: int, W: int) -> _Plan:
        if not self._flat.is_cuda:
            raise _lib.MsauError("msau_b200 has no CPU path: move the model to a CUDA device first (.cuda() / .to('cuda'))")
        key = (B, H, W, self._flat.device.index or 0)
        pl = self._plans.get(key)
        if pl is None:
            pl = _Plan(self._cfg, B, H, W, self._flat.device)
            n = int(_lib.lib().msau_param_count(pl.handle))
            assert n == self._numel, (n, self._numel)
            for name, value in self._options.items():
                pl.set_option(name, value)
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

    def _graph_entry(self, table: dict, key, pl_shape):
        """A captured graph is only valid for the workspace it was captured with: entries remember the plan's workspace
        generation and are dropped once the workspace has been reallocated (inference plan later used for training)."""
        ent = table.get(key)
        if ent is not None and ent[0] != self._plan(*pl_shape).generation:
            del table[key]
            ent = None
        return ent

    def predict_classes_graph(self, inp, layout: int = 0):
        """``predict_classes`` replayed from a CUDA graph: the ~150 kernel launches of one inference forward are captured once
        per (B, H, W, layout) and replayed with a single launch, which takes the host out of the latency of small batches
        (BASELINE.json config 1: one 512x512 page).  The input is copied into the graph's static buffer, the class map comes
        back as a fresh tensor; parameters are read from the flat buffer at every replay, so ``load_state_dict`` / training
        steps between calls are seen.  ``layout=3`` is refused: the feature table's pointer and row count would be baked into
        the graph while ``set_feature_table`` swaps the table per batch (use ``predict_classes``)."""
        if layout == 3:
            raise _lib.MsauError("predict_classes_graph does not take layout 3 (the per-batch feature table cannot be captured); "
                                 "use predict_classes")
        B, H, W = self._check_input(inp, layout)
        key = (B, H, W, layout, inp.dtype)
        ent = self._graph_entry(self._graphs, key, (B, H, W))
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
            ent = self._graphs[key] = (self._plan(B, H, W).generation, graph, static_in, static_out)
        _, graph, static_in, static_out = ent
        static_in.copy_(inp)
        graph.replay()
        return static_out.clone()

    # ------------------------------------------------------------------ loss + backward
    @staticmethod
    def _label_tensor(labels: torch.Tensor, device) -> Tuple[torch.Tensor, int]:
        if labels.dim() == 2:
            labels = labels.unsqueeze(0)
        if labels.dtype == torch.uint8:
            ld = 0
        else:
            labels = labels.to(torch.int64)
            ld = 1
        return labels.to(device).contiguous(), ld

    def _backward_from_last(self, labels: torch.Tensor, loss_scale: float = 1.0, grads: Optional[torch.Tensor] = None,
                            loss_spec: Optional[dict] = None, labels_aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        """msau_loss_backward_ex on the plan of the last training forward.  ``loss_spec`` None = MSAUWrapper.loss
        (model/model.py:446-459); dict(mode=1, weight_main, weight_aux, class_weights) = UNetLoss (model/training/cost.py:35-65)."""
        if self._last is None:
            raise _lib.MsauError("loss/backward needs a preceding forward() in training mode with grad enabled")
        pl, x, layout = self._last[:3]
        labels, ld = self._label_tensor(labels, x.device)
        if tuple(labels.shape) != (pl.B, pl.H, pl.W):
            raise ValueError(f"label_mask shape {tuple(labels.shape)} != {(pl.B, pl.H, pl.W)}")
        la_ptr = 0
        if labels_aux is not None:
            labels_aux, ld2 = self._label_tensor(labels_aux, x.device)
            if ld2 != ld or labels_aux.shape != labels.shape:
                raise ValueError("aux targets must have the shape and dtype of the main targets")
            la_ptr = labels_aux.data_ptr()
        spec = _lib.MsauLossSpec(0, 1.0, 1.0, None)
        cw = None
        if loss_spec:
            spec.mode = int(loss_spec.get("mode", 1))
            spec.weight_main = float(loss_spec.get("weight_main", 0.5))
            spec.weight_aux = float(loss_spec.get("weight_aux", 0.5))
            if loss_spec.get("class_weights") is not None:
                w = [float(v) for v in loss_spec["class_weights"]]
                if len(w) != self.n_class:
                    raise ValueError(f"class_weights needs {self.n_class} entries")
                cw = (C.c_float * len(w))(*w)
                spec.h_class_weights = C.cast(cw, C.POINTER(C.c_float))
        ws, ws_bytes = pl.ws_ptr(True)
        dev = x.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        self._loss_main = torch.empty((), dtype=torch.float32, device=dev)
        self._acc = torch.empty(2, dtype=torch.int32, device=dev)
        g = self.flat_grads if grads is None else grads
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().msau_loss_backward_ex(pl.handle, x.data_ptr(), layout, labels.data_ptr(), la_ptr, ld, C.byref(spec),
                                                        float(loss_scale), ws, ws_bytes, loss.data_ptr(), self._loss_main.data_ptr(),
                                                        self._acc.data_ptr(), g.data_ptr(), _lib.current_stream()))
        return loss

    def check_labels(self) -> None:
        """Raises if any label seen by a loss call since the last check was outside [0, n_class) (torch's CrossEntropyLoss
        raises at once; here such pixels contribute nothing and a sticky device flag records them).  Synchronises."""
        bad = 0
        for pl in self._plans.values():
            bad |= pl.error_flags()
        if bad & 1:
            raise IndexError(f"label_mask holds a class index outside [0, {self.n_class})")

    def last_accuracy(self) -> float:
        """Masked pixel accuracy of the last loss call: argmax(logits) == label over label != 0 (UNetLoss cost.py:44-50;
        evaluate() train_chargrid_funsd_msau.py:133-159), counted inside the loss kernel.  One 8-byte D2H read."""
        if self._acc is None:
            raise _lib.MsauError("last_accuracy() needs a preceding loss / train_step")
        c, k = (int(v) for v in self._acc.tolist())
        return c / k if k else float("nan")

    def _scratch_grads(self) -> torch.Tensor:
        if self._grad_new is None or self._grad_new.device != self._flat.device:
            self._grad_new = torch.zeros_like(self._flat)
        return self._grad_new

    def loss(self, out_grid, out_grid_aux, label_mask):
        """model/model.py:446-459 (mean over pages for B>1).  ``out_grid`` / ``out_grid_aux`` must be the logits
        returned by the immediately preceding ``forward``; the value and the gradients come from one fused
        CUDA pass, the returned scalar carries a grad_fn so ``loss.backward()`` fills ``p.grad`` -- accumulating into an
        existing ``p.grad`` like autograd does when ``zero_grad`` was not called.  A label outside [0, n_class) raises
        IndexError (one flag read per call; the fused ``train_step`` path leaves that check to ``check_labels()``)."""
        if self._last is None or out_grid.data_ptr() != self._last[3] or out_grid_aux.data_ptr() != self._last[4]:
            raise _lib.MsauError("loss(): pass the logits returned by the preceding forward() (training mode, grad enabled)")
        self._loss_value = self._backward_from_last(label_mask, grads=self._scratch_grads())
        self._last[0].error_flags() and self._raise_bad_label()
        if self._anchor is None or self._anchor.device != self._flat.device:
            self._anchor = torch.zeros((), device=self._flat.device, requires_grad=True)
        return _LossBackward.apply(self._anchor, self, label_mask)

    def _raise_bad_label(self):
        raise IndexError(f"label_mask holds a class index outside [0, {self.n_class})")

    def unet_loss(self, logits, tgt, aux_logits=None, aux_tgt=None, class_weights=None, weight_main: float = 0.5,
                  weight_aux: float = 0.5):
        """UNetLoss.forward (model/training/cost.py:35-65) on the logits of the preceding ``forward``: targets are one-hot
        [B, n_class, H, W] (arg-max taken on the device, :41,:52) or already class maps [B, H, W]; returns
        ``(acc, loss, final_loss)`` like the reference -- ``loss`` carries a grad_fn, ``acc`` is a Python float (the reference
        computes it on the host too), ``final_loss`` the main head's own term (None without aux logits)."""
        if self._last is None or logits.data_ptr() != self._last[3]:
            raise _lib.MsauError("unet_loss(): pass the logits returned by the preceding forward() (training mode, grad enabled)")
        if aux_logits is not None and aux_logits.data_ptr() != self._last[4]:
            raise _lib.MsauError("unet_loss(): aux_logits must be the aux logits of the preceding forward()")
        t = self.onehot_argmax(tgt) if tgt.dim() == 4 else tgt
        ta = None
        if aux_logits is not None and aux_tgt is not None:
            ta = self.onehot_argmax(aux_tgt) if aux_tgt.dim() == 4 else aux_tgt
            if ta.dtype != t.dtype:
                ta = ta.to(t.dtype)
        spec = dict(mode=1, weight_main=weight_main if aux_logits is not None else 1.0,
                    weight_aux=weight_aux if aux_logits is not None else 0.0, class_weights=class_weights)
        self._loss_value = self._backward_from_last(t, grads=self._scratch_grads(), loss_spec=spec, labels_aux=ta)
        self._last[0].error_flags() and self._raise_bad_label()
        if self._anchor is None or self._anchor.device != self._flat.device:
            self._anchor = torch.zeros((), device=self._flat.device, requires_grad=True)
        loss = _LossBackward.apply(self._anchor, self, t)
        final = self._loss_main.clone() if aux_logits is not None else None
        return self.last_accuracy(), loss, final

    def onehot_argmax(self, tgt: torch.Tensor) -> torch.Tensor:
        """torch.argmax(tgt, dim=1) of one-hot targets as a uint8 class map, on the device (cost.py:41,52)."""
        dt = {torch.uint8: 0, torch.int64: 1, torch.float32: 2}.get(tgt.dtype)
        if dt is None:
            tgt, dt = tgt.to(torch.int64), 1
        tgt = tgt.to(self._flat.device).contiguous()
        B, Cc, H, W = tgt.shape
        out = torch.empty((B, H, W), dtype=torch.uint8, device=tgt.device)
        with torch.cuda.device(tgt.device):
            _lib.check(_lib.lib().msau_onehot_argmax(tgt.data_ptr(), dt, B, Cc, H * W, 0, out.data_ptr(), _lib.current_stream()))
        return out

    def _live_mask(self) -> List[bool]:
        dead = f"msau_net.blocks.{self.num_blocks - 1}.downsamplingblock.layer_attentions."
        return [not k.startswith(dead) for k, _ in self._schema]

    def _assign_grads(self, scale=None):
        """loss.backward(): p.grad (views of the flat gradient buffer) += what the backward kernels wrote.  The last block's
        attention parameters keep grad=None exactly like the reference (their output is never read)."""
        new = self._scratch_grads()
        if scale is not None:
            new.mul_(scale.to(new.dtype))
        acc = self.flat_grads
        off = 0
        for p, (_, shape), live in zip(self._param_list, self._schema, self._live_mask()):
            n = p.numel()
            if live:
                view = acc[off:off + n].view(shape)
                fresh = new[off:off + n].view(shape)
                if p.grad is None:
                    view.copy_(fresh)
                    p.grad = view
                else:
                    p.grad.add_(fresh)
            off += n

    # ------------------------------------------------------------------ fused training step
    _OPT_KIND = {"adam": 0, "rmsprop": 1, "momentum": 2, "sgd": 2}

    def train_step(self, x: torch.Tensor, labels: torch.Tensor, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                   max_norm: float = 1.0, layout: int = 0, process_group=None, world_size: int = 1, use_graph=False,
                   optimizer: str = "adam", weight_decay: float = 0.0, loss_spec: Optional[dict] = None,
                   labels_aux: Optional[torch.Tensor] = None, graph_tail: bool = True):
        """One step of train_chargrid_funsd_msau.py:45-59 on a batch of pages: forward, masked CE (main + aux),
        backward, [NCCL all-reduce of the flat gradient over ``process_group``], clip_grad_norm(max_norm),
        Adam.  Returns the (local-batch) loss as a 0-d CUDA tensor; nothing synchronises with the host.
        ``optimizer`` / ``weight_decay`` / ``loss_spec`` / ``labels_aux`` select the alternative trainer's step instead
        (model/training/trainer.py:122-137: UNetLoss, RMSprop by default, no clipping -> ``max_norm=0``).
        ``use_graph``: the whole step -- forward + loss + backward (~400 launches on two streams) and, with ``graph_tail``, the
        all-reduce and the clip + optimiser kernels (the step count lives in a device counter) -- is captured once per input
        shape into a CUDA graph and replayed with one launch per step.
        ``use_graph="static"``: the caller keeps feeding the same (refilled) input tensors, so the graph reads them in place."""
        opt_args = (optimizer, float(lr), tuple(betas), float(eps), float(max_norm), float(weight_decay))
        if use_graph and layout != 3:
            return self._step_graph(x, labels, layout, world_size, process_group, use_graph == "static", opt_args, loss_spec, labels_aux,
                                    graph_tail)
        pl, xc, _, _, _, _ = self._run_forward(x, layout, True, False, want_logits=False)
        self._last = (pl, xc, layout, 0, 0)
        loss = self._backward_from_last(labels, loss_scale=1.0 / world_size, loss_spec=loss_spec, labels_aux=labels_aux)
        if world_size > 1:
            torch.distributed.all_reduce(self.flat_grads, group=process_group)
        self.optimizer_step(*opt_args)
        return loss

    def _step_graph(self, x, labels, layout, world_size, process_group, static, opt_args, loss_spec, labels_aux, graph_tail):
        """The train step as one CUDA-graph launch.  The graph reads private copies of the inputs (every call copies x / labels
        into them: 12 MB for an id-map batch of 16 pages); with ``static`` it is bound to the caller's own tensors instead, for
        loops that refill the same buffers in place (no copy; a dense 1.6 GB batch stays put)."""
        B, H, W = self._check_input(x, layout)
        labels, _ = self._label_tensor(labels, x.device)
        if labels_aux is not None:
            labels_aux, _ = self._label_tensor(labels_aux, x.device)
        spec_key = None if not loss_spec else tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in loss_spec.items()))
        key = (B, H, W, layout, x.dtype, labels.dtype, world_size, static, opt_args if graph_tail else None, spec_key,
               labels_aux is not None, graph_tail)
        ent = self._graph_entry(self._train_graphs, key, (B, H, W))
        if ent is None:
            if static:
                if not (x.is_contiguous() and labels.is_contiguous()):
                    raise ValueError("use_graph='static' needs contiguous input tensors")
                sx, sl, sa = x, labels, labels_aux           # the caller promises to keep feeding these very tensors
            else:
                sx, sl = x.contiguous().clone(), labels.contiguous().clone()
                sa = labels_aux.contiguous().clone() if labels_aux is not None else None
            dev = x.device

            def body():
                pl, xc, _, _, _, _ = self._run_forward(sx, layout, True, False, want_logits=False)
                self._last = (pl, xc, layout, 0, 0)
                loss = self._backward_from_last(sl, loss_scale=1.0 / world_size, loss_spec=loss_spec, labels_aux=sa)
                if graph_tail:
                    if world_size > 1:
                        torch.distributed.all_reduce(self.flat_grads, group=process_group)
                    self.optimizer_step(*opt_args)
                return loss

            self.flat_grads                               # allocate outside the capture
            self._opt_state(opt_args[0])
            # warm-up on a side stream: plan, workspace, side stream / events, kernel attributes, NCCL communicator.  The warm-up
            # must not change the model: parameters, optimiser state and step counter are restored afterwards
            keep = None
            if graph_tail:
                st = self._adam
                keep = (self._flat.clone(), st["s1"].clone(), st["s2"].clone(), st["step_dev"].clone())
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream(dev).wait_stream(side)
            if keep is not None:
                self._flat.copy_(keep[0]); st["s1"].copy_(keep[1]); st["s2"].copy_(keep[2]); st["step_dev"].copy_(keep[3])
            n0 = _lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may call CUDA
                sloss = body()
            pl = self._plan(B, H, W)
            ent = self._train_graphs[key] = (pl.generation, graph, sx, sl, sa, sloss, _lib.launch_count() - n0, self._acc, self._loss_main)
        _, graph, sx, sl, sa, sloss, n_launch, acc, lmain = ent
        if static:
            if x.data_ptr() != sx.data_ptr() or labels.data_ptr() != sl.data_ptr():
                raise _lib.MsauError("use_graph='static': the step must be fed the tensors the graph was captured with")
        else:
            sx.copy_(x)
            sl.copy_(labels)
            if sa is not None:
                sa.copy_(labels_aux)
        graph.replay()
        self._acc, self._loss_main = acc, lmain
        _lib.lib().msau_launch_count_add(n_launch)
        if not graph_tail:
            if world_size > 1:
                torch.distributed.all_reduce(self.flat_grads, group=process_group)
            self.optimizer_step(*opt_args)
        return sloss.clone()

    # ------------------------------------------------------------------ fused optimisers
    def _opt_state(self, optimizer: str) -> dict:
        kind = self._OPT_KIND.get(optimizer)
        if kind is None:
            raise ValueError(f"optimizer={optimizer!r}: expected one of {sorted(self._OPT_KIND)}")
        if self._adam is None or self._adam["kind"] != kind or self._adam["s1"].device != self._flat.device:
            dev = self._flat.device
            self._adam = dict(kind=kind, s1=torch.zeros_like(self._flat), s2=torch.zeros_like(self._flat),
                              step_dev=torch.zeros((), dtype=torch.int32, device=dev),
                              scratch=torch.empty(2048, dtype=torch.float32, device=dev),
                              total=torch.zeros((), dtype=torch.float32, device=dev))
        return self._adam

    def optimizer_step(self, optimizer: str = "adam", lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0,
                       weight_decay: float = 0.0):
        """clip_grad_norm(max_norm) + optimiser update on the flat buffers (train_chargrid_funsd_msau.py:58-59 with "adam";
        model/training/optimizer.py:4-30 with "rmsprop" (betas[0] = alpha) or "momentum" (betas[0] = momentum)).  The step count
        is a device counter, so the call is CUDA-graph capturable.  Returns the pre-clip gradient norm (0-d CUDA tensor)."""
        st = self._opt_state(optimizer)
        with torch.cuda.device(self._flat.device):
            _lib.check(_lib.lib().msau_optimizer_step(st["kind"], self._flat.data_ptr(), self.flat_grads.data_ptr(), st["s1"].data_ptr(),
                                                      st["s2"].data_ptr(), self._numel, 0, st["step_dev"].data_ptr(), lr, betas[0], betas[1],
                                                      eps, weight_decay, max_norm, st["scratch"].data_ptr(), st["total"].data_ptr(),
                                                      _lib.current_stream()))
        return st["total"]

    def adam_step(self, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0):
        return self.optimizer_step("adam", lr, betas, eps, max_norm)

    def optimizer_state_dict(self) -> dict:
        """State of the fused optimiser (moments + step) for checkpoints -- what ``save_checkpoint`` pickles with the optimizer
        object in the reference (utils/io_utils.py:83-105).  Keys follow torch: exp_avg / exp_avg_sq (Adam), square_avg (RMSprop),
        momentum_buffer (SGD), step."""
        if self._adam is None:
            return {}
        st = self._adam
        names = {0: ("exp_avg", "exp_avg_sq"), 1: (None, "square_avg"), 2: ("momentum_buffer", None)}[st["kind"]]
        out = dict(kind={0: "adam", 1: "rmsprop", 2: "momentum"}[st["kind"]], step=int(st["step_dev"].item()))
        for name, key in zip(names, ("s1", "s2")):
            if name:
                out[name] = st[key].detach().cpu().clone()
        return out

    def load_optimizer_state_dict(self, sd: dict) -> None:
        if not sd:
            self._adam = None
            return
        st = self._opt_state(sd["kind"])
        names = {0: ("exp_avg", "exp_avg_sq"), 1: (None, "square_avg"), 2: ("momentum_buffer", None)}[st["kind"]]
        for name, key in zip(names, ("s1", "s2")):
            if name:
                st[key].copy_(sd[name].to(st[key].device))
        st["step_dev"].fill_(int(sd["step"]))


MSAU = MSAUWrapper
