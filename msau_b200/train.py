"""Training driver + data-parallel step over the CUDA engine.

Mirrors train_chargrid_funsd_msau.py: ``train(dataset, model_instance, args, ...)`` (:16-118) and
``evaluate(dataset, model, args, ...)`` (:121-163) keep their signatures and loop shape (Adam lr=1e-4,
``clip_grad_norm(params, args.clip)``, per-epoch evaluation), minus the plotting / tensorboard / checkpoint-name
helpers that live in the reference's ``utils/io_utils.py`` (out of the hot path).

Data parallelism (new; the reference is single-device): pages are sharded over ranks,
every rank runs forward + loss + backward on its pages with the gradient pre-scaled by 1/world, ONE NCCL
all-reduce sums the flat 2.55 MB gradient buffer, and every rank applies the identical clip + Adam update, so the
replicas stay bit-identical.  ``shard_pages`` / ``allreduce_flat`` are plain host logic and are exercised on CPU with
the gloo backend in tests/test_dp_gloo.py.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_pages(n_pages: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced page range of ``rank`` (first ``n % world`` ranks take one extra page)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_pages, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def allreduce_flat(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the flat gradient buffer in place (gradients are pre-scaled by 1/world at the loss)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(flat_grad, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat_grad


def dp_gradient(local_grad_fn: Callable[[Sequence[int], float], Tuple[torch.Tensor, torch.Tensor]], n_pages: int,
                rank: int, world_size: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Data-parallel gradient of the batch loss (mean over pages of the per-page loss).

    ``local_grad_fn(page_indices, scale)`` returns (sum of the per-page losses of those pages, flat gradient of
    ``scale * that sum``).  With scale = 1/n_pages the SUM over ranks is the gradient of the batch-mean loss, whatever
    the (possibly uneven) sharding."""
    pages = shard_pages(n_pages, rank, world_size)
    loss_sum, flat = local_grad_fn(list(pages), 1.0 / n_pages)
    allreduce_flat(flat, group)
    loss = loss_sum.clone()
    if torch.distributed.is_available() and torch.distributed.is_initialized() and world_size > 1:
        torch.distributed.all_reduce(loss, group=group)
    return loss / n_pages, flat


class DataParallelTrainer:
    """One process per GPU; ``step`` = MSAUWrapper.train_step with the all-reduce between backward and Adam."""

    def __init__(self, model, lr: float = 1e-4, max_norm: float = 1.0, process_group=None):
        self.model, self.lr, self.max_norm, self.group = model, lr, max_norm, process_group
        dist = torch.distributed
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0

    def broadcast_parameters(self):
        if self.world > 1:
            torch.distributed.broadcast(self.model.flat_params, src=0, group=self.group)

    def step(self, x: torch.Tensor, labels: torch.Tensor, layout: int = 0) -> torch.Tensor:
        """x / labels = THIS rank's pages (equal count on every rank).  Returns the local mean loss (0-d CUDA tensor)."""
        return self.model.train_step(x, labels, lr=self.lr, max_norm=self.max_norm, layout=layout, process_group=self.group,
                                     world_size=self.world)


# ----------------------------------------------------------------------------- page feeder with (H, W) buckets
def page_grid_shape(words: dict) -> Tuple[int, int]:
    """(Hn, Wn) of the grid R1 / R2 rasterise a page into: data_generator_funsd_bert.py:49-61 (min_x/min_y/min_w/min_h and the
    max corner over the page's boxes) and :72-73 / :154-155 (``int((max - min) / min_size) + 1``).  Host arithmetic in fp64,
    the same expressions the device geometry kernel evaluates (msau_raster_geometry), so feeder buckets and device grids agree."""
    x = np.asarray(words["x"], np.float64); y = np.asarray(words["y"], np.float64)
    w = np.asarray(words["w"], np.float64); h = np.asarray(words["h"], np.float64)
    min_x, min_y, min_w, min_h = x.min(), y.min(), w.min(), h.min()
    max_x, max_y = (x + w).max(), (y + h).max()
    return int((max_y - min_y) / min_h) + 1, int((max_x - min_x) / min_w) + 1


class BucketedPageFeeder:
    """Variable-size page feeder (reference: FUNSDMaskDataLoader.getitem dgfb.py:188-230 hands the
    train loop one [1, D, Hn, Wn] page at a time, train_chargrid_funsd_msau.py:45-53, every page with its own grid size).

    Pages are grouped by their EXACT (Hn, Wn) -- padding a page to a larger grid would change the activations near its border
    (bias and LRN make the padded area non-zero), so only same-size pages share a batch -- and every bucket is cut into batches of
    at most ``max_pages`` pages.  A batch is ONE pinned host buffer (``raster.HostBatch``: CSR boxes + char ids + labels), uploaded
    with one asynchronous copy and rasterised on the device; the dense grid never exists on the host.  Iteration order is
    deterministic (``seed`` shuffles batches per epoch, like ``random.shuffle`` of the reference's index list would)."""

    def __init__(self, word_pages: Sequence[dict], line_pages: Sequence[dict], max_pages: int = 16, seed: Optional[int] = None,
                 with_chars: bool = True):
        from . import raster
        if len(word_pages) != len(line_pages):
            raise ValueError("word_pages and line_pages must describe the same pages")
        buckets = {}
        for i, wp in enumerate(word_pages):
            buckets.setdefault(page_grid_shape(wp), []).append(i)
        self.batches: List[Tuple[Tuple[int, int], List[int], "raster.HostBatch", "raster.HostBatch"]] = []
        for shape in sorted(buckets):
            idx = buckets[shape]
            for k in range(0, len(idx), max_pages):
                part = idx[k:k + max_pages]
                self.batches.append((shape, part, raster.HostBatch([word_pages[j] for j in part], with_chars=with_chars),
                                     raster.HostBatch([line_pages[j] for j in part], with_chars=False, with_labels=True)))
        self.seed, self.epoch = seed, 0
        self.n_pages = len(word_pages)

    def __len__(self):
        return len(self.batches)

    def shapes(self):
        return sorted({b[0] for b in self.batches})

    def __iter__(self):
        order = list(range(len(self.batches)))
        if self.seed is not None:
            np.random.RandomState(self.seed + self.epoch).shuffle(order)
        self.epoch += 1
        for k in order:
            yield self.batches[k]


def train_pages(feeder: BucketedPageFeeder, model, feat_table: torch.Tensor, epochs: int = 1, lr: float = 1e-4, max_norm: float = 1.0,
                use_chars: bool = True, process_group=None, world_size: int = 1, use_graph=False, checkpoint_every: int = 10,
                checkpoint_prefix: Optional[str] = None, on_step=None):
    """Fused train loop over host page records: per batch one H2D copy, device rasterisation (R1: int16 channel-id map when the
    table is one-hot, dense NCHW otherwise), ``train_step`` (forward + loss + backward [+ all-reduce] + clip + Adam), and the
    masked accuracy counted inside the loss kernel.  Checkpoints every ``checkpoint_every`` epochs as in
    train_chargrid_funsd_msau.py:100-102 (state_dict) plus the optimiser state the reference's final ``save_checkpoint`` keeps
    (utils/io_utils.py:83-105).  Returns per-epoch dicts(loss, acc, pages)."""
    from . import raster
    dev = model.flat_params.device
    table = feat_table.to(device=dev, dtype=torch.float64)
    one_hot = table.shape[0] == table.shape[1] and bool(torch.equal(table, torch.eye(table.shape[0], dtype=torch.float64, device=dev)))
    history = []
    for epoch in range(epochs):
        model.train()
        tot, correct, kept, pages = 0.0, 0, 0, 0
        accs = []
        for shape, idx, hw, hl in feeder:
            wb = raster.BoxBatch.from_host(hw, dev)
            lb = raster.BoxBatch.from_host(hl, dev)
            geom = wb.geometry()
            if one_hot:
                g = raster.raster_features(wb, geom, table, shape, use_chars, "ids")
                layout = 2
            else:
                g = raster.raster_features(wb, geom, table, shape, use_chars, "nchw")
                layout = 0
            lab = raster.raster_labels(lb, geom, shape)
            loss = model.train_step(g, lab, lr=lr, max_norm=max_norm, layout=layout, process_group=process_group, world_size=world_size,
                                    use_graph=use_graph)
            accs.append((loss, model._acc, len(idx)))
            if on_step is not None:
                on_step(epoch, shape, idx, loss)
        for loss, acc, n in accs:                         # one host read per batch, after the epoch's work is enqueued
            c, k = (int(v) for v in acc.tolist())
            tot += float(loss) * n; correct += c; kept += k; pages += n
        history.append(dict(epoch=epoch, loss=tot / max(pages, 1), acc=correct / kept if kept else float("nan"), pages=pages))
        if checkpoint_prefix and checkpoint_every and (epoch + 1) % checkpoint_every == 0:
            save_checkpoint(model, f"{checkpoint_prefix}_epoch{epoch + 1}.pt")
    if checkpoint_prefix:
        save_checkpoint(model, f"{checkpoint_prefix}_final.pt")
    return history


def save_checkpoint(model, path: str) -> None:
    """state_dict (reference key schema) + fused optimiser state, one file."""
    from collections import OrderedDict
    torch.save(dict(model_state=OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items()),
                    optimizer_state=model.optimizer_state_dict()), path)


def load_checkpoint(model, path: str) -> None:
    ck = torch.load(path, map_location="cpu")
    model.load_state_dict(ck["model_state"] if "model_state" in ck else ck)
    if "optimizer_state" in ck:
        model.load_optimizer_state_dict(ck["optimizer_state"])


# ----------------------------------------------------------------------------- reference-shaped driver
def train(dataset, model_instance, args, same_feat=True, val_dataset=None, test_dataset=None, writer=None, mask_nodes=True):
    """train_chargrid_funsd_msau.py:16-118 without the plotting helpers.  ``dataset`` yields dicts with "mask" [1,C,H,W] and
    "label" [1,H,W] (FUNSDMaskDataLoader.getitem, dgfb.py:216-222).  ``args.ckpt_prefix`` (optional) turns on the reference's
    checkpointing: the state_dict every 10 epochs (:100-102) and model + optimiser at the end (:113-117)."""
    device = torch.device("cuda")
    optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model_instance.parameters()), lr=0.0001)
    model_instance = model_instance.to(device)
    val_accs = []
    prefix = getattr(args, "ckpt_prefix", None)
    for epoch in range(args.num_epochs):
        avg_loss = 0.0
        model_instance.train()
        batch_idx = -1
        for batch_idx, data in enumerate(dataset):
            model_instance.zero_grad()
            V = data["mask"].float().to(device)
            label = data["label"].long().to(device)
            _, ypred, ypred_aux = model_instance(V)
            loss = model_instance.loss(ypred, ypred_aux, label)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model_instance.parameters(), args.clip)
            optimizer.step()
            avg_loss += float(loss.detach())
        avg_loss /= max(batch_idx + 1, 1)
        if writer is not None:
            writer.add_scalar("loss/avg_loss", avg_loss, epoch)
        if val_dataset is not None:
            val_accs.append(evaluate(val_dataset, model_instance, args, name="Validation")["acc"])
        if prefix and epoch % 10 == 0:
            model_instance.save(f"{prefix}_epoch{epoch}.pt")
    if prefix:
        torch.save(dict(model_state=model_instance.state_dict(), optimizer=optimizer.state_dict()), f"{prefix}_final.pt")
    return model_instance, val_accs


def evaluate(dataset, model, args, name="Validation", testing=False, max_num_examples=None):
    """train_chargrid_funsd_msau.py:121-163: argmax over classes on labelled pixels -> micro precision / recall / accuracy.
    Everything per pixel stays on the device: the uint8 arg-max map comes from the head kernel, the label x prediction counts
    over label != 0 from ``msau_confusion_counts``; the host reads one n_class x n_class int64 matrix at the end."""
    from . import _lib
    device = torch.device("cuda")
    model.eval()
    nc = model.n_class
    conf = torch.zeros((nc, nc), dtype=torch.int64, device=device)
    for batch_idx, data in enumerate(dataset):
        h0 = data["mask"].float().to(device)
        lab = data["label"]
        lab = (lab if lab.dtype == torch.uint8 else lab.long()).to(device).contiguous()
        pred = model.predict_classes(h0)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().msau_confusion_counts(pred.data_ptr(), lab.data_ptr(), 0 if lab.dtype == torch.uint8 else 1, pred.numel(),
                                                        nc, conf.data_ptr(), _lib.current_stream()))
        if max_num_examples is not None and (batch_idx + 1) * args.batch_size > max_num_examples:
            break
    c = conf.cpu().numpy()
    if testing:                                   # indices[indices == 0] = dataset.labels['other']  (:140-141)
        other = int(dataset.labels["other"])
        if other != 0:
            c[:, other] += c[:, 0]
            c[:, 0] = 0
    total = int(c.sum())
    acc = float(np.trace(c)) / total if total else 0.0
    # single-label micro precision == micro recall == accuracy (sklearn average="micro")
    return {"prec": acc, "recall": acc, "acc": acc, "confusion": c}
