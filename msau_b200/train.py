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


# ----------------------------------------------------------------------------- reference-shaped driver
def train(dataset, model_instance, args, same_feat=True, val_dataset=None, test_dataset=None, writer=None, mask_nodes=True):
    """train_chargrid_funsd_msau.py:16-118 without the plotting / checkpoint-name helpers.  ``dataset`` yields dicts with
    "mask" [1,C,H,W] and "label" [1,H,W] (FUNSDMaskDataLoader.getitem, dgfb.py:216-222)."""
    device = torch.device("cuda")
    optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model_instance.parameters()), lr=0.0001)
    model_instance = model_instance.to(device)
    val_accs = []
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
    return model_instance, val_accs


def evaluate(dataset, model, args, name="Validation", testing=False, max_num_examples=None):
    """train_chargrid_funsd_msau.py:121-163: argmax over classes on labelled pixels -> micro precision / recall / accuracy.
    The argmax runs on the device (uint8 class map), only labels and predictions of labelled pixels reach the host."""
    device = torch.device("cuda")
    model.eval()
    labels, preds = [], []
    for batch_idx, data in enumerate(dataset):
        h0 = data["mask"].float().to(device)
        instance_label = np.squeeze(data["label"].long().cpu().numpy())
        indices = model.predict_classes(h0)[0].cpu().numpy()
        keep = instance_label != 0
        indices = indices[keep].astype(np.int64)
        instance_label = instance_label[keep]
        if testing:
            indices[indices == 0] = dataset.labels["other"]
        labels.append(instance_label)
        preds.append(indices)
        if max_num_examples is not None and (batch_idx + 1) * args.batch_size > max_num_examples:
            break
    labels = np.hstack(labels).squeeze()
    preds = np.hstack(preds).squeeze()
    acc = float((labels == preds).mean()) if labels.size else 0.0
    # single-label micro precision == micro recall == accuracy (sklearn average="micro")
    return {"prec": acc, "recall": acc, "acc": acc}
