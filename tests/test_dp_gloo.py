"""Host-side data-parallel logic on CPU (gloo, world_size 2): sharding, flat all-reduce, and the identity
"sum over ranks of the 1/N-scaled shard gradients == gradient of the batch-mean loss" that the N-GPU train step relies
on.  The oracle stands in for the CUDA engine as the per-shard gradient function."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msau_b200.train import allreduce_flat, dp_gradient, shard_pages
from oracle import model as om
from oracle.synth import synth_input

CFG = om.MsauConfig(channels=8, n_class=3, scale_space_num=2, res_depth=1, feat_root=16)


def test_shard_pages_balanced_and_complete():
    for n, w in ((16, 1), (16, 2), (16, 8), (5, 2), (7, 4), (3, 8)):
        seen = []
        sizes = []
        for r in range(w):
            rg = shard_pages(n, r, w)
            seen += list(rg)
            sizes.append(len(rg))
        assert seen == list(range(n))
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_pages(4, 4, 4)


def _flat(grads, keys):
    return torch.cat([(grads[k] if grads[k] is not None else torch.zeros_like(SD[k])).reshape(-1) for k in keys])


SD = om.init_state_dict(CFG, 3)
KEYS = [k for k, _ in om.param_schema(CFG)]


def _local_grad_fn(x, labels):
    def fn(pages, scale):
        if not pages:
            return torch.zeros(()), torch.zeros(sum(v.numel() for v in SD.values()))
        leaves = {k: v.clone().requires_grad_(True) for k, v in SD.items()}
        idx = torch.tensor(pages)
        logits, aux = om.msau_forward(leaves, CFG, x[idx])
        loss_sum = om.batch_loss(logits, aux, labels[idx]) * len(pages)
        (loss_sum * scale).backward()
        return loss_sum.detach(), _flat({k: v.grad for k, v in leaves.items()}, KEYS)
    return fn


def _worker(rank, world, port, n_pages, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, labels = synth_input(CFG.channels, CFG.n_class, n_pages, 12, 10, 7)
    loss, flat = dp_gradient(_local_grad_fn(x, labels), n_pages, rank, world)
    t = torch.arange(4, dtype=torch.float32) + rank
    allreduce_flat(t)
    if rank == 0:
        torch.save(dict(loss=loss, flat=flat, t=t), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pages", [4, 5])
def test_dp_gradient_equals_single_process(tmp_path, n_pages):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, n_pages, out), nprocs=2, join=True)
    got = torch.load(out)
    x, labels = synth_input(CFG.channels, CFG.n_class, n_pages, 12, 10, 7)
    loss, _, _, grads = om.loss_and_grads(SD, CFG, x, labels)
    want = _flat(grads, KEYS)
    assert abs(float(got["loss"]) - float(loss)) < 1e-5
    assert (got["flat"] - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    assert torch.equal(got["t"], torch.tensor([1.0, 3.0, 5.0, 7.0]))
