"""torchrun worker of tests/test_round2_gpu.py::test_two_rank_nccl_step_equals_single_gpu_step_on_concatenated_batch.

Every rank takes its shard of a 4-page batch, runs 3 data-parallel train steps (eager launches) and the same 3 steps again
from the CUDA graph that also holds the NCCL all-reduce and the clip + Adam kernels; rank 0 writes the results."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import msau_b200  # noqa: E402
from msau_b200.train import shard_pages  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle.synth import synth_input  # noqa: E402


def main():
    out = sys.argv[1]
    import threading
    hard = threading.Timer(200.0, lambda: os._exit(3))          # a stuck collective must not outlive the test's own timeout
    hard.daemon = True
    hard.start()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = om.MsauConfig(channels=96, n_class=5, scale_space_num=3, res_depth=2, feat_root=8)
    sd = om.init_state_dict(cfg, 7)
    x, labels = synth_input(cfg.channels, cfg.n_class, 4, 64, 48, 8)
    mine = list(shard_pages(4, rank, world))
    xs, ls = x[mine].to(dev), labels[mine].to(dev)

    def run(use_graph):
        m = msau_b200.MSAUWrapper(cfg.channels, cfg.n_class, dict(final_act="softmax", featRoot=cfg.feat_root,
                                                                  scale_space_num=cfg.scale_space_num, res_depth=cfg.res_depth))
        m.load_state_dict(sd)
        m = m.to(dev).train()
        losses = [float(m.train_step(xs, ls, process_group=dist.group.WORLD, world_size=world, use_graph=use_graph)) for _ in range(3)]
        return m, losses

    m, losses = run(False)
    print(f"rank {rank}: eager steps done {losses}", file=sys.stderr, flush=True)
    mg, losses_g = run(True)
    print(f"rank {rank}: graph steps done {losses_g}", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    flat = m.flat_params.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    graph_ok = all(abs(a - b) <= 1e-5 * max(1.0, abs(a)) for a, b in zip(losses, losses_g))
    graph_ok = graph_ok and (mg.flat_params - m.flat_params).abs().max().item() <= 2.5e-4
    graph_ok = graph_ok and (mg.flat_params - m.flat_params).abs().mean().item() <= 2e-7
    if rank == 0:
        torch.save(dict(params={k: v.detach().cpu() for k, v in m.state_dict().items()}, losses=all_losses,
                        replicas_identical=all(torch.equal(g, gathered[0]) for g in gathered), graph_matches_eager=bool(graph_ok)), out)
    dist.barrier()
    # graphs holding NCCL kernels must be released before the communicator goes away; never let a stuck teardown hang the test
    import gc
    import threading
    del m, mg
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


if __name__ == "__main__":
    main()
