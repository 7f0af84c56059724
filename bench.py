#!/usr/bin/env python
"""bench.py -- pages/sec of the MSAU train step (fwd + loss + bwd + clip + Adam) on 512x512 chargrid pages.

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): batch 16 of synthetic 512x512 chargrid pages PER GPU (weak scaling),
MSAUWrapper(96, 5, featRoot=8, scale_space_num=4, res_depth=2), random-init weights, synthetic pages from the
SURVEY.md section 8(d) generator (198 words + 2 anchors, one-hot D=96).

One JSON line on stdout (rank 0):
  value     whole-job pages/s with the dense fp32 [16,96,512,512] batch already resident in HBM
  e2e       same step driven from HOST page records through the public API: pinned CSR boxes -> H2D -> device
            rasterisation (R1) -> train step -> loss read back, every step inside the timed region
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference step timed on the host cores (bounded sample)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(channels=96, n_class=5, scale_space_num=4, res_depth=2, feat_root=8)
PAGES_PER_GPU = 16
H = W = 512
METRIC = "pages/sec MSAU fwd+bwd 512x512 chargrid"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tf=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def synth_records(first_seed, n):
    from bench_inputs import synth_page
    wp, lp = [], []
    for i in range(n):
        w, l = synth_page(first_seed + i, H, W, 198)
        wp.append(w); lp.append(l)
    return wp, lp


# --------------------------------------------------------------------------------------------- reference arm
def cpu_train_pages(n_pages, warm, threads=None):
    """The reference step (train_chargrid_funsd_msau.py:46-59 restated in oracle/model.py) page by page on the host."""
    import numpy as np
    import torch
    from oracle import model as om
    from oracle import raster as orr
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = om.MsauConfig(**CFG)
    sd = om.init_state_dict(cfg, 0)
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v = {k: torch.zeros_like(t) for k, t in sd.items()}
    times = []
    for i in range(warm + n_pages):
        words, lines = orr.synth_page(1000 + i, H, W, 198)
        t0 = time.perf_counter()
        grid, label = orr.raster_word_chargrid(words, lines, np.eye(CFG["channels"]))     # R1, as dataset.getitem does
        x = torch.Tensor(grid).unsqueeze(0)
        lab = torch.from_numpy(label.astype(np.int64)).unsqueeze(0)
        loss, _, _, grads = om.loss_and_grads(sd, cfg, x, lab)
        dead = f"msau_net.blocks.{cfg.num_blocks - 1}.downsamplingblock.layer_attentions."
        grads = {k: (None if k.startswith(dead) else g) for k, g in grads.items()}
        om.clip_adam_step(sd, grads, m, v, step=i + 1)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, threads = cpu_train_pages(args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = 1e3 / ms
    sample = "1 page per step (rasterise R1 + fwd + loss + bwd + clip + Adam), torch-CPU fp32 oracle port of the reference step"
    print(json.dumps(dict(
        impl="reference", metric=METRIC, value=val, unit="pages/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="MSAU chargrid training step, synthetic 512x512 pages (configs[1])", model_kwargs=CFG,
                    pages_per_step=1, note="reference = pure-Python/PyTorch; no installable package, oracle port timed"),
        cpu_baseline=dict(value=val, unit="pages/s", cores=threads, kind="port", sample=sample),
        e2e=dict(value=val, unit="pages/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))



# --------------------------------------------------------------------------------------------- extra legs
def bench_c5(model, raster, rank, dev, timed, world, reps):
    """BASELINE.json configs[4]: 1024x768 chargrid inference, 64 pages per GPU (replicas only).  The pages are rasterised on the
    device into int16 channel-id maps once; a step = predict_classes (arg-max class map, uint8) over the 64 resident pages."""
    import torch
    from bench_inputs import synth_page
    Hc, Wc, n = 1024, 768, 64
    wp, lp = [], []
    for i in range(n):
        w, l = synth_page(5000 + rank * n + i, Hc, Wc, 198)
        wp.append(w); lp.append(l)
    table = torch.eye(CFG["channels"], dtype=torch.float64, device=dev)
    ids, _, _ = raster.rasterize_word_chargrid(wp, lp, table, out_hw=(Hc, Wc), layout="ids", device=dev)
    model.eval()
    chunk = 16

    def step():
        return model.predict_classes(ids, layout=2, pages_per_call=chunk)
    for _ in range(2):
        step()
    ms = timed(step, reps) / reps
    model.train()
    return dict(value=world * n / (ms * 1e-3), unit="pages/s", pages_per_gpu=n, pages_per_call=chunk, ms_per_step=ms,
                input="int16 channel-id maps [64,1024,768] resident in HBM (R1 rasterised on the device)", output="uint8 arg-max maps")


def gpu_library_baseline(grid, labels64, dev):
    """The reference step through PyTorch-eager on the same GPU (cuDNN / cuBLAS / ATen fp32, TF32 off): the oracle port of
    model/model.py + train_chargrid_funsd_msau.py:46-59 with its tensors on the device.  Batch 16 (mean over pages of the per-page
    loss) and page by page (the reference's own loop).  Same pages, same init as the native arm."""
    import torch
    from oracle import model as om
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = om.MsauConfig(**CFG)
    sd = {k: v.to(dev) for k, v in om.init_state_dict(cfg, 0).items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v = {k: torch.zeros_like(t) for k, t in sd.items()}
    dead = f"msau_net.blocks.{cfg.num_blocks - 1}.downsamplingblock.layer_attentions."
    step_no = [0]

    def step(x, lab):
        _, _, _, grads = om.loss_and_grads(sd, cfg, x, lab)
        grads = {k: (None if k.startswith(dead) else g) for k, g in grads.items()}
        step_no[0] += 1
        om.clip_adam_step(sd, grads, m, v, step=step_no[0])

    def time_it(fn, reps):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    B = grid.shape[0]
    out = dict(unit="pages/s", what="oracle port of the reference step on cuda (PyTorch-eager: cuDNN conv, ATen LRN / attention / CE / "
                                    "clip / Adam), fp32, TF32 off, same B200, CUDA events", kind="port")
    try:
        ms_b = time_it(lambda: step(grid, labels64), 3)
        out.update(value=B / (ms_b * 1e-3), ms_per_step=ms_b, batch=B)
    except torch.cuda.OutOfMemoryError as e:
        out.update(value=None, batch_error=str(e)[:120])
        torch.cuda.empty_cache()
    ms_p = time_it(lambda: [step(grid[i:i + 1], labels64[i:i + 1]) for i in range(min(B, 4))], 2) / min(B, 4)
    out.update(page_by_page_value=1e3 / ms_p, page_by_page_ms=ms_p)
    del sd, m, v
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- native arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--pages", type=int, default=PAGES_PER_GPU, help="pages per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-dense", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager kernel launches instead of the CUDA-graph replay of the step")
    ap.add_argument("--no-graph-tail", action="store_true", help="keep the all-reduce and clip + Adam outside the CUDA graph")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --pages per GPU (default 16); strong: global batch 16 split over the GPUs (16/N pages per GPU)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="engine option for A/B runs, e.g. --option pdl=0")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling / configs[4] / library-GPU-baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import msau_b200
    from msau_b200 import _lib, raster

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout clean for the single JSON line: library chatter (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    K, Wm, P = args.steps, max(args.warmup, 3), args.pages
    if args.scaling == "strong":
        P = max(PAGES_PER_GPU // world, 1)
    tail = not args.no_graph_tail

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # nothing under oracle/ is touched by this arm: the wrapper's own initialiser draws the reference's init distributions
    # (model/layers/layers.py:33-36,59-60) from a seeded CPU generator -> identical replicas on every rank
    model = msau_b200.MSAUWrapper(CFG["channels"], CFG["n_class"], dict(final_act="softmax", featRoot=CFG["feat_root"],
                                                                          scale_space_num=CFG["scale_space_num"], res_depth=CFG["res_depth"]))
    model.reset_parameters(seed=0)
    model = model.to(dev).train()
    for kv in args.option:
        name, val = kv.split("=")
        model.set_option(name, int(val))
    pg = dist.group.WORLD if world > 1 else None

    # ---- synthetic pages for this rank: records on the host, dense grid rasterised once for the resident run
    words, lines = synth_records(rank * P, P)
    table = torch.eye(CFG["channels"], dtype=torch.float64, device=dev)
    grid, label, _ = raster.rasterize_word_chargrid(words, lines, table, out_hw=(H, W), layout="nchw", device=dev)
    labels64 = label.long()

    # forward + loss + backward replayed from a CUDA graph (one launch instead of ~420; the all-reduce and clip + Adam stay
    # eager): same kernels, same work, but the step no longer depends on the host's launch rate.  --no-graph = eager launches.
    graph_static = "static"     # the resident batch is the same tensor every step: the graph reads it in place

    def step_resident():
        return model.train_step(grid, labels64, process_group=pg, world_size=world, use_graph=False if args.no_graph else graph_static,
                                graph_tail=tail)

    def step_resident_eager():
        return model.train_step(grid, labels64, process_group=pg, world_size=world)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    graph_note = None
    if not args.no_graph:
        try:        # the capture is exercised once before anything is timed; if this box cannot capture, time eager launches
            step_resident()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            graph_note = f"CUDA-graph capture failed ({type(e).__name__}: {e}); eager launches timed instead"
            sys.stderr.write(graph_note + "\n")
            args.no_graph = True
            model._train_graphs.clear()
    for _ in range(Wm):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_total = timed(step_resident, K)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = world * P / (ms_step * 1e-3)

    # ---- e2e: host page records -> device every step (public API: raster.rasterize_word_chargrid + train_step)
    h2d = [0]
    # the step's inputs in pinned host memory (what a data loader hands over): CSR page records, ~300 KB per 16 pages
    host_words = raster.HostBatch(words, with_chars=True)
    host_lines = raster.HostBatch(lines, with_chars=False, with_labels=True)

    def step_e2e():
        wb = raster.BoxBatch.from_host(host_words, dev)
        lb = raster.BoxBatch.from_host(host_lines, dev)
        geom = wb.geometry()
        # the chargrid as the int16 channel-id map (feature table = identity): the structured first layer consumes it directly
        g = raster.raster_features(wb, geom, table, (H, W), True, "ids")
        lab = raster.raster_labels(lb, geom, (H, W))
        h2d[0] = wb.h_bytes + lb.h_bytes
        loss = model.train_step(g, lab, layout=2, process_group=pg, world_size=world, use_graph=not args.no_graph, graph_tail=tail)
        return float(loss)            # D2H read of the step's result

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, K) / K
    e2e = dict(value=world * P / (ms_e2e * 1e-3), unit="pages/s", h2d_bytes_per_step=h2d[0], d2h_bytes_per_step=4,
               input="host page records (CSR boxes + char ids, pinned) rasterised on the device (R1 -> int16 id map) every step")

    # ---- e2e with the reference's literal input format: dense fp32 NCHW page tensor in pinned host memory
    e2e_dense = None
    if not args.no_e2e_dense:
        host_x = torch.empty(grid.shape, dtype=torch.float32).pin_memory()
        host_x.copy_(grid)
        host_l = torch.empty(labels64.shape, dtype=torch.int64).pin_memory()
        host_l.copy_(labels64)

        def step_dense():
            xd = host_x.to(dev, non_blocking=True)
            ld = host_l.to(dev, non_blocking=True)
            return float(model.train_step(xd, ld, process_group=pg, world_size=world))

        step_dense()
        ms_d = timed(step_dense, max(2, K // 2)) / max(2, K // 2)
        e2e_dense = dict(value=world * P / (ms_d * 1e-3), unit="pages/s",
                         h2d_bytes_per_step=host_x.numel() * 4 + host_l.numel() * 8, d2h_bytes_per_step=4,
                         input="dense fp32 [B,96,512,512] + int64 labels from pinned host memory (PCIe-bound)")
        del host_x, host_l

    # ---- per-kernel pass (CUDA events around every launch) -> roofline of the dominant kernel
    roof, breakdown = None, None
    # every rank runs the pass (the step contains the all-reduce); rank 0 reports its own kernels
    # (weight gradients normally overlap the data-gradient chain on a side stream; serialise them here so that every
    #  kernel's events measure that kernel alone)
    model.set_option("wgrad_side_stream", 0)
    _lib.profile_enable(True)
    for _ in range(K):
        step_resident_eager()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    model.set_option("wgrad_side_stream", 1)
    if rank == 0:
        pk = peaks()
        tot = sum(v["ms"] for v in rep.values())
        breakdown = {k: dict(share=round(v["ms"] / tot, 4), ms_per_step=round(v["ms"] / K, 4), launches_per_step=v["launches"] // K,
                             gbs=round(v["bytes"] / v["ms"] / 1e6, 1), tflops=round(v["flops"] / v["ms"] / 1e9, 2))
                     for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}
        top, tv = max(rep.items(), key=lambda kv: kv[1]["ms"])
        t_hbm = tv["bytes"] / (pk["hbm_gbs"] * 1e9)
        t_ten = tv["flops"] / (pk["tf"] * 1e12)
        if t_hbm >= t_ten:
            ach = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9
            roof = dict(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"], traffic=None)
        else:
            ach = tv["flops"] / (tv["ms"] * 1e-3) / 1e12
            roof = dict(bound="tensor", achieved=ach, peak=pk["tf"], unit="TFLOP/s", frac=ach / pk["tf"], traffic=None)
        try:   # DRAM bytes per launch of this kernel family, from the committed ncu launch list of the same step
            tfile = "traffic_r2.json" if os.path.exists(os.path.join(ROOT, "profiles", "traffic_r2.json")) else "traffic_r1.json"
            tr = json.load(open(os.path.join(ROOT, "profiles", tfile))).get(top)
            if tr:
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_source"] = f"profiles/{tfile}: " + tr["source"]
                roof["algorithmic_bytes_per_launch"] = tv["bytes"] / tv["launches"]
        except Exception:
            pass
        roof.update(kernel=top, peak_source=pk["src"], launches=tv["launches"], avg_launch_ms=tv["ms"] / tv["launches"],
                    share_of_step=tv["ms"] / tot, pass_="separate K-step pass with CUDA events around every launch, single stream")
    barrier()


    # ---- extra legs (all outside the timed region above; every rank takes part where a collective is involved)
    strong = c5 = gpu_lib = None
    if not args.no_extras:
        # (1) strong scaling of configs[1]: global batch 16 split over the GPUs
        Ps = max(PAGES_PER_GPU // world, 1)
        if world == 1 and P == PAGES_PER_GPU:
            strong = dict(pages_per_gpu=Ps, global_batch=Ps * world, ms_per_step=ms_step, value=value, unit="pages/s")
        else:
            gs, ls = grid[:Ps].contiguous(), labels64[:Ps].contiguous()

            def step_strong():
                return model.train_step(gs, ls, process_group=pg, world_size=world, use_graph=False if args.no_graph else "static",
                                        graph_tail=tail)
            for _ in range(3):
                step_strong()
            ms_s = timed(step_strong, K) / K
            strong = dict(pages_per_gpu=Ps, global_batch=Ps * world, ms_per_step=ms_s, value=Ps * world / (ms_s * 1e-3), unit="pages/s")
            del gs, ls
        # (2) BASELINE.json configs[4]: 1024x768 chargrid inference, 64 pages per GPU, replicas (no collective)
        try:
            c5 = bench_c5(model, raster, rank, dev, timed, world, max(3, K // 3))
        except Exception as e:  # noqa: BLE001
            c5 = dict(error=f"{type(e).__name__}: {e}")
    if rank == 0 and world == 1 and not args.no_extras:
        # (3) library-GPU baseline: the reference network written with torch ops (oracle port), same B200, cuDNN / cuBLAS fp32,
        #     TF32 off -- the "PyTorch-eager on the same GPU" bar of SURVEY.md 2b.  A reported baseline, like cpu_baseline.
        try:
            gpu_lib = gpu_library_baseline(grid, labels64, dev)
        except Exception as e:  # noqa: BLE001
            gpu_lib = dict(error=f"{type(e).__name__}: {e}")
    barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # (rank 0 at N = 1 only: under torchrun the other ranks share the host cores)
        times, threads = cpu_train_pages(3, 1)
        cpu = dict(value=len(times) / sum(times), unit="pages/s", cores=threads, kind="port",
                   sample="3 pages (after 1 warm-up) of the same workload: R1 rasterise + fwd + loss + bwd + clip + Adam, "
                          "torch-CPU fp32 oracle port of train_chargrid_funsd_msau.py:46-59")
    if rank == 0:
        out = dict(metric=METRIC, value=value, unit="pages/s", n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms_step,
                   higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                   dtype="f32 storage; tcgen05 bf16x3 split (hi*hi + lo*hi + hi*lo, fp32 accumulate) in forward / data gradients, "
                         "single-term bf16 operands in weight gradients and attention backward",
                   data="synthetic",
                   config=dict(workload=f"MSAU chargrid training step, batch {P} synthetic 512x512 pages per GPU (BASELINE.json configs[1])",
                               model_kwargs=CFG, pages_per_gpu=P, global_batch=world * P, parallelism=f"dp{world}",
                               l2="inputs (1.6 GB) and activations (>30 GB) are larger than L2; no flush needed",
                               step="fwd + masked CE (main+aux) + bwd + NCCL all-reduce (N>1) + clip_grad_norm(1.0) + Adam(1e-4)",
                               launch=(graph_note or "eager kernel launches") if args.no_graph else
                                      ("the whole step (fwd + loss + bwd + all-reduce + clip + Adam) replayed from ONE CUDA graph" if tail else
                                       "fwd + loss + bwd replayed from a CUDA graph (all-reduce, clip + Adam eager)") +
                                      "; gpu_launches counts the graph's kernels per replay"),
                   e2e=e2e, e2e_dense=e2e_dense, gpu_launches=int(launches), roofline=roof, kernel_breakdown=breakdown,
                   strong_scaling=strong, c5_inference_1024x768=c5, gpu_library_baseline=gpu_lib,
                   cpu_baseline=cpu, clocks=clocks)
        if e2e_dense is not None:     # the reference's literal input format next to the record-fed number, in the key the driver reads
            out["e2e"]["dense_input_value"] = e2e_dense["value"]
            out["e2e"]["dense_input_h2d_bytes_per_step"] = e2e_dense["h2d_bytes_per_step"]
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        # CUDA graphs that hold NCCL kernels must be gone before the communicator is torn down (destroy_process_group otherwise
        # waits for ever); a timer makes sure a stuck teardown can never keep a finished benchmark from exiting
        import gc
        barrier()
        model._train_graphs.clear(); model._graphs.clear()
        del model
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    main()
