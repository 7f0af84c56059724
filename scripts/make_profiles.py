"""Turn the raw ncu outputs a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/.

    python scripts/make_profiles.py <launches.csv> [<name>=<file.ncu-rep> ...]

* profiles/launches_r1_b16.csv + _summary.txt : per-launch device time / DRAM bytes of two train steps
* profiles/traffic_r1.json                    : average DRAM bytes per launch per kernel family (bench.py's roofline.traffic)
* profiles/ncu_r1_<name>_summary.txt          : selected `ncu --set full` metrics of one kernel capture
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
TAG = os.environ.get("MSAU_ROUND", "r2")          # file-name tag of the round the captures belong to

FAMILY = [("feature_fill", "feature_fill_kernels"), ("feature_ids", "feature_ids_kernel"), ("ccl_", "ccl_kernels"), ("class_closing", "class_closing_row_kernel"),
          ("conv3_tc_kernel<0, 2,", "conv3_tc_kernel_c32"), ("conv3_tc_kernel<1, 2,", "conv3_tc_kernel_c32"), ("conv3_tc_kernel<2, 2,", "conv3_tc_kernel_c32"),
          ("conv3_tc_kernel<3, 2,", "conv3_tc_kernel_c32"), ("conv3_tc_kernel<4, 2,", "conv3_tc_kernel_c32"), ("conv3_tc_kernel<5, 2,", "conv3_tc_kernel_c32"),
          ("wgrad_tc4", "wgrad_tc4_kernel"), ("wgrad_tc3", "wgrad_tc3_kernel"), ("wgrad_tc2", "wgrad_tc2_kernel"), ("wgrad_tc", "wgrad_tc_kernel"), ("conv3_tc", "conv3_tc_kernel"),
          ("conv_tc", "conv_tc_kernel"), ("conv1x1", "conv1x1_kernel"), ("relu_mask", "relu_mask_kernel"), ("attn_tc", "attn_kernels")]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID, iG = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    d = {}
    for r in rows[1:]:
        e = d.setdefault(r[iID], {"k": r[iK].split("(")[0].replace("void ", "").replace("msau::", "").replace("<unnamed>::", ""), "g": r[iG]})
        e[r[iM]] = float(r[iV].replace(",", ""))
    return list(d.values())


def main():
    src = sys.argv[1]
    if src != "-":
        launch_list(src)
    summaries()


def launch_list(src):
    shutil.copy(src, os.path.join(PROF, f"launches_{TAG}_b16.csv"))
    L = launches(src)
    tot = sum(v["gpu__time_duration.sum"] for v in L)
    agg = {}
    for v in L:
        a = agg.setdefault(v["k"][:40], [0, 0.0, 0.0])
        a[0] += 1; a[1] += v["gpu__time_duration.sum"]; a[2] += v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
    with open(os.path.join(PROF, f"launches_{TAG}_b16_summary.txt"), "w") as f:
        f.write(f"two train steps, B=16, 512x512 (scripts/prof_step.py 16 2) under ncu --clock-control none: {len(L)} launches, {tot / 1e6:.2f} ms of device time\n")
        f.write("(per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_breakdown, not absolutes)\n\n")
        for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:42s} n={n:4d} {t / 1e6:8.2f} ms {100 * t / tot:5.1f}%   DRAM {b / n / 1e6:8.1f} MB/launch  {b / t:7.0f} GB/s\n")
    fam = {}
    for v in L:
        for key, name in FAMILY:
            if key in v["k"]:
                a = fam.setdefault(name, [0, 0.0, 0.0])
                a[0] += 1; a[1] += v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0); a[2] += v["gpu__time_duration.sum"]
                break
    json.dump({k: dict(launches=n, dram_bytes_per_launch=b / n, ncu_ns_per_launch=t / n,
                       source="ncu dram__bytes_read.sum + dram__bytes_write.sum, launch list of scripts/prof_step.py 16 2")
               for k, (n, b, t) in fam.items()}, open(os.path.join(PROF, f"traffic_{TAG}.json"), "w"), indent=1)


def summaries():
    keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct")
    for arg in sys.argv[2:]:
        name, rep = arg.split("=")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        with open(os.path.join(PROF, f"ncu_{TAG}_{name}_summary.txt"), "w") as f:
            f.write(f"ncu --set full --clock-control none --import-source on, {os.path.basename(rep)}\n")
            for r in rows[2:]:
                f.write("---\n")
                for i, h in enumerate(hdr):
                    if h in ("Kernel Name", "Grid Size", "Block Size") or h in keep or "issue_stalled" in h and "pcsamp" not in h:
                        f.write(f"{h} = {r[i]} {rows[1][i]}\n")


if __name__ == "__main__":
    main()
