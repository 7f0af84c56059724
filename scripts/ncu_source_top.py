"""Top source lines of a kernel by warp-stall samples, from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.

    python scripts/ncu_source_top.py <source.csv> "<substring of the kernel's function name>" [n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    lines, cur_file, cur_fn, hdr = [], None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            cur_fn = r[1]; hdr = None; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or cur_fn is None or want not in cur_fn:
            continue
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        if len(r) > i_i and r[2] == "-" and r[i_s].isdigit():
            lines.append((int(r[i_s]), int(r[i_i] or 0), cur_file, r[0], r[1]))
    tot = sum(l[0] for l in lines) or 1
    toti = sum(l[1] for l in lines) or 1
    print(f"{want}: {tot} samples, {toti} warp instructions")
    for s, i, f, l, src in sorted(lines, reverse=True)[:n]:
        print(f"{100 * s / tot:5.1f}% samp {100 * i / toti:5.1f}% inst  {f}:{l}: {src.strip()[:100]}")


if __name__ == "__main__":
    main()
