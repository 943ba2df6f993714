"""Per-source-line totals of an `ncu --set full --import-source on` capture (kernels built with -lineinfo):
executed warp instructions and stall samples per line of CUDA source, optionally grouped into named line ranges.

    python tools/ncu_lines.py report.ncu-rep [--top 40] [--regions file.cu:lo-hi=name,...]
"""
import argparse
import csv
import io
import subprocess
from collections import defaultdict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regions", default="")
    ap.add_argument("--kernel", default="")
    args = ap.parse_args()
    cmd = ["ncu", "-i", args.report, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if args.kernel:
        cmd += ["-k", "regex:" + args.kernel]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, hdr = "", None
    per_line = defaultdict(lambda: [0, 0, ""])  # (file, line) -> [instructions, samples, text]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or r[0] in ("", "Function Name"):
            continue
        try:
            line = int(r[0])
            inst, samp = int(r[i_inst]), int(r[i_samp])
        except ValueError:
            continue
        e = per_line[(cur_file, line)]
        e[0] += inst
        e[1] += samp
        e[2] = r[1].strip()[:90]
    tot_i = sum(v[0] for v in per_line.values()) or 1
    tot_s = sum(v[1] for v in per_line.values()) or 1
    print(f"total warp instructions {tot_i:.4g}, stall samples {tot_s}")
    if args.regions:
        regs = []
        for item in args.regions.split(","):
            loc, name = item.split("=")
            f, rng = loc.split(":")
            lo, hi = rng.split("-")
            regs.append((f, int(lo), int(hi), name))
        agg = defaultdict(lambda: [0, 0])
        for (f, line), (inst, samp, _) in per_line.items():
            name = next((n for rf, lo, hi, n in regs if rf == f and lo <= line <= hi), "other")
            agg[name][0] += inst
            agg[name][1] += samp
        print("| region | warp instructions | share | stall samples | share |\n|---|---:|---:|---:|---:|")
        for name, (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"| {name} | {inst:.4g} | {inst / tot_i:.1%} | {samp} | {samp / tot_s:.1%} |")
    print("\n| file:line | warp instr | share | samples | share | source |\n|---|---:|---:|---:|---:|---|")
    for (f, line), (inst, samp, text) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: args.top]:
        print(f"| {f}:{line} | {inst:.4g} | {inst / tot_i:.1%} | {samp} | {samp / tot_s:.1%} | `{text}` |")


if __name__ == "__main__":
    main()
