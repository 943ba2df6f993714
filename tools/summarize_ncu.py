"""Turn the ncu artefacts brought back in gpurun_out/ into the small text/JSON summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/summarize_ncu.py full gpurun_out/k1_full.ncu-rep profiles/r01_k1_fabrik_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum.per_second",
    "smsp__inst_executed.sum", "sm__inst_executed.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"void |<unnamed>::|at::native::|\(anonymous namespace\)::", "", name)[:70]
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised -> compare SHARES)\n\n")
        f.write(f"source: {src}, {len(rows)} launches, {total / 1e6:.3f} ms total\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {cnt} | {ns / 1e6:.3f} | {ns / total:.1%} |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for vals in rows[2:]:
            f.write(f"## {vals[hdr.index('Kernel Name')][:100]}  grid {vals[hdr.index('Grid Size')]} block {vals[hdr.index('Block Size')]}\n\n")
            f.write("| metric | unit | value |\n|---|---|---:|\n")
            for i, h in enumerate(hdr):
                if h in KEEP or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                    f.write(f"| {h} | {units[i]} | {vals[i]} |\n")
            f.write("\n")
        src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src_csv)))
        if len(srows) > 3:
            sh = srows[1]
            data = srows[2:]
            tot = sum(int(r[sh.index("# Samples")]) for r in data) or 1
            stalls = {h: sum(int(r[i]) for r in data) for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h}
            f.write("### warp-state samples (source page, all SASS lines)\n\n| state | share |\n|---|---:|\n")
            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
                f.write(f"| {k} | {v / tot:.1%} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
