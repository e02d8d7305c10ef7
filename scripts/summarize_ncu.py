#!/usr/bin/env python
"""Summarise ncu captures brought back in gpurun_out/ into small tracked text files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_C4.csv profiles/r01_launches_C4.md
  python scripts/summarize_ncu.py report   gpurun_out/prof_eval_cam.ncu-rep profiles/r01_eval_cam.md
"""
import collections, csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: {src}; {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.1f} us total\n\n")
        f.write("| kernel | launches | total us | avg us | share | last grid x block |\n|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / 1e3 / a[0]:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} x {a[3]} |\n")


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"\n## {d.get('Kernel Name', '?')}  (launch id {d.get('ID')})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            stalls = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]), reverse=True)
            f.write("\nTop stall reasons (warps stalled per issue-active cycle): " + ", ".join(
                f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls[:6]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
