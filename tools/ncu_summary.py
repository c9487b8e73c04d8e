"""Markdown summary of `ncu --set full` reports: python tools/ncu_summary.py out.md rep1.ncu-rep [rep2 ...]
Reads each report with `ncu -i ... --page raw --csv` and keeps the numbers the roofline argument needs: duration, launch
geometry, occupancy, tensor-pipe activity, DRAM bytes and throughput, shared-memory wavefronts and the warp-stall
sampling breakdown (top reasons)."""
import csv
import io
import subprocess
import sys

KEEP = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("cluster", "launch__cluster_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/CTA", "launch__shared_mem_per_block_dynamic"),
    ("static smem/CTA", "launch__shared_mem_per_block_static"),
    ("waves/SM", "launch__waves_per_multiprocessor"),
    ("occupancy limit (regs/smem/warps) CTAs", None),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("SM busy % (sm__throughput)", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (of active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (elapsed, realtime)", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tensor instr executed", "sm__inst_executed_pipe_tensor.sum"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM GB/s", "dram__bytes.sum.per_second"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem wavefronts (LSU)", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem wavefronts (tensor core reads)", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"),
    ("smem bank conflicts (LSU)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("executed instr", "smsp__inst_executed.sum"),
    ("IPC (active)", "sm__inst_executed.avg.per_cycle_active"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def fmt(uv):
    u, v = uv
    try:
        f = float(v.replace(",", ""))
        v = f"{f:,.2f}" if abs(f) < 1e6 and f != int(f) else f"{int(f):,}"
    except ValueError:
        pass
    return f"{v} {u}".strip()


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    lines = ["# ncu --set full summaries", "",
             "Captured under gpurun with `--clock-control none --import-source on`; durations are cold-cache, serialised "
             "(never bench values). Raw reports stay in gpurun_out/ (scratch).", ""]
    for rep in reps:
        for k in load(rep):
            name = k.get("Kernel Name", ("", "?"))[1]
            lines += [f"## {name[:110]}", "", f"report: `{rep.split('/')[-1]}`", "", "| metric | value |", "|---|---|"]
            for label, key in KEEP:
                if key is None:
                    lim = [k.get(f"launch__occupancy_limit_{x}") for x in ("registers", "shared_mem", "warps")]
                    if all(lim):
                        lines.append(f"| {label} | {' / '.join(v[1] for v in lim)} |")
                    continue
                if key in k and k[key][1] != "":
                    lines.append(f"| {label} | {fmt(k[key])} |")
            stalls = []
            for h, uv in k.items():
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(uv[1]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                lines.append("| warp stalls per issue (top 6) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
            lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
