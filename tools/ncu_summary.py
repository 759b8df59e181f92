"""Extracts the judged metrics from an .ncu-rep (read here, no GPU needed) into a small text summary.

    python tools/ncu_summary.py gpurun_out/prof_assoc_scan_r01.ncu-rep profiles/assoc_scan_r01.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"kernel: {d.get('Kernel Name', '?')}   (ncu --set full --clock-control none; source: {rep})")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
        st = []
        for k in hdr:
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[k].replace(",", "")), k))
                except ValueError:
                    pass
        lines.append("  warp stall reasons per issued instruction (top 6):")
        for v, k in sorted(st, reverse=True)[:6]:
            lines.append(f"    {k[len(STALL):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
