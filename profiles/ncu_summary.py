"""Compact summary of an `ncu --set full` report for one kernel launch.
    python profiles/ncu_summary.py gpurun_out/adam_full_hint1.ncu-rep > profiles/ncu_adam_r1.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary: {path}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name[:120]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| `{k}` | {r[i]} | {units[i]} |")
        # warp-stall breakdown (per issue-active ratios)
        stalls = [(hdr[i], r[i]) for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled_") and hdr[i].endswith("_per_issue_active.ratio")]
        stalls = sorted(((h, float(v)) for h, v in stalls if v not in ("", "n/a")), key=lambda t: -t[1])[:6]
        print("\ntop warp stalls (warps stalled per issue-active cycle):\n")
        for h, v in stalls:
            print(f"* `{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}`: {v:.2f}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
