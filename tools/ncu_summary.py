"""Key metrics of an ncu report (one kernel).  Usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
def main(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = []
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            out.append(f"{h} = {v} {u}")
    return "\n".join(out)
if __name__ == "__main__":
    print(main(sys.argv[1]))
