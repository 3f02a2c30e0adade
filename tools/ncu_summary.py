"""Summarises an .ncu-rep (one kernel launch) into markdown: the numbers the roofline and the optimisation
decisions are read from.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of 64 warps"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (warp instructions / cycle / SM, max 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (max 32)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory pipeline throughput, % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate, %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput, % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors (32 B)"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local (stack) loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local (stack) stores"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, % busy"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe, % busy"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, % busy"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe, % busy"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe, % busy"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "wait", "not_selected", "branch_resolving", "short_scoreboard", "no_instruction",
               "math_pipe_throttle", "dispatch_stall", "lg_throttle", "mio_throttle", "barrier"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("## %s  (grid %s, block %s)\n" % (d.get("Kernel Name", "?")[:90], d.get("Grid Size", "?"), d.get("Block Size", "?")))
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("| %s (`%s`) | %s | %s |" % (label, k, d[k], u.get(k, "")))
        print("\nWarp stall reasons (warps stalled per issued instruction):\n")
        print("| reason | ratio |\n|---|---|")
        for n in STALL_NAMES:
            k = STALLS % n
            if k in d and d[k] != "":
                print("| %s | %s |" % (n, d[k]))
        print()


if __name__ == "__main__":
    main()
