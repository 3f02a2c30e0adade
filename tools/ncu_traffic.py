"""Writes profiles/traffic.json from one `ncu --set full` capture of the render kernel at the bench workload: DRAM bytes per
launch (the contract's `roofline.traffic`), L2 bytes (lts__t_bytes) and L1 global-load sectors per launch, and the launch
duration of the capture.  Usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep "<workload text>" <spp of the captured launch> > profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys


def num(x):
    return float(x.replace(",", "")) if x not in ("", None) else None


def main():
    rep, workload = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    d, u = dict(zip(hdr, rows[2])), dict(zip(hdr, units))

    def scaled(key, want):
        v = num(d.get(key, ""))
        if v is None:
            return None
        unit = u.get(key, "")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
                "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "sector": 1.0, "": 1.0}
        return v * mult.get(unit, 1.0)
    rd, wr = scaled("dram__bytes_read.sum", "byte"), scaled("dram__bytes_write.sum", "byte")
    out = {
        "dram_bytes_per_launch": (rd or 0.0) + (wr or 0.0), "dram_bytes_read": rd, "dram_bytes_write": wr,
        # bytes the L2 delivered to the SMs' L1s (lts__t_bytes.sum where the report has it, else the crossbar-to-L1 read counter)
        "lts_bytes_per_launch": scaled("lts__t_bytes.sum", "byte") or scaled("l1tex__m_xbar2l1tex_read_bytes.sum", "byte"),
        "lts_bytes_metric": "lts__t_bytes.sum" if scaled("lts__t_bytes.sum", "byte") else "l1tex__m_xbar2l1tex_read_bytes.sum",
        "tma_load_bytes_per_launch": scaled("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "byte"),
        "l1_global_load_sectors_per_launch": scaled("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sector"),
        "l1_hit_rate_pct": num(d.get("l1tex__t_sector_hit_rate.pct", "")), "l2_hit_rate_pct": num(d.get("lts__t_sector_hit_rate.pct", "")),
        "lts_throughput_pct_of_peak": num(d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", "")),
        "kernel_ms_of_capture": scaled("gpu__time_duration.sum", "ms"),
        "kernel": d.get("Kernel Name", "?")[:80], "grid": d.get("Grid Size"), "block": d.get("Block Size"),
        "workload": workload, "spp_of_capture": spp, "source": "ncu --set full --clock-control none, one launch, " + rep.split("/")[-1],
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
