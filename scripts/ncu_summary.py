"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of metrics the roofline needs.
    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [algorithmic_bytes] [algorithmic_flops]"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size','?')} block {d.get('Block Size','?')}")
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k]:>16s} {u[k]}")
        try:
            t = float(d["gpu__time_duration.sum"]); tu = u["gpu__time_duration.sum"]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[tu]
            rd = float(d["dram__bytes_read.sum"]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u["dram__bytes_write.sum"]]
            print(f"  traffic (dram read+write)  {(rd + wr) / 1e6:.1f} MB   -> {(rd + wr) / t_s / 1e9:.0f} GB/s")
            if len(sys.argv) > 2 and float(sys.argv[2]) > 0:
                print(f"  algorithmic bytes          {float(sys.argv[2]) / 1e6:.1f} MB   -> {float(sys.argv[2]) / t_s / 1e9:.0f} GB/s")
            if len(sys.argv) > 3 and float(sys.argv[3]) > 0:
                print(f"  algorithmic flops          {float(sys.argv[3]) / 1e9:.1f} GF   -> {float(sys.argv[3]) / t_s / 1e12:.0f} TFLOP/s (under ncu, cold)")
        except Exception as e:
            print("  (summary failed:", e, ")")
if __name__ == "__main__":
    main()
