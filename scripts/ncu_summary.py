"""Summarise an .ncu-rep (read here, no GPU needed): one line per captured launch with the metrics the roofline uses.
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.txt]"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "us", 1e-3 if False else 1.0), ("dram__bytes_read.sum", "rdMB", 1.0), ("dram__bytes_write.sum", "wrMB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__occupancy_limit_shared_mem", "occ_smem", 1.0),
        ("launch__occupancy_limit_registers", "occ_reg", 1.0), ("launch__grid_size", "grid", 1.0)]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    parts = [r[col["Kernel Name"]][:48].ljust(48)]
    for name, short, _ in WANT:
        if name in col:
            v, u = r[col[name]], units[col[name]]
            try:
                f = float(v.replace(",", ""))
                if u in ("ns", "nsecond"): f, u = f / 1e3, "us"
                if u == "byte": f, u = f / 1e6, "MB"
                if u == "Kbyte": f, u = f / 1e3, "MB"
                if u == "Gbyte": f, u = f * 1e3, "MB"
                v = f"{f:.1f}"
            except ValueError:
                pass
            parts.append(f"{short}={v}")
    print("  ".join(parts))
