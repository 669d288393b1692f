"""Turn the raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (read here, no GPU needed).
   python scripts/summarize_profiles.py r1"""
import collections, csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list of one CUDA-graph replay of the cfg2 step -------------------------------------------------------------
src = os.path.join(G, f"launches_{R}_cfg2_tf32_graph.csv")
rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
starts = [i for i, r in enumerate(rows) if "split_pair_kernel" in r["Kernel Name"]]
first = [i for i in starts if i == 0 or "split_pair_kernel" not in rows[i - 1]["Kernel Name"]]
rows = rows[first[-1]:]                      # the last replay
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["Kernel Name"].split("(")[0][:70], [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"launches_{R}_cfg2_tf32_graph.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node\n"
            f"# python scripts/ncu_graph_step.py --config cfg2 --precision tf32   (one replay of the captured training step, B=256)\n"
            f"# {len(rows)} kernel nodes, sum of node durations {tot:.1f} us (nodes of different streams overlap in the timed run;\n"
            f"# compare SHARES with bench.py, not absolutes)\n")
    f.write(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:72s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {v[1] / tot:7.3f}\n")
with open(os.path.join(P, f"launches_{R}_cfg2_tf32_graph.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "duration_ns"])
    for r in rows:
        w.writerow([r["ID"], r["Kernel Name"][:90], r["Grid Size"], r["Block Size"], r["Metric Value"]])

# ---- full-section captures ---------------------------------------------------------------------------------------------
WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("launch__occupancy_limit_registers", "occ_regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem_B")]
traffic = {}
for rep in (f"prof_{R}_step_kernels", f"prof_{R}_wgrad", f"prof_{R}_pyramid"):
    path = os.path.join(G, rep + ".ncu-rep")
    raw = os.path.join(G, rep + "_raw.csv")
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(path):
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rr = list(csv.reader(out.splitlines()))
    hdr, units = rr[0], rr[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(P, rep + ".txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ... -o {rep}   (see scripts/make_profiles.sh)\n")
        for r in rr[2:]:
            parts = [r[col["Kernel Name"]].split("(")[0][:52].ljust(52)]
            vals = {}
            for name, short in WANT:
                if name not in col:
                    continue
                v, u = r[col[name]], units[col[name]]
                try:
                    x = float(v.replace(",", ""))
                    if u in ("ns", "nsecond"): x /= 1e3
                    if u == "byte" and short.endswith("MB"): x /= 1e6
                    if u == "Kbyte": x /= 1e3
                    if u == "Gbyte": x *= 1e3
                    vals[short] = x
                    parts.append(f"{short}={x:.1f}")
                except ValueError:
                    parts.append(f"{short}={v}")
            f.write("  ".join(parts) + "\n")
            key = r[col["Kernel Name"]].split("(")[0][:52] + " grid=" + r[col["Grid Size"]]
            if "dram_rd_MB" in vals:
                traffic[key] = dict(dram_bytes=(vals["dram_rd_MB"] + vals["dram_wr_MB"]) * 1e6, us=vals.get("us"))
with open(os.path.join(P, f"traffic_{R}.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print("wrote", sorted(os.listdir(P)))
