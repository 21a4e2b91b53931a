"""Turn an .ncu-rep (ncu --set full) into the markdown/JSON summaries committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary
"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = {
    "time_us": ("gpu__time_duration.sum", 1e3 if units[hdr.index("gpu__time_duration.sum")] == "ms" else 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_read_MB": ("dram__bytes_read.sum", None),
    "dram_write_MB": ("dram__bytes_write.sum", None),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "regs": ("launch__registers_per_thread", 1),
    "inst_executed_M": ("smsp__inst_executed.sum", 1e-6),
    "pipe_fma_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "pipe_alu_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "pipe_xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    "pipe_lsu_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
}
def scale_bytes(unit):
    return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("grpg::", ""),
         "grid": r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""}
    for k, (m, sc) in cols.items():
        if m not in hdr:
            continue
        i = hdr.index(m)
        try:
            v = float(r[i])
        except ValueError:
            continue
        d[k] = round(v * (scale_bytes(units[i]) if sc is None else sc), 3)
    # every per-pipe utilisation metric the report holds (names differ between architectures)
    d["pipes"] = {}
    for i, m in enumerate(hdr):
        if "inst_executed_pipe_" in m and m.endswith("pct_of_peak_sustained_active"):
            try:
                d["pipes"][m.split("inst_executed_pipe_")[1].split(".")[0]] = round(float(r[i]), 2)
            except ValueError:
                pass
    res.append(d)
json.dump(res, open(out + ".json", "w"), indent=1)
keys = ["kernel", "time_us", "issue_active_pct", "dram_throughput_pct", "dram_read_MB", "dram_write_MB", "warps_active_pct",
        "regs", "inst_executed_M", "pipe_fma_pct", "pipe_alu_pct", "pipe_xu_pct", "pipe_lsu_pct"]
with open(out + ".md", "w") as f:
    f.write(f"# ncu --set full summary ({rep})\n\nPer-launch values; captured with --clock-control none on a B200 "
            "(tools/one_step.py: 2 M Gaussians, 1920x1280).  Times under ncu are serialised and cold-cache: compare "
            "shares, not absolutes (bench.py holds the CUDA-event numbers).\n\n")
    f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
    for d in res:
        f.write("| " + " | ".join(str(d.get(k, "")) for k in keys) + " |\n")
print("wrote", out + ".md", len(res), "kernels")
