"""Section-level instruction breakdown of a kernel from an .ncu-rep (ncu --set full --import-source on): the SASS lines
are grouped into plateaus of equal execution count (loop nests) and summed.

    python tools/source_breakdown.py gpurun_out/prof_final.ncu-rep blend_bwd [min_share]
"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1].split("(")[0]
hdr = rows[1]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Avg. Threads Executed")
cstall = hdr.index("# Samples")
lines = []
for r in rows[2:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    lines.append((r[cs].strip(), int(r[ci]), float(r[ct] or 0), int(r[cstall] or 0)))
total = sum(l[1] for l in lines)
tot_samples = sum(l[3] for l in lines) or 1
print(f"# {name}: {total / 1e6:.1f} M warp instructions, {len(lines)} SASS lines")
print("| section (first .. last instruction) | executions per line | SASS lines | warp instr. (M) | share | avg. active threads | stall-sample share |")
print("|---|---|---|---|---|---|---|")
# plateaus: consecutive lines whose execution counts stay within 3 % of the running plateau value
sec = []
for l in lines:
    if sec and abs(l[1] - sec[-1]["ref"]) <= 0.03 * max(sec[-1]["ref"], 1):
        s = sec[-1]; s["n"] += 1; s["sum"] += l[1]; s["thr"] += l[2] * l[1]; s["last"] = l[0]; s["samples"] += l[3]
    else:
        sec.append({"ref": l[1], "n": 1, "sum": l[1], "thr": l[2] * l[1], "first": l[0], "last": l[0], "samples": l[3]})
for s in sec:
    if s["sum"] < 0.004 * total:
        continue
    print(f"| `{s['first'][:38]}` .. `{s['last'][:38]}` | {s['ref'] / 1e6:.3f} M | {s['n']} | {s['sum'] / 1e6:.1f} | "
          f"{100 * s['sum'] / total:.1f} % | {s['thr'] / max(s['sum'], 1):.1f} | {100 * s['samples'] / tot_samples:.1f} % |")
