"""Hot source lines (warp-stall samples / instructions executed) per kernel of an .ncu-rep
captured with --import-source on.  Usage: ncu_hot_lines.py report.ncu-rep [kernel-index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections, cur = [], None
for r in rows:
    if r and r[0] in ("Kernel Name", "Function Name"):
        cur = {"name": r[1], "hdr": None, "lines": []}
        sections.append(cur)
        continue
    if cur is None:
        continue
    if r and r[0] in ("Address", "Line No"):
        cur["hdr"] = r
        continue
    if cur["hdr"] and len(r) >= len(cur["hdr"]) and r[0].isdigit():
        d = dict(zip(cur["hdr"], r))
        try:
            cur["lines"].append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), int(r[0]), d["Source"].strip()[:110]))
        except ValueError:
            pass
# the cuda,sass view lists each kernel twice (cuda view, sass view); keep those with lines
secs = [s for s in sections if s["lines"]]
s = secs[which]
tot_s = sum(l[0] for l in s["lines"]); tot_i = sum(l[1] for l in s["lines"])
print(f"kernel {which}/{len(secs)}: {s['name'][:100]}\nsamples {tot_s}, warp instructions {tot_i}")
for l in sorted(s["lines"], reverse=True)[:top]:
    print(f"{l[0]:7d} {100*l[0]/max(tot_s,1):5.1f}%  inst {l[1]:10d}  L{l[2]:<5d} {l[3]}")
