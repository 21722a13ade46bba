"""Source lines of one kernel sorted by executed warp instructions (needs --import-source on).
Usage: ncu_inst_lines.py report.ncu-rep [top]"""
import subprocess, csv, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []; sec = 0; fname = None
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if r and r[0] == "File Path": fname = r[1]
    if r and r[0] == "Function Name": sec += 1; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) >= len(hdr) and r[0].isdigit() and sec >= 1:
        d = dict(zip(hdr, r)); lines.append((I(d["Instructions Executed"]), I(d["# Samples"]), int(r[0]), fname))
tot = sum(l[0] for l in lines); tots = sum(l[1] for l in lines)
print(rep, "warp instructions", tot, "samples", tots)
srcs = {}
for ins, smp, ln, fn in sorted(lines, reverse=True)[:top]:
    if fn not in srcs:
        try: srcs[fn] = open(fn).read().split("\n")
        except OSError: srcs[fn] = []
    t = srcs[fn][ln - 1].strip()[:85] if ln - 1 < len(srcs[fn]) else "?"
    print(f"{ins:10d} {100*ins/tot:5.1f}% smp {100*smp/max(tots,1):5.1f}% {fn.split('/')[-1]}:{ln}: {t}")
