"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
per-kernel table (markdown).  Usage: summarize_launches.py launches.csv [title]"""
import csv, re, sys, collections

def short(name):
    m = re.search(r"(?:<unnamed>::|fused3d::)?(\w+)(?:<[^(]*>)?\(", name)
    base = m.group(1) if m else name[:40]
    t = re.search(r"(level_kernel<\w+, \d>|\w+_kernel<[^>(]*>)", name)
    return t.group(1) if t else base

def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    ours = collections.OrderedDict()
    other = 0.0
    for r in rows[1:]:
        name, ns = r[ki], float(r[vi].replace(",", ""))
        if "at::" in name or "elementwise" in name or "vectorized" in name or "reduce_kernel" in name:
            other += ns
            continue
        k = short(name)
        c = ours.setdefault(k, [0, 0.0, 0.0])
        c[0] += 1; c[1] += ns; c[2] = max(c[2], ns)
    tot = sum(c[1] for c in ours.values())
    print(f"### {title}\n")
    print(f"{len(rows)-1} launches captured; {tot/1e6:.3f} ms in mgard_b200 kernels, {other/1e6:.3f} ms in torch kernels (input generation / checks).\n")
    print("| kernel | launches | total ms | share | max launch ms |\n|---|---:|---:|---:|---:|")
    for k, c in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c[0]} | {c[1]/1e6:.3f} | {100*c[1]/tot:.1f}% | {c[2]/1e6:.3f} |")

main()
