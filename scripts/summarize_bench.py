"""Prints the headline numbers and the kernel tables of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
keys = ("value", "compress_ms", "decompress_ms", "compress_gbs", "decompress_gbs", "ratio", "bound_ok", "n_gpus")
print({k: d.get(k) for k in keys})
print("roofline", d.get("roofline"))
print("codec", d.get("roofline_codec"))
print("e2e", d.get("e2e"))
for f in d.get("kernel_breakdown", [])[:14]:
    print("  %-16s n=%6.1f  ms/step=%8.3f  max=%7.3f" % (f["kernel"], f["launches_per_step"], f["ms_per_step"], f["max_launch_ms"]))
c2 = d.get("c2")
if c2:
    print("C2", {k: c2.get(k) for k in keys if k in c2})
    print("C2 roofline", c2.get("roofline"))
    for f in c2.get("kernel_breakdown", [])[:14]:
        print("  %-16s n=%6.1f  ms/step=%8.3f  max=%7.3f" % (f["kernel"], f["launches_per_step"], f["ms_per_step"], f["max_launch_ms"]))
print("cpu", d.get("cpu_baseline"))
