"""Writes section 6a of DESIGN.md (between the NUMBERS markers) from the final bench lines
under profiles/: r2_final_bench.json (N = 1), r2_final_bench_n{2,4,8}.json and
r2_final_bench_reference.json when present."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda f: os.path.join(ROOT, "profiles", f)


def load(f):
    try:
        return json.load(open(P(f)))
    except Exception:
        return None


b = load("r2_final_bench.json")
out = ["### 6a. Numbers of the final code", ""]
if b:
    c2 = b.get("c2") or {}
    e2e = b.get("e2e") or {}
    cb = b.get("cpu_baseline") or {}
    out += ["`bench.py` default run (N = 1, `profiles/r2_final_bench.json`): the whole 2049³ fp32 domain (34.4 GB, 8 MaxDim",
            "sub-domains of 257 planes) on one B200, REL 1e-3, s = ∞, Huffman, dict 8192, block 20480; C2 = 513³ fp32, one",
            "sub-domain, as a sub-record of the same run.", "",
            "| | C5 2049³ (8 sub-domains) | C2 513³ | C2 end of round 1 |", "|---|---:|---:|---:|",
            f"| compress (device resident) | {b['compress_ms']:.1f} ms = {b['compress_gbs']:.0f} GB/s | {c2.get('compress_ms', 0):.2f} ms = {c2.get('compress_gbs', 0):.0f} GB/s | 2.55 ms = 212 GB/s |",
            f"| decompress (device resident) | {b['decompress_ms']:.1f} ms = {b['decompress_gbs']:.0f} GB/s | {c2.get('decompress_ms', 0):.2f} ms = {c2.get('decompress_gbs', 0):.0f} GB/s | 2.39 ms = 226 GB/s |",
            f"| `value` (2·N·4 B / (t_c + t_d)) | **{b['value']:.0f} GB/s** | {c2.get('value', 0):.0f} GB/s | 219 GB/s |",
            f"| `e2e` (pinned host buffers, H2D + D2H inside) | {e2e.get('value', 0):.1f} GB/s ({e2e.get('ms_per_step', 0):.0f} ms per step) | – | 34 GB/s |",
            f"| compression ratio | {b['ratio']:.4f} | {c2.get('ratio', 0):.4f} | 2.7284 |",
            f"| max abs error / bound | {b['max_abs_error']:.2e} / {b['error_bound']:.2e} | {c2.get('max_abs_error', 0):.2e} / {c2.get('error_bound', 0):.2e} | 2.58e-05 / 1.30e-03 |",
            f"| kernel launches per step | {b['gpu_launches'] / b['steps']:.0f} | – | 86 |",
            f"| whole codec on `B_alg` (compress / decompress) | {100 * b['roofline_codec']['compress_frac']:.1f} % / {100 * b['roofline_codec']['decompress_frac']:.1f} % | {100 * c2.get('roofline_codec', {}).get('compress_frac', 0):.1f} % / {100 * c2.get('roofline_codec', {}).get('decompress_frac', 0):.1f} % | 4.4 % / 4.7 % |",
            ""]
    if cb:
        mc = cb.get("mgard_cpu") or {}
        out += [f"Reference on the box's host, same run: MGARD-X SERIAL, one 257³ block per core on {cb.get('cores')} cores at once: "
                f"{cb.get('value', 0):.3f} GB/s (compress {cb.get('compress_gbs', 0):.3f}, decompress {cb.get('decompress_gbs', 0):.3f}); "
                f"MGARD-CPU `mgard::compress` stages (OpenMP, {mc.get('cores')} threads) on C1 129³ fp64: {mc.get('value') or 0:.3f} GB/s.", ""]
    for tag, rec in (("C5, one 257×2049² sub-domain", b), ("C2 513³", c2)):
        pk = rec.get("roofline_kernels") or []
        fam = {f["kernel"]: f for f in rec.get("kernel_breakdown") or []}
        if not pk:
            continue
        out += [f"Per kernel family, {tag} (live CUDA-event times of the same run; algorithmic bytes of the finest launch;",
                "DRAM bytes of that launch from `ncu --set full` where captured):", "",
                "| kernel family | launches / step | ms / step | finest launch ms | algorithmic MB | achieved GB/s | of 6545 GB/s | ncu DRAM MB |",
                "|---|---:|---:|---:|---:|---:|---:|---:|"]
        for k in pk:
            f = fam.get(k["kernel"], {})
            tr = f"{k['traffic'] / 1e6:.0f}" if k.get("traffic") else "–"
            out.append(f"| {k['kernel']} | {f.get('launches_per_step', 0):.0f} | {f.get('ms_per_step', 0):.2f} | {k['launch_ms']:.3f} | "
                       f"{k['algorithmic_bytes'] / 1e6:.0f} | {k['achieved']:.0f} | {100 * k['frac']:.1f}% | {tr} |")
        rf = rec.get("roofline") or {}
        if rf.get("dram_efficiency"):
            out += ["", f"Time-weighted DRAM efficiency of these launches (ncu DRAM bytes ÷ live time ÷ 6545 GB/s): {100 * rf['dram_efficiency']:.1f} %."]
        out.append("")
rows = []
for n in (1, 2, 4, 8):
    r = b if n == 1 else load(f"r2_final_bench_n{n}.json")
    if r:
        rows.append((n, r))
if len(rows) > 1:
    v1 = rows[0][1]["value"]
    out += ["Strong scaling on C5 (`bench.py --gpus N` under torchrun, max over ranks, NCCL all-reduce + all-gather inside the timed region):", "",
            "| GPUs | compress ms | decompress ms | `value` GB/s | efficiency | codec on `B_alg` (c / d) | `e2e` GB/s |", "|---:|---:|---:|---:|---:|---:|---:|"]
    for n, r in rows:
        e = (r.get("e2e") or {}).get("value")
        out.append(f"| {n} | {r['compress_ms']:.1f} | {r['decompress_ms']:.1f} | {r['value']:.0f} | {r['value'] / (n * v1):.3f} | "
                   f"{100 * r['roofline_codec']['compress_frac']:.1f} % / {100 * r['roofline_codec']['decompress_frac']:.1f} % | {e and f'{e:.0f}' or '–'} |")
    out.append("")
ref = load("r2_final_bench_reference.json")
if ref:
    out += [f"`bench.py --impl reference` on the same box: {ref['value']:.3f} GB/s ({ref['cpu_baseline']['cores']} cores, "
            f"{ref['cpu_baseline']['sample']}).", ""]
text = "\n".join(out)
path = os.path.join(ROOT, "DESIGN.md")
s = open(path).read()
if "<!-- NUMBERS -->" in s:
    s = s.replace("<!-- NUMBERS -->", "<!-- NUMBERS BEGIN -->\n" + text + "\n<!-- NUMBERS END -->")
else:
    s = re.sub(r"<!-- NUMBERS BEGIN -->.*<!-- NUMBERS END -->", lambda m: "<!-- NUMBERS BEGIN -->\n" + text + "\n<!-- NUMBERS END -->", s, flags=re.S)
open(path, "w").write(s)
print(text)
