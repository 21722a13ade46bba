"""Builds profiles/<out>_kernels.md and profiles/<out>_traffic.json from `ncu --set full`
captures of the finest-level launches of every kernel family.
Usage: make_profiles.py <out prefix, e.g. r2_ncu_c5> "<title line>" <report stems under gpurun_out/ ...>
(no arguments: the round-1 set)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[1] if len(sys.argv) > 1 else "r1_ncu"
TITLE = sys.argv[2] if len(sys.argv) > 2 else "513^3 fp32 bench field (scripts/prof_one.py); one launch per family; durations under ncu are cold-cache."
REPS = sys.argv[3:] if len(sys.argv) > 3 else ["r1_final_levels", "r1_final_huff", "r1_final_restore"]
FAM = [("norm_partial", "norm"), ("coef3d", "coef"), ("masstrans3d", "mass_trans"), ("thomas_tma", "thomas_contig"),
       ("thomas_smem", "thomas_contig"),
       ("thomas_strided", "thomas_strided"), ("restore3d", "restore"), ("quantize_linear", "quantize_hist"),
       ("codebook", "codebook"), ("chunk_bits", "chunk_bits"), ("encode_kernel", "encode"), ("encode_serial", "encode"), ("encode_sub", "encode"),
       ("decode_fast", "decode"), ("decode_serial", "decode"), ("decode_ring", "decode"), ("thomas_stream", "thomas_contig"),
       ("sort_outliers", "outlier_sort")]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]
traffic, sections, md = {}, {}, [f"# {OUT}: `ncu --set full --clock-control none` of the largest launch of every kernel family",
                   "", TITLE, ""]
for rep in REPS:
    # a report, or the `ncu -i report --page raw --csv` export of one (made on the GPU box
    # when the report itself is too large to bring back)
    path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if os.path.exists(os.path.join(ROOT, "gpurun_out", rep + ".csv")):
        out = open(os.path.join(ROOT, "gpurun_out", rep + ".csv")).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        fam = next((f for k, f in FAM if k in d["Kernel Name"]), None)
        if fam is None:
            continue
        # keep the longest launch of the family (the finest level)
        us0 = to_us(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
        if fam in traffic and traffic[fam]["ncu_duration_us"] >= us0:
            continue
        seen.add(fam)
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        us = to_us(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
        traffic[fam] = {"kernel": d["Kernel Name"][:80], "grid": d.get("Grid Size"), "block": d.get("Block Size"),
                        "dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_duration_us": us,
                        "dram_gbs_under_ncu": (rd + wr) / us / 1e3,
                        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                        "registers": int(d["launch__registers_per_thread"]), "source": rep + ".ncu-rep"}
        sec = [f"## {fam}: `{d['Kernel Name'][:100]}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n",
               "| metric | value | unit |\n|---|---:|---|"]
        for k in KEYS:
            if k in d and d[k] != "":
                sec.append(f"| {k} | {d[k]} | {u[k]} |")
        sec.append("")
        sections[fam] = sec  # the longest launch of the family so far
for fam in traffic:
    md += sections[fam]
json.dump(traffic, open(os.path.join(ROOT, "profiles", OUT + "_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", OUT + "_kernels.md"), "w").write("\n".join(md) + "\n")
for f, t in traffic.items():
    print(f"{f:16s} {t['ncu_duration_us']:9.1f} us  dram {(t['dram_read_bytes']+t['dram_write_bytes'])/1e6:8.1f} MB  {t['dram_gbs_under_ncu']:7.0f} GB/s  issue {t['issue_active_pct']:5.1f}%  regs {t['registers']}")
