#!/usr/bin/env python
"""tools/summarise_r2.py -- turn the round-2 ncu captures in gpurun_out/ (tools/profile_r2.sh) into the tracked evidence under profiles/:
  r2_launches.csv / .md    the ncu launch list of the bench command (per-launch durations; shares of a step)
  r2_full_raw.csv          `ncu --page raw --csv` of the --set full capture of every Mode A kernel of one frame
  r2_knn_raw.csv           the same for the Mode B gather kernel (knn_render_kernel)
  r2_kernels.md            the key metrics of both as a table
  r2_trace_ncu.json        what bench.py reads for roofline.traffic / roofline.ncu (dominant kernel, same configuration)
  r2_sass_*.txt            opcode histogram + excerpt of the SASS of trace_kernel<false> and knn_render_kernel<2>"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
PHOTONS = 16777216
short = lambda s: re.sub(r"<.*", "", s.split("(")[0].replace("pm::", "").replace("void ", ""))

# ---- launch list ----------------------------------------------------------------------------------------
src = open(os.path.join(go, "r2_launches.csv")).read()
open(os.path.join(prof, "r2_launches.csv"), "w").write("\n".join(l for l in src.splitlines() if l.startswith('"')) + "\n")
rows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    d.setdefault(short(r[ki]), []).append(v)
step = ("trace_kernel", "fold_volume_kernel", "build_map_kernel", "build_tables_kernel", "render_kernel")
tot = sum(sum(d[k]) / len(d[k]) for k in step if k in d)
with open(os.path.join(prof, "r2_launches.md"), "w") as f:
    f.write("# r2: ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps 10 --warmup 5 --no-extras "
            "--no-ref-cuda --no-cpu-baseline` (tools/profile_r2.sh; raw: r2_launches.csv)\n\n"
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
            "| kernel | launches captured | avg us | share of a step (our kernels) |\n|---|---|---|---|\n")
    for k, v in d.items():
        a = sum(v) / len(v)
        f.write("| %s | %d | %.1f | %s |\n" % (k[:70], len(v), a, ("%.1f%%" % (100 * a / tot)) if k in step else "(outside the step / runtime)"))
print(open(os.path.join(prof, "r2_launches.md")).read())

# ---- full captures ----------------------------------------------------------------------------------------
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed.sum", "thread instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %")]
seen = collections.OrderedDict()
for rep, raw in (("r2_full.ncu-rep", "r2_full_raw.csv"), ("r2_knn.ncu-rep", "r2_knn_raw.csv")):
    path = os.path.join(go, rep)
    if not os.path.exists(path):
        print("missing", path); continue
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(prof, raw), "w").write(out)
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ni = h.index("Kernel Name")
    for r in rows[2:]:
        k = short(r[ni])
        if k not in seen:
            seen[k] = {n: (r[h.index(n)], units[h.index(n)]) for n, _ in want if n in h}
with open(os.path.join(prof, "r2_kernels.md"), "w") as f:
    f.write("# r2: ncu --set full, one launch per kernel (raw: r2_full_raw.csv, r2_knn_raw.csv; commands: tools/profile_r2.sh)\n\n"
            "Mode A kernels: one frame of the bench configuration (16 777 216 photons, 1920x1080, media on).  knn_render_kernel: the Mode B "
            "gather of the same configuration (k = 50, 11 gathers per pixel).\n\n")
    f.write("| metric | " + " | ".join(seen) + " |\n|---|" + "---|" * len(seen) + "\n")
    for n, label in want:
        if not any(n in v for v in seen.values()):
            continue
        unit = next(v[n][1] for v in seen.values() if n in v)
        f.write("| %s [%s] | " % (label, unit) + " | ".join(seen[k].get(n, ("", ""))[0] for k in seen) + " |\n")
print(open(os.path.join(prof, "r2_kernels.md")).read())

def num(k, n):
    return float(seen[k][n][0].replace(",", "")) if k in seen and n in seen[k] else None
def to_bytes(k, n):
    v, u = num(k, n), seen[k][n][1] if k in seen and n in seen[k] else ""
    return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
if "trace_kernel" in seen:
    wi, ti = num("trace_kernel", "smsp__inst_executed.sum"), num("trace_kernel", "smsp__thread_inst_executed.sum")
    if ti is None and wi:   # warp instructions x average active lanes
        ti = wi * num("trace_kernel", "smsp__thread_inst_executed_per_inst_executed.ratio")
    json.dump({"photons": PHOTONS, "source": "profiles/r2_full_raw.csv (ncu --set full, tools/profile_r2.sh)", "kernel": "trace_kernel",
               "dram_bytes_per_launch": to_bytes("trace_kernel", "dram__bytes_read.sum") + to_bytes("trace_kernel", "dram__bytes_write.sum"),
               "issue_slots_busy_pct": num("trace_kernel", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warp_instructions": wi, "thread_instructions_per_photon": (ti / PHOTONS) if ti else None,
               "active_lanes_per_instruction": num("trace_kernel", "smsp__thread_inst_executed_per_inst_executed.ratio")},
              open(os.path.join(prof, "r2_trace_ncu.json"), "w"), indent=1)
    print(open(os.path.join(prof, "r2_trace_ncu.json")).read())

# ---- SASS ---------------------------------------------------------------------------------------------------
lib = os.path.join(ROOT, "cuda-photon-mapper_b200", "libpmb200.so")
for sym, name in (("trace_kernelILb0", "r2_sass_trace_kernel.txt"), ("knn_render_kernelILi2", "r2_sass_knn_render_kernel.txt")):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = txt.split("Function : ")
    blk = next((b for b in blocks if sym in b.split("\n")[0]), None)
    if blk is None:
        continue
    lines = [l for l in blk.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", re.sub(r"\s+/\*[0-9a-f]{4,}\*/\s+", "", l)).split()[0].split(".")[0].rstrip(";") for l in lines)
    with open(os.path.join(prof, name), "w") as f:
        f.write("cuobjdump -sass cuda-photon-mapper_b200/libpmb200.so, function %s\n%d instructions; opcode histogram:\n" % (blk.split("\n")[0], len(lines)))
        for op, c in ops.most_common():
            f.write("  %-12s %d\n" % (op, c))
        f.write("\nfirst 120 instructions:\n" + "\n".join(lines[:120]) + "\n")
    print(name, len(lines), ops.most_common(12))
