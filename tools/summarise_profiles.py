#!/usr/bin/env python
"""tools/summarise_profiles.py -- turn the ncu captures in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarise_profiles.py r1            # reads gpurun_out/r1_launches.csv and gpurun_out/r1_full.ncu-rep
Writes profiles/<tag>_launches.md (per-kernel share of a step), profiles/<tag>_kernels.md (ncu --set full key metrics
per kernel) and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).
"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
photons = int(sys.argv[2]) if len(sys.argv) > 2 else 16777216
go, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(prof, exist_ok=True)

# ---- launch list ----------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(go, tag + "_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)   # -> us
    name = re.sub(r"<.*", "", r[ki].split("(")[0].replace("pm::", "").replace("void ", ""))
    d.setdefault(name, []).append(v)
mine = {k: v for k, v in d.items() if k in ("trace_kernel", "fold_volume_kernel", "build_map_kernel", "build_tables_kernel", "render_kernel")}
tot = sum(sum(v) / len(v) for v in mine.values())
with open(os.path.join(prof, tag + "_launches.md"), "w") as f:
    f.write("# %s: ncu launch list (gpu__time_duration.sum, --clock-control none), `python bench.py --steps 10 --warmup 5 --no-ref-cuda --no-cpu-baseline --no-mode-b` (tools/profile_r1.sh)\n\n" % tag)
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    f.write("| kernel | launches captured | avg us | share of a step (our kernels) |\n|---|---|---|---|\n")
    for k, v in d.items():
        a = sum(v) / len(v)
        f.write("| %s | %d | %.1f | %s |\n" % (k[:70], len(v), a, ("%.1f%%" % (100 * a / tot)) if k in mine else ("(one-off table fill, outside the step)" if k == "mwc_table_kernel" else "(torch / memset)")))
print(open(os.path.join(prof, tag + "_launches.md")).read())

# ---- full capture -----------------------------------------------------------------------------------------
out = subprocess.run(["ncu", "-i", os.path.join(go, tag + "_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %")]
idx = {n: h.index(n) for n, _ in want if n in h}
ni = h.index("Kernel Name")
seen, traffic = {}, {}
for r in rows[2:]:
    k = re.sub(r"<.*", "", r[ni].split("(")[0].replace("pm::", "").replace("void ", ""))
    if k in seen:
        continue
    seen[k] = r
units = rows[1]
with open(os.path.join(prof, tag + "_kernels.md"), "w") as f:
    f.write("# %s: ncu --set full --clock-control none, one launch per kernel, bench configuration (16M photons, 1080p, media on)\n\n" % tag)
    f.write("| metric | " + " | ".join(seen) + " |\n|---|" + "---|" * len(seen) + "\n")
    for n, label in want:
        if n not in idx:
            continue
        f.write("| %s [%s] | " % (label, units[idx[n]]) + " | ".join(seen[k][idx[n]] for k in seen) + " |\n")
    f.write("\nSource: gpurun_out/%s_full.ncu-rep (scratch, not tracked); per-line views: `python tools/ncu_lines.py`.\n" % tag)

def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
for k, r in seen.items():
    a, b = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    traffic[k] = to_bytes(r[a], units[a]) + to_bytes(r[b], units[b])
issue = {}
for k, r in seen.items():
    try:
        issue[k] = float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
    except (KeyError, ValueError):
        pass
json.dump({"photons": photons, "source": tag + "_full.ncu-rep", "dram_bytes_per_launch": traffic, "issue_slots_busy_pct": issue},
          open(os.path.join(prof, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(prof, tag + "_kernels.md")).read())
