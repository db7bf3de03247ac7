#!/usr/bin/env python
"""tools/ncu_regions.py REP LIB KERNEL SYM LINE... -- warp-instruction share, lanes and hot code size of the regions of a kernel's
body delimited by source lines of the kernel's own file, with inlined code attributed to the OUTERMOST frame (nvdisasm -gi): where an
ncu source-page capture spends its instructions when most of the code is inlined helpers (development aid).

    python tools/ncu_regions.py gpurun_out/x.ncu-rep cuda-photon-mapper_b200/libpmb200.so trace_kernel trace_kernelILb0ELb1 496 570 719
"""
import sys, csv, io, os, re, subprocess, tempfile, collections


def outer_lines(lib, sym):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    table = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        insec, cur = False, ("?", 0)
        for ln in txt.splitlines():
            if ln.startswith("//-----"):
                insec = sym in ln and ".text." in ln
                continue
            if not insec:
                continue
            m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))   # the last one before an instruction is the outermost frame
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
            if m:
                table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, lib, kernel, sym = sys.argv[1:5]
    bounds = [int(x) for x in sys.argv[5:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
    ia, iex, ithr, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    tab = outer_lines(lib, sym)
    base = min(int(r[ia], 16) for r in body)
    reg = collections.defaultdict(lambda: [0.0, 0.0, 0, set(), 0.0])
    for r in body:
        off = int(r[ia], 16) - base
        f, l = tab.get(off, ("?", 0))
        k = sum(1 for b in bounds if l >= b)
        e, t = float(r[iex] or 0), float(r[ithr] or 0)
        g = reg[k]
        g[0] += e; g[1] += t; g[2] += 1; g[4] += float(r[isamp] or 0)
        if e > 0:
            g[3].add(off // 128)
    tot = sum(g[0] for g in reg.values()); ts = sum(g[4] for g in reg.values())
    for k in sorted(reg):
        g = reg[k]
        print("from line %4d: warp-instr %5.1f%% (%.3g)  samples %5.1f%%  lanes %4.1f  static %4d  executed code %.1f KB" % (
            bounds[k - 1] if k else 0, 100 * g[0] / tot, g[0], 100 * g[4] / ts, g[1] / max(g[0], 1), g[2], len(g[3]) * 128 / 1024))


if __name__ == "__main__":
    main()
