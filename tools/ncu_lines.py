#!/usr/bin/env python
"""tools/ncu_lines.py -- per-source-line view of an ncu capture without the GUI.

Joins `ncu -i REP --page source --csv` (per-SASS-address samples / instruction counts) with the line table
that `nvdisasm -g` prints for the cubin extracted from the built library (compile with -lineinfo).

    python tools/ncu_lines.py REP.ncu-rep cuda-photon-mapper_b200/libpmb200.so trace_kernel [--top 40] [--inline]
"""
import argparse, collections, csv, io, os, re, subprocess, sys, tempfile


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    table = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        insec, cur = False, ("?", 0)
        for ln in txt.splitlines():
            if ln.startswith("//-----"):
                insec = kernel in ln and ".text." in ln
                continue
            if not insec:
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("lib"); ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest individual SASS instructions")
    ap.add_argument("--sym", default=None, help="substring of the MANGLED symbol whose SASS the capture maps to (default: the kernel name); "
                                                 "needed when several instantiations share the name, e.g. trace_kernelILb0")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "-k", "regex:" + a.kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
    ia, isamp, iex, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tab = line_table(a.lib, a.sym or a.kernel)
    base = min(int(r[ia], 16) for r in body)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, collections.Counter()])
    sass = []
    for r in body:
        off = int(r[ia], 16) - base
        (f, l), ins = tab.get(off, (("?", 0), "?"))
        s, e, t = float(r[isamp] or 0), float(r[iex] or 0), float(r[ithr] or 0)
        g = agg[(f, l)]
        g[0] += s; g[1] += e; g[2] += t
        for i in stall_cols:
            v = float(r[i] or 0)
            if v:
                g[3][hdr[i]] += v
        sass.append((s, e, t, off, f, l, ins))
    ts = sum(v[0] for v in agg.values()) or 1
    te = sum(v[1] for v in agg.values()) or 1
    tt = sum(v[2] for v in agg.values()) or 1
    print("total samples %.0f, warp instructions %.3g, thread instructions %.3g, avg active lanes %.1f" % (ts, te, tt, tt / te))
    srcs = {}
    print("%-16s %5s %7s %7s %6s  %-28s %s" % ("file", "line", "samp%", "inst%", "lanes", "top stalls", "source"))
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
        if f not in srcs:
            p = None
            for d in ("cuda-photon-mapper_b200/csrc", "include", "."):
                q = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
                if os.path.exists(q):
                    p = q
            srcs[f] = open(p).read().splitlines() if p else []
        src = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        st = ",".join("%s:%.0f" % (k[6:], 100 * c / max(v[0], 1)) for k, c in v[3].most_common(2))
        print("%-16s %5d %6.1f%% %6.1f%% %6.1f  %-28s %s" % (f[:16], l, 100 * v[0] / ts, 100 * v[1] / te, v[2] / max(v[1], 1), st, src))
    if a.sass:
        print("\nhottest SASS:")
        for s, e, t, off, f, l, ins in sorted(sass, key=lambda x: -x[0])[: a.top]:
            print("%6.2f%% %6.2f%% lanes %4.1f  /*%04x*/ %-60s %s:%d" % (100 * s / ts, 100 * e / te, t / max(e, 1), off, ins[:60], f, l))


if __name__ == "__main__":
    main()
