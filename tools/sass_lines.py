#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (nvdisasm line info), optionally weighted by the
per-instruction execution counts of an `ncu --page source --csv --print-source sass` dump.
    python tools/sass_lines.py <cubin> <kernel-substring> [ncu_sass.csv]"""
import csv
import re
import subprocess
import sys


def main():
    cubin, pat = sys.argv[1], sys.argv[2]
    ncu = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur_fun, line, file_ = None, None, None
    insts = []  # (offset_index, file, line, text)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fun = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            file_, line = m.group(1).split("/")[-1], int(m.group(2))
            continue
        if cur_fun and pat in cur_fun:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                insts.append((int(m.group(1), 16), file_, line, m.group(2)))
    if not insts:
        print("no instructions matched", pat)
        return
    weights = None
    if ncu:
        rows = list(csv.reader(open(ncu)))
        h = rows[1]
        ia, ie, it = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        d = [(int(r[ia], 16), int(r[ie]), int(r[it])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
        base = d[0][0]
        weights = {a - base: (e, t) for a, e, t in d}
    agg = {}
    for off, f, l, t in insts:
        k = (f, l)
        a = agg.setdefault(k, [0, 0, 0])
        a[0] += 1
        if weights and off - insts[0][0] in weights:
            e, th = weights[off - insts[0][0]]
            a[1] += e
            a[2] += th
    tot_e = sum(a[1] for a in agg.values()) or 1
    print("static instructions:", len(insts))
    for (f, l), a in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1] or 0)):
        if weights:
            print("%-28s %5s  static %4d  exec %12d (%5.2f%%)  avg thr %5.1f" % (f, l, a[0], a[1], 100.0 * a[1] / tot_e, a[2] / max(a[1], 1)))
        else:
            print("%-28s %5s  static %4d" % (f, l, a[0]))


if __name__ == "__main__":
    main()
