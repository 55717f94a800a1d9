#!/usr/bin/env python3
"""Warp-stall samples of a captured kernel by SASS opcode (ncu --set full --import-source on).
  python tools/ncu_stalls.py <prof.ncu-rep> [kernel-substring]"""
import collections, csv, io, re, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
seen = set()
for blk in blocks:
    name, rest = blk.split("\n", 1)
    if want not in name or blk in seen:        # the source page lists every kernel once per view
        continue
    seen.add(blk)
    rows = list(csv.reader(io.StringIO(rest)))
    h, data = rows[0], rows[1:]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot, byop, cnt, execd = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for r in data:
        if len(r) < len(h):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        op = m.group(2) if m else r[ix["Source"]][:10]
        n, ex = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        cnt[op] += n
        execd[op] += ex
        for s in stalls:
            v = int(r[ix[s]])
            tot[s] += v
            byop[op][s] += v
    T, E = sum(cnt.values()), sum(execd.values())
    print("==", name[:110])
    print("   SASS instructions:", len(data), " warp instructions executed:", E, " samples:", T)
    print("   stalls %:", ", ".join(f"{k[6:]} {100 * v / T:.1f}" for k, v in tot.most_common(8)))
    for op, n in cnt.most_common(10):
        print(f"   {op:20s} samples {100 * n / T:5.1f}%  executed {100 * execd[op] / E:5.1f}%  ",
              ", ".join(f"{k[6:]} {100 * v / n:.0f}" for k, v in byop[op].most_common(4)))
