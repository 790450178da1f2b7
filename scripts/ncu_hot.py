#!/usr/bin/env python
"""Top stall sites of one launch in an .ncu-rep (source page, SASS view).  usage: ncu_hot.py REP LAUNCH_INDEX [N]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"launch {k}: {tot} samples, {len(body)} SASS instructions")
agg = {s: sum(int(r[idx[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {s: v for s, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for i, r in sorted(enumerate(body), key=lambda x: -int(x[1][idx["# Samples"]] or 0))[:n]:
    st = {s[6:]: int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0)}
    print(f"{int(r[idx['# Samples']]):6d} {100*int(r[idx['# Samples']])/max(tot,1):5.1f}%  #{i:5d} {r[idx['Source']].strip()[:90]:90s} {st}")
