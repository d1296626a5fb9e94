#!/usr/bin/env python
"""Hottest SASS lines of each kernel in an `ncu --page source --csv --print-source sass` export.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
kern, hdr, body = None, None, []


def flush():
	if not kern or not body:
		return
	i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
	stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
	tot = sum(int(r[i_s] or 0) for r in body)
	agg = {h: sum(int(r[j] or 0) for r in body) for j, h in stall_cols}
	print(f"=== {kern[:90]}  samples {tot}")
	print("   ", {h[6:]: round(100.0 * v / max(tot, 1), 1) for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
	ranked = sorted(range(len(body)), key=lambda k: -int(body[k][i_s] or 0))[:top]
	for k in sorted(ranked):
		r = body[k]
		why = max(stall_cols, key=lambda jh: int(r[jh[0]] or 0))[1][6:]
		print(f"  {int(r[i_s]):6d} {100.0 * int(r[i_s]) / max(tot, 1):5.1f}%  line {k:5d} exec={r[i_ex]:>8s} {why:10s} {r[i_src].strip()[:90]}")


for r in rows:
	if r and r[0] == "Kernel Name":
		flush()
		kern, hdr, body = r[1], None, []
	elif r and r[0] == "Address":
		hdr = r
	elif hdr is not None and r:
		body.append(r)
flush()
