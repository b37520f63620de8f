"""Stall-reason totals per region of the SASS (ncu source page). usage: ncu_stalls.py rep kidx a:b [a:b ...]"""
import csv, io, subprocess, sys
rep, kidx = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "Kernel Name": cur = []; tables.append(cur)
    elif cur is not None: cur.append(row)
t = tables[kidx]; hdr, rows = t[0], t[1:]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
regions = [tuple(int(x) for x in a.split(":")) for a in sys.argv[3:]] or [(0, len(rows))]
for a, b in regions:
    tot = {hdr[c]: sum(int(r[c] or 0) for r in rows[a:b]) for c in cols}
    s = sum(tot.values())
    inst = sum(int(r[hdr.index("Instructions Executed")]) for r in rows[a:b])
    print(f"[{a},{b}) samples {s} winst {inst}: " + ", ".join(f"{k[6:]} {100 * v / max(1, s):.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > s))
