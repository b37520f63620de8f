"""Per-region digest of the ncu source page (SASS): samples, executed warp-instructions, avg active threads.
usage: ncu_source_regions.py rep [kernel-index] — regions are split at the SASS landmarks given in REGIONS (regexes)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# the csv holds one table per kernel launch, each starting with a "Kernel Name" row
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
t = tables[kidx]
hdr = t["rows"][0]
rows = t["rows"][1:]
ci = {h: i for i, h in enumerate(hdr)}
S, I, T = ci["# Samples"], ci["Instructions Executed"], ci["Thread Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
tot_s = sum(int(r[S]) for r in rows); tot_i = sum(int(r[I]) for r in rows); tot_t = sum(int(r[T]) for r in rows)
print(t["name"], "samples", tot_s, "warp-inst", tot_i, "avg threads %.2f" % (tot_t / max(1, tot_i)))
if "--dump" in sys.argv:
    for k, r in enumerate(rows):
        print(f"{k:5d} {int(r[S]):7d} {int(r[I]):10d} {int(r[T]) / max(1, int(r[I])):5.1f}  {r[ci['Source']].strip()}")
    sys.exit(0)
# automatic regions: split where executed count changes by > 20 %
reg_start = 0
def flush(a, b):
    s = sum(int(r[S]) for r in rows[a:b]); i = sum(int(r[I]) for r in rows[a:b]); th = sum(int(r[T]) for r in rows[a:b])
    if i == 0 and s == 0:
        return
    print(f"[{a:5d},{b:5d}) n={b - a:4d} exec/inst {i / max(1, b - a):11.0f} samples {s:7d} ({100 * s / tot_s:5.1f}%) winst {100 * i / tot_i:5.1f}% thr {th / max(1, i):5.1f}  {rows[a][ci['Source']].strip()[:50]}")
for k in range(1, len(rows) + 1):
    if k == len(rows):
        flush(reg_start, k); break
    a, b = int(rows[k - 1][I]), int(rows[k][I])
    if abs(a - b) > 0.2 * max(a, b, 1):
        flush(reg_start, k); reg_start = k
