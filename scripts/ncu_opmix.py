"""Summarise an `ncu --page source --csv` export: executed warp-instructions by opcode, stall samples."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iT, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT or not r[iE].isdigit(): continue
    if not re.match(r'^\s*(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+', r[iS]): continue
    s = r[iS].strip()
    parts = s.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    n = int(r[iE] or 0); ops[op] += n; thr[op] += int(r[iT] or 0); samp[op] += int(r[iSamp] or 0); tot += n
print("total warp-instructions", tot, "thread-inst", sum(thr.values()), "avg active", sum(thr.values())/max(1,tot))
ALU = {"LOP3","PRMT","FMNMX","FMNMX3","IADD3","IADD","SHF","SEL","FSEL","ISETP","FSETP","LEA","POPC","FLO","VIMNMX","VIMNMX3","BREV","SGXT","PLOP3","MOV","I2FP","FSET","ISET","LOP","SHL","SHR","VOTE","VOTEU"}
FMA = {"FADD","FMUL","FFMA","IMAD","FADD2","FMUL2","FFMA2","HFMA2","HADD2"}
a = sum(v for k, v in ops.items() if k in ALU); f = sum(v for k, v in ops.items() if k in FMA)
print(f"ALU-class {a} ({100*a/tot:.1f}%)  FMA-class {f} ({100*f/tot:.1f}%)  other {tot-a-f} ({100*(tot-a-f)/tot:.1f}%)")
for op, n in ops.most_common(40):
    print(f"{op:10s} {n:12d} {100*n/tot:6.2f}%  avg_thr {thr[op]/max(1,n):5.1f}  stall_samples {samp[op]}")
