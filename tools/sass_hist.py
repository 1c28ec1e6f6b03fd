"""Tuning aid: opcode histogram (warp instructions executed, stall samples) from an `ncu --page source --csv` dump."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ex, st = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit() or not r[0].startswith('0x'): continue
    s = r[iS].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    op = s.split()[0].rstrip(";") if s else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED", "UBLKCP", "SYNCS")) else op.split(".")[0]
    ex[op] += int(r[iE] or 0); st[op] += int(r[iW] or 0)
te, ts = sum(ex.values()), sum(st.values())
print(f"total warp instructions {te}, stall samples {ts}")
for op, n in ex.most_common(28):
    print(f"  {op:14s} {n:12d} {100.0 * n / te:5.1f}%   samples {100.0 * st[op] / max(ts, 1):5.1f}%")
