"""Print the SASS around the hottest instructions of an .ncu-rep, plus per-region sample totals.
    python scripts/ncu_hot.py rep.ncu-rep [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[idx['# Samples']].strip().isdigit()]
base = int(data[0][idx['Address']], 16)
S = [int(r[idx['# Samples']]) for r in data]
ex = [int(r[idx['Instructions Executed']] or 0) if 'Instructions Executed' in idx else 0 for r in data]
tot = sum(S)
print("total samples", tot, "instrs", len(data))
# cumulative by 64-instruction blocks
for i in range(0, len(data), 64):
    s = sum(S[i:i + 64]); e = sum(ex[i:i + 64])
    print(f"  off {i*16:#7x}  samples {s:6d} ({100*s/tot:5.1f}%)  warp-inst {e}")
order = sorted(range(len(data)), key=lambda i: -S[i])[:12]
for o in sorted(order):
    print("----", hex(int(data[o][idx['Address']], 16) - base), S[o])
    for i in range(max(0, o - ctx), min(len(data), o + ctx + 1)):
        print(f"   {int(data[i][idx['Address']],16)-base:#7x} {S[i]:6d} {ex[i]:9d}  {data[i][idx['Source']][:90]}")
