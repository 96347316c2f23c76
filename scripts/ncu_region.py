"""Per-region stall breakdown of an .ncu-rep source page: python scripts/ncu_region.py rep lo_off hi_off [top]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; lo = int(sys.argv[2], 16); hi = int(sys.argv[3], 16); top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[idx['# Samples']].strip().isdigit()]
base = int(data[0][idx['Address']], 16)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
sel = [r for r in data if lo <= int(r[idx['Address']], 16) - base < hi]
tot = Counter(); n = 0
for r in sel:
    n += int(r[idx['# Samples']])
    for s in stalls:
        tot[s[6:]] += int(r[idx[s]] or 0)
print("region samples", n, "instrs", len(sel))
for k, v in tot.most_common(12):
    print(f"  {k:24s} {v:7d} {100*v/max(n,1):5.1f}%")
for r in sorted(sel, key=lambda r: -int(r[idx['# Samples']]))[:top]:
    st = {s[6:]: int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0) > 0.2 * int(r[idx['# Samples']] or 1)}
    print(f"  {int(r[idx['Address']],16)-base:#7x} {int(r[idx['# Samples']]):6d} {r[idx['Source']][:70]:70s} {st}")
