"""Opcode histogram of a kernel's SASS (whole kernel, or an address range = one loop body).

    python scripts/sass_hist.py <object-or-.so> <kernel-substring> [lo_hex hi_hex]

Used offline (no GPU) to count the instructions a loop body issues per iteration.
"""
import collections
import re
import subprocess
import sys


def kernel_sass(path, name):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, ins = None, []
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        if cur is None or name not in cur:
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return ins


def main():
    path, name = sys.argv[1], sys.argv[2]
    ins = kernel_sass(path, name)
    funcs = sorted({f for _, _, f in ins})
    print("functions:", len(funcs))
    for f in funcs:
        body = [(a, t) for a, t, g in ins if g == f]
        print("==", f[:110], len(body), "instructions")
        if len(sys.argv) >= 5:
            lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
            body = [(a, t) for a, t in body if lo <= a <= hi]
            print("range", hex(lo), hex(hi), len(body))
        else:
            for a, t in body:
                if re.search(r"\bBRA\b", t):
                    m = re.search(r"0x([0-9a-f]+)", t)
                    tgt = int(m.group(1), 16) if m else -1
                    print("   branch", hex(a), "->", hex(tgt), "BACK" if tgt <= a else "", t[:60])
        h = collections.Counter()
        for a, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0]
            h[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "MUFU")) and "." in op else "")] += 1
        for k, v in h.most_common():
            print(f"   {k:14s} {v}")


if __name__ == "__main__":
    main()
