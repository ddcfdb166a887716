"""Static (no GPU) evidence for profiles/: per-kernel registers / shared memory / spills from
`ptxas -v`, and the SASS mnemonics that prove which kernels use the tcgen05 tensor pipe (UTC*MMA),
TMEM loads (LDTM), TMA (UTMALDG / UTMASTG / UBLKCP) and warp match (MATCH).

    python tools/static_evidence.py > profiles/static_r1.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from hiarch_b200 import build as B

MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "MATCH.ANY",
             "SYNCS", "UTCBAR", "REDUX", "RED.E.ADD", "ATOMS")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"hia::\(anonymous namespace\)::|hia::<unnamed>::", "", o) for o in out]


def main():
    print("# ptxas -v (sm_100a), every kernel of libhiarch_b200.so")
    rows = []
    for src in B.sources():
        cmd = [B.NVCC] + B.FLAGS + ["-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
        txt = subprocess.run(cmd, capture_output=True, text=True).stderr
        cur = None
        for line in txt.split("\n"):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                rows.append((os.path.basename(src), cur, int(m.group(1)), int(m.group(3) or 0), spill))
                cur = None
    names = demangle([r[1] for r in rows])
    print(f"{'file':18s} {'regs':>4s} {'smem(static)':>12s} {'stack/spill':>12s}  kernel")
    for (f, _, regs, smem, spill), name in zip(rows, names):
        sp = "-" if spill == (0, 0, 0) else "/".join(map(str, spill))
        print(f"{f:18s} {regs:4d} {smem:12d} {sp:>12s}  {name.split('(')[0][:100]}")
    print()
    print("# SASS mnemonics per kernel (cuobjdump -sass), kernels with none of them omitted")
    sass = subprocess.run(["cuobjdump", "-sass", B.LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            for mn in MNEMONICS:
                if mn in line:
                    counts[cur][mn] += 1
    keys = [k for k, c in counts.items() if c]
    for k, name in zip(keys, demangle(keys)):
        c = counts[k]
        print(f"{name.split('(')[0][:90]:90s} " + " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn]))


if __name__ == "__main__":
    main()
