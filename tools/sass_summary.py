"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mma.sync -> HMMA/IMMA, plus registers.

    python tools/sass_summary.py quanta_b200/libquanta_b200.so profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib, out = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "IMMA", "SYNCS", "LDGSTS", "ATOM", "RED", "MEMBAR"]
kern, counts, total = None, {}, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1).split(".")[0]
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k] += 1
regs = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
with open(out, "w") as f:
    f.write(f"# SASS mnemonic counts per kernel of {lib} (cuobjdump -sass; sm_100a)\n")
    f.write("# UTCHMMA/UTCIMMA = tcgen05.mma kind::f16 / kind::i8, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store,\n")
    f.write("# HMMA = mma.sync (warp-level tensor core), SYNCS = mbarrier ops\n")
    for k, name in zip(counts, demangled):
        c = counts[k]
        if not any(c[x] for x in KEYS[:10]):
            continue
        short = re.sub(r"\(.*", "", name)[-120:]
        r = regs.get(k, ("?", "?"))
        f.write(f"\n{short}\n    instructions {total[k]}, registers {r[0]}, static smem {r[1]}\n    " +
                "  ".join(f"{x} {c[x]}" for x in KEYS if c[x]) + "\n")
print(open(out).read()[:3000])
