"""Opcode histogram per kernel of the shipped library (evidence that the hot path is tcgen05 / TMA native):
    python tools/sass_opcounts.py > profiles/r02_sass_opcounts.txt
Counts the SASS mnemonics that prove Blackwell-native code (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG / UBLKCP (TMA), and HMMA (legacy mma.sync -- expected: none)."""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "polus_b200", "libpolus_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "HMMA", "HGMMA")
per = defaultdict(Counter)
total_instr = Counter()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        total_instr[cur] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                full = op + (m.group(2) if w in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM") else "")
                per[cur][full] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(total_instr)} kernels, {sum(total_instr.values())} SASS instructions")
agg = Counter()
for k in sorted(total_instr, key=lambda k: -total_instr[k]):
    c = per[k]
    agg.update(c)
    if not c:
        continue
    print(f"\n{k}   [{total_instr[k]} instructions]")
    print("   " + "  ".join(f"{op}:{n}" for op, n in sorted(c.items())))
print("\n# library totals")
for op, n in sorted(agg.items()):
    print(f"   {op:28s} {n}")
print(f"   legacy tensor path (HMMA / HGMMA): {sum(n for op, n in agg.items() if op.startswith(('HMMA', 'HGMMA')))}")
