"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (tcgen05 = UTC*MMA, tcgen05.ld/st = LDTM/STTM,
TMA = UTMALDG/UTMASTG/UBLKCP, legacy mma.sync = HMMA) in the built library.  usage: python tools/sass_histogram.py > profiles/..."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "landiff_b200" / "lib" / "liblandiff_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "LDGSTS"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"SASS mnemonic counts per kernel of {lib.name} (cuobjdump -sass; sm_100a)")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for name, c in counts.items():
    print(f"{name[:78]:78s} {c['total']:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL':78s} {tot['total']:6d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
