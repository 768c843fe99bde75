"""Per-kernel SASS opcode summary of confignet_b200/lib/libconfignet_b200.so (cuobjdump -sass), so that the Blackwell-native
instructions can be audited without the binary:  python scripts/sass_summary.py > profiles/r02_sass_opcodes.txt
  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA unit, 1-D),
  UTMASTG / UTMALDG = cp.async.bulk.tensor store / load (TMA tensor), SYNCS = mbarrier ops, HMMA would be a legacy mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "confignet_b200", "lib", "libconfignet_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMASTG", "UTMALDG", "UTMAPF", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS",
        "SHFL", "ATOM", "RED", "LDC"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print("SASS opcode counts per kernel of %s (cuobjdump -sass; cubin architectures: %s)" % (os.path.relpath(lib, ROOT), ", ".join(arch)))
print("%-74s %s" % ("kernel", " ".join("%7s" % k for k in KEYS)))
tot = collections.Counter()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0][:72]
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f, flags=re.M)
    c = collections.Counter()
    for o in ops:
        for k in KEYS:
            if o == k or o.startswith(k + ".") or (k in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "LDC", "SHFL", "SYNCS") and o.startswith(k)):
                c[k] += 1
                break
    tot.update(c)
    print("%-74s %s" % (dem, " ".join("%7d" % c[k] for k in KEYS)))
print("%-74s %s" % ("TOTAL", " ".join("%7d" % tot[k] for k in KEYS)))
