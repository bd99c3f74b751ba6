"""cuobjdump -sass sutro_b200/libsutro_b200.so | python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
Per-kernel histogram of the Blackwell-specific / tensor / async SASS opcodes, so that the
"this is tcgen05 + TMA code" claim can be checked without the binary."""
import collections
import re
import subprocess
import sys

want = re.compile(r"\b(UTCHMMA[\w.]*|UTCQMMA[\w.]*|UTMALDG[\w.]*|UTMASTG[\w.]*|UBLKCP[\w.]*|LDTM[\w.]*|"
                  r"STTM[\w.]*|UTCBAR[\w.]*|HMMA[\w.]*|LDSM[\w.]*|SYNCS[\w.]*|MUFU[\w.]*|UCGABAR[\w.]*|"
                  r"LDGSTS[\w.]*|UTMAPF[\w.]*)")
cur, hist = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for op in want.findall(line):
            hist[cur][op] += 1
names = {}
if hist:
    dem = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.split("\n")
    names = dict(zip(hist, dem))
print("# SASS opcode histogram of sutro_b200/libsutro_b200.so (cuobjdump -sass, sm_100a): tensor / async /")
print("# Blackwell-specific opcodes per kernel.  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA")
print("# tensor load, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,")
print("# SYNCS = mbarrier, HMMA / LDSM = the legacy mma.sync path, MUFU.EX2 = exp2.\n")
for fn in sorted(hist, key=lambda f: names.get(f, f)):
    name = names.get(fn, fn).replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\((?!int\)).*", "", name)[:140]
    print(name)
    print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(hist[fn].items())))
