"""Per-kernel SASS opcode census of libfdsr.so: the mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
UTCHMMA[.2CTA] = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, UTCBAR = tcgen05.commit, HMMA would be
the legacy mma.sync path (there must be none).  python tools/sass_census.py [lib] > profiles/r2/sass_census.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "fastdiffsr_b200/libfdsr.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCATOMSWS", "HMMA",
        "MEMBAR.ALL.GPU", "CCTL.IVALL"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op.startswith(p):
            # the longer .2CTA forms are listed first so that they do not count as the plain form
            if p == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                continue
            if p == "UTCBAR" and op.startswith("UTCBAR.2CTA"):
                continue
            counts[cur][p] += 1
            break
demangle = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode census of {lib} (cuobjdump -sass), per kernel")
print("| kernel | " + " | ".join(pats) + " |")
print("|---|" + "---:|" * len(pats))
tot = collections.Counter()
for (k, c), name in zip(counts.items(), demangle):
    if not any(c.values()):
        continue
    name = re.sub(r"\(fdsr::ConvLayer, int\)|void |fdsr::", "", name)
    print(f"| `{name}` | " + " | ".join(str(c[p]) if c[p] else "" for p in pats) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot[p]) for p in pats) + " |")
