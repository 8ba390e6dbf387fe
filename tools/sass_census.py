"""SASS census of libaddk.so per kernel: the mnemonics that prove tcgen05 / TMEM / TMA are in the shipped binary
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
SYNCS = mbarrier).  Runs without a GPU:   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "add_gym_b200", "libaddk.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FFMA", "STG.E.128", "LDG.E.128",
      "STS.128", "LDS.128", "REDUX", "SHFL"]
counts, cur, regs = collections.OrderedDict(), None, {}
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
    for k in MN:
        if op.startswith(k):
            counts[cur][k] += 1
            break
    counts[cur]["_total"] += 1
print("# SASS census of %s (sm_100a), cuobjdump -sass; columns = instruction counts per kernel" % os.path.basename(lib))
demangle = subprocess.run(["c++filt"], input="\n".join(counts.keys()), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for (fn, c), name in zip(counts.items(), demangle):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    cols = " ".join("%s=%d" % (k, c[k]) for k in MN if c[k])
    print("%-64s total=%-6d %s" % (name[:64], c["_total"], cols))
    tot.update(c)
print("# library totals: " + " ".join("%s=%d" % (k, tot[k]) for k in MN if tot[k]))
