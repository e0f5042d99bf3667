"""SASS instruction census of libsrb200.so per kernel (cuobjdump -sass): proves which kernels carry tcgen05 / TMA / TMEM code.
Usage: python tools/sass_census.py > profiles/r2_sass_census.txt"""
import os, re, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200", "srb200", "libsrb200.so")
COLS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "SYNCS", "REDUX", "HMMA",
        "FFMA", "ACQBULK", "UTMACCTL", "USETMAXREG"]

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", out)
if names:
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
counts, order, cur = {}, [], None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["instrs"] += 1
        for c in COLS:
            if op.startswith(c):
                counts[cur][c] += 1


def short(n):
    d = demangle.get(n, n).replace("(anonymous namespace)::", "").replace("srb::", "")
    d = re.sub(r"^void ", "", d)
    return re.sub(r"\(.*", "", d)


print("SASS instruction census of libsrb200.so (cuobjdump -sass, sm_100a), per kernel: tcgen05 MMA (UTCHMMA), TMA tensor loads (UTMALDG),")
print("TMA L2 prefetch (UTMAPF), TMEM loads (LDTM), tcgen05.commit (UTCBAR), bulk copies (UBLKCP), mbarrier ops (SYNCS), plain FFMA.\n")
print("%-36s %8s " % ("kernel", "instrs") + " ".join("%8s" % c for c in COLS))
tot = collections.Counter()
for n in order:
    c = counts[n]
    tot.update(c)
    print("%-36s %8d " % (short(n)[:36], c["instrs"]) + " ".join("%8d" % c[k] for k in COLS))
print("%-36s %8d " % ("TOTAL", tot["instrs"]) + " ".join("%8d" % tot[k] for k in COLS))
