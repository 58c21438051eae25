#!/usr/bin/env python3
"""Join an `ncu --page source --csv --print-source sass` export with nvdisasm line info to attribute warp-stall samples and
executed instructions to source functions.
  python tools/sass_profile_join.py sass.csv[.gz] dis.txt KERNEL_MANGLED_SUBSTR KERNEL_INDEX [depth]
KERNEL_INDEX = which launch in the csv (0-based, among all "Kernel Name" blocks)."""
import bisect, collections, csv, glob, gzip, io, os, re, sys
path, dis, kern, kidx = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
blocks = txt.split('"Kernel Name"')[1:]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[kidx])))
print("kernel:", rows[0][1][:80])
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
prof = {}
for r in data:
    if len(r) < len(h): continue
    prof[int(r[ix["Address"]], 16) - base] = r
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
funcs = {}
for f in glob.glob(ROOT + "/shimmer_b200/csrc/*.cu*"):
    fl = []
    for n, l in enumerate(open(f), 1):
        m = re.match(r"^\s*(?:template\s*<[^>]*>\s*)?(?:SGD|__global__|static|inline|__device__)[^;=]*?\b(\w+)\s*\(", l)
        if m and m.group(1) not in ("__launch_bounds__", "if", "for", "while"):
            fl.append((n, m.group(1)))
        elif "__global__" in l:
            m = re.search(r"\)\s*(\w+)\s*\(", l)
            if m: fl.append((n, m.group(1)))
    funcs[os.path.basename(f)] = sorted(fl)
def fn(f, n):
    fl = funcs.get(f, []); k = bisect.bisect_right([x[0] for x in fl], n) - 1
    return fl[k][1] if k >= 0 else "?"
chain = []; fresh = True
S = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])   # samples, warp inst, thread inst, stall reasons
SITE = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = [0, 0, 0]
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m:
        fresh = True
        off = int(m.group(1), 16)
        r = prof.get(off)
        if r is None or not chain: continue
        smp = int(r[ix["# Samples"]]); wi = int(r[ix["Instructions Executed"]]); ti = int(r[ix["Thread Instructions Executed"]])
        lvl = chain[-1 - depth] if len(chain) > depth else chain[0]
        key = (lvl[0], fn(*lvl))
        e = S[key]; e[0] += smp; e[1] += wi; e[2] += ti
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v: e[3][c[6:]] += v
        s2 = SITE[chain[-1]]; s2[0] += smp; s2[1] += wi; s2[2] += ti
        tot[0] += smp; tot[1] += wi; tot[2] += ti
print("total samples %d, warp inst %d, lanes/inst %.1f" % (tot[0], tot[1], tot[2] / max(tot[1], 1)))
print("%7s %6s %6s %6s  %s   top stalls" % ("samples", "smp%", "inst%", "lanes", "function (%d below kernel)" % depth))
for k, e in sorted(S.items(), key=lambda x: -x[1][0])[:28]:
    st = ", ".join("%s %d%%" % (n, 100 * v / max(e[0], 1)) for n, v in e[3].most_common(3))
    print("%7d %5.1f%% %5.1f%% %6.1f  %s:%s   %s" % (e[0], 100 * e[0] / tot[0], 100 * e[1] / tot[1], e[2] / max(e[1], 1), k[0], k[1], st))
print("-- by call site in kernel body")
for k, e in sorted(SITE.items(), key=lambda x: -x[1][0])[:16]:
    print("%7d %5.1f%% %5.1f%% %6.1f  %s:%d" % (e[0], 100 * e[0] / tot[0], 100 * e[1] / tot[1], e[2] / max(e[1], 1), k[0], k[1]))
