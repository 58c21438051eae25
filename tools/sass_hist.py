#!/usr/bin/env python3
"""Static SASS size attribution: which source call sites / functions a kernel's instructions come from.
  cuobjdump -xelf all lib.so; nvdisasm --print-line-info-inline x.cubin > dis.txt
  python tools/sass_hist.py dis.txt KERNEL_SUBSTRING [depth]
Prints instruction counts by (outermost call-site line in the kernel) and by the function `depth` levels below it."""
import bisect, collections, glob, os, re, sys
dis, kern = sys.argv[1], sys.argv[2]
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
by_site = collections.Counter(); by_fn = collections.Counter(); by_inner = collections.Counter(); total = 0
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        fresh = True
        if not chain: continue
        total += 1
        by_site[chain[-1]] += 1
        lvl = chain[-1 - depth] if len(chain) > depth else chain[0]
        by_fn[(lvl[0], fn(*lvl))] += 1
        by_inner[(chain[0][0], fn(*chain[0]))] += 1
print(kern, "instructions:", total)
print("-- by call site in the kernel body"); [print("%6d %5.1f%%  %s:%d" % (c, 100 * c / total, f, n)) for (f, n), c in by_site.most_common(25)]
print("-- by function %d level(s) below the kernel" % depth); [print("%6d %5.1f%%  %s:%s" % (c, 100 * c / total, f, n)) for (f, n), c in by_fn.most_common(25)]
print("-- by innermost function"); [print("%6d %5.1f%%  %s:%s" % (c, 100 * c / total, f, n)) for (f, n), c in by_inner.most_common(25)]
