import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import orc
from shimmer_b200 import create_integrator, scenes
from test_gpu_parity import _ray_set
sc = scenes.tiny_scene("glass").build()
integ = create_integrator("wavefront", {}, sc)
n = 1 << 17
o, d = _ray_set(sc, n, seed=1)
tmax = np.full(n, np.inf, np.float32)
ref, rst = orc.trace(sc, o, d, tmax)
for stats in (False, True):
    got = integ.trace(o, d, tmax, want_stats=stats)
    if stats: got, gst = got
    bad = np.nonzero(got["prim"] != ref["prim"])[0]
    print("stats", stats, "mismatches", len(bad), "of", n)
    for i in bad[:8]:
        print(i, "o", o[i], "d", d[i], "got", got[i], "ref", ref[i])
    if stats: print(gst.nodes_visited, rst.nodes_visited, gst.tris_tested, rst.tris_tested)
nodes = sc.arrays["nodes"]
print("n_nodes", len(nodes), "multi-prim leaves", int((nodes["n_prims"]>1).sum()), "max", int(nodes["n_prims"].max()))
