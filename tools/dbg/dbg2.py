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
for i in (23667, 23907):
    for rep in (1, 2, 33, 64):
        oo = np.repeat(o[i:i+1], rep, 0); dd = np.repeat(d[i:i+1], rep, 0); tt = np.full(rep, np.inf, np.float32)
        got, gst = integ.trace(oo, dd, tt, want_stats=True)
        ref, rst = orc.trace(sc, oo, dd, tt)
        print(i, rep, got["prim"][:3], got["t"][:3], ref["prim"][:1], ref["t"][:1], gst.nodes_visited, rst.nodes_visited, gst.tris_tested, rst.tris_tested)
    # neighbours in original order
    lo = max(0, i - 40)
    got = integ.trace(o[lo:i+40], d[lo:i+40], np.full(i+40-lo, np.inf, np.float32))
    ref, _ = orc.trace(sc, o[lo:i+40], d[lo:i+40], np.full(i+40-lo, np.inf, np.float32))
    print("window mismatches", np.nonzero(got["prim"] != ref["prim"])[0])
