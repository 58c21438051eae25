#!/usr/bin/env python3
"""Regenerates tests/golden/*.  The reference (Rust, nightly, un-vendored crates) cannot be built or imported
in this image, so these fixtures are produced by the CPU oracle itself after it was pinned against the
reference's unit-test vectors (tests/test_oracle_kat.py).  They guard the oracle against drift and give the
GPU box (which has no /root/reference) fixed inputs/outputs.
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
from shimmer_b200 import scenes  # noqa: E402


def main():
    out = {}
    for kind in ["diffuse", "conductor", "mirror", "glass", "roughglass", "coated", "coatedrough", "ortho", "thinglass"] + list(scenes.TEXTURED_KINDS) + list(scenes.INSTANCED_KINDS) + list(scenes.VARIETY_KINDS) + list(scenes.INSTANCED_SHAPE_KINDS):
        sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
        film, st, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
        out[kind] = dict(closest_hit_rays=int(st.closest_hit_rays), shadow_rays=int(st.shadow_rays),
                         film_sum=[float(x) for x in film.sum(axis=0)])
    json.dump(out, open(os.path.join(HERE, "tiny_films.json"), "w"), indent=1)
    # ray-cast golden: first-hit records of a fixed ray set on the 64x64 Cornell scene
    sc = scenes.cornell_box(resolution=(64, 64)).build()
    rng = np.random.default_rng(11)
    n = 4096
    o = (rng.random((n, 3)).astype(np.float32) - 0.5) * np.float32(500) + np.array([0, 0, 1080], np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    hits, st = orc.trace(sc, o, d, np.full(n, np.inf, np.float32))
    np.savez_compressed(os.path.join(HERE, "cornell_raycast.npz"), o=o, d=d.astype(np.float32), prim=hits["prim"], t=hits["t"],
                        b=np.stack([hits["b0"], hits["b1"], hits["b2"]], 1), ng=hits["ng"],
                        nodes=np.uint64(st.nodes_visited), tris=np.uint64(st.tris_tested))
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()
