#!/usr/bin/env python3
"""Per-scene-kind agreement of the CUDA path with the oracle on identical random streams (run on the GPU box): for every tiny
scene kind the film tests use, the fraction of pixels whose luminance agrees within 1e-5 / 1e-4 / 2e-3 relative.  The numbers
set the per-kind bars of tests/test_gpu_parity.py::test_tiny_scene_films (KIND_BARS)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    import orc
    from shimmer_b200 import Options, create_integrator, scenes
    kinds = (["diffuse", "conductor", "mirror", "glass", "roughglass", "coated", "coatedrough", "ortho", "thinglass"] + list(scenes.TEXTURED_KINDS) +
             list(scenes.INSTANCED_KINDS) + list(scenes.SPHERE_KINDS) + list(scenes.PATCH_KINDS) + list(scenes.INSTANCED_SHAPE_KINDS) + list(scenes.VARIETY_KINDS))
    out = {}
    for kind in kinds:
        row = {}
        for res, spp in ((16, 4), (32, 16)):
            sc = scenes.tiny_scene(kind, resolution=(res, res)).build()
            integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp, "seed": 5})
            film = integ.render(Options()).copy(); integ.close()
            ref, _, _ = orc.render(sc, orc.make_params(seed=5, spp=spp))
            lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
            rel = np.abs(lg - lr) / np.maximum(lr, 0.05 * max(lr.mean(), 1e-12))
            row["%dx%dx%d" % (res, res, spp)] = [float((rel <= t).mean()) for t in (1e-5, 1e-4, 2e-3)] + [int((rel > 2e-3).sum())]
        out[kind] = row
        print("%-16s" % kind, row, flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_film_agreement.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
