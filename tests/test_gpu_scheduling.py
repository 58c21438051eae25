"""Scheduling knobs of the wavefront loop must not change the image: the stage barriers and CTA shapes of the shade kernels, the
number of wavefronts in flight, the wavefront width and the traversal kernels' vote-loop thresholds only reorder work.  Every
(pixel, sample) draws from its own random stream and the film is accumulated in f64, so films must agree to f64 round-off
(atomic adds commute up to the last bits of the f64 sums) and the per-pixel weights and ray counts exactly."""
import os

import numpy as np
import pytest

from shimmer_b200 import Options, create_integrator, scenes

pytestmark = pytest.mark.gpu


def _render(sc, cfg, spp, env, max_paths_in_flight=0):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp}, max_paths_in_flight=max_paths_in_flight)
        film = integ.render(Options(seed=3, pixel_samples=spp)).copy()
        st = integ.stats.as_dict()
        integ.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return film, st


KNOBS = [{"SG_SHADE_SYNC": "0"}, {"SG_SHADE_SYNC": "31"}, {"SG_SHADE_SYNC_TEX": "14"}, {"SG_OVERLAP": "1"}, {"SG_TRACE_DUAL": "3"},
         {"SG_TRACE_DUAL": "3", "SG_DUAL_LEVELS": "3", "SG_DUAL_LEVELS_SHADOW": "2", "SG_DUAL_REFILL": "5", "SG_DUAL_LEAF": "3"},
         {"SG_REFILL_THRESHOLD": "4", "SG_LEAF_THRESHOLD": "12", "SG_INTERIOR_BURST": "2"}]


@pytest.mark.parametrize("name", ["cornell", "glass", "instanced"])
def test_scheduling_knobs_do_not_change_the_film(name):
    cfg = scenes.CONFIGS[name]
    W, H = cfg["resolution"]
    c = 96
    x0 = (W - c) // 2; y0 = min(H - c, int(H * 0.55))
    sc = cfg["builder"](resolution=cfg["resolution"], crop=(x0, y0, x0 + c, y0 + c)).build()
    spp = 256                                               # 96 x 96 x 256 = 2.36 M paths: enough for two wavefronts in flight (>= 2 Mi)
    base, bst = _render(sc, cfg, spp, {})
    assert np.all(base[:, 3] == spp)
    for env in KNOBS:
        film, st = _render(sc, cfg, spp, env)
        assert np.array_equal(film[:, 3], base[:, 3]), env
        np.testing.assert_allclose(film[:, :3], base[:, :3], rtol=1e-12, atol=1e-300, err_msg=str(env))
        assert st["closest_hit_rays"] == bst["closest_hit_rays"] and st["shadow_rays"] == bst["shadow_rays"], env
    # a narrow wavefront (many small batches, two in flight) against the single wide one
    film, st = _render(sc, cfg, spp, {}, max_paths_in_flight=20000)
    np.testing.assert_allclose(film[:, :3], base[:, :3], rtol=1e-12, atol=1e-300)
    assert st["closest_hit_rays"] == bst["closest_hit_rays"] and st["shadow_rays"] == bst["shadow_rays"]
