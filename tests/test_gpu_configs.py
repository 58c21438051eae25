"""BASELINE.json's five configurations at their FULL resolution and sample counts: the CUDA path (through
the C ABI) against the oracle on crop windows of the real frame --
  * same-stream parity: identical per-(pixel, sample) random streams on both sides, 48x48 window, full spp;
  * converged-image parity (north-star level 3): the oracle in the REFERENCE's RNG mode (stream_mode=1: one sequential
    generator per worker thread, integrator.rs:250-263 -- completely different random numbers), RMSE and mean relative
    luminance error <= 1 %.
The window sits below the image centre (objects, floor and shadows in every scene), like tools/results_table.py's."""
import numpy as np
import pytest

import orc
from shimmer_b200 import Options, create_integrator, scenes

pytestmark = pytest.mark.gpu


def _window(name, c):
    W, H = scenes.CONFIGS[name]["resolution"]
    x0 = (W - c) // 2; y0 = min(H - c, int(H * 0.55))
    return (x0, y0, x0 + c, y0 + c)


def _lum(a):
    return 0.2126 * a[..., 0] + 0.7152 * a[..., 1] + 0.0722 * a[..., 2]


def _gpu_film(sc, cfg, spp, seed=0):
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp})
    film = integ.render(Options(seed=seed, pixel_samples=spp)).copy()
    img = integ.develop(film).reshape(-1, 3).astype(np.float64)
    st = integ.stats.as_dict()
    integ.close()
    return film, img, st


# (config, fraction of pixels that must agree within 2e-3, developed-image RMSE bar)
# C3: dielectric paths are ulp-chaotic in the reference algorithm itself -- scaling one component of every first-bounce direction
# by (1 + 2^-23) inside the ORACLE changes 31 of 2304 pixels of this very window and moves the image by RMSE 2.0e-3
# (tests/test_oracle_render.py::test_glass_config_is_sensitive_to_one_ulp): a reflect-or-refract choice `uc < R / (R + T)` or a
# total-internal-reflection test flips on a last-bit difference of sin / cos / atanh between CUDA's libm and glibc and the two
# paths then carry different energy.  The CUDA path stays at that level (observed RMSE 3e-3), far below the 1 % image bar.
SAME_STREAM = [("cornell", 0.999, 1e-3), ("glass", 0.97, 8e-3), ("instanced", 0.99, 3e-3), ("composite", 0.998, 1e-3)]


@pytest.mark.parametrize("name,frac,rmse_bar", SAME_STREAM, ids=[s[0] for s in SAME_STREAM])
def test_full_config_window_matches_the_oracle_on_the_same_streams(name, frac, rmse_bar):
    cfg = scenes.CONFIGS[name]; spp = cfg["spp"]
    win = _window(name, 48)
    sc = cfg["builder"](resolution=cfg["resolution"], crop=win).build()
    film, img_g, gst = _gpu_film(sc, cfg, spp)
    ref, rst, _ = orc.render(sc, orc.make_params(seed=0, spp=spp, max_depth=cfg["max_depth"]))
    assert np.array_equal(film[:, 3], ref[:, 3]) and np.all(ref[:, 3] == spp)            # weight sums are exact
    lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
    ok = np.abs(lg - lr) <= 2e-3 * np.maximum(lr, 0.05 * lr.mean())
    assert ok.mean() >= frac, f"{name}: only {ok.mean():.5f} of pixels within 2e-3"
    assert abs(lg.sum() - lr.sum()) / lr.sum() < 2e-3
    img_r = orc.develop(sc, ref).astype(np.float64)
    rmse = np.sqrt(np.mean((img_g - img_r) ** 2)) / np.mean(img_r)
    assert rmse <= rmse_bar, (name, rmse)
    assert abs(_lum(img_g).mean() - _lum(img_r).mean()) / _lum(img_r).mean() <= 1e-3
    # the same paths were traced: ray counts agree to a few flipped paths
    assert abs(int(gst["closest_hit_rays"]) - int(rst.closest_hit_rays)) <= 2e-3 * rst.closest_hit_rays
    assert abs(int(gst["shadow_rays"]) - int(rst.shadow_rays)) <= 2e-3 * max(rst.shadow_rays, 1)


def test_c3_converged_image_against_reference_rng_mode():
    """C3 (glass dispersion, wavelength termination: material.rs:603-635, sampled_wavelengths.rs:79-96) against the oracle in the
    reference's RNG mode.  The scene's noise is heavy-tailed (diffuse -> glass -> small light caustic paths found only by BSDF
    sampling): the ORACLE against ITSELF in its two RNG modes is still at 13 % per-pixel RMSE at 4096 spp and 4.4 % at 65536 spp,
    so a per-pixel 1 % bar is out of reach of any practical sample count -- for real shimmer too.  The converged image is
    therefore compared at 1/16 resolution: a 64x64 window at 8192 spp, box-filtered 16x16 -> 4x4 (2.1 M samples per compared
    value; oracle-vs-oracle calibration: 0.78 % at 4096 spp).  Bars: RMSE <= 1 %, mean relative luminance error <= 1 %."""
    cfg = scenes.CONFIGS["glass"]; spp = 8192; c = 64; B = 16
    sc = cfg["builder"](resolution=cfg["resolution"], crop=_window("glass", c)).build()
    film, img_g, _ = _gpu_film(sc, cfg, spp)
    ref, _, _ = orc.render(sc, orc.make_params(seed=0, spp=spp, max_depth=cfg["max_depth"]), stream_mode=1)
    img_r = orc.develop(sc, ref).astype(np.float64)
    assert np.all(film[:, 3] == spp)
    box = lambda x: x.reshape(c // B, B, c // B, B, 3).mean(axis=(1, 3))
    g, r = box(img_g), box(img_r)
    rmse = np.sqrt(np.mean((g - r) ** 2)) / np.mean(r)
    assert rmse <= 0.01, rmse
    assert abs(_lum(img_g).mean() - _lum(img_r).mean()) / _lum(img_r).mean() <= 0.01


def test_c4_converged_image_against_reference_rng_mode():
    """C4 (instanced + EWA / bilinear / bump image textures + 1024 emitters through the uniform light sampler and MIS): a 10x10
    crop at 262144 spp, per-pixel RMSE <= 1 % and mean relative luminance error <= 1 % against the oracle in the reference's RNG
    mode (oracle-vs-oracle calibration: 2.8 % at 16384 spp, i.e. 0.7 % here)."""
    cfg = scenes.CONFIGS["instanced"]; spp = 262144; c = 10
    x0, y0 = 955, 597
    sc = cfg["builder"](resolution=cfg["resolution"], crop=(x0, y0, x0 + c, y0 + c)).build()
    film, img_g, _ = _gpu_film(sc, cfg, spp)
    ref, _, _ = orc.render(sc, orc.make_params(seed=0, spp=spp, max_depth=cfg["max_depth"]), stream_mode=1)
    img_r = orc.develop(sc, ref).astype(np.float64)
    assert np.all(film[:, 3] == spp)
    rmse = np.sqrt(np.mean((img_g - img_r) ** 2)) / np.mean(img_r)
    assert rmse <= 0.01, rmse
    assert abs(_lum(img_g).mean() - _lum(img_r).mean()) / _lum(img_r).mean() <= 0.01


def test_c5_sample_range_split_adds_up_at_4k():
    """C5's multi-GPU decomposition at the real frame size: eight 1/8 sample ranges (what 8 GPUs render) accumulated into one
    3840x2160 film equal the one-call render of the same samples; every weight sum is exact.  (16 spp here: the property does not
    depend on the sample count.)"""
    import torch
    cfg = scenes.CONFIGS["composite"]; W, H = cfg["resolution"]; spp = 16
    sc = cfg["builder"](resolution=(W, H)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 1024})
    opts = Options(seed=0, pixel_samples=1024)
    full = torch.zeros((W * H, 4), dtype=torch.float64, device="cuda"); parts = torch.zeros_like(full)
    integ.render_device(opts, full.data_ptr(), sample_range=(0, spp))
    for r in range(8):
        integ.render_device(opts, parts.data_ptr(), sample_range=(r * spp // 8, (r + 1) * spp // 8))
    torch.cuda.synchronize()
    assert bool((full[:, 3] == spp).all()) and bool((parts[:, 3] == spp).all())
    assert torch.allclose(full, parts, rtol=1e-9, atol=1e-12)
    integ.close()
