"""Host-side logic: the native BVH builder reproduces the reference topology (checked against the oracle's
restatement of aggregate.rs:207-468), scene flattening invariants, tile decomposition."""
import ctypes as C

import numpy as np
import pytest

import orc
from shimmer_b200 import ffi, host, scenes


def _oracle_bvh(bounds):
    n = len(bounds)
    nodes = np.zeros(max(2 * n - 1, 1), dtype=np.dtype(ffi.SgBvhNode)); order = np.empty(n, np.uint32)
    k = orc.lib().orc_bvh_build(n, np.ascontiguousarray(bounds, np.float32).ctypes.data, nodes.ctypes.data, order.ctypes.data)
    return nodes[:k], order


def _host_bvh(bounds):
    n = len(bounds)
    nodes = np.zeros(max(2 * n - 1, 1), dtype=np.dtype(ffi.SgBvhNode)); order = np.empty(n, np.uint32)
    k = ffi.load_host_library().sh_bvh_build(n, np.ascontiguousarray(bounds, np.float32).ctypes.data, nodes.ctypes.data, order.ctypes.data)
    return nodes[:k], order


def _check_invariants(nodes, order, bounds):
    n = len(bounds)
    assert sorted(order.tolist()) == list(range(n))
    seen = 0
    stack = [0]
    while stack:
        i = stack.pop()
        nd = nodes[i]
        if nd["n_prims"] > 0:
            idx = order[nd["offset"]:nd["offset"] + nd["n_prims"]]
            assert np.all(bounds[idx, :3] >= nd["bmin"]) and np.all(bounds[idx, 3:] <= nd["bmax"])
            seen += int(nd["n_prims"])
        else:
            l, r = nodes[i + 1], nodes[nd["offset"]]
            assert nd["axis"] in (0, 1, 2) and nd["offset"] > i + 1
            assert np.array_equal(np.minimum(l["bmin"], r["bmin"]), nd["bmin"]) and np.array_equal(np.maximum(l["bmax"], r["bmax"]), nd["bmax"])
            stack += [i + 1, int(nd["offset"])]
    assert seen == n


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (37, 2), (1000, 3), (20000, 4)])
def test_host_bvh_equals_reference_restatement(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.random((n, 3)).astype(np.float32) * 10
    e = rng.random((n, 3)).astype(np.float32) * 0.3
    bounds = np.concatenate([c - e, c + e], axis=1).astype(np.float32)
    hn, ho = _host_bvh(bounds)
    on, oo = _oracle_bvh(bounds)
    assert len(hn) == len(on) == 2 * n - 1          # one primitive per leaf: maxnodeprims is ignored (aggregate.rs:41,60)
    assert hn.tobytes() == on.tobytes() and np.array_equal(ho, oo)
    _check_invariants(hn, ho, bounds)


def test_bvh_degenerate_inputs_make_multi_primitive_leaves():
    # coincident centroids -> "unusual edge case" leaf (aggregate.rs:345-355); zero-area bounds -> leaf (:326)
    b = np.array([[0, 0, 0, 1, 1, 1]] * 5 + [[2, 2, 2, 2, 2, 2]] * 3, np.float32)
    hn, ho = _host_bvh(b); on, oo = _oracle_bvh(b)
    assert hn.tobytes() == on.tobytes() and np.array_equal(ho, oo)
    leaves = hn[hn["n_prims"] > 0]
    assert sorted(leaves["n_prims"].tolist()) == [3, 5]
    _check_invariants(hn, ho, b)


def test_bvh_median_fallback_when_midpoint_partition_is_empty():
    # all centroids but one at the same coordinate on the split axis -> partition puts everything on one side
    # only if pmid rounds onto it; build a case: centroids {0, 1e-45...}: use large overlapping boxes
    c = np.zeros((8, 3), np.float32); c[:, 0] = [0, 0, 0, 0, 0, 0, 0, np.float32(1e-45)]
    b = np.concatenate([c - 1, c + 1], axis=1).astype(np.float32); b[:, 1:3] += np.arange(8)[:, None] * 0  # keep y,z equal
    hn, ho = _host_bvh(b); on, oo = _oracle_bvh(b)
    _check_invariants(hn, ho, b); _check_invariants(on, oo, b)
    assert len(hn) == len(on)


def test_cornell_flattening(cornell64):
    sc = cornell64
    assert sc.meta["n_triangles"] == 32 and sc.meta["n_lights"] == 2
    prims = sc.arrays["prims"]
    lit = prims[prims["light"] >= 0]
    assert len(lit) == 2 and set(lit["light"].tolist()) == {0, 1}
    # the light faces down (one-sided emission towards the room): normal = normalize((p0-p2) x (p1-p2)), triangle.rs:407
    m = sc.arrays["meshes"][7]
    P = sc.arrays["p"][m.first_vertex:m.first_vertex + m.n_vertices]; I = sc.arrays["idx"][m.first_index // 3:m.first_index // 3 + 2]
    for t in I:
        nrm = np.cross(P[t[0]] - P[t[2]], P[t[1]] - P[t[2]])
        assert nrm[1] < 0 and abs(nrm[0]) < 1e-3 and abs(nrm[2]) < 1e-3
    # camera-world rendering space: camera at the origin of render space (camera.rs:511-514)
    rfc = np.array(sc.desc.camera.render_from_camera[:]).reshape(4, 4)
    assert np.allclose(rfc[:3, 3], 0.0)
    # light scale = scale / spectrum_to_photometric(L) (light.rs:583)
    lt = sc.arrays["lights"][0]
    dense = sc.arrays["pool"][sc.arrays["spectra"][lt.spectrum].off_a:][:471]
    y = float(np.sum(dense.astype(np.float64) * host.cie("Y").astype(np.float64)))
    assert lt.scale * y == pytest.approx(20.0, rel=1e-4)
    assert lt.area == pytest.approx(0.5 * 130 * 105, rel=1e-5)


def test_mesh_scene_generator_is_deterministic_and_sized():
    a = scenes.mesh_scene(n_theta=40, n_phi=40, resolution=(32, 32)).build()
    b = scenes.mesh_scene(n_theta=40, n_phi=40, resolution=(32, 32)).build()
    assert a.meta["n_triangles"] == 2 * 40 * 39 + 4
    assert a.arrays["nodes"].tobytes() == b.arrays["nodes"].tobytes() and a.arrays["p"].tobytes() == b.arrays["p"].tobytes()
    # every primitive of the conductor mesh points at a conductor material
    kinds = np.array([a.arrays["materials"][int(m)].kind for m in a.arrays["prims"]["material"]])
    assert set(kinds.tolist()) == {ffi.SG_MATERIAL_DIFFUSE, ffi.SG_MATERIAL_CONDUCTOR}


def test_full_size_c2_triangle_count_formula():
    # BASELINE.json configs[1]: 708 x 708 lat-long grid ~ 1.0 M triangles
    assert 2 * 708 * 707 + 4 == 1001116


def test_orthographic_camera_rays_closed_form():
    """OrthographicCamera::generate_ray_differential (camera.rs:760-784): origin = camera_from_raster(p_film) on the z = 0
    plane of the screen window, direction +z, for every pixel; the oracle returns it in camera space as the reference does."""
    import orc
    from shimmer_b200 import scenes
    sc = scenes.tiny_scene("ortho", resolution=(16, 16)).build()
    assert sc.desc.camera.kind == 1
    xy = np.array([[x, y] for y in range(16) for x in range(16)], np.int32); si = np.zeros(len(xy), np.int32)
    rays, lam = orc.camera_rays(sc, orc.make_params(seed=3, spp=1, flags=1), xy, si)       # SG_OPT_DISABLE_PIXEL_JITTER: pixel centres
    assert np.array_equal(rays[:, 3:], np.tile(np.float32([0, 0, 1]), (256, 1)))
    want_x = -1.8 + (xy[:, 0] + 0.5) / 16 * 3.6; want_y = 1.8 - (xy[:, 1] + 0.5) / 16 * 3.6
    assert np.allclose(rays[:, 0], want_x, atol=1e-5) and np.allclose(rays[:, 1], want_y, atol=1e-5) and np.allclose(rays[:, 2], 0.0, atol=1e-6)
    film, st, _ = orc.render(sc, orc.make_params(seed=1, spp=16))
    assert np.isfinite(film).all() and film[:, :3].sum() > 0


def _find_interval(L, lam):
    """math.rs:322-333 with pred = L[i] <= lam (what PiecewiseLinearSpectrum::get runs, spectrum.rs:318-337)."""
    size = len(L)
    first, last = 1, size - 2
    while last > 0:
        half = last >> 1; middle = first + half
        if L[middle] <= lam:
            first = middle + 1; last -= half + 1
        else:
            last = half
    return min(max(first - 1, 0), size - 2)


def test_spectrum_interval_table_reproduces_find_interval():
    """sg_scene_create tabulates find_interval at every integer wavelength 360..830 for piecewise-linear spectra and the device walks
    forward from the entry of floor(lambda) (csrc/sg_host_tables.h, sg_shading.cuh spectrum_get): the walk must end on the interval
    the reference's binary search returns -- for knots on and off the integers, several knots inside one bin, duplicates, knots
    outside the visible range, and the named metal spectra the synthetic scenes use."""
    import ctypes as C
    from shimmer_b200 import ffi
    h = ffi.load_host_library()
    rng = np.random.default_rng(3)
    cases = [np.arange(360, 831, 20, dtype=np.float32), np.arange(360, 831, 10, dtype=np.float32),
             np.sort(rng.uniform(300, 900, 57)).astype(np.float32),
             np.sort(np.concatenate([rng.uniform(500, 503, 40), [360, 830]])).astype(np.float32),      # many knots in a few bins
             np.array([360, 400, 400, 400, 500.5, 500.5, 830], np.float32),                              # duplicates
             np.array([400, 700], np.float32), np.array([100, 200, 300], np.float32), np.array([900, 1000, 1100], np.float32)]
    lam = np.concatenate([np.arange(360, 831, dtype=np.float32), rng.uniform(360, 830, 4000).astype(np.float32),
                          np.nextafter(np.arange(361, 831, dtype=np.float32), np.float32(0))])          # just below every integer
    for L in cases:
        out = np.zeros(471, np.uint16)
        assert h.sh_spectrum_lut(L.ctypes.data, len(L), out.ctypes.data) == 1
        n = len(L)
        for x in lam:
            o = int(out[int(x) - 360])
            while o < n - 2 and L[o + 1] <= x:
                o += 1
            assert o == _find_interval(L, x), (L[:6], x)
    unsorted = np.array([500, 400, 600], np.float32)
    assert h.sh_spectrum_lut(unsorted.ctypes.data, 3, np.zeros(471, np.uint16).ctypes.data) == 0      # keeps the binary search
    one = np.array([500], np.float32)
    assert h.sh_spectrum_lut(one.ctypes.data, 1, np.zeros(471, np.uint16).ctypes.data) == 0
