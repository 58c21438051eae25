"""Deterministic generators for the benchmark configurations of BASELINE.json (C1..C5) and a few
tiny parity scenes.  Every generator is closed-form (no files, no RNG state outside a fixed seed)
so the oracle, the CUDA path and -- on a machine with a Rust toolchain -- real shimmer render
the same geometry.  Materials use `"spectrum"`-typed parameters only (never "rgb"), so the
hot path never needs the rgb2spec tables that are missing from the reference checkout
(SURVEY.md 8c).
"""
import math

import numpy as np

from . import host
from .host import SceneBuilder, Transform, named_spectrum

f32 = np.float32


def _quad(a, b, c, d):
    """Two triangles (a,b,c), (a,c,d)."""
    return np.array([a, b, c, d], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def _pl(pairs):
    lam = np.array([p[0] for p in pairs], np.float32); v = np.array([p[1] for p in pairs], np.float32)
    return ("pl", lam, v)


# Smooth synthetic stand-ins for the measured Cornell-box spectra (tabulated every 20 nm so they
# become PiecewiseLinearSpectrum in shimmer: "spectrum reflectance" [ l0 v0 l1 v1 ... ]).
def _white():
    return _pl([(l, 0.72 + 0.04 * math.sin((l - 400) / 300.0 * math.pi)) for l in range(360, 831, 20)])


def _red():
    return _pl([(l, 0.045 + 0.58 / (1.0 + math.exp(-(l - 598.0) / 9.0))) for l in range(360, 831, 10)])


def _green():
    return _pl([(l, 0.06 + 0.42 * math.exp(-((l - 535.0) / 42.0) ** 2)) for l in range(360, 831, 10)])


def _light_spectrum():
    # the classic Cornell light samples, extended flat to the visible range ends
    return _pl([(360, 0.0), (400, 0.0), (500, 8.0), (600, 15.6), (700, 18.4), (830, 18.4)])


def cornell_geometry():
    """Classic 555-unit Cornell layout: 5 walls, short + tall block, ceiling light quad."""
    quads = {}
    quads["floor"] = _quad((552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2))
    quads["ceiling"] = _quad((556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0))
    quads["back"] = _quad((549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2))
    quads["right"] = _quad((0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2))
    quads["left"] = _quad((552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0))
    # light sits 0.5 below the ceiling so no primary ray sees a t-tie between the two planes
    quads["light"] = _quad((343, 548.3, 227), (343, 548.3, 332), (213, 548.3, 332), (213, 548.3, 227))
    short = [((130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)),
             ((290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)),
             ((130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)),
             ((82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)),
             ((240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225))]
    tall = [((423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)),
            ((423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)),
            ((472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)),
            ((314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)),
            ((265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247))]

    def box(faces):
        P, I = [], []
        for f in faces:
            p, i = _quad(*f); I.append(i + len(P) * 4); P.append(p)
        return np.concatenate(P), np.concatenate(I)
    quads["short"] = box(short)
    quads["tall"] = box(tall)
    return quads


def cornell_box(resolution=(512, 512), crop=None, light_scale=20.0, add_to=None):
    """C1: synthetic Cornell box, diffuse + one two-triangle area light, fov 39.3 (BASELINE.json configs[0])."""
    b = add_to or SceneBuilder()
    if add_to is None:
        b.set_camera(pos=(278, 273, -800), look=(278, 273, 0), up=(0, 1, 0),
                     fov=2.0 * math.degrees(math.atan(0.0125 / 0.035)), resolution=resolution, crop=crop)
    white, red, green = b.diffuse(_white()), b.diffuse(_red()), b.diffuse(_green())
    g = cornell_geometry()
    for name, mat in (("floor", white), ("ceiling", white), ("back", white), ("right", green), ("left", red),
                      ("short", white), ("tall", white)):
        b.add_mesh(g[name][0], g[name][1], mat)
    b.add_mesh(g["light"][0], g["light"][1], white,
               area_light=dict(L=_light_spectrum(), scale=light_scale, two_sided=False))
    return b


def displaced_sphere(n_theta, n_phi, center, radius, amp=0.06, freq=7.0):
    """Lat-long tessellated sphere with a smooth radial displacement, vertices from closed-form
    trig in f64 (written as f32).  2*n_phi*(n_theta-1) triangles.  Poles are single vertices."""
    th = np.linspace(0.0, math.pi, n_theta + 1)[1:-1]            # interior rings
    ph = np.linspace(0.0, 2.0 * math.pi, n_phi, endpoint=False)
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    r = radius * (1.0 + amp * np.sin(freq * T + 0.3) * np.cos(freq * Pp * 0.5 + 1.1) + 0.5 * amp * np.sin(3.0 * freq * Pp + T))
    x = r * np.sin(T) * np.cos(Pp); y = r * np.cos(T); z = r * np.sin(T) * np.sin(Pp)
    ring = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    top = np.array([[0.0, radius, 0.0]]); bot = np.array([[0.0, -radius, 0.0]])
    P = np.concatenate([top, ring, bot]) + np.asarray(center, dtype=np.float64)
    nr = n_theta - 1
    idx = []
    j = np.arange(n_phi); jn = (j + 1) % n_phi
    idx.append(np.stack([np.zeros_like(j), 1 + jn, 1 + j], axis=1))                     # top cap
    for i in range(nr - 1):
        a = 1 + i * n_phi + j; bq = 1 + i * n_phi + jn; c = 1 + (i + 1) * n_phi + j; d = 1 + (i + 1) * n_phi + jn
        idx.append(np.stack([a, bq, d], axis=1)); idx.append(np.stack([a, d, c], axis=1))
    last = 1 + (nr - 1) * n_phi
    idx.append(np.stack([np.full_like(j, len(P) - 1), last + j, last + jn], axis=1))   # bottom cap
    I = np.concatenate(idx).astype(np.uint32)
    return P.astype(np.float32), I


def mesh_scene(n_theta=708, n_phi=708, resolution=(1024, 1024), crop=None, roughness=0.1):
    """C2: ~1 M-triangle displaced sphere split half diffuse / half conductor (Cu eta/k, roughness
    0.1) over a ground quad, one area-light quad (BASELINE.json configs[1])."""
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.6, -4.2), look=(0.0, 0.9, 0.0), up=(0, 1, 0), fov=38.0, resolution=resolution, crop=crop)
    white = b.diffuse(_white()); clay = b.diffuse(_red())
    copper = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=roughness)
    P, I = displaced_sphere(n_theta, n_phi, center=(0.0, 1.0, 0.0), radius=1.0)
    # split by triangle centroid x: left half diffuse, right half conductor (two meshes share vertices by copy)
    cx = P[I].mean(axis=1)[:, 0]
    for mask, mat in ((cx < 0.0, clay), (cx >= 0.0, copper)):
        sub = I[mask]
        used, inv = np.unique(sub.ravel(), return_inverse=True)
        b.add_mesh(P[used], inv.reshape(-1, 3).astype(np.uint32), mat)
    gp, gi = _quad((-6, -0.06, -6), (-6, -0.06, 6), (6, -0.06, 6), (6, -0.06, -6))
    b.add_mesh(gp, gi, white)
    lp, li = _quad((-1.5, 4.0, -1.5), (1.5, 4.0, -1.5), (1.5, 4.0, 1.5), (-1.5, 4.0, 1.5))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=12.0, two_sided=False))
    return b


def glass_scene(n_theta=128, n_phi=256, resolution=(1024, 1024), crop=None):
    """C3: glass prism + tessellated glass sphere with a NON-constant tabulated BK7 eta (-> wavelength
    termination, material.rs:609-620), small bright area light (BASELINE.json configs[2])."""
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.4, -4.5), look=(0.0, 0.7, 0.0), up=(0, 1, 0), fov=36.0, resolution=resolution, crop=crop)
    white = b.diffuse(_white())
    glass = b.dielectric(named_spectrum("glass-BK7"))
    P, I = displaced_sphere(n_theta, n_phi, center=(0.9, 0.75, 0.2), radius=0.75, amp=0.0)
    b.add_mesh(P, I, glass)
    # triangular prism along z
    a, bb, c = (-1.6, 0.0, -0.6), (-0.4, 0.0, -0.6), (-1.0, 1.04, -0.6)
    a2, b2, c2 = (-1.6, 0.0, 0.6), (-0.4, 0.0, 0.6), (-1.0, 1.04, 0.6)
    PP = np.array([a, bb, c, a2, b2, c2], np.float32)
    II = np.array([[0, 2, 1], [3, 4, 5], [0, 1, 4], [0, 4, 3], [1, 2, 5], [1, 5, 4], [2, 0, 3], [2, 3, 5]], np.uint32)
    b.add_mesh(PP, II, glass, object_from_world=Transform.translate((0.0, 0.002, 0.0)))
    gp, gi = _quad((-6, 0.0, -6), (-6, 0.0, 6), (6, 0.0, 6), (6, 0.0, -6))
    b.add_mesh(gp, gi, white)
    wp, wi = _quad((-6, 0.0, 3), (-6, 6, 3), (6, 6, 3), (6, 0.0, 3))
    b.add_mesh(wp, wi, white)
    lp, li = _quad((-2.6, 3.0, -1.0), (-2.2, 3.0, -1.0), (-2.2, 3.0, -0.6), (-2.6, 3.0, -0.6))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=900.0, two_sided=False))
    return b


def composite_scene(n_theta=708, n_phi=708, resolution=(3840, 2160), crop=None):
    """C5: the Cornell room enclosing the C2 mesh (BASELINE.json configs[4])."""
    b = SceneBuilder()
    b.set_camera(pos=(278, 273, -800), look=(278, 273, 0), up=(0, 1, 0),
                 fov=2.0 * math.degrees(math.atan(0.0125 / 0.035)) * 1.35, resolution=resolution, crop=crop)
    cornell_box(add_to=b)
    copper = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.1)
    P, I = displaced_sphere(n_theta, n_phi, center=(0.0, 0.0, 0.0), radius=1.0)
    xf = Transform.translate((300.0, 420.0, 150.0)) * Transform.scale(70.0, 70.0, 70.0)
    b.add_mesh(P, I, copper, object_from_world=xf)
    return b


def procedural_image(n=64, channels=3, quantize8=False):
    """Deterministic linear-valued test image (row 0 = top): checker x gradients, closed form.  quantize8: texels rounded to
    multiples of 1/255, i.e. exactly what an 8-bit PNG with a linear colour encoding decodes to (shimmer reads only PNG,
    image.rs:1140-1149) -- the BASELINE C4 scene uses it so that its PBRT export (pbrt_export.py) is texel-exact."""
    if quantize8:
        a = procedural_image(n, channels)
        return (np.rint(a.astype(np.float64) * 255.0).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    y, x = np.mgrid[0:n, 0:n].astype(np.float64) / n
    checker = (np.floor(x * 8.0) + np.floor(y * 8.0)) % 2.0
    r = 0.12 + 0.76 * checker
    g = 0.18 + 0.64 * x
    bl = 0.22 + 0.56 * y * (1.0 - 0.5 * checker)
    if channels == 1:
        return (0.2 + 0.6 * (0.5 + 0.5 * np.sin(9.0 * x + 2.0) * np.cos(7.0 * y))).astype(np.float32)
    return np.stack([r, g, bl], axis=-1).astype(np.float32)


def uv_sphere(n_theta, n_phi, center, radius):
    """Sphere with vertex normals and lat-long uv (duplicated seam column so uv stays continuous)."""
    th = np.linspace(0.0, math.pi, n_theta + 1)
    ph = np.linspace(0.0, 2.0 * math.pi, n_phi + 1)
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    N = np.stack([np.sin(T) * np.cos(Pp), np.cos(T), np.sin(T) * np.sin(Pp)], axis=-1).reshape(-1, 3)
    P = N * radius + np.asarray(center, dtype=np.float64)
    UV = np.stack([Pp / (2.0 * math.pi), 1.0 - T / math.pi], axis=-1).reshape(-1, 2)
    idx = []
    for i in range(n_theta):
        for j in range(n_phi):
            a = i * (n_phi + 1) + j; b = a + 1; c = a + n_phi + 1; d = c + 1
            if i > 0: idx.append((a, b, d))
            if i < n_theta - 1: idx.append((a, d, c))
    return P.astype(np.float32), np.asarray(idx, np.uint32), N.astype(np.float32), UV.astype(np.float32)


def uv_sphere_fast(n_theta, n_phi, center, radius, amp=0.0, freq=6.0):
    """uv_sphere() with vectorised index generation and an optional closed-form radial displacement (large meshes)."""
    th = np.linspace(0.0, math.pi, n_theta + 1)
    ph = np.linspace(0.0, 2.0 * math.pi, n_phi + 1)
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    N = np.stack([np.sin(T) * np.cos(Pp), np.cos(T), np.sin(T) * np.sin(Pp)], axis=-1).reshape(-1, 3)
    r = radius * (1.0 + amp * np.sin(freq * T) * np.cos(freq * Pp)).reshape(-1, 1)
    P = N * r + np.asarray(center, dtype=np.float64)
    UV = np.stack([Pp / (2.0 * math.pi), 1.0 - T / math.pi], axis=-1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n_theta), np.arange(n_phi), indexing="ij")
    a = (i * (n_phi + 1) + j).ravel(); b = a + 1; c = a + n_phi + 1; d = c + 1
    up = np.stack([a, b, d], axis=1)[(i > 0).ravel()]
    lo = np.stack([a, d, c], axis=1)[(i < n_theta - 1).ravel()]
    return P.astype(np.float32), np.concatenate([up, lo]).astype(np.uint32), N.astype(np.float32), UV.astype(np.float32)


def instanced_scene(n_theta=224, n_phi=224, grid=10, resolution=(1920, 1080), crop=None, lights=(16, 32), fix_instancing=True,
                    tex_size=512):
    """C4: one ~100 k-triangle uv-mapped object definition placed grid x grid (= 100) times with rotations and
    non-uniform scales (10 M instanced triangles), image textures (EWA on the instances, bilinear on the ground, a
    one-channel bump map on every other row) and 2 x lights[0] x lights[1] (= 1024) emissive triangles sampled through
    the uniform light sampler + MIS (BASELINE.json configs[3]).  Instance semantics: SG_SCENE_FIX_INSTANCING by
    default -- the reference's literal TransformedPrimitive (forward transform on shadow rays, inverse on normals,
    primitive.rs:172-175, transform.rs:573-597) is available with fix_instancing=False and is parity-tested on the
    tiny instanced scenes."""
    b = SceneBuilder()
    b.fix_instancing = fix_instancing
    half = 2.0 * grid
    b.set_camera(pos=(0.0, 0.55 * half, -1.25 * half), look=(0.0, 0.0, -0.1 * half), up=(0, 1, 0), fov=42.0, resolution=resolution, crop=crop)
    rgb_img, mono_img = procedural_image(tex_size, 3, quantize8=True), procedural_image(tex_size // 2, 1, quantize8=True)
    ground = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(2 * tex_size, 3, quantize8=True), filter="bilinear", su=8.0, sv=8.0))
    skin = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="ewa", su=4.0, sv=2.0, max_anisotropy=8.0))
    bumpy = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="trilinear", su=2.0, sv=2.0),
                      displacement_tex=b.image_texture(mono_img, filter="bilinear", su=12.0, sv=6.0, scale=0.02))
    white = b.diffuse(_white())
    P, I, Nn, UV = uv_sphere_fast(n_theta, n_phi, center=(0.0, 0.0, 0.0), radius=1.0, amp=0.05)
    objs = []
    for mat in (skin, bumpy):
        o = b.begin_object()
        b.add_mesh(P, I, mat, n=Nn, uv=UV, object=o)
        objs.append(o)
    for gy in range(grid):
        for gx in range(grid):
            k = gy * grid + gx
            x = (gx - 0.5 * (grid - 1)) * 4.0; z = (gy - 0.5 * (grid - 1)) * 4.0
            sx_, sy_, sz_ = 1.0 + 0.25 * math.sin(1.7 * k), 1.1 + 0.3 * math.cos(0.9 * k), 1.0 + 0.2 * math.sin(2.3 * k + 1.0)
            xf = (Transform.translate((x, 1.25 * sy_, z)) * Transform.rotate(37.0 * k, (0.3 * math.sin(k), 1.0, 0.2 * math.cos(k))) *
                  Transform.scale(sx_, sy_, sz_))
            b.add_instance(objs[gy & 1], xf)
    gp, gi = _quad((-half - 4, 0.0, -half - 4), (-half - 4, 0.0, half + 4), (half + 4, 0.0, half + 4), (half + 4, 0.0, -half - 4))
    b.add_mesh(gp, gi, ground, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    # many emitters: a ceiling of small downward-facing quads, one DiffuseAreaLight per triangle (scene.rs:609-622)
    ny, nx = lights
    LP, LI = [], []
    for iy in range(ny):
        for ix in range(nx):
            cx = (ix + 0.5) / nx * 2.0 * half - half; cz = (iy + 0.5) / ny * 2.0 * half - half; h = 0.22; y = 9.0
            p, idx = _quad((cx - h, y, cz - h), (cx + h, y, cz - h), (cx + h, y, cz + h), (cx - h, y, cz + h))
            LI.append(idx + 4 * len(LP)); LP.append(p)
    b.add_mesh(np.concatenate(LP), np.concatenate(LI), white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=160.0, two_sided=False))
    return b


SPHERE_KINDS = ("spheres", "spherestex", "spherelight")


def sphere_tiny_scene(kind="spheres", resolution=(32, 32), fix=False):
    """Shape "sphere" (shape/sphere.rs): a diffuse sphere, a glass sphere (translated + uniformly scaled), a copper partial
    sphere (zmin/zmax/phimax clipped, rotated + non-uniformly scaled: every quirk of Transform::apply(SurfaceInteraction)
    shows) over a ground quad; `spherestex` puts image textures (uv from phi/theta, dndu/dndv from the fundamental forms) on
    the spheres."""
    b = SceneBuilder()
    b.fix_instancing = fix
    b.set_camera(pos=(0.0, 1.5, -4.5), look=(0.0, 0.6, 0.0), up=(0, 1, 0), fov=42.0, resolution=resolution)
    white = b.diffuse(_white())
    if kind == "spherelight":
        # emissive spheres (sphere.rs:339-456): a small bright one seen from outside (cone sampling, the 2.90 pdf), a partial
        # one, and a large dim two-sided shell around the whole scene (reference point INSIDE: area sampling + pdf by intersection)
        b.add_sphere(0.6, b.diffuse(_green()), object_from_world=Transform.translate((-1.0, 0.6, 0.3)))
        b.add_sphere(0.5, b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.15),
                     object_from_world=Transform.translate((0.9, 0.5, -0.3)))
        b.add_sphere(0.25, white, object_from_world=Transform.translate((0.2, 2.2, 0.4)),
                     area_light=dict(L=named_spectrum("stdillum-D65"), scale=60.0, two_sided=False))
        b.add_sphere(0.2, white, z_min=-0.1, z_max=0.2, phi_max=270.0, object_from_world=Transform.translate((-1.6, 1.4, -0.8)) * Transform.rotate(40.0, (0, 1, 0)),
                     area_light=dict(L=named_spectrum("stdillum-D65"), scale=25.0, two_sided=True))
        b.add_sphere(9.0, white, object_from_world=Transform.translate((0.0, 1.0, 0.0)),
                     area_light=dict(L=named_spectrum("stdillum-D65"), scale=0.15, two_sided=True))
        gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, white)
        return b
    if kind == "spherestex":
        m0 = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="trilinear", su=4.0, sv=2.0))
        m2 = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="ewa", su=3.0, sv=3.0),
                       displacement_tex=b.image_texture(procedural_image(32, 1), filter="bilinear", su=8.0, sv=4.0, scale=0.03))
    else:
        m0 = b.diffuse(_green())
        m2 = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.2)
    glass = b.dielectric(("const", 1.5))
    b.add_sphere(0.6, m0, object_from_world=Transform.translate((-1.3, 0.6, 0.4)) * Transform.rotate(-90.0, (1, 0, 0)))
    b.add_sphere(1.0, glass, object_from_world=Transform.translate((0.1, 0.55, -0.6)) * Transform.scale(0.55, 0.55, 0.55))
    b.add_sphere(0.7, m2, z_min=-0.45, z_max=0.6, phi_max=300.0,
                 object_from_world=Transform.translate((1.4, 0.75, 0.5)) * Transform.rotate(-70.0, (1, 0.2, 0)) * Transform.scale(1.0, 0.8, 1.15))
    gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, white, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    lp, li = _quad((-0.6, 2.8, -0.6), (0.6, 2.8, -0.6), (0.6, 2.8, 0.6), (-0.6, 2.8, 0.6))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b


PATCH_KINDS = ("patches", "patchestex", "patchlight", "patchlightbent")


def patch_grid(n, size=2.0, amp=0.25):
    """(n+1)^2 vertices of a wavy height field with analytic normals and uv, n^2 bilinear patches (p00, p10, p01, p11)."""
    g = np.linspace(0.0, 1.0, n + 1)
    U, V = np.meshgrid(g, g, indexing="ij")
    X = (U - 0.5) * size; Z = (V - 0.5) * size
    Y = amp * np.sin(2.5 * X) * np.cos(2.0 * Z)
    dYdx = amp * 2.5 * np.cos(2.5 * X) * np.cos(2.0 * Z); dYdz = -amp * 2.0 * np.sin(2.5 * X) * np.sin(2.0 * Z)
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    N = np.stack([-dYdx, np.ones_like(X), -dYdz], axis=-1).reshape(-1, 3); N /= np.linalg.norm(N, axis=1, keepdims=True)
    UV = np.stack([U, V], axis=-1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (i * (n + 1) + j).ravel()
    # orientation chosen so that (p10 - p00) x (p01 - p00) points up (+y): u along z, v along x
    idx = np.stack([a, a + 1, a + n + 1, a + n + 2], axis=1)
    return P.astype(np.float32), idx.astype(np.uint32), N.astype(np.float32), UV.astype(np.float32)


def patch_tiny_scene(kind="patches", resolution=(32, 32)):
    """Shape "bilinearmesh" (shape/bilinear_patch.rs): a wavy 6x6 patch grid with vertex normals and uv (every patch is
    non-planar), a twisted single patch without normals (copper), a planar quad patch (the PLY-quad case) behind them."""
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 2.0, -4.0), look=(0.0, 0.4, 0.0), up=(0, 1, 0), fov=42.0, resolution=resolution)
    white = b.diffuse(_white())
    if kind == "patchestex":
        skin = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="ewa", su=3.0, sv=3.0),
                         displacement_tex=b.image_texture(procedural_image(32, 1), filter="bilinear", su=6.0, sv=6.0, scale=0.03))
        wall = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="trilinear", su=2.0, sv=2.0))
    else:
        skin = b.diffuse(_green()); wall = b.diffuse(_red())
    copper = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.2)
    P, I, N, UV = patch_grid(6)
    b.add_bilinear_mesh(P, I, skin, n=N, uv=UV, object_from_world=Transform.translate((0.0, 0.35, 0.0)))
    tw = np.array([[-0.5, 0.0, -0.3], [0.5, 0.0, -0.5], [-0.4, 1.0, 0.3], [0.6, 0.9, -0.2]], np.float32)
    b.add_bilinear_mesh(tw, [[0, 1, 2, 3]], copper, object_from_world=Transform.translate((1.3, 0.6, 0.6)) * Transform.rotate(25.0, (0, 1, 0)))
    quad = np.array([[-2.5, 0.0, 1.8], [2.5, 0.0, 1.8], [-2.5, 2.5, 1.8], [2.5, 2.5, 1.8]], np.float32)
    b.add_bilinear_mesh(quad, [[0, 1, 2, 3]], wall, uv=np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32), reverse_orientation=True)
    gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, white)
    if kind == "patchlight":        # a rectangular emissive patch: sampled by solid angle (sample_spherical_rectangle, bilinear_patch.rs:666-736)
        lq = np.array([[-0.6, 3.0, -0.9], [0.6, 3.0, -0.9], [-0.6, 3.0, 0.3], [0.6, 3.0, 0.3]], np.float32)
        b.add_bilinear_mesh(lq, [[0, 1, 2, 3]], white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
        return b
    if kind == "patchlightbent":    # a twisted two-sided emitter with uv + normals (area sampling with the bilinear warp) and a tiny
        lq = np.array([[-0.7, 2.8, -0.9], [0.6, 3.1, -0.8], [-0.6, 3.0, 0.4], [0.7, 2.7, 0.2]], np.float32)     # far rectangle (solid angle <= 1e-4)
        ln = np.tile(np.array([[0.0, -1.0, 0.0]], np.float32), (4, 1))
        b.add_bilinear_mesh(lq, [[0, 1, 2, 3]], white, n=ln, uv=np.array([[0.1, 0.0], [0.9, 0.1], [0.0, 1.0], [1.0, 0.8]], np.float32),
                            area_light=dict(L=named_spectrum("stdillum-D65"), scale=25.0, two_sided=True))
        fq = np.array([[-0.02, 14.0, -0.02], [0.02, 14.0, -0.02], [-0.02, 14.0, 0.02], [0.02, 14.0, 0.02]], np.float32)
        b.add_bilinear_mesh(fq, [[0, 1, 2, 3]], white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=4000.0, two_sided=False))
        return b
    lp, li = _quad((-0.6, 3.0, -0.9), (0.6, 3.0, -0.9), (0.6, 3.0, 0.3), (-0.6, 3.0, 0.3))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b


TEXTURED_KINDS = ("tex", "texewa", "texbump", "texcoated")
INSTANCED_KINDS = ("inst", "instrot", "instfix", "insttex")
INSTANCED_SHAPE_KINDS = ("instshapes", "instshapesfix", "instshapestex")     # spheres + bilinear patches inside object definitions


def instanced_tiny_scene(kind="inst", resolution=(32, 32), flatten=False):
    """Object instancing (primitive.rs:136-176): one object definition (a small sphere + a single-triangle object) placed
    several times.  `inst`: translations only (where the reference's closest-hit path is right and only its shadow-ray
    quirk shows); `instrot`: rotation + non-uniform scale (every quirk of transform.rs:573-609 shows); `instfix`: the
    same with SG_SCENE_FIX_INSTANCING; `insttex`: textured, normal-interpolated instances.  flatten=True bakes the
    instances into ordinary meshes (what `instfix` must reproduce)."""
    b = SceneBuilder()
    b.fix_instancing = kind == "instfix"
    b.set_camera(pos=(0.0, 1.4, -4.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=resolution)
    white = b.diffuse(_white())
    if kind == "insttex":
        mat = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="trilinear", su=2.0, sv=2.0))
        copper = b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0)
    else:
        mat = b.diffuse(_green())
        copper = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.15)
    P, I, Nn, UV = uv_sphere(8, 12, center=(0.0, 0.0, 0.0), radius=0.4)
    TP = np.array([[-0.3, -0.2, 0.0], [0.3, -0.2, 0.0], [0.0, 0.45, 0.1]], np.float32); TI = np.array([[0, 1, 2]], np.uint32)
    if kind in ("instrot", "instfix"):
        xfs = [Transform.translate((-1.2, 0.5, 0.3)) * Transform.rotate(35.0, (0, 1, 0.3)) * Transform.scale(1.0, 1.4, 0.8),
               Transform.translate((0.0, 0.6, 0.0)) * Transform.scale(1.3, 1.3, 1.3),
               Transform.translate((1.2, 0.45, -0.2)) * Transform.rotate(-50.0, (1, 0.2, 0))]
        txf = [Transform.translate((0.6, 1.3, 0.4)) * Transform.rotate(20.0, (0, 0, 1))]
    else:
        xfs = [Transform.translate((-1.2, 0.5, 0.3)), Transform.translate((0.0, 0.6, 0.0)), Transform.translate((1.2, 0.45, -0.2)),
               Transform.translate((0.5, 1.5, 0.8))]
        txf = [Transform.translate((-0.6, 1.3, 0.4)), Transform.translate((0.7, 1.2, -0.5))]
    use_n = kind == "insttex"
    if flatten:
        for xf in xfs:
            b.add_mesh(P, I, mat, n=Nn if use_n else None, uv=UV, object_from_world=xf)
        for xf in txf:
            b.add_mesh(TP, TI, copper, object_from_world=xf)
    else:
        sphere_obj = b.begin_object()
        b.add_mesh(P, I, mat, n=Nn if use_n else None, uv=UV, object=sphere_obj)
        tri_obj = b.begin_object()                                    # a one-primitive definition: no aggregate (scene.rs:821-832)
        b.add_mesh(TP, TI, copper, object=tri_obj)
        for xf in xfs:
            b.add_instance(sphere_obj, xf)
        for xf in txf:
            b.add_instance(tri_obj, xf)
    gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, white, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    lp, li = _quad((-0.6, 2.8, -0.6), (0.6, 2.8, -0.6), (0.6, 2.8, 0.6), (-0.6, 2.8, 0.6))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b


def instanced_shapes_tiny_scene(kind="instshapes", resolution=(32, 32), flatten=False):
    """Object definitions that hold shapes other than triangles (scene.rs:814-866 accepts any Shape): one definition with a partial
    sphere (own transform), a twisted bilinear patch and a triangle mesh behind its own BvhAggregate, and a one-primitive definition
    that is a bare sphere.  `instshapes`: literal TransformedPrimitive semantics with rotated / scaled instances; `instshapesfix`:
    SG_SCENE_FIX_INSTANCING; `instshapestex`: textured (uv of spheres / patches through the instance transform)."""
    b = SceneBuilder()
    b.fix_instancing = kind == "instshapesfix"
    b.set_camera(pos=(0.0, 1.4, -4.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=resolution)
    white = b.diffuse(_white())
    if kind == "instshapestex":
        mat = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="ewa", su=2.0, sv=2.0))
        metal = b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0)
    else:
        mat = b.diffuse(_green())
        metal = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.15)
    glass = b.dielectric(("const", 1.5))
    sph_xf = Transform.translate((0.0, 0.1, 0.0)) * Transform.rotate(-70.0, (1, 0, 0))
    patch_xf = Transform.translate((0.0, 0.0, 0.5))
    tw = np.array([[-0.5, -0.3, -0.2], [0.5, -0.3, -0.4], [-0.4, 0.5, 0.3], [0.6, 0.4, -0.1]], np.float32)
    tuv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)
    TP = np.array([[-0.4, -0.35, -0.3], [0.4, -0.35, -0.3], [0.0, -0.35, 0.4], [0.0, 0.2, 0.0]], np.float32)
    TI = np.array([[0, 1, 3], [1, 2, 3], [2, 0, 3]], np.uint32)
    xfs = [Transform.translate((-1.2, 0.6, 0.3)) * Transform.rotate(35.0, (0, 1, 0.3)) * Transform.scale(1.0, 1.4, 0.8),
           Transform.translate((0.0, 0.7, 0.0)) * Transform.scale(1.3, 1.3, 1.3),
           Transform.translate((1.2, 0.55, -0.2)) * Transform.rotate(-50.0, (1, 0.2, 0))]
    ball_xfs = [Transform.translate((0.6, 1.6, 0.4)) * Transform.scale(1.0, 0.6, 1.0), Transform.translate((-0.7, 1.5, -0.4))]
    if flatten:                                                   # the same shapes at the top level with composed transforms
        for xf in xfs:
            b.add_mesh(TP, TI, glass, object_from_world=xf)
        for xf in xfs:
            b.add_bilinear_mesh(tw, [[0, 1, 2, 3]], metal, uv=tuv, object_from_world=xf * patch_xf)
        for xf in xfs:
            b.add_sphere(0.35, mat, z_min=-0.2, z_max=0.3, phi_max=300.0, object_from_world=xf * sph_xf)
        for xf in ball_xfs:
            b.add_sphere(0.3, metal, object_from_world=xf)
    else:
        group = b.begin_object()
        b.add_sphere(0.35, mat, z_min=-0.2, z_max=0.3, phi_max=300.0, object=group, object_from_world=sph_xf)
        b.add_bilinear_mesh(tw, [[0, 1, 2, 3]], metal, uv=tuv, object=group, object_from_world=patch_xf)
        b.add_mesh(TP, TI, glass, object=group)
        ball = b.begin_object()                                   # a one-primitive definition: a bare sphere, no aggregate
        b.add_sphere(0.3, metal, object=ball)
        for xf in xfs:
            b.add_instance(group, xf)
        for xf in ball_xfs:
            b.add_instance(ball, xf)
    b.add_sphere(0.25, glass, object_from_world=Transform.translate((0.0, 0.25, -1.2)))          # a top-level sphere next to the instances
    gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, white, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    lp, li = _quad((-0.6, 2.8, -0.6), (0.6, 2.8, -0.6), (0.6, 2.8, 0.6), (-0.6, 2.8, 0.6))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b


VARIETY_KINDS = ("envmap", "envonly", "envrot", "coatedcond", "coatedcondrough", "coatedcondrefl", "normalmap", "texsph", "texcyl", "texplanar",
                 "mix", "mixtex", "mixnested", "textree", "textreemix", "texparams", "texparams2")


def procedural_envmap(n=32):
    """A square equal-area octahedral environment map (linear RGB): dim blue-grey sky, one bright warm blob and a dimmer
    cool one, so that the compensated distribution (light.rs:941-948) is sparse."""
    y, x = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    img = np.empty((n, n, 3), np.float32)
    base = 0.15 + 0.1 * np.sin(5.0 * x) * np.cos(3.0 * y)
    img[:, :, 0] = base; img[:, :, 1] = base * 1.1; img[:, :, 2] = base * 1.4
    b1 = np.exp(-((x - 0.62) ** 2 + (y - 0.40) ** 2) / 0.004); b2 = np.exp(-((x - 0.30) ** 2 + (y - 0.55) ** 2) / 0.01)
    img[:, :, 0] += 14.0 * b1 + 1.0 * b2; img[:, :, 1] += 11.0 * b1 + 2.0 * b2; img[:, :, 2] += 6.0 * b1 + 4.0 * b2
    return img.astype(np.float32)


def procedural_normal_map(n=32):
    """Tangent-space normal map (RGB = 0.5 + 0.5 n) of a bumpy surface."""
    y, x = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    nx = 0.35 * np.sin(2 * np.pi * 3 * x); ny = 0.35 * np.cos(2 * np.pi * 2 * y)
    nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.05))
    return (0.5 + 0.5 * np.stack([nx, ny, nz], axis=2)).astype(np.float32)


def variety_tiny_scene(kind, resolution=(32, 32)):
    """SURVEY 8f next-4 scenes: image-infinite lights, CoatedConductor, normal maps, non-UV texture mappings, Mix."""
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=resolution)
    white = b.diffuse(_white())
    guv = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32)
    ground = white
    area_light = True
    if kind in ("envmap", "envonly", "envrot"):
        xf = Transform.rotate(35.0, (0.3, 1.0, 0.2)) if kind == "envrot" else None
        b.add_image_infinite_light(procedural_envmap(32), scale=1.0 if kind != "envonly" else 2.0, light_from_world=xf,
                                   illuminance=3.0 if kind == "envrot" else None)
        mat = b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0) if kind == "envmap" else b.diffuse(_green())
        if kind == "envrot":
            mat = b.dielectric(("const", 1.5), roughness=0.1)
        area_light = kind != "envonly"
    elif kind == "coatedcond":
        mat = b.coated_conductor()
    elif kind == "coatedcondrough":
        mat = b.coated_conductor(conductor_eta=named_spectrum("metal-Au-eta"), conductor_k=named_spectrum("metal-Au-k"), interface_roughness=0.2,
                                 conductor_roughness=0.4, albedo=("const", 0.3), g=-0.2, thickness=0.03, interface_eta=named_spectrum("glass-BK7"))
    elif kind == "coatedcondrefl":
        mat = b.coated_conductor(reflectance=_red(), interface_roughness=0.05, remap=False, conductor_roughness=0.3)
    elif kind == "normalmap":
        nm = b.image_texture(procedural_normal_map(32))
        mat = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.15, normal_map=nm)
    elif kind in ("texsph", "texcyl", "texplanar"):
        rgb_img = procedural_image(64, 3)
        if kind == "texsph":
            mp = b.texture_mapping("spherical", texture_from_world=Transform.translate((0.0, -0.6, 0.0)))
        elif kind == "texcyl":
            mp = b.texture_mapping("cylindrical", texture_from_world=Transform.rotate(90.0, (1, 0, 0)) * Transform.translate((0.0, -0.6, 0.0)))
        else:
            mp = b.texture_mapping("planar", v1=(0.7, 0.0, 0.2), v2=(0.0, 0.3, 0.9), udelta=0.1, vdelta=0.25)
        mat = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="trilinear" if kind != "texplanar" else "ewa", mapping=mp))
        ground = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(32, 1), filter="bilinear", mapping=mp))
    elif kind == "mix":
        mat = b.mix(b.diffuse(_green()), b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.05), amount=0.4)
        ground = b.mix(white, b.diffuse(_red()), amount=0.7)
    elif kind == "mixtex":
        amt = b.image_texture(procedural_image(32, 1), filter="bilinear", su=2.0, sv=2.0)
        mat = b.mix(b.diffuse(_green()), b.dielectric(("const", 1.5)), amount_tex=amt)
        ground = b.mix(white, b.coated_diffuse(_red()), amount_tex=amt)
    elif kind == "mixnested":
        inner = b.mix(b.diffuse(_green()), b.diffuse(_red()), amount=0.5)
        mat = b.mix(inner, b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0), amount=0.3)
        ground = b.mix(white, white, amount=1.5)
    elif kind == "textree":
        # the non-image textures (texture.rs:180-310,:537-826): sphere reflectance = mix(rgb image, scaled(constant spectrum, float image),
        # amount = direction_mix(0.1, 0.95, +y)); ground reflectance = direction_mix(constant spectrum, one-channel image, tilted dir)
        # with a displacement = scaled(float image, constant 0.03)
        mono = b.image_texture(procedural_image(32, 1), filter="bilinear", su=3.0, sv=3.0)
        rgb = b.image_texture(procedural_image(64, 3), filter="trilinear")
        red = b.constant_texture(spectrum=b.spectrum(_red()))
        amt = b.direction_mix_texture(b.constant_texture(0.1), b.constant_texture(0.95), dir=(0.0, 1.0, 0.0))
        mat = b.diffuse(_white(), reflectance_tex=b.mix_texture(rgb, b.scaled_texture(red, mono), amt))
        gtex = b.direction_mix_texture(b.constant_texture(spectrum=b.spectrum(_green())), mono, dir=(0.3, 0.8, 0.1))
        ground = b.diffuse(_white(), reflectance_tex=gtex, displacement_tex=b.scaled_texture(mono, b.constant_texture(0.03)))
    elif kind == "textreemix":
        # a three-level float tree as the MixMaterial amount and a scaled-by-zero / amount-0 / amount-1 short-circuit on the ground
        mono = b.image_texture(procedural_image(32, 1), filter="bilinear", su=2.0, sv=2.0)
        inner = b.mix_texture(b.scaled_texture(mono, b.constant_texture(0.8)), b.constant_texture(0.9), mono)
        amt = b.scaled_texture(inner, b.constant_texture(1.1))
        mat = b.mix(b.diffuse(_green()), b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.1), amount_tex=amt)
        rgb = b.image_texture(procedural_image(64, 3), filter="ewa")
        zero = b.scaled_texture(rgb, b.constant_texture(0.0))
        gtex = b.mix_texture(b.mix_texture(zero, rgb, b.constant_texture(1.0)), zero, b.constant_texture(0.0))
        ground = b.diffuse(_white(), reflectance_tex=gtex)
    elif kind == "texparams":
        # texture-valued material parameters (SgMaterialTextures): a conductor whose roughness comes from an image and whose eta / k
        # blend copper into gold across the surface; a coated-diffuse ground with textured thickness, g, albedo and roughness
        mono = b.image_texture(procedural_image(32, 1), filter="bilinear", su=2.0, sv=2.0)
        rgb = b.image_texture(procedural_image(64, 3), filter="trilinear", su=4.0, sv=4.0)
        cs = lambda name: b.constant_texture(spectrum=b.spectrum(named_spectrum(name)))
        mat = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.3)
        b.set_material_textures(mat, u_roughness=b.scaled_texture(mono, b.constant_texture(0.4)), v_roughness=b.constant_texture(0.05),
                                spec_a=b.mix_texture(cs("metal-Cu-eta"), cs("metal-Au-eta"), mono), spec_b=b.mix_texture(cs("metal-Cu-k"), cs("metal-Au-k"), mono))
        ground = b.coated_diffuse(_red(), roughness=0.1, thickness=0.02, albedo=("const", 0.2), g=0.1)
        b.set_material_textures(ground, u_roughness=mono, v_roughness=mono, thickness=b.scaled_texture(mono, b.constant_texture(0.05)),
                                g=b.constant_texture(-0.3), spec_b=b.scaled_texture(rgb, b.constant_texture(0.5)))
    elif kind == "texparams2":
        # rough glass with an image roughness over a coated conductor whose conductor eta / k / roughness and interface roughness are textures
        mono = b.image_texture(procedural_image(32, 1), filter="bilinear", su=3.0, sv=1.0)
        cs = lambda name: b.constant_texture(spectrum=b.spectrum(named_spectrum(name)))
        mat = b.dielectric(("const", 1.5), roughness=0.2)
        b.set_material_textures(mat, u_roughness=b.scaled_texture(mono, b.constant_texture(0.3)), v_roughness=b.scaled_texture(mono, b.constant_texture(0.15)))
        ground = b.coated_conductor(conductor_eta=named_spectrum("metal-Au-eta"), conductor_k=named_spectrum("metal-Au-k"), interface_roughness=0.1,
                                    conductor_roughness=0.2, remap=False)
        b.set_material_textures(ground, u_roughness=b.constant_texture(0.02), spec_a=b.direction_mix_texture(cs("metal-Ag-eta"), cs("metal-Cu-eta"), dir=(0.0, 1.0, 0.0)),
                                spec_d=cs("metal-Cu-k"), u_roughness2=b.scaled_texture(mono, b.constant_texture(0.5)), v_roughness2=mono,
                                thickness=b.constant_texture(0.03), spec_b=b.constant_texture(0.1))
    else:
        raise ValueError(kind)
    P, I, Nn, UV = uv_sphere(10, 14, center=(0.0, 0.6, 0.0), radius=0.6)
    b.add_mesh(P, I, mat, n=Nn, uv=UV)
    gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, ground, uv=guv)
    if area_light:
        lp, li = _quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
        b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b


def tiny_scene(kind="diffuse", resolution=(32, 32)):
    """A few dozen triangles exercising one material each; used by the fast parity tests.  The `tex*` kinds add image
    textures (RGB + one-channel, every filter / wrap mode), bump mapping and specular ray-differential propagation."""
    if kind in INSTANCED_KINDS:
        return instanced_tiny_scene(kind, resolution)
    if kind in SPHERE_KINDS:
        return sphere_tiny_scene(kind, resolution)
    if kind in PATCH_KINDS:
        return patch_tiny_scene(kind, resolution)
    if kind in VARIETY_KINDS:
        return variety_tiny_scene(kind, resolution)
    if kind in INSTANCED_SHAPE_KINDS:
        return instanced_shapes_tiny_scene(kind, resolution)
    if kind == "ortho":
        # OrthographicCamera (camera.rs:657-827) looking down +z with up = y: render_from_camera is the identity in the
        # camera-world rendering space, where the reference's camera-space ray (see SgCameraKind) is also the right one.
        b = SceneBuilder()
        b.set_camera(pos=(0.0, 0.8, -3.0), look=(0.0, 0.8, 0.0), up=(0, 1, 0), fov=45.0, resolution=resolution, kind="orthographic",
                     screen_window=(-1.8, 1.8, -1.8, 1.8))
        ground = b.diffuse(_white(), reflectance_tex=b.image_texture(procedural_image(64, 3), filter="trilinear", su=3.0, sv=3.0))
        P, I, Nn, UV = uv_sphere(10, 14, center=(0.0, 0.6, 0.0), radius=0.6)
        b.add_mesh(P, I, b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0), n=Nn, uv=UV)
        gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, ground, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
        wp, wi = _quad((-3, 0.0, 2.0), (-3, 3, 2.0), (3, 3, 2.0), (3, 0.0, 2.0))
        b.add_mesh(wp, wi, ground, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
        lp, li = _quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
        b.add_mesh(lp, li, b.diffuse(_white()), area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
        return b
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=resolution)
    white = b.diffuse(_white())
    if kind == "diffuse":
        mat = b.diffuse(_green())
    elif kind == "conductor":
        mat = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.1)
    elif kind == "mirror":
        mat = b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0)
    elif kind == "glass":
        mat = b.dielectric(named_spectrum("glass-BK7"))
    elif kind == "roughglass":
        mat = b.dielectric(("const", 1.5), roughness=0.2)
    elif kind == "thinglass":
        mat = b.thin_dielectric(named_spectrum("glass-BK7"))
    elif kind == "coated":
        mat = b.coated_diffuse(_red())
    elif kind == "coatedrough":
        mat = b.coated_diffuse(_green(), roughness=0.15, albedo=("const", 0.4), g=0.3, thickness=0.05)
    elif kind in TEXTURED_KINDS:
        rgb_img, mono_img = procedural_image(64, 3), procedural_image(32, 1)
        guv = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32)
        if kind == "tex":        # bilinear RGB ground (tiled), trilinear one-channel sphere with interpolated normals
            ground = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="bilinear", su=3.0, sv=3.0, du=0.25))
            mat = b.diffuse(_white(), reflectance_tex=b.image_texture(mono_img, filter="trilinear", su=2.0, sv=1.0))
        elif kind == "texewa":   # EWA RGB ground seen directly and through a smooth mirror (specular reflection differentials)
            ground = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="ewa", su=5.0, sv=5.0, max_anisotropy=8.0))
            mat = b.conductor(named_spectrum("metal-Ag-eta"), named_spectrum("metal-Ag-k"), roughness=0.0)
        elif kind == "texbump":  # point-filtered, clamped, unbounded-spectrum ground with an image displacement; glass sphere (transmission differentials)
            ground = b.diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="point", wrap="clamp", su=1.5, sv=1.5, scale=0.9,
                                                                         spectrum_type="unbounded"),
                               displacement_tex=b.image_texture(mono_img, filter="bilinear", su=6.0, sv=6.0, scale=0.05))
            mat = b.dielectric(("const", 1.5))
        else:                    # coated diffuse with an inverted, black-wrapped texture + bump on the sphere
            ground = b.coated_diffuse(_white(), reflectance_tex=b.image_texture(rgb_img, filter="trilinear", wrap="black", su=1.3, sv=1.3, du=-0.15,
                                                                                 invert=True))
            mat = b.diffuse(_green(), displacement_tex=b.image_texture(mono_img, filter="ewa", su=4.0, sv=2.0, scale=0.02))
        P, I, Nn, UV = uv_sphere(10, 14, center=(0.0, 0.6, 0.0), radius=0.6)
        b.add_mesh(P, I, mat, n=Nn, uv=UV)
        gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, ground, uv=guv)
    else:
        raise ValueError(kind)
    if kind not in TEXTURED_KINDS:
        P, I = displaced_sphere(12, 16, center=(0.0, 0.6, 0.0), radius=0.6, amp=0.0)
        b.add_mesh(P, I, mat)
        gp, gi = _quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, white)
    lp, li = _quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    if kind == "mirror":
        b.add_uniform_infinite_light(("const", 1.0), scale=0.3)
    return b


CONFIGS = {
    "cornell": dict(builder=cornell_box, resolution=(512, 512), spp=16, max_depth=5,
                    desc="C1 synthetic Cornell box, 32 triangles, diffuse + area light, 512x512, 16 spp"),
    "mesh1m": dict(builder=mesh_scene, resolution=(1024, 1024), spp=64, max_depth=5,
                   desc="C2 procedural ~1M-triangle displaced sphere, diffuse + Cu conductor, 1024x1024, 64 spp"),
    "glass": dict(builder=glass_scene, resolution=(1024, 1024), spp=256, max_depth=5,
                  desc="C3 glass dispersion (tabulated BK7 eta), 1024x1024, 256 spp"),
    "instanced": dict(builder=instanced_scene, resolution=(1920, 1080), spp=128, max_depth=5,
                      desc="C4 100 instances x ~100k-triangle uv-mapped mesh (10M instanced triangles), EWA/bilinear/bump image textures, "
                           "1024 emissive triangles, 1920x1080, 128 spp"),
    "composite": dict(builder=composite_scene, resolution=(3840, 2160), spp=1024, max_depth=5,
                      desc="C5 Cornell + 1M-triangle mesh composite, 3840x2160, 1024 spp"),
}
