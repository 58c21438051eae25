"""PBRT-v4 emission of the synthetic scenes (shimmer_b200/pbrt_export.py): the emitted files describe the scene the GPU path
renders -- same meshes (binary PLY round trip is bit-exact), same spectra tables, same shape order -- using only directives
shimmer's parser implements, so that real shimmer can render C1..C5 on a machine with a Rust toolchain."""
import os
import re

import numpy as np
import pytest

from shimmer_b200 import scenes
from shimmer_b200.pbrt_export import read_ply, write_pbrt, write_png

# directives / types shimmer implements (loading/parser.rs, scene.rs, shape.rs:66-135, material.rs:66-160, texture.rs:108-130,433-480,
# light.rs:105-260, camera.rs:70-78, film.rs:102, sampler.rs:41, filter.rs:28)
DIRECTIVES = {"Option", "LookAt", "Camera", "Sampler", "Integrator", "PixelFilter", "Film", "WorldBegin", "Texture", "MakeNamedMaterial",
              "NamedMaterial", "AttributeBegin", "AttributeEnd", "ConcatTransform", "ReverseOrientation", "AreaLightSource", "LightSource",
              "Shape", "ObjectBegin", "ObjectEnd", "ObjectInstance"}
SHAPES = {"plymesh", "sphere"}
MATERIALS = {"diffuse", "conductor", "dielectric", "thindielectric", "coateddiffuse", "coatedconductor", "mix"}


def _directives(path):
    out = []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#") or line.startswith('"'):
            continue
        out.append(line)
    return out


def test_cornell_export_round_trips(tmp_path):
    b = scenes.cornell_box(resolution=(512, 512))
    path = write_pbrt(b, str(tmp_path), spp=16, max_depth=5)
    lines = _directives(path)
    assert all(l.split()[0] in DIRECTIVES for l in lines), [l for l in lines if l.split()[0] not in DIRECTIVES][:3]
    assert sum(l.startswith("Shape") for l in lines) == len(b.meshes)
    assert all(re.search(r'Shape "(\w+)"', l).group(1) in SHAPES for l in lines if l.startswith("Shape"))
    assert any(l.startswith('Sampler "independent" "integer pixelsamples" [ 16 ]') for l in lines)
    assert any(l.startswith('Film "rgb" "integer xresolution" [ 512 ]') for l in lines)
    assert sum(l.startswith("AreaLightSource") for l in lines) == sum(m["area_light"] is not None for m in b.meshes)
    # meshes: the PLY holds exactly the world-space vertices the builder was given, indices unchanged
    for mi, m in enumerate(b.meshes):
        ply = read_ply(os.path.join(str(tmp_path), "mesh%d.ply" % mi))
        assert np.array_equal(ply["verts"][:, :3], m["src"]["p"]) and np.array_equal(ply["faces"], m["idx"].astype(np.int32))
    # spectra: every table in the file is one of the builder's (lambda, value) tables, digit for digit after f32 parsing
    txt = open(path).read()
    tables = [np.array(t.split(), np.float32) for t in re.findall(r'"spectrum \w+" \[ ([^\]]+) \]', txt)]
    specs = list(b.spectra) + [m["area_light"]["L"] for m in b.meshes if m["area_light"] is not None]
    have = [np.stack([np.asarray(s[1], np.float32), np.asarray(s[2], np.float32)], 1).ravel() for s in specs if s[0] == "pl"]
    assert tables and all(any(len(t) == len(h) and np.array_equal(t, h) for h in have) for t in tables)


def test_instanced_textured_export(tmp_path):
    b = scenes.instanced_scene(n_theta=8, n_phi=8, grid=2, resolution=(64, 36), lights=(2, 2), tex_size=32)     # C4 in miniature
    path = write_pbrt(b, str(tmp_path), spp=8)
    assert "quantised" not in open(path).read()                                  # C4's textures are 8-bit exact
    lines = _directives(path)
    assert all(l.split()[0] in DIRECTIVES for l in lines)
    assert sum(l.startswith("ObjectInstance") or "ObjectInstance" in l for l in lines) == len(b.instances)
    assert sum(l.startswith("ObjectBegin") for l in lines) == b.n_objects == sum(l.startswith("ObjectEnd") for l in lines)
    mats = re.findall(r'MakeNamedMaterial "mat\d+" "string type" "(\w+)"', open(path).read())
    assert len(mats) == len(b.materials) and set(mats) <= MATERIALS
    # textures are declared before the materials that use them
    txt = open(path).read()
    for name in re.findall(r'"texture \w+" "(tex\d+)"', txt):
        assert txt.index('Texture "%s"' % name) < txt.index('"%s"' % name, txt.index("MakeNamedMaterial"))


def test_png_writer_is_exact_for_8_bit_images(tmp_path):
    img = scenes.procedural_image(32, 3, quantize8=True)
    assert write_png(str(tmp_path / "a.png"), img) == 0.0                       # C4's textures survive the PNG bit-exactly
    assert write_png(str(tmp_path / "b.png"), scenes.procedural_image(32, 3)) > 0.0
    raw = open(str(tmp_path / "a.png"), "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and raw[12:16] == b"IHDR"
    import struct
    import zlib
    w, h, depth, ctype = struct.unpack(">IIBB", raw[16:26])
    assert (w, h, depth, ctype) == (32, 32, 8, 2)
    idat = raw[raw.index(b"IDAT") + 4: raw.index(b"IEND") - 8]
    px = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(32, 1 + 32 * 3)[:, 1:].reshape(32, 32, 3)
    assert np.array_equal(px.astype(np.float32) / np.float32(255.0), img)


@pytest.mark.parametrize("kind", ["glass", "coated", "conductor", "spheres", "mix", "textree", "texparams"])
def test_material_and_shape_kinds_emit(kind, tmp_path):
    b = scenes.tiny_scene(kind, resolution=(16, 16))
    path = write_pbrt(b, str(tmp_path), spp=4)
    lines = _directives(path)
    assert all(l.split()[0] in DIRECTIVES for l in lines)
    mats = re.findall(r'MakeNamedMaterial "mat\d+" "string type" "(\w+)"', open(path).read())
    assert len(mats) == len(b.materials) and set(mats) <= MATERIALS
