"""Protobuf scene ingest (include/izpi_proto.h): the product's own wire/text decoder + ToScene
conversion, checked against (a) the reference's example scene (golden .izpi generated from the reference's .pbtxt by the
stock Python protobuf runtime, tests/golden/make_golden_proto.py), (b) messages built with the stock runtime and
converted independently here in Python, following internal/transport/transport.go."""
import ctypes as C
import os
import struct

import numpy as np
import pytest
from google.protobuf import text_format

import proto_schema as ps
from izpi_b200 import cuda, proto, scenes
from izpi_b200 import scene as S
from izpi_b200.build import build as build_lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "cornell_pyramid_spectral.izpi")


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build_lib()


def f32(x):
    return float(np.float32(x))


# ---- reading an izpi_scene_spec back ------------------------------------------------------------
def spec_view(c: S.SceneSpecC):
    prims = np.frombuffer(C.string_at(c.prims, c.n_prims * S.PRIM_DTYPE.itemsize), dtype=S.PRIM_DTYPE).copy() if c.n_prims else np.zeros(0, S.PRIM_DTYPE)
    mats = [S.MaterialSpec.from_buffer_copy(C.string_at(c.materials + i * C.sizeof(S.MaterialSpec), C.sizeof(S.MaterialSpec))) for i in range(c.n_materials)]
    texs = [S.TextureSpec.from_buffer_copy(C.string_at(c.textures + i * C.sizeof(S.TextureSpec), C.sizeof(S.TextureSpec))) for i in range(c.n_textures)]
    stex = [S.SpectralTextureSpec.from_buffer_copy(C.string_at(c.spectral_textures + i * C.sizeof(S.SpectralTextureSpec), C.sizeof(S.SpectralTextureSpec)))
            for i in range(c.n_spectral_textures)]
    return prims, mats, texs, stex


def spectral_desc(stex, i, texs=None):
    """A comparable description of spectral texture i."""
    if i < 0:
        return None
    t = stex[i]
    if t.type == S.SPEC_GAUSSIAN:
        return ("gauss", t.peak, t.centre, t.width)
    if t.type == S.SPEC_IMAGE:
        return ("spectral-image", texture_desc(texs, t.n))  # the index of the RGB image is an implementation detail
    w = np.frombuffer(C.string_at(t.wavelengths, 8 * t.n), dtype=np.float64)
    v = np.frombuffer(C.string_at(t.values, 8 * t.n), dtype=np.float64)
    return ("tab", tuple(w), tuple(v))


def texture_desc(texs, i):
    if i < 0:
        return None
    t = texs[i]
    if t.type == S.TEX_CONSTANT:
        return ("const", tuple(t.color))
    return ("image", t.width, t.height, t.pixels)


def material_desc(m, texs, stex):
    return (m.type, texture_desc(texs, m.tex), spectral_desc(stex, m.spectral_tex, texs), spectral_desc(stex, m.spectral_absorption_tex, texs),
            texture_desc(texs, m.normal_tex), texture_desc(texs, m.roughness_tex), texture_desc(texs, m.metalness_tex),
            m.compute_beer_lambert, tuple(m.v), m.s)


def resolved(c):
    """Primitives with their material replaced by its full description (material indices are an implementation detail)."""
    prims, mats, texs, stex = spec_view(c)
    descs = [material_desc(m, texs, stex) for m in mats]
    return [(int(p["type"]), tuple(p["p"]), int(p["wrap"]), descs[int(p["material"])]) for p in prims]


def camera_tuple(c):
    cam = c.camera
    return (tuple(cam.look_from), tuple(cam.look_at), tuple(cam.vup), cam.vfov, cam.aspect, cam.aperture, cam.focus_dist, cam.time0, cam.time1, cam.exposure)


# ---- golden: the reference's example scene --------------------------------------------------------
def test_golden_example_scene_matches_restated_config4():
    ps_scene = proto.ProtoScene.from_file(GOLDEN)
    assert ps_scene.name == "Cornell Box Transparent Pyramid Spectral"
    assert ps_scene.colour_representation == proto.COLOUR_SPECTRAL and ps_scene.sampler == cuda.SAMPLER_SPECTRAL
    assert ps_scene.num_parsed_triangles == 12
    ps_scene.to_scene(aspect_override=1.0)
    got = ps_scene.to_c()
    want = scenes.spectral_pyramid(1.0).to_c()  # config 4 as restated by hand from SURVEY.md A.8
    assert got.world_kind == S.WORLD_BVH4 and got.n_prims == want.n_prims == 22
    assert resolved(got) == resolved(want)
    assert camera_tuple(got) == camera_tuple(want)
    # and the host runtime builds the same BVH4 from both
    a, b = cuda.HostScene(ps_scene).bvh(), cuda.HostScene(scenes.spectral_pyramid(1.0)).bvh()
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


def test_golden_renders_identically_in_the_oracle(oracle_mod):
    """End to end on the CPU checker: the ingested scene and the hand-restated one give the same image."""
    sc = proto.ProtoScene.from_file(GOLDEN).to_scene(aspect_override=1.0)
    a, _ = oracle_mod.OracleScene(sc).render(24, 24, 2, max_depth=50, sampler=1, rng_mode=1, seed=5)
    b, _ = oracle_mod.OracleScene(scenes.spectral_pyramid(1.0)).render(24, 24, 2, max_depth=50, sampler=1, rng_mode=1, seed=5)
    assert a.tobytes() == b.tobytes()
    assert np.isfinite(a).all() and a[..., :3].max() > 0


def test_text_format_equals_binary():
    msg = ps.Scene()
    msg.ParseFromString(open(GOLDEN, "rb").read())
    for text in (text_format.MessageToString(msg), text_format.MessageToString(msg, as_one_line=True),
                 text_format.MessageToString(msg, use_short_repeated_primitives=True, pointy_brackets=True)):
        t = proto.ProtoScene(text.encode(), proto.TEXT).to_scene(aspect_override=1.0)
        b = proto.ProtoScene.from_file(GOLDEN).to_scene(aspect_override=1.0)
        assert resolved(t.to_c()) == resolved(b.to_c())
        assert camera_tuple(t.to_c()) == camera_tuple(b.to_c())
        assert t.name == b.name


def test_text_format_details():
    text = b'''
    # comment
    name: "a" "b" 'c\\x41\\n'   # adjacent strings concatenate
    colour_representation: 1
    camera < vfov: 1e1f aspect: 0.1 lookfrom { x: -inf y: 0x10 } >
    materials { key: "m" value { name: "m" type: METAL metal { albedo { x: .5 y: 5e-1 z: 0.5 } fuzz: 0.1 } } };
    objects { spheres: [ { radius: 2 material_name: "m" }, { radius: 3, material_name: "m" } ] }
    spectral_background { wavelengths: [380, 750] values: 1 values: 2 }
    total_triangles: 7 stream_triangles: true
    '''
    s = proto.ProtoScene(text, proto.TEXT)
    assert s.name == "abcA\n" and s.colour_representation == proto.COLOUR_RGB and s.total_triangles == 7 and s.stream_triangles
    w, v = s.background()
    assert list(w) == [380.0, 750.0] and list(v) == [1.0, 2.0]
    c = s.to_scene().to_c()
    assert c.camera.vfov == 10.0 and c.camera.aspect == f32(0.1) and c.camera.look_from[0] == -np.inf and c.camera.look_from[1] == 16.0
    prims, mats, _, _ = spec_view(c)
    assert [p["p"][3] for p in prims] == [2.0, 3.0] and mats[0].type == S.MAT_METAL and mats[0].s == f32(0.1)
    for bad, why in [(b"nme: 'x'", "unknown field"), (b"name: 'x' name: 'y'", "multiple times"), (b"camera { vfov: abc }", "invalid number"),
                     (b"camera { vfov: 1", "unexpected end"), (b"colour_representation: PURPLE", "unknown enum"),
                     (b"materials { value { lambert { albedo {} spectral_albedo {} } } }", "oneof")]:
        with pytest.raises(cuda.IzpiError) as e:
            proto.ProtoScene(bad, proto.TEXT)
        assert why in str(e.value), (bad, str(e.value))


# ---- messages built with the stock runtime, converted independently ---------------------------------
def _vec3(v, xyz):
    v.x, v.y, v.z = xyz


def _mixed_scene(spectral=False):
    m = ps.Scene(name="mixed", version="1")
    m.colour_representation = 2 if spectral else 1
    cam = m.camera
    _vec3(cam.lookfrom, (1.1, 2.2, 3.3)); _vec3(cam.lookat, (0.1, 0.2, 0.3)); _vec3(cam.vup, (0, 1, 0))
    cam.vfov, cam.aspect, cam.aperture, cam.focusdist, cam.time0, cam.time1, cam.exposure = 33.3, 1.7777, 0.05, 9.9, 0.25, 0.75, 1.3
    mt = m.materials
    a = mt["lam_const"]; a.name, a.type = "lam_const", 4; _vec3(a.lambert.albedo.constant.value, (0.1, 0.2, 0.3))
    a = mt["lam_img"]; a.name, a.type = "lam_img", 4; a.lambert.albedo.image.filename = "albedo.png"
    a = mt["lam_spec"]; a.name, a.type = "lam_spec", 4
    a.lambert.spectral_albedo.gaussian.peak_value, a.lambert.spectral_albedo.gaussian.center_wavelength, a.lambert.spectral_albedo.gaussian.width = 0.9, 540, 40
    a = mt["lam_spec_as_rgb"]; a.name, a.type = "lam_spec_as_rgb", 4; a.lambert.albedo.spectral_constant.neutral.reflectance = 0.3
    a = mt["glass"]; a.name, a.type = "glass", 1; a.dielectric.refidx = 1.5; a.dielectric.compute_beer_lambert_attenuation = True
    a = mt["glass_col"]; a.name, a.type = "glass_col", 1; a.dielectric.refidx = 1.45; _vec3(a.dielectric.absorption_coeff, (0.1, 0.0, 0.2))
    a = mt["glass_spec"]; a.name, a.type = "glass_spec", 1
    a.dielectric.spectral_refidx.tabulated.wavelengths.extend([380, 550, 750]); a.dielectric.spectral_refidx.tabulated.values.extend([1.52, 1.5, 1.48])
    a.dielectric.compute_beer_lambert_attenuation = True
    a = mt["glass_spec_col"]; a.name, a.type = "glass_spec_col", 1
    a.dielectric.spectral_refidx.neutral.reflectance = 1.5; a.dielectric.spectral_absorption_coeff.neutral.reflectance = 0.02
    a.dielectric.compute_beer_lambert_attenuation = True
    a = mt["lamp"]; a.name, a.type = "lamp", 2; _vec3(a.diffuselight.emit.constant.value, (15, 15, 15))
    a = mt["lamp_spec"]; a.name, a.type = "lamp_spec", 2; a.diffuselight.spectral_emit.from_light_source_library.light_source_name = "cie_illuminant_a_2856k"
    a = mt["lamp_custom"]; a.name, a.type = "lamp_custom", 2; a.diffuselight.spectral_emit.from_light_source_library.light_source_name = "my_led"
    a = mt["steel"]; a.name, a.type = "steel", 5; _vec3(a.metal.albedo, (0.7, 0.6, 0.5)); a.metal.fuzz = 0.2
    a = mt["pbr"]; a.name, a.type = "pbr", 6
    a.pbr.albedo.image.filename = "albedo.png"; a.pbr.roughness.image.filename = "rough.png"; _vec3(a.pbr.metalness.constant.value, (0.5, 0.5, 0.5))
    a.pbr.normal_map.image.filename = "normal.png"; _vec3(a.pbr.sss.constant.value, (0, 0, 0)); a.pbr.sss_radius = 0.1
    a = mt["pbr_const"]; a.name, a.type = "pbr_const", 6
    for t in (a.pbr.albedo, a.pbr.roughness, a.pbr.metalness, a.pbr.normal_map, a.pbr.sss):
        _vec3(t.constant.value, (0.2, 0.4, 0.6))
    a = mt["skipped"]; a.name = "skipped"  # MATERIAL_TYPE_UNSPECIFIED: silently skipped by toSceneMaterial
    names = [k for k in sorted(mt) if k != "skipped"]
    rng = np.random.default_rng(3)
    for i in range(40):
        t = m.objects.triangles.add()
        p = rng.uniform(-50, 50, (3, 3))
        _vec3(t.vertex0, p[0]); _vec3(t.vertex1, p[1]); _vec3(t.vertex2, p[2])
        uv = rng.uniform(0, 1, (3, 2))
        t.uv0.u, t.uv0.v = uv[0]; t.uv1.u, t.uv1.v = uv[1]; t.uv2.u, t.uv2.v = uv[2]
        _vec3(t.normal0, (0, 1, 0))  # ignored by toSceneTriangle
        t.material_name = names[i % len(names)]
    for i in range(7):
        s = m.objects.spheres.add()
        _vec3(s.center, rng.uniform(-50, 50, 3)); s.radius = float(rng.uniform(1, 5)); s.material_name = names[(3 * i) % len(names)]
    m.image_textures["albedo.png"].filename = "albedo.png"
    m.image_textures["rough.png"].filename = "rough.png"
    m.image_textures["normal.png"].filename = "normal.png"
    m.spectral_background.wavelengths.extend([380, 750]); m.spectral_background.values.extend([0.5, 0.25])
    return m


IMAGES = {"albedo.png": np.random.default_rng(1).uniform(0, 1, (4, 8, 4)), "rough.png": np.random.default_rng(2).uniform(0, 1, (2, 2, 4)),
          "normal.png": np.random.default_rng(3).uniform(0, 1, (3, 5, 4))}
LED = (np.array([400.0, 500, 600, 700]), np.array([0.1, 1.0, 0.6, 0.05]))


def _expected(m, spectral, images, aspect_override=0.0):
    """transport.go, restated in Python on top of the stock runtime's message accessors."""
    sc = S.SceneSpec(world_kind=S.WORLD_BVH4)
    img_tex = {}

    def spec_tex(t):
        which = t.WhichOneof("spectral_properties")
        if which == "gaussian":
            return sc.spectral_gaussian(t.gaussian.peak_value, t.gaussian.center_wavelength, t.gaussian.width)
        if which == "tabulated":
            return sc.spectral_tabulated(list(t.tabulated.wavelengths), list(t.tabulated.values))
        if which == "neutral":
            return sc.spectral_neutral(t.neutral.reflectance)
        name = t.from_light_source_library.light_source_name
        if name == "my_led":
            return sc.spectral_tabulated(*LED)
        temp = {"incandescent_2800k": 2800.0, "halogen_3200k": 3200.0, "cie_illuminant_a_2856k": 2856.0}[name]
        h, c, k = 6.62607015e-34, 2.99792458e8, 1.380649e-23  # spectral.NewBlackbodySPD
        w = 380.0 + 5.0 * np.arange(75)
        lam = w * 1e-9
        import math  # libm exp, as the C++ side (numpy's vectorised exp may differ in the last place)
        v = np.array([(2.0 * h * c * c) / ((x * x * x * x * x) * (math.exp(((h * c) / k) / (x * temp)) - 1.0)) for x in lam])
        return sc.spectral_tabulated(w, v / v.max())

    def tex(t):
        which = t.WhichOneof("texture_properties")
        if which == "constant":
            return sc.constant_texture((t.constant.value.x, t.constant.value.y, t.constant.value.z))
        if which == "image":
            if t.image.filename not in img_tex:
                img_tex[t.image.filename] = sc.image_texture(images[t.image.filename])
            return img_tex[t.image.filename]
        return sc.constant_texture((0.5, 0.5, 0.5))

    mats = {}
    for key in m.materials:
        a = m.materials[key]
        if a.type == 4:
            mats[a.name] = sc.lambertian(tex(a.lambert.albedo)) if a.lambert.WhichOneof("albedo_properties") == "albedo" else sc.spectral_lambertian(spec_tex(a.lambert.spectral_albedo))
        elif a.type == 1:
            d = a.dielectric
            if d.WhichOneof("refractive_index_properties") == "spectral_refidx":
                ref = spec_tex(d.spectral_refidx)
                if d.WhichOneof("absorption_properties") == "spectral_absorption_coeff":
                    mats[a.name] = sc.spectral_dielectric(ref, spec_tex(d.spectral_absorption_coeff), False)
                else:
                    mats[a.name] = sc.spectral_dielectric(ref, -1, d.compute_beer_lambert_attenuation)
            else:
                ab = (d.absorption_coeff.x, d.absorption_coeff.y, d.absorption_coeff.z)
                mats[a.name] = sc.colored_dielectric(d.refidx, ab) if any(ab) else sc.dielectric(d.refidx)
        elif a.type == 2:
            e = a.diffuselight
            mats[a.name] = sc.diffuse_light(tex(e.emit)) if e.WhichOneof("emission_properties") == "emit" else sc.spectral_diffuse_light(spec_tex(e.spectral_emit))
        elif a.type == 5:
            mats[a.name] = sc.metal((a.metal.albedo.x, a.metal.albedo.y, a.metal.albedo.z), a.metal.fuzz)
        elif a.type == 6:
            p = a.pbr
            alb, rough, metal, nrm = tex(p.albedo), tex(p.roughness), tex(p.metalness), tex(p.normal_map)
            tex(p.sss)
            sa = -1
            if spectral:
                t = sc.textures[alb]
                sa = sc.spectral_image(alb) if t.type == S.TEX_IMAGE else sc.spectral_neutral(0.299 * t.color[0] + 0.587 * t.color[1] + 0.114 * t.color[2])
            mats[a.name] = sc.pbr(alb, nrm, rough, metal, sa)
    for t in m.objects.triangles:
        v = [[t.vertex0.x, t.vertex0.y, t.vertex0.z], [t.vertex1.x, t.vertex1.y, t.vertex1.z], [t.vertex2.x, t.vertex2.y, t.vertex2.z]]
        uv = [[t.uv0.u, t.uv0.v], [t.uv1.u, t.uv1.v], [t.uv2.u, t.uv2.v]]
        sc.triangles(np.array([v]), mats[t.material_name], np.array([uv]))
    for s in m.objects.spheres:
        sc.sphere((s.center.x, s.center.y, s.center.z), s.radius, mats[s.material_name])
    c = m.camera
    sc.set_camera((c.lookfrom.x, c.lookfrom.y, c.lookfrom.z), (c.lookat.x, c.lookat.y, c.lookat.z), (c.vup.x, c.vup.y, c.vup.z), c.vfov,
                  aspect_override or c.aspect, c.aperture, c.focusdist, c.time0, c.time1, c.exposure)
    return sc


@pytest.mark.parametrize("spectral", [False, True])
@pytest.mark.parametrize("fmt", ["binary", "text"])
def test_every_material_kind(spectral, fmt):
    m = _mixed_scene(spectral)
    payload = m.SerializeToString() if fmt == "binary" else text_format.MessageToString(m).encode()
    s = proto.ProtoScene(payload, proto.BINARY if fmt == "binary" else proto.TEXT)
    assert sorted(s.image_filenames()) == ["albedo.png", "normal.png", "rough.png"] and s.image_filenames(displacement=True) == []
    s.to_scene(textures=IMAGES, light_sources={"my_led": LED})
    want = _expected(m, spectral, IMAGES)
    got_c, want_c = s.to_c(), want.to_c()
    g, w = resolved(got_c), resolved(want_c)

    def strip_ptr(r):  # image pixel pointers differ between the two specs; compare sizes here, contents below
        def fix(d):
            if isinstance(d, tuple) and d and d[0] == "image" and len(d) == 4:
                return ("image", d[1], d[2])
            return tuple(fix(x) for x in d) if isinstance(d, tuple) else d
        return [(a, b, c, fix(d)) for a, b, c, d in r]
    assert strip_ptr(g) == strip_ptr(w)
    assert camera_tuple(got_c) == camera_tuple(want_c)
    _, _, gt, _ = spec_view(got_c)
    for t in gt:
        if t.type == S.TEX_IMAGE:
            px = np.frombuffer(C.string_at(t.pixels, 32 * t.width * t.height), dtype=np.float64)
            assert any(px.tobytes() == np.ascontiguousarray(im).tobytes() for im in IMAGES.values())
    w_bg, v_bg = s.background()
    assert list(w_bg) == [380.0, 750.0] and list(v_bg) == [0.5, 0.25]
    # aspect override (leader.go:46)
    s2 = proto.ProtoScene(payload, proto.BINARY if fmt == "binary" else proto.TEXT).to_scene(aspect_override=2.5, textures=IMAGES, light_sources={"my_led": LED})
    assert s2.to_c().camera.aspect == 2.5


def _key(num, wt):
    return bytes([(num << 3) | wt])


def _ld(num, payload):
    assert len(payload) < 128
    return _key(num, 2) + bytes([len(payload)]) + payload


def test_wire_format_corner_cases():
    vec = lambda x, y, z: _key(1, 5) + struct.pack("<f", x) + _key(2, 5) + struct.pack("<f", y) + _key(3, 5) + struct.pack("<f", z)  # noqa: E731
    metal = _ld(1, b"m") + _key(2, 0) + bytes([5]) + _ld(7, _ld(1, vec(0.5, 0.5, 0.5)))
    # repeated floats UNPACKED (legal for proto3 parsers), an unknown field (number 15, varint) and an unknown length-delimited one
    tab_unpacked = b"".join(_key(1, 5) + struct.pack("<f", w) for w in (380.0, 750.0)) + b"".join(_key(2, 5) + struct.pack("<f", v) for v in (1.0, 2.0))
    tab_packed = _ld(1, struct.pack("<2f", 400.0, 700.0)) + _ld(2, struct.pack("<2f", 3.0, 4.0))
    tri = _ld(1, vec(0, 0, 0)) + _ld(2, vec(1, 0, 0)) + _ld(3, vec(0, 1, 0)) + _ld(10, b"m") + _key(15, 0) + bytes([1])
    # vertex0 split over two occurrences: the second merges into the first (only y overridden)
    tri2 = _ld(1, vec(5, 6, 7)) + _ld(1, _key(2, 5) + struct.pack("<f", 9.0)) + _ld(2, vec(1, 0, 0)) + _ld(3, vec(0, 1, 0)) + _ld(10, b"m")
    scene = (_ld(1, b"corner") + _ld(5, _ld(1, b"m") + _ld(2, metal)) + _ld(8, _ld(1, tri)) + _ld(8, _ld(1, tri2)) + _ld(14, b"junk")
             + _ld(11, tab_unpacked) + _ld(11, tab_packed) + _key(9, 0) + bytes([1]))
    stock = ps.Scene()
    stock.ParseFromString(scene)  # the stock runtime accepts the same bytes
    assert len(stock.objects.triangles) == 2 and stock.objects.triangles[1].vertex0.y == 9.0 and stock.objects.triangles[1].vertex0.x == 5.0
    s = proto.ProtoScene(scene)
    assert s.name == "corner" and s.stream_triangles and s.num_parsed_triangles == 2
    w, v = s.background()
    assert list(w) == list(stock.spectral_background.wavelengths) == [380.0, 750.0, 400.0, 700.0] and list(v) == [1.0, 2.0, 3.0, 4.0]
    prims, _, _, _ = spec_view(s.to_scene().to_c())
    assert tuple(prims[1]["p"][:3]) == (5.0, 9.0, 7.0)
    # oneof: the last member on the wire wins (refidx after spectral_refidx)
    diel = _ld(2, _ld(3, _key(1, 5) + struct.pack("<f", 1.7))) + _key(1, 5) + struct.pack("<f", 1.25)
    sc2 = _ld(5, _ld(1, b"g") + _ld(2, _ld(1, b"g") + _key(2, 0) + bytes([1]) + _ld(3, diel))) + _ld(8, _ld(2, _ld(1, vec(0, 0, 0)) + _key(2, 5) + struct.pack("<f", 1.0) + _ld(3, b"g")))
    stock2 = ps.Scene()
    stock2.ParseFromString(sc2)
    assert stock2.materials["g"].dielectric.WhichOneof("refractive_index_properties") == "refidx"
    _, mats, _, _ = spec_view(proto.ProtoScene(sc2).to_scene().to_c())
    assert mats[0].type == S.MAT_DIELECTRIC and mats[0].spectral_tex == -1 and mats[0].s == 1.25
    # duplicate map key: the last value replaces the first (Go map semantics)
    m1 = _ld(1, b"m") + _key(2, 0) + bytes([5]) + _ld(7, _key(2, 5) + struct.pack("<f", 0.1))
    m2 = _ld(1, b"m") + _key(2, 0) + bytes([5]) + _ld(7, _key(2, 5) + struct.pack("<f", 0.9))
    sc3 = _ld(5, _ld(1, b"k") + _ld(2, m1)) + _ld(5, _ld(1, b"k") + _ld(2, m2))
    _, mats, _, _ = spec_view(proto.ProtoScene(sc3).to_scene().to_c())
    assert len(mats) == 1 and mats[0].s == f32(0.9)
    # truncated / malformed input
    for bad in (scene[:-3], b"\x0a\xff", b"\x0b"):
        with pytest.raises(cuda.IzpiError):
            proto.ProtoScene(bad)


def test_streamed_triangles_follow_embedded_ones():
    m = _mixed_scene()
    resp = ps.StreamTrianglesResponse(total_triangles=2)
    for k in range(2):
        t = resp.triangles.add()
        _vec3(t.vertex0, (k, 0, 0)); _vec3(t.vertex1, (k, 1, 0)); _vec3(t.vertex2, (k, 0, 1))
        t.material_name = "steel"
    s = proto.ProtoScene(m.SerializeToString())
    s.append_triangles(resp.SerializeToString())
    assert s.num_parsed_triangles == 42
    prims, _, _, _ = spec_view(s.to_scene(textures=IMAGES, light_sources={"my_led": LED}).to_c())
    assert len(prims) == 49
    assert [int(t) for t in prims["type"][:42]] == [S.PRIM_TRIANGLE] * 42 and tuple(prims[41]["p"][:3]) == (1.0, 0.0, 0.0)
    assert [int(t) for t in prims["type"][42:]] == [S.PRIM_SPHERE] * 7  # triangles first, then spheres (transport.go:551-566)
    with pytest.raises(cuda.IzpiError):
        s.append_triangles(resp.SerializeToString())  # already converted


def test_conversion_errors_mirror_the_reference():
    def scene_with(fn):
        m = ps.Scene()
        fn(m)
        return proto.ProtoScene(m.SerializeToString())

    def missing_material(m):
        t = m.objects.triangles.add(); t.material_name = "nope"
    def missing_sphere_material(m):
        s = m.objects.spheres.add(); s.material_name = "nope"
    def missing_texture(m):
        a = m.materials["a"]; a.name, a.type = "a", 4; a.lambert.albedo.image.filename = "x.png"
    def lambert_without_albedo(m):
        a = m.materials["a"]; a.name, a.type = "a", 4
    def dielectric_without_refidx(m):
        a = m.materials["a"]; a.name, a.type = "a", 1; a.dielectric.compute_beer_lambert_attenuation = True
    def light_without_emit(m):
        a = m.materials["a"]; a.name, a.type = "a", 2
    def checker(m):
        a = m.materials["a"]; a.name, a.type = "a", 4; a.lambert.albedo.checker.odd.name = "o"
    def pbr_without_sss(m):
        a = m.materials["a"]; a.name, a.type = "a", 6
        for t in (a.pbr.albedo, a.pbr.roughness, a.pbr.metalness, a.pbr.normal_map):
            _vec3(t.constant.value, (0.5, 0.5, 0.5))
    def isotropic(m):
        a = m.materials["a"]; a.name, a.type = "a", 3; _vec3(a.isotropic.albedo.constant.value, (1, 1, 1))
    def displaced_without_map(m):
        a = m.materials["a"]; a.name, a.type = "a", 5
        t = m.objects.triangles.add(); t.material_name = "a"; t.operator = 1; t.displace.displacement_map = "d.png"

    for fn, why in [(missing_material, "material nope not found"), (missing_sphere_material, "material nope not found"),
                    (missing_texture, "texture x.png not found"), (lambert_without_albedo, "lambert material must have either albedo or spectral_albedo"),
                    (dielectric_without_refidx, "dielectric material must have either refidx or spectral_refidx"),
                    (light_without_emit, "diffuse light material must have either emit or spectral_emit"), (checker, "unknown texture type"),
                    (pbr_without_sss, "unknown texture type"), (isotropic, "isotropic"),
                    (displaced_without_map, "displacement map d.png not found")]:
        with pytest.raises(cuda.IzpiError) as e:
            scene_with(fn).to_scene()
        assert why in str(e.value), (fn.__name__, str(e.value))
    with pytest.raises(cuda.IzpiError):
        proto.ProtoScene(b"").to_c()  # to_scene() not called


def test_from_file_extension(tmp_path):
    with pytest.raises(ValueError):
        proto.ProtoScene.from_file(str(tmp_path / "scene.json"))
    p = tmp_path / "s.pbtxt"
    p.write_text('name: "t"')
    assert proto.ProtoScene.from_file(str(p)).name == "t"


# ---- on the device -----------------------------------------------------------------------------------
@pytest.mark.gpu
def test_ingested_scene_renders_like_the_restated_one():
    ctx = cuda.Context(0)
    sc = proto.ProtoScene.from_file(GOLDEN).to_scene(aspect_override=1.0)
    ctx.upload(cuda.HostScene(sc))
    a, rays_a = ctx.render(64, 64, 4, max_depth=50, sampler=sc.sampler, seed=11)
    ctx.upload(cuda.HostScene(scenes.spectral_pyramid(1.0)))
    b, rays_b = ctx.render(64, 64, 4, max_depth=50, sampler=cuda.SAMPLER_SPECTRAL, seed=11)
    assert a.tobytes() == b.tobytes() and rays_a == rays_b and a[..., :3].max() > 0
    ctx.close()


@pytest.mark.gpu
def test_displace_operator_goes_through_the_device_tessellator(oracle_mod):
    """Triangle.operator = DISPLACE (transport.go:633-646): one ApplyDisplacementMap call per triangle, results in place."""
    px = scenes.height_map(64, 32)
    m = ps.Scene()
    a = m.materials["a"]; a.name, a.type = "a", 5; _vec3(a.metal.albedo, (0.5, 0.5, 0.5))
    b = m.materials["b"]; b.name, b.type = "b", 5; _vec3(b.metal.albedo, (0.9, 0.5, 0.5))
    base = []
    for k, name in enumerate(["a", "b", "a"]):
        t = m.objects.triangles.add()
        v = np.array([[10.0 * k, 0, 0], [10.0 * k + 8, 0, 0], [10.0 * k, 8, 0]])
        uv = np.array([[0.1, 0.1], [0.4, 0.1], [0.1, 0.4]]) + 0.1 * k
        _vec3(t.vertex0, v[0]); _vec3(t.vertex1, v[1]); _vec3(t.vertex2, v[2])
        t.uv0.u, t.uv0.v = uv[0]; t.uv1.u, t.uv1.v = uv[1]; t.uv2.u, t.uv2.v = uv[2]
        t.material_name = name
        if k != 1:
            t.operator = 1
            t.displace.min, t.displace.max, t.displace.displacement_map = -0.5, 0.5, "h.exr"
        base.append(np.concatenate([np.float32(v).astype(np.float64).ravel(), np.float32(uv).astype(np.float64).ravel()]))
    m.displacement_maps["h.exr"].filename = "h.exr"
    ctx = cuda.Context(0)
    s = proto.ProtoScene(m.SerializeToString())
    assert s.image_filenames(displacement=True) == ["h.exr"]
    s.to_scene(displacement_maps={"h.exr": px}, displace_ctx=ctx)
    prims, mats, _, _ = spec_view(s.to_c())
    want0, _ = oracle_mod.apply_displacement(base[0][None], [0], px, -0.5, 0.5, per_triangle=True)
    want2, _ = oracle_mod.apply_displacement(base[2][None], [0], px, -0.5, 0.5, per_triangle=True)
    assert len(want0) > 1 and len(prims) == len(want0) + 1 + len(want2)
    assert prims["p"][:len(want0)].tobytes() == want0.tobytes()
    assert prims["p"][len(want0)].tobytes() == base[1].tobytes()
    assert prims["p"][len(want0) + 1:].tobytes() == want2.tobytes()
    assert mats[prims["material"][0]].v[0] == 0.5 and mats[prims["material"][len(want0)]].v[0] == f32(0.9)
    with pytest.raises(cuda.IzpiError) as e:
        proto.ProtoScene(m.SerializeToString()).to_scene(displacement_maps={"h.exr": px})
    assert "displace_ctx" in str(e.value)
    ctx.close()


# ---- the reference's own transport tests (internal/transport/transport_test.go), replayed ---------------------------------
def _pbr_constant_scene(spectral):
    """TestPBRMaterialTransformation's material (transport_test.go:12-75) on one triangle."""
    m = ps.Scene()
    m.colour_representation = 2 if spectral else 1
    a = m.materials["test_pbr"]
    a.name, a.type = "test_pbr", 6
    _vec3(a.pbr.albedo.constant.value, (1.0, 0.5, 0.2)); _vec3(a.pbr.roughness.constant.value, (0.5, 0.5, 0.5))
    _vec3(a.pbr.metalness.constant.value, (0.0, 0.0, 0.0)); _vec3(a.pbr.normal_map.constant.value, (0.5, 0.5, 1.0))
    _vec3(a.pbr.sss.constant.value, (0.0, 0.0, 0.0)); a.pbr.sss_radius = 0.0
    t = m.objects.triangles.add(); t.material_name = "test_pbr"
    _vec3(t.vertex1, (1, 0, 0)); _vec3(t.vertex2, (0, 1, 0))
    return m


def test_reference_pbr_material_transformation(oracle_mod):
    """transport_test.go:12-117: RGB -> plain PBR; SPECTRAL -> PBR with a spectral albedo whose value lies in [0, 1] at 380, 550
    and 650 nm (here: NewSpectralNeutral(luminance), transport.go:514-519)."""
    _, mats, texs, stex = spec_view(proto.ProtoScene(_pbr_constant_scene(False).SerializeToString()).to_scene().to_c())
    assert mats[0].type == S.MAT_PBR and mats[0].spectral_tex == -1
    assert texture_desc(texs, mats[0].tex) == ("const", (f32(1.0), f32(0.5), f32(0.2)))
    assert texture_desc(texs, mats[0].normal_tex) == ("const", (0.5, 0.5, 1.0))
    sc = proto.ProtoScene(_pbr_constant_scene(True).SerializeToString()).to_scene()
    _, mats, texs, stex = spec_view(sc.to_c())
    assert mats[0].type == S.MAT_PBR and mats[0].spectral_tex >= 0
    lum = 0.299 * f32(1.0) + 0.587 * f32(0.5) + 0.114 * f32(0.2)  # TestTextureToSpectralTexture's conversion (:119-138)
    osn = oracle_mod.OracleScene(sc)
    for lam in (380.0, 550.0, 650.0):
        v = osn.spectral_texture_value(mats[0].spectral_tex, lam)
        assert 0.0 <= v <= 1.0 and v == lum


def _library():
    import json
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lightsources.json")))


def _blackbody(k):
    """spectral.NewBlackbodySPD (spectral.go:275-320): Planck's law at 380..750 nm @ 5 nm, normalised to its maximum."""
    h, c, kb = 6.62607015e-34, 2.99792458e8, 1.380649e-23
    lam = (380.0 + 5.0 * np.arange(75)) * 1e-9
    v = (2.0 * h * c * c) / (lam ** 5 * (np.exp((h * c / kb) / (lam * k)) - 1.0))
    return v / v.max()


@pytest.mark.parametrize("name", sorted(_library()) + ["nonexistent_light_source"])
def test_reference_light_source_library_integration(oracle_mod, name):
    """transport_test.go:140-191, for EVERY key of the reference's library (lightsources.go:6-466) and for an unknown name (which
    falls back to CIE illuminant A, transport.go:483-490): the conversion succeeds stand-alone, the spectral texture is the
    library's table sample for sample (tests/golden/lightsources.json, generated from the reference by
    scripts/gen_lightsources.py), and its values at 400..700 nm are finite and non-negative."""
    m = ps.Scene()
    a = m.materials["l"]; a.name, a.type = "l", 2
    a.diffuselight.spectral_emit.from_light_source_library.light_source_name = name
    s = proto.ProtoScene(m.SerializeToString())
    s.to_scene()
    _, mats, _, stex = spec_view(s.to_c())
    d = spectral_desc(stex, mats[0].spectral_tex)
    lib = _library()
    want = lib.get(name, {"blackbody_k": 2856.0})
    want_v = np.asarray(want["values"]) if "values" in want else _blackbody(want["blackbody_k"])
    got_w, got_v = np.asarray(d[-2]), np.asarray(d[-1])
    assert np.array_equal(got_w, 380.0 + 5.0 * np.arange(75))
    if "values" in want:
        assert np.array_equal(got_v, want_v)
    else:
        np.testing.assert_allclose(got_v, want_v, rtol=1e-12)
    osn = oracle_mod.OracleScene(s)
    vals = [osn.spectral_texture_value(mats[0].spectral_tex, lam) for lam in (400.0, 500.0, 600.0, 700.0)]
    assert all(np.isfinite(x) and x >= 0.0 for x in vals)
    # a caller-supplied entry of the same name wins (a deployment with a patched library)
    s2 = proto.ProtoScene(m.SerializeToString())
    s2.to_scene(light_sources={name: (380.0 + 5.0 * np.arange(75), np.linspace(0.2, 1.0, 75))})
    d2 = spectral_desc(spec_view(s2.to_c())[3], 0)
    assert np.array_equal(np.asarray(d2[-1]), np.linspace(0.2, 1.0, 75))


def test_reference_spectral_texture_methods(oracle_mod):
    """transport_test.go:193-274: gaussian(1, 550, 40), neutral(0.73), tabulated {380,500,600,750} -> {0.1,0.5,0.8,0.3} evaluated
    at 550 nm are non-negative; the closed forms pin the values."""
    m = ps.Scene()
    a = m.materials["g"]; a.name, a.type = "g", 4
    a.lambert.spectral_albedo.gaussian.peak_value, a.lambert.spectral_albedo.gaussian.center_wavelength, a.lambert.spectral_albedo.gaussian.width = 1.0, 550, 40
    b = m.materials["n"]; b.name, b.type = "n", 4; b.lambert.spectral_albedo.neutral.reflectance = 0.73
    c = m.materials["t"]; c.name, c.type = "t", 4
    c.lambert.spectral_albedo.tabulated.wavelengths.extend([380, 500, 600, 750]); c.lambert.spectral_albedo.tabulated.values.extend([0.1, 0.5, 0.8, 0.3])
    s = proto.ProtoScene(m.SerializeToString()).to_scene()
    _, mats, _, _ = spec_view(s.to_c())
    by_name = {}
    osn = oracle_mod.OracleScene(s)
    for mat in mats:
        by_name[mat.spectral_tex] = osn.spectral_texture_value(mat.spectral_tex, 550.0)
    vals = sorted(by_name.values())
    # gaussian at its centre = peak (1.0); neutral = float32(0.73); tabulated = lerp(500 -> 0.5, 600 -> 0.8) at 550 in float32 inputs
    assert vals[2] == 1.0 and vals[1] == f32(0.73)
    assert abs(vals[0] - (f32(0.5) + 0.5 * (f32(0.8) - f32(0.5)))) < 1e-15
