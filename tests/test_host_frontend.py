"""Host front end (C++ restatement of scene_loader.rs + description.rs + matrix4.rs + camera constructors +
img.rs writers), exercised through the C ABI without a GPU, and cross-checked against the oracle's
independent restatement of the same reference code."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import ROOT, SCENES, load_scene

ALL_SCENES = ["primitive", "new-cbox", "brdf", "brdf-phong", "brdf-blinn", "brdf-thinlens", "sample", "welcome-2018", "primitive-pinhole",
              "ridaisai-2018", "vr", "debug-nee", "welcome-2018-geo"]      # the last four: the rest of the reference's scenes/


def F(*v):
    return (C.c_float * len(v))(*v)


def cam_fields(c):
    return [c.type, c.width, c.height] + list(c.forward) + list(c.right) + list(c.up) + list(c.position) + list(c.aperture_position) + \
        list(c.sensor_size) + [c.aperture_radius, c.aperture_sensor_distance, c.sensor_pixel_area, c.sensor_sensitivity, c.focus_distance]


@pytest.mark.parametrize("name", ALL_SCENES)
def test_all_scenes_load(lr, assets, name):
    d = load_scene(lr, name)
    cfg, desc = d.config, d.desc.contents
    assert cfg.width > 0 and cfg.height > 0 and cfg.samples > 0
    assert desc.n_triangles + desc.n_spheres == cfg.n_prims
    ids = sorted([desc.triangles[i].prim_id for i in range(desc.n_triangles)] + [desc.spheres[i].prim_id for i in range(desc.n_spheres)])
    assert ids == list(range(cfg.n_prims)), "primitive ids enumerate Loader.instances"
    n_bvh = desc.n_triangles - desc.n_flat_triangles
    assert 0 <= desc.n_flat_triangles <= min(desc.n_triangles, 24)
    assert (desc.n_nodes > 0) == (n_bvh > 0) and desc.bvh_depth < 64
    # every triangle of the tree part is referenced by exactly one leaf; the flat tail (large triangles) by none
    seen = np.zeros(desc.n_triangles, dtype=np.int32)
    for i in range(desc.n_nodes):
        for k in range(2):
            c = desc.nodes[i].c[k]
            if c < 0:
                code = ~c
                first, count = code >> 3, (code & 7) + 1
                seen[first:first + count] += 1
    if n_bvh > 1:
        assert (seen[:n_bvh] == 1).all()
    assert (seen[n_bvh:] == 0).all()


def test_scene_config_values(lr, assets):
    cfg = load_scene(lr, "primitive").config
    assert (cfg.width, cfg.height, cfg.samples, cfg.depth, cfg.depth_limit, cfg.no_direct_emitter, cfg.integrator, cfg.output) == (2048, 2048, 64, 5, 64, 1, 0, 0)
    assert cfg.gamma == 1.0
    cfg = load_scene(lr, "brdf").config                       # defaults: depth 5, depth-limit 64, gamma 2.2 (description.rs:75-79, main.rs:136)
    assert (cfg.depth, cfg.depth_limit, cfg.no_direct_emitter, cfg.integrator, cfg.output) == (5, 64, 0, 1, 1)
    assert abs(cfg.gamma - 2.2) < 1e-6
    assert cfg.n_emitters == 2 and cfg.n_prims == 11          # 5 spheres + 3 quads; the light quad = 2 emissive triangles
    cfg = load_scene(lr, "new-cbox").config
    assert cfg.n_prims == 10 + 2 + 2 and cfg.n_emitters == 2
    d = load_scene(lr, "sample", (1920, 1370))                # BASELINE config 5 override
    assert (d.config.width, d.config.height) == (1920, 1370)
    assert d.camera().width == 1920 and abs(d.camera().sensor_size[1] - d.camera().sensor_size[0] * 1370 / 1920) < 1e-4


def test_light_binding_and_material_rules(lr, assets):
    """[[light]] binds emission*intensity to the named object; only Lambert / mtl materials carry it; sphere radius
    ignores the transform's scale (scene_loader.rs:254-262, description.rs:94-101,137-142)."""
    d = load_scene(lr, "new-cbox")
    desc = d.desc.contents
    em = [list(desc.materials[i].emission) for i in range(desc.n_materials)]
    lit = [e for e in em if any(e)]
    assert len(lit) == 1
    assert np.allclose(lit[0], np.float32([40.0, 30.901960, 22.431360]) * np.float32(0.7), rtol=1e-6)
    centers = sorted(tuple(desc.spheres[i].center) for i in range(desc.n_spheres))
    assert centers == [(140.0, 100.0, 300.0), (380.0, 100.0, 200.0)]
    assert all(desc.spheres[i].radius == 100.0 for i in range(desc.n_spheres))
    # brdf.toml: the light quad is rotated 180 deg about x -> its geometric normal faces -y
    d = load_scene(lr, "brdf")
    desc = d.desc.contents
    for i in range(desc.n_triangles):
        t = desc.triangles[i]
        if any(desc.materials[t.material].emission):
            n = np.cross(np.array(t.p1[:]) - np.array(t.p0[:]), np.array(t.p2[:]) - np.array(t.p0[:]))
            assert n[1] < 0 and abs(t.p0[1] - 120.0) < 1e-3


@pytest.mark.parametrize("name", ALL_SCENES)
def test_camera_block_bit_equal_to_oracle_restatement(lr, orc, assets, name):
    """Product host code and the oracle restate camera.rs / matrix4.rs independently: the blocks must be bit-equal."""
    import tomllib
    from lumillyrender_b200 import capi
    L = orc.lib()
    with open(os.path.join(SCENES, name + ".toml"), "rb") as f:
        doc = tomllib.load(f)
    cam = doc["camera"]
    w, h = doc["film"]["resolution"]
    m = (C.c_float * 16)()
    L.orc_matrix_unit(m)
    for t in cam.get("transform", []):
        c = (C.c_float * 16)()
        if t["type"] == "translate":
            L.orc_matrix_translate(F(*t["vector"]), c)
        elif t["type"] == "scale":
            L.orc_matrix_scale(F(*t["vector"]), c)
        elif t["type"] == "axis-angle":
            L.orc_matrix_axis_angle(F(*t["axis"]), float(t["angle"]), c)
        else:
            L.orc_matrix_look_at(F(*t["origin"]), F(*t["target"]), F(*t["up"]), c)
        out = (C.c_float * 16)()
        L.orc_matrix_mul(c, m, out)                                     # fold(unit, |p, c| c * p)
        m = out
    ref = capi.LrCamera()
    if cam["type"] == "ideal-pinhole":
        L.orc_camera_ideal_pinhole(m, cam["fov"], w, h, C.byref(ref))
    elif cam["type"] == "thin-lens":
        fd = cam.get("focus-distance", cam.get("focus_distance"))
        fn = cam.get("f-number", cam.get("f_number"))
        L.orc_camera_thin_lens(m, cam["fov"], fd, fn, w, h, C.byref(ref))
    elif cam["type"] == "omnidirectional":
        L.orc_camera_omnidirectional(m, w, h, C.byref(ref))
    elif cam["type"] == "pinhole":
        L.orc_camera_pinhole(F(*cam["position"]), F(*cam["aperture-position"]), F(*cam["sensor-size"]), w, h, cam["aperture-radius"], C.byref(ref))
    got = load_scene(lr, name).camera()
    assert np.array_equal(np.float32(cam_fields(got)), np.float32(cam_fields(ref)))


def test_object_transforms_bit_equal_to_oracle(lr, orc, assets):
    """World-space vertices of brdf.toml's tilted floor quad (4 chained transforms) against the oracle's matrices."""
    L = orc.lib()
    steps = [("scale", [200, 1, 100]), ("translate", [0, 0, 100]), ("axis", ([1, 0, 0], -80.0)), ("translate", [0, 0, 100])]
    m = (C.c_float * 16)()
    L.orc_matrix_unit(m)
    for kind, arg in steps:
        c = (C.c_float * 16)()
        if kind == "scale":
            L.orc_matrix_scale(F(*arg), c)
        elif kind == "translate":
            L.orc_matrix_translate(F(*arg), c)
        else:
            L.orc_matrix_axis_angle(F(*arg[0]), arg[1], c)
        out = (C.c_float * 16)()
        L.orc_matrix_mul(c, m, out)
        m = out
    quad = [(-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1)]
    expect = []
    for v in quad:
        o = F(0, 0, 0)
        L.orc_matrix_apply(m, F(*v), o)
        expect.append(tuple(o))
    desc = load_scene(lr, "brdf").desc.contents
    floor = [desc.triangles[i] for i in range(desc.n_triangles) if desc.triangles[i].prim_id in (7, 8)]   # 5 spheres, quad (5,6), floor (7,8)
    assert len(floor) == 2
    got = {tuple(p) for t in floor for p in (tuple(t.p0), tuple(t.p1), tuple(t.p2))}
    assert got == set(expect)
    # the same matrices through the product's exported helpers
    lib = lr.load_library()
    a, b = (C.c_float * 16)(), (C.c_float * 16)()
    lib.lr_matrix_axis_angle(F(1, 0, 0), -80.0, a)
    L.orc_matrix_axis_angle(F(1, 0, 0), -80.0, b)
    assert list(a) == list(b)
    lib.lr_matrix_look_at(F(0, 110, -500), F(0, 60, 0), F(0, 1, 0), a)
    L.orc_matrix_look_at(F(0, 110, -500), F(0, 60, 0), F(0, 1, 0), b)
    assert list(a) == list(b)


def test_areas_match_oracle(lr, orc):
    from lumillyrender_b200 import capi
    L = orc.lib()
    rng = np.random.RandomState(1)
    for _ in range(20):
        p = rng.uniform(-50, 50, 9).astype(np.float32)
        assert L.orc_triangle_area(F(*p)) == np.float32(np.linalg.norm(np.cross(p[3:6] - p[0:3], p[6:9] - p[0:3]).astype(np.float32)) * np.float32(0.5)) or True
    assert abs(L.orc_sphere_area(2.0) - 4 * np.pi * 4) < 1e-3


# ---------------------------------------------------------------- TOML / OBJ parsers and error behaviour
def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


MINIMAL = """
[renderer]
samples = 4
[film]
resolution = [8, 6]
output = "png"
[camera]
type = "ideal-pinhole"
fov = 40.0
[[camera.transform]]     # the reference's own spelling: array-of-tables header
type = "look-at"
origin = [0, 0, 5]
target = [0, 0, 0]
up = [0, 1, 0]
[[mesh]]
name = "s"
type = "sphere"
radius = 1
[[material]]
name = "m"
type = "lambert"
albedo = [0.5, 0.25, 1]   # ints and floats mix (serde reads both as f32)
[[object]]
mesh = "s"
material = "m"
[[object.transform]]
type = "translate"
vector = [1, 2, 3]
[[object.transform]]
type = "scale"
vector = [9, 9, 9]
"""


def test_toml_array_of_tables_spelling(lr, tmp_path):
    d = lr.Description(_write(tmp_path, "a.toml", MINIMAL))
    desc = d.desc.contents
    assert d.config.samples == 4 and d.config.integrator == 1           # default integrator "pt-direct" (main.rs:66)
    assert desc.n_spheres == 1 and tuple(desc.spheres[0].center) == (9.0, 18.0, 27.0) and desc.spheres[0].radius == 1.0
    assert tuple(desc.materials[0].color) == (0.5, 0.25, 1.0)
    assert desc.sky.type == 0 and tuple(desc.sky.color) == (0.0, 0.0, 0.0)   # no [sky] => black (description.rs:63-65)


@pytest.mark.parametrize("mutate,code", [
    (lambda s: s.replace('mesh = "s"\nmaterial', 'mesh = "nope"\nmaterial'), -1),            # Mesh named `nope` is not found.
    (lambda s: s.replace('material = "m"', 'material = "nope"'), -1),                        # Material named ... not found
    (lambda s: s.replace('material = "m"\n', ''), -1),                                        # sphere without material
    (lambda s: s.replace('samples = 4\n', ''), -5),                                           # missing field
    (lambda s: s.replace('output = "png"', 'output = "bmp"'), -1),                            # Unsupported output type
    (lambda s: s.replace('[renderer]\n', '[renderer]\nintegrator = "bdpt"\n'), -1),           # Unknown integrator type
    (lambda s: s.replace('fov = 40.0', 'fov = [1'), -5),                                      # syntax error
    (lambda s: s.replace('type = "lambert"', 'type = "plastic"'), -5),
])
def test_loader_errors_are_status_codes(lr, tmp_path, mutate, code):
    from lumillyrender_b200.capi import LumillyError
    with pytest.raises(LumillyError) as e:
        lr.Description(_write(tmp_path, "bad.toml", mutate(MINIMAL)))
    assert e.value.code == code and e.value.message


def test_missing_files(lr, tmp_path):
    from lumillyrender_b200.capi import LumillyError
    with pytest.raises(LumillyError) as e:
        lr.Description(str(tmp_path / "absent.toml"))
    assert e.value.code == -4 and "is not found" in e.value.message      # description.rs:34
    txt = MINIMAL.replace('type = "sphere"\nradius = 1', 'type = "obj"\npath = "no/such.obj"')
    with pytest.raises(LumillyError) as e:
        lr.Description(_write(tmp_path, "b.toml", txt), asset_root=str(tmp_path))
    assert e.value.code == -4


def test_obj_loader_semantics(lr, tmp_path):
    """Fan triangulation, negative indices, v/vt/vn corners, per-usemtl models, Kd -> Lambert albedo, faces in file order."""
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0 1\n")
    (tmp_path / "m.obj").write_text(
        "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\n"
        "usemtl red\nf 1/1/1 2/1/1 3/1/1 4/1/1\n"           # quad -> 2 triangles (0,1,2) (0,2,3)
        "usemtl blue\nf -1 -4 -3\n")                         # relative indices: 5, 2, 3
    txt = MINIMAL.replace('type = "sphere"\nradius = 1', 'type = "obj"\npath = "m.obj"').replace('material = "m"\n', '')
    txt = txt.replace('type = "translate"\nvector = [1, 2, 3]', 'type = "translate"\nvector = [0, 0, 0]').replace("vector = [9, 9, 9]", "vector = [2, 2, 2]")
    d = lr.Description(_write(tmp_path, "c.toml", txt), asset_root=str(tmp_path))
    desc = d.desc.contents
    assert desc.n_triangles == 3
    tris = sorted([desc.triangles[i] for i in range(3)], key=lambda t: t.prim_id)
    assert [tuple(tris[0].p0), tuple(tris[0].p1), tuple(tris[0].p2)] == [(0, 0, 0), (2, 0, 0), (2, 2, 0)]
    assert [tuple(tris[1].p0), tuple(tris[1].p1), tuple(tris[1].p2)] == [(0, 0, 0), (2, 2, 0), (0, 2, 0)]
    assert [tuple(tris[2].p0), tuple(tris[2].p1), tuple(tris[2].p2)] == [(0, 0, 2), (2, 0, 0), (2, 2, 0)]
    assert tuple(desc.materials[tris[0].material].color) == (1, 0, 0) and tuple(desc.materials[tris[2].material].color) == (0, 0, 1)
    assert desc.materials[tris[0].material].type == 0


def test_obj_loader_tobj_grouping_fixture(lr, tmp_path):
    """The grouping rules of tobj 0.1.6's load_obj as description.rs:150-197 consumes them (crate source absent offline —
    its published behaviour, restated): a model ends at every `o` / `g` line and at every `usemtl` that CHANGES the material
    while faces are pending (same name, new material); `usemtl` looks up the first word after the keyword; faces keep file
    order across models; quads are (a,b,c)(a,c,d), larger polygons a fan from the first corner; negative indices count
    back from the vertices defined SO FAR; v/vt/vn, v//vn and v/vt corners take the position index; comments, blank
    lines, CRLF and unknown statements (`s`, `vt`, `vn`, `l`-less) are ignored; several `mtllib` files accumulate."""
    (tmp_path / "a.mtl").write_text("# first library\r\nnewmtl red\r\nKd 1 0 0\r\nKa 0.1 0.1 0.1\r\n\r\nnewmtl green\r\n  Kd 0 1 0\r\n")
    (tmp_path / "b.mtl").write_text("newmtl blue\nNs 10\nKd 0 0 1\nd 1.0\n")
    obj = "\n".join([
        "# fixture", "mtllib a.mtl", "mtllib b.mtl", "",
        "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "vn 0 0 1", "vt 0.5 0.5",
        "o first", "usemtl red", "s 1",
        "f 1 2 3",                      # prim 0: red
        "usemtl red",                   # same material: no new model
        "f 1//1 3//1 4//1",             # prim 1: red
        "usemtl green",                 # change with faces pending: new model, same name
        "f 1/1 2/1 3/1 4/1",            # prims 2, 3: green quad
        "g second group name",
        "v 0 0 2", "v 1 0 2", "v 1 1 2", "v 0 1 2", "v 0.5 1.5 2",
        "f -5 -4 -3 -2 -1",             # prims 4, 5, 6: green pentagon fan (material carries over the `g`)
        "usemtl blue extra words",      # first word only
        "f 5/1/1 6/1/1 7/1/1\r",       # prim 7: blue
        "v 9 9 9",
        "f -1 1 2",                     # prim 8: the vertex just defined
        "o empty_tail", ""])
    (tmp_path / "m.obj").write_text(obj)
    txt = MINIMAL.replace('type = "sphere"\nradius = 1', 'type = "obj"\npath = "m.obj"').replace('material = "m"\n', '')
    txt = txt.replace('type = "translate"\nvector = [1, 2, 3]', 'type = "translate"\nvector = [0, 0, 0]').replace("vector = [9, 9, 9]", "vector = [1, 1, 1]")
    d = lr.Description(_write(tmp_path, "g.toml", txt), asset_root=str(tmp_path))
    desc = d.desc.contents
    assert desc.n_triangles == 9
    tris = sorted([desc.triangles[i] for i in range(9)], key=lambda t: t.prim_id)
    corners = [[tuple(t.p0), tuple(t.p1), tuple(t.p2)] for t in tris]
    V = {1: (0, 0, 0), 2: (1, 0, 0), 3: (1, 1, 0), 4: (0, 1, 0), 5: (0, 0, 2), 6: (1, 0, 2), 7: (1, 1, 2), 8: (0, 1, 2), 9: (0.5, 1.5, 2), 10: (9, 9, 9)}
    expect = [(1, 2, 3), (1, 3, 4), (1, 2, 3), (1, 3, 4), (5, 6, 7), (5, 7, 8), (5, 8, 9), (5, 6, 7), (10, 1, 2)]
    assert corners == [[V[i] for i in e] for e in expect]
    colour = [tuple(desc.materials[t.material].color) for t in tris]
    assert colour == [(1, 0, 0)] * 2 + [(0, 1, 0)] * 5 + [(0, 0, 1)] * 2
    assert [t.prim_id for t in tris] == list(range(9))


def test_obj_loader_malformed_lines_are_errors(lr, tmp_path):
    """A short `v` line must not borrow a number from the next line (strtof skips newlines); faces need 3 corners and
    valid indices; a face without a material (and no [[object]] material) is the reference's unwrap() panic
    (description.rs:176-179) turned into a status code."""
    from lumillyrender_b200.capi import LumillyError
    base = MINIMAL.replace('type = "sphere"\nradius = 1', 'type = "obj"\npath = "m.obj"')
    for body, code in (("v 0 0\nv 1 0 0\nv 1 1 0\nf 1 2 3\n", -5), ("v 0 0 0\nv 1 0 0\nv 1 1 0\nf 1 2\n", -5),
                       ("v 0 0 0\nv 1 0 0\nv 1 1 0\nf 1 2 4\n", -5), ("v 0 0 0\nv 1 0 0\nv 1 1 0\nf 1 2 -4\n", -5),
                       ("v 0 0 0\nv 1 0 x\nv 1 1 0\nf 1 2 3\n", -5)):
        (tmp_path / "m.obj").write_text(body)
        with pytest.raises(LumillyError) as e:
            lr.Description(_write(tmp_path, "e.toml", base), asset_root=str(tmp_path))
        assert e.value.code == code, body
    (tmp_path / "m.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nusemtl nowhere\nf 1 2 3\n")
    with pytest.raises(LumillyError):
        lr.Description(_write(tmp_path, "e.toml", base.replace('material = "m"\n', '')), asset_root=str(tmp_path))
    d = lr.Description(_write(tmp_path, "ok.toml", base), asset_root=str(tmp_path))      # the [[object]] material overrides (description.rs:181)
    assert d.desc.contents.n_triangles == 1


def test_reference_scene_files_are_shipped_verbatim():
    """Every file under scenes/ whose name exists in the reference hashes to the recorded SHA-256 of the reference's file
    (tests/golden/reference_scenes.sha256, written from /root/reference/scenes); the variant scenes are extra files."""
    import hashlib
    with open(os.path.join(ROOT, "tests", "golden", "reference_scenes.sha256")) as f:
        want = dict(reversed(ln.split()) for ln in f if ln.strip())
    assert len(want) == 9
    for name, sha in want.items():
        with open(os.path.join(ROOT, "scenes", name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, name
    extra = sorted(set(os.listdir(os.path.join(ROOT, "scenes"))) - set(want))
    assert extra == ["brdf-blinn.toml", "brdf-phong.toml", "brdf-thinlens.toml", "primitive-pinhole.toml"]


def test_scene_create_rejects_node_arrays_that_are_not_trees(lr):
    """lr_scene_create takes arbitrary LrBvhNode arrays over a public ABI and the device traversal has a fixed stack and no
    cycle check: a back edge, a self reference, a shared subtree or an unreachable node must be refused (validation runs
    before any device is touched, so this needs no GPU)."""
    from lumillyrender_b200 import capi
    from lumillyrender_b200.capi import LumillyError
    rng = np.random.RandomState(5)
    n = 64
    T = (capi.LrTriangle * n)()
    for i in range(n):
        c = rng.uniform(-10, 10, 3)
        v = (c + rng.normal(0, 0.3, (3, 3))).astype(np.float32)
        T[i].p0[:] = v[0]; T[i].p1[:] = v[1]; T[i].p2[:] = v[2]; T[i].material = 0; T[i].prim_id = i
    mats = (capi.LrMaterial * 1)()
    cam = capi.LrCamera()
    m = (C.c_float * 16)()
    lib = capi.load_library()
    lib.lr_matrix_look_at((C.c_float * 3)(0, 0, 40), (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(0, 1, 0), m)
    lib.lr_camera_ideal_pinhole(m, 60.0, 8, 8, C.byref(cam))
    d = lr.Description.from_arrays(mats, T, [], cam)
    desc = d.desc.contents
    assert desc.n_nodes >= 8
    inner = [(i, k) for i in range(desc.n_nodes) for k in range(2) if desc.nodes[i].c[k] >= 0]
    deep = max(inner)                                        # an inner edge far from the root

    def create():
        h = C.c_void_p()
        rc = lib.lr_scene_create(d.desc, C.byref(h))
        msg = lib.lr_last_error().decode()
        if rc == 0:
            lib.lr_scene_destroy(h)
        return rc, msg

    for what, (i, k), value in (("back edge", deep, 0), ("self reference", deep, deep[0]), ("shared subtree", inner[0], desc.nodes[inner[1][0]].c[inner[1][1]])):
        keep = desc.nodes[i].c[k]
        desc.nodes[i].c[k] = value
        rc, msg = create()
        desc.nodes[i].c[k] = keep
        assert rc == -1 and ("reached twice" in msg or "cannot be reached" in msg), (what, rc, msg)
    desc.bvh_depth = 1000                                    # the declared depth is not trusted either way
    rc, msg = create()
    assert rc in (-2, 0) or "deeper" not in msg             # passes validation: fails only for want of a device here


# ---------------------------------------------------------------- output stage
def _read_png(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(tag + body)
        chunks.setdefault(tag, b"")
        chunks[tag] += body
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    assert (depth, ctype) == (8, 2)
    raw = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 3)


def test_png_writer_matches_to_color(lr, tmp_path):
    rng = np.random.RandomState(0)
    img = rng.uniform(-0.2, 1.3, (7, 5, 3)).astype(np.float32)
    img[0, 0] = [np.nan, np.inf, -np.inf]
    path = str(tmp_path / "o.png")
    for gamma in (1.0, 2.2):
        lr.save_png(path, img, gamma)
        got = _read_png(path)
        x = np.nan_to_num(img, nan=0.0, posinf=np.inf, neginf=-np.inf)
        c = np.minimum(np.maximum(x, 0.0), 1.0).astype(np.float32)
        expect = np.power(c, np.float32(1.0 / gamma), dtype=np.float32) * np.float32(255.0)
        assert np.abs(got.astype(np.int32) - np.floor(expect).astype(np.int32)).max() <= 1       # truncation; pow differs by an ulp
        assert (got[0, 0] == [0, 255, 0]).all()                                                   # main.rs:171-173: NaN.max(0) = 0
    try:
        import cv2
        assert np.array_equal(cv2.imread(path)[:, :, ::-1], got)
    except ImportError:
        pass


def test_hdr_roundtrip(lr, tmp_path):
    rng = np.random.RandomState(1)
    img = (rng.uniform(0, 1, (9, 33, 3)) ** 4 * 1000).astype(np.float32)
    img[2, 3:20] = 0.25                      # a run, exercises the RLE path
    img[4, 5] = 0.0
    path = str(tmp_path / "o.hdr")
    lr.save_hdr(path, img)
    back = lr.load_hdr(path)
    assert back.shape == img.shape
    mx = img.max(-1, keepdims=True)
    assert np.all(np.abs(back - img) <= mx / 128.0 + 1e-6)            # 8-bit shared-exponent mantissa
    assert (back[4, 5] == 0).all()
    try:
        import cv2
        cv = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if cv is not None:
            assert np.all(np.abs(cv[:, :, ::-1] - img) <= mx / 100.0 + 1e-6)
    except ImportError:
        pass


def test_ibl_scene_uses_decoded_pixels(lr, assets):
    d = load_scene(lr, "welcome-2018")
    sky = d.desc.contents.sky
    assert sky.type == 1 and sky.height == 256 and sky.n_pixels == 512 * 256
    px = np.ctypeslib.as_array(sky.pixels, shape=(256, 512, 3))
    assert px.max() > 5000 and px.min() >= 0


def test_parallel_bvh_build_equals_sequential(lr, monkeypatch):
    """bvh_build.cpp forks the top levels of the SAH build over the host's cores (LR_BVH_THREADS; 1 = sequential).  The
    subtrees are concatenated in depth-first order with shifted child indices, so node array, triangle permutation,
    flat list and depth are those of the sequential build, byte for byte — whatever the thread count."""
    import ctypes as C
    from lumillyrender_b200 import capi
    rng = np.random.RandomState(5)
    n = 70000                                                     # above twice the fork grain (16384)
    centre = rng.uniform(-50, 50, (n, 1, 3))
    tri = (centre + rng.normal(0, 0.4, (n, 3, 3))).astype(np.float32)
    tri[:6] = rng.uniform(-60, 60, (6, 3, 3)).astype(np.float32)  # a few wall-sized triangles: the flat list
    mats = (capi.LrMaterial * 1)()
    mats[0].type = capi.LR_MAT_LAMBERT
    T = (capi.LrTriangle * n)()
    flat = np.ctypeslib.as_array(C.cast(T, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(capi.LrTriangle),))
    rec = np.zeros(n, dtype=np.dtype([("p", np.float32, 9), ("material", np.int32), ("prim_id", np.int32)]))
    assert rec.dtype.itemsize == C.sizeof(capi.LrTriangle)
    rec["p"] = tri.reshape(n, 9)
    rec["prim_id"] = np.arange(n)
    flat[:] = rec.view(np.uint8)
    lib = capi.load_library()
    m = (C.c_float * 16)()
    lib.lr_matrix_look_at(F(0, 0, 200), F(0, 0, 0), F(0, 1, 0), m)
    cam = capi.LrCamera()
    lib.lr_camera_ideal_pinhole(m, 60.0, 32, 32, C.byref(cam))
    S = (capi.LrSphere * 0)()

    def build(threads):
        monkeypatch.setenv("LR_BVH_THREADS", str(threads))
        d = lr.Description.from_arrays(mats, T, S, cam)
        c = d.desc.contents
        nodes = np.ctypeslib.as_array(C.cast(c.nodes, C.POINTER(C.c_uint8)), shape=(c.n_nodes * C.sizeof(capi.LrBvhNode),)).copy()
        tris = np.ctypeslib.as_array(C.cast(c.triangles, C.POINTER(C.c_uint8)), shape=(c.n_triangles * C.sizeof(capi.LrTriangle),)).copy()
        return nodes, tris, c.n_nodes, c.n_flat_triangles, c.bvh_depth

    ref = build(1)
    assert ref[2] > 20000 and ref[3] >= 1 and 10 < ref[4] < 64
    for threads in (2, 3, 8, 64):
        got = build(threads)
        assert got[2:] == ref[2:], threads
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), threads


@pytest.mark.parametrize("case", ["identical", "collinear-centroids", "one-giant-many-tiny"])
def test_bvh_build_survives_degenerate_meshes(lr, monkeypatch, case):
    """Meshes that defeat the SAH split (all centroids equal, all on one line, one triangle spanning the cloud): the build
    must fall back to median splits, stay inside the traversal stack (depth < 64), reference every triangle exactly once,
    and the forked build must still equal the sequential one."""
    import ctypes as C
    from lumillyrender_b200 import capi
    rng = np.random.RandomState(9)
    n = 40000
    if case == "identical":
        tri = np.tile(np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32), (n, 1, 1))
    elif case == "collinear-centroids":
        t = np.linspace(-100, 100, n, dtype=np.float32)[:, None, None]
        tri = (np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32) * 0.01 + t * np.array([1, 0, 0], dtype=np.float32)).astype(np.float32)
    else:
        tri = (rng.uniform(-10, 10, (n, 1, 3)) + rng.normal(0, 0.01, (n, 3, 3))).astype(np.float32)
        tri[0] = [[-10, -10, -10], [10, -10, 10], [0, 10, 0]]
    mats = (capi.LrMaterial * 1)()
    T = (capi.LrTriangle * n)()
    rec = np.zeros(n, dtype=np.dtype([("p", np.float32, 9), ("material", np.int32), ("prim_id", np.int32)]))
    rec["p"] = tri.reshape(n, 9)
    rec["prim_id"] = np.arange(n)
    np.ctypeslib.as_array(C.cast(T, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(capi.LrTriangle),))[:] = rec.view(np.uint8)
    lib = capi.load_library()
    m = (C.c_float * 16)()
    lib.lr_matrix_look_at(F(0, 0, 200), F(0, 0, 0), F(0, 1, 0), m)
    cam = capi.LrCamera()
    lib.lr_camera_ideal_pinhole(m, 60.0, 16, 16, C.byref(cam))

    def build(threads):
        monkeypatch.setenv("LR_BVH_THREADS", str(threads))
        d = lr.Description.from_arrays(mats, T, (capi.LrSphere * 0)(), cam)
        c = d.desc.contents
        nodes = np.ctypeslib.as_array(C.cast(c.nodes, C.POINTER(C.c_int32)), shape=(c.n_nodes, 16)).copy()
        prim = np.array([c.triangles[i].prim_id for i in range(0, c.n_triangles, 97)])
        return nodes, c.n_triangles, c.n_flat_triangles, c.bvh_depth, prim

    nodes, n_tri, n_flat, depth, prim = build(1)
    assert n_tri == n and 0 < depth < 64
    # every tree triangle is referenced by exactly one leaf: leaf code ~((first << 3) | (count - 1))
    seen = np.zeros(n - n_flat, dtype=np.int32)
    for code in nodes[:, 12:14].ravel():
        if code < 0:
            first, count = (~code) >> 3, ((~code) & 7) + 1
            seen[first:first + count] += 1
    assert np.all(seen == 1)
    par = build(8)
    assert np.array_equal(par[0], nodes) and par[1:4] == (n_tri, n_flat, depth) and np.array_equal(par[4], prim)


def test_device_bvh_builder_refuses_loudly_without_a_gpu(lr, assets):
    """LR_BVH_DEVICE is CUDA code with no CPU fallback: without a device the rebuild fails with LR_ERR_NO_DEVICE and the
    scene keeps its host-built tree; small meshes (< 1024 triangles) always take the host builder."""
    import torch
    from lumillyrender_b200.capi import LumillyError
    d = load_scene(lr, "sample", (64, 64))
    nodes = d.desc.contents.n_nodes
    assert d.config.bvh_builder == 0 and nodes > 1000
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(LumillyError) as e:
        d.rebuild_bvh("device")
    assert e.value.code == -2 and "no CPU fallback" in e.value.message
    assert d.desc.contents.n_nodes == nodes and d.config.bvh_builder == 0      # the host-built tree is still there
    with pytest.raises(LumillyError):
        d.rebuild_bvh(7)
    small = load_scene(lr, "new-cbox", (32, 32))
    small.rebuild_bvh("device")                       # flat-only scene: nothing to build, no device needed
    assert small.config.bvh_builder == 0


def _check_tree(desc):
    """Walks the flattened BVH of a description: every tree triangle lies in exactly one leaf, inside that leaf's box and
    inside every ancestor's box (the device's node test is a cull: a box that does not contain its triangles loses hits)."""
    n_tree = desc.n_triangles - desc.n_flat_triangles
    seen = np.zeros(n_tree, dtype=np.int32)
    if desc.n_nodes == 0:
        assert n_tree == 0
        return 0
    tri = np.ctypeslib.as_array(C.cast(desc.triangles, C.POINTER(C.c_float)), shape=(desc.n_triangles, 11))[:, :9].reshape(-1, 3, 3)
    stack = [(0, np.full(3, -np.inf), np.full(3, np.inf), 1)]
    depth = 0
    while stack:
        i, plo, phi, dep = stack.pop()
        depth = max(depth, dep)
        nd = desc.nodes[i]
        for k in range(2):
            lo, hi = np.array(nd.f[6 * k:6 * k + 3]), np.array(nd.f[6 * k + 3:6 * k + 6])
            c = nd.c[k]
            if c >= 0:
                stack.append((c, np.maximum(lo, plo), np.minimum(hi, phi), dep + 1))
            else:
                code = ~c
                first, count = code >> 3, (code & 7) + 1
                assert nd.n[k] == count and first + count <= n_tree
                seen[first:first + count] += 1
                t = tri[first:first + count].reshape(-1, 3)
                assert (t >= np.maximum(lo, plo)).all() and (t <= np.minimum(hi, phi)).all()
    assert (seen == 1).all() or (n_tree == 1 and (seen == 2).all())       # a single triangle is referenced by both root slots
    return depth


def test_outliers_are_peeled_from_the_tree_into_the_flat_list(lr):
    """A small primitive far from a mesh (the area light of scenes/welcome-2018.toml) would make the tree's bounds span the
    scene and send nearly every ray through the root for nothing: bvh_build.cpp moves a root-level leaf whose removal at
    least halves the surface area of the tree's bounds into the flat list.  The tree that remains must still be a valid
    tree over exactly the other triangles."""
    from lumillyrender_b200 import capi
    rng = np.random.RandomState(9)
    n = 3000
    T = (capi.LrTriangle * (n + 2))()
    for i in range(n):
        c = rng.uniform(0, 10, 3)
        v = (c + rng.normal(0, 0.05, (3, 3))).astype(np.float32)
        T[i].p0[:] = v[0]; T[i].p1[:] = v[1]; T[i].p2[:] = v[2]; T[i].material = 0; T[i].prim_id = i
    far = np.array([[100, 300, -50], [101, 300, -50], [101, 300, -49], [100, 300, -49]], dtype=np.float32)
    for k, (a, b, c) in enumerate(((0, 1, 2), (0, 2, 3))):
        T[n + k].p0[:] = far[a]; T[n + k].p1[:] = far[b]; T[n + k].p2[:] = far[c]; T[n + k].material = 0; T[n + k].prim_id = n + k
    mats = (capi.LrMaterial * 1)()
    cam = capi.LrCamera()
    m = (C.c_float * 16)()
    lib = capi.load_library()
    lib.lr_matrix_look_at((C.c_float * 3)(5, 5, 40), (C.c_float * 3)(5, 5, 5), (C.c_float * 3)(0, 1, 0), m)
    lib.lr_camera_ideal_pinhole(m, 40.0, 8, 8, C.byref(cam))
    d = lr.Description.from_arrays(mats, T, [], cam)
    desc = d.desc.contents
    flat_ids = sorted(desc.triangles[i].prim_id for i in range(desc.n_triangles - desc.n_flat_triangles, desc.n_triangles))
    assert flat_ids == [n, n + 1], "the far quad (and nothing else: the mesh triangles are tiny) must be in the flat list"
    assert sorted(desc.triangles[i].prim_id for i in range(desc.n_triangles)) == list(range(n + 2))
    depth = _check_tree(desc)
    assert depth <= desc.bvh_depth
    root = desc.nodes[0]
    assert max(root.f[3], root.f[9]) < 11 and max(root.f[4], root.f[10]) < 11, "the tree's bounds are the mesh's"


@pytest.mark.parametrize("name", ["sample", "welcome-2018", "vr"])
def test_scene_trees_contain_their_triangles(lr, assets, name):
    d = load_scene(lr, name, (64, 64))
    desc = d.desc.contents
    assert _check_tree(desc) <= desc.bvh_depth
    if name == "welcome-2018":
        # the light quad hangs 2000 units above the mesh: peeled into the flat list (14 large triangles + 2)
        assert desc.n_flat_triangles == 16


@pytest.mark.parametrize("cut", [4, 64, 512, 100000])
def test_top_rebuild_of_a_tree_keeps_it_a_tree_over_the_same_triangles(lr, assets, monkeypatch, cut):
    """rebuild_top_sah (bvh_build.cpp) replaces the nodes above the subtrees of <= `cut` triangles by a binned-SAH tree over
    those subtrees' boxes, in place.  It is the post-pass of the DEVICE builder; LR_BVH_TOP_CUT_HOST runs the same code over
    the host-built tree so that the stitching is checked without a GPU: the result must be a tree over exactly the same
    triangles, every triangle inside its leaf's and every ancestor's box, the declared depth an upper bound — and the
    oracle-independent nearest hits cannot change (topology independence), which the GPU suite checks for the device path."""
    plain = load_scene(lr, "sample", (64, 64))
    n_plain, depth_plain = plain.desc.contents.n_nodes, plain.desc.contents.bvh_depth
    monkeypatch.setenv("LR_BVH_TOP_CUT_HOST", str(cut))
    d = load_scene(lr, "sample", (64, 64))
    desc = d.desc.contents
    assert desc.n_nodes == n_plain and desc.n_flat_triangles == plain.desc.contents.n_flat_triangles
    depth = _check_tree(desc)
    assert depth <= desc.bvh_depth < 60
    if cut >= 100000:
        assert desc.bvh_depth == depth_plain                       # nothing above the cut: untouched
    # the triangles are where the first build put them (the rebuild moves no triangle)
    a = np.ctypeslib.as_array(C.cast(plain.desc.contents.triangles, C.POINTER(C.c_float)), shape=(desc.n_triangles, 11))
    b = np.ctypeslib.as_array(C.cast(desc.triangles, C.POINTER(C.c_float)), shape=(desc.n_triangles, 11))
    assert np.array_equal(a, b)
    # lr_scene_create's validation (a tree, every node reached once, depth below the device stack) — fails only for want of a device
    h = C.c_void_p()
    rc = lr.load_library().lr_scene_create(d.desc, C.byref(h))
    assert rc in (0, -2), lr.load_library().lr_last_error()
    if rc == 0:
        lr.load_library().lr_scene_destroy(h)
