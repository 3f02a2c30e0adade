import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lr():
    import lumillyrender_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def assets(lr):
    """Synthetic stand-ins for the absent models/ tree; small bunny + small IBL keep the oracle fast."""
    return lr.ensure_assets(ROOT, bunny_tris=20000, ibl_height=256)


@pytest.fixture(scope="session")
def gpu(lr):
    lr.init(0)
    return lr.device_info()


def load_scene(lr, name, resolution=None):
    return lr.Description(os.path.join(SCENES, name + ".toml"), asset_root=ROOT, resolution=resolution)


def make_params(lr_mod, cfg, **kw):
    from lumillyrender_b200.capi import LrRenderParams
    p = LrRenderParams()
    p.integrator = kw.get("integrator", cfg.integrator)
    p.spp_begin = kw.get("spp_begin", 0)
    p.spp_count = kw.get("spp", cfg.samples)
    p.depth = kw.get("depth", cfg.depth)
    p.depth_limit = kw.get("depth_limit", cfg.depth_limit)
    p.no_direct_emitter = kw.get("no_direct_emitter", cfg.no_direct_emitter)
    p.seed = kw.get("seed", 0)
    crop = kw.get("crop")
    if crop:
        p.crop_x, p.crop_y, p.crop_w, p.crop_h = crop
    p.splits = kw.get("splits", 0)
    return p


def mc_agreement(mean_a, sumsq_a, n_a, mean_b, sumsq_b, n_b):
    """Per pixel/channel 3-sigma test of two Monte Carlo means (SURVEY.md §8d acceptance stats).
    Returns (fraction of channels within 3 sigma, z-score of the image-mean difference, relMSE)."""
    var_a = np.maximum(sumsq_a / n_a - mean_a ** 2, 0.0) * n_a / max(n_a - 1, 1)
    var_b = np.maximum(sumsq_b / n_b - mean_b ** 2, 0.0) * n_b / max(n_b - 1, 1)
    se = np.sqrt(var_a / n_a + var_b / n_b)
    diff = np.abs(mean_a - mean_b)
    ok = diff <= 3.0 * se + 1e-6                    # 1e-6: channels with zero variance on both sides agree to rounding, not to 0
    img_se = np.sqrt((var_a / n_a + var_b / n_b).sum()) / mean_a.size
    z = abs(float(mean_a.mean()) - float(mean_b.mean())) / max(img_se, 1e-12)
    relmse = float(np.mean((mean_a - mean_b) ** 2 / (mean_b ** 2 + 1e-2)))
    return float(ok.mean()), z, relmse


def form_factor_scene(lr_mod, albedo=0.6, emission=(10.0, 8.0, 6.0), half=(10.0, 15.0), height=50.0, point=(37.0, 41.0), with_mesh=False, sphere_light=0.0):
    """A Lambert floor (y = 0) under a rectangular Lambert emitter of black albedo, parallel to it and centred above
    `point` — which lies where the reference's hard-coded checker (lambert.rs:66-90) is 1 — seen by a 1-degree ideal
    pinhole.  The reflected radiance at the point has a closed form: albedo * L_e * F with F the point-to-rectangle
    form factor, for pt (emission found by BSDF sampling) and pt-direct (light sampling) alike.
    sphere_light = r > 0 replaces the rectangle by a spherical emitter of radius r centred at the same height: the form
    factor of a sphere seen from straight below is (r / height)^2 (Sphere::sample, sphere.rs:79-84, samples the whole
    sphere; its far side fails the closest-hit distance match of scene.rs:127-132).
    with_mesh adds a cloud of 200 small black triangles beside the point (it neither shades nor lights it, but the scene
    then has a BVH, so the kernels that traverse one are the ones tested).
    Returns (Description, analytic RGB)."""
    import math
    from lumillyrender_b200 import capi
    C = capi.C
    a, b = half
    px, pz = point
    mats = (capi.LrMaterial * 3)()
    mats[2].type = capi.LR_MAT_LAMBERT                      # black, no emission: the mesh cloud
    mats[2].color[:] = [0.0, 0.0, 0.0]
    mats[0].type = capi.LR_MAT_LAMBERT
    mats[0].color[:] = [albedo] * 3
    mats[1].type = capi.LR_MAT_LAMBERT
    mats[1].color[:] = [0.0, 0.0, 0.0]                      # black: a path that reaches the light ends there
    mats[1].emission[:] = list(emission)
    n_mesh = 200 if with_mesh else 0
    T = (capi.LrTriangle * (4 + n_mesh))()
    big = 4000.0
    quads = [([(-big, 0, -big), (-big, 0, big), (big, 0, big)], 0), ([(-big, 0, -big), (big, 0, big), (big, 0, -big)], 0),
             # the light faces down: (p1 - p0) x (p2 - p0) = -y
             ([(px - a, height, pz - b), (px + a, height, pz - b), (px + a, height, pz + b)], 1),
             ([(px - a, height, pz - b), (px + a, height, pz + b), (px - a, height, pz + b)], 1)]
    if sphere_light > 0.0:
        # the two light triangles shrink to a speck far away, without emission (the count and the indices stay the same)
        quads[2] = ([(9000.0, 1.0, 9000.0), (9000.1, 1.0, 9000.0), (9000.0, 1.0, 9000.1)], 2)
        quads[3] = ([(9001.0, 1.0, 9000.0), (9001.1, 1.0, 9000.0), (9001.0, 1.0, 9000.1)], 2)
    for i, (v, m) in enumerate(quads):
        T[i].p0[:] = v[0]; T[i].p1[:] = v[1]; T[i].p2[:] = v[2]
        T[i].material = m; T[i].prim_id = i
    rng = np.random.RandomState(11)
    for i in range(n_mesh):
        c = np.array([px + 60.0, 15.0, pz]) + rng.uniform(-6, 6, 3)
        v = (c + rng.normal(0, 0.8, (3, 3))).astype(np.float32)
        T[4 + i].p0[:] = v[0]; T[4 + i].p1[:] = v[1]; T[4 + i].p2[:] = v[2]
        T[4 + i].material = 2; T[4 + i].prim_id = 4 + i
    lib = capi.load_library()
    mtx = (C.c_float * 16)()
    # The reference's look_at stores the basis as rows and M*v applies it untransposed: a view tilted inside
    # the yz-plane comes out with its tilt mirrored.  Aiming at the mirror image of the point (height +25 above the eye)
    # therefore looks DOWN at it; the tests check with the primary-hit probe that the centre ray lands on the point.
    lib.lr_matrix_look_at((C.c_float * 3)(px, 25.0, pz + 20.0), (C.c_float * 3)(px, 50.0, pz), (C.c_float * 3)(0, 1, 0), mtx)
    cam = capi.LrCamera()
    lib.lr_camera_ideal_pinhole(mtx, 1.0, 8, 8, C.byref(cam))
    S = (capi.LrSphere * (1 if sphere_light > 0.0 else 0))()
    if sphere_light > 0.0:
        S[0].center[:] = [px, height, pz]; S[0].radius = sphere_light; S[0].material = 1; S[0].prim_id = 4 + n_mesh
    d = lr_mod.Description.from_arrays(mats, T, S, cam)     # sky: black by default

    def corner(x, y, h):                                     # form factor of a rectangle x * y with a corner above the point
        return (x / math.hypot(x, h) * math.atan(y / math.hypot(x, h)) + y / math.hypot(y, h) * math.atan(x / math.hypot(y, h))) / (2 * math.pi)
    f = 4.0 * corner(a, b, height) if sphere_light <= 0.0 else (sphere_light / height) ** 2
    return d, np.array([albedo * e * f for e in emission])


def lens_cos4_case(lr_mod, kind, w=12, h=8):
    """An empty scene under a radiance-1 sky seen through a lens camera ("thin-lens" or "pinhole" = the reference's realistic
    pinhole) and the image it must produce: per pixel the mean of cos^4 over the pixel area and the aperture disc, cos
    between (aperture point - sensor point) and the optical axis.  Returns (Description, expected HxW image)."""
    from lumillyrender_b200 import capi
    C = capi.C
    lib = capi.load_library()
    F = lambda *v: (C.c_float * len(v))(*v)
    cam = capi.LrCamera()
    if kind == "thin-lens":
        m = (C.c_float * 16)()
        lib.lr_matrix_look_at(F(0, 0, 10), F(0, 0, 0), F(0, 1, 0), m)
        assert lib.lr_camera_thin_lens(m, 70.0, 60.0, 1.4, w, h, C.byref(cam)) == 0
    else:
        assert lib.lr_camera_pinhole(F(0, 0, 60), F(0, 0, 10), F(70.0, 70.0 * h / w), w, h, 9.0, C.byref(cam)) == 0
    sky = capi.LrSky()
    sky.type = capi.LR_SKY_UNIFORM
    sky.color[:] = [1.0, 1.0, 1.0]
    mats = (capi.LrMaterial * 1)()
    d = lr_mod.Description.from_arrays(mats, (capi.LrTriangle * 0)(), (capi.LrSphere * 0)(), cam, sky)
    fwd, right, up = (np.array(list(v), dtype=np.float64) for v in (cam.forward, cam.right, cam.up))
    pos, apc = np.array(list(cam.position), dtype=np.float64), np.array(list(cam.aperture_position), dtype=np.float64)
    sx, sy, ra = cam.sensor_size[0], cam.sensor_size[1], cam.aperture_radius
    assert ra > 1.0 and cam.aperture_sensor_distance > 0
    k = 6                                                       # midpoint nodes per pixel axis; 24 x 48 over the disc (equal areas)
    jit = (np.arange(k) + 0.5) / k
    rad = np.sqrt((np.arange(24) + 0.5) / 24) * ra
    ang = (np.arange(48) + 0.5) / 48 * 2 * np.pi
    disc = (rad[:, None, None] * (np.cos(ang)[None, :, None] * right + np.sin(ang)[None, :, None] * up)).reshape(-1, 3)
    expect = np.zeros((h, w))
    for y in range(h):
        for x in range(w):
            px = ((x + jit) / w - 0.5) * sx                      # camera.rs:64-81
            py = ((y + jit) / h - 0.5) * sy
            sensor = pos - px[:, None, None] * right + py[None, :, None] * up
            v = (apc + disc)[None, None] - sensor[:, :, None, :]
            cos = (v @ fwd) / np.linalg.norm(v, axis=-1)
            expect[y, x] = np.mean(cos ** 4)
    assert expect.min() < 0.6 and expect.max() > 0.9, "the vignetting must be visible in this set-up"
    return d, expect
