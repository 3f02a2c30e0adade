"""Deterministic stand-ins for the reference's git-ignored assets (reference .gitignore:14,16).

Every `.obj/.mtl/.hdr` the scene files name is absent offline, so the same synthetic files
feed the oracle and the GPU path:
  models/simple/cbox.obj (+.mtl)            Cornell-box walls, 10 triangles, white/red/green Kd
  models/simple/quad.obj (+.mtl)            2x2 quad in the xz-plane at y=0, +y normal (inferred from
                                            reference scenes/brdf.toml:63-98, welcome-2018.toml:52-89)
  models/simple/cbox_luminaire.obj (+.mtl)  130x105 quad just under the ceiling, normal -y
  models/bunny/bunny.obj                    closed displaced-sphere mesh of the named triangle count
  models/ibl/14-Hamarikyu_Bridge_B_3k.hdr   synthetic 2H x H equirect: gradient + 1e4-radiance sun disc
"""
import os

import numpy as np

CBOX_OBJ = """# synthetic Cornell box walls (classic measurements), stand-in for models/simple/cbox.obj
mtllib cbox.mtl
v 552.8 0.0 0.0
v 0.0 0.0 0.0
v 0.0 0.0 559.2
v 549.6 0.0 559.2
v 556.0 548.8 0.0
v 556.0 548.8 559.2
v 0.0 548.8 559.2
v 0.0 548.8 0.0
o floor
usemtl white
f 1 2 3 4
o ceiling
usemtl white
f 5 6 7 8
o back_wall
usemtl white
f 4 3 7 6
o right_wall
usemtl green
f 3 2 8 7
o left_wall
usemtl red
f 1 4 6 5
"""
CBOX_MTL = """newmtl white
Kd 0.740063 0.742313 0.733934
newmtl red
Kd 0.366046 0.0371827 0.0416385
newmtl green
Kd 0.162928 0.408903 0.0833759
"""
QUAD_OBJ = """# unit quad: 2x2 in the xz-plane at y = 0, +y normal
mtllib quad.mtl
v -1.0 0.0 -1.0
v -1.0 0.0 1.0
v 1.0 0.0 1.0
v 1.0 0.0 -1.0
o quad
usemtl grey
f 1 2 3 4
"""
QUAD_MTL = """newmtl grey
Kd 0.8 0.8 0.8
"""
LUMINAIRE_OBJ = """# Cornell box light: 130 x 105 quad just under the ceiling, normal -y
mtllib cbox_luminaire.mtl
v 343.0 548.7 227.0
v 343.0 548.7 332.0
v 213.0 548.7 332.0
v 213.0 548.7 227.0
o luminaire
usemtl light
f 1 2 3 4
"""
LUMINAIRE_MTL = """newmtl light
Kd 0.78 0.78 0.78
"""


def _write_if_changed(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path):
        with open(path) as f:
            if f.read() == text:
                return
    with open(path, "w") as f:
        f.write(text)


def blob_mesh(n_tris, seed=1):
    """Closed lumpy 'bunny-sized' mesh: a lat-long grid on a displaced sphere, ~n_tris triangles.
    Returns (vertices float32 [nv,3], faces int32 [nf,3]); vertices roughly inside [-1,1] x [0,1.6] x [-1,1]."""
    rows = max(4, int(round(np.sqrt(n_tris / 4.0))))
    cols = max(8, int(round(n_tris / (2.0 * rows))))
    rng = np.random.RandomState(seed)
    theta = (np.arange(rows + 1, dtype=np.float64) / rows) * np.pi           # 0 .. pi
    phi = (np.arange(cols, dtype=np.float64) / cols) * 2.0 * np.pi
    T, P = np.meshgrid(theta, phi, indexing="ij")
    r = np.ones_like(T)
    for _ in range(12):                                                       # low-frequency lumps
        k_t, k_p = rng.randint(1, 6), rng.randint(0, 5)
        amp = 0.25 / (k_t + k_p)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, 2)
        r += amp * np.sin(k_t * T + ph1) * np.cos(k_p * P + ph2) * np.sin(T)
    r += 0.004 * np.sin(37.0 * T) * np.sin(41.0 * P) * np.sin(T)              # fine ripples
    x = r * np.sin(T) * np.cos(P)
    y = r * np.cos(T)
    z = r * np.sin(T) * np.sin(P)
    v = np.stack([0.72 * x - 0.2, 0.72 * y + 0.85, 0.72 * z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(rows)[:, None]
    j = np.arange(cols)[None, :]
    a = i * cols + j
    b = i * cols + (j + 1) % cols
    c = (i + 1) * cols + j
    d = (i + 1) * cols + (j + 1) % cols
    f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([b, c, d], -1).reshape(-1, 3)], 0).astype(np.int32)
    return v, f


def write_obj(path, v, f, header="synthetic mesh"):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as out:
        out.write("# %s: %d vertices, %d triangles\no mesh\n" % (header, len(v), len(f)))
        out.write("".join("v %.7g %.7g %.7g\n" % (p[0], p[1], p[2]) for p in v.tolist()))
        out.write("".join("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1) for t in f.tolist()))


def synth_ibl(height=1600):
    """Equirect RGB fp32 [H, 2H, 3]: sky/ground gradient plus one small, very bright sun disc."""
    w = 2 * height
    vv = (np.arange(height, dtype=np.float32) + 0.5) / height                 # 0 top .. 1 bottom
    uu = (np.arange(w, dtype=np.float32) + 0.5) / w
    V, U = np.meshgrid(vv, uu, indexing="ij")
    sky = np.stack([0.35 + 0.4 * (1 - V), 0.45 + 0.4 * (1 - V), 0.6 + 0.5 * (1 - V)], -1)
    ground = np.stack([0.25 + 0.1 * np.sin(6.2831853 * U), 0.22 + 0.05 * np.cos(12.566 * U), 0.18 + 0 * U], -1)
    img = np.where((V < 0.5)[..., None], sky, ground).astype(np.float32)
    theta, phi = V * np.pi, U * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    sun = np.array([0.45, 0.75, -0.48], dtype=np.float32)
    sun /= np.linalg.norm(sun)
    img[(d @ sun) > np.cos(np.radians(1.5))] = np.array([1.0e4, 0.9e4, 0.8e4], dtype=np.float32)
    return img


def ensure_assets(root, bunny_tris=144046, ibl_height=1600, need_bunny=True, need_ibl=True):
    """Creates the synthetic assets under `root`/models if absent (idempotent). Returns `root`.
    need_bunny also covers the dragon stand-in of scenes/vr.toml."""
    root = os.path.abspath(root)
    simple = os.path.join(root, "models", "simple")
    _write_if_changed(os.path.join(simple, "cbox.obj"), CBOX_OBJ)
    _write_if_changed(os.path.join(simple, "cbox.mtl"), CBOX_MTL)
    _write_if_changed(os.path.join(simple, "quad.obj"), QUAD_OBJ)
    _write_if_changed(os.path.join(simple, "quad.mtl"), QUAD_MTL)
    _write_if_changed(os.path.join(simple, "cbox_luminaire.obj"), LUMINAIRE_OBJ)
    _write_if_changed(os.path.join(simple, "cbox_luminaire.mtl"), LUMINAIRE_MTL)
    if need_bunny:
        bunny = os.path.join(root, "models", "bunny", "bunny.obj")
        tag = bunny + ".tris"
        have = None
        if os.path.exists(bunny) and os.path.exists(tag):
            with open(tag) as f:
                have = f.read().strip()
        if have != str(bunny_tris):
            v, f = blob_mesh(bunny_tris)
            write_obj(bunny, v, f, "procedural stand-in for the McGuire-archive bunny")
            with open(tag, "w") as fh:
                fh.write(str(bunny_tris))
        # scenes/vr.toml: models/stanford_dragon/dragon.obj (absent) -> a second procedural closed mesh, half the bunny's size
        dragon = os.path.join(root, "models", "stanford_dragon", "dragon.obj")
        tag = dragon + ".tris"
        have = None
        if os.path.exists(dragon) and os.path.exists(tag):
            with open(tag) as f:
                have = f.read().strip()
        n_dragon = max(2000, bunny_tris // 2)
        if have != str(n_dragon):
            v, f = blob_mesh(n_dragon, seed=7)
            write_obj(dragon, v * np.float32(5.0), f, "procedural stand-in for the Stanford dragon")
            with open(tag, "w") as fh:
                fh.write(str(n_dragon))
    if need_ibl:
        # welcome-2018.toml and ridaisai-2018.toml name two different (absent) .hdr files; both get the synthetic sky
        for name in ("14-Hamarikyu_Bridge_B_3k.hdr", "PaperMill_Ruins_E.hdr"):
            hdr = os.path.join(root, "models", "ibl", name)
            tag = hdr + ".height"
            have = None
            if os.path.exists(hdr) and os.path.exists(tag):
                with open(tag) as f:
                    have = f.read().strip()
            if have != str(ibl_height):
                from .renderer import save_hdr
                os.makedirs(os.path.dirname(hdr), exist_ok=True)
                save_hdr(hdr, synth_ibl(ibl_height))
                with open(tag, "w") as fh:
                    fh.write(str(ibl_height))
    return root
