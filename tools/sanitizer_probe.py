"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel of the library once — both render organisations
(pt over a BVH, pt-direct over a BVH, flat-only), the probes, the AOV kernel, the film path, split reduction, the device BVH
builder — at sizes a sanitizer run finishes in a minute or two."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lumillyrender_b200 as lr

lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=20000, ibl_height=256)
for name, res, spp in (("sample", (64, 48), 2), ("welcome-2018", (64, 48), 2), ("new-cbox", (32, 32), 2), ("vr", (48, 24), 2)):
    d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=ROOT, resolution=res)
    s = d.scene()
    for org in ("pool", "persistent"):
        os.environ["LR_ORGANISATION"] = org
        img, sq, st = s.render(spp=spp, seed=1, splits=1, sumsq=True)
        img2, _, _ = s.render(spp=spp, seed=1, splits=2)
    os.environ.pop("LR_ORGANISATION")
    s.render(spp=1, seed=1, count=True)
    s.trace_primary()
    o = np.zeros((64, 3), np.float32) + np.array(list(d.camera().aperture_position), np.float32)
    dd = np.random.RandomState(1).normal(size=(64, 3)).astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    s.trace_rays(o, dd, normals=True)
    s.trace_rays(o, dd, normals=True, render_query=True)
    s.render_aov("normal", spp=2, seed=1)
    s.render_aov("depth", spp=2, seed=1)
    f = s.film(sumsq=True, seed=1, splits=1)
    f.render(1); f.render(1); f.read(sumsq=True); f.close()
    if d.desc.contents.n_nodes > 0:
        d.rebuild_bvh("device")
        s2 = d.scene()
        a, _, _ = s2.render(spp=spp, seed=1, splits=1)
        assert np.array_equal(a, img, equal_nan=True), name
        s2.close()
    s.close()
    print("ok", name, st["rays"], flush=True)
