"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden files and cannot be run here, so these fixtures are outputs of the
oracle — the line-by-line restatement of the reference's algorithm — frozen at the commit that introduced them.
They serve two purposes: (1) they pin the ORACLE (tests/test_golden.py re-runs it on CPU and demands bit equality,
so an accidental change of the checker is caught), (2) they give the CUDA path fixed numbers to match on the GPU box.
Only scenes whose arithmetic is bit-defined on every platform are used (no powf / expf / acosf / atan2f: Lambert and
GGX materials, uniform sky, ideal-pinhole and thin-lens cameras with the specified sincos).

Fixtures per scene: nearest primary hit (primitive index, t) at two sensor jitters, and the replay render
(per-pixel SUM of `spp` samples, shared counter-based RNG, seed 11) with its ray count.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = {   # scene -> (resolution, spp)
    "primitive": ((96, 96), 8),
    "new-cbox": ((64, 64), 8),
    "brdf": ((96, 54), 8),
    "brdf-thinlens": ((96, 54), 4),
    "sample": ((64, 64), 4),
    "welcome-2018-geo": ((72, 52), 4),
    "debug-nee": ((64, 64), 8),
}
JITTERS = [(0.5, 0.5, 0.5, 0.5), (0.137, 0.859, 0.301, 0.644)]
ASSETS = dict(bunny_tris=20000, ibl_height=256)      # what tests/conftest.py::assets creates


def compute(lr, orc, name):
    from conftest import load_scene, make_params
    res, spp = GOLDEN[name]
    d = load_scene(lr, name, res)
    o = orc.OracleScene(d.desc, keepalive=d)
    out = {"spp": np.int32(spp), "resolution": np.int32(res)}
    for k, j in enumerate(JITTERS):
        prim, t = o.trace_primary(*j, traversal=0)
        out["prim%d" % k] = prim
        out["t%d" % k] = t
    s, sq, st = o.render(make_params(lr, d.config, spp=spp, seed=11), traversal=0, rng_mode=0, math_mode=1)
    out["sum"] = s
    out["rays"] = np.int64(st["rays"])
    return out


if __name__ == "__main__":
    import lumillyrender_b200 as lr
    from oracle import oracle_py as orc
    lr.load_library()
    orc.lib()
    lr.ensure_assets(ROOT, **ASSETS)
    here = os.path.dirname(os.path.abspath(__file__))
    for name in GOLDEN:
        g = compute(lr, orc, name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **g)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in g.items()})
