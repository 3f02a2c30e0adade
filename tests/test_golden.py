"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle must reproduce its frozen outputs bit for bit (pins the checker).
GPU: the CUDA path, through the C ABI, must match the same fixtures: primary-hit primitive index on >= 99.99 % of
pixels with bit-equal t, replay render with the identical ray count and per-pixel means within rtol 1e-4 on
>= 99.9 % of pixels (the residue is fp32 rounding of the iterative throughput form vs the reference's recursion)."""
import os
import sys

import numpy as np
import pytest

from conftest import load_scene

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg   # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", list(mg.GOLDEN))
def test_oracle_reproduces_golden(lr, orc, assets, name):
    g = _load(name)
    now = mg.compute(lr, orc, name)
    for k in g.files:
        assert np.array_equal(np.asarray(now[k]), g[k]), "oracle output `%s` of %s changed" % (k, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.GOLDEN))
def test_cuda_matches_golden(lr, assets, gpu, name):
    g = _load(name)
    res, spp = mg.GOLDEN[name]
    d = load_scene(lr, name, res)
    s = d.scene()
    for k, j in enumerate(mg.JITTERS):
        prim, t = s.trace_primary(*j)
        same = prim == g["prim%d" % k]
        assert same.mean() >= 0.9999, "index agreement %.6f" % same.mean()
        both = same & (prim >= 0)
        assert np.array_equal(t[both], g["t%d" % k][both]), "hit distances must be bit-identical"
    img, _, st = s.render(spp=spp, seed=11, splits=1)
    assert st["rays"] == int(g["rays"]), "path geometry must replay exactly"
    ref = g["sum"] / np.float32(spp)
    ok = np.isfinite(ref).all(-1)
    assert np.isclose(img, ref, rtol=1e-4, atol=1e-5).all(-1)[ok].mean() >= 0.999
