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
    ok = diff <= 3.0 * se + 1e-6 + 1e-4 * np.abs(mean_b)
    img_se = np.sqrt((var_a / n_a + var_b / n_b).sum()) / mean_a.size
    z = abs(float(mean_a.mean()) - float(mean_b.mean())) / max(img_se, 1e-12)
    relmse = float(np.mean((mean_a - mean_b) ** 2 / (mean_b ** 2 + 1e-2)))
    return float(ok.mean()), z, relmse
