"""North_star's second correctness criterion on the WHOLE film of the two mesh configs: the CUDA render and the CPU oracle (the
reference's algorithm, independent RNG streams) at equal spp — per channel |mean_g - mean_r| <= 3 sqrt(s2_g/n + s2_r/n), the
image-mean difference in standard errors, relMSE, NaN/Inf counts on both sides.
usage (GPU box): python tools/full_film_parity.py [spp]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lumillyrender_b200 as lr
from lumillyrender_b200.renderer import params_from_config
from oracle import oracle_py as orc

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=144046, ibl_height=1600)
print("# full-film statistical parity, CUDA path vs CPU oracle (faithful traversal, libm sincos, independent streams), %d spp, %d host threads" % (spp, os.cpu_count()))
for name, res in (("sample", (1920, 1370)), ("welcome-2018", (2138, 1536))):
    d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=ROOT, resolution=res)
    s = d.scene()
    o = orc.OracleScene(d.desc, keepalive=d)
    t0 = time.perf_counter()
    img, sq, st = s.render(spp=spp, seed=5, sumsq=True)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref_sum, ref_sq, ost = o.render(params_from_config(d.config, spp=spp, seed=99), traversal=0, rng_mode=1, math_mode=0)
    t_cpu = time.perf_counter() - t0
    ok = np.isfinite(img).all(-1) & np.isfinite(ref_sum).all(-1)
    a, b = img[ok].astype(np.float64), (ref_sum[ok] / spp).astype(np.float64)
    va = np.maximum(sq[ok] / spp - a ** 2, 0.0) * spp / (spp - 1)
    vb = np.maximum(ref_sq[ok] / spp - b ** 2, 0.0) * spp / (spp - 1)
    se = np.sqrt(va / spp + vb / spp)
    inf = se > 0
    frac = (np.abs(a - b)[inf] <= 3.0 * se[inf]).mean()
    exact = np.isclose(a[~inf], b[~inf], rtol=1e-5, atol=1e-7).mean() if (~inf).any() else 1.0
    z = abs(a.mean() - b.mean()) / (np.sqrt((se ** 2).sum()) / a.size)
    relmse = float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
    print("%s %dx%d (%d primitives): %d pixels, finite on both sides %.6f; channels within 3 sigma %.5f (99.7 %% expected), zero-variance channels equal %.5f; "
          "image mean gpu %.6f oracle %.6f, difference %.2f standard errors; relMSE %.4g; non-finite samples gpu %d oracle %d; "
          "rays/sample gpu %.4f oracle %.4f; render %.2f s on the GPU (%.0f Msamples/s incl. D2H), %.1f s on the CPU (%.2f Msamples/s)" % (
              name, res[0], res[1], d.config.n_prims, ok.size, ok.mean(), frac, exact, a.mean(), b.mean(), z, relmse, st["nonfinite_samples"],
              ost["nonfinite_samples"], st["rays"] / st["samples"], ost["rays"] / ost["samples"], t_gpu, st["samples"] / t_gpu / 1e6, t_cpu,
              ost["samples"] / t_cpu / 1e6), flush=True)
    s.close(); o.close()
