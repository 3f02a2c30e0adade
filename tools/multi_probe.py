"""Wall-clock throughput of lr_render_multi (one process, N GPUs) on the bench workload.
usage (GPU box): python tools/multi_probe.py [spp_total] [n_max]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import lumillyrender_b200 as lr

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_max = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=144046, need_ibl=False)
d = lr.Description(os.path.join(ROOT, "scenes", "sample.toml"), asset_root=ROOT, resolution=(1920, 1370))
ref = None
n = 1
while n <= n_max:
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        img, _, st = d.render_multi(list(range(n)), spp=spp, seed=rep)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, st)
    img, _, st = d.render_multi(list(range(n)), spp=spp, seed=0)
    if ref is None:
        ref = img
    err = float(np.abs(img - ref).max())
    # the handle form: set-up once, then renders only (what a host that renders more than once pays per render)
    t0 = time.perf_counter()
    ms = d.multi_scene(list(range(n)))
    setup = time.perf_counter() - t0
    walls = []
    for rep in range(5):
        t0 = time.perf_counter()
        h, _, sth = ms.render(spp=spp, seed=rep)
        walls.append(time.perf_counter() - t0)
    ms.close()
    print("lr_multi_render    N=%d spp=%d: set-up %.1f ms once, then per render min %.1f / median %.1f / max %.1f ms (%.0f Msamples/s at the median), "
          "slowest kernel %.1f ms" % (n, spp, setup * 1e3, min(walls) * 1e3, sorted(walls)[2] * 1e3, max(walls) * 1e3,
                                      sth["samples"] / sorted(walls)[2] / 1e6, sth["kernel_ms"]), flush=True)
    print("lr_render_multi N=%d spp=%d: wall %.1f ms (%.0f Msamples/s end to end, scene upload + render + reduce + D2H), slowest kernel %.1f ms, "
          "rays %d, max |diff| vs N=1 %.3g" % (n, spp, best[0] * 1e3, st["samples"] / best[0] / 1e6, best[1]["kernel_ms"], st["rays"], err), flush=True)
    n *= 2
