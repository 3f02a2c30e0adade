"""Pins a replay divergence (GPU ray count != oracle ray count under the shared RNG) on a pixel, a sample and a ray.
Bisects the crop by rows, pixels and sample indices through the C ABI, then walks the oracle's rays of that sample
(oracle orc_trace_path) through the device's two nearest-hit probes (strict query / the render kernels' query).
usage: python tools/replay_divergence.py <scene> <mesh triangles> <W>x<H> <crop x,y,w,h> <spp> <seed>"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lumillyrender_b200 as lr
from lumillyrender_b200.renderer import params_from_config
from oracle import oracle_py as orc

name, tris, res, crop, spp, seed = sys.argv[1], int(sys.argv[2]), tuple(int(v) for v in sys.argv[3].split("x")), tuple(int(v) for v in sys.argv[4].split(",")), int(sys.argv[5]), int(sys.argv[6])
lr.init(0)
root = lr.ensure_assets(os.path.join(tempfile.gettempdir(), "lumilly_div_%d" % tris), bunny_tris=tris, ibl_height=1600)
d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=root, resolution=res)
s = d.scene()
o = orc.OracleScene(d.desc, keepalive=d)


def rays(c, n, begin=0):
    _, _, st = s.render(spp=n, spp_begin=begin, seed=seed, splits=1, crop=c)
    _, _, ost = o.render(params_from_config(d.config, spp=n, spp_begin=begin, seed=seed, crop=c), traversal=0, rng_mode=0, math_mode=1, sumsq=False)
    return st["rays"], ost["rays"]


x0, y0, w, h = crop
print("whole crop:", rays(crop, spp))
found = []
for y in range(y0, y0 + h):
    g, r = rays((x0, y, w, 1), spp)
    if g != r:
        for x in range(x0, x0 + w):
            g, r = rays((x, y, 1, 1), spp)
            if g != r:
                for k in range(spp):
                    g, r = rays((x, y, 1, 1), 1, k)
                    if g != r:
                        found.append((x, y, k, g, r))
print("divergent (x, y, sample, gpu rays, oracle rays):", found)
for x, y, k, g, r in found[:4]:
    p = params_from_config(d.config, spp=1, spp_begin=k, seed=seed)
    org, dirs, prim, t = o.trace_path(p, x, y, k)
    ps, ts, ns = s.trace_rays(org, dirs, normals=True)
    pr, tr, nr = s.trace_rays(org, dirs, normals=True, render_query=True)
    pb, tb, nb = o.trace_rays(org, dirs, brute_force=True)
    for i in range(len(prim)):
        flag = "" if (ps[i] == prim[i] and pr[i] == prim[i] and ts[i] == t[i] and tr[i] == t[i]) else "   <-- differs"
        print("  ray %2d: oracle prim %8d t %.9g | strict prim %8d t %.9g | render-query prim %8d t %.9g | brute prim %8d t %.9g%s" % (
            i, prim[i], t[i], ps[i], ts[i], pr[i], tr[i], pb[i], tb[i], flag))
        if flag:
            print("          origin", org[i].tolist(), "direction", dirs[i].tolist())
            break
