"""Ad-hoc GPU probe: replay agreement per scene at several tolerances + first throughput numbers."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import lumillyrender_b200 as lr
from oracle import oracle_py as orc
from conftest import load_scene, make_params
from test_gpu_parity import SCENE_RES

lr.init(0)
print(lr.device_info())
lr.ensure_assets(ROOT, bunny_tris=20000, ibl_height=256)
for name, res in SCENE_RES.items():
    d = load_scene(lr, name, res)
    s = d.scene(); o = orc.OracleScene(d.desc, keepalive=d)
    spp = 8
    img, sq, st = s.render(spp=spp, seed=11, splits=1, sumsq=True)
    ref_sum, ref_sq, ost = o.render(make_params(lr, d.config, spp=spp, seed=11), traversal=0, rng_mode=0, math_mode=1)
    ref = ref_sum / spp
    fin = np.isfinite(ref).all(-1) & np.isfinite(img).all(-1)
    out = {}
    for rtol in (1e-6, 1e-5, 1e-4, 2e-3):
        out[rtol] = float(np.isclose(img, ref, rtol=rtol, atol=rtol * 0.1).all(-1)[fin].mean())
    print(name, 'rays', st['rays'], ost['rays'], 'nonfinite', st['nonfinite_samples'], ost['nonfinite_samples'], 'agree', out, 'identical px', float((img == ref).all(-1).mean()))

if '--bench' in sys.argv:
    lr.ensure_assets(ROOT, bunny_tris=144046, ibl_height=1600)
    for name, res, spp in [('primitive', (2048, 2048), 16), ('new-cbox', (256, 256), 64), ('brdf', (960, 540), 64), ('sample', (1920, 1370), 16), ('welcome-2018', (2138, 1536), 8)]:
        d = load_scene(lr, name, res)
        print(name, 'prims', d.config.n_prims, 'nodes', d.desc.contents.n_nodes, 'depth', d.desc.contents.bvh_depth, 'bvh build s', d.config.bvh_build_seconds)
        s = d.scene()
        for rep in range(3):
            t0 = time.time()
            img, _, st = s.render(spp=spp, seed=rep)
            dt = time.time() - t0
            print('  rep', rep, 'kernel_ms %.2f' % st['kernel_ms'], 'wall %.3f' % dt, 'Msamples/s %.1f' % (st['samples'] / st['kernel_ms'] / 1e3), 'Mrays/s %.1f' % (st['rays'] / st['kernel_ms'] / 1e3), 'rays/sample %.2f' % (st['rays'] / st['samples']), 'splits', st['splits'], 'mean', float(img.mean()))
        img, _, st = s.render(spp=min(spp, 4), seed=0, count=True)
        print('  per ray: nodes %.1f tris %.1f spheres %.1f' % (st['nodes_visited'] / st['rays'], st['tris_tested'] / st['rays'], st['spheres_tested'] / st['rays']))
        lr.save_png(os.path.join(ROOT, 'gpurun_out', 'gpu_%s.png' % name), img[::2, ::2] if img.shape[0] > 1200 else img, d.config.gamma)
    print('L2 read GB/s (48 MB):', lr.measure_l2_read_gbs(48 << 20, 20))
    print('HBM read GB/s (4 GB):', lr.measure_hbm_read_gbs(4 << 30, 2))
