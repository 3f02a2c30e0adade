"""Data point for BASELINE configs[3]'s "synthetic ~1M-triangle mesh": sample.toml and welcome-2018.toml with a
1,048,576-triangle stand-in (scene arrays beyond L1, close to L2 size).  Prints build time, depth and throughput."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lumillyrender_b200 as lr
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=1048576, ibl_height=1600)
for name, res, spp in [("sample", (1920, 1370), 16), ("welcome-2018", (2138, 1536), 8)]:
    d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=ROOT, resolution=res)
    c = d.desc.contents
    print(name, "triangles", c.n_triangles, "flat", c.n_flat_triangles, "nodes", c.n_nodes, "depth", c.bvh_depth, "bvh build s %.2f" % d.config.bvh_build_seconds, flush=True)
    s = d.scene()
    for rep in range(3):
        img, _, st = s.render(spp=spp, seed=rep)
    print("  kernel_ms %.2f Msamples/s %.1f Mrays/s %.1f mean %.4f" % (st["kernel_ms"], st["samples"] / st["kernel_ms"] / 1e3, st["rays"] / st["kernel_ms"] / 1e3, float(img.mean())), flush=True)
    img, _, st = s.render(spp=2, seed=0, count=True)
    print("  per ray: nodes %.2f tris %.2f" % (st["nodes_visited"] / st["rays"], st["tris_tested"] / st["rays"]), flush=True)
lr.ensure_assets(ROOT, bunny_tris=144046, ibl_height=1600)
