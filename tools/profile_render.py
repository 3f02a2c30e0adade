"""Small driver for ncu: one warm-up + one measured render of a BASELINE scene (default: the bench workload at 16 spp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lumillyrender_b200 as lr

name = sys.argv[1] if len(sys.argv) > 1 else "sample"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
res = {"sample": (1920, 1370), "welcome-2018": (2138, 1536), "primitive": (2048, 2048), "new-cbox": (256, 256), "brdf": (960, 540)}[name]
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=int(os.environ.get("BUNNY_TRIS", "144046")), ibl_height=1600, need_ibl=(name == "welcome-2018"))
d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=ROOT, resolution=res)
s = d.scene()
for i in range(2):
    img, _, st = s.render(spp=spp, seed=i)
    print(name, "kernel_ms %.2f Msamples/s %.1f Mrays/s %.1f" % (st["kernel_ms"], st["samples"] / st["kernel_ms"] / 1e3, st["rays"] / st["kernel_ms"] / 1e3))
