"""Kernel-organisation sweep: renders BASELINE scenes under the LR_* development knobs and prints throughput.
usage: python tools/sweep.py "LR_SCHED=0" "LR_SCHED=1,LR_SHADE_THRESH=8" ...   (each argument = one configuration)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lumillyrender_b200 as lr

SCENES = [("sample", (1920, 1370), 16), ("welcome-2018", (2138, 1536), 8), ("new-cbox", (256, 256), 64), ("brdf", (960, 540), 64),
          ("primitive", (2048, 2048), 16)]
only = os.environ.get("SWEEP_SCENES")
if only:
    SCENES = [s for s in SCENES if s[0] in only.split(",")]
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=144046, ibl_height=1600)
loaded = []
for name, res, spp in SCENES:
    d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=ROOT, resolution=res)
    loaded.append((name, spp, d, d.scene()))
keys = set()
for cfg in sys.argv[1:]:
    for k in keys:
        os.environ.pop(k, None)
    for kv in cfg.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            os.environ[k] = v
            keys.add(k)
    row = []
    for name, spp, d, s in loaded:
        best = None
        for rep in range(3):
            img, _, st = s.render(spp=spp, seed=rep)
            if best is None or st["kernel_ms"] < best[0]:
                best = (st["kernel_ms"], st["samples"] / st["kernel_ms"] / 1e3, st["rays"] / st["kernel_ms"] / 1e3, float(img.mean()))
        row.append("%s %.2fms %.0fMs/s %.0fMr/s r%d" % (name, best[0], best[1], best[2], st.get("gate_retraces", 0)))
    print("%-40s | %s" % (cfg, " | ".join(row)), flush=True)
