"""Wall-clock breakdown of the end-to-end call sequence bench.py times (scene upload, render with host buffers, destroy)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lumillyrender_b200 as lr
from lumillyrender_b200 import capi
lr.init(0)
lr.ensure_assets(ROOT, bunny_tris=144046, need_ibl=False)
d = lr.Description(os.path.join(ROOT, "scenes", "sample.toml"), asset_root=ROOT, resolution=(1920, 1370))
lib = capi.load_library()
host = torch.empty((1370, 1920, 3), dtype=torch.float32).pin_memory()
for i in range(int(os.environ.get("E2E_ITERS", "5"))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = d.scene()
    t1 = time.perf_counter()
    p = s.params(spp=64, seed=i)
    st = capi.LrStats()
    capi.check(lib.lr_render(s._s, C.byref(p), C.cast(host.data_ptr(), C.POINTER(C.c_float)), None, C.byref(st)))
    t2 = time.perf_counter()
    s.close()
    t3 = time.perf_counter()
    print("create %.1f ms  render call %.1f ms (kernel %.1f ms)  destroy %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st.kernel_ms, (t3 - t2) * 1e3), flush=True)
