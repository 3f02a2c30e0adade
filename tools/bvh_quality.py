"""What a tree costs the render: the host's binned-SAH tree against the device-built radix tree with its top rebuilt by SAH above
subtrees of <= cut triangles (LR_BVH_TOP_CUT; 0 = the plain radix tree).  Prints build time (second build of a kind) and the best
of three renders per tree.  usage (GPU box): python tools/bvh_quality.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lumillyrender_b200 as lr

lr.init(0)
roots = {144046: lr.ensure_assets(ROOT, bunny_tris=144046, ibl_height=1600)}
roots[1048576] = lr.ensure_assets(os.path.join(tempfile.gettempdir(), "lumilly_q_1m"), bunny_tris=1048576, ibl_height=1600)
for name, res, spp, tris in (("sample", (1920, 1370), 16, 144046), ("welcome-2018", (2138, 1536), 8, 144046), ("sample", (1920, 1370), 16, 1048576),
                             ("welcome-2018", (2138, 1536), 8, 1048576))[:int(os.environ.get("BVHQ_CASES", "4"))]:
    d = lr.Description(os.path.join(ROOT, "scenes", name + ".toml"), asset_root=roots[tris], resolution=res)
    for kind in (sys.argv[1:] or ("host", "0", "64", "256", "512", "2048", "8192")):
        if kind == "host":
            d.rebuild_bvh("host"); sec = d.rebuild_bvh("host")
        else:
            cut = kind.rpartition(":")[2]                       # "512" or "lbvh:512"
            os.environ["LR_BVH_TOP_CUT"] = cut
            d.rebuild_bvh("device"); sec = d.rebuild_bvh("device")
        s = d.scene()
        best = None
        for rep in range(3):
            img, _, st = s.render(spp=spp, seed=rep)
            best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
        print("%-13s %8d tris  tree %-16s build %8.1f ms (kernels %6.2f)  nodes %7d depth %2d  render %7.2f ms  %6.0f Msamples/s" % (
            name, d.config.n_prims, "host SAH" if kind == "host" else "device " + kind, 1e3 * sec, d.config.bvh_device_kernel_ms if kind != "host" else 0.0, d.desc.contents.n_nodes, d.desc.contents.bvh_depth,
            best["kernel_ms"], best["samples"] / best["kernel_ms"] / 1e3), flush=True)
        s.close()
