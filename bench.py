#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json metric: Msamples/s, Mrays/s).

Workload (BASELINE.json configs[4], the configuration the multi-GPU metric is quoted on):
  scenes/sample.toml at 1920x1370, pure path tracing, Cornell box + 144,046-triangle synthetic stand-in for the
  absent bunny.obj, spp sharded over the GPUs: every GPU renders SPP_PER_GPU sample indices of every pixel
  (weak scaling), accumulates per-pixel sums in HBM and ONE NCCL reduce to rank 0 sums the buffers.
A "step" is one full pass: render + reduce + normalise.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                   (the CPU restatement of the reference algorithm)

Prints ONE JSON line (rank 0).  `value` is device-timed with the scene resident in HBM; `e2e` goes through
the C ABI with host buffers (scene upload H2D + image D2H inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1370
SPP_PER_GPU = 64
BUNNY_TRIS = 144046
SCENE = "sample"
NODE_BYTES, TRI_BYTES, SPHERE_BYTES = 64, 48, 16      # 128-bit loads per visit: 4 / 3 / 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def oracle_sample(desc_owner, params_fn, spp, target_seconds=12.0):
    """Times the CPU restatement (faithful reference algorithm, all host threads) on a bounded pixel-strided
    sample of the same workload.  Returns (Msamples/s, Mrays/s, stats, description of the sample)."""
    from oracle import oracle_py as orc
    o = orc.OracleScene(desc_owner.desc, keepalive=desc_owner)
    cores = os.cpu_count() or 1
    # calibrate on a coarse stride, then pick the stride that gives ~target_seconds of wall time
    _, _, st = o.render(params_fn(spp=spp), traversal=0, rng_mode=1, math_mode=0, threads=cores, pixel_stride=16, sumsq=False)
    per_sample = max(st["render_seconds"], 1e-4) / max(st["samples"], 1)
    stride = 16
    for cand in (1, 2, 3, 4, 6, 8, 12, 16):
        n = ((WIDTH + cand - 1) // cand) * ((HEIGHT + cand - 1) // cand) * spp
        if n * per_sample <= target_seconds:
            stride = cand
            break
    p = params_fn(spp=spp)
    _, _, st = o.render(p, traversal=0, rng_mode=1, math_mode=0, threads=cores, pixel_stride=stride, sumsq=False)
    sec = st["render_seconds"]
    desc = "every %d-th pixel in x and y of the %dx%d film (%d pixels) at %d spp, faithful unordered BVH traversal" % (
        stride, WIDTH, HEIGHT, st["samples"] // spp, spp)
    return st["samples"] / sec / 1e6, st["rays"] / sec / 1e6, st, desc, cores


def reference_traversal(st):
    """What the reference's own traversal costs per ray ('for context'): bvh.rs:131-141 visits every node
    whose box the ray's LINE hits, unordered and unpruned, then fully tests every candidate leaf.  Counters of the
    restatement; bytes at the reference's sizes (AABB 36 B per node visited, aabb.rs:10-14; 3 x 12 B of vertices per
    primitive tested, triangle.rs:25-40)."""
    rays = max(st.get("rays", 0), 1)
    nodes, prims = st.get("nodes_visited", 0) / rays, st.get("prims_tested", 0) / rays
    return {"nodes_per_ray": nodes, "prims_per_ray": prims, "bytes_per_ray": 36.0 * nodes + 36.0 * prims}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import lumillyrender_b200 as lr
    lr.ensure_assets(ROOT, bunny_tris=BUNNY_TRIS, need_ibl=False)
    d = lr.Description(os.path.join(ROOT, "scenes", SCENE + ".toml"), asset_root=ROOT, resolution=(WIDTH, HEIGHT))
    from lumillyrender_b200.renderer import params_from_config

    def params_fn(spp):
        return params_from_config(d.config, spp=spp, seed=1)

    vals, rays = [], []
    desc, cores, st = "", 1, None
    budget = 150.0 / max(args.steps + args.warmup, 1)
    for i in range(args.warmup + args.steps):
        ms, mr, st, desc, cores = oracle_sample(d, params_fn, SPP_PER_GPU, target_seconds=min(12.0, budget))
        if i >= args.warmup:
            vals.append(ms)
            rays.append(mr)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "mrays_per_s": sum(rays) / len(rays),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (WIDTH * HEIGHT * SPP_PER_GPU * args.gpus) / (v * 1e6), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": desc,
                         "reference_traversal": reference_traversal(st),
                         "note": "C++ restatement of the reference CPU algorithm (oracle/); the Rust reference cannot be built offline"},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "scenes/sample.toml (BASELINE configs[4]) at %dx%d, integrator pt, Cornell box + %d-triangle synthetic bunny stand-in, "
                        "%d spp per GPU sharded by sample index, one NCCL reduce per step" % (WIDTH, HEIGHT, BUNNY_TRIS, SPP_PER_GPU),
            "resolution": [WIDTH, HEIGHT], "spp_per_gpu": SPP_PER_GPU, "spp_total": SPP_PER_GPU * n_gpus, "triangles": BUNNY_TRIS + 12,
            "integrator": "pt", "parallelism": "spp-range sharding x%d" % n_gpus,
            "l2": "256 MiB write between steps flushes L2 (scene arrays are ~10 MB and are re-read from L2 within a step by design)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import lumillyrender_b200 as lr
    from lumillyrender_b200.distributed import env_rank_world, render_sharded

    rank, local_rank, world = env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liblumilly_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lr.init(local_rank)
    if rank == 0:
        lr.ensure_assets(ROOT, bunny_tris=BUNNY_TRIS, need_ibl=False)
    if world > 1:
        dist.barrier()
    d = lr.Description(os.path.join(ROOT, "scenes", SCENE + ".toml"), asset_root=ROOT, resolution=(WIDTH, HEIGHT))
    s = d.scene()
    stream = torch.cuda.current_stream().cuda_stream
    accum = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    spp_total = SPP_PER_GPU * world

    def step(i):
        accum.zero_()
        render_sharded(lambda b, n: s.render_accumulate(accum.data_ptr(), None, stream=stream, spp_begin=b, spp=n, seed=1000 + i),
                       accum, spp_total, rank, world, dist=dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
        flush.fill_(i & 255)
    barrier()
    s.stats(stream)                                    # reset counters
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
        flush.fill_(i & 255)                           # L2 flush between timed iterations (outside the events)
    barrier()
    sampler.stop_flag = True
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    st = s.stats(stream)
    counts = torch.tensor([st["rays"], st["samples"], st["launches"]], dtype=torch.float64, device="cuda")
    kernel_ms = torch.tensor([st["kernel_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    rays, samples, launches = [float(x) for x in counts.tolist()]
    image_mean = float(accum.mean().item()) if rank == 0 else None

    # ---- e2e through the C ABI with host buffers: scene upload (H2D) + render + image download (D2H) per step
    host_img = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32).pin_memory()
    e2e_t = []
    h2d = d2h = 0
    import ctypes as C
    from lumillyrender_b200 import capi
    lib = capi.load_library()
    for i in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        s2 = d.scene()                                  # lr_scene_create: H2D of the flat arrays
        h2d = s2.h2d_bytes
        if world == 1:
            p = s2.params(spp=SPP_PER_GPU, seed=2000 + i)
            stats = capi.LrStats()
            capi.check(lib.lr_render(s2._s, C.byref(p), C.cast(host_img.data_ptr(), C.POINTER(C.c_float)), None, C.byref(stats)))
        else:
            accum.zero_()
            render_sharded(lambda b, n: s2.render_accumulate(accum.data_ptr(), None, stream=stream, spp_begin=b, spp=n, seed=2000 + i),
                           accum, spp_total, rank, world, dist=dist)
            if rank == 0:
                host_img.copy_(accum, non_blocking=True)
        torch.cuda.synchronize()
        s2.close()
        d2h = HEIGHT * WIDTH * 3 * 4
        barrier()
        if i >= 2:
            e2e_t.append(time.perf_counter() - t0)
    e2e_s = torch.tensor([sum(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = WIDTH * HEIGHT * spp_total * args.steps / float(e2e_s.item()) / 1e6

    if rank == 0:
        # ---- roofline inputs: algorithmic bytes per ray from ONE instrumented launch (outside the timed region)
        _, _, cst = s.render(spp=4, seed=7, count=True)
        b_ray = (NODE_BYTES * cst["nodes_visited"] + TRI_BYTES * cst["tris_tested"] + SPHERE_BYTES * cst["spheres_tested"]) / max(cst["rays"], 1)
        rays_per_launch = rays / max(launches, 1)
        launch_ms = float(kernel_ms.item()) / max(st["launches"], 1)
        achieved = b_ray * (st["rays"] / max(st["launches"], 1)) / (launch_ms * 1e-3) / 1e9
        peak, which = measured_peaks()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        try:
            l2_peak = lr.measure_l2_read_gbs(48 << 20, 20)
        except Exception:
            l2_peak = None
        line = {
            "metric": "Msamples/s", "value": samples / total_s / 1e6, "unit": "Msamples/s",
            "mrays_per_s": rays / total_s / 1e6, "rays_per_sample": rays / max(samples, 1),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "lr_scene_create (H2D of BVH/triangles/materials) + render + D2H of the fp32 image to pinned host memory, wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of %s)" % which,
                         "bytes_per_ray": b_ray, "nodes_per_ray": cst["nodes_visited"] / max(cst["rays"], 1),
                         "tris_per_ray": cst["tris_tested"] / max(cst["rays"], 1), "kernel": "render_pool_kernel<pt, tree> (pool.cuh)",
                         "kernel_ms_per_launch": launch_ms, "rays_per_launch": rays_per_launch / world},
            "roofline_l2": {"bound": "l2", "achieved": achieved, "peak": l2_peak, "unit": "GB/s",
                            "frac": (achieved / l2_peak) if l2_peak else None,
                            "peak_source": "lr_measure_l2_read_gbs: 48 MiB working set, ld.global.cg.v4, measured in this run"},
            "image_mean": image_mean,
        }
        if world == 1 and not args.no_cpu_baseline:
            from lumillyrender_b200.renderer import params_from_config
            ms, mr, ost, desc, cores = oracle_sample(d, lambda spp: params_from_config(d.config, spp=spp, seed=1), SPP_PER_GPU)
            line["cpu_baseline"] = {"value": ms, "unit": "Msamples/s", "mrays_per_s": mr, "cores": cores, "kind": "port", "sample": desc,
                                    "reference_traversal": reference_traversal(ost), "oracle_bvh_build_s": ost["build_seconds"]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
