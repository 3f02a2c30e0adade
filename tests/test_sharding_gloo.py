"""spp-range sharding host logic (lumillyrender_b200/distributed.py) under torch.distributed with the gloo
backend, world_size 2, on CPU.  The per-rank renderer is the CPU oracle (tests may use it); on GPUs bench.py
plugs lr_render_accumulate_device + NCCL into the same function."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_ranges_tile_exactly():
    from lumillyrender_b200.distributed import shard_range
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_c_abi_shard_rule_is_the_python_one(lr):
    """lr_shard_range (what lr_render_multi and a Rust / C++ host use) and distributed.shard_range (what bench.py's ranks
    use) are the same rule; bad requests are errors, not ranges.  No GPU needed."""
    import ctypes as C
    from lumillyrender_b200 import capi
    from lumillyrender_b200.distributed import shard_range
    lib = capi.load_library()
    for begin in (0, 5):
        for total in (0, 1, 7, 64, 1000):
            for world in (1, 2, 3, 8):
                for r in range(world):
                    b, c = C.c_int32(-1), C.c_int32(-1)
                    assert lib.lr_shard_range(begin, total, r, world, C.byref(b), C.byref(c)) == 0
                    lo, hi = shard_range(total, r, world)
                    assert (b.value, c.value) == (begin + lo, hi - lo)
    b, c = C.c_int32(), C.c_int32()
    for bad in ((0, 8, 2, 2), (0, 8, -1, 2), (0, 8, 0, 0), (0, -1, 0, 1), (-1, 8, 0, 1)):
        assert lib.lr_shard_range(*bad, C.byref(b), C.byref(c)) != 0
    assert lib.lr_shard_range(0, 8, 0, 1, None, C.byref(c)) != 0


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import lumillyrender_b200 as lr
    from lumillyrender_b200.distributed import render_sharded
    from lumillyrender_b200.renderer import params_from_config
    from oracle import oracle_py as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lr.ensure_assets(ROOT, need_bunny=False, need_ibl=False)
    d = lr.Description(os.path.join(ROOT, "scenes", "new-cbox.toml"), asset_root=ROOT, resolution=(24, 16))
    o = orc.OracleScene(d.desc, keepalive=d)
    accum = torch.zeros((16, 24, 3), dtype=torch.float32)

    def accumulate(begin, count):
        s, _, _ = o.render(params_from_config(d.config, spp=count, spp_begin=begin, seed=5), traversal=1, sumsq=False)
        accum.add_(torch.from_numpy(s))

    img = render_sharded(accumulate, accum, 10, rank, world, dist=dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), img.numpy())
        full, _, _ = o.render(params_from_config(d.config, spp=10, seed=5), traversal=1, sumsq=False)
        np.save(os.path.join(out_dir, "full.npy"), full / np.float32(10))
    else:
        assert img is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "sharded.npy")
    b = np.load(tmp_path / "full.npy")
    # identical sample set; only the fp32 summation order differs (two partial sums vs one running sum)
    assert np.allclose(a, b, rtol=2e-6, atol=1e-7)
    assert a.mean() > 0
