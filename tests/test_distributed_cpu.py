"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: segment sharding and the table reduce semantics."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pinot_b200.distributed import reduce_buffers, shard_segments


def test_shard_segments_covers_everything_once():
    for n, w in [(64, 8), (8, 8), (7, 2), (3, 4), (1, 1), (0, 2)]:
        got = [shard_segments(n, w, r) for r in range(w)]
        flat = [s for part in got for s in part]
        assert flat == list(range(n))
        assert max(len(p) for p in got) - min(len(p) for p in got) <= 1


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    G = 1000
    count = rng.integers(0, 5, size=G)
    sums = rng.integers(-10**12, 10**12, size=G) * (count > 0)
    fsum = rng.random(G) * (count > 0)
    mx = np.where(count > 0, rng.integers(0, 2**31 - 2, size=G) + 1, 0).astype(np.uint32)      # dictId + 1, 0 = empty
    mn = np.where(count > 0, rng.integers(0, 2**32 - 2, size=G), 0xFFFFFFFF).astype(np.uint32)  # raw order-preserving
    np.savez(os.path.join(out_dir, f"in{rank}.npz"), count=count, sums=sums, fsum=fsum, mx=mx, mn=mn)
    bufs = {"i64": torch.from_numpy(np.concatenate([count, sums]).astype(np.int64)),
            "f64": torch.from_numpy(fsum.copy()),
            "u32max": torch.from_numpy(mx.view(np.int32).copy()),
            "u32min": torch.from_numpy(mn.view(np.int32).copy())}
    reduce_buffers(bufs, dist, dst=0)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), i64=bufs["i64"].numpy(), f64=bufs["f64"].numpy(),
                 mx=bufs["u32max"].numpy().view(np.uint32), mn=bufs["u32min"].numpy().view(np.uint32))
    dist.destroy_process_group()


def test_reduce_buffers_world_size_2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ins = [np.load(tmp_path / f"in{r}.npz") for r in range(world)]
    out = np.load(tmp_path / "out.npz")
    G = 1000
    assert np.array_equal(out["i64"][:G], sum(i["count"] for i in ins))          # COUNT merge = a + b
    assert np.array_equal(out["i64"][G:], sum(i["sums"] for i in ins))           # SUM merge (exact integers)
    assert np.allclose(out["f64"], sum(i["fsum"] for i in ins), rtol=1e-12)
    assert np.array_equal(out["mx"], np.maximum(ins[0]["mx"], ins[1]["mx"]))     # MAX merge, 0 = empty loses
    assert np.array_equal(out["mn"], np.minimum(ins[0]["mn"], ins[1]["mn"]))     # MIN merge, 0xFFFFFFFF = empty loses
