"""Multi-GPU combine: one process per GPU, segments sharded whole, ONE reduce of the per-GPU group tables.

What this replaces in the reference: the JVM-side merge of per-segment results blocks in ``GroupByCombineOperator``
(``core/operator/combine/GroupByCombineOperator.java:102-165`` -> ``IndexedTable.upsert`` with
``AggregationFunction.merge``) and ``AggregationResultsBlockMerger`` (``.../merger/AggregationResultsBlockMerger.java:34-44``).

Segments are independent units (the reference already runs one task per segment), so the data path needs no
collective: every rank scans its own segments with ``PB200_Q_MERGE_SEGMENTS`` (device-side combine into one dense table
keyed by raw dictId key -- valid because the shards share dictionaries) and the tables meet exactly once:

    kind      contents                               reduce op        merge() it equals
    i64       COUNT(*) per group + integer SUMs       SUM              Count/Sum/Avg merge (a + b)
    f64       FLOAT/DOUBLE SUMs                       SUM              Sum/Avg merge
    u32max    MAX as (dictId + 1), 0 = empty          MAX              MaxAggregationFunction.merge
    u32min    MIN as dictId, 0xFFFFFFFF = empty       MIN              MinAggregationFunction.merge

``torch.distributed`` (NCCL over NVLink / NVSwitch) is plumbing only: the tensors handed to it alias the library's own
device buffers (``pb200_result_device_buffers``), nothing is copied.  The tables are O(groups), not O(rows): the
exchange is latency-bound (C4: 100 000 groups x (8+8+4) B = 2 MB per rank).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List


KINDS = {"i64": 0, "f64": 1, "u32max": 2, "u32min": 3}
_TORCH_DTYPES = {"i64": "int64", "f64": "float64", "u32max": "int32", "u32min": "int32"}


def shard_segments(num_segments: int, world_size: int, rank: int) -> List[int]:
    """Whole segments per rank, contiguous and balanced (64 segments over 8 GPUs -> 8 each)."""
    base, extra = divmod(num_segments, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def reduce_buffers(buffers: Dict[str, "torch.Tensor"], dist, dst: int = 0) -> None:
    """In-place reduce of the four table kinds to rank `dst` (works on CPU tensors with gloo and CUDA with nccl).

    The two u32 tables are carried by int32 tensors (torch has no uint32 collectives): the top bit is flipped before and
    after the MAX / MIN reduce so that the signed order equals the unsigned one.
    """
    import torch
    for kind, t in buffers.items():
        if t is None or t.numel() == 0:
            continue
        if kind in ("i64", "f64"):
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
        elif kind in ("u32max", "u32min"):
            # unsigned order == signed order after flipping the top bit: reduce in place, no widened copy
            t.bitwise_xor_(-2147483648)
            dist.reduce(t, dst=dst, op=dist.ReduceOp.MAX if kind == "u32max" else dist.ReduceOp.MIN)
            t.bitwise_xor_(-2147483648)
        else:
            raise KeyError(kind)


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a raw device pointer without copying."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def device_buffers(ctx, block) -> Dict[str, "torch.Tensor"]:
    """torch views over the dense group-table buffers of a merged results block (needs keep_handle=True)."""
    import torch
    out = {}
    if block.handle is None:
        raise ValueError("results block was not executed with merge=True, keep_handle=True")
    for kind, code in KINDS.items():
        p, n = C.c_void_p(), C.c_int64()
        rc = ctx.lib.pb200_result_device_buffers(block.handle, code, C.byref(p), C.byref(n))
        if rc != 0:
            from . import _lib
            _lib.check(rc)
        if not p.value or n.value == 0:
            out[kind] = None
            continue
        typestr = {"i64": "<i8", "f64": "<f8", "u32max": "<i4", "u32min": "<i4"}[kind]
        out[kind] = torch.as_tensor(_DeviceArray(p.value, n.value, typestr), device=f"cuda:{ctx.device}")
    return out


def combine_across_ranks(plan_maker, block, query, dist, dst: int = 0):
    """All ranks call this with their merged block; rank `dst` gets the block of the whole table, others get None."""
    import torch
    from .plan_maker import _read_result
    ctx = plan_maker.ctx
    bufs = device_buffers(ctx, block)  # pb200_execute returned after its stream finished: the tables are complete
    reduce_buffers(bufs, dist, dst)
    torch.cuda.synchronize(ctx.device)  # the extraction below runs on the library's own stream
    out = None
    if dist.get_rank() == dst:
        from . import _lib
        _lib.check(ctx.lib.pb200_result_finalize(ctx.handle, block.handle))
        out = _read_result(ctx, block.handle, query, 1, keep_handle=False)
    else:
        ctx.lib.pb200_result_free(block.handle)
    block.handle = None
    return out
