#!/usr/bin/env python
"""bench.py -- rows scanned/sec for Pinot's filter -> project -> aggregate path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], concretised in SURVEY.md section 8d "C2"): 8 segments x 100 M rows per GPU, 8
dict-encoded fixed-bit INT columns (cardinalities 10 .. 1 000 000 -> 4..20 bits), query

    SELECT SUM(c5), COUNT(*) FROM t WHERE c3 BETWEEN lo AND hi AND c6 > K        (2-predicate range filter, ~25 %)

A "step" is ONE pass of that query over all of the rank's segments through the reference-facing plugin call
(B200PlanMaker.execute_segments -> pb200h_execute -> one persistent scan kernel).

    value        rows/s with the segments resident in HBM (Pinot loads a segment once, then serves queries from it)
    roofline     algorithmic bytes of the scan kernel (sum of bitsPerElement/8 of the touched columns x rows) / its
                 CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
    e2e          the same call with HOST-resident index buffers: every step uploads the touched columns from pinned
                 host memory (H2D inside the timed region), scans, and reads the results back
    cpu_baseline the CPU oracle (a restatement of the Java operator chain, kind "port") on a bounded sample
    --impl reference   the CPU restatement alone, all host threads (there is no JVM in this image: SURVEY.md section 0)

Multi-GPU (torchrun, one rank per GPU): segments shard one set per GPU with no data-path collective except the final
reduce of the (tiny) result -- weak scaling, value = total rows of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CARDS = [10, 100, 1_000, 10_000, 65_536, 100_000, 1_000_000, 1_000_000]  # c0..c7 -> 4,7,10,14,16,17,20,20 bits
VALUE_STEP = [1, 1, 1, 3, 1, 7, 2, 2]
VALUE_BASE = [0, 0, 0, 5, 0, 11, 1, 1]
TOUCHED = ["c3", "c5", "c6"]
METRIC = "rows scanned/sec per box, filter+groupby on 100M-row segments, 1/2/4/8 GPU"


def bits_of(card: int) -> int:
    return 1 if card <= 2 else int(card - 1).bit_length()


def column_specs(rank: int, seg: int):
    return [{"name": f"c{c}", "cardinality": CARDS[c], "value_base": VALUE_BASE[c], "value_step": VALUE_STEP[c],
             "seed": 1000 + 104729 * rank + 131 * seg + c} for c in range(8)]


def query_text(selectivity: float) -> str:
    # each predicate keeps sqrt(selectivity) of the dictionary (uniform dictIds): c3 a centred range, c6 the upper tail
    f = selectivity ** 0.5
    n3 = CARDS[3]
    lo_id = int(n3 * (1 - f) / 2)
    hi_id = lo_id + int(round(n3 * f)) - 1
    k_id = int(round(CARDS[6] * (1 - f))) - 1
    lo, hi = VALUE_BASE[3] + VALUE_STEP[3] * lo_id, VALUE_BASE[3] + VALUE_STEP[3] * hi_id
    k = VALUE_BASE[6] + VALUE_STEP[6] * k_id
    return f"SELECT SUM(c5), COUNT(*) FROM benchTable WHERE c3 BETWEEN {lo} AND {hi} AND c6 > {k}"


def bytes_per_row() -> float:
    return sum(bits_of(CARDS[int(c[1:])]) for c in TOUCHED) / 8.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (t, sm, max, power, reasons[4])
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((time.time(), float(p[1]), float(p[2]), float(p[3]), p[4:8]))
            except (ValueError, IndexError):
                pass

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-5:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in inside for n, v in zip(names, s[4]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": max(s[2] for s in inside),
                "power_w_max": max(s[3] for s in inside), "samples": len(inside), "reasons": reasons}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_segment_for_oracle(sb, dev_seg, rows: int, name: str):
    """First `rows` docs of the touched columns of a device segment as oracle-readable Pinot buffers."""
    cols = []
    for cname in TOUCHED:
        info = dev_seg.column_info(cname)
        fwd = dev_seg.read_index(cname, "fwd")[: rows * info["bits"] // 8].copy()
        dct = dev_seg.read_index(cname, "dict")
        cols.append(sb.ColumnData(cname, sb.INT, True, info["bits"], info["cardinality"], False, 4, fwd, dct, None))
    return sb.SegmentData(name, rows, cols)


def time_oracle(oracle, segs, q, threads: int):
    """Runs the oracle on `segs` with `threads` worker threads (ctypes releases the GIL). Returns (seconds, results)."""
    results = [None] * len(segs)
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(segs):
                return
            results[i] = oracle.execute(segs[i], q)

    ts = [threading.Thread(target=work) for _ in range(min(threads, len(segs)))]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per segment")
    ap.add_argument("--selectivity", type=float, default=0.25)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-rows", type=int, default=8_000_000, help="rows per oracle work item")
    ap.add_argument("--quick", action="store_true", help="tuning runs: skip the e2e and cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 only

    import torch
    from pinot_b200 import sql
    from pinot_b200.plan_maker import B200Context, B200PlanMaker, IndexSegment

    dist = None
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    q = sql.parse(query_text(args.selectivity))
    bpr = bytes_per_row()
    config = {"workload": f"C2: {args.segments} segments x {args.rows} rows per GPU, 8 dict-encoded fixed-bit INT columns "
                          f"(bits 4,7,10,14,16,17,20,20); {query_text(args.selectivity)}",
              "segments_per_gpu": args.segments, "rows_per_segment": args.rows, "selectivity": args.selectivity,
              "touched_bits_per_row": int(bpr * 8), "l2_policy": "inputs larger than L2 (touched columns = "
              f"{args.segments * args.rows * bpr / 1e9:.2f} GB per step per GPU vs 126 MB L2)",
              "parallelism": f"segments sharded {args.segments}/GPU x {world} GPU, result reduced once"}

    # ---------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        from oracle import segment_builder as sb
        from oracle.pinot_oracle import oracle as get_oracle
        o = get_oracle()
        cores = os.cpu_count() or 1
        ctx = B200Context(local_rank)  # only to GENERATE the synthetic segment bytes (device generator), not to scan
        sample_rows = min(args.rows, args.cpu_sample_rows)
        items = max(cores, 1)
        dev_segs = [IndexSegment.synthetic(ctx, f"seg{s}", sample_rows, column_specs(0, s))
                    for s in range(min(args.segments, items))]
        base = [host_segment_for_oracle(sb, d, sample_rows, d.name) for d in dev_segs]
        for d in dev_segs:
            d.destroy()
        ctx.close()
        work = [base[i % len(base)] for i in range(items)]
        for _ in range(max(1, min(args.warmup, 1))):
            time_oracle(o, work[:cores], q, cores)
        steps = max(1, min(args.steps, 3))
        times = [time_oracle(o, work, q, cores)[0] for _ in range(steps)]
        sec = statistics.mean(times)
        rows = items * sample_rows
        val = rows / sec
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rows/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "rows/s", "cores": cores, "kind": "port",
                                 "sample": f"{items} work items x {sample_rows} rows (prefixes of the C2 segments), "
                                           f"{cores} threads, C++ restatement of the Java operator chain (no JVM in image)"},
                "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------- B200 arm
    ctx = B200Context(local_rank)
    pm = B200PlanMaker(ctx)
    t_gen = time.perf_counter()
    segs = [IndexSegment.synthetic(ctx, f"r{rank}s{s}", args.rows, column_specs(rank, s)) for s in range(args.segments)]
    gen_s = time.perf_counter() - t_gen
    rows_per_step = args.segments * args.rows

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    result_dev = torch.zeros(2, dtype=torch.float64, device=f"cuda:{local_rank}") if dist is not None else None

    def step():
        blocks = pm.execute_segments(segs, q)
        s = sum(float(b.doubles[0][0]) for b in blocks)
        c = sum(int(b.longs[1][0]) for b in blocks)
        if dist is not None:  # the one exchange step of the path: reduce the per-rank result to rank 0
            result_dev.copy_(torch.tensor([s, float(c)], dtype=torch.float64))
            dist.reduce(result_dev, dst=0)
        return s, c, blocks[0].device_ms

    for _ in range(args.warmup):
        first = step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        s, c, dms = step()
        kernel_ms.append(dms)
    barrier()
    elapsed = time.perf_counter() - t0
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    assert (s, c) == first[:2], "non-deterministic result across steps"
    expect = rows_per_step * args.selectivity
    assert abs(c - expect) < 0.02 * expect + 10, (c, expect)
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    ms_per_step = elapsed / args.steps * 1e3
    value = rows_per_step * world / (elapsed / args.steps)
    k_ms = statistics.mean(kernel_ms)
    peak, peak_src = measured_peak_gbs()
    achieved = rows_per_step * bpr / (k_ms * 1e-3) / 1e9

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "ms_per_step": ms_per_step, "kernel_ms": k_ms,
                              "achieved_gbs": achieved, "frac": achieved / peak, "clocks": clocks,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("PB200_")},
                              "selectivity": args.selectivity, "count": c}))
        for sgm in segs:
            sgm.destroy()
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- e2e: host-resident index buffers, H2D inside the timed region (rank-local; N>1: max over ranks) ----
    pinned = []
    h2d = 0
    for sgm in segs:
        cols = []
        for cname in TOUCHED:
            info = sgm.column_info(cname)
            fwd = sgm.read_index(cname, "fwd")
            pt = torch.empty(len(fwd), dtype=torch.uint8).pin_memory()
            pt.numpy()[:] = fwd
            dct = sgm.read_index(cname, "dict")
            cols.append((cname, info, pt, dct))
            h2d += len(fwd) + len(dct)
        pinned.append(cols)

    class _Col:  # duck type of IndexSegment.from_columns' column description
        def __init__(self, name, info, fwd, dct):
            self.name, self.data_type, self.has_dictionary = name, 0, True
            self.bits, self.cardinality, self.is_sorted, self.dict_entry_bytes = info["bits"], info["cardinality"], False, 4
            self.fwd, self.dict, self.inv = fwd, dct, None

    def e2e_step():
        loaded = [IndexSegment.from_columns(ctx, f"e2e{i}", args.rows, [_Col(n, inf, pt.numpy(), d) for n, inf, pt, d in cols])
                  for i, cols in enumerate(pinned)]
        blocks = pm.execute_segments(loaded, q)
        out = (sum(float(b.doubles[0][0]) for b in blocks), sum(int(b.longs[1][0]) for b in blocks))
        for l in loaded:
            l.destroy()
        return out

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        es = e2e_step()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    assert es == (s, c), "e2e result differs from the resident result"
    if dist is not None:
        t = torch.tensor([e2e_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = rows_per_step * world / (e2e_elapsed / args.e2e_steps)
    del pinned

    # ---- cpu_baseline: the oracle on a bounded sample, rank 0, N == 1 only ----
    cpu = None
    if rank == 0 and world == 1:
        from oracle import segment_builder as sb
        from oracle.pinot_oracle import oracle as get_oracle
        o = get_oracle()
        cores = os.cpu_count() or 1
        sample_rows = min(args.rows, args.cpu_sample_rows)
        base = [host_segment_for_oracle(sb, sgm, sample_rows, sgm.name) for sgm in segs[: min(len(segs), cores)]]
        work = [base[i % len(base)] for i in range(cores)]
        time_oracle(o, work[: min(4, cores)], q, cores)  # page-in
        sec, res = time_oracle(o, work, q, cores)
        # parity of the sample: the GPU path on the same prefix rows must agree with the oracle
        cpu = {"value": cores * sample_rows / sec, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"{cores} work items x {sample_rows} rows (prefixes of this run's segments), {cores} threads, "
                         "C++ restatement of the Java operator chain (oracle/); no JVM in the image"}

    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the scan kernel, from the committed `ncu --set full`
    # capture of this same workload (profiles/): only quoted when this run IS that workload
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_c2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("segments") == args.segments and tj.get("rows_per_segment") == args.rows and abs(tj.get("selectivity", -1) - args.selectivity) < 1e-9:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "peak_source": peak_src,
                             "kernel": "pb200::scan_kernel<6,false> (W=6 warps, 2 CTAs/SM)", "kernel_ms": k_ms,
                             "algorithmic_bytes_per_launch": rows_per_step * bpr},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 152 * args.segments, "steps": args.e2e_steps,
                        "note": "every step re-uploads the touched columns from pinned host memory (PCIe bound); "
                                "`value` is the same plugin call with the segments resident in HBM"},
                "gpu_launches": args.steps * 1, "clocks": clocks, "segment_generation_s": gen_s,
                "result": {"sum_c5": s, "count": c, "scope": "rank 0's segments",
                           "all_ranks": None if result_dev is None else {"sum_c5": float(result_dev[0].item()),
                                                                         "count": int(result_dev[1].item())}}}
        print(json.dumps(line))
    for sgm in segs:
        sgm.destroy()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
