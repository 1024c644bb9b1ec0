// pb200_scan.cuh -- the fused scan kernel: DocIdSet -> Projection -> Aggregation / GroupBy in one pass over HBM.
//
// What the reference does per segment, block (<=10 000 docs) at a time, on one CPU thread
//   FilterOperator tree -> DocIdSetOperator -> ProjectionOperator -> GroupByOperator/AggregationOperator
//   (core/operator/DocIdSetOperator.java:59-86, ProjectionOperator.java:68-79, query/GroupByOperator.java:101-140,
//    query/AggregationOperator.java:64-80; bit unpack in seglocal/io/reader/impl/FixedBitIntReader.java)
// is done here for ALL segments of a query by one persistent kernel:
//
//   * tile = consumer_warps x 1024 rows of every touched column; a producer thread streams tiles HBM -> shared memory
//     with TMA 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) through an N-stage full/empty mbarrier ring;
//   * each consumer thread owns 32 consecutive rows: unpacks its B big-endian words per column (pb200_unpack.cuh),
//     evaluates every filter leaf into a 32-bit row mask, combines masks with the filter's boolean program,
//     then aggregates the surviving rows: register accumulators + warp reduction + one atomic per warp for
//     aggregation-only queries, atomics into a dense group table (raw key = sum dictId_j * mult_j, the same key
//     DictionaryBasedGroupKeyGenerator computes, :311-346) for group-by;
//   * integer work only: no tensor cores; the bound is HBM bandwidth (algorithmic bytes = sum of bits/8 per row).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pb200_desc.h"
#include "pb200_unpack.cuh"

namespace pb200 {

// ------------------------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives (PTX ISA 8.x; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`; streaming data: L2 evict-first policy.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// leaf evaluation on a thread's 32 unpacked dictIds
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t eval_range(const uint32_t (&v)[32], uint32_t lo, uint32_t span) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= ((v[j] - lo) < span ? 1u : 0u) << j;
  return m;
}
__device__ __forceinline__ uint32_t eval_lut(const uint32_t (&v)[32], const uint32_t* __restrict__ bits) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= ((__ldg(bits + (v[j] >> 5)) >> (v[j] & 31)) & 1u) << j;
  return m;
}
// rows [row0, row0+32) against inclusive doc-id ranges
__device__ __forceinline__ uint32_t eval_doc_ranges(long long row0, const int32_t* __restrict__ r, int n) {
  uint32_t m = 0;
  for (int i = 0; i < n; ++i) {
    long long lo = (long long)__ldg(r + 2 * i) - row0, hi = (long long)__ldg(r + 2 * i + 1) - row0 + 1;  // [lo, hi)
    lo = lo < 0 ? 0 : lo;
    hi = hi > 32 ? 32 : hi;
    if (hi > lo) {
      uint32_t upto_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
      uint32_t upto_lo = (1u << lo) - 1u;  // lo < 32 here
      m |= upto_hi & ~upto_lo;
    }
  }
  return m;
}


// One value of a thread's 32-row group straight from the tile: bit offset j*bits inside the group's `bits` words.
__device__ __forceinline__ uint32_t read_one_group(const uint32_t* __restrict__ p, int j, int bits) {
  const int bit = j * bits;
  const int k = bit >> 5, s = bit & 31;
  const uint32_t hi = bswap32(p[k]);
  const uint32_t lo = (s + bits > 32) ? bswap32(p[k + 1]) : 0u;
  const uint32_t x = __funnelshift_l(lo, hi, s);
  return bits == 32 ? x : (x >> (32 - bits));
}
// Branch-free predicated gathers: the load is skipped (not just masked) for rows that did not survive the filter.
__device__ __forceinline__ int ldg_pred_s32(const int* p, uint32_t pred) {
  int x;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(x) : "l"(p), "r"(pred));
  return x;
}
__device__ __forceinline__ long long ldg_pred_s64(const long long* p, uint32_t pred) {
  long long x;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b64 %0, 0;\n\t@p ld.global.nc.b64 %0, [%1];\n\t}" : "=l"(x) : "l"(p), "r"(pred));
  return x;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  return x;
}
__device__ __forceinline__ uint32_t warp_min(uint32_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xFFFFFFFFu, x, o));
  return x;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xFFFFFFFFu, x, o));
  return x;
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
struct SmemHeader {
  SegDesc seg;            // consumer-side copy of the current segment's descriptor
  uint64_t full[8];       // stage filled by TMA
  uint64_t empty[8];      // stage drained by all consumer warps
};

template <int CW, bool GROUPBY>
__global__ void __launch_bounds__((CW + 1) * 32, (CW <= 3 ? 3 : 1))
scan_kernel(const __grid_constant__ QueryDesc q, const SegDesc* __restrict__ segs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem_raw);
  constexpr int kHdrBytes = (sizeof(SmemHeader) + 127) / 128 * 128;
  uint32_t* stages = reinterpret_cast<uint32_t*>(smem_raw + kHdrBytes);
  uint32_t* fstack = stages + (size_t)q.num_stages * q.stage_words;  // generic-filter mask stack (if !conj)
  constexpr int kConsumers = CW * 32;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool use_pipe = q.use_pipe != 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < q.num_stages; ++s) {
      mbar_init(&hdr->full[s], 1);
      mbar_init(&hdr->empty[s], CW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == CW) {
    // ============================== producer: one thread drives TMA ==============================
    if (lane == 0 && use_pipe) {
      const uint64_t policy = policy_evict_first();
      int sidx = 0, stage = 0;
      uint32_t phase = 0;
      for (long long T = blockIdx.x; T < q.total_tiles; T += gridDim.x) {
        while (T >= segs[sidx].first_tile + segs[sidx].num_tiles) ++sidx;
        const SegDesc* sd = segs + sidx;
        const long long t = T - sd->first_tile;
        mbar_wait(&hdr->empty[stage], phase ^ 1u);
        mbar_expect_tx(&hdr->full[stage], sd->stage_tx);
        uint32_t* dst = stages + (size_t)stage * q.stage_words;
        for (int s = 0; s < q.num_slots; ++s) {
          const uint32_t tb = sd->slots[s].tile_bytes;
          tma_load_1d(dst + sd->slots[s].stage_words, reinterpret_cast<const unsigned char*>(sd->slots[s].data) + t * tb,
                      tb, &hdr->full[stage], policy);
        }
        if (++stage == q.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ================================== consumers ==================================
  const int group = threadIdx.x;  // 32-row group inside the tile
  const SegDesc& sd = hdr->seg;

  // per-thread accumulators (aggregation-only kernel)
  unsigned long long cnt = 0;
  long long isum[kMaxAggs];
  double dsum[kMaxAggs];
  uint32_t mn[kMaxAggs], mx[kMaxAggs];
#pragma unroll
  for (int a = 0; a < kMaxAggs; ++a) { isum[a] = 0; dsum[a] = 0.0; mn[a] = 0xFFFFFFFFu; mx[a] = 0u; }

  auto flush = [&]() {
    // one atomic per warp per accumulator into the segment's AggAccum
    unsigned long long c = warp_sum(cnt);
    if (lane == 0 && c) atomicAdd(&sd.accum->count, c);
    cnt = 0;
    if (!GROUPBY) {
#pragma unroll
      for (int a = 0; a < kMaxAggs; ++a) {
        if (a < q.num_aggs && q.aggs[a].slot >= 0) {
          const int fn = q.aggs[a].function;
          if (fn == 1 || fn == 4) {  // SUM / AVG
            if (q.aggs[a].val_kind == VAL_DICT_F32 || q.aggs[a].val_kind == VAL_DICT_F64) {
              double d = warp_sum(dsum[a]);
              if (lane == 0) atomicAdd(&sd.accum->dsum[a], d);
            } else {
              long long s = warp_sum(isum[a]);
              if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&sd.accum->isum[a]), (unsigned long long)s);
            }
          } else if (fn == 2) {
            uint32_t m = warp_min(mn[a]);
            if (lane == 0) atomicMin(&sd.accum->min_id[a], m);
          } else if (fn == 3) {
            uint32_t m = warp_max(mx[a]);
            if (lane == 0) atomicMax(&sd.accum->max_id_plus1[a], m);
          }
          isum[a] = 0; dsum[a] = 0.0; mn[a] = 0xFFFFFFFFu; mx[a] = 0u;
        }
      }
    }
  };

  int sidx = -1, stage = 0;
  uint32_t phase = 0;
  for (long long T = blockIdx.x; T < q.total_tiles; T += gridDim.x) {
    // ---- segment change: flush accumulators, refresh the shared descriptor copy ----
    int ns = sidx < 0 ? 0 : sidx;
    while (T >= __ldg(&segs[ns].first_tile) + __ldg(&segs[ns].num_tiles)) ++ns;
    if (ns != sidx) {
      if (sidx >= 0) flush();
      consumer_bar_sync(kConsumers);  // everyone done reading the old descriptor
      const uint32_t* src = reinterpret_cast<const uint32_t*>(segs + ns);
      uint32_t* dstw = reinterpret_cast<uint32_t*>(&hdr->seg);
      for (int i = threadIdx.x; i < (int)(sizeof(SegDesc) / 4); i += kConsumers) dstw[i] = __ldg(src + i);
      consumer_bar_sync(kConsumers);
      sidx = ns;
    }
    const long long t = T - sd.first_tile;
    const long long row0 = t * q.tile_rows + (long long)group * kRowsPerThread;
    const long long left = sd.num_docs - row0;
    uint32_t m = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));

    if (use_pipe) mbar_wait(&hdr->full[stage], phase);
    const uint32_t* st = stages + (size_t)stage * q.stage_words;
    uint32_t v[32];

    // ---------------- phase 1: filter -> row mask ----------------
    // Conjunctions (the common case) evaluate leaf by leaf, most selective first (host order); once few rows per
    // thread survive, the remaining columns are probed per surviving row straight from the shared-memory tile
    // (the GPU form of ScanBasedDocIdIterator.applyAnd on the bitmap of survivors, AndDocIdSet.java:167-169)
    // instead of unpacking all 32 values.
    int v_slot = -1;  // which slot v[] currently holds (dense unpack), -1 = none
    if (q.num_nodes > 0) {
      if (q.conj) {
#pragma unroll
        for (int l = 0; l < kMaxLeaves; ++l) {
          if (l < q.num_leaves) {
            const LeafDesc& lf = sd.leaves[l];
            uint32_t lm;
            if (lf.kind == LEAF_ALL) lm = 0xFFFFFFFFu;
            else if (lf.kind == LEAF_NONE) lm = 0u;
            else if (lf.kind == LEAF_DOCMASK) lm = left > 0 ? __ldg(lf.bits + (row0 >> 5)) : 0u;
            else if (lf.kind == LEAF_DOCRANGES) lm = eval_doc_ranges(row0, lf.ranges, lf.num_ranges);
            else {
              const int wmax = __reduce_max_sync(0xFFFFFFFFu, __popc(m));
              const SlotDesc& sl = sd.slots[lf.slot];
              if (wmax == 0) {
                lm = 0u;  // nothing left to test (m == 0 in every lane)
              } else if (wmax <= q.sparse_max && v_slot != lf.slot) {
                const uint32_t* p = st + sl.stage_words + group * sl.bits;
                uint32_t keep = 0, mm = m;
                while (mm) {
                  const int j = __ffs(mm) - 1;
                  mm &= mm - 1;
                  const uint32_t id = read_one_group(p, j, sl.bits);
                  const bool hit = lf.kind == LEAF_RANGE ? ((id - lf.lo) < lf.span)
                                                         : (((__ldg(lf.bits + (id >> 5)) >> (id & 31)) & 1u) != 0u);
                  keep |= (hit ? 1u : 0u) << j;
                }
                lm = keep;  // bits outside m are irrelevant (m &= lm below)
              } else {
                if (v_slot != lf.slot) { unpack_group(sl.bits, st + sl.stage_words, group, v); v_slot = lf.slot; }
                lm = lf.kind == LEAF_RANGE ? eval_range(v, lf.lo, lf.span) : eval_lut(v, lf.bits);
              }
            }
            if (lf.negate) lm = ~lm;
            m &= lm;
          }
        }
      } else {
        uint32_t lm[kMaxLeaves];
#pragma unroll
        for (int l = 0; l < kMaxLeaves; ++l) {
          lm[l] = 0xFFFFFFFFu;
          if (l < q.num_leaves) {
            const LeafDesc& lf = sd.leaves[l];
            if (lf.kind == LEAF_NONE) lm[l] = 0u;
            else if (lf.kind == LEAF_DOCMASK) lm[l] = left > 0 ? __ldg(lf.bits + (row0 >> 5)) : 0u;
            else if (lf.kind == LEAF_DOCRANGES) lm[l] = eval_doc_ranges(row0, lf.ranges, lf.num_ranges);
          }
        }
        for (int s = 0; s < q.num_slots; ++s) {
          if (!(q.slot_roles[s] & ROLE_FILTER)) continue;
          unpack_group(sd.slots[s].bits, st + sd.slots[s].stage_words, group, v);
          v_slot = s;
#pragma unroll
          for (int l = 0; l < kMaxLeaves; ++l) {
            if (l < q.num_leaves && sd.leaves[l].slot == s) {
              const LeafDesc& lf = sd.leaves[l];
              if (lf.kind == LEAF_RANGE) lm[l] = eval_range(v, lf.lo, lf.span);
              else if (lf.kind == LEAF_LUT) lm[l] = eval_lut(v, lf.bits);
            }
          }
        }
#pragma unroll
        for (int l = 0; l < kMaxLeaves; ++l)
          if (l < q.num_leaves && sd.leaves[l].negate) lm[l] = ~lm[l];
        // postfix boolean program over masks; the stack lives in shared memory (column per thread: conflict free)
        uint32_t* stk = fstack + group;
        int sp = 0;
        for (int i = 0; i < q.num_nodes; ++i) {
          const int op = q.prog_op[i], arg = q.prog_arg[i];
          if (op == OP_LEAF) {
            uint32_t x = 0;
#pragma unroll
            for (int l = 0; l < kMaxLeaves; ++l) x = (l == arg) ? lm[l] : x;
            stk[(sp++) * kConsumers] = x;
          } else if (op == OP_NOT) {
            stk[(sp - 1) * kConsumers] = ~stk[(sp - 1) * kConsumers];
          } else {
            uint32_t x = stk[(sp - 1) * kConsumers];
            for (int c = 1; c < arg; ++c) {
              uint32_t y = stk[(sp - 1 - c) * kConsumers];
              x = op == OP_AND ? (x & y) : (x | y);
            }
            sp -= arg;
            stk[(sp++) * kConsumers] = x;
          }
        }
        m &= stk[0];
      }
    }

    // ---------------- phase 2: aggregate the surviving rows ----------------
    const int pc = __popc(m);
    cnt += pc;
    const int wmax2 = __reduce_max_sync(0xFFFFFFFFu, pc);
    if (wmax2 > 0 && wmax2 <= q.sparse_max) {
      // ---- sparse projection: per surviving row, read its dictIds from the tile (FixedBitIntReader.readUnchecked
      //      shape; what DataFetcher does with sparse docIds, core/common/DataFetcher.java:335-338) ----
      uint32_t mm = m;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        uint32_t g = 0;
        if (GROUPBY) {
#pragma unroll
          for (int gi = 0; gi < kMaxGroupBy; ++gi) {
            if (gi < q.num_group_by) {
              const SlotDesc& sl = sd.slots[q.group_slot[gi]];
              g += read_one_group(st + sl.stage_words + group * sl.bits, j, sl.bits) * sd.group_mult[gi];
            }
          }
          atomicAdd(sd.g_count + g, 1ull);
        }
#pragma unroll
        for (int a = 0; a < kMaxAggs; ++a) {
          if (a < q.num_aggs && q.aggs[a].slot >= 0) {
            const SlotDesc& sl = sd.slots[q.aggs[a].slot];
            const uint32_t id = read_one_group(st + sl.stage_words + group * sl.bits, j, sl.bits);
            const int fn = q.aggs[a].function, vk = q.aggs[a].val_kind;
            if (fn == 1 || fn == 4) {
              if (vk == VAL_DICT_F32 || vk == VAL_DICT_F64) {
                const double x = vk == VAL_DICT_F32 ? (double)__ldg(static_cast<const float*>(sd.dict[a]) + id)
                                                    : __ldg(static_cast<const double*>(sd.dict[a]) + id);
                if (GROUPBY) atomicAdd(sd.g_dsum[a] + g, x); else dsum[a] += x;
              } else {
                const long long x = vk == VAL_DICT_I32 ? (long long)__ldg(static_cast<const int*>(sd.dict[a]) + id)
                                    : vk == VAL_DICT_I64 ? __ldg(static_cast<const long long*>(sd.dict[a]) + id)
                                                         : (long long)(int)id;
                if (GROUPBY) atomicAdd(reinterpret_cast<unsigned long long*>(sd.g_isum[a] + g), (unsigned long long)x);
                else isum[a] += x;
              }
            } else if (fn == 2 || fn == 3) {
              const uint32_t x = id ^ (vk == VAL_RAW_I32 ? 0x80000000u : 0u);
              if (GROUPBY) { if (fn == 2) atomicMin(sd.g_min[a] + g, x); else atomicMax(sd.g_max[a] + g, x + 1u); }
              else { mn[a] = min(mn[a], x); mx[a] = max(mx[a], x + 1u); }
            } else if (fn == 5 && !GROUPBY) {
              atomicOr(sd.distinct_bits[a] + (id >> 5), 1u << (id & 31));
            }
          }
        }
      }
    } else if (wmax2 > 0) {
      // ---- dense projection: unpack all 32 values of each needed column ----
      uint32_t gid[32];
      if (GROUPBY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) gid[j] = 0;
        for (int s = 0; s < q.num_slots; ++s) {
          if (!(q.slot_roles[s] & ROLE_GROUP)) continue;
          if (v_slot != s) { unpack_group(sd.slots[s].bits, st + sd.slots[s].stage_words, group, v); v_slot = s; }
#pragma unroll
          for (int g = 0; g < kMaxGroupBy; ++g) {
            if (g < q.num_group_by && q.group_slot[g] == s) {
              const uint32_t mult = sd.group_mult[g];
#pragma unroll
              for (int j = 0; j < 32; ++j) gid[j] += v[j] * mult;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((m >> j) & 1u) atomicAdd(sd.g_count + gid[j], 1ull);
      }
      for (int s = 0; s < q.num_slots; ++s) {
        if (!(q.slot_roles[s] & ROLE_AGG)) continue;
        if (v_slot != s) { unpack_group(sd.slots[s].bits, st + sd.slots[s].stage_words, group, v); v_slot = s; }
#pragma unroll
        for (int a = 0; a < kMaxAggs; ++a) {
          if (a < q.num_aggs && q.aggs[a].slot == s) {
            const int fn = q.aggs[a].function;
            const int vk = q.aggs[a].val_kind;
            if (fn == 1 || fn == 4) {  // SUM / AVG: value = dictionary[dictId]
              if (vk == VAL_DICT_I32 || vk == VAL_RAW_I32) {
                const int* __restrict__ d = static_cast<const int*>(sd.dict[a]);
                int x[32];
                // all (predicated) gathers are issued before the first use: one L2 latency per tile, not 32
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = vk == VAL_RAW_I32 ? (((m >> j) & 1u) ? (int)v[j] : 0) : ldg_pred_s32(d + v[j], (m >> j) & 1u);
                if (GROUPBY) {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if ((m >> j) & 1u) atomicAdd(reinterpret_cast<unsigned long long*>(sd.g_isum[a] + gid[j]), (unsigned long long)(long long)x[j]);
                } else {
                  long long acc = 0;
#pragma unroll
                  for (int j = 0; j < 32; ++j) acc += x[j];
                  isum[a] += acc;
                }
              } else if (vk == VAL_DICT_I64) {
                const long long* __restrict__ d = static_cast<const long long*>(sd.dict[a]);
                long long x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = ldg_pred_s64(d + v[j], (m >> j) & 1u);
                if (GROUPBY) {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if ((m >> j) & 1u) atomicAdd(reinterpret_cast<unsigned long long*>(sd.g_isum[a] + gid[j]), (unsigned long long)x[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) isum[a] += x[j];
                }
              } else if (vk == VAL_DICT_F32) {
                const float* __restrict__ d = static_cast<const float*>(sd.dict[a]);
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __int_as_float(ldg_pred_s32(reinterpret_cast<const int*>(d + v[j]), (m >> j) & 1u));
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (GROUPBY) { if ((m >> j) & 1u) atomicAdd(sd.g_dsum[a] + gid[j], (double)x[j]); }
                  else dsum[a] += (double)x[j];
                }
              } else if (vk == VAL_DICT_F64) {
                const double* __restrict__ d = static_cast<const double*>(sd.dict[a]);
                double x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __longlong_as_double(ldg_pred_s64(reinterpret_cast<const long long*>(d + v[j]), (m >> j) & 1u));
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (GROUPBY) { if ((m >> j) & 1u) atomicAdd(sd.g_dsum[a] + gid[j], x[j]); }
                  else dsum[a] += x[j];
                }
              }
            } else if (fn == 2 || fn == 3) {  // MIN / MAX on dictIds (dictionaries are sorted: order preserving)
              const uint32_t bias = vk == VAL_RAW_I32 ? 0x80000000u : 0u;  // raw INT: signed -> unsigned order
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if ((m >> j) & 1u) {
                  const uint32_t x = v[j] ^ bias;
                  if (GROUPBY) {
                    if (fn == 2) atomicMin(sd.g_min[a] + gid[j], x);
                    else atomicMax(sd.g_max[a] + gid[j], x + 1u);
                  } else {
                    mn[a] = min(mn[a], x);
                    mx[a] = max(mx[a], x + 1u);
                  }
                }
              }
            } else if (fn == 5 && !GROUPBY) {  // DISTINCTCOUNT: bitset of dictIds (RoaringBitmap.addN in the reference)
              uint32_t* bits = sd.distinct_bits[a];
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if ((m >> j) & 1u) atomicOr(bits + (v[j] >> 5), 1u << (v[j] & 31));
            }
          }
        }
      }
    }

    if (use_pipe) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->empty[stage]);
      if (++stage == q.num_stages) { stage = 0; phase ^= 1u; }
    }
  }
  if (sidx >= 0) flush();
}

}  // namespace pb200
