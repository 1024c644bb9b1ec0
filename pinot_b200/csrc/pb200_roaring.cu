// pb200_roaring.cu -- device-side decode of RoaringBitmap inverted-index postings into a dense 1-bit-per-doc mask.
//
// Reference behaviour replaced: InvertedIndexFilterOperator.getNextBlockWithoutNullHandling
// (core/operator/filter/InvertedIndexFilterOperator.java:60-96): one dictId -> that bitmap; k dictIds ->
// ImmutableRoaringBitmap.or(bitmaps); NEQ / NOT_IN -> flip(0, numDocs) (done by the consumer of the mask as a negate
// flag).  Bitmaps are read in place from the index file bytes (BitmapInvertedIndexReader.getDocIds :45-62): portable
// RoaringFormatSpec serialization, little-endian, at arbitrary byte alignment inside the file.
//
// Mapping: blockIdx.y = requested bitmap, blocks stride over its containers; one CTA expands one container
// (array: a thread per value; run: a thread per run; bitmap: a thread per 32-bit word) with atomicOr into the mask.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "pb200_internal.h"

namespace pb200 {

__device__ __forceinline__ uint32_t ld16(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
__device__ __forceinline__ uint32_t ld32(const unsigned char* p) {
  return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

__global__ void __launch_bounds__(256) roaring_decode_kernel(const DecodeJob* __restrict__ jobs) {
  const DecodeJob ref = jobs[blockIdx.y];
  const unsigned char* __restrict__ inv = ref.inv;
  uint32_t* __restrict__ mask = ref.mask;
  const long long num_docs = ref.num_docs;
  if (ref.length < 8) return;  // empty bitmap: cookie + size 0
  const unsigned char* b = inv + ref.offset;
  const uint32_t cookie = ld32(b);
  const bool has_run = (cookie & 0xFFFFu) == 12347u;
  uint32_t n;
  unsigned long long pos;
  const unsigned char* run_flags = nullptr;
  if (has_run) {
    n = (cookie >> 16) + 1;
    pos = 4;
    run_flags = b + pos;
    pos += (n + 7) / 8;
  } else if (cookie == 12346u) {
    n = ld32(b + 4);
    pos = 8;
  } else {
    return;  // malformed; the host validated the first header, be defensive anyway
  }
  const unsigned char* desc = b + pos;
  pos += 4ull * n;
  const bool has_offsets = !has_run || n >= 4;
  const unsigned char* offs = b + pos;
  if (has_offsets) pos += 4ull * n;

  __shared__ unsigned long long s_start;
  for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
    const uint32_t key = ld16(desc + 4 * c);
    const uint32_t card = ld16(desc + 4 * c + 2) + 1;
    const bool is_run = has_run && ((run_flags[c >> 3] >> (c & 7)) & 1);
    unsigned long long start;
    if (has_offsets) {
      start = ld32(offs + 4 * c);
    } else {
      // < 4 containers with a run cookie carry no offset header: walk the (at most 3) predecessors
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long p = pos;
        for (uint32_t k = 0; k < c; k++) {
          const uint32_t kc = ld16(desc + 4 * k + 2) + 1;
          const bool kr = (run_flags[k >> 3] >> (k & 7)) & 1;
          if (kr) p += 2 + 4ull * ld16(b + p);
          else if (kc > 4096) p += 8192;
          else p += 2ull * kc;
        }
        s_start = p;
      }
      __syncthreads();
      start = s_start;
    }
    const unsigned char* cp = b + start;
    const long long base = (long long)key << 16;
    if (base >= num_docs) continue;
    uint32_t* mbase = mask + (base >> 5);
    if (is_run) {
      const uint32_t nruns = ld16(cp);
      for (uint32_t r = threadIdx.x; r < nruns; r += blockDim.x) {
        uint32_t s = ld16(cp + 2 + 4 * r), e = s + ld16(cp + 4 + 4 * r);  // inclusive [s, e]
        for (uint32_t w = s >> 5; w <= (e >> 5); w++) {
          uint32_t lo = w == (s >> 5) ? (s & 31) : 0, hi = w == (e >> 5) ? (e & 31) : 31;
          uint32_t bits = (hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
          atomicOr(mbase + w, bits);
        }
      }
    } else if (card > 4096) {
      for (uint32_t w = threadIdx.x; w < 2048; w += blockDim.x) {
        uint32_t x = ld32(cp + 4 * w);
        if (x) atomicOr(mbase + w, x);
      }
    } else {
      for (uint32_t i = threadIdx.x; i < card; i += blockDim.x) {
        uint32_t v = ld16(cp + 2 * i);
        atomicOr(mbase + (v >> 5), 1u << (v & 31));
      }
    }
  }
}

// One launch for ALL bitmaps a query needs (every segment, every inverted-index leaf): blockIdx.y = job.  Nothing here
// synchronises the stream; `jobs_dev` must stay alive until the stream has passed the kernel (the caller owns it).
int roaring_decode_batch(pb200_ctx* ctx, cudaStream_t stream, const std::vector<DecodeJob>& jobs, void* jobs_dev) {
  if (jobs.empty()) return PB200_OK;
  cudaError_t e = cudaMemcpyAsync(jobs_dev, jobs.data(), sizeof(DecodeJob) * jobs.size(), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { set_error("decode job upload failed: %s", cudaGetErrorString(e)); return PB200_E_CUDA; }
  long long max_docs = 0;
  for (const DecodeJob& j : jobs) max_docs = std::max(max_docs, j.num_docs);
  const long long containers = (max_docs + 65535) / 65536;
  for (size_t j0 = 0; j0 < jobs.size(); j0 += 65535) {
    const unsigned ny = (unsigned)std::min<size_t>(65535, jobs.size() - j0);
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(containers, 512)), ny);
    roaring_decode_kernel<<<grid, 256, 0, stream>>>((const DecodeJob*)jobs_dev + j0);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("roaring decode launch failed: %s", cudaGetErrorString(e)); return PB200_E_CUDA; }
  return PB200_OK;
}

}  // namespace pb200
