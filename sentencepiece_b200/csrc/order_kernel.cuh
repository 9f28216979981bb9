// order_kernel.cuh -- K0: length-bucketed processing order for the lane kernels.
//
// The lane kernels give each lane of a warp its own sentence, so a warp is busy for as long as
// its LONGEST sentence: with the sentences of a batch taken in input order the warp's lanes
// idle 20-50 % of the time (profiles/README.md, "lanes per instruction").  A counting sort of
// the sentence indices by byte length (4-byte buckets, longest first) makes the 32 sentences
// of a warp equally long and lets the dynamic scheduler start the expensive groups first.
// Only the ORDER of processing changes: every result is still stored under the sentence's own
// index, so the output is identical.  (The reference has no counterpart: its ThreadPool hands
// out sentences one by one, src/sentencepiece_processor.cc has no batch path.)
#ifndef SPM_B200_ORDER_KERNEL_CUH_
#define SPM_B200_ORDER_KERNEL_CUH_

#include <cstdint>

namespace spm_b200 {

constexpr uint32_t kOrderBuckets = 1024;
constexpr uint32_t kOrderThreads = 1024;
constexpr uint32_t kOrderPerThread = 4;

__device__ __forceinline__ uint32_t order_bucket(const uint64_t *offsets, uint32_t i) {
  const unsigned long long len = offsets[i + 1] - offsets[i];
  const uint32_t b = len >= 4ull * (kOrderBuckets - 1) ? kOrderBuckets - 1 : static_cast<uint32_t>(len >> 2);
  return kOrderBuckets - 1 - b;  // longest first
}

// The sort is segmented: sentences [g * seg, (g + 1) * seg) are ordered among themselves (grid.y = segments), so
// that a streamed batch can be processed piece by piece as it arrives; seg >= n sorts the whole batch.
// hist[g][b] += sentences of bucket b in segment g (hist zeroed by the caller)
__global__ void __launch_bounds__(kOrderThreads) order_hist_kernel(const uint64_t *offsets, uint32_t n, uint32_t seg,
                                                                    uint32_t *hist) {
  __shared__ uint32_t sh[kOrderBuckets];
  for (uint32_t b = threadIdx.x; b < kOrderBuckets; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const uint32_t in_seg = blockIdx.x * (kOrderThreads * kOrderPerThread);
  const uint32_t base = blockIdx.y * seg + in_seg;
#pragma unroll
  for (uint32_t j = 0; j < kOrderPerThread; ++j) {
    const uint32_t o = j * kOrderThreads + threadIdx.x;
    if (in_seg + o < seg && base + o < n) atomicAdd(&sh[order_bucket(offsets, base + o)], 1u);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < kOrderBuckets; b += blockDim.x)
    if (sh[b]) atomicAdd(&hist[blockIdx.y * kOrderBuckets + b], sh[b]);
}

// in-place exclusive scan of each segment's hist[kOrderBuckets]: one block of kOrderBuckets threads per segment
__global__ void __launch_bounds__(kOrderBuckets) order_scan_kernel(uint32_t *hist) {
  __shared__ uint32_t warp_sum[32];
  hist += blockIdx.x * kOrderBuckets;
  const uint32_t v = hist[threadIdx.x];
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if ((threadIdx.x & 31) >= static_cast<uint32_t>(d)) incl += t;
  }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = warp_sum[threadIdx.x];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, d);
      if (threadIdx.x >= static_cast<uint32_t>(d)) w += t;
    }
    warp_sum[threadIdx.x] = w;
  }
  __syncthreads();
  const uint32_t before = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0u;
  hist[threadIdx.x] = before + incl - v;
}

// order[cursor[b]++] = i for every sentence i of bucket b (cursor = the scanned histogram); one global
// atomic per (block, non-empty bucket), ranks inside the block through shared memory
__global__ void __launch_bounds__(kOrderThreads) order_scatter_kernel(const uint64_t *offsets, uint32_t n, uint32_t seg,
                                                                       uint32_t *cursor, uint32_t *order) {
  __shared__ uint32_t sh[kOrderBuckets];
  for (uint32_t b = threadIdx.x; b < kOrderBuckets; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const uint32_t in_seg = blockIdx.x * (kOrderThreads * kOrderPerThread);
  const uint32_t base = blockIdx.y * seg + in_seg;
  uint32_t bucket[kOrderPerThread], rank[kOrderPerThread];
#pragma unroll
  for (uint32_t j = 0; j < kOrderPerThread; ++j) {
    const uint32_t o = j * kOrderThreads + threadIdx.x;
    bucket[j] = 0xFFFFFFFFu; rank[j] = 0;
    if (in_seg + o < seg && base + o < n) {
      bucket[j] = order_bucket(offsets, base + o);
      rank[j] = atomicAdd(&sh[bucket[j]], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < kOrderBuckets; b += blockDim.x)
    if (sh[b]) sh[b] = atomicAdd(&cursor[blockIdx.y * kOrderBuckets + b], sh[b]);
  __syncthreads();
#pragma unroll
  for (uint32_t j = 0; j < kOrderPerThread; ++j) {
    const uint32_t o = j * kOrderThreads + threadIdx.x;
    if (bucket[j] != 0xFFFFFFFFu) order[blockIdx.y * seg + sh[bucket[j]] + rank[j]] = base + o;
  }
}

// Samples the batch's bytes for the share of ASCII spaces (the engine picks the unigram lane kernel's instantiation
// from it, engine.cu pick_fast_words): gridDim.x windows of 4096 bytes spread over [offsets[0], offsets[n]);
// out[0] += spaces, out[1] += bytes looked at.
__global__ void __launch_bounds__(256) sample_spaces_kernel(const uint8_t *bytes, const uint64_t *offsets, uint32_t n,
                                                            unsigned long long *out) {
  const unsigned long long lo = offsets[0], hi = offsets[n];
  if (hi <= lo) return;
  const unsigned long long total = hi - lo;
  const unsigned long long win = total < 4096ull ? total : 4096ull;
  const unsigned long long start = gridDim.x > 1 ? lo + (total - win) * blockIdx.x / (gridDim.x - 1) : lo;
  uint32_t spaces = 0, seen = 0;
  for (unsigned long long k = threadIdx.x; k < win; k += blockDim.x) {
    spaces += bytes[start + k] == 0x20u;
    ++seen;
  }
  for (int d = 16; d > 0; d >>= 1) {
    spaces += __shfl_xor_sync(0xFFFFFFFFu, spaces, d);
    seen += __shfl_xor_sync(0xFFFFFFFFu, seen, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, static_cast<unsigned long long>(spaces));
    atomicAdd(out + 1, static_cast<unsigned long long>(seen));
  }
}

}  // namespace spm_b200
#endif
