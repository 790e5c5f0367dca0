// K1 -- byte histogram (replaces reference include/encoder.h:136-150, Encoder::do_caculate_frequency).
//
// HBM-bound by intent: N bytes read once, 2 KiB written.  The limiter on the SM is the shared-memory
// atomic rate, so the bins are laid out to make every atomic conflict-free:
//   bins[b][slot]  (u32, b = byte value, slot = 0..223)   -> bank = slot mod 32 = lane, for every lane of a warp.
// A warp's 32 increments therefore always hit 32 distinct banks whatever the data looks like -- uniform bytes
// (random bins) and the skewed config (one byte value 99.97 % of the time, where a shared per-warp histogram would
// serialise 32-way on one address) run at the same rate.  224 slots x 256 bins x 4 B = 224 KiB, one CTA per SM,
// grid = SM count (persistent, grid-stride over 128-bit vectors).
// Each slot is shared by TWO threads of different warps (warp w and warp w + 7), one counting in the low and one in
// the high 16 bits of the word: 14 warps per SM instead of 7 feed the atomic pipe (with 7 the kernel sat at half of
// that pipe's rate, 1.75 warps per scheduler). A 16-bit counter cannot overflow because the bins are folded into the
// 64-bit result after every kHistEpochVecs vectors per thread (65 408 bytes + the two edge bytes of block 0).
#include "gh_common.cuh"

namespace gh {

constexpr int kHistSlots = 224;
constexpr int kHistThreads = 2 * kHistSlots;
constexpr int kHistSmemBytes = 256 * kHistSlots * 4;
constexpr int kHistUnroll = 4;  // 128-bit loads in flight per thread (x2: the next batch is requested before this one is consumed)
constexpr u64 kHistEpochVecs = 4088;  // per thread between folds: 4088 * 16 + 2 < 65536; a multiple of kHistUnroll

__device__ __forceinline__ void hist_add_word(u32* col, u32 inc, u32 w) {
  atomicAdd(col + (w & 0xffu) * kHistSlots, inc);
  atomicAdd(col + ((w >> 8) & 0xffu) * kHistSlots, inc);
  atomicAdd(col + ((w >> 16) & 0xffu) * kHistSlots, inc);
  atomicAdd(col + (w >> 24) * kHistSlots, inc);
}

__device__ __forceinline__ void hist_add_vec(u32* col, u32 inc, const uint4& v) {
  hist_add_word(col, inc, v.x);
  hist_add_word(col, inc, v.y);
  hist_add_word(col, inc, v.z);
  hist_add_word(col, inc, v.w);
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_kernel(const uint8_t* __restrict__ in, u64 n, u64* __restrict__ hist) {
  GH_DYNAMIC_SMEM(smem_raw);
  u32* bins = reinterpret_cast<u32*>(smem_raw);
  const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr unsigned kSlotWarps = kHistSlots / 32;
  u32* col = bins + (warp % kSlotWarps) * 32 + lane;
  const u32 inc = warp < kSlotWarps ? 1u : 1u << 16;
  for (unsigned k = t; k < 256u * kHistSlots; k += kHistThreads) bins[k] = 0;
  __syncthreads();

  // bytes before the first 16-byte boundary and after the last whole vector: block 0, one byte per thread
  const u64 misalign = (16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15;
  const u64 head = misalign < n ? misalign : n;
  const u64 nvec = (n - head) / 16;
  const u64 tail_start = head + nvec * 16;
  if (blockIdx.x == 0) {
    if (t < head) atomicAdd(col + u32(in[t]) * kHistSlots, inc);
    if (tail_start + t < n) atomicAdd(col + u32(in[tail_start + t]) * kHistSlots, inc);
  }

  const uint4* vec = reinterpret_cast<const uint4*>(in + head);
  const u64 stride = u64(gridDim.x) * kHistThreads;
  const u64 block_first = u64(blockIdx.x) * kHistThreads;  // the block's thread with the most vectors
  const u64 block_iters = nvec > block_first ? (nvec - block_first + stride - 1) / stride : 0;
  const u64 epochs = block_iters ? (block_iters + kHistEpochVecs - 1) / kHistEpochVecs : 1;  // block-uniform
  const u64 batch = u64(kHistUnroll) * stride;
  u64 i = block_first + t;
  for (u64 e = 0; e < epochs; ++e) {
    // this epoch: the thread's vectors below `lim`
    const u64 epoch_end = block_first + t + (e + 1) * kHistEpochVecs * stride;
    const u64 lim = epoch_end < nvec ? epoch_end : nvec;
    // Software pipeline: batch k+1 is requested before batch k is consumed, so every thread keeps kHistUnroll to
    // 2*kHistUnroll 128-bit loads outstanding. With one CTA per SM the bytes in flight are what bounds the kernel
    // (Little's law: ~45 KB per SM are needed to cover HBM latency at full bandwidth).
    if (i + (kHistUnroll - 1) * stride < lim) {
      uint4 cur[kHistUnroll];
#pragma unroll
      for (int k = 0; k < kHistUnroll; ++k) cur[k] = ldg128(vec + i + k * stride);
      i += batch;
      while (i + (kHistUnroll - 1) * stride < lim) {
        uint4 nxt[kHistUnroll];
#pragma unroll
        for (int k = 0; k < kHistUnroll; ++k) nxt[k] = ldg128(vec + i + k * stride);
#pragma unroll
        for (int k = 0; k < kHistUnroll; ++k) hist_add_vec(col, inc, cur[k]);
#pragma unroll
        for (int k = 0; k < kHistUnroll; ++k) cur[k] = nxt[k];
        i += batch;
      }
#pragma unroll
      for (int k = 0; k < kHistUnroll; ++k) hist_add_vec(col, inc, cur[k]);
    }
    for (; i < lim; i += stride) hist_add_vec(col, inc, ldg128(vec + i));

    __syncthreads();
    // fold the 224 slots (two 16-bit counters each) into the result and clear them: warp w takes bins w, w+14, ...
    for (unsigned b = warp; b < 256; b += kHistThreads / 32) {
      u32 s = 0;
#pragma unroll
      for (unsigned c = 0; c < kSlotWarps; ++c) {
        u32* word = bins + b * kHistSlots + lane + 32 * c;
        const u32 v = *word;
        *word = 0;
        s += (v & 0xffffu) + (v >> 16);
      }
      s = warp_sum(s);
      if (lane == 0 && s) atomicAdd(hist + b, u64(s));
    }
    __syncthreads();
  }
}

}  // namespace gh

extern "C" int gh_histogram(const uint8_t* d_in, uint64_t n, uint64_t* d_hist256, int accumulate, void* stream) {
  using namespace gh;
  if (!d_hist256 || (n && !d_in)) return GH_ERR_ARG;
  if (!accumulate) GH_CUDA_TRY(cudaMemsetAsync(d_hist256, 0, 256 * sizeof(uint64_t), (cudaStream_t)stream));
  if (n == 0) return GH_OK;
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaGetLastError());
  // per device, hence on every call (it costs nothing)
  GH_CUDA_TRY(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes));
  if (n > (1ull << 46)) return GH_ERR_ARG;
  const u64 nvec = n / 16 + 1;
  u64 blocks = (nvec + kHistThreads - 1) / kHistThreads;
  if (blocks > u64(sms)) blocks = u64(sms);
  GH_LAUNCH(hist_kernel, unsigned(blocks), kHistThreads, kHistSmemBytes, stream, d_in, (u64)n, (u64*)d_hist256);
  return check_launch();
}
