// K1 -- byte histogram (replaces reference include/encoder.h:136-150, Encoder::do_caculate_frequency).
//
// HBM-bound by intent: N bytes read once, 2 KiB written.  The limiter on the SM is the shared-memory
// atomic rate, so the bins are laid out to make every atomic conflict-free:
//   bins[b][t]  (u32, b = byte value, t = thread)   -> bank = t mod 32 for every lane of a warp.
// Each thread owns one column, i.e. the warp-private histogram is striped one lane per bank. A warp's 32
// increments therefore always hit 32 distinct banks whatever the data looks like -- uniform bytes (random
// bins) and the skewed config (one byte value 99.97 % of the time, where a shared per-warp histogram would
// serialise 32-way on one address) run at the same rate.  224 threads x 256 bins x 4 B = 224 KiB, one CTA
// per SM, grid = SM count (persistent, grid-stride over 128-bit vectors).
#include "gh_common.cuh"

namespace gh {

constexpr int kHistThreads = 224;
constexpr int kHistSmemBytes = 256 * kHistThreads * 4;
constexpr int kHistUnroll = 8;  // 128-bit loads in flight per thread (x2: the next batch is requested before this one is consumed)

__device__ __forceinline__ void hist_add_word(u32* col, u32 w) {
  atomicAdd(col + (w & 0xffu) * kHistThreads, 1u);
  atomicAdd(col + ((w >> 8) & 0xffu) * kHistThreads, 1u);
  atomicAdd(col + ((w >> 16) & 0xffu) * kHistThreads, 1u);
  atomicAdd(col + (w >> 24) * kHistThreads, 1u);
}

__device__ __forceinline__ void hist_add_vec(u32* col, const uint4& v) {
  hist_add_word(col, v.x);
  hist_add_word(col, v.y);
  hist_add_word(col, v.z);
  hist_add_word(col, v.w);
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_kernel(const uint8_t* __restrict__ in, u64 n, u64* __restrict__ hist) {
  GH_DYNAMIC_SMEM(smem_raw);
  u32* bins = reinterpret_cast<u32*>(smem_raw);
  const unsigned t = threadIdx.x;
  u32* col = bins + t;
#pragma unroll 8
  for (int b = 0; b < 256; ++b) col[b * kHistThreads] = 0;

  // bytes before the first 16-byte boundary and after the last whole vector: block 0, one byte per thread
  const u64 misalign = (16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15;
  const u64 head = misalign < n ? misalign : n;
  const u64 nvec = (n - head) / 16;
  const u64 tail_start = head + nvec * 16;
  if (blockIdx.x == 0) {
    if (t < head) atomicAdd(col + u32(in[t]) * kHistThreads, 1u);
    if (tail_start + t < n) atomicAdd(col + u32(in[tail_start + t]) * kHistThreads, 1u);
  }

  const uint4* vec = reinterpret_cast<const uint4*>(in + head);
  const u64 stride = u64(gridDim.x) * kHistThreads;
  u64 i = u64(blockIdx.x) * kHistThreads + t;
  // Software pipeline: batch k+1 is requested before batch k is consumed, so every thread keeps kHistUnroll to
  // 2*kHistUnroll 128-bit loads outstanding. With one 224-thread CTA per SM the bytes in flight are what bounds
  // the kernel (Little's law: ~45 KB per SM are needed to cover HBM latency at full bandwidth).
  const u64 batch = u64(kHistUnroll) * stride;
  if (i + (kHistUnroll - 1) * stride < nvec) {
    uint4 cur[kHistUnroll];
#pragma unroll
    for (int k = 0; k < kHistUnroll; ++k) cur[k] = ldg128(vec + i + k * stride);
    i += batch;
    while (i + (kHistUnroll - 1) * stride < nvec) {
      uint4 nxt[kHistUnroll];
#pragma unroll
      for (int k = 0; k < kHistUnroll; ++k) nxt[k] = ldg128(vec + i + k * stride);
#pragma unroll
      for (int k = 0; k < kHistUnroll; ++k) hist_add_vec(col, cur[k]);
#pragma unroll
      for (int k = 0; k < kHistUnroll; ++k) cur[k] = nxt[k];
      i += batch;
    }
#pragma unroll
    for (int k = 0; k < kHistUnroll; ++k) hist_add_vec(col, cur[k]);
  }
  for (; i < nvec; i += stride) hist_add_vec(col, ldg128(vec + i));

  __syncthreads();
  // fold the 224 columns: warp w takes bins w, w+7, ...; lanes stride over the columns
  const unsigned lane = t & 31, warp = t >> 5;
  for (unsigned b = warp; b < 256; b += kHistThreads / 32) {
    u32 s = 0;
#pragma unroll
    for (int c = 0; c < kHistThreads / 32; ++c) s += bins[b * kHistThreads + lane + 32 * c];
    s = warp_sum(s);
    if (lane == 0 && s) atomicAdd(hist + b, u64(s));
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
  // each thread's u32 counters hold at most the bytes it reads: n / (grid * 224) + 32 < 2^32 needs n < ~1.4e14
  if (n > (1ull << 46)) return GH_ERR_ARG;
  const u64 nvec = n / 16 + 1;
  u64 blocks = (nvec + kHistThreads - 1) / kHistThreads;
  if (blocks > u64(sms)) blocks = u64(sms);
  GH_LAUNCH(hist_kernel, unsigned(blocks), kHistThreads, kHistSmemBytes, stream, d_in, (u64)n, (u64*)d_hist256);
  return check_launch();
}
