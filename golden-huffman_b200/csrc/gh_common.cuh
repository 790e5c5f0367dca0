// Shared device/host helpers for the sm_100a kernels. Compiled by nvcc for the product
// (-gencode arch=compute_100a,code=sm_100a); the GH_EMUL branch exists only so tests/emul can run the same
// kernel logic on the CPU under a pthread shim (test infrastructure, never shipped).
#ifndef GH_COMMON_CUH_
#define GH_COMMON_CUH_

#ifndef GH_EMUL
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "gh_internal.h"

namespace gh {

int cuda_fail(cudaError_t e);  // records the error for gh_last_cuda_error(), returns GH_ERR_CUDA
void note_launch();            // bumps gh_launch_count()
// Optional per-kernel timing (gh_profile_enable): CUDA events recorded on the launching stream around a launch.
void profile_begin(const char* kernel_name, void* stream);
void profile_end(void* stream);
int sm_count();                // SMs of the current device (148 on B200)
int check_launch();            // cudaGetLastError -> status

#define GH_CUDA_TRY(expr)                              \
  do {                                                 \
    cudaError_t gh_e_ = (expr);                        \
    if (gh_e_ != cudaSuccess) return ::gh::cuda_fail(gh_e_); \
  } while (0)

#ifdef GH_EMUL
#define GH_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (::gh::note_launch(), gh_emul::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); }))
#define GH_DYNAMIC_SMEM(name) unsigned char* name = gh_emul::g_block.dyn_smem
#else
#define GH_LAUNCH(kernel, grid, block, smem, stream, ...)                              \
  (::gh::note_launch(), ::gh::profile_begin(#kernel, (void*)(stream)),                  \
   kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__), ::gh::profile_end((void*)(stream)))
#define GH_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

typedef unsigned long long u64;
typedef uint32_t u32;

// Stream bytes are one MSB-first bit string: a 32-bit window is the big-endian reading of 4 bytes.
__device__ __forceinline__ u32 be32(u32 little_endian_word) { return __byte_perm(little_endian_word, 0, 0x0123); }

// 128-bit read-only load (LDG.E.128.CONSTANT)
__device__ __forceinline__ uint4 ldg128(const uint4* p) { return __ldg(p); }

// Same load, but pinned to `dst`'s registers: as an opaque asm statement it cannot be merged with a sibling load
// on another path and copied over afterwards (a copy that waits for the load and defeats a software prefetch).
__device__ __forceinline__ void ldg128_into(uint4& dst, const uint4* p) {
#ifdef GH_EMUL
  dst = *p;
#else
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(dst.x), "=r"(dst.y), "=r"(dst.z), "=r"(dst.w)
               : "l"(p));
#endif
}

// Single-word flag|value messages between thread blocks: relaxed, GPU scope (served by L2, no system-scope
// round trip). One 64-bit word carries flag and value together, so no fence is needed around them.
__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
#ifdef GH_EMUL
  return __atomic_load_n(p, __ATOMIC_SEQ_CST);
#else
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v) {
#ifdef GH_EMUL
  __atomic_store_n(p, v, __ATOMIC_SEQ_CST);
#else
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// Ampere-style asynchronous 16-byte copy global -> shared (LDGSTS): the data never passes through registers, so a
// software prefetch cannot be undone by register copies that wait for the load.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef GH_EMUL
  memcpy(smem_dst, gsrc, 16);
#else
  const unsigned s = unsigned(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef GH_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all_but_one() {  // the older of two groups in flight has landed
#ifndef GH_EMUL
  asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}

// inclusive warp scan (Kogge-Stone over shuffles)
__device__ __forceinline__ u32 warp_inclusive_scan(u32 v, unsigned lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 up = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= unsigned(d)) v += up;
  }
  return v;
}

__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

}  // namespace gh
#endif
