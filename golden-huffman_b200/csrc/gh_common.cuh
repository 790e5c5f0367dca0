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

// One 32-byte sector per lane in a single request (LDG.E.ENL2.256, new with sm_100): lanes that stream their own
// subsequences touch one sector per request instead of half of one, which halves the L1 wavefronts of these
// fully divergent loads. `p` must be 32-byte aligned for the 256-bit form; a base that is only 16-byte aligned
// takes two 128-bit loads instead (block-uniform choice).
struct Unit8 {
  u32 w[8];
};
__device__ __forceinline__ Unit8 ldg_unit(const uint8_t* p, bool aligned32) {
  Unit8 u;
#ifdef GH_EMUL
  (void)aligned32;
  memcpy(u.w, p, 32);
#else
  if (aligned32) {
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u.w[0]), "=r"(u.w[1]), "=r"(u.w[2]), "=r"(u.w[3]), "=r"(u.w[4]), "=r"(u.w[5]), "=r"(u.w[6]),
                   "=r"(u.w[7])
                 : "l"(p));
  } else {
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u.w[0]), "=r"(u.w[1]), "=r"(u.w[2]), "=r"(u.w[3])
                 : "l"(p));
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u.w[4]), "=r"(u.w[5]), "=r"(u.w[6]), "=r"(u.w[7])
                 : "l"(p + 16));
  }
#endif
  return u;
}

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

// Bulk prefetch of `bytes` (multiple of 16, 16-byte aligned address) from DRAM into L2: one instruction, no registers,
// nothing to wait for (UBLKPF). A hint only.
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, u32 bytes) {
#ifdef GH_EMUL
  (void)gptr;
  (void)bytes;
#else
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
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

// Shared-memory loads by 32-bit shared-window address (LDS [reg + imm]): a table lookup whose byte offset was
// produced by a mask needs no further address arithmetic, and no generic-to-shared conversion is re-derived inside
// the loop (which is what indexing through a generic pointer compiles to).
#ifdef GH_EMUL
typedef const char* smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void* p) { return static_cast<const char*>(p); }
__device__ __forceinline__ u32 lds_u16(smem_addr_t base, u32 byte_off) {
  return *reinterpret_cast<const uint16_t*>(base + byte_off);
}
__device__ __forceinline__ uint2 lds_v2(smem_addr_t base, u32 byte_off) {
  return *reinterpret_cast<const uint2*>(base + byte_off);
}
__device__ __forceinline__ u32 lds_u32(smem_addr_t base, u32 byte_off) {
  return *reinterpret_cast<const u32*>(base + byte_off);
}
#else
typedef u32 smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void* p) {
  u32 a = u32(__cvta_generic_to_shared(p));
  asm volatile("" : "+r"(a));  // opaque: keeps the address in a register instead of re-deriving it at every use
  return a;
}
__device__ __forceinline__ u32 lds_u16(smem_addr_t base, u32 byte_off) {
  u32 v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
  return v;
}
__device__ __forceinline__ u32 lds_u32(smem_addr_t base, u32 byte_off) {
  u32 v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(smem_addr_t base, u32 byte_off) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + byte_off));
  return v;
}
#endif

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
