// TEST INFRASTRUCTURE ONLY -- a minimal "CUDA on pthreads" shim.
//
// There is no GPU in the build container, and GPU box time is scarce.  This header lets the *same* .cu
// sources of golden-huffman_b200/csrc be compiled by g++ (-DGH_EMUL -include cuda_emul.h -x c++) into
// tests/emul/_build/libgh_emul.so so the kernels' logic (indexing, barriers, look-back, bit packing,
// self-synchronising decode) can be checked against the oracle on the CPU, under ASan/TSan if wanted,
// before a B200 is spent on them.  It is NOT a CPU fallback: nothing in the product loads this library,
// and the product's libgh_b200.so contains no host implementation of any kernel.
//
// Model: a launch runs its blocks one after another (in blockIdx order); the threads of a block are real
// OS threads; __syncthreads is a pthread barrier; warp collectives exchange through a per-warp buffer
// guarded by a per-warp barrier (all 32 lanes must take part, as with a full mask on hardware).
#ifndef GH_CUDA_EMUL_H_
#define GH_CUDA_EMUL_H_

#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace gh_emul {

constexpr int kMaxThreads = 1024;
constexpr int kMaxDynSmem = 232448;

struct Block {
  pthread_barrier_t bar;
  pthread_barrier_t warp_bar[kMaxThreads / 32];
  uint64_t warp_xchg[kMaxThreads / 32][32];
  std::atomic<int> vote;
  // named barriers (bar.sync id, count): initialised on first use in a launch, destroyed when the launch ends
  pthread_barrier_t named_bar[16];
  std::atomic<int> named_state[16];  // 0 = unused, 1 = being initialised, 2 = ready
  unsigned char dyn_smem[kMaxDynSmem] __attribute__((aligned(128)));
};

extern Block g_block;
extern dim3 g_blockDim, g_gridDim;
extern uint3 g_blockIdx;
extern thread_local uint3 t_threadIdx;
extern int g_sm_count;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline unsigned lane() { return t_threadIdx.x & 31u; }
inline unsigned warp() { return t_threadIdx.x >> 5; }
inline void warp_sync() { pthread_barrier_wait(&g_block.warp_bar[warp()]); }
inline void named_barrier(unsigned id, unsigned count) {
  int expected = 0;
  if (g_block.named_state[id].compare_exchange_strong(expected, 1)) {
    pthread_barrier_init(&g_block.named_bar[id], nullptr, count);
    g_block.named_state[id].store(2);
  } else {
    while (g_block.named_state[id].load() != 2) std::this_thread::yield();
  }
  pthread_barrier_wait(&g_block.named_bar[id]);
}

template <class T>
inline uint64_t to_bits(T v) {
  uint64_t b = 0;
  static_assert(sizeof(T) <= 8, "shuffle payload too wide");
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

// every lane publishes v, then reads lane `src` (or keeps its own value when src is out of range)
template <class T>
inline T exchange(T v, int src, bool valid) {
  uint64_t* x = g_block.warp_xchg[warp()];
  x[lane()] = to_bits(v);
  warp_sync();
  T r = valid ? from_bits<T>(x[src & 31]) : v;
  warp_sync();
  return r;
}

}  // namespace gh_emul

#define threadIdx (gh_emul::t_threadIdx)
#define blockIdx (gh_emul::g_blockIdx)
#define blockDim (gh_emul::g_blockDim)
#define gridDim (gh_emul::g_gridDim)
#define warpSize 32

// ---- barriers and warp collectives ---------------------------------------------------------------
static inline void __syncthreads() { pthread_barrier_wait(&gh_emul::g_block.bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { gh_emul::warp_sync(); }
static inline int __syncthreads_or(int pred) {
  if (pred) gh_emul::g_block.vote.fetch_or(1);
  __syncthreads();
  int r = gh_emul::g_block.vote.load();
  __syncthreads();
  if (threadIdx.x == 0) gh_emul::g_block.vote.store(0);
  __syncthreads();
  return r;
}
static inline int __syncthreads_count(int pred) {
  if (pred) gh_emul::g_block.vote.fetch_add(1);
  __syncthreads();
  int r = gh_emul::g_block.vote.load();
  __syncthreads();
  if (threadIdx.x == 0) gh_emul::g_block.vote.store(0);
  __syncthreads();
  return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return gh_emul::exchange(v, src, true); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int src = int(gh_emul::lane()) - int(d);
  return gh_emul::exchange(v, src, src >= 0);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int src = int(gh_emul::lane()) + int(d);
  return gh_emul::exchange(v, src, src < 32);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return gh_emul::exchange(v, int(gh_emul::lane()) ^ m, true); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t* x = gh_emul::g_block.warp_xchg[gh_emul::warp()];
  x[gh_emul::lane()] = pred ? 1 : 0;
  gh_emul::warp_sync();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= unsigned(x[i] & 1) << i;
  gh_emul::warp_sync();
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  uint64_t* x = gh_emul::g_block.warp_xchg[gh_emul::warp()];
  x[gh_emul::lane()] = v;
  gh_emul::warp_sync();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r += unsigned(x[i]);
  gh_emul::warp_sync();
  return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }

// ---- atomics ---------------------------------------------------------------------------------------
template <class T>
static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T atomicCAS(T* p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
template <class T>
static inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T>
static inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

// ---- integer intrinsics ----------------------------------------------------------------------------
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz(unsigned(x)) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  uint64_t v = (uint64_t(y) << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned b = unsigned(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;
    r |= b << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
  uint64_t v = (uint64_t(hi) << 32) | lo;
  return unsigned((v << (shift & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  uint64_t v = (uint64_t(hi) << 32) | lo;
  return unsigned(v >> (shift & 31));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return unsigned((uint64_t(a) * b) >> 32); }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// ---- the slice of the runtime API the library uses -------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
typedef int* cudaEvent_t;
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new int(0); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new int(0); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = size_t(1) << 32; *total_b = size_t(1) << 33; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = (a == cudaDevAttrMultiProcessorCount) ? gh_emul::g_sm_count : gh_emul::kMaxDynSmem;
  return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
#define cudaStreamNonBlocking 1

#endif  // GH_CUDA_EMUL_H_
