// TEST INFRASTRUCTURE ONLY -- the launcher of the "CUDA on pthreads" shim (see cuda_emul.h).
#include "cuda_emul.h"

namespace gh_emul {

Block g_block;
dim3 g_blockDim, g_gridDim;
uint3 g_blockIdx;
thread_local uint3 t_threadIdx;
int g_sm_count = 2;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || nthreads > unsigned(kMaxThreads) || (nthreads % 32) != 0 || block.y != 1 || block.z != 1 ||
      smem > size_t(kMaxDynSmem)) {
    fprintf(stderr, "gh_emul: unsupported launch shape %u threads, %zu B smem\n", nthreads, smem);
    abort();
  }
  g_blockDim = block;
  g_gridDim = grid;
  pthread_barrier_init(&g_block.bar, nullptr, nthreads);
  for (unsigned w = 0; w < nthreads / 32; w++) pthread_barrier_init(&g_block.warp_bar[w], nullptr, 32);
  g_block.vote.store(0);
  for (int i = 0; i < 16; i++) g_block.named_state[i].store(0);

  // One pool of `nthreads` workers walks the grid: all workers execute block b, meet at a barrier, move on.
  // Blocks therefore run strictly one after another, in blockIdx order.
  const uint64_t nblocks = uint64_t(grid.x) * grid.y * grid.z;
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; t++) {
    pool.emplace_back([&, t]() {
      t_threadIdx = uint3{t, 0, 0};
      for (uint64_t b = 0; b < nblocks; b++) {
        if (t == 0) {
          g_blockIdx.x = unsigned(b % grid.x);
          g_blockIdx.y = unsigned((b / grid.x) % grid.y);
          g_blockIdx.z = unsigned(b / (uint64_t(grid.x) * grid.y));
        }
        pthread_barrier_wait(&g_block.bar);
        body();
        pthread_barrier_wait(&g_block.bar);
      }
    });
  }
  for (auto& th : pool) th.join();
  for (int i = 0; i < 16; i++)
    if (g_block.named_state[i].load() == 2) pthread_barrier_destroy(&g_block.named_bar[i]);
  pthread_barrier_destroy(&g_block.bar);
  for (unsigned w = 0; w < nthreads / 32; w++) pthread_barrier_destroy(&g_block.warp_bar[w]);
}

}  // namespace gh_emul

extern "C" void gh_emul_set_sm_count(int n) { gh_emul::g_sm_count = n > 0 ? n : 1; }
