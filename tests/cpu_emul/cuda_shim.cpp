// cuda_shim.cpp — TEST INFRASTRUCTURE ONLY (see cuda_shim.h).
#include "cuda_shim.h"

#include <condition_variable>
#include <map>
#include <mutex>

#include <barrier>
#include <memory>
#include <thread>
#include <vector>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
char* fc_emul_smem = nullptr;
static std::barrier<>* g_barrier = nullptr;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barriers;

void __syncthreads() { g_barrier->arrive_and_wait(); }
void __syncwarp() { g_warp_barriers[(threadIdx.y * blockDim.x + threadIdx.x) / 32]->arrive_and_wait(); }

void fc_emul_launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
  const unsigned nt = block.x * block.y;
  std::vector<char> shared(smem + 64);
  fc_emul_smem = shared.data();
  std::barrier<> bar((std::ptrdiff_t)nt);
  g_barrier = &bar;
  g_warp_barriers.clear();
  for (unsigned wi = 0; wi < (nt + 31) / 32; ++wi) {
    const unsigned lanes = (wi * 32 + 32 <= nt) ? 32 : nt - wi * 32;
    g_warp_barriers.emplace_back(new std::barrier<>((std::ptrdiff_t)lanes));
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([=, &body]() {
      blockDim = block;
      gridDim = grid;
      threadIdx = dim3(t % block.x, t / block.x, 0);
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = dim3(bx, by, bz);
            body();
            g_barrier->arrive_and_wait();  // block boundary: shared memory is reused by the next block
          }
    });
  }
  for (auto& t : th) t.join();
  g_barrier = nullptr;
  fc_emul_smem = nullptr;
}

// ---- named barrier stand-in (see fc_kernels.cuh)
namespace {
std::mutex g_named_m;
std::map<std::pair<int, int>, std::unique_ptr<std::barrier<>>> g_named;
}  // namespace
void fc_emul_named_barrier(int id, int count) {
  std::barrier<>* bar;
  {
    std::lock_guard<std::mutex> l(g_named_m);
    auto& slot = g_named[{id, count}];
    if (!slot) slot.reset(new std::barrier<>(count));
    bar = slot.get();
  }
  bar->arrive_and_wait();
}
