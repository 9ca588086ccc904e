// Host copy bandwidth probe (pageable -> pinned), the staging step of the C ABI's pageable path:
//   nvcc -O2 -o /tmp/host_copy_bw tools/host_copy_bw.cu && /tmp/host_copy_bw
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main() {
  const size_t total = (size_t)1 << 30, piece = (size_t)4 << 20;
  char *src = (char *)malloc(total), *dst = nullptr;
  cudaHostAlloc((void **)&dst, total, cudaHostAllocDefault);
  memset(src, 1, total);
  memset(dst, 2, total);
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  for (int nt : {1, 2, 4, 8, 12, 16}) {
    for (int mode = 0; mode < 2; mode++) {  // 0: every 4 MiB piece split over nt threads (spawned per piece); 1: nt long-lived threads, piece per thread
      auto t0 = std::chrono::steady_clock::now();
      if (mode == 0) {
        for (size_t off = 0; off < total; off += piece) {
          std::vector<std::thread> th;
          for (int k = 0; k < nt; k++) {
            const size_t a = piece * k / nt, b = piece * (k + 1) / nt;
            th.emplace_back([=] { memcpy(dst + off + a, src + off + a, b - a); });
          }
          for (auto &t : th) t.join();
        }
      } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; k++)
          th.emplace_back([=] { for (size_t off = piece * k; off < total; off += piece * nt) memcpy(dst + off, src + off, piece); });
        for (auto &t : th) t.join();
      }
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("threads %2d mode %d: %.1f GB/s\n", nt, mode, total / s / 1e9);
    }
  }
  return 0;
}
