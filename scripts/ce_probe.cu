// Which stream operations wait behind a large H2D copy running on another stream?
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <vector>
__global__ void k(unsigned *p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 3u + 1u; }
int main() {
  const size_t N = 40u << 20; // 160 MB
  unsigned *h, *d, *a, *b, *c2, *d2; void *tmp; unsigned *hs_pinned; unsigned hs_page[2];
  cudaHostAlloc(&h, N * 4, cudaHostAllocDefault); cudaMalloc(&d, N * 4);
  const int M = 4 << 20;
  cudaMalloc(&a, M * 4); cudaMalloc(&b, M * 4); cudaMalloc(&c2, M * 4); cudaMalloc(&d2, M * 4);
  cudaHostAlloc(&hs_pinned, 64, cudaHostAllocDefault);
  size_t tb = 0; cub::DoubleBuffer<unsigned> dk(a, b), dv(c2, d2);
  cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, M, 0, 17); cudaMalloc(&tmp, tb);
  cudaStream_t sa, sb; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  const char *names[] = {"kernel", "memsetAsync 64B", "D2D 16MB", "D2H 8B pinned", "D2H 8B pageable", "cub sort 4M", "kernel2"};
  cudaEvent_t e0, ev[8]; cudaEventCreate(&e0); for (auto &e : ev) cudaEventCreate(&e);
  for (int rep = 0; rep < 3; ++rep) {
    for (int withcopy = 0; withcopy < 2; ++withcopy) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, sb);
      if (withcopy) cudaMemcpyAsync(d, h, N * 4, cudaMemcpyHostToDevice, sa);
      k<<<M / 256, 256, 0, sb>>>(a, M); cudaEventRecord(ev[0], sb);
      cudaMemsetAsync(b, 0, 64, sb); cudaEventRecord(ev[1], sb);
      cudaMemcpyAsync(c2, a, M * 4, cudaMemcpyDeviceToDevice, sb); cudaEventRecord(ev[2], sb);
      cudaMemcpyAsync(hs_pinned, a, 8, cudaMemcpyDeviceToHost, sb); cudaEventRecord(ev[3], sb);
      cudaMemcpyAsync(hs_page, a, 8, cudaMemcpyDeviceToHost, sb); cudaEventRecord(ev[4], sb);
      cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, M, 0, 17, sb); cudaEventRecord(ev[5], sb);
      k<<<M / 256, 256, 0, sb>>>(a, M); cudaEventRecord(ev[6], sb);
      cudaDeviceSynchronize();
      if (rep == 2) {
        printf("--- %s a 160 MB H2D copy in flight on another stream\n", withcopy ? "WITH" : "without");
        for (int i = 0; i < 7; ++i) { float ms; cudaEventElapsedTime(&ms, e0, ev[i]); printf("  %-18s done at %7.3f ms\n", names[i], ms); }
      }
    }
  }
  // chunked copy with events; stream B waits on the event after the first 32 MB
  {
    const size_t chunk = 4u << 20; // 16 MB in elements of 4 B
    std::vector<cudaEvent_t> ce(10);
    for (auto &e : ce) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, sb);
      for (int i = 0; i < 10; ++i) {
        cudaMemcpyAsync(d + i * chunk, h + i * chunk, chunk * 4, cudaMemcpyHostToDevice, sa);
        cudaEventRecord(ce[i], sa);
      }
      cudaStreamWaitEvent(sb, ce[1], 0);
      k<<<M / 256, 256, 0, sb>>>(a, M); cudaEventRecord(ev[0], sb);
      cudaStreamWaitEvent(sb, ce[9], 0);
      k<<<M / 256, 256, 0, sb>>>(a, M); cudaEventRecord(ev[1], sb);
      cudaDeviceSynchronize();
      float m0, m1; cudaEventElapsedTime(&m0, e0, ev[0]); cudaEventElapsedTime(&m1, e0, ev[1]);
      if (rep == 2) printf("chunked copy: kernel after chunk 1 (32 MB) done at %.3f ms; after last chunk at %.3f ms\n", m0, m1);
    }
  }
  return 0;
}
