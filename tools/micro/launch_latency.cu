// Micro-benchmark (B200 box): where does the per-pass host<->device latency go?
//  (1) launch -> mapped-memory flag visible to the host, for small / large kernel parameter blocks and grids
//  (2) persistent kernel: host writes a flag in mapped memory -> kernel (polling over PCIe) acknowledges
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/launch_latency tools/micro/launch_latency.cu
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <atomic>
#include <algorithm>
#include <vector>
struct Small { unsigned long long* flag; unsigned long long seq; };
struct Big { unsigned long long* flag; unsigned long long seq; char pad[900]; };
template <class P> __global__ void k_flag(const __grid_constant__ P p) {
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(p.flag) = p.seq;
}
__global__ void k_persist(volatile unsigned long long* req, volatile unsigned long long* ack, int n, int payload_words, volatile unsigned* payload, unsigned* sink) {
  for (int i = 1; i <= n; ++i) {
    if (threadIdx.x == 0) { while (*req != (unsigned long long)i) {} }
    __syncwarp();
    unsigned v = 0;
    if (payload_words) for (int w = threadIdx.x; w < payload_words; w += 32) v += payload[w];
    if (payload_words) sink[threadIdx.x] = v;
    __syncwarp();
    if (threadIdx.x == 0) *ack = (unsigned long long)i;
  }
}
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class P> void run_launch(const char* name, int grid, unsigned long long* h_flag, unsigned long long* d_flag, cudaStream_t st) {
  std::vector<double> t, tl;
  P p; memset(&p, 0, sizeof(p)); p.flag = d_flag;
  for (int i = 1; i <= 300; ++i) {
    p.seq = i;
    double t0 = now_us();
    k_flag<P><<<grid, 128, 0, st>>>(p);
    double t1 = now_us();
    while (*reinterpret_cast<volatile unsigned long long*>(h_flag) != (unsigned long long)i) {}
    double t2 = now_us();
    if (i > 50) { t.push_back(t2 - t0); tl.push_back(t1 - t0); }
    cudaStreamSynchronize(st);
  }
  std::sort(t.begin(), t.end()); std::sort(tl.begin(), tl.end());
  printf("%-28s grid %5d params %4zu B: launch call %.2f us (median), call->flag visible %.2f us (median) p90 %.2f\n", name, grid, sizeof(P), tl[tl.size()/2], t[t.size()/2], t[t.size()*9/10]);
}
int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  unsigned long long *h, *d; cudaHostAlloc(&h, 4096, cudaHostAllocMapped); memset(h, 0, 4096); cudaHostGetDevicePointer((void**)&d, h, 0);
  run_launch<Small>("flag kernel, small params", 1, h, d, st); *h = 0;
  run_launch<Small>("flag kernel, small params", 1024, h, d, st); *h = 0;
  run_launch<Big>("flag kernel, big params", 1, h, d, st); *h = 0;
  run_launch<Big>("flag kernel, big params", 1024, h, d, st); *h = 0;
  // back-to-back without stream sync between (like the pass loop: launch as soon as the flag is seen)
  {
    std::vector<double> t; Big p; memset(&p, 0, sizeof(p)); p.flag = d;
    for (int i = 1; i <= 300; ++i) { p.seq = i; double t0 = now_us(); k_flag<Big><<<1024, 128, 0, st>>>(p); while (*reinterpret_cast<volatile unsigned long long*>(h) != (unsigned long long)i) {} if (i > 50) t.push_back(now_us() - t0); }
    std::sort(t.begin(), t.end()); printf("big params, 1024 CTAs, relaunch as soon as flag seen (no sync): %.2f us median\n", t[t.size()/2]);
    cudaStreamSynchronize(st); *h = 0;
  }
  for (int words : {0, 64}) {
    volatile unsigned long long* req = h + 8; volatile unsigned long long* ack = h + 16; *req = 0; *ack = 0;
    unsigned* sink; cudaMalloc(&sink, 256);
    const int n = 300;
    k_persist<<<1, 32, 0, st>>>(d + 8, d + 16, n, words, (volatile unsigned*)(d + 32), sink);
    std::vector<double> t;
    for (int i = 1; i <= n; ++i) {
      double t0 = now_us();
      *req = i;
      while (*ack != (unsigned long long)i) {}
      if (i > 50) t.push_back(now_us() - t0);
      double w = now_us(); while (now_us() - w < 5.0) {}   // let the kernel go back to polling
    }
    cudaStreamSynchronize(st);
    std::sort(t.begin(), t.end());
    printf("persistent handshake (host flag -> kernel poll over PCIe -> %d payload words -> ack visible): %.2f us median, p90 %.2f\n", words, t[t.size()/2], t[t.size()*9/10]);
  }
  return 0;
}
