// Allocation latency on the GPU box: cudaMalloc / cudaFree of several sizes, with and without other live allocations, and the
// virtual-memory route (cuMemAddressReserve once, cuMemCreate + cuMemMap + cuMemSetAccess per growth step).
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/micro/alloc_latency.cu -o /tmp/alloc_latency -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaFree(0);
  const size_t sizes[] = {1 << 20, 16 << 20, 64 << 20, 256 << 20, (size_t)1 << 30};
  for (int round = 0; round < 2; ++round) {
    for (size_t sz : sizes) {
      void* p = nullptr;
      double t0 = now_us();
      cudaMalloc(&p, sz);
      double t1 = now_us();
      cudaMemset(p, 0xFF, sz);
      cudaDeviceSynchronize();
      double t2 = now_us();
      cudaFree(p);
      double t3 = now_us();
      std::printf("round %d size %5zu MB: cudaMalloc %8.0f us, memset+sync %8.0f us, cudaFree %8.0f us\n", round, sz >> 20, t1 - t0, t2 - t1, t3 - t2);
    }
    if (round == 0) {   // keep 20 GB live, pinned mapped memory and a second stream, like the library does
      void* big = nullptr; cudaMalloc(&big, (size_t)20 << 30); cudaMemset(big, 0, (size_t)20 << 30);
      void* hp = nullptr; cudaHostAlloc(&hp, 1 << 20, cudaHostAllocMapped);
      cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
      cudaDeviceSynchronize();
    }
  }
  // virtual memory route
  CUdevice dev; cuDeviceGet(&dev, 0);
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  size_t gran = 0; cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  CUdeviceptr va = 0; const size_t va_size = (size_t)64 << 30;
  double t0 = now_us();
  CUresult r = cuMemAddressReserve(&va, va_size, 0, 0, 0);
  double t1 = now_us();
  std::printf("granularity %zu, cuMemAddressReserve(64 GB) rc %d: %.0f us\n", gran, (int)r, t1 - t0);
  size_t mapped = 0;
  for (size_t step : {(size_t)16 << 20, (size_t)64 << 20, (size_t)256 << 20, (size_t)1 << 30}) {
    CUmemGenericAllocationHandle h;
    double a0 = now_us();
    CUresult r1 = cuMemCreate(&h, step, &prop, 0);
    double a1 = now_us();
    CUresult r2 = cuMemMap(va + mapped, step, 0, h, 0);
    CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUresult r3 = cuMemSetAccess(va + mapped, step, &ad, 1);
    double a2 = now_us();
    cudaMemset((void*)(va + mapped), 1, step); cudaDeviceSynchronize();
    double a3 = now_us();
    std::printf("grow by %5zu MB: cuMemCreate %8.0f us (rc %d), map+access %8.0f us (rc %d %d), memset+sync %8.0f us\n", step >> 20, a1 - a0, (int)r1, a2 - a1, (int)r2, (int)r3, a3 - a2);
    mapped += step;
  }
  return 0;
}
