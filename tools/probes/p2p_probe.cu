// Peer-memory micro-benchmarks on 2 GPUs of one box (single process): what a kernel can move over NVLink and how
// long flags take. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_probe p2p_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode 0: push (local load, remote store); 1: pull (remote load, local store); unroll U 16-byte accesses in flight
template <int U>
__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long n4, unsigned long long* stamps, int fence) {
  const long stride = (long)gridDim.x * blockDim.x;
  unsigned long long t0 = gtime();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n4) v[u] = __ldcg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n4) __stcg(dst + i + u * stride, v[u]);
  }
  unsigned long long t1 = gtime();
  __syncthreads();
  if (fence && threadIdx.x == 0) __threadfence_system();
  unsigned long long t2 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; }
}

// TMA-less bulk copy: one thread per CTA moves its chunk with cp.async.bulk global -> shared -> global? (not here)

__global__ void pong_kernel(volatile uint32_t* my_flag, volatile uint32_t* peer_flag, int iters) {
  for (int i = 1; i <= iters; ++i) {
    while (*my_flag < (uint32_t)i) {}
    __threadfence_system();
    *peer_flag = (uint32_t)i;
  }
}
__global__ void ping_kernel(volatile uint32_t* my_flag, volatile uint32_t* peer_flag, int iters, unsigned long long* out) {
  unsigned long long t0 = gtime();
  for (int i = 1; i <= iters; ++i) {
    *peer_flag = (uint32_t)i;
    __threadfence_system();
    while (*my_flag < (uint32_t)i) {}
  }
  out[0] = (gtime() - t0) / iters;
}

int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  printf("devices: %d\n", n);
  if (n < 2) return 0;
  int can = 0, perf = 0, atom = 0;
  cudaDeviceCanAccessPeer(&can, 0, 1);
  cudaDeviceGetP2PAttribute(&perf, cudaDevP2PAttrPerformanceRank, 0, 1);
  cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, 0, 1);
  printf("peer access 0->1: %d, performance rank %d, native atomics %d\n", can, perf, atom);
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
  const long big = 64l << 20;
  float4 *a0, *a1; unsigned long long *st0; uint32_t *f0, *f1;
  CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, big)); CK(cudaMalloc(&st0, 64)); CK(cudaMalloc(&f0, 256)); CK(cudaMemset(f0, 0, 256)); CK(cudaMemset(a0, 1, big));
  CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, big)); CK(cudaMalloc(&f1, 256)); CK(cudaMemset(f1, 0, 256)); CK(cudaMemset(a1, 2, big));
  CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const long sizes[3] = {1340l << 10, 335l << 10, 64l << 20};
  for (int si = 0; si < 3; ++si) {
    const long bytes = sizes[si], n4 = bytes / 16;
    float ms;
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); cudaMemcpyPeerAsync(a1, 1, a0, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("memcpyPeer %8ld KB: %7.2f us  %6.1f GB/s\n", bytes >> 10, ms * 1e3, bytes / ms / 1e6);
    for (int mode = 0; mode < 2; ++mode)
      for (int grid = 32; grid <= 128; grid *= 2) {
        const float4* s = mode == 0 ? a0 : a1; float4* d = mode == 0 ? a1 : a0;
        unsigned long long h[2];
        for (int U = 1; U <= 4; U *= 4) {
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (U == 1) copy_kernel<1><<<grid, 256>>>(s, d, n4, st0, 1); else copy_kernel<4><<<grid, 256>>>(s, d, n4, st0, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
          }
          cudaEventElapsedTime(&ms, e0, e1);
          cudaMemcpy(h, st0, 16, cudaMemcpyDeviceToHost);
          printf("  %s %8ld KB grid %3d x256 unroll %d: kernel %7.2f us (%6.1f GB/s)  cta0: issue %5.2f us, fence %5.2f us\n", mode == 0 ? "push" : "pull",
                 bytes >> 10, grid, U, ms * 1e3, bytes / ms / 1e6, h[0] / 1e3, h[1] / 1e3);
        }
      }
  }
  // flag round trip
  CK(cudaSetDevice(1)); pong_kernel<<<1, 1>>>(f1, f0, 1000);
  CK(cudaSetDevice(0)); ping_kernel<<<1, 1>>>(f0, f1, 1000, st0);
  CK(cudaDeviceSynchronize());
  unsigned long long rt; cudaMemcpy(&rt, st0, 8, cudaMemcpyDeviceToHost);
  printf("flag round trip (store + fence.sys + remote poll, both ways): %.2f us\n", rt / 1e3);
  CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
  return 0;
}
