// HBM-bound streaming kernels: fill. (Elementwise expression programs live in program_kernels.cu.)
#include "egb_internal.hpp"

namespace egb {
namespace {
__global__ void fill_u32_kernel(uint32_t* __restrict__ dst, uint32_t value, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (n >> 2) : 0;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const uint4 v4 = make_uint4(value, value, value, value);
  for (size_t j = i; j < n4; j += stride) d4[j] = v4;
  for (size_t j = (n4 << 2) + i; j < n; j += stride) dst[j] = value;
}
// U(lo, hi) fill for TensorRandom tensors (model.nim:286-294, 310-314: refilled on every call).
// Counter-based: element i of tensor t in call c is a hash of (seed, c, t, i), so a run is reproducible and
// every random tensor of a call draws its own sequence (the reference calls newRandTensor per tensor).
__global__ void fill_uniform_kernel(float* __restrict__ dst, size_t n, float lo, float hi, uint64_t seed,
                                    uint64_t counter) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t z = seed * 0x9e3779b97f4a7c15ull + counter * 0xd1342543de82ef95ull + i + 1;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z = z ^ (z >> 31);
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // [0, 1)
    dst[i] = lo + (hi - lo) * u;
  }
}
}  // namespace

void launch_fill_uniform(Context& ctx, float* dst, size_t n, float lo, float hi, uint64_t seed, uint64_t counter,
                         uint64_t tensor, cudaStream_t st) {
  if (n == 0) return;
  // the tensor id selects an independent stream: folded into the seed with a full-avalanche mix
  {
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (tensor + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    seed = z ^ (z >> 31);
  }
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)ctx.sm_count * 8;
  if (blocks > cap) blocks = cap;
  {
    Launch l(ctx, KC_FILL, st);
    launch_kernel(ctx, fill_uniform_kernel, dim3((int)blocks), dim3(256), 0, st, dst, n, lo, hi, seed, counter);
  }
  EGB_CUDA(cudaGetLastError());
}

void launch_fill_u32(Context& ctx, uint32_t* dst, uint32_t value, size_t n, cudaStream_t st) {
  if (n == 0) return;
  size_t blocks = (n / 4 + 255) / 256 + 1;
  const size_t cap = (size_t)ctx.sm_count * 8;
  if (blocks > cap) blocks = cap;
  {
    Launch l(ctx, KC_FILL, st);
    launch_kernel(ctx, fill_u32_kernel, dim3((int)blocks), dim3(256), 0, st, dst, value, n);
  }
  EGB_CUDA(cudaGetLastError());
}
}  // namespace egb
