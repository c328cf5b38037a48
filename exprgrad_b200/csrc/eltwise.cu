// HBM-bound streaming kernels: fill. (Elementwise expression programs live in program_kernels.cu.)
#include "egb_internal.hpp"

namespace egb {
namespace {
__global__ void fill_u32_kernel(uint32_t* __restrict__ dst, uint32_t value, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (n >> 2) : 0;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const uint4 v4 = make_uint4(value, value, value, value);
  for (size_t j = i; j < n4; j += stride) d4[j] = v4;
  for (size_t j = (n4 << 2) + i; j < n; j += stride) dst[j] = value;
}
}  // namespace

void launch_fill_u32(Context& ctx, uint32_t* dst, uint32_t value, size_t n, cudaStream_t st) {
  if (n == 0) return;
  size_t blocks = (n / 4 + 255) / 256 + 1;
  const size_t cap = (size_t)ctx.sm_count * 8;
  if (blocks > cap) blocks = cap;
  {
    Launch l(ctx, KC_FILL, st);
    fill_u32_kernel<<<(int)blocks, 256, 0, st>>>(dst, value, n);
  }
  EGB_CUDA(cudaGetLastError());
}
}  // namespace egb
