// {I, dx, dy} gradient packing of a pyramid level with the intensity tile staged in shared memory by TMA
// (cp.async.bulk.tensor.2d + mbarrier) -- the tile-staged variant of k_pixelinfo (pba_kernels.cu), kept side by side with
// it for the A/B that settles whether TMA staging pays on this path (profiles/r02_ab.md, dpba_debug_pixelinfo_ab).
//
// Replaces the same reference function as k_pixelinfo: calculate_pixelinfo<1> (src/features/src/calculate_pixelinfo.cpp:
// 340-374, central differences, one-sided at the image border) feeding PixelMap's storage (features/camera/pixel_map.hpp).
// Output records are bit-identical to k_pixelinfo's: 32-byte texel pairs {I, dx, dy, 0}(x), {I, dx, dy, 0}(min(x + 1, W - 1)).
//
// One CTA = a tile of TW x TH pixels.  An elected thread arms an mbarrier with the byte count of the (TW + 8) x (TH + 2)
// box that starts 4 pixels left of / 1 row above the tile (inner box extent a multiple of 16 bytes; out-of-image elements
// are zero-filled by the TMA unit and never used: the border pixels take the one-sided formulas) and issues the bulk tensor
// copy; everybody waits on the barrier's phase, then reads the stencil from shared memory.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <tuple>

#include "pba_internal.h"

namespace pba {
namespace {

constexpr int TW = 64, TH = 8;            // pixels per CTA
constexpr int BW = TW + 8, BH = TH + 2;   // TMA box: 4 pixels of margin left / right (16-byte multiple), 1 row above / below
constexpr int XO = 4, YO = 1;             // position of the tile's first pixel inside the box

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(TW* TH) k_pixelinfo_tma(const __grid_constant__ CUtensorMap tmap, float4* __restrict__ dst,
                                                        int W, int H) {
  __shared__ __align__(128) float tile[BH][BW];
  __shared__ __align__(8) uint64_t bar;
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(BW * BH * sizeof(float)))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(&tile[0][0])),
        "l"(&tmap), "r"(x0 - XO), "r"(y0 - YO), "r"(bar_a)
        : "memory");
  }
  {  // everybody waits for phase 0 of the barrier: the box has landed
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar_a), "r"(0u)
          : "memory");
    }
  }
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  // the same arithmetic as pixelinfo_at (pba_kernels.cu), reading the stencil from the staged tile
  auto at = [&](int xx, int yy) { return tile[yy - y0 + YO][xx - x0 + XO]; };
  auto info = [&](int xx) {
    const float c = at(xx, y);
    float dx;
    if (xx == 0) dx = 1.0f * (at(1, y) - c);
    else if (xx == W - 1) dx = 1.0f * (c - at(xx - 1, y));
    else dx = 0.5f * (at(xx + 1, y) - at(xx - 1, y));
    const int yu = y == 0 ? y : y - 1, yb = y == H - 1 ? y : y + 1;
    const float dy = ((y == 0 || y == H - 1) ? 1.0f : 0.5f) * (at(xx, yb) - at(xx, yu));
    return make_float4(c, dx, dy, 0.f);
  };
  dst[2 * ((size_t)y * W + x)] = info(x);
  dst[2 * ((size_t)y * W + x) + 1] = info(min(x + 1, W - 1));
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

}  // namespace

// false: the tensor map could not be made (W * 4 not a multiple of 16, unaligned plane, no driver entry point); nothing
// was launched and the caller decides (the product path reports the error, it does not fall back silently)
bool launch_pixelinfo_tma(const float* I, float4* dst, int W, int H, cudaStream_t s) {
  EncodeFn enc = encode_fn();
  if (!enc || (W * sizeof(float)) % 16 != 0 || (reinterpret_cast<uintptr_t>(I) & 15) != 0) return false;
  static std::mutex mu;
  static std::map<std::tuple<const float*, int, int>, CUtensorMap> cache;
  CUtensorMap tm;
  {
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(I, W, H);
    auto it = cache.find(key);
    if (it == cache.end()) {
      const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
      const cuuint64_t strides[1] = {(cuuint64_t)W * sizeof(float)};
      const cuuint32_t box[2] = {BW, BH};
      const cuuint32_t estr[2] = {1, 1};
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(I), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
      if (cache.size() > 64) cache.clear();
      cache[key] = tm;
    } else {
      tm = it->second;
    }
  }
  dim3 g((W + TW - 1) / TW, (H + TH - 1) / TH);
  add_launches(1);
  k_pixelinfo_tma<<<g, TW * TH, 0, s>>>(tm, dst, W, H);
  return true;
}

}  // namespace pba
