// Does a TMA box with elementStrides = 2 (the stride-2 convolutions of the generic kernel) cost more L2->SM time than the
// same number of delivered bytes fetched densely?  Three variants deliver 128 px x 64 ch x 2 planes (32 KB) per tap:
//   0  dense:    rank-5 (C, W, H, N, plane), box (64, 128, 1, 1, 2), stride-1 conv pattern
//   1  strided:  same map, box (64, 256, 2, 1, 2), elementStrides (1, 2, 2, 1, 1), stride-2 conv pattern (what ships)
//   2  parity:   rank-5 (C, wpar, W/2, H, N) per plane, box (64, 1, 128, 1, 1), two loads (hi, lo) per tap; the stride-2
//                pattern without elementStrides (w = 2*w2 + wpar)
// Every CTA walks its own output rows; per output row the 9 taps are fetched like the convolution does (L2 hits).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
struct P { int mode, slots, rows, Hout, N; };
__global__ void __launch_bounds__(96, 1) k(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmhi,
                                          const __grid_constant__ CUtensorMap tmlo, P p, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8], empty[8];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int r = 0; r < p.rows; ++r) {
      const int row = blockIdx.x * p.rows + r;          // output row (ho, n) of this CTA
      const int ho = row % p.Hout, n = (row / p.Hout) % p.N;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t dst = base + (uint32_t)s * 32768u, f = smem_u32(&full[s]);
          mbar_expect_tx(f, 32768u);
          if (p.mode == 0) tma_load_5d(dst, &tm, 0, kw - 1, ho + kh - 1, n, 0, f);
          else if (p.mode == 1) tma_load_5d(dst, &tm, 0, kw - 1, 2 * ho + kh - 1, n, 0, f);
          else {
            // w = 2*x + kw - 1: kw = 1 -> parity 0, w2 = x; kw = 0 -> parity 1, w2 = x - 1; kw = 2 -> parity 1, w2 = x
            const int wp = kw == 1 ? 0 : 1, w2 = kw == 0 ? -1 : 0;
            tma_load_5d(dst, &tmhi, 0, wp, w2, 2 * ho + kh - 1, n, f);
            tma_load_5d(dst + 16384u, &tmlo, 0, wp, w2, 2 * ho + kh - 1, n, f);
          }
          if (++s == p.slots) { s = 0; ph ^= 1u; }
        }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < p.rows * 9; ++i) {
      mbar_wait(smem_u32(&full[s]), ph);
      mbar_arrive(smem_u32(&empty[s]));
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    }
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int grid = 148, N = 8, H = 1024, W = 256, C = 64;   // stage-3 enc2.conv1 input, 8 windows
  const size_t plane = (size_t)N * H * W * C * 2;
  void* buf; cudaMalloc(&buf, 2 * plane); cudaMemset(buf, 0, 2 * plane);
  unsigned long long* d_out; cudaMalloc(&d_out, 8 * grid);
  CUtensorMap m1, m2, mhi, mlo;
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)plane};
    cuuint32_t box1[5] = {64, 128, 1, 1, 2}, es1[5] = {1, 1, 1, 1, 1};
    cuuint32_t box2[5] = {64, 256, 2, 1, 2}, es2[5] = {1, 2, 2, 1, 1};
    CUresult r1 = enc(&m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box1, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box2, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsp[5] = {(cuuint64_t)C, 2, (cuuint64_t)W / 2, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t stridesp[4] = {(cuuint64_t)C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t boxp[5] = {64, 1, 128, 1, 1};
    CUresult r3 = enc(&mhi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dimsp, stridesp, boxp, es1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r4 = enc(&mlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (char*)buf + plane, dimsp, stridesp, boxp, es1,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode results %d %d %d %d\n", (int)r1, (int)r2, (int)r3, (int)r4);
  }
  const int dyn = 201 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  for (int mode : {0, 1, 2})
    for (int slots : {2, 3, 6}) {
      const int Hout = mode == 0 ? H : H / 2;
      P p{mode, slots, 24, Hout, N};
      for (int rep = 0; rep < 2; ++rep) k<<<grid, 96, dyn>>>(mode == 1 ? m2 : m1, mhi, mlo, p, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<unsigned long long> h(grid);
      cudaMemcpy(h.data(), d_out, 8 * grid, cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i];
      cyc /= grid;
      printf("mode %d (%s)  slots %d  %7.1f cycles per 32 KB tap  %6.2f B/cycle/SM  %s\n", mode,
             mode == 0 ? "dense" : mode == 1 ? "elementStrides 2" : "parity dims", slots, cyc / (24 * 9), 24 * 9 * 32768.0 / cyc,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
      fflush(stdout);
    }
  return 0;
}
