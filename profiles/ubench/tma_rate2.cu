// TMA issue-pattern micro-benchmark (follow-up of tma_rate.cu): K boxes per mbarrier / ring slot.
// Each box = 130 px x 32 ch x 1 row x 1 plane (8320 B) of a rank-5 NHWC split-bf16 tensor, SWIZZLE_64B.
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
struct P { int slots, nslots_total, K, mode, H, N; };
// box index b of this CTA -> (row h, image n, plane) without divisions: b>>1 = linear row, H is a power of two
__device__ __forceinline__ void box_coord(const P& p, int cta, int b, int& h, int& n, int& pl) {
  const int row = cta * 512 + (b >> 1);   // every CTA streams its own 512 rows (no L2 hot spot shared between SMs)
  h = row & (p.H - 1);
  n = (row >> 9) & (p.N - 1);
  pl = b & 1;
}
// mode 0: producer = one thread (wait empty, expect_tx, K loads); mode 1: K producer threads of one warp issue one load each
// (lane k loads box k; lane 0 does expect_tx first); mode 2: as 0 but the prefetch.tensormap is issued before the loop
__global__ void __launch_bounds__(96, 1) k(const __grid_constant__ CUtensorMap tm, P p, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[32], empty[32];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.mode == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
  }
  __syncthreads();
  const uint32_t slot_bytes = (uint32_t)p.K * 8320u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.nslots_total; ++i) {
      mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
      const int row0 = i * p.K;
      const uint32_t dst = base + (uint32_t)s * ((slot_bytes + 1023u) & ~1023u);
      if (p.mode == 1) {
        if (lane == 0) mbar_expect_tx(smem_u32(&full[s]), slot_bytes);
        __syncwarp();
        if (lane < p.K) {
          int h, n, pl;
          box_coord(p, blockIdx.x, row0 + lane, h, n, pl);
          tma_load_5d(dst + (uint32_t)lane * 8320u, &tm, 0, -1, h, n, pl, smem_u32(&full[s]));
        }
        __syncwarp();
      } else if (p.mode == 3) {
        // one box group per barrier, issued by a DIFFERENT lane each time (rotating over 8 lanes)
        if (lane == (i & 7)) {
          mbar_expect_tx(smem_u32(&full[s]), slot_bytes);
          for (int kk = 0; kk < p.K; ++kk) {
            int h, n, pl;
            box_coord(p, blockIdx.x, row0 + kk, h, n, pl);
            tma_load_5d(dst + (uint32_t)kk * 8320u, &tm, 0, -1, h, n, pl, smem_u32(&full[s]));
          }
        }
      } else if (lane == 0) {
        mbar_expect_tx(smem_u32(&full[s]), slot_bytes);
        for (int kk = 0; kk < p.K; ++kk) {
          int h, n, pl;
          box_coord(p, blockIdx.x, row0 + kk, h, n, pl);
          tma_load_5d(dst + (uint32_t)kk * 8320u, &tm, 0, -1, h, n, pl, smem_u32(&full[s]));
        }
      }
      __syncwarp();
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < p.nslots_total; ++i) {
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
  const int grid = 148, N = 16, H = 512, W = 256, C = 64;
  const size_t plane = (size_t)N * H * W * C * 2;
  void* buf; cudaMalloc(&buf, 2 * plane); cudaMemset(buf, 0, 2 * plane);
  unsigned long long* d_out; cudaMalloc(&d_out, 8 * grid);
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)plane};
  cuuint32_t box[5] = {32, 130, 1, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int dyn = 201 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  for (int mode : {0, 1, 3})
    for (int K : {1, 2, 4, 8, 16})
      for (int slots : {2, 4, 8}) {
        if ((size_t)slots * ((K * 8320 + 1023) / 1024 * 1024) > 200 * 1024) continue;
        P p{slots, 1024 / K, K, mode, H, N};
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 96, dyn>>>(m, p, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<unsigned long long> h(grid);
        cudaMemcpy(h.data(), d_out, 8 * grid, cudaMemcpyDeviceToHost);
        double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i];
        cyc /= grid;
        printf("mode %d  K=%2d boxes/barrier  slots %d  %7.1f cycles/box  %6.2f B/cycle/SM  %s\n", mode, K, slots, cyc / 1024, 1024.0 * 8320 / cyc,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
