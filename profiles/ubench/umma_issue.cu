#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ void umma_w(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d_tmem),
      "r"(a_lo32), "r"(b_lo32), "r"(hi32), "r"(idesc)
      : "memory");
}
// predicated form: every lane executes the asm, one elected lane issues
__device__ __forceinline__ void umma_wp(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32, uint32_t idesc, uint32_t pe) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d_tmem),
      "r"(a_lo32), "r"(b_lo32), "r"(hi32), "r"(idesc), "r"(pe)
      : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (uint32_t n = 0; n < (1u << 26); ++n) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// VAR 0: elect block around the unrolled row (as conv_tc_rows today, compile-time offsets)
// VAR 1: predicated MMAs, no branch
// VAR 2: elect block, but one MMA per (kw,ks) only (rate of the UTCHMMA itself with N stacked x3 wider)
// VAR 3: as VAR 0 plus one tcgen05.commit per row (to an mbarrier nobody waits on), VAR 4: commit every second row
template <int VAR, int BN, int NROWS>
__global__ void __launch_bounds__(128, 1) issue_kernel(int iters, int nslots, unsigned long long* out, int* fail, int random_data) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t rowbar[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&rowbar[i])));
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
  {
      uint32_t x = (uint32_t)i * 2654435761u + 12345u;
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) {
        x = x * 1664525u + 1013904223u;
        // two bf16 values with exponent in [0x3c, 0x40) (magnitudes 2^-7 .. 2): finite, both signs, random mantissas
        uint32_t lo = (x & 0x807fu) | (((x >> 8) & 3u) + 0x3cu) << 7, hi = ((x >> 16) & 0x807fu) | (((x >> 26) & 3u) + 0x3cu) << 7;
        w[j] = random_data ? (lo | (hi << 16)) : 0u;
      }
      reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t KB = 32, row_bytes = KB * 2;
  constexpr uint32_t a_plane = 9216, a_slot = 2 * a_plane;
  constexpr uint32_t b_kw_bytes = 2 * 3 * BN * KB * 2, b3_plane = 3 * BN * row_bytes;
  const uint32_t dhi = desc_hi(8 * row_bytes, 4);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * BN) >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t b_base = base + 8 * a_slot;
  if (warp == 1) {
    long long t0 = clock64();
    int as = 0;
    const uint32_t pe = elect_one() ? 1u : 0u;
    for (int it = 0; it < iters; ++it) {
      for (int r = 0; r < NROWS; ++r) {
        const uint32_t a_hi = desc_lo(base + (uint32_t)as * a_slot);
        const uint32_t a_lo = a_hi + (a_plane >> 4);
        const uint32_t d = tmem + (uint32_t)(r & 3) * BN;
        const uint32_t bsrc = desc_lo(b_base);
        if (VAR == 0 || VAR == 3 || VAR == 4) {
          if (elect_one()) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t ao = (kw * row_bytes + ks * 32) >> 4, bo = (kw * b_kw_bytes + ks * 32) >> 4;
                umma_w(d, a_hi + ao, bsrc + bo, dhi, idesc);
                umma_w(d, a_lo + ao, bsrc + bo, dhi, idesc);
                umma_w(d, a_hi + ao, bsrc + bo + (b3_plane >> 4), dhi, idesc);
              }
            if (VAR == 3 || (VAR == 4 && (r & 1)))
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&rowbar[as])) : "memory");
          }
          __syncwarp();
        } else if (VAR == 1) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t ao = (kw * row_bytes + ks * 32) >> 4, bo = (kw * b_kw_bytes + ks * 32) >> 4;
              umma_wp(d, a_hi + ao, bsrc + bo, dhi, idesc, pe);
              umma_wp(d, a_lo + ao, bsrc + bo, dhi, idesc, pe);
              umma_wp(d, a_hi + ao, bsrc + bo + (b3_plane >> 4), dhi, idesc, pe);
            }
        } else {
          if (elect_one()) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t ao = (kw * row_bytes + ks * 32) >> 4, bo = (kw * b_kw_bytes + ks * 32) >> 4;
                umma_w(d, a_hi + ao, bsrc + bo, dhi, idesc);
              }
          }
          __syncwarp();
        }
        if (++as == nslots) as = 0;
      }
    }
    if (elect_one()) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
      out[blockIdx.x] = (unsigned long long)(clock64() - t0);
      if (!ok) atomicAdd(fail, 1);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}
template <int VAR, int BN>
void run(const char* label, int random_data = 0) {
  unsigned long long* d_out; int* d_fail;
  const int grid = 148, iters = 64, NROWS = 10;
  cudaMalloc(&d_out, 8 * grid); cudaMalloc(&d_fail, 4); cudaMemset(d_fail, 0, 4);
  const int dyn = 201 * 1024 + 1024;
  cudaFuncSetAttribute(issue_kernel<VAR, BN, NROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  for (int rep = 0; rep < 2; ++rep) issue_kernel<VAR, BN, NROWS><<<grid, 128, dyn>>>(iters, 6, d_out, d_fail, random_data);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148]; int fail = 0;
  cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost); cudaMemcpy(&fail, d_fail, 4, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
  const double rows = (double)iters * NROWS, mmas = rows * (VAR == 2 ? 6 : 18);
  printf("%-50s %s BN=%d  cycles/row %.1f  cycles/MMA %.1f  floor/MMA %.1f %s%s\n", label, random_data ? "random" : "zeros ", BN, s / grid / rows, s / grid / mmas, 3 * BN / 2.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e), fail ? " TIMEOUT" : "");
}
int main() {
  run<0, 16>("elect block, 18 MMAs/row unrolled");
  run<0, 32>("elect block, 18 MMAs/row unrolled");
  run<0, 16>("elect block, 18 MMAs/row unrolled", 1);
  run<0, 32>("elect block, 18 MMAs/row unrolled", 1);
  run<3, 16>("same + tcgen05.commit after every row");
  run<3, 32>("same + tcgen05.commit after every row");
  run<4, 16>("same + tcgen05.commit after every 2nd row");
  run<4, 32>("same + tcgen05.commit after every 2nd row");
  run<2, 16>("elect block, 6 MMAs/row");
  run<2, 32>("elect block, 6 MMAs/row");
  return 0;
}
