// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands, M = 128 per CTA) as a function of N, the
// operand source (A from shared memory / TMEM), the swizzle mode and cta_group.  Operand contents are irrelevant for
// timing; shared memory holds zeros.  Prints cycles per MMA (clock64 of the issuing thread, first issue -> commit
// observed) next to the documented floor max(M,128)*N/(256*cta_group) per K=16 step.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (uint32_t n = 0; n < (1u << 26); ++n) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return true;
  }
  return false;
}

struct Cfg {
  int mode;     // 0: one SS MMA of width n0; 1: pair (n0 with A0, n1 with A1); 2: triple n0 (A0), n0 (A1), n0 (A0, other B);
                // 3: A from TMEM, width n0
  int n0, n1;
  int kb;       // channels per tile row: 64 (SWIZZLE_128B), 32 (SWIZZLE_64B), 16 (SWIZZLE_32B)
  int iters;    // tile passes; each pass issues kb/16 k-steps of the pattern
  int shift;    // 1: the A descriptor start is shifted by one tile row per k-step group (the row kernel's kw taps)
};

template <int CG>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(Cfg c, unsigned long long* out, int* fail) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t row_bytes = (uint32_t)c.kb * 2;
  const uint32_t layout = c.kb == 64 ? 2u : c.kb == 32 ? 4u : 6u;
  const uint32_t sbo = 8 * row_bytes;
  const uint32_t a0 = base, a1 = base + 40 * 1024, b0 = base + 80 * 1024, b1 = base + 120 * 1024;
  const uint32_t M = CG == 1 ? 128u : 256u;
  auto idesc = [&](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((M >> 4) << 24); };
  long long t0 = 0, t1 = 0;
  bool ok = true;
  if (warp == 1 && rank == 0) {
    const int ksteps = c.kb / 16;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const uint32_t sh = c.shift ? (uint32_t)(it % 3) * row_bytes : 0u;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t ko = (uint32_t)k * 32;
          const uint64_t dA0 = make_desc(a0 + sh + ko, sbo, layout), dA1 = make_desc(a1 + sh + ko, sbo, layout);
          const uint64_t dB0 = make_desc(b0 + ko, sbo, layout), dB1 = make_desc(b1 + ko, sbo, layout);
          if (c.mode == 0) {
            mma_ss<CG>(tmem, dA0, dB0, idesc(c.n0), 1u);
          } else if (c.mode == 1) {
            mma_ss<CG>(tmem, dA0, dB0, idesc(c.n0), 1u);
            mma_ss<CG>(tmem, dA1, dB0, idesc(c.n1), 1u);
          } else if (c.mode == 2) {
            mma_ss<CG>(tmem, dA0, dB0, idesc(c.n0), 1u);
            mma_ss<CG>(tmem, dA1, dB0, idesc(c.n0), 1u);
            mma_ss<CG>(tmem, dA0, dB1, idesc(c.n0), 1u);
          } else {
            mma_ts(tmem, tmem + 384 + (uint32_t)k * 8, dB0, idesc(c.n0), 1u);
          }
        }
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      ok = mbar_wait_bounded(smem_u32(&bar), 0);
      t1 = clock64();
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
      if (!ok) atomicAdd(fail, 1);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static void run(const char* label, int cg, Cfg c, int grid) {
  unsigned long long* d_out;
  int* d_fail;
  cudaMalloc(&d_out, sizeof(unsigned long long) * grid);
  cudaMalloc(&d_fail, sizeof(int));
  cudaMemset(d_out, 0, sizeof(unsigned long long) * grid);
  cudaMemset(d_fail, 0, sizeof(int));
  const int dyn = 161 * 1024 + 1024;
  if (cg == 1) {
    cudaFuncSetAttribute(umma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    for (int rep = 0; rep < 2; ++rep) umma_rate_kernel<1><<<grid, 128, dyn>>>(c, d_out, d_fail);
  } else {
    cudaFuncSetAttribute(umma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = dyn;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, umma_rate_kernel<2>, c, d_out, d_fail);
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> h(grid);
  int fail = 0;
  cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost);
  double sum = 0, mn = 1e30;
  int cnt = 0;
  for (int i = 0; i < grid; i += cg) {
    sum += (double)h[i];
    if ((double)h[i] < mn) mn = (double)h[i];
    ++cnt;
  }
  const int per_k = c.mode == 0 || c.mode == 3 ? 1 : c.mode == 1 ? 2 : 3;
  const double ksteps = (double)c.iters * (c.kb / 16);
  const double floor_k = c.mode == 0 || c.mode == 3 ? c.n0 / 2.0 : c.mode == 1 ? (c.n0 + c.n1) / 2.0 : 3 * c.n0 / 2.0;
  const double bytes_k = c.mode == 3 ? c.n0 * 32.0 / cg : per_k * 4096.0 + (c.mode == 1 ? (c.n0 + c.n1) : per_k * c.n0) * 32.0 / cg;
  printf("%-44s cg%d kb%d  cycles/k-step avg %8.1f min %8.1f  floor %6.1f  eff %.3f  smem B/cyc %.1f  %s%s\n", label, cg, c.kb,
         sum / cnt / ksteps, mn / ksteps, floor_k, floor_k / (sum / cnt / ksteps), bytes_k / (sum / cnt / ksteps),
         e == cudaSuccess ? "" : cudaGetErrorString(e), fail ? " TIMEOUT" : "");
  fflush(stdout);
  cudaFree(d_out);
  cudaFree(d_fail);
}

int main(int argc, char** argv) {
  const int grid = 148;
  const bool two = argc > 1 && atoi(argv[1]) == 2;
  char label[128];
  if (!two) {
    const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int kb : {64, 32}) {
      for (int n : ns) {
        snprintf(label, sizeof(label), "SS single N=%d", n);
        run(label, 1, Cfg{0, n, 0, kb, 512, 0}, grid);
      }
    }
    for (int n : ns) {
      snprintf(label, sizeof(label), "SS single N=%d, A start shifted per pass", n);
      run(label, 1, Cfg{0, n, 0, 32, 512, 1}, grid);
    }
    for (int n : ns) {
      snprintf(label, sizeof(label), "TS (A in TMEM) N=%d", n);
      run(label, 1, Cfg{3, n, 0, 64, 512, 0}, grid);
    }
    for (int bn : {16, 32, 64, 96, 128}) {
      snprintf(label, sizeof(label), "generic pair N=%d (A_hi) + N=%d (A_lo)", 2 * bn, bn);
      run(label, 1, Cfg{1, 2 * bn, bn, 64, 512, 0}, grid);
    }
    for (int n : {48, 96, 192}) {
      snprintf(label, sizeof(label), "row-kernel triple 3 x N=%d", n);
      run(label, 1, Cfg{2, n, 0, 32, 512, 1}, grid);
    }
  } else {
    const int ns[] = {32, 64, 96, 128, 192, 256};
    for (int n : ns) {
      snprintf(label, sizeof(label), "SS single M=256 (pair) N=%d", n);
      run(label, 2, Cfg{0, n, 0, 64, 512, 0}, grid);
    }
    for (int bn : {32, 64, 128}) {
      snprintf(label, sizeof(label), "pair: N=%d (A_hi) + N=%d (A_lo)", 2 * bn, bn);
      run(label, 2, Cfg{1, 2 * bn, bn, 64, 512, 0}, grid);
    }
  }
  return 0;
}
