// Micro-benchmark: the producer / MMA-issuer handshake skeleton of the row kernel (conv_tc_rows.cu) without any data
// movement.  Warp 0 stands in for the TMA producer (wait 'empty', arrive 'full'), warp 1 issues the 18 tcgen05.mma of a
// row and commits to 'empty'.  Reports cycles per row for several variants of the issuer loop, to separate the tensor
// pipe time (1008 cycles per N=96 row, umma_issue.cu) from the per-row loop overhead of the issuing warp.
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
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) { return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29); }
__device__ __forceinline__ void umma_w(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32, uint32_t idesc) {
  asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d_tmem),
               "r"(a_lo32), "r"(b_lo32), "r"(hi32), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

template <int BN>
__device__ __forceinline__ void issue18(uint32_t d, uint32_t a_hi, uint32_t bsrc, uint32_t dhi, uint32_t idesc) {
  constexpr uint32_t row_bytes = 64, a_plane = 9216, b_kw_bytes = 2 * 3 * BN * 64, b3_plane = 3 * BN * row_bytes;
  const uint32_t a_lo = a_hi + (a_plane >> 4);
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t ao = (kw * row_bytes + ks * 32) >> 4, bo = (kw * b_kw_bytes + ks * 32) >> 4;
      umma_w(d, a_hi + ao, bsrc + bo, dhi, idesc);
      umma_w(d, a_lo + ao, bsrc + bo, dhi, idesc);
      umma_w(d, a_hi + ao, bsrc + bo + (b3_plane >> 4), dhi, idesc);
    }
}

// VAR 0: as the row kernel (wait full[s], elect, 18 MMAs, commit empty[s]) ; VAR 1: no MMAs (handshake only)
// VAR 2: two rows per iteration (two waits, 36 MMAs, two commits) ; VAR 3: VAR 0 without waiting on 'full' (issuer never
// blocks on the producer: loop overhead of the issuer alone) ; VAR 4: VAR 2 with four rows per iteration
template <int VAR, int BN>
__global__ void __launch_bounds__(128, 1) pipe_kernel(int rows, int nslots, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(&done), 1);
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
  constexpr uint32_t a_slot = 18432;
  const uint32_t dhi = desc_hi(8 * 64, 4);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * BN) >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t b_base = base + 8 * a_slot;
  if (warp == 0) {   // producer stand-in
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < rows; ++i) {
      mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
      if (elect_one()) mbar_arrive(smem_u32(&full[s]));
      __syncwarp();
      if (++s == nslots) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    constexpr int STEP = VAR == 2 ? 2 : VAR == 4 ? 4 : 1;
    for (int i = 0; i < rows; i += STEP) {
      int ss[STEP];
#pragma unroll
      for (int j = 0; j < STEP; ++j) {
        ss[j] = s;
        if (VAR != 3) mbar_wait(smem_u32(&full[s]), ph);
        if (++s == nslots) { s = 0; ph ^= 1u; }
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bsrc = desc_lo(b_base);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < STEP; ++j) {
          const uint32_t a_hi = desc_lo(base + (uint32_t)ss[j] * a_slot);
          const uint32_t d = tmem + (uint32_t)((i + j) & 3) * BN;
          if (VAR != 1) issue18<BN>(d, a_hi, bsrc, dhi, idesc);
          umma_commit(smem_u32(&empty[ss[j]]));
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      umma_commit(smem_u32(&done));
      mbar_wait(smem_u32(&done), 0);
      out[blockIdx.x] = (unsigned long long)(clock64() - t0);
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
void run(const char* label, int nslots) {
  unsigned long long* d_out;
  const int grid = 148, rows = 1024;
  cudaMalloc(&d_out, 8 * grid);
  const int dyn = 201 * 1024 + 1024;
  cudaFuncSetAttribute(pipe_kernel<VAR, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  for (int rep = 0; rep < 2; ++rep) pipe_kernel<VAR, BN><<<grid, 128, dyn>>>(rows, nslots, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
  printf("%-62s BN=%d slots %d  cycles/row %7.1f %s\n", label, BN, nslots, s / grid / rows, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d_out);
}
int main() {
  for (int ns : {2, 4, 8}) run<0, 32>("row kernel skeleton: wait, elect, 18 MMAs, commit", ns);
  for (int ns : {4, 8}) run<1, 32>("handshake only (no MMAs)", ns);
  for (int ns : {4, 8}) run<2, 32>("two rows per iteration", ns);
  for (int ns : {4, 8}) run<4, 32>("four rows per iteration", ns);
  run<3, 32>("issuer never waits on the producer", 8);
  for (int ns : {4, 8}) run<0, 16>("row kernel skeleton: wait, elect, 18 MMAs, commit", ns);
  for (int ns : {4, 8}) run<2, 16>("two rows per iteration", ns);
  for (int ns : {4, 8}) run<4, 16>("four rows per iteration", ns);
  return 0;
}
