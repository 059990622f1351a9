// Micro-benchmark: TMA (cp.async.bulk.tensor.5d) delivery rate per SM for the row kernel's operand boxes as a function
// of the box shape (bytes per instruction) and the number of boxes in flight.  One CTA per SM streams row boxes of an
// NHWC split-bf16 tensor [2 planes][N][H][W][C] through a ring of shared-memory slots; a consumer thread frees a slot as
// soon as it has landed (no compute), so the measured rate is the TMA path alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (uint32_t n = 0; n < (1u << 24); ++n) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
struct P {
  int slots, rows_total;      // ring depth, rows streamed per CTA
  int instr_per_slot;         // 1: one box covers the slot; 2: one box per plane
  int rows_per_slot;          // image rows per slot
  int slot_bytes, box_bytes;
  int H, W, N, chunks;
  int rank, kc;   // tensor-map rank (3: one image plane only), channels per box
  int w_start;   // -1: the box hangs one pixel out of the image (zero fill = conv padding), 1: fully inside
};

__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, P p, unsigned long long* out, int* fail) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[24], empty[24];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nslots_total = p.rows_total / p.rows_per_slot;
  const int rows_per_img = p.H;
  if (threadIdx.x == 0) {   // producer
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < nslots_total; ++i) {
      if (!mbar_wait_bounded(smem_u32(&empty[s]), ph ^ 1u)) { atomicAdd(fail, 1); break; }
      // rows are spread over the tensor so that different CTAs read different data (no L2 sharing between SMs)
      const long long row = ((long long)blockIdx.x * nslots_total + i) * p.rows_per_slot;
      const int cc = (int)((row / (rows_per_img * p.N)) % p.chunks);
      const int n = (int)((row / rows_per_img) % p.N);
      const int h = (int)(row % rows_per_img);
      const uint32_t dst = base + (uint32_t)(s * p.slot_bytes);
      mbar_expect_tx(smem_u32(&full[s]), (uint32_t)(p.box_bytes * p.instr_per_slot));
      if (p.rank == 3) {
        tma_load_3d(dst, &tm, cc * p.kc, p.w_start, h, smem_u32(&full[s]));
      } else if (p.instr_per_slot == 1) {
        tma_load_5d(dst, &tm, cc * 32, p.w_start, h, n, 0, smem_u32(&full[s]));
      } else {
        tma_load_5d(dst, &tm, cc * 32, p.w_start, h, n, 0, smem_u32(&full[s]));
        tma_load_5d(dst + p.box_bytes, &tm, cc * 32, p.w_start, h, n, 1, smem_u32(&full[s]));
      }
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    }
    out[2 * blockIdx.x] = (unsigned long long)(clock64() - t0);
  } else if (threadIdx.x == 32) {   // consumer
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < nslots_total; ++i) {
      if (!mbar_wait_bounded(smem_u32(&full[s]), ph)) { atomicAdd(fail, 1); break; }
      mbar_arrive(smem_u32(&empty[s]));
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    }
    out[2 * blockIdx.x + 1] = (unsigned long long)(clock64() - t0);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int grid = 148;
  unsigned long long* d_out; int* d_fail;
  cudaMalloc(&d_out, 16 * grid); cudaMalloc(&d_fail, 4);
  // tensors: C channels per pixel; big = 8 x 512 x 256 pixels (beyond L2 for C = 64: 2 x 134 MB), small fits L2
  struct Shape { const char* name; int N, H, W, C; };
  Shape shapes[] = {{"L2-resident (1x512x256x64)", 1, 512, 256, 64}, {"DRAM (16x512x256x64)", 16, 512, 256, 64}};
  for (const Shape& sh : shapes) {
    const size_t plane = (size_t)sh.N * sh.H * sh.W * sh.C * 2;
    void* buf; cudaMalloc(&buf, 2 * plane); cudaMemset(buf, 0, 2 * plane);
    struct Var { const char* name; int rank, kc, planes_in_box, rows_in_box, instr_per_slot; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2p; };
    Var vars[] = {{"rank5 32ch x130px x1 row  x1 plane SW64 (2 instr/row)", 5, 32, 1, 1, 2, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank5 32ch x130px x4 rows x1 plane SW64", 5, 32, 1, 4, 1, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 32ch x130px x1 row  SW64", 3, 32, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 32ch x130px x4 rows SW64", 3, 32, 1, 4, 1, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 32ch x130px x1 row  no swizzle", 3, 32, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 32ch x130px x1 row  SW64 no L2 promotion", 3, 32, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE},
                  {"rank3 32ch x128px x1 row  SW64", 3, 32, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 64ch x130px x1 row  SW128 (contiguous 16.6 KB)", 3, 64, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank3 64ch x130px x2 rows SW128", 3, 64, 1, 2, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
                  {"rank5 64ch x130px x1 row x2 planes SW128", 5, 64, 2, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}};
    int vi = 0;
    for (const Var& v : vars) {
      const int px = vi == 6 ? 128 : 130;
      ++vi;
      CUtensorMap m;
      cuuint64_t dims[5] = {(cuuint64_t)sh.C, (cuuint64_t)sh.W, (cuuint64_t)sh.H * (v.rank == 3 ? sh.N : 1), (cuuint64_t)sh.N, 2};
      cuuint64_t strides[4] = {(cuuint64_t)sh.C * 2, (cuuint64_t)sh.W * sh.C * 2, (cuuint64_t)sh.H * sh.W * sh.C * 2, (cuuint64_t)plane};
      cuuint32_t box[5] = {(cuuint32_t)v.kc, (cuuint32_t)px, (cuuint32_t)v.rows_in_box, 1, (cuuint32_t)v.planes_in_box};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, v.rank, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       v.sw, v.l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      for (int wst : {-1}) for (int slots : {1, 4, 101, 102, 104, 108, 116}) {
        P p;
        const bool burst = slots > 100;
        if (burst) slots -= 100;
        p.w_start = wst; p.rank = v.rank; p.kc = v.kc;
        p.instr_per_slot = v.instr_per_slot; p.rows_per_slot = v.rows_in_box;
        p.box_bytes = px * v.kc * 2 * v.rows_in_box * v.planes_in_box;
        p.slot_bytes = (p.box_bytes * v.instr_per_slot + 1023) / 1024 * 1024;
        if (slots * p.slot_bytes > 200 * 1024) continue;
        p.slots = slots; p.rows_total = burst ? slots * v.rows_in_box : 512;
        p.H = sh.H * (v.rank == 3 ? sh.N : 1); p.W = sh.W; p.N = v.rank == 3 ? 1 : sh.N; p.chunks = sh.C / v.kc;
        cudaMemset(d_fail, 0, 4);
        const int dyn = 201 * 1024;
        cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        for (int rep = 0; rep < 2; ++rep) tma_rate_kernel<<<grid, 64, dyn>>>(m, p, d_out, d_fail);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<unsigned long long> h(2 * grid); int fail = 0;
        cudaMemcpy(h.data(), d_out, 16 * grid, cudaMemcpyDeviceToHost); cudaMemcpy(&fail, d_fail, 4, cudaMemcpyDeviceToHost);
        double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[2 * i + 1];
        cyc /= grid;
        const double ninstr = (double)p.rows_total / v.rows_in_box * v.instr_per_slot;
        const double bytes = ninstr * p.box_bytes;
        printf("%-28s %-56s %s %2d  %7.1f cycles/instr  %6.0f B/instr  %6.2f B/cycle/SM %s%s\n", sh.name, v.name, burst ? "burst" : "slots", slots,
               cyc / ninstr, (double)p.box_bytes, bytes / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e), fail ? " TIMEOUT" : "");
        fflush(stdout);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
