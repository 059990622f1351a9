// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a with fused folded-BN bias + activation.
//
// Replaces the reference's Conv2DBNActiv (lib/layers.py:8-26) for the dense 3x3 / 1x1 / strided / dilated
// layers that carry 99.8 % of the FLOPs.  GEMM view per CTA tile:
//     D[128 pixels][BN couts] += A[128 pixels][K] * B[BN][K]^T ,   K = taps * CinPad
//   * A is never materialised: for every (tap, 64/32/16-channel chunk) ONE TMA tiled load fetches the
//     shifted (dilated / strided, zero-filled out of bounds = conv padding) pixel box of the NHWC
//     split-bf16 activation, both planes (hi, lo) in one instruction, straight into the 128B/64B/32B
//     swizzled K-major layout tcgen05 consumes.
//   * B (BN-folded weights, split into bf16 hi/lo once at load time) arrives by TMA the same way.
//   * One elected thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32 in TMEM).
//     Three passes per k-step, hi*hi + lo*hi + hi*lo, give a ~2^-16 relative product error - the
//     precision the 1e-3 mask gate needs (single-pass bf16/fp16 measurably fails it).
//   * Warp-specialised persistent kernel: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc),
//     warps 2-5 epilogue (tcgen05.ld -> bias -> ReLU/LeakyReLU -> split to bf16 hi/lo -> channel slice of
//     the destination NHWC buffer, which is how concats are written in place).  smem ring of 3-6 stages,
//     two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdio.h>

#include <map>
#include <tuple>

#include "engine.h"
#include "tc_common.cuh"
#include "tc_plan.h"

namespace vr {

static constexpr int kMaxStages = 8;
static constexpr int kThreads = 192;

struct TcParams {
  int N, Ho, Wo, Wt, Ht, Nt, tiles_w, tiles_h, m_tiles, n_tiles;
  int stride, pad_h, pad_w, dil_h, dil_w, KW;
  int KB, cchunks, SUBS, total_sub, CinPadTC, BN, Cout, act, stages;
  int a_sub_bytes, b_sub_bytes, a_plane_bytes, b_plane_bytes;
  uint32_t idesc;    // N = BN
  uint32_t idesc2;   // N = 2*BN: one MMA against the stacked [B_hi ; B_lo] planes
  uint32_t sbo_bytes, layout_type;
  bf16* out_hi;
  bf16* out_lo;
  int64_t osn, osh;
  int osw;
  const float* bias;
  int tmem_cols;
};

// ------------------------------------------------------------------------------------------------
// KB: channels per operand sub-tile (64 / 32 / 16 = SWIZZLE_128B / 64B / 32B); a pipeline stage holds 64 / KB sub-tiles.
// The producer and the MMA issuer are single elected lanes running their whole loop nests with compile-time operand
// strides (see conv_tc_rows.cu: the tensor pipe queues only a few MMAs, so scalar work between the last MMA of a stage
// and the first of the next one is a bubble; ~500 cycles per stage were measured with the per-stage election and
// run-time strides of the first version).
template <int KB>
__global__ void __launch_bounds__(kThreads, 1)
    conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  constexpr int SUBS = 64 / KB;
  constexpr int kSteps = KB / 16;
  constexpr uint32_t kAPlane = 128 * KB * 2;     // one plane of an A sub-tile (128 pixels x KB channels)
  constexpr uint32_t kASub = 2 * kAPlane;        // hi + lo
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[256];   // folded-BN bias of every N tile, staged once (a global load per use stalled the epilogue)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_sub_bytes = (uint32_t)p.b_sub_bytes;
  const uint32_t stage_bytes = (uint32_t)SUBS * (kASub + b_sub_bytes);
  constexpr uint32_t b_region = (uint32_t)SUBS * kASub;   // B sub-tiles follow the A sub-tiles inside a stage
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int num_iters = (p.total_sub + SUBS - 1) / SUBS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tfull[a]), 1);
      mbar_init(smem_u32(&bar_tempty[a]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += blockDim.x) bias_s[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane runs the whole loop nest) =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.Wt;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.Ht;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.Nt;
        const int wbase = w0 * p.stride - p.pad_w, hbase = h0 * p.stride - p.pad_h;
        const int nrow = nt * p.BN;
        int cc = 0, kw = 0, kh = 0, sub = 0;   // (tap, channel chunk) of the next sub-tile, advanced without divisions
        for (int it = 0; it < num_iters; ++it) {
          mbar_wait(empty0 + (uint32_t)stage * 8u, phase ^ 1u);
          const int nsub = min(SUBS, p.total_sub - sub);
          const uint32_t full = full0 + (uint32_t)stage * 8u;
          const uint32_t sbase = smem_base + (uint32_t)stage * stage_bytes;
          mbar_expect_tx(full, (uint32_t)nsub * (kASub + b_sub_bytes));
#pragma unroll
          for (int j = 0; j < SUBS; ++j) {
            if (j < nsub) {
              tma_load_5d(sbase + (uint32_t)j * kASub, &tmA, cc * KB, wbase + kw * p.dil_w, hbase + kh * p.dil_h, n0, 0, full);
              tma_load_3d(sbase + b_region + (uint32_t)j * b_sub_bytes, &tmB, (kh * p.KW + kw) * p.CinPadTC + cc * KB, nrow, 0,
                          full);
              ++sub;
              if (++cc == p.cchunks) {
                cc = 0;
                if (++kw == p.KW) {
                  kw = 0;
                  ++kh;
                }
              }
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane runs the whole loop nest) =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t dhi = desc_hi(p.sbo_bytes, p.layout_type);
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      const uint32_t b_sub16 = b_sub_bytes >> 4;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * p.BN);   // [D1 | D2], see the issue loop
        mbar_wait(full0 + (uint32_t)stage * 8u, phase);
        int sub = 0;
        for (int it = 0; it < num_iters; ++it) {
          const int nsub = min(SUBS, p.total_sub - sub);
          sub += nsub;
          const uint32_t sdesc = desc_lo(smem_base + (uint32_t)stage * stage_bytes);
          const uint32_t empty = empty0 + (uint32_t)stage * 8u;
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
          // The hi and lo weight planes are adjacent in the stage ([BN rows hi][BN rows lo], same pitch), so A_hi meets
          // both in ONE N = 2*BN instruction: D1 += A_hi*B_hi, D2 += A_hi*B_lo.  A second N = BN instruction adds
          // A_lo*B_hi to D1.  Two instructions and one shared-memory pass over A_hi per k-step instead of three; the
          // epilogue adds D1 + D2.  The wait for the NEXT stage is issued before the last pair of this one, so that it
          // overlaps the products still queued in the tensor pipe.
#pragma unroll
          for (int j = 0; j < SUBS; ++j) {
            if (j < nsub) {
              const uint32_t a_hi = sdesc + (uint32_t)((j * kASub) >> 4);
              const uint32_t a_lo = a_hi + (kAPlane >> 4);
              const uint32_t b_hi = sdesc + (b_region >> 4) + (uint32_t)j * b_sub16;
#pragma unroll
              for (int k = 0; k < kSteps; ++k) {
                if (j == nsub - 1 && k == kSteps - 1 && it + 1 < num_iters) mbar_wait(full0 + (uint32_t)stage * 8u, phase);
                const uint32_t acc_flag = (it | j | k) ? 1u : 0u;
                umma_bf16_w(d_tmem, a_hi + 2u * k, b_hi + 2u * k, dhi, p.idesc2, acc_flag);
                umma_bf16_w(d_tmem, a_lo + 2u * k, b_hi + 2u * k, dhi, p.idesc, 1u);
              }
            }
          }
          umma_commit(empty);   // frees the slot once the MMAs have read it
        }
        umma_commit(smem_u32(&bar_tfull[acc]));       // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5 <-> TMEM lane quarters 2,3,0,1) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float slope = p.act == ACT_RELU ? 0.f : p.act == ACT_LEAKY ? 0.01f : 1.f;
    const int dw = row % p.Wt;
    const int dh = (row / p.Wt) % p.Ht;
    const int dn = row / (p.Wt * p.Ht);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int w0 = (mt % p.tiles_w) * p.Wt;
      const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.Ht;
      const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.Nt;
      const int n = n0 + dn;
      const bool valid = n < p.N;
      const int64_t obase = (int64_t)n * p.osn + (int64_t)(h0 + dh) * p.osh + (int64_t)(w0 + dw) * p.osw;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * 2 * p.BN) + ((uint32_t)(q * 32) << 16);
      int c0 = 0;
      for (; c0 + 32 <= p.BN; c0 += 32) {
        float v[32], v2[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        tmem_ld32(t_row + (uint32_t)(p.BN + c0), v2);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
        if (c0 + 32 >= p.BN) {   // all of this warp's TMEM reads are done: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        }
        if (valid) epilogue_store<2>(v, bias_s, nt * p.BN + c0, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
      }
      if (c0 < p.BN) {   // BN is a multiple of 16: one trailing 16-column group
        float v[16], v2[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        tmem_ld16(t_row + (uint32_t)(p.BN + c0), v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += v2[i];
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        if (valid) epilogue_store<1>(v, bias_s, nt * p.BN + c0, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
int g_tc_debug[8] = {0, 0, 0, 0, 0, 1, 1, 0};   // [5] fused decoder upsample, [6] zero-weight group skipping: on

static constexpr int kMaxDevices = 64;
const TcDevice& tc_device() {
  static TcDevice table[kMaxDevices];
  static TcDevice none;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return none;
  TcDevice& d = table[dev];
  if (!d.ok) {
    if (cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return none;
    cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem - 2048);
    cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem - 2048);
    cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem - 2048);
    tc_rows_set_attributes(d.max_smem);
    d.ok = true;
  }
  return d;
}

EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int KB) {
  return KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

uint16_t tc_f2bf(float f) {   // round-to-nearest-even, same as __float2bfloat16_rn for finite values
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
float tc_bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct TileGeom {
  int Wt, Ht, Nt;
  bool ok;
};
static TileGeom tile_geom(int Ho, int Wo) {
  TileGeom g{0, 0, 0, false};
  if (Wo <= 0 || Ho <= 0) return g;
  if (Wo >= 128) {
    if (Wo % 128) return g;
    g.Wt = 128; g.Ht = 1; g.Nt = 1;
  } else {
    if (128 % Wo) return g;
    g.Wt = Wo;
    g.Ht = 128 / Wo < Ho ? 128 / Wo : Ho;
    if (Ho % g.Ht) return g;
    if ((128 / Wo) % g.Ht) return g;
    g.Nt = 128 / (g.Wt * g.Ht);
  }
  g.ok = true;
  return g;
}

bool tc_supported(const ConvLayer& L, const ActView& in, const ActView& out) {
  if (!L.tc) return false;
  if (!(L.k == 1 || L.k == 3) || !(L.stride == 1 || L.stride == 2)) return false;
  TileGeom g = tile_geom(out.H, out.W);
  if (!g.ok) return false;
  if (g.Wt * L.stride > 256 || g.Ht * L.stride > 256) return false;
  if (in.sw % 8 || in.sh % 8 || in.sn % 8) return false;
  if ((reinterpret_cast<uintptr_t>(in.hi) | reinterpret_cast<uintptr_t>(in.lo)) & 15) return false;
  if (in.C <= 0 || in.N <= 0) return false;
  if ((in.H - 1) / L.stride + 1 != out.H || (in.W - 1) / L.stride + 1 != out.W) return false;
  return tc_encode_fn() != nullptr;
}

bool tc_can_fuse_upsample(const ConvLayer& L, const ActView& in, const ActView& out, const ActView& up_src) {
  if (!L.tc || g_tc_debug[5] != 1) return false;
  const TcConv& tc = *L.tc;
  if (!tc_rows_supported(L, tc, in, out)) return false;
  return up_src.C % 32 == 0 && up_src.C <= tc.rows.CinPadR && up_src.H * 2 == in.H && up_src.W * 2 == in.W &&
         up_src.sw % 8 == 0;
}

bool tc_prepare(ConvLayer& L, std::string& err, std::vector<void*>& allocs) {
  if (!(L.k == 1 || L.k == 3) || L.Cout < 4) return true;   // stays on the CUDA-core kernel
  auto tc = std::make_shared<TcConv>();
  tc->taps = L.k * L.k;
  tc->CinPadTC = round_up(L.CinPad, 16);
  tc->KB = tc->CinPadTC % 64 == 0 ? 64 : tc->CinPadTC % 32 == 0 ? 32 : 16;
  tc->cchunks = tc->CinPadTC / tc->KB;
  tc->SUBS = 64 / tc->KB;
  tc->Ktot = tc->taps * tc->CinPadTC;
  tc->CoutPadN = round_up(L.Cout, 16);
  tc->n_tiles = ceil_div(tc->CoutPadN, 128);
  tc->BN = round_up(ceil_div(tc->CoutPadN, tc->n_tiles), 16);
  const int rows = tc->n_tiles * tc->BN;
  std::vector<uint16_t> planes((size_t)2 * rows * tc->Ktot, 0);
  for (int co = 0; co < L.Cout; ++co)
    for (int t = 0; t < tc->taps; ++t)
      for (int ci = 0; ci < L.CinPad; ++ci) {
        const float w = L.w_host[((size_t)t * L.CinPad + ci) * L.CoutPad + co];
        const uint16_t hi = tc_f2bf(w);
        const uint16_t lo = tc_f2bf(w - tc_bf2f(hi));
        const size_t k = (size_t)t * tc->CinPadTC + ci;
        planes[(size_t)co * tc->Ktot + k] = hi;
        planes[((size_t)rows + co) * tc->Ktot + k] = lo;
      }
  std::vector<float> bias((size_t)rows, 0.f);
  for (int co = 0; co < L.Cout; ++co) bias[(size_t)co] = L.bias_host[(size_t)co];
  void* dw = nullptr;
  void* db = nullptr;
  if (cudaMalloc(&dw, planes.size() * 2) != cudaSuccess || cudaMalloc(&db, bias.size() * 4) != cudaSuccess) {
    err = "cudaMalloc failed while packing tensor-core weights for " + L.name;
    return false;
  }
  allocs.push_back(dw);
  allocs.push_back(db);
  cudaMemcpy(dw, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  tc->w_planes = (bf16*)dw;
  tc->bias = (float*)db;
  EncodeTiledFn enc = tc_encode_fn();
  if (!enc) {
    err = "cuTensorMapEncodeTiled is not available from the driver";
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)tc->Ktot, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)tc->Ktot * 2, (cuuint64_t)rows * tc->Ktot * 2};
  cuuint32_t box[3] = {(cuuint32_t)tc->KB, (cuuint32_t)tc->BN, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tc->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dw, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(tc->KB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    err = "cuTensorMapEncodeTiled(weights) failed for " + L.name + " code " + std::to_string((int)r);
    return false;
  }
  if (!tc_rows_prepare(L, *tc, err, allocs)) return false;
  L.tc = tc;
  return true;
}

cudaError_t tc_launch(ConvLayer& L, const ActView& in, const ActView& out, cudaStream_t s, std::string& err,
                      const ActView* up_src, const ActView* extra) {
  TcConv& tc = *L.tc;
  if (tc_rows_supported(L, tc, in, out)) return tc_rows_launch(L, tc, in, out, s, err, up_src, extra);
  if (up_src || extra) {
    err = "tc_launch: fused upsample is only implemented in the row-streaming kernel";
    return cudaErrorInvalidValue;
  }
  const TileGeom g = tile_geom(out.H, out.W);
  auto key = std::make_tuple((const void*)in.hi, (const void*)in.lo, in.N, in.H, in.W, in.C);
  auto it = tc.map_a.find(key);
  if (it == tc.map_a.end()) {
    CUtensorMap m;
    cuuint64_t dims[5] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N, 2};
    const int64_t plane = (const char*)in.lo - (const char*)in.hi;
    if (plane <= 0 || plane % 16) {
      err = "tc_launch: hi/lo planes must be 16-byte aligned with lo after hi";
      return cudaErrorInvalidValue;
    }
    cuuint64_t strides[4] = {(cuuint64_t)in.sw * 2, (cuuint64_t)in.sh * 2, (cuuint64_t)in.sn * 2, (cuuint64_t)plane};
    cuuint32_t box[5] = {(cuuint32_t)tc.KB, (cuuint32_t)(g.Wt * L.stride), (cuuint32_t)(g.Ht * L.stride),
                         (cuuint32_t)g.Nt, 2};
    cuuint32_t es[5] = {1, (cuuint32_t)L.stride, (cuuint32_t)L.stride, 1, 1};
    CUresult r = tc_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)in.hi, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(tc.KB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      err = "cuTensorMapEncodeTiled(activations) failed for " + L.name + " code " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
    it = tc.map_a.emplace(key, m).first;
  }
  TcParams p;
  p.N = out.N; p.Ho = out.H; p.Wo = out.W;
  p.Wt = g.Wt; p.Ht = g.Ht; p.Nt = g.Nt;
  p.tiles_w = out.W / g.Wt; p.tiles_h = out.H / g.Ht;
  p.m_tiles = p.tiles_w * p.tiles_h * ceil_div(out.N, g.Nt);
  p.n_tiles = tc.n_tiles;
  p.stride = L.stride; p.dil_h = L.dil_h; p.dil_w = L.dil_w;
  p.pad_h = L.dil_h * (L.k / 2); p.pad_w = L.dil_w * (L.k / 2);
  p.KW = L.k;
  p.KB = tc.KB; p.cchunks = tc.cchunks; p.SUBS = tc.SUBS; p.total_sub = tc.taps * tc.cchunks;
  p.CinPadTC = tc.CinPadTC; p.BN = tc.BN; p.Cout = L.Cout; p.act = L.act;
  p.a_plane_bytes = 128 * tc.KB * 2;
  p.b_plane_bytes = tc.BN * tc.KB * 2;
  p.a_sub_bytes = 2 * p.a_plane_bytes;
  p.b_sub_bytes = 2 * p.b_plane_bytes;
  const int stage_bytes = tc.SUBS * (p.a_sub_bytes + p.b_sub_bytes);
  const TcDevice& dv = tc_device();
  if (!dv.ok) {
    err = "tc_launch: cannot query the current device";
    return cudaErrorInvalidValue;
  }
  const int dyn = dv.max_smem - 2048;   // static barriers + staged bias live in the remaining 2 KiB
  p.stages = (dyn - 1024) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) {
    err = "tc_launch: shared memory too small for two pipeline stages";
    return cudaErrorInvalidValue;
  }
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
  // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(tc.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * tc.BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.sbo_bytes = (uint32_t)(8 * tc.KB * 2);
  p.layout_type = tc.KB == 64 ? 2u : tc.KB == 32 ? 4u : 6u;
  p.out_hi = out.hi; p.out_lo = out.lo;
  p.osn = out.sn; p.osh = out.sh; p.osw = out.sw;
  p.bias = tc.bias;
  int cols = 32;
  while (cols < 4 * tc.BN) cols <<= 1;   // two accumulators x [D1 | D2]
  p.tmem_cols = cols;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int grid = total_tiles < dv.num_sms ? total_tiles : dv.num_sms;
  if (tc.KB == 64)
    conv_tc_kernel<64><<<grid, kThreads, dyn, s>>>(it->second, tc.map_b, p);
  else if (tc.KB == 32)
    conv_tc_kernel<32><<<grid, kThreads, dyn, s>>>(it->second, tc.map_b, p);
  else
    conv_tc_kernel<16><<<grid, kThreads, dyn, s>>>(it->second, tc.map_b, p);
  return cudaGetLastError();
}

}  // namespace vr
