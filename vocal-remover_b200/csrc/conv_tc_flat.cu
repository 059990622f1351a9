// Flat-halo tcgen05 convolution: 3x3, stride 1, any dilation, feature maps of width <= 64 stored with zero pad
// pixels after every row (Buffer::Wp, engine.h).
//
// With the horizontal padding materialised as zero pixels, an image is ONE contiguous pixel sequence of pitch P
// and   out_flat[m] = sum_taps in_flat[m + (kh-1)*dh*P + (kw-1)*dw] * W[kh][kw]   holds for every position m
// (positions that fall on pad pixels produce junk that is simply not stored; rows above / below the image are the
// TMA out-of-bounds zero fill).  One CTA therefore owns MT*128 consecutive flat positions (MT accumulators in
// TMEM), loads the pixel segment [m0 - halo, m0 + MT*128 + halo) of a 16-channel chunk ONCE and forms all nine
// taps of all MT accumulators from it with UMMA descriptors whose start address is shifted by whole pixels
// (absolute-address swizzle, see conv_tc_rows.cu).  The 9 per-tap weight tiles stream through a small ring.
// Compared with the generic kernel (one TMA box per tap and per 128-pixel tile) the L2->SM operand traffic per
// MMA drops about 5x, which is what bounded the deep layers (enc3-5.conv2, ASPP dilated, dec3, dec4).
#include <stdio.h>

#include "engine.h"
#include "tc_common.cuh"
#include "tc_plan.h"

namespace vr {

static constexpr int kFlatThreads = 192;
static constexpr int kFlatKB = 16;          // channels per chunk: 32-byte rows, SWIZZLE_32B
static constexpr int kMaxSeg = 2;           // A segment slots
static constexpr int kMaxBSlots = 12;       // weight-tap ring

struct FlatParams {
  int N, H, W, P, tiles_per_img, n_tiles, total_tiles;
  int MT, chunks, CinPadTC, BN, Cout, act;
  int dil_h, dil_w, halo;
  int seg_pieces, seg_plane_bytes, seg_bytes, n_bslots, b_tap_bytes, b_plane_bytes;
  int acc_sets;   // 2: double-buffered accumulators (MT*BN*2 <= 512), 1: single
  uint32_t idesc;
  bf16* out_hi;
  bf16* out_lo;
  int64_t osn, osh;
  int osw;
  const float* bias;
  int tmem_cols;
};

__global__ void __launch_bounds__(kFlatThreads, 1)
    conv_tc_flat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[kMaxSeg];
  __shared__ __align__(8) uint64_t bar_aempty[kMaxSeg];
  __shared__ __align__(8) uint64_t bar_bfull[kMaxBSlots];
  __shared__ __align__(8) uint64_t bar_bempty[kMaxBSlots];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)(kMaxSeg * p.seg_bytes);
  const int M_tile = p.MT * 128;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < kMaxSeg; ++s) {
      mbar_init(smem_u32(&bar_afull[s]), 1);
      mbar_init(smem_u32(&bar_aempty[s]), 1);
    }
    for (int s = 0; s < p.n_bslots; ++s) {
      mbar_init(smem_u32(&bar_bfull[s]), 1);
      mbar_init(smem_u32(&bar_bempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 4);
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
    // ===================== TMA producer =====================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int n = mt / p.tiles_per_img;
      const int m0 = (mt - n * p.tiles_per_img) * M_tile;
      for (int cc = 0; cc < p.chunks; ++cc) {
        mbar_wait(smem_u32(&bar_aempty[as]), aph ^ 1u);
        const uint32_t afull = smem_u32(&bar_afull[as]);
        const uint32_t adst = a_base + (uint32_t)(as * p.seg_bytes);
        if (elect_one_sync()) {
          mbar_expect_tx(afull, (uint32_t)(2 * p.seg_pieces * 128 * kFlatKB * 2));
          for (int pc = 0; pc < p.seg_pieces; ++pc) {
            const int f = m0 - p.halo + pc * 128;
            tma_load_4d(adst + (uint32_t)(pc * 128 * kFlatKB * 2), &tmA, cc * kFlatKB, f, n, 0, afull);
            tma_load_4d(adst + (uint32_t)(p.seg_plane_bytes + pc * 128 * kFlatKB * 2), &tmA, cc * kFlatKB, f, n, 1,
                        afull);
          }
        }
        __syncwarp();
        if (++as == kMaxSeg) {
          as = 0;
          aph ^= 1u;
        }
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(smem_u32(&bar_bempty[bs]), bph ^ 1u);
          const uint32_t bfull = smem_u32(&bar_bfull[bs]);
          if (elect_one_sync()) {
            mbar_expect_tx(bfull, (uint32_t)p.b_tap_bytes);
            tma_load_3d(b_base + (uint32_t)(bs * p.b_tap_bytes), &tmB, tap * p.CinPadTC + cc * kFlatKB, nt * p.BN, 0,
                        bfull);
          }
          __syncwarp();
          if (++bs == p.n_bslots) {
            bs = 0;
            bph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    const uint32_t dhi = desc_hi(8 * kFlatKB * 2, 6);   // SWIZZLE_32B, 8-row groups of 256 B
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_set = tmem_base + (uint32_t)(acc * p.MT * p.BN);
      for (int cc = 0; cc < p.chunks; ++cc) {
        mbar_wait(smem_u32(&bar_afull[as]), aph);
        tc_fence_after();
        const uint32_t seg_hi = a_base + (uint32_t)(as * p.seg_bytes);
        const uint32_t seg_lo = seg_hi + (uint32_t)p.seg_plane_bytes;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(smem_u32(&bar_bfull[bs]), bph);
          tc_fence_after();
          const int kh = tap / 3, kw = tap - kh * 3;
          // segment pixel 0 is flat position m0 - halo; tap (kh,kw) of output m reads m + (kh-1)*dh*P + (kw-1)*dw
          const uint32_t shift = (uint32_t)((kh * p.dil_h * p.P + kw * p.dil_w) * kFlatKB * 2);
          const uint32_t b_hi = desc_lo(b_base + (uint32_t)(bs * p.b_tap_bytes));
          const uint32_t b_lo = desc_lo(b_base + (uint32_t)(bs * p.b_tap_bytes + p.b_plane_bytes));
          const uint32_t accumulate = (cc | tap) != 0 ? 1u : 0u;
          for (int j = 0; j < p.MT; ++j) {
            const uint32_t a_hi = desc_lo(seg_hi + shift + (uint32_t)(j * 128 * kFlatKB * 2));
            const uint32_t a_lo = desc_lo(seg_lo + shift + (uint32_t)(j * 128 * kFlatKB * 2));
            const uint32_t d_tmem = d_set + (uint32_t)(j * p.BN);
            if (elect_one_sync()) {
              umma_bf16_w(d_tmem, a_hi, b_hi, dhi, p.idesc, accumulate);
              umma_bf16_w(d_tmem, a_lo, b_hi, dhi, p.idesc, 1u);
              umma_bf16_w(d_tmem, a_hi, b_lo, dhi, p.idesc, 1u);
            }
          }
          __syncwarp();
          if (elect_one_sync()) umma_commit(smem_u32(&bar_bempty[bs]));
          if (++bs == p.n_bslots) {
            bs = 0;
            bph ^= 1u;
          }
        }
        if (elect_one_sync()) umma_commit(smem_u32(&bar_aempty[as]));
        if (++as == kMaxSeg) {
          as = 0;
          aph ^= 1u;
        }
      }
      if (elect_one_sync()) umma_commit(smem_u32(&bar_tfull[acc]));
      if (++acc == p.acc_sets) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float slope = p.act == ACT_RELU ? 0.f : p.act == ACT_LEAKY ? 0.01f : 1.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int n = mt / p.tiles_per_img;
      const int m0 = (mt - n * p.tiles_per_img) * M_tile;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_set = tmem_base + (uint32_t)(acc * p.MT * p.BN) + ((uint32_t)(q * 32) << 16);
      for (int j = 0; j < p.MT; ++j) {
        const int m = m0 + j * 128 + row;
        const int h = m / p.P;
        const int w = m - h * p.P;
        const bool valid = h < p.H && w < p.W;
        const int64_t obase = (int64_t)n * p.osn + (int64_t)h * p.osh + (int64_t)w * p.osw;
        int c0 = 0;
        for (; c0 + 32 <= p.BN; c0 += 32) {
          float v[32];
          tmem_ld32(t_set + (uint32_t)(j * p.BN + c0), v);
          if (j == p.MT - 1 && c0 + 32 >= p.BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
          }
          if (valid) epilogue_store<2>(v, bias_s, nt * p.BN + c0, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
        }
        if (c0 < p.BN) {
          float v[16];
          tmem_ld16(t_set + (uint32_t)(j * p.BN + c0), v);
          if (j == p.MT - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
          }
          if (valid) epilogue_store<1>(v, bias_s, nt * p.BN + c0, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
        }
      }
      if (++acc == p.acc_sets) {
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
// The weight matrix is the generic kernel's ([2][CoutPadN][taps*CinPadTC], K = tap*CinPadTC + ci); only the TMA
// box differs (16 channels, SWIZZLE_32B).
bool tc_flat_prepare(ConvLayer& L, TcConv& tc, std::string& err) {
  TcFlatPlan& F = tc.flat;
  F.ok = false;
  if (L.k != 3 || L.stride != 1) return true;
  cuuint64_t dims[3] = {(cuuint64_t)tc.Ktot, (cuuint64_t)(tc.n_tiles * tc.BN), 2};
  cuuint64_t strides[2] = {(cuuint64_t)tc.Ktot * 2, (cuuint64_t)(tc.n_tiles * tc.BN) * tc.Ktot * 2};
  cuuint32_t box[3] = {(cuuint32_t)kFlatKB, (cuuint32_t)tc.BN, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = tc_encode_fn()(&F.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)tc.w_planes, dims, strides, box,
                              es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    err = "cuTensorMapEncodeTiled(flat-kernel weights) failed for " + L.name + " code " + std::to_string((int)r);
    return false;
  }
  F.ok = true;
  return true;
}

bool tc_flat_supported(const ConvLayer& L, const TcConv& tc, const ActView& in, const ActView& out) {
  if (!tc.flat.ok || g_tc_debug[3] != 1) return false;   // experimental: opt-in (VR_FLAT=1), see DESIGN.md
  if (L.k != 3 || L.stride != 1) return false;
  if (out.W > 64 || in.W != out.W || in.H != out.H) return false;
  if (in.sh % in.sw) return false;
  const int P = (int)(in.sh / in.sw);
  if (P - in.W < L.dil_w) return false;            // needs >= dil_w zero pad pixels after every row
  if (in.sn != (int64_t)in.H * in.sh) return false;  // images must be whole [H][P] blocks (no band views)
  if (in.sw % 8) return false;
  if ((reinterpret_cast<uintptr_t>(in.hi) | reinterpret_cast<uintptr_t>(in.lo)) & 15) return false;
  return true;
}

cudaError_t tc_flat_launch(ConvLayer& L, TcConv& tc, const ActView& in, const ActView& out, cudaStream_t s,
                           std::string& err) {
  TcFlatPlan& F = tc.flat;
  const int P = (int)(in.sh / in.sw);
  ViewKey key = std::make_tuple((const void*)in.hi, (const void*)in.lo, in.N, in.H, in.W, in.C);
  auto it = F.map_a.find(key);
  if (it == F.map_a.end()) {
    CUtensorMap m;
    // flat view of one image: [H*P pixels][C], images and planes as outer dimensions
    cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)in.H * P, (cuuint64_t)in.N, 2};
    const int64_t plane = (const char*)in.lo - (const char*)in.hi;
    if (plane <= 0 || plane % 16) {
      err = "tc_flat_launch: hi/lo planes must be 16-byte aligned with lo after hi";
      return cudaErrorInvalidValue;
    }
    cuuint64_t strides[3] = {(cuuint64_t)in.sw * 2, (cuuint64_t)in.sn * 2, (cuuint64_t)plane};
    cuuint32_t box[4] = {(cuuint32_t)kFlatKB, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = tc_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in.hi, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      err = "cuTensorMapEncodeTiled(flat-kernel activations) failed for " + L.name + " code " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
    it = F.map_a.emplace(key, m).first;
  }
  static bool attr_set = false;
  static int num_sms = 0, max_smem = 0;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(conv_tc_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
    attr_set = true;
  }
  FlatParams p;
  p.N = out.N; p.H = out.H; p.W = out.W; p.P = P;
  p.dil_h = L.dil_h; p.dil_w = L.dil_w;
  p.halo = L.dil_h * P + L.dil_w;
  p.BN = tc.BN; p.n_tiles = tc.n_tiles; p.Cout = L.Cout; p.act = L.act;
  p.chunks = tc.CinPadTC / kFlatKB; p.CinPadTC = tc.CinPadTC;
  p.b_plane_bytes = tc.BN * kFlatKB * 2;
  p.b_tap_bytes = 2 * p.b_plane_bytes;
  const int budget = max_smem - 3072 - 1024;
  // accumulators: as many 128-position groups as fit TMEM and shared memory
  int MT = 512 / tc.BN;
  if (MT > 4) MT = 4;
  for (;; --MT) {
    const int pieces = ceil_div(MT * 128 + 2 * p.halo, 128);
    const int plane = round_up(pieces * 128 * kFlatKB * 2, 1024);
    const int left = budget - kMaxSeg * 2 * plane;
    if (left >= 4 * p.b_tap_bytes || MT == 1) {
      p.MT = MT; p.seg_pieces = pieces; p.seg_plane_bytes = plane; p.seg_bytes = 2 * plane;
      p.n_bslots = left / p.b_tap_bytes;
      break;
    }
  }
  if (p.n_bslots > kMaxBSlots) p.n_bslots = kMaxBSlots;
  if (p.n_bslots < 2) {
    err = "tc_flat_launch: shared memory too small for " + L.name;
    return cudaErrorInvalidValue;
  }
  p.acc_sets = (2 * p.MT * tc.BN <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_sets * p.MT * tc.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.tiles_per_img = ceil_div(out.H * P, p.MT * 128);
  p.total_tiles = p.tiles_per_img * out.N * tc.n_tiles;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(tc.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.out_hi = out.hi; p.out_lo = out.lo;
  p.osn = out.sn; p.osh = out.sh; p.osw = out.sw;
  p.bias = tc.bias;
  const int dyn = kMaxSeg * p.seg_bytes + p.n_bslots * p.b_tap_bytes + 1024;
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  conv_tc_flat_kernel<<<grid, kFlatThreads, dyn, s>>>(it->second, F.map_b, p);
  return cudaGetLastError();
}

}  // namespace vr
