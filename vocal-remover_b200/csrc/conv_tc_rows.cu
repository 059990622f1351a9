// Row-streaming tcgen05 convolution: 3x3, stride 1, dilation 1, output width a multiple of 128.
//
// The generic kernel (conv_tc.cu) re-fetches every input pixel from L2 once per tap (9x) and is bound
// by L2->SM bandwidth on the wide, shallow layers (enc1, enc2.conv2, dec2, dec1 of every BaseNet:
// 69 % of the convolution time in profiles/r01_launches_bench30s_v1_direct.csv).  Here one CTA owns a
// block of R=8 output rows x 128 pixels x BN couts with R accumulators resident in TMEM, and streams
// the R+2 input rows it needs through shared memory ONCE per 32-channel chunk:
//   * each input row (130 pixels incl. the +-1 halo, 32 channels, hi and lo plane) is one pair of TMA
//     loads; out-of-image rows / columns are zero-filled by TMA = the conv padding;
//   * a row feeds up to three output rows (kh = 0,1,2) and, for each, the three kw taps are the SAME
//     shared-memory tile read through UMMA descriptors whose start address is shifted by kw pixels
//     (+ kw * 64 B) - no data movement per tap.  Measured on B200: the swizzle is applied to the absolute
//     shared-memory address, so the shifted start needs NO matrix-base-offset (setting the field to
//     (addr>>7)&7 produces garbage);
//   * the weights of the three kh taps are STACKED along the MMA N dimension ([kh=2 | kh=1 | kh=0] x BN couts)
//     and the R accumulators sit in adjacent TMEM columns, so ONE tcgen05.mma of N = 3*BN adds an input
//     row's contribution to output rows r-2, r-1 and r at once;
//   * the 9-tap weight slab of the chunk (3 kw x [3*BN] x 32, hi+lo) is double-buffered in shared memory.
// L2->SM traffic per output pixel drops from 9 to (R+2)/R = 1.25 operand fetches.
//
// MMA issue (round 2): everything the issuer adds to a descriptor inside a row is a compile-time constant (the
// kernel is a template on BN; chunk width, slot and slab strides are constexpr).  With run-time strides ptxas kept
// the descriptor arithmetic in vector registers and moved the operands of every UTCHMMA through R2UR: the row
// kernel issued one MMA per 85-98 cycles whatever its N (profiles/r02_layers_before.tsv), i.e. the issuing thread
// was the limit.  The micro-benchmark profiles/ubench/umma_issue.cu measures the same 18-MMA row with constant
// offsets at 44 cycles per N=48 MMA and 56 per N=96 MMA - the shared-memory operand fetch of the tensor pipe,
// (4096 + 32 N) bytes at 128 B/cycle, which is the next bound (N >= 128 is needed for N/2 cycles).
//
// Fused decoder upsample (optional, Decoder of lib/layers.py:51-64): the leading `up_chunks` channel chunks of the
// input are F.interpolate(x2, bilinear, align_corners=True) of a tensor at half resolution.  Instead of reading a
// materialised up-sampled copy (4x the bytes, and the decoder layers are HBM-bound), nine producer warps
// interpolate each 130-pixel row from TMA-staged half-resolution rows into the swizzled operand slot
// (generic-proxy stores + fence.proxy.async + mbarrier arrive), bit-identical to upsample2x_kernel up to the
// order of the two blends.
#include <stdio.h>

#include "engine.h"
#include "tc_common.cuh"
#include "tc_plan.h"

namespace vr {

static constexpr int kInterpThreads = 288;          // 9 warps: 520 slot items in two passes, 272 source items in one
static constexpr int kRowsThreads = 192 + kInterpThreads;   // TMA, MMA, 4 epilogue warps + the interpolation warps
static constexpr int kMaxR = 8;                    // output rows per CTA tile
static constexpr int kRowPx = 130;                 // 128 + 2 halo pixels
static constexpr int kMaxASlots = 8;
static constexpr int kSrcPx = 68;                  // half-resolution pixels staged per source row (fused upsample)
static constexpr int kStageBytes = 4 * kSrcPx * 64;   // {hi,lo} x {y0,y1} x kSrcPx x 32 channels
static constexpr int kStages = 2;
static constexpr uint32_t kKB = 32;                 // channels per chunk (SWIZZLE_64B rows of 64 bytes)
static constexpr uint32_t kRowB = kKB * 2;          // bytes of one pixel of a chunk
static constexpr uint32_t kAPlane = 9216;           // round_up(kRowPx * kRowB, 1024)
static constexpr uint32_t kASlot = 2 * kAPlane;     // hi plane, lo plane

template <int BN>
struct RowsGeom {
  static constexpr uint32_t kBPlane = 3 * BN * kRowB;   // hi -> lo plane inside one kw slab ([kh=2|kh=1|kh=0] x BN rows)
  static constexpr uint32_t kBKw = 2 * kBPlane;         // one kw slab, both planes
  static constexpr uint32_t kBBuf = 3 * kBKw;           // the three kw slabs of a chunk
};

struct RowsParams {
  int N, H, W, tiles_w, tiles_h, n_tiles, total_tiles;
  int chunks, CinPadR, Cout, act;
  int n_aslots;
  bf16* out_hi;
  bf16* out_lo;
  int64_t osn, osh;
  int osw;
  const float* bias;
  // fused bilinear x2 producer for the first up_chunks chunks (0: everything comes from the TMA map)
  int up_chunks, xH, xW;
  unsigned long long kmask;   // bit g: some weight on input channels [8g, 8g+8) is non-zero (all ones = no skipping)
  int a_c_off;   // channel coordinate of chunk 0 in the TMA map (negative: the map holds only the skip tensor)
  int n_uslots;   // A slots [0, n_uslots) form the ring of the interpolation warps, [n_uslots, n_aslots) the TMA ring:
                  // one producer per ring (two producers sharing one ring can lap each other: the 1-bit phase
                  // parity cannot tell 'two uses behind' from 'up to date')
  float up_sh, up_sw;
};

// 8-channel groups of chunk cc that carry any non-zero weight (4 bits per 32-channel chunk)
__device__ __forceinline__ uint32_t chunk_groups(unsigned long long kmask, int cc) {
  const int sh = cc * 4;
  return sh + 4 <= 64 ? (uint32_t)(kmask >> sh) & 0xFu : 0xFu;
}

// tcgen05.mma with the accumulate flag as a compile-time constant (UPT / !UPT in SASS)
template <int ACC>
__device__ __forceinline__ void umma_c(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo32), "r"(b_lo32), "r"(hi32), "r"(idesc), "n"(ACC)
      : "memory");
}

// The three split-precision products of one 16-channel k-step: hi*hi + lo*hi + hi*lo
template <int BN, int ACC0>
__device__ __forceinline__ void umma_triple(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t dhi,
                                            uint32_t idesc) {
  umma_c<ACC0>(d, a_hi, b_hi, dhi, idesc);
  umma_c<1>(d, a_lo, b_hi, dhi, idesc);
  umma_c<1>(d, a_hi, b_hi + (RowsGeom<BN>::kBPlane >> 4), dhi, idesc);
}

// All MMAs of one input row of one chunk.  KSM: k-steps (16 channels) of the chunk that carry weights (bit 0 / 1).
// a_hi: descriptor low word of the slot's hi plane; b_row: low word of the weight rows of the first accumulator fed.
template <int BN, int KSM>
__device__ __forceinline__ void issue_row(uint32_t d, uint32_t a_hi, uint32_t b_row, uint32_t dhi, uint32_t idesc) {
  const uint32_t a_lo = a_hi + (kAPlane >> 4);
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if (!((KSM >> ks) & 1)) continue;
      const uint32_t ao = (uint32_t)(kw * kRowB + ks * 32) >> 4;
      const uint32_t bo = (uint32_t)(kw * RowsGeom<BN>::kBKw + ks * 32) >> 4;
      umma_triple<BN, 1>(d, a_hi + ao, a_lo + ao, b_row + bo, dhi, idesc);
    }
  }
}

// Same for a row that is the FIRST contribution to its newest accumulator (chunk 0, r < R): k-step 0 of tap kw = 0
// overwrites that accumulator (accumulate = 0) and accumulates into the `cnt - 1` older ones.
template <int BN, int KSM>
__device__ __forceinline__ void issue_row_fresh(uint32_t d, uint32_t a_hi, uint32_t b_row, uint32_t dhi, uint32_t idesc0,
                                                int cnt) {
  const uint32_t a_lo = a_hi + (kAPlane >> 4);
  const uint32_t n_old = (uint32_t)((cnt - 1) * BN);
  const uint32_t idesc_new = idesc0 | ((uint32_t)(BN >> 3) << 17);
  const uint32_t idesc_all = idesc0 | ((uint32_t)((cnt * BN) >> 3) << 17);
  if (cnt > 1) {
    const uint32_t idesc_old = idesc0 | ((n_old >> 3) << 17);
    umma_triple<BN, 1>(d, a_hi, a_lo, b_row, dhi, idesc_old);
  }
  umma_triple<BN, 0>(d + n_old, a_hi, a_lo, b_row + ((n_old * kRowB) >> 4), dhi, idesc_new);
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if ((kw == 0 && ks == 0) || !((KSM >> ks) & 1)) continue;
      const uint32_t ao = (uint32_t)(kw * kRowB + ks * 32) >> 4;
      const uint32_t bo = (uint32_t)(kw * RowsGeom<BN>::kBKw + ks * 32) >> 4;
      umma_triple<BN, 1>(d, a_hi + ao, a_lo + ao, b_row + bo, dhi, idesc_all);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kRowsThreads, 1)
    conv_tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmX, const RowsParams p) {
  typedef RowsGeom<BN> G;
  constexpr int R = kMaxR;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[kMaxASlots];
  __shared__ __align__(8) uint64_t bar_aempty[kMaxASlots];
  __shared__ __align__(8) uint64_t bar_bfull[2];
  __shared__ __align__(8) uint64_t bar_bempty[2];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ __align__(8) uint64_t bar_sfull[kStages];    // half-resolution source rows staged by TMA (fused upsample)
  __shared__ __align__(8) uint64_t bar_sempty[kStages];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[256];   // folded-BN bias of every N tile, staged once (a global load per use stalled the epilogue)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)p.n_aslots * kASlot;
  const uint32_t s_base = b_base + 2 * G::kBBuf;
  const uint32_t v_base = s_base + (uint32_t)(kStages * kStageBytes);   // fp32 vertically blended source row
  constexpr uint32_t kTmemCols = 2 * R * BN;   // 512 (BN=32) or 256 (BN=16): powers of two

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.n_aslots; ++s) {
      mbar_init(smem_u32(&bar_afull[s]), 1);
      mbar_init(smem_u32(&bar_aempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_bfull[s]), 1);
      mbar_init(smem_u32(&bar_bempty[s]), 1);
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 4);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_sfull[s]), 1);
      mbar_init(smem_u32(&bar_sempty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_tiles * BN; i += blockDim.x) bias_s[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged; one elected lane issues) =====================
    int as = p.n_uslots, bs = 0, ss = 0;
    uint32_t aph = 0, bph = 0, sph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int w0 = (mt % p.tiles_w) * 128;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * R;
      const int n = mt / p.tiles_h;
      for (int cc = 0; cc < p.chunks; ++cc) {
        mbar_wait(smem_u32(&bar_bempty[bs]), bph ^ 1u);
        const uint32_t bfull = smem_u32(&bar_bfull[bs]);
        const uint32_t bdst = b_base + (uint32_t)bs * G::kBBuf;
        if (elect_one_sync()) {
          mbar_expect_tx(bfull, G::kBBuf);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            tma_load_3d(bdst + (uint32_t)kw * G::kBKw, &tmB, kw * p.CinPadR + cc * (int)kKB, nt * 3 * BN, 0, bfull);
        }
        __syncwarp();
        if (++bs == 2) {
          bs = 0;
          bph ^= 1u;
        }
        for (int r = 0; r < R + 2; ++r) {
          if (cc < p.up_chunks) {
            // rows of this chunk are produced by the interpolation warps; stage their two half-resolution source
            // rows (hi and lo planes) in shared memory so that each source pixel crosses L2->SM once per row
            mbar_wait(smem_u32(&bar_sempty[ss]), sph ^ 1u);
            const uint32_t sfull = smem_u32(&bar_sfull[ss]);
            const uint32_t sdst = s_base + (uint32_t)(ss * kStageBytes);
            const int h = h0 - 1 + r;
            if (elect_one_sync()) {
              if (h >= 0 && h < p.H) {
                const float fy = p.up_sh * h;
                const int y0 = (int)fy;
                const int y1 = y0 + (y0 < p.xH - 1 ? 1 : 0);
                const int xs = (int)(p.up_sw * (w0 > 0 ? w0 - 1 : 0));
                mbar_expect_tx(sfull, (uint32_t)kStageBytes);
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                  tma_load_5d(sdst + (uint32_t)((pl * 2 + 0) * kSrcPx * 64), &tmX, cc * 32, xs, y0, n, pl, sfull);
                  tma_load_5d(sdst + (uint32_t)((pl * 2 + 1) * kSrcPx * 64), &tmX, cc * 32, xs, y1, n, pl, sfull);
                }
              } else {
                mbar_arrive(sfull);   // halo row outside the image: the interpolation warps write zeros
              }
            }
            __syncwarp();
            if (++ss == kStages) {
              ss = 0;
              sph ^= 1u;
            }
            continue;
          }
          mbar_wait(smem_u32(&bar_aempty[as]), aph ^ 1u);
          const uint32_t afull = smem_u32(&bar_afull[as]);
          const uint32_t adst = a_base + (uint32_t)as * kASlot;
          if (elect_one_sync()) {
            mbar_expect_tx(afull, (uint32_t)(2 * kRowPx * kRowB));
            tma_load_5d(adst, &tmA, cc * (int)kKB + p.a_c_off, w0 - 1, h0 - 1 + r, n, 0, afull);
            tma_load_5d(adst + kAPlane, &tmA, cc * (int)kKB + p.a_c_off, w0 - 1, h0 - 1 + r, n, 1, afull);
          }
          __syncwarp();
          if (++as == p.n_aslots) {
            as = p.n_uslots;
            aph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected lane runs the whole loop nest =====================
    // The tensor pipe queues only a few MMAs, so every cycle the issuing thread spends between the last MMA of a row
    // and the first MMA of the next one is a bubble in the pipe (measured: ~490 cycles of per-row scalar code made a
    // 920-cycle row take 1440).  Hence: the row loop is fully unrolled (accumulator offsets, weight-row offsets and the
    // N field of the instruction descriptor are immediates), per-chunk quantities are hoisted, and the lane election
    // happens once per kernel instead of once per row.
    if (elect_one_sync()) {
      int as_t = p.n_uslots, as_u = 0, bs = 0, acc = 0;
      uint32_t aph_t = 0, aph_u = 0, bph = 0, acc_phase = 0;
      const uint32_t dhi = desc_hi(8 * kRowB, 4u);   // SWIZZLE_64B, 8-row groups of 64-byte rows
      // instruction descriptor without the N field: D=f32, A=B=bf16, K-major, M=128
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
      const uint32_t a_lo0 = desc_lo(a_base);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_set = tmem_base + (uint32_t)(acc * R * BN);
        for (int cc = 0; cc < p.chunks; ++cc) {
          mbar_wait(smem_u32(&bar_bfull[bs]), bph);
          const uint32_t bsrc = desc_lo(b_base + (uint32_t)bs * G::kBBuf);
          const bool up = cc < p.up_chunks;
          // k-steps (16 channels = two groups) whose weights are all zero are not issued: exact, since the products
          // would be 0 (lstm / pad channel groups of the concat layouts); chunk 0 always keeps k-step 0 (accumulator init)
          const uint32_t gm = chunk_groups(p.kmask, cc);
          const uint32_t ksm = ((gm & 0x3u) ? 1u : 0u) | ((gm & 0xCu) ? 2u : 0u);
          // this chunk's ring of A slots: the interpolation ring [0, n_uslots) or the TMA ring [n_uslots, n_aslots)
          int as = up ? as_u : as_t;
          uint32_t aph = up ? aph_u : aph_t;
          const int ring_lo = up ? 0 : p.n_uslots, ring_hi = up ? p.n_uslots : p.n_aslots;
#pragma unroll
          for (int r = 0; r < R + 2; ++r) {
            // input row r feeds output rows o = r-kh; accumulators o_lo..o_hi are adjacent TMEM column blocks; the
            // weight rows are stacked [kh=2 | kh=1 | kh=0], the block of accumulator o_lo is kh = r - o_lo
            constexpr int kR = R;
            const int o_lo = r - 2 < 0 ? 0 : r - 2;
            const int o_hi = r > kR - 1 ? kR - 1 : r;
            const int cnt = o_hi - o_lo + 1;
            const uint32_t d_tmem = d_set + (uint32_t)(o_lo * BN);
            const uint32_t b_row = bsrc + (((uint32_t)((2 - (r - o_lo)) * BN) * kRowB) >> 4);
            const uint32_t idesc_all = idesc0 | ((uint32_t)((cnt * BN) >> 3) << 17);
            mbar_wait(afull0 + (uint32_t)as * 8u, aph);
            const uint32_t a_hi = a_lo0 + (uint32_t)as * (kASlot >> 4);
            if (cc == 0 && r <= kR - 1) {   // accumulator r receives its first product now
              if (ksm == 3u) issue_row_fresh<BN, 3>(d_tmem, a_hi, b_row, dhi, idesc0, cnt);
              else issue_row_fresh<BN, 1>(d_tmem, a_hi, b_row, dhi, idesc0, cnt);
            } else if (ksm == 3u) {
              issue_row<BN, 3>(d_tmem, a_hi, b_row, dhi, idesc_all);
            } else if (ksm == 1u) {
              issue_row<BN, 1>(d_tmem, a_hi, b_row, dhi, idesc_all);
            } else if (ksm == 2u) {
              issue_row<BN, 2>(d_tmem, a_hi, b_row, dhi, idesc_all);
            }
            umma_commit(aempty0 + (uint32_t)as * 8u);
            if (++as == ring_hi) {
              as = ring_lo;
              aph ^= 1u;
            }
          }
          if (up) {
            as_u = as;
            aph_u = aph;
          } else {
            as_t = as;
            aph_t = aph;
          }
          umma_commit(smem_u32(&bar_bempty[bs]));
          if (++bs == 2) {
            bs = 0;
            bph ^= 1u;
          }
        }
        umma_commit(smem_u32(&bar_tfull[acc]));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ===================== bilinear x2 producer (9 warps) =====================
    // Per A-slot row: (A) blend the two staged half-resolution source rows vertically into an fp32 row in shared
    // memory, (B) blend horizontally per output pixel, split to hi/lo and store in the SW64 slot layout.
    // align_corners=True weights as ATen upsample_bilinear2d / upsample2x_kernel (vertical blend first here).
    if (p.up_chunks > 0) {
      const int tid = threadIdx.x - 192;
      int as = 0, ss = 0;
      uint32_t aph = 0, sph = 0;
      float4* v0 = reinterpret_cast<float4*>(smem_raw + (v_base - smem_u32(smem_raw)));   // channels 8j..8j+3
      float4* v1 = v0 + kSrcPx * 4;                                                       // channels 8j+4..8j+7
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * 128;
        mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * R;
        const int xs = (int)(p.up_sw * (w0 > 0 ? w0 - 1 : 0));
        for (int cc = 0; cc < p.up_chunks; ++cc) {
          const uint32_t gmask = chunk_groups(p.kmask, cc);
          for (int r = 0; r < R + 2; ++r) {
            const int h = h0 - 1 + r;
            const bool row_ok = h >= 0 && h < p.H;
            mbar_wait(smem_u32(&bar_sfull[ss]), sph);
            if (row_ok) {
              const uint8_t* stage = smem_raw + (s_base - smem_u32(smem_raw)) + (size_t)ss * kStageBytes;
              const float fy = p.up_sh * h;
              const float ly = fy - (float)(int)fy, hy = 1.f - ly;
              const int rowb = kSrcPx * 64;
              for (int item = tid; item < kSrcPx * 4; item += kInterpThreads) {
                if (!((gmask >> (item & 3)) & 1u)) continue;   // channel group without weights: never read below
                // staged layout: [plane hi,lo][source row y0,y1][kSrcPx pixels from xs][32 channels]
                const int o = item * 16;
                const bf16x8 ah = *reinterpret_cast<const uint4*>(stage + o);
                const bf16x8 ch = *reinterpret_cast<const uint4*>(stage + rowb + o);
                const bf16x8 al = *reinterpret_cast<const uint4*>(stage + 2 * rowb + o);
                const bf16x8 cl = *reinterpret_cast<const uint4*>(stage + 3 * rowb + o);
                float a[8], c[8];
                unpack8(ah, al, a);
                unpack8(ch, cl, c);
                v0[item] = make_float4(hy * a[0] + ly * c[0], hy * a[1] + ly * c[1], hy * a[2] + ly * c[2],
                                       hy * a[3] + ly * c[3]);
                v1[item] = make_float4(hy * a[4] + ly * c[4], hy * a[5] + ly * c[5], hy * a[6] + ly * c[6],
                                       hy * a[7] + ly * c[7]);
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kInterpThreads) : "memory");
            if (tid == 0) mbar_arrive(smem_u32(&bar_sempty[ss]));   // the staged source rows are consumed
            mbar_wait(smem_u32(&bar_aempty[as]), aph ^ 1u);
            uint8_t* slot = smem_raw + (a_base - smem_u32(smem_raw)) + (size_t)as * kASlot;
            for (int item = tid; item < kRowPx * 4; item += kInterpThreads) {
              const int q = item >> 2, j = item & 3;          // pixel of the slot, 8-channel group
              const int w = w0 - 1 + q;
              bf16x8 oh = make_uint4(0, 0, 0, 0), ol = make_uint4(0, 0, 0, 0);
              if (row_ok && w >= 0 && w < p.W && ((gmask >> j) & 1u)) {
                const float fx = p.up_sw * w;
                const int x0 = (int)fx;
                const int x1 = x0 + (x0 < p.xW - 1 ? 1 : 0);
                const float lx = fx - x0, hx = 1.f - lx;
                const int i0 = (x0 - xs) * 4 + j, i1 = (x1 - xs) * 4 + j;
                const float4 pa = v0[i0], pb = v1[i0], qa = v0[i1], qb = v1[i1];
                float y[8];
                y[0] = hx * pa.x + lx * qa.x; y[1] = hx * pa.y + lx * qa.y;
                y[2] = hx * pa.z + lx * qa.z; y[3] = hx * pa.w + lx * qa.w;
                y[4] = hx * pb.x + lx * qb.x; y[5] = hx * pb.y + lx * qb.y;
                y[6] = hx * pb.z + lx * qb.z; y[7] = hx * pb.w + lx * qb.w;
                split8(y, oh, ol);
              }
              // SWIZZLE_64B (same pattern TMA writes and UMMA reads): 16-byte chunk j of 64-byte row q sits at
              // chunk j ^ ((q >> 1) & 3) because the XOR takes address bits [7,9) and the slot is 1 KiB aligned
              const int off = q * 64 + ((j ^ ((q >> 1) & 3)) << 4);
              *reinterpret_cast<uint4*>(slot + off) = oh;
              *reinterpret_cast<uint4*>(slot + kAPlane + off) = ol;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
            asm volatile("bar.sync 1, %0;" ::"n"(kInterpThreads) : "memory");
            if (tid == 0) mbar_arrive(smem_u32(&bar_afull[as]));
            if (++as == p.n_uslots) {
              as = 0;
              aph ^= 1u;
            }
            if (++ss == kStages) {
              ss = 0;
              sph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int px = q * 32 + lane;
    const float slope = p.act == ACT_RELU ? 0.f : p.act == ACT_LEAKY ? 0.01f : 1.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int w0 = (mt % p.tiles_w) * 128;
      mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * R;
      const int n = mt / p.tiles_h;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_set = tmem_base + (uint32_t)(acc * R * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int orow = 0; orow < R; ++orow) {
        const int64_t obase = (int64_t)n * p.osn + (int64_t)(h0 + orow) * p.osh + (int64_t)(w0 + px) * p.osw;
        const bool last_row = orow == R - 1;
        float v[BN];
        if (BN == 32) tmem_ld32(t_set + (uint32_t)(orow * BN), v);
        else tmem_ld16(t_set + (uint32_t)(orow * BN), v);
        if (last_row) {   // all of this warp's TMEM reads are done: hand the accumulator set back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        }
        epilogue_store<BN / 16>(v, bias_s, nt * BN, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
bool tc_rows_prepare(ConvLayer& L, TcConv& tc, std::string& err, std::vector<void*>& allocs) {
  TcRowsPlan& R = tc.rows;
  R.ok = false;
  if (L.k != 3 || L.stride != 1 || L.dil_h != 1 || L.dil_w != 1) return true;
  const int cout16 = round_up(L.Cout, 16);
  R.BN = cout16 == 16 ? 16 : 32;
  R.n_tiles = ceil_div(cout16, R.BN);
  if (R.n_tiles * R.BN > 256) return true;   // bias staging area of the kernel
  R.KB = (int)kKB;
  R.CinPadR = round_up(L.CinPad, R.KB);
  R.chunks = R.CinPadR / R.KB;
  // B[plane][nt*3*BN + (2-kh)*BN + co][kw*CinPadR + ci]: the three kh taps stacked along the MMA N dimension
  const int rows = R.n_tiles * R.BN;
  const int brows = 3 * rows;
  const int Ktot = 3 * R.CinPadR;
  std::vector<uint16_t> planes((size_t)2 * brows * Ktot, 0);
  for (int co = 0; co < L.Cout; ++co) {
    const int nt = co / R.BN, col = co % R.BN;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        for (int ci = 0; ci < L.CinPad; ++ci) {
          const float w = L.w_host[((size_t)(kh * 3 + kw) * L.CinPad + ci) * L.CoutPad + co];
          const uint16_t hi = tc_f2bf(w);
          const uint16_t lo = tc_f2bf(w - tc_bf2f(hi));
          const size_t row = (size_t)nt * 3 * R.BN + (size_t)(2 - kh) * R.BN + col;
          const size_t k = (size_t)kw * R.CinPadR + ci;
          planes[row * Ktot + k] = hi;
          planes[((size_t)brows + row) * Ktot + k] = lo;
        }
  }
  // which 8-channel input groups carry any weight at all (the lstm / pad groups of the concat layouts do not)
  R.kmask = ~0ull;
  if (R.CinPadR / 8 <= 64) {
    R.kmask = 0x3ull;   // k-step 0 of chunk 0 initialises the accumulators: never skipped
    for (int ci = 0; ci < L.CinPad; ++ci) {
      bool any = false;
      for (int t = 0; t < 9 && !any; ++t)
        for (int co = 0; co < L.Cout && !any; ++co) any = L.w_host[((size_t)t * L.CinPad + ci) * L.CoutPad + co] != 0.f;
      if (any) R.kmask |= 1ull << (ci / 8);
    }
  }
  std::vector<float> bias((size_t)rows, 0.f);
  for (int co = 0; co < L.Cout; ++co) bias[(size_t)co] = L.bias_host[(size_t)co];
  void* dw = nullptr;
  void* db = nullptr;
  if (cudaMalloc(&dw, planes.size() * 2) != cudaSuccess || cudaMalloc(&db, bias.size() * 4) != cudaSuccess) {
    err = "cudaMalloc failed while packing row-kernel weights for " + L.name;
    return false;
  }
  allocs.push_back(dw);
  allocs.push_back(db);
  cudaMemcpy(dw, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  R.w_planes = (bf16*)dw;
  R.bias = (float*)db;
  cuuint64_t dims[3] = {(cuuint64_t)Ktot, (cuuint64_t)brows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)Ktot * 2, (cuuint64_t)brows * Ktot * 2};
  cuuint32_t box[3] = {(cuuint32_t)R.KB, (cuuint32_t)(3 * R.BN), 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = tc_encode_fn()(&R.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dw, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    err = "cuTensorMapEncodeTiled(row-kernel weights) failed for " + L.name + " code " + std::to_string((int)r);
    return false;
  }
  R.ok = true;
  return true;
}

bool tc_rows_supported(const ConvLayer& L, const TcConv& tc, const ActView& in, const ActView& out) {
  if (!tc.rows.ok || g_tc_debug[1]) return false;
  if (L.k != 3 || L.stride != 1 || L.dil_h != 1 || L.dil_w != 1) return false;
  if (out.W % 128 || out.H % kMaxR || in.H != out.H || in.W != out.W) return false;
  if (in.sw % 8 || in.sh % 8 || in.sn % 8) return false;
  if ((reinterpret_cast<uintptr_t>(in.hi) | reinterpret_cast<uintptr_t>(in.lo)) & 15) return false;
  return true;
}

// up_src != nullptr: the first up_src->C channels of `in` are NOT read; they are produced inside the kernel as the
// bilinear x2 upsample of *up_src (half resolution).  Needs up_src->C % 32 == 0.
cudaError_t tc_rows_launch(ConvLayer& L, TcConv& tc, const ActView& in, const ActView& out, cudaStream_t s,
                           std::string& err, const ActView* up_src) {
  TcRowsPlan& R = tc.rows;
  if (up_src && (up_src->C % 32 || up_src->H * 2 != in.H || up_src->W * 2 != in.W || up_src->sw % 8 ||
                 up_src->C > R.CinPadR)) {
    err = "tc_rows_launch: fused upsample needs a half-resolution source of 32k channels";
    return cudaErrorInvalidValue;
  }
  ViewKey key = std::make_tuple((const void*)in.hi, (const void*)in.lo, in.N, in.H, in.W, in.C);
  auto it = R.map_a.find(key);
  if (it == R.map_a.end()) {
    CUtensorMap m;
    cuuint64_t dims[5] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N, 2};
    const int64_t plane = (const char*)in.lo - (const char*)in.hi;
    if (plane <= 0 || plane % 16) {
      err = "tc_rows_launch: hi/lo planes must be 16-byte aligned with lo after hi";
      return cudaErrorInvalidValue;
    }
    cuuint64_t strides[4] = {(cuuint64_t)in.sw * 2, (cuuint64_t)in.sh * 2, (cuuint64_t)in.sn * 2, (cuuint64_t)plane};
    cuuint32_t box[5] = {(cuuint32_t)R.KB, (cuuint32_t)kRowPx, 1, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = tc_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)in.hi, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      err = "cuTensorMapEncodeTiled(row-kernel activations) failed for " + L.name + " code " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
    it = R.map_a.emplace(key, m).first;
  }
  RowsParams p;
  p.N = out.N; p.H = out.H; p.W = out.W;
  p.tiles_w = out.W / 128; p.tiles_h = out.H / kMaxR; p.n_tiles = R.n_tiles;
  p.total_tiles = p.tiles_w * p.tiles_h * out.N * R.n_tiles;
  p.chunks = R.chunks; p.CinPadR = R.CinPadR; p.Cout = L.Cout; p.act = L.act;
  p.out_hi = out.hi; p.out_lo = out.lo;
  p.osn = out.sn; p.osh = out.sh; p.osw = out.sw;
  p.bias = R.bias;
  p.up_chunks = 0; p.xH = p.xW = 0;
  p.up_sh = p.up_sw = 0.f;
  p.a_c_off = 0;
  p.kmask = g_tc_debug[6] == 1 ? R.kmask : ~0ull;   // VR_KSKIP=0 issues the all-zero-weight channel groups too
  if (up_src) {
    // `in` is either the whole concat buffer (its first up_src->C channels are then never read) or only the skip
    // tensor, which starts at reduction index up_src->C
    if (in.C + up_src->C <= R.CinPadR) p.a_c_off = -up_src->C;
    p.up_chunks = up_src->C / 32;
    p.xH = up_src->H; p.xW = up_src->W;
    p.up_sh = in.H > 1 ? (float)(up_src->H - 1) / (float)(in.H - 1) : 0.f;   // as launch_upsample2x
    p.up_sw = in.W > 1 ? (float)(up_src->W - 1) / (float)(in.W - 1) : 0.f;
  }
  CUtensorMap map_x;
  if (up_src) {
    // half-resolution source, unswizzled 32-channel x kSrcPx-pixel row boxes (read back by the interpolation warps)
    cuuint64_t dims[5] = {(cuuint64_t)up_src->C, (cuuint64_t)up_src->W, (cuuint64_t)up_src->H, (cuuint64_t)up_src->N, 2};
    const int64_t plane = (const char*)up_src->lo - (const char*)up_src->hi;
    if (plane <= 0 || plane % 16) {
      err = "tc_rows_launch: upsample source hi/lo planes must be 16-byte aligned with lo after hi";
      return cudaErrorInvalidValue;
    }
    cuuint64_t strides[4] = {(cuuint64_t)up_src->sw * 2, (cuuint64_t)up_src->sh * 2, (cuuint64_t)up_src->sn * 2,
                             (cuuint64_t)plane};
    cuuint32_t box[5] = {32, (cuuint32_t)kSrcPx, 1, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = tc_encode_fn()(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)up_src->hi, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      err = "cuTensorMapEncodeTiled(upsample source) failed for " + L.name + " code " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
  }
  const TcDevice& dv = tc_device();
  if (!dv.ok) {
    err = "tc_rows_launch: cannot query the current device";
    return cudaErrorInvalidValue;
  }
  const int b_bytes = R.BN == 16 ? (int)(2 * RowsGeom<16>::kBBuf) : (int)(2 * RowsGeom<32>::kBBuf);
  const int stage_bytes = p.up_chunks > 0 ? kStages * kStageBytes + kSrcPx * 128 : 0;
  p.n_aslots = (dv.max_smem - 3072 - 1024 - b_bytes - stage_bytes) / (int)kASlot;
  if (p.n_aslots > kMaxASlots) p.n_aslots = kMaxASlots;
  if (p.n_aslots < 2) {
    err = "tc_rows_launch: shared memory too small";
    return cudaErrorInvalidValue;
  }
  p.n_uslots = p.up_chunks > 0 ? p.n_aslots / 2 : 0;
  if (p.up_chunks > 0 && g_tc_debug[4] >= 1 && g_tc_debug[4] <= p.n_aslots - 2) p.n_uslots = g_tc_debug[4];
  const int dyn = p.n_aslots * (int)kASlot + b_bytes + stage_bytes + 1024;
  const int grid = p.total_tiles < dv.num_sms ? p.total_tiles : dv.num_sms;   // persistent: one CTA per SM
  if (R.BN == 16)
    conv_tc_rows_kernel<16><<<grid, kRowsThreads, dyn, s>>>(it->second, R.map_b, up_src ? map_x : it->second, p);
  else
    conv_tc_rows_kernel<32><<<grid, kRowsThreads, dyn, s>>>(it->second, R.map_b, up_src ? map_x : it->second, p);
  return cudaGetLastError();
}

// cudaFuncSetAttribute is per device: called by tc_device() the first time a device is used (conv_tc.cu)
void tc_rows_set_attributes(int max_smem) {
  cudaFuncSetAttribute(conv_tc_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
  cudaFuncSetAttribute(conv_tc_rows_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(conv_tc_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
  cudaFuncSetAttribute(conv_tc_rows_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

}  // namespace vr
