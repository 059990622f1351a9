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

static constexpr int kInterpWarps = 10;             // each covers 7 staged source pixels (+1 neighbour): 70 >= kSrcPx
static constexpr int kInterpThreads = 32 * kInterpWarps;
static constexpr int kRowsThreads = 192 + kInterpThreads;   // TMA, MMA, 4 epilogue warps + the interpolation warps
static constexpr int kMaxR = 8;                    // output rows per CTA tile
static constexpr int kRowPx = 130;                 // 128 + 2 halo pixels
static constexpr int kBoxPx = 136;                 // pixels per TMA row box: makes one plane 17 x 512 B, so that the lo plane
                                                   // of the two-plane box starts on the SWIZZLE_64B repeat (8 rows x 64 B)
static constexpr int kMaxASlots = 8;
static constexpr int kSrcPx = 68;                  // half-resolution pixels a 130-pixel row interpolates from (fused upsample)
static constexpr uint32_t kKB = 32;                 // channels per chunk (SWIZZLE_64B rows of 64 bytes)
static constexpr uint32_t kRowB = kKB * 2;          // bytes of one pixel of a chunk
static constexpr uint32_t kAPlane = kBoxPx * kRowB; // 8704: hi plane, then lo plane (one two-plane TMA box per row)
static constexpr uint32_t kASlot = 2 * kAPlane;     // 17408 = 17 KiB

// Optional timeline of CTA 0 (builds with -DVR_TRACE only: vr_debug_set(0, 1), read back with vr_debug_trace): clock64 stamps of the three producer /
// consumer loops, to see which of them the others wait for.  [role][event index][3] : role 0 = MMA issuer (before the
// operand wait, after it, after the row's last MMA was issued), role 1 = TMA producer (before the slot wait, after the
// load was issued, 0), role 2 = interpolation warp 0 (row start, after the slot wait, after the arrive).
static constexpr int kTraceEvents = 2048;
__device__ unsigned long long g_rows_trace[3 * kTraceEvents * 3];

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
  const bf16* x_hi;   // half-resolution source (NHWC split-bf16), read with plain 16-byte loads by the producer warps
  const bf16* x_lo;
  int64_t xsn, xsh;
  int xsw;
  int trace;   // 1: CTA 0 records its timeline in g_rows_trace
  unsigned long long kmask;   // bit g: some weight on input channels [8g, 8g+8) is non-zero (all ones = no skipping)
  int a_c_off;   // channel coordinate of chunk 0 in the TMA map (negative: the map holds only the skip tensor)
  int l_chunk;   // >= 0: this chunk is read through the second activation map (tmL) from channel 0 (dec1's up-sampled
                 // LSTM channel group, kept in a buffer of its own so that the skip tensor stays dense)
  int n_uslots;   // A slots [0, n_uslots) form the ring of the interpolation warps, [n_uslots, n_aslots) the TMA ring:
                  // one producer per ring (two producers sharing one ring can lap each other: the 1-bit phase
                  // parity cannot tell 'two uses behind' from 'up to date')
  float up_sh, up_sw;
  // fused 1x1 convolution to ONE channel (the LSTM branch's input convolution, lib/layers.py:112,126): every output pixel
  // adds sum_c dot_w[c] * y[c] over this tile's output channels to dot_out[(n * H + h) * W + w] (pre-zeroed fp32 plane)
  const float* dot_w;
  float* dot_out;
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

// The MMAs of one input row of one chunk, in two parts so that the issuer can wait for the NEXT row's operands while the
// tensor pipe still has this row's last products queued: PART 0 = everything but the last k-step triple, PART 1 = that
// triple.  KSM: k-steps (16 channels) of the chunk that carry weights (bit 0 / 1).  a_hi: descriptor low word of the
// slot's hi plane; b_row: low word of the weight rows of the first accumulator fed.
template <int BN, int KSM, int PART>
__device__ __forceinline__ void issue_row(uint32_t d, uint32_t a_hi, uint32_t b_row, uint32_t dhi, uint32_t idesc) {
  const uint32_t a_lo = a_hi + (kAPlane >> 4);
  constexpr int kLastKs = (KSM & 2) ? 1 : 0;
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if (!((KSM >> ks) & 1)) continue;
      const bool last = kw == 2 && ks == kLastKs;
      if ((PART == 0) == last) continue;
      const uint32_t ao = (uint32_t)(kw * kRowB + ks * 32) >> 4;
      const uint32_t bo = (uint32_t)(kw * RowsGeom<BN>::kBKw + ks * 32) >> 4;
      umma_triple<BN, 1>(d, a_hi + ao, a_lo + ao, b_row + bo, dhi, idesc);
    }
  }
}

// Same for a row that is the FIRST contribution to its newest accumulator (chunk 0, r < R): k-step 0 of tap kw = 0
// overwrites that accumulator (accumulate = 0) and accumulates into the `cnt - 1` older ones; the rest is issue_row
// minus that k-step.
template <int BN, int KSM>
__device__ __forceinline__ void issue_row_fresh_head(uint32_t d, uint32_t a_hi, uint32_t b_row, uint32_t dhi, uint32_t idesc0,
                                                     int cnt) {
  const uint32_t a_lo = a_hi + (kAPlane >> 4);
  const uint32_t n_old = (uint32_t)((cnt - 1) * BN);
  const uint32_t idesc_new = idesc0 | ((uint32_t)(BN >> 3) << 17);
  const uint32_t idesc_all = idesc0 | ((uint32_t)((cnt * BN) >> 3) << 17);
  constexpr int kLastKs = (KSM & 2) ? 1 : 0;
  if (cnt > 1) {
    const uint32_t idesc_old = idesc0 | ((n_old >> 3) << 17);
    umma_triple<BN, 1>(d, a_hi, a_lo, b_row, dhi, idesc_old);
  }
  umma_triple<BN, 0>(d + n_old, a_hi, a_lo, b_row + ((n_old * kRowB) >> 4), dhi, idesc_new);
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if ((kw == 0 && ks == 0) || !((KSM >> ks) & 1) || (kw == 2 && ks == kLastKs)) continue;
      const uint32_t ao = (uint32_t)(kw * kRowB + ks * 32) >> 4;
      const uint32_t bo = (uint32_t)(kw * RowsGeom<BN>::kBKw + ks * 32) >> 4;
      umma_triple<BN, 1>(d, a_hi + ao, a_lo + ao, b_row + bo, dhi, idesc_all);
    }
  }
}

// sum_i w[co + i] * act(v[i] + bias[co + i]) over CNT accumulator columns: the fused single-channel 1x1 convolution on
// the fp32 activations of this tile (weights past Cout are zero)
template <int CNT>
__device__ __forceinline__ float dot_activated(const float* v, const float* bias_s, const float* dot_s, int co, float slope) {
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < CNT; i += 2) {
    const float t0 = v[i] + bias_s[co + i], t1 = v[i + 1] + bias_s[co + i + 1];
    d0 = fmaf(fmaxf(t0, 0.f) + slope * fminf(t0, 0.f), dot_s[co + i], d0);
    d1 = fmaf(fmaxf(t1, 0.f) + slope * fminf(t1, 0.f), dot_s[co + i + 1], d1);
  }
  return d0 + d1;
}

template <int BN>
__global__ void __launch_bounds__(kRowsThreads, 1)
    conv_tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmL, const RowsParams p) {
  typedef RowsGeom<BN> G;
  constexpr int R = kMaxR;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[kMaxASlots];
  __shared__ __align__(8) uint64_t bar_aempty[kMaxASlots];
  __shared__ __align__(8) uint64_t bar_bfull[2];
  __shared__ __align__(8) uint64_t bar_bempty[2];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[256];   // folded-BN bias of every N tile, staged once (a global load per use stalled the epilogue)
  __shared__ float dot_s[256];    // weights of the fused single-channel 1x1 convolution (zeros past Cout)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)p.n_aslots * kASlot;
  // accumulators: R rows x BN columns per set; two sets (the epilogue of tile i overlaps the MMAs of tile i+1) when
  // they fit the 512 TMEM columns, one set for BN = 64
  constexpr int kAccSets = 2 * R * BN <= 512 ? 2 : 1;
  constexpr uint32_t kTmemCols = kAccSets * R * BN;   // 512 (BN=64, 32) or 256 (BN=16): powers of two

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.n_aslots; ++s) {
      // slots of the interpolation ring are filled by kInterpWarps producers (one arrive each), the others by one TMA box
      mbar_init(smem_u32(&bar_afull[s]), s < p.n_uslots ? (uint32_t)kInterpWarps : 1u);
      mbar_init(smem_u32(&bar_aempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_bfull[s]), 1);
      mbar_init(smem_u32(&bar_bempty[s]), 1);
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_tiles * BN; i += blockDim.x) {
    bias_s[i] = __ldg(p.bias + i);
    dot_s[i] = p.dot_out && i < p.Cout ? __ldg(p.dot_w + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: one elected lane runs the whole loop nest =====================
    if (elect_one_sync()) {
      int as = p.n_uslots, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
#ifdef VR_TRACE
      const bool tr = p.trace && blockIdx.x == 0;
#else
      constexpr bool tr = false;
#endif
      int tn = 0;
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
          mbar_expect_tx(bfull, G::kBBuf);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            tma_load_3d(bdst + (uint32_t)kw * G::kBKw, &tmB, kw * p.CinPadR + cc * (int)kKB, nt * 3 * BN, 0, bfull);
          if (++bs == 2) {
            bs = 0;
            bph ^= 1u;
          }
          if (cc < p.up_chunks) continue;   // rows of this chunk are produced by the interpolation warps
          const bool from_l = cc == p.l_chunk;
          const int c0 = from_l ? 0 : cc * (int)kKB + p.a_c_off;
          for (int r = 0; r < R + 2; ++r) {
            const unsigned long long t0 = tr ? clock64() : 0ull;
            mbar_wait(aempty0 + (uint32_t)as * 8u, aph ^ 1u);
            const uint32_t afull = afull0 + (uint32_t)as * 8u;
            mbar_expect_tx(afull, kASlot);
            tma_load_5d(a_base + (uint32_t)as * kASlot, from_l ? &tmL : &tmA, c0, w0 - 1, h0 - 1 + r, n, 0, afull);
            if (tr && tn < kTraceEvents) {
              g_rows_trace[(1 * kTraceEvents + tn) * 3 + 0] = t0;
              g_rows_trace[(1 * kTraceEvents + tn) * 3 + 1] = clock64();
              g_rows_trace[(1 * kTraceEvents + tn) * 3 + 2] = 0ull;
              ++tn;
            }
            if (++as == p.n_aslots) {
              as = p.n_uslots;
              aph ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected lane runs the whole loop nest =====================
    // The tensor pipe queues only a few MMAs, so every cycle the issuing thread spends between the last MMA of a row
    // and the first MMA of the next one is a bubble in the pipe (measured: ~490 cycles of per-row scalar code made a
    // 920-cycle row take 1440).  Hence: the row loop is fully unrolled (accumulator offsets, weight-row offsets and the
    // N field of the instruction descriptor are immediates), per-chunk quantities are hoisted, the lane election
    // happens once per kernel, and the wait for the next row's operands is issued BEFORE the last k-step of the
    // current row so that it overlaps the products still queued.
    if (elect_one_sync()) {
      int as_t = p.n_uslots, as_u = 0, bs = 0, acc = 0;
      uint32_t aph_t = 0, aph_u = 0, bph = 0, acc_phase = 0;
      const uint32_t dhi = desc_hi(8 * kRowB, 4u);   // SWIZZLE_64B, 8-row groups of 64-byte rows
      // instruction descriptor without the N field: D=f32, A=B=bf16, K-major, M=128
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
      const uint32_t a_lo0 = desc_lo(a_base);
#ifdef VR_TRACE
      const bool tr = p.trace && blockIdx.x == 0;
#else
      constexpr bool tr = false;
#endif
      int tn = 0;
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
          unsigned long long t_w0 = tr ? clock64() : 0ull;
          mbar_wait(afull0 + (uint32_t)as * 8u, aph);
          unsigned long long t_w1 = tr ? clock64() : 0ull;
          // fully unrolled: measured 1080 cycles per steady N=96 row against 1235 with a rolled loop (timeline of CTA 0,
          // profiles/tools/trace_rows.py); the rows of chunk 0 cost ~1700 either way - they overlap the previous tile's
          // epilogue, whose tcgen05.ld traffic competes with the accumulator read-modify-write of the MMAs
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
            const uint32_t a_hi = a_lo0 + (uint32_t)as * (kASlot >> 4);
            const uint32_t aempty = aempty0 + (uint32_t)as * 8u;
            if (cc == 0 && r <= kR - 1) {   // accumulator r receives its first product now
              if (ksm == 3u) issue_row_fresh_head<BN, 3>(d_tmem, a_hi, b_row, dhi, idesc0, cnt);
              else issue_row_fresh_head<BN, 1>(d_tmem, a_hi, b_row, dhi, idesc0, cnt);
            } else if (ksm == 3u) {
              issue_row<BN, 3, 0>(d_tmem, a_hi, b_row, dhi, idesc_all);
            } else if (ksm == 1u) {
              issue_row<BN, 1, 0>(d_tmem, a_hi, b_row, dhi, idesc_all);
            } else if (ksm == 2u) {
              issue_row<BN, 2, 0>(d_tmem, a_hi, b_row, dhi, idesc_all);
            }
            if (++as == ring_hi) {
              as = ring_lo;
              aph ^= 1u;
            }
            unsigned long long t_n0 = 0ull, t_n1 = 0ull;
            if (r < R + 1) {   // next row of this chunk, while MMAs are queued
              if (tr) t_n0 = clock64();
              mbar_wait(afull0 + (uint32_t)as * 8u, aph);
              if (tr) t_n1 = clock64();
            }
            if (ksm == 3u) issue_row<BN, 3, 1>(d_tmem, a_hi, b_row, dhi, idesc_all);
            else if (ksm == 1u) issue_row<BN, 1, 1>(d_tmem, a_hi, b_row, dhi, idesc_all);
            else if (ksm == 2u) issue_row<BN, 2, 1>(d_tmem, a_hi, b_row, dhi, idesc_all);
            umma_commit(aempty);
            if (tr && tn < kTraceEvents) {
              g_rows_trace[(0 * kTraceEvents + tn) * 3 + 0] = t_w0;
              g_rows_trace[(0 * kTraceEvents + tn) * 3 + 1] = t_w1;
              g_rows_trace[(0 * kTraceEvents + tn) * 3 + 2] = clock64();
              ++tn;
            }
            t_w0 = t_n0;
            t_w1 = t_n1;
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
        if (++acc == kAccSets) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ===================== bilinear x2 producer (kInterpWarps autonomous warps) =====================
    // align_corners=True bilinear x2 (ATen upsample_bilinear2d / upsample2x_kernel weights; the vertical blend is done
    // first here).  Warp k owns the source pixels xs + [7k, 7k+7) of the row (+ pixel 7k+7 as right neighbour): lane =
    // (source pixel, 8-channel group) reads its two source rows (hi and lo plane: four 16-byte global loads, issued one
    // row AHEAD so that their latency overlaps the previous row's arithmetic), blends them vertically in registers,
    // fetches the right neighbour's blend by shuffle and emits the 2-3 output pixels whose left source pixel it is,
    // split to hi/lo, straight into the SWIZZLE_64B slot.  Which output pixels those are (and their horizontal weights
    // and slot offsets) depends only on the tile: computed once per tile.  No block-wide barrier and no shared-memory
    // staging (the tensor pipe already uses the full shared-memory bandwidth for its operands): every warp waits for
    // the slot (MMA commit) itself and arrives on the slot's mbarrier (count = kInterpWarps).
    if (p.up_chunks > 0) {
      const int wk = warp - 6;
      const int xi = lane >> 2, j = lane & 3;
      const int sx = 7 * wk + xi;              // source pixel of this lane, relative to xs
      const bool emit = xi < 7 && sx < kSrcPx; // xi == 7 only provides the neighbour of xi == 6
      int as = 0;
      uint32_t aph = 0;
      const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
      const float inv_sw = p.up_sw > 0.f ? 1.f / p.up_sw : 0.f;
#ifdef VR_TRACE
      const bool tr = p.trace && blockIdx.x == 0 && wk == 0;
#else
      constexpr bool tr = false;
#endif
      int tn = 0;
      const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int fills = my_tiles * p.up_chunks * (R + 2);
      // iteration state of the NEXT row to fetch (one ahead of the row being written); the tile decomposition
      // (divisions) is redone only when the fetch moves on to another tile
      int f_tile = blockIdx.x, f_cc = 0, f_r = 0, f_h0 = 0, f_X = 0;
      int64_t f_base = 0;
      bool f_px = false;
      auto fetch_tile = [&]() {
        int mt = f_tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * 128;
        mt /= p.tiles_w;
        f_h0 = (mt % p.tiles_h) * R - 1;
        f_X = (int)(p.up_sw * (w0 > 0 ? w0 - 1 : 0)) + sx;
        f_px = sx < kSrcPx && f_X < p.xW;
        f_base = (int64_t)(mt / p.tiles_h) * p.xsn + (int64_t)f_X * p.xsw + j * 8;
      };
      fetch_tile();
      // Two rows are in flight: the loads of row k+2 are issued right after the proxy fence of row k (fence.proxy.async
      // compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and the MEMBAR waits for every global load still in flight), so that
      // they have the whole of row k+1 to land before the next fence.  q* = row k+1, n* = row k+2.
      bf16x8 qah, qch, qal, qcl, nah, nch, nal, ncl;   // rows y0 / y1 of the hi plane, rows y0 / y1 of the lo plane
      float q_ly = 0.f, n_ly = 0.f;
      bool q_ok = false, n_ok = false;
      qah = qch = qal = qcl = nah = nch = nal = ncl = make_uint4(0, 0, 0, 0);
      auto fetch = [&]() {
        qah = nah; qch = nch; qal = nal; qcl = ncl;
        q_ly = n_ly;
        q_ok = n_ok;
        const int h = f_h0 + f_r;
        const bool grp = (chunk_groups(p.kmask, f_cc) >> j) & 1u;   // channel group without weights: zeros are written
        n_ok = f_tile < p.total_tiles && h >= 0 && h < p.H && grp && f_px;
        if (n_ok) {
          const float fy = p.up_sh * h;
          const int y0 = (int)fy;
          n_ly = fy - (float)y0;
          const int64_t o0 = f_base + (int64_t)y0 * p.xsh + f_cc * 32;
          const int64_t o1 = o0 + (y0 < p.xH - 1 ? p.xsh : 0);
          nah = ld128(p.x_hi + o0);
          nch = ld128(p.x_hi + o1);
          nal = ld128(p.x_lo + o0);
          ncl = ld128(p.x_lo + o1);
        }
        if (++f_r == R + 2) {
          f_r = 0;
          if (++f_cc == p.up_chunks) {
            f_cc = 0;
            f_tile += gridDim.x;
            if (f_tile < p.total_tiles) fetch_tile();
          }
        }
      };
      if (fills > 0) {
        fetch();   // row 0
        fetch();   // row 1 (row 0 moves to q*)
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt = tile / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * 128;
        const int X = (int)(p.up_sw * (w0 > 0 ? w0 - 1 : 0)) + sx;   // absolute source pixel
        // the output pixels w with (int)(up_sw * w) == X lie in [wc - 1, wc + 3] and there are at most 3 of them
        int e_off[3];
        float e_lx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) e_off[k] = -1;
        if (emit) {
          const int wc = (int)((float)X * inv_sw);
          int cnt = 0;
#pragma unroll
          for (int dw = -1; dw <= 3; ++dw) {
            const int w = wc + dw;
            const float fx = p.up_sw * (float)w;
            if (w < 0 || w >= p.W || (int)fx != X || w < w0 - 1 || w > w0 + 128) continue;
            const int q = w - (w0 - 1);
            // SWIZZLE_64B (same pattern TMA writes and UMMA reads): 16-byte chunk j of 64-byte row q sits at chunk
            // j ^ ((q >> 1) & 3) because the XOR takes address bits [7,9) and the planes are 512-byte aligned
            const int off = q * 64 + ((j ^ ((q >> 1) & 3)) << 4);
            const float lx = fx - (float)X;
            if (cnt == 0) { e_off[0] = off; e_lx[0] = lx; }
            else if (cnt == 1) { e_off[1] = off; e_lx[1] = lx; }
            else if (cnt == 2) { e_off[2] = off; e_lx[2] = lx; }
            ++cnt;
          }
        }
        // conv padding columns of the row: slot pixel 0 at the left image border, slot pixel 129 at the right one
        int z_off = -1;
        if (wk == 0 && lane < 8) {
          const int q = lane < 4 ? 0 : kRowPx - 1;
          const int w = w0 - 1 + q;
          if (w < 0 || w >= p.W) z_off = q * 64 + ((j ^ ((q >> 1) & 3)) << 4);
        }
        const bool last_px = X >= p.xW - 1;   // x1 = x0 on the last source pixel (upsample2x_kernel)
        for (int cc = 0; cc < p.up_chunks; ++cc) {
          for (int r = 0; r < R + 2; ++r) {
            const unsigned long long t0 = tr ? clock64() : 0ull;
            float v[8];
            if (q_ok) {
              float a[8], c[8];
              unpack8(qah, qal, a);
              unpack8(qch, qcl, c);
              const float ly = q_ly, hy = 1.f - q_ly;
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = hy * a[i] + ly * c[i];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            float nv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float t = __shfl_down_sync(0xffffffffu, v[i], 4);
              nv[i] = last_px ? v[i] : t;
            }
            const unsigned long long t1 = tr ? clock64() : 0ull;
            mbar_wait(aempty0 + (uint32_t)as * 8u, aph ^ 1u);
            const unsigned long long t2 = tr ? clock64() : 0ull;
            uint8_t* slot = smem_raw + (a_base - smem_u32(smem_raw)) + (size_t)as * kASlot;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (e_off[k] < 0) continue;
              const float lx = e_lx[k], hx = 1.f - lx;
              float y[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = hx * v[i] + lx * nv[i];
              bf16x8 oh, ol;
              split8(y, oh, ol);
              *reinterpret_cast<uint4*>(slot + e_off[k]) = oh;
              *reinterpret_cast<uint4*>(slot + kAPlane + e_off[k]) = ol;
            }
            if (z_off >= 0) {
              *reinterpret_cast<uint4*>(slot + z_off) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(slot + kAPlane + z_off) = make_uint4(0, 0, 0, 0);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(afull0 + (uint32_t)as * 8u);
            fetch();   // row k+2 (after the fence, see above); past the last row it only shifts the pipeline
            if (tr && lane == 0 && tn < kTraceEvents) {
              g_rows_trace[(2 * kTraceEvents + tn) * 3 + 0] = t0;
              g_rows_trace[(2 * kTraceEvents + tn) * 3 + 1] = t2 - t1;   // cycles spent waiting for the slot
              g_rows_trace[(2 * kTraceEvents + tn) * 3 + 2] = clock64();
            }
            ++tn;
            if (++as == p.n_uslots) {
              as = 0;
              aph ^= 1u;
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
        if (BN == 64) {
          float v[32], v2[32];
          tmem_ld32(t_set + (uint32_t)(orow * BN), v);
          tmem_ld32(t_set + (uint32_t)(orow * BN + 32), v2);
          if (last_row) {   // all of this warp's TMEM reads are done: hand the accumulator set back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
          }
          if (p.dot_out) {
            const float d = dot_activated<32>(v, bias_s, dot_s, nt * BN, slope) +
                            dot_activated<32>(v2, bias_s, dot_s, nt * BN + 32, slope);
            atomicAdd(p.dot_out + ((int64_t)n * p.H + (h0 + orow)) * p.W + (w0 + px), d);
          }
          epilogue_store<2>(v, bias_s, nt * BN, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
          epilogue_store<2>(v2, bias_s, nt * BN + 32, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
        } else {
          float v[BN];
          if (BN == 32) tmem_ld32(t_set + (uint32_t)(orow * BN), v);
          else tmem_ld16(t_set + (uint32_t)(orow * BN), v);
          if (last_row) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
          }
          if (p.dot_out)
            atomicAdd(p.dot_out + ((int64_t)n * p.H + (h0 + orow)) * p.W + (w0 + px),
                      dot_activated<BN>(v, bias_s, dot_s, nt * BN, slope));
          epilogue_store<(BN >= 32 ? 2 : 1)>(v, bias_s, nt * BN, p.Cout, slope, p.out_hi + obase, p.out_lo + obase);
        }
      }
      if (++acc == kAccSets) {
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
  // 64 output channels per tile for the decoder layers with a fused upsample: one N = 192 MMA per product instead of two
  // N = 96 ones (tensor-bound instead of shared-memory-bound) and every input row is interpolated once instead of once per
  // N tile (dec2: 3.6 -> 2.8 ms).  Plain TMA layers stay at 32: the 64-wide tile has a single accumulator set (no
  // epilogue overlap) and only four operand slots next to its 147 KB of weights, and measured slower there.
  R.BN = cout16 == 16 ? 16 : (L.rows_wide && cout16 % 64 == 0 ? 64 : 32);
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
static bool rows_activation_map(const ActView& v, CUtensorMap* m, std::string& err, const std::string& name) {
  cuuint64_t dims[5] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N, 2};
  const int64_t plane = (const char*)v.lo - (const char*)v.hi;
  if (plane <= 0 || plane % 16) {
    err = "tc_rows_launch: hi/lo planes must be 16-byte aligned with lo after hi";
    return false;
  }
  cuuint64_t strides[4] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2, (cuuint64_t)plane};
  cuuint32_t box[5] = {kKB, (cuuint32_t)kBoxPx, 1, 1, 2};   // both planes of one row in one box
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = tc_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)v.hi, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    err = "cuTensorMapEncodeTiled(row-kernel activations) failed for " + name + " code " + std::to_string((int)r);
    return false;
  }
  return true;
}

// extra != nullptr: the LAST channel chunk is read from *extra (channels [0, extra->C), zero-filled up to the chunk).
cudaError_t tc_rows_launch(ConvLayer& L, TcConv& tc, const ActView& in, const ActView& out, cudaStream_t s,
                           std::string& err, const ActView* up_src, const ActView* extra) {
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
    if (!rows_activation_map(in, &m, err, L.name)) return cudaErrorInvalidValue;
    it = R.map_a.emplace(key, m).first;
  }
  auto itl = it;
  if (extra) {
    if (extra->H != in.H || extra->W != in.W || extra->N != in.N || extra->C > R.KB || extra->sw % 8) {
      err = "tc_rows_launch: the extra last-chunk tensor must match the input geometry and fit one chunk";
      return cudaErrorInvalidValue;
    }
    ViewKey kl = std::make_tuple((const void*)extra->hi, (const void*)extra->lo, extra->N, extra->H, extra->W, extra->C);
    itl = R.map_l.find(kl);
    if (itl == R.map_l.end()) {
      CUtensorMap m;
      if (!rows_activation_map(*extra, &m, err, L.name)) return cudaErrorInvalidValue;
      itl = R.map_l.emplace(kl, m).first;
    }
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
  p.x_hi = p.x_lo = nullptr; p.xsn = p.xsh = 0; p.xsw = 0;
  p.trace = g_tc_debug[0] == 1 ? 1 : 0;
  p.up_sh = p.up_sw = 0.f;
  p.dot_w = L.dot_w; p.dot_out = L.dot_w ? L.dot_out : nullptr;
  p.a_c_off = 0;
  p.l_chunk = extra ? R.chunks - 1 : -1;
  p.kmask = g_tc_debug[6] == 1 ? R.kmask : ~0ull;   // VR_KSKIP=0 issues the all-zero-weight channel groups too
  if (up_src) {
    // `in` is either the whole concat buffer (its first up_src->C channels are then never read) or only the skip
    // tensor, which starts at reduction index up_src->C
    if (in.C + up_src->C + (extra ? R.KB : 0) <= R.CinPadR) p.a_c_off = -up_src->C;
    p.up_chunks = up_src->C / 32;
    p.xH = up_src->H; p.xW = up_src->W;
    p.x_hi = up_src->hi; p.x_lo = up_src->lo;
    p.xsn = up_src->sn; p.xsh = up_src->sh; p.xsw = up_src->sw;
    if ((reinterpret_cast<uintptr_t>(up_src->hi) | reinterpret_cast<uintptr_t>(up_src->lo)) & 15) {
      err = "tc_rows_launch: upsample source planes must be 16-byte aligned";
      return cudaErrorInvalidValue;
    }
    p.up_sh = in.H > 1 ? (float)(up_src->H - 1) / (float)(in.H - 1) : 0.f;   // as launch_upsample2x
    p.up_sw = in.W > 1 ? (float)(up_src->W - 1) / (float)(in.W - 1) : 0.f;
  }
  const TcDevice& dv = tc_device();
  if (!dv.ok) {
    err = "tc_rows_launch: cannot query the current device";
    return cudaErrorInvalidValue;
  }
  const int b_bytes = R.BN == 16 ? (int)(2 * RowsGeom<16>::kBBuf) : R.BN == 32 ? (int)(2 * RowsGeom<32>::kBBuf) : (int)(2 * RowsGeom<64>::kBBuf);
  p.n_aslots = (dv.max_smem - 3072 - 1024 - b_bytes) / (int)kASlot;
  if (p.n_aslots > kMaxASlots) p.n_aslots = kMaxASlots;
  if (p.n_aslots < (p.up_chunks > 0 ? 4 : 2)) {
    err = "tc_rows_launch: shared memory too small";
    return cudaErrorInvalidValue;
  }
  // the interpolation ring gets the larger share: most chunks of the decoder layers are up-sampled
  p.n_uslots = p.up_chunks > 0 ? p.n_aslots - p.n_aslots / 2 : 0;
  if (p.up_chunks > 0 && p.n_aslots - p.n_uslots < 2) p.n_uslots = p.n_aslots - 2;
  if (p.up_chunks > 0 && g_tc_debug[4] >= 1 && g_tc_debug[4] <= p.n_aslots - 2) p.n_uslots = g_tc_debug[4];
  const int dyn = p.n_aslots * (int)kASlot + b_bytes + 1024;
  const int grid = p.total_tiles < dv.num_sms ? p.total_tiles : dv.num_sms;   // persistent: one CTA per SM
  if (R.BN == 16)
    conv_tc_rows_kernel<16><<<grid, kRowsThreads, dyn, s>>>(it->second, R.map_b, itl->second, p);
  else if (R.BN == 32)
    conv_tc_rows_kernel<32><<<grid, kRowsThreads, dyn, s>>>(it->second, R.map_b, itl->second, p);
  else
    conv_tc_rows_kernel<64><<<grid, kRowsThreads, dyn, s>>>(it->second, R.map_b, itl->second, p);
  return cudaGetLastError();
}

// copies the timeline of the last traced launch (vr_debug_set(0, 1)) to the host: 3 roles x kTraceEvents x 3 stamps
int tc_rows_read_trace(unsigned long long* out, long long capacity) {
  const long long n = 3LL * kTraceEvents * 3;
  if (capacity < n) return -1;
  if (cudaMemcpyFromSymbol(out, g_rows_trace, sizeof(unsigned long long) * (size_t)n) != cudaSuccess) return -1;
  return (int)n;
}

// cudaFuncSetAttribute is per device: called by tc_device() the first time a device is used (conv_tc.cu)
void tc_rows_set_attributes(int max_smem) {
  cudaFuncSetAttribute(conv_tc_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
  cudaFuncSetAttribute(conv_tc_rows_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(conv_tc_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
  cudaFuncSetAttribute(conv_tc_rows_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(conv_tc_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 3072);
  cudaFuncSetAttribute(conv_tc_rows_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

}  // namespace vr
