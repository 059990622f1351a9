// Shared device/host definitions for the vocal-remover B200 hot path (sm_100a only).
//
// Activation storage ("split-bf16"): every activation tensor of the CascadedNet forward
// (reference lib/nets.py:82-117) lives in HBM as TWO NHWC bf16 planes, hi = bf16(x) and
// lo = bf16(x - hi).  hi+lo carries a 16-bit significand, which is what lets the tcgen05
// kind::f16 tensor-core convolution (conv_tc.cu) reach the 1e-3 mask parity gate with three
// bf16 passes (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM); a single bf16/fp16 pass fails it
// (DESIGN.md "Precision").  The planes cost the same 4 B/element as fp32 and are directly
// TMA-loadable as tensor-core operands.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vr {

typedef __nv_bfloat16 bf16;

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// View of an NHWC split-bf16 tensor (possibly a channel slice / band of a larger buffer).
struct ActView {
  bf16* hi;
  bf16* lo;
  int N, H, W, C;   // logical extent of the view (C = channels visible through it)
  int64_t sn;       // element stride between images
  int64_t sh;       // element stride between rows (H)
  int sw;           // element stride between pixels (= channel count of the underlying buffer)
};

// Parameters of one fused Conv2d(bias=False)+BatchNorm2d(eval)+activation layer
// (reference lib/layers.py:8-26) in implicit-GEMM form.
struct ConvParams {
  ActView in;        // in.C == CinPad (multiple of 8; zero weights on pad channels)
  ActView out;       // out.C == number of output channels stored (Cout)
  const float* w;    // fp32 [taps][CinPad][CoutPad], BN scale folded in
  const float* bias; // fp32 [CoutPad], folded BN shift
  int CinPad, Cout, CoutPad;
  int KH, KW, stride, pad_h, pad_w, dil_h, dil_w;
  int act;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;   // nn.LeakyReLU default slope (lib/layers.py:31)
  return v;
}

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float join_bf16(bf16 hi, bf16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

// 8 consecutive bf16 (16 bytes) <-> 8 floats
struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const bf16* hi, const bf16* lo, float* x) {
  bf16x8 a = *reinterpret_cast<const bf16x8*>(hi);
  bf16x8 b = *reinterpret_cast<const bf16x8*>(lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 fa = __bfloat1622float2(a.v[i]);
    float2 fb = __bfloat1622float2(b.v[i]);
    x[2 * i] = fa.x + fb.x;
    x[2 * i + 1] = fa.y + fb.y;
  }
}

__device__ __forceinline__ void split8(const float* x, bf16x8& h, bf16x8& l) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    float2 hf = __bfloat1622float2(hh);
    h.v[i] = hh;
    l.v[i] = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
  }
}

// Store `cnt` consecutive channels (fp32 values) as split-bf16 at element offset `off` of a pixel.
__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, const float* x, int cnt) {
  if (cnt == 8 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
    bf16x8 h, l;
    split8(x, h, l);
    *reinterpret_cast<bf16x8*>(hi) = h;
    *reinterpret_cast<bf16x8*>(lo) = l;
  } else {
    for (int i = 0; i < cnt; ++i) split_bf16(x[i], hi[i], lo[i]);
  }
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace vr
