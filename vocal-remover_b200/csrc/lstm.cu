// LSTM branch of BaseNet (reference lib/layers.py:108-133, wired at lib/nets.py:23,38):
//   1x1 conv (2n -> 1) + BN + ReLU  ->  (T, N, bins)  ->  BiLSTM(hidden = nout_lstm/2, gate order i,f,g,o)
//   -> Linear(nout_lstm -> bins) + BatchNorm1d(eval) + ReLU -> one extra channel of the dec1 input.
// 0.18 % of the FLOPs but a 128-step sequential dependency: the input projection is hoisted into
// one GEMM, and the recurrence runs as one persistent CTA per (window, direction) with its W_hh row
// held in registers and h exchanged through shared memory.  All math fp32 with accurate expf/tanhf.
#include "common.cuh"
#include "kernels.h"

namespace vr {

// ------------------------------------------------------------------------------------------------
__global__ void lstm_inconv_kernel(ActView in, const float* __restrict__ w, float bias, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)in.N * in.H * in.W;
  if (idx >= total) return;
  int t = (int)(idx % in.W);
  int64_t r = idx / in.W;
  int bin = (int)(r % in.H);
  int n = (int)(r / in.H);
  int64_t o = (int64_t)n * in.sn + (int64_t)bin * in.sh + (int64_t)t * in.sw;
  float acc = 0.f;
  for (int c = 0; c < in.C; c += 8) {
    float x[8];
    load8(in.hi + o + c, in.lo + o + c, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(x[i], __ldg(w + c + i), acc);
  }
  out[((int64_t)n * in.W + t) * in.H + bin] = fmaxf(acc + bias, 0.f);
}

cudaError_t launch_lstm_inconv(ActView in, const float* w, float bias, float* out, cudaStream_t stream) {
  int64_t total = (int64_t)in.N * in.H * in.W;
  if (total == 0) return cudaSuccess;
  lstm_inconv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, w, bias, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] * B[N][K]^T + bias[N]; 64x64 tile, 16-deep k slab, 256 threads x (4x4).
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                      const float* __restrict__ bias, float* __restrict__ C, int M,
                                                      int N, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int r = i >> 4, c = i & 15;
      int gm = m0 + r, gn = n0 + r, gk = k0 + c;
      As[c][r] = (gm < M && gk < K) ? A[(int64_t)gm * K + gk] : 0.f;
      Bs[c][r] = (gn < N && gk < K) ? B[(int64_t)gn * K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn < N) C[(int64_t)gm * N + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
    }
  }
}

cudaError_t launch_gemm_nt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                           cudaStream_t stream) {
  if (M == 0 || N == 0) return cudaSuccess;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
  gemm_nt_kernel<<<grid, 256, 0, stream>>>(A, B, bias, C, M, N, K);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// grid (N, 2 directions), block 4*HID threads.  Thread j owns gate row j of W_hh (registers).
template <int HID>
__global__ void __launch_bounds__(4 * HID) lstm_recurrence_kernel(const float* __restrict__ xp,
                                                                  const float* __restrict__ whh,
                                                                  float* __restrict__ hs, int T) {
  const int n = blockIdx.x, dir = blockIdx.y, j = threadIdx.x;
  __shared__ float h_s[HID];
  __shared__ float g_s[4 * HID];
  float wrow[HID];
#pragma unroll
  for (int k = 0; k < HID; ++k) wrow[k] = whh[((int64_t)dir * 4 * HID + j) * HID + k];
  if (j < HID) h_s[j] = 0.f;
  float c = 0.f;
  __syncthreads();
  const float* xp_n = xp + (int64_t)n * T * 8 * HID + dir * 4 * HID + j;
  int t = dir ? T - 1 : 0;
  float xnext = xp_n[(int64_t)t * 8 * HID];
  for (int s = 0; s < T; ++s) {
    const int tn = dir ? t - 1 : t + 1;
    // four independent partial sums: the 4-cycle FMA latency chain is HID/4 long instead of HID
    float g0 = xnext, g1 = 0.f, g2 = 0.f, g3 = 0.f;
    if (s + 1 < T) xnext = xp_n[(int64_t)tn * 8 * HID];
#pragma unroll
    for (int k = 0; k < HID; k += 4) {
      g0 = fmaf(wrow[k], h_s[k], g0);
      g1 = fmaf(wrow[k + 1], h_s[k + 1], g1);
      g2 = fmaf(wrow[k + 2], h_s[k + 2], g2);
      g3 = fmaf(wrow[k + 3], h_s[k + 3], g3);
    }
    g_s[j] = (g0 + g1) + (g2 + g3);
    __syncthreads();
    if (j < HID) {
      float ig = sigmoid_acc(g_s[j]);
      float fg = sigmoid_acc(g_s[HID + j]);
      float gg = tanhf(g_s[2 * HID + j]);
      float og = sigmoid_acc(g_s[3 * HID + j]);
      c = fg * c + ig * gg;
      float h = og * tanhf(c);
      h_s[j] = h;
      hs[((int64_t)n * T + t) * 2 * HID + dir * HID + j] = h;
    }
    __syncthreads();
    t = tn;
  }
}

cudaError_t launch_lstm_recurrence(const float* xp, const float* whh, float* hs, int N, int T, int hid,
                                   cudaStream_t stream) {
  if (N == 0) return cudaSuccess;
  dim3 grid((unsigned)N, 2);
  switch (hid) {
    case 8: lstm_recurrence_kernel<8><<<grid, 32, 0, stream>>>(xp, whh, hs, T); break;
    case 16: lstm_recurrence_kernel<16><<<grid, 64, 0, stream>>>(xp, whh, hs, T); break;
    case 32: lstm_recurrence_kernel<32><<<grid, 128, 0, stream>>>(xp, whh, hs, T); break;
    case 64: lstm_recurrence_kernel<64><<<grid, 256, 0, stream>>>(xp, whh, hs, T); break;
    case 128: lstm_recurrence_kernel<128><<<grid, 512, 0, stream>>>(xp, whh, hs, T); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// grid (T, N); block = 128 threads striding over bins.  wd is stored transposed: wdT[k][bin].
__global__ void lstm_dense_kernel(const float* __restrict__ hs, const float* __restrict__ wdT,
                                  const float* __restrict__ scale, const float* __restrict__ shift, int T, int K,
                                  int bins, ActView dst, int ch) {
  extern __shared__ float h_s[];
  const int t = blockIdx.x, n = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) h_s[k] = hs[((int64_t)n * T + t) * K + k];
  __syncthreads();
  for (int bin = threadIdx.x; bin < bins; bin += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(h_s[k], __ldg(wdT + (int64_t)k * bins + bin), acc);
    float y = fmaxf(fmaf(acc, scale[bin], shift[bin]), 0.f);
    int64_t o = (int64_t)n * dst.sn + (int64_t)bin * dst.sh + (int64_t)t * dst.sw + ch;
    split_bf16(y, dst.hi[o], dst.lo[o]);
  }
}

cudaError_t launch_lstm_dense(const float* hs, const float* wdT, const float* scale, const float* shift, int N, int T,
                              int K, int bins, ActView dst, int ch, cudaStream_t stream) {
  if (N == 0) return cudaSuccess;
  dim3 grid((unsigned)T, (unsigned)N);
  lstm_dense_kernel<<<grid, 128, K * sizeof(float), stream>>>(hs, wdT, scale, shift, T, K, bins, dst, ch);
  return cudaGetLastError();
}

}  // namespace vr
