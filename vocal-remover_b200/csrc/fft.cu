// Framed Hann-windowed STFT and window-sum-square-normalised overlap-add inverse STFT.
//
// Device replacement for the reference's two librosa calls (lib/spec_utils.py:26-31 and 157-165;
// librosa 0.10 semantics restated in SURVEY.md App. A and oracle/stft_oracle.py):
//   stft : zero centre padding n_fft/2, frame t = y_p[t*hop : t*hop+n_fft], periodic Hann, rfft
//          -> complex64 [2][n_fft/2+1][T], T = 1 + L/hop
//   istft: irfft(1/N) * Hann, overlap-add at hop, / sum(Hann^2) where > tiny, trim n_fft/2 both ends
// One CTA transforms one frame with a shared-memory radix-2 FFT.  Two real signals ride in one
// complex transform: the stereo pair (L + iR) forward, the two stems (instruments + i*vocals)
// inverse, so the mask multiply of Separator._postprocess (inference.py:32-36) is fused into the
// inverse transform's load.
#include "common.cuh"
#include "kernels.h"

namespace vr {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place radix-2 DIT on bit-reversed data in shared memory.  tw[q] = exp(-2 pi i q / NF), q < NF/2.
// INVERSE uses conj(tw).
template <bool INVERSE>
__device__ __forceinline__ void fft_inplace(float2* z, const float2* __restrict__ tw, int NF, int logn) {
  for (int s = 1; s <= logn; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = NF >> s;
    for (int k = threadIdx.x; k < (NF >> 1); k += blockDim.x) {
      const int j = k & (half - 1);
      const int i0 = ((k >> (s - 1)) << s) + j;
      const int i1 = i0 + half;
      float2 w = __ldg(tw + j * tstep);
      if (INVERSE) w.y = -w.y;
      const float2 a = z[i0];
      const float2 b = cmul(z[i1], w);
      z[i0] = make_float2(a.x + b.x, a.y + b.y);
      z[i1] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ wave, int64_t L, int NF, int logn,
                                                   int hop, float2* __restrict__ spec, int64_t T, int64_t t_first,
                                                   const float2* __restrict__ tw, const float* __restrict__ win) {
  extern __shared__ float2 z[];
  const int64_t t = t_first + blockIdx.x;
  const int64_t s0 = t * hop - NF / 2;
  for (int n = threadIdx.x; n < NF; n += blockDim.x) {
    const int64_t s = s0 + n;
    float l = 0.f, r = 0.f;
    if (s >= 0 && s < L) {
      l = wave[s];
      r = wave[L + s];
    }
    const float w = win[n];
    z[__brev((unsigned)n) >> (32 - logn)] = make_float2(l * w, r * w);
  }
  __syncthreads();
  fft_inplace<false>(z, tw, NF, logn);
  const int bins = NF / 2 + 1;
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    const float2 a = z[k];
    const float2 b = z[(NF - k) & (NF - 1)];
    spec[((int64_t)0 * bins + k) * T + t] = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    spec[((int64_t)1 * bins + k) * T + t] = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
  }
}

cudaError_t launch_stft(const float* wave, int64_t L, int n_fft, int hop, float2* spec, int64_t T, int64_t t0,
                        int64_t t1, const float2* twiddle, const float* window, cudaStream_t stream) {
  if (t1 <= t0) return cudaSuccess;
  int logn = 0;
  while ((1 << logn) < n_fft) ++logn;
  if ((1 << logn) != n_fft || n_fft > 4096 || n_fft < 64) return cudaErrorInvalidValue;
  stft_kernel<<<(unsigned)(t1 - t0), 256, n_fft * sizeof(float2), stream>>>(wave, L, n_fft, logn, hop, spec, T, t0,
                                                                             twiddle, window);
  return cudaGetLastError();
}

// grid (T, 2 channels): windowed inverse transform of one frame of one channel -> frames[stem][c][t][NF]
__global__ void __launch_bounds__(256) istft_frames_kernel(const float2* __restrict__ spec,
                                                           const float* __restrict__ mask, int NF, int logn,
                                                           int64_t T, int64_t t_first, int64_t nfr,
                                                           float* __restrict__ frames_a,
                                                           float* __restrict__ frames_b,
                                                           const float2* __restrict__ tw,
                                                           const float* __restrict__ win) {
  extern __shared__ float2 z[];
  const int64_t t = t_first + blockIdx.x;
  const int c = blockIdx.y;
  const int bins = NF / 2 + 1;
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    const int64_t gi = ((int64_t)c * bins + k) * T + t;
    float2 x = spec[gi];
    float2 ya, yb;
    if (mask != nullptr) {
      const float m = mask[gi];
      ya = make_float2(m * x.x, m * x.y);
      const float q = 1.f - m;
      yb = make_float2(q * x.x, q * x.y);
    } else {
      ya = x;
      yb = make_float2(0.f, 0.f);
    }
    if (k == 0 || k == NF / 2) {  // c2r transforms ignore the imaginary part of DC and Nyquist
      ya.y = 0.f;
      yb.y = 0.f;
    }
    const unsigned r0 = __brev((unsigned)k) >> (32 - logn);
    z[r0] = make_float2(ya.x - yb.y, ya.y + yb.x);
    if (k != 0 && k != NF / 2) {
      const unsigned r1 = __brev((unsigned)(NF - k)) >> (32 - logn);
      z[r1] = make_float2(ya.x + yb.y, -ya.y + yb.x);
    }
  }
  __syncthreads();
  fft_inplace<true>(z, tw, NF, logn);
  const float inv = 1.f / (float)NF;
  float* fa = frames_a + ((int64_t)c * nfr + blockIdx.x) * NF;
  float* fb = frames_b ? frames_b + ((int64_t)c * nfr + blockIdx.x) * NF : nullptr;
  for (int n = threadIdx.x; n < NF; n += blockDim.x) {
    const float w = win[n] * inv;
    fa[n] = z[n].x * w;
    if (fb) fb[n] = z[n].y * w;
  }
}

// Output samples [s0, s1) of each channel; the scratch holds frames [t_first, t_first + nfr) of the track.
__global__ void istft_ola_kernel(const float* __restrict__ frames_a, const float* __restrict__ frames_b, int NF,
                                 int hop, int64_t T, int64_t t_first, int64_t nfr, int64_t s0, int64_t s1, int64_t Lo,
                                 float* __restrict__ wave_a, float* __restrict__ wave_b,
                                 const float* __restrict__ win) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t span = s1 - s0;
  if (idx >= 2 * span) return;
  const int c = (int)(idx / span);
  const int64_t s = s0 + idx % span;
  const int64_t u = s + NF / 2;
  int64_t t1 = u / hop;
  if (t1 > T - 1) t1 = T - 1;
  int64_t t0 = (u - NF + hop) / hop;   // ceil((u - NF + 1) / hop) for u - NF + 1 > 0
  if (u - NF + 1 <= 0) t0 = 0;
  float acc_a = 0.f, acc_b = 0.f, wss = 0.f;
  for (int64_t t = t0; t <= t1; ++t) {
    const int n = (int)(u - t * hop);
    const float w = win[n];
    wss += w * w;
    const int64_t fi = ((int64_t)c * nfr + (t - t_first)) * NF + n;
    acc_a += frames_a[fi];
    if (frames_b) acc_b += frames_b[fi];
  }
  if (wss > 1.17549435e-38f) {   // np.finfo(float32).tiny
    acc_a /= wss;
    acc_b /= wss;
  }
  wave_a[(int64_t)c * Lo + s] = acc_a;
  if (wave_b) wave_b[(int64_t)c * Lo + s] = acc_b;
}

// frames scratch is provided by the caller through wave-independent workspace (see engine.cu)
cudaError_t launch_istft_frames(const float2* spec, const float* mask, int n_fft, int64_t T, int64_t t_first,
                                int64_t nfr, float* frames_a, float* frames_b, const float2* twiddle,
                                const float* window, cudaStream_t stream) {
  if (nfr <= 0) return cudaSuccess;
  int logn = 0;
  while ((1 << logn) < n_fft) ++logn;
  if ((1 << logn) != n_fft || n_fft > 4096 || n_fft < 64) return cudaErrorInvalidValue;
  dim3 grid((unsigned)nfr, 2);
  istft_frames_kernel<<<grid, 256, n_fft * sizeof(float2), stream>>>(spec, mask, n_fft, logn, T, t_first, nfr, frames_a,
                                                                     frames_b, twiddle, window);
  return cudaGetLastError();
}

cudaError_t launch_istft_ola(const float* frames_a, const float* frames_b, int n_fft, int hop, int64_t T,
                             int64_t t_first, int64_t nfr, int64_t s0, int64_t s1, float* wave_a, float* wave_b,
                             const float* window, cudaStream_t stream) {
  const int64_t Lo = (int64_t)hop * (T - 1);
  if (s1 <= s0) return cudaSuccess;
  istft_ola_kernel<<<(unsigned)((2 * (s1 - s0) + 255) / 256), 256, 0, stream>>>(frames_a, frames_b, n_fft, hop, T, t_first,
                                                                                nfr, s0, s1, Lo, wave_a, wave_b, window);
  return cudaGetLastError();
}

}  // namespace vr
