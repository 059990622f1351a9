// Times the two GEMMs and the recurrence of the LSTM branch (csrc/lstm.cu) in isolation at the stage-3 shapes of a
// 27-window batch: M = 27 * 128 (n, t) rows, bins = 512, hidden 64 per direction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../vocal-remover_b200/csrc -o lstm_gemm lstm_gemm.cu
#include <cstdio>
#include <vector>
#include "../../vocal-remover_b200/csrc/lstm.cu"
using namespace vr;
static float time_ms(cudaStream_t s, int reps, auto&& fn) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) fn();
  cudaEventRecord(a, s);
  for (int i = 0; i < reps; ++i) fn();
  cudaEventRecord(b, s);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}
int main() {
  const int N = 27, T = 128, hid = 64;
  for (int bins : {512, 256}) {
    const int gates = 8 * hid, NT = N * T;
    float *l0, *wih, *bih, *xp, *whh, *hs, *wd, *sc, *sh, *y;
    cudaMalloc(&l0, sizeof(float) * NT * bins); cudaMalloc(&wih, sizeof(float) * gates * bins); cudaMalloc(&bih, sizeof(float) * gates);
    cudaMalloc(&xp, sizeof(float) * NT * gates); cudaMalloc(&whh, sizeof(float) * 2 * 4 * hid * hid); cudaMalloc(&hs, sizeof(float) * NT * 2 * hid);
    cudaMalloc(&wd, sizeof(float) * bins * 2 * hid); cudaMalloc(&sc, sizeof(float) * bins); cudaMalloc(&sh, sizeof(float) * bins);
    cudaMalloc(&y, sizeof(float) * NT * bins);
    cudaMemset(l0, 0, sizeof(float) * NT * bins); cudaMemset(wih, 0, sizeof(float) * gates * bins); cudaMemset(bih, 0, sizeof(float) * gates);
    cudaMemset(whh, 0, sizeof(float) * 8 * hid * hid); cudaMemset(wd, 0, sizeof(float) * bins * 2 * hid); cudaMemset(sc, 0, sizeof(float) * bins);
    cudaMemset(sh, 0, sizeof(float) * bins);
    cudaStream_t s; cudaStreamCreate(&s);
    const float t1 = time_ms(s, 20, [&] { launch_lstm_input_projection(l0, 0.1f, wih, bih, xp, N, T, bins, gates, s); });
    const float t2 = time_ms(s, 20, [&] { launch_lstm_recurrence(xp, whh, hs, N, T, hid, s); });
    const float t3 = time_ms(s, 20, [&] { launch_lstm_dense(hs, wd, sc, sh, NT, 2 * hid, bins, y, s); });
    printf("bins %d: input projection %.1f us (%.1f TFLOP/s)  recurrence %.1f us  dense %.1f us (%.1f TFLOP/s)  %s\n", bins, t1 * 1e3,
           2.0 * NT * gates * bins / (t1 * 1e-3) / 1e12, t2 * 1e3, t3 * 1e3, 2.0 * NT * bins * 2 * hid / (t3 * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
