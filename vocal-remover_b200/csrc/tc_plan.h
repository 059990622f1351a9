// Host-side plan of one convolution layer for the tcgen05 kernels (conv_tc.cu, conv_tc_rows.cu).
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "engine.h"

namespace vr {

typedef std::tuple<const void*, const void*, int, int, int, int> ViewKey;

struct TcRowsPlan {   // row-streaming variant: 3x3, stride 1, dilation 1, W % 128 == 0 (conv_tc_rows.cu)
  bool ok = false;
  int KB = 32, CinPadR = 0, chunks = 0, BN = 0, n_tiles = 0;
  unsigned long long kmask = ~0ull;   // bit g: input channels [8g, 8g+8) carry a non-zero weight
  bf16* w_planes = nullptr;   // [2][n_tiles*BN][9*CinPadR]
  float* bias = nullptr;      // [n_tiles*BN]
  CUtensorMap map_b;
  std::map<ViewKey, CUtensorMap> map_a;
};

struct TcFlatPlan {   // flat-halo variant: 3x3, stride 1, W <= 64 with zero pad pixels after every row (conv_tc_flat.cu)
  bool ok = false;
  CUtensorMap map_b;
  std::map<ViewKey, CUtensorMap> map_a;
};

struct TcConv {
  int CinPadTC = 0, KB = 0, cchunks = 0, SUBS = 0, taps = 0, Ktot = 0, CoutPadN = 0, BN = 0, n_tiles = 0;
  bf16* w_planes = nullptr;   // [2][CoutPadN][Ktot]
  float* bias = nullptr;      // [n_tiles*BN]
  CUtensorMap map_b;
  std::map<ViewKey, CUtensorMap> map_a;
  TcRowsPlan rows;
  TcFlatPlan flat;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_encode_fn();
uint16_t tc_f2bf(float f);
float tc_bf2f(uint16_t h);

// conv_tc_rows.cu
bool tc_rows_prepare(ConvLayer& L, TcConv& tc, std::string& err, std::vector<void*>& allocs);
bool tc_rows_supported(const ConvLayer& L, const TcConv& tc, const ActView& in, const ActView& out);
cudaError_t tc_rows_launch(ConvLayer& L, TcConv& tc, const ActView& in, const ActView& out, cudaStream_t s,
                           std::string& err, const ActView* up_src = nullptr);
// conv_tc_flat.cu
bool tc_flat_prepare(ConvLayer& L, TcConv& tc, std::string& err);
bool tc_flat_supported(const ConvLayer& L, const TcConv& tc, const ActView& in, const ActView& out);
cudaError_t tc_flat_launch(ConvLayer& L, TcConv& tc, const ActView& in, const ActView& out, cudaStream_t s,
                           std::string& err);
extern int g_tc_debug[8];   // [0] unused, [1] disable the row kernel, [2] = 64: 64-channel chunks, [3] = 1 enables the experimental flat-halo kernel, [4] = k: k of the row slots feed the interpolation warps (default half), [5] = 1 (default): decoder upsample fused into the row kernel, [6] = 1: the row kernel skips channel groups whose weights are all zero, [7] = 1: tensor-core convolutions are launched with programmatic stream serialization (prologue overlaps the previous kernel's tail)

}  // namespace vr
