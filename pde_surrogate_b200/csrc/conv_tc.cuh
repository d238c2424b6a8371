// conv_tc.cuh — interface of the tcgen05 convolution kernels.
#pragma once
#include "conv.cuh"

namespace pdes {

struct TcConvArgs {
  ConvArgs c;        // geometry, prologue, epilogue (weights pointer `w` unused)
  const float* wtc;  // filter tiles packed by pack_tc_kernel
  int N;             // GEMM N = output channels padded to a multiple of 16 (<= 256)
  int KC;            // input channels per pipeline chunk (8, 16 or 32)
  int NB;            // filter-tile ring depth
  int nchunks;
  int S;             // accumulator sets in TMEM (K range spread over S accumulators)
  int prec;          // 0 = 3xTF32 (fp32 parity), 1 = single-pass TF32
};

struct TcPlan {
  int KC, nchunks, NB, S;
  size_t smem, pack_floats;
};

struct TcPackDesc {
  const float* w;  // OIHW
  float* dst;
  int Cout, Cin, KS, N, KC, nchunks, transpose;
};

struct TcWgradArgs {
  WgradArgs w;       // geometry / prologue (dw unused)
  float* dwp;        // staging gradient [tap][ci_pad][co_pad], accumulated with vector reductions
  int ci_pad, co_pad;
  int dbg;           // debug switches (env PDES_WG_DBG), 0 in production
};

struct TcWgradUnpack {
  float* dw;         // OIHW gradient (+=)
  float* dwp;        // staging buffer (read, then cleared)
  int Cout, Cin, KS, ci_pad, co_pad;
};

bool wgrad_tc_supported(int KS, int stride);
void wgrad_tc_dims(int Cin, int Cout, int* ci_pad, int* co_pad);
int launch_wgrad_tc(const TcWgradArgs& t, cudaStream_t st);
int launch_wgrad_unpack(const TcWgradUnpack* dev_table, int n, int max_elems, cudaStream_t st);

// tiling for a convolution whose GEMM-K operand has Cin_k channels and GEMM-N is N
void tc_plan(int KS, int Cin_k, int N, TcPlan* p);
bool tc_supported(int KS, int stride, int Cin_k, int N);
int launch_conv_tc(const TcConvArgs& t, cudaStream_t st);
int launch_pack_tc(const TcPackDesc* dev_table, int n, size_t max_elems, cudaStream_t st);

}  // namespace pdes
