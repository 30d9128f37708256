// Interface between the C-ABI dispatch (conv_mma.cu) and the tcgen05 / TMA implicit-GEMM kernel (conv_sm100.cu).
#pragma once
#include <cuda_runtime.h>

namespace sm100 {

struct Problem {
  // gathered tensor (NHWC bf16, channel slice): N x H x W x C with channel pitch a_pitch (elements)
  const void* a;
  int N, H, W, C, a_pitch;
  // B matrix: b_rows = GEMM N (output channels of this GEMM), b_cols = taps * b_cols_per_tap, K-major bf16
  const void* b;
  int b_rows, b_cols, b_cols_per_tap;
  int R, S, stride, pad, P, Q, flip;
  // optional explicit tap table (ntaps > 0): im2col offsets (dh, dw) and B column block of each tap
  int ntaps;
  int tap_dh[9], tap_dw[9], tap_b[9];
  // strided output rows (see conv_sm100.cu: Params::out_mode)
  int out_mode, o_mul, oh_add, ow_add, outH, outW;
  void* y;
  int y_pitch, y_off;
  const float* scale;
  const float* shift;
  const void* residual;
  double* stats;
  int stats_repl, act;
};

struct WgradProblem {
  const void* x;   // NHWC bf16 slice, N x H x W x C
  const void* dy;  // NHWC bf16 slice, N x P x Q x K
  int N, H, W, C, x_pitch;
  int K, y_pitch;
  int R, S, stride, pad, P, Q;
  float* dw;       // fp32 [K][R][S][C], accumulated into
};

bool enabled();
bool wgrad_supported(const WgradProblem& q);
int wgrad_launch(const WgradProblem& q, cudaStream_t st);
bool supported(const Problem& q);
int launch(const Problem& q, cudaStream_t st);
long long launch_count();
int read_trace(long long* host_out);  // 12 x 512 SM-clock stamps of the last SGB_DEBUG_SKIP&16 launch

}  // namespace sm100
