// Host-side helpers shared by the C-ABI translation units: error reporting, dtype tags, TMA map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace dvae {

// activation storage: bf16 or fp16 (tcgen05 kind::f16), or fp32 kept on the tf32 grid (kind::tf32)
enum DType : int { kBF16 = 0, kTF32 = 1, kF16 = 2, kF32 = 3 };   // kF32: strict fp32, CUDA cores (ops_simt.cu)

void set_last_error(const std::string& msg);

#define DVAE_CHECK_CUDA(expr)                                                                         \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      ::dvae::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +     \
                             __FILE__ + ":" + std::to_string(__LINE__));                              \
      return 2;                                                                                       \
    }                                                                                                 \
  } while (0)

#define DVAE_REQUIRE(cond, msg)                                                                       \
  do {                                                                                                \
    if (!(cond)) {                                                                                    \
      ::dvae::set_last_error(std::string("invalid argument: ") + (msg) + " [" #cond "] at " +         \
                             __FILE__ + ":" + std::to_string(__LINE__));                              \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

// rank-3 tiled tensor map, 128-byte swizzle, zero fill out of bounds.  dims/box are in elements
// (innermost first); strides in bytes for dims 1 and 2.
// mn_major: the operand is consumed MN-major by tcgen05.mma; 32-bit MN-major operands need the 32-byte-atom
// flavour of the 128-byte swizzle.
int encode_map3(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                bool mn_major = false);
// same, for boxes narrower than one swizzle row (inner box < 128 bytes, a multiple of 16): no swizzle
int encode_map3_narrow(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);

inline int ceil_div(long a, long b) { return static_cast<int>((a + b - 1) / b); }

int num_sms();

// BatchNorm statistics pass (ops_pointwise.cu): per-half column sums / sums of squares of y [halves*rows_half, C] into
// ws [halves][2][C] doubles (accumulating; the caller zeroes ws)
int bn_stats_launch(int dtype, const void* y, double* ws, int rows_half, int halves, int C, cudaStream_t st);

// Sequence-resident LSTM recurrence (ops_lstm_seq.cu): one launch for all T steps when W_hh fits in shared memory.
bool lstm_seq_supported(int H, int T);
void lstm_seq_set_stamps(long long* buf);
int lstm_seq_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, int D,
                 cudaStream_t st);
int lstm_seq_bwd(int dtype, const void* dh_all, const void* gates, const float* c_all, const void* whh_n, void* da_all,
                 int rows, int T, int H, int D, cudaStream_t st);

// Weight-stationary, time-resident forward recurrence for H = 512 / 1024 (ops_lstm_res.cu): one launch per layer, CTA pairs
// keep their slice of W_hh in shared memory and hand h_t to each other through L2.  lstm_res_fwd returns 3 when the grid
// cannot be co-resident on this device (the caller then takes the step-per-launch kernels).
bool lstm_res_supported(int dtype, int rows, int T, int H, int D);
void lstm_res_set_stamps(unsigned long long* buf);
int lstm_res_set_enabled(int on);   // on < 0: query only; returns the previous setting
int lstm_res_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, cudaStream_t st);

// Strict-fp32 mode (ops_simt.cu): CUDA-core fp32 implementations of every contraction, same index conventions as the
// tensor-core launchers
int simt_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* out, float* out_f32, long ldo, int M,
                    int N, int K, int relu, cudaStream_t st);
int simt_linear_dgrad(const float* dy, long lddy, const float* w, float* dx, float* dx_f32, const float* relu_mask, long ldx,
                      int M, int N, int K, cudaStream_t st);
int simt_linear_wgrad(const float* dy, long lddy, const float* x, long ldx, float* dw, long lddw, int M, int N, int K,
                      float alpha, cudaStream_t st);
int simt_conv5(const float* x, const float* wk, const float* bias, float* y, float* y_f32, int R, int T, int Cin, int Cout,
               bool dgrad, cudaStream_t st);
int simt_conv5_wgrad(const float* dy, const float* x, float* dwk, int R, int T, int Cin, int Cout, float alpha, cudaStream_t st);
int simt_lstm_fwd(float* xg, const float* whh_p, float* h_all, float* c_all, int rows, int T, int H, int D, cudaStream_t st);
int simt_lstm_bwd(const float* dh_all, const float* gates, const float* c_all, const float* whh_n, float* da_all, float* dc_ws,
                  int rows, int T, int H, int D, cudaStream_t st);
int simt_lstm_wgrad_hh(const float* da_all, const float* h_all, float* dwhh, int rows, int T, int H, int D, float alpha,
                       cudaStream_t st);

}  // namespace dvae
