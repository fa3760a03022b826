// Strict-fp32 mode (dtype tag kF32): every contraction of the hot path on the CUDA cores, fp32 operands, fp32 FMA
// accumulation, nothing rounded to a narrower grid anywhere.  This is the mode north_star's 1e-5 tolerance is stated
// for ("fp32 mode"); SURVEY.md section 7 step 4 names "a CUDA-core path" as the mechanism.  It exists for checking, not
// for speed: one tiled kernel (64 x 64 output tile, 4 x 4 outputs per thread, operands staged through shared memory)
// serves Linear / Conv1d(k=5) / LSTM forward, dgrad and wgrad through small accessor functors -- the same index
// conventions as the tensor-core launchers in ops_gemm.cu, so the engine above does not know the difference.
//
// Reference call sites (model/disentangled_vae.py): Conv1d :154-160,:178-189,:54-78; LSTM :163,:172,:193;
// Linear :165-171,:194.
#include <cuda_runtime.h>

#include "act_types.cuh"
#include "host_common.h"

namespace dvae {

constexpr int kSimtTile = 64;   // output tile (rows and columns)
constexpr int kSimtK = 16;      // reduction slab

// C(m, n) = sum_k A(m, k) * B(n, k); A / B are functors returning 0 outside their domain, E consumes the results.
template <class AF, class BF, class EF>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const AF a, const BF b, const EF e, const int M, const int N, const long K) {
  __shared__ float sa[kSimtK][kSimtTile + 1];
  __shared__ float sb[kSimtK][kSimtTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kSimtTile, n0 = blockIdx.x * kSimtTile;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long k0 = 0; k0 < K; k0 += kSimtK) {
    // 64 x 16 elements per operand, 256 threads: 4 each; consecutive threads walk k (contiguous for K-major operands)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int kk = idx & (kSimtK - 1), r = idx >> 4;
      const long k = k0 + kk;
      sa[kk][r] = (m0 + r < M && k < K) ? a(m0 + r, k) : 0.f;
      sb[kk][r] = (n0 + r < N && k < K) ? b(n0 + r, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSimtK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) e(m, n, acc[i][j]);
    }
}

template <class AF, class BF, class EF>
static int simt_gemm(const AF& a, const BF& b, const EF& e, int M, int N, long K, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid(ceil_div(N, kSimtTile), ceil_div(M, kSimtTile));
  simt_gemm_kernel<AF, BF, EF><<<grid, 256, 0, st>>>(a, b, e, M, N, K);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- accessors
struct RowMajor {          // element (i, k) of a matrix with row stride ld: p[i * ld + k]
  const float* p; long ld;
  __device__ __forceinline__ float operator()(int i, long k) const { return p[static_cast<long>(i) * ld + k]; }
};
struct ColMajor {          // element (i, k) = p[k * ld + i]: the transposed view
  const float* p; long ld;
  __device__ __forceinline__ float operator()(int i, long k) const { return p[k * ld + i]; }
};
// channels-last activation [R, T, C] seen by Conv1d(k=5, pad=2): row m = (r, t), reduction index k = (tap, c);
// frames outside the sequence are zero (taps never bleed across sequences)
struct ConvRows {
  const float* p; int T, C;
  __device__ __forceinline__ float operator()(int m, long k) const {
    const int tap = static_cast<int>(k / C), c = static_cast<int>(k - static_cast<long>(tap) * C);
    const int r = m / T, t = m - r * T + tap - 2;
    return (t >= 0 && t < T) ? p[(static_cast<long>(r) * T + t) * C + c] : 0.f;
  }
};
// dgrad weights: B(ci, k = (tap', co)) = wk[co][4 - tap'][ci]
struct ConvDgradW {
  const float* wk; int Cin, Cout;
  __device__ __forceinline__ float operator()(int ci, long k) const {
    const int tap = static_cast<int>(k / Cout), co = static_cast<int>(k - static_cast<long>(tap) * Cout);
    return wk[(static_cast<long>(co) * 5 + (4 - tap)) * Cin + ci];
  }
};
// wgrad: A(co, m) = dy[m][co];  B(n = (tap, ci), m = (r, t)) = x[r][t + tap - 2][ci]
struct ConvWgradX {
  const float* x; int T, C;
  __device__ __forceinline__ float operator()(int n, long m) const {
    const int tap = n / C, c = n - tap * C;
    const long r = m / T;
    const int t = static_cast<int>(m - r * T) + tap - 2;
    return (t >= 0 && t < T) ? x[(r * T + t) * C + c] : 0.f;
  }
};
// LSTM: one time step of a [rows, T, ld] tensor: element (r, k) = p[(r * T + t) * ld + off + k]; t out of range -> 0
struct StepRows {
  const float* p; int T, t; long ld, off;
  __device__ __forceinline__ float operator()(int r, long k) const {
    return (t >= 0 && t < T) ? p[(static_cast<long>(r) * T + t) * ld + off + k] : 0.f;
  }
};
// dW_hh: A(n, m = (r, t)) = da[r][t][off_a + n];  B(k, m) = h[r][t + shift][off_h + k] (zero outside the sequence)
struct SeqCols {
  const float* p; int T, shift; long ld, off;
  __device__ __forceinline__ float operator()(int i, long m) const {
    const long r = m / T;
    const int t = static_cast<int>(m - r * T) + shift;
    return (t >= 0 && t < T) ? p[(r * T + t) * ld + off + i] : 0.f;
  }
};

// ---- epilogues
struct EpiOut {            // out = act(v + bias) (+ fp32 mirror), optional ReLU-mask (dgrad of a ReLU layer)
  float* out; float* out_f32; const float* bias; const float* mask; long ldo; int relu;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    const long i = static_cast<long>(m) * ldo + n;
    if (mask && !(mask[i] > 0.f)) v = 0.f;
    if (out) out[i] = v;
    if (out_f32) out_f32[i] = v;
  }
};
struct EpiAccum {          // out += alpha * v (weight gradients: one thread owns each element, no atomics needed)
  float* out; long ldo; float alpha;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { out[static_cast<long>(m) * ldo + n] += alpha * v; }
};
struct EpiStepAdd {        // xg[r][t][off + n] += v  (recurrent part added to the x-projection in place)
  float* p; int T, t; long ld, off;
  __device__ __forceinline__ void operator()(int r, int n, float v) const { p[(static_cast<long>(r) * T + t) * ld + off + n] += v; }
};
struct EpiPlain {
  float* out; long ldo;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { out[static_cast<long>(m) * ldo + n] = v; }
};

// ---- LSTM cells (accurate expf / tanhf: this mode is the 1e-5 check)
__device__ __forceinline__ float sig_acc(float x) { return 1.f / (1.f + expf(-x)); }
// one time step, one direction: pre-activations (gate-interleaved: column 4u + g) -> activated gates in place, c, h
__global__ void simt_cell_fwd_kernel(float* __restrict__ xg, const float* __restrict__ c_prev, float* __restrict__ c_out,
                                     float* __restrict__ h_out, int rows, int H, long ldx, long ldh) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(rows) * H) return;
  const long r = i / H;
  const int u = static_cast<int>(i - r * H);
  float* g = xg + r * ldx + 4 * u;
  const float ig = sig_acc(g[0]), fg = sig_acc(g[1]), gg = tanhf(g[2]), og = sig_acc(g[3]);
  const float c = fg * (c_prev ? c_prev[r * ldh + u] : 0.f) + ig * gg;
  g[0] = ig; g[1] = fg; g[2] = gg; g[3] = og;
  c_out[r * ldh + u] = c;
  h_out[r * ldh + u] = og * tanhf(c);
}
// backward of one step: dh = dh_out + dh_rec; emits da (natural gate order: column g * H + u) and the new dc carry
__global__ void simt_cell_bwd_kernel(const float* __restrict__ dh_out, const float* __restrict__ dh_rec,
                                     const float* __restrict__ gates, const float* __restrict__ c_t,
                                     const float* __restrict__ c_prev, float* __restrict__ dc, float* __restrict__ da, int rows,
                                     int H, long ldx, long ldh, int dc_zero) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(rows) * H) return;
  const long r = i / H;
  const int u = static_cast<int>(i - r * H);
  const float* g = gates + r * ldx + 4 * u;
  const float ig = g[0], fg = g[1], gg = g[2], og = g[3];
  const float tc = tanhf(c_t[r * ldh + u]);
  const float dht = dh_out[r * ldh + u] + (dh_rec ? dh_rec[i] : 0.f);
  const float dct = (dc_zero ? 0.f : dc[i]) + dht * og * (1.f - tc * tc);
  float* o = da + r * ldx;
  o[0 * H + u] = dct * gg * ig * (1.f - ig);
  o[1 * H + u] = dct * (c_prev ? c_prev[r * ldh + u] : 0.f) * fg * (1.f - fg);
  o[2 * H + u] = dct * ig * (1.f - gg * gg);
  o[3 * H + u] = dht * tc * og * (1.f - og);
  dc[i] = dct * fg;
}

// =====================================================================================================================
int simt_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* out, float* out_f32, long ldo, int M,
                    int N, int K, int relu, cudaStream_t st) {
  return simt_gemm(RowMajor{x, ldx}, RowMajor{w, K}, EpiOut{out, out_f32, bias, nullptr, ldo, relu}, M, N, K, st);
}
int simt_linear_dgrad(const float* dy, long lddy, const float* w, float* dx, float* dx_f32, const float* relu_mask, long ldx,
                      int M, int N, int K, cudaStream_t st) {
  // dx[M,K] = dy[M,N] . w[N,K]
  return simt_gemm(RowMajor{dy, lddy}, ColMajor{w, K}, EpiOut{dx, dx_f32, nullptr, relu_mask, ldx, 0}, M, K, N, st);
}
int simt_linear_wgrad(const float* dy, long lddy, const float* x, long ldx, float* dw, long lddw, int M, int N, int K,
                      float alpha, cudaStream_t st) {
  // dw[N,K] += alpha * dy[M,N]^T . x[M,K]
  return simt_gemm(ColMajor{dy, lddy}, ColMajor{x, ldx}, EpiAccum{dw, lddw, alpha}, N, K, M, st);
}
int simt_conv5(const float* x, const float* wk, const float* bias, float* y, float* y_f32, int R, int T, int Cin, int Cout,
               bool dgrad, cudaStream_t st) {
  const int M = R * T;
  if (!dgrad)
    return simt_gemm(ConvRows{x, T, Cin}, RowMajor{wk, 5L * Cin}, EpiOut{y, y_f32, bias, nullptr, Cout, 0}, M, Cout, 5L * Cin, st);
  // dgrad: the caller passes x := dy [R,T,Cout], y := dx [R,T,Cin]
  return simt_gemm(ConvRows{x, T, Cout}, ConvDgradW{wk, Cin, Cout}, EpiOut{y, y_f32, nullptr, nullptr, Cin, 0}, M, Cin, 5L * Cout, st);
}
int simt_conv5_wgrad(const float* dy, const float* x, float* dwk, int R, int T, int Cin, int Cout, float alpha, cudaStream_t st) {
  return simt_gemm(ColMajor{dy, Cout}, ConvWgradX{x, T, Cin}, EpiAccum{dwk, 5L * Cin, alpha}, Cout, 5 * Cin, static_cast<long>(R) * T, st);
}

int simt_lstm_fwd(float* xg, const float* whh_p, float* h_all, float* c_all, int rows, int T, int H, int D, cudaStream_t st) {
  const long ldx = static_cast<long>(D) * 4 * H, ldh = static_cast<long>(D) * H;
  const long n_cell = static_cast<long>(rows) * H;
  for (int d = 0; d < D; ++d) {
    for (int s = 0; s < T; ++s) {
      const int t = d == 0 ? s : T - 1 - s, tp = d == 0 ? t - 1 : t + 1;
      if (s > 0) {
        if (int e = simt_gemm(StepRows{h_all, T, tp, ldh, static_cast<long>(d) * H}, RowMajor{whh_p + static_cast<long>(d) * 4 * H * H, H},
                              EpiStepAdd{xg, T, t, ldx, static_cast<long>(d) * 4 * H}, rows, 4 * H, H, st))
          return e;
      }
      simt_cell_fwd_kernel<<<ceil_div(n_cell, 256), 256, 0, st>>>(
          xg + static_cast<long>(t) * ldx + d * 4 * H, s > 0 ? c_all + static_cast<long>(tp) * ldh + d * H : nullptr,
          c_all + static_cast<long>(t) * ldh + d * H, h_all + static_cast<long>(t) * ldh + d * H, rows, H, T * ldx, T * ldh);
    }
  }
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// dc_ws: fp32 [2][D][rows][H] scratch: dc carry, then dh_rec
int simt_lstm_bwd(const float* dh_all, const float* gates, const float* c_all, const float* whh_n, float* da_all, float* dc_ws,
                  int rows, int T, int H, int D, cudaStream_t st) {
  const long ldx = static_cast<long>(D) * 4 * H, ldh = static_cast<long>(D) * H;
  const long n_cell = static_cast<long>(rows) * H;
  for (int d = 0; d < D; ++d) {
    float* dc = dc_ws + static_cast<long>(d) * n_cell;
    float* dh_rec = dc_ws + (static_cast<long>(D) + d) * n_cell;
    for (int s = 0; s < T; ++s) {
      const int t = d == 0 ? T - 1 - s : s;            // processing order is the reverse of the forward's
      const int tn = d == 0 ? t + 1 : t - 1;           // the step processed just before (its da feeds dh_rec)
      const int tp = d == 0 ? t - 1 : t + 1;           // the forward's previous step (c_prev)
      if (s > 0) {
        // dh_rec[r][u] = sum_n da[r][tn][d*4H + n] * whh_n[d][n][u]
        if (int e = simt_gemm(StepRows{da_all, T, tn, ldx, static_cast<long>(d) * 4 * H}, ColMajor{whh_n + static_cast<long>(d) * 4 * H * H, H},
                              EpiPlain{dh_rec, H}, rows, H, 4 * H, st))
          return e;
      }
      const bool has_prev = tp >= 0 && tp < T;
      simt_cell_bwd_kernel<<<ceil_div(n_cell, 256), 256, 0, st>>>(
          dh_all + static_cast<long>(t) * ldh + d * H, s > 0 ? dh_rec : nullptr, gates + static_cast<long>(t) * ldx + d * 4 * H,
          c_all + static_cast<long>(t) * ldh + d * H, has_prev ? c_all + static_cast<long>(tp) * ldh + d * H : nullptr, dc,
          da_all + static_cast<long>(t) * ldx + d * 4 * H, rows, H, T * ldx, T * ldh, s == 0 ? 1 : 0);
    }
  }
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// dW_hh[d][4H][H] += alpha * sum_{r,t} da[r,t,d,:]^T h_prev[r,t,d,:], h_prev = h[t-1] (forward) / h[t+1] (reverse direction)
int simt_lstm_wgrad_hh(const float* da_all, const float* h_all, float* dwhh, int rows, int T, int H, int D, float alpha,
                       cudaStream_t st) {
  const long ldx = static_cast<long>(D) * 4 * H, ldh = static_cast<long>(D) * H;
  for (int d = 0; d < D; ++d) {
    if (int e = simt_gemm(SeqCols{da_all, T, 0, ldx, static_cast<long>(d) * 4 * H}, SeqCols{h_all, T, d == 0 ? -1 : 1, ldh, static_cast<long>(d) * H},
                          EpiAccum{dwhh + static_cast<long>(d) * 4 * H * H, H, alpha}, 4 * H, H, static_cast<long>(rows) * T, st))
      return e;
  }
  return 0;
}

}  // namespace dvae
