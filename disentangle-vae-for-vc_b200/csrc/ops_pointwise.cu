// Memory-bound kernels of the Disentangled-VAE hot path (everything that is not a tensor-core contraction):
// weight re-layouts, NCL <-> channels-last packing, train/eval BatchNorm1d (stats, apply+activation, backward),
// column sums (bias gradients).  All are coalesced, 16/32-byte vectorised, fp32 (or fp64) accumulating.
//
// Reference semantics: torch.nn.BatchNorm1d in train mode (model/disentangled_vae.py:159,:182,:58) = biased batch
// variance for normalisation, unbiased for running_var, momentum 0.1, eps 1e-5; statistics are kept PER CALL
// (x1-call, x2-call = "halves") although both halves live in one tensor here.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "act_types.cuh"
#include "host_common.h"

namespace dvae {

using bf16 = __nv_bfloat16;
constexpr int kAct_None = 0, kAct_Relu = 1, kAct_Tanh = 2;

static inline int grid_for(long n, int block, int max_blocks = 148 * 16) {
  long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  return g < 1 ? 1 : static_cast<int>(g);
}

// ------------------------------------------------------------------------------------ weight preparation
template <typename AT>
__global__ void cast_kernel(const float* __restrict__ src, AT* __restrict__ dst, long n, float scale) {
  const long stride = static_cast<long>(gridDim.x) * blockDim.x * 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  for (long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (aligned && i + 8 <= n) {
      float v[8];
      Act8<float>::load(src + i, v);
      if (scale != 1.f) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= scale;
      }
      Act8<AT>::store(dst + i, v);
    } else {
      for (long j = i; j < n && j < i + 8; ++j) dst[j] = from_f32<AT>(src[j] * scale);
    }
  }
}

// w[Co][Ci][5] (torch Conv1d) -> wk[Co][5][Ci]: the reduction dim (ci) becomes contiguous per tap.
// wk_cat (may be null): [Co][5][3Ci] = [w_hi | w_hi | w_lo], the weight side of the split-precision first convolution
template <typename AT>
__global__ void conv_weight_kernel(const float* __restrict__ w, AT* __restrict__ wk, AT* __restrict__ wk_cat, int Co, int Ci) {
  const long n = static_cast<long>(Co) * 5 * Ci;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ci = i % Ci;
    const int k = (i / Ci) % 5;
    const int co = i / (5L * Ci);
    const float v = w[(static_cast<long>(co) * Ci + ci) * 5 + k];
    const AT hi = from_f32<AT>(v);
    wk[i] = hi;
    if (wk_cat != nullptr) {
      AT* row = wk_cat + (static_cast<long>(co) * 5 + k) * 3 * Ci;
      row[ci] = hi;
      row[Ci + ci] = hi;
      row[2 * Ci + ci] = from_f32<AT>(v - to_f32(hi));
    }
  }
}
// split-precision copy of a linear weight (row-wise), w_lo = round(w - w_hi): parts = 2: dst [N][2K] = [w_hi | w_lo];
// parts = 3: dst [N][3K] = [w_hi | w_hi | w_lo] (against an operand laid out [hi | lo | hi])
template <typename AT>
__global__ void cast_split_kernel(const float* __restrict__ src, AT* __restrict__ dst, long rows, long K, int parts) {
  const long n = rows * K;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / K, c = i - r * K;
    const float v = src[i];
    const AT hi = from_f32<AT>(v);
    AT* row = dst + r * parts * K;
    row[c] = hi;
    if (parts == 3) row[K + c] = hi;
    row[(parts - 1) * K + c] = from_f32<AT>(v - to_f32(hi));
  }
}
// dwk[Co][5][Ci] fp32 -> dw[Co][Ci][5] fp32 (gradient back in the parameter's layout)
__global__ void conv_wgrad_unpack_kernel(const float* __restrict__ dwk, float* __restrict__ dw, int Co, int Ci) {
  const long n = static_cast<long>(Co) * 5 * Ci;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = i % 5;
    const int ci = (i / 5) % Ci;
    const int co = i / (5L * Ci);
    dw[i] = dwk[(static_cast<long>(co) * 5 + k) * Ci + ci];
  }
}

__device__ __forceinline__ int gate_perm_src(int n, int H, int /*tile*/) {
  // destination row n = 4*u + g of the gate-interleaved layout <- source row g*H + u of the torch [i;f;g;o] layout
  return (n & 3) * H + (n >> 2);
}
template <typename AT>
__global__ void lstm_weight_kernel(const float* __restrict__ w, AT* __restrict__ dst, int H, int In, int tile) {
  const long n = 4L * H * In;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int col = i % In;
    const int row = i / In;
    dst[i] = from_f32<AT>(w[static_cast<long>(gate_perm_src(row, H, tile)) * In + col]);
  }
}
__global__ void lstm_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ dst,
                                 int H, int tile) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < 4 * H) {
    const int s = gate_perm_src(n, H, tile);
    dst[n] = b_ih[s] + b_hh[s];
  }
}

// a += b (both activation dtype); joins the two gradient branches that meet at the decoder output
template <typename AT>
__global__ void add_inplace_kernel(AT* __restrict__ a, const AT* __restrict__ b, long n) {
  const long stride = static_cast<long>(gridDim.x) * blockDim.x * 8;
  for (long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float u[8], v[8];
      Act8<AT>::load(a + i, u);
      Act8<AT>::load(b + i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] += v[k];
      Act8<AT>::store(a + i, u);
    } else {
      for (long j = i; j < n; ++j) a[j] = from_f32<AT>(to_f32(a[j]) + to_f32(b[j]));
    }
  }
}

// out (fp32) = a (fp32) + b (activation dtype): residual output that stays channels-last (AutoVC replica)
template <typename AT>
__global__ void add_f32_act_kernel(const float* __restrict__ a, const AT* __restrict__ b, float* __restrict__ out, long n) {
  const long stride = static_cast<long>(gridDim.x) * blockDim.x * 8;
  for (long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float u[8], v[8];
      Act8<float>::load(a + i, u);
      Act8<AT>::load(b + i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] += v[k];
      Act8<float>::store(out + i, u);
    } else {
      for (long j = i; j < n; ++j) out[j] = a[j] + to_f32(b[j]);
    }
  }
}

// ------------------------------------------------------------------------------------ all weight copies in one launch
// After every optimizer step every tensor-core copy of the parameters has to be re-derived (dvae_b200.engine.
// PreparedWeights.refresh).  Tensor by tensor that is ~55 small launches; this kernel does it in one, driven by a table:
// one descriptor per fp32 source tensor (what to derive from it and where to), one (descriptor, offset) pair per block.
struct PrepDesc {
  const float* src;
  const float* src2;   // kPrepLstmBias: the second bias vector
  void* dst0;
  void* dst1;          // optional second destination (see the kinds)
  long n;              // elements of src
  int kind, d0, d1, pad;
};
enum PrepKind : int {
  kPrepCast = 0,       // dst0[i] = act(src[i]);  dst1 (optional, src viewed [n / d1, d1]): rows [hi | lo] (d0 = 2) or [hi | hi | lo] (d0 = 3)
  kPrepConv = 1,       // src [d0 = Co][d1 = Ci][5] -> dst0 [Co][5][Ci];  dst1 (optional) [Co][5][3Ci] = [hi | hi | lo]
  kPrepLstmW = 2,      // src [4H][d1 = In] (d0 = H): dst0 natural cast, dst1 rows gate-interleaved (row 4u + g <- row g*H + u)
  kPrepCopy = 3,       // dst0 (fp32)[i] = src[i]
  kPrepLstmBias = 4,   // dst0 (fp32)[4u + g] = src[g*H + u] + src2[g*H + u]   (d0 = H)
};
template <typename AT>
__global__ void __launch_bounds__(256) prep_all_kernel(const PrepDesc* __restrict__ descs, const int* __restrict__ blk_desc,
                                                       const long* __restrict__ blk_off, int chunk) {
  const PrepDesc d = descs[blk_desc[blockIdx.x]];
  const long off = blk_off[blockIdx.x];
  const long end = min(d.n, off + chunk);
  for (long i = off + threadIdx.x; i < end; i += blockDim.x) {
    const float v = d.src[i];
    if (d.kind == kPrepCopy) {
      static_cast<float*>(d.dst0)[i] = v;
      continue;
    }
    if (d.kind == kPrepLstmBias) {
      const int H = d.d0, g = static_cast<int>(i / H), u = static_cast<int>(i - static_cast<long>(g) * H);
      static_cast<float*>(d.dst0)[4 * u + g] = v + d.src2[i];
      continue;
    }
    const AT hi = from_f32<AT>(v);
    if (d.kind == kPrepCast) {
      static_cast<AT*>(d.dst0)[i] = hi;
      if (d.dst1 != nullptr) {
        const long K = d.d1, r = i / K, c = i - r * K;
        AT* row = static_cast<AT*>(d.dst1) + r * d.d0 * K;
        row[c] = hi;
        if (d.d0 == 3) row[K + c] = hi;
        row[(d.d0 - 1) * K + c] = from_f32<AT>(v - to_f32(hi));
      }
    } else if (d.kind == kPrepConv) {
      const int Ci = d.d1;
      const int k = static_cast<int>(i % 5);
      const long rest = i / 5;
      const int ci = static_cast<int>(rest % Ci);
      const long co = rest / Ci;
      static_cast<AT*>(d.dst0)[(co * 5 + k) * Ci + ci] = hi;
      if (d.dst1 != nullptr) {
        AT* row = static_cast<AT*>(d.dst1) + (co * 5 + k) * 3 * Ci;
        row[ci] = hi;
        row[Ci + ci] = hi;
        row[2 * Ci + ci] = from_f32<AT>(v - to_f32(hi));
      }
    } else {   // kPrepLstmW
      const int H = d.d0;
      const long In = d.d1, r = i / In, c = i - r * In;
      const int g = static_cast<int>(r / H), u = static_cast<int>(r - static_cast<long>(g) * H);
      static_cast<AT*>(d.dst0)[i] = hi;
      static_cast<AT*>(d.dst1)[(4L * u + g) * In + c] = hi;
    }
  }
}

// ------------------------------------------------------------------------------------ layout packing
// x fp32 [R][C][T] (reference NCL) -> y act [R][T][C] (channels-last, the GEMM A-operand layout)
// split != 0: besides y (the rounded values, "hi") also y_cat [R][T][3C] = [hi | lo | hi] with lo = round(x - hi): the
// K-concatenated operand that, against weights laid out [w_hi | w_hi | w_lo], makes the first convolution see its fp32
// input and weights to ~2^-22 instead of 2^-11 (three products of a split-precision multiplication in one GEMM)
template <typename AT>
__global__ void ncl_to_cl_kernel(const float* __restrict__ x, AT* __restrict__ y, AT* __restrict__ y_cat, int C, int T) {
  extern __shared__ float tile[];  // [C][T+1]
  const long r = blockIdx.x;
  const float* xr = x + r * C * T;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    tile[c * (T + 1) + t] = xr[i];
  }
  __syncthreads();
  AT* yr = y + r * C * T;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    const float v = tile[c * (T + 1) + t];
    const AT hi = from_f32<AT>(v);
    yr[i] = hi;
    if (y_cat != nullptr) {
      AT* row = y_cat + (r * T + t) * 3L * C;
      row[c] = hi;
      row[C + c] = from_f32<AT>(v - to_f32(hi));
      row[2 * C + c] = hi;
    }
  }
}
// channels-last -> NCL fp32 with optional residual:  out[r][c][t] = a[r][t][c] (+ b[r][t][c]);  a is fp32 or act
template <typename TA, typename TB>
__global__ void cl_to_ncl_kernel(const TA* __restrict__ a, const TB* __restrict__ b, float* __restrict__ out_a,
                                 float* __restrict__ out_sum, int C, int T) {
  extern __shared__ float tile[];  // two [T][C+1] planes
  float* ta = tile;
  float* tb = tile + T * (C + 1);
  const long r = blockIdx.x;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    const float va = to_f32(a[r * C * T + i]);
    ta[t * (C + 1) + c] = va;
    if (b != nullptr) tb[t * (C + 1) + c] = va + to_f32(b[r * C * T + i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    if (out_a != nullptr) out_a[r * C * T + i] = ta[t * (C + 1) + c];
    if (out_sum != nullptr) out_sum[r * C * T + i] = tb[t * (C + 1) + c];
  }
}
// ---- conversion front / back end (model/variational_base_vae.py:335-348 chunking_mel, :295-296 time-concat + clamp), for
// many utterances of different lengths in one launch.  Utterance u is a [C][T_u] fp32 matrix at mel + mel_off[u]; it owns
// the chunks [chunk_first[u], chunk_first[u+1]) (T_u / 64 + 1 of them: the last one zero padded, a whole zero chunk when
// T_u % 64 == 0).  One block per chunk.
template <typename AT>
__global__ void chunk_mel_kernel(const float* __restrict__ mel, const long* __restrict__ mel_off, const int* __restrict__ t_len,
                                 const int* __restrict__ chunk_first, const int* __restrict__ chunk_utt, AT* __restrict__ x_cl,
                                 int C, int T) {
  extern __shared__ float tile[];  // [C][T+1]
  const long k = blockIdx.x;
  const int u = chunk_utt[k];
  const int t0 = (static_cast<int>(k) - chunk_first[u]) * T;
  const int Tu = t_len[u];
  const float* m = mel + mel_off[u];
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    tile[c * (T + 1) + t] = (t0 + t < Tu) ? m[static_cast<long>(c) * Tu + t0 + t] : 0.f;
  }
  __syncthreads();
  AT* y = x_cl + k * C * T;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    y[i] = from_f32<AT>(tile[c * (T + 1) + t]);
  }
}
// chunk k of utterance u, channels-last a (fp32) [+ b (act)] -> out_u [C][n_u * T] at out + out_off[u], frames
// [(k - chunk_first[u]) * T, +T); optionally clamped to [lo, hi]
template <typename AT>
__global__ void unchunk_mel_kernel(const float* __restrict__ a, const AT* __restrict__ b, const long* __restrict__ out_off,
                                   const int* __restrict__ chunk_first, const int* __restrict__ chunk_utt, float* __restrict__ out,
                                   int C, int T, int clamp, float lo, float hi) {
  extern __shared__ float tile[];  // [T][C+1]
  const long k = blockIdx.x;
  const int u = chunk_utt[k];
  const int kk = static_cast<int>(k) - chunk_first[u];
  const int n_u = chunk_first[u + 1] - chunk_first[u];
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    float v = a[k * C * T + i];
    if (b != nullptr) v += to_f32(b[k * C * T + i]);
    if (clamp) v = fminf(fmaxf(v, lo), hi);
    tile[t * (C + 1) + c] = v;
  }
  __syncthreads();
  float* o = out + out_off[u];
  const long ld = static_cast<long>(n_u) * T;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    o[c * ld + static_cast<long>(kk) * T + t] = tile[t * (C + 1) + c];
  }
}
// backward of the residual output: d_rec[r][t][c] = g_rec[r][c][t] + g_hat[r][c][t];  d_post = g_hat  (either may be null)
template <typename AT>
__global__ void recon_out_bwd_kernel(const float* __restrict__ g_rec, const float* __restrict__ g_hat, AT* __restrict__ d_rec,
                                     AT* __restrict__ d_post, int C, int T, float scale) {
  extern __shared__ float tile[];
  float* ta = tile;
  float* tb = tile + C * (T + 1);
  const long r = blockIdx.x;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    const float gh = g_hat ? scale * g_hat[r * C * T + i] : 0.f;
    const float gr = g_rec ? scale * g_rec[r * C * T + i] : 0.f;
    ta[c * (T + 1) + t] = gr + gh;
    tb[c * (T + 1) + t] = gh;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    d_rec[r * C * T + i] = from_f32<AT>(ta[c * (T + 1) + t]);
    d_post[r * C * T + i] = from_f32<AT>(tb[c * (T + 1) + t]);
  }
}

// ------------------------------------------------------------------------------------ BatchNorm1d
// All four streaming kernels share one mapping: a block owns a slab of 64 channels (8 threads x 8 channels = one 128-byte
// line per row in bf16) for `rb` consecutive rows of one statistics half; its 32 row lanes walk down the rows four at a
// time, so every thread keeps 4 (or 8, with two input tensors) independent 16-byte loads in flight -- the kernels are
// pure HBM streams and latency x bandwidth decides how many bytes have to be outstanding.  The reductions finish with
// 2 x 64 fp64 atomics per block (a 512-row block: 8x fewer than one atomic per channel per 64-row block).
// stats buffer (double) [halves][2][C]: sum, sum of squares.
constexpr int kBnRowsPerBlock = 64;    // granularity the callers guarantee for rows_half
constexpr int kBnSlab = 64;            // channels per block
constexpr int kBnLanes = 32;           // row lanes per block (256 threads)
constexpr int kBnUnroll = 4;

// block-level reduction of per-thread partial sums s[8], q[8] over the 32 row lanes, then fp64 atomics
__device__ __forceinline__ void bn_block_reduce(const float* s, const float* q, float* red, double* sums_half, int c_slab,
                                                int C) {
  const int rl = threadIdx.x >> 3, cg = threadIdx.x & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[(rl * 2 + 0) * kBnSlab + cg * 8 + i] = s[i];
    red[(rl * 2 + 1) * kBnSlab + cg * 8 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < 2 * kBnSlab) {
    const int which = threadIdx.x / kBnSlab, c = threadIdx.x - which * kBnSlab;
    float acc = 0.f;
#pragma unroll 8
    for (int l = 0; l < kBnLanes; ++l) acc += red[(l * 2 + which) * kBnSlab + c];
    if (c_slab + c < C) atomicAdd(sums_half + static_cast<long>(which) * C + c_slab + c, static_cast<double>(acc));
  }
}

template <typename AT>
__global__ void __launch_bounds__(256) bn_stats_kernel(const AT* __restrict__ y, double* __restrict__ sums, int rows_half, int C,
                                                       int rb) {
  __shared__ float red[kBnLanes * 2 * kBnSlab];
  const int rl = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const int c_slab = blockIdx.x * kBnSlab, c0 = c_slab + cg * 8;
  const long row0 = static_cast<long>(blockIdx.y) * rb;
  const int half = static_cast<int>(row0 / rows_half);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (c0 < C) {
    const AT* base = y + row0 * C + c0;
    for (int r = rl; r < rb; r += kBnLanes * kBnUnroll) {
      typename Act8<AT>::raw_t raw[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u)
        if (r + u * kBnLanes < rb) raw[u] = Act8<AT>::load_raw(base + static_cast<long>(r + u * kBnLanes) * C);
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        if (r + u * kBnLanes < rb) {
          float v[8];
          Act8<AT>::unpack(raw[u], v);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
        }
      }
    }
  }
  bn_block_reduce(s, q, red, sums + static_cast<long>(half) * 2 * C, c_slab, C);
}
// stat (fp32) [halves][4][C]: mean, rstd, scale = rstd*gamma, shift = beta - mean*scale.  Running statistics are
// updated half by half, in call order (x1 then x2), exactly like two consecutive nn.BatchNorm1d calls.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ stat, float* __restrict__ run_mean,
                                   float* __restrict__ run_var, long long* __restrict__ num_batches, int halves, int C,
                                   double n, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float rm = run_mean ? run_mean[c] : 0.f, rv = run_var ? run_var[c] : 1.f;
    for (int h = 0; h < halves; ++h) {
      const double mean = sums[(h * 2 + 0) * C + c] / n;
      double var = sums[(h * 2 + 1) * C + c] / n - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      const float rstd = rsqrtf(static_cast<float>(var) + eps);
      const float scale = rstd * gamma[c];
      stat[(h * 4 + 0) * C + c] = static_cast<float>(mean);
      stat[(h * 4 + 1) * C + c] = rstd;
      stat[(h * 4 + 2) * C + c] = scale;
      stat[(h * 4 + 3) * C + c] = beta[c] - static_cast<float>(mean) * scale;
      rm = (1.f - momentum) * rm + momentum * static_cast<float>(mean);
      rv = (1.f - momentum) * rv + momentum * static_cast<float>(var * (n / (n - 1.0)));
    }
    if (run_mean) run_mean[c] = rm;
    if (run_var) run_var[c] = rv;
  }
  if (c == 0 && num_batches) *num_batches += halves;
}
// eval mode: stat from the running statistics
__global__ void bn_eval_stat_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ run_mean, const float* __restrict__ run_var,
                                    float* __restrict__ stat, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float rstd = rsqrtf(run_var[c] + eps);
    const float scale = rstd * gamma[c];
    stat[0 * C + c] = run_mean[c];
    stat[1 * C + c] = rstd;
    stat[2 * C + c] = scale;
    stat[3 * C + c] = beta[c] - run_mean[c] * scale;
  }
}

// tanh of the postnet activations: the fast exp-based form for the tensor-core modes, libm's for strict fp32 (AT = float)
template <typename AT> __device__ __forceinline__ float act_tanh(float z) { return tanh_f(z); }
template <> __device__ __forceinline__ float act_tanh<float>(float z) { return tanhf(z); }
template <typename AT>
__device__ __forceinline__ float apply_act(float z, int act) {
  if (act == kAct_Relu) return fmaxf(z, 0.f);
  if (act == kAct_Tanh) return act_tanh<AT>(z);
  return z;
}
template <typename AT>
__device__ __forceinline__ float act_grad_from_z(float z, int act) {
  if (act == kAct_Relu) return z > 0.f ? 1.f : 0.f;
  if (act == kAct_Tanh) { const float t = act_tanh<AT>(z); return 1.f - t * t; }
  return 1.f;
}

// out = act(y * scale[h] + shift[h]); a thread keeps the affine constants of its 8 channels in registers.
// YT: storage type of y (the activation type, or float when the convolution output is kept unrounded)
template <typename AT, typename YT = AT>
__global__ void __launch_bounds__(256) bn_apply_kernel(const YT* __restrict__ y, AT* __restrict__ out,
                                                       const float* __restrict__ stat, long rows, int rows_half, int C, int act,
                                                       int rb) {
  const int rl = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const int c0 = blockIdx.x * kBnSlab + cg * 8;
  if (c0 >= C) return;
  const long row0 = static_cast<long>(blockIdx.y) * rb;
  const int h = static_cast<int>(row0 / rows_half);
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = stat[(h * 4 + 2) * C + c0 + k];
    sh[k] = stat[(h * 4 + 3) * C + c0 + k];
  }
  const int nr = static_cast<int>(min(static_cast<long>(rb), rows - row0));
  const YT* src = y + row0 * C + c0;
  AT* dst = out + row0 * C + c0;
  for (int r = rl; r < nr; r += kBnLanes * kBnUnroll) {
    typename Act8<YT>::raw_t raw[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u)
      if (r + u * kBnLanes < nr) raw[u] = Act8<YT>::load_raw(src + static_cast<long>(r + u * kBnLanes) * C);
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      if (r + u * kBnLanes < nr) {
        float v[8];
        Act8<YT>::unpack(raw[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = apply_act<AT>(fmaf(v[k], sc[k], sh[k]), act);
        Act8<AT>::store(dst + static_cast<long>(r + u * kBnLanes) * C, v);
      }
    }
  }
}

// backward pass 1: per (half, channel) sums of dz and dz*xhat, dz = dout * act'(z)
template <typename AT, typename YT = AT>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const AT* __restrict__ dout, const YT* __restrict__ y,
                                                            const float* __restrict__ stat, double* __restrict__ sums,
                                                            int rows_half, int C, int act, int rb) {
  __shared__ float red[kBnLanes * 2 * kBnSlab];
  constexpr int UNROLL = kBnUnroll;   // (halving it for a 32-bit y, as bn_bwd_apply_kernel does, measured slower here: 46.9 vs 41.9 us)
  const int rl = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const int c_slab = blockIdx.x * kBnSlab, c0 = c_slab + cg * 8;
  const long row0 = static_cast<long>(blockIdx.y) * rb;
  const int h = static_cast<int>(row0 / rows_half);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (c0 < C) {
    float mean[8], rstd[8], sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = stat[(h * 4 + 0) * C + c0 + i];
      rstd[i] = stat[(h * 4 + 1) * C + c0 + i];
      sc[i] = stat[(h * 4 + 2) * C + c0 + i];
      sh[i] = stat[(h * 4 + 3) * C + c0 + i];
    }
    const YT* ys = y + row0 * C + c0;
    const AT* dsrc = dout + row0 * C + c0;
    for (int r = rl; r < rb; r += kBnLanes * UNROLL) {
      typename Act8<YT>::raw_t ry[UNROLL];
      typename Act8<AT>::raw_t rd[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (r + u * kBnLanes < rb) {
          ry[u] = Act8<YT>::load_raw(ys + static_cast<long>(r + u * kBnLanes) * C);
          rd[u] = Act8<AT>::load_raw(dsrc + static_cast<long>(r + u * kBnLanes) * C);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (r + u * kBnLanes < rb) {
          float v[8], d[8];
          Act8<YT>::unpack(ry[u], v);
          Act8<AT>::unpack(rd[u], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dz = d[i] * act_grad_from_z<AT>(fmaf(v[i], sc[i], sh[i]), act);
            s[i] += dz;
            q[i] = fmaf(dz, (v[i] - mean[i]) * rstd[i], q[i]);
          }
        }
      }
    }
  }
  bn_block_reduce(s, q, red, sums + static_cast<long>(h) * 2 * C, c_slab, C);
}
// dgamma = sum_h sum dz*xhat, dbeta = sum_h sum dz; coef [halves][2][C] = (sum dz)/n, (sum dz*xhat)/n
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, float* __restrict__ coef, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int halves, int C, double n, double alpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dg = 0.0, db = 0.0;
  for (int h = 0; h < halves; ++h) {
    const double s = sums[(h * 2 + 0) * C + c], q = sums[(h * 2 + 1) * C + c];
    db += s; dg += q;
    coef[(h * 2 + 0) * C + c] = static_cast<float>(s / n);
    coef[(h * 2 + 1) * C + c] = static_cast<float>(q / n);
  }
  dgamma[c] = static_cast<float>(dg * alpha);
  dbeta[c] = static_cast<float>(db * alpha);
}
// backward pass 2: dy = scale * (dz - mean(dz) - xhat * mean(dz*xhat)); same streaming structure as bn_apply_kernel
template <typename AT, typename YT = AT>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const AT* __restrict__ dout, const YT* __restrict__ y,
                                                           const float* __restrict__ stat, const float* __restrict__ coef,
                                                           AT* __restrict__ dy, long rows, int rows_half, int C, int act,
                                                           int rb) {
  // a 32-bit y doubles the bytes (and registers) in flight per row: half the unroll keeps the occupancy
  constexpr int UNROLL = sizeof(YT) == 4 ? kBnUnroll / 2 : kBnUnroll;
  const int rl = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const int c0 = blockIdx.x * kBnSlab + cg * 8;
  if (c0 >= C) return;
  const long row0 = static_cast<long>(blockIdx.y) * rb;
  const int h = static_cast<int>(row0 / rows_half);
  float mean[8], rstd[8], sc[8], sh[8], k0[8], k1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = stat[(h * 4 + 0) * C + c0 + k];
    rstd[k] = stat[(h * 4 + 1) * C + c0 + k];
    sc[k] = stat[(h * 4 + 2) * C + c0 + k];
    sh[k] = stat[(h * 4 + 3) * C + c0 + k];
    k0[k] = coef[(h * 2 + 0) * C + c0 + k];
    k1[k] = coef[(h * 2 + 1) * C + c0 + k];
  }
  const int nr = static_cast<int>(min(static_cast<long>(rb), rows - row0));
  const YT* ys = y + row0 * C + c0;
  const AT* dsrc = dout + row0 * C + c0;
  AT* dst = dy + row0 * C + c0;
  for (int r = rl; r < nr; r += kBnLanes * UNROLL) {
    typename Act8<YT>::raw_t ry[UNROLL];
    typename Act8<AT>::raw_t rd[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (r + u * kBnLanes < nr) {
        ry[u] = Act8<YT>::load_raw(ys + static_cast<long>(r + u * kBnLanes) * C);
        rd[u] = Act8<AT>::load_raw(dsrc + static_cast<long>(r + u * kBnLanes) * C);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (r + u * kBnLanes < nr) {
        float v[8], d[8];
        Act8<YT>::unpack(ry[u], v);
        Act8<AT>::unpack(rd[u], d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dz = d[k] * act_grad_from_z<AT>(fmaf(v[k], sc[k], sh[k]), act);
          const float xh = (v[k] - mean[k]) * rstd[k];
          d[k] = sc[k] * (dz - k0[k] - xh * k1[k]);
        }
        Act8<AT>::store(dst + static_cast<long>(r + u * kBnLanes) * C, d);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ column sums (bias gradients)
template <typename AT>
__global__ void colsum_kernel(const AT* __restrict__ x, float* __restrict__ out, long rows, int C, long ldx, float alpha) {
  // block = 256 threads; thread owns 8 channels of one row lane; grid.x over row chunks, grid.y over channel chunks of 2048
  const int c0 = blockIdx.y * 2048;
  const int cw = min(2048, C - c0);
  const int tpr = cw >> 3;
  const int lanes = blockDim.x / tpr;
  const int rl = threadIdx.x / tpr, cg = threadIdx.x - rl * tpr;
  extern __shared__ float red[];  // [lanes][cw]
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  const long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (rl < lanes) {
    for (long r = r0 + rl; r < r1; r += lanes) {
      float v[8];
      Act8<AT>::load(x + r * ldx + c0 + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl * cw + cg * 8 + i] = s[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cw; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[l * cw + i];
    atomicAdd(out + c0 + i, acc * alpha);
  }
}

}  // namespace dvae

using namespace dvae;

#define DISPATCH_AT(dtype, ...)                              \
  do {                                                       \
    if ((dtype) == kBF16) { using AT = bf16; __VA_ARGS__; }  \
    else if ((dtype) == kF16) { using AT = __half; __VA_ARGS__; } \
    else if ((dtype) == kTF32) { using AT = tf32_t; __VA_ARGS__; } \
    else if ((dtype) == kF32) { using AT = float; __VA_ARGS__; } \
    else { set_last_error("unknown dtype tag"); return 1; }  \
  } while (0)

// y_f32: the BatchNorm input y is unrounded fp32 (fp16 mode) instead of the activation type
#define DISPATCH_AT_Y(dtype, y_f32, ...)                                        \
  do {                                                                          \
    if (y_f32) { using YT = float; DISPATCH_AT(dtype, __VA_ARGS__); }           \
    else { DISPATCH_AT(dtype, { using YT = AT; __VA_ARGS__; }); }               \
  } while (0)

// rows per block: the largest power-of-two multiple of 64 up to `want` that divides rows_half (blocks never straddle halves)
static int bn_rows_per_block(int rows_half, int want) {
  int rb = kBnRowsPerBlock;
  while (rb * 2 <= want && rows_half % (rb * 2) == 0) rb *= 2;
  return rb;
}

namespace dvae {
int bn_stats_launch(int dtype, const void* y, double* ws, int rows_half, int halves, int C, cudaStream_t st) {
  DVAE_REQUIRE(C % 8 == 0 && C <= 2048, "C must be a multiple of 8 (<= 2048)");
  DVAE_REQUIRE(rows_half % kBnRowsPerBlock == 0, "rows per half must be a multiple of 64");
  const long rows = static_cast<long>(rows_half) * halves;
  const int rb_red = bn_rows_per_block(rows_half, 512);
  const dim3 g_red(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_red));
  DISPATCH_AT(dtype, bn_stats_kernel<AT><<<g_red, 256, 0, st>>>((const AT*)y, ws, rows_half, C, rb_red));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace dvae

extern "C" {

int dvae_prep_cast(int dtype, const float* src, void* dst, long n, float scale, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_AT(dtype, cast_kernel<AT><<<grid_for((n + 7) / 8, 256), 256, 0, st>>>(src, (AT*)dst, n, scale));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// split-precision weights of the small linear layers: dst act [rows, parts*K] = [hi | lo] (parts = 2) or [hi | hi | lo]
// (parts = 3) of src fp32 [rows, K]
int dvae_prep_cast_split(int dtype, const float* src, void* dst, long rows, long K, int parts, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(parts == 2 || parts == 3, "parts must be 2 or 3");
  DISPATCH_AT(dtype, cast_split_kernel<AT><<<grid_for(rows * K, 256), 256, 0, st>>>(src, (AT*)dst, rows, K, parts));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_add_f32_act(int dtype, const float* a, const void* b, float* out, long n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_AT(dtype, add_f32_act_kernel<AT><<<grid_for((n + 7) / 8, 256), 256, 0, st>>>(a, (const AT*)b, out, n));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_copy_f32(const float* src, float* dst, long n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  cast_kernel<float><<<grid_for((n + 7) / 8, 256), 256, 0, st>>>(src, dst, n, 1.f);   // exact copy (no tf32 rounding)
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_add_inplace(int dtype, void* a, const void* b, long n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_AT(dtype, add_inplace_kernel<AT><<<grid_for((n + 7) / 8, 256), 256, 0, st>>>((AT*)a, (const AT*)b, n));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// Every tensor-core copy of the parameters in ONE launch.  descs: device array of 64-byte descriptors {src, src2, dst0, dst1,
// n (int64), kind, d0, d1, pad (int32)} -- see PrepKind in this file; blk_desc / blk_off: for each block, which descriptor and
// which element offset its chunk of `chunk` source elements starts at.
int dvae_prep_all(int dtype, const void* descs, const int* blk_desc, const long* blk_off, int num_blocks, int chunk, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  static_assert(sizeof(PrepDesc) == 56, "descriptor layout: dvae_b200/ops.py PrepTable packs 56-byte records");
  if (num_blocks <= 0) return 0;
  DISPATCH_AT(dtype, prep_all_kernel<AT><<<num_blocks, 256, 0, st>>>(static_cast<const PrepDesc*>(descs), blk_desc, blk_off, chunk));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_prep_conv_weight(int dtype, const float* w, void* wk, void* wk_cat, int Co, int Ci, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_AT(dtype, conv_weight_kernel<AT><<<grid_for(5L * Co * Ci, 256), 256, 0, st>>>(w, (AT*)wk, (AT*)wk_cat, Co, Ci));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_conv_wgrad_unpack(const float* dwk, float* dw, int Co, int Ci, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  conv_wgrad_unpack_kernel<<<grid_for(5L * Co * Ci, 256), 256, 0, st>>>(dwk, dw, Co, Ci);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_prep_lstm_weight(int dtype, const float* w, void* dst, int H, int In, int tile, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(tile % 4 == 0 && (4 * H) % tile == 0, "gate tile must divide 4H");
  DISPATCH_AT(dtype, lstm_weight_kernel<AT><<<grid_for(4L * H * In, 256), 256, 0, st>>>(w, (AT*)dst, H, In, tile));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_prep_lstm_bias(const float* b_ih, const float* b_hh, float* dst, int H, int tile, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  lstm_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, st>>>(b_ih, b_hh, dst, H, tile);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// y_cat (may be null): additionally the split-precision operand [R, T, 3C] = [hi | lo | hi] (see ncl_to_cl_kernel)
int dvae_pack_ncl_to_cl(int dtype, const float* x, void* y, void* y_cat, int R, int C, int T, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (R == 0) return 0;
  const int smem = C * (T + 1) * 4;
  DVAE_REQUIRE(smem <= 48 * 1024, "C*(T+1) tile must fit 48 KB of shared memory");
  DISPATCH_AT(dtype, ncl_to_cl_kernel<AT><<<R, 256, smem, st>>>(x, (AT*)y, (AT*)y_cat, C, T));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// out_a[r][c][t] = a[r][t][c] ; out_sum = a + b.  a_is_f32: a is fp32 (else act dtype); b is act dtype (may be null).
int dvae_unpack_cl_to_ncl(int dtype, const void* a, int a_is_f32, const void* b, float* out_a, float* out_sum, int R, int C,
                          int T, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (R == 0) return 0;
  const int smem = 2 * T * (C + 1) * 4;
  DVAE_REQUIRE(smem <= 48 * 1024, "2*T*(C+1) tile must fit 48 KB of shared memory");
  if (a_is_f32) {
    DISPATCH_AT(dtype, cl_to_ncl_kernel<float, AT><<<R, 256, smem, st>>>((const float*)a, (const AT*)b, out_a, out_sum, C, T));
  } else {
    DISPATCH_AT(dtype, cl_to_ncl_kernel<AT, AT><<<R, 256, smem, st>>>((const AT*)a, (const AT*)b, out_a, out_sum, C, T));
  }
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// chunking_mel for a batch of utterances (see chunk_mel_kernel): x_cl act [n_chunks, T, C]
int dvae_chunk_mel(int dtype, const float* mel, const long* mel_off, const int* t_len, const int* chunk_first,
                   const int* chunk_utt, void* x_cl, int n_chunks, int C, int T, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n_chunks == 0) return 0;
  const int smem = C * (T + 1) * 4;
  DVAE_REQUIRE(smem <= 48 * 1024, "C*(T+1) tile must fit 48 KB of shared memory");
  DISPATCH_AT(dtype, chunk_mel_kernel<AT><<<n_chunks, 256, smem, st>>>(mel, mel_off, t_len, chunk_first, chunk_utt, (AT*)x_cl, C, T));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// time-concatenation of the chunks of every utterance (+ residual b, + clamp): out_u fp32 [C, n_u * T] at out + out_off[u]
int dvae_unchunk_mel(int dtype, const float* a, const void* b, const long* out_off, const int* chunk_first, const int* chunk_utt,
                     float* out, int n_chunks, int C, int T, int clamp, float lo, float hi, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n_chunks == 0) return 0;
  const int smem = T * (C + 1) * 4;
  DVAE_REQUIRE(smem <= 48 * 1024, "T*(C+1) tile must fit 48 KB of shared memory");
  DISPATCH_AT(dtype, unchunk_mel_kernel<AT><<<n_chunks, 256, smem, st>>>(a, (const AT*)b, out_off, chunk_first, chunk_utt, out, C, T,
                                                                        clamp, lo, hi));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// scale: the factor the activation-gradient stream is carried at from here on (fp16 mode; 1 otherwise)
int dvae_recon_out_bwd(int dtype, const float* g_rec, const float* g_hat, void* d_rec, void* d_post, int R, int C, int T,
                       float scale, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (R == 0) return 0;
  const int smem = 2 * C * (T + 1) * 4;
  DVAE_REQUIRE(smem <= 48 * 1024, "tile must fit 48 KB of shared memory");
  DISPATCH_AT(dtype, recon_out_bwd_kernel<AT><<<R, 256, smem, st>>>(g_rec, g_hat, (AT*)d_rec, (AT*)d_post, C, T, scale));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Train-mode BatchNorm forward over y [halves*rows_half, C]: statistics per half, then act(y*scale+shift).
// ws: double [halves*2*C] scratch; stat: fp32 [halves*4*C] (kept for backward).
int dvae_bn_train_fwd(int dtype, const void* y, int y_f32, void* out, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, long long* num_batches, double* ws, float* stat, int rows_half, int halves, int C,
                      int act, float eps, float momentum, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(C % 8 == 0 && C <= 2048, "C must be a multiple of 8 (<= 2048)");
  DVAE_REQUIRE(rows_half % kBnRowsPerBlock == 0, "rows per half must be a multiple of 64");
  const long rows = static_cast<long>(rows_half) * halves;
  DVAE_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * halves * 2 * C, st));
  const int rb_red = bn_rows_per_block(rows_half, 512), rb_app = bn_rows_per_block(rows_half, 256);
  const dim3 g_red(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_red)), g_app(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_app));
  DISPATCH_AT_Y(dtype, y_f32, bn_stats_kernel<YT><<<g_red, 256, 0, st>>>((const YT*)y, ws, rows_half, C, rb_red));
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, gamma, beta, stat, run_mean, run_var, num_batches, halves, C,
                                                        static_cast<double>(rows_half), eps, momentum);
  DISPATCH_AT_Y(dtype, y_f32, bn_apply_kernel<AT, YT><<<g_app, 256, 0, st>>>((const YT*)y, (AT*)out, stat, rows, rows_half, C, act, rb_app));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// Second half of dvae_bn_train_fwd for a caller that already holds the statistics sums in ws (dvae_conv5_fwd_bnstats):
// finalise (mean / rstd / scale / shift, running statistics) and apply the affine transform + activation.
int dvae_bn_finalize_apply(int dtype, const void* y, int y_f32, void* out, const float* gamma, const float* beta, float* run_mean,
                           float* run_var, long long* num_batches, const double* ws, float* stat, int rows_half, int halves,
                           int C, int act, float eps, float momentum, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(C % 8 == 0 && C <= 2048, "C must be a multiple of 8 (<= 2048)");
  DVAE_REQUIRE(rows_half % kBnRowsPerBlock == 0, "rows per half must be a multiple of 64");
  const long rows = static_cast<long>(rows_half) * halves;
  const int rb_app = bn_rows_per_block(rows_half, 256);
  const dim3 g_app(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_app));
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, gamma, beta, stat, run_mean, run_var, num_batches, halves, C,
                                                        static_cast<double>(rows_half), eps, momentum);
  DISPATCH_AT_Y(dtype, y_f32, bn_apply_kernel<AT, YT><<<g_app, 256, 0, st>>>((const YT*)y, (AT*)out, stat, rows, rows_half, C, act, rb_app));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_bn_eval_fwd(int dtype, const void* y, void* out, const float* gamma, const float* beta, const float* run_mean,
                     const float* run_var, float* stat, long rows, int C, int act, float eps, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  bn_eval_stat_kernel<<<ceil_div(C, 128), 128, 0, st>>>(gamma, beta, run_mean, run_var, stat, C, eps);
  const int rows_half = rows > 0x7fffffffL ? 0x7fffffff : static_cast<int>(rows);
  const dim3 g_app(ceil_div(C, kBnSlab), static_cast<unsigned>(ceil_div(rows, 256)));
  if (rows > 0)
    DISPATCH_AT(dtype, bn_apply_kernel<AT><<<g_app, 256, 0, st>>>((const AT*)y, (AT*)out, stat, rows, rows_half > 0 ? rows_half : 1, C, act, 256));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// Train-mode BatchNorm backward (through the activation): dout -> dy, dgamma, dbeta.  coef: fp32 [halves*2*C] scratch.
// alpha scales the two parameter gradients (1 / gradient scale of the fp16 mode); dy stays at the stream's scale.
int dvae_bn_train_bwd(int dtype, const void* dout, const void* y, int y_f32, const float* stat, double* ws, float* coef, void* dy,
                      float* dgamma, float* dbeta, int rows_half, int halves, int C, int act, float alpha, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(C % 8 == 0 && C <= 2048, "C must be a multiple of 8 (<= 2048)");
  DVAE_REQUIRE(rows_half % kBnRowsPerBlock == 0, "rows per half must be a multiple of 64");
  const long rows = static_cast<long>(rows_half) * halves;
  DVAE_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * halves * 2 * C, st));
  const int rb_red = bn_rows_per_block(rows_half, 512), rb_app = bn_rows_per_block(rows_half, 256);
  const dim3 g_red(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_red)), g_app(ceil_div(C, kBnSlab), static_cast<unsigned>(rows / rb_app));
  DISPATCH_AT_Y(dtype, y_f32, bn_bwd_reduce_kernel<AT, YT><<<g_red, 256, 0, st>>>((const AT*)dout, (const YT*)y, stat, ws, rows_half, C, act, rb_red));
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, coef, dgamma, dbeta, halves, C, static_cast<double>(rows_half),
                                                            static_cast<double>(alpha));
  DISPATCH_AT_Y(dtype, y_f32, bn_bwd_apply_kernel<AT, YT><<<g_app, 256, 0, st>>>((const AT*)dout, (const YT*)y, stat, coef, (AT*)dy, rows, rows_half, C, act, rb_app));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out[C] (fp32) += alpha * column sums of x [rows, C] (row stride ldx)
int dvae_colsum(int dtype, const void* x, float* out, long rows, int C, long ldx, float alpha, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  if (rows == 0) return 0;
  const int ychunks = ceil_div(C, 2048);
  DVAE_REQUIRE(C <= 2048 || C % 2048 == 0, "C above 2048 must be a multiple of 2048");
  const int cw = C < 2048 ? C : 2048;
  const int tpr = cw / 8, lanes = 256 / tpr;
  long gx = (rows + 255) / 256;
  if (gx > 148 * 4) gx = 148 * 4;
  dim3 grid(static_cast<unsigned>(gx), ychunks);
  const int smem = lanes * cw * 4;
  DISPATCH_AT(dtype, colsum_kernel<AT><<<grid, 256, smem, st>>>((const AT*)x, out, rows, C, ldx, alpha));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
