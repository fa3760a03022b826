// The memory-bound tail of the Disentangled-VAE step:
//   * latent tail  (reparameterisation + pair-mean style posterior with stop-gradient on member 2 + the six
//                   concatenations)                         model/disentangled_vae.py:252-272, fwd and bwd
//   * fused loss   (4 x L1-sum / batch_size, 2 x KL, style KL, weighted total)   model/disentangled_vae.py:310-327
//   * speaker-group ops keyed by a group id per row: product of Gaussians (model/utils.py:13-75), group mean
//     (model/variational_base_vae.py:281-282), group-wise reparameterisation (model/utils.py:95-116), with a
//     differentiable backward for the product of Gaussians.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "act_types.cuh"
#include "host_common.h"

namespace dvae {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------ latent tail
// heads fp32 [2R, 2L]: columns [style_mu(S) | style_logvar(S) | content_mu(L-S) | content_logvar(L-S)];
// rows [0,R) = x1 call, [R,2R) = x2 call.
template <typename AT>
__global__ void latent_tail_fwd_kernel(const float* __restrict__ heads, const float* __restrict__ eps_c1,
                                       const float* __restrict__ eps_c2, const float* __restrict__ eps_s,
                                       AT* __restrict__ z, float* __restrict__ q1_mu, float* __restrict__ q1_lv,
                                       float* __restrict__ q2_mu, float* __restrict__ q2_lv, float* __restrict__ zs_mu,
                                       float* __restrict__ zs_lv, int R, int L, int S, int sample_content) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * L) return;
  const int i = idx / L, j = idx - i * L;
  const int Lc = L - S, W = 2 * L;
  const float* h1 = heads + static_cast<long>(i) * W;
  const float* h2 = heads + static_cast<long>(R + i) * W;
  if (j < S) {
    const float mu = (h1[j] + h2[j]) * 0.5f;                 // :259
    const float lv = (h1[S + j] + h2[S + j]) * 0.5f;         // :260
    const float zs = eps_s[i * S + j] * expf(0.5f * lv) + mu;  // :261 (always sampled)
    z[static_cast<long>(i) * L + j] = from_f32<AT>(zs);
    z[static_cast<long>(R + i) * L + j] = from_f32<AT>(zs);
    q1_mu[i * L + j] = mu; q2_mu[i * L + j] = mu;
    q1_lv[i * L + j] = lv; q2_lv[i * L + j] = lv;
    zs_mu[i * S + j] = mu; zs_lv[i * S + j] = lv;
  } else {
    const int k = j - S;
    const float mu1 = h1[2 * S + k], lv1 = h1[2 * S + Lc + k];
    const float mu2 = h2[2 * S + k], lv2 = h2[2 * S + Lc + k];
    const float z1 = sample_content ? eps_c1[i * Lc + k] * expf(0.5f * lv1) + mu1 : mu1;   // :252
    const float z2 = sample_content ? eps_c2[i * Lc + k] * expf(0.5f * lv2) + mu2 : mu2;   // :255
    z[static_cast<long>(i) * L + j] = from_f32<AT>(z1);
    z[static_cast<long>(R + i) * L + j] = from_f32<AT>(z2);
    q1_mu[i * L + j] = mu1; q1_lv[i * L + j] = lv1;
    q2_mu[i * L + j] = mu2; q2_lv[i * L + j] = lv2;
  }
}

__device__ __forceinline__ float ld_or0(const float* p, long i) { return p ? p[i] : 0.f; }

// dz fp32 [2R, L]; dq*, dzs* may be null.  dheads act [2R, 2L]; member 2's style columns get zero (detach, :257-258).
// gscale: dz arrives (and dheads leaves) at the gradient stream's scale; dq* / dzs* come from the loss unscaled.
template <typename AT>
__global__ void latent_tail_bwd_kernel(const float* __restrict__ heads, const float* __restrict__ eps_c1,
                                       const float* __restrict__ eps_c2, const float* __restrict__ eps_s,
                                       const float* __restrict__ dz, const float* __restrict__ dq1_mu,
                                       const float* __restrict__ dq1_lv, const float* __restrict__ dq2_mu,
                                       const float* __restrict__ dq2_lv, const float* __restrict__ dzs_mu,
                                       const float* __restrict__ dzs_lv, AT* __restrict__ dheads, int R, int L, int S,
                                       int sample_content, float gscale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * L) return;
  const int i = idx / L, j = idx - i * L;
  const int Lc = L - S, W = 2 * L;
  const float* h1 = heads + static_cast<long>(i) * W;
  const float* h2 = heads + static_cast<long>(R + i) * W;
  AT* d1 = dheads + static_cast<long>(i) * W;
  AT* d2 = dheads + static_cast<long>(R + i) * W;
  const float dz1 = dz[static_cast<long>(i) * L + j], dz2 = dz[static_cast<long>(R + i) * L + j];
  if (j < S) {
    const float lv = (h1[S + j] + h2[S + j]) * 0.5f;
    const float dzs = dz1 + dz2;
    const float dmu = dzs + gscale * ld_or0(dq1_mu, i * L + j) + gscale * ld_or0(dq2_mu, i * L + j) + gscale * ld_or0(dzs_mu, i * S + j);
    const float dlv = dzs * eps_s[i * S + j] * 0.5f * expf(0.5f * lv) + gscale * ld_or0(dq1_lv, i * L + j) +
                      gscale * ld_or0(dq2_lv, i * L + j) + gscale * ld_or0(dzs_lv, i * S + j);
    d1[j] = from_f32<AT>(0.5f * dmu);
    d1[S + j] = from_f32<AT>(0.5f * dlv);
    d2[j] = from_f32<AT>(0.f);
    d2[S + j] = from_f32<AT>(0.f);
  } else {
    const int k = j - S;
    const float lv1 = h1[2 * S + Lc + k], lv2 = h2[2 * S + Lc + k];
    float dmu1 = dz1 + gscale * ld_or0(dq1_mu, i * L + j), dlv1 = gscale * ld_or0(dq1_lv, i * L + j);
    float dmu2 = dz2 + gscale * ld_or0(dq2_mu, i * L + j), dlv2 = gscale * ld_or0(dq2_lv, i * L + j);
    if (sample_content) {
      dlv1 += dz1 * eps_c1[i * Lc + k] * 0.5f * expf(0.5f * lv1);
      dlv2 += dz2 * eps_c2[i * Lc + k] * 0.5f * expf(0.5f * lv2);
    }
    d1[2 * S + k] = from_f32<AT>(dmu1);
    d1[2 * S + Lc + k] = from_f32<AT>(dlv1);
    d2[2 * S + k] = from_f32<AT>(dmu2);
    d2[2 * S + Lc + k] = from_f32<AT>(dlv2);
  }
}

// ------------------------------------------------------------------------------------ fused loss
// acc (double) [8]: 0..3 L1 sums (x1/r1, x2/r2, x1/h1, x2/h2), 4..5 KL sums (q1, q2), 6 style sum; acc[7] unused.
// ticket: unsigned counter; the last block to finish writes the 8 outputs.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  return r;  // valid in thread 0
}

__global__ void loss_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ r1,
                                const float* __restrict__ r2, const float* __restrict__ h1, const float* __restrict__ h2,
                                long n, const float* __restrict__ q1_mu, const float* __restrict__ q1_lv,
                                const float* __restrict__ q2_mu, const float* __restrict__ q2_lv, long nq, int q_rows,
                                const float* __restrict__ s_mu, const float* __restrict__ s_lv, long ns, float batch_size,
                                float mse_cof, float kl_cof, double* __restrict__ acc, unsigned int* __restrict__ ticket,
                                float* __restrict__ out) {
  __shared__ float sh[32];
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const long n4 = n >> 2;
  for (long i = tid; i < n4; i += nthreads) {
    const float4 a1 = reinterpret_cast<const float4*>(x1)[i], a2 = reinterpret_cast<const float4*>(x2)[i];
    const float4 b1 = reinterpret_cast<const float4*>(r1)[i], b2 = reinterpret_cast<const float4*>(r2)[i];
    const float4 c1 = reinterpret_cast<const float4*>(h1)[i], c2 = reinterpret_cast<const float4*>(h2)[i];
    s[0] += fabsf(a1.x - b1.x) + fabsf(a1.y - b1.y) + fabsf(a1.z - b1.z) + fabsf(a1.w - b1.w);
    s[1] += fabsf(a2.x - b2.x) + fabsf(a2.y - b2.y) + fabsf(a2.z - b2.z) + fabsf(a2.w - b2.w);
    s[2] += fabsf(a1.x - c1.x) + fabsf(a1.y - c1.y) + fabsf(a1.z - c1.z) + fabsf(a1.w - c1.w);
    s[3] += fabsf(a2.x - c2.x) + fabsf(a2.y - c2.y) + fabsf(a2.z - c2.z) + fabsf(a2.w - c2.w);
  }
  for (long i = (n4 << 2) + tid; i < n; i += nthreads) {
    s[0] += fabsf(x1[i] - r1[i]); s[1] += fabsf(x2[i] - r2[i]);
    s[2] += fabsf(x1[i] - h1[i]); s[3] += fabsf(x2[i] - h2[i]);
  }
  for (long i = tid; i < nq; i += nthreads) {
    const float m1 = q1_mu[i], l1 = q1_lv[i], m2 = q2_mu[i], l2 = q2_lv[i];
    s[4] += 1.f + l1 - m1 * m1 - expf(l1);
    s[5] += 1.f + l2 - m2 * m2 - expf(l2);
  }
  for (long i = tid; i < ns; i += nthreads) {
    const float m = s_mu[i], l = s_lv[i];
    s[6] += 1.f + l - m * m - expf(l);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const float b = block_sum(s[k], sh);
    if (threadIdx.x == 0) atomicAdd(acc + k, static_cast<double>(b));
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    volatile double* a = acc;
    const double bs = batch_size;
    const double m1 = a[0] / bs, m2 = a[1] / bs, m1h = a[2] / bs, m2h = a[3] / bs;   // :314-318
    const double k1 = -0.5 * a[4] / q_rows, k2 = -0.5 * a[5] / q_rows;               // :320-321 (mean over rows)
    const double ks = -1.0 * a[6] / bs;                                              // :323
    out[0] = static_cast<float>(mse_cof * (m1 + m2 + m1h + m2h) + kl_cof * (k1 + k2)); // :325
    out[1] = static_cast<float>(m1); out[2] = static_cast<float>(m2);
    out[3] = static_cast<float>(m1h); out[4] = static_cast<float>(m2h);
    out[5] = static_cast<float>(k1); out[6] = static_cast<float>(k2); out[7] = static_cast<float>(ks);
  }
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

// gout fp32 [8]: upstream gradients of the 8 returned scalars (normally [1,0,...,0]).
__global__ void loss_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ r1,
                                const float* __restrict__ r2, const float* __restrict__ h1, const float* __restrict__ h2,
                                long n, const float* __restrict__ q1_mu, const float* __restrict__ q1_lv,
                                const float* __restrict__ q2_mu, const float* __restrict__ q2_lv, long nq, int q_rows,
                                const float* __restrict__ s_mu, const float* __restrict__ s_lv, long ns, float batch_size,
                                float mse_cof, float kl_cof, const float* __restrict__ gout, float* __restrict__ dr1,
                                float* __restrict__ dr2, float* __restrict__ dh1, float* __restrict__ dh2,
                                float* __restrict__ dq1_mu, float* __restrict__ dq1_lv, float* __restrict__ dq2_mu,
                                float* __restrict__ dq2_lv, float* __restrict__ ds_mu, float* __restrict__ ds_lv) {
  const float g0 = gout[0];
  const float c1 = (g0 * mse_cof + gout[1]) / batch_size, c2 = (g0 * mse_cof + gout[2]) / batch_size;
  const float c1h = (g0 * mse_cof + gout[3]) / batch_size, c2h = (g0 * mse_cof + gout[4]) / batch_size;
  const float ck1 = (g0 * kl_cof + gout[5]) * (-0.5f) / q_rows, ck2 = (g0 * kl_cof + gout[6]) * (-0.5f) / q_rows;
  const float cks = gout[7] * (-1.f) / batch_size;
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const long n4 = n >> 2;
  for (long i = tid; i < n4; i += nthreads) {
    const float4 a1 = reinterpret_cast<const float4*>(x1)[i], a2 = reinterpret_cast<const float4*>(x2)[i];
    const float4 b1 = reinterpret_cast<const float4*>(r1)[i], b2 = reinterpret_cast<const float4*>(r2)[i];
    const float4 e1 = reinterpret_cast<const float4*>(h1)[i], e2 = reinterpret_cast<const float4*>(h2)[i];
    // d|x - r|/dr = sign(r - x)
    reinterpret_cast<float4*>(dr1)[i] = make_float4(c1 * sgn(b1.x - a1.x), c1 * sgn(b1.y - a1.y), c1 * sgn(b1.z - a1.z), c1 * sgn(b1.w - a1.w));
    reinterpret_cast<float4*>(dr2)[i] = make_float4(c2 * sgn(b2.x - a2.x), c2 * sgn(b2.y - a2.y), c2 * sgn(b2.z - a2.z), c2 * sgn(b2.w - a2.w));
    reinterpret_cast<float4*>(dh1)[i] = make_float4(c1h * sgn(e1.x - a1.x), c1h * sgn(e1.y - a1.y), c1h * sgn(e1.z - a1.z), c1h * sgn(e1.w - a1.w));
    reinterpret_cast<float4*>(dh2)[i] = make_float4(c2h * sgn(e2.x - a2.x), c2h * sgn(e2.y - a2.y), c2h * sgn(e2.z - a2.z), c2h * sgn(e2.w - a2.w));
  }
  for (long i = (n4 << 2) + tid; i < n; i += nthreads) {
    dr1[i] = c1 * sgn(r1[i] - x1[i]); dr2[i] = c2 * sgn(r2[i] - x2[i]);
    dh1[i] = c1h * sgn(h1[i] - x1[i]); dh2[i] = c2h * sgn(h2[i] - x2[i]);
  }
  for (long i = tid; i < nq; i += nthreads) {
    dq1_mu[i] = ck1 * (-2.f * q1_mu[i]); dq1_lv[i] = ck1 * (1.f - expf(q1_lv[i]));
    dq2_mu[i] = ck2 * (-2.f * q2_mu[i]); dq2_lv[i] = ck2 * (1.f - expf(q2_lv[i]));
  }
  for (long i = tid; i < ns; i += nthreads) {
    ds_mu[i] = cks * (-2.f * s_mu[i]); ds_lv[i] = cks * (1.f - expf(s_lv[i]));
  }
}

// ------------------------------------------------------------------------------------ speaker-group ops
// gid int32 [B]: group index per row.  Any order is correct; rows of a group that are adjacent (the normal
// case: batches are sorted by speaker) are reduced in registers and with warp shuffles before touching memory.
constexpr int kModePoG = 0, kModeMean = 1, kModeRaw = 2;
constexpr int kGroupChunk = 8;  // consecutive rows per row-lane

// sorted labels -> gid: boundary flags + two-level exclusive scan (block = 1024 rows)
__global__ void seg_flag_count_kernel(const long long* __restrict__ labels, int* __restrict__ block_counts, long B) {
  __shared__ int sh[32];
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  int f = (i < B && i > 0 && labels[i] != labels[i - 1]) ? 1 : 0;
  int v = f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = t;
  }
}
__global__ void seg_scan_blocks_kernel(int* __restrict__ block_counts, int nblocks, int* __restrict__ num_groups) {
  // single block: exclusive scan of block_counts in place (nblocks <= 2^22 / 1024 in practice; loop in chunks)
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_counts[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_counts[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_groups = carry + 1;
}
__global__ void seg_write_gid_kernel(const long long* __restrict__ labels, const int* __restrict__ block_offsets,
                                     int* __restrict__ gid, long B) {
  __shared__ int sh[1024];
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int f = (i < B && i > 0 && labels[i] != labels[i - 1]) ? 1 : 0;
  sh[threadIdx.x] = f;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < B) gid[i] = block_offsets[blockIdx.x] + sh[threadIdx.x];
}

__device__ __forceinline__ void group_flush(float* __restrict__ acc, float* __restrict__ cnt, int g, int D, int c,
                                            const float4& a, const float4& b, float n) {
  float* pa = acc + (static_cast<long>(g) * 2 + 0) * D + c * 4;
  float* pb = acc + (static_cast<long>(g) * 2 + 1) * D + c * 4;
  atomicAdd(pa + 0, a.x); atomicAdd(pa + 1, a.y); atomicAdd(pa + 2, a.z); atomicAdd(pa + 3, a.w);
  atomicAdd(pb + 0, b.x); atomicAdd(pb + 1, b.y); atomicAdd(pb + 2, b.z); atomicAdd(pb + 3, b.w);
  if (cnt != nullptr && c == 0) atomicAdd(cnt + g, n);
}

// acc fp32 [G][2][D] += per-group sums of (u, v):  PoG: u = 1/var, v = mu/var (var = exp(logvar), exact zeros -> 1e-6);
// Mean / Raw: u = a, v = b.  cnt fp32 [G] += rows per group (may be null).
template <int CPR>  // float4 columns per row = D / 4  (1, 2, 4 or 8)
__global__ void group_accumulate_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                        const int* __restrict__ gid, float* __restrict__ acc, float* __restrict__ cnt,
                                        long B, int mode) {
  constexpr int D = CPR * 4;
  constexpr int RL = 32 / CPR;  // row lanes per warp
  const int lane = threadIdx.x & 31;
  const int q = lane / CPR, c = lane - q * CPR;
  const long warp_global = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const long rows_per_warp = RL * kGroupChunk;
  for (long base = warp_global * rows_per_warp; base < B; base += nwarps * rows_per_warp) {
    const long r0 = base + static_cast<long>(q) * kGroupChunk;
    float4 su = make_float4(0.f, 0.f, 0.f, 0.f), sv = su;
    float n = 0.f;
    int cur = -1;
    bool single = true;  // the whole chunk is one run
    // all loads of the chunk first (16 independent 16-byte loads per lane in flight): the kernel is a pure HBM stream
    float4 ua[kGroupChunk], va[kGroupChunk];
    int ga[kGroupChunk];
#pragma unroll
    for (int k = 0; k < kGroupChunk; ++k) {
      const long r = r0 + k;
      if (r < B) {
        ga[k] = __ldg(gid + r);
        ua[k] = __ldg(reinterpret_cast<const float4*>(a + r * D) + c);
        va[k] = __ldg(reinterpret_cast<const float4*>(b + r * D) + c);
      }
    }
#pragma unroll
    for (int k = 0; k < kGroupChunk; ++k) {
      const long r = r0 + k;
      if (r < B) {
        const int g = ga[k];
        float4 u = ua[k];
        float4 v = va[k];
        if (mode == kModePoG) {
          float4 var = make_float4(expf(v.x), expf(v.y), expf(v.z), expf(v.w));
          var.x = var.x == 0.f ? 1e-6f : var.x; var.y = var.y == 0.f ? 1e-6f : var.y;
          var.z = var.z == 0.f ? 1e-6f : var.z; var.w = var.w == 0.f ? 1e-6f : var.w;
          const float4 p = make_float4(1.f / var.x, 1.f / var.y, 1.f / var.z, 1.f / var.w);
          v = make_float4(u.x * p.x, u.y * p.y, u.z * p.z, u.w * p.w);
          u = p;
        }
        if (g != cur) {
          if (cur >= 0) {
            group_flush(acc, cnt, cur, D, c, su, sv, n);
            single = false;
          }
          cur = g; su = make_float4(0.f, 0.f, 0.f, 0.f); sv = su; n = 0.f;
        }
        su.x += u.x; su.y += u.y; su.z += u.z; su.w += u.w;
        sv.x += v.x; sv.y += v.y; sv.z += v.z; sv.w += v.w;
        n += 1.f;
      }
    }
    // Segmented warp reduction over the row lanes, keyed by group id: lanes whose whole chunk is one run and that sit
    // next to a lane with the same id are folded together with shuffles; only run heads touch memory.
    const int key = single ? cur : -2 - q;  // multi-run chunks never merge (their tail is flushed on its own)
    const int up_key = __shfl_up_sync(0xffffffffu, key, CPR);
    const bool head = (q == 0) || (up_key != key) || (key < 0);
    int hidx = head ? q : 0;
#pragma unroll
    for (int o = CPR; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, hidx, o);
      if (lane >= o) hidx = max(hidx, t);
    }
#pragma unroll
    for (int o = CPR; o < 32; o <<= 1) {
      const int h2 = __shfl_down_sync(0xffffffffu, hidx, o);
      const float4 u2 = make_float4(__shfl_down_sync(0xffffffffu, su.x, o), __shfl_down_sync(0xffffffffu, su.y, o),
                                    __shfl_down_sync(0xffffffffu, su.z, o), __shfl_down_sync(0xffffffffu, su.w, o));
      const float4 v2 = make_float4(__shfl_down_sync(0xffffffffu, sv.x, o), __shfl_down_sync(0xffffffffu, sv.y, o),
                                    __shfl_down_sync(0xffffffffu, sv.z, o), __shfl_down_sync(0xffffffffu, sv.w, o));
      const float n2 = __shfl_down_sync(0xffffffffu, n, o);
      if (lane + o < 32 && h2 == hidx) {
        su.x += u2.x; su.y += u2.y; su.z += u2.z; su.w += u2.w;
        sv.x += v2.x; sv.y += v2.y; sv.z += v2.z; sv.w += v2.w;
        n += n2;
      }
    }
    if (head && cur >= 0) group_flush(acc, cnt, cur, D, c, su, sv, n);
  }
}

// group results, computed ONCE per group (not per row as in the reference's Python loop):
// PoG: var_g = 1/acc0 (exact zero -> 1e-6), mu_g = acc1 * var_g, logvar_g = log var_g.   Mean: acc / cnt.
__global__ void group_table_kernel(const float* __restrict__ acc, const float* __restrict__ cnt, float* __restrict__ table,
                                   long G, int D, int mode) {
  const long total = G * D;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long g = i / D;
    const int d = static_cast<int>(i - g * D);
    const float u = acc[(g * 2 + 0) * D + d], v = acc[(g * 2 + 1) * D + d];
    float oa, ob;
    if (mode == kModePoG) {
      float var = 1.f / u;
      oa = v * var;
      var = var == 0.f ? 1e-6f : var;
      ob = logf(var);
    } else {
      const float inv = 1.f / cnt[g];
      oa = u * inv;
      ob = v * inv;
    }
    table[(g * 2 + 0) * D + d] = oa;
    table[(g * 2 + 1) * D + d] = ob;
  }
}
// rows <- their group's entry: a pure gather / streaming-store pass (the table is small and stays in L2)
template <int CPR>
__global__ void group_broadcast_kernel(const float* __restrict__ table, const int* __restrict__ gid,
                                       float* __restrict__ out_a, float* __restrict__ out_b, long B) {
  constexpr int D = CPR * 4;
  const long total = B * CPR;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / CPR;
    const int c = static_cast<int>(i - r * CPR);
    const int g = __ldg(gid + r);
    const float4 oa = __ldg(reinterpret_cast<const float4*>(table + (static_cast<long>(g) * 2 + 0) * D) + c);
    __stcs(reinterpret_cast<float4*>(out_a + r * D) + c, oa);
    if (out_b != nullptr) {
      const float4 ob = __ldg(reinterpret_cast<const float4*>(table + (static_cast<long>(g) * 2 + 1) * D) + c);
      __stcs(reinterpret_cast<float4*>(out_b + r * D) + c, ob);
    }
  }
}

// product-of-Gaussians backward per row.  acc_f = forward accumulators (P = sum p, M = sum p mu); acc_g = per-group sums
// of the incoming (d group_mu, d group_logvar).  d mu_i = dMu p_i / P;  d logvar_i = -p_i (dMu (mu_i - mu_g) - dLv) / P.
template <int CPR>
__global__ void group_pog_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                     const int* __restrict__ gid, const float* __restrict__ acc_f,
                                     const float* __restrict__ acc_g, float* __restrict__ dmu, float* __restrict__ dlv,
                                     long B) {
  constexpr int D = CPR * 4;
  const long total = B * D;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / D;
    const int d = static_cast<int>(i - r * D);
    const int g = gid[r];
    const float P = acc_f[(static_cast<long>(g) * 2 + 0) * D + d], M = acc_f[(static_cast<long>(g) * 2 + 1) * D + d];
    const float dMu = acc_g[(static_cast<long>(g) * 2 + 0) * D + d], dLv = acc_g[(static_cast<long>(g) * 2 + 1) * D + d];
    float var = expf(logvar[i]);
    var = var == 0.f ? 1e-6f : var;
    const float p = 1.f / var;
    const float mug = M / P;
    dmu[i] = dMu * p / P;
    dlv[i] = -p * (dMu * (mu[i] - mug) - dLv) / P;
  }
}

// z_i = exp(0.5 logvar_i) * eps[g(i)] + mu_i
__global__ void group_reparam_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                     const int* __restrict__ gid, const float* __restrict__ eps_group,
                                     float* __restrict__ z, long B, int D) {
  const long total = B * D;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / D;
    const int d = static_cast<int>(i - r * D);
    z[i] = expf(0.5f * logvar[i]) * eps_group[static_cast<long>(gid[r]) * D + d] + mu[i];
  }
}

static int blocks_for(long threads_needed, int block) {
  long g = (threads_needed + block - 1) / block;
  const long cap = static_cast<long>(num_sms()) * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
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

#define DISPATCH_CPR(D, ...)                                                         \
  do {                                                                               \
    switch ((D) / 4) {                                                               \
      case 1: { constexpr int CPR = 1; __VA_ARGS__; } break;                         \
      case 2: { constexpr int CPR = 2; __VA_ARGS__; } break;                         \
      case 4: { constexpr int CPR = 4; __VA_ARGS__; } break;                         \
      case 8: { constexpr int CPR = 8; __VA_ARGS__; } break;                         \
      default: set_last_error("group ops support D in {4, 8, 16, 32}"); return 1;    \
    }                                                                                \
  } while (0)

extern "C" {

int dvae_latent_tail_fwd(int dtype, const float* heads, const float* eps_c1, const float* eps_c2, const float* eps_s,
                         void* z, float* q1_mu, float* q1_lv, float* q2_mu, float* q2_lv, float* zs_mu, float* zs_lv, int R,
                         int L, int S, int sample_content, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (R == 0) return 0;
  DVAE_REQUIRE(S > 0 && S < L, "need 0 < speaker_size < latent_dim");
  DVAE_REQUIRE(!sample_content || (eps_c1 && eps_c2), "content noise required when sampling");
  DISPATCH_AT(dtype, latent_tail_fwd_kernel<AT><<<ceil_div((long)R * L, 256), 256, 0, st>>>(
                         heads, eps_c1, eps_c2, eps_s, (AT*)z, q1_mu, q1_lv, q2_mu, q2_lv, zs_mu, zs_lv, R, L, S, sample_content));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int dvae_latent_tail_bwd(int dtype, const float* heads, const float* eps_c1, const float* eps_c2, const float* eps_s,
                         const float* dz, const float* dq1_mu, const float* dq1_lv, const float* dq2_mu, const float* dq2_lv,
                         const float* dzs_mu, const float* dzs_lv, void* dheads, int R, int L, int S, int sample_content,
                         float gscale, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (R == 0) return 0;
  DISPATCH_AT(dtype, latent_tail_bwd_kernel<AT><<<ceil_div((long)R * L, 256), 256, 0, st>>>(
                         heads, eps_c1, eps_c2, eps_s, dz, dq1_mu, dq1_lv, dq2_mu, dq2_lv, dzs_mu, dzs_lv, (AT*)dheads, R, L, S,
                         sample_content, gscale));
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ws: 64 bytes of double accumulators + 4-byte ticket (72 bytes total, zeroed here).  out fp32 [8].
int dvae_loss_fwd(const float* x1, const float* x2, const float* r1, const float* r2, const float* h1, const float* h2, long n,
                  const float* q1_mu, const float* q1_lv, const float* q2_mu, const float* q2_lv, int q_rows, int L,
                  const float* s_mu, const float* s_lv, int S, float batch_size, float mse_cof, float kl_cof, void* ws,
                  float* out, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(batch_size > 0 && q_rows > 0, "batch_size and q_rows must be positive");
  DVAE_CHECK_CUDA(cudaMemsetAsync(ws, 0, 72, st));
  double* acc = static_cast<double*>(ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 64);
  const int blocks = blocks_for((n + 3) / 4, 256);
  loss_fwd_kernel<<<blocks, 256, 0, st>>>(x1, x2, r1, r2, h1, h2, n, q1_mu, q1_lv, q2_mu, q2_lv, (long)q_rows * L, q_rows,
                                          s_mu, s_lv, (long)q_rows * S, batch_size, mse_cof, kl_cof, acc, ticket, out);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int dvae_loss_bwd(const float* x1, const float* x2, const float* r1, const float* r2, const float* h1, const float* h2, long n,
                  const float* q1_mu, const float* q1_lv, const float* q2_mu, const float* q2_lv, int q_rows, int L,
                  const float* s_mu, const float* s_lv, int S, float batch_size, float mse_cof, float kl_cof,
                  const float* gout, float* dr1, float* dr2, float* dh1, float* dh2, float* dq1_mu, float* dq1_lv,
                  float* dq2_mu, float* dq2_lv, float* ds_mu, float* ds_lv, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  const int blocks = blocks_for((n + 3) / 4, 256);
  loss_bwd_kernel<<<blocks, 256, 0, st>>>(x1, x2, r1, r2, h1, h2, n, q1_mu, q1_lv, q2_mu, q2_lv, (long)q_rows * L, q_rows,
                                          s_mu, s_lv, (long)q_rows * S, batch_size, mse_cof, kl_cof, gout, dr1, dr2, dh1, dh2,
                                          dq1_mu, dq1_lv, dq2_mu, dq2_lv, ds_mu, ds_lv);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// labels sorted (equal ids adjacent).  gid int32 [B]; scratch int32 [ceil(B/1024)]; num_groups: device int.
int dvae_segment_ids_sorted(const long long* labels, int* gid, int* scratch, int* num_groups, long B, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B == 0) {
    DVAE_CHECK_CUDA(cudaMemsetAsync(num_groups, 0, sizeof(int), st));
    return 0;
  }
  const int nblocks = ceil_div(B, 1024);
  seg_flag_count_kernel<<<nblocks, 1024, 0, st>>>(labels, scratch, B);
  seg_scan_blocks_kernel<<<1, 1024, 0, st>>>(scratch, nblocks, num_groups);
  seg_write_gid_kernel<<<nblocks, 1024, 0, st>>>(labels, scratch, gid, B);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// mode 0: product of Gaussians (a = mu, b = logvar); 1: mean (a, b summed, cnt counted); 2: raw sums.
// acc fp32 [G][2][D] and cnt fp32 [G] must be zeroed by the caller.
int dvae_group_accumulate(int mode, const float* a, const float* b, const int* gid, float* acc, float* cnt, long B, int D,
                          void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  DVAE_REQUIRE(D % 4 == 0, "D must be a multiple of 4");
  DISPATCH_CPR(D, {
    const long rows_per_warp = (32 / CPR) * kGroupChunk;
    const long warps = (B + rows_per_warp - 1) / rows_per_warp;
    group_accumulate_kernel<CPR><<<blocks_for(warps * 32, 256), 256, 0, st>>>(a, b, gid, acc, cnt, B, mode);
  });
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// table: fp32 [G][2][D] scratch that receives the per-group results; rows then gather from it
int dvae_group_finalize(int mode, const float* acc, const float* cnt, const int* gid, float* table, float* out_a,
                        float* out_b, long B, long G, int D, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B == 0 || G == 0) return 0;
  group_table_kernel<<<blocks_for(G * D, 256), 256, 0, st>>>(acc, cnt, table, G, D, mode);
  DISPATCH_CPR(D, { group_broadcast_kernel<CPR><<<blocks_for(B * CPR, 256), 256, 0, st>>>(table, gid, out_a, out_b, B); });
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_group_pog_bwd(const float* mu, const float* logvar, const int* gid, const float* acc_f, const float* acc_g,
                       float* dmu, float* dlv, long B, int D, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  DISPATCH_CPR(D, { group_pog_bwd_kernel<CPR><<<blocks_for(B * D, 256), 256, 0, st>>>(mu, logvar, gid, acc_f, acc_g, dmu, dlv, B); });
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dvae_group_reparam(const float* mu, const float* logvar, const int* gid, const float* eps_group, float* z, long B, int D,
                       void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  group_reparam_kernel<<<blocks_for(B * D, 256), 256, 0, st>>>(mu, logvar, gid, eps_group, z, B, D);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
