// Fused Adam step over all parameter tensors in one launch.
//
// Reference: `self.optimizer = optim.Adam(self.model.parameters(), lr=...)` (model/disentangled_vae.py:304) stepped once
// per batch (model/variational_base_vae.py:69).  torch runs it as a handful of foreach kernels per step; here every
// element of every parameter is read and written exactly once: 16 bytes in (p, g, m, v), 12 bytes out (p, m, v), i.e.
// 28 B x 61.4 M parameters = 1.72 GB per step, an HBM stream.
//
// The arithmetic is torch.optim.Adam's default path (no weight decay, no amsgrad, maximize=False), in the same order:
//   m += (1 - beta1) * (g - m)                       (Tensor.lerp_)
//   v  = beta2 * v + (1 - beta2) * g * g             (mul_ then addcmul_)
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// Work is cut into fixed chunks; a table maps each block to (tensor, offset) so tensors of any size share one grid.
#include <cuda_runtime.h>

#include "host_common.h"

namespace dvae {

constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(float* const* __restrict__ ps, const float* const* __restrict__ gs, float* const* __restrict__ ms,
                  float* const* __restrict__ vs, const long* __restrict__ sizes, const int* __restrict__ blk_tensor,
                  const long* __restrict__ blk_off, int chunk, float w1, float beta2, float w2, float eps, float step_size,
                  float inv_bc2_sqrt, int* __restrict__ found_inf) {
  const int t = blk_tensor[blockIdx.x];
  const long off = blk_off[blockIdx.x];
  const long n = min(static_cast<long>(chunk), sizes[t] - off);
  float* __restrict__ p = ps[t] + off;
  const float* __restrict__ g = gs[t] + off;
  float* __restrict__ m = ms[t] + off;
  float* __restrict__ v = vs[t] + off;
  // An element whose gradient is not finite is left untouched (p, m, v) and reported through found_inf: the fp16 mode
  // carries gradients scaled into a 16-bit range, and an overflow there must not poison the weights (torch.optim.Adam
  // would write NaN; a GradScaler would skip the step -- the caller lowers the scale when the flag comes back).
  bool bad = false;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    if (!isfinite(gg)) { bad = true; return; }
    mm = mm + w1 * (gg - mm);
    vv = vv * beta2 + w2 * gg * gg;
    const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    // two independent 16-byte quads per tensor in flight per thread
    for (long i = threadIdx.x; i < n4; i += 2 * kAdamThreads) {
      const long j = i + kAdamThreads;
      const bool has2 = j < n4;
      float4 pa = p4[i], ga = __ldcs(g4 + i), ma = m4[i], va = v4[i];
      float4 pb, gb, mb, vb;
      if (has2) { pb = p4[j]; gb = __ldcs(g4 + j); mb = m4[j]; vb = v4[j]; }
      upd(pa.x, ga.x, ma.x, va.x); upd(pa.y, ga.y, ma.y, va.y); upd(pa.z, ga.z, ma.z, va.z); upd(pa.w, ga.w, ma.w, va.w);
      p4[i] = pa; m4[i] = ma; v4[i] = va;
      if (has2) {
        upd(pb.x, gb.x, mb.x, vb.x); upd(pb.y, gb.y, mb.y, vb.y); upd(pb.z, gb.z, mb.z, vb.z); upd(pb.w, gb.w, mb.w, vb.w);
        p4[j] = pb; m4[j] = mb; v4[j] = vb;
      }
    }
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += kAdamThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      upd(pp, g[i], mm, vv);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  } else {
    for (long i = threadIdx.x; i < n; i += kAdamThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      upd(pp, g[i], mm, vv);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
  if (bad && found_inf != nullptr) atomicOr(found_inf, 1);
}

}  // namespace dvae

extern "C" {

// One Adam step for every tensor in the tables (all arrays are DEVICE memory; `step` is the 1-based step count the bias
// corrections use).  blk_tensor / blk_off: for each block, which tensor and which element offset its chunk starts at.
// found_inf (device int, may be null): set to 1 when a gradient element was not finite; such elements are skipped.
int dvae_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long* sizes, const int* blk_tensor, const long* blk_off, int num_blocks, int chunk, double lr,
                   double beta1, double beta2, double eps, long step, int* found_inf, void* stream) {
  using namespace dvae;
  DVAE_REQUIRE(chunk > 0 && chunk % 4 == 0, "chunk must be a positive multiple of 4");
  DVAE_REQUIRE(step >= 1, "step counts from 1");
  if (num_blocks <= 0) return 0;
  // Hyper-parameters are doubles, like torch's Python scalars: 1 - beta and the bias corrections are formed in double and
  // rounded to fp32 once (1.f - 0.999f differs from float(1 - 0.999) by 1.3e-5 relative -- visible in exp_avg_sq).
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  adam_multi_kernel<<<num_blocks, kAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, sizes, blk_tensor, blk_off, chunk, static_cast<float>(1.0 - beta1),
      static_cast<float>(beta2), static_cast<float>(1.0 - beta2), static_cast<float>(eps), step_size, inv_bc2_sqrt, found_inf);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
