// Activation-storage helpers shared by the GEMM epilogues and the memory-bound kernels: 8-wide (16 / 32 byte)
// vector load/store of bf16 or fp32 activations, scalar conversions, fast gate non-linearities.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dvae {

template <typename T>
struct Act8;  // pack / unpack 8 consecutive activations
template <>
struct Act8<float> {
  struct raw_t { float4 a, b; };   // 8 values still "in flight": loads can be issued long before they are consumed
  static __device__ __forceinline__ raw_t load_raw(const float* p) {
    raw_t r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Act8<__nv_bfloat16> {
  using raw_t = uint4;
  static __device__ __forceinline__ raw_t load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const raw_t& u, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
// fp16 storage: 10 mantissa bits (the tf32 grid) at the bf16 byte count and tensor-core rate (tcgen05 kind::f16 takes
// either 16-bit format).  Range is the price: the engine keeps the activation-GRADIENT stream scaled by a power of two
// (DESIGN.md "Numerics") so that it stays inside fp16's normal range.
template <>
struct Act8<__half> {
  using raw_t = uint4;
  static __device__ __forceinline__ raw_t load_raw(const __half* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const raw_t& u, float* v) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void load(const __half* p, float* v) { unpack(load_raw(p), v); }
  static __device__ __forceinline__ void store(__half* p, const float* v) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
// fp32 storage whose values are kept on the tf32 grid (10 mantissa bits, round-to-nearest).  tcgen05 kind::tf32
// TRUNCATES the low 13 bits of its fp32 operands; truncation is a biased error that compounds through un-normalised
// layers, so every tensor that feeds a tf32 MMA is rounded once, when it is stored.
struct tf32_t {
  float x;
};
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
template <>
struct Act8<tf32_t> {
  using raw_t = Act8<float>::raw_t;
  static __device__ __forceinline__ raw_t load_raw(const tf32_t* p) { return Act8<float>::load_raw(reinterpret_cast<const float*>(p)); }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) { Act8<float>::unpack(r, v); }
  static __device__ __forceinline__ void load(const tf32_t* p, float* v) {
    Act8<float>::load(reinterpret_cast<const float*>(p), v);
  }
  static __device__ __forceinline__ void store(tf32_t* p, const float* v) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = round_tf32(v[i]);
    Act8<float>::store(reinterpret_cast<float*>(p), r);
  }
};
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<tf32_t>(tf32_t v) { return v.x; }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ tf32_t from_f32<tf32_t>(float v) { tf32_t t; t.x = round_tf32(v); return t; }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// 8 fp32 values -> one 16-byte chunk of a 16-bit storage type (round to nearest): the staged GEMM epilogues build their
// shared-memory boxes from these
template <typename T> __device__ __forceinline__ uint4 pack8_16(const float* v);
template <> __device__ __forceinline__ uint4 pack8_16<__nv_bfloat16>(const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
template <> __device__ __forceinline__ uint4 pack8_16<__half>(const float* v) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
// 4 fp32 values -> one 16-byte chunk of a 32-bit storage type (tf32_t: rounded onto the tf32 grid; float: as is)
template <typename T> __device__ __forceinline__ uint4 pack4_32(const float* v);
// one 16-bit storage element (raw bits) -> fp32
template <typename T> __device__ __forceinline__ float bits16_to_f32(unsigned short u);
template <> __device__ __forceinline__ float bits16_to_f32<__nv_bfloat16>(unsigned short u) { return __uint_as_float(static_cast<uint32_t>(u) << 16); }
template <> __device__ __forceinline__ float bits16_to_f32<__half>(unsigned short u) { return __half2float(__ushort_as_half(u)); }
// MMA operand format of a storage type: instruction-descriptor a_format / b_format (kind::f16: 0 = F16, 1 = BF16;
// kind::tf32: 2 = TF32)
template <typename T> struct MmaFmt;
template <> struct MmaFmt<__half> { static constexpr uint32_t value = 0; };
template <> struct MmaFmt<__nv_bfloat16> { static constexpr uint32_t value = 1; };
template <> struct MmaFmt<tf32_t> { static constexpr uint32_t value = 2; };
template <> struct MmaFmt<float> { static constexpr uint32_t value = 2; };
template <> __device__ __forceinline__ uint4 pack4_32<tf32_t>(const float* v) {
  return make_uint4(__float_as_uint(round_tf32(v[0])), __float_as_uint(round_tf32(v[1])), __float_as_uint(round_tf32(v[2])),
                    __float_as_uint(round_tf32(v[3])));
}
template <> __device__ __forceinline__ uint4 pack4_32<float>(const float* v) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
// 16 / sizeof(T) fp32 values -> one 16-byte chunk of storage type T
template <typename T> __device__ __forceinline__ uint4 pack_chunk(const float* v) {
  if constexpr (sizeof(T) == 2) return pack8_16<T>(v);
  else return pack4_32<T>(v);
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }


// Gate non-linearities of the LSTM cells, selected by the storage type.  fp32-on-the-tf32-grid storage keeps the exp-based
// forms (2 MUFU ops each).  bf16 storage uses the hardware tanh (1 MUFU op; absolute error ~5e-4, below the 2e-3 rounding
// step of a bf16 gate value near 1): the cell epilogues are MUFU-bound -- 10 special-function ops per hidden unit and row
// with the exp forms, 5 with these.  DVAE_GATE_APPROX=0 at compile time restores the exp forms everywhere.
#ifndef DVAE_GATE_APPROX
#define DVAE_GATE_APPROX 1
#endif
__device__ __forceinline__ float tanh_hw(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename ActT>
struct GateMath {
  static __device__ __forceinline__ float sig(float x) { return sigmoid_f(x); }
  static __device__ __forceinline__ float tnh(float x) { return tanh_f(x); }
};
#if DVAE_GATE_APPROX
template <>
struct GateMath<__nv_bfloat16> {
  static __device__ __forceinline__ float sig(float x) { return fmaf(tanh_hw(0.5f * x), 0.5f, 0.5f); }
  static __device__ __forceinline__ float tnh(float x) { return tanh_hw(x); }
};
#endif
// (fp16 storage keeps the exp forms: tanh.approx's 2^-11 absolute error is as large as fp16's own rounding step near 1
// and would cost the precision the format is chosen for.)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dvae
