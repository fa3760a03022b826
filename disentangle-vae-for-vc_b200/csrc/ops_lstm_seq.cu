// Sequence-resident LSTM recurrence for small hidden sizes (H = 64: the encoder BiLSTM, reference
// model/disentangled_vae.py:163 and autovc_replicate/proposed_autovc.py:41).
//
// The step-per-launch path (ops_gemm.cu) pays ~10 us per time step for a GEMM that is 0.03 GFLOP: launch latency,
// pipeline fill and a cold epilogue.  When W_hh fits in shared memory the whole recurrence of a row tile is independent
// of every other row tile, so one CTA can run all T steps without leaving the SM:
//
//   * W_hh (bf16 32 KB / tf32 64 KB) is fetched once by TMA and stays in shared memory for all T steps;
//   * the recurrent operand (h_{t-1} forward, da_{t+1} backward) never goes through global memory: the epilogue threads
//     write it straight into the 128-byte-swizzled K-major A tile that the next tcgen05.mma reads
//     (st.shared -> fence.proxy.async -> bar.sync -> MMA);
//   * the cell state c (forward) and the dc carry (backward) live in registers for the whole sequence;
//   * a CTA owns 32 rows.  tcgen05.mma is M = 128, so the 32 rows are replicated into all four 32-lane TMEM quarters:
//     warp q then reads "its" copy of the accumulator and owns hidden units [16q, 16q + 16) -- four warps share the
//     transcendental work of a row, and a 1024-row batch x 2 directions spreads over 64 SMs instead of 16;
//   * everything a step needs from global memory besides the accumulator (x-projection / saved gates / dh / c) is
//     requested one step ahead, so the per-step critical path is MMA -> tcgen05.ld -> cell -> st.shared -> MMA.
//
// Layouts are those of the step-per-launch path, so the two are interchangeable (DVAE_LSTM_SEQ=0 selects the other):
//   xg / gates [rows, T, D*4H] gate-interleaved (column 4u + g), h_all / c_all [rows, T, D*H],
//   da_all [rows, T, D*4H] natural torch order (g*H + u), whh_p [D][4H][H] interleaved rows, whh_n [D][4H][H] natural.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "act_types.cuh"
#include "host_common.h"
#include "ptx.cuh"
#include "umma_desc.cuh"

namespace dvae {

constexpr int kSeqH = 64;
constexpr int kSeqRows = 32;      // real rows per CTA (replicated x4 into the 128 MMA rows)
constexpr int kSeqThreads = 128;  // warp q: TMEM lane quarter q, hidden units [16q, 16q+16)
constexpr int kSeqTileA = 128 * 128;  // one k-block of the A operand: 128 rows x one 128-byte swizzle row

// 8 fp32 values -> 16-byte chunks of the storage dtype (bf16: 1 chunk, tf32: 2 chunks, rounded to nearest)
template <typename AT>
struct PackChunks;
template <>
struct PackChunks<__nv_bfloat16> {
  static constexpr int kPer8 = 1;
  static __device__ __forceinline__ void pack(const float* v, uint4* out) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(out);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  }
};
template <>
struct PackChunks<__half> {
  static constexpr int kPer8 = 1;
  static __device__ __forceinline__ void pack(const float* v, uint4* out) { *out = pack8_16<__half>(v); }
};
template <>
struct PackChunks<tf32_t> {
  static constexpr int kPer8 = 2;
  static __device__ __forceinline__ void pack(const float* v, uint4* out) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      out[j] = make_uint4(__float_as_uint(round_tf32(v[4 * j])), __float_as_uint(round_tf32(v[4 * j + 1])),
                          __float_as_uint(round_tf32(v[4 * j + 2])), __float_as_uint(round_tf32(v[4 * j + 3])));
  }
};

// Writes 16 consecutive K-elements [elem0, elem0 + 16) of this lane's row into the K-major, 128-byte-swizzled A operand
// (k-block kb = byte / 128, 16-byte chunk c of row r lives at chunk c ^ (r & 7)), once per row replica.
template <typename AT>
__device__ __forceinline__ void write_a_slice(uint32_t a_base, int lane, int elem0, const float* v16) {
  constexpr int EB = sizeof(AT);
  constexpr int P = PackChunks<AT>::kPer8;
  uint4 ch[2 * P];
  PackChunks<AT>::pack(v16, ch);
  PackChunks<AT>::pack(v16 + 8, ch + P);
  const int gc0 = elem0 * EB / 16;
#pragma unroll
  for (int j = 0; j < 2 * P; ++j) {
    const int gc = gc0 + j;
    const uint32_t off = static_cast<uint32_t>((gc >> 3) * kSeqTileA + lane * 128 + (((gc & 7) ^ (lane & 7)) << 4));
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) ptx::st_shared_v4(a_base + off + rep * (kSeqRows * 128), ch[j]);
  }
}

// Optional per-step phase stamps (SM clock) of CTA (0,0), thread 0: [step][8].  Off (nullptr) unless dvae_debug_seq_stamps()
// installed a buffer; one predictable branch per stamp.
__device__ __forceinline__ void seq_stamp(long long* stamps, int s, int slot) {
  if (stamps != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) stamps[s * 8 + slot] = clock64();
}

// =====================================================================================================================
// Forward: for s = 0..T-1 (time t = s, or T-1-s in the reverse direction)
//   a = xproj_t + h_prev . W_hh^T ; i,f,o = sigmoid, g = tanh ; c = f c + i g ; h = o tanh(c)
// saves the activated gates in place of xproj, c and h (what the backward pass and the next layer read).
// =====================================================================================================================
template <typename AT>
__global__ void __launch_bounds__(kSeqThreads, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmW, AT* __restrict__ xg, AT* __restrict__ h_all,
                    float* __restrict__ c_all, const int rows, const int T, const int D, long long* stamps) {
  constexpr int H = kSeqH;
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = H / BK;             // k-blocks: 1 (bf16) / 2 (tf32)
  constexpr int UMMA_K = 32 / EB;
  constexpr int TILE_W = 4 * H * 128;    // 256 weight rows x 128 B
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, 4 * H, false, false>();
  constexpr int TMEM_COLS = 4 * H;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t w_base = raw_addr + pad;
  const uint32_t a_base = w_base + KB * TILE_W;
  const uint32_t bar_base = a_base + KB * kSeqTileA;
  const uint32_t w_bar = bar_base, mma_bar = bar_base + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + KB * (TILE_W + kSeqTileA) + 16);

  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = blockIdx.y;
  const long r = static_cast<long>(blockIdx.x) * kSeqRows + lane;
  const bool row_ok = r < rows;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (q == 0) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 64 * q;   // my lane quarter, my 64 columns

  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(w_bar, KB * TILE_W);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) ptx::tma_load_3d(w_base + kb * TILE_W, &tmW, w_bar, kb * BK, 0, d);
  }

  const long ldx = static_cast<long>(T) * D * 4 * H, ldh = static_cast<long>(T) * D * H;
  AT* xrow = xg + r * ldx + d * 4 * H + 64 * q;        // + t * D*4H
  float* crow = c_all + r * ldh + d * H + 16 * q;      // + t * D*H
  AT* hrow = h_all + r * ldh + d * H + 16 * q;
  auto t_of = [&](int s) { return d == 0 ? s : T - 1 - s; };

  typename Act8<AT>::raw_t xr[8], xn[8];
  if (row_ok) {
    const AT* xp = xrow + static_cast<long>(t_of(0)) * D * 4 * H;
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[j] = Act8<AT>::load_raw(xp + 8 * j);
  }
  float c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.f;

  for (int s = 0; s < T; ++s) {
    const int t = t_of(s);
    if (row_ok && s + 1 < T) {   // next step's x-projection: in flight across this step's MMA wait
      const AT* xp = xrow + static_cast<long>(t_of(s + 1)) * D * 4 * H;
#pragma unroll
      for (int j = 0; j < 8; ++j) xn[j] = Act8<AT>::load_raw(xp + 8 * j);
    }
    float hn[16];
#pragma unroll
    for (int k = 0; k < 2; ++k) {   // 32 accumulator columns = 8 hidden units x (i, f, g, o)
      float a[32];
      if (s > 0) {
        if (k == 0) {
          seq_stamp(stamps, s, 0);
          ptx::mbar_wait(mma_bar, static_cast<uint32_t>(s - 1) & 1u);
          ptx::tc_fence_after();
          seq_stamp(stamps, s, 1);
        }
        ptx::tmem_ld_x32(taddr + 32 * k, a);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = 0.f;
      }
      float x[32];
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<AT>::unpack(xr[4 * k + j], x + 8 * j);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
      }
      float cn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ig = GateMath<AT>::sig(a[4 * i] + x[4 * i]), fg = GateMath<AT>::sig(a[4 * i + 1] + x[4 * i + 1]);
        const float gg = GateMath<AT>::tnh(a[4 * i + 2] + x[4 * i + 2]), og = GateMath<AT>::sig(a[4 * i + 3] + x[4 * i + 3]);
        a[4 * i] = ig; a[4 * i + 1] = fg; a[4 * i + 2] = gg; a[4 * i + 3] = og;
        cn[i] = fg * c[8 * k + i] + ig * gg;
        c[8 * k + i] = cn[i];
        hn[8 * k + i] = og * GateMath<AT>::tnh(cn[i]);
      }
      if (row_ok) {
        AT* gp = xrow + static_cast<long>(t) * D * 4 * H + 32 * k;
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<AT>::store(gp + 8 * j, a + 8 * j);
        Act8<float>::store(crow + static_cast<long>(t) * D * H + 8 * k, cn);
        Act8<AT>::store(hrow + static_cast<long>(t) * D * H + 8 * k, hn + 8 * k);
      }
    }
    if (s + 1 < T) {
      seq_stamp(stamps, s, 2);
      write_a_slice<AT>(a_base, lane, 16 * q, hn);
      seq_stamp(stamps, s, 3);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      seq_stamp(stamps, s, 4);
      __syncthreads();
      seq_stamp(stamps, s, 5);
      if (threadIdx.x == 0) {
        if (s == 0) ptx::mbar_wait(w_bar, 0);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t adesc = smem_desc(a_base + kb * kSeqTileA, 16, 1024, 2);
          const uint64_t bdesc = smem_desc(w_base + kb * TILE_W, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma<EB>(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(mma_bar);
        seq_stamp(stamps, s, 6);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[j] = xn[j];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (q == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================================
// Backward through time: for s = 0..T-1 (time t = T-1-s, or s in the reverse direction)
//   dh = dh_out_t + da_prev . W_hh   (da_prev: the gate gradients of the step processed just before, kept in shared memory)
//   cell backward (same arithmetic as lstm_cell_bwd_kernel in ops_gemm.cu) -> da_t (natural gate order), dc carry.
// =====================================================================================================================
template <typename AT>
__global__ void __launch_bounds__(kSeqThreads, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const AT* __restrict__ dh_all, const AT* __restrict__ gates,
                    const float* __restrict__ c_all, AT* __restrict__ da_all, const int rows, const int T, const int D,
                    long long* stamps) {
  constexpr int H = kSeqH;
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = 4 * H / BK;                 // k-blocks of the reduction over the 4H gate columns: 4 / 8
  constexpr int UMMA_K = 32 / EB;
  constexpr int B_BOXES = H * EB / 128;          // W_hh is read MN-major: 128-byte-wide boxes along the hidden dim
  constexpr int MN_BOX_BYTES = BK * 128;
  constexpr int STAGE_B = B_BOXES * MN_BOX_BYTES;
  constexpr uint32_t MN_LAYOUT = (EB == 4) ? 1u : 2u;
  constexpr uint32_t MN_SBO = (EB == 4) ? 512u : 1024u;
  constexpr uint32_t ADV_B = (UMMA_K * 128) >> 4;
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, H, false, true>();
  constexpr int TMEM_COLS = H;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t w_base = raw_addr + pad;
  const uint32_t a_base = w_base + KB * STAGE_B;
  const uint32_t bar_base = a_base + KB * kSeqTileA;
  const uint32_t w_bar = bar_base, mma_bar = bar_base + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + KB * (STAGE_B + kSeqTileA) + 16);

  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = blockIdx.y;
  const long r = static_cast<long>(blockIdx.x) * kSeqRows + lane;
  const bool row_ok = r < rows;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (q == 0) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 16 * q;   // my units = my 16 columns

  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(w_bar, KB * STAGE_B);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int i = 0; i < B_BOXES; ++i)
        ptx::tma_load_3d(w_base + kb * STAGE_B + i * MN_BOX_BYTES, &tmW, w_bar, i * BK, kb * BK, d);
  }

  const long ldx = static_cast<long>(T) * D * 4 * H, ldh = static_cast<long>(T) * D * H;
  const AT* dhrow = dh_all + r * ldh + d * H + 16 * q;          // + t * D*H
  const AT* grow = gates + r * ldx + d * 4 * H + 64 * q;        // + t * D*4H   (interleaved: my 16 units = 64 columns)
  const float* crow = c_all + r * ldh + d * H + 16 * q;
  AT* darow = da_all + r * ldx + d * 4 * H + 16 * q;            // + t * D*4H + g*H
  auto t_of = [&](int s) { return d == 0 ? T - 1 - s : s; };

  typename Act8<AT>::raw_t gr[8], gn[8], dhr[2], dhn[2];
  typename Act8<float>::raw_t cb_n[2];
  float ca[16], cb[16], dc[16];   // c at this step's time, c at the next processed step's time (= c_prev of the cell), carry
#pragma unroll
  for (int i = 0; i < 16; ++i) ca[i] = cb[i] = dc[i] = 0.f;
  if (row_ok) {
    const long t0 = t_of(0);
#pragma unroll
    for (int j = 0; j < 8; ++j) gr[j] = Act8<AT>::load_raw(grow + t0 * D * 4 * H + 8 * j);
#pragma unroll
    for (int j = 0; j < 2; ++j) dhr[j] = Act8<AT>::load_raw(dhrow + t0 * D * H + 8 * j);
    Act8<float>::load(crow + t0 * D * H, ca);
    Act8<float>::load(crow + t0 * D * H + 8, ca + 8);
    if (T > 1) {
      const long t1 = t_of(1);
      Act8<float>::load(crow + t1 * D * H, cb);
      Act8<float>::load(crow + t1 * D * H + 8, cb + 8);
    }
  }

  for (int s = 0; s < T; ++s) {
    const int t = t_of(s);
    if (row_ok && s + 1 < T) {
      const long t1 = t_of(s + 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) gn[j] = Act8<AT>::load_raw(grow + t1 * D * 4 * H + 8 * j);
#pragma unroll
      for (int j = 0; j < 2; ++j) dhn[j] = Act8<AT>::load_raw(dhrow + t1 * D * H + 8 * j);
      if (s + 2 < T) {
        const long t2 = t_of(s + 2);
        cb_n[0] = Act8<float>::load_raw(crow + t2 * D * H);
        cb_n[1] = Act8<float>::load_raw(crow + t2 * D * H + 8);
      }
    }
    float rec[16];
    if (s > 0) {
      seq_stamp(stamps, s, 0);
      ptx::mbar_wait(mma_bar, static_cast<uint32_t>(s - 1) & 1u);
      ptx::tc_fence_after();
      seq_stamp(stamps, s, 1);
      ptx::tmem_ld_x8(taddr, rec);
      ptx::tmem_ld_x8(taddr + 8, rec + 8);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) rec[i] = 0.f;
    }
    float dai[16], daf[16], dag[16], dao[16];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float g4[32], dh[8];
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<AT>::unpack(gr[4 * k + j], g4 + 8 * j);
        Act8<AT>::unpack(dhr[k], dh);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) g4[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dh[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int u = 8 * k + i;
        const float ig = g4[4 * i], fg = g4[4 * i + 1], gg = g4[4 * i + 2], og = g4[4 * i + 3];
        const float tc = GateMath<AT>::tnh(ca[u]);
        const float dht = dh[i] + rec[u];
        const float dct = dc[u] + dht * og * (1.f - tc * tc);
        dao[u] = dht * tc * og * (1.f - og);
        dai[u] = dct * gg * ig * (1.f - ig);
        dag[u] = dct * ig * (1.f - gg * gg);
        daf[u] = dct * ((s + 1 < T) ? cb[u] : 0.f) * fg * (1.f - fg);
        dc[u] = dct * fg;
      }
    }
    if (row_ok) {
      AT* o = darow + static_cast<long>(t) * D * 4 * H;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        Act8<AT>::store(o + 8 * k, dai + 8 * k);
        Act8<AT>::store(o + H + 8 * k, daf + 8 * k);
        Act8<AT>::store(o + 2 * H + 8 * k, dag + 8 * k);
        Act8<AT>::store(o + 3 * H + 8 * k, dao + 8 * k);
      }
    }
    if (s + 1 < T) {
      seq_stamp(stamps, s, 2);
      write_a_slice<AT>(a_base, lane, 0 * H + 16 * q, dai);
      write_a_slice<AT>(a_base, lane, 1 * H + 16 * q, daf);
      write_a_slice<AT>(a_base, lane, 2 * H + 16 * q, dag);
      write_a_slice<AT>(a_base, lane, 3 * H + 16 * q, dao);
      seq_stamp(stamps, s, 3);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      seq_stamp(stamps, s, 4);
      __syncthreads();
      seq_stamp(stamps, s, 5);
      if (threadIdx.x == 0) {
        if (s == 0) ptx::mbar_wait(w_bar, 0);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t adesc = smem_desc(a_base + kb * kSeqTileA, 16, 1024, 2);
          const uint64_t bdesc = smem_desc(w_base + kb * STAGE_B, MN_BOX_BYTES, MN_SBO, MN_LAYOUT);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma<EB>(tmem_base, adesc + 2 * k, bdesc + k * ADV_B, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(mma_bar);
        seq_stamp(stamps, s, 6);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) gr[j] = gn[j];
#pragma unroll
      for (int j = 0; j < 2; ++j) dhr[j] = dhn[j];
#pragma unroll
      for (int i = 0; i < 16; ++i) ca[i] = cb[i];
      if (row_ok && s + 2 < T) {
        Act8<float>::unpack(cb_n[0], cb);
        Act8<float>::unpack(cb_n[1], cb + 8);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (q == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================================
// TMA-staged variants.  Phase stamps of the kernels above (scripts/seq_lstm_stamps.py) show ~4000 of ~5300 clocks per
// step between "accumulator ready" and "cell done": with one row per lane every 16-byte global access of a warp touches
// 32 different 128-byte lines, and the LSU replays them one line at a time.  Here all per-step traffic moves as TMA boxes
// of 32 rows x 128 B (swizzle-128B, so a row-per-lane ld/st.shared.v4 is conflict free):
//   warp 4, lane 0  = DMA + MMA thread: prefetches the inputs of the next steps, and after the compute warps have arrived
//                     (named barrier 1) issues the next tcgen05.mma first and the TMA stores of the finished step second;
//   warps 0..3      = the cell arithmetic, between shared memory, TMEM and registers only.
// The recurrent operand tile A doubles as the store source of h (forward) / da (backward): replica 0 of each k-block is
// exactly a [32 rows x 128 B] swizzled box.  `a_free` tells the compute warps the store has finished reading it.
// =====================================================================================================================
constexpr int kSeqThreadsTma = 160;
constexpr int kSeqBoxBytes = kSeqRows * 128;
constexpr int kSeqBufs = 3;

// this lane's 16-byte chunk gc (8 chunks per 128-byte box) of a [boxes][32 rows][128 B] swizzled tile
__device__ __forceinline__ uint32_t box_chunk_addr(uint32_t base, int lane, int gc) {
  return base + static_cast<uint32_t>((gc >> 3) * kSeqBoxBytes + lane * 128 + (((gc & 7) ^ (lane & 7)) << 4));
}
template <typename ET> struct ChunkCvt;   // one 16-byte chunk <-> fp32 values
template <> struct ChunkCvt<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float* v) { Act8<__nv_bfloat16>::unpack(u, v); }
  static __device__ __forceinline__ uint4 pack(const float* v) {
    uint4 u;
    PackChunks<__nv_bfloat16>::pack(v, &u);
    return u;
  }
};
template <> struct ChunkCvt<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float* v) { Act8<__half>::unpack(u, v); }
  static __device__ __forceinline__ uint4 pack(const float* v) { return pack8_16<__half>(v); }
};
template <> struct ChunkCvt<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct ChunkCvt<tf32_t> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float* v) { ChunkCvt<float>::unpack(u, v); }
  static __device__ __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(round_tf32(v[0])), __float_as_uint(round_tf32(v[1])), __float_as_uint(round_tf32(v[2])),
                      __float_as_uint(round_tf32(v[3])));
  }
};
// NELEM consecutive elements of this lane's row, starting at element elem0, from / to a swizzled box tile
template <typename ET, int NELEM>
__device__ __forceinline__ void lds_slice(uint32_t base, int lane, int elem0, float* out) {
  constexpr int N = ChunkCvt<ET>::N;
#pragma unroll
  for (int j = 0; j < NELEM / N; ++j) ChunkCvt<ET>::unpack(ptx::ld_shared_v4(box_chunk_addr(base, lane, elem0 / N + j)), out + j * N);
}
template <typename ET, int NELEM>
__device__ __forceinline__ void sts_slice(uint32_t base, int lane, int elem0, const float* in) {
  constexpr int N = ChunkCvt<ET>::N;
#pragma unroll
  for (int j = 0; j < NELEM / N; ++j) ptx::st_shared_v4(box_chunk_addr(base, lane, elem0 / N + j), ChunkCvt<ET>::pack(in + j * N));
}

template <typename AT>
constexpr int seq_fwd_tma_smem() {
  constexpr int EB = sizeof(AT);
  constexpr int KB = kSeqH * EB / 128;
  return KB * (4 * kSeqH * 128 + kSeqTileA) + kSeqBufs * ((4 * kSeqH * EB / 128) + (kSeqH * 4 / 128)) * kSeqBoxBytes + 1024 + 128;
}

template <typename AT>
__global__ void __launch_bounds__(kSeqThreadsTma, 1)
lstm_seq_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmH, const int T,
                        long long* stamps) {
  constexpr int H = kSeqH;
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = H / BK;
  constexpr int UMMA_K = 32 / EB;
  constexpr int TILE_W = 4 * H * 128;
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, 4 * H, false, false>();
  constexpr int TMEM_COLS = 4 * H;
  constexpr int X_BOXES = 4 * H * EB / 128, X_BYTES = X_BOXES * kSeqBoxBytes;
  constexpr int C_BOXES = H * 4 / 128, C_BYTES = C_BOXES * kSeqBoxBytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t w_base = raw_addr + pad;
  const uint32_t a_base = w_base + KB * TILE_W;
  const uint32_t x_base = a_base + KB * kSeqTileA;
  const uint32_t c_base = x_base + kSeqBufs * X_BYTES;
  const uint32_t bar_base = c_base + kSeqBufs * C_BYTES;
  const uint32_t w_bar = bar_base, mma_bar = bar_base + 8, a_free = bar_base + 16;
  auto x_full = [&](int b) { return bar_base + 24u + 8u * b; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar_base - w_base) + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = blockIdx.y;
  const int row0 = blockIdx.x * kSeqRows;
  auto t_of = [&](int s) { return d == 0 ? s : T - 1 - s; };

  if (threadIdx.x == 0) {
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::mbar_init(a_free, 1);
    for (int b = 0; b < kSeqBufs; ++b) ptx::mbar_init(x_full(b), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ DMA + MMA warp
    auto load_x = [&](int s) {
      const int buf = s % kSeqBufs;
      ptx::mbar_expect_tx(x_full(buf), X_BYTES);
#pragma unroll
      for (int b = 0; b < X_BOXES; ++b)
        ptx::tma_load_3d(x_base + buf * X_BYTES + b * kSeqBoxBytes, &tmX, x_full(buf), d * 4 * H + b * BK, t_of(s), row0);
    };
    if (lane == 0) {
      ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmC); ptx::prefetch_tmap(&tmH);
      load_x(0);
      ptx::mbar_expect_tx(w_bar, KB * TILE_W);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) ptx::tma_load_3d(w_base + kb * TILE_W, &tmW, w_bar, kb * BK, 0, d);
      if (T > 1) load_x(1);
    }
    for (int s = 0; s < T; ++s) {
      ptx::bar_sync(1, kSeqThreadsTma);   // every compute thread has written step s's h / gates / c
      if (lane == 0) {
        ptx::tc_fence_after();
        const int t = t_of(s), buf = s % kSeqBufs;
        if (s + 1 < T) {
          if (s == 0) ptx::mbar_wait(w_bar, 0);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t adesc = smem_desc(a_base + kb * kSeqTileA, 16, 1024, 2);
            const uint64_t bdesc = smem_desc(w_base + kb * TILE_W, 16, 1024, 2);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma<EB>(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(mma_bar);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) ptx::tma_store_3d(&tmH, a_base + kb * kSeqTileA, d * H + kb * BK, t, row0);
        ptx::bulk_commit();
#pragma unroll
        for (int b = 0; b < X_BOXES; ++b)
          ptx::tma_store_3d(&tmX, x_base + buf * X_BYTES + b * kSeqBoxBytes, d * 4 * H + b * BK, t, row0);
#pragma unroll
        for (int b = 0; b < C_BOXES; ++b)
          ptx::tma_store_3d(&tmC, c_base + buf * C_BYTES + b * kSeqBoxBytes, d * H + b * 32, t, row0);
        ptx::bulk_commit();
        ptx::bulk_wait_read<1>();   // everything but the group just committed has been read: h(s), gates/c(s-1)
        ptx::mbar_arrive(a_free);
        if (s + 2 < T) load_x(s + 2);   // into the buffer of step s-1
      }
      __syncwarp();
    }
    if (lane == 0) ptx::bulk_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ compute warps: warp q = units [16q, 16q+16)
    const int q = warp;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 64 * q;
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int buf = s % kSeqBufs;
      const uint32_t xb = x_base + buf * X_BYTES, cb = c_base + buf * C_BYTES;
      ptx::mbar_wait(x_full(buf), static_cast<uint32_t>(s / kSeqBufs) & 1u);
      float x[64];
      lds_slice<AT, 64>(xb, lane, 64 * q, x);
      seq_stamp(stamps, s, 0);
      if (s > 0) {
        ptx::mbar_wait(mma_bar, static_cast<uint32_t>(s - 1) & 1u);
        ptx::tc_fence_after();
      }
      seq_stamp(stamps, s, 1);
      float hn[16];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float a[32];
        if (s > 0) {
          ptx::tmem_ld_x32(taddr + 32 * k, a);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = 0.f;
        }
        float cn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* xx = x + 32 * k + 4 * i;
          const float ig = GateMath<AT>::sig(a[4 * i] + xx[0]), fg = GateMath<AT>::sig(a[4 * i + 1] + xx[1]);
          const float gg = GateMath<AT>::tnh(a[4 * i + 2] + xx[2]), og = GateMath<AT>::sig(a[4 * i + 3] + xx[3]);
          a[4 * i] = ig; a[4 * i + 1] = fg; a[4 * i + 2] = gg; a[4 * i + 3] = og;
          cn[i] = fg * c[8 * k + i] + ig * gg;
          c[8 * k + i] = cn[i];
          hn[8 * k + i] = og * GateMath<AT>::tnh(cn[i]);
        }
        sts_slice<AT, 32>(xb, lane, 64 * q + 32 * k, a);      // activated gates, in place of the x-projection
        sts_slice<float, 8>(cb, lane, 16 * q + 8 * k, cn);
      }
      seq_stamp(stamps, s, 2);
      if (s > 0) ptx::mbar_wait(a_free, static_cast<uint32_t>(s - 1) & 1u);   // TMA store of h(s-1) has read the A tile
      write_a_slice<AT>(a_base, lane, 16 * q, hn);
      seq_stamp(stamps, s, 3);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      seq_stamp(stamps, s, 4);
      ptx::bar_arrive(1, kSeqThreadsTma);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <typename AT>
constexpr int seq_bwd_tma_smem() {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = 4 * kSeqH / BK;
  return KB * ((kSeqH * EB / 128) * BK * 128 + kSeqTileA) +
         kSeqBufs * ((4 * kSeqH * EB / 128) + (kSeqH * EB / 128) + (kSeqH * 4 / 128)) * kSeqBoxBytes + 1024 + 128;
}

template <typename AT>
__global__ void __launch_bounds__(kSeqThreadsTma, 1)
lstm_seq_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG,
                        const __grid_constant__ CUtensorMap tmDH, const __grid_constant__ CUtensorMap tmC,
                        const __grid_constant__ CUtensorMap tmDA, const float* __restrict__ c_all, const int rows, const int T,
                        const int D, long long* stamps) {
  constexpr int H = kSeqH;
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = 4 * H / BK;
  constexpr int UMMA_K = 32 / EB;
  constexpr int B_BOXES = H * EB / 128;
  constexpr int MN_BOX_BYTES = BK * 128;
  constexpr int STAGE_B = B_BOXES * MN_BOX_BYTES;
  constexpr uint32_t MN_LAYOUT = (EB == 4) ? 1u : 2u;
  constexpr uint32_t MN_SBO = (EB == 4) ? 512u : 1024u;
  constexpr uint32_t ADV_B = (UMMA_K * 128) >> 4;
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, H, false, true>();
  constexpr int TMEM_COLS = H;
  constexpr int G_BOXES = 4 * H * EB / 128, G_BYTES = G_BOXES * kSeqBoxBytes;
  constexpr int DH_BOXES = H * EB / 128, DH_BYTES = DH_BOXES * kSeqBoxBytes;
  constexpr int C_BOXES = H * 4 / 128, C_BYTES = C_BOXES * kSeqBoxBytes;
  constexpr int IN_BYTES = G_BYTES + DH_BYTES + C_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t w_base = raw_addr + pad;
  const uint32_t a_base = w_base + KB * STAGE_B;
  const uint32_t in_base = a_base + KB * kSeqTileA;
  const uint32_t bar_base = in_base + kSeqBufs * IN_BYTES;
  const uint32_t w_bar = bar_base, mma_bar = bar_base + 8, a_free = bar_base + 16;
  auto in_full = [&](int b) { return bar_base + 24u + 8u * b; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar_base - w_base) + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = blockIdx.y;
  const int row0 = blockIdx.x * kSeqRows;
  auto t_of = [&](int s) { return d == 0 ? T - 1 - s : s; };

  if (threadIdx.x == 0) {
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::mbar_init(a_free, 1);
    for (int b = 0; b < kSeqBufs; ++b) ptx::mbar_init(in_full(b), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ DMA + MMA warp
    auto load_in = [&](int s) {   // saved gates and dh of step s, c of the step processed after it (the cell's c_prev)
      const int buf = s % kSeqBufs;
      const uint32_t base = in_base + buf * IN_BYTES;
      const bool has_c = s + 1 < T;
      ptx::mbar_expect_tx(in_full(buf), G_BYTES + DH_BYTES + (has_c ? C_BYTES : 0));
#pragma unroll
      for (int b = 0; b < G_BOXES; ++b)
        ptx::tma_load_3d(base + b * kSeqBoxBytes, &tmG, in_full(buf), d * 4 * H + b * BK, t_of(s), row0);
#pragma unroll
      for (int b = 0; b < DH_BOXES; ++b)
        ptx::tma_load_3d(base + G_BYTES + b * kSeqBoxBytes, &tmDH, in_full(buf), d * H + b * BK, t_of(s), row0);
      if (has_c) {
#pragma unroll
        for (int b = 0; b < C_BOXES; ++b)
          ptx::tma_load_3d(base + G_BYTES + DH_BYTES + b * kSeqBoxBytes, &tmC, in_full(buf), d * H + b * 32, t_of(s + 1), row0);
      }
    };
    if (lane == 0) {
      ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmG); ptx::prefetch_tmap(&tmDH); ptx::prefetch_tmap(&tmC);
      ptx::prefetch_tmap(&tmDA);
      load_in(0);
      ptx::mbar_expect_tx(w_bar, KB * STAGE_B);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int i = 0; i < B_BOXES; ++i)
          ptx::tma_load_3d(w_base + kb * STAGE_B + i * MN_BOX_BYTES, &tmW, w_bar, i * BK, kb * BK, d);
      for (int s = 1; s < kSeqBufs && s < T; ++s) load_in(s);
    }
    for (int s = 0; s < T; ++s) {
      ptx::bar_sync(1, kSeqThreadsTma);   // compute warps have written da_t into the A tile and are done with IN[s % 3]
      if (lane == 0) {
        ptx::tc_fence_after();
        const int t = t_of(s);
        if (s + 1 < T) {
          if (s == 0) ptx::mbar_wait(w_bar, 0);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t adesc = smem_desc(a_base + kb * kSeqTileA, 16, 1024, 2);
            const uint64_t bdesc = smem_desc(w_base + kb * STAGE_B, MN_BOX_BYTES, MN_SBO, MN_LAYOUT);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma<EB>(tmem_base, adesc + 2 * k, bdesc + k * ADV_B, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(mma_bar);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) ptx::tma_store_3d(&tmDA, a_base + kb * kSeqTileA, d * 4 * H + kb * BK, t, row0);
        ptx::bulk_commit();
        if (s + kSeqBufs < T) load_in(s + kSeqBufs);
        ptx::bulk_wait_read<0>();
        ptx::mbar_arrive(a_free);
      }
      __syncwarp();
    }
    if (lane == 0) ptx::bulk_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ compute warps
    const int q = warp;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 16 * q;
    float ca[16], dc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ca[i] = dc[i] = 0.f;
    {
      const long r = static_cast<long>(row0) + lane;   // c at the first processed step: once, straight from global
      if (r < rows) {
        const float* cp = c_all + (r * T + t_of(0)) * D * H + d * H + 16 * q;
        Act8<float>::load(cp, ca);
        Act8<float>::load(cp + 8, ca + 8);
      }
    }
    for (int s = 0; s < T; ++s) {
      const int buf = s % kSeqBufs;
      const uint32_t gb = in_base + buf * IN_BYTES, dhb = gb + G_BYTES, cbb = dhb + DH_BYTES;
      ptx::mbar_wait(in_full(buf), static_cast<uint32_t>(s / kSeqBufs) & 1u);
      float g4[64], dh[16], cb[16];
      lds_slice<AT, 64>(gb, lane, 64 * q, g4);
      lds_slice<AT, 16>(dhb, lane, 16 * q, dh);
      if (s + 1 < T) {
        lds_slice<float, 16>(cbb, lane, 16 * q, cb);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) cb[i] = 0.f;
      }
      seq_stamp(stamps, s, 0);
      float rec[16];
      if (s > 0) {
        ptx::mbar_wait(mma_bar, static_cast<uint32_t>(s - 1) & 1u);
        ptx::tc_fence_after();
        seq_stamp(stamps, s, 1);
        ptx::tmem_ld_x8(taddr, rec);
        ptx::tmem_ld_x8(taddr + 8, rec + 8);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) rec[i] = 0.f;
      }
      float dai[16], daf[16], dag[16], dao[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ig = g4[4 * u], fg = g4[4 * u + 1], gg = g4[4 * u + 2], og = g4[4 * u + 3];
        const float tc = GateMath<AT>::tnh(ca[u]);
        const float dht = dh[u] + rec[u];
        const float dct = dc[u] + dht * og * (1.f - tc * tc);
        dao[u] = dht * tc * og * (1.f - og);
        dai[u] = dct * gg * ig * (1.f - ig);
        dag[u] = dct * ig * (1.f - gg * gg);
        daf[u] = dct * cb[u] * fg * (1.f - fg);
        dc[u] = dct * fg;
        ca[u] = cb[u];
      }
      seq_stamp(stamps, s, 2);
      if (s > 0) ptx::mbar_wait(a_free, static_cast<uint32_t>(s - 1) & 1u);   // TMA store of da(s-1) has read the A tile
      write_a_slice<AT>(a_base, lane, 0 * H + 16 * q, dai);
      write_a_slice<AT>(a_base, lane, 1 * H + 16 * q, daf);
      write_a_slice<AT>(a_base, lane, 2 * H + 16 * q, dag);
      write_a_slice<AT>(a_base, lane, 3 * H + 16 * q, dao);
      seq_stamp(stamps, s, 3);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      seq_stamp(stamps, s, 4);
      ptx::bar_arrive(1, kSeqThreadsTma);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
static long long* g_seq_stamps = nullptr;   // debug: device buffer [T][8] of SM-clock stamps (dvae_debug_seq_stamps)
void lstm_seq_set_stamps(long long* buf) { g_seq_stamps = buf; }

static bool seq_env_enabled() {
  static const int v = [] {
    const char* e = getenv("DVAE_LSTM_SEQ");
    return (e && *e) ? atoi(e) : 1;
  }();
  return v != 0;
}

// DVAE_LSTM_SEQ_IO=direct keeps the row-per-lane global accesses (A/B comparison); default: TMA-staged
static bool seq_tma_io() {
  static const bool v = [] {
    const char* e = getenv("DVAE_LSTM_SEQ_IO");
    return !(e && e[0] == 'd');
  }();
  return v;
}

bool lstm_seq_supported(int H, int T) { return seq_env_enabled() && H == kSeqH && T >= 1; }

template <typename AT>
static int lstm_seq_fwd_t(AT* xg, const AT* whh_p, AT* h_all, float* c_all, int rows, int T, int H, int D, cudaStream_t st) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = kSeqH / BK;
  DVAE_REQUIRE(H == kSeqH && (D == 1 || D == 2), "sequence-resident LSTM: H must be 64, D in {1,2}");
  if (rows <= 0) return 0;
  CUtensorMap tw;
  if (int e = encode_map3(&tw, whh_p, EB, H, 4 * H, D, (uint64_t)H * EB, (uint64_t)4 * H * H * EB, BK, 4 * H, 1)) return e;
  if (seq_tma_io()) {
    CUtensorMap tx, tc, th;
    if (int e = encode_map3(&tx, xg, EB, (uint64_t)D * 4 * H, T, rows, (uint64_t)D * 4 * H * EB, (uint64_t)T * D * 4 * H * EB, BK, 1,
                            kSeqRows))
      return e;
    if (int e = encode_map3(&tc, c_all, 4, (uint64_t)D * H, T, rows, (uint64_t)D * H * 4, (uint64_t)T * D * H * 4, 32, 1, kSeqRows))
      return e;
    if (int e = encode_map3(&th, h_all, EB, (uint64_t)D * H, T, rows, (uint64_t)D * H * EB, (uint64_t)T * D * H * EB, BK, 1, kSeqRows))
      return e;
    constexpr int smem_t = seq_fwd_tma_smem<AT>();
    static_assert(smem_t <= 227 * 1024, "sequence-resident LSTM forward: shared memory");
    auto kern_t = lstm_seq_fwd_tma_kernel<AT>;
    static bool configured_t = false;
    if (!configured_t) {
      DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern_t, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t));
      configured_t = true;
    }
    dim3 grid_t(ceil_div(rows, kSeqRows), D);
    kern_t<<<grid_t, kSeqThreadsTma, smem_t, st>>>(tw, tx, tc, th, T, g_seq_stamps);
    DVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  constexpr int smem = KB * (4 * kSeqH * 128 + kSeqTileA) + 1024 + 64;
  auto kern = lstm_seq_fwd_kernel<AT>;
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(rows, kSeqRows), D);
  kern<<<grid, kSeqThreads, smem, st>>>(tw, xg, h_all, c_all, rows, T, D, g_seq_stamps);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename AT>
static int lstm_seq_bwd_t(const AT* dh_all, const AT* gates, const float* c_all, const AT* whh_n, AT* da_all, int rows, int T,
                          int H, int D, cudaStream_t st) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int KB = 4 * kSeqH / BK;
  DVAE_REQUIRE(H == kSeqH && (D == 1 || D == 2), "sequence-resident LSTM: H must be 64, D in {1,2}");
  if (rows <= 0) return 0;
  CUtensorMap tw;
  if (int e = encode_map3(&tw, whh_n, EB, H, 4 * H, D, (uint64_t)H * EB, (uint64_t)4 * H * H * EB, BK, BK, 1, true)) return e;
  if constexpr (seq_bwd_tma_smem<AT>() <= 227 * 1024) {   // bf16; the tf32 tiles (A alone is 128 KB) leave no room for staging
    if (seq_tma_io()) {
      CUtensorMap tg, tdh, tc, tda;
      const uint64_t s1x = (uint64_t)D * 4 * H * EB, s2x = (uint64_t)T * D * 4 * H * EB;
      const uint64_t s1h = (uint64_t)D * H * EB, s2h = (uint64_t)T * D * H * EB;
      if (int e = encode_map3(&tg, gates, EB, (uint64_t)D * 4 * H, T, rows, s1x, s2x, BK, 1, kSeqRows)) return e;
      if (int e = encode_map3(&tdh, dh_all, EB, (uint64_t)D * H, T, rows, s1h, s2h, BK, 1, kSeqRows)) return e;
      if (int e = encode_map3(&tc, c_all, 4, (uint64_t)D * H, T, rows, (uint64_t)D * H * 4, (uint64_t)T * D * H * 4, 32, 1, kSeqRows))
        return e;
      if (int e = encode_map3(&tda, da_all, EB, (uint64_t)D * 4 * H, T, rows, s1x, s2x, BK, 1, kSeqRows)) return e;
      constexpr int smem_t = seq_bwd_tma_smem<AT>();
      auto kern_t = lstm_seq_bwd_tma_kernel<AT>;
      static bool configured_t = false;
      if (!configured_t) {
        DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern_t, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t));
        configured_t = true;
      }
      dim3 grid_t(ceil_div(rows, kSeqRows), D);
      kern_t<<<grid_t, kSeqThreadsTma, smem_t, st>>>(tw, tg, tdh, tc, tda, c_all, rows, T, D, g_seq_stamps);
      DVAE_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  constexpr int smem = KB * ((kSeqH * EB / 128) * BK * 128 + kSeqTileA) + 1024 + 64;
  auto kern = lstm_seq_bwd_kernel<AT>;
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(rows, kSeqRows), D);
  kern<<<grid, kSeqThreads, smem, st>>>(tw, dh_all, gates, c_all, da_all, rows, T, D, g_seq_stamps);
  DVAE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int lstm_seq_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, int D,
                 cudaStream_t st) {
  using bf16 = __nv_bfloat16;
  if (dtype == kBF16) return lstm_seq_fwd_t<bf16>((bf16*)xg, (const bf16*)whh_p, (bf16*)h_all, c_all, rows, T, H, D, st);
  if (dtype == kF16) return lstm_seq_fwd_t<__half>((__half*)xg, (const __half*)whh_p, (__half*)h_all, c_all, rows, T, H, D, st);
  if (dtype == kTF32) return lstm_seq_fwd_t<tf32_t>((tf32_t*)xg, (const tf32_t*)whh_p, (tf32_t*)h_all, c_all, rows, T, H, D, st);
  set_last_error("unknown dtype tag");
  return 1;
}

int lstm_seq_bwd(int dtype, const void* dh_all, const void* gates, const float* c_all, const void* whh_n, void* da_all,
                 int rows, int T, int H, int D, cudaStream_t st) {
  using bf16 = __nv_bfloat16;
  if (dtype == kBF16)
    return lstm_seq_bwd_t<bf16>((const bf16*)dh_all, (const bf16*)gates, c_all, (const bf16*)whh_n, (bf16*)da_all, rows, T, H, D, st);
  if (dtype == kF16)
    return lstm_seq_bwd_t<__half>((const __half*)dh_all, (const __half*)gates, c_all, (const __half*)whh_n, (__half*)da_all, rows, T, H, D, st);
  if (dtype == kTF32)
    return lstm_seq_bwd_t<tf32_t>((const tf32_t*)dh_all, (const tf32_t*)gates, c_all, (const tf32_t*)whh_n, (tf32_t*)da_all, rows,
                                  T, H, D, st);
  set_last_error("unknown dtype tag");
  return 1;
}

}  // namespace dvae
