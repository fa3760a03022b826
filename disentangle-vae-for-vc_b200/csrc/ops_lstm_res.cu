// Weight-stationary, time-resident LSTM forward recurrence for the large hidden sizes (H = 512: dec_lstm1, H = 1024:
// dec_lstm2; reference model/disentangled_vae.py:172,193,238,246 and autovc_replicate/proposed_autovc.py:73-75).
//
// The step-per-launch path (ops_gemm.cu) runs one GEMM launch per time step; a step is a strictly serial chain
//   launch + set-up -> pipeline fill -> MMA main loop -> cell epilogue -> stores -> next launch
// (profiles/r01_phase_timing_v2.txt: 8.5 us main loop + 5 us epilogue + 2-3 us launch at H = 1024), and every step
// streams the whole of W_hh (8 MB at H = 1024) from L2 once per 128-row tile.  This file replaces the chain by ONE launch
// per layer:
//
//   * CTA pairs (tcgen05 cta_group::2, M = 256).  A pair owns one SLICE of W_hh -- BN gate columns = BN / 4 hidden units x
//     (i,f,g,o) -- for the whole sequence and keeps it in shared memory (each CTA of the pair holds half: 128 KB at H = 1024).
//     Only the recurrent operand (h_{t-1}, 128 rows x H per CTA and item) streams through a TMA ring.
//   * 32 slices x 2 row groups = 64 pairs = 128 CTAs.  Every pair serves TWO 256-row blocks of the batch ("slots") and
//     ping-pongs between them: while the cell epilogue of slot 0 runs, is stored and is handed to the other CTAs, the tensor
//     core works on slot 1.  Two TMEM accumulators, one per slot.
//   * The CTAs that share a 128-row tile hand h_t to each other through L2, k-block by k-block: plain 32-byte stores ->
//     named barrier -> one release increment of the counter of (row tile, 64 hidden units) by a publishing warp; the TMA
//     producers poll those counters and load every k-block as soon as ITS two (four) producers have published.  One launch, no
//     grid-wide barrier; the MMAs of early k-blocks overlap the epilogues of the CTAs that are late.
//   * c stays in registers for all T steps; the x-projection of an item is requested while the previous item is finished;
//     activated gates leave through shared-memory staging + TMA (a store-agent warp), h and c as whole 32-byte sectors.
//
// Same MMAs in the same order as the step-per-launch kernels: h, c and the saved gates are BIT-IDENTICAL
// (tests/test_gemm_gpu.py::test_lstm_resident_forward_is_bit_identical).  Layouts are those of the step-per-launch path:
//   xg / gates [rows, T, 4H] gate-interleaved (column 4u + g), h_all / c_all [rows, T, H], whh_p [4H][H] interleaved rows.
// 16-bit storage (fp16 / bf16) only, rows a multiple of 512, unidirectional; anything else (and DVAE_LSTM_RES=0) takes the
// step-per-launch kernels.
//
// Measured (profiles/r02_lstm_resident.txt): 17.0 -> 12.8 us per step at H = 1024, 9.3 -> 7.6 at H = 512.  What bounds it now is the
// rate at which one SM pulls the recurrent operand through TMA (~60 GB/s: 0.27 us per 16 KB k-block whatever N is), then the
// hand-off chain (release ~1.9 us + propagation ~2.8 us).  Two resident BACKWARD designs were built, verified and dropped (git
// history: c50914d partial sums through L2, 92a8b57 clusters of 8 exchanging partial sums through DSMEM): both slower than the
// step-per-launch backward; numbers in profiles/r02_negative_experiments.txt.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include "act_types.cuh"
#include "host_common.h"
#include "ptx.cuh"
#include "umma_desc.cuh"

namespace dvae {

constexpr int kResThreads = 384;        // warp 0 TMA producer, 1 MMA issuer, 2..9 cell epilogue, 10 store agent, 11 publishing agent
constexpr int kResEpiThreads = 256;
constexpr int kResBarThreads = kResEpiThreads + 32;   // named barriers shared by the epilogue warps and one agent warp
constexpr int kResFlagStride = 32;      // 32-bit words per row tile: one 128-byte line holding its (up to 16) k-block counters
constexpr int kResMaxRowTiles = 8;      // 128-row tiles per launch (1024 rows)
constexpr int kResFlagWords = kResMaxRowTiles * kResFlagStride;
constexpr int kResFlagSlots = 8;
// hand-off counters; a launch uses one slot (zeroed by a memset node in front of it), slots rotate so that launches in
// flight on different streams do not share counters
__device__ unsigned int g_res_flags[kResFlagSlots][kResFlagWords];

struct ResParams {
  CUtensorMap tmA;    // recurrent operand, load: h_all {H, T, rows}, box {64, 1, 128}
  CUtensorMap tmW;    // weight slice, loaded once: whh_p {H, 4H, 1}, box {64, BN/2, 1}
  CUtensorMap tmG;    // activated gates, store: {4H, T, rows}, box {64, 1, 128}
  void* xg;           // x-projection in, activated gates out [rows, T, 4H]
  void* h_all;        // layer output [rows, T, H]
  float* c_all;       // cell states [rows, T, H]
  unsigned int* flags;
  unsigned long long* stamps;   // debug (dvae_debug_res_stamps): globaltimer stamps [T][2 slots][8 points] of CTA 0, or null
  int row0;           // first row of this launch
  int rows;           // rows of the whole tensor
  int T;
};

template <typename AT, int H, int BN>
struct ResCfg {
  static constexpr int KB = H / 64;                 // k-blocks per item
  static constexpr int WKB = (BN / 2) * 128;        // bytes of one resident k-block (this CTA's half of the B tile)
  static constexpr int W_BYTES = KB * WKB;
  static constexpr int STAGE = 128 * 128;           // one k-block of the recurrent operand: 128 rows x 128 B
  static constexpr int G_BOXES = BN * 2 / 128;      // 128-byte boxes of the staged gate tile
  static constexpr int STG = G_BOXES * 16384;       // staged gate tile (everything else leaves as direct 32-byte stores)
  static constexpr int BARS = 256;
  static constexpr int FIT = (232448 - 1024 - BARS - W_BYTES - STG) / STAGE;
  static constexpr int NST = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = W_BYTES + NST * STAGE + STG + BARS + 1024;
  static constexpr int NS = 4 * H / BN;             // weight slices
  static_assert(NST >= 2, "ring too shallow");
};

__device__ __forceinline__ void red_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");   // MEMBAR + RED: no separate __threadfence
}
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// polls are relaxed loads (an acquire load per poll would invalidate the SM's L1 every time); the consumer of a counter is a
// TMA load of L2, issued after the counter value has returned and after a proxy fence
__device__ __forceinline__ uint4 ld_relaxed_v4(const unsigned int* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// "I have finished READING the accumulator": a relaxed arrive on the barrier at this offset in the leader (even-ranked) CTA of
// the pair.  (A release arrive is a cluster-scope fence first and waits for the thread's outstanding global stores.)
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & ptx::kPeerBitMask) : "memory");
}
// 32-byte global store (one full sector per thread: row-per-lane epilogue stores without shared-memory staging)
__device__ __forceinline__ void st_global_v8(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

template <typename AT, int H, int BN>
__global__ void __launch_bounds__(kResThreads, 1) lstm_res_fwd_kernel(const __grid_constant__ ResParams p) {
  using Cfg = ResCfg<AT, H, BN>;
  constexpr int KB = Cfg::KB, NST = Cfg::NST, STAGE = Cfg::STAGE, WKB = Cfg::WKB, NS = Cfg::NS;
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, BN, false, false, 256>();   // both operands K-major, M = 256
  // a k-block of the recurrent operand = 64 hidden units of one time step; it is complete when ARR CTAs have published
  constexpr unsigned int ARR = 64u / (BN / 4);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t w_base = raw_addr + pad;
  const uint32_t ring_base = w_base + Cfg::W_BYTES;
  const uint32_t stg = ring_base + NST * STAGE;
  const uint32_t bar_base = stg + Cfg::STG;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * NST);
  auto acc_full_bar = [&](int a) { return bar_base + 8u * (2 * NST + 1 + a); };
  auto acc_empty_bar = [&](int a) { return bar_base + 8u * (2 * NST + 3 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::W_BYTES + NST * STAGE + Cfg::STG + 8 * (2 * NST + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cr = static_cast<int>(ptx::cluster_ctarank());   // rank in the pair; 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int nt = pair % NS, grp = pair / NS;                 // weight slice (gate-column tile), row group
  const int T = p.T;
  // 128-row tile (within this launch) of slot s: row block 2*grp + s, this CTA's half
  auto row_tile = [&](int s) { return (2 * grp + s) * 2 + cr; };
  unsigned int* ready = p.flags;                   // [row tile][k-block], one line per row tile
  const int my_kb = nt * (BN / 4) / 64;            // the k-block this CTA's hidden units belong to

  // debug stamps of CTA 0: 0 producer starts waiting, 1 first k-block ready, 2 loads issued, 3 MMAs committed, 4 epilogue
  // sees the accumulator, 5 tile done, 6 / 7 publishing agent before / after the release
  auto stamp = [&](int st, int s, int k) {
    if (p.stamps != nullptr && blockIdx.x == 0) p.stamps[(st * 2 + s) * 8 + k] = ptx::globaltimer_ns();
  };
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmW);
    ptx::prefetch_tmap(&p.tmG);
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(acc_full_bar(a), 1);
      ptx::mbar_init(acc_empty_bar(a), 2 * 8);   // one arrival per epilogue warp of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs: own weight half, own A rows)
    if (lane == 0) {
      if (cr == 0) ptx::mbar_expect_tx(wfull_bar, 2 * Cfg::W_BYTES);
      for (int kb = 0; kb < KB; ++kb) ptx::tma_load_3d_pair(w_base + kb * WKB, &p.tmW, wfull_bar, kb * 64, nt * BN + cr * (BN / 2), 0);
      int it = 0;
      for (int st = 1; st < T; ++st) {
        const unsigned int target = ARR * static_cast<unsigned int>(st);
        for (int s = 0; s < 2; ++s) {
          const int rt = row_tile(s);
          const unsigned int* cnt = ready + rt * kResFlagStride;
          uint32_t have = 0;   // bit kb: that k-block of h_{t-1} has been published by all its producers
          stamp(st, s, 0);
          for (int kb = 0; kb < KB; ++kb, ++it) {
            // data-flow start: a k-block is loaded as soon as ITS producers have published
            if (!((have >> kb) & 1u)) {
              const uint64_t t0 = ptx::globaltimer_ns();
              uint32_t spins = 0;
              for (;;) {
#pragma unroll
                for (int v = 0; v < KB / 4; ++v) {
                  const uint4 c4 = ld_relaxed_v4(cnt + 4 * v);
                  have |= (c4.x >= target ? 1u : 0u) << (4 * v) | (c4.y >= target ? 1u : 0u) << (4 * v + 1) |
                          (c4.z >= target ? 1u : 0u) << (4 * v + 2) | (c4.w >= target ? 1u : 0u) << (4 * v + 3);
                }
                if ((have >> kb) & 1u) break;
                if ((++spins & 255u) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {   // a protocol bug traps, it does not hang
                  printf("dvae_b200: resident LSTM hand-off timed out (block %d step %d slot %d k-block %d)\n", blockIdx.x, st, s, kb);
                  __trap();
                }
              }
              fence_proxy_async_all();
              if (kb == 0) stamp(st, s, 1);
            }
            const int sg = it % NST;
            const uint32_t ph = (it / NST) & 1;
            ptx::mbar_wait(empty_bar(sg), ph ^ 1u);
            if (cr == 0) ptx::mbar_expect_tx(full_bar(sg), 2 * STAGE);
            ptx::tma_load_3d_pair(ring_base + sg * STAGE, &p.tmA, full_bar(sg), kb * 64, st - 1, p.row0 + rt * 128);
          }
          stamp(st, s, 2);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (cr == 0) {
      ptx::mbar_wait(wfull_bar, 0);
      ptx::tc_fence_after();
      int it = 0;
      for (int st = 1; st < T; ++st) {
        for (int s = 0; s < 2; ++s) {
          ptx::mbar_wait(acc_empty_bar(s), ((st - 1) & 1) ^ 1u);   // both CTAs have drained this accumulator
          ptx::tc_fence_after();
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const int sg = it % NST;
            const uint32_t ph = (it / NST) & 1;
            ptx::mbar_wait(full_bar(sg), ph);
            ptx::tc_fence_after();
            if (lane == 0) {
              const uint64_t adesc = smem_desc(ring_base + sg * STAGE, 16, 1024, 2);
              const uint64_t bdesc = smem_desc(w_base + kb * WKB, 16, 1024, 2);
#pragma unroll
              for (int k = 0; k < 4; ++k)   // UMMA_K = 16 elements = 32 bytes: descriptor advance 2 (16-byte units)
                ptx::umma_pair<2>(tmem_base + s * BN, adesc + k * 2, bdesc + k * 2, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
              ptx::umma_commit_pair(empty_bar(sg));
            }
            __syncwarp();
          }
          if (lane == 0) {
            ptx::umma_commit_pair(acc_full_bar(s));
            stamp(st, s, 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------ publishing agent: h_t is written by the epilogue threads with plain
    // stores; bar 3 orders them before this thread's gpu-scope release (the pattern of a cooperative grid barrier).  A warp of
    // its own: the release fence then waits for nothing but those stores.
    for (int st = 0; st < T; ++st) {
      for (int s = 0; s < 2; ++s) {
        ptx::bar_sync(3, kResBarThreads);
        if (lane == 0) {
          stamp(st, s, 6);
          red_release_u32(ready + row_tile(s) * kResFlagStride + my_kb, 1u);
          stamp(st, s, 7);
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ store agent: staged gate tiles -> TMA stores
    // bar 1: "staging is full" (epilogue -> agent), bar 2: "staging may be overwritten" (agent -> epilogue); the producing
    // side of each only arrives, so neither ever blocks the other
    for (int st = 0; st < T; ++st) {
      for (int s = 0; s < 2; ++s) {
        ptx::bar_sync(1, kResBarThreads);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < Cfg::G_BOXES; ++b)
            ptx::tma_store_3d(&p.tmG, stg + b * 16384, nt * BN + b * 64, st, p.row0 + row_tile(s) * 128);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();
        }
        __syncwarp();
        ptx::bar_arrive(2, kResBarThreads);
      }
    }
  } else {
    // ------------------------------------------------------------ cell epilogue (warps 2..9 of both CTAs)
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;   // which half of the tile's columns
    const int row = q * 32 + lane;      // row inside this CTA's 128-row tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    constexpr int NCH = BN / 2 / 32;    // 32-column (8-unit) chunks per thread
    const uint32_t sg = stg;
    AT* xg = static_cast<AT*>(p.xg);
    AT* h_all = static_cast<AT*>(p.h_all);
    const long ldx = static_cast<long>(T) * 4 * H, ldc = static_cast<long>(T) * H;
    const int col0 = half * (BN / 2);
    auto x_ptr = [&](int st, int s) {
      return xg + (static_cast<long>(p.row0) + row_tile(s) * 128 + row) * ldx + static_cast<long>(st) * 4 * H + nt * BN + col0;
    };
    float cst[2][NCH * 8];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i) cst[s][i] = 0.f;
    // the x-projection of an item is requested while the previous item is being finished (its registers are free by then)
    typename Act8<AT>::raw_t xr[NCH][4];
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) xr[k][j] = Act8<AT>::load_raw(x_ptr(0, 0) + 32 * k + 8 * j);
    for (int st = 0; st < T; ++st) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const long m = static_cast<long>(p.row0) + row_tile(s) * 128 + row;
        AT* xp = x_ptr(st, s);
        if (st + 1 < T) prefetch_l2_line(xp + 4 * H);   // this slot's x-projection of the next step (HBM -> L2)
        if (st > 0) {
          ptx::mbar_wait(acc_full_bar(s), (st - 1) & 1);
          ptx::tc_fence_after();
        }
        if (threadIdx.x == 64) stamp(st, s, 4);
        if (st > 0 || s > 0) ptx::bar_sync(2, kResBarThreads);   // the previous gate tile has left the staging buffer
        float* co = p.c_all + m * ldc + static_cast<long>(st) * H + nt * (BN / 4) + col0 / 4;
        AT* ho = h_all + m * ldc + static_cast<long>(st) * H + nt * (BN / 4) + col0 / 4;
        uint4 hlo = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = col0 + 32 * k;
          float a[32];
          if (st > 0) {
            __syncwarp();
            ptx::tmem_ld_x32(tmem_base + s * BN + lane_addr + c, a);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = 0.f;
          }
          float x[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) Act8<AT>::unpack(xr[k][j], x + 8 * j);
          float hn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ig = GateMath<AT>::sig(a[4 * i] + x[4 * i]), fg = GateMath<AT>::sig(a[4 * i + 1] + x[4 * i + 1]);
            const float gg = GateMath<AT>::tnh(a[4 * i + 2] + x[4 * i + 2]), og = GateMath<AT>::sig(a[4 * i + 3] + x[4 * i + 3]);
            a[4 * i] = ig; a[4 * i + 1] = fg; a[4 * i + 2] = gg; a[4 * i + 3] = og;
            const float cn = fg * cst[s][8 * k + i] + ig * gg;
            cst[s][8 * k + i] = cn;
            hn[i] = og * GateMath<AT>::tnh(cn);
          }
          // h (what the other CTAs wait for): whole 32-byte sectors, straight from the registers
          if constexpr (NCH == 2) {
            if (k == 0) hlo = pack8_16<AT>(hn);
            else st_global_v8(ho, hlo, pack8_16<AT>(hn));
          } else {
            *reinterpret_cast<uint4*>(ho) = pack8_16<AT>(hn);
          }
          // activated gates: 32 columns = 64 bytes at byte 2c of the staged row (128-byte-swizzled boxes)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int byte = c * 2 + 16 * j;
            ptx::st_shared_v4(sg + static_cast<uint32_t>((byte >> 7) * 16384 + row * 128 + ((((byte & 127) >> 4) ^ (row & 7)) << 4)),
                              pack8_16<AT>(a + 8 * j));
          }
        }
        ptx::bar_arrive(3, kResBarThreads);   // h_t of this tile is stored: publish
        // c (nobody waits for it) only now: the publishing agent's release fence has fewer stores to wait for
#pragma unroll
        for (int k = 0; k < NCH; ++k)
          st_global_v8(co + 8 * k,
                       make_uint4(__float_as_uint(cst[s][8 * k]), __float_as_uint(cst[s][8 * k + 1]), __float_as_uint(cst[s][8 * k + 2]),
                                  __float_as_uint(cst[s][8 * k + 3])),
                       make_uint4(__float_as_uint(cst[s][8 * k + 4]), __float_as_uint(cst[s][8 * k + 5]), __float_as_uint(cst[s][8 * k + 6]),
                                  __float_as_uint(cst[s][8 * k + 7])));
        if (st > 0) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader_relaxed(acc_empty_bar(s));
        }
        ptx::fence_proxy_async_smem();
        ptx::bar_arrive(1, kResBarThreads);   // gate tile staged
        // next item's x-projection
        if (s == 0 || st + 1 < T) {
          const AT* xn = (s == 0) ? x_ptr(st, 1) : x_ptr(st + 1, 0);
#pragma unroll
          for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) xr[k][j] = Act8<AT>::load_raw(xn + 32 * k + 8 * j);
        }
        if (threadIdx.x == 64) stamp(st, s, 5);
      }
    }
    ptx::bar_sync(2, kResBarThreads);   // the store agent's last hand-back
    ptx::tc_fence_before();
  }
  ptx::cluster_sync();   // the leader's MMAs / commits touch the peer: nobody leaves before both are done
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------------------ host side
static int res_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// DVAE_LSTM_RES=0 (or dvae_set_lstm_resident(0)) keeps the step-per-launch kernels for every shape
static int g_res_on = -1;
int lstm_res_set_enabled(int on) {
  if (g_res_on < 0) g_res_on = res_env_int("DVAE_LSTM_RES", 1);
  const int prev = g_res_on;
  if (on >= 0) g_res_on = on;
  return prev;
}

bool lstm_res_supported(int dtype, int rows, int T, int H, int D) {
  return lstm_res_set_enabled(-1) != 0 && (dtype == kBF16 || dtype == kF16) && D == 1 && (H == 512 || H == 1024) && rows >= 512 &&
         rows % 512 == 0 && T >= 2 && num_sms() >= 128;
}

static unsigned long long* g_res_stamps = nullptr;
void lstm_res_set_stamps(unsigned long long* buf) { g_res_stamps = buf; }

static int res_flag_slot(unsigned int** out, cudaStream_t st) {
  static unsigned int* base = nullptr;
  static int next = 0;
  if (base == nullptr) DVAE_CHECK_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_res_flags));
  unsigned int* slot = base + static_cast<long>(next) * kResFlagWords;
  next = (next + 1) % kResFlagSlots;
  DVAE_CHECK_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned int) * kResFlagWords, st));
  *out = slot;
  return 0;
}

template <typename AT, int H, int BN>
static int res_launch(ResParams& p, int rows_l, cudaStream_t st) {
  using Cfg = ResCfg<AT, H, BN>;
  auto kern = lstm_res_fwd_kernel<AT, H, BN>;
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * Cfg::NS * (rows_l / 512));
  cfg.blockDim = dim3(kResThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // every CTA waits for others of the same launch: the whole grid has to be co-resident.  Ask the driver once per kernel; a
  // launch that does not fit is refused (status 3: the caller takes the step-per-launch kernels), never attempted.
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaLaunchConfig_t probe = cfg;
    probe.gridDim = dim3(2 * Cfg::NS * (kResMaxRowTiles * 128 / 512));
    int n = 0;
    DVAE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &probe));
    max_clusters = n;
  }
  if (static_cast<int>(cfg.gridDim.x / 2) > max_clusters) {
    set_last_error("resident LSTM: the device cannot hold " + std::to_string(cfg.gridDim.x / 2) + " CTA pairs of this kernel at once (max " +
                   std::to_string(max_clusters) + ")");
    return 3;
  }
  if (int e = res_flag_slot(&p.flags, st)) return e;
  p.stamps = g_res_stamps;
  DVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

// rows per launch: all CTAs of a launch must be co-resident (they wait for each other)
template <int NS>
static int res_rows_per_launch() {
  const int groups = num_sms() / 2 / NS;
  const int per = groups * 512;
  return per > kResMaxRowTiles * 128 ? kResMaxRowTiles * 128 : per;
}

template <typename AT, int H, int BN>
static int lstm_res_fwd_t(AT* xg, const AT* whh_p, AT* h_all, float* c_all, int rows, int T, cudaStream_t st) {
  using Cfg = ResCfg<AT, H, BN>;
  ResParams p{};
  if (int e = encode_map3(&p.tmA, h_all, 2, H, T, rows, (uint64_t)H * 2, (uint64_t)T * H * 2, 64, 1, 128)) return e;
  if (int e = encode_map3(&p.tmW, whh_p, 2, H, 4 * H, 1, (uint64_t)H * 2, (uint64_t)4 * H * H * 2, 64, BN / 2, 1)) return e;
  if (int e = encode_map3(&p.tmG, xg, 2, 4 * H, T, rows, (uint64_t)4 * H * 2, (uint64_t)T * 4 * H * 2, 64, 1, 128)) return e;
  p.xg = xg;
  p.h_all = h_all;
  p.c_all = c_all;
  p.rows = rows;
  p.T = T;
  const int per = res_rows_per_launch<Cfg::NS>();
  DVAE_REQUIRE(per >= 512, "resident LSTM needs at least 2 x slices SMs");
  for (int r0 = 0; r0 < rows; r0 += per) {
    p.row0 = r0;
    if (int e = res_launch<AT, H, BN>(p, rows - r0 < per ? rows - r0 : per, st)) return e;
  }
  return 0;
}

int lstm_res_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, cudaStream_t st) {
  if (dtype == kF16) {
    using AT = __half;
    return H == 1024 ? lstm_res_fwd_t<AT, 1024, 128>((AT*)xg, (const AT*)whh_p, (AT*)h_all, c_all, rows, T, st)
                     : lstm_res_fwd_t<AT, 512, 64>((AT*)xg, (const AT*)whh_p, (AT*)h_all, c_all, rows, T, st);
  }
  using AT = __nv_bfloat16;
  return H == 1024 ? lstm_res_fwd_t<AT, 1024, 128>((AT*)xg, (const AT*)whh_p, (AT*)h_all, c_all, rows, T, st)
                   : lstm_res_fwd_t<AT, 512, 64>((AT*)xg, (const AT*)whh_p, (AT*)h_all, c_all, rows, T, st);
}

}  // namespace dvae
