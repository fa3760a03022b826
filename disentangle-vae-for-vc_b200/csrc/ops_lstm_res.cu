// Weight-stationary, time-resident LSTM recurrence for the large hidden sizes (H = 512: dec_lstm1, H = 1024: dec_lstm2;
// reference model/disentangled_vae.py:172,193,238,246 and autovc_replicate/proposed_autovc.py:73-75).
//
// The step-per-launch path (ops_gemm.cu) runs one GEMM launch per time step; a step is a strictly serial chain
//   launch + set-up -> pipeline fill -> MMA main loop -> cell epilogue -> stores -> next launch
// (profiles/r01_phase_timing_v2.txt: 8.5 us main loop + 5 us epilogue + 2-3 us launch at H = 1024), and every step
// streams the whole of W_hh (8 MB at H = 1024) from L2 once per 128-row tile.  This file replaces the chain by ONE launch
// per layer and pass:
//
//   * CTA pairs (tcgen05 cta_group::2, M = 256).  A pair owns one SLICE of W_hh for the whole sequence -- forward: BN gate
//     columns (BN / 4 hidden units x i,f,g,o); backward: one gate's K range (H rows) x 128 hidden columns -- and keeps it
//     in shared memory (each CTA of the pair holds half: 128 KB at H = 1024).  Only the recurrent operand (h_{t-1} /
//     da_{t+1}, 128 rows x H per CTA and item) streams through a TMA ring.
//   * 32 slices x 2 row groups = 64 pairs = 128 CTAs.  Every pair serves TWO 256-row blocks of the batch ("slots") and
//     ping-pongs between them: while the cell epilogue of slot 0 runs, is stored and is handed to the other CTAs, the tensor
//     core works on slot 1.  Two TMEM accumulators, one per slot.
//   * The CTAs that share a 128-row tile hand the new recurrent operand to each other through L2: TMA store -> bulk-group
//     completion -> release increment of a per-row-tile counter; consumers poll the counter (acquire) before their TMA loads.
//     One launch, no grid-wide barrier: a row tile only waits for the CTAs that produce ITS columns.
//   * forward: c stays in registers for all T steps.  backward: the dc carry stays in registers; the K range of
//     dh_rec = da_{t+1} . W_hh is split over KQ = 4 pairs (one per gate) whose fp32 partial tiles meet in an L2-resident
//     scratch buffer (plain TMA stores, summed in a fixed order: deterministic), and each of the four CTAs then runs the cell
//     backward for a quarter of the tile's hidden units.
//
// Layouts are those of the step-per-launch path (the two are interchangeable; DVAE_LSTM_RES=0 selects the other):
//   xg / gates [rows, T, 4H] gate-interleaved (column 4u + g), h_all / c_all [rows, T, H], da_all [rows, T, 4H] natural torch
//   order (g*H + u), whh_p [4H][H] interleaved rows (forward B operand, K-major), whh_n [4H][H] natural (backward B operand,
//   MN-major).  16-bit storage (fp16 / bf16) only; rows must be a multiple of 512.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include "act_types.cuh"
#include "host_common.h"
#include "ptx.cuh"
#include "umma_desc.cuh"

namespace dvae {

constexpr int kResThreadsFwd = 384;     // warp 0 TMA producer, 1 MMA issuer, 2..9 cell epilogue, 10 store agent, 11 publishing agent
constexpr int kResThreadsBwd = 352;     // warp 0 TMA producer, 1 MMA issuer, 2..9 partial-sum exchange + cell backward, 10 publishing agent
constexpr int kResEpiThreads = 256;
constexpr int kResBarThreads = kResEpiThreads + 32;   // forward: named barriers shared by the epilogue warps and one agent warp
constexpr int kResKQ = 4;               // backward: K slices (one per gate)
constexpr int kResFlagStride = 32;      // 32-bit words between counters: one 128-byte line each
constexpr int kResMaxRowTiles = 8;      // 128-row tiles per launch (1024 rows)
constexpr int kResMaxNT = 8;            // backward: 128-column tiles of H
constexpr int kResFlagWords = (kResMaxRowTiles + kResMaxRowTiles * kResMaxNT) * kResFlagStride;
constexpr int kResFlagSlots = 8;
// hand-off counters; a launch uses one slot (zeroed by a memset node in front of it), slots rotate so that launches in
// flight on different streams do not share counters
__device__ unsigned int g_res_flags[kResFlagSlots][kResFlagWords];

struct ResParams {
  CUtensorMap tmA;    // recurrent operand, load:  fwd h_all {H, T, rows} / bwd da_all {4H, T, rows}; box {64, 1, 128}
  CUtensorMap tmW;    // weight slice, load once:  fwd whh_p {H, 4H, 1} box {64, BN/2, 1} / bwd W_hh^T {4H, H, 1} box {64, 64, 1}
  CUtensorMap tmO1;   // store: fwd activated gates {4H, T, rows} box {64, 1, 128}; unused by the backward
  void* xg;           // fwd: x-projection in, activated gates out;  bwd: saved activated gates
  float* c_all;
  void* h_all;        // fwd: layer output [rows, T, H];  bwd: da_all [rows, T, 4H] (output)
  const void* dh_all; // bwd: gradient wrt the layer output
  unsigned int* flags;
  unsigned long long* stamps;   // debug (dvae_debug_res_stamps): globaltimer stamps [T][2 slots][8 points] of CTA 0, or null
  int row0;           // first row of this launch
  int rows;           // rows of the whole tensor
  int T;
};

template <typename AT, int H, int BN, bool BWD>
struct ResCfg {
  static constexpr int KB = H / 64;                 // k-blocks per item
  static constexpr int WKB = (BN / 2) * 128;        // bytes of one resident k-block (this CTA's half of the B tile)
  static constexpr int W_BYTES = KB * WKB;
  static constexpr int STAGE = 128 * 128;           // one k-block of the recurrent operand: 128 rows x 128 B
  static constexpr int G_BOXES = BN * 2 / 128;      // fwd: 128-byte boxes of the staged gate tile
  static constexpr int H_ROW = BN / 4 * 2;          // fwd: bytes of h per row and tile
  // fwd: staged gate tile (everything else leaves as direct 32-byte stores).  bwd: receive buffers of the three other K
  // slices' partial sums for this CTA's quarter of the tile, [3][128 rows][32 fp32] (written by the peers through DSMEM)
  static constexpr int STG = BWD ? 3 * 16384 : G_BOXES * 16384;
  static constexpr int BARS = 256;
  static constexpr int FIT = (232448 - 1024 - BARS - W_BYTES - STG) / STAGE;
  static constexpr int NST = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = W_BYTES + NST * STAGE + STG + BARS + 1024;
  static constexpr int NS = BWD ? kResKQ * (H / BN) : 4 * H / BN;   // weight slices = producers of one row tile per step
  static_assert(NST >= 2, "ring too shallow");
  static_assert(!BWD || BN == 128, "backward tiles are 128 hidden columns wide");
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// spin until the counter reaches `target`; a protocol bug traps after 2 s instead of hanging the GPU
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const unsigned int* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// ---- distributed shared memory (cluster of 8 in the backward kernel)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// asynchronous remote store: 16 bytes into another CTA's shared memory, completion counted (bytes) on THAT CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
// Remote arrivals that only say "I have finished READING" (an accumulator in TMEM, a receive buffer): relaxed.  A release
// arrive is a cluster-scope fence first -- it waits for the thread's outstanding global stores (~1 us each, measured).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// ... on the barrier at this offset in the leader (even-ranked) CTA of this CTA's pair
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint32_t bar) { mbar_arrive_remote(bar & ptx::kPeerBitMask); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred P1;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait on a barrier whose arrivals come from other CTAs of the cluster (acquire at cluster scope); traps instead of hanging
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = ptx::globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
      printf("dvae_b200: resident LSTM cluster barrier timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
// tcgen05.commit of a CTA pair inside a larger cluster: arrive on the barrier at this offset in both CTAs of THIS pair
__device__ __forceinline__ void umma_commit_pair_at(uint32_t bar, uint32_t pair_first_rank) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3u << pair_first_rank))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// (polls are relaxed loads: an acquire load per poll invalidates the SM's L1 every time; one acquire fence follows)
__device__ __forceinline__ void wait_counter(const unsigned int* p, unsigned int target) {
  if (ld_relaxed_u32(p) < target) {
    const uint64_t t0 = ptx::globaltimer_ns();
    uint32_t spins = 0;
    while (ld_relaxed_u32(p) < target) {
      if ((++spins & 255u) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
        printf("dvae_b200: resident LSTM hand-off timed out (block %d thread %d target %u have %u)\n", blockIdx.x, threadIdx.x,
               target, ld_relaxed_u32(p));
        __trap();
      }
    }
  }
  (void)ld_acquire_u32(p);   // one acquire load instead of a full fence
}
// 32-byte global store (one full sector per thread: row-per-lane epilogue stores without shared-memory staging)
__device__ __forceinline__ void st_global_v8(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
// publish "this CTA's TMA stores of the step have landed": async-proxy writes -> generic release
__device__ __forceinline__ void publish(unsigned int* counter) {
  fence_proxy_async_all();
  red_release_u32(counter, 1u);   // (a release reduction is MEMBAR + RED: no separate __threadfence, which would be a second MEMBAR)
}

template <typename AT, int H, int BN, bool BWD>
__global__ void __launch_bounds__(BWD ? kResThreadsBwd : kResThreadsFwd, 1) lstm_res_kernel(const __grid_constant__ ResParams p) {
  using Cfg = ResCfg<AT, H, BN, BWD>;
  constexpr int KB = Cfg::KB, NST = Cfg::NST, STAGE = Cfg::STAGE, WKB = Cfg::WKB, NS = Cfg::NS;
  constexpr uint32_t IDESC = instr_desc_fmt<MmaFmt<AT>::value, BN, false, false, 256>();   // both operands K-major
  constexpr uint32_t ADV_B = 32 >> 4;   // descriptor advance per UMMA_K = 16 elements
  constexpr int KQ = kResKQ;
  // a k-block of the recurrent operand = 64 hidden units of one time step; it is complete when ARR CTAs have published
  constexpr unsigned int ARR = BWD ? 2u : 64u / (BN / 4);

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
  const uint32_t recv_full_bar = bar_base + 8u * (2 * NST + 5);   // backward: the peers' partial sums have arrived
  const uint32_t send_ok_bar = bar_base + 8u * (2 * NST + 6);     // backward: the peers have consumed what this CTA sent
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::W_BYTES + NST * STAGE + Cfg::STG + 8 * (2 * NST + 7));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // forward: clusters of 2 (one pair).  backward: clusters of 8 = the KQ pairs that share an output tile, rank = 2 * kq + cr
  const uint32_t crank = ptx::cluster_ctarank();
  const int cr = static_cast<int>(crank & 1u);   // rank in the pair; 0 = leader (issues the MMAs)
  const uint32_t pair_rank0 = crank & ~1u;       // cluster rank of this pair's leader
  const int pair = blockIdx.x >> 1;
  const int slice = pair % NS, grp = pair / NS;
  // forward: slice = gate-column tile.  backward: slice = (hidden-column tile nt, K slice kq); the KQ pairs of one nt are
  // neighbours (one cluster) and reduce into each other
  const int kq = BWD ? slice % KQ : 0;
  const int nt = BWD ? slice / KQ : slice;
  const int T = p.T;
  // 128-row tile (within this launch) of slot s: row block 2*grp + s, this CTA's half
  auto row_tile = [&](int s) { return (2 * grp + s) * 2 + cr; };
  unsigned int* ready = p.flags;                                                  // [row tile][k-block], one line per row tile
  // the k-block this CTA's output belongs to
  const int my_kb = BWD ? nt * 2 + kq / 2 : nt * (BN / 4) / 64;

  // debug stamps of CTA 0.  forward: 0 producer starts waiting, 1 first k-block ready, 2 loads issued, 3 MMAs committed,
  // 4 epilogue sees the accumulator, 5 tile done, 6 / 7 publishing agent before / after the release.  backward: 4 / 5 partial
  // sums: accumulator seen / published, 6 / 7 cell: partial sums seen / da published
  auto stamp = [&](int st, int s, int k) {
    if (p.stamps != nullptr && blockIdx.x == 0) p.stamps[(st * 2 + s) * 8 + k] = ptx::globaltimer_ns();
  };
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmW);
    if constexpr (!BWD) ptx::prefetch_tmap(&p.tmO1);
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(acc_full_bar(a), 1);
      ptx::mbar_init(acc_empty_bar(a), 2 * 8);   // one arrival per epilogue warp of both CTAs
    }
    ptx::mbar_init(recv_full_bar, 1);       // armed by this CTA with the bytes the three peers will send (st.async complete_tx)
    ptx::mbar_init(send_ok_bar, 3 * 8);     // ... and of the three CTAs this one sends to
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
      for (int kb = 0; kb < KB; ++kb) {
        ptx::tma_load_3d_pair(w_base + kb * WKB, &p.tmW, wfull_bar, (BWD ? kq * H : 0) + kb * 64, nt * BN + cr * (BN / 2), 0);
      }
      int it = 0;
      for (int st = 1; st < T; ++st) {
        const int t_a = BWD ? T - st : st - 1;   // time index of the recurrent operand
        const unsigned int target = ARR * static_cast<unsigned int>(st);
        for (int s = 0; s < 2; ++s) {
          const int rt = row_tile(s);
          const unsigned int* cnt = ready + rt * kResFlagStride;
          uint32_t have = 0;   // bit kb: that k-block of h_{t-1} / da_{t+1} has been published by all its producers
          if (!BWD) stamp(st, s, 0);
          for (int kb = 0; kb < KB; ++kb, ++it) {
            // data-flow start: a k-block is loaded as soon as ITS producers have published; the MMAs of the early k-blocks
            // overlap the epilogues of the CTAs that are late
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
                if ((++spins & 255u) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
                  printf("dvae_b200: resident LSTM hand-off timed out (block %d step %d slot %d k-block %d)\n", blockIdx.x, st, s, kb);
                  __trap();
                }
              }
              // the loads below are TMA (async proxy) reads of L2, issued after the counter value has returned
              fence_proxy_async_all();
              if (kb == 0 && !BWD) stamp(st, s, 1);
            }
            const int sg = it % NST;
            const uint32_t ph = (it / NST) & 1;
            ptx::mbar_wait(empty_bar(sg), ph ^ 1u);
            if (cr == 0) ptx::mbar_expect_tx(full_bar(sg), 2 * STAGE);
            ptx::tma_load_3d_pair(ring_base + sg * STAGE, &p.tmA, full_bar(sg), (BWD ? kq * H : 0) + kb * 64, t_a,
                                  p.row0 + rt * 128);
          }
          if (!BWD) stamp(st, s, 2);
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
              for (int k = 0; k < 4; ++k)
                ptx::umma_pair<2>(tmem_base + s * BN, adesc + k * 2, bdesc + k * ADV_B, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit_pair_at(empty_bar(sg), pair_rank0);
            }
            __syncwarp();
          }
          if (lane == 0) {
            umma_commit_pair_at(acc_full_bar(s), pair_rank0);
            stamp(st, s, 3);
          }
          __syncwarp();
        }
      }
    }
  } else if constexpr (!BWD) {
    // =============================================================================================== forward
    if (warp == 11) {
      // ---------------------------------------------------------- publishing agent: h_t is written by the epilogue threads with
      // plain stores; bar 3 orders them before this thread's gpu-scope release (the pattern of a cooperative grid barrier)
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
      // ---------------------------------------------------------- store agent: staged gate tiles -> TMA stores
      // bar 1: "staging is full" (epilogue -> agent), bar 2: "staging may be overwritten" (agent -> epilogue); the producing
      // side of each only arrives, so neither ever blocks the other
      for (int st = 0; st < T; ++st) {
        for (int s = 0; s < 2; ++s) {
          ptx::bar_sync(1, kResBarThreads);
          if (lane == 0) {
#pragma unroll
            for (int b = 0; b < Cfg::G_BOXES; ++b)
              ptx::tma_store_3d(&p.tmO1, stg + b * 16384, nt * BN + b * 64, st, p.row0 + row_tile(s) * 128);
            ptx::bulk_commit();
            ptx::bulk_wait_read<0>();
          }
          __syncwarp();
          ptx::bar_arrive(2, kResBarThreads);
        }
      }
    } else {
      // ---------------------------------------------------------- cell epilogue (warps 2..9 of both CTAs)
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
            // h (what the other CTAs wait for) and c: whole 32-byte sectors, straight from the registers
            if constexpr (NCH == 2) {
              if (k == 0) hlo = pack8_16<AT>(hn);
              else st_global_v8(ho, hlo, pack8_16<AT>(hn));
            } else {
              *reinterpret_cast<uint4*>(ho) = pack8_16<AT>(hn);
            }
            st_global_v8(co + 8 * k,
                         make_uint4(__float_as_uint(cst[s][8 * k]), __float_as_uint(cst[s][8 * k + 1]), __float_as_uint(cst[s][8 * k + 2]),
                                    __float_as_uint(cst[s][8 * k + 3])),
                         make_uint4(__float_as_uint(cst[s][8 * k + 4]), __float_as_uint(cst[s][8 * k + 5]), __float_as_uint(cst[s][8 * k + 6]),
                                    __float_as_uint(cst[s][8 * k + 7])));
            // activated gates: 32 columns = 64 bytes at byte 2c of the staged row (128-byte-swizzled boxes)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int byte = c * 2 + 16 * j;
              ptx::st_shared_v4(sg + static_cast<uint32_t>((byte >> 7) * 16384 + row * 128 + ((((byte & 127) >> 4) ^ (row & 7)) << 4)),
                                pack8_16<AT>(a + 8 * j));
            }
          }
          ptx::bar_arrive(3, kResBarThreads);   // h_t of this tile is stored: publish
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
  } else {
    // =============================================================================================== backward
    if (warp == 10) {
      // publishing agent: da_t is written by the epilogue threads with plain stores; bar 2 orders them before this thread's
      // gpu-scope release.  (A separate warp: the release fence would otherwise also wait for the epilogue thread's own
      // outstanding loads of the next item.)
      for (int st = 0; st < T; ++st) {
        for (int s = 0; s < 2; ++s) {
          ptx::bar_sync(2, kResBarThreads);
          if (lane == 0) {
            red_release_u32(ready + row_tile(s) * kResFlagStride + my_kb, 1u);
            stamp(st, s, 7);
          }
          __syncwarp();
        }
      }
    } else {
    // Epilogue warps 2..9 of every CTA.  The pair's accumulator holds ITS K slice (gate kq) of dh_rec for a 256 x 128 tile; the
    // four pairs of the cluster each own a quarter of the tile's hidden units for the cell backward.  Per item:
    //   1. the three quarters that belong to the other pairs go straight from TMEM into their receive buffers (DSMEM stores,
    //      then one remote mbarrier arrive per warp) -- no global memory, no fence, no flag;
    //   2. the own quarter is summed with the three received ones in a fixed order (deterministic), the senders are told that
    //      their data has been consumed, and the cell backward runs (inputs were requested before the accumulator wait);
    //   3. da_t leaves as direct 32-byte stores and is published to the consumers of the next step (all clusters).
    const AT* gates = static_cast<const AT*>(p.xg);
    const AT* dh_all = static_cast<const AT*>(p.dh_all);
    AT* da_all = static_cast<AT*>(p.h_all);
    const long ldx = static_cast<long>(T) * 4 * H, ldc = static_cast<long>(T) * H;
    constexpr int UPT = BN / KQ / 2;   // hidden units per thread in the cell phase
    static_assert(UPT == 16, "cell phase: 16 hidden units per thread");
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;   // which half of a quarter's 32 units
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int u0 = nt * BN + kq * (BN / KQ) + half * UPT;   // first hidden unit of this thread in the cell phase
    const uint32_t recv = stg;          // [3][128 rows][128 B], 16-byte chunks XOR-swizzled by (row & 7)
    // where this thread's 16 columns of a quarter live inside a receive buffer row: chunks 4*half .. 4*half+3
    auto chunk_off = [&](int j) { return static_cast<uint32_t>(row * 128 + (((4 * half + j) ^ (row & 7)) << 4)); };
    float dcst[2][UPT];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < UPT; ++i) dcst[s][i] = 0.f;
    // saved tensors of an item, requested while the previous item is being finished
    typename Act8<AT>::raw_t g_raw[8];
    typename Act8<float>::raw_t c_raw[2], cp_raw[2];
    typename Act8<AT>::raw_t dh_raw[2];
    auto request = [&](int st, int s) {
      const int t = T - 1 - st;
      const long m = static_cast<long>(p.row0) + row_tile(s) * 128 + row;
      const AT* gp = gates + m * ldx + static_cast<long>(t) * 4 * H + 4 * u0;
      const float* cp = p.c_all + m * ldc + static_cast<long>(t) * H + u0;
      const AT* dp = dh_all + m * ldc + static_cast<long>(t) * H + u0;
#pragma unroll
      for (int j = 0; j < 8; ++j) g_raw[j] = Act8<AT>::load_raw(gp + 8 * j);
#pragma unroll
      for (int j = 0; j < 2; ++j) c_raw[j] = Act8<float>::load_raw(cp + 8 * j);
      if (t > 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) cp_raw[j] = Act8<float>::load_raw(cp - H + 8 * j);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) dh_raw[j] = Act8<AT>::load_raw(dp + 8 * j);
      if (t > 0) {   // what this slot reads at the next step: HBM -> L2 now
        prefetch_l2_line(gp - 4 * H);
        prefetch_l2_line(dp - H);
        if (t > 1) prefetch_l2_line(cp - 2 * H);
      }
    };
    request(0, 0);
    for (int st = 0; st < T; ++st) {
      const int t = T - 1 - st;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int rt = row_tile(s);
        const long m = static_cast<long>(p.row0) + rt * 128 + row;
        float dh[UPT];
#pragma unroll
        for (int j = 0; j < 2; ++j) Act8<AT>::unpack(dh_raw[j], dh + 8 * j);
        if (st > 0) {
          const int item = (st - 1) * 2 + s;   // index among the items that have an accumulator
          ptx::mbar_wait(acc_full_bar(s), (st - 1) & 1);
          ptx::tc_fence_after();
          if (threadIdx.x == 64) stamp(st, s, 4);
          if (threadIdx.x == 64) ptx::mbar_expect_tx(recv_full_bar, 3 * 16384);   // what the three peers send for this item
          if (item > 0) mbar_wait_cluster(send_ok_bar, (item - 1) & 1);   // the peers have consumed the previous item's quarters
          // 1. the other pairs' quarters: TMEM -> their receive buffers
#pragma unroll
          for (int r = 1; r < KQ; ++r) {
            const int dq = (kq + r) & (KQ - 1);                      // destination pair
            const int slot = kq < dq ? kq : kq - 1;                  // sources are kept in ascending kq order
            const uint32_t dst = mapa_u32(recv + slot * 16384, static_cast<uint32_t>(dq * 2 + cr));
            const uint32_t dbar = mapa_u32(recv_full_bar, static_cast<uint32_t>(dq * 2 + cr));
            __syncwarp();
            float v[16];
            tmem_ld_x16(tmem_base + s * BN + lane_addr + dq * (BN / KQ) + half * UPT, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_async_v4(dst + chunk_off(j), make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])), dbar);
          }
          // own quarter
          float own[16];
          __syncwarp();
          tmem_ld_x16(tmem_base + s * BN + lane_addr + kq * (BN / KQ) + half * UPT, own);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader_relaxed(acc_empty_bar(s));
          if (threadIdx.x == 64) stamp(st, s, 5);
          // 2. sum the four K slices in ascending kq order
          mbar_wait_cluster(recv_full_bar, item & 1);
          if (threadIdx.x == 64) stamp(st, s, 6);
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            if (k == kq) {
#pragma unroll
              for (int i = 0; i < 16; ++i) dh[i] += own[i];
            } else {
              const int slot = k < kq ? k : k - 1;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = ptx::ld_shared_v4(recv + slot * 16384 + chunk_off(j));
                dh[4 * j] += __uint_as_float(u.x); dh[4 * j + 1] += __uint_as_float(u.y);
                dh[4 * j + 2] += __uint_as_float(u.z); dh[4 * j + 3] += __uint_as_float(u.w);
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int r = 1; r < KQ; ++r)
              mbar_arrive_remote(mapa_u32(send_ok_bar, static_cast<uint32_t>(((kq + r) & (KQ - 1)) * 2 + cr)));
          }
        }
        // 3. cell backward
        if (threadIdx.x == 64) stamp(st, s, 0);
        float g4[4 * UPT], cc[UPT], cpv[UPT];
#pragma unroll
        for (int j = 0; j < 8; ++j) Act8<AT>::unpack(g_raw[j], g4 + 8 * j);
#pragma unroll
        for (int j = 0; j < 2; ++j) Act8<float>::unpack(c_raw[j], cc + 8 * j);
        if (t > 0) {
#pragma unroll
          for (int j = 0; j < 2; ++j) Act8<float>::unpack(cp_raw[j], cpv + 8 * j);
        } else {
#pragma unroll
          for (int i = 0; i < UPT; ++i) cpv[i] = 0.f;
        }
        AT* dap = da_all + m * ldx + static_cast<long>(t) * 4 * H + u0;
        if (threadIdx.x == 64 && p.stamps != nullptr) {
          float acc = 0.f;
          for (int i = 0; i < 4 * UPT; ++i) acc += g4[i];
          for (int i = 0; i < UPT; ++i) acc += cc[i] + cpv[i] + dh[i];
          if (acc == 123.456f) printf("x");
          stamp(st, s, 1);
        }
        uint4 dalo[4];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float da[4][8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int u = 8 * ch + i;
            const float ig = g4[4 * u], fg = g4[4 * u + 1], gg = g4[4 * u + 2], og = g4[4 * u + 3];
            const float tc = GateMath<AT>::tnh(cc[u]);
            const float dht = dh[u];
            const float dct = dcst[s][u] + dht * og * (1.f - tc * tc);
            da[3][i] = dht * tc * og * (1.f - og);
            da[0][i] = dct * gg * ig * (1.f - ig);
            da[2][i] = dct * ig * (1.f - gg * gg);
            da[1][i] = dct * cpv[u] * fg * (1.f - fg);
            dcst[s][u] = dct * fg;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (ch == 0) dalo[g] = pack8_16<AT>(da[g]);
            else st_global_v8(dap + g * H, dalo[g], pack8_16<AT>(da[g]));
          }
        }
        if (threadIdx.x == 64) stamp(st, s, 2);
        ptx::bar_arrive(2, kResBarThreads);   // da_t of this tile is stored: the agent publishes it
        if (s == 0) request(st, 1);
        else if (st + 1 < T) request(st + 1, 0);
      }
    }
    ptx::tc_fence_before();
    }
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

// shapes the resident kernels handle (whether or not they are switched on)
bool lstm_res_shape_ok(int dtype, int rows, int T, int H, int D) {
  return (dtype == kBF16 || dtype == kF16) && D == 1 && (H == 512 || H == 1024) && rows >= 512 && rows % 512 == 0 && T >= 2 &&
         num_sms() >= 128;
}
bool lstm_res_supported(int dtype, int rows, int T, int H, int D) {
  return lstm_res_set_enabled(-1) != 0 && lstm_res_shape_ok(dtype, rows, T, H, D);
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

template <typename AT, int H, int BN, bool BWD>
static int res_launch(ResParams& p, int rows_l, cudaStream_t st) {
  using Cfg = ResCfg<AT, H, BN, BWD>;
  auto kern = lstm_res_kernel<AT, H, BN, BWD>;
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int groups = rows_l / 512;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * Cfg::NS * groups);
  cfg.blockDim = dim3(BWD ? kResThreadsBwd : kResThreadsFwd);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = BWD ? 2 * kResKQ : 2;   // backward: the KQ pairs of a tile exchange partial sums through DSMEM
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // every CTA waits for others of the same launch: the whole grid has to be co-resident (clusters of 8 need 8 free SMs of
  // one GPC each).  Ask the driver once per kernel; a launch that does not fit is refused, never attempted.
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaLaunchConfig_t probe = cfg;
    probe.gridDim = dim3(2 * Cfg::NS * (kResMaxRowTiles * 128 / 512));
    int n = 0;
    DVAE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &probe));
    max_clusters = n;
    if (res_env_int("DVAE_RES_VERBOSE", 0)) fprintf(stderr, "dvae_b200: resident LSTM kernel (bwd=%d, H=%d): max active clusters %d, smem %d\n", (int)BWD, H, n, Cfg::SMEM);
  }
  if (static_cast<int>(cfg.gridDim.x / attr[0].val.clusterDim.x) > max_clusters) {
    set_last_error("resident LSTM: the device cannot hold " + std::to_string(cfg.gridDim.x / attr[0].val.clusterDim.x) +
                   " clusters of this kernel at once (max " + std::to_string(max_clusters) + ")");
    return 3;
  }
  if (int e = res_flag_slot(&p.flags, st)) return e;
  p.stamps = g_res_stamps;
  DVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

// W_hh [4H][H] -> W_hh^T [H][4H] (16-bit elements): the backward's B operand, K-major like the forward's
__global__ void __launch_bounds__(256) res_transpose16_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int R, int C) {
  __shared__ uint16_t tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = src[static_cast<long>(r0 + i) * C + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) dst[static_cast<long>(c0 + i) * R + r0 + threadIdx.x] = tile[threadIdx.x][i];
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
  using Cfg = ResCfg<AT, H, BN, false>;
  ResParams p{};
  if (int e = encode_map3(&p.tmA, h_all, 2, H, T, rows, (uint64_t)H * 2, (uint64_t)T * H * 2, 64, 1, 128)) return e;
  if (int e = encode_map3(&p.tmW, whh_p, 2, H, 4 * H, 1, (uint64_t)H * 2, (uint64_t)4 * H * H * 2, 64, BN / 2, 1)) return e;
  if (int e = encode_map3(&p.tmO1, xg, 2, 4 * H, T, rows, (uint64_t)4 * H * 2, (uint64_t)T * 4 * H * 2, 64, 1, 128)) return e;
  p.xg = xg;
  p.h_all = h_all;
  p.c_all = c_all;
  p.rows = rows;
  p.T = T;
  const int per = res_rows_per_launch<Cfg::NS>();
  DVAE_REQUIRE(per >= 512, "resident LSTM needs at least 2 x slices SMs");
  for (int r0 = 0; r0 < rows; r0 += per) {
    p.row0 = r0;
    if (int e = res_launch<AT, H, BN, false>(p, rows - r0 < per ? rows - r0 : per, st)) return e;
  }
  return 0;
}

template <typename AT, int H>
static int lstm_res_bwd_t(const AT* dh_all, const AT* gates, const float* c_all, const AT* whh_n, AT* da_all, float* scratch, int rows,
                          int T, cudaStream_t st) {
  constexpr int BN = 128;
  using Cfg = ResCfg<AT, H, BN, true>;
  ResParams p{};
  if (int e = encode_map3(&p.tmA, da_all, 2, 4 * H, T, rows, (uint64_t)4 * H * 2, (uint64_t)T * 4 * H * 2, 64, 1, 128)) return e;
  AT* whh_t = reinterpret_cast<AT*>(scratch);   // [H][4H]
  res_transpose16_kernel<<<dim3(H / 32, 4 * H / 32), dim3(32, 8), 0, st>>>(reinterpret_cast<const uint16_t*>(whh_n),
                                                                         reinterpret_cast<uint16_t*>(whh_t), 4 * H, H);
  DVAE_CHECK_CUDA(cudaGetLastError());
  if (int e = encode_map3(&p.tmW, whh_t, 2, 4 * H, H, 1, (uint64_t)4 * H * 2, (uint64_t)4 * H * H * 2, 64, 64, 1)) return e;
  p.xg = const_cast<AT*>(gates);
  p.h_all = da_all;
  p.c_all = const_cast<float*>(c_all);
  p.dh_all = dh_all;
  p.rows = rows;
  p.T = T;
  const int per = res_rows_per_launch<Cfg::NS>();
  DVAE_REQUIRE(per >= 512, "resident LSTM needs at least 2 x slices SMs");
  for (int r0 = 0; r0 < rows; r0 += per) {
    p.row0 = r0;
    if (int e = res_launch<AT, H, BN, true>(p, rows - r0 < per ? rows - r0 : per, st)) return e;
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

// floats of scratch the resident backward wants: the transposed weight copy [H][4H] of 16-bit elements
long lstm_res_bwd_scratch_floats(int /*rows*/, int H) { return 2L * H * H; }

int lstm_res_bwd(int dtype, const void* dh_all, const void* gates, const float* c_all, const void* whh_n, void* da_all, float* part,
                 int rows, int T, int H, cudaStream_t st) {
  if (part == nullptr) {
    set_last_error("resident LSTM backward needs its scratch buffer (dvae_lstm_bwd_workspace)");
    return 1;
  }
  if (dtype == kF16) {
    using AT = __half;
    return H == 1024 ? lstm_res_bwd_t<AT, 1024>((const AT*)dh_all, (const AT*)gates, c_all, (const AT*)whh_n, (AT*)da_all, part, rows, T, st)
                     : lstm_res_bwd_t<AT, 512>((const AT*)dh_all, (const AT*)gates, c_all, (const AT*)whh_n, (AT*)da_all, part, rows, T, st);
  }
  using AT = __nv_bfloat16;
  return H == 1024 ? lstm_res_bwd_t<AT, 1024>((const AT*)dh_all, (const AT*)gates, c_all, (const AT*)whh_n, (AT*)da_all, part, rows, T, st)
                   : lstm_res_bwd_t<AT, 512>((const AT*)dh_all, (const AT*)gates, c_all, (const AT*)whh_n, (AT*)da_all, part, rows, T, st);
}

}  // namespace dvae
