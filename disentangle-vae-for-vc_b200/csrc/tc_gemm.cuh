// Warp-specialised tcgen05 GEMM core for sm_100a.
//
//   D[128 x BLOCK_N] (fp32, TMEM)  =  sum over k-blocks  A_tile * B_tile^T
//
// * operands are fetched by TMA (3-D tiled tensor maps, 128-byte swizzle) into a STAGES-deep shared
//   memory ring; one elected thread of warp 0 produces, one elected thread of warp 1 issues
//   tcgen05.mma (kind::f16 for bf16, kind::tf32 for fp32 storage), warps 2..5 drain the TMEM
//   accumulator through an epilogue functor.
// * every operand may be K-major (reduction dim contiguous) or MN-major (output dim contiguous);
//   MN-major tiles arrive as several 64-element-wide boxes.  This is what lets forward, dgrad and
//   wgrad of Linear / Conv1d(k=5) / LSTM all run on the same kernel without transposed copies.
// * the TMA box coordinates of k-block `kb` are an affine function of (j, tap, box, tile, z) described
//   by OperandWalk.  Conv1d is an implicit GEMM: the 5 taps are 5 shifted reads of the same
//   channels-last [R, T, C] tensor; rows that fall outside a sequence are zero-filled by TMA.
// * split-K: gridDim.z = batches * splits.  Weight-gradient epilogues accumulate with red.global.add; epilogues that
//   need the complete sum (LSTM cells, bias/activation stores) use an in-kernel fix-up: every split parks its fp32
//   partial tile in a workspace, takes a ticket, and the LAST split to arrive adds the others and runs the epilogue.
// * 8 epilogue warps (two per TMEM lane quarter, each owning half of the tile's columns); while the main loop runs
//   they prefetch the epilogue's global operands into L2.
#pragma once
#include <cuda_bf16.h>

#include "act_types.cuh"
#include "ptx.cuh"
#include "umma_desc.cuh"

namespace dvae {

constexpr int kGemmThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kEpilogueThreads = 256;

struct OperandWalk {
  int base[3];
  int per_j[3];     // k-block index inside a tap
  int per_tap[3];   // conv tap index
  int per_box[3];   // MN-major operands: successive 128-byte-wide boxes
  int per_tile[3];  // blockIdx.x for A, blockIdx.y for B
  int per_z[3];     // batch index (blockIdx.z / splits)
};

struct GemmShape {
  int M, N;        // logical output extent (epilogue masks against it)
  int num_kb;      // k-blocks per output tile (before split-K)
  int kb_per_tap;  // conv: k-blocks per tap; otherwise == num_kb
  int splits;      // split-K factor
  float* splitk_ws;  // fix-up workspace [tiles][splits][128][BLOCK_N] fp32 (epilogues with kFixup, splits > 1)
  int* tickets;      // [tiles] arrival counters, zero before first use, self-resetting
  int f16;           // 16-bit operands are IEEE fp16 rather than bf16 (same kernel, other a_format / b_format bits)
  int batches;       // persistent kernel, > 0: work items are ordered with the BATCH index fastest inside a split (z = split *
                     // batches + batch) instead of the split fastest.  Conv weight gradients: the batch is the filter tap, and
                     // the five taps of a K-slab read the same rows of dy and (shifted by a frame) of x -- running them side by
                     // side keeps that slab in L2 instead of re-reading both tensors from HBM once per tap
};
// instruction descriptor of this launch: the template's bf16 descriptor with the two format fields cleared for fp16
__device__ __forceinline__ uint32_t runtime_idesc(uint32_t idesc, const GemmShape& shp) {
  return shp.f16 ? (idesc & ~((7u << 7) | (7u << 10))) : idesc;
}

// Per-thread view of "this row's accumulator": TMEM, plus the parked partials of the other splits when this CTA is the
// last split to arrive.
struct AccSource {
  uint32_t taddr;        // TMEM address of this warp's lane quarter, column 0 of the tile
  bool has_acc;          // false: no MMA ran for this tile (first LSTM step) -> zeros
  const float* partial;  // workspace row base of split 0 for this thread's row, or nullptr
  int splits, my_split;
  long split_stride;     // elements between consecutive splits in the workspace
  template <int N>
  __device__ __forceinline__ void load(int col, float* v) const {
    if (has_acc) {
      if constexpr (N == 32) ptx::tmem_ld_x32(taddr + col, v);
      else ptx::tmem_ld_x8(taddr + col, v);
      ptx::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = 0.f;
    }
    if (partial != nullptr) {
      for (int sp = 0; sp < splits; ++sp) {
        if (sp == my_split) continue;
        const float4* src = reinterpret_cast<const float4*>(partial + sp * split_stride + col);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
          const float4 t = __ldcg(src + i);
          v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
        }
      }
    }
  }
};

// epilogues that define `static constexpr bool kStaged = true` write through shared-memory staging + TMA
template <class Epi, class = void>
struct epi_is_staged { static constexpr bool value = false; };
template <class Epi>
struct epi_is_staged<Epi, decltype((void)Epi::kStaged)> { static constexpr bool value = Epi::kStaged; };

// staged epilogues may drain a tile in several ROUNDS through a staging buffer that holds 1 / rounds of it (persistent
// kernel only): `static constexpr int rounds<BLOCK_N>()`; default 1
template <class Epi, int BLOCK_N, class = void>
struct epi_rounds { static constexpr int value = 1; };
template <class Epi, int BLOCK_N>
struct epi_rounds<Epi, BLOCK_N, decltype((void)Epi::template rounds<BLOCK_N>())> {
  static constexpr int value = Epi::template rounds<BLOCK_N>();
};

template <class Epi, int BLOCK_N>
__host__ __device__ constexpr int persistent_staging_bytes() {
  if constexpr (epi_is_staged<Epi>::value) return Epi::template staging_bytes<BLOCK_N>() / epi_rounds<Epi, BLOCK_N>::value;
  else return 0;
}

// Optional per-CTA phase stamps (globaltimer ns): [cta][0]=entry [1]=setup done [2]=producer done [3]=accumulator ready
// [4]=epilogue done.  Off (nullptr) unless dvae_debug_timing() installs a buffer; one predictable branch per stamp.
__device__ unsigned long long* g_phase_stamps = nullptr;
__device__ int g_phase_capacity = 0;
__device__ __forceinline__ void phase_stamp(int slot) {
  unsigned long long* buf = g_phase_stamps;
  if (buf != nullptr) {
    const long cta = (static_cast<long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta < g_phase_capacity) buf[cta * 5 + slot] = ptx::globaltimer_ns();
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// prefetch `bytes` starting at p (128-byte lines)
__device__ __forceinline__ void prefetch_l2_span(const void* p, int bytes) {
  const char* c = static_cast<const char*>(p);
  for (int o = 0; o < bytes; o += 128) prefetch_l2(c + o);
}

template <int BLOCK_N, int STAGES, int M_TILES = 1>
constexpr int gemm_smem_bytes() {
  return STAGES * (M_TILES * kBlockM * kSwizzleRow + BLOCK_N * kSwizzleRow) + 1024 /*align slack*/ + 256 /*barriers*/;
}

// M_TILES = 2: the CTA owns 256 output rows as two 128-row accumulators that share every B stage.  ncu shows the
// 128 x 256 tile already pulls ~11.6 TB/s from L2 (profiles/r01_ncu_gemm_v1.txt), i.e. the big GEMMs are bound by L2->SM
// bandwidth, not by the tensor pipe; the 256 x 256 tile moves 1/3 fewer operand bytes per FLOP.
//
// CTA_GROUP = 2: CTA pairs (clusters of two consecutive M tiles) run the 256 x BLOCK_N MMA of tcgen05 cta_group::2.  Each
// CTA stages its own 128 A rows and only HALF of the B tile (the tensor core reads both halves across the pair), so a
// stage is 16 + BLOCK_N/2 * 128 bytes and 227 KB hold 7 stages instead of 4.  That is what the recurrence step GEMMs
// need: their main loop is bound by the operand bytes one SM can keep in flight against ~1.7 us of loaded L2 latency,
// and the pair MMA needs 1/3 fewer bytes per FLOP per SM.  Both CTAs run a TMA producer (complete_tx goes to the leader's
// full barrier), only the leader (even rank) issues MMAs; its commits arrive on the empty / accumulator barriers of both.
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int ELEM_BYTES, class Epi, int M_TILES = 1, int CTA_GROUP = 1>
__global__ void __launch_bounds__(kGemmThreads, (CTA_GROUP == 1 && Epi::kCtasPerSm == 2 && STAGES * (M_TILES * kBlockM + BLOCK_N) * kSwizzleRow <= 100 * 1024) ? 2 : 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const OperandWalk wa, const OperandWalk wb, const GemmShape shp,
               const __grid_constant__ typename Epi::Params ep) {
  static_assert(CTA_GROUP == 1 || (CTA_GROUP == 2 && M_TILES == 1), "CTA pairs use the single-accumulator tile");
  constexpr bool PAIR = CTA_GROUP == 2;
  constexpr int BLOCK_K = kSwizzleRow / ELEM_BYTES;  // 64 bf16 / 32 tf32
  constexpr int UMMA_K = 32 / ELEM_BYTES;            // 16 / 8
  constexpr int TILE_A = kBlockM * kSwizzleRow;      // one 128-row A tile
  constexpr int STAGE_A = M_TILES * TILE_A;
  constexpr int TMEM_COLS = M_TILES * BLOCK_N;
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  constexpr int STAGE_B = BLOCK_N * kSwizzleRow / CTA_GROUP;   // a pair CTA holds half of the B tile
  constexpr int A_BOXES = A_MN ? (kBlockM * ELEM_BYTES / kSwizzleRow) : 1;
  constexpr int B_BOXES = B_MN ? (BLOCK_N * ELEM_BYTES / kSwizzleRow / CTA_GROUP) : 1;   // boxes this CTA loads
  static_assert(B_BOXES >= 1, "B tile too narrow to split over a CTA pair");
  constexpr int MN_BOX_BYTES = BLOCK_K * kSwizzleRow;  // one MN-major box: BLOCK_K rows of 128 B
  constexpr int A_BOX_BYTES = A_MN ? MN_BOX_BYTES : TILE_A;
  constexpr int B_BOX_BYTES = B_MN ? MN_BOX_BYTES : STAGE_B;
  constexpr uint32_t ADV_A = (A_MN ? UMMA_K * kSwizzleRow : 32) >> 4;  // descriptor advance per MMA (16 B units)
  constexpr uint32_t ADV_B = (B_MN ? UMMA_K * kSwizzleRow : 32) >> 4;
  constexpr uint32_t IDESC = instr_desc<ELEM_BYTES, BLOCK_N, A_MN, B_MN, kBlockM * CTA_GROUP>();
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "TMEM allocation must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  const uint32_t bar_base = smem_base + STAGES * (STAGE_A + STAGE_B);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int zb = blockIdx.z / shp.splits;
  const int split = blockIdx.z - zb * shp.splits;
  // split-K range of this CTA
  const int kb_chunk = (shp.num_kb + shp.splits - 1) / shp.splits;
  const int kb_begin = split * kb_chunk;
  const int kb_end = min(shp.num_kb, kb_begin + kb_chunk);
  const int num_local = max(0, kb_end - kb_begin);
  const int cr = PAIR ? static_cast<int>(ptx::cluster_ctarank()) : 0;   // rank in the pair; 0 = leader

  if (threadIdx.x == 0) phase_stamp(0);
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();   // the peer must see initialised barriers before its loads / commits arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the set-up above overlapped the previous kernel; from here on we read / write its data
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  if (threadIdx.x == 0) phase_stamp(1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (in a pair: both CTAs, own A rows + own B half)
    if (lane == 0) {
      for (int it = 0; it < num_local; ++it) {
        const int kb = kb_begin + it;
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(empty_bar(s), ph ^ 1u);
        if (!PAIR || cr == 0) ptx::mbar_expect_tx(full_bar(s), CTA_GROUP * (STAGE_A + STAGE_B));   // bytes of both CTAs
        const int tap = kb / shp.kb_per_tap;
        const int j = kb - tap * shp.kb_per_tap;
        const uint32_t sa = smem_base + s * (STAGE_A + STAGE_B);
        const uint32_t sb = sa + STAGE_A;
#pragma unroll
        for (int mt = 0; mt < M_TILES; ++mt) {
#pragma unroll
          for (int i = 0; i < A_BOXES; ++i) {
            int c[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
              c[d] = wa.base[d] + j * wa.per_j[d] + tap * wa.per_tap[d] + i * wa.per_box[d] +
                     (tile_m * M_TILES + mt) * wa.per_tile[d] + zb * wa.per_z[d];
            if constexpr (PAIR) ptx::tma_load_3d_pair(sa + i * A_BOX_BYTES, &tmA, full_bar(s), c[0], c[1], c[2]);
            else ptx::tma_load_3d(sa + mt * TILE_A + i * A_BOX_BYTES, &tmA, full_bar(s), c[0], c[1], c[2]);
          }
        }
#pragma unroll
        for (int i = 0; i < B_BOXES; ++i) {
          int c[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            c[d] = wb.base[d] + j * wb.per_j[d] + tap * wb.per_tap[d] + (cr * B_BOXES + i) * wb.per_box[d] +
                   tile_n * wb.per_tile[d] + zb * wb.per_z[d];
            if (PAIR && !B_MN) c[d] += cr * (wb.per_tile[d] / 2);   // K-major: second half of the tile's rows
          }
          if constexpr (PAIR) ptx::tma_load_3d_pair(sb + i * B_BOX_BYTES, &tmB, full_bar(s), c[0], c[1], c[2]);
          else ptx::tma_load_3d(sb + i * B_BOX_BYTES, &tmB, full_bar(s), c[0], c[1], c[2]);
        }
      }
      phase_stamp(2);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (in a pair: the leader only)
    const uint32_t idesc = runtime_idesc(IDESC, shp);
    if (!PAIR || cr == 0) {
      for (int it = 0; it < num_local; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(full_bar(s), ph);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + s * (STAGE_A + STAGE_B);
          const uint32_t sb = sa + STAGE_A;
          // K-major: 8-row x 128 B swizzle atoms, 1024 B apart.  MN-major: atoms of (128 B along MN) x (8 k-rows),
          // next atom along MN one TMA box further (LBO), next k-group SBO further.  32-bit MN-major operands
          // only exist in the 32-byte-atom swizzle (4 k-rows per atom): layout type 1, SBO 512.
          constexpr uint32_t MN_LAYOUT = (ELEM_BYTES == 4) ? 1u : 2u;
          constexpr uint32_t MN_SBO = (ELEM_BYTES == 4) ? 512u : 1024u;
          const uint64_t bdesc = B_MN ? smem_desc(sb, MN_BOX_BYTES, MN_SBO, MN_LAYOUT) : smem_desc(sb, 16, 1024, 2);
#pragma unroll
          for (int mt = 0; mt < M_TILES; ++mt) {
            const uint32_t sam = sa + mt * TILE_A;
            const uint64_t adesc = A_MN ? smem_desc(sam, MN_BOX_BYTES, MN_SBO, MN_LAYOUT) : smem_desc(sam, 16, 1024, 2);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if constexpr (PAIR)
                ptx::umma_pair<ELEM_BYTES>(tmem_base, adesc + k * ADV_A, bdesc + k * ADV_B, idesc, (it > 0 || k > 0) ? 1u : 0u);
              else
                ptx::umma<ELEM_BYTES>(tmem_base + mt * BLOCK_N, adesc + k * ADV_A, bdesc + k * ADV_B, idesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          if constexpr (PAIR) ptx::umma_commit_pair(empty_bar(s));   // frees the slot in both CTAs
          else ptx::umma_commit(empty_bar(s));                        // frees the smem slot once these MMAs retire
        }
        __syncwarp();
      }
      if (num_local > 0 && lane == 0) {
        if constexpr (PAIR) ptx::umma_commit_pair(tmem_full_bar);
        else ptx::umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;  // M_TILES == 1: which half of the tile's columns; M_TILES == 2: which accumulator
    const int mt = (M_TILES == 2) ? half : 0;
    const int row = mt * kBlockM + q * 32 + lane;             // row inside the CTA tile
    const int m = tile_m * (M_TILES * kBlockM) + row;
    const int n0 = tile_n * BLOCK_N;
    const int col0 = (M_TILES == 2) ? 0 : half * (BLOCK_N / 2);
    const int col1 = (M_TILES == 2) ? BLOCK_N : col0 + BLOCK_N / 2;
    Epi::template prefetch<BLOCK_N>(ep, m, n0, zb, col0, col1, shp);   // overlaps the main loop
    typename Epi::template Regs<BLOCK_N> regs;
    Epi::template preload<BLOCK_N>(ep, regs, m, n0, zb, col0, col1, shp);   // operands that do not depend on the MMA
    AccSource acc;
    acc.taddr = tmem_base + mt * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
    acc.has_acc = num_local > 0;
    acc.partial = nullptr;
    acc.splits = shp.splits;
    acc.my_split = split;
    acc.split_stride = static_cast<long>(M_TILES * kBlockM) * BLOCK_N;
    if (num_local > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
    if (threadIdx.x == 64) phase_stamp(3);
    bool run_epilogue = true;
    if (Epi::kFixup && shp.splits > 1) {
      const long tile_id = (static_cast<long>(zb) * gridDim.y + tile_n) * gridDim.x + tile_m;
      float* ws_row = shp.splitk_ws + (tile_id * shp.splits * (M_TILES * kBlockM) + row) * BLOCK_N;
      float* mine = ws_row + split * acc.split_stride;
      for (int c = col0; c < col1; c += 32) {
        __syncwarp();
        float v[32];
        acc.template load<32>(c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          __stcg(reinterpret_cast<float4*>(mine + c) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int* last_flag = reinterpret_cast<int*>(smem + STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 2));
      if (threadIdx.x == 64) {
        const int old = atomicAdd(shp.tickets + tile_id, 1);
        const int last = (old == shp.splits - 1) ? 1 : 0;
        if (last) shp.tickets[tile_id] = 0;   // self-reset for the next launch
        *last_flag = last;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      run_epilogue = (*last_flag != 0);
      if (run_epilogue) {
        __threadfence();
        acc.partial = ws_row;
      }
    }
    if constexpr (epi_is_staged<Epi>::value) {
      // Staged epilogue: the accumulator is complete, so every MMA has retired and every TMA load has landed -- the
      // operand ring is free.  The threads deposit their results there as 128-byte-swizzled boxes (row-per-lane
      // st.shared.v4 is conflict free, unlike row-per-lane global accesses, which cost one LSU transaction per lane),
      // and one thread hands the boxes to the TMA engine.
      static_assert(M_TILES == 1, "staged epilogues use the single-accumulator tile");
      static_assert(Epi::template staging_bytes<BLOCK_N>() <= STAGES * (STAGE_A + STAGE_B), "staging exceeds the ring");
      if (run_epilogue) Epi::template run_staged<BLOCK_N>(ep, acc, regs, row, m, n0, zb, col0, col1, shp, smem_base);
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        if (run_epilogue) Epi::template flush<BLOCK_N>(ep, smem_base, tile_m, tile_n, zb, shp);
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();   // shared memory must stay alive until the TMA engine has read it; the writes themselves
                                    // are ordered before grid completion, which is what the dependent kernel (PDL) waits for
      }
    } else {
      if (run_epilogue) Epi::template run<BLOCK_N>(ep, acc, regs, m, n0, zb, col0, col1, shp);
    }
    if (threadIdx.x == 64) phase_stamp(4);
    ptx::tc_fence_before();
  }
  if constexpr (PAIR) ptx::cluster_sync();   // the leader's commits / MMAs touch the peer: nobody leaves before both are done
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================
// Persistent variant for the large contractions (conv / linear / x-projection, forward, dgrad and wgrad): one CTA per
// SM loops over output tiles; the accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile
// i overlaps the main loop of tile i+1, and the TMA ring keeps running across tile boundaries (no cold start per tile).
// ncu on the one-tile-per-CTA kernel: tensor pipe active 52 %, of which the un-overlapped epilogue + tile start cost
// ~30 % (profiles/r01_phase_timing_v2.txt).  Split-K is supported for accumulate-type epilogues only (no fix-up).
// =====================================================================================
// CTA_GROUP = 2: clusters of two CTAs walk the tile list together (pair tile = two consecutive M tiles of the same N tile,
// batch and split) and run the 256 x BLOCK_N MMA of tcgen05 cta_group::2: each SM stages its own A rows and half of the B
// tile, i.e. 2/3 of the operand bytes per FLOP -- the large contractions are bound by L2->SM traffic (~97 GB/s per SM,
// profiles/r01_ncu_targets_v2.txt), not by the tensor pipe.  The leader issues the MMAs; its commits release the operand
// slot and publish the accumulator in both CTAs; the epilogue warps of both CTAs hand the accumulator back by arriving
// on the leader's barrier.
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int ELEM_BYTES, class Epi, int CTA_GROUP = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const OperandWalk wa, const OperandWalk wb, const GemmShape shp,
                          const __grid_constant__ typename Epi::Params ep, const int tiles_m, const int tiles_n,
                          const int num_tiles) {
  constexpr bool PAIR = CTA_GROUP == 2;
  constexpr int BLOCK_K = kSwizzleRow / ELEM_BYTES;
  constexpr int UMMA_K = 32 / ELEM_BYTES;
  constexpr int STAGE_A = kBlockM * kSwizzleRow;
  constexpr int STAGE_B = BLOCK_N * kSwizzleRow / CTA_GROUP;
  constexpr int A_BOXES = A_MN ? (kBlockM * ELEM_BYTES / kSwizzleRow) : 1;
  constexpr int B_BOXES = B_MN ? (BLOCK_N * ELEM_BYTES / kSwizzleRow / CTA_GROUP) : 1;
  static_assert(B_BOXES >= 1, "B tile too narrow to split over a CTA pair");
  constexpr int MN_BOX_BYTES = BLOCK_K * kSwizzleRow;
  constexpr int A_BOX_BYTES = A_MN ? MN_BOX_BYTES : STAGE_A;
  constexpr int B_BOX_BYTES = B_MN ? MN_BOX_BYTES : STAGE_B;
  constexpr uint32_t ADV_A = (A_MN ? UMMA_K * kSwizzleRow : 32) >> 4;
  constexpr uint32_t ADV_B = (B_MN ? UMMA_K * kSwizzleRow : 32) >> 4;
  constexpr uint32_t IDESC = instr_desc<ELEM_BYTES, BLOCK_N, A_MN, B_MN, kBlockM * CTA_GROUP>();
  constexpr int TMEM_COLS = 2 * BLOCK_N;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "two accumulators must fit TMEM");
  constexpr int RING = STAGES * (STAGE_A + STAGE_B);
  constexpr int STAGING = persistent_staging_bytes<Epi, BLOCK_N>();   // staged epilogues: one output tile, after the ring

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  const uint32_t bar_base = smem_base + RING + STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + RING + STAGING + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_chunk = (shp.num_kb + shp.splits - 1) / shp.splits;
  const int cr = PAIR ? static_cast<int>(ptx::cluster_ctarank()) : 0;   // rank in the pair; 0 = leader
  // work items: tiles, or pair tiles (two consecutive M tiles); walkers: CTAs, or clusters
  const int items = num_tiles / CTA_GROUP;
  const int walker = blockIdx.x / CTA_GROUP, walkers = gridDim.x / CTA_GROUP;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), CTA_GROUP * kEpilogueThreads / 32);   // one arrival per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync();
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  // item -> (tile_m, tile_n, batch zb, split): m fastest, so CTAs running side by side share the B (weight) tile in L2
  const int items_m = tiles_m / CTA_GROUP;
  auto decode = [&](int item, int& tm, int& tn, int& zb, int& split) {
    tm = (item % items_m) * CTA_GROUP + cr;
    const int rest = item / items_m;
    tn = rest % tiles_n;
    const int zz = rest / tiles_n;
    if (shp.batches > 0) {
      split = zz / shp.batches;
      zb = zz - split * shp.batches;
    } else {
      zb = zz / shp.splits;
      split = zz - zb * shp.splits;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (pair: both CTAs, own A rows + own half of B)
    if (lane == 0) {
      int it = 0;   // running k-block counter of this CTA: the ring never drains between tiles
      for (int tile = walker; tile < items; tile += walkers) {
        int tm, tn, zb, split;
        decode(tile, tm, tn, zb, split);
        const int kb_begin = split * kb_chunk;
        const int kb_end = min(shp.num_kb, kb_begin + kb_chunk);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          if (!PAIR || cr == 0) ptx::mbar_expect_tx(full_bar(s), CTA_GROUP * (STAGE_A + STAGE_B));
          const int tap = kb / shp.kb_per_tap;
          const int j = kb - tap * shp.kb_per_tap;
          const uint32_t sa = smem_base + s * (STAGE_A + STAGE_B);
          const uint32_t sb = sa + STAGE_A;
#pragma unroll
          for (int i = 0; i < A_BOXES; ++i) {
            int c[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
              c[d] = wa.base[d] + j * wa.per_j[d] + tap * wa.per_tap[d] + i * wa.per_box[d] + tm * wa.per_tile[d] +
                     zb * wa.per_z[d];
            if constexpr (PAIR) ptx::tma_load_3d_pair(sa + i * A_BOX_BYTES, &tmA, full_bar(s), c[0], c[1], c[2]);
            else ptx::tma_load_3d(sa + i * A_BOX_BYTES, &tmA, full_bar(s), c[0], c[1], c[2]);
          }
#pragma unroll
          for (int i = 0; i < B_BOXES; ++i) {
            int c[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              c[d] = wb.base[d] + j * wb.per_j[d] + tap * wb.per_tap[d] + (cr * B_BOXES + i) * wb.per_box[d] +
                     tn * wb.per_tile[d] + zb * wb.per_z[d];
              if (PAIR && !B_MN) c[d] += cr * (wb.per_tile[d] / 2);
            }
            if constexpr (PAIR) ptx::tma_load_3d_pair(sb + i * B_BOX_BYTES, &tmB, full_bar(s), c[0], c[1], c[2]);
            else ptx::tma_load_3d(sb + i * B_BOX_BYTES, &tmB, full_bar(s), c[0], c[1], c[2]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (pair: the leader only)
    constexpr uint32_t MN_LAYOUT = (ELEM_BYTES == 4) ? 1u : 2u;
    constexpr uint32_t MN_SBO = (ELEM_BYTES == 4) ? 512u : 1024u;
    const uint32_t idesc = runtime_idesc(IDESC, shp);
    if (!PAIR || cr == 0) {
      int it = 0, local = 0;
      for (int tile = walker; tile < items; tile += walkers, ++local) {
        int tm, tn, zb, split;
        decode(tile, tm, tn, zb, split);
        const int kb_begin = split * kb_chunk;
        const int num_local = max(0, min(shp.num_kb, kb_begin + kb_chunk) - kb_begin);
        const int as = local & 1;
        ptx::mbar_wait(tmem_empty_bar(as), ((local >> 1) & 1) ^ 1u);   // the epilogues have drained this accumulator
        ptx::tc_fence_after();
        for (int k0 = 0; k0 < num_local; ++k0, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(full_bar(s), ph);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_base + s * (STAGE_A + STAGE_B);
            const uint32_t sb = sa + STAGE_A;
            const uint64_t adesc = A_MN ? smem_desc(sa, MN_BOX_BYTES, MN_SBO, MN_LAYOUT) : smem_desc(sa, 16, 1024, 2);
            const uint64_t bdesc = B_MN ? smem_desc(sb, MN_BOX_BYTES, MN_SBO, MN_LAYOUT) : smem_desc(sb, 16, 1024, 2);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if constexpr (PAIR)
                ptx::umma_pair<ELEM_BYTES>(tmem_base + as * BLOCK_N, adesc + k * ADV_A, bdesc + k * ADV_B, idesc,
                                           (k0 > 0 || k > 0) ? 1u : 0u);
              else
                ptx::umma<ELEM_BYTES>(tmem_base + as * BLOCK_N, adesc + k * ADV_A, bdesc + k * ADV_B, idesc,
                                      (k0 > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (PAIR) ptx::umma_commit_pair(empty_bar(s));
            else ptx::umma_commit(empty_bar(s));
          }
          __syncwarp();
        }
        if (lane == 0) {
          if constexpr (PAIR) ptx::umma_commit_pair(tmem_full_bar(as));
          else ptx::umma_commit(tmem_full_bar(as));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int col0 = half * (BLOCK_N / 2), col1 = col0 + BLOCK_N / 2;
    int local = 0;
    // the accumulator goes back to the MMA issuer: in a pair that is the leader's barrier, for the warps of both CTAs
    auto release_acc = [&](int as) {
      if constexpr (PAIR) ptx::mbar_arrive_leader(tmem_empty_bar(as));
      else ptx::mbar_arrive(tmem_empty_bar(as));
    };
    for (int tile = walker; tile < items; tile += walkers, ++local) {
      int tm, tn, zb, split;
      decode(tile, tm, tn, zb, split);
      const int as = local & 1;
      const int m = tm * kBlockM + q * 32 + lane;
      const int n0 = tn * BLOCK_N;
      Epi::template prefetch<BLOCK_N>(ep, m, n0, zb, col0, col1, shp);
      typename Epi::template Regs<BLOCK_N> regs;
      Epi::template preload<BLOCK_N>(ep, regs, m, n0, zb, col0, col1, shp);
      const int kb_begin = split * kb_chunk;
      AccSource acc;
      acc.taddr = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      acc.has_acc = min(shp.num_kb, kb_begin + kb_chunk) > kb_begin;   // an empty trailing split contributes nothing
      acc.partial = nullptr;
      acc.splits = shp.splits;
      acc.my_split = split;
      acc.split_stride = 0;
      ptx::mbar_wait(tmem_full_bar(as), (local >> 1) & 1);
      ptx::tc_fence_after();
      if constexpr (epi_is_staged<Epi>::value) {
        // Output tile -> swizzled staging -> TMA store.  The accumulator is released as soon as it is in shared memory;
        // the store of tile i drains while the main loop of tile i+1 runs, and is only waited for when the staging
        // buffer is needed again.
        const uint32_t stage = smem_base + RING;
        constexpr int ROUNDS = epi_rounds<Epi, BLOCK_N>::value;
        if constexpr (ROUNDS == 1) {
          if (local > 0) {
            if (threadIdx.x == 64) ptx::bulk_wait_read<0>();
            asm volatile("bar.sync 2, 256;" ::: "memory");
          }
          Epi::template run_staged<BLOCK_N>(ep, acc, regs, q * 32 + lane, m, n0, zb, col0, col1, shp, stage);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(as);
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (threadIdx.x == 64) {
            Epi::template flush<BLOCK_N>(ep, stage, tm, tn, zb, shp);
            ptx::bulk_commit();
          }
          Epi::template after_stage<BLOCK_N>(ep, stage, static_cast<int>(threadIdx.x) - 64, tm, tn, zb, shp);   // reads the staged tile
        } else {
          // Wide (fp32) output tiles: the staging buffer holds BLOCK_N / ROUNDS columns, so the ring keeps its depth.  The
          // epilogue has the whole main loop of the next tile to hide in, so draining in rounds costs nothing.
          constexpr int RN = BLOCK_N / ROUNDS;
#pragma unroll 1
          for (int r = 0; r < ROUNDS; ++r) {
            if (local > 0 || r > 0) {
              if (threadIdx.x == 64) ptx::bulk_wait_read<0>();
              asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            const int c0 = r * RN + half * (RN / 2);
            Epi::template run_staged_r<BLOCK_N>(ep, acc, regs, q * 32 + lane, m, n0, zb, c0, c0 + RN / 2, shp, stage, r * RN);
            if (r == ROUNDS - 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc(as);
            }
            ptx::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
              Epi::template flush_r<BLOCK_N>(ep, stage, tm, tn, zb, shp, r * RN, RN);
              ptx::bulk_commit();
            }
            Epi::template after_stage_r<BLOCK_N>(ep, stage, static_cast<int>(threadIdx.x) - 64, tm, tn, zb, shp, r * RN, RN);
          }
        }
      } else {
        Epi::template run<BLOCK_N>(ep, acc, regs, m, n0, zb, col0, col1, shp);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(as);   // this warp no longer reads accumulator `as`
      }
    }
    if constexpr (epi_is_staged<Epi>::value) {
      if (threadIdx.x == 64) ptx::bulk_wait_read<0>();
    }
  }
  if constexpr (PAIR) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================
// Epilogues.  Each receives the TMEM address of its warp's 32-lane slice; thread `lane` owns
// output row m and reads BLOCK_N fp32 columns in chunks.
// =====================================================================================
// ---- out_f32[zb][m][n] += acc, as a TMA reduce-add of the staged tile (split-K partial sums meet in L2; no tickets,
// no row-per-lane atomics).  Staging: [BLOCK_N / 32 boxes][128 rows][128 B] fp32, swizzle-128B.
struct EpiReduceTma {
  static constexpr bool kFixup = false;
  static constexpr bool kStaged = true;
  static constexpr int kCtasPerSm = 1;
  struct Params {
    CUtensorMap tm_out;   // fp32 {N, M, batches}, box {32, 128, 1}
    // Optional: up to four strided row sets that the NEXT kernel in the stream streams from HBM (the LSTM cell backward
    // reads this time step's saved gates / cell states / output gradient).  The epilogue warps are idle while the main loop
    // runs, so they pull those lines into L2: the memory-bound consumer then runs at L2 rather than HBM bandwidth.
    const void* pf_ptr[4];
    long pf_row_stride[4];   // bytes between rows
    int pf_row_bytes[4];     // contiguous bytes per row
    int pf_rows;
  };
  template <int BLOCK_N>
  static __host__ __device__ constexpr int staging_bytes() { return BLOCK_N * 4 * kBlockM; }
  template <int BLOCK_N>
  static __device__ __forceinline__ void after_stage(const Params&, uint32_t, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params& p, int, int, int, int, int, const GemmShape&) {
    if ((threadIdx.x & 255) == 64) ptx::prefetch_tmap(&p.tm_out);
    if (p.pf_rows <= 0) return;
    const long cta = (static_cast<long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const long nthr = static_cast<long>(gridDim.x) * gridDim.y * gridDim.z * kEpilogueThreads;
    const long me = cta * kEpilogueThreads + (static_cast<int>(threadIdx.x) - 64);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const char* base = static_cast<const char*>(p.pf_ptr[k]);
      if (base == nullptr) continue;
      const int lpr = (p.pf_row_bytes[k] + 127) >> 7;
      const long total = static_cast<long>(p.pf_rows) * lpr;
      for (long i = me; i < total; i += nthr) {
        const long r = i / lpr;
        prefetch_l2(base + r * p.pf_row_stride[k] + ((i - r * lpr) << 7));
      }
    }
  }
  template <int BLOCK_N> struct Regs {};
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params&, Regs<BLOCK_N>&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run_staged(const Params&, const AccSource& acc, const Regs<BLOCK_N>&, int row, int m,
                                                    int n0, int zb, int col0, int col1, const GemmShape& shp,
                                                    uint32_t stage) {
#pragma unroll 1
    for (int c = col0; c < col1; c += 32) {
      __syncwarp();
      float v[32];
      acc.template load<32>(c, v);
      const uint32_t box = stage + static_cast<uint32_t>((c >> 5) * (kBlockM * 128) + row * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        ptx::st_shared_v4(box + ((j ^ (row & 7)) << 4),
                          make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                     __float_as_uint(v[4 * j + 3])));
    }
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void flush(const Params& p, uint32_t stage, int tile_m, int tile_n, int zb,
                                               const GemmShape& shp) {
#pragma unroll 1
    for (int b = 0; b < BLOCK_N / 32; ++b) {
      if (tile_n * BLOCK_N + b * 32 >= shp.N) break;
      ptx::tma_reduce_add_3d(&p.tm_out, stage + b * (kBlockM * 128), tile_n * BLOCK_N + b * 32, tile_m * kBlockM, zb);
    }
  }
};

// ---- out = act(acc + bias) -> activation dtype, optionally mirrored in fp32; optional relu-mask multiply
template <typename OutT>
struct EpiStore {
  static constexpr bool kFixup = true;
  static constexpr int kCtasPerSm = 2;
  struct Params {
    OutT* out;            // may be null
    float* out_f32;       // may be null
    const float* bias;    // [N] or null
    const OutT* mask;     // dgrad of a ReLU layer: zero the result where mask <= 0 (same layout as out); or null
    long ldo;             // row stride (elements) of out / out_f32 / mask
    long z_stride;        // batch stride (elements)
    int relu;
    OutT* out_cat;        // optional second output [M, 3N] = [hi | lo | hi], lo = round(v - hi): the K-concatenated
                          // split-precision operand of the next (small) GEMM, whose weight is laid out [w_hi | w_hi | w_lo]
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params& p, int m, int n0, int zb, int col0, int col1,
                                                  const GemmShape& shp) {
    if (p.mask != nullptr && m < shp.M && n0 + col0 < shp.N)
      prefetch_l2_span(p.mask + static_cast<long>(zb) * p.z_stride + static_cast<long>(m) * p.ldo + n0 + col0,
                       (col1 - col0) * static_cast<int>(sizeof(OutT)));
  }
  template <int BLOCK_N> struct Regs {};
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params&, Regs<BLOCK_N>&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const AccSource& acc, const Regs<BLOCK_N>&, int m, int n0,
                                             int zb, int col0, int col1, const GemmShape& shp) {
    const bool row_ok = m < shp.M;
    const long row_off = static_cast<long>(zb) * p.z_stride + static_cast<long>(m) * p.ldo;
#pragma unroll 1
    for (int c = col0; c < col1; c += 32) {
      if (n0 + c >= shp.N) break;
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the masked stores below
      float v[32];
      acc.template load<32>(c, v);
      const int nb = n0 + c;
      const bool full = (nb + 32 <= shp.N);
      if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || nb + i < shp.N) v[i] += __ldg(p.bias + nb + i);
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (row_ok) {
        const bool vec = full && ((p.ldo & 7) == 0);
        if (p.mask != nullptr) {
          const OutT* mk = p.mask + row_off + nb;
          if (vec) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float mv[8];
              Act8<OutT>::load(mk + 8 * g, mv);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 * g + i] = mv[i] > 0.f ? v[8 * g + i] : 0.f;
            }
          } else {
            for (int i = 0; i < 32; ++i)
              if (nb + i < shp.N) v[i] = to_f32(mk[i]) > 0.f ? v[i] : 0.f;
          }
        }
        if (p.out != nullptr) {
          OutT* o = p.out + row_off + nb;
          if (vec) {
#pragma unroll
            for (int g = 0; g < 4; ++g) Act8<OutT>::store(o + 8 * g, v + 8 * g);
          } else {
            for (int i = 0; i < 32; ++i)
              if (nb + i < shp.N) o[i] = from_f32<OutT>(v[i]);
          }
        }
        if (p.out_cat != nullptr) {
          OutT* oc = p.out_cat + static_cast<long>(zb) * 3 * p.z_stride + static_cast<long>(m) * 3 * shp.N + nb;
          for (int i = 0; i < 32; ++i) {
            if (nb + i < shp.N) {
              const OutT hi = from_f32<OutT>(v[i]);
              oc[i] = hi;
              oc[shp.N + i] = from_f32<OutT>(v[i] - to_f32(hi));
              oc[2 * shp.N + i] = hi;
            }
          }
        }
        if (p.out_f32 != nullptr) {
          float* o = p.out_f32 + row_off + nb;
          if (vec) {
#pragma unroll
            for (int g = 0; g < 4; ++g) Act8<float>::store(o + 8 * g, v + 8 * g);
          } else {
            for (int i = 0; i < 32; ++i)
              if (nb + i < shp.N) o[i] = v[i];
          }
        }
      }
    }
  }
};

// ---- out = act(acc + bias) -> activation dtype through staging + TMA store.  The direct version (EpiStore) writes one
// row per lane: 16 st.global.v4 per thread for a 128 x 256 bf16 tile, 32 LSU transactions per warp instruction, ~2 us per
// tile -- as long as the whole main loop when K <= 512 (LSTM input projections, 80-channel convolutions).
template <typename OutT>
struct EpiStoreTma {
  static constexpr bool kFixup = true;    // needs the complete sum: no split-K
  static constexpr bool kStaged = true;
  static constexpr int kCtasPerSm = 1;
  static constexpr int EB = sizeof(OutT);
  struct Params {
    CUtensorMap tm_out;   // OutT {N, M, batches}, box {128 / EB, 128, 1}
    const float* bias;    // [N] or null
    int relu;
    // optional BatchNorm statistics of the stored tile (train-mode nn.BatchNorm1d right after the convolution): per
    // column sum and sum of squares of the values AS STORED, added to stat_sums[half][2][N] (double), half = m / rows_half
    double* stat_sums;
    int rows_half;
  };
  template <int BLOCK_N>
  static __host__ __device__ constexpr int staging_bytes() { return BLOCK_N * EB * kBlockM; }
  // persistent kernel: 32-bit outputs of the widest tile (128 KB) are drained in two rounds of 128 columns (64 KB staging)
  template <int BLOCK_N>
  static __host__ __device__ constexpr int rounds() { return (EB == 4 && BLOCK_N == 256) ? 2 : 1; }
  // After the tile is staged (and its TMA store issued) every epilogue thread sums one column of the staged tile: the
  // statistics pass of the BatchNorm that follows costs no extra read of the activation tensor.  Column c of row r lives
  // in box c*EB/128 at chunk ((c*EB%128)/16) ^ (r&7): the 32 lanes of a warp read 32 consecutive columns of one row.
  template <int BLOCK_N>
  static __device__ __forceinline__ void after_stage(const Params& p, uint32_t stage, int et, int tile_m, int tile_n, int zb,
                                                     const GemmShape& shp) {
    after_stage_r<BLOCK_N>(p, stage, et, tile_m, tile_n, zb, shp, 0, BLOCK_N);
  }
  // the staged columns are [col_base, col_base + ncols) of the tile (one round)
  template <int BLOCK_N>
  static __device__ __forceinline__ void after_stage_r(const Params& p, uint32_t stage, int et, int tile_m, int tile_n, int,
                                                       const GemmShape& shp, int col_base, int ncols) {
    if (p.stat_sums == nullptr) return;
    constexpr int CPT = BLOCK_N / kEpilogueThreads;          // columns per thread (1 for BLOCK_N = 256)
    const int m0 = tile_m * kBlockM;
    const int nrows = min(kBlockM, shp.M - m0);
    const int half = m0 / p.rows_half;
#pragma unroll
    for (int cc = 0; cc < (CPT > 0 ? CPT : 1); ++cc) {
      const int c = (CPT > 0) ? et * CPT + cc : et;          // column inside the staged round
      if (c >= ncols) return;
      const int n = tile_n * BLOCK_N + col_base + c;
      if (n >= shp.N) continue;
      const int byte = c * EB;
      const uint32_t base = stage + static_cast<uint32_t>((byte >> 7) * (kBlockM * 128) + (byte & 15));
      const int chunk = (byte & 127) >> 4;
      float s = 0.f, q = 0.f;
      for (int r = 0; r < nrows; ++r) {
        const uint32_t a = base + static_cast<uint32_t>(r * 128 + ((chunk ^ (r & 7)) << 4));
        float v;
        if constexpr (EB == 2) {
          unsigned short u;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(a));
          v = bits16_to_f32<OutT>(u);
        } else {
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        }
        s += v;
        q = fmaf(v, v, q);
      }
      atomicAdd(p.stat_sums + (static_cast<long>(half) * 2 + 0) * shp.N + n, static_cast<double>(s));
      atomicAdd(p.stat_sums + (static_cast<long>(half) * 2 + 1) * shp.N + n, static_cast<double>(q));
    }
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N> struct Regs {};
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params&, Regs<BLOCK_N>&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run_staged(const Params& p, const AccSource& acc, const Regs<BLOCK_N>& regs, int row, int m,
                                                    int n0, int zb, int col0, int col1, const GemmShape& shp,
                                                    uint32_t stage) {
    run_staged_r<BLOCK_N>(p, acc, regs, row, m, n0, zb, col0, col1, shp, stage, 0);
  }
  // columns [col0, col1) of the tile go to staging column (c - col_base)
  template <int BLOCK_N>
  static __device__ __forceinline__ void run_staged_r(const Params& p, const AccSource& acc, const Regs<BLOCK_N>&, int row, int m,
                                                      int n0, int zb, int col0, int col1, const GemmShape& shp,
                                                      uint32_t stage, int col_base) {
#pragma unroll 1
    for (int c = col0; c < col1; c += 32) {
      if (n0 + c >= shp.N) break;
      __syncwarp();
      float v[32];
      acc.template load<32>(c, v);
      const int nb = n0 + c;
      if (p.bias != nullptr) {
        if (nb + 32 <= shp.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + i);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i < shp.N) v[i] += __ldg(p.bias + nb + i);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
#pragma unroll
      for (int j = 0; j < 32 * EB / 16; ++j) {
        const uint4 u = pack_chunk<OutT>(v + (16 / EB) * j);
        const int byte = (c - col_base) * EB + 16 * j;
        ptx::st_shared_v4(stage + static_cast<uint32_t>((byte >> 7) * (kBlockM * 128) + row * 128 +
                                                        ((((byte & 127) >> 4) ^ (row & 7)) << 4)),
                          u);
      }
    }
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void flush(const Params& p, uint32_t stage, int tile_m, int tile_n, int zb,
                                               const GemmShape& shp) {
    flush_r<BLOCK_N>(p, stage, tile_m, tile_n, zb, shp, 0, BLOCK_N);
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void flush_r(const Params& p, uint32_t stage, int tile_m, int tile_n, int zb,
                                                 const GemmShape& shp, int col_base, int ncols) {
    constexpr int GE = 128 / EB;
#pragma unroll 1
    for (int b = 0; b < ncols / GE; ++b) {
      const int n = tile_n * BLOCK_N + col_base + b * GE;
      if (n >= shp.N) break;
      ptx::tma_store_3d(&p.tm_out, stage + b * (kBlockM * 128), n, tile_m * kBlockM, zb);
    }
  }
};

// ---- out_f32 += acc  (split-K weight gradients; every split adds its own partial, no fix-up needed)
struct EpiAtomic {
  static constexpr bool kFixup = false;
  static constexpr int kCtasPerSm = 2;
  struct Params {
    float* out;
    long ldo;
    long z_stride;
    float alpha;   // every partial sum is scaled by alpha before it is added (1 / gradient scale of the fp16 mode)
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N> struct Regs {};
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params&, Regs<BLOCK_N>&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const AccSource& acc, const Regs<BLOCK_N>&, int m, int n0,
                                             int zb, int col0, int col1, const GemmShape& shp) {
    if (!acc.has_acc) return;
    const bool row_ok = m < shp.M;
    float* row = p.out + static_cast<long>(zb) * p.z_stride + static_cast<long>(m) * p.ldo;
#pragma unroll 1
    for (int c = col0; c < col1; c += 32) {
      if (n0 + c >= shp.N) break;
      __syncwarp();
      float v[32];
      acc.template load<32>(c, v);
      if (p.alpha != 1.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
      }
      const int nb = n0 + c;
      if (row_ok) {
        if (nb + 32 <= shp.N && (p.ldo & 3) == 0) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            ptx::red_add_v4(row + nb + 4 * g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        } else {
          for (int i = 0; i < 32; ++i)
            if (nb + i < shp.N) atomicAdd(row + nb + i, v[i]);
        }
      }
    }
  }
};

// ---- LSTM cell forward.  Gate-interleaved layout: column 4*u + g of the [rows, 4H] pre-activation is gate g (i,f,g,o)
// of hidden unit u (weight rows are permuted accordingly by prep_lstm_weight), so a thread's 8 units are 32 contiguous
// columns: one tcgen05.ld.x32 and one 64-byte run of xproj / gates.  a = acc + xproj;  c = s(f) c_prev + s(i) tanh(g);
// h = s(o) tanh(c).  Saves the activated gates (in place of xproj) and c for the backward pass.
template <typename ActT>
struct EpiLstmFwd {
  static constexpr bool kFixup = true;
  static constexpr int kCtasPerSm = 2;
  struct Params {
    const ActT* xproj;   // [rows, ldx] at time t (gate-interleaved), + zb * z_x
    const float* c_prev; // [rows, ldc] at the previous time step (null on the first step)
    float* c_out;        // [rows, ldc] at time t
    ActT* h_out;         // [rows, ldh] at time t (+ zb * z_h: direction offset in the concat output)
    ActT* gates;         // [rows, ldx] at time t, activated gates (same layout), + zb * z_x
    long ldx, ldc, ldh;
    // element offsets applied when zb == 1 (reverse direction of a bidirectional layer: other weights,
    // other half of the concat output, and a different time index)
    long z_x, z_c_prev, z_c_out, z_h;
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params& p, int m, int n0, int zb, int col0, int col1,
                                                  const GemmShape& shp) {
    if (m >= shp.M) return;
    prefetch_l2_span(p.xproj + zb * p.z_x + static_cast<long>(m) * p.ldx + n0 + col0,
                     (col1 - col0) * static_cast<int>(sizeof(ActT)));
    if (p.c_prev)
      prefetch_l2_span(p.c_prev + zb * p.z_c_prev + static_cast<long>(m) * p.ldc + (n0 + col0) / 4, (col1 - col0));
  }
  // everything the cell needs besides the accumulator, requested while the main loop is still running
  template <int BLOCK_N>
  struct Regs {
    static constexpr int NCH = BLOCK_N / 2 / 32;   // 32-column (8-unit) chunks owned by this thread
    typename Act8<ActT>::raw_t x[NCH][4];
    typename Act8<float>::raw_t c[NCH];
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params& p, Regs<BLOCK_N>& r, int m, int n0, int zb, int col0,
                                                 int col1, const GemmShape& shp) {
    if (m >= shp.M) return;
    const ActT* xp = p.xproj + zb * p.z_x + static_cast<long>(m) * p.ldx + n0 + col0;
    const float* cp = p.c_prev ? p.c_prev + zb * p.z_c_prev + static_cast<long>(m) * p.ldc + (n0 + col0) / 4 : nullptr;
#pragma unroll
    for (int k = 0; k < Regs<BLOCK_N>::NCH; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r.x[k][j] = Act8<ActT>::load_raw(xp + 32 * k + 8 * j);
      if (cp) r.c[k] = Act8<float>::load_raw(cp + 8 * k);
    }
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const AccSource& acc, const Regs<BLOCK_N>& r, int m, int n0,
                                             int zb, int col0, int col1, const GemmShape& shp) {
    const bool row_ok = m < shp.M;
    ActT* gs = p.gates + zb * p.z_x + static_cast<long>(m) * p.ldx + n0;
    float* co = p.c_out + zb * p.z_c_out + static_cast<long>(m) * p.ldc + n0 / 4;
    ActT* ho = p.h_out + zb * p.z_h + static_cast<long>(m) * p.ldh + n0 / 4;
#pragma unroll
    for (int k = 0; k < Regs<BLOCK_N>::NCH; ++k) {   // 32 columns = 8 hidden units x 4 gates
      const int c = col0 + 32 * k;
      __syncwarp();
      float a[32];
      acc.template load<32>(c, a);
      if (row_ok) {
        float x[32], cprev[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<ActT>::unpack(r.x[k][j], x + 8 * j);
        if (p.c_prev) Act8<float>::unpack(r.c[k], cprev);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cprev[i] = 0.f;
        }
        float cn[8], hn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ig = GateMath<ActT>::sig(a[4 * i] + x[4 * i]), fg = GateMath<ActT>::sig(a[4 * i + 1] + x[4 * i + 1]);
          const float gg = GateMath<ActT>::tnh(a[4 * i + 2] + x[4 * i + 2]), og = GateMath<ActT>::sig(a[4 * i + 3] + x[4 * i + 3]);
          a[4 * i] = ig; a[4 * i + 1] = fg; a[4 * i + 2] = gg; a[4 * i + 3] = og;
          cn[i] = fg * cprev[i] + ig * gg;
          hn[i] = og * GateMath<ActT>::tnh(cn[i]);
        }
        Act8<float>::store(co + c / 4, cn);
        Act8<ActT>::store(ho + c / 4, hn);
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<ActT>::store(gs + c + 8 * j, a + 8 * j);
      }
    }
  }
};

// ---- LSTM cell forward with staged stores: same arithmetic and the same (preloaded) global reads as EpiLstmFwd, but
// the three outputs of the step (activated gates, c, h) leave through shared-memory staging and TMA stores.  With one
// row per lane the direct version issues 28 16-byte stores per thread at BLOCK_N = 256, each touching 32 different
// 128-byte lines per warp: ~7000 LSU transactions per CTA and step, the largest item of the step's critical path.
template <typename ActT>
struct EpiLstmFwdTma : EpiLstmFwd<ActT> {
  using Base = EpiLstmFwd<ActT>;
  static constexpr bool kStaged = true;
  static constexpr int kCtasPerSm = 1;
  static constexpr int EB = sizeof(ActT);
  struct Params : Base::Params {
    CUtensorMap tm_g;   // gates  ActT {D*4H, T, rows}, box {128/EB, 1, 128}
    CUtensorMap tm_c;   // c      fp32 {D*H,  T, rows}, box {32, 1, 128}
    CUtensorMap tm_h;   // h      ActT {D*H,  T, rows}, box {min(128, BLOCK_N/4*EB)/EB, 1, 128}
    int t[2];           // time index per direction
    int H;
  };
  template <int BLOCK_N> static __host__ __device__ constexpr int g_boxes() { return BLOCK_N * EB / 128; }
  template <int BLOCK_N> static __host__ __device__ constexpr int c_boxes() { return BLOCK_N / 128; }
  template <int BLOCK_N> static __host__ __device__ constexpr int h_row_bytes() { return BLOCK_N / 4 * EB; }
  template <int BLOCK_N> static __host__ __device__ constexpr int h_boxes() { return h_row_bytes<BLOCK_N>() >= 128 ? h_row_bytes<BLOCK_N>() / 128 : 1; }
  template <int BLOCK_N> static __host__ __device__ constexpr int h_bytes() { return kBlockM * h_row_bytes<BLOCK_N>(); }
  template <int BLOCK_N>
  static __host__ __device__ constexpr int staging_bytes() {
    return (g_boxes<BLOCK_N>() + c_boxes<BLOCK_N>()) * kBlockM * 128 + h_bytes<BLOCK_N>();
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void after_stage(const Params&, uint32_t, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run_staged(const Params& p, const AccSource& acc, const typename Base::template Regs<BLOCK_N>& r,
                                                    int row, int m, int n0, int zb, int col0, int col1, const GemmShape& shp,
                                                    uint32_t stage) {
    const bool row_ok = m < shp.M;
    const uint32_t sg = stage, sc = sg + g_boxes<BLOCK_N>() * kBlockM * 128, sh = sc + c_boxes<BLOCK_N>() * kBlockM * 128;
    auto box_addr = [&](uint32_t base, int byte_in_row) {   // 16-byte chunk of this row in a [boxes][128 rows][128 B] tile
      return base + static_cast<uint32_t>((byte_in_row >> 7) * (kBlockM * 128) + row * 128 +
                                          ((((byte_in_row & 127) >> 4) ^ (row & 7)) << 4));
    };
#pragma unroll
    for (int k = 0; k < Base::template Regs<BLOCK_N>::NCH; ++k) {
      const int c = col0 + 32 * k;
      __syncwarp();
      float a[32];
      acc.template load<32>(c, a);
      float x[32], cprev[8];
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Act8<ActT>::unpack(r.x[k][j], x + 8 * j);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
      }
      if (row_ok && p.c_prev) Act8<float>::unpack(r.c[k], cprev);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) cprev[i] = 0.f;
      }
      float cn[8], hn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ig = GateMath<ActT>::sig(a[4 * i] + x[4 * i]), fg = GateMath<ActT>::sig(a[4 * i + 1] + x[4 * i + 1]);
        const float gg = GateMath<ActT>::tnh(a[4 * i + 2] + x[4 * i + 2]), og = GateMath<ActT>::sig(a[4 * i + 3] + x[4 * i + 3]);
        a[4 * i] = ig; a[4 * i + 1] = fg; a[4 * i + 2] = gg; a[4 * i + 3] = og;
        cn[i] = fg * cprev[i] + ig * gg;
        hn[i] = og * GateMath<ActT>::tnh(cn[i]);
      }
      // gates: 32 columns = 32*EB bytes at byte c*EB of the row
#pragma unroll
      for (int j = 0; j < 32 * EB / 16; ++j) {
        const uint4 u = pack_chunk<ActT>(a + (16 / EB) * j);
        ptx::st_shared_v4(box_addr(sg, c * EB + 16 * j), u);
      }
      // c: 8 units fp32 = 32 bytes at byte c (= 4 * (c / 4)) of the row
#pragma unroll
      for (int j = 0; j < 2; ++j)
        ptx::st_shared_v4(box_addr(sc, c + 16 * j),
                          make_uint4(__float_as_uint(cn[4 * j]), __float_as_uint(cn[4 * j + 1]), __float_as_uint(cn[4 * j + 2]),
                                     __float_as_uint(cn[4 * j + 3])));
      // h: 8 units = 8*EB bytes at byte (c / 4) * EB of the row
#pragma unroll
      for (int j = 0; j < 8 * EB / 16; ++j) {
        const uint4 u = pack_chunk<ActT>(hn + (16 / EB) * j);
        const int byte = (c / 4) * EB + 16 * j;
        if constexpr (h_row_bytes<BLOCK_N>() >= 128) ptx::st_shared_v4(box_addr(sh, byte), u);
        else ptx::st_shared_v4(sh + row * h_row_bytes<BLOCK_N>() + byte, u);   // narrow tile: unswizzled rows
      }
    }
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void flush(const Params& p, uint32_t stage, int tile_m, int tile_n, int zb,
                                               const GemmShape& shp) {
    const uint32_t sg = stage, sc = sg + g_boxes<BLOCK_N>() * kBlockM * 128, sh = sc + c_boxes<BLOCK_N>() * kBlockM * 128;
    const int t = p.t[zb], m0 = tile_m * kBlockM, n0 = tile_n * BLOCK_N;
    constexpr int GE = 128 / EB;   // elements per 128-byte box row
#pragma unroll
    for (int b = 0; b < h_boxes<BLOCK_N>(); ++b)   // h first: the next step's GEMM waits for it
      ptx::tma_store_3d(&p.tm_h, sh + b * (kBlockM * 128), zb * p.H + n0 / 4 + b * GE, t, m0);
#pragma unroll
    for (int b = 0; b < g_boxes<BLOCK_N>(); ++b)
      ptx::tma_store_3d(&p.tm_g, sg + b * (kBlockM * 128), zb * 4 * p.H + n0 + b * GE, t, m0);
#pragma unroll
    for (int b = 0; b < c_boxes<BLOCK_N>(); ++b)
      ptx::tma_store_3d(&p.tm_c, sc + b * (kBlockM * 128), zb * p.H + n0 / 4 + b * 32, t, m0);
  }
};

// ---- LSTM cell backward for one time step.  acc = dh_rec[m, unit] = da_{t+1} . W_hh (absent on the last
// step).  dh = dh_out + acc; emits da_t (natural torch gate order i,f,g,o: column g*H + unit) and the new
// dc carry.  The saved activated gates are in the forward's gate-interleaved layout (column 4*u + g).
template <typename ActT>
struct EpiLstmBwd {
  static constexpr bool kFixup = true;
  static constexpr int kCtasPerSm = 1;   // register-heavy epilogue (two chunks of operands in flight); grids are small
  struct Params {
    const ActT* dh_out;  // [rows, ldh] grad wrt this layer's output at time t (+ zb * z_h)
    const ActT* gates;   // activated gates at time t, gate-interleaved
    const float* c_t;    // [rows, ldc] at time t
    const float* c_prev; // at the previous time step or null
    float* dc;           // [rows, H] carry, in/out (+ zb * z_dc)
    ActT* da;            // [rows, lda] at time t, natural gate order (+ zb * z_a)
    long ldh, ldx, ldc, lda;
    long z_h, z_x, z_c, z_c_prev, z_dc, z_a;  // element offsets applied when zb == 1 (reverse direction)
    int H, dc_zero;                           // dc_zero: treat incoming carry as zero (first processed step)
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void prefetch(const Params& p, int m, int n0, int zb, int col0, int col1,
                                                  const GemmShape& shp) {
    if (m >= shp.M || n0 + col0 >= shp.N) return;
    const long r = m;
    const int u = n0 + col0, nu = col1 - col0;
    prefetch_l2_span(p.dh_out + zb * p.z_h + r * p.ldh + u, nu * static_cast<int>(sizeof(ActT)));
    prefetch_l2_span(p.gates + zb * p.z_x + r * p.ldx + 4 * u, 4 * nu * static_cast<int>(sizeof(ActT)));
    prefetch_l2_span(p.c_t + zb * p.z_c + r * p.ldc + u, nu * 4);
    if (p.c_prev) prefetch_l2_span(p.c_prev + zb * p.z_c_prev + r * p.ldc + u, nu * 4);
    if (!p.dc_zero) prefetch_l2_span(p.dc + zb * p.z_dc + r * p.H + u, nu * 4);
  }
  template <int BLOCK_N> struct Regs {};
  template <int BLOCK_N>
  static __device__ __forceinline__ void preload(const Params&, Regs<BLOCK_N>&, int, int, int, int, int, const GemmShape&) {}
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const AccSource& accs, const Regs<BLOCK_N>&, int m, int n0,
                                             int zb, int col0, int col1, const GemmShape& shp) {
    constexpr int CH = 2;   // 8-unit chunks per iteration (BLOCK_N / 2 columns per thread is always >= 32)
    const bool row_ok = m < shp.M;
    const long r = m;
    const ActT* dho = p.dh_out + zb * p.z_h + r * p.ldh;
    const ActT* gs = p.gates + zb * p.z_x + r * p.ldx;
    const float* ct = p.c_t + zb * p.z_c + r * p.ldc;
    const float* cp = p.c_prev ? p.c_prev + zb * p.z_c_prev + r * p.ldc : nullptr;
    float* dcp = p.dc + zb * p.z_dc + r * p.H;
    ActT* da = p.da + zb * p.z_a + r * p.lda;
#pragma unroll 1
    for (int c = col0; c < col1; c += 8 * CH) {
      if (n0 + c >= shp.N) break;
      // all global operands of both chunks first (latency-bound epilogue: maximise loads in flight)
      float dh[CH][8], g4[CH][32], cc[CH][8], cpv[CH][8], dc[CH][8];
      if (row_ok) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const int u = n0 + c + 8 * k;
          Act8<ActT>::load(dho + u, dh[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) Act8<ActT>::load(gs + 4 * u + 8 * j, g4[k] + 8 * j);
          Act8<float>::load(ct + u, cc[k]);
          if (cp) Act8<float>::load(cp + u, cpv[k]);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) cpv[k][i] = 0.f;
          }
          if (p.dc_zero) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dc[k][i] = 0.f;
          } else {
            Act8<float>::load(dcp + u, dc[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int u = n0 + c + 8 * k;
        __syncwarp();
        float acc[8];
        accs.template load<8>(c + 8 * k, acc);
        if (row_ok) {
          float dai[8], daf[8], dag[8], dao[8], dcn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ig = g4[k][4 * i], fg = g4[k][4 * i + 1], gg = g4[k][4 * i + 2], og = g4[k][4 * i + 3];
            const float tc = GateMath<ActT>::tnh(cc[k][i]);
            const float dht = dh[k][i] + acc[i];
            const float dct = dc[k][i] + dht * og * (1.f - tc * tc);
            dao[i] = dht * tc * og * (1.f - og);
            dai[i] = dct * gg * ig * (1.f - ig);
            dag[i] = dct * ig * (1.f - gg * gg);
            daf[i] = dct * cpv[k][i] * fg * (1.f - fg);
            dcn[i] = dct * fg;
          }
          Act8<float>::store(dcp + u, dcn);
          Act8<ActT>::store(da + 0 * p.H + u, dai);
          Act8<ActT>::store(da + 1 * p.H + u, daf);
          Act8<ActT>::store(da + 2 * p.H + u, dag);
          Act8<ActT>::store(da + 3 * p.H + u, dao);
        }
      }
    }
  }
};

}  // namespace dvae
