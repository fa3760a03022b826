// C-ABI launchers for every dense contraction on the Disentangled-VAE hot path.  All of them drive the
// single tcgen05 kernel in tc_gemm.cuh; what changes is the tensor-map geometry, the operand majors and
// the epilogue:
//
//   Linear   fwd : out[M,N]  = x[M,K] . w[N,K]^T  (+bias, relu)          A K-major , B K-major
//   Linear   dgrad: dx[M,K]  = dy[M,N] . w[N,K]                           A K-major , B MN-major
//   Linear   wgrad: dw[N,K] += dy[M,N]^T . x[M,K]                         A MN-major, B MN-major, split-K
//   Conv1d-5 fwd / dgrad / wgrad on channels-last [R,T,C] (implicit GEMM, taps = shifted TMA boxes)
//   LSTM     fwd step  : gates = xproj_t + h_{t-1} . W_hh^T  -> fused cell epilogue
//   LSTM     bwd step  : dh_rec = da_{t+1} . W_hh             -> fused cell-backward epilogue
//
// Reference call sites (model/disentangled_vae.py): Conv1d :154-160,:178-189,:54-78; LSTM :163,:172,:193;
// Linear :165-171,:194.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "host_common.h"
#include "tc_gemm.cuh"

namespace dvae {

static int env_int(const char* name, int dflt);
static int g_background_flag();
template <typename AT>
constexpr int kIsF16 = std::is_same<AT, __half>::value ? 1 : 0;

template <int BN>
struct Stages {
  static constexpr int value = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
};

// stages that fit 227 KB with the 256-row (two-accumulator) tile
template <int BN>
struct Stages2 {
  static constexpr int value = (BN == 256) ? 3 : (BN == 128 ? 4 : 5);
};

template <int BN, bool A_MN, bool B_MN, int EB, class Epi, int MT = 1, int ST = (MT == 2 ? Stages2<BN>::value : Stages<BN>::value)>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const OperandWalk& wa, const OperandWalk& wb,
                       const GemmShape& shp, const typename Epi::Params& ep, dim3 grid, cudaStream_t stream) {
  if (Epi::kFixup && shp.splits > 1 && (shp.splitk_ws == nullptr || shp.tickets == nullptr)) {
    set_last_error("split-K with a full-sum epilogue needs a fix-up workspace and tickets");
    return 1;
  }
  auto kern = tc_gemm_kernel<BN, ST, A_MN, B_MN, EB, Epi, MT>;
  constexpr int smem = gemm_smem_bytes<BN, ST, MT>();
  static bool configured = false;  // one flag per template instantiation
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
  static const int use_pdl = env_int("DVAE_PDL", 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  DVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, wa, wb, shp, ep));
  return 0;
}

// CTA-pair launch (tcgen05 cta_group::2): clusters of two consecutive M tiles; every CTA stages half of the B tile, so the
// ring is as deep as 227 KB allow.  The B tensor map of a K-major operand must have been encoded with BN / 2 rows per box.
template <int BN, bool A_MN, bool B_MN, int EB, class Epi>
static int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const OperandWalk& wa, const OperandWalk& wb,
                            const GemmShape& shp, const typename Epi::Params& ep, dim3 grid, cudaStream_t stream) {
  constexpr int STAGE = 128 * 128 + BN * 128 / 2;
  constexpr int ST_FIT = (227 * 1024 - 1280) / STAGE;
  constexpr int ST = ST_FIT > 10 ? 10 : ST_FIT;
  constexpr int smem = ST * STAGE + 1024 + 256;
  if ((Epi::kFixup && shp.splits > 1) || grid.x % 2 != 0) {
    set_last_error("CTA-pair GEMM needs an even number of M tiles and no fix-up split-K");
    return 1;
  }
  auto kern = tc_gemm_kernel<BN, ST, A_MN, B_MN, EB, Epi, 1, 2>;
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
  static const int use_pdl = env_int("DVAE_PDL", 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, wa, wb, shp, ep));
  return 0;
}

// Persistent launch: one CTA per SM looping over all (m, n, batch x split) tiles.  Used when there are at least two
// tiles per SM; smaller problems (LSTM steps, M = 1024 linears) keep the one-tile-per-CTA kernel.
template <int BN, bool A_MN, bool B_MN, int EB, class Epi, int CG = 1>
static int launch_gemm_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const OperandWalk& wa, const OperandWalk& wb,
                                  const GemmShape& shp, const typename Epi::Params& ep, dim3 grid, cudaStream_t stream) {
  static_assert(BN <= 256, "");
  constexpr int STAGING = persistent_staging_bytes<Epi, BN>();
  constexpr int STAGE = 128 * 128 + BN * 128 / CG;   // CG = 2 (CTA pairs): every CTA stages half of the B tile
  constexpr int ST_FULL = CG == 2 ? 10 : Stages<BN>::value + (BN == 128 ? 3 : (BN == 64 ? 4 : 0));   // 1 CTA / SM: use the whole 227 KB
  constexpr int ST_FIT = (227 * 1024 - 1280 - STAGING) / STAGE;        // what is left next to a staging tile
  constexpr int ST = ST_FIT < ST_FULL ? ST_FIT : ST_FULL;
  static_assert(ST >= 2, "persistent GEMM: ring too shallow");
  auto kern = tc_gemm_persistent_kernel<BN, ST, A_MN, B_MN, EB, Epi, CG>;
  constexpr int smem = ST * STAGE + 1024 + 256 + STAGING;
  static_assert(smem <= 227 * 1024, "persistent GEMM shared memory");
  static bool configured = false;
  if (!configured) {
    DVAE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_m = grid.x, tiles_n = grid.y;
  const long num_tiles = static_cast<long>(grid.x) * grid.y * grid.z;
  if (num_tiles == 0) return 0;
  if (shp.num_kb <= 0 || (Epi::kFixup && shp.splits > 1)) {
    set_last_error("persistent GEMM needs K > 0 and supports split-K only for accumulate epilogues");
    return 1;
  }
  if (CG == 2 && tiles_m % 2 != 0) {
    set_last_error("CTA-pair GEMM needs an even number of M tiles");
    return 1;
  }
  static const int use_pdl = env_int("DVAE_PDL", 1);
  long ctas = num_tiles < num_sms() ? num_tiles : num_sms();
  if (CG == 2) ctas &= ~1L;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, wa, wb, shp, ep, tiles_m, tiles_n, static_cast<int>(num_tiles)));
  return 0;
}

// CTA pairs (tcgen05 cta_group::2) for the persistent 128 x 256 kernels: DVAE_GEMM_PAIR (default 1) when the M tiles pair up.
// K-major B operands must then be encoded with BN / 2 rows per TMA box.
static bool gemm_pair(int tiles_m, int bn) {
  static const int want = env_int("DVAE_GEMM_PAIR", 1);
  return want != 0 && !g_background_flag() && bn == 256 && tiles_m % 2 == 0;
}

// Background mode (dvae_set_background): GEMMs that run on a side stream next to latency-critical kernels keep the
// one-tile-per-CTA kernel (96 KB CTAs that co-reside with the LSTM step kernels) instead of taking whole SMs.
static thread_local int g_background = 0;
static int g_background_flag() { return g_background; }

// DVAE_GEMM_PERSISTENT = 0 disables, 1 forces (tests); default: when the grid has >= 2 tiles per SM
static bool want_persistent(dim3 grid) {
  const char* v = getenv("DVAE_GEMM_PERSISTENT");
  if (v && *v == '0') return false;
  if (v && *v == '1') return true;
  if (g_background) return false;
  return static_cast<long>(grid.x) * grid.y * grid.z >= 2L * num_sms();
}

static OperandWalk zero_walk() {
  OperandWalk w;
  for (int d = 0; d < 3; ++d) w.base[d] = w.per_j[d] = w.per_tap[d] = w.per_box[d] = w.per_tile[d] = w.per_z[d] = 0;
  return w;
}

static int pick_bn(int N) { return N <= 64 ? 64 : (N <= 128 ? 128 : 256); }

// 256-row CTA tiles (two accumulators sharing the B stage) when the grid still fills the chip; DVAE_GEMM_MT=1|2 forces.
static int pick_mt(long M, int n_tiles_total) {
  const char* v = getenv("DVAE_GEMM_MT");
  if (v && *v == '1') return 1;
  if (v && *v == '2') return 2;
  (void)M; (void)n_tiles_total;
  return 1;   // measured: the 256-row tile is not faster (profiles/): the persistent 128-row kernel is the default
}

// Staged (TMA) output stores for the persistent kernel: eligible when the only output is the activation tensor.
// DVAE_GEMM_TMA_STORE: 0 never, 1 whenever eligible, default 2: when the main loop is short (<= 16 k-blocks per tile), i.e.
// when the direct row-per-lane stores would take as long as the MMAs.
static bool use_tma_store(const void* out, const float* out_f32, const void* mask, long ldo, int elem_bytes, int num_kb) {
  static const int mode = env_int("DVAE_GEMM_TMA_STORE", 2);
  if (mode == 0 || out == nullptr || out_f32 != nullptr || mask != nullptr || (ldo * elem_bytes) % 16 != 0) return false;
  static const int max_kb = env_int("DVAE_GEMM_TMA_STORE_MAX_KB", 16);
  return mode == 1 || num_kb <= max_kb;
}

static bool tma_store_possible(const void* out, const float* out_f32, const void* mask, long ldo, int elem_bytes) {
  static const int mode = env_int("DVAE_GEMM_TMA_STORE", 2);
  return mode != 0 && out != nullptr && out_f32 == nullptr && mask == nullptr && (ldo * elem_bytes) % 16 == 0;
}

// ------------------------------------------------------------------------------------ Linear
// b_terms = 2: w is the split-precision weight [N, 2K] = [w_hi | w_lo] and the product is x . w_hi^T + x . w_lo^T: the
// k-loop runs twice over the same A tiles ("tap" 1 moves only the B coordinate by K columns).  out_cat: see EpiStore.
template <typename AT>
static int linear_fwd_t(const AT* x, long ldx, const AT* w, const float* bias, AT* out, float* out_f32, long ldo, int M,
                        int N, int K, int relu, int bn_override, cudaStream_t st, int b_terms = 1, AT* out_cat = nullptr) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  CUtensorMap ta, tb;
  int BN = bn_override ? bn_override : pick_bn(N);
  // a long-K layer with few rows (enc_linear: 1024 x 2048 x 8192 = 64 tiles of 128 x 256) leaves most SMs idle: halve the tile
  if (!bn_override && BN == 256 && static_cast<long>(ceil_div(M, 128)) * ceil_div(N, 256) * 10 < 6L * num_sms()) BN = 128;
  const long KW = static_cast<long>(b_terms) * K;   // row length of w
  if (int e = encode_map3(&ta, x, EB, K, M, 1, ldx * EB, (uint64_t)M * ldx * EB, BK, 128, 1)) return e;
  if (int e = encode_map3(&tb, w, EB, KW, N, 1, (uint64_t)KW * EB, (uint64_t)N * KW * EB, BK, BN, 1)) return e;
  OperandWalk wa = zero_walk(), wb = zero_walk();
  wa.per_j[0] = BK; wa.per_tile[1] = 128;
  wb.per_j[0] = BK; wb.per_tile[1] = BN; wb.per_tap[0] = K;
  GemmShape shp{M, N, b_terms * ceil_div(K, BK), ceil_div(K, BK), 1};
  shp.f16 = kIsF16<AT>;
  typename EpiStore<AT>::Params ep{out, out_f32, bias, nullptr, ldo, 0, relu, out_cat};
  DVAE_REQUIRE(b_terms == 1 || b_terms == 2, "b_terms must be 1 or 2");
  const bool split3 = out_cat != nullptr;
  const int mt = pick_mt(M, ceil_div(N, BN));
  dim3 grid(ceil_div(M, 128 * mt), ceil_div(N, BN), 1);
  if (mt == 2) {
    switch (BN) {
      case 64: return launch_gemm<64, false, false, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm<128, false, false, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm<256, false, false, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  if (want_persistent(grid) && mt == 1 && gemm_pair(grid.x, BN)) {   // CTA pairs: half of the weight tile per CTA
    if (int e = encode_map3(&tb, w, EB, KW, N, 1, (uint64_t)KW * EB, (uint64_t)N * KW * EB, BK, BN / 2, 1)) return e;
    if (!split3 && use_tma_store(out, out_f32, nullptr, ldo, EB, shp.num_kb)) {
      typename EpiStoreTma<AT>::Params et;
      if (int e = encode_map3(&et.tm_out, out, EB, N, M, 1, (uint64_t)ldo * EB, (uint64_t)M * ldo * EB, BK, 128, 1)) return e;
      et.bias = bias; et.relu = relu; et.stat_sums = nullptr; et.rows_half = 0;
      return launch_gemm_persistent<256, false, false, EB, EpiStoreTma<AT>, 2>(ta, tb, wa, wb, shp, et, grid, st);
    }
    return launch_gemm_persistent<256, false, false, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  }
  if (want_persistent(grid)) {
    if (!split3 && use_tma_store(out, out_f32, nullptr, ldo, EB, shp.num_kb)) {
      typename EpiStoreTma<AT>::Params et;
      if (int e = encode_map3(&et.tm_out, out, EB, N, M, 1, (uint64_t)ldo * EB, (uint64_t)M * ldo * EB, BK, 128, 1)) return e;
      et.bias = bias; et.relu = relu; et.stat_sums = nullptr; et.rows_half = 0;
      switch (BN) {
        case 64: return launch_gemm_persistent<64, false, false, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
        case 128: return launch_gemm_persistent<128, false, false, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
        default: return launch_gemm_persistent<256, false, false, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
      }
    }
    switch (BN) {
      case 64: return launch_gemm_persistent<64, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm_persistent<128, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm_persistent<256, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  switch (BN) {
    case 64: return launch_gemm<64, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    case 128: return launch_gemm<128, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    default: return launch_gemm<256, false, false, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
  }
}

template <typename AT>
static int linear_dgrad_t(const AT* dy, long lddy, const AT* w, AT* dx, float* dx_f32, const AT* relu_mask, long ldx,
                          int M, int N, int K, int bn_override, cudaStream_t st) {
  // dx[M,K] = dy[M,N] . w[N,K]   (GEMM output dim = K, reduction = N)
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  CUtensorMap ta, tb;
  const int BN = bn_override ? bn_override : pick_bn(K);
  if (int e = encode_map3(&ta, dy, EB, N, M, 1, lddy * EB, (uint64_t)M * lddy * EB, BK, 128, 1)) return e;
  if (int e = encode_map3(&tb, w, EB, K, N, 1, (uint64_t)K * EB, (uint64_t)N * K * EB, BK, BK, 1, true)) return e;
  OperandWalk wa = zero_walk(), wb = zero_walk();
  wa.per_j[0] = BK; wa.per_tile[1] = 128;
  wb.per_j[1] = BK; wb.per_box[0] = BK; wb.per_tile[0] = BN;
  GemmShape shp{M, K, ceil_div(N, BK), ceil_div(N, BK), 1};
  shp.f16 = kIsF16<AT>;
  typename EpiStore<AT>::Params ep{dx, dx_f32, nullptr, relu_mask, ldx, 0, 0};
  const int mt = pick_mt(M, ceil_div(K, BN));
  dim3 grid(ceil_div(M, 128 * mt), ceil_div(K, BN), 1);
  if (mt == 2) {
    switch (BN) {
      case 64: return launch_gemm<64, false, true, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm<128, false, true, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm<256, false, true, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  if (want_persistent(grid) && mt == 1 && gemm_pair(grid.x, BN)) {   // B is MN-major: the pair splits its boxes
    if (use_tma_store(dx, dx_f32, relu_mask, ldx, EB, shp.num_kb)) {
      typename EpiStoreTma<AT>::Params et;
      if (int e = encode_map3(&et.tm_out, dx, EB, K, M, 1, (uint64_t)ldx * EB, (uint64_t)M * ldx * EB, BK, 128, 1)) return e;
      et.bias = nullptr; et.relu = 0; et.stat_sums = nullptr; et.rows_half = 0;
      return launch_gemm_persistent<256, false, true, EB, EpiStoreTma<AT>, 2>(ta, tb, wa, wb, shp, et, grid, st);
    }
    return launch_gemm_persistent<256, false, true, EB, EpiStore<AT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  }
  if (want_persistent(grid)) {
    if (use_tma_store(dx, dx_f32, relu_mask, ldx, EB, shp.num_kb)) {
      typename EpiStoreTma<AT>::Params et;
      if (int e = encode_map3(&et.tm_out, dx, EB, K, M, 1, (uint64_t)ldx * EB, (uint64_t)M * ldx * EB, BK, 128, 1)) return e;
      et.bias = nullptr; et.relu = 0; et.stat_sums = nullptr; et.rows_half = 0;
      switch (BN) {
        case 64: return launch_gemm_persistent<64, false, true, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
        case 128: return launch_gemm_persistent<128, false, true, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
        default: return launch_gemm_persistent<256, false, true, EB, EpiStoreTma<AT>>(ta, tb, wa, wb, shp, et, grid, st);
      }
    }
    switch (BN) {
      case 64: return launch_gemm_persistent<64, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm_persistent<128, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm_persistent<256, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  switch (BN) {
    case 64: return launch_gemm<64, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    case 128: return launch_gemm<128, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    default: return launch_gemm<256, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
  }
}

static int pick_splits(int tiles, int num_kb) {
  // fill ~2 CTAs per SM, never fewer than 4 k-blocks per split
  int want = ceil_div(2 * num_sms(), tiles > 0 ? tiles : 1);
  int cap = num_kb / 4 > 0 ? num_kb / 4 : 1;
  int s = want < cap ? want : cap;
  return s < 1 ? 1 : s;
}

// Split-K factor for a persistent launch: the work items (tiles x splits) are dealt round-robin to one CTA per SM, so what
// matters is how full the last wave is (40 tiles x 8 splits = 2.16 waves runs as 3), then fewer splits (less fp32
// reduction traffic).  At least 8 k-blocks per split.
static int pick_splits_persistent(int tiles, int num_kb) {
  const int sms = num_sms();
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= 24 && num_kb / s >= 8; ++s) {
    const double waves = static_cast<double>(tiles) * s / sms;
    const double eff = waves / static_cast<double>(ceil_div(static_cast<long>(tiles) * s, sms));
    const double score = eff - 0.004 * s;
    if (score > best_score) { best_score = score; best = s; }
  }
  return best;
}

template <typename AT>
static int linear_wgrad_t(const AT* dy, long lddy, const AT* x, long ldx, float* dw, long lddw, int M, int N, int K,
                          float alpha, cudaStream_t st) {
  // dw[N,K] += dy[M,N]^T . x[M,K]   (reduction over the M rows; both operands MN-major)
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  CUtensorMap ta, tb;
  // wide (128 x 256) persistent tiles with a wave-aware split-K when the launch may take whole SMs (see conv5_wgrad_t)
  static const int wide_env = env_int("DVAE_WGRAD_WIDE", 1);
  const bool wide = wide_env != 0 && !g_background && K % 256 == 0;
  const int BN = wide ? 256 : (K <= 64 ? 64 : 128);
  if (int e = encode_map3(&ta, dy, EB, N, M, 1, lddy * EB, (uint64_t)M * lddy * EB, BK, BK, 1, true)) return e;
  if (int e = encode_map3(&tb, x, EB, K, M, 1, ldx * EB, (uint64_t)M * ldx * EB, BK, BK, 1, true)) return e;
  OperandWalk wa = zero_walk(), wb = zero_walk();
  wa.per_j[1] = BK; wa.per_box[0] = BK; wa.per_tile[0] = 128;
  wb.per_j[1] = BK; wb.per_box[0] = BK; wb.per_tile[0] = BN;
  const int num_kb = ceil_div(M, BK);
  const int tiles = ceil_div(N, 128) * ceil_div(K, BN);
  int splits = pick_splits(tiles, num_kb);
  {
    dim3 probe(ceil_div(N, 128), ceil_div(K, BN), splits);
    if (want_persistent(probe) || wide) splits = pick_splits_persistent(tiles, num_kb);
  }
  GemmShape shp{N, K, num_kb, num_kb, splits};
  shp.f16 = kIsF16<AT>;
  EpiAtomic::Params ep{dw, lddw, 0, alpha};
  dim3 grid(ceil_div(N, 128), ceil_div(K, BN), shp.splits);
  if (wide && gemm_pair(grid.x, BN)) return launch_gemm_persistent<256, true, true, EB, EpiAtomic, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  if (wide) return launch_gemm_persistent<256, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  if (want_persistent(grid)) {
    if (BN == 64) return launch_gemm_persistent<64, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
    return launch_gemm_persistent<128, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  }
  if (BN == 64) return launch_gemm<64, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  return launch_gemm<128, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
}

// ------------------------------------------------------------------------------------ Conv1d k=5 pad=2
// activations channels-last [R, T, C]; weights re-laid as w_k[Cout][5][Cin] (prep_conv_weight).
static bool conv_tile_geometry(int T, int* box_t, int* box_r, int* tiles_per_seq) {
  if (T <= 128) {
    if (128 % T) return false;
    *box_t = T; *box_r = 128 / T; *tiles_per_seq = 0;  // 0: several sequences per tile
  } else {
    if (T % 128) return false;
    *box_t = 128; *box_r = 1; *tiles_per_seq = T / 128;
  }
  return true;
}

// YT: storage type of the output (AT, or float: the fp16 mode keeps the pre-BatchNorm convolution output unrounded)
template <typename AT, typename YT = AT>
static int conv5_fwd_t(const AT* x, const AT* wk, const float* bias, YT* y, float* y_f32, int R, int T, int Cin, int Cout,
                       bool dgrad, cudaStream_t st, double* bn_sums = nullptr, int rows_half = 0, bool* stats_fused = nullptr) {
  // bn_sums != nullptr: also accumulate the train-mode BatchNorm statistics of y (per column sum / sum of squares of the
  // stored values, per statistics half) -- fused into the staged store epilogue when that path is taken (*stats_fused).
  // fwd :  y[r,t,co]  = sum_{k,ci} x[r,t+k-2,ci] wk[co][k][ci]                     (B K-major)
  // dgrad: dx[r,t,ci] = sum_{k',co} dy[r,t+k'-2,co] wk[co][4-k'][ci]               (B MN-major)
  // In dgrad mode the caller passes x:=dy, y:=dx, and (Cin, Cout) are still the forward layer's.
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  constexpr int YB = sizeof(YT);
  constexpr int YK = 128 / YB;   // output elements per 128-byte TMA box row
  int box_t, box_r, tps;
  DVAE_REQUIRE(conv_tile_geometry(T, &box_t, &box_r, &tps), "T must divide 128 or be a multiple of 128");
  DVAE_REQUIRE(tps == 0, "T > 128 not wired for the conv path (model is locked to T = 64)");
  const int Ca = dgrad ? Cout : Cin;  // channels of the A operand (reduction)
  const int Cn = dgrad ? Cin : Cout;  // output channels of this GEMM
  CUtensorMap ta, tb;
  if (int e = encode_map3(&ta, x, EB, Ca, T, R, (uint64_t)Ca * EB, (uint64_t)T * Ca * EB, BK, box_t, box_r)) return e;
  OperandWalk wa = zero_walk(), wb = zero_walk();
  wa.base[1] = -2; wa.per_j[0] = BK; wa.per_tap[1] = 1; wa.per_tile[2] = box_r;
  const int BN = pick_bn(Cn);
  if (!dgrad) {
    if (int e = encode_map3(&tb, wk, EB, Cin, 5, Cout, (uint64_t)Cin * EB, (uint64_t)5 * Cin * EB, BK, 1, BN)) return e;
    wb.per_j[0] = BK; wb.per_tap[1] = 1; wb.per_tile[2] = BN;
  } else {
    if (int e = encode_map3(&tb, wk, EB, Cin, 5, Cout, (uint64_t)Cin * EB, (uint64_t)5 * Cin * EB, BK, 1, BK, true)) return e;
    wb.base[1] = 4; wb.per_tap[1] = -1; wb.per_j[2] = BK; wb.per_box[0] = BK; wb.per_tile[0] = BN;
  }
  const int kpt = ceil_div(Ca, BK);
  GemmShape shp{R * T, Cn, 5 * kpt, kpt, 1};
  shp.f16 = kIsF16<AT>;
  typename EpiStore<YT>::Params ep{y, y_f32, bias, nullptr, (long)Cn, 0, 0};
  static const int fuse_env = env_int("DVAE_BN_FUSE_STATS", 1);
  const bool want_stats = fuse_env != 0 && bn_sums != nullptr && !dgrad && rows_half > 0 && rows_half % 128 == 0;
  if (stats_fused) *stats_fused = false;
  const int mt = pick_mt((long)R * T, ceil_div(Cn, BN));
  dim3 grid(ceil_div((long)R * T, 128 * mt), ceil_div(Cn, BN), 1);
  if (mt == 2) {
    if (!dgrad) {
      switch (BN) {
        case 64: return launch_gemm<64, false, false, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
        case 128: return launch_gemm<128, false, false, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
        default: return launch_gemm<256, false, false, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      }
    }
    switch (BN) {
      case 64: return launch_gemm<64, false, true, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm<128, false, true, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm<256, false, true, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  if (want_persistent(grid) && mt == 1 && gemm_pair(grid.x, BN)) {   // CTA pairs (cta_group::2)
    if (!dgrad) {   // K-major filter tile: re-encode with half the rows per box
      if (int e = encode_map3(&tb, wk, EB, Cin, 5, Cout, (uint64_t)Cin * EB, (uint64_t)5 * Cin * EB, BK, 1, BN / 2)) return e;
    }
    if (use_tma_store(y, y_f32, nullptr, Cn, YB, shp.num_kb) || (want_stats && tma_store_possible(y, y_f32, nullptr, Cn, YB))) {
      typename EpiStoreTma<YT>::Params et;
      if (int e = encode_map3(&et.tm_out, y, YB, Cn, (uint64_t)R * T, 1, (uint64_t)Cn * YB, (uint64_t)R * T * Cn * YB, YK, 128, 1))
        return e;
      et.bias = bias; et.relu = 0;
      et.stat_sums = want_stats ? bn_sums : nullptr; et.rows_half = rows_half;
      if (want_stats && stats_fused) *stats_fused = true;
      if (!dgrad) return launch_gemm_persistent<256, false, false, EB, EpiStoreTma<YT>, 2>(ta, tb, wa, wb, shp, et, grid, st);
      return launch_gemm_persistent<256, false, true, EB, EpiStoreTma<YT>, 2>(ta, tb, wa, wb, shp, et, grid, st);
    }
    if (!dgrad) return launch_gemm_persistent<256, false, false, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
    return launch_gemm_persistent<256, false, true, EB, EpiStore<YT>, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  }
  if (want_persistent(grid) &&
      (use_tma_store(y, y_f32, nullptr, Cn, YB, shp.num_kb) || (want_stats && tma_store_possible(y, y_f32, nullptr, Cn, YB)))) {
    typename EpiStoreTma<YT>::Params et;
    if (int e = encode_map3(&et.tm_out, y, YB, Cn, (uint64_t)R * T, 1, (uint64_t)Cn * YB, (uint64_t)R * T * Cn * YB, YK, 128, 1))
      return e;
    et.bias = bias; et.relu = 0;
    et.stat_sums = want_stats ? bn_sums : nullptr; et.rows_half = rows_half;
    if (want_stats && stats_fused) *stats_fused = true;
    if (!dgrad) {
      switch (BN) {
        case 64: return launch_gemm_persistent<64, false, false, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
        case 128: return launch_gemm_persistent<128, false, false, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
        default: return launch_gemm_persistent<256, false, false, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
      }
    }
    switch (BN) {
      case 64: return launch_gemm_persistent<64, false, true, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
      case 128: return launch_gemm_persistent<128, false, true, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
      default: return launch_gemm_persistent<256, false, true, EB, EpiStoreTma<YT>>(ta, tb, wa, wb, shp, et, grid, st);
    }
  }
  if (want_persistent(grid)) {
    if (!dgrad) {
      switch (BN) {
        case 64: return launch_gemm_persistent<64, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
        case 128: return launch_gemm_persistent<128, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
        default: return launch_gemm_persistent<256, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
      }
    }
    switch (BN) {
      case 64: return launch_gemm_persistent<64, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm_persistent<128, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm_persistent<256, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  if (!dgrad) {
    switch (BN) {
      case 64: return launch_gemm<64, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
      case 128: return launch_gemm<128, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
      default: return launch_gemm<256, false, false, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
    }
  }
  switch (BN) {
    case 64: return launch_gemm<64, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
    case 128: return launch_gemm<128, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
    default: return launch_gemm<256, false, true, EB, EpiStore<YT>>(ta, tb, wa, wb, shp, ep, grid, st);
  }
}

template <typename AT>
static int conv5_wgrad_t(const AT* dy, const AT* x, float* dwk, int R, int T, int Cin, int Cout, float alpha, cudaStream_t st) {
  // dwk[co][k][ci] += sum_{r,t} dy[r,t,co] x[r,t+k-2,ci];  one GEMM batch (blockIdx.z) per tap, split-K over sequences
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  DVAE_REQUIRE(T % BK == 0, "T must be a multiple of the k-block");
  CUtensorMap ta, tb;
  if (int e = encode_map3(&ta, dy, EB, Cout, T, R, (uint64_t)Cout * EB, (uint64_t)T * Cout * EB, BK, BK, 1, true)) return e;
  if (int e = encode_map3(&tb, x, EB, Cin, T, R, (uint64_t)Cin * EB, (uint64_t)T * Cin * EB, BK, BK, 1, true)) return e;
  // 128 x 256 tiles when the filter is wide enough and the launch is persistent: the 128 x 128 tile reads 8 KB of operands
  // per 64 tensor-core clocks (exactly the shared-memory bandwidth) and moves 1.33x the L2->SM bytes per FLOP.
  static const int wide_env = env_int("DVAE_WGRAD_WIDE", 1);
  int BN = Cin <= 64 ? 64 : 128;
  const bool wide = wide_env != 0 && !g_background && Cin % 256 == 0;
  if (wide) BN = 256;
  OperandWalk wa = zero_walk(), wb = zero_walk();
  // k-block kb -> (sequence r = kb / (T/BK)  [the "tap" slot of the walk], time block j = kb % (T/BK))
  wa.per_j[1] = BK; wa.per_tap[2] = 1; wa.per_box[0] = BK; wa.per_tile[0] = 128;
  wb.base[1] = -2; wb.per_z[1] = 1; wb.per_j[1] = BK; wb.per_tap[2] = 1; wb.per_box[0] = BK; wb.per_tile[0] = BN;
  const int kpt = T / BK;
  const int num_kb = R * kpt;
  const int tiles = ceil_div(Cout, 128) * ceil_div(Cin, BN) * 5;
  int splits = pick_splits(tiles, num_kb);
  {
    dim3 probe(ceil_div(Cout, 128), ceil_div(Cin, BN), 5 * splits);
    if (want_persistent(probe) || wide) splits = pick_splits_persistent(tiles, num_kb);
  }
  GemmShape shp{Cout, Cin, num_kb, kpt, splits};
  shp.f16 = kIsF16<AT>;
  static const int tap_fast = env_int("DVAE_WGRAD_TAP_FAST", 1);
  shp.batches = tap_fast ? 5 : 0;   // persistent kernel: the 5 taps of a K-slab run side by side (L2 reuse of dy and x)
  EpiAtomic::Params ep{dwk, (long)5 * Cin, (long)Cin, alpha};
  dim3 grid(ceil_div(Cout, 128), ceil_div(Cin, BN), 5 * shp.splits);
  if (wide && gemm_pair(grid.x, BN)) return launch_gemm_persistent<256, true, true, EB, EpiAtomic, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  if (wide) return launch_gemm_persistent<256, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  if (want_persistent(grid)) {
    if (BN == 64) return launch_gemm_persistent<64, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
    return launch_gemm_persistent<128, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  }
  if (BN == 64) return launch_gemm<64, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  return launch_gemm<128, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
}

// ------------------------------------------------------------------------------------ LSTM
// Layouts (rows = sequences, D = directions):
//   xg    [rows, T, D*4H]  in: x-projection + biases, gate-interleaved per direction (column 4*u + g = gate g of
//                          unit u); out: activated gates (same layout, overwritten in place)
//   h_all [rows, T, D*H]   layer output, also the A operand of the next step (3-D TMA over {D*H, T, rows})
//   c_all [rows, T, D*H]   fp32 cell state
//   whh_p [D][4H][H]       recurrent weights, rows gate-interleaved like xg columns
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
// Gate tile of the forward recurrence (columns of one N tile = [i|f|g|o] x tile/4 hidden units).  Tunable for sweeps via
// DVAE_LSTM_FWD_TILE / DVAE_LSTM_FWD_TILE_SMALL (read once; the weight permutation follows dvae_lstm_gate_tile()).
static int lstm_fwd_bn(int H) {
  static const int big = env_int("DVAE_LSTM_FWD_TILE", 0);   // 0: by hidden size (profiles/r01_lstm_tile_sweep_v2.txt)
  static const int small = env_int("DVAE_LSTM_FWD_TILE_SMALL", 64);
  int t = (H == 64) ? small : (big ? big : (H >= 1024 ? 256 : 128));
  if (t != 64 && t != 128 && t != 256) t = 128;
  while (t > 4 * H) t >>= 1;
  return t;
}
static int lstm_bwd_bn(int H) {
  static const int v = env_int("DVAE_LSTM_BWD_TILE", 64);
  int t = (v == 64 || v == 128 || v == 256) ? v : 64;
  while (t > H) t >>= 1;
  return t;
}
// split-K of the backward step GEMM (K = 4H is long, the grid is small): fill ~2 CTAs per SM, >= 4 k-blocks per split
static int lstm_bwd_splits(int rows, int H, int D, int elem_bytes) {
  static const int forced = env_int("DVAE_LSTM_BWD_SPLITS", 0);
  const int tiles = ceil_div(rows, 128) * (H / lstm_bwd_bn(H)) * D;
  const int num_kb = 4 * H / (128 / elem_bytes);
  (void)tiles;
  int s = forced > 0 ? forced : 1;   // measured (profiles/r01_lstm_tile_sweep_v2.txt): splitting does not pay for these shapes
  if (s > num_kb / 4) s = num_kb / 4;
  return s < 1 ? 1 : s;
}

// CTA-pair MMA (cta_group::2) for the recurrence step kernels: DVAE_LSTM_PAIR=1, when the row tiles pair up.  Default
// off: measured on B200 (profiles/r01_lstm_cluster_multicast.txt) the pair shortens the backward step's main loop
// (8.96 -> 7.68 us at H = 1024) but lengthens the forward's (8.45 -> 10.75 us), and a cluster launch per time step adds
// ~1.5 us of set-up (cluster scheduling + two cluster barriers), so a step is 0.1 - 2.5 us slower end to end.
// (An earlier experiment shared the weight tile by plain TMA multicast across 1-SM MMAs instead; also slower.)
static bool lstm_pair(int row_tiles) {
  static const int want = env_int("DVAE_LSTM_PAIR", 0);
  return want != 0 && row_tiles % 2 == 0;
}

template <typename AT>
static int lstm_fwd_t(AT* xg, const AT* whh_p, AT* h_all, float* c_all, int rows, int T, int H, int D, cudaStream_t st) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  DVAE_REQUIRE(H % BK == 0 && (D == 1 || D == 2), "H must be a multiple of the k-block; D in {1,2}");
  const int BN = lstm_fwd_bn(H);
  DVAE_REQUIRE((4 * H) % BN == 0, "4H must be a multiple of the gate tile");
  CUtensorMap ta, tb;
  if (int e = encode_map3(&ta, h_all, EB, (uint64_t)D * H, T, rows, (uint64_t)D * H * EB, (uint64_t)T * D * H * EB, BK, 1, 128))
    return e;
  // CTA pairs (cta_group::2): each CTA of a pair of consecutive row tiles stages half of the W_hh tile
  const bool pair = lstm_pair(ceil_div(rows, 128)) && BN >= 128;
  if (int e = encode_map3(&tb, whh_p, EB, H, 4 * H, D, (uint64_t)H * EB, (uint64_t)4 * H * H * EB, BK, pair ? BN / 2 : BN, 1)) return e;
  const long ldx = (long)T * D * 4 * H, ldc = (long)T * D * H;
  // staged (TMA) stores of gates / c / h for the wide tiles; DVAE_LSTM_FWD_TMA=0 keeps the direct row-per-lane stores
  static const int tma_env = env_int("DVAE_LSTM_FWD_TMA", 1);
  const bool staged = (tma_env != 0 && BN >= 128) || pair;
  CUtensorMap tg, tc, th;
  if (staged) {
    if (int e = encode_map3(&tg, xg, EB, (uint64_t)D * 4 * H, T, rows, (uint64_t)D * 4 * H * EB, (uint64_t)T * D * 4 * H * EB, BK, 1, 128))
      return e;
    if (int e = encode_map3(&tc, c_all, 4, (uint64_t)D * H, T, rows, (uint64_t)D * H * 4, (uint64_t)T * D * H * 4, 32, 1, 128)) return e;
    const int h_row_bytes = BN / 4 * EB;
    if (int e = h_row_bytes >= 128
                    ? encode_map3(&th, h_all, EB, (uint64_t)D * H, T, rows, (uint64_t)D * H * EB, (uint64_t)T * D * H * EB, BK, 1, 128)
                    : encode_map3_narrow(&th, h_all, EB, (uint64_t)D * H, T, rows, (uint64_t)D * H * EB, (uint64_t)T * D * H * EB,
                                         BN / 4, 1, 128))
      return e;
  }
  for (int s = 0; s < T; ++s) {
    const int tf = s, tr = T - 1 - s;
    OperandWalk wa = zero_walk(), wb = zero_walk();
    wa.base[1] = tf - 1; wa.per_j[0] = BK; wa.per_tile[2] = 128; wa.per_z[0] = H; wa.per_z[1] = (tr + 1) - (tf - 1);
    wb.per_j[0] = BK; wb.per_tile[1] = BN; wb.per_z[2] = 1;
    GemmShape shp{rows, 4 * H, s == 0 ? 0 : H / BK, H / BK, 1};
    shp.f16 = kIsF16<AT>;
    auto fill = [&](typename EpiLstmFwd<AT>::Params& ep) {
      ep.xproj = xg + (long)tf * D * 4 * H;
      ep.gates = xg + (long)tf * D * 4 * H;
      ep.c_prev = s == 0 ? nullptr : c_all + (long)(tf - 1) * D * H;
      ep.c_out = c_all + (long)tf * D * H;
      ep.h_out = h_all + (long)tf * D * H;
      ep.ldx = ldx; ep.ldc = ldc; ep.ldh = ldc;
      ep.z_x = 4 * H + (long)(tr - tf) * D * 4 * H;
      ep.z_c_prev = H + (long)((tr + 1) - (tf - 1)) * D * H;
      ep.z_c_out = H + (long)(tr - tf) * D * H;
      ep.z_h = H + (long)(tr - tf) * D * H;
    };
    dim3 grid(ceil_div(rows, 128), 4 * H / BN, D);
    int e;
    if (staged) {
      typename EpiLstmFwdTma<AT>::Params ep;
      fill(ep);
      ep.tm_g = tg; ep.tm_c = tc; ep.tm_h = th;
      ep.t[0] = tf; ep.t[1] = tr; ep.H = H;
      if (pair)
        e = (BN == 256) ? launch_gemm_pair<256, false, false, EB, EpiLstmFwdTma<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                        : launch_gemm_pair<128, false, false, EB, EpiLstmFwdTma<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
      else
        e = (BN == 256) ? launch_gemm<256, false, false, EB, EpiLstmFwdTma<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                        : launch_gemm<128, false, false, EB, EpiLstmFwdTma<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    } else {
      typename EpiLstmFwd<AT>::Params ep;
      fill(ep);
      e = (BN == 256)   ? launch_gemm<256, false, false, EB, EpiLstmFwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
          : (BN == 128) ? launch_gemm<128, false, false, EB, EpiLstmFwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                        : launch_gemm<64, false, false, EB, EpiLstmFwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    }
    if (e) return e;
  }
  return 0;
}

// LSTM cell backward for one time step, elementwise over [rows, H] (both directions: blockIdx.y).  dh = dh_out + dh_rec;
// see EpiLstmBwd for the math.  One thread = 8 hidden units of one row: every access is a 16/32/64-byte run and
// consecutive threads touch consecutive runs, so the kernel streams at HBM/L2 bandwidth with full occupancy -- unlike the
// same arithmetic inside the GEMM epilogue, which is latency-bound on 8 warps (profiles/r01_phase_timing_v2.txt).
template <typename AT>
__global__ void lstm_cell_bwd_kernel(const AT* __restrict__ dh_out, float* __restrict__ dh_rec,
                                     const AT* __restrict__ gates, const float* __restrict__ c_t,
                                     const float* __restrict__ c_prev, float* __restrict__ dc, AT* __restrict__ da,
                                     int rows, int H, long ldh, long ldx, long z_h, long z_x, long z_cprev, long z_rec,
                                     int dc_zero, int zero_rec) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int d = blockIdx.y;
  const int upr = H >> 3;  // 8-unit chunks per row
  const long total = static_cast<long>(rows) * upr;
  dh_out += d * z_h; gates += d * z_x; c_t += d * z_h; da += d * z_x; dc += d * z_rec;
  if (c_prev) c_prev += d * z_cprev;
  if (dh_rec) dh_rec += d * z_rec;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / upr;
    const int u = static_cast<int>(i - r * upr) * 8;
    float dh[8], g4[32], cc[8], cpv[8], dcv[8], rec[8];
    Act8<AT>::load(dh_out + r * ldh + u, dh);
#pragma unroll
    for (int j = 0; j < 4; ++j) Act8<AT>::load(gates + r * ldx + 4 * u + 8 * j, g4 + 8 * j);
    Act8<float>::load(c_t + r * ldh + u, cc);
    if (c_prev) Act8<float>::load(c_prev + r * ldh + u, cpv);
    if (dh_rec) {
      Act8<float>::load(dh_rec + r * H + u, rec);
      if (zero_rec) {   // the next step's split-K GEMM accumulates into this buffer with TMA reduce-adds
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        Act8<float>::store(dh_rec + r * H + u, z);
      }
    }
    if (!dc_zero) Act8<float>::load(dc + r * H + u, dcv);
    float dai[8], daf[8], dag[8], dao[8], dcn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float ig = g4[4 * k], fg = g4[4 * k + 1], gg = g4[4 * k + 2], og = g4[4 * k + 3];
      const float tc = GateMath<AT>::tnh(cc[k]);
      const float dht = dh[k] + (dh_rec ? rec[k] : 0.f);
      const float dct = (dc_zero ? 0.f : dcv[k]) + dht * og * (1.f - tc * tc);
      dao[k] = dht * tc * og * (1.f - og);
      dai[k] = dct * gg * ig * (1.f - ig);
      dag[k] = dct * ig * (1.f - gg * gg);
      daf[k] = dct * (c_prev ? cpv[k] : 0.f) * fg * (1.f - fg);
      dcn[k] = dct * fg;
    }
    Act8<float>::store(dc + r * H + u, dcn);
    AT* o = da + r * ldx + u;
    Act8<AT>::store(o, dai);
    Act8<AT>::store(o + H, daf);
    Act8<AT>::store(o + 2 * H, dag);
    Act8<AT>::store(o + 3 * H, dao);
  }
}

// Backward through time.  dh_all [rows,T,D*H] is the gradient wrt the layer output; produces da_all
// [rows,T,D*4H] (pre-activation gate gradients, natural torch order i,f,g,o per direction).  whh_n is the
// natural-order copy [D][4H][H].  dc_ws: fp32 [2][D, rows, H] scratch (dc carry, then the dh_rec buffer of the split path).
template <typename AT>
static int lstm_bwd_t(const AT* dh_all, const AT* gates, const float* c_all, const AT* whh_n, AT* da_all, float* dc_ws,
                      float* splitk_ws, int* tickets, int rows, int T, int H, int D, cudaStream_t st) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  const int BN = lstm_bwd_bn(H);
  DVAE_REQUIRE(H % 64 == 0 && (D == 1 || D == 2), "H must be a multiple of 64; D in {1,2}");
  CUtensorMap ta, tb;
  if (int e = encode_map3(&ta, da_all, EB, (uint64_t)D * 4 * H, T, rows, (uint64_t)D * 4 * H * EB,
                          (uint64_t)T * D * 4 * H * EB, BK, 1, 128))
    return e;
  if (int e = encode_map3(&tb, whh_n, EB, H, 4 * H, D, (uint64_t)H * EB, (uint64_t)4 * H * H * EB, BK, BK, 1, true)) return e;
  const long ldh = (long)T * D * H, ldx = (long)T * D * 4 * H;
  static const int fused = env_int("DVAE_LSTM_BWD_FUSED", 0);
  if (!fused) {
    // Two kernels per step: (1) dh_rec[d][rows][H] (fp32) = da_{t+1} . W_hh, (2) the elementwise cell backward.
    // dc_ws holds [D][rows][H] carry followed by [D][rows][H] dh_rec.
    //
    // (1) is a K-heavy GEMM (K = 4H, N = H).  With one CTA per 128 x 64 output tile every CTA streams the whole K extent
    // of both operands and the step is bound by L2->SM bandwidth (192 MB per step at H = 1024).  "reduce" mode instead uses
    // wide tiles (128 x 256 / 128 x 128) and splits K over gridDim.z: the same SM count moves half the operand bytes, and
    // the partial sums meet in L2 through TMA reduce-adds of the staged tiles.  The cell kernel consumes dh_rec and
    // leaves it zeroed for the next step.
    float* dh_rec = dc_ws + (long)D * rows * H;
    static const int reduce_env = env_int("DVAE_LSTM_BWD_REDUCE", 1);
    const bool reduce = reduce_env != 0 && H >= 256 && H % 128 == 0;
    const int RBN = (H % 256 == 0) ? 256 : 128;
    int rsplits = 1;
    CUtensorMap trec;
    if (reduce) {
      static const int forced = env_int("DVAE_LSTM_BWD_SPLITS", 0);
      const int tiles = ceil_div(rows, 128) * (H / RBN) * D;
      const int num_kb = 4 * H / BK;
      rsplits = forced > 0 ? forced : 1;
      if (forced <= 0)
        while (tiles * rsplits * 2 <= num_sms() && num_kb / (rsplits * 2) >= 4) rsplits *= 2;
      if (int e = encode_map3(&trec, dh_rec, 4, H, rows, D, (uint64_t)H * 4, (uint64_t)rows * H * 4, 32, 128, 1)) return e;
      DVAE_CHECK_CUDA(cudaMemsetAsync(dh_rec, 0, sizeof(float) * D * rows * H, st));
    }
    for (int s = 0; s < T; ++s) {
      const int tf = T - 1 - s, tr = s;
      if (s > 0) {
        OperandWalk wa = zero_walk(), wb = zero_walk();
        wa.base[1] = tf + 1; wa.per_j[0] = BK; wa.per_tile[2] = 128; wa.per_z[0] = 4 * H; wa.per_z[1] = (tr - 1) - (tf + 1);
        if (reduce) {
          wb.per_j[1] = BK; wb.per_box[0] = BK; wb.per_tile[0] = RBN; wb.per_z[2] = 1;
          GemmShape shp{rows, H, 4 * H / BK, 4 * H / BK, rsplits, nullptr, nullptr};
          shp.f16 = kIsF16<AT>;
          EpiReduceTma::Params ep{};
          ep.tm_out = trec;
          static const int pf_env = env_int("DVAE_LSTM_BWD_PREFETCH", 1);
          if (pf_env && D == 1) {   // what this step's cell-backward kernel (next in the stream) will read: time index tf
            const AT* g_t = gates + (long)tf * D * 4 * H;
            const float* c_t = c_all + (long)tf * D * H;
            const AT* dh_t = dh_all + (long)tf * D * H;
            ep.pf_ptr[0] = g_t;  ep.pf_row_stride[0] = ldx * EB; ep.pf_row_bytes[0] = 4 * H * EB;
            ep.pf_ptr[1] = c_t;  ep.pf_row_stride[1] = ldh * 4;  ep.pf_row_bytes[1] = H * 4;
            ep.pf_ptr[2] = (s == T - 1) ? nullptr : c_t - (long)D * H;
            ep.pf_row_stride[2] = ldh * 4; ep.pf_row_bytes[2] = H * 4;
            ep.pf_ptr[3] = dh_t; ep.pf_row_stride[3] = ldh * EB; ep.pf_row_bytes[3] = H * EB;
            ep.pf_rows = rows;
          }
          dim3 grid(ceil_div(rows, 128), H / RBN, D * rsplits);
          int e;
          if (lstm_pair(ceil_div(rows, 128)))
            e = (RBN == 256) ? launch_gemm_pair<256, false, true, EB, EpiReduceTma>(ta, tb, wa, wb, shp, ep, grid, st)
                             : launch_gemm_pair<128, false, true, EB, EpiReduceTma>(ta, tb, wa, wb, shp, ep, grid, st);
          else
            e = (RBN == 256) ? launch_gemm<256, false, true, EB, EpiReduceTma>(ta, tb, wa, wb, shp, ep, grid, st)
                             : launch_gemm<128, false, true, EB, EpiReduceTma>(ta, tb, wa, wb, shp, ep, grid, st);
          if (e) return e;
        } else {
          wb.per_j[1] = BK; wb.per_box[0] = BK; wb.per_tile[0] = BN; wb.per_z[2] = 1;
          const int splits = splitk_ws == nullptr ? 1 : lstm_bwd_splits(rows, H, D, EB);
          GemmShape shp{rows, H, 4 * H / BK, 4 * H / BK, splits, splitk_ws, tickets};
          shp.f16 = kIsF16<AT>;
          typename EpiStore<AT>::Params ep{nullptr, dh_rec, nullptr, nullptr, (long)H, (long)rows * H, 0};
          dim3 grid(ceil_div(rows, 128), H / BN, D * splits);
          int e = (BN == 256)   ? launch_gemm<256, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                  : (BN == 128) ? launch_gemm<128, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                                : launch_gemm<64, false, true, EB, EpiStore<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
          if (e) return e;
        }
      }
      const long total = (long)rows * (H / 8);
      long gx = (total + 255) / 256;
      if (gx > 148 * 8) gx = 148 * 8;
      dim3 cgrid((unsigned)gx, D);
      {
        static const int use_pdl = env_int("DVAE_PDL", 1);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = cgrid;
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = use_pdl ? 1 : 0;
        DVAE_CHECK_CUDA(cudaLaunchKernelEx(
            &cfg, lstm_cell_bwd_kernel<AT>, (const AT*)(dh_all + (long)tf * D * H), (float*)(s > 0 ? dh_rec : nullptr),
            (const AT*)(gates + (long)tf * D * 4 * H), (const float*)(c_all + (long)tf * D * H),
            (const float*)((s == T - 1) ? nullptr : c_all + (long)(tf - 1) * D * H), dc_ws, da_all + (long)tf * D * 4 * H, rows, H,
            ldh, ldx, H + (long)(tr - tf) * D * H, 4 * H + (long)(tr - tf) * D * 4 * H,
            H + (long)((tr + 1) - (tf - 1)) * D * H, (long)rows * H, (int)(s == 0), (int)reduce));
      }
    }
    return 0;
  }
  for (int s = 0; s < T; ++s) {
    const int tf = T - 1 - s, tr = s;  // forward direction walks time backwards, reverse direction forwards
    OperandWalk wa = zero_walk(), wb = zero_walk();
    wa.base[1] = tf + 1; wa.per_j[0] = BK; wa.per_tile[2] = 128; wa.per_z[0] = 4 * H; wa.per_z[1] = (tr - 1) - (tf + 1);
    wb.per_j[1] = BK; wb.per_box[0] = BK; wb.per_tile[0] = BN; wb.per_z[2] = 1;
    const int splits = (s == 0 || splitk_ws == nullptr) ? 1 : lstm_bwd_splits(rows, H, D, EB);
    GemmShape shp{rows, H, s == 0 ? 0 : 4 * H / BK, 4 * H / BK, splits, splitk_ws, tickets};
    shp.f16 = kIsF16<AT>;
    typename EpiLstmBwd<AT>::Params ep;
    ep.dh_out = dh_all + (long)tf * D * H;
    ep.gates = gates + (long)tf * D * 4 * H;
    ep.c_t = c_all + (long)tf * D * H;
    ep.c_prev = (s == T - 1) ? nullptr : c_all + (long)(tf - 1) * D * H;
    ep.dc = dc_ws;
    ep.da = da_all + (long)tf * D * 4 * H;
    ep.ldh = ldh; ep.ldx = ldx; ep.ldc = ldh; ep.lda = ldx;
    ep.z_h = H + (long)(tr - tf) * D * H;
    ep.z_x = 4 * H + (long)(tr - tf) * D * 4 * H;
    ep.z_c = H + (long)(tr - tf) * D * H;
    ep.z_c_prev = H + (long)((tr + 1) - (tf - 1)) * D * H;
    ep.z_dc = (long)rows * H;
    ep.z_a = 4 * H + (long)(tr - tf) * D * 4 * H;
    ep.H = H; ep.dc_zero = (s == 0);
    dim3 grid(ceil_div(rows, 128), H / BN, D * splits);
    int e = (BN == 256)   ? launch_gemm<256, false, true, EB, EpiLstmBwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
            : (BN == 128) ? launch_gemm<128, false, true, EB, EpiLstmBwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st)
                          : launch_gemm<64, false, true, EB, EpiLstmBwd<AT>>(ta, tb, wa, wb, shp, ep, grid, st);
    if (e) return e;
  }
  return 0;
}

// dW_hh[d][4H][H] += sum_{r,t} da[r,t,d,:]^T h_prev[r,t,d,:]  with h_prev = h[t-1] (forward dir) / h[t+1] (reverse)
template <typename AT>
static int lstm_wgrad_hh_t(const AT* da_all, const AT* h_all, float* dwhh, int rows, int T, int H, int D, float alpha,
                           cudaStream_t st) {
  constexpr int EB = sizeof(AT);
  constexpr int BK = 128 / EB;
  DVAE_REQUIRE(T % BK == 0, "T must be a multiple of the k-block");
  CUtensorMap ta, tb;
  if (int e = encode_map3(&ta, da_all, EB, (uint64_t)D * 4 * H, T, rows, (uint64_t)D * 4 * H * EB,
                          (uint64_t)T * D * 4 * H * EB, BK, BK, 1, true))
    return e;
  if (int e = encode_map3(&tb, h_all, EB, (uint64_t)D * H, T, rows, (uint64_t)D * H * EB, (uint64_t)T * D * H * EB, BK, BK, 1, true))
    return e;
  static const int wide_env = env_int("DVAE_WGRAD_WIDE", 1);
  const bool wide = wide_env != 0 && !g_background && H % 256 == 0;
  const int BN = wide ? 256 : (H <= 64 ? 64 : 128);
  OperandWalk wa = zero_walk(), wb = zero_walk();
  wa.per_j[1] = BK; wa.per_tap[2] = 1; wa.per_box[0] = BK; wa.per_tile[0] = 128; wa.per_z[0] = 4 * H;
  wb.base[1] = -1; wb.per_z[1] = 2; wb.per_z[0] = H; wb.per_j[1] = BK; wb.per_tap[2] = 1; wb.per_box[0] = BK;
  wb.per_tile[0] = BN;
  const int kpt = T / BK, num_kb = rows * kpt;
  const int tiles = ceil_div(4 * H, 128) * ceil_div(H, BN) * D;
  GemmShape shp{4 * H, H, num_kb, kpt, wide ? pick_splits_persistent(tiles, num_kb) : pick_splits(tiles, num_kb)};
  shp.f16 = kIsF16<AT>;
  EpiAtomic::Params ep{dwhh, (long)H, (long)4 * H * H, alpha};
  dim3 grid(ceil_div(4 * H, 128), ceil_div(H, BN), D * shp.splits);
  if (wide && gemm_pair(grid.x, BN)) return launch_gemm_persistent<256, true, true, EB, EpiAtomic, 2>(ta, tb, wa, wb, shp, ep, grid, st);
  if (wide) return launch_gemm_persistent<256, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  if (BN == 64) return launch_gemm<64, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
  return launch_gemm<128, true, true, EB, EpiAtomic>(ta, tb, wa, wb, shp, ep, grid, st);
}

}  // namespace dvae

namespace dvae {
template <typename AT>
static int conv5_fwd_bnstats_t(int dtype, const void* x, const void* wk, const float* bias, void* y, int y_f32, int R, int T, int Cin,
                               int Cout, double* bn_ws, int rows_half, int halves, cudaStream_t st) {
  bool fused = false;
  int e;
  if (y_f32) {
    if constexpr (std::is_same<AT, __half>::value || std::is_same<AT, tf32_t>::value) {   // fp16: fp32 storage; tf32: no rounding
      e = conv5_fwd_t<AT, float>((const AT*)x, (const AT*)wk, bias, (float*)y, nullptr, R, T, Cin, Cout, false, st, bn_ws, rows_half, &fused);
    } else {
      set_last_error("unrounded fp32 convolution output is implemented for the fp16 and tf32 activation dtypes");
      return 1;
    }
  } else e = conv5_fwd_t<AT>((const AT*)x, (const AT*)wk, bias, (AT*)y, nullptr, R, T, Cin, Cout, false, st, bn_ws, rows_half, &fused);
  if (e) return e;
  if (!fused) return bn_stats_launch(y_f32 ? kF32 : dtype, y, bn_ws, rows_half, halves, Cout, st);
  return 0;
}
}  // namespace dvae

using namespace dvae;
using bf16 = __nv_bfloat16;
using f16 = __half;

// `AT` is the activation storage type selected by the dtype tag: bf16 / fp16 (tcgen05 kind::f16) or fp32 on the tf32 grid
#define DISPATCH_AT(dtype, ...)                                         \
  do {                                                                  \
    if ((dtype) == kBF16) { using AT = bf16; return __VA_ARGS__; }      \
    if ((dtype) == kF16) { using AT = f16; return __VA_ARGS__; }        \
    if ((dtype) == kTF32) { using AT = tf32_t; return __VA_ARGS__; }    \
    set_last_error("unknown dtype tag");                                \
    return 1;                                                           \
  } while (0)

extern "C" {

int dvae_linear_fwd(int dtype, const void* x, long ldx, const void* w, const float* bias, void* out, float* out_f32,
                    long ldo, int M, int N, int K, int relu, int block_n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_linear_fwd((const float*)x, ldx, (const float*)w, bias, (float*)out, out_f32, ldo, M, N, K, relu, st);
  DISPATCH_AT(dtype, linear_fwd_t<AT>((const AT*)x, ldx, (const AT*)w, bias, (AT*)out, out_f32, ldo, M, N, K, relu, block_n, st));
}
// Split-precision variants for the small linear layers (M = rows, not rows x frames: their cost is negligible):
// b_terms = 2: w is [N, 2K] = [w_hi | w_lo] (dvae_prep_cast_split) and out = x . (w_hi + w_lo)^T;
// out_cat (may be null): additionally act [M, 3N] = [hi | lo | hi] of the result, the K-concatenated operand of a following
// GEMM whose weight is laid out [w_hi | w_hi | w_lo] (dvae_prep_cast_split with parts = 3).
int dvae_linear_fwd_split(int dtype, const void* x, long ldx, const void* w, const float* bias, void* out, float* out_f32,
                          long ldo, void* out_cat, int M, int N, int K, int relu, int b_terms, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DISPATCH_AT(dtype, linear_fwd_t<AT>((const AT*)x, ldx, (const AT*)w, bias, (AT*)out, out_f32, ldo, M, N, K, relu, 0, st, b_terms, (AT*)out_cat));
}

int dvae_linear_dgrad(int dtype, const void* dy, long lddy, const void* w, void* dx, float* dx_f32, const void* relu_mask,
                      long ldx, int M, int N, int K, int block_n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_linear_dgrad((const float*)dy, lddy, (const float*)w, (float*)dx, dx_f32, (const float*)relu_mask, ldx, M, N, K, st);
  DISPATCH_AT(dtype, linear_dgrad_t<AT>((const AT*)dy, lddy, (const AT*)w, (AT*)dx, dx_f32, (const AT*)relu_mask, ldx, M, N, K, block_n, st));
}

// alpha scales what is added to dw (1 / gradient scale of the fp16 mode; 1 otherwise)
int dvae_linear_wgrad(int dtype, const void* dy, long lddy, const void* x, long ldx, float* dw, long lddw, int M, int N,
                      int K, float alpha, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_linear_wgrad((const float*)dy, lddy, (const float*)x, ldx, dw, lddw, M, N, K, alpha, st);
  DISPATCH_AT(dtype, linear_wgrad_t<AT>((const AT*)dy, lddy, (const AT*)x, ldx, dw, lddw, M, N, K, alpha, st));
}

int dvae_conv5_fwd(int dtype, const void* x, const void* wk, const float* bias, void* y, float* y_f32, int R, int T, int Cin,
                   int Cout, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_conv5((const float*)x, (const float*)wk, bias, (float*)y, y_f32, R, T, Cin, Cout, false, st);
  DISPATCH_AT(dtype, conv5_fwd_t<AT>((const AT*)x, (const AT*)wk, bias, (AT*)y, y_f32, R, T, Cin, Cout, false, st));
}

// Convolution + the statistics pass of the train-mode BatchNorm that follows it (ConvNorm -> BatchNorm1d,
// model/disentangled_vae.py:154-160): bn_ws [halves*2*Cout + 1] doubles receives per-half column sums / sums of squares of
// y (zeroed here).  Fused into the GEMM's staged store epilogue when possible, otherwise a separate reduction kernel.
// y_f32 != 0: y is stored as unrounded fp32 whatever the activation dtype (and BatchNorm then reads it with y_f32 set too)
int dvae_conv5_fwd_bnstats(int dtype, const void* x, const void* wk, const float* bias, void* y, int y_f32, int R, int T, int Cin,
                           int Cout, double* bn_ws, int rows_half, int halves, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  DVAE_REQUIRE(bn_ws != nullptr && halves >= 1 && (long)rows_half * halves == (long)R * T, "rows_half * halves must equal R * T");
  DVAE_CHECK_CUDA(cudaMemsetAsync(bn_ws, 0, sizeof(double) * ((long)halves * 2 * Cout + 1), st));
  if (dtype == kF32) {
    if (int e = simt_conv5((const float*)x, (const float*)wk, bias, (float*)y, nullptr, R, T, Cin, Cout, false, st)) return e;
    return bn_stats_launch(kF32, y, bn_ws, rows_half, halves, Cout, st);
  }
  DISPATCH_AT(dtype, conv5_fwd_bnstats_t<AT>(dtype, x, wk, bias, y, y_f32, R, T, Cin, Cout, bn_ws, rows_half, halves, st));
}

int dvae_conv5_dgrad(int dtype, const void* dy, const void* wk, void* dx, float* dx_f32, int R, int T, int Cin, int Cout,
                     void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_conv5((const float*)dy, (const float*)wk, nullptr, (float*)dx, dx_f32, R, T, Cin, Cout, true, st);
  DISPATCH_AT(dtype, conv5_fwd_t<AT>((const AT*)dy, (const AT*)wk, nullptr, (AT*)dx, dx_f32, R, T, Cin, Cout, true, st));
}

int dvae_conv5_wgrad(int dtype, const void* dy, const void* x, float* dwk, int R, int T, int Cin, int Cout, float alpha,
                     void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_conv5_wgrad((const float*)dy, (const float*)x, dwk, R, T, Cin, Cout, alpha, st);
  DISPATCH_AT(dtype, conv5_wgrad_t<AT>((const AT*)dy, (const AT*)x, dwk, R, T, Cin, Cout, alpha, st));
}

int dvae_lstm_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, int D,
                  void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_lstm_fwd((float*)xg, (const float*)whh_p, (float*)h_all, c_all, rows, T, H, D, st);
  if (lstm_seq_supported(H, T)) return lstm_seq_fwd(dtype, xg, whh_p, h_all, c_all, rows, T, H, D, st);
  if (lstm_res_supported(dtype, rows, T, H, D)) {
    const int e = lstm_res_fwd(dtype, xg, whh_p, h_all, c_all, rows, T, H, st);
    if (e != 3) return e;   // 3: the grid cannot be co-resident on this device -> step-per-launch kernels
  }
  DISPATCH_AT(dtype, lstm_fwd_t<AT>((AT*)xg, (const AT*)whh_p, (AT*)h_all, c_all, rows, T, H, D, st));
}

// splitk_ws / tickets: fix-up workspace for the split-K backward step (sizes from dvae_lstm_bwd_workspace); may be
// null (then no split-K).  tickets must be zero before the first use (they reset themselves afterwards).
int dvae_lstm_bwd(int dtype, const void* dh_all, const void* gates, const float* c_all, const void* whh_n, void* da_all,
                  float* dc_ws, float* splitk_ws, int* tickets, int rows, int T, int H, int D, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_lstm_bwd((const float*)dh_all, (const float*)gates, c_all, (const float*)whh_n, (float*)da_all, dc_ws, rows, T, H, D, st);
  if (lstm_seq_supported(H, T)) return lstm_seq_bwd(dtype, dh_all, gates, c_all, whh_n, da_all, rows, T, H, D, st);
  DISPATCH_AT(dtype, lstm_bwd_t<AT>((const AT*)dh_all, (const AT*)gates, c_all, (const AT*)whh_n, (AT*)da_all, dc_ws, splitk_ws,
                                    tickets, rows, T, H, D, st));
}
// floats of fix-up workspace and number of tickets dvae_lstm_bwd wants for this shape (0 floats: no split-K)
int dvae_lstm_bwd_workspace(int dtype, int rows, int H, int D, long* ws_floats, int* num_tickets) {
  const int eb = dtype == kTF32 ? 4 : 2;
  const int bn = lstm_bwd_bn(H);
  const int tiles = ceil_div(rows, 128) * (H / bn) * D;
  const int splits = lstm_bwd_splits(rows, H, D, eb);
  *num_tickets = tiles;
  *ws_floats = splits > 1 ? static_cast<long>(tiles) * splits * 128 * bn : 0;
  return 0;
}

int dvae_lstm_wgrad_hh(int dtype, const void* da_all, const void* h_all, float* dwhh, int rows, int T, int H, int D,
                       float alpha, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (dtype == kF32) return simt_lstm_wgrad_hh((const float*)da_all, (const float*)h_all, dwhh, rows, T, H, D, alpha, st);
  DISPATCH_AT(dtype, lstm_wgrad_hh_t<AT>((const AT*)da_all, (const AT*)h_all, dwhh, rows, T, H, D, alpha, st));
}

int dvae_lstm_gate_tile(int H) { return lstm_fwd_bn(H); }

// kernel launches dvae_lstm_fwd / dvae_lstm_bwd enqueue for this shape (host-side launch accounting)
int dvae_lstm_launches(int H, int T, int backward) {
  if (lstm_seq_supported(H, T)) return 1;
  static const int fused = env_int("DVAE_LSTM_BWD_FUSED", 0);
  return backward ? (fused ? T : 2 * T - 1) : T;
}

// same, for the exact call: dvae_lstm_fwd picks the time-resident kernel (one launch per 1024 rows) when the shape allows it
int dvae_lstm_launches_for(int dtype, int rows, int T, int H, int D, int backward) {
  if (!backward && dtype != kF32 && !lstm_seq_supported(H, T) && lstm_res_supported(dtype, rows, T, H, D)) return ceil_div(rows, 1024);
  return dvae_lstm_launches(H, T, backward);
}

// 1 / 0: use / do not use the time-resident kernel of ops_lstm_res.cu for the H = 512 / 1024 forward recurrences (default 1, or
// DVAE_LSTM_RES); a negative value only queries.  Returns the previous setting.
int dvae_set_lstm_resident(int on) { return lstm_res_set_enabled(on); }

// 1: subsequent GEMM launches from this host thread are background work (see want_persistent); 0: normal
int dvae_set_background(int on) {
  // DVAE_BACKGROUND=1 restores the small-footprint background kernels for side-stream GEMMs.  Default 0: since the LSTM step
  // kernels own whole SMs (staging needs the full shared memory) there is little room to co-run; the wide persistent
  // weight-gradient kernels (1050-1090 TFLOP/s vs 630) finish sooner even though they serialise (17.76 vs 18.02 ms/step).
  static const int allow = env_int("DVAE_BACKGROUND", 0);
  if (!allow) on = 0;
  g_background = on;
  return 0;
}

// debugging aid: install (or remove, buf = NULL) a device buffer of capacity*5 uint64 phase stamps written by every
// tc_gemm_kernel CTA (see phase_stamp in tc_gemm.cuh)
int dvae_debug_timing(unsigned long long* buf, int capacity) {
  DVAE_CHECK_CUDA(cudaMemcpyToSymbol(dvae::g_phase_stamps, &buf, sizeof(buf)));
  DVAE_CHECK_CUDA(cudaMemcpyToSymbol(dvae::g_phase_capacity, &capacity, sizeof(capacity)));
  return 0;
}

// debug: per-step SM-clock stamps [T][8] of CTA (0,0) of the sequence-resident LSTM kernels (nullptr: off)
int dvae_debug_seq_stamps(long long* buf) {
  lstm_seq_set_stamps(buf);
  return 0;
}

// debug: globaltimer stamps [T][2 slots][8 points] of CTA 0 of the time-resident H = 512 / 1024 kernels (nullptr: off)
int dvae_debug_res_stamps(unsigned long long* buf) {
  lstm_res_set_stamps(buf);
  return 0;
}

}  // extern "C"
