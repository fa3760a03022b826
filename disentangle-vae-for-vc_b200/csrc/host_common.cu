#include "host_common.h"

#include <cudaTypedefs.h>

#include <mutex>

namespace dvae {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

static int encode_map3_impl(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                            uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool mn_major,
                            bool narrow);

int encode_map3(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool mn_major) {
  return encode_map3_impl(out, base, elem_bytes, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, b2, mn_major, false);
}
int encode_map3_narrow(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  return encode_map3_impl(out, base, elem_bytes, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, b2, false, true);
}

static int encode_map3_impl(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                            uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool mn_major,
                            bool narrow) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 2;
  }
  DVAE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
  DVAE_REQUIRE(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides must be multiples of 16 bytes");
  if (narrow) DVAE_REQUIRE(b0 * elem_bytes < 128 && (b0 * elem_bytes) % 16 == 0, "narrow box: inner extent must be a multiple of 16 bytes below 128");
  else DVAE_REQUIRE(b0 * elem_bytes == 128, "inner box must span one 128-byte swizzle row");
  DVAE_REQUIRE(b1 <= 256 && b2 <= 256 && b1 >= 1 && b2 >= 1, "TMA box dims must be in [1,256]");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = narrow ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : (mn_major && elem_bytes == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                                : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u) eb=%d",
             static_cast<int>(r), (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, b0, b1, b2, elem_bytes);
    set_last_error(buf);
    return 2;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}

}  // namespace dvae

extern "C" {
// Bytes of caller-provided scratch an op needs (nothing in this library allocates): `op` names the buffer, n0..n2 its shape
// parameters.  Returns -1 for an unknown name.
//   "bn_stats"     (halves, C)      double sums of dvae_conv5_fwd_bnstats / dvae_bn_train_fwd / dvae_bn_train_bwd (`ws`)
//   "bn_stat"      (halves, C)      fp32 mean / rstd / scale / shift kept for the backward (`stat`)
//   "bn_bwd_coef"  (halves, C)      fp32 coefficients of dvae_bn_train_bwd (`coef`)
//   "loss"         ()               accumulators + ticket of dvae_loss_fwd (`ws`)
//   "lstm_bwd_dc"  (rows, H, D)     dc carry + dh_rec of dvae_lstm_bwd (`dc_ws`)
//   "segment_ids"  (B)              block counts of dvae_segment_ids_sorted (`scratch`)
//   "group_acc"    (G, D)           per-group accumulators of dvae_group_accumulate (`acc`; `cnt` is 4 * G more)
long dvae_workspace_bytes(const char* op, long n0, long n1, long n2) {
  const std::string s(op ? op : "");
  if (s == "bn_stats") return 8 * (n0 * 2 * n1 + 1);
  if (s == "bn_stat") return 4 * n0 * 4 * n1;
  if (s == "bn_bwd_coef") return 4 * n0 * 2 * n1;
  if (s == "loss") return 72;
  if (s == "lstm_bwd_dc") return 4 * 2 * n2 * n0 * n1;
  if (s == "segment_ids") return 4 * ((n0 + 1023) / 1024 > 0 ? (n0 + 1023) / 1024 : 1);
  if (s == "group_acc") return 4 * n0 * 2 * n1;
  return -1;
}
const char* dvae_last_error() { return dvae::last_error_cstr(); }
int dvae_version() { return 100; }
int dvae_sm_arch() { return 100; }  // built for sm_100a only
}
