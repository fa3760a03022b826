// Shared-memory matrix descriptors and the instruction descriptor of tcgen05.mma (sm_100a), shared by the tiled GEMM
// kernels (tc_gemm.cuh) and the sequence-resident LSTM kernels (ops_lstm_seq.cu).
#pragma once
#include <stdint.h>

namespace dvae {

constexpr int kBlockM = 128;      // every MMA in this library is M = 128 (cta_group::1, one TMEM lane per row)
constexpr int kSwizzleRow = 128;  // bytes per swizzle-128B row == BLOCK_K * ELEM_BYTES

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  // sm_100 shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
  // version=1 [46,48) | layout [61,64): SWIZZLE_128B = 2, SWIZZLE_128B_BASE32B = 1
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout_type) << 61);
}

// c_format F32 [4,6) | a_format [7,10) | b_format [10,13) | a_major 15 | b_major 16 | N>>3 [17,23) | M>>4 [24,29)
// FMT: kind::f16 0 = F16, 1 = BF16; kind::tf32 2 = TF32.  UMMA_M = 256 is the CTA-pair MMA (cta_group::2)
template <uint32_t FMT, int BLOCK_N, bool A_MN, bool B_MN, int UMMA_M = kBlockM>
__host__ __device__ constexpr uint32_t instr_desc_fmt() {
  return (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(BLOCK_N >> 3) << 17) | (static_cast<uint32_t>(UMMA_M >> 4) << 24);
}
// by element size: 2 bytes = BF16 (fp16 launches clear the format fields at run time, see runtime_idesc), 4 = TF32
template <int ELEM_BYTES, int BLOCK_N, bool A_MN, bool B_MN, int UMMA_M = kBlockM>
__host__ __device__ constexpr uint32_t instr_desc() {
  return instr_desc_fmt<(ELEM_BYTES == 2) ? 1u : 2u, BLOCK_N, A_MN, B_MN, UMMA_M>();
}

}  // namespace dvae
