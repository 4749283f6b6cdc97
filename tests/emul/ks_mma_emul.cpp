// CPU emulation of the tensor-core keyswitch (fhe_icp_b200/csrc/ks_mma.cu): the operand blocks are built by the
// product's own builders (ks_mma_layout.cuh is __host__ __device__), then read back the way the MMA's
// shared-memory descriptors address them -- K-major, no swizzle, element (row, k) at
// (k/16)*LBO + (row/8)*SBO + (row%8)*16 + k%16 -- contracted per (row tile, column tile, k block) as s8 x u8 -> s32,
// and recombined exactly as the kernel's epilogue does.  tests/test_ks_mma_emul.py compares with the oracle.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../fhe_icp_b200/csrc/ks_mma_layout.cuh"

using namespace fhe::kml;

extern "C" int emul_keyswitch_mma(const uint32_t* ksk32, const uint64_t* in, int64_t B, int kN, int n, int l, int beta,
                                  uint64_t* out) {
    if (kN % K_BLOCK != 0 || beta < 1 || beta > 8) return 1;
    const int kbs = kblocks(kN, l), nts = col_tiles(n);
    const int64_t mts = (B + M_TILE - 1) / M_TILE;
    std::vector<uint8_t> bt((size_t)nts * kbs * B_BYTES, 0xAB);       // poison: every byte must be written
    std::vector<int8_t> at((size_t)mts * kbs * A_BYTES, (int8_t)0x5A);
    for (int64_t g = 0; g < (int64_t)bt.size() / 16; ++g) build_b_chunk(g, ksk32, kN, l, n, bt.data());
    for (int64_t g = 0; g < mts * M_TILE * (kN / 16); ++g) build_a_chunks(g, in, B, kN, l, beta, at.data());
    std::vector<int32_t> acc((size_t)M_TILE * N_TILE);
    for (int64_t mt = 0; mt < mts; ++mt) {
        for (int nt = 0; nt < nts; ++nt) {
            std::fill(acc.begin(), acc.end(), 0);
            for (int kb = 0; kb < kbs; ++kb) {
                const int8_t* a = at.data() + ((size_t)mt * kbs + kb) * A_BYTES;
                const uint8_t* b = bt.data() + ((size_t)nt * kbs + kb) * B_BYTES;
                for (int k0 = 0; k0 < K_BLOCK; k0 += UMMA_K)          // one tcgen05.mma per 32 k
                    for (int row = 0; row < M_TILE; ++row)
                        for (int col = 0; col < N_TILE; ++col) {
                            int32_t s = 0;
                            for (int kk = k0; kk < k0 + UMMA_K; ++kk) s += (int32_t)a[a_elem(row, kk)] * (int32_t)b[b_elem(col, kk)];
                            acc[(size_t)row * N_TILE + col] += s;
                        }
            }
            for (int row = 0; row < M_TILE; ++row) {                 // epilogue: TMEM lane = row, 16 columns per load
                const int64_t r = mt * M_TILE + row;
                if (r >= B) continue;
                const uint32_t body = (uint32_t)((in[r * (int64_t)(kN + 1) + kN] + 0x80000000ULL) >> 32);
                for (int ch = 0; ch < N_TILE / 16; ++ch)
                    for (int i = 0; i < 4; ++i) {
                        const int c = word_of(nt, ch * 16 + 4 * i);
                        const int32_t* p = &acc[(size_t)row * N_TILE + ch * 16 + 4 * i];
                        const uint32_t v = recombine((uint32_t)p[0], (uint32_t)p[1], (uint32_t)p[2], (uint32_t)p[3]);
                        if (c <= n) out[r * (int64_t)(n + 1) + c] = (uint64_t)((c == n ? body : 0u) - v) << 32;
                    }
            }
        }
    }
    return 0;
}
