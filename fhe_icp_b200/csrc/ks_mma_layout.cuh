// ks_mma_layout.cuh -- index algebra of the tensor-core keyswitch (ks_mma.cu), __host__ __device__ so that
// tests/test_ks_mma_emul.py can rebuild the operand blocks on the CPU, walk them the way the MMA's shared-memory
// descriptors do (K-major, no swizzle: element (row, k) of a block at
//     (k / 16) * LBO + (row / 8) * SBO + (row % 8) * 16 + k % 16 )
// and check the contraction + byte-plane recombination against the oracle's 32-bit keyswitch without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#else
#define KM_HD inline
#endif

namespace fhe {
namespace kml {

constexpr int M_TILE = 128;                      // ciphertext rows per tile (UMMA M)
constexpr int N_TILE = 256;                      // byte columns per tile (UMMA N) = 64 output words
constexpr int K_BLOCK = 128;                     // k per pipeline stage
constexpr int UMMA_K = 32;                       // k per tcgen05.mma.kind::i8
constexpr int A_BYTES = M_TILE * K_BLOCK;        // 16 KB
constexpr int B_BYTES = N_TILE * K_BLOCK;        // 32 KB
constexpr uint32_t A_LBO = (M_TILE / 8) * 128;   // bytes between 16-byte k chunks of an A block (2048)
constexpr uint32_t B_LBO = (N_TILE / 8) * 128;   // ... of a B block (4096)
constexpr uint32_t SBO = 128;                    // bytes between 8-row groups

KM_HD int kblocks(int kN, int l) { return l * (kN / K_BLOCK); }
KM_HD int col_tiles(int n) { return (4 * (n + 1) + N_TILE - 1) / N_TILE; }

// K is level-major: k = lev * kN + j, so k-block kb holds coefficients j0 .. j0+127 of level lev
KM_HD int kb_of(int kN, int lev, int j) { return lev * (kN / K_BLOCK) + j / K_BLOCK; }

// byte offset, inside a block, of element (row, kk) with kk = k % 128
KM_HD uint32_t a_elem(int row, int kk) { return (uint32_t)(kk >> 4) * A_LBO + (uint32_t)(row >> 3) * SBO + (uint32_t)(row & 7) * 16 + (kk & 15); }
KM_HD uint32_t b_elem(int col, int kk) { return (uint32_t)(kk >> 4) * B_LBO + (uint32_t)(col >> 3) * SBO + (uint32_t)(col & 7) * 16 + (kk & 15); }

// byte column cc of column tile nt <-> (output word c, byte plane q)
KM_HD int word_of(int nt, int cc) { return nt * (N_TILE / 4) + (cc >> 2); }
KM_HD int plane_of(int cc) { return cc & 3; }

// balanced base-2^beta digit `lev` of torus word a (closest representative on l*beta bits first):
// the arithmetic of keyswitch_kernel / the oracle's keyswitch
KM_HD uint64_t digit_offsets(int l, int beta) {   // sum over digit positions of B/2: turns balanced digits into plain ones
    uint64_t offs = 0;
    for (int lev = 0; lev < l; ++lev) offs |= (1ULL << (beta - 1)) << (beta * lev);
    return offs;
}
KM_HD uint64_t digit_state(uint64_t a, int tot, uint64_t offs) { return ((a + (1ULL << (63 - tot))) >> (64 - tot)) + offs; }
KM_HD int digit_of(uint64_t state, int lev, int l, int beta) {
    const int sh = beta * (l - 1 - lev);
    return (int)((state >> sh) & ((1ULL << beta) - 1)) - (int)(1ULL << (beta - 1));
}

struct alignas(16) chunk16 { uint32_t w[4]; };   // one 16-byte k chunk of one row / byte column

// ---- the two block builders, one call per 16-byte chunk (a GPU thread each; the CPU emulation loops)
// key bytes: chunk g = (nt, kb, k16, n8, r) -> 16 consecutive k of byte-column cc = n8*8 + r of column tile nt
KM_HD void build_b_chunk(int64_t g, const uint32_t* ksk32, int kN, int l, int n, uint8_t* tiles) {
    const int kbs = kblocks(kN, l);
    const int r = (int)(g & 7);
    const int n8 = (int)((g >> 3) % (N_TILE / 8));
    int64_t rest = (g >> 3) / (N_TILE / 8);
    const int k16 = (int)(rest % (K_BLOCK / 16));
    rest /= (K_BLOCK / 16);
    const int kb = (int)(rest % kbs);
    const int nt = (int)(rest / kbs);
    const int cc = n8 * 8 + r;
    const int c = word_of(nt, cc), q = plane_of(cc);
    const int lev = kb / (kN / K_BLOCK);
    const int j0 = (kb % (kN / K_BLOCK)) * K_BLOCK + k16 * 16;
    chunk16 v = {{0, 0, 0, 0}};
    if (c <= n) {
        for (int i = 0; i < 16; ++i) {
            const uint32_t x = ksk32[((size_t)(j0 + i) * l + lev) * (size_t)(n + 1) + c];
            v.w[i >> 2] |= ((x >> (8 * q)) & 0xFFu) << (8 * (i & 3));
        }
    }
    *reinterpret_cast<chunk16*>(tiles + ((size_t)nt * kbs + kb) * B_BYTES + b_elem(cc, k16 * 16)) = v;
}
// digits: work item g = (mt, 16 consecutive coefficients, row) -> l chunks of 16 digit bytes, one per level.
// Consecutive g walk the rows of a tile first, so a warp's stores are contiguous.
KM_HD void build_a_chunks(int64_t g, const uint64_t* in, int64_t B, int kN, int l, int beta, int8_t* a_tiles) {
    const int chunks_per_row = kN / 16;
    const int r = (int)(g % M_TILE);
    int64_t rest = g / M_TILE;
    const int ch = (int)(rest % chunks_per_row);
    const int64_t mt = rest / chunks_per_row;
    const int64_t b = mt * M_TILE + r;
    const int jb = ch / (K_BLOCK / 16), k16 = ch % (K_BLOCK / 16);
    const uint64_t offs = digit_offsets(l, beta);
    uint64_t st[16];
    for (int i = 0; i < 16; ++i) st[i] = digit_state(b < B ? in[b * (int64_t)(kN + 1) + ch * 16 + i] : 0, l * beta, offs);
    for (int lev = 0; lev < l; ++lev) {
        chunk16 v = {{0, 0, 0, 0}};
        if (b < B)
            for (int i = 0; i < 16; ++i) v.w[i >> 2] |= ((uint32_t)digit_of(st[i], lev, l, beta) & 0xFFu) << (8 * (i & 3));
        const int kb = kb_of(kN, lev, jb * K_BLOCK);
        *reinterpret_cast<chunk16*>(a_tiles + ((size_t)(mt * kblocks(kN, l) + kb)) * A_BYTES + a_elem(r, k16 * 16)) = v;
    }
}

// the four s32 byte-plane sums of one output word -> the 32-bit accumulator contribution (mod 2^32)
KM_HD uint32_t recombine(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) { return c0 + (c1 << 8) + (c2 << 16) + (c3 << 24); }

}  // namespace kml
}  // namespace fhe
