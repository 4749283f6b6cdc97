// keyswitch.cu -- batched LWE keyswitch, big key (kN) -> small key (n)  (SURVEY.md A.4):
//   out = (0,...,0,b_in) - sum_j sum_lev dec_lev(a_j) * KSK[j][lev]
// Work per ciphertext: kN*l*(n+1) u64 MACs (7.6 M at the stated set); the 61 MB key is
// L2-resident.  Tiling: a CTA owns KS_TB ciphertexts x KS_COLS output columns and walks a
// slice of j; every KSK word it loads is used KS_TB times from registers.  Slices of j
// (split-K) fill the machine at small batch; partial sums meet through u64 atomics.
#include "common.cuh"
#include "kernels.h"

namespace fhe {

#ifndef KS_TB_VALUE
#define KS_TB_VALUE 8
#endif
constexpr int KS_TB = KS_TB_VALUE;  // ciphertexts per CTA
constexpr int KS_COLS = 128;  // output columns per CTA (one per thread)
constexpr int KS_JC = 32;     // j's decomposed per shared-memory refill
constexpr int KS_MAXL = 8;

__global__ void ks_init_kernel(const uint64_t* __restrict__ in, int64_t B, int64_t kN, int n,
                               uint64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (n + 1)) return;
    int64_t b = i / (n + 1);
    int w = (int)(i - b * (n + 1));
    out[i] = (w == n) ? in[b * (kN + 1) + kN] : 0;
}

__device__ __forceinline__ void ks_atomic_add(uint64_t* p, uint64_t v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ void ks_atomic_add(uint32_t* p, uint32_t v) { atomicAdd(p, v); }

// L = number of levels when known at compile time (0 = runtime l): with L fixed the L key words of one
// input coefficient are loaded back to back (independent loads in flight) before their MACs.
template <typename W, int L>
__global__ void __launch_bounds__(KS_COLS)
keyswitch_kernel(const W* __restrict__ ksk, const uint64_t* __restrict__ in, int64_t B, int64_t kN, int n,
                 int l_rt, int beta, int j_per_split, W* __restrict__ out) {
    __shared__ int32_t dig[KS_JC][KS_MAXL][KS_TB];
    const int l = L > 0 ? L : l_rt;
    const int col = blockIdx.x * KS_COLS + threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.y * KS_TB;
    const int64_t j0 = (int64_t)blockIdx.z * j_per_split;
    const int64_t j1 = min(j0 + (int64_t)j_per_split, kN);
    const int nb = (int)min((int64_t)KS_TB, B - b0);
    W acc[KS_TB];
#pragma unroll
    for (int t = 0; t < KS_TB; ++t) acc[t] = 0;
    const int tot = l * beta;
    const uint64_t Bm = (1ULL << beta) - 1, half = 1ULL << (beta - 1);
    uint64_t offs = 0;  // sum over digit positions of B/2: turns balanced digits into plain ones
    for (int lev = 0; lev < l; ++lev) offs |= half << (beta * lev);
    const size_t row = (size_t)(n + 1);
    for (int64_t jc = j0; jc < j1; jc += KS_JC) {
        __syncthreads();
        for (int e = threadIdx.x; e < KS_JC * KS_TB; e += KS_COLS) {
            const int jj = e / KS_TB, t = e - jj * KS_TB;
            const int64_t j = jc + jj;
            uint64_t a = (t < nb && j < j1) ? in[(b0 + t) * (kN + 1) + j] : 0;
            // closest representative on tot bits, then balanced base-2^beta digits
            uint64_t st = ((a + (1ULL << (63 - tot))) >> (64 - tot)) + offs;
            for (int lev = 0; lev < l; ++lev) {
                const int sh = beta * (l - 1 - lev);
                dig[jj][lev][t] = (t < nb && j < j1) ? (int32_t)((st >> sh) & Bm) - (int32_t)half : 0;
            }
        }
        __syncthreads();
        if (col <= n) {
            const int jn = (int)min((int64_t)KS_JC, j1 - jc);
            const W* kr = ksk + ((size_t)jc * l) * row + col;
            if (L > 0) {
                W kv[L > 0 ? L : 1], kn[L > 0 ? L : 1];
#pragma unroll
                for (int lev = 0; lev < L; ++lev) kv[lev] = kr[(size_t)lev * row];
                for (int jj = 0; jj < jn; ++jj) {
                    if (jj + 1 < jn) {  // next coefficient's key words are in flight behind these MACs
#pragma unroll
                        for (int lev = 0; lev < L; ++lev) kn[lev] = kr[((size_t)(jj + 1) * L + lev) * row];
                    }
#pragma unroll
                    for (int lev = 0; lev < L; ++lev) {
#pragma unroll
                        for (int t = 0; t < KS_TB; ++t) acc[t] -= (W)(int64_t)dig[jj][lev][t] * kv[lev];
                    }
#pragma unroll
                    for (int lev = 0; lev < L; ++lev) kv[lev] = kn[lev];
                }
            } else {
                for (int jj = 0; jj < jn; ++jj) {
                    for (int lev = 0; lev < l; ++lev) {
                        const W kv = kr[((size_t)jj * l + lev) * row];
#pragma unroll
                        for (int t = 0; t < KS_TB; ++t) acc[t] -= (W)(int64_t)dig[jj][lev][t] * kv;
                    }
                }
            }
        }
    }
    if (col <= n) {
        for (int t = 0; t < nb; ++t)
            if (acc[t]) ks_atomic_add(out + (b0 + t) * (n + 1) + col, acc[t]);
    }
}

template <typename W>
static void ks_dispatch(dim3 grid, cudaStream_t s, const W* ksk, const uint64_t* in, int64_t B, int64_t kN, int n, int l,
                        int beta, int j_per_split, W* out) {
    switch (l) {
        case 3: keyswitch_kernel<W, 3><<<grid, KS_COLS, 0, s>>>(ksk, in, B, kN, n, l, beta, j_per_split, out); break;
        case 4: keyswitch_kernel<W, 4><<<grid, KS_COLS, 0, s>>>(ksk, in, B, kN, n, l, beta, j_per_split, out); break;
        case 5: keyswitch_kernel<W, 5><<<grid, KS_COLS, 0, s>>>(ksk, in, B, kN, n, l, beta, j_per_split, out); break;
        default: keyswitch_kernel<W, 0><<<grid, KS_COLS, 0, s>>>(ksk, in, B, kN, n, l, beta, j_per_split, out); break;
    }
}

cudaError_t launch_keyswitch(const fhe_b200_pbs_params& p, const uint64_t* d_ksk, const uint64_t* d_in, int64_t B,
                             uint64_t* d_out, cudaStream_t s) {
    if (p.l_ks > KS_MAXL || p.l_ks * p.beta_ks > 62) return cudaErrorInvalidValue;
    const int64_t kN = (int64_t)p.k * p.N;
    const int64_t tot = B * (p.n + 1);
    ks_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_in, B, kN, p.n, d_out);
    count_launch();
    const int col_tiles = (p.n + 1 + KS_COLS - 1) / KS_COLS;
    const int64_t b_tiles = (B + KS_TB - 1) / KS_TB;
    if (b_tiles > 65535) return cudaErrorInvalidValue;
    // split-K so that at least ~4 CTAs per SM exist at small batch
    int splits = 1;
    while (splits < 64 && (int64_t)col_tiles * b_tiles * splits < 600 && (kN / (splits * 2)) >= KS_JC) splits *= 2;
    const int j_per_split = (int)((kN + splits - 1) / splits);
    dim3 grid(col_tiles, (unsigned)b_tiles, splits);
    ks_dispatch<uint64_t>(grid, s, d_ksk, d_in, B, kN, p.n, p.l_ks, p.beta_ks, j_per_split, d_out);
    count_launch();
    return cudaGetLastError();
}

// ---- 32-bit keyswitch ("KS32"): key and accumulation on the top 32 bits of the torus.  The output
// only has to be accurate to the small key's noise (2^-17 at the stated set); rounding the key to
// 32 bits adds ~2^-26.  One IMAD per MAC instead of three, half the key bytes.
__global__ void ksk_to_32_kernel(const uint64_t* __restrict__ ksk, int64_t words, uint32_t* __restrict__ ksk32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) ksk32[i] = (uint32_t)((ksk[i] + 0x80000000ULL) >> 32);
}

cudaError_t launch_ksk_to_32(const fhe_b200_pbs_params& p, const uint64_t* d_ksk, uint32_t* d_ksk32, cudaStream_t s) {
    const int64_t words = (int64_t)p.k * p.N * p.l_ks * (p.n + 1);
    ksk_to_32_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_ksk, words, d_ksk32);
    count_launch();
    return cudaGetLastError();
}

__global__ void ks32_init_kernel(const uint64_t* __restrict__ in, int64_t B, int64_t kN, int n, uint32_t* __restrict__ acc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (n + 1)) return;
    int64_t b = i / (n + 1);
    int w = (int)(i - b * (n + 1));
    acc[i] = (w == n) ? (uint32_t)((in[b * (kN + 1) + kN] + 0x80000000ULL) >> 32) : 0u;
}

__global__ void ks32_widen_kernel(const uint32_t* __restrict__ acc, int64_t words, uint64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) out[i] = (uint64_t)acc[i] << 32;
}

cudaError_t launch_keyswitch32(const fhe_b200_pbs_params& p, const uint32_t* d_ksk32, const uint64_t* d_in, int64_t B,
                               uint32_t* d_acc32, uint64_t* d_out, cudaStream_t s) {
    if (p.l_ks > KS_MAXL || p.l_ks * p.beta_ks > 62) return cudaErrorInvalidValue;
    const int64_t kN = (int64_t)p.k * p.N;
    const int64_t tot = B * (p.n + 1);
    ks32_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_in, B, kN, p.n, d_acc32);
    count_launch();
    const int col_tiles = (p.n + 1 + KS_COLS - 1) / KS_COLS;
    const int64_t b_tiles = (B + KS_TB - 1) / KS_TB;
    if (b_tiles > 65535) return cudaErrorInvalidValue;
    int splits = 1;
    while (splits < 64 && (int64_t)col_tiles * b_tiles * splits < 600 && (kN / (splits * 2)) >= KS_JC) splits *= 2;
    const int j_per_split = (int)((kN + splits - 1) / splits);
    dim3 grid(col_tiles, (unsigned)b_tiles, splits);
    ks_dispatch<uint32_t>(grid, s, d_ksk32, d_in, B, kN, p.n, p.l_ks, p.beta_ks, j_per_split, d_acc32);
    count_launch();
    ks32_widen_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_acc32, tot, d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fhe
