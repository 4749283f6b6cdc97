// ks_mma.cu -- the 32-bit LWE keyswitch as a dense int8 contraction on the 5th-generation tensor cores.
//
//   acc32[b][c] = hi32(b_in) . [c == n]  -  sum_{j < kN} sum_{lev < l} digit(b, j, lev) * ksk32[j][lev][c]   (mod 2^32)
//
// (the "KS32" form of keyswitch.cu).  The digits are balanced base-2^beta
// integers in [-2^(beta-1), 2^(beta-1)) -- int8 -- and every 32-bit key word is four unsigned bytes, so
//
//   C[b][4c + q] = sum_k digit[b][k] * byte_q(ksk32[k][c])        is a  [B x K] . [K x 4(n+1)]  int8 GEMM
//   acc32[b][c]  = init - sum_q (C[b][4c + q] << 8q)               (K = kN*l = 10240, |C| <= K*4*255 < 2^31)
//
// which is exactly what `tcgen05.mma.kind::i8` computes (s8 x u8 -> s32 in tensor memory).  This is the one
// dense contraction on the path (the external product's inner dimension is 2, the keyswitch's is
// 10240); the result is bit-identical to keyswitch_kernel<uint32_t> because every step is exact integer
// arithmetic.
//
// Data layout.  Both operands are stored in global memory ALREADY in the shared-memory layout the MMA reads
// (K-major, no swizzle: 8-row x 16-byte core matrices, 128 contiguous bytes each), one contiguous block per
// (tile, k-block), so a pipeline stage is filled by two 1-D TMA bulk copies and needs no tensor map:
//   A block [mt][kb] : 128 ciphertext rows x 128 k   = [k16 (8)][r8 (16)][8 rows][16 B]   16 KB  (ks_digits_kernel)
//   B block [nt][kb] : 256 byte-columns  x 128 k     = [k16 (8)][n8 (32)][8 cols][16 B]   32 KB  (built once per key)
// K is ordered level-major (k = lev*kN + j) so a k-block is 128 consecutive coefficients of one level.
//
// Kernel: one CTA per (row tile, column tile).  Warp 0: TMA producer (one lane), warp 1: MMA issuer (one lane;
// owns the TMEM allocation), warps 2..5: epilogue (TMEM -> registers -> recombine the four byte planes -> u64
// ciphertext words).  4-stage full/empty mbarrier ring; `tcgen05.commit` releases a stage when its MMAs retire.
#include "common.cuh"
#include "kernels.h"
#include "ks_mma_layout.cuh"

namespace fhe {
namespace {

constexpr int KM_M = kml::M_TILE;               // ciphertext rows per tile (UMMA M)
constexpr int KM_N = kml::N_TILE;               // byte columns per tile (UMMA N) = 64 output words
constexpr int KM_KB = kml::K_BLOCK;             // k per pipeline stage
constexpr int KM_UMMA_K = kml::UMMA_K;          // k per tcgen05.mma.kind::i8
constexpr int KM_STAGES = 4;
constexpr int KM_A_BYTES = kml::A_BYTES;        // 16 KB
constexpr int KM_B_BYTES = kml::B_BYTES;        // 32 KB
constexpr int KM_THREADS = 192;
constexpr uint32_t KM_TMEM_COLS = 256;
constexpr uint32_t KM_A_LBO = kml::A_LBO;       // bytes between 16-byte k chunks of A (2048)
constexpr uint32_t KM_B_LBO = kml::B_LBO;       // ... of B (4096)
constexpr uint32_t KM_SBO = kml::SBO;           // bytes between 8-row groups
constexpr size_t KM_SMEM = (size_t)KM_STAGES * (KM_A_BYTES + KM_B_BYTES) + 256;

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t km_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
// instruction descriptor: D = s32, A = signed 8-bit (digits), B = unsigned 8-bit (key bytes), both K-major
constexpr uint32_t KM_IDESC = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(KM_N >> 3) << 17) | ((uint32_t)(KM_M >> 4) << 24);

__device__ __forceinline__ void km_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(KM_IDESC), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void km_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void km_tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void km_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void km_tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace

// ---- key bytes in MMA block layout (once per key) and digits in MMA block layout (per batch): one thread per
// 16-byte chunk / per (row, 16 coefficients); the builders live in ks_mma_layout.cuh (shared with the CPU emulation)
__global__ void ksk32_to_mma_kernel(const uint32_t* __restrict__ ksk32, int kN, int l, int n, int64_t chunks,
                                    uint8_t* __restrict__ tiles) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < chunks) kml::build_b_chunk(g, ksk32, kN, l, n, tiles);
}

__global__ void ks_digits_kernel(const uint64_t* __restrict__ in, int64_t B, int kN, int l, int beta, int64_t items,
                                 int8_t* __restrict__ a_tiles) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < items) kml::build_a_chunks(g, in, B, kN, l, beta, a_tiles);
}

// ---- the contraction
__global__ void __launch_bounds__(KM_THREADS, 1)
ks_mma_kernel(const int8_t* __restrict__ a_tiles, const uint8_t* __restrict__ b_tiles, const uint64_t* __restrict__ in,
              int64_t B, int kN, int n, int kblocks, uint64_t* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)KM_STAGES * KM_A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)KM_STAGES * (KM_A_BYTES + KM_B_BYTES));
    uint64_t* empty = full + KM_STAGES;
    uint64_t* acc_done = empty + KM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.x, mt = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KM_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) km_tmem_alloc(tmem_slot, KM_TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer
            const int8_t* ga = a_tiles + (size_t)mt * kblocks * KM_A_BYTES;
            const uint8_t* gb = b_tiles + (size_t)nt * kblocks * KM_B_BYTES;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % KM_STAGES;
                mbar_wait(empty + s, (uint32_t)(((kb / KM_STAGES) & 1) ^ 1));
                mbar_expect_tx(full + s, KM_A_BYTES + KM_B_BYTES);
                tma_load_1d(sA + (size_t)s * KM_A_BYTES, ga + (size_t)kb * KM_A_BYTES, KM_A_BYTES, full + s);
                tma_load_1d(sB + (size_t)s * KM_B_BYTES, gb + (size_t)kb * KM_B_BYTES, KM_B_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % KM_STAGES;
                mbar_wait(full + s, (uint32_t)((kb / KM_STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + (size_t)s * KM_A_BYTES), b0 = smem_u32(sB + (size_t)s * KM_B_BYTES);
#pragma unroll
                for (int kk = 0; kk < KM_KB / KM_UMMA_K; ++kk) {
                    const uint64_t ad = km_smem_desc(a0 + kk * 2 * KM_A_LBO, KM_A_LBO, KM_SBO);
                    const uint64_t bd = km_smem_desc(b0 + kk * 2 * KM_B_LBO, KM_B_LBO, KM_SBO);
                    km_mma_i8(tmem_base, ad, bd, (kb | kk) != 0 ? 1u : 0u);
                }
                km_commit(empty + s);   // the stage is free once these MMAs have read it
            }
            km_commit(acc_done);        // accumulator complete
        }
    } else {  // ===== epilogue: warp w may touch TMEM lanes 32*(w % 4) .. +31
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int64_t b = (int64_t)mt * KM_M + row;
        mbar_wait(acc_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t body = b < B ? (uint32_t)((in[b * (int64_t)(kN + 1) + kN] + 0x80000000ULL) >> 32) : 0u;
        uint64_t* orow = out + b * (int64_t)(n + 1);
#pragma unroll 1
        for (int ch = 0; ch < KM_N / 16; ++ch) {
            uint32_t r[16];
            km_tmem_ld_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ch * 16), r);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = kml::word_of(nt, ch * 16 + 4 * i);
                const uint32_t v = kml::recombine(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                if (b < B && c <= n) orow[c] = (uint64_t)((c == n ? body : 0u) - v) << 32;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) km_tmem_dealloc(tmem_base, KM_TMEM_COLS);
}

static int km_col_tiles(int n) { return kml::col_tiles(n); }

bool keyswitch_mma_supported(const fhe_b200_pbs_params& p) {
    const int64_t kN = (int64_t)p.k * p.N;
    return kN % KM_KB == 0 && p.beta_ks >= 1 && p.beta_ks <= 8 && p.l_ks >= 1 && p.l_ks * p.beta_ks <= 62 &&
           (int64_t)p.l_ks * kN * (1LL << (p.beta_ks - 1)) * 255 < (1LL << 31);   // |digit| <= 2^(beta-1), byte <= 255: s32 never wraps
}

size_t keyswitch_mma_key_bytes(const fhe_b200_pbs_params& p) {
    const int64_t kN = (int64_t)p.k * p.N;
    return (size_t)km_col_tiles(p.n) * (size_t)(p.l_ks * (kN / KM_KB)) * KM_B_BYTES;
}

size_t keyswitch_mma_workspace_bytes(const fhe_b200_pbs_params& p, int64_t B) {
    const int64_t kN = (int64_t)p.k * p.N;
    return (size_t)((B + KM_M - 1) / KM_M) * (size_t)(p.l_ks * (kN / KM_KB)) * KM_A_BYTES;
}

cudaError_t launch_ksk32_to_mma(const fhe_b200_pbs_params& p, const uint32_t* d_ksk32, uint8_t* d_tiles, cudaStream_t s) {
    if (!keyswitch_mma_supported(p)) return cudaErrorInvalidValue;
    const int kN = p.k * p.N;
    const int64_t chunks = (int64_t)(keyswitch_mma_key_bytes(p) / 16);
    ksk32_to_mma_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(d_ksk32, kN, p.l_ks, p.n, chunks, d_tiles);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_keyswitch_mma(const fhe_b200_pbs_params& p, const uint8_t* d_tiles, const uint64_t* d_in, int64_t B,
                                 int8_t* d_work, uint64_t* d_out, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    if (!keyswitch_mma_supported(p)) return cudaErrorInvalidValue;
    const int kN = p.k * p.N;
    const int64_t m_tiles = (B + KM_M - 1) / KM_M;
    if (m_tiles > 65535) return cudaErrorInvalidValue;
    const int64_t dthreads = m_tiles * KM_M * (kN / 16);
    ks_digits_kernel<<<(unsigned)((dthreads + 255) / 256), 256, 0, s>>>(d_in, B, kN, p.l_ks, p.beta_ks, dthreads, d_work);
    count_launch();
    cudaError_t e = cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KM_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)km_col_tiles(p.n), (unsigned)m_tiles);
    ks_mma_kernel<<<grid, KM_THREADS, KM_SMEM, s>>>(d_work, d_tiles, d_in, B, kN, p.n, p.l_ks * (kN / KM_KB), d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fhe
