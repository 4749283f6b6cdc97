// lwe.cu -- LWE over the 64-bit torus: key sampling, batched encrypt / decrypt, and the
// batched encrypted dot product (LWE linear combination) that is the reference's whole
// compiled circuit (Concrete-ML LinearRegression, /root/reference/fhe_similarity.py:88-90,151;
// SURVEY.md Appendix A.3).  All arithmetic is wrapping u64 and therefore bit-exact.
//
// Roofline: every kernel here is HBM-bound (0.125 u64 MAC per byte for the dot product).
// Layout: ciphertext rows of `stride` u64 words (stride even => every row is 16-byte
// aligned), so a warp reads 512 contiguous bytes per 128-bit load instruction.
#include <algorithm>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "lwe_device.cuh"

namespace fhe {

// ----------------------------------------------------------------------------- secret key
// bit j of key `key_id` = bit (j%32) of word (j%128)/32 of Philox block j/128.
__global__ void secret_key_kernel(uint64_t key_seed, uint32_t key_id, int64_t dim, uint8_t* __restrict__ key) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    u32x4 r = rng_block(key_seed, FHE_B200_KIND_SK | (key_id << 8), 0, (uint32_t)(j >> 7));
    uint32_t w = (uint32_t)(j & 127) >> 5;
    uint32_t word = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
    key[j] = (word >> (j & 31)) & 1u;
}

cudaError_t launch_secret_key(uint64_t key_seed, uint32_t key_id, int64_t dim, uint8_t* d_key, cudaStream_t s) {
    if (dim <= 0) return cudaSuccess;
    secret_key_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, s>>>(key_seed, key_id, dim, d_key);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- encrypt
// Two kernels.  (1) lwe_body_noise_kernel, one THREAD per ciphertext: plaintext + rounded Gaussian error into the
// body slot (the Box-Muller chain is ~250 dependent FP64 instructions, so it runs lane-parallel).  (2) one WARP
// per ciphertext: lane L generates Philox blocks L, L+32, ... (two mask words each), stores them with one 128-bit
// store (512 B per warp instruction), accumulates <a, s> against the key bits held in shared memory and adds it
// to the body.  The unit of work is ONE ciphertext and the grid is the number of resident warps, so the static
// round-robin is balanced to a few percent (with 32-ciphertext units, 128 k ciphertexts filled only 4000 of
// 4736 resident warps and the kernel ran at the pace of the fullest SMs: 15 % lost).
constexpr int ENC_WARPS = 8;
#ifndef LCS_UNROLL
#define LCS_UNROLL 4
#endif
constexpr int LCS_UNROLL_N = LCS_UNROLL;   // independent Philox blocks in flight per thread (seeded dot product)
#ifndef LCS_MIN_CTAS
#define LCS_MIN_CTAS 5      // resident CTAs per SM the seeded dot product is compiled for (A/B: tools/build_variant.py)
#endif
#ifndef ENC_SEEDED_UNROLL
#define ENC_SEEDED_UNROLL 4
#endif
constexpr int ENC_SEEDED_UNROLL_N = ENC_SEEDED_UNROLL;   // ... per lane (seeded encryption)

// The error terms come from the client's SECRET noise seed -- never from the public mask seed that travels with seeded
// ciphertexts: an evaluator who could regenerate e would learn <a,s> + Delta*m exactly and solve for the key.
__global__ void lwe_body_noise_kernel(const int64_t* __restrict__ msgs, int64_t count, int shift, double sigma_abs,
                                      uint64_t noise_seed, uint64_t ct_base, uint32_t purpose, uint64_t* __restrict__ dst,
                                      int64_t dst_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t ndom = FHE_B200_KIND_NOISE | (purpose << 8);
    dst[i * dst_stride] = ((uint64_t)msgs[i] << shift) + (uint64_t)gaussian_i64(noise_seed, ndom, ct_base + (uint64_t)i, 0, sigma_abs);
}

template <typename Kernel>
static unsigned enc_grid(Kernel kernel, size_t smem, int64_t count) {
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ENC_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 4;
    const int64_t resident = (int64_t)sms * per_sm;
    return (unsigned)std::min<int64_t>((count + ENC_WARPS - 1) / ENC_WARPS, resident);
}

__global__ void __launch_bounds__(ENC_WARPS * 32)
lwe_encrypt_kernel(const uint8_t* __restrict__ key, int n, int64_t stride, int64_t count, uint64_t enc_seed,
                   uint64_t ct_base, uint32_t purpose, uint64_t* __restrict__ out) {
    extern __shared__ uint32_t skey[];  // packed key bits, ceil(n/32) words (+1 pad)
    pack_key_bits(key, n, skey);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t c = (int64_t)blockIdx.x * ENC_WARPS + (threadIdx.x >> 5); c < count; c += (int64_t)gridDim.x * ENC_WARPS) {
        uint64_t* ct = out + c * stride;
        const uint64_t pre = lane == 0 ? ct[n] : 0;   // plaintext + error, written by lwe_body_noise_kernel
        warp_lwe_encrypt(skey, n, stride, pre, 0, enc_seed, purpose, ct_base + (uint64_t)c, ct, lane);
    }
}

// Variant for a pipeline stage that runs NEXT TO the HBM-bound dot product of the previous chunk: key bits come
// pre-packed from global memory (no per-CTA prologue) and a CTA encrypts ENC_WARPS ciphertexts and exits, so SM
// slots keep turning over and the higher-priority dot-product CTAs get the share they need (a persistent grid
// would hold every slot until the whole chunk is encrypted).
__global__ void __launch_bounds__(ENC_WARPS * 32)
lwe_encrypt_packed_kernel(const uint32_t* __restrict__ kbits, int n, int64_t stride, int64_t count, uint64_t enc_seed,
                          uint64_t ct_base, uint32_t purpose, uint64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * ENC_WARPS + (threadIdx.x >> 5);
    if (c >= count) return;
    uint64_t* ct = out + c * stride;
    const uint64_t pre = lane == 0 ? ct[n] : 0;
    warp_lwe_encrypt(kbits, n, stride, pre, 0, enc_seed, purpose, ct_base + (uint64_t)c, ct, lane);
}

cudaError_t launch_lwe_encrypt_packed(const uint32_t* d_kbits, int n, int64_t stride, const int64_t* d_msgs, int64_t count,
                                      int shift, double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                                      uint32_t purpose, uint64_t* d_ct, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t grid = (count + ENC_WARPS - 1) / ENC_WARPS;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    lwe_body_noise_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_msgs, count, shift, sigma_abs, noise_seed, ct_base,
                                                                        purpose, d_ct + n, stride);
    count_launch();
    lwe_encrypt_packed_kernel<<<(unsigned)grid, ENC_WARPS * 32, 0, s>>>(d_kbits, n, stride, count, enc_seed, ct_base, purpose, d_ct);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_lwe_encrypt(const uint8_t* d_key, int n, int64_t stride, const int64_t* d_msgs, int64_t count,
                               int shift, double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                               uint32_t purpose, uint64_t* d_ct, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    lwe_body_noise_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_msgs, count, shift, sigma_abs, noise_seed, ct_base,
                                                                        purpose, d_ct + n, stride);
    count_launch();
    size_t smem = ((size_t)(n + 31) / 32 + 1) * sizeof(uint32_t);
    lwe_encrypt_kernel<<<enc_grid(lwe_encrypt_kernel, smem, count), ENC_WARPS * 32, smem, s>>>(d_key, n, stride, count, enc_seed,
                                                                                              ct_base, purpose, d_ct);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- phase / decrypt
// One warp per ciphertext: mu = b - <a, s>;  decode: m = (mu + Delta/2) >> shift (arithmetic).
constexpr int DEC_WARPS = 8;

__global__ void __launch_bounds__(DEC_WARPS * 32)
lwe_phase_kernel(const uint8_t* __restrict__ key, int n, int64_t stride, const uint64_t* __restrict__ cts,
                 int64_t count, int shift, bool decode, uint64_t* __restrict__ out) {
    extern __shared__ uint32_t skey[];
    pack_key_bits(key, n, skey);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (c >= count) return;
    const uint64_t* ct = cts + c * stride;
    uint64_t dot = 0;
    const int nvec = n / 2;  // full pairs of mask words
#pragma unroll 4
    for (int v = lane; v < nvec; v += 32) {
        u64x2 a = ld_stream_u64x2(ct + 2 * v);
        uint32_t bits = skey[(2 * v) >> 5] >> ((2 * v) & 31);
        dot += a.x & (0 - (uint64_t)(bits & 1u));
        dot += a.y & (0 - (uint64_t)((bits >> 1) & 1u));
    }
    if ((n & 1) && lane == 0) {
        int w = n - 1;
        dot += ct[w] & (0 - (uint64_t)((skey[w >> 5] >> (w & 31)) & 1u));
    }
    dot = warp_sum_u64(dot);
    if (lane == 0) {
        uint64_t mu = ct[n] - dot;
        if (decode) {
            uint64_t v = mu + (shift > 0 ? (1ULL << (shift - 1)) : 0ULL);
            out[c] = (uint64_t)((int64_t)v >> shift);
        } else {
            out[c] = mu;
        }
    }
}

cudaError_t launch_lwe_phase(const uint8_t* d_key, int n, int64_t stride, const uint64_t* d_ct, int64_t count,
                             int shift, bool decode, uint64_t* d_out, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    size_t smem = ((size_t)(n + 31) / 32 + 1) * sizeof(uint32_t);
    unsigned grid = (unsigned)((count + DEC_WARPS - 1) / DEC_WARPS);
    lwe_phase_kernel<<<grid, DEC_WARPS * 32, smem, s>>>(d_key, n, stride, d_ct, count, shift, decode, d_out);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- encrypted dot product
// out[b][m][:] = sum_j W[m][j] * ct[b][j][:].  The (row b, 16-byte column pair) space is
// flattened over the grid, so no thread idles whatever the row length.  Each thread streams
// the d ciphertexts of its row (stride*8 bytes apart) through a register double buffer:
// UNROLL 128-bit loads for rows j+UNROLL.. are issued before the MACs of rows j.. retire, which
// keeps UNROLL independent loads in flight per thread (measured: 5.3 -> 6.9 TB/s, see
// profiles/r1_lincomb_variant_sweep.txt; a TMA/mbarrier ring reached 6.8 TB/s).  Inputs are
// read exactly once (L1::no_allocate), the M output rows are written once.
constexpr int LC_THREADS = 256;
constexpr int LC_UNROLL = 8;

// ---- pushed scores (multi-GPU search) ---------------------------------------------------------------
// With a PushArgs destination the dot-product kernels do the gather themselves: every thread writes its
// two finished score words in the 32-bit wire form (modulus switch 2^64 -> 2^32, exactly
// lwe_modswitch32_kernel) straight into the CLIENT GPU's score board -- a peer mapping of the client's
// memory, so the stores travel over NVLink while the other CTAs are still streaming ciphertexts -- and
// the last CTA to finish publishes `step` in the client's arrival flag (release at system scope).
// Flow control: the board has two slots; the caller orders the launch behind a one-warp wait
// (peer_wait_kernel) on the credit flag the client bumps once it has decrypted the slot's previous
// contents.  The wait is NOT done inside this kernel: spinning CTAs would hold their SM slots and could
// lock the client's own decrypt kernels out of the GPU that has to produce the credit.
struct PushArgs {
    uint32_t* board;         // peer: this rank's rows of the slot, [B][M][stride] u32
    uint64_t* arrive;        // peer: arrival flag of (slot, rank) in the client's memory
    uint64_t step;           // value published on arrival (monotonic, > 0)
    uint32_t* counter;       // local: finished-CTA counter, left at zero
};

__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded spin: false on time-out (never hangs the GPU if a peer died)
__device__ __noinline__ bool spin_until_ge(const uint64_t* flag, uint64_t need, uint64_t timeout_ns) {
    if (ld_acquire_sys_u64(flag) >= need) return true;
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire_sys_u64(flag) < need) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(256);
    }
    return true;
}
__device__ __forceinline__ void push_store(uint32_t* o, uint64_t x, uint64_t y) {
    uint2 v;
    v.x = (uint32_t)((x + 0x80000000ULL) >> 32);
    v.y = (uint32_t)((y + 0x80000000ULL) >> 32);
    *reinterpret_cast<uint2*>(o) = v;
}
// every thread of the CTA calls this after its stores.  The barrier orders the CTA's stores before thread 0,
// whose system-scope fence is cumulative over them (the grid-sync pattern): one fence per CTA, not one per thread.
__device__ __forceinline__ void push_arrive(const PushArgs& p) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(p.counter, 1u);
        if (prev == gridDim.x - 1) {
            atomicExch(p.counter, 0u);
            __threadfence_system();
            st_release_sys_u64(p.arrive, p.step);
        }
    }
}

template <int M, bool SECOND_IS_SUM, int UNROLL>
__device__ __forceinline__ void lc_mac(const u64x2 (&x)[UNROLL], const int64_t* sW, int d, int j, uint64_t& a0x,
                                       uint64_t& a0y, uint64_t& a1x, uint64_t& a1y) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint64_t w = (uint64_t)sW[j + u];
        a0x += w * x[u].x;
        a0y += w * x[u].y;
        if (M == 2) {
            if (SECOND_IS_SUM) {
                a1x += x[u].x;
                a1y += x[u].y;
            } else {
                const uint64_t w1 = (uint64_t)sW[d + j + u];
                a1x += w1 * x[u].x;
                a1y += w1 * x[u].y;
            }
        }
    }
}

template <int M, bool SECOND_IS_SUM, int UNROLL, bool PUSH>
__device__ __forceinline__ void lincomb_thread(const uint64_t* __restrict__ ct, int d, int n_words, int64_t stride,
                                               int64_t g, const int64_t* sW, uint64_t bias0, uint64_t bias1,
                                               uint64_t* __restrict__ out, const PushArgs& push) {
    const int vecs = (int)(stride >> 1);
    const int64_t b = g / vecs;
    const int w0 = 2 * (int)(g - b * vecs);
    const uint64_t* p = ct + (size_t)b * d * stride + w0;
    uint64_t a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    const int d_main = d - d % UNROLL;
    if (d_main > 0) {
        u64x2 cur[UNROLL], nxt[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) cur[u] = ld_stream_u64x2(p + (size_t)u * stride);
        for (int j = 0; j < d_main; j += UNROLL) {
            if (j + UNROLL < d_main) {  // next block is in flight behind the MACs of this one
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) nxt[u] = ld_stream_u64x2(p + (size_t)(j + UNROLL + u) * stride);
            }
            lc_mac<M, SECOND_IS_SUM, UNROLL>(cur, sW, d, j, a0x, a0y, a1x, a1y);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) cur[u] = nxt[u];
        }
    }
    for (int j = d_main; j < d; ++j) {  // ragged tail (d % UNROLL rows)
        u64x2 x[1] = {ld_stream_u64x2(p + (size_t)j * stride)};
        lc_mac<M, SECOND_IS_SUM, 1>(x, sW, d, j, a0x, a0y, a1x, a1y);
    }
    // body word gets the clear bias; padding words are forced to zero
    const int nb = n_words - 1;
    if (w0 == nb) { a0x += bias0; a1x += bias1; }
    if (w0 + 1 == nb) { a0y += bias0; a1y += bias1; }
    if (w0 >= n_words) { a0x = 0; a1x = 0; }
    if (w0 + 1 >= n_words) { a0y = 0; a1y = 0; }
    if (PUSH) {
        uint32_t* o = push.board + (size_t)b * M * stride + w0;
        push_store(o, a0x, a0y);
        if (M == 2) push_store(o + stride, a1x, a1y);
    } else {
        uint64_t* o = out + (size_t)b * M * stride + w0;
        st_stream_u64x2(o, u64x2{a0x, a0y});
        if (M == 2) st_stream_u64x2(o + stride, u64x2{a1x, a1y});
    }
}

template <int M, bool SECOND_IS_SUM, int UNROLL, bool PUSH>
__global__ void __launch_bounds__(LC_THREADS)
lincomb_kernel(const uint64_t* __restrict__ ct, int d, int n_words, int64_t stride, int64_t total_vecs,
               const int64_t* __restrict__ W, uint64_t bias0, uint64_t bias1, uint64_t* __restrict__ out,
               const PushArgs push) {
    extern __shared__ int64_t sW[];  // [M][d]
    for (int i = threadIdx.x; i < M * d; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * LC_THREADS + threadIdx.x;
    if (g < total_vecs)
        lincomb_thread<M, SECOND_IS_SUM, UNROLL, PUSH>(ct, d, n_words, stride, g, sW, bias0, bias1, out, push);
    if (PUSH) push_arrive(push);  // one call site: every thread of the CTA reaches the same barrier
}

template <bool PUSH>
static cudaError_t launch_lincomb_impl(const uint64_t* d_ct, int64_t B, int d, int n, int64_t stride, const int64_t* d_W,
                                       int M, bool second_is_sum, int64_t bias0, int64_t bias1, int shift,
                                       uint64_t* d_out, const PushArgs& push, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    const int64_t total_vecs = B * (stride / 2);
    const int64_t grid64 = (total_vecs + LC_THREADS - 1) / LC_THREADS;
    if (grid64 > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)grid64;
    const size_t smem = (size_t)M * d * sizeof(int64_t);
    const uint64_t b0 = (uint64_t)bias0 << shift, b1 = (uint64_t)bias1 << shift;
    constexpr int U = LC_UNROLL;
    if (M == 1)
        lincomb_kernel<1, false, U, PUSH><<<grid, LC_THREADS, smem, s>>>(d_ct, d, n + 1, stride, total_vecs, d_W, b0, b1, d_out, push);
    else if (second_is_sum)
        lincomb_kernel<2, true, U, PUSH><<<grid, LC_THREADS, smem, s>>>(d_ct, d, n + 1, stride, total_vecs, d_W, b0, b1, d_out, push);
    else
        lincomb_kernel<2, false, U, PUSH><<<grid, LC_THREADS, smem, s>>>(d_ct, d, n + 1, stride, total_vecs, d_W, b0, b1, d_out, push);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_lincomb(const uint64_t* d_ct, int64_t B, int d, int n, int64_t stride, const int64_t* d_W, int M,
                           bool second_is_sum, int64_t bias0, int64_t bias1, int shift, uint64_t* d_out,
                           cudaStream_t s) {
    return launch_lincomb_impl<false>(d_ct, B, d, n, stride, d_W, M, second_is_sum, bias0, bias1, shift, d_out,
                                      PushArgs{}, s);
}

static PushArgs make_push(const fhe_b200_push& p) {
    PushArgs a;
    a.board = p.d_board32;
    a.arrive = p.d_arrive;
    a.step = p.step;
    a.counter = p.d_counter;
    return a;
}

cudaError_t launch_lincomb_push(const uint64_t* d_ct, int64_t B, int d, int n, int64_t stride, const int64_t* d_W, int M,
                                bool second_is_sum, int64_t bias0, int64_t bias1, int shift, const fhe_b200_push& push,
                                cudaStream_t s) {
    return launch_lincomb_impl<true>(d_ct, B, d, n, stride, d_W, M, second_is_sum, bias0, bias1, shift, nullptr,
                                     make_push(push), s);
}

// ----------------------------------------------------------------------------- seeded ciphertexts
// A fresh LWE ciphertext is (mask, body) with the mask a pure function of (seed, ciphertext id): the
// "seeded" form keeps only the 8-byte body and lets the evaluator regenerate the mask.  For this
// path it turns 1.46 MB per document into 1 KB (1 M documents = 1 GB instead of 1.46 TB, SURVEY.md
// section 7.2) and moves the dot product from the HBM roofline to the integer pipe (MASK_ROUNDS = 7 Philox rounds
// per 16 mask bytes).  Results are bit-identical to the materialised form.

// client: bodies only.  lwe_body_noise_kernel writes plaintext + error, then one warp per ciphertext adds <a, s>
// (exactly the arithmetic of warp_lwe_encrypt, nothing stored but the 8-byte body).
//
// The kernel is bound by the wide-multiplier pipe (fmaheavy; ncu profiles/r2_ncu_e2e_seeded_v1.txt), so it only
// generates what <a, s> needs.  Philox is counter-based and the key is fixed for the whole launch: a block both of
// whose key bits are 0 is never generated (a quarter of them for a uniform binary key), and a block with ONE key
// bit set needs one 64-bit word, which is one of the two products of the last round.  The CTA sorts the blocks
// into three lists in shared memory (both words / low word only / high word only); every word that is generated is
// then added unconditionally -- no per-word select.  Wrapping adds commute, so the (unordered) lists give the same
// body bit for bit.
constexpr int ENC_MAX_BLOCKS = 2048;     // mask blocks per ciphertext the shared-memory lists hold (n <= 4096)
static_assert(MASK_ROUNDS >= 3, "the block lists carry the first two rounds");

// What rounds 1-2 of mask block `blk` contribute that does not depend on the ciphertext id (counter =
// (blk, id_lo, id_hi, dom)): after round 1, z = hi(M0*blk) ^ dom ^ k1[0] and w = lo(M0*blk); round 2 multiplies that
// z by M1.  Computed once per CTA and list entry, so a (ciphertext, block) pair costs rounds 3..MASK_ROUNDS only.
struct __align__(16) EncEntry { uint32_t p_hi, p_lo, w1, blk; };

// the words of a mask block that the key selects, summed.  WORDS: 3 = both, 1 = low, 2 = high.
// (x1q_hi, x1q_lo) = M0 * x1 and y1 are the ciphertext's own share of rounds 1-2 (see the kernel).
// one IMAD.WIDE.U32, spelled in PTX so that ptxas cannot split it into IMAD.HI + IMAD (it does when the low half
// feeds an add, and both halves then occupy the multiplier)
__device__ __forceinline__ void mulwide32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}

// Accumulator of selected mask words in two limbs: lo64 += low 32 bits of a word, hi32 += its high 32 bits (mod 2^32);
// the sum is lo64 + (hi32 << 32).  Keeps the adds three-input IADD3s on the ALU pipe.
struct WordSum {
    uint64_t lo = 0;
    uint32_t hi = 0;
    __device__ __forceinline__ uint64_t value() const { return lo + ((uint64_t)hi << 32); }
};

// adds the words of a mask block that the key selects.  WORDS: 3 = both, 1 = low, 2 = high.
// (x1q_hi, x1q_lo) = M0 * x1 and y1 are the ciphertext's own share of rounds 1-2 (see the kernel).
template <int WORDS>
__device__ __forceinline__ void mask_block_sum(WordSum& acc, const PhiloxKeys& K, const EncEntry e, uint32_t y1,
                                               uint32_t x1q_hi, uint32_t x1q_lo) {
    u32x4 c{e.p_hi ^ y1 ^ K.k0[1], e.p_lo, x1q_hi ^ e.w1 ^ K.k1[1], x1q_lo};   // state after round 2
#pragma unroll
    for (int r = 2; r < MASK_ROUNDS - 1; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulwide32(0xD2511F53u, c.x, hi0, lo0);
        mulwide32(0xCD9E8D57u, c.z, hi1, lo1);
        c = u32x4{hi1 ^ c.y ^ K.k0[r], lo1, hi0 ^ c.w ^ K.k1[r], lo0};
    }
    uint32_t wl0 = 0, wh0 = 0, wl1 = 0, wh1 = 0;
    if (WORDS & 1) {   // low word = (hi1 ^ c.y ^ k0, lo1) of the last round
        uint32_t hi1;
        mulwide32(0xCD9E8D57u, c.z, hi1, wh0);
        wl0 = hi1 ^ c.y ^ K.k0[MASK_ROUNDS - 1];
    }
    if (WORDS & 2) {   // high word = (hi0 ^ c.w ^ k1, lo0)
        uint32_t hi0;
        mulwide32(0xD2511F53u, c.x, hi0, wh1);
        wl1 = hi0 ^ c.w ^ K.k1[MASK_ROUNDS - 1];
    }
    acc.lo += (uint64_t)wl0 + wl1;
    acc.hi += wh0 + wh1;
}

__global__ void __launch_bounds__(ENC_WARPS * 32)
lwe_encrypt_seeded_kernel(const uint8_t* __restrict__ key, int n, int64_t count, uint64_t enc_seed, uint64_t ct_base,
                          uint32_t purpose, uint64_t* __restrict__ bodies) {
    extern __shared__ __align__(16) uint32_t enc_smem[];     // [3][nblk] EncEntry lists, then the packed key bits
    __shared__ int cnt[4];
    const int nblk = (n + 1) / 2;
    EncEntry* list = reinterpret_cast<EncEntry*>(enc_smem);   // categories 1 (low), 2 (high), 3 (both)
    uint32_t* skey = enc_smem + 3 * (size_t)nblk * (sizeof(EncEntry) / sizeof(uint32_t));
    if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
    pack_key_bits(key, n, skey);
    __syncthreads();
    const uint32_t dom = FHE_B200_KIND_MASK | (purpose << 8);
    const PhiloxKeys K(enc_seed);
    for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
        const int w = 2 * blk;
        uint32_t b = (skey[w >> 5] >> (w & 31)) & 3u;
        if (w + 1 >= n) b &= 1u;
        if (b) {
            uint32_t hi0, lo0, p_hi, p_lo;
            mulhilo32(0xD2511F53u, (uint32_t)blk, hi0, lo0);
            mulhilo32(0xCD9E8D57u, hi0 ^ dom ^ K.k1[0], p_hi, p_lo);
            list[(b - 1) * nblk + atomicAdd(&cnt[b], 1)] = EncEntry{p_hi, p_lo, lo0, (uint32_t)blk};
        }
    }
    __syncthreads();
    const int n_lo = cnt[1], n_hi = cnt[2], n_both = cnt[3];
    const EncEntry *l_lo = list, *l_hi = list + nblk, *l_both = list + 2 * nblk;
    const int lane = threadIdx.x & 31;
    for (int64_t c = (int64_t)blockIdx.x * ENC_WARPS + (threadIdx.x >> 5); c < count; c += (int64_t)gridDim.x * ENC_WARPS) {
        const uint64_t id = ct_base + (uint64_t)c;
        const uint64_t pre = lane == 0 ? bodies[c] : 0;
        // the ciphertext's share of rounds 1-2: x1 = hi(M1*id_hi) ^ id_lo ^ k0[0], y1 = lo(M1*id_hi), then M0 * x1
        uint32_t h1, y1, q_hi, q_lo;
        mulhilo32(0xCD9E8D57u, (uint32_t)(id >> 32), h1, y1);
        mulhilo32(0xD2511F53u, h1 ^ (uint32_t)id ^ K.k0[0], q_hi, q_lo);
        WordSum acc;
#pragma unroll ENC_SEEDED_UNROLL_N
        for (int i = lane; i < n_both; i += 32) mask_block_sum<3>(acc, K, l_both[i], y1, q_hi, q_lo);
#pragma unroll ENC_SEEDED_UNROLL_N
        for (int i = lane; i < n_lo; i += 32) mask_block_sum<1>(acc, K, l_lo[i], y1, q_hi, q_lo);
#pragma unroll ENC_SEEDED_UNROLL_N
        for (int i = lane; i < n_hi; i += 32) mask_block_sum<2>(acc, K, l_hi[i], y1, q_hi, q_lo);
        const uint64_t dot = warp_sum_u64_redux(acc.value());
        if (lane == 0) bodies[c] = pre + dot;
    }
}

// any n: every block generated, words selected by AND masks (the form the lists replace)
__global__ void __launch_bounds__(ENC_WARPS * 32)
lwe_encrypt_seeded_generic_kernel(const uint8_t* __restrict__ key, int n, int64_t count, uint64_t enc_seed, uint64_t ct_base,
                                  uint32_t purpose, uint64_t* __restrict__ bodies) {
    extern __shared__ uint32_t skey[];
    pack_key_bits(key, n, skey);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t dom = FHE_B200_KIND_MASK | (purpose << 8);
    const PhiloxKeys K(enc_seed);
    const int nblk = (n + 1) / 2;
    for (int64_t c = (int64_t)blockIdx.x * ENC_WARPS + (threadIdx.x >> 5); c < count; c += (int64_t)gridDim.x * ENC_WARPS) {
        const uint64_t id = ct_base + (uint64_t)c;
        const uint64_t pre = lane == 0 ? bodies[c] : 0;
        uint64_t dot = 0;
#pragma unroll ENC_UNROLL_N
        for (int blk = lane; blk < nblk; blk += 32) {
            u32x4 r = rng_block(K, dom, id, (uint32_t)blk);
            const int w = 2 * blk;
            uint32_t bits = skey[w >> 5] >> (w & 31);
            dot += lo64(r) & (0 - (uint64_t)(bits & 1u));
            if (w + 1 < n) dot += hi64(r) & (0 - (uint64_t)((bits >> 1) & 1u));
        }
        dot = warp_sum_u64(dot);
        if (lane == 0) bodies[c] = pre + dot;
    }
}

static cudaError_t launch_encrypt_seeded_masks(const uint8_t* d_key, int n, int64_t count, uint64_t enc_seed, uint64_t ct_base,
                                               uint32_t purpose, uint64_t* d_bodies, cudaStream_t s) {
    const int nblk = (n + 1) / 2;
    if (nblk > ENC_MAX_BLOCKS) {
        size_t smem = ((size_t)(n + 31) / 32 + 1) * sizeof(uint32_t);
        lwe_encrypt_seeded_generic_kernel<<<enc_grid(lwe_encrypt_seeded_generic_kernel, smem, count), ENC_WARPS * 32, smem, s>>>(
            d_key, n, count, enc_seed, ct_base, purpose, d_bodies);
        count_launch();
        return cudaGetLastError();
    }
    size_t smem = ((size_t)(n + 31) / 32 + 1) * sizeof(uint32_t) + 3 * (size_t)nblk * sizeof(EncEntry);
    lwe_encrypt_seeded_kernel<<<enc_grid(lwe_encrypt_seeded_kernel, smem, count), ENC_WARPS * 32, smem, s>>>(
        d_key, n, count, enc_seed, ct_base, purpose, d_bodies);
    count_launch();
    return cudaGetLastError();
}

// quantize (the UniformQuantizer rule of quantize_kernel) + plaintext + error in one pass:
// the client path from float features to seeded ciphertexts needs no integer staging buffer
// With `query` (d floats) the feature is the clear product query[j] * X[i] the reference forms on the host before it
// calls the circuit (emb1 * emb2, batch_operations.py:226,273): one IEEE single-precision multiply, round to nearest --
// the same float32 value numpy produces -- so a search uploads the query and its documents instead of a product matrix
// built by the host.
__global__ void quantize_body_noise_kernel(const float* __restrict__ X, const float* __restrict__ query, int d, int64_t count,
                                           double scale, double zp, double qmin,
                                           double qmax, int shift, double sigma_abs, uint64_t noise_seed, uint64_t ct_base,
                                           uint32_t purpose, uint64_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float x = query ? __fmul_rn(query[i % d], X[i]) : X[i];
    double v = rint(__dadd_rn(__ddiv_rn((double)x, scale), zp));
    v = fmin(fmax(v, qmin), qmax);
    const uint32_t ndom = FHE_B200_KIND_NOISE | (purpose << 8);
    dst[i] = ((uint64_t)(int64_t)v << shift) + (uint64_t)gaussian_i64(noise_seed, ndom, ct_base + (uint64_t)i, 0, sigma_abs);
}

cudaError_t launch_lwe_encrypt_seeded_float(const uint8_t* d_key, int n, const float* d_X, const float* d_query, int d,
                                            int64_t count, double scale,
                                            int64_t zp, int64_t qmin, int64_t qmax, int shift, double sigma_abs,
                                            uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base, uint32_t purpose,
                                            uint64_t* d_bodies, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    quantize_body_noise_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_X, d_query, d, count, scale, (double)zp, (double)qmin,
                                                                             (double)qmax, shift, sigma_abs, noise_seed, ct_base,
                                                                             purpose, d_bodies);
    count_launch();
    return launch_encrypt_seeded_masks(d_key, n, count, enc_seed, ct_base, purpose, d_bodies, s);
}

cudaError_t launch_lwe_encrypt_seeded(const uint8_t* d_key, int n, const int64_t* d_msgs, int64_t count, int shift,
                                      double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                                      uint32_t purpose, uint64_t* d_bodies, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    lwe_body_noise_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_msgs, count, shift, sigma_abs, noise_seed, ct_base,
                                                                        purpose, d_bodies, 1);
    count_launch();
    return launch_encrypt_seeded_masks(d_key, n, count, enc_seed, ct_base, purpose, d_bodies, s);
}

// word w of the materialised ciphertext `id` whose body is `body`
__device__ __forceinline__ void seeded_pair(const PhiloxKeys& K, uint32_t dom, uint64_t id, int w0, int n, uint64_t body,
                                            uint64_t& x, uint64_t& y) {
    u32x4 r = rng_block(K, dom, id, (uint32_t)(w0 >> 1));
    x = w0 < n ? lo64(r) : (w0 == n ? body : 0);
    y = w0 + 1 < n ? hi64(r) : (w0 + 1 == n ? body : 0);
}

// materialise [count][stride] from bodies (interop / tests)
__global__ void lwe_expand_seeded_kernel(const uint64_t* __restrict__ bodies, int64_t count, int n, int64_t stride,
                                         uint64_t enc_seed, uint64_t ct_base, uint32_t purpose,
                                         uint64_t* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int vecs = (int)(stride >> 1);
    if (g >= count * vecs) return;
    const int64_t c = g / vecs;
    const int w0 = 2 * (int)(g - c * vecs);
    uint64_t x, y;
    seeded_pair(PhiloxKeys(enc_seed), FHE_B200_KIND_MASK | (purpose << 8), ct_base + (uint64_t)c, w0, n, bodies[c], x, y);
    st_stream_u64x2(out + c * stride + w0, u64x2{x, y});
}

cudaError_t launch_lwe_expand_seeded(const uint64_t* d_bodies, int64_t count, int n, int64_t stride, uint64_t enc_seed,
                                     uint64_t ct_base, uint32_t purpose, uint64_t* d_out, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t tot = count * (stride / 2);
    lwe_expand_seeded_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_bodies, count, n, stride, enc_seed, ct_base,
                                                                         purpose, d_out);
    count_launch();
    return cudaGetLastError();
}

// server: out[b][m][:] = sum_j W[m][j] * expand(seed, ct_base + b*d + j)[:], masks regenerated on the fly.
// Same thread -> (document, column pair) mapping as lincomb_kernel; no global loads except the d bodies.
//
// The kernel is bound by the wide-multiplier pipe (ncu, profiles/r2_ncu_e2e_seeded_v1.txt: math-pipe throttle is the
// top stall), and besides Philox the u64 x i64 MACs were on that pipe too (one IMAD.WIDE + two IMAD + an add each).
// Quantized weights span far less than 2^32, so the CTA shifts them to unsigned 32-bit values w' = w - min(w) and
// accumulates  lo64 += x_lo * w'  (one IMAD.WIDE with 64-bit addend)  and  hi32 += x_hi * w'  (one IMAD);
// sum_j w_j x_j = lo64 + (hi32 << 32) + min(w) * sum_j x_j  (mod 2^64, exact), and sum_j x_j is the second output
// the two-output circuit needs anyway.  Rows whose weights span 2^32 or more take the plain 64-bit MAC.
struct SeededRow {       // per weight row, CTA-uniform
    int64_t wmin;
    int narrow;          // max - min < 2^32
};

// Rounds 1-2 of a mask block split into what depends on the block only and what depends on the ciphertext only
// (see EncEntry): the CTA computes the ciphertext share once per (document, j) it touches -- every thread of a document
// needs the same one -- together with the shifted weight, so that the inner loop is one broadcast LDS.128 plus rounds
// 3..MASK_ROUNDS: 10 wide multiplies per block instead of 11 (+1 that ptxas split), on a kernel whose time is set by
// the number of IMAD.WIDE it issues (~7 cycles each per scheduler, measured: profiles/r2_seeded_kernel_times.txt).
struct __align__(16) CtShare { uint32_t y1, q_hi, q_lo, w; };
constexpr int LCS_MAX_TABLE = 2048;      // (document, j) entries a CTA may hold: 32 KB

template <int M, bool SECOND_IS_SUM, bool PUSH>
__global__ void __launch_bounds__(LC_THREADS, PUSH ? 4 : LCS_MIN_CTAS)
lincomb_seeded_kernel(const uint64_t* __restrict__ bodies, int d, int n, int64_t stride, int64_t total_vecs,
                      uint64_t enc_seed, uint64_t ct_base, uint32_t purpose, const int64_t* __restrict__ W,
                      uint64_t bias0, uint64_t bias1, uint64_t* __restrict__ out, const PushArgs push, int table_docs) {
    extern __shared__ __align__(16) int64_t sW[];    // [M*d] i64 weights, [table_docs*d] CtShare, [d] u32 (second row)
    __shared__ SeededRow rows[2];
    CtShare* tab = reinterpret_cast<CtShare*>(sW + ((M * d + 1) & ~1));     // 16-byte aligned
    uint32_t* sW32b = reinterpret_cast<uint32_t*>(tab + (size_t)table_docs * d);
    for (int i = threadIdx.x; i < M * d; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
    if (threadIdx.x < 32 * M) {                      // warp m: range of row m
        const int m = threadIdx.x >> 5, lane = threadIdx.x & 31;
        int64_t lo = INT64_MAX, hi = INT64_MIN;
        for (int j = lane; j < d; j += 32) { lo = min(lo, sW[m * d + j]); hi = max(hi, sW[m * d + j]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { rows[m].wmin = lo; rows[m].narrow = ((uint64_t)hi - (uint64_t)lo) < (1ULL << 32); }
    }
    __syncthreads();
    constexpr int MW = (M == 2 && !SECOND_IS_SUM) ? 2 : 1;    // rows that carry real weights
    bool narrow = rows[0].narrow && table_docs > 0;
    if (MW == 2) narrow = narrow && rows[1].narrow;
    const int vecs = (int)(stride >> 1);
    const uint32_t dom = FHE_B200_KIND_MASK | (purpose << 8);
    const PhiloxKeys K(enc_seed);
    const int64_t g0 = (int64_t)blockIdx.x * LC_THREADS;
    const int64_t b_first = g0 / vecs;
    if (narrow) {
        const int64_t g_last = min(g0 + LC_THREADS, total_vecs) - 1;
        const int nd = (int)(g_last / vecs - b_first) + 1;            // documents this CTA touches (<= table_docs)
        for (int i = threadIdx.x; i < nd * d; i += blockDim.x) {
            const int j = i % d;
            const uint64_t id = ct_base + (uint64_t)(b_first * d + i);
            uint32_t h1, y1, q_hi, q_lo;
            mulwide32(0xCD9E8D57u, (uint32_t)(id >> 32), h1, y1);
            mulwide32(0xD2511F53u, h1 ^ (uint32_t)id ^ K.k0[0], q_hi, q_lo);
            tab[i] = CtShare{y1, q_hi, q_lo, (uint32_t)((uint64_t)sW[j] - (uint64_t)rows[0].wmin)};
        }
        if (MW == 2)
            for (int j = threadIdx.x; j < d; j += blockDim.x) sW32b[j] = (uint32_t)((uint64_t)sW[d + j] - (uint64_t)rows[1].wmin);
        __syncthreads();
    }
    const int64_t g = g0 + threadIdx.x;
    // Does every lane of this warp that will run the table loop hold two MASK words?  Voted here, where the warp is
    // still converged (lanes past the end of the batch or in the padding behind the body do not enter the loop).
    bool plain;
    {
        const int64_t bq = g / vecs;
        const int wq = 2 * (int)(g - bq * vecs);
        plain = __all_sync(0xffffffffu, g >= total_vecs || wq > n || wq + 1 < n);
    }
    if (g < total_vecs) {
        const int64_t b = g / vecs;
        const int w0 = 2 * (int)(g - b * vecs);
        const uint64_t id0 = ct_base + (uint64_t)b * d;
        const bool has_body = (w0 == n) || (w0 + 1 == n);
        uint64_t a0x = 0, a0y = 0, a1x = 0, a1y = 0;
        if (w0 <= n && narrow) {
            // the block's share of rounds 1-2
            uint32_t hi0, w1, p_hi, p_lo;
            mulwide32(0xD2511F53u, (uint32_t)(w0 >> 1), hi0, w1);
            mulwide32(0xCD9E8D57u, hi0 ^ dom ^ K.k1[0], p_hi, p_lo);
            const CtShare* row = tab + (size_t)(b - b_first) * d;
            uint64_t l0x = 0, l0y = 0, l1x = 0, l1y = 0, sx = 0, sy = 0;   // low-limb accumulators, plain sums
            uint32_t h0x = 0, h0y = 0, h1x = 0, h1y = 0;                   // high-limb accumulators (mod 2^32)
            // The loop below generates MASK words only.  Which of a thread's two words are mask words is a per-thread
            // constant (all of them except in the one or two threads per document that hold the body / the padding), so
            // a warp without such a thread runs the loop with no selects at all and the others AND their words with
            // constant masks; the body's own term sum_j w_j body_j is added after the loop by the thread that holds it.
            // (Selecting mask / body / zero per word inside the loop cost ~10 predicated instructions per iteration,
            // several of them IMAD.MOVs on the very pipe that bounds the kernel.)
            const uint64_t mx = w0 < n ? ~0ULL : 0ULL, my = w0 + 1 < n ? ~0ULL : 0ULL;
            auto mac_loop = [&](auto masked) {
#pragma unroll LCS_UNROLL_N
                for (int j = 0; j < d; ++j) {
                    const CtShare e = row[j];
                    u32x4 c{p_hi ^ e.y1 ^ K.k0[1], p_lo, e.q_hi ^ w1 ^ K.k1[1], e.q_lo};   // state after round 2
#pragma unroll
                    for (int r = 2; r < MASK_ROUNDS; ++r) {
                        uint32_t m0h, m0l, m1h, m1l;
                        mulwide32(0xD2511F53u, c.x, m0h, m0l);
                        mulwide32(0xCD9E8D57u, c.z, m1h, m1l);
                        c = u32x4{m1h ^ c.y ^ K.k0[r], m1l, m0h ^ c.w ^ K.k1[r], m0l};
                    }
                    uint64_t x = lo64(c), y = hi64(c);
                    if (decltype(masked)::value) { x &= mx; y &= my; }
                    const uint32_t w = e.w;
                    l0x += (uint64_t)(uint32_t)x * w;  h0x += (uint32_t)(x >> 32) * w;
                    l0y += (uint64_t)(uint32_t)y * w;  h0y += (uint32_t)(y >> 32) * w;
                    sx += x;
                    sy += y;
                    if (MW == 2) {
                        const uint32_t wb = sW32b[j];
                        l1x += (uint64_t)(uint32_t)x * wb;  h1x += (uint32_t)(x >> 32) * wb;
                        l1y += (uint64_t)(uint32_t)y * wb;  h1y += (uint32_t)(y >> 32) * wb;
                    }
                }
            };
            if (plain) mac_loop(std::false_type{});
            else mac_loop(std::true_type{});
            if (has_body) {          // this thread's word at index n is the body: add sum_j w_j body_j to that word's sums
                uint64_t lb = 0, l1b = 0, sb = 0;
                uint32_t hb = 0, h1b = 0;
                for (int j = 0; j < d; ++j) {
                    const uint64_t body = bodies[b * d + j];
                    const uint32_t w = row[j].w;
                    lb += (uint64_t)(uint32_t)body * w;
                    hb += (uint32_t)(body >> 32) * w;
                    sb += body;
                    if (MW == 2) {
                        const uint32_t wb = sW32b[j];
                        l1b += (uint64_t)(uint32_t)body * wb;
                        h1b += (uint32_t)(body >> 32) * wb;
                    }
                }
                if (w0 == n) { l0x += lb; h0x += hb; sx += sb; l1x += l1b; h1x += h1b; }
                else { l0y += lb; h0y += hb; sy += sb; l1y += l1b; h1y += h1b; }
            }
            const uint64_t m0 = (uint64_t)rows[0].wmin;
            a0x = l0x + ((uint64_t)h0x << 32) + m0 * sx;
            a0y = l0y + ((uint64_t)h0y << 32) + m0 * sy;
            if (M == 2 && SECOND_IS_SUM) { a1x = sx; a1y = sy; }
            if (MW == 2) {
                const uint64_t m1 = (uint64_t)rows[1].wmin;
                a1x = l1x + ((uint64_t)h1x << 32) + m1 * sx;
                a1y = l1y + ((uint64_t)h1y << 32) + m1 * sy;
            }
        } else if (w0 <= n) {
#pragma unroll LCS_UNROLL_N
            for (int j = 0; j < d; ++j) {
                uint64_t x, y;
                seeded_pair(K, dom, id0 + j, w0, n, has_body ? bodies[b * d + j] : 0, x, y);
                const uint64_t w = (uint64_t)sW[j];
                a0x += w * x;
                a0y += w * y;
                if (M == 2) {
                    const uint64_t w1 = SECOND_IS_SUM ? 1ULL : (uint64_t)sW[d + j];
                    a1x += w1 * x;
                    a1y += w1 * y;
                }
            }
        }
        if (w0 == n) { a0x += bias0; a1x += bias1; }
        if (w0 + 1 == n) { a0y += bias0; a1y += bias1; }
        if (PUSH) {
            uint32_t* o = push.board + (size_t)b * M * stride + w0;
            push_store(o, a0x, a0y);
            if (M == 2) push_store(o + stride, a1x, a1y);
        } else {
            uint64_t* o = out + (size_t)b * M * stride + w0;
            st_stream_u64x2(o, u64x2{a0x, a0y});
            if (M == 2) st_stream_u64x2(o + stride, u64x2{a1x, a1y});
        }
    }
    if (PUSH) push_arrive(push);
}

template <bool PUSH>
static cudaError_t launch_lincomb_seeded_impl(const uint64_t* d_bodies, int64_t B, int d, int n, int64_t stride,
                                              uint64_t enc_seed, uint64_t ct_base, uint32_t purpose, const int64_t* d_W,
                                              int M, bool second_is_sum, int64_t bias0, int64_t bias1, int shift,
                                              uint64_t* d_out, const PushArgs& push, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    const int64_t total_vecs = B * (stride / 2);
    const int64_t grid64 = (total_vecs + LC_THREADS - 1) / LC_THREADS;
    if (grid64 > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)grid64;
    // documents a 256-thread CTA can touch: its first thread may sit anywhere inside a document
    const int64_t vecs = stride / 2;
    int table_docs = (int)((LC_THREADS + vecs - 2) / vecs) + 1;
    if ((int64_t)table_docs * d > LCS_MAX_TABLE) table_docs = 0;       // tiny ciphertexts x huge d: plain path
    const size_t smem = (size_t)((M * d + 1) & ~1) * sizeof(int64_t) + (size_t)table_docs * d * sizeof(CtShare) + (size_t)d * sizeof(uint32_t);
    const uint64_t b0 = (uint64_t)bias0 << shift, b1 = (uint64_t)bias1 << shift;
    if (M == 1)
        lincomb_seeded_kernel<1, false, PUSH><<<grid, LC_THREADS, smem, s>>>(d_bodies, d, n, stride, total_vecs, enc_seed, ct_base, purpose, d_W, b0, b1, d_out, push, table_docs);
    else if (second_is_sum)
        lincomb_seeded_kernel<2, true, PUSH><<<grid, LC_THREADS, smem, s>>>(d_bodies, d, n, stride, total_vecs, enc_seed, ct_base, purpose, d_W, b0, b1, d_out, push, table_docs);
    else
        lincomb_seeded_kernel<2, false, PUSH><<<grid, LC_THREADS, smem, s>>>(d_bodies, d, n, stride, total_vecs, enc_seed, ct_base, purpose, d_W, b0, b1, d_out, push, table_docs);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_lincomb_seeded(const uint64_t* d_bodies, int64_t B, int d, int n, int64_t stride, uint64_t enc_seed,
                                  uint64_t ct_base, uint32_t purpose, const int64_t* d_W, int M, bool second_is_sum,
                                  int64_t bias0, int64_t bias1, int shift, uint64_t* d_out, cudaStream_t s) {
    return launch_lincomb_seeded_impl<false>(d_bodies, B, d, n, stride, enc_seed, ct_base, purpose, d_W, M, second_is_sum,
                                             bias0, bias1, shift, d_out, PushArgs{}, s);
}

cudaError_t launch_lincomb_seeded_push(const uint64_t* d_bodies, int64_t B, int d, int n, int64_t stride,
                                       uint64_t enc_seed, uint64_t ct_base, uint32_t purpose, const int64_t* d_W, int M,
                                       bool second_is_sum, int64_t bias0, int64_t bias1, int shift,
                                       const fhe_b200_push& push, cudaStream_t s) {
    return launch_lincomb_seeded_impl<true>(d_bodies, B, d, n, stride, enc_seed, ct_base, purpose, d_W, M, second_is_sum,
                                            bias0, bias1, shift, nullptr, make_push(push), s);
}

// ----------------------------------------------------------------------------- score board flags
// client: the stream continues once every listed arrival flag has reached `value`
__global__ void peer_wait_kernel(const uint64_t* __restrict__ flags, int count, uint64_t value, uint64_t timeout_ns,
                                 uint32_t* __restrict__ status) {
    for (int i = threadIdx.x; i < count; i += blockDim.x)
        if (!spin_until_ge(flags + i, value, timeout_ns)) atomicExch(status, 1u);
}
// stream-ordered release store to (peer) flags: everything before it on the stream is visible first
__global__ void peer_signal_kernel(uint64_t* const* __restrict__ ptrs, int count, uint64_t value) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        __threadfence_system();
        st_release_sys_u64(ptrs[i], value);
    }
}

cudaError_t launch_peer_wait(const uint64_t* d_flags, int count, uint64_t value, uint32_t timeout_ms,
                             uint32_t* d_status, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    peer_wait_kernel<<<1, 32, 0, s>>>(d_flags, count, value, (uint64_t)timeout_ms * 1000000ULL, d_status);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_peer_signal(uint64_t* const* d_flag_ptrs, int count, uint64_t value, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    peer_signal_kernel<<<1, 32, 0, s>>>(d_flag_ptrs, count, value);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- 32-bit wire form
// Modulus switch 2^64 -> 2^32 of finished ciphertexts (scores on their way to the client): halves
// the bytes that cross NVLink.  Added noise: sum_i s_i e_i with e_i uniform in +-2^-33 of the torus,
// std ~ sqrt(n/24) * 2^-32 (2^-29 at n = 1423, against a decoding margin of 2^-23).
__global__ void lwe_modswitch32_kernel(const uint64_t* __restrict__ in, int64_t words, uint32_t* __restrict__ out) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < words) {
        u64x2 v = ld_stream_u64x2(in + i);
        uint2 o;
        o.x = (uint32_t)((v.x + 0x80000000ULL) >> 32);
        o.y = (uint32_t)((v.y + 0x80000000ULL) >> 32);
        *reinterpret_cast<uint2*>(out + i) = o;
    } else if (i < words) {
        out[i] = (uint32_t)((in[i] + 0x80000000ULL) >> 32);
    }
}

cudaError_t launch_lwe_modswitch32(const uint64_t* d_in, int64_t words, uint32_t* d_out, cudaStream_t s) {
    if (words <= 0) return cudaSuccess;
    int64_t vecs = (words + 1) / 2;
    lwe_modswitch32_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, s>>>(d_in, words, d_out);
    count_launch();
    return cudaGetLastError();
}

// decrypt of 32-bit ciphertexts: m = (b - <a,s> + Delta32/2) >> shift32, arithmetic on 32 bits
__global__ void __launch_bounds__(DEC_WARPS * 32)
lwe_decrypt32_kernel(const uint8_t* __restrict__ key, int n, int64_t stride, const uint32_t* __restrict__ cts,
                     int64_t count, int shift32, int64_t* __restrict__ out) {
    extern __shared__ uint32_t skey[];
    pack_key_bits(key, n, skey);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (c >= count) return;
    const uint32_t* ct = cts + c * stride;
    uint32_t dot = 0;
    for (int w = lane; w < n; w += 32) dot += ct[w] & (0u - ((skey[w >> 5] >> (w & 31)) & 1u));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        uint32_t v = ct[n] - dot + (shift32 > 0 ? (1u << (shift32 - 1)) : 0u);
        out[c] = (int64_t)((int32_t)v >> shift32);
    }
}

cudaError_t launch_lwe_decrypt32(const uint8_t* d_key, int n, int64_t stride, const uint32_t* d_ct, int64_t count,
                                 int shift32, int64_t* d_out, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    size_t smem = ((size_t)(n + 31) / 32 + 1) * sizeof(uint32_t);
    lwe_decrypt32_kernel<<<(unsigned)((count + DEC_WARPS - 1) / DEC_WARPS), DEC_WARPS * 32, smem, s>>>(
        d_key, n, stride, d_ct, count, shift32, d_out);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- accumulation
__global__ void accumulate_kernel(uint64_t* __restrict__ acc, const uint64_t* __restrict__ x, int64_t words) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < words) {
        u64x2 a = *reinterpret_cast<const u64x2*>(acc + i);
        u64x2 b = ld_stream_u64x2(x + i);
        a.x += b.x;
        a.y += b.y;
        *reinterpret_cast<u64x2*>(acc + i) = a;
    } else if (i < words) {
        acc[i] += x[i];
    }
}

cudaError_t launch_accumulate(uint64_t* d_acc, const uint64_t* d_x, int64_t words, cudaStream_t s) {
    if (words <= 0) return cudaSuccess;
    int64_t vecs = (words + 1) / 2;
    accumulate_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, s>>>(d_acc, d_x, words);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- encrypted x encrypted glue
// xy = floor((x+y)^2/4) - floor((x-y)^2/4): the two PBS inputs per dimension are the sum and the
// difference of the query's and the document's ciphertexts, shifted by a plaintext offset into the
// unsigned half of the message space.  out[b][j][0] = q[j] + y[b][j] + off, out[b][j][1] = q[j] - y[b][j] + off.
constexpr int PAIR_THREADS = 256;

__global__ void __launch_bounds__(PAIR_THREADS)
lwe_pair_addsub_kernel(const uint64_t* __restrict__ q, const uint64_t* __restrict__ y, int d, int words,
                       int64_t in_stride, uint64_t offset, uint64_t* __restrict__ out) {
    const int64_t row = blockIdx.x;  // b*d + j
    const int j = (int)(row % d);
    const uint64_t* qr = q + (size_t)j * in_stride;
    const uint64_t* yr = y + (size_t)row * in_stride;
    uint64_t* o0 = out + (size_t)row * 2 * words;
    uint64_t* o1 = o0 + words;
    for (int w = threadIdx.x; w < words; w += PAIR_THREADS) {
        const uint64_t a = qr[w], b = yr[w];
        const uint64_t off = w == words - 1 ? offset : 0;
        o0[w] = a + b + off;
        o1[w] = a - b + off;
    }
}

cudaError_t launch_lwe_pair_addsub(const uint64_t* d_q, const uint64_t* d_y, int64_t B, int d, int words,
                                   int64_t in_stride, uint64_t offset, uint64_t* d_out, cudaStream_t s) {
    if (B <= 0 || d <= 0) return cudaSuccess;
    lwe_pair_addsub_kernel<<<(unsigned)(B * d), PAIR_THREADS, 0, s>>>(d_q, d_y, d, words, in_stride, offset, d_out);
    count_launch();
    return cudaGetLastError();
}

// out[b][w] = sum_j (in[b][j][0][w] - in[b][j][1][w]): the encrypted score from the 2d bootstrapped squares
__global__ void __launch_bounds__(PAIR_THREADS)
lwe_pair_diff_sum_kernel(const uint64_t* __restrict__ in, int d, int words, int64_t out_stride,
                         uint64_t* __restrict__ out) {
    const int w = blockIdx.x * PAIR_THREADS + threadIdx.x;
    if (w >= words) {
        if (w < out_stride) out[(size_t)blockIdx.y * out_stride + w] = 0;  // row padding
        return;
    }
    const uint64_t* base = in + (size_t)blockIdx.y * d * 2 * words + w;
    uint64_t acc = 0;
    int j = 0;
    for (; j + 4 <= d; j += 4) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(base + (size_t)(2 * j + u) * words);
#pragma unroll
        for (int u = 0; u < 8; u += 2) acc += v[u] - v[u + 1];
    }
    for (; j < d; ++j) acc += base[(size_t)(2 * j) * words] - base[(size_t)(2 * j + 1) * words];
    out[(size_t)blockIdx.y * out_stride + w] = acc;
}

cudaError_t launch_lwe_pair_diff_sum(const uint64_t* d_in, int64_t B, int d, int words, int64_t out_stride,
                                     uint64_t* d_out, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {  // gridDim.y limit
        const unsigned nb = (unsigned)((B - b0) < 65535 ? (B - b0) : 65535);
        dim3 grid((unsigned)((out_stride + PAIR_THREADS - 1) / PAIR_THREADS), nb);
        lwe_pair_diff_sum_kernel<<<grid, PAIR_THREADS, 0, s>>>(d_in + (size_t)b0 * d * 2 * words, d, words, out_stride,
                                                              d_out + (size_t)b0 * out_stride);
        count_launch();
    }
    return cudaGetLastError();
}

// One bootstrap per dimension when each party also supplies an encryption of its own squared norm:
// 2*sum_j x_j*y_j = sum_j (x_j+y_j)^2 - sum_j x_j^2 - sum_j y_j^2.  pair_add builds the d sums,
// square_sum folds the d bootstrapped squares and subtracts the two norm ciphertexts.
__global__ void __launch_bounds__(PAIR_THREADS)
lwe_pair_add_kernel(const uint64_t* __restrict__ q, const uint64_t* __restrict__ y, int d, int words,
                    int64_t in_stride, uint64_t offset, uint64_t* __restrict__ out) {
    const int64_t row = blockIdx.x;  // b*d + j
    const int j = (int)(row % d);
    const uint64_t* qr = q + (size_t)j * in_stride;
    const uint64_t* yr = y + (size_t)row * in_stride;
    uint64_t* o = out + (size_t)row * words;
    for (int w = threadIdx.x; w < words; w += PAIR_THREADS) o[w] = qr[w] + yr[w] + (w == words - 1 ? offset : 0);
}

cudaError_t launch_lwe_pair_add(const uint64_t* d_q, const uint64_t* d_y, int64_t B, int d, int words,
                                int64_t in_stride, uint64_t offset, uint64_t* d_out, cudaStream_t s) {
    if (B <= 0 || d <= 0) return cudaSuccess;
    lwe_pair_add_kernel<<<(unsigned)(B * d), PAIR_THREADS, 0, s>>>(d_q, d_y, d, words, in_stride, offset, d_out);
    count_launch();
    return cudaGetLastError();
}

// out[b][w] = sum_j sq[b][j][w] - norm_q[w] - norm_y[b][w]
__global__ void __launch_bounds__(PAIR_THREADS)
lwe_square_sum_kernel(const uint64_t* __restrict__ sq, int d, int words, const uint64_t* __restrict__ norm_q,
                      const uint64_t* __restrict__ norm_y, int64_t norm_stride, int64_t out_stride,
                      uint64_t* __restrict__ out) {
    const int w = blockIdx.x * PAIR_THREADS + threadIdx.x;
    if (w >= words) {
        if (w < out_stride) out[(size_t)blockIdx.y * out_stride + w] = 0;  // row padding
        return;
    }
    const uint64_t* base = sq + (size_t)blockIdx.y * d * words + w;
    uint64_t acc = 0 - norm_q[w] - norm_y[(size_t)blockIdx.y * norm_stride + w];
    int j = 0;
    for (; j + 8 <= d; j += 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(base + (size_t)(j + u) * words);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; j < d; ++j) acc += base[(size_t)j * words];
    out[(size_t)blockIdx.y * out_stride + w] = acc;
}

cudaError_t launch_lwe_square_sum(const uint64_t* d_sq, int64_t B, int d, int words, const uint64_t* d_norm_q,
                                  const uint64_t* d_norm_y, int64_t norm_stride, int64_t out_stride, uint64_t* d_out,
                                  cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {  // gridDim.y limit
        const unsigned nb = (unsigned)((B - b0) < 65535 ? (B - b0) : 65535);
        dim3 grid((unsigned)((out_stride + PAIR_THREADS - 1) / PAIR_THREADS), nb);
        lwe_square_sum_kernel<<<grid, PAIR_THREADS, 0, s>>>(d_sq + (size_t)b0 * d * words, d, words, d_norm_q,
                                                           d_norm_y + (size_t)b0 * norm_stride, norm_stride, out_stride,
                                                           d_out + (size_t)b0 * out_stride);
        count_launch();
    }
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- packed GLWE results
// A GLWE ciphertext (A, B) under S carries N coefficients; the packed inner-product path needs `count`
// of them (first, first+step, ...).  Client: phase_q = B[idx] - sum_i S_i * a_i(idx), with the sample-
// extraction mask a_i(idx) = A[idx-i] (i <= idx), -A[N+idx-i] (i > idx), decoded as round(phase / 2^shift).
constexpr int GDEC_THREADS = 256;

__global__ void __launch_bounds__(GDEC_THREADS)
glwe_decrypt_coeffs_kernel(const uint8_t* __restrict__ S_big, const uint64_t* __restrict__ glwe, int N, int first,
                           int step, int count, int shift, int64_t* __restrict__ out) {
    extern __shared__ uint64_t gsm[];  // A[N], then the key bits
    uint64_t* A = gsm;
    uint32_t* Sb = reinterpret_cast<uint32_t*>(gsm + N);
    const uint64_t* g = glwe + (size_t)blockIdx.x * 2 * N;
    for (int x = threadIdx.x; x < N; x += GDEC_THREADS) A[x] = __ldcs(g + x);
    for (int w = threadIdx.x; w < N / 32; w += GDEC_THREADS) {
        const uint4* sp = reinterpret_cast<const uint4*>(S_big + w * 32);
        uint32_t bits = 0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {   // 32 key bytes -> 32 bits
            const uint4 q4 = sp[v];
            const uint32_t ws[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) bits |= ((ws[c] >> (8 * bb)) & 1u) << (v * 16 + c * 4 + bb);
        }
        Sb[w] = bits;
    }
    __syncthreads();
    // one warp per requested coefficient: no block-wide reduction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < count; q += GDEC_THREADS / 32) {
        const int idx = first + q * step;
        uint64_t acc = 0;
#pragma unroll 4
        for (int w = 0; w < N / 32; ++w) {
            const int i = w * 32 + lane;
            const uint64_t sel = 0 - (uint64_t)((Sb[w] >> lane) & 1u);
            const int m = idx - i;
            const uint64_t a = m >= 0 ? A[m] : (uint64_t)0 - A[m + N];
            acc += a & sel;
        }
        acc = warp_sum_u64(acc);
        if (lane == 0) {
            const uint64_t phase = g[(size_t)N + idx] - acc;
            out[(size_t)blockIdx.x * count + q] = (int64_t)((phase + (shift ? (1ull << (shift - 1)) : 0)) >> shift);
        }
    }
}

cudaError_t launch_glwe_decrypt_coeffs(const uint8_t* d_S_big, const uint64_t* d_glwe, int64_t G, int N, int first,
                                       int step, int count, int shift, int64_t* d_out, cudaStream_t s) {
    if (G <= 0 || count <= 0) return cudaSuccess;
    const size_t smem = (size_t)N * 8 + ((size_t)N / 64 + 1) * 8;
    glwe_decrypt_coeffs_kernel<<<(unsigned)G, GDEC_THREADS, smem, s>>>(d_S_big, d_glwe, N, first, step, count, shift, d_out);
    count_launch();
    return cudaGetLastError();
}

// Server-side sample extraction (when a score has to continue as an LWE ciphertext, e.g. into the
// encrypted threshold): out[(g*count + q)][0..N] = extract(glwe[g], first + q*step), rows out_stride apart.
__global__ void __launch_bounds__(PAIR_THREADS)
glwe_sample_extract_kernel(const uint64_t* __restrict__ glwe, int N, int first, int step, int count,
                           int64_t out_stride, uint64_t* __restrict__ out) {
    const int64_t row = blockIdx.x;  // g*count + q
    const int64_t gi = row / count;
    const int idx = first + (int)(row - gi * count) * step;
    const uint64_t* A = glwe + (size_t)gi * 2 * N;
    uint64_t* o = out + (size_t)row * out_stride;
    for (int i = threadIdx.x; i < out_stride; i += PAIR_THREADS) {
        uint64_t v = 0;
        if (i < N) {
            const int m = idx - i;
            v = m >= 0 ? A[m] : (uint64_t)0 - A[m + N];
        } else if (i == N) {
            v = A[(size_t)N + idx];
        }
        o[i] = v;
    }
}

cudaError_t launch_glwe_sample_extract(const uint64_t* d_glwe, int64_t G, int N, int first, int step, int count,
                                       int64_t out_stride, uint64_t* d_out, cudaStream_t s) {
    if (G <= 0 || count <= 0) return cudaSuccess;
    glwe_sample_extract_kernel<<<(unsigned)(G * count), PAIR_THREADS, 0, s>>>(d_glwe, N, first, step, count, out_stride,
                                                                             d_out);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- bit-extraction glue
// Exact encrypted threshold (sign of a wide message by LSB-first bit extraction): per step the
// ciphertext is scaled by a power of two (the wanted bit moves to the top of the torus, higher bits
// wrap away) and a plaintext offset is added to the body; the bootstrapped bit is then subtracted.
__global__ void __launch_bounds__(PAIR_THREADS)
lwe_shl_add_kernel(const uint64_t* __restrict__ in, int64_t in_stride, int64_t count, int words, int shift,
                   uint64_t offset, uint64_t* __restrict__ out, int64_t out_stride) {
    const int64_t idx = (int64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
    if (idx >= count * out_stride) return;
    const int64_t row = idx / out_stride;
    const int w = (int)(idx - row * out_stride);
    out[idx] = w < words ? (in[row * in_stride + w] << shift) + (w == words - 1 ? offset : 0) : 0;
}

cudaError_t launch_lwe_shl_add(const uint64_t* d_in, int64_t in_stride, int64_t count, int words, int shift,
                               uint64_t offset, uint64_t* d_out, int64_t out_stride, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t total = count * out_stride;
    lwe_shl_add_kernel<<<(unsigned)((total + PAIR_THREADS - 1) / PAIR_THREADS), PAIR_THREADS, 0, s>>>(
        d_in, in_stride, count, words, shift, offset, d_out, out_stride);
    count_launch();
    return cudaGetLastError();
}

// acc[row][w] -= x[row][w] (+ plain on the body), acc rows acc_stride apart, x rows `words` apart
__global__ void __launch_bounds__(PAIR_THREADS)
lwe_sub_plain_kernel(uint64_t* __restrict__ acc, int64_t acc_stride, const uint64_t* __restrict__ x, int64_t count,
                     int words, uint64_t plain) {
    const int64_t idx = (int64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
    if (idx >= count * words) return;
    const int64_t row = idx / words;
    const int w = (int)(idx - row * words);
    acc[row * acc_stride + w] -= x[idx] + (w == words - 1 ? plain : 0);
}

cudaError_t launch_lwe_sub_plain(uint64_t* d_acc, int64_t acc_stride, const uint64_t* d_x, int64_t count, int words,
                                 uint64_t plain, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t total = count * words;
    lwe_sub_plain_kernel<<<(unsigned)((total + PAIR_THREADS - 1) / PAIR_THREADS), PAIR_THREADS, 0, s>>>(
        d_acc, acc_stride, d_x, count, words, plain);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- quantize / finalize
// q = clip(rint(x / scale + zp), qmin, qmax) in float64, the UniformQuantizer rule
// (SURVEY.md Appendix A.1); IEEE div/add/rint => identical to numpy.
__global__ void quantize_kernel(const float* __restrict__ X, int64_t count, double scale, double zp, double qmin,
                                double qmax, int64_t* __restrict__ q) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double v = rint(__dadd_rn(__ddiv_rn((double)X[i], scale), zp));
    v = fmin(fmax(v, qmin), qmax);
    q[i] = (int64_t)v;
}

cudaError_t launch_quantize(const float* d_X, int64_t count, double scale, int64_t zp, int64_t qmin, int64_t qmax,
                            int64_t* d_q, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    quantize_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_X, count, scale, (double)zp, (double)qmin,
                                                                   (double)qmax, d_q);
    count_launch();
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- client: fused decrypt
// Decrypt + decode + dequantize of the M score ciphertexts of one document in one CTA (the client half
// of predict_encrypted, /root/reference/fhe_similarity.py:151-154).  8 warps share the M rows (4 warps
// per row at M = 2), every lane issues all of its 16-byte loads before the first use, and the key is
// read as pre-packed bit words, so a row costs one memory round trip instead of a 22-iteration
// dependent loop in a single warp (16 us -> ~5 us per 1000 documents, which is 4 % of the search step).
// WORD = uint64_t: native ciphertexts; uint32_t: the 32-bit wire form (arithmetic on 32 bits).
__global__ void pack_key_kernel(const uint8_t* __restrict__ key, int n, uint32_t* __restrict__ bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n / 32) return;
    uint32_t w = 0;
    for (int b = 0; b < 32; ++b) {
        const int j = i * 32 + b;
        if (j < n) w |= (uint32_t)(key[j] & 1u) << b;
    }
    bits[i] = w;
}

cudaError_t launch_pack_key(const uint8_t* d_key, int n, uint32_t* d_bits, cudaStream_t s) {
    const int words = n / 32 + 1;
    pack_key_kernel<<<(words + 127) / 128, 128, 0, s>>>(d_key, n, d_bits);
    count_launch();
    return cudaGetLastError();
}

template <typename WORD, int VW>
struct RowVec;
template <>
struct RowVec<uint64_t, 2> {
    uint64_t w[2];
    __device__ __forceinline__ void load(const uint64_t* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p));
    }
};
template <>
struct RowVec<uint32_t, 4> {
    uint32_t w[4];
    __device__ __forceinline__ void load(const uint32_t* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    }
};
template <>
struct RowVec<uint32_t, 2> {
    uint32_t w[2];
    __device__ __forceinline__ void load(const uint32_t* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    }
};

constexpr int SD_THREADS = 256;
constexpr int SD_DEPTH = 6;   // loads in flight per lane

// DPC = documents per CTA: the 8 warps are shared by the DPC * M ciphertexts of DPC consecutive documents.  The
// 32-bit wire form has half as many 16-byte vectors per row, so it takes two documents per CTA to keep ~6 loads
// in flight per lane -- and half the CTA slot-time, which matters on the client GPU of a multi-GPU search where
// this kernel runs next to the HBM-bound dot product of the next step.
template <typename WORD, int VW, int DPC>
__global__ void __launch_bounds__(SD_THREADS)
similarity_decrypt_kernel(const uint32_t* __restrict__ kbits, int n, int64_t stride, const WORD* __restrict__ cts, int64_t B,
                          int M, int shift, int64_t zp_w, int64_t q_bias, double out_scale, int64_t out_zp,
                          double* __restrict__ y, int64_t* __restrict__ q_y) {
    __shared__ WORD part[SD_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = (SD_THREADS / 32) / (M * DPC);         // warps per ciphertext (M = 1 or 2)
    const int ci = warp / wpc;                             // ciphertext of this CTA: document ci / M, output ci % M
    const int t = (warp - ci * wpc) * 32 + lane, T = wpc * 32;
    const int64_t doc0 = (int64_t)blockIdx.x * DPC;
    const bool live = doc0 + ci / M < B;
    const WORD* ct = cts + ((size_t)doc0 * M + ci) * stride;
    const int nvec = live ? n / VW : 0;
    WORD body[2] = {0, 0};
    if (threadIdx.x < DPC && doc0 + threadIdx.x < B) {   // thread dd finalises document dd: its bodies are in flight
        body[0] = cts[((size_t)(doc0 + threadIdx.x) * M) * stride + n];   // behind the mask loads
        if (M == 2) body[1] = cts[((size_t)(doc0 + threadIdx.x) * M + 1) * stride + n];
    }
    WORD dot = 0;
    for (int v0 = t; v0 < nvec; v0 += T * SD_DEPTH) {
        RowVec<WORD, VW> x[SD_DEPTH];
#pragma unroll
        for (int k = 0; k < SD_DEPTH; ++k) {
            const int v = v0 + k * T;
            if (v < nvec) x[k].load(ct + (size_t)v * VW);
        }
#pragma unroll
        for (int k = 0; k < SD_DEPTH; ++k) {
            const int v = v0 + k * T;
            if (v < nvec) {
                const uint32_t bits = __ldg(kbits + ((v * VW) >> 5)) >> ((v * VW) & 31);   // VW divides 32
#pragma unroll
                for (int i = 0; i < VW; ++i) dot += x[k].w[i] & (WORD)(0 - (WORD)((bits >> i) & 1u));
            }
        }
    }
    if (live && t < n - nvec * VW) {   // mask words past the last full vector
        const int w = nvec * VW + t;
        dot += ct[w] & (WORD)(0 - (WORD)((__ldg(kbits + (w >> 5)) >> (w & 31)) & 1u));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) part[warp] = dot;
    __syncthreads();
    if (threadIdx.x < DPC && doc0 + threadIdx.x < B) {
        const int dd = threadIdx.x;
        int64_t msg[2] = {0, 0};
#pragma unroll
        for (int mm = 0; mm < 2; ++mm) {
            if (mm >= M) break;
            WORD d = 0;
            for (int w = 0; w < wpc; ++w) d += part[(dd * M + mm) * wpc + w];
            const WORD mu = body[mm] - d;
            const WORD v = mu + (shift > 0 ? ((WORD)1 << (shift - 1)) : (WORD)0);
            if (sizeof(WORD) == 8) msg[mm] = (int64_t)v >> shift;
            else msg[mm] = (int64_t)((int32_t)v >> shift);
        }
        // q_y = m0 - zp_w * m1 + q_bias;  y = out_scale * (q_y - out_zp)   
        const int64_t q = msg[0] - zp_w * msg[1] + q_bias;
        if (q_y) q_y[doc0 + dd] = q;
        if (y) y[doc0 + dd] = __dmul_rn(out_scale, (double)(q - out_zp));
    }
}

cudaError_t launch_similarity_decrypt(const uint32_t* d_kbits, int n, int64_t stride, const void* d_cts, bool wire32,
                                      int64_t B, int M, int shift, int64_t zp_w, int64_t q_bias, double out_scale,
                                      int64_t out_zp, double* d_y, int64_t* d_q_y, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    if (B > 0x7fffffffLL || (M != 1 && M != 2)) return cudaErrorInvalidValue;
    if (!wire32)
        similarity_decrypt_kernel<uint64_t, 2, 1><<<(unsigned)B, SD_THREADS, 0, s>>>(
            d_kbits, n, stride, (const uint64_t*)d_cts, B, M, shift, zp_w, q_bias, out_scale, out_zp, d_y, d_q_y);
    else if (stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_cts) & 15) == 0)
        similarity_decrypt_kernel<uint32_t, 4, 2><<<(unsigned)((B + 1) / 2), SD_THREADS, 0, s>>>(
            d_kbits, n, stride, (const uint32_t*)d_cts, B, M, shift, zp_w, q_bias, out_scale, out_zp, d_y, d_q_y);
    else
        similarity_decrypt_kernel<uint32_t, 2, 1><<<(unsigned)B, SD_THREADS, 0, s>>>(
            d_kbits, n, stride, (const uint32_t*)d_cts, B, M, shift, zp_w, q_bias, out_scale, out_zp, d_y, d_q_y);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fhe
