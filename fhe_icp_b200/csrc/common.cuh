// common.cuh -- shared device helpers: deterministic counter-based randomness,
// streaming loads/stores, error plumbing.  sm_100a only.
//
// The randomness spec (DESIGN.md "Deterministic randomness") is implemented twice on
// purpose: here for the product and, independently, in oracle/fhe_oracle.c for the
// checker, so that "same keys, seeds and inputs" means bit-identical ciphertexts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fhe_b200.h"

namespace fhe {

// ----------------------------------------------------------------------------- Philox4x32-R
// Secret material (key bits, noise) is drawn with the standard 10 rounds.  PUBLIC masks -- which both the client and
// the evaluator regenerate, 2 x 32x32->64 multiplies per round on the scarce wide-multiplier pipe -- use
// MASK_ROUNDS = 7, the smallest round count Salmon et al. (SC'11, table 2) report as Crush-resistant (BigCrush clean);
// a mask only has to be uniform, it is public.  Same (seed, kind, object, block) interface for both.
struct u32x4 { uint32_t x, y, z, w; };
constexpr int MASK_ROUNDS = FHE_B200_MASK_ROUNDS;

__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    // one 32x32->64 multiply (IMAD.WIDE.U32 on the device) instead of separate lo / hi products
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}

template <int ROUNDS>
__host__ __device__ __forceinline__ u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c.x, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c, k0, k1);
}

// Round keys k + r*(W0,W1) of the MASK generator depend only on the seed: hot loops expand them once per thread.
struct PhiloxKeys {
    uint32_t k0[MASK_ROUNDS], k1[MASK_ROUNDS];
    __host__ __device__ __forceinline__ explicit PhiloxKeys(uint64_t seed) {
        uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < MASK_ROUNDS; ++r) {
            k0[r] = a;
            k1[r] = b;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
    }
};

// mask block (MASK_ROUNDS rounds) from pre-expanded round keys
__host__ __device__ __forceinline__ u32x4 rng_block(const PhiloxKeys& K, uint32_t domain, uint64_t obj, uint32_t blk) {
    u32x4 c{blk, (uint32_t)obj, (uint32_t)(obj >> 32), domain};
#pragma unroll
    for (int r = 0; r < MASK_ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c.x, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ c.y ^ K.k0[r];
        n.y = lo1;
        n.z = hi0 ^ c.w ^ K.k1[r];
        n.w = lo0;
        c = n;
    }
    return c;
}

// counter = (blk, obj_lo, obj_hi, domain), key = seed; 10 rounds (secret key bits, noise)
__host__ __device__ __forceinline__ u32x4 rng_block(uint64_t seed, uint32_t domain, uint64_t obj, uint32_t blk) {
    u32x4 c{blk, (uint32_t)obj, (uint32_t)(obj >> 32), domain};
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__host__ __device__ __forceinline__ uint64_t lo64(const u32x4& r) { return ((uint64_t)r.y << 32) | r.x; }
__host__ __device__ __forceinline__ uint64_t hi64(const u32x4& r) { return ((uint64_t)r.w << 32) | r.z; }

// ----------------------------------------------------------------------------- deterministic f64 math
// Only IEEE-exact operations, never contracted: explicit *_rn intrinsics on the device.
#ifdef __CUDA_ARCH__
#define FHE_DMUL(a, b) __dmul_rn((a), (b))
#define FHE_DADD(a, b) __dadd_rn((a), (b))
#define FHE_DSUB(a, b) __dsub_rn((a), (b))
#define FHE_DDIV(a, b) __ddiv_rn((a), (b))
#define FHE_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define FHE_DSQRT(a) __dsqrt_rn((a))
#else
#include <cmath>
#define FHE_DMUL(a, b) ((a) * (b))
#define FHE_DADD(a, b) ((a) + (b))
#define FHE_DSUB(a, b) ((a) - (b))
#define FHE_DDIV(a, b) ((a) / (b))
#define FHE_DFMA(a, b, c) std::fma((a), (b), (c))
#define FHE_DSQRT(a) std::sqrt((a))
#endif

__host__ __device__ __forceinline__ double det_log(double x) {
    const double LOGC[12] = {
        0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3, 0x1.c71c71c71c71cp-4,
        0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4, 0x1.e1e1e1e1e1e1ep-5,
        0x1.af286bca1af28p-5, 0x1.8618618618618p-5, 0x1.642c8590b2164p-5, 0x1.47ae147ae147bp-5};
#ifdef __CUDA_ARCH__
    uint64_t bits = (uint64_t)__double_as_longlong(x);
#else
    uint64_t bits;
    memcpy(&bits, &x, 8);
#endif
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    uint64_t mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
#ifdef __CUDA_ARCH__
    double m = __longlong_as_double((long long)mb);
#else
    double m;
    memcpy(&m, &mb, 8);
#endif
    if (m > 0x1.6a09e667f3bcdp+0) { m = FHE_DMUL(m, 0.5); e += 1; }
    double f = FHE_DDIV(FHE_DSUB(m, 1.0), FHE_DADD(m, 1.0));
    double s = FHE_DMUL(f, f);
    double p = LOGC[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = FHE_DFMA(p, s, LOGC[i]);
    double sp = FHE_DMUL(s, p);
    double g = FHE_DFMA(sp, 2.0, 2.0);
    double lm = FHE_DMUL(f, g);
    return FHE_DFMA((double)e, 0x1.62e42fefa39efp-1, lm);
}

__host__ __device__ __forceinline__ double det_cos2pi_k53(uint64_t k53) {
    const double COSC[11] = {
        0x1.0000000000000p+0, -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10,
        0x1.a01a01a01a01ap-16, -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37,
        0x1.ae7f3e733b81fp-45, -0x1.6827863b97d97p-53, 0x1.e542ba4020225p-62};
    const double SINC[10] = {
        -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19,
        -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49,
        -0x1.2f49b46814157p-57, 0x1.71b8ef6dcf572p-66};
    uint64_t q = (k53 + (1ULL << 50)) >> 51;
    int64_t r = (int64_t)k53 - (int64_t)(q << 51);
    double t = FHE_DMUL((double)r, 0x1p-53);
    double x = FHE_DMUL(t, 0x1.921fb54442d18p+2);
    double x2 = FHE_DMUL(x, x);
    double c = COSC[10];
#pragma unroll
    for (int i = 9; i >= 0; --i) c = FHE_DFMA(c, x2, COSC[i]);
    double sn = SINC[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) sn = FHE_DFMA(sn, x2, SINC[i]);
    double x3 = FHE_DMUL(x, x2);
    double s = FHE_DFMA(x3, sn, x);
    switch ((int)(q & 3)) {
        case 0: return c;
        case 1: return -s;
        case 2: return -c;
        default: return s;
    }
}

__host__ __device__ __forceinline__ double normal_from_block(const u32x4& r) {
    uint64_t k1 = lo64(r) >> 11;
    uint64_t k2 = hi64(r) >> 11;
    double u1 = FHE_DMUL((double)(k1 + 1), 0x1p-53);
    double lg = det_log(u1);
    double rad = FHE_DSQRT(FHE_DMUL(-2.0, lg));
    return FHE_DMUL(rad, det_cos2pi_k53(k2));
}

__host__ __device__ __forceinline__ int64_t gaussian_i64(uint64_t seed, uint32_t domain, uint64_t obj,
                                                        uint32_t blk, double sigma_abs) {
    double z = normal_from_block(rng_block(seed, domain, obj, blk));
    double v = FHE_DMUL(z, sigma_abs);
#ifdef __CUDA_ARCH__
    return (int64_t)__double2ll_rn(v);
#else
    return (int64_t)llrint(v);
#endif
}

// mask word w of ciphertext/row `obj`
__host__ __device__ __forceinline__ uint64_t mask_word(uint64_t seed, uint32_t purpose, uint64_t obj, int64_t w) {
    u32x4 c{(uint32_t)(w >> 1), (uint32_t)obj, (uint32_t)(obj >> 32), FHE_B200_KIND_MASK | (purpose << 8)};
    u32x4 r = philox4x32<MASK_ROUNDS>(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (w & 1) ? hi64(r) : lo64(r);
}

// ----------------------------------------------------------------------------- memory helpers
#ifdef __CUDACC__
struct __align__(16) u64x2 { uint64_t x, y; };

// streaming 128-bit load: read-only path, do not allocate in L1 (each byte is read once)
__device__ __forceinline__ u64x2 ld_stream_u64x2(const uint64_t* p) {
    u64x2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
// four independent streaming loads issued back to back (one asm block, so the compiler
// cannot interleave dependent math between them and shrink the memory-level parallelism)
__device__ __forceinline__ void ld_stream_u64x2_x4(const uint64_t* p0, const uint64_t* p1, const uint64_t* p2,
                                                   const uint64_t* p3, u64x2& a, u64x2& b, u64x2& c, u64x2& d) {
    asm volatile(
        "ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%8];\n\t"
        "ld.global.nc.L1::no_allocate.v2.u64 {%2, %3}, [%9];\n\t"
        "ld.global.nc.L1::no_allocate.v2.u64 {%4, %5}, [%10];\n\t"
        "ld.global.nc.L1::no_allocate.v2.u64 {%6, %7}, [%11];"
        : "=l"(a.x), "=l"(a.y), "=l"(b.x), "=l"(b.y), "=l"(c.x), "=l"(c.y), "=l"(d.x), "=l"(d.y)
        : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}
__device__ __forceinline__ void st_stream_u64x2(uint64_t* p, const u64x2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}
// ---- mbarrier + TMA bulk copy (cp.async.bulk, 1-D): global -> shared without registers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one non-blocking probe of the barrier's phase (the instruction itself takes ~90 clocks; issue it early)
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The same sum (mod 2^64) with three REDUX instructions instead of five dependent shuffle rounds: the word is cut into
// limbs of 22 / 22 / 20 bits, whose 32-lane sums fit 32 bits, and the three hardware reductions are independent.
__device__ __forceinline__ uint64_t warp_sum_u64_redux(uint64_t v) {
    const uint32_t a = (uint32_t)v & 0x3fffffu, b = (uint32_t)(v >> 22) & 0x3fffffu, c = (uint32_t)(v >> 44);
    const uint64_t sa = __reduce_add_sync(0xffffffffu, a), sb = __reduce_add_sync(0xffffffffu, b),
                   sc = __reduce_add_sync(0xffffffffu, c);
    return sa + (sb << 22) + (sc << 44);
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace fhe
