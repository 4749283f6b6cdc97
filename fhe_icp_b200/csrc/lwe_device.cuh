// lwe_device.cuh -- warp-level LWE encryption shared by lwe.cu (inputs) and keys.cu (KSK).
#pragma once
#include "common.cuh"

namespace fhe {

#ifndef ENC_UNROLL
#define ENC_UNROLL 2
#endif
constexpr int ENC_UNROLL_N = ENC_UNROLL;   // independent Philox blocks in flight per lane (encryption)

// pack the 0/1 key bytes into 32-bit words in shared memory (block-wide; caller syncs)
__device__ __forceinline__ void pack_key_bits(const uint8_t* __restrict__ key, int n, uint32_t* skey) {
    const int kw = (n + 31) / 32 + 1;
    const bool aligned = (reinterpret_cast<uintptr_t>(key) & 3) == 0;
    for (int i = threadIdx.x; i < kw; i += blockDim.x) {
        uint32_t w = 0;
        if (aligned && i * 32 + 32 <= n) {
            // 8 independent 4-byte loads; each byte holds 0/1
            const uint32_t* p = reinterpret_cast<const uint32_t*>(key + i * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t v = p[q];
                w |= ((v & 1u) | ((v >> 7) & 2u) | ((v >> 14) & 4u) | ((v >> 21) & 8u)) << (4 * q);
            }
        } else {
            for (int b = 0; b < 32; ++b) {
                int j = i * 32 + b;
                if (j < n) w |= (uint32_t)(key[j] & 1u) << b;
            }
        }
        skey[i] = w;
    }
}

// One warp writes ciphertext `id`: mask from Philox blocks (two words each, one 128-bit
// store per lane per block => 512 B per warp instruction), body = <a,s> + plaintext + e.
// `noise` is the ciphertext's rounded Gaussian error, already computed by the caller: the Box-Muller
// chain is ~250 dependent FP64 instructions, so callers evaluate it for 32 ciphertexts at once (one
// per lane) instead of on lane 0 of every ciphertext.
__device__ __forceinline__ void warp_lwe_encrypt(const uint32_t* skey, int n, int64_t stride, uint64_t plaintext,
                                                 int64_t noise, uint64_t seed, uint32_t purpose, uint64_t id,
                                                 uint64_t* __restrict__ ct, int lane) {
    const uint32_t dom = FHE_B200_KIND_MASK | (purpose << 8);
    const PhiloxKeys K(seed);
    uint64_t dot = 0;
    const int nblk = (n + 1) / 2;
#pragma unroll ENC_UNROLL_N
    for (int blk = lane; blk < nblk; blk += 32) {
        u32x4 r = rng_block(K, dom, id, (uint32_t)blk);
        uint64_t a0 = lo64(r), a1 = hi64(r);
        const int w = 2 * blk;
        uint32_t bits = skey[w >> 5] >> (w & 31);  // w even => bits w and w+1 share a word
        dot += a0 & (0 - (uint64_t)(bits & 1u));
        if (w + 1 < n) {
            dot += a1 & (0 - (uint64_t)((bits >> 1) & 1u));
            if ((stride & 1) == 0) st_stream_u64x2(ct + w, u64x2{a0, a1});
            else { ct[w] = a0; ct[w + 1] = a1; }
        } else {
            ct[w] = a0;  // n odd: the second word of the last block is unused
        }
    }
    dot = warp_sum_u64(dot);
    if (lane == 0) ct[n] = dot + plaintext + (uint64_t)noise;
    for (int64_t w = n + 1 + lane; w < stride; w += 32) ct[w] = 0;
}

}  // namespace fhe
