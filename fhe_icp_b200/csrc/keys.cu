// keys.cu -- evaluation-key generation on the device: keyswitching key and bootstrapping
// key (standard domain), bit-identical to the oracle's because both draw from the same
// counter-based streams and all arithmetic is wrapping u64.
// These are setup kernels, not the hot path.
#include "common.cuh"
#include "kernels.h"
#include "lwe_device.cuh"

namespace fhe {

// ksk[j][lev][0..n] = LWE_s( S_j << (64 - beta*(lev+1)) ), ciphertext id = j*l + lev
constexpr int KSK_WARPS = 8;

__global__ void __launch_bounds__(KSK_WARPS * 32)
ksk_gen_kernel(const uint8_t* __restrict__ S_big, const uint8_t* __restrict__ s_small, int n, int64_t kN, int l,
               int beta, double sigma_abs, uint64_t evk_seed, uint64_t* __restrict__ ksk) {
    extern __shared__ uint32_t skey[];
    pack_key_bits(s_small, n, skey);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * KSK_WARPS + (threadIdx.x >> 5);
    if (c >= kN * l) return;
    const int64_t j = c / l;
    const int lev = (int)(c - j * l);
    const uint64_t pt = (uint64_t)(S_big[j] & 1u) << (64 - beta * (lev + 1));
    int64_t e = 0;  // setup kernel: lane 0 evaluates the Gaussian, no batching needed
    if (lane == 0) e = gaussian_i64(evk_seed, FHE_B200_KIND_NOISE | (FHE_B200_PUR_KSK << 8), (uint64_t)c, 0, sigma_abs);
    warp_lwe_encrypt(skey, n, n + 1, pt, e, evk_seed, FHE_B200_PUR_KSK, (uint64_t)c, ksk + c * (n + 1), lane);
}

cudaError_t launch_ksk_gen(const fhe_b200_pbs_params& p, const uint8_t* d_S_big, const uint8_t* d_s_small,
                           uint64_t evk_seed, uint64_t* d_ksk, cudaStream_t s) {
    const int64_t kN = (int64_t)p.k * p.N;
    const int64_t cnt = kN * p.l_ks;
    size_t smem = ((size_t)(p.n + 31) / 32 + 1) * sizeof(uint32_t);
    ksk_gen_kernel<<<(unsigned)((cnt + KSK_WARPS - 1) / KSK_WARPS), KSK_WARPS * 32, smem, s>>>(
        d_S_big, d_s_small, p.n, kN, p.l_ks, p.beta_ks, p.sigma_lwe_abs, evk_seed, d_ksk);
    count_launch();
    return cudaGetLastError();
}

// bsk[i][t][lev][c][N]: row R = (i*(k+1) + t)*l + lev is a GLWE encryption of zero under S
// (mask polys c < k from the MASK stream, body = sum_c A_c * S_c + E) plus the gadget term
// s_i << (64 - beta*(lev+1)) on the constant coefficient of component t.
constexpr int BSK_THREADS = 256;

__global__ void __launch_bounds__(BSK_THREADS)
bsk_gen_kernel(const uint8_t* __restrict__ s_small, const uint8_t* __restrict__ S_big, int k, int N, int l, int beta,
               double sigma_abs, uint64_t evk_seed, int group, uint32_t purpose, uint64_t* __restrict__ bsk) {
    extern __shared__ uint64_t sm[];  // A[N] then S bits [N/32 words]
    uint64_t* A = sm;
    uint32_t* Sb = reinterpret_cast<uint32_t*>(sm + N);
    // group == 1: row R = (i*(k+1)+t)*l+lev encrypts s_i.  group == 3 (multi-bit, pairs of key bits):
    // R = ((i*3+g)*(k+1)+t)*l+lev encrypts s_2i*s_2i+1, s_2i*(1-s_2i+1), (1-s_2i)*s_2i+1 for g = 0,1,2.
    const int64_t R = blockIdx.x;
    const int lev = (int)(R % l);
    const int t = (int)((R / l) % (k + 1));
    const int64_t ig = R / ((int64_t)l * (k + 1));
    const int g = (int)(ig % group);
    const int i = (int)(ig / group);
    int bit;
    if (group == 1) {
        bit = s_small[i] & 1;
    } else {
        const int sa = s_small[2 * i] & 1, sb = s_small[2 * i + 1] & 1;
        bit = g == 0 ? (sa & sb) : (g == 1 ? (sa & (1 - sb)) : ((1 - sa) & sb));
    }
    uint64_t* row = bsk + (size_t)R * (k + 1) * N;
    constexpr int PER = 16;  // coefficients per thread, N <= BSK_THREADS * PER
    uint64_t body[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int x = threadIdx.x + u * BSK_THREADS;
        body[u] = x < N ? (uint64_t)gaussian_i64(evk_seed, FHE_B200_KIND_NOISE | (purpose << 8), (uint64_t)R,
                                                 (uint32_t)x, sigma_abs)
                        : 0;
    }
    for (int c = 0; c < k; ++c) {
        __syncthreads();
        for (int x = threadIdx.x; x < N; x += BSK_THREADS) {
            uint64_t a = mask_word(evk_seed, purpose, (uint64_t)R, (int64_t)c * N + x);
            A[x] = a;
            row[(size_t)c * N + x] = a;
        }
        for (int w = threadIdx.x; w < N / 32; w += BSK_THREADS) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; ++b) bits |= (uint32_t)(S_big[(size_t)c * N + w * 32 + b] & 1u) << b;
            Sb[w] = bits;
        }
        __syncthreads();
        // body[x] += sum_{y : S[y]=1} (X^y * A)[x] = sum_y S[y] * (x >= y ? A[x-y] : -A[x-y+N])
        for (int w = 0; w < N / 32; ++w) {
            uint32_t bits = Sb[w];
            while (bits) {
                const int y = w * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int x = threadIdx.x + u * BSK_THREADS;
                    if (x < N) body[u] += (x >= y) ? A[x - y] : (uint64_t)0 - A[x - y + N];
                }
            }
        }
    }
    const uint64_t gd = (uint64_t)bit << (64 - beta * (lev + 1));
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int x = threadIdx.x + u * BSK_THREADS;
        if (x < N) {
            uint64_t v = body[u];
            if (t == k && x == 0) v += gd;
            row[(size_t)k * N + x] = v;
        }
    }
    if (t < k) {
        __syncthreads();  // mask polynomials of this row are all written by this block
        if (threadIdx.x == 0) row[(size_t)t * N] += gd;
    }
}

// Generic GLWE row encryption for the packed inner-product path: row R of out [rows][k+1][N] is
// GLWE_S(0) + (msg_R(X) << shift_R) on component comp_R.  mode 0 (document vectors): msg_R = msgs +
// R*msg_stride, shift_R = shift, comp_R = k (body).  mode 1 (GGSW of one polynomial): R = t*l + lev,
// msg_R = msgs, shift_R = 64 - beta*(lev+1), comp_R = t.  Randomness of row R: object id_base + R,
// purpose GLWE.  Same structure as bsk_gen_kernel (body = sum_c A_c * S_c + E by sparse negacyclic adds).
__global__ void __launch_bounds__(BSK_THREADS)
glwe_encrypt_rows_kernel(const uint8_t* __restrict__ S_big, const int64_t* __restrict__ msgs, int64_t msg_stride,
                         int mode, int shift, int k, int N, int l, int beta, double sigma_abs, uint64_t seed,
                         uint64_t noise_seed, uint64_t id_base, uint64_t* __restrict__ out) {
    extern __shared__ uint64_t sm[];  // A[N] then S bits [N/32 words]
    uint64_t* A = sm;
    uint32_t* Sb = reinterpret_cast<uint32_t*>(sm + N);
    const int64_t R = blockIdx.x;
    const uint64_t id = id_base + (uint64_t)R;
    uint64_t* row = out + (size_t)R * (k + 1) * N;
    const int64_t* m = mode == 0 ? msgs + (size_t)R * msg_stride : msgs;
    const int sh = mode == 0 ? shift : 64 - beta * ((int)(R % l) + 1);
    const int comp = mode == 0 ? k : (int)(R / l);
    constexpr int PER = 16;
    uint64_t body[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int x = threadIdx.x + u * BSK_THREADS;
        body[u] = x < N ? (uint64_t)gaussian_i64(noise_seed, FHE_B200_KIND_NOISE | (FHE_B200_PUR_GLWE << 8), id, (uint32_t)x,
                                                 sigma_abs)
                        : 0;
    }
    for (int c = 0; c < k; ++c) {
        __syncthreads();
        for (int x = threadIdx.x; x < N; x += BSK_THREADS) {
            const uint64_t a = mask_word(seed, FHE_B200_PUR_GLWE, id, (int64_t)c * N + x);
            A[x] = a;
            row[(size_t)c * N + x] = a + (c == comp ? (uint64_t)m[x] << sh : 0);
        }
        for (int w = threadIdx.x; w < N / 32; w += BSK_THREADS) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; ++b) bits |= (uint32_t)(S_big[(size_t)c * N + w * 32 + b] & 1u) << b;
            Sb[w] = bits;
        }
        __syncthreads();
        for (int w = 0; w < N / 32; ++w) {
            uint32_t bits = Sb[w];
            while (bits) {
                const int y = w * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int x = threadIdx.x + u * BSK_THREADS;
                    if (x < N) body[u] += (x >= y) ? A[x - y] : (uint64_t)0 - A[x - y + N];
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int x = threadIdx.x + u * BSK_THREADS;
        if (x < N) row[(size_t)k * N + x] = body[u] + (comp == k ? (uint64_t)m[x] << sh : 0);
    }
}

cudaError_t launch_glwe_encrypt_rows(const fhe_b200_pbs_params& p, const uint8_t* d_S_big, const int64_t* d_msgs,
                                     int64_t rows, int64_t msg_stride, int mode, int shift, uint64_t seed,
                                     uint64_t noise_seed, uint64_t id_base, uint64_t* d_out, cudaStream_t s) {
    if (p.N > BSK_THREADS * 16) return cudaErrorInvalidValue;
    if (rows <= 0) return cudaSuccess;
    size_t smem = (size_t)p.N * 8 + (size_t)p.N / 8;
    glwe_encrypt_rows_kernel<<<(unsigned)rows, BSK_THREADS, smem, s>>>(d_S_big, d_msgs, msg_stride, mode, shift, p.k, p.N,
                                                                      p.l_pbs, p.beta_pbs, p.sigma_glwe_abs, seed, noise_seed, id_base,
                                                                      d_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_bsk_gen(const fhe_b200_pbs_params& p, const uint8_t* d_s_small, const uint8_t* d_S_big,
                           uint64_t evk_seed, uint64_t* d_bsk, cudaStream_t s) {
    if (p.N > BSK_THREADS * 16) return cudaErrorInvalidValue;
    const int64_t rows = (int64_t)p.n * (p.k + 1) * p.l_pbs;
    size_t smem = (size_t)p.N * 8 + (size_t)p.N / 8;
    bsk_gen_kernel<<<(unsigned)rows, BSK_THREADS, smem, s>>>(d_s_small, d_S_big, p.k, p.N, p.l_pbs, p.beta_pbs,
                                                            p.sigma_glwe_abs, evk_seed, 1, FHE_B200_PUR_BSK, d_bsk);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_bsk2_gen(const fhe_b200_pbs_params& p, const uint8_t* d_s_small, const uint8_t* d_S_big,
                            uint64_t evk_seed, uint64_t* d_bsk2, cudaStream_t s) {
    if (p.N > BSK_THREADS * 16 || (p.n & 1)) return cudaErrorInvalidValue;
    const int64_t rows = (int64_t)(p.n / 2) * 3 * (p.k + 1) * p.l_pbs;
    size_t smem = (size_t)p.N * 8 + (size_t)p.N / 8;
    bsk_gen_kernel<<<(unsigned)rows, BSK_THREADS, smem, s>>>(d_s_small, d_S_big, p.k, p.N, p.l_pbs, p.beta_pbs,
                                                            p.sigma_glwe_abs, evk_seed, 3, FHE_B200_PUR_BSK2, d_bsk2);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fhe
