/*
 * fhe_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See fhe_oracle.h.
 *
 * Restates, in plain C, the TFHE arithmetic the reference reaches through
 * `model.predict(X, fhe="execute")` (/root/reference/fhe_similarity.py:151) and the
 * keyswitch / programmable-bootstrap primitives that BASELINE.json's north star names
 * (algorithms as published for TFHE: Chillotti-Gama-
 * Georgieva-Izabachene, J. Cryptology 2020, and Concrete's "keyswitch -> PBS" atomic
 * pattern).  Build: `make -C oracle` (gcc -O2 -ffp-contract=off; see Makefile).
 *
 * Parity status: ciphertext-level "parity unpinned" (no goldens exist upstream);
 * pinned by Random123 KATs + decrypt==clear-integer-circuit + decrypt(PBS)==LUT.
 */
#include "fhe_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Philox4x32-R (Salmon et al., SC'11), counter-based, reproducible.   */
/* 10 rounds for secret streams (key bits, noise); ORC_MASK_ROUNDS = 7  */
/* (the paper's smallest Crush-resistant count) for the PUBLIC masks.   */
/* ------------------------------------------------------------------ */
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    orc_philox4x32_r(ctr, key, 10, out);
}

/* counter = (blk, obj_lo, obj_hi, domain); key = seed; round count by stream kind */
void orc_rng_block(uint64_t seed, uint32_t domain, uint64_t obj, uint32_t blk, uint32_t out[4]) {
    uint32_t ctr[4] = {blk, (uint32_t)obj, (uint32_t)(obj >> 32), domain};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_r(ctr, key, (domain & 0xff) == ORC_KIND_MASK ? ORC_MASK_ROUNDS : 10, out);
}

/* ------------------------------------------------------------------ */
/* Deterministic double-precision log / cos built only from IEEE-exact */
/* operations (+ - * / sqrt fma), so that gcc and nvcc agree bit for    */
/* bit.  This file is compiled with -ffp-contract=off.                  */
/* ------------------------------------------------------------------ */
static const double LOGC[12] = {
    0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3, 0x1.c71c71c71c71cp-4,
    0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4, 0x1.e1e1e1e1e1e1ep-5,
    0x1.af286bca1af28p-5, 0x1.8618618618618p-5, 0x1.642c8590b2164p-5, 0x1.47ae147ae147bp-5};
static const double COSC[11] = {
    0x1.0000000000000p+0, -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10,
    0x1.a01a01a01a01ap-16, -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37,
    0x1.ae7f3e733b81fp-45, -0x1.6827863b97d97p-53, 0x1.e542ba4020225p-62};
static const double SINC[10] = {
    -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19,
    -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49,
    -0x1.2f49b46814157p-57, 0x1.71b8ef6dcf572p-66};
#define ORC_LN2 0x1.62e42fefa39efp-1
#define ORC_TWO_PI 0x1.921fb54442d18p+2
#define ORC_SQRT2 0x1.6a09e667f3bcdp+0

/* natural log of a normal double in (0, 1] (any positive normal works) */
double orc_det_log(double x) {
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    uint64_t mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &mb, 8);
    if (m > ORC_SQRT2) { m = m * 0.5; e += 1; }
    double f = (m - 1.0) / (m + 1.0);
    double s = f * f;
    double p = LOGC[11];
    for (int i = 10; i >= 0; --i) p = fma(p, s, LOGC[i]);
    double sp = s * p;
    double g = fma(sp, 2.0, 2.0);
    double lm = f * g;
    return fma((double)e, ORC_LN2, lm);
}

/* cos(2*pi*k53/2^53) for a 53-bit integer phase */
double orc_det_cos2pi_k53(uint64_t k53) {
    uint64_t q = (k53 + (1ULL << 50)) >> 51;            /* nearest quarter turn, 0..4 */
    int64_t r = (int64_t)k53 - (int64_t)(q << 51);      /* |r| <= 2^50 */
    double t = (double)r * 0x1p-53;                      /* exact, in [-1/8, 1/8] */
    double x = t * ORC_TWO_PI;
    double x2 = x * x;
    double c = COSC[10];
    for (int i = 9; i >= 0; --i) c = fma(c, x2, COSC[i]);
    double sn = SINC[9];
    for (int i = 8; i >= 0; --i) sn = fma(sn, x2, SINC[i]);
    double x3 = x * x2;
    double s = fma(x3, sn, x);
    switch ((int)(q & 3)) {
        case 0: return c;
        case 1: return -s;
        case 2: return -c;
        default: return s;
    }
}

/* Box-Muller (cosine branch) from one Philox block */
double orc_normal_from_block(const uint32_t r[4]) {
    uint64_t k1 = (((uint64_t)r[1] << 32) | r[0]) >> 11;
    uint64_t k2 = (((uint64_t)r[3] << 32) | r[2]) >> 11;
    double u1 = (double)(k1 + 1) * 0x1p-53; /* (0, 1] */
    double lg = orc_det_log(u1);
    double rad = sqrt(-2.0 * lg);
    return rad * orc_det_cos2pi_k53(k2);
}

int64_t orc_gaussian(uint64_t seed, uint32_t domain, uint64_t obj, uint32_t blk, double sigma_abs) {
    uint32_t r[4];
    orc_rng_block(seed, domain, obj, blk, r);
    double z = orc_normal_from_block(r);
    return (int64_t)llrint(z * sigma_abs);
}

/* ------------------------------------------------------------------ */
/* LWE                                         */
/* ------------------------------------------------------------------ */
void orc_secret_key(uint64_t key_seed, uint32_t key_id, int64_t dim, uint8_t *s) {
    for (int64_t j = 0; j < dim; j += 128) {
        uint32_t r[4];
        orc_rng_block(key_seed, ORC_KIND_SK | (key_id << 8), 0, (uint32_t)(j / 128), r);
        for (int b = 0; b < 128 && j + b < dim; ++b) s[j + b] = (r[b / 32] >> (b % 32)) & 1u;
    }
}

static inline uint64_t mask_word(uint64_t seed, uint32_t purpose, uint64_t obj, int64_t w) {
    uint32_t r[4];
    orc_rng_block(seed, ORC_KIND_MASK | (purpose << 8), obj, (uint32_t)(w >> 1), r);
    return (w & 1) ? (((uint64_t)r[3] << 32) | r[2]) : (((uint64_t)r[1] << 32) | r[0]);
}

void orc_lwe_encrypt_batch(const uint8_t *s, int32_t n, int64_t stride, const int64_t *msgs,
                           int64_t count, int32_t shift, double sigma_abs, uint64_t enc_seed,
                           uint64_t noise_seed, uint64_t ct_base, uint32_t purpose, uint64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < count; ++c) {
        uint64_t id = ct_base + (uint64_t)c;
        uint64_t *ct = out + c * stride;
        uint64_t dot = 0;
        for (int64_t w = 0; w < n; w += 2) {
            uint32_t r[4];
            orc_rng_block(enc_seed, ORC_KIND_MASK | (purpose << 8), id, (uint32_t)(w >> 1), r);
            uint64_t a0 = ((uint64_t)r[1] << 32) | r[0];
            uint64_t a1 = ((uint64_t)r[3] << 32) | r[2];
            ct[w] = a0;
            if (s[w]) dot += a0;
            if (w + 1 < n) {
                ct[w + 1] = a1;
                if (s[w + 1]) dot += a1;
            }
        }
        int64_t e = orc_gaussian(noise_seed, ORC_KIND_NOISE | (purpose << 8), id, 0, sigma_abs);
        ct[n] = dot + ((uint64_t)msgs[c] << shift) + (uint64_t)e;
        for (int64_t w = n + 1; w < stride; ++w) ct[w] = 0;
    }
}

void orc_lwe_phase_batch(const uint8_t *s, int32_t n, int64_t stride, const uint64_t *ct,
                         int64_t count, uint64_t *phase) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < count; ++c) {
        const uint64_t *x = ct + c * stride;
        uint64_t dot = 0;
        for (int32_t j = 0; j < n; ++j)
            if (s[j]) dot += x[j];
        phase[c] = x[n] - dot;
    }
}

/* m = round(phase / 2^shift), as a signed (64-shift)-bit two's complement value */
void orc_lwe_decrypt_batch(const uint8_t *s, int32_t n, int64_t stride, const uint64_t *ct,
                           int64_t count, int32_t shift, int64_t *out) {
    uint64_t *ph = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(count > 0 ? count : 1));
    orc_lwe_phase_batch(s, n, stride, ct, count, ph);
    for (int64_t c = 0; c < count; ++c) {
        uint64_t v = ph[c] + (shift > 0 ? (1ULL << (shift - 1)) : 0);
        out[c] = (int64_t)v >> shift;
    }
    free(ph);
}

/* out[b][m][:] = sum_j W[m][j] * ct[b][j][:]  (+ bias[m] << shift on the body) */
void orc_lincomb_batch(const uint64_t *ct, int64_t B, int32_t d, int32_t n, int64_t stride,
                       const int64_t *W, int32_t M, const int64_t *bias, int32_t shift,
                       uint64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        for (int32_t m = 0; m < M; ++m) {
            uint64_t *o = out + ((size_t)b * M + m) * stride;
            for (int64_t w = 0; w < stride; ++w) o[w] = 0;
            for (int32_t j = 0; j < d; ++j) {
                const uint64_t *x = ct + ((size_t)b * d + j) * stride;
                uint64_t wv = (uint64_t)W[(size_t)m * d + j];
                for (int32_t w = 0; w <= n; ++w) o[w] += wv * x[w];
            }
            if (bias) o[n] += (uint64_t)bias[m] << shift;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Gadget decomposition (closest representative, balanced digits)      */
/* ------------------------------------------------------------------ */
static inline void decompose(uint64_t a, int l, int beta, int64_t *dig /* [l], dig[0] most significant */) {
    int tot = l * beta;
    uint64_t st = (a + (1ULL << (63 - tot))) >> (64 - tot); /* round to tot bits */
    uint64_t Bm = (1ULL << beta) - 1, half = 1ULL << (beta - 1);
    for (int lev = l - 1; lev >= 0; --lev) {
        uint64_t dg = st & Bm;
        st >>= beta;
        if (dg >= half) {
            dig[lev] = (int64_t)dg - (int64_t)(1ULL << beta);
            st += 1;
        } else {
            dig[lev] = (int64_t)dg;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Keyswitch                                   */
/* ksk[j][lev][n+1] = LWE_s( S_j * 2^(64 - beta*(lev+1)) )               */
/* ------------------------------------------------------------------ */
void orc_ksk_gen(const orc_pbs_params *p, const uint8_t *S_big, const uint8_t *s_small,
                 uint64_t evk_seed, uint64_t *ksk) {
    int64_t kN = (int64_t)p->k * p->N;
    int n = p->n, l = p->l_ks, beta = p->beta_ks;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < kN; ++j) {
        for (int lev = 0; lev < l; ++lev) {
            int64_t msg = S_big[j];
            orc_lwe_encrypt_batch(s_small, n, n + 1, &msg, 1, 64 - beta * (lev + 1),
                                  p->sigma_lwe_abs, evk_seed, evk_seed, (uint64_t)(j * l + lev),
                                  ORC_PUR_KSK, ksk + ((size_t)j * l + lev) * (n + 1));
        }
    }
}

void orc_keyswitch_batch(const orc_pbs_params *p, const uint64_t *ksk, const uint64_t *in,
                         int64_t B, uint64_t *out) {
    int64_t kN = (int64_t)p->k * p->N;
    int n = p->n, l = p->l_ks, beta = p->beta_ks;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const uint64_t *x = in + (size_t)b * (kN + 1);
        uint64_t *o = out + (size_t)b * (n + 1);
        for (int w = 0; w < n; ++w) o[w] = 0;
        o[n] = x[kN];
        int64_t dig[16];
        for (int64_t j = 0; j < kN; ++j) {
            decompose(x[j], l, beta, dig);
            for (int lev = 0; lev < l; ++lev) {
                if (!dig[lev]) continue;
                const uint64_t *kr = ksk + ((size_t)j * l + lev) * (n + 1);
                uint64_t dv = (uint64_t)dig[lev];
                for (int w = 0; w <= n; ++w) o[w] -= dv * kr[w];
            }
        }
    }
}

/* 32-bit keyswitch ("KS32"): the keyswitch output only has to be accurate to the small key's noise
 * (2^-17), so key and accumulation are carried on the top 32 bits of the torus.
 * ksk32 = round(ksk / 2^32); out = ((round(b_in / 2^32) - sum digit * ksk32) mod 2^32) << 32. */
void orc_ksk_to_32(const orc_pbs_params *p, const uint64_t *ksk, uint32_t *ksk32) {
    int64_t words = (int64_t)p->k * p->N * p->l_ks * (p->n + 1);
    for (int64_t i = 0; i < words; ++i) ksk32[i] = (uint32_t)((ksk[i] + 0x80000000ULL) >> 32);
}

void orc_keyswitch32_batch(const orc_pbs_params *p, const uint32_t *ksk32, const uint64_t *in,
                           int64_t B, uint64_t *out) {
    int64_t kN = (int64_t)p->k * p->N;
    int n = p->n, l = p->l_ks, beta = p->beta_ks;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const uint64_t *x = in + (size_t)b * (kN + 1);
        uint64_t *o = out + (size_t)b * (n + 1);
        uint32_t acc[4097];
        for (int w = 0; w < n; ++w) acc[w] = 0;
        acc[n] = (uint32_t)((x[kN] + 0x80000000ULL) >> 32);
        int64_t dig[16];
        for (int64_t j = 0; j < kN; ++j) {
            decompose(x[j], l, beta, dig);
            for (int lev = 0; lev < l; ++lev) {
                if (!dig[lev]) continue;
                const uint32_t *kr = ksk32 + ((size_t)j * l + lev) * (n + 1);
                uint32_t dv = (uint32_t)(int32_t)dig[lev];
                for (int w = 0; w <= n; ++w) acc[w] -= dv * kr[w];
            }
        }
        for (int w = 0; w <= n; ++w) o[w] = (uint64_t)acc[w] << 32;
    }
}

/* ------------------------------------------------------------------ */
/* Negacyclic FFT (size N/2 complex, folding + twisting)                */
/* ------------------------------------------------------------------ */
typedef struct {
    int N, M, logM;
    double *tw_re, *tw_im;   /* twist omega^j = exp(i*pi*j/N), j < M */
    double *w_re, *w_im;     /* exp(+2*pi*i*k/M), k < M */
    int *rev;
} fft_plan;

static fft_plan g_plan = {0, 0, 0, 0, 0, 0, 0, 0};

static const fft_plan *get_plan(int N) {
#pragma omp critical(orc_plan)
    {
        if (g_plan.N != N) {
            free(g_plan.tw_re); free(g_plan.tw_im); free(g_plan.w_re); free(g_plan.w_im); free(g_plan.rev);
            int M = N / 2, logM = 0;
            while ((1 << logM) < M) ++logM;
            g_plan.M = M; g_plan.logM = logM;
            g_plan.tw_re = (double *)malloc(sizeof(double) * M);
            g_plan.tw_im = (double *)malloc(sizeof(double) * M);
            g_plan.w_re = (double *)malloc(sizeof(double) * M);
            g_plan.w_im = (double *)malloc(sizeof(double) * M);
            g_plan.rev = (int *)malloc(sizeof(int) * M);
            for (int j = 0; j < M; ++j) {
                double a = M_PI * (double)j / (double)N;
                g_plan.tw_re[j] = cos(a); g_plan.tw_im[j] = sin(a);
                double b = 2.0 * M_PI * (double)j / (double)M;
                g_plan.w_re[j] = cos(b); g_plan.w_im[j] = sin(b);
                int r = 0;
                for (int t = 0; t < logM; ++t) if (j & (1 << t)) r |= 1 << (logM - 1 - t);
                g_plan.rev[j] = r;
            }
            g_plan.N = N;
        }
    }
    return &g_plan;
}

/* in-place radix-2 DIT, natural order in and out; sign=+1: exp(+2 pi i jk/M) */
static void fft_inplace(const fft_plan *pl, double *re, double *im, int sign) {
    int M = pl->M;
    for (int j = 0; j < M; ++j) {
        int r = pl->rev[j];
        if (r > j) {
            double t = re[j]; re[j] = re[r]; re[r] = t;
            t = im[j]; im[j] = im[r]; im[r] = t;
        }
    }
    for (int len = 2; len <= M; len <<= 1) {
        int half = len >> 1, step = M / len;
        for (int base = 0; base < M; base += len) {
            for (int t = 0; t < half; ++t) {
                double wr = pl->w_re[t * step], wi = sign * pl->w_im[t * step];
                double xr = re[base + t + half], xi = im[base + t + half];
                double vr = xr * wr - xi * wi, vi = xr * wi + xi * wr;
                double ur = re[base + t], ui = im[base + t];
                re[base + t] = ur + vr; im[base + t] = ui + vi;
                re[base + t + half] = ur - vr; im[base + t + half] = ui - vi;
            }
        }
    }
}

/* forward: real coefficient array (as doubles) of length N -> M complex bins (natural order) */
static void nega_forward(const fft_plan *pl, const double *coef, double *re, double *im) {
    int M = pl->M;
    for (int j = 0; j < M; ++j) {
        double a = coef[j], b = coef[j + M];
        re[j] = a * pl->tw_re[j] - b * pl->tw_im[j];
        im[j] = a * pl->tw_im[j] + b * pl->tw_re[j];
    }
    fft_inplace(pl, re, im, +1);
}

static inline uint64_t f64_to_torus(double x) {
    double r = rint(x * 0x1p-64);
    double y = x - r * 0x1p64;
    if (y >= 0x1p63) y -= 0x1p64;
    if (y < -0x1p63) y += 0x1p64;
    return (uint64_t)(int64_t)llrint(y);
}

/* inverse: bins -> N torus coefficients, ADDED into acc (wrapping) */
static void nega_inverse_add(const fft_plan *pl, double *re, double *im, uint64_t *acc) {
    int M = pl->M;
    fft_inplace(pl, re, im, -1);
    double inv = 1.0 / (double)M;
    for (int j = 0; j < M; ++j) {
        double zr = re[j] * inv, zi = im[j] * inv;
        double a = zr * pl->tw_re[j] + zi * pl->tw_im[j];  /* z * conj(tw) */
        double b = zi * pl->tw_re[j] - zr * pl->tw_im[j];
        acc[j] += f64_to_torus(a);
        acc[j + M] += f64_to_torus(b);
    }
}

/* out = a (small signed ints) * b (torus) mod X^N+1, via the FFT path (test hook) */
void orc_negacyclic_mul_fft(int32_t N, const int64_t *a, const uint64_t *b, uint64_t *out) {
    const fft_plan *pl = get_plan(N);
    int M = pl->M;
    double *buf = (double *)malloc(sizeof(double) * (size_t)N * 3);
    double *ca = buf, *ar = buf + N, *ai = ar + M, *br = buf + 2 * N, *bi = br + M;
    for (int j = 0; j < N; ++j) ca[j] = (double)a[j];
    nega_forward(pl, ca, ar, ai);
    for (int j = 0; j < N; ++j) ca[j] = (double)(int64_t)b[j];
    nega_forward(pl, ca, br, bi);
    for (int j = 0; j < M; ++j) {
        double r = ar[j] * br[j] - ai[j] * bi[j], i = ar[j] * bi[j] + ai[j] * br[j];
        ar[j] = r; ai[j] = i;
    }
    for (int j = 0; j < N; ++j) out[j] = 0;
    nega_inverse_add(pl, ar, ai, out);
    free(buf);
}

/* ------------------------------------------------------------------ */
/* Bootstrapping key                           */
/* bsk[i][t][lev][c][N]: row (t,lev) of GGSW_S(s_i); c<k mask polys, c=k body */
/* ------------------------------------------------------------------ */
void orc_bsk_gen(const orc_pbs_params *p, const uint8_t *s_small, const uint8_t *S_big,
                 uint64_t evk_seed, uint64_t *bsk) {
    int n = p->n, k = p->k, N = p->N, l = p->l_pbs, beta = p->beta_pbs;
    int64_t rows = (int64_t)n * (k + 1) * l;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t R = 0; R < rows; ++R) {
        int lev = (int)(R % l);
        int t = (int)((R / l) % (k + 1));
        int i = (int)(R / ((int64_t)l * (k + 1)));
        uint64_t *row = bsk + (size_t)R * (k + 1) * N;
        uint64_t *body = row + (size_t)k * N;
        for (int x = 0; x < N; ++x)
            body[x] = (uint64_t)orc_gaussian(evk_seed, ORC_KIND_NOISE | (ORC_PUR_BSK << 8),
                                             (uint64_t)R, (uint32_t)x, p->sigma_glwe_abs);
        for (int c = 0; c < k; ++c) {
            uint64_t *A = row + (size_t)c * N;
            for (int x = 0; x < N; ++x)
                A[x] = mask_word(evk_seed, ORC_PUR_BSK, (uint64_t)R, (int64_t)c * N + x);
            const uint8_t *S = S_big + (size_t)c * N;
            for (int y = 0; y < N; ++y) {
                if (!S[y]) continue;
                /* body += X^y * A (negacyclic) */
                for (int x = 0; x < N - y; ++x) body[x + y] += A[x];
                for (int x = N - y; x < N; ++x) body[x + y - N] -= A[x];
            }
        }
        if (s_small[i]) {
            uint64_t g = 1ULL << (64 - beta * (lev + 1));
            row[(size_t)t * N] += g; /* constant coefficient of component t */
        }
    }
}

/* ------------------------------------------------------------------ */
/* Packed encrypted inner products (leveled variant): documents are GLWE          */
/* encryptions of polynomials, the query is a GGSW of a polynomial, one external product per GLWE. */
/* Row R of `out` ([rows][k+1][N]) = GLWE_S(0) + msg_R(X) << shift_R on component comp_R, where     */
/*   mode 0 (vectors): msg_R = msgs + R*msg_stride, shift_R = shift, comp_R = k (body);            */
/*   mode 1 (GGSW of one polynomial): R = t*l + lev, msg_R = msgs, shift_R = 64 - beta*(lev+1),     */
/*                                    comp_R = t.                                                  */
/* Randomness: mask words / noise of row R come from object id (id_base + R), purpose ORC_PUR_GLWE. */
/* ------------------------------------------------------------------ */
void orc_glwe_encrypt_rows(const orc_pbs_params *p, const uint8_t *S_big, const int64_t *msgs, int64_t rows,
                           int64_t msg_stride, int32_t mode, int32_t shift, uint64_t seed, uint64_t noise_seed,
                           uint64_t id_base, uint64_t *out) {
    int k = p->k, N = p->N, l = p->l_pbs, beta = p->beta_pbs;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t R = 0; R < rows; ++R) {
        uint64_t id = id_base + (uint64_t)R;
        uint64_t *row = out + (size_t)R * (k + 1) * N;
        uint64_t *body = row + (size_t)k * N;
        for (int x = 0; x < N; ++x)
            body[x] = (uint64_t)orc_gaussian(noise_seed, ORC_KIND_NOISE | (ORC_PUR_GLWE << 8), id, (uint32_t)x,
                                             p->sigma_glwe_abs);
        for (int c = 0; c < k; ++c) {
            uint64_t *A = row + (size_t)c * N;
            for (int x = 0; x < N; ++x) A[x] = mask_word(seed, ORC_PUR_GLWE, id, (int64_t)c * N + x);
            const uint8_t *S = S_big + (size_t)c * N;
            for (int y = 0; y < N; ++y) {
                if (!S[y]) continue;
                for (int x = 0; x < N - y; ++x) body[x + y] += A[x];
                for (int x = N - y; x < N; ++x) body[x + y - N] -= A[x];
            }
        }
        const int64_t *m = mode == 0 ? msgs + (size_t)R * msg_stride : msgs;
        int sh = mode == 0 ? shift : 64 - beta * ((int)(R % l) + 1);
        int comp = mode == 0 ? k : (int)(R / l);
        uint64_t *dst = row + (size_t)comp * N;
        for (int x = 0; x < N; ++x) dst[x] += (uint64_t)m[x] << sh;
    }
}

/* out[b] = GGSW (Fourier, [t][lev][c][M] interleaved complex) external-product GLWE in[b]; [B][k+1][N] */
void orc_glwe_external_product_batch(const orc_pbs_params *p, const double *ggswf, const uint64_t *in, int64_t B,
                                     uint64_t *out) {
    int k = p->k, N = p->N, M = N / 2, l = p->l_pbs, beta = p->beta_pbs;
    const fft_plan *pl = get_plan(N);
#pragma omp parallel
    {
        double *co = (double *)malloc(sizeof(double) * (size_t)N);
        double *F = (double *)malloc(sizeof(double) * (size_t)(k + 1) * l * N);
        double *O = (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            const uint64_t *g = in + (size_t)b * (k + 1) * N;
            uint64_t *o = out + (size_t)b * (k + 1) * N;
            for (int t = 0; t <= k; ++t)
                for (int lev = 0; lev < l; ++lev) {
                    for (int x = 0; x < N; ++x) {
                        int64_t dg[16];
                        decompose(g[(size_t)t * N + x], l, beta, dg);
                        co[x] = (double)dg[lev];
                    }
                    double *f = F + ((size_t)t * l + lev) * N;
                    nega_forward(pl, co, f, f + M);
                }
            for (int c = 0; c <= k; ++c) {
                double *ore = O, *oim = O + M;
                for (int j = 0; j < M; ++j) { ore[j] = 0.0; oim[j] = 0.0; }
                for (int t = 0; t <= k; ++t)
                    for (int lev = 0; lev < l; ++lev) {
                        const double *f = F + ((size_t)t * l + lev) * N;
                        const double *gk = ggswf + ((((size_t)t) * l + lev) * (k + 1) + c) * N;
                        for (int j = 0; j < M; ++j) {
                            double gr = gk[2 * j], gi = gk[2 * j + 1];
                            ore[j] += f[j] * gr - f[j + M] * gi;
                            oim[j] += f[j] * gi + f[j + M] * gr;
                        }
                    }
                memset(o + (size_t)c * N, 0, sizeof(uint64_t) * N);
                nega_inverse_add(pl, ore, oim, o + (size_t)c * N);
            }
        }
        free(co); free(F); free(O);
    }
}

/* LWE sample extraction of coefficient `idx` of each GLWE: out [B][count][kN+1],
   out[b][q] = extract(in[b], first + q*step).  a_i = A_{idx-i} (i <= idx), -A_{N+idx-i} (i > idx). */
void orc_glwe_sample_extract(const orc_pbs_params *p, const uint64_t *in, int64_t B, int32_t first, int32_t step,
                             int32_t count, int64_t out_stride, uint64_t *out) {
    int k = p->k, N = p->N;
    for (int64_t b = 0; b < B; ++b)
        for (int q = 0; q < count; ++q) {
            int idx = first + q * step;
            const uint64_t *g = in + (size_t)b * (k + 1) * N;
            uint64_t *o = out + ((size_t)b * count + q) * out_stride;
            for (int c = 0; c < k; ++c)
                for (int i = 0; i < N; ++i)
                    o[(size_t)c * N + i] = i <= idx ? g[(size_t)c * N + idx - i] : (uint64_t)0 - g[(size_t)c * N + N + idx - i];
            o[(size_t)k * N] = g[(size_t)k * N + idx];
            for (int64_t w = (int64_t)k * N + 1; w < out_stride; ++w) o[w] = 0;
        }
}

/* bskf[i][t][lev][c][M] complex interleaved (re,im), natural bin order */
void orc_bsk_to_fourier(const orc_pbs_params *p, const uint64_t *bsk, double *bskf) {
    int N = p->N, M = N / 2;
    int64_t polys = (int64_t)p->n * (p->k + 1) * p->l_pbs * (p->k + 1);
    const fft_plan *pl = get_plan(N);
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)N * 2);
#pragma omp for schedule(static)
        for (int64_t q = 0; q < polys; ++q) {
            const uint64_t *src = bsk + (size_t)q * N;
            double *co = buf, *re = buf + N, *im = re + M;
            for (int x = 0; x < N; ++x) co[x] = (double)(int64_t)src[x];
            nega_forward(pl, co, re, im);
            double *dst = bskf + (size_t)q * N;
            for (int j = 0; j < M; ++j) { dst[2 * j] = re[j]; dst[2 * j + 1] = im[j]; }
        }
        free(buf);
    }
}

void orc_modswitch_batch(const orc_pbs_params *p, const uint64_t *in, int64_t B, int32_t *out) {
    int n = p->n, log2N2 = 0;
    while ((1 << log2N2) < 2 * p->N) ++log2N2;
    for (int64_t b = 0; b < B; ++b)
        for (int j = 0; j <= n; ++j) {
            uint64_t a = in[(size_t)b * (n + 1) + j];
            out[(size_t)b * (n + 1) + j] =
                (int32_t)((((a >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * p->N - 1));
        }
}

/* dst = X^r * src (negacyclic), r in [0, 2N) */
static void rotate_poly(int N, const uint64_t *src, int r, uint64_t *dst) {
    int neg = 0;
    if (r >= N) { r -= N; neg = 1; }
    for (int x = 0; x < N; ++x) {
        uint64_t v = (x >= r) ? src[x - r] : (uint64_t)0 - src[x - r + N];
        dst[x] = neg ? (uint64_t)0 - v : v;
    }
}

/* in: [B][n+1] under the small key.  out: [B][kN+1] under the big key.
 * luts: [n_luts][N] accumulator polynomials (already scaled / half-box rotated);
 * lut_index: [B] or NULL (=> LUT 0 for all). */
void orc_pbs_batch(const orc_pbs_params *p, const double *bskf, const uint64_t *in, int64_t B,
                   const uint64_t *luts, const int32_t *lut_index, uint64_t *out) {
    int n = p->n, k = p->k, N = p->N, M = N / 2, l = p->l_pbs, beta = p->beta_pbs;
    const fft_plan *pl = get_plan(N);
    int log2N2 = 0;
    while ((1 << log2N2) < 2 * N) ++log2N2;
#pragma omp parallel
    {
        uint64_t *acc = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(k + 1) * N);
        uint64_t *rot = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)N);
        double *co = (double *)malloc(sizeof(double) * (size_t)N);
        double *F = (double *)malloc(sizeof(double) * (size_t)(k + 1) * l * N); /* re[M], im[M] per (t,lev) */
        double *O = (double *)malloc(sizeof(double) * (size_t)N);
        int64_t *digs = (int64_t *)malloc(sizeof(int64_t) * (size_t)N * l);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            const uint64_t *ct = in + (size_t)b * (n + 1);
            const uint64_t *lut = luts + (size_t)(lut_index ? lut_index[b] : 0) * N;
            int bt = (int)((((ct[n] >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * N - 1));
            for (int c = 0; c < k; ++c) memset(acc + (size_t)c * N, 0, sizeof(uint64_t) * N);
            rotate_poly(N, lut, (2 * N - bt) % (2 * N), acc + (size_t)k * N);
            for (int i = 0; i < n; ++i) {
                int at = (int)((((ct[i] >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * N - 1));
                if (at == 0) continue;
                for (int t = 0; t <= k; ++t) {
                    uint64_t *a = acc + (size_t)t * N;
                    rotate_poly(N, a, at, rot);
                    for (int x = 0; x < N; ++x) {
                        int64_t dg[16];
                        decompose(rot[x] - a[x], l, beta, dg);
                        for (int lev = 0; lev < l; ++lev) digs[(size_t)lev * N + x] = dg[lev];
                    }
                    for (int lev = 0; lev < l; ++lev) {
                        for (int x = 0; x < N; ++x) co[x] = (double)digs[(size_t)lev * N + x];
                        double *f = F + ((size_t)t * l + lev) * N;
                        nega_forward(pl, co, f, f + M);
                    }
                }
                for (int c = 0; c <= k; ++c) {
                    double *ore = O, *oim = O + M;
                    for (int j = 0; j < M; ++j) { ore[j] = 0.0; oim[j] = 0.0; }
                    for (int t = 0; t <= k; ++t)
                        for (int lev = 0; lev < l; ++lev) {
                            const double *f = F + ((size_t)t * l + lev) * N;
                            const double *g = bskf + ((((size_t)i * (k + 1) + t) * l + lev) * (k + 1) + c) * N;
                            for (int j = 0; j < M; ++j) {
                                double gr = g[2 * j], gi = g[2 * j + 1];
                                ore[j] += f[j] * gr - f[j + M] * gi;
                                oim[j] += f[j] * gi + f[j + M] * gr;
                            }
                        }
                    nega_inverse_add(pl, ore, oim, acc + (size_t)c * N);
                }
            }
            /* sample extract coefficient 0 */
            uint64_t *o = out + (size_t)b * ((size_t)k * N + 1);
            for (int c = 0; c < k; ++c) {
                const uint64_t *A = acc + (size_t)c * N;
                o[(size_t)c * N] = A[0];
                for (int x = 1; x < N; ++x) o[(size_t)c * N + x] = (uint64_t)0 - A[N - x];
            }
            o[(size_t)k * N] = acc[(size_t)k * N];
        }
        free(acc); free(rot); free(co); free(F); free(O); free(digs);
    }
}

/* ------------------------------------------------------------------ */
/* Multi-bit blind rotation, grouping factor 2                          */
/* (Bourse, Minelli, Minihold, Paillier 2018; TFHE-rs "multi-bit PBS").  */
/* For the key-bit pair (s_a, s_b) = (s_{2i}, s_{2i+1}) the key holds    */
/* G1 = GGSW(s_a s_b), G2 = GGSW(s_a (1-s_b)), G3 = GGSW((1-s_a) s_b)   */
/* and one step is  ACC += sum_g (X^{e_g} - 1) * (G_g [.] ACC) with      */
/* e1 = a~_a + a~_b, e2 = a~_a, e3 = a~_b  (exactly one G_g encrypts 1,  */
/* or none): one gadget decomposition of ACC serves two key bits.        */
/* bsk2[i][g][t][lev][c][N], rows drawn from purpose ORC_PUR_BSK2.       */
/* ------------------------------------------------------------------ */
void orc_bsk2_gen(const orc_pbs_params *p, const uint8_t *s_small, const uint8_t *S_big,
                  uint64_t evk_seed, uint64_t *bsk2) {
    int n = p->n, k = p->k, N = p->N, l = p->l_pbs, beta = p->beta_pbs;
    int64_t rows = (int64_t)(n / 2) * 3 * (k + 1) * l;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t R = 0; R < rows; ++R) {
        int lev = (int)(R % l);
        int t = (int)((R / l) % (k + 1));
        int g = (int)((R / ((int64_t)l * (k + 1))) % 3);
        int i = (int)(R / ((int64_t)l * (k + 1) * 3));
        int sa = s_small[2 * i], sb = s_small[2 * i + 1];
        int bit = g == 0 ? (sa & sb) : (g == 1 ? (sa & (1 - sb)) : ((1 - sa) & sb));
        uint64_t *row = bsk2 + (size_t)R * (k + 1) * N;
        uint64_t *body = row + (size_t)k * N;
        for (int x = 0; x < N; ++x)
            body[x] = (uint64_t)orc_gaussian(evk_seed, ORC_KIND_NOISE | (ORC_PUR_BSK2 << 8),
                                             (uint64_t)R, (uint32_t)x, p->sigma_glwe_abs);
        for (int c = 0; c < k; ++c) {
            uint64_t *A = row + (size_t)c * N;
            for (int x = 0; x < N; ++x)
                A[x] = mask_word(evk_seed, ORC_PUR_BSK2, (uint64_t)R, (int64_t)c * N + x);
            const uint8_t *S = S_big + (size_t)c * N;
            for (int y = 0; y < N; ++y) {
                if (!S[y]) continue;
                for (int x = 0; x < N - y; ++x) body[x + y] += A[x];
                for (int x = N - y; x < N; ++x) body[x + y - N] -= A[x];
            }
        }
        if (bit) row[(size_t)t * N] += 1ULL << (64 - beta * (lev + 1));
    }
}

/* bskf2[i][g][t][lev][c][M] complex, natural bins (oracle layout; the GPU re-slices it) */
void orc_bsk2_to_fourier(const orc_pbs_params *p, const uint64_t *bsk2, double *bskf2) {
    orc_pbs_params q = *p;
    q.n = (p->n / 2) * 3;  /* same per-polynomial transform, 3 GGSWs per pair */
    orc_bsk_to_fourier(&q, bsk2, bskf2);
}

void orc_pbs_mb2_batch(const orc_pbs_params *p, const double *bskf2, const uint64_t *in, int64_t B,
                       const uint64_t *luts, const int32_t *lut_index, uint64_t *out) {
    int n = p->n, k = p->k, N = p->N, M = N / 2, l = p->l_pbs, beta = p->beta_pbs;
    const fft_plan *pl = get_plan(N);
    int log2N2 = 0;
    while ((1 << log2N2) < 2 * N) ++log2N2;
#pragma omp parallel
    {
        uint64_t *acc = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(k + 1) * N);
        uint64_t *prod = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)N);
        uint64_t *rot = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)N);
        uint64_t *delta = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(k + 1) * N);
        double *co = (double *)malloc(sizeof(double) * (size_t)N);
        double *F = (double *)malloc(sizeof(double) * (size_t)(k + 1) * l * N);
        double *O = (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            const uint64_t *ct = in + (size_t)b * (n + 1);
            const uint64_t *lut = luts + (size_t)(lut_index ? lut_index[b] : 0) * N;
            int bt = (int)((((ct[n] >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * N - 1));
            for (int c = 0; c < k; ++c) memset(acc + (size_t)c * N, 0, sizeof(uint64_t) * N);
            rotate_poly(N, lut, (2 * N - bt) % (2 * N), acc + (size_t)k * N);
            for (int i = 0; i < n / 2; ++i) {
                int ea = (int)((((ct[2 * i] >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * N - 1));
                int eb = (int)((((ct[2 * i + 1] >> (64 - log2N2 - 1)) + 1) >> 1) & (uint64_t)(2 * N - 1));
                int e[3] = {(ea + eb) % (2 * N), ea, eb};
                /* one decomposition of ACC for the pair */
                for (int t = 0; t <= k; ++t) {
                    const uint64_t *a = acc + (size_t)t * N;
                    for (int lev = 0; lev < l; ++lev) {
                        for (int x = 0; x < N; ++x) {
                            int64_t dg[16];
                            decompose(a[x], l, beta, dg);
                            co[x] = (double)dg[lev];
                        }
                        double *f = F + ((size_t)t * l + lev) * N;
                        nega_forward(pl, co, f, f + M);
                    }
                }
                memset(delta, 0, sizeof(uint64_t) * (size_t)(k + 1) * N);
                for (int g = 0; g < 3; ++g) {
                    if (e[g] == 0) continue; /* X^0 - 1 = 0 */
                    for (int c = 0; c <= k; ++c) {
                        double *ore = O, *oim = O + M;
                        for (int j = 0; j < M; ++j) { ore[j] = 0.0; oim[j] = 0.0; }
                        for (int t = 0; t <= k; ++t)
                            for (int lev = 0; lev < l; ++lev) {
                                const double *f = F + ((size_t)t * l + lev) * N;
                                const double *gk = bskf2 + (((((size_t)i * 3 + g) * (k + 1) + t) * l + lev) * (k + 1) + c) * N;
                                for (int j = 0; j < M; ++j) {
                                    double gr = gk[2 * j], gi = gk[2 * j + 1];
                                    ore[j] += f[j] * gr - f[j + M] * gi;
                                    oim[j] += f[j] * gi + f[j + M] * gr;
                                }
                            }
                        memset(prod, 0, sizeof(uint64_t) * N);
                        nega_inverse_add(pl, ore, oim, prod);          /* prod = (G_g [.] ACC)_c           */
                        rotate_poly(N, prod, e[g], rot);                /* rot  = X^{e_g} * prod            */
                        uint64_t *dl = delta + (size_t)c * N;
                        for (int x = 0; x < N; ++x) dl[x] += rot[x] - prod[x];
                    }
                }
                for (size_t x = 0; x < (size_t)(k + 1) * N; ++x) acc[x] += delta[x];
            }
            uint64_t *o = out + (size_t)b * ((size_t)k * N + 1);
            for (int c = 0; c < k; ++c) {
                const uint64_t *A = acc + (size_t)c * N;
                o[(size_t)c * N] = A[0];
                for (int x = 1; x < N; ++x) o[(size_t)c * N + x] = (uint64_t)0 - A[N - x];
            }
            o[(size_t)k * N] = acc[(size_t)k * N];
        }
        free(acc); free(prod); free(rot); free(delta); free(co); free(F); free(O);
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
