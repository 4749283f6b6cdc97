/*
 * fhe_b200.h -- C-ABI of the B200-native encrypted-similarity engine.
 *
 * This is the drop-in boundary for the reference's FHE hot path.  The reference
 * (shipstone-labs/fhe-icp) has no native interface of its own: it reaches the
 * arithmetic through Concrete-ML's Python estimator.  Each entry point below cites
 * the reference call it replaces (paths are relative to the reference tree).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (FHE_B200_OK) or an error code; the message for the
 *     calling thread is available from fhe_b200_last_error(); nothing throws.
 *   - `d_*` pointers are DEVICE pointers owned by the caller (e.g. torch tensors);
 *     `h_*` pointers are HOST pointers.  `stream` is a cudaStream_t passed as void*.
 *   - an LWE ciphertext is `stride` little-endian u64 words: n mask words, the body at
 *     index n, zero padding up to `stride` (stride even, >= n+1).
 *   - there is no CPU fallback: without a CUDA device ctx_create fails.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHE_B200_ABI_VERSION 2

enum {
    FHE_B200_OK = 0,
    FHE_B200_ERR_INVALID = 1, /* bad argument */
    FHE_B200_ERR_CUDA = 2,    /* CUDA runtime error (see last_error) */
    FHE_B200_ERR_NO_DEVICE = 3,
    FHE_B200_ERR_STATE = 4    /* object used before it was initialised */
};

/* RNG stream tags (DESIGN.md "Deterministic randomness").  Philox4x32 counter = (block, object_lo, object_hi,
 * kind | purpose << 8), key = seed.  Secret streams (SK, NOISE): 10 rounds; public MASK stream: FHE_B200_MASK_ROUNDS. */
#define FHE_B200_MASK_ROUNDS 7
enum { FHE_B200_KIND_SK = 1, FHE_B200_KIND_MASK = 2, FHE_B200_KIND_NOISE = 3 };
enum { FHE_B200_PUR_INPUT = 0, FHE_B200_PUR_KSK = 1, FHE_B200_PUR_BSK = 2, FHE_B200_PUR_BSK2 = 3, FHE_B200_PUR_GLWE = 4 };

typedef struct fhe_b200_ctx fhe_b200_ctx;
typedef struct fhe_b200_similarity fhe_b200_similarity;

/* TFHE parameter set for keyswitch + programmable bootstrap.  The reference circuit
 * fixes none (its compiled circuit has no table lookup, fhe_similarity.py:88-90);
 * these are the engine's own, stated in DESIGN.md. */
typedef struct {
    int32_t n;        /* small LWE dimension */
    int32_t k;        /* GLWE dimension */
    int32_t N;        /* polynomial size, power of two */
    int32_t l_pbs;    /* PBS gadget levels */
    int32_t beta_pbs; /* PBS gadget base log */
    int32_t l_ks;     /* keyswitch levels */
    int32_t beta_ks;  /* keyswitch base log */
    int32_t _pad;
    double sigma_lwe_abs;  /* noise std in 2^-64 torus steps */
    double sigma_glwe_abs;
} fhe_b200_pbs_params;

/* Quantized linear model + LWE parameters of one compiled similarity circuit.
 * Mirrors what `LinearRegression(n_bits).fit().compile()` fixes in the reference
 * (fhe_similarity.py:88-94,120). */
typedef struct {
    int32_t d;          /* input features (128 at the reference's configs) */
    int32_t n_bits;     /* input quantization bits */
    int32_t n;          /* LWE dimension */
    int32_t stride;     /* u64 words per ciphertext row */
    int32_t shift;      /* log2(Delta): message m is encoded as m << shift */
    int32_t two_outputs;/* 1: also return sum_j ct_j (needed when zp_w != 0) */
    double sigma_abs;   /* fresh-encryption noise std, 2^-64 torus steps */
    double x_scale;     /* input quantizer */
    int64_t x_zero_point;
    int64_t x_offset;   /* 2^(n_bits-1) for signed inputs */
    int64_t w_zero_point;
    int64_t q_bias;
    double out_scale;   /* y = out_scale * (q_y - out_zero_point) */
    int64_t out_zero_point;
    uint64_t key_seed;  /* CLIENT SECRET: secret key seed (key_id 2) */
    uint64_t noise_seed;/* CLIENT SECRET: seed of the encryption error terms.  Never equal to, or derivable from, the
                         * public mask seed (`enc_seed`) that travels with seeded ciphertexts: whoever can regenerate
                         * the errors learns <a,s> + Delta*m exactly and solves for the key.  Both secrets are ignored
                         * (and zeroed) by fhe_b200_similarity_create_evaluator. */
} fhe_b200_similarity_spec;

/* ---- context ---------------------------------------------------------------- */
int fhe_b200_abi_version(void);
const char *fhe_b200_last_error(void);
int fhe_b200_ctx_create(int device, fhe_b200_ctx **ctx);
int fhe_b200_ctx_destroy(fhe_b200_ctx *ctx);
int fhe_b200_device_info(fhe_b200_ctx *ctx, int32_t *sm_count, int32_t *cc_major,
                         int32_t *cc_minor, uint64_t *total_mem);
/* number of kernels this library has launched since ctx creation (bench.py's gpu_launches) */
uint64_t fhe_b200_launch_count(fhe_b200_ctx *ctx);
/* measured FP64 FMA peak of the device in TFLOP/s (denominator of the PBS roofline) */
int fhe_b200_probe_fp64(fhe_b200_ctx *ctx, double *tflops);

/* ---- client side: keys, encrypt, decrypt --------------------------------------
 * replaces the lazy keygen / encrypt / decrypt inside
 * `fhe_circuit.encrypt_run_decrypt` reached from fhe_similarity.py:151 */
int fhe_b200_secret_key(fhe_b200_ctx *ctx, uint64_t key_seed, uint32_t key_id, int64_t dim,
                        uint8_t *d_key, void *stream);
/* Randomness of ciphertext number ct_base + i: the MASK is a public function of (enc_seed, purpose, id) -- it may be
 * regenerated by anyone (seeded ciphertexts rely on that); the ERROR term is drawn from (noise_seed, purpose, id) and
 * noise_seed is a client secret.  An id must never be reused under the same (enc_seed, noise_seed): two ciphertexts
 * with equal mask and error differ exactly by Delta*(m1 - m2). */
int fhe_b200_lwe_encrypt(fhe_b200_ctx *ctx, const uint8_t *d_key, int32_t n, int64_t stride,
                         const int64_t *d_msgs, int64_t count, int32_t shift, double sigma_abs,
                         uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base, uint32_t purpose,
                         uint64_t *d_ct, void *stream);
int fhe_b200_lwe_phase(fhe_b200_ctx *ctx, const uint8_t *d_key, int32_t n, int64_t stride,
                       const uint64_t *d_ct, int64_t count, uint64_t *d_phase, void *stream);
int fhe_b200_lwe_decrypt(fhe_b200_ctx *ctx, const uint8_t *d_key, int32_t n, int64_t stride,
                         const uint64_t *d_ct, int64_t count, int32_t shift, int64_t *d_msgs,
                         void *stream);

/* ---- server side: the encrypted dot product -----------------------------------
 * out[b][m][:] = sum_j W[m][j] * ct[b][j][:]  (+ bias[m] << shift on the body)
 * replaces the `run` stage of predict(..., fhe="execute"), fhe_similarity.py:151.
 * d_ct [B][d][stride], d_W [M][d] (M = 1 or 2), h_bias [M] or NULL, d_out [B][M][stride]. */
int fhe_b200_lincomb(fhe_b200_ctx *ctx, const uint64_t *d_ct, int64_t B, int32_t d, int32_t n,
                     int64_t stride, const int64_t *d_W, int32_t M, const int64_t *h_bias,
                     int32_t shift, uint64_t *d_out, void *stream);
/* ---- seeded (compressed) ciphertexts ----------------------------------------------
 * A fresh ciphertext's mask is a function of (enc_seed, purpose, ct_base + index); the seeded form
 * keeps only the 8-byte body.  encrypt_seeded writes d_bodies[count]; expand_seeded materialises the
 * identical [count][stride] rows fhe_b200_lwe_encrypt would have written; lincomb_seeded evaluates
 * the encrypted dot product of B rows of d seeded ciphertexts (ids ct_base + b*d + j) regenerating
 * the masks on the fly.  Bit-identical to the materialised path; 1 KB instead of 1.46 MB per row. */
int fhe_b200_lwe_encrypt_seeded(fhe_b200_ctx *ctx, const uint8_t *d_key, int32_t n, const int64_t *d_msgs,
                                int64_t count, int32_t shift, double sigma_abs, uint64_t enc_seed,
                                uint64_t noise_seed, uint64_t ct_base, uint32_t purpose, uint64_t *d_bodies,
                                void *stream);
int fhe_b200_lwe_expand_seeded(fhe_b200_ctx *ctx, const uint64_t *d_bodies, int64_t count, int32_t n,
                               int64_t stride, uint64_t enc_seed, uint64_t ct_base, uint32_t purpose,
                               uint64_t *d_ct, void *stream);
int fhe_b200_lincomb_seeded(fhe_b200_ctx *ctx, const uint64_t *d_bodies, int64_t B, int32_t d, int32_t n,
                            int64_t stride, uint64_t enc_seed, uint64_t ct_base, uint32_t purpose,
                            const int64_t *d_W, int32_t M, const int64_t *h_bias, int32_t shift,
                            uint64_t *d_out, void *stream);
/* 32-bit wire form of finished ciphertexts (modulus switch 2^64 -> 2^32, same row stride in words):
 * halves the bytes gathered to the decrypting client; the added noise is stated in DESIGN.md. */
int fhe_b200_lwe_modswitch32(fhe_b200_ctx *ctx, const uint64_t *d_ct, int64_t count, int64_t stride,
                             uint32_t *d_ct32, void *stream);
/* ciphertext accumulation: d_acc[i] += d_x[i] over `words` u64 words (wrapping) */
int fhe_b200_accumulate(fhe_b200_ctx *ctx, uint64_t *d_acc, const uint64_t *d_x, int64_t words,
                        void *stream);

/* ---- encrypted x encrypted comparison glue -----------------------
 * Replaces the CLEAR element-wise product emb1 * emb2 of batch_operations.py:226,273 when both
 * vectors are encrypted: x*y = floor((x+y)^2/4) - floor((x-y)^2/4), two table lookups (PBS) per
 * dimension.  pair_addsub builds the PBS inputs from the query's d ciphertexts d_q [d][in_stride]
 * and the documents' d_y [B][d][in_stride] (the first `words` words of each row are the ciphertext):
 * d_out [B][d][2][words] = (q+y+offset, q-y+offset), `offset` added to the body (word words-1).  pair_diff_sum folds the bootstrapped squares back into one
 * score ciphertext per document: d_out [B][out_stride] = sum_j (d_in[b][j][0] - d_in[b][j][1]), words
 * words..out_stride-1 of each row zeroed (an even out_stride makes the rows decryptable in place). */
int fhe_b200_lwe_pair_addsub(fhe_b200_ctx *ctx, const uint64_t *d_q, const uint64_t *d_y, int64_t B, int32_t d,
                             int32_t words, int64_t in_stride, uint64_t offset, uint64_t *d_out, void *stream);
int fhe_b200_lwe_pair_diff_sum(fhe_b200_ctx *ctx, const uint64_t *d_in, int64_t B, int32_t d, int32_t words,
                               int64_t out_stride, uint64_t *d_out, void *stream);

/* ---- packed encrypted inner products: GLWE x GGSW (leveled; no bootstrap) ------------
 * The fast form of the both-encrypted comparison (same reference anchor: the clear product of
 * batch_operations.py:226,273).  Documents are packed N/slot per GLWE ciphertext (document b of a
 * group in coefficients slot*b ..), the query is a GGSW encryption of Q(X) = sum_j x_j X^(-j);
 * coefficient slot*b of GGSW(Q) [.] GLWE is sum_j x_j*y_{b,j}.  One external product per GLWE.
 *  glwe_encrypt_rows: row R of d_out [rows][k+1][N] = GLWE_S(0) + (msg_R << shift_R) on component comp_R;
 *    mode 0: msg_R = d_msgs + R*msg_stride (N coefficients), shift_R = shift, comp_R = k;
 *    mode 1 (GGSW of one polynomial, rows = (k+1)*l): R = t*l+lev, msg_R = d_msgs, shift_R = 64-beta*(lev+1),
 *    comp_R = t.  Row R draws its mask from (seed, id_base + R) -- public -- and its error from (noise_seed,
 *    id_base + R) -- client secret; purpose FHE_B200_PUR_GLWE.
 *  The GGSW goes to the Fourier domain with fhe_b200_bsk_to_fourier and a parameter copy with n = 1
 *    (layout [t][lev][c][N/2] complex).
 *  glwe_ggsw_dot: d_out [G][k+1][N] = GGSW [.] d_in[g] for all g (k = 1, l_pbs = 2).
 *  glwe_decrypt_coeffs (client): d_msgs [G][count] = round(phase of coefficient first+q*step / 2^shift).
 *  glwe_sample_extract: d_out [(g*count+q)][out_stride] = LWE (big key) of coefficient first+q*step. */
int fhe_b200_glwe_encrypt_rows(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_S_big,
                               const int64_t *d_msgs, int64_t rows, int64_t msg_stride, int32_t mode,
                               int32_t shift, uint64_t seed, uint64_t noise_seed, uint64_t id_base, uint64_t *d_out,
                               void *stream);
int fhe_b200_glwe_ggsw_dot(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const double *d_ggswf,
                           const uint64_t *d_in, int64_t G, uint64_t *d_out, void *stream);
int fhe_b200_glwe_decrypt_coeffs(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_S_big,
                                 const uint64_t *d_glwe, int64_t G, int32_t first, int32_t step, int32_t count,
                                 int32_t shift, int64_t *d_msgs, void *stream);
int fhe_b200_glwe_sample_extract(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint64_t *d_glwe,
                                 int64_t G, int32_t first, int32_t step, int32_t count, int64_t out_stride,
                                 uint64_t *d_out, void *stream);

/* One bootstrap per dimension instead of two when each party also sends an encryption of its own
 * squared norm (big key, same scale as the bootstrapped squares):
 *   2 * sum_j x_j*y_j = sum_j (x_j+y_j)^2 - sum_j x_j^2 - sum_j y_j^2.
 * pair_add: d_out [B][d][words] = q + y + offset.  square_sum: d_out [B][out_stride] =
 * sum_j d_sq[b][j] - d_norm_q - d_norm_y[b] (norm rows norm_stride words apart), padding zeroed. */
int fhe_b200_lwe_pair_add(fhe_b200_ctx *ctx, const uint64_t *d_q, const uint64_t *d_y, int64_t B, int32_t d,
                          int32_t words, int64_t in_stride, uint64_t offset, uint64_t *d_out, void *stream);
int fhe_b200_lwe_square_sum(fhe_b200_ctx *ctx, const uint64_t *d_sq, int64_t B, int32_t d, int32_t words,
                            const uint64_t *d_norm_q, const uint64_t *d_norm_y, int64_t norm_stride,
                            int64_t out_stride, uint64_t *d_out, void *stream);

/* ---- exact encrypted threshold glue -------------------------------
 * Replaces the CLEAR test `similarity >= min_similarity` of batch_operations.py:278 (and the score
 * buckets of fhe_cli.py:169-176) by the sign of (score - T) computed under encryption: LSB-first bit
 * extraction, one keyswitch + PBS per bit.  shl_add: d_out [count][out_stride] = (row << shift) with
 * `offset` added to the body (word words-1), words beyond zeroed; sub_plain: d_acc row -= d_x row, body additionally -= plain. */
int fhe_b200_lwe_shl_add(fhe_b200_ctx *ctx, const uint64_t *d_in, int64_t in_stride, int64_t count, int32_t words,
                         int32_t shift, uint64_t offset, uint64_t *d_out, int64_t out_stride, void *stream);
int fhe_b200_lwe_sub_plain(fhe_b200_ctx *ctx, uint64_t *d_acc, int64_t acc_stride, const uint64_t *d_x,
                           int64_t count, int32_t words, uint64_t plain, void *stream);

/* ---- keyswitch + programmable bootstrap ---------------------------------------
 * north-star primitives; not exercised by the reference's compiled circuit. */
int fhe_b200_ksk_gen(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_S_big,
                     const uint8_t *d_s_small, uint64_t evk_seed, uint64_t *d_ksk, void *stream);
int fhe_b200_bsk_gen(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_s_small,
                     const uint8_t *d_S_big, uint64_t evk_seed, uint64_t *d_bsk, void *stream);
uint64_t fhe_b200_ksk_words(const fhe_b200_pbs_params *p);
uint64_t fhe_b200_bsk_words(const fhe_b200_pbs_params *p);
int fhe_b200_bsk_to_fourier(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p,
                            const uint64_t *d_bsk, double *d_bskf, void *stream);
int fhe_b200_keyswitch(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint64_t *d_ksk,
                       const uint64_t *d_in, int64_t B, uint64_t *d_out, void *stream);
/* Multi-bit blind rotation, grouping factor 2 (two key bits per CMux, one gadget decomposition and one
 * FFT round trip per pair; n even, k = 1, l_pbs <= 2, and 2*beta_pbs <= 31 when l_pbs = 2).  bsk2 standard
 * domain: [n/2][3][k+1][l][k+1][N] u64; Fourier domain: [n/2][32 blocks][3][k+1][l][k+1][32] complex f64
 * (by frequency block so that the kernel streams it through a small shared-memory ring).  Same inputs / outputs as fhe_b200_pbs. */
uint64_t fhe_b200_bsk2_words(const fhe_b200_pbs_params *p);
int fhe_b200_bsk2_gen(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_s_small,
                      const uint8_t *d_S_big, uint64_t evk_seed, uint64_t *d_bsk2, void *stream);
int fhe_b200_bsk2_to_fourier(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint64_t *d_bsk2,
                             double *d_bskf2, void *stream);
int fhe_b200_pbs_mb2(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const double *d_bskf2,
                     const uint64_t *d_in, int64_t B, const uint64_t *d_luts,
                     const int32_t *d_lut_index, uint64_t *d_out, void *stream);
/* The same multi-bit blind rotation with FOUR warps per polynomial (8 points per thread, accumulator and twiddles in
 * registers; pbs_wide.cu): one ciphertext per CTA of 256 threads -- the latency kernel.  fhe_b200_pbs_mb2 itself runs
 * it for what is left of a batch after full waves of 4 x SMs ciphertexts (up to 3 x SMs); this entry point forces it for the tests and the batch sweep.
 * Same key, inputs and outputs as fhe_b200_pbs_mb2; k = 1, l_pbs = 1, n even. */
int fhe_b200_pbs_mb2_wide(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const double *d_bskf2,
                          const uint64_t *d_in, int64_t B, const uint64_t *d_luts,
                          const int32_t *d_lut_index, uint64_t *d_out, void *stream);
/* The same blind rotation with ONE ciphertext on a cluster of two CTAs (two SMs): CTA t owns polynomial t and streams
 * column t of the key, the spectra cross through distributed shared memory (pbs_wide.cu).  The lowest latency;
 * fhe_b200_pbs_mb2 runs it for batches up to half the SM count.  Same key, inputs and outputs as fhe_b200_pbs_mb2 (the
 * key is re-sliced by output column into stream-ordered scratch at every call). */
int fhe_b200_pbs_mb2_pair(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const double *d_bskf2,
                          const uint64_t *d_in, int64_t B, const uint64_t *d_luts,
                          const int32_t *d_lut_index, uint64_t *d_out, void *stream);
/* 32-bit keyswitch ("KS32"): key rounded to the top 32 torus bits, u32 accumulation; d_scratch32 holds
 * B*(n+1) u32 words; the result is written as u64 words with the low half zero. */
int fhe_b200_ksk_to_32(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint64_t *d_ksk,
                       uint32_t *d_ksk32, void *stream);
int fhe_b200_keyswitch32(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint32_t *d_ksk32,
                         const uint64_t *d_in, int64_t B, uint32_t *d_scratch32, uint64_t *d_out,
                         void *stream);
/* The same 32-bit keyswitch as a dense int8 contraction on the tensor cores (tcgen05.mma.kind::i8, s8 digits x
 * u8 key bytes -> s32 in tensor memory; ks_mma.cu).  Bit-identical to fhe_b200_keyswitch32.  ksk_to_mma lays
 * the 32-bit key out in MMA blocks (ksk_mma_bytes bytes, once per key); keyswitch_mma needs a per-call
 * workspace of keyswitch_mma_workspace_bytes(B) for the digit matrix.  Both sizes are 0 for parameter sets the
 * kernel does not cover (kN % 128 != 0, beta_ks > 8, or l_ks*kN*2^(beta_ks-1)*255 >= 2^31). */
uint64_t fhe_b200_ksk_mma_bytes(const fhe_b200_pbs_params *p);
uint64_t fhe_b200_keyswitch_mma_workspace_bytes(const fhe_b200_pbs_params *p, int64_t B);
int fhe_b200_ksk_to_mma(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint32_t *d_ksk32,
                        uint8_t *d_key_mma, void *stream);
int fhe_b200_keyswitch_mma(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const uint8_t *d_key_mma,
                           const uint64_t *d_in, int64_t B, int8_t *d_work, uint64_t *d_out, void *stream);
int fhe_b200_pbs(fhe_b200_ctx *ctx, const fhe_b200_pbs_params *p, const double *d_bskf,
                 const uint64_t *d_in, int64_t B, const uint64_t *d_luts,
                 const int32_t *d_lut_index, uint64_t *d_out, void *stream);

/* ---- the reference-facing call with HOST buffers ------------------------------
 * FHESimilarityModel.predict_encrypted(X) (fhe_similarity.py:142-160): quantize ->
 * encrypt d ciphertexts per row -> encrypted dot product -> decrypt -> dequantize,
 * every stage a kernel of this library, ciphertexts materialised in HBM between them. */
int fhe_b200_similarity_create(fhe_b200_ctx *ctx, const fhe_b200_similarity_spec *spec,
                               const int64_t *h_q_w, fhe_b200_similarity **sim);
/* Server side: the same model WITHOUT any secret material (key_seed / noise_seed of the spec are ignored and zeroed).
 * run / run_seeded / run_push work on it; encrypt, decrypt and predict_host return FHE_B200_ERR_INVALID.  This is what
 * the non-client ranks of a sharded search hold. */
int fhe_b200_similarity_create_evaluator(fhe_b200_ctx *ctx, const fhe_b200_similarity_spec *spec,
                                         const int64_t *h_q_w, fhe_b200_similarity **sim);
/* 1 if the 32-bit wire form of the scores (lwe_modswitch32 / run_push / decrypt32) keeps the decoding failure
 * probability at 2^-40 for this model: z * sqrt(sigma_out^2 + (n/2+1) * 2^64/12) < Delta/2; else 0 and those calls fail. */
int fhe_b200_similarity_wire32_supported(const fhe_b200_similarity *sim);
int fhe_b200_similarity_destroy(fhe_b200_similarity *sim);
int fhe_b200_similarity_predict_host(fhe_b200_similarity *sim, const float *h_X, int64_t B,
                                     uint64_t enc_seed, uint64_t ct_base, double *h_y,
                                     int64_t *h_q_y);
/* stage-wise variants on device buffers (client encrypt / server run / client decrypt) */
int fhe_b200_similarity_encrypt(fhe_b200_similarity *sim, const float *d_X, int64_t B,
                                uint64_t enc_seed, uint64_t ct_base, uint64_t *d_ct, void *stream);
int fhe_b200_similarity_run(fhe_b200_similarity *sim, const uint64_t *d_ct, int64_t B,
                            uint64_t *d_out, void *stream);
int fhe_b200_similarity_decrypt(fhe_b200_similarity *sim, const uint64_t *d_out, int64_t B,
                                double *d_y, int64_t *d_q_y, void *stream);
/* seeded variants: d_bodies [B][d] u64, ciphertext ids ct_base + b*d + j */
int fhe_b200_similarity_encrypt_seeded(fhe_b200_similarity *sim, const float *d_X, int64_t B,
                                       uint64_t enc_seed, uint64_t ct_base, uint64_t *d_bodies, void *stream);
/* Page-locked host memory for the host-buffer entry points.  Rows handed to predict_host[_seeded] from such a
 * buffer (or from any cudaHostAlloc / cudaHostRegister'ed range, e.g. a torch pinned tensor) are uploaded from where
 * they are; pageable rows are first staged through the model's own pinned buffer.  The reference passes plain numpy
 * arrays (fhe_similarity.py:151); this is the buffer a caller allocates when it wants the upload off its critical path. */
int fhe_b200_host_alloc(fhe_b200_ctx *ctx, uint64_t bytes, void **h_ptr);
int fhe_b200_host_free(fhe_b200_ctx *ctx, void *h_ptr);   /* ctx may be NULL */

/* same with the clear product fused in: feature (b, j) = d_query[j] * d_docs[b*d + j] in IEEE float32, the array the
 * reference builds on the host before every call (emb1 * emb2, /root/reference/batch_operations.py:226,273) */
int fhe_b200_similarity_encrypt_seeded_products(fhe_b200_similarity *sim, const float *d_query, const float *d_docs,
                                                int64_t B, uint64_t enc_seed, uint64_t ct_base, uint64_t *d_bodies,
                                                void *stream);
int fhe_b200_similarity_run_seeded(fhe_b200_similarity *sim, const uint64_t *d_bodies, int64_t B,
                                   uint64_t enc_seed, uint64_t ct_base, uint64_t *d_out, void *stream);
int fhe_b200_similarity_predict_host_seeded(fhe_b200_similarity *sim, const float *h_X, int64_t B,
                                            uint64_t enc_seed, uint64_t ct_base, double *h_y, int64_t *h_q_y);
/* same, for scores received in the 32-bit wire form */
int fhe_b200_similarity_decrypt32(fhe_b200_similarity *sim, const uint32_t *d_out32, int64_t B,
                                  double *d_y, int64_t *d_q_y, void *stream);

/* ---- multi-GPU search: scores pushed into the client GPU's memory ---------------------------------
 * replaces the per-document result collection of BatchProcessor.search_similar
 * (batch_operations.py:264-284) when the collection is sharded over the GPUs of one box: instead of
 * writing its encrypted scores locally and handing them to a collective, the dot-product kernel of
 * every rank stores them, in the 32-bit wire form, directly into the CLIENT rank's "score board" over
 * NVLink (a cudaIpc peer mapping) and publishes an arrival flag; the client waits for the flags,
 * decrypts the board and returns a credit per slot.  No host synchronisation, no collective call.
 * Every in-stream wait is bounded (timeout_ms): a dead peer sets a status word, it never hangs the GPU.
 *
 * peer_alloc: zero-filled device allocation that other processes of the box may map; `handle` receives
 * the 64-byte cudaIpcMemHandle to pass to them (any byte transport).  peer_open maps another process's
 * allocation (peer access is enabled on demand); peer_close unmaps; peer_free releases an allocation. */
#define FHE_B200_IPC_HANDLE_BYTES 64
int fhe_b200_peer_alloc(fhe_b200_ctx *ctx, uint64_t bytes, void **d_ptr, uint8_t *handle);
int fhe_b200_peer_open(fhe_b200_ctx *ctx, const uint8_t *handle, void **d_ptr);
int fhe_b200_peer_close(fhe_b200_ctx *ctx, void *d_ptr);
int fhe_b200_peer_free(fhe_b200_ctx *ctx, void *d_ptr);

/* Destination of one pushed evaluation.  d_board32 / d_arrive point into the client's allocation
 * (mapped with peer_open, or local on the client rank itself); d_counter is local to the caller.
 * Flow control is the caller's: before a slot is overwritten, order the launch behind
 * fhe_b200_peer_wait on the credit flag the client signals after decrypting the slot. */
typedef struct {
    uint32_t *d_board32; /* this rank's rows of the slot: [B][M][stride] u32 */
    uint64_t *d_arrive;  /* arrival flag of (slot, rank): set to `step` when every row is written */
    uint64_t step;       /* monotonic, > 0 */
    uint32_t *d_counter; /* local u32, zero before the first call; the kernel leaves it zero */
} fhe_b200_push;

/* fhe_b200_similarity_run / _run_seeded with the result pushed to the client's score board */
int fhe_b200_similarity_run_push(fhe_b200_similarity *sim, const uint64_t *d_ct, int64_t B,
                                 const fhe_b200_push *push, void *stream);
int fhe_b200_similarity_run_seeded_push(fhe_b200_similarity *sim, const uint64_t *d_bodies, int64_t B,
                                        uint64_t enc_seed, uint64_t ct_base, const fhe_b200_push *push,
                                        void *stream);
/* client: block the stream until d_flags[i] >= value for every i < count (bounded by timeout_ms;
 * on time-out *d_status is set to 1 and the stream continues).  Use it for flags that ANOTHER GPU
 * writes; order work of the same GPU with events -- two kernels of one GPU are not guaranteed to run
 * concurrently, so a kernel must not spin on a flag a later launch of the same GPU would set. */
int fhe_b200_peer_wait(fhe_b200_ctx *ctx, const uint64_t *d_flags, int32_t count, uint64_t value,
                       uint32_t timeout_ms, uint32_t *d_status, void *stream);
/* stream-ordered release store of `value` to each of the `count` flags whose (peer or local) addresses
 * are listed in the DEVICE array d_flag_ptrs: the client's credits, or the arrival of an empty shard */
int fhe_b200_peer_signal(fhe_b200_ctx *ctx, uint64_t *const *d_flag_ptrs, int32_t count, uint64_t value,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
