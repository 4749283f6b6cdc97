/*
 * fhe_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the arithmetic that the reference delegates to the
 * un-vendored Concrete stack (concrete-python==2.10.0 / concrete-ml==1.9.0,
 * /root/reference/requirements.txt:5,7).  The reference's own call sites for this
 * path are fhe_similarity.py:120 (compile), :151 (predict fhe="execute") and
 * batch_operations.py:233,276 (clear quantized predict).
 *
 * PARITY STATUS: "parity unpinned" for ciphertext-level values: the reference tree
 * holds no golden vectors and Concrete cannot be installed here.
 * What IS pinned: (1) Philox4x32-10 against the published Random123 known-answer
 * vectors, (2) the reference's only hot-path invariant -- decrypt(eval(encrypt(q)))
 * equals the clear quantized integer circuit exactly (test_fhe.py:56-57),
 * (3) decrypt(PBS(enc(m))) == LUT[m] for every m.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef FHE_ORACLE_H
#define FHE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RNG domain tags (counter word c3 = kind | purpose << 8) -- the spec is in DESIGN.md */
enum { ORC_KIND_SK = 1, ORC_KIND_MASK = 2, ORC_KIND_NOISE = 3 };
#define ORC_MASK_ROUNDS 7 /* Philox rounds of the public mask stream; secret streams use 10 */
enum { ORC_PUR_INPUT = 0, ORC_PUR_KSK = 1, ORC_PUR_BSK = 2, ORC_PUR_BSK2 = 3, ORC_PUR_GLWE = 4 };

typedef struct {
    int32_t n;        /* small LWE dimension */
    int32_t k;        /* GLWE dimension */
    int32_t N;        /* polynomial size (power of two) */
    int32_t l_pbs;    /* PBS decomposition levels */
    int32_t beta_pbs; /* PBS decomposition base log */
    int32_t l_ks;     /* keyswitch levels */
    int32_t beta_ks;  /* keyswitch base log */
    int32_t _pad;
    double sigma_lwe_abs;  /* std of small-key noise, in units of 2^-64 torus steps */
    double sigma_glwe_abs; /* std of GLWE noise, same units */
} orc_pbs_params;

/* ---- deterministic PRNG ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void orc_rng_block(uint64_t seed, uint32_t domain, uint64_t obj, uint32_t blk, uint32_t out[4]);
double orc_det_log(double x);
double orc_det_cos2pi_k53(uint64_t k53);
double orc_normal_from_block(const uint32_t r[4]);
int64_t orc_gaussian(uint64_t seed, uint32_t domain, uint64_t obj, uint32_t blk, double sigma_abs);

/* ---- LWE over the 64-bit torus ---- */
void orc_secret_key(uint64_t key_seed, uint32_t key_id, int64_t dim, uint8_t *s);
void orc_lwe_encrypt_batch(const uint8_t *s, int32_t n, int64_t stride, const int64_t *msgs,
                           int64_t count, int32_t shift, double sigma_abs, uint64_t enc_seed,
                           uint64_t noise_seed, uint64_t ct_base, uint32_t purpose, uint64_t *out);
void orc_lwe_phase_batch(const uint8_t *s, int32_t n, int64_t stride, const uint64_t *ct,
                         int64_t count, uint64_t *phase);
void orc_lwe_decrypt_batch(const uint8_t *s, int32_t n, int64_t stride, const uint64_t *ct,
                           int64_t count, int32_t shift, int64_t *out);
void orc_lincomb_batch(const uint64_t *ct, int64_t B, int32_t d, int32_t n, int64_t stride,
                       const int64_t *W, int32_t M, const int64_t *bias, int32_t shift,
                       uint64_t *out);

/* ---- keyswitch / PBS ---- */
void orc_ksk_gen(const orc_pbs_params *p, const uint8_t *S_big, const uint8_t *s_small,
                 uint64_t evk_seed, uint64_t *ksk);
void orc_bsk_gen(const orc_pbs_params *p, const uint8_t *s_small, const uint8_t *S_big,
                 uint64_t evk_seed, uint64_t *bsk);
void orc_bsk_to_fourier(const orc_pbs_params *p, const uint64_t *bsk, double *bskf);
void orc_keyswitch_batch(const orc_pbs_params *p, const uint64_t *ksk, const uint64_t *in,
                         int64_t B, uint64_t *out);
void orc_ksk_to_32(const orc_pbs_params *p, const uint64_t *ksk, uint32_t *ksk32);
void orc_keyswitch32_batch(const orc_pbs_params *p, const uint32_t *ksk32, const uint64_t *in,
                           int64_t B, uint64_t *out);
void orc_modswitch_batch(const orc_pbs_params *p, const uint64_t *in, int64_t B, int32_t *out);
void orc_pbs_batch(const orc_pbs_params *p, const double *bskf, const uint64_t *in, int64_t B,
                   const uint64_t *luts, const int32_t *lut_index, uint64_t *out);
void orc_bsk2_gen(const orc_pbs_params *p, const uint8_t *s_small, const uint8_t *S_big,
                  uint64_t evk_seed, uint64_t *bsk2);
void orc_bsk2_to_fourier(const orc_pbs_params *p, const uint64_t *bsk2, double *bskf2);
void orc_pbs_mb2_batch(const orc_pbs_params *p, const double *bskf2, const uint64_t *in, int64_t B,
                       const uint64_t *luts, const int32_t *lut_index, uint64_t *out);
void orc_glwe_encrypt_rows(const orc_pbs_params *p, const uint8_t *S_big, const int64_t *msgs, int64_t rows,
                           int64_t msg_stride, int32_t mode, int32_t shift, uint64_t seed, uint64_t noise_seed,
                           uint64_t id_base, uint64_t *out);
void orc_glwe_external_product_batch(const orc_pbs_params *p, const double *ggswf, const uint64_t *in, int64_t B,
                                     uint64_t *out);
void orc_glwe_sample_extract(const orc_pbs_params *p, const uint64_t *in, int64_t B, int32_t first, int32_t step,
                             int32_t count, int64_t out_stride, uint64_t *out);
void orc_negacyclic_mul_fft(int32_t N, const int64_t *a, const uint64_t *b, uint64_t *out);
int orc_num_threads(void);
void orc_set_num_threads(int t);

#ifdef __cplusplus
}
#endif
#endif
