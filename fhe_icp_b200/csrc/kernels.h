// kernels.h -- internal launcher interface between api.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fhe_b200.h"

namespace fhe {

void count_launch(int n = 1);  // api.cu: bumps the per-process launch counter

// lwe.cu
cudaError_t launch_secret_key(uint64_t key_seed, uint32_t key_id, int64_t dim, uint8_t* d_key, cudaStream_t s);
cudaError_t launch_lwe_encrypt(const uint8_t* d_key, int n, int64_t stride, const int64_t* d_msgs, int64_t count,
                               int shift, double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                               uint32_t purpose, uint64_t* d_ct, cudaStream_t s);
cudaError_t launch_lwe_encrypt_packed(const uint32_t* d_kbits, int n, int64_t stride, const int64_t* d_msgs, int64_t count,
                                      int shift, double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                                      uint32_t purpose, uint64_t* d_ct, cudaStream_t s);
cudaError_t launch_lwe_phase(const uint8_t* d_key, int n, int64_t stride, const uint64_t* d_ct, int64_t count,
                             int shift, bool decode, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_lincomb(const uint64_t* d_ct, int64_t B, int d, int n, int64_t stride, const int64_t* d_W, int M,
                           bool second_is_sum, int64_t bias0, int64_t bias1, int shift, uint64_t* d_out,
                           cudaStream_t s);
cudaError_t launch_lwe_encrypt_seeded(const uint8_t* d_key, int n, const int64_t* d_msgs, int64_t count, int shift,
                                      double sigma_abs, uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base,
                                      uint32_t purpose, uint64_t* d_bodies, cudaStream_t s);
cudaError_t launch_lwe_encrypt_seeded_float(const uint8_t* d_key, int n, const float* d_X, const float* d_query, int d,
                                            int64_t count, double scale,
                                            int64_t zp, int64_t qmin, int64_t qmax, int shift, double sigma_abs,
                                            uint64_t enc_seed, uint64_t noise_seed, uint64_t ct_base, uint32_t purpose,
                                            uint64_t* d_bodies, cudaStream_t s);
cudaError_t launch_lwe_expand_seeded(const uint64_t* d_bodies, int64_t count, int n, int64_t stride, uint64_t enc_seed,
                                     uint64_t ct_base, uint32_t purpose, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_lincomb_seeded(const uint64_t* d_bodies, int64_t B, int d, int n, int64_t stride, uint64_t enc_seed,
                                  uint64_t ct_base, uint32_t purpose, const int64_t* d_W, int M, bool second_is_sum,
                                  int64_t bias0, int64_t bias1, int shift, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_lincomb_push(const uint64_t* d_ct, int64_t B, int d, int n, int64_t stride, const int64_t* d_W, int M,
                                bool second_is_sum, int64_t bias0, int64_t bias1, int shift, const fhe_b200_push& push,
                                cudaStream_t s);
cudaError_t launch_lincomb_seeded_push(const uint64_t* d_bodies, int64_t B, int d, int n, int64_t stride,
                                       uint64_t enc_seed, uint64_t ct_base, uint32_t purpose, const int64_t* d_W, int M,
                                       bool second_is_sum, int64_t bias0, int64_t bias1, int shift,
                                       const fhe_b200_push& push, cudaStream_t s);
cudaError_t launch_peer_wait(const uint64_t* d_flags, int count, uint64_t value, uint32_t timeout_ms,
                             uint32_t* d_status, cudaStream_t s);
cudaError_t launch_peer_signal(uint64_t* const* d_flag_ptrs, int count, uint64_t value, cudaStream_t s);
cudaError_t launch_lwe_modswitch32(const uint64_t* d_in, int64_t words, uint32_t* d_out, cudaStream_t s);
cudaError_t launch_lwe_decrypt32(const uint8_t* d_key, int n, int64_t stride, const uint32_t* d_ct, int64_t count,
                                 int shift32, int64_t* d_out, cudaStream_t s);
cudaError_t launch_accumulate(uint64_t* d_acc, const uint64_t* d_x, int64_t words, cudaStream_t s);
cudaError_t launch_lwe_pair_addsub(const uint64_t* d_q, const uint64_t* d_y, int64_t B, int d, int words,
                                   int64_t in_stride, uint64_t offset, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_glwe_encrypt_rows(const fhe_b200_pbs_params& p, const uint8_t* d_S_big, const int64_t* d_msgs,
                                     int64_t rows, int64_t msg_stride, int mode, int shift, uint64_t seed, uint64_t noise_seed,
                                     uint64_t id_base, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_glwe_dot(const fhe_b200_pbs_params& p, const double* d_ggswf, const uint64_t* d_in, int64_t G,
                            uint64_t* d_out, int sm_count, cudaStream_t s);
cudaError_t launch_glwe_decrypt_coeffs(const uint8_t* d_S_big, const uint64_t* d_glwe, int64_t G, int N, int first,
                                       int step, int count, int shift, int64_t* d_out, cudaStream_t s);
cudaError_t launch_glwe_sample_extract(const uint64_t* d_glwe, int64_t G, int N, int first, int step, int count,
                                       int64_t out_stride, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_lwe_pair_add(const uint64_t* d_q, const uint64_t* d_y, int64_t B, int d, int words,
                                int64_t in_stride, uint64_t offset, uint64_t* d_out, cudaStream_t s);
cudaError_t launch_lwe_square_sum(const uint64_t* d_sq, int64_t B, int d, int words, const uint64_t* d_norm_q,
                                  const uint64_t* d_norm_y, int64_t norm_stride, int64_t out_stride, uint64_t* d_out,
                                  cudaStream_t s);
cudaError_t launch_lwe_shl_add(const uint64_t* d_in, int64_t in_stride, int64_t count, int words, int shift,
                               uint64_t offset, uint64_t* d_out, int64_t out_stride, cudaStream_t s);
cudaError_t launch_lwe_sub_plain(uint64_t* d_acc, int64_t acc_stride, const uint64_t* d_x, int64_t count, int words,
                                 uint64_t plain, cudaStream_t s);
cudaError_t launch_lwe_pair_diff_sum(const uint64_t* d_in, int64_t B, int d, int words, int64_t out_stride,
                                     uint64_t* d_out, cudaStream_t s);
cudaError_t launch_quantize(const float* d_X, int64_t count, double scale, int64_t zp, int64_t qmin, int64_t qmax,
                            int64_t* d_q, cudaStream_t s);
cudaError_t launch_pack_key(const uint8_t* d_key, int n, uint32_t* d_bits, cudaStream_t s);
cudaError_t launch_similarity_decrypt(const uint32_t* d_kbits, int n, int64_t stride, const void* d_cts, bool wire32,
                                      int64_t B, int M, int shift, int64_t zp_w, int64_t q_bias, double out_scale,
                                      int64_t out_zp, double* d_y, int64_t* d_q_y, cudaStream_t s);

// keys.cu
cudaError_t launch_ksk_gen(const fhe_b200_pbs_params& p, const uint8_t* d_S_big, const uint8_t* d_s_small,
                           uint64_t evk_seed, uint64_t* d_ksk, cudaStream_t s);
cudaError_t launch_bsk_gen(const fhe_b200_pbs_params& p, const uint8_t* d_s_small, const uint8_t* d_S_big,
                           uint64_t evk_seed, uint64_t* d_bsk, cudaStream_t s);

cudaError_t launch_bsk2_gen(const fhe_b200_pbs_params& p, const uint8_t* d_s_small, const uint8_t* d_S_big,
                            uint64_t evk_seed, uint64_t* d_bsk2, cudaStream_t s);

// keyswitch.cu
cudaError_t launch_keyswitch(const fhe_b200_pbs_params& p, const uint64_t* d_ksk, const uint64_t* d_in, int64_t B,
                             uint64_t* d_out, cudaStream_t s);

cudaError_t launch_ksk_to_32(const fhe_b200_pbs_params& p, const uint64_t* d_ksk, uint32_t* d_ksk32, cudaStream_t s);
cudaError_t launch_keyswitch32(const fhe_b200_pbs_params& p, const uint32_t* d_ksk32, const uint64_t* d_in, int64_t B,
                               uint32_t* d_acc32, uint64_t* d_out, cudaStream_t s);

// ks_mma.cu -- the 32-bit keyswitch as an int8 tensor-core contraction (tcgen05.mma.kind::i8)
bool keyswitch_mma_supported(const fhe_b200_pbs_params& p);
size_t keyswitch_mma_key_bytes(const fhe_b200_pbs_params& p);
size_t keyswitch_mma_workspace_bytes(const fhe_b200_pbs_params& p, int64_t B);
cudaError_t launch_ksk32_to_mma(const fhe_b200_pbs_params& p, const uint32_t* d_ksk32, uint8_t* d_tiles, cudaStream_t s);
cudaError_t launch_keyswitch_mma(const fhe_b200_pbs_params& p, const uint8_t* d_tiles, const uint64_t* d_in, int64_t B,
                                 int8_t* d_work, uint64_t* d_out, cudaStream_t s);

// pbs.cu
cudaError_t launch_bsk_to_fourier(const fhe_b200_pbs_params& p, const uint64_t* d_bsk, double* d_bskf,
                                  cudaStream_t s);
cudaError_t launch_pbs(const fhe_b200_pbs_params& p, const double* d_bskf, const uint64_t* d_in, int64_t B,
                       const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, int sm_count,
                       cudaStream_t s);
cudaError_t launch_bsk2_to_fourier(const fhe_b200_pbs_params& p, const uint64_t* d_bsk2, double* d_bskf2, cudaStream_t s);
cudaError_t launch_pbs_mb2(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                           const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, int sm_count,
                           cudaStream_t s);
bool pbs_params_supported(const fhe_b200_pbs_params& p, const char** why);
cudaError_t pbs_tables(const void** tables);

// pbs_wide.cu -- four-warps-per-polynomial multi-bit blind rotation, one ciphertext per CTA (256 threads, accumulator
// and twiddles in registers): the latency kernel, dispatched by launch_pbs_mb2 up to one ciphertext per SM.
cudaError_t launch_pbs_mb2_wide(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                                const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, cudaStream_t s);

// one ciphertext on a cluster of two CTAs (polynomial t and key column t on CTA t, spectra exchanged through distributed
// shared memory): the lowest latency, for batches up to half the SM count
cudaError_t launch_pbs_mb2_pair(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                                const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, cudaStream_t s);

// probe.cu
cudaError_t probe_fp64(int sm_count, double* tflops, cudaStream_t s);

}  // namespace fhe
