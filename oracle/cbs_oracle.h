/*
 * cbs_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's server-side AES-128 transciphering
 * path (code-perspective/temp-fhe-transciphering, Rust crate `auto-base-conv`
 * a.k.a. cbs_lib + the two server stage binaries), parameter set AES_TIGHT
 * (submission/cbs_lib/src/aes_instances.rs:76-97).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this library, and only as the CHECKER.  The product
 * (libcbs_b200) never links it.
 *
 * Third-party arithmetic the reference takes from crates that are not under
 * /root/reference (tfhe 0.5.4, concrete-fft 0.4.1; submission/Cargo.lock) is
 * restated from its published algorithm: negacyclic "folded + twisted" f64 FFT
 * of size N/2, signed balanced decomposition, external product, modulus switch,
 * sample extraction.  Intermediate parity with tfhe is therefore statistical
 * (same algorithm, different floating-point summation order); end-to-end
 * parity is pinned against the reference's own prebuilt binaries by
 * oracle/pin_against_reference.py (decrypted bits + output-noise statistics;
 * results committed in tests/golden/reference_pin.json).
 *
 * Layout conventions (all little-endian u64, wrapping arithmetic):
 *   LWE(n)      : [a_0..a_{n-1}, b]
 *   GLWE(k,N)   : [poly 0..k-1 mask][body], each poly N coefficients
 *   GGSW std    : [level 1..l][row 0..k][poly 0..k][N]      (level 1 = coarsest)
 *   GGSW Fourier: [level][row][poly][N/2] complex (re,im interleaved doubles)
 */
#ifndef CBS_ORACLE_H
#define CBS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* AES_TIGHT, submission/cbs_lib/src/aes_instances.rs:76-97 */
#define ORC_LWE_N      768
#define ORC_N          1024
#define ORC_K          2
#define ORC_BIG_N      (ORC_K * ORC_N)          /* 2048 */
#define ORC_PBS_BASE_LOG 23
#define ORC_PBS_LEVEL    1
#define ORC_KS_BASE_LOG  4
#define ORC_KS_LEVEL     3
#define ORC_KS_N         256
#define ORC_KS_IN_K      (ORC_BIG_N / ORC_KS_N)  /* 8 */
#define ORC_KS_OUT_K     (ORC_LWE_N / ORC_KS_N)  /* 3 */
#define ORC_AUTO_BASE_LOG 13
#define ORC_AUTO_LEVEL    3
#define ORC_AUTO_SPLIT    41
#define ORC_SS_BASE_LOG   17
#define ORC_SS_LEVEL      2
#define ORC_CBS_BASE_LOG  2
#define ORC_CBS_LEVEL     7
#define ORC_LOG_LUT_COUNT 3
#define ORC_NUM_AUTO      10

#define ORC_GLWE_WORDS  ((ORC_K + 1) * ORC_N)                     /* 3072 */
#define ORC_GGSW_WORDS  (ORC_CBS_LEVEL * (ORC_K + 1) * ORC_GLWE_WORDS) /* 64512 */

/* Fourier-domain evaluation keys built from standard-domain key material. */
typedef struct orc_keys orc_keys;

/*
 * bsk      : [768][1][3][3][1024]            (bsk.bin payload)
 * ksk      : [8][3][4][256]                  (ksk.bin payload)
 * auto_std : [10][2][3][3][1024], key index i <-> kappa = 1024/2^i + 1
 *            (standard-domain GLWE keyswitch keys, i.e. lo|hi<<41 recombined)
 * ss       : [2][2][3][3][1024]              (ss_key.bin payload)
 */
orc_keys *orc_keys_create(const uint64_t *bsk, const uint64_t *ksk,
                          const uint64_t *auto_std, const uint64_t *ss);
void orc_keys_destroy(orc_keys *k);

/* ---- primitives (exposed for stage-level parity tests) ---- */
uint64_t orc_modswitch(uint64_t x);
void orc_mono_mul(uint64_t *out, const uint64_t *in, int N, unsigned d);
void orc_eval_x_k(uint64_t *out, const uint64_t *in, int N, unsigned kappa);
void orc_sample_extract(uint64_t *lwe, const uint64_t *glwe, int k, int N, int t);
void orc_const_embed(uint64_t *glwe, const uint64_t *lwe, int k, int N);
/* digits[t][i], t = 0 is the FINEST level (level l), signed values */
void orc_decompose(int64_t *digits, const uint64_t *poly, int N, int base_log, int level);
void orc_fft_fwd_torus(double *out_c, const uint64_t *poly, int N);
void orc_fft_fwd_int(double *out_c, const int64_t *poly, int N);
void orc_fft_bwd_torus_add(uint64_t *poly, const double *in_c, int N);
void orc_external_product_add(uint64_t *out, const double *ggsw_f, const uint64_t *glwe,
                              int k, int N, int base_log, int level);
void orc_ggsw_to_fourier(double *out_f, const uint64_t *ggsw_std, int k, int N, int level);

/* ---- stages ---- */
/* a6: keyswitch_lwe_ciphertext_by_glwe_keyswitch  LWE(2048) -> LWE(768) */
void orc_lwe_keyswitch(const orc_keys *K, const uint64_t *in2049, uint64_t *out769);
/* a1: accumulator build + gen_blind_rotate_local_assign */
void orc_blind_rotate(const orc_keys *K, const uint64_t *lwe769, uint64_t *acc3072);
/* a2: level extraction + lwe_preprocessing + const embed  -> glev_pre[7][3072] */
void orc_glev_from_acc(const uint64_t *acc3072, uint64_t *glev_pre);
/* a3: trace_assign on one GLWE (in place) */
void orc_trace(const orc_keys *K, uint64_t *glwe3072);
/* one automorphism-keyswitch step i (1..10) : out = KS_kappa(in(X^kappa)) */
void orc_auto_step(const orc_keys *K, int i, const uint64_t *in3072, uint64_t *out3072);
/* a4: switch_scheme  glev[7][3072] -> ggsw_std[7][3][3072] */
void orc_scheme_switch(const orc_keys *K, const uint64_t *glev, uint64_t *ggsw_std);
/* a1..a4 : LWE(768) -> GGSW std */
void orc_circuit_bootstrap(const orc_keys *K, const uint64_t *lwe769, uint64_t *ggsw_std);
/* a7: 8 Fourier GGSW bits (LSB first) + 2 accumulators -> 8 LWE(2048) (bit t of acc a at 4a+t) */
void orc_lut8_eval(const double *ggsw_f8, const uint64_t *lut2x3072, uint64_t *out8x2049);
/* a8: known_rotate_keyed_lut for one block: luts [16][2][3072] */
void orc_known_rotate(const uint8_t *ct16, const uint64_t *luts, uint64_t *state128x2049);
/* a9 */
void orc_inv_mix_columns_precomp(uint64_t *st, const uint64_t *t9, const uint64_t *t11,
                                 const uint64_t *t13, const uint64_t *t14);
void orc_inv_shift_rows(uint64_t *st);

/*
 * Whole stage-7 path for nblocks independent ECB blocks (the reference does
 * exactly one: server_encrypted_aes_decryption.rs:28-191).
 * k10_9 : [4 (x9,x11,x13,x14)][16][2][3072]
 * k8_1  : [8 (round-1)][4][16][2][3072]
 * k0    : [16][2][3072]
 * out   : [nblocks][128][2049], MSB-first inside each byte.
 */
void orc_aes128_transcipher(const orc_keys *K, const uint8_t *ct, int nblocks,
                            const uint64_t *k10_9, const uint64_t *k8_1, const uint64_t *k0,
                            uint64_t *out);

/* forward direction / CTR mode: he_shift_rows aes_he.rs:348, he_mix_columns_precomp :441 */
void orc_shift_rows(uint64_t *st);
void orc_mix_columns_precomp(uint64_t *st /* x1 in, result out */, const uint64_t *t2, const uint64_t *t3);
void orc_aes128_ctr_transcipher(const orc_keys *K, const uint8_t *ct, int nblocks, const uint8_t *iv,
                                const uint64_t *kf_first, const uint64_t *kf_mid, const uint64_t *kf_last, uint64_t *out);

/* a10: max_of_two (server_encrypted_compute.rs:34-98).  ggsw std-domain, 16 each, MSB first.
 * reset_e = 0 reproduces the reference (glwe_e carried across output bits). */
void orc_max_of_two(const uint64_t *ggsw_a, const uint64_t *ggsw_b,
                    const uint64_t *lwe_a, const uint64_t *lwe_b, uint64_t *out16x2049,
                    int reset_e);
/* stage 8 for nvals 16-bit values (reference: exactly 8): in [nvals*16][2049] -> out [16][2049] */
void orc_max_u16(const orc_keys *K, const uint64_t *in, int nvals, uint64_t *out16x2049);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
