/*
 * daboracle.h -- CPU CHECKER for the DAB mode I receive chain.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.  The product
 * (abracadabra_b200/csrc) never includes, links or calls it.
 *
 * What it restates: the arithmetic of the reference's closed receiver core `dabsdr` 4.0.1
 * (/root/reference/lib/linux_x86_64/libdabsdr.so.4.0.1; interface lib/linux_x86_64/dabsdr.h:397-429; call sites
 * src/radiocontrol.cpp:81-104, :2235-2497, :2542-2602).  That core's source is NOT in the reference tree
 * (CMakeLists.txt:138-146 would build it from an absent dabsdr/ directory), so each function below follows the
 * published algorithm -- ETSI EN 300 401 (mode I) and ETSI TS 102 563 (DAB+) -- with the constants of SURVEY.md
 * Appendix A, and the restatement as a whole is pinned by running the binary itself: tests/test_oracle_vs_ref.py
 * and tests/golden/ hold event logs of libdabsdr 4.0.1 on synthetic files (AU bytes, FIB/RS/CRC counters, CFO,
 * TII ids) which this checker must reproduce.
 */
#ifndef DABORACLE_H
#define DABORACLE_H
#include <stdint.h>
#include "../include/b200dab_spec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t start_cu;      /* CU address 0..863 */
    int32_t size_cu;
    int32_t n_info;        /* information bits per logical frame = bitrate * 24 */
    int32_t n_seg;         /* protection profile: segments of L blocks (128 mother bits) with puncturing index PI */
    int32_t seg_len[4];
    int32_t seg_pi[4];
} orc_subch_t;

/* ---- sync */
void orc_block_power(const void *raw, int fmt, uint64_t nsamp, float *P /* nsamp/32 */);
void orc_null_metric(const float *P, int64_t nblk, float *q /* 4 planes of nblk */);
int  orc_candidates(const float *q, int64_t nblk, uint64_t nsamp, int64_t *cand, int max_cand);
int  orc_chain_frames(const int64_t *t_f, const float *cir_ratio, int n, int32_t *idx, int32_t *slot);
void orc_frame_sync(const void *raw, int fmt, uint64_t nsamp, int64_t null_start, int64_t *t_f, float *cfo_hz,
                    float *metric /* [4]: cp corr mag, int-cfo peak ratio, cir peak/mean, tau1 */);
/* ---- OFDM demod: 76 FFTs -> DQPSK -> frequency de-interleave -> int8 soft bits [75*3072]; spectra optional */
void orc_demod_frame(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz, int8_t *soft,
                     float *carriers /* optional [76*1536*2] */, float *symstat /* optional [75*2]: mean|v|, mean v^2 */);
void orc_fft2048(float *re_im /* 4096 floats, in place */, int inverse);
/* ---- channel decoding */
int  orc_depuncture_map(const orc_subch_t *sc, int32_t *map /* 4*(n_info+6) */);
void orc_viterbi(const int8_t *soft4 /* 4*steps */, int steps, uint8_t *bits /* steps-6 */);
void orc_fic_decode(const int8_t *soft_fic /* 9216 */, uint8_t *fibs /* 12*32 */, uint8_t *crc_ok /* 12 */);
void orc_msc_decode(const int8_t *soft_cifs /* [n_cif][55296] */, int n_cif, const orc_subch_t *sc,
                    uint8_t *frames /* [(n_cif-15)][n_info/8] */);
uint16_t orc_crc16(const uint8_t *p, int n);
uint16_t orc_firecode(const uint8_t *p /* 9 bytes */);
void orc_prbs(uint8_t *bits, int n);
/* ---- DAB+ */
int  orc_rs_decode(uint8_t *cw /* 120 */, int *bit_errors);   /* returns corrected symbols, -1 uncorrectable */
/* one superframe (5 logical frames = 120*s bytes): RS-correct in place, split into AUs.
 * returns number of AUs (0 if header invalid); au_off/au_len index into sf (payload without CRC). */
int  orc_superframe(uint8_t *sf, int s, int *au_off, int *au_len, int *au_crc_ok, int *rs_corrected, int *rs_bit_errors,
                    int *rs_uncorrectable, int *firecode_ok);
/* ---- TII */
void orc_tii_null(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz, float *spec2048 /* re,im interleaved 4096 */);
/* fold one null-symbol spectrum into 384 bins: sum over the four carrier quarters of |N[k]| */
void orc_tii_fold(const float *spec2048, float *fold384);
int  orc_tii_detect(const float *fold, const float *norm, const int32_t *sign, int n, float thr_k, uint8_t *mains, uint8_t *subs,
                    float *levels, int max_ids, float *spectrum384);
float orc_prs_rms(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz);
/* ---- whole chain (CPU baseline "port"): returns number of frames decoded */
int  orc_decode(const void *raw, int fmt, uint64_t nsamp, const orc_subch_t *sc, int n_sc,
                uint8_t *fibs /* [4*max_frames][96] */, uint8_t *fib_ok /* [12*max_frames] */,
                uint8_t **frames /* n_sc pointers, [4*max_frames-15][n_info/8] */, int64_t *t_f, float *cfo, int max_frames);

#ifdef __cplusplus
}
#endif
#endif
