/*
 * b200dab.h -- C ABI of the B200-native DAB mode I receive path (libb200dab.so).
 *
 * This is the thin layer the host side calls into CUDA through: plain pointers and sizes, no C++ or torch types.
 * It is the batch form of what the reference obtains from its closed core through the pull interface
 *     dabsdrRegisterInputFcn / dabsdrRegisterDummyInputFcn      (lib/linux_x86_64/dabsdr.h:405-406, feed:
 *                                                                src/input/inputdevice.cpp:70-131 getSamples/skipSamples)
 * and receives back through the five callbacks                  (dabsdr.h:390-394, src/radiocontrol.cpp:2235-2609).
 * The drop-in with the reference's own 24 entry points (include/dabsdr_abi.h, libdabsdr.so.4) is built on top of
 * these calls in abracadabra_b200/csrc/dabsdr_shim.cpp; INTEGRATION.md shows both bindings.
 *
 * Every function returns 0 on success and a negative value on failure (b200dab_last_error gives the text).
 * There is no CPU fallback: b200dab_create fails when no CUDA device is present.
 */
#ifndef B200DAB_H
#define B200DAB_H
#include <stddef.h>
#include <stdint.h>
#include "b200dab_spec.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200DAB_API __attribute__((visibility("default")))
#else
#define B200DAB_API
#endif

typedef struct b200dab_ctx b200dab_ctx;

/* Sub-channel organisation: what FIG 0/1 carries and RadioControl sees as SubChAddr/SubChSize/protection in
 * dabsdrServiceCompListItem_t (dabsdr.h:194-243).  The protection profile is given explicitly as up to four segments
 * of L 128-bit blocks with puncturing index PI (EN 300 401 11.3) so that EEP-A/B and UEP share one code path. */
typedef struct {
    int32_t subch_id;
    int32_t start_cu;        /* 0..863 */
    int32_t size_cu;
    int32_t bitrate;         /* kbit/s; information bits per 24 ms logical frame = 24 * bitrate */
    int32_t n_seg;
    int32_t seg_len[4];
    int32_t seg_pi[4];
    int32_t kind;            /* 0 = DAB (MP2) audio / data, 1 = DAB+ audio (superframes, RS, AUs) */
    int32_t uep_index;       /* -1 for EEP */
    int32_t prot_level;      /* EEP: 0..3 (+4 for option B); UEP: level 1..5 */
} b200dab_subch_t;

#define B200DAB_MAX_SUBCH 64
#define B200DAB_MEM_HOST   0
#define B200DAB_MEM_DEVICE 1

/* option flags of b200dab_decode */
#define B200DAB_OPT_TII        1u    /* run the TII detector */
#define B200DAB_OPT_NO_MSC     2u    /* stop after the FIC (ensemble discovery only) */
#define B200DAB_OPT_SYMSTAT    4u    /* keep per-symbol soft-bit statistics (SNR estimate) */
#define B200DAB_OPT_SPECTRUM   8u    /* per frame: 2048-bin power spectra of the PRS symbol and of the null symbol */

/* per access unit (DAB+), in stream order per sub-channel */
typedef struct {
    int32_t subch_index;     /* index into the sub-channel table of the result */
    int32_t superframe;      /* superframe number within the sub-channel */
    int32_t cif;             /* logical frame (CIF row) the superframe starts at */
    int32_t offset;          /* byte offset of the payload inside that sub-channel's frame buffer */
    int32_t length;          /* payload bytes (CRC stripped) */
    int32_t crc_ok;
    int32_t header;          /* superframe byte 2 (dac_rate, sbr, channel mode, ps, surround) */
} b200dab_au_t;

typedef struct {
    int32_t n_ids;
    int32_t n_tii_frames;
    float norm, noise;
    uint8_t main_id[TII_MAX_IDS];
    uint8_t sub_id[TII_MAX_IDS];
    float level[TII_MAX_IDS];
    float spectrum[TII_SPECTRUM_BINS];
} b200dab_tii_t;

/* per DAB+ superframe (five logical frames), in (sub-channel, row) order */
typedef struct {
    int32_t subch_index;
    int32_t row;             /* logical frame (CIF row) the superframe starts at */
    int32_t firecode_ok;     /* fire code of the header after RS correction */
    int32_t rs_corrected, rs_bit_errors, rs_uncorrectable;   /* corrected bytes, corrected bits, uncorrectable codewords (of bitrate/8) */
    int32_t n_aus;
    int32_t header;
} b200dab_sf_t;

/* All pointers below refer to pinned host memory owned by the context; valid until the next decode/destroy. */
typedef struct {
    int32_t n_candidates;            /* null-symbol candidates examined */
    int32_t n_frames;                /* frames accepted */
    int32_t n_slots;                 /* frame slots spanned (n_slots - n_frames = lost frames) */
    const int64_t *t_f;              /* [n_slots] FFT window start of symbol 0, -1 for a lost slot */
    const float *cfo_hz;             /* [n_slots] */
    const float *cir_ratio;          /* [n_slots] PRS correlation sharpness */
    const uint8_t *fibs;             /* [n_slots * 12][32] */
    const uint8_t *fib_ok;           /* [n_slots * 12] */
    int32_t n_subch;
    const b200dab_subch_t *subch;    /* [n_subch] as used for MSC decoding (given or discovered from FIG 0/1) */
    const uint8_t *const *frames;    /* [n_subch] -> [n_rows][24 * s] logical frames (RS-corrected in place for DAB+) */
    const int32_t *frame_rows;       /* [n_subch] rows = n_cif - 15 */
    int32_t n_aus;
    const b200dab_au_t *aus;         /* [n_aus] */
    int64_t rs_corrected, rs_bit_errors, rs_uncorrectable, firecode_fail, superframes;
    int32_t n_tii;
    const b200dab_tii_t *tii;        /* [n_tii] one per 40-frame window */
    const float *symstat;            /* [n_slots * 75 * 2] mean |soft|, mean soft^2 before scaling (OPT_SYMSTAT) */
    float snr_db;                    /* decision-directed estimate over all frames (OPT_SYMSTAT), on the scale of the reference's snr10
                                      * (dabsdr.h:153; include/b200dab_snr_cal.h) */
    int32_t gpu_launches;            /* kernels launched by this call */
    /* per-frame TII material for streaming callers (OPT_TII): folded null spectra, reference magnitude, TII/reference sign */
    const float *tii_fold;           /* [n_slots][384] */
    const float *tii_norm;           /* [n_slots] */
    const int32_t *tii_sign;         /* [n_slots] +1 TII null, -1 reference null, 0 unknown */
    const float *spectrum;           /* [n_slots][2][2048] linear power, FFT order: [0] PRS symbol, [1] null symbol (OPT_SPECTRUM) */
    /* DAB+ superframe layer in detail (what the periodic statistics and a chunked decode are put together from) */
    int32_t n_sf;
    const b200dab_sf_t *sf;          /* [n_sf] every superframe of the aligned runs */
    const uint8_t *const *sf_flags;  /* [n_subch] -> [frame_rows] 1 = fire code of that logical frame matches BEFORE RS (NULL: not DAB+) */
    const int32_t *sf_phase;         /* [n_subch] row % 5 of the superframe starts (most fire-code hits, or forced); -1: none / not DAB+ */
} b200dab_result_t;

B200DAB_API int b200dab_version(void);
B200DAB_API int b200dab_create(b200dab_ctx **out, int device);
B200DAB_API void b200dab_destroy(b200dab_ctx *ctx);
B200DAB_API const char *b200dab_last_error(b200dab_ctx *ctx);
B200DAB_API void *b200dab_stream(b200dab_ctx *ctx);
/* page-locked host memory for sample buffers handed to b200dab_decode with B200DAB_MEM_HOST: the upload then runs at the full PCIe
 * rate and asynchronously (a pageable buffer is staged by the driver at a fraction of it) */
B200DAB_API void *b200dab_host_alloc(b200dab_ctx *ctx, size_t bytes);
B200DAB_API void b200dab_host_free(b200dab_ctx *ctx, void *p);                 /* cudaStream_t the context launches on */

/* protection profile helpers (EN 300 401 tables 7/8 and 11.3.2); fill seg_len/seg_pi/size_cu/bitrate */
B200DAB_API int b200dab_profile_eep(int bitrate, int level /*1..4*/, int option /*0=A,1=B*/, b200dab_subch_t *out);
B200DAB_API int b200dab_profile_uep(int table_index, b200dab_subch_t *out);
/* parse FIBs (CRC-checked by the caller) into a sub-channel table: FIG 0/1 + FIG 0/2 (ASCTy 63 = DAB+) */
B200DAB_API int b200dab_parse_subch(const uint8_t *fibs, const uint8_t *fib_ok, int n_fibs, b200dab_subch_t *out, int max_out);

/* Shared-memory layout of the OFDM kernel's spectrum buffer (host-side table, exported for tests and tools): out[p], p < 2048, is
 * the slot of the FFT output at position p of the digit-reversed result (0xffff = unused bin), out[2048 + n], n < 1536, the slot
 * of QPSK symbol n.  n must be >= 3584; returns the number of entries or -1. */
B200DAB_API int b200dab_spectrum_layout(uint16_t *out, int n);

/* Bit of the 64-bit survivor word (role 0's word | ~bit-reverse(role 1's word) << 32) in which the Viterbi trace kernel finds the
 * decision of `state` for a forward (backward = 0: state after the step) or backward (state before the step) trellis step; the
 * index function of the kernel itself, exported for the layout test.  -1 for a state outside 0..63. */
B200DAB_API int b200dab_viterbi_decision_index(int state, int backward);
/* Same for the four-lane layout (long codewords): the four 16-bit halves of a step, lane a at bit 16 a.  Returns
 * bit index | invert << 6: the decision is that bit of the word XOR invert.  tools/vit4_model.py derives the tables. */
B200DAB_API int b200dab_viterbi4_decision_index(int state, int backward);
/* Codewords of at least `min_steps_four_lane` trellis steps are decoded with four lanes per codeword pair (shorter dependent
 * chains for the long sub-channels), shorter ones with two; 0 = four lanes for everything, a huge value = two lanes.  Both
 * modes give identical bits (tests/test_gpu_parity.py). */
B200DAB_API int b200dab_set_viterbi_split(b200dab_ctx *ctx, int min_steps_four_lane);

/* Ensemble / service summary out of decoded FIBs: the fields RadioControl receives in dabsdrNtfEnsemble_t and the service
 * list (dabsdr.h:169-318), for batch callers that do not go through the callback interface (band scan output). */
typedef struct {
    uint32_t sid;
    char label[17];
    uint8_t n_components;
    uint8_t subch_id[4];
    uint8_t ascty[4];           /* 63 = DAB+, 0 = MP2 (TMId 0 components only) */
} b200dab_service_t;
typedef struct {
    uint32_t ueid;              /* ECC << 16 | EId */
    int32_t lto, int_table, alarm;
    char label[17];
    int32_t first_cif_count;    /* CIF counter of the first decoded FIG 0/0, -1 if none */
    int32_t n_services;
    b200dab_service_t services[64];
} b200dab_ensemble_t;
B200DAB_API int b200dab_parse_ensemble(const uint8_t *fibs, const uint8_t *fib_ok, int n_fibs, b200dab_ensemble_t *out);

/* Programme associated data of decoded DAB+ access units (host-side byte work; what the reference core reports through
 * dabsdrDynamicLabelCBFunc_t / dabsdrDataGroupCBFunc_t, dabsdr.h:81-96).  Feed the AUs of one service in stream order; every
 * complete, CRC-correct X-PAD data group is reported: kind 0 = dynamic label data group (prefix + characters + CRC),
 * kind >= 4 = MSC data group of the X-PAD application whose start sub-field has that application type. */
typedef struct b200dab_pad b200dab_pad;
typedef void (*b200dab_pad_cb)(void *user, int kind, const uint8_t *data, int len);
B200DAB_API b200dab_pad *b200dab_pad_create(void);
B200DAB_API void b200dab_pad_destroy(b200dab_pad *p);
B200DAB_API void b200dab_pad_reset(b200dab_pad *p);
B200DAB_API int b200dab_pad_feed_au(b200dab_pad *p, const uint8_t *au, int len, b200dab_pad_cb cb, void *user);   /* 1 = AU carried PAD */
/* one logical frame of a packet-mode sub-channel (EN 300 401 sec. 5.3.2): complete MSC data groups of packet address `addr`
 * are reported with kind -1 */
B200DAB_API int b200dab_pad_feed_packets(b200dab_pad *p, int addr, const uint8_t *frame, int len, b200dab_pad_cb cb, void *user);
B200DAB_API int b200dab_pad_feed_mp2(b200dab_pad *p, const uint8_t *frame, int len, b200dab_pad_cb cb, void *user); /* DAB (MPEG Layer II) frame */

/* Whole hot path over one recording (or chunk): sync -> FFT -> DQPSK -> de-interleave -> Viterbi -> FIB CRC ->
 * [FIG 0/1 discovery] -> MSC Viterbi -> RS/AU -> TII.  iq: interleaved I,Q in `fmt`; mem says where it lives. */
B200DAB_API int b200dab_decode(b200dab_ctx *ctx, const void *iq, uint64_t n_samples, int fmt, int mem,
                               const b200dab_subch_t *subch, int n_subch, uint32_t flags, b200dab_result_t *res);

/* Superframe alignment of the NEXT b200dab_decode only: phase[i] >= 0 forces the superframes of sub-channel i (index into the
 * sub-channel table of that decode) to start at rows = phase[i] (mod 5) instead of taking the residue with most fire-code hits
 * of this call.  A recording decoded in time chunks votes over all chunks (abracadabra_b200/shard.py: decode_chunked). */
B200DAB_API int b200dab_set_superframe_phase(b200dab_ctx *ctx, const int32_t *phase, int n);

/* per-kernel device times (CUDA events on the context stream) of the last b200dab_decode; names ';'-separated */
B200DAB_API int b200dab_set_profiling(b200dab_ctx *ctx, int on);
B200DAB_API int b200dab_get_kernel_times(b200dab_ctx *ctx, char *names, int names_cap, float *ms, int max_n);
/* same for the call `back` calls ago (0 = last; the events of the last 8 calls are kept) */
B200DAB_API int b200dab_get_kernel_times_back(b200dab_ctx *ctx, int back, char *names, int names_cap, float *ms, int max_n);
/* same, plus where each kernel started on the stream (ms after the first kernel of that call): shows the gaps between kernels */
B200DAB_API int b200dab_get_kernel_timeline(b200dab_ctx *ctx, int back, char *names, int names_cap, float *start_ms, float *ms, int max_n);

/* ---- stage entry points (host buffers in, host buffers out): the kernels of the path one by one, for parity tests
 *      against oracle/ and for profiling.  They launch exactly the kernels b200dab_decode launches. */
B200DAB_API int b200dab_stage_coarse(b200dab_ctx *ctx, const void *iq, uint64_t n_samples, int fmt, float *P /* n/32 */,
                                     float *S /* n/32 */, int64_t *cand, int32_t *n_cand, int max_cand);
B200DAB_API int b200dab_stage_fine(b200dab_ctx *ctx, const void *iq, uint64_t n_samples, int fmt, const int64_t *cand, int n_cand,
                                   int64_t *t_f, float *cfo_hz, float *metric /* [n_cand][4] */);
B200DAB_API int b200dab_stage_demod(b200dab_ctx *ctx, const void *iq, uint64_t n_samples, int fmt, const int64_t *t_f,
                                    const float *cfo_hz, int n_frames, int8_t *soft /* [n_frames][230400] */);
B200DAB_API int b200dab_stage_viterbi(b200dab_ctx *ctx, const int8_t *soft4 /* [n_cw][4*steps] */, int n_cw, int steps,
                                      uint8_t *bits /* [n_cw][steps-6] one bit per byte, before descrambling */);
B200DAB_API int b200dab_stage_fic(b200dab_ctx *ctx, const int8_t *soft /* [n_frames][230400] */, int n_frames,
                                  uint8_t *fibs /* [n_frames*12][32] */, uint8_t *fib_ok);
B200DAB_API int b200dab_stage_msc(b200dab_ctx *ctx, const int8_t *soft /* [n_frames][230400] */, int n_frames,
                                  const b200dab_subch_t *sc, uint8_t *frames /* [4*n_frames-15][24*bitrate/8] */);
B200DAB_API int b200dab_stage_dabplus(b200dab_ctx *ctx, uint8_t *frames /* [n_rows][pitch], corrected in place */, int n_rows,
                                      int bitrate, b200dab_au_t *aus, int32_t *n_aus, int max_aus, int64_t *stats /* [5] */);
/* the spectrum callback's arithmetic on given windows: out[i] = |FFT|^2 (2048 bins, FFT order, unnormalised) of the 2048 raw samples
 * starting at sample w0[i] -- what dabsdrSpectrumCBFunc_t delivers (dabsdr.h:393; pinned by tests/golden/c1_spectrum_u8_30db) */
B200DAB_API int b200dab_stage_spectrum(b200dab_ctx *ctx, const void *iq, uint64_t n_samples, int fmt, const int64_t *w0, int n_w,
                                       float *out /* [n_w][2048] */);
B200DAB_API int b200dab_stage_tii_fold(b200dab_ctx *ctx, const float *nulls /* [n][2048] complex64 */, int n, float *fold /* [n][384] */);
B200DAB_API int b200dab_stage_tii_detect(b200dab_ctx *ctx, const float *fold /* [n][384] */, const float *norm /* [n] */,
                                         const int32_t *sign /* [n] */, int n, float thr_k, b200dab_tii_t *out);

#ifdef __cplusplus
}
#endif
#endif
