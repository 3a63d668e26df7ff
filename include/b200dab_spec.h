/*
 * b200dab_spec.h -- numeric specification of the receive chain (framing constants, thresholds, fixed-point
 * formats).  Shared by the CUDA implementation (abracadabra_b200/csrc) and by the CPU checker (oracle/), so that
 * both follow the same integer decisions.  Mode I constants: ETSI EN 300 401, validated against libdabsdr 4.0.1
 * (SURVEY.md Appendix A).
 */
#ifndef B200DAB_SPEC_H
#define B200DAB_SPEC_H

#define DAB_FS            2048000
#define DAB_T_NULL        2656
#define DAB_T_SYM         2552
#define DAB_T_CP          504
#define DAB_NFFT          2048
#define DAB_NCARR         1536
#define DAB_NSYM          76          /* PRS + 75 data symbols */
#define DAB_T_FRAME       196608
#define DAB_BITS_PER_SYM  3072
#define DAB_FRAME_SOFT    (75 * 3072) /* soft bits per transmission frame: 3 FIC symbols + 72 MSC symbols */
#define DAB_FIC_SOFT      9216
#define DAB_CIF_BITS      55296
#define DAB_CIFS_PER_FRAME 4
#define DAB_FIC_PUNCT     2304
#define DAB_FIC_INFO      768
#define DAB_TDI_DEPTH     16          /* time de-interleaver span in CIFs */

/* sample formats of the raw feed (reference: src/input/rawfileinput.h:42-46; f32 = the float input function) */
#define B200DAB_FMT_U8   0
#define B200DAB_FMT_S16  1
#define B200DAB_FMT_F32  2

/* --- coarse time sync on 32-sample block powers */
#define SYNC_BLK          32
#define SYNC_WIN          83          /* 2656 / 32 */
#define SYNC_FRAME_BLKS   6144        /* 196608 / 32 */
#define SYNC_THR_ACQ      0.75f       /* null/PRS energy ratio to acquire */
#define SYNC_CIR_MIN      40.0f       /* PRS channel-impulse-response peak / mean power needed to accept a frame */
#define SYNC_FRAME_MARGIN 160         /* samples a frame may extend beyond its nominal end (fine timing slack) */

/* --- fine sync */
#define CP_CORR_LO        160         /* cyclic-prefix correlation uses CP samples [LO, HI) of every symbol */
#define CP_CORR_HI        344
#define PRS_WIN_EARLY     252         /* the PRS analysis window starts this many samples before the nominal useful part */
#define CFO_INT_RANGE     16          /* integer CFO search +-16 carriers (+-16 kHz) */
#define FINE_BACKOFF      24          /* FFT windows start this many samples before the first detected path */
#define FINE_PATH_SPAN    300         /* first path searched up to this many samples before the CIR maximum */
#define FINE_PATH_THR     0.25f       /* relative power of the earliest accepted path */

/* --- soft bits: int8, positive = bit 0; per-symbol gain puts the mean |soft| at SOFT_TARGET */
#define SOFT_TARGET       48.0f
#define SOFT_MAX          127

/* --- Viterbi: K=7, polynomials 133,171,145,133 (octal, MSB = current bit); state = last 6 inputs, newest = LSB */
#define VIT_STATES        64
#define VIT_START_PENALTY (-8192)
#define VIT_NORM_PERIOD   16          /* every 16 steps all metrics are re-referenced to state 0 */

/* --- TII */
#define TII_PAIRS         192
#define TII_SPECTRUM_BINS 384
#define TII_MAX_IDS       24
#define TII_NOTIFY_FRAMES 40

#endif
