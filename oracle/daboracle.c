/*
 * daboracle.c -- CPU CHECKER for the DAB mode I receive chain.  TEST INFRASTRUCTURE ONLY (see daboracle.h).
 *
 * Plain scalar C restatement of the arithmetic inside the reference's closed `dabsdr` 4.0.1 core
 * (lib/linux_x86_64/libdabsdr.so.4.0.1, interface lib/linux_x86_64/dabsdr.h:397-429), following
 * ETSI EN 300 401 / TS 102 563 and SURVEY.md Appendix A.  Pinned by tests/test_oracle_vs_ref.py against event
 * logs of the binary.  The sample conversion follows src/input/rawfileinput.cpp:643-711 (u8: x-128, s16: x).
 */
#include "daboracle.h"
#include "../include/b200dab_tii_gain.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= tables */
static const int8_t H_TAB[4][32] = {
    {0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1, 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1},
    {0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0, 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0},
    {0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3, 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3},
    {0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2, 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2}};
static const int8_t PRS_NEG[24][2] = {{0,1},{1,2},{2,0},{3,1},{0,3},{1,2},{2,2},{3,3},{0,2},{1,1},{2,2},{3,3},
                                      {0,1},{1,2},{2,3},{3,3},{0,2},{1,2},{2,2},{3,1},{0,1},{1,3},{2,1},{3,2}};
static const int8_t PRS_POS[24][2] = {{0,3},{3,1},{2,1},{1,1},{0,2},{3,2},{2,1},{1,0},{0,2},{3,2},{2,3},{1,3},
                                      {0,0},{3,2},{2,1},{1,3},{0,3},{3,3},{2,3},{1,0},{0,3},{3,0},{2,1},{1,1}};
static const uint32_t PI_MASK[24] = {
    0xC8888888u,0xC888C888u,0xC8C8C888u,0xC8C8C8C8u,0xCCC8C8C8u,0xCCC8CCC8u,0xCCCCCCC8u,0xCCCCCCCCu,
    0xECCCCCCCu,0xECCCECCCu,0xECECECCCu,0xECECECECu,0xEEECECECu,0xEEECEEECu,0xEEEEEEECu,0xEEEEEEEEu,
    0xFEEEEEEEu,0xFEEEFEEEu,0xFEFEFEEEu,0xFEFEFEFEu,0xFFFEFEFEu,0xFFFEFFFEu,0xFFFFFFFEu,0xFFFFFFFFu};
static const int TDI_DELAY[16] = {0,8,4,12,2,10,6,14,1,9,5,13,3,11,7,15};

static int g_init;
static int8_t  g_prs_q[DAB_NCARR];        /* PRS phase (units of pi/2) by carrier index c (ascending k) */
static int16_t g_fdi_c[DAB_NCARR];        /* QPSK symbol n -> carrier index c */
static int16_t g_c_bin[DAB_NCARR];        /* carrier index -> FFT bin */
static float   g_tw[DAB_NFFT][2];         /* exp(-2 pi i k / 2048) */
static int8_t  g_bm_sign[32][4];          /* +1/-1 of the 4 code bits for butterfly j, input 0 */
static uint8_t g_gf_exp[512], g_gf_log[256];
static uint8_t g_prbs[65536];

static inline int carrier_k(int c) { return c < 768 ? c - 768 : c - 767; }

static void init_tables(void)
{
    if (g_init) return;
    for (int b = 0; b < 24; b++)
        for (int j = 0; j < 32; j++) {
            g_prs_q[b * 32 + j] = (H_TAB[PRS_NEG[b][0]][j] + PRS_NEG[b][1]) & 3;
            g_prs_q[768 + b * 32 + j] = (H_TAB[PRS_POS[b][0]][j] + PRS_POS[b][1]) & 3;
        }
    for (int c = 0; c < DAB_NCARR; c++) g_c_bin[c] = (int16_t)((carrier_k(c) + 2048) & 2047);
    int pi = 0, n = 0;
    for (int i = 1; i < 2048; i++) {
        pi = (13 * pi + 511) & 2047;
        if (pi >= 256 && pi <= 1792 && pi != 1024) {
            int k = pi - 1024;
            g_fdi_c[n++] = (int16_t)(k < 0 ? k + 768 : k + 767);
        }
    }
    for (int k = 0; k < DAB_NFFT; k++) {
        double a = -2.0 * M_PI * k / DAB_NFFT;
        g_tw[k][0] = (float)cos(a);
        g_tw[k][1] = (float)sin(a);
    }
    /* code polynomials with bit d <-> a[i-d] (octal 133,171,145,133 have MSB = a[i]) */
    static const int polyr[4] = {0155, 0117, 0123, 0155};
    for (int j = 0; j < 32; j++) {
        int w = j << 1;           /* state j (newest = LSB), input bit 0 */
        for (int q = 0; q < 4; q++) g_bm_sign[j][q] = (__builtin_popcount(w & polyr[q]) & 1) ? -1 : 1;
    }
    int x = 1;
    for (int i = 0; i < 255; i++) {
        g_gf_exp[i] = (uint8_t)x; g_gf_log[x] = (uint8_t)i;
        x <<= 1; if (x & 0x100) x ^= 0x11D;
    }
    for (int i = 255; i < 512; i++) g_gf_exp[i] = g_gf_exp[i - 255];
    orc_prbs(g_prbs, 65536);
    g_init = 1;
}

void orc_prbs(uint8_t *bits, int n)
{
    unsigned reg = 0x1FF;                      /* bit 0 = newest */
    for (int i = 0; i < n; i++) {
        unsigned b = ((reg >> 8) ^ (reg >> 4)) & 1;
        bits[i] = (uint8_t)b;
        reg = ((reg << 1) | b) & 0x1FF;
    }
}

uint16_t orc_crc16(const uint8_t *p, int n)
{
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < n; i++) {
        crc ^= (uint16_t)(p[i] << 8);
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return (uint16_t)~crc;
}

uint16_t orc_firecode(const uint8_t *p)
{
    uint16_t crc = 0;
    for (int i = 0; i < 9; i++) {
        crc ^= (uint16_t)(p[i] << 8);
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x782F) : (uint16_t)(crc << 1);
    }
    return crc;
}

/* ======================================================================= sample access */
static inline void get_iq(const void *raw, int fmt, int64_t n, float *i, float *q)
{
    if (fmt == B200DAB_FMT_U8) {
        const uint8_t *p = (const uint8_t *)raw + 2 * n;
        *i = (float)((int)p[0] - 128); *q = (float)((int)p[1] - 128);
    } else if (fmt == B200DAB_FMT_S16) {
        const int16_t *p = (const int16_t *)raw + 2 * n;
        *i = (float)p[0]; *q = (float)p[1];
    } else {
        const float *p = (const float *)raw + 2 * n;
        *i = p[0]; *q = p[1];
    }
}

/* ======================================================================= coarse sync */
void orc_block_power(const void *raw, int fmt, uint64_t nsamp, float *P)
{
    int64_t nblk = (int64_t)(nsamp / SYNC_BLK);
    for (int64_t b = 0; b < nblk; b++) {
        float acc = 0.f;
        for (int i = 0; i < SYNC_BLK; i++) {
            float x, y;
            get_iq(raw, fmt, b * SYNC_BLK + i, &x, &y);
            acc += fmaf(x, x, y * y);
        }
        P[b] = acc;
    }
}

void orc_null_metric(const float *P, int64_t nblk, float *q)
{
    /* S[b] = sum of 83 block powers from b (sequential float adds); planes 0..2: S[b] / S[b + 83 k], k = 1..3; plane 3: S[b+83] - S[b] */
    float *S = (float *)malloc(sizeof(float) * (size_t)(nblk > 0 ? nblk : 1));
    for (int64_t b = 0; b + SYNC_WIN <= nblk; b++) {
        float acc = 0.f;
        for (int j = 0; j < SYNC_WIN; j++) acc += P[b + j];
        S[b] = acc;
    }
    for (int k = 1; k <= 3; k++)
        for (int64_t b = 0; b < nblk; b++) {
            float v = 1e30f;
            if (b + (k + 1) * SYNC_WIN <= nblk && S[b + k * SYNC_WIN] > 0.f) v = S[b] / S[b + k * SYNC_WIN];
            q[(k - 1) * nblk + b] = v;
        }
    for (int64_t b = 0; b < nblk; b++)      /* plane 3: energy step S[b+83] - S[b], peaks at the null -> PRS edge */
        q[3 * nblk + b] = (b + 2 * SYNC_WIN <= nblk) ? S[b + SYNC_WIN] - S[b] : -1e30f;
    free(S);
}

/* candidate null starts: blocks where all three energy ratios are below the acquisition threshold and the energy
 * step is the maximum of its +-83-block neighbourhood (first index wins ties).  Ascending order. */
int orc_candidates(const float *q, int64_t nblk, uint64_t nsamp, int64_t *cand, int max_cand)
{
    const float *q2 = q + nblk, *q3 = q + 2 * nblk, *d = q + 3 * nblk;
    int n = 0;
    for (int64_t b = 0; b < nblk && n < max_cand; b++) {
        if (!(q[b] < SYNC_THR_ACQ && q2[b] < SYNC_THR_ACQ && q3[b] < SYNC_THR_ACQ)) continue;
        if ((uint64_t)(b * SYNC_BLK + DAB_T_FRAME + SYNC_FRAME_MARGIN) > nsamp) break;   /* incomplete frame */
        int64_t lo = b - SYNC_WIN, hi = b + SYNC_WIN;
        if (lo < 0) lo = 0;
        if (hi >= nblk) hi = nblk - 1;
        int ok = 1;
        for (int64_t i = lo; i <= hi && ok; i++)
            if (d[i] > d[b] || (d[i] == d[b] && i < b)) ok = 0;
        if (ok) cand[n++] = b * SYNC_BLK;
    }
    return n;
}

/* keep candidates whose PRS correlation is sharp, at most one per transmission frame; slot = frame number
 * relative to the first accepted frame (gaps are lost frames). */
int orc_chain_frames(const int64_t *t_f, const float *cir_ratio, int n, int32_t *idx, int32_t *slot)
{
    int nf = 0;
    int64_t last = 0;
    for (int i = 0; i < n; i++) {
        if (!(cir_ratio[i] >= SYNC_CIR_MIN)) continue;
        if (nf > 0 && t_f[i] - last < (int64_t)(DAB_T_FRAME * 9 / 10)) continue;
        idx[nf] = i;
        slot[nf] = nf == 0 ? 0 : slot[nf - 1] + (int32_t)((t_f[i] - last + DAB_T_FRAME / 2) / DAB_T_FRAME);   /* relative to the previous frame: immune to clock drift */
        last = t_f[i];
        nf++;
    }
    return nf;
}

/* ======================================================================= FFT */
void orc_fft2048(float *x, int inverse)
{
    init_tables();
    /* bit reversal */
    for (int i = 0, j = 0; i < DAB_NFFT; i++) {
        if (i < j) {
            float tr = x[2 * i], ti = x[2 * i + 1];
            x[2 * i] = x[2 * j]; x[2 * i + 1] = x[2 * j + 1];
            x[2 * j] = tr; x[2 * j + 1] = ti;
        }
        int m = DAB_NFFT >> 1;
        while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
        j |= m;
    }
    for (int len = 2; len <= DAB_NFFT; len <<= 1) {
        int half = len >> 1, step = DAB_NFFT / len;
        for (int i = 0; i < DAB_NFFT; i += len)
            for (int k = 0; k < half; k++) {
                float wr = g_tw[k * step][0], wi = inverse ? -g_tw[k * step][1] : g_tw[k * step][1];
                float *a = x + 2 * (i + k), *b = x + 2 * (i + k + half);
                float tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr; b[1] = a[1] - ti;
                a[0] += tr; a[1] += ti;
            }
    }
}

/* load 2048 samples starting at `start`, de-rotated by cfo: multiply by exp(-2 pi i cfo (m0 + m)/fs) */
static void load_derot(const void *raw, int fmt, uint64_t nsamp, int64_t start, double cfo_hz, int64_t m0, float *x)
{
    double cyc_per = cfo_hz / (double)DAB_FS;
    double base = cyc_per * (double)m0;
    base -= floor(base);
    for (int m = 0; m < DAB_NFFT; m++) {
        int64_t n = start + m;
        float xi = 0.f, xq = 0.f;
        if (n >= 0 && (uint64_t)n < nsamp) get_iq(raw, fmt, n, &xi, &xq);
        float cyc = (float)base + (float)cyc_per * (float)m;
        float ph = -6.283185307179586f * (cyc - rintf(cyc));
        float c = cosf(ph), s = sinf(ph);
        x[2 * m] = xi * c - xq * s;
        x[2 * m + 1] = xi * s + xq * c;
    }
}

/* ======================================================================= fine sync */
void orc_frame_sync(const void *raw, int fmt, uint64_t nsamp, int64_t null_start, int64_t *t_f, float *cfo_hz, float *metric)
{
    init_tables();
    /* (a) fractional CFO from the cyclic prefix of all 76 symbols */
    double ar = 0, ai = 0;
    for (int l = 0; l < DAB_NSYM; l++) {
        int64_t base = null_start + DAB_T_NULL + (int64_t)DAB_T_SYM * l;
        float sr = 0.f, si = 0.f;
        for (int n = CP_CORR_LO; n < CP_CORR_HI; n++) {
            float a, b, c, d;
            if ((uint64_t)(base + n + DAB_NFFT) >= nsamp) break;
            get_iq(raw, fmt, base + n, &a, &b);
            get_iq(raw, fmt, base + n + DAB_NFFT, &c, &d);
            sr += a * c + b * d;            /* x[n] * conj(x[n+N]) */
            si += b * c - a * d;
        }
        ar += sr; ai += si;
    }
    float frac = (float)(-atan2(ai, ar) / (2.0 * M_PI) * 1000.0);
    if (metric) metric[0] = (float)sqrt(ar * ar + ai * ai);
    /* (b) PRS: integer CFO and channel impulse response */
    static float X[2 * DAB_NFFT], C[2 * DAB_NFFT];
    int64_t w0 = null_start + DAB_T_NULL + DAB_T_CP - PRS_WIN_EARLY;
    load_derot(raw, fmt, nsamp, w0, frac, 0, X);
    orc_fft2048(X, 0);
    int best_s = 0; float best_m = -1.f, sum_m = 0.f;
    for (int s = -CFO_INT_RANGE; s <= CFO_INT_RANGE; s++) {
        float sr = 0.f, si = 0.f;
        for (int c = 0; c + 1 < DAB_NCARR; c++) {
            if (c == 767) continue;                        /* pair would straddle DC */
            int b0 = (g_c_bin[c] + s + 2048) & 2047, b1 = (g_c_bin[c + 1] + s + 2048) & 2047;
            float dr = X[2 * b1] * X[2 * b0] + X[2 * b1 + 1] * X[2 * b0 + 1];
            float di = X[2 * b1 + 1] * X[2 * b0] - X[2 * b1] * X[2 * b0 + 1];
            int dq = (g_prs_q[c + 1] - g_prs_q[c]) & 3;   /* multiply by conj(exp(j pi/2 dq)) */
            float rr, ri;
            switch (dq) { case 0: rr = dr; ri = di; break; case 1: rr = di; ri = -dr; break;
                          case 2: rr = -dr; ri = -di; break; default: rr = -di; ri = dr; }
            sr += rr; si += ri;
        }
        float m = sr * sr + si * si;
        sum_m += m;
        if (m > best_m) { best_m = m; best_s = s; }
    }
    if (metric) metric[1] = best_m / (sum_m / (2 * CFO_INT_RANGE + 1) + 1e-30f);
    memset(C, 0, sizeof C);
    for (int c = 0; c < DAB_NCARR; c++) {
        int b = (g_c_bin[c] + best_s + 2048) & 2047;
        float xr = X[2 * b], xi = X[2 * b + 1], rr, ri;
        switch (g_prs_q[c]) { case 0: rr = xr; ri = xi; break; case 1: rr = xi; ri = -xr; break;
                              case 2: rr = -xr; ri = -xi; break; default: rr = -xi; ri = xr; }
        C[2 * g_c_bin[c]] = rr; C[2 * g_c_bin[c] + 1] = ri;
    }
    orc_fft2048(C, 1);
    int peak = 0; float pmax = -1.f, psum = 0.f;
    for (int t = 0; t < DAB_NFFT; t++) {
        float p = C[2 * t] * C[2 * t] + C[2 * t + 1] * C[2 * t + 1];
        C[2 * t] = p;
        psum += p;
        if (p > pmax) { pmax = p; peak = t; }
    }
    int tau1 = peak;
    for (int d = FINE_PATH_SPAN; d >= 1; d--) {
        int t = (peak - d + 2048) & 2047;
        if (C[2 * t] >= FINE_PATH_THR * pmax) { tau1 = peak - d; break; }
    }
    tau1 = (tau1 + 2048) & 2047;
    if (tau1 >= 1536) tau1 -= 2048;
    if (metric) { metric[2] = pmax / (psum / DAB_NFFT + 1e-30f); metric[3] = (float)tau1; }
    *t_f = w0 + tau1 - FINE_BACKOFF;
    *cfo_hz = 1000.f * (float)best_s + frac;
}

/* ======================================================================= OFDM demod */
void orc_demod_frame(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz, int8_t *soft, float *carriers, float *symstat)
{
    init_tables();
    static float X[2 * DAB_NFFT];
    static float Z[2][2 * DAB_NCARR];
    static float Y[2 * DAB_NCARR];
    for (int l = 0; l < DAB_NSYM; l++) {
        load_derot(raw, fmt, nsamp, t_f + (int64_t)DAB_T_SYM * l, (double)cfo_hz, (int64_t)DAB_T_SYM * l, X);
        orc_fft2048(X, 0);
        float *z = Z[l & 1], *zp = Z[(l & 1) ^ 1];
        for (int c = 0; c < DAB_NCARR; c++) { z[2 * c] = X[2 * g_c_bin[c]]; z[2 * c + 1] = X[2 * g_c_bin[c] + 1]; }
        if (carriers) memcpy(carriers + (size_t)l * 2 * DAB_NCARR, z, sizeof(float) * 2 * DAB_NCARR);
        if (l == 0) continue;
        float sabs = 0.f, ssq = 0.f;
        for (int c = 0; c < DAB_NCARR; c++) {
            float yr = z[2 * c] * zp[2 * c] + z[2 * c + 1] * zp[2 * c + 1];
            float yi = z[2 * c + 1] * zp[2 * c] - z[2 * c] * zp[2 * c + 1];
            Y[2 * c] = yr; Y[2 * c + 1] = yi;
            sabs += fabsf(yr) + fabsf(yi);
            ssq += yr * yr + yi * yi;
        }
        float mean = sabs / (float)DAB_BITS_PER_SYM;
        float g = mean > 0.f ? SOFT_TARGET / mean : 0.f;
        if (symstat) { symstat[2 * (l - 1)] = mean; symstat[2 * (l - 1) + 1] = ssq / (float)DAB_BITS_PER_SYM; }
        int8_t *o = soft + (size_t)(l - 1) * DAB_BITS_PER_SYM;
        for (int n = 0; n < DAB_NCARR; n++) {
            int c = g_fdi_c[n];
            float a = rintf(Y[2 * c] * g), b = rintf(Y[2 * c + 1] * g);
            if (a > SOFT_MAX) a = SOFT_MAX; if (a < -SOFT_MAX) a = -SOFT_MAX;
            if (b > SOFT_MAX) b = SOFT_MAX; if (b < -SOFT_MAX) b = -SOFT_MAX;
            o[n] = (int8_t)a; o[n + DAB_NCARR] = (int8_t)b;
        }
    }
}

/* ======================================================================= channel decoding */
int orc_depuncture_map(const orc_subch_t *sc, int32_t *map)
{
    int pos = 0, idx = 0;
    for (int s = 0; s < sc->n_seg; s++) {
        uint32_t m = PI_MASK[sc->seg_pi[s] - 1];
        for (int l = 0; l < sc->seg_len[s] * 4; l++)
            for (int i = 0; i < 32; i++) map[pos++] = ((m >> (31 - i)) & 1) ? idx++ : -1;
    }
    for (int i = 0; i < 24; i++) map[pos++] = ((0xCCCCCCu >> (23 - i)) & 1) ? idx++ : -1;
    return (pos == 4 * (sc->n_info + 6)) ? idx : -1;
}

void orc_viterbi(const int8_t *soft4, int steps, uint8_t *bits)
{
    /* Maximum-likelihood sequence of the zero-terminated K=7 code, found bidirectionally (same formulation as the CUDA
     * kernel): forward recursion over steps [0, m) from state 0, backward recursion over steps [m, T) from state 0,
     * m = steps/2 rounded down to a multiple of 32; the state maximising alpha + beta at stage m (lowest state on ties)
     * lies on the best path.  Tie rules (include/b200dab_spec.h; chosen to be invariant under complementing all state and
     * input bits, the symmetry the CUDA kernel's lane split relies on): forward, equal candidates into state n -> the
     * predecessor whose oldest bit equals n & 1; backward, equal continuations from state s -> the input bit equal to
     * s >> 5.  With a unique best path this equals one-directional Viterbi. */
    init_tables();
    const int m = ((steps >> 1) >> 5) << 5;
    int32_t M[64], N[64];
    uint64_t *dec = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)steps);
    for (int i = 0; i < 64; i++) M[i] = VIT_START_PENALTY;
    M[0] = 0;
    for (int t = 0; t < m; t++) {
        const int8_t *s = soft4 + 4 * t;
        uint64_t d = 0;
        for (int j = 0; j < 32; j++) {
            int32_t beta = s[0] * g_bm_sign[j][0] + s[1] * g_bm_sign[j][1] + s[2] * g_bm_sign[j][2] + s[3] * g_bm_sign[j][3];
            int32_t c0 = M[j] + beta, c1 = M[j + 32] - beta;
            if (c1 > c0) { N[2 * j] = c1; d |= 1ull << (2 * j); } else N[2 * j] = c0;
            c0 = M[j] - beta; c1 = M[j + 32] + beta;
            if (c1 >= c0) { N[2 * j + 1] = c1; d |= 1ull << (2 * j + 1); } else N[2 * j + 1] = c0;
        }
        dec[t] = d;
        memcpy(M, N, sizeof M);
    }
    int32_t Bm[64], Bn[64];
    for (int i = 0; i < 64; i++) Bm[i] = VIT_START_PENALTY;
    Bm[0] = 0;
    for (int t = steps - 1; t >= m; t--) {
        const int8_t *s = soft4 + 4 * t;
        uint64_t d = 0;
        for (int j = 0; j < 32; j++) {
            int32_t beta = s[0] * g_bm_sign[j][0] + s[1] * g_bm_sign[j][1] + s[2] * g_bm_sign[j][2] + s[3] * g_bm_sign[j][3];
            /* state j: input 0 -> +beta to state 2j, input 1 -> -beta to state 2j+1; state j+32: signs flipped */
            int32_t c0 = Bm[2 * j] + beta, c1 = Bm[2 * j + 1] - beta;
            if (c1 > c0) { Bn[j] = c1; d |= 1ull << j; } else Bn[j] = c0;
            c0 = Bm[2 * j] - beta; c1 = Bm[2 * j + 1] + beta;
            if (c1 >= c0) { Bn[j + 32] = c1; d |= 1ull << (j + 32); } else Bn[j + 32] = c0;
        }
        dec[t] = d;
        memcpy(Bm, Bn, sizeof Bm);
    }
    int sm = 0; int32_t best = M[0] + Bm[0];
    for (int q = 1; q < 64; q++) if (M[q] + Bm[q] > best) { best = M[q] + Bm[q]; sm = q; }
    int s = sm;
    for (int t = m - 1; t >= 0; t--) {
        bits[t] = (uint8_t)(s & 1);
        int d = (int)((dec[t] >> s) & 1);
        s = (s >> 1) | (d << 5);
    }
    s = sm;
    for (int t = m; t < steps; t++) {
        int b = (int)((dec[t] >> s) & 1);
        if (t < steps - 6) bits[t] = (uint8_t)b;
        s = ((s << 1) | b) & 63;
    }
    free(dec);
}

void orc_fic_decode(const int8_t *soft_fic, uint8_t *fibs, uint8_t *crc_ok)
{
    init_tables();
    orc_subch_t fic = {0, 36, DAB_FIC_INFO, 2, {21, 3, 0, 0}, {16, 15, 0, 0}};
    static int32_t map[4 * (DAB_FIC_INFO + 6)];
    int steps = DAB_FIC_INFO + 6;
    orc_depuncture_map(&fic, map);
    for (int v = 0; v < 4; v++) {
        int8_t dp[4 * (DAB_FIC_INFO + 6)];
        uint8_t bits[DAB_FIC_INFO];
        for (int i = 0; i < 4 * steps; i++) dp[i] = map[i] >= 0 ? soft_fic[v * DAB_FIC_PUNCT + map[i]] : 0;
        orc_viterbi(dp, steps, bits);
        for (int i = 0; i < DAB_FIC_INFO; i++) bits[i] ^= g_prbs[i];
        for (int f = 0; f < 3; f++) {
            uint8_t *fib = fibs + (v * 3 + f) * 32;
            for (int i = 0; i < 32; i++) {
                unsigned x = 0;
                for (int b = 0; b < 8; b++) x = (x << 1) | bits[f * 256 + i * 8 + b];
                fib[i] = (uint8_t)x;
            }
            uint16_t crc = orc_crc16(fib, 30);
            crc_ok[v * 3 + f] = (crc == (uint16_t)((fib[30] << 8) | fib[31]));
        }
    }
}

void orc_msc_decode(const int8_t *soft_cifs, int n_cif, const orc_subch_t *sc, uint8_t *frames)
{
    init_tables();
    int steps = sc->n_info + 6;
    int32_t *map = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)steps);
    int8_t *dp = (int8_t *)malloc(4 * (size_t)steps);
    uint8_t *bits = (uint8_t *)malloc((size_t)sc->n_info);
    orc_depuncture_map(sc, map);
    int nbytes = sc->n_info / 8;
    for (int r = 0; r + DAB_TDI_DEPTH - 1 < n_cif; r++) {
        for (int i = 0; i < 4 * steps; i++) {
            int p = map[i];
            dp[i] = p >= 0 ? soft_cifs[(size_t)(r + TDI_DELAY[p & 15]) * DAB_CIF_BITS + (size_t)sc->start_cu * 64 + p] : 0;
        }
        orc_viterbi(dp, steps, bits);
        uint8_t *o = frames + (size_t)r * nbytes;
        for (int i = 0; i < nbytes; i++) {
            unsigned x = 0;
            for (int b = 0; b < 8; b++) x = (x << 1) | (bits[i * 8 + b] ^ g_prbs[i * 8 + b]);
            o[i] = (uint8_t)x;
        }
    }
    free(map); free(dp); free(bits);
}

/* ======================================================================= RS(120,110) over GF(2^8)/0x11D, roots a^0..a^9 */
static inline uint8_t gmul(uint8_t a, uint8_t b) { return (a && b) ? g_gf_exp[g_gf_log[a] + g_gf_log[b]] : 0; }
static inline uint8_t ginv(uint8_t a) { return g_gf_exp[255 - g_gf_log[a]]; }

int orc_rs_decode(uint8_t *cw, int *bit_errors)
{
    init_tables();
    const int n = 120, nroots = 10;
    uint8_t S[10];
    int nz = 0;
    for (int i = 0; i < nroots; i++) {
        uint8_t s = 0;
        for (int j = 0; j < n; j++) s = (uint8_t)(gmul(s, g_gf_exp[i]) ^ cw[j]);   /* Horner at alpha^i */
        S[i] = s; nz |= s;
    }
    if (bit_errors) *bit_errors = 0;
    if (!nz) return 0;
    /* Berlekamp-Massey */
    uint8_t L[12] = {1}, B[12] = {1}, T[12];
    int el = 0, m = 1; uint8_t bb = 1;
    for (int r = 0; r < nroots; r++) {
        uint8_t d = S[r];
        for (int i = 1; i <= el; i++) d ^= gmul(L[i], S[r - i]);
        if (d == 0) { m++; continue; }
        memcpy(T, L, sizeof T);
        uint8_t coef = gmul(d, ginv(bb));
        for (int i = 0; i + m < 12; i++) L[i + m] ^= gmul(coef, B[i]);
        if (2 * el <= r) { el = r + 1 - el; memcpy(B, T, sizeof B); bb = d; m = 1; } else m++;
    }
    if (el > nroots / 2) return -1;
    /* Chien search: error at codeword index j <-> locator X = alpha^(n-1-j) */
    int pos[5], np = 0;
    for (int j = 0; j < n; j++) {
        int p = n - 1 - j;                       /* power of alpha */
        uint8_t xinv = g_gf_exp[(255 - p) % 255];
        uint8_t v = 0, xp = 1;
        for (int i = 0; i <= el; i++) { v ^= gmul(L[i], xp); xp = gmul(xp, xinv); }
        if (v == 0) { if (np < 5) pos[np] = j; np++; }
    }
    if (np != el) return -1;
    /* Forney, first consecutive root = 0:  e = X * Omega(X^-1) / Lambda'(X^-1) */
    uint8_t Om[10] = {0};
    for (int i = 0; i < nroots; i++) {
        uint8_t v = 0;
        for (int k = 0; k <= i && k <= el; k++) v ^= gmul(L[k], S[i - k]);
        Om[i] = v;
    }
    int be = 0;
    for (int e = 0; e < np; e++) {
        int p = n - 1 - pos[e];
        uint8_t X = g_gf_exp[p % 255], xinv = g_gf_exp[(255 - p) % 255];
        uint8_t num = 0, xp = 1;
        for (int i = 0; i < nroots; i++) { num ^= gmul(Om[i], xp); xp = gmul(xp, xinv); }
        uint8_t den = 0;
        for (int i = 1; i <= el; i += 2) {        /* formal derivative: odd terms */
            uint8_t t = L[i];
            uint8_t pw = 1;
            for (int k = 0; k < i - 1; k++) pw = gmul(pw, xinv);
            den ^= gmul(t, pw);
        }
        if (den == 0) return -1;
        uint8_t ev = gmul(gmul(X, num), ginv(den));
        cw[pos[e]] ^= ev;
        be += __builtin_popcount(ev);
    }
    if (bit_errors) *bit_errors = be;
    return np;
}

int orc_superframe(uint8_t *sf, int s, int *au_off, int *au_len, int *au_crc_ok, int *rs_corrected, int *rs_bit_errors,
                   int *rs_uncorrectable, int *firecode_ok)
{
    init_tables();
    int corr = 0, be = 0, unc = 0;
    for (int j = 0; j < s; j++) {
        uint8_t cw[120];
        for (int i = 0; i < 120; i++) cw[i] = sf[j + s * i];
        int b = 0, r = orc_rs_decode(cw, &b);
        if (r < 0) unc++;
        else if (r > 0) { corr += r; be += b; for (int i = 0; i < 120; i++) sf[j + s * i] = cw[i]; }
    }
    if (rs_corrected) *rs_corrected = corr;
    if (rs_bit_errors) *rs_bit_errors = be;
    if (rs_uncorrectable) *rs_uncorrectable = unc;
    int fc = (orc_firecode(sf + 2) == (uint16_t)((sf[0] << 8) | sf[1])) && (sf[0] | sf[1] | sf[2]);
    if (firecode_ok) *firecode_ok = fc;
    if (!fc) return 0;
    int dac = (sf[2] >> 6) & 1, sbr = (sf[2] >> 5) & 1;
    int num = dac ? (sbr ? 3 : 6) : (sbr ? 2 : 4);
    int first = dac ? (sbr ? 6 : 11) : (sbr ? 5 : 8);
    int st[8];
    st[0] = first;
    for (int i = 1; i < num; i++) {
        int bitpos = 24 + 12 * (i - 1);
        int byte = bitpos >> 3;
        st[i] = (bitpos & 7) ? (((sf[byte] & 0x0F) << 8) | sf[byte + 1]) : ((sf[byte] << 4) | (sf[byte + 1] >> 4));
    }
    st[num] = 110 * s;
    for (int i = 0; i < num; i++) if (st[i + 1] <= st[i] + 2 || st[i + 1] > 110 * s) return 0;
    for (int i = 0; i < num; i++) {
        int len = st[i + 1] - st[i];
        au_off[i] = st[i]; au_len[i] = len - 2;
        uint16_t crc = orc_crc16(sf + st[i], len - 2);
        au_crc_ok[i] = crc == (uint16_t)((sf[st[i] + len - 2] << 8) | sf[st[i] + len - 1]);
    }
    return num;
}

/* ======================================================================= TII */
void orc_tii_null(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz, float *spec)
{
    /* useful part of the null symbol ends where the PRS cyclic prefix begins */
    load_derot(raw, fmt, nsamp, t_f - DAB_T_SYM, (double)cfo_hz, -(int64_t)DAB_T_SYM, spec);
    orc_fft2048(spec, 0);
}

void orc_tii_fold(const float *spec, float *fold)
{
    init_tables();
    /* carrier pair p (0..191) of comb quarter qd (0..3): carriers 384*qd + 2p and + 2p + 1 (ascending-k index).
     * fold[2p + e] = sum over the four quarters of |N[k]| -- magnitudes, as the reference's levels are linear in amplitude */
    for (int i = 0; i < TII_SPECTRUM_BINS; i++) {
        float a = 0.f;
        for (int qd = 0; qd < 4; qd++) {
            int b = g_c_bin[384 * qd + i];
            a += sqrtf(spec[2 * b] * spec[2 * b] + spec[2 * b + 1] * spec[2 * b + 1]);
        }
        fold[i] = a * B200DAB_TII_GAIN[i];
    }
}

static int tii_pattern(int main_id)
{
    int n = 0;
    for (int v = 0; v < 256; v++) if (__builtin_popcount(v) == 4) { if (n == main_id) return v; n++; }
    return 0;
}

/* accumulate sign[f] * fold[f] over n frames (+1 = TII null, -1 = TII-free reference null), then the comb search.
 * norm[f] = rms PRS carrier magnitude of frame f.  Returns the number of ids; ids ordered by sub then main. */
int orc_tii_detect(const float *fold, const float *norm, const int32_t *sign, int n, float thr_k, uint8_t *mains, uint8_t *subs,
                   float *levels, int max_ids, float *spectrum /* 384 */)
{
    float spec[TII_SPECTRUM_BINS], pair[TII_PAIRS], absv[TII_PAIRS];
    for (int t = 0; t < TII_SPECTRUM_BINS; t++) {
        float a = 0.f;
        for (int f = 0; f < n; f++) if (sign[f]) a += (float)sign[f] * fold[(size_t)f * TII_SPECTRUM_BINS + t];
        spec[t] = a;
    }
    float nm = 0.f;
    for (int f = 0; f < n; f++) if (sign[f] > 0) nm += norm[f];
    nm *= 4.f;
    for (int p = 0; p < TII_PAIRS; p++) { pair[p] = spec[2 * p] + spec[2 * p + 1]; absv[p] = fabsf(pair[p]); }
    /* median and maximum of |pair| */
    for (int i = 1; i < TII_PAIRS; i++) { float v = absv[i]; int j = i - 1; while (j >= 0 && absv[j] > v) { absv[j + 1] = absv[j]; j--; } absv[j + 1] = v; }
    float noise = absv[TII_PAIRS / 2], peak = absv[TII_PAIRS - 1];
    /* A transmitter (main, sub) is present when the MEAN of the four comb blocks of its pattern reaches thr_k x noise (and 2 % of the
     * strongest pair) and none of the four is below half of that noise multiple; when more than four of the eight blocks of that sub
     * id are lit (several transmitters share it), only combs of even strength are kept.  Level = mean of the transmitter's eight
     * bins relative to the strongest carrier PAIR (half its sum) -- the normaliser the reference's levels imply within 0.8 % over 80
     * multipath recordings (the strongest single bin reads 0.1-0.5 dB low in noise).  Both rules were fitted against libdabsdr 4.0.1
     * on the config-3 recordings (tools/tii_ref_spread.py): 366 of 368 ids agree. */
    float thr_mean = fmaxf(thr_k * noise, 0.02f * peak), thr_lit = 0.5f * thr_k * noise;
    int cnt = 0;
    for (int sub = 0; sub < 24 && cnt < max_ids; sub++)
        for (int m = 0; m < 70 && cnt < max_ids; m++) {
            unsigned pat = (unsigned)tii_pattern(m);
            int lit = 0; float mn = 1e30f, mx = 0.f, sum = 0.f;
            for (int b = 0; b < 8; b++) {
                float v = pair[24 * b + sub];
                if (v > thr_lit) lit++;
                if (pat & (0x80u >> b)) { mn = fminf(mn, v); mx = fmaxf(mx, v); sum += v; }
            }
            int hit = sum >= 4.f * thr_mean && mn > thr_lit;
            if (hit && lit > 4) hit = mn >= 0.5f * mx;
            if (!hit) continue;
            mains[cnt] = (uint8_t)m; subs[cnt] = (uint8_t)sub; levels[cnt] = peak > 0.f ? (sum / 8.f) / (0.5f * peak) : 0.f;
            cnt++;
        }
    if (spectrum) for (int t = 0; t < TII_SPECTRUM_BINS; t++) spectrum[t] = fmaxf(spec[t], 0.f);
    return cnt;
}

/* rms magnitude of the PRS carriers of a frame (TII level reference) */
float orc_prs_rms(const void *raw, int fmt, uint64_t nsamp, int64_t t_f, float cfo_hz)
{
    init_tables();
    static float X[2 * DAB_NFFT];
    load_derot(raw, fmt, nsamp, t_f, (double)cfo_hz, 0, X);
    orc_fft2048(X, 0);
    float acc = 0.f;
    for (int c = 0; c < DAB_NCARR; c++) { int b = g_c_bin[c]; acc += X[2 * b] * X[2 * b] + X[2 * b + 1] * X[2 * b + 1]; }
    return sqrtf(acc / (float)DAB_NCARR);
}

/* ======================================================================= whole chain */
int orc_decode(const void *raw, int fmt, uint64_t nsamp, const orc_subch_t *sc, int n_sc, uint8_t *fibs, uint8_t *fib_ok,
               uint8_t **frames, int64_t *t_f, float *cfo, int max_frames)
{
    init_tables();
    int64_t nblk = (int64_t)(nsamp / SYNC_BLK);
    float *P = (float *)malloc(sizeof(float) * (size_t)(nblk + 1));
    float *q = (float *)malloc(sizeof(float) * 4 * (size_t)(nblk + 1));
    int max_cand = (int)(nblk / SYNC_WIN) + 2;
    int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * (size_t)max_cand);
    int64_t *ct = (int64_t *)malloc(sizeof(int64_t) * (size_t)max_cand);
    float *cc = (float *)malloc(sizeof(float) * (size_t)max_cand), *cm = (float *)malloc(sizeof(float) * (size_t)max_cand);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_cand), *slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_cand);
    orc_block_power(raw, fmt, nsamp, P);
    orc_null_metric(P, nblk, q);
    int nc = orc_candidates(q, nblk, nsamp, cand, max_cand);
    for (int i = 0; i < nc; i++) {
        float m[4];
        orc_frame_sync(raw, fmt, nsamp, cand[i], &ct[i], &cc[i], m);
        cm[i] = m[2];
    }
    int nf = orc_chain_frames(ct, cm, nc, idx, slot);
    int n_slots = nf ? slot[nf - 1] + 1 : 0;
    if (n_slots > max_frames) { n_slots = max_frames; while (nf && slot[nf - 1] >= n_slots) nf--; }
    int8_t *soft = (int8_t *)malloc((size_t)DAB_FRAME_SOFT);
    int8_t *cifs = (int8_t *)calloc((size_t)(n_slots > 0 ? n_slots : 1) * 4 * DAB_CIF_BITS, 1);   /* lost frames = erasures */
    memset(fibs, 0, (size_t)n_slots * 384);
    memset(fib_ok, 0, (size_t)n_slots * 12);
    for (int f = 0; f < n_slots; f++) { t_f[f] = -1; cfo[f] = 0.f; }
    for (int k = 0; k < nf; k++) {
        int f = slot[k], i = idx[k];
        t_f[f] = ct[i]; cfo[f] = cc[i];
        orc_demod_frame(raw, fmt, nsamp, ct[i], cc[i], soft, NULL, NULL);
        orc_fic_decode(soft, fibs + (size_t)f * 384, fib_ok + (size_t)f * 12);
        memcpy(cifs + (size_t)f * 4 * DAB_CIF_BITS, soft + DAB_FIC_SOFT, 4 * DAB_CIF_BITS);
    }
    for (int i = 0; i < n_sc; i++) orc_msc_decode(cifs, 4 * n_slots, &sc[i], frames[i]);
    free(P); free(q); free(cand); free(ct); free(cc); free(cm); free(idx); free(slot); free(soft); free(cifs);
    return n_slots;
}
