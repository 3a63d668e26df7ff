// k_tii.cuh -- Transmitter Identification Information: null-symbol FFT, carrier-pair fold, comb search.
//
// Replaces the TII detector inside the reference core, whose results reach RadioControl as DABSDR_NID_TII
// (dabsdrNtfTii_t, dabsdr.h:372-384; consumer src/radiocontrol.cpp:2479-2493, src/tii/tiibackend.cpp:70-88) and whose
// carrier layout is restated in the open source at src/dabtables.cpp:2338-2444 (pattern table, pair = 24 b + sub).
// Black-box behaviour reproduced: TII lives in the nulls of frames whose CIF
// counter is 0 mod 8; the other nulls are subtracted as noise reference; 384-bin spectrum = 192 pairs x 2 carriers
// folded over the four 384-carrier quarters; levels are linear in amplitude and relative to the data carriers.
#pragma once
#include "common.cuh"
#include "k_ofdm.cuh"
#include "../../include/b200dab_tii_gain.h"

// per-bin gain of the reference core's TII spectrum (measured, see the header); applied when folding
__constant__ float c_tii_gain[TII_SPECTRUM_BINS];

// one CTA per frame: fold[f][2p+e] = sum over the 4 quarters of |N[k]|, k = 384 q + 2p + e ; norm[f] = rms PRS carrier magnitude
template <int FMT>
__global__ void __launch_bounds__(256) k_tii_fold(const void* __restrict__ raw, int64_t nsamp, const FrameInfo* __restrict__ frames,
                                                  const ChainState* __restrict__ chain, int range_idx, int skip, int limit,
                                                  const float2* __restrict__ tw, float* __restrict__ fold, float* __restrict__ norm)
{
    __shared__ float2 buf[FFT_BUF_LEN];
    __shared__ float red[32];
    const int t = threadIdx.x;
    int fi, f_end;
    chain_range(chain, range_idx, skip, limit, fi, f_end);   // the grid is an upper bound of the frames of the range
    fi += blockIdx.x;
    if (fi >= f_end) return;
    const FrameInfo fr = frames[fi];
    const double cyc_per = (double)fr.cfo_hz / (double)DAB_FS;
    const float cpp = (float)cyc_per;
    // PRS symbol -> reference carrier magnitude
    load_symbol<FMT>(buf, raw, nsamp, fr.t_f, 0.f, cpp);
    fft2048_smem(buf, tw);
    float acc = 0.f;
    for (int c = t; c < DAB_NCARR; c += 256) {
        float2 v = buf[fft_phys(fft_pos(carrier_bin(c)))];
        acc += v.x * v.x + v.y * v.y;
    }
    acc = block_sum(acc, red);
    if (t == 0) norm[fr.slot] = sqrtf(acc / (float)DAB_NCARR);
    __syncthreads();
    // null symbol: its useful part ends where the PRS cyclic prefix begins
    double base = cyc_per * (double)(-(int64_t)DAB_T_SYM);
    base -= floor(base);
    load_symbol<FMT>(buf, raw, nsamp, fr.t_f - DAB_T_SYM, (float)base, cpp);
    fft2048_smem(buf, tw);
    float* o = fold + (size_t)fr.slot * TII_SPECTRUM_BINS;
    for (int i = t; i < TII_SPECTRUM_BINS; i += 256) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float2 v = buf[fft_phys(fft_pos(carrier_bin(384 * q + i)))];
            a += sqrtf(v.x * v.x + v.y * v.y);
        }
        o[i] = a * c_tii_gain[i];
    }
}

// stage variant for pre-cut null symbols: in [n][2048] complex64 -> fold [n][384]
__global__ void __launch_bounds__(256) k_tii_fold_nulls(const float2* __restrict__ nulls, const float2* __restrict__ tw, float* __restrict__ fold)
{
    __shared__ float2 buf[FFT_BUF_LEN];
    const int t = threadIdx.x;
    const float2* src = nulls + (size_t)blockIdx.x * DAB_NFFT;
#pragma unroll
    for (int r = 0; r < 8; r++) buf[fft_phys(t + 256 * r)] = __ldg(src + t + 256 * r);
    fft2048_smem(buf, tw);
    float* o = fold + (size_t)blockIdx.x * TII_SPECTRUM_BINS;
    for (int i = t; i < TII_SPECTRUM_BINS; i += 256) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float2 v = buf[fft_phys(fft_pos(carrier_bin(384 * q + i)))];
            a += sqrtf(v.x * v.x + v.y * v.y);
        }
        o[i] = a * c_tii_gain[i];
    }
}

// signal-spectrum callback material (dabsdr.h:393, consumer src/radiocontrol.cpp:2604-2609): per frame two 2048-bin linear power
// Where the reference takes the 2048 samples of its spectrum callbacks from (probed, tests/golden/c1_spectrum_u8_30db: every
// callback equals |FFT|^2 of a block of RAW samples, sum ratio 1.00000): SIGNAL = the first 2048 samples of the 2552-sample block
// it pulls for the third symbol of the frame, NULL / TII = the first 2048 samples of the 2656-sample block of the null symbol; with
// its timing locked both blocks begin 4 samples ahead of the symbol (while it is still pulling in they sit up to 400 samples
// earlier).  Relative to our t_f (FFT window of the PRS = frame start + 2656 + 504 - 24 for a single path):
#define SPECTRUM_SIGNAL_OFFSET (2 * DAB_T_SYM - DAB_T_CP + 24 - 4)      /* frame start + 7756 */
#define SPECTRUM_NULL_OFFSET (-(DAB_T_NULL + DAB_T_CP - 24) - 4)        /* frame start - 4 */
// spectra in natural FFT order, unnormalised |FFT|^2 of the raw samples like the reference's: [0] SIGNAL, [1] NULL / TII.
// One CTA per frame.
template <int FMT>
__global__ void __launch_bounds__(256) k_spectrum(const void* __restrict__ raw, int64_t nsamp, const FrameInfo* __restrict__ frames,
                                                  const float2* __restrict__ tw, float* __restrict__ spec)
{
    __shared__ float2 buf[FFT_BUF_LEN];
    const int t = threadIdx.x;
    const FrameInfo fr = frames[blockIdx.x];
#pragma unroll
    for (int which = 0; which < 2; which++) {
        load_symbol<FMT>(buf, raw, nsamp, fr.t_f + (which == 0 ? SPECTRUM_SIGNAL_OFFSET : SPECTRUM_NULL_OFFSET), 0.f, 0.f);
        fft2048_smem(buf, tw);
        float* o = spec + ((size_t)fr.slot * 2 + which) * DAB_NFFT;
        for (int k = t; k < DAB_NFFT; k += 256) {
            const float2 v = buf[fft_phys(fft_pos(k))];
            o[k] = v.x * v.x + v.y * v.y;
        }
        __syncthreads();
    }
}

// |FFT|^2 of the 2048 raw samples starting at w0[blockIdx.x] (stage entry point: parity of the spectrum callback on given windows)
template <int FMT>
__global__ void __launch_bounds__(256) k_spectrum_at(const void* __restrict__ raw, int64_t nsamp, const long long* __restrict__ w0,
                                                     const float2* __restrict__ tw, float* __restrict__ spec)
{
    __shared__ float2 buf[FFT_BUF_LEN];
    load_symbol<FMT>(buf, raw, nsamp, w0[blockIdx.x], 0.f, 0.f);
    fft2048_smem(buf, tw);
    float* o = spec + (size_t)blockIdx.x * DAB_NFFT;
    for (int k = threadIdx.x; k < DAB_NFFT; k += 256) {
        const float2 v = buf[fft_phys(fft_pos(k))];
        o[k] = v.x * v.x + v.y * v.y;
    }
}

struct TiiResult {
    int n_ids;
    int n_tii_frames;
    float norm;                 // sum of the reference magnitudes of the TII frames * 4
    float noise;
    unsigned char main_id[TII_MAX_IDS];
    unsigned char sub_id[TII_MAX_IDS];
    float level[TII_MAX_IDS];
    float spectrum[TII_SPECTRUM_BINS];
};

__constant__ unsigned char c_tii_pattern[70];

// one CTA (384 threads) per window of frames [w0, w1): acc = sum sign[f] * fold[f]; then the comb search.
// sign[f]: +1 TII null, -1 reference null, 0 unusable.
__global__ void __launch_bounds__(384) k_tii_detect(const float* __restrict__ fold, const float* __restrict__ norm, const int* __restrict__ sign,
                                                    const int2* __restrict__ windows, float thr_k, TiiResult* __restrict__ res)
{
    __shared__ float spec[TII_SPECTRUM_BINS];
    __shared__ float pair[TII_PAIRS];
    __shared__ float sorted_rank[TII_PAIRS];
    __shared__ float s_noise, s_norm, s_peak, s_binmax;
    __shared__ int s_ntii, s_count;
    __shared__ unsigned short hits[64];
    __shared__ float warp_max[12];
    const int t = threadIdx.x;
    const int2 win = windows[blockIdx.x];
    float a = 0.f;
    for (int f = win.x; f < win.y; f++) {
        const int sg = sign[f];
        if (sg) a += (float)sg * fold[(size_t)f * TII_SPECTRUM_BINS + t];
    }
    spec[t] = a;
    if (t == 0) {
        float nm = 0.f; int nt = 0;
        for (int f = win.x; f < win.y; f++) if (sign[f] > 0) { nm += norm[f]; nt++; }
        s_norm = nm * 4.f; s_ntii = nt; s_count = 0;
    }
    __syncthreads();
    if (t < TII_PAIRS) pair[t] = spec[2 * t] + spec[2 * t + 1];
    __syncthreads();
    // noise level: median of |pair| (rank selection, 192 values)
    if (t < TII_PAIRS) {
        const float v = fabsf(pair[t]);
        int rank = 0;
        for (int j = 0; j < TII_PAIRS; j++) { float u = fabsf(pair[j]); rank += (u < v || (u == v && j < t)) ? 1 : 0; }
        sorted_rank[rank] = v;
    }
    __syncthreads();
    // strongest bin of the spectrum (level reference), block-wide maximum
    {
        float bm = spec[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if ((t & 31) == 0) warp_max[t >> 5] = bm;
    }
    __syncthreads();
    if (t == 0) {
        s_noise = sorted_rank[TII_PAIRS / 2]; s_peak = sorted_rank[TII_PAIRS - 1];
        float bm = 0.f;
        for (int i = 0; i < 12; i++) bm = fmaxf(bm, warp_max[i]);
        s_binmax = bm;
    }
    __syncthreads();
    // comb search: thread (sub, main) tests the four blocks of the pattern (rule: oracle/daboracle.c orc_tii_detect); hits go to a short list
    const float thr_mean = fmaxf(thr_k * s_noise, 0.02f * s_peak), thr_lit = 0.5f * thr_k * s_noise;
    for (int i = t; i < 24 * 70; i += 384) {
        const int sub = i / 70, m = i - sub * 70;
        const unsigned pat = c_tii_pattern[m];
        float mn = 1e30f, mx = 0.f, sum = 0.f; int cnt = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const float v = pair[24 * b + sub];
            if (v > thr_lit) cnt++;
            if (pat & (0x80u >> b)) { mn = fminf(mn, v); mx = fmaxf(mx, v); sum += v; }
        }
        bool hit = sum >= 4.f * thr_mean && mn > thr_lit;
        if (hit && cnt > 4) hit = mn >= 0.5f * mx;      // several transmitters share this sub id: keep consistent combs only
        if (hit) { const int k = atomicAdd(&s_count, 1); if (k < 64) hits[k] = (unsigned short)i; }
    }
    __syncthreads();
    if (t == 0) {
        TiiResult* r = res + blockIdx.x;
        int nh = min(s_count, 64);
        for (int i = 1; i < nh; i++) { unsigned short v = hits[i]; int j = i - 1; while (j >= 0 && hits[j] > v) { hits[j + 1] = hits[j]; j--; } hits[j + 1] = v; }
        int n = 0;
        for (int k = 0; k < nh && n < TII_MAX_IDS; k++) {      // index = sub * 70 + main: ascending sub, then main (order of the reference)
            const int sub = hits[k] / 70, m = hits[k] - sub * 70;
            const unsigned pat = c_tii_pattern[m];
            float sum = 0.f;
            for (int b = 0; b < 8; b++) if (pat & (0x80u >> b)) sum += pair[24 * b + sub];
            r->main_id[n] = (unsigned char)m; r->sub_id[n] = (unsigned char)sub;
            // like the reference: mean of the transmitter's 8 bins relative to the strongest carrier pair (half its sum)
            r->level[n] = s_peak > 0.f ? (sum / 8.f) / (0.5f * s_peak) : 0.f;
            n++;
        }
        r->n_ids = n; r->n_tii_frames = s_ntii; r->norm = s_norm; r->noise = s_noise;
    }
    res[blockIdx.x].spectrum[t] = fmaxf(spec[t], 0.f);
}
