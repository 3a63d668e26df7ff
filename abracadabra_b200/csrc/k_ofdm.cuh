// k_ofdm.cuh -- OFDM demodulation: CFO de-rotation + 2048-point FFT + DQPSK + frequency de-interleave -> int8 soft bits.
//
// Replaces the reference core's per-symbol ffts_execute + differential demodulator + frequency de-interleaver
// (SURVEY.md sec. 8 a5/a6).  One CTA owns a run of 19 consecutive symbols of one transmission frame, so each symbol
// spectrum is produced once (plus one reference symbol per run) and never leaves shared memory: the only HBM traffic is
// the raw IQ read (2/4/8 bytes per sample) and the 3072 soft bits written per symbol.
#pragma once
#include "common.cuh"

struct FrameInfo {
    long long t_f;     // FFT window start of symbol 0 (PRS)
    float cfo_hz;
    int slot;          // frame slot in the soft-bit buffer (gaps = lost frames)
};

#define OFDM_SYMS_PER_CTA 19
#define OFDM_GROUPS 4

// fdi_pos[n] = physical shared-memory index of the FFT output that carries QPSK symbol n
template <int FMT>
__global__ void __launch_bounds__(256, 2) k_ofdm_demod(const void* __restrict__ raw, int64_t nsamp, const FrameInfo* __restrict__ frames,
                                                    const float2* __restrict__ tw, const uint16_t* __restrict__ fdi_pos,
                                                    int8_t* __restrict__ soft, float* __restrict__ symstat)
{
    __shared__ float2 bufA[FFT_BUF_LEN];
    __shared__ float2 bufB[FFT_BUF_LEN];
    __shared__ float red[32];
    const int t = threadIdx.x;
    const int fi = blockIdx.x / OFDM_GROUPS, gi = blockIdx.x % OFDM_GROUPS;
    const FrameInfo fr = frames[fi];
    const int l_ref = OFDM_SYMS_PER_CTA * gi;
    int l_end = l_ref + OFDM_SYMS_PER_CTA;
    if (l_end > DAB_NSYM - 1) l_end = DAB_NSYM - 1;
    const double cyc_per = (double)fr.cfo_hz / (double)DAB_FS;
    const float cpp = (float)cyc_per;
    float2* cur = bufA;
    float2* prev = bufB;
    int8_t* out_frame = soft + (size_t)fr.slot * DAB_FRAME_SOFT;

    // thread-invariant pieces hoisted out of the symbol loop: FFT plan (addresses + 21 twiddles), de-interleaver positions,
    // the rotor that advances the de-rotation by 256 samples, and whether the whole run lies inside the buffer
    FftPlan plan;
    fft_plan_init(plan, tw);
    int ps[8];
    {
        const uint2 p0 = __ldg(reinterpret_cast<const uint2*>(fdi_pos + 4 * t));
        const uint2 p1 = __ldg(reinterpret_cast<const uint2*>(fdi_pos + 1024 + 4 * (t & 127)));
        ps[0] = p0.x & 0xffff; ps[1] = p0.x >> 16; ps[2] = p0.y & 0xffff; ps[3] = p0.y >> 16;
        ps[4] = p1.x & 0xffff; ps[5] = p1.x >> 16; ps[6] = p1.y & 0xffff; ps[7] = p1.y >> 16;
    }
    float2 step;
    {
        double c256 = cyc_per * 256.0;
        c256 -= floor(c256 + 0.5);
        sincospif((float)(-2.0 * c256), &step.y, &step.x);
    }
    const int64_t first = fr.t_f + (int64_t)DAB_T_SYM * l_ref, last = fr.t_f + (int64_t)DAB_T_SYM * l_end + DAB_NFFT;
    const bool inside = first >= 0 && last <= nsamp;
    RawSym<FMT> rs;
    if (inside) fetch_symbol_fast<FMT>(rs, raw, first); else fetch_symbol<FMT>(rs, raw, nsamp, first);
    for (int l = l_ref; l <= l_end; l++) {
        double base = cyc_per * (double)((int64_t)DAB_T_SYM * l);
        base -= floor(base);
        store_symbol_rotor<FMT>(cur, plan.pA, rs, (float)base, cpp, step);
        if (l < l_end) {                                                             // in flight during the FFT
            const int64_t nx = fr.t_f + (int64_t)DAB_T_SYM * (l + 1);
            if (inside) fetch_symbol_fast<FMT>(rs, raw, nx); else fetch_symbol<FMT>(rs, raw, nsamp, nx);
        }
        fft2048_planned(cur, plan);
        if (l > l_ref) {
            // DQPSK against the previous symbol, de-interleaved order: thread handles n = 4t..4t+3 and (t < 128) 1024+4t..
            float2 y[8];
            float sabs = 0.f, ssq = 0.f;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (h == 1 && t >= 128) break;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 v = cmul_conj(cur[ps[4 * h + i]], prev[ps[4 * h + i]]);
                    y[4 * h + i] = v;
                    sabs += fabsf(v.x) + fabsf(v.y);
                    ssq += v.x * v.x + v.y * v.y;
                }
            }
            const float tot = block_sum(sabs, red);
            const float mean = tot / (float)DAB_BITS_PER_SYM;
            const float g = mean > 0.f ? SOFT_TARGET / mean : 0.f;
            if (symstat) {
                const float tsq = block_sum(ssq, red);
                if (t == 0) {
                    float* st = symstat + ((size_t)fr.slot * 75 + (l - 1)) * 2;
                    st[0] = mean; st[1] = tsq / (float)DAB_BITS_PER_SYM;
                }
            }
            int8_t* o = out_frame + (size_t)(l - 1) * DAB_BITS_PER_SYM;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (h == 1 && t >= 128) break;
                const int n0 = h * 1024 + 4 * t;
                uint32_t wr = 0, wi = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    int a = __float2int_rn(y[4 * h + i].x * g), b = __float2int_rn(y[4 * h + i].y * g);
                    a = max(-SOFT_MAX, min(SOFT_MAX, a));
                    b = max(-SOFT_MAX, min(SOFT_MAX, b));
                    wr |= (uint32_t)(a & 0xff) << (8 * i);
                    wi |= (uint32_t)(b & 0xff) << (8 * i);
                }
                *reinterpret_cast<uint32_t*>(o + n0) = wr;
                *reinterpret_cast<uint32_t*>(o + DAB_NCARR + n0) = wi;
            }
        }
        float2* tmp = cur; cur = prev; prev = tmp;
        __syncthreads();
    }
}
