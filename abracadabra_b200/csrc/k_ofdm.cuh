// k_ofdm.cuh -- OFDM demodulation: CFO de-rotation + 2048-point FFT + DQPSK + frequency de-interleave -> int8 soft bits.
//
// Replaces the reference core's per-symbol ffts_execute + differential demodulator + frequency de-interleaver
// (SURVEY.md sec. 8 a5/a6).  One CTA owns a run of 19 consecutive symbols of one transmission frame, so each symbol
// spectrum is produced once (plus one reference symbol per run) and never leaves shared memory: the only HBM traffic is
// the raw IQ read (2/4/8 bytes per sample) and the 3072 soft bits written per symbol.
#pragma once
#include "common.cuh"

#define OFDM_SYMS_PER_CTA 19
#define OFDM_GROUPS 4

// Last FFT pass (radix 4 on the quadruples at pD0 / pD1) fused with the frequency de-interleaver: output X[k] goes straight to
// spec[n], n = QPSK symbol index of carrier k (fdi_inv, 0xffff = unused bin), so the DQPSK stage reads both spectra linearly.
// The two 16-byte halves of a quadruple are fetched in an order that depends on bit 2 of the lane: consecutive lanes are 32 bytes
// apart, and without the swap lanes l and l+4 of a quarter-warp would hit the same banks.
__device__ __forceinline__ void fft_pass_d_scatter(const float2* __restrict__ work, const FftPlan& pl, float2* __restrict__ spec,
                                                   const uint32_t (&wp)[4])
{
    const int s = (threadIdx.x >> 2) & 1;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float2* p = work + (h ? pl.pD1 : pl.pD0);
        const float4 qa = *reinterpret_cast<const float4*>(p + 2 * s), qb = *reinterpret_cast<const float4*>(p + 2 * (s ^ 1));
        const float4 q0 = s ? qb : qa, q1 = s ? qa : qb;
        float2 x0 = make_float2(q0.x, q0.y), x1 = make_float2(q0.z, q0.w), x2 = make_float2(q1.x, q1.y), x3 = make_float2(q1.z, q1.w);
        float2 s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = csub(x1, x3);
        d1 = make_float2(d1.y, -d1.x);
        const uint32_t n0 = wp[2 * h] & 0xffffu, n1 = wp[2 * h] >> 16, n2 = wp[2 * h + 1] & 0xffffu, n3 = wp[2 * h + 1] >> 16;
        if (n0 != 0xffffu) spec[n0] = cadd(s0, s1);
        if (n1 != 0xffffu) spec[n1] = cadd(d0, d1);
        if (n2 != 0xffffu) spec[n2] = csub(s0, s1);
        if (n3 != 0xffffu) spec[n3] = csub(d0, d1);
    }
    __syncthreads();
}

// slots [first, end) hold no frame (lost frames in front of the one this CTA demodulates): they read as erasures; each of the
// frame's four CTAs clears a quarter of them.  Rare, kept out of line so that it costs the demodulation loop no registers.
__device__ __noinline__ void ofdm_clear_lost(int8_t* __restrict__ soft, int first, int end, int gi)
{
    for (int sl = first; sl < end; sl++) {
        uint4* z = reinterpret_cast<uint4*>(soft + (size_t)sl * DAB_FRAME_SOFT) + gi * (DAB_FRAME_SOFT / 16 / OFDM_GROUPS);
        for (int i = threadIdx.x; i < DAB_FRAME_SOFT / 16 / OFDM_GROUPS; i += 256) z[i] = make_uint4(0, 0, 0, 0);
    }
}

// fdi_inv[p], p < 2048: address inside the spectrum buffer of the FFT output at logical position p of the digit-reversed result
// (0xffff = unused bin); fdi_inv[2048 + n]: address of QPSK symbol n (a conflict-free slot inside its group of 16, see init_tables)
template <int FMT>
__global__ void __launch_bounds__(256, 3) k_ofdm_demod(const void* __restrict__ raw, int64_t nsamp, const FrameInfo* __restrict__ frames,
                                                    const ChainState* __restrict__ chain, int range_idx, int skip, int limit,
                                                    const float2* __restrict__ tw, const uint16_t* __restrict__ fdi_inv,
                                                    int8_t* __restrict__ soft, float* __restrict__ symstat)
{
    __shared__ float2 work[FFT_BUF_LEN];
    __shared__ float2 specA[DAB_NCARR];
    __shared__ float2 specB[DAB_NCARR];
    __shared__ float red[32];
    float2* spec = specA;
    float2* prev = specB;
    const int t = threadIdx.x;
    const int gi = blockIdx.x % OFDM_GROUPS;
    // the frame list and its length live on the device (chained by k_fine_sync): the grid is an upper bound, CTAs beyond it leave
    int fi, f_end;
    chain_range(chain, range_idx, skip, limit, fi, f_end);
    fi += blockIdx.x / OFDM_GROUPS;
    if (fi >= f_end) return;
    const FrameInfo fr = frames[fi];
    if (fi > 0 && frames[fi - 1].slot + 1 < fr.slot) ofdm_clear_lost(soft, frames[fi - 1].slot + 1, fr.slot, gi);
    else if (fi == 0 && fr.slot > 0) ofdm_clear_lost(soft, 0, fr.slot, gi);
    const int l_ref = OFDM_SYMS_PER_CTA * gi;
    int l_end = l_ref + OFDM_SYMS_PER_CTA;
    if (l_end > DAB_NSYM - 1) l_end = DAB_NSYM - 1;
    const double cyc_per = (double)fr.cfo_hz / (double)DAB_FS;
    const float cpp = (float)cyc_per;
    int8_t* out_frame = soft + (size_t)fr.slot * DAB_FRAME_SOFT;

    // thread-invariant pieces hoisted out of the symbol loop: FFT plan (addresses + 21 twiddles), de-interleaver targets of the
    // thread's eight FFT outputs, the rotor that advances the de-rotation by 256 samples, whether the run lies inside the buffer
    FftPlan plan;
    fft_plan_init(plan, tw);
    uint32_t wp[4];
    {
        const uint2 p0 = __ldg(reinterpret_cast<const uint2*>(fdi_inv + 4 * t));
        const uint2 p1 = __ldg(reinterpret_cast<const uint2*>(fdi_inv + 1024 + 4 * t));
        wp[0] = p0.x; wp[1] = p0.y; wp[2] = p1.x; wp[3] = p1.y;
    }
    int rp[6];
#pragma unroll
    for (int j = 0; j < 6; j++) rp[j] = __ldg(fdi_inv + DAB_NFFT + t + 256 * j);
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
        // the de-rotated samples x[t + 256 r] are this thread's own inputs of the first radix-8 pass: they stay in registers.
        // `work` is free: its last readers (pass D of the previous symbol) are behind the barrier that ends fft_pass_d_scatter
        float2 x[8];
        symbol_rotor_regs<FMT>(x, rs, (float)base, cpp, step);
        if (l < l_end) {                                                             // in flight during the FFT
            const int64_t nx = fr.t_f + (int64_t)DAB_T_SYM * (l + 1);
            if (inside) fetch_symbol_fast<FMT>(rs, raw, nx); else fetch_symbol<FMT>(rs, raw, nsamp, nx);
        }
        fft2048_planned_abc_regs(work, plan, x);
        fft_pass_d_scatter(work, plan, spec, wp);
        if (l > l_ref) {
            // DQPSK against the previous symbol: thread handles QPSK symbols n = t + 256 j, stored at rp[j]
            float2 y[6];
            float sabs = 0.f, ssq = 0.f;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float2 v = cmul_conj(spec[rp[j]], prev[rp[j]]);
                y[j] = v;
                sabs += fabsf(v.x) + fabsf(v.y);
                ssq += v.x * v.x + v.y * v.y;
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
            int8_t* o = out_frame + (size_t)(l - 1) * DAB_BITS_PER_SYM + t;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                int a = __float2int_rn(y[j].x * g), b = __float2int_rn(y[j].y * g);
                a = max(-SOFT_MAX, min(SOFT_MAX, a));
                b = max(-SOFT_MAX, min(SOFT_MAX, b));
                o[256 * j] = (int8_t)a;                                              // a warp writes 32 consecutive bytes: one sector
                o[DAB_NCARR + 256 * j] = (int8_t)b;
            }
        }
        // no barrier needed here: the next writes into `prev` (as the new `spec`) come after the four barriers of the next FFT
        { float2* tmp = spec; spec = prev; prev = tmp; }
    }
}
