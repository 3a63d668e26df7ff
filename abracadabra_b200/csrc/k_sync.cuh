// k_sync.cuh -- time/frequency synchronisation kernels.
//
// Replaces, inside the closed reference core, "null-symbol/PRS time sync and coarse/fine CFO estimation"
// (observable results: dabsdrNtfPeriodic_t.freqOffset, dabsdr.h:155, and sync level).
// Unlike the reference's per-frame tracking loop, every stage here is stateless and batch-parallel:
//   k_block_power  : one thread per 32-sample block                       (HBM-bound: reads the whole recording once)
//   k_window_sum   : S[b] = sum of 83 block powers                        (tiny)
//   k_candidates   : threshold + neighbourhood maximum of the energy step (tiny)
//   k_fine_sync    : one CTA per candidate: CP correlation, PRS FFT, integer CFO, channel impulse response
#pragma once
#include "common.cuh"

struct FineOut {
    long long t_f;      // sample index where symbol 0 (PRS) FFT window starts
    float cfo_hz;
    float cp_mag;
    float cfo_ratio;
    float cir_ratio;
    float tau1;
    int pad_[2];
};

template <int FMT>
__global__ void __launch_bounds__(256) k_block_power(const void* __restrict__ raw, int64_t nblk, float* __restrict__ P)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    float acc = 0.f;
    if (FMT == B200DAB_FMT_U8) {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(raw) + b * 64);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint4 w = __ldg(p + v);
            uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // two samples per 32-bit word, same order as the scalar checker: i0 q0 i1 q1
                float i0 = (float)((int)(ws[k] & 0xff) - 128), q0 = (float)((int)((ws[k] >> 8) & 0xff) - 128);
                float i1 = (float)((int)((ws[k] >> 16) & 0xff) - 128), q1 = (float)((int)(ws[k] >> 24) - 128);
                acc += __fmaf_rn(i0, i0, __fmul_rn(q0, q0));
                acc += __fmaf_rn(i1, i1, __fmul_rn(q1, q1));
            }
        }
    } else if (FMT == B200DAB_FMT_S16) {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(raw) + b * 128);
#pragma unroll
        for (int v = 0; v < 8; v++) {
            uint4 w = __ldg(p + v);
            uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float i0 = (float)(short)(ws[k] & 0xffff), q0 = (float)(short)(ws[k] >> 16);
                acc += __fmaf_rn(i0, i0, __fmul_rn(q0, q0));
            }
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(raw) + b * 256);
#pragma unroll
        for (int v = 0; v < 16; v++) {
            float4 w = __ldg(p + v);
            acc += __fmaf_rn(w.x, w.x, __fmul_rn(w.y, w.y));
            acc += __fmaf_rn(w.z, w.z, __fmul_rn(w.w, w.w));
        }
    }
    P[b] = acc;
}

// Four consecutive outputs per thread: every block power is read from shared memory once and added into the (up to four)
// running sums it belongs to.  Each sum still receives its 83 terms in ascending order, one __fadd_rn at a time, so the result is
// bit-identical to the scalar checker; the kernel goes from 83 shared-memory reads per output to 21.5.
#define WSUM_PER_THREAD 4
__device__ __forceinline__ int wsum_phys(int x) { return x + (x >> 5); }       // one pad word per 32: 4-word lane strides stay conflict-free
__global__ void __launch_bounds__(256) k_window_sum(const float* __restrict__ P, int64_t nblk, float* __restrict__ S)
{
    constexpr int TILE = 256 * WSUM_PER_THREAD, NEED = TILE + SYNC_WIN + WSUM_PER_THREAD - 1;
    __shared__ float tile[NEED + NEED / 32 + 2];
    const int64_t b0 = (int64_t)blockIdx.x * TILE;
    for (int i = threadIdx.x; i < NEED; i += 256) tile[wsum_phys(i)] = (b0 + i < nblk) ? __ldg(P + b0 + i) : 0.f;
    __syncthreads();
    const int o = threadIdx.x * WSUM_PER_THREAD;               // thread t owns outputs b0 + 4 t .. + 3
    float acc[WSUM_PER_THREAD] = {0.f, 0.f, 0.f, 0.f};
    // term j = i - k of output k is tile[o + i]: j ascending for every output
#pragma unroll
    for (int i = 0; i < WSUM_PER_THREAD - 1; i++) {            // head: outputs k <= i have started
        const float p = tile[wsum_phys(o + i)];
#pragma unroll
        for (int k = 0; k <= i; k++) acc[k] = __fadd_rn(acc[k], p);
    }
#pragma unroll 4
    for (int i = WSUM_PER_THREAD - 1; i < SYNC_WIN; i++) {     // body: all four
        const float p = tile[wsum_phys(o + i)];
#pragma unroll
        for (int k = 0; k < WSUM_PER_THREAD; k++) acc[k] = __fadd_rn(acc[k], p);
    }
#pragma unroll
    for (int i = SYNC_WIN; i < SYNC_WIN + WSUM_PER_THREAD - 1; i++) {   // tail: outputs k > i - 83 are still open
        const float p = tile[wsum_phys(o + i)];
#pragma unroll
        for (int k = 0; k < WSUM_PER_THREAD; k++) if (k > i - SYNC_WIN) acc[k] = __fadd_rn(acc[k], p);
    }
#pragma unroll
    for (int k = 0; k < WSUM_PER_THREAD; k++) {
        const int64_t b = b0 + o + k;
        if (b < nblk) S[b] = (b + SYNC_WIN <= nblk) ? acc[k] : 0.f;
    }
}

__device__ __forceinline__ float energy_step(const float* __restrict__ S, int64_t nblk, int64_t b)
{
    return (b + 2 * SYNC_WIN <= nblk) ? __fsub_rn(__ldg(S + b + SYNC_WIN), __ldg(S + b)) : -1e30f;
}

__device__ __forceinline__ float ratio_k(const float* __restrict__ S, int64_t nblk, int64_t b, int k)
{
    if (b + (int64_t)(k + 1) * SYNC_WIN <= nblk) {
        float den = __ldg(S + b + k * SYNC_WIN);
        if (den > 0.f) return __fdiv_rn(__ldg(S + b), den);
    }
    return 1e30f;
}

// appends candidate sample positions (unordered) to cand[]; sorted afterwards by k_sort_small
__global__ void __launch_bounds__(256) k_candidates(const float* __restrict__ S, int64_t nblk, int64_t nsamp,
                                                    long long* __restrict__ cand, int* __restrict__ n_cand, int max_cand)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    if (!(ratio_k(S, nblk, b, 1) < SYNC_THR_ACQ && ratio_k(S, nblk, b, 2) < SYNC_THR_ACQ && ratio_k(S, nblk, b, 3) < SYNC_THR_ACQ)) return;
    if (b * SYNC_BLK + DAB_T_FRAME + SYNC_FRAME_MARGIN > nsamp) return;
    int64_t lo = b - SYNC_WIN, hi = b + SYNC_WIN;
    if (lo < 0) lo = 0;
    if (hi >= nblk) hi = nblk - 1;
    const float d = energy_step(S, nblk, b);
    for (int64_t i = lo; i <= hi; i++) {
        float di = energy_step(S, nblk, i);
        if (di > d || (di == d && i < b)) return;
    }
    int slot = atomicAdd(n_cand, 1);
    if (slot < max_cand) cand[slot] = b * SYNC_BLK;
}

// rank sort of a short list (n <= a few thousand), single CTA
__global__ void __launch_bounds__(256) k_sort_small(const long long* __restrict__ in, const int* __restrict__ n_in, int max_n,
                                                    long long* __restrict__ out)
{
    int n = *n_in;
    if (n > max_n) n = max_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        long long v = in[i];
        int rank = 0;
        for (int j = 0; j < n; j++) rank += (in[j] < v) ? 1 : 0;
        out[rank] = v;
    }
}

// carrier pairs (c, c+1) grouped by the PRS phase step between them (class q: PRS[c+1] = PRS[c] * exp(j pi/2 q));
// cls_bin[i] = FFT bin of the lower carrier; cls_off[q] .. cls_off[q+1] = entries of class q
__constant__ int c_cls_off[5];

// one CTA (256 threads) per candidate null start
template <int FMT>
__global__ void __launch_bounds__(256) k_fine_sync(const void* __restrict__ raw, int64_t nsamp, const long long* __restrict__ cand,
                                                   const int* __restrict__ n_cand_p, int max_cand, const float2* __restrict__ tw,
                                                   const int8_t* __restrict__ prs_q, const uint16_t* __restrict__ cls_bin,
                                                   FineOut* __restrict__ out)
{
    __shared__ float2 X[FFT_BUF_LEN];
    __shared__ float2 C[FFT_BUF_LEN];
    __shared__ float red[32];
    __shared__ float shift_m[2 * CFO_INT_RANGE + 1];
    __shared__ int s_best, s_peak;
    __shared__ float s_pmax;

    int n_cand = *n_cand_p;
    if (n_cand > max_cand) n_cand = max_cand;
    const int ci = blockIdx.x;
    if (ci >= n_cand) return;
    const int t = threadIdx.x;
    const int64_t null_start = cand[ci];
    FftPlan plan;
    fft_plan_init(plan, tw);

    // (a) cyclic-prefix correlation over all 76 symbols -> fractional CFO
    float sr = 0.f, si = 0.f;
    const int span = CP_CORR_HI - CP_CORR_LO;      // 184
    const int64_t cp0 = null_start + DAB_T_NULL + CP_CORR_LO;
    if (FMT == B200DAB_FMT_U8 && (cp0 & 3) == 0 && (span & 3) == 0 &&
        cp0 + (int64_t)DAB_T_SYM * (DAB_NSYM - 1) + span + DAB_NFFT < nsamp) {
        // aligned fast path: four samples (8 bytes) per load
        const int quads = span / 4;
        for (int i = t; i < DAB_NSYM * quads; i += 256) {
            const int l = i / quads, n = 4 * (i - l * quads);
            const int64_t p = cp0 + (int64_t)DAB_T_SYM * l + n;
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(raw) + 2 * p));
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(raw) + 2 * (p + DAB_NFFT)));
            const uint32_t aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y};
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float ax = (float)((int)((aw[k] >> (16 * h)) & 0xff) - 128), ay = (float)((int)((aw[k] >> (16 * h + 8)) & 0xff) - 128);
                    const float bx = (float)((int)((bw[k] >> (16 * h)) & 0xff) - 128), by = (float)((int)((bw[k] >> (16 * h + 8)) & 0xff) - 128);
                    sr += ax * bx + ay * by;
                    si += ay * bx - ax * by;
                }
        }
    } else {
        for (int i = t; i < DAB_NSYM * span; i += 256) {
            int l = i / span, n = i - l * span;
            int64_t p = cp0 + (int64_t)DAB_T_SYM * l + n;
            if (p + DAB_NFFT < nsamp) {
                float2 a = load_iq<FMT>(raw, p), b = load_iq<FMT>(raw, p + DAB_NFFT);
                sr += a.x * b.x + a.y * b.y;
                si += a.y * b.x - a.x * b.y;
            }
        }
    }
    sr = block_sum(sr, red);
    si = block_sum(si, red);
    const float frac = -atan2f(si, sr) * (1000.f / 6.283185307179586f);
    const float cp_mag = sqrtf(sr * sr + si * si);

    // (b) PRS window, de-rotated by the fractional offset
    const int64_t w0 = null_start + DAB_T_NULL + DAB_T_CP - PRS_WIN_EARLY;
    const float cpp = (float)((double)frac / (double)DAB_FS);
    load_symbol<FMT>(X, raw, nsamp, w0, 0.f, cpp);
    fft2048_planned(X, plan);

    // integer CFO: D[b] = X[b+1] X*[b] for every bin once; then for each of the 33 shifts the sum over the carrier pairs of
    // D[bin + s] rotated by the PRS phase step, evaluated class by class so that the inner loop is load + add only
    for (int b = t; b < DAB_NFFT; b += 256)
        C[b] = cmul_conj(X[fft_phys(fft_pos((b + 1) & 2047))], X[fft_phys(fft_pos(b))]);
    __syncthreads();
    {
        const int w = t >> 5, lane = t & 31;
        for (int si_ = w; si_ < 2 * CFO_INT_RANGE + 1; si_ += 8) {
            const int s = si_ - CFO_INT_RANGE;
            float ar = 0.f, ai = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float qr = 0.f, qi = 0.f;
                for (int i = c_cls_off[q] + lane; i < c_cls_off[q + 1]; i += 32) {
                    const float2 d = C[(__ldg(cls_bin + i) + s + 2048) & 2047];
                    qr += d.x; qi += d.y;
                }
                const float2 r = rot_mq(make_float2(qr, qi), q);
                ar += r.x; ai += r.y;
            }
            ar = warp_sum(ar); ai = warp_sum(ai);
            if (lane == 0) shift_m[si_] = ar * ar + ai * ai;
        }
    }
    __syncthreads();
    if (t == 0) {
        int best = 0; float bm = -1.f, sum = 0.f;
        for (int i = 0; i < 2 * CFO_INT_RANGE + 1; i++) { float m = shift_m[i]; sum += m; if (m > bm) { bm = m; best = i; } }
        s_best = best - CFO_INT_RANGE;
        red[0] = bm / (sum / (float)(2 * CFO_INT_RANGE + 1) + 1e-30f);
    }
    __syncthreads();
    const int best_s = s_best;
    const float cfo_ratio = red[0];
    __syncthreads();

    // (c) channel impulse response: IFFT of X[k+s] * conj(PRS[k]) via conj(FFT(conj(.)))
    for (int i = t; i < FFT_BUF_LEN; i += 256) C[i] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int c = t; c < DAB_NCARR; c += 256) {
        int b = (carrier_bin(c) + best_s + 2048) & 2047;
        float2 x = rot_mq(X[fft_phys(fft_pos(b))], prs_q[c]);
        C[fft_phys(carrier_bin(c))] = make_float2(x.x, -x.y);          // conj -> forward FFT gives conj(IFFT)
    }
    fft2048_planned(C, plan);
    // power profile in natural delay order, stored back into X (reuse) as floats
    float* pw = reinterpret_cast<float*>(X);
    __syncthreads();
    float lmax = -1.f; int lidx = 0; float lsum = 0.f;
    for (int tau = t; tau < DAB_NFFT; tau += 256) {
        float2 v = C[fft_phys(fft_pos(tau))];
        float p = v.x * v.x + v.y * v.y;
        pw[tau] = p;
        lsum += p;
        if (p > lmax) { lmax = p; lidx = tau; }      // ascending tau per thread: first maximum kept
    }
    // block argmax with lowest index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, lmax, o);
        int oi = __shfl_xor_sync(0xffffffffu, lidx, o);
        if (om > lmax || (om == lmax && oi < lidx)) { lmax = om; lidx = oi; }
    }
    __shared__ float wmax[8];
    __shared__ int widx[8];
    if ((t & 31) == 0) { wmax[t >> 5] = lmax; widx[t >> 5] = lidx; }
    const float psum = block_sum(lsum, red);
    if (t == 0) {
        float bm = wmax[0]; int bi = widx[0];
        for (int i = 1; i < 8; i++) if (wmax[i] > bm || (wmax[i] == bm && widx[i] < bi)) { bm = wmax[i]; bi = widx[i]; }
        s_peak = bi; s_pmax = bm;
    }
    __syncthreads();
    if (t == 0) {
        const int peak = s_peak; const float pmax = s_pmax;
        int tau1 = peak;
        for (int d = FINE_PATH_SPAN; d >= 1; d--) {
            int tt = (peak - d + 2048) & 2047;
            if (pw[tt] >= FINE_PATH_THR * pmax) { tau1 = peak - d; break; }
        }
        tau1 = (tau1 + 2048) & 2047;
        if (tau1 >= 1536) tau1 -= 2048;
        FineOut o;
        o.t_f = w0 + tau1 - FINE_BACKOFF;
        o.cfo_hz = 1000.f * (float)best_s + frac;
        o.cp_mag = cp_mag;
        o.cfo_ratio = cfo_ratio;
        o.cir_ratio = pmax / (psum / (float)DAB_NFFT + 1e-30f);
        o.tau1 = (float)tau1;
        o.pad_[0] = o.pad_[1] = 0;
        out[ci] = o;
    }
}
