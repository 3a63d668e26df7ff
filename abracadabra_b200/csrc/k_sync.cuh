// k_sync.cuh -- time/frequency synchronisation kernels.
//
// Replaces, inside the closed reference core, "null-symbol/PRS time sync and coarse/fine CFO estimation"
// (observable results: dabsdrNtfPeriodic_t.freqOffset, dabsdr.h:155, and sync level).
// Unlike the reference's per-frame tracking loop, every stage here is stateless and batch-parallel:
//   k_coarse       : power of 32-sample blocks, S[b] = sum of 83 block powers, threshold + neighbourhood maximum of the energy
//                    step, sorted candidate list -- one launch (HBM-bound: reads the whole recording once)
//   k_fine_sync    : one CTA per candidate: CP correlation, PRS FFT, integer CFO, channel impulse response; the last CTA chains
//                    the frames (slot numbers, lost frames) on the device
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

// power of one 32-sample block, summed in the order of the scalar checker
template <int FMT>
__device__ __forceinline__ float block_power_one(const void* __restrict__ raw, int64_t b)
{
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
    return acc;
}

template <int FMT>
__global__ void __launch_bounds__(256) k_block_power(const void* __restrict__ raw, int64_t nblk, float* __restrict__ P)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk) P[b] = block_power_one<FMT>(raw, b);
}

// Everything of the coarse synchronisation behind the block powers in ONE launch (was window sum -> candidates -> sort): a CTA owns
// CO_TILE (1712) consecutive blocks, fetches their powers P plus the halo the candidate test looks at (83 blocks below, 335 above) into shared
// memory, computes the window sums S over them and runs the candidate test for its own blocks.  Candidates of a tile are at least 84
// blocks apart (neighbourhood maximum), so a tile holds at most CO_CAP of them; they leave the CTA in ascending order, and the last CTA
// to finish strings the tiles together: the list is sorted by construction, no atomically appended list and no sort kernel.
// Arithmetic and summation order are those of the kernels this replaces (and of the scalar checker): each S receives its 83 terms in
// ascending order, one __fadd_rn at a time.
#define WSUM_PER_THREAD 8                                /* outputs per thread: every block power is read once and added into up to eight running sums */
#define CO_S_LEN 2048                                    /* window sums per CTA = 256 threads x 8: S[b0 - 83 .. b0 - 83 + 2048) */
#define CO_TILE (CO_S_LEN - 4 * SYNC_WIN - 4)            /* blocks owned by a CTA: the test of block b looks at S[b - 83 .. b + 249] */
#define CO_P_LEN (CO_S_LEN + SYNC_WIN - 1)
#define CO_CAP 24
__device__ __forceinline__ int wsum_phys(int x) { return x + (x >> 5); }       // one pad word per 32: lane strides of 4 or 8 words stay conflict-free
__global__ void __launch_bounds__(256) k_coarse(const float* __restrict__ P, int64_t nblk, int64_t nsamp, float* __restrict__ S,
                                                long long* __restrict__ cand_tile, int* __restrict__ cand_cnt, int* __restrict__ ticket,
                                                long long* __restrict__ cand, int* __restrict__ n_cand, int max_cand)
{
    __shared__ float sP[CO_P_LEN + CO_P_LEN / 32 + 2];
    __shared__ float sS[CO_S_LEN];
    __shared__ int s_list[CO_CAP];
    __shared__ int s_n, s_last;
    const int t = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * CO_TILE, g0 = b0 - SYNC_WIN;      // g0: global block index of sP[0] / sS[0]
    if (t == 0) s_n = 0;
    for (int i = t; i < CO_P_LEN; i += 256) {
        const int64_t g = g0 + i;
        sP[wsum_phys(i)] = (g >= 0 && g < nblk) ? __ldg(P + g) : 0.f;
    }
    __syncthreads();
    for (int quad = t; quad * WSUM_PER_THREAD < CO_S_LEN; quad += 256) {
        const int o = quad * WSUM_PER_THREAD;                  // outputs o .. o + 3; term j = i - k of output k is sP[o + i]
        float acc[WSUM_PER_THREAD];
#pragma unroll
        for (int k = 0; k < WSUM_PER_THREAD; k++) acc[k] = 0.f;
#pragma unroll
        for (int i = 0; i < WSUM_PER_THREAD - 1; i++) {        // head: outputs k <= i have started
            const float p = sP[wsum_phys(o + i)];
#pragma unroll
            for (int k = 0; k <= i; k++) acc[k] = __fadd_rn(acc[k], p);
        }
#pragma unroll 4
        for (int i = WSUM_PER_THREAD - 1; i < SYNC_WIN; i++) { // body: all four
            const float p = sP[wsum_phys(o + i)];
#pragma unroll
            for (int k = 0; k < WSUM_PER_THREAD; k++) acc[k] = __fadd_rn(acc[k], p);
        }
#pragma unroll
        for (int i = SYNC_WIN; i < SYNC_WIN + WSUM_PER_THREAD - 1; i++) {   // tail: outputs k > i - 83 are still open
            const float p = (o + i < CO_P_LEN) ? sP[wsum_phys(o + i)] : 0.f;
#pragma unroll
            for (int k = 0; k < WSUM_PER_THREAD; k++) if (k > i - SYNC_WIN) acc[k] = __fadd_rn(acc[k], p);
        }
#pragma unroll
        for (int k = 0; k < WSUM_PER_THREAD; k++) {
            const int64_t g = g0 + o + k;
            const float v = (g >= 0 && g + SYNC_WIN <= nblk) ? acc[k] : 0.f;
            if (o + k < CO_S_LEN) sS[o + k] = v;
            if (o + k >= SYNC_WIN && o + k < SYNC_WIN + CO_TILE && g < nblk) S[g] = v;
        }
    }
    __syncthreads();
    // candidate test for the owned blocks: null/PRS energy ratios, then the maximum of the energy step in the +-83-block neighbourhood
    auto Sg = [&](int64_t g) -> float { return sS[(int)(g - g0)]; };
    auto estep = [&](int64_t g) -> float { return (g + 2 * SYNC_WIN <= nblk) ? __fsub_rn(Sg(g + SYNC_WIN), Sg(g)) : -1e30f; };
    auto ratio = [&](int64_t g, int k) -> float {
        if (g + (int64_t)(k + 1) * SYNC_WIN <= nblk) {
            const float den = Sg(g + k * SYNC_WIN);
            if (den > 0.f) return __fdiv_rn(Sg(g), den);
        }
        return 1e30f;
    };
    for (int i = t; i < CO_TILE; i += 256) {
        const int64_t b = b0 + i;
        if (b >= nblk) break;
        if (!(ratio(b, 1) < SYNC_THR_ACQ && ratio(b, 2) < SYNC_THR_ACQ && ratio(b, 3) < SYNC_THR_ACQ)) continue;
        if (b * SYNC_BLK + DAB_T_FRAME + SYNC_FRAME_MARGIN > nsamp) continue;
        int64_t lo = b - SYNC_WIN, hi = b + SYNC_WIN;
        if (lo < 0) lo = 0;
        if (hi >= nblk) hi = nblk - 1;
        const float d = estep(b);
        bool is_max = true;
        for (int64_t j = lo; j <= hi; j++) {
            const float dj = estep(j);
            if (dj > d || (dj == d && j < b)) { is_max = false; break; }
        }
        if (!is_max) continue;
        const int slot = atomicAdd(&s_n, 1);
        if (slot < CO_CAP) s_list[slot] = i;
    }
    __syncthreads();
    if (t == 0) {
        int n = min(s_n, CO_CAP);
        for (int a = 1; a < n; a++) {                            // at most a couple of dozen entries
            const int v = s_list[a];
            int c = a - 1;
            while (c >= 0 && s_list[c] > v) { s_list[c + 1] = s_list[c]; c--; }
            s_list[c + 1] = v;
        }
        for (int a = 0; a < n; a++) cand_tile[(size_t)blockIdx.x * CO_CAP + a] = (b0 + s_list[a]) * SYNC_BLK;
        cand_cnt[blockIdx.x] = n;
        __threadfence();
        const int tk = atomicAdd(ticket, 1);
        s_last = (tk == (int)gridDim.x - 1);
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last CTA: tiles in order -> one sorted list
    __shared__ int s_run;
    if (t == 0) s_run = 0;
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    __shared__ int wtot[8];
    for (int base = 0; base < (int)gridDim.x; base += 256) {
        const int tile = base + t;
        const int n = tile < (int)gridDim.x ? __ldcg(cand_cnt + tile) : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        int pos = s_run + incl - n;
        for (int w = 0; w < warp; w++) pos += wtot[w];
        for (int a = 0; a < n; a++)
            if (pos + a < max_cand) cand[pos + a] = __ldcg(cand_tile + (size_t)tile * CO_CAP + a);
        __syncthreads();
        if (t == 0) { int sum = 0; for (int w = 0; w < 8; w++) sum += wtot[w]; s_run += sum; }
        __syncthreads();
    }
    if (t == 0) *n_cand = s_run;
}

// carrier pairs (c, c+1) grouped by the PRS phase step between them (class q: PRS[c+1] = PRS[c] * exp(j pi/2 q));
// cls_bin[i] = FFT bin of the lower carrier; cls_off[q] .. cls_off[q+1] = entries of class q
__constant__ int c_cls_off[5];

// one CTA (256 threads) per candidate null start
template <int FMT>
__global__ void __launch_bounds__(256) k_fine_sync(const void* __restrict__ raw, int64_t nsamp, const long long* __restrict__ cand,
                                                   const int* __restrict__ n_cand_p, int max_cand, const float2* __restrict__ tw,
                                                   const int8_t* __restrict__ prs_q, const uint16_t* __restrict__ cls_bin,
                                                   FineOut* __restrict__ out,
                                                   long long range_off, int range_idx, int max_slots, ChainState* __restrict__ chain,
                                                   FrameInfo* __restrict__ frames, float* __restrict__ frame_cir, int* __restrict__ ticket)
{
    __shared__ __align__(16) float2 X[FFT_BUF_LEN];
    __shared__ __align__(16) float2 C[FFT_BUF_LEN];
    __shared__ float red[32];
    __shared__ float shift_m[2 * CFO_INT_RANGE + 1];
    __shared__ int s_best, s_peak;
    __shared__ float s_pmax;

    int n_cand = *n_cand_p;
    if (n_cand > max_cand) n_cand = max_cand;
    const int t = threadIdx.x;
    FftPlan plan;
    fft_plan_init(plan, tw);
    // a CTA takes candidates blockIdx.x, blockIdx.x + gridDim.x, ...: the grid is sized for the frames a range can hold, and a range
    // with (many) more candidates than frames -- noise-like input -- is still looked at completely
    for (int ci = blockIdx.x; ci < n_cand; ci += gridDim.x) {
    const int64_t null_start = cand[ci];

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
    __syncthreads();                                    // shared buffers are reused by the next candidate
    }
    if (!chain) return;
    // ---- chaining (last CTA to finish): candidates in ascending order; a candidate is a frame if its channel impulse response
    //      stands out and it lies at least 0.9 frame periods behind the previous frame; the slot (frame number) advances by the
    //      rounded distance to the previous frame, so gaps = lost frames = erasures, and a sample-clock offset that accumulates to
    //      more than half a frame over a long recording cannot shift the CIF numbering.
    __shared__ int s_is_last;
    if (t == 0) {
        __threadfence();                                // thread 0 wrote this CTA's results: one fence, not one per thread
        const int tk = atomicAdd(ticket, 1);
        s_is_last = (tk == (int)gridDim.x - 1);
        if (s_is_last) *ticket = 0;
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    // Candidates are taken 1024 at a time: (1) four consecutive candidates per thread, those whose impulse response stands out are
    // compacted in order into shared memory; (2) if every one of them lies at least 0.9 frame periods behind its predecessor -- the
    // normal case -- all are frames and their slots are a prefix sum of the rounded distances: fully parallel; (3) otherwise (a
    // false candidate next to a frame) one thread walks the compacted list.
    struct Ent { long long t; float cfo, cir; };                // 16 bytes
    Ent* sE = reinterpret_cast<Ent*>(X);                        // FFT_BUF_LEN * 8 / 16 = 1152 entries
    int* sSlot = reinterpret_cast<int*>(C);                     // slot of entry e (-1: not a frame)
    __shared__ int c_frames, c_slot, s_wsum[8], s_carry;
    __shared__ long long c_last;
    const int lane = t & 31, warp = t >> 5;
    auto excl_scan = [&](int v, int& total) -> int {            // exclusive prefix over the CTA in thread order
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        __syncthreads();
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const int x = s_wsum[w]; if (w < warp) before += x; tot += x; }
        total = tot;
        return before + incl - v;
    };
    auto frames_between = [](long long gap) -> int {            // rounded distance in frame periods
        if (gap >= 0 && gap < (1ll << 30)) return (int)(((unsigned)gap + DAB_T_FRAME / 2) / (unsigned)DAB_T_FRAME);
        return (int)((gap + DAB_T_FRAME / 2) / DAB_T_FRAME);
    };
    if (t == 0) { c_frames = chain->n_frames; c_last = chain->last_t; c_slot = c_frames > 0 ? frames[c_frames - 1].slot : -1; }
    __syncthreads();
    for (int base = 0; base < n_cand; base += 1024) {
        Ent mine[4];
        int nv = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = base + 4 * t + q;
            if (i < n_cand) {
                const FineOut* f = out + i;                             // written by other CTAs: read from L2
                const float cir = __ldcg(&f->cir_ratio);
                if (cir >= SYNC_CIR_MIN) {
                    Ent e;
                    e.t = __ldcg(&f->t_f) + range_off;
                    e.cfo = __ldcg(&f->cfo_hz); e.cir = cir;
                    mine[nv++] = e;
                }
            }
        }
        int n_valid;
        const int at = excl_scan(nv, n_valid);
#pragma unroll
        for (int q = 0; q < 4; q++) if (q < nv) sE[at + q] = mine[q];
        __syncthreads();
        const int nf0 = c_frames, slot0 = c_slot;
        const long long last0 = c_last;
        // (2) the parallel case
        int ok = 1;
        for (int e = t; e < n_valid; e += 256) {
            const bool has_prev = e > 0 || nf0 > 0;
            const long long prev = e > 0 ? sE[e - 1].t : last0;
            if (has_prev && sE[e].t - prev < (long long)(DAB_T_FRAME * 9 / 10)) ok = 0;
        }
        if (__syncthreads_and(ok)) {
            if (t == 0) s_carry = slot0;
            __syncthreads();
            for (int e0 = 0; e0 < n_valid; e0 += 256) {
                const int e = e0 + t;
                int inc = 0;
                if (e < n_valid) {
                    const bool has_prev = e > 0 || nf0 > 0;
                    inc = has_prev ? frames_between(sE[e].t - (e > 0 ? sE[e - 1].t : last0)) : 1;       // the first frame of the call gets slot 0
                }
                int tot;
                const int before = excl_scan(inc, tot);
                if (e < n_valid) sSlot[e] = s_carry + before + inc;
                __syncthreads();
                if (t == 0) s_carry += tot;
                __syncthreads();
            }
        } else {
            // (3) sequential walk
            if (t == 0) {
                int nf = nf0, slot_prev = slot0;
                long long last = last0;
                for (int e = 0; e < n_valid; e++) {
                    const long long t_abs = sE[e].t;
                    if (nf > 0 && t_abs - last < (long long)(DAB_T_FRAME * 9 / 10)) { sSlot[e] = -1; continue; }   // seen already / too close
                    const int slot = nf == 0 ? 0 : slot_prev + frames_between(t_abs - last);
                    sSlot[e] = slot;
                    nf++; slot_prev = slot; last = t_abs;
                }
            }
            __syncthreads();
        }
        // frames = entries with a slot below max_slots (slots ascend, so what is cut off is a tail), written out in order
        int acc = 0;
        for (int e0 = 0; e0 < n_valid; e0 += 256) {
            const int e = e0 + t;
            const int is_frame = (e < n_valid && sSlot[e] >= 0 && sSlot[e] < max_slots) ? 1 : 0;
            int tot;
            const int before = excl_scan(is_frame, tot);
            if (is_frame) {
                FrameInfo fi;
                fi.t_f = sE[e].t; fi.cfo_hz = sE[e].cfo; fi.slot = sSlot[e];
                frames[nf0 + acc + before] = fi;
                frame_cir[nf0 + acc + before] = sE[e].cir;
            }
            // the last frame of the chunk carries the chain on
            if (is_frame && before == tot - 1) { c_last = sE[e].t; c_slot = sSlot[e]; }
            acc += tot;
            __syncthreads();
        }
        if (t == 0) c_frames = nf0 + acc;
        __syncthreads();
    }
    if (t == 0) {
        if (range_idx == 0) chain->range_first[0] = 0;
        if (range_idx < CHAIN_MAX_RANGES) chain->range_first[range_idx + 1] = c_frames;
        chain->n_frames = c_frames;
        chain->last_t = c_last;
        chain->n_candidates += n_cand;
    }
}
