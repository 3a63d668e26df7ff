// common.cuh -- shared device helpers: sample loads, block reductions, the 2048-point shared-memory FFT.
//
// Part of the B200-native DAB mode I receive path that stands in for the reference's closed libdabsdr core
// (reference interface: lib/linux_x86_64/dabsdr.h:397-429; feed conversion: src/input/rawfileinput.cpp:643-711).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200dab_spec.h"

#define CUDA_CHECK_RET(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->set_error(#x, e_); return -1; } } while (0)

// ------------------------------------------------------------------------------------------------ sample access
// u8: float(x - 128), s16: float(x), f32: as is -- exactly the conversion RawFileWorker::run applies.
template <int FMT>
__device__ __forceinline__ float2 load_iq(const void* __restrict__ raw, int64_t n)
{
    if (FMT == B200DAB_FMT_U8) {
        uchar2 v = __ldg(reinterpret_cast<const uchar2*>(raw) + n);
        return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128));
    } else if (FMT == B200DAB_FMT_S16) {
        short2 v = __ldg(reinterpret_cast<const short2*>(raw) + n);
        return make_float2((float)v.x, (float)v.y);
    } else {
        return __ldg(reinterpret_cast<const float2*>(raw) + n);
    }
}

__device__ __forceinline__ float2 load_iq_rt(const void* __restrict__ raw, int fmt, int64_t n)
{
    if (fmt == B200DAB_FMT_U8) return load_iq<B200DAB_FMT_U8>(raw, n);
    if (fmt == B200DAB_FMT_S16) return load_iq<B200DAB_FMT_S16>(raw, n);
    return load_iq<B200DAB_FMT_F32>(raw, n);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by exp(-j pi/2 * q), q = 0..3
__device__ __forceinline__ float2 rot_mq(float2 a, int q)
{
    switch (q & 3) { case 0: return a; case 1: return make_float2(a.y, -a.x); case 2: return make_float2(-a.x, -a.y); default: return make_float2(-a.y, a.x); }
}

// ------------------------------------------------------------------------------------------------ frame list (device-resident)
struct FrameInfo {
    long long t_f;     // FFT window start of symbol 0 (PRS)
    float cfo_hz;
    int slot;          // frame slot in the soft-bit buffer (gaps = lost frames)
};
// Frames chained on the device (k_fine_sync, last CTA): what the host used to do between the fine synchronisation and the OFDM
// demodulation, so that no host round trip sits in the middle of the front end.  One state per decode call, ranges in ascending order.
#define CHAIN_MAX_RANGES 2048
struct ChainState {
    int n_frames;                 // frames chained so far (all ranges of this call)
    int n_candidates;             // candidates looked at so far
    long long last_t;             // FFT window start (absolute) of the last chained frame
    int range_first[CHAIN_MAX_RANGES + 1];   // frames [range_first[c], range_first[c + 1]) were found in range c
};
// frames [first, end) a launch works on: range_idx >= 0: the frames found in that range, < 0: all frames; then `skip` frames are
// left out at the front and at most `limit` frames (counted from the unskipped start) are taken
__device__ __forceinline__ void chain_range(const ChainState* __restrict__ chain, int range_idx, int skip, int limit, int& first, int& end)
{
    int a = 0, b = chain->n_frames;
    if (range_idx >= 0) { a = chain->range_first[range_idx]; b = chain->range_first[range_idx + 1]; }
    end = (limit < b - a) ? a + limit : b;
    first = a + skip;
}

// ------------------------------------------------------------------------------------------------ reductions (256-thread CTAs)
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; `red` is a shared float[32]; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red)
{
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = (l < nw) ? red[l] : 0.f;
    r = warp_sum(r);
    return r;
}

// ------------------------------------------------------------------------------------------------ FFT 2048 (256 threads)
// In-place decimation-in-frequency, radix 8 x 8 x 8 x 4, on a padded shared buffer of float2.
// Input in natural order at phys(n); output X[k] lands at phys(fft_pos(k)) -- the digit reversal is never undone,
// every consumer goes through fft_pos().
#define FFT_BUF_LEN 2304                                   // 2048 + 4 pad per 32
__device__ __host__ __forceinline__ int fft_phys(int p) { return p + ((p >> 5) << 2); }
__device__ __host__ __forceinline__ int fft_pos(int k) { return ((k & 7) << 8) | (((k >> 3) & 7) << 5) | (((k >> 6) & 7) << 2) | (k >> 9); }

__device__ __forceinline__ void radix8(float2 (&v)[8])
{
    const float h = 0.70710678118654752440f;
    // stage 1
    float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
    float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    // twiddles W8^0..3 on the difference branch: 1, (1-j)/sqrt2, -j, (-1-j)/sqrt2
    a5 = make_float2((a5.x + a5.y) * h, (a5.y - a5.x) * h);
    a6 = make_float2(a6.y, -a6.x);
    a7 = make_float2((a7.y - a7.x) * h, (-a7.x - a7.y) * h);
    // stage 2 (two radix-4 halves)
    float2 b0 = cadd(a0, a2), b2 = csub(a0, a2);
    float2 b1 = cadd(a1, a3), b3 = csub(a1, a3);
    b3 = make_float2(b3.y, -b3.x);
    float2 b4 = cadd(a4, a6), b6 = csub(a4, a6);
    float2 b5 = cadd(a5, a7), b7 = csub(a5, a7);
    b7 = make_float2(b7.y, -b7.x);
    // stage 3; outputs in natural frequency order k = 0..7
    v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
    v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
    v[1] = cadd(b4, b5); v[5] = csub(b4, b5);
    v[3] = cadd(b6, b7); v[7] = csub(b6, b7);
}

// tw: global table exp(-2 pi i k / 2048), k = 0..2047
__device__ __forceinline__ void fft2048_smem(float2* __restrict__ buf, const float2* __restrict__ tw)
{
    const int t = threadIdx.x;
    float2 v[8];
    // pass A: over n3 (stride 256), twiddle W2048^(t*k)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = buf[fft_phys(t + 256 * r)];
    radix8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], __ldg(tw + ((t * k) & 2047)));
#pragma unroll
    for (int k = 0; k < 8; k++) buf[fft_phys(t + 256 * k)] = v[k];
    __syncthreads();
    // pass B: inside each 256-block, stride 32, twiddle W256^(r*k)
    {
        const int blk = t >> 5, r = t & 31, base = blk << 8;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = buf[fft_phys(base + r + 32 * i)];
        radix8(v);
#pragma unroll
        for (int k = 1; k < 8; k++) v[k] = cmul(v[k], __ldg(tw + ((r * k) << 3)));
#pragma unroll
        for (int k = 0; k < 8; k++) buf[fft_phys(base + r + 32 * k)] = v[k];
    }
    __syncthreads();
    // pass C: inside each 32-block, stride 4, twiddle W32^(n0*k)
    {
        const int blk = t >> 2, n0 = t & 3, base = blk << 5;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = buf[fft_phys(base + n0 + 4 * i)];
        radix8(v);
#pragma unroll
        for (int k = 1; k < 8; k++) v[k] = cmul(v[k], __ldg(tw + ((n0 * k) << 6)));
#pragma unroll
        for (int k = 0; k < 8; k++) buf[fft_phys(base + n0 + 4 * k)] = v[k];
    }
    __syncthreads();
    // pass D: radix 4 on consecutive quadruples; two per thread
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int g = t + 256 * h;
        float2* p = buf + fft_phys(4 * g);
        float4 q0 = *reinterpret_cast<float4*>(p), q1 = *reinterpret_cast<float4*>(p + 2);
        float2 x0 = make_float2(q0.x, q0.y), x1 = make_float2(q0.z, q0.w), x2 = make_float2(q1.x, q1.y), x3 = make_float2(q1.z, q1.w);
        float2 s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = csub(x1, x3);
        d1 = make_float2(d1.y, -d1.x);
        float2 y0 = cadd(s0, s1), y2 = csub(s0, s1), y1 = cadd(d0, d1), y3 = csub(d0, d1);
        *reinterpret_cast<float4*>(p) = make_float4(y0.x, y0.y, y1.x, y1.y);
        *reinterpret_cast<float4*>(p + 2) = make_float4(y2.x, y2.y, y3.x, y3.y);
    }
    __syncthreads();
}

// Same transform with everything that only depends on the thread hoisted out by the caller: shared-memory base offsets
// (the padded layout makes every pass a base + constant stride) and the 21 twiddles of passes A/B/C held in registers.
struct FftPlan {
    int pA, pB, pC, pD0, pD1;
    float2 twA[7], twB[7], twC[7];
};
__device__ __forceinline__ void fft_plan_init(FftPlan& pl, const float2* __restrict__ tw)
{
    const int t = threadIdx.x;
    pl.pA = fft_phys(t);                                        // + 288 r
    pl.pB = fft_phys(((t >> 5) << 8) + (t & 31));               // + 36 i
    pl.pC = (t >> 2) * 36 + (t & 3);                            // + 4 i
    pl.pD0 = fft_phys(4 * t);
    pl.pD1 = fft_phys(4 * (t + 256));
#pragma unroll
    for (int k = 1; k < 8; k++) {
        pl.twA[k - 1] = __ldg(tw + ((t * k) & 2047));
        pl.twB[k - 1] = __ldg(tw + (((t & 31) * k) << 3));
        pl.twC[k - 1] = __ldg(tw + (((t & 3) * k) << 6));
    }
}
// passes A, B, C only: the caller finishes with its own radix-4 pass over the quadruples at pD0 / pD1.
// _regs variant: the thread already holds its eight pass-A inputs x[t + 256 r] (they are exactly the samples it loaded), so the
// input never goes through shared memory; the caller guarantees that nobody still reads `buf` (a barrier since the last use).
__device__ __forceinline__ void fft2048_planned_bc(float2* __restrict__ buf, const FftPlan& pl);
__device__ __forceinline__ void fft2048_planned_abc_regs(float2* __restrict__ buf, const FftPlan& pl, float2 (&v)[8])
{
    radix8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], pl.twA[k - 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) buf[pl.pA + 288 * k] = v[k];
    __syncthreads();
    fft2048_planned_bc(buf, pl);
}
__device__ __forceinline__ void fft2048_planned_abc(float2* __restrict__ buf, const FftPlan& pl)
{
    float2 v[8];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = buf[pl.pA + 288 * r];
    fft2048_planned_abc_regs(buf, pl, v);
}
__device__ __forceinline__ void fft2048_planned_bc(float2* __restrict__ buf, const FftPlan& pl)
{
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = buf[pl.pB + 36 * i];
    radix8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], pl.twB[k - 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) buf[pl.pB + 36 * k] = v[k];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = buf[pl.pC + 4 * i];
    radix8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], pl.twC[k - 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) buf[pl.pC + 4 * k] = v[k];
    __syncthreads();
}
__device__ __forceinline__ void fft2048_planned(float2* __restrict__ buf, const FftPlan& pl)
{
    fft2048_planned_abc(buf, pl);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        float2* p = buf + (h ? pl.pD1 : pl.pD0);
        float4 q0 = *reinterpret_cast<float4*>(p), q1 = *reinterpret_cast<float4*>(p + 2);
        float2 x0 = make_float2(q0.x, q0.y), x1 = make_float2(q0.z, q0.w), x2 = make_float2(q1.x, q1.y), x3 = make_float2(q1.z, q1.w);
        float2 s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = csub(x1, x3);
        d1 = make_float2(d1.y, -d1.x);
        float2 y0 = cadd(s0, s1), y2 = csub(s0, s1), y1 = cadd(d0, d1), y3 = csub(d0, d1);
        *reinterpret_cast<float4*>(p) = make_float4(y0.x, y0.y, y1.x, y1.y);
        *reinterpret_cast<float4*>(p + 2) = make_float4(y2.x, y2.y, y3.x, y3.y);
    }
    __syncthreads();
}

// carrier index c (0..1535, ascending k = -768..-1, 1..768) -> FFT bin
__device__ __host__ __forceinline__ int carrier_bin(int c) { return c < 768 ? c + 1280 : c - 767; }

// raw samples of one symbol held in registers between the (early) global load and the (late) conversion, so that the
// next symbol's loads are in flight while the current symbol is transformed
template <int FMT> struct RawSym;
template <> struct RawSym<B200DAB_FMT_U8> { uchar2 v[8]; };
template <> struct RawSym<B200DAB_FMT_S16> { short2 v[8]; };
template <> struct RawSym<B200DAB_FMT_F32> { float2 v[8]; };

template <int FMT>
__device__ __forceinline__ void fetch_symbol(RawSym<FMT>& r, const void* __restrict__ raw, int64_t nsamp, int64_t start)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t n = start + t + 256 * i;
        const bool ok = n >= 0 && n < nsamp;
        if (!ok) n = 0;
        if (FMT == B200DAB_FMT_U8) { uchar2 x = __ldg(reinterpret_cast<const uchar2*>(raw) + n); if (!ok) x = make_uchar2(128, 128); reinterpret_cast<uchar2*>(r.v)[i] = x; }
        else if (FMT == B200DAB_FMT_S16) { short2 x = __ldg(reinterpret_cast<const short2*>(raw) + n); if (!ok) x = make_short2(0, 0); reinterpret_cast<short2*>(r.v)[i] = x; }
        else { float2 x = __ldg(reinterpret_cast<const float2*>(raw) + n); if (!ok) x = make_float2(0.f, 0.f); reinterpret_cast<float2*>(r.v)[i] = x; }
    }
}

// whole symbol inside the buffer: no per-sample bounds checks
template <int FMT>
__device__ __forceinline__ void fetch_symbol_fast(RawSym<FMT>& r, const void* __restrict__ raw, int64_t start)
{
    const int64_t n0 = start + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (FMT == B200DAB_FMT_U8) reinterpret_cast<uchar2*>(r.v)[i] = __ldg(reinterpret_cast<const uchar2*>(raw) + n0 + 256 * i);
        else if (FMT == B200DAB_FMT_S16) reinterpret_cast<short2*>(r.v)[i] = __ldg(reinterpret_cast<const short2*>(raw) + n0 + 256 * i);
        else reinterpret_cast<float2*>(r.v)[i] = __ldg(reinterpret_cast<const float2*>(raw) + n0 + 256 * i);
    }
}

// de-rotation by a rotor recurrence: one sincos per thread and symbol, then w_{r+1} = w_r * step (step = rotation over 256 samples)
template <int FMT>
__device__ __forceinline__ void store_symbol_rotor(float2* __restrict__ buf, int pA, const RawSym<FMT>& r, float base_cyc, float cpp, float2 step)
{
    float cyc = base_cyc + cpp * (float)threadIdx.x;
    float ph = -6.283185307179586f * (cyc - rintf(cyc));
    float2 w;
    __sincosf(ph, &w.y, &w.x);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float2 x;
        if (FMT == B200DAB_FMT_U8) { uchar2 v = reinterpret_cast<const uchar2*>(r.v)[i]; x = make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
        else if (FMT == B200DAB_FMT_S16) { short2 v = reinterpret_cast<const short2*>(r.v)[i]; x = make_float2((float)v.x, (float)v.y); }
        else x = reinterpret_cast<const float2*>(r.v)[i];
        buf[pA + 288 * i] = cmul(x, w);
        w = cmul(w, step);
    }
}

// same conversion + de-rotation, result kept in registers: v[i] = x[t + 256 i] (the pass-A inputs of this thread)
template <int FMT>
__device__ __forceinline__ void symbol_rotor_regs(float2 (&out)[8], const RawSym<FMT>& r, float base_cyc, float cpp, float2 step)
{
    float cyc = base_cyc + cpp * (float)threadIdx.x;
    float ph = -6.283185307179586f * (cyc - rintf(cyc));
    float2 w;
    __sincosf(ph, &w.y, &w.x);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float2 x;
        if (FMT == B200DAB_FMT_U8) { uchar2 v = reinterpret_cast<const uchar2*>(r.v)[i]; x = make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
        else if (FMT == B200DAB_FMT_S16) { short2 v = reinterpret_cast<const short2*>(r.v)[i]; x = make_float2((float)v.x, (float)v.y); }
        else x = reinterpret_cast<const float2*>(r.v)[i];
        out[i] = cmul(x, w);
        w = cmul(w, step);
    }
}

template <int FMT>
__device__ __forceinline__ void store_symbol(float2* __restrict__ buf, const RawSym<FMT>& r, float base_cyc, float cpp)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int m = t + 256 * i;
        float2 x;
        if (FMT == B200DAB_FMT_U8) { uchar2 v = reinterpret_cast<const uchar2*>(r.v)[i]; x = make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
        else if (FMT == B200DAB_FMT_S16) { short2 v = reinterpret_cast<const short2*>(r.v)[i]; x = make_float2((float)v.x, (float)v.y); }
        else x = reinterpret_cast<const float2*>(r.v)[i];
        float cyc = base_cyc + cpp * (float)m;
        float ph = -6.283185307179586f * (cyc - rintf(cyc));
        float s, c;
        __sincosf(ph, &s, &c);
        buf[fft_phys(m)] = make_float2(x.x * c - x.y * s, x.x * s + x.y * c);
    }
}

// load 2048 samples starting at `start` into buf (natural order), de-rotated: x[m] * exp(-2 pi i (base_cyc + cpp * m))
template <int FMT>
__device__ __forceinline__ void load_symbol(float2* __restrict__ buf, const void* __restrict__ raw, int64_t nsamp, int64_t start,
                                            float base_cyc, float cpp)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int m = t + 256 * r;
        const int64_t n = start + m;
        float2 x = make_float2(0.f, 0.f);
        if (n >= 0 && n < nsamp) x = load_iq<FMT>(raw, n);
        float cyc = base_cyc + cpp * (float)m;
        float ph = -6.283185307179586f * (cyc - rintf(cyc));
        float s, c;
        __sincosf(ph, &s, &c);
        buf[fft_phys(m)] = make_float2(x.x * c - x.y * s, x.x * s + x.y * c);
    }
}
