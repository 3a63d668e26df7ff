// k_viterbi.cuh -- time de-interleave + de-puncture (gather) and K=7 rate-1/4 Viterbi decoding with full traceback.
//
// Replaces the reference core's time de-interleaver, de-puncturer, float Viterbi decoder and energy-dispersal
// descrambler (code: polynomials 133,171,145,133, puncturing vectors PI 1..24 + tail,
// 16-CIF delay table {0,8,4,12,...}).
//
// Layout: codewords of one sub-channel (one per CIF) are decoded 32 at a time, one codeword per thread, so every
// global access is a coalesced 128/256-byte row:
//   vin [group][step][32 lanes] uint32 : the 4 de-punctured int8 soft bits of that trellis step (0 = punctured)
//   dec [group][step][32 lanes] uint2  : the 64 survivor decisions of that step
// Path metrics are int16, two states per 32-bit register ((j, j+16) and (j+32, j+48)), updated with the packed
// VIADD.16x2 / VIMNMX.S16x2 instructions; VIMNMX delivers both decision predicates of a register pair for free.
#pragma once
#include "common.cuh"

struct VitJob {
    long long soft_base;   // byte offset of column 0 of row 0 inside a frame's soft block
    long long vin_off;     // uint32 index of this job's first group in vin
    long long dec_off;     // uint2 index in dec
    long long out_off;     // byte offset in out
    int row_pitch;         // bytes between the (up to 4) rows inside one frame's soft block
    int use_tdi;           // 1: bit p comes from row r + delay[p & 15]
    int row0;              // absolute row (CIF / FIC vector index) of codeword 0
    int max_row;           // last absolute row that exists in the soft buffer
    int n_rows;            // number of codewords
    int steps;             // trellis steps = n_info + 6
    int n_info;
    int stepinfo_off;      // index into the stepinfo table; entry [steps] is the sentinel (total punctured bits << 4)
    int out_pitch;         // bytes per decoded row = n_info / 8
    int kind;              // 0 FIC, 1 MSC
    int pad_[2];
};
struct VitGroup { int job; int g_in_job; };

__constant__ int c_tdi_delay[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

#define GATHER_STEPS 128
#define GATHER_PITCHW 131        /* odd: lanes reading the same column of consecutive rows hit distinct banks */

// grid (n_groups, ceil(max_steps / 64)), 256 threads
__global__ void __launch_bounds__(256) k_gather(const int8_t* __restrict__ soft, const VitJob* __restrict__ jobs,
                                                const VitGroup* __restrict__ groups, const uint32_t* __restrict__ stepinfo,
                                                uint32_t* __restrict__ vin)
{
    __shared__ uint32_t tile[47 * GATHER_PITCHW];
    __shared__ uint2 soff[GATHER_STEPS];
    const VitGroup grp = groups[blockIdx.x];
    const VitJob job = jobs[grp.job];
    const int s0 = blockIdx.y * GATHER_STEPS;
    if (s0 >= job.steps) return;
    const int s1 = min(s0 + GATHER_STEPS, job.steps);
    const uint32_t* si = stepinfo + job.stepinfo_off;
    const int p0 = (int)(__ldg(si + s0) >> 4), p1 = (int)(__ldg(si + s1) >> 4);
    const int p0a = p0 & ~3;
    const int nwords = (p1 - p0a + 3) >> 2;              // <= 129
    const int nrows = job.use_tdi ? 47 : 32;
    const int rbase = job.row0 + grp.g_in_job * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < nrows; row += 8) {        // one warp per row: coalesced 128-byte pieces, no index division
        int r = rbase + row;
        if (r > job.max_row) r = job.max_row;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(soft + (size_t)(r >> 2) * DAB_FRAME_SOFT + (size_t)job.soft_base +
                                                                (size_t)(r & 3) * job.row_pitch + p0a);
        for (int w = lane; w < nwords; w += 32) tile[row * GATHER_PITCHW + w] = __ldg(src + w);
    }
    // every row gets a zero word in its last (otherwise unused) column: punctured bits read it, so the inner loop is branch-free
    if (threadIdx.x < nrows) tile[threadIdx.x * GATHER_PITCHW + (GATHER_PITCHW - 1)] = 0;
    // byte offsets (relative to the lane's own row) of the four soft bits of every step of this chunk, computed once
    for (int i = threadIdx.x; i < (s1 - s0); i += 256) {
        const uint32_t info = __ldg(si + s0 + i);
        int p = (int)(info >> 4);
        unsigned short o4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (info & (1u << q)) {
                const int d = job.use_tdi ? c_tdi_delay[p & 15] : 0;
                o4[q] = (unsigned short)(d * (GATHER_PITCHW * 4) + (p - p0a));
                p++;
            } else o4[q] = (unsigned short)((GATHER_PITCHW - 1) * 4);
        }
        soff[i] = make_uint2((uint32_t)o4[0] | ((uint32_t)o4[1] << 16), (uint32_t)o4[2] | ((uint32_t)o4[3] << 16));
    }
    __syncthreads();
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile) + lane * (GATHER_PITCHW * 4);
    uint32_t* out = vin + job.vin_off + ((size_t)grp.g_in_job * job.steps) * 32 + lane;
    for (int s = s0 + warp; s < s1; s += 8) {
        const uint2 o = soff[s - s0];                                   // same for all lanes: broadcast
        const uint32_t b0 = tb[o.x & 0xffffu], b1 = tb[o.x >> 16], b2 = tb[o.y & 0xffffu], b3 = tb[o.y >> 16];
        out[(size_t)s * 32] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
}

// ------------------------------------------------------------------------------------------------ ACS
__host__ __device__ constexpr int vit_parity7(int x) { return (x ^ (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4) ^ (x >> 5) ^ (x >> 6)) & 1; }
// code bit q (0..2; bit 3 repeats bit 0) for state j (newest input = LSB) and input 0; polynomials bit-reversed so that
// bit d of the window (state << 1 | input) is a[i-d]
__host__ __device__ constexpr int vit_cbit(int j, int q) { return vit_parity7((j << 1) & (q == 0 ? 0155 : q == 1 ? 0117 : 0123)); }

template <int J>
__device__ __forceinline__ void vit_bfly(const uint32_t (&L)[16], const uint32_t (&H)[16], uint32_t (&E)[16], uint32_t (&O)[16],
                                         const uint32_t (&BP)[8], uint32_t (&dl)[4], uint32_t (&dh)[4])
{
    // low halves: butterfly of states (J, J+32) -> (2J, 2J+1); high halves: (J+16, J+48) -> (2J+32, 2J+33)
    uint32_t& dlo = dl[J & 3];          // four partial decision words: independent chains instead of one serial one
    uint32_t& dhi = dh[J & 3];
    constexpr int k = vit_cbit(J, 0) * 4 + vit_cbit(J, 1) * 2 + vit_cbit(J, 2);   // BP[k] = +beta, BP[7-k] = -beta
    bool ph, pl;
    const uint32_t t0 = L[J] + BP[k], t1 = H[J] + BP[7 - k];       // plain 32-bit adds: see the bias note in k_viterbi
    E[J] = __vibmax_s16x2(t0, t1, &ph, &pl);          // predicate = (t0 >= t1): survivor from the lower predecessor
    if (!pl) dlo |= 1u << (2 * J);
    if (!ph) dhi |= 1u << (2 * J);
    const uint32_t u0 = L[J] + BP[7 - k], u1 = H[J] + BP[k];
    O[J] = __vibmax_s16x2(u0, u1, &ph, &pl);
    if (!pl) dlo |= 1u << (2 * J + 1);
    if (!ph) dhi |= 1u << (2 * J + 1);
}
template <int J> struct VitUnroll {
    static __device__ __forceinline__ void run(const uint32_t (&L)[16], const uint32_t (&H)[16], uint32_t (&E)[16], uint32_t (&O)[16],
                                               const uint32_t (&BP)[8], uint32_t (&dl)[4], uint32_t (&dh)[4])
    {
        VitUnroll<J - 1>::run(L, H, E, O, BP, dl, dh);
        vit_bfly<J>(L, H, E, O, BP, dl, dh);
    }
};
template <> struct VitUnroll<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[16], const uint32_t (&)[16], uint32_t (&)[16], uint32_t (&)[16],
                                               const uint32_t (&)[8], uint32_t (&)[4], uint32_t (&)[4]) {}
};

// ---- backward butterfly: packed inputs A[J] = (b'[2J], b'[2J+32]), B[J] = (b'[2J+1], b'[2J+33]) (metrics of the NEXT trellis stage),
//      outputs X[J] = (b[J], b[J+16]), Y[J] = (b[J+32], b[J+48]); decision bit = chosen input bit of that state
template <int J>
__device__ __forceinline__ void vit_bfly_bwd(const uint32_t (&A)[16], const uint32_t (&B)[16], uint32_t (&X)[16], uint32_t (&Y)[16],
                                             const uint32_t (&BP)[8], uint32_t (&dl)[4], uint32_t (&dh)[4])
{
    uint32_t& dlo = dl[J & 3];
    uint32_t& dhi = dh[J & 3];
    constexpr int k = vit_cbit(J, 0) * 4 + vit_cbit(J, 1) * 2 + vit_cbit(J, 2);
    bool ph, pl;
    const uint32_t t0 = A[J] + BP[k], t1 = B[J] + BP[7 - k];         // state J / J+16: input 0 -> +beta, input 1 -> -beta
    X[J] = __vibmax_s16x2(t0, t1, &ph, &pl);
    if (!pl) dlo |= 1u << (2 * J);                                    // same bit pattern as the forward butterfly (the compiler
    if (!ph) dhi |= 1u << (2 * J);                                    // turns it into predicated adds); decoded in the trace below
    const uint32_t u0 = A[J] + BP[7 - k], u1 = B[J] + BP[k];         // state J+32 / J+48: all code bits flipped
    Y[J] = __vibmax_s16x2(u0, u1, &ph, &pl);
    if (!pl) dlo |= 1u << (2 * J + 1);
    if (!ph) dhi |= 1u << (2 * J + 1);
}
template <int J> struct VitUnrollBwd {
    static __device__ __forceinline__ void run(const uint32_t (&A)[16], const uint32_t (&B)[16], uint32_t (&X)[16], uint32_t (&Y)[16],
                                               const uint32_t (&BP)[8], uint32_t (&dl)[4], uint32_t (&dh)[4])
    {
        VitUnrollBwd<J - 1>::run(A, B, X, Y, BP, dl, dh);
        vit_bfly_bwd<J>(A, B, X, Y, BP, dl, dh);
    }
};
template <> struct VitUnrollBwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[16], const uint32_t (&)[16], uint32_t (&)[16], uint32_t (&)[16],
                                               const uint32_t (&)[8], uint32_t (&)[4], uint32_t (&)[4]) {}
};

__device__ __forceinline__ void vit_branch_metrics(uint32_t w, uint32_t (&BP)[8])
{
    const int a0 = (int8_t)(w), a1 = (int8_t)(w >> 8), a2 = (int8_t)(w >> 16), a3 = (int8_t)(w >> 24);
    const int T = a0 + a3, X1 = a1 + a2, X2 = a1 - a2;
    // BP[k], k = c0*4 + c1*2 + c2 (1 = code bit 1 = negative contribution): low half  (1-2c0) T + x,
    //                                                                      high half -(1-2c0) T + x  (state + 16 flips c0, c3)
    const int xs[4] = {X1, X2, -X2, -X1};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        BP[q] = (uint32_t)((-T + xs[q]) * 65536 + (T + xs[q]));
        BP[4 + q] = (uint32_t)((T + xs[q]) * 65536 + (-T + xs[q]));
    }
}

#define VIT_BIAS 16384u
#define VIT_PENV ((VIT_BIAS + VIT_START_PENALTY) * 65537u)

struct VitSmem {
    uint32_t ring[2][2][VIT_NORM_PERIOD][32];      // [direction][buffer][step][lane] soft words
    unsigned short meet[2][64][32];                // [direction][state][lane] path metrics at the meeting stage
};

// Bidirectional decoding: the code is zero-terminated at BOTH ends, so warp 0 runs the trellis forward over steps [0, m) from
// state 0 while warp 1 runs it backward over steps [m, T) from state 0; at stage m the state maximising alpha + beta lies on
// the maximum-likelihood path, warp 0 traces back from it and warp 1 traces forward.  Same path as one-directional Viterbi
// whenever the best path is unique; the tie rules (include/b200dab_spec.h) are mirrored by the CPU checker.  This halves the
// latency of the longest codeword, which is what bounds the kernel, with no per-step synchronisation between the warps.
//   decisions: dec[group][step][lane] uint2, bit s of (x | y << 32):
//     steps <  m : 1 = the survivor into state s came from predecessor (s >> 1) + 32
//     steps >= m : 1 = the best continuation from state s takes input bit 1; here state s sits at word (s >> 4) & 1,
//                  bit 2 (s & 15) + (s >> 5)
// one 64-thread CTA per group of 32 codewords; grid = n_groups (longest groups first)
__global__ void __launch_bounds__(64) k_viterbi(const uint32_t* __restrict__ vin, uint2* __restrict__ dec, const VitJob* __restrict__ jobs,
                                                const VitGroup* __restrict__ groups, const uint32_t* __restrict__ prbs_w,
                                                uint8_t* __restrict__ out)
{
    __shared__ VitSmem sm;
    const VitGroup grp = groups[blockIdx.x];
    const VitJob job = jobs[grp.job];
    const int lane = threadIdx.x & 31, dir = threadIdx.x >> 5;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;                         // meeting stage, multiple of 32
    const uint32_t* in = vin + job.vin_off + ((size_t)grp.g_in_job * steps) * 32 + lane;
    uint2* d = dec + job.dec_off + ((size_t)grp.g_in_job * steps) * 32 + lane;
    constexpr int PF = VIT_NORM_PERIOD;

    if (dir == 0) {
        // ------------------------------------------------------------ forward over steps [0, m)
        uint32_t L[16], H[16], E[16], O[16];
#pragma unroll
        for (int j = 0; j < 16; j++) { L[j] = VIT_PENV; H[j] = VIT_PENV; }
        L[0] = ((VIT_BIAS + VIT_START_PENALTY) << 16) | VIT_BIAS;   // state 0 starts at 0, state 16 at the penalty
        auto prefetch = [&](int buf, int t0) {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.ring[0][buf][i][lane]);
                const uint32_t* src = in + (size_t)min(t0 + i, steps - 1) * 32;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        prefetch(0, 0);
        int buf = 0;
        for (int t0 = 0; t0 < m; t0 += PF, buf ^= 1) {
            if (t0 + PF < m) { prefetch(buf ^ 1, t0 + PF); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            {   // re-reference all metrics to state 0 (keeps the int16 range; constants do not change any decision)
                const uint32_t c = (uint32_t)(((int)VIT_BIAS - (int)(L[0] & 0xffffu)) * 65537);
#pragma unroll
                for (int j = 0; j < 16; j++) { L[j] += c; H[j] += c; }
            }
            const int nst = min(PF, m - t0);
#pragma unroll 2
            for (int i = 0; i < nst; i++) {
                uint32_t BP[8];
                vit_branch_metrics(sm.ring[0][buf][i][lane], BP);
                uint32_t dl[4] = {0, 0, 0, 0}, dh[4] = {0, 0, 0, 0};
                VitUnroll<15>::run(L, H, E, O, BP, dl, dh);
                d[(size_t)(t0 + i) * 32] = make_uint2((dl[0] | dl[1]) | (dl[2] | dl[3]), (dh[0] | dh[1]) | (dh[2] | dh[3]));
                // re-pair: L'[q] = (new[q], new[q+16]), H'[q] = (new[q+32], new[q+48])
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    L[2 * j] = __byte_perm(E[j], E[j + 8], 0x5410);     H[2 * j] = __byte_perm(E[j], E[j + 8], 0x7632);
                    L[2 * j + 1] = __byte_perm(O[j], O[j + 8], 0x5410); H[2 * j + 1] = __byte_perm(O[j], O[j + 8], 0x7632);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            sm.meet[0][j][lane] = (unsigned short)(L[j] & 0xffffu);      sm.meet[0][j + 16][lane] = (unsigned short)(L[j] >> 16);
            sm.meet[0][j + 32][lane] = (unsigned short)(H[j] & 0xffffu); sm.meet[0][j + 48][lane] = (unsigned short)(H[j] >> 16);
        }
    } else {
        // ------------------------------------------------------------ backward over steps [m, T), T-1 first
        uint32_t A[16], B[16], X[16], Y[16];
#pragma unroll
        for (int j = 0; j < 16; j++) { A[j] = VIT_PENV; B[j] = VIT_PENV; }
        A[0] = ((VIT_BIAS + VIT_START_PENALTY) << 16) | VIT_BIAS;   // must end in state 0: A[0] = (b[0], b[32])
        auto prefetch = [&](int buf, int t_hi) {                    // ring slot i holds step t_hi - i
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.ring[1][buf][i][lane]);
                const uint32_t* src = in + (size_t)max(t_hi - i, 0) * 32;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        prefetch(0, steps - 1);
        int buf = 0;
        for (int t_hi = steps - 1; t_hi >= m; t_hi -= PF, buf ^= 1) {
            if (t_hi - PF >= m) { prefetch(buf ^ 1, t_hi - PF); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            {
                const uint32_t c = (uint32_t)(((int)VIT_BIAS - (int)(A[0] & 0xffffu)) * 65537);
#pragma unroll
                for (int j = 0; j < 16; j++) { A[j] += c; B[j] += c; }
            }
            const int nst = min(PF, t_hi - m + 1);
#pragma unroll 2
            for (int i = 0; i < nst; i++) {
                uint32_t BP[8];
                vit_branch_metrics(sm.ring[1][buf][i][lane], BP);
                uint32_t dl[4] = {0, 0, 0, 0}, dh[4] = {0, 0, 0, 0};
                VitUnrollBwd<15>::run(A, B, X, Y, BP, dl, dh);
                d[(size_t)(t_hi - i) * 32] = make_uint2((dl[0] | dl[1]) | (dl[2] | dl[3]), (dh[0] | dh[1]) | (dh[2] | dh[3]));
                // re-pair: A'[q] = (b[2q], b[2q+32]), B'[q] = (b[2q+1], b[2q+33]) from X[q] = (b[q], b[q+16]), Y[q] = (b[q+32], b[q+48])
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    A[j] = __byte_perm(X[2 * j], Y[2 * j], 0x5410);         B[j] = __byte_perm(X[2 * j + 1], Y[2 * j + 1], 0x5410);
                    A[j + 8] = __byte_perm(X[2 * j], Y[2 * j], 0x7632);     B[j + 8] = __byte_perm(X[2 * j + 1], Y[2 * j + 1], 0x7632);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            sm.meet[1][2 * j][lane] = (unsigned short)(A[j] & 0xffffu);     sm.meet[1][2 * j + 32][lane] = (unsigned short)(A[j] >> 16);
            sm.meet[1][2 * j + 1][lane] = (unsigned short)(B[j] & 0xffffu); sm.meet[1][2 * j + 33][lane] = (unsigned short)(B[j] >> 16);
        }
    }
    __syncthreads();
    // ---- meeting state: maximum of alpha + beta, lowest state on ties
    int s = 0, best = -1;
#pragma unroll 8
    for (int q = 0; q < 64; q++) {
        const int v = (int)sm.meet[0][q][lane] + (int)sm.meet[1][q][lane];
        if (v > best) { best = v; s = q; }
    }
    const int row = grp.g_in_job * 32 + lane;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + job.out_off + (size_t)row * job.out_pitch);
    const bool live = row < job.n_rows;
    if (dir == 0) {
        // traceback over steps m-1 .. 0
        for (int c = m / 32 - 1; c >= 0; c--) {
            uint32_t word = 0;
#pragma unroll 8
            for (int i = 31; i >= 0; i--) {
                const uint2 dd = d[(size_t)(32 * c + i) * 32];
                word |= (uint32_t)(s & 1) << (31 - i);
                const uint32_t wsel = (s & 32) ? dd.y : dd.x;
                const int bit = (wsel >> (s & 31)) & 1;
                s = (s >> 1) | (bit << 5);
            }
            if (live) orow[c] = __byte_perm(word, 0, 0x0123) ^ __ldg(prbs_w + c);
        }
    } else {
        // trace forward over steps m .. n_info-1 (the six tail steps carry no information)
        for (int c = m / 32; c < job.n_info / 32; c++) {
            uint32_t word = 0;
#pragma unroll 8
            for (int i = 0; i < 32; i++) {
                const uint2 dd = d[(size_t)(32 * c + i) * 32];
                const uint32_t wsel = (s & 16) ? dd.y : dd.x;
                const int bit = (wsel >> (2 * (s & 15) + (s >> 5))) & 1;
                word |= (uint32_t)bit << (31 - i);
                s = ((s << 1) | bit) & 63;
            }
            if (live) orow[c] = __byte_perm(word, 0, 0x0123) ^ __ldg(prbs_w + c);
        }
    }
}
