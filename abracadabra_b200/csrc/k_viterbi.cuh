// k_viterbi.cuh -- time de-interleave + de-puncture (gather) and K=7 rate-1/4 Viterbi decoding with full traceback.
//
// Replaces the reference core's time de-interleaver, de-puncturer, float Viterbi decoder and energy-dispersal
// descrambler (code: polynomials 133,171,145,133, puncturing vectors PI 1..24 + tail,
// 16-CIF delay table {0,8,4,12,...}).
//
// Layout: codewords of one sub-channel (one per CIF) are handled in groups of 32, so every global access is a coalesced row:
//   vin [group][step][32 codewords] uint32 : the 4 de-punctured int8 soft bits of that trellis step (0 = punctured)
//   dec [group][step][32 codewords] uint2  : the 64 survivor decisions of that step (x: lane role 0, y: lane role 1)
// Path metrics are int16, two states per 32-bit register, updated with packed 16x2 adds and VIMNMX.S16x2; VIMNMX delivers
// both decision predicates of a register pair for free.  Each codeword is spread over TWO lanes (32 states each), and each
// direction of the bidirectional recursion is its own one-warp CTA, so one recording yields ~4000 short work items that the
// hardware deals to whichever SM sub-partition frees up first.
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
__host__ __device__ constexpr int vit_kidx(int j) { return vit_cbit(j, 0) * 4 + vit_cbit(j, 1) * 2 + vit_cbit(j, 2); }

// Two lanes per codeword.  Role 0 owns the butterflies whose index has bit 3 clear; role 1 owns the others but works in
// COMPLEMENTED coordinates (physical state = logical state ^ 63, physical input = logical input ^ 1): the trellis is invariant
// under that map, so both roles run the same code on physical butterflies 0..7, both keep the results whose next butterfly
// index is < 8 and hand the rest to the partner with one SHFL.xor each, and a received state index is simply the complement
// of the sender's.  Complementing the 7-bit window flips the code bits of the odd-weight polynomials (133, 171) and leaves
// the even-weight one (145) alone: role 1 negates soft bits 0, 1 and 3.  The tie rules (include/b200dab_spec.h) are
// invariant under the same map, which is what makes the two roles agree with the scalar formulation bit for bit.
//
// forward registers : L[i] = (a[i], a[i+16]), H[i] = (a[i+32], a[i+48]), i = 0..7 (physical states)
//   butterfly i: E[i] = (a'[2i], a'[2i+32]), O[i] = (a'[2i+1], a'[2i+33]); decision word bit (n & 15) + 16 (n >> 5) for
//   physical new state n, 1 = survivor from physical predecessor (n >> 1) + 32
// Decision bits are collected with predicated adds of a constant.  The ALU pipe (VIMNMX, PRMT, IADD3) is the busy one in this
// kernel and issues one warp instruction per two cycles, the FMA pipe (IMAD) likewise: half of the bit-sets are therefore
// written as a multiply-add by a register holding 1, which ptxas has to keep on the FMA pipe.
__device__ __forceinline__ void vit_setbit_alu(uint32_t& d, bool p, uint32_t bit) { if (p) d |= bit; }
__device__ __forceinline__ void vit_setbit_fma(uint32_t& d, bool p, uint32_t bit, uint32_t one)
{
    if (p) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(one), "r"(bit));
}

template <int I>
__device__ __forceinline__ void vit_bfly(const uint32_t (&L)[8], const uint32_t (&H)[8], uint32_t (&E)[8], uint32_t (&O)[8],
                                         const uint32_t (&BP)[8], uint32_t (&d)[4], uint32_t one)
{
    uint32_t& dd = d[I & 3];              // four partial decision words: independent chains instead of one serial one
    constexpr int k = vit_kidx(I);        // BP[k] = +beta, BP[7-k] = -beta
    bool ph, pl;
    const uint32_t t0 = L[I] + BP[k], t1 = H[I] + BP[7 - k];       // plain 32-bit adds: see the bias note in k_viterbi
    E[I] = __vibmax_s16x2(t0, t1, &ph, &pl);          // even new state: the lower predecessor survives a tie
    vit_setbit_alu(dd, !pl, 1u << (2 * I));
    vit_setbit_fma(dd, !ph, 1u << (2 * I + 16), one);
    const uint32_t u0 = L[I] + BP[7 - k], u1 = H[I] + BP[k];
    O[I] = __vibmax_s16x2(u1, u0, &ph, &pl);          // odd new state: the upper predecessor survives a tie
    vit_setbit_alu(dd, pl, 1u << (2 * I + 1));
    vit_setbit_fma(dd, ph, 1u << (2 * I + 17), one);
}
template <int I> struct VitUnroll {
    static __device__ __forceinline__ void run(const uint32_t (&L)[8], const uint32_t (&H)[8], uint32_t (&E)[8], uint32_t (&O)[8],
                                               const uint32_t (&BP)[8], uint32_t (&d)[4], uint32_t one)
    {
        VitUnroll<I - 1>::run(L, H, E, O, BP, d, one);
        vit_bfly<I>(L, H, E, O, BP, d, one);
    }
};
template <> struct VitUnroll<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[8], const uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8],
                                               const uint32_t (&)[8], uint32_t (&)[4], uint32_t) {}
};

// backward registers: A[i] = (b'[2i], b'[2i+32]), B[i] = (b'[2i+1], b'[2i+33]) (metrics of the NEXT trellis stage), outputs
//   X[i] = (b[i], b[i+16]), Y[i] = (b[i+32], b[i+48]); decision word bit 2 (s & 7) + (s >> 5) + 16 ((s >> 4) & 1) for
//   physical state s (the forward pattern again: ptxas turns other layouts into a slower P2R sequence), 1 = the best
//   continuation takes physical input bit 1
template <int I>
__device__ __forceinline__ void vit_bfly_bwd(const uint32_t (&A)[8], const uint32_t (&B)[8], uint32_t (&X)[8], uint32_t (&Y)[8],
                                             const uint32_t (&BP)[8], uint32_t (&d)[4], uint32_t one)
{
    uint32_t& dd = d[I & 3];
    constexpr int k = vit_kidx(I);
    bool ph, pl;
    const uint32_t t0 = A[I] + BP[k], t1 = B[I] + BP[7 - k];         // states i, i+16: input 0 -> +beta; input 0 wins a tie
    X[I] = __vibmax_s16x2(t0, t1, &ph, &pl);
    vit_setbit_alu(dd, !pl, 1u << (2 * I));
    vit_setbit_fma(dd, !ph, 1u << (2 * I + 16), one);
    const uint32_t u0 = A[I] + BP[7 - k], u1 = B[I] + BP[k];         // states i+32, i+48: code bits flipped; input 1 wins a tie
    Y[I] = __vibmax_s16x2(u1, u0, &ph, &pl);
    vit_setbit_alu(dd, pl, 1u << (2 * I + 1));
    vit_setbit_fma(dd, ph, 1u << (2 * I + 17), one);
}
template <int I> struct VitUnrollBwd {
    static __device__ __forceinline__ void run(const uint32_t (&A)[8], const uint32_t (&B)[8], uint32_t (&X)[8], uint32_t (&Y)[8],
                                               const uint32_t (&BP)[8], uint32_t (&d)[4], uint32_t one)
    {
        VitUnrollBwd<I - 1>::run(A, B, X, Y, BP, d, one);
        vit_bfly_bwd<I>(A, B, X, Y, BP, d, one);
    }
};
template <> struct VitUnrollBwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[8], const uint32_t (&)[8], uint32_t (&)[8], uint32_t (&)[8],
                                               const uint32_t (&)[8], uint32_t (&)[4], uint32_t) {}
};

// sg = +1 (role 0) or -1 (role 1)
__device__ __forceinline__ void vit_branch_metrics(uint32_t w, int sg, uint32_t (&BP)[8])
{
    const int a0 = (int8_t)(w), a1 = (int8_t)(w >> 8), a2 = (int8_t)(w >> 16), a3 = (int8_t)(w >> 24);
    const int T = (a0 + a3) * sg, X1 = a1 * sg + a2, X2 = a1 * sg - a2;
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
#define VIT_MEET_WORDS (2 * 64 * 32)        /* unsigned shorts per group: [direction][state][codeword] */

// Bidirectional decoding: the code is zero-terminated at BOTH ends, so one work item runs the trellis forward over steps
// [0, m) from state 0 while another runs it backward over steps [m, T) from state 0; at stage m the state maximising
// alpha + beta lies on the maximum-likelihood path (k_vit_trace picks it, traces back over [0, m) and forward over [m, T)).
// Same path as one-directional Viterbi whenever the best path is unique; the tie rules are mirrored by the CPU checker.
// Metrics carry a bias of 16384 so that both halves of a register stay non-negative and a packed add is one 32-bit add.
// grid = 4 * n_groups one-warp CTAs (longest groups first): item = (group, which 16 codewords, direction)
__global__ void __launch_bounds__(32) k_viterbi(const uint32_t* __restrict__ vin, uint32_t* __restrict__ dec, unsigned short* __restrict__ meet,
                                                const VitJob* __restrict__ jobs, const VitGroup* __restrict__ groups)
{
    constexpr int PF = VIT_NORM_PERIOD;
    __shared__ uint32_t ring[2][PF][32];
    const int item = blockIdx.x;
    const VitGroup grp = groups[item >> 2];
    const VitJob job = jobs[grp.job];
    const int lane = threadIdx.x, role = lane & 1, dir = item & 1;
    const int cw = ((item >> 1) & 1) * 16 + (lane >> 1);
    const int sg = role ? -1 : 1;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;                         // meeting stage, multiple of 32
    const uint32_t* in = vin + job.vin_off + ((size_t)grp.g_in_job * steps) * 32 + cw;
    uint32_t* d = dec + (job.dec_off + ((size_t)grp.g_in_job * steps) * 32 + cw) * 2 + role;
    unsigned short* mt = meet + ((size_t)(item >> 2) * 2 + dir) * (64 * 32) + cw;
    const uint32_t flip = role ? 63u : 0u;
    const uint32_t one = job.steps > 0 ? 1u : 0u;                   // always 1, but not to the compiler: see vit_setbit_fma

    if (dir == 0) {
        // ------------------------------------------------------------ forward over steps [0, m)
        uint32_t L[8], H[8], E[8], O[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { L[j] = VIT_PENV; H[j] = VIT_PENV; }
        if (role == 0) L[0] = ((VIT_BIAS + VIT_START_PENALTY) << 16) | VIT_BIAS;   // state 0 starts at 0, state 16 at the penalty
        auto prefetch = [&](int buf, int t0) {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[buf][i][lane]);
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
                uint32_t c = (uint32_t)(((int)VIT_BIAS - (int)(L[0] & 0xffffu)) * 65537);
                c = __shfl_sync(0xffffffffu, c, lane & ~1);
#pragma unroll
                for (int j = 0; j < 8; j++) { L[j] += c; H[j] += c; }
            }
            const int nst = min(PF, m - t0);
#pragma unroll 2
            for (int i = 0; i < nst; i++) {
                uint32_t BP[8];
                vit_branch_metrics(ring[buf][i][lane], sg, BP);
                uint32_t dw[4] = {0, 0, 0, 0};
                VitUnroll<7>::run(L, H, E, O, BP, dw, one);
                d[(size_t)(t0 + i) * 64] = (dw[0] | dw[1]) | (dw[2] | dw[3]);
                // E/O[0..3] stay (new states 0..7, 32..39); the partner's E/O[4..7] arrive complemented: its (8+2u+b, 40+2u+b)
                // are my (55-2u-b, 23-2u-b)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t re = __shfl_xor_sync(0xffffffffu, E[4 + u], 1), ro = __shfl_xor_sync(0xffffffffu, O[4 + u], 1);
                    const int je = 7 - 2 * u, jo = 6 - 2 * u;           // received E -> odd slot, received O -> even slot
                    const uint32_t own_o = O[je >> 1], own_e = E[jo >> 1];
                    L[je] = __byte_perm(own_o, re, 0x7610); H[je] = __byte_perm(own_o, re, 0x5432);
                    L[jo] = __byte_perm(own_e, ro, 0x7610); H[jo] = __byte_perm(own_e, ro, 0x5432);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            mt[((uint32_t)j ^ flip) * 32] = (unsigned short)(L[j] & 0xffffu);        mt[((uint32_t)(j + 16) ^ flip) * 32] = (unsigned short)(L[j] >> 16);
            mt[((uint32_t)(j + 32) ^ flip) * 32] = (unsigned short)(H[j] & 0xffffu); mt[((uint32_t)(j + 48) ^ flip) * 32] = (unsigned short)(H[j] >> 16);
        }
    } else {
        // ------------------------------------------------------------ backward over steps [m, T), T-1 first
        uint32_t A[8], B[8], X[8], Y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { A[j] = VIT_PENV; B[j] = VIT_PENV; }
        if (role == 0) A[0] = ((VIT_BIAS + VIT_START_PENALTY) << 16) | VIT_BIAS;   // must end in state 0: A[0] = (b[0], b[32])
        auto prefetch = [&](int buf, int t_hi) {                    // ring slot i holds step t_hi - i
#pragma unroll
            for (int i = 0; i < PF; i++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[buf][i][lane]);
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
                uint32_t c = (uint32_t)(((int)VIT_BIAS - (int)(A[0] & 0xffffu)) * 65537);
                c = __shfl_sync(0xffffffffu, c, lane & ~1);
#pragma unroll
                for (int j = 0; j < 8; j++) { A[j] += c; B[j] += c; }
            }
            const int nst = min(PF, t_hi - m + 1);
#pragma unroll 1
            for (int i = 0; i < nst; i++) {
                uint32_t BP[8];
                vit_branch_metrics(ring[buf][i][lane], sg, BP);
                uint32_t dw[4] = {0, 0, 0, 0};
                VitUnrollBwd<7>::run(A, B, X, Y, BP, dw, one);
                d[(size_t)(t_hi - i) * 64] = (dw[0] | dw[1]) | (dw[2] | dw[3]);
                // low halves (states 0..7, 32..39) stay; the partner's high halves (its 16+i, 48+i) are my 47-i, 15-i:
                // its packed (Y[i].hi, X[i].hi) is my (b[15-i], b[47-i]) = A[(15-i)/2] for odd i, B[(14-i)/2] for even i
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t so = __byte_perm(Y[7 - 2 * j], X[7 - 2 * j], 0x7632), se = __byte_perm(Y[6 - 2 * j], X[6 - 2 * j], 0x7632);
                    A[4 + j] = __shfl_xor_sync(0xffffffffu, so, 1);
                    B[4 + j] = __shfl_xor_sync(0xffffffffu, se, 1);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    A[j] = __byte_perm(X[2 * j], Y[2 * j], 0x5410);
                    B[j] = __byte_perm(X[2 * j + 1], Y[2 * j + 1], 0x5410);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            mt[((uint32_t)(2 * j) ^ flip) * 32] = (unsigned short)(A[j] & 0xffffu);     mt[((uint32_t)(2 * j + 32) ^ flip) * 32] = (unsigned short)(A[j] >> 16);
            mt[((uint32_t)(2 * j + 1) ^ flip) * 32] = (unsigned short)(B[j] & 0xffffu); mt[((uint32_t)(2 * j + 33) ^ flip) * 32] = (unsigned short)(B[j] >> 16);
        }
    }
}

// Meeting state (maximum of alpha + beta, lowest state on ties), traceback over [0, m) (warp 0) and trace forward over
// [m, n_info) (warp 1), energy-dispersal XOR on the way out.  grid = n_groups, 64 threads, one codeword per lane.
// With w = x | (~brev(y)) << 32 of a decision pair, the bit of logical state n is bit (n&15 | (n>>1)&16 | (n&16)<<1) of w for a
// forward step and bit ((s&7)<<1 | s>>5 | s&16 | (s&8)<<2) for a backward step (role 1 wrote complemented positions and values).
// The state recursion is a dependent chain, so the 32 decision pairs of the next chunk are fetched while this one is traced.
__device__ __forceinline__ void vit_load_chunk(uint2 (&b)[32], const uint2* d, int c)
{
#pragma unroll
    for (int i = 0; i < 32; i++) b[i] = __ldcs(d + (size_t)(32 * c + i) * 32);
}
__global__ void __launch_bounds__(64) k_vit_trace(const uint2* __restrict__ dec, const unsigned short* __restrict__ meet,
                                                  const VitJob* __restrict__ jobs, const VitGroup* __restrict__ groups,
                                                  const uint32_t* __restrict__ prbs_w, uint8_t* __restrict__ out)
{
    const VitGroup grp = groups[blockIdx.x];
    const VitJob job = jobs[grp.job];
    const int lane = threadIdx.x & 31, dir = threadIdx.x >> 5;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;
    const uint2* d = dec + job.dec_off + ((size_t)grp.g_in_job * steps) * 32 + lane;
    const int c_first = dir == 0 ? m / 32 - 1 : m / 32, c_end = dir == 0 ? -1 : job.n_info / 32, c_step = dir == 0 ? -1 : 1;
    uint2 cur[32], nxt[32];
    if (c_first != c_end) vit_load_chunk(cur, d, c_first);
    const unsigned short* mt = meet + (size_t)blockIdx.x * VIT_MEET_WORDS + lane;
    uint32_t s = 0;
    int best = -1;
#pragma unroll 8
    for (int q = 0; q < 64; q++) {
        const int v = (int)mt[q * 32] + (int)mt[(64 + q) * 32];
        if (v > best) { best = v; s = (uint32_t)q; }
    }
    const int row = grp.g_in_job * 32 + lane;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + job.out_off + (size_t)row * job.out_pitch);
    const bool live = row < job.n_rows;
    for (int c = c_first; c != c_end; c += c_step) {
        const bool more = (c + c_step) != c_end;
        if (more) vit_load_chunk(nxt, d, c + c_step);
        uint32_t word = 0;
        if (dir == 0) {
            // traceback: s is the state after step 32 c + i
#pragma unroll
            for (int i = 31; i >= 0; i--) {
                const unsigned long long w = (unsigned long long)cur[i].x | ((unsigned long long)(~__brev(cur[i].y)) << 32);
                word |= (s & 1u) << (31 - i);
                const uint32_t idx = (s & 15u) | ((s >> 1) & 16u) | ((s & 16u) << 1);
                s = (s >> 1) | (((uint32_t)(w >> idx) & 1u) << 5);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const unsigned long long w = (unsigned long long)cur[i].x | ((unsigned long long)(~__brev(cur[i].y)) << 32);
                const uint32_t idx = ((s & 7u) << 1) | (s >> 5) | (s & 16u) | ((s & 8u) << 2);
                const uint32_t bit = (uint32_t)(w >> idx) & 1u;
                word |= bit << (31 - i);
                s = ((s << 1) | bit) & 63u;
            }
        }
        if (live) orow[c] = __byte_perm(word, 0, 0x0123) ^ __ldg(prbs_w + c);
        if (more) {
#pragma unroll
            for (int i = 0; i < 32; i++) cur[i] = nxt[i];
        }
    }
}
