// k_viterbi.cuh -- time de-interleave + de-puncture (gather) and K=7 rate-1/4 Viterbi decoding with full traceback.
//
// Replaces the reference core's time de-interleaver, de-puncturer, float Viterbi decoder and energy-dispersal
// descrambler (code: polynomials 133,171,145,133, puncturing vectors PI 1..24 + tail,
// 16-CIF delay table {0,8,4,12,...}).
//
// Layout: codewords of one sub-channel (one per CIF) are handled in groups of 32, so every global access is a coalesced row:
//   vin [group][step][32 codewords] uint32 : the 4 de-punctured int8 soft bits of that trellis step (0 = punctured)
//   dec [group][step][32 codewords] uint2  : the 64 survivor decisions of that step (x: lane role 0, y: lane role 1)
// Path metrics are int16, two codewords per 32-bit register, updated with packed 16x2 adds and VIMNMX.S16x2; VIMNMX delivers
// both decision predicates of a register pair for free.  A register holds the same state of TWO codewords, the 64 states of the
// pair are spread over TWO lanes (32 registers each), and each direction of the bidirectional recursion is its own one-warp CTA.
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

// One warp = one group of 32 codewords = 16 lane PAIRS; a pair works on TWO codewords at once, codeword p in the low and codeword
// 16 + p in the high 16 bits of every register, so a packed add / VIMNMX.S16x2 serves both and -- unlike packing two states of
// one codeword -- no register ever has to be taken apart again: where a survivor goes is pure register naming.
// The 64 states are split over the two lanes of the pair by their LSB.  Role 0 owns the even states; role 1 owns the odd ones but
// works in COMPLEMENTED coordinates (physical state = logical state ^ 63, physical input = logical input ^ 1): the trellis is
// invariant under that map, so both roles run the same code on physical even states, and a state handed to the partner arrives as
// the complement of the sender's index.  Complementing the 7-bit window flips the code bits of the odd-weight polynomials
// (133, 171) and leaves the even-weight one (145) alone: role 1 negates soft bits 0, 1 and 3.  The tie rules
// (include/b200dab_spec.h) are invariant under the same map, which makes the two roles agree with the scalar formulation bit for bit.
//
// registers: M[x] = metric of physical state 2x (x = 0..31) of both codewords.
//   forward  butterfly j (0..15): predecessors 2j = M[j] and 2j+32 = M[j+16]; survivor of new state 4j stays (M'[2j]), the one of
//            4j+1 belongs to the partner, where it is state 62-4j = M'[31-2j]: one SHFL.BFLY, no PRMT
//   backward butterfly j: successors 4j = M[2j] and 4j+1 = the partner's M[31-2j]; results are states 2j = M'[j], 2j+32 = M'[j+16]
// Branch metrics: beta_q = +-a0 +-a1 +-a2 +-a3 is one IDP.4A against a coefficient word that already holds the role's signs;
// beta_(7-k) = -beta_k, and the negative of a packed pair is the negative of the 32-bit word, so four registers serve all eight
// code-bit patterns (the sign goes into the add).
// Decision bits without predicates.  VIMNMX.S16x2 can hand out its two comparison results as predicates, but there are only seven
// predicate registers and the predicates arrive ~13 cycles after issue, so a lone warp spent more than half of its time waiting
// for them (ncu, stall "wait" on every first predicated consumer).  Instead: winner - loser-candidate + 0x7fff per half is
// >= 0x8000 exactly when the OTHER candidate won (both halves >= 0, no borrow between them), i.e. one IADD3 puts the two
// decisions into bits 15 and 31; PRMT with sign replication turns the four decision bits of a butterfly into four bytes of
// 0x00 / 0xff, and one LOP3 bit-select drops them into bit J & 7 of the four bytes of an accumulator.  Same instruction count as
// two predicated bit-sets per maximum, no predicate anywhere, and every instruction is free to move.
//   accumulator J >> 3: byte 0 = out0 decisions of codeword p, byte 1 = out1 of p, byte 2 = out0 of 16 + p, byte 3 = out1 of 16 + p
//   out0: bit = in1 taken.  out1: bit = in0 taken (complemented once per step with a constant).
template <int J>
__device__ __forceinline__ void vit_bfly(uint32_t in0, uint32_t in1, uint32_t& out0, uint32_t& out1, const uint32_t (&BP)[4], uint32_t (&acc)[2])
{
    constexpr int k = vit_kidx(2 * J);
    constexpr bool neg = k >= 4;
    constexpr int q = neg ? 7 - k : k;
    const uint32_t t0 = neg ? in0 - BP[q] : in0 + BP[q], t1 = neg ? in1 + BP[q] : in1 - BP[q];     // plain 32-bit adds: see the bias note
    out0 = __vmaxs2(t0, t1);                             // in0 survives a tie: the bit says t1 > t0
    const uint32_t z0 = out0 - t0 + 0x7fff7fffu;
    const uint32_t u0 = neg ? in0 + BP[q] : in0 - BP[q], u1 = neg ? in1 - BP[q] : in1 + BP[q];
    out1 = __vmaxs2(u0, u1);                             // in1 survives a tie: the bit says u0 > u1
    const uint32_t z1 = out1 - u1 + 0x7fff7fffu;
    uint32_t w;                                          // bytes (z0.1, z1.1, z0.3, z1.3), each replaced by eight copies of its MSB
    asm("prmt.b32 %0, %1, %2, 0xfbd9;" : "=r"(w) : "r"(z0), "r"(z1));
    constexpr uint32_t m = 0x01010101u << (J & 7);
    acc[J >> 3] = (acc[J >> 3] & ~m) | (w & m);          // one LOP3
}
template <int J> struct VitFwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[32], uint32_t (&N)[32], const uint32_t (&BP)[4], uint32_t (&acc)[2])
    {
        VitFwd<J - 1>::run(M, N, BP, acc);
        uint32_t e, o;
        vit_bfly<J>(M[J], M[J + 16], e, o, BP, acc);
        N[2 * J] = e;
        N[31 - 2 * J] = __shfl_xor_sync(0xffffffffu, o, 1);
    }
};
template <> struct VitFwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[32], uint32_t (&)[32], const uint32_t (&)[4], uint32_t (&)[2]) {}
};
template <int J> struct VitBwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[32], uint32_t (&N)[32], const uint32_t (&BP)[4], uint32_t (&acc)[2])
    {
        VitBwd<J - 1>::run(M, N, BP, acc);
        const uint32_t r = __shfl_xor_sync(0xffffffffu, M[31 - 2 * J], 1);
        vit_bfly<J>(M[2 * J], r, N[J], N[J + 16], BP, acc);
    }
};
template <> struct VitBwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[32], uint32_t (&)[32], const uint32_t (&)[4], uint32_t (&)[2]) {}
};

#define VIT_BIAS 16384u
#define VIT_PENV ((VIT_BIAS + VIT_START_PENALTY) * 65537u)
#define VIT_MEET_WORDS (2 * 64 * 32)        /* unsigned shorts per group: [direction][state][codeword] */

// one trellis step of a lane: branch metrics of both codewords from the ring, 16 butterflies M -> N, two decision words out
template <bool BWD>
__device__ __forceinline__ void vit_step(const uint32_t (&M)[32], uint32_t (&N)[32], uint32_t wA, uint32_t wB, const uint32_t (&C)[4],
                                         uint32_t* dst)
{
    uint32_t BP[4];
#pragma unroll
    for (int q = 0; q < 4; q++) BP[q] = (uint32_t)__dp4a((int)wA, (int)C[q], __dp4a((int)wB, (int)C[q], 0) << 16);
    uint32_t acc[2] = {0, 0};
    if (BWD) VitBwd<15>::run(M, N, BP, acc);
    else VitFwd<15>::run(M, N, BP, acc);
    // codeword p: bytes 0, 1 of both accumulators; codeword 16 + p: bytes 2, 3; the out1 bytes are complemented (1 = in1 was taken)
    dst[0] = __byte_perm(acc[0], acc[1], 0x5410) ^ 0xff00ff00u;
    dst[32] = __byte_perm(acc[0], acc[1], 0x7632) ^ 0xff00ff00u;
}

// Bidirectional decoding: the code is zero-terminated at BOTH ends, so one work item runs the trellis forward over steps
// [0, m) from state 0 while another runs it backward over steps [m, T) from state 0; at stage m the state maximising
// alpha + beta lies on the maximum-likelihood path (k_vit_trace picks it, traces back over [0, m) and forward over [m, T)).
// Same path as one-directional Viterbi whenever the best path is unique; the tie rules are mirrored by the CPU checker.
// Metrics carry a bias of 16384 so that both halves of a register stay non-negative and a packed add is one 32-bit add.
// grid = 2 * n_groups one-warp CTAs (longest groups first): item = (group, direction)
__global__ void __launch_bounds__(32) k_viterbi(const uint32_t* __restrict__ vin, uint32_t* __restrict__ dec, unsigned short* __restrict__ meet,
                                                const VitJob* __restrict__ jobs, const VitGroup* __restrict__ groups)
{
    constexpr int PF = VIT_NORM_PERIOD;
    static_assert(PF % 2 == 0, "the step loop is unrolled by two");
    __shared__ uint32_t ring[2][PF][32];
    const int item = blockIdx.x;
    const VitGroup grp = groups[item >> 1];
    const VitJob job = jobs[grp.job];
    const int lane = threadIdx.x, role = lane & 1, p = lane >> 1, dir = item & 1;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;                         // meeting stage, multiple of 32
    const uint32_t* in = vin + job.vin_off + ((size_t)grp.g_in_job * steps) * 32 + lane;
    uint32_t* d = dec + (job.dec_off + ((size_t)grp.g_in_job * steps) * 32 + p) * 2 + role;     // codeword 16 + p: 32 words further on
    unsigned short* mt = meet + ((size_t)(item >> 1) * 2 + dir) * (64 * 32) + p;
    const uint32_t flip = role ? 63u : 0u;
    // coefficient bytes (a0, a1, a2, a3) of beta_0..3 = T + X1, T + X2, T - X2, T - X1 with T = +-(a0 + a3), X1/X2 = +-a1 +/- a2
    const uint32_t C[4] = {role ? 0xff01ffffu : 0x01010101u, role ? 0xffffffffu : 0x01ff0101u,
                           role ? 0xff0101ffu : 0x0101ff01u, role ? 0xffff01ffu : 0x01ffff01u};

    uint32_t M[32], N[32];
#pragma unroll
    for (int j = 0; j < 32; j++) M[j] = VIT_PENV;
    if (role == 0) M[0] = VIT_BIAS * 65537u;                        // both codewords start (forward) / end (backward) in state 0
    auto prefetch = [&](int buf, int t_first, int t_step) {         // ring slot i holds step t_first + i * t_step
#pragma unroll
        for (int i = 0; i < PF; i++) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[buf][i][lane]);
            const uint32_t* src = in + (size_t)min(max(t_first + i * t_step, 0), steps - 1) * 32;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto renorm = [&]() {           // re-reference all metrics to state 0 (keeps the int16 range; constants do not change any decision)
        const uint32_t c = __shfl_sync(0xffffffffu, VIT_BIAS * 65537u - M[0], lane & ~1);
#pragma unroll
        for (int j = 0; j < 32; j++) M[j] += c;
    };
    const int t_begin = dir == 0 ? 0 : steps - 1, t_step = dir == 0 ? 1 : -1, n_total = dir == 0 ? m : steps - m;
    prefetch(0, t_begin, t_step);
    int buf = 0;
    for (int done = 0; done < n_total; done += PF, buf ^= 1) {
        const int t0 = t_begin + done * t_step;
        __syncwarp();                                               // everybody is through with the buffer that is refilled now
        if (done + PF < n_total) { prefetch(buf ^ 1, t0 + PF * t_step, t_step); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                               // a lane reads the words its neighbours fetched
        renorm();
        const int nst = min(PF, n_total - done);
        int i = 0;
        if (dir == 0) {
#pragma unroll 1
            for (; i + 1 < nst; i += 2) {           // two steps per iteration: M -> N -> M, no register copies
                vit_step<false>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 + i) * 64);
                vit_step<false>(N, M, ring[buf][i + 1][p], ring[buf][i + 1][16 + p], C, d + (size_t)(t0 + i + 1) * 64);
            }
            if (i < nst) {
                vit_step<false>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 + i) * 64);
#pragma unroll
                for (int j = 0; j < 32; j++) M[j] = N[j];
            }
        } else {
#pragma unroll 1
            for (; i + 1 < nst; i += 2) {
                vit_step<true>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 - i) * 64);
                vit_step<true>(N, M, ring[buf][i + 1][p], ring[buf][i + 1][16 + p], C, d + (size_t)(t0 - i - 1) * 64);
            }
            if (i < nst) {
                vit_step<true>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 - i) * 64);
#pragma unroll
                for (int j = 0; j < 32; j++) M[j] = N[j];
            }
        }
    }
    // metrics at the meeting stage: M[x] = physical state 2x
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t s = (uint32_t)(2 * j) ^ flip;
        mt[s * 32] = (unsigned short)(M[j] & 0xffffu);
        mt[s * 32 + 16] = (unsigned short)(M[j] >> 16);
    }
}

// Meeting state (maximum of alpha + beta, lowest state on ties), traceback over [0, m) (warp 0) and trace forward over
// [m, n_info) (warp 1), energy-dispersal XOR on the way out.  grid = n_groups, 64 threads, one codeword per lane.
// A role's decision word holds the decision of butterfly J (0..15) for its first / second output at bit 16 (J >> 3) + (J & 7) and
// 8 bits above.  Role 0 works on logical states: forward, new state n (n & 2 clear) is output n & 1 of butterfly n >> 2; backward,
// state s (s & 1 clear) is output s >> 5 of butterfly (s >> 1) & 15.  Role 1 holds the complemented states with complemented
// decisions, and the complement of (J, output) is (15 - J, 1 - output) = bit position 31 - pos: with w = x | (~brev(y)) << 32 of
// a decision pair (x: role 0, y: role 1) both roles are read with the same index formula.
// The state recursion is a dependent chain, so the 32 decision pairs of the next chunk are fetched while this one is traced.
__host__ __device__ __forceinline__ uint32_t vit_dec_index_fwd(uint32_t n) { return ((n >> 5) << 4) | ((n & 1u) << 3) | ((n >> 2) & 7u) | ((n & 2u) << 4); }
__host__ __device__ __forceinline__ uint32_t vit_dec_index_bwd(uint32_t s) { return (((s >> 4) & 1u) << 4) | ((s >> 5) << 3) | ((s >> 1) & 7u) | ((s & 1u) << 5); }
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
                const uint32_t idx = vit_dec_index_fwd(s);
                s = (s >> 1) | (((uint32_t)(w >> idx) & 1u) << 5);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const unsigned long long w = (unsigned long long)cur[i].x | ((unsigned long long)(~__brev(cur[i].y)) << 32);
                const uint32_t idx = vit_dec_index_bwd(s);
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
