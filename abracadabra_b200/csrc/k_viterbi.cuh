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
    int mode4;             // 1: four-lane survivor-word layout (set by the host together with the work items)
    int late;              // 1: traced by the second k_vit_trace launch (DAB+ sub-channels; the others go first so that their copy to the host starts early)
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
// K = code-bit pattern of the butterfly (vit_kidx of the state behind in0), BIT = bit of the four accumulator bytes it writes.
// Branch metrics come as P[q] = c + beta_q and Q[q] = c - beta_q with c = VIT_C = 512 > max |beta|: both halves of both words are
// non-negative, so (i) adding them is a plain 32-bit add with no borrow between the codewords and (ii) they are proper packed
// operands for VIADDMNMX.S16x2 = max(a + b, c) per half, which does the second candidate's add and the maximum in ONE instruction.
// Every candidate of a step carries the same extra c, which no decision sees; the renormalisation takes the drift out again.
template <int K, int BIT>
__device__ __forceinline__ void vit_bfly_k(uint32_t in0, uint32_t in1, uint32_t& out0, uint32_t& out1, const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t& acc)
{
    constexpr int k = K;
    constexpr bool neg = k >= 4;
    constexpr int q = neg ? 7 - k : k;
    const uint32_t up = neg ? Q[q] : P[q], dn = neg ? P[q] : Q[q];     // in0 + beta / in1 - beta for the even successor, swapped for the odd one
    const uint32_t t0 = in0 + up;                        // plain 32-bit add: see the bias note
    out0 = __viaddmax_s16x2(in1, dn, t0);                // max(in1 - beta, t0); in0 survives a tie: the bit says t1 > t0
    const uint32_t z0 = out0 - t0 + 0x7fff7fffu;
    const uint32_t u1 = in1 + up;
    out1 = __viaddmax_s16x2(in0, dn, u1);                // max(in0 - beta, u1); in1 survives a tie: the bit says u0 > u1
    const uint32_t z1 = out1 - u1 + 0x7fff7fffu;
    uint32_t w;                                          // bytes (z0.1, z1.1, z0.3, z1.3), each replaced by eight copies of its MSB
    asm("prmt.b32 %0, %1, %2, 0xfbd9;" : "=r"(w) : "r"(z0), "r"(z1));
    constexpr uint32_t m = 0x01010101u << BIT;
    acc = (acc & ~m) | (w & m);                          // one LOP3
}
#define VIT_C 512u
// P[q], Q[q] of both codewords of a lane from their soft words (wA: low halves, wB: high halves) and the lane's coefficient words
__device__ __forceinline__ void vit_branch(uint32_t wA, uint32_t wB, const uint32_t (&C)[4], uint32_t (&P)[4], uint32_t (&Q)[4])
{
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int hi = __dp4a((int)wB, (int)C[q], (int)VIT_C);
        P[q] = (uint32_t)__dp4a((int)wA, (int)C[q], hi * 65536 + (int)VIT_C);
        Q[q] = (2u * VIT_C) * 65537u - P[q];
    }
}
template <int J>
__device__ __forceinline__ void vit_bfly(uint32_t in0, uint32_t in1, uint32_t& out0, uint32_t& out1, const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t (&acc)[2])
{
    vit_bfly_k<vit_kidx(2 * J), (J & 7)>(in0, in1, out0, out1, P, Q, acc[J >> 3]);
}
template <int J> struct VitFwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[32], uint32_t (&N)[32], const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t (&acc)[2])
    {
        VitFwd<J - 1>::run(M, N, P, Q, acc);
        uint32_t e, o;
        vit_bfly<J>(M[J], M[J + 16], e, o, P, Q, acc);
        N[2 * J] = e;
        N[31 - 2 * J] = __shfl_xor_sync(0xffffffffu, o, 1);
    }
};
template <> struct VitFwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[32], uint32_t (&)[32], const uint32_t (&)[4], const uint32_t (&)[4], uint32_t (&)[2]) {}
};
template <int J> struct VitBwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[32], uint32_t (&N)[32], const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t (&acc)[2])
    {
        VitBwd<J - 1>::run(M, N, P, Q, acc);
        const uint32_t r = __shfl_xor_sync(0xffffffffu, M[31 - 2 * J], 1);
        vit_bfly<J>(M[2 * J], r, N[J], N[J + 16], P, Q, acc);
    }
};
template <> struct VitBwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[32], uint32_t (&)[32], const uint32_t (&)[4], const uint32_t (&)[4], uint32_t (&)[2]) {}
};

// ------------------------------------------------------------------------------------------------ four-lane split
// The same recursion with the 64 states of a codeword pair spread over FOUR lanes (16 registers each), 16 codewords per warp:
// half the dependent instruction chain per trellis step (8 butterflies instead of 16), which is what bounds the launch for the
// long sub-channels (a lone warp issues once per two cycles, so a chain of T steps costs T x instructions-per-step x 2 cycles
// however idle the machine is).  tools/vit4_model.py is the scalar model of this layout and checks it against the plain recursion.
//   lane a (0..3) of a quad owns the states with state & 3 == a; logical register r = state >> 2 lives in physical register
//   r ^ X[a], X = {11, 0, 6, 13}.
//   slot m (0..7): butterfly on physical registers m and m + 8.  Lanes 1, 2: in0 = state j, in1 = j + 32 and out0 / out1 are the
//   even / odd successor; lanes 0, 3 (bit 3 of X set): in0 = j + 32, in1 = j, out0 / out1 = odd / even successor -- the same
//   butterfly code with the same branch metric, and the tie rules fall on the right candidates in both cases.
//   forward : out0 of every lane goes through ONE shuffle (lane a reads lane R1[a] = {2, 0, 1, 3}) into register 2m ^ 6,
//             out1 through a second one (R2 = {0, 2, 3, 1}) into 2m ^ 13: no select on either side, register naming only.
//   backward: the reverse flow -- in0 = register 2m ^ 6 of lane R1inv[a] = {1, 2, 0, 3}, in1 = register 2m ^ 13 of lane
//             R2inv[a] = {0, 3, 1, 2}; results stay in registers m, m + 8.
//   branch metrics: slot m has the code-bit pattern of state 4m; what the lane's own state bits and X[a] add is a sign per soft
//   bit, folded into the IDP.4A coefficient words (lanes 0, 3: all four negated; lanes 1, 2: bits 1, 2 negated).
// Survivor words: one 16-bit half per lane, codeword and step (bit m: out0 of slot m, bit 8 + m: out1, 1 = in1 taken); two
// consecutive steps share a 32-bit word (even step in the low half):  dec32[((t / 2 * 2 + (a >> 1)) * 32 + codeword) * 2 + (a & 1)],
// the same 8 bytes per codeword and step as the two-lane layout.
#define VIT4_X0 11
#define VIT4_C1 6
#define VIT4_C2 13
template <int M_> struct Vit4Fwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[16], uint32_t (&N)[16], const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t& acc, int s1, int s2)
    {
        Vit4Fwd<M_ - 1>::run(M, N, P, Q, acc, s1, s2);
        uint32_t e, o;
        vit_bfly_k<vit_kidx(4 * M_), M_>(M[M_], M[M_ + 8], e, o, P, Q, acc);
        N[(2 * M_) ^ VIT4_C1] = __shfl_sync(0xffffffffu, e, s1);
        N[(2 * M_) ^ VIT4_C2] = __shfl_sync(0xffffffffu, o, s2);
    }
};
template <> struct Vit4Fwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[16], uint32_t (&)[16], const uint32_t (&)[4], const uint32_t (&)[4], uint32_t&, int, int) {}
};
template <int M_> struct Vit4Bwd {
    static __device__ __forceinline__ void run(const uint32_t (&M)[16], uint32_t (&N)[16], const uint32_t (&P)[4], const uint32_t (&Q)[4], uint32_t& acc, int s1, int s2)
    {
        Vit4Bwd<M_ - 1>::run(M, N, P, Q, acc, s1, s2);
        const uint32_t i0 = __shfl_sync(0xffffffffu, M[(2 * M_) ^ VIT4_C1], s1);
        const uint32_t i1 = __shfl_sync(0xffffffffu, M[(2 * M_) ^ VIT4_C2], s2);
        vit_bfly_k<vit_kidx(4 * M_), M_>(i0, i1, N[M_], N[M_ + 8], P, Q, acc);
    }
};
template <> struct Vit4Bwd<-1> {
    static __device__ __forceinline__ void run(const uint32_t (&)[16], uint32_t (&)[16], const uint32_t (&)[4], const uint32_t (&)[4], uint32_t&, int, int) {}
};
template <bool BWD>
__device__ __forceinline__ uint32_t vit4_step(const uint32_t (&M)[16], uint32_t (&N)[16], uint32_t wA, uint32_t wB, const uint32_t (&C)[4], int s1, int s2)
{
    uint32_t P[4], Q[4];
    vit_branch(wA, wB, C, P, Q);
    uint32_t acc = 0;
    if (BWD) Vit4Bwd<7>::run(M, N, P, Q, acc, s1, s2);
    else Vit4Fwd<7>::run(M, N, P, Q, acc, s1, s2);
    return acc;         // bytes: out0 of codeword p, out1 of p, out0 of 8 + p, out1 of 8 + p
}

#define VIT_BIAS 16384u
// Metrics are re-referenced to state 0 every VIT_RENORM steps.  Range: state 0 = 16384 after it, the others within +-6096 of it
// (6 steps x 1016), the penalised start states at 8192; a step adds 4..1020 (c +- beta), so after 8 steps every half lies in
// [8224, 30640]: never negative, never beyond int16 -- what the packed adds and the signed packed maximum need.
#define VIT_RENORM 8
#define VIT_PENV ((VIT_BIAS + VIT_START_PENALTY) * 65537u)
#define VIT_MEET_WORDS (2 * 64 * 32)        /* unsigned shorts per group: [direction][state][codeword] */

// one trellis step of a lane: branch metrics of both codewords from the ring, 16 butterflies M -> N, two decision words out
template <bool BWD>
__device__ __forceinline__ void vit_step(const uint32_t (&M)[32], uint32_t (&N)[32], uint32_t wA, uint32_t wB, const uint32_t (&C)[4],
                                         uint32_t* dst)
{
    uint32_t P[4], Q[4];
    vit_branch(wA, wB, C, P, Q);
    uint32_t acc[2] = {0, 0};
    if (BWD) VitBwd<15>::run(M, N, P, Q, acc);
    else VitFwd<15>::run(M, N, P, Q, acc);
    // codeword p: bytes 0, 1 of both accumulators; codeword 16 + p: bytes 2, 3; the out1 bytes are complemented (1 = in1 was taken)
    dst[0] = __byte_perm(acc[0], acc[1], 0x5410) ^ 0xff00ff00u;
    dst[32] = __byte_perm(acc[0], acc[1], 0x7632) ^ 0xff00ff00u;
}

// Work item of k_viterbi: one warp.  Two-lane mode: (group, direction) = 32 codewords; four-lane mode: (group, direction, half) =
// 16 codewords.  The host sorts the items by the length of their dependent instruction chain, longest first.
struct VitItem { int group; int what; };      // what: bit 0 direction, bit 1 half (four-lane), bit 2 four-lane mode

// Bidirectional decoding: the code is zero-terminated at BOTH ends, so one work item runs the trellis forward over steps
// [0, m) from state 0 while another runs it backward over steps [m, T) from state 0; at stage m the state maximising
// alpha + beta lies on the maximum-likelihood path (k_vit_trace picks it, traces back over [0, m) and forward over [m, T)).
// Same path as one-directional Viterbi whenever the best path is unique; the tie rules are mirrored by the CPU checker.
// Metrics carry a bias of 16384 so that both halves of a register stay non-negative and a packed add is one 32-bit add.
__device__ __forceinline__ void vit_run2(uint32_t (&ring)[2][VIT_NORM_PERIOD][32], const uint32_t* __restrict__ vin, uint32_t* __restrict__ dec,
                                         unsigned short* __restrict__ meet, const VitJob& job, int group, int g_in_job, int dir)
{
    constexpr int PF = VIT_NORM_PERIOD;
    static_assert(PF % 2 == 0, "the step loop is unrolled by two");
    const int lane = threadIdx.x & 31, role = lane & 1, p = lane >> 1;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;                         // meeting stage, multiple of 32
    const uint32_t* in = vin + job.vin_off + ((size_t)g_in_job * steps) * 32 + lane;
    uint32_t* d = dec + (job.dec_off + ((size_t)g_in_job * steps) * 32 + p) * 2 + role;     // codeword 16 + p: 32 words further on
    unsigned short* mt = meet + ((size_t)group * 2 + dir) * (64 * 32) + p;
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
        const int nst = min(PF, n_total - done);
        int i = 0;
        if (dir == 0) {
#pragma unroll 1
            for (; i + 1 < nst; i += 2) {           // two steps per iteration: M -> N -> M, no register copies
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
                vit_step<false>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 + i) * 64);
                vit_step<false>(N, M, ring[buf][i + 1][p], ring[buf][i + 1][16 + p], C, d + (size_t)(t0 + i + 1) * 64);
            }
            if (i < nst) {
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
                vit_step<false>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 + i) * 64);
#pragma unroll
                for (int j = 0; j < 32; j++) M[j] = N[j];
            }
        } else {
#pragma unroll 1
            for (; i + 1 < nst; i += 2) {
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
                vit_step<true>(M, N, ring[buf][i][p], ring[buf][i][16 + p], C, d + (size_t)(t0 - i) * 64);
                vit_step<true>(N, M, ring[buf][i + 1][p], ring[buf][i + 1][16 + p], C, d + (size_t)(t0 - i - 1) * 64);
            }
            if (i < nst) {
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
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

// four-lane item: codewords 16 half + p (low register halves) and 16 half + 8 + p (high halves) of the group, p = lane >> 2.
// The number of steps of either direction is even (steps and the meeting stage are), so steps always come in pairs.
__device__ __forceinline__ void vit_run4(uint32_t (&ring)[2][VIT_NORM_PERIOD][32], const uint32_t* __restrict__ vin, uint32_t* __restrict__ dec,
                                         unsigned short* __restrict__ meet, const VitJob& job, int group, int g_in_job, int dir, int half)
{
    constexpr int PF = VIT_NORM_PERIOD;
    const int lane = threadIdx.x & 31, a = lane & 3, p = lane >> 2, cw = 16 * half + p;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;
    const uint32_t* in = vin + job.vin_off + ((size_t)g_in_job * steps) * 32 + 16 * half + (lane & 15);
    // survivor word of step pair t2, lane a, codeword c: dec32[((2 t2 + (a >> 1)) * 32 + c) * 2 + (a & 1)]
    uint32_t* d = dec + (job.dec_off + ((size_t)g_in_job * steps) * 32) * 2 + ((size_t)(a >> 1) * 32 + cw) * 2 + (a & 1);
    unsigned short* mt = meet + ((size_t)group * 2 + dir) * (64 * 32) + cw;
    const bool outer = (a == 0) || (a == 3);
    const uint32_t C[4] = {outer ? 0xffffffffu : 0x01ffff01u, outer ? 0xff01ffffu : 0x0101ff01u,
                           outer ? 0xffff01ffu : 0x01ff0101u, outer ? 0xff0101ffu : 0x01010101u};
    const int quad = lane & ~3;
    // shuffle sources: forward R1 = {2, 0, 1, 3}, R2 = {0, 2, 3, 1}; backward their inverses {1, 2, 0, 3}, {0, 3, 1, 2}
    const int s1 = quad | (int)((dir == 0 ? 0x3102u : 0x3021u) >> (4 * a) & 3u);
    const int s2 = quad | (int)((dir == 0 ? 0x1320u : 0x2130u) >> (4 * a) & 3u);

    uint32_t M[16], N[16];
#pragma unroll
    for (int j = 0; j < 16; j++) M[j] = VIT_PENV;
    if (a == 0) M[VIT4_X0] = VIT_BIAS * 65537u;                     // state 0: lane 0, logical register 0
    auto prefetch = [&](int buf, int t_first, int t_step) {         // lanes 0..15 fetch the even ring slots, lanes 16..31 the odd ones
#pragma unroll
        for (int i = 0; i < PF / 2; i++) {
            const int slot = 2 * i + (lane >> 4);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[buf][slot][16 * half + (lane & 15)]);
            const uint32_t* src = in + (size_t)min(max(t_first + slot * t_step, 0), steps - 1) * 32;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto renorm = [&]() {
        const uint32_t c = __shfl_sync(0xffffffffu, VIT_BIAS * 65537u - M[VIT4_X0], quad);
#pragma unroll
        for (int j = 0; j < 16; j++) M[j] += c;
    };
    const int t_begin = dir == 0 ? 0 : steps - 1, t_step = dir == 0 ? 1 : -1, n_total = dir == 0 ? m : steps - m;
    prefetch(0, t_begin, t_step);
    int buf = 0;
    for (int done = 0; done < n_total; done += PF, buf ^= 1) {
        const int t0 = t_begin + done * t_step;
        __syncwarp();
        if (done + PF < n_total) { prefetch(buf ^ 1, t0 + PF * t_step, t_step); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int nst = min(PF, n_total - done);
        if (dir == 0) {
#pragma unroll 1
            for (int i = 0; i < nst; i += 2) {
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
                const uint32_t e = vit4_step<false>(M, N, ring[buf][i][cw], ring[buf][i][cw + 8], C, s1, s2);
                const uint32_t o = vit4_step<false>(N, M, ring[buf][i + 1][cw], ring[buf][i + 1][cw + 8], C, s1, s2);
                uint32_t* dst = d + (size_t)(t0 + i) * 64;           // pair (t0 + i) / 2: 128 words per pair
                dst[0] = __byte_perm(e, o, 0x5410) ^ 0xff00ff00u;
                dst[16] = __byte_perm(e, o, 0x7632) ^ 0xff00ff00u;   // codeword + 8
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < nst; i += 2) {
                if ((i & (VIT_RENORM - 1)) == 0) renorm();
                const uint32_t o = vit4_step<true>(M, N, ring[buf][i][cw], ring[buf][i][cw + 8], C, s1, s2);          // odd step t0 - i
                const uint32_t e = vit4_step<true>(N, M, ring[buf][i + 1][cw], ring[buf][i + 1][cw + 8], C, s1, s2);  // even step t0 - i - 1
                uint32_t* dst = d + (size_t)(t0 - i - 1) * 64;
                dst[0] = __byte_perm(e, o, 0x5410) ^ 0xff00ff00u;
                dst[16] = __byte_perm(e, o, 0x7632) ^ 0xff00ff00u;
            }
        }
    }
    // metrics at the meeting stage: physical register j of lane a = logical register j ^ X[a] = state 4 (j ^ X[a]) + a
    const uint32_t xa = (0xd60bu >> (4 * a)) & 15u;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t s = 4u * ((uint32_t)j ^ xa) + (uint32_t)a;
        mt[s * 32] = (unsigned short)(M[j] & 0xffffu);
        mt[s * 32 + 8] = (unsigned short)(M[j] >> 16);
    }
}

// Persistent kernel: CTAs of four warps (one per SM sub-partition), a few CTAs per SM; every warp takes the next work item from
// a global queue until it is empty.  The items are sorted by the length of their dependent instruction chain, longest first, so
// the warps of every sub-partition stay busy until the short items at the end of the list run out (list scheduling); with
// one-warp CTAs dealt by the hardware the sub-partitions of an SM ended up with 0.64-1.09 M instructions each (profiles/r01k).
// `counter` is zero on entry; k_vit_trace, which always follows on the stream, clears it again.
#define VIT_CTA_WARPS 4
__global__ void __launch_bounds__(32 * VIT_CTA_WARPS) k_viterbi(const uint32_t* __restrict__ vin, uint32_t* __restrict__ dec, unsigned short* __restrict__ meet,
                                                                  const VitJob* __restrict__ jobs, const VitGroup* __restrict__ groups,
                                                                  const VitItem* __restrict__ items, int n_items, int* __restrict__ counter, int* __restrict__ done)
{
    __shared__ uint32_t ring[VIT_CTA_WARPS][2][VIT_NORM_PERIOD][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Programmatic dependent launch: the trace kernel queued behind this one may start as soon as every CTA of this grid is running
    // (its CTAs then wait for the groups they trace, see k_vit_trace); without the launch attribute this is a no-op.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n_items) break;
        const VitItem item = items[idx];
        const VitGroup grp = groups[item.group];
        const VitJob job = jobs[grp.job];
        if (item.what & 4) vit_run4(ring[warp], vin, dec, meet, job, item.group, grp.g_in_job, item.what & 1, (item.what >> 1) & 1);
        else vit_run2(ring[warp], vin, dec, meet, job, item.group, grp.g_in_job, item.what & 1);
        if (done) {
            __threadfence();                            // every lane's decision words and metrics are out before the counter moves
            __syncwarp();
            if (lane == 0) atomicAdd(done + item.group, 1);
        }
        __syncwarp();
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
// Four-lane layout: w = the four 16-bit halves of a step (lane a at bit 16 a).  The function gives bit index | invert << 6:
// forward (state n after the step) -> MSB of the predecessor = bit ^ invert; backward (state s before the step) -> input bit.
// Closed forms of the tables tools/vit4_model.py derives from the register maps: base[n & 7] ^ (n >> 3) and base[s & 3] ^ (s >> 2).
__host__ __device__ __forceinline__ uint32_t vit4_dec_index_fwd(uint32_t n)
{
    const unsigned long long base = 0x757d2e261810434bull;                           // bytes 75, 67, 16, 24, 38, 46, 125, 117
    return ((uint32_t)(base >> (8 * (n & 7u))) & 0x7fu) ^ (n >> 3);
}
__host__ __device__ __forceinline__ uint32_t vit4_dec_index_bwd(uint32_t s)
{
    const uint32_t base = 0x7d26104bu;                                               // bytes 75, 16, 38, 125
    return ((base >> (8 * (s & 3u))) & 0x7fu) ^ (s >> 2);
}
__device__ __forceinline__ void vit_load_chunk(uint2 (&b)[32], const uint2* d, int c)
{
#pragma unroll
    for (int i = 0; i < 32; i++) b[i] = __ldcg(d + (size_t)(32 * c + i) * 32);        // L2: written by other SMs, possibly while this kernel is already running
}
// the 64 survivor bits of step i (0..31) of a chunk, two-lane layout: role 0's word | ~bit-reverse(role 1's word) << 32
__device__ __forceinline__ unsigned long long vit_word2(const uint2 (&b)[32], int i)
{
    return (unsigned long long)b[i].x | ((unsigned long long)(~__brev(b[i].y)) << 32);
}
// four-lane layout: b[2 k] = (lane 0, lane 1), b[2 k + 1] = (lane 2, lane 3) of step pair k; even step in the low halves
__device__ __forceinline__ unsigned long long vit_word4(const uint2 (&b)[32], int i)
{
    const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
    const uint2 lo = b[(i >> 1) * 2], hi = b[(i >> 1) * 2 + 1];
    return (unsigned long long)__byte_perm(lo.x, lo.y, sel) | ((unsigned long long)__byte_perm(hi.x, hi.y, sel) << 32);
}
template <bool FOUR>
__device__ __forceinline__ void vit_trace_run(const uint2* __restrict__ d, const VitJob& job, int dir, int m, uint32_t s, bool live,
                                              uint32_t* __restrict__ orow, const uint32_t* __restrict__ prbs_w)
{
    const int c_first = dir == 0 ? m / 32 - 1 : m / 32, c_end = dir == 0 ? -1 : job.n_info / 32, c_step = dir == 0 ? -1 : 1;
    uint2 cur[32], nxt[32];
    if (c_first != c_end) vit_load_chunk(cur, d, c_first);
    for (int c = c_first; c != c_end; c += c_step) {
        const bool more = (c + c_step) != c_end;
        if (more) vit_load_chunk(nxt, d, c + c_step);
        uint32_t word = 0;
        if (dir == 0) {
            // traceback: s is the state after step 32 c + i
#pragma unroll
            for (int i = 31; i >= 0; i--) {
                const unsigned long long w = FOUR ? vit_word4(cur, i) : vit_word2(cur, i);
                word |= (s & 1u) << (31 - i);
                const uint32_t idx = FOUR ? vit4_dec_index_fwd(s) : vit_dec_index_fwd(s);
                s = (s >> 1) | ((((uint32_t)(w >> (idx & 63u)) ^ (idx >> 6)) & 1u) << 5);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const unsigned long long w = FOUR ? vit_word4(cur, i) : vit_word2(cur, i);
                const uint32_t idx = FOUR ? vit4_dec_index_bwd(s) : vit_dec_index_bwd(s);
                const uint32_t bit = ((uint32_t)(w >> (idx & 63u)) ^ (idx >> 6)) & 1u;
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
__global__ void __launch_bounds__(64) k_vit_trace(const uint2* dec, const unsigned short* meet,
                                                  const VitJob* __restrict__ jobs, const VitGroup* __restrict__ groups,
                                                  const uint32_t* __restrict__ prbs_w, uint8_t* __restrict__ out, int* __restrict__ vit_counter, int phase,
                                                  const int* __restrict__ order, const int* done, uint8_t* __restrict__ host_out, int host_vec)
{
    if (vit_counter && blockIdx.x == 0 && threadIdx.x == 0) *vit_counter = 0;   // the work queue of the k_viterbi launch in front of this one (not when the two run side by side)
    // order: CTA -> group, in the order in which the groups' recursions are expected to finish (host simulation of the k_viterbi
    // schedule).  done: launched as a programmatic dependent of k_viterbi, this kernel runs NEXT TO it, on the registers and warp
    // slots its persistent CTAs leave free: a CTA waits until the recursions of its group (two, or four in four-lane mode) have
    // published their decisions, traces, and makes room for the next.  Every k_viterbi CTA is resident before the first CTA of this
    // kernel starts, and the recursions wait for nothing: no deadlock.  What is left behind the end of k_viterbi is the trace of the
    // groups that finished last instead of the whole 0.14 ms.
    const int g = order ? order[blockIdx.x] : (int)blockIdx.x;
    const VitGroup grp = groups[g];
    const VitJob job = jobs[grp.job];
    if (phase >= 0 && job.late != phase) return;                    // two launches: phase 0 = jobs whose output is final after the trace
    if (done) {
        if (threadIdx.x == 0) {
            const int need = job.mode4 ? 4 : 2;
            while (*reinterpret_cast<const volatile int*>(done + g) < need) __nanosleep(1000);
            __threadfence();
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, dir = threadIdx.x >> 5;
    const int steps = job.steps;
    const int m = ((steps >> 1) >> 5) << 5;
    const uint2* d = dec + job.dec_off + ((size_t)grp.g_in_job * steps) * 32 + lane;
    const unsigned short* mt = meet + (size_t)g * VIT_MEET_WORDS + lane;
    uint32_t s = 0;
    int best = -1;
#pragma unroll 8
    for (int q = 0; q < 64; q++) {
        const int v = (int)__ldcg(mt + q * 32) + (int)__ldcg(mt + (64 + q) * 32);
        if (v > best) { best = v; s = (uint32_t)q; }
    }
    const int row = grp.g_in_job * 32 + lane;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + job.out_off + (size_t)row * job.out_pitch);
    const bool live = row < job.n_rows;
    if (job.mode4) vit_trace_run<true>(d, job, dir, m, s, live, orow, prbs_w);
    else vit_trace_run<false>(d, job, dir, m, s, live, orow, prbs_w);
    // host_out: the logical frames go straight to the caller's (pinned, device-visible) result buffer, group by group, while the
    // recursions of the other groups are still running -- instead of one device-to-host copy of everything behind the last trace, on
    // the critical path of a step.  The group's rows are contiguous (32 x out_pitch bytes at the same offset in both buffers); the CTA
    // re-reads what its lanes wrote row by row (L2) and stores it in long consecutive pieces, which cross PCIe as full packets.
    if (host_out) {
        __syncthreads();                                            // both warps' rows are written
        const int r0 = grp.g_in_job * 32;
        const int rows_here = min(32, job.n_rows - r0);
        if (rows_here > 0) {
            const size_t boff = (size_t)job.out_off + (size_t)r0 * job.out_pitch;      // multiple of 768: out_pitch is a multiple of 24
            const int nbytes = rows_here * job.out_pitch;
            int done_bytes = 0;
            if (host_vec == 16) {                                   // a warp's store = 512 consecutive bytes
                const int n16 = nbytes >> 4;
                const uint4* src = reinterpret_cast<const uint4*>(out + boff);
                uint4* dst = reinterpret_cast<uint4*>(host_out + boff);
                for (int i = threadIdx.x; i < n16; i += 64) dst[i] = __ldcg(src + i);
                done_bytes = n16 << 4;
            } else {
                const int n8 = nbytes >> 3;
                const uint2* src = reinterpret_cast<const uint2*>(out + boff);
                uint2* dst = reinterpret_cast<uint2*>(host_out + boff);
                for (int i = threadIdx.x; i < n8; i += 64) dst[i] = __ldcg(src + i);
                done_bytes = n8 << 3;
            }
            for (int i = done_bytes + threadIdx.x; i < nbytes; i += 64) host_out[boff + i] = __ldcg(out + boff + i);
        }
    }
}
