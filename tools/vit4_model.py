#!/usr/bin/env python
"""Scalar model of the FOUR-LANE Viterbi lane split (k_viterbi4 in csrc/k_viterbi.cuh): checks the register maps, the shuffle
sources, the per-lane sign flips and the decision look-up tables against the plain formulation of include/b200dab_spec.h /
oracle/daboracle.c:orc_viterbi.  Prints the tables the kernel uses.  Pure host arithmetic, no GPU.
"""
import sys
import numpy as np

POLY = [0o155, 0o117, 0o123]          # bit-reversed 133, 171, 145 (bit d of (state << 1 | input) is a[i-d]); soft bit 3 repeats 0
X = [11, 0, 6, 13]                    # physical register of logical register r in lane a: r ^ X[a]
C1, C2 = 6, 13                        # forward: out0 of slot m lands in register 2m ^ C1, out1 in 2m ^ C2
R1SRC = [2, 0, 1, 3]                  # forward round 1 (out0): lane a reads from lane R1SRC[a]
R2SRC = [0, 2, 3, 1]                  # forward round 2 (out1)
R1INV = [R1SRC.index(a) for a in range(4)]   # backward: lane a fetches in0 from lane R1INV[a] (register 2m ^ C1)
R2INV = [R2SRC.index(a) for a in range(4)]


def parity(x):
    return bin(x).count("1") & 1


def cbit(j, q):
    return parity((j << 1) & POLY[q])


def kidx(j):
    return cbit(j, 0) * 4 + cbit(j, 1) * 2 + cbit(j, 2)


def lane_flips(a):
    """sign flips of soft bits 0..3 in lane a relative to the compile-time pattern of slot m (state 4m)"""
    f = [cbit(4 * (X[a] & 7), q) ^ cbit(a, q) for q in range(3)]
    return f + [f[0]]


def coef(q, i, a):
    """coefficient of soft bit i in beta_q (q = 0..3 = code-bit pattern c0 c1 c2 = 0 q1 q0) for lane a"""
    c = [0, (q >> 1) & 1, q & 1, 0][i]
    return (-1 if c else 1) * (-1 if lane_flips(a)[i] else 1)


def ref_bm(j, s):
    return sum(int(s[i]) * (1 - 2 * cbit(j, min(i, 3) if i < 3 else 0)) for i in range(4))


def ref_forward(M, s):
    N = np.zeros(64, np.int64); d = 0
    for j in range(32):
        b = ref_bm(j, s)
        c0, c1 = M[j] + b, M[j + 32] - b
        if c1 > c0: N[2 * j] = c1; d |= 1 << (2 * j)
        else: N[2 * j] = c0
        c0, c1 = M[j] - b, M[j + 32] + b
        if c1 >= c0: N[2 * j + 1] = c1; d |= 1 << (2 * j + 1)
        else: N[2 * j + 1] = c0
    return N, d


def ref_backward(B, s):
    N = np.zeros(64, np.int64); d = 0
    for j in range(32):
        b = ref_bm(j, s)
        c0, c1 = B[2 * j] + b, B[2 * j + 1] - b
        if c1 > c0: N[j] = c1; d |= 1 << j
        else: N[j] = c0
        c0, c1 = B[2 * j] - b, B[2 * j + 1] + b
        if c1 >= c0: N[j + 32] = c1; d |= 1 << (j + 32)
        else: N[j + 32] = c0
    return N, d


def to_lanes(M):
    L = np.zeros((4, 16), np.int64)
    for s in range(64):
        L[s & 3, (s >> 2) ^ X[s & 3]] = M[s]
    return L


def from_lanes(L):
    M = np.zeros(64, np.int64)
    for s in range(64):
        M[s] = L[s & 3, (s >> 2) ^ X[s & 3]]
    return M


def bfly(m, in0, in1, BP):
    k = kidx(4 * m); neg = k >= 4; q = 7 - k if neg else k
    bm = -BP[q] if neg else BP[q]
    t0, t1 = in0 + bm, in1 - bm
    o0 = max(t0, t1); d0 = int(t1 > t0)
    u0, u1 = in0 - bm, in1 + bm
    o1 = max(u0, u1); d1 = int(not (u0 > u1))
    return o0, o1, d0, d1


def lane_bp(a, s):
    return [sum(int(s[i]) * coef(q, i, a) for i in range(4)) for q in range(4)]


def step_fwd(L, s):
    out0 = np.zeros((4, 8), np.int64); out1 = np.zeros((4, 8), np.int64); half = [0] * 4
    for a in range(4):
        BP = lane_bp(a, s)
        for m in range(8):
            o0, o1, d0, d1 = bfly(m, L[a, m], L[a, m + 8], BP)
            out0[a, m], out1[a, m] = o0, o1
            half[a] |= (d0 << m) | (d1 << (8 + m))
    N = np.zeros((4, 16), np.int64)
    for a in range(4):
        for m in range(8):
            N[a, (2 * m) ^ C1] = out0[R1SRC[a], m]
            N[a, (2 * m) ^ C2] = out1[R2SRC[a], m]
    return N, half


def step_bwd(L, s):
    N = np.zeros((4, 16), np.int64); half = [0] * 4
    for a in range(4):
        BP = lane_bp(a, s)
        for m in range(8):
            i0 = L[R1INV[a], (2 * m) ^ C1]
            i1 = L[R2INV[a], (2 * m) ^ C2]
            o0, o1, d0, d1 = bfly(m, i0, i1, BP)
            N[a, m], N[a, m + 8] = o0, o1
            half[a] |= (d0 << m) | (d1 << (8 + m))
    return N, half


def build_luts():
    """byte per state: bit index into the 64-bit word (lane a's 16 bits at 16 a) | invert << 6"""
    fwd = [0] * 64; bwd = [0] * 64
    for n in range(64):                                  # forward: state n AFTER the step; decision = MSB of the predecessor
        t, p = n & 3, (n >> 2) ^ X[n & 3]
        if ((p ^ C1) & 1) == 0: m, rnd, src = (p ^ C1) >> 1, 0, R1SRC[t]
        else: m, rnd, src = (p ^ C2) >> 1, 1, R2SRC[t]
        fwd[n] = (16 * src + 8 * rnd + m) | ((X[src] >> 3) << 6)
    for s in range(64):                                  # backward: state s BEFORE the step; decision = input bit
        a, p = s & 3, (s >> 2) ^ X[s & 3]
        bwd[s] = (16 * a + 8 * (p >> 3) + (p & 7)) | ((X[a] >> 3) << 6)
    return fwd, bwd


def main():
    rng = np.random.default_rng(5)
    fwd, bwd = build_luts()
    for trial in range(40):
        M = rng.integers(-3000, 3000, 64)
        if trial % 3 == 0: M[rng.integers(0, 64, 20)] = 100          # provoke ties
        s = rng.integers(-127, 128, 4) if trial % 4 else np.array([0, 0, 0, 0])
        Nr, dr = ref_forward(M, s)
        Nl, half = step_fwd(to_lanes(M), s)
        assert np.array_equal(from_lanes(Nl), Nr), "forward metrics"
        w = sum(h << (16 * a) for a, h in enumerate(half))
        for n in range(64):
            bit = ((w >> (fwd[n] & 63)) & 1) ^ (fwd[n] >> 6)
            assert bit == (dr >> n) & 1, ("forward decision", n)
        Br, db = ref_backward(M, s)
        Bl, half = step_bwd(to_lanes(M), s)
        assert np.array_equal(from_lanes(Bl), Br), "backward metrics"
        w = sum(h << (16 * a) for a, h in enumerate(half))
        for st in range(64):
            bit = ((w >> (bwd[st] & 63)) & 1) ^ (bwd[st] >> 6)
            assert bit == (db >> st) & 1, ("backward decision", st)
    print("four-lane model agrees with the plain recursion (metrics and decisions, ties included)")
    print("X =", X, " C1, C2 =", C1, C2, " R1SRC =", R1SRC, " R2SRC =", R2SRC, " R1INV =", R1INV, " R2INV =", R2INV)
    for a in range(4):
        print(f"lane {a}: flips {lane_flips(a)}  C words", [hex(sum(((coef(q, i, a) & 0xff) << (8 * i)) for i in range(4))) for q in range(4)])
    print("slot patterns k(4m):", [kidx(4 * m) for m in range(8)])
    print("fwd lut:", ",".join(str(v) for v in fwd))
    print("bwd lut:", ",".join(str(v) for v in bwd))


if __name__ == "__main__":
    sys.exit(main())
