// k_dabplus.cuh -- FIB CRC, DAB+ superframe synchronisation (fire code), RS(120,110) correction, AU split + CRC.
//
// Replaces the reference core's FIB CRC check, firecode search, libfec decode_rs_char calls and AU CRC check
// (observable through dabsdrNtfPeriodic_t fibErrorCntr/rs*/mscCrc* (dabsdr.h:151-166) and the
// access units handed to the audio callback (dabsdr.h:71-78)).
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint16_t crc16_ccitt_dev(const uint8_t* __restrict__ p, int n, uint16_t init, uint16_t poly)
{
    uint32_t crc = init;
    for (int i = 0; i < n; i++) {
        crc ^= (uint32_t)p[i] << 8;
#pragma unroll
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000u) ? ((crc << 1) ^ poly) & 0xffffu : (crc << 1) & 0xffffu;
    }
    return (uint16_t)crc;
}

// one thread per FIB (32 bytes): ok[i] = CRC-16 (0x1021, init 0xFFFF, inverted) matches
__global__ void __launch_bounds__(256) k_fib_crc(const uint8_t* __restrict__ fibs, int n_fib, uint8_t* __restrict__ ok)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_fib) return;
    const uint8_t* f = fibs + (size_t)i * 32;
    uint16_t crc = (uint16_t)~crc16_ccitt_dev(f, 30, 0xFFFF, 0x1021);
    ok[i] = crc == (uint16_t)((f[30] << 8) | f[31]);
}

struct DabPlusJob {
    long long frames_off;   // byte offset of row 0 of this sub-channel's logical frames
    int n_rows;             // logical frames available
    int pitch;              // bytes per logical frame = 24 * s
    int s;                  // RS codewords per superframe = bitrate / 8
    int sf_off;             // index of this job's first superframe slot in the per-superframe arrays
    int max_sf;             // slots reserved = n_rows / 5
    int flag_off;           // index into the per-row firecode flag array
    int subch_index;        // index of the sub-channel in the caller's table (copied into the AU records)
    int forced_phase;       // >= 0: superframes start at rows = forced_phase (mod 5), no vote (chunked decoding: the vote is global)
    int pad_;
};
struct SfInfo {             // per superframe slot
    int valid;              // 1 = slot used (inside the aligned run)
    int row;                // first logical frame
    int rs_corrected, rs_bit_errors, rs_uncorrectable;
    int firecode_ok;
    int n_aus;
    int au_off[6], au_len[6], au_ok[6];
    int header;             // superframe byte 2
};

// Result records in the layout of the C ABI (b200dab_au_t: seven int32) and the totals the host used to add up from SfInfo.
struct AuRecord { int subch_index, superframe, cif, offset, length, crc_ok, header; };
struct DabPlusTotals { int n_aus, superframes, rs_corrected, rs_bit_errors, rs_uncorrectable, firecode_fail, n_patches, pad_; };
// One corrected byte: where it lies inside the logical-frame buffer and what it is now.  The host receives the logical frames of the DAB+
// sub-channels as they were traced -- before the RS correction (k_vit_trace writes them to host memory) -- and applies this list to
// its copy (b200dab_decode).
struct RsPatch { uint32_t offset, value; };

// ---- GF(2^8) / 0x11D
struct GfTables { uint8_t exp[512]; uint8_t log[256]; };
__device__ __forceinline__ uint8_t gf_mul(const GfTables& g, uint8_t a, uint8_t b) { return (a && b) ? g.exp[g.log[a] + g.log[b]] : 0; }
__device__ __forceinline__ uint8_t gf_inv(const GfTables& g, uint8_t a) { return g.exp[255 - g.log[a]]; }

// Constant tables of the superframe kernel, built once on the host (b200dab_ctx::init_tables) and copied to shared memory per CTA:
//   fb[t]  = t * g(x), g(x) = prod_{i < 10} (x - alpha^i): what coefficient t leaving at the top adds to the remainder (3 words = 10 bytes)
//   crc[k][v] = CRC-16 (0x1021) state after byte v followed by k zero bytes (slicing by four)
struct SfTables {
    uint32_t fb[256][3];
    uint16_t crc[4][256];
    GfTables gf;
};

// fire code of one logical frame: CRC-16 (0x782F, init 0) over bytes 2..10 against bytes 0..1; an all-zero header never counts
__device__ __forceinline__ int firecode_ok(const uint8_t* f)
{
    const uint16_t fc = crc16_ccitt_dev(f + 2, 9, 0, 0x782F);
    return (fc == (uint16_t)((f[0] << 8) | f[1])) && ((f[0] | f[1] | f[2]) != 0);
}

// One CTA per DAB+ sub-channel: fire code over bytes 2..10 of every logical frame (pre-RS, like the reference, which needs it to
// find the alignment), the residue (row mod 5) with most hits, and the superframe slots that follow from it.  (Was two launches:
// a scan with a thread per row and a vote with a CTA per sub-channel; the tail of a step is latency-bound, every launch counts.)
#define FIRE_THREADS 1024
__global__ void __launch_bounds__(FIRE_THREADS) k_fire_align(const uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs,
                                                             uint8_t* __restrict__ flags, SfInfo* __restrict__ sf, int* __restrict__ residue_out,
                                                             DabPlusTotals* __restrict__ totals)
{
    __shared__ int part[FIRE_THREADS / 32][5];
    if (blockIdx.x == 0 && threadIdx.x < (int)(sizeof(DabPlusTotals) / 4)) reinterpret_cast<int*>(totals)[threadIdx.x] = 0;   // k_sf_decode / k_au_list add to it
    __shared__ int s_best;
    const DabPlusJob job = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    int cnt[5] = {0, 0, 0, 0, 0};
    for (int r = tid; r < job.n_rows; r += FIRE_THREADS) {
        // the first 12 bytes of the row in two loads (rows are 8-byte aligned: the pitch is a multiple of 24)
        const uint8_t* f = frames + job.frames_off + (size_t)r * job.pitch;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(f));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(f + 8));
        uint8_t h[12];
#pragma unroll
        for (int i = 0; i < 4; i++) { h[i] = (uint8_t)(a.x >> (8 * i)); h[4 + i] = (uint8_t)(a.y >> (8 * i)); h[8 + i] = (uint8_t)(b >> (8 * i)); }
        const int ok = firecode_ok(h);
        flags[job.flag_off + r] = (uint8_t)ok;
        const int k = r % 5;
#pragma unroll
        for (int q = 0; q < 5; q++) cnt[q] += (k == q) ? ok : 0;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
        if ((tid & 31) == 0) part[tid >> 5][k] = cnt[k];
    }
    __syncthreads();
    if (tid == 0) {
        int best = -1, bc = 0;
        for (int k = 0; k < 5; k++) {
            int c = 0;
            for (int w = 0; w < FIRE_THREADS / 32; w++) c += part[w][k];
            if (c > bc) { bc = c; best = k; }           // first maximum
        }
        if (job.forced_phase >= 0) best = job.forced_phase % 5;
        s_best = best;
        residue_out[blockIdx.x] = best;
    }
    __syncthreads();
    const int best = s_best;
    for (int m = tid; m < job.max_sf; m += FIRE_THREADS) {
        SfInfo* o = sf + job.sf_off + m;
        const int row = best + 5 * m;
        o->valid = (best >= 0 && row + 5 <= job.n_rows) ? 1 : 0;
        o->row = row;
        o->rs_corrected = o->rs_bit_errors = o->rs_uncorrectable = 0;
        o->firecode_ok = 0; o->n_aus = 0; o->header = 0;
    }
}

// RS(120,110) correction of codeword j (bytes j, j + s, j + 2 s, ... of the superframe in shared memory); corrected bytes go to
// `buf` and to `glob`.  Returns the number of corrected bytes (bit errors in *bit_errors), or -1 if the word is uncorrectable.
__device__ __noinline__ int rs_correct(uint8_t* buf, uint8_t* glob, int j, int s, uint32_t w0, uint32_t w1, uint32_t w2, const GfTables& gf, int* bit_errors,
                                       RsPatch* patches, int patch_cap, int* n_patches, uint32_t glob_off)
{
    const int n = 120, nroots = 10;
    uint8_t S[10];
    {
        // the syndromes are the remainder evaluated at alpha^i, since g(alpha^i) = 0
        uint8_t r[10];
#pragma unroll
        for (int q = 0; q < 10; q++) r[q] = (uint8_t)((q < 4 ? w0 : q < 8 ? w1 : w2) >> (8 * (q & 3)));
        for (int i = 0; i < nroots; i++) {
            uint8_t v = 0;
            for (int q = 9; q >= 0; q--) {
                v = v ? gf.exp[gf.log[v] + i] : 0;            // * alpha^i
                v ^= r[q];
            }
            S[i] = v;
        }
    }
    // Berlekamp-Massey
    uint8_t Lm[12], B[12], T[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { Lm[i] = 0; B[i] = 0; }
    Lm[0] = 1; B[0] = 1;
    int el = 0, mm = 1; uint8_t bb = 1;
    for (int r = 0; r < nroots; r++) {
        uint8_t d = S[r];
        for (int i = 1; i <= el; i++) d ^= gf_mul(gf, Lm[i], S[r - i]);
        if (d == 0) { mm++; continue; }
        for (int i = 0; i < 12; i++) T[i] = Lm[i];
        const uint8_t coef = gf_mul(gf, d, gf_inv(gf, bb));
        for (int i = 0; i + mm < 12; i++) Lm[i + mm] ^= gf_mul(gf, coef, B[i]);
        if (2 * el <= r) { el = r + 1 - el; for (int i = 0; i < 12; i++) B[i] = T[i]; bb = d; mm = 1; } else mm++;
    }
    if (el > nroots / 2) return -1;
    // Chien search over the 120 positions
    int pos[5], np = 0;
    for (int k = 0; k < n; k++) {
        const int p = n - 1 - k;
        const uint8_t xinv = gf.exp[(255 - p) % 255];
        uint8_t v = 0, xp = 1;
        for (int i = 0; i <= el; i++) { v ^= gf_mul(gf, Lm[i], xp); xp = gf_mul(gf, xp, xinv); }
        if (v == 0) { if (np < 5) pos[np] = k; np++; }
    }
    if (np != el) return -1;
    // Forney (first root alpha^0): e = X * Omega(X^-1) / Lambda'(X^-1)
    uint8_t Om[10];
    for (int i = 0; i < nroots; i++) {
        uint8_t v = 0;
        for (int k = 0; k <= i && k <= el; k++) v ^= gf_mul(gf, Lm[k], S[i - k]);
        Om[i] = v;
    }
    uint8_t ev[5];
    for (int e = 0; e < np; e++) {
        const int p = n - 1 - pos[e];
        const uint8_t X = gf.exp[p % 255], xinv = gf.exp[(255 - p) % 255];
        uint8_t num = 0, xp = 1;
        for (int i = 0; i < nroots; i++) { num ^= gf_mul(gf, Om[i], xp); xp = gf_mul(gf, xp, xinv); }
        uint8_t den = 0;
        for (int i = 1; i <= el; i += 2) {
            uint8_t pw = 1;
            for (int k = 0; k < i - 1; k++) pw = gf_mul(gf, pw, xinv);
            den ^= gf_mul(gf, Lm[i], pw);
        }
        if (den == 0) return -1;
        ev[e] = gf_mul(gf, gf_mul(gf, X, num), gf_inv(gf, den));
    }
    int be = 0;
    for (int e = 0; e < np; e++) {
        const int at = j + s * pos[e];
        buf[at] ^= ev[e];
        glob[at] = buf[at];
        if (patches) {
            const int idx = atomicAdd(n_patches, 1);
            if (idx < patch_cap) patches[idx] = RsPatch{glob_off + (uint32_t)at, buf[at]};
        }
        be += __popc((unsigned)ev[e]);
    }
    *bit_errors = be;
    return np;
}


// Everything behind the alignment in ONE launch.  A CTA takes SF_PER_CTA superframe slots of one sub-channel: its warps stage the
// 120 s bytes of their superframes in shared memory (coalesced); then, with the whole CTA re-divided for each phase so that the lanes
// are full, (1) one thread per RS(120,110) codeword -- SF_PER_CTA x s of them -- computes the remainder modulo g(x) (three table words
// per received byte) and runs Berlekamp-Massey / Chien / Forney only when it is non-zero; (2) one thread per superframe checks the
// fire code again and parses the header into the access-unit table; (3) one thread per access unit checks its CRC (slicing by four,
// straight from shared memory: the dependent loop no longer waits for global loads).  Was four launches (k_rs_decode, k_au_split,
// k_au_crc + the statistics), each of them short and latency-bound at the very end of a step.
#define SF_PER_CTA 8
#define SF_MAX_BYTES (120 * 24)            /* s <= 24 (192 kbit/s) */
__global__ void __launch_bounds__(32 * SF_PER_CTA) k_sf_decode(uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs,
                                                                SfInfo* __restrict__ sf, const SfTables* __restrict__ tabs_g,
                                                                int* __restrict__ naus, DabPlusTotals* __restrict__ totals,
                                                                RsPatch* __restrict__ patches, int patch_cap)
{
    __shared__ SfTables tabs;
    __shared__ __align__(16) uint8_t sfbuf[SF_PER_CTA][SF_MAX_BYTES];
    __shared__ int s_tot[5];
    __shared__ int s_valid[SF_PER_CTA], s_row[SF_PER_CTA], s_rs[SF_PER_CTA][3], s_num[SF_PER_CTA], s_st[SF_PER_CTA][7];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t < 5) s_tot[t] = 0;
    if (t < SF_PER_CTA) { s_rs[t][0] = s_rs[t][1] = s_rs[t][2] = 0; s_num[t] = 0; }
    for (int i = t; i < (int)(sizeof(SfTables) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&tabs)[i] = __ldg(reinterpret_cast<const uint32_t*>(tabs_g) + i);
    const DabPlusJob job = jobs[blockIdx.y];
    const int s = job.s, bytes = 120 * s;
    const int m0 = blockIdx.x * SF_PER_CTA;
    {   // phase 0: warp w stages superframe m0 + w (five consecutive logical frames)
        const int m = m0 + warp;
        int valid = 0, row = 0;
        if (m < job.max_sf) { const SfInfo* info = sf + job.sf_off + m; valid = info->valid && s <= 24; row = info->row; }
        if (lane == 0) { s_valid[warp] = valid; s_row[warp] = row; }
        if (valid) {
            const uint8_t* glob = frames + job.frames_off + (size_t)row * job.pitch;
            for (int i = lane * 8; i < bytes; i += 256) *reinterpret_cast<uint2*>(sfbuf[warp] + i) = *reinterpret_cast<const uint2*>(glob + i);
        }
    }
    __syncthreads();
    // phase 1: Reed-Solomon, one thread per codeword (w, j): bytes j, j + s, j + 2 s, ... of superframe w
    if (t < SF_PER_CTA * s) {
        const int w = t / s, j = t - w * s;
        if (s_valid[w]) {
            uint8_t* buf = sfbuf[w];
            // Remainder of the received word modulo g(x), Horner style: R <- R x + c - t g(x) with t the coefficient that leaves at the
            // top.  R = 0 <=> all syndromes are 0 (the common case).
            uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll 4
            for (int k = 0; k < 120; k++) {
                const uint32_t c = buf[j + s * k];
                const uint32_t tt = (w2 >> 8) & 0xffu;
                w2 = ((w2 << 8) | (w1 >> 24)) & 0xffffu;
                w1 = (w1 << 8) | (w0 >> 24);
                w0 = (w0 << 8) | c;
                w0 ^= tabs.fb[tt][0]; w1 ^= tabs.fb[tt][1]; w2 ^= tabs.fb[tt][2];
            }
            if (w0 | w1 | w2) {
                int be = 0;
                const size_t goff = (size_t)job.frames_off + (size_t)s_row[w] * job.pitch;
                const int np = rs_correct(buf, frames + goff, j, s, w0, w1, w2, tabs.gf, &be, patches, patch_cap, &totals->n_patches, (uint32_t)goff);
                if (np < 0) atomicAdd(&s_rs[w][2], 1);
                else { atomicAdd(&s_rs[w][0], np); atomicAdd(&s_rs[w][1], be); }
            }
        }
    }
    __syncthreads();                                    // corrections are in shared memory
    // phase 2: fire code after RS, header -> access-unit table, one thread per superframe
    if (t < SF_PER_CTA && m0 + t < job.max_sf) {
        const int m = m0 + t;
        SfInfo* info = sf + job.sf_off + m;
        int num = 0;
        if (s_valid[t]) {
            const uint8_t* buf = sfbuf[t];
            const int fok = firecode_ok(buf);
            int st[7] = {0, 0, 0, 0, 0, 0, 0};
            if (fok) {
                const int dac = (buf[2] >> 6) & 1, sbr = (buf[2] >> 5) & 1;
                num = dac ? (sbr ? 3 : 6) : (sbr ? 2 : 4);
                st[0] = dac ? (sbr ? 6 : 11) : (sbr ? 5 : 8);
#pragma unroll
                for (int i = 1; i < 6; i++) {
                    if (i < num) {
                        const int bitpos = 24 + 12 * (i - 1), byte = bitpos >> 3;
                        st[i] = (bitpos & 7) ? (((buf[byte] & 0x0F) << 8) | buf[byte + 1]) : ((buf[byte] << 4) | (buf[byte + 1] >> 4));
                    }
                }
#pragma unroll
                for (int i = 0; i < 7; i++) if (i == num) st[i] = 110 * s;
                bool good = true;
#pragma unroll
                for (int i = 0; i < 6; i++) if (i < num && (st[i + 1] <= st[i] + 2 || st[i + 1] > 110 * s)) good = false;
                if (!good) num = 0;
            }
#pragma unroll
            for (int i = 0; i < 7; i++) s_st[t][i] = st[i];
            s_num[t] = num;
            info->rs_corrected = s_rs[t][0]; info->rs_bit_errors = s_rs[t][1]; info->rs_uncorrectable = s_rs[t][2];
            info->firecode_ok = fok;
            info->header = buf[2];
            info->n_aus = num;
            atomicAdd(&s_tot[0], 1);
            if (s_rs[t][0]) atomicAdd(&s_tot[1], s_rs[t][0]);
            if (s_rs[t][1]) atomicAdd(&s_tot[2], s_rs[t][1]);
            if (s_rs[t][2]) atomicAdd(&s_tot[3], s_rs[t][2]);
            if (!fok) atomicAdd(&s_tot[4], 1);
        }
        naus[job.sf_off + m] = num;
    }
    __syncthreads();
    // phase 3: access-unit CRCs, one thread per (superframe, access unit)
    if (t < SF_PER_CTA * 6) {
        const int w = t / 6, a = t - 6 * w;
        if (a < s_num[w]) {
            SfInfo* info = sf + job.sf_off + m0 + w;
            const int my_off = s_st[w][a], my_end = s_st[w][a + 1];
            const uint8_t* p = sfbuf[w] + my_off;
            const int len = my_end - my_off - 2;
            uint32_t crc = 0xFFFF;
            int i = 0;
            for (; i + 4 <= len; i += 4)
                crc = (uint32_t)tabs.crc[3][((crc >> 8) ^ p[i]) & 0xff] ^ tabs.crc[2][(crc ^ p[i + 1]) & 0xff] ^ tabs.crc[1][p[i + 2]] ^ tabs.crc[0][p[i + 3]];
            for (; i < len; i++) crc = ((crc << 8) & 0xffffu) ^ tabs.crc[0][((crc >> 8) ^ p[i]) & 0xff];
            crc = (~crc) & 0xffffu;
            info->au_off[a] = my_off; info->au_len[a] = len;
            info->au_ok[a] = crc == (uint32_t)((p[len] << 8) | p[len + 1]);
        }
    }
    // RS statistics: one set of global atomics per CTA
    if (t < 5 && s_tot[t]) atomicAdd(&totals->superframes + t, s_tot[t]);     // superframes, rs_corrected, rs_bit_errors, rs_uncorrectable, firecode_fail
}

// Access-unit list in (sub-channel, superframe) order -- the layout of the result (b200dab_au_t): every CTA adds up the access-unit
// counts of the slots in front of its own 256 (a coalesced read of a few KB), scans its own, and every thread writes the records of
// its slot.  One short, fully parallel launch instead of a count / write pair whose second half walked the CTA sums serially.
__global__ void __launch_bounds__(256) k_au_list(const DabPlusJob* __restrict__ jobs, int n_jobs, const SfInfo* __restrict__ sf, const int* __restrict__ naus,
                                                 int total_sf, AuRecord* __restrict__ out, DabPlusTotals* __restrict__ totals)
{
    __shared__ int red[8];
    __shared__ int s_base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int first = blockIdx.x * 256, i = first + t;
    int before = 0;
    for (int k = t; k < first; k += 256) before += __ldg(naus + k);
    const int n = i < total_sf ? __ldg(naus + i) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 0) red[warp] = before;
    __syncthreads();
    if (t == 0) { int b = 0; for (int w = 0; w < 8; w++) b += red[w]; s_base = b; }
    __syncthreads();
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    int pos = s_base + incl - n;
    for (int w = 0; w < warp; w++) pos += red[w];
    if (t == 255) atomicAdd(&totals->n_aus, pos + n - s_base);
    if (n > 0) {
        int k = 0;
        while (k + 1 < n_jobs && i >= jobs[k + 1].sf_off) k++;
        const SfInfo info = sf[i];
        const int m = i - jobs[k].sf_off, pitch = jobs[k].pitch, sub = jobs[k].subch_index;
        for (int a = 0; a < n; a++) {
            AuRecord r;
            r.subch_index = sub; r.superframe = m; r.cif = info.row;
            r.offset = info.row * pitch + info.au_off[a]; r.length = info.au_len[a]; r.crc_ok = info.au_ok[a]; r.header = info.header;
            out[pos + a] = r;
        }
    }
}
