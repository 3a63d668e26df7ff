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

// firecode over bytes 2..10 of every logical frame (pre-RS, like the reference, which needs it to find alignment)
__global__ void __launch_bounds__(256) k_firecode_scan(const uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs,
                                                       uint8_t* __restrict__ flags)
{
    const DabPlusJob job = jobs[blockIdx.y];
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= job.n_rows) return;
    const uint8_t* f = frames + job.frames_off + (size_t)r * job.pitch;
    uint16_t fc = crc16_ccitt_dev(f + 2, 9, 0, 0x782F);
    flags[job.flag_off + r] = (fc == (uint16_t)((f[0] << 8) | f[1])) && ((f[0] | f[1] | f[2]) != 0);
}

// one CTA per job: pick the residue (row mod 5) with most firecode hits; mark the superframe slots
#define SF_ALIGN_THREADS 256
__global__ void __launch_bounds__(SF_ALIGN_THREADS) k_sf_align(const DabPlusJob* __restrict__ jobs, const uint8_t* __restrict__ flags,
                                                               SfInfo* __restrict__ sf, int* __restrict__ residue_out)
{
    __shared__ int part[SF_ALIGN_THREADS / 32][5];
    __shared__ int s_best;
    const DabPlusJob job = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    int cnt[5] = {0, 0, 0, 0, 0};
    for (int r = tid; r < job.n_rows; r += SF_ALIGN_THREADS)
        if (flags[job.flag_off + r]) cnt[r % 5]++;
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
            for (int w = 0; w < SF_ALIGN_THREADS / 32; w++) c += part[w][k];
            if (c > bc) { bc = c; best = k; }
        }
        if (job.forced_phase >= 0) best = job.forced_phase % 5;
        s_best = best;
        residue_out[blockIdx.x] = best;
    }
    __syncthreads();
    const int best = s_best;
    for (int m = tid; m < job.max_sf; m += SF_ALIGN_THREADS) {
        SfInfo* o = sf + job.sf_off + m;
        const int row = best + 5 * m;
        o->valid = (best >= 0 && row + 5 <= job.n_rows) ? 1 : 0;
        o->row = row;
        o->rs_corrected = o->rs_bit_errors = o->rs_uncorrectable = 0;
        o->firecode_ok = 0; o->n_aus = 0; o->header = 0;
    }
}

// ---- GF(2^8) / 0x11D
struct GfTables { uint8_t exp[512]; uint8_t log[256]; };
__device__ __forceinline__ uint8_t gf_mul(const GfTables& g, uint8_t a, uint8_t b) { return (a && b) ? g.exp[g.log[a] + g.log[b]] : 0; }
__device__ __forceinline__ uint8_t gf_inv(const GfTables& g, uint8_t a) { return g.exp[255 - g.log[a]]; }

// one thread per RS codeword; grid (ceil(max_sf * s / 128), n_jobs), 128 threads
__global__ void __launch_bounds__(128) k_rs_decode(uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs,
                                                   SfInfo* __restrict__ sf, const GfTables* __restrict__ gf_g)
{
    __shared__ GfTables gf;
    __shared__ uint8_t gen[12];                 // g(x) = prod_{i < 10} (x - alpha^i) = x^10 + gen[9] x^9 + ... + gen[0]
    __shared__ uint32_t fb[256][3];             // fb[t] = t * (gen[0..3] | gen[4..7] | gen[8..9]): what t * g(x) adds to the remainder
    for (int i = threadIdx.x; i < (int)sizeof(GfTables); i += blockDim.x) reinterpret_cast<uint8_t*>(&gf)[i] = reinterpret_cast<const uint8_t*>(gf_g)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t g[12] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};           // g = 1, then times (x + alpha^i)
        for (int i = 0; i < 10; i++) {
            const uint8_t a = gf.exp[i];
            for (int j = i + 1; j > 0; j--) g[j] = g[j - 1] ^ gf_mul(gf, g[j], a);
            g[0] = gf_mul(gf, g[0], a);
        }
        for (int j = 0; j < 12; j++) gen[j] = j < 10 ? g[j] : 0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        uint32_t w[3] = {0, 0, 0};
        for (int j = 0; j < 10; j++) w[j >> 2] |= (uint32_t)gf_mul(gf, (uint8_t)t, gen[j]) << (8 * (j & 3));
        fb[t][0] = w[0]; fb[t][1] = w[1]; fb[t][2] = w[2];
    }
    __syncthreads();
    const DabPlusJob job = jobs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= job.max_sf * job.s) return;
    const int m = idx / job.s, j = idx - m * job.s;
    SfInfo* info = sf + job.sf_off + m;
    if (!info->valid) return;
    uint8_t* base = frames + job.frames_off + (size_t)info->row * job.pitch;
    const int s = job.s, n = 120, nroots = 10;
    // Remainder of the received word modulo g(x), Horner style: R <- R x + c - t g(x) with t the coefficient that leaves at the top.
    // The ten remainder bytes live in three words and t g(x) is one table row, so a received byte costs three shared-memory loads
    // instead of the twenty of ten separate syndrome recursions.  R = 0 <=> all syndromes are 0 (the common case); otherwise the
    // syndromes are R evaluated at alpha^i, since g(alpha^i) = 0.
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (int k = 0; k < n; k++) {
        const uint32_t c = base[j + s * k];
        const uint32_t t = (w2 >> 8) & 0xffu;
        w2 = ((w2 << 8) | (w1 >> 24)) & 0xffffu;
        w1 = (w1 << 8) | (w0 >> 24);
        w0 = (w0 << 8) | c;
        w0 ^= fb[t][0]; w1 ^= fb[t][1]; w2 ^= fb[t][2];
    }
    if (!(w0 | w1 | w2)) return;
    uint8_t S[10];
    {
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
    if (el > nroots / 2) { atomicAdd(&info->rs_uncorrectable, 1); return; }
    // Chien search over the 120 positions
    int pos[5], np = 0;
    for (int k = 0; k < n; k++) {
        const int p = n - 1 - k;
        const uint8_t xinv = gf.exp[(255 - p) % 255];
        uint8_t v = 0, xp = 1;
        for (int i = 0; i <= el; i++) { v ^= gf_mul(gf, Lm[i], xp); xp = gf_mul(gf, xp, xinv); }
        if (v == 0) { if (np < 5) pos[np] = k; np++; }
    }
    if (np != el) { atomicAdd(&info->rs_uncorrectable, 1); return; }
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
        if (den == 0) { atomicAdd(&info->rs_uncorrectable, 1); return; }
        ev[e] = gf_mul(gf, gf_mul(gf, X, num), gf_inv(gf, den));
    }
    int be = 0;
    for (int e = 0; e < np; e++) { base[j + s * pos[e]] ^= ev[e]; be += __popc((unsigned)ev[e]); }
    atomicAdd(&info->rs_corrected, np);
    atomicAdd(&info->rs_bit_errors, be);
}

// one thread per superframe slot: firecode after RS, header parse -> AU offsets (the CRCs follow in k_au_crc)
__global__ void __launch_bounds__(64) k_au_split(const uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs, SfInfo* __restrict__ sf)
{
    const DabPlusJob job = jobs[blockIdx.y];
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= job.max_sf) return;
    SfInfo* info = sf + job.sf_off + m;
    if (!info->valid) return;
    const uint8_t* p = frames + job.frames_off + (size_t)info->row * job.pitch;
    const int s = job.s;
    const uint16_t fc = crc16_ccitt_dev(p + 2, 9, 0, 0x782F);
    const int fok = (fc == (uint16_t)((p[0] << 8) | p[1])) && ((p[0] | p[1] | p[2]) != 0);
    info->firecode_ok = fok;
    info->header = p[2];
    if (!fok) return;
    const int dac = (p[2] >> 6) & 1, sbr = (p[2] >> 5) & 1;
    const int num = dac ? (sbr ? 3 : 6) : (sbr ? 2 : 4);
    const int first = dac ? (sbr ? 6 : 11) : (sbr ? 5 : 8);
    int st[7];
    st[0] = first;
    for (int i = 1; i < num; i++) {
        const int bitpos = 24 + 12 * (i - 1), byte = bitpos >> 3;
        st[i] = (bitpos & 7) ? (((p[byte] & 0x0F) << 8) | p[byte + 1]) : ((p[byte] << 4) | (p[byte + 1] >> 4));
    }
    st[num] = 110 * s;
    for (int i = 0; i < num; i++) if (st[i + 1] <= st[i] + 2 || st[i + 1] > 110 * s) return;
    for (int i = 0; i < num; i++) { info->au_off[i] = st[i]; info->au_len[i] = st[i + 1] - st[i] - 2; info->au_ok[i] = 0; }
    info->n_aus = num;
}

// one thread per (superframe, access unit): table-driven CRC-16 (0x1021, init 0xFFFF, inverted) over the AU payload
__global__ void __launch_bounds__(128) k_au_crc(const uint8_t* __restrict__ frames, const DabPlusJob* __restrict__ jobs, SfInfo* __restrict__ sf)
{
    // slicing by four: tbl[k][v] = CRC state after byte v followed by k zero bytes, so four message bytes are folded in with four
    // independent look-ups instead of a chain of four
    __shared__ uint16_t tbl[4][256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t)i << 8;
#pragma unroll
        for (int b = 0; b < 8; b++) c = (c & 0x8000u) ? ((c << 1) ^ 0x1021u) & 0xffffu : (c << 1) & 0xffffu;
        tbl[0][i] = (uint16_t)c;
    }
    __syncthreads();
    for (int k = 1; k < 4; k++) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            const uint32_t c = tbl[k - 1][i];
            tbl[k][i] = (uint16_t)(((c << 8) & 0xffffu) ^ tbl[0][c >> 8]);
        }
        __syncthreads();
    }
    const DabPlusJob job = jobs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = idx / 6, a = idx - 6 * m;
    if (m >= job.max_sf) return;
    SfInfo* info = sf + job.sf_off + m;
    if (!info->valid || a >= info->n_aus) return;
    const uint8_t* p = frames + job.frames_off + (size_t)info->row * job.pitch + info->au_off[a];
    const int len = info->au_len[a];
    uint32_t crc = 0xFFFF;
    int i = 0;
    for (; i + 4 <= len; i += 4)
        crc = (uint32_t)tbl[3][((crc >> 8) ^ p[i]) & 0xff] ^ tbl[2][(crc ^ p[i + 1]) & 0xff] ^ tbl[1][p[i + 2]] ^ tbl[0][p[i + 3]];
    for (; i < len; i++) crc = ((crc << 8) & 0xffffu) ^ tbl[0][((crc >> 8) ^ p[i]) & 0xff];
    crc = (~crc) & 0xffffu;
    info->au_ok[a] = crc == (uint32_t)((p[len] << 8) | p[len + 1]);
}

// Result records in the layout of the C ABI (b200dab_au_t: seven int32) and the totals the host used to add up from SfInfo.
struct AuRecord { int subch_index, superframe, cif, offset, length, crc_ok, header; };
struct DabPlusTotals { int n_aus, superframes, rs_corrected, rs_bit_errors, rs_uncorrectable, firecode_fail, pad_[2]; };

// Compaction in two small launches over all superframe slots in (job, superframe) order -- the order the host loop produced:
// k_au_count: per CTA of 256 slots, the number of access units and the partial RS statistics;
// k_au_write: prefix over the (<= a few dozen) CTA sums + scan inside the CTA, then every slot writes its records.
__global__ void __launch_bounds__(256) k_au_count(const SfInfo* __restrict__ sf, int total_sf, int* __restrict__ block_sums)
{
    __shared__ int acc[6];
    const int t = threadIdx.x, i = blockIdx.x * 256 + t;
    if (t < 6) acc[t] = 0;
    __syncthreads();
    int v[6] = {0, 0, 0, 0, 0, 0};
    if (i < total_sf && sf[i].valid) {
        const SfInfo* p = sf + i;
        v[0] = p->n_aus; v[1] = 1; v[2] = p->rs_corrected; v[3] = p->rs_bit_errors; v[4] = p->rs_uncorrectable; v[5] = p->firecode_ok ? 0 : 1;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
        int x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((t & 31) == 0 && x) atomicAdd(&acc[q], x);
    }
    __syncthreads();
    if (t < 6) block_sums[blockIdx.x * 6 + t] = acc[t];
}

__global__ void __launch_bounds__(256) k_au_write(const DabPlusJob* __restrict__ jobs, int n_jobs, const SfInfo* __restrict__ sf, int total_sf,
                                                  const int* __restrict__ block_sums, AuRecord* __restrict__ out, DabPlusTotals* __restrict__ totals)
{
    __shared__ int warp_tot[8];
    __shared__ int job_off[65], job_pitch[64], job_sub[64];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, i = blockIdx.x * 256 + t;
    if (n_jobs > 64) n_jobs = 64;
    if (t < n_jobs) { job_off[t] = jobs[t].sf_off; job_pitch[t] = jobs[t].pitch; job_sub[t] = jobs[t].subch_index; }
    if (t == n_jobs) job_off[t] = total_sf;
    int base = 0;
    for (int b = 0; b < (int)blockIdx.x; b++) base += block_sums[b * 6];
    if (blockIdx.x == 0 && t == 0) {
        DabPlusTotals r = {0, 0, 0, 0, 0, 0, {0, 0}};
        for (int b = 0; b < (int)gridDim.x; b++) {
            const int* q = block_sums + b * 6;
            r.n_aus += q[0]; r.superframes += q[1]; r.rs_corrected += q[2]; r.rs_bit_errors += q[3]; r.rs_uncorrectable += q[4]; r.firecode_fail += q[5];
        }
        *totals = r;
    }
    SfInfo info;
    int n = 0;
    if (i < total_sf) { info = sf[i]; if (info.valid) n = info.n_aus; }
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int pos = base + incl - n;
    for (int w = 0; w < warp; w++) pos += warp_tot[w];
    if (n > 0) {
        int k = 0;
        while (k + 1 < n_jobs && i >= job_off[k + 1]) k++;
        const int m = i - job_off[k], pitch = job_pitch[k], sub = job_sub[k];
        for (int a = 0; a < n; a++) {
            AuRecord r;
            r.subch_index = sub; r.superframe = m; r.cif = info.row;
            r.offset = info.row * pitch + info.au_off[a]; r.length = info.au_len[a]; r.crc_ok = info.au_ok[a]; r.header = info.header;
            out[pos + a] = r;
        }
    }
}
