// tables.h -- host-side construction of the constant tables of DAB mode I (ETSI EN 300 401) used by the kernels,
// and the protection-profile tables (EEP: 11.3.2, UEP: table 8 + table 31).  Values validated against libdabsdr 4.0.1
// (SURVEY.md Appendix A); the UEP table can also be read at .rodata+0xE40 of that binary.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/b200dab.h"

namespace dabtab {

static const int8_t H_TAB[4][32] = {
    {0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1, 0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1},
    {0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0, 0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0},
    {0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3, 0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3},
    {0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2, 0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2}};
static const int8_t PRS_NEG[24][2] = {{0,1},{1,2},{2,0},{3,1},{0,3},{1,2},{2,2},{3,3},{0,2},{1,1},{2,2},{3,3},
                                      {0,1},{1,2},{2,3},{3,3},{0,2},{1,2},{2,2},{3,1},{0,1},{1,3},{2,1},{3,2}};
static const int8_t PRS_POS[24][2] = {{0,3},{3,1},{2,1},{1,1},{0,2},{3,2},{2,1},{1,0},{0,2},{3,2},{2,3},{1,3},
                                      {0,0},{3,2},{2,1},{1,3},{0,3},{3,3},{2,3},{1,0},{0,3},{3,0},{2,1},{1,1}};
static const uint32_t PI_MASK[24] = {
    0xC8888888u,0xC888C888u,0xC8C8C888u,0xC8C8C8C8u,0xCCC8C8C8u,0xCCC8CCC8u,0xCCCCCCC8u,0xCCCCCCCCu,
    0xECCCCCCCu,0xECCCECCCu,0xECECECCCu,0xECECECECu,0xEEECECECu,0xEEECEEECu,0xEEEEEEECu,0xEEEEEEEEu,
    0xFEEEEEEEu,0xFEEEFEEEu,0xFEFEFEEEu,0xFEFEFEFEu,0xFFFEFEFEu,0xFFFEFFFEu,0xFFFFFFFEu,0xFFFFFFFFu};

inline void prs_quarter(int8_t* q /*1536*/)
{
    for (int b = 0; b < 24; b++)
        for (int j = 0; j < 32; j++) {
            q[b * 32 + j] = (int8_t)((H_TAB[PRS_NEG[b][0]][j] + PRS_NEG[b][1]) & 3);
            q[768 + b * 32 + j] = (int8_t)((H_TAB[PRS_POS[b][0]][j] + PRS_POS[b][1]) & 3);
        }
}

// QPSK symbol n -> carrier index c (ascending k)
inline void freq_deint(int16_t* c_of_n /*1536*/)
{
    int pi = 0, n = 0;
    for (int i = 1; i < 2048; i++) {
        pi = (13 * pi + 511) & 2047;
        if (pi >= 256 && pi <= 1792 && pi != 1024) {
            int k = pi - 1024;
            c_of_n[n++] = (int16_t)(k < 0 ? k + 768 : k + 767);
        }
    }
}

inline void twiddles(float* tw /*2048*2*/)
{
    for (int k = 0; k < 2048; k++) {
        double a = -2.0 * M_PI * k / 2048.0;
        tw[2 * k] = (float)std::cos(a);
        tw[2 * k + 1] = (float)std::sin(a);
    }
}

inline void prbs_bits(uint8_t* bits, int n)
{
    unsigned reg = 0x1FF;
    for (int i = 0; i < n; i++) {
        unsigned b = ((reg >> 8) ^ (reg >> 4)) & 1;
        bits[i] = (uint8_t)b;
        reg = ((reg << 1) | b) & 0x1FF;
    }
}

// little-endian 32-bit words of the PRBS byte stream (bits MSB first inside each byte)
inline void prbs_words(uint32_t* w, int nwords)
{
    std::vector<uint8_t> bits((size_t)nwords * 32);
    prbs_bits(bits.data(), nwords * 32);
    for (int c = 0; c < nwords; c++) {
        uint32_t v = 0;
        for (int byte = 0; byte < 4; byte++) {
            unsigned x = 0;
            for (int b = 0; b < 8; b++) x = (x << 1) | bits[(size_t)c * 32 + byte * 8 + b];
            v |= x << (8 * byte);
        }
        w[c] = v;
    }
}

inline void gf_tables(uint8_t* exp512, uint8_t* log256)
{
    int x = 1;
    std::memset(log256, 0, 256);
    for (int i = 0; i < 255; i++) {
        exp512[i] = (uint8_t)x; log256[x] = (uint8_t)i;
        x <<= 1; if (x & 0x100) x ^= 0x11D;
    }
    for (int i = 255; i < 512; i++) exp512[i] = exp512[i - 255];
}

inline void tii_patterns(uint8_t* pat /*70*/)
{
    int n = 0;
    for (int v = 0; v < 256; v++) if (__builtin_popcount(v) == 4) pat[n++] = (uint8_t)v;
}

// ---- protection profiles
struct UepRow { uint16_t bitrate, size_cu; uint8_t level; uint8_t L[4]; uint8_t PI[4]; uint8_t pad; };
static const UepRow UEP[64] = {
    {32,16,5,{3,4,17,0},{5,3,2,0},0},{32,21,4,{3,3,18,0},{11,6,5,0},0},{32,24,3,{3,4,14,3},{15,9,6,8},0},{32,29,2,{3,4,14,3},{22,13,8,13},0},{32,35,1,{3,5,13,3},{24,17,12,17},4},
    {48,24,5,{4,3,26,3},{5,4,2,3},0},{48,29,4,{3,4,26,3},{9,6,4,6},0},{48,35,3,{3,4,26,3},{15,10,6,9},4},{48,42,2,{3,4,26,3},{24,14,8,15},0},{48,52,1,{3,5,25,3},{24,18,13,18},0},
    {56,29,5,{6,10,23,3},{5,4,2,3},0},{56,35,4,{6,10,23,3},{9,6,4,5},0},{56,42,3,{6,12,21,3},{16,7,6,9},0},{56,52,2,{6,10,23,3},{23,13,8,13},8},
    {64,32,5,{6,9,31,2},{5,3,2,3},0},{64,42,4,{6,9,33,0},{11,6,5,0},0},{64,48,3,{6,12,27,3},{16,8,6,9},0},{64,58,2,{6,10,29,3},{23,13,8,13},8},{64,70,1,{6,11,28,3},{24,18,12,18},4},
    {80,40,5,{6,10,41,3},{6,3,2,3},0},{80,52,4,{6,10,41,3},{11,6,5,6},0},{80,58,3,{6,11,40,3},{16,8,6,7},0},{80,70,2,{6,10,41,3},{23,13,8,13},8},{80,84,1,{6,10,41,3},{24,17,12,18},4},
    {96,48,5,{7,9,53,3},{5,4,2,4},0},{96,58,4,{7,10,52,3},{9,6,4,6},0},{96,70,3,{6,12,51,3},{16,9,6,10},4},{96,84,2,{6,10,53,3},{22,12,9,12},0},{96,104,1,{6,13,50,3},{24,18,13,19},0},
    {112,58,5,{14,17,50,3},{5,4,2,5},0},{112,70,4,{11,21,49,3},{9,6,4,8},0},{112,84,3,{11,23,47,3},{16,8,6,9},0},{112,104,2,{11,21,49,3},{23,12,9,14},4},
    {128,64,5,{12,19,62,3},{5,3,2,4},0},{128,84,4,{11,21,61,3},{11,6,5,7},0},{128,96,3,{11,22,60,3},{16,9,6,10},4},{128,116,2,{11,21,61,3},{22,12,9,14},0},{128,140,1,{11,20,62,3},{24,17,13,19},8},
    {160,80,5,{11,19,87,3},{5,4,2,4},0},{160,104,4,{11,23,83,3},{11,6,5,9},0},{160,116,3,{11,24,82,3},{16,8,6,11},0},{160,140,2,{11,21,85,3},{22,11,9,13},0},{160,168,1,{11,22,84,3},{24,18,12,19},0},
    {192,96,5,{11,20,110,3},{6,4,2,5},0},{192,116,4,{11,22,108,3},{10,6,4,9},0},{192,140,3,{11,24,106,3},{16,10,6,11},0},{192,168,2,{11,20,110,3},{22,13,9,13},8},{192,208,1,{11,21,109,3},{24,20,13,24},0},
    {224,116,5,{12,22,131,3},{8,6,2,6},4},{224,140,4,{12,26,127,3},{12,8,4,11},0},{224,168,3,{11,20,134,3},{16,10,7,9},0},{224,208,2,{11,22,132,3},{24,16,10,15},0},{224,232,1,{11,24,130,3},{24,20,12,20},4},
    {256,128,5,{11,24,154,3},{6,5,2,5},0},{256,168,4,{11,24,154,3},{12,9,5,10},4},{256,192,3,{11,27,151,3},{16,10,7,10},0},{256,232,2,{11,22,156,3},{24,14,10,13},8},{256,280,1,{11,26,152,3},{24,19,14,18},4},
    {320,160,5,{11,26,200,3},{8,5,2,6},4},{320,208,4,{11,25,201,3},{13,9,5,10},8},{320,280,2,{11,26,200,3},{24,17,9,17},0},
    {384,192,5,{11,27,247,3},{8,6,2,7},0},{384,280,3,{11,24,250,3},{16,9,7,10},4},{384,416,1,{12,28,245,3},{24,20,14,23},8}};

inline int profile_eep(int bitrate, int level, int option, b200dab_subch_t* o)
{
    if (level < 1 || level > 4 || bitrate <= 0) return -1;
    o->bitrate = bitrate; o->n_seg = 2; o->uep_index = -1; o->prot_level = (level - 1) + (option ? 4 : 0);
    for (int i = 0; i < 4; i++) { o->seg_len[i] = 0; o->seg_pi[i] = 0; }
    if (option == 0) {
        if (bitrate % 8) return -1;
        int n = bitrate / 8;
        switch (level) {
        case 1: o->seg_len[0] = 6 * n - 3; o->seg_pi[0] = 24; o->seg_len[1] = 3; o->seg_pi[1] = 23; o->size_cu = 12 * n; break;
        case 2:
            if (n == 1) { o->seg_len[0] = 5; o->seg_pi[0] = 13; o->seg_len[1] = 1; o->seg_pi[1] = 12; }
            else { o->seg_len[0] = 2 * n - 3; o->seg_pi[0] = 14; o->seg_len[1] = 4 * n + 3; o->seg_pi[1] = 13; }
            o->size_cu = 8 * n; break;
        case 3: o->seg_len[0] = 6 * n - 3; o->seg_pi[0] = 8; o->seg_len[1] = 3; o->seg_pi[1] = 7; o->size_cu = 6 * n; break;
        default: o->seg_len[0] = 4 * n - 3; o->seg_pi[0] = 3; o->seg_len[1] = 2 * n + 3; o->seg_pi[1] = 2; o->size_cu = 4 * n; break;
        }
    } else {
        if (bitrate % 32) return -1;
        int m = bitrate / 32;
        static const int pis[4][3] = {{10, 9, 27}, {6, 5, 21}, {4, 3, 18}, {2, 1, 15}};
        o->seg_len[0] = 24 * m - 3; o->seg_pi[0] = pis[level - 1][0];
        o->seg_len[1] = 3; o->seg_pi[1] = pis[level - 1][1];
        o->size_cu = pis[level - 1][2] * m;
    }
    return 0;
}

inline int profile_uep(int idx, b200dab_subch_t* o)
{
    if (idx < 0 || idx > 63) return -1;
    const UepRow& r = UEP[idx];
    o->bitrate = r.bitrate; o->size_cu = r.size_cu; o->uep_index = idx; o->prot_level = r.level; o->n_seg = 0;
    for (int i = 0; i < 4; i++) { o->seg_len[i] = 0; o->seg_pi[i] = 0; }
    for (int i = 0; i < 4; i++) if (r.L[i]) { o->seg_len[o->n_seg] = r.L[i]; o->seg_pi[o->n_seg] = r.PI[i]; o->n_seg++; }
    return 0;
}

// stepinfo[k] = (index of the first kept bit of step k in the punctured stream) << 4 | keep mask of the 4 mother bits;
// stepinfo[steps] = total << 4.  Returns the number of punctured bits, or -1 if the profile does not match n_info.
inline int build_stepinfo(const b200dab_subch_t& sc, std::vector<uint32_t>& out)
{
    const int n_info = sc.bitrate * 24, steps = n_info + 6;
    std::vector<uint8_t> keep;
    keep.reserve((size_t)steps * 4);
    for (int s = 0; s < sc.n_seg; s++) {
        if (sc.seg_pi[s] < 1 || sc.seg_pi[s] > 24) return -1;
        uint32_t m = PI_MASK[sc.seg_pi[s] - 1];
        for (int l = 0; l < sc.seg_len[s] * 4; l++)
            for (int i = 0; i < 32; i++) keep.push_back((m >> (31 - i)) & 1);
    }
    for (int i = 0; i < 24; i++) keep.push_back((0xCCCCCCu >> (23 - i)) & 1);
    if ((int)keep.size() != steps * 4) return -1;
    int idx = 0;
    for (int k = 0; k < steps; k++) {
        uint32_t mask = 0;
        int first = idx;
        for (int q = 0; q < 4; q++) if (keep[(size_t)4 * k + q]) { mask |= 1u << q; idx++; }
        out.push_back(((uint32_t)first << 4) | mask);
    }
    out.push_back((uint32_t)idx << 4);
    return idx;
}

}  // namespace dabtab
