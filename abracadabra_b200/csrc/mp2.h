// mp2.h -- ISO/IEC 11172-3 Layer II frame check for DAB (MP2) audio sub-channels: header sanity + CRC-16 over the header's
// last 16 bits, the bit allocation and the scale-factor selection information.  Byte-level host work done by the drop-in
// before an MP2 frame is handed to the audio callback (dabsdr.h:71-78); the reference core delivers auLen = 0 for a frame
// that fails this check.
#pragma once
#include <cstdint>

namespace mp2 {

struct BitReader {
    const uint8_t* p; int n; int pos = 0;
    BitReader(const uint8_t* p_, int n_) : p(p_), n(n_) {}
    int get(int bits) { int v = 0; for (int i = 0; i < bits; i++, pos++) v = (v << 1) | ((pos >> 3) < n ? (p[pos >> 3] >> (7 - (pos & 7))) & 1 : 0); return v; }
};

inline uint16_t crc_update(uint16_t crc, int value, int bits)
{
    for (int i = bits - 1; i >= 0; i--) {
        const int b = ((value >> i) & 1) ^ (crc >> 15);
        crc = (uint16_t)(crc << 1);
        if (b) crc ^= 0x8005;
    }
    return crc;
}

// bits per allocation field, by sub-band, for the four MPEG-1 tables (B.2a-d) and the MPEG-2 LSF table
inline int nbal(int table, int sb)
{
    switch (table) {
    case 0: case 1: return sb < 11 ? 4 : sb < 23 ? 3 : 2;     // B.2a (27 sub-bands), B.2b (30)
    case 2: case 3: return sb < 2 ? 4 : 3;                    // B.2c (8), B.2d (12)
    default: return sb < 4 ? 4 : sb < 11 ? 3 : 2;             // LSF (30)
    }
}
inline int sblimit(int table) { static const int l[5] = {27, 30, 8, 12, 30}; return l[table]; }

// returns 1 = CRC ok, 0 = CRC mismatch, -1 = not a Layer II frame with CRC protection
inline int check_frame(const uint8_t* f, int len)
{
    if (len < 8 || f[0] != 0xFF || (f[1] & 0xF0) != 0xF0) return -1;
    const int id = (f[1] >> 3) & 1, layer = (f[1] >> 1) & 3, prot = f[1] & 1;
    if (layer != 2) return -1;                                 // '10' = Layer II
    if (prot) return -1;                                       // no CRC in the stream
    const int br_idx = f[2] >> 4, fs_idx = (f[2] >> 2) & 3, mode = f[3] >> 6, mode_ext = (f[3] >> 4) & 3;
    if (br_idx == 0 || br_idx == 15 || fs_idx == 3) return -1;
    const int nch = mode == 3 ? 1 : 2;
    int table;
    if (id == 0) table = 4;
    else {
        static const int br1[15] = {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384};
        const int br_ch = br1[br_idx] / nch;
        if (fs_idx == 1) table = br_ch <= 48 ? 2 : 0;          // 48 kHz
        else if (fs_idx == 0) table = br_ch <= 48 ? 2 : br_ch <= 80 ? 0 : 1;   // 44.1 kHz
        else table = br_ch <= 48 ? 3 : br_ch <= 80 ? 0 : 1;    // 32 kHz
    }
    const int sbl = sblimit(table);
    int bound = sbl;
    if (mode == 1) bound = 4 * (mode_ext + 1);
    if (bound > sbl) bound = sbl;
    uint16_t crc = 0xFFFF;
    crc = crc_update(crc, f[2], 8);
    crc = crc_update(crc, f[3], 8);
    BitReader br(f, len);
    br.pos = 48;                                               // after header (32) and CRC word (16)
    int alloc[30][2];
    for (int sb = 0; sb < sbl; sb++) {
        const int nb = nbal(table, sb);
        if (sb < bound) for (int ch = 0; ch < nch; ch++) { alloc[sb][ch] = br.get(nb); crc = crc_update(crc, alloc[sb][ch], nb); }
        else { alloc[sb][0] = br.get(nb); crc = crc_update(crc, alloc[sb][0], nb); if (nch == 2) alloc[sb][1] = alloc[sb][0]; }
    }
    for (int sb = 0; sb < sbl; sb++)
        for (int ch = 0; ch < nch; ch++)
            if (alloc[sb][ch]) crc = crc_update(crc, br.get(2), 2);
    return crc == (uint16_t)((f[4] << 8) | f[5]) ? 1 : 0;
}

}  // namespace mp2
