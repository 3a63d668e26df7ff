// pad.h -- programme associated data of an audio frame: F-PAD / X-PAD field parsing, dynamic-label and MSC data group
// assembly (ETSI EN 300 401 sec. 7.4; DAB+ carriage: TS 102 563 sec. 5.4).  Host-side byte work of the drop-in: the
// reference core hands complete dynamic-label data groups (2-byte prefix + characters + CRC, CRC already verified) to
// dabsdrDynamicLabelCBFunc_t (dabsdr.h:81-86; consumer src/data/dldecoder.cpp:82-160) and complete MSC data groups of a started
// X-PAD user application to dabsdrDataGroupCBFunc_t (dabsdr.h:89-96; consumer src/data/mscdatagroup.cpp:30-160).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace pad {

inline uint16_t crc16(const uint8_t* p, size_t n)
{
    uint16_t crc = 0xFFFF;
    for (size_t i = 0; i < n; i++) {
        crc ^= (uint16_t)p[i] << 8;
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return (uint16_t)~crc;
}
inline bool crc_ok(const uint8_t* p, size_t n) { return n >= 2 && crc16(p, n - 2) == (uint16_t)((p[n - 2] << 8) | p[n - 1]); }

// DAB+ access unit -> PAD bytes: the PAD travels in a data_stream_element (id 4) that is the first syntactic element of the AU
inline bool dabplus_pad(const uint8_t* au, int len, const uint8_t** out, int* n)
{
    if (len < 3 || (au[0] >> 5) != 4) return false;
    const bool align = au[0] & 1;
    int count = au[1], pos = 2;
    if (count == 255) { count += au[2]; pos = 3; }
    (void)align;                                  // the element header is 16 (24) bits: the data is byte aligned either way
    if (count < 2 || pos + count > len) return false;
    *out = au + pos; *n = count;
    return true;
}

// DAB (MPEG Layer II) audio frame -> PAD bytes: the frame ends with [X-PAD, reversed] [ScF-CRC] [F-PAD byte L-1, byte L]
// (EN 300 401 sec. 7.4 / annex B.3).  The ScF-CRC is 2 bytes below 56 kbit/s per channel at 48 kHz and 4 bytes otherwise.
// `tmp` receives the up to 196 bytes before the ScF-CRC followed by the F-PAD, i.e. the layout Decoder::feed expects.
inline bool mp2_pad(const uint8_t* f, int len, std::vector<uint8_t>& tmp, int* scf_len = nullptr)
{
    if (len < 16 || f[0] != 0xFF || (f[1] & 0xF0) != 0xF0) return false;
    static const int br1[15] = {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384};
    const int id = (f[1] >> 3) & 1, br_idx = f[2] >> 4, nch = (f[3] >> 6) == 3 ? 1 : 2;
    if (br_idx == 0 || br_idx == 15) return false;
    const int scf = (id == 1 && br1[br_idx] / nch < 56) ? 2 : 4;
    if (scf_len) *scf_len = scf;
    const int avail = len - 2 - scf - 6;                         // keep clear of header + CRC
    if (avail <= 0) return false;
    const int nx = avail < 196 ? avail : 196;
    tmp.resize((size_t)nx + 2);
    memcpy(tmp.data(), f + len - 2 - scf - nx, (size_t)nx);
    tmp[(size_t)nx] = f[len - 2]; tmp[(size_t)nx + 1] = f[len - 1];
    return true;
}

static const int kXpadLen[8] = {4, 6, 8, 12, 16, 24, 32, 48};

struct Decoder {
    struct Sub { int app; int len; };
    std::vector<Sub> last;               // sub-fields of the previous X-PAD (a field without contents indicator continues them)
    std::vector<uint8_t> dl; bool dl_on = false;
    std::vector<uint8_t> dg; bool dg_on = false; int dg_need = -1; int dg_app = 0;
    std::vector<uint8_t> x;              // X-PAD in logical order

    void reset() { last.clear(); dl.clear(); dg.clear(); dl_on = dg_on = false; dg_need = -1; }

    static int dl_length(const std::vector<uint8_t>& d)
    {
        if (d.size() < 2) return -1;
        if (!(d[0] & 0x10)) return 2 + (d[0] & 0x0F) + 1 + 2;            // message segment: prefix + characters + CRC
        switch (d[0] & 0x0F) {
            case 1: return 4;                                            // clear display
            case 2: return 2 + (d[1] & 0x0F) + 1 + 2;                    // DL Plus command (TS 102 980)
            default: return 4;
        }
    }

    // on_dl(const uint8_t*, int) for every CRC-correct dynamic-label data group; on_dg(int xpad_app_type, const uint8_t*, int)
    // for every CRC-correct MSC data group of X-PAD application `xpad_app_type` (the type of its start sub-field)
    template <class DlFn, class DgFn>
    void subfield(int app, const uint8_t* p, int n, DlFn&& on_dl, DgFn&& on_dg)
    {
        if (app == 1) {                                                   // data group length indicator
            dg_need = -1; dg_on = false;
            if (n >= 4 && crc_ok(p, 4)) dg_need = ((p[0] & 0x3F) << 8) | p[1];
        } else if (app == 2 || app == 3) {                                // dynamic label: start / continuation
            if (app == 2) { dl.clear(); dl_on = true; }
            if (!dl_on) return;
            dl.insert(dl.end(), p, p + n);
            const int need = dl_length(dl);
            if (need > 0 && (int)dl.size() >= need) {
                if (crc_ok(dl.data(), (size_t)need)) on_dl(dl.data(), need);
                dl_on = false; dl.clear();
            }
        } else if (app >= 4 && app < 31) {                                // user applications: even = start, odd = continuation
            if (!(app & 1)) {
                dg.clear(); dg_on = dg_need > 0; dg_app = app;
            } else if (!dg_on || app != dg_app + 1) return;
            if (!dg_on) return;
            dg.insert(dg.end(), p, p + n);
            if ((int)dg.size() >= dg_need) {
                if (crc_ok(dg.data(), (size_t)dg_need)) on_dg(dg_app, dg.data(), dg_need);
                dg_on = false; dg.clear(); dg_need = -1;
            }
        }
    }

    // pad: the PAD field of one audio frame in transmission order (X-PAD bytes reversed, then F-PAD byte L-1 and byte L)
    // returns the number of X-PAD bytes in use (contents indicators + sub-fields)
    template <class DlFn, class DgFn>
    int feed(const uint8_t* padb, int n, DlFn&& on_dl, DgFn&& on_dg)
    {
        if (n < 2) return 0;
        const uint8_t f0 = padb[n - 2], f1 = padb[n - 1];
        if ((f0 >> 6) != 0) return 0;                                     // F-PAD type 00 only
        const int ind = (f0 >> 4) & 3;
        const bool ci = (f1 >> 1) & 1;
        int nx = n - 2;
        if (ind == 0 || nx <= 0) return 0;
        if (ind == 1) { if (nx < 4) return 0; padb += nx - 4; nx = 4; }   // short X-PAD: the four bytes next to the F-PAD
        else if (ind != 2) return 0;
        x.resize((size_t)nx);
        for (int i = 0; i < nx; i++) x[(size_t)i] = padb[nx - 1 - i];
        int pos = 0;
        std::vector<Sub> subs;
        if (ind == 1) {
            if (ci) { subs.push_back({x[0] & 0x1F, 3}); pos = 1; }
            else if (!last.empty()) subs.push_back({last.back().app | 1, 4});      // continuation of the same application
        } else if (ci) {
            for (int k = 0; k < 4 && pos < nx; k++) {
                const uint8_t c = x[(size_t)pos++];
                const int app = c & 0x1F;
                if (app == 0) break;                                      // end marker
                if (app == 31) { pos++; continue; }                       // extended application type: not used
                subs.push_back({app, kXpadLen[c >> 5]});
            }
        } else {
            // no contents indicator: same layout as before, every sub-field continues its application
            for (auto& s : last) subs.push_back({s.app == 1 ? 1 : (s.app | 1), s.len});
            if (!subs.empty() && ind == 2) { int tot = 0; for (auto& s : subs) tot += s.len; if (tot > nx) subs.clear(); }
        }
        for (auto& s : subs) {
            if (s.app == 0) continue;
            if (pos + s.len > nx) break;
            subfield(s.app, x.data() + pos, s.len, on_dl, on_dg);
            pos += s.len;
        }
        if (!subs.empty()) last = subs;
        return pos;
    }
};

}  // namespace pad

// Packet mode (EN 300 401 sec. 5.3.2 / 5.3.3): a sub-channel is a sequence of packets of 24 / 48 / 72 / 96 bytes, each with a
// 3-byte header (length, continuity index, first / last flags, 10-bit address, command flag, useful data length), a data field
// and a CRC-16; the useful data of the packets of one address between a `first` and a `last` flag form one MSC data group.
// The reference core hands the complete group (header .. CRC) to dabsdrDataGroupCBFunc_t with id = DABSDR_ID_DATA.
namespace packet {

struct Assembler {
    int addr = -1;
    std::vector<uint8_t> dg;
    bool on = false;
    int ci = -1;

    void reset(int address) { addr = address; dg.clear(); on = false; ci = -1; }

    // one logical frame of the sub-channel; on_dg(const uint8_t*, int) per complete data group
    template <class Fn>
    void frame(const uint8_t* f, int len, Fn&& on_dg)
    {
        int pos = 0;
        while (pos + 24 <= len) {
            const uint8_t* p = f + pos;
            const int plen = 24 * ((p[0] >> 6) + 1);
            if (pos + plen > len || !pad::crc_ok(p, (size_t)plen)) { pos += 24; continue; }     // damaged: resynchronise on the 24-byte grid
            pos += plen;
            const int a = ((p[0] & 3) << 8) | p[1];
            if (a != addr) continue;                                                           // other component or padding (address 0)
            const int cix = (p[0] >> 4) & 3, first = (p[0] >> 3) & 1, last = (p[0] >> 2) & 1, useful = p[2] & 0x7F;
            if (useful > plen - 5) { on = false; continue; }
            if (first) { dg.clear(); on = true; }
            else if (!on || cix != ((ci + 1) & 3)) { on = false; ci = cix; continue; }           // a packet of this group was lost
            ci = cix;
            dg.insert(dg.end(), p + 3, p + 3 + useful);
            if (last) {
                const bool has_crc = !dg.empty() && (dg[0] & 0x40);
                if (dg.size() >= 2 && (!has_crc || pad::crc_ok(dg.data(), dg.size()))) on_dg(dg.data(), (int)dg.size());
                on = false;
            }
        }
    }
};

}  // namespace packet
