// fig.h -- Fast Information Group database (ETSI EN 300 401 clause 6, 8; TS 101 756 tables).
//
// Host-side, byte-level work that turns CRC-checked FIBs coming off the GPU into the structures the dabsdr API hands to
// RadioControl (reference: lib/linux_x86_64/dabsdr.h:169-318; consumer src/radiocontrol.cpp:1381-1617).  Inside the
// reference this lives in the closed core ("FIG0/1/2 extension %d not supported yet" strings).
// Implemented: FIG 0/0 0/1 0/2 0/3 0/5 0/8 0/9 0/10 0/13 0/17 0/18 0/19, FIG 1/0 1/1 1/4 1/5.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>
#include "tables.h"

namespace fig {

struct Label { char str[17] = {0}; uint16_t char_field = 0; uint8_t charset = 0; bool valid = false; };

struct SubCh { bool valid = false; b200dab_subch_t p; };

struct UserApp { uint16_t type = 0; uint8_t data_len = 0; uint8_t data[23] = {0}; };

struct Component {
    uint8_t tmid = 0;
    uint8_t ascty_dscty = 0;
    int subch_id = -1;          // stream modes
    uint16_t scid = 0;          // packet mode
    int16_t packet_addr = -1;
    uint8_t dg_flag = 0;
    uint8_t ps = 0, ca = 0;
    int scids = -1;             // from FIG 0/8; primary = 0
    uint8_t lang = 0;
    Label label;
    std::vector<UserApp> user_apps;
};

struct Service {
    uint32_t sid = 0;
    uint8_t caid = 0, local = 0;
    std::vector<Component> comps;
    Label label;
    uint8_t pty_s = 255, pty_d = 255;
    uint16_t asu = 0;
    std::vector<uint8_t> clusters;
};

struct Database {
    bool have_eid = false;
    uint16_t eid = 0;
    uint8_t ecc = 0, int_table = 0, alarm = 0;
    int8_t lto = 0;
    bool have_ecc = false;
    Label label;
    int cif_count = -1;                 // of the last FIG 0/0 seen
    SubCh subch[64];
    std::vector<Service> services;      // in order of first appearance
    uint32_t date_hours_minutes = 0;    // FIG 0/10: MJD << 14 | LSI << 13 | hours << 6 | minutes
    uint16_t sec_msec = 0;
    uint64_t version = 0;               // bumps when something changed
    // changes the caller is told about without asking (dabsdr.h: NID_USER_APP_UPDATE, NID_PTY, NID_ANNOUNCEMENT_SUPPORT,
    // NID_ANNOUNCEMENT_SWITCHING); filled by parse_fib in FIG order, drained by the owner
    struct Event {
        enum Kind { UA_UPDATE, PTY, ASU, ASW, NEW_EID, RECONFIGURATION } kind;
        uint32_t sid = 0;
        uint8_t scids = 0;
        int n_asw = 0;
        struct { uint8_t cluster, subch; uint16_t flags; } asw[8];
    };
    std::vector<Event> events;
    // Multiplex reconfiguration (EN 300 401 sec. 6.4, 6.5): while FIG 0/0 carries change flags, the MCI extensions sent with C/N = 1
    // describe the NEXT configuration and are kept apart; one frame before the CIF counter reaches the announced occurrence the next
    // configuration becomes the current one and the owner is told (DABSDR_NID_RECONFIGURATION, dabsdr.h:112).
    SubCh next_subch[64];
    std::vector<Service> next_services;
    int change_flags = 0, occurrence = -1;
    bool reconfigured = false;          // this announcement has been acted upon

    void reset() { *this = Database(); }

    void apply_next_configuration()
    {
        // Sub-channel organisation: the announced one takes over at once, so the decoders of the running services keep their
        // (possibly moved) sub-channels without a gap.  Service organisation: dropped and rebuilt from the FIGs of the new
        // configuration, which arrive with C/N = 0 from the change on -- like the reference, whose first service list after a
        // reconfiguration is empty and whose second one is complete (golden c1_reconf_u8_30db); RadioControl asks again until it is
        // (src/radiocontrol.cpp:1427-1436).
        bool any = false;
        for (int i = 0; i < 64; i++) any = any || next_subch[i].valid;
        if (any) for (int i = 0; i < 64; i++) subch[i] = next_subch[i];
        services.clear();
        next_services.clear();
        for (auto& n : next_subch) n = SubCh();
        version++;
    }

    Service* find(uint32_t sid) { for (auto& s : services) if (s.sid == sid) return &s; return nullptr; }
    Service& get(uint32_t sid)
    {
        if (Service* s = find(sid)) return *s;
        services.emplace_back();
        services.back().sid = sid;
        version++;
        return services.back();
    }
    Component* find_comp_by_subch(int subch_id)
    {
        for (auto& s : services) for (auto& c : s.comps) if (c.tmid != 3 && c.subch_id == subch_id) return &c;
        return nullptr;
    }

    void set_label(Label& l, const uint8_t* p, uint8_t charset)
    {
        if (l.valid && !memcmp(l.str, p, 16)) return;
        memcpy(l.str, p, 16); l.str[16] = 0;
        l.char_field = (uint16_t)((p[16] << 8) | p[17]);
        l.charset = charset;
        l.valid = true;
        version++;
    }

    void parse_fib(const uint8_t* p)
    {
        int i = 0;
        while (i < 30) {
            const uint8_t h = p[i];
            if (h == 0xFF) break;
            const int type = h >> 5, len = h & 31;
            if (len == 0 || i + 1 + len > 30) break;
            const uint8_t* d = p + i + 1;
            if (type == 0) fig0(d, len);
            else if (type == 1) fig1(d, len);
            i += 1 + len;
        }
    }

    void fig0(const uint8_t* d, int len)
    {
        const int ext = d[0] & 31, pd = (d[0] >> 5) & 1, oe = (d[0] >> 6) & 1, cn = d[0] >> 7;
        if (oe && ext != 0) return;                        // other-ensemble information is not kept
        if (cn && (ext == 1 || ext == 2 || ext == 3 || ext == 8)) {
            // MCI of the next configuration: parsed with the same code into the tables kept for it
            std::swap(services, next_services);
            for (int i = 0; i < 64; i++) std::swap(subch[i], next_subch[i]);
            uint8_t tmp[32];
            memcpy(tmp, d, (size_t)len);
            tmp[0] &= 0x7F;
            const uint64_t v = version;
            fig0(tmp, len);
            version = v;
            std::swap(services, next_services);
            for (int i = 0; i < 64; i++) std::swap(subch[i], next_subch[i]);
            return;
        }
        int k = 1;
        switch (ext) {
        case 0:
            if (len >= 5) {
                const uint16_t e = (uint16_t)((d[1] << 8) | d[2]);
                if (have_eid && e != eid) { Event ev; ev.kind = Event::NEW_EID; events.push_back(ev); }      // another ensemble on this frequency
                if (!have_eid || e != eid) { eid = e; have_eid = true; version++; }
                alarm = (d[3] >> 5) & 1;
                cif_count = (d[3] & 0x1F) * 250 + d[4];
                const int flags = d[3] >> 6;
                if (flags && len >= 6) {
                    if (!change_flags || occurrence != d[5]) reconfigured = false;
                    change_flags = flags; occurrence = d[5];
                    // the change takes effect with the frame that follows this one (a frame carries four CIF counts)
                    const int ahead = ((occurrence - d[4]) % 250 + 250) % 250;
                    if (!reconfigured && ahead <= 4) {
                        reconfigured = true;
                        apply_next_configuration();
                        Event ev; ev.kind = Event::RECONFIGURATION; events.push_back(ev);
                    }
                } else if (!flags) { change_flags = 0; occurrence = -1; reconfigured = false; }
            }
            break;
        case 1:
            while (k + 3 <= len) {
                const int id = d[k] >> 2, start = ((d[k] & 3) << 8) | d[k + 1];
                b200dab_subch_t sc;
                memset(&sc, 0, sizeof sc);
                int ok = -1;
                if (d[k + 2] & 0x80) {
                    if (k + 4 > len) break;
                    const int opt = (d[k + 2] >> 4) & 7, lvl = (d[k + 2] >> 2) & 3, size = ((d[k + 2] & 3) << 8) | d[k + 3];
                    int bitrate = 0;
                    if (opt == 0) { static const int dv[4] = {12, 8, 6, 4}; bitrate = size / dv[lvl] * 8; }
                    else if (opt == 1) { static const int dv[4] = {27, 21, 18, 15}; bitrate = size / dv[lvl] * 32; }
                    if (bitrate > 0) ok = dabtab::profile_eep(bitrate, lvl + 1, opt, &sc);
                    if (ok == 0 && sc.size_cu != size) ok = -1;
                    k += 4;
                } else {
                    ok = dabtab::profile_uep(d[k + 2] & 0x3F, &sc);
                    k += 3;
                }
                if (ok == 0) {
                    sc.subch_id = id; sc.start_cu = start;
                    if (!subch[id].valid || subch[id].p.start_cu != start || subch[id].p.size_cu != sc.size_cu || subch[id].p.bitrate != sc.bitrate) {
                        sc.kind = subch[id].p.kind;        // FIG 0/2 may have told us already
                        subch[id].p = sc; subch[id].valid = true; version++;
                    }
                }
            }
            break;
        case 2:
            while (k < len) {
                const int sidlen = pd ? 4 : 2;
                if (k + sidlen + 1 > len) break;
                uint32_t sid = 0;
                for (int b = 0; b < sidlen; b++) sid = (sid << 8) | d[k + b];
                k += sidlen;
                const int ncomp = d[k] & 0x0F, caid = (d[k] >> 4) & 7, local = d[k] >> 7;
                k++;
                Service& s = get(sid);
                s.caid = (uint8_t)caid; s.local = (uint8_t)local;
                for (int c = 0; c < ncomp && k + 2 <= len; c++, k += 2) {
                    Component nc;
                    nc.tmid = d[k] >> 6;
                    if (nc.tmid == 3) { nc.scid = (uint16_t)(((d[k] & 0x3F) << 6) | (d[k + 1] >> 2)); }
                    else { nc.ascty_dscty = d[k] & 0x3F; nc.subch_id = d[k + 1] >> 2; }
                    nc.ps = (d[k + 1] >> 1) & 1; nc.ca = d[k + 1] & 1;
                    bool found = false;
                    for (auto& e : s.comps)
                        if (e.tmid == nc.tmid && ((nc.tmid == 3 && e.scid == nc.scid) || (nc.tmid != 3 && e.subch_id == nc.subch_id))) {
                            found = true;
                            if (nc.tmid == 3) nc.ascty_dscty = e.ascty_dscty;          // packet mode: the DSCTy comes from FIG 0/3
                            if (e.ascty_dscty != nc.ascty_dscty || e.ps != nc.ps) { e.ascty_dscty = nc.ascty_dscty; e.ps = nc.ps; e.ca = nc.ca; version++; }
                        }
                    if (!found) { if (nc.ps) nc.scids = 0; s.comps.push_back(nc); version++; }
                    if (nc.tmid == 0 && nc.subch_id >= 0 && subch[nc.subch_id].p.kind != (nc.ascty_dscty == 63 ? 1 : 0)) {
                        subch[nc.subch_id].p.kind = nc.ascty_dscty == 63 ? 1 : 0;
                        if (subch[nc.subch_id].valid) version++;
                    }
                }
            }
            break;
        case 3:
            while (k + 5 <= len) {
                const uint16_t scid = (uint16_t)((d[k] << 4) | (d[k + 1] >> 4));
                const int scca = d[k + 1] & 1;
                const uint8_t dg = d[k + 2] >> 7, dscty = d[k + 2] & 0x3F;
                const int sub = d[k + 3] >> 2, addr = ((d[k + 3] & 3) << 8) | d[k + 4];
                k += 5 + (scca ? 2 : 0);
                for (auto& s : services) for (auto& c : s.comps)
                    if (c.tmid == 3 && c.scid == scid && (c.subch_id != sub || c.packet_addr != addr)) {
                        c.subch_id = sub; c.packet_addr = (int16_t)addr; c.dg_flag = dg; c.ascty_dscty = dscty; version++;
                    }
            }
            break;
        case 5:
            while (k < len) {
                if (d[k] & 0x80) { k += 3; continue; }          // long form (SCId): not used by audio services
                if (k + 2 > len) break;
                const int sub = d[k] & 0x3F;
                if (Component* c = find_comp_by_subch(sub)) if (c->lang != d[k + 1]) { c->lang = d[k + 1]; version++; }
                k += 2;
            }
            break;
        case 8:
            while (k < len) {
                const int sidlen = pd ? 4 : 2;
                if (k + sidlen + 2 > len) break;
                uint32_t sid = 0;
                for (int b = 0; b < sidlen; b++) sid = (sid << 8) | d[k + b];
                k += sidlen;
                const int extf = d[k] >> 7, scids = d[k] & 0x0F;
                k++;
                if (d[k] & 0x80) {                               // long form: SCId
                    if (k + 2 > len) break;
                    const uint16_t scid = (uint16_t)(((d[k] & 0x0F) << 8) | d[k + 1]);
                    if (Service* s = find(sid)) for (auto& c : s->comps) if (c.tmid == 3 && c.scid == scid && c.scids != scids) { c.scids = scids; version++; }
                    k += 2;
                } else {
                    const int sub = d[k] & 0x3F;
                    if (Service* s = find(sid)) for (auto& c : s->comps) if (c.tmid != 3 && c.subch_id == sub && c.scids != scids) { c.scids = scids; version++; }
                    k += 1;
                }
                if (extf) k++;
            }
            break;
        case 9:
            if (len >= 4) {
                const int v = d[1] & 0x1F;
                const int8_t l = (int8_t)((d[1] & 0x20) ? -v : v);
                if (!have_ecc || ecc != d[2] || lto != l) version++;
                lto = l; ecc = d[2]; int_table = d[3]; have_ecc = true;
            }
            break;
        case 10:
            if (len >= 5) {
                const uint32_t mjd = ((uint32_t)(d[1] & 0x7F) << 10) | ((uint32_t)d[2] << 2) | (d[3] >> 6);
                const int utc = (d[3] >> 3) & 1;
                const uint32_t hours = ((d[3] & 7) << 2) | (d[4] >> 6), minutes = d[4] & 0x3F;
                const uint32_t lsi = (d[3] >> 5) & 1;
                date_hours_minutes = (mjd << 14) | (lsi << 13) | (hours << 6) | minutes;   // packing RadioControl unpacks (src/dabtables.cpp:124-157)
                if (utc && len >= 7) sec_msec = (uint16_t)(((d[5] >> 2) << 10) | (((d[5] & 3) << 8) | d[6]));
                else sec_msec = 0;
            }
            break;
        case 13:
            while (k < len) {
                const int sidlen = pd ? 4 : 2;
                if (k + sidlen + 1 > len) break;
                uint32_t sid = 0;
                for (int b = 0; b < sidlen; b++) sid = (sid << 8) | d[k + b];
                k += sidlen;
                const int scids = d[k] >> 4, nua = d[k] & 0x0F;
                k++;
                Service* s = find(sid);
                for (int u = 0; u < nua && k + 2 <= len; u++) {
                    UserApp ua;
                    ua.type = (uint16_t)((d[k] << 3) | (d[k + 1] >> 5));
                    const int dl = d[k + 1] & 0x1F;
                    k += 2;
                    if (k + dl > len) { k = len; break; }
                    ua.data_len = (uint8_t)(dl > 23 ? 23 : dl);
                    memcpy(ua.data, d + k, ua.data_len);
                    k += dl;
                    if (s) for (auto& c : s->comps) if (c.scids == scids) {
                        bool have = false;
                        for (auto& e : c.user_apps) if (e.type == ua.type) have = true;
                        if (!have) {
                            c.user_apps.push_back(ua); version++;
                            bool queued = false;
                            for (auto& e : events) if (e.kind == Event::UA_UPDATE && e.sid == sid && e.scids == scids) queued = true;
                            if (!queued) { Event e; e.kind = Event::UA_UPDATE; e.sid = sid; e.scids = (uint8_t)scids; events.push_back(e); }
                        }
                    }
                }
            }
            break;
        case 17:
            while (k + 4 <= len) {
                const uint32_t sid = (uint32_t)((d[k] << 8) | d[k + 1]);
                const int sd = d[k + 2] >> 7, lf = (d[k + 2] >> 5) & 1, cf = (d[k + 2] >> 4) & 1;
                int q = k + 3;
                if (lf) q++;
                if (q >= len) break;
                const uint8_t pty = d[q] & 0x1F;
                q++;
                if (cf) q++;
                if (Service* s = find(sid)) {
                    uint8_t& slot = sd ? s->pty_d : s->pty_s;
                    if (slot != pty) { slot = pty; version++; Event e; e.kind = Event::PTY; e.sid = sid; events.push_back(e); }
                }
                k = q;
            }
            break;
        case 18:
            while (k + 5 <= len) {
                const uint32_t sid = (uint32_t)((d[k] << 8) | d[k + 1]);
                const uint16_t asu = (uint16_t)((d[k + 2] << 8) | d[k + 3]);
                const int nc = d[k + 4] & 0x1F;
                if (Service* s = find(sid)) {
                    std::vector<uint8_t> cl(d + k + 5, d + k + 5 + (k + 5 + nc <= len ? nc : 0));
                    if (s->asu != asu || s->clusters != cl) {
                        s->asu = asu; s->clusters = cl; version++;
                        Event e; e.kind = Event::ASU; e.sid = sid; events.push_back(e);
                    }
                }
                k += 5 + nc;
            }
            break;
        case 19: {
            // announcement switching: one notification per received FIG 0/19, as the reference core does
            Event e; e.kind = Event::ASW;
            while (k + 4 <= len) {
                const int region = (d[k + 3] >> 6) & 1;
                if (e.n_asw < 7) { e.asw[e.n_asw].cluster = d[k]; e.asw[e.n_asw].flags = (uint16_t)((d[k + 1] << 8) | d[k + 2]); e.asw[e.n_asw].subch = d[k + 3] & 0x3F; e.n_asw++; }
                k += 4 + (region ? 1 : 0);
            }
            if (e.n_asw) events.push_back(e);
            break; }
        default:
            break;
        }
    }

    void fig1(const uint8_t* d, int len)
    {
        const int charset = d[0] >> 4, ext = d[0] & 7;
        switch (ext) {
        case 0:
            if (len >= 21) set_label(label, d + 3, (uint8_t)charset);
            break;
        case 1:
            if (len >= 21) { Service& s = get((uint32_t)((d[1] << 8) | d[2])); set_label(s.label, d + 3, (uint8_t)charset); }
            break;
        case 4:
            if (len >= 22) {
                const int pd = d[1] >> 7, scids = d[1] & 0x0F;
                if (!pd) if (Service* s = find((uint32_t)((d[2] << 8) | d[3]))) for (auto& c : s->comps) if (c.scids == scids) set_label(c.label, d + 4, (uint8_t)charset);
            }
            break;
        case 5:
            if (len >= 23) { Service& s = get(((uint32_t)d[1] << 24) | ((uint32_t)d[2] << 16) | ((uint32_t)d[3] << 8) | d[4]); set_label(s.label, d + 5, (uint8_t)charset); }
            break;
        default:
            break;
        }
    }
};

}  // namespace fig
