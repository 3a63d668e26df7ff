// dabsdr_shim.cpp -- drop-in for the reference's closed receiver core: exports the 24 entry points of the dabsdr C API
// (include/dabsdr_abi.h; reference interface lib/linux_x86_64/dabsdr.h:397-429, SONAME libdabsdr.so.4) on top of the
// CUDA pipeline in libb200dab.so, so that AbracaDABra's RadioControl (src/radiocontrol.cpp) links against it unchanged.
//
// Contract reproduced (SURVEY.md sec. 8 b):
//   * one worker thread per handle, named "dabsdr"; every callback and both input functions run on it;
//   * requests are non-blocking, queued, and answered by a notification with the matching nid;
//   * every pData / pAuData pointer is valid only during the callback; list notifications carry an accessor;
//   * Tune(0) -> NID_TUNE(0); Tune(f) -> NID_TUNE(f), NID_RESET(INIT); SetTII -> immediate empty NID_TII;
//     the first NID_PERIODIC after SetPeriodicNotify has pData == NULL (guard at src/radiocontrol.cpp:2405).
//
// Streaming on a batch engine: the worker pulls float IQ through the registered input function
// (src/input/inputdevice.cpp:70-108 getSamples) into a sliding window of ADV + OVL transmission frames and runs the
// whole GPU chain on the window; the overlap (time de-interleaver depth + one DAB+ superframe) is simply decoded
// again, and results are de-duplicated by absolute CIF number.  The dummy input function is never needed: the GPU
// demodulates every symbol, so nothing is skipped.
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <pthread.h>
#include <vector>

#include "b200dab.h"
#include "dabsdr_abi.h"
#include "fig.h"
#include "mp2.h"
#include "pad.h"

namespace {

constexpr int ADV_FRAMES_MIN = 4;              // new frames per GPU pass of a live stream (384 ms of signal)
constexpr int ADV_FRAMES_MAX = 64;             // ... of a source that delivers faster than real time (a file): the batch grows up to this
                                               // (B200DAB_DROPIN_MAX_BATCH in the environment lowers the limit; 4 = always behave as live)
constexpr int OVL_FRAMES = 6;                  // re-decoded overlap: 16 CIFs of interleaver memory + 5 CIFs of superframe
constexpr int64_t WIN_SAMPLES_MAX = (int64_t)(ADV_FRAMES_MAX + OVL_FRAMES) * DAB_T_FRAME + 4096;

enum ReqType { R_TUNE, R_ENSEMBLE, R_SERVICE_LIST, R_SERVICE_COMPS, R_USER_APPS, R_ANN_SUPPORT, R_SELECT, R_STOP, R_XPAD, R_PERIODIC, R_TII, R_SPECTRUM, R_EXIT };
struct Request { ReqType type; uint32_t a = 0; uint32_t b = 0; int c = 0; };

struct Decoder {               // one of the (up to two) audio decoder instances of the API
    bool active = false;
    uint32_t sid = 0;
    uint8_t scids = 0;
    int subch_id = -1;
    bool dabplus = false;
    int64_t last_row = -1;     // absolute logical-frame number delivered last (MP2)
    int64_t last_sf_row = -1;  // absolute row of the last delivered superframe (DAB+)
    int64_t last_stat_row = -1;// absolute row of the last superframe that went into the RS statistics
    uint8_t mp2_drc = 0;       // DRC value for the NEXT MP2 frame: F-PAD of the frame before it (dabsdr.h:50 mp2DRC)
    pad::Decoder xpad;         // X-PAD state of this audio service
    bool xpad_app[32] = {};    // X-PAD application types started with dabsdrRequest_XPadAppStart
};

constexpr int MAX_DATA_DEC = 4;
struct DataDecoder {           // packet-mode service component started on DABSDR_ID_DATA
    bool active = false;
    uint32_t sid = 0;
    uint8_t scids = 0;
    int subch_id = -1;
    uint16_t scid = 0;
    uint16_t ua_type = 0xFFFF;
    int64_t last_row = -1;
    packet::Assembler as;
};

}  // namespace

struct dabsdr_s {
    dabsdrInputFunc_t in = nullptr, dummy = nullptr;
    dabsdrAudioCBFunc_t audio_cb = nullptr; void* audio_ctx = nullptr;
    dabsdrDynamicLabelCBFunc_t dl_cb = nullptr; void* dl_ctx = nullptr;
    dabsdrDataGroupCBFunc_t dg_cb = nullptr; void* dg_ctx = nullptr;
    dabsdrSpectrumCBFunc_t spec_cb = nullptr; void* spec_ctx = nullptr;
    dabsdrNotificationCBFunc_t ntf_cb = nullptr; void* ntf_ctx = nullptr;

    pthread_t thr{};
    bool started = false;
    std::mutex mx;
    std::condition_variable cv;
    std::deque<Request> q;

    b200dab_ctx* gpu = nullptr;
    uint32_t freq = 0;
    int sync_level = 0;
    fig::Database db;
    Decoder dec[2];
    DataDecoder data[MAX_DATA_DEC];
    struct Win { float* p = nullptr; float* data() { return p; } } win;   // interleaved I,Q: pinned host memory (b200dab_host_alloc)
    float* win_data = nullptr;
    std::vector<float> win_fallback;
    int adv_frames = ADV_FRAMES_MIN;   // new frames per pass: doubles while the input function delivers much faster than real time
    int adv_max = ADV_FRAMES_MAX;
    int fast_pulls = 0;
    bool at_end = false;               // the source has run out (zero-filled chunks): batches stay small
    int64_t win_samples() const { return (int64_t)(adv_frames + OVL_FRAMES) * DAB_T_FRAME + 4096; }
    int64_t adv_samples() const { return (int64_t)adv_frames * DAB_T_FRAME; }
    double t_pull0 = 0;                // wall clock when the pull of the current batch began
    int64_t pulled0 = 0;               // samples in the window at that time
    int64_t have = 0;                  // complex samples in win
    int64_t abs_pos = 0;               // absolute index of win[0]
    int64_t last_frame_abs = -1;       // absolute t_f of the newest frame already accounted for
    int64_t frame_no = -1;             // absolute number of that frame
    bool cif_known = false;
    // statistics for NID_PERIODIC
    int period = -1;
    int64_t next_periodic = 0;
    uint32_t fib_err = 0, crc_ok = 0, crc_err = 0, audio_bytes = 0, pad_bytes = 0, rs_unc = 0, rs_bit = 0, rs_bytes = 0;
    float last_cfo = 0.f, last_snr = 0.f;
    // TII
    bool tii_on = false; int tii_mode = 1;
    std::vector<float> tii_fold, tii_norm; std::vector<int32_t> tii_sign;
    std::vector<float> tii_spectrum = std::vector<float>(TII_SPECTRUM_BINS, 0.f);
    bool spectrum_on = false;
    // accessor snapshots (valid during a callback)
    std::vector<dabsdrServiceListItem_t> snap_services;
    std::vector<dabsdrServiceCompListItem_t> snap_comps;
    std::vector<dabsdrUserAppListItem_t> snap_uas;
    std::vector<uint8_t> pad_tmp;      // PAD of an MP2 frame, re-joined across the ScF-CRC
};

namespace {

double wall_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void notify(dabsdr_s* h, dabsdrNotificationId_t nid, dabsdrNotificationStatus_t st, const void* data, uint16_t len)
{
    if (!h->ntf_cb) return;
    dabsdrNotificationCBData_t n;
    n.nid = nid; n.status = st; n.len = len; n.pData = data;
    h->ntf_cb(&n, h->ntf_ctx);
}

void fill_label(dabsdrLabel_t& o, const fig::Label& l)
{
    memset(&o, 0, sizeof o);
    if (l.valid) { memcpy(o.str, l.str, 16); o.str[16] = 0; o.charField = l.char_field; o.charset = l.charset; }
}

int get_service_item(dabsdrHandle_t h, uint8_t idx, dabsdrServiceListItem_t* out)
{
    if (!h || !out || idx >= h->snap_services.size()) return -1;
    *out = h->snap_services[idx];
    return 0;
}
int get_comp_item(dabsdrHandle_t h, uint8_t idx, dabsdrServiceCompListItem_t* out)
{
    if (!h || !out || idx >= h->snap_comps.size()) return -1;
    *out = h->snap_comps[idx];
    return 0;
}
int get_ua_item(dabsdrHandle_t h, uint8_t idx, dabsdrUserAppListItem_t* out)
{
    if (!h || !out || idx >= h->snap_uas.size()) return -1;
    *out = h->snap_uas[idx];
    return 0;
}
int get_tii_spectrum(dabsdrHandle_t h, float* buf)
{
    if (!h || !buf) return -1;
    memcpy(buf, h->tii_spectrum.data(), TII_SPECTRUM_BINS * sizeof(float));     // 384 floats: src/radiocontrol.h:280
    return 0;
}

void send_sync(dabsdr_s* h, int level)
{
    if (level == h->sync_level) return;
    h->sync_level = level;
    dabsdrNtfSyncStatus_t s;
    s.syncLevel = (dabsdrSyncLevel_t)level;
    s.snr10 = (int16_t)lrintf(h->last_snr * 10.f);
    notify(h, DABSDR_NID_SYNC_STATUS, DABSDR_NSTAT_SUCCESS, &s, sizeof s);
}

void send_tii(dabsdr_s* h, const b200dab_tii_t* t)
{
    dabsdrNtfTii_t n;
    memset(&n, 0, sizeof n);
    n.getSpectrumTii = get_tii_spectrum;
    if (t) {
        n.numIds = (uint8_t)t->n_ids;
        for (int i = 0; i < t->n_ids && i < 24; i++) { n.id[i].main = t->main_id[i]; n.id[i].sub = t->sub_id[i]; n.id[i].level = t->level[i]; }
    }
    notify(h, DABSDR_NID_TII, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
}

// unsolicited notifications raised by the FIGs just parsed (SURVEY.md sec. 8 b; consumers src/radiocontrol.cpp:2323, :2441-2478)
void send_fig_events(dabsdr_s* h)
{
    bool new_eid = false;
    for (const fig::Database::Event& e : h->db.events) {
        switch (e.kind) {
        case fig::Database::Event::NEW_EID: new_eid = true; break;
        case fig::Database::Event::RECONFIGURATION:
            // the caller clears its service list and asks for the new one (src/radiocontrol.cpp:178-190)
            notify(h, DABSDR_NID_RECONFIGURATION, DABSDR_NSTAT_SUCCESS, nullptr, 0);
            break;
        case fig::Database::Event::UA_UPDATE: {
            dabsdrNtfUserAppUpdate_t n; memset(&n, 0, sizeof n);
            n.SId = e.sid; n.SCIdS = e.scids;
            notify(h, DABSDR_NID_USER_APP_UPDATE, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
            break; }
        case fig::Database::Event::PTY: {
            dabsdrNtfPTy_t n; memset(&n, 0, sizeof n);
            n.SId = e.sid;
            if (fig::Service* s = h->db.find(e.sid)) { n.s = s->pty_s; n.d = s->pty_d; }
            notify(h, DABSDR_NID_PTY, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
            break; }
        case fig::Database::Event::ASU: {
            dabsdrNtfAnnouncementSupport_t n; memset(&n, 0, sizeof n);
            n.SId = e.sid;
            if (fig::Service* s = h->db.find(e.sid)) {
                n.ASu = (uint16_t)(s->asu & 0xFFFE);                 // the alarm flag (b0) is implied and reported as 0 by the reference
                n.numClusterIds = (uint8_t)std::min<size_t>(7, s->clusters.size()); memcpy(n.clusterIds, s->clusters.data(), n.numClusterIds);
            }
            notify(h, DABSDR_NID_ANNOUNCEMENT_SUPPORT, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
            break; }
        case fig::Database::Event::ASW: {
            dabsdrNtfAnnouncementSwitching_t n; memset(&n, 0, sizeof n);         // the list ends with clusterId 0 (:1822)
            for (int i = 0; i < e.n_asw && i < 7; i++) { n.asw[i].clusterId = e.asw[i].cluster; n.asw[i].subChId = e.asw[i].subch; n.asw[i].ASwFlags = e.asw[i].flags; }
            notify(h, DABSDR_NID_ANNOUNCEMENT_SWITCHING, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
            break; }
        }
    }
    h->db.events.clear();
    if (new_eid) {
        // the ensemble identifier changed while tuned: everything learnt so far belongs to the old multiplex.  The reference raises
        // NID_RESET(NEW_EID) and keeps its decoders running; RadioControl answers by tuning again (src/radiocontrol.cpp:122-126)
        fig::Database fresh;
        for (int i = 0; i < 64; i++) fresh.subch[i] = h->db.subch[i];        // running decoders keep their sub-channels until FIG 0/1 says otherwise
        h->db = fresh;
        const int32_t flag = DABSDR_RESET_NEW_EID;
        notify(h, DABSDR_NID_RESET, DABSDR_NSTAT_SUCCESS, &flag, sizeof flag);
    }
}

void reset_rx(dabsdr_s* h)
{
    h->db.reset();
    h->dec[0] = Decoder(); h->dec[1] = Decoder();
    for (auto& d : h->data) d = DataDecoder();
    h->have = 0; h->abs_pos = 0; h->last_frame_abs = -1; h->frame_no = -1;
    h->fib_err = h->crc_ok = h->crc_err = h->audio_bytes = h->pad_bytes = h->rs_unc = h->rs_bit = h->rs_bytes = 0;
    h->tii_fold.clear(); h->tii_norm.clear(); h->tii_sign.clear();
    std::fill(h->tii_spectrum.begin(), h->tii_spectrum.end(), 0.f);
    h->next_periodic = h->period >= 0 ? ((int64_t)DAB_T_FRAME << h->period) : 0;
}

fig::Component* find_component(dabsdr_s* h, uint32_t sid, uint8_t scids, fig::Service** svc_out)
{
    fig::Service* s = h->db.find(sid);
    if (svc_out) *svc_out = s;
    if (!s) return nullptr;
    for (auto& c : s->comps) {
        const int cs = c.scids >= 0 ? c.scids : (c.ps ? 0 : -1);
        if (cs == scids) return &c;
    }
    return nullptr;
}

void fill_component(dabsdr_s* h, const fig::Component& c, dabsdrServiceCompListItem_t& o)
{
    memset(&o, 0, sizeof o);
    o.SCIdS = (uint8_t)(c.scids >= 0 ? c.scids : (c.ps ? 0 : 255));      // secondary component without FIG 0/8 yet: 255, like the reference
    o.SubChId = (uint8_t)(c.subch_id >= 0 ? c.subch_id : 0);
    o.SubChAddr = -1;
    o.ps = (uint8_t)(c.ps ? 2 : 0);            // raw bit mask, non-zero = primary
    o.lang = c.lang; o.CAflag = c.ca; o.TMId = c.tmid;
    o.numUserApps = (uint8_t)c.user_apps.size();
    fill_label(o.label, c.label);
    if (c.subch_id >= 0 && h->db.subch[c.subch_id].valid) {
        const b200dab_subch_t& p = h->db.subch[c.subch_id].p;
        o.SubChAddr = (int16_t)p.start_cu;
        o.SubChSize = (uint16_t)p.size_cu;
        if (p.uep_index >= 0) { o.protectionLevel = (uint8_t)p.prot_level; o.uepIdx = (uint8_t)p.uep_index; }
        else { o.protectionLevel = (uint8_t)(6 + p.prot_level); o.fecScheme = 0; }      // DabProtectionLevel: src/dabtables.h:35-55
        if (c.tmid == 0) { o.streamAudio.ASCTy = c.ascty_dscty; o.streamAudio.bitRate = (uint16_t)p.bitrate; }
        else if (c.tmid == 1) { o.streamData.DSCTy = c.ascty_dscty; o.streamData.bitRate = (uint16_t)p.bitrate; }
    }
    if (c.tmid == 3) {
        o.packetData.DSCTy = c.ascty_dscty; o.packetData.SCId = c.scid; o.packetData.DGflag = c.dg_flag; o.packetData.packetAddress = c.packet_addr;
    }
}

// ---- request handling (worker thread, between GPU passes) -----------------------------------------------------------
bool handle_request(dabsdr_s* h, const Request& r)
{
    switch (r.type) {
    case R_EXIT:
        return false;
    case R_TUNE: {
        h->freq = r.a;
        reset_rx(h);
        h->sync_level = 0;
        uint32_t f = r.a;
        notify(h, DABSDR_NID_TUNE, DABSDR_NSTAT_SUCCESS, &f, sizeof f);
        if (f != 0) {
            dabsdrNtfResetFlags_t fl = DABSDR_RESET_INIT;
            notify(h, DABSDR_NID_RESET, DABSDR_NSTAT_SUCCESS, &fl, sizeof fl);
            if (h->tii_on) send_tii(h, nullptr);
        }
        break; }
    case R_ENSEMBLE: {
        dabsdrNtfEnsemble_t e;
        memset(&e, 0, sizeof e);
        if (h->db.have_eid) {
            e.frequency = h->freq;
            e.ueid = ((uint32_t)(h->db.have_ecc ? h->db.ecc : 0) << 16) | h->db.eid;
            e.LTO = h->db.lto; e.intTable = h->db.int_table; e.alarm = h->db.alarm;
            fill_label(e.label, h->db.label);
        }
        notify(h, DABSDR_NID_ENSEMBLE_INFO, DABSDR_NSTAT_SUCCESS, &e, sizeof e);
        break; }
    case R_SERVICE_LIST: {
        h->snap_services.clear();
        for (auto& s : h->db.services) {
            if (s.comps.empty()) continue;
            dabsdrServiceListItem_t it;
            memset(&it, 0, sizeof it);
            it.sid = s.sid; it.CAId = s.caid; it.pty.s = s.pty_s; it.pty.d = s.pty_d;
            fill_label(it.label, s.label);
            h->snap_services.push_back(it);
        }
        dabsdrNtfServiceList_t n;
        n.numServices = (uint8_t)h->snap_services.size();
        n.getServiceListItem = get_service_item;
        notify(h, DABSDR_NID_SERVICE_LIST, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
        break; }
    case R_SERVICE_COMPS: {
        h->snap_comps.clear();
        fig::Service* s = h->db.find(r.a);
        if (s) for (auto& c : s->comps) { dabsdrServiceCompListItem_t it; fill_component(h, c, it); h->snap_comps.push_back(it); }
        dabsdrNtfServiceComponentList_t n;
        n.SId = r.a; n.numServiceComponents = (uint8_t)h->snap_comps.size(); n.getServiceComponentListItem = get_comp_item;
        notify(h, DABSDR_NID_SERVICE_COMPONENT_LIST, s ? DABSDR_NSTAT_SUCCESS : DABSDR_NSTAT_SERVICE_NOT_FOUND, &n, sizeof n);
        break; }
    case R_USER_APPS: {
        h->snap_uas.clear();
        fig::Service* s = nullptr;
        fig::Component* c = find_component(h, r.a, (uint8_t)r.b, &s);
        if (c) for (auto& u : c->user_apps) {
            dabsdrUserAppListItem_t it;
            memset(&it, 0, sizeof it);
            it.type = u.type; it.dataLen = u.data_len; memcpy(it.data, u.data, u.data_len);
            h->snap_uas.push_back(it);
        }
        dabsdrNtfUserAppList_t n;
        n.SId = r.a; n.SCIdS = (uint8_t)r.b; n.numUserApps = (uint8_t)h->snap_uas.size(); n.getUserAppListItem = get_ua_item;
        notify(h, DABSDR_NID_USER_APP_LIST, c ? DABSDR_NSTAT_SUCCESS : DABSDR_NSTAT_SERVICE_NOT_FOUND, &n, sizeof n);
        break; }
    case R_ANN_SUPPORT: {
        dabsdrNtfAnnouncementSupport_t n;
        memset(&n, 0, sizeof n);
        n.SId = r.a;
        fig::Service* s = h->db.find(r.a);
        if (s) { n.ASu = (uint16_t)(s->asu & 0xFFFE); n.numClusterIds = (uint8_t)std::min<size_t>(7, s->clusters.size()); memcpy(n.clusterIds, s->clusters.data(), n.numClusterIds); }
        notify(h, DABSDR_NID_ANNOUNCEMENT_SUPPORT, s ? DABSDR_NSTAT_SUCCESS : DABSDR_NSTAT_SERVICE_NOT_FOUND, &n, sizeof n);
        break; }
    case R_SELECT: {
        dabsdrNtfServiceSelection_t n;
        n.SId = r.a; n.SCIdS = (uint8_t)r.b; n.id = (dabsdrDecoderId_t)r.c;
        dabsdrNotificationStatus_t st = DABSDR_NSTAT_SUCCESS;
        fig::Service* s = nullptr;
        fig::Component* c = find_component(h, r.a, (uint8_t)r.b, &s);
        if (!c) st = DABSDR_NSTAT_SERVICE_NOT_FOUND;
        else if (r.c == DABSDR_ID_DATA) {
            // packet-mode data component (reference caller: src/radiocontrol.cpp:638); stream-mode data is not supported
            DataDecoder* slot = nullptr;
            for (auto& d : h->data) if (d.active && d.sid == r.a && d.scids == (uint8_t)r.b) slot = &d;
            if (!slot) for (auto& d : h->data) if (!d.active) { slot = &d; break; }
            if (c->tmid != 3 || !slot) st = DABSDR_NSTAT_SERVICE_NOT_SUPPORTED;
            else if (c->subch_id < 0 || !h->db.subch[c->subch_id].valid || c->packet_addr < 0) st = DABSDR_NSTAT_SERVICE_NOT_READY;
            else {
                *slot = DataDecoder();
                slot->active = true; slot->sid = r.a; slot->scids = (uint8_t)r.b; slot->subch_id = c->subch_id; slot->scid = c->scid;
                slot->ua_type = c->user_apps.empty() ? 0xFFFF : c->user_apps[0].type;
                slot->as.reset(c->packet_addr);
                if (h->frame_no >= 0) slot->last_row = h->frame_no * 4 + 3;
            }
        }
        else if (r.c != DABSDR_ID_AUDIO_PRIMARY && r.c != DABSDR_ID_AUDIO_SECONDARY) st = DABSDR_NSTAT_SERVICE_NOT_SUPPORTED;
        else if (c->tmid != 0) st = DABSDR_NSTAT_SERVICE_NOT_SUPPORTED;
        else if (c->subch_id < 0 || !h->db.subch[c->subch_id].valid) st = DABSDR_NSTAT_SERVICE_NOT_READY;
        else {
            Decoder& d = h->dec[r.c];
            d = Decoder();
            d.active = true; d.sid = r.a; d.scids = (uint8_t)r.b; d.subch_id = c->subch_id; d.dabplus = (c->ascty_dscty == 63);
            if (h->frame_no >= 0) { d.last_row = h->frame_no * 4 + 3; d.last_sf_row = d.last_row; }
        }
        notify(h, DABSDR_NID_SERVICE_SELECTION, st, &n, sizeof n);
        break; }
    case R_STOP: {
        dabsdrNtfServiceStop_t n;
        n.SId = r.a; n.SCIdS = (uint8_t)r.b; n.id = (dabsdrDecoderId_t)r.c;
        if (r.c == 0 || r.c == 1) h->dec[r.c].active = false;
        else if (r.c == DABSDR_ID_DATA) for (auto& d : h->data) if (d.active && d.sid == r.a && d.scids == (uint8_t)r.b) d.active = false;
        notify(h, DABSDR_NID_SERVICE_STOP, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
        break; }
    case R_XPAD: {
        if (r.c >= 0 && r.c < 2 && r.a < 32) h->dec[r.c].xpad_app[r.a] = r.b != 0;
        dabsdrNtfXpadAppStartStop_t n;
        n.appType = (uint8_t)r.a; n.start = (int8_t)r.b;
        notify(h, DABSDR_NID_XPAD_APP_START_STOP, DABSDR_NSTAT_SUCCESS, &n, sizeof n);
        break; }
    case R_PERIODIC:
        h->period = (int)r.a;
        h->next_periodic = h->abs_pos + h->have + ((int64_t)DAB_T_FRAME << h->period);
        notify(h, DABSDR_NID_PERIODIC, DABSDR_NSTAT_SUCCESS, nullptr, 0);       // first one carries no data
        break;
    case R_TII:
        h->tii_on = r.a != 0; h->tii_mode = r.c;
        h->tii_fold.clear(); h->tii_norm.clear(); h->tii_sign.clear();
        std::fill(h->tii_spectrum.begin(), h->tii_spectrum.end(), 0.f);
        if (h->tii_on && h->freq != 0) send_tii(h, nullptr);     // untuned: the empty list follows NID_RESET of the next tune
        break;
    case R_SPECTRUM:
        h->spectrum_on = r.a != 0;
        break;
    }
    return true;
}

// ---- one GPU pass over the current window --------------------------------------------------------------------------
// A complete, CRC-correct MSC data group arrived in the X-PAD of audio decoder `did`: hand it to the data group callback if the
// caller started that X-PAD application (RadioControl::dabXPadAppStart, src/radiocontrol.cpp:612); the user application type is
// the one FIG 0/13 announces for this X-PAD application type (first byte of the user application data, bits 4..0).
void deliver_xpad_dg(dabsdr_s* h, dabsdrDecoderId_t did, Decoder& d, int app, const uint8_t* p, int n)
{
    if (!h->dg_cb || app < 0 || app >= 32 || !d.xpad_app[app]) return;
    uint16_t ua_type = 0xFFFF;
    fig::Service* s = nullptr;
    if (fig::Component* c = find_component(h, d.sid, d.scids, &s))
        for (auto& u : c->user_apps) if (u.data_len >= 1 && (u.data[0] & 0x1F) == app) ua_type = u.type;
    dabsdrDataGroupCBData_t c;
    c.id = did; c.SCId = 0xFFFF;                              // the reference reports SCId 0xFFFF for X-PAD applications
    c.userAppType = ua_type; c.dgLen = (uint16_t)n; c.pDgData = p;
    h->dg_cb(&c, h->dg_ctx);
}

void send_periodic(dabsdr_s* h);

void process_window(dabsdr_s* h)
{
    // sub-channels to decode: those of the active decoders (none -> FIC only)
    b200dab_subch_t sc[2 + MAX_DATA_DEC];
    int n_sc = 0, sc_dec[2 + MAX_DATA_DEC];
    for (int& x : sc_dec) x = -1;
    for (int k = 0; k < 2; k++)
        if (h->dec[k].active && h->db.subch[h->dec[k].subch_id].valid) {
            sc[n_sc] = h->db.subch[h->dec[k].subch_id].p;
            sc[n_sc].kind = h->dec[k].dabplus ? 1 : 0;
            sc_dec[n_sc++] = k;
        }
    const int n_audio_sc = n_sc;
    for (auto& d : h->data) {                                   // packet-mode sub-channels (several components may share one)
        if (!d.active || !h->db.subch[d.subch_id].valid) continue;
        bool have = false;
        for (int i = n_audio_sc; i < n_sc; i++) if (sc[i].subch_id == d.subch_id) have = true;
        if (have) continue;
        sc[n_sc] = h->db.subch[d.subch_id].p;
        sc[n_sc].kind = 0;
        n_sc++;
    }
    uint32_t flags = B200DAB_OPT_SYMSTAT | (h->tii_on ? B200DAB_OPT_TII : 0u) | (n_sc == 0 ? B200DAB_OPT_NO_MSC : 0u) |
                     (h->spectrum_on && h->spec_cb ? B200DAB_OPT_SPECTRUM : 0u);
    b200dab_result_t r;
    if (b200dab_decode(h->gpu, h->win.data(), (uint64_t)h->have, B200DAB_FMT_F32, B200DAB_MEM_HOST, n_sc ? sc : nullptr, n_sc, flags, &r) != 0) {
        fprintf(stderr, "dabsdr(b200): GPU decode failed: %s\n", b200dab_last_error(h->gpu));
        return;
    }
    if (r.n_frames == 0) { send_sync(h, DABSDR_SYNC_LEVEL_NO_SYNC); return; }
    h->last_snr = r.snr_db;
    // absolute frame numbers of the window's slots, counted from the newest frame already accounted for
    std::vector<int64_t> slot_no(r.n_slots, -1);
    std::vector<int> fresh;                                     // slots holding a frame seen for the first time, in order
    {
        int64_t no = h->frame_no, last_abs = h->last_frame_abs;
        const int64_t anchor_no = h->frame_no, anchor_tf = h->last_frame_abs;
        for (int s = 0; s < r.n_slots; s++) {
            if (r.t_f[s] < 0) continue;
            const int64_t abs_tf = h->abs_pos + r.t_f[s];
            int64_t n;
            if (anchor_tf < 0) n = (last_abs < 0) ? 0 : no + (abs_tf - last_abs + DAB_T_FRAME / 2) / DAB_T_FRAME;
            else {
                const int64_t d = abs_tf - anchor_tf;
                n = anchor_no + (d >= 0 ? (d + DAB_T_FRAME / 2) / DAB_T_FRAME : -((-d + DAB_T_FRAME / 2) / DAB_T_FRAME));
            }
            slot_no[s] = n;
            if (n <= no && last_abs >= 0) continue;             // decoded already in an earlier window
            no = n;
            last_abs = abs_tf;
            fresh.push_back(s);
        }
    }
    // Results are handed out in the order and rhythm of a live stream whatever the size of the batch: ADV_FRAMES_MIN new frames at a
    // time -- their FIC, then the logical frames / access units that are complete once those frames are in (row <= 4 (n + 1) - 16 for
    // newest frame n: 15 CIFs of interleaver memory), TII every 40 frames, the periodic notification when its time has come.
    for (size_t f0 = 0; f0 < fresh.size() || f0 == 0; f0 += ADV_FRAMES_MIN) {
    const size_t f1 = std::min(fresh.size(), f0 + (size_t)ADV_FRAMES_MIN);
    const bool last_part = f1 >= fresh.size();
    bool any_fic = false;
    for (size_t fi = f0; fi < f1; fi++) {
        const int s = fresh[fi];
        h->frame_no = slot_no[s]; h->last_frame_abs = h->abs_pos + r.t_f[s];
        // ---- a frame seen for the first time: FIC
        h->last_cfo = r.cfo_hz[s];
        for (int f = 0; f < 12; f++) {
            if (r.fib_ok[(size_t)s * 12 + f]) { h->db.parse_fib(r.fibs + ((size_t)s * 12 + f) * 32); any_fic = true; send_fig_events(h); }
            else h->fib_err++;
        }
        if (r.spectrum && h->spectrum_on && h->spec_cb) {
            // like the reference: per frame one spectrum of the null symbol (TII when that null carries TII) and one SIGNAL spectrum
            const float* sp = r.spectrum + (size_t)s * 2 * DAB_NFFT;
            const bool tii_null = r.tii_sign && r.tii_sign[s] > 0;
            h->spec_cb(sp + DAB_NFFT, tii_null ? DABSDR_SPECT_TII : DABSDR_SPECT_NULL, h->spec_ctx);
            h->spec_cb(sp, DABSDR_SPECT_SIGNAL, h->spec_ctx);
        }
        if (h->tii_on && r.tii_fold) {
            h->tii_fold.insert(h->tii_fold.end(), r.tii_fold + (size_t)s * TII_SPECTRUM_BINS, r.tii_fold + (size_t)(s + 1) * TII_SPECTRUM_BINS);
            h->tii_norm.push_back(r.tii_norm[s]);
            h->tii_sign.push_back(r.tii_sign[s]);
        }
    }
    send_sync(h, any_fic || h->sync_level == DABSDR_SYNC_LEVEL_FIC ? DABSDR_SYNC_LEVEL_FIC : DABSDR_SYNC_LEVEL_ON_NULL);
    const int64_t row_limit = last_part ? INT64_MAX : 4 * (h->frame_no + 1) - 16;

    // ---- MSC: logical frame `row` of the window starts at CIF (row) = slot row/4, CIF row%4
    if (r.frames && n_sc > n_audio_sc) {
        auto abs_row = [&](int row) -> int64_t { const int s = row >> 2; return (s < r.n_slots && slot_no[s] >= 0) ? slot_no[s] * 4 + (row & 3) : -1; };
        for (auto& d : h->data) {
            if (!d.active) continue;
            for (int i = n_audio_sc; i < r.n_subch && i < n_sc; i++) {
                if (r.subch[i].subch_id != d.subch_id) continue;
                const int rows = r.frame_rows[i], pitch = r.subch[i].bitrate * 3;
                for (int row = 0; row < rows; row++) {
                    const int64_t ar = abs_row(row);
                    if (ar < 0 || ar <= d.last_row) continue;
                    if (ar > row_limit) break;
                    d.last_row = ar;
                    d.as.frame(r.frames[i] + (size_t)row * pitch, pitch, [&](const uint8_t* p, int n) {
                        if (!h->dg_cb) return;
                        dabsdrDataGroupCBData_t c;
                        c.id = DABSDR_ID_DATA; c.SCId = d.scid; c.userAppType = d.ua_type; c.dgLen = (uint16_t)n; c.pDgData = p;
                        h->dg_cb(&c, h->dg_ctx);
                    });
                }
            }
        }
    }
    if (r.frames && n_audio_sc > 0) {
        for (int i = 0; i < r.n_subch && i < n_audio_sc; i++) {
            Decoder& d = h->dec[sc_dec[i]];
            const int rows = r.frame_rows[i], pitch = r.subch[i].bitrate * 3;
            auto abs_row = [&](int row) -> int64_t { const int s = row >> 2; return (s < r.n_slots && slot_no[s] >= 0) ? slot_no[s] * 4 + (row & 3) : -1; };
            if (!d.dabplus) {
                for (int row = 0; row < rows; row++) {
                    const int64_t ar = abs_row(row);
                    if (ar < 0 || ar <= d.last_row) continue;
                    if (ar > row_limit) break;
                    d.last_row = ar;
                    const uint8_t* fr = r.frames[i] + (size_t)row * pitch;
                    const bool hdr_ok = mp2::check_frame(fr, pitch) == 1;      // ISO 11172-3 CRC; a failing frame is delivered empty
                    if (sc_dec[i] == 0) { if (hdr_ok) h->crc_ok++; else h->crc_err++; h->audio_bytes += pitch; }
                    if (h->audio_cb) {
                        dabsdrAudioCBData_t a;
                        // header = dynamic range control of this frame, which travels in the F-PAD of the PREVIOUS frame: byte L >> 2
                        // when bit 0 of byte L-1 is set ("byte L carries DRC data", EN 300 401 sec. 7.4.1), else 0 -- as the reference
                        // reports it (golden multi_mp2_u8_25db; consumer src/audiodecoder.cpp:285-326: gain = 10^(mp2DRC / 80))
                        a.id = (dabsdrDecoderId_t)sc_dec[i]; a.ASCTy = 0; a.header.raw = d.mp2_drc; a.auLen = (uint16_t)(hdr_ok ? pitch : 0); a.pAuData = fr;
                        h->audio_cb(&a, h->audio_ctx);
                    }
                    d.mp2_drc = (hdr_ok && (fr[pitch - 2] & 1)) ? (uint8_t)(fr[pitch - 1] >> 2) : 0;
                    int scf = 0;
                    if (hdr_ok && pad::mp2_pad(fr, pitch, h->pad_tmp, &scf)) {
                        const dabsdrDecoderId_t did = (dabsdrDecoderId_t)sc_dec[i];
                        const int used = d.xpad.feed(h->pad_tmp.data(), (int)h->pad_tmp.size(),
                            [&](const uint8_t* p, int n) {
                                if (!h->dl_cb) return;
                                dabsdrDynamicLabelCBData_t c; c.id = did; c.len = (uint16_t)n; c.pData = p;
                                h->dl_cb(&c, h->dl_ctx);
                            },
                            [&](int app, const uint8_t* p, int n) { deliver_xpad_dg(h, did, d, app, p, n); });
                        // the reference counts X-PAD + F-PAD + ScF-CRC + 1 per frame (probed: 15 / 21 / 25 for 6 / 12 / 16 data bytes)
                        if (sc_dec[i] == 0) h->pad_bytes += (uint32_t)(used + 2 + scf + 1);
                    }
                }
            } else {
                int64_t newest = d.last_sf_row;
                for (int k = 0; k < r.n_aus; k++) {
                    const b200dab_au_t& au = r.aus[k];
                    if (au.subch_index != i) continue;
                    const int64_t ar = abs_row(au.cif);
                    if (ar < 0 || ar <= d.last_sf_row || ar + 4 > row_limit) continue;
                    if (ar > newest) newest = ar;
                    if (sc_dec[i] == 0) { if (au.crc_ok) h->crc_ok++; else h->crc_err++; h->audio_bytes += au.length; }
                    if (h->audio_cb) {
                        dabsdrAudioCBData_t a;
                        a.id = (dabsdrDecoderId_t)sc_dec[i]; a.ASCTy = 63;
                        a.header.raw = (uint8_t)((au.header & 0x7F) | (au.crc_ok ? 0 : 0x80));       // bit 7 = concealment request
                        a.auLen = (uint16_t)au.length; a.pAuData = r.frames[i] + au.offset;
                        h->audio_cb(&a, h->audio_ctx);
                    }
                    // programme associated data of this AU: complete dynamic-label groups go out right after the audio callback
                    const uint8_t* padb; int padn;
                    if (au.crc_ok && pad::dabplus_pad(r.frames[i] + au.offset, au.length, &padb, &padn)) {
                        if (sc_dec[i] == 0) h->pad_bytes += (uint32_t)padn;
                        const dabsdrDecoderId_t did = (dabsdrDecoderId_t)sc_dec[i];
                        d.xpad.feed(padb, padn,
                            [&](const uint8_t* p, int n) {
                                if (!h->dl_cb) return;
                                dabsdrDynamicLabelCBData_t c; c.id = did; c.len = (uint16_t)n; c.pData = p;
                                h->dl_cb(&c, h->dl_ctx);
                            },
                            [&](int app, const uint8_t* p, int n) { deliver_xpad_dg(h, did, d, app, p, n); });
                    }
                }
                d.last_sf_row = newest;
                // RS statistics of the primary decoder, superframe by superframe, each counted once (dabsdr.h:159-165).  The reference
                // counts rsBytes = 24 * (bitrate / 8) per superframe -- 288 at 96 kbit/s (golden c1_u8_40db: 1440 / 1728 / 2016 for
                // 5 / 6 / 7 superframes per period) -- and starts with the second superframe it decodes; the application shows
                // rsBitErrors * 0.125 / rsBytes (src/ensembleinfobackend.cpp:438-440).
                if (sc_dec[i] == 0) {
                    for (int k = 0; k < r.n_sf; k++) {
                        const b200dab_sf_t& q = r.sf[k];
                        if (q.subch_index != i) continue;
                        const int64_t ar = abs_row(q.row);
                        if (ar < 0 || ar <= d.last_stat_row || ar + 4 > row_limit) continue;
                        if (d.last_stat_row >= 0) h->rs_bytes += (uint32_t)(24 * (r.subch[i].bitrate / 8));
                        h->rs_bit += (uint32_t)q.rs_bit_errors;
                        h->rs_unc += (uint32_t)q.rs_uncorrectable;
                        d.last_stat_row = ar;
                    }
                }
            }
        }
    }

    // ---- TII: one notification per 40 new frames, spectrum cumulative like the reference
    while (h->tii_on && h->tii_norm.size() >= TII_NOTIFY_FRAMES) {
        b200dab_tii_t t;
        const float thr_k = h->tii_mode == DABSDR_TII_MODE_CONSERVATIVE ? 10.f : 6.f;
        if (b200dab_stage_tii_detect(h->gpu, h->tii_fold.data(), h->tii_norm.data(), h->tii_sign.data(), TII_NOTIFY_FRAMES, thr_k, &t) == 0) {
            for (int b = 0; b < TII_SPECTRUM_BINS; b++) h->tii_spectrum[b] += t.spectrum[b];
            send_tii(h, &t);
        }
        h->tii_fold.erase(h->tii_fold.begin(), h->tii_fold.begin() + (size_t)TII_NOTIFY_FRAMES * TII_SPECTRUM_BINS);
        h->tii_norm.erase(h->tii_norm.begin(), h->tii_norm.begin() + TII_NOTIFY_FRAMES);
        h->tii_sign.erase(h->tii_sign.begin(), h->tii_sign.begin() + TII_NOTIFY_FRAMES);
    }
    // the periodic notification at the feed position a live stream would have reached with these frames in
    const int64_t feed = last_part ? h->abs_pos + h->have : h->last_frame_abs + DAB_T_FRAME + DAB_T_FRAME / 2;
    while (h->period >= 0 && feed >= h->next_periodic) {
        send_periodic(h);
        h->next_periodic += (int64_t)DAB_T_FRAME << h->period;
    }
    if (last_part) break;
    }   // parts of ADV_FRAMES_MIN new frames
}

void send_periodic(dabsdr_s* h)
{
    dabsdrNtfPeriodic_t p;
    memset(&p, 0, sizeof p);
    p.syncLevel = (dabsdrSyncLevel_t)h->sync_level;
    p.snr10 = (int16_t)lrintf(h->last_snr * 10.f);
    p.freqOffset = (int32_t)lrintf(h->last_cfo * 10.f);
    p.dateHoursMinutes = h->db.date_hours_minutes; p.secMsec = h->db.sec_msec;
    p.fibErrorCntr = (uint16_t)std::min<uint32_t>(h->fib_err, 65535);
    p.mscCrcOkCntr = (uint8_t)std::min<uint32_t>(h->crc_ok, 255); p.mscCrcErrorCntr = (uint8_t)std::min<uint32_t>(h->crc_err, 255);
    p.audioServiceBytes = (uint16_t)std::min<uint32_t>(h->audio_bytes, 65535);
    p.padBytes = (uint16_t)std::min<uint32_t>(h->pad_bytes, 65535);       // bytes of the data_stream_elements that carry the PAD
    p.rsUncorrectableCntr = (uint16_t)std::min<uint32_t>(h->rs_unc, 65535); p.rsBitErrors = (uint16_t)std::min<uint32_t>(h->rs_bit, 65535);
    p.rsBytes = (uint16_t)std::min<uint32_t>(h->rs_bytes, 65535);
    h->fib_err = h->crc_ok = h->crc_err = h->audio_bytes = h->pad_bytes = h->rs_unc = h->rs_bit = h->rs_bytes = 0;
    notify(h, DABSDR_NID_PERIODIC, DABSDR_NSTAT_SUCCESS, &p, sizeof p);
}

void* worker(void* arg)
{
    dabsdr_s* h = (dabsdr_s*)arg;
    pthread_setname_np(pthread_self(), "dabsdr");
    for (;;) {
        // requests are served between GPU passes; while untuned the worker sleeps on the mailbox
        {
            std::unique_lock<std::mutex> lk(h->mx);
            if (h->freq == 0 || !h->in || !h->gpu) h->cv.wait(lk, [&] { return !h->q.empty(); });
            while (!h->q.empty()) {
                Request r = h->q.front();
                h->q.pop_front();
                lk.unlock();
                const bool go = handle_request(h, r);
                if (!go) { if (h->gpu) { if (h->win_fallback.empty()) b200dab_host_free(h->gpu, h->win_data); b200dab_destroy(h->gpu); } h->gpu = nullptr; h->win_data = nullptr; return nullptr; }
                lk.lock();
            }
        }
        if (h->freq == 0 || !h->in || !h->gpu) continue;
        // pull one chunk (blocks inside the input function = flow control); requests are looked at again after every chunk,
        // i.e. every 8 ms of signal, so a request/notification round trip costs less than one transmission frame
        {
            if (h->t_pull0 == 0) { h->t_pull0 = wall_ms(); h->pulled0 = h->have; }
            const int64_t n = std::min<int64_t>(16384, h->win_samples() - h->have);
            float* dst = h->win.data() + 2 * h->have;
            h->in(dst, (uint16_t)n);
            h->have += n;
            // A chunk of exact zeros is the end of the source (InputDevice zero-fills after flush(), src/input/inputdevice.cpp:95-108;
            // a receiver never delivers 16384 exact zeros): what a large batch still holds is decoded now instead of waiting for it to fill.
            bool zeros = dst[0] == 0.f && dst[1] == 0.f;
            for (int64_t i = 0; zeros && i < 2 * n; i++) zeros = dst[i] == 0.f;
            if (zeros && !h->at_end && h->adv_frames > ADV_FRAMES_MIN && h->have > (int64_t)(ADV_FRAMES_MIN + OVL_FRAMES) * DAB_T_FRAME + 4096) {
                h->at_end = true;
                h->t_pull0 = 0; h->fast_pulls = 0;
                process_window(h);
                const int64_t keep = (int64_t)OVL_FRAMES * DAB_T_FRAME + 4096, adv = h->have - keep;
                memmove(h->win.data(), h->win.data() + 2 * adv, (size_t)keep * 2 * sizeof(float));
                h->have = keep; h->abs_pos += adv;
                h->adv_frames = ADV_FRAMES_MIN;
                continue;
            }
            if (!zeros) h->at_end = false;
        }
        if (h->have < h->win_samples()) continue;
        // A live stream delivers one second of signal per second; a file (RawFileInput without pacing, a scanner replay) much faster.
        // Latency only matters for the former, so the batch of new frames per GPU pass doubles (up to ADV_FRAMES_MAX) while the pull
        // of a batch takes less than a quarter of its signal time, and falls back to ADV_FRAMES_MIN as soon as it does not.
        {
            const double took = wall_ms() - h->t_pull0, signal_ms = (double)(h->have - h->pulled0) * 1000.0 / DAB_FS;
            h->t_pull0 = 0;
            if (took < 0.25 * signal_ms) h->fast_pulls++; else { h->fast_pulls = 0; }
        }
        const int64_t adv = h->adv_samples();
        process_window(h);
        while (h->period >= 0 && h->abs_pos + h->have >= h->next_periodic) {
            send_periodic(h);
            h->next_periodic += (int64_t)DAB_T_FRAME << h->period;
        }
        // slide by the frames of this batch
        memmove(h->win.data(), h->win.data() + 2 * adv, (size_t)(h->have - adv) * 2 * sizeof(float));
        h->have -= adv;
        h->abs_pos += adv;
        if (h->fast_pulls >= 2 && !h->at_end && h->adv_frames < h->adv_max) h->adv_frames = std::min(h->adv_max, h->adv_frames * 2);
        else if (h->fast_pulls == 0) h->adv_frames = ADV_FRAMES_MIN;
    }
}

void post(dabsdr_s* h, const Request& r)
{
    if (!h) return;
    std::lock_guard<std::mutex> lk(h->mx);
    h->q.push_back(r);
    h->cv.notify_one();
}

}  // namespace

extern "C" {

DABSDR_API uint8_t dabsdrInit(dabsdrHandle_t* handle)
{
    if (!handle) return 1;
    *handle = new dabsdr_s();
    if (const char* e = getenv("B200DAB_DROPIN_MAX_BATCH")) (*handle)->adv_max = std::max(ADV_FRAMES_MIN, std::min(ADV_FRAMES_MAX, atoi(e)));
    return 0;
}

DABSDR_API void dabsdr(dabsdrHandle_t h)
{
    if (!h || h->started) return;
    // The CUDA context and the (pinned) sample window are set up here, once, before the worker starts: RadioControl calls dabsdr()
    // at application start (src/radiocontrol.cpp:93), long before the first Tune.
    if (!h->gpu && b200dab_create(&h->gpu, 0) != 0) {
        fprintf(stderr, "dabsdr(b200): no CUDA device -- this build has no CPU fallback\n");
        h->gpu = nullptr;
    }
    if (h->gpu && !h->win_data) h->win_data = (float*)b200dab_host_alloc(h->gpu, (size_t)WIN_SAMPLES_MAX * 2 * sizeof(float));
    if (!h->win_data) { h->win_fallback.resize((size_t)WIN_SAMPLES_MAX * 2); h->win_data = h->win_fallback.data(); }
    h->win.p = h->win_data;
    h->started = pthread_create(&h->thr, nullptr, worker, h) == 0;
}

DABSDR_API void dabsdrDeinit(dabsdrHandle_t* handle)
{
    if (!handle || !*handle) return;
    dabsdr_s* h = *handle;
    if (h->started) {
        Request r; r.type = R_EXIT;
        post(h, r);
        pthread_join(h->thr, nullptr);
    }
    delete h;
    *handle = nullptr;
}

DABSDR_API void dabsdrGetVersion(dabsdrVersion_t* v)
{
    if (!v) return;
    v->major = 4; v->minor = 0; v->patch = 1; v->flags = 0xB2;        // API level of libdabsdr 4.0.1; flags mark the B200 build
}

DABSDR_API void dabsdrRegisterInputFcn(dabsdrHandle_t h, dabsdrInputFunc_t f) { if (h) h->in = f; }
DABSDR_API void dabsdrRegisterDummyInputFcn(dabsdrHandle_t h, dabsdrInputFunc_t f) { if (h) h->dummy = f; }
DABSDR_API void dabsdrRegisterAudioCb(dabsdrHandle_t h, dabsdrAudioCBFunc_t f, void* c) { if (h) { h->audio_cb = f; h->audio_ctx = c; } }
DABSDR_API void dabsdrRegisterDynamicLabelCb(dabsdrHandle_t h, dabsdrDynamicLabelCBFunc_t f, void* c) { if (h) { h->dl_cb = f; h->dl_ctx = c; } }
DABSDR_API void dabsdrRegisterDataGroupCb(dabsdrHandle_t h, dabsdrDataGroupCBFunc_t f, void* c) { if (h) { h->dg_cb = f; h->dg_ctx = c; } }
DABSDR_API void dabsdrRegisterSignalSpectrumCb(dabsdrHandle_t h, dabsdrSpectrumCBFunc_t f, void* c) { if (h) { h->spec_cb = f; h->spec_ctx = c; } }
DABSDR_API void dabsdrRegisterNotificationCb(dabsdrHandle_t h, dabsdrNotificationCBFunc_t f, void* c) { if (h) { h->ntf_cb = f; h->ntf_ctx = c; } }

DABSDR_API void dabsdrRequest_Tune(dabsdrHandle_t h, uint32_t frequency) { Request r; r.type = R_TUNE; r.a = frequency; post(h, r); }
DABSDR_API void dabsdrRequest_GetEnsemble(dabsdrHandle_t h) { Request r; r.type = R_ENSEMBLE; post(h, r); }
DABSDR_API void dabsdrRequest_GetServiceList(dabsdrHandle_t h) { Request r; r.type = R_SERVICE_LIST; post(h, r); }
DABSDR_API void dabsdrRequest_GetServiceComponents(dabsdrHandle_t h, uint32_t SId) { Request r; r.type = R_SERVICE_COMPS; r.a = SId; post(h, r); }
DABSDR_API void dabsdrRequest_GetUserAppList(dabsdrHandle_t h, uint32_t SId, uint8_t SCIdS) { Request r; r.type = R_USER_APPS; r.a = SId; r.b = SCIdS; post(h, r); }
DABSDR_API void dabsdrRequest_GetAnnouncementSupport(dabsdrHandle_t h, uint32_t SId) { Request r; r.type = R_ANN_SUPPORT; r.a = SId; post(h, r); }
DABSDR_API void dabsdrRequest_ServiceSelection(dabsdrHandle_t h, uint32_t SId, uint8_t SCIdS, dabsdrDecoderId_t id)
{
    Request r; r.type = R_SELECT; r.a = SId; r.b = SCIdS; r.c = (int)id; post(h, r);
}
DABSDR_API void dabsdrRequest_ServiceStop(dabsdrHandle_t h, uint32_t SId, uint8_t SCIdS, dabsdrDecoderId_t id)
{
    Request r; r.type = R_STOP; r.a = SId; r.b = SCIdS; r.c = (int)id; post(h, r);
}
DABSDR_API void dabsdrRequest_XPadAppStart(dabsdrHandle_t h, uint8_t appType, int8_t startRequest, dabsdrDecoderId_t id)
{
    Request r; r.type = R_XPAD; r.a = appType; r.b = (uint32_t)(uint8_t)startRequest; r.c = (int)id; post(h, r);
}
DABSDR_API void dabsdrRequest_SetPeriodicNotify(dabsdrHandle_t h, uint8_t period, uint32_t cfg) { Request r; r.type = R_PERIODIC; r.a = period; r.b = cfg; post(h, r); }
DABSDR_API void dabsdrRequest_SetTII(dabsdrHandle_t h, uint8_t ena, dabsdrTiiMode_t mode) { Request r; r.type = R_TII; r.a = ena; r.c = (int)mode; post(h, r); }
DABSDR_API void dabsdrRequest_SignalSpectrum(dabsdrHandle_t h, uint8_t ena) { Request r; r.type = R_SPECTRUM; r.a = ena; post(h, r); }
DABSDR_API void dabsdrRequest_Exit(dabsdrHandle_t h) { Request r; r.type = R_EXIT; post(h, r); }

}  // extern "C"
