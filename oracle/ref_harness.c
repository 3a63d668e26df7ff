/*
 * ref_harness.c -- TEST INFRASTRUCTURE (oracle side).  Not part of the product path.
 *
 * Drives a library exporting the dabsdr C API (the reference's closed libdabsdr.so.4.0.1, or this repo's
 * drop-in built from abracadabra_b200/csrc/dabsdr_shim.cpp) over a raw IQ file exactly the way the reference's
 * RadioControl does (reference: src/radiocontrol.cpp:81-104 init, :131-141 sync -> GetEnsemble,
 * :1381-1617 ensemble -> service list -> components -> ServiceSelection, :2479-2493 TII), with the sample feed
 * of RawFileInput (reference: src/input/rawfileinput.cpp:643-711 conversion, src/input/inputdevice.cpp:70-131
 * getSamples/skipSamples) but without the 50 ms pacing timer.
 *
 * Every callback is written to a JSON-lines event log.  The feed is LOCK-STEPPED: whenever a notification has
 * been delivered, the next input call blocks until the main thread has handled it and issued its follow-up
 * requests, so request timing is deterministic in signal time (SURVEY.md Appendix B "alignment").
 *
 * The ABI types are declared in include/dabsdr_abi.h (this repo's own restatement of lib/linux_x86_64/dabsdr.h).
 *
 * usage: ref_harness <iqfile> <u8|s16|f32> <out.jsonl|-> [--tii MODE] [--service N | --sid HEX] [--spectrum]
 *                    [--freq KHZ] [--tail-frames N] [--no-au-data] [--second N] [--free-run]
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include "dabsdr_abi.h"

/* ------------------------------------------------------------------ sample feed */
static const uint8_t *g_file;
static size_t g_file_bytes;
static int g_fmt;              /* 0 u8, 1 s16, 2 f32 */
static uint64_t g_nsamples;    /* complex samples in file */
static uint64_t g_pos;         /* next complex sample */
static uint64_t g_tail_samples = 8 * 196608ull;
static volatile int g_eof_seen, g_finished;
static struct timespec g_t0, g_t_eof;

/* ------------------------------------------------------------------ event queue (worker -> main) */
typedef struct ev_s {
    struct ev_s *next;
    int kind;                  /* nid (>0) or -1 tick, -2 eof */
    int status;
    uint64_t pos;
    uint8_t payload[4096];
    int payload_len;
    uint32_t sids[64]; int nsids;                     /* SERVICE_LIST shadow (accessor only valid inside the callback) */
    uint32_t sc_sid; int sc_n, sc_addr0, sc_scids0;   /* SERVICE_COMPONENT_LIST shadow */
    uint8_t xpad_apps[8]; int n_xpad_apps;            /* USER_APP_LIST shadow: X-PAD application types */
    int n_pkt; int pkt_scids[8]; int pkt_addr[8];     /* packet-mode components of the SERVICE_COMPONENT_LIST */
    int raw_len;                                      /* `len` as the core reported it */
} ev_t;

static pthread_mutex_t g_mx = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv_main = PTHREAD_COND_INITIALIZER, g_cv_worker = PTHREAD_COND_INITIALIZER;
static ev_t *g_head, *g_tail;
static int g_pending;

static FILE *g_log;
static pthread_mutex_t g_logmx = PTHREAD_MUTEX_INITIALIZER;
static int g_log_au_data = 1;

static dabsdrHandle_t g_h;

static void post_event(ev_t *e)
{
    e->next = NULL;
    e->pos = g_pos;
    pthread_mutex_lock(&g_mx);
    if (g_tail) g_tail->next = e; else g_head = e;
    g_tail = e;
    g_pending++;
    pthread_cond_signal(&g_cv_main);
    pthread_mutex_unlock(&g_mx);
}

static void wait_main_idle(void)
{
    pthread_mutex_lock(&g_mx);
    while (g_pending > 0) pthread_cond_wait(&g_cv_worker, &g_mx);
    pthread_mutex_unlock(&g_mx);
}

static uint64_t g_next_tick = 196608;
static int g_free_run;     /* --free-run: no lock-step gating (throughput measurements) */

static void advance(uint64_t n)
{
    uint64_t before = g_pos;
    g_pos += n;
    if (before < g_nsamples && g_pos >= g_nsamples && !g_eof_seen) {
        clock_gettime(CLOCK_MONOTONIC, &g_t_eof);
        g_eof_seen = 1;
    }
    if (g_pos >= g_next_tick) {
        g_next_tick += 196608;
        ev_t *e = calloc(1, sizeof *e);
        e->kind = -1;
        post_event(e);
    }
    if (g_pos >= g_nsamples + g_tail_samples && !g_finished) {
        g_finished = 1;
        ev_t *e = calloc(1, sizeof *e);
        e->kind = -2;
        post_event(e);
    }
}

static void input_fn(float buf[], uint16_t n)
{
    if (!g_free_run) wait_main_idle();
    uint64_t p = g_pos;
    uint64_t avail = p < g_nsamples ? g_nsamples - p : 0;
    uint32_t m = n < avail ? n : (uint32_t)avail;
    if (g_fmt == 0) {
        const uint8_t *s = g_file + 2 * p;
        for (uint32_t i = 0; i < 2 * m; i++) buf[i] = (float)(s[i] - 128);
    } else if (g_fmt == 1) {
        const int16_t *s = (const int16_t *)g_file + 2 * p;
        for (uint32_t i = 0; i < 2 * m; i++) buf[i] = (float)s[i];
    } else {
        memcpy(buf, (const float *)g_file + 2 * p, 2 * m * sizeof(float));
    }
    if (m < n) memset(buf + 2 * m, 0, 2 * (size_t)(n - m) * sizeof(float));
    advance(n);
}

static void dummy_fn(float buf[], uint16_t n)
{
    (void)buf;
    if (!g_free_run) wait_main_idle();
    advance(n);
}

/* ------------------------------------------------------------------ logging helpers */
static void log_hex(FILE *f, const uint8_t *p, int n)
{
    static const char hx[] = "0123456789abcdef";
    for (int i = 0; i < n; i++) { fputc(hx[p[i] >> 4], f); fputc(hx[p[i] & 15], f); }
}

static void log_label(FILE *f, const dabsdrLabel_t *l)
{
    fprintf(f, "\"label\":\"");
    for (int i = 0; i < 16 && l->str[i]; i++) {
        unsigned char c = (unsigned char)l->str[i];
        if (c == '"' || c == '\\' || c < 32 || c > 126) fprintf(f, "\\u%04x", c); else fputc(c, f);
    }
    fprintf(f, "\",\"charField\":%u,\"charset\":%u", l->charField, l->charset);
}

/* ------------------------------------------------------------------ callbacks (worker thread) */
static uint64_t g_n_au, g_au_bytes;

static void audio_cb(dabsdrAudioCBData_t *p, void *ctx)
{
    (void)ctx;
    g_n_au++;
    g_au_bytes += p->auLen;
    if (!g_log) return;
    pthread_mutex_lock(&g_logmx);
    fprintf(g_log, "{\"ev\":\"audio\",\"pos\":%llu,\"id\":%d,\"ascty\":%u,\"header\":%u,\"len\":%u,\"data\":\"",
            (unsigned long long)g_pos, (int)p->id, p->ASCTy, p->header.raw, p->auLen);
    if (g_log_au_data) log_hex(g_log, p->pAuData, p->auLen);
    fprintf(g_log, "\"}\n");
    pthread_mutex_unlock(&g_logmx);
}

static void dl_cb(dabsdrDynamicLabelCBData_t *p, void *ctx)
{
    (void)ctx;
    if (!g_log) return;
    pthread_mutex_lock(&g_logmx);
    fprintf(g_log, "{\"ev\":\"dl\",\"pos\":%llu,\"id\":%d,\"len\":%u,\"data\":\"", (unsigned long long)g_pos, (int)p->id, p->len);
    log_hex(g_log, p->pData, p->len);
    fprintf(g_log, "\"}\n");
    pthread_mutex_unlock(&g_logmx);
}

static void dg_cb(dabsdrDataGroupCBData_t *p, void *ctx)
{
    (void)ctx;
    if (!g_log) return;
    pthread_mutex_lock(&g_logmx);
    fprintf(g_log, "{\"ev\":\"dg\",\"pos\":%llu,\"id\":%d,\"scid\":%u,\"uatype\":%u,\"len\":%u,\"data\":\"",
            (unsigned long long)g_pos, (int)p->id, p->SCId, p->userAppType, p->dgLen);
    log_hex(g_log, p->pDgData, p->dgLen);
    fprintf(g_log, "\"}\n");
    pthread_mutex_unlock(&g_logmx);
}

static uint64_t g_n_spec;
static int g_spectrum_bins;            /* --spectrum-bins N: the first N spectrum callbacks are logged with all 2048 bins */
static void spectrum_cb(const float *p, dabsdrSpectrum_t type, void *ctx)
{
    (void)ctx;
    g_n_spec++;
    if (!g_log) return;
    if ((int)g_n_spec <= g_spectrum_bins) {
        pthread_mutex_lock(&g_logmx);
        fprintf(g_log, "{\"ev\":\"spectrum_bins\",\"pos\":%llu,\"type\":%d,\"bins\":[", (unsigned long long)g_pos, (int)type);
        for (int i = 0; i < 2048; i++) fprintf(g_log, "%s%.7g", i ? "," : "", (double)p[i]);
        fprintf(g_log, "]}\n");
        pthread_mutex_unlock(&g_logmx);
    }
    /* 2048 linear-power bins in FFT order (src/radiocontrol.cpp:2604-2609, src/signalbackend.cpp:383-431) */
    double sum = 0, mx = 0; int arg = 0;
    for (int i = 0; i < 2048; i++) { sum += p[i]; if (p[i] > mx) { mx = p[i]; arg = i; } }
    double inband = 0;
    for (int i = 1; i <= 768; i++) inband += p[i] + p[2048 - i];
    pthread_mutex_lock(&g_logmx);
    fprintf(g_log, "{\"ev\":\"spectrum\",\"pos\":%llu,\"type\":%d,\"sum\":%.6g,\"max\":%.6g,\"argmax\":%d,\"inband\":%.6g,\"b0\":%.6g,\"b1\":%.6g,\"b900\":%.6g}\n",
            (unsigned long long)g_pos, (int)type, sum, mx, arg, inband, (double)p[0], (double)p[1], (double)p[900]);
    pthread_mutex_unlock(&g_logmx);
}

static void log_notification(const ev_t *e);

static void notif_cb(dabsdrNotificationCBData_t *p, void *ctx)
{
    (void)ctx;
    ev_t *e = calloc(1, sizeof *e);
    e->kind = (int)p->nid;
    e->status = (int)p->status;
    e->payload_len = -1;
    if (p->pData) {
        int n = p->len < sizeof e->payload ? p->len : (int)sizeof e->payload;
        /* TUNE / RESET: the caller dereferences pData as a 32-bit value whatever `len` says (src/radiocontrol.cpp:2261, :2437) */
        if ((p->nid == DABSDR_NID_TUNE || p->nid == DABSDR_NID_RESET) && n < 4) n = 4;
        /* same for the selection / stop structures (src/radiocontrol.cpp:2364-2372, :2378-2386) */
        if ((p->nid == DABSDR_NID_SERVICE_SELECTION || p->nid == DABSDR_NID_SERVICE_STOP) && n < (int)sizeof(dabsdrNtfServiceSelection_t))
            n = (int)sizeof(dabsdrNtfServiceSelection_t);
        /* unsolicited notifications: the caller reads / copies the whole structure (src/radiocontrol.cpp:2325, :2457, :2469) */
        if (p->nid == DABSDR_NID_USER_APP_UPDATE && n < (int)sizeof(dabsdrNtfUserAppUpdate_t)) n = (int)sizeof(dabsdrNtfUserAppUpdate_t);
        if (p->nid == DABSDR_NID_ANNOUNCEMENT_SWITCHING && n < (int)sizeof(dabsdrNtfAnnouncementSwitching_t)) n = (int)sizeof(dabsdrNtfAnnouncementSwitching_t);
        if (p->nid == DABSDR_NID_PTY && n < (int)sizeof(dabsdrNtfPTy_t)) n = (int)sizeof(dabsdrNtfPTy_t);
        if (p->nid == DABSDR_NID_ANNOUNCEMENT_SUPPORT && n < (int)sizeof(dabsdrNtfAnnouncementSupport_t)) n = (int)sizeof(dabsdrNtfAnnouncementSupport_t);
        e->raw_len = p->len;
        memcpy(e->payload, p->pData, n);
        e->payload_len = n;
    }
    /* list accessors are only valid inside the callback: resolve them here */
    static FILE *devnull;
    if (!g_log && !devnull) devnull = fopen("/dev/null", "w");
    if (p->pData) {
        FILE *lf = g_log ? g_log : devnull;
#define g_log lf
        pthread_mutex_lock(&g_logmx);
        switch (p->nid) {
        case DABSDR_NID_SERVICE_LIST: {
            const dabsdrNtfServiceList_t *l = p->pData;
            fprintf(g_log, "{\"ev\":\"service_list\",\"pos\":%llu,\"n\":%u,\"items\":[", (unsigned long long)g_pos, l->numServices);
            for (int i = 0; i < l->numServices; i++) {
                dabsdrServiceListItem_t it; memset(&it, 0, sizeof it);
                l->getServiceListItem(g_h, i, &it);
                if (i < 64) { e->sids[i] = it.sid; e->nsids = i + 1; }
                fprintf(g_log, "%s{\"sid\":%u,", i ? "," : "", it.sid);
                log_label(g_log, &it.label);
                fprintf(g_log, ",\"pty_s\":%u,\"pty_d\":%u,\"caid\":%u}", it.pty.s, it.pty.d, it.CAId);
            }
            fprintf(g_log, "]}\n");
            break; }
        case DABSDR_NID_SERVICE_COMPONENT_LIST: {
            const dabsdrNtfServiceComponentList_t *l = p->pData;
            fprintf(g_log, "{\"ev\":\"sc_list\",\"pos\":%llu,\"sid\":%u,\"n\":%u,\"items\":[", (unsigned long long)g_pos, l->SId, l->numServiceComponents);
            for (int i = 0; i < l->numServiceComponents; i++) {
                dabsdrServiceCompListItem_t it; memset(&it, 0, sizeof it);
                l->getServiceComponentListItem(g_h, i, &it);
                if (i == 0) { e->sc_addr0 = it.SubChAddr; e->sc_scids0 = it.SCIdS; }
                if (it.TMId == 3 && e->n_pkt < 8) { e->pkt_scids[e->n_pkt] = it.SCIdS; e->pkt_addr[e->n_pkt] = it.packetData.packetAddress; e->n_pkt++; }
                e->sc_sid = l->SId; e->sc_n = l->numServiceComponents;
                fprintf(g_log, "%s{\"scids\":%u,\"subchid\":%u,\"addr\":%d,\"size\":%u,\"prot\":%u,\"uep_fec\":%u,\"ps\":%u,\"lang\":%u,\"ca\":%u,\"tmid\":%u,\"ascty\":%u,\"bitrate\":%u,\"numUA\":%u,",
                        i ? "," : "", it.SCIdS, it.SubChId, it.SubChAddr, it.SubChSize, it.protectionLevel, it.uepIdx, it.ps, it.lang,
                        it.CAflag, it.TMId, it.streamAudio.ASCTy, it.streamAudio.bitRate, it.numUserApps);
                log_label(g_log, &it.label);
                if (it.TMId == 3)
                    fprintf(g_log, ",\"dscty\":%u,\"scid\":%u,\"dg\":%u,\"paddr\":%d", it.packetData.DSCTy, it.packetData.SCId, it.packetData.DGflag,
                            it.packetData.packetAddress);
                fprintf(g_log, "}");
            }
            fprintf(g_log, "]}\n");
            break; }
        case DABSDR_NID_USER_APP_LIST: {
            const dabsdrNtfUserAppList_t *l = p->pData;
            fprintf(g_log, "{\"ev\":\"ua_list\",\"pos\":%llu,\"status\":%d,\"sid\":%u,\"scids\":%u,\"n\":%u,\"items\":[", (unsigned long long)g_pos,
                    (int)p->status, l->SId, l->SCIdS, l->numUserApps);
            e->n_xpad_apps = 0;
            for (int i = 0; i < l->numUserApps; i++) {
                dabsdrUserAppListItem_t it; memset(&it, 0, sizeof it);
                l->getUserAppListItem(g_h, i, &it);
                fprintf(g_log, "%s{\"type\":%u,\"len\":%u,\"data\":\"", i ? "," : "", it.type, it.dataLen);
                log_hex(g_log, it.data, it.dataLen);
                fprintf(g_log, "\"}");
                /* X-PAD user application: first data byte = CA flags + X-PAD application type (reference: src/radiocontrol.cpp:1546-1560) */
                if (it.dataLen >= 2 && e->n_xpad_apps < 8) e->xpad_apps[e->n_xpad_apps++] = it.data[0] & 0x1F;
            }
            fprintf(g_log, "]}\n");
            break; }
        case DABSDR_NID_TII: {
            const dabsdrNtfTii_t *t = p->pData;
            float spec[400]; memset(spec, 0, sizeof spec);
            if (t->getSpectrumTii) t->getSpectrumTii(g_h, spec);
            fprintf(g_log, "{\"ev\":\"tii\",\"pos\":%llu,\"n\":%u,\"ids\":[", (unsigned long long)g_pos, t->numIds);
            for (int i = 0; i < t->numIds && i < 24; i++)
                fprintf(g_log, "%s[%u,%u,%.9g]", i ? "," : "", t->id[i].main, t->id[i].sub, t->id[i].level);
            fprintf(g_log, "],\"spectrum\":[");
            for (int i = 0; i < 384; i++) fprintf(g_log, "%s%.9g", i ? "," : "", spec[i]);
            fprintf(g_log, "]}\n");
            break; }
        default: break;
        }
        pthread_mutex_unlock(&g_logmx);
#undef g_log
    }
    e->pos = g_pos;
    log_notification(e);
    post_event(e);
}

/* ------------------------------------------------------------------ main-thread protocol (mirrors RadioControl) */
static uint32_t g_freq = 227360;
static int g_data_scids = -1, g_selected_data;   /* SCIdS of a packet-mode component to start on the data decoder */
static int g_xpad_apps;                 /* ask for the user application list and start the X-PAD applications */
static int g_want_service_idx = -1;     /* index into the service list, -1 = none */
static int g_want_second_idx = -1;
static uint32_t g_want_sid;
static uint32_t g_sids[64]; static int g_nsids;
static int g_need_ensemble, g_have_ensemble, g_selected, g_selected2, g_fic_sync;
static uint32_t g_retry_sid; static int g_retry_ticks;
static int g_list_retry_ticks;          /* empty service list: ask again one frame later (reference: src/radiocontrol.cpp:1427-1436, 100 ms timer) */
static uint64_t g_fib_err, g_rs_bit, g_rs_unc, g_crc_ok, g_crc_err, g_n_periodic;

/* logging happens inside the callback (exact callback order); handle_event only runs the RadioControl-like protocol */
static void log_notification(const ev_t *e)
{
    FILE *f = g_log;
    if (!f) return;
    pthread_mutex_lock(&g_logmx);
    switch (e->kind) {
    case DABSDR_NID_SYNC_STATUS: {
        const dabsdrNtfSyncStatus_t *s = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"sync\",\"pos\":%llu,\"level\":%d,\"snr10\":%d}\n", (unsigned long long)e->pos, (int)s->syncLevel, s->snr10);
        break; }
    case DABSDR_NID_TUNE:
        fprintf(f, "{\"ev\":\"tune\",\"pos\":%llu,\"freq\":%u}\n", (unsigned long long)e->pos, e->payload_len >= 4 ? *(const uint32_t *)e->payload : 0);
        break;
    case DABSDR_NID_RESET:
        fprintf(f, "{\"ev\":\"reset\",\"pos\":%llu,\"flag\":%d}\n", (unsigned long long)e->pos, e->payload_len >= 4 ? *(const int *)e->payload : -1);
        break;
    case DABSDR_NID_ENSEMBLE_INFO: {
        const dabsdrNtfEnsemble_t *en = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"ensemble\",\"pos\":%llu,\"status\":%d,\"freq\":%u,\"ueid\":%u,\"lto\":%d,\"inttable\":%u,\"alarm\":%u,",
                (unsigned long long)e->pos, e->status, en->frequency, en->ueid, en->LTO, en->intTable, en->alarm);
        log_label(f, &en->label);
        fprintf(f, "}\n");
        break; }
    case DABSDR_NID_SERVICE_SELECTION: {
        const dabsdrNtfServiceSelection_t *s = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"selection\",\"pos\":%llu,\"status\":%d,\"sid\":%u,\"scids\":%u,\"id\":%d}\n",
                (unsigned long long)e->pos, e->status, s->SId, s->SCIdS, (int)s->id);
        break; }
    case DABSDR_NID_PERIODIC: {
        if (e->payload_len < 0) { fprintf(f, "{\"ev\":\"periodic\",\"pos\":%llu,\"null\":1}\n", (unsigned long long)e->pos); break; }
        const dabsdrNtfPeriodic_t *pp = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"periodic\",\"pos\":%llu,\"level\":%d,\"snr10\":%d,\"freqOffset\":%d,\"date\":%u,\"secmsec\":%u,\"fibErr\":%u,"
                   "\"crcOk\":%u,\"crcErr\":%u,\"audioBytes\":%u,\"padBytes\":%u,\"rsUnc\":%u,\"rsBitErr\":%u,\"rsBytes\":%u}\n",
                (unsigned long long)e->pos, (int)pp->syncLevel, pp->snr10, pp->freqOffset, pp->dateHoursMinutes, pp->secMsec,
                pp->fibErrorCntr, pp->mscCrcOkCntr, pp->mscCrcErrorCntr, pp->audioServiceBytes, pp->padBytes,
                pp->rsUncorrectableCntr, pp->rsBitErrors, pp->rsBytes);
        break; }
    case DABSDR_NID_SERVICE_LIST: case DABSDR_NID_SERVICE_COMPONENT_LIST: case DABSDR_NID_TII: case DABSDR_NID_USER_APP_LIST:
        break;                      /* logged with their accessor data in notif_cb */
    case DABSDR_NID_USER_APP_UPDATE: {
        const dabsdrNtfUserAppUpdate_t *u = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"ua_update\",\"pos\":%llu,\"status\":%d,\"sid\":%u,\"scids\":%u}\n", (unsigned long long)e->pos, e->status, u->SId, u->SCIdS);
        break; }
    case DABSDR_NID_PTY: {
        const dabsdrNtfPTy_t *t = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"pty\",\"pos\":%llu,\"status\":%d,\"sid\":%u,\"s\":%u,\"d\":%u}\n", (unsigned long long)e->pos, e->status, t->SId, t->s, t->d);
        break; }
    case DABSDR_NID_ANNOUNCEMENT_SWITCHING: {
        const dabsdrNtfAnnouncementSwitching_t *a = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"asw\",\"pos\":%llu,\"status\":%d,\"len\":%d,\"asw\":[", (unsigned long long)e->pos, e->status, e->raw_len);
        for (int i = 0; i < 8; i++) fprintf(f, "%s[%u,%u,%u]", i ? "," : "", a->asw[i].clusterId, a->asw[i].subChId, a->asw[i].ASwFlags);
        fprintf(f, "]}\n");
        break; }
    case DABSDR_NID_RECONFIGURATION:
        fprintf(f, "{\"ev\":\"reconfiguration\",\"pos\":%llu,\"status\":%d}\n", (unsigned long long)e->pos, e->status);
        break;
    case DABSDR_NID_ANNOUNCEMENT_SUPPORT: {
        const dabsdrNtfAnnouncementSupport_t *a = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"asu\",\"pos\":%llu,\"status\":%d,\"len\":%d,\"sid\":%u,\"asu\":%u,\"n\":%u,\"clusters\":[", (unsigned long long)e->pos, e->status, e->raw_len, a->SId, a->ASu, a->numClusterIds);
        for (int i = 0; i < a->numClusterIds && i < 7; i++) fprintf(f, "%s%u", i ? "," : "", a->clusterIds[i]);
        fprintf(f, "]}\n");
        break; }
    case DABSDR_NID_XPAD_APP_START_STOP: {
        const dabsdrNtfXpadAppStartStop_t *x = (const void *)e->payload;
        fprintf(f, "{\"ev\":\"xpad_app\",\"pos\":%llu,\"status\":%d,\"app\":%u,\"start\":%d}\n", (unsigned long long)e->pos, e->status, x->appType, x->start);
        break; }
    default:
        fprintf(f, "{\"ev\":\"other\",\"pos\":%llu,\"nid\":%d,\"status\":%d,\"len\":%d}\n", (unsigned long long)e->pos, e->kind, e->status, e->payload_len);
        break;
    }
    pthread_mutex_unlock(&g_logmx);
}

static void handle_event(ev_t *e)
{
    switch (e->kind) {
    case -1: /* tick, once per frame of signal */
        if (g_need_ensemble && g_fic_sync && !g_have_ensemble) dabsdrRequest_GetEnsemble(g_h);
        if (g_retry_ticks > 0 && --g_retry_ticks == 0) dabsdrRequest_GetServiceComponents(g_h, g_retry_sid);
        if (g_list_retry_ticks > 0 && --g_list_retry_ticks == 0) dabsdrRequest_GetServiceList(g_h);
        break;
    case DABSDR_NID_SYNC_STATUS: {
        const dabsdrNtfSyncStatus_t *s = (const void *)e->payload;
        if (s->syncLevel == DABSDR_SYNC_LEVEL_FIC && !g_fic_sync) {
            g_fic_sync = 1;
            dabsdrRequest_SetPeriodicNotify(g_h, 3, 0);
            if (!g_have_ensemble) { dabsdrRequest_GetEnsemble(g_h); g_need_ensemble = 1; }
        }
        if (s->syncLevel == DABSDR_SYNC_LEVEL_NO_SYNC) g_fic_sync = 0;
        break; }
    case DABSDR_NID_ENSEMBLE_INFO: {
        const dabsdrNtfEnsemble_t *en = (const void *)e->payload;
        /* accepted only once ECC != 0 and label non-empty (reference: src/radiocontrol.cpp:1396) */
        if (e->payload_len > 0 && ((en->ueid >> 16) & 0xFF) != 0 && en->label.str[0] != 0 && !g_have_ensemble) {
            g_have_ensemble = 1; g_need_ensemble = 0;
            dabsdrRequest_GetServiceList(g_h);
        }
        break; }
    case DABSDR_NID_RECONFIGURATION:
        /* reference: src/radiocontrol.cpp:178-190 -- the service list is dropped and requested again */
        g_nsids = 0;
        dabsdrRequest_GetServiceList(g_h);
        break;
    case DABSDR_NID_SERVICE_LIST:
        if (e->nsids == 0) { g_list_retry_ticks = 1; break; }
        g_nsids = e->nsids;
        for (int i = 0; i < e->nsids; i++) g_sids[i] = e->sids[i];
        /* reference: src/radiocontrol.cpp:1427-1468 -- one GetServiceComponents per service */
        for (int i = 0; i < e->nsids; i++) dabsdrRequest_GetServiceComponents(g_h, e->sids[i]);
        break;
    case DABSDR_NID_SERVICE_COMPONENT_LIST: {
        int want1 = 0, want2 = 0;
        if (g_want_sid && e->sc_sid == g_want_sid) want1 = 1;
        if (g_want_service_idx >= 0 && g_want_service_idx < g_nsids && g_sids[g_want_service_idx] == e->sc_sid) want1 = 1;
        if (g_want_second_idx >= 0 && g_want_second_idx < g_nsids && g_sids[g_want_second_idx] == e->sc_sid) want2 = 1;
        if ((want1 && !g_selected) || (want2 && !g_selected2)) {
            if (e->sc_n > 0 && e->sc_addr0 >= 0) {   /* reference: src/radiocontrol.cpp:1492-1498 re-requests while addr < 0 */
                dabsdrRequest_ServiceSelection(g_h, e->sc_sid, (uint8_t)e->sc_scids0,
                                               want1 ? DABSDR_ID_AUDIO_PRIMARY : DABSDR_ID_AUDIO_SECONDARY);
                if (want1) g_selected = 1; else g_selected2 = 1;
            } else { g_retry_sid = e->sc_sid; g_retry_ticks = 2; }
        }
        /* secondary packet-mode data component: started like RadioControl does (src/radiocontrol.cpp:638) once its packet address
         * is known (re-request while it is negative, :1494) */
        if (g_data_scids >= 0 && want1 && !g_selected_data) {
            int found = 0;
            for (int i = 0; i < e->n_pkt; i++) if (e->pkt_scids[i] == g_data_scids) {
                found = 1;
                if (e->pkt_addr[i] >= 0) { dabsdrRequest_ServiceSelection(g_h, e->sc_sid, (uint8_t)g_data_scids, DABSDR_ID_DATA); g_selected_data = 1; }
            }
            if (!g_selected_data && (found || e->n_pkt == 0)) { g_retry_sid = e->sc_sid; g_retry_ticks = 2; }
        }
        break; }
    case DABSDR_NID_SERVICE_SELECTION: {
        /* reference: after a successful audio selection RadioControl asks for the user applications of the component
         * (src/radiocontrol.cpp:1530-1545) and starts the X-PAD ones (:612) */
        const dabsdrNtfServiceSelection_t *s = (const void *)e->payload;
        if (g_xpad_apps && e->status == 0 && s->id == DABSDR_ID_AUDIO_PRIMARY) dabsdrRequest_GetUserAppList(g_h, s->SId, s->SCIdS);
        break; }
    case DABSDR_NID_USER_APP_LIST:
        for (int i = 0; i < e->n_xpad_apps; i++) dabsdrRequest_XPadAppStart(g_h, e->xpad_apps[i], 1, DABSDR_ID_AUDIO_PRIMARY);
        break;
    case DABSDR_NID_PERIODIC: {
        if (e->payload_len < 0) break;
        const dabsdrNtfPeriodic_t *pp = (const void *)e->payload;
        g_fib_err += pp->fibErrorCntr; g_rs_bit += pp->rsBitErrors; g_rs_unc += pp->rsUncorrectableCntr;
        g_crc_ok += pp->mscCrcOkCntr; g_crc_err += pp->mscCrcErrorCntr; g_n_periodic++;
        break; }
    default:
        break;
    }
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s <iqfile> <u8|s16|f32> <out.jsonl|-> [options]\n", argv[0]); return 2; }
    const char *path = argv[1];
    g_fmt = !strcmp(argv[2], "u8") ? 0 : !strcmp(argv[2], "s16") ? 1 : 2;
    int tii = -1, spectrum = 0;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--tii") && i + 1 < argc) tii = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--service") && i + 1 < argc) g_want_service_idx = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--second") && i + 1 < argc) g_want_second_idx = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sid") && i + 1 < argc) g_want_sid = (uint32_t)strtoul(argv[++i], NULL, 16);
        else if (!strcmp(argv[i], "--spectrum")) spectrum = 1;
        else if (!strcmp(argv[i], "--spectrum-bins") && i + 1 < argc) { spectrum = 1; g_spectrum_bins = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--freq") && i + 1 < argc) g_freq = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--tail-frames") && i + 1 < argc) g_tail_samples = 196608ull * atoi(argv[++i]);
        else if (!strcmp(argv[i], "--no-au-data")) g_log_au_data = 0;
        else if (!strcmp(argv[i], "--free-run")) g_free_run = 1;
        else if (!strcmp(argv[i], "--xpad-apps")) g_xpad_apps = 1;
        else if (!strcmp(argv[i], "--data-scids") && i + 1 < argc) g_data_scids = atoi(argv[++i]);
    }
    int fd = open(path, O_RDONLY);
    if (fd < 0) { perror(path); return 1; }
    struct stat st; fstat(fd, &st);
    g_file_bytes = st.st_size;
    g_file = mmap(NULL, g_file_bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (g_file == MAP_FAILED) { perror("mmap"); return 1; }
    g_nsamples = g_file_bytes / (g_fmt == 0 ? 2 : g_fmt == 1 ? 4 : 8);
    if (strcmp(argv[3], "-")) { g_log = fopen(argv[3], "w"); if (!g_log) { perror(argv[3]); return 1; } }
    if (g_log && getenv("HARNESS_LINEBUF")) setvbuf(g_log, NULL, _IOLBF, 0);      /* debugging aid: keep the log if the core crashes */

    dabsdrVersion_t v; dabsdrGetVersion(&v);
    if (dabsdrInit(&g_h) != 0) { fprintf(stderr, "dabsdrInit failed\n"); return 1; }
    dabsdrRegisterInputFcn(g_h, input_fn);
    dabsdrRegisterDummyInputFcn(g_h, dummy_fn);
    dabsdrRegisterNotificationCb(g_h, notif_cb, NULL);
    dabsdrRegisterDynamicLabelCb(g_h, dl_cb, NULL);
    dabsdrRegisterDataGroupCb(g_h, dg_cb, NULL);
    dabsdrRegisterAudioCb(g_h, audio_cb, NULL);
    dabsdrRegisterSignalSpectrumCb(g_h, spectrum_cb, NULL);
    if (g_log) fprintf(g_log, "{\"ev\":\"version\",\"major\":%u,\"minor\":%u,\"patch\":%u,\"flags\":%u,\"samples\":%llu}\n",
                       v.major, v.minor, v.patch, v.flags, (unsigned long long)g_nsamples);
    /* hold the worker at its first input call until the initial requests are in */
    pthread_mutex_lock(&g_mx); g_pending = 1; pthread_mutex_unlock(&g_mx);
    dabsdr(g_h);
    if (tii >= 0) dabsdrRequest_SetTII(g_h, 1, (dabsdrTiiMode_t)tii);
    if (spectrum) dabsdrRequest_SignalSpectrum(g_h, 1);
    clock_gettime(CLOCK_MONOTONIC, &g_t0);
    dabsdrRequest_Tune(g_h, g_freq);
    pthread_mutex_lock(&g_mx); g_pending = 0; pthread_cond_broadcast(&g_cv_worker); pthread_mutex_unlock(&g_mx);

    int done = 0;
    while (!done) {
        pthread_mutex_lock(&g_mx);
        while (!g_head) pthread_cond_wait(&g_cv_main, &g_mx);
        ev_t *e = g_head;
        g_head = e->next;
        if (!g_head) g_tail = NULL;
        pthread_mutex_unlock(&g_mx);

        handle_event(e);
        if (e->kind == -2) done = 1;
        free(e);
        pthread_mutex_lock(&g_mx);
        g_pending--;
        if (g_pending == 0) pthread_cond_broadcast(&g_cv_worker);
        pthread_mutex_unlock(&g_mx);
    }
    struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
    dabsdrRequest_Exit(g_h);
    /* let the worker run into the exit request */
    pthread_mutex_lock(&g_mx); g_pending = 0; pthread_cond_broadcast(&g_cv_worker); pthread_mutex_unlock(&g_mx);
    usleep(50000);
    dabsdrDeinit(&g_h);
    double sec = (g_t_eof.tv_sec - g_t0.tv_sec) + 1e-9 * (g_t_eof.tv_nsec - g_t0.tv_nsec);
    if (!g_eof_seen) sec = (t1.tv_sec - g_t0.tv_sec) + 1e-9 * (t1.tv_nsec - g_t0.tv_nsec);
    printf("{\"samples\":%llu,\"seconds\":%.6f,\"msps\":%.3f,\"aus\":%llu,\"au_bytes\":%llu,\"fib_err\":%llu,\"rs_bit_err\":%llu,"
           "\"rs_unc\":%llu,\"crc_ok\":%llu,\"crc_err\":%llu,\"periodic\":%llu,\"spectra\":%llu,\"version\":\"%u.%u.%u\"}\n",
           (unsigned long long)g_nsamples, sec, g_nsamples / sec / 1e6, (unsigned long long)g_n_au, (unsigned long long)g_au_bytes,
           (unsigned long long)g_fib_err, (unsigned long long)g_rs_bit, (unsigned long long)g_rs_unc, (unsigned long long)g_crc_ok,
           (unsigned long long)g_crc_err, (unsigned long long)g_n_periodic, (unsigned long long)g_n_spec, v.major, v.minor, v.patch);
    if (g_log) fclose(g_log);
    return 0;
}
