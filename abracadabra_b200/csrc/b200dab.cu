// b200dab.cu -- host side of libb200dab.so: context, device memory, stage orchestration and the C ABI of include/b200dab.h.
//
// The pipeline is the B200-native stand-in for the reference's closed receiver core (libdabsdr 4.0.1; interface
// lib/linux_x86_64/dabsdr.h:397-429): it is fed with what RawFileInput/InputDevice deliver (raw u8/s16 IQ, or the float
// stream of getSamples: src/input/inputdevice.cpp:70-108) and produces what RadioControl receives (FIBs -> ensemble
// structures, access units, statistics, TII lists).  No CPU fallback exists: without a CUDA device creation fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include "common.cuh"
#include "k_sync.cuh"
#include "k_ofdm.cuh"
#include "k_viterbi.cuh"
#include "k_dabplus.cuh"
#include "k_tii.cuh"
#include "tables.h"
#include "fig.h"
#include "pad.h"
#include "../../include/b200dab_snr_cal.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(#x, e_); return -1; } } while (0)
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->set_error(#x, e_); return -1; } } while (0)

// Small host->device table upload done by SMs reading pinned host memory (unified addressing) instead of the copy engine:
// on the copy engine it would queue behind the bulk IQ upload of another receiver context sharing the GPU.
__global__ void __launch_bounds__(256) k_fetch(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src_host, size_t nwords)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src_host[i];
}

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const bool g_trace = getenv("B200DAB_TRACE") != nullptr;
#define TRACE(msg) do { if (g_trace) fprintf(stderr, "[trace] %-28s %.3f ms\n", msg, now_ms() - ctx->t_trace0); } while (0)

// Host-to-device uploads of the receiver contexts sharing a GPU (ReceiverPool) are chained: the copy stream of a context waits for the
// last chunk of the upload queued before it, so recordings cross PCIe one after the other, back to back -- not interleaved chunk by
// chunk (every recording would arrive late) and without the copy engine idling between them.
#include <mutex>
static std::mutex g_upload_mu;
static cudaEvent_t g_upload_last[64] = {};        // per device: event behind the last chunk of the most recently queued upload

#define EARLY_FIC_FRAMES 16      /* frames of the early FIC slice: the first two prefixes (8, 16) of the sub-channel discovery */

#define RS_PATCH_PREFIX 8192       /* RS corrections copied to the host together with their count (64 KB) */
struct b200dab_ctx {
    int device = 0;
    double t_trace0 = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int launches = 0;
    cudaEvent_t ev_msc = nullptr, ev_d2h = nullptr;   // ordering of the split device-to-host copy of the logical frames
    cudaEvent_t ev_rs = nullptr, ev_tii = nullptr;    // behind k_rs_decode (DAB+ frames final) / behind k_tii_detect
    // Logical frames written to the host result buffer by k_vit_trace itself, group by group, next to the recursions that are still
    // running (B200DAB_FRAME_EXPORT=0: one device-to-host copy per class of sub-channels behind the trace / the RS correction instead).
    // The frames of the DAB+ sub-channels arrive as traced; what the RS decoder changes afterwards follows as a list of (offset, byte)
    // that the host applies to its copy.  The list has room for every byte the decoder can change (five per codeword); its first
    // rs_patch_prefix entries are copied blindly -- the count arrives with the totals -- and a recording with more corrections than that
    // (close to the decoding threshold) gets its corrected DAB+ frames copied again.
    bool frame_export = true;
    bool export_frames = false;        // set by the caller of msc() for this call
    bool frames_exported = false;      // answer of msc(): h_msc is being filled by the trace kernel
    uint8_t* trace_host_out = nullptr;
    int export_vec = 16;               // bytes per lane of the trace kernel's stores to host memory (B200DAB_EXPORT_VEC: 8 or 16)
    size_t rs_patch_prefix = RS_PATCH_PREFIX;     // B200DAB_RS_PATCH_PREFIX: fewer entries copied blindly (tests of the overflow path)
    bool rs_patch_mode = false;
    size_t rs_patch_cap = 0;
    // Everything the DAB+ kernels hand to the host lives in ONE device buffer and comes back in ONE copy (a small copy costs its latency,
    // and five of them in a row sat at the very end of a step): totals | access-unit records | SfInfo | fire-code flags | alignment
    // per sub-channel | RS patch list (only its first entries are copied).  Offsets in bytes, same layout in h_dpres.
    struct DpLayout { size_t aus = 0, sf = 0, flags = 0, res = 0, patch = 0, end = 0; } dp;
    DevBuf d_dpres;
    HostBuf h_dpres;
    template <typename T> T* dpp(size_t off) { return reinterpret_cast<T*>(d_dpres.as<uint8_t>() + off); }
    template <typename T> const T* hpp(size_t off) { return reinterpret_cast<const T*>(h_dpres.as<uint8_t>() + off); }
    int vit_ctas_per_sm = 2;       // persistent k_viterbi CTAs (four warps, one per sub-partition) per SM (tools/vit_split_sweep.py sweeps this)
    int n_sm = 148;
    DevBuf d_vit_counter;          // work queue heads of k_viterbi, one per table set + one for the stage entry point; zero between launches
    // Dynamic shared memory nobody uses, sized so that exactly vit_ctas_per_sm CTAs fit on an SM: the persistent CTAs are then spread
    // evenly however the SMs drain from the kernel in front (without it the block scheduler may pile four of them on the SMs that
    // are free first and leave others with one: the work queue evens out the items, not the issue slots).
    bool vit_smem_cap = true;
    bool trace_pdl = true;             // k_vit_trace(msc) as a programmatic dependent launch next to k_viterbi (B200DAB_TRACE_PDL=0: behind it)
    int vit_smem_slack = 4096;         // bytes per CTA left free below the cap: the trace CTAs running next to the recursions need 1 KB of (system) shared memory each
    size_t vit_dyn_smem() const { return vit_smem_cap ? (size_t)(227 * 1024 / vit_ctas_per_sm) - sizeof(uint32_t) * VIT_CTA_WARPS * 2 * VIT_NORM_PERIOD * 32 - (size_t)vit_smem_slack : 0; }
    unsigned vit_grid(size_t n_items) const { return (unsigned)std::min<size_t>((size_t)n_sm * vit_ctas_per_sm, (n_items + VIT_CTA_WARPS - 1) / VIT_CTA_WARPS); }
    // constant tables on the device
    DevBuf d_tw, d_fdi_pos, d_prs_q, d_prbs_w, d_gf, d_cls_bin, d_naus;   // d_gf: SfTables of k_sf_decode; d_naus: access units per superframe slot
    // pipeline buffers
    DevBuf d_raw, d_P, d_S, d_cand_raw, d_cand, d_ncand, d_fine, d_frames, d_soft, d_symstat;
    DevBuf d_jobs, d_groups, d_items, d_stepinfo, d_vin, d_dec, d_meet, d_fic_out, d_fib_ok, d_msc_out;
    struct VitTables {              // device-resident job / group / step tables of one call site, kept between calls
        bool valid = false;
        std::vector<VitJob> jobs_in;
        std::vector<uint32_t> stepinfo;
        size_t n_groups = 0, n_items = 0, words = 0;
        int max_steps = 0, vit4_min_steps = -1;
        DevBuf d_jobs, d_groups, d_stepinfo, d_items;
        DevBuf d_vin, d_dec, d_meet;   // work buffers of the set: the FIC chain runs on its own stream next to the MSC chain
        DevBuf d_order, d_done;        // trace order (CTA -> group); finished recursions per group, cleared before every launch
    } vit_tables[3];                   // [0] FIC, [1] MSC, [2] FIC of the first frames only (early sub-channel discovery)
    int vit_set = -1;                  // table set forced by the caller (-1: by job kind)
    cudaEvent_t ev_first_trace = nullptr;   // set by the caller: recorded behind the first of two k_vit_trace launches (see run_viterbi)
    bool first_trace_recorded = false;
    int vit4_min_steps = 0;            // codewords of at least this many trellis steps run in four-lane mode (k_viterbi.cuh); 0 = all
    // Work items of k_viterbi for `groups` (sorted), longest dependent instruction chain first.  trace_order (optional): the groups in
    // the order in which a simulation of the kernel's own schedule (n_warps persistent warps, each taking the next item when it is
    // through with its last) expects their recursions to be finished -- the order in which k_vit_trace, running next to k_viterbi as
    // its programmatic dependent, takes them (ties: the longer trace first).
    void build_items(const std::vector<VitJob>& jobs, const std::vector<VitGroup>& groups, std::vector<VitItem>& items,
                     std::vector<int>* trace_order = nullptr, int n_warps = 1) const
    {
        items.clear();
        std::vector<std::pair<long long, VitItem>> tmp;
        for (size_t g = 0; g < groups.size(); g++) {
            const VitJob& j = jobs[groups[g].job];
            // instructions per trellis step of one warp (tools/sass_loops.py): 186 (two lanes per codeword pair), 107 (four lanes)
            if (j.mode4) for (int w = 0; w < 4; w++) tmp.push_back({(long long)j.steps * 107, VitItem{(int)g, 4 | w}});
            else for (int w = 0; w < 2; w++) tmp.push_back({(long long)j.steps * 186, VitItem{(int)g, w}});
        }
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<long long, VitItem>& a, const std::pair<long long, VitItem>& b) { return a.first > b.first; });
        for (auto& t : tmp) items.push_back(t.second);
        if (!trace_order) return;
        n_warps = std::max(n_warps, 1);
        std::vector<int> left(groups.size());
        std::vector<long long> ready_at(groups.size(), 0);
        for (size_t g = 0; g < groups.size(); g++) left[g] = jobs[groups[g].job].mode4 ? 4 : 2;
        typedef std::pair<long long, int> Ev;                      // (end time, group) of a busy warp; min-heap on the time
        std::vector<Ev> heap;
        auto later = [](const Ev& a, const Ev& b) { return a.first > b.first; };
        long long now = 0;
        auto retire = [&]() {
            std::pop_heap(heap.begin(), heap.end(), later);
            const Ev e = heap.back(); heap.pop_back();
            now = e.first;
            if (--left[e.second] == 0) ready_at[e.second] = e.first;
        };
        for (auto& t : tmp) {
            if ((int)heap.size() >= n_warps) retire();
            heap.push_back({now + t.first, t.second.group}); std::push_heap(heap.begin(), heap.end(), later);
        }
        while (!heap.empty()) retire();
        trace_order->resize(groups.size());
        for (size_t g = 0; g < groups.size(); g++) (*trace_order)[g] = (int)g;
        std::stable_sort(trace_order->begin(), trace_order->end(), [&](int a, int b) {
            return ready_at[a] != ready_at[b] ? ready_at[a] < ready_at[b] : jobs[groups[a].job].steps > jobs[groups[b].job].steps; });
    }
    cudaStream_t cur = nullptr;        // stream the helpers below launch on (nullptr: `stream`)
    cudaStream_t S() const { return cur ? cur : stream; }
    cudaStream_t fic_stream = nullptr;
    cudaStream_t aux_stream = nullptr;       // DAB+ half of the back end behind k_viterbi: trace + superframe kernels, next to the trace of the other sub-channels
    cudaEvent_t ev_vit = nullptr, ev_aux = nullptr;
    bool late_on_aux = false;
    cudaEvent_t ev_early_demod = nullptr, ev_early_done = nullptr, ev_early_fic = nullptr, ev_demod = nullptr, ev_fic_stream = nullptr;
    DevBuf d_fic_out_early, d_fib_ok_early;
    HostBuf h_fibs_early, h_fib_ok_early;
    int early_frames = 0;              // > 0: front_end splits its first demodulation launch and starts the early FIC slice
    bool early_launched = false;
    DevBuf d_dpjobs, d_fold, d_norm, d_sign, d_windows, d_tii;
    DevBuf d_tmp_a, d_tmp_b, d_spec;
    HostBuf h_spec;
    HostBuf h_fine, h_frames, h_fibs, h_fib_ok, h_msc, h_tii, h_symstat, h_misc, h_fold, h_norm;
    HostBuf h_jobs[3], h_groups[3], h_stepinfo[3], h_items[3], h_order[3], h_dpjobs, h_sign, h_windows;   // pinned staging of small uploads ([0] FIC, [1] MSC)
    std::vector<int32_t> r_sign;
    std::vector<std::pair<b200dab_subch_t, std::vector<uint32_t>>> stepinfo_cache;   // de-puncturing tables by protection profile
    // result storage
    std::vector<int64_t> r_tf;
    std::vector<float> r_cfo, r_cir;
    std::vector<b200dab_subch_t> r_subch;
    std::vector<const uint8_t*> r_frames;
    std::vector<int32_t> r_rows;
    std::vector<b200dab_au_t> r_aus;
    std::vector<b200dab_sf_t> r_sf;
    std::vector<const uint8_t*> r_sf_flags;
    std::vector<int32_t> r_sf_phase, forced_phase;
    std::vector<b200dab_tii_t> r_tii;

    // optional per-kernel timing (CUDA events on `stream`), enabled by b200dab_set_profiling
    bool profile = false;
    // The events of the last PROF_RING calls are kept, so a caller can time K calls back to back and read the events afterwards
    // (reading them costs ~0.1 ms of host time per call, which does not belong inside a timed region).
    static constexpr int PROF_RING = 8;
    struct ProfEv { const char* name; cudaEvent_t a, b; };
    std::vector<ProfEv> prof_sets[PROF_RING];
    size_t prof_counts[PROF_RING] = {};
    int prof_slot = 0;
    void prof_next() { prof_slot = (prof_slot + 1) % PROF_RING; prof_counts[prof_slot] = 0; }
    void prof_begin(const char* name)
    {
        if (!profile) return;
        std::vector<ProfEv>& prof = prof_sets[prof_slot];
        size_t& used = prof_counts[prof_slot];
        if (used == prof.size()) { ProfEv e; e.name = name; cudaEventCreate(&e.a); cudaEventCreate(&e.b); prof.push_back(e); }
        prof[used].name = name;
        cudaEventRecord(prof[used].a, S());
    }
    void prof_end()
    {
        if (!profile) return;
        cudaEventRecord(prof_sets[prof_slot][prof_counts[prof_slot]].b, S());
        prof_counts[prof_slot]++;
    }

    void set_error(const char* what, cudaError_t e)
    {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        err = buf;
    }
    void set_error(const char* what) { err = what; }

    int fetch(void* d_dst, const void* h_pinned, size_t bytes)
    {
        const size_t nw = (bytes + 3) / 4;
        if (!nw) return 0;
        const unsigned grid = (unsigned)std::min<size_t>(64, (nw + 255) / 256);
        k_fetch<<<grid, 256, 0, S()>>>(reinterpret_cast<uint32_t*>(d_dst), reinterpret_cast<const uint32_t*>(h_pinned), nw);
        launches += 1;
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    int init_tables();
    // device-side stages (all asynchronous on `stream` unless noted)
    int upload_raw(const void* iq, uint64_t n, int fmt, int mem, const void** d_out);
    int coarse(const void* d_iq, uint64_t n, int fmt, int* n_cand_out, int* max_cand_out);
    int fine(const void* d_iq, uint64_t n, int fmt, int grid, int max_cand, long long range_off, int chain_range_idx, int max_slots);
    int prepare_slots(int max_slots, bool want_stat, bool want_tii);
    int demod(const void* d_iq, uint64_t n_visible, int fmt, int range_idx, int skip, int limit, int bound, bool want_stat, bool want_tii);
    int tii_fold(const void* d_iq, uint64_t n_visible, int fmt, int range_idx, int skip, int limit, int bound);
    DevBuf d_chain, d_frame_cir;       // ChainState + per-frame CIR ratio, written by the last CTA of k_fine_sync
    HostBuf h_chain, h_frame_cir;
    cudaEvent_t ev_chain = nullptr, ev_frames = nullptr;
    int fe_ranges = 0;                 // front-end ranges queued in this call
    bool defer_tii_fold = false;   // single-range decode: the fold is launched after the FIC so that it runs while the host parses the FIGs
    cudaEvent_t ev_fic = nullptr;
    int front_end(const void* d_iq, int fmt, uint64_t off, uint64_t len, bool want_stat, bool want_tii, int max_slots);
    // front-end state across the ranges of one decode
    size_t fe_frames = 0; int fe_candidates = 0;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    int run_viterbi(const std::vector<VitJob>& jobs_in, const std::vector<uint32_t>& stepinfo, uint8_t* d_out);
    int fic(int n_slots, bool early = false);
    int early_fic(const void* d_iq, uint64_t n_visible, int fmt, bool want_stat, bool want_tii);
    int msc(int n_slots, const std::vector<b200dab_subch_t>& subch, std::vector<size_t>& out_off, size_t* total);
    int dabplus(uint8_t* d_frames, const std::vector<DabPlusJob>& jobs, int total_sf, int total_rows);
    // A DISCOVERED sub-channel whose description cannot be decoded (capacity units beyond the CIF, a protection profile that does not fit
    // its size: what a FIB with a false-positive CRC at low SNR can produce) is left out -- the others are still decoded.  A table given
    // by the caller is taken as is (msc() fails on an inconsistent entry).
    int dropped_subch = 0;
    // De-puncturing table of a sub-channel (nullptr: inconsistent protection profile).  It depends on the bitrate and the (L, PI) segments
    // only, and ensembles rarely change between calls: built once per profile -- the sub-channel discovery sits between the early FIC
    // and the MSC chain on the critical path of a step, and building the tables of a full ensemble takes 0.1 ms.  The pointer is valid
    // until the next call.
    bool stepinfo_cached = true;
    const std::vector<uint32_t>* stepinfo_for(const b200dab_subch_t& q)
    {
        if (q.n_seg < 0 || q.n_seg > (int)(sizeof q.seg_len / sizeof q.seg_len[0])) return nullptr;
        if (!stepinfo_cached) stepinfo_cache.clear();             // B200DAB_STEPINFO_CACHE=0: built every time (A/B)
        for (auto& e : stepinfo_cache)
            if (e.first.bitrate == q.bitrate && e.first.n_seg == q.n_seg && !memcmp(e.first.seg_len, q.seg_len, sizeof q.seg_len) &&
                !memcmp(e.first.seg_pi, q.seg_pi, sizeof q.seg_pi)) return &e.second;
        std::vector<uint32_t> t;
        if (dabtab::build_stepinfo(q, t) < 0) return nullptr;
        if (stepinfo_cache.size() > 64) stepinfo_cache.clear();
        stepinfo_cache.emplace_back(q, std::move(t));
        return &stepinfo_cache.back().second;
    }
    void drop_undecodable(std::vector<b200dab_subch_t>& v)
    {
        size_t k = 0;
        for (size_t i = 0; i < v.size(); i++) {
            const std::vector<uint32_t>* t = stepinfo_for(v[i]);
            const int np = t ? (int)(t->back() >> 4) : -1;
            const bool ok = np >= 0 && np <= v[i].size_cu * 64 && v[i].start_cu >= 0 && v[i].start_cu + v[i].size_cu <= 864 && v[i].bitrate > 0;
            if (ok) v[k++] = v[i]; else dropped_subch++;
        }
        v.resize(k);
    }
};

static size_t fmt_bytes(int fmt) { return fmt == B200DAB_FMT_U8 ? 2 : fmt == B200DAB_FMT_S16 ? 4 : 8; }
#define FRONT_CHUNK_FRAMES 64     /* host uploads are pipelined in chunks of 64 transmission frames (25 MB of u8) */

// Spectrum-buffer layout of k_ofdm_demod (see the comment inside): out[p], p < 2048 = address of the FFT output at logical position p
// of the digit-reversed result (0xffff = unused bin); out[2048 + n], n < 1536 = address of QPSK symbol n.  Returns 0, or -1 if the
// colouring is not proper (cannot happen: Koenig).  Host-only, also exported for the tests (b200dab_spectrum_layout).
static int build_spectrum_layout(std::vector<uint16_t>& pos)
{
    std::vector<int16_t> c_of_n(DAB_NCARR);
    dabtab::freq_deint(c_of_n.data());
    // k_ofdm_demod scatters the spectrum straight into de-interleaved order.  Where QPSK symbol n lives inside the 1536-entry
    // spectrum buffer is free as long as it stays inside its group of 16 consecutive n (the DQPSK stage reads n = t + 256 j, so a
    // half-warp reads one such group).  The slot inside the group is chosen so that the half-warp STORES of the last FFT pass are
    // bank-conflict free as well: every n is an edge between its store group (half-warp, output index) and its read group, and a
    // proper 16-colouring of that bipartite graph (Koenig: max degree 16) gives every node 16 distinct 8-byte banks.
    //   table[p]        (p < 2048)  = address of the FFT output at logical position p of the digit-reversed result, 0xffff = unused bin
    //   table[2048 + n] (n < 1536)  = address of QPSK symbol n
    std::vector<int> n_of_p(DAB_NFFT, -1);
    for (int n = 0; n < DAB_NCARR; n++) n_of_p[fft_pos(carrier_bin(c_of_n[n]))] = n;
    const int NL = 128, NR = DAB_NCARR / 16;
    std::vector<int> colour(DAB_NCARR, -1), left_of(DAB_NCARR, -1);
    std::vector<std::vector<int>> at_l(NL, std::vector<int>(16, -1)), at_r(NR, std::vector<int>(16, -1));   // node x colour -> edge
    for (int t = 0; t < 256; t++)
        for (int h = 0; h < 2; h++)
            for (int j = 0; j < 4; j++) {
                const int n = n_of_p[4 * (t + 256 * h) + j];
                if (n >= 0) left_of[n] = (t >> 4) * 8 + h * 4 + j;
            }
    for (int n = 0; n < DAB_NCARR; n++) {
        const int u = left_of[n], v = n >> 4;
        int a = 0, b = 0;
        while (at_l[u][a] >= 0) a++;
        while (at_r[v][b] >= 0) b++;
        if (at_r[v][a] >= 0) {
            // colour a is taken at v: walk the a/b alternating path that starts at v and swap the two colours along it
            std::vector<int> path;
            int node = v, side = 1, want = a;
            for (;;) {
                const int e = side ? at_r[node][want] : at_l[node][want];
                if (e < 0) break;
                path.push_back(e);
                node = side ? left_of[e] : (e >> 4);
                side ^= 1;
                want = (want == a) ? b : a;
            }
            for (int e : path) { at_l[left_of[e]][colour[e]] = -1; at_r[e >> 4][colour[e]] = -1; }
            for (int e : path) { colour[e] = (colour[e] == a) ? b : a; at_l[left_of[e]][colour[e]] = e; at_r[e >> 4][colour[e]] = e; }
        }
        colour[n] = a; at_l[u][a] = n; at_r[v][a] = n;
    }
    pos.assign(DAB_NFFT + DAB_NCARR, 0xffff);
    for (int n = 0; n < DAB_NCARR; n++) pos[DAB_NFFT + n] = (uint16_t)(16 * (n >> 4) + colour[n]);
    for (int q = 0; q < DAB_NFFT; q++) if (n_of_p[q] >= 0) pos[q] = pos[DAB_NFFT + n_of_p[q]];
    for (int u = 0; u < NL; u++) for (int c = 0; c < 16; c++) for (int c2 = c + 1; c2 < 16; c2++)
        if (at_l[u][c] >= 0 && at_l[u][c] == at_l[u][c2]) return -1;
    return 0;
}

int b200dab_ctx::init_tables()
{
    std::vector<float> tw(4096);
    dabtab::twiddles(tw.data());
    CK(d_tw.ensure(tw.size() * 4));
    CK(cudaMemcpy(d_tw.p, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
    std::vector<int16_t> c_of_n(DAB_NCARR);
    dabtab::freq_deint(c_of_n.data());
    std::vector<uint16_t> pos;
    if (build_spectrum_layout(pos)) { set_error("de-interleaver colouring failed"); return -1; }
    CK(d_fdi_pos.ensure(pos.size() * 2));
    CK(cudaMemcpy(d_fdi_pos.p, pos.data(), pos.size() * 2, cudaMemcpyHostToDevice));
    std::vector<int8_t> prs(DAB_NCARR);
    dabtab::prs_quarter(prs.data());
    CK(d_prs_q.ensure(prs.size()));
    CK(cudaMemcpy(d_prs_q.p, prs.data(), prs.size(), cudaMemcpyHostToDevice));
    {   // carrier pairs by PRS phase step (integer CFO search)
        std::vector<uint16_t> cls;
        int off[5] = {0, 0, 0, 0, 0};
        for (int q = 0; q < 4; q++) {
            off[q] = (int)cls.size();
            for (int c = 0; c + 1 < DAB_NCARR; c++) {
                if (c == 767) continue;                                 // the pair would straddle DC
                if (((prs[c + 1] - prs[c]) & 3) == q) cls.push_back((uint16_t)carrier_bin(c));
            }
        }
        off[4] = (int)cls.size();
        CK(d_cls_bin.ensure(cls.size() * 2));
        CK(cudaMemcpy(d_cls_bin.p, cls.data(), cls.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpyToSymbol(c_cls_off, off, sizeof off));
    }
    std::vector<uint32_t> pw(512);
    dabtab::prbs_words(pw.data(), 512);
    CK(d_prbs_w.ensure(pw.size() * 4));
    CK(cudaMemcpy(d_prbs_w.p, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice));
    {   // tables of k_sf_decode: GF(2^8), t * g(x) of the RS(120,110) generator, CRC-16 slicing by four
        static SfTables tb;
        dabtab::gf_tables(tb.gf.exp, tb.gf.log);
        auto mul = [&](uint8_t a, uint8_t b) -> uint8_t { return (a && b) ? tb.gf.exp[tb.gf.log[a] + tb.gf.log[b]] : 0; };
        uint8_t g[12] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};           // g = 1, then times (x + alpha^i), i = 0..9
        for (int i = 0; i < 10; i++) {
            const uint8_t a = tb.gf.exp[i];
            for (int j = i + 1; j > 0; j--) g[j] = g[j - 1] ^ mul(g[j], a);
            g[0] = mul(g[0], a);
        }
        for (int t = 0; t < 256; t++) {
            uint32_t w[3] = {0, 0, 0};
            for (int j = 0; j < 10; j++) w[j >> 2] |= (uint32_t)mul((uint8_t)t, g[j]) << (8 * (j & 3));
            tb.fb[t][0] = w[0]; tb.fb[t][1] = w[1]; tb.fb[t][2] = w[2];
        }
        for (int i = 0; i < 256; i++) {
            uint32_t c = (uint32_t)i << 8;
            for (int b = 0; b < 8; b++) c = (c & 0x8000u) ? ((c << 1) ^ 0x1021u) & 0xffffu : (c << 1) & 0xffffu;
            tb.crc[0][i] = (uint16_t)c;
        }
        for (int k = 1; k < 4; k++)
            for (int i = 0; i < 256; i++) {
                const uint32_t c = tb.crc[k - 1][i];
                tb.crc[k][i] = (uint16_t)(((c << 8) & 0xffffu) ^ tb.crc[0][c >> 8]);
            }
        CK(d_gf.ensure(sizeof tb));
        CK(cudaMemcpy(d_gf.p, &tb, sizeof tb, cudaMemcpyHostToDevice));
    }
    uint8_t pat[70];
    dabtab::tii_patterns(pat);
    CK(cudaMemcpyToSymbol(c_tii_pattern, pat, 70));
    CK(cudaMemcpyToSymbol(c_tii_gain, B200DAB_TII_GAIN, sizeof(B200DAB_TII_GAIN)));
    return 0;
}

int b200dab_ctx::upload_raw(const void* iq, uint64_t n, int fmt, int mem, const void** d_out)
{
    if (mem == B200DAB_MEM_DEVICE) { *d_out = iq; return 0; }
    const size_t bytes = (size_t)n * fmt_bytes(fmt);
    CK(d_raw.ensure(bytes + 64));
    CK(cudaMemcpyAsync(d_raw.p, iq, bytes, cudaMemcpyHostToDevice, stream));
    *d_out = d_raw.p;
    return 0;
}

// ------------------------------------------------------------------------------------------------ coarse sync
// block powers, then one launch for window sums, candidate test and the sorted candidate list (d_cand, d_ncand)
int b200dab_ctx::coarse(const void* d_iq, uint64_t n, int fmt, int* n_cand_out, int* max_cand_out)
{
    const int64_t nblk = (int64_t)(n / SYNC_BLK);
    const int max_cand = (int)(nblk / SYNC_WIN) + 2;
    *max_cand_out = max_cand;
    if (n_cand_out) *n_cand_out = 0;
    if (nblk <= 0) return 0;
    const unsigned tiles = (unsigned)((nblk + CO_TILE - 1) / CO_TILE);
    CK(d_P.ensure((size_t)nblk * 4));
    CK(d_S.ensure((size_t)nblk * 4));
    CK(d_cand_raw.ensure((size_t)tiles * (CO_CAP * 8 + 4)));
    CK(d_cand.ensure((size_t)max_cand * 8));
    CK(d_ncand.ensure(16));                                   // [0] candidates, [1] ticket of k_coarse, [2] ticket of k_fine_sync
    CK(cudaMemsetAsync(d_ncand.p, 0, 16, stream));
    long long* tile_list = d_cand_raw.as<long long>();
    int* tile_cnt = reinterpret_cast<int*>(tile_list + (size_t)tiles * CO_CAP);
    int* nc = d_ncand.as<int>();
    const unsigned pgrid = (unsigned)((nblk + 255) / 256);
    prof_begin("k_block_power");
    if (fmt == B200DAB_FMT_U8) k_block_power<B200DAB_FMT_U8><<<pgrid, 256, 0, stream>>>(d_iq, nblk, d_P.as<float>());
    else if (fmt == B200DAB_FMT_S16) k_block_power<B200DAB_FMT_S16><<<pgrid, 256, 0, stream>>>(d_iq, nblk, d_P.as<float>());
    else k_block_power<B200DAB_FMT_F32><<<pgrid, 256, 0, stream>>>(d_iq, nblk, d_P.as<float>());
    prof_end();
    prof_begin("k_coarse");
    k_coarse<<<tiles, 256, 0, stream>>>(d_P.as<float>(), nblk, (int64_t)n, d_S.as<float>(), tile_list, tile_cnt, nc + 1, d_cand.as<long long>(), nc, max_cand);
    prof_end();
    launches += 1;
    launches += 1;
    CK(cudaGetLastError());
    if (n_cand_out) {                     // stage entry point: the caller wants the count now
        int c = 0;
        CK(cudaMemcpyAsync(&c, d_ncand.p, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (c > max_cand) c = max_cand;
        *n_cand_out = c;
    }
    return 0;
}

// Fine synchronisation of the candidates in d_cand (grid CTAs take them in turn); with chain_range >= 0 the last CTA chains the
// frames on the device (d_chain, d_frames, d_frame_cir), range_off = first sample of the range inside the recording.
int b200dab_ctx::fine(const void* d_iq, uint64_t n, int fmt, int grid, int max_cand, long long range_off, int chain_range_idx, int max_slots)
{
    if (grid <= 0) return 0;
    CK(d_fine.ensure((size_t)max_cand * sizeof(FineOut)));
    const float2* tw = d_tw.as<float2>();
    const int8_t* prs = d_prs_q.as<int8_t>();
    ChainState* ch = chain_range_idx >= 0 ? d_chain.as<ChainState>() : nullptr;
    int* nc = d_ncand.as<int>();
    prof_begin("k_fine_sync");
    if (fmt == B200DAB_FMT_U8)
        k_fine_sync<B200DAB_FMT_U8><<<grid, 256, 0, stream>>>(d_iq, (int64_t)n, d_cand.as<long long>(), nc, max_cand, tw, prs, d_cls_bin.as<uint16_t>(), d_fine.as<FineOut>(),
                                                              range_off, chain_range_idx, max_slots, ch, d_frames.as<FrameInfo>(), d_frame_cir.as<float>(), nc + 2);
    else if (fmt == B200DAB_FMT_S16)
        k_fine_sync<B200DAB_FMT_S16><<<grid, 256, 0, stream>>>(d_iq, (int64_t)n, d_cand.as<long long>(), nc, max_cand, tw, prs, d_cls_bin.as<uint16_t>(), d_fine.as<FineOut>(),
                                                               range_off, chain_range_idx, max_slots, ch, d_frames.as<FrameInfo>(), d_frame_cir.as<float>(), nc + 2);
    else
        k_fine_sync<B200DAB_FMT_F32><<<grid, 256, 0, stream>>>(d_iq, (int64_t)n, d_cand.as<long long>(), nc, max_cand, tw, prs, d_cls_bin.as<uint16_t>(), d_fine.as<FineOut>(),
                                                               range_off, chain_range_idx, max_slots, ch, d_frames.as<FrameInfo>(), d_frame_cir.as<float>(), nc + 2);
    prof_end();
    launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ OFDM demod
// soft-bit / statistics / TII buffers for up to max_slots frames; the frame chain of the call starts empty
int b200dab_ctx::prepare_slots(int max_slots, bool want_stat, bool want_tii)
{
    CK(d_frames.ensure((size_t)(max_slots + 1) * sizeof(FrameInfo)));
    CK(d_frame_cir.ensure((size_t)(max_slots + 1) * 4));
    CK(h_frames.ensure((size_t)(max_slots + 1) * sizeof(FrameInfo)));
    CK(h_frame_cir.ensure((size_t)(max_slots + 1) * 4));
    CK(h_chain.ensure(sizeof(ChainState)));
    CK(d_chain.ensure(sizeof(ChainState)));
    CK(cudaMemsetAsync(d_chain.p, 0, 32, stream));            // counters; range_first[] is written before it is read
    // Every frame that is found writes all of its soft bits; only the slots of LOST frames have to read as erasures, and those are
    // cleared by the demodulation CTAs of the frame behind the gap -- not the whole buffer (144 MB for 60 s) on every call.
    CK(d_soft.ensure((size_t)max_slots * DAB_FRAME_SOFT + 256));
    if (want_stat) {
        CK(d_symstat.ensure((size_t)max_slots * 150 * 4));
        CK(cudaMemsetAsync(d_symstat.p, 0, (size_t)max_slots * 150 * 4, stream));
    }
    if (want_tii) {
        CK(d_fold.ensure((size_t)max_slots * TII_SPECTRUM_BINS * 4));
        CK(d_norm.ensure((size_t)max_slots * 4));
        CK(cudaMemsetAsync(d_fold.p, 0, (size_t)max_slots * TII_SPECTRUM_BINS * 4, stream));
        CK(cudaMemsetAsync(d_norm.p, 0, (size_t)max_slots * 4, stream));
    }
    return 0;
}

// OFDM demodulation (and TII fold) of the frames of range `range_idx` of the device-side chain (< 0: all frames), leaving out `skip`
// frames at the front and taking at most `limit`; `bound` = upper bound of the number of frames (grid size); n_visible = samples
// valid in d_iq
int b200dab_ctx::demod(const void* d_iq, uint64_t n_visible, int fmt, int range_idx, int skip, int limit, int bound, bool want_stat, bool want_tii)
{
    if (bound <= 0) return 0;
    const FrameInfo* dfr = d_frames.as<FrameInfo>();
    const ChainState* ch = d_chain.as<ChainState>();
    float* st = want_stat ? d_symstat.as<float>() : nullptr;
    const unsigned grid = (unsigned)bound * OFDM_GROUPS;
    const float2* tw = d_tw.as<float2>();
    const uint16_t* fp = d_fdi_pos.as<uint16_t>();
    const int64_t n = (int64_t)n_visible;
    cudaStream_t stream = S();
    prof_begin("k_ofdm_demod");
    if (fmt == B200DAB_FMT_U8) k_ofdm_demod<B200DAB_FMT_U8><<<grid, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, fp, d_soft.as<int8_t>(), st);
    else if (fmt == B200DAB_FMT_S16) k_ofdm_demod<B200DAB_FMT_S16><<<grid, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, fp, d_soft.as<int8_t>(), st);
    else k_ofdm_demod<B200DAB_FMT_F32><<<grid, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, fp, d_soft.as<int8_t>(), st);
    prof_end();
    launches += 1;
    CK(cudaGetLastError());
    if (want_tii && !defer_tii_fold) {                    // the fold only reads the samples: it stays on the main stream
        cudaStream_t keep = cur;
        cur = nullptr;
        const int rc = tii_fold(d_iq, n_visible, fmt, range_idx, skip, limit, bound);
        cur = keep;
        return rc;
    }
    return 0;
}

// null-symbol spectra of the same selection of frames, folded to the 384 TII bins
int b200dab_ctx::tii_fold(const void* d_iq, uint64_t n_visible, int fmt, int range_idx, int skip, int limit, int bound)
{
    if (bound <= 0) return 0;
    const FrameInfo* dfr = d_frames.as<FrameInfo>();
    const ChainState* ch = d_chain.as<ChainState>();
    const float2* tw = d_tw.as<float2>();
    const int64_t n = (int64_t)n_visible;
    cudaStream_t stream = S();
    prof_begin("k_tii_fold");
    if (fmt == B200DAB_FMT_U8) k_tii_fold<B200DAB_FMT_U8><<<bound, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, d_fold.as<float>(), d_norm.as<float>());
    else if (fmt == B200DAB_FMT_S16) k_tii_fold<B200DAB_FMT_S16><<<bound, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, d_fold.as<float>(), d_norm.as<float>());
    else k_tii_fold<B200DAB_FMT_F32><<<bound, 256, 0, stream>>>(d_iq, n, dfr, ch, range_idx, skip, limit, tw, d_fold.as<float>(), d_norm.as<float>());
    prof_end();
    launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// FIC of the first `early_frames` frames on the FIC stream, behind their demodulation: FIG 0/1 and 0/2 repeat every few frames, so
// the sub-channel organisation is known (and the MSC chain can be queued) while the rest of the recording is still demodulated.
// Queued blindly (the frame list is still on the device): whether the slice is usable -- more than early_frames frames, none of the
// first ones lost -- is decided by the host when it has the list; an unusable slice is simply ignored.
int b200dab_ctx::early_fic(const void* d_iq, uint64_t n_visible, int fmt, bool want_stat, bool want_tii)
{
    const int k = early_frames;
    // The first frames are demodulated on the FIC stream as well, next to the rest of the recording on the main stream (64 CTAs on
    // their own would leave the GPU idle for the length of a demodulation CTA).  The FIC stream joins the main stream here (buffers
    // cleared at the start of the call, frames chained) and the main stream waits for these frames before anything reads the soft bits.
    // (ev_early_demod was recorded on the main stream before the rest of the recording was queued there)
    CK(cudaStreamWaitEvent(fic_stream, ev_early_demod, 0));
    cur = fic_stream;
    int rc = demod(d_iq, n_visible, fmt, 0, 0, k, k, want_stat, want_tii);
    if (!rc) rc = cudaEventRecord(ev_early_done, fic_stream) == cudaSuccess ? 0 : -1;
    if (!rc) rc = fic(k, true);
    cur = nullptr;
    if (rc) return -1;
    CK(cudaStreamWaitEvent(stream, ev_early_done, 0));
    CK(h_fibs_early.ensure((size_t)k * 12 * 32));
    CK(h_fib_ok_early.ensure((size_t)k * 12));
    CK(cudaMemcpyAsync(h_fibs_early.p, d_fic_out_early.p, (size_t)k * 12 * 32, cudaMemcpyDeviceToHost, fic_stream));
    CK(cudaMemcpyAsync(h_fib_ok_early.p, d_fib_ok_early.p, (size_t)k * 12, cudaMemcpyDeviceToHost, fic_stream));
    CK(cudaEventRecord(ev_early_fic, fic_stream));
    early_launched = true;
    return 0;
}

// Front end over the sample range [off, off + len) of the device buffer: coarse sync, fine sync + chaining with the frames found so
// far (on the device), OFDM demodulation of the new frames -- queued without a host round trip; the host reads the frame list once,
// behind the last range.  Ranges must come in ascending order and overlap by more than one frame so that a frame cut by a range
// boundary is found whole in the next range.
int b200dab_ctx::front_end(const void* d_iq, int fmt, uint64_t off, uint64_t len, bool want_stat, bool want_tii, int max_slots)
{
    const uint8_t* base = reinterpret_cast<const uint8_t*>(d_iq) + off * fmt_bytes(fmt);
    const int range_idx = fe_ranges;
    if (range_idx >= CHAIN_MAX_RANGES) { set_error("too many front-end ranges in one call"); return -1; }
    int max_cand = 0;
    if (coarse(base, len, fmt, nullptr, &max_cand)) return -1;
    if (max_cand <= 0) return 0;
    // one CTA per candidate up to twice the frames the range can hold; more candidates than that (noise-like input) are taken in turn
    const int bound_frames = (int)(len / DAB_T_FRAME) + 2;
    const int grid = std::min(max_cand, 2 * bound_frames + 64);
    if (fine(base, len, fmt, grid, max_cand, (long long)off, range_idx, max_slots)) return -1;
    fe_ranges++;
    CK(cudaEventRecord(ev_chain, stream));                // the frame list of everything up to this range is final here
    if (early_frames > 0 && !early_launched && range_idx == 0 && bound_frames > early_frames + 2) {
        // the first frames on their own, the early FIC slice behind them, then the rest
        CK(cudaEventRecord(ev_early_demod, stream));
        if (demod(d_iq, off + len, fmt, range_idx, early_frames, 1 << 30, bound_frames, want_stat, want_tii)) return -1;
        return early_fic(d_iq, off + len, fmt, want_stat, want_tii);
    }
    return demod(d_iq, off + len, fmt, range_idx, 0, 1 << 30, bound_frames, want_stat, want_tii);
}

// ------------------------------------------------------------------------------------------------ Viterbi (gather + ACS + traceback)
// jobs_in: vin_off/dec_off are filled here; out_off, row0, ... by the caller.  Soft bits are read from d_soft.
int b200dab_ctx::run_viterbi(const std::vector<VitJob>& jobs_in, const std::vector<uint32_t>& stepinfo, uint8_t* d_out)
{
    const int st = vit_set >= 0 ? vit_set : (jobs_in[0].kind == 0 ? 0 : 1);   // each call site owns its tables and work buffers
    cudaStream_t stream = S();
    VitTables& vt = vit_tables[st];
    // The job / group / step tables only depend on the sub-channel organisation and the number of CIFs: when they equal those of
    // the previous call (every window of a stream, every recording of one multiplex) what is already on the device is reused.
    const bool same = vt.valid && vt.vit4_min_steps == vit4_min_steps && vt.jobs_in.size() == jobs_in.size() && vt.stepinfo.size() == stepinfo.size() &&
                      !memcmp(vt.jobs_in.data(), jobs_in.data(), jobs_in.size() * sizeof(VitJob)) &&
                      !memcmp(vt.stepinfo.data(), stepinfo.data(), stepinfo.size() * 4);
    if (!same) {
        vt.valid = false;
        std::vector<VitJob> jobs = jobs_in;
        std::vector<VitGroup> groups;
        size_t words = 0;
        int max_steps = 0;
        for (size_t j = 0; j < jobs.size(); j++) {
            const int ng = (jobs[j].n_rows + 31) / 32;
            jobs[j].mode4 = (jobs[j].steps >= vit4_min_steps && jobs[j].steps % 2 == 0) ? 1 : 0;
            jobs[j].vin_off = (long long)words;
            jobs[j].dec_off = (long long)words;
            words += (size_t)ng * 32 * jobs[j].steps;
            for (int g = 0; g < ng; g++) groups.push_back(VitGroup{(int)j, g});
            max_steps = std::max(max_steps, jobs[j].steps);
        }
        vt.n_groups = groups.size(); vt.words = words; vt.max_steps = max_steps;
        if (groups.empty()) return 0;
        std::stable_sort(groups.begin(), groups.end(), [&](const VitGroup& a, const VitGroup& b) { return jobs[a.job].steps > jobs[b.job].steps; });
        std::vector<VitItem> items;
        std::vector<int> order;
        size_t n_it = 0;
        for (auto& g : groups) n_it += jobs[g.job].mode4 ? 4 : 2;
        build_items(jobs, groups, items, &order, (int)vit_grid(n_it) * VIT_CTA_WARPS);
        vt.n_items = items.size();
        CK(h_order[st].ensure(order.size() * 4 + 16));
        CK(vt.d_order.ensure(order.size() * 4 + 16));
        memcpy(h_order[st].p, order.data(), order.size() * 4);
        if (fetch(vt.d_order.p, h_order[st].p, order.size() * 4)) { set_error("k_fetch launch failed"); return -1; }
        // small tables go through pinned staging owned by the context (one set per call site), so no stream sync is needed here
        CK(h_items[st].ensure(items.size() * sizeof(VitItem) + 16));
        CK(vt.d_items.ensure(items.size() * sizeof(VitItem) + 16));
        memcpy(h_items[st].p, items.data(), items.size() * sizeof(VitItem));
        if (fetch(vt.d_items.p, h_items[st].p, items.size() * sizeof(VitItem))) { set_error("k_fetch launch failed"); return -1; }
        CK(h_jobs[st].ensure(jobs.size() * sizeof(VitJob) + 16));
        CK(h_groups[st].ensure(groups.size() * sizeof(VitGroup) + 16));
        CK(h_stepinfo[st].ensure(stepinfo.size() * 4 + 16));
        CK(vt.d_jobs.ensure(jobs.size() * sizeof(VitJob) + 16));
        CK(vt.d_groups.ensure(groups.size() * sizeof(VitGroup) + 16));
        CK(vt.d_stepinfo.ensure(stepinfo.size() * 4 + 16));
        memcpy(h_jobs[st].p, jobs.data(), jobs.size() * sizeof(VitJob));
        memcpy(h_groups[st].p, groups.data(), groups.size() * sizeof(VitGroup));
        memcpy(h_stepinfo[st].p, stepinfo.data(), stepinfo.size() * 4);
        if (fetch(vt.d_jobs.p, h_jobs[st].p, jobs.size() * sizeof(VitJob)) || fetch(vt.d_groups.p, h_groups[st].p, groups.size() * sizeof(VitGroup)) ||
            fetch(vt.d_stepinfo.p, h_stepinfo[st].p, stepinfo.size() * 4)) { set_error("k_fetch launch failed"); return -1; }
        vt.jobs_in = jobs_in; vt.stepinfo = stepinfo; vt.vit4_min_steps = vit4_min_steps; vt.valid = true;
    }
    if (vt.n_groups == 0) return 0;
    const size_t n_groups = vt.n_groups;
    DevBuf &d_vin = vt.d_vin, &d_dec = vt.d_dec, &d_meet = vt.d_meet;
    CK(d_vin.ensure(vt.words * 4));
    CK(d_dec.ensure(vt.words * 8));
    CK(d_meet.ensure(n_groups * VIT_MEET_WORDS * sizeof(unsigned short)));
    const VitJob* dj = vt.d_jobs.as<VitJob>();
    const VitGroup* dg = vt.d_groups.as<VitGroup>();
    // k_vit_trace as a programmatic dependent of k_viterbi (MSC set): it starts while the recursions are still running and takes the
    // groups as they finish (done[]: recursions finished per group, cleared here).  B200DAB_TRACE_PDL=0: plain launches (A/B).
    const bool pdl = trace_pdl && st == 1;
    if (pdl) {
        CK(vt.d_done.ensure((n_groups + 1) * 4));               // + the work queue head of k_viterbi: the trace kernel, which clears it
        CK(cudaMemsetAsync(vt.d_done.p, 0, (n_groups + 1) * 4, stream));   //   otherwise, now runs while the queue is in use
    }
    int* d_done = pdl ? vt.d_done.as<int>() : nullptr;
    int* d_head = pdl ? vt.d_done.as<int>() + n_groups : d_vit_counter.as<int>() + st;
    dim3 ggrid((unsigned)n_groups, (unsigned)((vt.max_steps + GATHER_STEPS - 1) / GATHER_STEPS));
    prof_begin(st == 0 ? "k_gather(fic)" : st == 2 ? "k_gather(early fic)" : "k_gather(msc)");
    k_gather<<<ggrid, 256, 0, stream>>>(d_soft.as<int8_t>(), dj, dg, vt.d_stepinfo.as<uint32_t>(), d_vin.as<uint32_t>());
    prof_end();
    prof_begin(st == 0 ? "k_viterbi(fic)" : st == 2 ? "k_viterbi(early fic)" : "k_viterbi(msc)");
    k_viterbi<<<vit_grid(vt.n_items), 32 * VIT_CTA_WARPS, st == 1 ? vit_dyn_smem() : 0, stream>>>(d_vin.as<uint32_t>(), d_dec.as<uint32_t>(), d_meet.as<unsigned short>(), dj, dg, vt.d_items.as<VitItem>(),
                                                                        (int)vt.n_items, d_head, d_done);
    if (!pdl) prof_end();
    const char* tname = st == 0 ? "k_vit_trace(fic)" : st == 2 ? "k_vit_trace(early fic)" : "k_vit_trace(msc)";
    bool two = false;
    if (pdl) {
        // (profiling: one pair of events around both kernels -- an event between them would be a stream dependency and keep the
        // trace from starting early)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_groups); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_vit_trace, (const uint2*)d_dec.as<uint2>(), (const unsigned short*)d_meet.as<unsigned short>(), dj, dg,
                              (const uint32_t*)d_prbs_w.as<uint32_t>(), d_out, (int*)nullptr, -1, (const int*)vt.d_order.as<int>(), (const int*)d_done,
                              st == 1 ? trace_host_out : (uint8_t*)nullptr, export_vec));
        prof_end();
        launches += 1;
    } else {
        // Sub-channels whose logical frames are final after the trace (DAB / data) are traced first, so that their copy to the host runs
        // under the trace of the DAB+ sub-channels and the superframe kernels behind it (ev_first_trace: recorded between the two launches).
        bool any_late = false, any_early = false;
        for (auto& j : vt.jobs_in) { if (j.late) any_late = true; else any_early = true; }
        two = any_late && any_early && ev_first_trace != nullptr;
        if (two) {
            // the two halves run next to each other: both are chains of dependent steps that leave most of the GPU idle on their own
            CK(cudaEventRecord(ev_vit, stream));
            CK(cudaStreamWaitEvent(aux_stream, ev_vit, 0));
        }
        for (int ph = 0; ph < (two ? 2 : 1); ph++) {
            cudaStream_t keep = cur;
            if (two && ph == 1) cur = aux_stream;
            prof_begin(tname);
            k_vit_trace<<<(unsigned)n_groups, 64, 0, S()>>>(d_dec.as<uint2>(), d_meet.as<unsigned short>(), dj, dg, d_prbs_w.as<uint32_t>(), d_out, d_head, two ? ph : -1,
                                                            vt.d_order.as<int>(), nullptr, st == 1 ? trace_host_out : nullptr, export_vec);
            prof_end();
            cur = keep;
            launches += 1;
            if (two && ph == 0) CK(cudaEventRecord(ev_first_trace, stream));
        }
    }
    late_on_aux = two;
    first_trace_recorded = two;
    launches += 2;
    CK(cudaGetLastError());
    return 0;
}

static VitJob make_job(const b200dab_subch_t& sc, int kind, int n_rows, int max_row, int stepinfo_off, long long out_off)
{
    VitJob j;
    memset(&j, 0, sizeof j);
    j.kind = kind;
    j.n_info = sc.bitrate * 24;
    j.steps = j.n_info + 6;
    j.n_rows = n_rows;
    j.row0 = 0;
    j.max_row = max_row;
    j.stepinfo_off = stepinfo_off;
    j.out_off = out_off;
    j.out_pitch = j.n_info / 8;
    if (kind == 0) { j.soft_base = 0; j.row_pitch = DAB_FIC_PUNCT; j.use_tdi = 0; }
    else { j.soft_base = DAB_FIC_SOFT + (long long)sc.start_cu * 64; j.row_pitch = DAB_CIF_BITS; j.use_tdi = 1; }
    return j;
}

int b200dab_ctx::fic(int n_slots, bool early)
{
    cudaStream_t stream = S();
    DevBuf& d_fic_out = early ? d_fic_out_early : this->d_fic_out;
    DevBuf& d_fib_ok = early ? d_fib_ok_early : this->d_fib_ok;
    b200dab_subch_t ficp;
    memset(&ficp, 0, sizeof ficp);
    ficp.bitrate = 32; ficp.n_seg = 2; ficp.seg_len[0] = 21; ficp.seg_pi[0] = 16; ficp.seg_len[1] = 3; ficp.seg_pi[1] = 15;
    std::vector<uint32_t> si;
    if (dabtab::build_stepinfo(ficp, si) != DAB_FIC_PUNCT) { set_error("FIC puncturing table inconsistent"); return -1; }
    const int n_rows = 4 * n_slots;
    std::vector<VitJob> jobs(1, make_job(ficp, 0, n_rows, n_rows - 1, 0, 0));
    CK(d_fic_out.ensure((size_t)(n_rows + 32) * 96));
    CK(d_fib_ok.ensure((size_t)n_rows * 3 + 64));
    vit_set = early ? 2 : -1;
    const int rc = run_viterbi(jobs, si, d_fic_out.as<uint8_t>());
    vit_set = -1;
    if (rc) return -1;
    const int n_fib = n_rows * 3;
    prof_begin(early ? "k_fib_crc(early)" : "k_fib_crc");
    k_fib_crc<<<(n_fib + 255) / 256, 256, 0, stream>>>(d_fic_out.as<uint8_t>(), n_fib, d_fib_ok.as<uint8_t>());
    prof_end();
    launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int b200dab_ctx::msc(int n_slots, const std::vector<b200dab_subch_t>& subch, std::vector<size_t>& out_off, size_t* total)
{
    const int n_cif = 4 * n_slots, n_rows = n_cif - (DAB_TDI_DEPTH - 1);
    out_off.assign(subch.size(), 0);
    *total = 0;
    if (n_rows <= 0 || subch.empty()) return 0;
    std::vector<uint32_t> si;
    std::vector<VitJob> jobs;
    size_t off = 0;
    for (size_t i = 0; i < subch.size(); i++) {
        const int si_off = (int)si.size();
        int np = -1;
        if (const std::vector<uint32_t>* hit = stepinfo_for(subch[i])) { si.insert(si.end(), hit->begin(), hit->end()); np = (int)(hit->back() >> 4); }
        if (np < 0 || np > subch[i].size_cu * 64 || subch[i].start_cu < 0 || subch[i].start_cu + subch[i].size_cu > 864) {
            set_error("sub-channel protection profile inconsistent with its size");
            return -1;
        }
        out_off[i] = off;
        jobs.push_back(make_job(subch[i], 1, n_rows, n_cif - 1, si_off, (long long)off));
        jobs.back().late = subch[i].kind == 1 ? 1 : 0;       // DAB+: its frames are final only after the RS correction
        off += (size_t)((n_rows + 31) / 32 * 32) * (size_t)(subch[i].bitrate * 3);
    }
    *total = off;
    CK(d_msc_out.ensure(off + 256));
    frames_exported = export_frames && off < (1ull << 32);     // (patch offsets are 32 bit)
    if (frames_exported) {
        if (off + 64 > h_msc.cap) CK(cudaDeviceSynchronize());    // growing the buffer frees it: nothing may still be writing into it
        CK(h_msc.ensure(off + 64));
        trace_host_out = h_msc.as<uint8_t>();
    }
    const int rc = run_viterbi(jobs, si, d_msc_out.as<uint8_t>());
    trace_host_out = nullptr;
    if (rc) frames_exported = false;
    return rc;
}

int b200dab_ctx::dabplus(uint8_t* d_fr, const std::vector<DabPlusJob>& jobs, int total_sf, int total_rows)
{
    if (jobs.empty() || total_sf <= 0) return 0;
    CK(d_dpjobs.ensure(jobs.size() * sizeof(DabPlusJob) + 16));
    CK(h_dpjobs.ensure(jobs.size() * sizeof(DabPlusJob) + 16));
    memcpy(h_dpjobs.p, jobs.data(), jobs.size() * sizeof(DabPlusJob));
    if (fetch(d_dpjobs.p, h_dpjobs.p, jobs.size() * sizeof(DabPlusJob))) { set_error("k_fetch launch failed"); return -1; }
    int max_sf = 0;
    for (auto& j : jobs) { max_sf = std::max(max_sf, j.max_sf); if (j.s > 24 || j.s < 1 || j.pitch != 24 * j.s) { set_error("DAB+ sub-channel bitrate out of range"); return -1; } }
    const unsigned nj = (unsigned)jobs.size();
    rs_patch_cap = 0;
    if (rs_patch_mode) for (auto& j : jobs) rs_patch_cap += (size_t)j.max_sf * j.s * 5;
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    static_assert(sizeof(DabPlusTotals) == 32, "the access-unit records follow the totals at byte 32");
    dp.aus = sizeof(DabPlusTotals);
    dp.sf = al16(dp.aus + (size_t)total_sf * 6 * sizeof(AuRecord));
    dp.flags = al16(dp.sf + (size_t)total_sf * sizeof(SfInfo));
    dp.res = al16(dp.flags + (size_t)total_rows);
    dp.patch = al16(dp.res + (size_t)nj * 4);
    dp.end = dp.patch + rs_patch_cap * sizeof(RsPatch);
    CK(d_dpres.ensure(dp.end + 64));
    CK(d_naus.ensure((size_t)total_sf * 4 + 64));
    DabPlusTotals* d_tot = dpp<DabPlusTotals>(0);
    prof_begin("k_fire_align");
    k_fire_align<<<nj, FIRE_THREADS, 0, S()>>>(d_fr, d_dpjobs.as<DabPlusJob>(), dpp<uint8_t>(dp.flags), dpp<SfInfo>(dp.sf), dpp<int>(dp.res), d_tot);
    prof_end();
    // RS correction, fire code, access-unit table and access-unit CRCs of a superframe by one warp, in one launch
    prof_begin("k_sf_decode");
    k_sf_decode<<<dim3((max_sf + SF_PER_CTA - 1) / SF_PER_CTA, nj), 32 * SF_PER_CTA, 0, S()>>>(d_fr, d_dpjobs.as<DabPlusJob>(), dpp<SfInfo>(dp.sf), d_gf.as<SfTables>(),
                                                                                              d_naus.as<int>(), d_tot,
                                                                                              rs_patch_mode ? dpp<RsPatch>(dp.patch) : nullptr, (int)std::min<size_t>(rs_patch_cap, 1u << 30));
    prof_end();
    if (ev_rs) CK(cudaEventRecord(ev_rs, S()));        // the logical frames are final from here on: the caller starts their copy
    // access-unit list and RS statistics in the layout of the result, put together on the device
    prof_begin("k_au_list");
    k_au_list<<<(total_sf + 255) / 256, 256, 0, S()>>>(d_dpjobs.as<DabPlusJob>(), (int)nj, dpp<SfInfo>(dp.sf), d_naus.as<int>(), total_sf, dpp<AuRecord>(dp.aus), d_tot);
    prof_end();
    launches += 3;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ FIG 0/1 + 0/2 (sub-channel discovery)
static int parse_subch_impl(const uint8_t* fibs, const uint8_t* ok, int n_fibs, b200dab_subch_t* out, int max_out)
{
    struct Ent { bool have = false; b200dab_subch_t sc; bool dabplus = false; bool audio = false; };
    std::vector<Ent> tab(64);
    for (int f = 0; f < n_fibs; f++) {
        if (ok && !ok[f]) continue;
        const uint8_t* p = fibs + (size_t)f * 32;
        int i = 0;
        while (i < 30) {
            const uint8_t h = p[i];
            if (h == 0xFF) break;
            const int type = h >> 5, len = h & 31;
            if (len == 0 || i + 1 + len > 30) break;
            const uint8_t* d = p + i + 1;
            if (type == 0) {
                const int ext = d[0] & 31, pd = (d[0] >> 5) & 1, cn = d[0] >> 7;
                int k = 1;
                // C/N = 1 on the MCI extensions announces the NEXT multiplex configuration (EN 300 401 sec. 6.4, sent for up to 6 s
                // ahead of a reconfiguration): it must not overwrite the organisation the MSC is decoded with now
                if (cn && (ext == 1 || ext == 2)) { i += 1 + len; continue; }
                if (ext == 1) {
                    while (k + 3 <= len) {
                        const int id = d[k] >> 2, start = ((d[k] & 3) << 8) | d[k + 1];
                        b200dab_subch_t sc;
                        memset(&sc, 0, sizeof sc);
                        sc.subch_id = id; sc.start_cu = start;
                        int okp = -1;
                        if (d[k + 2] & 0x80) {
                            if (k + 4 > len) break;
                            const int opt = (d[k + 2] >> 4) & 7, lvl = (d[k + 2] >> 2) & 3, size = ((d[k + 2] & 3) << 8) | d[k + 3];
                            int bitrate = 0;
                            if (opt == 0) { static const int div[4] = {12, 8, 6, 4}; bitrate = size / div[lvl] * 8; }
                            else if (opt == 1) { static const int div[4] = {27, 21, 18, 15}; bitrate = size / div[lvl] * 32; }
                            if (bitrate > 0) okp = dabtab::profile_eep(bitrate, lvl + 1, opt, &sc);
                            if (okp == 0 && sc.size_cu != size) okp = -1;
                            k += 4;
                        } else {
                            okp = dabtab::profile_uep(d[k + 2] & 0x3F, &sc);
                            k += 3;
                        }
                        if (okp == 0) { sc.subch_id = id; sc.start_cu = start; tab[id].have = true; bool dp = tab[id].dabplus; tab[id].sc = sc; tab[id].dabplus = dp; }
                    }
                } else if (ext == 2) {
                    while (k < len) {
                        const int sidlen = pd ? 4 : 2;
                        if (k + sidlen + 1 > len) break;
                        k += sidlen;
                        const int ncomp = d[k] & 0x0F;
                        k++;
                        for (int c = 0; c < ncomp && k + 2 <= len; c++, k += 2) {
                            const int tmid = d[k] >> 6;
                            if (tmid == 0) {
                                const int ascty = d[k] & 0x3F, id = d[k + 1] >> 2;
                                tab[id].audio = true;
                                tab[id].dabplus = (ascty == 63);
                            }
                        }
                    }
                }
            }
            i += 1 + len;
        }
    }
    int n = 0;
    for (int id = 0; id < 64 && n < max_out; id++)
        if (tab[id].have) { out[n] = tab[id].sc; out[n].kind = tab[id].dabplus ? 1 : 0; n++; }
    std::sort(out, out + n, [](const b200dab_subch_t& a, const b200dab_subch_t& b) { return a.start_cu < b.start_cu; });
    return n;
}

// ================================================================================================ C ABI
extern "C" {

int b200dab_version(void) { return 0x000100; }

int b200dab_create(b200dab_ctx** out, int device)
{
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return -2;      // no CPU fallback
    if (device < 0 || device >= ndev) return -3;
    if (cudaSetDevice(device) != cudaSuccess) return -4;
    b200dab_ctx* ctx = new b200dab_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -5; }
    if (const char* e = getenv("B200DAB_VIT_CTAS")) ctx->vit_ctas_per_sm = std::max(1, atoi(e));
    if (const char* e = getenv("B200DAB_VIT_SMEMCAP")) ctx->vit_smem_cap = atoi(e) != 0;
    if (const char* e = getenv("B200DAB_TRACE_PDL")) ctx->trace_pdl = atoi(e) != 0;
    if (const char* e = getenv("B200DAB_VIT_SMEM_SLACK")) ctx->vit_smem_slack = std::max(0, atoi(e));
    if (const char* e = getenv("B200DAB_FRAME_EXPORT")) ctx->frame_export = atoi(e) != 0;
    if (const char* e = getenv("B200DAB_EXPORT_VEC")) ctx->export_vec = atoi(e) == 16 ? 16 : 8;
    if (const char* e = getenv("B200DAB_RS_PATCH_PREFIX")) ctx->rs_patch_prefix = (size_t)std::max(0, std::min(RS_PATCH_PREFIX, atoi(e)));
    if (const char* e = getenv("B200DAB_STEPINFO_CACHE")) ctx->stepinfo_cached = atoi(e) != 0;
    cudaFuncSetAttribute(k_viterbi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (const char* e = getenv("B200DAB_VIT4_MIN_STEPS")) ctx->vit4_min_steps = atoi(e);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (ctx->d_vit_counter.ensure(64) != cudaSuccess || cudaMemset(ctx->d_vit_counter.p, 0, 64) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return -5; }
    // The FIC stream has priority: its few small CTAs are placed as soon as slots free up, ahead of the thousands of pending CTAs of
    // the demodulation / MSC kernels they run next to.
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&ctx->fic_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return -5; }
    cudaEventCreateWithFlags(&ctx->ev_early_demod, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_early_fic, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_early_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_frames, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_demod, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_fic_stream, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_vit, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming);
    if (ctx->init_tables()) { fprintf(stderr, "b200dab: %s\n", ctx->err.c_str()); cudaStreamDestroy(ctx->stream); delete ctx; return -6; }
    *out = ctx;
    return 0;
}

void b200dab_destroy(b200dab_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* db[] = {&ctx->d_tw, &ctx->d_fdi_pos, &ctx->d_prs_q, &ctx->d_prbs_w, &ctx->d_gf, &ctx->d_naus, &ctx->d_cls_bin, &ctx->d_raw, &ctx->d_P, &ctx->d_S, &ctx->d_cand_raw,
                    &ctx->d_cand, &ctx->d_ncand, &ctx->d_fine, &ctx->d_frames, &ctx->d_chain, &ctx->d_frame_cir, &ctx->d_soft, &ctx->d_symstat, &ctx->d_jobs, &ctx->d_groups,
                    &ctx->d_stepinfo, &ctx->vit_tables[0].d_jobs, &ctx->vit_tables[0].d_groups, &ctx->vit_tables[0].d_stepinfo, &ctx->vit_tables[1].d_jobs,
                    &ctx->vit_tables[1].d_groups, &ctx->vit_tables[1].d_stepinfo, &ctx->vit_tables[2].d_jobs, &ctx->vit_tables[2].d_groups, &ctx->vit_tables[2].d_stepinfo, &ctx->vit_tables[0].d_vin, &ctx->vit_tables[0].d_dec, &ctx->vit_tables[0].d_meet, &ctx->vit_tables[1].d_vin, &ctx->vit_tables[1].d_dec, &ctx->vit_tables[1].d_meet, &ctx->vit_tables[2].d_vin, &ctx->vit_tables[2].d_dec, &ctx->vit_tables[2].d_meet, &ctx->vit_tables[0].d_items, &ctx->vit_tables[1].d_items, &ctx->vit_tables[2].d_items, &ctx->vit_tables[0].d_done, &ctx->vit_tables[1].d_done, &ctx->vit_tables[2].d_done, &ctx->vit_tables[0].d_order, &ctx->vit_tables[1].d_order, &ctx->vit_tables[2].d_order, &ctx->d_items, &ctx->d_vit_counter, &ctx->d_fic_out_early, &ctx->d_fib_ok_early, &ctx->d_vin, &ctx->d_dec, &ctx->d_meet, &ctx->d_fic_out, &ctx->d_fib_ok, &ctx->d_msc_out, &ctx->d_dpjobs,
                    &ctx->d_fold, &ctx->d_norm, &ctx->d_sign, &ctx->d_windows, &ctx->d_tii, &ctx->d_tmp_a, &ctx->d_tmp_b, &ctx->d_spec, &ctx->d_dpres};
    for (auto* b : db) b->release();
    HostBuf* hb[] = {&ctx->h_fine, &ctx->h_frames, &ctx->h_chain, &ctx->h_frame_cir, &ctx->h_fibs, &ctx->h_fib_ok, &ctx->h_msc, &ctx->h_tii, &ctx->h_symstat, &ctx->h_misc, &ctx->h_fold, &ctx->h_norm,
                     &ctx->h_spec, &ctx->h_jobs[0], &ctx->h_jobs[1], &ctx->h_jobs[2], &ctx->h_groups[0], &ctx->h_groups[1], &ctx->h_groups[2], &ctx->h_stepinfo[0], &ctx->h_stepinfo[1], &ctx->h_stepinfo[2], &ctx->h_items[0], &ctx->h_items[1], &ctx->h_items[2], &ctx->h_order[0], &ctx->h_order[1], &ctx->h_order[2], &ctx->h_fibs_early, &ctx->h_fib_ok_early, &ctx->h_dpjobs, &ctx->h_sign, &ctx->h_windows, &ctx->h_dpres};
    cudaStreamSynchronize(ctx->fic_stream); if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream); if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (auto* b : hb) b->release();
    {
        std::lock_guard<std::mutex> lk(g_upload_mu);
        for (auto e : ctx->copy_events) { if (g_upload_last[ctx->device & 63] == e) g_upload_last[ctx->device & 63] = nullptr; cudaEventDestroy(e); }
    }
    if (ctx->ev_msc) { cudaEventDestroy(ctx->ev_msc); cudaEventDestroy(ctx->ev_d2h); cudaEventDestroy(ctx->ev_rs); cudaEventDestroy(ctx->ev_tii); }
    if (ctx->ev_fic) cudaEventDestroy(ctx->ev_fic);
    if (ctx->fic_stream) { cudaEventDestroy(ctx->ev_early_demod); cudaEventDestroy(ctx->ev_early_done); cudaEventDestroy(ctx->ev_early_fic); cudaEventDestroy(ctx->ev_demod); cudaEventDestroy(ctx->ev_fic_stream); cudaStreamDestroy(ctx->fic_stream); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_vit) { cudaEventDestroy(ctx->ev_vit); cudaEventDestroy(ctx->ev_aux); }
    if (ctx->ev_chain) { cudaEventDestroy(ctx->ev_chain); cudaEventDestroy(ctx->ev_frames); }
    for (auto& set : ctx->prof_sets) for (auto& e : set) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* b200dab_host_alloc(b200dab_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
void b200dab_host_free(b200dab_ctx* ctx, void* p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFreeHost(p);
}

const char* b200dab_last_error(b200dab_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void* b200dab_stream(b200dab_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int b200dab_profile_eep(int bitrate, int level, int option, b200dab_subch_t* out)
{
    if (!out) return -1;
    return dabtab::profile_eep(bitrate, level, option, out);
}
int b200dab_profile_uep(int idx, b200dab_subch_t* out)
{
    if (!out) return -1;
    return dabtab::profile_uep(idx, out);
}
int b200dab_parse_subch(const uint8_t* fibs, const uint8_t* fib_ok, int n_fibs, b200dab_subch_t* out, int max_out)
{
    if (!fibs || !out) return -1;
    return parse_subch_impl(fibs, fib_ok, n_fibs, out, max_out);
}

int b200dab_set_superframe_phase(b200dab_ctx* ctx, const int32_t* phase, int n)
{
    if (!ctx || n < 0 || (n > 0 && !phase)) return -1;
    ctx->forced_phase.assign(phase, phase + n);
    return 0;
}

int b200dab_set_profiling(b200dab_ctx* ctx, int on)
{
    if (!ctx) return -1;
    ctx->profile = on != 0;
    return 0;
}

// per-kernel device times of the last b200dab_decode (CUDA events on the context stream); names are ';'-separated
int b200dab_get_kernel_times(b200dab_ctx* ctx, char* names, int names_cap, float* ms, int max_n)
{
    return b200dab_get_kernel_times_back(ctx, 0, names, names_cap, ms, max_n);
}

int b200dab_get_kernel_times_back(b200dab_ctx* ctx, int back, char* names, int names_cap, float* ms, int max_n)
{
    return b200dab_get_kernel_timeline(ctx, back, names, names_cap, nullptr, ms, max_n);
}

int b200dab_get_kernel_timeline(b200dab_ctx* ctx, int back, char* names, int names_cap, float* start_ms, float* ms, int max_n)
{
    if (!ctx || back < 0 || back >= b200dab_ctx::PROF_RING) return -1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    const int slot = (ctx->prof_slot - back + 2 * b200dab_ctx::PROF_RING) % b200dab_ctx::PROF_RING;
    const std::vector<b200dab_ctx::ProfEv>& prof = ctx->prof_sets[slot];
    int n = 0;
    std::string all;
    for (size_t i = 0; i < ctx->prof_counts[slot] && n < max_n; i++, n++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, prof[i].a, prof[i].b) != cudaSuccess) t = -1.f;
        if (ms) ms[n] = t;
        if (start_ms && cudaEventElapsedTime(&start_ms[n], prof[0].a, prof[i].a) != cudaSuccess) start_ms[n] = -1.f;
        if (!all.empty()) all += ";";
        all += prof[i].name;
    }
    if (names && names_cap > 0) { strncpy(names, all.c_str(), (size_t)names_cap - 1); names[names_cap - 1] = 0; }
    return n;
}

struct b200dab_pad { pad::Decoder d; std::vector<uint8_t> tmp; packet::Assembler pk; };
b200dab_pad* b200dab_pad_create(void) { return new b200dab_pad(); }
void b200dab_pad_destroy(b200dab_pad* p) { delete p; }
void b200dab_pad_reset(b200dab_pad* p) { if (p) p->d.reset(); }
int b200dab_pad_feed_au(b200dab_pad* p, const uint8_t* au, int len, b200dab_pad_cb cb, void* user)
{
    const uint8_t* pb; int pn;
    if (!p || !au || !pad::dabplus_pad(au, len, &pb, &pn)) return 0;
    p->d.feed(pb, pn, [&](const uint8_t* d, int n) { if (cb) cb(user, 0, d, n); }, [&](int app, const uint8_t* d, int n) { if (cb) cb(user, app, d, n); });
    return 1;
}

int b200dab_pad_feed_packets(b200dab_pad* p, int addr, const uint8_t* f, int len, b200dab_pad_cb cb, void* user)
{
    if (!p || !f) return 0;
    if (p->pk.addr != addr) p->pk.reset(addr);
    p->pk.frame(f, len, [&](const uint8_t* d, int n) { if (cb) cb(user, -1, d, n); });
    return 1;
}

int b200dab_pad_feed_mp2(b200dab_pad* p, const uint8_t* f, int len, b200dab_pad_cb cb, void* user)
{
    if (!p || !f || !pad::mp2_pad(f, len, p->tmp)) return 0;
    p->d.feed(p->tmp.data(), (int)p->tmp.size(), [&](const uint8_t* d, int n) { if (cb) cb(user, 0, d, n); }, [&](int app, const uint8_t* d, int n) { if (cb) cb(user, app, d, n); });
    return 1;
}

int b200dab_viterbi_decision_index(int state, int backward)
{
    if (state < 0 || state > 63) return -1;
    return (int)(backward ? vit_dec_index_bwd((uint32_t)state) : vit_dec_index_fwd((uint32_t)state));
}

int b200dab_viterbi4_decision_index(int state, int backward)
{
    if (state < 0 || state > 63) return -1;
    return (int)(backward ? vit4_dec_index_bwd((uint32_t)state) : vit4_dec_index_fwd((uint32_t)state));
}

int b200dab_set_viterbi_split(b200dab_ctx* ctx, int min_steps_four_lane)
{
    if (!ctx) return -1;
    ctx->vit4_min_steps = min_steps_four_lane;
    return 0;
}

int b200dab_spectrum_layout(uint16_t* out, int n)
{
    std::vector<uint16_t> pos;
    if (!out || n < DAB_NFFT + DAB_NCARR || build_spectrum_layout(pos)) return -1;
    memcpy(out, pos.data(), pos.size() * sizeof(uint16_t));
    return (int)pos.size();
}

int b200dab_parse_ensemble(const uint8_t* fibs, const uint8_t* fib_ok, int n_fibs, b200dab_ensemble_t* out)
{
    if (!fibs || !out) return -1;
    fig::Database db;
    memset(out, 0, sizeof *out);
    out->first_cif_count = -1;
    for (int f = 0; f < n_fibs; f++) {
        if (fib_ok && !fib_ok[f]) continue;
        db.parse_fib(fibs + (size_t)f * 32);
        if (out->first_cif_count < 0 && db.cif_count >= 0) out->first_cif_count = db.cif_count;
    }
    out->ueid = ((uint32_t)(db.have_ecc ? db.ecc : 0) << 16) | db.eid;
    out->lto = db.lto; out->int_table = db.int_table; out->alarm = db.alarm;
    if (db.label.valid) memcpy(out->label, db.label.str, 16);
    for (auto& sv : db.services) {
        if (sv.comps.empty() || out->n_services >= 64) continue;
        b200dab_service_t& o = out->services[out->n_services++];
        o.sid = sv.sid;
        if (sv.label.valid) memcpy(o.label, sv.label.str, 16);
        for (auto& c : sv.comps) if (c.tmid == 0 && o.n_components < 4) { o.subch_id[o.n_components] = (uint8_t)c.subch_id; o.ascty[o.n_components] = c.ascty_dscty; o.n_components++; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ whole path
static int decode_impl(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, int mem, const b200dab_subch_t* subch_in, int n_subch_in,
                       uint32_t flags, b200dab_result_t* res);

int b200dab_decode(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, int mem, const b200dab_subch_t* subch_in, int n_subch_in,
                   uint32_t flags, b200dab_result_t* res)
{
    const int rc = decode_impl(ctx, iq, n, fmt, mem, subch_in, n_subch_in, flags, res);
    if (rc != 0 && ctx) {
        // An error exit may leave copies and kernels in flight on any of the context's streams: they read the caller's buffer and
        // the context's pinned staging, which the caller / the next call are free to reuse once this returns.
        const std::string keep = ctx->err;
        cudaStream_t all[] = {ctx->stream, ctx->fic_stream, ctx->aux_stream, ctx->copy_stream};
        for (cudaStream_t st : all) if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
        ctx->err = keep;
        ctx->cur = nullptr;
    }
    return rc;
}

static int decode_impl(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, int mem, const b200dab_subch_t* subch_in, int n_subch_in,
                       uint32_t flags, b200dab_result_t* res)
{
    if (!ctx || !iq || !res || fmt < 0 || fmt > 2) return -1;
    cudaSetDevice(ctx->device);
    memset(res, 0, sizeof *res);
    ctx->launches = 0;
    ctx->prof_next();
    ctx->t_trace0 = now_ms();
    ctx->r_aus.clear(); ctx->r_tii.clear(); ctx->r_subch.clear(); ctx->r_frames.clear(); ctx->r_rows.clear();
    ctx->r_sf.clear(); ctx->r_sf_flags.clear(); ctx->r_sf_phase.clear();
    const std::vector<int32_t> forced_phase = std::move(ctx->forced_phase);     // one-shot
    ctx->forced_phase.clear();
    const bool want_stat = (flags & B200DAB_OPT_SYMSTAT) != 0;
    const bool want_tii = (flags & B200DAB_OPT_TII) != 0;
    bool deferred_tii = false;
    const int max_slots = (int)(n / DAB_T_FRAME) + 2;
    ctx->fe_frames = 0; ctx->fe_candidates = 0; ctx->fe_ranges = 0;
    // early sub-channel discovery pays when the sub-channels have to be found and there is an MSC to decode behind a long front end
    ctx->early_launched = false;
    ctx->cur = nullptr;
    ctx->early_frames = (!(subch_in && n_subch_in > 0) && !(flags & B200DAB_OPT_NO_MSC)) ? EARLY_FIC_FRAMES : 0;
    if (ctx->prepare_slots(max_slots, want_stat, want_tii)) return -1;
    // ---- front end: sync + OFDM demodulation.  A host buffer is uploaded in chunks on a second stream and each chunk is
    //      synchronised and demodulated while the next one is still crossing PCIe; a device buffer is done in one range.
    const void* d_iq = iq;
    const uint64_t chunk = (uint64_t)FRONT_CHUNK_FRAMES * DAB_T_FRAME;
    const uint64_t lap = (uint64_t)DAB_T_FRAME + 2 * DAB_T_NULL + SYNC_FRAME_MARGIN + 4 * SYNC_WIN * SYNC_BLK;
    if (mem == B200DAB_MEM_HOST && n > 2 * chunk) {
        const size_t bps = fmt_bytes(fmt);
        CKC(ctx->d_raw.ensure((size_t)n * bps + 64));
        d_iq = ctx->d_raw.p;
        if (!ctx->copy_stream) CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        const int n_chunks = (int)((n + chunk - 1) / chunk);
        while ((int)ctx->copy_events.size() < n_chunks) { cudaEvent_t e; CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->copy_events.push_back(e); }
        {
            std::lock_guard<std::mutex> lk(g_upload_mu);
            cudaEvent_t& last = g_upload_last[ctx->device & 63];
            if (last) CKC(cudaStreamWaitEvent(ctx->copy_stream, last, 0));
            for (int c = 0; c < n_chunks; c++) {
                const uint64_t a0 = (uint64_t)c * chunk, a1 = std::min<uint64_t>(n, a0 + chunk);
                CKC(cudaMemcpyAsync((uint8_t*)ctx->d_raw.p + a0 * bps, (const uint8_t*)iq + a0 * bps, (size_t)(a1 - a0) * bps, cudaMemcpyHostToDevice, ctx->copy_stream));
                CKC(cudaEventRecord(ctx->copy_events[c], ctx->copy_stream));
            }
            last = ctx->copy_events[n_chunks - 1];
        }
        static const bool trace = getenv("B200DAB_TRACE") != nullptr;
        auto t_start = std::chrono::steady_clock::now();
        for (int c = 0; c < n_chunks; c++) {
            if (trace) fprintf(stderr, "[trace] chunk %d begin %.3f ms\n", c, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
            const uint64_t a1 = std::min<uint64_t>(n, (uint64_t)(c + 1) * chunk);
            uint64_t a0 = (uint64_t)c * chunk;
            a0 = a0 > lap ? a0 - lap : 0;
            a0 &= ~(uint64_t)(SYNC_BLK * 8 - 1);                    // keep 128-bit loads of the block-power kernel aligned
            CKC(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[c], 0));
            if (ctx->front_end(d_iq, fmt, a0, a1 - a0, want_stat, want_tii, max_slots)) return -1;
        }
        if (trace) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[trace] front end done %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count()); }
    } else {
        if (ctx->upload_raw(iq, n, fmt, mem, &d_iq)) return -1;
        ctx->defer_tii_fold = want_tii;
        const int rc = ctx->front_end(d_iq, fmt, 0, n, want_stat, want_tii, max_slots);
        ctx->defer_tii_fold = false;
        if (rc) return -1;
        deferred_tii = want_tii;
    }
    TRACE("front end submitted");
    // ---- the frame list was put together on the device (last CTA of every k_fine_sync launch): it comes back on the copy stream,
    //      behind the last chaining step, while the GPU is busy with the demodulation that was queued behind it
    if (!ctx->copy_stream) CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (ctx->fe_ranges > 0) {
        CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chain, 0));
        CKC(cudaMemcpyAsync(ctx->h_chain.p, ctx->d_chain.p, 32, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CKC(cudaMemcpyAsync(ctx->h_frames.p, ctx->d_frames.p, (size_t)max_slots * sizeof(FrameInfo), cudaMemcpyDeviceToHost, ctx->copy_stream));
        CKC(cudaMemcpyAsync(ctx->h_frame_cir.p, ctx->d_frame_cir.p, (size_t)max_slots * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CKC(cudaEventRecord(ctx->ev_frames, ctx->copy_stream));
        CKC(cudaEventSynchronize(ctx->ev_frames));
        TRACE("frame list on the host");
        const ChainState* cs = ctx->h_chain.as<ChainState>();
        ctx->fe_frames = (size_t)std::min(cs->n_frames, max_slots);
        ctx->fe_candidates = cs->n_candidates;
    }
    res->n_candidates = ctx->fe_candidates;
    const int n_frames = (int)ctx->fe_frames;
    const FrameInfo* frames = ctx->h_frames.as<FrameInfo>();
    const int n_slots = n_frames ? frames[n_frames - 1].slot + 1 : 0;
    res->n_frames = n_frames; res->n_slots = n_slots;
    ctx->r_tf.assign(n_slots, -1); ctx->r_cfo.assign(n_slots, 0.f); ctx->r_cir.assign(n_slots, 0.f);
    for (int i = 0; i < n_frames; i++) { ctx->r_tf[frames[i].slot] = frames[i].t_f; ctx->r_cfo[frames[i].slot] = frames[i].cfo_hz; ctx->r_cir[frames[i].slot] = ctx->h_frame_cir.as<float>()[i]; }
    res->t_f = ctx->r_tf.data(); res->cfo_hz = ctx->r_cfo.data(); res->cir_ratio = ctx->r_cir.data();
    // the early FIC slice was queued blindly: it is usable if its frames exist and none of them was lost
    if (ctx->early_launched && !(n_frames > ctx->early_frames && frames[ctx->early_frames - 1].slot == ctx->early_frames - 1)) ctx->early_launched = false;
    if (n_frames == 0) { res->gpu_launches = ctx->launches; return 0; }
    if (flags & B200DAB_OPT_SPECTRUM) {
        CKC(ctx->d_spec.ensure((size_t)n_slots * 2 * DAB_NFFT * 4));
        CKC(ctx->h_spec.ensure((size_t)n_slots * 2 * DAB_NFFT * 4));
        CKC(cudaMemsetAsync(ctx->d_spec.p, 0, (size_t)n_slots * 2 * DAB_NFFT * 4, ctx->stream));
        const FrameInfo* dfr = ctx->d_frames.as<FrameInfo>();
        const float2* tw = ctx->d_tw.as<float2>();
        const int64_t nn = (int64_t)n;
        ctx->prof_begin("k_spectrum");
        if (fmt == B200DAB_FMT_U8) k_spectrum<B200DAB_FMT_U8><<<n_frames, 256, 0, ctx->stream>>>(d_iq, nn, dfr, tw, ctx->d_spec.as<float>());
        else if (fmt == B200DAB_FMT_S16) k_spectrum<B200DAB_FMT_S16><<<n_frames, 256, 0, ctx->stream>>>(d_iq, nn, dfr, tw, ctx->d_spec.as<float>());
        else k_spectrum<B200DAB_FMT_F32><<<n_frames, 256, 0, ctx->stream>>>(d_iq, nn, dfr, tw, ctx->d_spec.as<float>());
        ctx->prof_end();
        ctx->launches += 1;
        CKC(cudaGetLastError());
        CKC(cudaMemcpyAsync(ctx->h_spec.p, ctx->d_spec.p, (size_t)n_slots * 2 * DAB_NFFT * 4, cudaMemcpyDeviceToHost, ctx->stream));
        res->spectrum = ctx->h_spec.as<float>();
    }
    // ---- FIC: its own stream behind the demodulation, so that it runs next to the MSC chain once the sub-channels are known
    if (!ctx->ev_fic) CKC(cudaEventCreateWithFlags(&ctx->ev_fic, cudaEventDisableTiming));
    CKC(cudaEventRecord(ctx->ev_demod, ctx->stream));
    CKC(cudaStreamWaitEvent(ctx->fic_stream, ctx->ev_demod, 0));
    ctx->cur = ctx->fic_stream;
    int fic_rc = 0;
    if (deferred_tii) fic_rc = ctx->tii_fold(d_iq, n, fmt, -1, 0, 1 << 30, n_frames);     // under the MSC gather instead of in front of it
    if (!fic_rc) fic_rc = ctx->fic(n_slots);
    ctx->cur = nullptr;
    if (fic_rc) return -1;
    const size_t fib_bytes = (size_t)n_slots * 12 * 32;
    CKC(ctx->h_fibs.ensure(fib_bytes));
    CKC(ctx->h_fib_ok.ensure((size_t)n_slots * 12));
    CKC(cudaMemcpyAsync(ctx->h_fibs.p, ctx->d_fic_out.p, fib_bytes, cudaMemcpyDeviceToHost, ctx->fic_stream));
    CKC(cudaMemcpyAsync(ctx->h_fib_ok.p, ctx->d_fib_ok.p, (size_t)n_slots * 12, cudaMemcpyDeviceToHost, ctx->fic_stream));
    CKC(cudaEventRecord(ctx->ev_fic, ctx->fic_stream));
    TRACE("fic submitted");
    res->fibs = ctx->h_fibs.as<uint8_t>(); res->fib_ok = ctx->h_fib_ok.as<uint8_t>();
    // ---- sub-channel organisation: given by the caller or discovered from FIG 0/1, 0/2, which repeat every few frames: growing
    //      prefixes of the FIC (8, 16, 32, ... frames) are parsed until the table stops changing.  The first two prefixes come
    //      from the early FIC slice, which is through long before the whole FIC is; only if they disagree the host waits for it.
    bool fic_synced = false;
    auto subch_from = [&](const uint8_t* fibs, const uint8_t* ok, int have_slots, bool may_wait, bool* stable) -> int {
        ctx->r_subch.resize(B200DAB_MAX_SUBCH);
        std::vector<b200dab_subch_t> prev;
        int ns = 0;
        *stable = false;
        for (int upto = std::min(have_slots, 8); ; upto = std::min(have_slots, upto * 2)) {
            ns = parse_subch_impl(fibs, ok, upto * 12, ctx->r_subch.data(), B200DAB_MAX_SUBCH);
            const bool same = (int)prev.size() == ns && (ns == 0 || !memcmp(prev.data(), ctx->r_subch.data(), (size_t)ns * sizeof(b200dab_subch_t)));
            if (same && ns > 0) { *stable = true; break; }
            if (upto >= have_slots) { *stable = may_wait; break; }       // the whole FIC has been looked at: that is the answer
            prev.assign(ctx->r_subch.begin(), ctx->r_subch.begin() + ns);
        }
        return ns;
    };
    if (subch_in && n_subch_in > 0) ctx->r_subch.assign(subch_in, subch_in + n_subch_in);
    else {
        bool stable = false;
        int ns = 0;
        if (ctx->early_launched) {
            CKC(cudaEventSynchronize(ctx->ev_early_fic));
            TRACE("early fic synced");
            ns = subch_from(ctx->h_fibs_early.as<uint8_t>(), ctx->h_fib_ok_early.as<uint8_t>(), ctx->early_frames, false, &stable);
        }
        if (!stable) {
            CKC(cudaEventSynchronize(ctx->ev_fic));
            fic_synced = true;
            TRACE("fic synced");
            ns = subch_from(res->fibs, res->fib_ok, n_slots, true, &stable);
        }
        ctx->r_subch.resize(ns);
    }
    if (!(subch_in && n_subch_in > 0)) ctx->drop_undecodable(ctx->r_subch);
    res->n_subch = (int)ctx->r_subch.size();
    res->subch = ctx->r_subch.data();
    TRACE("sub-channels known");
    // the copy stream carries every larger device-to-host copy, each behind the event of the kernel that finishes its source
    if (!ctx->copy_stream) CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->ev_msc) {
        CKC(cudaEventCreateWithFlags(&ctx->ev_msc, cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&ctx->ev_d2h, cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ctx->ev_rs, cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&ctx->ev_tii, cudaEventDisableTiming));
    }
    bool copies_pending = false;
    int n_win = 0;
    std::vector<size_t> out_off;
    size_t msc_bytes = 0;
    const int n_rows = 4 * n_slots - (DAB_TDI_DEPTH - 1);
    std::vector<DabPlusJob> dpjobs;
    std::vector<int> dp_subch;
    int total_sf = 0, total_rows = 0;
    auto tii_block = [&]() -> int {
    // ---- TII detection needs to know which nulls carry TII: CIF counter (FIG 0/0) = 0 mod 8
    if (want_tii) {
        ctx->r_sign.assign(n_slots, 0);
        std::vector<int32_t>& sign = ctx->r_sign;
        for (int s = 0; s < n_slots; s++) {
            if (ctx->r_tf[s] < 0) continue;
            const uint8_t* fib = res->fibs + (size_t)s * 384;          // FIB 0 of the first CIF carries FIG 0/0 first
            if (res->fib_ok[(size_t)s * 12] && (fib[0] >> 5) == 0 && (fib[1] & 31) == 0 && (fib[0] & 31) >= 5) {
                const int cif = ((fib[4] & 0x1F) * 250 + fib[5]);
                sign[s] = (cif % 8 == 0) ? 1 : (cif % 8 == 4 ? -1 : 0);
            }
        }
        // frames whose FIG 0/0 was lost: alternate from a decoded neighbour
        for (int s = 1; s < n_slots; s++) if (!sign[s] && sign[s - 1] && ctx->r_tf[s] >= 0) sign[s] = -sign[s - 1];
        for (int s = n_slots - 2; s >= 0; s--) if (!sign[s] && sign[s + 1] && ctx->r_tf[s] >= 0) sign[s] = -sign[s + 1];
        std::vector<int2> wins;
        for (int w = 0; w + TII_NOTIFY_FRAMES <= n_slots; w += TII_NOTIFY_FRAMES) wins.push_back(make_int2(w, w + TII_NOTIFY_FRAMES));
        if (wins.empty()) wins.push_back(make_int2(0, n_slots & ~1));
        n_win = (int)wins.size();
        CKC(ctx->d_sign.ensure((size_t)n_slots * 4));
        CKC(ctx->d_windows.ensure(wins.size() * sizeof(int2)));
        CKC(ctx->d_tii.ensure(wins.size() * sizeof(TiiResult)));
        CKC(ctx->h_sign.ensure((size_t)n_slots * 4));
        CKC(ctx->h_windows.ensure(wins.size() * sizeof(int2)));
        memcpy(ctx->h_sign.p, sign.data(), (size_t)n_slots * 4);
        memcpy(ctx->h_windows.p, wins.data(), wins.size() * sizeof(int2));
        if (ctx->fetch(ctx->d_sign.p, ctx->h_sign.p, (size_t)n_slots * 4) || ctx->fetch(ctx->d_windows.p, ctx->h_windows.p, wins.size() * sizeof(int2))) return -1;
        k_tii_detect<<<n_win, 384, 0, ctx->S()>>>(ctx->d_fold.as<float>(), ctx->d_norm.as<float>(), ctx->d_sign.as<int>(), ctx->d_windows.as<int2>(),
                                                     6.0f, ctx->d_tii.as<TiiResult>());
        ctx->launches += 1;
        CKC(cudaGetLastError());
        CKC(ctx->h_tii.ensure(wins.size() * sizeof(TiiResult)));
        CKC(ctx->h_fold.ensure((size_t)n_slots * TII_SPECTRUM_BINS * 4));
        CKC(ctx->h_norm.ensure((size_t)n_slots * 4));
        CKC(cudaEventRecord(ctx->ev_tii, ctx->S()));
        CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_tii, 0));
        CKC(cudaMemcpyAsync(ctx->h_tii.p, ctx->d_tii.p, wins.size() * sizeof(TiiResult), cudaMemcpyDeviceToHost, ctx->copy_stream));
        CKC(cudaMemcpyAsync(ctx->h_fold.p, ctx->d_fold.p, (size_t)n_slots * TII_SPECTRUM_BINS * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CKC(cudaMemcpyAsync(ctx->h_norm.p, ctx->d_norm.p, (size_t)n_slots * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        copies_pending = true;
        res->tii_fold = ctx->h_fold.as<float>(); res->tii_norm = ctx->h_norm.as<float>(); res->tii_sign = ctx->r_sign.data();
    }
    return 0;
    };
    bool exported = false;              // the trace kernel wrote the logical frames to the host buffer itself
    bool patching = false;              // ... those of the DAB+ sub-channels as traced: the RS corrections are applied here
    size_t n_patch_copied = 0;
    auto msc_block = [&]() -> int {
    exported = false; patching = false; n_patch_copied = 0;
    if (!(flags & B200DAB_OPT_NO_MSC) && n_rows > 0 && !ctx->r_subch.empty()) {
        ctx->ev_first_trace = ctx->ev_msc;
        ctx->export_frames = ctx->frame_export;
        const int msc_rc = ctx->msc(n_slots, ctx->r_subch, out_off, &msc_bytes);
        ctx->export_frames = false;
        ctx->ev_first_trace = nullptr;
        if (msc_rc) return -1;
        exported = ctx->frames_exported;
        TRACE("msc submitted");
        for (size_t i = 0; i < ctx->r_subch.size(); i++) {
            if (ctx->r_subch[i].kind != 1) continue;
            DabPlusJob j;
            j.frames_off = (long long)out_off[i]; j.n_rows = n_rows; j.pitch = ctx->r_subch[i].bitrate * 3; j.s = ctx->r_subch[i].bitrate / 8;
            j.sf_off = total_sf; j.max_sf = n_rows / 5; j.flag_off = total_rows; j.subch_index = (int)i;
            j.forced_phase = i < forced_phase.size() ? forced_phase[i] : -1; j.pad_ = 0;
            total_sf += j.max_sf; total_rows += n_rows;
            dpjobs.push_back(j); dp_subch.push_back((int)i);
        }
        // The logical frames of the DAB (MP2) and data sub-channels are final once the trace kernel has run: their copy to the host
        // goes to the copy stream and runs under the DAB+ kernels (fire code, RS); the DAB+ sub-channels follow on the copy stream
        // as soon as the RS correction is through, under the access-unit kernels.  Adjacent sub-channels of the same class are
        // copied in one piece.
        CKC(ctx->h_msc.ensure(msc_bytes + 64));
        auto copy_class = [&](bool dabplus_class, cudaStream_t st) -> int {
            const size_t n = ctx->r_subch.size();
            for (size_t i = 0; i < n;) {
                if ((ctx->r_subch[i].kind == 1) != dabplus_class) { i++; continue; }
                size_t j = i;
                while (j + 1 < n && (ctx->r_subch[j + 1].kind == 1) == dabplus_class) j++;
                const size_t a = out_off[i], b = (j + 1 < n) ? out_off[j + 1] : msc_bytes;
                CKC(cudaMemcpyAsync((uint8_t*)ctx->h_msc.p + a, (const uint8_t*)ctx->d_msc_out.p + a, b - a, cudaMemcpyDeviceToHost, st));
                i = j + 1;
            }
            return 0;
        };
        static const bool no_split = getenv("B200DAB_NO_SPLIT_COPY") != nullptr;      // A/B switch (tools/ab_step.py)
        const bool split = !dpjobs.empty() && !no_split && !exported;
        patching = exported && !dpjobs.empty() && total_sf > 0;
        if (split) {
            if (!ctx->first_trace_recorded) CKC(cudaEventRecord(ctx->ev_msc, ctx->stream));      // (else: recorded between the two trace launches)
            CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_msc, 0));
            if (copy_class(false, ctx->copy_stream)) return -1;
            copies_pending = true;
        }
        cudaStream_t back = ctx->late_on_aux ? ctx->aux_stream : ctx->stream;      // where the DAB+ half of the back end runs
        ctx->cur = ctx->late_on_aux ? ctx->aux_stream : nullptr;
        ctx->rs_patch_mode = patching;
        const int dp_rc = ctx->dabplus(ctx->d_msc_out.as<uint8_t>(), dpjobs, total_sf, total_rows);
        ctx->rs_patch_mode = false;
        ctx->cur = nullptr;
        if (dp_rc) return -1;
        if (!patching) {
            if (split) {
                CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rs, 0));
                if (copy_class(true, ctx->copy_stream)) return -1;
            } else if (!exported) CKC(cudaMemcpyAsync(ctx->h_msc.p, ctx->d_msc_out.p, msc_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (total_sf > 0) {
            // totals, access-unit records, superframe details, flags, alignments and (the frames being on the host already, as traced) the
            // first RS corrections: one copy
            static_assert(sizeof(AuRecord) == sizeof(b200dab_au_t), "AU record layout");
            n_patch_copied = patching ? std::min<size_t>(ctx->rs_patch_cap, ctx->rs_patch_prefix) : 0;
            const size_t len = ctx->dp.patch + n_patch_copied * sizeof(RsPatch);
            CKC(ctx->h_dpres.ensure(len + 64));
            CKC(cudaMemcpyAsync(ctx->h_dpres.p, ctx->d_dpres.p, len, cudaMemcpyDeviceToHost, back));
        }
        if (ctx->late_on_aux) { CKC(cudaEventRecord(ctx->ev_aux, ctx->aux_stream)); CKC(cudaStreamWaitEvent(ctx->stream, ctx->ev_aux, 0)); }
    }
    TRACE("msc+dab+ submitted");
    return 0;
    };
    // ---- MSC chain on the main stream; the FIC chain is still running (or already through) on its own
    if (msc_block()) return -1;
    // ---- everything below needs the whole FIC on the host
    if (!fic_synced) { CKC(cudaEventSynchronize(ctx->ev_fic)); TRACE("fic synced"); }
    if (!(subch_in && n_subch_in > 0) && !(flags & B200DAB_OPT_NO_MSC)) {
        // The table the MSC chain was queued with came from a prefix of the FIC (so that the GPU never waits for the host).  What
        // counts is the organisation the WHOLE FIC of this call describes: a recording that starts in noise can show a stable but
        // incomplete table in its first frames.  If the two differ (rare), the MSC chain runs again with the full table.
        std::vector<b200dab_subch_t> full(B200DAB_MAX_SUBCH);
        const int nf = parse_subch_impl(res->fibs, res->fib_ok, n_slots * 12, full.data(), B200DAB_MAX_SUBCH);
        full.resize(nf);
        ctx->drop_undecodable(full);
        const bool same = full.size() == ctx->r_subch.size() && (full.empty() || !memcmp(full.data(), ctx->r_subch.data(), full.size() * sizeof(b200dab_subch_t)));
        if (!same) {
            TRACE("sub-channel table revised from the whole FIC");
            ctx->r_subch = full;
            res->n_subch = (int)ctx->r_subch.size(); res->subch = ctx->r_subch.data();
            dpjobs.clear(); dp_subch.clear(); total_sf = 0; total_rows = 0; msc_bytes = 0; out_off.clear();
            if (msc_block()) return -1;
        }
    }
    ctx->cur = ctx->fic_stream;
    const int tii_rc = tii_block();
    ctx->cur = nullptr;
    if (tii_rc) return -1;
    CKC(cudaEventRecord(ctx->ev_fic_stream, ctx->fic_stream));
    CKC(cudaStreamWaitEvent(ctx->stream, ctx->ev_fic_stream, 0));
    if (want_stat) {
        CKC(ctx->h_symstat.ensure((size_t)n_slots * 150 * 4));
        CKC(cudaMemcpyAsync(ctx->h_symstat.p, ctx->d_symstat.p, (size_t)n_slots * 150 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (copies_pending) {
        CKC(cudaEventRecord(ctx->ev_d2h, ctx->copy_stream));
        CKC(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h, 0));
    }
    TRACE("all submitted");
    CKC(cudaStreamSynchronize(ctx->stream));
    TRACE("final sync");
    // ---- RS corrections of the DAB+ frames that reached the host as traced
    if (patching && total_sf > 0) {
        const int n_patches = ctx->hpp<DabPlusTotals>(0)->n_patches;
        if (n_patches >= 0 && (size_t)n_patches <= n_patch_copied) {
            const RsPatch* e = ctx->hpp<RsPatch>(ctx->dp.patch);
            uint8_t* fr = ctx->h_msc.as<uint8_t>();
            for (int i = 0; i < n_patches; i++) if (e[i].offset < msc_bytes) fr[e[i].offset] = (uint8_t)e[i].value;
        } else {
            // more corrections than the prefix that was copied (a recording close to the decoding threshold): the corrected frames themselves
            for (size_t i = 0; i < ctx->r_subch.size(); i++) {
                if (ctx->r_subch[i].kind != 1) continue;
                const size_t a = out_off[i], b = (i + 1 < ctx->r_subch.size()) ? out_off[i + 1] : msc_bytes;
                CKC(cudaMemcpyAsync((uint8_t*)ctx->h_msc.p + a, (const uint8_t*)ctx->d_msc_out.p + a, b - a, cudaMemcpyDeviceToHost, ctx->stream));
            }
            CKC(cudaStreamSynchronize(ctx->stream));
        }
    }
    // ---- assemble the result
    if (msc_bytes > 0) {
        for (size_t i = 0; i < ctx->r_subch.size(); i++) { ctx->r_frames.push_back(ctx->h_msc.as<uint8_t>() + out_off[i]); ctx->r_rows.push_back(n_rows); }
        res->frames = ctx->r_frames.data(); res->frame_rows = ctx->r_rows.data();
        if (total_sf > 0) {
            // the access-unit list and the RS totals were put together on the device (k_au_compact)
            const DabPlusTotals* tot = ctx->hpp<DabPlusTotals>(0);
            res->superframes = tot->superframes; res->rs_corrected = tot->rs_corrected; res->rs_bit_errors = tot->rs_bit_errors;
            res->rs_uncorrectable = tot->rs_uncorrectable; res->firecode_fail = tot->firecode_fail;
            res->n_aus = tot->n_aus;
            res->aus = ctx->hpp<b200dab_au_t>(ctx->dp.aus);
            // the superframe layer in detail
            ctx->r_sf_flags.assign(ctx->r_subch.size(), nullptr);
            ctx->r_sf_phase.assign(ctx->r_subch.size(), -1);
            const SfInfo* si = ctx->hpp<SfInfo>(ctx->dp.sf);
            ctx->r_sf.reserve((size_t)total_sf);
            for (size_t k = 0; k < dpjobs.size(); k++) {
                const DabPlusJob& j = dpjobs[k];
                ctx->r_sf_flags[j.subch_index] = ctx->hpp<uint8_t>(ctx->dp.flags) + j.flag_off;
                ctx->r_sf_phase[j.subch_index] = ctx->hpp<int32_t>(ctx->dp.res)[k];
                for (int m = 0; m < j.max_sf; m++) {
                    const SfInfo& q = si[j.sf_off + m];
                    if (!q.valid) continue;
                    ctx->r_sf.push_back(b200dab_sf_t{j.subch_index, q.row, q.firecode_ok, q.rs_corrected, q.rs_bit_errors, q.rs_uncorrectable, q.n_aus, q.header});
                }
            }
            res->n_sf = (int)ctx->r_sf.size(); res->sf = ctx->r_sf.data();
            res->sf_flags = ctx->r_sf_flags.data(); res->sf_phase = ctx->r_sf_phase.data();
        }
    }
    if (want_tii && n_win > 0) {
        static_assert(sizeof(TiiResult) == sizeof(b200dab_tii_t), "TII result layout");
        ctx->r_tii.resize(n_win);
        memcpy(ctx->r_tii.data(), ctx->h_tii.p, (size_t)n_win * sizeof(TiiResult));
        res->n_tii = n_win; res->tii = ctx->r_tii.data();
    }
    if (want_stat) {
        res->symstat = ctx->h_symstat.as<float>();
        double snr_acc = 0; int cnt = 0;
        for (int s = 0; s < n_slots; s++) {
            if (ctx->r_tf[s] < 0) continue;
            for (int l = 0; l < 75; l++) {
                const float mu = res->symstat[((size_t)s * 75 + l) * 2], m2 = res->symstat[((size_t)s * 75 + l) * 2 + 1];
                const double var = (double)m2 - (double)mu * mu;
                if (mu > 0 && var > 0) { snr_acc += (double)mu * mu / var; cnt++; }
            }
        }
        // decision-directed estimate, put on the scale of the reference's snr10 (include/b200dab_snr_cal.h: measured, not modelled)
        res->snr_db = cnt ? b200dab_snr_calibrated((float)(10.0 * log10(snr_acc / cnt))) : 0.f;
    }
    res->gpu_launches = ctx->launches;
    TRACE("result assembled");
    return 0;
}

// ------------------------------------------------------------------------------------------------ stage entry points
int b200dab_stage_coarse(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, float* P, float* S, int64_t* cand, int32_t* n_cand, int max_cand_out)
{
    if (!ctx || !iq) return -1;
    cudaSetDevice(ctx->device);
    const void* d_iq;
    if (ctx->upload_raw(iq, n, fmt, B200DAB_MEM_HOST, &d_iq)) return -1;
    int nc = 0, mc = 0;
    if (ctx->coarse(d_iq, n, fmt, &nc, &mc)) return -1;
    const size_t nblk = n / SYNC_BLK;
    if (P) CKC(cudaMemcpy(P, ctx->d_P.p, nblk * 4, cudaMemcpyDeviceToHost));
    if (S) CKC(cudaMemcpy(S, ctx->d_S.p, nblk * 4, cudaMemcpyDeviceToHost));
    if (nc > max_cand_out) nc = max_cand_out;
    if (cand && nc > 0) CKC(cudaMemcpy(cand, ctx->d_cand.p, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    if (n_cand) *n_cand = nc;
    return 0;
}

int b200dab_stage_fine(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, const int64_t* cand, int n_cand, int64_t* t_f, float* cfo, float* metric)
{
    if (!ctx || !iq || !cand || n_cand <= 0) return -1;
    cudaSetDevice(ctx->device);
    const void* d_iq;
    if (ctx->upload_raw(iq, n, fmt, B200DAB_MEM_HOST, &d_iq)) return -1;
    CKC(ctx->d_cand.ensure((size_t)n_cand * 8));
    CKC(ctx->d_ncand.ensure(16));
    CKC(cudaMemcpyAsync(ctx->d_cand.p, cand, (size_t)n_cand * 8, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemsetAsync(ctx->d_ncand.p, 0, 16, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_ncand.p, &n_cand, 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    if (ctx->fine(d_iq, n, fmt, n_cand, n_cand, 0, -1, 0)) return -1;
    std::vector<FineOut> f(n_cand);
    CKC(cudaMemcpyAsync(f.data(), ctx->d_fine.p, (size_t)n_cand * sizeof(FineOut), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_cand; i++) {
        if (t_f) t_f[i] = f[i].t_f;
        if (cfo) cfo[i] = f[i].cfo_hz;
        if (metric) { metric[4 * i] = f[i].cp_mag; metric[4 * i + 1] = f[i].cfo_ratio; metric[4 * i + 2] = f[i].cir_ratio; metric[4 * i + 3] = f[i].tau1; }
    }
    return 0;
}

int b200dab_stage_demod(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, const int64_t* t_f, const float* cfo, int n_frames, int8_t* soft)
{
    if (!ctx || !iq || !t_f || !cfo || n_frames <= 0 || !soft) return -1;
    cudaSetDevice(ctx->device);
    const void* d_iq;
    if (ctx->upload_raw(iq, n, fmt, B200DAB_MEM_HOST, &d_iq)) return -1;
    if (ctx->prepare_slots(n_frames, false, false)) return -1;
    FrameInfo* fr = ctx->h_frames.as<FrameInfo>();
    for (int i = 0; i < n_frames; i++) { fr[i].t_f = t_f[i]; fr[i].cfo_hz = cfo[i]; fr[i].slot = i; }
    ChainState* cs = ctx->h_chain.as<ChainState>();
    cs->n_frames = n_frames; cs->n_candidates = 0; cs->last_t = 0;
    CKC(cudaMemcpyAsync(ctx->d_frames.p, fr, (size_t)n_frames * sizeof(FrameInfo), cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_chain.p, cs, 32, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->demod(d_iq, n, fmt, -1, 0, 1 << 30, n_frames, false, false)) return -1;
    CKC(cudaMemcpyAsync(soft, ctx->d_soft.p, (size_t)n_frames * DAB_FRAME_SOFT, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// raw Viterbi on already de-punctured soft bits: packs them into the kernel's [group][step][lane] layout on the host
int b200dab_stage_viterbi(b200dab_ctx* ctx, const int8_t* soft4, int n_cw, int steps, uint8_t* bits)
{
    if (!ctx || !soft4 || n_cw <= 0 || steps < 38 || ((steps - 6) % 32) != 0 || !bits) { if (ctx) ctx->set_error("stage_viterbi: steps-6 must be a multiple of 32"); return -1; }
    cudaSetDevice(ctx->device);
    const int ng = (n_cw + 31) / 32;
    std::vector<uint32_t> vin((size_t)ng * steps * 32, 0);
    for (int c = 0; c < n_cw; c++)
        for (int t = 0; t < steps; t++) {
            uint32_t w;
            memcpy(&w, soft4 + ((size_t)c * steps + t) * 4, 4);
            vin[((size_t)(c / 32) * steps + t) * 32 + (c % 32)] = w;
        }
    VitJob job;
    memset(&job, 0, sizeof job);
    job.n_rows = n_cw; job.steps = steps; job.n_info = steps - 6; job.out_pitch = job.n_info / 8;
    job.mode4 = (steps >= ctx->vit4_min_steps && steps % 2 == 0) ? 1 : 0;
    std::vector<VitGroup> groups;
    for (int g = 0; g < ng; g++) groups.push_back(VitGroup{0, g});
    std::vector<VitItem> items;
    ctx->build_items(std::vector<VitJob>(1, job), groups, items);
    const size_t words = vin.size();
    std::vector<uint32_t> zero_prbs(512, 0);
    CKC(ctx->d_jobs.ensure(sizeof job));
    CKC(ctx->d_groups.ensure(groups.size() * sizeof(VitGroup)));
    CKC(ctx->d_items.ensure(items.size() * sizeof(VitItem)));
    CKC(cudaMemcpyAsync(ctx->d_items.p, items.data(), items.size() * sizeof(VitItem), cudaMemcpyHostToDevice, ctx->stream));
    CKC(ctx->d_vin.ensure(words * 4));
    CKC(ctx->d_dec.ensure(words * 8));
    CKC(ctx->d_tmp_a.ensure(512 * 4));
    CKC(ctx->d_tmp_b.ensure((size_t)ng * 32 * job.out_pitch + 64));
    CKC(cudaMemcpyAsync(ctx->d_jobs.p, &job, sizeof job, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_groups.p, groups.data(), groups.size() * sizeof(VitGroup), cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_vin.p, vin.data(), words * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_tmp_a.p, zero_prbs.data(), 512 * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof_next();
    CKC(ctx->d_meet.ensure((size_t)ng * VIT_MEET_WORDS * sizeof(unsigned short)));
    ctx->prof_begin("k_viterbi(stage)");
    k_viterbi<<<ctx->vit_grid(items.size()), 32 * VIT_CTA_WARPS, 0, ctx->stream>>>(ctx->d_vin.as<uint32_t>(), ctx->d_dec.as<uint32_t>(), ctx->d_meet.as<unsigned short>(),
                                                                                  ctx->d_jobs.as<VitJob>(), ctx->d_groups.as<VitGroup>(), ctx->d_items.as<VitItem>(),
                                                                                  (int)items.size(), ctx->d_vit_counter.as<int>() + 3, nullptr);
    ctx->prof_end();
    ctx->prof_begin("k_vit_trace(stage)");
    k_vit_trace<<<ng, 64, 0, ctx->stream>>>(ctx->d_dec.as<uint2>(), ctx->d_meet.as<unsigned short>(), ctx->d_jobs.as<VitJob>(), ctx->d_groups.as<VitGroup>(),
                                           ctx->d_tmp_a.as<uint32_t>(), ctx->d_tmp_b.as<uint8_t>(), ctx->d_vit_counter.as<int>() + 3, -1, nullptr, nullptr, nullptr, 0);
    ctx->prof_end();
    ctx->launches += 2;
    CKC(cudaGetLastError());
    std::vector<uint8_t> packed((size_t)n_cw * job.out_pitch);
    CKC(cudaMemcpyAsync(packed.data(), ctx->d_tmp_b.p, packed.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < n_cw; c++)
        for (int i = 0; i < job.n_info; i++) bits[(size_t)c * job.n_info + i] = (packed[(size_t)c * job.out_pitch + i / 8] >> (7 - (i & 7))) & 1;
    return 0;
}

static int upload_soft(b200dab_ctx* ctx, const int8_t* soft, int n_frames)
{
    CKC(ctx->d_soft.ensure((size_t)n_frames * DAB_FRAME_SOFT + 256));
    CKC(cudaMemcpyAsync(ctx->d_soft.p, soft, (size_t)n_frames * DAB_FRAME_SOFT, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int b200dab_stage_fic(b200dab_ctx* ctx, const int8_t* soft, int n_frames, uint8_t* fibs, uint8_t* fib_ok)
{
    if (!ctx || !soft || n_frames <= 0) return -1;
    cudaSetDevice(ctx->device);
    if (upload_soft(ctx, soft, n_frames)) return -1;
    if (ctx->fic(n_frames)) return -1;
    if (fibs) CKC(cudaMemcpyAsync(fibs, ctx->d_fic_out.p, (size_t)n_frames * 384, cudaMemcpyDeviceToHost, ctx->stream));
    if (fib_ok) CKC(cudaMemcpyAsync(fib_ok, ctx->d_fib_ok.p, (size_t)n_frames * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dab_stage_msc(b200dab_ctx* ctx, const int8_t* soft, int n_frames, const b200dab_subch_t* sc, uint8_t* frames)
{
    if (!ctx || !soft || !sc || !frames || n_frames < 4) return -1;
    cudaSetDevice(ctx->device);
    if (upload_soft(ctx, soft, n_frames)) return -1;
    std::vector<b200dab_subch_t> v(1, *sc);
    std::vector<size_t> off;
    size_t total = 0;
    if (ctx->msc(n_frames, v, off, &total)) return -1;
    const int n_rows = 4 * n_frames - 15;
    CKC(cudaMemcpyAsync(frames, ctx->d_msc_out.p, (size_t)n_rows * sc->bitrate * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dab_stage_dabplus(b200dab_ctx* ctx, uint8_t* frames, int n_rows, int bitrate, b200dab_au_t* aus, int32_t* n_aus, int max_aus, int64_t* stats)
{
    if (!ctx || !frames || n_rows < 5 || bitrate % 8) return -1;
    cudaSetDevice(ctx->device);
    const int pitch = bitrate * 3;
    CKC(ctx->d_msc_out.ensure((size_t)n_rows * pitch + 256));
    CKC(cudaMemcpyAsync(ctx->d_msc_out.p, frames, (size_t)n_rows * pitch, cudaMemcpyHostToDevice, ctx->stream));
    DabPlusJob j;
    j.frames_off = 0; j.n_rows = n_rows; j.pitch = pitch; j.s = bitrate / 8; j.sf_off = 0; j.max_sf = n_rows / 5; j.flag_off = 0;
    j.subch_index = 0; j.forced_phase = -1; j.pad_ = 0;
    std::vector<DabPlusJob> jobs(1, j);
    if (ctx->dabplus(ctx->d_msc_out.as<uint8_t>(), jobs, j.max_sf, n_rows)) return -1;
    std::vector<SfInfo> sf(j.max_sf);
    CKC(cudaMemcpyAsync(frames, ctx->d_msc_out.p, (size_t)n_rows * pitch, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaMemcpyAsync(sf.data(), ctx->dpp<SfInfo>(ctx->dp.sf), sf.size() * sizeof(SfInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    int n = 0;
    int64_t st[5] = {0, 0, 0, 0, 0};
    for (int m = 0; m < j.max_sf; m++) {
        const SfInfo& s = sf[m];
        if (!s.valid) continue;
        st[0]++; st[1] += s.rs_corrected; st[2] += s.rs_bit_errors; st[3] += s.rs_uncorrectable; st[4] += s.firecode_ok ? 0 : 1;
        for (int a = 0; a < s.n_aus && n < max_aus; a++) {
            b200dab_au_t au;
            au.subch_index = 0; au.superframe = m; au.cif = s.row; au.offset = s.row * pitch + s.au_off[a]; au.length = s.au_len[a];
            au.crc_ok = s.au_ok[a]; au.header = s.header;
            if (aus) aus[n] = au;
            n++;
        }
    }
    if (n_aus) *n_aus = n;
    if (stats) memcpy(stats, st, sizeof st);
    return 0;
}

int b200dab_stage_tii_fold(b200dab_ctx* ctx, const float* nulls, int n, float* fold)
{
    if (!ctx || !nulls || n <= 0 || !fold) return -1;
    cudaSetDevice(ctx->device);
    CKC(ctx->d_tmp_a.ensure((size_t)n * DAB_NFFT * 8));
    CKC(ctx->d_fold.ensure((size_t)n * TII_SPECTRUM_BINS * 4));
    CKC(cudaMemcpyAsync(ctx->d_tmp_a.p, nulls, (size_t)n * DAB_NFFT * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_tii_fold_nulls<<<n, 256, 0, ctx->stream>>>(ctx->d_tmp_a.as<float2>(), ctx->d_tw.as<float2>(), ctx->d_fold.as<float>());
    ctx->launches += 1;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(fold, ctx->d_fold.p, (size_t)n * TII_SPECTRUM_BINS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dab_stage_spectrum(b200dab_ctx* ctx, const void* iq, uint64_t n, int fmt, const int64_t* w0, int n_w, float* out)
{
    if (!ctx || !iq || !w0 || n_w <= 0 || !out || fmt < 0 || fmt > 2) return -1;
    cudaSetDevice(ctx->device);
    const void* d_iq;
    if (ctx->upload_raw(iq, n, fmt, B200DAB_MEM_HOST, &d_iq)) return -1;
    CKC(ctx->d_tmp_a.ensure((size_t)n_w * 8));
    CKC(ctx->d_spec.ensure((size_t)n_w * DAB_NFFT * 4));
    CKC(cudaMemcpyAsync(ctx->d_tmp_a.p, w0, (size_t)n_w * 8, cudaMemcpyHostToDevice, ctx->stream));
    const float2* tw = ctx->d_tw.as<float2>();
    if (fmt == B200DAB_FMT_U8) k_spectrum_at<B200DAB_FMT_U8><<<n_w, 256, 0, ctx->stream>>>(d_iq, (int64_t)n, ctx->d_tmp_a.as<long long>(), tw, ctx->d_spec.as<float>());
    else if (fmt == B200DAB_FMT_S16) k_spectrum_at<B200DAB_FMT_S16><<<n_w, 256, 0, ctx->stream>>>(d_iq, (int64_t)n, ctx->d_tmp_a.as<long long>(), tw, ctx->d_spec.as<float>());
    else k_spectrum_at<B200DAB_FMT_F32><<<n_w, 256, 0, ctx->stream>>>(d_iq, (int64_t)n, ctx->d_tmp_a.as<long long>(), tw, ctx->d_spec.as<float>());
    ctx->launches += 1;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(out, ctx->d_spec.p, (size_t)n_w * DAB_NFFT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dab_stage_tii_detect(b200dab_ctx* ctx, const float* fold, const float* norm, const int32_t* sign, int n, float thr_k, b200dab_tii_t* out)
{
    if (!ctx || !fold || !norm || !sign || n <= 0 || !out) return -1;
    cudaSetDevice(ctx->device);
    CKC(ctx->d_fold.ensure((size_t)n * TII_SPECTRUM_BINS * 4));
    CKC(ctx->d_norm.ensure((size_t)n * 4));
    CKC(ctx->d_sign.ensure((size_t)n * 4));
    CKC(ctx->d_windows.ensure(sizeof(int2)));
    CKC(ctx->d_tii.ensure(sizeof(TiiResult)));
    const int2 win = make_int2(0, n);
    CKC(cudaMemcpyAsync(ctx->d_fold.p, fold, (size_t)n * TII_SPECTRUM_BINS * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_norm.p, norm, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_sign.p, sign, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->d_windows.p, &win, sizeof win, cudaMemcpyHostToDevice, ctx->stream));
    k_tii_detect<<<1, 384, 0, ctx->stream>>>(ctx->d_fold.as<float>(), ctx->d_norm.as<float>(), ctx->d_sign.as<int>(), ctx->d_windows.as<int2>(), thr_k,
                                            ctx->d_tii.as<TiiResult>());
    ctx->launches += 1;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(out, ctx->d_tii.p, sizeof(TiiResult), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
