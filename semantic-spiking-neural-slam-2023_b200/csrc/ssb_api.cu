// Host side of the C-ABI declared in include/sspslam_b200.h: plan upload, arenas,
// the per-step launch sequence, probes, event timing.  No torch, no Python types.
#include "ssb_kernels.cuh"
#include "../../include/sspslam_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define SSB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(-2, std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
    } while (0)

enum Kind { K_SMALL = 0, K_WIDE, K_DEC, K_PES, K_SCAN, K_PICK, K_GATE, K_LIN, K_ADV, K_BEGIN, K_VOJA, K_EXTRA, K_PHIST, K_PFOLD, K_NKINDS };   // K_EXTRA: further kernels of a multi-kernel scope; K_PES = the sparse decode, K_PHIST / K_PFOLD the history append and the fold

struct HostArray {
    std::vector<unsigned char> bytes;
};

struct CleanupDev {
    float* cx = nullptr;
    float* pval = nullptr;
    int* pidx = nullptr;
    int* idx = nullptr;
    const double* s64 = nullptr;
    int n_chunks = 0, rows_per_chunk = 0, tile_rows = 0;
    // tensor-core scan (k_cleanup_scan_tc): pre-tiled 3xTF32 grid, K padded to a multiple of 8
    bool tc = false;
    float* stc = nullptr;
    int kp = 0, n_tiles = 0, tr = 128;
    // K-blocked tensor-core scan (k_cleanup_scan_tck) for operand widths whose tiles do not fit in shared memory
    bool tck = false;
    float* stck = nullptr;   // grid, pre-tiled per (tile, K block)
    float* xt = nullptr;     // queries, tiled per (trial block, K block) by k_scan_xtiles every step
    int n_kb = 0;
};

// grid rows per tensor-core tile: 128, or 64 when the operand tiles of 128 rows do not fit in shared memory; 0 = no fit
int scan_tc_rows(int dpad) {
    const int kp = (dpad + 7) / 8 * 8;
    // SSB_SCAN_TR=64: 64-row tiles even when 128 fit (114 KB instead of 172 KB of shared memory at d = 55, so a scan CTA
    // can share its SM with a CTA of the streaming kernels)
    const char* e = getenv("SSB_SCAN_TR");
    const bool force64 = e && atoi(e) == 64;
    for (int tr : {128, 64})
        if (!(force64 && tr == 128) && (size_t)(2 * 128 + 4 * tr) * kp * sizeof(float) <= 216 * 1024) return tr;
    return 0;
}

// SSB_SCAN=ffma forces the FFMA scan (the measured comparison in DESIGN.md); default is the tcgen05 scan
// whenever its operand tiles fit in shared memory.
bool scan_tc_allowed(int dpad) {
    const char* e = getenv("SSB_SCAN");
    if (e && std::string(e) == "ffma") return false;
    return scan_tc_rows(dpad) > 0;
}

bool scan_ffma_forced() {
    const char* e = getenv("SSB_SCAN");
    return e && std::string(e) == "ffma";
}

// Grid-scan geometry: a CTA scans rows_per_chunk grid rows in shared-memory tiles of tile_rows rows for 4 trial
// groups.  Enough chunks for ~8 CTAs per SM (the scan is FFMA-bound and needs resident warps), at most
// SSB_SCAN_MAX_CHUNKS (each chunk leaves TOPK candidates per trial); tiles of at most 24 KB.
void scan_geometry(int G, int dpad, int n_groups, CleanupDev* cd) {
    const int group_ctas = (n_groups + 3) / 4;
    if (scan_tc_allowed(dpad)) {   // one CTA per SM: trial blocks x grid chunks ~ 148
        cd->tc = true;
        cd->kp = (dpad + 7) / 8 * 8;
        cd->tr = scan_tc_rows(dpad);
        cd->n_tiles = (G + cd->tr - 1) / cd->tr;
        cd->n_chunks = std::max(1, std::min(std::min(cd->n_tiles, SSB_SCAN_MAX_CHUNKS), 148 / std::max(1, group_ctas)));
        cd->rows_per_chunk = cd->tile_rows = cd->tr;
        return;
    }
    if (!scan_ffma_forced()) {     // wide operands (d = 649): both operands stream through K blocks of 32 columns
        cd->tck = true;
        cd->n_kb = (dpad + SSB_SCK_KB - 1) / SSB_SCK_KB;
        cd->tr = 128;
        cd->n_tiles = (G + 127) / 128;
        cd->n_chunks = std::max(1, std::min(std::min(cd->n_tiles, SSB_SCAN_MAX_CHUNKS), 148 / std::max(1, group_ctas)));
        cd->rows_per_chunk = cd->tile_rows = 128;
        return;
    }
    int want = std::max(1, (148 * 8 + group_ctas - 1) / group_ctas);
    want = std::min(std::min(want, SSB_SCAN_MAX_CHUNKS), std::max(1, G / 16));
    int rows = (G + want - 1) / want;
    rows = (rows + 1) & ~1;                              // two rows per iteration
    cd->rows_per_chunk = rows;
    cd->n_chunks = (G + rows - 1) / rows;
    const int max_tile = std::max(2, ((24 * 1024) / (dpad * (int)sizeof(float))) & ~1);
    cd->tile_rows = std::min(rows, max_tile);
}

}  // namespace

struct LevelInfo {
    int n_static = 0, n_voja = 0;   // wide ensembles of the level by kernel flavour
    bool dec_needs_voja = false;    // a static decoder of the level reads a Voja ensemble's activities
};

struct ssb_sim {
    int device = 0;
    int n_trials = 0, B = 0, n_groups = 0;
    cudaStream_t stream = nullptr;
    bool finalized = false;
    std::map<std::string, HostArray> arrays;
    std::map<std::string, double> scalars;
    // plan (device)
    int* d_csr_ptr = nullptr;
    int2 *d_ent0 = nullptr, *d_ent1 = nullptr;
    float* d_W = nullptr;
    int *d_small = nullptr, *d_big = nullptr, *d_dec = nullptr, *d_pes = nullptr, *d_cleanup = nullptr, *d_gate = nullptr;
    int* d_lin_rows = nullptr;
    float* d_lin_ab = nullptr;
    // dense blocks of the row program (rows sharing one column list), found at finalize
    int *d_dense_items = nullptr, *d_dense_desc = nullptr, *d_dense_cols = nullptr, *d_dense_rows = nullptr;
    float* d_dense_T = nullptr;
    struct LinSeg { int csr_row0 = 0, n_csr = 0, item0 = 0, n_items = 0, rec0 = 0, n_recs = 0, tc0 = 0, n_tc = 0, tc_max_kb = 0, tc_max_tiles = 0; };
    SsbLinTcBlock* d_lin_tc = nullptr;      // large dense blocks served by k_lin_tck (tcgen05), per segment [tc0, tc0 + n_tc)
    float *d_lin_ttk = nullptr, *d_lin_xt = nullptr;
    int* d_lin_recs = nullptr;
    std::vector<LinSeg> lin_segs;            // one per level + the early end-of-step rows (need only level-0 narrow ensembles) + the rest
    long long n_dense_rows = 0, n_dense_blocks = 0;
    float* d_ntypes = nullptr;
    double* d_s64 = nullptr;
    std::vector<int> h_stages, h_small, h_big, h_dec, h_cleanup, h_pes;
    cudaGraphExec_t step_graph = nullptr;   // graph_steps consecutive steps (the step counter lives on the device)
    int graph_steps = 0;
    int graph_phase = 0;                    // steps_done mod (PES window) the graph was captured at
    bool use_graph = true;
    bool debug_sync = false, debug_failed = false;
    // independent kernels of one dependency level run on side streams (fork/join with events; under
    // capture these become parallel branches of the step graph)
    bool parallel = true;
    cudaStream_t aux[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // B, C, D, E, F (input prefetch), G (early end-of-step rows)
    std::vector<int> level_deps;            // [n_levels][n_levels] producer-kind bits (lowering.py); empty = wait for everything
    cudaStream_t io_h2d = nullptr, io_d2h = nullptr;   // copy streams of ssb_run_steps_io (one per DMA direction)
    std::vector<cudaEvent_t> io_events;
    std::vector<cudaEvent_t> dep_pool;
    size_t dep_used = 0;
    int pes_level = -1;
    bool pes_needs_static = false;
    std::vector<LevelInfo> levels;
    float* d_enc_t = nullptr;               // static wide encoders pre-tiled for k_wide_static_tc (3xTF32 hi | lo)
    int* d_enc_t_off = nullptr;
    std::vector<int> enc_t_off;             // host copy: -1 = not eligible
    float* d_dec_wt = nullptr;              // static decoders pre-tiled for k_decode_tc (3xTF32 hi | lo)
    int* d_dec_wt_off = nullptr;
    int dec_tc_n = 64;                      // operand tile width of k_decode_tc (64 or 128 output columns)
    int dec_tc_nt = 1;                      // column tiles per decoder
    std::vector<char> dec_tc_level;         // per level: every decoder of the level can use the tensor-core kernel
    // on-device input synthesis (ssb_synth_setup): replaces k_begin and the input tables
    bool synth_on = false;
    SsbSynth synth;
    float *syn_path = nullptr, *syn_vel = nullptr, *syn_lm = nullptr, *syn_phases = nullptr, *syn_lmsp = nullptr;
    float *syn_cos = nullptr, *syn_sin = nullptr;
    int* syn_idx = nullptr;
    long long syn_step0 = 0;
    int syn_steps = 0;
    SsbPesDefer pes_h = {nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};   // deferred PES history (K = 0: off)
    // K-blocked tensor-core encode of static wide ensembles with very wide inputs (k_wide_static_tck), per level
    struct TckLevel {
        bool on = false;
        SsbTckItems items;
        int* d_desc6 = nullptr;
        int n_kb = 0, n_tiles_max = 0;
        long long xt_stride = 0;
    };
    std::vector<TckLevel> tck_levels;
    float *d_etk = nullptr, *d_xtk = nullptr;
    bool pes_fused = false;                 // the sparse decode runs inside k_wide_voja (every PES pre-ensemble is a Voja ensemble)
    std::vector<int> pes_of_big;            // wide-ensemble descriptor -> PES descriptor it feeds, or -1
    int* d_pes_hdesc = nullptr;
    size_t pes_pad_smem = 0, voja_pad_smem = 0;   // experiment knobs: extra dynamic smem lowers residency
    std::map<int, int> wide_chunk_cache;   // launch geometry of the wide-ensemble kernels, decided once
    long long kind_per_graph[16] = {0};     // launches per graph replay by kind (counted while capturing)
    std::vector<size_t> s64_offsets;
    int n_levels = 0, n_lin = 0, lin0 = 0, n_lin_early = 0, n_lin_fused = 0, n_lvl0_res = 0, n_pes = 0, n_small_total = 0;
    // sizes
    long long nv = 0, nf = 0, nt = 0, nn = 0, n_act = 0, n_lenc = 0, n_ldec = 0, n_afilt = 0, n_probe = 0;
    long long tab_row0 = 0, n_part = 0, n_counters = 0;
    int chunk_cap = 0;
    // arenas
    float *vec = nullptr, *tab = nullptr, *st = nullptr, *act = nullptr, *lenc = nullptr, *ldec = nullptr;
    float *afilt = nullptr, *probe = nullptr, *part = nullptr;
    float* wpt = nullptr;                   // per-trial static weights [G][n_wpt][32] (scalar per_trial_weights = 1)
    long long n_wpt = 0;
    bool lin_six = false;                   // SSB_LIN_MINB=6: always the 6-CTAs-per-SM variant of k_lin
    bool dec_sparse = false;                // SSB_DECODE=sparse: shared static decoders through the sparse per-trial walk (spiking runs)
    bool per_trial = false;                 // plan lowered with per-trial static weights (wide ensembles: encoders in lenc, decoders in ldec)
    int* counters = nullptr;
    int* aflag = nullptr;
    long long* dyn = nullptr;
    std::vector<CleanupDev> cleanups;
    int* cidx = nullptr;  // [n_cleanup][B] (trial-major, not tiled)
    // host mirrors of dyn
    long long steps_done = 0, tab_step0 = 0, probe_step0 = 0;
    long long step_base = 0;                // absolute step of i_rel = 0 for the launches being issued (host mirror of dyn[])
    int tab_steps = 0;
    // timing
    bool profiling = false;
    bool timeline = false;                  // profiling with the dependency streams kept (per-launch start / end times)
    cudaEvent_t ev_timeline0 = nullptr;
    std::vector<float> tl_start, tl_end;
    std::vector<int> tl_kind;
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;
    cudaEvent_t ev_mark[4] = {nullptr, nullptr, nullptr, nullptr};
    bool run_timed = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_used;  // (kind, index of first event of the pair)
    float kind_ms[K_NKINDS] = {0};
    long long kind_launches[K_NKINDS] = {0};
    long long total_launches = 0;
    SsbCtx ctx;
};

namespace {

template <typename T>
int upload_array(ssb_sim* s, const char* name, T** dst, size_t* count = nullptr) {
    auto it = s->arrays.find(name);
    size_t bytes = it == s->arrays.end() ? 0 : it->second.bytes.size();
    if (count) *count = bytes / sizeof(T);
    size_t alloc = bytes ? bytes : sizeof(T) * 4;
    SSB_CUDA(cudaMalloc((void**)dst, alloc));
    SSB_CUDA(cudaMemset(*dst, 0, alloc));
    if (bytes) SSB_CUDA(cudaMemcpy(*dst, it->second.bytes.data(), bytes, cudaMemcpyHostToDevice));
    return 0;
}

std::vector<int> host_ints(ssb_sim* s, const char* name) {
    auto it = s->arrays.find(name);
    if (it == s->arrays.end()) return {};
    const int* p = reinterpret_cast<const int*>(it->second.bytes.data());
    return std::vector<int>(p, p + it->second.bytes.size() / sizeof(int));
}

long long iscalar(ssb_sim* s, const char* name) {
    auto it = s->scalars.find(name);
    return it == s->scalars.end() ? 0 : (long long)(it->second + 0.5);
}

int alloc_rows(float** p, long long rows, int B) {
    size_t bytes = (size_t)(rows > 0 ? rows : 1) * B * sizeof(float);
    SSB_CUDA(cudaMalloc((void**)p, bytes));
    SSB_CUDA(cudaMemset(*p, 0, bytes));
    return 0;
}

struct ArenaRef {
    float* ptr;
    long long rows;
    bool tiled;
};

int arena(ssb_sim* s, const char* name, ArenaRef* out) {
    std::string n(name);
    if (n == "st") *out = {s->st, s->nn, true};
    else if (n == "act") *out = {s->act, s->n_act, true};
    else if (n == "lenc") *out = {s->lenc, s->n_lenc, true};
    else if (n == "ldec") *out = {s->ldec, s->n_ldec, true};
    else if (n == "afilt") *out = {s->afilt, 2 * s->n_afilt, true};
    else if (n == "vec") *out = {s->vec, s->nv, true};
    else if (n == "wpt") *out = {s->wpt, s->n_wpt, true};
    else if (n == "cidx") *out = {reinterpret_cast<float*>(s->cidx), (long long)s->cleanups.size(), false};
    else return fail(-3, "unknown arena '" + n + "'");
    return 0;
}

// Host rows are [n_rows][B] (trial contiguous); device arenas are tiled [G][arena_rows][32].
// One strided 2-D copy per trial group moves n_rows lines of 128 bytes.
int copy_rows(ssb_sim* s, float* dev_base, long long arena_rows, size_t row0, size_t n_rows, float* host, bool to_device,
              cudaStream_t st = nullptr) {
    const size_t B = s->B;
    if (!st) st = s->stream;
    for (int g = 0; g < s->n_groups; ++g) {
        float* d = dev_base + ((size_t)g * arena_rows + row0) * 32;
        float* h = host + (size_t)g * 32;
        if (to_device)
            SSB_CUDA(cudaMemcpy2DAsync(d, 128, h, B * sizeof(float), 128, n_rows, cudaMemcpyHostToDevice, st));
        else
            SSB_CUDA(cudaMemcpy2DAsync(h, B * sizeof(float), d, 128, 128, n_rows, cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

cudaEvent_t* next_events(ssb_sim* s, int kind) {
    size_t used = s->ev_used.size() * 2;
    while (s->ev_pool.size() < used + 2) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s->ev_pool.push_back(e);
    }
    s->ev_used.push_back({kind, (int)used});
    return &s->ev_pool[used];
}

struct LaunchTimer {
    ssb_sim* s;
    cudaEvent_t* ev = nullptr;
    int kind;
    cudaStream_t st;
    // `st_`: the stream the launch goes to (timeline mode keeps the dependency streams, so the event pair must sit on it)
    LaunchTimer(ssb_sim* s_, int kind_, cudaStream_t st_ = nullptr) : s(s_), kind(kind_), st(st_ ? st_ : s_->stream) {
        s->kind_launches[kind]++;
        s->total_launches++;
        if (s->profiling) {
            ev = next_events(s, kind);
            cudaEventRecord(ev[0], st);
        }
    }
    ~LaunchTimer() {
        if (ev) cudaEventRecord(ev[1], st);
        if (s->debug_sync && !s->debug_failed) {   // SSB_DEBUG_SYNC=1: find the launch that faults
            cudaError_t e = cudaStreamSynchronize(s->stream);
            if (e != cudaSuccess) {
                s->debug_failed = true;
                fprintf(stderr, "[ssb] kernel kind %d (launch #%lld) failed: %s\n", kind, s->total_launches,
                        cudaGetErrorString(e));
            }
        }
    }
};

int collect_profile(ssb_sim* s) {
    if (s->ev_used.empty()) return 0;
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->timeline) {
        SSB_CUDA(cudaDeviceSynchronize());
        s->tl_start.clear();
        s->tl_end.clear();
        s->tl_kind.clear();
    }
    for (auto& u : s->ev_used) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->ev_pool[u.second], s->ev_pool[u.second + 1]);
        s->kind_ms[u.first] += ms;
        if (s->timeline && s->ev_timeline0) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, s->ev_timeline0, s->ev_pool[u.second]);
            cudaEventElapsedTime(&b, s->ev_timeline0, s->ev_pool[u.second + 1]);
            s->tl_start.push_back(a);
            s->tl_end.push_back(b);
            s->tl_kind.push_back(u.first);
        }
    }
    s->ev_used.clear();
    return 0;
}

// Grid rows -> 3xTF32 operand tiles in UMMA core-matrix order: [tile][hi|lo][k/4][row/8][row%8][k%4].
int build_scan_tiles(const float* S32, int G, int dpad, CleanupDev* cd) {
    const int kp = cd->kp, tr = cd->tr, part = tr * kp;
    std::vector<float> t((size_t)cd->n_tiles * 2 * part, 0.f);
    for (int g = 0; g < G; ++g) {
        const int tile = g / tr, r = g % tr;
        float* hi = &t[(size_t)tile * 2 * part];
        float* lo = hi + part;
        for (int k = 0; k < dpad; ++k) {
            const float x = S32[(size_t)g * dpad + k];
            const float h = ssb_tf32_round(x);
            const size_t off = ((size_t)(k / 4) * (tr / 8) + r / 8) * 32 + (r % 8) * 4 + k % 4;
            hi[off] = h;
            lo[off] = ssb_tf32_round(x - h);
        }
    }
    SSB_CUDA(cudaMalloc((void**)&cd->stc, t.size() * sizeof(float)));
    SSB_CUDA(cudaMemcpy(cd->stc, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// Grid rows -> per (tile, K block) operand blocks [hi | lo][KB/4][16][8][4]; also allocates the query tiles.
int build_scan_tiles_k(const float* S32, int G, int dpad, int n_groups, CleanupDev* cd) {
    const int n_kb = cd->n_kb, part = SSB_SCK_PART;
    std::vector<float> t((size_t)cd->n_tiles * n_kb * 2 * part, 0.f);
    for (int g = 0; g < G; ++g) {
        const int tile = g / 128, r = g % 128;
        for (int k = 0; k < dpad; ++k) {
            const int kb = k / SSB_SCK_KB, kk = k % SSB_SCK_KB;
            float* hi = &t[((size_t)tile * n_kb + kb) * 2 * part];
            float* lo = hi + part;
            const float x = S32[(size_t)g * dpad + k];
            const float h = ssb_tf32_round(x);
            const size_t off = ((size_t)(kk / 4) * 16 + r / 8) * 32 + (r % 8) * 4 + kk % 4;
            hi[off] = h;
            lo[off] = ssb_tf32_round(x - h);
        }
    }
    SSB_CUDA(cudaMalloc((void**)&cd->stck, t.size() * sizeof(float)));
    SSB_CUDA(cudaMemcpy(cd->stck, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t xt_floats = (size_t)((n_groups + 3) / 4) * n_kb * 2 * part;
    SSB_CUDA(cudaMalloc((void**)&cd->xt, xt_floats * sizeof(float)));
    SSB_CUDA(cudaMemset(cd->xt, 0, xt_floats * sizeof(float)));
    return 0;
}

void launch_scan_tck(cudaStream_t st, bool csr, const SsbCtx& c, const int* desc, const CleanupDev& cd, int n_groups) {
    const int quads = (n_groups + 3) / 4;
    if (csr) k_scan_xtiles<true><<<dim3(cd.n_kb, quads), 128, 0, st>>>(c, desc, cd.cx, cd.xt, cd.n_kb, n_groups, 0);
    else k_scan_xtiles<false><<<dim3(cd.n_kb, quads), 128, 0, st>>>(c, desc, cd.cx, cd.xt, cd.n_kb, n_groups, 0);
    const size_t smem = (size_t)SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float);
    k_cleanup_scan_tck<<<dim3(cd.n_chunks, quads), 320, smem, st>>>(desc, cd.stck, cd.xt, cd.pval, cd.pidx, cd.n_kb, cd.n_tiles,
                                                                  n_groups, cd.n_chunks * SSB_TOPK * 2);
}

void launch_scan_tc(cudaStream_t st, bool csr, const SsbCtx& c, const int* desc, const CleanupDev& cd, int n_groups) {
    dim3 grid(cd.n_chunks, (n_groups + 3) / 4);
    const size_t smem = (size_t)(2 * 128 + 4 * cd.tr) * cd.kp * sizeof(float);
    const int n_cand = cd.n_chunks * SSB_TOPK * 2;
#define SSB_SCAN_TC(CSR, TR) \
    k_cleanup_scan_tc<CSR, TR><<<grid, 256, smem, st>>>(c, desc, cd.stc, cd.cx, cd.pval, cd.pidx, cd.kp, cd.n_tiles, n_groups, n_cand)
    if (cd.tr == 128) {
        if (csr) SSB_SCAN_TC(true, 128);
        else SSB_SCAN_TC(false, 128);
    } else {
        if (csr) SSB_SCAN_TC(true, 64);
        else SSB_SCAN_TC(false, 64);
    }
#undef SSB_SCAN_TC
}

// candidates per trial left by the scan (the tensor-core scan keeps two lists per chunk)
int scan_n_cand(const CleanupDev& cd) { return cd.n_chunks * SSB_TOPK * ((cd.tc || cd.tck) ? 2 : 1); }

// relative near-tie band of the fp64 re-score: fp32 FFMA bound, or the 3xTF32 bound (3 * 2^-22 + accumulation)
float scan_eps_floor(const CleanupDev& cd) { return (cd.tc || cd.tck) ? 1.6e-5f : 0.f; }

template <int DP>
void launch_scan(cudaStream_t st, bool csr, const SsbCtx& c, const int* desc, const float* S, const CleanupDev& cd,
                 int dpad, int n_groups, int i_rel) {
    const int nw = 4;                                    // warps (= trial groups) per CTA
    const size_t smem = (size_t)cd.tile_rows * dpad * sizeof(float);
    dim3 grid(cd.n_chunks, (n_groups + nw - 1) / nw);
    const int n_cand = cd.n_chunks * SSB_TOPK;
    if (csr)
        k_cleanup_scan<DP, true><<<grid, 32 * nw, smem, st>>>(c, desc, S, cd.cx, cd.pval, cd.pidx, cd.rows_per_chunk,
                                                              cd.tile_rows, n_groups, n_cand, i_rel);
    else
        k_cleanup_scan<DP, false><<<grid, 32 * nw, smem, st>>>(c, desc, S, cd.cx, cd.pval, cd.pidx, cd.rows_per_chunk,
                                                               cd.tile_rows, n_groups, n_cand, i_rel);
}

void dispatch_scan(cudaStream_t st, bool csr, int dpad, int n_groups, const SsbCtx& c, const int* desc, const float* S,
                   const CleanupDev& cd, int i_rel) {
    if (cd.tc) launch_scan_tc(st, csr, c, desc, cd, n_groups);
    else if (cd.tck) launch_scan_tck(st, csr, c, desc, cd, n_groups);
    else if (dpad == 56) launch_scan<56>(st, csr, c, desc, S, cd, dpad, n_groups, i_rel);
    else if (dpad == 100) launch_scan<100>(st, csr, c, desc, S, cd, dpad, n_groups, i_rel);
    else launch_scan<0>(st, csr, c, desc, S, cd, dpad, n_groups, i_rel);
}

void scan_smem_optin() {
    const int lim = 200 * 1024;
    cudaFuncSetAttribute(k_cleanup_scan<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_cleanup_scan<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_cleanup_scan_tc<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_cleanup_scan_tc<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_cleanup_scan_tc<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_cleanup_scan_tc<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_cleanup_scan_tck, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float)));
}

// Wide ensembles of one level: one launch per (kernel flavour, width class); the items of a launch are
// passed by value.  The chunk (neurons per CTA) shrinks for very wide ensembles so that the staged
// encoder tile fits in shared memory.
template <int DP>
void launch_wide_class(ssb_sim* s, cudaStream_t st, const int* stage, bool voja, int cls, int i_rel, bool dry = false) {
    SsbItemList items;
    items.n = 0;
    int max_n = 0, max_dpad = 0, max_dims = 0, max_jn = 0;
    for (int i = 0; i < stage[3] && items.n < 15; ++i) {
        const int* d = &s->h_big[(stage[2] + i) * 16];
        const int c = d[2] == 56 ? 0 : (d[2] == 100 ? 1 : 2);
        if (c != cls || ((d[9] & 1) != 0) != voja) continue;
        items.idx[items.n++] = stage[2] + i;
        max_n = std::max(max_n, d[0]);
        max_dpad = std::max(max_dpad, d[2]);
        max_dims = std::max(max_dims, d[1]);
        max_jn = std::max(max_jn, d[11]);
    }
    if (items.n == 0) return;
    const int units = s->n_groups * items.n;          // CTAs per neuron chunk
    if (!voja) {
        // neurons per CTA: the multiple of 4 that wastes the least time on partial waves (resident CTAs per SM
        // from the occupancy calculator, since the staged encoder tile grows with the chunk)
        auto smem_of = [&](int ch) {
            return (size_t)(ch * max_dpad + ch + ch * max_jn + ch * 32 + max_dpad * 32 + max_jn * 32) * sizeof(float);
        };
        const int key = (DP << 20) | (max_n << 6) | (units & 63);
        auto it = s->wide_chunk_cache.find(key);
        int chunk = it == s->wide_chunk_cache.end() ? 0 : it->second;
        if (!chunk) {
            long long best_cost = -1;
            for (int ch = 16; ch <= 128; ch += 4) {
                if (smem_of(ch) > 96 * 1024) break;
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wide_static<DP>, 128, smem_of(ch));
                if (occ < 1) break;
                const long long ctas = (long long)((max_n + ch - 1) / ch) * units;
                const long long waves = (ctas + 148LL * occ - 1) / (148LL * occ);
                const long long cost = waves * (ch + 12);
                if (best_cost < 0 || cost < best_cost) best_cost = cost, chunk = ch;
            }
            if (!chunk) chunk = 16;
            s->wide_chunk_cache[key] = chunk;
        }
        if (dry) return;
        dim3 grid((max_n + chunk - 1) / chunk, s->n_groups, items.n);
        k_wide_static<DP><<<grid, 128, smem_of(chunk), st>>>(s->ctx, s->d_big, items, chunk, i_rel);
    } else {
        // very long encoder rows (d = 649).  Default: the CTA-cooperative kernel (two whole-neuron tiles in shared memory,
        // the input slice of each warp in registers) - as long as two tiles fit and a warp's slice is <= 96 rows (d <= 768);
        // SSB_VOJA=stream selects the per-warp ring kernel, SSB_VOJA=cta forces the cooperative one for any generic width.
        const char* vj = getenv("SSB_VOJA");
        const bool vj_stream = vj && std::string(vj) == "stream", vj_cta = vj && std::string(vj) == "cta";
        const size_t cta_smem = ((size_t)2 * max_dims * 32 + SSB_VC_NW * 32 + 32 + (size_t)max_jn * 32) * sizeof(float);
        const bool cta_fits = cta_smem <= 220 * 1024 && (max_dims + SSB_VC_NW - 1) / SSB_VC_NW <= 96;
        if (DP == 0 && cta_fits && !vj_stream && (vj_cta || (size_t)max_dims * 128 * 2 > 96 * 1024)) {
            const int key = (3 << 28) | (max_n << 6) | (units & 63);
            auto it = s->wide_chunk_cache.find(key);
            int chunk = it == s->wide_chunk_cache.end() ? 0 : it->second;
            if (!chunk) {
                cudaFuncSetAttribute(k_wide_voja_cta<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                cudaFuncSetAttribute(k_wide_voja_cta<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                const int per_unit = std::max(1, 148 / std::max(1, units));        // one CTA per SM: one wave
                chunk = std::max(2, (max_n + per_unit - 1) / per_unit);
                s->wide_chunk_cache[key] = chunk;
            }
            if (dry) return;
            dim3 grid((max_n + chunk - 1) / chunk, s->n_groups, items.n);
            if ((max_dims + SSB_VC_NW - 1) / SSB_VC_NW <= 48)
                k_wide_voja_cta<48><<<grid, 32 * (SSB_VC_NW + 1), cta_smem, st>>>(s->ctx, s->d_big, items, chunk, i_rel);
            else
                k_wide_voja_cta<96><<<grid, 32 * (SSB_VC_NW + 1), cta_smem, st>>>(s->ctx, s->d_big, items, chunk, i_rel);
            return;
        }
        if (DP == 0 && (size_t)max_dims * 128 * 2 > 96 * 1024) {
            // very long encoder rows (d = 649): stream them through per-warp rings of sub-tiles (k_wide_voja_stream)
            const size_t smem = (size_t)(max_dpad * 32 + max_jn * 32 + 8 * SSB_VS_NB * SSB_VS_SUB * 32) * sizeof(float);
            const int key = (1 << 29) | (max_n << 6) | (units & 63);
            auto it = s->wide_chunk_cache.find(key);
            int chunk = it == s->wide_chunk_cache.end() ? 0 : it->second;
            if (!chunk) {
                cudaFuncSetAttribute(k_wide_voja_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                const int per_unit = std::max(1, 148 / std::max(1, units));        // one CTA per SM: one wave
                chunk = (max_n + per_unit - 1) / per_unit;
                chunk = std::max(8, (chunk + 7) / 8 * 8);
                s->wide_chunk_cache[key] = chunk;
            }
            if (dry) return;
            dim3 grid((max_n + chunk - 1) / chunk, s->n_groups, items.n);
            k_wide_voja_stream<<<grid, 256, smem, st>>>(s->ctx, s->d_big, items, chunk, i_rel);
            return;
        }
        // tiles in flight per warp: measured on B200 (configs[1], 1 024 trials) 3 / 2 / 1 -> 65.2 / 62.7 / 62.1 us per launch,
        // step 264.9 / 262.7 / 262.6 us (profiles/r02g_perf_voja_ring_depth.log): a shallower ring leaves room for more CTAs
        int nwarps = 4, nb = 2;
        if (const char* e = getenv("SSB_VOJA_NB")) nb = std::max(1, std::min(SSB_VOJA_NB, atoi(e)));    // tuning knob
        SsbPesFuse pf;
        memset(&pf, 0, sizeof(pf));
        pf.desc = s->d_pes;
        pf.hdesc = s->d_pes_hdesc;
        pf.h = s->pes_h;
        bool fuse = false;
        for (int k = 0; k < 15; ++k) {
            pf.item[k] = (s->pes_fused && k < items.n) ? s->pes_of_big[items.idx[k]] : -1;
            fuse = fuse || pf.item[k] >= 0;
        }
        auto smem_of = [&](int nw) {
            return (size_t)(max_dpad * 32 + max_jn * 32 + nw * nb * max_dims * 32) * sizeof(float) + s->voja_pad_smem;
        };
        // very wide ensembles (d = 649: 83 KB per encoder tile): fewer tiles in flight, then fewer warps
        while (nb > 1 && smem_of(1) > 200 * 1024) --nb;
        while (nwarps > 1 && smem_of(nwarps) > 200 * 1024) nwarps >>= 1;
        // one wave: every resident CTA slot gets one contiguous neuron range of a trial group
        const int key = (1 << 30) | (DP << 20) | (max_n << 6) | (units & 63);
        auto it = s->wide_chunk_cache.find(key);
        int chunk = it == s->wide_chunk_cache.end() ? 0 : it->second;
        if (!chunk) {
            int occ = 0;
            if (fuse) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wide_voja<DP, true>, 32 * nwarps, smem_of(nwarps));
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wide_voja<DP, false>, 32 * nwarps, smem_of(nwarps));
            const int slots = 148 * std::max(1, occ);
            const int per_unit = std::max(1, slots / std::max(1, units));
            chunk = (max_n + per_unit - 1) / per_unit;
            chunk = std::max(nwarps, (chunk + nwarps - 1) / nwarps * nwarps);
            s->wide_chunk_cache[key] = chunk;
        }
        if (dry) return;
        dim3 grid((max_n + chunk - 1) / chunk, s->n_groups, items.n);
        if (fuse) k_wide_voja<DP, true><<<grid, 32 * nwarps, smem_of(nwarps), st>>>(s->ctx, s->d_big, items, chunk, i_rel, nb, pf);
        else k_wide_voja<DP, false><<<grid, 32 * nwarps, smem_of(nwarps), st>>>(s->ctx, s->d_big, items, chunk, i_rel, nb, pf);
    }
}

bool launch_wide_tc(ssb_sim* s, cudaStream_t st, const int* stage, bool dry);

// Static wide ensembles whose input is wider than the register-specialised / whole-K tensor-core kernels take (d = 649):
// encoder tiles per (128-neuron tile, 32-column K block), one X tile buffer per ensemble.  A level is served by
// k_wide_static_tck only if every static wide ensemble of the level qualifies.
int build_encode_tiles_k(ssb_sim* s) {
    s->tck_levels.assign(s->n_levels, ssb_sim::TckLevel());
    const char* e = getenv("SSB_ENCODE");
    if (e && std::string(e) == "ffma") return 0;
    const float* hW = reinterpret_cast<const float*>(s->arrays["weights"].bytes.data());
    const int quads = (s->n_groups + 3) / 4;
    std::vector<float> etk;
    size_t xt_floats = 0;
    for (int lvl = 0; lvl < s->n_levels; ++lvl) {
        const int* st = &s->h_stages[lvl * 12];
        ssb_sim::TckLevel& L = s->tck_levels[lvl];
        std::vector<int> idx;
        bool ok = true;
        int dpad0 = 0;
        for (int i = 0; i < st[3]; ++i) {
            const int b = st[2] + i;
            const int* d = &s->h_big[b * 16];
            if (d[9] & 1) continue;                               // Voja ensembles keep per-trial encoders
            if (d[2] <= 104 || d[11] > 4 || (dpad0 && d[2] != dpad0)) ok = false;
            dpad0 = d[2];
            idx.push_back(b);
        }
        if (!ok || idx.empty() || idx.size() > 15) continue;
        L.n_kb = (dpad0 + SSB_SCK_KB - 1) / SSB_SCK_KB;
        L.xt_stride = (long long)quads * L.n_kb * 2 * SSB_SCK_PART;
        L.items.n = (int)idx.size();
        std::vector<int> desc6;
        for (size_t k = 0; k < idx.size(); ++k) {
            const int* d = &s->h_big[idx[k] * 16];
            const int n = d[0], dims = d[1], dpad = d[2], enc_off = d[5];
            const int n_tiles = (n + 127) / 128;
            L.n_tiles_max = std::max(L.n_tiles_max, n_tiles);
            L.items.idx[k] = idx[k];
            L.items.e_off[k] = (long long)etk.size();
            L.items.x_off[k] = (long long)(xt_floats + k * L.xt_stride);
            etk.resize(etk.size() + (size_t)n_tiles * L.n_kb * 2 * SSB_SCK_PART, 0.f);
            float* base = &etk[L.items.e_off[k]];
            for (int nn = 0; nn < n; ++nn) {
                const int tile = nn / 128, r = nn % 128;
                for (int kk0 = 0; kk0 < dpad; ++kk0) {
                    const int kb = kk0 / SSB_SCK_KB, kk = kk0 % SSB_SCK_KB;
                    float* hi = base + ((size_t)tile * L.n_kb + kb) * 2 * SSB_SCK_PART;
                    float* lo = hi + SSB_SCK_PART;
                    const float x = hW[(size_t)enc_off + (size_t)nn * dpad + kk0];
                    const float h = ssb_tf32_round(x);
                    const size_t off = ((size_t)(kk / 4) * 16 + r / 8) * 32 + (r % 8) * 4 + kk % 4;
                    hi[off] = h;
                    lo[off] = ssb_tf32_round(x - h);
                }
            }
            desc6.insert(desc6.end(), {n, dims, dpad, 0, d[7], 0});
        }
        xt_floats += idx.size() * (size_t)L.xt_stride;
        SSB_CUDA(cudaMalloc((void**)&L.d_desc6, desc6.size() * sizeof(int)));
        SSB_CUDA(cudaMemcpy(L.d_desc6, desc6.data(), desc6.size() * sizeof(int), cudaMemcpyHostToDevice));
        L.on = true;
    }
    if (etk.empty()) return 0;
    SSB_CUDA(cudaMalloc((void**)&s->d_etk, etk.size() * sizeof(float)));
    SSB_CUDA(cudaMemcpy(s->d_etk, etk.data(), etk.size() * sizeof(float), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaMalloc((void**)&s->d_xtk, std::max<size_t>(xt_floats, 8) * sizeof(float)));
    SSB_CUDA(cudaMemset(s->d_xtk, 0, std::max<size_t>(xt_floats, 8) * sizeof(float)));
    SSB_CUDA(cudaFuncSetAttribute(k_wide_static_tck, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float))));
    return 0;
}

bool launch_wide_tck(ssb_sim* s, cudaStream_t st, const int* stage, bool dry) {
    if (!s->d_etk) return false;
    const int lvl = (int)((stage - s->h_stages.data()) / 12);
    if (lvl < 0 || lvl >= (int)s->tck_levels.size() || !s->tck_levels[lvl].on) return false;
    if (dry) return true;
    const ssb_sim::TckLevel& L = s->tck_levels[lvl];
    const int quads = (s->n_groups + 3) / 4;
    k_scan_xtiles<true><<<dim3(L.n_kb, quads, L.items.n), 128, 0, st>>>(s->ctx, L.d_desc6, nullptr, s->d_xtk + L.items.x_off[0],
                                                                       L.n_kb, s->n_groups, L.xt_stride);
    const int n_chunks = std::max(1, std::min(L.n_tiles_max, 148 / std::max(1, quads * L.items.n)));
    const size_t smem = (size_t)SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float);
    k_wide_static_tck<<<dim3(n_chunks, quads, L.items.n), 320, smem, st>>>(s->ctx, s->d_big, L.items, s->d_etk, s->d_xtk, L.n_kb);
    s->kind_launches[K_EXTRA]++;
    s->total_launches++;
    return true;
}

void launch_wide(ssb_sim* s, cudaStream_t st, const int* stage, bool voja, int i_rel, bool dry = false) {
    if (!voja && launch_wide_tc(s, st, stage, dry)) return;
    if (!voja && launch_wide_tck(s, st, stage, dry)) return;
    launch_wide_class<56>(s, st, stage, voja, 0, i_rel, dry);
    launch_wide_class<100>(s, st, stage, voja, 1, i_rel, dry);
    launch_wide_class<0>(s, st, stage, voja, 2, i_rel, dry);
}

void wide_smem_optin() {
    const int lim = 200 * 1024;
    cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_static<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_static<100>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_static<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<56, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<100, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<56, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<100, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_wide_voja<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
}

cudaEvent_t dep_event(ssb_sim* s) {
    if (s->dep_used == s->dep_pool.size()) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        s->dep_pool.push_back(e);
    }
    return s->dep_pool[s->dep_used++];
}

// `to` waits for everything enqueued on `from` so far
void stream_dep(ssb_sim* s, cudaStream_t from, cudaStream_t to) {
    if (from == to) return;
    cudaEvent_t e = dep_event(s);
    cudaEventRecord(e, from);
    cudaStreamWaitEvent(to, e, 0);
}

// Deferred PES: history arenas + per-decoder rows.  SSB_PES_DEFER=0 keeps the every-step read-modify-write kernel;
// SSB_PES_DEFER=4|8 sets the window (default 8).
int setup_pes_defer(ssb_sim* s) {
    int K = 8;
    if (const char* e = getenv("SSB_PES_DEFER")) K = atoi(e) == 4 ? 4 : 8;
    if (s->n_pes == 0) return 0;
    // neuron chunks of the sparse decode: enough CTAs (chunks x groups x column tiles) for ~3 resident CTAs per SM
    int tiles_total = 0;
    for (int i = 0; i < s->n_pes; ++i) tiles_total += (ssb_pes_jp(s->h_pes[i * 13 + 1]) + SSB_PES_JT - 1) / SSB_PES_JT;
    // CTAs = chunks x 4 trial octets x groups x column tiles: one wave of resident CTAs (every warp then walks a long
    // neuron range, whose spikes it compacts first, so the dependent round trips stay few)
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pes_defer<8>, 256, 4096);
    int chunks = (148 * std::max(1, occ)) / std::max(1, 4 * tiles_total * s->n_groups);
    chunks = std::max(1, std::min(chunks, 32));
    if (const char* e = getenv("SSB_PES_CHUNKS")) chunks = std::max(1, std::min(atoi(e), 64));
    std::vector<int> hd;
    int rows_e = 0, rows_f = 0, rows_p = 0;
    for (int i = 0; i < s->n_pes; ++i) {
        int* d = &s->h_pes[i * 13];
        d[10] = std::max(1, std::min(chunks, d[0] / 16));
        hd.insert(hd.end(), {rows_e, rows_f, rows_p, i});
        rows_e += K * d[1];
        rows_f += K * d[0];
        // split-K partials: the decode kernel's own chunks, or (fused decode) the chunks of the Voja ensemble launch
        rows_p += std::max(d[10], std::min(d[0], 592 / std::max(1, s->n_groups) + 1)) * (d[1] + K);
    }
    SSB_CUDA(cudaMemcpy(s->d_pes, s->h_pes.data(), s->h_pes.size() * sizeof(int), cudaMemcpyHostToDevice));
    SsbPesDefer& h = s->pes_h;
    h.rows_e = rows_e;
    h.rows_f = rows_f;
    h.rows_p = rows_p;
    if (alloc_rows(&h.hist_e, rows_e, s->B) || alloc_rows(&h.hist_f, rows_f, s->B) || alloc_rows(&h.part, rows_p, s->B)) return -2;
    SSB_CUDA(cudaMalloc((void**)&h.counters, (size_t)s->n_pes * s->n_groups * sizeof(int)));
    SSB_CUDA(cudaMemset(h.counters, 0, (size_t)s->n_pes * s->n_groups * sizeof(int)));
    hd.resize(hd.size() + 8, 0);
    SSB_CUDA(cudaMalloc((void**)&s->d_pes_hdesc, hd.size() * sizeof(int)));
    SSB_CUDA(cudaMemcpy(s->d_pes_hdesc, hd.data(), hd.size() * sizeof(int), cudaMemcpyHostToDevice));
    h.K = K;
    SSB_CUDA(cudaFuncSetAttribute(k_pes_fold<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SSB_PES_FS * 32 * (int)sizeof(float)));
    SSB_CUDA(cudaFuncSetAttribute(k_pes_fold<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SSB_PES_FS * 32 * (int)sizeof(float)));
    return 0;
}

void launch_pes_fold(ssb_sim* s, cudaStream_t st, int i_rel, int force) {
    // neuron chunks of the streaming fold: ~4 CTAs per SM
    int max_n = 0, max_jp = 4;
    for (int i = 0; i < s->n_pes; ++i) {
        max_n = std::max(max_n, s->h_pes[i * 13]);
        max_jp = std::max(max_jp, (s->h_pes[i * 13 + 1] + 3) & ~3);
    }
    // decoders up to 56 columns: the CTA-cooperative fold (ring of 4-neuron tiles AND their history factors by TMA, history
    // terms in registers): 85.8 us per fold on configs[1] = 5.2 TB/s, against 153 us for the warp-task kernel below, which
    // stays for wider decoders and as SSB_PES_FOLD=tasks.  (With the factors as plain loads - a dependent L2 round trip per
    // neuron - the same kernel took 194 - 231 us: profiles/r02i_perf_pes_fold_cta.log.)
    const char* fold_env = getenv("SSB_PES_FOLD");
    const bool fold_cta = !(fold_env && std::string(fold_env) == "tasks");
    if (max_jp <= 56 && fold_cta) {
        const size_t smem = (size_t)SSB_PFC_NT * SSB_PFC_NPT * 32 * (max_jp + s->pes_h.K) * sizeof(float);
        const int chunks = std::max(1, std::min(148 / std::max(1, s->n_groups * s->n_pes), (max_n + 15) / 16));
        dim3 grid(chunks, s->n_groups, s->n_pes);
        if (s->pes_h.K == 4) {
            cudaFuncSetAttribute(k_pes_fold_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_pes_fold_cta<4><<<grid, 288, smem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, chunks);
        } else {
            cudaFuncSetAttribute(k_pes_fold_cta<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_pes_fold_cta<8><<<grid, 288, smem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, chunks);
        }
        k_pes_clear<<<dim3((s->pes_h.rows_e + 3) / 4, s->n_groups), 128, 0, st>>>(s->ctx, s->pes_h, i_rel, force);
        return;
    }
    // neuron chunks: one wave of resident CTAs (occupancy of the 69 KB / 256-thread kernel), so no partial second wave
    const size_t fsm = (size_t)s->pes_h.K * SSB_PES_FS * 32 * sizeof(float);
    int occ = 0;
    if (s->pes_h.K == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pes_fold<4>, 256, fsm);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pes_fold<8>, 256, fsm);
    const int chunks = std::max(1, std::min(148 * std::max(1, occ) / std::max(1, s->n_groups * s->n_pes), (max_n + 7) / 8));
    dim3 grid(chunks, s->n_groups, s->n_pes);
    const size_t smem = (size_t)s->pes_h.K * SSB_PES_FS * 32 * sizeof(float);
    if (s->pes_h.K == 4) k_pes_fold<4><<<grid, 256, smem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, chunks, i_rel, force);
    else k_pes_fold<8><<<grid, 256, smem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, chunks, i_rel, force);
    k_pes_clear<<<dim3((s->pes_h.rows_e + 3) / 4, s->n_groups), 128, 0, st>>>(s->ctx, s->pes_h, i_rel, force);
}

// Fold the pending history into the decoders (before any host read / write of the ldec arena).
int pes_flush(ssb_sim* s) {
    if (s->pes_h.K == 0 || !s->finalized) return 0;
    launch_pes_fold(s, s->stream, 0, 1);
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

void launch_pes(ssb_sim* s, cudaStream_t st, int i_rel) {
    int max_chunks = 1, max_rows = 0, max_jt = 1, max_per = 1;
    for (int i = 0; i < s->n_pes; ++i) {
        const int* d = &s->h_pes[i * 13];
        max_chunks = std::max(max_chunks, d[10]);
        max_per = std::max(max_per, (d[0] + d[10] - 1) / d[10]);
        max_rows = std::max(max_rows, d[0] + d[1]);
        max_jt = std::max(max_jt, (ssb_pes_jp(d[1]) + SSB_PES_JT - 1) / SSB_PES_JT);
    }
    if (!s->pes_fused) {     // (fused: the sparse decode already ran inside the Voja ensemble kernel of this step)
        LaunchTimer t(s, K_PES, st);
        dim3 dgrid(max_chunks * 4, s->n_groups, s->n_pes * max_jt);
        const size_t fsmem = (size_t)max_per * 9 * sizeof(int);               // flag words of one neuron chunk + 8 spike lists
        if (s->pes_h.K == 4) k_pes_defer<4><<<dgrid, 256, fsmem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, max_jt, i_rel);
        else k_pes_defer<8><<<dgrid, 256, fsmem, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, max_jt, i_rel);
    }
    {
        LaunchTimer t(s, K_PHIST, st);
        k_pes_hist<<<dim3((max_rows + 3) / 4, s->n_groups, s->n_pes), 128, 0, st>>>(s->ctx, s->pes_h, s->d_pes, s->d_pes_hdesc, i_rel);
    }
    // the host mirrors the step counter, so the fold is launched only after the last slot of a window
    // (a captured graph bakes this in; it is replayed only from steps with the same phase, see ssb_run_steps)
    if ((int)((s->step_base + i_rel) % s->pes_h.K) == s->pes_h.K - 1) {
        LaunchTimer t(s, K_PFOLD, st);
        launch_pes_fold(s, st, i_rel, 1);
        s->kind_launches[K_EXTRA] += 1;              // k_pes_clear
        s->total_launches += 1;
    }
}

// SSB_ENCODE=tc selects the tensor-core wide-ensemble kernel.  Measured on B200 (BASELINE configs[1], 1024 trials):
// k_wide_static (FFMA, 20 warps/SM) 36.5 us; k_wide_static_tc 55 us with 8 warps and per-neuron bias loads, 45.7 us
// with bias / direct-current weights staged in shared memory, 27.4 us with 16 warps x 16 TMEM columns.  Both kernels
// are issue-bound (LIF update + addressing ~50 instructions per neuron and trial group; the dot product adds 56 FFMA),
// but the tcgen05 kernel holds a whole SM (114 KB smem, 512 threads) while the FFMA one shares SMs with the other
// dependency streams: the step changes by ~1 us (284.8 vs 286.0), inside run-to-run noise, so FFMA stays the default.
bool encode_tc_allowed() {
    const char* e = getenv("SSB_ENCODE");
    return e && std::string(e) == "tc";
}

// Static wide-ensemble encoders [n][dpad] -> 64-neuron K-major operand tiles (hi | lo) in UMMA core-matrix order.
int build_encode_tiles(ssb_sim* s) {
    const int n_big = (int)(s->h_big.size() / 16);
    s->enc_t_off.assign(n_big + 8, -1);
    if (n_big == 0 || !encode_tc_allowed()) return 0;
    const float* hW = reinterpret_cast<const float*>(s->arrays["weights"].bytes.data());
    std::vector<float> et;
    for (int i = 0; i < n_big; ++i) {
        const int* d = &s->h_big[i * 16];
        const int n = d[0], dpad = d[2], enc_off = d[5];
        const int kp = (dpad + 7) / 8 * 8;
        if ((d[9] & 1) || kp > 104 || d[11] > 4) continue;     // Voja ensembles keep per-trial encoders
        const int n_tiles = (n + SSB_ETC_N - 1) / SSB_ETC_N, part = SSB_ETC_N * kp;
        s->enc_t_off[i] = (int)et.size();
        et.resize(et.size() + (size_t)n_tiles * 2 * part, 0.f);
        float* base = &et[s->enc_t_off[i]];
        for (int nn = 0; nn < n; ++nn) {
            float* hi = base + (size_t)(nn / SSB_ETC_N) * 2 * part;
            float* lo = hi + part;
            const int r = nn % SSB_ETC_N;
            for (int k = 0; k < dpad; ++k) {
                const float x = hW[(size_t)enc_off + (size_t)nn * dpad + k];
                const float h = ssb_tf32_round(x);
                const size_t o = ((size_t)(k / 4) * 8 + r / 8) * 32 + (r % 8) * 4 + k % 4;
                hi[o] = h;
                lo[o] = ssb_tf32_round(x - h);
            }
        }
    }
    et.resize(et.size() + 8, 0.f);
    SSB_CUDA(cudaMalloc((void**)&s->d_enc_t, et.size() * sizeof(float)));
    SSB_CUDA(cudaMemcpy(s->d_enc_t, et.data(), et.size() * sizeof(float), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaMalloc((void**)&s->d_enc_t_off, s->enc_t_off.size() * sizeof(int)));
    SSB_CUDA(cudaMemcpy(s->d_enc_t_off, s->enc_t_off.data(), s->enc_t_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaFuncSetAttribute(k_wide_static_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    SSB_CUDA(cudaFuncSetAttribute(k_wide_static_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    return 0;
}

// Static wide ensembles of a level on the tensor-core kernel: one launch per operand width.  Returns false if
// some ensemble of the level is not eligible (the caller then uses the FFMA kernels for the whole level).
bool launch_wide_tc(ssb_sim* s, cudaStream_t st, const int* stage, bool dry) {
    if (!s->d_enc_t) return false;
    std::map<int, SsbItemList> by_kp;
    std::map<int, int> max_n;
    for (int i = 0; i < stage[3]; ++i) {
        const int idx = stage[2] + i;
        const int* d = &s->h_big[idx * 16];
        if (d[9] & 1) continue;
        if (s->enc_t_off[idx] < 0) return false;
        const float* nt = reinterpret_cast<const float*>(s->arrays["ntypes"].bytes.data()) + (size_t)d[8] * 8;
        const int kp = ((d[2] + 7) / 8 * 8) * 2 + (nt[5] != 0.f ? 1 : 0);       // launch class: operand width, LIF variant
        SsbItemList& L = by_kp[kp];
        if (by_kp.count(kp) == 1 && max_n.count(kp) == 0) L.n = 0;
        if (L.n >= 15) return false;
        L.idx[L.n++] = idx;
        max_n[kp] = std::max(max_n[kp], d[0]);
    }
    if (dry) return true;
    for (auto& kv : by_kp) {
        const int kp = kv.first >> 1;
        const bool fast = kv.first & 1;
        const SsbItemList& L = kv.second;
        const int quads = (s->n_groups + 3) / 4;
        const int n_tiles = (max_n[kv.first] + SSB_ETC_N - 1) / SSB_ETC_N;
        const size_t smem = (size_t)(2 * 128 + 4 * SSB_ETC_N) * kp * sizeof(float);
        const int per_sm = smem <= 110 * 1024 ? 2 : 1;
        const int chunks_wanted = std::max(1, 148 * per_sm / std::max(1, quads * L.n));
        const int tpc = std::max(1, (n_tiles + chunks_wanted - 1) / chunks_wanted);
        dim3 grid((n_tiles + tpc - 1) / tpc, quads, L.n);
        if (fast) k_wide_static_tc<true><<<grid, 512, smem, st>>>(s->ctx, s->d_big, L, s->d_enc_t, s->d_enc_t_off, kp, tpc);
        else k_wide_static_tc<false><<<grid, 512, smem, st>>>(s->ctx, s->d_big, L, s->d_enc_t, s->d_enc_t_off, kp, tpc);
    }
    return true;
}

// SSB_DECODE=ffma forces the FFMA decoder kernel (measured comparison); default is tcgen05 when the output fits 64 columns.
bool decode_tc_allowed() {
    const char* e = getenv("SSB_DECODE");
    return !(e && std::string(e) == "ffma");
}

// Static decoders [n][jpad] -> per K stage: Wd^T as N x KS K-major operand tiles (hi | lo) in UMMA core-matrix order.
// (N, KS) = (64, 64) when every decoder of the plan fits 64 columns, else (128, 32) up to 128 columns.
int build_decode_tiles(ssb_sim* s) {
    const int n_dec = (int)(s->h_dec.size() / 9);
    s->dec_tc_level.assign(s->n_levels, 0);
    if (const char* e = getenv("SSB_DECODE")) s->dec_sparse = std::string(e) == "sparse" && !s->per_trial;
    if (s->dec_sparse) {      // every decoder of the plan must not be split (the sparse walk writes the rows itself)
        for (int i = 0; i < n_dec; ++i) s->h_dec[i * 9 + 6] = 1;
        return 0;
    }
    if (n_dec == 0 || !decode_tc_allowed() || s->per_trial) return 0;      // per-trial decoders: no shared GEMM operand
    int max_jpad = 0;
    for (int i = 0; i < n_dec; ++i) max_jpad = std::max(max_jpad, s->h_dec[i * 9 + 2]);
    const int N = max_jpad <= 64 ? 64 : 128, KS = max_jpad <= 64 ? 64 : 32;
    s->dec_tc_n = N;
    s->dec_tc_nt = (max_jpad + N - 1) / N;              // column tiles per decoder (1 up to 128 columns; 6 for d = 649)
    const int n_nt = s->dec_tc_nt;
    const float* hW = reinterpret_cast<const float*>(s->arrays["weights"].bytes.data());
    std::vector<float> wt;
    std::vector<int> off(n_dec + 8, -1);
    const int part = N * KS;
    for (int i = 0; i < n_dec; ++i) {
        const int* d = &s->h_dec[i * 9];
        const int n = d[0], jpad = d[2], w_off = d[4];
        const int n_stages = (n + KS - 1) / KS;
        off[i] = (int)wt.size();
        wt.resize(wt.size() + (size_t)n_nt * n_stages * 2 * part, 0.f);
        float* base = &wt[off[i]];
        for (int k = 0; k < n; ++k) {
            const int st = k / KS, kk = k % KS;
            for (int jg = 0; jg < jpad; ++jg) {
                const int j = jg % N;
                float* hi = base + ((size_t)(jg / N) * n_stages + st) * 2 * part;
                float* lo = hi + part;
                const float x = hW[(size_t)w_off + (size_t)k * jpad + jg];
                const float h = ssb_tf32_round(x);
                const size_t o = ((size_t)(kk / 4) * (N / 8) + j / 8) * 32 + (j % 8) * 4 + kk % 4;
                hi[o] = h;
                lo[o] = ssb_tf32_round(x - h);
            }
        }
    }
    for (int lvl = 0; lvl < s->n_levels; ++lvl) s->dec_tc_level[lvl] = s->h_stages[lvl * 12 + 5] > 0 ? 1 : 0;
    wt.resize(wt.size() + 8, 0.f);
    SSB_CUDA(cudaMalloc((void**)&s->d_dec_wt, wt.size() * sizeof(float)));
    SSB_CUDA(cudaMemcpy(s->d_dec_wt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaMalloc((void**)&s->d_dec_wt_off, off.size() * sizeof(int)));
    SSB_CUDA(cudaMemcpy(s->d_dec_wt_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaFuncSetAttribute(k_decode_tc<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SSB_CUDA(cudaFuncSetAttribute(k_decode_tc<128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return 0;
}

// One k_lin launch = the dense items of a segment + its packed records + its remaining CSR rows, each x G groups.
void launch_lin(ssb_sim* s, cudaStream_t st, int seg, int i_rel) {
    const ssb_sim::LinSeg& L = s->lin_segs[seg];
    const int G = s->n_groups;
    const int rec_per_cta = 4 * SSB_REC_PER_WARP;
    const long long blocks = ((long long)L.n_items + (L.n_recs + rec_per_cta - 1) / rec_per_cta + (L.n_csr + 3) / 4) * G;
    if (L.n_tc > 0) {      // the segment's large dense blocks: gather the source rows into operand tiles, then the tcgen05 GEMM
        const int quads = (G + 3) / 4;
        k_lin_xtiles<<<dim3(L.tc_max_kb, quads, L.n_tc), 128, 0, st>>>(s->ctx, s->d_lin_tc + L.tc0, s->d_dense_cols, s->d_lin_xt, i_rel);
        const int n_chunks = std::max(1, std::min(L.tc_max_tiles, 148 / std::max(1, quads * L.n_tc)));
        const size_t smem = (size_t)SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float);
        k_lin_tck<<<dim3(n_chunks, quads, L.n_tc), 320, smem, st>>>(s->ctx, s->d_lin_tc + L.tc0, s->d_lin_ttk, s->d_lin_xt,
                                                                   s->d_dense_rows, i_rel);
        s->kind_launches[K_EXTRA] += 2;
        s->total_launches += 2;
    }
    if (blocks <= 0) return;
    SsbLinArgs a;
    a.rows = s->d_lin_rows + (size_t)L.csr_row0 * 5;
    a.ab = s->d_lin_ab + (size_t)L.csr_row0 * 2;
    a.n_rows = L.n_csr;
    a.items = s->d_dense_items + (size_t)L.item0 * 8;
    a.n_items = L.n_items;
    a.ddesc = s->d_dense_desc;
    a.dT = s->d_dense_T;
    a.dcols = s->d_dense_cols;
    a.drows = s->d_dense_rows;
    a.recs = s->d_lin_recs + (size_t)L.rec0 * 32;
    a.n_recs = L.n_recs;
    const long long waves6 = (blocks + 148 * 6 - 1) / (148 * 6), waves8 = (blocks + 148 * 8 - 1) / (148 * 8);
    if (waves8 < waves6 && !s->lin_six) k_lin<8><<<(unsigned)blocks, 128, 0, st>>>(s->ctx, a, i_rel);
    else k_lin<6><<<(unsigned)blocks, 128, 0, st>>>(s->ctx, a, i_rel);
}

// Split the row program of every launch segment into dense blocks and CSR rows.  Rows (of one view) with an
// identical column list and at least SSB_DENSE_MIN_K entries form a block when there are >= SSB_DENSE_MIN_R of them.
#define SSB_DENSE_MIN_K 16
#define SSB_DENSE_MIN_R 8
int build_lin_program(ssb_sim* s) {
    const std::vector<int> rows3 = host_ints(s, "lin_rows"), ptr = host_ints(s, "csr_ptr");
    const std::vector<int> e0 = host_ints(s, "csr_ent0"), e1 = host_ints(s, "csr_ent1");   // (row, coefficient bits) pairs
    const float* ab = s->arrays.count("lin_ab") ? reinterpret_cast<const float*>(s->arrays["lin_ab"].bytes.data()) : nullptr;
    const size_t n_rows = rows3.size() / 3;
    if ((size_t)(s->lin0 + s->n_lin + s->n_lin_fused) != n_rows) return fail(-1, "ssb_finalize: lin_rows segments do not add up");
    std::vector<int> rows5, items, ddesc, dcols, drows, recs;
    std::vector<float> ab2, dT, ttk;
    std::vector<SsbLinTcBlock> tcb;
    size_t xt_floats = 0;
    const char* lin_env = getenv("SSB_LIN");
    const bool lin_tc_on = !(lin_env && std::string(lin_env) == "ffma");
    s->lin_segs.assign(s->n_levels + 4, ssb_sim::LinSeg());
    if (s->n_lin_early < 0 || s->n_lin_early > s->n_lin) return fail(-1, "ssb_finalize: n_lin_early out of range");
    if (s->n_lin_fused > 0 && (s->n_levels < 1 || s->n_lvl0_res < 0 || s->n_lvl0_res > s->h_stages[11]))
        return fail(-1, "ssb_finalize: fused row program without a level 0");
    for (int seg = 0; seg <= s->n_levels + 3; ++seg) {
        // segments after the levels: n_levels = early end-of-step rows [lin0, + n_lin_early); n_levels + 1 = the other
        // end-of-step rows; n_levels + 2 = those AND the next step's level-0 rows (step fusion); n_levels + 3 = level 0's
        // previous-view rows, the residual first launch of a step whose level-0 rows were evaluated by the previous step
        int r0, nr;
        if (seg < s->n_levels) r0 = s->h_stages[seg * 12 + 10], nr = s->h_stages[seg * 12 + 11];
        else if (seg == s->n_levels) r0 = s->lin0, nr = s->n_lin_early;
        else if (seg == s->n_levels + 1) r0 = s->lin0 + s->n_lin_early, nr = s->n_lin - s->n_lin_early;
        else if (seg == s->n_levels + 2) r0 = s->lin0 + s->n_lin_early, nr = s->n_lin_fused > 0 ? s->n_lin - s->n_lin_early + s->n_lin_fused : 0;
        else r0 = s->n_lin_fused > 0 ? s->h_stages[10] + s->h_stages[11] - s->n_lvl0_res : 0, nr = s->n_lin_fused > 0 ? s->n_lvl0_res : 0;
        ssb_sim::LinSeg& L = s->lin_segs[seg];
        L.csr_row0 = (int)(rows5.size() / 5);
        L.item0 = (int)(items.size() / 8);
        L.tc0 = (int)tcb.size();
        // group candidate rows by (view, column list)
        std::map<std::vector<int>, std::vector<int>> groups;
        for (int r = r0; r < r0 + nr; ++r) {
            const int src = rows3[r * 3], kind = rows3[r * 3 + 1];
            if (kind == 2 || kind == 5) continue;            // activity rows (trace / neuron probe) read the act arena
            if (src < 0 || (size_t)src + 1 >= ptr.size()) return fail(-1, "ssb_finalize: lin_rows CSR row out of range");
            const int lo = ptr[src], hi = ptr[src + 1];
            if (hi - lo < SSB_DENSE_MIN_K) continue;
            std::vector<int> key;
            key.reserve(hi - lo + 1);
            key.push_back(kind == 4 ? 1 : 0);
            for (int p = lo; p < hi; ++p) key.push_back(e0[(size_t)p * 2]);
            groups[key].push_back(r);
        }
        std::vector<char> is_dense(nr, 0);
        for (auto& kv : groups) {
            const std::vector<int>& members = kv.second;
            if ((int)members.size() < SSB_DENSE_MIN_R) continue;
            const int K = (int)kv.first.size() - 1;
            const int kpad = (K + SSB_DENSE_SLAB - 1) / SSB_DENSE_SLAB * SSB_DENSE_SLAB;
            const int R = (int)members.size();
            const bool tc_block = lin_tc_on && R >= 128 && K >= 256;      // a real GEMM: K-blocked tcgen05 path
            const int block = (int)(ddesc.size() / 8);
            const int t_off = (int)dT.size(), cols_off = (int)dcols.size(), rows_off = (int)(drows.size() / 4);
            const int lo0 = ptr[rows3[members[0] * 3]];
            for (int par = 0; par < 2; ++par)
                for (int k = 0; k < kpad; ++k) dcols.push_back(k < K ? (par ? e1 : e0)[(size_t)(lo0 + k) * 2] : 0);
            for (int m : members) {
                const int lo = ptr[rows3[m * 3]];
                for (int k = 0; k < kpad; ++k) {
                    float v = 0.f;
                    if (k < K) memcpy(&v, &e0[(size_t)(lo + k) * 2 + 1], 4);
                    dT.push_back(v);
                }
                int abits = 0, bbits = 0;
                if (ab) {
                    memcpy(&abits, &ab[(size_t)m * 2], 4);
                    memcpy(&bbits, &ab[(size_t)m * 2 + 1], 4);
                }
                drows.insert(drows.end(), {rows3[m * 3 + 1], rows3[m * 3 + 2], abits, bbits});
                is_dense[m - r0] = 1;
            }
            ddesc.insert(ddesc.end(), {R, kpad, t_off, cols_off, rows_off, 0, 0, 0});
            (void)block;
            if (tc_block) {
                SsbLinTcBlock b;
                b.R = R;
                b.K = K;
                b.n_kb = (K + SSB_SCK_KB - 1) / SSB_SCK_KB;
                b.n_tiles = (R + 127) / 128;
                b.cols_off = cols_off;
                b.kpad = kpad;
                b.view = kv.first[0];
                b.rows_off = rows_off;
                b.t_off = (long long)ttk.size();
                b.x_off = (long long)xt_floats;
                ttk.resize(ttk.size() + (size_t)b.n_tiles * b.n_kb * 2 * SSB_SCK_PART, 0.f);
                float* base = &ttk[b.t_off];
                for (int r = 0; r < R; ++r) {
                    const int tile = r / 128, rr = r % 128;
                    for (int k = 0; k < K; ++k) {
                        const int kb = k / SSB_SCK_KB, kk = k % SSB_SCK_KB;
                        float* hi = base + ((size_t)tile * b.n_kb + kb) * 2 * SSB_SCK_PART;
                        float* lo = hi + SSB_SCK_PART;
                        const float x = dT[(size_t)t_off + (size_t)r * kpad + k];
                        const float h = ssb_tf32_round(x);
                        const size_t off = ((size_t)(kk / 4) * 16 + rr / 8) * 32 + (rr % 8) * 4 + kk % 4;
                        hi[off] = h;
                        lo[off] = ssb_tf32_round(x - h);
                    }
                }
                xt_floats += (size_t)((s->n_groups + 3) / 4) * b.n_kb * 2 * SSB_SCK_PART;
                tcb.push_back(b);
                L.n_tc++;
                L.tc_max_kb = std::max(L.tc_max_kb, b.n_kb);
                L.tc_max_tiles = std::max(L.tc_max_tiles, b.n_tiles);
            } else {
                for (int row0 = 0; row0 < R; row0 += SSB_DENSE_RCH)
                    items.insert(items.end(), {t_off + row0 * kpad, cols_off, kpad, rows_off + row0,
                                               std::min(SSB_DENSE_RCH, R - row0), kv.first[0], 0, 0});
            }
            s->n_dense_rows += R;
            s->n_dense_blocks++;
        }
        L.rec0 = (int)(recs.size() / 32);
        for (int r = r0; r < r0 + nr; ++r) {
            if (is_dense[r - r0]) continue;
            const int src = rows3[r * 3], kind = rows3[r * 3 + 1];
            if (kind == 2 && s->pes_h.K > 0) continue;      // deferred PES: k_pes_hist updates the activity traces
            const bool from_act = kind == 2 || kind == 5;
            const int lo = from_act ? 0 : ptr[src], hi = from_act ? 0 : ptr[src + 1];
            const float fa = ab ? ab[(size_t)r * 2] : 0.f, fb = ab ? ab[(size_t)r * 2 + 1] : 1.f;
            if (hi - lo <= 8 && kind != 5) {              // one 128-byte record
                int w[32] = {0};
                w[0] = kind;
                w[1] = rows3[r * 3 + 2];
                memcpy(&w[2], &fa, 4);
                memcpy(&w[3], &fb, 4);
                for (int e = 0; e < hi - lo; ++e) {
                    w[4 + e] = e0[(size_t)(lo + e) * 2];
                    w[12 + e] = e1[(size_t)(lo + e) * 2];
                    w[20 + e] = e0[(size_t)(lo + e) * 2 + 1];
                }
                w[28] = src;
                recs.insert(recs.end(), w, w + 32);
                continue;
            }
            rows5.insert(rows5.end(), {src, kind, rows3[r * 3 + 2], lo, hi});
            ab2.push_back(fa);
            ab2.push_back(fb);
        }
        L.n_recs = (int)(recs.size() / 32) - L.rec0;
        L.n_csr = (int)(rows5.size() / 5) - L.csr_row0;
        L.n_items = (int)(items.size() / 8) - L.item0;
    }
    auto up_i = [&](std::vector<int>& v, int** dst) {
        v.resize(v.size() + 8, 0);
        if (cudaMalloc((void**)dst, v.size() * sizeof(int)) != cudaSuccess) return 1;
        return cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ? 1 : 0;
    };
    auto up_f = [&](std::vector<float>& v, float** dst) {
        v.resize(v.size() + 8, 0.f);
        if (cudaMalloc((void**)dst, v.size() * sizeof(float)) != cudaSuccess) return 1;
        return cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ? 1 : 0;
    };
    if (up_i(rows5, &s->d_lin_rows) || up_f(ab2, &s->d_lin_ab) || up_i(items, &s->d_dense_items) || up_i(ddesc, &s->d_dense_desc) ||
        up_i(dcols, &s->d_dense_cols) || up_i(drows, &s->d_dense_rows) || up_f(dT, &s->d_dense_T) || up_i(recs, &s->d_lin_recs))
        return fail(-2, "ssb_finalize: row program upload failed");
    if (!tcb.empty()) {
        SSB_CUDA(cudaMalloc((void**)&s->d_lin_tc, tcb.size() * sizeof(SsbLinTcBlock)));
        SSB_CUDA(cudaMemcpy(s->d_lin_tc, tcb.data(), tcb.size() * sizeof(SsbLinTcBlock), cudaMemcpyHostToDevice));
        SSB_CUDA(cudaMalloc((void**)&s->d_lin_ttk, ttk.size() * sizeof(float)));
        SSB_CUDA(cudaMemcpy(s->d_lin_ttk, ttk.data(), ttk.size() * sizeof(float), cudaMemcpyHostToDevice));
        SSB_CUDA(cudaMalloc((void**)&s->d_lin_xt, xt_floats * sizeof(float)));
        SSB_CUDA(cudaMemset(s->d_lin_xt, 0, xt_floats * sizeof(float)));
        SSB_CUDA(cudaFuncSetAttribute(k_lin_tck, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(SSB_SCK_NST * 4 * SSB_SCK_PART * sizeof(float))));
    }
    return 0;
}

// One simulator step = this launch sequence; `i_rel` is the step's offset from the device-side step
// counter dyn[0], which ssb_run_steps advances once per call (or once per replayed graph).
//
// Dependency streams (parallel branches of the captured graph):
//   A  k_begin, the level's materialised rows (k_lin), narrow ensembles, final k_lin
//   B  Voja-learned wide ensembles -> k_pes          (the two HBM-heavy streaming kernels)
//   C  grid clean-up scan -> pick, gate
//   D  static wide ensembles -> static decoders
// A level's successor waits for C and D only: nothing inside a step reads what k_pes writes (its outputs
// reach consumers through Lowpass filters, i.e. the final k_lin), so B joins A just before that launch.
// The step's input rows (parity copy of step dyn[0] + i_rel): copied from the resident tables or synthesised.
void launch_inputs(ssb_sim* s, cudaStream_t st, int i_rel) {
    const SsbCtx& c = s->ctx;
    const int G = s->n_groups;
    if (s->synth_on) {
        LaunchTimer t(s, K_BEGIN, st);
        const int d = s->synth.d;
        const size_t small_smem = (size_t)(2 * d * 32 + s->synth.n_lm * d + s->synth.n_lm * s->synth.dim * 32) * sizeof(float) +
                                  (size_t)s->synth.n_lm * 32 + 32;
        if (d <= 128 && s->synth.n_lm < 255 && small_smem <= 200 * 1024) k_synth<true><<<G, 256, small_smem, st>>>(c, s->synth, i_rel);
        else k_synth<false><<<G, 256, (size_t)2 * d * 32 * sizeof(float), st>>>(c, s->synth, i_rel);
    } else if (s->nt > 0) {
        LaunchTimer t(s, K_BEGIN, st);
        dim3 grid(((int)s->nt + 3) / 4, G);
        k_begin<<<grid, 128, 0, st>>>(c, i_rel);
    }
}

// `have_inputs`: the previous step of this launch sequence already produced this step's input rows;
// `prefetch_next`: produce the next step's input rows on stream C while this step runs (they go to the other parity copy).
//
// Level scheduling.  The lowering's level_deps[L][Lp] says which producer kinds of level Lp the sink rows of level L read
// (bit 0 narrow ensembles, 1 static decoders, 2 clean-up nodes, 3 gate nodes).  The chain of a level (its k_lin rows and
// its narrow ensembles) runs on stream A when it needs the narrow ensembles of an earlier level (stream order then
// carries that dependency), otherwise on stream E, waiting only for the events of the producers it really reads — in
// SLAM the landmark circular convolution (level 1) needs the clean-up index and the OVC decode, not the 14 000 VCO
// neurons of level 0, so it overlaps them instead of queueing behind them.
int one_step(ssb_sim* s, int i_rel, bool have_inputs = false, bool prefetch_next = false) {
    const SsbCtx& c = s->ctx;
    const int G = s->n_groups;
    const bool par = s->parallel && (!s->profiling || s->timeline) && !s->debug_sync;
    cudaStream_t A = s->stream, B = par ? s->aux[0] : A, C = par ? s->aux[1] : A, D = par ? s->aux[2] : A;
    cudaStream_t E = par ? s->aux[3] : A, F = par ? s->aux[4] : A, Gs = par ? s->aux[5] : A;
    bool early_used = false;
    const bool any_inputs = s->synth_on || s->nt > 0;
    if (any_inputs && !(have_inputs && par)) launch_inputs(s, A, i_rel);
    const bool prefetch = any_inputs && prefetch_next && par;
    if (prefetch) {                 // everything of the previous step is behind A here, so the other copy is free
        // its own stream: on C the clean-up chain queued behind it (k_synth: one CTA per trial group, ~35 us of latency)
        stream_dep(s, A, F);
        launch_inputs(s, F, i_rel + 1);
    }
    const int NL = s->n_levels;
    // events of the producers of every level (null = that level has no such producer)
    std::vector<cudaEvent_t> ev_small(NL, nullptr), ev_dec(NL, nullptr), ev_pick(NL, nullptr), ev_gate(NL, nullptr);
    auto mark = [&](cudaStream_t st) {
        cudaEvent_t e = dep_event(s);
        cudaEventRecord(e, st);
        return e;
    };
    auto wait_on = [&](cudaStream_t st, cudaEvent_t e) {
        if (e) cudaStreamWaitEvent(st, e, 0);
    };
    bool pes_done = s->n_pes == 0, b_used = false, c_used = false, d_used = false, e_used = false;
    // step fusion: a step whose predecessor in this batch ran the fused end-of-step launch finds its level-0 sink rows
    // materialised; only the previous-view rows (PES error) are left, and only the PES chain waits for them
    const bool fused_prev = par && have_inputs && s->n_lin_fused > 0;
    const bool fuse_next = par && prefetch_next && s->n_lin_fused > 0;
    cudaEvent_t ev_res = nullptr;
    for (int lvl = 0; lvl < NL; ++lvl) {
        const int* st = &s->h_stages[lvl * 12];
        const LevelInfo& li = s->levels[lvl];
        // which stream carries this level's chain, and what it waits for
        cudaStream_t T = A;
        if (lvl > 0 && par) {
            int need_small = s->level_deps.empty() ? 1 : 0, any = 0;
            for (int lp = 0; lp < lvl && !s->level_deps.empty(); ++lp) {
                need_small |= s->level_deps[(size_t)lvl * NL + lp] & 1;
                any |= s->level_deps[(size_t)lvl * NL + lp];
            }
            // (every event E waits for is downstream of this step's first k_lin on A, which orders E after the previous step)
            if (!need_small && any) T = E;
            for (int lp = 0; lp < lvl; ++lp) {
                const int m = s->level_deps.empty() ? 15 : s->level_deps[(size_t)lvl * NL + lp];
                if (m & 1) wait_on(T, ev_small[lp]);
                if (m & 2) wait_on(T, ev_dec[lp]);
                if (m & 4) wait_on(T, ev_pick[lp]);
                if (m & 8) wait_on(T, ev_gate[lp]);
            }
            if (T == E) e_used = true;
        }
        if (lvl == 0 && fused_prev) {
            if (s->n_lvl0_res > 0) {
                stream_dep(s, A, Gs);
                LaunchTimer t(s, K_LIN, Gs);
                launch_lin(s, Gs, s->n_levels + 3, i_rel);
                ev_res = mark(Gs);
                early_used = true;
            }
        } else if (st[11] > 0) {   // materialise this level's sink rows (ensemble / node inputs, PES errors)
            LaunchTimer t(s, K_LIN, T);
            launch_lin(s, T, lvl, i_rel);
        }
        const bool pes_here = !pes_done && lvl == s->pes_level;
        const bool useB = li.n_voja > 0 || pes_here;
        const bool useC = st[7] > 0 || st[9] > 0;
        const bool useD = li.n_static > 0 || st[5] > 0;
        // the clean-up chain goes first: its one-CTA-per-SM scan must not queue behind kernels that fill the SMs
        if (useC) stream_dep(s, T, C);
        for (int i = 0; i < st[7]; ++i) {
            const int ci = st[6] + i;
            const int* d = &s->h_cleanup[ci * 6];
            const CleanupDev& cd = s->cleanups[ci];
            {
                LaunchTimer t(s, K_SCAN, C);
                dispatch_scan(C, true, d[2], G, c, s->d_cleanup + ci * 6, s->d_W + d[3], cd, i_rel);
                if (cd.tck) {                                    // + k_scan_xtiles
                    s->kind_launches[K_EXTRA]++;
                    s->total_launches++;
                }
            }
            {
                LaunchTimer t(s, K_PICK, C);
                k_cleanup_pick<<<G, 256, 0, C>>>(d[1], d[2], scan_n_cand(cd), cd.cx, cd.pval, cd.pidx, cd.s64,
                                                 s->d_W + d[3], s->vec, (int)s->nv, d[5], cd.idx, nullptr, 0, 0,
                                                 scan_eps_floor(cd));
            }
        }
        if (st[7] > 0) ev_pick[lvl] = mark(C);
        if (st[9] > 0) {
            LaunchTimer t(s, K_GATE, C);
            dim3 grid(G, st[9]);
            k_gate<<<grid, 256, 0, C>>>(c, s->d_gate, st[8], i_rel);
            ev_gate[lvl] = mark(C);
        }
        if (useD) stream_dep(s, T, D);
        if (useB) stream_dep(s, T, B);
        b_used = b_used || useB;
        c_used = c_used || useC;
        d_used = d_used || useD;
        if (li.n_static > 0) {
            LaunchTimer t(s, K_WIDE, D);
            launch_wide(s, D, st, false, i_rel);
        }
        if (li.n_voja > 0) {
            LaunchTimer t(s, K_VOJA, B);
            launch_wide(s, B, st, true, i_rel);
        }
        if (pes_here) {   // every PES pre-ensemble has produced its activities
            wait_on(B, ev_res);
            if (s->pes_needs_static && li.n_static > 0) stream_dep(s, D, B);
            launch_pes(s, B, i_rel);
            pes_done = true;
        }
        if (st[5] > 0) {
            if (li.dec_needs_voja) stream_dep(s, B, D);   // recorded after the Voja kernel (and k_pes, if any)
            LaunchTimer t(s, K_DEC, D);
            int max_chunks = 1;
            size_t smem = 0;
            for (int i = 0; i < st[5]; ++i) {
                const int* d = &s->h_dec[(st[4] + i) * 9];
                max_chunks = std::max(max_chunks, d[6]);
                const size_t per = (d[0] + d[6] - 1) / d[6];
                smem = std::max(smem, (per * d[2] + 4 * per * 32) * sizeof(float));
            }
            dim3 grid(max_chunks, (G + 3) / 4, st[5]);
            if (s->dec_tc_level[lvl]) grid.z = st[5] * s->dec_tc_nt;
            if (s->per_trial) {
                int max_jp = 4;
                for (int i = 0; i < st[5]; ++i) max_jp = std::max(max_jp, (s->h_dec[(st[4] + i) * 9 + 1] + 3) & ~3);
                const int n_jt = (max_jp + SSB_PES_JT - 1) / SSB_PES_JT;
                k_decode_pt<false><<<dim3(4, G, st[5] * n_jt), 256, 0, D>>>(c, s->d_dec, st[4], n_jt);
            } else if (s->dec_sparse) {
                int max_jp = 8;
                for (int i = 0; i < st[5]; ++i) max_jp = std::max(max_jp, s->h_dec[(st[4] + i) * 9 + 2]);
                const int n_jt = (max_jp + SSB_PES_JT - 1) / SSB_PES_JT;
                k_decode_pt<true><<<dim3(4, G, st[5] * n_jt), 256, 0, D>>>(c, s->d_dec, st[4], n_jt);
            } else if (s->dec_tc_level[lvl] && s->dec_tc_n == 64)
                k_decode_tc<64, 64><<<grid, 256, (size_t)(4 * 128 * 64 + 4 * 64 * 64) * sizeof(float), D>>>(
                    c, s->d_dec, st[4], s->d_dec_wt, s->d_dec_wt_off, s->dec_tc_nt);
            else if (s->dec_tc_level[lvl])
                k_decode_tc<128, 32><<<grid, 256, (size_t)(4 * 128 * 32 + 4 * 128 * 32) * sizeof(float), D>>>(
                    c, s->d_dec, st[4], s->d_dec_wt, s->d_dec_wt_off, s->dec_tc_nt);
            else
                k_decode<<<grid, 128, smem, D>>>(c, s->d_dec, st[4]);
            ev_dec[lvl] = mark(D);
        }
        if (st[1] > 0) {
            LaunchTimer t(s, K_SMALL, T);
            // items are sorted by neuron count (descending): the leading ones get a whole CTA per trial group
            int n_split = 0;
            while (n_split < st[1] && s->h_small[(st[0] + n_split) * 9] >= 128) ++n_split;
            const int packed_warps = (st[1] - n_split) * G;
            const int blocks = n_split * G + (packed_warps + 3) / 4;
            if (s->wpt) k_ens_small_pt<<<blocks, 128, 0, T>>>(c, s->d_small + st[0] * 9, st[1], n_split);
            else k_ens_small<<<blocks, 128, 0, T>>>(c, s->d_small + st[0] * 9, st[1], n_split);
            ev_small[lvl] = mark(T);
        }
        if (lvl == 0 && s->n_lin_early > 0) {
            // end-of-step rows that need nothing but the step-start columns and level 0's narrow ensembles (the VCO filters):
            // on their own stream, next to the other chains, instead of in the launch that waits for everything
            if (par) {
                cudaEvent_t e = ev_small[0] ? ev_small[0] : mark(T);
                wait_on(Gs, e);
            }
            LaunchTimer t(s, K_LIN, Gs);
            launch_lin(s, Gs, s->n_levels, i_rel);
            early_used = par;
        }
    }
    // the end-of-step rows read everything: join every stream that was used
    if (prefetch) stream_dep(s, F, A);
    if (c_used) stream_dep(s, C, A);
    if (d_used) stream_dep(s, D, A);
    if (e_used) stream_dep(s, E, A);
    if (b_used) stream_dep(s, B, A);
    if (early_used) stream_dep(s, Gs, A);
    if (!pes_done) launch_pes(s, A, i_rel);
    if (fuse_next) {            // end-of-step rows + the next step's level-0 rows in one launch (its tables were prefetched on F)
        LaunchTimer t(s, K_LIN, A);
        launch_lin(s, A, s->n_levels + 2, i_rel);
    } else if (s->n_lin - s->n_lin_early > 0) {
        LaunchTimer t(s, K_LIN, A);
        launch_lin(s, A, s->n_levels + 1, i_rel);
    }
    return 0;
}

void advance(ssb_sim* s, int n) {
    LaunchTimer t(s, K_ADV);
    k_advance<<<1, 1, 0, s->stream>>>(s->dyn, n);
}

// Capture `n` consecutive steps (+ one counter advance) into an executable graph.  Step parity, table
// row and probe row are derived on the device from dyn[], so one graph is valid for any starting step.
int build_graph(ssb_sim* s, int n) {
    cudaGraph_t g = nullptr;
    const bool prof = s->profiling;
    s->profiling = false;
    long long saved[K_NKINDS];
    memcpy(saved, s->kind_launches, sizeof(saved));
    const long long saved_total = s->total_launches;
    s->dep_used = 0;
    s->step_base = s->steps_done;
    s->graph_phase = s->pes_h.K > 0 ? (int)(s->steps_done % s->pes_h.K) : 0;
    SSB_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; ++i) one_step(s, i, i > 0, i + 1 < n);
    advance(s, n);
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    for (int k = 0; k < K_NKINDS; ++k) s->kind_per_graph[k] = s->kind_launches[k] - saved[k];
    memcpy(s->kind_launches, saved, sizeof(saved));
    s->total_launches = saved_total;
    s->profiling = prof;
    if (e != cudaSuccess) return fail(-2, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(-2, std::string("graph instantiate: ") + cudaGetErrorString(e));
    s->graph_steps = n;
    return 0;
}

int push_dyn(ssb_sim* s) {
    long long h[4] = {s->steps_done, s->tab_step0, s->probe_step0, s->syn_step0};
    SSB_CUDA(cudaMemcpyAsync(s->dyn, h, sizeof(h), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

}  // namespace

extern "C" {

const char* ssb_last_error(void) { return g_err.c_str(); }
const char* ssb_version(void) { return "sspslam_b200 0.1 (sm_100a)"; }

int ssb_create(int device, int n_trials, ssb_sim** out) {
    if (!out || n_trials <= 0) return fail(-1, "ssb_create: bad arguments");
    int count = 0;
    SSB_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(-1, "ssb_create: no such CUDA device");
    SSB_CUDA(cudaSetDevice(device));
    ssb_sim* s = new ssb_sim();
    s->device = device;
    s->n_trials = n_trials;
    s->B = (n_trials + 31) / 32 * 32;
    s->n_groups = s->B / 32;
    if (const char* e = getenv("SSB_DEBUG_SYNC")) s->debug_sync = e[0] == '1';
    if (s->debug_sync) s->use_graph = false;
    SSB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    {   // the long HBM-streaming chain (B) and the decode chain (D) are scheduled ahead of the rest
        int pr_lo = 0, pr_hi = 0;
        SSB_CUDA(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
        // the clean-up chain C (short, and its scan needs whole SMs) and the chain of the later levels E go first, then the
        // long HBM-streaming chain B and the decode chain D; the narrow-ensemble kernels on A fill what is left
        const int pr_mid = std::min(pr_lo, pr_hi + 1);
        // SSB_PRIOS = "b,c,d,e": distance of each chain's priority from the highest one (tuning knob; default 1,0,1,0)
        int off[4] = {1, 0, 1, 0};
        if (const char* e = getenv("SSB_PRIOS")) sscanf(e, "%d,%d,%d,%d", &off[0], &off[1], &off[2], &off[3]);
        if (const char* pd = getenv("SSB_PRIO_D")) off[2] = pd[0] == '1' ? 0 : 1;
        for (int i = 0; i < 4; ++i)
            SSB_CUDA(cudaStreamCreateWithPriority(&s->aux[i], cudaStreamNonBlocking, std::min(pr_lo, pr_hi + std::max(0, off[i]))));
        SSB_CUDA(cudaStreamCreateWithPriority(&s->aux[4], cudaStreamNonBlocking, pr_mid));
        SSB_CUDA(cudaStreamCreateWithPriority(&s->aux[5], cudaStreamNonBlocking, pr_mid));
    }
    if (const char* e = getenv("SSB_SERIAL")) s->parallel = e[0] != '1';
    if (const char* e = getenv("SSB_LIN_MINB")) s->lin_six = e[0] == '6';
    if (const char* e = getenv("SSB_PES_PAD")) s->pes_pad_smem = (size_t)atoi(e) * 1024;
    if (const char* e = getenv("SSB_VOJA_PAD")) s->voja_pad_smem = (size_t)atoi(e) * 1024;
    SSB_CUDA(cudaEventCreate(&s->ev_run0));
    SSB_CUDA(cudaEventCreate(&s->ev_run1));
    *out = s;
    return 0;
}

int ssb_set_array(ssb_sim* s, const char* name, const void* data, size_t bytes) {
    if (!s || !name || (!data && bytes)) return fail(-1, "ssb_set_array: bad arguments");
    if (s->finalized) return fail(-1, "ssb_set_array: plan already finalized");
    HostArray& a = s->arrays[name];
    a.bytes.assign((const unsigned char*)data, (const unsigned char*)data + bytes);
    return 0;
}

int ssb_set_scalar(ssb_sim* s, const char* name, double value) {
    if (!s || !name) return fail(-1, "ssb_set_scalar: bad arguments");
    s->scalars[name] = value;
    return 0;
}

int ssb_finalize(ssb_sim* s) {
    if (!s) return fail(-1, "ssb_finalize: null handle");
    if (s->finalized) return fail(-1, "ssb_finalize: already finalized");
    SSB_CUDA(cudaSetDevice(s->device));
    s->nv = iscalar(s, "nv");
    s->nf = iscalar(s, "nf");
    s->nt = iscalar(s, "nt");
    s->tab_row0 = iscalar(s, "tab_row0");
    s->nn = iscalar(s, "nn");
    s->n_act = iscalar(s, "n_act");
    s->n_lenc = iscalar(s, "n_lenc");
    s->n_ldec = iscalar(s, "n_ldec");
    s->n_afilt = iscalar(s, "n_afilt");
    s->n_probe = iscalar(s, "n_probe");
    s->n_part = iscalar(s, "n_part");
    s->n_counters = iscalar(s, "n_jtiles") * s->n_groups;
    s->n_levels = (int)iscalar(s, "n_levels");
    s->pes_level = s->scalars.count("pes_level") ? (int)s->scalars["pes_level"] : -1;
    s->chunk_cap = (int)iscalar(s, "chunk_cap");
    if (s->nv < 1 || s->chunk_cap < 1 || s->n_levels < 1) return fail(-1, "ssb_finalize: plan scalars missing");
    const double dt = s->scalars.count("dt") ? s->scalars["dt"] : 0.001;

    size_t cnt = 0;
    if (upload_array(s, "csr_ptr", &s->d_csr_ptr)) return -2;
    if (upload_array(s, "csr_ent0", &s->d_ent0)) return -2;
    if (upload_array(s, "csr_ent1", &s->d_ent1)) return -2;
    if (upload_array(s, "weights", &s->d_W)) return -2;
    if (upload_array(s, "ens_small", &s->d_small, &cnt)) return -2;
    s->n_small_total = (int)(cnt / 9);
    if (upload_array(s, "ens_big", &s->d_big)) return -2;
    if (upload_array(s, "dec", &s->d_dec)) return -2;
    if (upload_array(s, "pes", &s->d_pes, &cnt)) return -2;
    s->n_pes = (int)(cnt / 13);
    if (upload_array(s, "cleanup", &s->d_cleanup)) return -2;
    if (upload_array(s, "gate", &s->d_gate)) return -2;
    s->lin0 = (int)iscalar(s, "lin0");
    s->n_lin = (int)iscalar(s, "n_lin");
    s->n_lin_early = (int)iscalar(s, "n_lin_early");
    s->n_lin_fused = (int)iscalar(s, "n_lin_fused");
    s->n_lvl0_res = (int)iscalar(s, "n_lvl0_res");
    if (upload_array(s, "ntypes", &s->d_ntypes)) return -2;
    if (upload_array(s, "cleanup_s64", &s->d_s64)) return -2;
    s->h_stages = host_ints(s, "stages");
    s->h_small = host_ints(s, "ens_small");
    s->h_big = host_ints(s, "ens_big");
    s->h_dec = host_ints(s, "dec");
    s->h_cleanup = host_ints(s, "cleanup");
    s->h_pes = host_ints(s, "pes");
    s->level_deps = host_ints(s, "level_deps");
    if (s->level_deps.size() != (size_t)s->n_levels * s->n_levels) s->level_deps.clear();
    if (const char* e = getenv("SSB_LEVEL_DEPS")) if (e[0] == '0') s->level_deps.clear();   // A/B switch: level barriers
    if ((int)s->h_stages.size() != s->n_levels * 12) return fail(-1, "ssb_finalize: stages array has wrong size");
    s->per_trial = iscalar(s, "per_trial_weights") != 0;
    if (int rc = setup_pes_defer(s)) return rc;        // decides whether k_pes_hist owns the PES activity traces
    if (int rc = build_lin_program(s)) return rc;
    if (int rc = build_decode_tiles(s)) return rc;
    if (int rc = build_encode_tiles(s)) return rc;
    if (int rc = build_encode_tiles_k(s)) return rc;
    s->levels.assign(s->n_levels, LevelInfo());
    for (int lvl = 0; lvl < s->n_levels; ++lvl) {
        const int* st = &s->h_stages[lvl * 12];
        LevelInfo& li = s->levels[lvl];
        auto act_is_voja = [&](int act0) {   // does this activity range belong to a Voja ensemble of the level?
            for (int i = 0; i < st[3]; ++i) {
                const int* d = &s->h_big[(st[2] + i) * 16];
                if (act0 >= d[4] && act0 < d[4] + d[0]) return (d[9] & 1) != 0;
            }
            return true;                     // produced at an earlier level: already joined, but stay safe
        };
        for (int i = 0; i < st[3]; ++i) {
            if (s->h_big[(st[2] + i) * 16 + 9] & 1) li.n_voja++;
            else li.n_static++;
        }
        for (int i = 0; i < st[5]; ++i)
            if (act_is_voja(s->h_dec[(st[4] + i) * 9 + 3])) li.dec_needs_voja = true;
        if (lvl == s->pes_level)
            for (int i = 0; i < s->n_pes; ++i)
                if (!act_is_voja(s->h_pes[i * 13 + 4])) s->pes_needs_static = true;
    }
    {   // fused deferred-PES decode: every PES descriptor must be fed by a Voja ensemble of the PES level, and one decoder
        // pass (56 rows + K history rows per warp) must fit in the warp's encoder ring (nb * dims rows; nb >= 1)
        const int n_big = (int)(s->h_big.size() / 16);
        s->pes_of_big.assign(n_big + 1, -1);
        const char* e = getenv("SSB_PES_FUSE");
        // opt-in (SSB_PES_FUSE=1): measured on B200 the fused post-pass lengthens the bandwidth-bound Voja kernel by more
        // than the separate sparse-decode kernel costs (too little memory-level parallelism on 2 CTAs per SM)
        bool ok = s->pes_h.K > 0 && s->n_pes > 0 && s->pes_level >= 0 && (e && e[0] == '1');
        for (int i = 0; ok && i < s->n_pes; ++i) {
            const int* pd = &s->h_pes[i * 13];
            const int* st = &s->h_stages[s->pes_level * 12];
            int found = -1;
            for (int b = st[2]; b < st[2] + st[3]; ++b) {
                const int* d = &s->h_big[b * 16];
                if ((d[9] & 1) && d[4] == pd[4] && d[0] == pd[0]) found = b;
            }
            if (found < 0) {
                ok = false;
                break;
            }
            const int* d = &s->h_big[found * 16];
            int nb = 2;                                   // as launch_wide_class sizes the ring of a warp
            while (nb > 1 && (size_t)(d[2] * 32 + d[11] * 32 + nb * d[1] * 32) * sizeof(float) > 200 * 1024) --nb;
            if (std::min(pd[1], 56) + 8 > nb * d[1] || (size_t)d[1] * 128 * 2 > 96 * 1024) ok = false;   // (stream-class ensembles: no fusion)
            else s->pes_of_big[found] = i;
        }
        s->pes_fused = ok;
        if (!ok) s->pes_of_big.assign(n_big + 1, -1);
    }
    for (size_t i = 0; i + 8 < s->h_small.size(); i += 9)
        if (s->h_small[i + 8] > SSB_SM_WMAX || (s->h_small[i + 8] & 3))
            return fail(-1, "ssb_finalize: narrow-ensemble weight stride out of range");

    const int B = s->B;
    if (s->per_trial) {
        // every trial has its own network seed: the seed-dependent static weights (narrow ensembles' packed rows, wide
        // ensembles' bias / Voja scale) become one more per-trial arena; wide encoders live in lenc, wide decoders in ldec
        // (the lowering flags those ensembles and points the decoder descriptors at ldec rows).
        if (!s->arrays.count("weights_pt")) return fail(-1, "ssb_finalize: per_trial_weights needs the weights_pt array");
        for (size_t i = 0; i + 15 < s->h_big.size(); i += 16)
            if (!(s->h_big[i + 9] & 1) || !(s->h_big[i + 9] & 4))
                return fail(-1, "ssb_finalize: per-trial plan with a shared-weight wide ensemble");
        s->n_wpt = (long long)(s->arrays["weights_pt"].bytes.size() / sizeof(float));
        if (alloc_rows(&s->wpt, s->n_wpt, B)) return -2;
    }
    if (alloc_rows(&s->vec, s->nv, B)) return -2;
    if (alloc_rows(&s->st, s->nn, B)) return -2;
    if (alloc_rows(&s->act, s->n_act, B)) return -2;
    if (alloc_rows(&s->lenc, s->n_lenc, B)) return -2;
    if (alloc_rows(&s->ldec, s->n_ldec, B)) return -2;
    if (alloc_rows(&s->afilt, 2 * s->n_afilt, B)) return -2;
    if (alloc_rows(&s->part, s->n_part, B)) return -2;
    if (alloc_rows(&s->tab, (long long)s->chunk_cap * std::max(1LL, s->nt), B)) return -2;
    if (alloc_rows(&s->probe, (long long)s->chunk_cap * std::max(1LL, s->n_probe), B)) return -2;
    SSB_CUDA(cudaMalloc((void**)&s->counters, (size_t)std::max(1LL, s->n_counters) * sizeof(int)));
    SSB_CUDA(cudaMemset(s->counters, 0, (size_t)std::max(1LL, s->n_counters) * sizeof(int)));
    SSB_CUDA(cudaMalloc((void**)&s->dyn, 4 * sizeof(long long)));
    SSB_CUDA(cudaMemset(s->dyn, 0, 4 * sizeof(long long)));
    {   // vec row 0 = ones, for every trial group
        std::vector<float> ones(32, 1.0f);
        for (int g = 0; g < s->n_groups; ++g)
            SSB_CUDA(cudaMemcpy(s->vec + (size_t)g * s->nv * 32, ones.data(), 32 * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int n_cleanup = (int)(s->h_cleanup.size() / 6);
    s->cleanups.resize(n_cleanup);
    if (n_cleanup) {
        SSB_CUDA(cudaMalloc((void**)&s->cidx, (size_t)n_cleanup * B * sizeof(int)));
        SSB_CUDA(cudaMemset(s->cidx, 0, (size_t)n_cleanup * B * sizeof(int)));
    }
    size_t s64_off = 0;
    const size_t s64_total = s->arrays.count("cleanup_s64") ? s->arrays["cleanup_s64"].bytes.size() / sizeof(double) : 0;
    for (int i = 0; i < n_cleanup; ++i) {
        const int* d = &s->h_cleanup[i * 6];
        CleanupDev& cd = s->cleanups[i];
        scan_geometry(d[0], d[2], s->n_groups, &cd);
        if (alloc_rows(&cd.cx, d[2], B)) return -2;
        if (alloc_rows(&cd.pval, scan_n_cand(cd), B)) return -2;
        if (alloc_rows(reinterpret_cast<float**>(&cd.pidx), scan_n_cand(cd), B)) return -2;
        cd.idx = s->cidx + (size_t)i * B;
        const size_t need = (size_t)d[0] * d[1];
        if (s64_off + need <= s64_total) {
            cd.s64 = s->d_s64 + s64_off;
            s64_off += need;
        }
        if (cd.tc || cd.tck) {
            const float* hW = reinterpret_cast<const float*>(s->arrays["weights"].bytes.data());
            if (cd.tc ? build_scan_tiles(hW + d[3], d[0], d[2], &cd) : build_scan_tiles_k(hW + d[3], d[0], d[2], s->n_groups, &cd))
                return -2;
        }
    }
    // opt-in to large dynamic shared memory for very wide ensembles (d = 649)
    wide_smem_optin();
    scan_smem_optin();

    for (int lvl = 0; lvl < s->n_levels; ++lvl) {   // decide the wide-ensemble launch geometry now (occupancy queries)
        launch_wide(s, nullptr, &s->h_stages[lvl * 12], false, 0, true);
        launch_wide(s, nullptr, &s->h_stages[lvl * 12], true, 0, true);
    }

    SsbCtx& c = s->ctx;
    c.G = s->n_groups;
    c.nv = (int)s->nv;
    c.nf = (int)s->nf;
    c.nt = (int)s->nt;
    c.tab_row0 = (int)s->tab_row0;
    c.nn = (int)std::max(1LL, s->nn);
    c.n_act = (int)std::max(1LL, s->n_act);
    c.n_lenc = (int)std::max(1LL, s->n_lenc);
    c.n_ldec = (int)std::max(1LL, s->n_ldec);
    c.n_afilt = (int)s->n_afilt;
    c.n_probe = (int)s->n_probe;
    c.n_part = (int)std::max(1LL, s->n_part);
    c.tab_cap = s->chunk_cap;
    c.probe_cap = s->chunk_cap;
    c.dt = (float)dt;
    c.vec = s->vec;
    c.tab = s->tab;
    c.st = s->st;
    c.act = s->act;
    SSB_CUDA(cudaMalloc((void**)&s->aflag, (size_t)std::max(1LL, s->n_act) * s->n_groups * sizeof(int)));
    SSB_CUDA(cudaMemset(s->aflag, 0, (size_t)std::max(1LL, s->n_act) * s->n_groups * sizeof(int)));
    c.aflag = s->aflag;
    c.lenc = s->lenc;
    c.ldec = s->ldec;
    c.afilt = s->afilt;
    c.probe = s->probe;
    c.part = s->part;
    c.counters = s->counters;
    c.W = s->d_W;
    c.wpt = s->wpt;
    c.n_wpt = (int)std::max(1LL, s->n_wpt);
    c.csr_ptr = s->d_csr_ptr;
    c.ent0 = s->d_ent0;
    c.ent1 = s->d_ent1;
    c.ntypes = s->d_ntypes;
    c.dyn = s->dyn;
    s->finalized = true;
    SSB_CUDA(cudaDeviceSynchronize());
    return 0;
}

int ssb_upload(ssb_sim* s, const char* name, size_t row0, size_t n_rows, const float* host) {
    if (!s || !s->finalized || !host) return fail(-1, "ssb_upload: bad arguments");
    ArenaRef a;
    if (arena(s, name, &a)) return -3;
    if ((long long)(row0 + n_rows) > a.rows) return fail(-1, "ssb_upload: rows out of range");
    SSB_CUDA(cudaSetDevice(s->device));
    if (a.ptr == s->ldec && pes_flush(s)) return -2;
    if (a.ptr == s->ldec) {
        // learned decoders are not row-per-trial-vector: a group's block is [neuron][trial][JP] (ssb_pes.cuh); the host
        // passes the device order, [group][n_rows][32] floats, for the row range of one decoder
        for (int g = 0; g < s->n_groups; ++g)
            SSB_CUDA(cudaMemcpyAsync(a.ptr + ((size_t)g * a.rows + row0) * 32, host + (size_t)g * n_rows * 32,
                                     n_rows * 32 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    } else if (a.tiled) {
        if (copy_rows(s, a.ptr, std::max(1LL, a.rows), row0, n_rows, const_cast<float*>(host), true)) return -2;
    } else {
        SSB_CUDA(cudaMemcpyAsync(a.ptr + row0 * s->B, host, n_rows * s->B * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    }
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int ssb_download(ssb_sim* s, const char* name, size_t row0, size_t n_rows, float* host) {
    if (!s || !s->finalized || !host) return fail(-1, "ssb_download: bad arguments");
    ArenaRef a;
    if (arena(s, name, &a)) return -3;
    if ((long long)(row0 + n_rows) > a.rows) return fail(-1, "ssb_download: rows out of range");
    SSB_CUDA(cudaSetDevice(s->device));
    if (a.ptr == s->ldec && pes_flush(s)) return -2;
    if (a.ptr == s->ldec) {
        for (int g = 0; g < s->n_groups; ++g)
            SSB_CUDA(cudaMemcpyAsync(host + (size_t)g * n_rows * 32, a.ptr + ((size_t)g * a.rows + row0) * 32,
                                     n_rows * 32 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    } else if (a.tiled) {
        if (copy_rows(s, a.ptr, std::max(1LL, a.rows), row0, n_rows, host, false)) return -2;
    } else {
        SSB_CUDA(cudaMemcpyAsync(host, a.ptr + row0 * s->B, n_rows * s->B * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int ssb_set_tables(ssb_sim* s, const float* host, long long step0, int n_steps) {
    if (!s || !s->finalized) return fail(-1, "ssb_set_tables: bad handle");
    if (n_steps < 0 || n_steps > s->chunk_cap) return fail(-1, "ssb_set_tables: n_steps exceeds chunk_cap");
    SSB_CUDA(cudaSetDevice(s->device));
    if (s->nt > 0 && n_steps > 0) {
        if (!host) return fail(-1, "ssb_set_tables: null host buffer");
        // host [n_steps][nt][B] -> device [G][chunk_cap][nt][32]: one strided copy per trial group
        if (copy_rows(s, s->tab, (long long)s->chunk_cap * s->nt, 0, (size_t)n_steps * s->nt, const_cast<float*>(host), true))
            return -2;
    }
    s->tab_step0 = step0;
    s->tab_steps = n_steps;
    return 0;
}

int ssb_rebase_tables(ssb_sim* s, long long step0) {
    if (!s || !s->finalized) return fail(-1, "ssb_rebase_tables: bad handle");
    s->tab_step0 = step0;
    return 0;
}

int ssb_run_steps(ssb_sim* s, int n_steps) {
    if (!s || !s->finalized) return fail(-1, "ssb_run_steps: bad handle");
    if (n_steps <= 0) return 0;
    if (n_steps > s->chunk_cap) return fail(-1, "ssb_run_steps: n_steps exceeds chunk_cap (probe buffer)");
    if (s->synth_on) {
        if (s->steps_done < s->syn_step0 || s->steps_done + n_steps > s->syn_step0 + s->syn_steps)
            return fail(-4, "ssb_run_steps: resident step indices (ssb_synth_steps) do not cover the requested steps");
    } else if (s->nt > 0 && (s->steps_done < s->tab_step0 || s->steps_done + n_steps > s->tab_step0 + s->tab_steps))
        return fail(-4, "ssb_run_steps: resident input tables do not cover the requested steps");
    SSB_CUDA(cudaSetDevice(s->device));
    s->probe_step0 = s->steps_done;
    if (push_dyn(s)) return -2;
    SSB_CUDA(cudaEventRecord(s->ev_run0, s->stream));
    int i = 0;
    if (s->use_graph && !s->profiling) {
        const int gs = 16;
        if (n_steps >= gs && !s->step_graph) {
            if (build_graph(s, gs)) return -2;
        }
        const bool phase_ok = s->pes_h.K == 0 || (int)(s->steps_done % s->pes_h.K) == s->graph_phase;
        if (s->step_graph && phase_ok) {
            int replays = 0;
            for (; i + s->graph_steps <= n_steps; i += s->graph_steps, ++replays)
                SSB_CUDA(cudaGraphLaunch(s->step_graph, s->stream));
            for (int k = 0; k < K_NKINDS; ++k) {   // account the replayed kernel launches
                s->kind_launches[k] += (long long)replays * s->kind_per_graph[k];
                s->total_launches += (long long)replays * s->kind_per_graph[k];
            }
        }
    }
    if (i < n_steps) {
        const int rest = n_steps - i;
        s->dep_used = 0;
        s->step_base = s->steps_done + i;
        for (int r = 0; r < rest; ++r) {
            if (one_step(s, r, r > 0, r + 1 < rest)) return -2;
        }
        advance(s, rest);
    }
    SSB_CUDA(cudaEventRecord(s->ev_run1, s->stream));
    s->run_timed = true;
    s->steps_done += n_steps;
    SSB_CUDA(cudaGetLastError());
    return 0;
}

// ssb_set_tables + ssb_run_steps + ssb_read_probes as one software pipeline over 16-step sub-chunks: the input tables of
// sub-chunk j+1 are copied (H2D stream) while sub-chunk j computes, and the probe rows of sub-chunk j go back to the
// host (D2H stream) while sub-chunk j+1 computes.  Asynchronous: ssb_io_wait() returns once steps and copies are done.
int ssb_run_steps_io(ssb_sim* s, const float* host_tables, int n_steps, float* host_probes) {
    if (!s || !s->finalized) return fail(-1, "ssb_run_steps_io: bad handle");
    if (n_steps <= 0) return 0;
    if (n_steps > s->chunk_cap) return fail(-1, "ssb_run_steps_io: n_steps exceeds chunk_cap");
    const bool copy_tables = s->nt > 0 && !s->synth_on;
    if (copy_tables && !host_tables) return fail(-1, "ssb_run_steps_io: null table buffer");
    if (s->synth_on && (s->steps_done < s->syn_step0 || s->steps_done + n_steps > s->syn_step0 + s->syn_steps))
        return fail(-4, "ssb_run_steps_io: resident step indices (ssb_synth_steps) do not cover the requested steps");
    SSB_CUDA(cudaSetDevice(s->device));
    if (!s->io_h2d) {
        SSB_CUDA(cudaStreamCreateWithFlags(&s->io_h2d, cudaStreamNonBlocking));
        SSB_CUDA(cudaStreamCreateWithFlags(&s->io_d2h, cudaStreamNonBlocking));
    }
    const int sub = 16;
    const int n_sub = (n_steps + sub - 1) / sub;
    while ((int)s->io_events.size() < 2 * n_sub + 1) {
        cudaEvent_t e;
        SSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->io_events.push_back(e);
    }
    const bool graphs = s->use_graph && !s->profiling;
    if (graphs && n_steps >= sub && !s->step_graph) {
        if (build_graph(s, sub)) return -2;
    }
    s->tab_step0 = s->steps_done;
    s->tab_steps = n_steps;
    s->probe_step0 = s->steps_done;
    if (push_dyn(s)) return -2;
    SSB_CUDA(cudaEventRecord(s->ev_run0, s->stream));
    // the copy streams start after everything already queued on the compute stream (previous readers of the arenas)
    SSB_CUDA(cudaEventRecord(s->io_events[2 * n_sub], s->stream));
    SSB_CUDA(cudaStreamWaitEvent(s->io_h2d, s->io_events[2 * n_sub], 0));
    for (int j = 0; j < n_sub; ++j) {
        const int j0 = j * sub, jn = std::min(sub, n_steps - j0);
        if (copy_tables) {
            if (copy_rows(s, s->tab, (long long)s->chunk_cap * s->nt, (size_t)j0 * s->nt, (size_t)jn * s->nt,
                          const_cast<float*>(host_tables) + (size_t)j0 * s->nt * s->B, true, s->io_h2d))
                return -2;
            SSB_CUDA(cudaEventRecord(s->io_events[2 * j], s->io_h2d));
        }
    }
    for (int j = 0; j < n_sub; ++j) {
        const int j0 = j * sub, jn = std::min(sub, n_steps - j0);
        if (copy_tables) SSB_CUDA(cudaStreamWaitEvent(s->stream, s->io_events[2 * j], 0));
        const bool phase_ok = s->pes_h.K == 0 || (int)((s->steps_done + j0) % s->pes_h.K) == s->graph_phase;
        if (graphs && s->step_graph && jn == s->graph_steps && phase_ok) {
            SSB_CUDA(cudaGraphLaunch(s->step_graph, s->stream));
            for (int k = 0; k < K_NKINDS; ++k) {
                s->kind_launches[k] += s->kind_per_graph[k];
                s->total_launches += s->kind_per_graph[k];
            }
        } else {
            s->dep_used = 0;
            s->step_base = s->steps_done + j0;
            for (int r = 0; r < jn; ++r)
                if (one_step(s, r, r > 0, r + 1 < jn)) return -2;
            advance(s, jn);
        }
        if (host_probes && s->n_probe > 0) {
            SSB_CUDA(cudaEventRecord(s->io_events[2 * j + 1], s->stream));
            SSB_CUDA(cudaStreamWaitEvent(s->io_d2h, s->io_events[2 * j + 1], 0));
            if (copy_rows(s, s->probe, (long long)s->chunk_cap * s->n_probe, (size_t)j0 * s->n_probe, (size_t)jn * s->n_probe,
                          host_probes + (size_t)j0 * s->n_probe * s->B, false, s->io_d2h))
                return -2;
        }
    }
    SSB_CUDA(cudaEventRecord(s->ev_run1, s->stream));
    s->run_timed = true;
    s->steps_done += n_steps;
    SSB_CUDA(cudaGetLastError());
    return 0;
}

int ssb_io_wait(ssb_sim* s) {
    if (!s || !s->finalized) return fail(-1, "ssb_io_wait: bad handle");
    SSB_CUDA(cudaSetDevice(s->device));
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->io_d2h) SSB_CUDA(cudaStreamSynchronize(s->io_d2h));
    SSB_CUDA(cudaGetLastError());
    return 0;
}

int ssb_synth_setup(ssb_sim* s, const int* cfg, const float* fparams, const double* phases, const double* lm_sp,
                    const float* path_rows, const float* vel_rows, const float* lm_rows) {
    if (!s || !s->finalized || !cfg || !fparams || !phases || !path_rows || !vel_rows)
        return fail(-1, "ssb_synth_setup: bad arguments");
    if (s->synth_on) return fail(-1, "ssb_synth_setup: already configured");
    SSB_CUDA(cudaSetDevice(s->device));
    SsbSynth& y = s->synth;
    memset(&y, 0, sizeof(y));
    y.dim = cfg[0];
    y.d = cfg[1];
    y.n_lm = cfg[2];
    y.T = cfg[3];
    y.vel_col = cfg[4];
    y.init_col = cfg[5];
    y.lmvec_col = cfg[6];
    y.lmsp_col = cfg[7];
    y.nolm_col = cfg[8];
    y.view_rad = fparams[0];
    y.none_value = fparams[1];
    if (y.dim < 1 || y.dim > 3 || y.d < 1 || y.T < 2 || y.n_lm < 0) return fail(-1, "ssb_synth_setup: bad sizes");
    if (y.n_lm > 0 && (!lm_sp || !lm_rows)) return fail(-1, "ssb_synth_setup: landmark arrays missing");
    const int cols[5] = {y.vel_col, y.init_col, y.lmvec_col, y.lmsp_col, y.nolm_col};
    const int widths[5] = {y.dim, y.d, y.d, y.d, 1};
    for (int i = 0; i < 5; ++i)
        if (cols[i] >= 0 && cols[i] + widths[i] > s->nt) return fail(-1, "ssb_synth_setup: input column outside the table rows");
    if ((size_t)2 * y.d * 32 * sizeof(float) > 200 * 1024) return fail(-1, "ssb_synth_setup: ssp_dim too large");
    const size_t B = s->B;
    auto rows_up = [&](float** dst, const float* host, long long rows) {
        if (alloc_rows(dst, rows, (int)B)) return 1;
        if (rows > 0 && copy_rows(s, *dst, rows, 0, (size_t)rows, const_cast<float*>(host), true)) return 1;
        return 0;
    };
    if (rows_up(&s->syn_path, path_rows, (long long)y.T * y.dim) || rows_up(&s->syn_vel, vel_rows, (long long)y.T * y.dim)) return -2;
    if (y.n_lm > 0 && rows_up(&s->syn_lm, lm_rows, (long long)y.n_lm * y.dim)) return -2;
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<float> ph((size_t)y.d * y.dim), sp((size_t)std::max(1, y.n_lm) * y.d, 0.f), ct((size_t)y.d * y.d), st((size_t)y.d * y.d);
    for (size_t i = 0; i < ph.size(); ++i) ph[i] = (float)phases[i];
    for (size_t i = 0; i < (size_t)y.n_lm * y.d; ++i) sp[i] = (float)lm_sp[i];
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < y.d; ++k)
        for (int m = 0; m < y.d; ++m) {
            const double ang = two_pi * (double)(((long long)k * m) % y.d) / (double)y.d;
            ct[(size_t)k * y.d + m] = (float)cos(ang);
            st[(size_t)k * y.d + m] = (float)sin(ang);
        }
    auto up = [&](float** dst, const std::vector<float>& v) {
        if (cudaMalloc((void**)dst, v.size() * sizeof(float)) != cudaSuccess) return 1;
        return cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ? 1 : 0;
    };
    if (up(&s->syn_phases, ph) || up(&s->syn_lmsp, sp) || up(&s->syn_cos, ct) || up(&s->syn_sin, st))
        return fail(-2, "ssb_synth_setup: upload failed");
    SSB_CUDA(cudaMalloc((void**)&s->syn_idx, (size_t)s->chunk_cap * 4 * sizeof(int)));
    SSB_CUDA(cudaMemset(s->syn_idx, 0, (size_t)s->chunk_cap * 4 * sizeof(int)));
    SSB_CUDA(cudaFuncSetAttribute(k_synth<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SSB_CUDA(cudaFuncSetAttribute(k_synth<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    y.path = s->syn_path;
    y.vel = s->syn_vel;
    y.lm = s->syn_lm;
    y.phases = s->syn_phases;
    y.lm_sp = s->syn_lmsp;
    y.cosT = s->syn_cos;
    y.sinT = s->syn_sin;
    y.idx = s->syn_idx;
    s->synth_on = true;
    if (s->step_graph) {                     // a graph captured with k_begin is stale now
        cudaGraphExecDestroy(s->step_graph);
        s->step_graph = nullptr;
    }
    return 0;
}

int ssb_synth_steps(ssb_sim* s, const int* idx, long long step0, int n_steps) {
    if (!s || !s->finalized || !s->synth_on || !idx) return fail(-1, "ssb_synth_steps: bad arguments");
    if (n_steps < 0 || n_steps > s->chunk_cap) return fail(-1, "ssb_synth_steps: n_steps exceeds chunk_cap");
    SSB_CUDA(cudaSetDevice(s->device));
    std::vector<int> h((size_t)n_steps * 4, 0);
    for (int i = 0; i < n_steps; ++i) {
        const int ip = idx[i * 3], ic = idx[i * 3 + 1];
        if (ip < 0 || ip >= s->synth.T || ic < 0 || ic >= s->synth.T) return fail(-1, "ssb_synth_steps: path index out of range");
        h[(size_t)i * 4] = ip;
        h[(size_t)i * 4 + 1] = ic;
        h[(size_t)i * 4 + 2] = idx[i * 3 + 2];
    }
    // the stream is idle between calls of the synchronous run entry points; a plain ordered copy is enough
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    SSB_CUDA(cudaMemcpy(s->syn_idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    s->syn_step0 = step0;
    s->syn_steps = n_steps;
    return 0;
}

int ssb_read_probes(ssb_sim* s, float* host, long long step0, int n_steps) {
    if (!s || !s->finalized || !host) return fail(-1, "ssb_read_probes: bad arguments");
    if (step0 < s->probe_step0 || step0 + n_steps > s->steps_done || n_steps < 0)
        return fail(-4, "ssb_read_probes: steps not in the probe buffer");
    if (s->n_probe == 0 || n_steps == 0) return 0;
    SSB_CUDA(cudaSetDevice(s->device));
    // device [G][chunk_cap][n_probe][32] -> host [n_steps][n_probe][B]
    if (copy_rows(s, s->probe, (long long)s->chunk_cap * s->n_probe, (size_t)(step0 - s->probe_step0) * s->n_probe,
                  (size_t)n_steps * s->n_probe, host, false))
        return -2;
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

long long ssb_n_steps(ssb_sim* s) { return s ? s->steps_done : -1; }
int ssb_n_trials_padded(ssb_sim* s) { return s ? s->B : -1; }

int ssb_sync(ssb_sim* s) {
    if (!s) return fail(-1, "ssb_sync: null handle");
    SSB_CUDA(cudaSetDevice(s->device));
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    SSB_CUDA(cudaGetLastError());
    return 0;
}

int ssb_reset(ssb_sim* s) {
    if (!s || !s->finalized) return fail(-1, "ssb_reset: bad handle");
    SSB_CUDA(cudaSetDevice(s->device));
    const size_t B = s->B;
    auto zero = [&](float* p, long long rows) {
        return cudaMemsetAsync(p, 0, (size_t)std::max(1LL, rows) * B * sizeof(float), s->stream);
    };
    SSB_CUDA(zero(s->vec, s->nv));
    SSB_CUDA(zero(s->st, s->nn));
    SSB_CUDA(zero(s->act, s->n_act));
    SSB_CUDA(zero(s->lenc, s->n_lenc));
    SSB_CUDA(zero(s->ldec, s->n_ldec));
    SSB_CUDA(zero(s->afilt, 2 * s->n_afilt));
    if (s->pes_h.K > 0) {
        SSB_CUDA(zero(s->pes_h.hist_e, s->pes_h.rows_e));
        SSB_CUDA(zero(s->pes_h.hist_f, s->pes_h.rows_f));
        SSB_CUDA(cudaMemsetAsync(s->pes_h.counters, 0, (size_t)s->n_pes * s->n_groups * sizeof(int), s->stream));
    }
    SSB_CUDA(cudaMemsetAsync(s->counters, 0, (size_t)std::max(1LL, s->n_counters) * sizeof(int), s->stream));
    {
        std::vector<float> ones(32, 1.0f);
        for (int g = 0; g < s->n_groups; ++g)
            SSB_CUDA(cudaMemcpyAsync(s->vec + (size_t)g * s->nv * 32, ones.data(), 32 * sizeof(float), cudaMemcpyHostToDevice,
                                     s->stream));
        SSB_CUDA(cudaStreamSynchronize(s->stream));
    }
    s->steps_done = 0;
    s->probe_step0 = 0;
    if (push_dyn(s)) return -2;
    SSB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

void ssb_destroy(ssb_sim* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    void* ptrs[] = {s->d_csr_ptr, s->d_ent0, s->d_ent1, s->d_W, s->d_small, s->d_big, s->d_dec, s->d_pes, s->d_cleanup,
                    s->d_gate, s->d_lin_rows, s->d_lin_ab, s->d_dense_items, s->d_dense_desc, s->d_dense_cols, s->d_dense_rows, s->d_dense_T, s->d_lin_recs, s->d_dec_wt, s->d_dec_wt_off, s->d_enc_t, s->d_enc_t_off, s->pes_h.hist_e, s->pes_h.hist_f, s->pes_h.part,
                    s->pes_h.counters, s->d_pes_hdesc, s->aflag, s->syn_path, s->syn_vel, s->syn_lm, s->syn_phases, s->syn_lmsp,
                    s->syn_cos, s->syn_sin, s->syn_idx, s->d_etk, s->d_xtk, s->d_lin_tc, s->d_lin_ttk, s->d_lin_xt, s->d_ntypes, s->d_s64, s->vec, s->tab, s->st, s->act,
                    s->lenc, s->ldec, s->afilt, s->probe, s->part, s->counters, s->dyn, s->cidx, s->wpt};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& cd : s->cleanups) {
        if (cd.cx) cudaFree(cd.cx);
        if (cd.pval) cudaFree(cd.pval);
        if (cd.pidx) cudaFree(cd.pidx);
        if (cd.stc) cudaFree(cd.stc);
        if (cd.stck) cudaFree(cd.stck);
        if (cd.xt) cudaFree(cd.xt);
    }
    for (auto& L : s->tck_levels)
        if (L.d_desc6) cudaFree(L.d_desc6);
    if (s->step_graph) cudaGraphExecDestroy(s->step_graph);
    for (auto e : s->dep_pool) cudaEventDestroy(e);
    for (auto e : s->io_events) cudaEventDestroy(e);
    if (s->io_h2d) cudaStreamDestroy(s->io_h2d);
    if (s->io_d2h) cudaStreamDestroy(s->io_d2h);
    for (auto a : s->aux)
        if (a) cudaStreamDestroy(a);
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    for (auto e : s->ev_mark)
        if (e) cudaEventDestroy(e);
    if (s->ev_timeline0) cudaEventDestroy(s->ev_timeline0);
    if (s->ev_run0) cudaEventDestroy(s->ev_run0);
    if (s->ev_run1) cudaEventDestroy(s->ev_run1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int ssb_set_profiling(ssb_sim* s, int on) {
    if (!s) return fail(-1, "ssb_set_profiling: null handle");
    if (collect_profile(s)) return -2;
    s->profiling = on != 0;
    s->timeline = on == 2;
    if (s->timeline) {
        SSB_CUDA(cudaSetDevice(s->device));
        if (!s->ev_timeline0) SSB_CUDA(cudaEventCreate(&s->ev_timeline0));
        SSB_CUDA(cudaStreamSynchronize(s->stream));
        SSB_CUDA(cudaEventRecord(s->ev_timeline0, s->stream));
    }
    for (int k = 0; k < K_NKINDS; ++k) {
        s->kind_ms[k] = 0.f;
        s->kind_launches[k] = 0;
    }
    return 0;
}

int ssb_last_run_ms(ssb_sim* s, float* ms) {
    if (!s || !ms) return fail(-1, "ssb_last_run_ms: bad arguments");
    if (!s->run_timed) return fail(-1, "ssb_last_run_ms: no run recorded");
    SSB_CUDA(cudaSetDevice(s->device));
    SSB_CUDA(cudaEventSynchronize(s->ev_run1));
    SSB_CUDA(cudaEventElapsedTime(ms, s->ev_run0, s->ev_run1));
    return 0;
}

int ssb_kernel_times(ssb_sim* s, float* ms_per_kind, long long* launches_per_kind, int n_kinds) {
    if (!s) return fail(-1, "ssb_kernel_times: null handle");
    if (collect_profile(s)) return -2;
    for (int k = 0; k < n_kinds && k < K_NKINDS; ++k) {
        if (ms_per_kind) ms_per_kind[k] = s->kind_ms[k];
        if (launches_per_kind) launches_per_kind[k] = s->kind_launches[k];
    }
    return 0;
}

long long ssb_total_launches(ssb_sim* s) { return s ? s->total_launches : -1; }

int ssb_timeline(ssb_sim* s, float* start_ms, float* end_ms, int* kinds, int max_n, int* n_out) {
    if (!s || !n_out) return fail(-1, "ssb_timeline: bad arguments");
    if (collect_profile(s)) return -2;
    const int n = (int)s->tl_kind.size();
    *n_out = n;
    for (int i = 0; i < n && i < max_n; ++i) {
        if (start_ms) start_ms[i] = s->tl_start[i];
        if (end_ms) end_ms[i] = s->tl_end[i];
        if (kinds) kinds[i] = s->tl_kind[i];
    }
    return 0;
}

int ssb_mark(ssb_sim* s, int slot) {
    if (!s || slot < 0 || slot >= 4) return fail(-1, "ssb_mark: bad arguments");
    SSB_CUDA(cudaSetDevice(s->device));
    if (!s->ev_mark[slot]) SSB_CUDA(cudaEventCreate(&s->ev_mark[slot]));
    SSB_CUDA(cudaEventRecord(s->ev_mark[slot], s->stream));
    return 0;
}

int ssb_mark_elapsed_ms(ssb_sim* s, int slot_a, int slot_b, float* ms) {
    if (!s || !ms || slot_a < 0 || slot_a >= 4 || slot_b < 0 || slot_b >= 4 || !s->ev_mark[slot_a] || !s->ev_mark[slot_b])
        return fail(-1, "ssb_mark_elapsed_ms: bad arguments or unrecorded mark");
    SSB_CUDA(cudaSetDevice(s->device));
    SSB_CUDA(cudaEventSynchronize(s->ev_mark[slot_b]));
    SSB_CUDA(cudaEventElapsedTime(ms, s->ev_mark[slot_a], s->ev_mark[slot_b]));
    return 0;
}

void* ssb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "ssb_host_alloc: cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void ssb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------ stand-alone SSP kernels
int ssb_ssp_encode(int device, const double* a_scaled, const double* x, double* out, long long n_points, int n, int d) {
    if (!a_scaled || !x || !out || n_points < 0 || n <= 0 || d <= 0) return fail(-1, "ssb_ssp_encode: bad arguments");
    if (n_points == 0) return 0;
    SSB_CUDA(cudaSetDevice(device));
    double *dA = nullptr, *dx = nullptr, *dout = nullptr;
    SSB_CUDA(cudaMalloc((void**)&dA, (size_t)d * n * sizeof(double)));
    SSB_CUDA(cudaMalloc((void**)&dx, (size_t)n_points * n * sizeof(double)));
    SSB_CUDA(cudaMalloc((void**)&dout, (size_t)n_points * d * sizeof(double)));
    SSB_CUDA(cudaMemcpy(dA, a_scaled, (size_t)d * n * sizeof(double), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaMemcpy(dx, x, (size_t)n_points * n * sizeof(double), cudaMemcpyHostToDevice));
    const long long max_grid = 1 << 30;
    for (long long p0 = 0; p0 < n_points; p0 += max_grid) {
        const long long np = std::min(max_grid, n_points - p0);
        k_ssp_encode<<<(unsigned)np, 128, 2 * d * sizeof(double)>>>(dA, dx + p0 * n, dout + p0 * d, np, n, d);
    }
    SSB_CUDA(cudaGetLastError());
    SSB_CUDA(cudaMemcpy(out, dout, (size_t)n_points * d * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dA);
    cudaFree(dx);
    cudaFree(dout);
    return 0;
}

int ssb_ssp_decode_argmax(int device, const double* sample_ssps, const double* queries, int* idx_out, long long n_q,
                          long long n_samples, int d) {
    if (!sample_ssps || !queries || !idx_out || n_q < 0 || n_samples <= 0 || d <= 0)
        return fail(-1, "ssb_ssp_decode_argmax: bad arguments");
    if (n_q == 0) return 0;
    if (n_samples >= 0x7fffffff) return fail(-1, "ssb_ssp_decode_argmax: too many samples");
    SSB_CUDA(cudaSetDevice(device));
    const int dpad = (d + 3) / 4 * 4;
    const int G = (int)n_samples;
    // grid in float32 (padded rows) for the scan, float64 for the near-tie re-score
    std::vector<float> s32((size_t)G * dpad, 0.f);
    for (long long g = 0; g < G; ++g)
        for (int k = 0; k < d; ++k) s32[(size_t)g * dpad + k] = (float)sample_ssps[(size_t)g * d + k];
    float *dS32 = nullptr, *cx = nullptr, *pval = nullptr;
    double *dS64 = nullptr, *dq = nullptr;
    int *pidx = nullptr, *didx = nullptr;
    const int B = (int)std::min<long long>((n_q + 31) / 32 * 32, 1 << 14);
    CleanupDev cd;
    scan_geometry(G, dpad, B / 32, &cd);
    SSB_CUDA(cudaMalloc((void**)&dS32, s32.size() * sizeof(float)));
    SSB_CUDA(cudaMalloc((void**)&dS64, (size_t)G * d * sizeof(double)));
    SSB_CUDA(cudaMalloc((void**)&dq, (size_t)n_q * d * sizeof(double)));
    SSB_CUDA(cudaMalloc((void**)&cx, (size_t)dpad * B * sizeof(float)));
    SSB_CUDA(cudaMalloc((void**)&pval, (size_t)scan_n_cand(cd) * B * sizeof(float)));
    SSB_CUDA(cudaMalloc((void**)&pidx, (size_t)scan_n_cand(cd) * B * sizeof(int)));
    SSB_CUDA(cudaMalloc((void**)&didx, (size_t)B * sizeof(int)));
    SSB_CUDA(cudaMemcpy(dS32, s32.data(), s32.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (cd.tc && build_scan_tiles(s32.data(), G, dpad, &cd)) return -2;
    if (cd.tck && build_scan_tiles_k(s32.data(), G, dpad, B / 32, &cd)) return -2;
    SSB_CUDA(cudaMemcpy(dS64, sample_ssps, (size_t)G * d * sizeof(double), cudaMemcpyHostToDevice));
    SSB_CUDA(cudaMemcpy(dq, queries, (size_t)n_q * d * sizeof(double), cudaMemcpyHostToDevice));
    scan_smem_optin();
    int hdesc[6] = {G, d, dpad, 0, 0, 0};
    int* ddesc = nullptr;
    SSB_CUDA(cudaMalloc((void**)&ddesc, sizeof(hdesc)));
    SSB_CUDA(cudaMemcpy(ddesc, hdesc, sizeof(hdesc), cudaMemcpyHostToDevice));
    SsbCtx c;
    memset(&c, 0, sizeof(c));
    c.G = B / 32;
    cd.cx = cx;
    cd.pval = pval;
    cd.pidx = pidx;
    for (long long q0 = 0; q0 < n_q; q0 += B) {
        const long long nb = std::min<long long>(B, n_q - q0);
        k_decode_prep<<<(B + 127) / 128, 128>>>(dq, cx, n_q, B, d, dpad, q0);
        dispatch_scan(nullptr, false, dpad, B / 32, c, ddesc, dS32, cd, 0);
        // near-ties are re-scored against the float64 grid with the float64 query
        k_cleanup_pick<<<B / 32, 256>>>(d, dpad, scan_n_cand(cd), cx, pval, pidx, dS64, dS32, nullptr, 0, 0, didx, dq,
                                        q0, n_q, scan_eps_floor(cd));
        SSB_CUDA(cudaGetLastError());
        SSB_CUDA(cudaMemcpy(idx_out + q0, didx, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (cd.stc) cudaFree(cd.stc);
    if (cd.stck) cudaFree(cd.stck);
    if (cd.xt) cudaFree(cd.xt);
    cudaFree(dS32);
    cudaFree(dS64);
    cudaFree(dq);
    cudaFree(cx);
    cudaFree(pval);
    cudaFree(pidx);
    cudaFree(didx);
    cudaFree(ddesc);
    return 0;
}

}  // extern "C"
