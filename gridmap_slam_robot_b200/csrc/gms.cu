// gms.cu — libgms.so: the CUDA (sm_100a) implementation of include/gms.h.
// No CPU fallback: every entry point runs the kernels of kernels.cuh or fails with an error code.
// Host-side arithmetic here is limited to what the reference does once per GridMap / Odometry
// construction (grid size, Gaussian taps, log-odds constants, noise sigmas), with the same
// promotions as the Java source (cited inline).
#include "../../include/gms.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#include "kernels.cuh"

using namespace gms;

#define EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_create_err;

struct PhaseSpan {
    int phase;
    cudaEvent_t a, b;
};

// one scan on the device: the table as uploaded + what k_pack_beams derives from it
struct BeamSet {
    unsigned char* blob = nullptr;  // one allocation [xy | dist | hit]: a host scan arrives with ONE H2D copy
    int cap = 0;
    double2* xy = nullptr;     // {localX, localY} of every beam (private copy: the map integration reads it later)
    double* dist = nullptr;    // Measurement.distance (metres)
    uint8_t* hit = nullptr;    // wasHit     (xy / dist / hit are views into blob: packed for B beams after an
                               //             upload, at the capacity offsets otherwise)
    void view(int B) {
        xy = reinterpret_cast<double2*>(blob);
        dist = reinterpret_cast<double*>(blob + (size_t)B * 16);
        hit = blob + (size_t)B * 24;
    }
    float* meas = nullptr;     // (float) distance / resolution (GridMap.java:188)
    double2* hit_xy = nullptr; // compacted hit beams (scoring reads only those, GridMap.java:269-270)
    int* num_hit = nullptr;
    double* rmax2 = nullptr;   // squared length of the longest hit beam (range guard of k_score_sorted<G, 2>)
};

}  // namespace

struct gms_handle {
    gms_config cfg{};
    Geometry g{};
    int W = 0, H = 0, P = 0, lo = 0, cnt = 0, S = 0, dev = 0, resample_mode = 0;
    size_t cells = 0;
    float world_w = 0, world_h = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // shared map: side streams let independent chains of one step overlap (likelihood refresh next to
    // motion + heading sort; map integration next to resampling); joined back before anything depends on them
    cudaStream_t side_a = nullptr, side_b = nullptr;
    cudaEvent_t ev_fork_a = nullptr, ev_done_a = nullptr, ev_fork_b = nullptr, ev_done_b = nullptr;
    bool overlap = true;  // GMS_NO_OVERLAP=1 serialises everything on one stream
    bool b_pending = false;  // a map integration is still running on side_b (joined by the next consumer)
    bool coop = true;     // cooperative launches available (always on sm_100; GMS_NO_COOP=1 is a debugging aid)
    // particle state (double buffered for resampling)
    float4* pose[2] = {nullptr, nullptr};
    double* w[2] = {nullptr, nullptr};
    double* lw[2] = {nullptr, nullptr};
    int* slot[2] = {nullptr, nullptr};
    int cur = 0, slot_cur = 0;
    int* parents = nullptr;
    void* cdf = nullptr;
    // maps
    CellCounts* counts = nullptr;
    double* lik = nullptr;
    double* fac = nullptr;  // shared map only: per-cell scoring factor (GridMap.java:284-288)
    alignas(64) CUtensorMap lik_tmap{};  // 3-D tensor map of the counter arena {W, H, S} for k_likelihood_tma
    bool lik_tma = false;
    int4* rect = nullptr;        // explored bounding box per slot
    uint32_t* dirty = nullptr;   // dirty-tile bitmap per slot: the PENDING set (read by the next likelihood refresh)
    uint32_t* dirty_alt = nullptr;  // shared map with the self-listing refresh: the buffer the last refresh consumed
    bool alt_needs_clear = false;
    bool poses_sharded = false;  // peer exchange: the current pose array holds only this rank's block up to date
    // Per-particle maps: likelihoodData is VIRTUAL (kernels.cuh, k_score_pp): the step scores against the counters,
    // no field is stored.  What Java's array would hold is described by `field_state` for every local slot at once,
    // plus a few per-slot overrides left by the per-map operator entry points; getters materialise it on demand.
    //   FIELD_ZERO        never computed since reset (GridMap.java:110-111: all 0.0)
    //   FIELD_FRESH       blur(codes of the slot's counters now)
    //   FIELD_BEFORE_LAST blur(codes of the counters as they were before the last step's scan was integrated):
    //                     computeLikelihoodMap runs before integrateObservation (SLAM.java:93-105), so that is what
    //                     the array holds between two updates; recovered by subtracting that scan's increments
    //                     (same pose, same beam table, same kernel => same cells) from a scratch copy
    enum { FIELD_ZERO = 0, FIELD_FRESH = 1, FIELD_BEFORE_LAST = 2 };
    int field_state = FIELD_ZERO;
    int field_B = 0;              // beams of the scan a FIELD_BEFORE_LAST field excludes (it is cur_beams())
    struct FieldOverride { bool fresh; double* snap; };  // fresh: blur(counters now); else an explicit copy
    std::unordered_map<int, FieldOverride> field_ovr;    // slot -> override
    CellCounts* fld_counts = nullptr;  // scratch counter map for the subtraction (allocated on first use)
    // Deferred integration (per-particle maps): an update leaves integrateObservation (SLAM.java:103-105) PENDING.  If a
    // resampling comes next, only the particles it selected as parents are integrated (the others are dropped with
    // their maps: dead work); anything else that reads or writes a map, and the next update, integrates all first.
    bool force_resample = false;       // inside gms_resample: the selection kernels ignore Stats.do_resample
    bool integ_pending = false;
    bool defer_integration = true;     // GMS_DEFER_INTEGRATION=0: integrate inside every update (round-1 order)
    int field_bset = 0;                // beam-table set a FIELD_BEFORE_LAST field refers to
    int integ_B = 0, integ_pose_buf = 0, integ_slot_buf = 0, integ_bset = 0;
    int* used = nullptr;               // [0] = n, [1..n] = local particles the last resampling selected as parents
    unsigned long long mapseq = 0;     // per-particle maps across ranks: sequence number of k_maps_final
    float4* upd_pose[2] = {nullptr, nullptr};  // poses the last scan was integrated from, once gms_set_poses has
    bool use_upd_pose = false;                 // replaced the live ones (otherwise those are still pose[cur])
    bool self_list = false;      // shared map, bitmap small enough: the refresh builds its own work list
    int* word_off = nullptr;
    int2* tile_list = nullptr;
    int tiles_per_map = 0;
    int4* dup_rect = nullptr;
    int *dup_src = nullptr, *dup_dst = nullptr, *dup_src_rank = nullptr, *dup_level = nullptr, *scratch2p = nullptr;
    // peer mappings of every rank's arenas (cudaIpc; one process per GPU on one node)
    PeerTable peers{};
    bool peers_ready = false;
    void* ipc_opened[kMaxRanks][GMS_IPC_NUM_HANDLES] = {};
    // peer exchange of the log-weights: double-buffered receive buffers + arrival flags + peer pose arrays
    double* xlw[2] = {nullptr, nullptr};
    unsigned long long* xflags = nullptr;
    unsigned* xticket = nullptr;
    double* peer_xlw[2][kMaxRanks] = {};
    unsigned long long* peer_flags[kMaxRanks] = {};
    const float4* peer_pose[2][kMaxRanks] = {};
    unsigned long long xseq = 0;
    bool direct = false;
    // sharded normalise / resample (multi-rank shared map on the peer path): exchange areas (double-buffered by step
    // parity), flags of the four rounds, and peer views of the arrays a remote parent is read from
    unsigned char* xarea[2] = {nullptr, nullptr};
    unsigned long long* xflags4 = nullptr;
    int coarse_cap = 0;
    size_t xarea_bytes = 0;
    unsigned char* peer_xarea[2][kMaxRanks] = {};
    unsigned long long* peer_xflags4[kMaxRanks] = {};
    const unsigned long long* peer_cdf[kMaxRanks] = {};
    const double* peer_w[2][kMaxRanks] = {};
    const double* peer_lw[2][kMaxRanks] = {};
    const int* peer_parents[kMaxRanks] = {};
    bool exact_sums = false;     // GMS_SHARDED=1 (any handle): the normalise with exact 128-bit sums, so that a sharded
                                 // multi-rank run and a single-rank run agree bit for bit
    bool sharded_post = false;   // enabled by gms_ipc_import for shared maps (GMS_SHARDED=0 keeps the replicated path)
    bool small_fused = true;     // LITERAL resampling of <= 2048 particles on one rank as one launch (GMS_SMALL_FUSED=0: three)
    bool xpull = true;           // replicated peer exchange as a pull inside k_norm_tiles (XPull); GMS_PULL=0: k_xpush_lw
    bool tile_fx_sharded = false;  // np.fx holds the tile sums of the local block only
    bool blocks_stale = false;   // pose / w / lw / parents hold only this rank's block: getters copy the rest from peers
    // beams: two step sets (the shared-map integration of step N may still read set N%2 on the side stream while
    // step N+1 uploads into the other) + one set for the GridMap operator entry points / the pose-optimiser hook
    int bcap = 0;
    BeamSet bs[2], ops;
    int bset = 0;
    double *raw_angle = nullptr, *raw_dist = nullptr;  // raw sweep for the fused de-skew
    double* d_normals = nullptr;
    // combined-map fusion scratch (allocated on first use)
    double *comb_prod = nullptr, *comb_log = nullptr, *comb_lik = nullptr;
    CellCounts* comb_sign = nullptr;
    uint32_t* comb_dirty = nullptr;
    int* comb_off = nullptr;
    int2* comb_list = nullptr;
    // heading sort for k_score_sorted
    SortBufs sort{};
    // tile partials of normalise / neff / weighted pose
    int ntiles = 0;
    NormPartials np{};
    double* wp_part = nullptr;
    unsigned* wp_counter = nullptr;
    bool tile_fx_valid = false;
    bool wpose_valid = false;  // Stats.weighted_pose is current (computed inside the normalise / resample kernels)
    bool fold_wpose = false;   // host entry points (gms_update, gms_resample): the caller will ask for the weighted pose
                               // right away (GridMapApp.java:192), so it is folded into the step's kernels (+3-5 us
                               // each); the *_dev entry points leave it to gms_get_weighted_pose
    // multi-rank shared map: the step's resampling selects only this rank's children; the rest is selected
    // lazily if a getter asks for the full arrays before the next step (which overwrites them anyway)
    bool resample_partial = false;
    bool partial_force = false;
    double partial_u01 = 0;
    unsigned long long partial_count = 0;
    int pp_warps = 4;           // warps per particle in k_score_pp (GMS_PP_WARPS=4|8)
    bool copy_bulk = true;      // single-rank map copies on the TMA engine (GMS_COPY_BULK=0: per-thread 16-byte copies)
    int copy_chunks = 16;       // CTAs per copied map (GMS_COPY_CHUNKS; 4 / 8 / 16 / 32 -> 0.239 / 0.231 / 0.226 / 0.240 ms, K2pp)
    bool score_dynamic = false;  // k_score_sorted draws its work items from a counter when they exceed the resident warps
                                // (GMS_SCORE_DYNAMIC=1; measured no faster at 100k particles: 0.110 vs 0.108 ms)
    unsigned* score_work = nullptr;
    int score_g = 0;  // sub-threads per particle in k_score_sorted (0 = automatic; GMS_SCORE_G overrides: tuning knob)
    int score_v = 7;  // variant of k_score_sorted (GMS_SCORE_V=0..7, identical cell indices; kernels.cuh / DESIGN.md §4.19):
                      // 2 = folded magic constant + per-particle range guard + padded factor field; 7 = 2 with the index
                      // validation amortised over the eight lookups of a batch (measured fastest: 0.102 vs 0.108 ms, K4)
    int num_sms = 148;
    int* ray_maxlen = nullptr;
    // GMS_UPDATE_SORTED scratch (allocated on first use)
    uint32_t *sk_keys = nullptr, *sk_sorted = nullptr, *sk_unique = nullptr;
    int *sk_runlen = nullptr, *sk_nruns = nullptr;
    void* sk_temp = nullptr;
    size_t sk_temp_bytes = 0, sk_cap = 0;
    // shared-map update: recorded ray cells
    uint32_t* ray_cells = nullptr;
    int* ray_count = nullptr;
    float2* ray_start = nullptr;
    int ray_cap = 0;
    // exchange (collective path), scratch, stats
    ExchangeRec *xlocal = nullptr, *xglobal = nullptr;
    void* d_tmp = nullptr;  // max(cells*8, P*24) bytes
    size_t d_tmp_bytes = 0;
    float4* tmp_pose = nullptr;
    int* tmp_slot = nullptr;
    double* tmp_lw = nullptr;
    Stats* st = nullptr;
    Stats* h_st = nullptr;  // pinned
    // pinned staging ring for the host-buffer entry points: slot k is reused only after the copy enqueued from
    // it has completed (one event per slot), so an upload never synchronises the stream
    static constexpr int kStageSlots = 4;
    unsigned char* h_stage[kStageSlots] = {};
    size_t h_stage_bytes[kStageSlots] = {};
    cudaEvent_t h_stage_ev[kStageSlots] = {};
    int stage_next = 0;
    bool stats_valid = false;
    // A4: optional CPU pose optimiser between motion and scoring (GridMap.findBestPoseOptim); default none = identity
    gms_pose_optimizer_fn opt_fn = nullptr;
    void* opt_user = nullptr;
    float* h_opt_poses = nullptr;  // pinned, 3 * cnt
    // step bookkeeping
    uint64_t step = 0, resample_count = 0;
    bool have_update = false, pending = false;
    double pend_dtheta = 0;
    int pend_B = 0;
    // profiling
    bool profile = false;
    std::vector<PhaseSpan> spans;
    std::vector<cudaEvent_t> pool;
    double phase_ms[GMS_PHASE_COUNT] = {0};
    int64_t phase_launches[GMS_PHASE_COUNT] = {0};
    int64_t launches = 0;
    std::string err;
};

namespace {

int flush_integration(gms_handle* h);

int fail(gms_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
int cuda_fail(gms_handle* h, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    (void)cudaGetLastError();
    return fail(h, e == cudaErrorMemoryAllocation ? GMS_ERR_OOM : GMS_ERR_CUDA, m);
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);     \
    } while (0)
// ENTER: every entry point; joins a map integration still running on the side stream.  ENTER_STEP: the
// step entry points, which hand that dependency to the likelihood chain instead (step_begin).
#define ENTER_STEP(h)                                              \
    if (!(h)) return GMS_ERR_INVALID_ARG;                          \
    CK(cudaSetDevice((h)->dev))
#define ENTER_KEEP(h)                                              \
    ENTER_STEP(h);                                                 \
    if ((h)->b_pending) {                                          \
        CK(cudaStreamWaitEvent((h)->stream, (h)->ev_done_b, 0));   \
        (h)->b_pending = false;                                    \
    }
// ENTER_KEEP: entry points that neither read nor write a map (resampling, particle-array getters, bookkeeping): a
// pending integration of per-particle maps stays pending.  ENTER: everything else integrates first.
#define ENTER(h)                                                   \
    ENTER_KEEP(h);                                                 \
    if ((h)->integ_pending) {                                      \
        int rc_flush_ = flush_integration(h);                      \
        if (rc_flush_) return rc_flush_;                           \
    }

struct Phase {  // RAII: CUDA events around one phase when profiling is on
    gms_handle* h;
    int phase;
    cudaEvent_t a = nullptr, b = nullptr;
    Phase(gms_handle* h_, int p) : h(h_), phase(p) {
        if (!h->profile) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!h->pool.empty()) { e = h->pool.back(); h->pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        a = get(); b = get();
        cudaEventRecord(a, h->stream);
    }
    ~Phase() {
        if (!a) return;
        cudaEventRecord(b, h->stream);
        h->spans.push_back({phase, a, b});
    }
};
inline void count_launch(gms_handle* h, int phase) {
    h->launches++;
    h->phase_launches[phase]++;
}
#define LAUNCH(phase, ...)                                         \
    do {                                                           \
        __VA_ARGS__;                                               \
        count_launch(h, phase);                                    \
        CK(cudaGetLastError());                                    \
    } while (0)
// cooperative launch (grid-wide barriers inside the kernel): every CTA of the grid is co-resident
#define LAUNCH_COOP(phase, kernel, grid, block, smem, ...)                                               \
    do {                                                                                                 \
        void* args_[] = {__VA_ARGS__};                                                                   \
        CK(cudaLaunchCooperativeKernel((const void*)(kernel), dim3(grid), dim3(block), args_, smem, h->stream)); \
        count_launch(h, phase);                                                                          \
    } while (0)

int flush_profile(gms_handle* h) {
    if (h->spans.empty()) return GMS_OK;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->side_a));  // phases of the shared-map step are timed on the stream they run on
    CK(cudaStreamSynchronize(h->side_b));
    for (auto& s : h->spans) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, s.a, s.b));
        h->phase_ms[s.phase] += ms;
        h->pool.push_back(s.a);
        h->pool.push_back(s.b);
    }
    h->spans.clear();
    return GMS_OK;
}

inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// Util.generateGaussianKernel Util.java:428-455 (host, once per GridMap: GridMap.java:94-95)
void gaussian_kernel(double sigma, int size, double* values) {
    const double norm = 1.0 / (std::sqrt(2 * M_PI) * sigma);
    const double coeff = 2 * sigma * sigma;
    double total = 0;
    for (int x = -size; x <= size; x++) {
        const double gv = norm * std::exp((double)(-x * x) / coeff);
        values[x + size] = gv;
        total += gv;
    }
    for (int i = 0; i < 2 * size + 1; i++) values[i] /= total;
}
double log_odds(double p) { return std::log(p / (1.0 - p)); }  // Util.java:35-37 (1.0f - odds promotes)
int java_d2i_host(double d) {
    if (d != d) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return -2147483647 - 1;
    return (int)d;
}

void free_beamset(BeamSet& b) {
    cudaFree(b.blob); cudaFree(b.meas); cudaFree(b.hit_xy); cudaFree(b.num_hit);
    cudaFree(b.rmax2);
    b = BeamSet{};
}

void free_all(gms_handle* h) {
    if (!h) return;
    cudaSetDevice(h->dev);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->side_a) cudaStreamSynchronize(h->side_a);
    if (h->side_b) cudaStreamSynchronize(h->side_b);
    for (int q = 0; q < kMaxRanks; q++)
        for (int k = 0; k < GMS_IPC_NUM_HANDLES; k++)
            if (h->ipc_opened[q][k]) cudaIpcCloseMemHandle(h->ipc_opened[q][k]);
    cudaFree(h->xlw[0]); cudaFree(h->xlw[1]); cudaFree(h->xflags); cudaFree(h->xticket);
    cudaFree(h->xarea[0]); cudaFree(h->xarea[1]); cudaFree(h->xflags4);
    for (int i = 0; i < 2; i++) { cudaFree(h->pose[i]); cudaFree(h->w[i]); cudaFree(h->lw[i]); cudaFree(h->slot[i]); }
    cudaFree(h->parents); cudaFree(h->cdf); cudaFree(h->counts); cudaFree(h->lik); cudaFree(h->fac); cudaFree(h->rect);
    cudaFree(h->score_work);
    cudaFree(h->fld_counts); cudaFree(h->upd_pose[0]); cudaFree(h->upd_pose[1]); cudaFree(h->used);
    for (auto& kv : h->field_ovr) cudaFree(kv.second.snap);
    cudaFree(h->dirty); cudaFree(h->dirty_alt); cudaFree(h->word_off); cudaFree(h->tile_list); cudaFree(h->dup_rect);
    cudaFree(h->dup_src_rank); cudaFree(h->dup_level); cudaFree(h->dup_src); cudaFree(h->dup_dst); cudaFree(h->scratch2p);
    free_beamset(h->bs[0]); free_beamset(h->bs[1]); free_beamset(h->ops);
    cudaFree(h->d_normals); cudaFree(h->xlocal); cudaFree(h->xglobal);
    cudaFree(h->ray_cells); cudaFree(h->ray_count); cudaFree(h->ray_start); cudaFree(h->ray_maxlen);
    cudaFree(h->sk_keys); cudaFree(h->sk_sorted); cudaFree(h->sk_unique); cudaFree(h->sk_runlen); cudaFree(h->sk_nruns);
    cudaFree(h->sk_temp);
    cudaFree(h->raw_angle); cudaFree(h->raw_dist); cudaFree(h->comb_prod); cudaFree(h->comb_log); cudaFree(h->comb_lik); cudaFree(h->comb_sign);
    cudaFree(h->comb_dirty); cudaFree(h->comb_off); cudaFree(h->comb_list);
    cudaFree(h->sort.hist); cudaFree(h->sort.chunk_total); cudaFree(h->sort.offs); cudaFree(h->sort.key);
    cudaFree(h->sort.rank); cudaFree(h->sort.order);
    cudaFree(h->np.m); cudaFree(h->np.idx); cudaFree(h->np.s); cudaFree(h->np.ws); cudaFree(h->np.q); cudaFree(h->np.fx);
    cudaFree(h->np.x128); cudaFree(h->np.counter); cudaFree(h->wp_part); cudaFree(h->wp_counter);
    cudaFree(h->d_tmp); cudaFree(h->tmp_pose); cudaFree(h->tmp_slot); cudaFree(h->tmp_lw); cudaFree(h->st);
    if (h->h_st) cudaFreeHost(h->h_st);
    if (h->h_opt_poses) cudaFreeHost(h->h_opt_poses);
    for (int k = 0; k < gms_handle::kStageSlots; k++) {
        if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
        if (h->h_stage_ev[k]) cudaEventDestroy(h->h_stage_ev[k]);
    }
    for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : h->pool) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_a) cudaStreamDestroy(h->side_a);
    if (h->side_b) cudaStreamDestroy(h->side_b);
    for (cudaEvent_t e : {h->ev_fork_a, h->ev_done_a, h->ev_fork_b, h->ev_done_b})
        if (e) cudaEventDestroy(e);
    delete h;
}

// Next slot of the pinned staging ring, at least `bytes` large.  Blocks only if the copy enqueued from this
// slot kStageSlots uploads ago has not finished yet (it has: a step lies in between).
int stage_acquire(gms_handle* h, size_t bytes, unsigned char** out, int* slot) {
    const int k = h->stage_next;
    h->stage_next = (k + 1) % gms_handle::kStageSlots;
    if (!h->h_stage_ev[k]) CK(cudaEventCreateWithFlags(&h->h_stage_ev[k], cudaEventDisableTiming));
    else CK(cudaEventSynchronize(h->h_stage_ev[k]));
    if (bytes > h->h_stage_bytes[k]) {
        if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
        h->h_stage[k] = nullptr;
        h->h_stage_bytes[k] = 0;
        const size_t n = bytes + bytes / 2 + 4096;
        CK(cudaMallocHost((void**)&h->h_stage[k], n));
        h->h_stage_bytes[k] = n;
    }
    *out = h->h_stage[k];
    *slot = k;
    return GMS_OK;
}
int stage_release(gms_handle* h, int slot) {  // after the last copy from the slot has been enqueued
    CK(cudaEventRecord(h->h_stage_ev[slot], h->stream));
    return GMS_OK;
}

int alloc_beamset(gms_handle* h, BeamSet& b, int cap) {
    CK(cudaMalloc((void**)&b.blob, (size_t)cap * 25));
    b.cap = cap;
    b.view(cap);
    CK(cudaMalloc((void**)&b.meas, (size_t)cap * 4));
    CK(cudaMalloc((void**)&b.hit_xy, (size_t)cap * 16));
    CK(cudaMalloc((void**)&b.num_hit, 4));
    CK(cudaMemset(b.num_hit, 0, 4));
    CK(cudaMalloc((void**)&b.rmax2, 8));
    CK(cudaMemset(b.rmax2, 0, 8));
    return GMS_OK;
}

int ensure_beams(gms_handle* h, int B) {
    if (B <= h->bcap) return GMS_OK;
    if (B > GMS_MAX_BEAMS) return fail(h, GMS_ERR_INVALID_ARG, "too many beams (GMS_MAX_BEAMS)");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->side_a));
    CK(cudaStreamSynchronize(h->side_b));
    h->b_pending = false;
    if (int rc_ = flush_integration(h)) return rc_;  // a pending integration reads the tables that are about to go
    CK(cudaStreamSynchronize(h->stream));
    // a virtual per-particle field in state FIELD_BEFORE_LAST is defined through the last step's beam table: carry it over
    const bool keep_scan = h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->field_state == gms_handle::FIELD_BEFORE_LAST;
    BeamSet old_scan = h->bs[h->field_bset];
    if (keep_scan) h->bs[h->field_bset] = BeamSet{};
    free_beamset(h->bs[0]); free_beamset(h->bs[1]); free_beamset(h->ops);
    cudaFree(h->ray_cells); cudaFree(h->ray_count); cudaFree(h->ray_start);
    cudaFree(h->raw_angle); cudaFree(h->raw_dist);
    h->raw_angle = h->raw_dist = nullptr;
    h->ray_cells = nullptr; h->ray_count = nullptr; h->ray_start = nullptr;
    h->bcap = 0;
    const int cap = ((B + 255) / 256) * 256;
    int rc;
    for (BeamSet* b : {&h->bs[0], &h->bs[1], &h->ops})
        if ((rc = alloc_beamset(h, *b, cap))) return rc;
    if (keep_scan) {
        BeamSet& nb = h->bs[h->field_bset];
        const size_t n = (size_t)h->field_B;
        CK(cudaMemcpy(nb.xy, old_scan.xy, n * 16, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(nb.hit, old_scan.hit, n, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(nb.meas, old_scan.meas, n * 4, cudaMemcpyDeviceToDevice));
        free_beamset(old_scan);
    }
    CK(cudaMalloc((void**)&h->raw_angle, (size_t)cap * 8));
    CK(cudaMalloc((void**)&h->raw_dist, (size_t)cap * 8));
    if (h->cfg.map_mode == GMS_MAP_SHARED) {
        // a ray visits at most W + H - 1 in-bounds cells (4-connected, monotone) + the extra steps
        h->ray_cap = h->W + h->H + h->cfg.extra_steps + 4;
        CK(cudaMalloc((void**)&h->ray_cells, (size_t)cap * h->ray_cap * 4));
        CK(cudaMalloc((void**)&h->ray_count, (size_t)cap * 4));
        CK(cudaMalloc((void**)&h->ray_start, sizeof(float2)));
    }
    h->bcap = cap;
    return GMS_OK;
}

// argument check shared by the step entry points; runs BEFORE flip_beams so that a rejected call leaves the
// beam-table parity (and with it the side stream's read set) untouched
int check_beams(gms_handle* h, int B, const void* a, const void* b, const void* c, const char* who) {
    if (B < 0 || (B > 0 && (!a || !b || !c))) return fail(h, GMS_ERR_INVALID_ARG, std::string(who) + ": bad beam arrays");
    if (B > GMS_MAX_BEAMS) return fail(h, GMS_ERR_INVALID_ARG, std::string(who) + ": too many beams (GMS_MAX_BEAMS)");
    return GMS_OK;
}

// start of a step: switch to the beam-table set the side stream is not reading
void flip_beams(gms_handle* h) { h->bset ^= 1; }
inline BeamSet& cur_beams(gms_handle* h) { return h->bs[h->bset]; }

int fetch_stats(gms_handle* h) {
    if (h->stats_valid) return GMS_OK;
    CK(cudaMemcpyAsync(h->h_st, h->st, sizeof(Stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_st->xerror)
        return fail(h, GMS_ERR_STATE, "peer exchange: a rank's log-weights did not arrive (rank out of step?); the "
                                      "step was abandoned before it changed any state");
    h->stats_valid = true;
    return GMS_OK;
}

PoseTable pose_table(const gms_handle* h, int buf) {
    PoseTable t{};
    t.cnt = h->cnt;
    for (int q = 0; q < h->cfg.nranks; q++)
        t.p[q] = (h->direct && q != h->cfg.rank) ? h->peer_pose[buf][q] : h->pose[buf];
    return t;
}

// ---- the step, as stream-ordered launches ---------------------------------------------------------
int launch_pack(gms_handle* h, BeamSet& b, const double* d_xy, const double* d_dist, const uint8_t* d_hit, int B) {
    LAUNCH(GMS_PHASE_SCORE, k_pack_beams<<<1, 256, 0, h->stream>>>((const double2*)d_xy, d_dist, d_hit, B, h->g.res_f,
                                                                   b.hit_xy, b.meas, b.xy, b.hit, b.num_hit, b.rmax2));
    return GMS_OK;
}

// blur launch over a work list (k_lik_scan / k_lik_emit) or, with `sl.bitmap`, over the bitmap itself
int launch_blur(gms_handle* h, const CellCounts* counts, double* lik, double* fac, const int2* list, SelfList sl,
                long long max_tiles, bool allow_tma) {
    const int k = h->g.khalf, th = kTileH + 2 * k, tw = (kTileW + 2 * k + 3) & ~3;
    const size_t smem = (size_t)th * kTileW * 8 + (size_t)th * tw * 4 + (sl.bitmap ? (size_t)(sl.nwords + 1) * 4 : 0);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(max_tiles, (long long)h->num_sms * 6));
    if (allow_tma && h->lik_tma) {
        const size_t smem_t = 2 * (((size_t)kTmaTileW * kTmaTileH * 8 + 127) / 128 * 128) + smem;
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_likelihood_tma<<<grid, 256, smem_t, h->stream>>>(h->lik_tmap, lik, fac, list, h->st, sl, h->g));
    } else if (k == 3)
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_likelihood<3><<<grid, 256, smem, h->stream>>>(counts, lik, fac, list, h->st, sl, h->g));
    else
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_likelihood<0><<<grid, 256, smem, h->stream>>>(counts, lik, fac, list, h->st, sl, h->g));
    return GMS_OK;
}

int launch_likelihood(gms_handle* h) {
    Phase ph(h, GMS_PHASE_LIKELIHOOD);
    if (h->self_list) {
        // shared map: ONE launch; the refresh reads the pending bitmap, new marks go to the other buffer
        if (h->alt_needs_clear) CK(cudaMemsetAsync(h->dirty_alt, 0, (size_t)h->g.tile_words * 4, h->stream));
        const SelfList sl{h->dirty, h->g.tile_words};
        int rc = launch_blur(h, h->counts, h->lik, h->fac, nullptr, sl, h->tiles_per_map, true);
        if (rc) return rc;
        std::swap(h->dirty, h->dirty_alt);
        h->alt_needs_clear = true;
        return GMS_OK;
    }
    // one map (the shared map with a bitmap too large for the self-listing refresh)
    const int nwords = h->g.tile_words;
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_scan<<<1, 1024, 0, h->stream>>>(h->dirty, nwords, h->word_off, h->st));
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_emit<<<blocks_for(nwords, 256), 256, 0, h->stream>>>(
                                     h->dirty, nwords, h->g.tile_words, h->word_off, h->tile_list));
    return launch_blur(h, h->counts, h->lik, h->fac, h->tile_list, SelfList{nullptr, 0}, h->tiles_per_map, true);
}

// every tile of ONE counter map -> `out` (GridMap.computeLikelihoodMap on a whole map; per-particle maps materialise
// their virtual field with it, the combined-map fusion its sign map)
int launch_blur_whole(gms_handle* h, const CellCounts* map, double* out, uint32_t* bitmap, int* word_off, int2* list) {
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_fill_dirty<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                     bitmap, 1, h->g.tile_words, h->tiles_per_map));
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_scan<<<1, 1024, 0, h->stream>>>(bitmap, h->g.tile_words, word_off, h->st));
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_emit<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                     bitmap, h->g.tile_words, h->g.tile_words, word_off, list));
    return launch_blur(h, map, out, nullptr, list, SelfList{nullptr, 0}, h->tiles_per_map, false);
}

// Thread-per-particle scoring in heading order pays off when one shared field serves many particles
// (see k_score_sorted); per-particle maps and small particle sets keep one warp per particle.
// sub-threads per particle in k_score_sorted: enough threads to fill the machine (>= ~200k), at most one warp per particle
int score_subthreads(const gms_handle* h, int cnt) {
    if (h->score_g) return h->score_g;
    int G = 1;
    while (G < 32 && (long long)cnt * G < 200000) G *= 2;
    return G;
}
// The heading sort only matters when a warp holds several particles (G < 32): with one warp per particle the
// 32 lanes look up beams of the SAME particle, whatever the order (K3: 10k particles -> no sort, plain k_motion).
bool use_sorted_score(const gms_handle* h) {
    return h->cfg.map_mode == GMS_MAP_SHARED && h->coop && score_subthreads(h, h->cnt) < 32;
}
// every shared-map scoring of the step runs k_score_sorted (factor field + FMA cell index)
bool use_fac_score(const gms_handle* h) { return h->cfg.map_mode == GMS_MAP_SHARED; }

template <int G, int V>
int launch_score_sorted(gms_handle* h, unsigned grid, size_t smem, const float4* pose, int lo, int cnt, const BeamSet& b,
                        const int* order, double* lw, ExchangeRec* xlocal) {
    static bool attr_done[64] = {};  // per instantiation and device: scans above 3072 beams need more than 48 KB
    if (!attr_done[h->dev & 63]) {
        CK(cudaFuncSetAttribute(k_score_sorted<G, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done[h->dev & 63] = true;
    }
    unsigned* work = nullptr;
    if (h->score_dynamic) {  // more items than resident warps: persistent CTAs draw them from a counter
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_sorted<G, V>, 128, smem));
        const unsigned resident = (unsigned)std::max(1, per_sm) * (unsigned)h->num_sms;
        if (grid > resident + resident / 8) {
            CK(cudaMemsetAsync(h->score_work, 0, 4, h->stream));
            work = h->score_work;
            grid = resident;
        }
    }
    LAUNCH(GMS_PHASE_SCORE, k_score_sorted<G, V><<<grid, 128, smem, h->stream>>>(pose, lo, cnt, b.hit_xy, b.num_hit, h->fac,
                                                                                 order, lw, xlocal, b.rmax2, work, h->g));
    return GMS_OK;
}

int launch_score(gms_handle* h, const BeamSet& b, const float4* pose, int lo, int cnt, const int* slot, double* lw,
                 ExchangeRec* xlocal, int B, bool sorted, const double* field) {
    Phase ph(h, GMS_PHASE_SCORE);
    const size_t smem = std::max<size_t>(16, (size_t)B * 16);
    if (sorted) {
        const int G = score_subthreads(h, cnt);
        const int* order = (use_sorted_score(h) && pose == h->pose[h->cur] && cnt == h->cnt) ? h->sort.order : nullptr;
        const unsigned grid = blocks_for((long long)cnt * G, 128);
#define SCORE_G(GG)                                                                                            \
    case GG:                                                                                                   \
        return h->score_v == 7   ? launch_score_sorted<GG, 7>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 6 ? launch_score_sorted<GG, 6>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 5 ? launch_score_sorted<GG, 5>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 4 ? launch_score_sorted<GG, 4>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 3 ? launch_score_sorted<GG, 3>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 2 ? launch_score_sorted<GG, 2>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
               : h->score_v == 1 ? launch_score_sorted<GG, 1>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal) \
                                 : launch_score_sorted<GG, 0>(h, grid, smem, pose, lo, cnt, b, order, lw, xlocal);
        switch (G) {
            SCORE_G(1) SCORE_G(2) SCORE_G(4) SCORE_G(8) SCORE_G(16) SCORE_G(32)
        }
#undef SCORE_G
        return fail(h, GMS_ERR_STATE, "launch_score: bad sub-thread count");
    }
    const unsigned grid = std::min<unsigned>(blocks_for(cnt, 8), (unsigned)h->num_sms * 8);
    if (!field) {  // per-particle maps, field = blur(codes of the counters now): evaluated where it is read
        // warps per particle: a fixed property of the handle's process (never of the particle or rank count), so the
        // reduction shape — and with it every log-weight bit — is the same on every rank
        if (h->g.khalf == 3 && (h->W & 1) == 0) {
            if (h->pp_warps == 8)
                LAUNCH(GMS_PHASE_SCORE, k_score_pp<3, 8><<<cnt, 256, 0, h->stream>>>(pose, lo, cnt, b.hit_xy, b.num_hit, h->counts,
                                                                                      slot, lw, xlocal, h->g));
            else
                LAUNCH(GMS_PHASE_SCORE, k_score_pp<3, 4><<<cnt, 128, 0, h->stream>>>(pose, lo, cnt, b.hit_xy, b.num_hit, h->counts,
                                                                                      slot, lw, xlocal, h->g));
        } else {
            LAUNCH(GMS_PHASE_SCORE, k_score_pp<0, 4><<<cnt, 128, 0, h->stream>>>(pose, lo, cnt, b.hit_xy, b.num_hit, h->counts,
                                                                                  slot, lw, xlocal, h->g));
        }
        return GMS_OK;
    }
    // one stored field (the shared map's, or a materialised per-particle one: `slot` is then null)
    LAUNCH(GMS_PHASE_SCORE, k_score<<<grid, 256, smem, h->stream>>>(pose, lo, cnt, b.hit_xy, b.num_hit, field, slot, lw,
                                                                     xlocal, h->g));
    return GMS_OK;
}

// GMS_UPDATE_SORTED: the atomic-free scatter.  The key count is read back (one 4-byte D2H + sync) so the sort
// covers only the cells the scan produced; the mode exists to be measured against the default.
int launch_sorted_update(gms_handle* h, const BeamSet& b, int Bpad) {
    int maxlen = 0;
    CK(cudaMemcpyAsync(&maxlen, h->ray_maxlen, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const size_t n = (size_t)maxlen * Bpad;
    if (n == 0) return GMS_OK;
    if (n > h->sk_cap) {
        cudaFree(h->sk_keys); cudaFree(h->sk_sorted); cudaFree(h->sk_unique); cudaFree(h->sk_runlen); cudaFree(h->sk_temp);
        h->sk_keys = h->sk_sorted = h->sk_unique = nullptr; h->sk_runlen = nullptr; h->sk_temp = nullptr; h->sk_cap = 0;
        const size_t cap = n + n / 2;
        CK(cudaMalloc((void**)&h->sk_keys, cap * 4));
        CK(cudaMalloc((void**)&h->sk_sorted, cap * 4));
        CK(cudaMalloc((void**)&h->sk_unique, cap * 4));
        CK(cudaMalloc((void**)&h->sk_runlen, cap * 4));
        if (!h->sk_nruns) CK(cudaMalloc((void**)&h->sk_nruns, 4));
        size_t t1 = 0, t2 = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, t1, h->sk_keys, h->sk_sorted, (int)cap, 0, 32, h->stream));
        CK(cub::DeviceRunLengthEncode::Encode(nullptr, t2, h->sk_sorted, h->sk_unique, h->sk_runlen, h->sk_nruns, (int)cap, h->stream));
        h->sk_temp_bytes = std::max(t1, t2);
        CK(cudaMalloc(&h->sk_temp, h->sk_temp_bytes));
        h->sk_cap = cap;
    }
    LAUNCH(GMS_PHASE_MAP_UPDATE, k_ray_keys<<<148 * 4, 256, 0, h->stream>>>(h->ray_cells, Bpad, maxlen, h->ray_count,
                                                                            h->ray_start, b.meas, b.hit, h->sk_keys, h->g));
    size_t tb = h->sk_temp_bytes;
    CK(cub::DeviceRadixSort::SortKeys(h->sk_temp, tb, h->sk_keys, h->sk_sorted, (int)n, 0, 32, h->stream));
    tb = h->sk_temp_bytes;
    CK(cub::DeviceRunLengthEncode::Encode(h->sk_temp, tb, h->sk_sorted, h->sk_unique, h->sk_runlen, h->sk_nruns, (int)n,
                                          h->stream));
    h->launches += 8;  // CUB: 3-4 kernels per call
    LAUNCH(GMS_PHASE_MAP_UPDATE, k_apply_runs<<<148 * 4, 256, 0, h->stream>>>(h->sk_unique, h->sk_runlen, h->sk_nruns,
                                                                              h->counts, h->dirty, h->g));
    return GMS_OK;
}

// shared map: integrate the scan from the strongest pose (Stats.strongest_pose, snapshotted by the normalise kernel)
int launch_shared_update(gms_handle* h, const BeamSet& b, int B) {
    if (B <= 0) return GMS_OK;
    Phase ph(h, GMS_PHASE_MAP_UPDATE);
    const int Bpad = ((B + 31) / 32) * 32;
    const bool sorted = h->cfg.update_mode == GMS_UPDATE_SORTED;
    if (sorted) CK(cudaMemsetAsync(h->ray_maxlen, 0, 4, h->stream));
    uint32_t* stale = h->alt_needs_clear ? h->dirty_alt : nullptr;
    LAUNCH(GMS_PHASE_MAP_UPDATE, k_ray_integrate<<<Bpad / kRaysPerCta, 256, 0, h->stream>>>(
                                     b.xy, B, Bpad, h->st, h->ray_cells, h->ray_cap, h->ray_count, h->ray_start,
                                     h->ray_maxlen, h->rect, b.meas, b.hit, h->counts, h->dirty, stale,
                                     stale ? h->g.tile_words : 0, sorted ? 1 : 0, h->g));
    h->alt_needs_clear = false;
    if (sorted) return launch_sorted_update(h, b, Bpad);
    return GMS_OK;
}

// per-particle maps (or one explicit {pose, slot} pair): one thread per (particle, beam) ray
int launch_map_update(gms_handle* h, const BeamSet& b, const float4* pose, int lo, int cnt, const int* slot, int B,
                      const int* ulist = nullptr, const int* n_used = nullptr) {
    if (B <= 0) return GMS_OK;
    Phase ph(h, GMS_PHASE_MAP_UPDATE);
    const long long total = (long long)cnt * B;
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE) {  // no dirty-tile bookkeeping: fire-and-forget reductions
        // a listed launch may spread its rays 8 to a warp (kSparseRays): cover that mapping too
        const long long threads = ulist ? std::max<long long>(total, std::min<long long>(total, kSparseRays) * 4) : total;
        LAUNCH(GMS_PHASE_MAP_UPDATE, k_map_update_red<false><<<blocks_for(threads, 128), 128, 0, h->stream>>>(
                                         pose, lo, cnt, b.xy, b.meas, b.hit, B, h->counts, slot, h->rect, ulist, n_used,
                                         h->st, h->g));
    }
    else
        LAUNCH(GMS_PHASE_MAP_UPDATE, k_map_update<<<blocks_for(total, 128), 128, 0, h->stream>>>(
                                         pose, lo, cnt, b.xy, b.meas, b.hit, B, h->counts, slot, h->rect, h->dirty, h->g));
    return GMS_OK;
}

// the pending integration of the last update: every local particle, or (`used`) only the parents a resampling selected
int integrate_pending(gms_handle* h, const int* ulist, const int* n_used) {
    if (!h->integ_pending) return GMS_OK;
    h->integ_pending = false;
    int rc = launch_map_update(h, h->bs[h->integ_bset], h->pose[h->integ_pose_buf], h->lo, h->cnt,
                               h->slot[h->integ_slot_buf] + h->lo, h->integ_B, ulist, n_used);
    if (rc) return rc;
    h->field_state = gms_handle::FIELD_BEFORE_LAST;  // likelihoodData now lags the counters by this scan
    h->field_B = h->integ_B;
    h->field_bset = h->integ_bset;
    return GMS_OK;
}
int flush_integration(gms_handle* h) { return integrate_pending(h, nullptr, nullptr); }

// A4 — GridMap.findBestPoseOptim GridMap.java:348-369 as a CPU hook between motion and scoring: the local
// particles' poses go to the host, the callback may replace them, they come back.  Default: no hook (the
// reference's optimiser returns its start pose), nothing of this runs.
int run_pose_optimizer(gms_handle* h, const double* d_xy, const double* d_dist, const uint8_t* d_hit, int B,
                       double d_center, double d_theta) {
    if (!h->h_opt_poses) CK(cudaMallocHost((void**)&h->h_opt_poses, (size_t)h->cnt * 12));
    float* tmp = (float*)h->d_tmp;
    LAUNCH(GMS_PHASE_COUNT - 1, k_pose_unpack<<<blocks_for(h->cnt, 256), 256, 0, h->stream>>>(h->pose[h->cur] + h->lo, tmp, h->cnt));
    CK(cudaMemcpyAsync(h->h_opt_poses, tmp, (size_t)h->cnt * 12, cudaMemcpyDeviceToHost, h->stream));
    // the hook sees host copies of the scan (the *_dev entry points hold it on the device)
    std::vector<double> xy((size_t)2 * B), dist((size_t)B);
    std::vector<uint8_t> hit((size_t)B);
    if (B > 0) {
        CK(cudaMemcpyAsync(xy.data(), d_xy, (size_t)B * 16, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(dist.data(), d_dist, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(hit.data(), d_hit, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    const int rc = h->opt_fn(h->opt_user, h, h->lo, h->cnt, h->h_opt_poses, xy.data(), dist.data(), hit.data(), B, d_center,
                             d_theta);
    if (rc != 0) return fail(h, GMS_ERR_STATE, "pose optimiser hook returned " + std::to_string(rc));
    CK(cudaMemcpyAsync(tmp, h->h_opt_poses, (size_t)h->cnt * 12, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(GMS_PHASE_COUNT - 1, k_pose_pack<<<blocks_for(h->cnt, 256), 256, 0, h->stream>>>(tmp, h->pose[h->cur] + h->lo, h->cnt));
    return GMS_OK;
}

void clear_field_overrides(gms_handle* h) {
    for (auto& kv : h->field_ovr) cudaFree(kv.second.snap);
    h->field_ovr.clear();
}

// replicated peer exchange in its pull form: the fixed-tile normalise reads every rank's block from its owner
bool pulling(const gms_handle* h) {
    return h->cfg.nranks > 1 && h->direct && !h->sharded_post && h->xpull && !h->exact_sums;
}

int step_begin(gms_handle* h, const double* d_xy, const double* d_dist, const uint8_t* d_hit, int B, double d_center,
               double d_theta, const double* d_normals) {
    const gms_config& c = h->cfg;
    h->stats_valid = false;
    h->resample_partial = false;  // this step overwrites every non-local particle of the gathered arrays anyway
    const bool shared = c.map_mode == GMS_MAP_SHARED;
    const bool hook = h->opt_fn != nullptr;
    const bool fork = shared && h->overlap && !hook;
    int rc = flush_integration(h);  // an update without a resampling in between: integrate the previous scan everywhere
    if (rc) return rc;
    rc = ensure_beams(h, B);  // grow the beam tables (if needed) while every stream can still be drained from here
    if (rc) return rc;
    BeamSet& bs = cur_beams(h);
    if (d_xy != (const double*)bs.xy) bs.view(bs.cap);  // device-resident scan: private copies at the capacity offsets
    if (fork) {  // likelihood refresh of the shared map: independent of the beams and of the motion update
        cudaStream_t main = h->stream;
        CK(cudaEventRecord(h->ev_fork_a, main));
        CK(cudaStreamWaitEvent(h->side_a, h->ev_fork_a, 0));
        if (h->b_pending) {  // the previous step's map integration feeds this refresh (and, through it, main)
            CK(cudaStreamWaitEvent(h->side_a, h->ev_done_b, 0));
            h->b_pending = false;
        }
        h->stream = h->side_a;
        rc = launch_likelihood(h);
        h->stream = main;
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_done_a, h->side_a));
    }
    if (h->b_pending) {  // not forking: join the side stream the plain way
        CK(cudaStreamWaitEvent(h->stream, h->ev_done_b, 0));
        h->b_pending = false;
    }
    // Odometry.recalculateStdDev Odometry.java:60-69
    MotionArgs ma{};
    ma.pose = h->pose[h->cur]; ma.lo = h->lo; ma.cnt = h->cnt; ma.normals = d_normals; ma.seed = c.seed; ma.step = h->step;
    ma.d_center = d_center; ma.d_theta = d_theta;
    ma.sd_c = (c.noise_center_base + std::fabs(d_center) * c.noise_center_gain) / 2;
    ma.sd_t = c.noise_theta_base_deg * (M_PI / 180.0) + c.noise_theta_gain * std::fabs(d_theta);
    const bool sorted = use_sorted_score(h);
    bool packed = false;
    {
        Phase ph(h, GMS_PHASE_MOTION);
        PackArgs pk{(const double2*)d_xy, d_dist, d_hit, B, h->g.res_f, bs.hit_xy, bs.meas, bs.xy, bs.hit, bs.num_hit, bs.rmax2};
        int do_pack = hook ? 0 : 1;
        if (sorted) {  // motion + heading sort + beam packing: one cooperative launch
            const unsigned grid = std::min<unsigned>(blocks_for(h->cnt, 1024), (unsigned)h->num_sms);
            LAUNCH_COOP(GMS_PHASE_MOTION, k_motion_sort, grid, 1024, 0, &ma, &h->sort, &pk, &do_pack);
        } else {       // motion + beam packing
            LAUNCH(GMS_PHASE_MOTION, k_motion<<<blocks_for(h->cnt, 256), 256, 0, h->stream>>>(ma, pk, do_pack));
        }
        packed = do_pack != 0;
    }
    if (!shared) {
        // computeLikelihoodMap of every particle (SLAM.java:93): the field is virtual, so this is bookkeeping only —
        // from here on every slot's likelihoodData is blur(codes of its counters now)
        clear_field_overrides(h);
        h->field_state = gms_handle::FIELD_FRESH;
        h->use_upd_pose = false;
    } else if (!fork) {  // computeLikelihoodMap precedes findBestPoseOptim and probabilityOf (SLAM.java:93-99)
        rc = launch_likelihood(h);
        if (rc) return rc;
    }
    if (hook) {
        rc = run_pose_optimizer(h, d_xy, d_dist, d_hit, B, d_center, d_theta);
        if (rc) return rc;
    }
    if (!packed) {
        rc = launch_pack(h, bs, d_xy, d_dist, d_hit, B);
        if (rc) return rc;
    }
    if (fork) CK(cudaStreamWaitEvent(h->stream, h->ev_done_a, 0));
    // (pull exchange: the log-weights go straight into this rank's exchange buffer of the coming exchange; the
    // normalise files every rank's values, this rank's included, in lw[cur])
    rc = launch_score(h, bs, h->pose[h->cur], h->lo, h->cnt, shared ? nullptr : h->slot[h->slot_cur] + h->lo,
                      pulling(h) ? h->xlw[(h->xseq + 1) & 1] : h->lw[h->cur],
                      (c.nranks > 1 && !h->direct) ? h->xlocal : nullptr, B, use_fac_score(h), shared ? h->lik : nullptr);
    if (rc) return rc;

    const bool skip = std::fabs(d_theta) > (M_PI / 180.0) * c.skip_update_deg;  // SLAM.java:82
    if (!shared && !skip && B > 0) {
        // integrateObservation (SLAM.java:103-105) is left pending: see gms_handle::integ_pending
        h->integ_pending = true;
        h->integ_B = B; h->integ_pose_buf = h->cur; h->integ_slot_buf = h->slot_cur; h->integ_bset = h->bset;
        if (!h->defer_integration && (rc = flush_integration(h))) return rc;
    }
    if (c.nranks > 1 && h->direct && !h->sharded_post && !pulling(h)) {
        // this rank's log-weights -> every rank's receive buffer (NVLink stores), then one flag per receiver.  With
        // per-particle maps the flag also certifies that this rank's maps are final for the step (peers may pull
        // them when resampling), so the push is enqueued after the map integration.
        Phase ph(h, GMS_PHASE_EXCHANGE);
        XPush xp{};
        xp.nranks = c.nranks;
        for (int q = 0; q < xp.nranks; q++) {
            xp.dst[q] = q == c.rank ? h->xlw[(h->xseq + 1) & 1] : h->peer_xlw[(h->xseq + 1) & 1][q];
            xp.flag[q] = q == c.rank ? h->xflags : h->peer_flags[q];
        }
        const unsigned grid = std::max(1u, std::min<unsigned>(blocks_for((long long)c.nranks * h->cnt / 2, 1024 * 4), 64u));
        LAUNCH(GMS_PHASE_EXCHANGE, k_xpush_lw<<<grid, 1024, 0, h->stream>>>(h->lw[h->cur], h->lo, h->cnt, xp, c.rank,
                                                                          h->xseq + 1, h->xticket));
    }
    h->pending = true;
    h->pend_dtheta = d_theta;
    h->pend_B = B;
    h->poses_sharded = c.nranks > 1 && h->direct;
    return GMS_OK;
}

SelectArgs select_args(gms_handle* h, int from, int to, double u01, unsigned long long count, int m_begin, int m_count) {
    SelectArgs a{};
    a.cdf = h->cdf; a.P = h->P; a.u01 = u01; a.seed = h->cfg.seed; a.resample_count = count;
    a.parents = h->parents; a.st = h->st; a.poses_in = pose_table(h, from);
    a.w_in = h->w[from]; a.lw_in = h->lw[from];
    a.pose_out = h->pose[to]; a.w_out = h->w[to]; a.lw_out = h->lw[to];
    a.m_begin = m_begin; a.m_count = m_count;
    a.force = h->force_resample ? 1 : 0;
    a.wp_part = nullptr; a.wp_counter = nullptr;
    return a;
}

// children [m_begin, m_begin + m_count) of the resampling whose CDF is in h->cdf; in = buffers `from`, out = `to`
int launch_select(gms_handle* h, int from, int to, double u01, unsigned long long count, int m_begin, int m_count) {
    if (m_count <= 0) return GMS_OK;
    const SelectArgs a = select_args(h, from, to, u01, count, m_begin, m_count);
    if (h->resample_mode == GMS_RESAMPLE_FIXED)
        LAUNCH(GMS_PHASE_RESAMPLE, k_select<true><<<blocks_for(m_count, 256), 256, 0, h->stream>>>(a));
    else
        LAUNCH(GMS_PHASE_RESAMPLE, k_select<false><<<blocks_for(m_count, 256), 256, 0, h->stream>>>(a));
    return GMS_OK;
}

// Peer exchange: between an update and the next full gather every rank holds only its own block of the moved
// poses — and, with the sharded normalise / resample, of the weights, log-weights and parent indices.  A getter
// that needs all of them copies the other blocks out of their owners' arrays (peer mappings).  The caller keeps
// the ranks between steps while it reads (documented in gms.h).
int materialize_poses(gms_handle* h) {
    if (!h->direct) return GMS_OK;
    if (h->blocks_stale) {
        RemoteBlocks rb{};
        rb.poses = pose_table(h, h->cur);
        for (int q = 0; q < h->cfg.nranks; q++) {
            const bool me = q == h->cfg.rank;
            rb.w[q] = me ? h->w[h->cur] : h->peer_w[h->cur][q];
            rb.lw[q] = me ? h->lw[h->cur] : h->peer_lw[h->cur][q];
            rb.parents[q] = me ? h->parents : h->peer_parents[q];
        }
        rb.cnt = h->cnt; rb.lo = h->lo; rb.P = h->P;
        LAUNCH(GMS_PHASE_COUNT - 1, k_fill_remote_blocks<<<blocks_for(h->P, 256), 256, 0, h->stream>>>(
                                        rb, h->pose[h->cur], h->w[h->cur], h->lw[h->cur], h->parents));
        h->blocks_stale = false;
        h->poses_sharded = false;
        return GMS_OK;
    }
    if (!h->poses_sharded) return GMS_OK;
    const PoseTable t = pose_table(h, h->cur);
    LAUNCH(GMS_PHASE_COUNT - 1, k_pose_fill_remote<<<blocks_for(h->P, 256), 256, 0, h->stream>>>(t, h->pose[h->cur], h->lo,
                                                                                           h->cnt, h->P));
    h->poses_sharded = false;
    return GMS_OK;
}

// the exchange descriptor of the sharded normalise / resample kernels for the step with sequence number h->xseq
Shard shard_of(const gms_handle* h) {
    Shard sh{};
    sh.nranks = h->cfg.nranks; sh.rank = h->cfg.rank; sh.seq = h->xseq; sh.coarse_cap = h->coarse_cap;
    const int par = (int)(h->xseq & 1);
    for (int q = 0; q < sh.nranks; q++) {
        const bool me = q == h->cfg.rank;
        sh.area[q] = me ? h->xarea[par] : h->peer_xarea[par][q];
        sh.flags[q] = me ? h->xflags4 : h->peer_xflags4[q];
        sh.cdfseg[q] = me ? static_cast<const unsigned long long*>(h->cdf) : h->peer_cdf[q];
        sh.w[q] = me ? h->w[h->cur] : h->peer_w[h->cur][q];
        sh.lw[q] = me ? h->lw[h->cur] : h->peer_lw[h->cur][q];
    }
    return sh;
}

// the part of a local-only resampling that was skipped (see resample_partial)
int complete_resample(gms_handle* h) {
    if (h->blocks_stale) { int rc_ = materialize_poses(h); if (rc_) return rc_; }  // sharded normalise / resample
    if (!h->resample_partial) return GMS_OK;
    h->resample_partial = false;
    h->stats_valid = false;  // Stats.strongest_now may be found among the children selected now
    const int from = h->cur ^ 1, to = h->cur;
    const bool forced = h->force_resample;
    h->force_resample = h->partial_force;  // complete it the way it was started
    int rc = launch_select(h, from, to, h->partial_u01, h->partial_count, 0, h->lo);
    if (!rc) rc = launch_select(h, from, to, h->partial_u01, h->partial_count, h->lo + h->cnt, h->P - h->lo - h->cnt);
    h->force_resample = forced;
    return rc;
}

// Per-map operator results (explicit fields left by gms_map_compute_likelihood / the modifying gms_map_* calls) after
// a single-rank resampling: the child in new slot ns inherits the override of its parent's old slot.  Slow path
// (three small D2H copies + a synchronisation); only the per-map operator entry points create overrides.
int remap_field_overrides(gms_handle* h, int old_slots) {
    std::vector<int> parents((size_t)h->P), so((size_t)h->P), sn((size_t)h->P);
    CK(cudaMemcpyAsync(parents.data(), h->parents, (size_t)h->P * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(so.data(), h->slot[old_slots], (size_t)h->P * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sn.data(), h->slot[h->slot_cur], (size_t)h->P * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::unordered_map<int, gms_handle::FieldOverride> next;
    for (int m = 0; m < h->P; m++) {
        auto it = h->field_ovr.find(so[parents[m]]);
        if (it == h->field_ovr.end()) continue;
        gms_handle::FieldOverride o{it->second.fresh, nullptr};
        if (it->second.snap) {
            CK(cudaMalloc((void**)&o.snap, h->cells * 8));
            CK(cudaMemcpyAsync(o.snap, it->second.snap, h->cells * 8, cudaMemcpyDeviceToDevice, h->stream));
        }
        next[sn[m]] = o;
    }
    CK(cudaStreamSynchronize(h->stream));
    clear_field_overrides(h);
    h->field_ovr.swap(next);
    return GMS_OK;
}

int launch_resample(gms_handle* h, double u01, bool local_only = false) {
    const int P = h->P;
    bool parents_listed = false;  // k_resample_small has already listed the selected parents
    h->stats_valid = false;  // Stats.strongest_now changes
    {
        Phase ph(h, GMS_PHASE_RESAMPLE);
        const int nxt = h->cur ^ 1;
        const int m_begin = local_only ? h->lo : 0, m_count = local_only ? h->cnt : P;
        if (h->sharded_post && h->direct && h->tile_fx_valid && h->tile_fx_sharded && h->resample_mode == GMS_RESAMPLE_FIXED) {
            // every rank builds its own CDF segment and selects its own children (k_resample_shard)
            SelectArgs a = select_args(h, h->cur, nxt, u01, h->resample_count, h->lo, h->cnt);
            a.wp_counter = h->wp_counter;
            const unsigned long long* fx = h->np.fx;
            int ntiles = (h->cnt + 1023) / 1024, lo = h->lo, cnt = h->cnt;
            Shard sh = shard_of(h);
            const unsigned grid = (unsigned)std::max(1, std::min(h->num_sms * kNormCtasPerSm, std::max(ntiles, (cnt + kNormThreads - 1) / kNormThreads)));
            LAUNCH_COOP(GMS_PHASE_RESAMPLE, k_resample_shard, grid, kNormThreads, 0, &a, &fx, &ntiles, &lo, &cnt, &sh);
            h->resample_partial = false;
            h->poses_sharded = false;
            h->blocks_stale = true;
            h->wpose_valid = false;
            h->cur = nxt;
            h->tile_fx_valid = false;
            h->resample_count++;
            return GMS_OK;
        }
        if (h->blocks_stale) { int rc_ = materialize_poses(h); if (rc_) return rc_; }  // the replicated CDF needs every rank's weights
        if (h->resample_mode == GMS_RESAMPLE_FIXED) {
            if (!h->tile_fx_valid)  // tile sums of trunc(w * 2^60): by-product of k_norm_coop / k_neff
                LAUNCH(GMS_PHASE_RESAMPLE, k_neff<<<h->ntiles, 1024, 0, h->stream>>>(h->w[h->cur], P, h->ntiles, h->np, h->st));
            SelectArgs a = select_args(h, h->cur, nxt, u01, h->resample_count, m_begin, m_count);
            const bool fold = !local_only && h->fold_wpose;  // the whole new generation passes through this launch
            if (fold) {
                a.wp_part = h->wp_part;
                a.wp_counter = h->wp_counter;
            }
            h->wpose_valid = fold;
            const unsigned long long* fx = h->np.fx;
            int ntiles = h->ntiles;
            const unsigned grid = (unsigned)std::max(1, std::min(h->num_sms * kNormCtasPerSm, std::max(ntiles, (m_count + kNormThreads - 1) / kNormThreads)));
            LAUNCH_COOP(GMS_PHASE_RESAMPLE, k_resample_coop, grid, kNormThreads, 0, &a, &fx, &ntiles);
        } else if (h->small_fused && h->cfg.nranks == 1 && P <= kCdfChunk) {
            // CDF + selection (+ the list of selected parents, when a scan waits for it) in one single-CTA launch
            h->wpose_valid = false;
            const bool list = h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->integ_pending;
            if (list && !h->used) CK(cudaMalloc((void**)&h->used, (size_t)(h->cnt + 1) * 4));  // [0]: length, [1..]: the list
            const SelectArgs a = select_args(h, h->cur, nxt, u01, h->resample_count, 0, P);
            LAUNCH(GMS_PHASE_RESAMPLE, k_resample_small<<<1, 1024, 0, h->stream>>>(a, h->lo, h->cnt, list ? h->used + 1 : nullptr,
                                                                                   list ? h->used : nullptr));
            parents_listed = list;
        } else {
            h->wpose_valid = false;
            LAUNCH(GMS_PHASE_RESAMPLE, k_cdf_literal<<<1, 256, 0, h->stream>>>(h->w[h->cur], P, (double*)h->cdf, h->st,
                                                                              h->force_resample ? 1 : 0));
            int rc = launch_select(h, h->cur, nxt, u01, h->resample_count, m_begin, m_count);
            if (rc) return rc;
        }
        h->resample_partial = local_only;
        h->poses_sharded = false;  // the selection gathers every child it selects (the rest follows in complete_resample)
        h->partial_u01 = u01;
        h->partial_count = h->resample_count;
        h->partial_force = h->force_resample;
        h->cur = nxt;
        h->tile_fx_valid = false;
    }
    h->resample_count++;
    if (h->cfg.map_mode != GMS_MAP_PER_PARTICLE) return GMS_OK;
    if (h->cfg.nranks > 1 && !h->field_ovr.empty())
        return fail(h, GMS_ERR_STATE, "per-particle maps across ranks: per-map operator results (gms_map_*) cannot follow "
                                      "their particles through a resampling; call an update first");
    const int old_slots = h->slot_cur;
    if (h->cfg.nranks > 1 && !h->peers_ready)
        return fail(h, GMS_ERR_STATE, "per-particle maps across ranks: call gms_ipc_import before resampling");
    if (h->integ_pending) {  // the scan goes into the maps of the selected parents only: the others are dropped
        if (!parents_listed) {
            if (!h->used) CK(cudaMalloc((void**)&h->used, (size_t)(h->cnt + 1) * 4));  // [0]: length, [1..]: the list
            CK(cudaMemsetAsync(h->used, 0, 4, h->stream));
            LAUNCH(GMS_PHASE_MAP_UPDATE, k_list_parents<<<blocks_for(P, 256), 256, 0, h->stream>>>(h->parents, P, h->lo, h->cnt,
                                                                                                h->used + 1, h->used));
        }
        int rc_ = integrate_pending(h, h->used + 1, h->used);
        if (rc_) return rc_;
    }
    if (h->cfg.nranks > 1) {
        {   // every rank's maps are final before any rank pulls one (k_maps_final)
            Phase ph(h, GMS_PHASE_EXCHANGE);
            Shard sh{};
            sh.nranks = h->cfg.nranks; sh.rank = h->cfg.rank; sh.seq = ++h->mapseq;
            for (int q = 0; q < sh.nranks; q++) sh.flags[q] = q == h->cfg.rank ? h->xflags4 : h->peer_xflags4[q];
            LAUNCH(GMS_PHASE_EXCHANGE, k_maps_final<<<1, 32, 0, h->stream>>>(sh, h->st));
        }
        Phase ph(h, GMS_PHASE_MAP_COPY);
        const int nxt = h->slot_cur ^ 1;
        LAUNCH(GMS_PHASE_MAP_COPY, k_assign_slots_mr<<<h->cfg.nranks, 1024, 0, h->stream>>>(
                                       h->parents, P, h->cnt, h->cfg.nranks, h->S, h->cfg.rank, h->slot[h->slot_cur],
                                       h->slot[nxt], h->dup_src_rank, h->dup_src, h->dup_dst, h->dup_level, h->scratch2p,
                                       h->st));
        h->slot_cur = nxt;
        const int chunks = std::max(1, std::min(32, h->H / 16));
        for (int level = 0; level < 2; level++) {  // 0: pulls + copies of old-generation maps, 1: copies of pulled replicas
            LAUNCH(GMS_PHASE_MAP_COPY, k_job_rects<<<blocks_for(h->cnt, 256), 256, 0, h->stream>>>(
                                           h->dup_src_rank, h->dup_src, h->dup_dst, h->dup_level, level, h->dup_rect,
                                           h->rect, h->st, h->peers, h->g));
            const bool bulk = h->copy_bulk && ((h->cells | (size_t)h->W) & 1) == 0;
            if (level == 0 || !bulk)  // pulls over NVLink (level 0), and every copy without the TMA path
                LAUNCH(GMS_PHASE_MAP_COPY, k_copy_maps<<<(unsigned)((long long)chunks * h->cnt), 256, 0, h->stream>>>(
                                               h->counts, h->dup_src, h->dup_dst, h->dup_rect, h->st, h->cells, h->W, chunks,
                                               h->dup_src_rank, h->peers, h->dup_level, level, bulk ? h->cfg.rank : -1));
            if (bulk) {  // copies whose source map is on this rank: TMA engine
                const int bchunks = std::max(1, std::min(h->copy_chunks, h->H / 16));
                LAUNCH(GMS_PHASE_MAP_COPY, k_copy_maps_bulk<<<(unsigned)((long long)bchunks * h->cnt), 32, kCpStages * kCpSeg, h->stream>>>(
                                               h->counts, h->dup_src, h->dup_dst, h->dup_rect, h->st, h->cells, h->W, bchunks,
                                               h->dup_src_rank, h->cfg.rank, h->dup_level, level));
            }
        }
    } else {
        Phase ph(h, GMS_PHASE_MAP_COPY);
        const int nxt = h->slot_cur ^ 1;
        LAUNCH(GMS_PHASE_MAP_COPY, k_assign_slots<<<1, 1024, 0, h->stream>>>(
                                       h->parents, P, h->slot[h->slot_cur], h->slot[nxt], h->dup_src, h->dup_dst,
                                       h->dup_rect, h->rect, h->scratch2p, h->st, h->g));
        h->slot_cur = nxt;
        if (h->copy_bulk && ((h->cells | (size_t)h->W) & 1) == 0) {  // TMA engine: rows start 16-byte aligned
            const int chunks = std::max(1, std::min(h->copy_chunks, h->H / 16));
            LAUNCH(GMS_PHASE_MAP_COPY, k_copy_maps_bulk<<<(unsigned)((long long)chunks * P), 32, kCpStages * kCpSeg, h->stream>>>(
                                           h->counts, h->dup_src, h->dup_dst, h->dup_rect, h->st, h->cells, h->W, chunks,
                                           nullptr, 0, nullptr, 0));
        } else {
            const int chunks = std::max(1, std::min(32, h->H / 16));
            LAUNCH(GMS_PHASE_MAP_COPY, k_copy_maps<<<(unsigned)((long long)chunks * P), 256, 0, h->stream>>>(
                                           h->counts, h->dup_src, h->dup_dst, h->dup_rect, h->st, h->cells, h->W, chunks,
                                           nullptr, h->peers, nullptr, 0, -1));
        }
    }
    // the virtual likelihood field follows the particles: a child inherits its parent's
    if (h->use_upd_pose) {  // ... and the pose its parent's last scan was integrated from
        LAUNCH(GMS_PHASE_COUNT - 1, k_gather_pose<<<blocks_for(P, 256), 256, 0, h->stream>>>(h->parents, h->upd_pose[0],
                                                                                              h->upd_pose[1], P));
        std::swap(h->upd_pose[0], h->upd_pose[1]);
    }
    if (!h->field_ovr.empty()) return remap_field_overrides(h, old_slots);
    return GMS_OK;
}

int step_end(gms_handle* h, int policy, double u01) {
    if (!h->pending) return fail(h, GMS_ERR_STATE, "update_end without update_begin");
    const gms_config& c = h->cfg;
    h->stats_valid = false;
    const double* lw_src = h->lw[h->cur];
    const unsigned long long* xflags = nullptr;
    const bool sharded = c.nranks > 1 && h->direct && h->sharded_post;
    if (c.nranks > 1) {
        if (sharded) {  // every rank normalises its own block; three tiny exchange rounds inside the kernel
            h->xseq++;
        } else if (h->direct) {  // pushed by the peers (k_xpush_lw): the normalise kernel waits for every sender's flag
            h->xseq++;
            lw_src = h->xlw[h->xseq & 1];
            xflags = h->xflags;
        } else {  // filled by the caller's all-gather
            LAUNCH(GMS_PHASE_EXCHANGE, k_import_exchange<<<blocks_for(h->P, 256), 256, 0, h->stream>>>(
                                           h->xglobal, h->P, h->lw[h->cur], h->pose[h->cur]));
        }
    }
    {
        Phase ph(h, GMS_PHASE_NORMALISE);
        NormArgs a{};
        a.lw = lw_src; a.lw_store = lw_src == h->lw[h->cur] ? nullptr : h->lw[h->cur];
        a.w = h->w[h->cur]; a.poses = pose_table(h, h->cur); a.P = h->P; a.policy = policy;
        a.lo = sharded ? h->lo : 0; a.cnt = sharded ? h->cnt : h->P; a.ntiles = (a.cnt + 1023) / 1024;
        a.np = h->np; a.st = h->st; a.xflags = xflags; a.nranks = c.nranks; a.seq = h->xseq;
        if (pulling(h)) {  // every block is read from its owner's exchange buffer; the flags are raised in the kernel
            a.lw = nullptr;
            a.lw_store = h->lw[h->cur];
            a.pull.cnt = h->cnt;
            a.pull.myrank = c.rank;
            const int par = (int)(h->xseq & 1);
            for (int q = 0; q < c.nranks; q++) {
                a.pull.src[q] = q == c.rank ? h->xlw[par] : h->peer_xlw[par][q];
                a.pull.flag[q] = q == c.rank ? h->xflags : h->peer_flags[q];
            }
        }
        a.pose_local = (h->fold_wpose && (c.nranks == 1 || !h->direct)) ? h->pose[h->cur] : nullptr;
        a.wp_part = h->wp_part;
        if (sharded) a.sh = shard_of(h);
        h->wpose_valid = a.pose_local != nullptr;
        h->tile_fx_sharded = sharded;
        if (sharded) h->blocks_stale = true;  // w / lw of the other ranks' blocks are not computed here
        if (h->exact_sums) {  // GMS_SHARDED=1: exact, order-free sums (two grid barriers)
            const unsigned grid = (unsigned)std::max(1, std::min(a.ntiles, h->num_sms * kNormCtasPerSm));
            LAUNCH_COOP(GMS_PHASE_NORMALISE, k_norm_coop, grid, kNormThreads, 0, &a);
        } else {              // fixed tiles, one grid barrier
            const unsigned grid = (unsigned)std::max(1, std::min(a.ntiles, h->num_sms * 6));
            LAUNCH_COOP(GMS_PHASE_NORMALISE, k_norm_tiles, grid, kNormThreads, 0, &a);
        }
        h->tile_fx_valid = true;
    }
    const bool skip = std::fabs(h->pend_dtheta) > (M_PI / 180.0) * c.skip_update_deg;
    bool forked = false;
    if (c.map_mode == GMS_MAP_SHARED && !skip) {
        // map integration from the strongest pose: touches the map only, resampling touches the particle
        // arrays only (it gathers into the other buffer), so the two chains run side by side
        cudaStream_t main = h->stream;
        forked = h->overlap;
        if (forked) {
            CK(cudaEventRecord(h->ev_fork_b, main));
            CK(cudaStreamWaitEvent(h->side_b, h->ev_fork_b, 0));
            h->stream = h->side_b;
        }
        int rc = launch_shared_update(h, cur_beams(h), h->pend_B);
        h->stream = main;
        if (rc) return rc;
        if (forked) CK(cudaEventRecord(h->ev_done_b, h->side_b));
    }
    h->step++;
    h->have_update = true;
    h->pending = false;
    int rc = GMS_OK;
    if (policy != GMS_RESAMPLE_NEVER)
        rc = launch_resample(h, u01, c.nranks > 1 && c.map_mode == GMS_MAP_SHARED);
    if (forked) h->b_pending = true;  // joined by the next step's likelihood chain or the next other call
    return rc;
}

int slot_of(gms_handle* h, int particle, int* slot) {
    if (h->cfg.map_mode == GMS_MAP_SHARED) { *slot = 0; return GMS_OK; }
    if (particle < h->lo || particle >= h->lo + h->cnt)
        return fail(h, GMS_ERR_INVALID_ARG, "particle index is not held by this handle");
    CK(cudaMemcpyAsync(slot, h->slot[h->slot_cur] + particle, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

// ---- per-particle maps: the virtual likelihood field, materialised for the entry points that expose it ----------
// `*out` = likelihoodData of the particle in slot s as Java would hold it now (see gms_handle::field_state); the
// pointer (the handle's one-map buffer or an explicit copy) is valid until the next call that materialises a field.
int field_of(gms_handle* h, int particle, int s, const double** out) {
    if (h->cfg.map_mode == GMS_MAP_SHARED) { *out = h->lik; return GMS_OK; }
    const CellCounts* map = h->counts + (size_t)s * h->cells;
    int state = h->integ_pending ? (int)gms_handle::FIELD_FRESH : h->field_state;  // pending: the counters still lack the scan
    auto it = h->field_ovr.find(s);
    if (it != h->field_ovr.end()) {
        if (it->second.snap) { *out = it->second.snap; return GMS_OK; }
        state = gms_handle::FIELD_FRESH;
    }
    *out = h->lik;
    if (state == gms_handle::FIELD_ZERO) {
        CK(cudaMemsetAsync(h->lik, 0, h->cells * 8, h->stream));
        return GMS_OK;
    }
    if (state == gms_handle::FIELD_BEFORE_LAST) {
        // counters before the last scan = counters now - that scan's increments, replayed from the pose it was
        // integrated from (the particle's pose: resampling copies it with the map, motion has not run since)
        if (!h->fld_counts) CK(cudaMalloc((void**)&h->fld_counts, h->cells * sizeof(CellCounts)));
        CK(cudaMemcpyAsync(h->fld_counts, map, h->cells * sizeof(CellCounts), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemsetAsync(h->tmp_slot, 0, sizeof(int), h->stream));
        const BeamSet& b = h->bs[h->field_bset];
        const float4* poses = h->use_upd_pose ? h->upd_pose[0] : h->pose[h->cur];
        LAUNCH(GMS_PHASE_COUNT - 1, k_map_update_red<true><<<blocks_for(h->field_B, 128), 128, 0, h->stream>>>(
                                        poses, particle, 1, b.xy, b.meas, b.hit, h->field_B, h->fld_counts, h->tmp_slot,
                                        nullptr, nullptr, nullptr, h->st, h->g));
        map = h->fld_counts;
    }
    return launch_blur_whole(h, map, h->lik, h->dirty, h->word_off, h->tile_list);
}
// before a per-map operator changes the counters of slot s: likelihoodData must keep what it holds (Java only
// rebuilds it in computeLikelihoodMap), so the virtual field becomes an explicit copy
int pin_field(gms_handle* h, int particle, int s) {
    if (h->cfg.map_mode == GMS_MAP_SHARED) return GMS_OK;
    auto it = h->field_ovr.find(s);
    if (it != h->field_ovr.end() && it->second.snap) return GMS_OK;
    const double* f = nullptr;
    int rc = field_of(h, particle, s, &f);
    if (rc) return rc;
    double* snap = nullptr;
    CK(cudaMalloc((void**)&snap, h->cells * 8));
    CK(cudaMemcpyAsync(snap, f, h->cells * 8, cudaMemcpyDeviceToDevice, h->stream));
    h->field_ovr[s] = gms_handle::FieldOverride{false, snap};
    return GMS_OK;
}

// host scan -> device beam set `b` through the pinned staging ring: one packed block, no stream synchronisation
int upload_beams(gms_handle* h, BeamSet& b, const double* xy, const double* dist, const uint8_t* hit, int B,
                 const double* normals = nullptr) {
    int rc = ensure_beams(h, B);
    if (rc) return rc;
    if (B == 0 && !normals) return GMS_OK;
    const size_t off_n = (((size_t)B * 25 + 63) / 64) * 64;
    unsigned char* s = nullptr;
    int slot = 0;
    if ((rc = stage_acquire(h, off_n + (normals ? (size_t)h->cnt * 16 : 0) + 64, &s, &slot))) return rc;
    if (B > 0) {
        std::memcpy(s, xy, (size_t)B * 16);
        std::memcpy(s + (size_t)B * 16, dist ? (const void*)dist : (const void*)xy, (size_t)B * 8);
        std::memcpy(s + (size_t)B * 24, hit, (size_t)B);
        b.view(B);  // [xy | dist | hit] packed for B beams: one copy
        CK(cudaMemcpyAsync(b.blob, s, (size_t)B * 25, cudaMemcpyHostToDevice, h->stream));
    }
    if (normals) {
        std::memcpy(s + off_n, normals, (size_t)h->cnt * 16);
        CK(cudaMemcpyAsync(h->d_normals, s + off_n, (size_t)h->cnt * 16, cudaMemcpyHostToDevice, h->stream));
    }
    return stage_release(h, slot);
}

int do_reset(gms_handle* h) {
    const int P = h->P;
    for (int i = 0; i < 2; i++)
        LAUNCH(GMS_PHASE_COUNT - 1, k_init_particles<<<blocks_for(P, 256), 256, 0, h->stream>>>(h->pose[i], h->w[i],
                                                                                               h->lw[i], h->parents, P));
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE)
        for (int i = 0; i < 2; i++)
            LAUNCH(GMS_PHASE_COUNT - 1, k_iota_mod<<<blocks_for(h->P, 256), 256, 0, h->stream>>>(h->slot[i], h->P, h->cnt));
    // GridMap.createMapData(null) GridMap.java:106-117: logData = logOdds(0.5) = 0.0 (no counts),
    // likelihoodData = 0.0 until the first computeLikelihoodMap; the whole map is dirty.
    CK(cudaMemsetAsync(h->counts, 0, (size_t)h->S * h->cells * sizeof(CellCounts), h->stream));
    CK(cudaMemsetAsync(h->lik, 0, h->cells * sizeof(double), h->stream));  // the shared map's field (per-particle maps:
                                                                          // scratch; their field is virtual, FIELD_ZERO)
    // nothing explored yet; every tile needs its first likelihood build
    LAUNCH(GMS_PHASE_COUNT - 1, k_fill_rect<<<blocks_for(h->S, 256), 256, 0, h->stream>>>(
                                    h->rect, h->S, make_int4(0x7fffffff, 0x7fffffff, -1, -1)));
    LAUNCH(GMS_PHASE_COUNT - 1, k_fill_dirty<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                    h->dirty, 1, h->g.tile_words, h->tiles_per_map));
    clear_field_overrides(h);
    h->field_state = gms_handle::FIELD_ZERO;
    h->use_upd_pose = false;
    h->integ_pending = false;
    if (h->dirty_alt) CK(cudaMemsetAsync(h->dirty_alt, 0, (size_t)h->g.tile_words * 4, h->stream));
    h->alt_needs_clear = false;
    CK(cudaMemsetAsync(h->st, 0, sizeof(Stats), h->stream));
    h->cur = 0; h->slot_cur = 0;
    h->step = 0; h->resample_count = 0;
    h->have_update = false; h->pending = false; h->stats_valid = false; h->tile_fx_valid = false;
    h->resample_partial = false; h->wpose_valid = false;
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
EXPORT int gms_config_default(gms_config* cfg) {
    if (!cfg) return GMS_ERR_INVALID_ARG;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = (uint32_t)sizeof *cfg;
    cfg->num_particles = 500;        // SLAM.java:50
    cfg->map_width_m = 6.0f;         // SLAM.java:57
    cfg->map_height_m = 6.0f;
    cfg->resolution = 0.05f;
    cfg->origin_x = -3.0f;
    cfg->origin_y = -3.0f;
    cfg->sensor_max_range = 10.0f;   // SensorModel.java:20
    cfg->z_hit = 0.9;                // GridMap.java:259
    cfg->hit_tolerance = 2.0f;       // GridMap.java:223
    cfg->extra_steps = 2;            // GridMap.java:210
    cfg->p_free = 0.30f;             // SensorModel.java:23-24
    cfg->p_occ = 0.9f;
    cfg->noise_center_base = 0.01;   // Odometry.java:63-64
    cfg->noise_center_gain = 0.05;
    cfg->noise_theta_base_deg = 5;
    cfg->noise_theta_gain = 0.1;
    cfg->skip_update_deg = 30;       // SLAM.java:82
    cfg->likelihood_sigma_num = 0.05;  // GridMap.java:94
    cfg->map_mode = GMS_MAP_PER_PARTICLE;
    cfg->resample_mode = GMS_RESAMPLE_AUTO;
    cfg->device = 0;
    cfg->rank = 0;
    cfg->nranks = 1;
    cfg->seed = 0x5EEDull;
    return GMS_OK;
}

EXPORT int gms_create(const gms_config* cfg, gms_handle** out) {
    gms_handle* h = nullptr;  // for the CK / fail macros before the handle exists
    if (!cfg || !out) return fail(nullptr, GMS_ERR_INVALID_ARG, "gms_create: NULL argument");
    if (cfg->struct_size != sizeof(gms_config))
        return fail(nullptr, GMS_ERR_INVALID_ARG, "gms_create: gms_config.struct_size mismatch");
    if (cfg->nranks > kMaxRanks) return fail(nullptr, GMS_ERR_INVALID_ARG, "gms_create: at most 16 ranks");
    if (cfg->num_particles < 1 || !(cfg->resolution > 0) || !(cfg->map_width_m > 0) || !(cfg->map_height_m > 0) ||
        cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks || cfg->extra_steps < 0 ||
        (cfg->map_mode != GMS_MAP_PER_PARTICLE && cfg->map_mode != GMS_MAP_SHARED) || cfg->resample_mode < 0 ||
        cfg->resample_mode > 2 || cfg->num_particles % cfg->nranks != 0 || cfg->update_mode < 0 || cfg->update_mode > 1)
        return fail(nullptr, GMS_ERR_INVALID_ARG, "gms_create: invalid configuration");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1 || cfg->device < 0 || cfg->device >= ndev) {
        (void)cudaGetLastError();
        return fail(nullptr, GMS_ERR_CUDA,
                    std::string("gms_create: no usable CUDA device (libgms has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range"));
    }
    h = new (std::nothrow) gms_handle();
    if (!h) return fail(nullptr, GMS_ERR_OOM, "gms_create: out of host memory");
    h->cfg = *cfg;
    h->dev = cfg->device;
    // GridMap ctor GridMap.java:80-100
    h->W = java_d2i_host(std::ceil((double)(cfg->map_width_m / cfg->resolution)));
    h->H = java_d2i_host(std::ceil((double)(cfg->map_height_m / cfg->resolution)));
    h->world_w = (float)h->W * cfg->resolution;
    h->world_h = (float)h->H * cfg->resolution;
    const double sigma = std::sqrt(cfg->likelihood_sigma_num / (double)cfg->resolution);
    const int half = java_d2i_host(std::ceil(sigma * 3));
    if (2 * half + 1 > 31 || h->W < 1 || h->H < 1 || (long long)h->W * h->H > (1LL << 31) - 1) {
        delete h;
        return fail(nullptr, GMS_ERR_INVALID_ARG, "gms_create: kernel too wide or grid size out of range");
    }
    // the shared-map ray kernels pack a cell as x | y << 16, the sorted update keys a cell as index << 2
    if (cfg->map_mode == GMS_MAP_SHARED &&
        (h->W > 65535 || h->H > 65535 ||
         (cfg->update_mode == GMS_UPDATE_SORTED && (long long)h->W * h->H >= (1LL << 30)))) {
        delete h;
        return fail(nullptr, GMS_ERR_INVALID_ARG,
                    "gms_create: a shared map is limited to 65535 cells per side (2^30 cells with GMS_UPDATE_SORTED)");
    }
    Geometry& g = h->g;
    g.W = h->W; g.H = h->H;
    g.extra_steps = cfg->extra_steps;
    g.khalf = half; g.ktaps = 2 * half + 1;
    g.res_f = cfg->resolution;
    g.tol_half = cfg->hit_tolerance / 2;
    g.res = (double)cfg->resolution; g.posx = (double)cfg->origin_x; g.posy = (double)cfg->origin_y;
    g.inv_res = 1.0 / g.res;
    {
        int bits = 1;
        while ((1 << bits) <= std::max(h->W, h->H)) bits++;  // max(W, H) < 2^bits
        g.fx_k = std::min(20, 30 - bits);
        g.fx_magic = std::ldexp(1.5, 52 - g.fx_k);
        unsigned long long u;
        std::memcpy(&u, &g.fx_magic, 8);
        g.fx_hi = (int)(u >> 32);
        g.fx_margin = std::max(8, 1 << std::max(0, g.fx_k - 14));
        g.fac_lp = bits;  // 2^bits > max(W, H) - 1 ... the padded side of the factor field
        while (g.fac_lp > 0 && (1 << (g.fac_lp - 1)) >= std::max(h->W, h->H)) g.fac_lp--;
        g.fac_pitch = 1 << g.fac_lp;
    }
    g.tiles_x = (h->W + kTileW - 1) / kTileW;
    g.tiles_y = (h->H + kTileH - 1) / kTileH;
    g.tile_words = (g.tiles_x * g.tiles_y + 31) / 32;
    h->tiles_per_map = g.tiles_x * g.tiles_y;
    g.z_hit = cfg->z_hit;
    g.uniform_term = 1.0 / (double)cfg->sensor_max_range;                          // GridMap.java:286
    g.random_term = (1 - cfg->z_hit) * 1.0 / (double)cfg->sensor_max_range;        // GridMap.java:288
    g.l_free = log_odds((double)cfg->p_free);
    g.l_occ = log_odds((double)cfg->p_occ);
    gaussian_kernel(sigma, half, g.kernel);
    h->P = cfg->num_particles;
    h->cnt = h->P / cfg->nranks;
    h->lo = cfg->rank * h->cnt;
    // per-particle maps across ranks keep cnt spare slots so that a resampling exchange never writes a slot
    // of the old generation (other ranks may still be pulling from it)
    h->S = cfg->map_mode == GMS_MAP_SHARED ? 1 : (cfg->nranks > 1 ? 2 * h->cnt : h->cnt);
    h->cells = (size_t)h->W * h->H;
    h->resample_mode = cfg->resample_mode == GMS_RESAMPLE_AUTO
                           ? (h->P <= 2048 ? GMS_RESAMPLE_LITERAL : GMS_RESAMPLE_FIXED)
                           : cfg->resample_mode;
    auto bail = [&](int rc) {
        std::string m = h->err;
        free_all(h);
        g_create_err = m;
        return rc;
    };
#define CKC(call)                                                                \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) return bail(cuda_fail(h, e_, #call));             \
    } while (0)
    CKC(cudaSetDevice(h->dev));
    CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CKC(cudaStreamCreateWithFlags(&h->side_a, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->side_b, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&h->ev_fork_a, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_done_a, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_fork_b, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_done_b, cudaEventDisableTiming));
    if (const char* e = std::getenv("GMS_NO_OVERLAP")) h->overlap = std::atoi(e) == 0;
    if (const char* e = std::getenv("GMS_NO_COOP")) h->coop = std::atoi(e) == 0;
    {
        int coop_ok = 0;
        CKC(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, h->dev));
        if (!coop_ok) return bail(fail(h, GMS_ERR_CUDA, "gms_create: the device does not support cooperative launches"));
    }
    const size_t P = (size_t)h->P;
    for (int i = 0; i < 2; i++) {
        CKC(cudaMalloc((void**)&h->pose[i], P * sizeof(float4)));
        CKC(cudaMalloc((void**)&h->w[i], P * 8));
        CKC(cudaMalloc((void**)&h->lw[i], P * 8));
        CKC(cudaMalloc((void**)&h->slot[i], P * 4));  // global table: particle m -> slot in its rank's arena
    }
    CKC(cudaMalloc((void**)&h->parents, P * 4));
    CKC(cudaMalloc((void**)&h->cdf, P * 8));
    CKC(cudaMalloc((void**)&h->counts, (size_t)h->S * h->cells * sizeof(CellCounts)));
    // ONE field: the shared map's likelihoodData; with per-particle maps the field is virtual (k_score_pp) and this
    // is the buffer a getter materialises one particle's field into
    CKC(cudaMalloc((void**)&h->lik, h->cells * sizeof(double)));
    if (cfg->map_mode == GMS_MAP_SHARED) {  // + one sentinel element holding 1.0 for out-of-map end points
        // padded square (side 2^fac_lp) + one sentinel element, all 1.0 until the first refresh writes the map's cells
        const size_t nfac = (size_t)g.fac_pitch * g.fac_pitch + 1;
        CKC(cudaMalloc((void**)&h->fac, nfac * sizeof(double)));
        k_fill_f64<<<(unsigned)std::min<size_t>((nfac + 255) / 256, 148 * 16), 256, 0, h->stream>>>(h->fac, nfac, 1.0);
        CKC(cudaGetLastError());
    }
    CKC(cudaMalloc((void**)&h->rect, (size_t)h->S * sizeof(int4)));
    CKC(cudaMalloc((void**)&h->dirty, (size_t)g.tile_words * 4));  // dirty tiles of the shared map / scratch work list
    if (cfg->map_mode == GMS_MAP_SHARED && g.tile_words <= kSelfListWords &&
        !(std::getenv("GMS_SELF_LIST") && std::atoi(std::getenv("GMS_SELF_LIST")) == 0)) {
        CKC(cudaMalloc((void**)&h->dirty_alt, (size_t)g.tile_words * 4));
        h->self_list = true;
    }
    CKC(cudaMalloc((void**)&h->word_off, (size_t)g.tile_words * 4));
    CKC(cudaMalloc((void**)&h->tile_list, (size_t)h->tiles_per_map * sizeof(int2)));
    CKC(cudaMalloc((void**)&h->dup_rect, P * sizeof(int4)));
    CKC(cudaMalloc((void**)&h->dup_src, P * 4));
    CKC(cudaMalloc((void**)&h->dup_dst, P * 4));
    CKC(cudaMalloc((void**)&h->dup_src_rank, P * 4));
    CKC(cudaMalloc((void**)&h->dup_level, P * 4));
    CKC(cudaMalloc((void**)&h->scratch2p, std::max(2 * P, (size_t)2 * cfg->nranks * h->S) * 4));
    CKC(cudaMalloc((void**)&h->d_normals, (size_t)h->cnt * 16));
    CKC(cudaMalloc((void**)&h->xlocal, (size_t)h->cnt * sizeof(ExchangeRec)));
    CKC(cudaMalloc((void**)&h->xglobal, P * sizeof(ExchangeRec)));
    if (cfg->nranks > 1) {
        CKC(cudaMalloc((void**)&h->xlw[0], P * 8));
        CKC(cudaMalloc((void**)&h->xlw[1], P * 8));
        CKC(cudaMalloc((void**)&h->xflags, kMaxRanks * 8));
        CKC(cudaMemset(h->xflags, 0, kMaxRanks * 8));
        CKC(cudaMalloc((void**)&h->xticket, 4));
        CKC(cudaMemset(h->xticket, 0, 4));
        h->coarse_cap = (h->cnt + kCoarseStep - 1) / kCoarseStep + 1;
        h->xarea_bytes = sizeof(XArea) + (size_t)kMaxRanks * h->coarse_cap * 8;
        for (int i = 0; i < 2; i++) {
            CKC(cudaMalloc((void**)&h->xarea[i], h->xarea_bytes));
            CKC(cudaMemset(h->xarea[i], 0, h->xarea_bytes));
        }
        CKC(cudaMalloc((void**)&h->xflags4, 4 * kMaxRanks * 8));
        CKC(cudaMemset(h->xflags4, 0, 4 * kMaxRanks * 8));
    }
    h->d_tmp_bytes = std::max(h->cells * 8, P * 24);
    CKC(cudaMalloc(&h->d_tmp, h->d_tmp_bytes));
    CKC(cudaMalloc((void**)&h->tmp_pose, sizeof(float4)));
    CKC(cudaMalloc((void**)&h->tmp_slot, sizeof(int)));
    CKC(cudaMalloc((void**)&h->tmp_lw, sizeof(double)));
    CKC(cudaMalloc((void**)&h->st, sizeof(Stats)));
    CKC(cudaMalloc((void**)&h->sort.hist, (size_t)kSortBins * 4));
    CKC(cudaMalloc((void**)&h->sort.chunk_total, (size_t)kSortChunks * 4));
    CKC(cudaMalloc((void**)&h->sort.offs, kSortBins * 4));
    CKC(cudaMemset(h->sort.hist, 0, (size_t)kSortBins * 4));
    CKC(cudaMalloc((void**)&h->sort.key, (size_t)h->cnt * 4));
    CKC(cudaMalloc((void**)&h->sort.rank, (size_t)h->cnt * 4));
    CKC(cudaMalloc((void**)&h->sort.order, (size_t)h->cnt * 4));
    h->ntiles = (h->P + 1023) / 1024;
    if (const char* e = std::getenv("GMS_SHARDED")) h->exact_sums = std::atoi(e) != 0;
    if (const char* e = std::getenv("GMS_SCORE_DYNAMIC")) h->score_dynamic = std::atoi(e) != 0;
    if (const char* e = std::getenv("GMS_PULL")) h->xpull = std::atoi(e) != 0;
    if (const char* e = std::getenv("GMS_SMALL_FUSED")) h->small_fused = std::atoi(e) != 0;
    // by the TOTAL particle count (the same on every rank): 100 particles x 360 beams 0.043 -> 0.033 ms with 8 warps,
    // 1000 particles 0.061 -> 0.073 ms
    h->pp_warps = h->P <= 256 ? 8 : 4;
    if (const char* e = std::getenv("GMS_PP_WARPS")) h->pp_warps = std::atoi(e) == 8 ? 8 : 4;
    if (const char* e = std::getenv("GMS_COPY_BULK")) h->copy_bulk = std::atoi(e) != 0;
    if (const char* e = std::getenv("GMS_COPY_CHUNKS")) h->copy_chunks = std::max(1, std::min(64, std::atoi(e)));
    CKC(cudaFuncSetAttribute(k_copy_maps_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, kCpStages * kCpSeg));
    CKC(cudaMalloc((void**)&h->score_work, 4));
    if (const char* e = std::getenv("GMS_DEFER_INTEGRATION")) h->defer_integration = std::atoi(e) != 0;
    if (const char* e = std::getenv("GMS_SCORE_V")) h->score_v = std::max(0, std::min(7, std::atoi(e)));
    if (const char* e = std::getenv("GMS_SCORE_G")) { const int v = std::atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) h->score_g = v; }
    { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, h->dev) == cudaSuccess) h->num_sms = prop.multiProcessorCount; }
    {
        const size_t nt = (size_t)h->ntiles;
        const size_t np_cap = std::max<size_t>(nt, 1024);  // (m, idx): one entry per CTA of k_norm_coop; s: per tile
        CKC(cudaMalloc((void**)&h->np.m, np_cap * 8));
        CKC(cudaMalloc((void**)&h->np.idx, np_cap * 4));
        CKC(cudaMalloc((void**)&h->np.s, np_cap * 8));
        CKC(cudaMalloc((void**)&h->np.ws, nt * 8));
        CKC(cudaMalloc((void**)&h->np.q, nt * 8));
        CKC(cudaMalloc((void**)&h->np.fx, nt * 8));
        CKC(cudaMalloc((void**)&h->np.x128, nt * 48));
        CKC(cudaMalloc((void**)&h->np.counter, 4));
        CKC(cudaMemset(h->np.counter, 0, 4));
        CKC(cudaMalloc((void**)&h->wp_part, (nt * (1024 / kNormThreads) + 4) * 32));  // per chunk of kNormThreads children
        CKC(cudaMalloc((void**)&h->wp_counter, 4));
        CKC(cudaMemset(h->wp_counter, 0, 4));
        CKC(cudaMalloc((void**)&h->ray_maxlen, 4));
        CKC(cudaMemset(h->ray_maxlen, 0, 4));
    }
    CKC(cudaMallocHost((void**)&h->h_st, sizeof(Stats)));
    CKC(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKC(cudaFuncSetAttribute(k_likelihood<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CKC(cudaFuncSetAttribute(k_likelihood<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CKC(cudaFuncSetAttribute(k_likelihood_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    if (cfg->map_mode == GMS_MAP_SHARED && g.khalf == 3 && (h->W & 1) == 0 && !(std::getenv("GMS_LIK_TMA") && std::atoi(std::getenv("GMS_LIK_TMA")) == 0)) {
        // TMA descriptor of counts[S][H][W] (8-byte cells); box = one tile + halo.  Any failure keeps the plain kernel.
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t dims[3] = {(cuuint64_t)h->W, (cuuint64_t)h->H, (cuuint64_t)h->S};
            const cuuint64_t strides[2] = {(cuuint64_t)h->W * 8, (cuuint64_t)h->cells * 8};
            const cuuint32_t box[3] = {(cuuint32_t)kTmaTileW, (cuuint32_t)kTmaTileH, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = ((EncodeFn)fn)(&h->lik_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, h->counts, dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            h->lik_tma = r == CUDA_SUCCESS;
        }
        (void)cudaGetLastError();
    }
    int rc = ensure_beams(h, 1024);
    if (rc) return bail(rc);
    rc = do_reset(h);
    if (rc) return bail(rc);
    *out = h;
    return GMS_OK;
}

EXPORT int gms_destroy(gms_handle* h) {
    free_all(h);
    return GMS_OK;
}

EXPORT const char* gms_last_error(const gms_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

EXPORT int gms_get_info(const gms_handle* h, gms_info* info) {
    if (!h || !info) return GMS_ERR_INVALID_ARG;
    std::memset(info, 0, sizeof *info);
    info->abi_version = GMS_ABI_VERSION;
    info->is_cuda = 1;
    info->grid_w = h->W; info->grid_h = h->H;
    info->num_particles = h->P; info->local_begin = h->lo; info->local_count = h->cnt;
    info->num_slots = h->S; info->kernel_taps = h->g.ktaps; info->resample_mode = h->resample_mode;
    std::memcpy(info->kernel, h->g.kernel, sizeof(double) * h->g.ktaps);
    info->l_free = h->g.l_free; info->l_occ = h->g.l_occ;
    info->world_w = h->world_w; info->world_h = h->world_h;
    return GMS_OK;
}

EXPORT int gms_reset(gms_handle* h) {
    ENTER_KEEP(h);
    return do_reset(h);
}

EXPORT int gms_update(gms_handle* h, const double* beam_xy, const double* beam_dist, const uint8_t* beam_hit,
                      int32_t B, double d_center, double d_theta, const double* normals, double* neff_out) {
    ENTER_STEP(h);
    if (int rc_ = check_beams(h, B, beam_xy, beam_dist, beam_hit, "gms_update")) return rc_;
    if (h->cfg.nranks != 1) return fail(h, GMS_ERR_STATE, "gms_update: multi-rank handles use update_begin/end");
    flip_beams(h);
    h->fold_wpose = true;
    BeamSet& bs = cur_beams(h);
    int rc = upload_beams(h, bs, beam_xy, beam_dist, beam_hit, B, normals);
    if (rc) return rc;
    rc = step_begin(h, (const double*)bs.xy, bs.dist, bs.hit, B, d_center, d_theta, normals ? h->d_normals : nullptr);
    if (rc) return rc;
    rc = step_end(h, GMS_RESAMPLE_NEVER, 0.0);
    if (rc) return rc;
    rc = fetch_stats(h);
    if (rc) return rc;
    if (neff_out) *neff_out = h->h_st->neff;
    return GMS_OK;
}

EXPORT int gms_resample(gms_handle* h, double u01) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (u01 >= 1.0) return fail(h, GMS_ERR_INVALID_ARG, "gms_resample: u01 must be < 1");
    h->fold_wpose = true;
    // SLAM.resample() returns nothing: the work is only enqueued; every getter synchronises before it reads.
    // The kernels resample whatever the last step's policy decided on the device (SelectArgs::force).
    h->force_resample = true;
    const int rc = launch_resample(h, u01);
    h->force_resample = false;
    return rc;
}

EXPORT int gms_calculate_neff(gms_handle* h, double* neff_out) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!neff_out) return GMS_ERR_INVALID_ARG;
    LAUNCH(GMS_PHASE_NORMALISE, k_neff<<<h->ntiles, 1024, 0, h->stream>>>(h->w[h->cur], h->P, h->ntiles, h->np, h->st));
    h->tile_fx_valid = true;
    h->stats_valid = false;
    int rc = fetch_stats(h);
    if (rc) return rc;
    *neff_out = h->h_st->neff_query;
    return GMS_OK;
}

EXPORT int gms_get_weighted_pose(gms_handle* h, float pose[3]) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!pose) return GMS_ERR_INVALID_ARG;
    if (!h->wpose_valid) {  // otherwise the last normalise / resample launch already produced it
        { int rc_ = materialize_poses(h); if (rc_) return rc_; }
        LAUNCH(GMS_PHASE_NORMALISE, k_weighted_pose<<<h->ntiles, 1024, 0, h->stream>>>(
                                        h->w[h->cur], h->pose[h->cur], h->P, h->ntiles, h->wp_part, h->wp_counter, h->st));
        h->stats_valid = false;
        h->wpose_valid = true;
    }
    int rc = fetch_stats(h);
    if (rc) return rc;
    std::memcpy(pose, h->h_st->weighted_pose, 3 * sizeof(float));
    return GMS_OK;
}

EXPORT int gms_get_strongest(gms_handle* h, int32_t* index, float pose[3], double* weight) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }  // the first child may belong to another rank's block
    int rc = fetch_stats(h);
    if (rc) return rc;
    if (!h->have_update) {  // SLAM.reset: strongestParticle = particles.get(0) (SLAM.java:75)
        if (index) *index = -1;
        if (pose) pose[0] = pose[1] = pose[2] = 0.f;
        if (weight) *weight = 1.0 / h->P;
        return GMS_OK;
    }
    // Java keeps referencing the Particle object that was strongest at the last update; after a resampling that
    // object lives on as its first child (its map slot is inherited), so that is the index reported here
    if (index) *index = h->h_st->strongest_now;
    if (pose) std::memcpy(pose, h->h_st->strongest_pose, 3 * sizeof(float));
    if (weight) *weight = h->h_st->strongest_w;
    return GMS_OK;
}

EXPORT int gms_get_poses(gms_handle* h, float* xyt) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!xyt) return GMS_ERR_INVALID_ARG;
    { int rc_ = materialize_poses(h); if (rc_) return rc_; }
    LAUNCH(GMS_PHASE_COUNT - 1,
           k_pose_unpack<<<blocks_for(h->P, 256), 256, 0, h->stream>>>(h->pose[h->cur], (float*)h->d_tmp, h->P));
    CK(cudaMemcpyAsync(xyt, h->d_tmp, (size_t)h->P * 12, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

static int copy_out(gms_handle* h, void* dst, const void* src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}
EXPORT int gms_get_weights(gms_handle* h, double* w) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!w) return GMS_ERR_INVALID_ARG;
    return copy_out(h, w, h->w[h->cur], (size_t)h->P * 8);
}
EXPORT int gms_get_log_weights(gms_handle* h, double* lw) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!lw) return GMS_ERR_INVALID_ARG;
    return copy_out(h, lw, h->lw[h->cur], (size_t)h->P * 8);
}
EXPORT int gms_get_parents(gms_handle* h, int32_t* parents) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!parents) return GMS_ERR_INVALID_ARG;
    return copy_out(h, parents, h->parents, (size_t)h->P * 4);
}

EXPORT int gms_get_map(gms_handle* h, int32_t particle, int32_t kind, void* dst, size_t bytes) {
    ENTER(h);
    if (!dst) return GMS_ERR_INVALID_ARG;
    if (kind < GMS_MAP_LOG || kind > GMS_MAP_OCC_COUNT) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: unknown kind");
    const size_t esz = (kind == GMS_MAP_LOG || kind == GMS_MAP_LIKELIHOOD) ? 8 : 4;
    if (bytes != h->cells * esz) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: size mismatch");
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    const CellCounts* c = h->counts + (size_t)s * h->cells;
    const unsigned nb = blocks_for((long long)h->cells, 256);
    if (kind == GMS_MAP_LIKELIHOOD) {
        const double* f = nullptr;
        if ((rc = field_of(h, particle, s, &f))) return rc;
        return copy_out(h, dst, f, bytes);
    }
    if (kind == GMS_MAP_LOG)
        LAUNCH(GMS_PHASE_COUNT - 1,
               k_counts_to_log<<<nb, 256, 0, h->stream>>>(c, (double*)h->d_tmp, h->cells, h->g.l_free, h->g.l_occ));
    else
        LAUNCH(GMS_PHASE_COUNT - 1, k_counts_split<<<nb, 256, 0, h->stream>>>(c, (uint32_t*)h->d_tmp, h->cells,
                                                                             kind == GMS_MAP_OCC_COUNT));
    return copy_out(h, dst, h->d_tmp, bytes);
}

EXPORT int gms_set_poses(gms_handle* h, const float* xyt) {
    ENTER(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!xyt) return GMS_ERR_INVALID_ARG;
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->field_state == gms_handle::FIELD_BEFORE_LAST && !h->use_upd_pose) {
        // the virtual likelihood fields are defined through the poses the last scan was integrated from: keep them
        { int rc_ = materialize_poses(h); if (rc_) return rc_; }
        for (int i = 0; i < 2; i++)
            if (!h->upd_pose[i]) CK(cudaMalloc((void**)&h->upd_pose[i], (size_t)h->P * sizeof(float4)));
        CK(cudaMemcpyAsync(h->upd_pose[0], h->pose[h->cur], (size_t)h->P * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
        h->use_upd_pose = true;
    }
    CK(cudaMemcpyAsync(h->d_tmp, xyt, (size_t)h->P * 12, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(GMS_PHASE_COUNT - 1,
           k_pose_pack<<<blocks_for(h->P, 256), 256, 0, h->stream>>>((const float*)h->d_tmp, h->pose[h->cur], h->P));
    CK(cudaStreamSynchronize(h->stream));
    h->wpose_valid = false;
    return GMS_OK;
}
EXPORT int gms_set_weights(gms_handle* h, const double* w) {
    ENTER_KEEP(h);
    { int rc_ = complete_resample(h); if (rc_) return rc_; }
    if (!w) return GMS_ERR_INVALID_ARG;
    CK(cudaMemcpyAsync(h->w[h->cur], w, (size_t)h->P * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tile_fx_valid = false;
    h->wpose_valid = false;
    return GMS_OK;
}
EXPORT int gms_set_map_counts(gms_handle* h, int32_t particle, const uint32_t* nf, const uint32_t* no) {
    ENTER(h);
    if (!nf || !no) return GMS_ERR_INVALID_ARG;
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    if ((rc = pin_field(h, particle, s))) return rc;
    uint32_t* tmp = (uint32_t*)h->d_tmp;  // cells*8 bytes: two u32 planes
    CK(cudaMemcpyAsync(tmp, nf, h->cells * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(tmp + h->cells, no, h->cells * 4, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(GMS_PHASE_COUNT - 1, k_counts_join<<<blocks_for((long long)h->cells, 256), 256, 0, h->stream>>>(
                                    h->counts + (size_t)s * h->cells, tmp, tmp + h->cells, h->cells));
    LAUNCH(GMS_PHASE_COUNT - 1, k_fill_rect<<<1, 1, 0, h->stream>>>(h->rect + s, 1, make_int4(0, 0, h->W - 1, h->H - 1)));
    if (h->cfg.map_mode == GMS_MAP_SHARED)
        LAUNCH(GMS_PHASE_COUNT - 1, k_fill_dirty<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                        h->dirty, 1, h->g.tile_words, h->tiles_per_map));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

// ---- GridMap operators on one map ---------------------------------------------------------------
EXPORT int gms_map_apply_measurement(gms_handle* h, int32_t particle, float sx, float sy, float ex, float ey,
                                     float meas, int32_t was_hit) {
    ENTER(h);
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    if ((rc = pin_field(h, particle, s))) return rc;
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE)
        LAUNCH(GMS_PHASE_MAP_UPDATE, k_apply_one_red<<<1, 1, 0, h->stream>>>(h->counts + (size_t)s * h->cells, h->rect + s, sx, sy,
                                                                            ex, ey, meas, was_hit, h->g));
    else
        LAUNCH(GMS_PHASE_MAP_UPDATE, k_apply_one<<<1, 1, 0, h->stream>>>(h->counts, h->rect, h->dirty, sx, sy, ex, ey, meas,
                                                                        was_hit, h->g));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

static int stage_pose_slot(gms_handle* h, const float pose[3], int s) {
    const float4 p = make_float4(pose[0], pose[1], pose[2], 0.f);
    CK(cudaMemcpyAsync(h->tmp_pose, &p, sizeof p, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tmp_slot, &s, sizeof s, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));  // the sources are stack variables
    return GMS_OK;
}

EXPORT int gms_map_integrate_observation(gms_handle* h, int32_t particle, const float pose[3], const double* bxy,
                                         const double* bdist, const uint8_t* bhit, int32_t B) {
    ENTER(h);
    if (!pose || B < 0 || (B > 0 && (!bxy || !bdist || !bhit))) return GMS_ERR_INVALID_ARG;
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    if ((rc = pin_field(h, particle, s))) return rc;
    if ((rc = upload_beams(h, h->ops, bxy, bdist, bhit, B))) return rc;
    if ((rc = launch_pack(h, h->ops, (const double*)h->ops.xy, h->ops.dist, h->ops.hit, B))) return rc;
    if ((rc = stage_pose_slot(h, pose, s))) return rc;
    if ((rc = launch_map_update(h, h->ops, h->tmp_pose, 0, 1, h->tmp_slot, B))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

EXPORT int gms_map_compute_likelihood(gms_handle* h, int32_t particle) {
    ENTER(h);
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE) {  // virtual field: from now on blur(codes of this slot's counters)
        auto it = h->field_ovr.find(s);
        if (it != h->field_ovr.end()) cudaFree(it->second.snap);
        h->field_ovr[s] = gms_handle::FieldOverride{true, nullptr};
        return GMS_OK;
    }
    // the whole map (GridMap.computeLikelihoodMap has no notion of dirty tiles)
    LAUNCH(GMS_PHASE_LIKELIHOOD, k_fill_dirty<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                     h->dirty, 1, h->g.tile_words, h->tiles_per_map));
    if ((rc = launch_likelihood(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

EXPORT int gms_map_probability_of(gms_handle* h, int32_t particle, const float pose[3], const double* bxy,
                                  const uint8_t* bhit, int32_t B, double* log_prob, double* prob) {
    ENTER(h);
    if (!pose || B < 0 || (B > 0 && (!bxy || !bhit))) return GMS_ERR_INVALID_ARG;
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    if ((rc = upload_beams(h, h->ops, bxy, nullptr, bhit, B))) return rc;
    if ((rc = launch_pack(h, h->ops, (const double*)h->ops.xy, h->ops.dist, h->ops.hit, B))) return rc;
    const double* f = nullptr;
    if ((rc = field_of(h, particle, s, &f))) return rc;
    if ((rc = stage_pose_slot(h, pose, 0))) return rc;
    if ((rc = launch_score(h, h->ops, h->tmp_pose, 0, 1, nullptr, h->tmp_lw, nullptr, B, false, f))) return rc;
    double lw = 0;
    if ((rc = copy_out(h, &lw, h->tmp_lw, 8))) return rc;
    if (log_prob) *log_prob = lw;
    if (prob) *prob = std::exp(lw);
    return GMS_OK;
}

EXPORT int gms_trace_rays(gms_handle* h, const float* rays, int32_t n, int32_t extra, int32_t* cells_xy, int32_t cap,
                          int32_t* counts) {
    ENTER_KEEP(h);
    if (!rays || !counts || n < 0 || cap < 0 || extra < 0 || (cap > 0 && !cells_xy)) return GMS_ERR_INVALID_ARG;
    if (n == 0) return GMS_OK;
    float4* d_rays = nullptr;
    int2* d_cells = nullptr;
    int* d_counts = nullptr;
    auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_cells); cudaFree(d_counts); };
    const size_t cb = (size_t)n * std::max(cap, 1) * sizeof(int2);
    cudaError_t e;
    if ((e = cudaMalloc((void**)&d_rays, (size_t)n * 16)) != cudaSuccess ||
        (e = cudaMalloc((void**)&d_cells, cb)) != cudaSuccess ||
        (e = cudaMalloc((void**)&d_counts, (size_t)n * 4)) != cudaSuccess) {
        cleanup();
        return cuda_fail(h, e, "gms_trace_rays: cudaMalloc");
    }
    int rc = GMS_OK;
    do {
        if ((e = cudaMemcpyAsync(d_rays, rays, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(d_cells, 0xff, cb, h->stream)) != cudaSuccess) break;
        k_trace_rays<<<blocks_for(n, 128), 128, 0, h->stream>>>(d_rays, n, extra, h->W, h->H, d_cells, cap, d_counts);
        count_launch(h, GMS_PHASE_MAP_UPDATE);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if (cap > 0 &&
            (e = cudaMemcpyAsync(cells_xy, d_cells, (size_t)n * cap * 8, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
            break;
        if ((e = cudaMemcpyAsync(counts, d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(h->stream);
    } while (0);
    if (e != cudaSuccess) rc = cuda_fail(h, e, "gms_trace_rays");
    cleanup();
    return rc;
}

// Odometry(int,int) Odometry.java:41-55; MathUtil.PI is float pi (MathUtil.java:21); Robot.java:8-14
EXPORT int gms_odometry_from_counts(int32_t left, int32_t right, double* d_center, double* d_theta) {
    if (!d_center || !d_theta) return GMS_ERR_INVALID_ARG;
    const double pif = (double)(float)M_PI;
    const double dl = (double)left / 960 * pif * 0.063;
    const double dr = (double)right / 960 * pif * 0.063;
    *d_center = (dl + dr) / 2;
    *d_theta = (dr - dl) / 0.22;
    return GMS_OK;
}

// ---- device-resident / multi-rank -----------------------------------------------------------------
EXPORT int gms_update_begin_dev(gms_handle* h, const double* d_xy, const double* d_dist, const uint8_t* d_hit,
                                int32_t B, double d_center, double d_theta, const double* d_normals) {
    ENTER_STEP(h);
    if (int rc_ = check_beams(h, B, d_xy, d_dist, d_hit, "gms_update_begin_dev")) return rc_;
    flip_beams(h);
    h->fold_wpose = false;
    return step_begin(h, d_xy, d_dist, d_hit, B, d_center, d_theta, d_normals);
}
EXPORT int gms_join_streams(gms_handle* h) {
    ENTER_KEEP(h);  // the main stream now waits for the map integration running on the side stream
    return GMS_OK;
}
EXPORT int gms_set_pose_optimizer(gms_handle* h, gms_pose_optimizer_fn fn, void* user) {
    ENTER_KEEP(h);
    h->opt_fn = fn;
    h->opt_user = user;
    return GMS_OK;
}
EXPORT int gms_update_end_dev(gms_handle* h, int32_t policy, double u01) {
    ENTER_KEEP(h);
    if (policy < 0 || policy > 2 || u01 >= 1.0) return fail(h, GMS_ERR_INVALID_ARG, "bad resample policy / u01");
    return step_end(h, policy, u01);
}
EXPORT int gms_step_dev(gms_handle* h, const double* d_xy, const double* d_dist, const uint8_t* d_hit, int32_t B,
                        double d_center, double d_theta, const double* d_normals, int32_t policy, double u01) {
    ENTER_STEP(h);
    if (h->cfg.nranks != 1) return fail(h, GMS_ERR_STATE, "gms_step_dev: multi-rank handles use update_begin/end");
    if (int rc_ = check_beams(h, B, d_xy, d_dist, d_hit, "gms_step_dev")) return rc_;
    if (policy < 0 || policy > 2 || u01 >= 1.0) return fail(h, GMS_ERR_INVALID_ARG, "bad resample policy / u01");
    flip_beams(h);
    h->fold_wpose = false;
    int rc = step_begin(h, d_xy, d_dist, d_hit, B, d_center, d_theta, d_normals);
    if (rc) return rc;
    return step_end(h, policy, u01);
}
EXPORT int gms_exchange_buffers(gms_handle* h, void** dl, size_t* lb, void** dg, size_t* gb) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (dl) *dl = h->xlocal;
    if (lb) *lb = (size_t)h->cnt * sizeof(ExchangeRec);
    if (dg) *dg = h->xglobal;
    if (gb) *gb = (size_t)h->P * sizeof(ExchangeRec);
    return GMS_OK;
}
EXPORT int gms_read_neff(gms_handle* h, double* neff) {
    ENTER_KEEP(h);
    if (!neff) return GMS_ERR_INVALID_ARG;
    int rc = fetch_stats(h);
    if (rc) return rc;
    *neff = h->h_st->neff;
    return GMS_OK;
}
EXPORT int gms_sync(gms_handle* h) {
    ENTER_KEEP(h);  // joins a pending map integration into the main stream first
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}
EXPORT int gms_set_stream(gms_handle* h, void* s) {
    ENTER_KEEP(h);
    CK(cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return GMS_OK;
}
EXPORT int gms_profile_enable(gms_handle* h, int32_t on) {
    ENTER_KEEP(h);
    int rc = flush_profile(h);
    h->profile = on != 0;
    return rc;
}
EXPORT int gms_profile_read(gms_handle* h, double* ms, int64_t* launches) {
    ENTER_KEEP(h);
    int rc = flush_profile(h);
    if (rc) return rc;
    for (int i = 0; i < GMS_PHASE_COUNT; i++) {
        if (ms) ms[i] = h->phase_ms[i];
        if (launches) launches[i] = h->phase_launches[i];
    }
    return GMS_OK;
}
EXPORT int gms_profile_reset(gms_handle* h) {
    ENTER_KEEP(h);
    int rc = flush_profile(h);
    for (int i = 0; i < GMS_PHASE_COUNT; i++) { h->phase_ms[i] = 0; h->phase_launches[i] = 0; }
    return rc;
}
EXPORT int gms_launch_count(gms_handle* h, int64_t* n) {
    if (!h || !n) return GMS_ERR_INVALID_ARG;
    *n = h->launches;
    return GMS_OK;
}

// ---- per-particle maps across ranks: peer mappings (one process per GPU on one node) ----------------
EXPORT int gms_ipc_export(gms_handle* h, void* handles) {
    ENTER_KEEP(h);
    if (!handles) return GMS_ERR_INVALID_ARG;
    if (h->cfg.nranks < 2) return fail(h, GMS_ERR_STATE, "gms_ipc_export: single-rank handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == GMS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t* out = static_cast<cudaIpcMemHandle_t*>(handles);
    void* ptr[GMS_IPC_NUM_HANDLES] = {h->counts, h->lik, h->rect, h->dirty, h->xlw[0], h->xlw[1], h->xflags,
                                      h->pose[0], h->pose[1], h->w[0], h->w[1], h->lw[0], h->lw[1], h->parents,
                                      h->cdf, h->xarea[0], h->xarea[1], h->xflags4};
    for (int k = 0; k < GMS_IPC_NUM_HANDLES; k++) CK(cudaIpcGetMemHandle(&out[k], ptr[k]));
    return GMS_OK;
}
EXPORT int gms_ipc_import(gms_handle* h, const void* all_handles) {
    ENTER_KEEP(h);
    if (!all_handles) return GMS_ERR_INVALID_ARG;
    if (h->cfg.nranks < 2) return fail(h, GMS_ERR_STATE, "gms_ipc_import: single-rank handle");
    const cudaIpcMemHandle_t* in = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int q = 0; q < h->cfg.nranks; q++) {
        void* ptr[GMS_IPC_NUM_HANDLES] = {h->counts, h->lik, h->rect, h->dirty, h->xlw[0], h->xlw[1], h->xflags,
                                          h->pose[0], h->pose[1], h->w[0], h->w[1], h->lw[0], h->lw[1], h->parents,
                                          h->cdf, h->xarea[0], h->xarea[1], h->xflags4};
        if (q != h->cfg.rank)
            for (int k = 0; k < GMS_IPC_NUM_HANDLES; k++) {
                if (h->ipc_opened[q][k]) { ptr[k] = h->ipc_opened[q][k]; continue; }
                CK(cudaIpcOpenMemHandle(&ptr[k], in[q * GMS_IPC_NUM_HANDLES + k], cudaIpcMemLazyEnablePeerAccess));
                h->ipc_opened[q][k] = ptr[k];
            }
        h->peers.counts[q] = static_cast<const CellCounts*>(ptr[0]);
        h->peers.rect[q] = static_cast<const int4*>(ptr[2]);
        h->peer_xlw[0][q] = static_cast<double*>(ptr[4]);
        h->peer_xlw[1][q] = static_cast<double*>(ptr[5]);
        h->peer_flags[q] = static_cast<unsigned long long*>(ptr[6]);
        h->peer_pose[0][q] = static_cast<const float4*>(ptr[7]);
        h->peer_pose[1][q] = static_cast<const float4*>(ptr[8]);
        h->peer_w[0][q] = static_cast<const double*>(ptr[9]);
        h->peer_w[1][q] = static_cast<const double*>(ptr[10]);
        h->peer_lw[0][q] = static_cast<const double*>(ptr[11]);
        h->peer_lw[1][q] = static_cast<const double*>(ptr[12]);
        h->peer_parents[q] = static_cast<const int*>(ptr[13]);
        h->peer_cdf[q] = static_cast<const unsigned long long*>(ptr[14]);
        h->peer_xarea[0][q] = static_cast<unsigned char*>(ptr[15]);
        h->peer_xarea[1][q] = static_cast<unsigned char*>(ptr[16]);
        h->peer_xflags4[q] = static_cast<unsigned long long*>(ptr[17]);
    }
    h->peers_ready = true;
    h->direct = true;
    // shared map, GMS_SHARDED=1: every rank normalises and resamples only its own block (three tiny exchange rounds
    // per step inside the normalise kernel + one in the resampling).  Measured slower than the default — the push of
    // the log-weights + replicated normalise / CDF — at every rank count (8 x 100k particles: 0.276 vs 0.264 ms per
    // step, 4 x: 0.251 vs 0.235, 2 x: 0.245 vs 0.226; each round costs ~8 us of NVLink flag latency + grid barrier),
    // so it stays an option.
    h->sharded_post = h->cfg.map_mode == GMS_MAP_SHARED && std::getenv("GMS_SHARDED") &&
                      std::atoi(std::getenv("GMS_SHARDED")) != 0;
    return GMS_OK;
}

// ---- rows adjacent to the path ------------------------------------------------------
namespace {
// raw sweep -> device (raw_angle, raw_dist, all_hit) -> de-skewed beam table (all_xy, in_dist)
int upload_raw_and_deskew(gms_handle* h, BeamSet& b, const double* angle, const double* dist, const uint8_t* hit, int B,
                          double d_center, double d_theta, const double* normals = nullptr) {
    int rc = ensure_beams(h, B);
    if (rc) return rc;
    if (B == 0 && !normals) return GMS_OK;
    const size_t off_n = (((size_t)B * 17 + 63) / 64) * 64;
    unsigned char* s = nullptr;
    int slot = 0;
    if ((rc = stage_acquire(h, off_n + (normals ? (size_t)h->cnt * 16 : 0) + 64, &s, &slot))) return rc;
    b.view(b.cap);
    if (B > 0) {
        std::memcpy(s, angle, (size_t)B * 8);
        std::memcpy(s + (size_t)B * 8, dist, (size_t)B * 8);
        if (hit) std::memcpy(s + (size_t)B * 16, hit, (size_t)B);
        CK(cudaMemcpyAsync(h->raw_angle, s, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->raw_dist, s + (size_t)B * 8, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
        if (hit) CK(cudaMemcpyAsync(b.hit, s + (size_t)B * 16, (size_t)B, cudaMemcpyHostToDevice, h->stream));
        LAUNCH(GMS_PHASE_COUNT - 1, k_deskew<<<blocks_for(B, 256), 256, 0, h->stream>>>(h->raw_angle, h->raw_dist, B, d_center,
                                                                                       d_theta, b.xy, b.dist));
    }
    if (normals) {
        std::memcpy(s + off_n, normals, (size_t)h->cnt * 16);
        CK(cudaMemcpyAsync(h->d_normals, s + off_n, (size_t)h->cnt * 16, cudaMemcpyHostToDevice, h->stream));
    }
    return stage_release(h, slot);
}
}  // namespace

EXPORT int gms_deskew(gms_handle* h, const double* angle, const double* dist, int32_t B, double d_center,
                      double d_theta, double* out_xy, double* out_dist) {
    ENTER_KEEP(h);
    if (B < 0 || (B > 0 && (!angle || !dist || !out_xy || !out_dist))) return fail(h, GMS_ERR_INVALID_ARG, "gms_deskew: bad arrays");
    int rc = upload_raw_and_deskew(h, h->ops, angle, dist, nullptr, B, d_center, d_theta);
    if (rc || B == 0) return rc;
    CK(cudaMemcpyAsync(out_xy, h->ops.xy, (size_t)B * 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(out_dist, h->ops.dist, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

EXPORT int gms_update_raw(gms_handle* h, const double* angle, const double* dist, const uint8_t* hit, int32_t B,
                          double d_center, double d_theta, const double* normals, double* neff_out) {
    ENTER_STEP(h);
    if (int rc_ = check_beams(h, B, angle, dist, hit, "gms_update_raw")) return rc_;
    if (h->cfg.nranks != 1) return fail(h, GMS_ERR_STATE, "gms_update_raw: multi-rank handles use update_begin/end");
    flip_beams(h);
    h->fold_wpose = true;
    BeamSet& bs = cur_beams(h);
    int rc = upload_raw_and_deskew(h, bs, angle, dist, hit, B, d_center, d_theta, normals);
    if (rc) return rc;
    if ((rc = step_begin(h, (const double*)bs.xy, bs.dist, bs.hit, B, d_center, d_theta, normals ? h->d_normals : nullptr))) return rc;
    if ((rc = step_end(h, GMS_RESAMPLE_NEVER, 0.0))) return rc;
    if ((rc = fetch_stats(h))) return rc;
    if (neff_out) *neff_out = h->h_st->neff;
    return GMS_OK;
}

EXPORT int gms_render_map(gms_handle* h, int32_t particle, int32_t likelihood, uint32_t* abgr_out) {
    ENTER(h);
    if (!abgr_out) return GMS_ERR_INVALID_ARG;
    int s;
    int rc = slot_of(h, particle, &s);
    if (rc) return rc;
    const double* f = h->lik;
    if (likelihood && (rc = field_of(h, particle, s, &f))) return rc;
    LAUNCH(GMS_PHASE_COUNT - 1, k_render<<<blocks_for((long long)h->cells, 256), 256, 0, h->stream>>>(
                                    h->counts + (size_t)s * h->cells, f, h->cells, likelihood,
                                    h->g.l_free, h->g.l_occ, (uint32_t*)h->d_tmp));
    return copy_out(h, abgr_out, h->d_tmp, h->cells * 4);
}

namespace {
int ensure_combined(gms_handle* h) {
    if (h->comb_log) return GMS_OK;
    CK(cudaMalloc((void**)&h->comb_prod, h->cells * 8));
    CK(cudaMalloc((void**)&h->comb_log, h->cells * 8));
    CK(cudaMalloc((void**)&h->comb_lik, h->cells * 8));
    CK(cudaMalloc((void**)&h->comb_sign, h->cells * sizeof(CellCounts)));
    CK(cudaMalloc((void**)&h->comb_dirty, (size_t)h->g.tile_words * 4));
    CK(cudaMalloc((void**)&h->comb_off, (size_t)h->g.tile_words * 4));
    CK(cudaMalloc((void**)&h->comb_list, (size_t)h->tiles_per_map * sizeof(int2)));
    return GMS_OK;
}
}  // namespace

EXPORT int gms_combined_map_begin_dev(gms_handle* h, void** d_product, size_t* bytes) {
    ENTER(h);
    if (h->cfg.map_mode != GMS_MAP_PER_PARTICLE)
        return fail(h, GMS_ERR_UNSUPPORTED, "gms_combined_map: per-particle maps only (a shared map is its own fusion)");
    int rc = ensure_combined(h);
    if (rc) return rc;
    LAUNCH(GMS_PHASE_COUNT - 1, k_combine_product<<<blocks_for((long long)h->cells, 256), 256, 0, h->stream>>>(
                                    h->counts, h->slot[h->slot_cur] + h->lo, h->cnt, h->cells, h->g.l_free, h->g.l_occ,
                                    h->comb_prod));
    if (d_product) *d_product = h->comb_prod;
    if (bytes) *bytes = h->cells * 8;
    return GMS_OK;
}

EXPORT int gms_combined_map_end(gms_handle* h, double* log_out, double* lik_out) {
    ENTER(h);
    if (!h->comb_prod) return fail(h, GMS_ERR_STATE, "gms_combined_map_end without gms_combined_map_begin_dev");
    LAUNCH(GMS_PHASE_COUNT - 1, k_combine_finish<<<blocks_for((long long)h->cells, 256), 256, 0, h->stream>>>(
                                    h->comb_prod, h->cells, h->comb_log, h->comb_sign));
    if (log_out) CK(cudaMemcpyAsync(log_out, h->comb_log, h->cells * 8, cudaMemcpyDeviceToHost, h->stream));
    if (lik_out) {
        // GridMap.computeLikelihoodMap(combinedGrid): the same tile kernel on the sign map
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_fill_dirty<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                         h->comb_dirty, 1, h->g.tile_words, h->tiles_per_map));
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_scan<<<1, 1024, 0, h->stream>>>(h->comb_dirty, h->g.tile_words, h->comb_off,
                                                                           h->st));
        LAUNCH(GMS_PHASE_LIKELIHOOD, k_lik_emit<<<blocks_for(h->g.tile_words, 256), 256, 0, h->stream>>>(
                                         h->comb_dirty, h->g.tile_words, h->g.tile_words, h->comb_off, h->comb_list));
        int rc = launch_blur(h, h->comb_sign, h->comb_lik, nullptr, h->comb_list, SelfList{nullptr, 0}, h->tiles_per_map, false);
        if (rc) return rc;
        CK(cudaMemcpyAsync(lik_out, h->comb_lik, h->cells * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    h->stats_valid = false;
    CK(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

EXPORT int gms_combined_map(gms_handle* h, double* log_out, double* lik_out) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (h->cfg.nranks != 1)
        return fail(h, GMS_ERR_STATE, "gms_combined_map: multi-rank handles use gms_combined_map_begin_dev, a PRODUCT "
                                      "all-reduce of the returned buffer over the ranks, gms_combined_map_end");
    int rc = gms_combined_map_begin_dev(h, nullptr, nullptr);
    if (rc) return rc;
    return gms_combined_map_end(h, log_out, lik_out);
}
