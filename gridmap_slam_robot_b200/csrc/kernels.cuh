// kernels.cuh — the sm_100a kernels of the SLAM hot path.
// No tensor cores: nothing here is a dense contraction.  Data layout (HBM), per handle:
//   pose   float4[P]                 {x, y, theta, 0}      (Pose.java:21-34)
//   lw, w  f64[P]                    ln(product) of the last update / normalised weight
//   counts CellCounts[S][H][W]       8 B per cell, row-major idx = x + y*W (GridMap.java:135)
//   lik    f64[S][H][W]              likelihood field (GridMapData.likelihoodData)
//   rect   int4[S]                   bounding box of every cell touched since reset ("explored")
//   dirty  u32[S][tile_words]        bitmap of likelihood tiles whose thresholded codes changed
#pragma once
#include <cuda.h>  // CUtensorMap
#include <cooperative_groups.h>
#include <type_traits>
#include "device_math.cuh"

namespace gms {
namespace cg = cooperative_groups;

struct Stats {
    double neff;          // SLAM.calculateNeff of the last update
    double lw_max;        // max log-weight (the strongest particle's)
    double sum_exp;       // sum exp(lw - lw_max)
    double strongest_w;   // normalised weight of the strongest particle
    double neff_query;    // result slot of gms_calculate_neff
    float strongest_pose[4];
    float weighted_pose[4];
    int strongest;        // first arg-max of lw (SLAM.java:110-115 keeps the first maximum: strict >)
    int do_resample;      // decided on the device from the policy (GridMapApp.java:185)
    int num_hit;          // beams with wasHit (GridMap.java:269-270)
    int num_dup;          // map copies of the last resample
    int num_tiles;        // likelihood work-list length
    int xerror;           // peer exchange: a peer's log-weights did not arrive in time (every later kernel of the
                          // step returns early, so no state is mutated from stale records)
    int strongest_now;    // current index of the particle that was strongest at the last update: after a
                          // resampling its first child (GridMapApp keeps drawing strongestParticle.m), -1 if none
    int pad[1];
};

struct ExchangeRec {  // 24 B, gms.h "Exchange record"
    double lw;
    float x, y, t;
    uint32_t pad;
};

struct NormPartials {   // partial results of normalise: per score CTA (m, idx, s) and per 1024-particle tile (ws, q, fx)
    double* m;          // tile max of lw
    int* idx;           // first index of the tile max
    double* s;          // sum exp(lw - tile max)
    double* ws;         // sum of normalised weights of the tile
    double* q;          // sum of squared normalised weights of the tile
    unsigned long long* fx;  // sum of trunc(w * 2^60) of the tile (feeds the fixed-point CDF)
    unsigned long long* x128;  // 6 u64 per tile: exact 128-bit sums of exp(lw - M), w^2, w (U128 below)
    unsigned* counter;  // last-block-done ticket
};
#define kNegInf (__longlong_as_double((long long)0xfff0000000000000ULL))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kMaxRanks = 16;
struct PeerTable {  // per-particle maps across ranks: every rank's arenas, mapped into this process (cudaIpc)
    const CellCounts* counts[kMaxRanks];
    const int4* rect[kMaxRanks];
};

// Peer exchange (multi-rank, replaces the NCCL all-gather).  Only the f64 log-weights travel: after scoring,
// k_xpush_lw stores this rank's block of lw straight into EVERY rank's receive buffer (peer-mapped, NVLink) with
// coalesced 8-byte stores — 8 B per particle and peer instead of the 24-byte {lw, pose} record — and the last
// CTA to finish raises one flag per receiver carrying the exchange sequence number (st.release.sys after
// __threadfence_system).  The consumer (k_norm_coop) polls the flags itself (ld.acquire.sys, bounded spin), so
// no separate wait / import kernels run.  Poses are NOT exchanged: the resampling reads a remote parent's
// 16-byte pose through the peer mapping of that rank's pose array (PoseTable) — one dependent NVLink read per
// child whose parent lives elsewhere, and only for this rank's own children.
// (History: pushing 24-byte records from inside the scoring kernel doubled the scoring time on 8 GPUs; pushing
// whole {lw, pose} blocks + an import pass cost 19 MB of NVLink stores per rank and step at 8 x 100k.)
struct XPush {
    int nranks;
    double* dst[kMaxRanks];               // receive buffer (f64[P]) of every rank for this exchange parity
    unsigned long long* flag[kMaxRanks];  // flag[q] = rank q's flag array (one u64 per sender)
};
__global__ void __launch_bounds__(1024) k_xpush_lw(const double* __restrict__ lw, int lo, int cnt, XPush xp,
                                                   int myrank, unsigned long long seq,
                                                   unsigned* __restrict__ ticket) {
    __shared__ bool s_last;
    // two particles (16 bytes) per store; destination-major so that a warp's stores form full 128-byte lines
    const int pairs = cnt >> 1;
    const long long total = (long long)xp.nranks * pairs;
    const bool aligned = ((lo | cnt) & 1) == 0;
    if (aligned) {
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (long long)gridDim.x * blockDim.x) {
            const int q = (int)(e / pairs), j = (int)(e - (long long)q * pairs);
            reinterpret_cast<double2*>(xp.dst[q] + lo)[j] = reinterpret_cast<const double2*>(lw + lo)[j];
        }
    } else {
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)xp.nranks * cnt;
             e += (long long)gridDim.x * blockDim.x) {
            const int q = (int)(e / cnt), i = lo + (int)(e - (long long)q * cnt);
            xp.dst[q][i] = lw[i];
        }
    }
    // one system-scope fence per CTA: the barrier orders the CTA's stores before thread 0's fence (cumulativity),
    // the ticket orders every CTA's fence before the last CTA's flag stores
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        *ticket = 0u;
    }
    __syncthreads();
    if ((int)threadIdx.x < xp.nranks)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(xp.flag[threadIdx.x] + myrank), "l"(seq) : "memory");
}
// The same exchange as a PULL (default; GMS_PULL=0 keeps the push above).  The scoring kernel writes this rank's
// block of log-weights into its OWN exchange buffer; nothing is stored over NVLink and no exchange kernel runs.
// The normalise kernel that follows in stream order raises this rank's flag on every rank first (the kernel
// boundary has made the scoring kernel's stores visible at this GPU's L2, which is where a peer's NVLink read is
// served), waits for every rank's flag, and reads each 1024-particle tile from the buffer of the rank that owns
// it (8-byte __ldcg loads through the peer mapping: L1 is bypassed, the requester's L2 does not
// cache peer memory), filing the values in the handle's own lw array.  Per rank and step 8 B x (P - cnt) arrive
// over NVLink in one round trip instead of 8 B x cnt x R stores, two system fences and a launch (8 x 100k
// particles: exchange phase 0.023 ms -> 0).  Buffers are double-buffered by exchange parity: a rank rewrites a
// buffer two steps later, and it cannot get there before every peer has raised its next flag, which a peer does
// only after its reads of this step (stream order).
struct XPull {
    int cnt;                              // particles per rank; 0: not pulling
    int myrank;
    const double* src[kMaxRanks];         // every rank's exchange buffer (f64[P], GLOBAL index) for this parity
    unsigned long long* flag[kMaxRanks];  // flag[q] = rank q's flag array (one u64 per sender)
};
// (Raising the flags from the scoring kernel's last CTA instead — a launch gap earlier — was measured and dropped: the
// per-CTA system-scope fence ahead of the ticket cost the scoring kernel 14 us, the normalise gained 5.)
// every rank's pose array of the current generation, indexable by GLOBAL particle index (single rank, or
// records imported by an all-gather: every entry is the local array)
struct PoseTable {
    const float4* p[kMaxRanks];
    int cnt;  // particles per rank
    __device__ __forceinline__ float4 at(int i) const { return p[i / cnt][i]; }
};

// ------------------------------------------------------------------------------------------------
// beam table: compaction of the hit beams (scoring reads only those, GridMap.java:269-270) and the
// per-beam measured distance in cells, (float) m.distance / resolution (GridMap.java:188).
// ------------------------------------------------------------------------------------------------
// Executed by ONE whole CTA (any block size that is a multiple of 32, <= 1024).
__device__ __forceinline__ void pack_beams_cta(const double2* __restrict__ in_xy, const double* __restrict__ in_dist,
                                               const uint8_t* __restrict__ in_hit, int B, float res_f,
                                               double2* __restrict__ hit_xy, float* __restrict__ meas,
                                               double2* __restrict__ all_xy, uint8_t* __restrict__ all_hit,
                                               int* __restrict__ num_hit, double* __restrict__ rmax2) {
    __shared__ int s_warp[32];
    __shared__ double s_r[32];
    __shared__ int s_base;
    double r2 = 0.0;  // longest hit beam (squared length, metres): range guard of k_score_sorted<G, 2>
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + tid;
        const bool hit = b < B && in_hit[b] != 0;
        if (hit) {
            const double2 v = in_xy[b];
            const double d2 = v.x * v.x + v.y * v.y;
            r2 = d2 > r2 || d2 != d2 ? d2 : r2;  // a NaN beam poisons the guard: exact path
        }
        if (b < B) {
            meas[b] = (float)in_dist[b] / res_f;
            if (all_xy != in_xy) all_xy[b] = in_xy[b];  // private copy: the map integration reads it later
            if (all_hit != in_hit) all_hit[b] = in_hit[b];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wid; k++) off += s_warp[k];
        if (hit) hit_xy[off + __popc(m & ((1u << lane) - 1))] = in_xy[b];
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int k = 0; k < nw; k++) t += s_warp[k];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) *num_hit = s_base;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, r2, o);
        r2 = v > r2 || v != v ? v : r2;
    }
    if (lane == 0) s_r[wid] = r2;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < nw; k++) r2 = s_r[k] > r2 || s_r[k] != s_r[k] ? s_r[k] : r2;
        *rmax2 = r2;
    }
}
__global__ void __launch_bounds__(256) k_pack_beams(const double2* __restrict__ in_xy,
                                                    const double* __restrict__ in_dist,
                                                    const uint8_t* __restrict__ in_hit, int B, float res_f,
                                                    double2* __restrict__ hit_xy, float* __restrict__ meas,
                                                    double2* __restrict__ all_xy, uint8_t* __restrict__ all_hit,
                                                    int* __restrict__ num_hit, double* __restrict__ rmax2) {
    pack_beams_cta(in_xy, in_dist, in_hit, B, res_f, hit_xy, meas, all_xy, all_hit, num_hit, rmax2);
}

// ------------------------------------------------------------------------------------------------
// A2 — SLAM.sampleMotionModel SLAM.java:155-163 + Odometry.apply Odometry.java:77-96.
// One thread per local particle; z = {z_d, z_theta} injected or Philox(seed, global index, step).
// ------------------------------------------------------------------------------------------------
// 65536 heading buckets of 2*pi/65536 rad: ~30 particles per bucket at 100k particles spread over +-15 deg.
// Same-address atomics with a return value serialise at ~70 ns each (ncu: 8192 buckets made k_motion 23 us,
// 65536 buckets 10 us), so the bucket count is chosen for low contention, not for ordering precision.
constexpr int kSortBins = 65536;
constexpr int kSortChunks = kSortBins / 1024;

struct MotionArgs {
    float4* pose;
    int lo, cnt;
    const double* normals;
    uint64_t seed, step;
    double d_center, d_theta, sd_c, sd_t;
};
__device__ __forceinline__ float4 motion_one(const MotionArgs& a, int li) {
    const int i = a.lo + li;
    double zd, zt;
    if (a.normals) {
        zd = a.normals[2 * li];
        zt = a.normals[2 * li + 1];
    } else {
        philox_normals(a.seed, (uint32_t)i, a.step, zd, zt);
    }
    const double d = a.sd_c * zd + a.d_center;   // NormalDistribution.sample(): sd * z + mean
    const double th = a.sd_t * zt + a.d_theta;
    float4 p = a.pose[i];
    p.z = (float)angle_constrain((double)p.z + th);
    p.x = (float)((double)p.x + (double)cos_f(p.z) * d);
    p.y = (float)((double)p.y + (double)sin_f(p.z) * d);
    a.pose[i] = p;
    return p;
}
struct PackArgs {
    const double2* in_xy;
    const double* in_dist;
    const uint8_t* in_hit;
    int B;
    float res_f;
    double2* hit_xy;
    float* meas;
    double2* all_xy;
    uint8_t* all_hit;
    int* num_hit;
    double* rmax2;
};
// motion of every local particle; with `do_pack` the last CTA of the grid also packs the beam table (one launch less
// on the per-particle path, where nothing else shares the launch)
__global__ void __launch_bounds__(256) k_motion(MotionArgs a, PackArgs pk, int do_pack) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < a.cnt) motion_one(a, li);
    if (do_pack && blockIdx.x == gridDim.x - 1)
        pack_beams_cta(pk.in_xy, pk.in_dist, pk.in_hit, pk.B, pk.res_f, pk.hit_xy, pk.meas, pk.all_xy, pk.all_hit, pk.num_hit, pk.rmax2);
}

// Motion + heading sort + beam packing in ONE cooperative launch (shared map, many particles).  The sort only
// fixes a PROCESSING ORDER for k_score_sorted (which thread computes which particle): the rank inside a bucket
// is whatever order the atomics land in, and nothing downstream depends on it — the scoring kernel performs
// no reduction across particles, and every sum over particles (normalise, Neff, CDF) runs over fixed
// 1024-particle tiles of the particle index, so results are run-to-run bit-identical.
//   phase 1: motion update, bucket histogram (atomic with return = rank in bucket); last CTA packs the beams
//   phase 2: exclusive scan of the histogram, 1024 buckets per chunk, chunks strided over the CTAs
//   phase 3: scatter: order[chunk base + bucket offset + rank] = particle
struct SortBufs {
    unsigned *hist, *offs, *chunk_total, *key, *rank;
    int* order;
};
__global__ void __launch_bounds__(1024) k_motion_sort(MotionArgs a, SortBufs sb, PackArgs pk, int do_pack) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_base[kSortChunks];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int stride = gridDim.x * blockDim.x;
    for (int li = blockIdx.x * blockDim.x + tid; li < a.cnt; li += stride) {
        const float4 p = motion_one(a, li);
        const unsigned b = min(__float2uint_rz((p.z + 3.14159274f) * (kSortBins / 6.28318548f)), (unsigned)kSortBins - 1u);
        sb.key[li] = b;
        sb.rank[li] = atomicAdd(sb.hist + b, 1u);
    }
    if (do_pack && blockIdx.x == gridDim.x - 1)
        pack_beams_cta(pk.in_xy, pk.in_dist, pk.in_hit, pk.B, pk.res_f, pk.hit_xy, pk.meas, pk.all_xy, pk.all_hit, pk.num_hit, pk.rmax2);
    grid.sync();
    for (int c = blockIdx.x; c < kSortChunks; c += gridDim.x) {
        const int i = c * 1024 + tid;
        const unsigned v = sb.hist[i];
        sb.hist[i] = 0u;  // re-armed for the next step
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        __syncthreads();
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        unsigned wv = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, wv, o);
            if (lane >= o) wv += u;
        }
        const unsigned wbase = wid > 0 ? __shfl_sync(0xffffffffu, wv, wid - 1) : 0u;
        sb.offs[i] = wbase + inc - v;
        if (tid == 1023) sb.chunk_total[c] = wbase + inc;
    }
    grid.sync();
    if (tid < kSortChunks) {  // exclusive prefix of the 64 chunk totals (two warps, shuffle scan)
        const unsigned t = sb.chunk_total[tid];
        unsigned inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        s_base[tid] = inc - t;
    }
    __syncthreads();
    if (tid >= 32 && tid < kSortChunks) s_base[tid] += s_base[31] + sb.chunk_total[31];
    __syncthreads();
    for (int li = blockIdx.x * blockDim.x + tid; li < a.cnt; li += stride) {
        const unsigned k = sb.key[li];
        sb.order[s_base[k >> 10] + sb.offs[k] + sb.rank[li]] = li;
    }
}

// ------------------------------------------------------------------------------------------------
// A3 — GridMap.computeLikelihoodMap GridMap.java:233-250 + Util.doGaussianBlurdSeparable
// Util.java:378-426, restricted to the tiles that intersect each slot's dirty rectangle (+ the
// kernel half-width).  Cells outside keep their value: the field only depends on the sign of the
// log-odds within `khalf` cells, so an untouched neighbourhood reproduces itself bit for bit.
// Same f64 operation order as Java (tap index ascending, mul then add, no FMA) => bit-exact.
// ------------------------------------------------------------------------------------------------
// Work list: every set bit of the per-slot dirty-tile bitmaps becomes one {slot, tile} item.
// k_lik_scan (one CTA): exclusive scan of the per-word popcounts.  k_lik_emit: expands and clears the words.
__global__ void __launch_bounds__(1024) k_lik_scan(const uint32_t* __restrict__ dirty, int nwords,
                                                   int* __restrict__ word_off, Stats* __restrict__ st) {
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    const int per = (nwords + 1023) / 1024;
    const int i0 = min(nwords, tid * per), i1 = min(nwords, i0 + per);
    int sum = 0;
    for (int i = i0; i < i1; i++) sum += __popc(dirty[i]);
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - sum;
    for (int i = i0; i < i1; i++) {
        word_off[i] = run;
        run += __popc(dirty[i]);
    }
    if (tid == 1023) st->num_tiles = s_part[1023];
}
__global__ void __launch_bounds__(256) k_lik_emit(uint32_t* __restrict__ dirty, int nwords, int tile_words,
                                                  const int* __restrict__ word_off, int2* __restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t w = dirty[i];
    if (!w) return;
    dirty[i] = 0u;
    const int slot = i / tile_words, base = (i - slot * tile_words) * 32;
    int o = word_off[i];
    while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        list[o++] = make_int2(slot, base + bit);
    }
}

// Shared map (one slot, <= kSelfListWords bitmap words): the blur kernels build the work list themselves —
// every CTA scans the (tiny) dirty bitmap into a shared-memory prefix of popcounts and picks its tiles from
// it, so the likelihood refresh is ONE launch on the critical path map update -> refresh -> scoring instead of
// three.  The bitmap is double buffered by the host: nobody clears the buffer a refresh is reading.
constexpr int kSelfListWords = 2048;
struct SelfList {
    const uint32_t* bitmap;  // nullptr: use the {slot, tile} list built by k_lik_scan / k_lik_emit
    int nwords;
};
// all threads of the CTA; returns the number of dirty tiles, fills s_pref[0..nwords] (exclusive prefix)
__device__ __forceinline__ int self_list_build(const SelfList& sl, int* s_pref) {
    __shared__ int s_wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthreads = blockDim.x;
    int carry = 0;
    for (int base = 0; base < sl.nwords; base += nthreads) {
        const int i = base + tid;
        const int v = i < sl.nwords ? __popc(sl.bitmap[i]) : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        __syncthreads();
        if (lane == 31) s_wsum[wid] = inc;
        __syncthreads();
        int wv = lane < (nthreads >> 5) ? s_wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, wv, o);
            if (lane >= o) wv += u;
        }
        const int wbase = wid > 0 ? __shfl_sync(0xffffffffu, wv, wid - 1) : 0;
        const int total = __shfl_sync(0xffffffffu, wv, 31);
        if (i < sl.nwords) s_pref[i] = carry + wbase + inc - v;
        carry += total;
    }
    __syncthreads();
    if (tid == 0) s_pref[sl.nwords] = carry;
    __syncthreads();
    return carry;
}
// the t-th dirty tile (t < total): binary search over the prefix, then the n-th set bit of that word
__device__ __forceinline__ int2 self_list_item(const SelfList& sl, const int* s_pref, int t) {
    int lo = 0, hi = sl.nwords - 1;
    while (lo < hi) {  // last word whose exclusive prefix is <= t
        const int mid = (lo + hi + 1) >> 1;
        if (s_pref[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const uint32_t w = sl.bitmap[lo];
    const int bit = __fns(w, 0, t - s_pref[lo] + 1);
    return make_int2(0, lo * 32 + bit);
}

// Thresholded code {0, 1, 2} = {free, unknown, occupied}; the two cheap branches agree with cell_code().
__device__ __forceinline__ int cell_code_fast(const CellCounts c, const Geometry& g) {
    if (c.n_occ == 0) return c.n_free ? 0 : 1;
    if (c.n_free == 0) return 2;
    return cell_code(c.n_free, c.n_occ, g);
}

// Persistent CTAs walk the work list; one 64x32-cell tile per iteration.
//   smem s_t[(TH+2k)][tw_pad] f32 codes {0, .5, 1} (exact in f32), s_h[(TH+2k)][TW] f64 horizontal pass.
// KH = 3 (the reference's 0.05 m cells: 7 taps) is register-blocked: a thread produces 8 neighbouring
// outputs of a pass from 14 inputs held in registers, taps fully unrolled, eight independent f64
// accumulation chains (each in Java's order: tap index ascending, multiply then add, no FMA).
// KH = 0 is the generic run-time-width version.
template <int KH>
__global__ void __launch_bounds__(256) k_likelihood(const CellCounts* __restrict__ counts,
                                                    double* __restrict__ lik, double* __restrict__ fac,
                                                    const int2* __restrict__ list, const Stats* __restrict__ st,
                                                    SelfList sl, Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = KH ? KH : g.khalf;
    const int tw = KH ? ((kTileW + 2 * KH + 3) & ~3) : kTileW + 2 * k;  // KH: rows padded to 16 bytes
    const int th = kTileH + 2 * k;
    double* s_h = reinterpret_cast<double*>(smem_raw);               // th * kTileW
    float* s_t = reinterpret_cast<float*>(s_h + th * kTileW);        // th * tw
    int* s_pref = reinterpret_cast<int*>(s_t + th * tw);             // self-list launches only: nwords + 1 (the host
                                                                     // adds the bytes; keeps 3 CTAs/SM otherwise)
    const int tid = threadIdx.x;
    const int num_tiles = sl.bitmap ? self_list_build(sl, s_pref) : st->num_tiles;
    const size_t cells = (size_t)g.W * g.H;
    double kr[2 * KH + 1];
#pragma unroll
    for (int i = 0; i < 2 * KH + 1; i++) kr[i] = g.kernel[i];
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int2 item = sl.bitmap ? self_list_item(sl, s_pref, t) : list[t];
        const int ox = (item.y % g.tiles_x) * kTileW, oy = (item.y / g.tiles_x) * kTileH;
        const CellCounts* cmap = counts + (size_t)item.x * cells;
        double* out = lik + (size_t)item.x * cells;
        // 1. threshold the log-odds against logOdds(0.5) == 0.0 (GridMap.java:238-245); cells outside
        //    the map contribute 0 — Java skips those taps, and total + k*0.0 == total.
        const int twu = kTileW + 2 * k;  // used columns
        if (KH) {
            // all global loads of the tile + halo are issued before the first one is consumed
            constexpr int kPer = ((kTileH + 2 * KH) * (kTileW + 2 * KH) + 255) / 256;
            CellCounts c[kPer];
            int where[kPer];  // smem offset, or -1 outside the tile list / -2 outside the map
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                const int e = tid + j * 256;
                where[j] = -1;
                c[j] = CellCounts{0u, 0u};
                if (e < th * twu) {
                    const int ly = e / twu, lx = e - ly * twu;
                    const int gx = ox + lx - k, gy = oy + ly - k;
                    where[j] = ly * tw + lx;
                    if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H) c[j] = cmap[(size_t)gx + (size_t)gy * g.W];
                    else where[j] = -2 - where[j];
                }
            }
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                if (where[j] >= 0) s_t[where[j]] = 0.5f * (float)cell_code_fast(c[j], g);
                else if (where[j] <= -2) s_t[-2 - where[j]] = 0.0f;
            }
        } else {
            for (int e = tid; e < th * twu; e += 256) {
                const int ly = e / twu, lx = e - ly * twu;
                const int gx = ox + lx - k, gy = oy + ly - k;
                float code = 0.0f;
                if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H)
                    code = 0.5f * (float)cell_code_fast(cmap[(size_t)gx + (size_t)gy * g.W], g);
                s_t[ly * tw + lx] = code;
            }
        }
        __syncthreads();
        if (KH) {
            // 2. horizontal pass (Util.java:387-403): item = (row, 8-column segment)
            for (int it = tid; it < th * (kTileW / 8); it += 256) {
                const int ly = it / (kTileW / 8), x0 = (it - ly * (kTileW / 8)) * 8;
                const float4* row4 = reinterpret_cast<const float4*>(s_t + ly * tw + x0);
                double c[16];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 v = row4[q];
                    c[4 * q] = (double)v.x; c[4 * q + 1] = (double)v.y; c[4 * q + 2] = (double)v.z; c[4 * q + 3] = (double)v.w;
                }
                double* dst = s_h + ly * kTileW + x0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    double total = 0.0;
#pragma unroll
                    for (int i = 0; i < 2 * KH + 1; i++) total += kr[i] * c[j + i];
                    dst[j] = total;
                }
            }
            __syncthreads();
            // 3. vertical pass (Util.java:409-424): thread = (column, group of 8 rows)
            {
                const int lx = tid & (kTileW - 1), y0 = (tid / kTileW) * 8;
                const int gx = ox + lx;
                double hcol[8 + 2 * KH];
#pragma unroll
                for (int i = 0; i < 8 + 2 * KH; i++) hcol[i] = s_h[(y0 + i) * kTileW + lx];
                if (gx < g.W) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int gy = oy + y0 + j;
                        double total = 0.0;
#pragma unroll
                        for (int i = 0; i < 2 * KH + 1; i++) total += kr[i] * hcol[j + i];
                        if (gy < g.H) {
                            out[(size_t)gx + (size_t)gy * g.W] = total;
                            // shared map: the per-lookup factor of GridMap.probabilityOf (GridMap.java:284-288)
                            // is a pure function of the cell: evaluated once per cell here, not per lookup
                            if (fac) fac[(size_t)gx + (size_t)gy * g.fac_pitch] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
                        }
                    }
                }
            }
        } else {
            for (int e = tid; e < th * kTileW; e += 256) {
                const int ly = e / kTileW, lx = e - ly * kTileW;
                const float* row = s_t + ly * tw + lx;
                double total = 0.0;
                for (int i = 0; i < g.ktaps; i++) total += g.kernel[i] * (double)row[i];
                s_h[e] = total;
            }
            __syncthreads();
            for (int e = tid; e < kTileH * kTileW; e += 256) {
                const int ly = e / kTileW, lx = e - ly * kTileW;
                const int gx = ox + lx, gy = oy + ly;
                if (gx < g.W && gy < g.H) {
                    const double* col = s_h + ly * kTileW + lx;
                    double total = 0.0;
                    for (int i = 0; i < g.ktaps; i++) total += g.kernel[i] * col[i * kTileW];
                    out[(size_t)gx + (size_t)gy * g.W] = total;
                    if (fac) fac[(size_t)gx + (size_t)gy * g.fac_pitch] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
                }
            }
        }
        __syncthreads();
    }
}

// TMA variant of k_likelihood<3> (KH = 3, even W): the tile + halo of the counter map is fetched by ONE
// cp.async.bulk.tensor.3d request ({x, y, slot} box of the 3-D tensor counts[S][H][W], 8-byte elements;
// out-of-map coordinates are zero-filled by the TMA unit) that completes on an mbarrier, instead of 2660
// per-thread loads with their address arithmetic.  The innermost start coordinate must be 16-byte aligned
// (measured: an odd cell offset raises "illegal instruction"), so the box starts at x0 - 4 and is 72 cells
// wide; the 3-cell halo is columns 1..70 of it.  Passes 2 and 3 are those of k_likelihood<3>.
constexpr int kTmaTileW = kTileW + 8, kTmaTileH = kTileH + 6, kTmaPadX = 4;
__global__ void __launch_bounds__(256) k_likelihood_tma(const __grid_constant__ CUtensorMap tmap,
                                                        double* __restrict__ lik, double* __restrict__ fac,
                                                        const int2* __restrict__ list, const Stats* __restrict__ st,
                                                        SelfList sl, Geometry g) {
    constexpr int KH = 3;
    constexpr int tw = (kTileW + 2 * KH + 3) & ~3, th = kTileH + 2 * KH;
    extern __shared__ __align__(128) unsigned char smem_tma[];
    __shared__ __align__(8) uint64_t s_bar[2];
    constexpr int kRawBytes = (kTmaTileW * kTmaTileH * 8 + 127) & ~127;
    // two raw buffers: the TMA request of the NEXT tile is in flight while this tile is blurred
    double* s_h = reinterpret_cast<double*>(smem_tma + 2 * kRawBytes);  // th * kTileW
    float* s_t = reinterpret_cast<float*>(s_h + th * kTileW);           // th * tw
    int* s_pref = reinterpret_cast<int*>(s_t + th * tw);                // self-list launches only (host adds the bytes)
    const int tid = threadIdx.x;
    const int num_tiles = sl.bitmap ? self_list_build(sl, s_pref) : st->num_tiles;
    const size_t cells = (size_t)g.W * g.H;
    auto item_of = [&](int t) -> int2 { return sl.bitmap ? self_list_item(sl, s_pref, t) : list[t]; };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double kr[2 * KH + 1];
#pragma unroll
    for (int i = 0; i < 2 * KH + 1; i++) kr[i] = g.kernel[i];
    auto fetch = [&](int t, int buf) {  // one elected thread: arm the barrier, issue the 3-D box load
        const int2 item = item_of(t);
        const int ox = (item.y % g.tiles_x) * kTileW, oy = (item.y / g.tiles_x) * kTileH;
        const uint32_t bar = smem_u32(&s_bar[buf]);
        const uint32_t bytes = kTmaTileW * kTmaTileH * 8;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the buffer are done
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(smem_tma + buf * kRawBytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(ox - kTmaPadX),
              "r"(oy - KH), "r"(item.x), "r"(bar)
            : "memory");
    };
    if (tid == 0 && (int)blockIdx.x < num_tiles) fetch(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, it++) {
        const int buf = it & 1;
        const int2 item = item_of(t);
        const int ox = (item.y % g.tiles_x) * kTileW, oy = (item.y / g.tiles_x) * kTileH;
        double* out = lik + (size_t)item.x * cells;
        if (tid == 0 && t + (int)gridDim.x < num_tiles) fetch(t + gridDim.x, buf ^ 1);
        const CellCounts* s_c = reinterpret_cast<const CellCounts*>(smem_tma + buf * kRawBytes);
        const uint32_t bar = smem_u32(&s_bar[buf]);
        const uint32_t phase = (uint32_t)(it >> 1) & 1u;
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar), "r"(phase)
                : "memory");
        }
        // 1. threshold (GridMap.java:238-245); cells outside the map contribute 0, not the "unknown" 0.5 that
        //    the zero-filled counters would threshold to
        constexpr int twu = kTileW + 2 * KH;
        for (int e = tid; e < th * twu; e += 256) {
            const int ly = e / twu, lx = e - ly * twu;
            const int gx = ox + lx - KH, gy = oy + ly - KH;
            float code = 0.0f;
            if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H)
                code = 0.5f * (float)cell_code_fast(s_c[ly * kTmaTileW + lx + (kTmaPadX - KH)], g);
            s_t[ly * tw + lx] = code;
        }
        __syncthreads();
        // 2. horizontal pass (Util.java:387-403): item = (row, 8-column segment)
        for (int seg = tid; seg < th * (kTileW / 8); seg += 256) {
            const int ly = seg / (kTileW / 8), x0 = (seg - ly * (kTileW / 8)) * 8;
            const float4* row4 = reinterpret_cast<const float4*>(s_t + ly * tw + x0);
            double c[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 v = row4[q];
                c[4 * q] = (double)v.x; c[4 * q + 1] = (double)v.y; c[4 * q + 2] = (double)v.z; c[4 * q + 3] = (double)v.w;
            }
            double* dst = s_h + ly * kTileW + x0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double total = 0.0;
#pragma unroll
                for (int i = 0; i < 2 * KH + 1; i++) total += kr[i] * c[j + i];
                dst[j] = total;
            }
        }
        __syncthreads();
        // 3. vertical pass (Util.java:409-424): thread = (column, group of 8 rows)
        {
            const int lx = tid & (kTileW - 1), y0 = (tid / kTileW) * 8;
            const int gx = ox + lx;
            double hcol[8 + 2 * KH];
#pragma unroll
            for (int i = 0; i < 8 + 2 * KH; i++) hcol[i] = s_h[(y0 + i) * kTileW + lx];
            if (gx < g.W) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int gy = oy + y0 + j;
                    double total = 0.0;
#pragma unroll
                    for (int i = 0; i < 2 * KH + 1; i++) total += kr[i] * hcol[j + i];
                    if (gy < g.H) {
                        out[(size_t)gx + (size_t)gy * g.W] = total;
                        if (fac) fac[(size_t)gx + (size_t)gy * g.fac_pitch] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// A5 — GridMap.probabilityOf GridMap.java:261-294.  One warp per particle, hit beams across lanes.
// The beam table is staged once per CTA with a 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) and
// completion is signalled on an mbarrier; persistent CTAs then stride over particles.
// Each lane multiplies its factors (<= ceil(B/32) of them, each in [0.01, 0.91]; the exponent is peeled off
// every 64 factors so the product never underflows), takes one log, and the 32 logs are summed with a fixed
// xor-shuffle tree (deterministic).
// ------------------------------------------------------------------------------------------------

// move the binary exponent of a positive normal double into `exp2`: exact (power-of-two scaling)
__device__ __forceinline__ void peel_exponent(double& mant, int& exp2) {
    const int hi = __double2hiint(mant);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    if (e == -1023) return;  // zero / subnormal (only reachable with z_hit == 1 configurations): ln() handles it
    mant = __hiloint2double(hi - (e << 20), __double2loint(mant));
    exp2 += e;
}

__global__ void __launch_bounds__(256) k_score(const float4* __restrict__ pose, int lo, int cnt,
                                               const double2* __restrict__ hit_xy, const int* __restrict__ num_hit,
                                               const double* __restrict__ lik, const int* __restrict__ slot,
                                               double* __restrict__ lw, ExchangeRec* __restrict__ xlocal,
                                               Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    double2* s_xy = reinterpret_cast<double2*>(smem_raw);
    const int nh = *num_hit;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t bar = smem_u32(&s_bar);
    if (nh > 0) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)nh * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(s_xy)),
                "l"(hit_xy), "r"(bytes), "r"(bar)
                : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
    const size_t cells = (size_t)g.W * g.H;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int li = blockIdx.x * (blockDim.x >> 5) + wid; li < cnt; li += warps) {
        const int i = lo + li;
        const float4 p = pose[i];
        const Xform t(p.x, p.y, p.z);
        const double* field = lik + (slot ? (size_t)slot[li] * cells : 0);
        double prod = 1.0;
        int exp2 = 0, it = 0;
        for (int b0 = 0; b0 < nh; b0 += 128, it++) {
            double f[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int b = b0 + u * 32 + lane;
                f[u] = 1.0;
                if (b < nh) {
                    const double2 m = s_xy[b];
                    const int gx = cell_of(t.tx(m.x, m.y) - g.posx, g.res, g.inv_res);  // (int): toward zero
                    const int gy = cell_of(t.ty(m.x, m.y) - g.posy, g.res, g.inv_res);
                    if (!(gx < 0 || gy < 0 || gx >= g.W || gy >= g.H)) {
                        const double val = __ldg(field + ((size_t)gx + (size_t)gy * g.W));
                        f[u] = val == 0.5 ? g.uniform_term : g.z_hit * val + g.random_term;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) prod *= f[u];
            // factors are in [0.01, 0.91]: 64 of them cannot underflow a normalised mantissa, more can (a lane
            // takes up to GMS_MAX_BEAMS / 32 = 400) — peel the exponent off exactly every 16 rounds
            if ((it & 15) == 15) peel_exponent(prod, exp2);
        }
        peel_exponent(prod, exp2);
        const double l = warp_sum(log(prod) + (double)exp2 * 0.6931471805599453);
        if (lane == 0) {
            lw[i] = l;
            if (xlocal) {
                ExchangeRec r;
                r.lw = l; r.x = p.x; r.y = p.y; r.t = p.z; r.pad = 0;
                xlocal[li] = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Per-particle maps: the likelihood field evaluated WHERE IT IS READ.
// GridMap.computeLikelihoodMap (GridMap.java:233-250) rebuilds likelihoodData[W*H] of every particle before the
// scan is scored, and GridMap.probabilityOf (GridMap.java:261-294) then reads it at B_hit cells.  The field at a
// cell is a pure function of the thresholded codes within khalf cells (separable blur, Util.java:378-426), so the
// step evaluates it only at the cells the scan looks up: (2*khalf+1)^2 counter pairs per lookup, 7 x 7 for the
// reference's 0.05 m cells, straight from the counter map — O(B * k^2) per particle and step instead of
// O(touched tiles * k), and no per-particle field in HBM at all (the reference spends ~3/4 of its step in that
// rebuild; the tile kernel took 0.51 of the 2.24 ms K2pp step and the field doubled every map copy).
// Same f64 operation order as k_likelihood (tap index ascending, multiply then add, no FMA; rows and columns
// outside the map contribute 0.0) => the value is bit-identical to the stored field's.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double code_of(const CellCounts c, const Geometry& g) {
    return 0.5 * (double)cell_code_fast(c, g);  // {0, .5, 1}: what k_likelihood keeps as f32 in shared memory
}
// KH = 3, even W (rows and slots start 16-byte aligned): the 7 cells of a row are covered by four aligned pairs
template <int KH>
__device__ __forceinline__ double field_at(const CellCounts* __restrict__ cmap, int gx, int gy, const Geometry& g,
                                           const double* __restrict__ kr) {
    double total = 0.0;
    if constexpr (KH == 3) {
        const int xs = gx - 3, xa = xs & ~1, o = xs - xa;  // first tap column, its aligned pair, 0 / 1
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const int y = gy + r - 3;
            double h = 0.0;
            if (y >= 0 && y < g.H) {
                const uint4* row = reinterpret_cast<const uint4*>(cmap + (size_t)y * g.W);
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x = xa + 2 * q;
                    v[q] = (x >= 0 && x < g.W) ? __ldg(row + (x >> 1)) : make_uint4(0u, 0u, 0u, 0u);
                }
                double c[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x = xa + 2 * q;
                    const bool in = x >= 0 && x < g.W;
                    c[2 * q] = in ? code_of(CellCounts{v[q].x, v[q].y}, g) : 0.0;
                    c[2 * q + 1] = in ? code_of(CellCounts{v[q].z, v[q].w}, g) : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 7; i++) h += kr[i] * (o ? c[i + 1] : c[i]);
            }
            total += kr[r] * h;
        }
    } else {
        const int k = g.khalf;
        for (int r = 0; r < g.ktaps; r++) {
            const int y = gy + r - k;
            double h = 0.0;
            if (y >= 0 && y < g.H) {
                const CellCounts* row = cmap + (size_t)y * g.W;
                for (int i = 0; i < g.ktaps; i++) {
                    const int x = gx + i - k;
                    const double c = (x >= 0 && x < g.W) ? code_of(row[x], g) : 0.0;
                    h += g.kernel[i] * c;
                }
            }
            total += g.kernel[r] * h;
        }
    }
    return total;
}

// A5 for per-particle maps: one CTA of kPpWarps warps per particle, hit beams across the CTA's threads (beam b goes
// to thread b mod 128: a thread multiplies its factors in beam order, takes one log, each warp sums its 32 logs with
// the fixed xor tree and thread 0 adds the kPpWarps partial sums in warp order) — one fixed reduction shape for every
// particle count, so log-weights are bit-identical run to run and for every rank count.  Four warps per particle
// because a lookup is a burst of 28 independent 16-byte loads followed by ~110 f64 operations: with one warp per
// particle 1000 particles left the SMs 12 % occupied and the kernel latency-bound (ncu, r02j: 126 us).
template <int KH, int kPpWarps>
__global__ void __launch_bounds__(kPpWarps * 32) k_score_pp(const float4* __restrict__ pose, int lo, int cnt,
                                                           const double2* __restrict__ hit_xy,
                                                           const int* __restrict__ num_hit,
                                                           const CellCounts* __restrict__ counts,
                                                           const int* __restrict__ slot, double* __restrict__ lw,
                                                           ExchangeRec* __restrict__ xlocal, Geometry g) {
    __shared__ double s_part[kPpWarps];
    const int li = blockIdx.x;
    if (li >= cnt) return;
    const int nh = *num_hit;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double kr[2 * KH + 1];
#pragma unroll
    for (int i = 0; i < 2 * KH + 1; i++) kr[i] = g.kernel[i];
    const int i = lo + li;
    const float4 p = pose[i];
    const Xform t(p.x, p.y, p.z);
    const CellCounts* cmap = counts + (size_t)slot[li] * ((size_t)g.W * g.H);
    double prod = 1.0;
    int exp2 = 0, it = 0;
    for (int b = tid; b < nh; b += kPpWarps * 32, it++) {
        const double2 m = __ldg(hit_xy + b);
        const int gx = cell_of(t.tx(m.x, m.y) - g.posx, g.res, g.inv_res);  // (int): toward zero
        const int gy = cell_of(t.ty(m.x, m.y) - g.posy, g.res, g.inv_res);
        if (!(gx < 0 || gy < 0 || gx >= g.W || gy >= g.H)) {
            const double val = field_at<KH>(cmap, gx, gy, g, kr);
            prod *= val == 0.5 ? g.uniform_term : g.z_hit * val + g.random_term;
        }
        if ((it & 63) == 63) peel_exponent(prod, exp2);  // factors in [0.01, 0.91]: no underflow between two peels
    }
    peel_exponent(prod, exp2);
    const double l = warp_sum(log(prod) + (double)exp2 * 0.6931471805599453);
    if (lane == 0) s_part[wid] = l;
    __syncthreads();
    if (tid == 0) {
        double tot = s_part[0];
#pragma unroll
        for (int w = 1; w < kPpWarps; w++) tot += s_part[w];
        lw[i] = tot;
        if (xlocal) {
            ExchangeRec r;
            r.lw = tot; r.x = p.x; r.y = p.y; r.t = p.z; r.pad = 0;
            xlocal[li] = r;
        }
    }
}

// Shared map, many particles: ONE THREAD per particle, the 32 lanes of a warp take 32 particles that
// are neighbours in heading (k_motion / k_sort_*).  Why (ncu, profiles/r01_*): with beams across lanes a
// warp-wide gather touches ~17-25 different 128-byte lines, and the L1 tag stage retires one line per
// clock, so scoring sat at ~1 lookup/clk/SM (0.25 ms for 72 M lookups) regardless of instruction
// count.  Heading-sorted neighbours look up the SAME beam at nearly the same cell, so a warp-wide
// gather touches 1-3 lines.  The product runs over the beams in Java's order; every 64 factors the
// exponent is peeled off exactly (power-of-two scaling commutes with rounding), so mant * 2^exp2 is
// Java's product bit for bit wherever that does not underflow, and ln() is taken once.
// G threads share a particle (beam u*G + gsub goes to sub-thread gsub): G = 1 for ~1e5 particles, up to 32
// (= one warp per particle) for small sets, so the grid always fills the machine.
// V selects how the fast path validates its fixed-point cell index (same accepted set, same results):
// 0 = mask / subtract / compare per coordinate (measured, default); 1 = ALU-lean form — the margin is a power
// of two, so "fraction within margin of an integer" is ((i + margin) & M) == 0, x and y share one unsigned min
// and one LOP3 covers both high-word tests (GMS_SCORE_V=1; the ALU pipe is this kernel's busiest).
// V = 3 / 4: the code of V = 2 / 0 compiled for 8 resident CTAs per SM (<= 64 registers) instead of 5-6 (measured
// slower: the spills land in the beam loop).  V = 5: V = 2 with a software-pipelined beam loop.
template <int G, int V = 0>
__global__ void __launch_bounds__(128, (V == 3 || V == 4) ? 8 : 1) k_score_sorted(const float4* __restrict__ pose, int lo, int cnt,
                                                      const double2* __restrict__ hit_xy,
                                                      const int* __restrict__ num_hit, const double* __restrict__ fac,
                                                      const int* __restrict__ order, double* __restrict__ lw,
                                                      ExchangeRec* __restrict__ xlocal,
                                                      const double* __restrict__ rmax2, unsigned* __restrict__ work,
                                                      Geometry g) {
    constexpr int K = (V == 3 || V >= 5) ? 2 : (V == 4 ? 0 : V);  // which index-validation code
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    double2* s_xy = reinterpret_cast<double2*>(smem_raw);
    const int nh = *num_hit;
    const int tid = threadIdx.x;
    const uint32_t bar = smem_u32(&s_bar);
    if (nh > 0) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)nh * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(s_xy)),
                "l"(hit_xy), "r"(bytes), "r"(bar)
                : "memory");
        }
    }
    // Work item = one warp's worth of (particle, sub-thread) pairs.  `work` == null: one item per warp, the grid covers
    // them all (small sets).  Otherwise the grid is exactly the resident CTAs and every warp draws items from the
    // counter until none are left: at 100k particles the static grid was 1563 CTAs on 592 resident ones = 2.64 waves,
    // i.e. a third wave 64 % full (12 % of the kernel's time idle); drawn items keep every warp busy to the end.
    // Which warp scores a particle does not enter its result.
    // A CTA draws its warps' items TOGETHER, as one run of consecutive items: neighbours in the heading order read the
    // same lines of the factor field, and the static grid's L1 hit rate (82 %) comes from the four warps of a CTA
    // working on four consecutive items at the same time.  (Drawn warp by warp — the first version of this mode —
    // consecutive items scatter over the SMs: 0.110 ms against the static grid's 0.097.)
    const int lane = tid & 31, wid = tid >> 5, nw = (int)(blockDim.x >> 5);
    const int nitems = (int)(((long long)cnt * G + 31) / 32);
    __shared__ unsigned s_next;
    auto next_item = [&]() -> int {
        __syncthreads();  // every warp has taken its item of the previous draw
        if (tid == 0) s_next = atomicAdd(work, (unsigned)nw);
        __syncthreads();
        return (int)s_next + wid;
    };
    int item = work ? next_item() : (int)(blockIdx.x * nw + wid);
    if (nh > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
  // (the condition is uniform over the CTA — next_item() holds barriers; a warp whose item lies beyond the last one
  // runs the body with idle lanes: li = -1 below)
  for (; item - wid < nitems; item = work ? next_item() : nitems + wid) {
    const int t = (item * 32 + lane) / G, gsub = (item * 32 + lane) % G;
    const int li = t < cnt ? (order ? order[t] : t) : -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (li >= 0) p = pose[lo + li];
    const Xform x(p.x, p.y, p.z);
    const int nhe = li >= 0 ? nh : 0;  // idle sub-threads still take part in the shuffles below
    // Fast path for (int) ((world - position) / resolution), GridMap.java:273-274: q~ = the same quantity
    // evaluated with two FMAs from per-particle constants.  |q~ - q_java| < 1e-9 for every finite input
    // with |q| < 2^31, so when q~ is further than 1e-5 from both neighbouring integers the truncation is
    // the one Java computes; otherwise (or NaN / saturation) the literal expression with the f64 division
    // decides.  Result: bit-identical cell indices at 2 DFMA instead of 4 DMUL/DADD + subtract + divide.
    const double cinv = x.c * g.inv_res, sinv = x.s * g.inv_res;
    const double pqx = (x.px - g.posx) * g.inv_res, pqy = (x.py - g.posy) * g.inv_res;
    double mant = 1.0;
    int exp2 = 0;
    const unsigned uW = (unsigned)g.W, uH = (unsigned)g.H, pitch = (unsigned)g.fac_pitch;
    const unsigned oob = pitch * pitch;  // sentinel element (1.0) behind the padded square
    const int fk = g.fx_k;
    const unsigned fmask = (1u << fk) - 1u, fmarg = (unsigned)g.fx_margin, fspan = (1u << fk) - 2u * fmarg;
    const unsigned fnear = fmask & ~(2u * fmarg - 1u);  // V == 1: the fraction bits above the 2*margin window
    // exact (slow) evaluation of one beam: the literal Java expression incl. the f64 division
    auto factor_exact = [&](const double2 m) -> double {
        {   // far outside the map (more than a cell beyond an edge): no lookup, no division needed
            const double qx = fma(m.x, cinv, fma(-m.y, sinv, pqx));
            const double qy = fma(m.x, sinv, fma(m.y, cinv, pqy));
            if (qx < -2.0 || qy < -2.0 || qx > (double)uW + 1.0 || qy > (double)uH + 1.0) return 1.0;
        }
        const int gx = java_d2i((x.tx(m.x, m.y) - g.posx) / g.res);
        const int gy = java_d2i((x.ty(m.x, m.y) - g.posy) / g.res);
        if ((unsigned)gx < uW && (unsigned)gy < uH) return __ldg(fac + ((unsigned)gy * pitch + (unsigned)gx));
        return 1.0;
    };
    // V == 2 only.  The magic constant is folded into the per-particle offsets (q~ + magic comes out of the two
    // FMAs directly: 5 instead of 7 FP64-pipe instructions per lookup); that adds at most one more unit of
    // 2^-k of rounding, far inside the acceptance margin of >= 8 units (tests/test_fastpath_bound.py).
    // The high-word test of the other variants (|q| within the fixed-point range) is hoisted out of the beam
    // loop: with R = the longest hit beam of the scan in cells, every q of this particle lies within
    // |pq| + R * 1.000001 + 2, so one per-particle comparison covers all beams; a particle that fails it
    // (poses thousands of cells outside the map) takes the exact path for every beam.
    const double pqxm = pqx + g.fx_magic, pqym = pqy + g.fx_magic;
    bool fast_ok = true;
    if constexpr (K == 2) {
        const double R = sqrt(*rmax2) * g.inv_res * 1.000001 + 2.0;
        const double lim = (double)(1u << (31 - fk)) - 2.0;
        fast_ok = fabs(pqx) + R < lim && fabs(pqy) + R < lim;  // NaN poses fail too
    }
    const unsigned fshift = 32u - (unsigned)fk;
    const unsigned fmul = 1u << fshift, fadd = fmarg << fshift, fthr = (2u * fmarg) << fshift;
    const unsigned bound = 1u << (fk + g.fac_lp);  // (ix | iy) below it  <=>  both cells inside the padded square
    auto peel = [&]() { peel_exponent(mant, exp2); };
    int b0 = 0, it = 0;
    const int nfast = (K == 2 && !fast_ok) ? 0 : nhe;
    if constexpr (V == 5 || V == 6) {
        constexpr int NB = V == 6 ? 4 : 8;  // beams per batch
        // Software-pipelined form of the K == 2 loop: the eight gathers of batch i+1 are issued BEFORE the products
        // of batch i are formed, so a warp always has a batch of loads in flight (ncu on the plain loop: 22 %
        // occupancy at 98 registers, long_scoreboard + wait the top stalls; the loop body is index math ->
        // 8 loads -> wait -> multiply tree, serialised per warp).
        auto index8 = [&](int base, unsigned (&idx)[NB]) -> unsigned {
            unsigned bad = 0;
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const double2 m = s_xy[base + u * G + gsub];
                const double tx = fma(m.x, cinv, fma(-m.y, sinv, pqxm));
                const double ty = fma(m.x, sinv, fma(m.y, cinv, pqym));
                const unsigned ix = (unsigned)__double2loint(tx), iy = (unsigned)__double2loint(ty);
                const unsigned ux = ix * fmul + fadd, uy = iy * fmul + fadd;
                const bool ok = min(ux, uy) >= fthr && (ix | iy) < bound;
                bad |= ok ? 0u : (1u << u);
                idx[u] = ok ? __umulhi(iy, fmul) * pitch + __umulhi(ix, fmul) : oob;
            }
            return bad;
        };
        double f[NB];
        unsigned bad = 0;
        if (NB * G <= nfast) {
            unsigned idx[NB];
            bad = index8(0, idx);
#pragma unroll
            for (int u = 0; u < NB; u++) f[u] = __ldg(fac + idx[u]);
        }
#pragma unroll 2
        for (; b0 + NB * G <= nfast; b0 += NB * G, it++) {
            double fn[NB];
            unsigned badn = 0;
            const bool more = b0 + 2 * NB * G <= nfast;
            if (more) {
                unsigned idx[NB];
                badn = index8(b0 + NB * G, idx);
#pragma unroll
                for (int u = 0; u < NB; u++) fn[u] = __ldg(fac + idx[u]);
            }
            if (bad) {
#pragma unroll
                for (int u = 0; u < NB; u++)
                    if (bad & (1u << u)) f[u] = factor_exact(s_xy[b0 + u * G + gsub]);
            }
            if constexpr (NB == 8) mant *= ((f[0] * f[1]) * (f[2] * f[3])) * ((f[4] * f[5]) * (f[6] * f[7]));
            else mant *= (f[0] * f[1]) * (f[2] * f[3]);
            if ((it & (64 / NB - 1)) == 64 / NB - 1) peel();
            if (more) {
#pragma unroll
                for (int u = 0; u < NB; u++) f[u] = fn[u];
                bad = badn;
            }
        }
    } else if constexpr (V == 7) {
        // (Compiled for 5 / 6 resident CTAs per SM — 96 / 80 registers, no / 20 bytes of spills — this loop measured
        // 0.1005 / 0.0976 ms against 0.0976 ms at 4 CTAs: occupancy is not what limits it.)
        // V = 2 with the validation amortised over the batch: instead of a compare / select / flag per lookup (the
        // ALU pipe is this kernel's busiest: 43 % against 23 % FP64 and 13 % FMA, ncu r02), the eight lookups share
        // one running unsigned minimum of the shifted fractions and one running OR of the fixed-point coordinates —
        // a 3-input min and a 3-input LOP per lookup — and are judged once.  The index is masked into the padded
        // square, so a coordinate outside it (judged afterwards) still loads from inside the array.  A batch that
        // fails (probability ~2e-3: one of 16 coordinates within the margin of a cell border) is redone lookup by
        // lookup with V = 2's test: same accepted set per lookup, same factors.
        const unsigned idxmask = pitch * pitch - 1u;
        for (; b0 + 8 * G <= nfast; b0 += 8 * G, it++) {
            unsigned idx[8];
            unsigned m8 = 0xffffffffu, o8 = 0u;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double2 m = s_xy[b0 + u * G + gsub];
                const double tx = fma(m.x, cinv, fma(-m.y, sinv, pqxm));
                const double ty = fma(m.x, sinv, fma(m.y, cinv, pqym));
                const unsigned ix = (unsigned)__double2loint(tx), iy = (unsigned)__double2loint(ty);
                m8 = min(m8, min(ix * fmul + fadd, iy * fmul + fadd));
                o8 |= ix | iy;
                idx[u] = (__umulhi(iy, fmul) * pitch + __umulhi(ix, fmul)) & idxmask;
            }
            double f[8];
#pragma unroll
            for (int u = 0; u < 8; u++) f[u] = __ldg(fac + idx[u]);
            if (!(m8 >= fthr && o8 < bound)) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const double2 m = s_xy[b0 + u * G + gsub];
                    const double tx = fma(m.x, cinv, fma(-m.y, sinv, pqxm));
                    const double ty = fma(m.x, sinv, fma(m.y, cinv, pqym));
                    const unsigned ix = (unsigned)__double2loint(tx), iy = (unsigned)__double2loint(ty);
                    const bool ok = min(ix * fmul + fadd, iy * fmul + fadd) >= fthr && (ix | iy) < bound;
                    if (!ok) f[u] = factor_exact(m);
                }
            }
            mant *= ((f[0] * f[1]) * (f[2] * f[3])) * ((f[4] * f[5]) * (f[6] * f[7]));
            if ((it & 7) == 7) peel();
        }
    } else
    for (; b0 + 8 * G <= nfast; b0 += 8 * G, it++) {
        unsigned idx[8];
        unsigned bad = 0;
        // branch-free fast path for 8 beams (their dependency chains interleave); beams whose q~ is too
        // close to an integer are only flagged here and redone exactly below.  End points outside the
        // map read the sentinel fac[W*H] == 1.0 (GridMap.java:276: such beams do not multiply).
        if constexpr (K == 2) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double2 m = s_xy[b0 + u * G + gsub];
                const double tx = fma(m.x, cinv, fma(-m.y, sinv, pqxm));
                const double ty = fma(m.x, sinv, fma(m.y, cinv, pqym));
                const unsigned ix = (unsigned)__double2loint(tx), iy = (unsigned)__double2loint(ty);
                // fraction test on the FMA pipe: (i + margin) * 2^(32-k) keeps only the fraction bits, shifted to
                // the top; it is >= 2*margin * 2^(32-k)  <=>  the fraction lies in [margin, 2^k - margin)
                const unsigned ux = ix * fmul + fadd, uy = iy * fmul + fadd;
                const bool ok = min(ux, uy) >= fthr && (ix | iy) < bound;
                bad |= ok ? 0u : (1u << u);
                idx[u] = ok ? __umulhi(iy, fmul) * pitch + __umulhi(ix, fmul) : oob;  // cell = i >> k
            }
        } else
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double2 m = s_xy[b0 + u * G + gsub];  // G distinct addresses per warp: shared-memory broadcast
            const double qx = fma(m.x, cinv, fma(-m.y, sinv, pqx));
            const double qy = fma(m.x, sinv, fma(m.y, cinv, pqy));
            // q + 1.5*2^(52-k): the low mantissa word now holds round(q * 2^k) as an integer, the high word
            // is a known constant exactly when 0 <= q < 2^(31-k).  Cell = I >> k, fraction = I & (2^k - 1).
            const double tx = qx + g.fx_magic, ty = qy + g.fx_magic;
            const int ix = __double2loint(tx), iy = __double2loint(ty);
            const unsigned gx = (unsigned)ix >> fk, gy = (unsigned)iy >> fk;
            bool ok;
            if constexpr (K == 1) {
                // fmarg is a power of two (Geometry): frac in [fmarg, 2^k - fmarg)  <=>  ((i + fmarg) & fnear) != 0
                const unsigned hi = ((unsigned)__double2hiint(tx) ^ (unsigned)g.fx_hi) |
                                    ((unsigned)__double2hiint(ty) ^ (unsigned)g.fx_hi);
                const unsigned nx = ((unsigned)ix + fmarg) & fnear, ny = ((unsigned)iy + fmarg) & fnear;
                ok = hi == 0u && min(nx, ny) != 0u;
            } else {
                const bool inrange = __double2hiint(tx) == g.fx_hi && __double2hiint(ty) == g.fx_hi;
                const unsigned frx = (unsigned)ix & fmask, fry = (unsigned)iy & fmask;
                // accepted when both fractions are >= margin away from 0 and 1: no integer lies between q~ and
                // Java's quotient, so both truncate to the same cell
                ok = inrange && (frx - fmarg) < fspan && (fry - fmarg) < fspan;
            }
            const bool inb = gx < uW && gy < uH;
            bad |= ok ? 0u : (1u << u);
            idx[u] = (ok && inb) ? gy * pitch + gx : oob;
        }
        double f[8];
#pragma unroll
        for (int u = 0; u < 8; u++) f[u] = __ldg(fac + idx[u]);  // 8 independent gathers in flight
        if (bad) {
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (bad & (1u << u)) f[u] = factor_exact(s_xy[b0 + u * G + gsub]);
        }
        // four independent partial products (a serial chain of 8 dependent f64 multiplies stalls the warp);
        // the order differs from Java's left-to-right product by rounding only (|d ln w| <= 1e-13)
        mant *= ((f[0] * f[1]) * (f[2] * f[3])) * ((f[4] * f[5]) * (f[6] * f[7]));
        if ((it & 7) == 7) peel();  // factors are in [0.01, 0.91]: 64 of them cannot underflow a normalised mantissa
    }
    // beams left over after the last full batch (none at G = 2 with 720 beams; 208 of 720 at G = 32, i.e. one warp
    // per particle, K3): the same fast path one lookup at a time — round 2's first half sent them through the exact
    // division and K3's scoring took 45 us for 7.2 M lookups
    for (int b = b0 + gsub, j = 0; b < nhe; b += G, j++) {
        const double2 m = s_xy[b];
        double f;
        if (K == 2 && b < nfast) {
            const double tx = fma(m.x, cinv, fma(-m.y, sinv, pqxm));
            const double ty = fma(m.x, sinv, fma(m.y, cinv, pqym));
            const unsigned ix = (unsigned)__double2loint(tx), iy = (unsigned)__double2loint(ty);
            const bool ok = min(ix * fmul + fadd, iy * fmul + fadd) >= fthr && (ix | iy) < bound;
            f = ok ? __ldg(fac + (__umulhi(iy, fmul) * pitch + __umulhi(ix, fmul))) : factor_exact(m);
        } else {
            f = factor_exact(m);
        }
        mant *= f;
        if ((j & 31) == 31) peel();
    }
    peel();
    if (G > 1) {  // combine the sub-threads' partial products: mantissas in [1, 2), at most 2^5 after the tree
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            mant *= __shfl_xor_sync(0xffffffffu, mant, o);
            exp2 += __shfl_xor_sync(0xffffffffu, exp2, o);
        }
    }
    const bool writer = li >= 0 && gsub == 0;
    const double l = log(mant) + (double)exp2 * 0.6931471805599453;
    if (writer) {
        lw[lo + li] = l;
        if (xlocal) {
            ExchangeRec r;
            r.lw = l; r.x = p.x; r.y = p.y; r.t = p.z; r.pad = 0;
            xlocal[li] = r;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A8..A11 — GridMap.integrateObservation GridMap.java:173-191 + applyMeasurement :194-228.
// Per-particle maps: one thread per (particle, beam) ray.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_map_update(const float4* __restrict__ pose, int lo, int cnt,
                                                    const double2* __restrict__ all_xy,
                                                    const float* __restrict__ meas,
                                                    const uint8_t* __restrict__ hit, int B,
                                                    CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                    int4* __restrict__ rect, uint32_t* __restrict__ dirty, Geometry g) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)cnt * B) return;
    const int li = (int)(gid / B);
    const int b = (int)(gid - (long long)li * B);
    const int s = slot[li];
    const float4 p = pose[lo + li];
    const Xform t(p.x, p.y, p.z);
    const float sx = (float)((t.tx(0.0, 0.0) - g.posx) / g.res);
    const float sy = (float)((t.ty(0.0, 0.0) - g.posy) / g.res);
    const double2 m = all_xy[b];
    const float ex = (float)((t.tx(m.x, m.y) - g.posx) / g.res);
    const float ey = (float)((t.ty(m.x, m.y) - g.posy) / g.res);
    CellBox box;
    apply_measurement(counts + (size_t)s * ((size_t)g.W * g.H), dirty + (size_t)s * g.tile_words, g, sx, sy, ex, ey,
                      meas[b], hit[b] != 0, box);
    if (box.x1 >= 0) {
        int* r = reinterpret_cast<int*>(rect + s);
        atomicMin(r + 0, box.x0);
        atomicMin(r + 1, box.y0);
        atomicMax(r + 2, box.x1);
        atomicMax(r + 3, box.y1);
    }
}

// The same without dirty-tile bookkeeping (fire-and-forget RED): the step's kernel for per-particle maps now that
// their likelihood field is evaluated on demand.  NEG: subtract (see apply_measurement_red); `rect` may be null.
// `ulist` / `n_used` (nullable): integrate only the listed local particles — when a resampling follows the update,
// a particle without children is dropped together with its map (SLAM.java:133-153 builds the new list from copies
// of the selected particles), so integrating the scan into it is dead work; the resampling lists the parents it
// selected (k_list_parents) and only those are integrated.  A short list (a handful of parents is the usual
// outcome) leaves the machine almost empty and the kernel bound by the latency of the ray loop, whose REDs
// serialise over the distinct sectors of a warp's lanes: such launches put 8 rays instead of 32 into a warp.
constexpr int kSparseRays = 65536;  // up to this many rays: 8 per warp
template <bool NEG>
__global__ void __launch_bounds__(128) k_map_update_red(const float4* __restrict__ pose, int lo, int cnt,
                                                        const double2* __restrict__ all_xy,
                                                        const float* __restrict__ meas,
                                                        const uint8_t* __restrict__ hit, int B,
                                                        CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                        int4* __restrict__ rect, const int* __restrict__ ulist,
                                                        const int* __restrict__ n_used,
                                                        const Stats* __restrict__ st, Geometry g) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = gid;
    long long nrays = (long long)cnt * B;
    if (ulist) {
        if (st->xerror) return;
        nrays = (long long)(*n_used) * B;
        if (nrays <= kSparseRays) {
            if ((threadIdx.x & 31) >= 8) return;
            r = (gid >> 5) * 8 + (threadIdx.x & 31);
        }
    }
    if (r >= nrays) return;
    int li = (int)(r / B);
    const int b = (int)(r - (long long)li * B);
    if (ulist) li = ulist[li];
    const int s = slot[li];
    const float4 p = pose[lo + li];
    const Xform t(p.x, p.y, p.z);
    const float sx = (float)((t.tx(0.0, 0.0) - g.posx) / g.res);
    const float sy = (float)((t.ty(0.0, 0.0) - g.posy) / g.res);
    const double2 m = all_xy[b];
    const float ex = (float)((t.tx(m.x, m.y) - g.posx) / g.res);
    const float ey = (float)((t.ty(m.x, m.y) - g.posy) / g.res);
    CellBox box;
    apply_measurement_red<NEG>(counts + (size_t)s * ((size_t)g.W * g.H), g, sx, sy, ex, ey, meas[b], hit[b] != 0, box);
    if (rect && box.x1 >= 0) {
        int* r = reinterpret_cast<int*>(rect + s);
        atomicMin(r + 0, box.x0);
        atomicMin(r + 1, box.y0);
        atomicMax(r + 2, box.x1);
        atomicMax(r + 3, box.y1);
    }
}

// Shared map (one scan per step, only B rays): the DDA of a ray is inherently sequential (f32 error term,
// RayIterator.java:112-130), but the per-cell work (sqrt, inverse sensor model, counter update) is not.
// k_ray_integrate does both in ONE launch: a CTA owns 4 consecutive rays (180 CTAs for 720 beams); 4 lanes of warp 0
// walk them in lockstep and record the cells {x | y << 16} (cell k of ray b at ray_cells[k * Bpad + b]), then all 8
// warps classify and accumulate the CTA's (cell, ray) pairs in parallel, four independent loads per thread.
// The walk loop keeps everything in registers (grid size, increments) and steps four cells per iteration:
// the round-1 loop re-loaded W/H from the constant bank inside a predicate chain and took ~130 cycles per
// cell (24 us for 720 rays); this one is bounded by the f32 add -> compare dependency of the error term.
// `walk_only` (GMS_UPDATE_SORTED) stops after the walk: k_ray_keys + sort + k_apply_runs take over.
constexpr int kRaysPerCta = 4;
__global__ void __launch_bounds__(256) k_ray_integrate(const double2* __restrict__ all_xy, int B, int Bpad,
                                                       const Stats* __restrict__ st,
                                                       uint32_t* __restrict__ ray_cells, int cap,
                                                       int* __restrict__ ray_count, float2* __restrict__ ray_start,
                                                       int* __restrict__ ray_maxlen, int4* __restrict__ rect,
                                                       const float* __restrict__ meas,
                                                       const uint8_t* __restrict__ hit,
                                                       CellCounts* __restrict__ counts, uint32_t* __restrict__ dirty,
                                                       uint32_t* __restrict__ stale_bitmap, int stale_words,
                                                       int walk_only, Geometry g) {
    __shared__ int s_len[kRaysPerCta];
    __shared__ int s_max;
    const int tid = threadIdx.x;
    // the dirty-tile buffer the refresh of THIS step consumed (host double buffering) is re-armed here: this
    // kernel is ordered after that refresh and before the next writer of the buffer
    for (int i = blockIdx.x * blockDim.x + tid; i < stale_words; i += gridDim.x * blockDim.x) stale_bitmap[i] = 0u;
    if (st->xerror) return;
    // the strongest pose as k_norm_coop snapshotted it: this kernel may run on a side stream while the next
    // step's motion update already rewrites the pose array
    const float4 p = make_float4(st->strongest_pose[0], st->strongest_pose[1], st->strongest_pose[2], 0.f);
    const Xform t(p.x, p.y, p.z);
    const float sx = (float)((t.tx(0.0, 0.0) - g.posx) / g.res);
    const float sy = (float)((t.ty(0.0, 0.0) - g.posy) / g.res);
    const int b0 = blockIdx.x * kRaysPerCta;
    if (tid < 32) {
        const int b = b0 + tid;
        int c = 0;
        if (tid < kRaysPerCta && b < B) {
            const double2 m = all_xy[b];
            const float ex = (float)((t.tx(m.x, m.y) - g.posx) / g.res);
            const float ey = (float)((t.ty(m.x, m.y) - g.posy) / g.res);
            if (b == 0) *ray_start = make_float2(sx, sy);
            RayIter it;
            it.init(sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, g.extra_steps);
            int x = it.x, y = it.y, n = min(it.n, cap);
            const int xi = it.x_inc, yi = it.y_inc;
            const float dx = it.dx, dy = it.dy;
            float err = it.error;
            const int fx = x, fy = y;
            uint32_t* out = ray_cells + b;
            // A step moves one cell along one axis, so the first min(room to the map border) + 1 cells cannot be
            // out of bounds: they are walked without the bounds test (RayIterator.hasNext reduces to n > 0), the
            // loop is then bounded by the f32 compare -> select -> add chain of the error term alone.
            int n_fast = 0;
            if ((unsigned)x < (unsigned)g.W && (unsigned)y < (unsigned)g.H) {
                const int room_x = xi > 0 ? g.W - 1 - x : (xi < 0 ? x : 0x3fffffff);
                const int room_y = yi > 0 ? g.H - 1 - y : (yi < 0 ? y : 0x3fffffff);
                n_fast = min(n, min(room_x, room_y) + 1);
            }
#pragma unroll 4
            for (int k = 0; k < n_fast; k++) {
                out[(size_t)k * Bpad] = (uint32_t)x | ((uint32_t)y << 16);
                const bool up = err > 0.0f;  // RayIterator.next
                y += up ? yi : 0;
                x += up ? 0 : xi;
                err += up ? -dx : dy;
            }
            c = n_fast;
            n -= n_fast;
            while (n > 0 && (unsigned)x < (unsigned)g.W && (unsigned)y < (unsigned)g.H) {  // RayIterator.hasNext
                out[(size_t)c * Bpad] = (uint32_t)x | ((uint32_t)y << 16);
                c++;
                const bool up = err > 0.0f;
                y += up ? yi : 0;
                x += up ? 0 : xi;
                err += up ? -dx : dy;
                n--;
            }
            if (c > 0) {  // the walk is monotone in x and y: its box is spanned by the first and the last cell
                const uint32_t last = out[(size_t)(c - 1) * Bpad];
                const int lx = (int)(last & 0xffffu), ly = (int)(last >> 16);
                int* r = reinterpret_cast<int*>(rect);
                atomicMin(r + 0, min(fx, lx));
                atomicMin(r + 1, min(fy, ly));
                atomicMax(r + 2, max(fx, lx));
                atomicMax(r + 3, max(fy, ly));
            }
        }
        if (tid < kRaysPerCta) {
            if (b < Bpad) ray_count[b] = c;
            s_len[tid] = c;
        }
        int mx = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (tid == 0) {
            s_max = mx;
            if (mx > 0) atomicMax(ray_maxlen, mx);
        }
    }
    __syncthreads();  // also orders warp 0's global stores before the other warps' loads (same CTA)
    if (walk_only) return;
    const int total = s_max * kRaysPerCta;
    for (int e0 = tid; e0 < total; e0 += 4 * 256) {
        uint32_t cell[4];
        int bb[4], cls[4];
        unsigned long long old[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {  // four independent loads in flight per thread
            const int e = e0 + u * 256;
            const int k = e / kRaysPerCta, j = e % kRaysPerCta;
            bb[u] = -1;
            if (e < total && k < s_len[j]) {
                bb[u] = b0 + j;
                cell[u] = __ldcg(ray_cells + (size_t)k * Bpad + bb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {  // ... then four independent atomics (the returned pair is only needed below)
            cls[u] = 0;
            if (bb[u] < 0) continue;
            const int cx = (int)(cell[u] & 0xffffu), cy = (int)(cell[u] >> 16);
            const float dX = sx - ((float)cx + 0.5f);
            const float dY = sy - ((float)cy + 0.5f);
            const float dist = __fsqrt_rn(dX * dX + dY * dY);
            cls[u] = inverse_sensor_class(dist, meas[bb[u]], hit[bb[u]] != 0, g.tol_half);
            if (cls[u] != 0)
                old[u] = atomicAdd(reinterpret_cast<unsigned long long*>(counts + ((size_t)cx + (size_t)cy * g.W)),
                                   cls[u] == 1 ? 1ull : (1ull << 32));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (cls[u] != 0 && code_flips((uint32_t)old[u], (uint32_t)(old[u] >> 32), cls[u], g))
                mark_dirty(dirty, (int)(cell[u] & 0xffffu), (int)(cell[u] >> 16), g);
    }
}

// ---- atomic-free alternative (GMS_UPDATE_SORTED): keys -> sort -> run lengths -> one writer per cell ----
// key = cell index << 2 | class (1 free, 2 occupied); class-0 cells and padding get the sentinel 0xFFFFFFFF.
__global__ void __launch_bounds__(256) k_ray_keys(const uint32_t* __restrict__ ray_cells, int Bpad, int maxlen,
                                                  const int* __restrict__ ray_count,
                                                  const float2* __restrict__ ray_start,
                                                  const float* __restrict__ meas, const uint8_t* __restrict__ hit,
                                                  uint32_t* __restrict__ keys, Geometry g) {
    const long long total = (long long)maxlen * Bpad;
    const float2 s = *ray_start;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e / Bpad), b = (int)(e - (long long)k * Bpad);
        uint32_t key = 0xffffffffu;
        if (k < ray_count[b]) {
            const uint32_t cell = ray_cells[e];
            const int cx = (int)(cell & 0xffffu), cy = (int)(cell >> 16);
            const float dX = s.x - ((float)cx + 0.5f);
            const float dY = s.y - ((float)cy + 0.5f);
            const float dist = __fsqrt_rn(dX * dX + dY * dY);
            const int cls = inverse_sensor_class(dist, meas[b], hit[b] != 0, g.tol_half);
            if (cls != 0) key = ((uint32_t)(cx + cy * g.W) << 2) | (uint32_t)cls;
        }
        keys[e] = key;
    }
}
// one thread per run; the thread of a cell's FIRST run applies all (<= 2) runs of that cell with plain stores
__global__ void __launch_bounds__(256) k_apply_runs(const uint32_t* __restrict__ ukeys, const int* __restrict__ runlen,
                                                    const int* __restrict__ num_runs, CellCounts* __restrict__ counts,
                                                    uint32_t* __restrict__ dirty, Geometry g) {
    const int n = *num_runs;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = ukeys[i];
        if (key == 0xffffffffu) continue;
        const uint32_t cell = key >> 2;
        if (i > 0 && (ukeys[i - 1] >> 2) == cell) continue;  // second run of this cell: applied by the first
        CellCounts c = counts[cell];
        const int before = cell_code(c.n_free, c.n_occ, g);
        if ((key & 3u) == 1u) c.n_free += (uint32_t)runlen[i]; else c.n_occ += (uint32_t)runlen[i];
        if (i + 1 < n && (ukeys[i + 1] >> 2) == cell && ukeys[i + 1] != 0xffffffffu) c.n_occ += (uint32_t)runlen[i + 1];
        counts[cell] = c;
        if (cell_code(c.n_free, c.n_occ, g) != before) mark_dirty(dirty, (int)(cell % (uint32_t)g.W), (int)(cell / (uint32_t)g.W), g);
    }
}

// single ray given in grid coordinates (gms_map_apply_measurement)
__global__ void k_apply_one(CellCounts* __restrict__ counts, int4* __restrict__ rect, uint32_t* __restrict__ dirty,
                            float sx, float sy, float ex, float ey, float meas, int was_hit, Geometry g) {
    CellBox box;
    apply_measurement(counts, dirty, g, sx, sy, ex, ey, meas, was_hit != 0, box);
    if (box.x1 >= 0) {
        rect->x = min(rect->x, box.x0); rect->y = min(rect->y, box.y0);
        rect->z = max(rect->z, box.x1); rect->w = max(rect->w, box.y1);
    }
}

// RayIterator cell sequences (gms_trace_rays)
__global__ void k_trace_rays(const float4* __restrict__ rays, int n, int extra, int W, int H,
                             int2* __restrict__ cells, int cap, int* __restrict__ counts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float4 q = rays[r];
    RayIter it;
    it.init(q.x, q.y, q.z, q.w, extra);
    int c = 0;
    while (it.has_next(W, H)) {
        if (c < cap) cells[(size_t)cap * r + c] = make_int2(it.x, it.y);
        it.advance();
        c++;
    }
    counts[r] = c;
}

// ------------------------------------------------------------------------------------------------
// A6 — normalise (SLAM.java:119-121), Neff (:180-190), strongest (:110-115), weighted pose (:165-178).
// P <= ~1e6 doubles: latency-bound.  ceil(P/1024) CTAs of 1024 threads; per-CTA partials go to global
// memory and are combined in a FIXED order (by every CTA redundantly, or by the last CTA to finish), so
// the totals are run-to-run and rank-to-rank bit-identical.  (An 8-CTA cluster/DSMEM version measured
// 33 us here: it confines the exp/divide work to 8 SMs.)
// ------------------------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce_1024(T v, Op op, T* s_buf) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) s_buf[wid] = v;
    __syncthreads();
    v = s_buf[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// N sums at once over a 1024-thread block (one shared-memory exchange instead of N): fixed tree, result in every thread
template <int N>
__device__ __forceinline__ void block_sum_vec_1024(double (&v)[N], double* s_buf /* 32 * N */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < N; k++) s_buf[k * 32 + wid] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; k++) {
        v[k] = s_buf[k * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
}
struct SumOp { __device__ double operator()(double a, double b) const { return a + b; } };
struct SumU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };

// The same reductions for blocks of NT = 64 .. 1024 threads (NT / 32 a power of two): fixed trees, result in every thread.
template <int NT, typename T, typename Op>
__device__ __forceinline__ T block_reduce_nt(T v, Op op, T* s_buf /* NT / 32 */) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) s_buf[wid] = v;
    __syncthreads();
    v = s_buf[lane & (NW - 1)];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int NT, int N>
__device__ __forceinline__ void block_sum_vec_nt(double (&v)[N], double* s_buf /* (NT / 32) * N */) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < N; k++) s_buf[k * NW + wid] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; k++) {
        v[k] = s_buf[k * NW + (lane & (NW - 1))];
#pragma unroll
        for (int o = NW / 2; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
}
template <int NT>
__device__ __forceinline__ void block_argmax_nt(double& best, int& bi, double* s_key, int* s_idx) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { s_key[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    best = s_key[lane & (NW - 1)]; bi = s_idx[lane & (NW - 1)];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
}

// (max, first index of the max) over a block; result valid in every thread
__device__ __forceinline__ void block_argmax_1024(double& best, int& bi, double* s_key, int* s_idx) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { s_key[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    best = s_key[lane]; bi = s_idx[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
}

// ---- exact sums ---------------------------------------------------------------------------------------------
// Sums over particles that must not depend on how the particles are partitioned (tiles, CTAs, ranks) are
// accumulated EXACTLY: every non-negative term x < 2^27 is converted to the integer floor(x * 2^100) (128 bits)
// and the integers are added.  Integer addition is associative, so any grouping gives the same total; the
// truncation error is < 2^-100 per term.  (Up to 2^20 terms of at most 2^127 / 2^20 each: no overflow.)
struct U128 {
    unsigned long long lo, hi;
};
__device__ __forceinline__ U128 fx100(double x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const int e = (int)((bits >> 52) & 0x7ffull);
    U128 r{0ull, 0ull};
    if (e == 0 || e == 0x7ff || (long long)bits < 0) return r;  // zero / subnormal / non-finite / negative: no contribution
    const unsigned long long mant = (bits & 0xfffffffffffffull) | (1ull << 52);
    const int sh = e - 975;  // x = mant * 2^(e - 1075); floor(x * 2^100) = mant shifted by e - 975
    if (sh >= 64) { r.hi = mant << (sh - 64); }
    else if (sh > 0) { r.lo = mant << sh; r.hi = mant >> (64 - sh); }
    else if (sh > -64) { r.lo = mant >> (-sh); }
    return r;
}
__device__ __forceinline__ void add128(U128& a, const U128 b) {
    const unsigned long long lo = a.lo + b.lo;
    a.hi += b.hi + (lo < a.lo ? 1ull : 0ull);
    a.lo = lo;
}
__device__ __forceinline__ double to_double100(const U128 a) {  // deterministic function of the integer
    return (double)a.hi * 0x1p-36 + (double)a.lo * 0x1p-100;
}
template <int NT>
__device__ __forceinline__ U128 block_sum128_nt(U128 v, unsigned long long* s_buf /* 2 * NT / 32 */) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        U128 u;
        u.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        u.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        add128(v, u);
    }
    __syncthreads();
    if (lane == 0) { s_buf[2 * wid] = v.lo; s_buf[2 * wid + 1] = v.hi; }
    __syncthreads();
    v.lo = s_buf[2 * (lane & (NW - 1))]; v.hi = s_buf[2 * (lane & (NW - 1)) + 1];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) {
        U128 u;
        u.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        u.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        add128(v, u);
    }
    return v;
}

// ---- sharded normalise / resample: what ranks tell each other (a few hundred bytes per step and rank) -----------
// Multi-rank shared map on the peer path: every rank normalises ONLY its own block of particles and selects ONLY its
// own children.  Three tiny all-to-all rounds replace the exchange of per-particle data:
//   A  {max log-weight of the block, its first index, that particle's pose}        -> global M, strongest, its pose
//   B  exact sum of exp(lw - M) over the block                                     -> S
//   C  fixed-point weight total of the block (CDF offsets), exact sums of w, w^2   -> Neff, resample decision
// and, only when a resampling follows,
//   D  every 64th value of the block's (block-relative) fixed-point CDF            -> children find their parent's
//      rank and 64-particle bucket locally, and finish the search with <= 6 reads of that rank's CDF segment
// Each record is stored into every rank's XArea (peer-mapped, double-buffered by step parity), followed by
// __threadfence_system and a release store of the step sequence number into that rank's flag for the round;
// consumers poll their own flags (ld.acquire.sys, bounded spin -> Stats.xerror).
constexpr int kCoarseStep = 64;
struct XRecA { double m; int idx; float x, y, t; int pad; };
struct XRecB { unsigned long long lo, hi; };
struct XRecC { unsigned long long T, q_lo, q_hi, w_lo, w_hi, pad; };
struct XArea {  // followed in memory by kMaxRanks coarse tables of `coarse_cap` u64 each
    XRecA a[kMaxRanks];
    XRecB b[kMaxRanks];
    XRecC c[kMaxRanks];
};
struct Shard {           // by value in the kernel arguments; nranks == 0: not sharded
    int nranks, rank;
    unsigned long long seq;
    unsigned char* area[kMaxRanks];         // every rank's XArea of this step's parity (area[rank] = my own)
    unsigned long long* flags[kMaxRanks];   // every rank's flag block: [round][sender]
    const unsigned long long* cdfseg[kMaxRanks];  // every rank's block-relative CDF segment (k_resample_coop)
    const double* w[kMaxRanks];             // every rank's weight / log-weight / pose arrays of the current generation
    const double* lw[kMaxRanks];
    int coarse_cap;
};
__device__ __forceinline__ unsigned long long* coarse_of(unsigned char* area, int coarse_cap, int sender) {
    return reinterpret_cast<unsigned long long*>(area + sizeof(XArea)) + (size_t)sender * coarse_cap;
}
// threads 0 .. nranks-1 of the calling CTA: wait until every rank's flag of `round` carries `seq`
__device__ __forceinline__ void shard_wait(const Shard& sh, int round, Stats* st) {
    if ((int)threadIdx.x < sh.nranks) {
        const unsigned long long* f = sh.flags[sh.rank] + round * kMaxRanks + threadIdx.x;
        unsigned long long v = 0;
        for (long long spins = 0; spins < 8000000; spins++) {  // ~4 s with the sleeps: never hang the device
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= sh.seq) break;
            __nanosleep(100);
        }
        if (v < sh.seq) st->xerror = 1;
    }
    __syncthreads();  // the acquiring threads' view is handed to the whole CTA
}
// thread q < nranks of ONE CTA: after it stored its record into rank q's area: fence, then raise my flag there
__device__ __forceinline__ void shard_signal(const Shard& sh, int round) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sh.flags[threadIdx.x] + round * kMaxRanks + sh.rank), "l"(sh.seq)
                 : "memory");
}

// Normalise + Neff + strongest in ONE cooperative launch (grid barriers inside), over the block [lo, lo + cnt) of
// the particle index: the whole set (single rank, or lw replicated by k_xpush_lw / an all-gather) or this rank's
// shard (see above).  M and the first arg-max are order-independent by nature; S, sum w and sum w^2 are exact
// integer sums (U128); the CDF is u64 fixed point.  So the weights, Neff, the strongest particle and the parent
// indices do not depend on the grid size, on which CTA handled which tile, on the scoring kernel's processing
// order, or on how many ranks share the particle set (run-to-run, rank-to-rank bit-identical).
//   phase 0  (replicated peer exchange only) wait until every rank's log-weights have landed
//   phase 1  tile maxima -> grid barrier -> block max [round A] -> M (SLAM.java:110-115 keeps the first maximum)
//   phase 2  e_i = exp(lw_i - M), exact tile sums -> grid barrier -> block sum [round B] -> S
//   phase 3  w_i = e_i / S (SLAM.java:119-121); tile sums of trunc(w * 2^60), exact tile sums of w and w^2
//   final    (last CTA to finish) block totals [round C] -> Neff = (sum w)^2 / sum w^2 (SLAM.java:180-190),
//            strongest pose snapshot, resample decision
struct NormArgs {
    const double* lw;    // log-weights, indexed by GLOBAL particle index (own array, or the lw receive buffer)
    double* lw_store;    // replicated peer exchange: received values are also filed in the handle's own lw array
    double* w;
    PoseTable poses;
    int P, lo, cnt, ntiles, policy;  // ntiles = ceil(cnt / 1024), tiles start at lo
    NormPartials np;
    Stats* st;
    const unsigned long long* xflags;  // replicated peer exchange: arrival flags of the log-weights, else nullptr
    int nranks;
    unsigned long long seq;
    const float4* pose_local;  // fold the weighted pose (SLAM.getWeightedPose) into this launch: all poses of the block
    double* wp_part;           // are local (single rank / imported records), else nullptr
    Shard sh;
    XPull pull;                // replicated peer exchange, pull form (k_norm_tiles only): see XPull
};
constexpr int kNormThreads = 256;   // 4 consecutive particles per thread: a 1024-particle tile per CTA iteration
constexpr int kNormCtasPerSm = 4;   // <= 64 registers (the exact 128-bit sums need them): 592 co-resident CTAs
__global__ void __launch_bounds__(kNormThreads, kNormCtasPerSm) k_norm_coop(NormArgs a) {
    cg::grid_group grid = cg::this_grid();
    constexpr int NT = kNormThreads, NW = NT / 32;
    __shared__ double s_key[NW];
    __shared__ int s_idx[NW];
    __shared__ double s_v[3 * NW];
    __shared__ unsigned long long s_u[2 * NW];
    __shared__ bool s_last;
    const int tid = threadIdx.x, G = gridDim.x;
    const bool sharded = a.sh.nranks > 0;
    const int end = a.lo + a.cnt;
    if (a.xflags) {
        if (tid < a.nranks) {  // every CTA polls for itself: no extra grid barrier
            unsigned long long v = 0;
            long long spins = 0;
            for (; spins < 8000000; spins++) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.xflags + tid) : "memory");
                if (v >= a.seq) break;
                __nanosleep(200);
            }
            if (v < a.seq) a.st->xerror = 1;
        }
        __syncthreads();
    }
    // phase 1: per tile, (max, first arg-max)
    for (int t = blockIdx.x; t < a.ntiles; t += G) {
        const int i0 = a.lo + t * 1024 + tid * 4;
        double best = kNegInf;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < end) {
                const double v = __ldcg(a.lw + i0 + j);
                if (a.lw_store) a.lw_store[i0 + j] = v;
                if (v > best) { best = v; bi = i0 + j; }  // ascending index: a later equal value never replaces
            }
        }
        block_argmax_nt<NT>(best, bi, s_key, s_idx);
        if (tid == 0) { a.np.m[t] = best; a.np.idx[t] = bi; }
    }
    __threadfence();
    grid.sync();
    if (*(volatile int*)&a.st->xerror) return;  // uniform: set (if at all) before the barrier
    double best = kNegInf;
    int bi = 0x7fffffff;
    for (int c = tid; c < a.ntiles; c += NT) {
        const double v = __ldcg(a.np.m + c);
        const int vi = __ldcg(a.np.idx + c);
        if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
    }
    block_argmax_nt<NT>(best, bi, s_key, s_idx);
    float4 bpose = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sharded) {
        if (blockIdx.x == 0 && tid < a.sh.nranks) {  // round A
            const float4 p = a.poses.at(bi);  // my own block: local
            XRecA r; r.m = best; r.idx = bi; r.x = p.x; r.y = p.y; r.t = p.z; r.pad = 0;
            reinterpret_cast<XArea*>(a.sh.area[tid])->a[a.sh.rank] = r;
            shard_signal(a.sh, 0);
        }
        shard_wait(a.sh, 0, a.st);
        const XArea* mine = reinterpret_cast<const XArea*>(a.sh.area[a.sh.rank]);
        best = kNegInf; bi = 0x7fffffff;
        for (int q = 0; q < a.sh.nranks; q++) {
            const volatile XRecA* r = &mine->a[q];
            const double rm = r->m;
            const int ri = r->idx;
            if (rm > best || (rm == best && ri < bi)) { best = rm; bi = ri; bpose = make_float4(r->x, r->y, r->t, 0.f); }
        }
    }
    // phase 2: e_i = exp(lw_i - M) (parked in w), exact tile sums
    for (int t = blockIdx.x; t < a.ntiles; t += G) {
        const int i0 = a.lo + t * 1024 + tid * 4;
        U128 acc{0ull, 0ull};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < end) {
                const double e = exp(__ldcg(a.lw + i0 + j) - best);
                a.w[i0 + j] = e;
                add128(acc, fx100(e));
            }
        }
        acc = block_sum128_nt<NT>(acc, s_u);
        if (tid == 0) { a.np.x128[6 * t] = acc.lo; a.np.x128[6 * t + 1] = acc.hi; }
    }
    __threadfence();
    grid.sync();
    U128 sacc{0ull, 0ull};
    for (int c = tid; c < a.ntiles; c += NT) add128(sacc, U128{__ldcg(a.np.x128 + 6 * c), __ldcg(a.np.x128 + 6 * c + 1)});
    sacc = block_sum128_nt<NT>(sacc, s_u);
    if (sharded) {
        if (blockIdx.x == 0 && tid < a.sh.nranks) {  // round B
            reinterpret_cast<XArea*>(a.sh.area[tid])->b[a.sh.rank] = XRecB{sacc.lo, sacc.hi};
            shard_signal(a.sh, 1);
        }
        shard_wait(a.sh, 1, a.st);
        const XArea* mine = reinterpret_cast<const XArea*>(a.sh.area[a.sh.rank]);
        sacc = U128{0ull, 0ull};
        for (int q = 0; q < a.sh.nranks; q++) {
            const volatile XRecB* r = &mine->b[q];
            add128(sacc, U128{r->lo, r->hi});
        }
    }
    const double S = to_double100(sacc);
    // phase 3: w_i = e_i / S; tile sums
    for (int t = blockIdx.x; t < a.ntiles; t += G) {
        const int i0 = a.lo + t * 1024 + tid * 4;
        U128 qacc{0ull, 0ull}, wacc{0ull, 0ull};
        double v[3] = {0.0, 0.0, 0.0};  // w*x, w*y, w*angleConstrain(theta) (folded weighted pose only)
        unsigned long long fx = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = i0 + j;
            if (i < end) {
                const double wi = a.w[i] / S;
                a.w[i] = wi;
                fx += (unsigned long long)(wi * 0x1p60);
                add128(wacc, fx100(wi));
                add128(qacc, fx100(wi * wi));
                if (a.pose_local) {  // SLAM.getWeightedPose SLAM.java:165-178
                    const float4 p = a.pose_local[i];
                    v[0] += (double)p.x * wi; v[1] += (double)p.y * wi; v[2] += angle_constrain((double)p.z) * wi;
                }
            }
        }
        qacc = block_sum128_nt<NT>(qacc, s_u);
        wacc = block_sum128_nt<NT>(wacc, s_u);
        fx = block_reduce_nt<NT>(fx, SumU64(), s_u);
        if (a.pose_local) block_sum_vec_nt<NT, 3>(v, s_v);
        if (tid == 0) {
            a.np.fx[t] = fx;
            a.np.x128[6 * t + 2] = qacc.lo; a.np.x128[6 * t + 3] = qacc.hi;
            a.np.x128[6 * t + 4] = wacc.lo; a.np.x128[6 * t + 5] = wacc.hi;
            if (a.pose_local) { a.wp_part[4 * t] = v[0]; a.wp_part[4 * t + 1] = v[1]; a.wp_part[4 * t + 2] = v[2]; }
        }
    }
    // the last CTA to finish folds the tile sums and publishes the step's statistics
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(a.np.counter, 1u) == (unsigned)G - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    U128 qt{0ull, 0ull}, wt{0ull, 0ull};
    unsigned long long T = 0;
    double f[3] = {0.0, 0.0, 0.0};
    for (int c = tid; c < a.ntiles; c += NT) {
        add128(qt, U128{__ldcg(a.np.x128 + 6 * c + 2), __ldcg(a.np.x128 + 6 * c + 3)});
        add128(wt, U128{__ldcg(a.np.x128 + 6 * c + 4), __ldcg(a.np.x128 + 6 * c + 5)});
        T += __ldcg(a.np.fx + c);
        if (a.pose_local) { f[0] += __ldcg(a.wp_part + 4 * c); f[1] += __ldcg(a.wp_part + 4 * c + 1); f[2] += __ldcg(a.wp_part + 4 * c + 2); }
    }
    qt = block_sum128_nt<NT>(qt, s_u);
    wt = block_sum128_nt<NT>(wt, s_u);
    T = block_reduce_nt<NT>(T, SumU64(), s_u);
    if (a.pose_local) block_sum_vec_nt<NT, 3>(f, s_v);
    if (sharded) {
        if (tid < a.sh.nranks) {  // round C
            XRecC r; r.T = T; r.q_lo = qt.lo; r.q_hi = qt.hi; r.w_lo = wt.lo; r.w_hi = wt.hi; r.pad = 0;
            reinterpret_cast<XArea*>(a.sh.area[tid])->c[a.sh.rank] = r;
            shard_signal(a.sh, 2);
        }
        shard_wait(a.sh, 2, a.st);
        const XArea* mine = reinterpret_cast<const XArea*>(a.sh.area[a.sh.rank]);
        qt = U128{0ull, 0ull}; wt = U128{0ull, 0ull};
        for (int q = 0; q < a.sh.nranks; q++) {
            const volatile XRecC* r = &mine->c[q];
            add128(qt, U128{r->q_lo, r->q_hi});
            add128(wt, U128{r->w_lo, r->w_hi});
        }
    }
    if (tid == 0) {
        Stats* st = a.st;
        const double sa = to_double100(wt), sq = to_double100(qt);
        if (a.pose_local) {
            st->weighted_pose[0] = (float)(f[0] / sa);
            st->weighted_pose[1] = (float)(f[1] / sa);
            st->weighted_pose[2] = (float)(f[2] / sa);
        }
        const double neff = (sa * sa) / sq;
        st->neff = neff;
        st->lw_max = best;
        st->sum_exp = S;
        st->strongest = bi;
        st->strongest_now = bi;
        st->strongest_w = 1.0 / S;
        const float4 p = sharded ? bpose : a.poses.at(bi);
        st->strongest_pose[0] = p.x; st->strongest_pose[1] = p.y; st->strongest_pose[2] = p.z;
        st->do_resample = a.policy == 2 || (a.policy == 1 && neff < (double)(a.P / 2));  // GridMapApp.java:185
        *a.np.counter = 0u;
    }
}

// The default normalise: the same step over FIXED tiles of 1024 consecutive particle indices with plain f64 sums —
// each tile reduced by a fixed shuffle tree, tile results combined in tile order, S = sum_t s_t * exp(m_t - M) from
// per-tile maxima — which needs ONE grid barrier instead of two (the exact sums above need M before the first
// exponential).  Independent of the grid size, of which CTA handled a tile, of the scoring kernel's order and of the
// rank count as long as every rank normalises the whole set (single rank, replicated peer exchange, all-gather):
// the default.  17 us against 24 us at 100k particles.  k_norm_coop (exact, order-free sums) serves GMS_SHARDED=1.
__global__ void __launch_bounds__(kNormThreads, 6) k_norm_tiles(NormArgs a) {
    cg::grid_group grid = cg::this_grid();
    constexpr int NT = kNormThreads, NW = NT / 32;
    __shared__ double s_key[NW];
    __shared__ int s_idx[NW];
    __shared__ double s_d[NW];
    __shared__ double s_v[5 * NW];
    __shared__ unsigned long long s_u[NW];
    __shared__ double s_tile[1024];
    __shared__ bool s_last;
    const int tid = threadIdx.x, G = gridDim.x;
    const bool pull = a.pull.cnt > 0;
    if (pull && blockIdx.x == 0 && tid < a.nranks) {  // this rank's block is final (kernel boundary): tell every rank
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.pull.flag[tid] + a.pull.myrank), "l"(a.seq) : "memory");
    }
    if (a.xflags) {
        if (tid < a.nranks) {  // every CTA polls for itself: no extra grid barrier
            unsigned long long v = 0;
            long long spins = 0;
            for (; spins < 8000000; spins++) {  // ~4 s with the sleeps: never hang the device
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.xflags + tid) : "memory");
                if (v >= a.seq) break;
                __nanosleep(200);
            }
            if (v < a.seq) a.st->xerror = 1;
        }
        __syncthreads();  // the acquiring threads' view is handed to the whole CTA
    }
    const double* lw2 = pull ? a.lw_store : a.lw;  // phase 2 re-reads the values: pulled ones from the local copy phase 1 filed
    // phase 1: per fixed tile, (max, first arg-max, sum exp(lw - tile max))
    for (int t = blockIdx.x; t < a.ntiles; t += G) {
        const int i0 = t * 1024 + tid * 4;
        double v[4];
        if (pull && (a.pull.cnt & 1) == 0) {
            // each value from the exchange buffer of the rank that owns the particle.  The tile is staged through
            // shared memory so that a warp's request is 512 contiguous bytes and every 32-byte sector crosses NVLink
            // once (a thread's own four values are 32 bytes apart from its neighbour's: read directly, every sector
            // would be requested four times — measured +10 us at 2 x 100k particles)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int e = 2 * (tid + k * NT), i = t * 1024 + e;  // cnt and P are even: a pair never straddles ranks
                double2 d = make_double2(kNegInf, kNegInf);
                if (i < a.P) d = __ldcg(reinterpret_cast<const double2*>(a.pull.src[i / a.pull.cnt] + i));
                s_tile[e] = d.x;
                s_tile[e + 1] = d.y;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = s_tile[tid * 4 + j];
            __syncthreads();
        } else if (pull) {
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = i0 + j < a.P ? __ldcg(a.pull.src[(i0 + j) / a.pull.cnt] + i0 + j) : kNegInf;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = i0 + j < a.P ? __ldcg(a.lw + i0 + j) : kNegInf;
        }
        if (a.lw_store)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i0 + j < a.P) a.lw_store[i0 + j] = v[j];
        double best = kNegInf;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (v[j] > best) { best = v[j]; bi = i0 + j; }  // ascending index: a later equal value never replaces
        block_argmax_nt<NT>(best, bi, s_key, s_idx);
        double e = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) e += i0 + j < a.P ? exp(v[j] - best) : 0.0;
        const double sum = block_reduce_nt<NT>(e, SumOp(), s_d);
        if (tid == 0) { a.np.m[t] = best; a.np.idx[t] = bi; a.np.s[t] = sum; }
    }
    __threadfence();
    grid.sync();
    if (*(volatile int*)&a.st->xerror) return;  // uniform: set (if at all) before the barrier
    // every CTA combines the tile partials in the same fixed order -> (M, first arg-max, S)
    double best = kNegInf;
    int bi = 0x7fffffff;
    for (int c = tid; c < a.ntiles; c += NT) {
        const double v = __ldcg(a.np.m + c);
        const int vi = __ldcg(a.np.idx + c);
        if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
    }
    block_argmax_nt<NT>(best, bi, s_key, s_idx);
    double acc = 0.0;
    for (int c = tid; c < a.ntiles; c += NT) acc += __ldcg(a.np.s + c) * exp(__ldcg(a.np.m + c) - best);
    const double S = block_reduce_nt<NT>(acc, SumOp(), s_d);
    // phase 2: w_i = exp(lw_i - M) / S, tile sums of w, w^2, trunc(w * 2^60) (+ weighted pose terms)
    for (int t = blockIdx.x; t < a.ntiles; t += G) {
        const int i0 = t * 1024 + tid * 4;
        double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // w, w^2, w*x, w*y, w*angleConstrain(theta)
        unsigned long long fx = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = i0 + j;
            if (i < a.P) {
                const double wi = exp(__ldcg(lw2 + i) - best) / S;
                a.w[i] = wi;
                v[0] += wi; v[1] += wi * wi;
                fx += (unsigned long long)(wi * 0x1p60);
                if (a.pose_local) {  // SLAM.getWeightedPose SLAM.java:165-178
                    const float4 p = a.pose_local[i];
                    v[2] += (double)p.x * wi; v[3] += (double)p.y * wi; v[4] += angle_constrain((double)p.z) * wi;
                }
            }
        }
        block_sum_vec_nt<NT, 5>(v, s_v);
        fx = block_reduce_nt<NT>(fx, SumU64(), s_u);
        if (tid == 0) {
            a.np.ws[t] = v[0]; a.np.q[t] = v[1]; a.np.fx[t] = fx;
            if (a.pose_local) { a.wp_part[4 * t] = v[2]; a.wp_part[4 * t + 1] = v[3]; a.wp_part[4 * t + 2] = v[4]; }
        }
    }
    // the last CTA to finish folds the tile sums (fixed order) and publishes the step's statistics
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(a.np.counter, 1u) == (unsigned)G - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double f[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int c = tid; c < a.ntiles; c += NT) {
        f[0] += __ldcg(a.np.ws + c);
        f[1] += __ldcg(a.np.q + c);
        if (a.pose_local) {
            f[2] += __ldcg(a.wp_part + 4 * c); f[3] += __ldcg(a.wp_part + 4 * c + 1); f[4] += __ldcg(a.wp_part + 4 * c + 2);
        }
    }
    block_sum_vec_nt<NT, 5>(f, s_v);
    const double sa = f[0], sq = f[1];
    if (tid == 0) {
        Stats* st = a.st;
        if (a.pose_local) {
            st->weighted_pose[0] = (float)(f[2] / sa);
            st->weighted_pose[1] = (float)(f[3] / sa);
            st->weighted_pose[2] = (float)(f[4] / sa);
        }
        const double neff = (sa * sa) / sq;
        st->neff = neff;
        st->lw_max = best;
        st->sum_exp = S;
        st->strongest = bi;
        st->strongest_now = bi;
        st->strongest_w = 1.0 / S;
        const float4 p = a.poses.at(bi);
        st->strongest_pose[0] = p.x; st->strongest_pose[1] = p.y; st->strongest_pose[2] = p.z;
        st->do_resample = a.policy == 2 || (a.policy == 1 && neff < (double)(a.P / 2));  // GridMapApp.java:185
        *a.np.counter = 0u;
    }
}

// SLAM.calculateNeff on the current weights (after set_weights / resample); also the fixed-point tile
// sums k_cdf_fixed needs.  Same last-block pattern.
__global__ void __launch_bounds__(1024) k_neff(const double* __restrict__ w, int P, int ntiles, NormPartials np,
                                               Stats* __restrict__ st) {
    __shared__ double s_d[32];
    __shared__ unsigned long long s_u[32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * 1024 + tid;
    const double wi = i < P ? w[i] : 0.0;
    const double ws = block_reduce_1024(wi, SumOp(), s_d);
    const double q = block_reduce_1024(wi * wi, SumOp(), s_d);
    const unsigned long long fx = block_reduce_1024((unsigned long long)(wi * 0x1p60), SumU64(), s_u);
    if (tid == 0) {
        np.ws[blockIdx.x] = ws;
        np.q[blockIdx.x] = q;
        np.fx[blockIdx.x] = fx;
        __threadfence();
        s_last = atomicAdd(np.counter, 1u) == (unsigned)ntiles - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double a = 0.0, b = 0.0;
    for (int c = tid; c < ntiles; c += 1024) {
        a += __ldcg(np.ws + c);
        b += __ldcg(np.q + c);
    }
    a = block_reduce_1024(a, SumOp(), s_d);
    b = block_reduce_1024(b, SumOp(), s_d);
    if (tid == 0) {
        st->neff_query = (a * a) / b;
        *np.counter = 0u;
    }
}

// SLAM.getWeightedPose SLAM.java:165-178 (plain, not circular, mean of angleConstrain(theta))
__global__ void __launch_bounds__(1024) k_weighted_pose(const double* __restrict__ w,
                                                        const float4* __restrict__ pose, int P, int ntiles,
                                                        double* __restrict__ part /* 4 * ntiles */,
                                                        unsigned* __restrict__ counter, Stats* __restrict__ st) {
    __shared__ double s_d[32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * 1024 + tid;
    double xs = 0, ys = 0, ts = 0, ws = 0;
    if (i < P) {
        const float4 p = pose[i];
        const double wi = w[i];
        xs = (double)p.x * wi;
        ys = (double)p.y * wi;
        ts = angle_constrain((double)p.z) * wi;
        ws = wi;
    }
    xs = block_reduce_1024(xs, SumOp(), s_d);
    ys = block_reduce_1024(ys, SumOp(), s_d);
    ts = block_reduce_1024(ts, SumOp(), s_d);
    ws = block_reduce_1024(ws, SumOp(), s_d);
    if (tid == 0) {
        part[4 * blockIdx.x + 0] = xs; part[4 * blockIdx.x + 1] = ys;
        part[4 * blockIdx.x + 2] = ts; part[4 * blockIdx.x + 3] = ws;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == (unsigned)ntiles - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    xs = ys = ts = ws = 0.0;
    for (int c = tid; c < ntiles; c += 1024) {
        xs += __ldcg(part + 4 * c + 0); ys += __ldcg(part + 4 * c + 1);
        ts += __ldcg(part + 4 * c + 2); ws += __ldcg(part + 4 * c + 3);
    }
    xs = block_reduce_1024(xs, SumOp(), s_d);
    ys = block_reduce_1024(ys, SumOp(), s_d);
    ts = block_reduce_1024(ts, SumOp(), s_d);
    ws = block_reduce_1024(ws, SumOp(), s_d);
    if (tid == 0) {
        st->weighted_pose[0] = (float)(xs / ws);
        st->weighted_pose[1] = (float)(ys / ws);
        st->weighted_pose[2] = (float)(ts / ws);
        *counter = 0u;
    }
}

// multi-rank: unpack the all-gathered exchange records into the global lw / pose arrays
__global__ void __launch_bounds__(256) k_import_exchange(const ExchangeRec* __restrict__ xg, int P,
                                                         double* __restrict__ lw, float4* __restrict__ pose) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const ExchangeRec r = xg[i];
    lw[i] = r.lw;
    pose[i] = make_float4(r.x, r.y, r.t, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// A7 — SLAM.resample SLAM.java:133-153.
// ------------------------------------------------------------------------------------------------
// LITERAL CDF: Java's sequential f64 running sum, c_i = c_{i-1} + w_i in particle order.  The rounding after
// every addition makes the chain inherently serial, so the kernel only keeps everything else off it: 256 threads
// stage a chunk of weights in shared memory (all loads in flight at once: one DRAM latency instead of one per 32
// values), ONE thread replays the dependent additions over shared memory, sixteen operands loaded ahead of their
// additions, and all threads store the chunk coalesced.  (Round 2's first version, one warp replaying the chain
// through shuffles with a global load every 32 values, took 24 us for 1000 particles.)
constexpr int kCdfChunk = 2048;
__global__ void __launch_bounds__(256) k_cdf_literal(const double* __restrict__ w, int P, double* __restrict__ cdf,
                                                     Stats* __restrict__ st, int force) {
    if (!(force || st->do_resample) || st->xerror) return;
    __shared__ double s_w[kCdfChunk];
    const int tid = threadIdx.x;
    if (tid == 0) st->strongest_now = -1;
    double c = 0.0;  // 0.0 + w[0] == w[0]: same as Java's c = particles.get(0).weight (SLAM.java:137)
    for (int base = 0; base < P; base += kCdfChunk) {
        const int n = min(kCdfChunk, P - base);
        double v[kCdfChunk / 256];
#pragma unroll
        for (int j = 0; j < kCdfChunk / 256; j++) v[j] = tid + j * 256 < n ? w[base + tid + j * 256] : 0.0;
#pragma unroll
        for (int j = 0; j < kCdfChunk / 256; j++) s_w[tid + j * 256] = v[j];
        __syncthreads();
        if (tid == 0) {
            int i = 0;
            for (; i + 16 <= n; i += 16) {
                double x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = s_w[i + j];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    c = c + x[j];
                    s_w[i + j] = c;
                }
            }
            for (; i < n; i++) {
                c = c + s_w[i];
                s_w[i] = c;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += 256) cdf[base + i] = s_w[i];
        __syncthreads();
    }
}

// index selection: for m = 1..P, U = r + (m-1)*1.0/P, first i with !(U > c_i), clamped to P-1.
// ... and the new generation is gathered right here: copies of the chosen parents (Particle(Particle)
// SLAM.java:41-45: weight and pose are copied; weights are NOT reset to 1/N).  A parent's pose is read through
// the PoseTable: with the peer exchange it may live in another rank's pose array (one 16-byte NVLink read).
struct SelectArgs {
    const void* cdf;     // f64 (LITERAL) or u64 fixed point (FIXED), P entries
    int P;
    double u01;
    uint64_t seed, resample_count;
    int* parents;
    Stats* st;
    PoseTable poses_in;
    const double* w_in;
    const double* lw_in;
    float4* pose_out;
    double* w_out;
    double* lw_out;
    int m_begin, m_count;  // children [m_begin, m_begin + m_count)
    int force;             // SLAM.resample() called explicitly: resample whatever the step's policy decided
    double* wp_part;       // k_resample_coop selecting ALL children: SLAM.getWeightedPose of the new generation comes
    unsigned* wp_counter;  // out of the same kernel (tile partials + last-CTA ticket); nullptr otherwise
};
__device__ __forceinline__ int select_stride(int P) {
    int stride = 32;
    while ((P + stride - 1) / stride > 2048) stride <<= 1;
    return stride;
}
template <bool FIXED, bool SMEM = false>  // SMEM: a.cdf points into shared memory (k_resample_small)
__device__ __forceinline__ void select_child(const SelectArgs& a, int m0, const unsigned long long* s_coarse, int stride,
                                             int ncoarse, double u01, float4& pose_o, double& w_o,
                                             const void* cdf_smem = nullptr) {
    using Key = typename std::conditional<FIXED, unsigned long long, double>::type;
    const int P = a.P;
    const double r = u01 * 1.0 / (double)P;
    auto key_of = [&](int m) -> Key {
        const double U = r + (double)m * 1.0 / (double)P;
        return FIXED ? (Key)(unsigned long long)(U * 0x1p60) : (Key)U;
    };
    // two-level search: every `stride`-th CDF value (<= 2048 of them) is staged in shared memory with one
    // round of independent loads, which replaces the top ~11 dependent global probes of a plain bisection
    const Key* cdf = static_cast<const Key*>(SMEM ? cdf_smem : a.cdf);
    const Key key = key_of(m0);
    const Key* coarse = reinterpret_cast<const Key*>(s_coarse);
    int lo = 0, hi = ncoarse - 1;
    while (lo < hi) {  // first segment whose last CDF value is not below the key
        const int mid = (lo + hi) >> 1;
        if (key > coarse[mid]) lo = mid + 1; else hi = mid;
    }
    hi = min(P - 1, (lo + 1) * stride - 1);
    lo = lo * stride;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key > (SMEM ? cdf[mid] : __ldcg(cdf + mid))) lo = mid + 1; else hi = mid;
    }
    a.parents[m0] = lo;
    pose_o = a.poses_in.at(lo);
    w_o = a.w_in[lo];
    a.pose_out[m0] = pose_o;
    a.w_out[m0] = w_o;
    a.lw_out[m0] = __ldcg(a.lw_in + lo);
    // the strongest particle of the last update lives on as its FIRST child (slam.py / gms_get_strongest)
    const int sb = a.st->strongest;
    if (lo == sb && (m0 == 0 || sb == 0 ? m0 == 0 : !(key_of(m0 - 1) > (SMEM ? cdf[sb - 1] : __ldcg(cdf + sb - 1)))))
        a.st->strongest_now = m0;
}

// LITERAL mode (P <= 2048 by default): Java's sequential f64 CDF (k_cdf_literal) + this selection kernel.
// Also used to complete a local-only FIXED selection (multi-rank: the children other ranks own).
template <bool FIXED>
__global__ void __launch_bounds__(256) k_select(SelectArgs a) {
    __shared__ unsigned long long s_coarse[2048];
    const int m0 = a.m_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (a.st->xerror) return;
    const bool resample = a.force || a.st->do_resample != 0;  // uniform over the grid
    const int stride = select_stride(a.P), ncoarse = (a.P + stride - 1) / stride;
    if (resample) {
        const unsigned long long* raw = static_cast<const unsigned long long*>(a.cdf);  // 8-byte keys either way
        for (int j = threadIdx.x; j < ncoarse; j += 256) s_coarse[j] = __ldcg(raw + min(a.P - 1, (j + 1) * stride - 1));
        __syncthreads();
    }
    if (m0 >= a.m_begin + a.m_count) return;
    if (!resample) {
        a.parents[m0] = m0;
        a.pose_out[m0] = a.poses_in.at(m0);
        a.w_out[m0] = a.w_in[m0];
        a.lw_out[m0] = __ldcg(a.lw_in + m0);
        return;
    }
    const double u01 = a.u01 < 0.0 ? philox_uniform(a.seed, a.resample_count) : a.u01;
    float4 po;
    double wo;
    select_child<FIXED>(a, m0, s_coarse, stride, ncoarse, u01, po, wo);
}

// LITERAL mode on one rank with at most kCdfChunk particles (the reference's own sizes: K1, K2, K2pp): k_cdf_literal,
// k_select<false> and — per-particle maps with a pending scan — k_list_parents as ONE launch of one CTA.  The CDF
// never leaves shared memory for the search (it is still written out: getters and the oracle comparison read it);
// the same chain, the same two-level search and the same strongest_now rule as the separate kernels, so parents,
// poses and weights are the same bits.  Saves two launches and a memset per resampling (~10 us of launch gaps).
__global__ void __launch_bounds__(1024) k_resample_small(SelectArgs a, int lo, int cnt, int* __restrict__ ulist,
                                                         int* __restrict__ n_used) {
    __shared__ double s_cdf[kCdfChunk];
    __shared__ unsigned long long s_coarse[kCdfChunk / 32];
    const int tid = threadIdx.x, P = a.P;
    if (a.st->xerror) return;
    const bool resample = a.force || a.st->do_resample != 0;
    if (tid == 0 && n_used) *n_used = 0;
    if (resample) {
#pragma unroll
        for (int j = 0; j < kCdfChunk / 1024; j++) {
            const int i = tid + j * 1024;
            s_cdf[i] = i < P ? a.w_in[i] : 0.0;
        }
        if (tid == 0) a.st->strongest_now = -1;
        __syncthreads();
        if (tid == 0) {  // Java's sequential f64 running sum (SLAM.java:137-147), see k_cdf_literal
            double c = 0.0;
            int i = 0;
            for (; i + 16 <= P; i += 16) {
                double x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = s_cdf[i + j];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    c = c + x[j];
                    s_cdf[i + j] = c;
                }
            }
            for (; i < P; i++) {
                c = c + s_cdf[i];
                s_cdf[i] = c;
            }
        }
        __syncthreads();
        double* cdf_out = static_cast<double*>(const_cast<void*>(a.cdf));
        const int stride = select_stride(P), ncoarse = (P + stride - 1) / stride;  // stride == 32 here
        for (int i = tid; i < P; i += 1024) cdf_out[i] = s_cdf[i];
        for (int j = tid; j < ncoarse; j += 1024) s_coarse[j] = (unsigned long long)__double_as_longlong(s_cdf[min(P - 1, (j + 1) * stride - 1)]);
        __syncthreads();
        const double u01 = a.u01 < 0.0 ? philox_uniform(a.seed, a.resample_count) : a.u01;
        for (int m0 = tid; m0 < P; m0 += 1024) {
            float4 po;
            double wo;
            select_child<false, true>(a, m0, s_coarse, stride, ncoarse, u01, po, wo, s_cdf);
        }
    } else {  // the generation is carried over unchanged
        for (int m0 = tid; m0 < P; m0 += 1024) {
            a.parents[m0] = m0;
            a.pose_out[m0] = a.poses_in.at(m0);
            a.w_out[m0] = a.w_in[m0];
            a.lw_out[m0] = __ldcg(a.lw_in + m0);
        }
    }
    if (ulist) {  // k_list_parents
        __syncthreads();
        for (int m = tid; m < P; m += 1024) {
            const int p = a.parents[m];
            if ((m == 0 || a.parents[m - 1] != p) && p >= lo && p < lo + cnt) ulist[atomicAdd(n_used, 1)] = p - lo;
        }
    }
}

// FIXED mode: CDF + selection in ONE cooperative launch.
//   phase 1  u64 fixed-point CDF, c_i = sum_{j<=i} trunc(w_j * 2^60): integer addition is associative, so the
//            parallel scan (fixed-point tile sums of k_norm_coop / k_neff + a block-wide scan per tile) equals
//            the sequential walk bit for bit on any number of threads / CTAs / ranks
//   phase 2  selection + gather of the children [m_begin, m_begin + m_count)
__global__ void __launch_bounds__(kNormThreads, kNormCtasPerSm) k_resample_coop(SelectArgs a,
                                                                   const unsigned long long* __restrict__ tile_fx,
                                                                   int ntiles) {
    cg::grid_group grid = cg::this_grid();
    constexpr int NT = kNormThreads, NW = NT / 32;
    __shared__ unsigned long long s_coarse[2048];
    __shared__ unsigned long long s_u[NW];
    __shared__ double s_v[4 * NW];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x;
    if (a.st->xerror) return;        // uniform over the grid
    if (!(a.force || a.st->do_resample)) {  // uniform over the grid: the generation is carried over unchanged
        for (int m0 = a.m_begin + blockIdx.x * NT + tid; m0 < a.m_begin + a.m_count; m0 += G * NT) {
            a.parents[m0] = m0;
            a.pose_out[m0] = a.poses_in.at(m0);
            a.w_out[m0] = a.w_in[m0];
            a.lw_out[m0] = __ldcg(a.lw_in + m0);
        }
        return;
    }
    unsigned long long* cdf = static_cast<unsigned long long*>(const_cast<void*>(a.cdf));
    for (int t = blockIdx.x; t < ntiles; t += G) {  // a tile = 1024 particles = 4 consecutive ones per thread
        unsigned long long acc = 0;
        for (int c = tid; c < t; c += NT) acc += __ldcg(tile_fx + c);
        const unsigned long long carry = block_reduce_nt<NT>(acc, SumU64(), s_u);
        const int i0 = t * 1024 + tid * 4;
        unsigned long long x[4], run = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            run += i0 + j < a.P ? (unsigned long long)(a.w_in[i0 + j] * 0x1p60) : 0ull;
            x[j] = run;  // inclusive within the thread
        }
        unsigned long long v = run;  // inclusive scan of the thread totals across the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        __syncthreads();
        if (lane == 31) s_u[wid] = v;
        __syncthreads();
        unsigned long long wbase = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) wbase += k < wid ? s_u[k] : 0ull;
        const unsigned long long excl = carry + wbase + v - run;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (i0 + j < a.P) cdf[i0 + j] = excl + x[j];
    }
    if (blockIdx.x == 0 && tid == 0) a.st->strongest_now = -1;
    __threadfence();
    grid.sync();
    const int stride = select_stride(a.P), ncoarse = (a.P + stride - 1) / stride;
    for (int j = tid; j < ncoarse; j += NT) s_coarse[j] = __ldcg(cdf + min(a.P - 1, (j + 1) * stride - 1));
    __syncthreads();
    const double u01 = a.u01 < 0.0 ? philox_uniform(a.seed, a.resample_count) : a.u01;
    const int nchunks = (a.m_count + NT - 1) / NT;
    for (int c = blockIdx.x; c < nchunks; c += G) {  // fixed chunks of NT children, one per thread (fixed reduction order)
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        const int m0 = a.m_begin + c * NT + tid;
        if (m0 < a.m_begin + a.m_count) {
            float4 po;
            double wo;
            select_child<true>(a, m0, s_coarse, stride, ncoarse, u01, po, wo);
            v[0] = (double)po.x * wo; v[1] = (double)po.y * wo; v[2] = angle_constrain((double)po.z) * wo; v[3] = wo;
        }
        if (a.wp_part) {
            block_sum_vec_nt<NT, 4>(v, s_v);
            if (tid == 0) { a.wp_part[4 * c] = v[0]; a.wp_part[4 * c + 1] = v[1]; a.wp_part[4 * c + 2] = v[2]; a.wp_part[4 * c + 3] = v[3]; }
        }
    }
    if (!a.wp_part) return;
    // SLAM.getWeightedPose of the new generation: the last CTA to finish folds the chunk sums in chunk order
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(a.wp_counter, 1u) == (unsigned)G - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double f[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = tid; c < nchunks; c += NT)
#pragma unroll
        for (int k = 0; k < 4; k++) f[k] += __ldcg(a.wp_part + 4 * c + k);
    block_sum_vec_nt<NT, 4>(f, s_v);
    if (tid == 0) {
        a.st->weighted_pose[0] = (float)(f[0] / f[3]);
        a.st->weighted_pose[1] = (float)(f[1] / f[3]);
        a.st->weighted_pose[2] = (float)(f[2] / f[3]);
        *a.wp_counter = 0u;
    }
}

// Sharded resampling (multi-rank shared map on the peer path): this rank builds the CDF segment of ITS particles and
// selects ITS children; see "sharded normalise / resample" above for the protocol.
//   phase 1  block-relative u64 fixed-point CDF of the block (tile sums of k_norm_coop + a block scan per tile);
//            every 64th value goes to every rank's coarse table (round D)
//   phase 2  per child: rank of the parent from the ranks' weight totals (round C), 64-particle bucket from the local
//            copy of that rank's coarse table, <= 6 probes of that rank's CDF segment (NVLink reads when remote), then
//            the parent's pose, weight and log-weight through the peer mappings
// The integers are those of the replicated CDF (c_i = offset of the rank + block-relative prefix), so the parents are
// the ones a single rank computes.
__global__ void __launch_bounds__(kNormThreads, kNormCtasPerSm) k_resample_shard(SelectArgs a, const unsigned long long* __restrict__ tile_fx,
                                                                    int ntiles, int lo, int cnt, Shard sh) {
    constexpr int NT = kNormThreads, NW = NT / 32;
    __shared__ unsigned long long s_u[NW];
    __shared__ unsigned long long s_base[kMaxRanks + 1];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x;
    Stats* st = a.st;
    if (st->xerror) return;        // uniform over the grid
    if (!(a.force || st->do_resample)) {  // uniform over the grid (and over the ranks): the generation is carried over unchanged
        for (int m0 = lo + blockIdx.x * NT + tid; m0 < lo + cnt; m0 += G * NT) {
            a.parents[m0] = m0;
            a.pose_out[m0] = a.poses_in.at(m0);
            a.w_out[m0] = a.w_in[m0];
            a.lw_out[m0] = a.lw_in[m0];
        }
        return;
    }
    unsigned long long* seg = static_cast<unsigned long long*>(const_cast<void*>(a.cdf));  // [cnt], block-relative
    for (int t = blockIdx.x; t < ntiles; t += G) {
        unsigned long long acc = 0;
        for (int c = tid; c < t; c += NT) acc += __ldcg(tile_fx + c);
        const unsigned long long carry = block_reduce_nt<NT>(acc, SumU64(), s_u);
        const int k0 = t * 1024 + tid * 4;  // position inside the block
        unsigned long long x[4], run = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            run += k0 + j < cnt ? (unsigned long long)(a.w_in[lo + k0 + j] * 0x1p60) : 0ull;
            x[j] = run;
        }
        unsigned long long v = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        __syncthreads();
        if (lane == 31) s_u[wid] = v;
        __syncthreads();
        unsigned long long wbase = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) wbase += k < wid ? s_u[k] : 0ull;
        const unsigned long long excl = carry + wbase + v - run;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j;
            if (k >= cnt) continue;
            const unsigned long long c = excl + x[j];
            seg[k] = c;
            if (((k + 1) & (kCoarseStep - 1)) == 0 || k == cnt - 1)  // round D: coarse sample j = k / 64 to every rank
                for (int q = 0; q < sh.nranks; q++) coarse_of(sh.area[q], sh.coarse_cap, sh.rank)[k / kCoarseStep] = c;
        }
    }
    __threadfence_system();  // segment + coarse samples visible system-wide before this thread's CTA takes its ticket
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.wp_counter, 1u) == (unsigned)G - 1u;
    __syncthreads();
    if (s_last) {
        if (tid == 0) { __threadfence_system(); *a.wp_counter = 0u; }
        __syncthreads();
        if (tid < sh.nranks) shard_signal(sh, 3);
    }
    shard_wait(sh, 3, st);
    if (*(volatile int*)&st->xerror) return;
    // offsets of the ranks' segments in the global CDF (round C totals), in shared memory
    const XArea* mine = reinterpret_cast<const XArea*>(sh.area[sh.rank]);
    if (tid == 0) {
        unsigned long long b = 0;
        for (int q = 0; q < sh.nranks; q++) { s_base[q] = b; b += reinterpret_cast<const volatile XRecC*>(&mine->c[q])->T; }
        s_base[sh.nranks] = b;
    }
    __syncthreads();
    const int P = a.P, ncoarse = (cnt + kCoarseStep - 1) / kCoarseStep;
    const double u01 = a.u01 < 0.0 ? philox_uniform(a.seed, a.resample_count) : a.u01;
    const double r = u01 * 1.0 / (double)P;
    auto key_of = [&](int m) -> unsigned long long {
        const double U = r + (double)m * 1.0 / (double)P;
        return (unsigned long long)(U * 0x1p60);
    };
    // global CDF value of particle i (any rank): offset + block-relative prefix
    auto cdf_of = [&](int i) -> unsigned long long { const int q = i / cnt; return s_base[q] + __ldcg(sh.cdfseg[q] + (i - q * cnt)); };
    for (int m0 = lo + blockIdx.x * NT + tid; m0 < lo + cnt; m0 += G * NT) {
        const unsigned long long key = key_of(m0);
        int q = 0;
        while (q < sh.nranks - 1 && key > s_base[q + 1]) q++;  // first rank whose last CDF value is not below the key
        const unsigned long long kk = key - s_base[q];
        const unsigned long long* coarse = coarse_of(sh.area[sh.rank], sh.coarse_cap, q);
        int l = 0, hgh = ncoarse - 1;
        while (l < hgh) {  // first bucket whose last value is not below the key
            const int mid = (l + hgh) >> 1;
            if (kk > __ldcg(coarse + mid)) l = mid + 1; else hgh = mid;
        }
        hgh = min(cnt - 1, (l + 1) * kCoarseStep - 1);
        l = l * kCoarseStep;
        const unsigned long long* rs = sh.cdfseg[q];
        while (l < hgh) {
            const int mid = (l + hgh) >> 1;
            if (kk > __ldcg(rs + mid)) l = mid + 1; else hgh = mid;
        }
        const int parent = q * cnt + l;
        a.parents[m0] = parent;
        a.pose_out[m0] = a.poses_in.at(parent);
        a.w_out[m0] = __ldcg(sh.w[q] + parent);
        a.lw_out[m0] = __ldcg(sh.lw[q] + parent);
    }
    // where the strongest particle of the update lives now: its first child (every rank computes the same number)
    if (blockIdx.x == 0 && tid == 0) {
        const int sb = st->strongest;
        int first = 0;
        if (sb > 0) {
            const unsigned long long cprev = cdf_of(sb - 1);
            int l = 0, hgh = P;  // smallest m with key(m) > cprev (key is non-decreasing in m)
            while (l < hgh) {
                const int mid = (l + hgh) >> 1;
                if (key_of(mid) > cprev) hgh = mid; else l = mid + 1;
            }
            first = l;
        }
        const bool has_child = first < P && (sb == P - 1 || !(key_of(first) > cdf_of(sb)));
        st->strongest_now = has_child ? first : -1;
    }
}

// getters on a sharded handle: copy the other ranks' blocks out of their owners' arrays (peer mappings)
struct RemoteBlocks {
    PoseTable poses;
    const double* w[kMaxRanks];
    const double* lw[kMaxRanks];
    const int* parents[kMaxRanks];
    int cnt, lo, P;
};
__global__ void k_fill_remote_blocks(RemoteBlocks rb, float4* __restrict__ pose, double* __restrict__ w,
                                     double* __restrict__ lw, int* __restrict__ parents) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rb.P || (i >= rb.lo && i < rb.lo + rb.cnt)) return;
    const int q = i / rb.cnt;
    pose[i] = rb.poses.p[q][i];
    w[i] = rb.w[q][i];
    lw[i] = rb.lw[q][i];
    parents[i] = rb.parents[q][i];
}

// Per-particle maps: slot assignment.  parents[] is non-decreasing, so the first child of a parent is
// where parents[m] != parents[m-1]; it keeps the parent's slot (no copy).  Every further child
// ("duplicate") takes, in order, the slot of a parent that has no child at all.  One CTA; two
// exclusive scans (duplicates, dead parents).
__global__ void __launch_bounds__(1024) k_assign_slots(const int* __restrict__ parents, int P,
                                                       const int* __restrict__ slot_in, int* __restrict__ slot_out,
                                                       int* __restrict__ dup_src, int* __restrict__ dup_dst,
                                                       int4* __restrict__ dup_rect, int4* __restrict__ rect,
                                                       int* __restrict__ scratch /* 2P */, Stats* __restrict__ st,
                                                       Geometry g) {
    __shared__ int s_a[1024], s_b[1024];
    const int tid = threadIdx.x;
    const int per = (P + 1023) / 1024;
    const int i0 = tid * per, i1 = min(P, i0 + per);
    int* used = scratch;
    int* free_slots = scratch + P;
    for (int i = i0; i < i1; i++) used[i] = 0;
    __syncthreads();
    for (int m = i0; m < i1; m++) used[parents[m]] = 1;  // benign race: all writers store 1
    __syncthreads();
    int nd = 0, nf = 0;
    for (int i = i0; i < i1; i++) {
        nd += (i > 0 && parents[i] == parents[i - 1]) ? 1 : 0;
        nf += used[i] ? 0 : 1;
    }
    s_a[tid] = nd;
    s_b[tid] = nf;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int va = tid >= o ? s_a[tid - o] : 0, vb = tid >= o ? s_b[tid - o] : 0;
        __syncthreads();
        s_a[tid] += va;
        s_b[tid] += vb;
        __syncthreads();
    }
    int rd = s_a[tid] - nd, rf = s_b[tid] - nf;
    for (int i = i0; i < i1; i++)
        if (!used[i]) free_slots[rf++] = slot_in[i];
    __syncthreads();
    for (int m = i0; m < i1; m++) {
        const int p = parents[m];
        if (m > 0 && p == parents[m - 1]) {
            const int d = free_slots[rd], sp = slot_in[p];
            dup_src[rd] = sp;
            dup_dst[rd] = d;
            slot_out[m] = d;
            // Outside the explored boxes both maps are blank (identical), so only the union of the two
            // boxes has to move; the child inherits the parent's box.  Source slots are never destinations
            // (a parent with a child is not dead).
            const int4 a = rect[sp], b = rect[d];
            dup_rect[rd] = make_int4(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
            rect[d] = a;
            rd++;
        } else {
            slot_out[m] = slot_in[p];
        }
    }
    if (tid == 1023) st->num_dup = s_a[1023];
}

// Per-particle maps sharded over R ranks.  Particle m lives on rank m / cnt, in slot
// gslot[m] of that rank's arena of S = 2*cnt slots.  Every rank runs this kernel on the same inputs and
// obtains the same global table; it records copy jobs only for its own particles.
//   keep : the first child that a parent has ON ITS OWN RANK stays in the parent's slot (no copy)
//   need : every other child takes, in order, a slot of its rank that the OLD generation does not occupy
//          (there are >= cnt of those), and is filled by a local copy or a pull from the parent's rank.
// No slot of the old generation is written during the exchange, so concurrent pulls by other ranks read
// consistent maps; the caller's barrier after the pulls releases the old slots.
__device__ __forceinline__ int block_excl_scan_1024(int v, int* s_buf, int* total) {
    const int tid = threadIdx.x;
    __syncthreads();
    s_buf[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int u = tid >= o ? s_buf[tid - o] : 0;
        __syncthreads();
        s_buf[tid] += u;
        __syncthreads();
    }
    if (total) *total = s_buf[1023];
    return s_buf[tid] - v;
}

__global__ void __launch_bounds__(1024) k_assign_slots_mr(const int* __restrict__ parents, int P, int cnt, int R, int S,
                                                          int myrank, const int* __restrict__ gslot_in,
                                                          int* __restrict__ gslot_out, int* __restrict__ job_src_rank,
                                                          int* __restrict__ job_src_slot, int* __restrict__ job_dst,
                                                          int* __restrict__ job_level, int* __restrict__ scratch /* 2*R*S */,
                                                          Stats* __restrict__ st) {
    __shared__ int s_buf[1024];
    const int tid = threadIdx.x;
    int* occ = scratch;
    int* freelist = scratch + R * S;
    // one CTA per rank q: the ranks' tables are independent of each other (a single CTA walking all R ranks cost
    // ~40 us of the 8-GPU resampling)
    {
        const int q = blockIdx.x;
        for (int i = tid; i < S; i += 1024) occ[q * S + i] = 0;
        __syncthreads();
        for (int p = q * cnt + tid; p < (q + 1) * cnt; p += 1024) occ[q * S + gslot_in[p]] = 1;
        __syncthreads();
        const int per_s = (S + 1023) / 1024;
        const int s0 = min(S, tid * per_s), s1 = min(S, s0 + per_s);
        int nf = 0;
        for (int x = s0; x < s1; x++) nf += occ[q * S + x] ? 0 : 1;
        int rf = block_excl_scan_1024(nf, s_buf, nullptr);
        for (int x = s0; x < s1; x++)
            if (!occ[q * S + x]) freelist[q * S + rf++] = x;
        const int per_m = (cnt + 1023) / 1024;
        const int m0 = q * cnt + min(cnt, tid * per_m), m1 = min((q + 1) * cnt, m0 + per_m);
        int nn = 0;
        for (int m = m0; m < m1; m++) {
            const int p = parents[m];
            const bool keep = (p / cnt == q) && (m == q * cnt || parents[m - 1] != p);
            nn += keep ? 0 : 1;
        }
        int total = 0;
        int rn = block_excl_scan_1024(nn, s_buf, &total);  // also orders the freelist writes before the reads
        for (int m = m0; m < m1; m++) {
            const int p = parents[m];
            const bool keep = (p / cnt == q) && (m == q * cnt || parents[m - 1] != p);
            if (keep) {
                gslot_out[m] = gslot_in[p];
            } else {
                const int d = freelist[q * S + rn];
                gslot_out[m] = d;
                if (q == myrank) {
                    // A remote parent is pulled over NVLink ONCE per rank (by its first child here, level 0);
                    // its further children on this rank copy that local replica afterwards (level 1).  Without
                    // this a heavy parent's rank serves every one of its children: an NVLink hot spot.
                    const bool remote = p / cnt != q;
                    const bool first_here = m == q * cnt || parents[m - 1] != p;
                    if (remote && !first_here) {
                        int mf = m;  // first child of p on this rank, and the slot it was given
                        while (mf > q * cnt && parents[mf - 1] == p) mf--;
                        job_src_rank[rn] = q;
                        job_src_slot[rn] = freelist[q * S + rn - (m - mf)];
                        job_level[rn] = 1;
                    } else {
                        job_src_rank[rn] = p / cnt;
                        job_src_slot[rn] = gslot_in[p];
                        job_level[rn] = 0;
                    }
                    job_dst[rn] = d;
                }
                rn++;
            }
        }
        if (q == myrank && tid == 0) st->num_dup = total;
    }
}

// copy rectangle of every job (union of the two explored boxes); the child inherits the parent's box.  The
// parent's box is read from the parent's rank.
__global__ void __launch_bounds__(256) k_job_rects(const int* __restrict__ job_src_rank,
                                                   const int* __restrict__ job_src_slot,
                                                   const int* __restrict__ job_dst, const int* __restrict__ job_level,
                                                   int level, int4* __restrict__ job_rect, int4* __restrict__ rect,
                                                   const Stats* __restrict__ st, PeerTable peers, Geometry g) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= st->num_dup || job_level[k] != level) return;
    const int4 a = peers.rect[job_src_rank[k]][job_src_slot[k]];
    const int d = job_dst[k];
    const int4 b = rect[d];
    job_rect[k] = make_int4(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
    rect[d] = a;
}

// GridMap.createMapData(other) GridMap.java:118-124 copies logData and likelihoodData of the parent.  Here a map IS
// its counter pairs (logData in closed form; the field is evaluated on demand from them), so the copy moves 8 bytes
// per cell — restricted to the rectangle k_assign_slots computed (identical result, see there).
// grid = chunks_per_map * max_dups; CTAs beyond num_dup exit.  Rows are moved as 16-byte vectors, 8 per thread in flight.
// With `dup_src_rank` the source slot lives in another rank's arena: the same kernel then PULLS the rows
// over NVLink through the peer mappings of PeerTable (cudaIpc), one 16-byte load per lane.
__global__ void __launch_bounds__(256) k_copy_maps(CellCounts* __restrict__ counts, const int* __restrict__ dup_src,
                                                   const int* __restrict__ dup_dst, const int4* __restrict__ dup_rect,
                                                   const Stats* __restrict__ st, size_t cells, int W,
                                                   int chunks_per_map, const int* __restrict__ dup_src_rank,
                                                   PeerTable peers, const int* __restrict__ job_level, int level,
                                                   int skip_rank /* jobs whose source is this rank are left to
                                                                    k_copy_maps_bulk; -1: none */) {
    const int k = blockIdx.x / chunks_per_map;
    if (k >= st->num_dup) return;
    if (job_level && job_level[k] != level) return;
    if (dup_src_rank && dup_src_rank[k] == skip_rank) return;
    const int chunk = blockIdx.x - k * chunks_per_map;
    const int src = dup_src[k], dst = dup_dst[k];
    const CellCounts* src_counts = dup_src_rank ? peers.counts[dup_src_rank[k]] : counts;
    const int4 r = dup_rect[k];
    if (r.x > r.z || r.y > r.w) return;
    const int rows = r.w - r.y + 1;
    const int per = (rows + chunks_per_map - 1) / chunks_per_map;
    const int y0 = r.y + chunk * per, y1 = min(r.w + 1, y0 + per);
    const CellCounts* cs = src_counts + (size_t)src * cells;
    CellCounts* cd = counts + (size_t)dst * cells;
    if (((cells | (size_t)W) & 1) == 0) {  // even row length and slot size: rows start 16-byte aligned
        const int x0 = r.x & ~1, n2 = ((r.z | 1) - x0 + 1) / 2;  // pairs of cells per row
        const uint4* cs4 = reinterpret_cast<const uint4*>(cs);
        uint4* cd4 = reinterpret_cast<uint4*>(cd);
        const int total = (y1 - y0) * n2;  // (row, pair) flattened: 8 independent 16-byte loads in flight per thread
        for (int e0 = threadIdx.x; e0 < total; e0 += 8 * 256) {
            size_t o[8];
            uint4 a[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = e0 + u * 256;
                const int yy = e / n2, xx = e - yy * n2;
                o[u] = ((size_t)(y0 + yy) * W + x0) / 2 + xx;
                if (e < total) a[u] = cs4[o[u]];  // plain loads: many children read the same parent (L2 hits)
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (e0 + u * 256 < total) __stcs(cd4 + o[u], a[u]);  // written once, read next step at the earliest
        }
    } else {
        const int n = r.z - r.x + 1;
        for (int y = y0; y < y1; y++) {
            const size_t o = (size_t)y * W + r.x;
            for (int i = threadIdx.x; i < n; i += 256) cd[o + i] = cs[o + i];
        }
    }
}

// The same copy on the TMA engine (source on this rank, 16-byte aligned rows): ONE thread per CTA moves row segments of up
// to kCpSeg bytes global -> shared (cp.async.bulk + mbarrier::complete_tx) -> global (cp.async.bulk.global.shared +
// bulk groups) through a ring of kCpStages buffers, kCpLead loads ahead of the stores.  No per-element index
// arithmetic, no registers holding data: the per-thread version above spends 52 % of its issue slots on addresses
// (ncu r02j) and reaches 3.9 TB/s of writes.
constexpr int kCpStages = 8, kCpLead = 6, kCpSeg = 4096;
__global__ void __launch_bounds__(32) k_copy_maps_bulk(CellCounts* __restrict__ counts, const int* __restrict__ dup_src,
                                                       const int* __restrict__ dup_dst,
                                                       const int4* __restrict__ dup_rect,
                                                       const Stats* __restrict__ st, size_t cells, int W,
                                                       int chunks_per_map, const int* __restrict__ dup_src_rank,
                                                       int myrank, const int* __restrict__ job_level, int level) {
    extern __shared__ __align__(128) unsigned char s_ring[];  // kCpStages * kCpSeg
    __shared__ __align__(8) uint64_t s_bar[kCpStages];
    const int k = blockIdx.x / chunks_per_map;
    if (k >= st->num_dup || threadIdx.x != 0) return;
    // multi-rank: only the jobs of this level whose source map lives on this rank (pulls stay with k_copy_maps)
    if (job_level && (job_level[k] != level || dup_src_rank[k] != myrank)) return;
    const int chunk = blockIdx.x - k * chunks_per_map;
    const int4 r = dup_rect[k];
    if (r.x > r.z || r.y > r.w) return;
    const int rows = r.w - r.y + 1;
    const int per = (rows + chunks_per_map - 1) / chunks_per_map;
    const int y0 = r.y + chunk * per, y1 = min(r.w + 1, y0 + per);
    if (y0 >= y1) return;
    const int x0 = r.x & ~1;
    const unsigned rowbytes = (unsigned)(((r.z | 1) - x0 + 1) * (int)sizeof(CellCounts));
    const int segs = (int)((rowbytes + kCpSeg - 1) / kCpSeg);
    const int total = (y1 - y0) * segs;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(counts + (size_t)dup_src[k] * cells);
    unsigned char* dst = reinterpret_cast<unsigned char*>(counts + (size_t)dup_dst[k] * cells);
    for (int i = 0; i < kCpStages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto seg_of = [&](int i, size_t& off, unsigned& bytes) {
        const int yy = i / segs, sg = i - yy * segs;
        off = ((size_t)(y0 + yy) * W + x0) * sizeof(CellCounts) + (size_t)sg * kCpSeg;
        bytes = min((unsigned)kCpSeg, rowbytes - (unsigned)sg * kCpSeg);
    };
    for (int i = 0; i < total + kCpLead; i++) {
        if (i < total) {
            const int sgi = i % kCpStages;
            // the store that last read this stage (iteration i - kCpStages) must have finished reading shared memory
            if (i >= kCpStages) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kCpStages - 1 - kCpLead) : "memory");
            size_t off; unsigned bytes;
            seg_of(i, off, bytes);
            const uint32_t bar = smem_u32(&s_bar[sgi]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(s_ring + sgi * kCpSeg)), "l"(src + off), "r"(bytes), "r"(bar) : "memory");
        }
        const int j = i - kCpLead;
        if (j >= 0) {
            const int sgj = j % kCpStages;
            const uint32_t bar = smem_u32(&s_bar[sgj]), phase = (uint32_t)(j / kCpStages) & 1u;
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            size_t off; unsigned bytes;
            seg_of(j, off, bytes);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst + off), "r"(smem_u32(s_ring + sgj * kCpSeg)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the last store's read
}

// ------------------------------------------------------------------------------------------------
// rows adjacent to the path
// ------------------------------------------------------------------------------------------------
// GridMapApp.onHandleData GridMapApp.java:140-175 + Measurement(double x, double y, boolean, int)
// Observation.java:69-76.  MathUtil.cos/sin(double) = FastMath (commons-math3) -> cos_fixed / sin_fixed
// (device_math.cuh): the same operation sequence as the oracle's, so the de-skewed beams are bit-identical.
__global__ void __launch_bounds__(256) k_deskew(const double* __restrict__ angle, const double* __restrict__ dist,
                                                int n, double d_center, double d_theta, double2* __restrict__ out_xy,
                                                double* __restrict__ out_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d_i = (double)(-(n - i)) / (double)n;
    const double delta_theta = d_theta * d_i;
    const double delta_x = d_center * d_i;
    const double a = angle[i] + delta_theta;
    const double x_a = dist[i] * cos_fixed(a) + delta_x;
    const double y_a = dist[i] * sin_fixed(a);
    out_xy[i] = make_double2(x_a, y_a);
    out_dist[i] = sqrt(x_a * x_a + y_a * y_a);
}

// Util.invLogOdds(double) Util.java:46-48: 1.0f - 1.0f / (1 + Math.exp(log))
__device__ __forceinline__ double inv_log_odds(double l) { return 1.0 - 1.0 / (1.0 + exp(l)); }

// GridMap.render GridMap.java:371-388: value -> LUT index (int)(value * 255) -> gray (int)(255 * (i / 256f))
__global__ void __launch_bounds__(256) k_render(const CellCounts* __restrict__ counts, const double* __restrict__ lik,
                                                size_t n, int likelihood, double l_free, double l_occ,
                                                uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float value;
    if (likelihood) value = (float)lik[i];
    else {
        const double l = (double)counts[i].n_free * l_free + (double)counts[i].n_occ * l_occ;
        value = (float)(1.0 - inv_log_odds(l));
    }
    int idx = (int)(value * 255.0f);
    idx = min(max(idx, 0), 255);  // Java would throw outside [0, 255]; values are probabilities
    const float ratio = (float)idx / 256.0f;
    const uint32_t c = (uint32_t)(int)(255.0f * ratio);
    out[i] = ((255u << 24) | (c << 16) | (c << 8) | c) & 0xfeffffffu;  // Color.colorToFloatBits
}

// GridMapApp.calculateCombined GridMapApp.java:439-458 in two steps, so that ranks can multiply their partial
// products in between (multi-rank: one PRODUCT all-reduce of W*H doubles):
//   k_combine_product: prod_p (1 - invLogOdds(logData_p)) over the handle's local particles, in particle order
//   k_combine_finish : logOdds(1 - product) and the pseudo-counts whose sign equals the combined log-odds' sign,
//                      so the likelihood tile kernel can blur it (GridMap.computeLikelihoodMap(combinedGrid))
__global__ void __launch_bounds__(256) k_combine_product(const CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                         int n_local, size_t cells, double l_free, double l_occ,
                                                         double* __restrict__ prod_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    double product = 1.0;
    for (int p = 0; p < n_local; p++) {
        const CellCounts c = counts[(size_t)slot[p] * cells + i];
        const double l = (double)c.n_free * l_free + (double)c.n_occ * l_occ;
        product *= 1.0 - inv_log_odds(l);
    }
    prod_out[i] = product;
}
__global__ void __launch_bounds__(256) k_combine_finish(const double* __restrict__ prod, size_t cells,
                                                        double* __restrict__ log_out, CellCounts* __restrict__ sign_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const double odds = 1.0 - prod[i];
    const double v = log(odds / (1.0 - odds));  // Util.logOdds(double)
    log_out[i] = v;
    sign_out[i] = v > 0.0 ? CellCounts{0u, 1u} : (v < 0.0 ? CellCounts{1u, 0u} : CellCounts{0u, 0u});
}

// ---- small utilities ----
// mark every tile of `nslots` slots dirty (bits beyond the last tile stay clear)
__global__ void k_fill_dirty(uint32_t* dirty, int nslots, int tile_words, int ntiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots * tile_words) return;
    const int wi = i % tile_words;
    const int left = ntiles - wi * 32;
    dirty[i] = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
}
__global__ void k_fill_f64(double* a, size_t n, double v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void k_fill_rect(int4* rect, int S, int4 v) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) rect[s] = v;
}
__global__ void k_init_particles(float4* pose, double* w, double* lw, int* parents, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    pose[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    w[i] = 1.0 / (double)P;  // SLAM.java:71
    lw[i] = 0.0;
    parents[i] = i;
}
__global__ void k_iota_mod(int* a, int n, int mod) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i % mod;
}
__global__ void k_counts_to_log(const CellCounts* __restrict__ c, double* __restrict__ out, size_t n, double l_free,
                                double l_occ) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)c[i].n_free * l_free + (double)c[i].n_occ * l_occ;
}
__global__ void k_counts_split(const CellCounts* __restrict__ c, uint32_t* __restrict__ out, size_t n, int which) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = which ? c[i].n_occ : c[i].n_free;
}
__global__ void k_counts_join(CellCounts* __restrict__ c, const uint32_t* __restrict__ nf,
                              const uint32_t* __restrict__ no, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = CellCounts{nf[i], no[i]};
}
__global__ void k_pose_pack(const float* __restrict__ xyt, float4* __restrict__ pose, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) pose[i] = make_float4(xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2], 0.f);
}
// multi-rank peer exchange: poses are not replicated by the step; a getter that needs all of them first copies
// the remote blocks out of their owners' pose arrays (peer mappings)
__global__ void k_pose_fill_remote(PoseTable poses, float4* __restrict__ local, int lo, int cnt, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P && (i < lo || i >= lo + cnt)) local[i] = poses.at(i);
}
// The parents the last resampling selected, as a list of LOCAL particle indices (every rank holds the whole
// parents[] of a per-particle-map resampling; a parent counts if any rank's child descends from it).  parents[] is
// non-decreasing (systematic resampling, SLAM.java:139-150; the identity when the step did not resample), so a
// parent's first child is where the value changes: each parent is listed exactly once, in no particular order
// (the integration adds integers).
__global__ void k_list_parents(const int* __restrict__ parents, int P, int lo, int cnt, int* __restrict__ ulist,
                               int* __restrict__ n_used) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P) return;
    const int p = parents[m];
    if ((m == 0 || parents[m - 1] != p) && p >= lo && p < lo + cnt) ulist[atomicAdd(n_used, 1)] = p - lo;
}
// Per-particle maps across ranks: the maps a resampling copies are final only once their owners have integrated
// the scan into them (deferred until the parents are known, see k_map_update_red).  One CTA: raise my flag on every
// rank (after the integration kernel, in stream order), then wait for every rank's flag — the copy kernels behind
// this one in the stream may pull from any rank.
__global__ void k_maps_final(Shard sh, Stats* st) {
    if ((int)threadIdx.x < sh.nranks) shard_signal(sh, 0);
    shard_wait(sh, 0, st);
}
__global__ void k_gather_pose(const int* __restrict__ parents, const float4* __restrict__ in, float4* __restrict__ out, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) out[i] = in[parents[i]];
}
// GridMap.applyMeasurement on one per-particle map (no dirty-tile bookkeeping)
__global__ void k_apply_one_red(CellCounts* __restrict__ counts, int4* __restrict__ rect, float sx, float sy, float ex,
                                float ey, float meas, int was_hit, Geometry g) {
    CellBox box;
    apply_measurement_red<false>(counts, g, sx, sy, ex, ey, meas, was_hit != 0, box);
    if (box.x1 >= 0) *rect = make_int4(min(rect->x, box.x0), min(rect->y, box.y0), max(rect->z, box.x1), max(rect->w, box.y1));
}
__global__ void k_pose_unpack(const float4* __restrict__ pose, float* __restrict__ xyt, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        const float4 p = pose[i];
        xyt[3 * i] = p.x; xyt[3 * i + 1] = p.y; xyt[3 * i + 2] = p.z;
    }
}

}  // namespace gms
