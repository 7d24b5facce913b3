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
#include <type_traits>
#include "device_math.cuh"

namespace gms {

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
    int xerror;           // fused exchange: a peer's records did not arrive in time
    int pad[2];
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
    unsigned long long* fx;  // sum of trunc(w * 2^60) of the tile (feeds k_cdf_fixed)
    unsigned* counter;  // last-block-done ticket
};
#define kNegInf (__longlong_as_double((long long)0xfff0000000000000ULL))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kMaxRanks = 16;
struct PeerTable {  // per-particle maps across ranks: every rank's arenas, mapped into this process (cudaIpc)
    const CellCounts* counts[kMaxRanks];
    const double* lik[kMaxRanks];
    const int4* rect[kMaxRanks];
    const uint32_t* dirty[kMaxRanks];
};

// Peer exchange (multi-rank, replaces the NCCL all-gather): after scoring, k_xpush stores this rank's block of
// {log-weight, pose} straight into EVERY rank's receive buffer (peer-mapped, NVLink) with coalesced 8/16-byte
// stores; a flag per (receiver, sender) carries the step sequence number (k_xsignal / k_xwait).  Receive
// buffer layout: f64 lw[P] followed by float4 pose[P].  (Pushing the 24-byte records from inside the scoring
// kernel was tried first: fine on 2-4 GPUs, but 7 x 100k scattered 24-byte NVLink writes per rank doubled the
// scoring time on 8.)
struct XPush {
    int nranks;
    unsigned char* dst[kMaxRanks];
};
__global__ void __launch_bounds__(256) k_xpush(const double* __restrict__ lw, const float4* __restrict__ pose, int lo,
                                               int cnt, int P, XPush xp) {
    const long long total = (long long)xp.nranks * cnt;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(e / cnt), i = lo + (int)(e - (long long)q * cnt);
        double* dl = reinterpret_cast<double*>(xp.dst[q]);
        float4* dp = reinterpret_cast<float4*>(dl + P);
        dl[i] = lw[i];
        dp[i] = pose[i];
    }
}
__global__ void __launch_bounds__(256) k_import_soa(const unsigned char* __restrict__ xg, int P, int lo, int cnt,
                                                    double* __restrict__ lw, float4* __restrict__ pose) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || (i >= lo && i < lo + cnt)) return;  // the local block is already in place
    const double* sl = reinterpret_cast<const double*>(xg);
    const float4* sp = reinterpret_cast<const float4*>(sl + P);
    lw[i] = sl[i];
    pose[i] = sp[i];
}
struct XFlags {
    unsigned long long* flag[kMaxRanks];  // flag[q] = rank q's flag array (one u64 per sender)
};
__global__ void k_xsignal(XFlags f, int R, int myrank, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q < R) {
        __threadfence_system();  // the records written by the preceding kernel are visible before the flag
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.flag[q] + myrank), "l"(seq) : "memory");
    }
}
__global__ void k_xwait(const unsigned long long* __restrict__ my_flags, int R, unsigned long long seq,
                        Stats* __restrict__ st) {
    const int q = threadIdx.x;
    if (q >= R) return;
    unsigned long long v = 0;
    for (long long spins = 0; spins < 8000000; spins++) {  // ~4 s with the sleeps: never hang the device
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + q) : "memory");
        if (v >= seq) return;
        __nanosleep(500);
    }
    st->xerror = 1;
}

// ------------------------------------------------------------------------------------------------
// beam table: compaction of the hit beams (scoring reads only those, GridMap.java:269-270) and the
// per-beam measured distance in cells, (float) m.distance / resolution (GridMap.java:188).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_beams(const double2* __restrict__ in_xy,
                                                    const double* __restrict__ in_dist,
                                                    const uint8_t* __restrict__ in_hit, int B, float res_f,
                                                    double2* __restrict__ hit_xy, float* __restrict__ meas,
                                                    double2* __restrict__ all_xy, uint8_t* __restrict__ all_hit,
                                                    Stats* __restrict__ st) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + tid;
        const bool hit = b < B && in_hit[b] != 0;
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
            for (int k = 0; k < 8; k++) t += s_warp[k];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) st->num_hit = s_base;
}

// ------------------------------------------------------------------------------------------------
// A2 — SLAM.sampleMotionModel SLAM.java:155-163 + Odometry.apply Odometry.java:77-96.
// One thread per local particle; z = {z_d, z_theta} injected or Philox(seed, global index, step).
// ------------------------------------------------------------------------------------------------
// 65536 heading buckets of 2*pi/65536 rad: ~30 particles per bucket at 100k particles spread over +-15 deg.
// Same-address atomics with a return value serialise at ~70 ns each (ncu: 8192 buckets made k_motion 23 us,
// 65536 buckets 10 us), so the bucket count is chosen for low contention, not for ordering precision.
constexpr int kSortBins = 65536;
constexpr int kSortCtas = kSortBins / 1024;

__global__ void __launch_bounds__(256) k_motion(float4* __restrict__ pose, int lo, int cnt,
                                                const double* __restrict__ normals, uint64_t seed,
                                                uint64_t step, double d_center, double d_theta, double sd_c,
                                                double sd_t, unsigned* __restrict__ sort_hist,
                                                unsigned* __restrict__ sort_key, unsigned* __restrict__ sort_rank) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= cnt) return;
    const int i = lo + li;
    double zd, zt;
    if (normals) {
        zd = normals[2 * li];
        zt = normals[2 * li + 1];
    } else {
        philox_normals(seed, (uint32_t)i, step, zd, zt);
    }
    const double d = sd_c * zd + d_center;   // NormalDistribution.sample(): sd * z + mean
    const double th = sd_t * zt + d_theta;
    float4 p = pose[i];
    p.z = (float)angle_constrain((double)p.z + th);
    p.x = (float)((double)p.x + (double)cos_f(p.z) * d);
    p.y = (float)((double)p.y + (double)sin_f(p.z) * d);
    pose[i] = p;
    if (sort_hist) {
        // processing order for k_score_sorted (not part of the arithmetic): bucket by heading; the rank
        // inside a bucket is whatever order the atomics land in — any order gives the same weights
        const unsigned b = min(__float2uint_rz((p.z + 3.14159274f) * (kSortBins / 6.28318548f)), (unsigned)kSortBins - 1u);
        sort_key[li] = b;
        sort_rank[li] = atomicAdd(sort_hist + b, 1u);
    }
}

// exclusive scan of the heading histogram: CTA c scans buckets [1024c, 1024c+1024) (coalesced) and publishes
// its total; the scatter adds the totals of the CTAs before it.  Re-zeroes the histogram for the next step.
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned* __restrict__ hist, unsigned* __restrict__ offs,
                                                    unsigned* __restrict__ cta_total) {
    __shared__ unsigned s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i = blockIdx.x * 1024 + tid;
    const unsigned v = hist[i];
    hist[i] = 0u;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    unsigned wv = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, wv, o);
        if (lane >= o) wv += u;
    }
    const unsigned wbase = wid > 0 ? __shfl_sync(0xffffffffu, wv, wid - 1) : 0u;
    offs[i] = wbase + inc - v;
    if (tid == 1023) cta_total[blockIdx.x] = wbase + inc;
}

__global__ void __launch_bounds__(256) k_sort_scatter(const unsigned* __restrict__ offs,
                                                      const unsigned* __restrict__ cta_total,
                                                      const unsigned* __restrict__ key,
                                                      const unsigned* __restrict__ rank, int cnt,
                                                      int* __restrict__ order) {
    __shared__ unsigned s_base[kSortCtas];
    if (threadIdx.x < kSortCtas) {  // exclusive prefix of the 64 CTA totals (two warps, shuffle scan)
        const int lane = threadIdx.x & 31;
        const unsigned t = cta_total[threadIdx.x];
        unsigned inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        s_base[threadIdx.x] = inc - t;
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < kSortCtas) {
        unsigned first_half = s_base[31] + cta_total[31];
        s_base[threadIdx.x] += first_half;
    }
    __syncthreads();
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < cnt) {
        const unsigned k = key[li];
        order[s_base[k >> 10] + offs[k] + rank[li]] = li;
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
                                                   int* __restrict__ word_off, Stats* __restrict__ st,
                                                   int* __restrict__ ray_maxlen) {
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
    if (tid == 1023) {
        st->num_tiles = s_part[1023];
        *ray_maxlen = 0;  // consumed by the previous step's k_ray_apply; re-armed for this step's k_ray_walk
    }
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
                                                    Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = KH ? KH : g.khalf;
    const int tw = KH ? ((kTileW + 2 * KH + 3) & ~3) : kTileW + 2 * k;  // KH: rows padded to 16 bytes
    const int th = kTileH + 2 * k;
    double* s_h = reinterpret_cast<double*>(smem_raw);               // th * kTileW
    float* s_t = reinterpret_cast<float*>(s_h + th * kTileW);        // th * tw
    const int tid = threadIdx.x;
    const int num_tiles = st->num_tiles;
    const size_t cells = (size_t)g.W * g.H;
    double kr[2 * KH + 1];
#pragma unroll
    for (int i = 0; i < 2 * KH + 1; i++) kr[i] = g.kernel[i];
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int2 item = list[t];
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
                            if (fac) fac[(size_t)gx + (size_t)gy * g.W] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
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
                    if (fac) fac[(size_t)gx + (size_t)gy * g.W] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
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
                                                        Geometry g) {
    constexpr int KH = 3;
    constexpr int tw = (kTileW + 2 * KH + 3) & ~3, th = kTileH + 2 * KH;
    extern __shared__ __align__(128) unsigned char smem_tma[];
    __shared__ __align__(8) uint64_t s_bar[2];
    constexpr int kRawBytes = (kTmaTileW * kTmaTileH * 8 + 127) & ~127;
    // two raw buffers: the TMA request of the NEXT tile is in flight while this tile is blurred
    double* s_h = reinterpret_cast<double*>(smem_tma + 2 * kRawBytes);  // th * kTileW
    float* s_t = reinterpret_cast<float*>(s_h + th * kTileW);           // th * tw
    const int tid = threadIdx.x;
    const int num_tiles = st->num_tiles;
    const size_t cells = (size_t)g.W * g.H;
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
        const int2 item = list[t];
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
        const int2 item = list[t];
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
                        if (fac) fac[(size_t)gx + (size_t)gy * g.W] = total == 0.5 ? g.uniform_term : g.z_hit * total + g.random_term;
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
// Each lane multiplies its factors (<= ceil(B/32) of them, each in [0.01, 0.91]: no underflow), takes
// one log, and the 32 logs are summed with a fixed xor-shuffle tree (deterministic).
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_score(const float4* __restrict__ pose, int lo, int cnt,
                                               const double2* __restrict__ hit_xy, const Stats* __restrict__ st,
                                               const double* __restrict__ lik, const int* __restrict__ slot,
                                               double* __restrict__ lw, ExchangeRec* __restrict__ xlocal,
                                               Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    double2* s_xy = reinterpret_cast<double2*>(smem_raw);
    const int nh = st->num_hit;
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
        for (int b0 = 0; b0 < nh; b0 += 128) {
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
        }
        const double l = warp_sum(log(prod));
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
template <int G, int V = 0>
__global__ void __launch_bounds__(128) k_score_sorted(const float4* __restrict__ pose, int lo, int cnt,
                                                      const double2* __restrict__ hit_xy,
                                                      const Stats* __restrict__ st, const double* __restrict__ fac,
                                                      const int* __restrict__ order, double* __restrict__ lw,
                                                      ExchangeRec* __restrict__ xlocal, NormPartials np,
                                                      int emit_partials, Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_red[4];
    __shared__ int s_redi[4];
    double2* s_xy = reinterpret_cast<double2*>(smem_raw);
    const int nh = st->num_hit;
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
    const int t = (blockIdx.x * blockDim.x + tid) / G, gsub = (blockIdx.x * blockDim.x + tid) % G;
    const int li = t < cnt ? (order ? order[t] : t) : -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (li >= 0) p = pose[lo + li];
    const Xform x(p.x, p.y, p.z);
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
    const unsigned uW = (unsigned)g.W, uH = (unsigned)g.H, oob = uW * uH;
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
        if ((unsigned)gx < uW && (unsigned)gy < uH) return __ldg(fac + ((unsigned)gy * uW + (unsigned)gx));
        return 1.0;
    };
    auto peel = [&]() {  // move the exponent of mant into exp2: exact (power-of-two scaling)
        const int hi = __double2hiint(mant);
        const int e = ((hi >> 20) & 0x7ff) - 1023;
        mant = __hiloint2double(hi - (e << 20), __double2loint(mant));
        exp2 += e;
    };
    int b0 = 0, it = 0;
    for (; b0 + 8 * G <= nhe; b0 += 8 * G, it++) {
        unsigned idx[8];
        unsigned bad = 0;
        // branch-free fast path for 8 beams (their dependency chains interleave); beams whose q~ is too
        // close to an integer are only flagged here and redone exactly below.  End points outside the
        // map read the sentinel fac[W*H] == 1.0 (GridMap.java:276: such beams do not multiply).
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
            if constexpr (V == 1) {
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
            idx[u] = (ok && inb) ? gy * uW + gx : oob;
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
    for (int b = b0 + gsub; b < nhe; b += G) mant *= factor_exact(s_xy[b]);
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
    if (!emit_partials) return;
    // epilogue (single-rank shared map): this CTA's (max, first arg-max, sum exp(lw - max)) for k_normalise,
    // which saves the separate pass over lw.  The combination in k_normalise is order-independent.
    const int lane = tid & 31, wid = tid >> 5;
    double best = writer ? l : kNegInf;
    int bi = writer ? lo + li : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_red[wid] = best; s_redi[wid] = bi; }
    __syncthreads();
    best = s_red[0]; bi = s_redi[0];
#pragma unroll
    for (int k = 1; k < 4; k++)
        if (k < (int)(blockDim.x >> 5) && (s_red[k] > best || (s_red[k] == best && s_redi[k] < bi))) { best = s_red[k]; bi = s_redi[k]; }
    double e = writer ? exp(l - best) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    __syncthreads();
    if (lane == 0) s_red[wid] = e;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) sum += s_red[k];
        np.m[blockIdx.x] = best;
        np.idx[blockIdx.x] = bi;
        np.s[blockIdx.x] = sum;
    }
}

// ------------------------------------------------------------------------------------------------
// A8..A11 — GridMap.integrateObservation GridMap.java:173-191 + applyMeasurement :194-228.
// One thread per (particle, beam) ray; per-particle maps, or the shared map from the strongest pose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_map_update(const float4* __restrict__ pose, int lo, int cnt,
                                                    const double2* __restrict__ all_xy,
                                                    const float* __restrict__ meas,
                                                    const uint8_t* __restrict__ hit, int B,
                                                    CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                    int4* __restrict__ rect, uint32_t* __restrict__ dirty,
                                                    const Stats* __restrict__ st, int shared, Geometry g) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = shared ? (long long)B : (long long)cnt * B;
    if (gid >= total) return;
    const int li = shared ? 0 : (int)(gid / B);
    const int b = (int)(gid - (long long)li * B);
    const int s = shared ? 0 : slot[li];
    const float4 p = pose[shared ? st->strongest : lo + li];
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

// Shared map (one scan per step, only B rays): the DDA of a ray is inherently sequential (f32 error
// term, RayIterator.java:112-130), but the per-cell work (sqrt, inverse sensor model, counter update)
// is not.  Pass 1 walks each ray once and records its cells {x | y << 16}; pass 2 classifies and
// accumulates all cells of all rays in parallel.
__global__ void __launch_bounds__(64) k_ray_walk(const float4* __restrict__ pose,
                                                 const double2* __restrict__ all_xy, int B, int Bpad,
                                                 const Stats* __restrict__ st, uint32_t* __restrict__ ray_cells,
                                                 int cap, int* __restrict__ ray_count,
                                                 float2* __restrict__ ray_start, int* __restrict__ ray_maxlen,
                                                 int4* __restrict__ rect, Geometry g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bpad) return;
    if (b >= B) {
        ray_count[b] = 0;
        return;
    }
    // the strongest pose as k_normalise snapshotted it: this kernel may run on a side stream while the next
    // step's motion update already rewrites the pose array
    (void)pose;
    const float4 p = make_float4(st->strongest_pose[0], st->strongest_pose[1], st->strongest_pose[2], 0.f);
    const Xform t(p.x, p.y, p.z);
    const float sx = (float)((t.tx(0.0, 0.0) - g.posx) / g.res);
    const float sy = (float)((t.ty(0.0, 0.0) - g.posy) / g.res);
    const double2 m = all_xy[b];
    const float ex = (float)((t.tx(m.x, m.y) - g.posx) / g.res);
    const float ey = (float)((t.ty(m.x, m.y) - g.posy) / g.res);
    if (b == 0) *ray_start = make_float2(sx, sy);
    RayIter it;
    it.init(sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, g.extra_steps);
    // cell k of ray b lives at ray_cells[k * Bpad + b]: lanes (consecutive rays) advance in lockstep, so
    // every store of the walk and every load of k_ray_apply is one coalesced line
    uint32_t* out = ray_cells + b;
    int c = 0;
    const int fx = it.x, fy = it.y;
    int lx = it.x, ly = it.y;
    while (it.has_next(g.W, g.H) && c < cap) {
        lx = it.x; ly = it.y;
        out[(size_t)c * Bpad] = (uint32_t)lx | ((uint32_t)ly << 16);
        c++;
        it.advance();
    }
    ray_count[b] = c;
    if (c > 0) {
        atomicMax(ray_maxlen, c);
        int* r = reinterpret_cast<int*>(rect);
        atomicMin(r + 0, min(fx, lx));
        atomicMin(r + 1, min(fy, ly));
        atomicMax(r + 2, max(fx, lx));
        atomicMax(r + 3, max(fy, ly));
    }
}

// persistent grid-stride over the (cell k, ray b) pairs (ray_maxlen is re-zeroed by the next k_lik_scan)
__global__ void __launch_bounds__(256) k_ray_apply(const uint32_t* __restrict__ ray_cells, int Bpad,
                                                   const int* __restrict__ ray_count, int* __restrict__ ray_maxlen,
                                                   const float2* __restrict__ ray_start,
                                                   const float* __restrict__ meas, const uint8_t* __restrict__ hit,
                                                   CellCounts* __restrict__ counts, uint32_t* __restrict__ dirty,
                                                   Geometry g) {
    const int maxlen = *ray_maxlen;
    const long long total = (long long)maxlen * Bpad;
    const float2 s = *ray_start;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e / Bpad), b = (int)(e - (long long)k * Bpad);
        if (k >= ray_count[b]) continue;
        const uint32_t cell = ray_cells[e];
        const int cx = (int)(cell & 0xffffu), cy = (int)(cell >> 16);
        const float dX = s.x - ((float)cx + 0.5f);
        const float dY = s.y - ((float)cy + 0.5f);
        const float dist = __fsqrt_rn(dX * dX + dY * dY);
        const int cls = inverse_sensor_class(dist, meas[b], hit[b] != 0, g.tol_half);
        if (cls != 0) bump_cell(counts, dirty, cx, cy, cls, g);
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
struct SumOp { __device__ double operator()(double a, double b) const { return a + b; } };
struct SumU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };

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

// pass 1: per-tile (max, first arg-max, sum exp(lw - tile max))
__global__ void __launch_bounds__(1024) k_softmax_partials(const double* __restrict__ lw, int P, NormPartials np) {
    __shared__ double s_key[32];
    __shared__ int s_idx[32];
    __shared__ double s_d[32];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const double v = i < P ? lw[i] : kNegInf;
    double best = v;
    int bi = i < P ? i : 0x7fffffff;
    block_argmax_1024(best, bi, s_key, s_idx);
    const double e = i < P ? exp(v - best) : 0.0;
    const double sum = block_reduce_1024(e, SumOp(), s_d);
    if (threadIdx.x == 0) {
        np.m[blockIdx.x] = best;
        np.idx[blockIdx.x] = bi;
        np.s[blockIdx.x] = sum;
    }
}

// pass 2: every CTA combines the tile partials in the same fixed order -> (M, first arg-max, S); then
// w_i = exp(lw_i - M) / S for its tile, tile sums; the last CTA to finish folds the tile sums (fixed
// order) into Neff (SLAM.java:180-190: 1 / sum (w / sum w)^2, evaluated as (sum w)^2 / sum w^2) and
// publishes the step's statistics.
__global__ void __launch_bounds__(1024) k_normalise(const double* __restrict__ lw, double* __restrict__ w,
                                                    const float4* __restrict__ pose, int P, int ntiles, int nparts,
                                                    int policy, NormPartials np, Stats* __restrict__ st) {
    __shared__ double s_key[32];
    __shared__ int s_idx[32];
    __shared__ double s_d[32];
    __shared__ unsigned long long s_u[32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    double best = kNegInf;
    int bi = 0x7fffffff;
    for (int c = tid; c < nparts; c += 1024) {
        const double v = np.m[c];
        const int vi = np.idx[c];
        if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
    }
    block_argmax_1024(best, bi, s_key, s_idx);
    double acc = 0.0;
    for (int c = tid; c < nparts; c += 1024) acc += np.s[c] * exp(np.m[c] - best);
    const double S = block_reduce_1024(acc, SumOp(), s_d);
    const int i = blockIdx.x * 1024 + tid;
    double wi = 0.0;
    if (i < P) {
        wi = exp(lw[i] - best) / S;
        w[i] = wi;
    }
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
        const double neff = (a * a) / b;
        st->neff = neff;
        st->lw_max = best;
        st->sum_exp = S;
        st->strongest = bi;
        st->strongest_w = 1.0 / S;
        const float4 p = pose[bi];
        st->strongest_pose[0] = p.x; st->strongest_pose[1] = p.y; st->strongest_pose[2] = p.z;
        st->do_resample = policy == 2 || (policy == 1 && neff < (double)(P / 2));  // GridMapApp.java:185
        *np.counter = 0u;
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
// LITERAL CDF: Java's sequential f64 running sum, c_i = c_{i-1} + w_i in particle order.  One warp:
// coalesced 32-wide loads, the dependent add chain is replayed through shuffles.
__global__ void __launch_bounds__(32) k_cdf_literal(const double* __restrict__ w, int P, double* __restrict__ cdf,
                                                    const Stats* __restrict__ st) {
    if (!st->do_resample) return;
    const int lane = threadIdx.x;
    double c = 0.0;  // 0.0 + w[0] == w[0]: same as Java's c = particles.get(0).weight (SLAM.java:137)
    for (int base = 0; base < P; base += 32) {
        const double v = base + lane < P ? w[base + lane] : 0.0;
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double wj = __shfl_sync(0xffffffffu, v, j);
            c = c + wj;
            if (lane == j) mine = c;
        }
        if (base + lane < P) cdf[base + lane] = mine;
    }
}

// FIXED CDF: u64 fixed point trunc(w * 2^60); integer addition is associative, so a parallel scan equals
// the sequential walk bit for bit on any number of threads / CTAs / ranks.  Tile c (1024 particles) adds
// the fixed-point tile sums of tiles < c (produced by k_normalise / k_neff) to a block-wide scan.
__global__ void __launch_bounds__(1024) k_cdf_fixed(const double* __restrict__ w, int P,
                                                    const unsigned long long* __restrict__ tile_fx,
                                                    unsigned long long* __restrict__ cdf,
                                                    const Stats* __restrict__ st) {
    if (!st->do_resample) return;
    __shared__ unsigned long long s_u[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long acc = 0;
    for (int c = tid; c < (int)blockIdx.x; c += 1024) acc += tile_fx[c];
    const unsigned long long carry = block_reduce_1024(acc, SumU64(), s_u);
    const int i = blockIdx.x * 1024 + tid;
    unsigned long long v = i < P ? (unsigned long long)(w[i] * 0x1p60) : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    __syncthreads();
    if (lane == 31) s_u[wid] = v;
    __syncthreads();
    unsigned long long wsum = s_u[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, wsum, o);
        if (lane >= o) wsum += u;
    }
    const unsigned long long warp_excl = __shfl_sync(0xffffffffu, wsum, max(wid, 1) - 1);
    if (i < P) cdf[i] = carry + (wid > 0 ? warp_excl : 0ull) + v;
}

// index selection: for m = 1..P, U = r + (m-1)*1.0/P, first i with !(U > c_i), clamped to P-1.
// ... and the new generation is gathered right here: copies of the chosen parents (Particle(Particle)
// SLAM.java:41-45: weight and pose are copied; weights are NOT reset to 1/N).
template <bool FIXED>
__global__ void __launch_bounds__(256) k_select(const void* __restrict__ cdf_raw, int P, double u01, uint64_t seed,
                                                uint64_t resample_count, int* __restrict__ parents,
                                                const Stats* __restrict__ st, const float4* __restrict__ pose_in,
                                                const double* __restrict__ w_in, const double* __restrict__ lw_in,
                                                float4* __restrict__ pose_out, double* __restrict__ w_out,
                                                double* __restrict__ lw_out, int m_begin, int m_count) {
    __shared__ unsigned long long s_coarse[2048];
    const int m0 = m_begin + blockIdx.x * blockDim.x + threadIdx.x;  // children [m_begin, m_begin + m_count)
    const bool resample = st->do_resample != 0;  // uniform over the grid
    int stride = 32;
    while ((P + stride - 1) / stride > 2048) stride <<= 1;
    const int ncoarse = (P + stride - 1) / stride;
    if (resample) {
        const unsigned long long* raw = static_cast<const unsigned long long*>(cdf_raw);  // 8-byte keys either way
        for (int j = threadIdx.x; j < ncoarse; j += 256) s_coarse[j] = raw[min(P - 1, (j + 1) * stride - 1)];
        __syncthreads();
    }
    if (m0 >= m_begin + m_count) return;
    if (!resample) {
        parents[m0] = m0;
        pose_out[m0] = pose_in[m0];
        w_out[m0] = w_in[m0];
        lw_out[m0] = lw_in[m0];
        return;
    }
    if (u01 < 0.0) u01 = philox_uniform(seed, resample_count);
    const double r = u01 * 1.0 / (double)P;
    const double U = r + (double)m0 * 1.0 / (double)P;
    // two-level search: every `stride`-th CDF value (<= 2048 of them) is staged in shared memory with one
    // round of independent loads, which replaces the top ~11 dependent global probes of a plain bisection
    using Key = typename std::conditional<FIXED, unsigned long long, double>::type;
    const Key* cdf = static_cast<const Key*>(cdf_raw);
    const Key key = FIXED ? (Key)(unsigned long long)(U * 0x1p60) : (Key)U;
    Key* coarse = reinterpret_cast<Key*>(s_coarse);
    int lo = 0, hi = ncoarse - 1;
    while (lo < hi) {  // first segment whose last CDF value is not below the key
        const int mid = (lo + hi) >> 1;
        if (key > coarse[mid]) lo = mid + 1; else hi = mid;
    }
    hi = min(P - 1, (lo + 1) * stride - 1);
    lo = lo * stride;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key > cdf[mid]) lo = mid + 1; else hi = mid;
    }
    parents[m0] = lo;
    pose_out[m0] = pose_in[lo];
    w_out[m0] = w_in[lo];
    lw_out[m0] = lw_in[lo];
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
            // boxes (+ the blur half-width for the likelihood field) has to move; the child inherits
            // the parent's box.  Source slots are never destinations (a parent with a child is not dead).
            const int4 a = rect[sp], b = rect[d];
            int4 r = make_int4(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
            if (r.x <= r.z && r.y <= r.w)
                r = make_int4(max(r.x - g.khalf, 0), max(r.y - g.khalf, 0), min(r.z + g.khalf, g.W - 1),
                              min(r.w + g.khalf, g.H - 1));
            dup_rect[rd] = r;
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
                                                          int* __restrict__ job_level, uint32_t* __restrict__ dirty,
                                                          int tile_words, int* __restrict__ scratch /* 2*R*S */,
                                                          Stats* __restrict__ st) {
    __shared__ int s_buf[1024];
    const int tid = threadIdx.x;
    int* occ = scratch;
    int* freelist = scratch + R * S;
    for (int i = tid; i < R * S; i += 1024) occ[i] = 0;
    __syncthreads();
    for (int p = tid; p < P; p += 1024) occ[(p / cnt) * S + gslot_in[p]] = 1;
    __syncthreads();
    for (int q = 0; q < R; q++) {
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
    // slots of this rank that the new generation does not occupy: drop their pending likelihood tiles
    __syncthreads();
    for (int i = tid; i < S; i += 1024) occ[i] = 0;
    __syncthreads();
    for (int m = myrank * cnt + tid; m < (myrank + 1) * cnt; m += 1024) occ[gslot_out[m]] = 1;
    __syncthreads();
    for (int i = tid; i < S * tile_words; i += 1024)
        if (!occ[i / tile_words]) dirty[i] = 0u;
}

// copy rectangle of every job (union of the two explored boxes + blur half-width); the child inherits the
// parent's box.  The parent's box is read from the parent's rank.
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
    int4 r = make_int4(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
    if (r.x <= r.z && r.y <= r.w)
        r = make_int4(max(r.x - g.khalf, 0), max(r.y - g.khalf, 0), min(r.z + g.khalf, g.W - 1), min(r.w + g.khalf, g.H - 1));
    job_rect[k] = r;
    rect[d] = a;
}

// GridMap.createMapData(other) GridMap.java:118-124: both arrays of the parent are copied — restricted to
// the rectangle k_assign_slots computed (identical result, see there) — plus the dirty-tile bitmap.
// grid = chunks_per_map * max_dups; CTAs beyond num_dup exit.  Rows are moved as 16-byte vectors.
// With `dup_src_rank` the source slot lives in another rank's arena: the same kernel then PULLS the rows
// over NVLink through the peer mappings of PeerTable (cudaIpc), one 16-byte load per lane.
__global__ void __launch_bounds__(256) k_copy_maps(CellCounts* __restrict__ counts, double* __restrict__ lik,
                                                   uint32_t* __restrict__ dirty, const int* __restrict__ dup_src,
                                                   const int* __restrict__ dup_dst, const int4* __restrict__ dup_rect,
                                                   const Stats* __restrict__ st, size_t cells, int W, int tile_words,
                                                   int chunks_per_map, const int* __restrict__ dup_src_rank,
                                                   PeerTable peers, const int* __restrict__ job_level, int level) {
    const int k = blockIdx.x / chunks_per_map;
    if (k >= st->num_dup) return;
    if (job_level && job_level[k] != level) return;
    const int chunk = blockIdx.x - k * chunks_per_map;
    const int src = dup_src[k], dst = dup_dst[k];
    const CellCounts* src_counts = counts;
    const double* src_lik = lik;
    const uint32_t* src_dirty = dirty;
    if (dup_src_rank) {
        const int q = dup_src_rank[k];
        src_counts = peers.counts[q];
        src_lik = peers.lik[q];
        src_dirty = peers.dirty[q];
    }
    if (chunk == 0)
        for (int i = threadIdx.x; i < tile_words; i += 256)
            dirty[(size_t)dst * tile_words + i] = src_dirty[(size_t)src * tile_words + i];
    const int4 r = dup_rect[k];
    if (r.x > r.z || r.y > r.w) return;
    const int rows = r.w - r.y + 1;
    const int per = (rows + chunks_per_map - 1) / chunks_per_map;
    const int y0 = r.y + chunk * per, y1 = min(r.w + 1, y0 + per);
    const CellCounts* cs = src_counts + (size_t)src * cells;
    CellCounts* cd = counts + (size_t)dst * cells;
    const double* ls = src_lik + (size_t)src * cells;
    double* ld = lik + (size_t)dst * cells;
    if (((cells | (size_t)W) & 1) == 0) {  // even row length and slot size: rows start 16-byte aligned
        const int x0 = r.x & ~1, n2 = ((r.z | 1) - x0 + 1) / 2;  // pairs of cells per row
        const uint4* cs4 = reinterpret_cast<const uint4*>(cs);
        uint4* cd4 = reinterpret_cast<uint4*>(cd);
        const uint4* ls4 = reinterpret_cast<const uint4*>(ls);
        uint4* ld4 = reinterpret_cast<uint4*>(ld);
        const int total = (y1 - y0) * n2;  // (row, pair) flattened: 4 independent 16-byte loads per array in flight
        for (int e0 = threadIdx.x; e0 < total; e0 += 4 * 256) {
            size_t o[4];
            uint4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * 256;
                const int yy = e / n2, xx = e - yy * n2;
                o[u] = ((size_t)(y0 + yy) * W + x0) / 2 + xx;
                if (e < total) { a[u] = cs4[o[u]]; b[u] = ls4[o[u]]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e0 + u * 256 < total) { cd4[o[u]] = a[u]; ld4[o[u]] = b[u]; }
        }
    } else {
        const int n = r.z - r.x + 1;
        for (int y = y0; y < y1; y++) {
            const size_t o = (size_t)y * W + r.x;
            for (int i = threadIdx.x; i < n; i += 256) {
                cd[o + i] = cs[o + i];
                ld[o + i] = ls[o + i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// rows adjacent to the path
// ------------------------------------------------------------------------------------------------
// GridMapApp.onHandleData GridMapApp.java:140-175 + Measurement(double x, double y, boolean, int)
// Observation.java:69-76.  MathUtil.cos/sin(double) = FastMath (commons-math3) -> CUDA cos/sin: f64
// results agree to ~1 ulp, not bit for bit.
__global__ void __launch_bounds__(256) k_deskew(const double* __restrict__ angle, const double* __restrict__ dist,
                                                int n, double d_center, double d_theta, double2* __restrict__ out_xy,
                                                double* __restrict__ out_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d_i = (double)(-(n - i)) / (double)n;
    const double delta_theta = d_theta * d_i;
    const double delta_x = d_center * d_i;
    const double a = angle[i] + delta_theta;
    const double x_a = dist[i] * cos(a) + delta_x;
    const double y_a = dist[i] * sin(a);
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

// GridMapApp.calculateCombined GridMapApp.java:439-458 (product in particle order); also emits the
// pseudo-counts whose sign equals the combined log-odds' sign, so k_likelihood can blur it.
__global__ void __launch_bounds__(256) k_combine(const CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                 int P, size_t cells, double l_free, double l_occ,
                                                 double* __restrict__ log_out, CellCounts* __restrict__ sign_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    double product = 1.0;
    for (int p = 0; p < P; p++) {
        const CellCounts c = counts[(size_t)slot[p] * cells + i];
        const double l = (double)c.n_free * l_free + (double)c.n_occ * l_occ;
        product *= 1.0 - inv_log_odds(l);
    }
    const double odds = 1.0 - product;
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
__global__ void k_pose_unpack(const float4* __restrict__ pose, float* __restrict__ xyt, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        const float4 p = pose[i];
        xyt[3 * i] = p.x; xyt[3 * i + 1] = p.y; xyt[3 * i + 2] = p.z;
    }
}

}  // namespace gms
