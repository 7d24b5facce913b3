// kernels.cuh — the sm_100a kernels of the SLAM hot path.
// No tensor cores: nothing here is a dense contraction.  Data layout (HBM), per handle:
//   pose   float4[P]                 {x, y, theta, 0}      (Pose.java:21-34)
//   lw, w  f64[P]                    ln(product) of the last update / normalised weight
//   counts CellCounts[S][H][W]       8 B per cell, row-major idx = x + y*W (GridMap.java:135)
//   lik    f64[S][H][W]              likelihood field (GridMapData.likelihoodData)
//   rect   int4[S]                   cells modified since the slot's last likelihood rebuild
#pragma once
#include <cooperative_groups.h>

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
    int pad[3];
};

struct ExchangeRec {  // 24 B, gms.h "Exchange record"
    double lw;
    float x, y, t;
    uint32_t pad;
};

// ------------------------------------------------------------------------------------------------
// beam table: compaction of the hit beams (scoring reads only those, GridMap.java:269-270) and the
// per-beam measured distance in cells, (float) m.distance / resolution (GridMap.java:188).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_beams(const double2* __restrict__ in_xy,
                                                    const double* __restrict__ in_dist,
                                                    const uint8_t* __restrict__ in_hit, int B, float res_f,
                                                    double2* __restrict__ hit_xy, float* __restrict__ meas,
                                                    Stats* __restrict__ st) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + tid;
        const bool hit = b < B && in_hit[b] != 0;
        if (b < B) meas[b] = (float)in_dist[b] / res_f;
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
__global__ void __launch_bounds__(256) k_motion(float4* __restrict__ pose, int lo, int cnt,
                                                const double* __restrict__ normals, uint64_t seed,
                                                uint64_t step, double d_center, double d_theta, double sd_c,
                                                double sd_t) {
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
}

// ------------------------------------------------------------------------------------------------
// A3 — GridMap.computeLikelihoodMap GridMap.java:233-250 + Util.doGaussianBlurdSeparable
// Util.java:378-426, restricted to the tiles that intersect each slot's dirty rectangle (+ the
// kernel half-width).  Cells outside keep their value: the field only depends on the sign of the
// log-odds within `khalf` cells, so an untouched neighbourhood reproduces itself bit for bit.
// Same f64 operation order as Java (tap index ascending, mul then add, no FMA) => bit-exact.
// ------------------------------------------------------------------------------------------------
constexpr int kTileW = 64, kTileH = 32;

// One CTA: per-slot tile ranges + exclusive scan -> work list; resets the dirty rectangles.
__global__ void __launch_bounds__(1024) k_lik_worklist(int4* __restrict__ rect, int S, int W, int H, int khalf,
                                                       int4* __restrict__ tile_desc, int* __restrict__ tile_off,
                                                       Stats* __restrict__ st) {
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    const int per = (S + 1023) / 1024;
    const int s0 = tid * per, s1 = min(S, s0 + per);
    int sum = 0;
    for (int s = s0; s < s1; s++) {
        int4 r = rect[s];
        int4 d = make_int4(0, 0, 0, 0);
        if (r.x <= r.z && r.y <= r.w) {
            const int x0 = max(r.x - khalf, 0) / kTileW, x1 = min(r.z + khalf, W - 1) / kTileW;
            const int y0 = max(r.y - khalf, 0) / kTileH, y1 = min(r.w + khalf, H - 1) / kTileH;
            d = make_int4(x0, y0, x1 - x0 + 1, y1 - y0 + 1);
        }
        tile_desc[s] = d;
        sum += d.z * d.w;
        rect[s] = make_int4(0x7fffffff, 0x7fffffff, -1, -1);
    }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - sum;
    for (int s = s0; s < s1; s++) {
        tile_off[s] = run;
        run += tile_desc[s].z * tile_desc[s].w;
    }
    if (tid == 1023) {
        tile_off[S] = s_part[1023];
        st->num_tiles = s_part[1023];
    }
}

// Persistent CTAs walk the work list.  smem: s_t[(TH+2k)][TW+2k] f32 codes {0, .5, 1} (exact in f32),
// s_h[(TH+2k)][TW] f64 horizontal pass.
__global__ void __launch_bounds__(256) k_likelihood(const CellCounts* __restrict__ counts,
                                                    double* __restrict__ lik, const int4* __restrict__ tile_desc,
                                                    const int* __restrict__ tile_off, int S,
                                                    const Stats* __restrict__ st, Geometry g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = g.khalf;
    const int tw = kTileW + 2 * k, th = kTileH + 2 * k;
    double* s_h = reinterpret_cast<double*>(smem_raw);               // th * kTileW
    float* s_t = reinterpret_cast<float*>(s_h + th * kTileW);        // th * tw
    const int tid = threadIdx.x;
    const int num_tiles = st->num_tiles;
    const size_t cells = (size_t)g.W * g.H;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        // slot of tile t: last s with tile_off[s] <= t
        int lo = 0, hi = S - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tile_off[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int4 d = tile_desc[lo];
        const int r = t - tile_off[lo];
        const int ox = (d.x + r % d.z) * kTileW, oy = (d.y + r / d.z) * kTileH;
        const CellCounts* cmap = counts + (size_t)lo * cells;
        double* out = lik + (size_t)lo * cells;
        // 1. threshold the log-odds against logOdds(0.5) == 0.0 (GridMap.java:238-245); cells outside
        //    the map contribute 0 — Java skips those taps, and total + k*0.0 == total.
        for (int e = tid; e < th * tw; e += 256) {
            const int ly = e / tw, lx = e - ly * tw;
            const int gx = ox + lx - k, gy = oy + ly - k;
            float code = 0.0f;
            if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H) {
                const CellCounts c = cmap[(size_t)gx + (size_t)gy * g.W];
                const double v = (double)c.n_free * g.l_free + (double)c.n_occ * g.l_occ;
                code = v > 0.0 ? 1.0f : (v < 0.0 ? 0.0f : 0.5f);
            }
            s_t[e] = code;
        }
        __syncthreads();
        // 2. horizontal pass (Util.java:387-403)
        for (int e = tid; e < th * kTileW; e += 256) {
            const int ly = e / kTileW, lx = e - ly * kTileW;
            const float* row = s_t + ly * tw + lx;
            double total = 0.0;
            for (int i = 0; i < g.ktaps; i++) total += g.kernel[i] * (double)row[i];
            s_h[e] = total;
        }
        __syncthreads();
        // 3. vertical pass (Util.java:409-424)
        for (int e = tid; e < kTileH * kTileW; e += 256) {
            const int ly = e / kTileW, lx = e - ly * kTileW;
            const int gx = ox + lx, gy = oy + ly;
            if (gx < g.W && gy < g.H) {
                const double* col = s_h + ly * kTileW + lx;
                double total = 0.0;
                for (int i = 0; i < g.ktaps; i++) total += g.kernel[i] * col[i * kTileW];
                out[(size_t)gx + (size_t)gy * g.W] = total;
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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// ------------------------------------------------------------------------------------------------
// A8..A11 — GridMap.integrateObservation GridMap.java:173-191 + applyMeasurement :194-228.
// One thread per (particle, beam) ray; per-particle maps, or the shared map from the strongest pose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_map_update(const float4* __restrict__ pose, int lo, int cnt,
                                                    const double2* __restrict__ all_xy,
                                                    const float* __restrict__ meas,
                                                    const uint8_t* __restrict__ hit, int B,
                                                    CellCounts* __restrict__ counts, const int* __restrict__ slot,
                                                    int4* __restrict__ rect, const Stats* __restrict__ st,
                                                    int shared, Geometry g) {
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
    apply_measurement(counts + (size_t)s * ((size_t)g.W * g.H), g, sx, sy, ex, ey, meas[b], hit[b] != 0, box);
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
                                                 const double2* __restrict__ all_xy, int B,
                                                 const Stats* __restrict__ st, uint32_t* __restrict__ ray_cells,
                                                 int cap, int* __restrict__ ray_count,
                                                 float2* __restrict__ ray_start, int4* __restrict__ rect,
                                                 Geometry g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float4 p = pose[st->strongest];
    const Xform t(p.x, p.y, p.z);
    const float sx = (float)((t.tx(0.0, 0.0) - g.posx) / g.res);
    const float sy = (float)((t.ty(0.0, 0.0) - g.posy) / g.res);
    const double2 m = all_xy[b];
    const float ex = (float)((t.tx(m.x, m.y) - g.posx) / g.res);
    const float ey = (float)((t.ty(m.x, m.y) - g.posy) / g.res);
    if (b == 0) *ray_start = make_float2(sx, sy);
    RayIter it;
    it.init(sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, g.extra_steps);
    uint32_t* out = ray_cells + (size_t)b * cap;
    int c = 0;
    int fx = it.x, fy = it.y, lx = it.x, ly = it.y;
    while (it.has_next(g.W, g.H) && c < cap) {
        lx = it.x; ly = it.y;
        out[c++] = (uint32_t)lx | ((uint32_t)ly << 16);
        it.advance();
    }
    ray_count[b] = c;
    if (c > 0) {
        int* r = reinterpret_cast<int*>(rect);
        atomicMin(r + 0, min(fx, lx));
        atomicMin(r + 1, min(fy, ly));
        atomicMax(r + 2, max(fx, lx));
        atomicMax(r + 3, max(fy, ly));
    }
}

__global__ void __launch_bounds__(256) k_ray_apply(const uint32_t* __restrict__ ray_cells, int cap,
                                                   const int* __restrict__ ray_count,
                                                   const float2* __restrict__ ray_start,
                                                   const float* __restrict__ meas, const uint8_t* __restrict__ hit,
                                                   CellCounts* __restrict__ counts, Geometry g) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ray_count[b]) return;
    const uint32_t cell = ray_cells[(size_t)b * cap + k];
    const int cx = (int)(cell & 0xffffu), cy = (int)(cell >> 16);
    const float2 s = *ray_start;
    const float dX = s.x - ((float)cx + 0.5f);
    const float dY = s.y - ((float)cy + 0.5f);
    const float dist = __fsqrt_rn(dX * dX + dY * dY);
    const int cls = inverse_sensor_class(dist, meas[b], hit[b] != 0, g.tol_half);
    if (cls != 0) atomicAdd(reinterpret_cast<uint32_t*>(counts + ((size_t)cx + (size_t)cy * g.W)) + (cls - 1), 1u);
}

// single ray given in grid coordinates (gms_map_apply_measurement)
__global__ void k_apply_one(CellCounts* __restrict__ counts, int4* __restrict__ rect, float sx, float sy,
                            float ex, float ey, float meas, int was_hit, Geometry g) {
    CellBox box;
    apply_measurement(counts, g, sx, sy, ex, ey, meas, was_hit != 0, box);
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
// P <= ~1e6 doubles: latency-bound reductions.  One thread-block CLUSTER of 8 CTAs x 1024 threads;
// CTA partials are exchanged through distributed shared memory (DSMEM) and combined in rank order, so
// every CTA — and every rank of a multi-GPU run — obtains bit-identical totals (fixed tree).
// ------------------------------------------------------------------------------------------------
constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = kClusterCtas * 1024;

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

// block partial -> cluster total (identical in every thread of every CTA of the cluster)
__device__ __forceinline__ double cluster_sum(cg::cluster_group& cl, double v, double* s_buf, double* s_slot) {
    const double b = block_reduce_1024(v, SumOp(), s_buf);
    if (threadIdx.x == 0) *s_slot = b;
    cl.sync();
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < kClusterCtas; r++) t += *cl.map_shared_rank(s_slot, r);
    cl.sync();  // the slot may be rewritten by the next reduction
    return t;
}

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    k_normalise(const double* __restrict__ lw, double* __restrict__ w, const float4* __restrict__ pose, int P,
                int policy, Stats* __restrict__ st) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double s_d[32];
    __shared__ double s_key[32];
    __shared__ int s_idx[32];
    __shared__ double s_slot;
    __shared__ double s_best;
    __shared__ int s_besti;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gt = cl.block_rank() * 1024 + tid;
    // pass 1: max and its FIRST index (strict > keeps the first maximum, SLAM.java:110-115)
    double best = __longlong_as_double(0xfff0000000000000LL);  // -inf
    int bi = 0x7fffffff;
    for (int i = gt; i < P; i += kClusterThreads) {
        const double v = lw[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_key[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    best = s_key[lane]; bi = s_idx[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (tid == 0) { s_best = best; s_besti = bi; }
    cl.sync();
#pragma unroll
    for (int r = 0; r < kClusterCtas; r++) {
        const double ov = *cl.map_shared_rank(&s_best, r);
        const int oi = *cl.map_shared_rank(&s_besti, r);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    // pass 2: e_i = exp(lw_i - max), S = sum e_i
    double acc = 0.0;
    for (int i = gt; i < P; i += kClusterThreads) {
        const double e = exp(lw[i] - best);
        w[i] = e;
        acc += e;
    }
    const double S = cluster_sum(cl, acc, s_d, &s_slot);
    // pass 3: w_i = e_i / S and their sum (SLAM.calculateNeff recomputes it, SLAM.java:181-183)
    acc = 0.0;
    for (int i = gt; i < P; i += kClusterThreads) {
        const double v = w[i] / S;
        w[i] = v;
        acc += v;
    }
    const double ws = cluster_sum(cl, acc, s_d, &s_slot);
    // pass 4: sum (w/ws)^2 (SLAM.java:185-187)
    acc = 0.0;
    for (int i = gt; i < P; i += kClusterThreads) {
        const double v = w[i] / ws;
        acc += v * v;
    }
    const double sq = cluster_sum(cl, acc, s_d, &s_slot);
    if (gt == 0) {
        const double neff = 1.0 / sq;
        st->neff = neff;
        st->lw_max = best;
        st->sum_exp = S;
        st->strongest = bi;
        st->strongest_w = exp(lw[bi] - best) / S;
        const float4 p = pose[bi];
        st->strongest_pose[0] = p.x; st->strongest_pose[1] = p.y; st->strongest_pose[2] = p.z;
        st->do_resample = policy == 2 || (policy == 1 && neff < (double)(P / 2));  // GridMapApp.java:185
    }
}

// SLAM.calculateNeff on the current weights (after set_weights / resample)
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    k_neff(const double* __restrict__ w, int P, Stats* __restrict__ st) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double s_d[32];
    __shared__ double s_slot;
    const int gt = cl.block_rank() * 1024 + threadIdx.x;
    double acc = 0.0;
    for (int i = gt; i < P; i += kClusterThreads) acc += w[i];
    const double ws = cluster_sum(cl, acc, s_d, &s_slot);
    acc = 0.0;
    for (int i = gt; i < P; i += kClusterThreads) {
        const double v = w[i] / ws;
        acc += v * v;
    }
    const double sq = cluster_sum(cl, acc, s_d, &s_slot);
    if (gt == 0) st->neff_query = 1.0 / sq;
}

// SLAM.getWeightedPose SLAM.java:165-178 (plain, not circular, mean of angleConstrain(theta))
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    k_weighted_pose(const double* __restrict__ w, const float4* __restrict__ pose, int P, Stats* __restrict__ st) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double s_d[32];
    __shared__ double s_slot;
    const int gt = cl.block_rank() * 1024 + threadIdx.x;
    double xs = 0, ys = 0, ts = 0, ws = 0;
    for (int i = gt; i < P; i += kClusterThreads) {
        const float4 p = pose[i];
        const double wi = w[i];
        xs += (double)p.x * wi;
        ys += (double)p.y * wi;
        ts += angle_constrain((double)p.z) * wi;
        ws += wi;
    }
    xs = cluster_sum(cl, xs, s_d, &s_slot);
    ys = cluster_sum(cl, ys, s_d, &s_slot);
    ts = cluster_sum(cl, ts, s_d, &s_slot);
    ws = cluster_sum(cl, ws, s_d, &s_slot);
    if (gt == 0) {
        st->weighted_pose[0] = (float)(xs / ws);
        st->weighted_pose[1] = (float)(ys / ws);
        st->weighted_pose[2] = (float)(ts / ws);
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

// FIXED CDF: u64 fixed point trunc(w * 2^60); integer addition is associative, so the cluster-wide
// scan equals the sequential walk bit for bit on any number of threads / CTAs / ranks.  Each CTA of
// the 8-CTA cluster owns a contiguous chunk: chunk totals travel through DSMEM, then a coalesced
// tile-by-tile block scan (warp shuffles) writes the inclusive prefix.
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    k_cdf_fixed(const double* __restrict__ w, int P, unsigned long long* __restrict__ cdf,
                const Stats* __restrict__ st) {
    if (!st->do_resample) return;  // uniform over the cluster
    cg::cluster_group cl = cg::this_cluster();
    __shared__ unsigned long long s_u[32];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)cl.block_rank();
    const int per = (((P + kClusterCtas - 1) / kClusterCtas + 1023) / 1024) * 1024;
    const int i0 = min(P, rank * per), i1 = min(P, i0 + per);
    unsigned long long acc = 0;
    for (int i = i0 + tid; i < i1; i += 1024) acc += (unsigned long long)(w[i] * 0x1p60);
    const unsigned long long tot = block_reduce_1024(acc, SumU64(), s_u);
    if (tid == 0) s_total = tot;
    cl.sync();
    unsigned long long carry = 0;
    for (int r = 0; r < rank; r++) carry += *cl.map_shared_rank(&s_total, r);
    cl.sync();
    for (int base = i0; base < i1; base += 1024) {
        const int i = base + tid;
        unsigned long long v = i < i1 ? (unsigned long long)(w[i] * 0x1p60) : 0ull;
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
        const unsigned long long tile_total = __shfl_sync(0xffffffffu, wsum, 31);
        if (i < i1) cdf[i] = carry + (wid > 0 ? warp_excl : 0ull) + v;
        carry += tile_total;
    }
}

// index selection: for m = 1..P, U = r + (m-1)*1.0/P, first i with !(U > c_i), clamped to P-1.
template <bool FIXED>
__global__ void __launch_bounds__(256) k_select(const void* __restrict__ cdf_raw, int P, double u01, uint64_t seed,
                                                uint64_t resample_count, int* __restrict__ parents,
                                                const Stats* __restrict__ st) {
    const int m0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (m0 >= P) return;
    if (!st->do_resample) {
        parents[m0] = m0;
        return;
    }
    if (u01 < 0.0) u01 = philox_uniform(seed, resample_count);
    const double r = u01 * 1.0 / (double)P;
    const double U = r + (double)m0 * 1.0 / (double)P;
    int lo = 0, hi = P - 1;
    if (FIXED) {
        const unsigned long long* cdf = static_cast<const unsigned long long*>(cdf_raw);
        const unsigned long long Uq = (unsigned long long)(U * 0x1p60);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (Uq > cdf[mid]) lo = mid + 1; else hi = mid;
        }
    } else {
        const double* cdf = static_cast<const double*>(cdf_raw);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (U > cdf[mid]) lo = mid + 1; else hi = mid;
        }
    }
    parents[m0] = lo;
}

// new generation = copies of the chosen parents (Particle(Particle) SLAM.java:41-45: weight and pose
// are copied; weights are NOT reset to 1/N)
__global__ void __launch_bounds__(256) k_gather(const int* __restrict__ parents, int P,
                                                const float4* __restrict__ pose_in, const double* __restrict__ w_in,
                                                const double* __restrict__ lw_in, float4* __restrict__ pose_out,
                                                double* __restrict__ w_out, double* __restrict__ lw_out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P) return;
    const int p = parents[m];
    pose_out[m] = pose_in[p];
    w_out[m] = w_in[p];
    lw_out[m] = lw_in[p];
}

// Per-particle maps: slot assignment.  parents[] is non-decreasing, so the first child of a parent is
// where parents[m] != parents[m-1]; it keeps the parent's slot (no copy).  Every further child
// ("duplicate") takes, in order, the slot of a parent that has no child at all.  One CTA; two
// exclusive scans (duplicates, dead parents).
__global__ void __launch_bounds__(1024) k_assign_slots(const int* __restrict__ parents, int P,
                                                       const int* __restrict__ slot_in, int* __restrict__ slot_out,
                                                       int* __restrict__ dup_src, int* __restrict__ dup_dst,
                                                       int* __restrict__ scratch /* 2P */, Stats* __restrict__ st) {
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
            const int d = free_slots[rd];
            dup_src[rd] = slot_in[p];
            dup_dst[rd] = d;
            slot_out[m] = d;
            rd++;
        } else {
            slot_out[m] = slot_in[p];
        }
    }
    if (tid == 1023) st->num_dup = s_a[1023];
}

// GridMap.createMapData(other) GridMap.java:118-124: both arrays of the parent are copied (+ the dirty
// rectangle that travels with them).  grid = chunks_per_map * max_dups; CTAs beyond num_dup exit.
__global__ void __launch_bounds__(256) k_copy_maps(CellCounts* __restrict__ counts, double* __restrict__ lik,
                                                   int4* __restrict__ rect, const int* __restrict__ dup_src,
                                                   const int* __restrict__ dup_dst, const Stats* __restrict__ st,
                                                   size_t cells, int chunks_per_map) {
    const int k = blockIdx.x / chunks_per_map;
    if (k >= st->num_dup) return;
    const int chunk = blockIdx.x - k * chunks_per_map;
    const int src = dup_src[k], dst = dup_dst[k];
    const CellCounts* cs = counts + (size_t)src * cells;
    CellCounts* cd = counts + (size_t)dst * cells;
    const double* ls = lik + (size_t)src * cells;
    double* ld = lik + (size_t)dst * cells;
    if ((cells & 1) == 0) {  // slot bases are 16-byte aligned: move 2 cells per access
        const size_t n16 = cells / 2;
        const size_t per = (n16 + chunks_per_map - 1) / chunks_per_map;
        const size_t a = (size_t)chunk * per, b = min(n16, a + per);
        const uint4* cs4 = reinterpret_cast<const uint4*>(cs);
        uint4* cd4 = reinterpret_cast<uint4*>(cd);
        const uint4* ls4 = reinterpret_cast<const uint4*>(ls);
        uint4* ld4 = reinterpret_cast<uint4*>(ld);
        for (size_t i = a + threadIdx.x; i < b; i += 256) {
            cd4[i] = cs4[i];
            ld4[i] = ls4[i];
        }
    } else {
        const size_t per = (cells + chunks_per_map - 1) / chunks_per_map;
        const size_t a = (size_t)chunk * per, b = min(cells, a + per);
        for (size_t i = a + threadIdx.x; i < b; i += 256) {
            cd[i] = cs[i];
            ld[i] = ls[i];
        }
    }
    if (chunk == 0 && threadIdx.x == 0) rect[dst] = rect[src];
}

// ---- small utilities ----
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
__global__ void k_iota(int* a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
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
