// device_math.cuh — Java-exact scalar arithmetic shared by every kernel of libgms.
//
// The translation unit is compiled with --fmad=false: Java never contracts a*b+c, and the
// bit-exact claims (ray cells, hit/miss counts, f32 poses, f64 likelihood field) depend on it.
// References are to java/GridMapGL/src/main/java/com/fmsz/gridmapgl/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gms {

constexpr double kPi = 3.141592653589793;       // Math.PI
constexpr double kTwoPi = 6.283185307179586;    // Math.PI * 2

// Constants of one handle, passed by value to kernels (lives in the kernel parameter bank).
struct Geometry {
    int W, H;
    int extra_steps;          // GridMap.java:210
    int ktaps, khalf;         // GridMap.java:94-95
    float res_f;              // GridMap.resolution
    float tol_half;           // hitTolerance / 2 (SensorModel.java:35-37)
    double res, posx, posy;   // (double) resolution / position: the promotions Java performs
    double inv_res;           // 1.0 / res — only for the guarded fast path of cell_of()
    // fixed-point fast path of k_score_sorted: q + fx_magic puts round(q * 2^fx_k) into the low mantissa word
    double fx_magic;          // 1.5 * 2^(52 - fx_k)
    int fx_k;                 // fraction bits: max(W, H) * 2^fx_k < 2^30
    int fx_hi;                // high word of fx_magic == high word of q + fx_magic for 0 <= q < 2^(31 - fx_k)
    int fx_margin;            // acceptance margin in units of 2^-fx_k (>= 8 units >> |q~ - q_java| + rounding)
    int fac_pitch, fac_lp;    // shared map: the factor field is a padded square of side fac_pitch = 2^fac_lp >= max(W, H)
                              // (1.0 outside the map: such end points do not multiply, GridMap.java:276)
    int tiles_x, tiles_y;     // likelihood tiles (kTileW x kTileH cells) per map
    int tile_words;           // 32-bit words of one slot's dirty-tile bitmap
    double z_hit;             // GridMap.java:259
    double uniform_term;      // 1.0 / SENSOR_MAX_RANGE                (GridMap.java:286)
    double random_term;       // zRandom * 1.0 / SENSOR_MAX_RANGE      (GridMap.java:288)
    double l_free, l_occ;     // Util.logOdds(P_FREE / P_OCCUPPIED)    (Util.java:35-37)
    double kernel[31];        // Util.generateGaussianKernel           (Util.java:428-455)
};

// JLS 5.1.3 (int) of a double: NaN -> 0, saturating, truncation toward zero.  cvt.rzi.s32.f64 truncates and
// saturates, but returns INT_MIN for NaN on sm_100 (measured: tests/test_gpu_parity.py::test_degenerate_scans),
// so NaN is handled explicitly.
__device__ __forceinline__ int java_d2i(double d) { return d != d ? 0 : __double2int_rz(d); }

// (int) ((world - position) / resolution) of GridMap.java:273-274 without the f64 division on the fast
// path.  q = t * (1/res) is within 2 ulp of the correctly rounded quotient t / res, i.e. within
// 1.4e-6 for every |q| < 2^31; when q is further than 1e-5 from both neighbouring integers no integer
// lies between q and t / res, so both truncate to the same cell.  Otherwise (probability ~2e-5 per
// lookup; also NaN / saturating inputs) the exact division decides.  Result is bit-identical to Java.
__device__ __forceinline__ int cell_of(double t, double res, double inv_res) {
    const double q = t * inv_res;
    const int n = __double2int_rz(q);
    const double fr = fabs(q - (double)n);  // exact (Sterbenz); in [0, 1) for in-range q
    if (fr > 1e-5 && fr < 1.0 - 1e-5) return n;
    return java_d2i(t / res);
}

// MathUtil.angleConstrain MathUtil.java:65-72.  The loops are replicated literally (they are NOT the
// identity on in-range input: +2pi then -2pi rounds).  Deliberate divergence: |a| > 1e6 or
// non-finite input returns unchanged where Java would spin (or hang on +-inf).
__device__ __forceinline__ double angle_constrain(double a) {
    if (!(fabs(a) <= 1e6)) return a;
    while (a < kPi) a += kTwoPi;
    while (a > kPi) a -= kTwoPi;
    return a;
}

// ---- sin / cos as a FIXED sequence of IEEE f64 operations ------------------------------------------------------
// MathUtil.sin/cos delegate to FastMath (commons-math3 3.6.1, not in the reference tree); FastMath, libm and
// CUDA's sin/cos are all accurate to about an ulp and all different in the last bit now and then.  So that the
// CUDA path and the oracle cannot differ — not even with probability 2^-29 per f32-rounded use, and not in the f64
// de-skew where an ulp can move an end point across a cell border — both evaluate the classic published scheme
// (Sun's fdlibm 5.3: e_rem_pio2.c medium-size reduction by Cody-Waite with a 33+33+53-bit pi/2, k_sin.c / k_cos.c
// minimax polynomials on [-pi/4, pi/4]; < 1 ulp) with the operations in the order written here: +, -, *, rint and
// integer tests only, no fused multiply-add (--fmad=false), so every step is correctly rounded and identical on
// any IEEE machine.  Arguments beyond 2^20 * pi/2 (never an angle on this path) and non-finite ones go to CUDA's
// sin / cos.  The oracle restates the same scheme in C (oracle/gms_ref.c), oracle/pyref.py a third time in Python.
namespace trig {
constexpr double S1 = -0x1.5555555555549p-3, S2 = 0x1.111111110f8a6p-7, S3 = -0x1.a01a019c161d5p-13,
                 S4 = 0x1.71de357b1fe7dp-19, S5 = -0x1.ae5e68a2b9cebp-26, S6 = 0x1.5d93a5acfd57cp-33;
constexpr double C1 = 0x1.555555555554cp-5, C2 = -0x1.6c16c16c15177p-10, C3 = 0x1.a01a019cb1590p-16,
                 C4 = -0x1.27e4f809c52adp-22, C5 = 0x1.1ee9ebdb4b1c4p-29, C6 = -0x1.8fae9be8838d4p-37;
constexpr double kInvPio2 = 0x1.45f306dc9c883p-1;  // 53 bits of 2/pi
constexpr double kPio2_1 = 0x1.921fb54400000p+0, kPio2_1t = 0x1.0b4611a626331p-34;  // first 33 bits of pi/2, rest
constexpr double kPio2_2 = 0x1.0b4611a600000p-34, kPio2_2t = 0x1.3198a2e037073p-69;  // second 33 bits, rest
constexpr double kPio2_3 = 0x1.3198a2e000000p-69, kPio2_3t = 0x1.b839a252049c1p-104;  // third 33 bits, rest
constexpr double kPio4 = 0x1.921fb54442d18p-1;
constexpr double kMedium = 1647099.0;  // < 2^20 * pi/2: fn * kPio2_1 is exact below it

__device__ __forceinline__ double kernel_sin(double x, double y) {  // |x| <= pi/4, y: tail of x
    const double z = x * x, v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
__device__ __forceinline__ double kernel_cos(double x, double y) {
    const double z = x * x;
    double w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}
__device__ __forceinline__ int biased_exponent(double x) { return (__double2hiint(x) >> 20) & 0x7ff; }
// x = n * pi/2 + (y0 + y1), |y0 + y1| <= pi/4 (+ an ulp); returns n
__device__ __forceinline__ int rem_pio2(double x, double& y0, double& y1) {
    if (fabs(x) <= kPio4) { y0 = x; y1 = 0.0; return 0; }
    const double fn = rint(x * kInvPio2);
    double r = x - fn * kPio2_1, w = fn * kPio2_1t;  // good to 85 bits
    y0 = r - w;
    const int ex = biased_exponent(x);
    if (ex - biased_exponent(y0) > 16) {  // cancellation: x is close to a multiple of pi/2 — second round, 118 bits
        double t = r;
        w = fn * kPio2_2;
        r = t - w;
        w = fn * kPio2_2t - ((t - r) - w);
        y0 = r - w;
        if (ex - biased_exponent(y0) > 49) {  // third round, 151 bits: covers every f64 below kMedium
            t = r;
            w = fn * kPio2_3;
            r = t - w;
            w = fn * kPio2_3t - ((t - r) - w);
            y0 = r - w;
        }
    }
    y1 = (r - y0) - w;
    return (int)fn;
}
}  // namespace trig
__device__ __forceinline__ double sin_fixed(double x) {
    if (!(fabs(x) < trig::kMedium)) return sin(x);
    double a, b;
    const int n = trig::rem_pio2(x, a, b) & 3;
    const double v = (n & 1) ? trig::kernel_cos(a, b) : trig::kernel_sin(a, b);
    return (n & 2) ? -v : v;
}
__device__ __forceinline__ double cos_fixed(double x) {
    if (!(fabs(x) < trig::kMedium)) return cos(x);
    double a, b;
    const int n = trig::rem_pio2(x, a, b) & 3;
    const double v = (n & 1) ? trig::kernel_sin(a, b) : trig::kernel_cos(a, b);
    return ((n + 1) & 2) ? -v : v;
}

// MathUtil.cos(float) / sin(float) MathUtil.java:30-40: (float) FastMath.cos((double) radians).
__device__ __forceinline__ float cos_f(float r) { return (float)cos_fixed((double)r); }
__device__ __forceinline__ float sin_f(float r) { return (float)sin_fixed((double)r); }

// Transform.fromRobotToWorld Transform.java:13-32: cos/sin rounded to f32, then widened.
struct Xform {
    double c, s, px, py;
    __device__ __forceinline__ Xform(float x, float y, float theta)
        : c((double)cos_f(theta)), s((double)sin_f(theta)), px((double)x), py((double)y) {}
    __device__ __forceinline__ double tx(double x, double y) const { return x * c - y * s + px; }
    __device__ __forceinline__ double ty(double x, double y) const { return x * s + y * c + py; }
};

// ---- Philox4x32-10: motion noise / resampling uniform when the caller does not inject draws ----
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b, bool centred) {
    uint64_t k = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return centred ? ((double)k + 0.5) * 0x1p-53 : (double)k * 0x1p-53;
}
__device__ __forceinline__ void philox_normals(uint64_t seed, uint32_t gidx, uint64_t step, double& zd,
                                               double& zt) {
    uint32_t c[4] = {gidx, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u53(c[0], c[1], true), u2 = u53(c[2], c[3], false);
    double r = sqrt(-2.0 * log(u1));
    double a = 6.283185307179586 * u2;
    zd = r * cos_fixed(a);
    zt = r * sin_fixed(a);
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t n) {
    uint32_t c[4] = {(uint32_t)n, (uint32_t)(n >> 32), 0u, 1u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u53(c[0], c[1], false);
}

// ---- RayIterator RayIterator.java:65-130 ----
struct RayIter {
    int x, y, x_inc, y_inc, n;
    float dx, dy, error;

    __device__ __forceinline__ void init(float x0, float y0, float x1, float y1, int additional) {
        dx = fabsf(x1 - x0);
        dy = fabsf(y1 - y0);
        const double fx0 = floor((double)x0), fy0 = floor((double)y0);
        x = java_d2i(fx0);
        y = java_d2i(fy0);
        n = 1 + additional;
        if (dx == 0.0f) {
            x_inc = 0;
            error = __int_as_float(0x7f800000);
        } else if (x1 > x0) {
            x_inc = 1;
            n += java_d2i(floor((double)x1) - (double)x);
            error = (float)((fx0 + 1.0 - (double)x0) * (double)dy);
        } else {
            x_inc = -1;
            n += x - java_d2i(floor((double)x1));
            error = (float)(((double)x0 - fx0) * (double)dy);
        }
        if (dy == 0.0f) {
            y_inc = 0;
            error = error - __int_as_float(0x7f800000);
        } else if (y1 > y0) {
            y_inc = 1;
            n += java_d2i(floor((double)y1)) - y;
            error = (float)((double)error - (fy0 + 1.0 - (double)y0) * (double)dx);
        } else {
            y_inc = -1;
            n += y - java_d2i(floor((double)y1));
            error = (float)((double)error - ((double)y0 - fy0) * (double)dx);
        }
    }
    __device__ __forceinline__ bool has_next(int W, int H) const {
        return n > 0 && !(x < 0 || x >= W || y < 0 || y >= H);
    }
    __device__ __forceinline__ void advance() {
        if (error > 0.0f) {
            y += y_inc;
            error -= dx;
        } else {
            x += x_inc;
            error += dy;
        }
        n -= 1;
    }
};

// SensorModel.inverseSensorModel SensorModel.java:31-41 -> 0 prior (+= 0), 1 free, 2 occupied.
__device__ __forceinline__ int inverse_sensor_class(float current, float measured, bool was_hit,
                                                    float tol_half) {
    if (!was_hit) return current < measured ? 1 : 0;
    if (current < measured - tol_half) return 1;
    if (current > measured + tol_half) return 0;
    return 2;
}

// A map cell: the two integer counters that represent Java's f64 log-odds exactly
// (every increment is one of L_free, L_occ, 0.0 — GridMap.java:223, SensorModel.java:23-25).
struct __align__(8) CellCounts {
    uint32_t n_free, n_occ;
};

constexpr int kTileW = 64, kTileH = 32;  // likelihood tile (cells)

// Thresholded code of a cell, GridMap.java:238-245: log-odds > 0 -> 1, < 0 -> 0, == 0 -> 0.5, from the
// closed form of the counters.  The same expression decides which tiles are rebuilt and what they hold.
__device__ __forceinline__ double cell_value(uint32_t n_free, uint32_t n_occ, const Geometry& g) {
    return (double)n_free * g.l_free + (double)n_occ * g.l_occ;
}
__device__ __forceinline__ int cell_code(uint32_t n_free, uint32_t n_occ, const Geometry& g) {
    const double v = cell_value(n_free, n_occ, g);
    return v > 0.0 ? 2 : (v < 0.0 ? 0 : 1);
}
// Does adding one free (cls 1) / occupied (cls 2) increment to (n_free, n_occ) change the code?
// L_free < 0 < L_occ: a never-occupied cell only flips on its first free hit, a never-freed cell only on
// its first occupied hit; everything else needs the two evaluations.
__device__ __forceinline__ bool code_flips(uint32_t n_free, uint32_t n_occ, int cls, const Geometry& g) {
    if (cls == 1) {
        if (n_occ == 0) return n_free == 0;
        return cell_code(n_free, n_occ, g) != cell_code(n_free + 1, n_occ, g);
    }
    if (n_free == 0) return n_occ == 0;
    return cell_code(n_free, n_occ, g) != cell_code(n_free, n_occ + 1, g);
}
// The likelihood value of a cell depends on the codes within khalf cells: mark every tile whose output
// can change when the code of (cx, cy) changes.
__device__ __forceinline__ void mark_dirty(uint32_t* __restrict__ bitmap, int cx, int cy, const Geometry& g) {
    const int tx0 = max(cx - g.khalf, 0) / kTileW, tx1 = min(cx + g.khalf, g.W - 1) / kTileW;
    const int ty0 = max(cy - g.khalf, 0) / kTileH, ty1 = min(cy + g.khalf, g.H - 1) / kTileH;
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
            const int t = ty * g.tiles_x + tx;
            atomicOr(bitmap + (t >> 5), 1u << (t & 31));
        }
}
// One counter increment.  The 64-bit atomic returns the cell as it was, so exactly one thread observes
// each transition: integer adds commute (deterministic counts), and a tile is queued for a likelihood
// rebuild only when the thresholded code of one of its cells really changed.
__device__ __forceinline__ void bump_cell(CellCounts* __restrict__ map, uint32_t* __restrict__ bitmap, int cx, int cy,
                                          int cls, const Geometry& g) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(map + ((size_t)cx + (size_t)cy * g.W));
    const unsigned long long old = atomicAdd(p, cls == 1 ? 1ull : (1ull << 32));
    if (code_flips((uint32_t)old, (uint32_t)(old >> 32), cls, g)) mark_dirty(bitmap, cx, cy, g);
}

// GridMap.applyMeasurement GridMap.java:194-228 on integer counters.
struct CellBox {
    int x0, y0, x1, y1;
};
__device__ __forceinline__ void apply_measurement(CellCounts* __restrict__ map, uint32_t* __restrict__ bitmap,
                                                  const Geometry& g, float sx, float sy, float ex, float ey,
                                                  float meas, bool was_hit, CellBox& box) {
    RayIter it;
    it.init(sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, g.extra_steps);
    box.x0 = box.y0 = 0x7fffffff;
    box.x1 = box.y1 = -1;
    if (!it.has_next(g.W, g.H)) return;
    box.x0 = box.x1 = it.x;
    box.y0 = box.y1 = it.y;
    int lx = it.x, ly = it.y;
    // The returned counters are only needed for the (rare) dirty-tile marking.  Four increments are issued
    // back to back before the first result is inspected, so four atomics are in flight per thread instead
    // of one (a branch on the returned value would otherwise stall the warp for a full L2 round trip).
    while (it.has_next(g.W, g.H)) {
        int cx[4], cy[4], cl[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // four DDA steps (statically indexed registers, predicated past the end)
            const bool live = it.has_next(g.W, g.H);
            cx[j] = it.x;
            cy[j] = it.y;
            cl[j] = 0;
            if (live) {
                lx = it.x;
                ly = it.y;
                const float dX = sx - ((float)lx + 0.5f);
                const float dY = sy - ((float)ly + 0.5f);
                const float dist = __fsqrt_rn(dX * dX + dY * dY);  // (float) Math.sqrt((double) f32) == sqrt.rn.f32
                cl[j] = inverse_sensor_class(dist, meas, was_hit, g.tol_half);
                it.advance();
            }
        }
        unsigned long long old[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cl[j] != 0)
                old[j] = atomicAdd(reinterpret_cast<unsigned long long*>(map + ((size_t)cx[j] + (size_t)cy[j] * g.W)),
                                   cl[j] == 1 ? 1ull : (1ull << 32));
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cl[j] != 0 && code_flips((uint32_t)old[j], (uint32_t)(old[j] >> 32), cl[j], g))
                mark_dirty(bitmap, cx[j], cy[j], g);
    }
    // the walk is monotone in x and in y: its bounding box is spanned by the first and last cell
    box.x0 = min(box.x0, lx); box.x1 = max(box.x1, lx);
    box.y0 = min(box.y0, ly); box.y1 = max(box.y1, ly);
}

// The same walk with fire-and-forget reductions (RED.E.ADD.64: no value returns, nothing waits for L2): used
// where no dirty-tile bookkeeping hangs on the transition, i.e. per-particle maps, whose likelihood field is
// evaluated on demand from the counters (k_score_pp).  NEG subtracts the increments again: the getter of a
// particle's likelihoodData needs the counters as they were BEFORE the last scan was integrated.
template <bool NEG>
__device__ __forceinline__ void apply_measurement_red(CellCounts* __restrict__ map, const Geometry& g, float sx, float sy,
                                                      float ex, float ey, float meas, bool was_hit, CellBox& box) {
    RayIter it;
    it.init(sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, g.extra_steps);
    box.x0 = box.y0 = 0x7fffffff;
    box.x1 = box.y1 = -1;
    if (!it.has_next(g.W, g.H)) return;
    box.x0 = box.x1 = it.x;
    box.y0 = box.y1 = it.y;
    int lx = it.x, ly = it.y;
    const unsigned long long inc_free = NEG ? ~0ull : 1ull;                      // pair - 1 (no borrow: n_free >= 1)
    const unsigned long long inc_occ = NEG ? (0ull - (1ull << 32)) : (1ull << 32);
    // Four DDA steps, then their four classifications and reductions: the DDA's error term is the only serial
    // dependence (RayIterator.java:112-130), so the square roots and the REDs of a group overlap with the next
    // group's walk.  Matters when few rays are in flight (only the selected parents' maps take a scan: the kernel
    // is then bound by this loop's latency, 74 us for ~10 maps before the unrolling).
    while (it.has_next(g.W, g.H)) {
        int cx[4], cy[4];
        bool live[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            live[j] = it.has_next(g.W, g.H);
            cx[j] = it.x;
            cy[j] = it.y;
            if (live[j]) {
                lx = it.x;
                ly = it.y;
                it.advance();
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (!live[j]) continue;
            const float dX = sx - ((float)cx[j] + 0.5f);
            const float dY = sy - ((float)cy[j] + 0.5f);
            const float dist = __fsqrt_rn(dX * dX + dY * dY);  // (float) Math.sqrt((double) f32) == sqrt.rn.f32
            const int cls = inverse_sensor_class(dist, meas, was_hit, g.tol_half);
            if (cls != 0)
                atomicAdd(reinterpret_cast<unsigned long long*>(map + ((size_t)cx[j] + (size_t)cy[j] * g.W)),
                          cls == 1 ? inc_free : inc_occ);
        }
    }
    box.x0 = min(box.x0, lx); box.x1 = max(box.x1, lx);
    box.y0 = min(box.y0, ly); box.y1 = max(box.y1, ly);
}

// ---- warp helpers ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gms
