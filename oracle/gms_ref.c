/*
 * gms_ref.c — CPU ORACLE for the grid-map SLAM hot path.  TEST INFRASTRUCTURE, NOT PRODUCT:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  libgms.so (the CUDA product) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference (antbern/gridmap-slam-robot) ships no tests, no golden vectors and
 * cannot be compiled here (Java 8 + Kotlin glm + commons-math3; no JDK in this image).  This file is
 * a literal restatement of the Java source text, pinned only by (1) the hand-derived known-answer
 * vectors of SURVEY.md Appendix B and (2) an independent pure-Python restatement
 * (oracle/pyref.py -> tests/golden/).  See DESIGN.md "Oracle".
 *
 * Every function cites the reference lines it follows; paths are relative to
 * java/GridMapGL/src/main/java/com/fmsz/gridmapgl/.  Java numeric semantics that matter:
 * strict left-to-right evaluation, no fused multiply-add (build with -ffp-contract=off),
 * float op double -> double, compound assignment narrows, (int) truncates toward zero and saturates.
 *
 * Third-party arithmetic not in the reference tree (commons-math3 3.6.1, build.gradle:59):
 * FastMath.sin/cos -> sin_fixed / cos_fixed below (a published < 1 ulp scheme as a fixed operation sequence);
 * NormalDistribution.sample() = sd * nextGaussian() + mean -> the standard normal draws are INPUTS
 * (injected) or come from the Philox generator below; BOBYQAOptimizer with an objective that is
 * identically zero (Odometry.java:99-103) -> identity on the start pose (GridMap.java:348-369).
 */
#define _GNU_SOURCE
#include "../include/gms.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

struct gms_handle {
    gms_config cfg;
    int W, H, P, lo, cnt, S;
    int ktaps;
    double kernel[32];
    double l_free, l_occ;
    float world_w, world_h;
    int resample_mode;
    int threads;
    /* global (size P) particle state */
    float *px, *py, *pt;
    double *w;     /* canonical normalised weights (log domain -> exp)              */
    double *lw;    /* ln of the un-normalised product of the last update            */
    double *wlit;  /* Java's literal weights: product / weightSum (NaN when it NaNs) */
    int32_t *parents;
    int32_t *slot; /* particle (local index) -> map slot                            */
    /* maps */
    double **logd, **lik;
    uint32_t **nfree, **nocc;
    double *prob_scratch, *tmp_scratch; /* per-thread scratch, threads * W*H */
    /* scalars */
    double neff, neff_lit;
    int32_t strongest, strongest_lit;
    int32_t strongest_now;        /* where the strongest particle of the last update lives now (first child) */
    float strongest_pose[3];      /* its pose / weight at that update (Java keeps the Particle object) */
    double strongest_w;
    double *comb_prod;            /* combined-map fusion: product over the local particles */
    gms_pose_optimizer_fn opt_fn; /* A4 hook: GridMap.findBestPoseOptim (default NULL = identity) */
    void *opt_user;
    uint64_t step, resample_count;
    int have_update;
    /* exchange blocks for the begin/end split ("device" == host for the oracle) */
    unsigned char *xlocal, *xglobal;
    /* pending step inputs between begin and end */
    int pend_B;
    double *pend_xy, *pend_dist;
    uint8_t *pend_hit;
    double pend_dtheta;
    /* wall-clock seconds per phase of the step, accumulated when threads == 1 (the per-phase
     * split of the CPU path): motion, likelihood field, scoring, map integration, normalise, resample */
    double phase[6];
    char err[256];
};

static inline double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#define PHASE_T0(h) double t0__ = (h)->threads == 1 ? now_s() : 0.0
#define PHASE_ADD(h, k)                                  \
    do {                                                 \
        if ((h)->threads == 1) {                         \
            double t1__ = now_s();                       \
            (h)->phase[k] += t1__ - t0__;                \
            t0__ = t1__;                                 \
        }                                                \
    } while (0)

static __thread char g_create_err[256];

static int fail(gms_handle *h, int code, const char *msg) {
    if (h) snprintf(h->err, sizeof h->err, "%s", msg);
    else snprintf(g_create_err, sizeof g_create_err, "%s", msg);
    return code;
}

/* ---------------------------------------------------------------------------------------------
 * Java primitive conversions
 * ------------------------------------------------------------------------------------------- */
/* JLS 5.1.3: (int) of a double: NaN -> 0, saturate, else truncate toward zero. */
static inline int32_t java_d2i(double d) {
    if (d != d) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return (-2147483647 - 1);
    return (int32_t)d;
}

/* MathUtil.angleConstrain MathUtil.java:65-72 (literal loops: NOT the identity on in-range input). */
static double angle_constrain(double a) {
    while (a < M_PI) a += M_PI * 2;
    while (a > M_PI) a -= M_PI * 2;
    return a;
}
/* MathUtil.sin / cos (MathUtil.java:30-48) delegate to FastMath.sin / cos of commons-math3, which is not in the
 * reference tree.  FastMath, glibc and CUDA are each accurate to about an ulp and disagree in the last bit now
 * and then, so the oracle does not call libm here: it evaluates the classic published scheme — Sun's fdlibm 5.3,
 * e_rem_pio2.c (medium arguments: Cody-Waite reduction by pi/2 held as 33 + 33 + 53 bits, up to three rounds),
 * k_sin.c and k_cos.c (minimax polynomials on [-pi/4, pi/4]), error < 1 ulp — written out as a fixed sequence of
 * IEEE operations (no contraction: -ffp-contract=off), which any implementation can reproduce bit for bit; the
 * CUDA library does (device_math.cuh), oracle/pyref.py restates it a third time.  |x| >= 2^20 * pi/2 (never an
 * angle on this path) and non-finite arguments go to libm.  tests/test_trig.py holds it within 1 ulp of an
 * exact (mpmath) sine / cosine. */
static const double TS1 = -0x1.5555555555549p-3, TS2 = 0x1.111111110f8a6p-7, TS3 = -0x1.a01a019c161d5p-13,
                    TS4 = 0x1.71de357b1fe7dp-19, TS5 = -0x1.ae5e68a2b9cebp-26, TS6 = 0x1.5d93a5acfd57cp-33;
static const double TC1 = 0x1.555555555554cp-5, TC2 = -0x1.6c16c16c15177p-10, TC3 = 0x1.a01a019cb1590p-16,
                    TC4 = -0x1.27e4f809c52adp-22, TC5 = 0x1.1ee9ebdb4b1c4p-29, TC6 = -0x1.8fae9be8838d4p-37;
static double poly_sin(double x, double tail) { /* k_sin.c, the form that takes the tail of the argument */
    double z = x * x;
    double v = z * x;
    double r = TS2 + z * (TS3 + z * (TS4 + z * (TS5 + z * TS6)));
    return x - ((z * (0.5 * tail - v * r) - tail) - v * TS1);
}
static double poly_cos(double x, double tail) { /* k_cos.c */
    double z = x * x;
    double zz = z * z;
    double r = z * (TC1 + z * (TC2 + z * TC3)) + (zz * zz) * (TC4 + z * (TC5 + z * TC6));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * tail));
}
static int exponent_field(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int)((u >> 52) & 0x7ff);
}
/* e_rem_pio2.c, medium case: quadrant count, reduced argument as head + tail */
static int reduce_pio2(double x, double *head, double *tail) {
    static const double inv = 0x1.45f306dc9c883p-1;
    static const double p1 = 0x1.921fb54400000p+0, p1t = 0x1.0b4611a626331p-34;
    static const double p2 = 0x1.0b4611a600000p-34, p2t = 0x1.3198a2e037073p-69;
    static const double p3 = 0x1.3198a2e000000p-69, p3t = 0x1.b839a252049c1p-104;
    if (fabs(x) <= 0x1.921fb54442d18p-1) { *head = x; *tail = 0.0; return 0; }
    double fn = rint(x * inv); /* default rounding mode: to nearest, ties to even */
    double r = x - fn * p1;
    double w = fn * p1t;
    double y = r - w;
    int ex = exponent_field(x);
    if (ex - exponent_field(y) > 16) {
        double t = r;
        w = fn * p2;
        r = t - w;
        w = fn * p2t - ((t - r) - w);
        y = r - w;
        if (ex - exponent_field(y) > 49) {
            t = r;
            w = fn * p3;
            r = t - w;
            w = fn * p3t - ((t - r) - w);
            y = r - w;
        }
    }
    *head = y;
    *tail = (r - y) - w;
    return (int)fn;
}
static double sin_fixed(double x) {
    if (!(fabs(x) < 1647099.0)) return sin(x);
    double a, b;
    int q = reduce_pio2(x, &a, &b) & 3;
    switch (q) {
    case 0: return poly_sin(a, b);
    case 1: return poly_cos(a, b);
    case 2: return -poly_sin(a, b);
    default: return -poly_cos(a, b);
    }
}
static double cos_fixed(double x) {
    if (!(fabs(x) < 1647099.0)) return cos(x);
    double a, b;
    int q = reduce_pio2(x, &a, &b) & 3;
    switch (q) {
    case 0: return poly_cos(a, b);
    case 1: return -poly_sin(a, b);
    case 2: return -poly_cos(a, b);
    default: return poly_sin(a, b);
    }
}
/* MathUtil.cos(float)/sin(float) MathUtil.java:30-40: (float) FastMath.cos((double) radians). */
static inline float cos_f(float r) { return (float)cos_fixed((double)r); }
static inline float sin_f(float r) { return (float)sin_fixed((double)r); }

/* Util.logOdds(double) Util.java:35-37: Math.log(odds / (1.0f - odds)). */
static double log_odds(double p) { return log(p / (1.0 - p)); }

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011) — the generator shared with the CUDA library for the
 * "draw on the device" path.  Not part of the reference (its RNGs are unseeded: Odometry.java:27,
 * SLAM.java:136); parity of this path is oracle-vs-CUDA only.
 * ------------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
/* 53-bit uniform from two words; offset=1 gives (k+0.5)/2^53 in (0,1), offset=0 gives k/2^53 in [0,1). */
static double u53(uint32_t a, uint32_t b, int centred) {
    uint64_t k = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return centred ? ((double)k + 0.5) * 0x1p-53 : (double)k * 0x1p-53;
}
/* stream 0: the two standard normals {z_d, z_theta} of global particle `gidx` at step `step`. */
static void philox_normals(uint64_t seed, uint32_t gidx, uint64_t step, double *zd, double *zt) {
    uint32_t c[4] = {gidx, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u53(c[0], c[1], 1), u2 = u53(c[2], c[3], 0);
    double r = sqrt(-2.0 * log(u1));
    double a = 6.283185307179586 * u2;
    *zd = r * cos_fixed(a);
    *zt = r * sin_fixed(a);
}
/* stream 1: the resampling uniform of resample number n. */
static double philox_uniform(uint64_t seed, uint64_t n) {
    uint32_t c[4] = {(uint32_t)n, (uint32_t)(n >> 32), 0u, 1u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u53(c[0], c[1], 0);
}

/* ---------------------------------------------------------------------------------------------
 * Util.generateGaussianKernel Util.java:428-455
 * ------------------------------------------------------------------------------------------- */
static void generate_gaussian_kernel(double sigma, int size, double *values) {
    int ksize = size * 2 + 1;
    double norm = 1.0 / (sqrt(2 * M_PI) * sigma);
    double coeff = 2 * sigma * sigma;
    double total = 0;
    for (int x = -size; x <= size; x++) {
        /* Java: -x * x / coeff with int x: ((-x) * x) as int, then / double */
        double g = norm * exp((double)(-x * x) / coeff);
        values[x + size] = g;
        total += g;
    }
    for (int i = 0; i < ksize; i++) values[i] /= total;
}

/* ---------------------------------------------------------------------------------------------
 * Util.doGaussianBlurdSeparable Util.java:378-426 (out-of-range taps are skipped, no renormalisation)
 * ------------------------------------------------------------------------------------------- */
static void blur_separable(const double *in, double *out, double *tmp, int width, int height,
                           const double *kernel, int ktaps, int threads) {
    int k = (ktaps - 1) / 2;
    /* rows are independent: the OpenMP split (reference arm only, threads > 1) is bit-identical */
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1 && !omp_in_parallel())
#endif
    for (int y = 0; y < height; y++) {
        int yi = y * width;
        for (int x = 0; x < width; x++) {
            double total = 0;
            for (int i = -k; i <= k; i++) {
                int x2 = x + i;
                if (x2 >= 0 && x2 < width) total += kernel[i + k] * in[yi + x2];
            }
            out[yi + x] = total;
        }
    }
    memcpy(tmp, out, sizeof(double) * (size_t)width * height);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1 && !omp_in_parallel())
#endif
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            double total = 0;
            for (int i = -k; i <= k; i++) {
                int y2 = y + i;
                if (y2 >= 0 && y2 < height) total += kernel[i + k] * tmp[x + y2 * width];
            }
            out[x + y * width] = total;
        }
    }
}

/* GridMap.computeLikelihoodMap GridMap.java:233-250: threshold against logOdds(0.5) == 0.0, then blur. */
static void compute_likelihood(const gms_handle *h, const double *logd, double *lik, double *prob,
                               double *tmp) {
    size_t n = (size_t)h->W * h->H;
    for (size_t i = 0; i < n; i++) {
        if (logd[i] > 0.0) prob[i] = 1;
        else if (logd[i] < 0.0) prob[i] = 0;
        else prob[i] = 0.5;
    }
    blur_separable(prob, lik, tmp, h->W, h->H, h->kernel, h->ktaps, h->threads);
}

/* ---------------------------------------------------------------------------------------------
 * RayIterator RayIterator.java:65-130
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int x, y, width, height, x_inc, y_inc, n;
    float dx, dy, error;
} ray_iter;

static void ray_init(ray_iter *it, int width, int height, float x0, float y0, float x1, float y1,
                     int additional) {
    it->width = width;
    it->height = height;
    it->dx = fabsf(x1 - x0);
    it->dy = fabsf(y1 - y0);
    it->x = java_d2i(floor((double)x0));
    it->y = java_d2i(floor((double)y0));
    it->n = 1 + additional;
    if (it->dx == 0) {
        it->x_inc = 0;
        it->error = INFINITY;
    } else if (x1 > x0) {
        it->x_inc = 1;
        it->n += java_d2i(floor((double)x1) - it->x);
        it->error = (float)((floor((double)x0) + 1 - (double)x0) * (double)it->dy);
    } else {
        it->x_inc = -1;
        it->n += it->x - java_d2i(floor((double)x1));
        it->error = (float)(((double)x0 - floor((double)x0)) * (double)it->dy);
    }
    if (it->dy == 0) {
        it->y_inc = 0;
        it->error -= INFINITY; /* float - float */
    } else if (y1 > y0) {
        it->y_inc = 1;
        it->n += java_d2i(floor((double)y1)) - it->y;
        it->error = (float)((double)it->error - (floor((double)y0) + 1 - (double)y0) * (double)it->dx);
    } else {
        it->y_inc = -1;
        it->n += it->y - java_d2i(floor((double)y1));
        it->error = (float)((double)it->error - ((double)y0 - floor((double)y0)) * (double)it->dx);
    }
}
static inline int ray_has_next(const ray_iter *it) {
    return it->n > 0 && !(it->x < 0 || it->x >= it->width || it->y < 0 || it->y >= it->height);
}
static inline void ray_next(ray_iter *it, int *cx, int *cy) {
    *cx = it->x;
    *cy = it->y;
    if (it->error > 0) {
        it->y += it->y_inc;
        it->error -= it->dx;
    } else {
        it->x += it->x_inc;
        it->error += it->dy;
    }
    it->n -= 1;
}

/* SensorModel.inverseSensorModel SensorModel.java:31-41 -> class: 0 = prior, 1 = free, 2 = occupied. */
static inline int inverse_sensor_class(float current, float measured, int was_hit, float tol) {
    if (!was_hit) return current < measured ? 1 : 0;
    if (current < measured - tol / 2) return 1;
    if (current > measured + tol / 2) return 0;
    return 2;
}

/* GridMap.applyMeasurement GridMap.java:194-228 */
static void apply_measurement(const gms_handle *h, double *logd, uint32_t *nfree, uint32_t *nocc,
                              float sx, float sy, float ex, float ey, float meas, int was_hit) {
    ray_iter it;
    ray_init(&it, h->W, h->H, sx + 0.5f, sy + 0.5f, ex + 0.5f, ey + 0.5f, h->cfg.extra_steps);
    while (ray_has_next(&it)) {
        int cx, cy;
        ray_next(&it, &cx, &cy);
        float dX = sx - ((float)cx + 0.5f);
        float dY = sy - ((float)cy + 0.5f);
        float distance = (float)sqrt((double)(dX * dX + dY * dY));
        int cls = inverse_sensor_class(distance, meas, was_hit, h->cfg.hit_tolerance);
        size_t idx = (size_t)cx + (size_t)cy * h->W;
        /* logData[idx] += Util.logOdds(p): three possible increments (logOdds(0.5) == 0.0) */
        if (cls == 1) {
            logd[idx] += h->l_free;
            nfree[idx] += 1;
        } else if (cls == 2) {
            logd[idx] += h->l_occ;
            nocc[idx] += 1;
        } else {
            logd[idx] += 0.0;
        }
    }
}

/* Transform.fromRobotToWorld Transform.java:13-32 */
typedef struct { double cos, sin, px, py; } xform;
static inline xform robot_to_world(float x, float y, float theta) {
    xform t = {(double)cos_f(theta), (double)sin_f(theta), (double)x, (double)y};
    return t;
}
static inline double tx(const xform *t, double x, double y) { return x * t->cos - y * t->sin + t->px; }
static inline double ty(const xform *t, double x, double y) { return x * t->sin + y * t->cos + t->py; }

/* GridMap.integrateObservation GridMap.java:173-191 */
static void integrate_observation(const gms_handle *h, double *logd, uint32_t *nfree, uint32_t *nocc,
                                  float px, float py, float pt, const double *bxy,
                                  const double *bdist, const uint8_t *bhit, int B) {
    xform t = robot_to_world(px, py, pt);
    double posx = (double)h->cfg.origin_x, posy = (double)h->cfg.origin_y;
    double res = (double)h->cfg.resolution;
    float sx = (float)((tx(&t, 0, 0) - posx) / res);
    float sy = (float)((ty(&t, 0, 0) - posy) / res);
    for (int b = 0; b < B; b++) {
        float ex = (float)((tx(&t, bxy[2 * b], bxy[2 * b + 1]) - posx) / res);
        float ey = (float)((ty(&t, bxy[2 * b], bxy[2 * b + 1]) - posy) / res);
        float meas = (float)bdist[b] / h->cfg.resolution;
        apply_measurement(h, logd, nfree, nocc, sx, sy, ex, ey, meas, bhit[b] != 0);
    }
}

/* GridMap.probabilityOf GridMap.java:261-294.  Returns Java's product; *log_out = sum of ln(factor)
 * (the log-domain value the CUDA path computes; finite where the product underflows). */
static double probability_of(const gms_handle *h, const double *lik, float px, float py, float pt,
                             const double *bxy, const uint8_t *bhit, int B, double *log_out) {
    double product = 1, lsum = 0;
    xform t = robot_to_world(px, py, pt);
    double posx = (double)h->cfg.origin_x, posy = (double)h->cfg.origin_y;
    double res = (double)h->cfg.resolution;
    double zhit = h->cfg.z_hit, zrandom = 1 - zhit;
    double range = (double)h->cfg.sensor_max_range;
    for (int b = 0; b < B; b++) {
        if (!bhit[b]) continue;
        int gx = java_d2i((tx(&t, bxy[2 * b], bxy[2 * b + 1]) - posx) / res);
        int gy = java_d2i((ty(&t, bxy[2 * b], bxy[2 * b + 1]) - posy) / res);
        if (!(gx < 0 || gy < 0 || gx >= h->W || gy >= h->H)) {
            double val = lik[(size_t)gx + (size_t)gy * h->W];
            double f;
            if (val == 0.5) f = 1.0 / range;
            else f = zhit * val + zrandom * 1.0 / range;
            product *= f;
            lsum += log(f);
        }
    }
    if (log_out) *log_out = lsum;
    return product;
}

/* SLAM.sampleMotionModel SLAM.java:155-163 + Odometry.apply Odometry.java:77-96 with the two
 * NormalDistribution.sample() results written as sd * z + mean (commons-math3 3.6.1). */
static void motion_sample(const gms_handle *h, float *x, float *y, float *theta, double d_center,
                          double d_theta, double zd, double zt) {
    const gms_config *c = &h->cfg;
    /* Odometry.recalculateStdDev Odometry.java:60-69 */
    double sd_c = (c->noise_center_base + fabs(d_center) * c->noise_center_gain) / 2;
    double sd_t = c->noise_theta_base_deg * (M_PI / 180.0) + c->noise_theta_gain * fabs(d_theta);
    double d = sd_c * zd + d_center;
    double th = sd_t * zt + d_theta;
    *theta = (float)angle_constrain((double)*theta + th);
    *x = (float)((double)*x + (double)cos_f(*theta) * d);
    *y = (float)((double)*y + (double)sin_f(*theta) * d);
}

/* ---------------------------------------------------------------------------------------------
 * lifecycle
 * ------------------------------------------------------------------------------------------- */
EXPORT int gms_config_default(gms_config *cfg) {
    if (!cfg) return GMS_ERR_INVALID_ARG;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = (uint32_t)sizeof *cfg;
    cfg->num_particles = 500;       /* SLAM.java:50 */
    cfg->map_width_m = 6.0f;        /* SLAM.java:57 */
    cfg->map_height_m = 6.0f;
    cfg->resolution = 0.05f;
    cfg->origin_x = -3.0f;
    cfg->origin_y = -3.0f;
    cfg->sensor_max_range = 10.0f;  /* SensorModel.java:20 */
    cfg->z_hit = 0.9;               /* GridMap.java:259 */
    cfg->hit_tolerance = 2.0f;      /* GridMap.java:223 */
    cfg->extra_steps = 2;           /* GridMap.java:210 */
    cfg->p_free = 0.30f;            /* SensorModel.java:23-24 */
    cfg->p_occ = 0.9f;
    cfg->noise_center_base = 0.01;  /* Odometry.java:63-64 */
    cfg->noise_center_gain = 0.05;
    cfg->noise_theta_base_deg = 5;
    cfg->noise_theta_gain = 0.1;
    cfg->skip_update_deg = 30;      /* SLAM.java:82 */
    cfg->likelihood_sigma_num = 0.05; /* GridMap.java:94 */
    cfg->map_mode = GMS_MAP_PER_PARTICLE;
    cfg->resample_mode = GMS_RESAMPLE_AUTO;
    cfg->device = 0;
    cfg->rank = 0;
    cfg->nranks = 1;
    cfg->seed = 0x5EEDull;
    return GMS_OK;
}

static void free_handle(gms_handle *h) {
    if (!h) return;
    free(h->px); free(h->py); free(h->pt); free(h->w); free(h->lw); free(h->wlit);
    free(h->parents); free(h->slot);
    for (int s = 0; s < h->S; s++) {
        if (h->logd) free(h->logd[s]);
        if (h->lik) free(h->lik[s]);
        if (h->nfree) free(h->nfree[s]);
        if (h->nocc) free(h->nocc[s]);
    }
    free(h->logd); free(h->lik); free(h->nfree); free(h->nocc);
    free(h->prob_scratch); free(h->tmp_scratch); free(h->comb_prod);
    free(h->xlocal); free(h->xglobal);
    free(h->pend_xy); free(h->pend_dist); free(h->pend_hit);
    free(h);
}

EXPORT int gms_reset(gms_handle *h);

EXPORT int gms_create(const gms_config *cfg, gms_handle **out) {
    if (!cfg || !out) return fail(NULL, GMS_ERR_INVALID_ARG, "gms_create: NULL argument");
    if (cfg->struct_size != sizeof(gms_config))
        return fail(NULL, GMS_ERR_INVALID_ARG, "gms_create: gms_config.struct_size mismatch");
    if (cfg->num_particles < 1 || !(cfg->resolution > 0) || !(cfg->map_width_m > 0) ||
        !(cfg->map_height_m > 0) || cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks ||
        cfg->extra_steps < 0 || (cfg->map_mode != GMS_MAP_PER_PARTICLE && cfg->map_mode != GMS_MAP_SHARED) ||
        cfg->resample_mode < 0 || cfg->resample_mode > 2 || cfg->num_particles % cfg->nranks != 0)
        return fail(NULL, GMS_ERR_INVALID_ARG, "gms_create: invalid configuration");
    gms_handle *h = (gms_handle *)calloc(1, sizeof *h);
    if (!h) return fail(NULL, GMS_ERR_OOM, "gms_create: out of memory");
    h->cfg = *cfg;
    /* GridMap ctor GridMap.java:80-100: float division, Math.ceil, float product */
    h->W = java_d2i(ceil((double)(cfg->map_width_m / cfg->resolution)));
    h->H = java_d2i(ceil((double)(cfg->map_height_m / cfg->resolution)));
    h->world_w = (float)h->W * cfg->resolution;
    h->world_h = (float)h->H * cfg->resolution;
    double sigma = sqrt(cfg->likelihood_sigma_num / (double)cfg->resolution);
    int half = java_d2i(ceil(sigma * 3));
    h->ktaps = 2 * half + 1;
    if (h->ktaps > 31 || h->W < 1 || h->H < 1) {
        free(h);
        return fail(NULL, GMS_ERR_INVALID_ARG, "gms_create: kernel too wide or empty grid");
    }
    generate_gaussian_kernel(sigma, half, h->kernel);
    h->l_free = log_odds((double)cfg->p_free);
    h->l_occ = log_odds((double)cfg->p_occ);
    h->P = cfg->num_particles;
    h->cnt = h->P / cfg->nranks;
    h->lo = cfg->rank * h->cnt;
    h->S = cfg->map_mode == GMS_MAP_SHARED ? 1 : h->cnt;
    h->resample_mode = cfg->resample_mode == GMS_RESAMPLE_AUTO
                           ? (h->P <= 2048 ? GMS_RESAMPLE_LITERAL : GMS_RESAMPLE_FIXED)
                           : cfg->resample_mode;
    h->threads = 1;
    size_t n = (size_t)h->W * h->H;
    h->px = calloc(h->P, sizeof(float)); h->py = calloc(h->P, sizeof(float)); h->pt = calloc(h->P, sizeof(float));
    h->w = calloc(h->P, sizeof(double)); h->lw = calloc(h->P, sizeof(double)); h->wlit = calloc(h->P, sizeof(double));
    h->parents = calloc(h->P, sizeof(int32_t)); h->slot = calloc(h->cnt, sizeof(int32_t));
    h->logd = calloc(h->S, sizeof(double *)); h->lik = calloc(h->S, sizeof(double *));
    h->nfree = calloc(h->S, sizeof(uint32_t *)); h->nocc = calloc(h->S, sizeof(uint32_t *));
    h->xlocal = calloc((size_t)h->cnt, 24); h->xglobal = calloc((size_t)h->P, 24);
    int ok = h->px && h->py && h->pt && h->w && h->lw && h->wlit && h->parents && h->slot && h->logd &&
             h->lik && h->nfree && h->nocc && h->xlocal && h->xglobal;
    for (int s = 0; ok && s < h->S; s++) {
        h->logd[s] = malloc(n * sizeof(double)); h->lik[s] = malloc(n * sizeof(double));
        h->nfree[s] = malloc(n * sizeof(uint32_t)); h->nocc[s] = malloc(n * sizeof(uint32_t));
        ok = h->logd[s] && h->lik[s] && h->nfree[s] && h->nocc[s];
    }
    if (!ok) {
        free_handle(h);
        return fail(NULL, GMS_ERR_OOM, "gms_create: out of memory");
    }
    gms_reset(h);
    *out = h;
    return GMS_OK;
}

EXPORT int gms_destroy(gms_handle *h) {
    free_handle(h);
    return GMS_OK;
}

EXPORT const char *gms_last_error(const gms_handle *h) { return h ? h->err : g_create_err; }

EXPORT int gms_get_info(const gms_handle *h, gms_info *info) {
    if (!h || !info) return GMS_ERR_INVALID_ARG;
    memset(info, 0, sizeof *info);
    info->abi_version = GMS_ABI_VERSION;
    info->is_cuda = 0;
    info->grid_w = h->W; info->grid_h = h->H;
    info->num_particles = h->P; info->local_begin = h->lo; info->local_count = h->cnt;
    info->num_slots = h->S; info->kernel_taps = h->ktaps; info->resample_mode = h->resample_mode;
    memcpy(info->kernel, h->kernel, sizeof(double) * h->ktaps);
    info->l_free = h->l_free; info->l_occ = h->l_occ;
    info->world_w = h->world_w; info->world_h = h->world_h;
    return GMS_OK;
}

/* SLAM.reset SLAM.java:65-77 + GridMap.createMapData(null) GridMap.java:106-117 */
EXPORT int gms_reset(gms_handle *h) {
    if (!h) return GMS_ERR_INVALID_ARG;
    size_t n = (size_t)h->W * h->H;
    for (int i = 0; i < h->P; i++) {
        h->px[i] = h->py[i] = h->pt[i] = 0;
        h->w[i] = 1.0 / h->P;
        h->wlit[i] = 1.0 / h->P;
        h->lw[i] = 0;
        h->parents[i] = i;
    }
    for (int i = 0; i < h->cnt; i++) h->slot[i] = h->cfg.map_mode == GMS_MAP_SHARED ? 0 : i;
    for (int s = 0; s < h->S; s++) {
        for (size_t i = 0; i < n; i++) h->logd[s][i] = log_odds(0.5);
        memset(h->lik[s], 0, n * sizeof(double));
        memset(h->nfree[s], 0, n * sizeof(uint32_t));
        memset(h->nocc[s], 0, n * sizeof(uint32_t));
    }
    h->strongest = 0; h->strongest_lit = 0;
    h->neff = h->neff_lit = (double)h->P;
    h->step = 0; h->resample_count = 0; h->have_update = 0;
    return GMS_OK;
}

static int ensure_scratch(gms_handle *h) {
    if (h->prob_scratch) return 1;
    /* per-thread scratch only where threads blur different maps (per-particle mode) */
    size_t n = (size_t)h->W * h->H * (size_t)(h->cfg.map_mode == GMS_MAP_SHARED ? 1 : h->threads);
    h->prob_scratch = malloc(n * sizeof(double));
    h->tmp_scratch = malloc(n * sizeof(double));
    return h->prob_scratch && h->tmp_scratch;
}

/* ---------------------------------------------------------------------------------------------
 * normalisation, Neff, strongest — canonical (log domain) and literal (Java) side by side
 * ------------------------------------------------------------------------------------------- */
static double neff_of(const double *w, int n) { /* SLAM.calculateNeff SLAM.java:180-190 */
    double sum = 0;
    for (int i = 0; i < n; i++) sum += w[i];
    double sq = 0;
    for (int i = 0; i < n; i++) sq += (w[i] / sum) * (w[i] / sum);
    return 1.0 / sq;
}

static void normalise(gms_handle *h) {
    int P = h->P;
    /* literal: SLAM.java:87-121 (weightSum accumulates in particle order; strict > keeps the first max) */
    double wsum = 0;
    int best = 0;
    for (int i = 0; i < P; i++) {
        wsum += h->wlit[i];
        if (i > 0 && h->wlit[i] > h->wlit[best]) best = i;
    }
    h->strongest_lit = best;
    for (int i = 0; i < P; i++) h->wlit[i] /= wsum;
    h->neff_lit = neff_of(h->wlit, P);
    /* canonical: w_i = exp(lw_i - max) / sum; first arg-max */
    int cb = 0;
    for (int i = 1; i < P; i++)
        if (h->lw[i] > h->lw[cb]) cb = i;
    double m = h->lw[cb], s = 0;
    for (int i = 0; i < P; i++) {
        h->w[i] = exp(h->lw[i] - m);
        s += h->w[i];
    }
    for (int i = 0; i < P; i++) h->w[i] /= s;
    h->strongest = cb;
    h->strongest_now = cb;
    h->strongest_pose[0] = h->px[cb]; h->strongest_pose[1] = h->py[cb]; h->strongest_pose[2] = h->pt[cb];
    h->strongest_w = h->w[cb];
    h->neff = neff_of(h->w, P);
}

/* ---------------------------------------------------------------------------------------------
 * resampling: index selection
 * ------------------------------------------------------------------------------------------- */
/* SLAM.resample SLAM.java:133-153, index part.  LITERAL: Java's sequential f64 running sum (with the
 * i <= n-1 clamp Java lacks: it would throw IndexOutOfBounds, SURVEY.md Appendix B2 last row).
 * FIXED: the same walk over u64 fixed point (trunc(w * 2^60), trunc(U * 2^60)). */
EXPORT int gmsref_resample_indices(const double *w, int32_t n, double u01, int32_t mode, int32_t *out) {
    if (!w || !out || n < 1) return GMS_ERR_INVALID_ARG;
    double r = u01 * 1.0 / n;
    if (mode == GMS_RESAMPLE_FIXED) {
        uint64_t c = (uint64_t)(w[0] * 0x1p60);
        int i = 0;
        for (int m = 1; m <= n; m++) {
            double U = r + (m - 1) * 1.0 / n;
            uint64_t Uq = (uint64_t)(U * 0x1p60);
            while (Uq > c && i < n - 1) {
                i++;
                c += (uint64_t)(w[i] * 0x1p60);
            }
            out[m - 1] = i;
        }
    } else {
        double c = w[0];
        int i = 0;
        for (int m = 1; m <= n; m++) {
            double U = r + (m - 1) * 1.0 / n;
            while (U > c && i < n - 1) {
                i++;
                c += w[i];
            }
            out[m - 1] = i;
        }
    }
    return GMS_OK;
}

/* ---------------------------------------------------------------------------------------------
 * the step
 * ------------------------------------------------------------------------------------------- */
struct xrec { double lw; float x, y, t; uint32_t pad; };

/* local phase: motion -> likelihood -> scoring (-> per-particle map integration) */
static int update_local(gms_handle *h, const double *bxy, const double *bdist, const uint8_t *bhit,
                        int B, double d_center, double d_theta, const double *normals) {
    if (!ensure_scratch(h)) return fail(h, GMS_ERR_OOM, "out of memory (scratch)");
    int skip = fabs(d_theta) > (M_PI / 180.0) * h->cfg.skip_update_deg; /* SLAM.java:82 */
    size_t n = (size_t)h->W * h->H;
    int shared = h->cfg.map_mode == GMS_MAP_SHARED;
    if (shared) {
        PHASE_T0(h);
        compute_likelihood(h, h->logd[0], h->lik[0], h->prob_scratch, h->tmp_scratch);
        PHASE_ADD(h, 1);
    }
    if (h->opt_fn) {
        /* A4 hook installed: the loop of SLAM.java:88 is run in two sweeps — motion sample + likelihood field of
         * every particle, then the hook (findBestPoseOptim, SLAM.java:97) on all poses, then the weights.  The
         * particles are independent, so the split changes nothing else. */
        float *xyt = malloc(sizeof(float) * 3 * (size_t)(h->cnt > 0 ? h->cnt : 1));
        for (int li = 0; li < h->cnt; li++) {
            int i = h->lo + li;
            double zd, zt;
            if (normals) { zd = normals[2 * li]; zt = normals[2 * li + 1]; }
            else philox_normals(h->cfg.seed, (uint32_t)i, h->step, &zd, &zt);
            motion_sample(h, &h->px[i], &h->py[i], &h->pt[i], d_center, d_theta, zd, zt);
            if (!shared) compute_likelihood(h, h->logd[h->slot[li]], h->lik[h->slot[li]], h->prob_scratch, h->tmp_scratch);
            xyt[3 * li] = h->px[i]; xyt[3 * li + 1] = h->py[i]; xyt[3 * li + 2] = h->pt[i];
        }
        int rc = h->opt_fn(h->opt_user, h, h->lo, h->cnt, xyt, bxy, bdist, bhit, B, d_center, d_theta);
        if (rc == 0)
            for (int li = 0; li < h->cnt; li++) {
                int i = h->lo + li;
                h->px[i] = xyt[3 * li]; h->py[i] = xyt[3 * li + 1]; h->pt[i] = xyt[3 * li + 2];
            }
        free(xyt);
        if (rc != 0) return fail(h, GMS_ERR_STATE, "pose optimiser hook failed");
    }
    const int hooked = h->opt_fn != NULL;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(h->threads)
#endif
    for (int li = 0; li < h->cnt; li++) {
        int i = h->lo + li;
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double zd, zt;
        PHASE_T0(h);
        int s = h->slot[li];
        if (!hooked) {
            if (normals) { zd = normals[2 * li]; zt = normals[2 * li + 1]; }
            else philox_normals(h->cfg.seed, (uint32_t)i, h->step, &zd, &zt);
            motion_sample(h, &h->px[i], &h->py[i], &h->pt[i], d_center, d_theta, zd, zt); /* SLAM.java:90 */
            PHASE_ADD(h, 0);
            if (!shared)
                compute_likelihood(h, h->logd[s], h->lik[s], h->prob_scratch + n * tid,
                                   h->tmp_scratch + n * tid);                               /* SLAM.java:93 */
            PHASE_ADD(h, 1);
        }
        /* SLAM.java:97 findBestPoseOptim: objective == 0 -> start pose (identity) unless a hook replaced it */
        double lw;
        h->wlit[i] = probability_of(h, h->lik[s], h->px[i], h->py[i], h->pt[i], bxy, bhit, B, &lw); /* :99 */
        h->lw[i] = lw;
        PHASE_ADD(h, 2);
        if (!shared && !skip)
            integrate_observation(h, h->logd[s], h->nfree[s], h->nocc[s], h->px[i], h->py[i], h->pt[i],
                                  bxy, bdist, bhit, B);                                 /* SLAM.java:102-107 */
        PHASE_ADD(h, 3);
    }
    struct xrec *xl = (struct xrec *)h->xlocal;
    for (int li = 0; li < h->cnt; li++) {
        int i = h->lo + li;
        xl[li].lw = h->lw[i]; xl[li].x = h->px[i]; xl[li].y = h->py[i]; xl[li].t = h->pt[i]; xl[li].pad = 0;
    }
    return GMS_OK;
}

static void do_resample(gms_handle *h, double u01);

/* global phase: import all particles' (lw, pose), normalise, strongest, shared-map integration */
static int update_global(gms_handle *h, const double *bxy, const double *bdist, const uint8_t *bhit,
                         int B, double d_theta, int policy, double u01) {
    const struct xrec *xg = (const struct xrec *)h->xglobal;
    for (int i = 0; i < h->P; i++) {
        if (i >= h->lo && i < h->lo + h->cnt) continue;
        h->lw[i] = xg[i].lw; h->px[i] = xg[i].x; h->py[i] = xg[i].y; h->pt[i] = xg[i].t;
        h->wlit[i] = exp(xg[i].lw); /* literal products of remote particles are not exchanged */
    }
    PHASE_T0(h);
    normalise(h);
    PHASE_ADD(h, 4);
    int skip = fabs(d_theta) > (M_PI / 180.0) * h->cfg.skip_update_deg;
    if (h->cfg.map_mode == GMS_MAP_SHARED && !skip) {
        int b = h->strongest;
        integrate_observation(h, h->logd[0], h->nfree[0], h->nocc[0], h->px[b], h->py[b], h->pt[b], bxy,
                              bdist, bhit, B);
        PHASE_ADD(h, 3);
    }
    h->step++;
    h->have_update = 1;
    if (policy == GMS_RESAMPLE_ALWAYS || (policy == GMS_RESAMPLE_IF_NEFF_LOW && h->neff < h->P / 2))
        do_resample(h, u01);
    return GMS_OK;
}

EXPORT int gms_update(gms_handle *h, const double *beam_xy, const double *beam_dist,
                      const uint8_t *beam_hit, int32_t B, double d_center, double d_theta,
                      const double *normals, double *neff_out) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (B < 0 || (B > 0 && (!beam_xy || !beam_dist || !beam_hit)))
        return fail(h, GMS_ERR_INVALID_ARG, "gms_update: bad beam arrays");
    if (B > GMS_MAX_BEAMS) return fail(h, GMS_ERR_INVALID_ARG, "gms_update: too many beams (GMS_MAX_BEAMS)");
    if (h->cfg.nranks != 1) return fail(h, GMS_ERR_STATE, "gms_update: multi-rank handles use begin/end");
    int rc = update_local(h, beam_xy, beam_dist, beam_hit, B, d_center, d_theta, normals);
    if (rc) return rc;
    memcpy(h->xglobal, h->xlocal, (size_t)h->P * 24);
    rc = update_global(h, beam_xy, beam_dist, beam_hit, B, d_theta, GMS_RESAMPLE_NEVER, 0);
    if (neff_out) *neff_out = h->neff;
    return rc;
}

/* SLAM.resample SLAM.java:133-153 incl. the deep copy Particle(Particle) SLAM.java:41-45 */
static void do_resample(gms_handle *h, double u01) {
    int P = h->P;
    PHASE_T0(h);
    if (u01 < 0) u01 = philox_uniform(h->cfg.seed, h->resample_count);
    h->resample_count++;
    gmsref_resample_indices(h->w, P, u01, h->resample_mode, h->parents);
    /* the Particle object GridMapApp holds as strongestParticle lives on as its first child (same map slot) */
    h->strongest_now = -1;
    for (int m = 0; m < P; m++)
        if (h->parents[m] == h->strongest) { h->strongest_now = m; break; }
    float *nx = malloc(sizeof(float) * P), *ny = malloc(sizeof(float) * P), *nt = malloc(sizeof(float) * P);
    double *nw = malloc(sizeof(double) * P), *nl = malloc(sizeof(double) * P), *nwl = malloc(sizeof(double) * P);
    for (int m = 0; m < P; m++) {
        int p = h->parents[m];
        nx[m] = h->px[p]; ny[m] = h->py[p]; nt[m] = h->pt[p];
        nw[m] = h->w[p]; nl[m] = h->lw[p]; nwl[m] = h->wlit[p];
    }
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->cfg.nranks == 1) {
        /* literal: every child gets a fresh copy of both arrays (GridMap.createMapData(other)
         * GridMap.java:106-124).  Survivors keep their slot; duplicates take the slots of dead parents. */
        size_t n = (size_t)h->W * h->H;
        int32_t *newslot = malloc(sizeof(int32_t) * P);
        char *used = calloc(P, 1);
        for (int m = 0; m < P; m++) used[h->parents[m]] = 1;
        int nfreeslots = 0;
        int32_t *freeslots = malloc(sizeof(int32_t) * P);
        for (int p = 0; p < P; p++)
            if (!used[p]) freeslots[nfreeslots++] = h->slot[p];
        int k = 0;
        for (int m = 0; m < P; m++) {
            int p = h->parents[m];
            if (m == 0 || h->parents[m - 1] != p) newslot[m] = h->slot[p];
            else {
                int d = freeslots[k++], s = h->slot[p];
                memcpy(h->logd[d], h->logd[s], n * sizeof(double));
                memcpy(h->lik[d], h->lik[s], n * sizeof(double));
                memcpy(h->nfree[d], h->nfree[s], n * sizeof(uint32_t));
                memcpy(h->nocc[d], h->nocc[s], n * sizeof(uint32_t));
                newslot[m] = d;
            }
        }
        memcpy(h->slot, newslot, sizeof(int32_t) * P);
        free(newslot); free(used); free(freeslots);
    }
    memcpy(h->px, nx, sizeof(float) * P); memcpy(h->py, ny, sizeof(float) * P); memcpy(h->pt, nt, sizeof(float) * P);
    memcpy(h->w, nw, sizeof(double) * P); memcpy(h->lw, nl, sizeof(double) * P); memcpy(h->wlit, nwl, sizeof(double) * P);
    free(nx); free(ny); free(nt); free(nw); free(nl); free(nwl);
    PHASE_ADD(h, 5);
}

EXPORT int gms_resample(gms_handle *h, double u01) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (u01 >= 1.0) return fail(h, GMS_ERR_INVALID_ARG, "gms_resample: u01 must be < 1");
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->cfg.nranks != 1)
        return fail(h, GMS_ERR_UNSUPPORTED, "oracle: per-particle maps are single-rank only");
    do_resample(h, u01);
    return GMS_OK;
}

EXPORT int gms_calculate_neff(gms_handle *h, double *neff_out) {
    if (!h || !neff_out) return GMS_ERR_INVALID_ARG;
    *neff_out = neff_of(h->w, h->P);
    return GMS_OK;
}

/* SLAM.getWeightedPose SLAM.java:165-178 */
static void weighted_pose(const gms_handle *h, const double *w, float out[3]) {
    double xs = 0, ys = 0, ts = 0, ws = 0;
    for (int i = 0; i < h->P; i++) {
        xs += (double)h->px[i] * w[i];
        ys += (double)h->py[i] * w[i];
        ts += angle_constrain((double)h->pt[i]) * w[i];
        ws += w[i];
    }
    out[0] = (float)(xs / ws); out[1] = (float)(ys / ws); out[2] = (float)(ts / ws);
}
EXPORT int gms_get_weighted_pose(gms_handle *h, float pose[3]) {
    if (!h || !pose) return GMS_ERR_INVALID_ARG;
    weighted_pose(h, h->w, pose);
    return GMS_OK;
}
EXPORT int gms_get_strongest(gms_handle *h, int32_t *index, float pose[3], double *weight) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (!h->have_update) { /* SLAM.reset: strongestParticle = particles.get(0) (SLAM.java:75) */
        if (index) *index = -1;
        if (pose) { pose[0] = pose[1] = pose[2] = 0.f; }
        if (weight) *weight = 1.0 / h->P;
        return GMS_OK;
    }
    if (index) *index = h->strongest_now;
    if (pose) { pose[0] = h->strongest_pose[0]; pose[1] = h->strongest_pose[1]; pose[2] = h->strongest_pose[2]; }
    if (weight) *weight = h->strongest_w;
    return GMS_OK;
}
EXPORT int gms_get_poses(gms_handle *h, float *xyt) {
    if (!h || !xyt) return GMS_ERR_INVALID_ARG;
    for (int i = 0; i < h->P; i++) { xyt[3 * i] = h->px[i]; xyt[3 * i + 1] = h->py[i]; xyt[3 * i + 2] = h->pt[i]; }
    return GMS_OK;
}
EXPORT int gms_get_weights(gms_handle *h, double *w) {
    if (!h || !w) return GMS_ERR_INVALID_ARG;
    memcpy(w, h->w, sizeof(double) * h->P);
    return GMS_OK;
}
EXPORT int gms_get_log_weights(gms_handle *h, double *lw) {
    if (!h || !lw) return GMS_ERR_INVALID_ARG;
    memcpy(lw, h->lw, sizeof(double) * h->P);
    return GMS_OK;
}
EXPORT int gms_get_parents(gms_handle *h, int32_t *parents) {
    if (!h || !parents) return GMS_ERR_INVALID_ARG;
    memcpy(parents, h->parents, sizeof(int32_t) * h->P);
    return GMS_OK;
}

static int slot_of(gms_handle *h, int particle, int *slot) {
    if (h->cfg.map_mode == GMS_MAP_SHARED) { *slot = 0; return 1; }
    if (particle < h->lo || particle >= h->lo + h->cnt) return 0;
    *slot = h->slot[particle - h->lo];
    return 1;
}

EXPORT int gms_get_map(gms_handle *h, int32_t particle, int32_t kind, void *dst, size_t bytes) {
    if (!h || !dst) return GMS_ERR_INVALID_ARG;
    int s;
    if (!slot_of(h, particle, &s)) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: particle not on this handle");
    size_t n = (size_t)h->W * h->H;
    switch (kind) {
    case GMS_MAP_LOG: {
        if (bytes != n * 8) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: size mismatch");
        /* canonical value: closed form of the counts (what libgms returns) */
        double *d = (double *)dst;
        for (size_t i = 0; i < n; i++) d[i] = (double)h->nfree[s][i] * h->l_free + (double)h->nocc[s][i] * h->l_occ;
        return GMS_OK;
    }
    case GMS_MAP_LIKELIHOOD:
        if (bytes != n * 8) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: size mismatch");
        memcpy(dst, h->lik[s], bytes);
        return GMS_OK;
    case GMS_MAP_FREE_COUNT:
    case GMS_MAP_OCC_COUNT:
        if (bytes != n * 4) return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: size mismatch");
        memcpy(dst, kind == GMS_MAP_FREE_COUNT ? h->nfree[s] : h->nocc[s], bytes);
        return GMS_OK;
    default:
        return fail(h, GMS_ERR_INVALID_ARG, "gms_get_map: unknown kind");
    }
}

/* ---- oracle-only views of Java's literal state -------------------------------------------- */
/* logData exactly as Java accumulates it: sequential f64 "+=" in beam-then-ray order. */
EXPORT int gmsref_get_literal_log(gms_handle *h, int32_t particle, double *dst) {
    int s;
    if (!h || !dst || !slot_of(h, particle, &s)) return GMS_ERR_INVALID_ARG;
    memcpy(dst, h->logd[s], sizeof(double) * (size_t)h->W * h->H);
    return GMS_OK;
}
/* Java's weights after SLAM.java:120-121 (NaN when weightSum underflowed to 0), its Neff and its
 * strongest particle (strict > on the raw products). */
EXPORT int gmsref_get_literal_weights(gms_handle *h, double *w, double *neff, int32_t *strongest) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (w) memcpy(w, h->wlit, sizeof(double) * h->P);
    if (neff) *neff = h->neff_lit;
    if (strongest) *strongest = h->strongest_lit;
    return GMS_OK;
}
/* seconds spent per phase since the last call with reset != 0 (threads == 1 only; see struct gms_handle) */
EXPORT int gmsref_phase_seconds(gms_handle *h, double out[6], int32_t reset) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (out) memcpy(out, h->phase, sizeof h->phase);
    if (reset) memset(h->phase, 0, sizeof h->phase);
    return GMS_OK;
}
EXPORT int gmsref_set_threads(gms_handle *h, int32_t n) {
    if (!h || n < 1) return GMS_ERR_INVALID_ARG;
    free(h->prob_scratch); free(h->tmp_scratch);
    h->prob_scratch = h->tmp_scratch = NULL;
    h->threads = n;
    return GMS_OK;
}
/* stand-alone numeric helpers for unit tests */
EXPORT int gmsref_blur(const double *in, double *out, int32_t w, int32_t hgt, const double *kernel, int32_t ktaps) {
    double *tmp = malloc(sizeof(double) * (size_t)w * hgt);
    if (!tmp) return GMS_ERR_OOM;
    blur_separable(in, out, tmp, w, hgt, kernel, ktaps, 1);
    free(tmp);
    return GMS_OK;
}
EXPORT int gmsref_motion(gms_handle *h, float *xyt, double d_center, double d_theta, double zd, double zt) {
    if (!h || !xyt) return GMS_ERR_INVALID_ARG;
    motion_sample(h, &xyt[0], &xyt[1], &xyt[2], d_center, d_theta, zd, zt);
    return GMS_OK;
}
EXPORT int gmsref_philox_normals(uint64_t seed, uint32_t gidx, uint64_t step, double *zd, double *zt) {
    philox_normals(seed, gidx, step, zd, zt);
    return GMS_OK;
}
EXPORT double gmsref_philox_uniform(uint64_t seed, uint64_t n) { return philox_uniform(seed, n); }
EXPORT double gmsref_angle_constrain(double a) { return angle_constrain(a); }

/* ---- state injection ------------------------------------------------------------------------ */
EXPORT int gms_set_poses(gms_handle *h, const float *xyt) {
    if (!h || !xyt) return GMS_ERR_INVALID_ARG;
    for (int i = 0; i < h->P; i++) { h->px[i] = xyt[3 * i]; h->py[i] = xyt[3 * i + 1]; h->pt[i] = xyt[3 * i + 2]; }
    return GMS_OK;
}
EXPORT int gms_set_weights(gms_handle *h, const double *w) {
    if (!h || !w) return GMS_ERR_INVALID_ARG;
    memcpy(h->w, w, sizeof(double) * h->P);
    memcpy(h->wlit, w, sizeof(double) * h->P);
    return GMS_OK;
}
EXPORT int gms_set_map_counts(gms_handle *h, int32_t particle, const uint32_t *nf, const uint32_t *no) {
    int s;
    if (!h || !nf || !no || !slot_of(h, particle, &s)) return GMS_ERR_INVALID_ARG;
    size_t n = (size_t)h->W * h->H;
    memcpy(h->nfree[s], nf, n * 4);
    memcpy(h->nocc[s], no, n * 4);
    for (size_t i = 0; i < n; i++) h->logd[s][i] = (double)nf[i] * h->l_free + (double)no[i] * h->l_occ;
    return GMS_OK;
}

/* ---- GridMap operators ------------------------------------------------------------------------ */
EXPORT int gms_map_apply_measurement(gms_handle *h, int32_t particle, float sx, float sy, float ex,
                                     float ey, float meas, int32_t was_hit) {
    int s;
    if (!h || !slot_of(h, particle, &s)) return GMS_ERR_INVALID_ARG;
    apply_measurement(h, h->logd[s], h->nfree[s], h->nocc[s], sx, sy, ex, ey, meas, was_hit != 0);
    return GMS_OK;
}
EXPORT int gms_map_integrate_observation(gms_handle *h, int32_t particle, const float pose[3],
                                         const double *bxy, const double *bdist, const uint8_t *bhit,
                                         int32_t B) {
    int s;
    if (!h || !pose || !slot_of(h, particle, &s) || B < 0) return GMS_ERR_INVALID_ARG;
    integrate_observation(h, h->logd[s], h->nfree[s], h->nocc[s], pose[0], pose[1], pose[2], bxy, bdist, bhit, B);
    return GMS_OK;
}
EXPORT int gms_map_compute_likelihood(gms_handle *h, int32_t particle) {
    int s;
    if (!h || !slot_of(h, particle, &s)) return GMS_ERR_INVALID_ARG;
    if (!ensure_scratch(h)) return fail(h, GMS_ERR_OOM, "out of memory (scratch)");
    compute_likelihood(h, h->logd[s], h->lik[s], h->prob_scratch, h->tmp_scratch);
    return GMS_OK;
}
EXPORT int gms_map_probability_of(gms_handle *h, int32_t particle, const float pose[3], const double *bxy,
                                  const uint8_t *bhit, int32_t B, double *log_prob, double *prob) {
    int s;
    if (!h || !pose || !slot_of(h, particle, &s) || B < 0) return GMS_ERR_INVALID_ARG;
    double lw, p = probability_of(h, h->lik[s], pose[0], pose[1], pose[2], bxy, bhit, B, &lw);
    if (log_prob) *log_prob = lw;
    if (prob) *prob = p;
    return GMS_OK;
}
EXPORT int gms_trace_rays(gms_handle *h, const float *rays, int32_t num_rays, int32_t extra,
                          int32_t *cells_xy, int32_t cap, int32_t *counts) {
    if (!h || !rays || !counts || num_rays < 0 || cap < 0 || (cap > 0 && !cells_xy)) return GMS_ERR_INVALID_ARG;
    for (int r = 0; r < num_rays; r++) {
        ray_iter it;
        ray_init(&it, h->W, h->H, rays[4 * r], rays[4 * r + 1], rays[4 * r + 2], rays[4 * r + 3], extra);
        int c = 0;
        while (ray_has_next(&it)) {
            int cx, cy;
            ray_next(&it, &cx, &cy);
            if (c < cap) { cells_xy[2 * ((size_t)cap * r + c)] = cx; cells_xy[2 * ((size_t)cap * r + c) + 1] = cy; }
            c++;
        }
        counts[r] = c;
    }
    return GMS_OK;
}
/* Odometry(int,int) Odometry.java:41-55; MathUtil.PI is FLOAT pi (MathUtil.java:21), Robot.java:8-14 */
EXPORT int gms_odometry_from_counts(int32_t left, int32_t right, double *d_center, double *d_theta) {
    if (!d_center || !d_theta) return GMS_ERR_INVALID_ARG;
    double pif = (double)(float)M_PI;
    double dl = (double)left / 960 * pif * 0.063;
    double dr = (double)right / 960 * pif * 0.063;
    *d_center = (dl + dr) / 2;
    *d_theta = (dr - dl) / 0.22;
    return GMS_OK;
}

/* ---- begin/end split ("device" pointers are host pointers for the oracle) --------------------- */
EXPORT int gms_exchange_buffers(gms_handle *h, void **dl, size_t *lb, void **dg, size_t *gb) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (dl) *dl = h->xlocal;
    if (lb) *lb = (size_t)h->cnt * 24;
    if (dg) *dg = h->xglobal;
    if (gb) *gb = (size_t)h->P * 24;
    return GMS_OK;
}
EXPORT int gms_update_begin_dev(gms_handle *h, const double *bxy, const double *bdist, const uint8_t *bhit,
                                int32_t B, double d_center, double d_theta, const double *normals) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (B < 0 || (B > 0 && (!bxy || !bdist || !bhit))) return fail(h, GMS_ERR_INVALID_ARG, "bad beam arrays");
    if (B > GMS_MAX_BEAMS) return fail(h, GMS_ERR_INVALID_ARG, "too many beams (GMS_MAX_BEAMS)");
    if (h->cfg.map_mode == GMS_MAP_PER_PARTICLE && h->cfg.nranks != 1)
        return fail(h, GMS_ERR_UNSUPPORTED, "oracle: per-particle maps are single-rank only");
    free(h->pend_xy); free(h->pend_dist); free(h->pend_hit);
    h->pend_xy = malloc(sizeof(double) * 2 * (B + 1)); h->pend_dist = malloc(sizeof(double) * (B + 1));
    h->pend_hit = malloc(B + 1);
    memcpy(h->pend_xy, bxy, sizeof(double) * 2 * B); memcpy(h->pend_dist, bdist, sizeof(double) * B);
    memcpy(h->pend_hit, bhit, B);
    h->pend_B = B; h->pend_dtheta = d_theta;
    return update_local(h, bxy, bdist, bhit, B, d_center, d_theta, normals);
}
EXPORT int gms_update_end_dev(gms_handle *h, int32_t policy, double u01) {
    if (!h || !h->pend_xy) return GMS_ERR_STATE;
    if (h->cfg.nranks == 1) memcpy(h->xglobal, h->xlocal, (size_t)h->P * 24);
    return update_global(h, h->pend_xy, h->pend_dist, h->pend_hit, h->pend_B, h->pend_dtheta, policy, u01);
}
EXPORT int gms_step_dev(gms_handle *h, const double *bxy, const double *bdist, const uint8_t *bhit, int32_t B,
                        double d_center, double d_theta, const double *normals, int32_t policy, double u01) {
    int rc = gms_update_begin_dev(h, bxy, bdist, bhit, B, d_center, d_theta, normals);
    if (rc) return rc;
    return gms_update_end_dev(h, policy, u01);
}
EXPORT int gms_read_neff(gms_handle *h, double *neff) {
    if (!h || !neff) return GMS_ERR_INVALID_ARG;
    *neff = h->neff;
    return GMS_OK;
}
EXPORT int gms_sync(gms_handle *h) { return h ? GMS_OK : GMS_ERR_INVALID_ARG; }
EXPORT int gms_join_streams(gms_handle *h) { return h ? GMS_OK : GMS_ERR_INVALID_ARG; }
EXPORT int gms_set_pose_optimizer(gms_handle *h, gms_pose_optimizer_fn fn, void *user) {
    if (!h) return GMS_ERR_INVALID_ARG;
    h->opt_fn = fn;
    h->opt_user = user;
    return GMS_OK;
}
EXPORT int gms_set_stream(gms_handle *h, void *s) { (void)s; return h ? GMS_OK : GMS_ERR_INVALID_ARG; }
EXPORT int gms_profile_enable(gms_handle *h, int32_t on) { (void)on; return h ? GMS_OK : GMS_ERR_INVALID_ARG; }
EXPORT int gms_profile_read(gms_handle *h, double *ms, int64_t *launches) {
    if (!h) return GMS_ERR_INVALID_ARG;
    for (int i = 0; i < GMS_PHASE_COUNT; i++) { if (ms) ms[i] = 0; if (launches) launches[i] = 0; }
    return GMS_OK;
}
EXPORT int gms_profile_reset(gms_handle *h) { return h ? GMS_OK : GMS_ERR_INVALID_ARG; }
EXPORT int gms_launch_count(gms_handle *h, int64_t *n) { if (!h || !n) return GMS_ERR_INVALID_ARG; *n = 0; return GMS_OK; }
EXPORT int gms_ipc_export(gms_handle *h, void *handles) { (void)handles; return h ? fail(h, GMS_ERR_UNSUPPORTED, "oracle: no device arenas") : GMS_ERR_INVALID_ARG; }
EXPORT int gms_ipc_import(gms_handle *h, const void *all) { (void)all; return h ? fail(h, GMS_ERR_UNSUPPORTED, "oracle: no device arenas") : GMS_ERR_INVALID_ARG; }

/* ---- rows adjacent to the path ---------------------------------------------------- */
/* GridMapApp.onHandleData GridMapApp.java:140-175 + Measurement(x, y, wasHit, 0) Observation.java:69-76 */
static void deskew(const double *angle, const double *dist, int n, double d_center, double d_theta, double *xy, double *od) {
    for (int i = 0; i < n; i++) {
        double d_i = -(n - i) / (double)n;
        double delta_theta = d_theta * d_i;
        double delta_x = d_center * d_i;
        double x_a = dist[i] * cos_fixed(angle[i] + delta_theta) + delta_x;
        double y_a = dist[i] * sin_fixed(angle[i] + delta_theta);
        xy[2 * i] = x_a; xy[2 * i + 1] = y_a;
        od[i] = sqrt(x_a * x_a + y_a * y_a);
    }
}
EXPORT int gms_deskew(gms_handle *h, const double *angle, const double *dist, int32_t n, double d_center, double d_theta,
                      double *out_xy, double *out_dist) {
    if (!h || n < 0 || (n > 0 && (!angle || !dist || !out_xy || !out_dist))) return GMS_ERR_INVALID_ARG;
    deskew(angle, dist, n, d_center, d_theta, out_xy, out_dist);
    return GMS_OK;
}
EXPORT int gms_update_raw(gms_handle *h, const double *angle, const double *dist, const uint8_t *hit, int32_t n,
                          double d_center, double d_theta, const double *normals, double *neff_out) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (n < 0 || (n > 0 && (!angle || !dist || !hit))) return fail(h, GMS_ERR_INVALID_ARG, "gms_update_raw: bad arrays");
    if (n > GMS_MAX_BEAMS) return fail(h, GMS_ERR_INVALID_ARG, "gms_update_raw: too many beams (GMS_MAX_BEAMS)");
    double *xy = malloc(sizeof(double) * 2 * (n + 1)), *od = malloc(sizeof(double) * (n + 1));
    deskew(angle, dist, n, d_center, d_theta, xy, od);
    int rc = gms_update(h, xy, od, hit, n, d_center, d_theta, normals, neff_out);
    free(xy); free(od);
    return rc;
}
static double inv_log_odds(double l) { return 1.0 - 1.0 / (1 + exp(l)); } /* Util.java:46-48 */
/* GridMap.render GridMap.java:371-388, Util.getColorBitsGrayscale Util.java:106-108, Color.java:62-66 */
EXPORT int gms_render_map(gms_handle *h, int32_t particle, int32_t likelihood, uint32_t *out) {
    int s;
    if (!h || !out || !slot_of(h, particle, &s)) return GMS_ERR_INVALID_ARG;
    size_t n = (size_t)h->W * h->H;
    for (size_t i = 0; i < n; i++) {
        float value = likelihood ? (float)h->lik[s][i] : (float)(1.0 - inv_log_odds(h->logd[s][i]));
        int idx = (int)(value * 255);
        if (idx < 0) idx = 0;
        if (idx > 255) idx = 255;
        float ratio = idx / (float)256;
        uint32_t c = (uint32_t)(int)(255 * ratio);
        out[i] = ((255u << 24) | (c << 16) | (c << 8) | c) & 0xfeffffffu;
    }
    return GMS_OK;
}
/* GridMapApp.calculateCombined GridMapApp.java:439-458 */
EXPORT int gms_combined_map_begin_dev(gms_handle *h, void **d_product, size_t *bytes) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (h->cfg.map_mode != GMS_MAP_PER_PARTICLE)
        return fail(h, GMS_ERR_UNSUPPORTED, "gms_combined_map: per-particle maps only");
    size_t n = (size_t)h->W * h->H;
    if (!h->comb_prod) h->comb_prod = malloc(n * sizeof(double));
    if (!h->comb_prod) return fail(h, GMS_ERR_OOM, "out of memory (combined map)");
    for (size_t i = 0; i < n; i++) { /* GridMapApp.java:447-452, particle order */
        double product = 1;
        for (int p = 0; p < h->cnt; p++) product *= 1 - inv_log_odds(h->logd[h->slot[p]][i]);
        h->comb_prod[i] = product;
    }
    if (d_product) *d_product = h->comb_prod;
    if (bytes) *bytes = n * sizeof(double);
    return GMS_OK;
}
EXPORT int gms_combined_map_end(gms_handle *h, double *log_out, double *lik_out) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (!h->comb_prod) return fail(h, GMS_ERR_STATE, "gms_combined_map_end without begin");
    if (!ensure_scratch(h)) return fail(h, GMS_ERR_OOM, "out of memory (scratch)");
    size_t n = (size_t)h->W * h->H;
    double *comb = malloc(n * sizeof(double)), *lik = malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) comb[i] = log_odds(1 - h->comb_prod[i]); /* GridMapApp.java:454 */
    compute_likelihood(h, comb, lik, h->prob_scratch, h->tmp_scratch);     /* GridMapApp.java:457 */
    if (log_out) memcpy(log_out, comb, n * sizeof(double));
    if (lik_out) memcpy(lik_out, lik, n * sizeof(double));
    free(comb); free(lik);
    return GMS_OK;
}
EXPORT int gms_combined_map(gms_handle *h, double *log_out, double *lik_out) {
    if (!h) return GMS_ERR_INVALID_ARG;
    if (h->cfg.nranks != 1) return fail(h, GMS_ERR_STATE, "gms_combined_map: multi-rank handles use begin / all-reduce / end");
    int rc = gms_combined_map_begin_dev(h, NULL, NULL);
    if (rc) return rc;
    return gms_combined_map_end(h, log_out, lik_out);
}
