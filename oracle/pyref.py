"""pyref.py — second, independent restatement of the reference's hot path in pure Python.

TEST INFRASTRUCTURE (see oracle/gms_ref.c header): used only by oracle/make_golden.py to produce
tests/golden/*.npz, which pin the C oracle.  PARITY UNPINNED by the reference itself (it has no
tests and cannot run here); two independently written restatements that agree bit-for-bit, plus
the hand-derived vectors of SURVEY.md Appendix B, are the best pin available.

Arithmetic model: Python floats are IEEE binary64.  A Java `float` is held as a Python float whose
value is representable in binary32; every float-typed operation is computed in binary64 and
rounded with f32(), which equals the correctly rounded binary32 operation for + - * / sqrt
(53 >= 2*24+2, so the double rounding is innocuous).  No fused multiply-add anywhere.

References are to java/GridMapGL/src/main/java/com/fmsz/gridmapgl/.
"""
import math
import struct

import numpy as np

INF = float("inf")


def f32(x):
    with np.errstate(all="ignore"):
        return float(np.float32(x))


def d2i(d):
    """Java (int) cast of a double, JLS 5.1.3."""
    if d != d:
        return 0
    if d >= 2147483647.0:
        return 2147483647
    if d <= -2147483648.0:
        return -2147483648
    return int(d)  # truncates toward zero


def angle_constrain(a):  # MathUtil.java:65-72
    while a < math.pi:
        a += math.pi * 2
    while a > math.pi:
        a -= math.pi * 2
    return a


# FastMath.sin / cos (commons-math3, not in the reference tree) -> the published fdlibm 5.3 scheme (e_rem_pio2.c medium
# reduction, k_sin.c, k_cos.c; < 1 ulp) as a fixed sequence of IEEE operations — the same sequence oracle/gms_ref.c and
# the CUDA library evaluate, so all three agree bit for bit (Python floats are IEEE f64, no contraction).
_H = float.fromhex
_TS = [_H(v) for v in ("-0x1.5555555555549p-3", "0x1.111111110f8a6p-7", "-0x1.a01a019c161d5p-13",
                       "0x1.71de357b1fe7dp-19", "-0x1.ae5e68a2b9cebp-26", "0x1.5d93a5acfd57cp-33")]
_TC = [_H(v) for v in ("0x1.555555555554cp-5", "-0x1.6c16c16c15177p-10", "0x1.a01a019cb1590p-16",
                       "-0x1.27e4f809c52adp-22", "0x1.1ee9ebdb4b1c4p-29", "-0x1.8fae9be8838d4p-37")]
_INV_PIO2 = _H("0x1.45f306dc9c883p-1")
_PIO2 = [(_H("0x1.921fb54400000p+0"), _H("0x1.0b4611a626331p-34")), (_H("0x1.0b4611a600000p-34"), _H("0x1.3198a2e037073p-69")),
         (_H("0x1.3198a2e000000p-69"), _H("0x1.b839a252049c1p-104"))]
_PIO4 = _H("0x1.921fb54442d18p-1")


def _exponent_field(x):
    return (struct.unpack("<Q", struct.pack("<d", x))[0] >> 52) & 0x7FF


def _poly_sin(x, tail):
    z = x * x
    v = z * x
    r = _TS[1] + z * (_TS[2] + z * (_TS[3] + z * (_TS[4] + z * _TS[5])))
    return x - ((z * (0.5 * tail - v * r) - tail) - v * _TS[0])


def _poly_cos(x, tail):
    z = x * x
    zz = z * z
    r = z * (_TC[0] + z * (_TC[1] + z * _TC[2])) + (zz * zz) * (_TC[3] + z * (_TC[4] + z * _TC[5]))
    hz = 0.5 * z
    w = 1.0 - hz
    return w + (((1.0 - w) - hz) + (z * r - x * tail))


def _reduce_pio2(x):
    if abs(x) <= _PIO4:
        return 0, x, 0.0
    n = round(x * _INV_PIO2)  # ties to even, like rint
    fn = float(n)
    r = x - fn * _PIO2[0][0]
    w = fn * _PIO2[0][1]
    y = r - w
    ex = _exponent_field(x)
    for (hi, lo), gap in ((_PIO2[1], 16), (_PIO2[2], 49)):
        if ex - _exponent_field(y) <= gap:
            break
        t = r
        w = fn * hi
        r = t - w
        w = fn * lo - ((t - r) - w)
        y = r - w
    return n, y, (r - y) - w


def sin_fixed(x):
    x = float(x)
    if not abs(x) < 1647099.0:
        return math.sin(x) if math.isfinite(x) else math.nan
    n, a, b = _reduce_pio2(x)
    q = n & 3
    return (_poly_sin(a, b), _poly_cos(a, b), -_poly_sin(a, b), -_poly_cos(a, b))[q]


def cos_fixed(x):
    x = float(x)
    if not abs(x) < 1647099.0:
        return math.cos(x) if math.isfinite(x) else math.nan
    n, a, b = _reduce_pio2(x)
    q = n & 3
    return (_poly_cos(a, b), -_poly_sin(a, b), -_poly_cos(a, b), _poly_sin(a, b))[q]


def cos_f(theta_f):  # MathUtil.java:36-38 (float overload)
    return f32(cos_fixed(theta_f))


def sin_f(theta_f):  # MathUtil.java:30-32
    return f32(sin_fixed(theta_f))


def log_odds(p):  # Util.java:35-37
    return math.log(p / (1.0 - p))


P_FREE = f32(0.30)  # SensorModel.java:23
P_OCC = f32(0.9)  # SensorModel.java:24


def gaussian_kernel(sigma, size):  # Util.java:428-455
    norm = 1.0 / (math.sqrt(2 * math.pi) * sigma)
    coeff = 2 * sigma * sigma
    vals, total = [], 0.0
    for x in range(-size, size + 1):
        g = norm * math.exp((-x * x) / coeff)
        vals.append(g)
        total += g
    return [v / total for v in vals]


def blur(inp, width, height, kernel):  # Util.java:378-426
    k = (len(kernel) - 1) // 2
    out = [0.0] * (width * height)
    for y in range(height):
        for x in range(width):
            total = 0.0
            for i in range(-k, k + 1):
                x2 = x + i
                if 0 <= x2 < width:
                    total += kernel[i + k] * inp[y * width + x2]
            out[y * width + x] = total
    tmp = list(out)
    for y in range(height):
        for x in range(width):
            total = 0.0
            for i in range(-k, k + 1):
                y2 = y + i
                if 0 <= y2 < height:
                    total += kernel[i + k] * tmp[x + y2 * width]
            out[x + y * width] = total
    return out


def ray_cells(W, H, x0, y0, x1, y1, extra):
    """RayIterator.java:65-130.  Arguments are binary32-valued.  Returns (cells, n_init, err_init)."""
    dx = f32(abs(f32(x1 - x0)))
    dy = f32(abs(f32(y1 - y0)))
    x = d2i(math.floor(x0))
    y = d2i(math.floor(y0))
    n = 1 + extra
    if dx == 0:
        xi, err = 0, INF
    elif x1 > x0:
        xi = 1
        n += d2i(math.floor(x1) - x)
        err = f32((math.floor(x0) + 1 - x0) * dy)
    else:
        xi = -1
        n += x - d2i(math.floor(x1))
        err = f32((x0 - math.floor(x0)) * dy)
    if dy == 0:
        yi = 0
        err = f32(err - INF)
    elif y1 > y0:
        yi = 1
        n += d2i(math.floor(y1)) - y
        err = f32(err - (math.floor(y0) + 1 - y0) * dx)
    else:
        yi = -1
        n += y - d2i(math.floor(y1))
        err = f32(err - (y0 - math.floor(y0)) * dx)
    n0, e0 = n, err
    cells = []
    while n > 0 and not (x < 0 or x >= W or y < 0 or y >= H):
        cells.append((x, y))
        if err > 0:
            y += yi
            err = f32(err - dx)
        else:
            x += xi
            err = f32(err + dy)
        n -= 1
    return cells, n0, e0


def inverse_sensor_class(cur, meas, hit, tol):  # SensorModel.java:31-41; 0 prior, 1 free, 2 occupied
    if not hit:
        return 1 if cur < meas else 0
    half = f32(tol / 2)
    if cur < f32(meas - half):
        return 1
    if cur > f32(meas + half):
        return 0
    return 2


class PyGridMap:
    """GridMap.java — geometry + per-map operators (maps are dicts of flat lists)."""

    def __init__(self, width, height, resolution, pos, max_range=10.0, z_hit=0.9, tol=2.0, extra=2):
        self.res = f32(resolution)
        self.posx, self.posy = f32(pos[0]), f32(pos[1])
        self.W = d2i(math.ceil(f32(f32(width) / self.res)))  # GridMap.java:85
        self.H = d2i(math.ceil(f32(f32(height) / self.res)))
        sigma = math.sqrt(0.05 / self.res)  # GridMap.java:94
        self.kernel = gaussian_kernel(sigma, d2i(math.ceil(sigma * 3)))
        self.max_range = f32(max_range)
        self.z_hit = z_hit
        self.tol = f32(tol)
        self.extra = extra
        self.l_free = log_odds(P_FREE)
        self.l_occ = log_odds(P_OCC)

    def create_map(self):  # GridMap.java:106-117
        n = self.W * self.H
        return {"log": [log_odds(0.5)] * n, "lik": [0.0] * n, "nfree": [0] * n, "nocc": [0] * n}

    def copy_map(self, m):  # GridMap.java:118-124
        return {k: list(v) for k, v in m.items()}

    def compute_likelihood(self, m):  # GridMap.java:233-250
        prob = [1.0 if v > 0.0 else (0.0 if v < 0.0 else 0.5) for v in m["log"]]
        m["lik"] = blur(prob, self.W, self.H, self.kernel)

    def apply_measurement(self, m, sx, sy, ex, ey, meas, hit, trace=None):  # GridMap.java:194-228
        cells, _, _ = ray_cells(self.W, self.H, f32(sx + 0.5), f32(sy + 0.5), f32(ex + 0.5), f32(ey + 0.5),
                                self.extra)
        for cx, cy in cells:
            dX = f32(sx - f32(float(cx) + 0.5))
            dY = f32(sy - f32(float(cy) + 0.5))
            dist = f32(math.sqrt(f32(f32(dX * dX) + f32(dY * dY))))
            c = inverse_sensor_class(dist, meas, hit, self.tol)
            idx = cx + cy * self.W
            if c == 1:
                m["log"][idx] += self.l_free
                m["nfree"][idx] += 1
            elif c == 2:
                m["log"][idx] += self.l_occ
                m["nocc"][idx] += 1
            else:
                m["log"][idx] += 0.0
            if trace is not None:
                trace.append((cx, cy, c))

    def _xform(self, pose):  # Transform.java:13-32
        c, s = cos_f(pose[2]), sin_f(pose[2])
        px, py = pose[0], pose[1]
        return (lambda x, y: x * c - y * s + px), (lambda x, y: x * s + y * c + py)

    def integrate_observation(self, m, beams, pose):  # GridMap.java:173-191
        tx, ty = self._xform(pose)
        sx = f32((tx(0.0, 0.0) - self.posx) / self.res)
        sy = f32((ty(0.0, 0.0) - self.posy) / self.res)
        for (lx, ly, dist, hit) in beams:
            ex = f32((tx(lx, ly) - self.posx) / self.res)
            ey = f32((ty(lx, ly) - self.posy) / self.res)
            meas = f32(f32(dist) / self.res)
            self.apply_measurement(m, sx, sy, ex, ey, meas, hit)

    def probability_of(self, m, beams, pose):  # GridMap.java:261-294 -> (product, sum of logs)
        product, lsum = 1.0, 0.0
        tx, ty = self._xform(pose)
        z_random = 1 - self.z_hit
        for (lx, ly, dist, hit) in beams:
            if not hit:
                continue
            gx = d2i((tx(lx, ly) - self.posx) / self.res)
            gy = d2i((ty(lx, ly) - self.posy) / self.res)
            if not (gx < 0 or gy < 0 or gx >= self.W or gy >= self.H):
                val = m["lik"][gx + gy * self.W]
                if val == 0.5:
                    f = 1.0 / self.max_range
                else:
                    f = self.z_hit * val + z_random * 1.0 / self.max_range
                product *= f
                lsum += math.log(f)
        return product, lsum


def motion_sample(pose, d_center, d_theta, zd, zt):
    """SLAM.java:155-163, Odometry.java:60-69,77-96; NormalDistribution.sample() = sd*z + mean."""
    sd_c = (0.01 + abs(d_center) * 0.05) / 2
    sd_t = 5 * (math.pi / 180.0) + 0.1 * abs(d_theta)
    d = sd_c * zd + d_center
    th = sd_t * zt + d_theta
    theta = f32(angle_constrain(pose[2] + th))
    x = f32(pose[0] + cos_f(theta) * d)
    y = f32(pose[1] + sin_f(theta) * d)
    return (x, y, theta)


def resample_indices(w, u01):  # SLAM.java:133-153 (+ clamp at n-1; Java would throw)
    n = len(w)
    r = u01 * 1.0 / n
    c, i, out = w[0], 0, []
    for m in range(1, n + 1):
        U = r + (m - 1) * 1.0 / n
        while U > c and i < n - 1:
            i += 1
            c += w[i]
        out.append(i)
    return out


def resample_indices_fixed(w, u01):
    """canonical associative variant: u64 fixed point, trunc(w*2^60) and trunc(U*2^60)."""
    n = len(w)
    r = u01 * 1.0 / n
    q = [int(x * 2.0 ** 60) for x in w]
    c, i, out = q[0], 0, []
    for m in range(1, n + 1):
        U = int((r + (m - 1) * 1.0 / n) * 2.0 ** 60)
        while U > c and i < n - 1:
            i += 1
            c += q[i]
        out.append(i)
    return out


def neff(w):  # SLAM.java:180-190
    s = 0.0
    for x in w:
        s += x
    sq = 0.0
    for x in w:
        sq += (x / s) * (x / s)
    return 1.0 / sq


class PySLAM:
    """SLAM.java — per-particle maps (the reference's mode) or the shared-map extension."""

    def __init__(self, n, gm, shared=False):
        self.gm, self.n, self.shared = gm, n, shared
        self.poses = [(0.0, 0.0, 0.0)] * n  # SLAM.java:65-77
        self.w = [1.0 / n] * n
        self.wlit = [1.0 / n] * n
        self.lw = [0.0] * n
        self.maps = [gm.create_map() for _ in range(1 if shared else n)]
        self.strongest = 0
        self.strongest_lit = 0

    def update(self, beams, d_center, d_theta, normals):  # SLAM.java:80-131
        skip = abs(d_theta) > (math.pi / 180.0) * 30
        gm = self.gm
        if self.shared:
            gm.compute_likelihood(self.maps[0])
        prods = []
        for i in range(self.n):
            self.poses[i] = motion_sample(self.poses[i], d_center, d_theta, normals[2 * i], normals[2 * i + 1])
            m = self.maps[0 if self.shared else i]
            if not self.shared:
                gm.compute_likelihood(m)
            p, l = gm.probability_of(m, beams, self.poses[i])
            prods.append(p)
            self.lw[i] = l
            if not self.shared and not skip:
                gm.integrate_observation(m, beams, self.poses[i])
        # literal normalisation (SLAM.java:87-121)
        wsum, best = 0.0, 0
        for i, p in enumerate(prods):
            wsum += p
            if i > 0 and p > prods[best]:
                best = i
        self.strongest_lit = best
        with np.errstate(all="ignore"):
            self.wlit = [float(np.float64(p) / np.float64(wsum)) for p in prods]
        # canonical normalisation (log domain)
        cb = 0
        for i in range(1, self.n):
            if self.lw[i] > self.lw[cb]:
                cb = i
        e = [math.exp(l - self.lw[cb]) for l in self.lw]
        s = 0.0
        for x in e:
            s += x
        self.w = [x / s for x in e]
        self.strongest = cb
        if self.shared and not skip:
            gm.integrate_observation(self.maps[0], beams, self.poses[cb])
        return neff(self.w)

    def resample(self, u01, fixed=False):  # SLAM.java:133-153
        idx = resample_indices_fixed(self.w, u01) if fixed else resample_indices(self.w, u01)
        self.poses = [self.poses[i] for i in idx]
        self.w = [self.w[i] for i in idx]
        self.wlit = [self.wlit[i] for i in idx]
        self.lw = [self.lw[i] for i in idx]
        if not self.shared:
            self.maps = [self.gm.copy_map(self.maps[i]) for i in idx]
        return idx

    def weighted_pose(self):  # SLAM.java:165-178
        xs = ys = ts = ws = 0.0
        for (x, y, t), w in zip(self.poses, self.w):
            xs += x * w
            ys += y * w
            ts += angle_constrain(t) * w
            ws += w
        return (f32(xs / ws), f32(ys / ws), f32(ts / ws))
