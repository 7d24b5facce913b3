"""Backend-generic parity checks: each takes a binding.Library (CPU oracle or CUDA product) and
asserts it against the hand-derived Appendix B vectors or the pyref golden fixtures.

Tolerance classes (DESIGN.md "Parity"):
  bit-exact : ray cell sequences, per-cell (nFree, nOcc), parent indices, poses (f32), likelihood
              field (f64, same operation order), strongest index
  toleranced: log-weights |d| <= 1e-9 (lane-product + log vs sequential sum of logs),
              normalised weights rel 1e-9, Neff rel 1e-9, log-odds |d| <= 1e-9*(n+1),
              weighted pose |d| <= 1e-6
"""
import math
import os

import numpy as np

import appendix_b as AB
from gridmap_slam_robot_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LW_TOL = 1e-9
W_RTOL = 1e-9


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def small_handle(lib, P=1, mode=B.MAP_PER_PARTICLE, **kw):
    """60 x 50 map of the pyref fixtures: 3.0 m x 2.5 m at 0.05 m, origin (-1.5, -1.25)."""
    return lib.create(num_particles=P, map_width_m=3.0, map_height_m=2.5, resolution=0.05, origin_x=-1.5,
                      origin_y=-1.25, map_mode=mode, **kw)


def grid10(lib):
    return lib.create(num_particles=1, map_width_m=0.5, map_height_m=0.5, resolution=0.05, origin_x=0.0,
                      origin_y=0.0)


# ---------------------------------------------------------------- Appendix B (hand-derived) ----
def check_constants(lib):
    h = lib.create(num_particles=2)
    assert (h.W, h.H) == (120, 120)
    assert h.info.l_free == AB.L_FREE
    assert h.info.l_occ == AB.L_OCC
    assert h.info.kernel_taps == 7
    k = list(h.info.kernel[:7])
    assert k[:4] == AB.KERNEL_HALF and k[4:] == AB.KERNEL_HALF[2::-1]
    h.close()
    for width, cells in AB.GRID_SIZES:
        h = lib.create(num_particles=1, map_width_m=width, map_height_m=0.05, resolution=0.05,
                       map_mode=B.MAP_SHARED)
        assert h.W == cells and h.H == 1
        h.close()


def check_appendix_b_rays(lib):
    h = grid10(lib)
    assert (h.W, h.H) == (10, 10)
    for rid, s, e, n, err0, cells, cls_hit, cls_miss in AB.RAYS:
        ray = np.asarray([s[0] + 0.5, s[1] + 0.5, e[0] + 0.5, e[1] + 0.5], np.float32)
        got, cnt = h.trace_rays(ray, extra=2)
        assert cnt[0] == len(cells), rid
        assert [tuple(c) for c in got[0, : cnt[0]]] == cells, rid
        meas = float(np.float32(math.hypot(e[0] - s[0], e[1] - s[1])))
        for hit, classes in ((1, cls_hit), (0, cls_miss)):
            h.reset()
            h.map_apply_measurement(0, s[0], s[1], e[0], e[1], meas, hit)
            nf, no = h.get_map(0, B.MAP_FREE_COUNT), h.get_map(0, B.MAP_OCC_COUNT)
            ef, eo = np.zeros((10, 10), np.uint32), np.zeros((10, 10), np.uint32)
            for (cx, cy), c in zip(cells, classes):
                if c == "F":
                    ef[cy, cx] += 1
                elif c == "O":
                    eo[cy, cx] += 1
            assert np.array_equal(nf, ef) and np.array_equal(no, eo), (rid, hit)
    h.close()


def check_appendix_b_resample(lib):
    for w, u, parents in AB.RESAMPLE:
        h = lib.create(num_particles=len(w), map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED,
                       resample_mode=B.RESAMPLE_LITERAL)
        h.set_weights(w)
        h.resample(u)
        assert h.parents().tolist() == parents, (w, u)
        h.close()
    w, u = AB.RESAMPLE_CLAMP
    h = lib.create(num_particles=4, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED,
                   resample_mode=B.RESAMPLE_LITERAL)
    h.set_weights(w)
    h.resample(u)  # Java would throw; the build clamps the walk at N-1
    p = h.parents().tolist()
    assert p[-1] == 3 and all(a <= b for a, b in zip(p, p[1:]))
    h.set_weights(AB.NEFF[0])
    assert h.calculate_neff() == AB.NEFF[1] or abs(h.calculate_neff() - AB.NEFF[1]) < 1e-12
    h.close()


def check_blank_likelihood(lib):
    """B3: blank map -> exactly 0.5 at >= 3 cells from every border, strictly less within 3 cells."""
    h = small_handle(lib)
    h.map_compute_likelihood(0)
    lik = h.get_map(0, B.MAP_LIKELIHOOD)
    assert np.all(lik[3:-3, 3:-3] == 0.5)
    border = np.ones_like(lik, bool)
    border[3:-3, 3:-3] = False
    assert np.all(lik[border] < 0.5)
    # single occupied cell: 0.5 + 0.5*k[i]*k[j] on its 7x7 support
    nf, no = np.zeros((h.H, h.W), np.uint32), np.zeros((h.H, h.W), np.uint32)
    no[20, 30] = 1
    h.set_map_counts(0, nf, no)
    h.map_compute_likelihood(0)
    lik = h.get_map(0, B.MAP_LIKELIHOOD)
    k = np.asarray(h.info.kernel[:7])
    np.testing.assert_allclose(lik[17:24, 27:34], 0.5 + 0.5 * np.outer(k, k), rtol=0, atol=1e-15)
    assert np.all(lik[3:-3, 3:-3][(np.abs(np.arange(3, h.H - 3)[:, None] - 20) > 3) |
                                  (np.abs(np.arange(3, h.W - 3)[None, :] - 30) > 3)] == 0.5)
    h.close()


# ---------------------------------------------------------------- pyref fixtures ----------------
def check_rays(lib):
    g = golden("pyref_rays.npz")
    W, H = int(g["W"]), int(g["H"])
    h = lib.create(num_particles=1, map_width_m=W * 0.05, map_height_m=H * 0.05, resolution=0.05,
                   map_mode=B.MAP_SHARED)
    assert (h.W, h.H) == (W, H)
    cells, counts = h.trace_rays(g["rays"], extra=int(g["extra"]), cap=g["cells"].shape[1])
    assert np.array_equal(counts, g["counts"])
    assert np.array_equal(cells, g["cells"])
    # a truncated capacity still reports the full count
    cells2, counts2 = h.trace_rays(g["rays"], extra=int(g["extra"]), cap=5)
    assert np.array_equal(counts2, g["counts"])
    assert np.array_equal(cells2, g["cells"][:, :5])
    h.close()


def check_apply(lib):
    g = golden("pyref_apply.npz")
    h = small_handle(lib)
    for a, hit in zip(g["args"], g["hits"]):
        h.map_apply_measurement(0, *[float(v) for v in a], int(hit))
    assert np.array_equal(h.get_map(0, B.MAP_FREE_COUNT), g["nfree"])
    assert np.array_equal(h.get_map(0, B.MAP_OCC_COUNT), g["nocc"])
    n = g["nfree"].astype(np.float64) + g["nocc"]
    assert np.all(np.abs(h.get_map(0, B.MAP_LOG) - g["log"]) <= 1e-9 * (n + 1))
    h.close()


def check_blur(lib):
    g = golden("pyref_blur.npz")
    h = small_handle(lib)
    assert np.array_equal(np.asarray(h.info.kernel[: h.info.kernel_taps]), g["kernel"])
    h.set_map_counts(0, g["nfree"], g["nocc"])
    h.map_compute_likelihood(0)
    lik = h.get_map(0, B.MAP_LIKELIHOOD)
    assert np.array_equal(lik, g["lik"])  # bit-exact f64: same operation order, no FMA
    h.close()


def check_resample(lib):
    g = golden("pyref_resample.npz")
    for k in range(int(g["num_cases"])):
        w, u = g[f"w{k}"], float(g[f"u{k}"])
        for mode, key in ((B.RESAMPLE_LITERAL, "lit"), (B.RESAMPLE_FIXED, "fix")):
            h = lib.create(num_particles=len(w), map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED,
                           resample_mode=mode)
            h.set_weights(w)
            assert abs(h.calculate_neff() / float(g[f"neff{k}"]) - 1) < 1e-12
            h.resample(u)
            assert np.array_equal(h.parents(), g[f"{key}{k}"]), (k, key)
            assert np.array_equal(h.weights(), w[g[f"{key}{k}"]])  # weights are copied, not reset (SLAM.java:42)
            h.close()


def check_slam(lib, shared, via_dev=None):
    """Replay of the pyref SLAM fixture.  `via_dev(handle, xy, dist, hit, dc, dth, normals)` lets the
    CUDA tests route the update through the device-resident entry points instead of gms_update."""
    g = golden("pyref_slam_shared.npz" if shared else "pyref_slam_pp.npz")
    P, steps = int(g["P"]), int(g["steps"])
    h = small_handle(lib, P=P, mode=B.MAP_SHARED if shared else B.MAP_PER_PARTICLE,
                     resample_mode=B.RESAMPLE_LITERAL)
    idx0, _, _ = h.strongest()
    assert idx0 == -1
    for s in range(steps):
        args = (g[f"xy{s}"], g[f"dist{s}"], g[f"hit{s}"], float(g[f"dc{s}"]), float(g[f"dth{s}"]), g["normals"][s])
        neff = via_dev(h, *args) if via_dev else h.update(*args)
        assert np.array_equal(h.poses(), g[f"poses_upd{s}"]), s  # f32 poses bit-exact
        np.testing.assert_allclose(h.log_weights(), g[f"lw{s}"], rtol=0, atol=LW_TOL)
        np.testing.assert_allclose(h.weights(), g[f"w{s}"], rtol=W_RTOL, atol=1e-300)
        assert abs(neff / float(g[f"neff{s}"]) - 1) < 1e-9
        finite = np.isfinite(g[f"wlit{s}"]).all() and g[f"wlit{s}"].sum() > 0
        if finite:  # Java's own (product) weights, where they did not underflow
            np.testing.assert_allclose(h.weights(), g[f"wlit{s}"], rtol=1e-9, atol=1e-300)
        idx, pose, w = h.strongest()
        assert idx == int(g[f"strongest{s}"])
        assert np.array_equal(pose, g[f"poses_upd{s}"][idx])
        np.testing.assert_allclose(h.weighted_pose(), g[f"wpose{s}"], rtol=0, atol=1e-6)
        if f"parents{s}" in g.files:
            h.resample(float(g["uniforms"][s]))
            assert np.array_equal(h.parents(), g[f"parents{s}"]), s
        assert np.array_equal(h.poses(), g[f"poses{s}"]), s
        for m in range(1 if shared else P):
            assert np.array_equal(h.get_map(m, B.MAP_FREE_COUNT), g[f"nfree{s}"][m]), (s, m)
            assert np.array_equal(h.get_map(m, B.MAP_OCC_COUNT), g[f"nocc{s}"][m]), (s, m)
            assert np.array_equal(h.get_map(m, B.MAP_LIKELIHOOD), g[f"lik{s}"][m]), (s, m)
            n = g[f"nfree{s}"][m].astype(np.float64) + g[f"nocc{s}"][m]
            assert np.all(np.abs(h.get_map(m, B.MAP_LOG) - g[f"log{s}"][m]) <= 1e-9 * (n + 1))
    h.close()


def check_motion(lib):
    """Motion model through gms_update on a map with no hits (weights irrelevant): P particles, one
    step per fixture row group.  Poses must be bit-exact f32."""
    g = golden("pyref_motion.npz")
    poses, out = g["poses"], g["out"]
    n = poses.shape[0]
    # each fixture row has its own odometry: run them one particle at a time in groups sharing (dc, dth)
    for i in range(0, n, 1):
        if i >= 48:
            break
        h = lib.create(num_particles=1, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
        h.set_poses(poses[i])
        h.update(np.zeros((0, 2)), np.zeros(0), np.zeros(0, np.uint8), float(g["d_center"][i]),
                 float(g["d_theta"][i]), g["z"][i])
        assert np.array_equal(h.poses()[0], out[i]), i
        h.close()


def check_errors(lib):
    cfg = lib.default_config(num_particles=0)
    try:
        lib.create(cfg)
        raise AssertionError("num_particles=0 accepted")
    except B.GmsError as e:
        assert e.code == B.ERR_INVALID_ARG
    h = small_handle(lib, P=2)
    try:
        h.get_map(5, B.MAP_LOG)
        raise AssertionError("bad particle accepted")
    except B.GmsError as e:
        assert e.code == B.ERR_INVALID_ARG
    try:
        h.get_map(0, 99)
        raise AssertionError("bad kind accepted")
    except B.GmsError as e:
        assert e.code == B.ERR_INVALID_ARG
    try:
        h.resample(1.5)
        raise AssertionError("u01 >= 1 accepted")
    except B.GmsError as e:
        assert e.code == B.ERR_INVALID_ARG
    # a scan longer than GMS_MAX_BEAMS is rejected and leaves the state alone
    n = B.MAX_BEAMS + 1
    poses_before = h.poses()
    try:
        h.update(np.zeros((n, 2)), np.ones(n), np.ones(n, np.uint8), 0.0, 0.0, np.zeros(4))
        raise AssertionError("GMS_MAX_BEAMS + 1 beams accepted")
    except B.GmsError as e:
        assert e.code == B.ERR_INVALID_ARG
    assert np.array_equal(h.poses(), poses_before)
    # empty scan: every weight is the empty product 1 -> uniform weights, Neff = P
    neff = h.update(np.zeros((0, 2)), np.zeros(0), np.zeros(0, np.uint8), 0.0, 0.0, np.zeros(4))
    assert abs(neff - 2.0) < 1e-12
    assert np.array_equal(h.weights(), [0.5, 0.5])
    h.close()


# ---------------------------------------------------------------- CUDA vs oracle replay ----------
def replay_compare(cuda, oracle, P, beams, steps, grid_m, mode, max_range=10.0, resample_every=2,
                   resample_mode=B.RESAMPLE_LITERAL, map_particles=(0,), via_dev=None, seed=7):
    """Same seeded synthetic-room scans and injected draws through both libraries, compared after
    every step.  Bit-exact: poses, per-cell counts, likelihood field, parents, strongest index.
    Toleranced: log-weights 1e-9 abs, weights/Neff 1e-9 rel, weighted pose 2e-6 abs."""
    from gridmap_slam_robot_b200 import synth

    scans = synth.make_scans(steps, beams, max_range=max_range)
    normals, uniforms = synth.make_draws(steps, P, seed=seed)
    kw = dict(num_particles=P, map_width_m=grid_m, map_height_m=grid_m, origin_x=-grid_m / 2, origin_y=-grid_m / 2,
              map_mode=mode, resample_mode=resample_mode)
    g, o = cuda.create(**kw), oracle.create(**kw)
    try:
        for s, sc in enumerate(scans):
            args = (sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
            ng = via_dev(g, *args) if via_dev else g.update(*args)
            no = o.update(*args)
            assert np.array_equal(g.poses(), o.poses()), f"step {s}: poses"
            np.testing.assert_allclose(g.log_weights(), o.log_weights(), rtol=0, atol=LW_TOL, err_msg=f"step {s}")
            np.testing.assert_allclose(g.weights(), o.weights(), rtol=W_RTOL, atol=1e-300, err_msg=f"step {s}")
            assert abs(ng / no - 1) < 1e-9, f"step {s}: neff {ng} vs {no}"
            gi, gp, gw = g.strongest()
            oi, op, ow = o.strongest()
            assert gi == oi and np.array_equal(gp, op) and abs(gw / ow - 1) < 1e-9, f"step {s}: strongest"
            np.testing.assert_allclose(g.weighted_pose(), o.weighted_pose(), rtol=0, atol=2e-6)
            if resample_every and s % resample_every == resample_every - 1:
                g.resample(float(uniforms[s]))
                o.resample(float(uniforms[s]))
                assert np.array_equal(g.parents(), o.parents()), f"step {s}: parents"
                assert np.array_equal(g.poses(), o.poses()), f"step {s}: poses after resample"
                assert np.array_equal(g.weights(), o.weights()[...]) or np.allclose(
                    g.weights(), o.weights(), rtol=W_RTOL, atol=1e-300)
            for p in map_particles:
                for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT, B.MAP_LIKELIHOOD):
                    a, b = g.get_map(p, kind), o.get_map(p, kind)
                    assert np.array_equal(a, b), f"step {s}: map kind {kind} of particle {p}: {np.sum(a != b)} cells"
        return g.launch_count()
    finally:
        g.close()
        o.close()


# ---------------------------------------------------------------- rows adjacent to the path (§8f) ----
def _raw_sweeps(steps, beams):
    """Raw sweeps the way the robot link delivers them: angle, distance, wasHit (+ odometry)."""
    from gridmap_slam_robot_b200 import synth

    scans = synth.make_scans(steps, beams)
    ang = 2.0 * np.pi * np.arange(beams) / beams
    return [(ang, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta) for sc in scans]


def _pyref():
    import importlib.util

    spec = importlib.util.spec_from_file_location("pyref", os.path.join(ROOT, "oracle", "pyref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def trig_probe_angles():
    """Angles that exercise every branch of the fixed-sequence sin / cos: the no-reduction interval, all four
    quadrants, one / two / three reduction rounds (arguments next to multiples of pi/2, incl. the f64 closest to
    one), signed zeros, the edge of the medium range and what lies beyond it."""
    rng = np.random.default_rng(2026)
    k = np.arange(-64, 65, dtype=np.float64)
    near = np.concatenate([np.nextafter(k * (np.pi / 2), np.inf), k * (np.pi / 2), np.nextafter(k * (np.pi / 2), -np.inf),
                           k * (np.pi / 4)])
    return np.concatenate([rng.uniform(-np.pi, np.pi, 4000), rng.uniform(-40.0, 40.0, 4000), rng.uniform(-1.6e6, 1.6e6, 2000),
                           rng.uniform(-1e-6, 1e-6, 200), near,
                           [0.0, -0.0, 5e-324, 1e-300, float.fromhex("0x1.921fb54442d18p-1"), -float.fromhex("0x1.921fb54442d18p-1"),
                            float.fromhex("0x1.921fb54442d19p-1"), 355.0, 52174.0, 1146408.0,
                            1647098.9999, -1647098.9999, 1647099.0, 1.0e7, -3.0e9]])


def check_trig_probe(lib):
    """sin / cos as the library evaluates them, read through gms_deskew (dist = 1, no motion: x = cos(a), y = sin(a)),
    against oracle/pyref.py's restatement of the same operation sequence: bit for bit (beyond the medium range
    every implementation defers to its libm: there a few ulp)."""
    ref = _pyref()
    a = trig_probe_angles()
    h = lib.create(num_particles=1, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
    xy, od = h.deskew(a, np.ones_like(a), 0.0, 0.0)
    h.close()
    c = np.asarray([ref.cos_fixed(v) for v in a])
    s = np.asarray([ref.sin_fixed(v) for v in a])
    med = np.abs(a) < 1647099.0
    assert np.array_equal(xy[med, 0].view(np.uint64), c[med].view(np.uint64)), "cos differs from the fixed sequence"
    assert np.array_equal(xy[med, 1].view(np.uint64), s[med].view(np.uint64)), "sin differs from the fixed sequence"
    np.testing.assert_allclose(xy[~med, 0], c[~med], rtol=0, atol=1e-15)
    np.testing.assert_allclose(xy[~med, 1], s[~med], rtol=0, atol=1e-15)
    assert np.array_equal(od, np.sqrt(xy[:, 0] * xy[:, 0] + xy[:, 1] * xy[:, 1]))


def check_deskew_reference_formula(lib):
    """GridMapApp.java:140-175 restated in Python with oracle/pyref.py's sin / cos (the fixed operation sequence all
    three implementations share): oracle AND CUDA must match it bit for bit."""
    ref = _pyref()
    h = lib.create(num_particles=1, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
    rng = np.random.default_rng(11)
    n = 97
    ang, dist = rng.uniform(-np.pi, np.pi, n), rng.uniform(0.05, 11.0, n)
    dc, dth = 0.07, -0.21
    d_i = -(n - np.arange(n)) / float(n)
    a = ang + dth * d_i
    x = dist * np.asarray([ref.cos_fixed(v) for v in a]) + dc * d_i
    y = dist * np.asarray([ref.sin_fixed(v) for v in a])
    xy, od = h.deskew(ang, dist, dc, dth)
    assert np.array_equal(xy[:, 0], x) and np.array_equal(xy[:, 1], y)
    assert np.array_equal(od, np.sqrt(x * x + y * y))
    h.close()


def check_next_rows(cuda, oracle):
    """CUDA vs oracle: fused de-skew + update on raw sweeps, renderer hand-off, combined-map fusion."""
    P, beams, steps = 12, 180, 5
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    g, o = cuda.create(**kw), oracle.create(**kw)
    from gridmap_slam_robot_b200 import synth

    normals, uniforms = synth.make_draws(steps, P)
    for s, (ang, dist, hit, dc, dth) in enumerate(_raw_sweeps(steps, beams)):
        ng = g.update_raw(ang, dist, hit, dc, dth, normals[s])
        no = o.update_raw(ang, dist, hit, dc, dth, normals[s])
        assert abs(ng / no - 1) < 1e-9
        assert np.array_equal(g.poses(), o.poses())
        np.testing.assert_allclose(g.log_weights(), o.log_weights(), rtol=0, atol=1e-9)
        if s % 2:
            g.resample(float(uniforms[s]))
            o.resample(float(uniforms[s]))
            assert np.array_equal(g.parents(), o.parents())
    # the de-skewed beams are bit-identical (both libraries evaluate sin / cos as the same operation sequence), so
    # the integer counters are too
    for p in range(P):
        for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT):
            a, b = g.get_map(p, kind), o.get_map(p, kind)
            assert np.array_equal(a, b), (p, kind, np.sum(a != b))
    # renderer hand-off: packed ABGR gray words
    for lik in (False, True):
        a, b = g.render_map(0, lik), o.render_map(0, lik)
        assert a.dtype == np.uint32 and np.all((a >> 24) == 0xFE)  # alpha 255 & 0xfe mask
        assert np.mean(a != b) < 1e-3
        lvl = (a & 0xFF).astype(int) - (b & 0xFF).astype(int)
        assert np.abs(lvl).max() <= 1 or np.mean(a != b) < 1e-4
    # combined-map fusion
    lg, lk = g.combined_map()
    lo, lko = o.combined_map()
    finite = np.isfinite(lo)
    assert np.array_equal(np.isfinite(lg), finite)
    np.testing.assert_allclose(lg[finite], lo[finite], rtol=1e-9, atol=1e-9)
    assert np.mean(lk != lko) < 1e-3
    g.close()
    o.close()


def check_truncation_toward_zero(lib):
    """GridMap.java:273-274: gridX = (int) ((x - position.x) / resolution) truncates TOWARD ZERO, so an end point
    up to one cell left of / below the map still reads column / row 0 (it is not skipped).  Both scoring
    kernels (one warp per particle; heading-sorted fixed-point path) must reproduce that."""
    for P in (1, 5000):  # 5000 particles on a shared map take the k_score_sorted path inside gms_update
        h = lib.create(num_particles=P, map_width_m=3.0, map_height_m=2.5, resolution=0.05, origin_x=-1.5,
                       origin_y=-1.25, map_mode=B.MAP_SHARED)
        nf, no = np.zeros((h.H, h.W), np.uint32), np.zeros((h.H, h.W), np.uint32)
        no[:, 0] = 1  # column 0 occupied -> likelihood > 0.5 there
        no[0, :] = 1
        h.set_map_counts(0, nf, no)
        # robot at the origin, heading 0: beam end points at world (-1.5 - 0.02, 0.3) and (0.4, -1.25 - 0.03):
        # (x - pos)/res = -0.4 and -0.6 -> (int) 0 in Java; a floor() would give -1 and skip the beam
        xy = np.asarray([[-1.52, 0.3], [0.4, -1.28], [-1.5 - 0.051, 0.3]], np.float64)
        dist = np.hypot(xy[:, 0], xy[:, 1])
        hit = np.ones(3, np.uint8)
        h.map_compute_likelihood(0)
        lik = h.get_map(0, B.MAP_LIKELIHOOD)
        expect = 0.0
        for gx, gy in ((0, int((0.3 + 1.25) / 0.05)), (int((0.4 + 1.5) / 0.05), 0)):
            v = lik[gy, gx]
            assert v != 0.5
            expect += np.log(0.9 * v + (1 - 0.9) * 1.0 / 10.0)
        # third beam: (x - pos)/res = -1.02 -> (int) -1 -> out of the map -> skipped
        lp, _ = h.map_probability_of(0, (0.0, 0.0, 0.0), xy, hit)
        assert abs(lp - expect) < 1e-12, (lp, expect)
        h.set_poses(np.zeros((P, 3), np.float32))
        # zero motion: sd*z + mean with z = 0 and dCenter = dTheta = 0 leaves every pose at the origin
        h.update(xy, dist, hit, 0.0, 0.0, np.zeros((P, 2)))
        np.testing.assert_allclose(h.log_weights(), expect, rtol=0, atol=1e-12)
        h.close()


def check_underflow(lib):
    """Java's product weights underflow (NaN after normalisation); the build's log-domain weights do not."""
    g = golden("pyref_underflow.npz")
    assert np.isnan(g["wlit"]).all()  # what the reference itself would produce (SLAM.java:121)
    h = small_handle(lib, P=3, mode=B.MAP_SHARED)
    neff = h.update(g["xy"], g["dist"], g["hit"], 0.0, 0.0, g["normals"])
    np.testing.assert_allclose(h.log_weights(), g["lw"], rtol=0, atol=LW_TOL)
    assert np.all(g["lw"] < -700)  # below ln(DBL_MIN): exp() of it is 0
    np.testing.assert_allclose(h.weights(), g["w"], rtol=W_RTOL, atol=0)
    assert abs(neff - float(g["neff"])) < 1e-9 and abs(neff - 3.0) < 1e-9
    h.close()


# ---------------------------------------------------------------- full BASELINE sizes, sampled --------
def check_full_size_sampled(lib, oracle, P, beams, grid_m, max_range=10.0, steps=3, sample=48, seed=11):
    """Shared-map step at sizes where a whole oracle replay would take minutes, checked piecewise against the
    oracle's per-map operators (which finish in seconds):
      motion      a random sample of particles against Odometry.apply on the previous pose, bit for bit
      field       the ONE shared likelihood field, bit for bit (GridMap.computeLikelihoodMap on the oracle's map)
      scoring     the sample's log-weights against GridMap.probabilityOf on that field, |d| <= 1e-9
      normalise   weights / Neff / strongest recomputed in numpy from the library's own log-weights
      integration per-cell counts bit for bit: the oracle integrates the same scan from the library's strongest pose
      resampling  parents against the oracle's resampler fed the library's weights; children's poses and weights
    """
    import ctypes as C

    from gridmap_slam_robot_b200 import synth

    scans = synth.make_scans(steps, beams, max_range=max_range)
    normals, uniforms = synth.make_draws(steps, P, seed=seed)
    kw = dict(map_width_m=grid_m, map_height_m=grid_m, origin_x=-grid_m / 2, origin_y=-grid_m / 2,
              map_mode=B.MAP_SHARED)
    g = lib.create(num_particles=P, **kw)
    o = oracle.create(num_particles=1, **kw)
    r = oracle.create(num_particles=P, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
    assert g.info.resample_mode == r.info.resample_mode
    motion = oracle.dll.gmsref_motion
    motion.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_double, C.c_double, C.c_double, C.c_double]
    rng = np.random.default_rng(seed)
    try:
        prev = g.poses()
        for s, sc in enumerate(scans):
            neff = g.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
            poses, lw, w = g.poses(), g.log_weights(), g.weights()
            pick = np.unique(np.concatenate([rng.choice(P, size=min(sample, P), replace=False),
                                             [0, P - 1, int(np.argmax(lw)), int(np.argmin(lw))]]))
            for i in pick:
                p = (C.c_float * 3)(*prev[i])
                assert motion(o.h, p, sc.d_center, sc.d_theta, float(normals[s][i, 0]), float(normals[s][i, 1])) == 0
                assert np.array_equal(np.asarray(p[:], np.float32), poses[i]), f"step {s}: motion of particle {i}"
            o.map_compute_likelihood(0)
            assert np.array_equal(g.get_map(0, B.MAP_LIKELIHOOD), o.get_map(0, B.MAP_LIKELIHOOD)), f"step {s}: field"
            for i in pick:
                lp, _ = o.map_probability_of(0, poses[i], sc.beam_xy, sc.beam_hit)
                assert abs(lp - lw[i]) <= LW_TOL, f"step {s}: log-weight of particle {i}: {lw[i]} vs {lp}"
            e = np.exp(lw - lw.max())
            wn = e / e.sum()
            np.testing.assert_allclose(w, wn, rtol=1e-9, atol=1e-300, err_msg=f"step {s}: weights")
            assert abs(neff * np.sum(wn * wn) - 1) < 1e-9, f"step {s}: neff"
            si, sp, sw = g.strongest()
            assert si == int(np.argmax(lw)) and np.array_equal(sp, poses[si]) and abs(sw / wn[si] - 1) < 1e-9
            wp = g.weighted_pose()
            assert abs(wp[0] - np.dot(wn, poses[:, 0].astype(np.float64))) < 2e-6
            assert abs(wp[1] - np.dot(wn, poses[:, 1].astype(np.float64))) < 2e-6
            o.map_integrate_observation(0, sp, sc.beam_xy, sc.beam_dist, sc.beam_hit)
            for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT):
                a, b = g.get_map(0, kind), o.get_map(0, kind)
                assert np.array_equal(a, b), f"step {s}: counts kind {kind}: {np.sum(a != b)} cells differ"
            u = float(uniforms[s])
            r.set_weights(w)
            r.resample(u)
            g.resample(u)
            par = g.parents()
            assert np.array_equal(par, r.parents()), f"step {s}: parents"
            assert np.array_equal(g.poses(), poses[par]) and np.array_equal(g.weights(), w[par])
            prev = g.poses()
    finally:
        g.close()
        o.close()
        r.close()


def check_full_size_sampled_pp(lib, oracle, P, beams, grid_m, max_range=10.0, steps=3, sample=6, seed=13):
    """Per-particle maps (the reference's mode) at full size, a sample of particles per step against the
    oracle's per-map operators: each sampled particle's field, log-weight and integrated counts are reproduced
    from its own counts before the step; after resampling every sampled child's map equals its parent's."""
    from gridmap_slam_robot_b200 import synth

    scans = synth.make_scans(steps, beams, max_range=max_range)
    normals, uniforms = synth.make_draws(steps, P, seed=seed)
    kw = dict(map_width_m=grid_m, map_height_m=grid_m, origin_x=-grid_m / 2, origin_y=-grid_m / 2)
    g = lib.create(num_particles=P, map_mode=B.MAP_PER_PARTICLE, **kw)
    o = oracle.create(num_particles=1, map_mode=B.MAP_SHARED, **kw)
    r = oracle.create(num_particles=P, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
    assert g.info.resample_mode == r.info.resample_mode
    rng = np.random.default_rng(seed)
    counts = lambda h, i: (h.get_map(i, B.MAP_FREE_COUNT), h.get_map(i, B.MAP_OCC_COUNT))  # noqa: E731
    try:
        for s, sc in enumerate(scans):
            pick = np.unique(np.concatenate([rng.choice(P, size=min(sample, P), replace=False), [0, P - 1]]))
            before = {int(i): counts(g, int(i)) for i in pick}
            neff = g.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
            poses, lw, w = g.poses(), g.log_weights(), g.weights()
            for i in before:
                o.set_map_counts(0, *before[i])
                o.map_compute_likelihood(0)
                assert np.array_equal(g.get_map(i, B.MAP_LIKELIHOOD), o.get_map(0, B.MAP_LIKELIHOOD)), (s, i, "field")
                lp, _ = o.map_probability_of(0, poses[i], sc.beam_xy, sc.beam_hit)
                assert abs(lp - lw[i]) <= LW_TOL, (s, i, lw[i], lp)
                o.map_integrate_observation(0, poses[i], sc.beam_xy, sc.beam_dist, sc.beam_hit)
                for a, b in zip(counts(g, i), counts(o, 0)):
                    assert np.array_equal(a, b), (s, i, "counts", int(np.sum(a != b)))
            e = np.exp(lw - lw.max())
            wn = e / e.sum()
            np.testing.assert_allclose(w, wn, rtol=1e-9, atol=1e-300)
            assert abs(neff * np.sum(wn * wn) - 1) < 1e-9
            assert g.strongest()[0] == int(np.argmax(lw))
            u = float(uniforms[s])
            r.set_weights(w)
            r.resample(u)
            rpar = r.parents()
            kids = np.unique(np.concatenate([rng.choice(P, size=min(sample, P), replace=False), [0, P - 1]]))
            parent_maps = {int(m): counts(g, int(rpar[m])) + (g.get_map(int(rpar[m]), B.MAP_LIKELIHOOD),) for m in kids}
            g.resample(u)
            assert np.array_equal(g.parents(), rpar), f"step {s}: parents"
            assert np.array_equal(g.poses(), poses[rpar])
            for m, (pf, po, pl) in parent_maps.items():  # deep copy of both arrays: GridMap.java:106-124
                cf, co = counts(g, m)
                assert np.array_equal(cf, pf) and np.array_equal(co, po), (s, m, "child counts")
                assert np.array_equal(g.get_map(m, B.MAP_LIKELIHOOD), pl), (s, m, "child field")
    finally:
        g.close()
        o.close()
        r.close()
