"""GPU (-m gpu): the CUDA hot path, called through the C-ABI, against
  (1) the hand-derived Appendix B vectors, (2) the committed pyref golden fixtures,
  (3) the C oracle on seeded synthetic-room replays, (4) size-independent properties at full size.
"""
import ctypes

import numpy as np
import pytest

import checks
from gridmap_slam_robot_b200 import binding as B

pytestmark = pytest.mark.gpu


def test_no_cpu_fallback_symbols(cuda):
    h = cuda.create(num_particles=4, map_mode=B.MAP_SHARED)
    assert h.info.is_cuda == 1
    h.close()


def test_constants(cuda):
    checks.check_constants(cuda)


def test_appendix_b_rays(cuda):
    checks.check_appendix_b_rays(cuda)


def test_appendix_b_resample(cuda):
    checks.check_appendix_b_resample(cuda)


def test_blank_likelihood(cuda):
    checks.check_blank_likelihood(cuda)


def test_golden_rays(cuda):
    checks.check_rays(cuda)


def test_golden_apply(cuda):
    checks.check_apply(cuda)


def test_golden_blur(cuda):
    checks.check_blur(cuda)


def test_golden_resample(cuda):
    checks.check_resample(cuda)


def test_golden_motion(cuda):
    checks.check_motion(cuda)


@pytest.mark.parametrize("shared", [False, True])
def test_golden_slam(cuda, shared):
    checks.check_slam(cuda, shared)


def test_errors(cuda):
    checks.check_errors(cuda)


def _via_dev(policy=B.POLICY_NEVER):
    """Route an update through the device-resident entry point with torch-owned device buffers."""
    import torch

    def run(h, xy, dist, hit, dc, dth, normals):
        dev = torch.device("cuda:0")
        t_xy = torch.from_numpy(np.ascontiguousarray(xy, np.float64)).to(dev)
        t_d = torch.from_numpy(np.ascontiguousarray(dist, np.float64)).to(dev)
        t_h = torch.from_numpy(np.ascontiguousarray(hit, np.uint8)).to(dev)
        t_n = torch.from_numpy(np.ascontiguousarray(normals, np.float64)).to(dev)
        torch.cuda.synchronize()
        Bn = int(t_d.numel())
        h.step_dev(t_xy.data_ptr() if Bn else 0, t_d.data_ptr() if Bn else 0, t_h.data_ptr() if Bn else 0, Bn, dc,
                   dth, t_n.data_ptr(), policy)
        return h.read_neff()

    return run


@pytest.mark.parametrize("shared", [False, True])
def test_golden_slam_device_entry(cuda, shared):
    checks.check_slam(cuda, shared, via_dev=_via_dev())


# ---- CUDA vs oracle on the synthetic room ----
def test_replay_k1_per_particle(cuda, oracle):
    """K1 geometry (20 m room, 400^2 cells, 360 beams, 10 m range => misses exercised), fewer particles."""
    n = checks.replay_compare(cuda, oracle, P=24, beams=360, steps=8, grid_m=20.0, mode=B.MAP_PER_PARTICLE,
                              map_particles=(0, 11, 23))
    assert n > 0


def test_replay_k1_shared(cuda, oracle):
    checks.replay_compare(cuda, oracle, P=100, beams=360, steps=10, grid_m=20.0, mode=B.MAP_SHARED)


def test_replay_small_map_out_of_bounds(cuda, oracle):
    """12 m map around an 18 m room: rays leave the grid, end points fall outside (skipped in scoring)."""
    checks.replay_compare(cuda, oracle, P=16, beams=180, steps=6, grid_m=12.0, mode=B.MAP_PER_PARTICLE,
                          map_particles=(0, 15))


def test_replay_fixed_point_resampling(cuda, oracle):
    checks.replay_compare(cuda, oracle, P=300, beams=90, steps=6, grid_m=20.0, mode=B.MAP_SHARED,
                          resample_mode=B.RESAMPLE_FIXED, resample_every=1)


def test_replay_720_beams_1024_grid_shared(cuda, oracle):
    """K2/K3-style: 720 beams underflow Java's product (0.1^720): log-domain weights stay finite."""
    checks.replay_compare(cuda, oracle, P=64, beams=720, steps=4, grid_m=51.2, mode=B.MAP_SHARED, max_range=12.0)


def test_replay_device_entry(cuda, oracle):
    checks.replay_compare(cuda, oracle, P=32, beams=360, steps=5, grid_m=20.0, mode=B.MAP_PER_PARTICLE,
                          via_dev=_via_dev(), map_particles=(0, 31))


def test_determinism_and_profile(cuda):
    """Same inputs twice => byte-identical state (integer atomics commute; fixed reduction trees)."""
    from gridmap_slam_robot_b200 import synth

    P, steps = 256, 4
    scans = synth.make_scans(steps, 360)
    normals, uniforms = synth.make_draws(steps, P)
    outs = []
    for _ in range(2):
        h = cuda.create(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
        h.profile_enable(True)
        for s, sc in enumerate(scans):
            h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
            h.resample(float(uniforms[s]))
        ms, launches = h.profile_read()
        # per-particle maps: no likelihood-field launches in the step (the field is evaluated where the scan reads it)
        assert launches["score"] >= steps and launches["map_update"] >= steps and ms["score"] > 0
        outs.append((h.poses().tobytes(), h.weights().tobytes(), h.get_map(5, B.MAP_FREE_COUNT).tobytes(),
                     h.get_map(5, B.MAP_LIKELIHOOD).tobytes(), h.parents().tobytes()))
        h.close()
    assert outs[0] == outs[1]


def test_device_philox_matches_oracle(cuda, oracle):
    """normals == NULL: both sides draw from Philox4x32-10 keyed by (seed, global index, step).  The
    Box-Muller transcendental functions differ by ulps between libm and CUDA, so poses agree to 1e-6."""
    from gridmap_slam_robot_b200 import synth

    sc = synth.make_scans(1, 90)[0]
    kw = dict(num_particles=512, map_mode=B.MAP_SHARED, seed=1234)
    g, o = cuda.create(**kw), oracle.create(**kw)
    for _ in range(3):
        g.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, 0.05, 0.01)
        o.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, 0.05, 0.01)
    np.testing.assert_allclose(g.poses(), o.poses(), rtol=0, atol=1e-6)
    p = g.poses()
    assert 0.1 < p[:, 0].mean() < 0.2 and p[:, 2].std() > 0.05
    g.resample(-1.0)
    o.resample(-1.0)
    assert np.array_equal(g.parents(), o.parents())
    g.close()
    o.close()


# ---- size-independent properties at BASELINE sizes ----
def test_full_size_k2_properties(cuda):
    """K2: 1k particles x 360 beams, 1024^2 shared grid.  Properties: weights sum to 1, Neff in [1, P],
    parents sorted and in range, counts only grow, every likelihood value in [0, 1], and the map
    touched only inside the scan's reach."""
    from gridmap_slam_robot_b200 import synth

    P = 1000
    h = cuda.create(num_particles=P, map_width_m=51.2, map_height_m=51.2, origin_x=-25.6, origin_y=-25.6,
                    map_mode=B.MAP_SHARED)
    assert (h.W, h.H) == (1024, 1024)
    scans = synth.make_scans(6, 360)
    prev = np.zeros((1024, 1024), np.uint64)
    for s, sc in enumerate(scans):
        neff = h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta)
        w = h.weights()
        assert abs(w.sum() - 1) < 1e-9 and 1.0 <= neff <= P + 1e-6
        h.resample(-1.0)
        par = h.parents()
        assert par.min() >= 0 and par.max() < P and np.all(np.diff(par) >= 0)
        tot = h.get_map(0, B.MAP_FREE_COUNT).astype(np.uint64) + h.get_map(0, B.MAP_OCC_COUNT)
        assert np.all(tot >= prev)
        prev = tot
    lik = h.get_map(0, B.MAP_LIKELIHOOD)
    assert lik.min() >= 0.0 and lik.max() <= 1.0 + 1e-12
    ys, xs = np.nonzero(prev)
    # room is +-9 m, robot within 6.1 m of the origin, 10 m range (+2 cells): |coord| < 16.2 m
    assert np.all(np.abs((xs + 0.5) * 0.05 - 25.6) < 16.2) and np.all(np.abs((ys + 0.5) * 0.05 - 25.6) < 16.2)
    h.close()


# ---- paths that only large particle sets / unusual geometry reach ----
def test_replay_sorted_scoring_path(cuda, oracle):
    """>= 4096 particles on a shared map: heading-sorted k_score_sorted<G> with the fixed-point cell index,
    717 beams (not a multiple of 8: remainder loop) and a 12 m map (end points outside: sentinel + exact path)."""
    checks.replay_compare(cuda, oracle, P=6000, beams=717, steps=4, grid_m=12.0, mode=B.MAP_SHARED, max_range=30.0,
                          resample_mode=B.RESAMPLE_FIXED, resample_every=1)
    checks.replay_compare(cuda, oracle, P=5000, beams=360, steps=3, grid_m=51.2, mode=B.MAP_SHARED, max_range=30.0,
                          resample_mode=B.RESAMPLE_FIXED, resample_every=2)


def test_replay_other_resolution_generic_blur(cuda, oracle):
    """0.02 m cells: sigma = sqrt(0.05/0.02) -> 11 taps: the generic (run-time width) likelihood kernel; odd,
    non-square grid (scalar copy path)."""
    from gridmap_slam_robot_b200 import synth

    P, steps = 10, 4
    kw = dict(num_particles=P, map_width_m=7.0 + 0.02, map_height_m=5.0 + 3 * 0.02, resolution=0.02, origin_x=-3.5,
              origin_y=-2.5, resample_mode=B.RESAMPLE_LITERAL)
    g, o = cuda.create(**kw), oracle.create(**kw)
    assert g.info.kernel_taps == o.info.kernel_taps == 11 and (g.W * g.H) % 2 == 1
    scans = synth.make_scans(steps, 120, max_range=3.0)
    normals, uniforms = synth.make_draws(steps, P)
    for s, sc in enumerate(scans):
        a = (sc.beam_xy * 0.3, sc.beam_dist * 0.3, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        ng, no = g.update(*a), o.update(*a)
        assert abs(ng / no - 1) < 1e-9 and np.array_equal(g.poses(), o.poses())
        np.testing.assert_allclose(g.log_weights(), o.log_weights(), rtol=0, atol=1e-9)
        g.resample(float(uniforms[s]))
        o.resample(float(uniforms[s]))
        assert np.array_equal(g.parents(), o.parents())
        for p in (0, P - 1):
            for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT, B.MAP_LIKELIHOOD):
                assert np.array_equal(g.get_map(p, kind), o.get_map(p, kind)), (s, p, kind)
    g.close()
    o.close()


def test_degenerate_scans(cuda, oracle):
    """All-miss scan (no scoring factors: uniform weights), zero-length and NaN/huge beams, particles far
    outside the map: both sides agree, nothing crashes."""
    P = 64
    kw = dict(num_particles=P, map_width_m=6.0, map_height_m=6.0, origin_x=-3.0, origin_y=-3.0, map_mode=B.MAP_SHARED)
    g, o = cuda.create(**kw), oracle.create(**kw)
    rng = np.random.default_rng(2)
    z = rng.standard_normal((P, 2))
    n = 40
    ang = 2 * np.pi * np.arange(n) / n
    dist = np.full(n, 10.0)
    xy = np.stack([dist * np.cos(ang), dist * np.sin(ang)], 1)
    for hit in (np.zeros(n, np.uint8), np.ones(n, np.uint8)):
        ng, no = g.update(xy, dist, hit, 0.0, 0.0, z), o.update(xy, dist, hit, 0.0, 0.0, z)
        assert abs(ng / no - 1) < 1e-9
        np.testing.assert_allclose(g.log_weights(), o.log_weights(), rtol=0, atol=1e-9)
    xy2, d2 = xy.copy(), dist.copy()
    xy2[0] = (0.0, 0.0); d2[0] = 0.0              # zero-length ray: start cell three times
    xy2[1] = (1e30, -1e30); d2[1] = 1e30          # saturating (int) casts
    xy2[2] = (np.nan, 1.0); d2[2] = np.nan        # NaN -> (int) 0
    hit = np.ones(n, np.uint8)
    ng, no = g.update(xy2, d2, hit, 0.05, 0.0, z), o.update(xy2, d2, hit, 0.05, 0.0, z)
    assert np.array_equal(np.isnan(g.log_weights()), np.isnan(o.log_weights()))
    fin = np.isfinite(o.log_weights())
    np.testing.assert_allclose(g.log_weights()[fin], o.log_weights()[fin], rtol=0, atol=1e-9)
    for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT):
        assert np.array_equal(g.get_map(0, kind), o.get_map(0, kind))
    far = np.tile(np.asarray([[500.0, -500.0, 0.3]], np.float32), (P, 1))
    g.set_poses(far); o.set_poses(far)
    ng, no = g.update(xy, dist, hit, 0.0, 0.0, z), o.update(xy, dist, hit, 0.0, 0.0, z)
    assert abs(ng / no - 1) < 1e-9 and np.array_equal(g.poses(), o.poses())
    g.close()
    o.close()


def test_literal_vs_fixed_resampling_agree(cuda):
    """The associative fixed-point CDF picks the same parents as Java's sequential f64 sum on random weights
    (a disagreement needs U within ~1e-16 of a CDF step); the count is reported, and must be 0 here."""
    rng = np.random.default_rng(9)
    bad = 0
    for n in (500, 2048, 20000):
        w = rng.random(n) ** 6
        w /= w.sum()
        par = []
        for mode in (B.RESAMPLE_LITERAL, B.RESAMPLE_FIXED):
            h = cuda.create(num_particles=n, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED, resample_mode=mode)
            h.set_weights(w)
            h.resample(0.123456789)
            par.append(h.parents())
            h.close()
        bad += int(np.sum(par[0] != par[1]))
    assert bad == 0


def test_device_side_resample_policy(cuda, oracle):
    """GMS_RESAMPLE_IF_NEFF_LOW decides on the device (no host round trip) what GridMapApp.java:185 decides
    on the host: resample iff neff < P/2.  The device-policy run must equal a run where the host applies
    the same rule with gms_update + gms_resample, and the oracle doing the same."""
    import torch

    from gridmap_slam_robot_b200 import synth

    P, steps, beams = 512, 10, 180
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0,
              map_mode=B.MAP_SHARED, resample_mode=B.RESAMPLE_LITERAL)
    scans = synth.make_scans(steps, beams)
    normals, uniforms = synth.make_draws(steps, P)
    dev = torch.device("cuda:0")
    gd, gh, o = cuda.create(**kw), cuda.create(**kw), oracle.create(**kw)
    resampled = 0
    for s, sc in enumerate(scans):
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (sc.beam_xy, sc.beam_dist, sc.beam_hit, normals[s])]
        torch.cuda.synchronize()
        gd.step_dev(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), beams, sc.d_center, sc.d_theta, t[3].data_ptr(),
                    B.POLICY_IF_NEFF_LOW, float(uniforms[s]))
        neff_d = gd.read_neff()
        for h in (gh, o):
            neff = h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
            if neff < P // 2:
                h.resample(float(uniforms[s]))
                resampled += h is gh
        assert abs(neff_d / neff - 1) < 1e-9
        assert np.array_equal(gd.poses(), gh.poses()) and np.array_equal(gd.poses(), o.poses()), s
        np.testing.assert_allclose(gd.weights(), o.weights(), rtol=1e-9, atol=1e-300)
        assert np.array_equal(gd.get_map(0, B.MAP_FREE_COUNT), o.get_map(0, B.MAP_FREE_COUNT))
    assert 0 < resampled < steps  # the schedule exercised both branches
    for h in (gd, gh, o):
        h.close()


def test_strongest_survives_resampling_like_the_reference(cuda, oracle):
    """GridMapApp keeps `strongestParticle` across the resample (GridMapApp.java:188-190): its pose and weight are
    those of the last update even though the particle list was replaced, and its map (GridMapApp.java:376-393
    draws strongestParticle.m) is the one its FIRST CHILD inherits — so the reported index follows it there."""
    from gridmap_slam_robot_b200 import synth

    P = 200
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    sc = synth.make_scans(2, 90)
    z = synth.make_draws(2, P)[0]
    seen = []
    for lib in (cuda, oracle):
        h = lib.create(**kw)
        h.update(sc[0].beam_xy, sc[0].beam_dist, sc[0].beam_hit, 0.05, 0.01, z[0])
        h.update(sc[1].beam_xy, sc[1].beam_dist, sc[1].beam_hit, 0.05, 0.01, z[1])
        idx, pose, w = h.strongest()
        assert np.array_equal(pose, h.poses()[idx]) and w == h.weights()[idx] == h.weights().max()
        occ_before = h.get_map(idx, B.MAP_OCC_COUNT).copy()
        h.resample(0.77)
        idx2, pose2, w2 = h.strongest()
        parents = h.parents()
        assert idx2 == int(np.flatnonzero(parents == idx)[0])  # the strongest particle always survives
        assert np.array_equal(pose2, pose) and w2 == w
        assert np.array_equal(h.poses()[idx2], pose)
        assert np.array_equal(h.get_map(idx2, B.MAP_OCC_COUNT), occ_before)
        seen.append((idx, idx2))
        h.close()
    assert seen[0] == seen[1]


def _halfway_hook(first, poses, xy, dist, hit, dc, dth):
    """A deterministic stand-in for GridMap.findBestPoseOptim: pulls every pose 25 % towards the origin heading."""
    poses[:, 2] = (poses[:, 2] * np.float32(0.75)).astype(np.float32)
    poses[:, 0] = (poses[:, 0] + np.float32(0.01)).astype(np.float32)


@pytest.mark.parametrize("mode", [B.MAP_PER_PARTICLE, B.MAP_SHARED])
def test_pose_optimizer_hook_matches_oracle(cuda, oracle, mode):
    """A4 (GridMap.findBestPoseOptim, SLAM.java:97) as a CPU hook between motion and scoring: same hook on both
    libraries => same poses (bit-exact), weights and maps; removing the hook restores the identity default."""
    from gridmap_slam_robot_b200 import synth

    P, steps = 48, 4
    scans = synth.make_scans(steps, 180)
    normals, uniforms = synth.make_draws(steps, P)
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0, map_mode=mode,
              resample_mode=B.RESAMPLE_LITERAL)
    g, o = cuda.create(**kw), oracle.create(**kw)
    calls = []
    for h in (g, o):
        h.set_pose_optimizer(lambda first, poses, *a, _h=h: (calls.append((first, len(poses))), _halfway_hook(first, poses, *a)))
    for s, sc in enumerate(scans):
        if s == steps - 1:  # last step without the hook
            g.set_pose_optimizer(None)
            o.set_pose_optimizer(None)
        ng = g.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        no = o.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        assert np.array_equal(g.poses(), o.poses()), f"step {s}"
        np.testing.assert_allclose(g.log_weights(), o.log_weights(), rtol=0, atol=checks.LW_TOL)
        assert abs(ng / no - 1) < 1e-9
        g.resample(float(uniforms[s]))
        o.resample(float(uniforms[s]))
        assert np.array_equal(g.parents(), o.parents())
    for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT, B.MAP_LIKELIHOOD):
        assert np.array_equal(g.get_map(0, kind), o.get_map(0, kind))
    assert len(calls) == 2 * (steps - 1) and all(c == (0, P) for c in calls)
    g.close()
    o.close()


def test_pose_optimizer_hook_may_call_the_objective(cuda, oracle):
    """The hook may evaluate GridMap.probabilityOf (the reference's objective, GridMap.java:355-359) on the handle:
    a two-candidate search per particle gives the same choice on both libraries."""
    from gridmap_slam_robot_b200 import synth

    P = 12
    scans = synth.make_scans(3, 90)
    normals, _ = synth.make_draws(3, P)
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    res = []
    for lib in (cuda, oracle):
        h = lib.create(**kw)

        def hook(first, poses, xy, dist, hit, dc, dth, _h=h):
            for k in range(len(poses)):
                cand = poses[k].copy()
                cand[2] += np.float32(0.02)
                a = _h.map_probability_of(first + k, poses[k], xy, hit)[0]
                b = _h.map_probability_of(first + k, cand, xy, hit)[0]
                if b > a + 1e-6:
                    poses[k] = cand

        h.set_pose_optimizer(hook)
        for s, sc in enumerate(scans):
            h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        res.append((h.poses().copy(), h.log_weights().copy(), h.get_map(3, B.MAP_OCC_COUNT).copy()))
        h.close()
    assert np.array_equal(res[0][0], res[1][0])
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=0, atol=checks.LW_TOL)
    assert np.array_equal(res[0][2], res[1][2])


def test_scoring_many_beams_does_not_underflow(cuda, oracle):
    """Per-particle scoring near GMS_MAX_BEAMS on an explored map: a lane multiplies ~390 factors of 0.01..0.91;
    the running product must not underflow (the oracle sums ln per beam)."""
    rng = np.random.default_rng(11)
    Bn = 12500
    ang = rng.uniform(-np.pi, np.pi, Bn)
    dist = rng.uniform(1.0, 8.0, Bn)
    xy = np.stack([dist * np.cos(ang), dist * np.sin(ang)], 1)
    hit = np.ones(Bn, np.uint8)
    kw = dict(num_particles=2, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    out = []
    for lib in (cuda, oracle):
        h = lib.create(**kw)
        pose = np.array([0.1, -0.2, 0.3], np.float32)
        sub = slice(0, 2000)
        h.map_integrate_observation(0, pose, xy[sub], dist[sub], hit[sub])
        h.map_compute_likelihood(0)
        out.append(h.map_probability_of(0, pose, xy, hit))
        h.close()
    assert np.isfinite(out[0][0]) and out[0][0] < -2000.0  # far below ln(DBL_MIN) = -708
    assert abs(out[0][0] - out[1][0]) <= 1e-9 * abs(out[1][0])


def test_sorted_scoring_validation_variant_many_beams(cuda, oracle, monkeypatch):
    """GMS_SCORE_V=1 (ALU-lean index validation) with more than 3072 hit beams needs > 48 KB of dynamic shared
    memory in every instantiation; results equal the default variant's and the oracle's."""
    from gridmap_slam_robot_b200 import synth

    P, Bn = 4200, 3600
    scans = synth.make_scans(2, Bn, max_range=30.0)
    normals, _ = synth.make_draws(2, P)
    kw = dict(num_particles=P, map_width_m=51.2, map_height_m=51.2, origin_x=-25.6, origin_y=-25.6, map_mode=B.MAP_SHARED)
    lws = []
    for v in ("0", "1", "2", "5", "6"):
        monkeypatch.setenv("GMS_SCORE_V", v)
        h = cuda.create(**kw)
        for s, sc in enumerate(scans):
            h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        lws.append(h.log_weights().copy())
        h.close()
    assert np.array_equal(lws[0], lws[1]) and np.array_equal(lws[0], lws[2]) and np.array_equal(lws[0], lws[3])
    np.testing.assert_allclose(lws[4], lws[0], rtol=0, atol=1e-10)  # 4-beam batches: another product tree
    o = oracle.create(**kw)
    for s, sc in enumerate(scans):
        o.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
    np.testing.assert_allclose(lws[0], o.log_weights(), rtol=0, atol=checks.LW_TOL)
    o.close()


def test_determinism_large_shared_sorted_path(cuda):
    """>= 8k particles on a shared map, device Philox, resampling every step: the heading-sorted scoring, the
    cooperative normalise / resample kernels and the side streams give byte-identical results run to run."""
    import torch

    from gridmap_slam_robot_b200 import synth

    P, steps = 16384, 6
    scans = synth.make_scans(steps, 360, max_range=30.0)
    dev = torch.device("cuda:0")
    outs = []
    for _ in range(3):
        h = cuda.create(num_particles=P, map_width_m=51.2, map_height_m=51.2, origin_x=-25.6, origin_y=-25.6,
                        map_mode=B.MAP_SHARED, seed=99)
        neffs = []
        for s, sc in enumerate(scans):
            t_xy, t_d, t_h = (torch.from_numpy(a).to(dev) for a in (sc.beam_xy, sc.beam_dist, sc.beam_hit))
            torch.cuda.synchronize()
            h.step_dev(t_xy.data_ptr(), t_d.data_ptr(), t_h.data_ptr(), sc.num_beams, sc.d_center, sc.d_theta, None,
                       B.POLICY_ALWAYS, -1.0)
            neffs.append(h.read_neff())
        outs.append((np.array(neffs).tobytes(), h.poses().tobytes(), h.weights().tobytes(), h.log_weights().tobytes(),
                     h.parents().tobytes(), h.get_map(0, B.MAP_FREE_COUNT).tobytes(),
                     h.get_map(0, B.MAP_LIKELIHOOD).tobytes(), h.weighted_pose().tobytes()))
        h.close()
    assert outs[0] == outs[1] == outs[2]


def test_shared_map_side_limit(cuda):
    """The shared-map ray kernels pack a cell as x | y << 16: wider grids are rejected at creation."""
    with pytest.raises(B.GmsError) as e:
        cuda.create(num_particles=4, map_mode=B.MAP_SHARED, map_width_m=3400.0, map_height_m=1.0, resolution=0.05)
    assert e.value.code == B.ERR_INVALID_ARG


def test_sorted_update_mode_equals_atomic(cuda, oracle):
    """GMS_UPDATE_SORTED (atomic-free: key sort + run lengths + one writer per cell) produces the same counts,
    dirty tiles and likelihood field as the default integer atomics — and as the oracle."""
    from gridmap_slam_robot_b200 import synth

    P, steps = 64, 5
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0, map_mode=B.MAP_SHARED)
    a, s_, o = cuda.create(**kw), cuda.create(update_mode=B.UPDATE_SORTED, **kw), oracle.create(**kw)
    scans = synth.make_scans(steps, 360)
    normals, uniforms = synth.make_draws(steps, P)
    for i, sc in enumerate(scans):
        for h in (a, s_, o):
            h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[i])
            h.resample(float(uniforms[i]))
        for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT, B.MAP_LIKELIHOOD):
            m = a.get_map(0, kind)
            assert np.array_equal(m, s_.get_map(0, kind)) and np.array_equal(m, o.get_map(0, kind)), (i, kind)
    for h in (a, s_, o):
        h.close()


def test_truncation_toward_zero(cuda):
    checks.check_truncation_toward_zero(cuda)


def test_weight_underflow_log_domain(cuda):
    checks.check_underflow(cuda)


# ---- BASELINE configs 2 and 3 at their full sizes, piecewise against the oracle's per-map operators ----
def test_full_size_k3_sampled_parity(cuda, oracle):
    """K3: 10k particles x 720 beams, 2048^2 shared grid, 12 m range."""
    checks.check_full_size_sampled(cuda, oracle, P=10000, beams=720, grid_m=102.4, max_range=12.0)


def test_full_size_k4_sampled_parity(cuda, oracle):
    """K4: 100k particles x 720 beams, 4096^2 shared grid, every beam hits (the bench workload)."""
    checks.check_full_size_sampled(cuda, oracle, P=100000, beams=720, grid_m=204.8, max_range=30.0)


def test_full_size_per_particle_maps_sampled_parity(cuda, oracle):
    """BASELINE config 1 with the reference's own map semantics (and config 4's per-GPU share): 1k particles x 360
    beams, each particle with its own 1024^2 map."""
    checks.check_full_size_sampled_pp(cuda, oracle, P=1000, beams=360, grid_m=51.2)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_per_particle_operator_sequences_match_oracle(cuda, oracle, seed):
    """Per-particle maps keep no likelihood field and integrate a scan only where a map survives:
    random interleavings of updates, resamplings, pose injection and the per-map GridMap operators must leave
    exactly what the reference's eager arrays would hold (GridMap.java:106-124,173-250; SLAM.java:80-153) —
    counters and likelihoodData bit for bit, parents equal, log-probabilities to 1e-9."""
    from gridmap_slam_robot_b200 import synth

    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    P, beams, steps = 6, 90, 40
    kw = dict(num_particles=P, map_width_m=10.0, map_height_m=10.0, origin_x=-5.0, origin_y=-5.0,
              map_mode=B.MAP_PER_PARTICLE, resample_mode=B.RESAMPLE_LITERAL)
    g, o = cuda.create(**kw), oracle.create(**kw)
    scans = synth.make_scans(steps, beams, max_range=4.0)
    normals, uniforms = synth.make_draws(steps, P, seed=seed)
    cells = g.W * g.H

    def compare(tag):
        for i in range(P):
            for kind in (B.MAP_FREE_COUNT, B.MAP_OCC_COUNT, B.MAP_LIKELIHOOD):
                assert np.array_equal(g.get_map(i, kind), o.get_map(i, kind)), (tag, i, kind)
        assert np.array_equal(g.poses(), o.poses()), tag
        assert np.array_equal(g.parents(), o.parents()), tag

    k = 0
    for it in range(steps):
        op = rng.choice(["update", "update", "update_resample", "resample", "set_poses", "apply", "integrate",
                         "compute", "prob", "set_counts", "read_field", "read_counts"])
        p = int(rng.integers(P))
        sc = scans[k % steps]
        if op in ("update", "update_resample"):
            d_theta = sc.d_theta if rng.random() > 0.15 else 1.0  # > 30 degrees: the update skips the integration
            for h in (g, o):
                h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, d_theta, normals[k % steps])
            k += 1
            if op == "update_resample":
                for h in (g, o):
                    h.resample(float(uniforms[it]))
        elif op == "resample":
            for h in (g, o):
                h.resample(float(uniforms[it]))
        elif op == "set_poses":
            xyt = rng.uniform(-1.0, 1.0, size=(P, 3)).astype(np.float32)
            for h in (g, o):
                h.set_poses(xyt)
        elif op == "apply":
            a = rng.uniform(20.0, 180.0, size=4).astype(np.float32)
            for h in (g, o):
                h.map_apply_measurement(p, float(a[0]), float(a[1]), float(a[2]), float(a[3]), 30.0, bool(it & 1))
        elif op == "integrate":
            pose = rng.uniform(-1.0, 1.0, size=3).astype(np.float32)
            for h in (g, o):
                h.map_integrate_observation(p, pose, sc.beam_xy, sc.beam_dist, sc.beam_hit)
        elif op == "compute":
            for h in (g, o):
                h.map_compute_likelihood(p)
        elif op == "prob":
            pose = rng.uniform(-0.5, 0.5, size=3).astype(np.float32)
            lg, _ = g.map_probability_of(p, pose, sc.beam_xy, sc.beam_hit)
            lo_, _ = o.map_probability_of(p, pose, sc.beam_xy, sc.beam_hit)
            assert abs(lg - lo_) <= 1e-9 * max(1.0, abs(lo_)), (it, lg, lo_)
        elif op == "set_counts":
            nf = rng.integers(0, 3, size=cells).astype(np.uint32)
            no = rng.integers(0, 2, size=cells).astype(np.uint32)
            for h in (g, o):
                h.set_map_counts(p, nf, no)
        elif op == "read_field":
            assert np.array_equal(g.get_map(p, B.MAP_LIKELIHOOD), o.get_map(p, B.MAP_LIKELIHOOD)), (it, op)
        else:
            assert np.array_equal(g.get_map(p, B.MAP_FREE_COUNT), o.get_map(p, B.MAP_FREE_COUNT)), (it, op)
        if it % 4 == 3:
            compare((it, op))
    compare("end")
    g.close()
    o.close()
