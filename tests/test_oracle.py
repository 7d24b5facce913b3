"""CPU: the C oracle against (1) the hand-derived vectors of SURVEY.md Appendix B and (2) the golden
fixtures written by the independent pure-Python restatement (oracle/make_golden.py)."""
import numpy as np
import pytest

import checks


def test_constants(oracle):
    checks.check_constants(oracle)


def test_appendix_b_rays(oracle):
    checks.check_appendix_b_rays(oracle)


def test_appendix_b_resample(oracle):
    checks.check_appendix_b_resample(oracle)


def test_blank_likelihood(oracle):
    checks.check_blank_likelihood(oracle)


def test_golden_rays(oracle):
    checks.check_rays(oracle)


def test_golden_apply(oracle):
    checks.check_apply(oracle)


def test_golden_blur(oracle):
    checks.check_blur(oracle)


def test_golden_resample(oracle):
    checks.check_resample(oracle)


def test_golden_motion(oracle):
    checks.check_motion(oracle)


@pytest.mark.parametrize("shared", [False, True])
def test_golden_slam(oracle, shared):
    checks.check_slam(oracle, shared)


def test_errors(oracle):
    checks.check_errors(oracle)


def test_sign_of_count_form_is_robust():
    """The CUDA path thresholds nFree*L_free + nOcc*L_occ instead of Java's sequentially accumulated
    f64 sum.  The two can only disagree in sign if the exact value is within rounding error of 0;
    show the closest approach over all count pairs up to 4096 is > 1e-4 (rounding error < 1e-11)."""
    import appendix_b as AB

    nf = np.arange(0, 4097, dtype=np.float64)[:, None]
    no = np.arange(0, 4097, dtype=np.float64)[None, :]
    v = np.abs(nf * AB.L_FREE + no * AB.L_OCC)
    v[0, 0] = np.inf
    assert v.min() > 1e-4


def test_truncation_toward_zero(oracle):
    checks.check_truncation_toward_zero(oracle)


def test_weight_underflow_log_domain(oracle):
    checks.check_underflow(oracle)


def test_full_size_sampled_check_is_self_consistent(oracle):
    """The piecewise full-size check of the GPU suite (checks.check_full_size_sampled), run here on the oracle
    itself at a size a CPU finishes in seconds: the per-map operators compose to exactly the step."""
    checks.check_full_size_sampled(oracle, oracle, P=3000, beams=90, grid_m=25.6, steps=3, sample=16)


def test_full_size_sampled_pp_check_is_self_consistent(oracle):
    checks.check_full_size_sampled_pp(oracle, oracle, P=24, beams=90, grid_m=25.6, steps=3, sample=4)


def test_pose_optimizer_hook_default_is_identity(oracle):
    """A4 (GridMap.findBestPoseOptim, SLAM.java:97): a hook that changes nothing gives the results of no hook, bit
    for bit; a hook that moves the poses is what gets scored and integrated."""
    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import synth

    P = 20
    scans = synth.make_scans(3, 90)
    normals, _ = synth.make_draws(3, P)
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    outs = []
    for hook in (None, lambda first, poses, *a: None, lambda first, poses, *a: poses.__setitem__(slice(None), poses + np.float32(0.05))):
        h = oracle.create(**kw)
        if hook:
            h.set_pose_optimizer(hook)
        for s, sc in enumerate(scans):
            h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta, normals[s])
        outs.append((h.poses().copy(), h.log_weights().copy(), h.get_map(2, B.MAP_FREE_COUNT).copy()))
        h.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert not np.array_equal(outs[0][0], outs[2][0]) and not np.array_equal(outs[0][2], outs[2][2])


def test_strongest_index_follows_its_first_child(oracle):
    from gridmap_slam_robot_b200 import synth

    P = 64
    sc = synth.make_scans(2, 90)
    z = synth.make_draws(2, P)[0]
    h = oracle.create(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    assert h.strongest()[0] == -1
    for s in range(2):
        h.update(sc[s].beam_xy, sc[s].beam_dist, sc[s].beam_hit, 0.05, 0.01, z[s])
    idx, pose, w = h.strongest()
    h.resample(0.31)
    idx2, pose2, w2 = h.strongest()
    assert idx2 == int(np.flatnonzero(h.parents() == idx)[0]) and np.array_equal(pose, pose2) and w == w2
    assert np.array_equal(h.poses()[idx2], pose)
    h.close()
