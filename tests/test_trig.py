"""sin / cos on the hot path (MathUtil.java:30-48 -> FastMath, commons-math3, not in the reference tree) are evaluated
by the oracle, by oracle/pyref.py and by the CUDA library as ONE fixed sequence of IEEE f64 operations (the published
fdlibm 5.3 scheme), so none of them can differ from the others in the last bit.  Here: the sequence is accurate
(< 1 ulp against an exact sine / cosine, i.e. the same class as FastMath / libm), the C oracle and the Python
restatement agree bit for bit on every branch of it, and (GPU) so does the CUDA library."""
import math

import numpy as np
import pytest

import checks


def test_fixed_sequence_is_accurate_to_an_ulp():
    mp = pytest.importorskip("mpmath")
    ref = checks._pyref()
    mp.mp.prec = 400
    worst = 0.0
    for x in checks.trig_probe_angles()[::3]:
        x = float(x)
        if not abs(x) < 1647099.0:
            continue
        for f, g in ((ref.sin_fixed, mp.sin), (ref.cos_fixed, mp.cos)):
            exact = g(mp.mpf(x))
            v = f(x)
            u = math.ulp(float(exact)) if float(exact) != 0.0 else 5e-324
            worst = max(worst, float(abs((mp.mpf(v) - exact) / u)))
    assert worst < 1.0, worst


def test_fixed_sequence_close_to_libm():
    """Independent of mpmath: within 1 ulp of glibc (itself < 1 ulp), equal in the vast majority of cases."""
    ref = checks._pyref()
    a = checks.trig_probe_angles()
    a = a[np.abs(a) < 1647099.0]
    s = np.asarray([ref.sin_fixed(float(v)) for v in a])
    c = np.asarray([ref.cos_fixed(float(v)) for v in a])
    for got, want in ((s, np.sin(a)), (c, np.cos(a))):
        assert np.all(np.abs(got - want) <= 2 * np.spacing(np.abs(want)))
        assert np.mean(got == want) > 0.9
    # signs and symmetry on the no-reduction interval
    assert ref.sin_fixed(0.0) == 0.0 and math.copysign(1.0, ref.sin_fixed(-0.0)) == -1.0 and ref.cos_fixed(0.0) == 1.0
    assert math.isnan(ref.sin_fixed(float("inf"))) and math.isnan(ref.cos_fixed(float("nan")))


def test_trig_probe_oracle(oracle):
    checks.check_trig_probe(oracle)


@pytest.mark.gpu
def test_trig_probe_cuda(cuda):
    checks.check_trig_probe(cuda)


def test_fixed_sequence_symmetries():
    """Exact identities of the operation sequence (they follow from IEEE negation being exact and the reduction being
    odd in x): sin(-x) == -sin(x), cos(-x) == cos(x) bit for bit; sin^2 + cos^2 within 3 ulp of 1."""
    ref = checks._pyref()
    for x in checks.trig_probe_angles()[::2]:
        x = float(x)
        if not abs(x) < 1647099.0:
            continue
        s, c = ref.sin_fixed(x), ref.cos_fixed(x)
        assert ref.sin_fixed(-x) == -s and ref.cos_fixed(-x) == c, x
        assert abs(s * s + c * c - 1.0) <= 3 * 2.220446049250313e-16, x


def test_motion_step_uses_the_fixed_sequence(oracle):
    """Odometry.apply (Odometry.java:77-96): x' = (float)(x + (double)(float)cos(theta') * d) — the oracle's f32 pose after
    one noise-free motion step equals the value computed here from pyref's cos_fixed / sin_fixed, for headings
    spread over every quadrant (libm's cos differs from the sequence in ~3 % of f64 results, so a library call
    slipping back in would show up as an f32 mismatch only with probability 2^-29 per pose: the f64 probe through
    gms_deskew in test_trig_probe_* is the sharp test; this one pins the call site)."""
    ref = checks._pyref()
    from gridmap_slam_robot_b200 import binding as B

    P = 256
    h = oracle.create(num_particles=P, map_width_m=4.0, map_height_m=4.0, origin_x=-2.0, origin_y=-2.0,
                      map_mode=B.MAP_SHARED)
    rng = np.random.default_rng(7)
    poses = np.stack([rng.uniform(-1, 1, P), rng.uniform(-1, 1, P), rng.uniform(-3.1, 3.1, P)], 1).astype(np.float32)
    h.set_poses(poses)
    xy = np.zeros((0, 2)); dist = np.zeros(0); hit = np.zeros(0, np.uint8)
    dc, dth = 0.125, 0.0625
    h.update(xy, dist, hit, dc, dth, np.zeros((P, 2)))
    got = h.poses()
    for i in range(P):
        want = ref.motion_sample(tuple(float(v) for v in poses[i]), dc, dth, 0.0, 0.0)
        assert tuple(float(v) for v in got[i]) == tuple(want), i
    h.close()
