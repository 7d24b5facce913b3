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
