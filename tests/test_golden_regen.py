"""The committed fixtures under tests/golden/ are exactly what oracle/make_golden.py (the independent pure-Python
restatement) writes: regenerate them into a scratch directory and compare array by array."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_fixtures_are_reproducible(tmp_path):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.OUT = str(tmp_path)
    for fn in (mg.gen_rays, mg.gen_apply, mg.gen_blur, mg.gen_motion, mg.gen_resample):
        fn()
    mg.gen_slam(False)
    mg.gen_slam(True)
    mg.gen_underflow()
    committed = os.path.join(ROOT, "tests", "golden")
    names = sorted(f for f in os.listdir(committed) if f.endswith(".npz"))
    assert names == sorted(os.listdir(tmp_path))
    for name in names:
        a, b = np.load(os.path.join(committed, name)), np.load(os.path.join(tmp_path, name))
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k], equal_nan=True), (name, k)
