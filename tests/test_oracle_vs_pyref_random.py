"""The C oracle against the independent pure-Python restatement (oracle/pyref.py) on seeded random inputs that are
NOT in tests/golden/: a second line of defence against a fixture that happens to miss a branch.  CPU only."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pyref  # noqa: E402

from gridmap_slam_robot_b200 import binding as B  # noqa: E402

SEEDS = [1, 2, 3, 4, 5]


@pytest.mark.parametrize("seed", SEEDS)
def test_rays(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(3, 45)), int(rng.integers(3, 45))
    extra = int(rng.integers(0, 4))
    h = oracle.create(num_particles=1, map_width_m=W * 0.05, map_height_m=H * 0.05, resolution=0.05,
                      map_mode=B.MAP_SHARED)
    assert (h.W, h.H) == (W, H)
    rays = rng.uniform(-3, 48, size=(120, 4)).astype(np.float32)
    rays[:20] = np.floor(rays[:20])  # integer coordinates: the error term hits exact zeros
    rays[20:30, 2:] = rays[20:30, :2]  # zero-length
    cells, counts = h.trace_rays(rays, extra=extra)
    for i, r in enumerate(rays):
        c, _, _ = pyref.ray_cells(W, H, *(float(v) for v in r), extra)
        assert counts[i] == len(c), (seed, i, r)
        assert [tuple(x) for x in cells[i, : len(c)]] == [tuple(x) for x in c], (seed, i, r)
    h.close()


def random_beams(rng, n):
    ang = rng.uniform(0, 2 * math.pi, n)
    dist = rng.uniform(0.05, 1.4, n)
    hit = rng.random(n) < 0.75
    xy = np.stack([dist * np.cos(ang), dist * np.sin(ang)], 1)
    return xy, dist, hit.astype(np.uint8)


@pytest.mark.parametrize("seed", SEEDS)
def test_integrate_likelihood_probability(oracle, seed):
    """GridMap.integrateObservation -> computeLikelihoodMap -> probabilityOf from random poses, twice over."""
    rng = np.random.default_rng(2000 + seed)
    gm = pyref.PyGridMap(2.0, 1.5, 0.05, (-1.0, -0.75))
    m = gm.create_map()
    h = oracle.create(num_particles=1, map_width_m=2.0, map_height_m=1.5, resolution=0.05, origin_x=-1.0,
                      origin_y=-0.75)
    assert (h.W, h.H) == (gm.W, gm.H)
    for rnd in range(2):
        pose = np.asarray([rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4), rng.uniform(-math.pi, math.pi)], np.float32)
        xy, dist, hit = random_beams(rng, 24)
        beams = [(float(xy[b, 0]), float(xy[b, 1]), float(dist[b]), bool(hit[b])) for b in range(len(dist))]
        gm.integrate_observation(m, beams, tuple(float(v) for v in pose))
        h.map_integrate_observation(0, pose, xy, dist, hit)
        assert np.array_equal(h.get_map(0, B.MAP_FREE_COUNT).ravel(), np.asarray(m["nfree"], np.uint32))
        assert np.array_equal(h.get_map(0, B.MAP_OCC_COUNT).ravel(), np.asarray(m["nocc"], np.uint32))
        gm.compute_likelihood(m)
        h.map_compute_likelihood(0)
        assert np.array_equal(h.get_map(0, B.MAP_LIKELIHOOD).ravel(), np.asarray(m["lik"], np.float64))
        for _ in range(4):
            q = np.asarray([rng.uniform(-1.1, 1.1), rng.uniform(-0.9, 0.9), rng.uniform(-4, 4)], np.float32)
            prod, lsum = gm.probability_of(m, beams, tuple(float(v) for v in q))
            lp, p = h.map_probability_of(0, q, xy, hit)
            assert abs(lp - lsum) <= 1e-12 * max(1.0, abs(lsum))
            assert p == prod or abs(p / prod - 1) < 1e-12
    h.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_motion(oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    h = oracle.create(num_particles=1, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED)
    motion = oracle.dll.gmsref_motion
    motion.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_double, C.c_double, C.c_double, C.c_double]
    for _ in range(200):
        pose = np.asarray([rng.uniform(-9, 9), rng.uniform(-9, 9), rng.uniform(-math.pi, math.pi)], np.float32)
        dc, dth = float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-1.2, 1.2))
        zd, zt = (float(v) for v in rng.standard_normal(2) * rng.choice([1.0, 4.0]))
        p = (C.c_float * 3)(*pose)
        assert motion(h.h, p, dc, dth, zd, zt) == 0
        want = pyref.motion_sample(tuple(float(v) for v in pose), dc, dth, zd, zt)
        assert tuple(np.float32(v) for v in p[:]) == tuple(np.float32(v) for v in want)
    h.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_resample_and_neff(oracle, seed):
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(2, 400))
    w = rng.random(n) ** int(rng.choice([1, 3, 9]))
    w[rng.integers(0, n, size=n // 4)] = 0.0
    w[int(rng.integers(0, n))] += 1e-3  # never all zero
    w /= w.sum()
    u = float(rng.random())
    for mode, fn in ((B.RESAMPLE_LITERAL, pyref.resample_indices), (B.RESAMPLE_FIXED, pyref.resample_indices_fixed)):
        h = oracle.create(num_particles=n, map_width_m=0.5, map_height_m=0.5, map_mode=B.MAP_SHARED, resample_mode=mode)
        h.set_weights(w)
        assert abs(h.calculate_neff() / pyref.neff(list(map(float, w))) - 1) < 1e-12
        h.resample(u)
        assert h.parents().tolist() == fn(list(map(float, w)), u), (seed, mode)
        h.close()
