"""make_golden.py — writes tests/golden/*.npz from the pure-Python restatement (oracle/pyref.py).

Run from the repo root:  python oracle/make_golden.py
The fixtures pin the C oracle (tests/test_oracle.py, CPU) and the CUDA path
(tests/test_gpu_parity.py, -m gpu).  They are NOT outputs of the reference itself — the Java
reference cannot run in this image (no JDK); see the header of oracle/gms_ref.c.
Inputs are drawn from numpy PCG64 with fixed seeds; every float32-typed input is stored as float32
and every double as float64, so the files are exact.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def f32a(x):
    return np.asarray(x, dtype=np.float32)


def gen_rays():
    rng = np.random.Generator(np.random.PCG64(101))
    W, H, extra = 37, 29, 2
    rays = []
    # random rays, some starting or ending outside the grid
    for _ in range(160):
        rays.append(rng.uniform(-4, 41, size=4) * [1, 29 / 37, 1, 29 / 37])
    # degenerate: dx == 0, dy == 0, zero length, exact diagonals, integer coordinates, reversed
    for _ in range(12):
        x, y, x1, y1 = rng.uniform(1, 27, size=4)
        rays += [[x, y, x, y1], [x, y, x1, y], [x, y, x, y]]
        a, b, n = int(x), int(y), int(rng.integers(1, 9))
        rays += [[a, b, a + n, b + n], [a + n, b + n, a, b], [a + 0.5, b + 0.5, a + 0.5 + n, b + 0.5 - n], [a, b, a + n, b]]
    rays = f32a(rays)
    cap = 2 * (W + H) + extra + 4
    cells = np.full((len(rays), cap, 2), -1, np.int32)
    counts = np.zeros(len(rays), np.int32)
    n0 = np.zeros(len(rays), np.int32)
    e0 = np.zeros(len(rays), np.float32)
    for i, r in enumerate(rays):
        c, n, e = pyref.ray_cells(W, H, *(float(v) for v in r), extra)
        counts[i] = len(c)
        n0[i], e0[i] = n, e
        if c:
            cells[i, : len(c)] = c
    np.savez_compressed(os.path.join(OUT, "pyref_rays.npz"), W=W, H=H, extra=extra, rays=rays, cells=cells,
                        counts=counts, n0=n0, e0=e0)


def small_map():
    # 3.0 m x 2.5 m at 0.05 m -> 60 x 50 cells, lower-left at (-1.5, -1.25)
    return pyref.PyGridMap(3.0, 2.5, 0.05, (-1.5, -1.25))


def gen_apply():
    rng = np.random.Generator(np.random.PCG64(202))
    gm = small_map()
    m = gm.create_map()
    n = 96
    args = np.zeros((n, 5), np.float32)
    hits = np.zeros(n, np.uint8)
    for i in range(n):
        sx, sy = rng.uniform(5, 55), rng.uniform(5, 45)
        ang, ln = rng.uniform(0, 2 * math.pi), rng.uniform(0.0, 40.0)
        ex, ey = sx + ln * math.cos(ang), sy + ln * math.sin(ang)
        meas = ln * rng.choice([1.0, 1.0, 0.7, 1.2])
        args[i] = (sx, sy, ex, ey, meas)
        hits[i] = rng.random() < 0.7
    for i in range(n):
        a = [float(v) for v in args[i]]
        gm.apply_measurement(m, a[0], a[1], a[2], a[3], a[4], bool(hits[i]))
    np.savez_compressed(os.path.join(OUT, "pyref_apply.npz"), args=args, hits=hits,
                        nfree=np.asarray(m["nfree"], np.uint32).reshape(gm.H, gm.W),
                        nocc=np.asarray(m["nocc"], np.uint32).reshape(gm.H, gm.W),
                        log=np.asarray(m["log"], np.float64).reshape(gm.H, gm.W))


def gen_blur():
    rng = np.random.Generator(np.random.PCG64(303))
    gm = small_map()
    # counts -> sign pattern; includes F U O rows that cancel to ~0.5
    nfree = rng.integers(0, 3, size=(gm.H, gm.W)).astype(np.uint32)
    nocc = (rng.random((gm.H, gm.W)) < 0.25).astype(np.uint32)
    nfree[rng.random((gm.H, gm.W)) < 0.4] = 0
    nfree[10, 10:13] = (1, 0, 0)
    nocc[10, 10:13] = (0, 0, 1)
    m = gm.create_map()
    m["log"] = [float(f) * gm.l_free + float(o) * gm.l_occ for f, o in zip(nfree.ravel(), nocc.ravel())]
    gm.compute_likelihood(m)
    np.savez_compressed(os.path.join(OUT, "pyref_blur.npz"), nfree=nfree, nocc=nocc,
                        lik=np.asarray(m["lik"], np.float64).reshape(gm.H, gm.W),
                        kernel=np.asarray(gm.kernel, np.float64))


def gen_motion():
    rng = np.random.Generator(np.random.PCG64(404))
    n = 256
    poses = f32a(np.stack([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(-math.pi, math.pi, n)], 1))
    poses[:8, 2] = f32a([math.pi, -math.pi, 0.0, 3.1415925, -3.1415925, 1e-8, 3.0, -3.0])
    dc = rng.uniform(-0.2, 0.3, n)
    dt = rng.uniform(-0.7, 0.7, n)
    z = rng.standard_normal((n, 2))
    out = np.zeros((n, 3), np.float32)
    for i in range(n):
        out[i] = pyref.motion_sample(tuple(float(v) for v in poses[i]), float(dc[i]), float(dt[i]), float(z[i, 0]),
                                     float(z[i, 1]))
    np.savez_compressed(os.path.join(OUT, "pyref_motion.npz"), poses=poses, d_center=dc, d_theta=dt, z=z, out=out)


def gen_resample():
    rng = np.random.Generator(np.random.PCG64(505))
    cases = {}
    k = 0
    for n in (1, 2, 7, 64, 500):
        for _ in range(4):
            w = rng.random(n) ** rng.choice([1, 4, 12])
            if n > 2 and rng.random() < 0.5:
                w[rng.integers(0, n, size=n // 3)] = 0.0
            w = w / w.sum()
            u = float(rng.random())
            cases[f"w{k}"] = w
            cases[f"u{k}"] = np.float64(u)
            cases[f"lit{k}"] = np.asarray(pyref.resample_indices(list(map(float, w)), u), np.int32)
            cases[f"fix{k}"] = np.asarray(pyref.resample_indices_fixed(list(map(float, w)), u), np.int32)
            cases[f"neff{k}"] = np.float64(pyref.neff(list(map(float, w))))
            k += 1
    cases["num_cases"] = np.int32(k)
    np.savez_compressed(os.path.join(OUT, "pyref_resample.npz"), **cases)


def tiny_scans(rng, steps, B):
    """A robot in a ~1 m box: short beams so the 60x50 map is exercised incl. its borders."""
    scans = []
    ang = 2 * math.pi * np.arange(B) / B
    for s in range(steps):
        dist = rng.uniform(0.35, 1.1, B)
        hit = rng.random(B) < 0.8
        dist = np.where(hit, dist, 1.6)  # misses run out of the map on the short sides
        xy = np.stack([dist * np.cos(ang), dist * np.sin(ang)], 1)
        dc = float(rng.uniform(0.0, 0.06))
        dth = float(rng.uniform(-0.12, 0.12)) if s != 3 else 0.6  # step 3: |dTheta| > 30 deg -> skip map update
        scans.append((xy, dist, hit.astype(np.uint8), dc, dth))
    return scans


def gen_slam(shared):
    rng = np.random.Generator(np.random.PCG64(606 + int(shared)))
    gm = small_map()
    P, B, steps = 6, 20, 6
    slam = pyref.PySLAM(P, gm, shared=shared)
    scans = tiny_scans(rng, steps, B)
    normals = rng.standard_normal((steps, P, 2))
    uniforms = rng.random(steps)
    out = {"P": P, "B": B, "steps": steps, "normals": normals, "uniforms": uniforms, "shared": int(shared)}
    for s, (xy, dist, hit, dc, dth) in enumerate(scans):
        beams = [(float(xy[b, 0]), float(xy[b, 1]), float(dist[b]), bool(hit[b])) for b in range(B)]
        ne = slam.update(beams, dc, dth, [float(v) for v in normals[s].ravel()])
        out[f"xy{s}"], out[f"dist{s}"], out[f"hit{s}"] = xy, dist, hit
        out[f"dc{s}"], out[f"dth{s}"] = np.float64(dc), np.float64(dth)
        out[f"neff{s}"] = np.float64(ne)
        out[f"lw{s}"] = np.asarray(slam.lw, np.float64)
        out[f"w{s}"] = np.asarray(slam.w, np.float64)
        out[f"wlit{s}"] = np.asarray(slam.wlit, np.float64)
        out[f"strongest{s}"] = np.int32(slam.strongest)
        out[f"poses_upd{s}"] = f32a(slam.poses)
        out[f"wpose{s}"] = f32a(slam.weighted_pose())
        if s in (1, 2, 4):  # resample on a fixed schedule (the caller decides: GridMapApp.java:185-186)
            out[f"parents{s}"] = np.asarray(slam.resample(float(uniforms[s])), np.int32)
        out[f"poses{s}"] = f32a(slam.poses)
        nm = len(slam.maps)
        out[f"nfree{s}"] = np.asarray([m["nfree"] for m in slam.maps], np.uint32).reshape(nm, gm.H, gm.W)
        out[f"nocc{s}"] = np.asarray([m["nocc"] for m in slam.maps], np.uint32).reshape(nm, gm.H, gm.W)
        out[f"lik{s}"] = np.asarray([m["lik"] for m in slam.maps], np.float64).reshape(nm, gm.H, gm.W)
        out[f"log{s}"] = np.asarray([m["log"] for m in slam.maps], np.float64).reshape(nm, gm.H, gm.W)
    np.savez_compressed(os.path.join(OUT, "pyref_slam_shared.npz" if shared else "pyref_slam_pp.npz"), **out)


def gen_underflow():
    """400 hit beams on a blank map: Java's product 0.1^400 underflows to 0 -> weightSum 0 -> NaN weights
    (SLAM.java:121); the log-domain weights stay finite and uniform."""
    rng = np.random.Generator(np.random.PCG64(808))
    gm = small_map()
    P, B = 3, 400
    slam = pyref.PySLAM(P, gm, shared=True)
    ang = 2 * math.pi * np.arange(B) / B
    dist = rng.uniform(0.4, 1.0, B)
    xy = np.stack([dist * np.cos(ang), dist * np.sin(ang)], 1)
    hit = np.ones(B, np.uint8)
    beams = [(float(xy[b, 0]), float(xy[b, 1]), float(dist[b]), True) for b in range(B)]
    normals = np.zeros((P, 2))
    ne = slam.update(beams, 0.0, 0.0, [0.0] * (2 * P))
    np.savez_compressed(os.path.join(OUT, "pyref_underflow.npz"), xy=xy, dist=dist, hit=hit, normals=normals,
                        neff=np.float64(ne), lw=np.asarray(slam.lw, np.float64), w=np.asarray(slam.w, np.float64),
                        wlit=np.asarray(slam.wlit, np.float64))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_rays()
    gen_apply()
    gen_blur()
    gen_motion()
    gen_resample()
    gen_slam(False)
    gen_slam(True)
    gen_underflow()
    print("golden fixtures written to", os.path.normpath(OUT))
