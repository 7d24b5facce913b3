"""Synthetic inputs for tests and bench.

World: axis-aligned square room with walls at +-9 m plus four fixed interior boxes; the robot drives
a 3 m-radius circle (dCenter = 0.05 m, dTheta = 0.05/3 rad per step, well below the 30 degree
skip threshold of SLAM.java:82).  A scan is B beams at angle 2*pi*b/B in the robot frame, exact
ray/segment intersection + N(0, 0.01 m) range noise; ranges beyond `max_range` become misses with
distance = max_range, hit = false (ConnectionThread.java:78-79).  Beams are delivered the way the
hot path consumes them (Observation.Measurement, Observation.java:44-51):
localX = distance*cos(angle), localY = distance*sin(angle), distance, wasHit.
"""
from __future__ import annotations

import dataclasses

import numpy as np

ROOM_HALF = 9.0
BOXES = ((-6.0, -5.0, -4.0, -2.5), (3.5, 4.0, 6.5, 6.0), (4.5, -6.5, 6.0, -4.0), (-5.5, 3.0, -3.5, 5.5))
TRAJ_RADIUS = 3.0
STEP_LEN = 0.05


def _segments():
    segs = []
    h = ROOM_HALF
    rects = ((-h, -h, h, h),) + BOXES
    for (x0, y0, x1, y1) in rects:
        segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
    return np.asarray(segs, dtype=np.float64)


_SEGS = _segments()


def raycast(x, y, angles):
    """Exact distance from (x, y) along world-frame `angles` to the nearest wall segment."""
    dx, dy = np.cos(angles)[:, None], np.sin(angles)[:, None]
    ax, ay, bx, by = (_SEGS[None, :, i] for i in range(4))
    ex, ey = bx - ax, by - ay
    den = dx * ey - dy * ex
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((ax - x) * ey - (ay - y) * ex) / den
        s = ((ax - x) * dy - (ay - y) * dx) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (s >= 0.0) & (s <= 1.0)
    t = np.where(ok, t, np.inf)
    return t.min(axis=1)


def true_pose(step):
    """True pose after `step` odometry increments.  The robot starts at the world origin with heading
    0 — the pose SLAM.reset gives every particle (SLAM.java:65-77) — so the SLAM frame IS the world
    frame; the circle is centred at (0, R)."""
    a = step * (STEP_LEN / TRAJ_RADIUS)
    return TRAJ_RADIUS * np.sin(a), TRAJ_RADIUS * (1.0 - np.cos(a)), a


@dataclasses.dataclass
class Scan:
    beam_xy: np.ndarray  # f64 [B, 2]: localX, localY
    beam_dist: np.ndarray  # f64 [B]
    beam_hit: np.ndarray  # u8  [B]
    d_center: float
    d_theta: float

    @property
    def num_beams(self):
        return int(self.beam_dist.shape[0])

    @property
    def num_hits(self):
        return int(self.beam_hit.sum())


def make_scans(num_steps, num_beams, max_range=10.0, seed=20260101, noise_sd=0.01, room_scale=1.0):
    """`num_steps` consecutive scans along the trajectory (scan k is taken at true_pose(k + 1)).
    `room_scale` enlarges the room and its boxes about the origin (bench.py's gather-stress workload); the
    trajectory stays the 3 m circle."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rel = 2.0 * np.pi * np.arange(num_beams, dtype=np.float64) / num_beams
    scans = []
    for k in range(num_steps):
        x, y, th = true_pose(k + 1)
        dist = room_scale * raycast(x / room_scale, y / room_scale, rel + th) + rng.normal(0.0, noise_sd, size=num_beams)
        hit = dist <= max_range
        dist = np.where(hit, np.maximum(dist, 0.02), max_range)
        xy = np.stack([dist * np.cos(rel), dist * np.sin(rel)], axis=1)
        scans.append(Scan(np.ascontiguousarray(xy), np.ascontiguousarray(dist), hit.astype(np.uint8),
                          STEP_LEN, STEP_LEN / TRAJ_RADIUS))
    return scans


def make_draws(num_steps, num_particles, seed=7):
    """Injected randomness shared by oracle and CUDA: standard normals [step][P][2] ({z_d, z_theta},
    the two NormalDistribution.sample() calls of Odometry.java:80-81) and one resampling uniform
    per step (Math.random(), SLAM.java:136)."""
    rng = np.random.Generator(np.random.Philox(seed))
    normals = rng.standard_normal(size=(num_steps, num_particles, 2))
    uniforms = rng.random(size=num_steps)
    return normals, uniforms
