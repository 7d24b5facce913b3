"""Rows adjacent to the hot path: recording file format, scan de-skew, renderer
hand-off, combined-map fusion."""
import struct

import numpy as np
import pytest

import checks
from gridmap_slam_robot_b200 import binding as B
from gridmap_slam_robot_b200 import recording


def test_recording_known_bytes():
    """Hand-assembled file, byte for byte what DataOutputStream writes (DataRecorder.java:381-399)."""
    raw = bytes([0xFF]) + struct.pack(">h", 1) + struct.pack(">f", 1.5) + struct.pack(">dd", 0.05, -0.01) + \
        struct.pack(">h", 2) + struct.pack(">ddB", 0.25, 3.5, 1) + struct.pack(">ddB", 0.5, 10.0, 0)
    frames = recording.loads(raw)
    assert len(frames) == 1 and frames[0].time_stamp == 1.5 and frames[0].d_center == 0.05 and frames[0].d_theta == -0.01
    assert frames[0].angle.tolist() == [0.25, 0.5] and frames[0].distance.tolist() == [3.5, 10.0]
    assert frames[0].was_hit.tolist() == [1, 0]
    assert recording.dumps(frames) == raw
    with pytest.raises(ValueError):
        recording.loads(b"\x00" + raw[1:])


def test_recording_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    frames = [recording.RecordedFrame(float(np.float32(i * 0.1)), rng.normal(), rng.normal(), rng.uniform(-3, 3, n),
                                      rng.uniform(0, 12, n), (rng.random(n) < 0.8).astype(np.uint8))
              for i, n in enumerate((0, 1, 90, 720))]
    p = tmp_path / "rec.bin"
    recording.save(p, frames)
    back = recording.load(p)
    assert len(back) == 4
    for a, b in zip(frames, back):
        assert (a.time_stamp, a.d_center, a.d_theta) == (b.time_stamp, b.d_center, b.d_theta)
        assert np.array_equal(a.angle, b.angle) and np.array_equal(a.distance, b.distance)
        assert np.array_equal(a.was_hit, b.was_hit)


def test_deskew_oracle(oracle):
    checks.check_deskew_reference_formula(oracle)


def test_replay_recording_through_oracle(oracle, tmp_path):
    """A recorded session replayed through update_raw (the reference's replay path)."""
    sweeps = checks._raw_sweeps(3, 60)
    frames = [recording.RecordedFrame(0.1 * i, dc, dth, a, d, h) for i, (a, d, h, dc, dth) in enumerate(sweeps)]
    recording.save(tmp_path / "s.bin", frames)
    h = oracle.create(num_particles=8, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    for f in recording.load(tmp_path / "s.bin"):
        neff = h.update_raw(f.angle, f.distance, f.was_hit, f.d_center, f.d_theta, np.zeros((8, 2)))
        assert 1.0 <= neff <= 8.0 + 1e-9
    assert h.get_map(0, B.MAP_OCC_COUNT).sum() > 0
    h.close()


@pytest.mark.gpu
def test_deskew_cuda(cuda):
    checks.check_deskew_reference_formula(cuda)


@pytest.mark.gpu
def test_next_rows_cuda_vs_oracle(cuda, oracle):
    checks.check_next_rows(cuda, oracle)
