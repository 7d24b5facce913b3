"""Reader / writer of the reference's recording files.

Format (DataRecorder.save/load DataRecorder.java:381-436, ObjectSerializer.java:36-83; Java
DataOutputStream => big endian):

    u8    0xFF                                   header byte
    i16   nFrames
    per frame:
      f32   timeStamp
      f64   dCenter, f64 dTheta                  ObjectSerializer.writeOdometry
      i16   n                                    ObjectSerializer.writeObservation
      n x ( f64 angle, f64 distance, u8 wasHit ) ObjectSerializer.writeMeasurement

A loaded frame is the RAW sweep (angle, distance, wasHit): feed it to Handle.update_raw /
gms_update_raw, which applies the de-skew of GridMapApp.onHandleData before the SLAM step — the
replay path of the reference (DataRecorder.update -> DataEventHandler.publish -> onHandleData).
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

HEADER = 0xFF


@dataclasses.dataclass
class RecordedFrame:
    time_stamp: float
    d_center: float
    d_theta: float
    angle: np.ndarray  # f64 [n]
    distance: np.ndarray  # f64 [n]
    was_hit: np.ndarray  # u8 [n]


_MEAS = np.dtype([("angle", ">f8"), ("distance", ">f8"), ("hit", "u1")])


def dumps(frames) -> bytes:
    if len(frames) > 32767:
        raise ValueError("the format stores the frame count in a Java short")
    out = [struct.pack(">Bh", HEADER, len(frames))]
    for f in frames:
        n = int(len(f.angle))
        if n > 32767:
            raise ValueError("the format stores the measurement count in a Java short")
        out.append(struct.pack(">fddh", f.time_stamp, f.d_center, f.d_theta, n))
        rec = np.empty(n, _MEAS)
        rec["angle"], rec["distance"], rec["hit"] = f.angle, f.distance, np.asarray(f.was_hit, np.uint8) != 0
        out.append(rec.tobytes())
    return b"".join(out)


def loads(data: bytes):
    if len(data) < 3 or data[0] != HEADER:  # DataRecorder.java:411-416
        raise ValueError(f"header byte is not correct: wanted {HEADER}, got {data[0] if data else None}")
    (nframes,) = struct.unpack_from(">h", data, 1)
    off, frames = 3, []
    for _ in range(nframes):
        t, dc, dt, n = struct.unpack_from(">fddh", data, off)
        off += 22
        rec = np.frombuffer(data, _MEAS, count=n, offset=off)
        off += n * _MEAS.itemsize
        frames.append(RecordedFrame(t, dc, dt, rec["angle"].astype(np.float64), rec["distance"].astype(np.float64),
                                    (rec["hit"] != 0).astype(np.uint8)))
    return frames


def save(path, frames):
    with open(path, "wb") as f:
        f.write(dumps(frames))


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())
