"""B200-native hot path of gridmap-slam-robot's SLAM backend (java/GridMapGL, com.fmsz.gridmapgl.slam).

csrc/   hand-written CUDA (sm_100a) kernels + the C-ABI library libgms.so (include/gms.h)
binding ctypes binding of the C-ABI
slam    host-side mirror of the reference's SLAM / GridMap / Observation / Odometry / Pose classes
synth   synthetic scans
"""
from . import binding  # noqa: F401

__all__ = ["binding"]
