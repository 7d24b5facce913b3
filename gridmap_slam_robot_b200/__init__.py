"""B200-native hot path of gridmap-slam-robot's SLAM backend (java/GridMapGL, com.fmsz.gridmapgl.slam).

csrc/      hand-written CUDA (sm_100a) kernels + the C-ABI library libgms.so (include/gms.h)
binding    ctypes binding of the C-ABI
slam       host-side mirror of the reference's SLAM / GridMap / Observation / Odometry / Pose classes
parallel   multi-rank stepping (one process per GPU)
recording  reader / writer of the reference's recording files
synth      synthetic scans
"""
from . import binding  # noqa: F401

__all__ = ["binding", "slam", "parallel", "recording", "synth"]


def __getattr__(name):  # the other modules load on first use (parallel pulls in torch)
    if name in ("slam", "parallel", "recording", "synth", "build"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
