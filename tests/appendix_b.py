"""Hand-derived known-answer vectors of SURVEY.md Appendix B (NOT from the reference's tests — it has none).

B1: GridMap.applyMeasurement (GridMap.java:194-228) + RayIterator (RayIterator.java:65-130) on a 10x10
grid; inputs are applyMeasurement's grid-coordinate floats, measuredDistance = (float)hypot(end-start),
2 extra steps.  Classes: F = += L_free, O = += L_occ, 0 = += 0.
"""
INF = float("inf")
NAN = float("nan")

RAYS = [
    # id, start, end, n, err0, cells, classes(hit), classes(miss)
    ("K1", (1, 1), (4, 2), 7, -1.0, [(1, 1), (2, 1), (3, 1), (3, 2), (4, 2), (5, 2), (6, 2)], "FFOOO00", "FFFF000"),
    ("K2", (4, 2), (1, 1), 7, -1.0, [(4, 2), (3, 2), (2, 2), (2, 1), (1, 1), (0, 1)], "FFFFOO", "FFFFF0"),
    ("K3", (2, 1), (2, 4), 6, INF, [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (2, 6)], "FFOO00", "FFF000"),
    ("K4", (5, 3), (2, 3), 6, -INF, [(5, 3), (4, 3), (3, 3), (2, 3), (1, 3), (0, 3)], "FFFOO0", "FFFF00"),
    ("K5", (3, 3), (3, 3), 3, NAN, [(3, 3), (3, 3), (3, 3)], "OOO", "000"),
    ("K6", (7, 7), (12, 9), 10, -1.5, [(7, 7), (8, 7), (8, 8), (9, 8)], "FFFF", "FFFF"),
    ("K7", (-2, 1), (3, 1), 8, -INF, [], "", ""),
    ("K8", (1, 1), (3, 3), 7, 0.0, [(1, 1), (2, 1), (2, 2), (3, 2), (3, 3), (4, 3), (4, 4)], "FFOOO00", "FFF0000"),
    ("K9", (1.25, 2.75), (5.6, 0.3), 11, -0.47499996423721313,
     [(1, 3), (2, 3), (2, 2), (3, 2), (3, 1), (4, 1), (5, 1), (5, 0), (6, 0), (7, 0)], "FFFFFFOOO0", "FFFFFFFF00"),
]

# B2: SLAM.resample (SLAM.java:133-153): weights, u = Math.random() draw, parent indices
RESAMPLE = [
    ([0.1, 0.2, 0.3, 0.4], 0.5, [1, 2, 3, 3]),
    ([0.1, 0.2, 0.3, 0.4], 0.2, [0, 1, 2, 3]),
    ([0.25, 0.25, 0.25, 0.25], 0.0, [0, 0, 1, 2]),
    ([0.25, 0.25, 0.25, 0.25], 0.5, [0, 1, 2, 3]),
    ([0.0, 0.0, 1.0, 0.0], 0.3, [2, 2, 2, 2]),
]
# last row of B2: Java throws IndexOutOfBounds; the build clamps to N-1
RESAMPLE_CLAMP = ([0.1, 0.2, 0.3, 0.3999999], 0.9999999)
NEFF = ([0.1, 0.2, 0.3, 0.4], 3.333333333333333)

# Appendix A constants
L_FREE = float.fromhex("-0x1.b1d104890c701p-1")
L_OCC = float.fromhex("0x1.193ea571eca66p+1")
KERNEL_HALF = [float.fromhex(s) for s in
               ("0x1.228633cc7fbd4p-8", "0x1.ba69e9bf23990p-5", "0x1.efb0b0ccafee0p-3", "0x1.98a0a325232ddp-2")]
GRID_SIZES = [(6.0, 120), (20.0, 400), (51.2, 1024), (102.4, 2048), (204.8, 4096)]
