#!/usr/bin/env python
"""bench.py — particle-beam scores/s and ms per full SLAM step (score + resample + map update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload K4|K2|K3|K5] [--impl reference]

A "step" is one pass of the hot path over one synthetic scan: motion sampling, likelihood-field
refresh, scan scoring of every particle, weight normalisation / Neff / strongest, map integration and
systematic resampling (every step: the policy is GMS_RESAMPLE_ALWAYS so that no work is skipped).

Default workload at N=1 is K4, the configuration BASELINE.json's target is quoted on (100k particles x
720 beams, 4096^2 shared grid); it fits one GPU.  With N>1 ranks the particles are sharded
(100k per GPU, weak scaling; --scaling strong shards BASELINE's fixed 100k) and the only data-path
exchange is that of the f64 log-weights: every rank's normalise kernel reads the other ranks' blocks over NVLink
(peer mappings).

One JSON line on stdout (rank 0):
  value        device-resident throughput: scans already in HBM, one CUDA-event pair per step on the stream the
               library launches on, closed AFTER the library's side streams have been joined (so the map
               integration that runs next to the resampling is inside the window), L2 flushed between steps
  back_to_back the same steps enqueued without flushes or gaps, one event pair around all of them (steady state:
               the likelihood refresh of step t+1 overlaps the tail of step t, as in production)
  e2e          the same step through the host-buffer C-ABI calls the Java shim would make
               (GridMapApp.java:178-192), H2D/D2H inside the timed region, driven from Python (ctypes binding)
  e2e_native   the same calls from a compiled host loop (csrc/e2e_host.cpp, a plain C-ABI client): what the boundary
               costs without an interpreter between the calls
  parity_check (N > 1) rank-count invariance, checked before anything is timed: a seeded replay on N ranks equals
               the same replay on one rank (parents, pose bytes, weights, every map); the run aborts on a mismatch
  extra        (default workload only) short timed runs of BASELINE config 4 as written (100k particles in total,
               sharded: strong scaling) and config 5 (per-particle 1024^2 maps, 1k per GPU, maps migrating)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: particles, beams, grid metres (0.05 m cells), map mode, scan max range
    "K1": dict(P=100, B=360, grid_m=20.0, mode="per_particle", max_range=30.0,
               desc="100 particles x 360 beams, 400^2 per-particle maps (the reference's own mode and size class)"),
    "K2": dict(P=1000, B=360, grid_m=51.2, mode="shared", max_range=30.0,
               desc="1k particles x 360 beams, 1024^2 shared grid"),
    "K2pp": dict(P=1000, B=360, grid_m=51.2, mode="per_particle", max_range=30.0,
                 desc="1k particles x 360 beams, 1024^2 per-particle maps (reference semantics)"),
    "K3": dict(P=10000, B=720, grid_m=102.4, mode="shared", max_range=12.0,
               desc="10k particles x 720 beams, 2048^2 shared grid, 12 m range"),
    "K4": dict(P=100000, B=720, grid_m=204.8, mode="shared", max_range=30.0,
               desc="100k particles x 720 beams, 4096^2 shared grid"),
    "K5": dict(P=1000, B=360, grid_m=51.2, mode="per_particle", max_range=30.0,
               desc="FastSLAM-style: 1k particles per GPU, each with a 1024^2 map"),
    # gather stress (north_star's regime for the scoring kernel): the particles are spread over the whole room and
    # the room is scaled up, so the factor-field footprint the lookups touch exceeds L1 and a warp's 32 particles
    # no longer share lines (see --spread)
    "K4g": dict(P=100000, B=720, grid_m=204.8, mode="shared", max_range=150.0, room_scale=5.0, spread=True,
                desc="gather stress: 100k particles spread over a 90 m room x 720 beams, 4096^2 shared grid"),
}
SCAN_RING = 16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = [], set(), False, None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


UPDATE_MODE = 0  # --update-mode sorted: the atomic-free sort + run-length scatter (measured alternative)


def make_handle(lib, wl, P, rank=0, nranks=1, device=0):
    from gridmap_slam_robot_b200 import binding as B

    g = wl["grid_m"]
    return lib.create(num_particles=P, map_width_m=g, map_height_m=g, origin_x=-g / 2, origin_y=-g / 2,
                      map_mode=B.MAP_SHARED if wl["mode"] == "shared" else B.MAP_PER_PARTICLE,
                      rank=rank, nranks=nranks, device=device, seed=20260101, update_mode=UPDATE_MODE)


def make_scans(wl, n):
    from gridmap_slam_robot_b200 import synth

    return synth.make_scans(n, wl["B"], max_range=wl["max_range"], room_scale=wl.get("room_scale", 1.0))


def config_of(name, wl, P_total, world, W=None, H=None, hits=None, update_mode="atomic"):
    """The `config` object: identical (as a dict) for the GPU arm and the reference arm of one (workload, N, scaling)."""
    cells = int(round(wl["grid_m"] / 0.05))
    ring = make_scans(wl, SCAN_RING)
    return {"workload": f"{name}: {wl['desc']}", "particles_total": P_total, "particles_per_gpu": P_total // world,
            "beams": wl["B"], "grid": f"{W or cells}x{H or cells}", "map_mode": wl["mode"],
            "resample": "every step (GMS_RESAMPLE_ALWAYS)", "motion_noise": "device Philox",
            "map_update": update_mode, "scan_ring": SCAN_RING,
            "parallelism": f"particles sharded over {world} rank(s)",
            "beams_scored": float(np.mean([s.num_hits for s in ring])),
            "l2": "GPU arm: flushed between timed steps (256 MiB memset, untimed), see back_to_back for the unflushed "
                  "figure; reference arm: host caches left as they are",
            "timed_window": "GPU arm: per step a CUDA event pair, closed after the library's side streams have been "
                            "joined; reference arm: wall clock around update + resample + weighted pose of each step"}


def score_bytes(P, B, s=8):
    """algorithmic bytes of one scoring launch = P*B*s + P*(12+8) + B*17."""
    return P * B * s + P * 20 + B * 17


def run_reference(args, wl):
    """The reference arm: the CPU restatement of the Java path (oracle/libgms_ref.so; the JVM reference
    cannot run in this image) on all host cores, same config / metric / unit / steps / warm-up."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import ctypes

    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import build as b

    lib = B.Library(b.build_oracle())
    cores = os.cpu_count() or 1
    world = max(1, args.gpus)
    # same configuration as the GPU arm at this N: weak scaling multiplies the particle count
    P = wl["P"] * world if args.scaling == "weak" else wl["P"]
    # bounded sample: shared-map workloads run at full size; per-particle-map workloads (cost linear in particles:
    # each owns a map) time a slice of the particles at full beam count / map size
    P_s = min(P, args.ref_particles if args.ref_particles else (P if wl["mode"] == "shared" else 64))
    h = make_handle(lib, wl, P_s)
    lib.dll.gmsref_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.dll.gmsref_set_threads(h.h, cores)
    steps, warm = args.steps, args.warmup
    ring = make_scans(wl, SCAN_RING)
    scored, t_total, done = 0, 0.0, 0
    t_start = time.perf_counter()
    for s in range(warm + steps):
        sc = ring[s % SCAN_RING]
        t0 = time.perf_counter()
        h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta)
        h.resample(-1.0)
        h.weighted_pose()
        dt = time.perf_counter() - t0
        if s >= warm:
            t_total += dt
            scored += P_s * sc.num_hits
            done += 1
        if time.perf_counter() - t_start > 240 and done >= 3:  # keep the whole run within a few minutes
            break
    value = scored / t_total
    sample = (f"{P_s} of {P} particles x {wl['B']} beams on the full {h.W}x{h.H} grid, {done} timed steps after {warm} "
              f"warm-up; C restatement of the Java path (not JVM), OpenMP over particles on {cores} threads")
    line = {
        "impl": "reference", "metric": "particle_beam_scores_per_s", "value": value, "unit": "scores/s",
        "n_gpus": args.gpus, "steps": done, "warmup": warm, "ms_per_step": 1e3 * t_total / done,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, wl, P, world, h.W, h.H, update_mode=args.update_mode),
        "cpu_baseline": {"value": value, "unit": "scores/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


PHASES = ("motion", "likelihood_field", "scoring", "map_update", "normalise", "resample")


def cpu_baseline(wl, workload_name, steps=2, warm=1):
    """Oracle on ONE host core (the reference is single threaded: SLAM.java:88 on the GL render thread),
    a bounded sample of the same workload, with the per-phase split SURVEY.md §8d asks for."""
    import ctypes

    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import build as b

    lib = B.Library(b.build_oracle())
    lib.dll.gmsref_phase_seconds.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int32]
    P = wl["P"]
    P_s = min(P, P if wl["mode"] == "shared" or P <= 100 else 8)
    h = make_handle(lib, wl, P_s)
    scans = make_scans(wl, warm + steps)
    t_total, scored, per_step = 0.0, 0, []
    ph = (ctypes.c_double * 6)()
    for s, sc in enumerate(scans):
        if s == warm:
            lib.dll.gmsref_phase_seconds(h.h, ph, 1)
        t0 = time.perf_counter()
        h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta)
        h.resample(-1.0)
        h.weighted_pose()
        dt = time.perf_counter() - t0
        if s >= warm:
            t_total += dt
            per_step.append(dt)
            scored += P_s * sc.num_hits
    lib.dll.gmsref_phase_seconds(h.h, ph, 0)
    h.close()
    per_step.sort()
    return {"value": scored / t_total, "unit": "scores/s", "cores": 1, "kind": "port",
            "ms_per_step_sample": 1e3 * t_total / steps,
            "ms_per_step_median": 1e3 * per_step[len(per_step) // 2],
            "ms_per_step_p95": 1e3 * per_step[min(len(per_step) - 1, int(0.95 * len(per_step)))],
            "phase_ms_per_step": {k: 1e3 * ph[i] / steps for i, k in enumerate(PHASES)},
            "host": {"nproc": os.cpu_count(), "cpu": cpu_model()},
            "sample": f"{P_s} of {P} particles x {wl['B']} beams, full {wl['grid_m']} m grid, {steps} steps after {warm} "
                      f"warm-up; C restatement of the Java path (not JVM), single thread"}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def native_e2e(lib_path, wl, P_total, scans, steps, warm=5):
    """The same end-to-end step from a compiled host: csrc/e2e_host (C++) drives gms_update + gms_resample +
    gms_get_strongest + gms_get_weighted_pose through the C-ABI with host beam arrays — what the Java shim's FFM
    downcalls cost, without the Python interpreter between the calls.  Runs after this process has released its
    handle; returns None if the helper is not built."""
    import struct
    import subprocess
    import tempfile

    from gridmap_slam_robot_b200 import build as b

    try:
        exe = b.build_e2e_host()
    except Exception:
        return None
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(struct.pack("<ii", len(scans), wl["B"]))
        for sc in scans:
            f.write(struct.pack("<dd", sc.d_center, sc.d_theta))
            f.write(np.ascontiguousarray(sc.beam_xy, np.float64).tobytes())
            f.write(np.ascontiguousarray(sc.beam_dist, np.float64).tobytes())
            f.write(np.ascontiguousarray(sc.beam_hit, np.uint8).tobytes())
        path = f.name
    try:
        out = subprocess.run([exe, lib_path, path, str(P_total), str(wl["grid_m"]), "1" if wl["mode"] == "shared" else "0",
                              str(steps), str(warm)], capture_output=True, text=True, timeout=300)
        if out.returncode != 0:
            return {"error": (out.stderr or out.stdout)[-300:]}
        r = json.loads(out.stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)
    return {"value": r["value"], "unit": "scores/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
            "h2d_bytes_per_step": wl["B"] * 25, "d2h_bytes_per_step": 208,
            "host": "compiled C++ loop over the C-ABI (gridmap_slam_robot_b200/csrc/e2e_host.cpp): gms_update + "
                    "gms_resample + gms_get_strongest + gms_get_weighted_pose, host beam arrays"}


def kernel_counters(workload):
    """ncu counters of the workload's dominant kernel (profiles/kernel_counters.json, written from this round's
    `ncu --set full` capture): DRAM bytes and executed warp instructions per launch."""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        with open(p) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def time_workload(ctx, name, wl, P_total, steps, warmup, e2e_steps):
    """One timed run of `wl` with P_total particles over all ranks of ctx; returns the measurements."""
    import torch

    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import synth

    lib, dist, dev, rank, world, stream = ctx["lib"], ctx["dist"], ctx["dev"], ctx["rank"], ctx["world"], ctx["stream"]
    assert P_total % world == 0
    h = make_handle(lib, wl, P_total, rank=rank, nranks=world, device=dev.index)
    h.set_stream(stream.cuda_stream)
    runner = None
    if world > 1:
        from gridmap_slam_robot_b200 import parallel

        runner = parallel.ShardedStepper(h, dist, dev)
    nscan = SCAN_RING
    scans = make_scans(wl, nscan)
    Bn = wl["B"]
    t_xy = torch.from_numpy(np.stack([s.beam_xy for s in scans])).to(dev)
    t_d = torch.from_numpy(np.stack([s.beam_dist for s in scans])).to(dev)
    t_h = torch.from_numpy(np.stack([s.beam_hit for s in scans])).to(dev)
    hits = [s.num_hits for s in scans]
    flush = ctx["flush"]

    def step(i, policy=B.POLICY_ALWAYS):
        k = i % nscan
        a = (t_xy[k].data_ptr(), t_d[k].data_ptr(), t_h[k].data_ptr(), Bn, scans[k].d_center, scans[k].d_theta)
        if runner:
            runner.step(*a, policy=policy)
        else:
            h.step_dev(*a, None, policy, -1.0)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    respread = None
    if wl.get("spread"):
        # gather stress: the particles are scattered over the whole room with random headings before every step
        # (gms_set_poses, untimed), so neighbouring lanes no longer look up neighbouring cells; the resampling of
        # the step collapses the cloud again, hence the re-spread
        half = 9.0 * wl.get("room_scale", 1.0) - 1.0
        rng = np.random.default_rng(5)
        xyt = np.stack([rng.uniform(-half, half, P_total), rng.uniform(-half, half, P_total),
                        rng.uniform(-np.pi, np.pi, P_total)], 1).astype(np.float32)

        def respread():
            h.set_poses(xyt)
    for i in range(warmup):
        if respread:
            respread()
        step(i)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    h.profile_reset()
    h.profile_enable(True)
    launches0 = h.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    scored = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if respread:
            respread()
        flush.zero_()  # evict the likelihood field / particle arrays from L2 (not timed)
        ev[i][0].record(stream)
        step(warmup + i)
        h.join_streams()  # the map integration on the library's side stream belongs to this step's window
        ev[i][1].record(stream)
        scored += P_total * hits[(warmup + i) % nscan]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    phase_ms, phase_launches = h.profile_read()
    h.profile_enable(False)
    launches = h.launch_count() - launches0
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    # steady state: the same steps back to back, no flush, no join between steps, one event pair around all
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(steps):
        step(warmup + steps + i)
    h.join_streams()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    t_b2b = e0.elapsed_time(e1) * 1e-3
    scored_b2b = sum(P_total * hits[(warmup + steps + i) % nscan] for i in range(steps))
    t_never = None
    if wl["mode"] == "per_particle":
        # the other regime of the per-particle path: no resampling, so EVERY particle's map takes the scan (with a
        # resampling only the selected parents' maps do: the rest is dropped); same steps, back to back
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step(warmup + 2 * steps - 1, B.POLICY_NEVER)  # leaves its scan pending, like every step of the loop below
        barrier()
        h.profile_reset()
        h.profile_enable(True)
        e2.record(stream)
        for i in range(steps):  # each step integrates its predecessor's scan into every map, then scores its own
            step(warmup + 2 * steps + i, B.POLICY_NEVER)
        e3.record(stream)
        barrier()
        phase_never, _ = h.profile_read()
        h.profile_enable(False)
        t_never = e2.elapsed_time(e3) * 1e-3 / steps
    if dist:
        t = torch.tensor([t_dev, t_b2b, t_never or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_b2b = float(t[0].item()), float(t[1].item())
        t_never = float(t[2].item()) if t_never is not None else None
    neff = h.read_neff()

    # ---- e2e: host buffers, copies inside the timed region ----
    e2e = None
    pinned = [(torch.from_numpy(s.beam_xy).pin_memory(), torch.from_numpy(s.beam_dist).pin_memory(),
               torch.from_numpy(s.beam_hit).pin_memory()) for s in scans]
    if e2e_steps and world == 1:
        h.set_stream(None)

        host = [(xy.numpy(), d.numpy(), hh.numpy()) for xy, d, hh in pinned]  # the caller's arrays (pinned memory)

        def e2e_step(i):
            k = i % nscan
            xy, d, hh = host[k]
            h.update(xy, d, hh, scans[k].d_center, scans[k].d_theta, None)
            h.resample(-1.0)
            h.strongest()
            return h.weighted_pose()

        for i in range(3):
            e2e_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc = 0
        for i in range(e2e_steps):
            e2e_step(3 + i)
            sc += P_total * hits[(3 + i) % nscan]
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        stats_bytes = 104  # sizeof(Stats): neff, strongest index/pose/weight, weighted pose
        e2e = {"value": sc / te, "unit": "scores/s", "ms_per_step": 1e3 * te / e2e_steps, "steps": e2e_steps,
               "h2d_bytes_per_step": Bn * 25, "d2h_bytes_per_step": 2 * stats_bytes,
               "calls": "gms_update + gms_resample + gms_get_strongest + gms_get_weighted_pose "
                        "(GridMapApp.java:178-192), host beam arrays in pinned memory"}
        h.set_stream(stream.cuda_stream)
    elif e2e_steps:
        # e2e at N ranks: every rank stages the scan from pinned host memory each step (H2D inside the timed
        # region), steps through the peer exchange, and reads Neff back (D2H + sync) like SLAM.update's caller
        d_xy, d_d, d_h = torch.empty_like(t_xy[0]), torch.empty_like(t_d[0]), torch.empty_like(t_h[0])

        def e2e_step_mr(i):
            k = i % nscan
            d_xy.copy_(pinned[k][0], non_blocking=True)
            d_d.copy_(pinned[k][1], non_blocking=True)
            d_h.copy_(pinned[k][2], non_blocking=True)
            runner.step(d_xy.data_ptr(), d_d.data_ptr(), d_h.data_ptr(), Bn, scans[k].d_center, scans[k].d_theta,
                        policy=B.POLICY_ALWAYS)
            return h.read_neff()

        for i in range(3):
            e2e_step_mr(i)
        barrier()
        t0 = time.perf_counter()
        sc = 0
        for i in range(e2e_steps):
            e2e_step_mr(3 + i)
            sc += P_total * hits[(3 + i) % nscan]
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        e2e = {"value": sc / te, "unit": "scores/s", "ms_per_step": 1e3 * te / e2e_steps, "steps": e2e_steps,
               "h2d_bytes_per_step": Bn * 25 * world, "d2h_bytes_per_step": 104 * world,
               "calls": "per rank: H2D of the scan from pinned memory + gms_update_begin_dev (scoring into the rank's "
                        "exchange buffer) + gms_update_end_dev (normalise pulls every rank's log-weights over NVLink; "
                        "resample) + gms_read_neff"}
    W, H = h.W, h.H
    barrier()
    h.close()
    return dict(name=name, wl=wl, P_total=P_total, P_local=P_total // world, W=W, H=H, steps=steps, warmup=warmup,
                t_dev=t_dev, t_b2b=t_b2b, t_never=t_never, phase_never=phase_never if t_never is not None else None,
                scored=scored, scored_b2b=scored_b2b, t_wall=t_wall, phase_ms=phase_ms,
                phase_launches=phase_launches, launches=int(launches), neff=neff, e2e=e2e, clocks=sampler.result(),
                hits=float(np.mean(hits)), scans=scans)


def roofline_of(r, args):
    """The roofline object of the run's dominant kernel: ALGORITHMIC bytes per launch
    / average launch duration (CUDA events around the phase, measured in this run) / measured HBM peak; beside it
    what ncu says actually bounds the kernel (profiles/README.md) and the issue-slot utilisation."""
    from gridmap_slam_robot_b200 import synth

    peak, peak_src = peaks()
    wl, steps = r["wl"], r["steps"]
    kc = kernel_counters(r["name"]) or {}
    sm_hz = (r["clocks"].get("sm_mhz") or 1965.0) * 1e6
    if wl["mode"] == "shared":
        score_ms = r["phase_ms"]["score"] / max(1, steps)
        sb = score_bytes(r["P_local"], int(r["hits"]))
        achieved = sb / (score_ms * 1e-3) / 1e9
        inst = kc.get("warp_instructions_per_launch")
        out = {"kernel": "k_score_sorted", "bound": kc.get("bound", "issue"), "achieved": achieved, "peak": peak,
               "unit": "GB/s", "frac": achieved / peak, "traffic": kc.get("dram_bytes_per_launch"),
               "traffic_source": kc.get("source"), "peak_source": peak_src, "algorithmic_bytes_per_launch": sb,
               "sector_bytes_per_launch": r["P_local"] * int(r["hits"]) * 32, "launch_ms": score_ms,
               "launches": steps,
               "note": "frac = SURVEY 8(d) algorithmic bytes / launch time / measured HBM copy peak.  ncu: the gathers "
                       "of this workload hit L1/L2 (DRAM traffic = `traffic`), the kernel is bound by instruction "
                       "issue; issue_frac = executed warp instructions / (SMs x 4 schedulers x SM clock x launch time)"}
        if inst:
            out["warp_instructions_per_launch"] = inst
            out["issue_frac"] = inst / (148 * 4 * sm_hz * score_ms * 1e-3)
        if kc.get("l2_sectors_per_launch"):
            # what the gather really moves: 32-byte sectors from L2 (and HBM behind it), from the ncu capture
            out["l2_sector_gbs"] = kc["l2_sectors_per_launch"] * 32 / (score_ms * 1e-3) / 1e9
            out["l2_hit_rate_pct"] = kc.get("l2_hit_rate_pct")
            out["lts_throughput_pct_of_peak_ncu"] = kc.get("lts_throughput_pct_of_peak")
            if kc.get("dram_bytes_per_launch"):
                out["dram_gbs"] = kc["dram_bytes_per_launch"] / (score_ms * 1e-3) / 1e9
                out["dram_frac_of_peak"] = out["dram_gbs"] / peak
        return out
    # per-particle maps: the scatter (map update) dominates.  SURVEY.md §8d: C * 2 * s_log bytes with s_log = 4 (one
    # counter of the 8-byte pair is read and written per ray cell); C from the true poses (n_r = 3 + |dfloor x| +
    # |dfloor y| per ray, RayIterator.java:65-104), so approximate per particle.
    Bn = wl["B"]
    cells_per_scan = []
    for k, sc in enumerate(r["scans"]):
        x, y, th = synth.true_pose(k + 1)
        a = 2.0 * np.pi * np.arange(Bn) / Bn + th
        ex, ey = x + sc.beam_dist * np.cos(a), y + sc.beam_dist * np.sin(a)
        res = 0.05
        cells_per_scan.append(float(np.sum(3 + np.abs(np.floor(ex / res) - np.floor(x / res)) +
                                           np.abs(np.floor(ey / res) - np.floor(y / res)))))
    ub = r["P_local"] * float(np.mean(cells_per_scan)) * 8.0
    # measured in the no-resample regime, where every particle's map takes the scan (with a resampling only the
    # selected parents' maps do, and the step is dominated by the streaming map copies instead)
    upd_ms = (r.get("phase_never") or r["phase_ms"])["map_update"] / max(1, steps)
    copy = None
    if kc.get("copy") and r["phase_ms"].get("map_copy"):
        # the resampling regime is dominated by the streaming map copies: DRAM bytes of the launch (ncu: the child
        # maps written once; the parents' reads hit L2) / the event-timed copy phase of this run
        cp_ms = r["phase_ms"]["map_copy"] / max(1, steps)
        cb = kc["copy"]["dram_bytes_per_launch"]
        copy = {"kernel": kc["copy"]["kernel"], "bound": "hbm", "achieved": cb / (cp_ms * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": cb / (cp_ms * 1e-3) / 1e9 / peak, "traffic": cb,
                "traffic_source": kc["copy"]["source"], "launch_ms": cp_ms, "regime": "resampling every step"}
    return {"kernel": "k_map_update_red", "bound": kc.get("bound", "l2_atomic"), "achieved": ub / (upd_ms * 1e-3) / 1e9,
            "copy": copy,
            "peak": peak, "unit": "GB/s", "frac": ub / (upd_ms * 1e-3) / 1e9 / peak,
            "traffic": kc.get("dram_bytes_per_launch"), "traffic_source": kc.get("source"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ub, "launch_ms": upd_ms,
            "regime": "no_resample (every map integrates the scan)",
            "note": "scatter of counter increments: see DESIGN.md §3 for what bounds it"}


def summarise(r, args, world, scaling):
    """The compact object of one timed run (used for `extra.*`)."""
    out = {"workload": f"{r['name']}: {r['wl']['desc']}", "particles_total": r["P_total"], "scaling": scaling,
           "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": 1e3 * r["t_dev"] / r["steps"],
           "value": r["scored"] / r["t_dev"], "unit": "scores/s",
           "back_to_back": {"ms_per_step": 1e3 * r["t_b2b"] / r["steps"], "value": r["scored_b2b"] / r["t_b2b"]},
           "phases_ms_per_step": {k: v / r["steps"] for k, v in r["phase_ms"].items() if v > 0},
           "gpu_launches": r["launches"], "clocks": r["clocks"], "neff_last": r["neff"]}
    if r["e2e"]:
        out["e2e"] = r["e2e"]
    if r.get("t_never") is not None:
        out["no_resample"] = no_resample_of(r)
    return out


def no_resample_of(r):
    return {"ms_per_step": 1e3 * r["t_never"],
            "phases_ms_per_step": {k: v / r["steps"] for k, v in (r.get("phase_never") or {}).items() if v > 0},
            "note": "same scans with GMS_RESAMPLE_NEVER, back to back: every particle's map integrates the scan "
                    "(with a resampling only the selected parents' maps do, the others are dropped with their particles)"}


def run_gpu(args, wl):
    import torch

    from gridmap_slam_robot_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    lib = B.load()
    # a non-default torch stream shared with the library, so torch.cuda.Event brackets the library's
    # launches (the legacy default stream has handle 0, which gms_set_stream reads as "own stream")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = dict(lib=lib, dist=dist, dev=dev, rank=rank, world=world, stream=stream,
               flush=torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev))  # > 126 MB L2

    # ---- rank-count invariance, before anything is timed (N > 1) ----
    parity = None
    if world > 1 and not args.no_parity:
        from gridmap_slam_robot_b200 import rankcheck

        ok_s, det_s = rankcheck.check(lib, dist, dev, rank, world, per_particle=False, P=16384, beams=360, steps=6,
                                      grid_m=51.2)
        ok_p, det_p = rankcheck.check(lib, dist, dev, rank, world, per_particle=True, P=64 * world, beams=180, steps=5,
                                      grid_m=20.0)
        parity = {"ranks": world, "shared": ok_s, "per_particle": ok_p, "shared_detail": det_s,
                  "per_particle_detail": det_p}
        if not (ok_s and ok_p):
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "rank-count invariance violated"}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    P_total = wl["P"] * world if args.scaling == "weak" else wl["P"]
    e2e_steps = max(3, min(args.steps, 200))
    r = time_workload(ctx, args.workload, wl, P_total, args.steps, args.warmup, e2e_steps)
    extra = {}
    if args.workload == "K4" and args.scaling == "weak" and not args.no_extra:
        xs, xw = max(3, min(args.steps, 10)), 3
        if world > 1:  # BASELINE config 4 as written: 100k particles in total, sharded over the ranks
            extra["k4_strong"] = summarise(time_workload(ctx, "K4", wl, wl["P"], xs, xw, xs), args, world, "strong")
        k5 = WORKLOADS["K5"]  # BASELINE config 5: per-particle 1024^2 maps, 1k per GPU, P2P map copy on resample
        extra["k5"] = summarise(time_workload(ctx, "K5", k5, k5["P"] * world, xs, xw, xs), args, world, "weak")
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    line = {
        "metric": "particle_beam_scores_per_s", "value": r["scored"] / r["t_dev"], "unit": "scores/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["t_dev"] / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args.workload, wl, P_total, world, r["W"], r["H"], update_mode=args.update_mode),
        "back_to_back": {"ms_per_step": 1e3 * r["t_b2b"] / args.steps, "value": r["scored_b2b"] / r["t_b2b"],
                         "note": "same steps, no L2 flush, no gaps: one event pair around all of them"},
        "ms_per_step_wall": 1e3 * r["t_wall"] / args.steps,
        "phases_ms_per_step": {k: v / args.steps for k, v in r["phase_ms"].items() if v > 0},
        "neff_last": r["neff"],
        "roofline": roofline_of(r, args),
        "clocks": r["clocks"],
        "gpu_launches": r["launches"],
    }
    if r["e2e"]:
        line["e2e"] = r["e2e"]
    if world == 1 and r["e2e"]:
        ne = native_e2e(lib.path, wl, P_total, r["scans"], max(3, min(args.steps, 200)))
        if ne:
            line["e2e_native"] = ne
    if r.get("t_never") is not None:
        line["no_resample"] = no_resample_of(r)
    if parity:
        line["parity_check"] = parity
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl, args.workload)
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="K4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-particles", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.k4_strong / extra.k5 runs")
    ap.add_argument("--no-parity", action="store_true", help="skip the rank-count invariance check (N > 1)")
    ap.add_argument("--update-mode", default="atomic", choices=["atomic", "sorted"])
    ap.add_argument("--cpu-replay", type=int, default=0, metavar="STEPS",
                    help="only time the single-thread CPU restatement for STEPS steps of the workload (per-phase "
                         "split; SURVEY.md 8d asks for K1 x 500) and print its cpu_baseline object")
    args = ap.parse_args()
    if args.cpu_replay:
        wl = WORKLOADS[args.workload]
        print(json.dumps({"workload": f"{args.workload}: {wl['desc']}",
                          "cpu_baseline": cpu_baseline(wl, args.workload, steps=args.cpu_replay, warm=1)}), flush=True)
        return
    global UPDATE_MODE
    UPDATE_MODE = 1 if args.update_mode == "sorted" else 0
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == "__main__":
    main()
