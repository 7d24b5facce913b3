#!/usr/bin/env python
"""bench.py — particle-beam scores/s and ms per full SLAM step (score + resample + map update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload K4|K2|K3|K5] [--impl reference]

A "step" is one pass of the hot path over one synthetic scan: motion sampling, likelihood-field
refresh, scan scoring of every particle, weight normalisation / Neff / strongest, map integration and
systematic resampling (every step: the policy is GMS_RESAMPLE_ALWAYS so that no work is skipped).

Default workload at N=1 is K4, the configuration BASELINE.json's target is quoted on (100k particles x
720 beams, 4096^2 shared grid); it fits one GPU.  With N>1 ranks the particles are sharded
(100k per GPU, weak scaling; --scaling strong shards BASELINE's fixed 100k) and the only data-path
collective is the all-gather of the 24-byte {log-weight, pose} records.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (scans already in HBM,
CUDA-event timed, L2 flushed between steps); `e2e` = the same step through the host-buffer C-ABI calls
the Java shim would make (GridMapApp.java:178-192), H2D/D2H inside the timed region.
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


def score_bytes(P, B, s=8):
    """algorithmic bytes of one scoring launch = P*B*s + P*(12+8) + B*17."""
    return P * B * s + P * 20 + B * 17


def run_reference(args, wl):
    """The reference arm: the CPU restatement of the Java path (oracle/libgms_ref.so; the JVM reference
    cannot run in this image) on the host cores, same config / metric / unit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import build as b
    from gridmap_slam_robot_b200 import synth
    import ctypes

    lib = B.Library(b.build_oracle())
    cores = os.cpu_count() or 1
    # same configuration as the GPU arm at this N: weak scaling multiplies the particle count
    P = wl["P"] * max(1, args.gpus) if args.scaling == "weak" else wl["P"]
    # bounded sample: shared-map workloads run at full size for a few steps; per-particle-map workloads
    # (cost linear in particles: each owns a map) time a slice of the particles at full beam count/map size
    P_s = min(P, args.ref_particles if args.ref_particles else (P if wl["mode"] == "shared" else 64))
    h = make_handle(lib, wl, P_s)
    lib.dll.gmsref_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.dll.gmsref_set_threads(h.h, cores)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    scans = synth.make_scans(warm + steps, wl["B"], max_range=wl["max_range"])
    scored = 0
    t_total = 0.0
    for s, sc in enumerate(scans):
        t0 = time.perf_counter()
        h.update(sc.beam_xy, sc.beam_dist, sc.beam_hit, sc.d_center, sc.d_theta)
        h.resample(-1.0)
        h.weighted_pose()
        dt = time.perf_counter() - t0
        if s >= warm:
            t_total += dt
            scored += P_s * sc.num_hits
    value = scored / t_total
    sample = (f"{P_s} of {P} particles x {wl['B']} beams on the full {h.W}x{h.H} grid, {steps} steps after {warm} "
              f"warm-up; C restatement of the Java path (not JVM), OpenMP over particles")
    line = {
        "impl": "reference", "metric": "particle_beam_scores_per_s", "value": value, "unit": "scores/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_total / steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "particles_total": P, "map_mode": wl["mode"],
                   "resample": "every step"},
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
    from gridmap_slam_robot_b200 import synth

    lib = B.Library(b.build_oracle())
    lib.dll.gmsref_phase_seconds.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int32]
    P = wl["P"]
    P_s = min(P, P if wl["mode"] == "shared" or P <= 100 else 8)
    h = make_handle(lib, wl, P_s)
    scans = synth.make_scans(warm + steps, wl["B"], max_range=wl["max_range"])
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


def run_gpu(args, wl):
    import torch

    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import synth

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
    P_total = wl["P"] * world if args.scaling == "weak" else wl["P"]
    assert P_total % world == 0
    h = make_handle(lib, wl, P_total, rank=rank, nranks=world, device=local_rank)
    # a non-default torch stream shared with the library, so torch.cuda.Event brackets the library's
    # launches (the legacy default stream has handle 0, which gms_set_stream reads as "own stream")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h.set_stream(stream.cuda_stream)
    runner = None
    if world > 1:
        from gridmap_slam_robot_b200 import parallel

        runner = parallel.ShardedStepper(h, dist, dev)

    nscan = SCAN_RING
    scans = synth.make_scans(nscan, wl["B"], max_range=wl["max_range"])
    Bn = wl["B"]
    t_xy = torch.from_numpy(np.stack([s.beam_xy for s in scans])).to(dev)
    t_d = torch.from_numpy(np.stack([s.beam_dist for s in scans])).to(dev)
    t_h = torch.from_numpy(np.stack([s.beam_hit for s in scans])).to(dev)
    hits = [s.num_hits for s in scans]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(i):
        k = i % nscan
        a = (t_xy[k].data_ptr(), t_d[k].data_ptr(), t_h[k].data_ptr(), Bn, scans[k].d_center, scans[k].d_theta)
        if runner:
            runner.step(*a, policy=B.POLICY_ALWAYS)
        else:
            h.step_dev(*a, None, B.POLICY_ALWAYS, -1.0)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    h.profile_reset()
    h.profile_enable(True)
    launches0 = h.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scored = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # evict the likelihood field / particle arrays from L2 (not timed)
        ev[i][0].record(stream)
        step(args.warmup + i)
        ev[i][1].record(stream)
        scored += P_total * hits[(args.warmup + i) % nscan]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    phase_ms, phase_launches = h.profile_read()
    h.profile_enable(False)
    launches = h.launch_count() - launches0
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    if dist:
        t = torch.tensor([t_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    neff = h.read_neff()

    # ---- e2e: host buffers through the public C-ABI calls, copies inside the timed region (N=1) ----
    e2e = None
    if world == 1:
        h.set_stream(None)
        pinned = [(torch.from_numpy(s.beam_xy).pin_memory(), torch.from_numpy(s.beam_dist).pin_memory(),
                   torch.from_numpy(s.beam_hit).pin_memory()) for s in scans]
        n_e2e = max(3, min(args.steps, 200))

        def e2e_step(i):
            k = i % nscan
            xy, d, hh = pinned[k]
            h.update(xy.numpy(), d.numpy(), hh.numpy(), scans[k].d_center, scans[k].d_theta, None)
            h.resample(-1.0)
            h.strongest()
            return h.weighted_pose()

        for i in range(3):
            e2e_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc = 0
        for i in range(n_e2e):
            e2e_step(3 + i)
            sc += P_total * hits[(3 + i) % nscan]
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        stats_bytes = 104  # sizeof(Stats): neff, strongest index/pose/weight, weighted pose
        e2e = {"value": sc / te, "unit": "scores/s", "ms_per_step": 1e3 * te / n_e2e, "steps": n_e2e,
               "h2d_bytes_per_step": Bn * 25, "d2h_bytes_per_step": 2 * stats_bytes,
               "calls": "gms_update + gms_resample + gms_get_strongest + gms_get_weighted_pose "
                        "(GridMapApp.java:178-192), host beam arrays in pinned memory"}
    if world > 1:
        # e2e at N ranks: every rank stages the scan from pinned host memory each step (H2D inside the timed
        # region), steps through the all-gather, and reads Neff back (D2H + sync) like SLAM.update's caller
        pinned = [(torch.from_numpy(s.beam_xy).pin_memory(), torch.from_numpy(s.beam_dist).pin_memory(),
                   torch.from_numpy(s.beam_hit).pin_memory()) for s in scans]
        d_xy, d_d, d_h = torch.empty_like(t_xy[0]), torch.empty_like(t_d[0]), torch.empty_like(t_h[0])
        n_e2e = max(3, min(args.steps, 200))

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
        for i in range(n_e2e):
            e2e_step_mr(3 + i)
            sc += P_total * hits[(3 + i) % nscan]
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        e2e = {"value": sc / te, "unit": "scores/s", "ms_per_step": 1e3 * te / n_e2e, "steps": n_e2e,
               "h2d_bytes_per_step": Bn * 25 * world, "d2h_bytes_per_step": 104 * world,
               "calls": "per rank: H2D of the scan from pinned memory + gms_update_begin_dev + ncclAllGather + "
                        "gms_update_end_dev(resample) + gms_read_neff"}
    if rank != 0:
        return
    peak, peak_src = peaks()
    P_local = P_total // world
    n_score = max(1, phase_launches["score"] // 2)  # pack + score kernels share the phase counter
    score_ms = phase_ms["score"] / max(1, args.steps)
    sb = score_bytes(P_local, int(np.mean(hits)))
    achieved = sb / (score_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "score_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        except Exception:
            traffic = None
    value = scored / t_dev
    roofline = {"kernel": "k_score_sorted", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sb, "sector_bytes_per_launch": P_local * int(np.mean(hits)) * 32,
                "launch_ms": score_ms, "launches": n_score}
    if wl["mode"] == "per_particle":
        # per-particle maps: the scatter (k_map_update) dominates.  SURVEY.md §8d: C * 2 * s_log bytes with
        # s_log = 4 (one counter of the 8-byte pair is read and written per ray cell); C from the true poses
        # (n_r = 3 + |dfloor x| + |dfloor y| per ray, RayIterator.java:65-104), so approximate per particle.
        cells_per_scan = []
        for k, sc in enumerate(scans):
            x, y, th = synth.true_pose(k + 1)
            a = 2.0 * np.pi * np.arange(Bn) / Bn + th
            ex, ey = x + sc.beam_dist * np.cos(a), y + sc.beam_dist * np.sin(a)
            res = 0.05
            cells_per_scan.append(float(np.sum(3 + np.abs(np.floor(ex / res) - np.floor(x / res)) +
                                               np.abs(np.floor(ey / res) - np.floor(y / res)))))
        ub = P_local * float(np.mean(cells_per_scan)) * 8.0
        upd_ms = phase_ms["map_update"] / max(1, args.steps)
        roofline = {"kernel": "k_map_update", "bound": "hbm", "achieved": ub / (upd_ms * 1e-3) / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": ub / (upd_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ub, "launch_ms": upd_ms,
                    "note": "scatter of 64-bit counter atomics: bound by L2 atomic throughput, not by HBM bandwidth"}
    line = {
        "metric": "particle_beam_scores_per_s", "value": value, "unit": "scores/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "particles_total": P_total,
                   "particles_per_gpu": P_local, "beams": Bn, "beams_scored": float(np.mean(hits)),
                   "grid": f"{h.W}x{h.H}", "map_mode": wl["mode"], "resample": "every step (GMS_RESAMPLE_ALWAYS)",
                   "motion_noise": "device Philox", "map_update": args.update_mode, "l2": "flushed between timed steps (256 MiB memset, untimed)",
                   "scan_ring": nscan, "parallelism": f"particles sharded over {world} rank(s)"},
        "ms_per_step_wall": 1e3 * t_wall / args.steps,
        "phases_ms_per_step": {k: v / args.steps for k, v in phase_ms.items() if v > 0},
        "neff_last": neff,
        "roofline": roofline,
        "clocks": sampler.result(),
        "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
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
