"""Rank-count invariance of the sharded SLAM step (SURVEY.md §8e "invariants to test").

The reference's loop over particles is sequential and every particle is independent within a step
(SLAM.java:88-117); resampling walks the weights in particle order (SLAM.java:133-153).  Sharding the particle
set over R ranks must not change any of that: parent indices, f32 poses and integer map counts have to be
identical for R in {1, 2, 4, 8}.  This module runs the SAME seeded replay (device Philox keyed by the global
particle index, so the draws are rank-invariant by construction) on all ranks of an initialised process group
and on a single-rank handle on rank 0, and compares.  It only drives the product library (libgms.so): it is used
by `bench.py` (every multi-GPU run carries a `parity_check` object and fails loudly on a mismatch) and by
`tests/test_multigpu_nccl.py`.
"""
from __future__ import annotations

import hashlib

import numpy as np

from . import binding as B
from . import parallel, synth


def _digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _run(handle, stepper, scans, torch, dev, maps_of, check_steps):
    """Steps `handle` through `scans` (resampling every step, device-drawn noise) and records the state."""
    rec = []
    for s, sc in enumerate(scans):
        t_xy = torch.from_numpy(sc.beam_xy).to(dev)
        t_d = torch.from_numpy(sc.beam_dist).to(dev)
        t_h = torch.from_numpy(sc.beam_hit).to(dev)
        torch.cuda.synchronize(dev)
        args = (t_xy.data_ptr(), t_d.data_ptr(), t_h.data_ptr(), sc.num_beams, sc.d_center, sc.d_theta)
        if stepper is not None:
            stepper.step(*args, policy=B.POLICY_ALWAYS, u01=-1.0)
        else:
            handle.step_dev(*args, None, B.POLICY_ALWAYS, -1.0)
        neff = handle.read_neff()  # synchronises this rank's stream
        if stepper is not None:
            stepper.dist.barrier()  # ... and now every rank's: the getters below read other ranks' blocks through peer mappings
        strongest = handle.strongest()
        entry = {"neff": neff, "parents": handle.parents().copy(), "poses": handle.poses().copy(),
                 "weights": handle.weights().copy(), "strongest": (strongest[0], strongest[1].tobytes())}
        if s in check_steps:
            entry["maps"] = {p: _digest(handle.get_map(p, B.MAP_FREE_COUNT), handle.get_map(p, B.MAP_OCC_COUNT),
                                        handle.get_map(p, B.MAP_LIKELIHOOD)) for p in maps_of(handle)}
        rec.append(entry)
        if stepper is not None:
            stepper.dist.barrier()  # no rank starts the next step while another still reads through peer mappings
    return rec


def check(lib, dist, dev, rank, world, *, per_particle, P, beams, steps, grid_m, seed=4242, max_range=12.0):
    """Returns (ok, detail).  Collective: every rank of the group must call it with the same arguments."""
    import torch

    mode = B.MAP_PER_PARTICLE if per_particle else B.MAP_SHARED
    kw = dict(num_particles=P, map_width_m=grid_m, map_height_m=grid_m, origin_x=-grid_m / 2, origin_y=-grid_m / 2,
              map_mode=mode, resample_mode=B.RESAMPLE_FIXED, seed=seed)
    scans = synth.make_scans(steps, beams, max_range=max_range)
    check_steps = {steps // 2, steps - 1}
    stream = torch.cuda.Stream(dev)
    prev = torch.cuda.current_stream(dev)
    torch.cuda.set_stream(stream)
    try:
        h = lib.create(rank=rank, nranks=world, device=dev.index, **kw)
        h.set_stream(stream.cuda_stream)
        stepper = parallel.ShardedStepper(h, dist, dev)

        def local_maps(hh):
            if not per_particle:
                return (0,)
            return range(hh.info.local_begin, hh.info.local_begin + hh.info.local_count)

        mine = _run(h, stepper, scans, torch, dev, local_maps, check_steps)
        dist.barrier()
        h.close()
        ref = [None]
        if rank == 0:
            h1 = lib.create(rank=0, nranks=1, device=dev.index, **kw)
            h1.set_stream(stream.cuda_stream)
            ref[0] = _run(h1, None, scans, torch, dev, (lambda hh: range(P)) if per_particle else (lambda hh: (0,)),
                          check_steps)
            h1.close()
        dist.broadcast_object_list(ref, src=0)
        ref = ref[0]
    finally:
        torch.cuda.set_stream(prev)
    problems = []
    max_w_rel = 0.0
    for s, (a, b) in enumerate(zip(mine, ref)):
        if not np.array_equal(a["parents"], b["parents"]):
            problems.append(f"step {s}: parents differ in {int(np.sum(a['parents'] != b['parents']))} places")
        if a["poses"].tobytes() != b["poses"].tobytes():
            problems.append(f"step {s}: pose bytes differ")
        if a["strongest"] != b["strongest"]:
            problems.append(f"step {s}: strongest differs")
        rel = float(np.max(np.abs(a["weights"] - b["weights"]) / np.maximum(b["weights"], 1e-300)))
        max_w_rel = max(max_w_rel, rel)
        if rel > 1e-12 or abs(a["neff"] / b["neff"] - 1) > 1e-12:
            problems.append(f"step {s}: weights differ by {rel:.3e} (rel), neff {a['neff']} vs {b['neff']}")
        for p, dg in a.get("maps", {}).items():
            if dg != b["maps"][p]:
                problems.append(f"step {s}: map of particle {p} (counts / likelihood field) differs")
    ok = torch.tensor([0 if problems else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    detail = {"ranks": world, "particles": P, "beams": beams, "steps": steps, "grid": f"{int(grid_m / 0.05)}^2",
              "map_mode": "per_particle" if per_particle else "shared",
              "compared": "parents, f32 pose bytes, strongest, weights (rel 1e-12), Neff, per-cell counts + likelihood "
                          "field digests" + (" of every particle's map (maps migrate between ranks every step)"
                                             if per_particle else ""),
              "max_weight_rel_diff_rank0": max_w_rel, "problems_rank0": problems[:5]}
    return bool(ok.item() == 1), detail
