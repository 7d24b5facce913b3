"""Debug (2 GPUs): first divergence between a 2-rank and a 1-rank per-particle run.  torchrun --nproc-per-node 2."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gridmap_slam_robot_b200 import binding as B, parallel, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
lib = B.load()
for (P, inject, u_dev) in ((128, False, True), (128, True, False), (128, True, True), (128, False, False), (48, False, True)):
    kw = dict(num_particles=P, map_width_m=20.0, map_height_m=20.0, origin_x=-10.0, origin_y=-10.0, map_mode=0, resample_mode=2, seed=4242)
    steps = 4
    scans = synth.make_scans(steps, 180, max_range=12.0)
    normals, uniforms = synth.make_draws(steps, P)
    def run(h, stepper, lo, cnt):
        out = []
        for s, sc in enumerate(scans):
            t = [torch.from_numpy(a).to(dev) for a in (sc.beam_xy, sc.beam_dist, sc.beam_hit)]
            tn = torch.from_numpy(np.ascontiguousarray(normals[s, lo:lo + cnt])).to(dev) if inject else None
            torch.cuda.synchronize()
            args = (t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), sc.num_beams, sc.d_center, sc.d_theta, tn.data_ptr() if inject else None)
            u = -1.0 if u_dev else float(uniforms[s])
            if stepper: stepper.step(*args, policy=B.POLICY_NEVER)
            else: h.update_begin_dev(*args); h.update_end_dev(B.POLICY_NEVER, 0.0)
            neff = h.read_neff()
            lw = h.log_weights().copy(); poses0 = h.poses().copy()
            maps0 = {p: h.get_map(p, B.MAP_FREE_COUNT).sum() + 7 * h.get_map(p, B.MAP_OCC_COUNT).sum() for p in range(lo, lo + cnt)}
            if stepper: dist.barrier()
            h.resample(u)
            if stepper and stepper.migrates: dist.all_reduce(stepper._token)
            par = h.parents().copy(); poses1 = h.poses().copy()
            maps1 = {p: (int(h.get_map(p, B.MAP_FREE_COUNT).sum()), int(h.get_map(p, B.MAP_OCC_COUNT).sum()), float(h.get_map(p, B.MAP_LIKELIHOOD).sum())) for p in range(lo, lo + cnt)}
            out.append(dict(lw=lw, poses0=poses0, maps0=maps0, par=par, poses1=poses1, maps1=maps1, neff=neff))
            if stepper: dist.barrier()
        return out
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    h = lib.create(rank=rank, nranks=world, device=rank, **kw); h.set_stream(stream.cuda_stream)
    st = parallel.ShardedStepper(h, dist, dev)
    mine = run(h, st, h.info.local_begin, h.info.local_count)
    dist.barrier(); h.close()
    ref = [None]
    if rank == 0:
        h1 = lib.create(rank=0, nranks=1, device=0, **kw); h1.set_stream(stream.cuda_stream)
        ref[0] = run(h1, None, 0, P); h1.close()
    dist.broadcast_object_list(ref, src=0); ref = ref[0]
    msgs = []
    for s, (a, b) in enumerate(zip(mine, ref)):
        for k in ("lw", "poses0", "par", "poses1"):
            if not np.array_equal(a[k], b[k]):
                bad = np.flatnonzero((a[k] != b[k]).reshape(len(a[k]), -1).any(axis=1))
                msgs.append(f"step {s} {k} differs at {bad[:8].tolist()} (n={len(bad)})")
        for k in ("maps0", "maps1"):
            bad = [p for p in a[k] if a[k][p] != b[k][p]]
            if bad: msgs.append(f"step {s} {k} differs for particles {bad[:8]} (n={len(bad)}) e.g. {a[k][bad[0]]} vs {b[k][bad[0]]} parent={b['par'][bad[0]]}")
        if msgs: break
    print(f"[rank {rank}] P={P} inject={inject} u_dev={u_dev}:", msgs if msgs else "identical", flush=True)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
dist.barrier(); dist.destroy_process_group()
