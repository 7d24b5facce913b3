"""CPU, world_size 2 and 4 over gloo: the multi-rank stepping logic (gridmap_slam_robot_b200/parallel.py)
around the oracle library.  Rank-count invariance: parents, poses, weights and the
integer map counts of an R-rank run equal those of the 1-rank run on the same inputs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_SO = os.path.join(ROOT, "oracle", "libgms_ref.so")


def _config(P):
    return dict(num_particles=P, map_width_m=12.0, map_height_m=12.0, origin_x=-6.0, origin_y=-6.0, map_mode=1,
                resample_mode=2, seed=99)


def _inputs(steps, P, beams):
    from gridmap_slam_robot_b200 import synth

    scans = synth.make_scans(steps, beams, max_range=10.0)
    normals, uniforms = synth.make_draws(steps, P)
    return scans, normals, uniforms


def _run_steps(h, stepper, scans, normals, uniforms, lo, cnt):
    from gridmap_slam_robot_b200 import binding as B

    out = []
    for s, sc in enumerate(scans):
        nz = np.ascontiguousarray(normals[s, lo:lo + cnt])
        xy, d, hit = (np.ascontiguousarray(a) for a in (sc.beam_xy, sc.beam_dist, sc.beam_hit))
        args = (xy.ctypes.data, d.ctypes.data, hit.ctypes.data, d.size, sc.d_center, sc.d_theta, nz.ctypes.data)
        if stepper:
            stepper.step(*args, policy=B.POLICY_ALWAYS, u01=float(uniforms[s]))
        else:
            h.update_begin_dev(*args)
            h.update_end_dev(B.POLICY_ALWAYS, float(uniforms[s]))
        out.append((h.read_neff(), h.parents().copy(), h.poses().copy(), h.weights().copy(),
                    h.get_map(0, B.MAP_FREE_COUNT).copy(), h.get_map(0, B.MAP_OCC_COUNT).copy()))
    return out


def _worker(rank, world, port, P, steps, beams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import parallel

    lib = B.Library(ORACLE_SO)
    h = lib.create(rank=rank, nranks=world, **_config(P))
    assert h.info.local_count == P // world and h.info.local_begin == rank * (P // world)
    stepper = parallel.ShardedStepper(h, dist, "cpu")
    scans, normals, uniforms = _inputs(steps, P, beams)
    res = _run_steps(h, stepper, scans, normals, uniforms, h.info.local_begin, h.info.local_count)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_equal_one_rank(oracle, world):
    from gridmap_slam_robot_b200 import binding as B

    P, steps, beams = 64, 4, 60
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, steps, beams, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    h = oracle.create(**_config(P))
    scans, normals, uniforms = _inputs(steps, P, beams)
    single = _run_steps(h, None, scans, normals, uniforms, 0, P)
    for r in range(world):
        for s in range(steps):
            a, b = results[r][s], single[s]
            assert abs(a[0] - b[0]) < 1e-9 * b[0]
            for k in range(1, 6):
                assert np.array_equal(a[k], b[k]), (r, s, k)
    h.close()
