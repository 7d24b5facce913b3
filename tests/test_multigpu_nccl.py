"""GPU (-m gpu), needs >= 2 devices: particles sharded over ranks, NCCL all-gather of the exchange
records on the library's stream.  Rank-count invariance: the R-rank run equals the
1-rank run on the same inputs — parents, poses and integer map counts bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _cfg(P, mode=1):
    g = 51.2 if mode == 1 else 20.0
    return dict(num_particles=P, map_width_m=g, map_height_m=g, origin_x=-g / 2, origin_y=-g / 2, map_mode=mode,
                resample_mode=2, seed=4242)


def _steps(h, stepper, scans, normals, uniforms, lo, cnt, dev, map_particles=(0,)):
    import torch

    from gridmap_slam_robot_b200 import binding as B

    out = []
    for s, sc in enumerate(scans):
        t_xy = torch.from_numpy(sc.beam_xy).to(dev)
        t_d = torch.from_numpy(sc.beam_dist).to(dev)
        t_h = torch.from_numpy(sc.beam_hit).to(dev)
        t_n = torch.from_numpy(np.ascontiguousarray(normals[s, lo:lo + cnt])).to(dev)
        torch.cuda.synchronize()
        args = (t_xy.data_ptr(), t_d.data_ptr(), t_h.data_ptr(), sc.num_beams, sc.d_center, sc.d_theta, t_n.data_ptr())
        if stepper:
            stepper.step(*args, policy=B.POLICY_ALWAYS, u01=float(uniforms[s]))
        else:
            h.update_begin_dev(*args)
            h.update_end_dev(B.POLICY_ALWAYS, float(uniforms[s]))
        neff = h.read_neff()  # synchronises this rank's stream
        if stepper:
            stepper.dist.barrier()  # ... and now every rank's: the getters read other ranks' blocks through peer mappings
        maps = {p: (h.get_map(p, B.MAP_FREE_COUNT).copy(), h.get_map(p, B.MAP_OCC_COUNT).copy(),
                    h.get_map(p, B.MAP_LIKELIHOOD).copy()) for p in map_particles}
        out.append((neff, h.parents().copy(), h.poses().copy(), h.weights().copy(), maps))
        if stepper:
            stepper.dist.barrier()  # no rank starts the next step while another still reads through peer mappings
    return out


def _worker(rank, world, port, P, steps, beams, q, mode=1, peer=True, sharded=False):
    import torch
    import torch.distributed as dist

    os.environ["GMS_SHARDED"] = "1" if sharded else "0"  # read by gms_ipc_import
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import parallel, synth

    h = B.load().create(rank=rank, nranks=world, device=rank, **_cfg(P, mode))
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    stepper = parallel.ShardedStepper(h, dist, dev, use_peer_memory=peer or mode == 0)
    assert stepper.direct == (peer or mode == 0)
    scans = synth.make_scans(steps, beams, max_range=12.0)
    normals, uniforms = synth.make_draws(steps, P)
    lo, cnt = h.info.local_begin, h.info.local_count
    mp_ = (0,) if mode == 1 else tuple(range(lo, lo + cnt))
    res = _steps(h, stepper, scans, normals, uniforms, lo, cnt, dev, map_particles=mp_)
    comb = stepper.combined_map() if mode == 0 else None  # fusion over the particles of all ranks
    q.put((rank, res, comb))
    dist.barrier()
    dist.destroy_process_group()


def _run(cuda, P, steps, beams, world, mode, peer=True, sharded=False):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    from gridmap_slam_robot_b200 import synth

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, steps, beams, q, mode, peer, sharded)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    results = {r: res for r, res, _ in got}
    combined = {r: c for r, _, c in got}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    os.environ["GMS_SHARDED"] = "1" if sharded else "0"  # the single-rank run uses the same (exact-sum) normalise
    h = cuda.create(**_cfg(P, mode))
    os.environ.pop("GMS_SHARDED", None)
    scans = synth.make_scans(steps, beams, max_range=12.0)
    normals, uniforms = synth.make_draws(steps, P)
    single = _steps(h, None, scans, normals, uniforms, 0, P, torch.device("cuda", 0),
                    map_particles=(0,) if mode == 1 else tuple(range(P)))
    for r in range(world):
        for s in range(steps):
            a, b = results[r][s], single[s]
            assert abs(a[0] / b[0] - 1) < 1e-12
            assert np.array_equal(a[1], b[1]), (r, s, "parents")
            assert np.array_equal(a[2], b[2]), (r, s, "poses")
            np.testing.assert_allclose(a[3], b[3], rtol=1e-12, atol=0)
            for p, (nf, no, lik) in a[4].items():
                assert np.array_equal(nf, b[4][p][0]) and np.array_equal(no, b[4][p][1]), (r, s, p, "counts")
                assert np.array_equal(lik, b[4][p][2]), (r, s, p, "likelihood")
    if mode == 0:  # GridMapApp.calculateCombined over all ranks == over one rank (product order differs: tolerance)
        lg1, lk1 = h.combined_map()
        for r in range(world):
            lg, lk = combined[r]
            # 1 - prod cancels where every particle says "free": the rounding of a differently associated product is
            # amplified there, so the bar is 1e-6 relative (single rank vs the oracle, same order: 1e-9)
            np.testing.assert_allclose(lg, lg1, rtol=1e-6, atol=1e-9)
            assert np.mean(lk == lk1) > 0.999
    h.close()


WORLDS = [2, 4, 8]


@pytest.mark.parametrize("world", WORLDS)
def test_ranks_equal_one_gpu(cuda, world):
    """Peer exchange: log-weights pushed into every rank's receive buffer (NVLink), poses read through peer maps."""
    _run(cuda, P=8192, steps=6, beams=360, world=world, mode=1)


@pytest.mark.parametrize("world", WORLDS)
def test_ranks_equal_one_gpu_sharded_normalise(cuda, world):
    """GMS_SHARDED=1: every rank normalises and resamples only its own block (exact 128-bit sums, four tiny exchange
    rounds); same bar."""
    _run(cuda, P=8192, steps=6, beams=360, world=world, mode=1, sharded=True)


def test_two_gpus_equal_one_gpu_nccl_all_gather(cuda):
    """Same run with the explicit NCCL all-gather between begin and end."""
    _run(cuda, P=8192, steps=4, beams=360, world=2, mode=1, peer=False)


@pytest.mark.parametrize("world", WORLDS)
def test_per_particle_maps_migrate(cuda, world):
    """K5-style: every particle owns a map; resampling every step moves maps between the GPUs (peer pull over
    NVLink).  Every particle's counts and likelihood field equal the 1-GPU run."""
    _run(cuda, P=24 * world, steps=5, beams=180, world=world, mode=0)


def _rankcheck_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gridmap_slam_robot_b200 import binding as B
    from gridmap_slam_robot_b200 import rankcheck

    lib = B.load()
    out = []
    out.append(rankcheck.check(lib, dist, dev, rank, world, per_particle=False, P=16384, beams=360, steps=6, grid_m=51.2))
    out.append(rankcheck.check(lib, dist, dev, rank, world, per_particle=True, P=64 * world, beams=180, steps=5, grid_m=20.0))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", WORLDS)
def test_bench_parity_check(cuda, world):
    """The check bench.py runs before timing at N > 1 (device Philox keyed by the global particle index,
    resampling every step): N ranks == 1 rank for parents, pose bytes, weights and every map."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rankcheck_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        for ok, detail in results[r]:
            assert ok, (r, detail)
