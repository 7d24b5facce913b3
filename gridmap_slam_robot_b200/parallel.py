"""Multi-rank stepping: one process per GPU, particles block-partitioned over ranks.

Two data planes for the one exchange a step needs (the weights, before resampling):

* PEER path (CUDA handles, default): every rank maps the others' receive buffers, flags and pose arrays with
  cudaIpc.  After scoring, each rank stores its block of f64 log-weights (8 B per particle) into every rank's
  receive buffer over NVLink and raises a flag; the normalise kernel of every rank waits for the flags itself.
  Poses are not exchanged at all: the resampling reads a remote parent's 16-byte pose through the peer mapping.
  No collective, no host synchronisation, no import pass.
* COLLECTIVE path (the oracle under gloo in the CPU tests, or `use_peer_memory=False`): all-gather of 24-byte
  records {f64 log-weight, f32 x, y, theta, u32 pad} between begin and end.

Everything after the exchange (normalise, Neff, strongest, CDF, parent indices, shared-map integration) is
computed redundantly and bit-identically on every rank over fixed 1024-particle tiles of the global particle
index, so no broadcast follows and the results do not depend on the number of ranks.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import binding as B


class _CudaBlock:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v3)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def wrap_block(ptr, nbytes, device):
    """A uint8 torch tensor aliasing `nbytes` at `ptr` (device memory for CUDA devices, host otherwise)."""
    if torch.device(device).type == "cuda":
        return torch.as_tensor(_CudaBlock(ptr, nbytes), device=device)
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return torch.from_numpy(np.ctypeslib.as_array(buf))


class ShardedStepper:
    def __init__(self, handle: B.Handle, dist, device, use_peer_memory=True):
        self.h, self.dist, self.device = handle, dist, device
        dl, lb, dg, gb = handle.exchange_buffers()
        self.local = wrap_block(dl, lb, device)
        self.glob = wrap_block(dg, gb, device)
        assert gb == lb * dist.get_world_size()
        # per-particle maps: resampling moves maps between GPUs (SLAM.java:41-45 deep copy); every rank
        # maps the other ranks' arenas (cudaIpc) so its copy kernel can pull parents' maps over NVLink
        self.migrates = handle.cfg.map_mode == B.MAP_PER_PARTICLE and dist.get_world_size() > 1
        # CUDA handles map each other's buffers (cudaIpc): PEER path, and per-particle maps can migrate.
        # The oracle (CPU tests, gloo) has no device arenas: it keeps the explicit all-gather.
        self.direct = False
        if handle.info.is_cuda and dist.get_world_size() > 1 and use_peer_memory:
            mine = handle.ipc_export()
            every = [None] * dist.get_world_size()
            dist.all_gather_object(every, mine)
            handle.ipc_import(b"".join(every))
            self.direct = True
        if self.migrates:
            self._token = torch.zeros(1, dtype=torch.int32, device=device)

    def step(self, d_xy, d_dist, d_hit, num_beams, d_center, d_theta, d_normals=None, policy=B.POLICY_NEVER,
             u01=-1.0):
        """One SLAM step across all ranks.  Pointers are device pointers (host pointers for the oracle)."""
        self.h.update_begin_dev(d_xy, d_dist, d_hit, num_beams, d_center, d_theta, d_normals)
        if not self.direct:
            # NCCL / gloo: ordered after the begin kernels on the current stream, and the end kernels after it
            self.dist.all_gather_into_tensor(self.glob, self.local)
        self.h.update_end_dev(policy, u01)
        if self.migrates and policy != B.POLICY_NEVER:
            # stream-ordered barrier: no rank starts the next map update before every pull has finished
            self.dist.all_reduce(self._token)

    def combined_map(self):
        """GridMapApp.calculateCombined (GridMapApp.java:439-458) over the particles of ALL ranks: every rank
        multiplies its own particles' factors, one PRODUCT all-reduce of W*H doubles joins them."""
        ptr, nbytes = self.h.combined_map_begin()
        prod = wrap_block(ptr, nbytes, self.device).view(torch.float64)
        self.dist.all_reduce(prod, op=self.dist.ReduceOp.PRODUCT)
        return self.h.combined_map_end()
