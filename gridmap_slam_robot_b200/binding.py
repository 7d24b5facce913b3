"""ctypes binding of include/gms.h.

`load()` opens the CUDA product library (csrc/libgms.so) and fails loudly when it is missing or
when no CUDA device is usable — there is no CPU fallback.  `Library(path)` binds any shared object
that implements the header; the tests use it to drive the CPU oracle through the same code.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBGMS_PATH = os.path.join(_HERE, "csrc", "libgms.so")

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
MAP_PER_PARTICLE, MAP_SHARED = 0, 1
RESAMPLE_AUTO, RESAMPLE_LITERAL, RESAMPLE_FIXED = 0, 1, 2
MAP_LOG, MAP_LIKELIHOOD, MAP_FREE_COUNT, MAP_OCC_COUNT = 0, 1, 2, 3
POLICY_NEVER, POLICY_IF_NEFF_LOW, POLICY_ALWAYS = 0, 1, 2
UPDATE_ATOMIC, UPDATE_SORTED = 0, 1
IPC_NUM_HANDLES = 18
MAX_BEAMS = 12800  # GMS_MAX_BEAMS
PHASES = ("motion", "likelihood", "score", "normalise", "map_update", "resample", "map_copy", "exchange", "other")


class GmsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gms error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_particles", C.c_int32),
        ("map_width_m", C.c_float), ("map_height_m", C.c_float), ("resolution", C.c_float),
        ("origin_x", C.c_float), ("origin_y", C.c_float), ("sensor_max_range", C.c_float),
        ("z_hit", C.c_double), ("hit_tolerance", C.c_float), ("extra_steps", C.c_int32),
        ("p_free", C.c_float), ("p_occ", C.c_float),
        ("noise_center_base", C.c_double), ("noise_center_gain", C.c_double),
        ("noise_theta_base_deg", C.c_double), ("noise_theta_gain", C.c_double),
        ("skip_update_deg", C.c_double), ("likelihood_sigma_num", C.c_double),
        ("map_mode", C.c_int32), ("resample_mode", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("update_mode", C.c_int32),
        ("seed", C.c_uint64),
    ]


class Info(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("is_cuda", C.c_int32), ("grid_w", C.c_int32), ("grid_h", C.c_int32),
        ("num_particles", C.c_int32), ("local_begin", C.c_int32), ("local_count", C.c_int32),
        ("num_slots", C.c_int32), ("kernel_taps", C.c_int32), ("resample_mode", C.c_int32),
        ("kernel", C.c_double * 32), ("l_free", C.c_double), ("l_occ", C.c_double),
        ("world_w", C.c_double), ("world_h", C.c_double),
    ]


_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_P = C.POINTER
POSE_OPTIMIZER_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f64, _f64)

# every symbol include/gms.h declares: name -> argtypes (restype is int unless noted)
SYMBOLS = {
    "gms_config_default": [_P(Config)],
    "gms_create": [_P(Config), _P(_vp)],
    "gms_destroy": [_vp],
    "gms_last_error": [_vp],
    "gms_get_info": [_vp, _P(Info)],
    "gms_reset": [_vp],
    "gms_update": [_vp, _vp, _vp, _vp, _i32, _f64, _f64, _vp, _P(_f64)],
    "gms_resample": [_vp, _f64],
    "gms_calculate_neff": [_vp, _P(_f64)],
    "gms_get_weighted_pose": [_vp, _vp],
    "gms_get_strongest": [_vp, _P(_i32), _vp, _P(_f64)],
    "gms_get_poses": [_vp, _vp],
    "gms_get_weights": [_vp, _vp],
    "gms_get_log_weights": [_vp, _vp],
    "gms_get_parents": [_vp, _vp],
    "gms_get_map": [_vp, _i32, _i32, _vp, C.c_size_t],
    "gms_set_poses": [_vp, _vp],
    "gms_set_weights": [_vp, _vp],
    "gms_set_map_counts": [_vp, _i32, _vp, _vp],
    "gms_map_apply_measurement": [_vp, _i32, _f32, _f32, _f32, _f32, _f32, _i32],
    "gms_map_integrate_observation": [_vp, _i32, _vp, _vp, _vp, _vp, _i32],
    "gms_map_compute_likelihood": [_vp, _i32],
    "gms_map_probability_of": [_vp, _i32, _vp, _vp, _vp, _i32, _P(_f64), _P(_f64)],
    "gms_trace_rays": [_vp, _vp, _i32, _i32, _vp, _i32, _vp],
    "gms_odometry_from_counts": [_i32, _i32, _P(_f64), _P(_f64)],
    "gms_step_dev": [_vp, _vp, _vp, _vp, _i32, _f64, _f64, _vp, _i32, _f64],
    "gms_sync": [_vp],
    "gms_join_streams": [_vp],
    "gms_set_pose_optimizer": [_vp, _vp, _vp],
    "gms_set_stream": [_vp, _vp],
    "gms_profile_enable": [_vp, _i32],
    "gms_profile_read": [_vp, _vp, _vp],
    "gms_profile_reset": [_vp],
    "gms_launch_count": [_vp, _P(_i64)],
    "gms_exchange_buffers": [_vp, _P(_vp), _P(C.c_size_t), _P(_vp), _P(C.c_size_t)],
    "gms_update_begin_dev": [_vp, _vp, _vp, _vp, _i32, _f64, _f64, _vp],
    "gms_update_end_dev": [_vp, _i32, _f64],
    "gms_read_neff": [_vp, _P(_f64)],
    "gms_ipc_export": [_vp, _vp],
    "gms_ipc_import": [_vp, _vp],
    "gms_deskew": [_vp, _vp, _vp, _i32, _f64, _f64, _vp, _vp],
    "gms_update_raw": [_vp, _vp, _vp, _vp, _i32, _f64, _f64, _vp, _P(_f64)],
    "gms_render_map": [_vp, _i32, _i32, _vp],
    "gms_combined_map": [_vp, _vp, _vp],
    "gms_combined_map_begin_dev": [_vp, _P(_vp), _P(C.c_size_t)],
    "gms_combined_map_end": [_vp, _vp, _vp],
}


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.__array_interface__["data"][0]  # the address, without building a ctypes helper object (a.ctypes)


def _arr(a, dtype, n=None):
    # the common case costs nothing: an ndarray of the right dtype that is already C-contiguous is passed through
    if not (type(a) is np.ndarray and a.dtype == dtype and a.flags.c_contiguous):
        a = np.ascontiguousarray(a, dtype=dtype)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} elements, got {a.size}")
    return a


class Library:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.dll = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
        for name, args in SYMBOLS.items():
            fn = getattr(self.dll, name)  # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = C.c_char_p if name == "gms_last_error" else C.c_int

    def default_config(self, **kw) -> Config:
        cfg = Config()
        rc = self.dll.gms_config_default(C.byref(cfg))
        if rc:
            raise GmsError(rc, "gms_config_default")
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise AttributeError(f"gms_config has no field {k!r}")
            setattr(cfg, k, v)
        return cfg

    def create(self, cfg: Config | None = None, **kw) -> "Handle":
        cfg = cfg if cfg is not None else self.default_config(**kw)
        return Handle(self, cfg)

    def odometry_from_counts(self, left, right):
        dc, dt = _f64(), _f64()
        rc = self.dll.gms_odometry_from_counts(left, right, C.byref(dc), C.byref(dt))
        if rc:
            raise GmsError(rc, "gms_odometry_from_counts")
        return dc.value, dt.value


class Handle:
    """One gms_handle.  Methods mirror the C entry points 1:1 and take/return numpy arrays."""

    def __init__(self, lib: Library, cfg: Config):
        self.lib, self.dll = lib, lib.dll
        self.h = _vp()
        rc = self.dll.gms_create(C.byref(cfg), C.byref(self.h))
        if rc:
            msg = self.dll.gms_last_error(None)
            self.h = None
            raise GmsError(rc, (msg or b"").decode())
        self.cfg = cfg
        self.info = Info()
        self._ck(self.dll.gms_get_info(self.h, C.byref(self.info)))
        self.W, self.H, self.P = self.info.grid_w, self.info.grid_h, self.info.num_particles

    def _ck(self, rc):
        if rc:
            raise GmsError(rc, (self.dll.gms_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.dll.gms_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- step ----
    def reset(self):
        self._ck(self.dll.gms_reset(self.h))

    def update(self, beam_xy, beam_dist, beam_hit, d_center, d_theta, normals=None):
        d = _arr(beam_dist, np.float64)
        B = d.size
        xy = _arr(beam_xy, np.float64, 2 * B)
        hit = _arr(beam_hit, np.uint8, B)
        nz = None if normals is None else _arr(normals, np.float64, 2 * self.info.local_count)
        neff = _f64()
        self._ck(self.dll.gms_update(self.h, _ptr(xy), _ptr(d), _ptr(hit), B, d_center, d_theta, _ptr(nz),
                                     C.byref(neff)))
        return neff.value

    def resample(self, u01=-1.0):
        self._ck(self.dll.gms_resample(self.h, u01))

    def calculate_neff(self):
        v = _f64()
        self._ck(self.dll.gms_calculate_neff(self.h, C.byref(v)))
        return v.value

    def weighted_pose(self):
        p = np.zeros(3, np.float32)
        self._ck(self.dll.gms_get_weighted_pose(self.h, _ptr(p)))
        return p

    def strongest(self):
        idx, w, p = _i32(), _f64(), np.zeros(3, np.float32)
        self._ck(self.dll.gms_get_strongest(self.h, C.byref(idx), _ptr(p), C.byref(w)))
        return idx.value, p, w.value

    def poses(self):
        p = np.zeros((self.P, 3), np.float32)
        self._ck(self.dll.gms_get_poses(self.h, _ptr(p)))
        return p

    def weights(self):
        w = np.zeros(self.P, np.float64)
        self._ck(self.dll.gms_get_weights(self.h, _ptr(w)))
        return w

    def log_weights(self):
        w = np.zeros(self.P, np.float64)
        self._ck(self.dll.gms_get_log_weights(self.h, _ptr(w)))
        return w

    def parents(self):
        p = np.zeros(self.P, np.int32)
        self._ck(self.dll.gms_get_parents(self.h, _ptr(p)))
        return p

    def get_map(self, particle, kind):
        dt = np.float64 if kind in (MAP_LOG, MAP_LIKELIHOOD) else np.uint32
        m = np.zeros((self.H, self.W), dt)
        self._ck(self.dll.gms_get_map(self.h, particle, kind, _ptr(m), m.nbytes))
        return m

    # ---- injection ----
    def set_poses(self, xyt):
        self._ck(self.dll.gms_set_poses(self.h, _ptr(_arr(xyt, np.float32, 3 * self.P))))

    def set_weights(self, w):
        self._ck(self.dll.gms_set_weights(self.h, _ptr(_arr(w, np.float64, self.P))))

    def set_map_counts(self, particle, n_free, n_occ):
        n = self.W * self.H
        self._ck(self.dll.gms_set_map_counts(self.h, particle, _ptr(_arr(n_free, np.uint32, n)),
                                             _ptr(_arr(n_occ, np.uint32, n))))

    # ---- GridMap operators ----
    def map_apply_measurement(self, particle, sx, sy, ex, ey, meas, hit):
        self._ck(self.dll.gms_map_apply_measurement(self.h, particle, sx, sy, ex, ey, meas, int(bool(hit))))

    def map_integrate_observation(self, particle, pose, beam_xy, beam_dist, beam_hit):
        d = _arr(beam_dist, np.float64)
        B = d.size
        self._ck(self.dll.gms_map_integrate_observation(
            self.h, particle, _ptr(_arr(pose, np.float32, 3)), _ptr(_arr(beam_xy, np.float64, 2 * B)), _ptr(d),
            _ptr(_arr(beam_hit, np.uint8, B)), B))

    def map_compute_likelihood(self, particle):
        self._ck(self.dll.gms_map_compute_likelihood(self.h, particle))

    def map_probability_of(self, particle, pose, beam_xy, beam_hit):
        hit = _arr(beam_hit, np.uint8)
        B = hit.size
        lp, p = _f64(), _f64()
        self._ck(self.dll.gms_map_probability_of(self.h, particle, _ptr(_arr(pose, np.float32, 3)),
                                                 _ptr(_arr(beam_xy, np.float64, 2 * B)), _ptr(hit), B,
                                                 C.byref(lp), C.byref(p)))
        return lp.value, p.value

    def trace_rays(self, rays, extra=2, cap=None):
        rays = _arr(rays, np.float32).reshape(-1, 4)
        n = rays.shape[0]
        cap = cap if cap is not None else 2 * (self.W + self.H) + extra + 4
        cells = np.full((n, cap, 2), -1, np.int32)
        counts = np.zeros(n, np.int32)
        self._ck(self.dll.gms_trace_rays(self.h, _ptr(rays), n, extra, _ptr(cells), cap, _ptr(counts)))
        return cells, counts

    # ---- rows adjacent to the path ----
    def deskew(self, angle, dist, d_center, d_theta):
        a, d = _arr(angle, np.float64), _arr(dist, np.float64)
        xy, od = np.zeros((a.size, 2), np.float64), np.zeros(a.size, np.float64)
        self._ck(self.dll.gms_deskew(self.h, _ptr(a), _ptr(d), a.size, d_center, d_theta, _ptr(xy), _ptr(od)))
        return xy, od

    def update_raw(self, angle, dist, hit, d_center, d_theta, normals=None):
        a = _arr(angle, np.float64)
        d, hh = _arr(dist, np.float64, a.size), _arr(hit, np.uint8, a.size)
        nz = None if normals is None else _arr(normals, np.float64, 2 * self.info.local_count)
        neff = _f64()
        self._ck(self.dll.gms_update_raw(self.h, _ptr(a), _ptr(d), _ptr(hh), a.size, d_center, d_theta, _ptr(nz),
                                         C.byref(neff)))
        return neff.value

    def render_map(self, particle, likelihood=False):
        out = np.zeros((self.H, self.W), np.uint32)
        self._ck(self.dll.gms_render_map(self.h, particle, int(likelihood), _ptr(out)))
        return out

    def combined_map(self):
        lg, lk = np.zeros((self.H, self.W), np.float64), np.zeros((self.H, self.W), np.float64)
        self._ck(self.dll.gms_combined_map(self.h, _ptr(lg), _ptr(lk)))
        return lg, lk

    def combined_map_begin(self):
        ptr, nb = _vp(), C.c_size_t()
        self._ck(self.dll.gms_combined_map_begin_dev(self.h, C.byref(ptr), C.byref(nb)))
        return ptr.value, nb.value

    def combined_map_end(self):
        lg, lk = np.zeros((self.H, self.W), np.float64), np.zeros((self.H, self.W), np.float64)
        self._ck(self.dll.gms_combined_map_end(self.h, _ptr(lg), _ptr(lk)))
        return lg, lk

    # ---- device-resident / multi-rank ----
    def step_dev(self, d_xy, d_dist, d_hit, B, d_center, d_theta, d_normals=None, policy=POLICY_NEVER, u01=-1.0):
        self._ck(self.dll.gms_step_dev(self.h, _ptr(d_xy), _ptr(d_dist), _ptr(d_hit), B, d_center, d_theta,
                                       _ptr(d_normals), policy, u01))

    def update_begin_dev(self, d_xy, d_dist, d_hit, B, d_center, d_theta, d_normals=None):
        self._ck(self.dll.gms_update_begin_dev(self.h, _ptr(d_xy), _ptr(d_dist), _ptr(d_hit), B, d_center,
                                               d_theta, _ptr(d_normals)))

    def update_end_dev(self, policy=POLICY_NEVER, u01=-1.0):
        self._ck(self.dll.gms_update_end_dev(self.h, policy, u01))

    def exchange_buffers(self):
        dl, dg, lb, gb = _vp(), _vp(), C.c_size_t(), C.c_size_t()
        self._ck(self.dll.gms_exchange_buffers(self.h, C.byref(dl), C.byref(lb), C.byref(dg), C.byref(gb)))
        return dl.value, lb.value, dg.value, gb.value

    def read_neff(self):
        v = _f64()
        self._ck(self.dll.gms_read_neff(self.h, C.byref(v)))
        return v.value

    def ipc_export(self) -> bytes:
        buf = (C.c_ubyte * (IPC_NUM_HANDLES * 64))()
        self._ck(self.dll.gms_ipc_export(self.h, C.addressof(buf)))
        return bytes(buf)

    def ipc_import(self, all_handles: bytes):
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        self._ck(self.dll.gms_ipc_import(self.h, C.addressof(buf)))

    def sync(self):
        self._ck(self.dll.gms_sync(self.h))

    def join_streams(self):
        self._ck(self.dll.gms_join_streams(self.h))

    def set_pose_optimizer(self, fn=None):
        """A4 hook (GridMap.findBestPoseOptim): fn(first_particle, poses[count,3] f32 in/out, beam_xy[B,2], beam_dist[B],
        beam_hit[B], d_center, d_theta) -> None, or None to restore the identity default."""
        if fn is None:
            self._opt_cb = None
            self._ck(self.dll.gms_set_pose_optimizer(self.h, None, None))
            return

        def tramp(user, hptr, first, count, poses, bxy, bdist, bhit, nb, dc, dt):
            try:
                p = np.ctypeslib.as_array(C.cast(poses, C.POINTER(C.c_float)), shape=(count, 3))
                xy = np.ctypeslib.as_array(C.cast(bxy, C.POINTER(C.c_double)), shape=(nb, 2)) if nb else np.zeros((0, 2))
                d = np.ctypeslib.as_array(C.cast(bdist, C.POINTER(C.c_double)), shape=(nb,)) if nb else np.zeros(0)
                hh = np.ctypeslib.as_array(C.cast(bhit, C.POINTER(C.c_uint8)), shape=(nb,)) if nb else np.zeros(0, np.uint8)
                fn(first, p, xy, d, hh, dc, dt)
                return 0
            except Exception:  # a Python exception must not unwind through the C frames
                import traceback

                traceback.print_exc()
                return 1

        self._opt_cb = POSE_OPTIMIZER_FN(tramp)
        self._ck(self.dll.gms_set_pose_optimizer(self.h, C.cast(self._opt_cb, _vp), None))

    def set_stream(self, stream_ptr):
        self._ck(self.dll.gms_set_stream(self.h, stream_ptr))

    def profile_enable(self, on=True):
        self._ck(self.dll.gms_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.dll.gms_profile_reset(self.h))

    def profile_read(self):
        ms = np.zeros(len(PHASES), np.float64)
        n = np.zeros(len(PHASES), np.int64)
        self._ck(self.dll.gms_profile_read(self.h, _ptr(ms), _ptr(n)))
        return dict(zip(PHASES, ms.tolist())), dict(zip(PHASES, n.tolist()))

    def launch_count(self):
        v = _i64()
        self._ck(self.dll.gms_launch_count(self.h, C.byref(v)))
        return v.value


_LIB = None


def load() -> Library:
    """The CUDA product library.  Raises if csrc/libgms.so has not been built."""
    global _LIB
    if _LIB is None:
        _LIB = Library(LIBGMS_PATH)
    return _LIB
