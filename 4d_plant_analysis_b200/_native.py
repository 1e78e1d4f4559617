"""ctypes binding of liblocalicp.so (C ABI declared in include/local_icp.h).

There is no CPU path behind these calls: if the shared library has not been built, or no sm_100
device is present, every entry point raises.  ``build()`` compiles the library in-tree with nvcc for
sm_100a (no GPU needed to compile).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("LIP_LIBRARY") or os.path.join(PKG_DIR, "liblocalicp.so")   # (LIP_LIBRARY: tuning variants)
SOURCES = ["local_icp.cu", "lip_kernels.cuh", "lip_device.cuh"]
HEADER = os.path.join(ROOT, "include", "local_icp.h")

DIR_T21, DIR_T12, DIR_BOTH = 1, 2, 3
ABI_VERSION = 2          # == LIP_ABI_VERSION of include/local_icp.h; bump both whenever a signature or struct layout changes

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC", "-DLIP_ARCH=100"]

# name -> (restype, argtypes); must list every symbol include/local_icp.h declares
_D, _I64, _I32 = POINTER(c_double), POINTER(c_int64), POINTER(c_int32)
SIGNATURES = {
    "lip_default_params": (None, [c_void_p]),
    "lip_create": (c_int, [POINTER(c_void_p), c_int]),
    "lip_destroy": (None, [c_void_p]),
    "lip_last_error": (c_char_p, [c_void_p]),
    "lip_set_params": (c_int, [c_void_p, c_void_p]),
    "lip_set_map_mode": (c_int, [c_void_p, c_int]),
    "lip_host_alloc": (c_int, [POINTER(c_void_p), c_uint64]),
    "lip_set_input_validation": (c_int, [c_void_p, c_int]),
    "lip_host_free": (c_int, [c_void_p]),
    "lip_batch_plan": (c_int, [c_void_p, c_int64, _I64, _I64, _I64, c_int]),
    "lip_batch_run_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lip_batch_run_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lip_set_map_dtype": (c_int, [c_void_p, c_int]),
    "lip_series_run_host": (c_int, [c_void_p, c_void_p, c_int64, _I64, _I64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lip_series_run_device": (c_int, [c_void_p, c_void_p, c_int64, _I64, _I64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lip_batch_get_transforms": (c_int, [c_void_p, c_int, _D, _D, _I32, _I32, _D, _D, _I32, _I32]),
    "lip_batch_get_landmark_nn": (c_int, [c_void_p, c_int, _I32, _D]),
    "lip_batch_get_counters": (c_int, [c_void_p, _I64]),
    "lip_batch_get_pair_counts": (c_int, [c_void_p, _I64]),
    "lip_set_profiling": (c_int, [c_void_p, c_int]),
    "lip_batch_get_stage_ms": (c_int, [c_void_p, POINTER(ctypes.c_float)]),
    "lip_batch_get_bin_offsets": (c_int, [c_void_p, _I32]),
    "lip_batch_evaluate_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, _D, _I64]),
    "lip_batch_evaluate_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, _I64, _I64, c_int64, c_void_p]),
    "lip_batch_refine_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lip_batch_match_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p, _D]),
    "lip_batch_interpolate_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "lip_batch_interpolate_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p]),
    "lip_icp_host": (c_int, [c_void_p, c_int64, _D, _I64, _D, _I64, _D, _D, _I32, _I32]),
    "lip_nn_host": (c_int, [c_void_p, _D, c_int64, _D, c_int64, _D, _I64]),
    "lip_kabsch_host": (c_int, [c_void_p, _D, _D, c_int64, _D, _D, _I32]),
    "lip_triangle_angles_host": (c_int, [c_void_p, _D, c_int64, _D]),
    "lip_version": (c_int, []),
    "lip_compiled_arch": (c_int, []),
}


class LipParams(ctypes.Structure):
    """mirror of struct lip_params"""
    _fields_ = [("max_iter", c_int32), ("cost_rise_tol", c_double), ("cost_flat_tol", c_double),
                ("obtuse_eps", c_double), ("cos_den_eps", c_double), ("reject_dist", c_double)]


class NativeError(RuntimeError):
    pass


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile liblocalicp.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "local_icp.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise NativeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None
_lock = threading.RLock()


def load():
    """dlopen the library and declare the prototypes.  Never builds implicitly on a box without nvcc
    sources changed: a missing library is an error, not a reason to fall back."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.lip_version.restype, lib.lip_version.argtypes = c_int, []
        if lib.lip_version() != ABI_VERSION:     # an old binary against new prototypes would crash or, worse, return wrong maps
            raise NativeError(f"{LIB_PATH} has ABI version {lib.lip_version()}, the Python layer expects {ABI_VERSION}: "
                              "rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def _ptr(a, ctype):
    return a.ctypes.data_as(POINTER(ctype)) if a is not None else None


def _f64(a, shape_last=3):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None:
        a = a.reshape(-1, shape_last)
    return a


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Context:
    """One lip_ctx: a CUDA stream plus grow-only device workspace on one GPU.  A context is single-threaded: plan, run and
    the getters of one call sequence belong together.  Code that shares a context between threads (the drop-in functions
    all use ``default_context()``) holds ``ctx.lock`` around the whole sequence."""

    def __init__(self, device: int = 0):
        self.lock = threading.RLock()
        self.lib = load()
        h = c_void_p()
        rc = self.lib.lip_create(byref(h), int(device))
        if rc != 0:
            msg = self.lib.lip_last_error(None)
            raise NativeError(f"lip_create(device={device}) failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h
        self.device = int(device)
        self._plan_key = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.lip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.lip_last_error(self.h)
            err = f"liblocalicp error {rc}: {msg.decode() if msg else ''}"
            raise ValueError(err) if rc == -2 else NativeError(err)

    # ---- parameters ---------------------------------------------------------------------------
    def default_params(self) -> LipParams:
        p = LipParams()
        self.lib.lip_default_params(byref(p))
        return p

    def set_params(self, **kw):
        p = self.default_params()
        for k, v in kw.items():
            setattr(p, k, v)
        self._check(self.lib.lip_set_params(self.h, byref(p)))

    def set_map_mode(self, mode: int):
        """0 = landmark-bin tiles on the CUDA cores (default), 1 = hashed grid, 2 = the tiles with the scores on the tensor cores (tcgen05 + TMEM); same maps in every mode; plan again afterwards."""
        self._check(self.lib.lip_set_map_mode(self.h, int(mode)))

    # ---- batch --------------------------------------------------------------------------------
    def set_input_validation(self, on: bool):
        """run_host raises ValueError on NaN / infinity in the inputs (checked on the library's copy threads)"""
        self._check(self.lib.lip_set_input_validation(self.h, int(bool(on))))

    def plan(self, off1, off2, lm_off, directions=DIR_BOTH):
        off1, off2, lm_off = _i64(off1), _i64(off2), _i64(lm_off)
        S = len(off1) - 1
        if len(off2) != S + 1 or len(lm_off) != S + 1:
            raise ValueError("off1, off2 and lm_off must have the same length")
        self._check(self.lib.lip_batch_plan(self.h, S, _ptr(off1, c_int64), _ptr(off2, c_int64), _ptr(lm_off, c_int64),
                                            int(directions)))
        self.S, self.n1, self.n2, self.sumK, self.directions = S, int(off1[-1]), int(off2[-1]), int(lm_off[-1]), int(directions)

    map_dtype = np.int64

    def set_map_dtype(self, dtype):
        """entry type of the maps returned by run_host / run_series / run_device: np.int64 (default, as the reference), np.int32,
        or np.uint16 (every mesh of the batch must then have fewer than 65,536 vertices)"""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.uint16), np.dtype(np.int32), np.dtype(np.int64)):
            raise ValueError("map dtype must be uint16, int32 or int64")
        self._check(self.lib.lip_set_map_dtype(self.h, 8 * dtype.itemsize))
        self.map_dtype = dtype.type

    def _maps_out(self, T21, T12):
        for T, n in ((T21, self.n2), (T12, self.n1)):
            if T is not None and (T.dtype != self.map_dtype or len(T) != n or not T.flags.c_contiguous):
                raise ValueError("output maps must be contiguous arrays of the context's map dtype and the planned size")
        if T21 is None and self.directions & DIR_T21:
            T21 = np.empty(self.n2, self.map_dtype)
        if T12 is None and self.directions & DIR_T12:
            T12 = np.empty(self.n1, self.map_dtype)
        return T21, T12

    def run_series(self, frames, row1, row2, lm1, lm2, T21=None, T12=None):
        """time series: ``frames`` (n_rows,3) holds every mesh once; row1[s] / row2[s] = first row of mesh1 / mesh2 of segment
        pair s of the planned batch (meshes may be shared between pairs); one upload of every mesh."""
        frames, lm1, lm2, row1, row2 = _f64(frames), _f64(lm1), _f64(lm2), _i64(row1), _i64(row2)
        if len(row1) != self.S or len(row2) != self.S or len(lm1) != self.sumK or len(lm2) != self.sumK:
            raise ValueError("array sizes do not match the planned batch")
        T21, T12 = self._maps_out(T21, T12)
        self._check(self.lib.lip_series_run_host(self.h, frames.ctypes.data, len(frames), _ptr(row1, c_int64), _ptr(row2, c_int64),
                                                 lm1.ctypes.data, lm2.ctypes.data,
                                                 T21.ctypes.data if T21 is not None else None,
                                                 T12.ctypes.data if T12 is not None else None))
        return T21, T12

    def run_series_ptr(self, frames, n_rows, row1, row2, lm1, lm2, T21, T12):
        """same, raw addresses of (pinned) host buffers; row1 / row2 are int64 numpy arrays"""
        self._check(self.lib.lip_series_run_host(self.h, frames, int(n_rows), _ptr(row1, c_int64), _ptr(row2, c_int64), lm1, lm2, T21, T12))

    def run_series_device(self, d_frames, n_rows, row1, row2, d_lm1, d_lm2, d_T21, d_T12, stream=0):
        """device addresses, enqueued on `stream`; no sync.  row1 / row2: int64 numpy arrays (host)"""
        self._check(self.lib.lip_series_run_device(self.h, d_frames, int(n_rows), _ptr(row1, c_int64), _ptr(row2, c_int64),
                                                   d_lm1, d_lm2, d_T21, d_T12, stream))

    def run_host(self, xyz1, xyz2, lm1, lm2, T21=None, T12=None):
        """host numpy in -> host numpy out (H2D, kernels, D2H, sync inside)."""
        xyz1, xyz2, lm1, lm2 = _f64(xyz1), _f64(xyz2), _f64(lm1), _f64(lm2)
        if len(xyz1) != self.n1 or len(xyz2) != self.n2 or len(lm1) != self.sumK or len(lm2) != self.sumK:
            raise ValueError("array sizes do not match the planned batch")
        T21, T12 = self._maps_out(T21, T12)
        self._check(self.lib.lip_batch_run_host(self.h, xyz1.ctypes.data, xyz2.ctypes.data, lm1.ctypes.data, lm2.ctypes.data,
                                                T21.ctypes.data if T21 is not None else None,
                                                T12.ctypes.data if T12 is not None else None))
        return T21, T12

    def run_host_ptr(self, xyz1, xyz2, lm1, lm2, T21, T12):
        """same, raw addresses (pinned buffers of the bench)."""
        self._check(self.lib.lip_batch_run_host(self.h, xyz1, xyz2, lm1, lm2, T21, T12))

    def run_device(self, xyz1, xyz2, lm1, lm2, T21, T12, stream=0):
        """device addresses in/out, enqueued on `stream` (0 = the context's own stream); no sync."""
        self._check(self.lib.lip_batch_run_device(self.h, xyz1, xyz2, lm1, lm2, T21, T12, stream))

    def evaluate(self, xyz1, xyz2, lm1, lm2, T21, T12=None, voxel_size=1.0, sample_proportion=0.1):
        """scores (S,3) [continuity, landmark precision, cycle] and the integer counts (S,6) behind them"""
        xyz1, xyz2, lm1, lm2 = _f64(xyz1), _f64(xyz2), _f64(lm1), _f64(lm2)
        T21 = _i64(T21)
        T12 = _i64(T12) if T12 is not None else None
        if len(xyz1) != self.n1 or len(xyz2) != self.n2 or len(lm1) != self.sumK or len(lm2) != self.sumK or len(T21) != self.n2 \
                or (T12 is not None and len(T12) != self.n1):
            raise ValueError("array sizes do not match the planned batch")
        scores, counts = np.empty((self.S, 3)), np.empty((self.S, 6), np.int64)
        self._check(self.lib.lip_batch_evaluate_host(self.h, xyz1.ctypes.data, xyz2.ctypes.data, lm1.ctypes.data, lm2.ctypes.data,
                                                     T21.ctypes.data, T12.ctypes.data if T12 is not None else None,
                                                     float(voxel_size), float(sample_proportion), _ptr(scores, c_double),
                                                     _ptr(counts, c_int64)))
        return scores, counts

    def evaluate_device(self, d_xyz1, d_xyz2, d_lm1, d_lm2, d_T21, d_T12, d_counts, voxel_size=1.0, sample_proportion=0.1, stream=0,
                        row1=None, row2=None, n_rows=0):
        """the score counts (S,6) int64 of device-resident maps, written to the device address d_counts on `stream`; no sync.
        row1 / row2 (int64 numpy): series layout, d_xyz1 = the frames array of n_rows rows."""
        self._check(self.lib.lip_batch_evaluate_device(self.h, d_xyz1, d_xyz2, d_lm1, d_lm2, d_T21, d_T12, float(voxel_size),
                                                       float(sample_proportion), d_counts,
                                                       _ptr(row1, c_int64) if row1 is not None else None,
                                                       _ptr(row2, c_int64) if row2 is not None else None, int(n_rows), stream))

    def match(self, xyz1, xyz2, lm1, lm2, refine=True, voxel_size=1.0, sample_proportion=0.5):
        """hot path + refinement + evaluation of the planned batch in one call: (T21, T12, scores (S,3))"""
        xyz1, xyz2, lm1, lm2 = _f64(xyz1), _f64(xyz2), _f64(lm1), _f64(lm2)
        if len(xyz1) != self.n1 or len(xyz2) != self.n2 or len(lm1) != self.sumK or len(lm2) != self.sumK:
            raise ValueError("array sizes do not match the planned batch")
        T21, T12, scores = np.empty(self.n2, np.int64), np.empty(self.n1, np.int64), np.empty((self.S, 3))
        self._check(self.lib.lip_batch_match_host(self.h, xyz1.ctypes.data, xyz2.ctypes.data, lm1.ctypes.data, lm2.ctypes.data,
                                                  int(bool(refine)), float(voxel_size), float(sample_proportion),
                                                  T21.ctypes.data, T12.ctypes.data, _ptr(scores, c_double)))
        return T21, T12, scores

    def refine(self, xyz1, xyz2, T21, T12):
        """double_direction_refinement of the planned batch: returns the refined copy of T21 (int64)"""
        xyz1, xyz2 = _f64(xyz1), _f64(xyz2)
        out = np.array(T21, dtype=np.int64, copy=True)
        T12 = _i64(T12)
        if len(xyz1) != self.n1 or len(xyz2) != self.n2 or len(out) != self.n2 or len(T12) != self.n1:
            raise ValueError("array sizes do not match the planned batch")
        self._check(self.lib.lip_batch_refine_host(self.h, xyz1.ctypes.data, xyz2.ctypes.data, out.ctypes.data, T12.ctypes.data))
        return out

    def interpolate(self, xyz1, xyz2, T21, T12, alpha):
        """interpolation_segment_level of the planned batch: (n1 + n2, 3) float64, per segment pair the n2 rows
        (1 - alpha) * xyz1[T21] + alpha * xyz2 followed by the n1 rows (1 - alpha) * xyz1 + alpha * xyz2[T12]"""
        xyz1, xyz2, T21, T12 = _f64(xyz1), _f64(xyz2), _i64(T21), _i64(T12)
        if len(xyz1) != self.n1 or len(xyz2) != self.n2 or len(T21) != self.n2 or len(T12) != self.n1:
            raise ValueError("array sizes do not match the planned batch")
        out = np.empty((self.n1 + self.n2, 3), np.float64)
        self._check(self.lib.lip_batch_interpolate_host(self.h, xyz1.ctypes.data, xyz2.ctypes.data, T21.ctypes.data, T12.ctypes.data,
                                                        float(alpha), out.ctypes.data))
        return out

    def interpolate_device(self, d_xyz1, d_xyz2, d_T21, d_T12, alpha, d_out, stream=0):
        """the same on device pointers (ints), enqueued on `stream`; no synchronisation"""
        self._check(self.lib.lip_batch_interpolate_device(self.h, d_xyz1, d_xyz2, d_T21, d_T12, float(alpha), d_out, stream))

    def transforms(self, filled=True):
        K = self.sumK
        out = {}
        R21, t21, R12, t12 = (np.empty((K, 3, 3)), np.empty((K, 3)), np.empty((K, 3, 3)), np.empty((K, 3)))
        i21, s21, i12, s12 = (np.empty(K, np.int32) for _ in range(4))
        self._check(self.lib.lip_batch_get_transforms(self.h, int(bool(filled)), _ptr(R21, c_double), _ptr(t21, c_double),
                                                      _ptr(i21, c_int32), _ptr(s21, c_int32), _ptr(R12, c_double),
                                                      _ptr(t12, c_double), _ptr(i12, c_int32), _ptr(s12, c_int32)))
        out.update(R21=R21, t21=t21, iters21=i21, status21=s21, R12=R12, t12=t12, iters12=i12, status12=s12)
        return out

    def landmark_nn(self, side, distances=True):
        """two nearest landmarks (local indices) of every vertex of side 1 / 2 and, unless distances=False, their distances"""
        n = self.n1 if side == 1 else self.n2
        nn, d = np.empty((n, 2), np.int32), (np.empty((n, 2)) if distances else None)
        self._check(self.lib.lip_batch_get_landmark_nn(self.h, side, _ptr(nn, c_int32), _ptr(d, c_double) if distances else None))
        return nn, d

    def counters(self):
        c = np.zeros(8, np.int64)
        self._check(self.lib.lip_batch_get_counters(self.h, _ptr(c, c_int64)))
        return dict(small_icps=int(c[0]), big_icps=int(c[1]), icp_f64_rescans=int(c[2]), map_f64_rescans=int(c[3]),
                    map_retry=int(c[4]), map_far=int(c[5]), launches=int(c[6]), map_ring1_pairs=int(c[7]))

    STAGES = ("landmark_partition_kernel", "icp_schedule_kernel", "icp_group_kernel", "icp_big_kernel", "grid_build_kernel",
              "blend_kernel", "final_map_kernel", "retry_nn_kernel", "far_nn_kernel")

    def pair_counts(self):
        c = np.zeros(3, np.int64)
        self._check(self.lib.lip_batch_get_pair_counts(self.h, _ptr(c, c_int64)))
        return dict(icp_group_kernel=int(c[0]), icp_big_kernel=int(c[1]), final_map_kernel=int(c[2]))

    def set_profiling(self, on=True):
        self._check(self.lib.lip_set_profiling(self.h, int(bool(on))))

    def stage_ms(self):
        ms = np.zeros(len(self.STAGES), np.float32)
        self._check(self.lib.lip_batch_get_stage_ms(self.h, ms.ctypes.data_as(POINTER(ctypes.c_float))))
        return dict(zip(self.STAGES, (float(x) for x in ms)))

    def bin_offsets(self):
        off = np.empty(2 * self.sumK + 1, np.int32)
        self._check(self.lib.lip_batch_get_bin_offsets(self.h, _ptr(off, c_int32)))
        return off

    # ---- single functions ---------------------------------------------------------------------
    def icp(self, src, src_off, tgt, tgt_off):
        src, tgt, src_off, tgt_off = _f64(src), _f64(tgt), _i64(src_off), _i64(tgt_off)
        P = len(src_off) - 1
        R, t = np.empty((P, 3, 3)), np.empty((P, 3))
        it, st = np.empty(P, np.int32), np.empty(P, np.int32)
        self._check(self.lib.lip_icp_host(self.h, P, _ptr(src, c_double), _ptr(src_off, c_int64), _ptr(tgt, c_double),
                                          _ptr(tgt_off, c_int64), _ptr(R, c_double), _ptr(t, c_double),
                                          _ptr(it, c_int32), _ptr(st, c_int32)))
        return R, t, it, st

    def nn(self, src, tgt):
        src, tgt = _f64(src), _f64(tgt)
        d, i = np.empty(len(src)), np.empty(len(src), np.int64)
        self._check(self.lib.lip_nn_host(self.h, _ptr(src, c_double), len(src), _ptr(tgt, c_double), len(tgt),
                                         _ptr(d, c_double), _ptr(i, c_int64)))
        return d, i

    def kabsch(self, A, B):
        A, B = _f64(A), _f64(B)
        R, t, st = np.empty((3, 3)), np.empty(3), np.zeros(1, np.int32)
        self._check(self.lib.lip_kabsch_host(self.h, _ptr(A, c_double), _ptr(B, c_double), len(A), _ptr(R, c_double),
                                             _ptr(t, c_double), _ptr(st, c_int32)))
        return R, t, int(st[0])

    def triangle_angles(self, abc):
        abc = _f64(abc)
        out = np.empty_like(abc)
        self._check(self.lib.lip_triangle_angles_host(self.h, _ptr(abc, c_double), len(abc), _ptr(out, c_double)))
        return out


_default = {}


def default_context(device: int | None = None) -> Context:
    """process-wide context per device (device defaults to LOCAL_RANK, else 0).  Shared by every drop-in function: hold
    ``ctx.lock`` around plan + run + getters."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        if device not in _default:
            _default[device] = Context(device)
        return _default[device]
