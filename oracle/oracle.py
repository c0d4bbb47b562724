"""ctypes loader for oracle/sasa_oracle.c (test infrastructure only; see the C header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _build(target: str, *extra: str) -> None:
    subprocess.run(["make", "-C", _HERE, target, *extra], check=True, stdout=subprocess.DEVNULL)


def _cpu_tag() -> str:
    """Short hash of this host's CPU model + ISA flags: the -march=native build is only valid on the CPU it was
    built on, and built .so files travel from the build container to the GPU box."""
    import hashlib
    sig = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith(("model name", "flags")):
                    sig += line
                if line.startswith("flags"):
                    break
    except OSError:
        pass
    return hashlib.sha1(sig.encode()).hexdigest()[:10]


class Oracle:
    def __init__(self, fast: bool = False):
        name = f"liboracle_fast_{_cpu_tag()}.so" if fast else "liboracle.so"
        path = os.path.join(_HERE, name)
        src = os.path.join(_HERE, "sasa_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            if fast:
                _build("fast", f"FAST_OUT={name}")
            else:
                _build("all")
        self.lib = L = C.CDLL(path)
        fp, u32p, u64p, u8p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                               C.POINTER(C.c_uint8))
        L.oracle_sphere_points.argtypes = [C.c_size_t, fp, fp, fp]
        L.oracle_sphere_points.restype = None
        L.oracle_calculate_sasa_internal.argtypes = [fp, u64p, C.c_size_t, C.c_float, C.c_size_t, C.c_int, C.c_int,
                                                     fp, u32p, u32p, u32p, C.c_double]
        L.oracle_calculate_sasa_internal.restype = C.c_int
        L.oracle_neighbor_lists.argtypes = [fp, u64p, C.c_size_t, C.c_float, C.c_float, C.c_float, u32p, u32p, fp]
        L.oracle_neighbor_lists.restype = C.c_long
        L.oracle_segment_sums.argtypes = [fp, u32p, C.c_size_t, fp]
        L.oracle_segment_sums.restype = None
        L.oracle_protein_totals.argtypes = [fp, C.c_size_t, u32p, u8p, C.c_size_t, fp]
        L.oracle_protein_totals.restype = None
        L.oracle_run_batch.argtypes = [fp, u64p, C.c_size_t, C.c_float, C.c_size_t, C.c_int, C.c_int, fp, u32p,
                                       u32p, u64p, fp]
        L.oracle_run_batch.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int

    @staticmethod
    def _p(a, ty):
        return None if a is None else a.ctypes.data_as(C.POINTER(ty))

    def sphere_points(self, n: int) -> np.ndarray:
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        self.lib.oracle_sphere_points(n, self._p(x, C.c_float), self._p(y, C.c_float), self._p(z, C.c_float))
        return np.stack([x, y, z], axis=1)

    def calculate_sasa_internal(self, xyzr, probe=1.4, n_points=100, threads=1, lanes=8, ids=None,
                                want_k=False, boundary_tol=None):
        """Returns dict(sasa, counts[, k][, boundary]) for one structure (rows A1-A4)."""
        xyzr = np.ascontiguousarray(xyzr, dtype=np.float32).reshape(-1, 4)
        n = xyzr.shape[0]
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        sasa = np.zeros(n, np.float32)
        counts = np.zeros(n, np.uint32)
        k = np.zeros(n, np.uint32) if want_k else None
        bnd = np.zeros(n, np.uint32) if boundary_tol is not None else None
        rc = self.lib.oracle_calculate_sasa_internal(
            self._p(xyzr, C.c_float), self._p(ids, C.c_uint64), n, probe, n_points, threads, lanes,
            self._p(sasa, C.c_float), self._p(counts, C.c_uint32), self._p(k, C.c_uint32), self._p(bnd, C.c_uint32),
            float(boundary_tol or 0.0))
        if rc:
            raise ValueError(f"oracle: non-finite or invalid input (rc={rc})")
        out = dict(sasa=sasa, counts=counts)
        if want_k:
            out["k"] = k
        if bnd is not None:
            out["boundary"] = bnd
        return out

    def neighbor_lists(self, xyzr, probe, max_radius, cell_size=None, ids=None):
        xyzr = np.ascontiguousarray(xyzr, dtype=np.float32).reshape(-1, 4)
        n = xyzr.shape[0]
        cell = float(np.float32(probe) + np.float32(max_radius)) if cell_size is None else cell_size
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        off = np.zeros(n + 1, np.uint32)
        total = self.lib.oracle_neighbor_lists(self._p(xyzr, C.c_float), self._p(ids, C.c_uint64), n, probe,
                                               max_radius, cell, self._p(off, C.c_uint32), None, None)
        if total < 0:
            raise ValueError("oracle: invalid input")
        idx = np.zeros(max(total, 1), np.uint32)
        thr = np.zeros(max(total, 1), np.float32)
        self.lib.oracle_neighbor_lists(self._p(xyzr, C.c_float), self._p(ids, C.c_uint64), n, probe, max_radius,
                                       cell, self._p(off, C.c_uint32), self._p(idx, C.c_uint32),
                                       self._p(thr, C.c_float))
        return [idx[off[i]:off[i + 1]].copy() for i in range(n)], [thr[off[i]:off[i + 1]].copy() for i in range(n)]

    def segment_sums(self, atom_sasa, seg_be):
        atom_sasa = np.ascontiguousarray(atom_sasa, np.float32)
        seg_be = np.ascontiguousarray(seg_be, np.uint32).reshape(-1, 2)
        out = np.zeros(seg_be.shape[0], np.float32)
        self.lib.oracle_segment_sums(self._p(atom_sasa, C.c_float), self._p(seg_be, C.c_uint32), seg_be.shape[0],
                                     self._p(out, C.c_float))
        return out

    def protein_totals(self, atom_sasa, seg_be, seg_polar):
        atom_sasa = np.ascontiguousarray(atom_sasa, np.float32)
        seg_be = np.ascontiguousarray(seg_be, np.uint32).reshape(-1, 2)
        seg_polar = np.ascontiguousarray(seg_polar, np.uint8)
        out = np.zeros(3, np.float32)
        self.lib.oracle_protein_totals(self._p(atom_sasa, C.c_float), atom_sasa.shape[0],
                                       self._p(seg_be, C.c_uint32), self._p(seg_polar, C.c_uint8), seg_be.shape[0],
                                       self._p(out, C.c_float))
        return out

    def run_batch(self, xyzr, struct_off, probe=1.4, n_points=100, lanes=8, n_threads=0, seg_be=None,
                  struct_seg_off=None, want_counts=True):
        """Directory-mode analogue: structures over host threads, each single-threaded."""
        xyzr = np.ascontiguousarray(xyzr, dtype=np.float32).reshape(-1, 4)
        struct_off = np.ascontiguousarray(struct_off, dtype=np.uint64)
        n = xyzr.shape[0]
        sasa = np.zeros(n, np.float32)
        counts = np.zeros(n, np.uint32) if want_counts else None
        out_seg = None
        if seg_be is not None:
            seg_be = np.ascontiguousarray(seg_be, np.uint32).reshape(-1, 2)
            struct_seg_off = np.ascontiguousarray(struct_seg_off, np.uint64)
            out_seg = np.zeros(seg_be.shape[0], np.float32)
        rc = self.lib.oracle_run_batch(self._p(xyzr, C.c_float), self._p(struct_off, C.c_uint64),
                                       struct_off.shape[0] - 1, probe, n_points, lanes, n_threads,
                                       self._p(sasa, C.c_float), self._p(counts, C.c_uint32),
                                       self._p(seg_be, C.c_uint32), self._p(struct_seg_off, C.c_uint64),
                                       self._p(out_seg, C.c_float))
        if rc:
            raise ValueError(f"oracle: invalid input (rc={rc})")
        return dict(sasa=sasa, counts=counts, seg=out_seg)

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())


_CACHE = {}


def load(fast: bool = False) -> Oracle:
    if fast not in _CACHE:
        _CACHE[fast] = Oracle(fast)
    return _CACHE[fast]
