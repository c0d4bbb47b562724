"""Thin Python host layer over the C ABI: Engine (a context on one GPU) and Batch (a reusable topology).

numpy arrays are host buffers (pipelined H2D / kernels / D2H inside the library);
torch CUDA tensors are device buffers (``Batch.run_device``, no copies).  Nothing here computes SASA
on the CPU: every numeric result comes out of libsasa_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib


class SasaB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[sasa_b200 status {code}] {message}")
        self.code = code


class _PinnedBlock:
    """Owns one cudaHostAlloc block and exposes it through the numpy array interface."""

    def __init__(self, L, ptr, nbytes, shape, dtype):
        self._L, self._ptr = L, ptr
        self.__array_interface__ = dict(shape=tuple(int(x) for x in shape), typestr=dtype.str,
                                        data=(ptr, False), version=3)

    def __del__(self):
        try:
            self._L.sasa_b200_free_pinned(self._ptr)
        except Exception:
            pass


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())   # torch tensor


def _stream(stream) -> int:
    """cudaStream_t handle for the device entry points.  None = torch's CURRENT stream, so that work enqueued here
    is ordered with the caller's other torch / NCCL work (the legacy default stream is passed as cudaStreamLegacy = 1,
    because NULL means "the context's own stream" in the C ABI); an int is passed through as a raw handle."""
    if stream is None:
        import torch
        h = int(torch.cuda.current_stream().cuda_stream)
        return h if h != 0 else 1
    return int(stream)


def index_radii(radii):
    """Palette + one-byte indices of a radius column (the indexed-radius wire format), or None when it holds more than 256
    distinct values."""
    pal, idx = np.unique(np.asarray(radii, np.float32), return_inverse=True)
    if pal.shape[0] > 256:
        return None
    return pal.astype(np.float32), idx.astype(np.uint8)


def _np(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class BatchResult:
    counts: Optional[np.ndarray] = None
    atom_sasa: Optional[np.ndarray] = None
    seg_sasa: Optional[np.ndarray] = None
    protein: Optional[np.ndarray] = None   # (S, 3): global, polar, non-polar
    stats: Optional[dict] = None


class Engine:
    """One sasa_b200 context on one CUDA device."""

    def __init__(self, device: int = -1):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.sasa_b200_create(device, C.byref(h))
        if rc != _lib.OK:
            raise SasaB200Error(rc, self._L.sasa_b200_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.sasa_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _lib.OK:
            raise SasaB200Error(rc, self._L.sasa_b200_last_error(self._h).decode())

    @staticmethod
    def sphere_points(n_points: int) -> np.ndarray:
        out = np.empty((n_points, 3), np.float32)
        rc = _lib.load().sasa_b200_sphere_points(n_points, out.ctypes.data)
        if rc != _lib.OK:
            raise SasaB200Error(rc, "sphere_points: invalid argument")
        return out

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array backed by page-locked memory (freed when the array is collected)."""
        dtype = np.dtype(dtype)
        shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        rc = self._L.sasa_b200_alloc_pinned(n, C.byref(p))
        if rc != _lib.OK:
            raise SasaB200Error(rc, self._L.sasa_b200_last_error(None).decode())
        return np.asarray(_PinnedBlock(self._L, p.value, n, shape, dtype))

    def calculate_sasa_internal(self, xyzr, ids=None, probe_radius=1.4, n_points=100, threads=-1,
                                want_counts=False):
        """The reference's raw-atoms entry point (src/lib.rs:249-298) through the C ABI."""
        xyzr = _np(xyzr, np.float32, (-1, 4))
        n = xyzr.shape[0]
        ids = _np(ids, np.uint64)
        out = np.zeros(n, np.float32)
        counts = np.zeros(n, np.uint32) if want_counts else None
        self._check(self._L.sasa_b200_calculate_sasa_internal(
            self._h, _ptr(xyzr), _ptr(ids), n, probe_radius, n_points, threads, _ptr(out), _ptr(counts)))
        return (out, counts) if want_counts else out

    def batch(self, struct_off, seg_be=None, struct_seg_off=None, seg_polar=None) -> "Batch":
        return Batch(self, struct_off, seg_be, struct_seg_off, seg_polar)


class Batch:
    """Topology of a batch of structures (CSR offsets + output segments), reusable across runs."""

    def __init__(self, engine: Engine, struct_off, seg_be=None, struct_seg_off=None, seg_polar=None):
        self.engine = engine
        self._L = engine._L
        self.struct_off = _np(struct_off, np.uint64)
        self.S = self.struct_off.shape[0] - 1
        self.n_atoms = int(self.struct_off[-1]) if self.S >= 0 and self.struct_off.size else 0
        self.seg_be = _np(seg_be, np.uint32, (-1, 2))
        self.struct_seg_off = _np(struct_seg_off, np.uint64)
        self.seg_polar = _np(seg_polar, np.uint8)
        self.n_seg = 0 if self.seg_be is None else self.seg_be.shape[0]
        h = C.c_void_p()
        engine._check(self._L.sasa_b200_batch_create(
            engine._h, _ptr(self.struct_off), max(self.S, 0), _ptr(self.seg_be), _ptr(self.struct_seg_off),
            _ptr(self.seg_polar), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.engine, "_h", None):
            self._L.sasa_b200_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _params(probe_radius, n_points, simd_lanes, flags):
        return _lib.Params(probe_radius, n_points, simd_lanes, -1, flags)

    def _host_outputs(self, want, out):
        res = BatchResult()
        alloc = (lambda shape, dt: self.engine.pinned_empty(shape, dt)) if out == "pinned" else np.zeros
        if "counts" in want:
            res.counts = alloc(self.n_atoms, np.uint32)
        if "atom" in want:
            res.atom_sasa = alloc(self.n_atoms, np.float32)
        if "seg" in want and self.n_seg:
            res.seg_sasa = alloc(self.n_seg, np.float32)
        if "protein" in want:
            res.protein = alloc((max(self.S, 0), 3), np.float32)
        return res

    def run_host(self, xyzr, id_class=None, probe_radius=1.4, n_points=100, simd_lanes=8, flags=0,
                 want=("counts", "atom", "seg", "protein"), result: Optional[BatchResult] = None) -> BatchResult:
        xyzr = _np(xyzr, np.float32, (-1, 4))
        assert xyzr.shape[0] == self.n_atoms, "xyzr does not match struct_off"
        id_class = _np(id_class, np.uint32)
        res = result if result is not None else self._host_outputs(want, "numpy")
        outs = _lib.Outputs(_ptr(res.counts), _ptr(res.atom_sasa), _ptr(res.seg_sasa), _ptr(res.protein))
        prm = self._params(probe_radius, n_points, simd_lanes, flags)
        st = _lib.Stats()
        self.engine._check(self._L.sasa_b200_batch_run_host(self._h, _ptr(xyzr), _ptr(id_class), C.byref(prm),
                                                           C.byref(outs), C.byref(st)))
        res.stats = st.as_dict()
        return res

    def submit_host(self, xyzr, id_class=None, probe_radius=1.4, n_points=100, simd_lanes=8, flags=0,
                    want=("counts", "atom", "seg", "protein"), result: Optional[BatchResult] = None) -> "Job":
        """Asynchronous run_host: returns at once with a Job; `Job.wait()` blocks until the outputs are in host memory.
        Inputs and outputs should be page-locked (``Engine.pinned_empty``) and must stay alive until wait returns."""
        xyzr = _np(xyzr, np.float32, (-1, 4))
        assert xyzr.shape[0] == self.n_atoms, "xyzr does not match struct_off"
        id_class = _np(id_class, np.uint32)
        res = result if result is not None else self._host_outputs(want, "pinned")
        outs = _lib.Outputs(_ptr(res.counts), _ptr(res.atom_sasa), _ptr(res.seg_sasa), _ptr(res.protein))
        prm = self._params(probe_radius, n_points, simd_lanes, flags)
        h = C.c_void_p()
        self.engine._check(self._L.sasa_b200_batch_submit_host(self._h, _ptr(xyzr), _ptr(id_class), C.byref(prm),
                                                              C.byref(outs), C.byref(h)))
        return Job(self, h, res, (xyzr, id_class))

    def run_frames_host(self, xyz, radii, probe_radius=1.4, n_points=100, simd_lanes=8, flags=0,
                        want=("protein",), result: Optional[BatchResult] = None) -> BatchResult:
        """MD form: xyz (F, N, 3) float32 + radii (N,) shared by all frames."""
        xyz = _np(xyz, np.float32, (-1, 3))
        radii = _np(radii, np.float32)
        assert xyz.shape[0] == self.n_atoms
        res = result if result is not None else self._host_outputs(want, "numpy")
        outs = _lib.Outputs(_ptr(res.counts), _ptr(res.atom_sasa), _ptr(res.seg_sasa), _ptr(res.protein))
        prm = self._params(probe_radius, n_points, simd_lanes, flags)
        st = _lib.Stats()
        self.engine._check(self._L.sasa_b200_batch_run_frames_host(self._h, _ptr(xyz), _ptr(radii), C.byref(prm),
                                                                  C.byref(outs), C.byref(st)))
        res.stats = st.as_dict()
        return res

    def run_indexed_host(self, xyz, radius_index, palette, id_class=None, probe_radius=1.4, n_points=100, simd_lanes=8, flags=0,
                         want=("counts", "atom", "seg", "protein"), result: Optional[BatchResult] = None) -> BatchResult:
        """Indexed-radius form (13 B/atom on the wire): xyz (N, 3) float32, radius_index (N,) uint8 into palette (<= 256,)."""
        xyz = _np(xyz, np.float32, (-1, 3))
        radius_index = _np(radius_index, np.uint8)
        palette = _np(palette, np.float32)
        assert xyz.shape[0] == self.n_atoms == radius_index.shape[0]
        id_class = _np(id_class, np.uint32)
        res = result if result is not None else self._host_outputs(want, "numpy")
        outs = _lib.Outputs(_ptr(res.counts), _ptr(res.atom_sasa), _ptr(res.seg_sasa), _ptr(res.protein))
        prm = self._params(probe_radius, n_points, simd_lanes, flags)
        st = _lib.Stats()
        self.engine._check(self._L.sasa_b200_batch_run_indexed_host(self._h, _ptr(xyz), _ptr(radius_index), _ptr(palette),
                                                                   palette.shape[0], _ptr(id_class), C.byref(prm), C.byref(outs),
                                                                   C.byref(st)))
        res.stats = st.as_dict()
        return res

    def run_device(self, d_xyzr, d_id_class=None, probe_radius=1.4, n_points=100, simd_lanes=8, flags=0,
                   counts=None, atom_sasa=None, seg_sasa=None, protein=None, stream=None):
        """All tensors are torch CUDA tensors; enqueues on `stream` (None = torch's current stream, else a raw
        cudaStream_t handle) and returns without synchronising."""
        outs = _lib.Outputs(_ptr(counts), _ptr(atom_sasa), _ptr(seg_sasa), _ptr(protein))
        prm = self._params(probe_radius, n_points, simd_lanes, flags)
        self.engine._check(self._L.sasa_b200_batch_run_device(self._h, _ptr(d_xyzr), _ptr(d_id_class), C.byref(prm),
                                                             C.byref(outs), _stream(stream)))

    def run_atom_range_device(self, d_xyzr, rank: int, n_ranks: int, d_id_class=None, probe_radius=1.4, n_points=100,
                              simd_lanes=8, counts=None, atom_sasa=None, stream=None):
        """Atom-range split (cfg5): slice `rank` of `n_ranks` of every structure's cell-sorted atom order; the
        output tensors are zero outside the slice, ready for an all-reduce(SUM) across ranks."""
        prm = self._params(probe_radius, n_points, simd_lanes, 0)
        self.engine._check(self._L.sasa_b200_batch_run_atom_range_device(
            self._h, _ptr(d_xyzr), _ptr(d_id_class), C.byref(prm), rank, n_ranks, _ptr(counts), _ptr(atom_sasa),
            _stream(stream)))

    def run_atom_range_peers_device(self, d_xyzr, rank: int, n_ranks: int, peer_counts=None, peer_atom_sasa=None, d_id_class=None,
                                    probe_radius=1.4, n_points=100, simd_lanes=8, stream=None):
        """Atom-range split with the exchange fused into the kernel: peer_counts / peer_atom_sasa are sequences of n_ranks
        raw device pointers (ints), valid on this device, of every rank's output vector (symmetric-memory buffer_ptrs)."""
        prm = self._params(probe_radius, n_points, simd_lanes, 0)
        pc = (C.c_void_p * n_ranks)(*[int(x) for x in peer_counts]) if peer_counts is not None else None
        pa = (C.c_void_p * n_ranks)(*[int(x) for x in peer_atom_sasa]) if peer_atom_sasa is not None else None
        self.engine._check(self._L.sasa_b200_batch_run_atom_range_peers_device(
            self._h, _ptr(d_xyzr), _ptr(d_id_class), C.byref(prm), rank, n_ranks, pc, pa, _stream(stream)))

    def run_atom_range_host(self, xyzr, rank: int, n_ranks: int, id_class=None, probe_radius=1.4, n_points=100,
                            simd_lanes=8) -> BatchResult:
        xyzr = _np(xyzr, np.float32, (-1, 4))
        assert xyzr.shape[0] == self.n_atoms, "xyzr does not match struct_off"
        id_class = _np(id_class, np.uint32)
        res = BatchResult(counts=np.zeros(self.n_atoms, np.uint32), atom_sasa=np.zeros(self.n_atoms, np.float32))
        prm = self._params(probe_radius, n_points, simd_lanes, 0)
        st = _lib.Stats()
        self.engine._check(self._L.sasa_b200_batch_run_atom_range_host(
            self._h, _ptr(xyzr), _ptr(id_class), C.byref(prm), rank, n_ranks, _ptr(res.counts), _ptr(res.atom_sasa),
            C.byref(st)))
        res.stats = st.as_dict()
        return res

    def reduce_device(self, d_atom_sasa, seg_sasa=None, protein=None, stream=None):
        """Level sums of a finished per-atom SASA vector (torch CUDA tensors)."""
        self.engine._check(self._L.sasa_b200_batch_reduce_device(self._h, _ptr(d_atom_sasa), _ptr(seg_sasa), _ptr(protein),
                                                                _stream(stream)))

    def sync(self) -> dict:
        st = _lib.Stats()
        self.engine._check(self._L.sasa_b200_batch_sync(self._h, C.byref(st)))
        return st.as_dict()


class Job:
    """A host-buffer run in flight (sasa_b200_batch_submit_host .. sasa_b200_job_wait)."""

    def __init__(self, batch: Batch, handle, result: BatchResult, keep):
        self.batch, self._h, self.result, self._keep = batch, handle, result, keep

    def wait(self) -> BatchResult:
        if self._h is None:
            return self.result
        st = _lib.Stats()
        h, self._h = self._h, None
        self.batch.engine._check(self.batch._L.sasa_b200_job_wait(h, C.byref(st)))
        self.result.stats = st.as_dict()
        self._keep = None
        return self.result
