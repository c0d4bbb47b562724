"""ctypes access to the C++ host layer (libsasa_b200_host.so, include/sasa_b200.hpp) for tests and tools."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsasa_b200_host.so")
CLI_PATH = os.path.join(HERE, "sasa_b200_cli")
LEVELS = {"atom": 0, "residue": 1, "chain": 2, "protein": 3}

_lib = None


class HostError(RuntimeError):
    """kind = SASACalcError variant name, or "IO"."""

    def __init__(self, text: str):
        super().__init__(text)
        self.kind = text.split(":", 1)[0]


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m rustsasa_b200.build`")
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.sasa_b200_host_pack.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, sz]
        L.sasa_b200_host_pack.restype = vp
        L.sasa_b200_host_pack_atoms.argtypes = [vp]
        L.sasa_b200_host_pack_atoms.restype = sz
        L.sasa_b200_host_pack_segments.argtypes = [vp]
        L.sasa_b200_host_pack_segments.restype = sz
        L.sasa_b200_host_pack_copy.argtypes = [vp, vp, vp, vp, vp]
        L.sasa_b200_host_pack_copy.restype = None
        L.sasa_b200_host_pack_free.argtypes = [vp]
        L.sasa_b200_host_pack_free.restype = None
        L.sasa_b200_host_process_json.argtypes = [C.c_char_p, C.c_int, C.c_float, sz, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_char_p, C.c_char_p, sz, C.c_char_p, sz]
        L.sasa_b200_host_process_json.restype = C.c_long
        L.sasa_b200_host_format.argtypes = [C.c_int, C.c_int, vp, sz, C.c_char_p, sz]
        L.sasa_b200_host_format.restype = C.c_long
        L.sasa_b200_host_writeback.argtypes = [C.c_char_p, C.c_int, vp, sz, C.c_int, C.c_int, C.c_char_p, sz, C.c_char_p, sz]
        L.sasa_b200_host_writeback.restype = C.c_long
        L.sasa_b200_host_serialize_chain_id.argtypes = [C.c_char_p]
        L.sasa_b200_host_serialize_chain_id.restype = C.c_long
        L.sasa_b200_host_get_radius.argtypes = [C.c_char_p, C.c_char_p]
        L.sasa_b200_host_get_radius.restype = C.c_float
        L.sasa_b200_host_flatten.argtypes = [C.c_char_p, C.c_char_p, sz]
        L.sasa_b200_host_flatten.restype = vp
        L.sasa_b200_host_flat_sizes.argtypes = [vp, vp]
        L.sasa_b200_host_flat_sizes.restype = None
        L.sasa_b200_host_flat_copy.argtypes = [vp, vp, vp, vp]
        L.sasa_b200_host_flat_copy.restype = None
        L.sasa_b200_host_flat_free.argtypes = [vp]
        L.sasa_b200_host_flat_free.restype = None
        _lib = L
    return _lib


def flatten(path: str):
    """The C++ reader's hierarchy of a file, one row per atom in hierarchy order (all models, all conformers):
    dict(xyz, occupancy, b_factor, serial, res_serial, model_index, model_serial, hetero, is_h, conformer_index,
    residue_index, chain, icode, altloc, resname, name, element, n_models, n_chains, n_residues, n_conformers)."""
    L = load()
    err = C.create_string_buffer(512)
    h = L.sasa_b200_host_flatten(path.encode(), err, 512)
    if not h:
        raise HostError(err.value.decode())
    try:
        sizes = (C.c_size_t * 6)()
        L.sasa_b200_host_flat_sizes(h, sizes)
        n = sizes[0]
        xyzob = np.zeros((n, 5), np.float64)
        ints = np.zeros((n, 8), np.int64)
        text = C.create_string_buffer(max(1, sizes[5]))
        L.sasa_b200_host_flat_copy(h, xyzob.ctypes.data, ints.ctypes.data, text)
    finally:
        L.sasa_b200_host_flat_free(h)
    rows = [r.split("|") for r in text.raw[:sizes[5]].decode().split("\n")[:-1]] if n else []
    col = lambda k: [r[k] for r in rows]   # noqa: E731
    return dict(xyz=xyzob[:, :3], occupancy=xyzob[:, 3], b_factor=xyzob[:, 4], serial=ints[:, 0], res_serial=ints[:, 1],
                model_index=ints[:, 2], model_serial=ints[:, 3], hetero=ints[:, 4].astype(bool), is_h=ints[:, 5].astype(bool),
                conformer_index=ints[:, 6], residue_index=ints[:, 7], chain=col(0), icode=col(1), altloc=col(2), resname=col(3),
                name=col(4), element=col(5), n_models=int(sizes[1]), n_chains=int(sizes[2]), n_residues=int(sizes[3]),
                n_conformers=int(sizes[4]))


def pack(path: str, level: str = "residue", include_hydrogens=False, include_hetatms=False, allow_vdw_fallback=False,
         read_radii_from_occupancy=False, radii_file: Optional[str] = None):
    """C++ build_atoms_and_mapping on a file -> dict(xyzr, ids, seg_be, seg_polar)."""
    L = load()
    err = C.create_string_buffer(512)
    h = L.sasa_b200_host_pack(path.encode(), LEVELS[level], include_hydrogens, include_hetatms, allow_vdw_fallback,
                              read_radii_from_occupancy, radii_file.encode() if radii_file else None, err, 512)
    if not h:
        raise HostError(err.value.decode())
    try:
        n, g = L.sasa_b200_host_pack_atoms(h), L.sasa_b200_host_pack_segments(h)
        xyzr = np.zeros((n, 4), np.float32)
        ids = np.zeros(n, np.uint64)
        seg = np.zeros((g, 2), np.uint32)
        pol = np.zeros(g, np.uint8)
        L.sasa_b200_host_pack_copy(h, xyzr.ctypes.data, ids.ctypes.data, seg.ctypes.data, pol.ctypes.data)
    finally:
        L.sasa_b200_host_pack_free(h)
    return dict(xyzr=xyzr, ids=ids, seg_be=seg, seg_polar=pol)


def process_json(path: str, level: str = "residue", probe_radius=1.4, n_points=100, include_hydrogens=False,
                 include_hetatms=False, allow_vdw_fallback=False, read_radii_from_occupancy=False,
                 radii_file: Optional[str] = None) -> str:
    """SASAOptions<level>::process through the C++ layer and the GPU -> serde-style JSON text."""
    L = load()
    out = C.create_string_buffer(64 << 20)
    err = C.create_string_buffer(512)
    n = L.sasa_b200_host_process_json(path.encode(), LEVELS[level], probe_radius, n_points, include_hydrogens, include_hetatms,
                                      allow_vdw_fallback, read_radii_from_occupancy,
                                      radii_file.encode() if radii_file else None, out, len(out), err, 512)
    if n < 0:
        raise HostError(err.value.decode())
    return out.raw[:n].decode()


def format_values(values, xml=False, kind="atom") -> str:
    v = np.ascontiguousarray(values, np.float32)
    out = C.create_string_buffer(1 << 20)
    n = load().sasa_b200_host_format(int(xml), LEVELS[kind], v.ctypes.data, v.size, out, len(out))
    return out.raw[:n].decode()


def writeback(path: str, kind: str, values, fmt: str = "pdb", bad_serial: bool = False) -> str:
    """sasa_result_to_protein_object with hand-made values on a file, then the pdbtbx-style text (fmt: "pdb" | "cif")."""
    L = load()
    v = np.ascontiguousarray(values, dtype=np.float32)
    out = C.create_string_buffer(1 << 24)
    err = C.create_string_buffer(512)
    n = L.sasa_b200_host_writeback(path.encode(), LEVELS[kind], v.ctypes.data, v.size, int(bad_serial), 1 if fmt == "cif" else 0,
                                   out, len(out), err, 512)
    if n < 0:
        raise HostError(err.value.decode())
    return out.value.decode()
