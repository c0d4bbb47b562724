"""Host-side mirror of the reference's public interface for the hot path.

Same names, argument meaning and error behaviour as the crate (rust-sasa 0.9.2):

* ``calculate_sasa_internal(atoms, probe_radius, n_points, threads)`` -- src/lib.rs:249-298
* ``SASAOptions`` builder with ``AtomLevel | ResidueLevel | ChainLevel | ProteinLevel`` and
  ``process()`` -- src/options.rs:60-70, :496-618; defaults probe 1.4, n_points 100, threads -1,
  hydrogens / HETATM / vdW fallback / occupancy radii all off (:498-510)
* result structs ``ResidueResult``, ``ChainResult``, ``ProteinResult`` -- src/structures/atomic.rs:26-70
* ``SASACalcError`` variants -- src/options.rs:466-494 (raised host-side, before / after the numeric core)

The numeric core is the C ABI (include/sasa_b200.h); there is no CPU path here.
``process_many`` is the batched form the CLI's directory mode (src/main.rs:342-480) maps onto.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .engine import Batch, Engine
from .structure import (PackedStructure, SASACalcError, Structure, build_atoms_and_mapping,  # noqa: F401
                        load_radii_from_file, read_structure)

_ENGINE: Optional[Engine] = None


def default_engine() -> Engine:
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine()
    return _ENGINE


@dataclass
class Atom:
    """src/structures/atomic.rs:13-24."""
    position: Sequence[float]
    radius: float
    id: int
    parent_id: Optional[int] = None


@dataclass
class ChainResult:
    name: str
    value: float


@dataclass
class ResidueResult:
    serial_number: int
    insertion_code: str
    value: float
    name: str
    is_polar: bool
    chain_id: str


@dataclass
class ProteinResult:
    global_total: float
    polar_total: float
    non_polar_total: float


class AtomLevel:
    name = "atom"


class ResidueLevel:
    name = "residue"


class ChainLevel:
    name = "chain"


class ProteinLevel:
    name = "protein"


def calculate_sasa_internal(atoms, probe_radius: float = 1.4, n_points: int = 100, threads: int = -1,
                            engine: Optional[Engine] = None) -> np.ndarray:
    """SASA per atom.  ``atoms`` is a sequence of :class:`Atom` or an (N, 4) float32 array of x, y, z, radius
    (ids then default to the index).  ``threads`` is accepted and ignored (the GPU path has no thread pool)."""
    eng = engine or default_engine()
    if isinstance(atoms, np.ndarray):
        return eng.calculate_sasa_internal(atoms, None, probe_radius, n_points, threads)
    atoms = list(atoms)
    if not atoms:
        return np.zeros(0, np.float32)
    xyzr = np.array([[a.position[0], a.position[1], a.position[2], a.radius] for a in atoms], dtype=np.float32)
    ids = np.array([a.id & 0xFFFFFFFFFFFFFFFF for a in atoms], dtype=np.uint64)
    return eng.calculate_sasa_internal(xyzr, ids, probe_radius, n_points, threads)


def _dense_classes(ids: np.ndarray) -> Optional[np.ndarray]:
    uniq, inv = np.unique(ids, return_inverse=True)
    return None if uniq.shape[0] == ids.shape[0] else inv.astype(np.uint32)


class SASAOptions:
    """``SASAOptions::<Level>::new()`` with the reference's ``with_*`` builder methods."""

    def __init__(self, level=ResidueLevel):
        self.level = level
        self.probe_radius = 1.4
        self.n_points = 100
        self.threads = -1
        self.include_hydrogens = False
        self.radii_config = None
        self.allow_vdw_fallback = False
        self.include_hetatms = False
        self.read_radii_from_occupancy = False

    # convenience constructors (src/options.rs:565-587)
    @classmethod
    def atom_level(cls): return cls(AtomLevel)
    @classmethod
    def residue_level(cls): return cls(ResidueLevel)
    @classmethod
    def chain_level(cls): return cls(ChainLevel)
    @classmethod
    def protein_level(cls): return cls(ProteinLevel)

    def with_probe_radius(self, radius): self.probe_radius = float(radius); return self
    def with_include_hetatms(self, v): self.include_hetatms = bool(v); return self
    def with_n_points(self, points): self.n_points = int(points); return self
    def with_read_radii_from_occupancy(self, v): self.read_radii_from_occupancy = bool(v); return self
    def with_threads(self, threads): self.threads = int(threads); return self
    def with_include_hydrogens(self, v): self.include_hydrogens = bool(v); return self
    def with_allow_vdw_fallback(self, v): self.allow_vdw_fallback = bool(v); return self

    def with_radii_file(self, path):
        try:
            self.radii_config = load_radii_from_file(path)
        except OSError as e:
            raise SASACalcError("RadiiFileLoad", f"Failed to load radii file: {e}") from e
        return self

    def _pack(self, st: Structure) -> PackedStructure:
        return build_atoms_and_mapping(st, self.level.name, self.radii_config, self.allow_vdw_fallback,
                                       self.include_hydrogens, self.include_hetatms, self.read_radii_from_occupancy)

    def process(self, st: Structure, engine: Optional[Engine] = None):
        """One structure -> the level's result type (src/options.rs:606-618)."""
        return self.process_many([st], engine)[0]

    def process_many(self, structures: Sequence[Structure], engine: Optional[Engine] = None) -> List:
        """Batched form: all structures go through ONE pipelined batch call.  A structure whose extraction
        fails is reported as its ``SASACalcError`` in the result list (directory mode logs and continues,
        src/main.rs:447-453)."""
        eng = engine or default_engine()
        packed: List[Optional[PackedStructure]] = []
        results: List = [None] * len(structures)
        for i, st in enumerate(structures):
            try:
                packed.append(self._pack(st))
            except SASACalcError as e:
                packed.append(None)
                results[i] = e
        good = [i for i, p in enumerate(packed) if p is not None]
        if not good:
            return results
        xyzr = np.concatenate([packed[i].xyzr for i in good]) if good else np.zeros((0, 4), np.float32)
        struct_off = np.cumsum([0] + [packed[i].xyzr.shape[0] for i in good]).astype(np.uint64)
        seg_be = np.concatenate([packed[i].seg_be for i in good]).astype(np.uint32)
        seg_off = np.cumsum([0] + [packed[i].seg_be.shape[0] for i in good]).astype(np.uint64)
        polar = np.concatenate([packed[i].seg_polar for i in good]).astype(np.uint8)
        # Atom.id equality classes (atoms with equal ids never occlude each other); None when all distinct
        cls_parts, base, any_dup = [], 0, False
        for i in good:
            c = _dense_classes(packed[i].ids)
            any_dup = any_dup or c is not None
            n = packed[i].ids.shape[0]
            cls_parts.append((np.arange(n, dtype=np.uint32) if c is None else c) + np.uint32(base))
            base += n
        id_class = np.concatenate(cls_parts) if any_dup else None
        level = self.level.name
        want = {"atom": ("atom",), "residue": ("seg",), "chain": ("seg",), "protein": ("protein",)}[level]
        batch = Batch(eng, struct_off, seg_be if seg_be.shape[0] else None, seg_off if seg_be.shape[0] else None,
                      polar if seg_be.shape[0] else None)
        try:
            res = batch.run_host(xyzr, id_class, self.probe_radius, self.n_points, want=want)
        finally:
            batch.close()
        for k, i in enumerate(good):
            p = packed[i]
            a0, a1 = int(struct_off[k]), int(struct_off[k + 1])
            g0, g1 = int(seg_off[k]), int(seg_off[k + 1])
            if level == "atom":
                results[i] = res.atom_sasa[a0:a1].copy()
            elif level == "residue":
                results[i] = [ResidueResult(m["serial_number"], m["insertion_code"], float(v), m["name"],
                                            m["is_polar"], m["chain_id"])
                              for m, v in zip(p.seg_meta, res.seg_sasa[g0:g1])] if g1 > g0 else []
            elif level == "chain":
                results[i] = [ChainResult(m["name"], float(v)) for m, v in zip(p.seg_meta, res.seg_sasa[g0:g1])] \
                    if g1 > g0 else []
            else:
                t = res.protein[k]
                results[i] = ProteinResult(float(t[0]), float(t[1]), float(t[2]))
        return results
