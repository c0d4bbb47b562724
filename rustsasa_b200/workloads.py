"""Seeded synthetic workloads of the shapes BASELINE.json names (there is no network for real proteomes).

cfg2  proteome batch   4,400 AlphaFold-sized structures: fragments of real coordinate sets (the committed
                       fixtures extracted from the reference's tests/data), cut to N(2400, 400) atoms
                       clipped to [400, 6000], each under a random rigid transform plus sigma = 0.05 A jitter
cfg3  MD trajectory    one ~5,000-atom protein fragment, frames = base + sigma = 0.3 A displacement
cfg4  large assembly   jittered FCC lattice at protein number density (0.057 / A^3) carved to a sphere
cfg5  capsid           same lattice carved to a spherical shell (30 A thick)
(SURVEY.md 8d).  Everything is deterministic in `seed`.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = os.path.join(ROOT, "tests", "golden", "structures.npz")
SEED = 20261017

# ProtOr radius frequencies measured on example.cif
RADII = np.array([1.88, 1.61, 1.64, 1.42, 1.76, 1.46, 1.77], dtype=np.float32)
RADII_P = np.array([0.399, 0.179, 0.169, 0.142, 0.078, 0.029, 0.005])
RADII_P = RADII_P / RADII_P.sum()


@dataclass
class BatchData:
    xyzr: np.ndarray            # (N, 4) float32
    struct_off: np.ndarray      # (S+1,) uint64
    seg_be: np.ndarray          # (G, 2) uint32, relative to each structure's first atom
    struct_seg_off: np.ndarray  # (S+1,) uint64
    seg_polar: np.ndarray       # (G,) uint8
    name: str = ""

    @property
    def n_atoms(self):
        return int(self.xyzr.shape[0])

    @property
    def n_structures(self):
        return int(self.struct_off.shape[0] - 1)


class _Templates:
    def __init__(self):
        d = np.load(FIXTURES)
        self.names = [str(n) for n in d["names"]]
        self.atom_off = d["atom_off"]
        self.seg_off = d["seg_off"]
        self.xyz = (d["milli"].astype(np.float64) / 1000.0).astype(np.float32)
        self.rad = d["radii_table"][d["radius_idx"]].astype(np.float32)
        self.seg_be = d["seg_be"].astype(np.int64)
        self.polar = d["polar"]
        # templates with enough atoms to cut fragments from
        self.sizes = np.diff(self.atom_off)


_TEMPLATES: Optional[_Templates] = None


def templates() -> _Templates:
    global _TEMPLATES
    if _TEMPLATES is None:
        _TEMPLATES = _Templates()
    return _TEMPLATES


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _fragment(T: _Templates, rng, target_atoms: int):
    """Consecutive non-empty residues of one template totalling ~target_atoms atoms."""
    big = np.nonzero(T.sizes >= target_atoms)[0]
    t = int(rng.choice(big)) if big.size else int(np.argmax(T.sizes))
    s0, s1 = int(T.seg_off[t]), int(T.seg_off[t + 1])
    be = T.seg_be[s0:s1]
    keep = np.nonzero(be[:, 1] > be[:, 0])[0]
    be = be[keep]
    polar = T.polar[s0:s1][keep]
    # residue ranges are contiguous and ascending within a template: choose a start so that enough atoms follow
    ends = be[:, 1]
    total = int(ends[-1])
    if total <= target_atoms:
        i0, i1 = 0, len(be)
    else:
        ok = np.nonzero(total - be[:, 0] >= target_atoms)[0]
        i0 = int(rng.choice(ok))
        i1 = int(np.searchsorted(ends, be[i0, 0] + target_atoms, side="left")) + 1
        i1 = min(i1, len(be))
    a0, a1 = int(be[i0, 0]), int(be[i1 - 1, 1])
    base = int(T.atom_off[t])
    xyz = T.xyz[base + a0: base + a1]
    rad = T.rad[base + a0: base + a1]
    seg = (be[i0:i1] - a0).astype(np.uint32)
    return xyz, rad, seg, polar[i0:i1]


def proteome_batch(n_structures: int = 4400, seed: int = SEED, mean_atoms: float = 2400.0, sd_atoms: float = 400.0,
                   lo: int = 400, hi: int = 6000, jitter: float = 0.05) -> BatchData:
    """cfg2: AlphaFold-like proteome batch."""
    T = templates()
    rng = np.random.default_rng(seed)
    xs, rs, segs, pols, offs, soffs = [], [], [], [], [0], [0]
    for _ in range(n_structures):
        target = int(np.clip(rng.normal(mean_atoms, sd_atoms), lo, hi))
        xyz, rad, seg, pol = _fragment(T, rng, target)
        R = _random_rotation(rng)
        c = xyz.mean(axis=0, dtype=np.float64)
        shift = rng.uniform(-50.0, 50.0, size=3)
        out = (xyz.astype(np.float64) - c) @ R.T + shift + rng.normal(scale=jitter, size=xyz.shape)
        xs.append(out.astype(np.float32))
        rs.append(rad)
        segs.append(seg)
        pols.append(pol)
        offs.append(offs[-1] + xyz.shape[0])
        soffs.append(soffs[-1] + seg.shape[0])
    xyzr = np.concatenate([np.concatenate(xs), np.concatenate(rs)[:, None]], axis=1).astype(np.float32)
    return BatchData(np.ascontiguousarray(xyzr), np.asarray(offs, np.uint64), np.concatenate(segs).astype(np.uint32),
                     np.asarray(soffs, np.uint64), np.concatenate(pols).astype(np.uint8),
                     name=f"cfg2 proteome batch: {n_structures} structures")


@dataclass
class FramesData:
    xyz: np.ndarray        # (F, N, 3) float32
    radii: np.ndarray      # (N,) float32
    seg_be: np.ndarray     # (G, 2) uint32
    seg_polar: np.ndarray  # (G,) uint8
    name: str = ""


def md_trajectory(n_frames: int = 10000, n_atoms: int = 5000, seed: int = SEED, sigma: float = 0.3) -> FramesData:
    """cfg3: frames of one ~n_atoms protein fragment, each = base + N(0, sigma) displacement."""
    T = templates()
    rng = np.random.default_rng(seed + 3)
    xyz, rad, seg, pol = _fragment(T, rng, n_atoms)
    frames = np.empty((n_frames,) + xyz.shape, np.float32)
    step = 256
    for f0 in range(0, n_frames, step):
        f1 = min(f0 + step, n_frames)
        frames[f0:f1] = xyz[None] + rng.normal(scale=sigma, size=(f1 - f0,) + xyz.shape).astype(np.float32)
    return FramesData(frames, rad.copy(), seg, pol.astype(np.uint8), name=f"cfg3 MD: {n_frames} x {xyz.shape[0]} atoms")


def _lattice(rng, r_outer: float, r_inner: float, density: float = 0.057, jitter: float = 0.25):
    a = (4.0 / density) ** (1.0 / 3.0)
    n = int(np.ceil(r_outer / a)) + 1
    g = np.arange(-n, n + 1, dtype=np.float64) * a
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    pts = []
    for b in basis:
        X, Y, Z = np.meshgrid(g + b[0], g + b[1], g + b[2], indexing="ij")
        P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        d = np.linalg.norm(P, axis=1)
        pts.append(P[(d <= r_outer) & (d >= r_inner)])
    P = np.concatenate(pts)
    P = P[rng.permutation(P.shape[0])]
    # order atoms along a space-filling-ish sweep so that "residues" of 8 consecutive atoms are spatially compact
    key = np.lexsort((P[:, 0] // (2 * a), P[:, 1] // (2 * a), P[:, 2] // (2 * a)))
    P = P[key] + rng.normal(scale=jitter, size=P.shape)
    return P.astype(np.float32)


def _single(P: np.ndarray, rng, name: str) -> BatchData:
    n = P.shape[0]
    rad = rng.choice(RADII, size=n, p=RADII_P).astype(np.float32)
    nres = (n + 7) // 8
    b = np.arange(nres, dtype=np.uint32) * 8
    seg = np.stack([b, np.minimum(b + 8, n).astype(np.uint32)], axis=1)
    return BatchData(np.ascontiguousarray(np.concatenate([P, rad[:, None]], axis=1).astype(np.float32)),
                     np.array([0, n], np.uint64), seg, np.array([0, nres], np.uint64),
                     (rng.random(nres) < 0.25).astype(np.uint8), name=name)


def large_assembly(n_atoms: int = 150000, seed: int = SEED) -> BatchData:
    """cfg4: globule of ~n_atoms atoms at protein density."""
    rng = np.random.default_rng(seed + 4)
    r = (3.0 * n_atoms / (4.0 * np.pi * 0.057)) ** (1.0 / 3.0)
    P = _lattice(rng, r, 0.0)
    return _single(P, rng, f"cfg4 assembly: {P.shape[0]} atoms")


def capsid_shell(n_atoms: int = 1000000, thickness: float = 30.0, seed: int = SEED) -> BatchData:
    """cfg5: spherical shell of ~n_atoms atoms, `thickness` A thick."""
    rng = np.random.default_rng(seed + 5)
    # solve 4/3 pi (R^3 - (R-t)^3) * rho = n for R
    lo, hi = thickness, 5000.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if 4.0 / 3.0 * np.pi * (mid ** 3 - (mid - thickness) ** 3) * 0.057 < n_atoms:
            lo = mid
        else:
            hi = mid
    P = _lattice(rng, hi, hi - thickness)
    return _single(P, rng, f"cfg5 capsid shell: {P.shape[0]} atoms")


def describe(b: BatchData) -> str:
    sizes = np.diff(b.struct_off.astype(np.int64))
    return json.dumps(dict(name=b.name, structures=b.n_structures, atoms=b.n_atoms, segments=int(b.seg_be.shape[0]),
                           atoms_min=int(sizes.min()) if sizes.size else 0, atoms_max=int(sizes.max()) if sizes.size else 0,
                           atoms_mean=float(sizes.mean()) if sizes.size else 0.0))
