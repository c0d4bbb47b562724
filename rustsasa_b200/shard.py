"""Multi-GPU sharding of the hot path: one process per GPU, structures (or MD frames) split over ranks.

Structures are independent, so a batch shards with NO data-path collective: every rank takes
a contiguous, cost-balanced range of structures, runs it through its own engine (own context, own stream
queue with H2D / kernel / D2H overlap) and writes a disjoint slice of the result arrays.  What crosses
ranks is only the final host-side gather of results to rank 0 -- over a gloo (CPU) process group, because
the results already sit in host memory; NCCL is used only by the atom-range split of one giant assembly
(`run_atom_range`, BASELINE cfg5), where per-rank partial count vectors are summed on the devices.

This is what the reference's directory mode (rayon `par_iter` over files, src/main.rs:375, :439) becomes
on an 8-GPU box.  Nothing here computes SASA: `compute` is the engine call (`Batch.run_host`), injected so
that the CPU-only gloo tests can drive the same code with the oracle as the checker.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def structure_cost(struct_off: np.ndarray, n_points: int = 100) -> np.ndarray:
    """Relative cost of each structure: atoms x (n_points + setup).  Neighbour density is uniform in
    proteins, so the point-neighbour test count is proportional to atoms x n_points; the constant covers
    the per-atom gather / sort work that does not scale with n_points."""
    n = np.diff(np.asarray(struct_off, dtype=np.int64))
    return n.astype(np.float64) * (float(n_points) + 40.0)


def partition_structures(struct_off: Sequence[int], n_parts: int, n_points: int = 100,
                         cost: Optional[np.ndarray] = None) -> np.ndarray:
    """Contiguous partition of S structures into `n_parts` ranges with balanced cumulative cost.

    Returns `bounds` of length n_parts + 1 (bounds[0] = 0, bounds[-1] = S); rank r owns structures
    [bounds[r], bounds[r + 1]).  Contiguity keeps every rank's inputs and outputs one slice of the CSR
    arrays.  Ranges may be empty when S < n_parts."""
    struct_off = np.asarray(struct_off, dtype=np.int64)
    S = struct_off.shape[0] - 1
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    if S <= 0:
        return np.zeros(n_parts + 1, dtype=np.int64)
    c = structure_cost(struct_off, n_points) if cost is None else np.asarray(cost, dtype=np.float64)
    if c.shape[0] != S:
        raise ValueError("cost must have one entry per structure")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    bounds = np.zeros(n_parts + 1, dtype=np.int64)
    bounds[-1] = S
    for r in range(1, n_parts):
        target = total * r / n_parts
        # the boundary whose cumulative cost is closest to the ideal split point
        k = int(np.searchsorted(cum, target, side="left"))
        if k > 0 and (k > S or abs(cum[k - 1] - target) <= abs(cum[k] - target)):
            k -= 1
        bounds[r] = min(max(k, bounds[r - 1]), S)
    return bounds


@dataclass
class Shard:
    """One rank's slice of a CSR batch (offsets rebased to the slice)."""
    s0: int
    s1: int
    a0: int
    a1: int
    g0: int
    g1: int
    xyzr: np.ndarray
    struct_off: np.ndarray
    seg_be: Optional[np.ndarray]
    struct_seg_off: Optional[np.ndarray]
    seg_polar: Optional[np.ndarray]
    id_class: Optional[np.ndarray]


def take_shard(bounds: np.ndarray, rank: int, xyzr: np.ndarray, struct_off: np.ndarray,
               seg_be: Optional[np.ndarray] = None, struct_seg_off: Optional[np.ndarray] = None,
               seg_polar: Optional[np.ndarray] = None, id_class: Optional[np.ndarray] = None) -> Shard:
    s0, s1 = int(bounds[rank]), int(bounds[rank + 1])
    struct_off = np.asarray(struct_off, dtype=np.uint64)
    a0, a1 = int(struct_off[s0]), int(struct_off[s1])
    g0 = g1 = 0
    sb = so = sp = None
    if seg_be is not None:
        struct_seg_off = np.asarray(struct_seg_off, dtype=np.uint64)
        g0, g1 = int(struct_seg_off[s0]), int(struct_seg_off[s1])
        sb = np.ascontiguousarray(np.asarray(seg_be, np.uint32).reshape(-1, 2)[g0:g1])
        so = (struct_seg_off[s0:s1 + 1] - np.uint64(g0)).astype(np.uint64)
        sp = None if seg_polar is None else np.ascontiguousarray(np.asarray(seg_polar, np.uint8)[g0:g1])
    cls = None
    if id_class is not None:
        # classes only matter through equality inside a structure: the slice can be used as is
        cls = np.ascontiguousarray(np.asarray(id_class, np.uint32)[a0:a1])
    return Shard(s0, s1, a0, a1, g0, g1, np.asarray(xyzr, np.float32).reshape(-1, 4)[a0:a1],
                 (struct_off[s0:s1 + 1] - np.uint64(a0)).astype(np.uint64), sb, so, sp, cls)


@dataclass
class ShardedResult:
    """Full-batch results, valid on the destination rank (None elsewhere)."""
    counts: Optional[np.ndarray] = None
    atom_sasa: Optional[np.ndarray] = None
    seg_sasa: Optional[np.ndarray] = None
    protein: Optional[np.ndarray] = None
    bounds: Optional[np.ndarray] = None


_FIELDS = (("counts", np.uint32, "atoms"), ("atom_sasa", np.float32, "atoms"), ("seg_sasa", np.float32, "segs"),
           ("protein", np.float32, "structs3"))


def _gather_slices(local: Optional[np.ndarray], sizes: List[int], dtype, group, rank: int, world: int, dst: int):
    """Variable-size gather of one result array to `dst` over a host (gloo) group: every rank knows every
    slice size from the deterministic partition, so no size exchange is needed."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return None if local is None else np.asarray(local, dtype=dtype).reshape(-1).copy()
    mine = torch.from_numpy(np.ascontiguousarray(np.asarray(local, dtype=dtype).reshape(-1)).view(np.uint8).copy())
    if rank == dst:
        bufs = [torch.empty(sizes[r] * np.dtype(dtype).itemsize, dtype=torch.uint8) for r in range(world)]
        reqs = [dist.irecv(bufs[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
                for r in range(world) if r != dst and sizes[r]]
        bufs[dst] = mine
        for q in reqs:
            q.wait()
        return np.concatenate([b.numpy().view(dtype) for b in bufs])
    if sizes[rank]:
        dist.send(mine, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    return None


def run_sharded(compute: Callable[[Shard], "object"], xyzr: np.ndarray, struct_off: np.ndarray,
                seg_be: Optional[np.ndarray] = None, struct_seg_off: Optional[np.ndarray] = None,
                seg_polar: Optional[np.ndarray] = None, id_class: Optional[np.ndarray] = None,
                n_points: int = 100, want: Tuple[str, ...] = ("counts", "atom_sasa", "seg_sasa", "protein"),
                rank: Optional[int] = None, world: Optional[int] = None, host_group=None, dst: int = 0) -> ShardedResult:
    """Run one CSR batch over all ranks of the job and gather the results on `dst`.

    `compute(shard)` runs the shard on this rank's GPU and returns an object with the attributes named in
    `want` (a `BatchResult`).  `host_group` is a gloo process group for the host-side gather (required when
    the default group is NCCL; None = the default group)."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    struct_off = np.asarray(struct_off, dtype=np.uint64)
    bounds = partition_structures(struct_off, world, n_points)
    sh = take_shard(bounds, rank, xyzr, struct_off, seg_be, struct_seg_off, seg_polar, id_class)
    res = compute(sh) if sh.s1 > sh.s0 else None
    seg_off = None if seg_be is None else np.asarray(struct_seg_off, dtype=np.uint64)
    sizes = {
        "atoms": [int(struct_off[bounds[r + 1]] - struct_off[bounds[r]]) for r in range(world)],
        "segs": [0] * world if seg_off is None else [int(seg_off[bounds[r + 1]] - seg_off[bounds[r]]) for r in range(world)],
        "structs3": [3 * int(bounds[r + 1] - bounds[r]) for r in range(world)],
    }
    out = ShardedResult(bounds=bounds)
    for name, dtype, kind in _FIELDS:
        if name not in want or (kind == "segs" and seg_off is None):
            continue
        local = getattr(res, name) if res is not None else np.zeros(0, dtype)
        if local is None:
            raise ValueError(f"compute() did not return `{name}`")
        got = _gather_slices(local, sizes[kind], dtype, host_group, rank, world, dst)
        if rank == dst:
            setattr(out, name, got.reshape(-1, 3) if kind == "structs3" else got)
    return out


def run_atom_range(compute_range: Callable[[int, int], Tuple["object", "object"]], rank: Optional[int] = None,
                   world: Optional[int] = None, group=None):
    """Atom-range split of ONE large structure (BASELINE cfg5): every rank holds all atoms, rebuilds the cell
    list redundantly and evaluates slice `rank` of the cell-sorted atom order; the per-rank partial vectors
    (zero outside the slice) are then summed across ranks -- the path's one real exchange step.

    `compute_range(rank, world)` returns (counts, atom_sasa) as torch tensors on this rank's device (int32 /
    float32, zero outside the slice); the tensors are all-reduced in place (NCCL on GPUs) and returned.
    Adding zeros is exact in both integer and IEEE arithmetic, so the result is bit-identical to a
    single-GPU run."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    counts, atom_sasa = compute_range(rank, world)
    if world > 1:
        w1 = dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w2 = dist.all_reduce(atom_sasa, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w1.wait()
        w2.wait()
    return counts, atom_sasa


class PeerVectors:
    """Per-atom output vectors of an atom-range split in torch symmetric memory: every rank allocates the same two vectors,
    the rendezvous maps all of them into every rank's address space (NVLink / NVSwitch peer memory), and each rank's atoms
    kernel writes the values of the atoms it owns straight into all of them (`Batch.run_atom_range_peers_device`) -- the
    all-reduce of `run_atom_range` and its zero-fill disappear; what remains is a barrier on either side."""

    def __init__(self, n_atoms: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.counts = symm_mem.empty(n_atoms, dtype=torch.int32, device="cuda")
        self.atom_sasa = symm_mem.empty(n_atoms, dtype=torch.float32, device="cuda")
        self._hc = symm_mem.rendezvous(self.counts, self.group)
        self._ha = symm_mem.rendezvous(self.atom_sasa, self.group)
        self.rank, self.world = self._hc.rank, self._hc.world_size
        self.count_ptrs = list(self._hc.buffer_ptrs)
        self.atom_ptrs = list(self._ha.buffer_ptrs)

    def barrier(self):
        """Cross-rank barrier on the current stream (signal pads of the symmetric allocation)."""
        self._hc.barrier()


def run_atom_range_peers(launch: Callable[[int, int, "PeerVectors"], None], peers: "PeerVectors"):
    """One step of the atom-range split with peer writes: barrier (nobody still reads the vectors), this rank's kernels
    (`launch(rank, world, peers)` enqueues `run_atom_range_peers_device` on the current stream), barrier (every rank's values
    have landed).  Returns (counts, atom_sasa): complete on every rank, bit-identical to a single-GPU run."""
    peers.barrier()
    launch(peers.rank, peers.world, peers)
    peers.barrier()
    return peers.counts, peers.atom_sasa
