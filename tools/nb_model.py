#!/usr/bin/env python
"""Offline workload statistics of the fused kernel's neighbour search + cap path on the cfg2 generator (analysis aid, not
part of the product): atoms per occupied cell, candidates per cell, 32-candidate steps per atom, neighbours per atom,
cap rounds, and what an "inner rows first" early exit would save.

usage: nb_model.py [n_structures]
"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from rustsasa_b200 import workloads as W
from rustsasa_b200.engine import Engine
from tools.cap_model import build_table, oct_uv

PROBE = 1.4
SLACK = 1e-3


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    data = W.proteome_batch(ns, seed=W.SEED)
    P = Engine.sphere_points(100).astype(np.float64)
    N, L = 128, 64
    inner, outer, _ = build_table(P, N, L)
    off = data.struct_off.astype(np.int64)
    T = dict(atoms=0, cells=0, cand_cell=0, steps=0, steps_inner=0, k=0, rounds=0, full_inner9=0, full_all=0, full_r32=0,
             k_le32=0, k_le64=0, zero=0, k9=0, cells1=0, cells2=0, pairs_atoms=0, k_inner=0, full_near=0)
    hist_cell = np.zeros(16, int)
    for s in range(ns):
        a = data.xyzr[off[s]:off[s + 1]].astype(np.float64)
        n = a.shape[0]
        rmax = a[:, 3].max()
        c = 0.5 * (2 * rmax + 2 * PROBE + SLACK) * 1.0002
        mn = a[:, :3].min(0)
        ci = np.floor((a[:, :3] - mn) / c).astype(int)
        dims = ci.max(0) + 1
        cid = (ci[:, 2] * dims[1] + ci[:, 1]) * dims[0] + ci[:, 0]
        order = np.argsort(cid, kind="stable")
        a, ci, cid = a[order], ci[order], cid[order]
        ncell = int(dims.prod())
        cnt = np.bincount(cid, minlength=ncell)
        start = np.concatenate([[0], np.cumsum(cnt)])
        occ_cells = np.nonzero(cnt)[0]
        T["cells"] += occ_cells.size
        for v in cnt[occ_cells]:
            hist_cell[min(v, 15)] += 1
        grid = cnt.reshape(dims[2], dims[1], dims[0])
        for cc in occ_cells:
            cx, cy, cz = cc % dims[0], (cc // dims[0]) % dims[1], cc // (dims[0] * dims[1])
            z0, z1 = max(cz - 2, 0), min(cz + 2, dims[2] - 1)
            y0, y1 = max(cy - 2, 0), min(cy + 2, dims[1] - 1)
            x0, x1 = max(cx - 2, 0), min(cx + 2, dims[0] - 1)
            tot = int(grid[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1].sum())
            zi0, zi1 = max(cz - 1, 0), min(cz + 1, dims[2] - 1)
            yi0, yi1 = max(cy - 1, 0), min(cy + 1, dims[1] - 1)
            tot_in = int(grid[zi0:zi1 + 1, yi0:yi1 + 1, x0:x1 + 1].sum())
            m = cnt[cc]
            T["cand_cell"] += tot
            T["steps"] += m * ((tot + 31) // 32)
            T["steps_inner"] += m * ((tot_in + 31) // 32)
            T["pairs_atoms"] += (m // 2) * 2
        # neighbours per atom (tight cutoff)
        from scipy.spatial import cKDTree
        tree = cKDTree(a[:, :3])
        lists = tree.query_ball_point(a[:, :3], 2 * rmax + 2 * PROBE + SLACK)
        for i, js in enumerate(lists):
            js = np.array([j for j in js if j != i], int)
            T["atoms"] += 1
            if js.size == 0:
                T["zero"] += 1
                continue
            v = a[i, :3] - a[js, :3]
            d2 = (v * v).sum(1)
            keep = d2 <= (a[i, 3] + a[js, 3] + 2 * PROBE + SLACK) ** 2
            js, v, d2 = js[keep], v[keep], d2[keep]
            k = js.size
            T["k"] += k
            T["rounds"] += (k + 31) // 32
            T["k_le32"] += k <= 32
            T["k_le64"] += k <= 64
            if k == 0:
                continue
            r = a[i, 3] + PROBE
            lim = ((a[js, 3] + PROBE) ** 2 - d2 - r * r) / (2 * r)
            vm = np.sqrt(d2)
            cc_ = lim / vm
            u, w = oct_uv(v / vm[:, None])
            iu = np.clip(((u + 1) * (N / 2)).astype(int), 0, N - 1)
            iv = np.clip(((w + 1) * (N / 2)).astype(int), 0, N - 1)
            l = np.clip(np.floor((cc_ + 1) * (L / 2)).astype(int) + 1, 0, L + 1)
            inn = inner[l, iv, iu]
            full_all = inn.any(0).all()
            T["full_all"] += full_all
            # neighbours in the 9 inner rows (|dy| <= 1, |dz| <= 1 cells)
            dcell = ci[js] - ci[i]
            in9 = (np.abs(dcell[:, 1]) <= 1) & (np.abs(dcell[:, 2]) <= 1)
            T["k9"] += int(in9.sum())
            if in9.any():
                T["full_inner9"] += inn[in9].any(0).all()
            near = d2 < 5.0 ** 2
            if near.any():
                T["full_near"] += inn[near].any(0).all()
            # first 32 in list order (cell-sorted order ~ js ascending)
            o = np.argsort(js)
            T["full_r32"] += inn[o[:32]].any(0).all()
    A = T["atoms"]
    print(f"structures {ns} atoms {A}  occupied cells {T['cells']}  atoms/cell {A / T['cells']:.2f}")
    print("atoms-per-cell histogram (cells):", hist_cell.tolist())
    print(f"candidates per occupied cell {T['cand_cell'] / T['cells']:.1f}; gather steps/atom {T['steps'] / A:.2f}; "
          f"inner-9-row steps/atom {T['steps_inner'] / A:.2f}; atoms in pairs {T['pairs_atoms'] / A:.3f}")
    print(f"k {T['k'] / A:.2f}  rounds/atom {T['rounds'] / A:.3f}  k<=32 {T['k_le32'] / A:.3f}  k<=64 {T['k_le64'] / A:.3f}  k in inner 9 rows {T['k9'] / A:.2f}")
    print(f"fully covered by inner masks: all neighbours {T['full_all'] / A:.3f}; inner-9-row neighbours only {T['full_inner9'] / A:.3f}; "
          f"d<5A only {T['full_near'] / A:.3f}; first 32 (sorted order) {T['full_r32'] / A:.3f}")


if __name__ == "__main__":
    main()
