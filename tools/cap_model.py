#!/usr/bin/env python
"""Offline model of the cap-table occlusion path (analysis aid, not part of the product).

Every neighbour j of atom i occludes a spherical cap of i's test sphere: point p is occluded iff p . v^ < c with
v = c_i - c_j and c = limit / |v|.  A table indexed by (octahedral direction bin of v^, level bin of c) holds two
point masks: `inner` (occluded for every (v^, c) of the bin) and `outer` (occluded for some).  Points in the OR of the
inner masks are done; only (point, neighbour) pairs with the point in outer & ~inner need the exact test.  This
script measures, on real structures, how many points stay undecided and how many exact pair tests remain, and checks
that the decided points agree with the brute-force result.

usage: cap_model.py [N] [L] [names...]
"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tests.golden_data import Golden
from rustsasa_b200.engine import Engine
from tools.phase_model import atoms_entries


def oct_dir(u, v):
    """octahedral (u, v) in [-1, 1]^2 -> unit vector"""
    z = 1.0 - np.abs(u) - np.abs(v)
    x = np.where(z >= 0, u, (1 - np.abs(v)) * np.sign(u + 1e-300))
    y = np.where(z >= 0, v, (1 - np.abs(u)) * np.sign(v + 1e-300))
    d = np.stack([x, y, z], -1)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def oct_uv(d):
    s = np.abs(d).sum(-1)
    u, v = d[..., 0] / s, d[..., 1] / s
    neg = d[..., 2] < 0
    u2 = np.where(neg, (1 - np.abs(v)) * np.where(u >= 0, 1.0, -1.0), u)
    v2 = np.where(neg, (1 - np.abs(u)) * np.where(v >= 0, 1.0, -1.0), v)
    return u2, v2


def build_table(P, N, L, eps_ang=1e-3, eps_c=1e-4):
    """inner/outer boolean tables (L + 2, N, N, n_points)"""
    n = P.shape[0]
    node = np.linspace(-1, 1, N + 1)
    cu = 0.5 * (node[:-1] + node[1:])
    U, V = np.meshgrid(cu, cu, indexing="xy")           # [iv, iu]
    centre = oct_dir(U, V)                                # (N, N, 3)
    rho = np.zeros((N, N))
    for du in (0, 1):
        for dv in (0, 1):
            Uc, Vc = np.meshgrid(node[du:N + du], node[dv:N + dv], indexing="xy")
            corner = oct_dir(Uc, Vc)
            ang = np.arccos(np.clip((centre * corner).sum(-1), -1, 1))
            rho = np.maximum(rho, ang)
    rho = rho + eps_ang
    alpha = np.arccos(np.clip(centre @ P.T, -1, 1))      # (N, N, n)
    dmax = np.cos(np.maximum(alpha - rho[..., None], 0.0))        # max of p.u over the bin
    dmin = np.cos(np.minimum(alpha + rho[..., None], np.pi))      # min of p.u over the bin
    lev = np.linspace(-1, 1, L + 1)
    c_lo = np.concatenate([[-np.inf], lev])              # level 0: c < -1 ; level L+1: c >= 1
    c_hi = np.concatenate([lev, [np.inf]])
    inner = dmax[None] < (c_lo[:, None, None, None] - eps_c)
    outer = dmin[None] < (c_hi[:, None, None, None] + eps_c)
    return inner, outer, rho


ORDER = "list"


def main():
    global ORDER
    import os
    ORDER = os.environ.get("ORDER", "list")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    names = sys.argv[3:] or ["example.cif"]
    P = Engine.sphere_points(100).astype(np.float64)
    inner, outer, rho = build_table(P, N, L)
    print(f"N={N} L={L} table {(L + 2) * N * N * 32 / 1e6:.2f} MB  rho {rho.min():.3f}..{rho.max():.3f}")
    g = Golden()
    tot = dict(atoms=0, undecided=0, pairs=0, exposed=0, k=0, maxpairs=0, ring=0, bad=0, maxlane=0, rounds=0, iters=0)
    hist = []
    for name in names:
        s = g.structure(name)
        for i, v, lim, d2 in atoms_entries(s["xyzr"]):
            k = len(lim)
            tot["atoms"] += 1
            if k == 0:
                continue
            vm = np.sqrt(d2)
            c = lim / vm
            u, w = oct_uv(v / vm[:, None])
            iu = np.clip(((u + 1) * (N / 2)).astype(int), 0, N - 1)
            iv = np.clip(((w + 1) * (N / 2)).astype(int), 0, N - 1)
            l = np.clip(np.floor((c + 1) * (L / 2)).astype(int) + 1, 0, L + 1)
            inn = inner[l, iv, iu]             # (k, n)
            out = outer[l, iv, iu]
            occ = (P @ v.T) < lim[None, :]     # (n, k)
            # safety: inner => occ, occ => outer
            tot["bad"] += int((inn & ~occ.T).sum() + (occ.T & ~out).sum())
            covered = inn.any(0)
            ring = out & ~inn
            tot["ring"] += int(ring.sum())
            tot["undecided"] += int((~covered).sum())
            # rounds of 32 neighbours in list order (ORDER: "list" = as enumerated, "near" = d2 < 16 first, "dist" = sorted)
            o = {"list": np.arange(k), "near": np.argsort(~(d2 < 16.0), kind="stable"), "dist": np.argsort(d2)}[ORDER]
            cov = np.zeros(P.shape[0], bool)
            for r0 in range(0, k, 32):
                sel = o[r0:r0 + 32]
                cov |= inn[sel].any(0)
                rp = ring[sel] & ~cov[None, :]
                npairs = int(rp.sum())
                tot["pairs"] += npairs
                tot["maxlane"] += int(rp.sum(1).max())
                for w0 in range(0, P.shape[0], 32):
                    tot["iters"] += int(rp[:, w0:w0 + 32].sum(1).max())
                tot["rounds"] += 1
                cov |= (rp & occ.T[sel]).any(0)
            assert np.array_equal(~cov, ~occ.any(1))
            tot["exposed"] += int((~occ.any(1)).sum())
            tot["k"] += k
    a = tot["atoms"]
    print(f"atoms {a}  k {tot['k'] / a:.1f}  exposed/atom {tot['exposed'] / a:.2f}  undecided/atom {tot['undecided'] / a:.2f}  "
          f"ring pts/neighbour {tot['ring'] / tot['k']:.2f}  exact pairs/atom {tot['pairs'] / a:.1f} "
          f"max pairs of one lane {tot['maxlane'] / a:.2f}  test rounds/atom {tot['rounds'] / a:.2f}  violations {tot['bad']}")
    print(f"order {ORDER}: ring-loop iterations/atom (sum over words of the longest lane) {tot['iters'] / a:.2f}")


if __name__ == "__main__":
    main()
