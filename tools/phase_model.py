#!/usr/bin/env python
"""Offline cost model of the per-atom occlusion pipeline (phase 1 depth m, entry ordering, tile shapes).
Pure analysis aid: builds the point x neighbour occlusion matrix of sample atoms with numpy and counts the
warp instructions the tight kernel would issue under different policies.  Not part of the product."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from tests.golden_data import Golden
from rustsasa_b200.engine import Engine

def points(n):
    return Engine.sphere_points(n)

def atoms_entries(xyzr, probe=1.4):
    from scipy.spatial import cKDTree
    t = cKDTree(xyzr[:, :3])
    rmax = xyzr[:, 3].max()
    pairs = t.query_ball_point(xyzr[:, :3], 2 * rmax + 2 * probe + 1e-3)
    for i, js in enumerate(pairs):
        js = np.array([j for j in js if j != i])
        if js.size == 0:
            yield i, np.zeros((0, 3)), np.zeros(0), np.zeros(0); continue
        v = xyzr[i, :3] - xyzr[js, :3]
        d2 = (v * v).sum(1)
        keep = d2 <= (xyzr[i, 3] + xyzr[js, 3] + 2 * probe + 1e-3) ** 2
        v, d2, js = v[keep], d2[keep], js[keep]
        r = xyzr[i, 3] + probe
        lim = ((xyzr[js, 3] + probe) ** 2 - d2 - r * r) / (2 * r)
        yield i, v, lim, d2

def tile_cost(ns, nent, setup=22, per_it=8):
    if ns == 0 or nent <= 0: return 0
    c = 0
    for b in range(0, ns, 32):
        n = min(32, ns - b)
        G = 1 << int(np.ceil(np.log2(n))) if n > 1 else 1
        c += setup + int(np.ceil(nent / (32 // G))) * per_it
    return c

def main():
    g = Golden()
    P = points(100)[:96]
    names = sys.argv[1:] or ["example.cif", "151L_H3.pdb"]
    res = {}
    for name in names:
        s = g.structure(name)
        for i, v, lim, d2 in atoms_entries(s["xyzr"]):
            k = len(lim)
            if k == 0: continue
            occ = (P @ v.T) < lim[None, :]          # (96, k)
            orders = {
                "near4": np.argsort(~(d2 < 16.0), kind="stable"),
                "dist": np.argsort(d2),
                "limit": np.argsort(-lim),
            }
            for oname, o in orders.items():
                oc = occ[:, o]
                alive = ~np.logical_or.accumulate(oc, axis=1)       # alive[:, m-1] after m entries
                nfront = int((d2 < 16.0).sum())
                for pol in ("cur", 4, 8, 12, 16, 20, 24, "all"):
                    if pol == "cur": m = min(k, min(max(nfront, 4), 16))
                    elif pol == "all": m = k
                    else: m = min(k, pol)
                    ns = int(alive[:, m - 1].sum()) if m > 0 else 96
                    c = 13 * m + 6 + (tile_cost(ns, k - m) + (12 if ns and m < k else 0))
                    key = (oname, pol)
                    a = res.setdefault(key, [0, 0, 0])
                    a[0] += c; a[1] += ns; a[2] += 1
    print(f"{'order':8s} {'m policy':>8s} {'instr/atom':>10s} {'survivors':>10s}")
    for (oname, pol), (c, ns, n) in sorted(res.items(), key=lambda kv: kv[1][0]):
        print(f"{oname:8s} {str(pol):>8s} {c / n:10.1f} {ns / n:10.2f}")

if __name__ == "__main__":
    main()
