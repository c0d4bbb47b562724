"""Loader for the committed fixtures under tests/golden/ (made by tools/make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Golden:
    def __init__(self):
        self.vdw = np.load(os.path.join(HERE, "golden", "example_cif_vdw.npz"))
        d = np.load(os.path.join(HERE, "golden", "structures.npz"))
        self._d = {k: d[k] for k in d.files}
        self.names = [str(n) for n in self._d["names"]]
        self.chains = json.loads(str(self._d["chains_json"]))
        self.freesasa = json.loads(str(self._d["freesasa_json"]))
        self._index = {n: i for i, n in enumerate(self.names)}

    def structure(self, name):
        """dict(xyzr (N,4) f32, seg_be (R,2) u32 residue ranges, polar (R,) u8, res_chain (R,), chains [str])."""
        i = self._index[name]
        d = self._d
        a0, a1 = d["atom_off"][i], d["atom_off"][i + 1]
        s0, s1 = d["seg_off"][i], d["seg_off"][i + 1]
        xyz = (d["milli"][a0:a1].astype(np.float64) / 1000.0).astype(np.float32)
        r = d["radii_table"][d["radius_idx"][a0:a1]].astype(np.float32)
        return dict(name=name, xyzr=np.ascontiguousarray(np.concatenate([xyz, r[:, None]], axis=1)),
                    seg_be=np.ascontiguousarray(d["seg_be"][s0:s1]), polar=np.ascontiguousarray(d["polar"][s0:s1]),
                    res_chain=d["res_chain"][s0:s1], chains=self.chains[name],
                    res_names=[str(x) for x in d["res_names"][s0:s1]])

    def sizes(self):
        return {n: int(self._d["atom_off"][i + 1] - self._d["atom_off"][i]) for i, n in enumerate(self.names)}

    def chain_ranges(self, s):
        """Chain-level [begin, end) atom ranges of a structure() dict (atoms of a chain are contiguous)."""
        out = []
        for ci in range(len(s["chains"])):
            idx = np.nonzero(s["res_chain"] == ci)[0]
            be = s["seg_be"][idx]
            out.append((int(be[:, 0].min()), int(be[:, 1].max())))
        return np.asarray(out, dtype=np.uint32).reshape(-1, 2)
