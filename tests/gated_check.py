"""Gated single-launch host pipeline (sasa_api.cu: one kernel over the chunk-major queue, the copy stream raising a ready
counter, results written straight into page-locked host buffers) against the per-chunk launches and the oracle.  The
pipeline is chosen per call: page-locked outputs of moderate size take it, pageable outputs take the per-chunk path -- so
the same process can run both on the same inputs.  Run as a script by tests/test_gpu_gated.py with SASA_B200_CHUNK_ATOMS
small enough that the test batches span many chunks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import load  # noqa: E402
from rustsasa_b200 import Engine, workloads as W  # noqa: E402
from rustsasa_b200.engine import index_radii  # noqa: E402

ok = True


def check(name, cond):
    global ok
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    ok = ok and bool(cond)


def same(a, b):
    return all(np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True) for k in ("counts", "atom_sasa", "seg_sasa", "protein")
               if getattr(a, k) is not None)


orc = load(fast=True)
eng = Engine(0)
d = W.proteome_batch(int(os.environ.get("GATED_CHECK_STRUCTURES", "160")), seed=11, mean_atoms=900.0, sd_atoms=250.0, lo=200, hi=2000)
b = eng.batch(d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
h_x = eng.pinned_empty(d.xyzr.shape, np.float32)
h_x[...] = d.xyzr

# float4 form, all four outputs: pinned outputs (gated) == pageable outputs (per-chunk launches) == oracle
pin = b.run_host(h_x, result=b._host_outputs(("counts", "atom", "seg", "protein"), "pinned"))
pag = b.run_host(h_x)
o = orc.run_batch(d.xyzr, d.struct_off, 1.4, 100, seg_be=d.seg_be, struct_seg_off=d.struct_seg_off)
check(f"gated run is one launch ({pin.stats['gpu_launches']}), per-chunk run is many ({pag.stats['gpu_launches']})",
      pin.stats["gpu_launches"] == 1 and pag.stats["gpu_launches"] > 4)
check("float4 form: gated == per-chunk", same(pin, pag))
check("float4 form: gated counts == oracle", np.array_equal(np.asarray(pin.counts), o["counts"]))
check("float4 form: gated residue sums == oracle", np.array_equal(np.asarray(pin.seg_sasa), o["seg"]))

# indexed-radius wire format
pal, idx = index_radii(d.xyzr[:, 3])
h3 = eng.pinned_empty((d.n_atoms, 3), np.float32)
h3[...] = d.xyzr[:, :3]
hi = eng.pinned_empty((d.n_atoms,), np.uint8)
hi[...] = idx
pin2 = b.run_indexed_host(h3, hi, pal, result=b._host_outputs(("counts", "seg"), "pinned"))
check("indexed form: gated one launch", pin2.stats["gpu_launches"] == 1)
check("indexed form: gated == float4 gated", np.array_equal(np.asarray(pin2.counts), np.asarray(pin.counts)) and
      np.array_equal(np.asarray(pin2.seg_sasa), np.asarray(pin.seg_sasa)))

# id classes (duplicate ids inside structures): classes are copied chunk by chunk as well
cls = (np.arange(d.n_atoms, dtype=np.uint32) // 2).astype(np.uint32)
pin3 = b.run_host(h_x, id_class=cls, result=b._host_outputs(("counts",), "pinned"))
pag3 = b.run_host(h_x, id_class=cls, want=("counts",))
check("id classes: gated == per-chunk", np.array_equal(np.asarray(pin3.counts), pag3.counts))

# a non-finite structure in the middle: error code, NaN outputs for that structure only, identical on both paths
bad = d.xyzr.copy()
s_bad = d.n_structures // 2
bad[int(d.struct_off[s_bad]) + 3, 1] = np.nan
h_x[...] = bad
res_p = b._host_outputs(("counts", "atom", "seg"), "pinned")
res_q = b._host_outputs(("counts", "atom", "seg"), "numpy")
codes = []
for res in (res_p, res_q):
    try:
        b.run_host(h_x, want=("counts", "atom", "seg"), result=res)
        codes.append(0)
    except Exception as e:   # SasaB200Error carries the code
        codes.append(getattr(e, "code", -1))
check(f"non-finite input: both paths report it ({codes})", codes[0] == codes[1] and codes[0] != 0)
check("non-finite input: gated == per-chunk (NaN-filled structure, the rest computed)", same(res_p, res_q))
h_x[...] = d.xyzr

# MD frames through the fused unpack
md = W.md_trajectory(n_frames=int(os.environ.get("GATED_CHECK_FRAMES", "48")), n_atoms=1500)
F, NA = md.xyz.shape[:2]
off = np.arange(F + 1, dtype=np.uint64) * NA
G = len(md.seg_be)
bf = eng.batch(off, np.tile(md.seg_be, (F, 1)), np.arange(F + 1, dtype=np.uint64) * G, np.tile(md.seg_polar, F))
hx = eng.pinned_empty((F * NA, 3), np.float32)
hx[...] = md.xyz.reshape(-1, 3)
fp = bf.run_frames_host(hx, md.radii, result=bf._host_outputs(("protein",), "pinned"))
fq = bf.run_frames_host(hx, md.radii)
check(f"frames: gated ({fp.stats['gpu_launches']} launch) == per-chunk ({fq.stats['gpu_launches']})",
      fp.stats["gpu_launches"] == 1 and np.array_equal(np.asarray(fp.protein), fq.protein))

# submit / wait: two jobs of two batches in flight, waited out of order
j1 = b.submit_host(h_x, want=("seg",))
j2 = bf.submit_frames_host(hx, md.radii) if hasattr(bf, "submit_frames_host") else None
r1 = j1.wait()
check("submit / wait: gated job == run_host", np.array_equal(np.asarray(r1.seg_sasa), np.asarray(pin.seg_sasa)))
if j2 is not None:
    r2 = j2.wait()
    check("submit / wait: frames job == run_frames_host", np.array_equal(np.asarray(r2.protein), fq.protein))
b.close()
bf.close()
eng.close()
print("ALL OK" if ok else "FAILURES", flush=True)
sys.exit(0 if ok else 1)
