"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): per-atom exposed-point counts EXACT; per-atom SASA, residue / chain /
protein sums within 1e-4 relative or 1e-3 A^2 absolute -- in fact asserted bit-identical here, because
the kernels use the reference's arithmetic and summation order.  Run on the B200 box: pytest -m gpu.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PROBE = 1.4
REL, ABS = 1e-4, 1e-3     # the north-star floating-point tolerance


@pytest.fixture(scope="module")
def eng():
    from rustsasa_b200 import Engine
    e = Engine()
    yield e
    e.close()


def close_enough(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= np.maximum(ABS, REL * np.abs(b)))


def test_sphere_points_match_oracle(eng, oracle):
    for n in (1, 7, 100, 960, 5000):
        assert np.array_equal(eng.sphere_points(n), oracle.sphere_points(n))


def test_golden_vector_exact(eng, golden):
    """The reference's own KAT (tests/units.rs:17-43, tests/common/data.rs): exact integer counts."""
    g = golden.vdw
    sasa, counts = eng.calculate_sasa_internal(g["xyzr"], None, PROBE, 100, -1, want_counts=True)
    assert np.array_equal(counts, g["gold_counts"])
    assert int(counts.sum()) == 16843
    assert np.abs(sasa - g["gold_sasa"]).max() < 1e-5          # reference asserts epsilon = 25.0
    assert close_enough(sasa, g["gold_sasa"])


@pytest.mark.parametrize("name", ["example.cif", "151L_H3.pdb", "bad_seqadv_1A06.pdb", "2drt", "4xfj"])
@pytest.mark.parametrize("n_points", [100, 960])
def test_structure_all_levels_bit_exact(eng, oracle, golden, name, n_points):
    s = golden.structure(name)
    o = oracle.calculate_sasa_internal(s["xyzr"], PROBE, n_points)
    n = s["xyzr"].shape[0]
    # residue level
    b = eng.batch([0, n], s["seg_be"], [0, len(s["seg_be"])], s["polar"])
    r = b.run_host(s["xyzr"], n_points=n_points)
    assert np.array_equal(r.counts, o["counts"])
    assert np.array_equal(r.atom_sasa, o["sasa"])
    assert np.array_equal(r.seg_sasa, oracle.segment_sums(o["sasa"], s["seg_be"]))
    assert np.array_equal(r.protein[0], oracle.protein_totals(o["sasa"], s["seg_be"], s["polar"]))
    assert r.stats["streamed_atoms"] == 0
    b.close()
    # chain level
    ch = golden.chain_ranges(s)
    b = eng.batch([0, n], ch, [0, len(ch)])
    r = b.run_host(s["xyzr"], n_points=n_points, want=("seg",))
    want = oracle.segment_sums(o["sasa"], ch)
    assert close_enough(r.seg_sasa, want)
    if n <= 16384:
        assert np.array_equal(r.seg_sasa, want)
    b.close()


def test_reference_totals(eng, golden):
    """tests/units.rs:117 (20131.227 @ 960 points) and Appendix A values."""
    s = golden.structure("example.cif")
    n = s["xyzr"].shape[0]
    b = eng.batch([0, n], s["seg_be"], [0, len(s["seg_be"])], s["polar"])
    r = b.run_host(s["xyzr"], n_points=960, want=("counts", "protein"))
    assert int(r.counts.sum()) == 157016
    assert r.protein[0, 0] == np.float32(20131.227)
    r = b.run_host(s["xyzr"], n_points=100, want=("counts", "protein"))
    assert int(r.counts.sum()) == 16326 and abs(float(r.protein[0, 0]) - 20097.68) < 0.01
    b.close()


@pytest.mark.parametrize("n_points,lanes", [(1, 8), (3, 4), (37, 8), (50, 8), (100, 4), (100, 16), (128, 8), (129, 16),
                                            (200, 8), (200, 16), (333, 16), (1000, 16)])
def test_point_counts_and_lane_rules(eng, oracle, golden, n_points, lanes):
    """Body / tail split for every mirrored reference build (src/lib.rs:104-106, :162-218)."""
    s = golden.structure("151L_H3.pdb")
    n = s["xyzr"].shape[0]
    b = eng.batch([0, n])
    r = b.run_host(s["xyzr"], n_points=n_points, simd_lanes=lanes, want=("counts", "atom"))
    o = oracle.calculate_sasa_internal(s["xyzr"], PROBE, n_points, lanes=lanes)
    assert np.array_equal(r.counts, o["counts"])
    assert np.array_equal(r.atom_sasa, o["sasa"])
    b.close()


@pytest.mark.parametrize("probe", [0.0, 0.5, 1.4, 2.0, 3.5])
def test_probe_radii(eng, oracle, golden, probe):
    s = golden.structure("bad_seqadv_1A06.pdb")
    sasa, counts = eng.calculate_sasa_internal(s["xyzr"], None, probe, 100, -1, want_counts=True)
    o = oracle.calculate_sasa_internal(s["xyzr"], probe, 100)
    assert np.array_equal(counts, o["counts"]) and np.array_equal(sasa, o["sasa"])


def _concat(structs):
    xyzr = np.concatenate([s["xyzr"] for s in structs])
    off = np.cumsum([0] + [s["xyzr"].shape[0] for s in structs]).astype(np.uint64)
    seg = np.concatenate([s["seg_be"] for s in structs])
    soff = np.cumsum([0] + [len(s["seg_be"]) for s in structs]).astype(np.uint64)
    pol = np.concatenate([s["polar"] for s in structs])
    return xyzr, off, seg, soff, pol


def test_whole_quality_set_in_one_batch(eng, golden):
    """All 90 fixture structures (317 .. 32k atoms: every shared-memory bucket and the large path) in ONE
    batch call vs the oracle per structure; also restates tests/quality.rs RMSE on the GPU results."""
    from oracle import load
    fast = load(fast=True)
    structs = [golden.structure(n) for n in golden.names]
    xyzr, off, seg, soff, pol = _concat(structs)
    b = eng.batch(off, seg, soff, pol)
    r = b.run_host(xyzr)
    fs, ours = [], []
    for k, s in enumerate(structs):
        a0, a1, g0, g1 = int(off[k]), int(off[k + 1]), int(soff[k]), int(soff[k + 1])
        o = fast.calculate_sasa_internal(s["xyzr"], PROBE, 100)
        assert np.array_equal(r.counts[a0:a1], o["counts"]), s["name"]
        assert np.array_equal(r.atom_sasa[a0:a1], o["sasa"]), s["name"]
        assert np.array_equal(r.seg_sasa[g0:g1], fast.segment_sums(o["sasa"], s["seg_be"])), s["name"]
        assert close_enough(r.protein[k], fast.protein_totals(o["sasa"], s["seg_be"], s["polar"])), s["name"]
        if s["name"] in golden.freesasa:
            segv = r.seg_sasa[g0:g1].astype(np.float64)
            for ci, label in enumerate(s["chains"]):
                if label in golden.freesasa[s["name"]]:
                    fs.append(golden.freesasa[s["name"]][label])
                    ours.append(segv[s["res_chain"] == ci].sum())
    rmse = float(np.sqrt(np.mean((np.array(fs) - np.array(ours)) ** 2)))
    assert rmse <= 43.99 + 20.0 and abs(rmse - 43.99) < 0.5
    b.close()


def test_streaming_kernel_and_boundary_stats(eng, oracle, golden):
    from rustsasa_b200 import _lib
    s = golden.structure("example.cif")
    n = s["xyzr"].shape[0]
    b = eng.batch([0, n])
    fastp = b.run_host(s["xyzr"], want=("counts",))
    stream = b.run_host(s["xyzr"], want=("counts",), flags=_lib.FLAG_FORCE_STREAMING)
    assert np.array_equal(fastp.counts, stream.counts)
    assert stream.stats["streamed_atoms"] == n and fastp.stats["streamed_atoms"] == 0
    st = b.run_host(s["xyzr"], want=("counts",), flags=_lib.FLAG_BOUNDARY_STATS)
    o = oracle.calculate_sasa_internal(s["xyzr"], PROBE, 100, boundary_tol=1e-5)
    assert np.array_equal(st.counts, o["counts"])
    assert st.stats["boundary_points"] == int(o["boundary"].sum())
    b.close()


def test_duplicate_ids(eng, oracle):
    xyzr = np.array([[0, 0, 0, 2.0], [1.0, 0, 0, 2.0], [0.5, 1.0, 0.0, 1.5]], np.float32)
    ids = np.array([7, 7, 9], np.uint64)
    sasa, counts = eng.calculate_sasa_internal(xyzr, ids, PROBE, 100, -1, want_counts=True)
    o = oracle.calculate_sasa_internal(xyzr, PROBE, 100, ids=ids)
    assert np.array_equal(counts, o["counts"]) and np.array_equal(sasa, o["sasa"])
    sasa2, counts2 = eng.calculate_sasa_internal(xyzr, np.array([1, 2, 3], np.uint64), PROBE, 100, -1, want_counts=True)
    assert np.array_equal(counts2, oracle.calculate_sasa_internal(xyzr, PROBE, 100)["counts"])
    assert counts2[0] < counts[0]


def test_multi_model_id_classes_in_batch(eng, oracle, golden):
    """Two overlaid copies of a structure sharing ids (a multi-model file): copies never occlude each other."""
    s = golden.structure("2drt")
    n = s["xyzr"].shape[0]
    xyzr = np.concatenate([s["xyzr"], s["xyzr"] + np.array([0.3, 0, 0, 0], np.float32)])
    cls = np.concatenate([np.arange(n), np.arange(n)]).astype(np.uint32)
    b = eng.batch([0, 2 * n])
    r = b.run_host(xyzr, cls, want=("counts",))
    o = oracle.calculate_sasa_internal(xyzr, PROBE, 100, ids=cls.astype(np.uint64))
    assert np.array_equal(r.counts, o["counts"])
    b.close()


def _area(r):
    return 4.0 * np.pi * r * r


def test_sanity_closed_forms(eng):
    """tests/sanity.rs:20-157 through the raw-atoms entry point, 50,000 points, 0.5 %."""
    from rustsasa_b200 import Atom, calculate_sasa_internal
    n, tol = 50000, 0.005
    A = lambda x, y, z, r, i: Atom([x, y, z], r, i)   # noqa: E731
    one = calculate_sasa_internal([A(0, 0, 0, 2.0, 1)], PROBE, n, 1, engine=eng)
    assert abs(one[0] / _area(3.4) - 1) < tol
    two = calculate_sasa_internal([A(0, 0, 0, 2.0, 1), A(10, 0, 0, 2.0, 2)], PROBE, n, 1, engine=eng)
    assert abs(two[0] / _area(3.4) - 1) < tol and abs(two[1] / _area(3.4) - 1) < tol
    assert abs(float(two.sum()) / (2 * _area(3.4)) - 1) < tol
    r, d = 3.4, 4.0
    exp = _area(r) - 2 * np.pi * r * (r - d / 2)
    ov = calculate_sasa_internal([A(0, 0, 0, 2.0, 1), A(d, 0, 0, 2.0, 2)], PROBE, n, 1, engine=eng)
    assert abs(ov[0] / exp - 1) < tol and abs(ov[1] / exp - 1) < tol
    cont = calculate_sasa_internal([A(0, 0, 0, 10.0, 1), A(2, 0, 0, 2.0, 2)], PROBE, n, 1, engine=eng)
    assert abs(cont[0] / _area(11.4) - 1) < tol and cont[1] <= tol
    d = 5.0
    cap = 2 * np.pi * r * (r - d / 2)
    ch = calculate_sasa_internal([A(0, 0, 0, 2.0, 1), A(d, 0, 0, 2.0, 2), A(2 * d, 0, 0, 2.0, 3)], PROBE, n, 1, engine=eng)
    assert abs(ch[0] / (_area(r) - cap) - 1) < tol and abs(ch[2] / (_area(r) - cap) - 1) < tol
    assert abs(ch[1] / (_area(r) - 2 * cap) - 1) < tol
    assert calculate_sasa_internal([], PROBE, n, 1, engine=eng).shape == (0,)


def test_edge_cases(eng, oracle):
    from rustsasa_b200 import SasaB200Error
    # empty batch, empty structures between real ones, single atom, coincident atoms
    b = eng.batch([0])
    r = b.run_host(np.zeros((0, 4), np.float32))
    assert r.counts.shape == (0,) and r.protein.shape == (0, 3)
    b.close()
    xyzr = np.array([[1, 2, 3, 1.5], [5, 5, 5, 1.8], [5, 5, 5, 1.8], [5.5, 5, 5, 1.2]], np.float32)
    b = eng.batch([0, 0, 1, 1, 4, 4], np.array([[0, 1], [0, 0], [0, 3], [1, 3]], np.uint32), [0, 0, 1, 1, 4, 4],
                  np.array([1, 0, 1, 0], np.uint8))
    r = b.run_host(xyzr)
    o1 = oracle.calculate_sasa_internal(xyzr[:1], PROBE, 100)
    o2 = oracle.calculate_sasa_internal(xyzr[1:], PROBE, 100)
    assert list(r.counts[:1]) == [100] and np.array_equal(r.counts[1:], o2["counts"])
    assert np.array_equal(r.atom_sasa, np.concatenate([o1["sasa"], o2["sasa"]]))
    assert r.seg_sasa[1] == 0.0                                   # empty residue reports 0.0 (tests/io.rs:164-224)
    assert r.seg_sasa[3] == oracle.segment_sums(o2["sasa"], [[1, 3]])[0]
    assert np.all(r.protein[0] == 0) and np.all(r.protein[1, 0] == o1["sasa"][0])
    b.close()
    # non-finite input: the reference panics; the C ABI reports status 4 and does not crash
    bad = xyzr.copy()
    bad[2, 1] = np.nan
    with pytest.raises(SasaB200Error) as ei:
        eng.calculate_sasa_internal(bad)
    assert ei.value.code == 4
    assert eng.calculate_sasa_internal(xyzr).shape == (4,)        # the context stays usable
    # invalid arguments
    with pytest.raises(SasaB200Error):
        eng.batch([0, 4]).run_host(xyzr, n_points=0)
    with pytest.raises(SasaB200Error):
        eng.batch([0, 4]).run_host(xyzr, simd_lanes=5)
    with pytest.raises(SasaB200Error):
        eng.batch([0, 4], np.array([[0, 9]], np.uint32), [0, 1])


def test_dense_cluster_overflows_to_streaming(eng, oracle):
    """More than 128 neighbours per atom (never a protein, but legal input): the staging area overflows and
    the streaming kernel takes over with identical results."""
    rng = np.random.default_rng(5)
    xyz = rng.normal(scale=2.0, size=(600, 3)).astype(np.float32)
    xyzr = np.concatenate([xyz, np.full((600, 1), 1.7, np.float32)], axis=1)
    b = eng.batch([0, 600])
    r = b.run_host(xyzr, want=("counts",))
    assert np.array_equal(r.counts, oracle.calculate_sasa_internal(xyzr, PROBE, 100)["counts"])
    assert r.stats["streamed_atoms"] > 0
    b.close()


def test_sparse_and_elongated_structures(eng, oracle):
    """Cell grid growth: a 2,000 A long string of atoms and widely scattered atoms."""
    rng = np.random.default_rng(11)
    n = 1500
    line = np.stack([np.linspace(0, 2000, n), rng.normal(size=n), rng.normal(size=n)], axis=1)
    scat = rng.uniform(-400, 400, size=(n, 3))
    for P in (line, scat):
        xyzr = np.concatenate([P, rng.choice([1.4, 1.6, 1.9], size=(n, 1))], axis=1).astype(np.float32)
        sasa, counts = eng.calculate_sasa_internal(xyzr, None, PROBE, 100, -1, want_counts=True)
        assert np.array_equal(counts, oracle.calculate_sasa_internal(xyzr, PROBE, 100)["counts"])


def test_frames_entry_point(eng, oracle):
    from rustsasa_b200 import workloads as W
    md = W.md_trajectory(n_frames=6, n_atoms=1200)
    F, N = md.xyz.shape[:2]
    off = (np.arange(F + 1) * N).astype(np.uint64)
    seg = np.tile(md.seg_be, (F, 1))
    soff = (np.arange(F + 1) * len(md.seg_be)).astype(np.uint64)
    b = eng.batch(off, seg, soff, np.tile(md.seg_polar, F))
    r = b.run_frames_host(md.xyz, md.radii, want=("counts", "seg", "protein"))
    for f in range(F):
        xyzr = np.concatenate([md.xyz[f], md.radii[:, None]], axis=1)
        o = oracle.calculate_sasa_internal(xyzr, PROBE, 100)
        assert np.array_equal(r.counts[f * N:(f + 1) * N], o["counts"])
        assert np.array_equal(r.protein[f], oracle.protein_totals(o["sasa"], md.seg_be, md.seg_polar))
    b.close()


def test_large_structure_path(eng, golden):
    """cfg4-style single large structure (global cell list) vs the oracle, atom level."""
    from oracle import load
    from rustsasa_b200 import workloads as W
    fast = load(fast=True)
    a = W.large_assembly(40000)
    sasa, counts = eng.calculate_sasa_internal(a.xyzr, None, PROBE, 100, -1, want_counts=True)
    o = fast.calculate_sasa_internal(a.xyzr, PROBE, 100, threads=-1)
    assert np.array_equal(counts, o["counts"]) and np.array_equal(sasa, o["sasa"])
    c = W.capsid_shell(60000)
    b = eng.batch(c.struct_off, c.seg_be, c.struct_seg_off, c.seg_polar)
    r = b.run_host(c.xyzr, n_points=960)
    o = fast.calculate_sasa_internal(c.xyzr, PROBE, 960, threads=-1)
    assert np.array_equal(r.counts, o["counts"])
    assert np.array_equal(r.seg_sasa, fast.segment_sums(o["sasa"], c.seg_be))
    assert close_enough(r.protein[0], fast.protein_totals(o["sasa"], c.seg_be, c.seg_polar))
    b.close()


def test_proteome_batch_sample_and_invariants(eng):
    """cfg2 shape at reduced count: oracle parity on every structure, then size-independent properties at
    larger scale -- permutation equivariance of per-atom counts, and batch == one-by-one."""
    from oracle import load
    from rustsasa_b200 import workloads as W
    fast = load(fast=True)
    data = W.proteome_batch(300)
    b = eng.batch(data.struct_off, data.seg_be, data.struct_seg_off, data.seg_polar)
    r = b.run_host(data.xyzr, want=("counts", "seg"))
    o = fast.run_batch(data.xyzr, data.struct_off, seg_be=data.seg_be, struct_seg_off=data.struct_seg_off)
    assert np.array_equal(r.counts, o["counts"])
    assert np.array_equal(r.seg_sasa, o["seg"])
    assert r.stats["streamed_atoms"] == 0
    # device-resident entry point gives the same answer
    import torch
    d_xyzr = torch.from_numpy(data.xyzr).cuda()
    d_counts = torch.zeros(data.n_atoms, dtype=torch.int32, device="cuda")
    d_seg = torch.zeros(len(data.seg_be), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    b.run_device(d_xyzr, counts=d_counts, seg_sasa=d_seg)
    b.sync()
    assert np.array_equal(d_counts.cpu().numpy().view(np.uint32), o["counts"])
    assert np.array_equal(d_seg.cpu().numpy(), o["seg"])
    b.close()
    # permutation equivariance (exact: the per-pair arithmetic does not depend on atom order)
    rng = np.random.default_rng(0)
    s0, s1 = int(data.struct_off[3]), int(data.struct_off[4])
    x = data.xyzr[s0:s1]
    perm = rng.permutation(x.shape[0])
    c1 = eng.calculate_sasa_internal(x, want_counts=True)[1]
    c2 = eng.calculate_sasa_internal(x[perm], want_counts=True)[1]
    assert np.array_equal(c1[perm], c2) and np.array_equal(c1, o["counts"][s0:s1])


def test_options_api_process_many(eng, tmp_path):
    """SASAOptions mirror: levels, filters and error behaviour on a hand-written PDB."""
    from rustsasa_b200 import (ProteinLevel, ResidueLevel, ChainLevel, AtomLevel, SASACalcError, SASAOptions,
                               read_structure)
    pdb = tmp_path / "mini.pdb"
    lines = [
        "ATOM      1  N   ALA A   1      11.104   6.134  -6.504  1.00  0.00           N",
        "ATOM      2  CA  ALA A   1      11.639   6.071  -5.147  1.00  0.00           C",
        "ATOM      3  C   ALA A   1      13.090   5.613  -5.183  1.00  0.00           C",
        "ATOM      4  O   ALA A   1      13.704   5.557  -6.250  1.00  0.00           O",
        "ATOM      5  CB  ALA A   1      10.813   5.126  -4.289  1.00  0.00           C",
        "ATOM      6  H   ALA A   1      10.500   6.900  -6.700  1.00  0.00           H",
        "ATOM      7  N   SER B   2      13.645   5.285  -4.023  1.00  0.00           N",
        "ATOM      8  CA  SER B   2      15.032   4.833  -3.923  1.00  0.00           C",
        "ATOM      9  C   SER B   2      15.261   4.076  -2.622  1.00  0.00           C",
        "ATOM     10  O   SER B   2      14.338   3.873  -1.833  1.00  0.00           O",
        "ATOM     11  CB  SER B   2      15.994   6.014  -4.011  1.00  0.00           C",
        "ATOM     12  OG  SER B   2      15.800   6.900  -2.920  1.00  0.00           O",
        "HETATM   13  O   HOH B 101      20.000   6.900  -2.920  1.00  0.00           O",
        "END",
    ]
    pdb.write_text("\n".join(lines) + "\n")
    st = read_structure(str(pdb))
    atoms = SASAOptions(AtomLevel).process(st, eng)
    assert atoms.shape == (11,)                       # hydrogen and HETATM dropped
    res = SASAOptions(ResidueLevel).process(st, eng)
    assert [(r.name, r.chain_id, r.is_polar) for r in res] == [("ALA", "A", False), ("SER", "B", True), ("HOH", "B", False)]
    assert res[2].value == 0.0                        # excluded HETATM residue reports 0.0
    chains = SASAOptions(ChainLevel).process(st, eng)
    assert [c.name for c in chains] == ["A", "B"]
    prot = SASAOptions(ProteinLevel).process(st, eng)
    assert abs(prot.global_total - float(atoms.sum(dtype=np.float32))) < 1e-3
    assert abs(prot.polar_total - res[1].value) < 1e-4 and abs(chains[0].value - res[0].value) < 1e-4
    assert abs(prot.polar_total + prot.non_polar_total - prot.global_total) < 1e-2
    het = SASAOptions(AtomLevel).with_include_hetatms(True).process(st, eng)
    assert het.shape == (12,)
    # RadiusMissing is returned per structure by process_many (directory mode logs and continues) ...
    out = SASAOptions(AtomLevel).with_include_hydrogens(True).process_many([st, st], eng)
    assert all(isinstance(o, SASACalcError) and o.kind == "RadiusMissing" for o in out)
    ok = SASAOptions(AtomLevel).with_include_hydrogens(True).with_allow_vdw_fallback(True).process(st, eng)
    assert ok.shape == (12,)


def test_atom_range_split_sums_to_single_gpu(eng, golden):
    """cfg5 mechanics on one GPU: the per-rank partial vectors are disjoint and their sum is the full result."""
    import torch
    from rustsasa_b200 import workloads as W
    a = W.large_assembly(30000)
    b = eng.batch(a.struct_off, a.seg_be, a.struct_seg_off, a.seg_polar)
    full = b.run_host(a.xyzr)
    parts = [b.run_atom_range_host(a.xyzr, r, 3) for r in range(3)]
    touched = sum((p.atom_sasa != 0) | (p.counts != 0) for p in parts)
    assert touched.max() == 1
    sizes = [int(((p.atom_sasa != 0) | (p.counts != 0)).sum()) for p in parts]
    assert max(sizes) - min(sizes) < 0.2 * a.n_atoms / 3          # slices of the sorted order are near-equal
    counts = parts[0].counts + parts[1].counts + parts[2].counts
    atom = parts[0].atom_sasa + parts[1].atom_sasa + parts[2].atom_sasa
    assert np.array_equal(counts, full.counts) and np.array_equal(atom, full.atom_sasa)
    d_atom = torch.from_numpy(atom).cuda()
    d_seg = torch.zeros(len(a.seg_be), dtype=torch.float32, device="cuda")
    d_prot = torch.zeros(3, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    b.reduce_device(d_atom, d_seg, d_prot)
    b.sync()
    assert np.array_equal(d_seg.cpu().numpy(), full.seg_sasa)
    assert close_enough(d_prot.cpu().numpy(), full.protein[0])
    b.close()
    # several (small) structures in one range-mode batch, with id classes, 960 points
    structs = [golden.structure(n) for n in ("example.cif", "2drt")]
    xyzr, off, seg, soff, pol = _concat(structs)
    b = eng.batch(off, seg, soff, pol)
    full = b.run_host(xyzr, n_points=960, want=("counts", "atom"))
    parts = [b.run_atom_range_host(xyzr, r, 2, n_points=960) for r in range(2)]
    assert np.array_equal(parts[0].counts + parts[1].counts, full.counts)
    assert np.array_equal(parts[0].atom_sasa + parts[1].atom_sasa, full.atom_sasa)
    from rustsasa_b200 import SasaB200Error
    with pytest.raises(SasaB200Error):
        b.run_atom_range_host(xyzr, 2, 2)
    b.close()


def test_sharded_run_matches_single(eng):
    """The multi-GPU sharding path (rustsasa_b200.shard) with the real engine: shards run one after another on
    this GPU and reassembled equal the unsharded batch."""
    from rustsasa_b200 import workloads as W
    from rustsasa_b200.shard import partition_structures, run_sharded, take_shard
    d = W.proteome_batch(40)

    def compute(sh):
        bb = eng.batch(sh.struct_off, sh.seg_be, sh.struct_seg_off, sh.seg_polar)
        try:
            return bb.run_host(sh.xyzr, sh.id_class)
        finally:
            bb.close()
    whole = run_sharded(compute, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar, rank=0, world=1)
    bounds = partition_structures(d.struct_off, 3)
    pieces = [compute(take_shard(bounds, r, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)) for r in range(3)]
    assert np.array_equal(np.concatenate([p.counts for p in pieces]), whole.counts)
    assert np.array_equal(np.concatenate([p.seg_sasa for p in pieces]), whole.seg_sasa)
    assert np.array_equal(np.concatenate([p.protein for p in pieces]), whole.protein)


@pytest.mark.parametrize("seed", range(6))
def test_cap_table_randomised_geometry(eng, oracle, seed):
    """Stress of the cap-table occlusion (sasa_cap.cuh) away from protein statistics: random point counts <= 128 and lane
    rules, probe radii 0..3 A, radii 0.4..3.2 A (cap levels from barely touching to a neighbour swallowing the atom),
    densities from sparse gas to 3x protein density, coincident and nearly coincident centres (the degenerate bin), and
    structures on both the fused and the large-structure path.  Counts must equal the oracle's EXACTLY: a table bin that
    decided a point the reference's arithmetic decides differently would show up here."""
    rng = np.random.default_rng(9000 + seed)
    n_points = int(rng.choice([1, 5, 31, 32, 33, 64, 96, 100, 127, 128]))
    lanes = int(rng.choice([4, 8, 16]))
    probe = float(rng.choice([0.0, 0.3, 1.4, 2.2, 3.0]))
    structs = []
    for t in range(10):
        n = int(rng.integers(1, 900))
        density = float(rng.choice([0.005, 0.03, 0.057, 0.12, 0.17]))
        side = (n / density) ** (1.0 / 3.0)
        xyz = rng.uniform(0.0, side, size=(n, 3)) + rng.uniform(-300, 300, size=3)
        rad = rng.uniform(0.4, 3.2, size=n) if t % 2 else rng.choice([1.42, 1.61, 1.76, 1.88], size=n)
        if n > 20:   # coincident / nearly coincident centres and an atom inside a much larger one
            xyz[1] = xyz[0]
            xyz[3] = xyz[2] + 1e-4
            xyz[5] = xyz[4] + [0.0, 0.0, 3e-4]
            rad[7], xyz[7] = 3.2, xyz[6] + 0.2
        structs.append(np.concatenate([xyz, rad[:, None]], axis=1).astype(np.float32))
    big = rng.uniform(0.0, 95.0, size=(40000, 3))          # one structure for the large path (cap table on global atoms)
    structs.append(np.concatenate([big, rng.uniform(1.2, 2.0, size=(40000, 1))], axis=1).astype(np.float32))
    off = np.cumsum([0] + [s.shape[0] for s in structs]).astype(np.uint64)
    xyzr = np.concatenate(structs)
    b = eng.batch(off)
    r = b.run_host(xyzr, probe_radius=probe, n_points=n_points, simd_lanes=lanes, want=("counts", "atom"))
    for i, s in enumerate(structs):
        o = oracle.calculate_sasa_internal(s, probe, n_points, lanes=lanes, threads=0 if s.shape[0] > 10000 else 1)
        a0, a1 = int(off[i]), int(off[i + 1])
        assert np.array_equal(r.counts[a0:a1], o["counts"]), (seed, i, n_points, lanes, probe)
        assert np.array_equal(r.atom_sasa[a0:a1], o["sasa"]), (seed, i)
    b.close()


@pytest.mark.parametrize("seed", range(5))
def test_chunked_cap_table_randomised_geometry(eng, seed):
    """The chunked cap table of the large-structure path (capm_atom, 128 < n_points <= 1024): random point counts around the
    chunk boundaries, lane rules, probe radii, radii from 0.4 to 3.2 A, coincident centres, a dense cluster (more than 128
    neighbours: cold path) and a cell block with more than 288 candidates (staging overflow), one structure with id
    classes (generic kernel).  Counts and areas must equal the oracle's exactly."""
    from oracle import load
    fast = load(fast=True)
    rng = np.random.default_rng(7700 + seed)
    n_points = int([129, 200, 256, 257, 500, 513, 960, 1000, 1023, 1024][int(rng.integers(0, 10))])
    lanes = int(rng.choice([4, 8, 16]))
    probe = float(rng.choice([0.0, 0.8, 1.4, 2.2]))
    structs = []
    for t in range(3):
        n = int(rng.integers(6000, 14000))
        density = float(rng.choice([0.02, 0.057, 0.09]))
        side = (n / density) ** (1.0 / 3.0)
        xyz = rng.uniform(0.0, side, size=(n, 3)) + rng.uniform(-200, 200, size=3)
        rad = rng.uniform(0.4, 3.2, size=n) if t == 1 else rng.choice([1.42, 1.61, 1.76, 1.88], size=n)
        xyz[1] = xyz[0]
        xyz[3] = xyz[2] + 1e-4
        rad[7], xyz[7] = 3.2, xyz[6] + 0.2
        if t == 2:   # 300 atoms inside a 5 A ball: > 128 neighbours each and > 288 candidates in the cell block
            xyz[100:400] = xyz[50] + rng.normal(scale=1.6, size=(300, 3))
        structs.append(np.concatenate([xyz, rad[:, None]], axis=1).astype(np.float32))
    for t in range(5):   # structures of the fused kernels: the same table on shared-memory atoms (sasa_small_kernel)
        n = int(rng.integers(40, 3000))
        density = float(rng.choice([0.01, 0.057, 0.1]))
        side = (n / density) ** (1.0 / 3.0)
        xyz = rng.uniform(0.0, side, size=(n, 3)) + rng.uniform(-200, 200, size=3)
        rad = rng.uniform(0.4, 3.2, size=n) if t % 2 else rng.choice([1.42, 1.61, 1.76, 1.88], size=n)
        xyz[1] = xyz[0]
        xyz[3] = xyz[2] + 1e-4
        structs.append(np.concatenate([xyz, rad[:, None]], axis=1).astype(np.float32))
    off = np.cumsum([0] + [s.shape[0] for s in structs]).astype(np.uint64)
    xyzr = np.concatenate(structs)
    b = eng.batch(off)
    r = b.run_host(xyzr, probe_radius=probe, n_points=n_points, simd_lanes=lanes, want=("counts", "atom"))
    want = [fast.calculate_sasa_internal(s, probe, n_points, lanes=lanes, threads=-1) for s in structs]
    for i, o in enumerate(want):
        a0, a1 = int(off[i]), int(off[i + 1])
        assert np.array_equal(r.counts[a0:a1], o["counts"]), (seed, i, n_points, lanes, probe)
        assert np.array_equal(r.atom_sasa[a0:a1], o["sasa"]), (seed, i)
    assert r.stats["streamed_atoms"] > 0
    # the same atoms with id classes (every atom its own class except two pairs): generic kernel, same answers
    cls = np.arange(xyzr.shape[0], dtype=np.uint32)
    r2 = b.run_host(xyzr, id_class=cls, probe_radius=probe, n_points=n_points, simd_lanes=lanes, want=("counts",))
    assert np.array_equal(r2.counts, r.counts)
    b.close()


def test_large_path_non_finite_and_empty(eng):
    """A non-finite coordinate in a structure of the large path: status 4, outputs blanked, context usable; an atom-range
    batch with an empty structure."""
    from rustsasa_b200 import SasaB200Error
    from rustsasa_b200 import workloads as W
    a = W.large_assembly(8000)
    bad = a.xyzr.copy()
    bad[4321, 2] = np.inf
    with pytest.raises(SasaB200Error) as ei:
        eng.calculate_sasa_internal(bad)
    assert ei.value.code == 4
    good = eng.calculate_sasa_internal(a.xyzr)
    assert np.isfinite(good).all() and good.shape[0] == a.n_atoms
    n = a.n_atoms
    b = eng.batch([0, 0, n, n])
    parts = [b.run_atom_range_host(a.xyzr, r, 2) for r in range(2)]
    assert np.array_equal(parts[0].atom_sasa + parts[1].atom_sasa, good)
    b.close()


def test_submit_wait_jobs_overlap_and_match(eng, oracle, golden):
    """sasa_b200_batch_submit_host / sasa_b200_job_wait: several jobs of different batches in flight at once, waited out of
    order, equal the synchronous results; a second submit on a busy batch is refused."""
    from rustsasa_b200 import SasaB200Error
    from rustsasa_b200 import workloads as W
    datas = [W.proteome_batch(24, seed=500 + i) for i in range(3)]
    batches = [eng.batch(d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar) for d in datas]
    pinned = []
    for d in datas:
        h = eng.pinned_empty((d.n_atoms, 4), np.float32)
        h[...] = d.xyzr
        pinned.append(h)
    jobs = [b.submit_host(h) for b, h in zip(batches, pinned)]
    with pytest.raises(SasaB200Error):
        batches[0].submit_host(pinned[0])
    for i in (2, 0, 1):
        r = jobs[i].wait()
        s = batches[i].run_host(datas[i].xyzr)
        assert np.array_equal(np.asarray(r.counts), s.counts) and np.array_equal(np.asarray(r.seg_sasa), s.seg_sasa)
        assert np.array_equal(np.asarray(r.protein), s.protein)
        assert r.stats["gpu_launches"] >= 1 and r.stats["n_atoms"] == datas[i].n_atoms
    # a non-finite value is reported by wait, and the batch is usable afterwards
    bad = pinned[1].copy()
    bad[5, 0] = np.nan
    j = batches[1].submit_host(bad)
    with pytest.raises(SasaB200Error) as ei:
        j.wait()
    assert ei.value.code == 4
    again = batches[1].submit_host(pinned[1]).wait()
    assert np.array_equal(np.asarray(again.counts), batches[1].run_host(datas[1].xyzr).counts)
    for b in batches:
        b.close()


def test_concurrent_single_structure_callers(eng, oracle, golden):
    """The reference's directory mode calls the engine from every worker thread (src/main.rs:375, :439): 16 threads calling
    sasa_b200_calculate_sasa_internal on one context get correct results and overlap (each call has its own stream and
    workspace), instead of queueing on a context-wide lock."""
    import threading
    import time
    names = ["example.cif", "151L_H3.pdb", "2drt", "4xfj"]
    xs = [golden.structure(n)["xyzr"] for n in names]
    want = [oracle.calculate_sasa_internal(x, PROBE, 100) for x in xs]
    for x in xs:
        eng.calculate_sasa_internal(x)                      # warm the point set and one slot
    reps = 60

    def loop(tid, out):
        ok = True
        for r in range(reps):
            k = (tid + r) % len(xs)
            sasa, counts = eng.calculate_sasa_internal(xs[k], None, PROBE, 100, -1, want_counts=True)
            ok = ok and np.array_equal(counts, want[k]["counts"]) and np.array_equal(sasa, want[k]["sasa"])
        out[tid] = ok

    def run(nthreads):
        out = [False] * nthreads
        th = [threading.Thread(target=loop, args=(t, out)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        assert all(out)
        return nthreads * reps / dt

    run(16)                                                  # creates the slots
    single, many = run(1), run(16)
    print(f"calls/s: 1 thread {single:.0f}, 16 threads {many:.0f} ({many / single:.1f}x)")
    assert many > 1.5 * single                # python's GIL caps these threads; the C++ callers below are the measurement


def test_concurrent_callers_from_cpp_threads(tmp_path):
    """tools/abi_latency.cpp: 16 std::threads on one context through the C ABI itself.  The calls overlap (own stream, own
    workspace, no context-wide lock): the concurrent call rate must be several times the single caller's."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_latency")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tools", "abi_latency.cpp"), "-L", os.path.join(root, "rustsasa_b200"), "-lsasa_b200",
                    "-Wl,-rpath," + os.path.join(root, "rustsasa_b200"), "-o", exe], check=True)
    out = subprocess.run([exe, "2622", "16", "300"], check=True, stdout=subprocess.PIPE, text=True).stdout
    r = json.loads(out.strip().splitlines()[-1])
    print(r)
    assert r["errors"] == 0
    assert r["speedup"] >= 5.0, r


def test_single_structure_levels_through_run_batch(eng, oracle, golden):
    """sasa_b200_run_batch with S = 1 (what SASAOptions::process does for one file) takes the one-structure path with its
    level sums: residue / protein outputs equal the oracle's, for a small and a large structure, with and without ids."""
    import ctypes as C
    from rustsasa_b200 import _lib
    L = eng._L
    for name in ("example.cif", "1jz8"):
        s = golden.structure(name)
        x = np.ascontiguousarray(s["xyzr"], np.float32)
        n, g = x.shape[0], len(s["seg_be"])
        off = np.array([0, n], np.uint64)
        soff = np.array([0, g], np.uint64)
        seg_be = np.ascontiguousarray(s["seg_be"], np.uint32)
        pol = np.ascontiguousarray(s["polar"], np.uint8)
        counts = np.zeros(n, np.uint32); atom = np.zeros(n, np.float32); seg = np.zeros(g, np.float32); prot = np.zeros(3, np.float32)
        outs = _lib.Outputs(counts.ctypes.data, atom.ctypes.data, seg.ctypes.data, prot.ctypes.data)
        prm = _lib.Params(PROBE, 100, 8, -1, 0)
        st = _lib.Stats()
        for cls in (None, np.arange(n, dtype=np.uint32)):
            rc = L.sasa_b200_run_batch(eng._h, x.ctypes.data, None if cls is None else cls.ctypes.data, off.ctypes.data, 1,
                                       seg_be.ctypes.data, soff.ctypes.data, pol.ctypes.data, C.byref(prm), C.byref(outs), C.byref(st))
            assert rc == 0, L.sasa_b200_last_error(eng._h)
            o = oracle.calculate_sasa_internal(x, PROBE, 100)
            assert np.array_equal(counts, o["counts"]) and np.array_equal(atom, o["sasa"])
            assert np.array_equal(seg, oracle.segment_sums(o["sasa"], s["seg_be"]))
            want = oracle.protein_totals(o["sasa"], s["seg_be"], s["polar"])
            assert np.array_equal(prot, want) if n <= 16384 else close_enough(prot, want)
            assert st.gpu_launches <= 8


def test_full_size_configs_match_oracle_fingerprints(eng):
    """BASELINE cfg4 (150k atoms x 100 points) and cfg5 (1M atoms x 960 points) at FULL size on one GPU: per-atom counts
    equal the oracle's, through their sha256 + sum stored in tests/golden/cfg_hashes.json (tools/make_cfg_hashes.py runs the
    CPU oracle; 12 s at 1M x 960).  Both the whole-structure call and the atom-range split (3 ranks, summed) are checked."""
    import hashlib
    import json
    import os
    from rustsasa_b200 import workloads as W
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = json.load(open(os.path.join(root, "tests", "golden", "cfg_hashes.json")))
    for key, data in (("cfg4", W.large_assembly(150000)), ("cfg5", W.capsid_shell(1000000))):
        g = gold[key]
        assert data.n_atoms == g["atoms"]
        assert hashlib.sha256(np.ascontiguousarray(data.xyzr).tobytes()).hexdigest() == g["xyzr_sha256"], "workload generator changed"
        b = eng.batch(data.struct_off)
        r = b.run_host(data.xyzr, n_points=g["n_points"], want=("counts", "atom"))
        assert int(r.counts.astype(np.int64).sum()) == g["sum_counts"]
        assert hashlib.sha256(np.ascontiguousarray(r.counts, dtype="<u4").tobytes()).hexdigest() == g["sha256_counts"]
        assert r.stats["streamed_atoms"] == 0
        parts = [b.run_atom_range_host(data.xyzr, k, 3, n_points=g["n_points"]) for k in range(3)]
        assert np.array_equal(parts[0].counts + parts[1].counts + parts[2].counts, r.counts)
        assert np.array_equal(parts[0].atom_sasa + parts[1].atom_sasa + parts[2].atom_sasa, r.atom_sasa)
        b.close()


def test_indexed_radius_wire_format_equals_float4(eng, golden):
    """sasa_b200_batch_run_indexed_host (12 B coordinates + 1 B palette index per atom) gives the float4 form's results bit for
    bit: a proteome sample through the fused kernels (all levels, with id classes too) and a batch with a large structure."""
    from rustsasa_b200 import SasaB200Error
    from rustsasa_b200 import workloads as W
    from rustsasa_b200.engine import index_radii
    d = W.proteome_batch(40, seed=77)
    pal, idx = index_radii(d.xyzr[:, 3])
    assert pal.shape[0] <= 16
    b = eng.batch(d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
    ref = b.run_host(d.xyzr)
    got = b.run_indexed_host(d.xyzr[:, :3], idx, pal)
    for k in ("counts", "atom_sasa", "seg_sasa", "protein"):
        assert np.array_equal(getattr(ref, k), getattr(got, k)), k
    cls = np.arange(d.n_atoms, dtype=np.uint32)
    cls[1] = cls[0]
    assert np.array_equal(b.run_indexed_host(d.xyzr[:, :3], idx, pal, id_class=cls).counts, b.run_host(d.xyzr, id_class=cls).counts)
    with pytest.raises(SasaB200Error):
        b.run_indexed_host(d.xyzr[:, :3], idx, np.zeros(300, np.float32))
    b.close()
    big = W.large_assembly(9000)
    s = golden.structure("example.cif")
    xyzr = np.concatenate([s["xyzr"], big.xyzr])
    off = np.array([0, s["xyzr"].shape[0], xyzr.shape[0]], np.uint64)
    pal, idx = index_radii(xyzr[:, 3])
    b = eng.batch(off)
    ref = b.run_host(xyzr, want=("counts", "atom"))
    got = b.run_indexed_host(xyzr[:, :3], idx, pal, want=("counts", "atom"))
    assert np.array_equal(ref.counts, got.counts) and np.array_equal(ref.atom_sasa, got.atom_sasa)
    b.close()
