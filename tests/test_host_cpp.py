"""C++ host layer (include/sasa_b200.hpp) on CPU: extraction parity with the Python mirror of row A0, error kinds,
radius tables and the JSON / XML writers.  The engine itself is not called here (no GPU)."""
import glob
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")
REF_DATA = "/root/reference/tests/data"


@pytest.fixture(scope="module")
def host():
    from rustsasa_b200 import build, host_lib
    build.build_host()
    host_lib.load()
    return host_lib


def py_pack(path, level, **kw):
    from rustsasa_b200.structure import build_atoms_and_mapping, read_structure
    return build_atoms_and_mapping(read_structure(path), level, None, kw.get("allow_vdw_fallback", False),
                                   kw.get("include_hydrogens", False), kw.get("include_hetatms", False),
                                   kw.get("read_radii_from_occupancy", False))


def same_partition(a, b):
    """ids only matter through equality: the two id vectors must induce the same partition of the atoms."""
    _, ia = np.unique(a, return_inverse=True)
    _, ib = np.unique(b, return_inverse=True)
    fa = {}
    return all(fa.setdefault(x, y) == y for x, y in zip(ia.tolist(), ib.tolist())) and len(set(ia)) == len(set(ib))


def check_same(host, path, level, **kw):
    from rustsasa_b200.structure import SASACalcError
    try:
        want = py_pack(path, level, **kw)
    except SASACalcError as e:
        with pytest.raises(host.HostError) as ei:
            host.pack(path, level, **kw)
        assert ei.value.kind == e.kind, path
        return None
    got = host.pack(path, level, **kw)
    assert np.array_equal(got["xyzr"], want.xyzr), (path, level)
    assert np.array_equal(got["seg_be"], want.seg_be.reshape(-1, 2)), (path, level)
    assert np.array_equal(got["seg_polar"], want.seg_polar), (path, level)
    assert same_partition(got["ids"], want.ids), (path, level)
    return got


@pytest.mark.parametrize("name", ["mini_altloc.pdb", "mini_models.pdb", "mini.cif"])
@pytest.mark.parametrize("level", ["atom", "residue", "chain", "protein"])
def test_extraction_matches_python_mirror_on_committed_files(host, name, level):
    path = os.path.join(DATA, name)
    check_same(host, path, level)
    check_same(host, path, level, include_hetatms=True, allow_vdw_fallback=True)
    check_same(host, path, level, include_hydrogens=True, allow_vdw_fallback=True)
    check_same(host, path, level, read_radii_from_occupancy=True, include_hetatms=True)


def test_reference_rules_on_hand_written_files(host):
    """first conformer only + blank-altloc atoms appended; hydrogens / HETATM dropped; insertion codes are separate
    residues; multi-model files overlay models (all models are walked, ids repeat across models)."""
    alt = host.pack(os.path.join(DATA, "mini_altloc.pdb"), "residue")
    # ALA 5 heavy atoms; SER: conformer A (CB, OG) then the 4 blank atoms; TYR 2A 5 atoms; GLY 4; HOH and ZN dropped
    assert alt["xyzr"].shape[0] == 5 + 6 + 5 + 4
    assert alt["seg_be"].tolist() == [[0, 5], [5, 11], [11, 16], [16, 20], [20, 20], [20, 20]]
    assert alt["seg_polar"].tolist() == [0, 1, 1, 0, 0, 0]
    assert np.allclose(alt["xyzr"][5, :3], [15.994, 6.014, -4.011])          # alt-loc A atoms come first
    assert alt["xyzr"][0, 3] == np.float32(1.64) and alt["xyzr"][1, 3] == np.float32(1.88)   # ProtOr ALA N, CA
    ch = host.pack(os.path.join(DATA, "mini_altloc.pdb"), "chain")
    assert ch["seg_be"].tolist() == [[0, 16], [16, 20]]
    with pytest.raises(host.HostError) as ei:
        host.pack(os.path.join(DATA, "mini_altloc.pdb"), "residue", include_hetatms=True)
    assert ei.value.kind == "RadiusMissing"                                   # ZN has no ProtOr entry (strict mode)
    mm = host.pack(os.path.join(DATA, "mini_models.pdb"), "atom")
    assert mm["xyzr"].shape[0] == 8 and len(set(mm["ids"].tolist())) == 4     # two models, ids repeat
    res = host.pack(os.path.join(DATA, "mini_models.pdb"), "residue")
    assert res["seg_be"].tolist() == [[4, 8], [4, 8]]                        # HashMap::insert: last writer wins
    cif = host.pack(os.path.join(DATA, "mini.cif"), "residue")
    # chain = auth_asym_id: the water (auth chain A) is the third residue of chain A, before nucleotide chain R
    assert cif["xyzr"].shape[0] == 17 and cif["seg_be"].tolist() == [[0, 8], [8, 15], [15, 15], [15, 17]]
    with pytest.raises(host.HostError) as ei:
        host.pack(os.path.join(DATA, "does_not_exist.pdb"), "residue")
    assert ei.value.kind == "IO"


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference test data not mounted")
def test_extraction_matches_python_mirror_on_reference_data(host):
    """All three tests/data/pdbs files and a spread of the 88 quality-set files, every level."""
    files = sorted(glob.glob(os.path.join(REF_DATA, "pdbs", "*")))
    files += sorted(glob.glob(os.path.join(REF_DATA, "freesasa_pdbs", "*.pdb")))[::8]
    assert len(files) >= 10
    n_atoms = 0
    for path in files:
        for level in ("atom", "residue", "chain", "protein"):
            got = check_same(host, path, level)
            if got is not None and level == "atom":
                n_atoms += got["xyzr"].shape[0]
        check_same(host, path, "residue", include_hetatms=True, allow_vdw_fallback=True)
    assert n_atoms > 30000


def test_cpp_extraction_reproduces_committed_fixture(host, golden):
    """tests/golden/structures.npz was produced by the Python mirror from the reference's files; where those files
    are mounted the C++ reader must reproduce the stored arrays exactly."""
    path = os.path.join(REF_DATA, "pdbs", "example.cif")
    if not os.path.exists(path):
        pytest.skip("reference test data not mounted")
    s = golden.structure("example.cif")
    got = host.pack(path, "residue")
    assert np.array_equal(got["xyzr"], s["xyzr"]) and np.array_equal(got["seg_be"], s["seg_be"])
    assert np.array_equal(got["seg_polar"], s["polar"])


def test_radii_and_chain_keys(host):
    from rustsasa_b200.structure import get_protor_radius, serialize_chain_id
    L = host.load()
    for res, atom in [("ASN", "CA"), ("ASN", "N"), ("CYS", "SG"), ("TYR", "OH"), ("A", "O5'"), ("HOH", "O")]:
        assert L.sasa_b200_host_get_radius(res.encode(), atom.encode()) == np.float32(get_protor_radius(res, atom))
    # tests/units.rs:211-277
    assert L.sasa_b200_host_get_radius(b"ASN", b"CA") == np.float32(1.88)
    assert L.sasa_b200_host_get_radius(b"CYS", b"SG") == np.float32(1.77)
    assert L.sasa_b200_host_get_radius(b"XXX", b"CA") == -1.0
    for s in ("A", "Z", "AA", "K", "a", "A1", "", "1", "AB"):
        assert L.sasa_b200_host_serialize_chain_id(s.encode()) == serialize_chain_id(s)


def test_writers(host):
    js = host.format_values([0.0, 25.0, 1.5, 12.34375, 1e-7, 20131.227, 123456792.0])
    assert js == '{"Atom":[0.0,25.0,1.5,12.34375,1e-7,20131.227,123456790.0]}'
    assert json.loads(js)["Atom"][5] == pytest.approx(20131.227)
    # every f32 survives the round trip through the shortest representation
    rng = np.random.default_rng(0)
    v = (rng.random(2000) * 400).astype(np.float32)
    back = np.array(json.loads(host.format_values(v))["Atom"], np.float32)
    assert np.array_equal(back, v)
    p = host.format_values([3.5, 1.25, 2.25], kind="protein")
    assert p == '{"Protein":{"global_total":3.5,"polar_total":1.25,"non_polar_total":2.25}}'
    x = host.format_values([3.5, 1.25, 2.25], xml=True, kind="protein")
    assert x == ("<Protein><global_total>3.5</global_total><polar_total>1.25</polar_total>"
                 "<non_polar_total>2.25</non_polar_total></Protein>")
    assert host.format_values([1.0, 2.5], xml=True) == "<Atom>1.0</Atom><Atom>2.5</Atom>"


def test_cli_argument_errors(host, tmp_path):
    """Failure modes of tests/integration.rs that need no device: missing input, directory without --format,
    bad radii file."""
    import subprocess
    cli = host.CLI_PATH
    r = subprocess.run([cli, str(tmp_path / "nope.pdb"), str(tmp_path / "o.json")], capture_output=True, text=True)
    assert r.returncode != 0 and "does not exist" in r.stderr
    (tmp_path / "in").mkdir()
    r = subprocess.run([cli, str(tmp_path / "in"), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode != 0 and "--format" in r.stderr
    r = subprocess.run([cli, os.path.join(DATA, "mini.cif"), str(tmp_path / "o.json"), "-r", str(tmp_path / "no.config")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "radii" in r.stderr.lower()
    r = subprocess.run([cli, "only_one_arg"], capture_output=True, text=True)
    assert r.returncode == 2
