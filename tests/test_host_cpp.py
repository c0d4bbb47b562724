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
    # quick-xml writes primitives with `to_string()` (Rust's Display): shortest digits, no exponent, no trailing ".0"
    assert host.format_values([1.0, 2.5, 1e-7, 1.5e10, 0.0], xml=True) == \
        "<Atom>1</Atom><Atom>2.5</Atom><Atom>0.0000001</Atom><Atom>15000000000</Atom><Atom>0</Atom>"
    back = np.array([float(t) for t in host.format_values(v, xml=True).replace("</Atom>", "").split("<Atom>")[1:]], np.float32)
    assert np.array_equal(back, v)


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


# ---- B-factor write-back (src/utils/io.rs:20-64) and the coordinate writers after pdbtbx::save -----------------------
# The four inputs below are the ones of the reference's own unit tests (src/utils/io.rs:73-247).
IO_ATOM = """ATOM      1  N   ALA A   1      20.154  16.967  25.000  1.00 10.00           N
ATOM      2  CA  ALA A   1      19.030  16.155  25.000  1.00 15.00           C
ATOM      3  C   ALA A   1      17.948  16.712  25.000  1.00 20.00           C
END
"""
IO_RESIDUE = """ATOM      1  N   ALA A   1      20.154  16.967  25.000  1.00 10.00           N
ATOM      2  CA  ALA A   1      19.030  16.155  25.000  1.00 15.00           C
ATOM      3  N   GLY A   2      17.948  16.712  25.000  1.00 20.00           N
ATOM      4  CA  GLY A   2      16.500  17.000  25.000  1.00 25.00           C
END
"""
IO_CHAIN = """ATOM      1  N   ALA A   1      20.154  16.967  25.000  1.00 10.00           N
ATOM      2  CA  ALA A   1      19.030  16.155  25.000  1.00 15.00           C
ATOM      3  N   GLY B   1      17.948  16.712  25.000  1.00 20.00           N
ATOM      4  CA  GLY B   1      16.500  17.000  25.000  1.00 25.00           C
END
"""


def b_factors_of_pdb_text(text):
    return [float(l[60:66]) for l in text.splitlines() if l.startswith(("ATOM", "HETATM"))]


def test_sasa_result_to_protein_object_reference_cases(host, tmp_path):
    """test_sasa_result_to_protein_object_{atom,residue,chain,protein} of src/utils/io.rs:73-247."""
    f = tmp_path / "a.pdb"
    f.write_text(IO_ATOM)
    assert b_factors_of_pdb_text(host.writeback(str(f), "atom", [5.0, 10.0, 15.0])) == [5.0, 10.0, 15.0]
    f.write_text(IO_RESIDUE)
    assert b_factors_of_pdb_text(host.writeback(str(f), "residue", [100.0, 200.0])) == [100.0, 100.0, 200.0, 200.0]
    f.write_text(IO_CHAIN)
    assert b_factors_of_pdb_text(host.writeback(str(f), "chain", [300.0, 400.0])) == [300.0, 300.0, 400.0, 400.0]
    f.write_text(IO_ATOM)
    assert b_factors_of_pdb_text(host.writeback(str(f), "protein", [500.0, 200.0, 300.0])) == [500.0, 500.0, 500.0]


def test_writeback_failure_modes(host, tmp_path):
    """Where the reference returns Err or panics: result shorter than pdb.atoms() (what an atom-level result with filtered
    hydrogens / HETATMs runs into, src/utils/io.rs:26-29), residue serial mismatch (:37), negative / non-finite values
    (pdbtbx set_b_factor)."""
    f = tmp_path / "a.pdb"
    f.write_text(IO_ATOM)
    with pytest.raises(host.HostError, match="index out of bounds"):
        host.writeback(str(f), "atom", [5.0, 10.0])
    with pytest.raises(host.HostError, match="negative"):
        host.writeback(str(f), "atom", [5.0, -1.0, 2.0])
    with pytest.raises(host.HostError, match="not finite"):
        host.writeback(str(f), "protein", [float("nan"), 0.0, 0.0])
    f.write_text(IO_RESIDUE)
    with pytest.raises(host.HostError, match="serial_number"):
        host.writeback(str(f), "residue", [1.0, 2.0], bad_serial=True)


def test_pdb_writer_follows_pdbtbx_field_rules(host, tmp_path):
    """pdbtbx/src/save/pdb.rs:112-127, :520-581: every sized field is the last `width` characters of its text, leading
    zeros trimmed, LEFT-aligned (so serial and residue numbers are left-aligned), TER after every chain, END last; values
    wider than their field lose their leading characters (1234.5 in a {:6.2} field prints as 234.50)."""
    f = tmp_path / "a.pdb"
    f.write_text(IO_CHAIN.replace("  1.00 25.00           C", "  0.50 25.00           C1+"))
    text = host.writeback(str(f), "chain", [300.0, 1234.5])
    assert text.splitlines() == [
        "ATOM  1     N    ALA A1         20.154  16.967  25.000  1.00300.00          N ",
        "ATOM  2     CA   ALA A1         19.030  16.155  25.000  1.00300.00          C ",
        "TER2          ALA A1   ",
        "ATOM  3     N    GLY B1         17.948  16.712  25.000  1.00234.50          N ",
        "ATOM  4     CA   GLY B1         16.500  17.000  25.000  0.50234.50          C 1+",
        "TER4          GLY B1   ",
        "END",
    ]


def test_mmcif_writer_and_round_trip(host, tmp_path):
    """_atom_site loop of pdbtbx/src/save/mmcif.rs:262-412 (column set, base-26 label chain, 1-based label_seq_id, aligned
    columns, print_float) and: what the writers emit, the reader takes back with the same atoms and B-factors."""
    f = tmp_path / "a.pdb"
    f.write_text(IO_CHAIN)
    cif = host.writeback(str(f), "atom", [1.0, 2.5, 0.125, 1.4235263], fmt="cif")
    rows = [l.split() for l in cif.splitlines() if l.startswith("ATOM")]
    assert cif.startswith("data_?\n#\n_entry.id   ?\n#\n_audit_conform.dict_name       mmcif_pdbx.dic\n")
    assert cif.rstrip().endswith("#") and cif.count("_atom_site.") == 19
    assert rows[0] == ["ATOM", "0", "N", "N", ".", "ALA", "B", "A", "1", "1", "1", ".", "20.154", "16.967", "25.0", "1.0", "1.0",
                       "0", "0"]   # model number 0: pdbtbx's default for PDB files without MODEL records (read/pdb/parser.rs:98)
    assert rows[3][1] == "3" and rows[3][6:11] == ["C", "B", "2", "1", "1"] and rows[3][16] == "1.42353"
    assert len({len(l) for l in cif.splitlines() if l.startswith("ATOM")}) == 1      # aligned table
    # round trip through both formats
    g = tmp_path / "b.cif"
    g.write_text(cif)
    back = host.writeback(str(g), "protein", [7.0, 0.0, 0.0], fmt="pdb")
    assert b_factors_of_pdb_text(back) == [7.0] * 4
    a = host.pack(str(f), "chain", allow_vdw_fallback=True)
    b = host.pack(str(g), "chain", allow_vdw_fallback=True)
    assert np.array_equal(a["xyzr"], b["xyzr"]) and np.array_equal(a["seg_be"], b["seg_be"])


def test_writeback_on_altloc_and_multimodel_files(host):
    """Atom-level write-back indexes ALL atoms of the hierarchy (every conformer of every model): the result vector must
    be that long, which is the reference's behaviour (src/utils/io.rs:26-29) and the reason its CLI's atom-level pdb/cif
    output only works on files without hydrogens, HETATMs and alternative locations."""
    for name in ["mini_altloc.pdb", "mini_models.pdb"]:
        path = os.path.join(DATA, name)
        n_all = sum(1 for _ in b_factors_of_pdb_text(host.writeback(path, "protein", [1.0, 0.0, 0.0])))
        vals = np.arange(n_all, dtype=np.float32)
        assert b_factors_of_pdb_text(host.writeback(path, "atom", vals)) == vals.tolist()
        if n_all > 1:
            with pytest.raises(host.HostError):
                host.writeback(path, "atom", vals[:-1])


def test_number_fields_parse_like_the_mirror(host, tmp_path):
    """The allocation-free field parser (from_chars + fallbacks) against the Python mirror's float(): explicit '+' signs,
    exponents, missing occupancy / B-factor columns, short lines, CRLF line ends, serial / residue-number wrap-around."""
    lines = [
        "ATOM      1  N   ALA A   1     +20.154 -16.967  2.5e+1  1.00 10.00           N",
        "ATOM      2  CA  ALA A   1      19.030  16.155  25.000",
        "ATOM  99999  C   ALA A9999      17.948  16.712  25.000  0.50               C",
        "ATOM      0  O   ALA A   0      16.500  17.000  25.000  1.00  0.00           O",
        "HETATM    1  O   HOH A   1      1.0e1   +.5     -0.     1.00  0.00           O",
    ]
    f = tmp_path / "n.pdb"
    f.write_bytes(("\r\n".join(lines) + "\r\nEND\r\n").encode())
    for level in ("atom", "residue", "protein"):
        got = check_same(host, str(f), level, include_hetatms=True, allow_vdw_fallback=True)
        assert got is not None and got["xyzr"].shape[0] == 5
    got = host.pack(str(f), "atom", include_hetatms=True, allow_vdw_fallback=True)
    assert np.allclose(got["xyzr"][0, :3], [20.154, -16.967, 25.0]) and np.allclose(got["xyzr"][4, :3], [10.0, 0.5, 0.0])


def test_decimal_fast_path_matches_correctly_rounded_parse(host, tmp_path):
    """The reader's fast path for plain decimals (integer mantissa / exact power of ten, one IEEE division) must give the double a
    correctly rounded decimal parser gives -- Python's float() here, Rust's str::parse::<f64> in the reference -- and hence the same
    f32 coordinate: 3,000 random %8.3f fields of every magnitude and sign the column holds, compared bit for bit."""
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.uniform(-999.999, 9999.999, 2000), rng.uniform(-1.0, 1.0, 700), rng.uniform(-99.0, 99.0, 300)])
    lines = []
    for i in range(0, vals.size, 3):
        x, y, z = vals[i:i + 3]
        lines.append("ATOM  %5d  CA  ALA A%4d    %8.3f%8.3f%8.3f  1.00  0.00           C" % (i // 3 + 1, i // 3 + 1, x, y, z))
    f = tmp_path / "r.pdb"
    f.write_text("\n".join(lines) + "\nEND\n")
    got = host.pack(str(f), "atom")
    want = np.array([[float("%8.3f" % v) for v in vals[i:i + 3]] for i in range(0, vals.size, 3)], np.float64).astype(np.float32)
    assert got["xyzr"].shape[0] == want.shape[0]
    assert np.array_equal(got["xyzr"][:, :3].view(np.uint32), want.view(np.uint32))
    assert check_same(host, str(f), "residue") is not None


def test_residue_and_chain_json_shape(host):
    """serde's externally tagged enum with the struct fields in declaration order (src/structures/atomic.rs:26-70,
    SURVEY.md 8f row f-3)."""
    js = host.format_values([25.0, 0.5, 101.25], kind="residue")
    assert js == ('{"Residue":[{"serial_number":1,"insertion_code":"","value":25.0,"name":"MET","is_polar":false,"chain_id":"A"},'
                  '{"serial_number":2,"insertion_code":"","value":0.5,"name":"MET","is_polar":false,"chain_id":"A"},'
                  '{"serial_number":3,"insertion_code":"B","value":101.25,"name":"SER","is_polar":true,"chain_id":"A"}]}')
    assert host.format_values([1.5, 2.0], kind="chain") == '{"Chain":[{"name":"A","value":1.5},{"name":"B","value":2.0}]}'
    rng = np.random.default_rng(1)
    v = (rng.random(700) * 300).astype(np.float32)
    back = json.loads(host.format_values(v, kind="residue"))["Residue"]
    assert [r["serial_number"] for r in back] == list(range(1, 701))
    assert np.array_equal(np.array([r["value"] for r in back], np.float32), v)


def test_cli_rejects_malformed_numbers_with_a_usage_error(host):
    """clap in the reference reports an invalid numeric value and exits with code 2; an uncaught std::invalid_argument
    (abort) is not acceptable.  No engine call is made, so this runs without a GPU."""
    import subprocess
    for bad in (["-n", "abc"], ["-n", "0"], ["-n", "-5"], ["-p", "1.x"], ["-t", "many"], ["--tile", "0"], ["--devices", "x"], ["-n"]):
        r = subprocess.run([host.CLI_PATH] + bad + ["in", "out"], capture_output=True, text=True)
        assert r.returncode == 2, (bad, r.returncode, r.stderr[:200])
        assert "error:" in r.stderr and "Usage:" in r.stderr
