// capi.cpp -- small extern "C" surface over the C++ host layer so that the Python test-suite can exercise it
// (extraction parity against rustsasa_b200/structure.py on CPU, end-to-end JSON on the GPU box).
#include <cstring>

#include "../../../include/sasa_b200.hpp"

using namespace rust_sasa;

namespace {
void put_error(char *err, size_t n, const std::string &msg) {
    if (err && n) {
        std::strncpy(err, msg.c_str(), n - 1);
        err[n - 1] = '\0';
    }
}
OptionValues make_opts(float probe, size_t n_points, int include_h, int include_het, int vdw_fallback, int occ_radii,
                       const char *radii_file) {
    OptionValues o;
    o.probe_radius = probe;
    o.n_points = n_points;
    o.include_hydrogens = include_h != 0;
    o.include_hetatms = include_het != 0;
    o.allow_vdw_fallback = vdw_fallback != 0;
    o.read_radii_from_occupancy = occ_radii != 0;
    if (radii_file && *radii_file) o.radii_config = std::make_shared<const RadiiConfig>(load_radii_from_file(radii_file));
    return o;
}
}  // namespace

#define HOST_API extern "C" __attribute__((visibility("default")))

HOST_API void *sasa_b200_host_pack(const char *path, int level, int include_h, int include_het, int vdw_fallback, int occ_radii,
                                   const char *radii_file, char *err, size_t errlen) {
    try {
        const OptionValues o = make_opts(1.4f, 100, include_h, include_het, vdw_fallback, occ_radii, radii_file);
        return new Packed(build_atoms_and_mapping(pdb::open(path), (LevelKind)level, o));
    } catch (const SASACalcError &e) {
        put_error(err, errlen, std::string(e.kind_name()) + ": " + e.what());
    } catch (const std::exception &e) {
        put_error(err, errlen, std::string("IO: ") + e.what());
    }
    return nullptr;
}
HOST_API size_t sasa_b200_host_pack_atoms(const void *p) { return static_cast<const Packed *>(p)->n_atoms(); }
HOST_API size_t sasa_b200_host_pack_segments(const void *p) { return static_cast<const Packed *>(p)->seg_polar.size(); }
HOST_API void sasa_b200_host_pack_copy(const void *h, float *xyzr, uint64_t *ids, uint32_t *seg_be, uint8_t *polar) {
    const Packed &p = *static_cast<const Packed *>(h);
    if (xyzr) std::memcpy(xyzr, p.xyzr.data(), p.xyzr.size() * 4);
    if (ids) std::memcpy(ids, p.ids.data(), p.ids.size() * 8);
    if (seg_be) std::memcpy(seg_be, p.seg_be.data(), p.seg_be.size() * 4);
    if (polar) std::memcpy(polar, p.seg_polar.data(), p.seg_polar.size());
}
HOST_API void sasa_b200_host_pack_free(void *p) { delete static_cast<Packed *>(p); }

// path -> SASAOptions<level>::process -> JSON text (needs the GPU).  Returns the JSON length, or -1 with err filled.
HOST_API long sasa_b200_host_process_json(const char *path, int level, float probe, size_t n_points, int include_h, int include_het,
                                          int vdw_fallback, int occ_radii, const char *radii_file, char *out, size_t outlen, char *err,
                                          size_t errlen) {
    try {
        const OptionValues o = make_opts(probe, n_points, include_h, include_het, vdw_fallback, occ_radii, radii_file);
        const pdb::PDB st = pdb::open(path);
        auto res = process_many({&st}, (LevelKind)level, o);
        if (auto *e = std::get_if<SASACalcError>(&res[0])) throw *e;
        const std::string js = sasa_result_to_json(std::get<SASAResult>(res[0]));
        if (js.size() + 1 > outlen) {
            put_error(err, errlen, "IO: output buffer too small");
            return -1;
        }
        std::memcpy(out, js.c_str(), js.size() + 1);
        return (long)js.size();
    } catch (const SASACalcError &e) {
        put_error(err, errlen, std::string(e.kind_name()) + ": " + e.what());
    } catch (const std::exception &e) {
        put_error(err, errlen, std::string("IO: ") + e.what());
    }
    return -1;
}

// writers on hand-made values (CPU-only formatting tests): kind 0 atom, 1 residue, 2 chain, 3 protein
HOST_API long sasa_b200_host_format(int xml, int kind, const float *values, size_t n, char *out, size_t outlen) {
    SASAResult r;
    if (kind == 3 && n >= 3) {
        r = ProteinResult{values[0], values[1], values[2]};
    } else if (kind == 1) {   // residues with made-up metadata: serial i + 1, every third one polar SER with insertion code "B"
        std::vector<ResidueResult> v;
        for (size_t i = 0; i < n; ++i)
            v.push_back(ResidueResult{(std::ptrdiff_t)i + 1, i % 3 == 2 ? "B" : "", values[i], i % 3 == 2 ? "SER" : "MET", i % 3 == 2, "A"});
        r = std::move(v);
    } else if (kind == 2) {
        std::vector<ChainResult> v;
        for (size_t i = 0; i < n; ++i) v.push_back(ChainResult{std::string(1, (char)('A' + i % 26)), values[i]});
        r = std::move(v);
    } else {
        r = std::vector<float>(values, values + n);
    }
    const std::string s = xml ? sasa_result_to_xml(r) : sasa_result_to_json(r);
    if (s.size() + 1 > outlen) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (long)s.size();
}
// B-factor write-back on hand-made values (CPU-only tests of src/utils/io.rs:20-64 and of the coordinate writers):
// kind 0 atom (values[i] per atom), 1 residue (values[i] per residue of the file, in order), 2 chain, 3 protein
// (values[0..3]); `bad_serial` != 0 corrupts the first residue's serial number to exercise the reference's assert.
// Output text: format 0 = PDB, 1 = mmCIF.  Returns the text length, or -1 with err filled.
HOST_API long sasa_b200_host_writeback(const char *path, int kind, const float *values, size_t n, int bad_serial, int format,
                                       char *out, size_t outlen, char *err, size_t errlen) {
    try {
        pdb::PDB st = pdb::open(path);
        SASAResult r;
        if (kind == 0) {
            r = std::vector<float>(values, values + n);
        } else if (kind == 1) {
            std::vector<ResidueResult> v;
            for (const auto &m : st.models)
                for (const auto &ch : m.chains)
                    for (const auto &res : ch.residues) {
                        if (v.size() >= n) break;
                        const std::string name = res.name().value_or("");
                        v.push_back(ResidueResult{res.serial + (bad_serial && v.empty() ? 1 : 0), res.icode, values[v.size()], name,
                                                  is_polar_residue(name), ch.id});
                    }
            r = std::move(v);
        } else if (kind == 2) {
            std::vector<ChainResult> v;
            for (const auto &m : st.models)
                for (const auto &ch : m.chains) {
                    if (v.size() >= n) break;
                    v.push_back(ChainResult{ch.id, values[v.size()]});
                }
            r = std::move(v);
        } else {
            r = ProteinResult{values[0], values[1], values[2]};
        }
        sasa_result_to_protein_object(st, r);
        const std::string text = format == 1 ? pdb::to_mmcif_string(st, "?") : pdb::to_pdb_string(st);
        if (text.size() + 1 > outlen) {
            put_error(err, errlen, "IO: output buffer too small");
            return -1;
        }
        std::memcpy(out, text.c_str(), text.size() + 1);
        return (long)text.size();
    } catch (const std::exception &e) {
        put_error(err, errlen, std::string("ProteinSerialization: ") + e.what());
    }
    return -1;
}
HOST_API long sasa_b200_host_serialize_chain_id(const char *s) { return (long)serialize_chain_id(s); }
HOST_API float sasa_b200_host_get_radius(const char *res, const char *atom) {
    auto r = get_radius(res, atom, nullptr);
    return r ? *r : -1.0f;
}

// ---- the reader's hierarchy, flattened for the reader tests (pdbtbx/tests/*.rs restated in tests/test_reader_pins.py) ----
// Every atom of the file in hierarchy order (models -> chains -> residues -> conformers -> atoms, like PDB::atoms()).
namespace {
struct FlatAtoms {
    std::vector<double> xyzob;          // x, y, z, occupancy, b per atom
    std::vector<long long> ints;        // serial, residue serial, model index, model serial, hetero, is_hydrogen, conformer index, residue index per atom
    std::string text;                   // per atom "chain|icode|altloc|resname|name|element\n"
    size_t models = 0, chains = 0, residues = 0, conformers = 0;
};
}  // namespace

HOST_API void *sasa_b200_host_flatten(const char *path, char *err, size_t errlen) {
    try {
        const pdb::PDB st = pdb::open(path);
        auto *f = new FlatAtoms();
        f->models = st.models.size();
        long long mi = 0, ri = 0;
        for (const auto &m : st.models) {
            f->chains += m.chains.size();
            for (const auto &c : m.chains)
                for (const auto &r : c.residues) {
                    ++f->residues;
                    long long ci = 0;
                    for (const auto &cf : r.conformers) {
                        ++f->conformers;
                        for (const auto &a : cf.atoms) {
                            f->xyzob.insert(f->xyzob.end(), {a.x, a.y, a.z, a.occupancy, a.b_factor});
                            f->ints.insert(f->ints.end(), {(long long)a.serial, (long long)r.serial, mi, (long long)m.serial, (long long)a.hetero,
                                                           (long long)(a.element == "H"), ci, ri});
                            f->text += c.id + "|" + r.icode + "|" + cf.altloc + "|" + cf.name + "|" + a.name + "|" + a.element + "\n";
                        }
                        ++ci;
                    }
                    ++ri;
                }
            ++mi;
        }
        return f;
    } catch (const std::exception &e) {
        put_error(err, errlen, std::string("IO: ") + e.what());
    }
    return nullptr;
}
HOST_API void sasa_b200_host_flat_sizes(const void *h, size_t *out /* atoms, models, chains, residues, conformers, text bytes */) {
    const FlatAtoms &f = *static_cast<const FlatAtoms *>(h);
    out[0] = f.ints.size() / 8; out[1] = f.models; out[2] = f.chains; out[3] = f.residues; out[4] = f.conformers; out[5] = f.text.size();
}
HOST_API void sasa_b200_host_flat_copy(const void *h, double *xyzob, long long *ints, char *text) {
    const FlatAtoms &f = *static_cast<const FlatAtoms *>(h);
    if (xyzob) std::memcpy(xyzob, f.xyzob.data(), f.xyzob.size() * sizeof(double));
    if (ints) std::memcpy(ints, f.ints.data(), f.ints.size() * sizeof(long long));
    if (text) std::memcpy(text, f.text.data(), f.text.size());
}
HOST_API void sasa_b200_host_flat_free(void *h) { delete static_cast<FlatAtoms *>(h); }
