// writers.cpp -- result serialisation, src/utils/io.rs:11-18.
//
// JSON is serde_json's rendering of the externally tagged `SASAResult` enum (src/structures/atomic.rs:62-70), fields in
// declaration order, f32 printed shortest-round-trip like ryu ("25.0", "0.0", "1e-7"; XML prints the same digits the way Rust's Display does: "25", "0", "0.0000001").  XML follows quick-xml's serde
// serializer: a newtype variant holding a sequence becomes one element per item named after the variant.  The XML
// shape could not be checked against the crate here (no Rust toolchain); the JSON shape is pinned by the reference's
// own test helpers (tests/common/io.rs).  B-factor write-back (src/utils/io.rs:20-64) and the coordinate-section writers after
// pdbtbx::save are at the end of this file.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/sasa_b200.hpp"

namespace rust_sasa {
namespace {

// f32 the way ryu's pretty printer (what serde_json uses) lays it out: the shortest round-trip digit string D with
// decimal exponent, printed as an integer with ".0", a plain decimal, "0.000ddd", or d.ddde[-]x outside [1e-5, 1e13).
std::string fmt_f32(float v) {
    if (std::isnan(v) || std::isinf(v)) return "null";   // serde_json writes non-finite floats as null
    if (v == 0.0f) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    const std::string sci(buf, r.ptr);
    const size_t e = sci.find('e');
    std::string digits;
    for (char c : sci.substr(0, e))
        if (c != '.') digits += c;
    const int exp10 = std::atoi(sci.c_str() + e + 1);
    const int len = (int)digits.size(), kk = exp10 + 1, k = kk - len;
    std::string out = std::signbit(v) ? "-" : "";
    if (k >= 0 && kk <= 13) out += digits + std::string((size_t)k, '0') + ".0";
    else if (kk > 0 && kk <= 13) out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    else if (kk > -6 && kk <= 0) out += "0." + std::string((size_t)(-kk), '0') + digits;
    else if (len == 1) out += digits + "e" + std::to_string(kk - 1);
    else out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    return out;
}

std::string json_str(const std::string &s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) {
                    char b[8];
                    std::snprintf(b, sizeof b, "\\u%04x", c);
                    o += b;
                } else {
                    o += (char)c;
                }
        }
    }
    return o + "\"";
}

std::string xml_text(const std::string &s) {
    std::string o;
    for (char c : s) {
        if (c == '&') o += "&amp;";
        else if (c == '<') o += "&lt;";
        else if (c == '>') o += "&gt;";
        else o += c;
    }
    return o;
}

// f32 the way quick-xml's serde serializer writes primitives: `value.to_string()`, i.e. Rust's Display for f32 -- the
// shortest round-trip digits laid out positionally, never with an exponent and without a trailing ".0" ("25", "0.1",
// "0.0000001", "15000000000"); NaN / inf / -inf by name.  (serde_json, above, prints the same digits ryu-style.)
std::string xml_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    if (v == 0.0f) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    const std::string sci(buf, r.ptr);
    const size_t e = sci.find('e');
    std::string digits;
    for (char c : sci.substr(0, e))
        if (c != '.') digits += c;
    const int kk = std::atoi(sci.c_str() + e + 1) + 1, len = (int)digits.size();   // value = 0.DIGITS x 10^kk
    std::string out = std::signbit(v) ? "-" : "";
    if (kk >= len) out += digits + std::string((size_t)(kk - len), '0');
    else if (kk > 0) out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    else out += "0." + std::string((size_t)(-kk), '0') + digits;
    return out;
}

}  // namespace

std::string sasa_result_to_json(const SASAResult &result) {
    std::string o;
    if (auto *v = std::get_if<std::vector<float>>(&result)) {
        o = "{\"Atom\":[";
        for (size_t i = 0; i < v->size(); ++i) { if (i) o += ','; o += fmt_f32((*v)[i]); }
        o += "]}";
    } else if (auto *v = std::get_if<std::vector<ResidueResult>>(&result)) {
        o.reserve(32 + v->size() * 112);   // appended in place: the chain of operator+ temporaries cost 0.75 us per residue
        o = "{\"Residue\":[";
        char num[24];
        for (size_t i = 0; i < v->size(); ++i) {
            const ResidueResult &r = (*v)[i];
            if (i) o += ',';
            o += "{\"serial_number\":";
            o.append(num, std::to_chars(num, num + sizeof num, (long long)r.serial_number).ptr);
            o += ",\"insertion_code\":";
            o += json_str(r.insertion_code);
            o += ",\"value\":";
            o += fmt_f32(r.value);
            o += ",\"name\":";
            o += json_str(r.name);
            o += r.is_polar ? ",\"is_polar\":true,\"chain_id\":" : ",\"is_polar\":false,\"chain_id\":";
            o += json_str(r.chain_id);
            o += '}';
        }
        o += "]}";
    } else if (auto *v = std::get_if<std::vector<ChainResult>>(&result)) {
        o = "{\"Chain\":[";
        for (size_t i = 0; i < v->size(); ++i) {
            if (i) o += ',';
            o += "{\"name\":" + json_str((*v)[i].name) + ",\"value\":" + fmt_f32((*v)[i].value) + "}";
        }
        o += "]}";
    } else {
        const ProteinResult &p = std::get<ProteinResult>(result);
        o = "{\"Protein\":{\"global_total\":" + fmt_f32(p.global_total) + ",\"polar_total\":" + fmt_f32(p.polar_total) +
            ",\"non_polar_total\":" + fmt_f32(p.non_polar_total) + "}}";
    }
    return o;
}

std::string sasa_result_to_xml(const SASAResult &result) {
    std::string o;
    if (auto *v = std::get_if<std::vector<float>>(&result)) {
        for (float f : *v) o += "<Atom>" + xml_f32(f) + "</Atom>";
    } else if (auto *v = std::get_if<std::vector<ResidueResult>>(&result)) {
        for (const ResidueResult &r : *v) {
            o += "<Residue><serial_number>" + std::to_string(r.serial_number) + "</serial_number>";
            o += r.insertion_code.empty() ? "<insertion_code/>" : "<insertion_code>" + xml_text(r.insertion_code) + "</insertion_code>";
            o += "<value>" + xml_f32(r.value) + "</value><name>" + xml_text(r.name) + "</name><is_polar>" +
                 (r.is_polar ? "true" : "false") + "</is_polar><chain_id>" + xml_text(r.chain_id) + "</chain_id></Residue>";
        }
    } else if (auto *v = std::get_if<std::vector<ChainResult>>(&result)) {
        for (const ChainResult &c : *v) o += "<Chain><name>" + xml_text(c.name) + "</name><value>" + xml_f32(c.value) + "</value></Chain>";
    } else {
        const ProteinResult &p = std::get<ProteinResult>(result);
        o = "<Protein><global_total>" + xml_f32(p.global_total) + "</global_total><polar_total>" + xml_f32(p.polar_total) +
            "</polar_total><non_polar_total>" + xml_f32(p.non_polar_total) + "</non_polar_total></Protein>";
    }
    return o;
}


// ---- src/utils/io.rs:20-64 -----------------------------------------------------------------------------------------
namespace {

// pdbtbx Atom::set_b_factor (pdbtbx/src/structs/atom.rs:313-330)
void set_b_factor(pdb::AtomRec &a, double v) {
    if (!std::isfinite(v))
        throw std::runtime_error("The value of the new b_factor is not finite for atom " + std::to_string(a.serial) + " value " + std::to_string(v));
    if (v < 0.0)
        throw std::runtime_error("The value of the new b_factor is negative for atom " + std::to_string(a.serial) + " value " + std::to_string(v));
    a.b_factor = v;
}

template <class F>
void for_each_atom(pdb::Residue &r, F &&f) {
    for (auto &c : r.conformers)
        for (auto &a : c.atoms) f(a);
}

}  // namespace

void sasa_result_to_protein_object(pdb::PDB &original_pdb, const SASAResult &result) {
    if (auto *v = std::get_if<std::vector<float>>(&result)) {
        // src/utils/io.rs:25-30: the i-th atom of pdb.atoms_mut() -- ALL atoms, whatever process() filtered -- gets v[i]
        size_t i = 0;
        for (auto &m : original_pdb.models)
            for (auto &ch : m.chains)
                for (auto &r : ch.residues)
                    for_each_atom(r, [&](pdb::AtomRec &a) {
                        if (i >= v->size())   // the reference panics here (index out of bounds)
                            throw std::runtime_error("index out of bounds: the result holds " + std::to_string(v->size()) +
                                                     " atoms but the structure has more (hydrogens / HETATMs / alternative conformers were filtered)");
                        set_b_factor(a, (double)(*v)[i++]);
                    });
    } else if (auto *v = std::get_if<std::vector<ResidueResult>>(&result)) {
        size_t i = 0;
        for (auto &m : original_pdb.models)
            for (auto &ch : m.chains)
                for (auto &r : ch.residues) {
                    if (i >= v->size()) throw std::runtime_error("index out of bounds: more residues in the structure than in the result");
                    const ResidueResult &item = (*v)[i];
                    if (r.serial != item.serial_number)   // assert!(residue.serial_number() == item.serial_number)
                        throw std::runtime_error("assertion failed: residue.serial_number() == item.serial_number");
                    for_each_atom(r, [&](pdb::AtomRec &a) { set_b_factor(a, (double)item.value); });
                    ++i;
                }
    } else if (auto *v = std::get_if<std::vector<ChainResult>>(&result)) {
        size_t i = 0;
        for (auto &m : original_pdb.models)
            for (auto &ch : m.chains) {
                for (auto &r : ch.residues)
                    for_each_atom(r, [&](pdb::AtomRec &a) {
                        if (i >= v->size()) throw std::runtime_error("index out of bounds: more chains in the structure than in the result");
                        if ((*v)[i].name != ch.id) throw std::runtime_error("assertion failed: v[i].name == id");
                        set_b_factor(a, (double)(*v)[i].value);
                    });
                ++i;
            }
    } else {
        const ProteinResult &pr = std::get<ProteinResult>(result);
        for (auto &m : original_pdb.models)
            for (auto &ch : m.chains)
                for (auto &r : ch.residues) for_each_atom(r, [&](pdb::AtomRec &a) { set_b_factor(a, (double)pr.global_total); });
    }
}

// ---- pdbtbx::save, coordinate section (pdbtbx/src/save/pdb.rs:110-145, :507-613; save/mmcif.rs:88-112, :262-412) -----
namespace pdb {
namespace {

std::string fixed(double v, int width, int prec) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%*.*f", width, prec, v);
    return buf;
}

// One field of pdbtbx's get_line: width 0 appends the text; otherwise the LAST `width` characters, leading '0's trimmed
// (an all-zero cell prints as "0"), LEFT-aligned and space-padded to the width.
void field(std::string &line, size_t width, const std::string &text) {
    if (width == 0) { line += text; return; }
    const std::string cell = text.substr(text.size() - std::min(width, text.size()));
    size_t z = 0;
    while (z < cell.size() && cell[z] == '0') ++z;
    std::string out = (!cell.empty() && z == cell.size()) ? std::string("0") : cell.substr(z);
    if (out.size() < width) out.append(width - out.size(), ' ');
    line += out;
}

std::string element_symbol(const std::string &upper_sym) {   // pdbtbx Element::symbol: "C", "Cl", "Fe"
    std::string s = upper_sym;
    for (size_t i = 1; i < s.size(); ++i) s[i] = (char)std::tolower((unsigned char)s[i]);
    return s;
}

std::string pdb_charge(int charge) {   // pdbtbx/src/structs/atom.rs:357-368
    if (charge == 0 || charge < -9 || charge > 9) return "";
    return std::string(1, (char)('0' + std::abs(charge))) + (charge < 0 ? '-' : '+');
}

// Print a float with at least one and at most five decimals (pdbtbx/src/save/mmcif.rs:405-412); Rust's `{}` of an f64 is
// the shortest round-trip form, which std::to_chars gives as well.
std::string print_float(double num) {
    const double rounded = std::round(num * 100000.0) / 100000.0;
    if (std::fabs(std::round(rounded) - rounded) < 2.220446049250313e-16) return std::to_string((long long)std::trunc(rounded)) + ".0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, rounded, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

std::string number_to_base26(size_t num) {   // pdbtbx/src/structs/helper.rs:38-46
    std::string out(1, (char)('A' + num % 26));
    num /= 26;
    while (num != 0) { out.push_back((char)('A' + num % 26)); num /= 26; }
    return std::string(out.rbegin(), out.rend());
}

bool chain_has_atoms(const Chain &c) {
    for (const auto &r : c.residues)
        for (const auto &cf : r.conformers)
            if (!cf.atoms.empty()) return true;
    return false;
}

void write_text(const std::string &path, const std::string &text) {
    FILE *fh = std::fopen(path.c_str(), "wb");
    if (!fh) throw std::runtime_error("Could not open file " + path + " for writing");
    const bool ok = std::fwrite(text.data(), 1, text.size(), fh) == text.size();
    if (std::fclose(fh) != 0 || !ok) throw std::runtime_error("Could not write file " + path);
}

}  // namespace

std::string to_pdb_string(const PDB &pdb) {
    std::string out;
    const bool multiple_models = pdb.models.size() > 1;
    for (const Model &model : pdb.models) {
        if (multiple_models) out += "MODEL        " + std::to_string(model.serial) + "\n";
        for (const Chain &chain : model.chains) {
            if (!chain_has_atoms(chain)) continue;
            const AtomRec *last_atom = nullptr;
            const Conformer *last_conformer = nullptr;
            const Residue *last_residue = nullptr;
            auto atom_line = [&](std::string &line, const AtomRec &a, const Conformer &cf, const Residue &r) {
                field(line, 5, std::to_string(a.serial));
                field(line, 0, " ");
                field(line, 4, a.name);
                field(line, 1, cf.altloc.empty() ? " " : cf.altloc);
                field(line, 4, cf.name);
                field(line, 1, chain.id);
                field(line, 4, std::to_string(r.serial));
                field(line, 1, r.icode.empty() ? " " : r.icode);
            };
            for (const Residue &r : chain.residues) {
                if (!r.conformers.empty()) { last_residue = &r; last_conformer = &r.conformers.back(); }
                for (const Conformer &cf : r.conformers)
                    for (const AtomRec &a : cf.atoms) {
                        std::string line;
                        field(line, 6, a.hetero ? "HETATM" : "ATOM  ");
                        atom_line(line, a, cf, r);
                        field(line, 0, "   ");
                        field(line, 8, fixed(a.x, 8, 3));
                        field(line, 8, fixed(a.y, 8, 3));
                        field(line, 8, fixed(a.z, 8, 3));
                        field(line, 6, fixed(a.occupancy, 6, 2));
                        field(line, 6, fixed(a.b_factor, 6, 2));
                        field(line, 0, "          ");
                        field(line, 2, element_symbol(a.element));
                        field(line, 0, pdb_charge(a.charge));
                        out += line + "\n";
                        last_atom = &a;
                    }
            }
            if (last_atom && last_conformer && last_residue) {
                std::string line = "TER";
                field(line, 5, std::to_string(last_atom->serial));
                field(line, 0, "      ");
                field(line, 3, last_conformer->name);
                field(line, 0, " ");
                field(line, 1, chain.id);
                field(line, 4, std::to_string(last_residue->serial));
                out += line + "\n";
            }
        }
        if (multiple_models) out += "ENDMDL\n";
    }
    out += "END\n";
    return out;
}

std::string to_mmcif_string(const PDB &pdb, const std::string &name) {
    std::string out = "data_" + name + "\n#\n_entry.id   " + name +
                      "\n#\n_audit_conform.dict_name       mmcif_pdbx.dic\n_audit_conform.dict_version    5.338\n"
                      "_audit_conform.dict_location   http://mmcif.pdb.org/dictionaries/ascii/mmcif_pdbx.dic\n";
    out += "loop_\n_atom_site.group_PDB\n_atom_site.id\n_atom_site.type_symbol\n_atom_site.label_atom_id\n_atom_site.label_alt_id\n"
           "_atom_site.label_comp_id\n_atom_site.label_asym_id\n_atom_site.auth_asym_id\n_atom_site.label_entity_id\n"
           "_atom_site.label_seq_id\n_atom_site.auth_seq_id\n_atom_site.pdbx_PDB_ins_code\n_atom_site.Cartn_x\n_atom_site.Cartn_y\n"
           "_atom_site.Cartn_z\n_atom_site.occupancy\n_atom_site.B_iso_or_equiv\n_atom_site.pdbx_formal_charge\n"
           "_atom_site.pdbx_PDB_model_num\n";
    std::vector<std::vector<std::string>> lines;
    for (const Model &model : pdb.models) {
        size_t chain_index = 0;
        for (const Chain &chain : model.chains) {
            ++chain_index;
            for (size_t ri = 0; ri < chain.residues.size(); ++ri) {
                const Residue &r = chain.residues[ri];
                for (const Conformer &cf : r.conformers)
                    for (const AtomRec &a : cf.atoms)
                        lines.push_back({a.hetero ? "HETATM" : "ATOM", a.id, element_symbol(a.element), a.name,
                                         cf.altloc.empty() ? "." : cf.altloc, cf.name, number_to_base26(chain_index), chain.id,
                                         std::to_string(chain_index), std::to_string(ri + 1), std::to_string(r.serial),
                                         r.icode.empty() ? "." : r.icode, print_float(a.x), print_float(a.y), print_float(a.z),
                                         print_float(a.occupancy), print_float(a.b_factor), std::to_string(a.charge),
                                         std::to_string(model.serial)});
            }
        }
    }
    if (!lines.empty()) {
        std::vector<size_t> sizes(lines[0].size(), 1);
        for (const auto &l : lines)
            for (size_t i = 0; i < l.size(); ++i) sizes[i] = std::max(sizes[i], l[i].size());
        for (const auto &l : lines) {
            out += l[0] + std::string(sizes[0] - l[0].size(), ' ');
            for (size_t i = 1; i < l.size(); ++i) {
                out += ' ';
                const bool blank = l[i].find_first_not_of(" \t\r\n") == std::string::npos;
                if (!blank) out += l[i] + std::string(sizes[i] - l[i].size(), ' ');
                else out += "?" + std::string(sizes[i] - 1, ' ');
            }
            out += '\n';
        }
    }
    out += "#\n";
    return out;
}

void save_pdb(const PDB &pdb, const std::string &path) { write_text(path, to_pdb_string(pdb)); }
void save_mmcif(const PDB &pdb, const std::string &path) {
    write_text(path, to_mmcif_string(pdb, "?"));   // no identifier is kept by the reader: pdbtbx writes "?" for None
}
void save(const PDB &pdb, const std::string &path) {
    const size_t dot = path.find_last_of('.');
    std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
    for (char &c : ext) c = (char)std::tolower((unsigned char)c);
    if (ext == "pdb") save_pdb(pdb, path);
    else if (ext == "cif" || ext == "mmcif") save_mmcif(pdb, path);
    else throw std::runtime_error("Incorrect extension: could not determine the type of the given file, make it .pdb or .cif");
}

}  // namespace pdb

}  // namespace rust_sasa
