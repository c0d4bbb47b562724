// structure.cpp -- host-side reader and radius tables in front of the hot path.
//
// A columnar ATOM/HETATM reader for PDB and mmCIF that reproduces the ordering rules the reference inherits from
// its pdbtbx fork, because they define the atom order of every per-atom output:
//   * hierarchy and first-seen ordering      pdbtbx/src/read/pdb/parser.rs:144-251 (chains and residues in
//     insertion order per model, conformers keyed by (residue name, alt-loc)); serial / residue-number wrap
//     handling :167-173
//   * mmCIF columns                           pdbtbx/src/read/mmcif/parser.rs:456-600 (auth_asym_id / auth_seq_id
//     with label_* fallback, serial number = running atom count per model, hetero from group_PDB, model from
//     pdbx_PDB_model_num)
//   * blank alt-loc atoms appended to every other conformer     pdbtbx/src/validate.rs:302-325
//   * element resolution                      pdbtbx/src/structs/atom.rs:74-87
// Coordinates are parsed as double and cast to float at extraction, like the reference (f64 as f32).
// It is not a general structure parser (no symmetry, bonds or validation).
#include <algorithm>
#include <cctype>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_set>

#include "../../../include/sasa_b200.hpp"

namespace rust_sasa {

// ---- radii ----------------------------------------------------------------------------------------------------
namespace {
const char kProtorConfig[] =
#include "protor_config.inc"
    ;

std::vector<std::string_view> split_ws(std::string_view s) {
    std::vector<std::string_view> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < s.size() && !std::isspace((unsigned char)s[j])) ++j;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

// isspace of the "C" locale, inline (twelve calls per atom record)
inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

__attribute__((always_inline)) inline std::string_view trim(std::string_view s) {
    while (!s.empty() && is_space(s.front())) s.remove_prefix(1);
    while (!s.empty() && is_space(s.back())) s.remove_suffix(1);
    return s;
}

std::string upper(std::string_view s) {
    std::string r(s);
    for (char &c : r) c = (char)std::toupper((unsigned char)c);
    return r;
}
}  // namespace

// src/utils/consts.rs:31-81: "types:" section maps a type name to a radius, "atoms:" maps (residue, atom) to a type.
RadiiConfig parse_radii_config(std::string_view content) {
    std::unordered_map<std::string, float> types;
    RadiiConfig atoms;
    enum { None, Types, Atoms } section = None;
    size_t pos = 0;
    while (pos <= content.size()) {
        size_t eol = content.find('\n', pos);
        if (eol == std::string_view::npos) eol = content.size();
        std::string_view line = trim(content.substr(pos, eol - pos));
        pos = eol + 1;
        if (line.empty() || line[0] == '#' || line.substr(0, 5) == "name:") continue;
        if (line == "types:") { section = Types; continue; }
        if (line == "atoms:") { section = Atoms; continue; }
        auto parts = split_ws(line);
        if (section == Types && parts.size() >= 2) {
            char *end = nullptr;
            std::string v(parts[1]);
            const float r = std::strtof(v.c_str(), &end);
            if (end && *end == '\0') types[std::string(parts[0])] = r;
        } else if (section == Atoms && parts.size() >= 3) {
            auto it = types.find(std::string(parts[2]));
            if (it != types.end()) atoms[std::string(parts[0])][std::string(parts[1])] = it->second;
        }
    }
    return atoms;
}

RadiiConfig load_radii_from_file(const std::string &path) {
    std::ifstream fh(path);
    if (!fh) throw SASACalcError(SASACalcError::Kind::RadiiFileLoad, "Failed to load radii file: cannot open " + path);
    std::stringstream ss;
    ss << fh.rdbuf();
    return parse_radii_config(ss.str());
}

const RadiiConfig &protor_radii() {
    static const RadiiConfig table = parse_radii_config(kProtorConfig);
    return table;
}

std::optional<float> get_protor_radius(const std::string &residue, const std::string &atom) {
    const auto &t = protor_radii();
    auto r = t.find(residue);
    if (r == t.end()) return std::nullopt;
    auto a = r->second.find(atom);
    if (a == r->second.end()) return std::nullopt;
    return a->second;
}

std::optional<float> get_radius(const std::string &residue, const std::string &atom, const RadiiConfig *custom) {
    if (custom) {
        auto r = custom->find(residue);
        if (r != custom->end()) {
            auto a = r->second.find(atom);
            if (a != r->second.end()) return a->second;
        }
    }
    return get_protor_radius(residue, atom);
}

std::ptrdiff_t serialize_chain_id(std::string_view s) {
    std::ptrdiff_t result = 0;
    for (char c : s)
        if ((unsigned char)c < 128 && std::isalpha((unsigned char)c)) result = result * 10 + (std::toupper((unsigned char)c) - 64);
    return result;
}

bool is_polar_residue(const std::string &name) {
    static const std::unordered_set<std::string> polar = {"SER", "THR", "CYS", "ASN", "GLN", "TYR"};
    return polar.count(name) != 0;
}

const char *SASACalcError::kind_name() const {
    switch (kind_) {
        case Kind::ElementMissing: return "ElementMissing";
        case Kind::VanDerWaalsMissing: return "VanDerWaalsMissing";
        case Kind::RadiusMissing: return "RadiusMissing";
        case Kind::AtomMapToLevelElementFailed: return "AtomMapToLevelElementFailed";
        case Kind::FailedToGetResidueName: return "FailedToGetResidueName";
        case Kind::RadiiFileLoad: return "RadiiFileLoad";
        case Kind::Device: return "Device";
    }
    return "?";
}

// ---- hierarchy ----------------------------------------------------------------------------------------------------
namespace pdb {

std::optional<std::string> Residue::name() const {
    if (conformers.empty()) return std::nullopt;
    for (const Conformer &c : conformers)
        if (c.name != conformers[0].name) return std::nullopt;
    return conformers[0].name;
}

std::size_t PDB::atom_count() const {
    std::size_t n = 0;
    for (const Model &m : models)
        for (const Chain &c : m.chains)
            for (const Residue &r : c.residues)
                for (const Conformer &f : r.conformers) n += f.atoms.size();
    return n;
}

namespace {

const std::unordered_set<std::string> &elements() {
    static const std::unordered_set<std::string> set = [] {
        const char *all =
            "H HE LI BE B C N O F NE NA MG AL SI P S CL AR K CA SC TI V CR MN FE CO NI CU ZN GA GE AS SE BR KR RB SR Y ZR "
            "NB MO TC RU RH PD AG CD IN SN SB TE I XE CS BA LA CE PR ND PM SM EU GD TB DY HO ER TM YB LU HF TA W RE OS IR "
            "PT AU HG TL PB BI PO AT RN FR RA AC TH PA U NP PU AM CM BK CF ES FM MD NO LR RF DB SG BH HS MT DS RG CN NH FL "
            "MC LV TS OG";
        std::unordered_set<std::string> s;
        for (auto t : split_ws(all)) s.emplace(t);
        return s;
    }();
    return set;
}

// pdbtbx Atom::new: the element column if it names an element, else the atom name, else its first letter if CHNOS.
std::string resolve_element(std::string_view element, std::string_view atom_name) {
    element = trim(element);
    if (element.size() == 1) {   // the common case, no allocation: a one-letter symbol of the organic set
        const char c = (char)std::toupper((unsigned char)element[0]);
        if (std::strchr("CNOSHPFIBKVWYU", c)) return std::string(1, c);
    }
    const std::string e = upper(element);
    if (elements().count(e)) return e;
    const std::string n = upper(trim(atom_name));
    if (elements().count(n)) return n;
    if (!n.empty() && std::strchr("CHNOS", n[0])) return std::string(1, n[0]);
    return std::string();
}

// Where the previous atom went: consecutive atoms of a file almost always share chain, residue and conformer, so the
// hash lookups (and the key strings they need) are only paid at the boundaries.
struct AddCursor {
    const Model *model = nullptr;
    size_t chain = 0, residue = 0, conformer = 0;
    std::ptrdiff_t res_serial = 0;
    bool valid = false;
};

void add_atom(Model &model, std::string_view chain_id, std::ptrdiff_t res_serial, std::string_view icode,
              std::string_view res_name, std::string_view altloc, AtomRec &&atom, AddCursor *cur = nullptr) {
    if (cur && cur->valid && cur->model == &model && cur->res_serial == res_serial) {
        Chain &chain = model.chains[cur->chain];
        Residue &res = chain.residues[cur->residue];
        Conformer &conf = res.conformers[cur->conformer];
        if (chain.id == chain_id && res.icode == icode && conf.name == res_name && conf.altloc == altloc) {
            conf.atoms.push_back(std::move(atom));
            ++model.atom_count;
            return;
        }
    }
    const std::string chain_key(chain_id);
    auto ci = model.chain_index.find(chain_key);
    if (ci == model.chain_index.end()) {
        ci = model.chain_index.emplace(chain_key, model.chains.size()).first;
        model.chains.emplace_back();
        model.chains.back().id = chain_key;
        model.chains.back().residue_index.reserve(1024);   // growth by rehashing was 8 rehashes per 500-residue chain
        model.chains.back().residues.reserve(256);
    }
    Chain &chain = model.chains[ci->second];
    std::string rkey = std::to_string(res_serial);
    rkey += '|';
    rkey += icode;
    auto ri = chain.residue_index.find(rkey);
    if (ri == chain.residue_index.end()) {
        ri = chain.residue_index.emplace(rkey, chain.residues.size()).first;
        chain.residues.emplace_back();
        chain.residues.back().serial = res_serial;
        chain.residues.back().icode = std::string(icode);
    }
    Residue &res = chain.residues[ri->second];
    size_t fi = res.conformers.size();
    for (size_t i = 0; i < res.conformers.size(); ++i)
        if (res.conformers[i].name == res_name && res.conformers[i].altloc == altloc) { fi = i; break; }
    if (fi == res.conformers.size()) {
        res.conformers.emplace_back();
        res.conformers.back().name = std::string(res_name);
        res.conformers.back().altloc = std::string(altloc);
        res.conformers.back().atoms.reserve(16);   // one allocation per residue instead of four or five doublings
    }
    res.conformers[fi].atoms.push_back(std::move(atom));
    ++model.atom_count;
    if (cur) {
        cur->model = &model;
        cur->chain = ci->second;
        cur->residue = ri->second;
        cur->conformer = fi;
        cur->res_serial = res_serial;
        cur->valid = true;
    }
}

// pdbtbx/src/validate.rs:302-325: the blank-altloc conformer is removed and its atoms appended to every other one.
void reshuffle_conformers(PDB &pdb) {
    for (Model &m : pdb.models)
        for (Chain &c : m.chains)
            for (Residue &r : c.residues) {
                if (r.conformers.size() <= 1) continue;
                int blank = -1;
                for (size_t i = 0; i < r.conformers.size(); ++i)
                    if (r.conformers[i].altloc.empty()) blank = (int)i;
                if (blank < 0) continue;
                Conformer shared = std::move(r.conformers[blank]);
                r.conformers.erase(r.conformers.begin() + blank);
                const double count = (double)(r.conformers.size() + 1);
                for (Conformer &f : r.conformers)
                    for (const AtomRec &a : shared.atoms) {
                        AtomRec b = a;
                        b.occupancy = a.occupancy / count;
                        f.atoms.push_back(std::move(b));
                    }
            }
}

// strtol / strtod semantics on a trimmed field without the temporary string: an optional sign, then from_chars (which
// rounds correctly, like Rust's str::parse the reference uses); anything from_chars cannot take falls back to strtod.
long parse_long(std::string_view s, bool *ok = nullptr) {
    s = trim(s);
    std::string_view t = s;
    if (!t.empty() && t.front() == '+') t.remove_prefix(1);
    long v = 0;
    const auto r = std::from_chars(t.data(), t.data() + t.size(), v, 10);
    if (ok) *ok = !s.empty() && r.ec == std::errc() && r.ptr == t.data() + t.size();
    return r.ec == std::errc() ? v : 0;
}

__attribute__((noinline)) double parse_double_slow(std::string_view s, std::string_view t) {
    double v = 0.0;
    const auto r = std::from_chars(t.data(), t.data() + t.size(), v, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == t.data() + t.size()) return v;
    const std::string tmp(s);
    return std::strtod(tmp.c_str(), nullptr);
}

double parse_double(std::string_view s) {
    s = trim(s);
    std::string_view t = s;
    if (!t.empty() && t.front() == '+') t.remove_prefix(1);
    {   // fast path for plain decimals (what coordinate, occupancy and B-factor columns hold): an integer mantissa below 2^53
        // over an exact power of ten is ONE correctly rounded division -- the same double a correctly rounded decimal parser
        // (from_chars in parse_double_slow, Rust's str::parse::<f64> in the reference) returns
        static const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
        const char *q = t.data(), *const end = q + t.size();
        const bool neg = q < end && *q == '-';
        if (neg) ++q;
        unsigned long long m = 0;
        int digits = 0, frac = -1;
        for (; q < end; ++q) {
            const unsigned d = (unsigned)(*q - '0');
            if (d <= 9) {
                m = m * 10 + d;
                ++digits;
                if (frac >= 0) ++frac;
            } else if (*q == '.' && frac < 0) {
                frac = 0;
            } else {
                break;
            }
        }
        if (q == end && digits > 0 && digits <= 15) {
            const double v = (double)m / kPow10[frac < 0 ? 0 : frac];
            return neg ? -v : v;
        }
    }
    return parse_double_slow(s, t);
}

void flush_model(PDB &pdb, Model &model) {
    if (!model.chains.empty()) pdb.models.push_back(std::move(model));
}

}  // namespace

namespace {
std::string read_file(const std::string &path) {
    FILE *fh = std::fopen(path.c_str(), "rb");
    if (!fh) throw std::runtime_error("cannot open " + path);
    std::string buf;
    char chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof chunk, fh)) > 0) buf.append(chunk, n);
    std::fclose(fh);
    return buf;
}

// Upper-cased, trimmed copy of a short fixed-width field into `out` (no heap: these fit the small-string buffer).
// (the "C" locale's toupper -- the process never sets another -- written out: std::toupper is a call per character)
__attribute__((always_inline)) inline void upper_trim(std::string_view f, std::string &out) {
    f = trim(f);
    out.assign(f);
    for (char &c : out)
        if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
}
}  // namespace

PDB read_pdb(const std::string &path) {
    const std::string text = read_file(path);
    PDB pdb;
    Model model;
    AddCursor cur;
    long serial_add = 0, res_add = 0, last_serial = -1, last_res = -1;
    unsigned long long atom_id = 0;   // pdbtbx's id iterator (pdbtbx/src/read/pdb/parser.rs:116)
    // blank chain IDs take the letter of an 'A'..'Z' cycle that advances on every TER record and is never reset, not even
    // by MODEL (pdbtbx/src/read/pdb/parser.rs:113-115, :176-180, :511): TER-separated chains without IDs (typical MD
    // output) become chains A, B, C, ...
    int chain_letter = 0;
    std::ptrdiff_t model_no = 0;      // number of the last MODEL record (0 without one), parser.rs:98, :305
    std::string name, resname;
    char col[81];
    size_t pos = 0;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string_view raw(text.data() + pos, eol - pos);
        pos = eol + 1;
        if (!raw.empty() && raw.back() == '\r') raw.remove_suffix(1);
        if (raw.size() <= 6) {   // short lines: only TER counts (pdbtbx lexer.rs:87-91)
            if (raw.size() > 2 && raw.compare(0, 3, "TER") == 0) chain_letter = (chain_letter + 1) % 26;
            continue;
        }
        const bool is_atom = raw.compare(0, 6, "ATOM  ") == 0, is_het = raw.compare(0, 6, "HETATM") == 0;
        if (is_atom || is_het) {
            // the record as 80 columns, space-padded
            const size_t len = std::min<size_t>(raw.size(), 80);
            std::memcpy(col, raw.data(), len);
            std::memset(col + len, ' ', 80 - len);
            col[80] = 0;
            const std::string_view line(col, 80);
            bool ok = false;
            long serial = parse_long(line.substr(6, 5), &ok);
            if (!ok) serial = 0;
            upper_trim(line.substr(12, 4), name);
            const char alt = line[16];
            upper_trim(line.substr(17, 3), resname);
            const char chain_c = line[21];
            const long resseq = parse_long(line.substr(22, 4));
            const char icode = line[26];
            const std::string_view occ_s = trim(line.substr(54, 6));
            if (serial == 0 && last_serial == 99999) serial_add += 100000;
            if (resseq == 0 && last_res == 9999) res_add += 10000;
            AtomRec a;
            a.hetero = is_het;
            a.serial = (std::size_t)(serial + serial_add);
            a.name = name;
            a.x = parse_double(line.substr(30, 8));
            a.y = parse_double(line.substr(38, 8));
            a.z = parse_double(line.substr(46, 8));
            a.occupancy = occ_s.empty() ? 1.0 : parse_double(occ_s);
            {
                const std::string_view b_s = trim(line.substr(60, 6));
                a.b_factor = b_s.empty() ? 0.0 : parse_double(b_s);
                // columns 79-80: charge as digit + sign ("1+", "2-")
                const char cd = line[78], cs = line[79];
                if (cd >= '0' && cd <= '9' && (cs == '+' || cs == '-')) a.charge = (cs == '-' ? -1 : 1) * (cd - '0');
            }
            a.element = resolve_element(line.substr(76, 2), name);
            {
                char idb[24];
                const auto r = std::to_chars(idb, idb + sizeof idb, atom_id++);
                a.id.assign(idb, r.ptr);
            }
            const char chain_s[1] = {is_space(chain_c) ? (char)('A' + chain_letter) : chain_c};
            add_atom(model, std::string_view(chain_s, 1), resseq + res_add, icode == ' ' ? std::string_view() : std::string_view(&line[26], 1),
                     resname, alt == ' ' ? std::string_view() : std::string_view(&line[16], 1), std::move(a), &cur);
            last_serial = serial;
            last_res = resseq;
        } else if (raw.compare(0, 6, "MODEL ") == 0) {
            // the model in progress is closed by the NEXT MODEL record (or MASTER, or the end of the file), not by ENDMDL,
            // which pdbtbx ignores (parser.rs:282-306, lexer.rs:82): atoms between ENDMDL and MODEL stay in the old model
            model.serial = model_no;
            flush_model(pdb, model);
            cur.valid = false;
            bool ok = false;
            const long no = parse_long(raw.substr(6), &ok);
            model_no = ok ? no : 0;
            model = Model();
        } else if (raw.compare(0, 6, "MASTER") == 0) {
            model.serial = model_no;
            flush_model(pdb, model);   // parser.rs:435-456
            cur.valid = false;
            model = Model();
        } else if (raw.compare(0, 6, "TER   ") == 0) {
            chain_letter = (chain_letter + 1) % 26;   // lexer.rs:83, :89; parser.rs:511
        }
    }
    model.serial = model_no;
    flush_model(pdb, model);
    reshuffle_conformers(pdb);
    return pdb;
}

namespace {

// Next token of a CIF text at or after p (STAR tokenisation as far as _atom_site needs it): whitespace separates tokens,
// '#' at a token start comments out the rest of the line, a quote only closes before whitespace or the end of the text,
// and a ';' in the first column opens a text field that runs to the next line starting with ';'.  `bare` tells an
// unquoted token (which may be a tag or a keyword) from a quoted value.  Returns false at the end of the text.
bool cif_next(const char *&p, const char *const begin, const char *const end, std::string_view &tok, bool &bare) {
    for (;;) {
        while (p < end && is_space(*p)) ++p;
        if (p >= end) return false;
        if (*p == '#') {
            while (p < end && *p != '\n') ++p;
            continue;
        }
        break;
    }
    if (*p == ';' && (p == begin || p[-1] == '\n')) {
        const char *q = p + 1;
        const char *stop = q;
        for (;;) {   // the closing ';' is the first character of a line
            while (stop < end && *stop != '\n') ++stop;
            if (stop >= end || (stop + 1 < end && stop[1] == ';')) break;
            ++stop;
        }
        tok = trim(std::string_view(q, (size_t)(stop - q)));
        p = stop + 2 <= end ? stop + 2 : end;
        bare = false;
        return true;
    }
    if (*p == '\'' || *p == '"') {
        const char qc = *p;
        const char *q = p + 1;
        while (q < end && !(*q == qc && (q + 1 == end || is_space(q[1]))) && *q != '\n') ++q;
        tok = std::string_view(p + 1, (size_t)(q - p - 1));
        p = q < end && *q == qc ? q + 1 : q;
        bare = false;
        return true;
    }
    const char *q = p;
    while (q < end && !is_space(*q)) ++q;
    tok = std::string_view(p, (size_t)(q - p));
    p = q;
    bare = true;
    return true;
}

bool cif_ends_loop(std::string_view t) {
    if (t.empty()) return false;
    if (t[0] == '_') return true;
    auto starts = [&](const char *kw) {
        const size_t n = std::strlen(kw);
        if (t.size() < n) return false;
        for (size_t i = 0; i < n; ++i)
            if (std::tolower((unsigned char)t[i]) != kw[i]) return false;
        return true;
    };
    return starts("loop_") || starts("data_") || starts("save_") || starts("global_") || starts("stop_");
}

}  // namespace

PDB read_mmcif(const std::string &path) {
    const std::string text = read_file(path);
    PDB pdb;
    std::unordered_map<long, size_t> model_index;
    const char *const begin = text.data(), *const end = text.data() + text.size();
    std::vector<std::string_view> tok;
    std::string name, resname;
    AddCursor cur;
    const char *p = begin;
    std::string_view t;
    bool bare = false;
    bool have = cif_next(p, begin, end, t, bare);
    while (have) {
        if (!(bare && t.size() == 5 && cif_ends_loop(t) && t[4] == '_')) {   // not "loop_"
            have = cif_next(p, begin, end, t, bare);
            continue;
        }
        // a loop: its tags, then -- for _atom_site -- its values as a token stream (rows may wrap across lines and hold
        // quoted or ';' text values; pdbtbx parses loops token-wise too, pdbtbx/src/read/mmcif/lexer.rs)
        std::unordered_map<std::string, int> col;
        int ncol = 0;
        bool site = false;
        have = cif_next(p, begin, end, t, bare);
        while (have && bare && !t.empty() && t[0] == '_') {
            if (t.rfind("_atom_site.", 0) == 0) {
                site = true;
                col[std::string(t.substr(11))] = ncol;
            }
            ++ncol;
            have = cif_next(p, begin, end, t, bare);
        }
        if (!site || ncol == 0) continue;   // some other loop: its values are skipped by the scan above
        auto idx = [&](const char *key) { auto it = col.find(key); return it == col.end() ? -1 : it->second; };
        const int c_model = idx("pdbx_PDB_model_num"), c_group = idx("group_PDB"), c_atom = idx("label_atom_id"),
                  c_comp = idx("label_comp_id"), c_aseq = idx("auth_seq_id"), c_lseq = idx("label_seq_id"),
                  c_achain = idx("auth_asym_id"), c_lchain = idx("label_asym_id"), c_x = idx("Cartn_x"), c_y = idx("Cartn_y"),
                  c_z = idx("Cartn_z"), c_occ = idx("occupancy"), c_b = idx("B_iso_or_equiv"), c_id = idx("id"),
                  c_charge = idx("pdbx_formal_charge"), c_sym = idx("type_symbol"), c_icode = idx("pdbx_PDB_ins_code"),
                  c_alt = idx("label_alt_id");
        // value of a column; false when the column is absent or holds "." / "?"
        auto get = [&](int c, std::string_view &v) {
            if (c < 0) return false;
            v = tok[(size_t)c];
            return !(v == "." || v == "?");
        };
        tok.clear();
        while (have && !(bare && cif_ends_loop(t))) {
            tok.push_back(t);
            have = cif_next(p, begin, end, t, bare);
            if ((int)tok.size() < ncol) continue;
            {
                std::string_view v;
                const long model_no = get(c_model, v) ? parse_long(v) : 1;
                auto mi = model_index.find(model_no);
                if (mi == model_index.end()) {
                    mi = model_index.emplace(model_no, pdb.models.size()).first;
                    pdb.models.emplace_back();
                    pdb.models.back().serial = model_no;
                    cur.valid = false;   // the models vector may have moved
                }
                Model &model = pdb.models[mi->second];
                const bool hetero = get(c_group, v) && v == "HETATM";
                if (get(c_atom, v)) upper_trim(v, name); else name.clear();
                if (get(c_comp, v)) upper_trim(v, resname); else resname.clear();
                std::string_view seq, chain, icode, alt;
                const bool has_seq = get(c_aseq, seq) || get(c_lseq, seq);
                const bool has_chain = get(c_achain, chain) || get(c_lchain, chain);
                AtomRec a;
                a.hetero = hetero;
                a.serial = model.atom_count;
                a.name = name;
                a.x = get(c_x, v) ? parse_double(v) : 0.0;
                a.y = get(c_y, v) ? parse_double(v) : 0.0;
                a.z = get(c_z, v) ? parse_double(v) : 0.0;
                a.occupancy = get(c_occ, v) ? parse_double(v) : 1.0;
                a.b_factor = get(c_b, v) ? parse_double(v) : 0.0;
                if (get(c_id, v)) a.id.assign(v);
                else a.id = std::to_string(model.atom_count);
                if (get(c_charge, v) && !v.empty()) a.charge = (int)parse_long(v);
                a.element = resolve_element(get(c_sym, v) ? v : std::string_view(), name);
                if (!get(c_icode, icode)) icode = std::string_view();
                if (!get(c_alt, alt)) alt = std::string_view();
                add_atom(model, has_chain ? chain : std::string_view(), has_seq ? parse_long(seq) : 0, icode, resname, alt, std::move(a),
                         &cur);
            }
            tok.clear();
        }
        if (!tok.empty())
            throw std::runtime_error("mmCIF _atom_site loop ends inside a row (" + std::to_string(tok.size()) + " of " +
                                     std::to_string(ncol) + " values) in " + path);
    }
    reshuffle_conformers(pdb);
    return pdb;
}

PDB open(const std::string &path) {
    const size_t dot = path.find_last_of('.');
    std::string ext = dot == std::string::npos ? "" : path.substr(dot);
    for (char &c : ext) c = (char)std::tolower((unsigned char)c);
    if (ext == ".cif" || ext == ".mmcif") return read_mmcif(path);
    return read_pdb(path);
}

}  // namespace pdb
}  // namespace rust_sasa
