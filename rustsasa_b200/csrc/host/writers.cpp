// writers.cpp -- result serialisation, src/utils/io.rs:11-18.
//
// JSON is serde_json's rendering of the externally tagged `SASAResult` enum (src/structures/atomic.rs:62-70), fields in
// declaration order, f32 printed shortest-round-trip like ryu ("25.0", "0.0", "1e-7").  XML follows quick-xml's serde
// serializer: a newtype variant holding a sequence becomes one element per item named after the variant.  The XML
// shape could not be checked against the crate here (no Rust toolchain); the JSON shape is pinned by the reference's
// own test helpers (tests/common/io.rs).  B-factor write-back into PDB / mmCIF (src/utils/io.rs:20-64) is not provided.
#include <charconv>
#include <cmath>
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

std::string xml_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    return fmt_f32(v);
}

}  // namespace

std::string sasa_result_to_json(const SASAResult &result) {
    std::string o;
    if (auto *v = std::get_if<std::vector<float>>(&result)) {
        o = "{\"Atom\":[";
        for (size_t i = 0; i < v->size(); ++i) { if (i) o += ','; o += fmt_f32((*v)[i]); }
        o += "]}";
    } else if (auto *v = std::get_if<std::vector<ResidueResult>>(&result)) {
        o = "{\"Residue\":[";
        for (size_t i = 0; i < v->size(); ++i) {
            const ResidueResult &r = (*v)[i];
            if (i) o += ',';
            o += "{\"serial_number\":" + std::to_string(r.serial_number) + ",\"insertion_code\":" + json_str(r.insertion_code) +
                 ",\"value\":" + fmt_f32(r.value) + ",\"name\":" + json_str(r.name) + ",\"is_polar\":" + (r.is_polar ? "true" : "false") +
                 ",\"chain_id\":" + json_str(r.chain_id) + "}";
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

}  // namespace rust_sasa
