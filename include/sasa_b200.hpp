// sasa_b200.hpp -- C++ host layer over the C ABI (include/sasa_b200.h): RustSASA's public interface for the
// hot path, restated in C++ because the crate's own language (Rust) has no toolchain in this build image.
//
// Same names, argument meaning, defaults and error behaviour as the crate (rust-sasa 0.9.2):
//   rust_sasa::calculate_sasa_internal(atoms, probe_radius, n_points, threads)        src/lib.rs:249-298
//   rust_sasa::SASAOptions<AtomLevel|ResidueLevel|ChainLevel|ProteinLevel>            src/options.rs:60-76, :496-618
//       ::with_probe_radius / with_n_points / with_threads / with_include_hydrogens / with_include_hetatms /
//         with_radii_file / with_allow_vdw_fallback / with_read_radii_from_occupancy / process(pdb)
//   rust_sasa::{ChainResult, ResidueResult, ProteinResult, SASAResult}                src/structures/atomic.rs:26-70
//   rust_sasa::SASACalcError                                                          src/options.rs:466-494
//   rust_sasa::{get_radius, get_protor_radius, load_radii_from_file, serialize_chain_id}
//                                                                                     src/utils.rs:24-56, src/utils/consts.rs:31-91
//   rust_sasa::{sasa_result_to_json, sasa_result_to_xml, sasa_result_to_protein_object} src/utils/io.rs:11-64
//   rust_sasa::pdb::{PDB, open}   the part of pdbtbx's hierarchy the path reads       pdbtbx/src/read/**, structs/**
// plus `process_many`, the batched form that the CLI's directory mode (src/main.rs:342-480) maps onto: all
// structures of a tile go through ONE pipelined sasa_b200_batch_run_host call.
//
// Every numeric result comes out of libsasa_b200.so (CUDA, sm_100a).  There is no CPU fallback: without a device
// the first call throws SASACalcError{Kind::Device}.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <variant>
#include <vector>

namespace rust_sasa {

// ---- src/structures/atomic.rs -------------------------------------------------------------------------------
struct Atom {
    std::array<float, 3> position;
    float radius;
    std::size_t id;
    std::optional<std::ptrdiff_t> parent_id;
};

struct ChainResult {
    std::string name;
    float value;
};

struct ResidueResult {
    std::ptrdiff_t serial_number;
    std::string insertion_code;
    float value;
    std::string name;
    bool is_polar;
    std::string chain_id;
};

struct ProteinResult {
    float global_total, polar_total, non_polar_total;
};

// enum SASAResult { Atom(Vec<f32>), Residue(Vec<ResidueResult>), Chain(Vec<ChainResult>), Protein(ProteinResult) }
using SASAResult = std::variant<std::vector<float>, std::vector<ResidueResult>, std::vector<ChainResult>, ProteinResult>;

// ---- src/options.rs:466-494 ----------------------------------------------------------------------------------
class SASACalcError : public std::runtime_error {
public:
    enum class Kind {
        ElementMissing,
        VanDerWaalsMissing,
        RadiusMissing,
        AtomMapToLevelElementFailed,
        FailedToGetResidueName,
        RadiiFileLoad,
        Device,   // not in the reference: the CUDA engine reported an error (no device, non-finite input, ...)
    };
    SASACalcError(Kind kind, const std::string &message) : std::runtime_error(message), kind_(kind) {}
    Kind kind() const { return kind_; }
    const char *kind_name() const;

private:
    Kind kind_;
};

// ---- radii (src/utils/consts.rs:31-91, src/utils.rs:35-56) ---------------------------------------------------
using RadiiConfig = std::unordered_map<std::string, std::unordered_map<std::string, float>>;
RadiiConfig parse_radii_config(std::string_view content);
RadiiConfig load_radii_from_file(const std::string &path);            // throws SASACalcError{RadiiFileLoad}
const RadiiConfig &protor_radii();                                     // the embedded radii/protor.config
std::optional<float> get_protor_radius(const std::string &residue, const std::string &atom);
std::optional<float> get_radius(const std::string &residue, const std::string &atom, const RadiiConfig *custom);
std::ptrdiff_t serialize_chain_id(std::string_view s);                 // src/utils.rs:24-33 (lossy on purpose)
bool is_polar_residue(const std::string &name);                        // POLAR_AMINO_ACIDS, src/utils/consts.rs:7-16

// ---- the slice of pdbtbx::PDB the path reads -------------------------------------------------------------------
namespace pdb {

struct AtomRec {
    bool hetero;
    std::size_t serial;
    std::string name;
    double x, y, z;
    double occupancy;
    std::string element;   // upper-case symbol, empty = unknown
    double b_factor = 0.0; // what sasa_result_to_protein_object overwrites
    int charge = 0;
    std::string id;        // pdbtbx Atom::id: _atom_site.id (mmCIF) or the 0-based running atom count of the file (PDB)
};

struct Conformer {
    std::string name;      // residue name
    std::string altloc;    // empty = none
    std::vector<AtomRec> atoms;
};

struct Residue {
    std::ptrdiff_t serial;
    std::string icode;     // empty = none
    std::vector<Conformer> conformers;
    // Some(name) iff all conformers agree (pdbtbx Residue::name)
    std::optional<std::string> name() const;
};

struct Chain {
    std::string id;
    std::vector<Residue> residues;                                   // first-seen order
    std::unordered_map<std::string, std::size_t> residue_index;      // (serial, icode) -> position
};

struct Model {
    std::ptrdiff_t serial = 0;
    std::vector<Chain> chains;                                       // first-seen order
    std::unordered_map<std::string, std::size_t> chain_index;
    std::size_t atom_count = 0;
};

struct PDB {
    std::vector<Model> models;
    std::size_t atom_count() const;   // every atom of every conformer
};

PDB read_pdb(const std::string &path);      // throws std::runtime_error on I/O failure
PDB read_mmcif(const std::string &path);
PDB open(const std::string &path);          // by extension: .cif / .mmcif -> mmCIF, else PDB

// Coordinate-section writers after pdbtbx::save (pdbtbx/src/save/pdb.rs:507-613, save/mmcif.rs:225-412, StrictnessLevel::Loose):
// MODEL / ATOM / HETATM / TER / ENDMDL / END with pdbtbx's field rules, and the _atom_site loop with its columns.  Header
// records (HEADER, REMARK, CRYST1, SCALE, ...) are not kept by the reader above and are not written.
void save_pdb(const PDB &pdb, const std::string &path);     // throws std::runtime_error when the file cannot be written
void save_mmcif(const PDB &pdb, const std::string &path);
void save(const PDB &pdb, const std::string &path);         // by extension: .pdb / .cif / .mmcif, anything else is an error
std::string to_pdb_string(const PDB &pdb);
std::string to_mmcif_string(const PDB &pdb, const std::string &name = "sasa_b200");

}  // namespace pdb

// ---- src/lib.rs:249-298 ----------------------------------------------------------------------------------------
// `threads` is accepted for signature compatibility and ignored (the GPU path has no thread pool).
std::vector<float> calculate_sasa_internal(const std::vector<Atom> &atoms, float probe_radius = 1.4f,
                                           std::size_t n_points = 100, std::ptrdiff_t threads = -1);

// ---- src/options.rs ----------------------------------------------------------------------------------------------
enum class LevelKind { Atom, Residue, Chain, Protein };

struct AtomLevel {
    using Output = std::vector<float>;
    static constexpr LevelKind kind = LevelKind::Atom;
};
struct ResidueLevel {
    using Output = std::vector<ResidueResult>;
    static constexpr LevelKind kind = LevelKind::Residue;
};
struct ChainLevel {
    using Output = std::vector<ChainResult>;
    static constexpr LevelKind kind = LevelKind::Chain;
};
struct ProteinLevel {
    using Output = ProteinResult;
    static constexpr LevelKind kind = LevelKind::Protein;
};

// Level-independent option block (defaults: src/options.rs:498-510).
struct OptionValues {
    float probe_radius = 1.4f;
    std::size_t n_points = 100;
    std::ptrdiff_t threads = -1;
    bool include_hydrogens = false;
    std::shared_ptr<const RadiiConfig> radii_config;   // None = ProtOr only
    bool allow_vdw_fallback = false;
    bool include_hetatms = false;
    bool read_radii_from_occupancy = false;
};

// Output of build_atoms_and_mapping (row A0) in the wire format of the C ABI.
struct Packed {
    std::vector<float> xyzr;            // 4 floats per atom
    std::vector<std::uint64_t> ids;     // Atom.id (FNV-1a of (altloc, serial)); only equality matters
    std::vector<std::uint32_t> seg_be;  // 2 per segment: [begin, end) atom range
    std::vector<std::uint8_t> seg_polar;
    std::vector<ResidueResult> residue_meta;   // Residue / Protein levels (value filled in later)
    std::vector<ChainResult> chain_meta;       // Chain level
    std::size_t n_atoms() const { return ids.size(); }
};

Packed build_atoms_and_mapping(const pdb::PDB &pdb, LevelKind level, const OptionValues &opt);

// One entry per input structure: the level's result or the error that structure raised
// (directory mode logs and continues, src/main.rs:447-453).
using ProcessOutcome = std::variant<SASAResult, SASACalcError>;
std::vector<ProcessOutcome> process_many(const std::vector<const pdb::PDB *> &pdbs, LevelKind level, const OptionValues &opt);
// Same, from already extracted atoms (what a caller with its own parser uses).
// device < 0: the process's default engine device (set_device / SASA_B200_DEVICE / 0).  Calls on different devices run
// concurrently (one engine context, one staging buffer each): directory mode deals its tiles round-robin over the devices.
std::vector<ProcessOutcome> process_packed(const std::vector<const Packed *> &packed, LevelKind level, const OptionValues &opt,
                                           int device = -1);

template <class Level>
class SASAOptions {
public:
    static SASAOptions create() { return SASAOptions(); }   // SASAOptions::<Level>::new()
    SASAOptions &with_probe_radius(float radius) { opt_.probe_radius = radius; return *this; }
    SASAOptions &with_include_hetatms(bool v) { opt_.include_hetatms = v; return *this; }
    SASAOptions &with_n_points(std::size_t points) { opt_.n_points = points; return *this; }
    SASAOptions &with_read_radii_from_occupancy(bool v) { opt_.read_radii_from_occupancy = v; return *this; }
    SASAOptions &with_threads(std::ptrdiff_t threads) { opt_.threads = threads; return *this; }
    SASAOptions &with_include_hydrogens(bool v) { opt_.include_hydrogens = v; return *this; }
    SASAOptions &with_allow_vdw_fallback(bool v) { opt_.allow_vdw_fallback = v; return *this; }
    SASAOptions &with_radii_file(const std::string &path) {
        opt_.radii_config = std::make_shared<const RadiiConfig>(load_radii_from_file(path));
        return *this;
    }
    const OptionValues &values() const { return opt_; }

    // src/options.rs:606-618; throws SASACalcError where the reference returns Err.
    typename Level::Output process(const pdb::PDB &pdb) const {
        auto out = process_many({&pdb}, Level::kind, opt_);
        if (auto *err = std::get_if<SASACalcError>(&out[0])) throw *err;
        return std::get<typename Level::Output>(std::get<SASAResult>(std::move(out[0])));
    }
    std::vector<ProcessOutcome> process_many(const std::vector<const pdb::PDB *> &pdbs) const {
        return rust_sasa::process_many(pdbs, Level::kind, opt_);
    }

private:
    OptionValues opt_;
};

// ---- src/utils/io.rs:11-18 ---------------------------------------------------------------------------------------
std::string sasa_result_to_json(const SASAResult &result);   // serde_json::to_string of the externally tagged enum
std::string sasa_result_to_xml(const SASAResult &result);    // quick_xml::se::to_string
// src/utils/io.rs:20-64: writes the result into the B-factors of `original_pdb`, with the reference's own iteration rules
// (Atom: the i-th atom of pdb.atoms() over ALL atoms -- every conformer, hydrogens and HETATMs included -- gets v[i];
// Residue / Chain: the i-th residue / chain over all models, the serial number / chain id is checked; Protein: every atom
// gets global_total).  Throws std::runtime_error where the reference returns Err or panics (index past the result
// vector, mismatched residue, negative or non-finite value).
void sasa_result_to_protein_object(pdb::PDB &original_pdb, const SASAResult &result);

// Optional: create the engine context and the per-n_points tables now (e.g. on a helper thread while files are parsed).
void warm_up(const OptionValues &opt, int device = -1);
// CUDA devices visible to this process (0 when there is none).
int device_count();

// Engine selection for this process: CUDA device ordinal used by every call above (default: SASA_B200_DEVICE or 0).
void set_device(int device);

}  // namespace rust_sasa
