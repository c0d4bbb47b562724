// options.cpp -- the level API over the C ABI: atom extraction (row A0), the batched engine call, and the mapping of
// its outputs back to the reference's result structs.
//
//   build_atoms_and_mapping   src/options.rs:151-189 (Atom), :234-287 (Residue), :317-365 (Chain), :412-464 (Protein),
//                             build_atom! :81-116, get_radius src/utils.rs:40-56, combine_hash src/utils.rs:83-87
//   process / process_many    src/options.rs:606-618; directory mode src/main.rs:342-480
//   calculate_sasa_internal   src/lib.rs:249-298
// The numeric part of process_atoms (sequential f32 sums per residue / chain / protein) runs on the GPU inside the
// same launch (seg_sasa / protein outputs of sasa_b200_batch_run_host); strings and metadata stay here.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "../../../include/sasa_b200.h"
#include "../../../include/sasa_b200.hpp"

namespace rust_sasa {

namespace {

// Alvarez (2013) van der Waals radii as tabulated by pdbtbx/src/structs/elements.rs:631-647 ff.
const std::unordered_map<std::string, float> &vdw_radii() {
    static const std::unordered_map<std::string, float> t = {
        {"H", 1.2f},   {"HE", 1.43f}, {"LI", 2.12f}, {"BE", 1.98f}, {"B", 1.91f},  {"C", 1.77f},  {"N", 1.66f},  {"O", 1.5f},
        {"F", 1.46f},  {"NE", 1.58f}, {"NA", 2.5f},  {"MG", 2.51f}, {"AL", 2.25f}, {"SI", 2.19f}, {"P", 1.9f},   {"S", 1.89f},
        {"CL", 1.82f}, {"AR", 1.83f}, {"K", 2.73f},  {"CA", 2.62f}, {"SC", 2.58f}, {"TI", 2.46f}, {"V", 2.42f},  {"CR", 2.45f},
        {"MN", 2.45f}, {"FE", 2.44f}, {"CO", 2.4f},  {"NI", 2.4f},  {"CU", 2.38f}, {"ZN", 2.39f}, {"GA", 2.32f}, {"GE", 2.29f},
        {"AS", 1.88f}, {"SE", 1.82f}, {"BR", 1.86f}, {"KR", 2.25f}};
    return t;
}

// FNV-1a over the bytes Rust's `(&str, usize).hash()` feeds the hasher: the string bytes, a 0xff terminator, then the
// usize in little-endian order (src/utils.rs:83-87).  Only equality of ids is observable on the path.
std::uint64_t atom_id_hash(const std::string &altloc, std::size_t serial) {
    std::uint64_t h = 0xcbf29ce484222325ull;
    auto feed = [&h](unsigned char b) { h = (h ^ b) * 0x100000001b3ull; };
    for (unsigned char c : altloc) feed(c);
    feed(0xff);
    for (int i = 0; i < 8; ++i) feed((unsigned char)((std::uint64_t)serial >> (8 * i)));
    return h;
}

// One engine context and one grow-only pinned staging buffer per device, created on first use.
struct DeviceState {
    sasa_b200_ctx *ctx = nullptr;
    std::mutex pin_mu;   // one tile at a time owns the staging buffer of a device
    void *pin = nullptr;
    size_t pin_bytes = 0;
};
int g_device = -1;
std::mutex g_ctx_mu;
std::map<int, std::unique_ptr<DeviceState>> g_devices;

DeviceState &device_state(int device) {
    DeviceState *st = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (device < 0) device = g_device;
        if (device < 0) {
            const char *e = std::getenv("SASA_B200_DEVICE");
            device = e ? std::atoi(e) : 0;
        }
        auto &slot = g_devices[device];
        if (!slot) slot = std::make_unique<DeviceState>();
        st = slot.get();
    }
    // context creation (hundreds of milliseconds) happens outside the registry lock: devices come up in parallel
    std::lock_guard<std::mutex> lk(st->pin_mu);
    if (!st->ctx && sasa_b200_create(device, &st->ctx) != SASA_B200_OK) {
        st->ctx = nullptr;
        throw SASACalcError(SASACalcError::Kind::Device, std::string("sasa_b200_create failed: ") + sasa_b200_last_error(nullptr));
    }
    return *st;
}

sasa_b200_ctx *context(int device = -1) { return device_state(device).ctx; }

[[noreturn]] void throw_device(sasa_b200_ctx *ctx, const char *what) {
    throw SASACalcError(SASACalcError::Kind::Device, std::string(what) + ": " + sasa_b200_last_error(ctx));
}

// Dense equality classes of the ids of one structure, or empty when all ids are distinct (the usual case, decided by a
// sort of a copy; the hash-map ranking only runs for files that really repeat an id, e.g. multi-model files).
std::vector<std::uint32_t> id_classes(const std::vector<std::uint64_t> &ids, std::uint32_t base) {
    {
        std::vector<std::uint64_t> sorted(ids);
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end()) return {};
    }
    std::unordered_map<std::uint64_t, std::uint32_t> rank;
    rank.reserve(ids.size() * 2);
    std::vector<std::uint32_t> cls(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) cls[i] = base + rank.emplace(ids[i], (std::uint32_t)rank.size()).first->second;
    if (rank.size() == ids.size()) cls.clear();
    return cls;
}

// Host-side loops over the structures of a tile (packing, result assembly) on all cores.
template <class F>
void parallel_for(size_t n, F &&f) {
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)std::min<size_t>(n / 64 + 1, 64)));
    if (nt <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&] {
            for (size_t i0; (i0 = next.fetch_add(32)) < n;)
                for (size_t i = i0; i < std::min(n, i0 + 32); ++i) f(i);
        });
    for (auto &t : pool) t.join();
}

// Grow-only pinned staging buffer of a device (a cudaHostAlloc per tile cost tens of milliseconds); st.pin_mu is held.
void *pinned_reserve(DeviceState &st, size_t bytes) {
    if (bytes <= st.pin_bytes) return st.pin;
    if (st.pin) sasa_b200_free_pinned(st.pin);
    st.pin = nullptr;
    st.pin_bytes = 0;
    const size_t want = bytes + bytes / 4 + (1 << 20);
    if (sasa_b200_alloc_pinned(want, &st.pin) != SASA_B200_OK) return nullptr;
    st.pin_bytes = want;
    return st.pin;
}

}  // namespace

// Creates the engine context and the tables of this point count ahead of the first real call (CUDA context creation,
// module load and the cap table are a few hundred milliseconds): the CLI runs this while it parses its first files.
int device_count() {
    int n = 0;
    return sasa_b200_device_count(&n) == SASA_B200_OK ? n : 0;
}

void warm_up(const OptionValues &opt, int device) {
    const bool trace = std::getenv("SASA_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    sasa_b200_ctx *ctx = context(device);
    if (trace) std::fprintf(stderr, "[sasa_b200] context of device %d created at %.3f s\n", device, since());
    const float one[4] = {0.0f, 0.0f, 0.0f, 1.5f};
    float out = 0.0f;
    sasa_b200_calculate_sasa_internal(ctx, one, nullptr, 1, opt.probe_radius, opt.n_points, opt.threads, &out, nullptr);
    if (trace) std::fprintf(stderr, "[sasa_b200] first call (point set, cap table, first launch) done at %.3f s\n", since());
}

void set_device(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_device = device;
}

// ---- row A0 ---------------------------------------------------------------------------------------------------------
Packed build_atoms_and_mapping(const pdb::PDB &pdb, LevelKind level, const OptionValues &opt) {
    Packed out;
    const RadiiConfig *custom = opt.radii_config.get();
    {
        const size_t n = pdb.atom_count();
        out.xyzr.reserve(4 * n);
        out.ids.reserve(n);
    }
    // get_radius (src/utils.rs:40-56: custom[residue][atom], else ProtOr[residue][atom]) with the residue's two inner tables
    // looked up once per residue instead of once per atom
    using Inner = std::unordered_map<std::string, float>;
    const Inner *custom_res = nullptr, *protor_res = nullptr;
    auto set_residue = [&](const std::string &resname) {   // called once per residue, before its atoms are pushed
        custom_res = protor_res = nullptr;
        if (custom) {
            auto r = custom->find(resname);
            if (r != custom->end()) custom_res = &r->second;
        }
        auto r = protor_radii().find(resname);
        if (r != protor_radii().end()) protor_res = &r->second;
    };
    auto radius_of = [&](const std::string &atom) -> std::optional<float> {
        if (custom_res) {
            auto a = custom_res->find(atom);
            if (a != custom_res->end()) return a->second;
        }
        if (protor_res) {
            auto a = protor_res->find(atom);
            if (a != protor_res->end()) return a->second;
        }
        return std::nullopt;
    };
    auto push = [&](const pdb::AtomRec &a, const std::string &resname, const std::string &altloc) {
        if (a.element.empty()) throw SASACalcError(SASACalcError::Kind::ElementMissing, "Element missing for atom");
        if (a.element == "H" && !opt.include_hydrogens) return;
        if (a.hetero && !opt.include_hetatms) return;
        float radius;
        if (opt.read_radii_from_occupancy) {
            radius = (float)a.occupancy;
        } else if (auto r = radius_of(a.name)) {
            radius = *r;
        } else if (opt.allow_vdw_fallback) {
            auto it = vdw_radii().find(a.element);
            if (it == vdw_radii().end())
                throw SASACalcError(SASACalcError::Kind::VanDerWaalsMissing, "Van der Waals radius missing for element");
            radius = it->second;
        } else {
            throw SASACalcError(SASACalcError::Kind::RadiusMissing, "Radius not found for residue '" + resname + "' atom '" + a.name +
                                                                        "' of type '" + a.element + "'.");
        }
        out.xyzr.push_back((float)a.x);
        out.xyzr.push_back((float)a.y);
        out.xyzr.push_back((float)a.z);
        out.xyzr.push_back(radius);
        out.ids.push_back(atom_id_hash(altloc, a.serial));
    };
    auto residue_name = [](const pdb::Residue &r) {
        auto n = r.name();
        if (!n && !r.conformers.empty())
            throw SASACalcError(SASACalcError::Kind::FailedToGetResidueName, "Failed to get residue name");
        return n;
    };

    if (level == LevelKind::Atom) {
        for (const pdb::Model &m : pdb.models)
            for (const pdb::Chain &c : m.chains)
                for (const pdb::Residue &r : c.residues) {
                    auto name = residue_name(r);
                    if (r.conformers.empty()) continue;
                    const pdb::Conformer &conf = r.conformers[0];   // first conformer only (src/options.rs:255)
                    set_residue(*name);
                    for (const pdb::AtomRec &a : conf.atoms) push(a, *name, conf.altloc);
                }
        return out;
    }

    // key -> range with HashMap::insert semantics (last writer wins), then one segment per hierarchy element in order
    std::map<std::string, std::pair<std::uint32_t, std::uint32_t>> key_range;
    std::vector<std::string> seg_keys;
    for (const pdb::Model &m : pdb.models)
        for (const pdb::Chain &c : m.chains) {
            const std::uint32_t chain_begin = (std::uint32_t)out.n_atoms();
            for (const pdb::Residue &r : c.residues) {
                auto name = residue_name(r);
                const std::uint32_t begin = (std::uint32_t)out.n_atoms();
                const std::string rkey = c.id + "\x1f" + std::to_string(r.serial) + "\x1f" + r.icode;
                if (!r.conformers.empty()) {
                    const pdb::Conformer &conf = r.conformers[0];
                    // ProteinLevel hashes ("", serial) for the atom id (src/options.rs:453)
                    set_residue(*name);
                    const std::string no_altloc;
                    for (const pdb::AtomRec &a : conf.atoms) push(a, *name, level == LevelKind::Protein ? no_altloc : conf.altloc);
                    if (level != LevelKind::Chain) key_range[rkey] = {begin, (std::uint32_t)out.n_atoms()};
                }
                if (level != LevelKind::Chain) {
                    seg_keys.push_back(rkey);
                    const std::string nm = name ? *name : std::string();
                    out.residue_meta.push_back(ResidueResult{r.serial, r.icode, 0.0f, nm, is_polar_residue(nm), c.id});
                }
            }
            if (level == LevelKind::Chain) {
                const std::string ckey = std::to_string(serialize_chain_id(c.id));   // lossy key, collisions preserved
                key_range[ckey] = {chain_begin, (std::uint32_t)out.n_atoms()};
                seg_keys.push_back(ckey);
                out.chain_meta.push_back(ChainResult{c.id, 0.0f});
            }
        }
    out.seg_be.reserve(2 * seg_keys.size());
    for (size_t k = 0; k < seg_keys.size(); ++k) {
        auto it = key_range.find(seg_keys[k]);
        if (it == key_range.end())
            throw SASACalcError(SASACalcError::Kind::AtomMapToLevelElementFailed, "Failed to map atoms back to level element");
        out.seg_be.push_back(it->second.first);
        out.seg_be.push_back(it->second.second);
        out.seg_polar.push_back(level != LevelKind::Chain && out.residue_meta[k].is_polar ? 1 : 0);
    }
    return out;
}

// ---- the batched engine call ----------------------------------------------------------------------------------------
std::vector<ProcessOutcome> process_packed(const std::vector<const Packed *> &packed, LevelKind level, const OptionValues &opt,
                                           int device) {
    std::vector<ProcessOutcome> results;
    const size_t S = packed.size();
    std::vector<std::uint64_t> struct_off(S + 1, 0), seg_off(S + 1, 0);
    for (size_t s = 0; s < S; ++s) {
        struct_off[s + 1] = struct_off[s] + packed[s]->n_atoms();
        seg_off[s + 1] = seg_off[s] + packed[s]->seg_polar.size();
    }
    const size_t N = struct_off[S], G = seg_off[S];
    DeviceState &dev = device_state(device);
    sasa_b200_ctx *ctx = dev.ctx;
    if (S == 1 && N > 0) {
        // One structure (SASAOptions::process, src/options.rs:606-618): the engine's one-structure call runs on a stream and
        // workspace of its own, so callers on many threads overlap (the reference's directory mode, src/main.rs:375).
        const Packed &p = *packed[0];
        const std::vector<std::uint32_t> cls = id_classes(p.ids, 0);
        std::vector<float> atom(level == LevelKind::Atom ? N : 0), seg(G);
        float prot[3] = {0.0f, 0.0f, 0.0f};
        sasa_b200_params prm{opt.probe_radius, (std::uint32_t)opt.n_points, 8, (std::int32_t)opt.threads, 0};
        sasa_b200_outputs outs{nullptr, level == LevelKind::Atom ? atom.data() : nullptr,
                               (level == LevelKind::Residue || level == LevelKind::Chain) && G ? seg.data() : nullptr,
                               level == LevelKind::Protein ? prot : nullptr};
        if (sasa_b200_run_batch(ctx, p.xyzr.data(), cls.empty() ? nullptr : cls.data(), struct_off.data(), 1, G ? p.seg_be.data() : nullptr,
                                G ? seg_off.data() : nullptr, G ? p.seg_polar.data() : nullptr, &prm, &outs, nullptr) != SASA_B200_OK)
            throw_device(ctx, "run_batch");
        switch (level) {
            case LevelKind::Atom: results.emplace_back(SASAResult(std::move(atom))); break;
            case LevelKind::Residue: {
                std::vector<ResidueResult> v = p.residue_meta;
                for (size_t k = 0; k < v.size(); ++k) v[k].value = seg[k];
                results.emplace_back(SASAResult(std::move(v)));
                break;
            }
            case LevelKind::Chain: {
                std::vector<ChainResult> v = p.chain_meta;
                for (size_t k = 0; k < v.size(); ++k) v[k].value = seg[k];
                results.emplace_back(SASAResult(std::move(v)));
                break;
            }
            case LevelKind::Protein: results.emplace_back(SASAResult(ProteinResult{prot[0], prot[1], prot[2]})); break;
        }
        return results;
    }
    // pinned staging: the pipelined host entry point overlaps H2D, kernels and D2H across chunks
    std::lock_guard<std::mutex> pin_lock(dev.pin_mu);   // one tile at a time owns the staging buffer of a device
    const size_t in_bytes = N * 16, out_atom = level == LevelKind::Atom ? N * 4 : 0, out_seg = G * 4, out_prot = S * 12;
    void *pin = pinned_reserve(dev, in_bytes + out_atom + out_seg + out_prot + 64);
    if (!pin) throw_device(nullptr, "alloc_pinned");
    float *h_xyzr = static_cast<float *>(pin);
    float *h_atom = reinterpret_cast<float *>(static_cast<char *>(pin) + in_bytes);
    float *h_seg = reinterpret_cast<float *>(static_cast<char *>(pin) + in_bytes + out_atom);
    float *h_prot = reinterpret_cast<float *>(static_cast<char *>(pin) + in_bytes + out_atom + out_seg);
    std::vector<std::uint32_t> seg_be(2 * G), id_class;
    std::vector<std::uint8_t> polar(G);
    std::vector<std::vector<std::uint32_t>> dup(S);   // per structure: id classes, empty when all ids are distinct
    parallel_for(S, [&](size_t s) {
        const Packed &p = *packed[s];
        if (p.n_atoms()) std::memcpy(h_xyzr + 4 * struct_off[s], p.xyzr.data(), p.n_atoms() * 16);
        if (!p.seg_polar.empty()) {
            std::memcpy(seg_be.data() + 2 * seg_off[s], p.seg_be.data(), p.seg_be.size() * 4);
            std::memcpy(polar.data() + seg_off[s], p.seg_polar.data(), p.seg_polar.size());
        }
        dup[s] = id_classes(p.ids, (std::uint32_t)struct_off[s]);
    });
    bool any_dup = false;
    for (size_t s = 0; s < S; ++s) {
        if (dup[s].empty()) continue;
        if (!any_dup) {   // first structure with duplicate ids: every other atom gets a class of its own
            id_class.resize(N);
            for (size_t i = 0; i < N; ++i) id_class[i] = (std::uint32_t)i;
            any_dup = true;
        }
        std::memcpy(id_class.data() + struct_off[s], dup[s].data(), dup[s].size() * 4);
    }
    sasa_b200_batch *batch = nullptr;
    if (sasa_b200_batch_create(ctx, struct_off.data(), S, G ? seg_be.data() : nullptr, G ? seg_off.data() : nullptr,
                               G ? polar.data() : nullptr, &batch) != SASA_B200_OK)
        throw_device(ctx, "batch_create");
    sasa_b200_params prm{opt.probe_radius, (std::uint32_t)opt.n_points, 8, (std::int32_t)opt.threads, 0};
    sasa_b200_outputs outs{nullptr, level == LevelKind::Atom ? h_atom : nullptr,
                           (level == LevelKind::Residue || level == LevelKind::Chain) && G ? h_seg : nullptr,
                           level == LevelKind::Protein ? h_prot : nullptr};
    const int rc = sasa_b200_batch_run_host(batch, h_xyzr, any_dup ? id_class.data() : nullptr, &prm, &outs, nullptr);
    sasa_b200_batch_destroy(batch);
    if (rc != SASA_B200_OK) throw_device(ctx, "batch_run_host");
    results.assign(S, ProcessOutcome(SASAResult(std::vector<float>())));
    parallel_for(S, [&](size_t s) {
        const Packed &p = *packed[s];
        switch (level) {
            case LevelKind::Atom:
                results[s] = SASAResult(std::vector<float>(h_atom + struct_off[s], h_atom + struct_off[s + 1]));
                break;
            case LevelKind::Residue: {
                std::vector<ResidueResult> v = p.residue_meta;
                for (size_t k = 0; k < v.size(); ++k) v[k].value = h_seg[seg_off[s] + k];
                results[s] = SASAResult(std::move(v));
                break;
            }
            case LevelKind::Chain: {
                std::vector<ChainResult> v = p.chain_meta;
                for (size_t k = 0; k < v.size(); ++k) v[k].value = h_seg[seg_off[s] + k];
                results[s] = SASAResult(std::move(v));
                break;
            }
            case LevelKind::Protein:
                results[s] = SASAResult(ProteinResult{h_prot[3 * s], h_prot[3 * s + 1], h_prot[3 * s + 2]});
                break;
        }
    });
    return results;
}

std::vector<ProcessOutcome> process_many(const std::vector<const pdb::PDB *> &pdbs, LevelKind level, const OptionValues &opt) {
    std::vector<std::optional<Packed>> packed(pdbs.size());
    std::vector<std::optional<SASACalcError>> errors(pdbs.size());
    std::vector<const Packed *> good;
    for (size_t i = 0; i < pdbs.size(); ++i) {
        try {
            packed[i] = build_atoms_and_mapping(*pdbs[i], level, opt);
            good.push_back(&*packed[i]);
        } catch (const SASACalcError &e) {
            errors[i] = e;
        }
    }
    std::vector<ProcessOutcome> computed = good.empty() ? std::vector<ProcessOutcome>() : process_packed(good, level, opt);
    std::vector<ProcessOutcome> out;
    out.reserve(pdbs.size());
    size_t g = 0;
    for (size_t i = 0; i < pdbs.size(); ++i) {
        if (errors[i]) out.emplace_back(*errors[i]);
        else out.emplace_back(std::move(computed[g++]));
    }
    return out;
}

// ---- src/lib.rs:249-298 -----------------------------------------------------------------------------------------------
std::vector<float> calculate_sasa_internal(const std::vector<Atom> &atoms, float probe_radius, std::size_t n_points,
                                           std::ptrdiff_t threads) {
    if (atoms.empty()) return {};   // tests/sanity.rs:148-157
    std::vector<float> xyzr(4 * atoms.size());
    std::vector<std::uint64_t> ids(atoms.size());
    for (size_t i = 0; i < atoms.size(); ++i) {
        xyzr[4 * i + 0] = atoms[i].position[0];
        xyzr[4 * i + 1] = atoms[i].position[1];
        xyzr[4 * i + 2] = atoms[i].position[2];
        xyzr[4 * i + 3] = atoms[i].radius;
        ids[i] = (std::uint64_t)atoms[i].id;
    }
    std::vector<float> out(atoms.size());
    sasa_b200_ctx *ctx = context();
    if (sasa_b200_calculate_sasa_internal(ctx, xyzr.data(), ids.data(), atoms.size(), probe_radius, n_points, threads, out.data(),
                                          nullptr) != SASA_B200_OK)
        throw_device(ctx, "calculate_sasa_internal");   // the reference panics on non-finite input
    return out;
}

}  // namespace rust_sasa
