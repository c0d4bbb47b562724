// cli.cpp -- `sasa_b200_cli`: the reference CLI's two modes (src/main.rs) over the batched GPU engine.
//
//   sasa_b200_cli <input> <output> [-o atom|residue|chain|protein] [-f json|xml] [-n N] [-p R] [-H] [-r FILE] [-a]
//                 [-e] [-t T] [-R]
// Flags, defaults and behaviour follow src/main.rs:56-106: a directory input requires --format (:554-559) and is
// processed as ONE batched pipeline -- parse in parallel -> pack -> one engine call per tile of structures -> write
// in parallel (what src/main.rs:342-480 does with a rayon par_iter over files and a single-threaded engine call
// each); per-file failures are collected, reported at the end and do not change the exit code (:447-479); outputs
// are named {stem}.{ext} (:414-416).  A single-file failure exits non-zero.  `-t` is accepted and ignored.
// Output formats: json, xml, and pdb / cif (B-factor write-back, src/utils/io.rs:20-64 + the coordinate writers of writers.cpp).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>

#include "../../../include/sasa_b200.hpp"

namespace fs = std::filesystem;
using namespace rust_sasa;

namespace {

struct Args {
    std::string input, output, depth = "residue", format;
    std::size_t n_points = 100;
    float probe_radius = 1.4f;
    bool include_hydrogens = false, allow_vdw_fallback = false, include_hetatms = false, read_radii_from_occupancy = false;
    std::string radii_file;
    std::ptrdiff_t threads = -1;
    std::size_t tile = 4096;   // structures per engine call in directory mode
};

[[noreturn]] void usage(const char *msg) {
    if (msg) std::fprintf(stderr, "error: %s\n\n", msg);
    std::fprintf(stderr,
                 "Usage: sasa_b200_cli [OPTIONS] <INPUT> <OUTPUT>\n"
                 "  -o, --output-depth <atom|residue|chain|protein>   [default: residue]\n"
                 "  -f, --format <json|xml|pdb|cif> required for directories, else inferred from the extension (default json)\n"
                 "  -n, --n-points <N>             [default: 100]\n"
                 "  -p, --probe-radius <R>         [default: 1.4]\n"
                 "  -H, --include-hydrogens\n  -r, --radii-file <FILE>\n  -a, --allow-vdw-fallback\n  -e, --include-hetatms\n"
                 "  -t, --threads <T>              accepted, ignored (GPU path)\n  -R, --read-radii-from-occupancy\n");
    std::exit(2);
}

Args parse(int argc, char **argv) {
    Args a;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string s = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) usage(("missing value for " + s).c_str());
            return argv[++i];
        };
        if (s == "-o" || s == "--output-depth") a.depth = val();
        else if (s == "-f" || s == "--format") a.format = val();
        else if (s == "-n" || s == "--n-points") a.n_points = std::stoul(val());
        else if (s == "-p" || s == "--probe-radius") a.probe_radius = std::stof(val());
        else if (s == "-H" || s == "--include-hydrogens") a.include_hydrogens = true;
        else if (s == "-r" || s == "--radii-file") a.radii_file = val();
        else if (s == "-a" || s == "--allow-vdw-fallback") a.allow_vdw_fallback = true;
        else if (s == "-e" || s == "--include-hetatms") a.include_hetatms = true;
        else if (s == "-t" || s == "--threads") a.threads = std::stol(val());
        else if (s == "-R" || s == "--read-radii-from-occupancy") a.read_radii_from_occupancy = true;
        else if (s == "--tile") a.tile = std::stoul(val());
        else if (s == "-h" || s == "--help") usage(nullptr);
        else if (!s.empty() && s[0] == '-' && s.size() > 1 && !std::isdigit((unsigned char)s[1])) usage(("unknown option " + s).c_str());
        else pos.push_back(s);
    }
    if (pos.size() != 2) usage("expected <INPUT> and <OUTPUT>");
    a.input = pos[0];
    a.output = pos[1];
    return a;
}

LevelKind level_of(const std::string &d) {
    if (d == "atom") return LevelKind::Atom;
    if (d == "residue") return LevelKind::Residue;
    if (d == "chain") return LevelKind::Chain;
    if (d == "protein") return LevelKind::Protein;
    usage("output depth must be atom, residue, chain or protein");
}

bool known_format(const std::string &f) { return f == "json" || f == "xml" || f == "pdb" || f == "cif"; }
bool structure_format(const std::string &f) { return f == "pdb" || f == "cif"; }

// src/main.rs:203-226: xml / json text, or the input structure with the result written into its B-factors
std::string render(const SASAResult &r, const std::string &format, const pdb::PDB *original) {
    if (format == "xml") return sasa_result_to_xml(r);
    if (!structure_format(format)) return sasa_result_to_json(r);
    pdb::PDB copy = *original;
    sasa_result_to_protein_object(copy, r);
    return format == "pdb" ? pdb::to_pdb_string(copy) : pdb::to_mmcif_string(copy, "?");
}

bool write_file(const fs::path &p, const std::string &text, std::string *err) {
    std::ofstream fh(p, std::ios::binary);
    if (!fh) { *err = "cannot write " + p.string(); return false; }
    fh << text;
    return (bool)fh;
}

template <class F>
void parallel_for(size_t n, F &&f) {
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)n));
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&] {
            for (size_t i; (i = next.fetch_add(1)) < n;) f(i);
        });
    for (auto &t : pool) t.join();
}

}  // namespace

int main(int argc, char **argv) {
    const Args args = parse(argc, argv);
    const LevelKind level = level_of(args.depth);
    OptionValues opt;
    opt.probe_radius = args.probe_radius;
    opt.n_points = args.n_points;
    opt.threads = args.threads;
    opt.include_hydrogens = args.include_hydrogens;
    opt.allow_vdw_fallback = args.allow_vdw_fallback;
    opt.include_hetatms = args.include_hetatms;
    opt.read_radii_from_occupancy = args.read_radii_from_occupancy;
    try {
        if (!args.radii_file.empty()) opt.radii_config = std::make_shared<const RadiiConfig>(load_radii_from_file(args.radii_file));
    } catch (const SASACalcError &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    std::error_code ec;
    if (!fs::exists(args.input, ec)) {
        std::fprintf(stderr, "error: input '%s' does not exist\n", args.input.c_str());
        return 1;
    }
    std::string format = args.format;
    if (fs::is_directory(args.input, ec)) {
        if (format.empty()) {
            std::fprintf(stderr, "error: --format is required when processing a directory\n");
            return 1;
        }
        if (!known_format(format)) {
            std::fprintf(stderr, "error: invalid value '%s' for '--format' (json, xml, pdb, cif)\n", format.c_str());
            return 1;
        }
        fs::create_directories(args.output, ec);
        if (ec || !fs::is_directory(args.output)) {
            std::fprintf(stderr, "error: cannot create output directory '%s'\n", args.output.c_str());
            return 1;
        }
        std::vector<fs::path> files;
        for (const auto &entry : fs::directory_iterator(args.input))
            if (entry.is_regular_file()) files.push_back(entry.path());
        std::sort(files.begin(), files.end());
        std::mutex err_mu;
        std::vector<std::string> errors;
        auto add_error = [&](const std::string &m) { std::lock_guard<std::mutex> lk(err_mu); errors.push_back(m); };
        const auto t0 = std::chrono::steady_clock::now();
        size_t atoms_total = 0;
        double t_parse = 0.0, t_engine = 0.0, t_write = 0.0;
        auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
        // the engine context and its tables come up while the first tile is being parsed
        double t_warm = 0.0;
        std::thread warm([&] { try { warm_up(opt); } catch (...) {} t_warm = since(t0); });
        for (size_t f0 = 0; f0 < files.size(); f0 += args.tile) {
            const size_t f1 = std::min(files.size(), f0 + args.tile), n = f1 - f0;
            // 1. parse + extract in parallel
            auto tp = std::chrono::steady_clock::now();
            std::vector<std::optional<Packed>> packed(n);
            std::vector<std::optional<pdb::PDB>> kept(structure_format(format) ? n : 0);   // only the write-back formats need them
            parallel_for(n, [&](size_t i) {
                const fs::path &p = files[f0 + i];
                try {
                    pdb::PDB st = pdb::open(p.string());
                    packed[i] = build_atoms_and_mapping(st, level, opt);
                    if (!kept.empty()) kept[i] = std::move(st);
                } catch (const std::exception &e) {
                    add_error("Error processing " + p.stem().string() + ": " + e.what());
                }
            });
            t_parse += since(tp);
            // 2. one engine call for the tile
            if (warm.joinable()) warm.join();
            tp = std::chrono::steady_clock::now();
            std::vector<const Packed *> good;
            std::vector<size_t> good_idx;
            for (size_t i = 0; i < n; ++i)
                if (packed[i]) { good.push_back(&*packed[i]); good_idx.push_back(i); atoms_total += packed[i]->n_atoms(); }
            if (good.empty()) { t_engine += since(tp); continue; }
            std::vector<ProcessOutcome> out;
            try {
                out = process_packed(good, level, opt);
            } catch (const SASACalcError &e) {
                // a device-level failure (e.g. a non-finite coordinate somewhere in the tile): retry one by one so that
                // only the offending files are reported
                out.clear();
                for (const Packed *p : good) {
                    try { out.push_back(process_packed({p}, level, opt)[0]); }
                    catch (const SASACalcError &e1) { out.emplace_back(e1); }
                }
            }
            t_engine += since(tp);
            // 3. serialise + write in parallel
            tp = std::chrono::steady_clock::now();
            parallel_for(good.size(), [&](size_t k) {
                const fs::path &p = files[f0 + good_idx[k]];
                if (auto *err = std::get_if<SASACalcError>(&out[k])) {
                    add_error("Error processing " + p.stem().string() + ": " + err->what());
                    return;
                }
                std::string werr;
                try {
                    const pdb::PDB *orig = kept.empty() ? nullptr : &*kept[good_idx[k]];
                    if (!write_file(fs::path(args.output) / (p.stem().string() + "." + format),
                                    render(std::get<SASAResult>(out[k]), format, orig), &werr))
                        add_error("Error processing " + p.stem().string() + ": " + werr);
                } catch (const std::exception &e) {   // CLIError::ProteinSerialization
                    add_error("Error processing " + p.stem().string() + ": " + e.what());
                }
            });
            t_write += since(tp);
        }
        if (warm.joinable()) warm.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!errors.empty()) {
            std::fprintf(stderr, "\nThe following errors occurred during processing:\n");
            for (const auto &e : errors) std::fprintf(stderr, "  - %s\n", e.c_str());
            std::fprintf(stderr, "\nTotal errors: %zu\n", errors.size());
        } else {
            std::printf("All files processed successfully!\n");
        }
        std::printf("%zu files, %zu atoms in %.3f s (%.2f M atoms/s end to end incl. parsing and writing; parse+extract %.3f s, "
                    "pack+engine %.3f s, serialise+write %.3f s; engine start-up, overlapped with the first parse: %.3f s)\n", files.size(),
                    atoms_total, dt, atoms_total / dt / 1e6, t_parse, t_engine, t_write, t_warm);
        return 0;
    }
    // single-file mode (src/main.rs:483-523)
    if (format.empty()) {
        std::string ext = fs::path(args.output).extension().string();
        if (!ext.empty()) ext.erase(0, 1);
        format = known_format(ext) ? ext : "json";   // OutputFormat::from_file_extension: anything else is JSON (src/main.rs:45-53)
    }
    if (!known_format(format)) {
        std::fprintf(stderr, "error: invalid value '%s' for '--format' (json, xml, pdb, cif)\n", format.c_str());
        return 1;
    }
    if (fs::is_directory(args.output, ec)) {
        std::fprintf(stderr, "error: output path '%s' is a directory\n", args.output.c_str());
        return 1;
    }
    try {
        const pdb::PDB st = pdb::open(args.input);
        auto out = process_many({&st}, level, opt);
        if (auto *err = std::get_if<SASACalcError>(&out[0])) throw *err;
        std::string werr;
        if (!write_file(args.output, render(std::get<SASAResult>(out[0]), format, &st), &werr)) throw std::runtime_error(werr);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
