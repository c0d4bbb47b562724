// cli.cpp -- `sasa_b200_cli`: the reference CLI's two modes (src/main.rs) over the batched GPU engine.
//
//   sasa_b200_cli <input> <output> [-o atom|residue|chain|protein] [-f json|xml] [-n N] [-p R] [-H] [-r FILE] [-a]
//                 [-e] [-t T] [-R]
// Flags, defaults and behaviour follow src/main.rs:56-106: a directory input requires --format (:554-559) and is
// processed as ONE batched pipeline -- parse in parallel -> pack -> one engine call per tile of structures -> write
// in parallel (what src/main.rs:342-480 does with a rayon par_iter over files and a single-threaded engine call
// each); per-file failures are collected, reported at the end and do not change the exit code (:447-479); outputs
// are named {stem}.{ext} (:414-416).  A single-file failure exits non-zero.  `-t` is accepted and ignored.
// Directory mode is a three-stage pipeline over tiles of files: every host core parses and extracts; one engine thread per
// GPU (--devices N | all; default 1) takes finished tiles round-robin and makes ONE engine call per tile; results are
// serialised and written in parallel.  The engine contexts come up while the first tiles are being parsed.
// Output formats: json, xml, and pdb / cif (B-factor write-back, src/utils/io.rs:20-64 + the coordinate writers of writers.cpp).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
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
    std::size_t tile = 512;    // structures per engine call in directory mode
    int devices = 1;           // GPUs used by directory mode (0 = all)
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
                 "  -t, --threads <T>              accepted, ignored (GPU path)\n  -R, --read-radii-from-occupancy\n"
                 "      --devices <N|all>          GPUs used in directory mode [default: 1]\n"
                 "      --tile <N>                 files per engine call in directory mode [default: 512]\n");
    std::exit(2);
}

// Numeric option values: a malformed or out-of-range number is a usage error (clap's behaviour in the reference: message and
// exit code 2), never an uncaught exception.
template <class T, class F>
T number(const std::string &opt, const std::string &text, F conv) {
    try {
        size_t used = 0;
        const T v = conv(text, &used);
        if (used != text.size()) throw std::invalid_argument(text);
        return v;
    } catch (const std::exception &) {
        usage(("invalid value '" + text + "' for '" + opt + "'").c_str());
    }
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
        auto as_size = [&](const std::string &t) {
            if (!t.empty() && t[0] == '-') usage(("invalid value '" + t + "' for '" + s + "'").c_str());
            return number<unsigned long>(s, t, [](const std::string &x, size_t *u) { return std::stoul(x, u); });
        };
        if (s == "-o" || s == "--output-depth") a.depth = val();
        else if (s == "-f" || s == "--format") a.format = val();
        else if (s == "-n" || s == "--n-points") {
            a.n_points = as_size(val());
            if (a.n_points < 1) usage("--n-points must be at least 1");
        }
        else if (s == "-p" || s == "--probe-radius") a.probe_radius = number<float>(s, val(), [](const std::string &x, size_t *u) { return std::stof(x, u); });
        else if (s == "-H" || s == "--include-hydrogens") a.include_hydrogens = true;
        else if (s == "-r" || s == "--radii-file") a.radii_file = val();
        else if (s == "-a" || s == "--allow-vdw-fallback") a.allow_vdw_fallback = true;
        else if (s == "-e" || s == "--include-hetatms") a.include_hetatms = true;
        else if (s == "-t" || s == "--threads") a.threads = number<long>(s, val(), [](const std::string &x, size_t *u) { return std::stol(x, u); });
        else if (s == "-R" || s == "--read-radii-from-occupancy") a.read_radii_from_occupancy = true;
        else if (s == "--tile") {
            a.tile = as_size(val());
            if (a.tile < 1) usage("--tile must be at least 1");
        }
        else if (s == "--devices") {
            const std::string v = val();
            a.devices = v == "all" ? 0 : (int)as_size(v);
        }
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
        auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
        // devices: one engine thread each; with a single device the process hides the others from CUDA before its first
        // CUDA call (initialising eight GPUs to use one costs most of a second of start-up)
        int n_dev = args.devices;
        if (n_dev == 1 && !std::getenv("CUDA_VISIBLE_DEVICES")) {
            const char *e = std::getenv("SASA_B200_DEVICE");
            setenv("CUDA_VISIBLE_DEVICES", e ? e : "0", 1);
            if (e) setenv("SASA_B200_DEVICE", "0", 1);
        }
        if (n_dev != 1) {
            const int have = device_count();
            n_dev = n_dev == 0 ? have : std::min(n_dev, have);
            if (n_dev < 1) {
                std::fprintf(stderr, "error: no CUDA device available (this program has no CPU fallback)\n");
                return 1;
            }
        }
        const size_t n_files = files.size(), n_tiles = (n_files + args.tile - 1) / args.tile;
        const bool keep_structures = structure_format(format);   // only the write-back formats need the hierarchy again
        std::vector<std::optional<Packed>> packed(n_files);
        std::vector<std::optional<pdb::PDB>> kept(keep_structures ? n_files : 0);
        // stage 1 -> 2: a tile is ready when all of its files are parsed; parsers stay at most kAhead tiles in front of the engines
        std::vector<std::atomic<size_t>> left(n_tiles);
        for (size_t t = 0; t < n_tiles; ++t) left[t] = std::min(n_files, (t + 1) * args.tile) - t * args.tile;
        std::mutex mu;
        std::condition_variable cv_ready, cv_room;
        std::vector<char> ready(n_tiles, 0);
        size_t tiles_done = 0;
        const size_t kAhead = std::max<size_t>(4 * (size_t)n_dev, 8 + (size_t)(2000000 / std::max<size_t>(1, args.tile * 2400)));
        std::atomic<size_t> next_file{0}, atoms_total{0};
        std::atomic<long long> parse_us{0}, engine_us{0}, write_us{0};
        double t_warm = 0.0, t_parse_done = 0.0;
        auto parser = [&] {
            for (size_t i; (i = next_file.fetch_add(1)) < n_files;) {
                const size_t t = i / args.tile;
                {   // bounded run-ahead
                    std::unique_lock<std::mutex> lk(mu);
                    cv_room.wait(lk, [&] { return t < tiles_done + kAhead; });
                }
                const auto tp = std::chrono::steady_clock::now();
                const fs::path &p = files[i];
                try {
                    pdb::PDB st = pdb::open(p.string());
                    packed[i] = build_atoms_and_mapping(st, level, opt);
                    if (keep_structures) kept[i] = std::move(st);
                } catch (const std::exception &e) {
                    add_error("Error processing " + p.stem().string() + ": " + e.what());
                }
                parse_us += (long long)(since(tp) * 1e6);
                if (left[t].fetch_sub(1) == 1) {
                    std::lock_guard<std::mutex> lk(mu);
                    ready[t] = 1;
                    cv_ready.notify_all();
                }
            }
        };
        // stage 2 + 3: engine thread of device d takes tiles d, d + n_dev, ... in order
        auto engine = [&](int d) {
            try { warm_up(opt, n_dev == 1 ? -1 : d); } catch (...) {}
            if (d == 0) t_warm = since(t0);
            for (size_t t = (size_t)d; t < n_tiles; t += (size_t)n_dev) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_ready.wait(lk, [&] { return ready[t] != 0; });
                }
                const size_t f0 = t * args.tile, f1 = std::min(n_files, f0 + args.tile);
                auto tp = std::chrono::steady_clock::now();
                std::vector<const Packed *> good;
                std::vector<size_t> good_idx;
                size_t atoms = 0;
                for (size_t i = f0; i < f1; ++i)
                    if (packed[i]) { good.push_back(&*packed[i]); good_idx.push_back(i); atoms += packed[i]->n_atoms(); }
                atoms_total += atoms;
                std::vector<ProcessOutcome> out;
                if (!good.empty()) {
                    try {
                        out = process_packed(good, level, opt, n_dev == 1 ? -1 : d);
                    } catch (const SASACalcError &e) {
                        // a device-level failure (e.g. a non-finite coordinate somewhere in the tile): retry one by one so that
                        // only the offending files are reported
                        out.clear();
                        for (const Packed *p : good) {
                            try { out.push_back(process_packed({p}, level, opt, n_dev == 1 ? -1 : d)[0]); }
                            catch (const SASACalcError &e1) { out.emplace_back(e1); }
                        }
                    }
                }
                engine_us += (long long)(since(tp) * 1e6);
                tp = std::chrono::steady_clock::now();
                parallel_for(good.size(), [&](size_t k) {
                    const fs::path &p = files[good_idx[k]];
                    if (auto *err = std::get_if<SASACalcError>(&out[k])) {
                        add_error("Error processing " + p.stem().string() + ": " + err->what());
                        return;
                    }
                    std::string werr;
                    try {
                        const pdb::PDB *orig = keep_structures ? &*kept[good_idx[k]] : nullptr;
                        if (!write_file(fs::path(args.output) / (p.stem().string() + "." + format),
                                        render(std::get<SASAResult>(out[k]), format, orig), &werr))
                            add_error("Error processing " + p.stem().string() + ": " + werr);
                    } catch (const std::exception &e) {   // CLIError::ProteinSerialization
                        add_error("Error processing " + p.stem().string() + ": " + e.what());
                    }
                });
                write_us += (long long)(since(tp) * 1e6);
                for (size_t i = f0; i < f1; ++i) {   // the tile's memory goes back before the next one is taken
                    packed[i].reset();
                    if (keep_structures) kept[i].reset();
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    ++tiles_done;
                    cv_room.notify_all();
                }
            }
        };
        std::vector<std::thread> engines, parsers;
        for (int d = 0; d < n_dev; ++d) engines.emplace_back(engine, d);
        const unsigned n_parsers = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)std::max<size_t>(1, n_files)));
        for (unsigned k = 0; k < n_parsers; ++k) parsers.emplace_back(parser);
        for (auto &t : parsers) t.join();
        t_parse_done = since(t0);
        for (auto &t : engines) t.join();
        const double t_parse = parse_us / 1e6 / n_parsers, t_engine = engine_us / 1e6 / n_dev, t_write = write_us / 1e6 / n_dev;
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!errors.empty()) {
            std::fprintf(stderr, "\nThe following errors occurred during processing:\n");
            for (const auto &e : errors) std::fprintf(stderr, "  - %s\n", e.c_str());
            std::fprintf(stderr, "\nTotal errors: %zu\n", errors.size());
        } else {
            std::printf("All files processed successfully!\n");
        }
        std::printf("%zu files, %zu atoms in %.3f s (%.2f M atoms/s end to end incl. parsing and writing) on %d GPU(s), %u parser threads, "
                    "%zu tiles: parse+extract %.3f s per thread (all parsed at %.3f s), pack+engine %.3f s and serialise+write %.3f s per "
                    "engine thread; engine start-up, overlapped with parsing: %.3f s\n", files.size(), atoms_total.load(), dt,
                    atoms_total.load() / dt / 1e6, n_dev, n_parsers, n_tiles, t_parse, t_parse_done, t_engine, t_write, t_warm);
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
