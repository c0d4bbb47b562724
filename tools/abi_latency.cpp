// abi_latency.cpp -- the C ABI driven straight from C++ threads: latency of sasa_b200_calculate_sasa_internal for one caller
// and the call rate of T concurrent callers on ONE context (the reference's directory mode calls the engine from every
// rayon worker, src/main.rs:375, :439).  Prints one JSON line.
//   g++ -O2 -std=c++17 -pthread -Iinclude tools/abi_latency.cpp -Lrustsasa_b200 -lsasa_b200 -Wl,-rpath,$PWD/rustsasa_b200 -o /tmp/abi_latency
//   /tmp/abi_latency [n_atoms=2622] [threads=16] [reps=400]
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "sasa_b200.h"

static std::vector<float> globule(size_t n, unsigned seed) {
    // jittered cubic lattice at protein number density (0.057 atoms / A^3), carved to a ball
    std::mt19937 rng(seed);
    std::normal_distribution<float> jit(0.0f, 0.25f);
    const float a = std::cbrt(1.0f / 0.057f);
    const float R = std::cbrt(3.0f * n / (4.0f * 3.14159265f * 0.057f));
    const int m = (int)std::ceil(R / a) + 1;
    const float radii[4] = {1.88f, 1.61f, 1.64f, 1.42f};
    std::vector<float> v;
    for (int i = -m; i <= m && v.size() < 4 * n; ++i)
        for (int j = -m; j <= m && v.size() < 4 * n; ++j)
            for (int k = -m; k <= m && v.size() < 4 * n; ++k) {
                const float x = i * a, y = j * a, z = k * a;
                if (x * x + y * y + z * z > R * R) continue;
                v.push_back(x + jit(rng)); v.push_back(y + jit(rng)); v.push_back(z + jit(rng));
                v.push_back(radii[rng() & 3]);
            }
    return v;
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 2622;
    const int threads = argc > 2 ? atoi(argv[2]) : 16;
    const int reps = argc > 3 ? atoi(argv[3]) : 400;
    sasa_b200_ctx *ctx = nullptr;
    if (sasa_b200_create(0, &ctx) != SASA_B200_OK) {
        std::fprintf(stderr, "create failed: %s\n", sasa_b200_last_error(nullptr));
        return 1;
    }
    std::vector<std::vector<float>> xs;
    for (int t = 0; t < threads; ++t) xs.push_back(globule(n, 1000 + t));
    std::atomic<int> bad{0};
    auto loop = [&](int t, int r) {
        const size_t na = xs[t].size() / 4;
        std::vector<float> out(na);
        for (int i = 0; i < r; ++i)
            if (sasa_b200_calculate_sasa_internal(ctx, xs[t].data(), nullptr, na, 1.4f, 100, -1, out.data(), nullptr) != SASA_B200_OK) ++bad;
    };
    auto run = [&](int nt, int r) {
        std::vector<std::thread> th;
        const auto t0 = std::chrono::steady_clock::now();
        for (int t = 0; t < nt; ++t) th.emplace_back(loop, t, r);
        for (auto &x : th) x.join();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    run(threads, 8);   // warm-up: point set, cap table, one slot per thread
    const double t1 = run(1, reps);
    const double tn = run(threads, reps);
    const double single = reps / t1, many = (double)threads * reps / tn;
    std::printf("{\"atoms\": %zu, \"threads\": %d, \"us_per_call_single\": %.1f, \"calls_per_s_single\": %.0f, "
                "\"calls_per_s_concurrent\": %.0f, \"speedup\": %.2f, \"errors\": %d}\n",
                xs[0].size() / 4, threads, 1e6 / single, single, many, many / single, bad.load());
    sasa_b200_destroy(ctx);
    return bad.load() ? 2 : 0;
}
