// cap_table_check.cu -- host-only check of the cap tables of sasa_cap.cuh (built by tests/test_cap_table.py with nvcc, run on the
// CPU).  For random directions v^ and levels c it looks the bin up with the kernel's binning arithmetic (in float, exact
// reciprocals -- the bins' margins cover the approximate ones) and verifies against brute force in double precision the one
// property the exactness argument rests on:   inner bit  =>  p . v^ < c   and   p . v^ < c  =>  inner or ring bit,
// i.e. the table never decides a point that the reference's test could decide differently.
// usage: cap_table_check n_points samples     (n_points <= 128: the single table; above: the chunked tables, 64 x 64 x 66)
#include "../../rustsasa_b200/csrc/sasa_cap.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace sasa;

static void points(uint32_t n, std::vector<float> &h) {   // generate_sphere_points, src/lib.rs:43-66
    h.resize(3 * (size_t)n);
    const float inc = (2.0f * 3.14159265358979323846f) * 1.618034f, inv = 1.0f / (float)n;
    for (uint32_t i = 0; i < n; ++i) {
        const float fi = (float)i, incl = std::acos(1.0f - 2.0f * (fi * inv)), az = inc * fi, si = std::sin(incl);
        h[i] = si * std::cos(az); h[n + i] = si * std::sin(az); h[2 * (size_t)n + i] = std::cos(incl);
    }
}

// cap_bin / cap_bin_rt on the host: direction grid N, L level bins
static void bin_of(float x, float y, float z, float c, int N, int L, int &iu, int &iv, int &l) {
    const float s = 1.0f / (std::fabs(x) + std::fabs(y) + std::fabs(z));
    float u = x * s, v = y * s;
    if (z < 0.0f) {
        const float uu = std::copysign(1.0f - std::fabs(v), u);
        v = std::copysign(1.0f - std::fabs(u), v);
        u = uu;
    }
    const float half = 0.5f * (float)N, scale = 0.5f * (float)N * (1.0f - 1.0f / 65536.0f);
    iu = (int)std::floor(std::fma(u, scale, half));
    iv = (int)std::floor(std::fma(v, scale, half));
    const float lf = std::floor(std::fma(c, 0.5f * (float)L, 0.5f * (float)L + 1.0f));
    l = (int)std::min(std::max(lf, 0.0f), (float)(L + 1));
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 100;
    const long samples = argc > 2 ? atol(argv[2]) : 1000000;
    std::vector<float> h;
    points(n, h);
    std::mt19937_64 rng(20261018 + n);
    std::normal_distribution<double> g(0.0, 1.0);
    std::uniform_real_distribution<double> uc(-1.25, 1.25);
    long bad_inner = 0, bad_cover = 0, ring_bits = 0, inner_bits = 0;
    if (n <= 128) {
        std::vector<uint32_t> t(kCapTableBins * 8);
        cap_build_table(n, h.data(), h.data() + n, h.data() + 2 * (size_t)n, t.data());
        for (long s = 0; s < samples; ++s) {
            double d[3] = {g(rng), g(rng), g(rng)};
            // every tenth sample sits on an axis plane or an octant boundary, where the octahedral map folds
            if (s % 10 == 0) d[s / 10 % 3] = 0.0;
            if (s % 10 == 1) d[1] = d[0];
            const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (len < 1e-3) continue;
            const double c = s % 7 == 0 ? -1.0 + 2.0 * (double)(rng() % (kCapL + 1)) / kCapL : uc(rng);   // also exactly on level boundaries
            int iu, iv, l;
            bin_of((float)(d[0] / len), (float)(d[1] / len), (float)(d[2] / len), (float)c, kCapN, kCapL, iu, iv, l);
            if (iu < 0 || iu >= kCapN || iv < 0 || iv >= kCapN) { printf("bin out of range\n"); return 2; }
            const uint32_t *e = t.data() + (((size_t)l * kCapN + iv) * kCapN + iu) * 8;
            // the device computes with the float direction and level: compare against those
            const double fx = (float)(d[0] / len), fy = (float)(d[1] / len), fz = (float)(d[2] / len), fc = (float)c;
            const double flen = std::sqrt(fx * fx + fy * fy + fz * fz);
            for (uint32_t p = 0; p < n; ++p) {
                const bool in = (e[p >> 5] >> (p & 31)) & 1u, rg = (e[4 + (p >> 5)] >> (p & 31)) & 1u;
                const double dot = ((double)h[p] * fx + (double)h[n + p] * fy + (double)h[2 * (size_t)n + p] * fz) / flen;
                const bool occ = dot < fc;
                bad_inner += in && !occ;
                bad_cover += occ && !(in || rg);
                ring_bits += rg; inner_bits += in;
            }
        }
    } else {
        const CapDims D = cap_dims(64, kCapmL, 8);
        const size_t words = cap_multi_words(D), stride = (size_t)4 << D.nchp_shift;
        std::vector<uint32_t> tin(words), trg(words);
        cap_build_table_multi(n, h.data(), h.data() + n, h.data() + 2 * (size_t)n, D, tin.data(), trg.data());
        for (long s = 0; s < samples; ++s) {
            double d[3] = {g(rng), g(rng), g(rng)};
            if (s % 10 == 0) d[s / 10 % 3] = 0.0;
            const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (len < 1e-3) continue;
            const double c = s % 7 == 0 ? -1.0 + 2.0 * (double)(rng() % (kCapmL + 1)) / kCapmL : uc(rng);
            int iu, iv, l;
            bin_of((float)(d[0] / len), (float)(d[1] / len), (float)(d[2] / len), (float)c, D.n, kCapmL, iu, iv, l);
            const size_t b = ((size_t)l * D.n + iv) * D.n + iu;
            const double fx = (float)(d[0] / len), fy = (float)(d[1] / len), fz = (float)(d[2] / len), fc = (float)c;
            const double flen = std::sqrt(fx * fx + fy * fy + fz * fz);
            for (uint32_t p = 0; p < n; ++p) {
                const bool in = (tin[b * stride + (p >> 5)] >> (p & 31)) & 1u, rg = (trg[b * stride + (p >> 5)] >> (p & 31)) & 1u;
                const double dot = ((double)h[p] * fx + (double)h[n + p] * fy + (double)h[2 * (size_t)n + p] * fz) / flen;
                const bool occ = dot < fc;
                bad_inner += in && !occ;
                bad_cover += occ && !(in || rg);
                ring_bits += rg; inner_bits += in;
            }
        }
    }
    printf("n=%u samples=%ld inner_bits=%ld ring_bits=%ld wrong_inner=%ld uncovered=%ld\n", n, samples, inner_bits, ring_bits, bad_inner, bad_cover);
    return bad_inner || bad_cover ? 1 : 0;
}
