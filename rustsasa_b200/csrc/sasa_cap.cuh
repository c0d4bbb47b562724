// sasa_cap.cuh -- cap-table occlusion: the exact Shrake-Rupley point test with (almost) no point tests.
//
// Neighbour j occludes a spherical cap of atom i's test sphere: the reference's test dot(p, v) < limit
// (src/lib.rs:128-147; v = c_i - c_j) is p . v^ < c with c = limit / |v|.  For a fixed point set the answer depends
// only on the direction v^ and the level c, so it is tabulated: v^ is binned on an octahedral N x N grid, c on L
// uniform levels of [-1, 1] (plus one level below and one above), and every bin holds two point masks
//     inner : points occluded for EVERY (v^, c) of the bin  -> decided without arithmetic
//     ring  : points occluded for SOME but not all          -> decided by the reference's exact test
// and points in neither are exposed to this neighbour for every (v^, c) of the bin.  Per atom, one lane per neighbour
// looks its bin up (32 B from L2), the inner masks are OR-reduced across the warp, and only (ring point, neighbour)
// pairs whose point is still uncovered run the exact arithmetic of sasa_device.cuh -- about 10-25 tests per atom
// instead of ~1,500 (tools/cap_model.py).  The result is the reference's, bit for bit: the table never decides a
// point the exact test could decide differently, because the bins are built in double precision with a margin
// (kCapEpsAng on the direction, kCapEpsC on the level) that is 100x the worst float rounding of the binning
// arithmetic below and of the reference's own dot product, and 100x smaller than a bin.
#pragma once
#include <cmath>
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "sasa_device.cuh"

#ifndef SASA_CAP_N
#define SASA_CAP_N 128         // direction bins per axis of the octahedral square (even)
#endif
#ifndef SASA_CAP_L
#define SASA_CAP_L 128         // level bins over c in [-1, 1].  Thinner rings mean fewer exact tests: 64 -> 96 -> 128 levels gave
                               // 1,916 -> 1,938 -> 1,946 M atoms/s (477.5 -> 470.6 -> 466.5 instructions per atom) for a 34.6 -> 51 -> 68 MB
                               // table that still sits in the 126 MB L2 (gpurun_out r04f / r04g); 96 x 96 directions x 128 / 192 levels
                               // (38 / 57 MB) measured slower than 128 x 128 x 128
#endif
#ifndef SASA_CAP_RING1
#define SASA_CAP_RING1 1       // ring tests: one loop over the lane's whole 128-bit mask (0: one loop per 32-point word)
#endif
#ifndef SASA_CAP_LD256
#define SASA_CAP_LD256 1       // one 256-bit load per bin (LDG.E.256, sm_100) instead of two 128-bit loads
#endif
#ifndef SASA_CAP_R2
#define SASA_CAP_R2 1          // rounds 0 and 1 (the first 64 neighbours) are fetched together: one OR-reduction for both, the
                               // ring tests of round 0 already see the inner masks of round 1, and both fetches overlap
#endif
#ifndef SASA_CAP_LAZYDIV
#define SASA_CAP_LAZYDIV 1     // the IEEE division of the entry's limit (src/lib.rs:135-136) only in lanes that really run an exact test; the
#endif                         // table lookup takes limit ~ numerator x rcp(2r), well inside the bins' margins.  Round 1: 560.6 -> 557.3
                               // instructions per atom, +0.1 % (off).  With the 128-level table fewer warps reach a ring test at all:
                               // 448.0 -> 441.1 instructions per atom, 2,019 -> 2,029 M atoms/s (gpurun_out r04k); on
#ifndef SASA_CAP_FULLX
#define SASA_CAP_FULLX 2       // 1: an atom whose points are ALL inside inner masks after the first reduction (42 % of the atoms of a
#endif                         // protein) returns 0 there: no ring masks, no ring loops, no second reduction; 2: the same test as three
                               // logic instructions for n_points > 96 instead of four population counts (+0.6 %, r04k)
#ifndef SASA_CAP_NOBR
#define SASA_CAP_NOBR 1        // 1: cap_fetch computes the bin in every lane and selects the special bins instead of branching around
#endif                         // the arithmetic: 477.5 -> 464.3 warp instructions per atom, 1,916 -> 1,958 M atoms/s (gpurun_out r04g)
// Tried and dropped (gpurun_out r04j): the entry's IEEE division as an explicit three-FFMA sequence with the reciprocal of 2r
// refined once per atom and no FCHK range check (bit-identical to __fdiv_rn on a 2.6 M-operand self-test over 0.02 <= 2r <= 4096):
// 449.1 -> 448.0 warp instructions per atom, no change in time -- __fdiv_rn's fast path already is that sequence plus three
// instructions, and the range guard costs as many.
#ifndef SASA_CAP_TEX
#define SASA_CAP_TEX 1         // 1: the fused kernel reads the table through the TEXTURE path (two tex1Dfetch<uint4> per bin) instead of one
#endif                         // LDG.E.256.  Once the instruction count was down to 441 per atom the LSU data pipe of L1TEX had become the
                               // top unit (81 % busy: 74 shared-memory wavefronts + 43 for the table -- one per lane, every bin in a line
                               // of its own -- of the 143 cycles an SM has per atom, profiles/r05a_cfg2_summary.txt); the texture pipe
                               // sat idle.  tools/tex_bench.cu measured the two paths alone and against shared-memory traffic (random
                               // 32-byte bins of a 68 MB table, 32 warps per SM: LDG 0.86 -> 0.59 bins per cycle per SM when scattered
                               // LDS.128 compete, textures 0.66 -> 0.69).  Fused kernel: 2,032 -> 2,104 M atoms/s, issue slots 77.6 -> 81.8 %
                               // busy (gpurun_out r05b); 2: only the first round through textures (2,099); 3: fetch predicated on the
                               // lane having a neighbour (2,068)
#ifndef SASA_CAP_PF
#define SASA_CAP_PF 0          // fetch the next round's masks before the current round's ring tests
#endif

namespace sasa {

constexpr int kCapN = SASA_CAP_N, kCapL = SASA_CAP_L;
constexpr int kCapmL = 64;                                   // level bins of the chunked tables (128 < n <= 1024): 2 x 34.6 MB at 64 x 64 directions
constexpr int kCapLevels = kCapL + 2;                       // level 0: c < -1, level L + 1: c >= 1
constexpr size_t kCapBins = (size_t)kCapLevels * kCapN * kCapN;
constexpr size_t kCapBinDegenerate = kCapBins;               // extra bin: no inner points, every point in the ring
constexpr size_t kCapBinEmpty = kCapBins + 1;                // extra bin: nothing (lanes without a neighbour)
constexpr size_t kCapTableBins = kCapBins + 2;
constexpr double kCapEpsAng = 1.0e-3;                        // radians added to every direction bin's radius
constexpr double kCapEpsC = 1.0e-4;                          // widening of every level interval
constexpr float kCapMinV2 = 1.0e-6f;                         // |v|^2 below this: no table, every point takes the exact test

// ---- host: the table of one point set (n <= 128), 2 x uint4 per bin: {inner, ring} -------------------------------------
// Direction bin (iu, iv) is the cell [-1 + 2 iu/N, -1 + 2 (iu+1)/N] x [...] of the octahedral square.  Its cell
// edges are great-circle arcs (u = const is a plane through the origin inside each octant; with N even the octant
// boundaries and the fold |u| + |v| = 1 run through grid nodes), so the largest angle between the cell centre and any
// direction of the cell is attained at one of the four corners: rho = max corner angle + kCapEpsAng.  For a point at
// angle alpha from the centre, p . u over the cell lies within [cos(min(alpha + rho, pi)), cos(max(alpha - rho, 0))].
inline void cap_oct_dir(double u, double v, double d[3]) {
    const double z = 1.0 - std::fabs(u) - std::fabs(v);
    double x = u, y = v;
    if (z < 0.0) {
        x = (1.0 - std::fabs(v)) * (u >= 0.0 ? 1.0 : -1.0);
        y = (1.0 - std::fabs(u)) * (v >= 0.0 ? 1.0 : -1.0);
    }
    const double inv = 1.0 / std::sqrt(x * x + y * y + z * z);
    d[0] = x * inv; d[1] = y * inv; d[2] = z * inv;
}

inline void cap_build_table(uint32_t n, const float *px, const float *py, const float *pz, uint32_t *tab /* 8 * kCapTableBins */) {
    const int N = kCapN, L = kCapL;
    memset(tab, 0, kCapTableBins * 8 * sizeof(uint32_t));
    for (uint32_t p = 0; p < n; ++p) tab[kCapBinDegenerate * 8 + 4 + (p >> 5)] |= 1u << (p & 31);
    // level l covers c in [lo(l), hi(l)): lo(0) = -inf, lo(l) = -1 + 2 (l - 1) / L, hi(l) = lo(l + 1), hi(L + 1) = +inf
    std::vector<double> lo(kCapLevels), hi(kCapLevels);
    for (int l = 0; l < kCapLevels; ++l) {
        lo[l] = l == 0 ? -INFINITY : -1.0 + 2.0 * (l - 1) / L;
        hi[l] = l == kCapLevels - 1 ? INFINITY : -1.0 + 2.0 * l / L;
    }
    // Rows of the direction grid are independent: spread over the host threads.  One row of direction bins at a time: per (bin,
    // point) the two level thresholds, then the masks of all levels of the row in a local block laid out [level][iu] (532 KB at
    // 130 levels x 128 bins), copied out level by level as contiguous 4 KB runs.  (Testing both predicates at every level and
    // setting single bits across the table's N * N * 32-byte level stride took 0.63 s on 8 threads for 128 levels; this form
    // 0.15 s, byte-identical.)
    auto rows = [&](int iv0, int iv1) {
    std::vector<uint32_t> loc((size_t)kCapLevels * N * 8);
    for (int iv = iv0; iv < iv1; ++iv) {
        std::fill(loc.begin(), loc.end(), 0u);
        for (int iu = 0; iu < N; ++iu) {
            const double u0 = -1.0 + 2.0 * iu / N, u1 = -1.0 + 2.0 * (iu + 1) / N;
            const double v0 = -1.0 + 2.0 * iv / N, v1 = -1.0 + 2.0 * (iv + 1) / N;
            double c[3], k[3];
            cap_oct_dir(0.5 * (u0 + u1), 0.5 * (v0 + v1), c);
            double rho = 0.0;
            const double cu[4] = {u0, u1, u0, u1}, cv[4] = {v0, v0, v1, v1};
            for (int t = 0; t < 4; ++t) {
                cap_oct_dir(cu[t], cv[t], k);
                rho = std::max(rho, std::acos(std::min(1.0, std::max(-1.0, c[0] * k[0] + c[1] * k[1] + c[2] * k[2]))));
            }
            rho += kCapEpsAng;
            // both predicates are monotone in the level (lo and hi increase): point p is in the ring of levels [lout, lin) and in
            // the inner mask of levels [lin, levels)
            uint8_t lin[128], lout[128];
            static_assert(kCapLevels < 256, "levels must fit a byte");
            for (uint32_t p = 0; p < n; ++p) {
                // normalise the float point (it is unit length only to float rounding)
                const double x = px[p], y = py[p], z = pz[p], len = std::sqrt(x * x + y * y + z * z);
                const double alpha = std::acos(std::min(1.0, std::max(-1.0, (c[0] * x + c[1] * y + c[2] * z) / len)));
                // the device compares the UNnormalised point: p . v^ = len * cos(angle)
                const double dmax = len * std::cos(std::max(alpha - rho, 0.0)), dmin = len * std::cos(std::min(alpha + rho, M_PI));
                int a = 0, b = 0;
                while (a < kCapLevels && !(dmin < hi[a] + kCapEpsC)) ++a;      // first level where the point is occluded for some (v^, c)
                b = a;
                while (b < kCapLevels && !(dmax < lo[b] - kCapEpsC)) ++b;      // first level where it is occluded whatever (v^, c)
                lout[p] = (uint8_t)a;
                lin[p] = (uint8_t)b;
            }
            for (int l = 0; l < kCapLevels; ++l) {
                uint32_t *e = loc.data() + ((size_t)l * N + iu) * 8;
                for (uint32_t p = 0; p < n; ++p) {
                    const uint32_t in = l >= (int)lin[p], rg = (l >= (int)lout[p]) & !in;
                    e[p >> 5] |= in << (p & 31);
                    e[4 + (p >> 5)] |= rg << (p & 31);
                }
            }
        }
        for (int l = 0; l < kCapLevels; ++l)
            memcpy(tab + (((size_t)l * N + iv) * N) * 8, loc.data() + (size_t)l * N * 8, (size_t)N * 32);
    }
    };
    const int nt = std::max(1, std::min<int>((int)std::thread::hardware_concurrency(), 32));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(rows, N * t / nt, N * (t + 1) / nt);
    rows(0, N / nt);
    for (auto &th : pool) th.join();
}

// ---- host: chunked tables for 128 < n <= 1024 points -------------------------------------------------------------------
// The point set is cut into chunks of 128 consecutive points (the golden spiral orders points by latitude, so a chunk is
// a latitude band); bin b holds, for every chunk c < nchp (the power of two >= the number of chunks), one 128-bit inner
// mask at inner[b * nchp + c] and one ring mask at ring[b * nchp + c].  Inner and ring live in separate arrays because the
// kernel's first pass reads inner masks only.  Same bins, margins and semantics as cap_build_table; grid size N and level
// count L are run-time values here (a 64 x 64 x 66 table of 8 chunks is 2 x 34.6 MB and stays in the 126 MB L2).
inline CapDims cap_dims(int N, int L, int nchunks) {
    CapDims D;
    D.half = 0.5f * (float)N;
    D.scale = 0.5f * (float)N * (1.0f - 1.0f / 65536.0f);
    D.lhalf = 0.5f * (float)L;
    D.n = N;
    D.levels = L + 2;
    D.nchp_shift = 0;
    while ((1 << D.nchp_shift) < nchunks) ++D.nchp_shift;
    D.bin_degenerate = (unsigned)((size_t)D.levels * N * N);
    D.bin_empty = D.bin_degenerate + 1;
    return D;
}
inline size_t cap_multi_words(const CapDims &D) { return ((size_t)D.levels * D.n * D.n + 2) * ((size_t)4 << D.nchp_shift); }

inline void cap_build_table_multi(uint32_t n, const float *px, const float *py, const float *pz, const CapDims &D,
                                  uint32_t *inner, uint32_t *ring /* cap_multi_words(D) each */) {
    const int N = D.n, L = D.levels - 2, levels = D.levels;
    const size_t stride = (size_t)4 << D.nchp_shift;   // u32 words per bin
    memset(inner, 0, cap_multi_words(D) * sizeof(uint32_t));
    memset(ring, 0, cap_multi_words(D) * sizeof(uint32_t));
    for (uint32_t p = 0; p < n; ++p) ring[(size_t)D.bin_degenerate * stride + (p >> 5)] |= 1u << (p & 31);
    std::vector<double> lo(levels), hi(levels);
    for (int l = 0; l < levels; ++l) {
        lo[l] = l == 0 ? -INFINITY : -1.0 + 2.0 * (l - 1) / L;
        hi[l] = l == levels - 1 ? INFINITY : -1.0 + 2.0 * l / L;
    }
    auto rows = [&](int iv0, int iv1) {
        for (int iv = iv0; iv < iv1; ++iv)
            for (int iu = 0; iu < N; ++iu) {
                const double u0 = -1.0 + 2.0 * iu / N, u1 = -1.0 + 2.0 * (iu + 1) / N;
                const double v0 = -1.0 + 2.0 * iv / N, v1 = -1.0 + 2.0 * (iv + 1) / N;
                double c[3], k[3];
                cap_oct_dir(0.5 * (u0 + u1), 0.5 * (v0 + v1), c);
                double rho = 0.0;
                const double cu[4] = {u0, u1, u0, u1}, cv[4] = {v0, v0, v1, v1};
                for (int t = 0; t < 4; ++t) {
                    cap_oct_dir(cu[t], cv[t], k);
                    rho = std::max(rho, std::acos(std::min(1.0, std::max(-1.0, c[0] * k[0] + c[1] * k[1] + c[2] * k[2]))));
                }
                rho += kCapEpsAng;
                for (uint32_t p = 0; p < n; ++p) {
                    const double x = px[p], y = py[p], z = pz[p], len = std::sqrt(x * x + y * y + z * z);
                    const double alpha = std::acos(std::min(1.0, std::max(-1.0, (c[0] * x + c[1] * y + c[2] * z) / len)));
                    const double dmax = len * std::cos(std::max(alpha - rho, 0.0)), dmin = len * std::cos(std::min(alpha + rho, M_PI));
                    const uint32_t bit = 1u << (p & 31);
                    const size_t word = p >> 5;   // chunk (p >> 7) * 4 + word within the chunk: consecutive
                    // both predicates are monotone in the level: ring for levels [a, b), inner for [b, levels), nothing below a
                    int a = 0;
                    while (a < levels && !(dmin < hi[a] + kCapEpsC)) ++a;
                    int b = a;
                    while (b < levels && !(dmax < lo[b] - kCapEpsC)) ++b;
                    for (int l = a; l < b; ++l) ring[(((size_t)l * N + iv) * N + iu) * stride + word] |= bit;
                    for (int l = b; l < levels; ++l) inner[(((size_t)l * N + iv) * N + iu) * stride + word] |= bit;
                }
            }
    };
    const int nt = std::max(1, std::min<int>((int)std::thread::hardware_concurrency(), 32));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(rows, N * t / nt, N * (t + 1) / nt);
    rows(0, N / nt);
    for (auto &th : pool) th.join();
}

// ---- device ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cap_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cap_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Bin of entry e = (vx, vy, vz, limit) with vmag = |v|^2 >= kCapMinV2.  Approximate reciprocals are fine here: the bins'
// margins absorb 1e-4 in c and 1e-3 rad in direction, these errors are ~1e-6.
__device__ __forceinline__ int cap_bin(const float4 e, float vmag) {
    const float c = e.w * cap_rsqrt(vmag);
    const float s = cap_rcp(fabsf(e.x) + fabsf(e.y) + fabsf(e.z));
    float u = e.x * s, v = e.y * s;
    if (e.z < 0.0f) {
        const float uu = copysignf(1.0f - fabsf(v), u);
        v = copysignf(1.0f - fabsf(u), v);
        u = uu;
    }
    // |u|, |v| <= 1 + 1e-6: the scale just below N/2 keeps floor() inside [0, N - 1] without clamps (a direction that
    // lands one ulp into the neighbouring bin is covered by that bin's angular margin)
    constexpr float kHalf = 0.5f * kCapN, kScale = 0.5f * kCapN * (1.0f - 1.0f / 65536.0f);
    const int iu = __float2int_rd(fmaf(u, kScale, kHalf));
    const int iv = __float2int_rd(fmaf(v, kScale, kHalf));
    // c is unbounded (and __float2int_rd saturates): level 0 below -1, level L + 1 from +1 on
    const int l = min(max(__float2int_rd(fmaf(c, 0.5f * kCapL, 0.5f * kCapL + 1.0f)), 0), kCapLevels - 1);
    return (l * kCapN + iv) * kCapN + iu;
}

// The reference's test of point `pt` against this lane's neighbour e (tail points: unfused dot, <=).
__device__ __forceinline__ bool cap_exact(const float4 *pts, int pt, const float4 e, int nbody) {
    const float4 P = pts[pt];
    return pt >= nbody ? dot_tail(P.x, P.y, P.z, e) <= e.w : dot_body(P.x, P.y, P.z, e) < e.w;
}

// Exact tests of this lane's neighbour e against its ring points (m0..m3: points 0-31, .., 96-127 still uncovered);
// points found occluded are OR-ed into a0..a3.
__device__ __forceinline__ void cap_ring_tests(unsigned m0, unsigned m1, unsigned m2, unsigned m3, float4 e, float two_r,
                                               const float4 *pts, int nbody, unsigned &a0, unsigned &a1, unsigned &a2,
                                               unsigned &a3) {
#if SASA_CAP_LAZYDIV
    // e.w still holds the numerator (t_j - |v|^2) - r^2: the reference's limit is its IEEE quotient by 2r
    if ((m0 | m1 | m2 | m3) != 0u) e.w = __fdiv_rn(e.w, two_r);
#else
    (void)two_r;
#endif
#if SASA_CAP_RING1
    // every trip takes the highest point left in the lane's 128-bit mask; the warp runs max-over-lanes trips
    while (__any_sync(kFull, (m0 | m1 | m2 | m3) != 0u)) {
        if ((m0 | m1 | m2 | m3) != 0u) {
            const int w = m3 ? 3 : m2 ? 2 : m1 ? 1 : 0;
            const unsigned m = m3 ? m3 : m2 ? m2 : m1 ? m1 : m0;
            const int b = 31 - __clz(m);
            const unsigned bit = 1u << b;
            const unsigned hit = cap_exact(pts, 32 * w + b, e, nbody) ? bit : 0u;
            if (w == 3) { m3 ^= bit; a3 |= hit; }
            else if (w == 2) { m2 ^= bit; a2 |= hit; }
            else if (w == 1) { m1 ^= bit; a1 |= hit; }
            else { m0 ^= bit; a0 |= hit; }
        }
    }
#else
    unsigned *const mm[4] = {&m0, &m1, &m2, &m3};
    unsigned *const aa[4] = {&a0, &a1, &a2, &a3};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        unsigned m = *mm[w], hit = 0u;
        while (__any_sync(kFull, m != 0u)) {
            if (m) {
                const int b = 31 - __clz(m);
                const unsigned bit = 1u << b;
                m ^= bit;
                if (cap_exact(pts, 32 * w + b, e, nbody)) hit |= bit;
            }
        }
        *aa[w] |= hit;
    }
#endif
}

struct CapRound {
    float4 e;       // this lane's neighbour: (vx, vy, vz, limit), the reference's arithmetic (SASA_CAP_LAZYDIV: .w = the limit's
                    // numerator; cap_ring_tests divides where a test is due)
    uint4 in, rg;   // its bin's inner and ring masks
};

// make_entry without the division: (vx, vy, vz, (t_j - |v|^2) - r^2), same operations in the same order (src/lib.rs:128-136).
__device__ __forceinline__ float4 make_entry_num(const float4 ai, const float4 aj, float probe, float r2, float *vmag_out) {
    const float vx = __fsub_rn(ai.x, aj.x), vy = __fsub_rn(ai.y, aj.y), vz = __fsub_rn(ai.z, aj.z);
    const float vmag = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
    const float tj = __fadd_rn(aj.w, probe);
    const float t = __fmul_rn(tj, tj);
    *vmag_out = vmag;
    return make_float4(vx, vy, vz, __fsub_rn(__fsub_rn(t, vmag), r2));
}

// Lane `lane` of the round starting at q0: entry + table fetch (lanes past k fetch the empty bin).
// Atoms: accessor of the cell-sorted atoms (shared-memory array in the fused kernel, global array in the large-structure path);
// IdxT: type of the candidate positions.
template <bool TEX = false, class Atoms, class IdxT>
__device__ __forceinline__ CapRound cap_fetch(const uint4 *__restrict__ tab, const Atoms &s_atom, const float4 ai, float probe,
                                              float r2, float two_r, float inv_two_r, const IdxT *cand, int q, int k,
                                              unsigned long long tex = 0ull) {
    CapRound R;
    const bool valid = q < k;
    const float4 aj = s_atom(valid ? (int)cand[q] : 0);
    float vmag;
#if SASA_CAP_LAZYDIV
    R.e = make_entry_num(ai, aj, probe, r2, &vmag);
    float4 eb = R.e;
    eb.w = R.e.w * inv_two_r;   // approximate limit: good for the table lookup only
#else
    R.e = make_entry(ai, aj, probe, r2, two_r, &vmag);
    const float4 eb = R.e;
    (void)inv_two_r;
#endif
    // (nearly) coincident centres: no direction -- the degenerate bin sends every point to the exact test
#if SASA_CAP_NOBR
    // branch-free: the bin arithmetic runs in every lane (it cannot trap: rsqrt(0) = inf, float -> int conversions saturate) and
    // two selects pick the special bins
    int bin = cap_bin(eb, vmag);
    bin = vmag >= kCapMinV2 ? bin : (int)kCapBinDegenerate;
    bin = valid ? bin : (int)kCapBinEmpty;
#else
    const int bin = valid ? (vmag >= kCapMinV2 ? cap_bin(eb, vmag) : (int)kCapBinDegenerate) : (int)kCapBinEmpty;
#endif
    if (TEX && (SASA_CAP_TEX != 2 || q < 32)) {   // SASA_CAP_TEX 2: the first round through the texture pipe, later rounds through LSU
#if SASA_CAP_TEX == 3
        R.in = make_uint4(0u, 0u, 0u, 0u);
        R.rg = R.in;
        if (valid) {
            R.in = tex1Dfetch<uint4>((cudaTextureObject_t)tex, 2 * bin);
            R.rg = tex1Dfetch<uint4>((cudaTextureObject_t)tex, 2 * bin + 1);
        }
#else
        R.in = tex1Dfetch<uint4>((cudaTextureObject_t)tex, 2 * bin);
        R.rg = tex1Dfetch<uint4>((cudaTextureObject_t)tex, 2 * bin + 1);
#endif
        return R;
    }
    const uint4 *b = tab + 2 * (size_t)(unsigned)bin;
#if SASA_CAP_LD256
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(R.in.x), "=r"(R.in.y), "=r"(R.in.z), "=r"(R.in.w), "=r"(R.rg.x), "=r"(R.rg.y), "=r"(R.rg.z), "=r"(R.rg.w)
        : "l"(b));
#else
    R.in = __ldg(b);
    R.rg = __ldg(b + 1);
#endif
    return R;
}

// One round of an atom, given this lane's entry and masks: OR of the inner masks, exact tests of the ring points left.
__device__ __forceinline__ void cap_round(const CapRound &R, float two_r, const float4 *pts, int nbody, unsigned &a0, unsigned &a1,
                                          unsigned &a2, unsigned &a3) {
    a0 |= R.in.x; a1 |= R.in.y; a2 |= R.in.z; a3 |= R.in.w;
    const unsigned c0 = __reduce_or_sync(kFull, a0), c1 = __reduce_or_sync(kFull, a1),
                   c2 = __reduce_or_sync(kFull, a2), c3 = __reduce_or_sync(kFull, a3);
    cap_ring_tests(R.rg.x & ~c0, R.rg.y & ~c1, R.rg.z & ~c2, R.rg.w & ~c3, R.e, two_r, pts, nbody, a0, a1, a2, a3);
}

__device__ __forceinline__ int cap_covered(unsigned a0, unsigned a1, unsigned a2, unsigned a3) {
    return __popc(__reduce_or_sync(kFull, a0)) + __popc(__reduce_or_sync(kFull, a1)) +
           __popc(__reduce_or_sync(kFull, a2)) + __popc(__reduce_or_sync(kFull, a3));
}

// One atom: neighbours cand[0, k) (positions in the cell-sorted shared atom array).  Returns the exposed-point count.
// Lane q of round r owns neighbour 32 r + q: it builds the entry with the reference's arithmetic (make_entry), fetches
// its bin's masks, and -- after the warp-wide OR of the inner masks -- runs the exact test on its own ring points that
// are still uncovered.  Hits are folded into the next OR.
template <bool TEX = false, class Atoms, class IdxT>
__device__ __forceinline__ int cap_atom(const uint4 *__restrict__ tab, const Atoms &s_atom, const float4 ai, float probe,
                                        const IdxT *cand, int k, const float4 *pts, int n_points, int nbody, unsigned long long tex = 0ull) {
    const int lane = lane_id();
    const float r = __fadd_rn(ai.w, probe);
    const float r2 = __fmul_rn(r, r), two_r = __fmul_rn(2.0f, r);
    const float inv_two_r = cap_rcp(two_r);
    unsigned a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;      // this lane's inner masks and exact hits, not yet reduced
#if SASA_CAP_R2
    {
        const CapRound A = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, lane, k, tex);
        CapRound B;
        B.e = make_float4(0.f, 0.f, 0.f, 0.f);
        B.in = make_uint4(0u, 0u, 0u, 0u);
        B.rg = B.in;
        if (k > 32) B = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, 32 + lane, k, tex);
        a0 = A.in.x | B.in.x; a1 = A.in.y | B.in.y; a2 = A.in.z | B.in.z; a3 = A.in.w | B.in.w;
        const unsigned c0 = __reduce_or_sync(kFull, a0), c1 = __reduce_or_sync(kFull, a1),
                       c2 = __reduce_or_sync(kFull, a2), c3 = __reduce_or_sync(kFull, a3);
#if SASA_CAP_FULLX
        // the masks hold bits of existing points only, so full coverage is a population count
#if SASA_CAP_FULLX == 2
        // more than 96 points: words 0-2 must be full and word 3 must hold its n - 96 low bits (three instructions instead of seven)
        if (n_points > 96 ? ((c0 & c1 & c2) == 0xffffffffu && c3 == (0xffffffffu >> (128 - n_points)))
                          : (__popc(c0) + __popc(c1) + __popc(c2) + __popc(c3) == n_points))
            return 0;
#else
        if (__popc(c0) + __popc(c1) + __popc(c2) + __popc(c3) == n_points) return 0;
#endif
#endif
        cap_ring_tests(A.rg.x & ~c0, A.rg.y & ~c1, A.rg.z & ~c2, A.rg.w & ~c3, A.e, two_r, pts, nbody, a0, a1, a2, a3);
        if (k > 32) cap_ring_tests(B.rg.x & ~c0, B.rg.y & ~c1, B.rg.z & ~c2, B.rg.w & ~c3, B.e, two_r, pts, nbody, a0, a1, a2, a3);
    }
#pragma unroll 1
    for (int q0 = 64; q0 < k; q0 += 32) {
        const CapRound R = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, q0 + lane, k, tex);
        cap_round(R, two_r, pts, nbody, a0, a1, a2, a3);
    }
#elif SASA_CAP_PF
    CapRound R = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, lane, k, tex);
#pragma unroll 1
    for (int q0 = 0; q0 < k; q0 += 32) {
        CapRound Nx = R;
        if (q0 + 32 < k) Nx = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, q0 + 32 + lane, k, tex);
        cap_round(R, two_r, pts, nbody, a0, a1, a2, a3);
        R = Nx;
    }
#else
#pragma unroll 1
    for (int q0 = 0; q0 < k; q0 += 32) {
        const CapRound R = cap_fetch<TEX>(tab, s_atom, ai, probe, r2, two_r, inv_two_r, cand, q0 + lane, k, tex);
        cap_round(R, two_r, pts, nbody, a0, a1, a2, a3);
    }
#endif
    return n_points - cap_covered(a0, a1, a2, a3);
}

// ---- chunked cap tables: 128 < n_points <= 1024 --------------------------------------------------------------------------
// Tried and dropped (gpurun_out r05i): prefetch.global.L1 of each neighbour's inner-mask line as soon as its bin is known
// (cfg5 1.618 -> 1.627 ms).
#ifndef SASA_CAPM_UNROLL
#define SASA_CAPM_UNROLL 2     // pass-1 loop unroll.  Measured on cfg5 (gpurun_out r02l): 2 -> 1.60 ms, 4 -> 1.93 ms (spills), 8 -> 1.62 ms;
#endif                         // a 96^2 / 128^2 direction grid gives 1.55 / 1.54 ms for 2.3x / 4x the table (kept at 64^2)
// Bin of entry e with run-time grid dimensions (same arithmetic as cap_bin).
__device__ __forceinline__ unsigned cap_bin_rt(const float4 e, float vmag, const CapDims &D) {
    const float c = e.w * cap_rsqrt(vmag);
    const float s = cap_rcp(fabsf(e.x) + fabsf(e.y) + fabsf(e.z));
    float u = e.x * s, v = e.y * s;
    if (e.z < 0.0f) {
        const float uu = copysignf(1.0f - fabsf(v), u);
        v = copysignf(1.0f - fabsf(u), v);
        u = uu;
    }
    const int iu = __float2int_rd(fmaf(u, D.scale, D.half));
    const int iv = __float2int_rd(fmaf(v, D.scale, D.half));
    const int l = min(max(__float2int_rd(fmaf(c, D.lhalf, D.lhalf + 1.0f)), 0), D.levels - 1);
    return (unsigned)((l * D.n + iv) * D.n + iu);
}

// One atom against the chunked table.  NCHP (2, 4 or 8) lanes share a neighbour, one lane per 128-point chunk: lane l
// owns chunk l mod NCHP of neighbour (l div NCHP) of every round of 32 / NCHP neighbours, so the 128-bit coverage of a
// chunk stays in four registers of the lanes that own it and a neighbour's NCHP masks are one contiguous read.
//   bins     lane q of a round of 32: entry (reference arithmetic) -> bin, packed with the candidate index into nbp[q]
//   pass 1   OR of the inner masks over all neighbours; a buried atom (every point inside some inner cap -- the common
//            case: 94 % of all points are) ends here with no point test at all
//   pass 2   per (neighbour, chunk): ring points not yet covered take the reference's exact test; hits join the coverage
// Returns the exposed-point count.  nb: candidate indices (into `atoms`) of the k <= 128 neighbours; nbp: 128 u32 of scratch.
// IDX_BITS: width of a candidate index in the packed (bin, index) word: 9 for the staged strips of the large-structure path,
// 13 for the shared-memory atoms of the fused kernel (bins need 19 bits at 64 x 64 x 66).
template <int NCHP, int IDX_BITS, class Atoms, class IdxT>
__device__ __forceinline__ int capm_atom(const uint4 *__restrict__ tin, const uint4 *__restrict__ trg, const CapDims &D,
                                         const Atoms &atoms, const float4 ai, float probe, const IdxT *nb, int k, uint32_t *nbp,
                                         const float4 *__restrict__ pts4, int n_points, int nbody) {
    static_assert(NCHP == 2 || NCHP == 4 || NCHP == 8, "chunk lanes");
    constexpr int PER = 32 / NCHP;
    const int lane = lane_id();
    const int ch = lane & (NCHP - 1), sub = lane / NCHP;
    const float r = __fadd_rn(ai.w, probe);
    const float r2 = __fmul_rn(r, r), two_r = __fmul_rn(2.0f, r);
#pragma unroll 1
    for (int q0 = 0; q0 < k; q0 += 32) {
        const int q = q0 + lane;
        if (q < k) {
            const unsigned j = (unsigned)nb[q];
            float vmag;
            const float4 e = make_entry(ai, atoms((int)j), probe, r2, two_r, &vmag);
            const unsigned bin = vmag >= kCapMinV2 ? cap_bin_rt(e, vmag, D) : D.bin_degenerate;
            nbp[q] = (bin << IDX_BITS) | j;
        }
    }
    __syncwarp();
    // points of this lane's chunk that exist: [128 ch, 128 ch + 128) intersected with [0, n_points)
    unsigned vm[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const int left = n_points - (128 * ch + 32 * w);
        vm[w] = left >= 32 ? 0xffffffffu : (left > 0 ? (0xffffffffu >> (32 - left)) : 0u);
    }
    unsigned a[4] = {0u, 0u, 0u, 0u};
    constexpr int kFetchUnroll = SASA_CAPM_UNROLL;
#pragma unroll kFetchUnroll
    for (int q0 = 0; q0 < k; q0 += PER) {
        const int q = q0 + sub;
        if (q < k) {
            const uint4 m = __ldg(tin + (((size_t)(nbp[q] >> IDX_BITS)) << D.nchp_shift) + ch);
            a[0] |= m.x; a[1] |= m.y; a[2] |= m.z; a[3] |= m.w;
        }
    }
#pragma unroll
    for (int d = NCHP; d < 32; d <<= 1)
#pragma unroll
        for (int w = 0; w < 4; ++w) a[w] |= __shfl_xor_sync(kFull, a[w], d);
    const bool open = ((vm[0] & ~a[0]) | (vm[1] & ~a[1]) | (vm[2] & ~a[2]) | (vm[3] & ~a[3])) != 0u;
    if (!__any_sync(kFull, open)) return 0;
    // pass 2: exact tests of the uncovered ring points
#pragma unroll 1
    for (int q0 = 0; q0 < k; q0 += PER) {
        const int q = q0 + sub;
        unsigned m[4] = {0u, 0u, 0u, 0u};
        unsigned j = 0;
        if (q < k && open) {
            const unsigned pk = nbp[q];
            j = pk & ((1u << IDX_BITS) - 1u);
            const uint4 g = __ldg(trg + (((size_t)(pk >> IDX_BITS)) << D.nchp_shift) + ch);
            m[0] = g.x & vm[0] & ~a[0]; m[1] = g.y & vm[1] & ~a[1]; m[2] = g.z & vm[2] & ~a[2]; m[3] = g.w & vm[3] & ~a[3];
        }
        if (!__any_sync(kFull, (m[0] | m[1] | m[2] | m[3]) != 0u)) continue;
        if ((m[0] | m[1] | m[2] | m[3]) != 0u) {
            float vmag;
            const float4 e = make_entry(ai, atoms((int)j), probe, r2, two_r, &vmag);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                unsigned mw = m[w];
                while (mw) {
                    const int b = 31 - __clz(mw);
                    const unsigned bit = 1u << b;
                    mw ^= bit;
                    const int pt = 128 * ch + 32 * w + b;
                    const float4 P = __ldg(pts4 + pt);
                    const bool hit = pt >= nbody ? dot_tail(P.x, P.y, P.z, e) <= e.w : dot_body(P.x, P.y, P.z, e) < e.w;
                    if (hit) a[w] |= bit;
                }
            }
        }
    }
    // hits found by the other lanes of the same chunk
#pragma unroll
    for (int d = NCHP; d < 32; d <<= 1)
#pragma unroll
        for (int w = 0; w < 4; ++w) a[w] |= __shfl_xor_sync(kFull, a[w], d);
    int exposed = 0;
    if (sub == 0) exposed = __popc(vm[0] & ~a[0]) + __popc(vm[1] & ~a[1]) + __popc(vm[2] & ~a[2]) + __popc(vm[3] & ~a[3]);
    return __reduce_add_sync(kFull, exposed);
}

}  // namespace sasa
