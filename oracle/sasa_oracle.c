/*
 * sasa_oracle.c -- CPU restatement of RustSASA's Shrake-Rupley hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed CPU baseline.
 * The product path (rustsasa_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  The restatement reproduces, exactly, all 2,622 integer
 * exposed-point counts recoverable from the reference's own golden vector
 * tests/common/data.rs:4-238 (FIXED_LOW_RES_ATOMS, produced by the real crate;
 * tests/units.rs:17-43) and global_total = 20131.227 at 960 points
 * (tests/units.rs:117) -- see tests/test_oracle_golden.py.  The Rust crate itself
 * cannot be built here (no cargo/rustc), so there is no oracle/_ref binary.
 *
 * Every function cites the reference lines (relative to /root/reference) it
 * follows.  Arithmetic is IEEE-754 binary32 with the same association and the
 * same fused/unfused choices as the reference; build with -ffp-contract=off so
 * that the compiler never fuses what Rust does not fuse, and fmaf() is used
 * exactly where the reference calls mul_add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* src/utils/consts.rs:18-19 -- f32 constants, folded in f32 like rustc does. */
static const float GOLDEN_RATIO_F = 1.618034f;
static inline float angle_increment(void) {
    const float two_pi = 2.0f * 3.14159265358979323846f; /* 2.0 * std::f32::consts::PI */
    return two_pi * GOLDEN_RATIO_F;
}

/* ---------------------------------------------------------------------------
 * A1: generate_sphere_points, src/lib.rs:43-66.  Golden-section spiral, points
 * not centred; t = i * (1/n); libm acosf/sinf/cosf (what f32::acos/sin_cos/cos
 * lower to on linux-gnu).
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_sphere_points(size_t n_points, float *x, float *y, float *z) {
    const float inv_n = 1.0f / (float)n_points;
    const float inc = angle_increment();
    for (size_t i = 0; i < n_points; ++i) {
        const float fi = (float)i;
        const float t = fi * inv_n;
        const float inclination = acosf(1.0f - 2.0f * t);
        const float azimuth = inc * fi;
        const float sa = sinf(azimuth), ca = cosf(azimuth);
        const float si = sinf(inclination);
        x[i] = si * ca;
        y[i] = si * sa;
        z[i] = cosf(inclination);
    }
}

/* Atom as seen by the path: src/structures/atomic.rs:13-24 (position, radius,
 * id).  xyzr is the packed float4 the product uses; ids may be NULL (= all
 * distinct, i.e. id == index). */
typedef struct {
    float thr; /* threshold_squared = (r_j + probe)^2, atomic.rs:5-10 */
    uint32_t idx;
} nb_t;

typedef struct {
    nb_t *v;
    uint32_t len, cap;
} nbvec_t;

static inline void nb_push(nbvec_t *l, float thr, uint32_t idx) {
    if (l->len == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 80; /* Vec::with_capacity(80), spatial_grid.rs:213 */
        l->v = (nb_t *)realloc(l->v, (size_t)l->cap * sizeof(nb_t));
    }
    l->v[l->len].thr = thr;
    l->v[l->len].idx = idx;
    l->len++;
}

/* Rust `f32 as u32`: saturating, NaN -> 0. */
static inline uint32_t f32_as_u32(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
static inline int32_t f32_as_i32(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

typedef struct {
    uint32_t *atom_indices;
    float *px, *py, *pz, *rad;
    uint32_t *cell_starts;
    uint32_t dims[3];
    size_t num_cells;
    int32_t (*offs)[3];
    size_t n_offs;
} grid_t;

static inline size_t cell_index(const float *pos, const float *minb, float inv, const uint32_t *dims) {
    /* get_cell_index_static, spatial_grid.rs:133-143 */
    uint32_t x = f32_as_u32((pos[0] - minb[0]) * inv);
    uint32_t y = f32_as_u32((pos[1] - minb[1]) * inv);
    uint32_t z = f32_as_u32((pos[2] - minb[2]) * inv);
    return (size_t)(uint32_t)(x + y * dims[0] + z * dims[0] * dims[1]);
}

/* SpatialGrid::new, spatial_grid.rs:28-106 (+ calculate_bounds :108-130,
 * compute_half_shell_offsets :174-192). */
static int grid_build(grid_t *g, const float *xyzr, size_t n, float cell_size, float max_search_radius) {
    float minb[3] = {INFINITY, INFINITY, INFINITY}, maxb[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            minb[k] = fminf(minb[k], xyzr[4 * i + k]);
            maxb[k] = fmaxf(maxb[k], xyzr[4 * i + k]);
        }
    for (int k = 0; k < 3; ++k) {
        minb[k] -= cell_size;
        maxb[k] += cell_size;
    }
    const float inv = 1.0f / cell_size;
    for (int k = 0; k < 3; ++k) g->dims[k] = f32_as_u32(ceilf((maxb[k] - minb[k]) * inv)) + 1u;
    g->num_cells = (size_t)(uint32_t)(g->dims[0] * g->dims[1] * g->dims[2]);

    int32_t ext = f32_as_i32(ceilf(max_search_radius / cell_size));
    if (ext < 0 || ext > 64) return -3;
    size_t cap = (size_t)(2 * ext + 1) * (2 * ext + 1) * (2 * ext + 1);
    g->offs = (int32_t(*)[3])malloc(cap * sizeof(*g->offs));
    g->n_offs = 0;
    for (int32_t dz = -ext; dz <= ext; ++dz)
        for (int32_t dy = -ext; dy <= ext; ++dy)
            for (int32_t dx = -ext; dx <= ext; ++dx)
                if (dz > 0 || (dz == 0 && dy > 0) || (dz == 0 && dy == 0 && dx >= 0)) {
                    g->offs[g->n_offs][0] = dx;
                    g->offs[g->n_offs][1] = dy;
                    g->offs[g->n_offs][2] = dz;
                    g->n_offs++;
                }

    uint32_t *counts = (uint32_t *)calloc(g->num_cells + 1, sizeof(uint32_t));
    g->cell_starts = (uint32_t *)calloc(g->num_cells + 1, sizeof(uint32_t));
    uint32_t *cells = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!counts || !g->cell_starts || !cells) return -1;
    for (size_t i = 0; i < n; ++i) {
        size_t c = cell_index(&xyzr[4 * i], minb, inv, g->dims);
        if (c >= g->num_cells) { /* the reference panics (index out of bounds) on non-finite input */
            free(counts);
            free(cells);
            return -2;
        }
        cells[i] = (uint32_t)c;
        counts[c]++;
    }
    for (size_t c = 0; c < g->num_cells; ++c) g->cell_starts[c + 1] = g->cell_starts[c] + counts[c];
    g->atom_indices = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    g->px = (float *)malloc((n ? n : 1) * sizeof(float));
    g->py = (float *)malloc((n ? n : 1) * sizeof(float));
    g->pz = (float *)malloc((n ? n : 1) * sizeof(float));
    g->rad = (float *)malloc((n ? n : 1) * sizeof(float));
    memcpy(counts, g->cell_starts, g->num_cells * sizeof(uint32_t)); /* write_pos */
    for (size_t i = 0; i < n; ++i) {
        uint32_t wp = counts[cells[i]]++;
        g->atom_indices[wp] = (uint32_t)i;
        g->px[wp] = xyzr[4 * i + 0];
        g->py[wp] = xyzr[4 * i + 1];
        g->pz[wp] = xyzr[4 * i + 2];
        g->rad[wp] = xyzr[4 * i + 3];
    }
    free(counts);
    free(cells);
    return 0;
}

static void grid_free(grid_t *g) {
    free(g->atom_indices);
    free(g->px);
    free(g->py);
    free(g->pz);
    free(g->rad);
    free(g->cell_starts);
    free(g->offs);
}

static inline uint64_t atom_id(const uint64_t *ids, size_t i) { return ids ? ids[i] : (uint64_t)i; }

/* process_self_cell / process_neighbor_cells, spatial_grid.rs:282-436 (the two
 * bodies are identical except for the j range). */
static inline void pair_sweep(const grid_t *g, const uint64_t *ids, size_t i, size_t j0, size_t j1, float probe,
                              float max_radius, float max_sr_sq, nbvec_t *nb) {
    const uint32_t oi = g->atom_indices[i];
    const float xi = g->px[i], yi = g->py[i], zi = g->pz[i], ri = g->rad[i];
    const uint64_t id_i = atom_id(ids, oi);
    const float sr_i = ri + max_radius + 2.0f * probe;
    const float sr_i_sq = sr_i * sr_i;
    for (size_t j = j0; j < j1; ++j) {
        const uint32_t oj = g->atom_indices[j];
        if (atom_id(ids, oj) == id_i) continue;
        const float dx = xi - g->px[j], dy = yi - g->py[j], dz = zi - g->pz[j];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > max_sr_sq) continue;
        const float rj = g->rad[j];
        const float sr_j = rj + max_radius + 2.0f * probe;
        const float sr_j_sq = sr_j * sr_j;
        if (d2 <= sr_i_sq) {
            const float tj = rj + probe;
            nb_push(&nb[oi], tj * tj, oj);
        }
        if (d2 <= sr_j_sq) {
            const float ti = ri + probe;
            nb_push(&nb[oj], ti * ti, oi);
        }
    }
}

typedef struct {
    const float *xyzr;
    const float *c;
} sortctx_t;
static __thread sortctx_t g_sortctx;
static int nb_cmp(const void *a, const void *b) {
    /* sort_neighbors_by_distance, spatial_grid.rs:438-465; ties broken by idx
     * (sort_unstable's tie order is unspecified and result-neutral). */
    const nb_t *na = (const nb_t *)a, *nb = (const nb_t *)b;
    const float *c = g_sortctx.c;
    const float *pa = &g_sortctx.xyzr[4 * (size_t)na->idx], *pb = &g_sortctx.xyzr[4 * (size_t)nb->idx];
    const float ax = c[0] - pa[0], ay = c[1] - pa[1], az = c[2] - pa[2];
    const float bx = c[0] - pb[0], by = c[1] - pb[1], bz = c[2] - pb[2];
    const float da = ax * ax + ay * ay + az * az;
    const float db = bx * bx + by * by + bz * bz;
    if (da < db) return -1;
    if (da > db) return 1;
    return (na->idx > nb->idx) - (na->idx < nb->idx);
}

/* precompute_neighbors, src/lib.rs:69-84 + build_all_neighbor_lists,
 * spatial_grid.rs:195-278.  Serial, like the reference. */
static int build_neighbors(const float *xyzr, const uint64_t *ids, size_t n, float probe, float max_radii,
                           nbvec_t *nb) {
    const float cell_size = probe + max_radii;
    const float max_search_radius = max_radii + max_radii + 2.0f * probe;
    grid_t g;
    memset(&g, 0, sizeof g);
    int rc = grid_build(&g, xyzr, n, cell_size, max_search_radius);
    if (rc) {
        grid_free(&g);
        return rc;
    }
    const float max_sr_sq = max_search_radius * max_search_radius;
    const uint32_t dxy = g.dims[0] * g.dims[1];
    for (size_t ca = 0; ca < g.num_cells; ++ca) {
        const size_t sa = g.cell_starts[ca], ea = g.cell_starts[ca + 1];
        if (sa == ea) continue;
        const int32_t cz = (int32_t)((uint32_t)ca / dxy);
        const uint32_t rem = (uint32_t)ca % dxy;
        const int32_t cy = (int32_t)(rem / g.dims[0]), cx = (int32_t)(rem % g.dims[0]);
        for (size_t o = 0; o < g.n_offs; ++o) {
            const int32_t bx = cx + g.offs[o][0], by = cy + g.offs[o][1], bz = cz + g.offs[o][2];
            if (bx < 0 || by < 0 || bz < 0) continue;
            if ((uint32_t)bx >= g.dims[0] || (uint32_t)by >= g.dims[1] || (uint32_t)bz >= g.dims[2]) continue;
            const size_t cb = (size_t)((uint32_t)bx + (uint32_t)by * g.dims[0] + (uint32_t)bz * dxy);
            const size_t sb = g.cell_starts[cb], eb = g.cell_starts[cb + 1];
            if (sb == eb) continue;
            const int is_self = g.offs[o][0] == 0 && g.offs[o][1] == 0 && g.offs[o][2] == 0;
            for (size_t i = sa; i < ea; ++i) {
                if (is_self)
                    pair_sweep(&g, ids, i, i + 1, ea, probe, max_radii, max_sr_sq, nb);
                else
                    pair_sweep(&g, ids, i, sb, eb, probe, max_radii, max_sr_sq, nb);
            }
        }
    }
    for (size_t i = 0; i < n; ++i) {
        if (nb[i].len <= 1) continue;
        g_sortctx.xyzr = xyzr;
        g_sortctx.c = &xyzr[4 * i];
        qsort(nb[i].v, nb[i].len, sizeof(nb_t), nb_cmp);
    }
    grid_free(&g);
    return 0;
}

/* ---------------------------------------------------------------------------
 * A3: AtomSasaKernel::with_simd, src/lib.rs:94-224, for one atom.
 *   lanes in {4, 8, 16} selects which pulp build is mirrored (NEON / AVX2 V3 /
 *   AVX-512 V4): points [0, lanes*floor(n/lanes)) take the SIMD body
 *   (fma-nested dot, strict <), the last n mod lanes points take the scalar
 *   tail (unfused dot, <=).  Returns the exposed-point count (an exact integer
 *   in the reference's f32 accumulator for every n < 2^24).
 * ------------------------------------------------------------------------- */
#define LMAX 16
static inline __attribute__((always_inline)) uint32_t atom_count_l(const float *xyzr, const uint64_t *ids, size_t i, const nb_t *nb, uint32_t k,
                                  const float *sx, const float *sy, const float *sz, size_t n_points, float probe,
                                  int lanes) {
    const float *ci = &xyzr[4 * i];
    const float r = ci[3] + probe;
    const float r2 = r * r;
    const uint64_t id_i = atom_id(ids, i);
    const size_t n_body = (n_points / (size_t)lanes) * (size_t)lanes;
    uint32_t accessible = 0;

    for (size_t p0 = 0; p0 < n_body; p0 += (size_t)lanes) {
        int occ[LMAX];
        for (int l = 0; l < lanes; ++l) occ[l] = 0;
        for (uint32_t q = 0; q < k; ++q) {
            const size_t j = nb[q].idx;
            if (atom_id(ids, j) == id_i) continue;
            const float vx = ci[0] - xyzr[4 * j + 0];
            const float vy = ci[1] - xyzr[4 * j + 1];
            const float vz = ci[2] - xyzr[4 * j + 2];
            const float vmag = vx * vx + vy * vy + vz * vz;
            const float limit = (nb[q].thr - vmag - r2) / (2.0f * r);
            int all = 1;
            for (int l = 0; l < lanes; ++l) {
                const float dot = fmaf(sx[p0 + l], vx, fmaf(sy[p0 + l], vy, sz[p0 + l] * vz));
                occ[l] |= (dot < limit);
                all &= occ[l];
            }
            if (all) break;
        }
        for (int l = 0; l < lanes; ++l) accessible += !occ[l];
    }
    /* scalar remainder, lib.rs:162-218 (the last-hit cache only reorders the
     * search; the predicate is an OR over the whole list either way). */
    for (size_t p = n_body; p < n_points; ++p) {
        int occluded = 0;
        for (uint32_t q = 0; q < k && !occluded; ++q) {
            const size_t j = nb[q].idx;
            if (atom_id(ids, j) == id_i) continue;
            const float vx = ci[0] - xyzr[4 * j + 0];
            const float vy = ci[1] - xyzr[4 * j + 1];
            const float vz = ci[2] - xyzr[4 * j + 2];
            const float vmag = vx * vx + vy * vy + vz * vz;
            const float limit = (nb[q].thr - vmag - r2) / (2.0f * r);
            const float dot = sx[p] * vx + sy[p] * vy + sz[p] * vz;
            if (dot <= limit) occluded = 1;
        }
        accessible += !occluded;
    }
    return accessible;
}

/* Constant-lane instantiations so the chunk loops compile to real SIMD in the
 * -O3 -march=native baseline build (what pulp's V3/V4/NEON dispatch does). */
static uint32_t atom_count(const float *xyzr, const uint64_t *ids, size_t i, const nb_t *nb, uint32_t k,
                           const float *sx, const float *sy, const float *sz, size_t n_points, float probe,
                           int lanes) {
    switch (lanes) {
    case 4: return atom_count_l(xyzr, ids, i, nb, k, sx, sy, sz, n_points, probe, 4);
    case 8: return atom_count_l(xyzr, ids, i, nb, k, sx, sy, sz, n_points, probe, 8);
    default: return atom_count_l(xyzr, ids, i, nb, k, sx, sy, sz, n_points, probe, 16);
    }
}

/* Area expression, lib.rs:220-222: ((4*PI_f32 * r2) * count) * (1/n). */
static inline float atom_area(float radius, float probe, uint32_t count, size_t n_points) {
    const float r = radius + probe;
    const float r2 = r * r;
    const float surface_area = (4.0f * 3.14159265358979323846f) * r2;
    const float inv_n = 1.0f / (float)n_points;
    return surface_area * (float)count * inv_n;
}

/* Number of sphere points of atom i whose test against some listed neighbour
 * lies within tol (A^2) of that neighbour's sphere: |d^2 - R_j^2| =
 * |2 R_i (dot - limit)| <= tol (SURVEY.md 8b "numerics contract").  Evaluated
 * in double from the f32 dot/limit the reference would compute. */
static uint32_t atom_boundary(const float *xyzr, const uint64_t *ids, size_t i, const nb_t *nb, uint32_t k,
                              const float *sx, const float *sy, const float *sz, size_t n_points, float probe,
                              int lanes, double tol) {
    const float *ci = &xyzr[4 * i];
    const float r = ci[3] + probe;
    const float r2 = r * r;
    const uint64_t id_i = atom_id(ids, i);
    const size_t n_body = (n_points / (size_t)lanes) * (size_t)lanes;
    uint32_t nbnd = 0;
    for (size_t p = 0; p < n_points; ++p) {
        int hit = 0;
        for (uint32_t q = 0; q < k && !hit; ++q) {
            const size_t j = nb[q].idx;
            if (atom_id(ids, j) == id_i) continue;
            const float vx = ci[0] - xyzr[4 * j + 0];
            const float vy = ci[1] - xyzr[4 * j + 1];
            const float vz = ci[2] - xyzr[4 * j + 2];
            const float vmag = vx * vx + vy * vy + vz * vz;
            const float limit = (nb[q].thr - vmag - r2) / (2.0f * r);
            const float dot = p < n_body ? fmaf(sx[p], vx, fmaf(sy[p], vy, sz[p] * vz))
                                         : sx[p] * vx + sy[p] * vy + sz[p] * vz;
            if (fabs(2.0 * (double)r * ((double)dot - (double)limit)) <= tol) hit = 1;
        }
        nbnd += (uint32_t)hit;
    }
    return nbnd;
}

/* ---------------------------------------------------------------------------
 * A4: calculate_sasa_internal, src/lib.rs:249-298.
 *   threads == 1 -> atoms sequential; anything else -> atoms across all cores
 *   (the reference's rayon global pool; here OpenMP).  The neighbour build is
 *   serial either way, exactly like the reference.
 *   Outputs (each nullable): out_sasa[n], out_counts[n], out_k[n] (= neighbour
 *   list length, the k_i of SURVEY.md 8d), out_boundary[n].
 *   Returns 0, or <0 on non-finite input (where the reference panics).
 * ------------------------------------------------------------------------- */
ORACLE_API int oracle_calculate_sasa_internal(const float *xyzr, const uint64_t *ids, size_t n, float probe,
                                              size_t n_points, int threads, int lanes, float *out_sasa,
                                              uint32_t *out_counts, uint32_t *out_k, uint32_t *out_boundary,
                                              double boundary_tol) {
    if (lanes != 4 && lanes != 8 && lanes != 16) return -3;
    if (n == 0) return 0; /* empty in -> empty out, tests/sanity.rs:148-157 */
    float *sx = (float *)malloc(3 * (n_points ? n_points : 1) * sizeof(float));
    float *sy = sx + n_points, *sz = sy + n_points;
    if (n_points) oracle_sphere_points(n_points, sx, sy, sz);

    float max_radii = 0.0f; /* fold(0.0, f32::max), lib.rs:259-262 */
    for (size_t i = 0; i < n; ++i) max_radii = fmaxf(max_radii, xyzr[4 * i + 3]);

    nbvec_t *nb = (nbvec_t *)calloc(n, sizeof(nbvec_t));
    int rc = build_neighbors(xyzr, ids, n, probe, max_radii, nb);
    if (rc == 0) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) if (threads != 1)
#endif
        for (size_t i = 0; i < n; ++i) {
            const uint32_t c = atom_count(xyzr, ids, i, nb[i].v, nb[i].len, sx, sy, sz, n_points, probe, lanes);
            if (out_counts) out_counts[i] = c;
            if (out_sasa) out_sasa[i] = atom_area(xyzr[4 * i + 3], probe, c, n_points);
            if (out_k) out_k[i] = nb[i].len;
            if (out_boundary)
                out_boundary[i] =
                    atom_boundary(xyzr, ids, i, nb[i].v, nb[i].len, sx, sy, sz, n_points, probe, lanes, boundary_tol);
        }
    }
    for (size_t i = 0; i < n; ++i) free(nb[i].v);
    free(nb);
    free(sx);
    (void)threads;
    return rc;
}

/* Neighbour lists as CSR, for restating tests/units.rs:131-209 (SpatialGrid with
 * an explicit cell size).  off has n+1 entries; idx/thr may be NULL to query
 * sizes only.  Returns total entries or <0. */
ORACLE_API long oracle_neighbor_lists(const float *xyzr, const uint64_t *ids, size_t n, float probe, float max_radius,
                                      float cell_size, uint32_t *off, uint32_t *idx, float *thr) {
    nbvec_t *nb = (nbvec_t *)calloc(n ? n : 1, sizeof(nbvec_t));
    grid_t g;
    memset(&g, 0, sizeof g);
    const float max_search_radius = max_radius + max_radius + 2.0f * probe;
    int rc;
    if (cell_size == probe + max_radius) {
        rc = build_neighbors(xyzr, ids, n, probe, max_radius, nb);
    } else {
        /* explicit-cell variant used by the reference's unit test */
        rc = grid_build(&g, xyzr, n, cell_size, max_search_radius);
        if (rc == 0) {
            const float max_sr_sq = max_search_radius * max_search_radius;
            const uint32_t dxy = g.dims[0] * g.dims[1];
            for (size_t ca = 0; ca < g.num_cells; ++ca) {
                const size_t sa = g.cell_starts[ca], ea = g.cell_starts[ca + 1];
                if (sa == ea) continue;
                const int32_t cz = (int32_t)((uint32_t)ca / dxy);
                const uint32_t rem = (uint32_t)ca % dxy;
                const int32_t cy = (int32_t)(rem / g.dims[0]), cx = (int32_t)(rem % g.dims[0]);
                for (size_t o = 0; o < g.n_offs; ++o) {
                    const int32_t bx = cx + g.offs[o][0], by = cy + g.offs[o][1], bz = cz + g.offs[o][2];
                    if (bx < 0 || by < 0 || bz < 0) continue;
                    if ((uint32_t)bx >= g.dims[0] || (uint32_t)by >= g.dims[1] || (uint32_t)bz >= g.dims[2])
                        continue;
                    const size_t cb = (size_t)((uint32_t)bx + (uint32_t)by * g.dims[0] + (uint32_t)bz * dxy);
                    const size_t sb = g.cell_starts[cb], eb = g.cell_starts[cb + 1];
                    if (sb == eb) continue;
                    const int is_self = !g.offs[o][0] && !g.offs[o][1] && !g.offs[o][2];
                    for (size_t i = sa; i < ea; ++i)
                        pair_sweep(&g, ids, i, is_self ? i + 1 : sb, is_self ? ea : eb, probe, max_radius, max_sr_sq,
                                   nb);
                }
            }
            for (size_t i = 0; i < n; ++i) {
                if (nb[i].len <= 1) continue;
                g_sortctx.xyzr = xyzr;
                g_sortctx.c = &xyzr[4 * i];
                qsort(nb[i].v, nb[i].len, sizeof(nb_t), nb_cmp);
            }
        }
        grid_free(&g);
    }
    long total = rc;
    if (rc == 0) {
        total = 0;
        for (size_t i = 0; i < n; ++i) {
            if (off) off[i] = (uint32_t)total;
            for (uint32_t q = 0; q < nb[i].len; ++q) {
                if (idx) idx[total + q] = nb[i].v[q].idx;
                if (thr) thr[total + q] = nb[i].v[q].thr;
            }
            total += nb[i].len;
        }
        if (off) off[n] = (uint32_t)total;
    }
    for (size_t i = 0; i < n; ++i) free(nb[i].v);
    free(nb);
    return total;
}

/* simd_sum, src/utils.rs:14-22: plain sequential f32 sum in slice order. */
static inline float seq_sum(const float *v, size_t b, size_t e) {
    float t = 0.0f;
    for (size_t i = b; i < e; ++i) t += v[i];
    return t;
}

/* A5 numeric part: process_atoms for Residue/Chain level, src/options.rs:195-232
 * and :292-315 -- each output segment is the sequential f32 sum of a contiguous
 * atom range [seg_be[2k], seg_be[2k+1]). */
ORACLE_API void oracle_segment_sums(const float *atom_sasa, const uint32_t *seg_be, size_t n_seg, float *out) {
    for (size_t k = 0; k < n_seg; ++k) out[k] = seq_sum(atom_sasa, seg_be[2 * k], seg_be[2 * k + 1]);
}

/* A5 ProteinLevel, src/options.rs:370-410: global_total = sequential sum over
 * all atoms; polar/non-polar = running f32 sums of residue sums in residue
 * order.  out3 = {global, polar, non_polar}. */
ORACLE_API void oracle_protein_totals(const float *atom_sasa, size_t n_atoms, const uint32_t *seg_be,
                                      const uint8_t *seg_polar, size_t n_seg, float *out3) {
    float polar = 0.0f, nonpolar = 0.0f;
    for (size_t k = 0; k < n_seg; ++k) {
        const float s = seq_sum(atom_sasa, seg_be[2 * k], seg_be[2 * k + 1]);
        if (seg_polar[k])
            polar += s;
        else
            nonpolar += s;
    }
    out3[0] = seq_sum(atom_sasa, 0, n_atoms);
    out3[1] = polar;
    out3[2] = nonpolar;
}

/* ---------------------------------------------------------------------------
 * CLI directory-mode analogue (src/main.rs:342-480): one structure per task,
 * tasks spread over all host cores, each structure single-threaded
 * (threads = 1, main.rs:375,439) -- the CPU baseline for batch configs.
 * struct_off has n_struct+1 atom offsets into xyzr.  Residue sums are written
 * when seg_be != NULL (seg_be entries are relative to the structure's first
 * atom; struct_seg_off has n_struct+1 offsets into the segment table).
 * ------------------------------------------------------------------------- */
ORACLE_API int oracle_run_batch(const float *xyzr, const uint64_t *struct_off, size_t n_struct, float probe,
                                size_t n_points, int lanes, int n_threads, float *out_sasa, uint32_t *out_counts,
                                const uint32_t *seg_be, const uint64_t *struct_seg_off, float *out_seg) {
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (size_t s = 0; s < n_struct; ++s) {
        const size_t a0 = struct_off[s], a1 = struct_off[s + 1];
        int rc = oracle_calculate_sasa_internal(xyzr + 4 * a0, NULL, a1 - a0, probe, n_points, 1, lanes,
                                                out_sasa + a0, out_counts ? out_counts + a0 : NULL, NULL, NULL, 0.0);
        if (rc) bad = rc;
        if (seg_be && out_seg) {
            const size_t g0 = struct_seg_off[s], g1 = struct_seg_off[s + 1];
            oracle_segment_sums(out_sasa + a0, seg_be + 2 * g0, g1 - g0, out_seg + g0);
        }
    }
    (void)n_threads;
    return bad;
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
