// sasa_small.cuh -- the fused per-structure kernel: one CTA owns one structure at a time and keeps
// it entirely in shared memory from the raw float4 atoms to the residue / chain / protein sums:
//
//   bounds -> cell grid -> counting sort (shared-memory atomics) -> per-atom neighbour gather ->
//   occlusion test -> per-atom counts/areas -> segment sums
//
// replacing, for structures of up to `nmax` atoms, the reference's SpatialGrid::new +
// build_all_neighbor_lists (src/structures/spatial_grid.rs:28-465), AtomSasaKernel::with_simd
// (src/lib.rs:94-224), the driver loop (src/lib.rs:249-298) and the numeric part of process_atoms
// (src/options.rs:195-232, :292-315, :370-410).  CTAs are persistent and pull structures from an
// atomic queue ordered largest-first by the host.
#pragma once
#include "sasa_device.cuh"

namespace sasa {

struct SmemAtoms {
    const float4 *a;
    __device__ __forceinline__ float4 operator()(int j) const { return a[j]; }
};

struct SmallLayout {
    size_t atom, ent, pts, val, cls, cellw, red, misc, orig, cand, total;
};

__host__ __device__ inline SmallLayout small_layout(uint32_t nmax, uint32_t cmax, int nwarps, bool has_cls) {
    SmallLayout L;
    size_t o = 0;
    L.atom = o;  o += (size_t)nmax * 16;
    L.ent = o;   o += (size_t)nwarps * kNbCap * 16;
    L.pts = o;   o += 128 * 16;
    L.val = o;   o += (size_t)nmax * 4;
    L.cls = o;   o += has_cls ? (size_t)nmax * 4 : 0;
    L.cellw = o; o += (((size_t)cmax + 2 + 1) / 2) * 4;
    L.red = o;   o += 32 * 8 * 4;
    L.misc = o;  o += 64;
    L.orig = o;  o += (size_t)nmax * 2;
    L.cand = o;  o += (size_t)nwarps * kQueueCap * 2;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

template <int NT>
__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float *red) {
    // returns the reduction in every thread; red must hold NT/32 floats; caller syncs between uses
    const int lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o = __shfl_xor_sync(kFull, v, d);
        v = is_max ? fmaxf(v, o) : fminf(v, o);
    }
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < NT / 32; ++i) r = is_max ? fmaxf(r, red[i]) : fminf(r, red[i]);
    __syncthreads();
    return r;
}

template <int NT, int MINB, bool HAS_CLS>
__global__ void __launch_bounds__(NT, MINB) sasa_small_kernel(const KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = NT / 32;
    const SmallLayout L = small_layout(p.nmax, p.cmax, NW, HAS_CLS);
    float4 *s_atom = reinterpret_cast<float4 *>(smem + L.atom);
    float4 *s_ent = reinterpret_cast<float4 *>(smem + L.ent);
    float4 *s_ptab = reinterpret_cast<float4 *>(smem + L.pts);
    float *s_val = reinterpret_cast<float *>(smem + L.val);
    uint16_t *s_cellid = reinterpret_cast<uint16_t *>(smem + L.val);  // aliases s_val during the sort
    uint16_t *s_rank = s_cellid + p.nmax;
    uint32_t *s_cls = HAS_CLS ? reinterpret_cast<uint32_t *>(smem + L.cls) : nullptr;
    uint32_t *s_cellw = reinterpret_cast<uint32_t *>(smem + L.cellw);
    uint16_t *s_cell = reinterpret_cast<uint16_t *>(smem + L.cellw);
    float *s_red = reinterpret_cast<float *>(smem + L.red);
    int *s_misc = reinterpret_cast<int *>(smem + L.misc);
    uint16_t *s_orig = reinterpret_cast<uint16_t *>(smem + L.orig);
    uint16_t *s_cand = reinterpret_cast<uint16_t *>(smem + L.cand);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 *w_ent = s_ent + warp * kNbCap;
    uint16_t *w_cand = s_cand + warp * kQueueCap;
    const bool force_stream = (p.flags & 2u) != 0, stats = (p.flags & 1u) != 0, use_cache = (p.flags & 4u) == 0;
    // the whole point set as a float4 table when it fits (n_points <= 128): phase 2 fetches survivors from it
    const float4 *s_pts = p.n_points <= 128 ? s_ptab : nullptr;
    if (tid < 128) {
        const bool v = (uint32_t)tid < p.n_points;
        s_ptab[tid] = make_float4(v ? __ldg(p.px + tid) : 0.f, v ? __ldg(p.py + tid) : 0.f, v ? __ldg(p.pz + tid) : 0.f, 0.f);
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[0] = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t w = (uint32_t)s_misc[0];
        if (w >= p.n_work) break;
        const uint32_t sid = p.order[w];
        const uint32_t a0 = p.struct_off[sid];
        const int N = (int)(p.struct_off[sid + 1] - a0);
        const float4 *gat = p.xyzr + a0;

        // ---- bounds, r_max, finiteness -------------------------------------------------------
        float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY,
              rmax = 0.0f;
        bool finite = true;
        for (int i = tid; i < N; i += NT) {
            const float4 a = __ldg(gat + i);
            finite = finite && isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(a.w);
            mnx = fminf(mnx, a.x); mny = fminf(mny, a.y); mnz = fminf(mnz, a.z);
            mxx = fmaxf(mxx, a.x); mxy = fmaxf(mxy, a.y); mxz = fmaxf(mxz, a.z);
            rmax = fmaxf(rmax, a.w);
        }
        mnx = block_reduce_minmax<NT>(mnx, false, s_red);
        mny = block_reduce_minmax<NT>(mny, false, s_red);
        mnz = block_reduce_minmax<NT>(mnz, false, s_red);
        mxx = block_reduce_minmax<NT>(mxx, true, s_red);
        mxy = block_reduce_minmax<NT>(mxy, true, s_red);
        mxz = block_reduce_minmax<NT>(mxz, true, s_red);
        rmax = block_reduce_minmax<NT>(rmax, true, s_red);
        const float bad = block_reduce_minmax<NT>(finite ? 0.0f : 1.0f, true, s_red);
        if (bad != 0.0f) {
            // the reference panics on non-finite input; report it and emit NaN for this structure
            if (tid == 0) atomicExch(p.err_flag, 4);
            const float qn = __int_as_float(0x7fc00000);
            for (int i = tid; i < N; i += NT) {
                if (p.out_counts) p.out_counts[a0 + i] = 0u;
                if (p.out_atom) p.out_atom[a0 + i] = qn;
            }
            if (p.seg_be && p.out_seg)
                for (uint32_t k = p.struct_seg_off[sid] + tid; k < p.struct_seg_off[sid + 1]; k += NT) p.out_seg[k] = qn;
            if (p.out_protein && tid < 3) p.out_protein[3 * (size_t)sid + tid] = qn;
            continue;
        }

        // ---- cell grid: cell edge >= half the largest possible pair cutoff, grown until it fits ----
        Grid g;
        {
            const float cutoff = (2.0f * rmax + 2.0f * p.probe + kCutSlack) * kCellSafety;
            float c = 0.5f * cutoff;
            const float ex = fmaxf(mxx - mnx, 0.0f), ey = fmaxf(mxy - mny, 0.0f), ez = fmaxf(mxz - mnz, 0.0f);
            float fx, fy, fz;
            for (int it = 0; it < 64; ++it) {
                fx = floorf(ex / c) + 1.0f; fy = floorf(ey / c) + 1.0f; fz = floorf(ez / c) + 1.0f;
                const float nc = fx * fy * fz;
                if (nc <= (float)p.cmax) break;
                c *= fmaxf(1.05f, cbrtf(nc / (float)p.cmax));
            }
            g.minx = mnx; g.miny = mny; g.minz = mnz;
            g.inv_c = 1.0f / c;
            g.nx = (int)fx; g.ny = (int)fy; g.nz = (int)fz;
            g.e = (c >= cutoff) ? 1 : 2;
        }
        const int ncell = g.nx * g.ny * g.nz;

        // ---- counting sort into cells ---------------------------------------------------------
        for (int i = tid; i < (ncell + 2 + 1) / 2; i += NT) s_cellw[i] = 0u;
        __syncthreads();
        for (int i = tid; i < N; i += NT) {
            const float4 a = __ldg(gat + i);
            const int c = (cell_coord(a.z, g.minz, g.inv_c, g.nz) * g.ny + cell_coord(a.y, g.miny, g.inv_c, g.ny)) * g.nx +
                          cell_coord(a.x, g.minx, g.inv_c, g.nx);
            const uint32_t old = atomicAdd(&s_cellw[c >> 1], (c & 1) ? 0x10000u : 1u);
            s_cellid[i] = (uint16_t)c;
            s_rank[i] = (uint16_t)((c & 1) ? (old >> 16) : (old & 0xffffu));
        }
        __syncthreads();
        {   // exclusive scan of the u16 counts -> cell starts; s_cell[ncell] = N
            const int per = (ncell + NT - 1) / NT;
            const int b = min(tid * per, ncell), e = min(b + per, ncell);
            int sum = 0;
            for (int c = b; c < e; ++c) sum += s_cell[c];
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            int *s_wsum = reinterpret_cast<int *>(s_red);
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int base = 0;
            for (int i = 0; i < warp; ++i) base += s_wsum[i];
            int run = base + incl - sum;
            for (int c = b; c < e; ++c) {
                const int cnt = s_cell[c];
                s_cell[c] = (uint16_t)run;
                run += cnt;
            }
            if (tid == 0) s_cell[ncell] = (uint16_t)N;
            __syncthreads();
        }
        for (int i = tid; i < N; i += NT) {
            const int at = (int)s_cell[s_cellid[i]] + (int)s_rank[i];
            s_atom[at] = __ldg(gat + i);
            s_orig[at] = (uint16_t)i;
            if (HAS_CLS) s_cls[at] = p.cls[a0 + i];
        }
        if (tid == 0) s_misc[1] = 0;
        __syncthreads();   // s_cellid / s_rank are dead from here on: s_val may be written

        // ---- per-atom work: warps pull atoms (in cell order) from a shared counter ---------------
        const SmemAtoms atoms{s_atom};
        unsigned long long pairs = 0, streamed = 0;
        constexpr int kFetch = SASA_FETCH;   // consecutive cell-sorted atoms per fetch: neighbours in the list share cells
        CandCache<uint16_t> cc;
#if SASA_PRELOAD
        PointChunk pre;
        load_chunk(p, s_pts, 0, pre);
        const PointChunk *prep = p.n_points <= 128 ? &pre : nullptr;
#else
        const PointChunk *prep = nullptr;
#endif
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_misc[1], kFetch);
            base = __shfl_sync(kFull, base, 0);
            if (base >= N) break;
            cc.cell = -1;
            cc.total = -1;
            const int pend = min(base + kFetch, N);
            for (int pos = base; pos < pend; ++pos) {
                const float4 ai = s_atom[pos];
                float cnt;
                int k = -1;
                if (!force_stream && !stats) {
                    const int cx = cell_coord(ai.x, g.minx, g.inv_c, g.nx), cy = cell_coord(ai.y, g.miny, g.inv_c, g.ny),
                              cz = cell_coord(ai.z, g.minz, g.inv_c, g.nz);
                    const int cid = (cz * g.ny + cy) * g.nx + cx;
                    if (use_cache && cid != cc.cell) fill_cache(g, s_cell, cx, cy, cz, cid, cc);
                    k = cc.total >= 0 ? gather_cached(p, atoms, s_cls, pos, ai, cc, w_cand)
                                      : gather_candidates(p, g, atoms, s_cell, s_cls, pos, ai, w_cand);
                }
                if (k >= 0) {
                    const float r = __fadd_rn(ai.w, p.probe);
                    const int nfront = build_entries(p, atoms, ai, __fmul_rn(r, r), __fmul_rn(2.0f, r), w_cand, k, w_ent);
                    cnt = atom_fast(p, w_ent, k, nfront, w_cand, s_pts, prep);
                    pairs += (unsigned)k;
                } else {
                    cnt = stats ? atom_streaming<SmemAtoms, uint16_t, true>(p, g, atoms, s_cell, s_cls, pos, w_ent, p.stat)
                                : atom_streaming<SmemAtoms, uint16_t, false>(p, g, atoms, s_cell, s_cls, pos, w_ent, p.stat);
                    streamed += 1;
                }
                if (lane == 0) s_val[s_orig[pos]] = cnt;
                __syncwarp();
            }
        }
        if (lane == 0 && p.stat) {
            if (pairs) atomicAdd(p.stat + 1, pairs);
            if (streamed) atomicAdd(p.stat + 2, streamed);
        }
        __syncthreads();

        // ---- outputs: per-atom counts / areas (coalesced), then sums in the reference's order ------
        for (int i = tid; i < N; i += NT) {
            const float cnt = s_val[i];
            const float area = atom_area(__ldg(gat + i).w, p.probe, cnt, p.inv_n);
            if (p.out_counts) p.out_counts[a0 + i] = (uint32_t)cnt;
            if (p.out_atom) p.out_atom[a0 + i] = area;
            s_val[i] = area;
        }
        __syncthreads();
        if (p.seg_be) {
            const uint32_t g0 = p.struct_seg_off[sid], g1 = p.struct_seg_off[sid + 1];
            if (p.out_seg) {
                // simd_sum (src/utils.rs:14-22): sequential f32 sum in atom order, one thread per segment
                for (uint32_t k = g0 + tid; k < g1; k += NT) {
                    const uint2 be = p.seg_be[k];
                    float t = 0.0f;
                    for (uint32_t i = be.x; i < be.y && i < (uint32_t)N; ++i) t = __fadd_rn(t, s_val[i]);
                    p.out_seg[k] = t;
                }
            }
            if (p.out_protein && tid == 32 % NT) {
                // polar / non-polar: running sums of residue sums in residue order (src/options.rs:376-403)
                float polar = 0.0f, nonpolar = 0.0f;
                for (uint32_t k = g0; k < g1; ++k) {
                    const uint2 be = p.seg_be[k];
                    float t = 0.0f;
                    for (uint32_t i = be.x; i < be.y && i < (uint32_t)N; ++i) t = __fadd_rn(t, s_val[i]);
                    if (p.seg_polar && p.seg_polar[k]) polar = __fadd_rn(polar, t);
                    else nonpolar = __fadd_rn(nonpolar, t);
                }
                p.out_protein[3 * (size_t)sid + 1] = polar;
                p.out_protein[3 * (size_t)sid + 2] = nonpolar;
            }
        }
        if (p.out_protein && tid == 0) {
            float t = 0.0f;   // global_total = simd_sum(atom_sasa), src/options.rs:404
            for (int i = 0; i < N; ++i) t = __fadd_rn(t, s_val[i]);
            p.out_protein[3 * (size_t)sid + 0] = t;
            if (!p.seg_be) { p.out_protein[3 * (size_t)sid + 1] = 0.0f; p.out_protein[3 * (size_t)sid + 2] = t; }
        }
    }
}

}  // namespace sasa
