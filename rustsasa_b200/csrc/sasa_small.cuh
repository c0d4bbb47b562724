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
//
// Template parameters: NT threads, MINB resident CTAs per SM, HAS_CLS (Atom.id equality classes present),
// FAST (n_points <= 128: the tight pipeline of sasa_fast.cuh; otherwise the generic chunked routines).
#pragma once
#include "sasa_device.cuh"
#include "sasa_fast.cuh"

namespace sasa {

struct SmemAtoms {
    const float4 *a;
    __device__ __forceinline__ float4 operator()(int j) const { return a[j]; }
};

// Shared-memory layout.  Everything the per-atom loops touch sits at compile-time offsets (given the warp
// count), so no base pointer has to stay in a register: point table | per-warp entries | per-warp index
// lists | reduction scratch | misc | atoms | per-atom values | [classes] | original indices | cell table.
struct SmallLayout {
    size_t pts, ent, cand, red, misc, atom, val, cls, orig, cellw, total;
};

__host__ __device__ constexpr size_t small_fixed_bytes(int nwarps) {
    return 128 * 16 + (size_t)nwarps * kNbCap * 16 + (size_t)nwarps * kQueueCap * 2 + 32 * 8 * 4 + 64;
}

__host__ __device__ inline SmallLayout small_layout(uint32_t nmax, uint32_t cmax, int nwarps, bool has_cls) {
    SmallLayout L;
    size_t o = 0;
    L.pts = o;   o += 128 * 16;
    L.ent = o;   o += (size_t)nwarps * kNbCap * 16;
    L.cand = o;  o += (size_t)nwarps * kQueueCap * 2;
    L.red = o;   o += 32 * 8 * 4;
    L.misc = o;  o += 64;
    L.atom = o;  o += (size_t)nmax * 16;
    L.val = o;   o += (size_t)nmax * 4;
    L.cls = o;   o += has_cls ? (size_t)nmax * 4 : 0;
    L.orig = o;  o += (size_t)nmax * 2;
    o = (o + 3) & ~(size_t)3;
    L.cellw = o; o += (((size_t)cmax + 2 + 1) / 2) * 4;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// min or max of 8 per-thread values over the block; result valid in every thread.
template <int NT>
__device__ __forceinline__ void block_minmax8(float (&v)[8], float *red) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], __shfl_xor_sync(kFull, v[k], d));
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 8; ++k) red[w * 8 + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float r = red[k];
        for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i * 8 + k]);
        v[k] = r;
    }
    __syncthreads();
}

template <int NT, int MINB, bool HAS_CLS, bool FAST>
__global__ void __launch_bounds__(NT, MINB) sasa_small_kernel(const KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = NT / 32;
    constexpr size_t kOffEnt = 128 * 16, kOffCand = kOffEnt + (size_t)NW * kNbCap * 16,
                     kOffRed = kOffCand + (size_t)NW * kQueueCap * 2, kOffMisc = kOffRed + 32 * 8 * 4,
                     kOffAtom = kOffMisc + 64;
    static_assert(kOffAtom == small_fixed_bytes(NW), "layout mismatch");
    float4 *const s_ptab = reinterpret_cast<float4 *>(smem);
    float4 *const s_atom = reinterpret_cast<float4 *>(smem + kOffAtom);
    float *const s_red = reinterpret_cast<float *>(smem + kOffRed);
    int *const s_misc = reinterpret_cast<int *>(smem + kOffMisc);
    const SmallLayout L = small_layout(p.nmax, p.cmax, NW, HAS_CLS);
    float *s_val = reinterpret_cast<float *>(smem + L.val);
    uint16_t *s_cellid = reinterpret_cast<uint16_t *>(smem + L.val);  // aliases s_val during the sort
    uint16_t *s_rank = s_cellid + p.nmax;
    uint32_t *s_cls = HAS_CLS ? reinterpret_cast<uint32_t *>(smem + L.cls) : nullptr;
    uint16_t *s_orig = reinterpret_cast<uint16_t *>(smem + L.orig);
    uint32_t *s_cellw = reinterpret_cast<uint32_t *>(smem + L.cellw);
    uint16_t *s_cell = reinterpret_cast<uint16_t *>(smem + L.cellw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 *const w_ent = reinterpret_cast<float4 *>(smem + kOffEnt) + warp * kNbCap;
    uint16_t *const w_cand = reinterpret_cast<uint16_t *>(smem + kOffCand) + warp * kQueueCap;
    const bool force_stream = (p.flags & 2u) != 0, stats = (p.flags & 1u) != 0;
    // the whole point set as a float4 table when it fits (n_points <= 128)
    const float4 *s_pts = p.n_points <= 128 ? s_ptab : nullptr;
    if (tid < 128) {
        const bool v = (uint32_t)tid < p.n_points;
        s_ptab[tid] = make_float4(v ? __ldg(p.px + tid) : 0.f, v ? __ldg(p.py + tid) : 0.f, v ? __ldg(p.pz + tid) : 0.f, 0.f);
    }
    const int nbody = (int)min(p.n_points, p.n_body), nsl = (nbody + 31) >> 5;

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

        // ---- bounds, r_max, finiteness (one fused block reduction: maxima of {-min, max, r, bad}) -------------
        float red8[8] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0f, 0.0f};
        for (int i = tid; i < N; i += NT) {
            const float4 a = __ldg(gat + i);
            const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(a.w);
            red8[0] = fmaxf(red8[0], -a.x); red8[1] = fmaxf(red8[1], -a.y); red8[2] = fmaxf(red8[2], -a.z);
            red8[3] = fmaxf(red8[3], a.x);  red8[4] = fmaxf(red8[4], a.y);  red8[5] = fmaxf(red8[5], a.z);
            red8[6] = fmaxf(red8[6], a.w);
            if (!fin) red8[7] = 1.0f;
        }
        block_minmax8<NT>(red8, s_red);
        const float mnx = -red8[0], mny = -red8[1], mnz = -red8[2], mxx = red8[3], mxy = red8[4], mxz = red8[5],
                    rmax = red8[6];
        if (red8[7] != 0.0f) {
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

        // ---- per-atom work: warps pull runs of consecutive cell-sorted atoms from a shared counter ----------
        const SmemAtoms atoms{s_atom};
        unsigned pairs = 0, streamed = 0;
        constexpr int kFetch = SASA_FETCH;
        CandCache<uint16_t> cc;
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_misc[1], kFetch);
            base = __shfl_sync(kFull, base, 0);
            if (base >= N) break;
            int cell_end = -1;   // atoms [.., cell_end) share the cached candidate list
            cc.total = -1;
            const int pend = min(base + kFetch, N);
            for (int pos = base; pos < pend; ++pos) {
                const float4 ai = s_atom[pos];
                int cnt;
                int k = -1;
                if (!force_stream && !stats) {
                    if (pos >= cell_end) {
                        const int cx = cell_coord(ai.x, g.minx, g.inv_c, g.nx), cy = cell_coord(ai.y, g.miny, g.inv_c, g.ny),
                                  cz = cell_coord(ai.z, g.minz, g.inv_c, g.nz);
                        const int cid = (cz * g.ny + cy) * g.nx + cx;
                        cell_end = (int)s_cell[cid + 1];
                        fill_cache(g, s_cell, cx, cy, cz, cid, cc);
                    }
                    if (cc.total >= 0) {
                        if (FAST) k = fast_gather<HAS_CLS>(s_atom, s_cls, pos, ai, ai.w + 2.0f * p.probe + kCutSlack, cc, w_cand);
                        else k = gather_cached(p, atoms, s_cls, pos, ai, cc, w_cand);
                    } else {
                        k = gather_candidates(p, g, atoms, s_cell, s_cls, pos, ai, w_cand);
                    }
                    if (k > kNbCap) k = -1;
                }
                if (k >= 0) {
                    const float r = __fadd_rn(ai.w, p.probe);
                    if (FAST) {
                        const int nfront = fast_entries(s_atom, ai, p.probe, __fmul_rn(r, r), __fmul_rn(2.0f, r), p.near2, w_cand, k, w_ent);
                        cnt = fast_atom(p, w_ent, k, nfront, s_ptab, w_cand, nbody, nsl);
                    } else {
                        const int nfront = build_entries(p, atoms, ai, __fmul_rn(r, r), __fmul_rn(2.0f, r), w_cand, k, w_ent);
                        cnt = (int)atom_fast(p, w_ent, k, nfront, w_cand, s_pts);
                    }
                    pairs += (unsigned)k;
                } else {
                    cnt = (int)(stats ? atom_streaming<SmemAtoms, uint16_t, true>(p, g, atoms, s_cell, s_cls, pos, w_ent, p.stat)
                                      : atom_streaming<SmemAtoms, uint16_t, false>(p, g, atoms, s_cell, s_cls, pos, w_ent, p.stat));
                    streamed += 1;
                }
                if (lane == 0) s_val[s_orig[pos]] = (float)cnt;
                __syncwarp();
            }
        }
        if (lane == 0 && p.stat) {
            if (pairs) atomicAdd(p.stat + 1, (unsigned long long)pairs);
            if (streamed) atomicAdd(p.stat + 2, (unsigned long long)streamed);
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
