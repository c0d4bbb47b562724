// sasa_large.cuh -- structures too large for the fused shared-memory kernel (ribosome / capsid scale).
//
// Same algorithm as sasa_small.cuh with the cell list in global memory: bounds -> dense cell grid ->
// counting sort by cell (global atomics + three-kernel exclusive scan) -> one warp per atom gathers its
// neighbours from the cell-sorted float4 array through L1/L2 and runs the same occlusion routines.
// Replaces SpatialGrid::new / build_all_neighbor_lists (src/structures/spatial_grid.rs:28-465), which the
// reference runs serially, for N up to 2^32 / 16 atoms.
#pragma once
#include "sasa_cap.cuh"
#include "sasa_device.cuh"

namespace sasa {

struct GlobalAtoms {
    const float4 *a;
    __device__ __forceinline__ float4 operator()(int j) const { return __ldg(a + j); }
};

struct LargeHeader {
    unsigned enc[8];   // order-preserving encodings of min xyz, max xyz, rmax; [7] = non-finite flag
    Grid grid;
    int ncell;
    unsigned next_atom;   // work counter of large_atoms_kernel: next cell-sorted position to hand out
    unsigned range_end;   // atom-range split: end of this rank's slice (a cell boundary)
};

struct LargeWorkspace {
    uint32_t cap_atoms = 0, cap_cells = 0;
    float4 *sorted = nullptr;
    uint32_t *orig = nullptr, *cellid = nullptr, *rank = nullptr, *cls_sorted = nullptr, *cells = nullptr,
             *blocksum = nullptr;
    float *val = nullptr;
    LargeHeader *hdr = nullptr;
};

constexpr int kScanItems = 2048;  // cells per scan block (256 threads x 8)

__device__ __forceinline__ unsigned enc_f(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void large_init_kernel(LargeHeader *h, unsigned first_atom) {
    if (threadIdx.x < 3) h->enc[threadIdx.x] = 0xffffffffu;          // running minima
    else if (threadIdx.x < 8) h->enc[threadIdx.x] = 0u;              // running maxima, flag
    if (threadIdx.x == 0) { h->next_atom = first_atom; h->range_end = 0u; }
}

__global__ void __launch_bounds__(256) large_bounds_kernel(const float4 *__restrict__ at, int N, LargeHeader *h) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, rmax = 0.0f;
    bool finite = true;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 a = __ldg(at + i);
        finite = finite && isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(a.w);
        mn[0] = fminf(mn[0], a.x); mn[1] = fminf(mn[1], a.y); mn[2] = fminf(mn[2], a.z);
        mx[0] = fmaxf(mx[0], a.x); mx[1] = fmaxf(mx[1], a.y); mx[2] = fmaxf(mx[2], a.z);
        rmax = fmaxf(rmax, a.w);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(kFull, mn[k], d));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(kFull, mx[k], d));
        }
        rmax = fmaxf(rmax, __shfl_xor_sync(kFull, rmax, d));
    }
    const bool any_bad = __any_sync(kFull, !finite);
    if ((threadIdx.x & 31) == 0) {
        if (any_bad) atomicExch(&h->enc[7], 1u);
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicMin(&h->enc[k], enc_f(mn[k]));
                atomicMax(&h->enc[3 + k], enc_f(mx[k]));
            }
            atomicMax(&h->enc[6], enc_f(rmax));
        }
    }
}

__global__ void large_grid_kernel(LargeHeader *h, float probe, uint32_t cmax, int *err_flag) {
    if (threadIdx.x != 0) return;
    if (h->enc[7]) {
        atomicExch(err_flag, 4);
        h->ncell = 0;
        h->grid = Grid{0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0};
        return;
    }
    const float mnx = dec_f(h->enc[0]), mny = dec_f(h->enc[1]), mnz = dec_f(h->enc[2]);
    const float mxx = dec_f(h->enc[3]), mxy = dec_f(h->enc[4]), mxz = dec_f(h->enc[5]), rmax = dec_f(h->enc[6]);
    const float cutoff = (2.0f * rmax + 2.0f * probe + kCutSlack) * kCellSafety;
    float c = 0.5f * cutoff;
    const float ex = fmaxf(mxx - mnx, 0.0f), ey = fmaxf(mxy - mny, 0.0f), ez = fmaxf(mxz - mnz, 0.0f);
    float fx = 1.f, fy = 1.f, fz = 1.f;
    for (int it = 0; it < 64; ++it) {
        fx = floorf(ex / c) + 1.0f; fy = floorf(ey / c) + 1.0f; fz = floorf(ez / c) + 1.0f;
        const float nc = fx * fy * fz;
        if (nc <= (float)cmax) break;
        c *= fmaxf(1.05f, cbrtf(nc / (float)cmax));
    }
    Grid g;
    g.minx = mnx; g.miny = mny; g.minz = mnz;
    g.inv_c = 1.0f / c;
    g.nx = (int)fx; g.ny = (int)fy; g.nz = (int)fz;
    g.e = (c >= cutoff) ? 1 : 2;
    h->grid = g;
    h->ncell = g.nx * g.ny * g.nz;
}

__global__ void __launch_bounds__(256) large_zero_kernel(const LargeHeader *h, uint32_t *cells) {
    const int n = h->ncell + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cells[i] = 0u;
}

__global__ void __launch_bounds__(256) large_count_kernel(const float4 *__restrict__ at, int N, const LargeHeader *h,
                                                          uint32_t *cells, uint32_t *cellid, uint32_t *rank) {
    const Grid g = h->grid;
    if (h->ncell == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 a = __ldg(at + i);
        const int c = (cell_coord(a.z, g.minz, g.inv_c, g.nz) * g.ny + cell_coord(a.y, g.miny, g.inv_c, g.ny)) * g.nx +
                      cell_coord(a.x, g.minx, g.inv_c, g.nx);
        cellid[i] = (uint32_t)c;
        rank[i] = atomicAdd(&cells[c], 1u);
    }
}

// Exclusive scan of cells[0, ncell) in three kernels (block sums -> scan of sums -> local scan + offset).
__global__ void __launch_bounds__(256) large_scan1_kernel(const LargeHeader *h, const uint32_t *cells, uint32_t *blocksum) {
    const int n = h->ncell;
    const int b0 = blockIdx.x * kScanItems;
    if (b0 >= n) return;
    uint32_t s = 0;
    for (int i = b0 + threadIdx.x; i < min(b0 + kScanItems, n); i += 256) s += cells[i];
    __shared__ uint32_t red[8];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(kFull, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        blocksum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) large_scan2_kernel(const LargeHeader *h, uint32_t *blocksum) {
    const int nb = (h->ncell + kScanItems - 1) / kScanItems;
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? blocksum[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t base = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += wsum[w];
        if (i < nb) blocksum[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = base + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) large_scan3_kernel(const LargeHeader *h, uint32_t *cells, const uint32_t *blocksum, uint32_t N) {
    const int n = h->ncell;
    const int b0 = blockIdx.x * kScanItems;
    if (b0 >= n) return;
    // each thread owns 8 consecutive cells
    const int t0 = b0 + threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = (t0 + k < n) ? cells[t0 + k] : 0u;
        s += v[k];
    }
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    __shared__ uint32_t wsum[8];
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t run = blocksum[blockIdx.x] + incl - s;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) run += wsum[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (t0 + k < n) cells[t0 + k] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cells[n] = N;
}

__global__ void __launch_bounds__(256) large_scatter_kernel(const float4 *__restrict__ at, const uint32_t *__restrict__ cls,
                                                            int N, const LargeHeader *h, const uint32_t *cells,
                                                            const uint32_t *cellid, const uint32_t *rank, float4 *sorted,
                                                            uint32_t *orig, uint32_t *cls_sorted) {
    if (h->ncell == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const uint32_t at_pos = cells[cellid[i]] + rank[i];
        sorted[at_pos] = __ldg(at + i);
        orig[at_pos] = (uint32_t)i;
        if (cls) cls_sorted[at_pos] = cls[i];
    }
}

// Atom-range split: turn the ideal cut points N*r/n and N*(r+1)/n into CELL boundaries of the sorted order.  The
// order of atoms inside a cell comes from atomics and differs from GPU to GPU, but the cell starts are the same
// everywhere, so slices cut at cell boundaries partition the atoms identically on every rank.
__global__ void large_range_kernel(LargeHeader *h, const uint32_t *cells, uint32_t target_lo, uint32_t target_hi, uint32_t N) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = h->ncell;
    auto first_start_at_or_after = [&](uint32_t target) -> uint32_t {
        if (target >= N || n == 0) return N;
        int lo = 0, hi = n;            // cells[n] == N
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cells[mid] >= target) hi = mid;
            else lo = mid + 1;
        }
        return cells[lo];
    };
    h->next_atom = first_start_at_or_after(target_lo);
    h->range_end = first_start_at_or_after(target_hi);
}

// One warp per atom; warps pull consecutive cell-sorted atoms from a global counter (which starts at the first
// position of this launch's range) so that the warps of a CTA share candidate cells in L1.  Only sorted positions
// below N are evaluated: N = the structure's atom count, or the end of this rank's slice in the atom-range split.
__global__ void __launch_bounds__(256, 2) large_atoms_kernel(const KParams p, int N, uint32_t a0, LargeHeader *h,
                                                             const float4 *__restrict__ sorted, const uint32_t *__restrict__ orig,
                                                             const uint32_t *__restrict__ cls_sorted, const uint32_t *__restrict__ cells,
                                                             float *val) {
    __shared__ __align__(16) float4 s_ent[8 * kNbCap];
    __shared__ uint32_t s_cand[8 * kNbCap];       // u32 candidate positions; reused as the u16 survivor queue
    static_assert(kNbCap * 2 >= kQueueCap, "survivor queue must fit the candidate list");
    __shared__ __align__(16) float4 s_ptab[128];
    if (h->ncell == 0) return;
    if (N < 0) N = (int)h->range_end;   // atom-range split: the slice end was computed on the device
    const Grid g = h->grid;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    float4 *w_ent = s_ent + warp * kNbCap;
    uint32_t *w_cand = s_cand + warp * kNbCap;
    const GlobalAtoms atoms{sorted};
    const bool force_stream = (p.flags & 2u) != 0, stats = (p.flags & 1u) != 0, use_cache = (p.flags & 4u) == 0;
    unsigned long long pairs = 0, streamed = 0;
    const float4 *s_pts = p.n_points <= 128 ? s_ptab : nullptr;
    const bool use_cap = p.cap != nullptr && p.n_points <= 128;
    const int nbody = (int)min(p.n_points, p.n_body);
    if (threadIdx.x < 128) {
        const bool v = threadIdx.x < p.n_points;
        s_ptab[threadIdx.x] = make_float4(v ? __ldg(p.px + threadIdx.x) : 0.f, v ? __ldg(p.py + threadIdx.x) : 0.f,
                                          v ? __ldg(p.pz + threadIdx.x) : 0.f, 0.f);
    }
    __syncthreads();
    constexpr unsigned kFetch = SASA_FETCH;
    CandCache<uint32_t> cc;
    for (;;) {
        unsigned base_u = 0;
        if (lane == 0) base_u = atomicAdd(&h->next_atom, kFetch);
        const int base = (int)__shfl_sync(kFull, base_u, 0);
        if (base >= N) break;
        cc.cell = -1;
        cc.total = -1;
        const int pend = min(base + (int)kFetch, N);
        for (int pos = base; pos < pend; ++pos) {
            const float4 ai = atoms(pos);
            float cnt;
            int k = -1;
            if (!force_stream && !stats) {
                const int cx = cell_coord(ai.x, g.minx, g.inv_c, g.nx), cy = cell_coord(ai.y, g.miny, g.inv_c, g.ny),
                          cz = cell_coord(ai.z, g.minz, g.inv_c, g.nz);
                const int cid = (cz * g.ny + cy) * g.nx + cx;
                if (use_cache && cid != cc.cell) fill_cache(g, cells, cx, cy, cz, cid, cc);
                k = cc.total >= 0 ? gather_cached(p, atoms, cls_sorted, pos, ai, cc, w_cand)
                                  : gather_candidates(p, g, atoms, cells, cls_sorted, pos, ai, w_cand);
            }
            if (k >= 0 && use_cap) {
                // n_points <= 128: the cap-table occlusion of the fused kernel (sasa_cap.cuh), atoms read from global memory
                cnt = (float)cap_atom(p.cap, atoms, ai, p.probe, w_cand, k, s_ptab, (int)p.n_points, nbody);
                pairs += (unsigned)k;
            } else if (k >= 0) {
                const float r = __fadd_rn(ai.w, p.probe);
                const int nfront = build_entries(p, atoms, ai, __fmul_rn(r, r), __fmul_rn(2.0f, r), w_cand, k, w_ent);
                cnt = atom_fast(p, w_ent, k, nfront, reinterpret_cast<uint16_t *>(w_cand), s_pts);
                pairs += (unsigned)k;
            } else {
                cnt = stats ? atom_streaming<GlobalAtoms, uint32_t, true>(p, g, atoms, cells, cls_sorted, pos, w_ent, p.stat)
                            : atom_streaming<GlobalAtoms, uint32_t, false>(p, g, atoms, cells, cls_sorted, pos, w_ent, p.stat);
                streamed += 1;
            }
            if (lane == 0) {
                const uint32_t oi = orig[pos];
                const float area = atom_area(ai.w, p.probe, cnt, p.inv_n);
                val[oi] = area;
                if (p.out_counts) p.out_counts[a0 + oi] = (uint32_t)cnt;
                if (p.out_atom) p.out_atom[a0 + oi] = area;
            }
            __syncwarp();
        }
    }
    if (lane == 0 && p.stat) {
        if (pairs) atomicAdd(p.stat + 1, pairs);
        if (streamed) atomicAdd(p.stat + 2, streamed);
    }
}

__device__ __forceinline__ float warp_sum_range(const float *v, uint32_t b, uint32_t e) {
    float t = 0.0f;
    for (uint32_t i = b + lane_id(); i < e; i += 32) t += v[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(kFull, t, d);
    return t;
}

constexpr uint32_t kSeqSumMax = 16384;  // longer ranges are summed by a warp (order differs: <= 1e-6 relative)

// Segment / protein sums of one large structure.  Ranges up to kSeqSumMax atoms are summed sequentially in
// atom order like simd_sum (src/utils.rs:14-22); longer ones by a warp-shuffle tree.
__global__ void __launch_bounds__(256) large_sums_kernel(const KParams p, uint32_t sid, int N, const float *val,
                                                         const LargeHeader *h) {
    const bool bad = h != nullptr && h->ncell == 0 && N > 0;   // h == nullptr: stand-alone reduction of finished values
    const float qn = __int_as_float(0x7fc00000);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, nwarp = nth >> 5;
    if (bad) {
        const uint32_t a0 = p.struct_off[sid];
        for (int i = gtid; i < N; i += nth) {
            if (p.out_counts) p.out_counts[a0 + i] = 0u;
            if (p.out_atom) p.out_atom[a0 + i] = qn;
        }
    }
    if (p.seg_be && p.out_seg) {
        const uint32_t g0 = p.struct_seg_off[sid], g1 = p.struct_seg_off[sid + 1];
        for (uint32_t k = g0 + gtid; k < g1; k += nth) {
            const uint2 be = p.seg_be[k];
            if (be.y - be.x > kSeqSumMax) continue;
            float t = 0.0f;
            for (uint32_t i = be.x; i < be.y; ++i) t = __fadd_rn(t, val[i]);
            p.out_seg[k] = bad ? qn : t;
        }
        for (uint32_t k = g0 + gwarp; k < g1; k += nwarp) {
            const uint2 be = p.seg_be[k];
            if (be.y - be.x <= kSeqSumMax) continue;
            const float t = warp_sum_range(val, be.x, be.y);
            if (lane_id() == 0) p.out_seg[k] = bad ? qn : t;
        }
    }
    if (p.out_protein && blockIdx.x == 0) {
        // global total
        if (threadIdx.x < 32) {
            float t = 0.0f;
            if ((uint32_t)N <= kSeqSumMax) {
                if (threadIdx.x == 0) for (int i = 0; i < N; ++i) t = __fadd_rn(t, val[i]);
            } else {
                t = warp_sum_range(val, 0, (uint32_t)N);
            }
            if (threadIdx.x == 0) {
                p.out_protein[3 * (size_t)sid + 0] = bad ? qn : t;
                if (!p.seg_be) { p.out_protein[3 * (size_t)sid + 1] = 0.0f; p.out_protein[3 * (size_t)sid + 2] = bad ? qn : t; }
            }
        } else if (threadIdx.x == 32 && p.seg_be) {
            const uint32_t g0 = p.struct_seg_off[sid], g1 = p.struct_seg_off[sid + 1];
            float polar = 0.0f, nonpolar = 0.0f;
            for (uint32_t k = g0; k < g1; ++k) {
                const uint2 be = p.seg_be[k];
                float t = 0.0f;
                for (uint32_t i = be.x; i < be.y; ++i) t = __fadd_rn(t, val[i]);
                if (p.seg_polar && p.seg_polar[k]) polar = __fadd_rn(polar, t);
                else nonpolar = __fadd_rn(nonpolar, t);
            }
            p.out_protein[3 * (size_t)sid + 1] = bad ? qn : polar;
            p.out_protein[3 * (size_t)sid + 2] = bad ? qn : nonpolar;
        }
    }
}

// xyz[frame][atom][3] + radii[atom] -> float4 {x, y, z, r}; `phase` = index within the frame of element 0.
__global__ void __launch_bounds__(256) pack_frames_kernel(const float *__restrict__ xyz, const float *__restrict__ radii,
                                                          float4 *__restrict__ out, uint32_t n, uint32_t per_frame,
                                                          uint32_t phase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = (phase + i) % per_frame;
    out[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __ldg(radii + a));
}

inline void large_release(LargeWorkspace &w) {
    cudaFree(w.sorted); cudaFree(w.orig); cudaFree(w.cellid); cudaFree(w.rank); cudaFree(w.cls_sorted);
    cudaFree(w.cells); cudaFree(w.blocksum); cudaFree(w.val); cudaFree(w.hdr);
    w = LargeWorkspace{};
}

inline int large_reserve(LargeWorkspace &w, uint32_t n_atoms) {
    if (n_atoms <= w.cap_atoms) return 0;
    large_release(w);
    const uint64_t want_cells = std::min<uint64_t>(1ull << 26, std::max<uint64_t>(1ull << 20, 8ull * n_atoms));
    const size_t nb = (want_cells + kScanItems - 1) / kScanItems + 1;
    if (cudaMalloc(&w.sorted, (size_t)n_atoms * 16) != cudaSuccess || cudaMalloc(&w.orig, (size_t)n_atoms * 4) != cudaSuccess ||
        cudaMalloc(&w.cellid, (size_t)n_atoms * 4) != cudaSuccess || cudaMalloc(&w.rank, (size_t)n_atoms * 4) != cudaSuccess ||
        cudaMalloc(&w.cls_sorted, (size_t)n_atoms * 4) != cudaSuccess || cudaMalloc(&w.val, (size_t)n_atoms * 4) != cudaSuccess ||
        cudaMalloc(&w.cells, (want_cells + 2) * 4) != cudaSuccess || cudaMalloc(&w.blocksum, nb * 4) != cudaSuccess ||
        cudaMalloc(&w.hdr, sizeof(LargeHeader)) != cudaSuccess) {
        large_release(w);
        return 3;  // SASA_B200_ERR_OUT_OF_MEMORY
    }
    w.cap_atoms = n_atoms;
    w.cap_cells = (uint32_t)want_cells;
    return 0;
}

// Enqueue the whole pipeline for each large structure of one launch group.  `order` / `off` are host arrays.
// range_n > 1 selects the atom-range split (BASELINE cfg5): the cell list is built for the whole structure, but
// only slice `range_rank` of `range_n` of the cell-sorted atom order is evaluated, and no sums are produced.
inline int large_enqueue(int sm_count, LargeWorkspace &w, const KParams &kp, const uint32_t *order, uint32_t n_work,
                         const uint32_t *off, cudaStream_t st, uint32_t *launches, uint32_t range_rank = 0,
                         uint32_t range_n = 1) {
    for (uint32_t q = 0; q < n_work; ++q) {
        const uint32_t sid = order[q];
        const uint32_t a0 = off[sid];
        const int N = (int)(off[sid + 1] - a0);
        if ((uint32_t)N > w.cap_atoms) return 5;
        const float4 *at = kp.xyzr + a0;
        const uint32_t *cls = kp.cls ? kp.cls + a0 : nullptr;
        const int gb = std::min((N + 255) / 256, sm_count * 8);
        const int cell_blocks = (int)((w.cap_cells + kScanItems - 1) / kScanItems);
        const uint32_t lo = (uint32_t)((uint64_t)N * range_rank / range_n), hi = (uint32_t)((uint64_t)N * (range_rank + 1) / range_n);
        large_init_kernel<<<1, 32, 0, st>>>(w.hdr, 0u);
        large_bounds_kernel<<<gb, 256, 0, st>>>(at, N, w.hdr);
        large_grid_kernel<<<1, 32, 0, st>>>(w.hdr, kp.probe, w.cap_cells, kp.err_flag);
        large_zero_kernel<<<sm_count * 4, 256, 0, st>>>(w.hdr, w.cells);
        large_count_kernel<<<gb, 256, 0, st>>>(at, N, w.hdr, w.cells, w.cellid, w.rank);
        large_scan1_kernel<<<cell_blocks, 256, 0, st>>>(w.hdr, w.cells, w.blocksum);
        large_scan2_kernel<<<1, 1024, 0, st>>>(w.hdr, w.blocksum);
        large_scan3_kernel<<<cell_blocks, 256, 0, st>>>(w.hdr, w.cells, w.blocksum, (uint32_t)N);
        large_scatter_kernel<<<gb, 256, 0, st>>>(at, cls, N, w.hdr, w.cells, w.cellid, w.rank, w.sorted, w.orig, w.cls_sorted);
        if (range_n > 1) {
            large_range_kernel<<<1, 32, 0, st>>>(w.hdr, w.cells, lo, hi, (uint32_t)N);
            ++*launches;
        }
        const int ga = std::max(1, std::min((int)(hi - lo + 7) / 8, sm_count * 2));
        large_atoms_kernel<<<ga, 256, 0, st>>>(kp, range_n > 1 ? -1 : N, a0, w.hdr, w.sorted, w.orig, cls ? w.cls_sorted : nullptr,
                                               w.cells, w.val);
        *launches += 10;
        // level sums; also blanks the outputs of a structure with non-finite input (kp carries no segments and
        // no protein output in the atom-range split, where the per-atom values of other ranks are missing)
        large_sums_kernel<<<sm_count, 256, 0, st>>>(kp, sid, N, w.val, w.hdr);
        ++*launches;
        if (cudaGetLastError() != cudaSuccess) return 2;
    }
    return 0;
}

}  // namespace sasa
