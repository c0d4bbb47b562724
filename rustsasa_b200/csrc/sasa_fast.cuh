// sasa_fast.cuh -- the tight per-atom pipeline of the fused kernel for n_points <= 128 (the reference's
// default, 100 points, is the headline configuration): every loop below is written to compile to a
// branch-free, predicate-accumulating SASS body; the generic routines in sasa_device.cuh cover every other
// case (more points, the large-structure path, list overflow, boundary statistics).
//
// Per atom (one warp):
//   1. gather   the cell's flattened candidate list is cached in registers and shared by the atoms of a
//               cell; 32 candidates per step are distance-tested and compacted into a u16 index list
//   2. entries  (vx, vy, vz, limit) per accepted neighbour, "near" neighbours packed first
//   3. phase 1  all body points (3 slots per lane for n = 100) against the first m entries,
//               FMUL + 2 FFMA + FSETP.LT.OR per test
//   4. phase 2  the survivors against the remaining entries as a G x (32/G) tile of (survivor, entry) pairs
//   5. tail     the n mod lanes tail points (unfused dot, <=) against all entries, same tile form
#pragma once
#include "sasa_device.cuh"

namespace sasa {

// ---- 1. gather through the register cache ----------------------------------------------------------------
template <bool HAS_CLS>
__device__ __forceinline__ int fast_gather(const float4 *s_atom, const uint32_t *s_cls, int pos, const float4 ai,
                                           float reach_i, const CandCache<uint16_t> &cc, uint16_t *cand) {
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const uint32_t cls_i = HAS_CLS ? s_cls[pos] : 0u;
    int k = 0;
#pragma unroll
    for (int w = 0; w < kCacheWin; ++w) {
        if (32 * w < cc.total) {   // warp-uniform
            const int j = cc.get(w);
            const float4 aj = s_atom[j];
            const float dx = ai.x - aj.x, dy = ai.y - aj.y, dz = ai.z - aj.z;
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float cut = reach_i + aj.w;
            // membership is result-neutral (any superset of the overlapping pairs gives the same counts), so this
            // test may use contracted arithmetic; the 1e-3 A slack absorbs its rounding
            bool acc = (d2 <= cut * cut) & (lane < cc.total - 32 * w) & (j != pos);
            if (HAS_CLS) acc = acc & (s_cls[j] != cls_i);
            const unsigned m = __ballot_sync(kFull, acc);
            const int at = k + __popc(m & lt);
            if (acc & (at < kQueueCap)) cand[at] = (uint16_t)j;
            k += __popc(m);
        }
    }
    __syncwarp();
    return k;
}

// ---- 2. entries ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int fast_entries(const float4 *s_atom, const float4 ai, float probe, float r2, float two_r,
                                            float near2, const uint16_t *cand, int k, float4 *ent) {
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    int nfront = 0, nback = k - 1;
    for (int q0 = 0; q0 < k; q0 += 32) {
        const int q = q0 + lane;
        const bool valid = q < k;
        const float4 aj = s_atom[valid ? (int)cand[q] : 0];
        float vmag;
        const float4 e = make_entry(ai, aj, probe, r2, two_r, &vmag);
        const bool near = valid & (vmag < near2);
        const unsigned mn = __ballot_sync(kFull, near);
        const unsigned mv = __ballot_sync(kFull, valid);
        const unsigned mf = mv & ~mn;
        const int at = near ? nfront + __popc(mn & lt) : nback - __popc(mf & lt);
        if (valid) ent[at] = e;
        nfront += __popc(mn);
        nback -= __popc(mf);
    }
    __syncwarp();
    return nfront;
}

// ---- 3. phase 1 ------------------------------------------------------------------------------------------------
// Returns per-slot "still exposed" ballots through live[]; NSL = slots holding body points.
template <int NSL>
__device__ __forceinline__ void fast_phase1(const float4 *ent, int m, const float4 *pts, int nbody, unsigned (&live)[4]) {
    const int lane = lane_id();
    float4 p0 = pts[lane], p1, p2, p3;
    if (NSL > 1) p1 = pts[32 + lane];
    if (NSL > 2) p2 = pts[64 + lane];
    if (NSL > 3) p3 = pts[96 + lane];
    bool o0 = lane >= nbody, o1 = 32 + lane >= nbody, o2 = 64 + lane >= nbody, o3 = 96 + lane >= nbody;
#pragma unroll 2
    for (int q = 0; q < m; ++q) {
        const float4 e = ent[q];
        o0 = o0 || (dot_body(p0.x, p0.y, p0.z, e) < e.w);
        if (NSL > 1) o1 = o1 || (dot_body(p1.x, p1.y, p1.z, e) < e.w);
        if (NSL > 2) o2 = o2 || (dot_body(p2.x, p2.y, p2.z, e) < e.w);
        if (NSL > 3) o3 = o3 || (dot_body(p3.x, p3.y, p3.z, e) < e.w);
    }
    live[0] = __ballot_sync(kFull, !o0);
    live[1] = NSL > 1 ? __ballot_sync(kFull, !o1) : 0u;
    live[2] = NSL > 2 ? __ballot_sync(kFull, !o2) : 0u;
    live[3] = NSL > 3 ? __ballot_sync(kFull, !o3) : 0u;
}

// ---- 4./5. (survivor x entry) tile ---------------------------------------------------------------------------------
// ns (<= G) points listed in queue[0, ns) against entries [q0, k): lane l owns point (l mod G) and entry offset
// (l div G); one step tests 32/G entries against every point.  Returns how many points no entry occludes.
template <int G, bool TAIL>
__device__ __forceinline__ int fast_tile(const float4 *ent, int q0, int k, const float4 *pts, const uint16_t *queue, int ns) {
    constexpr int kStep = 32 / G;
    const int lane = lane_id();
    const int sidx = lane & (G - 1);
    const float4 pt = pts[sidx < ns ? (int)queue[sidx] : 0];
    bool hit = false;
    for (int q = q0 + lane / G; q < k; q += kStep) {
        const float4 e = ent[q];
        if (TAIL) hit = hit | (dot_tail(pt.x, pt.y, pt.z, e) <= e.w);
        else hit = hit | (dot_body(pt.x, pt.y, pt.z, e) < e.w);
    }
    unsigned mk = __ballot_sync(kFull, hit);
    if (G <= 16) mk |= mk >> 16;
    if (G <= 8) mk |= mk >> 8;
    if (G <= 4) mk |= mk >> 4;
    if (G <= 2) mk |= mk >> 2;
    if (G <= 1) mk |= mk >> 1;
    const unsigned valid = ns >= 32 ? 0xffffffffu : ((1u << ns) - 1u);
    return __popc(~mk & valid);
}

template <bool TAIL>
__device__ __forceinline__ int fast_tile_any(const float4 *ent, int q0, int k, const float4 *pts, const uint16_t *queue, int ns) {
    if (ns <= 4) return fast_tile<4, TAIL>(ent, q0, k, pts, queue, ns);
    if (ns <= 8) return fast_tile<8, TAIL>(ent, q0, k, pts, queue, ns);
    if (ns <= 16) return fast_tile<16, TAIL>(ent, q0, k, pts, queue, ns);
    return fast_tile<32, TAIL>(ent, q0, k, pts, queue, ns);
}

// One atom, n_points <= 128, complete neighbour list in ent[0, k) with nfront near entries first.
// nbody = min(n_points, n_body) body points; tail points are [nbody, n_points).
__device__ __forceinline__ int fast_atom(const KParams &p, const float4 *ent, int k, int nfront, const float4 *pts,
                                         uint16_t *queue, int nbody, int nsl) {
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int m = min(k, min(max(nfront, p.m_min), p.m_max));
    unsigned live[4];
    if (nsl == 3) fast_phase1<3>(ent, m, pts, nbody, live);
    else if (nsl == 4) fast_phase1<4>(ent, m, pts, nbody, live);
    else if (nsl == 2) fast_phase1<2>(ent, m, pts, nbody, live);
    else if (nsl == 1) fast_phase1<1>(ent, m, pts, nbody, live);
    else live[0] = live[1] = live[2] = live[3] = 0u;
    int exposed = 0;
    const int ns = __popc(live[0]) + __popc(live[1]) + __popc(live[2]) + __popc(live[3]);
    if (m == k) {
        exposed = ns;
    } else if (ns) {
        // survivors -> queue (slot-major)
        int at = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if ((live[s] >> lane) & 1u) queue[at + __popc(live[s] & lt)] = (uint16_t)(32 * s + lane);
            at += __popc(live[s]);
        }
        __syncwarp();
        for (int b = 0; b < ns; b += 32) exposed += fast_tile_any<false>(ent, m, k, pts, queue + b, min(32, ns - b));
        __syncwarp();
    }
    const int ntail = (int)p.n_points - nbody;
    if (ntail) {
        if (k == 0) {
            exposed += ntail;
        } else {
            if (lane < ntail) queue[lane] = (uint16_t)(nbody + lane);
            __syncwarp();
            exposed += fast_tile_any<true>(ent, 0, k, pts, queue, ntail);
            __syncwarp();
        }
    }
    return exposed;
}

}  // namespace sasa
