// solve.cuh -- K3: exact transportation solver, one warp per document pair.
//
// Replaces pyemd's emd_hat_gd_metric<double> from the quantisation of the costs onward
// (SURVEY.md 8(c) S4, S6(c)-(f)):  Cn = 1e6 / maxC, iC = floor(D * Cn + 0.5), the integer
// minimum of sum f_ij * iC_ij over flows that ship the lighter side completely (the heavier
// side's surplus leaves through a zero-cost dummy column -- value-equivalent to pyemd's
// threshold node), and the un-normalisation  dist = opt / PQn / Cn + (maxSum - minSum) * maxC.
//
// pyemd runs successive shortest paths over adjacency lists; that shape is wrong for a warp.
// Here the residual problem (m supplying rows x nc columns) lives in shared memory as dense
// int32 cost and flow matrices and the warp runs a primal-dual (Hungarian-style) method:
// rows are admitted one at a time, each Dijkstra runs over the dense bipartite residual graph
// with one column (or KC columns) per lane -- relax a row = one shared-memory read per lane,
// pick the next column = one REDUX.MIN + one ballot.  All quantities are integers below 2^31,
// so the optimum is exact and equals the reference's regardless of the pivot path.
#pragma once
#include "common.cuh"

namespace wmd {

// pairs a warp claims per atomic: problem sizes vary from 1 x 1 to 17 x 17, and with 8 pairs per claim a
// third of the warps sat idle through the tail of every launch (K3 9.0 -> 7.1 ms per 1 M pairs)
constexpr unsigned kSolveClaim = 1;

struct SolveArgs {
    DocSide s1, s2;
    int64_t p0;
    int32_t npairs;
    int32_t cls;                      // class this launch serves (kClsA also finalises kClsNone pairs with status 0)
    int32_t mr, mc;                   // row / column capacity of the per-warp matrices
    int32_t ldc;                      // column pitch (odd)
    int32_t use_global;               // matrices in global scratch (class C)
    int32_t *scratch;                 // [warps, 2 * mr * ldc] when use_global
    const int32_t *ip1, *ip2;
    const int32_t *u12, *meta;
    const double *pqn, *extra;
    const float *tiles;
    int64_t tile_stride;
    const float *maxc;
    unsigned int *counter;            // work-claim counter for this launch (zeroed by the host)
    double *out;
    int32_t *status;
};

__host__ __device__ inline size_t solve_smem_per_warp(int mr, int mc, int ldc, bool use_global)
{
    size_t ints = (size_t)mr * 4 /* u, rowdist, rowpred, supply */ + (size_t)mr /* ridx */ + 2 * (size_t)mc /* cidx, way */;
    if (!use_global) ints += 2 * (size_t)mr * ldc;
    return ints * 4;
}

template <int KC>
__device__ __forceinline__ void relax_row(const int *cost, int ldc, int nc, int row, int di, int ui, int lane,
                                          const int (&v)[KC], unsigned used, int (&minv)[KC], int (&way)[KC])
{
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = lane + 32 * k;
        if (c < nc && !((used >> k) & 1u)) {
            const int cand = di + cost[row * ldc + c] - ui - v[k];
            if (cand < minv[k]) { minv[k] = cand; way[k] = row; }
        }
    }
}

// Exact min-cost of shipping supply[] (rows) into deficit[] (columns; sum equal). Returns sum f*c.
template <int KC>
__device__ long long transport_solve(int m, int nc, int ldc, const int *cost, int *flow,
                                     int *su, int *srowdist, int *srowpred, const int *ssupply, int *sway,
                                     int (&deficit)[KC], int lane)
{
    int v[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) v[k] = 0;
    for (int i = lane; i < m; i += kWarp) su[i] = 0;
    for (int x = lane; x < m * ldc; x += kWarp) flow[x] = 0;
    __syncwarp();

    for (int r = 0; r < m; ++r) {
        int sup = ssupply[r];
        while (sup > 0) {
            int minv[KC], way[KC];
            unsigned used = 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) { minv[k] = kIntInf; way[k] = -1; }
            for (int i = lane; i < m; i += kWarp) srowdist[i] = (i == r) ? 0 : -1;
            __syncwarp();
            relax_row<KC>(cost, ldc, nc, r, 0, su[r], lane, v, used, minv, way);
            int D, j0, def;
            for (;;) {
                int best = kIntInf, bk = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (!((used >> k) & 1u) && minv[k] < best) { best = minv[k]; bk = k; }
                const int delta = __reduce_min_sync(kFull, best);
                if (delta >= kIntInf) return -1;                 // unbalanced input: cannot happen, never spin
                const int jl = __ffs(__ballot_sync(kFull, best == delta)) - 1;
                const int jk = __shfl_sync(kFull, bk, jl);
                int mydef = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k) if (k == jk) mydef = deficit[k];
                def = __shfl_sync(kFull, mydef, jl);
                if (lane == jl) used |= 1u << jk;
                j0 = jl + 32 * jk;
                D = delta;
                if (def > 0) break;
                // column j0 is saturated: every row shipping into it joins the tree at distance delta
                for (int base = 0; base < m; base += kWarp) {
                    const int i = base + lane;
                    bool isnew = false;
                    if (i < m && srowdist[i] < 0 && flow[i * ldc + j0] > 0) {
                        isnew = true; srowdist[i] = delta; srowpred[i] = j0;
                    }
                    unsigned mask = __ballot_sync(kFull, isnew);
                    while (mask) {
                        const int b = __ffs(mask) - 1;
                        mask &= mask - 1;
                        relax_row<KC>(cost, ldc, nc, base + b, delta, su[base + b], lane, v, used, minv, way);
                    }
                }
            }
            // dual update: tree nodes move by (D - their distance); others keep their potentials
            for (int i = lane; i < m; i += kWarp) {
                const int dd = srowdist[i];
                if (dd >= 0) su[i] += D - dd;
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if ((used >> k) & 1u) v[k] -= D - minv[k];
                const int c = lane + 32 * k;
                if (c < nc) sway[c] = way[k];
            }
            __syncwarp();
            // augment along the tree path j0 -> ... -> r
            int amt = min(sup, def);
            for (int j = j0;;) {
                const int i = sway[j];
                if (i == r) break;
                const int jp = srowpred[i];
                amt = min(amt, flow[i * ldc + jp]);
                j = jp;
            }
            __syncwarp();
            if (lane == 0) {
                for (int j = j0;;) {
                    const int i = sway[j];
                    flow[i * ldc + j] += amt;
                    if (i == r) break;
                    const int jp = srowpred[i];
                    flow[i * ldc + jp] -= amt;
                    j = jp;
                }
            }
            sup -= amt;
#pragma unroll
            for (int k = 0; k < KC; ++k) if (lane + 32 * k == j0) deficit[k] -= amt;
            __syncwarp();
        }
    }
    long long tot = 0;
    for (int x = lane; x < m * ldc; x += kWarp) {
        const int c = x % ldc;
        if (c < nc) tot += (long long)flow[x] * (long long)cost[x];
    }
    return warp_sum_ll(tot);
}


// ------------------------------------------------------------------------------------------------
// Class A (m <= 32 rows, nc <= 32 columns incl. the dummy): the same primal-dual method with the
// whole dual / tree state in registers.  Lane L is both row L and column L:
//   as a column: potential v, remaining deficit, tentative distance minv, tree predecessor way
//   as a row:    potential u, tree distance rdist, predecessor column rpred
// The tree and the used-column set are warp-uniform bitmasks, and cmask[j] (shared memory) is the
// bitmask of rows currently shipping into column j, so "which rows join the tree when column j
// saturates" is one broadcast load instead of a scan of the flow matrix.  An augmenting path is
// walked ONCE with shuffles, hop k landing in lane k; bottleneck (REDUX.MIN) and push then run in
// parallel over the hops.  Only the dense int32 cost and flow matrices and cmask live in shared memory.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t solve_small_smem_per_warp(int mr, int mc, int ldc)
{
    return ((size_t)2 * mr * ldc + mr + mc + 32) * 4;
}

__device__ __forceinline__ long long transport_solve_small(int m, int nc, int ldc, const int *cost, int *flow, unsigned *cmask,
                                                           int supply, int deficit, int lane)
{
    int u = 0, v = 0;
    for (int x = lane; x < m * ldc; x += kWarp) flow[x] = 0;
    cmask[lane] = 0;
    __syncwarp();
    const bool iscol = lane < nc;
    for (int r = 0; r < m; ++r) {
        int sup = __shfl_sync(kFull, supply, r);
        while (sup > 0) {
            unsigned used = 0, tree = 1u << r;
            int rdist = 0, rpred = -1;
            int minv = kIntInf, way = r;
            {
                const int ur = __shfl_sync(kFull, u, r);
                if (iscol) minv = cost[r * ldc + lane] - ur - v;
            }
            int delta, jl, def;
            for (;;) {
                const int key = ((used >> lane) & 1u) ? kIntInf : minv;
                delta = __reduce_min_sync(kFull, key);
                if (delta >= kIntInf) return -1;                 // unbalanced input: cannot happen, never spin
                jl = __ffs(__ballot_sync(kFull, key == delta)) - 1;
                used |= 1u << jl;
                def = __shfl_sync(kFull, deficit, jl);
                if (def > 0) {
                    // A deficit column reached straight from the root row: ship along the single arc (r, jl)
                    // and, if the row still holds supply, KEEP the search going -- only a forward arc out of
                    // the root changed, every label stays a feasible potential, and jl (now saturated) is
                    // expanded like any other saturated column.  Restarting here re-selected all the columns
                    // this row had already filled.  Longer paths change reverse arcs the tree rests on: those
                    // leave the loop and restart the search after the augmentation.
                    if (__shfl_sync(kFull, way, jl) != r) break;
                    const int amt = min(sup, def);
                    if (lane == jl) { flow[r * ldc + jl] += amt; cmask[jl] |= 1u << r; deficit -= amt; }
                    __syncwarp();
                    sup -= amt;
                    def -= amt;
                    if (sup == 0) break;                         // def may be > 0: the standard end of a search
                }
                unsigned nr = cmask[jl] & ~tree;                 // rows shipping into the saturated column
                tree |= nr;
                if ((nr >> lane) & 1u) { rdist = delta; rpred = jl; }
                while (nr) {
                    const int i = __ffs(nr) - 1;
                    nr &= nr - 1;
                    const int ui = __shfl_sync(kFull, u, i);
                    if (iscol && !((used >> lane) & 1u)) {
                        const int cand = delta + cost[i * ldc + lane] - ui - v;
                        if (cand < minv) { minv = cand; way = i; }
                    }
                }
            }
            if ((tree >> lane) & 1u) u += delta - rdist;         // dual update (tree nodes only)
            if ((used >> lane) & 1u) v -= delta - minv;
            if (sup == 0) break;                                 // the row emptied on a direct arc
            // tree path jl -> ... -> r, walked once: hop k = (row pi ships into column pj, and stops shipping
            // amt into its tree predecessor column pjp) lands in lane k
            int pi = 0, pj = 0, pjp = -1, nh = 0;
            for (int j = jl;; ++nh) {
                const int i = __shfl_sync(kFull, way, j);
                const int jp = __shfl_sync(kFull, rpred, i);     // -1 for the root row
                if (lane == nh) { pi = i; pj = j; pjp = i == r ? -1 : jp; }
                if (i == r) { ++nh; break; }
                j = jp;
            }
            const bool hop = lane < nh;
            const int frev = (hop && pjp >= 0) ? flow[pi * ldc + pjp] : kIntInf;
            const int amt = min(min(sup, def), __reduce_min_sync(kFull, frev));
            if (hop) {
                flow[pi * ldc + pj] += amt;
                atomicOr(&cmask[pj], 1u << pi);
                if (pjp >= 0) {
                    flow[pi * ldc + pjp] = frev - amt;
                    if (frev == amt) atomicAnd(&cmask[pjp], ~(1u << pi));
                }
            }
            __syncwarp();
            sup -= amt;
            if (lane == jl) deficit -= amt;
        }
    }
    long long tot = 0;
    if (iscol)
        for (int i = 0; i < m; ++i) tot += (long long)flow[i * ldc + lane] * (long long)cost[i * ldc + lane];
    return warp_sum_ll(tot);
}

__global__ void __launch_bounds__(256)
emd_solve_small_kernel(const __grid_constant__ SolveArgs A)
{
    extern __shared__ __align__(16) int smem_i[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ldc = A.ldc;
    int *cost = smem_i + (size_t)wib * (solve_small_smem_per_warp(A.mr, A.mc, ldc) / 4);
    int *flow = cost + A.mr * ldc;
    int *sridx = flow + A.mr * ldc;
    int *scidx = sridx + A.mr;
    unsigned *cmask = reinterpret_cast<unsigned *>(scidx + A.mc);
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);

    for (;;) {
        int q0 = 0;
        if (lane == 0) q0 = (int)atomicAdd(A.counter, kSolveClaim);
        q0 = __shfl_sync(kFull, q0, 0);
        if (q0 >= A.npairs) break;
        const int q1 = min(A.npairs, q0 + (int)kSolveClaim);
        for (int q = q0; q < q1; ++q) {
            const int meta = A.meta[q];
            if ((meta & 7) != kClsA) continue;
            const int64_t p = A.p0 + q;
            const float maxc_f = A.maxc[q];
            if (!(maxc_f > 0.f)) {                               // S4: all-zero distance matrix
                if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; }
                continue;
            }
            const int uu = A.u12[q];
            const int u1 = uu & 0xffff, u2 = uu >> 16;
            const bool swap = (meta & kMetaSwap) != 0;
            int64_t a1, a2; int l;
            doc_span(A.s1, p, a1, l); doc_span(A.s2, p, a2, l);
            const int32_t *ipR = swap ? A.ip2 + slot_off(A.s2, tok2, q, a2) : A.ip1 + slot_off(A.s1, tok1, q, a1);   // supplying side
            const int32_t *ipC = swap ? A.ip1 + slot_off(A.s1, tok1, q, a1) : A.ip2 + slot_off(A.s2, tok2, q, a2);
            const int uR = swap ? u2 : u1, uC = swap ? u1 : u2;
            int m = 0, n = 0, sumR = 0, sumC = 0;
            for (int base = 0; base < uR; base += kWarp) {       // compact the residual rows: (mass << 8) | index
                const int i = base + lane;
                const int x = i < uR ? ipR[i] : 0;
                const unsigned bal = __ballot_sync(kFull, x > 0);
                if (x > 0) sridx[m + __popc(bal & ((1u << lane) - 1))] = (x << 8) | i;
                m += __popc(bal);
                sumR += x;
            }
            for (int base = 0; base < uC; base += kWarp) {
                const int j = base + lane;
                const int x = j < uC ? ipC[j] : 0;
                const unsigned bal = __ballot_sync(kFull, x > 0);
                if (x > 0) scidx[n + __popc(bal & ((1u << lane) - 1))] = (x << 8) | j;
                n += __popc(bal);
                sumC += x;
            }
            sumR = warp_sum(sumR); sumC = warp_sum(sumC);
            __syncwarp();
            const double Cn = __ddiv_rn(1000000.0, (double)maxc_f);
            long long opt = 0;
            if (n > 0 && m > 0) {
                const int diff = sumR - sumC;                    // >= 0 by the choice of the supplying side
                const int nc = n + (diff > 0 ? 1 : 0);
                const int packedR = lane < m ? sridx[lane] : 0;
                const int packedC = lane < n ? scidx[lane] : 0;
                const int supply = packedR >> 8;
                const int deficit = lane < n ? (packedC >> 8) : (lane == n ? diff : 0);
                // quantised costs of the residual sub-tile (S6(d)); the dummy column costs 0
                const float *tile = A.tiles + (int64_t)q * A.tile_stride;
                const int j = packedC & 0xff;
                for (int rI = 0; rI < m; ++rI) {
                    const int i = __shfl_sync(kFull, packedR, rI) & 0xff;
                    int ic = 0;
                    if (lane < n) {
                        const float dv = swap ? tile[j * u2 + i] : tile[i * u2 + j];
                        ic = (int)floor(__dadd_rn(__dmul_rn((double)dv, Cn), 0.5));
                    }
                    if (lane < nc) cost[rI * ldc + lane] = ic;
                }
                __syncwarp();
                opt = transport_solve_small(m, nc, ldc, cost, flow, cmask, supply, deficit, lane);
            }
            if (lane == 0) {
                double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
                dist = __ddiv_rn(dist, A.pqn[q]);                 // S6(f)
                dist = __ddiv_rn(dist, Cn);
                dist = __dadd_rn(dist, __dmul_rn(A.extra[q], (double)maxc_f));
                A.out[p] = dist;
            }
            __syncwarp();
        }
    }
}

template <int KC>
__global__ void __launch_bounds__(256)
emd_solve_kernel(const __grid_constant__ SolveArgs A)
{
    extern __shared__ __align__(16) int smem_i[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = solve_smem_per_warp(A.mr, A.mc, A.ldc, A.use_global) / 4;
    int *sb = smem_i + (size_t)wib * per_warp;
    int *su = sb, *srowdist = sb + A.mr, *srowpred = sb + 2 * A.mr, *ssupply = sb + 3 * A.mr;
    int *sridx = sb + 4 * A.mr, *scidx = sb + 5 * A.mr, *sway = scidx + A.mc;
    int *cost, *flow;
    if (A.use_global) {
        cost = A.scratch + ((size_t)blockIdx.x * wpb + wib) * 2 * (size_t)A.mr * A.ldc;
        flow = cost + (size_t)A.mr * A.ldc;
    } else {
        cost = sway + A.mc;
        flow = cost + (size_t)A.mr * A.ldc;
    }
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);

    for (;;) {
        int q0 = 0;
        if (lane == 0) q0 = (int)atomicAdd(A.counter, kSolveClaim);
        q0 = __shfl_sync(kFull, q0, 0);
        if (q0 >= A.npairs) break;
        const int q1 = min(A.npairs, q0 + (int)kSolveClaim);
        for (int q = q0; q < q1; ++q) {
            const int meta = A.meta[q];
            const int cls = meta & 7;
            const int64_t p = A.p0 + q;
            if (cls == kClsNone) continue;                       // early-out already written by K1
            if (cls != A.cls) continue;
            const float maxc_f = A.maxc[q];
            if (!(maxc_f > 0.f)) {                               // S4: all-zero distance matrix
                if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; }
                continue;
            }
            const int u = A.u12[q];
            const int u1 = u & 0xffff, u2 = u >> 16;
            const bool swap = (meta & kMetaSwap) != 0;
            int64_t a1, a2; int l;
            doc_span(A.s1, p, a1, l); doc_span(A.s2, p, a2, l);
            const int32_t *ipR = swap ? A.ip2 + slot_off(A.s2, tok2, q, a2) : A.ip1 + slot_off(A.s1, tok1, q, a1);   // supplying side
            const int32_t *ipC = swap ? A.ip1 + slot_off(A.s1, tok1, q, a1) : A.ip2 + slot_off(A.s2, tok2, q, a2);
            const int uR = swap ? u2 : u1, uC = swap ? u1 : u2;
            // compact the residual rows / columns
            int m = 0, n = 0, sumR = 0, sumC = 0;
            for (int base = 0; base < uR; base += kWarp) {
                const int i = base + lane;
                const int x = i < uR ? ipR[i] : 0;
                const unsigned bal = __ballot_sync(kFull, x > 0);
                if (x > 0) { const int pos = m + __popc(bal & ((1u << lane) - 1)); sridx[pos] = i; ssupply[pos] = x; }
                m += __popc(bal);
                sumR += x;
            }
            int deficit[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) deficit[k] = 0;
            for (int base = 0; base < uC; base += kWarp) {
                const int j = base + lane;
                const int x = j < uC ? ipC[j] : 0;
                const unsigned bal = __ballot_sync(kFull, x > 0);
                if (x > 0) { const int pos = n + __popc(bal & ((1u << lane) - 1)); scidx[pos] = j; sway[pos] = x; /* staging */ }
                n += __popc(bal);
                sumC += x;
            }
            sumR = warp_sum(sumR); sumC = warp_sum(sumC);
            __syncwarp();
            long long opt = 0;
            if (n > 0 && m > 0) {
                const int diff = sumR - sumC;                    // >= 0 by the choice of the supplying side
                const int nc = n + (diff > 0 ? 1 : 0);
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int c = lane + 32 * k;
                    deficit[k] = c < n ? sway[c] : (c == n && diff > 0 ? diff : 0);
                }
                __syncwarp();
                // quantised costs of the residual sub-tile (S6(d)); dummy column costs 0
                const double Cn = __ddiv_rn(1000000.0, (double)maxc_f);
                const float *tile = A.tiles + (int64_t)q * A.tile_stride;
                for (int rI = 0; rI < m; ++rI) {
                    const int i = sridx[rI];
                    for (int c = lane; c < nc; c += kWarp) {
                        int ic = 0;
                        if (c < n) {
                            const int j = scidx[c];
                            const float dv = swap ? tile[(int64_t)j * u2 + i] : tile[(int64_t)i * u2 + j];
                            ic = (int)floor(__dadd_rn(__dmul_rn((double)dv, Cn), 0.5));
                        }
                        cost[rI * A.ldc + c] = ic;
                    }
                }
                __syncwarp();
                opt = transport_solve<KC>(m, nc, A.ldc, cost, flow, su, srowdist, srowpred, ssupply, sway, deficit, lane);
            }
            if (lane == 0) {
                const double Cn = __ddiv_rn(1000000.0, (double)maxc_f);
                double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
                dist = __ddiv_rn(dist, A.pqn[q]);                 // S6(f)
                dist = __ddiv_rn(dist, Cn);
                dist = __dadd_rn(dist, __dmul_rn(A.extra[q], (double)maxc_f));
                A.out[p] = dist;
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// WMD_MODE_EXACT: the real-valued transportation optimum in FP64 -- no 1e6 grid, no cancellation
// (an additive mode: the reference's pyemd never computes it).  Rows = the unique
// tokens of doc1 with the nBOW weights count/len as supplies, columns = doc2's; costs are the float32
// distances widened to double.  Same primal-dual method as transport_solve<KC> on doubles; masses
// below kExactTol are treated as shipped (the two weight vectors sum to 1 only up to rounding).
// One warp per pair; cost / flow matrices in L2-resident global scratch, duals and tree in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr double kExactTol = 1e-13;

__device__ __forceinline__ double warp_min_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(kFull, v, o); v = w < v ? w : v; }
    return v;
}

struct ExactArgs {
    DocSide s1, s2;
    int64_t p0;
    int32_t npairs;
    int32_t mr, mc;                   // capacity of the per-warp arrays: rows, columns
    int32_t ldc;
    const int32_t *u12;
    const double *wt1, *wt2;          // nBOW weights at the pairs' work slots (K1, exact flag)
    const float *tiles;
    int64_t tile_stride;
    const float *maxc;
    double *scratch;                  // [warps, 2 * mr * ldc]
    unsigned int *counter;
    double *out;
    int32_t *status;
};

__host__ __device__ inline size_t exact_smem_per_warp(int mr, int mc) { return (size_t)mr * (8 + 8 + 8 + 4) + (size_t)mc * 4 + 8; }

template <int KC>
__device__ double transport_solve_f64(int m, int nc, int ldc, const double *cost, double *flow, double *su, double *srowdist,
                                      int *srowpred, const double *ssupply, int *sway, double (&deficit)[KC], int lane)
{
    const double kBig = 1e300;
    double v[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) v[k] = 0.0;
    for (int i = lane; i < m; i += kWarp) su[i] = 0.0;
    for (int x = lane; x < m * ldc; x += kWarp) flow[x] = 0.0;
    __syncwarp();
    for (int r = 0; r < m; ++r) {
        double sup = ssupply[r];
        while (sup > kExactTol) {
            double minv[KC]; int way[KC];
            unsigned used = 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) { minv[k] = kBig; way[k] = -1; }
            for (int i = lane; i < m; i += kWarp) srowdist[i] = (i == r) ? 0.0 : -1.0;
            __syncwarp();
            {
                const double ur = su[r];
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int c = lane + 32 * k;
                    if (c < nc) { minv[k] = cost[r * ldc + c] - ur - v[k]; way[k] = r; }
                }
            }
            double D = 0.0, def = 0.0; int j0 = 0;
            for (;;) {
                double best = kBig; int bk = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (!((used >> k) & 1u) && minv[k] < best) { best = minv[k]; bk = k; }
                const double delta = warp_min_f64(best);
                if (!(delta < kBig)) return -1.0;                // every column used: the remaining supply is rounding noise
                const int jl = __ffs(__ballot_sync(kFull, best == delta)) - 1;
                const int jk = __shfl_sync(kFull, bk, jl);
                double mydef = 0.0;
#pragma unroll
                for (int k = 0; k < KC; ++k) if (k == jk) mydef = deficit[k];
                def = __shfl_sync(kFull, mydef, jl);
                if (lane == jl) used |= 1u << jk;
                j0 = jl + 32 * jk;
                D = delta;
                if (def > kExactTol) break;
                for (int base = 0; base < m; base += kWarp) {    // rows shipping into the saturated column join the tree
                    const int i = base + lane;
                    bool isnew = false;
                    if (i < m && srowdist[i] < 0.0 && flow[i * ldc + j0] > kExactTol) { isnew = true; srowdist[i] = delta; srowpred[i] = j0; }
                    unsigned mask = __ballot_sync(kFull, isnew);
                    while (mask) {
                        const int row = base + __ffs(mask) - 1;
                        mask &= mask - 1;
                        const double ui = su[row];
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const int c = lane + 32 * k;
                            if (c < nc && !((used >> k) & 1u)) {
                                const double cand = delta + cost[row * ldc + c] - ui - v[k];
                                if (cand < minv[k]) { minv[k] = cand; way[k] = row; }
                            }
                        }
                    }
                }
            }
            for (int i = lane; i < m; i += kWarp) { const double dd = srowdist[i]; if (dd >= 0.0) su[i] += D - dd; }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if ((used >> k) & 1u) v[k] -= D - minv[k];
                const int c = lane + 32 * k;
                if (c < nc) sway[c] = way[k];
            }
            __syncwarp();
            double amt = sup < def ? sup : def;
            for (int j = j0;;) {
                const int i = sway[j];
                if (i == r) break;
                const int jp = srowpred[i];
                const double f = flow[i * ldc + jp];
                amt = f < amt ? f : amt;
                j = jp;
            }
            __syncwarp();
            if (lane == 0) {
                for (int j = j0;;) {
                    const int i = sway[j];
                    flow[i * ldc + j] += amt;
                    if (i == r) break;
                    const int jp = srowpred[i];
                    flow[i * ldc + jp] -= amt;
                    j = jp;
                }
            }
            sup -= amt;
#pragma unroll
            for (int k = 0; k < KC; ++k) if (lane + 32 * k == j0) deficit[k] -= amt;
            __syncwarp();
            if (!(amt > 0.0)) break;                             // a zero-flow tree arc: nothing more to ship from this row
        }
    }
    double tot = 0.0;
    for (int x = lane; x < m * ldc; x += kWarp) {
        const int c = x % ldc;
        if (c < nc) tot += flow[x] * cost[x];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
    return tot;
}

template <int KC>
__global__ void __launch_bounds__(128)
emd_solve_exact_kernel(const __grid_constant__ ExactArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_x[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char *sb = smem_x + (size_t)wib * ((exact_smem_per_warp(A.mr, A.mc) + 15) & ~(size_t)15);
    double *su = reinterpret_cast<double *>(sb), *srowdist = su + A.mr, *ssupply = srowdist + A.mr;
    int *srowpred = reinterpret_cast<int *>(ssupply + A.mr), *sway = srowpred + A.mr;
    double *cost = A.scratch + ((size_t)blockIdx.x * wpb + wib) * 2 * (size_t)A.mr * A.ldc;
    double *flow = cost + (size_t)A.mr * A.ldc;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    for (;;) {
        int q0 = 0;
        if (lane == 0) q0 = (int)atomicAdd(A.counter, kSolveClaim);
        q0 = __shfl_sync(kFull, q0, 0);
        if (q0 >= A.npairs) break;
        const int q1 = min(A.npairs, q0 + (int)kSolveClaim);
        for (int q = q0; q < q1; ++q) {
            const int uu = A.u12[q];
            const int u1 = uu & 0xffff, u2 = uu >> 16;
            if (u1 == 0 || u2 == 0) continue;                    // early-out already written by K1
            const int64_t p = A.p0 + q;
            const float maxc_f = A.maxc[q];
            if (!(maxc_f > 0.f)) {                               // S4: all-zero distance matrix
                if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; }
                continue;
            }
            int64_t a1, a2; int l;
            doc_span(A.s1, p, a1, l); doc_span(A.s2, p, a2, l);
            const double *w1 = A.wt1 + slot_off(A.s1, tok1, q, a1), *w2 = A.wt2 + slot_off(A.s2, tok2, q, a2);
            for (int i = lane; i < u1; i += kWarp) ssupply[i] = w1[i];
            double deficit[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) { const int c = lane + 32 * k; deficit[k] = c < u2 ? w2[c] : 0.0; }
            const float *tile = A.tiles + (int64_t)q * A.tile_stride;
            for (int x = lane; x < u1 * u2; x += kWarp) { const int i = x / u2, j = x - i * u2; cost[i * A.ldc + j] = (double)tile[x]; }
            __syncwarp();
            const double opt = transport_solve_f64<KC>(u1, u2, A.ldc, cost, flow, su, srowdist, srowpred, ssupply, sway, deficit, lane);
            if (lane == 0) A.out[p] = opt < 0.0 ? __longlong_as_double(0x7ff8000000000000LL) : opt;
            __syncwarp();
        }
    }
}

}  // namespace wmd
