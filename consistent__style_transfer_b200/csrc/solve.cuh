// solve.cuh -- K3: exact transportation solver, one warp per document pair.
//
// Replaces pyemd's emd_hat_gd_metric<double> from the quantisation of the costs onward
// (SURVEY.md 8(c) S4, S6(c)-(f)):  Cn = 1e6 / maxC, iC = floor(D * Cn + 0.5), the integer
// minimum of sum f_ij * iC_ij over flows that ship the lighter side completely (the heavier
// side's surplus leaves through a zero-cost dummy column -- value-equivalent to pyemd's
// threshold node), and the un-normalisation  dist = opt / PQn / Cn + (maxSum - minSum) * maxC.
//
// pyemd runs successive shortest paths over adjacency lists; that shape is wrong for a warp.
// Here the residual problem (m supplying rows x nc columns) is a dense int32 cost matrix and the
// warp runs a primal-dual (Hungarian-style) method: rows are admitted one at a time, each Dijkstra
// runs over the dense bipartite residual graph with one column (or KC columns) per lane -- relax a
// row = one read per lane and column word, pick the next column = one REDUX.MIN + one ballot.  All quantities are integers below 2^31,
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
    int32_t _pad;
    int32_t *scratch;                 // wide classes: [warps, 2 * mr * ldc] quantised costs + flow
    const int32_t *ip1, *ip2;
    const int32_t *u12, *meta;
    const double *pqn, *extra;
    const float *tiles;
    int64_t tile_stride;
    const float *maxc;
    unsigned int *counter;            // work-claim counter for this launch (zeroed by the host)
    double *out;
    int32_t *status;
    // list mode: the launch serves the *nlist pairs list[0..) of the chunk (fused.cuh leaves them behind)
    const int32_t *list;
    const unsigned int *nlist;
    // gather mode (template flag GATHER): no cost tiles exist; costs come straight from the V x V word-distance
    // table through the pairs' unique rows (K1), and the kernel finds maxC itself (kept in maxc_w for the epilogue)
    const float *D;
    int64_t V;
    const int32_t *rows1, *rows2;
    float *maxc_w;
    OutFan fan;                       // further copies of out / status (peers of a sharded job)
};

// Reads of the word-distance table are single-use 32-byte sectors scattered over V x V floats: they bypass L1 and
// are the first lines L2 gives up, so that they do not push the solvers' cost matrices out.
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float ldg_once(const float *ptr, uint64_t policy)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(policy));
    return v;
}

// Largest entry of the u1 x u2 tile of D addressed by the unique rows r1 / r2 (pyemd's maxC is over the FULL matrix).
__device__ __forceinline__ float gather_tile_max(const float *D, int64_t V, const int32_t *r1, const int32_t *r2, int u1, int u2, int lane)
{
    unsigned mx = 0;
    const int ncell = u1 * u2;
    const float inv = 1.0f / (float)u2;
    const uint64_t once = l2_evict_first_policy();
    for (int c = lane; c < ncell; c += kWarp) {
        const int i = (int)(((float)c + 0.5f) * inv);                // c / u2: exact for c < 2^16, u2 <= 256
        const int j = c - i * u2;
        mx = max(mx, __float_as_uint(ldg_once(D + (int64_t)__ldg(r1 + i) * V + __ldg(r2 + j), once)));
    }
    return __uint_as_float(__reduce_max_sync(kFull, mx));            // distances are >= 0: uint order == float order
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

// Shared-memory accesses of the class-A solver go through 32-bit shared addresses: with generic
// pointers the compiler rebuilt the CTA's shared window base (S2UR SR_CgaCtaId + UMOV + ULEA) in front
// of every LDS of the select / expand / relax loop -- 4 % of the kernel's instructions.
__device__ __forceinline__ unsigned smem_addr(const void *p)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));                                   // opaque: kept in a register, not rematerialised
    return a;
}
__device__ __forceinline__ int lds32(unsigned a)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(unsigned a, int v)
{
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void atom_or_s32(unsigned a, unsigned v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void atom_and_s32(unsigned a, unsigned v)
{
    asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// deficit0 (optional, shared memory, [32]): the columns' deficits as they were handed in.  With it the optimum is read
// off the duals -- at termination every arc that carries flow is tight, so sum f c = sum u_r supply_r + sum v_c deficit_c
// exactly (integers) -- instead of a pass over the flow and cost matrices.
__device__ __forceinline__ long long transport_solve_small(int m, int nc, int ldc, const int *cost_p, int *flow_p, unsigned *cmask_p,
                                                           int supply, int deficit, int lane, const int *deficit0 = nullptr)
{
    const unsigned cost = smem_addr(cost_p), flow = smem_addr(flow_p), cmask = smem_addr(cmask_p);
    const unsigned ld4 = (unsigned)ldc * 4u, lane4 = (unsigned)lane * 4u;
    int u = 0, v = 0;
    for (int x = lane; x < m * ldc; x += kWarp) sts32(flow + 4u * x, 0);
    sts32(cmask + lane4, 0);
    __syncwarp();
    const bool iscol = lane < nc;
    const unsigned lbit = 1u << lane;
    // Start from reduced costs instead of zero duals (the Hungarian method's row and column reduction): u_r = the row's
    // cheapest arc, v_c = what is left of the column's cheapest.  Every row then owns a tight arc, and one greedy pass ships
    // along tight arcs into columns that still have a deficit -- any flow on tight arcs under feasible duals is a valid
    // state of the primal-dual method, and the optimum it ends in is the same integer.  On Yelp-shape pairs this replaces
    // 6.7 of 14.5 searches and 11 of 46 column selections per pair by ~25 instructions per row (tools/solver_model.py).
    int srem = supply;                                           // supply still to ship (lane = row); `supply` stays as handed in
    {
        if (lane < m) {
            int best = kIntInf;
            for (int c = 0; c < nc; ++c) best = min(best, lds32(cost + lane * ld4 + 4u * c));
            u = best;
        }
        int best = kIntInf;
        for (int r = 0; r < m; ++r) {
            const int ur = __shfl_sync(kFull, u, r);
            if (iscol) best = min(best, lds32(cost + r * ld4 + lane4) - ur);
        }
        if (iscol) v = best;
        for (int r = 0; r < m; ++r) {
            const int ur = __shfl_sync(kFull, u, r), sr = __shfl_sync(kFull, srem, r);
            const bool open = iscol && deficit > 0 && lds32(cost + r * ld4 + lane4) - ur - v == 0;
            const unsigned cand = __ballot_sync(kFull, open);
            if (cand == 0 || sr == 0) continue;
            const int j = __ffs(cand) - 1;
            const int amt = min(sr, __shfl_sync(kFull, deficit, j));
            if (lane == j) {
                sts32(flow + r * ld4 + lane4, amt);
                sts32(cmask + lane4, lds32(cmask + lane4) | (1u << r));
                deficit -= amt;
            }
            if (lane == r) srem -= amt;
        }
        __syncwarp();
    }
    for (int r = 0; r < m; ++r) {
        int sup = __shfl_sync(kFull, srem, r);
        while (sup > 0) {
            unsigned used = 0, tree = 1u << r;
            int rdist = 0, rpred = -1;
            int minv = kIntInf, way = r;
            {
                const int ur = __shfl_sync(kFull, u, r);
                if (iscol) minv = lds32(cost + r * ld4 + lane4) - ur - v;
            }
            int jl = 0, def = 0;
            int key = minv;                                      // kIntInf once the column is used
            int delta = __reduce_min_sync(kFull, key);
            // the loop condition only fails on unbalanced input (cannot happen): never spin
            while (delta < kIntInf) {
                jl = __ffs(__ballot_sync(kFull, key == delta)) - 1;
                used |= 1u << jl;
                def = __shfl_sync(kFull, deficit, jl);
                if (def > 0) {
                    // A column with a deficit: ship along the tree path and, when that leaves the tree intact,
                    // KEEP the search going.  Every label stays the exact distance in the new residual graph as
                    // long as no reverse arc of the path ran empty (the arcs the path added are tight, the ones it
                    // used are still there), so if the row still holds supply the column -- now saturated -- is
                    // expanded like any other.  Restarting instead re-selected every column the row had already
                    // been through: 51 -> 44 selections and 16 -> 14 searches per Yelp-shape pair.
                    int amt;
                    bool intact = true;
                    if (__shfl_sync(kFull, way, jl) == r) {      // the single arc (r, jl)
                        amt = min(sup, def);
                        if (lane == jl) {
                            const unsigned fa = flow + r * ld4 + lane4;
                            sts32(fa, lds32(fa) + amt);
                            sts32(cmask + lane4, lds32(cmask + lane4) | (1u << r));
                            deficit -= amt;
                        }
                    } else {
                        // tree path jl -> ... -> r, walked once: hop k = (row pi ships into column pj, and stops
                        // shipping amt into its tree predecessor column pjp) lands in lane k
                        int pi = 0, pj = 0, pjp = -1, nh = 0;
                        for (int j = jl;; ++nh) {
                            const int i = __shfl_sync(kFull, way, j);
                            const int jp = __shfl_sync(kFull, rpred, i);     // -1 for the root row
                            if (lane == nh) { pi = i; pj = j; pjp = i == r ? -1 : jp; }
                            if (i == r) { ++nh; break; }
                            j = jp;
                        }
                        const bool hop = lane < nh;
                        const unsigned frow = flow + pi * ld4;
                        const int frev = (hop && pjp >= 0) ? lds32(frow + 4u * pjp) : kIntInf;
                        const int bott = __reduce_min_sync(kFull, frev);
                        amt = min(min(sup, def), bott);
                        intact = amt < bott;                     // no reverse arc of the path runs empty
                        if (hop) {
                            sts32(frow + 4u * pj, lds32(frow + 4u * pj) + amt);
                            atom_or_s32(cmask + 4u * pj, 1u << pi);
                            if (pjp >= 0) {
                                sts32(frow + 4u * pjp, frev - amt);
                                if (frev == amt) atom_and_s32(cmask + 4u * pjp, ~(1u << pi));
                            }
                        }
                        if (lane == jl) deficit -= amt;
                    }
                    __syncwarp();
                    sup -= amt;
                    def -= amt;
                    if (sup == 0 || !intact) break;              // the row is empty, or the search has to start again
                }
                unsigned nr = (unsigned)lds32(cmask + 4u * jl) & ~tree;      // rows shipping into the saturated column
                tree |= nr;
                if (nr & lbit) { rdist = delta; rpred = jl; }
                while (nr) {
                    const int i = __ffs(nr) - 1;
                    nr &= nr - 1;
                    const int ui = __shfl_sync(kFull, u, i);
                    if (iscol && !(used & lbit)) {
                        const int cand = delta + lds32(cost + i * ld4 + lane4) - ui - v;
                        if (cand < minv) { minv = cand; way = i; }
                    }
                }
                key = (used & lbit) ? kIntInf : minv;
                delta = __reduce_min_sync(kFull, key);
            }
            asm volatile("" : "+r"(delta));                        // opaque: the error exit is tested here, once per search,
            if (delta >= kIntInf) return -1;                     // not threaded into the loop's back edge (3 BREAKs per step)
            if (tree & lbit) u += delta - rdist;                 // dual update (tree nodes only)
            if (used & lbit) v -= delta - minv;
        }
    }
    long long tot = 0;
    if (deficit0) tot = (long long)u * (long long)supply + (long long)v * (long long)deficit0[lane];
    else if (iscol)
        for (int i = 0; i < m; ++i) tot += (long long)lds32(flow + i * ld4 + lane4) * (long long)lds32(cost + i * ld4 + lane4);
    return warp_sum_ll(tot);
}

// Everything the pair's set-up needs (offsets, the cost normaliser) dies before the solver runs: the solver's own
// state fills the 48 registers that let five blocks of eight warps share an SM (at 64 registers the kernel ran 8 %
// fewer instructions 5 % slower: 47 % instead of 58 % resident warps, 75 % instead of 85 % issue slots used).
template <bool GATHER>
__global__ void __launch_bounds__(128, 9)
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

    const int npairs = A.nlist ? (int)*A.nlist : A.npairs;
    for (;;) {
        int q = 0;
        if (lane == 0) q = (int)atomicAdd(A.counter, 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= npairs) break;
        if (A.list) q = A.list[q];
        const int meta = A.meta[q];
        if ((meta & kMetaCls) != kClsA) continue;
        float maxc_f = 0.f;
        if (!GATHER) maxc_f = A.maxc[q];
        else {
            const int uu = A.u12[q];
            int64_t tok1, tok2, a1, a2; int l;
            doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l);
            doc_span(A.s1, A.p0 + q, a1, l); doc_span(A.s2, A.p0 + q, a2, l);
            maxc_f = gather_tile_max(A.D, A.V, A.rows1 + slot_off(A.s1, tok1, q, a1), A.rows2 + slot_off(A.s2, tok2, q, a2),
                                     uu & 0xffff, uu >> 16, lane);
            if (lane == 0) A.maxc_w[q] = maxc_f;
        }
        if (!(maxc_f > 0.f)) {                                   // S4: all-zero distance matrix
            if (lane == 0) { A.out[A.p0 + q] = __longlong_as_double(0x7ff0000000000000LL); A.status[A.p0 + q] = 3; fan_score(A.fan, A.p0 + q, __longlong_as_double(0x7ff0000000000000LL)); fan_status(A.fan, A.p0 + q, 3); }
            continue;
        }
        long long opt = 0;
        {
            const int uu = A.u12[q];
            const int u1 = uu & 0xffff, u2 = uu >> 16;
            const bool swap = (meta & kMetaSwap) != 0;
            int64_t tok1, tok2, a1, a2; int l;
            doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l);
            doc_span(A.s1, A.p0 + q, a1, l); doc_span(A.s2, A.p0 + q, a2, l);
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
            if (n > 0 && m > 0) {
                const double Cn = __ddiv_rn(1000000.0, (double)maxc_f);
                const int diff = sumR - sumC;                    // >= 0 by the choice of the supplying side
                const int nc = n + (diff > 0 ? 1 : 0);
                const int packedR = lane < m ? sridx[lane] : 0;
                const int packedC = lane < n ? scidx[lane] : 0;
                // The balanced problem is symmetric, so the solver's rows may be either side.  Rows = the side with
                // MORE nodes: a search then wades through fewer columns before it meets one with a deficit
                // (Yelp-shape pairs: 70 -> 50 column selections per pair; profiles/README.md).  flip: rows = the
                // lighter side plus the surplus as a zero-cost dummy ROW, columns = the supplying side.
                const bool flip = m < nc;
                const int mm = flip ? nc : m, ncc = flip ? m : nc;
                const int nrow = flip ? n : m, ncol = flip ? m : n;              // real (non-dummy) rows / columns
                const int rowP = flip ? packedC : packedR, colP = flip ? packedR : packedC;
                const int supply = lane < nrow ? (rowP >> 8) : ((flip && lane == nrow) ? diff : 0);
                const int deficit = lane < ncol ? (colP >> 8) : ((!flip && lane == ncol) ? diff : 0);
                // quantised costs of the residual sub-tile (S6(d)); the dummy row / column costs 0
                const float *tile = GATHER ? nullptr : A.tiles + (int64_t)q * A.tile_stride;
                const int cidx = colP & 0xff;
                const bool rows_doc1 = swap == flip;                             // the rows are doc1's tokens
                const int32_t *rowtab = nullptr;                                 // GATHER: table rows of the solver's rows / of this lane's column
                int64_t coloff = 0;
                if (GATHER) {
                    const int32_t *r1 = A.rows1 + slot_off(A.s1, tok1, q, a1), *r2 = A.rows2 + slot_off(A.s2, tok2, q, a2);
                    rowtab = rows_doc1 ? r1 : r2;
                    coloff = lane < ncol ? (int64_t)__ldg((rows_doc1 ? r2 : r1) + cidx) : 0;
                }
                for (int rI = 0; rI < mm; ++rI) {
                    const int ridx = __shfl_sync(kFull, rowP, rI) & 0xff;
                    int ic = 0;
                    if (lane < ncol && rI < nrow) {
                        float dv;
                        if (GATHER) dv = __ldg(A.D + (int64_t)__ldg(rowtab + ridx) * A.V + coloff);       // D is symmetric
                        else dv = rows_doc1 ? tile[ridx * u2 + cidx] : tile[cidx * u2 + ridx];
                        ic = (int)floor(__dadd_rn(__dmul_rn((double)dv, Cn), 0.5));
                    }
                    if (lane < ncc) cost[rI * ldc + lane] = ic;
                }
                sridx[lane] = deficit;                           // the compaction list is dead: it keeps the deficits for the dual objective
                __syncwarp();
                opt = transport_solve_small(mm, ncc, ldc, cost, flow, cmask, supply, deficit, lane, sridx);
            }
        }
        if (lane == 0) {
            const double maxc_d = (double)(GATHER ? A.maxc_w[q] : A.maxc[q]);
            const double Cn = __ddiv_rn(1000000.0, maxc_d);       // recomputed: nothing of the set-up stays live across the solver
            double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
            dist = __ddiv_rn(dist, A.pqn[q]);                     // S6(f)
            dist = __ddiv_rn(dist, Cn);
            dist = __dadd_rn(dist, __dmul_rn(A.extra[q], maxc_d));
            A.out[A.p0 + q] = dist;
            fan_score(A.fan, A.p0 + q, dist);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// WMD_MODE_EXACT: the real-valued transportation optimum in FP64 -- no 1e6 grid, no cancellation
// (an additive mode: the reference's pyemd never computes it).  Rows = the unique
// tokens of doc1 with the nBOW weights count/len as supplies, columns = doc2's; costs are the float32
// distances widened to double.  Same primal-dual method as transport_solve_small on doubles; masses
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
