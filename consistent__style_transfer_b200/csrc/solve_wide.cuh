// solve_wide.cuh -- K3 for residual problems above 32 x 32: exact transportation solver, one warp per pair,
// instances by the number of 32-column words KC of the SHORTER side.
//
// Same primal-dual method as transport_solve_small (solve.cuh) and the same integer optimum as pyemd's
// emd_hat_gd_metric; what differs from the class-A kernel is where the state lives:
//   * rows = the side with MORE nodes (the balanced problem is symmetric; when the lighter side is the larger one
//     the surplus becomes a zero-cost dummy ROW), columns = the shorter side.  Everything a search does per step
//     -- pick the next column, relax a row -- walks the column words, so the instance is chosen by the shorter
//     side alone: a 180 x 40 problem runs in <2>, not in the square instance of its longer side.
//   * per-lane registers hold only the column side: potential v, tentative distance minv (lane L = columns
//     L + 32k, k < KC).  The row side -- potential u, remaining supply, tree predecessor -- the column deficits, the
//     tree predecessor of every column and cmask[j] (bit rows shipping into column j) sit in shared memory and are
//     read by broadcast; the tree set is one register per lane (lane w owns rows 32w .. 32w+31).  The row count
//     therefore costs no registers and no template parameter, and the kernels run at 16 .. 32 warps per SM where
//     register arrays for both sides (the previous version) allowed 12 .. 16.  The supplies, read once per row, live in
//     the per-warp global scratch: in shared memory they would cost the 5- and 6-word classes their sixth block per SM.
//   * a row that joins the tree at distance d gets u -= d at once and every tree row u += D when the search ends
//     at distance D (same update as u += D - d, no per-row distance array).
//   * rows that join the tree in one step are relaxed two at a time: both cost rows are requested before either
//     is compared (the matrices of the larger classes overflow L2; a step is one DRAM round trip).
// Quantised costs and the flow matrix live in per-warp global scratch, pitch 32 * KC ints so that a row word is
// one aligned 128-byte line; the flow matrix is only touched where cmask has a bit and is never cleared.
#pragma once
#include "common.cuh"
#include "solve.cuh"

namespace wmd {

// per-warp shared memory: u [mrp] ints, deficit [32 KC] ints, cmask [32 KC * krp] words, rpred [mrp] + way, cany [32 KC] shorts
// per-warp global scratch: cost, flow [mr * 32 KC] ints, srem [mrp] ints (read once per row: it would cost the 5- and
// 6-word classes their sixth block per SM in shared memory)
__host__ __device__ inline int wide_krp(int mr) { return (mr + 31) >> 5; }
template <int KC>
__host__ __device__ inline size_t solve_wide_smem_per_warp(int mr)
{
    const size_t krp = (size_t)wide_krp(mr), mrp = 32 * krp, mcp = 32 * KC;
    const size_t bytes = 4 * (mrp + mcp + mcp * krp) + 2 * (mrp + 2 * mcp);
    return (bytes + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t solve_wide_scratch_ints_per_warp(int mr, int ldc) { return (size_t)2 * mr * ldc + 32 * (size_t)wide_krp(mr); }

template <int KC>
__device__ long long transport_solve_wide(const int mm, const int ncc, const int krp, const int *__restrict__ cost, int *__restrict__ flow,
                                          int *u, int *srem, int *deficit, unsigned *cmask, short *rpred, short *way,
                                          unsigned short *cany, const int lane)
{
    constexpr int ldc = 32 * KC;
    int v[KC], minv[KC];
    unsigned usedb, invalb = 0;                                      // bit k: this lane's column of word k is used / does not exist
    const unsigned lbit = 1u << lane;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        v[k] = 0;
        if (lane + 32 * k >= ncc) invalb |= 1u << k;                 // columns >= ncc do not exist: permanently "used"
        cany[lane + 32 * k] = 0;
    }
    for (int i = lane; i < 32 * krp; i += kWarp) u[i] = 0;
    for (int x = lane; x < ncc * krp; x += kWarp) cmask[x] = 0;
    __syncwarp();

    // Reduced-cost start with one greedy pass over tight arcs (see transport_solve_small) -- for problems of up to 64 rows
    // only: the python model (tools/solver_model.py) counts 36 % fewer column selections at 64 tokens, no change at 128 and
    // 50 % MORE at 256, where the greedy shipments along near-tied arcs have to be re-routed one by one.
    if (mm <= 64) {
        for (int r = 0; r < mm; ++r) {                               // u_r = the row's cheapest arc
            int best = kIntInf;
#pragma unroll
            for (int k = 0; k < KC; ++k) if (!((invalb >> k) & 1u)) best = min(best, cost[r * ldc + lane + 32 * k]);
            best = __reduce_min_sync(kFull, best);
            if (lane == 0) u[r] = best;
        }
        __syncwarp();
        {                                                            // v_c = what is left of the column's cheapest arc
            int best[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) best[k] = kIntInf;
            for (int r = 0; r < mm; ++r) {
                const int ur = u[r];
#pragma unroll
                for (int k = 0; k < KC; ++k) best[k] = min(best[k], cost[r * ldc + lane + 32 * k] - ur);
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) if (!((invalb >> k) & 1u)) v[k] = best[k];
        }
        for (int r = 0; r < mm; ++r) {                               // every row ships along a tight arc into a column with a deficit
            const int ur = u[r], sr = srem[r];
            int jk = -1, jl = 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const bool open = !((invalb >> k) & 1u) && deficit[lane + 32 * k] > 0 && cost[r * ldc + lane + 32 * k] - ur - v[k] == 0;
                const unsigned cand = __ballot_sync(kFull, open);
                if (jk < 0 && cand) { jk = k; jl = __ffs(cand) - 1; }
            }
            if (jk < 0 || sr == 0) continue;
            const int j0 = jl + 32 * jk;
            const int amt = min(sr, deficit[j0]);
            __syncwarp();
            if (lane == 0) {
                flow[r * ldc + j0] = amt;
                cmask[j0 * krp + (r >> 5)] |= 1u << (r & 31);
                cany[j0] |= (unsigned short)(1u << (r >> 5));
                deficit[j0] -= amt;
                srem[r] = sr - amt;
            }
            __syncwarp();
        }
    }

    unsigned treew = 0;                                              // lane w: tree bits of rows 32w .. 32w + 31
    int steps = 0;                                                   // hard bound on column selections: a pair never spins
    for (int r = 0; r < mm; ++r) {
        int sup = srem[r];
        if (sup == 0) continue;
        int crow[KC];                                                // the row's own costs serve all of its searches
#pragma unroll
        for (int k = 0; k < KC; ++k) crow[k] = cost[r * ldc + lane + 32 * k];
        while (sup > 0) {
            treew = lane == (r >> 5) ? (1u << (r & 31)) : 0u;
            {
                const int ur = u[r];
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    minv[k] = ((invalb >> k) & 1u) ? kIntInf : crow[k] - ur - v[k];
                    way[lane + 32 * k] = (short)r;
                }
                usedb = invalb;
            }
            __syncwarp();
            int delta, j0;
            for (;;) {
                int best = kIntInf, bk = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int key = ((usedb >> k) & 1u) ? kIntInf : minv[k];
                    if (key < best) { best = key; bk = k; }
                }
                delta = __reduce_min_sync(kFull, best);
                if (delta >= kIntInf || ++steps > (1 << 24)) return -1;      // unbalanced input: cannot happen, never spin
                const int jl = __ffs(__ballot_sync(kFull, best == delta)) - 1;
                const int jk = KC > 1 ? __shfl_sync(kFull, bk, jl) : 0;
                j0 = jl + 32 * jk;
                if (lane == jl) usedb |= 1u << bk;
                int def = deficit[j0];
                if (def > 0) {
                    // ship along the tree path and keep searching while the tree stays intact (see transport_solve_small)
                    int amt;
                    bool intact = true;
                    if (way[j0] == r) {
                        amt = min(sup, def);
                        if (lane == 0) {
                            unsigned *cw = cmask + j0 * krp + (r >> 5);
                            const unsigned w = *cw, bit = 1u << (r & 31);
                            int *f = flow + r * ldc + j0;
                            *f = (w & bit) ? *f + amt : amt;
                            *cw = w | bit;
                            cany[j0] |= (unsigned short)(1u << (r >> 5));
                        }
                    } else {
                        // tree path j0 -> ... -> r in pieces of 32 hops (hop h = row pi starts shipping into column pj
                        // and stops shipping amt into its tree predecessor column pjp); pass 0 finds the bottleneck, the
                        // push follows at once when the path fits one piece, otherwise pass 1 walks it again
                        int bott = kIntInf;
                        amt = min(sup, def);
                        for (int pass = 0; pass < 2; ++pass) {
                            int j = j0;
                            bool done = false, single = true;
                            for (int piece = 0; !done; ++piece) {
                                if (piece > 16) return -1;           // a path has at most 2 * 257 hops: never spin
                                int pi = 0, pj = 0, pjp = -1, nh = 0;
                                for (; nh < kWarp;) {
                                    const int i = way[j];
                                    const int jp = rpred[i];
                                    if (lane == nh) { pi = i; pj = j; pjp = i == r ? -1 : jp; }
                                    ++nh;
                                    if (i == r) { done = true; break; }
                                    j = jp;
                                }
                                if (!done) single = false;
                                const bool hop = lane < nh;
                                const int frev = (hop && pjp >= 0) ? flow[pi * ldc + pjp] : kIntInf;
                                if (pass == 0) { bott = min(bott, __reduce_min_sync(kFull, frev)); amt = min(amt, bott); }
                                if ((pass == 0 && single && done) || pass == 1) {
                                    if (hop) {
                                        const unsigned bit = 1u << (pi & 31);
                                        const unsigned old = atomicOr(&cmask[pj * krp + (pi >> 5)], bit);
                                        cany[pj] |= (unsigned short)(1u << (pi >> 5));      // this hop alone owns column pj
                                        int *f = flow + pi * ldc + pj;
                                        *f = (old & bit) ? *f + amt : amt;
                                        if (pjp >= 0) {
                                            flow[pi * ldc + pjp] = frev - amt;
                                            if (frev == amt) atomicAnd(&cmask[pjp * krp + (pi >> 5)], ~bit);
                                        }
                                    }
                                    __syncwarp();
                                }
                            }
                            if (single) break;
                        }
                        intact = amt < bott;                         // no reverse arc of the path ran empty
                    }
                    if (lane == 0) deficit[j0] = def - amt;
                    __syncwarp();
                    sup -= amt;
                    def -= amt;
                    if (sup == 0 || !intact) break;                  // the row is empty, or the search has to start again
                }
                // rows shipping into the saturated column join the tree at distance delta, two at a time
                // cany[j0]: the row words of cmask[j0] that may hold a bit (set with every shipment, cleared only here)
                int pend = -1;
                for (unsigned wm = cany[j0]; wm; wm &= wm - 1) {
                    const int w = __ffs(wm) - 1;
                    const unsigned cm = cmask[j0 * krp + w];
                    if (cm == 0) { if (lane == 0) cany[j0] &= (unsigned short)~(1u << w); continue; }
                    unsigned nr = cm & ~__shfl_sync(kFull, treew, w);
                    if (lane == w) treew |= nr;
                    while (nr) {
                        const int i = 32 * w + __ffs(nr) - 1;
                        nr &= nr - 1;
                        if (pend < 0) { pend = i; continue; }
                        const int ba = delta - u[pend], bb = delta - u[i];
                        const int *ra = cost + pend * ldc + lane, *rb = cost + i * ldc + lane;
                        int ca[KC], cb[KC];
#pragma unroll
                        for (int k = 0; k < KC; ++k) { ca[k] = ra[32 * k]; cb[k] = rb[32 * k]; }
                        if (lane == 0) { u[pend] = -ba; u[i] = -bb; rpred[pend] = (short)j0; rpred[i] = (short)j0; }
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            if (!((usedb >> k) & 1u)) {
                                const int xa = ba + ca[k] - v[k], xb = bb + cb[k] - v[k];
                                if (xa < minv[k]) { minv[k] = xa; way[lane + 32 * k] = (short)pend; }
                                if (xb < minv[k]) { minv[k] = xb; way[lane + 32 * k] = (short)i; }
                            }
                        }
                        pend = -1;
                    }
                }
                if (pend >= 0) {
                    const int ba = delta - u[pend];
                    const int *ra = cost + pend * ldc + lane;
                    int ca[KC];
#pragma unroll
                    for (int k = 0; k < KC; ++k) ca[k] = ra[32 * k];
                    if (lane == 0) { u[pend] = -ba; rpred[pend] = (short)j0; }
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        if (!((usedb >> k) & 1u)) {
                            const int xa = ba + ca[k] - v[k];
                            if (xa < minv[k]) { minv[k] = xa; way[lane + 32 * k] = (short)pend; }
                        }
                    }
                }
                __syncwarp();
            }
            // dual update (tree nodes only): joined rows already carry u - (their distance); the root's distance is 0
            for (int w = 0; w < krp; ++w) {
                const unsigned t = __shfl_sync(kFull, treew, w);
                if (t & lbit) u[32 * w + lane] += delta;
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) if (((usedb & ~invalb) >> k) & 1u) v[k] -= delta - minv[k];
            __syncwarp();
        }
    }
    long long tot = 0;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = lane + 32 * k;
        if (c < ncc) {
            for (int w = 0; w < krp; ++w) {
                unsigned bits = cmask[c * krp + w];
                while (bits) {
                    const int i = 32 * w + __ffs(bits) - 1;
                    bits &= bits - 1;
                    tot += (long long)flow[i * ldc + c] * (long long)cost[i * ldc + c];
                }
            }
        }
    }
    return warp_sum_ll(tot);
}

// Longest first: the pairs of a chunk in descending order of their work estimate (meta bits 8 .. 15), one counting sort in
// a single block.  A pair of the larger classes runs for milliseconds; claimed in arrival order the longest ones may start
// last and the chunk waits for them with most of the SMs idle.
struct ListSortArgs {
    const int32_t *list;              // pairs to order (launch-local indices), or nullptr: 0 .. npairs - 1
    const unsigned int *nlist;        // its length on the device, or nullptr: npairs
    int32_t npairs;
    const int32_t *meta;
    int32_t *sorted;
};
__global__ void __launch_bounds__(1024)
list_sort_kernel(const __grid_constant__ ListSortArgs A)
{
    __shared__ unsigned int hist[256], cur[256];
    const int n = A.nlist ? (int)*A.nlist : A.npairs;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int q = A.list ? A.list[i] : i;
        atomicAdd(&hist[(A.meta[q] >> kMetaWorkShift) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int k = 255; k >= 0; --k) { cur[k] = run; run += hist[k]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int q = A.list ? A.list[i] : i;
        A.sorted[atomicAdd(&cur[(A.meta[q] >> kMetaWorkShift) & 255], 1u)] = q;
    }
}

template <int KC, bool GATHER, int MINB>
__global__ void __launch_bounds__(128, MINB)
emd_solve_wide_kernel(const __grid_constant__ SolveArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_w[];
    constexpr int ldc = 32 * KC;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int krp = wide_krp(A.mr), mrp = 32 * krp;
    unsigned char *sb = smem_w + (size_t)wib * solve_wide_smem_per_warp<KC>(A.mr);
    int *u = reinterpret_cast<int *>(sb), *deficit = u + mrp;
    unsigned *cmask = reinterpret_cast<unsigned *>(deficit + ldc);
    short *rpred = reinterpret_cast<short *>(cmask + (size_t)ldc * krp), *way = rpred + mrp;
    unsigned short *cany = reinterpret_cast<unsigned short *>(way + ldc);
    int *listR = u, *listC = reinterpret_cast<int *>(cmask);         // compaction lists: dead before the solver clears u / cmask
    int *cost, *flow, *srem;
    {
        const size_t w = (size_t)blockIdx.x * wpb + wib, mat = (size_t)A.mr * ldc;
        cost = A.scratch + w * solve_wide_scratch_ints_per_warp(A.mr, ldc); flow = cost + mat; srem = flow + mat;
    }
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    const int npairs = A.nlist ? (int)*A.nlist : A.npairs;

    for (;;) {
        int q = 0;
        if (lane == 0) q = (int)atomicAdd(A.counter, 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= npairs) break;
        if (A.list) q = A.list[q];
        const int meta = A.meta[q];
        if ((meta & kMetaCls) != A.cls) continue;
        const int64_t p = A.p0 + q;
        const int uu = A.u12[q];
        const int u1 = uu & 0xffff, u2 = uu >> 16;
        const bool swap = (meta & kMetaSwap) != 0;
        int64_t a1, a2; int l;
        doc_span(A.s1, p, a1, l); doc_span(A.s2, p, a2, l);
        const int64_t o1 = slot_off(A.s1, tok1, q, a1), o2 = slot_off(A.s2, tok2, q, a2);
        const int32_t *r1 = nullptr, *r2 = nullptr;
        if (GATHER) { r1 = A.rows1 + o1; r2 = A.rows2 + o2; }
        float maxc_f = GATHER ? 0.f : A.maxc[q];
        if (!GATHER && !(maxc_f > 0.f)) {                            // S4: all-zero distance matrix
            if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; fan_score(A.fan, p, kInf); fan_status(A.fan, p, 3); }
            continue;
        }
        const int32_t *ipR = swap ? A.ip2 + o2 : A.ip1 + o1;         // supplying side
        const int32_t *ipC = swap ? A.ip1 + o1 : A.ip2 + o2;
        const int uR = swap ? u2 : u1, uC = swap ? u1 : u2;
        // Compact the residual nodes to the front of the lists as (mass << 8) | index; nodes without residual mass (the
        // metric cancellation emptied them) go to the back, index only: pyemd's maxC is over the FULL tile.
        int m = 0, n = 0, zR = 0, zC = 0, sumR = 0, sumC = 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int base = 0; base < uR; base += kWarp) {
            const int i = base + lane;
            const int x = i < uR ? ipR[i] : 0;
            const unsigned bal = __ballot_sync(kFull, x > 0), zal = __ballot_sync(kFull, i < uR && x <= 0);
            if (x > 0) listR[m + __popc(bal & lt)] = (x << 8) | i;
            else if (i < uR) listR[mrp - 1 - (zR + __popc(zal & lt))] = i;
            m += __popc(bal); zR += __popc(zal);
            sumR += x;
        }
        for (int base = 0; base < uC; base += kWarp) {
            const int j = base + lane;
            const int x = j < uC ? ipC[j] : 0;
            const unsigned bal = __ballot_sync(kFull, x > 0), zal = __ballot_sync(kFull, j < uC && x <= 0);
            if (x > 0) listC[n + __popc(bal & lt)] = (x << 8) | j;
            else if (j < uC) listC[mrp - 1 - (zC + __popc(zal & lt))] = j;
            n += __popc(bal); zC += __popc(zal);
            sumC += x;
        }
        sumR = warp_sum(sumR); sumC = warp_sum(sumC);
        __syncwarp();
        double Cn = 0.0;
        long long opt = 0;
        bool zero_matrix = false;
        {
            const int diff = sumR - sumC;                            // >= 0 by the choice of the supplying side
            const int nc = n + (diff > 0 ? 1 : 0);
            // rows = the side with more nodes; flip: the lighter side plus the surplus as a zero-cost dummy ROW are the
            // rows, the supplying side the columns
            const bool flip = m < nc;
            const int mm = flip ? nc : m, ncc = flip ? m : nc;
            const int nrow = flip ? n : m, ncol = flip ? m : n;      // real (non-dummy) rows / columns
            const int zrow = flip ? zC : zR, zcol = flip ? zR : zC;  // nodes without residual mass on either side
            const int *rowL = flip ? listC : listR, *colL = flip ? listR : listC;
            const bool rows_doc1 = swap == flip;                     // the rows are doc1's tokens
            int cj[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                const int pc = c < ncol ? colL[c] : 0;
                cj[k] = pc & 0xff;
                if (GATHER) cj[k] = c < ncol ? __ldg((rows_doc1 ? r2 : r1) + cj[k]) : 0;      // table row of the column's token
                deficit[c] = c < ncol ? pc >> 8 : ((!flip && c == ncol) ? diff : 0);
            }
            if (GATHER) {
                // One pass over the table: every cell of the full tile is fetched once -- the residual sub-tile lands in
                // the cost matrix as float bits and is quantised in place once maxC is known.
                const uint64_t once = l2_evict_first_policy();
                const int32_t *rtab = rows_doc1 ? r1 : r2, *ctab = rows_doc1 ? r2 : r1;
                unsigned mx = 0;
                const int T = nrow + zrow;
                int *rowtab = srem;                                  // table row of every token of the row document (srem is still free)
                for (int t = lane; t < T; t += kWarp) rowtab[t] = __ldg(rtab + (t < nrow ? rowL[t] & 0xff : rowL[mrp - 1 - (t - nrow)]));
                __syncwarp();
                constexpr int RB = KC <= 2 ? 4 : 2;                  // rows in flight: RB * KC independent gathers per lane
                for (int t0 = 0; t0 < T; t0 += RB) {                 // all tokens of the row document x residual columns
                    unsigned bits[RB][KC];
#pragma unroll
                    for (int b = 0; b < RB; ++b) {
                        const float *drow = A.D + (int64_t)rowtab[min(t0 + b, T - 1)] * A.V;     // D is symmetric
#pragma unroll
                        for (int k = 0; k < KC; ++k) bits[b][k] = lane + 32 * k < ncol ? __float_as_uint(ldg_once(drow + cj[k], once)) : 0u;
                    }
#pragma unroll
                    for (int b = 0; b < RB; ++b) {
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            mx = max(mx, bits[b][k]);                // distances are >= 0: uint order == float order
                            if (t0 + b < nrow && lane + 32 * k < ncol) cost[(t0 + b) * ldc + lane + 32 * k] = (int)bits[b][k];
                        }
                    }
                }
                for (int x = lane; x < T * zcol; x += kWarp) {                      // ... x columns without residual mass
                    const int t = x / zcol, z = x - t * zcol;
                    const int j = colL[mrp - 1 - z];
                    mx = max(mx, __float_as_uint(ldg_once(A.D + (int64_t)rowtab[t] * A.V + __ldg(ctab + j), once)));
                }
                mx = __reduce_max_sync(kFull, mx);
                maxc_f = __uint_as_float(mx);
                zero_matrix = mx == 0;
            }
            if (!zero_matrix) Cn = __ddiv_rn(1000000.0, (double)maxc_f);
            if (!zero_matrix && n > 0 && m > 0) {
                // quantised costs of the residual sub-tile (S6(d)); the dummy row / column costs 0
                const float *tile = GATHER ? nullptr : A.tiles + (int64_t)q * A.tile_stride;
                __syncwarp();
                constexpr int QB = KC <= 2 ? 8 : KC <= 4 ? 4 : 2;    // rows in flight: the values come back from L2 or DRAM
                for (int r0 = 0; r0 < mm; r0 += QB) {
                    float dv[QB][KC];
#pragma unroll
                    for (int b = 0; b < QB; ++b) {
                        const int rI = min(r0 + b, mm - 1);
                        const int i = rI < nrow ? rowL[rI] & 0xff : 0;
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const int c = lane + 32 * k;
                            dv[b][k] = 0.f;
                            if (c < ncol && rI < nrow) {
                                if (GATHER) dv[b][k] = __int_as_float(cost[rI * ldc + c]);
                                else dv[b][k] = rows_doc1 ? tile[(int64_t)i * u2 + cj[k]] : tile[(int64_t)cj[k] * u2 + i];
                            }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < QB; ++b) {
                        const int rI = r0 + b;
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const int c = lane + 32 * k;
                            int ic = 0;
                            if (c < ncol && rI < nrow) ic = (int)floor(__dadd_rn(__dmul_rn((double)dv[b][k], Cn), 0.5));
                            if (c < ncc && rI < mm) cost[rI * ldc + c] = ic;
                        }
                    }
                }
                // supplies last: srem does not alias the lists, u and cmask (which do) are cleared by the solver
                for (int i = lane; i < mm; i += kWarp) srem[i] = i < nrow ? rowL[i] >> 8 : diff;
                __syncwarp();
                opt = transport_solve_wide<KC>(mm, ncc, krp, cost, flow, u, srem, deficit, cmask, rpred, way, cany, lane);
            }
        }
        if (zero_matrix) {                                           // S4: all-zero distance matrix
            if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; fan_score(A.fan, p, kInf); fan_status(A.fan, p, 3); }
            __syncwarp();
            continue;
        }
        if (lane == 0) {
            double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
            dist = __ddiv_rn(dist, A.pqn[q]);                         // S6(f)
            dist = __ddiv_rn(dist, Cn);
            dist = __dadd_rn(dist, __dmul_rn(A.extra[q], (double)maxc_f));
            A.out[p] = dist;
            fan_score(A.fan, p, dist);
        }
        __syncwarp();
    }
}

}  // namespace wmd
