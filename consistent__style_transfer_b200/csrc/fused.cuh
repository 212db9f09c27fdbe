// fused.cuh -- the table-mode pair kernel: K1 -> K2 -> K3 of one document pair in ONE warp, nothing in between
// touches global memory.
//
// With the V x V word-distance table D resident (wmd_set_distance_table: built once per embedding table by the
// cost kernels themselves, so every entry is bit-identical to what they compute for a pair) a pair needs no
// embedding rows at all, and the three launches of the direct path collapse into a persistent warp-per-pair loop:
//   * nBOW (spec S1, S2, S5, S6(a)-(d)) for documents of <= 32 tokens per side, one token per lane: duplicates
//     by MATCH.ANY, canonical order by a rank count over the first occurrences (shuffles, no O(n^2) shared-memory
//     loops), the FP64 mass sums of both histograms accumulated in ONE sequential loop (even lanes: sumP, odd
//     lanes: sumQ) in Dictionary id order;
//   * the u1 x u2 cost tile gathered from D (4 B per cell) into shared memory, maxC by REDUX;
//   * quantisation of the residual sub-tile and the class-A transport solve of solve.cuh, unchanged.
// Pairs this kernel cannot take -- a document longer than 32 tokens, or a residual problem above 32 x 32 -- are
// appended to `biglist` and served by the list-mode launches of nbow_pairs_kernel and the gather-mode solvers
// (solve.cuh), which read the same table: no cost tiles exist in this mode, so nothing in a launch depends on the
// longest document of the batch (BASELINE north_star: "batch scheduler that packs pairs by length").
#pragma once
#include "common.cuh"
#include "solve.cuh"

namespace wmd {

struct FusedArgs {
    DocSide s1, s2;
    Vocab vc;
    const float *D;                   // [V, V] float32 word distances
    int64_t p0;
    int32_t npairs;
    int32_t cap;                      // row / column capacity of the per-warp matrices (<= 32)
    int32_t ldc;                      // column pitch (odd)
    int32_t _pad;
    unsigned int *counter;            // work-claim counter, zeroed by the host
    int32_t *biglist;                 // launch-local indices of the pairs left to the general path
    unsigned int *nbig;               // zeroed by the host
    unsigned long long *stats;        // [6], as PairWork::stats
    double *out;
    int32_t *status;
    OutFan fan;                       // further copies of out / status (peers of a sharded job)
};

constexpr int kFusedScratchInts = 6 * 32 + 32 + 2 * 64;      // nBOW scratch: 6 key/row/count arrays, partner map, two FP64 sequences

// both matrices are padded to a multiple of 4 ints: the FP64 scratch behind them stays 8-byte aligned
__host__ __device__ inline size_t fused_cost_ints(int cap, int ldc) { return ((size_t)cap * ldc + 3) & ~(size_t)3; }
__host__ __device__ inline size_t fused_flow_ints(int cap, int ldc)
{
    const size_t a = fused_cost_ints(cap, ldc);
    return a > (size_t)kFusedScratchInts ? a : (size_t)kFusedScratchInts;
}
// per warp: cost [cap * ldc], flow (aliased with the nBOW scratch and the float tile), cmask [32], sridx [32], scidx [32], keep [4 doubles]
__host__ __device__ inline size_t fused_smem_per_warp(int cap, int ldc)
{
    return (fused_cost_ints(cap, ldc) + fused_flow_ints(cap, ldc) + 32 + 32 + 32 + 8) * 4;
}

// Unique in-vocabulary rows of one document of <= 32 tokens, one token per lane.  On return lane i < u holds the
// i-th unique key / row / count in canonical order (INT_MAX / 0 / 0 above u).  sk / sr / sc: per-warp [32] ints.
__device__ __forceinline__ void fused_side(const DocSide &s, const Vocab &vc, int64_t a, int nraw, int lane, int *sk, int *sr, int *sc,
                                           int &key_o, int &row_o, int &cnt_o, int &u, int &nvalid)
{
    int row = -1;
    if (lane < nraw) {
        const int id = s.ids[a + lane];
        row = id;
        if (s.has_pad && id == s.pad_id) row = -1;
        else if (vc.map) row = (id >= 0 && (int64_t)id < vc.nmap) ? vc.map[id] : -1;
        if (row < 0 || (int64_t)row >= vc.V) row = -1;
    }
    const bool valid = row >= 0;
    const int key = valid ? (vc.rank ? vc.rank[row] : row) : INT_MAX;
    nvalid = __popc(__ballot_sync(kFull, valid));
    const unsigned same = __match_any_sync(kFull, key);
    const bool first = valid && (__ffs(same) - 1 == lane);
    const unsigned fm = __ballot_sync(kFull, first);
    u = __popc(fm);
    // canonical position = number of smaller unique keys: the unique keys go to shared memory unsorted (sc is free
    // until the counts land in it) and every first occurrence counts against broadcast reads
    if (first) sc[__popc(fm & ((1u << lane) - 1u))] = key;
    __syncwarp();
    int pos = 0;
    for (int j = 0; j < u; ++j) pos += (sc[j] < key);
    __syncwarp();
    if (first) { sk[pos] = key; sr[pos] = row; sc[pos] = __popc(same); }
    __syncwarp();
    const bool mine = lane < u;
    key_o = mine ? sk[lane] : INT_MAX;
    row_o = mine ? sr[lane] : 0;
    cnt_o = mine ? sc[lane] : 0;
}

// FAN: the instance that also stores into the peers' arrays (wmd_set_fanout); the single-GPU instance carries none of it
template <int MINB, bool FAN>
__global__ void __launch_bounds__(128, MINB)
wmd_fused_small_kernel(const __grid_constant__ FusedArgs A)
{
    extern __shared__ __align__(16) int smem_i[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ldc = A.ldc, cap = A.cap;
    int *cost = smem_i + (size_t)wib * (fused_smem_per_warp(cap, ldc) / 4);
    int *flow = cost + fused_cost_ints(cap, ldc);
    unsigned *cmask = reinterpret_cast<unsigned *>(flow + fused_flow_ints(cap, ldc));
    int *sridx = reinterpret_cast<int *>(cmask) + 32;
    int *scidx = sridx + 32;
    double *keep = reinterpret_cast<double *>(scidx + 32);
    // nBOW scratch and the float tile live where the flow matrix will be (dead before the solver clears it)
    int *sk1 = flow, *sr1 = flow + 32, *sc1 = flow + 64, *sk2 = flow + 96, *sr2 = flow + 128, *sc2 = flow + 160;
    int *spart2 = flow + 192;
    double *seqP = reinterpret_cast<double *>(flow + 224), *seqQ = seqP + 32;
    float *tileF = reinterpret_cast<float *>(flow);
    const unsigned lt = (1u << lane) - 1u;
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);

    unsigned long long st_tok = 0, st_unq = 0, st_cells = 0, st_solved = 0;
    int st_mr = 0, st_mc = 0;

    for (;;) {
        int q = 0;
        if (lane == 0) q = (int)atomicAdd(A.counter, 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= A.npairs) break;
        const int64_t p = A.p0 + q;
        int64_t a1, a2; int n1raw, n2raw;
        doc_span(A.s1, p, a1, n1raw);
        doc_span(A.s2, p, a2, n2raw);
        if (n1raw > 32 || n2raw > 32) {                              // long documents: the general path
            if (lane == 0) A.biglist[atomicAdd(A.nbig, 1u)] = q;
            continue;
        }
        int K1, R1, C1, u1, n1, K2, R2, C2, u2, n2;
        fused_side(A.s1, A.vc, a1, n1raw, lane, sk1, sr1, sc1, K1, R1, C1, u1, n1);
        fused_side(A.s2, A.vc, a2, n2raw, lane, sk2, sr2, sc2, K2, R2, C2, u2, n2);
        (void)K2;
        if (n1 == 0 || n2 == 0) {                                    // S1
            if (lane == 0) { A.out[p] = kInf; A.status[p] = 1; if (FAN) { fan_score(A.fan, p, kInf); fan_status(A.fan, p, 1); } }
            st_tok += n1raw + n2raw;
            continue;
        }
        if (u1 == 1 && u2 == 1 && __shfl_sync(kFull, R1, 0) == __shfl_sync(kFull, R2, 0)) {      // S2
            if (lane == 0) { A.out[p] = 0.0; A.status[p] = 2; if (FAN) { fan_score(A.fan, p, 0.0); fan_status(A.fan, p, 2); } }
            st_tok += n1raw + n2raw;
            continue;
        }
        // partners: the side-2 entry holding the same table row (metric cancellation, S6(a))
        int part1 = -1;
        for (int j = 0; j < u2; ++j) { if (sk2[j] == K1) part1 = j; }
        spart2[lane] = -1;
        __syncwarp();
        if (part1 >= 0) spart2[part1] = lane;
        __syncwarp();
        const int part2 = spart2[lane];
        // S5: nBOW weights; each lane also fetches its partner's weight
        const double w1 = lane < u1 ? __ddiv_rn((double)C1, (double)n1) : 0.0;
        const double w2 = lane < u2 ? __ddiv_rn((double)C2, (double)n2) : 0.0;
        const double wq = __shfl_sync(kFull, w2, part1 < 0 ? 0 : part1);      // side-1 lane: Q of its row (if any)
        const double wp = __shfl_sync(kFull, w1, part2 < 0 ? 0 : part2);      // side-2 lane: P of its row (if any)
        // S6(b): both sums sequentially in Dictionary id order (doc1's ids first) -- even lanes add P, odd lanes Q
        {
            const unsigned pm = __ballot_sync(kFull, part1 >= 0);
            const bool q_only = lane < u2 && part2 < 0;
            const unsigned qm = __ballot_sync(kFull, q_only);
            if (lane < u1) seqP[lane] = w1;
            if (part1 >= 0) seqQ[__popc(pm & lt)] = wq;
            if (q_only) seqQ[__popc(pm) + __popc(qm & lt)] = w2;
        }
        __syncwarp();
        double sumP, sumQ;
        {
            const double *seq = (lane & 1) ? seqQ : seqP;
            const int len = (lane & 1) ? u2 : u1;
            const int both = min(u1, u2), longest = max(u1, u2);
            double acc = 0.0;
            int k = 0;
            for (; k < both; ++k) acc = __dadd_rn(acc, seq[k]);
            for (; k < longest; ++k) { if (k < len) acc = __dadd_rn(acc, seq[k]); }
            sumP = __shfl_sync(kFull, acc, 0);
            sumQ = __shfl_sync(kFull, acc, 1);
        }
        const double maxSum = sumP < sumQ ? sumQ : sumP;
        const double minSum = sumP < sumQ ? sumP : sumQ;
        const double PQn = __ddiv_rn(1000000.0, maxSum);               // S6(c)
        // S6(a),(d): residual masses on the 1e6 grid
        int ip1 = 0, ip2 = 0;
        if (lane < u1) {
            const double Q = part1 >= 0 ? wq : 0.0;
            const double res = (w1 < Q) ? 0.0 : __dsub_rn(w1, Q);
            ip1 = (int)floor(__dadd_rn(__dmul_rn(res, PQn), 0.5));
        }
        if (lane < u2) {
            const double P = part2 >= 0 ? wp : 0.0;
            const double res = (P < w2) ? __dsub_rn(w2, P) : 0.0;
            ip2 = (int)floor(__dadd_rn(__dmul_rn(res, PQn), 0.5));
        }
        const int sP = warp_sum(ip1), sQ = warp_sum(ip2);
        const bool swap = sQ > sP;                                   // heavier side supplies
        const int xR = swap ? ip2 : ip1, xC = swap ? ip1 : ip2;      // residual masses: supplying side / other side
        const unsigned balR = __ballot_sync(kFull, xR > 0), balC = __ballot_sync(kFull, xC > 0);
        const int m = __popc(balR), n = __popc(balC);
        const int diff = swap ? sQ - sP : sP - sQ;                   // >= 0
        const int nc = n + (diff > 0 ? 1 : 0);
        if (m > 32 || nc > 32) {                                     // a class-B residual (needs a 32-token side): the general path
            if (lane == 0) A.biglist[atomicAdd(A.nbig, 1u)] = q;
            __syncwarp();
            continue;
        }
        st_tok += n1raw + n2raw; st_unq += u1 + u2; st_cells += (unsigned long long)u1 * u2; st_solved += 1;
        st_mr = max(st_mr, m); st_mc = max(st_mc, nc);
        if (lane == 0) { keep[0] = PQn; keep[1] = __dsub_rn(maxSum, minSum); }
        if (xR > 0) sridx[__popc(balR & lt)] = (xR << 8) | lane;     // compact the residual rows / columns: (mass << 8) | index
        if (xC > 0) scidx[__popc(balC & lt)] = (xC << 8) | lane;
        __syncwarp();                                                // the sequences are dead: the tile may overwrite them
        // K2: the u1 x u2 tile from the word-distance table, and its maximum (pyemd's maxC is over the FULL matrix)
        unsigned mx = 0;
        {
            const int ncell = u1 * u2;
            const float inv = 1.0f / (float)u2;
            for (int c0 = 0; c0 < ncell; c0 += kWarp) {
                const int c = c0 + lane;
                const bool live = c < ncell;
                const int cc = live ? c : 0;
                const int i = (int)(((float)cc + 0.5f) * inv);       // c / u2: exact for c < 2^16, u2 <= 256
                const int j = cc - i * u2;
                const int ri = __shfl_sync(kFull, R1, i), rj = __shfl_sync(kFull, R2, j);
                if (live) {
                    const float v = __ldg(A.D + (int64_t)ri * A.vc.V + rj);
                    tileF[c] = v;
                    mx = max(mx, __float_as_uint(v));                // distances are >= 0: uint order == float order
                }
            }
            mx = __reduce_max_sync(kFull, mx);
        }
        if (mx == 0) {                                               // S4: all-zero distance matrix
            if (lane == 0) { A.out[p] = kInf; A.status[p] = 3; if (FAN) { fan_score(A.fan, p, kInf); fan_status(A.fan, p, 3); } }
            __syncwarp();
            continue;
        }
        const float maxc_f = __uint_as_float(mx);
        __syncwarp();
        long long opt = 0;
        if (n > 0 && m > 0) {
            const double Cn = __ddiv_rn(1000000.0, (double)maxc_f);
            const int packedR = lane < m ? sridx[lane] : 0;
            const int packedC = lane < n ? scidx[lane] : 0;
            // rows = the side with more nodes (see emd_solve_small_kernel); flip: the lighter side plus the surplus as a
            // zero-cost dummy ROW are the rows, the supplying side the columns
            const bool flip = m < nc;
            const int mm = flip ? nc : m, ncc = flip ? m : nc;
            const int nrow = flip ? n : m, ncol = flip ? m : n;
            const int rowP = flip ? packedC : packedR, colP = flip ? packedR : packedC;
            const int supply = lane < nrow ? (rowP >> 8) : ((flip && lane == nrow) ? diff : 0);
            const int deficit = lane < ncol ? (colP >> 8) : ((!flip && lane == ncol) ? diff : 0);
            const bool rows_doc1 = swap == flip;
            // quantised costs of the residual sub-tile (S6(d)), 32 cells per step; the dummy row / column costs 0
            {
                const int ncell = mm * ncc;
                const float inv = 1.0f / (float)ncc;
                for (int c0 = 0; c0 < ncell; c0 += kWarp) {
                    const int c = c0 + lane;
                    const bool live = c < ncell;
                    const int cc = live ? c : 0;
                    const int rI = (int)(((float)cc + 0.5f) * inv);      // c / ncc (exact: c < 2^16)
                    const int cI = cc - rI * ncc;
                    const int ridx = __shfl_sync(kFull, rowP, rI) & 0xff, cidx = __shfl_sync(kFull, colP, cI) & 0xff;
                    int ic = 0;
                    if (live && rI < nrow && cI < ncol) {
                        const float dv = rows_doc1 ? tileF[ridx * u2 + cidx] : tileF[cidx * u2 + ridx];
                        ic = (int)floor(__dadd_rn(__dmul_rn((double)dv, Cn), 0.5));
                    }
                    if (live) cost[rI * ldc + cI] = ic;
                }
            }
            sridx[lane] = deficit;                               // the compaction list is dead: it keeps the deficits for the dual objective
            __syncwarp();
            opt = transport_solve_small(mm, ncc, ldc, cost, flow, cmask, supply, deficit, lane, sridx);
        }
        if (lane == 0) {
            const double maxc_d = (double)maxc_f;
            const double Cn = __ddiv_rn(1000000.0, maxc_d);
            double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
            dist = __ddiv_rn(dist, keep[0]);                         // S6(f)
            dist = __ddiv_rn(dist, Cn);
            dist = __dadd_rn(dist, __dmul_rn(keep[1], maxc_d));
            A.out[p] = dist;
            A.status[p] = 0;
            if (FAN) { fan_score(A.fan, p, dist); fan_status(A.fan, p, 0); }
        }
        __syncwarp();
    }
    // per-launch totals (algorithmic-bytes figure of the roofline)
    if (lane == 0) {
        atomicAdd(&A.stats[0], st_tok); atomicAdd(&A.stats[1], st_unq);
        atomicAdd(&A.stats[2], st_cells); atomicAdd(&A.stats[3], st_solved);
        atomicMax(&A.stats[4], (unsigned long long)st_mr); atomicMax(&A.stats[5], (unsigned long long)st_mc);
    }
}

}  // namespace wmd
