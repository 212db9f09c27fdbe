// nbow.cuh -- K1: nBOW builder + emd_hat pre-processing, one warp per document pair.
//
// Replaces (per pair) gensim's OOV filter, Dictionary/doc2bow and nbow() and the first half of
// pyemd's emd_hat_gd_metric (SURVEY.md 8(c) S1, S2, S5, S6(a)-(d) for the masses):
//   - token id -> table row (token map), OOV / pad dropped                      [S1]
//   - unique rows in canonical order (rank table or row number), int32 counts   [S2, S5]
//   - early-outs: empty side -> +inf (status 1); one-token union -> 0.0 (status 2)
//   - FP64 weights count/len, metric cancellation of shared tokens              [S6(a)]
//   - sumP / sumQ accumulated sequentially in Dictionary id order               [S6(b)]
//   - PQn = 1e6 / max(sumP, sumQ); integer masses floor(x * PQn + 0.5)          [S6(c),(d)]
//   - picks the supplying side and the solver class from the residual problem size.
// Integer outputs are bit-exact; the FP64 ones use only IEEE add/mul/div in the reference's order.
#pragma once
#include "common.cuh"

namespace wmd {

struct PairWork {
    int32_t *rows1, *cnt1, *ip1;      // per token slot of side 1 (first u1 entries of a document's slot)
    int32_t *rows2, *cnt2, *ip2;
    int32_t *u12;                     // per pair: u1 | u2 << 16
    int32_t *meta;                    // per pair: class | swap flag
    double *pqn, *extra;              // per pair: PQn, maxSum - minSum
    unsigned long long *stats;        // [6]: tokens, uniques, cells, solved pairs, max rows, max cols
    double *wt1, *wt2;                // WMD_MODE_EXACT: nBOW weights count/len per token slot (else unused)
    int32_t exact;                    // != 0: no cancellation / quantisation, weights out, every pair class A
    int32_t _pad;
    OutFan fan;                       // further copies of out / status (wmd_set_fanout; pyemd mode only)
};

// Unique in-vocabulary rows of one document, sorted by key. All lanes return (u, nvalid).
// smem (ints, each [Lp]): skey, skeyF, srowt, scntT, srow, scnt
__device__ __forceinline__ int side_unique(const DocSide &s, int64_t start, int nraw, const Vocab &vc, int lane,
                                           int *skey, int *skeyF, int *srowt, int *scntT, int *srow, int *scnt,
                                           int &nvalid)
{
    int nv = 0;
    for (int t = lane; t < nraw; t += kWarp) {
        int id = s.ids[start + t];
        int row = id;
        if (s.has_pad && id == s.pad_id) row = -1;
        else if (vc.map) row = (id >= 0 && (int64_t)id < vc.nmap) ? vc.map[id] : -1;
        if (row < 0 || (int64_t)row >= vc.V) row = -1;
        int key = row < 0 ? INT_MAX : (vc.rank ? vc.rank[row] : row);
        skey[t] = key;
        srowt[t] = row;
        nv += (row >= 0);
    }
    nv = warp_sum(nv);
    __syncwarp();
    int nfirst = 0;
    for (int t = lane; t < nraw; t += kWarp) {
        int key = skey[t];
        int cnt = 0;
        bool first = key != INT_MAX;
        if (first) {
            for (int q = 0; q < nraw; ++q) {
                bool eq = skey[q] == key;
                cnt += eq;
                first = first && !(eq && q < t);
            }
        }
        skeyF[t] = first ? key : INT_MAX;
        scntT[t] = cnt;
        nfirst += first;
    }
    nfirst = warp_sum(nfirst);
    __syncwarp();
    for (int t = lane; t < nraw; t += kWarp) {
        int key = skeyF[t];
        if (key != INT_MAX) {
            int pos = 0;
            for (int q = 0; q < nraw; ++q) pos += (skeyF[q] < key);
            srow[pos] = srowt[t];
            scnt[pos] = scntT[t];
        }
    }
    __syncwarp();
    nvalid = nv;
    return nfirst;
}

// smem per warp: 12 int arrays + 2 double arrays + 2 int arrays, all [Lp]
__host__ __device__ inline size_t nbow_smem_per_warp(int Lp) { return (size_t)Lp * (14 * 4 + 2 * 8); }

__global__ void __launch_bounds__(256)
nbow_pairs_kernel(DocSide s1, DocSide s2, Vocab vc, int64_t p0, int32_t npairs, int32_t Lp,
                  PairWork w, double *out, int32_t *status, const int32_t *list, const unsigned int *nlist)
{
    // list mode (list != nullptr): the launch serves the *nlist pairs list[0..) of the chunk instead of all of them
    if (nlist) npairs = (int32_t)*nlist;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    __shared__ unsigned long long s_stats[4];
    __shared__ int s_max[2];
    if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();

    unsigned char *base = smem_raw + (size_t)wib * nbow_smem_per_warp(Lp);
    double *w1 = reinterpret_cast<double *>(base);
    double *w2 = w1 + Lp;
    int *ibase = reinterpret_cast<int *>(w2 + Lp);
    int *skey = ibase, *skeyF = ibase + Lp, *srowt = ibase + 2 * Lp, *scntT = ibase + 3 * Lp;
    int *srow1 = ibase + 4 * Lp, *scnt1 = ibase + 5 * Lp, *srow2 = ibase + 6 * Lp, *scnt2 = ibase + 7 * Lp;
    int *part1 = ibase + 8 * Lp, *part2 = ibase + 9 * Lp;

    int64_t tokbase1, tokbase2;
    { int l; doc_span(s1, p0, tokbase1, l); doc_span(s2, p0, tokbase2, l); }

    unsigned long long st_tok = 0, st_unq = 0, st_cells = 0, st_solved = 0;
    int st_mr = 0, st_mc = 0;

    for (int qq = blockIdx.x * wpb + wib; qq < npairs; qq += gridDim.x * wpb) {
        const int q = list ? list[qq] : qq;
        const int64_t p = p0 + q;
        int64_t a1, a2; int n1raw, n2raw;
        doc_span(s1, p, a1, n1raw);
        doc_span(s2, p, a2, n2raw);
        int n1, n2;
        const int u1 = side_unique(s1, a1, n1raw, vc, lane, skey, skeyF, srowt, scntT, srow1, scnt1, n1);
        const int u2 = side_unique(s2, a2, n2raw, vc, lane, skey, skeyF, srowt, scntT, srow2, scnt2, n2);
        const int64_t o1 = slot_off(s1, tokbase1, q, a1), o2 = slot_off(s2, tokbase2, q, a2);
        st_tok += n1raw + n2raw;

        int meta = kClsNone;
        if (n1 == 0 || n2 == 0) {                                        // S1
            if (lane == 0) { out[p] = __longlong_as_double(0x7ff0000000000000LL); status[p] = 1; w.u12[q] = 0; w.meta[q] = 0; fan_score(w.fan, p, __longlong_as_double(0x7ff0000000000000LL)); fan_status(w.fan, p, 1); }
            continue;
        }
        if (u1 == 1 && u2 == 1 && srow1[0] == srow2[0]) {                // S2
            if (lane == 0) { out[p] = 0.0; status[p] = 2; w.u12[q] = 0; w.meta[q] = 0; fan_score(w.fan, p, 0.0); fan_status(w.fan, p, 2); }
            continue;
        }
        // S5: nBOW weights
        const double dn1 = (double)n1, dn2 = (double)n2;
        for (int i = lane; i < u1; i += kWarp) { w1[i] = __ddiv_rn((double)scnt1[i], dn1); part1[i] = -1; }
        for (int j = lane; j < u2; j += kWarp) { w2[j] = __ddiv_rn((double)scnt2[j], dn2); part2[j] = -1; }
        __syncwarp();
        if (w.exact) {                                                   // WMD_MODE_EXACT: rows, counts and weights only
            for (int i = lane; i < u1; i += kWarp) { w.rows1[o1 + i] = srow1[i]; w.cnt1[o1 + i] = scnt1[i]; w.wt1[o1 + i] = w1[i]; }
            for (int j = lane; j < u2; j += kWarp) { w.rows2[o2 + j] = srow2[j]; w.cnt2[o2 + j] = scnt2[j]; w.wt2[o2 + j] = w2[j]; }
            if (lane == 0) { w.u12[q] = u1 | (u2 << 16); w.meta[q] = kClsA; w.pqn[q] = 1.0; w.extra[q] = 0.0; status[p] = 0; }
            st_unq += u1 + u2; st_cells += (unsigned long long)u1 * u2; st_solved += 1;
            st_mr = max(st_mr, u1); st_mc = max(st_mc, u2);
            __syncwarp();
            continue;
        }
        for (int i = lane; i < u1; i += kWarp) {
            const int r = srow1[i];
            for (int j = 0; j < u2; ++j)
                if (srow2[j] == r) { part1[i] = j; part2[j] = i; }
        }
        __syncwarp();
        // S6(b): sums over the original histograms in Dictionary id order (doc1's ids first)
        double sumP = 0.0, sumQ = 0.0;
        for (int i = 0; i < u1; ++i) {
            sumP = __dadd_rn(sumP, w1[i]);
            const int j = part1[i];
            if (j >= 0) sumQ = __dadd_rn(sumQ, w2[j]);
        }
        for (int j = 0; j < u2; ++j)
            if (part2[j] < 0) sumQ = __dadd_rn(sumQ, w2[j]);
        const double maxSum = sumP < sumQ ? sumQ : sumP;
        const double minSum = sumP < sumQ ? sumP : sumQ;
        const double PQn = __ddiv_rn(1000000.0, maxSum);                 // S6(c)
        // S6(a),(d): residual masses on the 1e6 grid
        int sP = 0, sQ = 0, nP = 0, nQ = 0;
        for (int i = lane; i < u1; i += kWarp) {
            const double P = w1[i];
            const int j = part1[i];
            const double Q = j >= 0 ? w2[j] : 0.0;
            const double res = (P < Q) ? 0.0 : __dsub_rn(P, Q);
            const int ip = (int)floor(__dadd_rn(__dmul_rn(res, PQn), 0.5));
            w.rows1[o1 + i] = srow1[i]; w.cnt1[o1 + i] = scnt1[i]; w.ip1[o1 + i] = ip;
            sP += ip; nP += (ip > 0);
        }
        for (int j = lane; j < u2; j += kWarp) {
            const double Q = w2[j];
            const int i = part2[j];
            const double P = i >= 0 ? w1[i] : 0.0;
            const double res = (P < Q) ? __dsub_rn(Q, P) : 0.0;
            const int iq = (int)floor(__dadd_rn(__dmul_rn(res, PQn), 0.5));
            w.rows2[o2 + j] = srow2[j]; w.cnt2[o2 + j] = scnt2[j]; w.ip2[o2 + j] = iq;
            sQ += iq; nQ += (iq > 0);
        }
        sP = warp_sum(sP); sQ = warp_sum(sQ); nP = warp_sum(nP); nQ = warp_sum(nQ);
        const bool swap = sQ > sP;                                       // heavier side supplies
        const int m = swap ? nQ : nP;
        const int nc = (swap ? nP : nQ) + ((sP != sQ) ? 1 : 0);
        const int cls = solver_class(m, nc);
        meta = cls | (swap ? kMetaSwap : 0) | (min(255, (m * nc) >> 8) << kMetaWorkShift);
        if (lane == 0) {
            w.u12[q] = u1 | (u2 << 16);
            w.meta[q] = meta;
            w.pqn[q] = PQn;
            w.extra[q] = __dsub_rn(maxSum, minSum);
            status[p] = 0;
            fan_status(w.fan, p, 0);
        }
        st_unq += u1 + u2; st_cells += (unsigned long long)u1 * u2; st_solved += 1;
        st_mr = max(st_mr, m); st_mc = max(st_mc, nc);
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(&s_stats[0], st_tok); atomicAdd(&s_stats[1], st_unq);
        atomicAdd(&s_stats[2], st_cells); atomicAdd(&s_stats[3], st_solved);
        atomicMax(&s_max[0], st_mr); atomicMax(&s_max[1], st_mc);
    }
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&w.stats[threadIdx.x], s_stats[threadIdx.x]);
    if (threadIdx.x == 4) atomicMax(&w.stats[4], (unsigned long long)s_max[0]);
    if (threadIdx.x == 5) atomicMax(&w.stats[5], (unsigned long long)s_max[1]);
}

// Stand-alone nBOW of documents (wmd_nbow_host): rows / counts / weights at the CSR offsets.
__global__ void __launch_bounds__(256)
nbow_docs_kernel(DocSide s, Vocab vc, int32_t ndocs, int32_t Lp,
                 int32_t *rows, int32_t *counts, double *weights, int32_t *uniq, int32_t *nval)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    int *ibase = reinterpret_cast<int *>(smem_raw + (size_t)wib * nbow_smem_per_warp(Lp));
    int *skey = ibase, *skeyF = ibase + Lp, *srowt = ibase + 2 * Lp, *scntT = ibase + 3 * Lp;
    int *srow = ibase + 4 * Lp, *scnt = ibase + 5 * Lp;
    for (int q = blockIdx.x * wpb + wib; q < ndocs; q += gridDim.x * wpb) {
        int64_t a; int nraw;
        doc_span(s, q, a, nraw);
        int n;
        const int u = side_unique(s, a, nraw, vc, lane, skey, skeyF, srowt, scntT, srow, scnt, n);
        const double dn = (double)n;
        for (int i = lane; i < u; i += kWarp) {
            rows[a + i] = srow[i]; counts[a + i] = scnt[i];
            if (weights) weights[a + i] = __ddiv_rn((double)scnt[i], dn);
        }
        if (lane == 0) { uniq[q] = u; if (nval) nval[q] = n; }
        __syncwarp();
    }
}

}  // namespace wmd
