// rwmd.cuh -- K5: Relaxed WMD lower bound from the cost tiles of K2, one warp per pair.
//
// Not in the reference (Kusner et al. 2015, "From Word Embeddings To Document Distances");
// used for pruning in all-pairs mode.  l1 = sum_i w1[i] * min_j c[i][j], l2 = sum_j w2[j] *
// min_i c[i][j] with the nBOW weights count/len, accumulated sequentially in canonical order
// in FP64 (separately rounded multiply and add); lb = max(l1, l2).  Argmins take the lowest
// index on ties and are bit-exact against oracle/wmd_oracle.py:rwmd_pair.
#pragma once
#include "common.cuh"

namespace wmd {

struct RwmdArgs {
    DocSide s1, s2;
    int64_t p0;
    int32_t npairs;
    int32_t Lp;                       // per-warp term buffer length
    const int32_t *cnt1, *cnt2;
    const int32_t *u12;
    const float *tiles;
    int64_t tile_stride;
    const int32_t *status;            // global pair indexing (p)
    double *lb, *l1, *l2;             // global pair indexing; l1/l2 may be null
    int32_t *argmin_rows, *argmin_cols;   // at the documents' own offsets (chunk-relative), may be null
    int32_t am_abs;                   // != 0: the argmin arrays are indexed by absolute CSR offsets (device entry)
    int32_t _pad;
};

__global__ void __launch_bounds__(256)
rwmd_pairs_kernel(const __grid_constant__ RwmdArgs A)
{
    extern __shared__ __align__(16) double sterm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *term = sterm + (size_t)wib * A.Lp;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    for (int q = blockIdx.x * wpb + wib; q < A.npairs; q += gridDim.x * wpb) {
        const int64_t p = A.p0 + q;
        int64_t a1, a2; int l;
        doc_span(A.s1, p, a1, l); doc_span(A.s2, p, a2, l);
        const int64_t o1 = slot_off(A.s1, tok1, q, a1), o2 = slot_off(A.s2, tok2, q, a2);
        const int64_t m1 = A.am_abs ? a1 : o1, m2 = A.am_abs ? a2 : o2;
        const int st = A.status[p];
        if (st == 1) {
            if (lane == 0) { A.lb[p] = __longlong_as_double(0x7ff0000000000000LL); if (A.l1) A.l1[p] = A.lb[p]; if (A.l2) A.l2[p] = A.lb[p]; }
            continue;
        }
        if (st == 2) {
            if (lane == 0) {
                A.lb[p] = 0.0; if (A.l1) A.l1[p] = 0.0; if (A.l2) A.l2[p] = 0.0;
                if (A.argmin_rows) A.argmin_rows[m1] = 0;
                if (A.argmin_cols) A.argmin_cols[m2] = 0;
            }
            continue;
        }
        const int u = A.u12[q];
        const int u1 = u & 0xffff, u2 = u >> 16;
        const float *tile = A.tiles + (int64_t)q * A.tile_stride;
        int n1 = 0, n2 = 0;
        for (int i = lane; i < u1; i += kWarp) n1 += A.cnt1[o1 + i];
        for (int j = lane; j < u2; j += kWarp) n2 += A.cnt2[o2 + j];
        n1 = warp_sum(n1); n2 = warp_sum(n2);
        double s1 = 0.0, s2 = 0.0;
        // rows
        for (int i = lane; i < u1; i += kWarp) {
            float best = tile[(int64_t)i * u2]; int bj = 0;
            for (int j = 1; j < u2; ++j) { const float c = tile[(int64_t)i * u2 + j]; if (c < best) { best = c; bj = j; } }
            term[i] = __dmul_rn(__ddiv_rn((double)A.cnt1[o1 + i], (double)n1), (double)best);
            if (A.argmin_rows) A.argmin_rows[m1 + i] = bj;
        }
        __syncwarp();
        for (int i = 0; i < u1; ++i) s1 = __dadd_rn(s1, term[i]);
        __syncwarp();
        // columns
        for (int j = lane; j < u2; j += kWarp) {
            float best = tile[j]; int bi = 0;
            for (int i = 1; i < u1; ++i) { const float c = tile[(int64_t)i * u2 + j]; if (c < best) { best = c; bi = i; } }
            term[j] = __dmul_rn(__ddiv_rn((double)A.cnt2[o2 + j], (double)n2), (double)best);
            if (A.argmin_cols) A.argmin_cols[m2 + j] = bi;
        }
        __syncwarp();
        for (int j = 0; j < u2; ++j) s2 = __dadd_rn(s2, term[j]);
        __syncwarp();
        if (lane == 0) {
            A.lb[p] = s1 < s2 ? s2 : s1;
            if (A.l1) A.l1[p] = s1;
            if (A.l2) A.l2[p] = s2;
        }
    }
}

}  // namespace wmd
