// emd.cuh -- batched pyemd.emd(first_histogram, second_histogram, distance_matrix, extra_mass_penalty)
// on small histograms, one warp per problem.
//
// Stands behind the reference's direct pyemd call, evaluate/auto/transfer_intensity.py:8-11
// (calculate_emd: two class-probability vectors and an all-ones matrix), and is the general form of
// the arithmetic the WMD path specialises: emd_hat_gd_metric<double>
//   (a) cancel the common mass bin by bin, (b) sumP / sumQ sequentially over the ORIGINAL histograms and
//   maxC over the whole matrix, (c) PQn = 1e6 / max(sumP, sumQ), Cn = 1e6 / maxC, (d) floor(x * n + 0.5)
//   quantisation, (e) exact integer transportation -- the heavier side supplies, its surplus leaves at
//   zero cost (the dummy column), and as upstream the matrix is read as C[supplier bin][consumer bin]
//   even when the histograms swap roles -- (f) opt / PQn / Cn + (maxSum - minSum) * penalty, where
//   penalty = maxC when extra_mass_penalty == -1.
// Bit-faithful to oracle/emd_hat.c:emd_hat_gd_metric_double (tests/test_gpu_emd.py).
#pragma once
#include "common.cuh"
#include "solve.cuh"

namespace wmd {

constexpr int kEmdMaxBins = 31;          // columns incl. the dummy must fit the 32 lanes of the register-resident solver

struct EmdArgs {
    const double *P, *Q;                 // [nprob, n]
    const double *D;                     // [n, n] (shared) or [nprob, n, n]
    int64_t nprob;
    int32_t n;
    int32_t shared_d;
    double extra_mass_penalty;
    double *out;                         // [nprob]
};

__host__ __device__ inline size_t emd_smem_per_warp(int n) { return ((size_t)2 * n * (n + 1) + 3 * 32) * 4; }

__global__ void __launch_bounds__(128)
emd_hat_batch_kernel(const __grid_constant__ EmdArgs A)
{
    extern __shared__ __align__(16) int smem_i[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n = A.n, ldc = n + 1;
    int *cost = smem_i + (size_t)wib * (emd_smem_per_warp(n) / 4);
    int *flow = cost + n * ldc;
    int *sridx = flow + n * ldc;
    int *scidx = sridx + 32;
    unsigned *cmask = reinterpret_cast<unsigned *>(scidx + 32);
    for (int64_t pr = (int64_t)blockIdx.x * wpb + wib; pr < A.nprob; pr += (int64_t)gridDim.x * wpb) {
        const double *D = A.D + (A.shared_d ? 0 : pr * (int64_t)n * n);
        const double p = lane < n ? A.P[pr * n + lane] : 0.0;
        const double q = lane < n ? A.Q[pr * n + lane] : 0.0;
        // (a)
        const double rp = p < q ? 0.0 : __dsub_rn(p, q);
        const double rq = p < q ? __dsub_rn(q, p) : 0.0;
        // (b)
        double sumP = 0.0, sumQ = 0.0;
        for (int i = 0; i < n; ++i) {
            sumP = __dadd_rn(sumP, __shfl_sync(kFull, p, i));
            sumQ = __dadd_rn(sumQ, __shfl_sync(kFull, q, i));
        }
        double maxC = D[0];
        if (lane < n)
            for (int j = 0; j < n; ++j) { const double c = D[lane * n + j]; if (c > maxC) maxC = c; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double c = __shfl_xor_sync(kFull, maxC, o); if (c > maxC) maxC = c; }
        const double minSum = sumP < sumQ ? sumP : sumQ;
        const double maxSum = sumP < sumQ ? sumQ : sumP;
        // (c), (d)
        const double PQn = __ddiv_rn(1000000.0, maxSum);
        const double Cn = __ddiv_rn(1000000.0, maxC);
        const int iP = (int)floor(__dadd_rn(__dmul_rn(rp, PQn), 0.5));
        const int iQ = (int)floor(__dadd_rn(__dmul_rn(rq, PQn), 0.5));
        const int sP = warp_sum(iP), sQ = warp_sum(iQ);
        const bool swap = sQ > sP;
        const int iS = swap ? iQ : iP, iD = swap ? iP : iQ;                 // supplier / consumer masses of this bin
        const unsigned balS = __ballot_sync(kFull, iS > 0), balD = __ballot_sync(kFull, iD > 0);
        const int m = __popc(balS), nn = __popc(balD);
        if (iS > 0) sridx[__popc(balS & ((1u << lane) - 1))] = (iS << 5) | lane;
        if (iD > 0) scidx[__popc(balD & ((1u << lane) - 1))] = (iD << 5) | lane;
        __syncwarp();
        long long opt = 0;
        if (m > 0 && nn > 0) {
            const int diff = (swap ? sQ - sP : sP - sQ);
            const int nc = nn + (diff > 0 ? 1 : 0);
            const int packedR = lane < m ? sridx[lane] : 0;
            const int packedC = lane < nn ? scidx[lane] : 0;
            const int supply = packedR >> 5;
            const int deficit = lane < nn ? (packedC >> 5) : (lane == nn ? diff : 0);
            const int cbin = packedC & 31;
            for (int rI = 0; rI < m; ++rI) {
                const int rbin = __shfl_sync(kFull, packedR, rI) & 31;
                int ic = 0;
                if (lane < nn) ic = (int)floor(__dadd_rn(__dmul_rn(D[rbin * n + cbin], Cn), 0.5));
                if (lane < nc) cost[rI * ldc + lane] = ic;
            }
            __syncwarp();
            opt = transport_solve_small(m, nc, ldc, cost, flow, cmask, supply, deficit, lane);
        }
        if (lane == 0) {
            double dist = opt < 0 ? __longlong_as_double(0x7ff8000000000000LL) : (double)opt;
            dist = __ddiv_rn(dist, PQn);                                   // (f)
            dist = __ddiv_rn(dist, Cn);
            const double pen = A.extra_mass_penalty == -1.0 ? maxC : A.extra_mass_penalty;
            dist = __dadd_rn(dist, __dmul_rn(__dsub_rn(maxSum, minSum), pen));
            A.out[pr] = dist;
        }
        __syncwarp();
    }
}

}  // namespace wmd
