// allpairs.cuh -- all-pairs top-k WMD with RWMD pruning (BASELINE config 4; not in the reference).
//
// For every document i of set A, the k documents j of set B with the smallest WMD(i, j)
// (ties: lower j), exact -- the values are the same pyemd-grid optima the pair path returns.
// Prefetch-and-prune after Kusner et al. 2015, with the relaxed bound evaluated the linear-complexity
// way of Atasu et al. 2017 (LC-RWMD) so that the 10^10 bounds of a 100k x 100k problem cost ~10 FMAs each:
//   D[a][b]   float32 distance between every two table rows, computed by the K2 cost kernels
//             themselves on blocks of consecutive rows (bit-identical to the pair path's cost tiles);
//   Z_B[w][j] = min over the tokens t of B_j of D[w][t]        (one pass over B, V x |B| floats)
//   L1(i, j)  = sum over tokens w of A_i of weight(w) * Z_B[w][j]      (cost of moving A_i's mass greedily)
//   L2(i, j)  = sum over tokens t of B_j of weight(t) * Z_A[t][i]
//   LB(i, j)  = max(L1, L2) <= WMD_real(i, j)
// Round 1 solves exactly the k candidates with the smallest LB per row; their k-th distance thr_i
// bounds the answer, so round 2 solves every remaining j with LB(i, j) <= thr_i + kLbMargin and
// nothing else can enter the top k.  kLbMargin covers the gap between the real-valued optimum that
// RWMD bounds and the 1e-6-grid optimum pyemd returns: costs round by <= 0.5e-6 * maxC per unit of
// mass and the <= 512 masses by <= 0.5e-6 each, i.e. |WMD_pyemd - WMD_real| < 3e-4 * maxC / 2 ... we
// use 1e-3, an absolute bound that is generous for unit vectors (maxC <= 2) and costs nothing.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace wmd {

// Relative slack of the round-2 threshold: thr = kth * (1 + kLbMarginRel) + kLbMarginRel * dmax.  It covers the gap
// between the real-valued optimum that RWMD bounds and the 1e6-grid optimum pyemd returns: every cost is rounded by
// <= 0.5e-6 * maxC and every one of the <= 2 * 256 masses by <= 0.5e-6 of the total, so
// |WMD_pyemd - WMD_real| <= (0.5e-6 + 512 * 0.5e-6) * maxC < 2.6e-4 * maxC, and maxC <= dmax, the largest entry of
// the word-distance table (so the margin follows the scale of an un-normalised embedding table).
constexpr double kLbMarginRel = 3e-4;

// ---- D: scatter the cost tiles of row-block pairs into the V x V table --------------------------
// pair q = (bi, bj), bi <= bj, blocks of BS consecutive table rows.
__global__ void dtab_make_pairs_kernel(int32_t V, int32_t BS, int32_t nb, int64_t q0, int32_t npairs,
                                       int32_t *rows1, int32_t *rows2, int32_t *u12, int32_t *bij)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npairs) return;
    // unrank (bi, bj) from the global pair number over the upper triangle, row-major
    int64_t g = q0 + q;
    int bi = 0;
    {
        // row bi holds nb - bi pairs; solve by float estimate then fix up
        const double nbd = (double)nb;
        double est = nbd + 0.5 - sqrt((nbd + 0.5) * (nbd + 0.5) - 2.0 * (double)g);
        bi = (int)est;
        if (bi < 0) bi = 0;
        if (bi >= nb) bi = nb - 1;
        while (bi > 0 && (int64_t)bi * nb - (int64_t)bi * (bi - 1) / 2 > g) --bi;
        while ((int64_t)(bi + 1) * nb - (int64_t)(bi + 1) * bi / 2 <= g) ++bi;
    }
    const int64_t before = (int64_t)bi * nb - (int64_t)bi * (bi - 1) / 2;
    const int bj = bi + (int)(g - before);
    const int n1 = min(BS, V - bi * BS), n2 = min(BS, V - bj * BS);
    for (int k = 0; k < BS; ++k) {
        rows1[(int64_t)q * BS + k] = k < n1 ? bi * BS + k : 0;
        rows2[(int64_t)q * BS + k] = k < n2 ? bj * BS + k : 0;
    }
    u12[q] = n1 | (n2 << 16);
    bij[2 * q] = bi; bij[2 * q + 1] = bj;
}

__global__ void dtab_scatter_kernel(int32_t V, int32_t BS, int32_t npairs, const int32_t *u12, const int32_t *bij,
                                    const float *tiles, int64_t tile_stride, float *D)
{
    const int q = blockIdx.x;
    if (q >= npairs) return;
    const int u = u12[q], n1 = u & 0xffff, n2 = u >> 16;
    const int bi = bij[2 * q], bj = bij[2 * q + 1];
    const float *t = tiles + (int64_t)q * tile_stride;
    for (int c = threadIdx.x; c < n1 * n2; c += blockDim.x) {
        const int i = c / n2, j = c - i * n2;
        const float v = t[c];
        const int64_t a = (int64_t)bi * BS + i, b = (int64_t)bj * BS + j;
        D[a * V + b] = v;
        D[b * V + a] = v;
    }
}

__global__ void table_max_kernel(const float *D, int64_t n, unsigned int *maxbits)
{
    unsigned m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(D[i]));                  // distances are >= 0: uint order == float order
    m = __reduce_max_sync(kFull, m);
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits, m);
}

// ---- D16: the word-distance table rounded DOWN to half precision ----------------------------------
// The bounds are sums of table entries, so entries rounded down keep every bound a lower bound; half the bytes
// of the float table for the two kernels whose time is its traffic (z_build16_kernel, and through Z the bound tiles).
__global__ void dtab_to_half_kernel(const float *D, int64_t n, __half *D16)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        D16[i] = __float2half_rd(D[i]);
}

// ---- packed document lists for the bound kernel: (table row, weight) per unique token --------------
// weight = count / valid tokens rounded DOWN to float32, so that every partial sum of the bound stays below the
// exact one.  Entries live at the documents' CSR offsets like rows / counts.
__global__ void pack_lists_kernel(const int32_t *rows, const int32_t *cnt, const int64_t *off, const int32_t *uniq,
                                  const int32_t *nval, int64_t ndocs, int2 *lists)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= ndocs) return;
    const int64_t a = off[j];
    const int u = uniq[j];
    const float n = (float)nval[j];
    for (int k = 0; k < u; ++k)
        lists[a + k] = make_int2(rows[a + k], __float_as_int(__fdiv_rd((float)cnt[a + k], n)));
}

// ---- Z[w][j] = min_t D16[t][w] over the in-vocabulary rows t of document j -----------------------
// rows / uniq: per-document unique rows at the CSR offsets (nbow_docs_kernel). Documents without
// any in-vocabulary token get +inf (every WMD with them is +inf).
struct ZArgs {
    const __half *D16; int32_t V;
    const int32_t *rows; const int64_t *off; const int32_t *uniq;
    int64_t doc0; int32_t ndocs;              // documents [doc0, doc0 + ndocs) -> columns [0, ndocs)
    __half *Z; int64_t ldz;                   // Z[w * ldz + (j - doc0)]
};

// block = 256 words x 32 documents: warp y owns one document, its lanes eight consecutive words each (one 16-byte load
// per token: 512 B of a table row per warp); the document's row ids are fetched once, 32 per lane, and handed round by
// shuffles, four table reads in flight per lane.  ncu (profiles/README.md): the kernel is bound by instruction issue,
// not by the L2 -- eight words per load instead of four halved its instructions; staging the table's hot rows or whole
// column strips in shared memory (both tried) only moved the limit to the shared-memory pipe.
__device__ __forceinline__ uint4 hmin8(uint4 a, uint4 b)
{
    uint4 r;
    *reinterpret_cast<__half2 *>(&r.x) = __hmin2(*reinterpret_cast<const __half2 *>(&a.x), *reinterpret_cast<const __half2 *>(&b.x));
    *reinterpret_cast<__half2 *>(&r.y) = __hmin2(*reinterpret_cast<const __half2 *>(&a.y), *reinterpret_cast<const __half2 *>(&b.y));
    *reinterpret_cast<__half2 *>(&r.z) = __hmin2(*reinterpret_cast<const __half2 *>(&a.z), *reinterpret_cast<const __half2 *>(&b.z));
    *reinterpret_cast<__half2 *>(&r.w) = __hmin2(*reinterpret_cast<const __half2 *>(&a.w), *reinterpret_cast<const __half2 *>(&b.w));
    return r;
}

constexpr int kZWords = 256;                  // words per block

__global__ void __launch_bounds__(1024)
z_build16_kernel(const __grid_constant__ ZArgs A)
{
    __shared__ uint4 tile[32][33];            // [doc][word octet]
    const int x = threadIdx.x, y = threadIdx.y;
    const int w0 = blockIdx.y * kZWords, j0 = blockIdx.x * 32;
    const int j = j0 + y, w = w0 + 8 * x;
    const unsigned short infb = 0x7c00;
    uint4 best = make_uint4(0x7c007c00u, 0x7c007c00u, 0x7c007c00u, 0x7c007c00u);
    const bool fast = (A.V & 7) == 0 && w + 7 < A.V;               // rows of D16 are 16-byte aligned and this lane's octet is whole
    if (j < A.ndocs) {                                             // warp-uniform (y is the warp)
        const int64_t a = A.off[A.doc0 + j];
        const int u = A.uniq[A.doc0 + j];
        for (int k0 = 0; k0 < u; k0 += kWarp) {
            const int mine = k0 + x < u ? A.rows[a + k0 + x] : 0;
            const int nk = min(kWarp, u - k0);
            for (int k = 0; k < nk; k += 4) {                      // four table reads in flight per lane: one at a time the loop waits on each
                int rr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) rr[c] = __shfl_sync(kFull, mine, min(k + c, nk - 1));     // a repeated row does not change a minimum
                if (fast) {
                    uint4 v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = __ldg(reinterpret_cast<const uint4 *>(A.D16 + (int64_t)rr[c] * A.V + w));
                    best = hmin8(best, hmin8(hmin8(v[0], v[1]), hmin8(v[2], v[3])));
                } else if (w < A.V) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const __half *r = A.D16 + (int64_t)rr[c] * A.V;
                        unsigned h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) h[e] = w + e < A.V ? __half_as_ushort(r[w + e]) : infb;
                        best = hmin8(best, make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16)));
                    }
                }
            }
        }
    }
    tile[y][x] = best;
    __syncthreads();
    // transposed write: thread (x = document, y = word octet) stores eight rows of Z
    const int jw = j0 + x, ww = w0 + 8 * y;
    if (jw < A.ndocs && ww < A.V) {
        const uint4 v = tile[x][y];
        const unsigned q[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (ww + c < A.V) A.Z[(int64_t)(ww + c) * A.ldz + jw] = __ushort_as_half((unsigned short)((q[c >> 1] >> (16 * (c & 1))) & 0xffffu));
    }
}

// ---- LB tile: 128 query rows x 128 corpus columns per block ---------------------------------------
// LB(i, j) = max(L1, L2),  L1 = sum over the tokens of A_i of weight * Z_B[row][j],  L2 = sum over the tokens of B_j of
// weight * Z_A[row][i].  Every operand is rounded down and every FMA rounds toward zero (all terms are >= 0), so the
// float32 result never exceeds the real-valued bound.  A warp task is one document and 128 documents of the other
// side (4 per lane, one 8-byte load of Z per term): the document's (row, weight) list is warp-uniform, the Z loads are
// coalesced, ~3 instructions per (pair, term).  L2 tasks run first and leave their tile transposed in shared memory.
struct LbArgs {
    const int2 *listA; const int64_t *offA; const int32_t *uniqA; const int32_t *nvalA;       // query side (set A)
    int64_t i0; int32_t ni;                                                                   // rows [i0, i0 + ni)
    const int2 *listB; const int64_t *offB; const int32_t *uniqB; const int32_t *nvalB;       // corpus side (set B)
    int32_t nB;
    const __half *ZB; int64_t ldzb;           // [V][ldzb], ldzb a multiple of 128
    const __half *ZA; int64_t ldza;           // [V][ldza] for this block of query rows, ldza a multiple of 128
    float *LB; int64_t ldlb;                  // [ni][ldlb]
};

constexpr int kLbTile = 128;
constexpr int kLbPitch = kLbTile + 1;
constexpr int kLbWarps = 16;
// Shared-memory transpose without bank conflicts: element (j, i) of the 128 x 128 tile lives at row (j % 4) * 32 + j / 4,
// column (i % 4) * 32 + i / 4 -- a lane that owns four consecutive i (or j) then touches 32 different banks per access.
__device__ __forceinline__ int lb_swz(int t) { return (t & 3) * 32 + (t >> 2); }

__device__ __forceinline__ void lb_fma4(float wgt, uint2 raw, float (&acc)[4])
{
    const float2 z01 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
    const float2 z23 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    acc[0] = __fmaf_rz(wgt, z01.x, acc[0]); acc[1] = __fmaf_rz(wgt, z01.y, acc[1]);
    acc[2] = __fmaf_rz(wgt, z23.x, acc[2]); acc[3] = __fmaf_rz(wgt, z23.y, acc[3]);
}

// Terms four at a time: the list entries are warp-uniform, the four Z loads are independent, so four L2 reads are in
// flight per lane instead of one (the loop was bound by the latency of one load per term: 43 % issue-active, 16 % of the
// L2's throughput in ncu).
__device__ __forceinline__ void lb_accumulate(const int2 *list, int u, const __half *Z, int64_t ldz, int col, float (&acc)[4])
{
    int k = 0;
    for (; k + 4 <= u; k += 4) {
        int2 e[4];
        uint2 raw[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = __ldg(list + k + c);              // warp-uniform
#pragma unroll
        for (int c = 0; c < 4; ++c) raw[c] = __ldg(reinterpret_cast<const uint2 *>(Z + (int64_t)e[c].x * ldz + col));
#pragma unroll
        for (int c = 0; c < 4; ++c) lb_fma4(__int_as_float(e[c].y), raw[c], acc);
    }
    for (; k < u; ++k) {
        const int2 e = __ldg(list + k);
        lb_fma4(__int_as_float(e.y), __ldg(reinterpret_cast<const uint2 *>(Z + (int64_t)e.x * ldz + col)), acc);
    }
}

__global__ void __launch_bounds__(32 * kLbWarps)
lb_tile16_kernel(const __grid_constant__ LbArgs A)
{
    extern __shared__ float l2t[];                                           // [128 corpus docs][kLbPitch] : L2(i, j) at [j][i]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // query blocks vary fastest: the column slice of Z_B of one corpus block stays L2-hot across the query blocks
    const int ib = blockIdx.x * kLbTile, jb = blockIdx.y * kLbTile;
    const float kInf = __int_as_float(0x7f800000);
    // L2(i, j): one corpus document per task, lanes along i
    for (int t = warp; t < kLbTile; t += kLbWarps) {
        const int j = jb + t;
        float acc[4] = { kInf, kInf, kInf, kInf };
        if (j < A.nB && A.nvalB[j] > 0) {
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
            lb_accumulate(A.listB + A.offB[j], A.uniqB[j], A.ZA, A.ldza, ib + 4 * lane, acc);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) l2t[lb_swz(t) * kLbPitch + c * 32 + lane] = acc[c];            // i = 4 lane + c
    }
    __syncthreads();
    // L1(i, j): one query document per task, lanes along j
    for (int t = warp; t < kLbTile; t += kLbWarps) {
        const int i = ib + t;
        if (i >= A.ni) break;
        float acc[4] = { kInf, kInf, kInf, kInf };
        if (A.nvalA[A.i0 + i] > 0) {
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
            lb_accumulate(A.listA + A.offA[A.i0 + i], A.uniqA[A.i0 + i], A.ZB, A.ldzb, jb + 4 * lane, acc);
        }
        float4 o;
        float *op = &o.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) op[c] = fmaxf(acc[c], l2t[(c * 32 + lane) * kLbPitch + lb_swz(t)]);     // j = 4 lane + c; inf if either side is empty
        *reinterpret_cast<float4 *>(A.LB + (int64_t)i * A.ldlb + jb + 4 * lane) = o;              // columns >= nB are padding
    }
}

// ---- per-row k-th smallest of LB -------------------------------------------------------------------
// All values are >= 0 or +inf, so the float bit patterns order like the values.  Common case (k <= 256):
// every thread keeps the minimum of its strided share; the k-th smallest of those 256 minima, T, is an
// upper bound of the row's k-th smallest (k different threads hold a value <= T), and the few values
// <= T are collected and ranked.  A radix select over the bits (4 histogram passes) is the fallback
// when more than kKthCap values qualify (heavy ties) or k > 256.
constexpr int kKthCap = 2048;

__device__ void row_kth_radix(const unsigned *row, int32_t n, int32_t k, unsigned *hist, unsigned *s_pk, float *out)
{
    unsigned prefix = 0, mask = 0, kk = (unsigned)k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += 256) {
            const unsigned v = row[j];
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned acc = 0, b = 0;
            for (; b < 256; ++b) { if (acc + hist[b] >= kk) break; acc += hist[b]; }
            s_pk[0] = prefix | (b << shift);
            s_pk[1] = kk - acc;
        }
        __syncthreads();
        prefix = s_pk[0]; kk = s_pk[1]; mask |= 0xffu << shift;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = __uint_as_float(prefix);
}

__global__ void __launch_bounds__(256)
row_kth_kernel(const float *LB, int64_t ldlb, int32_t n, int32_t k, float *kth)
{
    __shared__ unsigned s_min[256];
    __shared__ unsigned s_list[kKthCap];
    __shared__ unsigned s_T, s_cnt, s_pk[2];
    const unsigned *row = reinterpret_cast<const unsigned *>(LB + (int64_t)blockIdx.x * ldlb);
    if (k > n) { if (threadIdx.x == 0) kth[blockIdx.x] = __int_as_float(0x7f800000); return; }
    if (k <= 256 && n >= 256) {
        unsigned m = 0xffffffffu;
        for (int j = threadIdx.x; j < n; j += 256) m = min(m, row[j]);
        s_min[threadIdx.x] = m;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int rank = 0;
        for (int t = 0; t < 256; ++t) { const unsigned o = s_min[t]; rank += (o < m) || (o == m && t < (int)threadIdx.x); }
        if (rank == k - 1) s_T = m;
        __syncthreads();
        const unsigned T = s_T;
        for (int j = threadIdx.x; j < n; j += 256) {
            const unsigned v = row[j];
            if (v <= T) { const unsigned pos = atomicAdd(&s_cnt, 1u); if (pos < kKthCap) s_list[pos] = v; }
        }
        __syncthreads();
        const unsigned cnt = s_cnt;
        if (cnt <= kKthCap) {
            for (unsigned a = threadIdx.x; a < cnt; a += 256) {
                const unsigned v = s_list[a];
                unsigned r = 0;
                for (unsigned b = 0; b < cnt; ++b) { const unsigned o = s_list[b]; r += (o < v) || (o == v && b < a); }
                if (r == (unsigned)(k - 1)) kth[blockIdx.x] = __uint_as_float(v);
            }
            return;
        }
        __syncthreads();
    }
    row_kth_radix(row, n, k, s_min, s_pk, kth + blockIdx.x);
}

// ---- candidate lists: j with lo_i < LB(i, j) <= hi_i, ascending j ---------------------------------
// One block per row, each of its 8 warps owns a contiguous segment of the row and never talks to the others: pass 0
// leaves one count per (row, warp), the scan turns them into offsets, pass 1 fills at those offsets (ballot + popc
// inside the warp).  The first version synchronised the block twice per 256 columns and ran at 1.5 TB/s; the passes
// over the bound matrix are pure streaming.  lo == nullptr: no lower limit.
constexpr int kCandWarps = 8;

__global__ void __launch_bounds__(256)
cand_rows_kernel(const float *LB, int64_t ldlb, int32_t n, const float *lo, const float *hi, int32_t fill,
                 int32_t *counts /* [rows * 8] */, const int64_t *offsets /* [rows * 8 + 1] */, int32_t row0, int32_t *ci, int32_t *cj)
{
    const int r = blockIdx.x;
    const float *row = LB + (int64_t)r * ldlb;
    const float h = hi[r];
    const bool has_lo = lo != nullptr;
    const float l = has_lo ? lo[r] : 0.f;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = ((n + kCandWarps - 1) / kCandWarps + 127) & ~127;           // columns per warp, a multiple of 128
    const int j0 = wid * seg, j1 = min(n, j0 + seg);
    const unsigned lt = (1u << lane) - 1u;
    int total = 0;
    int64_t base = fill ? offsets[(int64_t)r * kCandWarps + wid] : 0;
    for (int jb = j0; jb < j1; jb += 128) {                                     // lane: columns jb + 4 lane .. + 3, one 16-byte load
        const int j = jb + 4 * lane;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < j1) v = __ldcs(reinterpret_cast<const float4 *>(row + j));      // rows are padded to a multiple of 128 columns
        const float vv[4] = { v.x, v.y, v.z, v.w };
        unsigned mine = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (j + c < j1 && vv[c] <= h && (!has_lo || vv[c] > l)) mine |= 1u << c;
        if (!fill) { total += __popc(mine); continue; }
        // ascending j: everything of lower lanes first, then this lane's own earlier columns
        int before = 0, all = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned bal = __ballot_sync(kFull, (mine >> c) & 1u);
            before += __popc(bal & lt); all += __popc(bal);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if ((mine >> c) & 1u) { const int64_t pos = base + before + __popc(mine & ((1u << c) - 1u)); ci[pos] = row0 + r; cj[pos] = j + c; }
        base += all;
    }
    if (!fill) total = __reduce_add_sync(kFull, total);
    if (!fill && lane == 0) counts[(int64_t)r * kCandWarps + wid] = total;
}

// exclusive scan of up to a few thousand row counts (one block)
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t *counts, int32_t n, int64_t *offsets /* n + 1 */)
{
    __shared__ long long s_part[1024];
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    long long sum = 0;
    for (int i = b; i < e; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int t = 0; t < 1024; ++t) { const long long v = s_part[t]; s_part[t] = acc; acc += v; }
        offsets[n] = acc;
    }
    __syncthreads();
    long long acc = s_part[threadIdx.x];
    for (int i = b; i < e; ++i) { offsets[i] = acc; acc += counts[i]; }
}

// ---- per-row top-k merge ---------------------------------------------------------------------------
// Row r merges its current best list (top_j / top_d, kcur[r] valid entries, sorted) with the new
// candidates [offsets[r], offsets[r + 1]) and keeps the k smallest by (distance, j).  Distances are
// >= 0 or +inf (NaN never occurs), so their bit patterns order like the values.
__global__ void __launch_bounds__(256)
topk_merge_kernel(int32_t k, const int64_t *offsets /* [rows * kCandWarps + 1] */, const int32_t *cj, const double *cd,
                  int32_t *top_j, double *top_d, int32_t *kcur, float *thr /* k-th distance or +inf, rounded up */, float dmax)
{
    extern __shared__ unsigned char sm_raw[];
    unsigned long long *okey = reinterpret_cast<unsigned long long *>(sm_raw);      // [k] merged keys
    int *oj = reinterpret_cast<int *>(okey + k);                                     // [k]
    __shared__ unsigned long long s_bk[8];
    __shared__ int s_bj[8];
    __shared__ unsigned long long s_lastk;
    __shared__ int s_lastj;
    const int r = blockIdx.x;
    const int64_t b = offsets[(int64_t)r * kCandWarps], e = offsets[(int64_t)(r + 1) * kCandWarps];
    const int nold = kcur[r];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t total = (e - b) + nold;
    const int nout = (int)(total < k ? total : k);
    unsigned long long lastk = 0; int lastj = -1;                 // everything <= (lastk, lastj) is already taken
    for (int o = 0; o < nout; ++o) {
        unsigned long long bk = ~0ull; int bj = 0x7fffffff;
        for (int64_t c = threadIdx.x; c < total; c += 256) {
            unsigned long long key; int j;
            if (c < nold) { key = (unsigned long long)__double_as_longlong(top_d[(int64_t)r * k + c]); j = top_j[(int64_t)r * k + c]; }
            else { key = (unsigned long long)__double_as_longlong(cd[b + (c - nold)]); j = cj[b + (c - nold)]; }
            const bool after = o == 0 || key > lastk || (key == lastk && j > lastj);
            if (after && (key < bk || (key == bk && j < bj))) { bk = key; bj = j; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(kFull, bk, s);
            const int ojj = __shfl_xor_sync(kFull, bj, s);
            if (ok < bk || (ok == bk && ojj < bj)) { bk = ok; bj = ojj; }
        }
        if (lane == 0) { s_bk[wid] = bk; s_bj[wid] = bj; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w)
                if (s_bk[w] < bk || (s_bk[w] == bk && s_bj[w] < bj)) { bk = s_bk[w]; bj = s_bj[w]; }
            okey[o] = bk; oj[o] = bj; s_lastk = bk; s_lastj = bj;
        }
        __syncthreads();
        lastk = s_lastk; lastj = s_lastj;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nout; o += 256) {
        top_d[(int64_t)r * k + o] = __longlong_as_double((long long)okey[o]);
        top_j[(int64_t)r * k + o] = oj[o];
    }
    if (threadIdx.x == 0) {
        kcur[r] = nout;
        float t = __int_as_float(0x7f800000);
        if (nout == k) t = __double2float_ru(__longlong_as_double((long long)okey[k - 1]) * (1.0 + kLbMarginRel) + kLbMarginRel * (double)dmax);
        thr[r] = t;
    }
}

}  // namespace wmd
