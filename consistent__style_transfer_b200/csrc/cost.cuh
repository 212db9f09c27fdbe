// cost.cuh -- K2: embedding-row gather + per-pair Euclidean cost tile.
//
// Replaces gensim wmdistance's python double loop:
//     D[i, j] = sqrt(np_sum((wv[t_i] - wv[t_j]) ** 2))          (float32, numpy summation order)
// for every unique doc1 token i and unique doc2 token j of a pair, plus the tile maximum
// (pyemd's maxC, S6(b)).  Bit-exact with numpy: every subtract, multiply, add and the square
// root are separately rounded float32 operations (no FMA), accumulated in numpy's
// FLOAT_pairwise_sum order -- eight strided accumulators per leaf block, leaves combined by
// the recursion tree that SumPlan flattens.
//
// Mapping.  A CTA stages the table rows of a *group* of consecutive pairs in shared memory
// (each row read from L2/HBM exactly once per pair, 128-bit coalesced loads), then all threads
// sweep a flattened list of lane tasks.  A task is half of a 2x2 cell tile: two adjacent lanes
// own the accumulator quads r[0..3] and r[4..7] of the same four cells and meet with one
// shuffle per leaf, so shared-memory traffic is 4 LDS.128 per 48 FP32 operations.
#pragma once
#include "common.cuh"

namespace wmd {

constexpr int kCostThreads = 256;
constexpr int kGroupMax = 32;        // pairs per staged group
constexpr int kPlanDepth = 8;

struct CostArgs {
    Vocab vc;
    SumPlan plan;
    DocSide s1, s2;                  // only .off / .L are used (document slots)
    int64_t p0;                      // first pair of this chunk
    int32_t npairs;                  // pairs in this chunk
    int32_t tb;                      // max rows per side of a staged unit (<= 32)
    int32_t rcap;                    // row capacity of the staging buffer
    int32_t ldr;                     // floats between staged rows (multiple of 4; ldr/4 odd)
    const int32_t *rows1, *rows2;    // from K1
    const int32_t *u12;
    float *tiles;                    // [npairs, tile_stride]
    int64_t tile_stride;
    float *maxc;                     // [npairs]
};

struct CostUnit {
    int32_t q;                       // pair (chunk-local)
    int32_t rowbase;                 // first staged row
    int32_t i0, ni, j0, nj;          // sub-block of the pair's tile
    int32_t u2;                      // tile row pitch
    int32_t taskbase;
    int64_t o1, o2;                  // token-slot offsets of the pair
};

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__device__ __forceinline__ void sq_acc_init(float4 &acc, const float4 a, const float4 b)
{
    float t;
    t = __fsub_rn(a.x, b.x); acc.x = __fmul_rn(t, t);
    t = __fsub_rn(a.y, b.y); acc.y = __fmul_rn(t, t);
    t = __fsub_rn(a.z, b.z); acc.z = __fmul_rn(t, t);
    t = __fsub_rn(a.w, b.w); acc.w = __fmul_rn(t, t);
}
__device__ __forceinline__ void sq_acc(float4 &acc, const float4 a, const float4 b)
{
    float t;
    t = __fsub_rn(a.x, b.x); acc.x = __fadd_rn(acc.x, __fmul_rn(t, t));
    t = __fsub_rn(a.y, b.y); acc.y = __fadd_rn(acc.y, __fmul_rn(t, t));
    t = __fsub_rn(a.z, b.z); acc.z = __fadd_rn(acc.z, __fmul_rn(t, t));
    t = __fsub_rn(a.w, b.w); acc.w = __fadd_rn(acc.w, __fmul_rn(t, t));
}
__device__ __forceinline__ float quad_sum(const float4 a)
{
    return __fadd_rn(__fadd_rn(a.x, a.y), __fadd_rn(a.z, a.w));
}

// One leaf block of numpy's pairwise sum for the four cells (a0,b0) (a0,b1) (a1,b0) (a1,b1).
// `half` selects accumulators r[0..3] (0) or r[4..7] (1); both lanes of a pair return the same sums.
__device__ __forceinline__ void leaf_2x2(const float *a0, const float *a1, const float *b0, const float *b1,
                                         int start, int len, int half, float (&res)[4])
{
    if (len < 8) {                                     // numpy: plain sequential loop
        res[0] = res[1] = res[2] = res[3] = 0.f;
        for (int e = start; e < start + len; ++e) {
            const float x0 = a0[e], x1 = a1[e], y0 = b0[e], y1 = b1[e];
            float t;
            t = __fsub_rn(x0, y0); res[0] = __fadd_rn(res[0], __fmul_rn(t, t));
            t = __fsub_rn(x0, y1); res[1] = __fadd_rn(res[1], __fmul_rn(t, t));
            t = __fsub_rn(x1, y0); res[2] = __fadd_rn(res[2], __fmul_rn(t, t));
            t = __fsub_rn(x1, y1); res[3] = __fadd_rn(res[3], __fmul_rn(t, t));
        }
        return;
    }
    const int nfull = len - (len & 7);
    int e = start + 4 * half;
    float4 c00, c01, c10, c11;
    {
        const float4 x0 = lds4(a0 + e), x1 = lds4(a1 + e), y0 = lds4(b0 + e), y1 = lds4(b1 + e);
        sq_acc_init(c00, x0, y0); sq_acc_init(c01, x0, y1); sq_acc_init(c10, x1, y0); sq_acc_init(c11, x1, y1);
    }
    const int eend = start + nfull;
#pragma unroll 2
    for (e += 8; e < eend; e += 8) {
        const float4 x0 = lds4(a0 + e), x1 = lds4(a1 + e), y0 = lds4(b0 + e), y1 = lds4(b1 + e);
        sq_acc(c00, x0, y0); sq_acc(c01, x0, y1); sq_acc(c10, x1, y0); sq_acc(c11, x1, y1);
    }
    float p[4] = { quad_sum(c00), quad_sum(c01), quad_sum(c10), quad_sum(c11) };
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float o = __shfl_xor_sync(kFull, p[c], 1);
        res[c] = half ? __fadd_rn(o, p[c]) : __fadd_rn(p[c], o);   // (r0+r1+r2+r3) + (r4+..+r7)
    }
    for (int t = eend; t < start + len; ++t) {         // the len % 8 tail, sequential
        const float x0 = a0[t], x1 = a1[t], y0 = b0[t], y1 = b1[t];
        float s;
        s = __fsub_rn(x0, y0); res[0] = __fadd_rn(res[0], __fmul_rn(s, s));
        s = __fsub_rn(x0, y1); res[1] = __fadd_rn(res[1], __fmul_rn(s, s));
        s = __fsub_rn(x1, y0); res[2] = __fadd_rn(res[2], __fmul_rn(s, s));
        s = __fsub_rn(x1, y1); res[3] = __fadd_rn(res[3], __fmul_rn(s, s));
    }
}

// sqrt(sum((a-b)^2)) for the four cells of a lane task, numpy order.
__device__ __forceinline__ void dist_2x2(const SumPlan &plan, const float *a0, const float *a1,
                                         const float *b0, const float *b1, int half, float (&out)[4])
{
    float st[kPlanDepth][4];
    int sp = 0;
    for (int o = 0; o < plan.nops; ++o) {
        float r[4];
        leaf_2x2(a0, a1, b0, b1, plan.start[o], plan.len[o], half, r);
#pragma unroll
        for (int c = 0; c < 4; ++c) st[sp][c] = r[c];
        ++sp;
        for (int k = 0; k < plan.adds[o]; ++k) {
            --sp;
#pragma unroll
            for (int c = 0; c < 4; ++c) st[sp - 1][c] = __fadd_rn(st[sp - 1][c], st[sp][c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = __fsqrt_rn(st[0][c]);
}

// Stage `nrows` table rows (ids from K1) of unit `un` and run its tasks. Shared by both kernels.
__device__ __forceinline__ void stage_rows(const CostArgs &A, const CostUnit *units, int nunits, int total_rows,
                                           float *rowsbuf)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int d4 = A.vc.d >> 2;
    for (int rr = wib; rr < total_rows; rr += wpb) {
        int g = 0;
        while (g + 1 < nunits && units[g + 1].rowbase <= rr) ++g;
        const CostUnit &un = units[g];
        const int local = rr - un.rowbase;
        int row;
        if (local < un.ni) row = A.rows1[un.o1 + un.i0 + local];
        else               row = A.rows2[un.o2 + un.j0 + (local - un.ni)];
        const float4 *src = reinterpret_cast<const float4 *>(A.vc.table + (int64_t)row * A.vc.ld);
        float4 *dst = reinterpret_cast<float4 *>(rowsbuf + (size_t)rr * A.ldr);
        for (int k = lane; k < d4; k += kWarp) dst[k] = __ldg(src + k);
        const int rem = A.vc.d & 3;                    // d not a multiple of 4: scalar tail
        if (lane < rem) rowsbuf[(size_t)rr * A.ldr + 4 * d4 + lane] = __ldg(A.vc.table + (int64_t)row * A.vc.ld + 4 * d4 + lane);
    }
}

__device__ __forceinline__ void run_tasks(const CostArgs &A, const CostUnit *units, int nunits, int total_tasks,
                                          const float *rowsbuf, unsigned *umax)
{
    const int tid = threadIdx.x;
    for (int tbase = 0; tbase < total_tasks; tbase += blockDim.x) {
        int t = tbase + tid;
        const bool live = t < total_tasks;
        if (!live) t = (tid & 1);                      // clamp, keep the lane pair together
        int g = 0;
        while (g + 1 < nunits && units[g + 1].taskbase <= t) ++g;
        const CostUnit &un = units[g];
        const int local = t - un.taskbase;
        const int half = local & 1;
        const int tile = local >> 1;
        const int TI = (un.ni + 1) >> 1, TJ = (un.nj + 1) >> 1;
        const int ti = tile / TJ, tj = tile - ti * TJ;
        const int i0 = ti, i1 = ti + TI, j0 = tj, j1 = tj + TJ;
        const bool vi1 = i1 < un.ni, vj1 = j1 < un.nj;
        const float *a0 = rowsbuf + (size_t)(un.rowbase + i0) * A.ldr;
        const float *a1 = rowsbuf + (size_t)(un.rowbase + (vi1 ? i1 : i0)) * A.ldr;
        const float *b0 = rowsbuf + (size_t)(un.rowbase + un.ni + j0) * A.ldr;
        const float *b1 = rowsbuf + (size_t)(un.rowbase + un.ni + (vj1 ? j1 : j0)) * A.ldr;
        float v[4];
        dist_2x2(A.plan, a0, a1, b0, b1, half, v);
        if (live) {
            float *tile_p = A.tiles + (int64_t)un.q * A.tile_stride;
            float mx = 0.f;
            // half 0 stores row i0, half 1 stores row i1 (both lanes hold all four values)
            if (half == 0) {
                float *rp = tile_p + (int64_t)(un.i0 + i0) * un.u2 + un.j0;
                rp[j0] = v[0]; mx = v[0];
                if (vj1) { rp[j1] = v[1]; mx = fmaxf(mx, v[1]); }
            } else if (vi1) {
                float *rp = tile_p + (int64_t)(un.i0 + i1) * un.u2 + un.j0;
                rp[j0] = v[2]; mx = v[2];
                if (vj1) { rp[j1] = v[3]; mx = fmaxf(mx, v[3]); }
            }
            atomicMax(&umax[g], __float_as_uint(mx));   // distances are >= 0: uint order == float order
        }
    }
}

// Small pairs (u1 <= tb and u2 <= tb): persistent CTAs, each owning a contiguous slice of pairs.
__global__ void __launch_bounds__(kCostThreads)
cost_tiles_kernel(const __grid_constant__ CostArgs A)
{
    extern __shared__ __align__(16) float rowsbuf[];
    __shared__ CostUnit units[kGroupMax];
    __shared__ unsigned umax[kGroupMax];
    __shared__ int s_n, s_rows, s_tasks, s_adv;

    const int per = (A.npairs + gridDim.x - 1) / gridDim.x;
    int cur = blockIdx.x * per;
    const int end = min(A.npairs, cur + per);
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }

    while (cur < end) {
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int q = cur + lane;
            int u1 = 0, u2 = 0;
            if (q < end) { const int u = A.u12[q]; u1 = u & 0xffff; u2 = u >> 16; }
            if (u1 > A.tb || u2 > A.tb) { u1 = 0; u2 = 0; }           // handled by the large-pair kernel
            const int rows = u1 + u2;
            const int tasks = ((u1 + 1) >> 1) * ((u2 + 1) >> 1) * 2;
            int rs = rows, ts = tasks;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int r = __shfl_up_sync(kFull, rs, o), t2 = __shfl_up_sync(kFull, ts, o);
                if (lane >= o) { rs += r; ts += t2; }
            }
            const unsigned fits = __ballot_sync(kFull, rs <= A.rcap && q < end);
            const int cnt = (fits == kFull) ? 32 : (__ffs(~fits) - 1);
            if (lane < cnt) {
                CostUnit un;
                un.q = q; un.rowbase = rs - rows; un.i0 = 0; un.ni = u1; un.j0 = 0; un.nj = u2; un.u2 = u2;
                un.taskbase = ts - tasks;
                int64_t a; int l;
                doc_span(A.s1, A.p0 + q, a, l); un.o1 = a - tok1;
                doc_span(A.s2, A.p0 + q, a, l); un.o2 = a - tok2;
                units[lane] = un;
                umax[lane] = 0u;
            }
            if (lane == cnt - 1) { s_n = cnt; s_rows = rs; s_tasks = ts; s_adv = cnt; }
        }
        __syncthreads();
        const int n = s_n, total_rows = s_rows, total_tasks = s_tasks;
        stage_rows(A, units, n, total_rows, rowsbuf);
        __syncthreads();
        run_tasks(A, units, n, total_tasks, rowsbuf, umax);
        __syncthreads();
        if (threadIdx.x < n) {
            const CostUnit &un = units[threadIdx.x];
            if (un.ni > 0) A.maxc[un.q] = __uint_as_float(umax[threadIdx.x]);
        }
        cur += s_adv;
        __syncthreads();
    }
}

// Large pairs (a side with more than tb unique rows): one CTA per pair, tb x tb blocks in turn.
__global__ void __launch_bounds__(kCostThreads)
cost_tiles_large_kernel(const __grid_constant__ CostArgs A)
{
    extern __shared__ __align__(16) float rowsbuf[];
    __shared__ CostUnit unit;
    __shared__ unsigned umax;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    for (int q = blockIdx.x; q < A.npairs; q += gridDim.x) {
        const int u = A.u12[q];
        const int u1 = u & 0xffff, u2 = u >> 16;
        if (u1 <= A.tb && u2 <= A.tb) continue;
        if (threadIdx.x == 0) umax = 0u;
        for (int bi = 0; bi < u1; bi += A.tb)
            for (int bj = 0; bj < u2; bj += A.tb) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    CostUnit un;
                    un.q = q; un.rowbase = 0; un.i0 = bi; un.ni = min(A.tb, u1 - bi);
                    un.j0 = bj; un.nj = min(A.tb, u2 - bj); un.u2 = u2; un.taskbase = 0;
                    int64_t a; int l;
                    doc_span(A.s1, A.p0 + q, a, l); un.o1 = a - tok1;
                    doc_span(A.s2, A.p0 + q, a, l); un.o2 = a - tok2;
                    unit = un;
                }
                __syncthreads();
                stage_rows(A, &unit, 1, unit.ni + unit.nj, rowsbuf);
                __syncthreads();
                run_tasks(A, &unit, 1, ((unit.ni + 1) >> 1) * ((unit.nj + 1) >> 1) * 2, rowsbuf, &umax);
            }
        __syncthreads();
        if (threadIdx.x == 0) A.maxc[q] = __uint_as_float(umax);
        __syncthreads();
    }
}

// init_sims(replace=True): v /= sqrt((v ** 2).sum(-1)) per row, float32, numpy summation order.
// One-off at table load (src/wmd.py:54); one thread per row.
__global__ void normalize_rows_kernel(float *table, int64_t V, int32_t d, int32_t ld, SumPlan plan)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= V) return;
    float *v = table + r * ld;
    float st[kPlanDepth];
    int sp = 0;
    for (int o = 0; o < plan.nops; ++o) {
        const int s = plan.start[o], len = plan.len[o];
        float res;
        if (len < 8) {
            res = 0.f;
            for (int e = s; e < s + len; ++e) res = __fadd_rn(res, __fmul_rn(v[e], v[e]));
        } else {
            float acc[8];
            for (int k = 0; k < 8; ++k) acc[k] = __fmul_rn(v[s + k], v[s + k]);
            const int nfull = len - (len & 7);
            for (int e = 8; e < nfull; e += 8)
                for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(v[s + e + k], v[s + e + k]));
            res = __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3])),
                            __fadd_rn(__fadd_rn(acc[4], acc[5]), __fadd_rn(acc[6], acc[7])));
            for (int e = nfull; e < len; ++e) res = __fadd_rn(res, __fmul_rn(v[s + e], v[s + e]));
        }
        st[sp++] = res;
        for (int k = 0; k < plan.adds[o]; ++k) { --sp; st[sp - 1] = __fadd_rn(st[sp - 1], st[sp]); }
    }
    const float nrm = __fsqrt_rn(st[0]);
    for (int e = 0; e < d; ++e) v[e] = __fdiv_rn(v[e], nrm);
}

}  // namespace wmd
