// cost.cuh -- K2: embedding-row gather + per-pair Euclidean cost tile.
//
// Replaces gensim wmdistance's python double loop:
//     D[i, j] = sqrt(np_sum((wv[t_i] - wv[t_j]) ** 2))          (float32, numpy summation order)
// for every unique doc1 token i and unique doc2 token j of a pair, plus the tile maximum
// (pyemd's maxC, S6(b)).  Bit-exact with numpy: every subtract, multiply, add and the square
// root are separately rounded float32 operations (no FMA contraction), accumulated in numpy's
// FLOAT_pairwise_sum order -- eight strided accumulators per leaf block, leaves combined by
// the recursion tree that SumPlan flattens.
//
// Structure (sm_100a).  One persistent CTA per SM runs a producer/consumer ring:
//   * warp 0 (producer) walks its slice of pairs, packs consecutive pairs into a stage
//     ("group"), and gathers their table rows HBM/L2 -> shared memory with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx), one copy per row, each row fetched once per pair;
//   * the consumer warps wait on the stage's "full" mbarrier, pull tile tasks from a shared
//     counter, and release the stage through its "empty" mbarrier; they never barrier with each
//     other, so a warp that runs out of tiles in one stage starts on the next.
// A tile task is a 2x4 block of cells owned by 2*PL adjacent lanes: lane (leaf l, half h) keeps
// the accumulator quads r[4h..4h+3] of leaf block l for all eight cells, i.e. 6 LDS.128 per 96
// float operations, which balances the 128 B/clk shared-memory port against the FP32 pipe.
// The float math is issued as packed FADD2/FFMA2 (sub.f32x2, fma.f32x2 with a -0.0 addend that
// ptxas cannot see, add.f32x2): identical IEEE roundings, half the issue slots.
#pragma once
#include "common.cuh"

namespace wmd {

constexpr int kCostConsumerWarps = 16;
constexpr int kCostThreads = 32 * (1 + kCostConsumerWarps);
constexpr int kGroupMax = 32;        // pairs per staged group
constexpr int kPlanDepth = 8;
constexpr int kMaxStages = 4;

struct CostArgs {
    Vocab vc;
    SumPlan plan;
    DocSide s1, s2;                  // only .off / .L are used (document slots)
    int64_t p0;                      // first pair of this chunk
    int32_t npairs;                  // pairs in this chunk
    int32_t tb;                      // max rows per side of a staged unit (<= 32)
    int32_t rcap;                    // row capacity of one stage
    int32_t ldr;                     // floats between staged rows (multiple of 4)
    int32_t stages;                  // ring depth (2..kMaxStages)
    int32_t pl;                      // leaf blocks processed in parallel by one tile (1, 2 or 4)
    int32_t rowbytes;                // bytes copied per row (ld * 4, multiple of 16)
    int32_t _pad;
    unsigned long long negzero2;     // 0x8000000080000000: (-0.0f, -0.0f), opaque to ptxas
    const int32_t *rows1, *rows2;    // from K1
    const int32_t *u12;
    float *tiles;                    // [npairs, tile_stride]
    int64_t tile_stride;
    float *maxc;                     // [npairs]
};

struct CostUnit {
    int32_t q;                       // pair (chunk-local)
    int32_t rowbase;                 // first staged row
    int32_t i0, ni, j0, nj;          // sub-block of the pair's tile (doc1 rows x doc2 rows)
    int32_t u2;                      // tile row pitch
    int32_t tilebase;                // first tile task of the unit inside its group
    int32_t tr;                      // 1: the 2-side of the 2x4 tile runs along doc2
    int32_t _pad;
    int64_t o1, o2;                  // token-slot offsets of the pair
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > (1ll << 33)) __trap();       // ~4 s: a protocol bug must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_row_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// x*x rounded once: fma with a -0.0 addend held in a register ptxas cannot fold (it would
// otherwise turn mul+add into one FFMA2 and lose numpy's intermediate rounding)
__device__ __forceinline__ f32x2 sq2(f32x2 a, f32x2 nz) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(a), "l"(nz)); return r; }
__device__ __forceinline__ float lo_f(f32x2 a) { return __uint_as_float((unsigned)(a & 0xffffffffull)); }
__device__ __forceinline__ float hi_f(f32x2 a) { return __uint_as_float((unsigned)(a >> 32)); }

struct Q4 { f32x2 lo, hi; };          // four consecutive floats
__device__ __forceinline__ Q4 ldq(const float *p)
{
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    return Q4{ v.x, v.y };
}
__device__ __forceinline__ float quad_sum(const Q4 a)
{
    return __fadd_rn(__fadd_rn(lo_f(a.lo), hi_f(a.lo)), __fadd_rn(lo_f(a.hi), hi_f(a.hi)));
}

// One leaf block [start, start+len) of numpy's pairwise sum for the 2x4 cells (a_r, b_c), c = 4r + cc.
// half selects accumulators r[0..3] or r[4..7].  The two lanes of a pair meet in a reduce-scatter:
// afterwards lane `half` holds the four cells c = 2j + half (j = 0..3) in res[j], tail included.
__device__ __forceinline__ void leaf_2x4(const float *const (&a)[2], const float *const (&b)[4], int start, int len,
                                         int half, f32x2 nz, float (&res)[4])
{
    const float *bk[2] = { half ? b[1] : b[0], half ? b[3] : b[2] };      // columns of the kept cells: half, 2 + half
    if (len < 8) {                                     // numpy: plain sequential loop (only when d < 8)
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = 0.f;
        for (int e = start; e < start + len; ++e) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t = __fsub_rn(a[j >> 1][e], bk[j & 1][e]);
                res[j] = __fadd_rn(res[j], __fmul_rn(t, t));
            }
        }
        return;
    }
    const int nfull = len - (len & 7);
    int e = start + 4 * half;
    const int eend = start + nfull;
    Q4 acc[8];
    {
        Q4 x[2], y[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) x[r] = ldq(a[r] + e);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = ldq(b[c] + e);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[r * 4 + c].lo = sq2(sub2(x[r].lo, y[c].lo), nz);
                acc[r * 4 + c].hi = sq2(sub2(x[r].hi, y[c].hi), nz);
            }
    }
#pragma unroll 2
    for (e += 8; e < eend; e += 8) {
        Q4 x[2], y[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) x[r] = ldq(a[r] + e);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = ldq(b[c] + e);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[r * 4 + c].lo = add2(acc[r * 4 + c].lo, sq2(sub2(x[r].lo, y[c].lo), nz));
                acc[r * 4 + c].hi = add2(acc[r * 4 + c].hi, sq2(sub2(x[r].hi, y[c].hi), nz));
            }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float p0 = quad_sum(acc[2 * j]), p1 = quad_sum(acc[2 * j + 1]);
        const float mine = half ? p1 : p0, theirs = half ? p0 : p1;
        const float o = __shfl_xor_sync(kFull, theirs, 1);
        res[j] = __fadd_rn(mine, o);                               // (r0+r1+r2+r3) + (r4+..+r7); fadd commutes
    }
    for (int t = eend; t < start + len; ++t) {                     // the len % 8 tail, sequential, kept cells only
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float s = __fsub_rn(a[j >> 1][t], bk[j & 1][t]);
            res[j] = __fadd_rn(res[j], __fmul_rn(s, s));
        }
    }
}

// Distances of one tile task.  PL = 1: the lane pair walks the whole postfix program and each lane
// ends with the four cells c = 2j + half.  PL = 2 / 4: leaf l of a balanced tree is summed by lane
// pair l and the tree is closed by further reduce-scatter steps, leaving 2 / 1 cells per lane.
// out[k] is cell cell0 + k * cstep of the tile.
template <int PL>
__device__ __forceinline__ void dist_2x4(const CostArgs &A, const float *const (&a)[2], const float *const (&b)[4],
                                         int sub, float (&out)[4 / PL], int &cell0, int &cstep)
{
    const int half = sub & 1;
    if (PL == 1) {
        float st[kPlanDepth][4];
        int sp = 0;
        for (int o = 0; o < A.plan.nops; ++o) {
            float r[4];
            leaf_2x4(a, b, A.plan.start[o], A.plan.len[o], half, A.negzero2, r);
#pragma unroll
            for (int c = 0; c < 4; ++c) st[sp][c] = r[c];
            ++sp;
            for (int k = 0; k < A.plan.adds[o]; ++k) {
                --sp;
#pragma unroll
                for (int c = 0; c < 4; ++c) st[sp - 1][c] = __fadd_rn(st[sp - 1][c], st[sp][c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4 / PL; ++c) out[c] = __fsqrt_rn(st[0][c]);
        cell0 = half; cstep = 2;
    } else {
        const int l = sub >> 1, l0 = l & 1;
        float r[4];
        leaf_2x4(a, b, A.plan.start[l], A.plan.len[l], half, A.negzero2, r);
        float k2[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {                              // L0 + L1 (and L2 + L3): keep cells with bit 1 == l0
            const float mine = l0 ? r[2 * m + 1] : r[2 * m], theirs = l0 ? r[2 * m] : r[2 * m + 1];
            const float o = __shfl_xor_sync(kFull, theirs, 2);
            k2[m] = __fadd_rn(mine, o);
        }
        if (PL == 2) {
#pragma unroll
            for (int m = 0; m < 4 / PL; ++m) out[m] = __fsqrt_rn(k2[m & 1]);
            cell0 = 2 * l0 + half; cstep = 4;
        } else {
            const int l1 = l >> 1;                                 // (L0+L1) + (L2+L3): keep the cell with bit 2 == l1
            const float mine = l1 ? k2[1] : k2[0], theirs = l1 ? k2[0] : k2[1];
            const float o = __shfl_xor_sync(kFull, theirs, 4);
            out[0] = __fsqrt_rn(__fadd_rn(mine, o));
            cell0 = sub; cstep = 8;
        }
    }
}

__device__ __forceinline__ int unit_tiles(int ni, int nj, int tr)
{
    const int na = tr ? nj : ni, nb = tr ? ni : nj;
    return ((na + 1) >> 1) * ((nb + 3) >> 2);
}
__device__ __forceinline__ int pick_orientation(int ni, int nj)
{
    const int p0 = ((ni + 1) >> 1) * 2 * ((nj + 3) >> 2) * 4;      // padded cells, 2-side along doc1
    const int p1 = ((nj + 1) >> 1) * 2 * ((ni + 3) >> 2) * 4;
    return p1 < p0 ? 1 : 0;
}

// Executes one warp-wide batch of tile tasks [t0, t0 + 32 / (2 PL)) of a group.
template <int PL>
__device__ __forceinline__ void run_tile_batch(const CostArgs &A, const CostUnit *units, int nunits, int ntiles, int t0,
                                               const float *rowsbuf, unsigned *umax)
{
    constexpr int LPT = 2 * PL;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT;
    int t = t0 + lane / LPT;
    const bool live = t < ntiles;
    if (!live) t = t0;                                   // clamp: recompute a valid tile, discard
    int g = 0;
    while (g + 1 < nunits && units[g + 1].tilebase <= t) ++g;
    const CostUnit &un = units[g];
    const int local = t - un.tilebase;
    const int na = un.tr ? un.nj : un.ni, nb = un.tr ? un.ni : un.nj;
    const int abase = un.rowbase + (un.tr ? un.ni : 0), bbase = un.rowbase + (un.tr ? 0 : un.ni);
    const int TI = (na + 1) >> 1, TJ = (nb + 3) >> 2;
    const int ti = local / TJ, tj = local - ti * TJ;
    int ia[2], jb[4];
    const float *a[2], *b[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) { ia[r] = ti + r * TI; a[r] = rowsbuf + (size_t)(abase + (ia[r] < na ? ia[r] : ti)) * A.ldr; }
#pragma unroll
    for (int c = 0; c < 4; ++c) { jb[c] = tj + c * TJ; b[c] = rowsbuf + (size_t)(bbase + (jb[c] < nb ? jb[c] : tj)) * A.ldr; }
    float v[4 / PL];
    int cell0, cstep;
    dist_2x4<PL>(A, a, b, sub, v, cell0, cstep);
    if (live) {
        float *tile_p = A.tiles + (int64_t)un.q * A.tile_stride;
        float mx = 0.f;
#pragma unroll
        for (int k = 0; k < 4 / PL; ++k) {
            const int c = cell0 + k * cstep;
            const int r = c >> 2, cc = c & 3;
            const int iar = r ? ia[1] : ia[0];
            const int jbc = cc == 0 ? jb[0] : (cc == 1 ? jb[1] : (cc == 2 ? jb[2] : jb[3]));
            if (iar < na && jbc < nb) {
                const int i = un.i0 + (un.tr ? jbc : iar);
                const int j = un.j0 + (un.tr ? iar : jbc);
                tile_p[(int64_t)i * un.u2 + j] = v[k];
                mx = fmaxf(mx, v[k]);
            }
        }
        atomicMax(&umax[g], __float_as_uint(mx));        // distances are >= 0: uint order == float order
    }
}

struct CostStage {
    CostUnit units[kGroupMax];
    unsigned umax[kGroupMax];
    int nunits, nrows, ntiles, taskctr;
};

// Small pairs (u1 <= tb and u2 <= tb): producer/consumer ring, one CTA per SM.
template <int PL>
__global__ void __launch_bounds__(kCostThreads, 1)
cost_tiles_kernel(const __grid_constant__ CostArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ CostStage stage[kMaxStages];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = A.stages;
    const size_t stage_floats = (size_t)A.rcap * A.ldr;
    float *rows_all = reinterpret_cast<float *>(smem_raw);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kCostConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int per = (A.npairs + gridDim.x - 1) / gridDim.x;
    const int begin = min(A.npairs, (int)blockIdx.x * per);
    const int end = min(A.npairs, begin + per);

    if (warp == 0) {
        // ------------------------------ producer ------------------------------
        int64_t tok1, tok2;
        { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
        int cur = begin;
        int it = 0;
        for (;; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            CostStage &G = stage[s];
            mbar_wait(&empty_bar[s], ph ^ 1u);                      // stage drained (passes at once the first time round)
            if (it >= S) {                                          // epilogue of the group that lived here
                if (lane < G.nunits && G.units[lane].ni > 0) A.maxc[G.units[lane].q] = __uint_as_float(G.umax[lane]);
            }
            __syncwarp();
            if (cur >= end) {
                // no more work: publish the stop marker once per remaining stage so that every consumer sees it
                if (lane == 0) { G.nunits = 0; G.nrows = 0; G.ntiles = -1; G.taskctr = 0; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
                break;
            }
            const int q = cur + lane;
            int u1 = 0, u2 = 0;
            if (q < end) { const int u = A.u12[q]; u1 = u & 0xffff; u2 = u >> 16; }
            if (u1 > A.tb || u2 > A.tb) { u1 = 0; u2 = 0; }         // handled by the large-pair kernel
            const int tr = pick_orientation(u1, u2);
            const int rows = u1 + u2;
            const int tiles = (u1 > 0) ? unit_tiles(u1, u2, tr) : 0;
            int rs = rows, ts = tiles;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int r = __shfl_up_sync(kFull, rs, o), t2 = __shfl_up_sync(kFull, ts, o);
                if (lane >= o) { rs += r; ts += t2; }
            }
            const unsigned fits = __ballot_sync(kFull, rs <= A.rcap && q < end);
            const int cnt = (fits == kFull) ? 32 : (__ffs(~fits) - 1);   // >= 1: one pair always fits
            if (lane < cnt) {
                CostUnit un;
                un.q = q; un.rowbase = rs - rows; un.i0 = 0; un.ni = u1; un.j0 = 0; un.nj = u2; un.u2 = u2;
                un.tilebase = ts - tiles; un.tr = tr; un._pad = 0;
                int64_t aa; int l;
                doc_span(A.s1, A.p0 + q, aa, l); un.o1 = aa - tok1;
                doc_span(A.s2, A.p0 + q, aa, l); un.o2 = aa - tok2;
                G.units[lane] = un;
                G.umax[lane] = 0u;
            }
            const int total_rows = __shfl_sync(kFull, rs, cnt - 1);
            const int total_tiles = __shfl_sync(kFull, ts, cnt - 1);
            if (lane == 0) { G.nunits = cnt; G.nrows = total_rows; G.ntiles = total_tiles; G.taskctr = 0; }
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)total_rows * (uint32_t)A.rowbytes);
            __syncwarp();
            float *rowsbuf = rows_all + (size_t)s * stage_floats;
            for (int rr = lane; rr < total_rows; rr += kWarp) {
                int g = 0;
                while (g + 1 < cnt && G.units[g + 1].rowbase <= rr) ++g;
                const CostUnit &un = G.units[g];
                const int local = rr - un.rowbase;
                const int row = local < un.ni ? A.rows1[un.o1 + local] : A.rows2[un.o2 + (local - un.ni)];
                tma_row_g2s(rowsbuf + (size_t)rr * A.ldr, A.vc.table + (int64_t)row * A.vc.ld, (uint32_t)A.rowbytes, &full_bar[s]);
            }
            cur += cnt;
        }
        // drain: the groups still in flight hand their maxima over when their stage empties
        for (int k = 1; k < S; ++k) {
            const int it2 = it + k;
            const int s = it2 % S;
            if (it2 < S) continue;                                  // that stage was never used
            const uint32_t ph = (uint32_t)(it2 / S) & 1u;
            CostStage &G = stage[s];
            mbar_wait(&empty_bar[s], ph ^ 1u);
            if (lane < G.nunits && G.units[lane].ni > 0) A.maxc[G.units[lane].q] = __uint_as_float(G.umax[lane]);
            __syncwarp();
        }
    } else {
        // ------------------------------ consumers ------------------------------
        constexpr int TPW = 32 / (2 * PL);
        for (int it = 0;; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            CostStage &G = stage[s];
            mbar_wait(&full_bar[s], ph);
            const int ntiles = G.ntiles;
            if (ntiles < 0) break;                                  // stop marker
            const int nunits = G.nunits;
            const float *rowsbuf = rows_all + (size_t)s * stage_floats;
            for (;;) {
                int t0 = 0;
                if (lane == 0) t0 = atomicAdd(&G.taskctr, TPW);
                t0 = __shfl_sync(kFull, t0, 0);
                if (t0 >= ntiles) break;
                run_tile_batch<PL>(A, G.units, nunits, ntiles, t0, rowsbuf, G.umax);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
}

// Large pairs (a side with more than tb unique rows): one CTA per pair, tb x tb blocks in turn,
// staged with plain coalesced 128-bit loads (rare path: documents longer than 32 unique tokens).
template <int PL>
__global__ void __launch_bounds__(kCostThreads)
cost_tiles_large_kernel(const __grid_constant__ CostArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *rowsbuf = reinterpret_cast<float *>(smem_raw);
    __shared__ CostUnit unit;
    __shared__ unsigned umax;
    constexpr int TPW = 32 / (2 * PL);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const int d4 = A.vc.ld >> 2;
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
                    un.j0 = bj; un.nj = min(A.tb, u2 - bj); un.u2 = u2; un.tilebase = 0;
                    un.tr = pick_orientation(un.ni, un.nj); un._pad = 0;
                    int64_t aa; int l;
                    doc_span(A.s1, A.p0 + q, aa, l); un.o1 = aa - tok1;
                    doc_span(A.s2, A.p0 + q, aa, l); un.o2 = aa - tok2;
                    unit = un;
                }
                __syncthreads();
                const int nrows = unit.ni + unit.nj;
                for (int rr = warp; rr < nrows; rr += nwarps) {
                    const int row = rr < unit.ni ? A.rows1[unit.o1 + unit.i0 + rr] : A.rows2[unit.o2 + unit.j0 + (rr - unit.ni)];
                    const float4 *src = reinterpret_cast<const float4 *>(A.vc.table + (int64_t)row * A.vc.ld);
                    float4 *dst = reinterpret_cast<float4 *>(rowsbuf + (size_t)rr * A.ldr);
                    for (int k = lane; k < d4; k += kWarp) dst[k] = __ldg(src + k);
                }
                __syncthreads();
                const int ntiles = unit_tiles(unit.ni, unit.nj, unit.tr);
                for (int t0 = warp * TPW; t0 < ntiles; t0 += nwarps * TPW)
                    run_tile_batch<PL>(A, &unit, 1, ntiles, t0, rowsbuf, &umax);
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
