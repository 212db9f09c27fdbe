// cost_fast.cuh -- K2 fast path (v5): planned stages, TMA producer warps, descriptor-driven consumers.
//
// Same arithmetic as cost.cuh (bit-exact numpy order, see there); this file only changes who does
// the bookkeeping.  v1..v4 of the kernel all landed at ~750 us per 65 536 Yelp-shape pairs with very
// different bottlenecks (ncu, profiles/README.md): a single producer warp that could not issue row
// copies fast enough (v2, v3), or per-warp pipelines whose bookkeeping and 8-copies-per-row chunking
// cost 3x the useful instructions (v4).  Here
//   * cost_plan_kernel packs consecutive pairs into STAGES (<= R table rows, <= 64 tile tasks) once,
//     in parallel, and writes for every stage the row list and one 16-byte descriptor per 2x4 tile
//     task (staged row slots, valid extents, output strides);
//   * cost_tiles_fast_kernel is one persistent CTA per SM with a ring of S stages in shared memory:
//     3 producer warps (stage k belongs to producer k % 3; stages are dealt to CTAs round-robin, so
//     no claim protocol is needed) gather whole table rows L2 -> shared memory with one 1-D TMA
//     bulk copy per row (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) plus one for the
//     stage's descriptors; 12 consumer warps wait on the stage's "full" barrier, pull batches of
//     tile tasks from a shared counter, decode one descriptor (a single LDS.128) and run the
//     2x4-cell leaf loops; a warp that finds no tiles left moves on and releases the stage through
//     its "empty" barrier.
// Pairs that do not fit a stage (u1 + u2 > R or more than 64 tile tasks: long documents) are cut by the
// plan into blocks of at most split_block_max(R) rows per side, every block a stage of its own; the
// general kernel in cost.cuh only runs when this path is switched off.
#pragma once
#include <cstddef>
#include "cost.cuh"

namespace wmd {

constexpr int kFastProducers = 3;
constexpr int kFastConsumers = 12;
constexpr int kFastThreads = 32 * (kFastProducers + kFastConsumers);
constexpr int kStageRowsMax = 96;
constexpr int kStageTilesMax = 64;
constexpr int kFastMaxStages = 12;
constexpr int kStageDescBytes = kStageTilesMax * 16;

struct TileDesc {                      // 16 bytes, read by the consumers as one uint4
    uint8_t ra[2];                     // staged row slots of the 2-side
    uint8_t rb[4];                     // staged row slots of the 4-side
    uint8_t na, nb;                    // valid rows on each side (1..2, 1..4)
    uint16_t sr, sc;                   // output strides of r (2-side) and c (4-side), in floats
    uint16_t q;                        // pair, launch-local (a launch holds <= 65 536 pairs)
    uint16_t off;                      // cell (r = 0, c = 0) inside the pair's tile
};
static_assert(sizeof(TileDesc) == 16, "TileDesc must be 16 bytes");

struct StageRec {
    int32_t nrows, ntiles, _r[2];
    int32_t rows[kStageRowsMax];       // table row of every staged row slot
    TileDesc tiles[kStageTilesMax];
};
static_assert(sizeof(StageRec) % 16 == 0 && offsetof(StageRec, tiles) % 16 == 0, "TMA alignment");

struct PlanArgs {
    DocSide s1, s2;
    int64_t p0;
    int32_t npairs;
    int32_t R, T;                      // stage capacity: rows, tile tasks
    int32_t _pad;
    const int32_t *rows1, *rows2, *u12;
    StageRec *stages;
    unsigned int *nstages;             // zeroed by the host
};

// Rows and tile descriptors of one unit: the ni x nj block at (i0, j0) of pair q's u1 x u2 tile.
// r1 / r2 point at the block's first table row on each side.
__device__ __forceinline__ void emit_unit(StageRec &S, int rowbase, int tilebase, int q, const int32_t *r1, const int32_t *r2,
                                          int ni, int nj, int i0, int j0, int u2)
{
    for (int k = 0; k < ni; ++k) S.rows[rowbase + k] = r1[k];
    for (int k = 0; k < nj; ++k) S.rows[rowbase + ni + k] = r2[k];
    const int tr = pick_orientation(ni, nj);
    const int na = tr ? nj : ni, nb = tr ? ni : nj;
    const int abase = rowbase + (tr ? ni : 0), bbase = rowbase + (tr ? 0 : ni);
    const int TI = (na + 1) >> 1, TJ = (nb + 3) >> 2;
    const unsigned sr = tr ? (unsigned)TI : (unsigned)(TI * u2);
    const unsigned sc = tr ? (unsigned)(TJ * u2) : (unsigned)TJ;
    const int tiles = TI * TJ;
    for (int t = 0; t < tiles; ++t) {
        const int ti = t / TJ, tj = t - ti * TJ;
        const int va = (ti + TI < na) ? 2 : 1;
        int vb = 1;
#pragma unroll
        for (int c = 1; c < 4; ++c) vb += (tj + c * TJ < nb);
        unsigned rowsA[2], rowsB[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) rowsA[r] = (unsigned)(abase + (r < va ? ti + r * TI : ti));
#pragma unroll
        for (int c = 0; c < 4; ++c) rowsB[c] = (unsigned)(bbase + (c < vb ? tj + c * TJ : tj));
        const unsigned off = tr ? (unsigned)((i0 + tj) * u2 + j0 + ti) : (unsigned)((i0 + ti) * u2 + j0 + tj);
        uint4 w;
        w.x = rowsA[0] | (rowsA[1] << 8) | (rowsB[0] << 16) | (rowsB[1] << 24);
        w.y = rowsB[2] | (rowsB[3] << 8) | ((unsigned)va << 16) | ((unsigned)vb << 24);
        w.z = sr | (sc << 16);
        w.w = (unsigned)q | (off << 16);
        *reinterpret_cast<uint4 *>(&S.tiles[tilebase + t]) = w;
    }
}

// Side length of the blocks a pair that does not fit one stage is cut into (every block is a stage of its own)
__host__ __device__ inline int split_block_max(int R) { return R / 2 < 20 ? R / 2 : 20; }

// One warp per block of 32 consecutive pairs (lane = pair).  Pairs that fit a stage are packed greedily with
// their neighbours; longer ones are cut into blocks of at most split_block_max(R) rows per side, one stage
// each (a table row is then fetched once per block it takes part in).
__global__ void __launch_bounds__(128)
cost_plan_kernel(const __grid_constant__ PlanArgs A)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    const int bmax = split_block_max(A.R);
    for (int blk = gw; blk * 32 < A.npairs; blk += nw) {
        const int q = blk * 32 + lane;
        int u1 = 0, u2 = 0;
        if (q < A.npairs) { const int u = A.u12[q]; u1 = u & 0xffff; u2 = u >> 16; }
        if (u1 == 0 || u2 == 0) { u1 = 0; u2 = 0; }
        const bool big = u1 > 0 && !fast_fits(u1, u2, A.R, A.T);
        // geometry of a split pair
        const int nbi = big ? (u1 + bmax - 1) / bmax : 0, nbj = big ? (u2 + bmax - 1) / bmax : 0;
        const int BI = big ? (u1 + nbi - 1) / nbi : 0, BJ = big ? (u2 + nbj - 1) / nbj : 0;
        const int nblk = nbi * nbj;
        const int pu1 = big ? 0 : u1, pu2 = big ? 0 : u2;             // what takes part in the greedy packing
        const int rows = pu1 + pu2;
        const int tiles = pu1 > 0 ? unit_tiles(pu1, pu2, pick_orientation(pu1, pu2)) : 0;
        int rs = rows, ts = tiles, bs = nblk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int r = __shfl_up_sync(kFull, rs, o), t2 = __shfl_up_sync(kFull, ts, o), b2 = __shfl_up_sync(kFull, bs, o);
            if (lane >= o) { rs += r; ts += t2; bs += b2; }
        }
        const int total_big = __shfl_sync(kFull, bs, 31);
        // greedy packing of consecutive pairs into stages (stage numbers are reserved once per warp:
        // one same-address atomic per stage serialises in L2 and cost 40 us per launch)
        int my_stage = -1, my_rowbase = 0, my_tilebase = 0;
        int hdr_r = 0, hdr_t = 0;                                     // set on the first lane of every stage
        int start = 0, base_r = 0, base_t = 0, nlocal = 0;
        while (start < 32) {
            const bool fit = lane >= start && rs - base_r <= A.R && ts - base_t <= A.T;
            const unsigned fm = __ballot_sync(kFull, fit) >> start;
            const unsigned nf = ~fm;
            const int cnt = nf ? (__ffs(nf) - 1) : 32;               // >= 1: one pair always fits
            const int end = min(32, start + max(cnt, 1));
            const int tot_r = __shfl_sync(kFull, rs, end - 1) - base_r;
            const int tot_t = __shfl_sync(kFull, ts, end - 1) - base_t;
            if (tot_r > 0) {
                if (lane >= start && lane < end) { my_stage = nlocal; my_rowbase = rs - rows - base_r; my_tilebase = ts - tiles - base_t; }
                if (lane == start) { hdr_r = tot_r; hdr_t = tot_t; }
                ++nlocal;
            }
            base_r += tot_r; base_t += tot_t; start = end;
        }
        int sbase = 0;
        if (lane == 0 && nlocal + total_big) sbase = (int)atomicAdd(A.nstages, (unsigned)(nlocal + total_big));
        sbase = __shfl_sync(kFull, sbase, 0);
        if (my_stage >= 0) my_stage += sbase;
        if (hdr_r > 0) { A.stages[my_stage].nrows = hdr_r; A.stages[my_stage].ntiles = hdr_t; }
        int64_t a1 = 0, a2 = 0;
        if (u1 > 0) { int l; doc_span(A.s1, A.p0 + q, a1, l); doc_span(A.s2, A.p0 + q, a2, l); }
        const int64_t o1 = u1 > 0 ? slot_off(A.s1, tok1, q, a1) : 0, o2 = u1 > 0 ? slot_off(A.s2, tok2, q, a2) : 0;
        // every packed pair writes its own rows and tile descriptors
        if (my_stage >= 0 && rows > 0)
            emit_unit(A.stages[my_stage], my_rowbase, my_tilebase, q, A.rows1 + o1, A.rows2 + o2, u1, u2, 0, 0, u2);
        // split pairs: the warp emits the blocks of one pair together, a block (= a stage) per lane
        unsigned bigmask = __ballot_sync(kFull, big);
        while (bigmask) {
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const int bq = __shfl_sync(kFull, q, src), bu1 = __shfl_sync(kFull, u1, src), bu2 = __shfl_sync(kFull, u2, src);
            const int bnbj = __shfl_sync(kFull, nbj, src), bBI = __shfl_sync(kFull, BI, src), bBJ = __shfl_sync(kFull, BJ, src);
            const int bn = __shfl_sync(kFull, nblk, src);
            const int first = sbase + nlocal + __shfl_sync(kFull, bs - nblk, src);
            const long long bo1 = __shfl_sync(kFull, (long long)o1, src), bo2 = __shfl_sync(kFull, (long long)o2, src);
            for (int b = lane; b < bn; b += kWarp) {
                const int bi = b / bnbj, bj = b - bi * bnbj;
                const int i0 = bi * bBI, j0 = bj * bBJ;
                const int ni = min(bBI, bu1 - i0), nj = min(bBJ, bu2 - j0);
                StageRec &S = A.stages[first + b];
                S.nrows = ni + nj;
                S.ntiles = unit_tiles(ni, nj, pick_orientation(ni, nj));
                emit_unit(S, 0, 0, bq, A.rows1 + bo1 + i0, A.rows2 + bo2 + j0, ni, nj, i0, j0, bu2);
            }
        }
    }
}

// Upper bound of the stages cost_plan_kernel can emit for Bc pairs with at most ml1 x ml2 unique rows and tok1 / tok2
// token slots in total: sum over pairs of ceil(u1 / b) * ceil(u2 / b) <= ml2 * tok1 / b^2 + (tok1 + tok2) / b + Bc.
inline int64_t plan_stage_bound(int64_t Bc, int ml1, int ml2, int64_t tok1, int64_t tok2, int R, int T)
{
    // nothing is split when even the largest possible pair fits a stage (rows AND tile tasks, as fast_fits)
    const int a = ml1 < ml2 ? ml1 : ml2, c = ml1 < ml2 ? ml2 : ml1;
    const int t0 = ((a + 1) / 2) * ((c + 3) / 4), t1 = ((c + 1) / 2) * ((a + 3) / 4);
    if (ml1 + ml2 <= R && (t0 < t1 ? t1 : t0) <= T) return Bc;
    const int64_t b = split_block_max(R) > 0 ? split_block_max(R) : 1;
    tok1 = tok1 < Bc * (int64_t)ml1 ? tok1 : Bc * (int64_t)ml1;
    tok2 = tok2 < Bc * (int64_t)ml2 ? tok2 : Bc * (int64_t)ml2;
    return (int64_t)ml2 * tok1 / (b * b) + (tok1 + tok2) / b + 2 * Bc + 64;
}

struct FastArgs {
    Vocab vc;
    SumPlan plan;
    int32_t R;                         // rows per stage
    int32_t S;                         // ring depth (2..kFastMaxStages)
    int32_t ldr;                       // floats between staged rows (ldr / 4 odd: conflict-free LDS.128)
    int32_t rowbytes;                  // bytes copied per row (ld * 4, multiple of 16)
    int32_t common_iters;              // PL > 1: 8-float iterations of the shortest parallel leaf (lockstep part of the loop)
    int32_t _pad;
    unsigned long long negzero2;
    const StageRec *stages;
    const unsigned int *nstages;
    float *tiles;
    int64_t tile_stride;
    unsigned int *maxc;                // float bits, zeroed by the host
};

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One leaf block [start, start + len) of numpy's pairwise sum for the 2x4 cells (a_r, b_c), c = 4r + cc.
// half selects accumulators r[0..3] or r[4..7]; the two lanes of a pair meet in a reduce-scatter:
// afterwards lane `half` holds the four cells c = 2j + half (j = 0..3) in res[j], tail included.
// `common` (warp-uniform, >= 1) is the number of 8-float iterations every lane of the warp runs in
// lockstep; a lane whose leaf is longer finishes its extra iterations afterwards.  Lanes that run ahead
// of their neighbours (as the compiler's own unroll-remainder placement made the 84-float leaf do) shift
// their 32-byte window onto another leaf's banks: ncu showed 8 wavefronts per LDS.128 instead of 4.
__device__ __forceinline__ void leaf_2x4(const float *const (&a)[2], const float *const (&b)[4], int start, int len, int common,
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
    e += 8;
#pragma unroll 4
    for (int it = 1; it < common; ++it, e += 8) {                 // uniform trip count: the warp stays in lockstep
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
#pragma unroll 1
    for (; e < eend; e += 8) {                                     // the longer leaves' extra iterations
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
    const int tail = len & 7;                                      // the len % 8 tail: sequential adds, kept cells only
    if (tail == 4) {                                               // d = 100, 300, ...: one quad per row, packed sub / square
        const Q4 x0 = ldq(a[0] + eend), x1 = ldq(a[1] + eend), y0 = ldq(bk[0] + eend), y1 = ldq(bk[1] + eend);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const Q4 &x = (j >> 1) ? x1 : x0;
            const Q4 &y = (j & 1) ? y1 : y0;
            const f32x2 lo = sq2(sub2(x.lo, y.lo), nz), hi = sq2(sub2(x.hi, y.hi), nz);
            res[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(res[j], lo_f(lo)), hi_f(lo)), lo_f(hi)), hi_f(hi));
        }
    } else {
        for (int t = eend; t < start + len; ++t) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float s = __fsub_rn(a[j >> 1][t], bk[j & 1][t]);
                res[j] = __fadd_rn(res[j], __fmul_rn(s, s));
            }
        }
    }
}

// Distances of one tile task.  PL = 1: the lane pair walks the whole postfix program and each lane
// ends with the four cells c = 2j + half.  PL = 2 / 4: leaf l of a balanced tree is summed by lane
// pair l and the tree is closed by further reduce-scatter steps, leaving 2 / 1 cells per lane.
// out[k] is cell cell0 + k * cstep of the tile.
template <int PL>
__device__ __forceinline__ void dist_2x4(const SumPlan &plan, int common, f32x2 nz, const float *const (&a)[2], const float *const (&b)[4],
                                         int sub, float (&out)[4 / PL], int &cell0, int &cstep)
{
    const int half = sub & 1;
    if (PL == 1) {
        float st[kPlanDepth][4];
        int sp = 0;
        for (int o = 0; o < plan.nops; ++o) {
            float r[4];
            leaf_2x4(a, b, plan.start[o], plan.len[o], plan.len[o] >> 3, half, nz, r);
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
        for (int c = 0; c < 4 / PL; ++c) out[c] = __fsqrt_rn(st[0][c]);
        cell0 = half; cstep = 2;
    } else {
        const int l = sub >> 1, l0 = l & 1;
        float r[4];
        leaf_2x4(a, b, plan.start[l], plan.len[l], common, half, nz, r);
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

// One warp-wide batch of tile tasks [t0, t0 + 32 / (2 PL)) of a stage.
template <int PL>
__device__ __forceinline__ void run_desc_batch(const FastArgs &A, const unsigned char *stage, int ntiles, int t0)
{
    constexpr int LPT = 2 * PL;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT;
    int t = t0 + lane / LPT;
    const bool live = t < ntiles;
    if (!live) t = t0;                                   // clamp: recompute a valid tile, discard
    const uint4 w = *reinterpret_cast<const uint4 *>(stage + (size_t)t * 16);
    const float *rowsbuf = reinterpret_cast<const float *>(stage + kStageDescBytes);
    const float *a[2] = { rowsbuf + (w.x & 0xffu) * A.ldr, rowsbuf + ((w.x >> 8) & 0xffu) * A.ldr };
    const float *b[4] = { rowsbuf + ((w.x >> 16) & 0xffu) * A.ldr, rowsbuf + (w.x >> 24) * A.ldr,
                          rowsbuf + (w.y & 0xffu) * A.ldr, rowsbuf + ((w.y >> 8) & 0xffu) * A.ldr };
    const int na = (w.y >> 16) & 0xff, nb = w.y >> 24;
    float v[4 / PL];
    int cell0, cstep;
    dist_2x4<PL>(A.plan, A.common_iters, A.negzero2, a, b, sub, v, cell0, cstep);
    float mx = 0.f;
    const unsigned q = w.w & 0xffffu;
    if (live) {
        float *tile_p = A.tiles + (int64_t)q * A.tile_stride + (w.w >> 16);
        const int sr = w.z & 0xffff, sc = w.z >> 16;
#pragma unroll
        for (int k = 0; k < 4 / PL; ++k) {
            const int c = cell0 + k * cstep;
            const int r = c >> 2, cc = c & 3;
            if (r < na && cc < nb) {
                tile_p[r * sr + cc * sc] = v[k];
                mx = fmaxf(mx, v[k]);
            }
        }
    }
    unsigned mxb = __float_as_uint(mx);                  // distances are >= 0: uint order == float order
#pragma unroll
    for (int o = 1; o < LPT; o <<= 1) mxb = max(mxb, __shfl_xor_sync(kFull, mxb, o));
    if (live && sub == 0 && mxb) atomicMax(A.maxc + q, mxb);
}

template <int PL>
__global__ void __launch_bounds__(kFastThreads, 1)
cost_tiles_fast_kernel(const __grid_constant__ FastArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[kFastMaxStages], empty_bar[kFastMaxStages];
    __shared__ int s_ntiles[kFastMaxStages], s_taskctr[kFastMaxStages];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = A.S;
    const size_t stage_bytes = (size_t)kStageDescBytes + (size_t)A.R * A.ldr * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kFastConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nst = (int)*A.nstages;
    const int P = min(kFastProducers, S);                // a slot's next use may only be armed by a producer at most S stages ahead

    if (warp < kFastProducers) {
        // ------------------------------ producers ------------------------------
        if (warp >= P) return;
        int s = warp % S;                                          // slot and use count of stage `it`, kept incrementally
        uint32_t ph = (uint32_t)(warp / S) & 1u;
        for (int it = warp;; it += P) {
            const int idx = (int)blockIdx.x + it * (int)gridDim.x;
            if (idx >= nst) break;
            const StageRec *G = A.stages + idx;
            const int nrows = G->nrows, ntiles = G->ntiles;
            const int row0 = lane < nrows ? G->rows[lane] : 0;
            const int row1 = lane + 32 < nrows ? G->rows[lane + 32] : 0;
            const int row2 = lane + 64 < nrows ? G->rows[lane + 64] : 0;
            mbar_wait(&empty_bar[s], ph ^ 1u);                      // slot drained (passes at once the first time round)
            unsigned char *base = smem_raw + (size_t)s * stage_bytes;
            if (lane == 0) {
                s_ntiles[s] = ntiles; s_taskctr[s] = 0;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)nrows * (uint32_t)A.rowbytes + (uint32_t)ntiles * 16u);
                tma_row_g2s(smem_u32(base), G->tiles, (uint32_t)ntiles * 16u, &full_bar[s]);
            }
            __syncwarp();
            const uint32_t rows_u32 = smem_u32(base + kStageDescBytes);
            const uint32_t pitch = (uint32_t)A.ldr * 4u;
            if (lane < nrows)
                tma_row_g2s(rows_u32 + (uint32_t)lane * pitch, A.vc.table + (int64_t)row0 * A.vc.ld, (uint32_t)A.rowbytes, &full_bar[s]);
            if (lane + 32 < nrows)
                tma_row_g2s(rows_u32 + (uint32_t)(lane + 32) * pitch, A.vc.table + (int64_t)row1 * A.vc.ld, (uint32_t)A.rowbytes, &full_bar[s]);
            if (lane + 64 < nrows)
                tma_row_g2s(rows_u32 + (uint32_t)(lane + 64) * pitch, A.vc.table + (int64_t)row2 * A.vc.ld, (uint32_t)A.rowbytes, &full_bar[s]);
            s += P;
            while (s >= S) { s -= S; ph ^= 1u; }
        }
    } else {
        // ------------------------------ consumers ------------------------------
        constexpr int TPW = 32 / (2 * PL);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0;; ++it) {
            const int idx = (int)blockIdx.x + it * (int)gridDim.x;
            if (idx >= nst) break;
            mbar_wait(&full_bar[s], ph);
            const int ntiles = s_ntiles[s];
            const unsigned char *base = smem_raw + (size_t)s * stage_bytes;
            // the next batch is claimed before the current one runs, so the shared-memory atomic and the
            // broadcast are off the critical path (a consumer over-claims once per stage: harmless)
            int t0 = 0;
            if (lane == 0) t0 = atomicAdd(&s_taskctr[s], TPW);
            t0 = __shfl_sync(kFull, t0, 0);
            while (t0 < ntiles) {
                int tn = 0;
                if (lane == 0) tn = atomicAdd(&s_taskctr[s], TPW);
                run_desc_batch<PL>(A, base, ntiles, t0);
                t0 = __shfl_sync(kFull, tn, 0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 from the word-distance table (optional, wmd_set_distance_table): with D[V][V] resident (built once per
// embedding table by the kernels above, so every entry is bit-identical to what they would compute for
// the pair) a pair's tile is a gather of u1 x u2 floats -- 4 bytes per cell from HBM / L2 instead of
// 3 d rounded FP32 operations.  One warp per pair; the tile and its maximum land where K2b puts them.
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
    DocSide s1, s2;
    int64_t p0;
    int32_t npairs;
    int32_t _pad;
    const int32_t *rows1, *rows2, *u12;
    const float *D;
    int64_t V;
    float *tiles;
    int64_t tile_stride;
    unsigned int *maxc;
};

__global__ void __launch_bounds__(256)
cost_gather_kernel(const __grid_constant__ GatherArgs A)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }
    for (int q = gw; q < A.npairs; q += nw) {
        const int u = A.u12[q];
        const int u1 = u & 0xffff, u2 = u >> 16;
        if (u1 == 0 || u2 == 0) continue;
        int64_t a1, a2; int l;
        doc_span(A.s1, A.p0 + q, a1, l); doc_span(A.s2, A.p0 + q, a2, l);
        const int32_t *r1 = A.rows1 + slot_off(A.s1, tok1, q, a1), *r2 = A.rows2 + slot_off(A.s2, tok2, q, a2);
        float *tile = A.tiles + (int64_t)q * A.tile_stride;
        const int ncell = u1 * u2;
        const float inv = 1.0f / (float)u2;
        unsigned mx = 0;
        for (int c = lane; c < ncell; c += kWarp) {
            const int i = (int)(((float)c + 0.5f) * inv);         // c / u2: exact for c < 2^16, u2 <= 256
            const int j = c - i * u2;
            const float v = __ldg(A.D + (int64_t)__ldg(r1 + i) * A.V + __ldg(r2 + j));
            tile[c] = v;
            mx = max(mx, __float_as_uint(v));                      // distances are >= 0: uint order == float order
        }
        mx = __reduce_max_sync(kFull, mx);
        if (lane == 0) A.maxc[q] = mx;
    }
}

}  // namespace wmd
