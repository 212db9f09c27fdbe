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
// Structure (sm_100a), v4: every warp is an autonomous software pipeline; there is no producer
// warp and no CTA-wide barrier (v3 spent > 50 % of its samples spinning on the producer).
//   * a warp claims pairs from a global counter, cuts them into "units" (up to 4 blocks of
//     consecutive pairs, <= 16 tile tasks, <= 24 table rows) and streams the unit's rows
//     L2 -> shared memory in CHUNKS of a few dozen floats per row with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), one copy per row per chunk,
//     double-buffered per warp: chunk k+1 (or the first chunk of the next unit) is in flight
//     while chunk k is being summed.  A chunk never straddles a leaf of numpy's recursion.
//   * a tile task is a 2x4 block of cells owned by a lane pair: lane `half` keeps the accumulator
//     quad r[4h..4h+3] of the current leaf for all eight cells in registers ACROSS chunks, i.e.
//     6 LDS.128 per 96 float operations, which balances the 128 B/clk shared-memory port against
//     the FP32 pipe; leaf sums meet in a reduce-scatter (4 cells per lane) and the recursion tree
//     is a 4-deep register stack.
// The float math is issued as packed FADD2/FFMA2 (sub.f32x2, fma.f32x2 with a -0.0 addend that
// ptxas cannot see, add.f32x2): identical IEEE roundings, half the issue slots.
// Shared memory per warp is 2 x 24 rows x chunk pitch (~8 KB), so 16 warps fit in ~130 KB and the
// latency-bound solver (K3) can be co-resident on the same SM from the engine's other stream.
#pragma once
#include "common.cuh"

namespace wmd {

constexpr int kCostWarps = 4;             // warps per CTA (each one autonomous)
constexpr int kCostThreads = 32 * kCostWarps;
constexpr int kUnitSubs = 4;              // blocks (of consecutive pairs) per unit
constexpr int kUnitRows = 24;             // staged rows per unit
constexpr int kUnitTiles = 16;            // tile tasks per unit = lane pairs per warp
constexpr int kStackDepth = 4;            // register stack of leaf sums (d <= 1024)
constexpr int kMaxChunks = 96;
constexpr int kClaim = 8;                 // pairs claimed per atomic

struct CostChunk {
    uint16_t foff;                        // first float of the chunk inside a table row (multiple of 8)
    uint16_t bytes;                       // bytes copied per row (multiple of 16)
    uint16_t niter;                       // 8-float iterations (0 for a sequential leaf)
    uint8_t tail;                         // floats after the 8-wide part (only on the last chunk of a leaf)
    uint8_t flags;                        // 1 = first chunk of a leaf, 2 = last chunk, 4 = sequential leaf (len < 8)
    uint8_t adds;                         // stack pops after the leaf (only on its last chunk)
    uint8_t _p[3];
};

struct CostArgs {
    Vocab vc;
    DocSide s1, s2;                  // only .off / .L are used (document slots)
    int64_t p0;                      // first pair of this chunk of pairs
    int32_t npairs;                  // pairs in this launch
    int32_t nchunks;
    int32_t pitch;                   // bytes between staged rows (== 32 mod 64: conflict-free LDS.128)
    int32_t fast_R, fast_T;          // pairs that fit a planned stage (cost_fast.cuh) are skipped; 0 = take everything
    int32_t _pad;
    unsigned long long negzero2;     // 0x8000000080000000: (-0.0f, -0.0f), opaque to ptxas
    const int32_t *rows1, *rows2;    // from K1
    const int32_t *u12;
    float *tiles;                    // [npairs, tile_stride]
    int64_t tile_stride;
    unsigned int *maxc;              // [npairs] float bits, zeroed by the host before the launch
    unsigned int *counter;           // work-claim counter, zeroed by the host
    CostChunk chunks[kMaxChunks];
};

struct CostSub {
    int32_t q;                       // pair (launch-local)
    int32_t i0, ni, j0, nj;          // block of the pair's tile (doc1 rows x doc2 rows)
    int32_t u2;                      // tile row pitch
    int32_t tr;                      // 1: the 2-side of the 2x4 tile runs along doc2
    int32_t tilebase;                // first tile task of the block inside its unit
    int32_t rowbase;                 // first staged row
    int32_t _pad;
    int64_t o1, o2;                  // token-slot offsets of the pair
};

struct CostWarpState {
    CostSub subs[2][kUnitSubs];
    unsigned long long bar[2];
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 28)) __trap();                 // seconds: a protocol bug must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_row_g2s(uint32_t dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

__device__ __forceinline__ int unit_tiles(int ni, int nj, int tr)
{
    const int na = tr ? nj : ni, nb = tr ? ni : nj;
    return ((na + 1) >> 1) * ((nb + 3) >> 2);
}
__device__ __forceinline__ int pick_orientation(int ni, int nj)
{
    const int p0 = ((ni + 1) >> 1) * ((nj + 3) >> 2);              // tile tasks, 2-side along doc1
    const int p1 = ((nj + 1) >> 1) * ((ni + 3) >> 2);
    return p1 < p0 ? 1 : 0;
}

// A pair goes to the planned fast path (cost_fast.cuh) when its rows and tile tasks fit one stage.
__device__ __forceinline__ bool fast_fits(int u1, int u2, int R, int T)
{
    return u1 > 0 && u2 > 0 && u1 + u2 <= R && unit_tiles(u1, u2, pick_orientation(u1, u2)) <= T;
}

// One chunk of a leaf for the 2x4 cells (a_r, b_c), c = 4r + cc.  `half` selects accumulators
// r[0..3] or r[4..7].  acc persists in registers from the first chunk of a leaf to its last; on the
// last chunk the two lanes of a pair meet in a reduce-scatter and lane `half` receives the four
// cells c = 2j + half (j = 0..3) in res[j], tail included.  Returns true when res is valid.
__device__ __forceinline__ bool chunk_2x4(const CostChunk ck, const float *const (&a)[2], const float *const (&b)[4],
                                          int half, f32x2 nz, Q4 (&acc)[8], float (&res)[4])
{
    const float *bk[2] = { half ? b[1] : b[0], half ? b[3] : b[2] };      // columns of the kept cells: half, 2 + half
    if (ck.flags & 4) {                                // numpy: plain sequential loop (only when d < 8)
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = 0.f;
        for (int e = 0; e < ck.tail; ++e) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t = __fsub_rn(a[j >> 1][e], bk[j & 1][e]);
                res[j] = __fadd_rn(res[j], __fmul_rn(t, t));
            }
        }
        return true;
    }
    int e = 4 * half;
    const int eend = 8 * ck.niter;
    if (ck.flags & 1) {
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
        e += 8;
    }
#pragma unroll 2
    for (; e < eend; e += 8) {
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
    if (!(ck.flags & 2)) return false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float p0 = quad_sum(acc[2 * j]), p1 = quad_sum(acc[2 * j + 1]);
        const float mine = half ? p1 : p0, theirs = half ? p0 : p1;
        const float o = __shfl_xor_sync(kFull, theirs, 1);
        res[j] = __fadd_rn(mine, o);                               // (r0+r1+r2+r3) + (r4+..+r7); fadd commutes
    }
    for (int t = eend; t < eend + ck.tail; ++t) {                  // the len % 8 tail, sequential, kept cells only
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float s = __fsub_rn(a[j >> 1][t], bk[j & 1][t]);
            res[j] = __fadd_rn(res[j], __fmul_rn(s, s));
        }
    }
    return true;
}

// Splits a pair's u1 x u2 tile into nbi x nbj blocks of at most kUnitTiles tile tasks and
// kUnitRows rows each (balanced block sizes; the fewest blocks that fit).  Warp-uniform scalar code.
__device__ __forceinline__ void choose_blocks(int u1, int u2, int &BI, int &BJ, int &nbi, int &nbj)
{
    int best = 0x7fffffff;
    BI = 1; BJ = 1; nbi = u1; nbj = u2;
    for (int a = 1; a <= u1; ++a) {
        if (a >= best) break;
        const int bi = (u1 + a - 1) / a;
        for (int b = 1; b <= u2; ++b) {
            if (a * b >= best) break;
            const int bj = (u2 + b - 1) / b;
            if (bi + bj <= kUnitRows && unit_tiles(bi, bj, pick_orientation(bi, bj)) <= kUnitTiles) {
                best = a * b; BI = bi; BJ = bj; nbi = a; nbj = b;
                break;
            }
        }
        if (bi == 1) break;
    }
}

__global__ void __launch_bounds__(kCostThreads)
cost_tiles_kernel(const __grid_constant__ CostArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ CostWarpState wstate[kCostWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane & 1, lp = lane >> 1;
    CostWarpState &W = wstate[warp];
    const uint32_t bufbytes = (uint32_t)kUnitRows * (uint32_t)A.pitch;
    unsigned char *ring = smem_raw + (size_t)warp * 2 * bufbytes;
    const uint32_t ring_u32 = smem_u32(ring);

    if (lane == 0) {
        mbar_init(&W.bar[0], 1); mbar_init(&W.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    int64_t tok1, tok2;
    { int l; doc_span(A.s1, A.p0, tok1, l); doc_span(A.s2, A.p0, tok2, l); }

    // ---- pair / block iterator (warp-uniform) ----
    int qa = 0, qb = 0;                       // claimed pairs [qa, qb)
    bool have_pair = false;
    int pq = 0, pu1 = 0, pu2 = 0, BI = 1, BJ = 1, nbi = 0, nbj = 0, bi = 0, bj = 0;
    int64_t po1 = 0, po2 = 0;

    auto next_pair = [&]() -> bool {
        for (;;) {
            if (qa >= qb) {
                int q0 = 0;
                if (lane == 0) q0 = (int)atomicAdd(A.counter, (unsigned)kClaim);
                q0 = __shfl_sync(kFull, q0, 0);
                if (q0 >= A.npairs) return false;
                qa = q0; qb = min(A.npairs, q0 + kClaim);
            }
            const int q = qa++;
            const int u = __ldg(A.u12 + q);
            const int u1 = u & 0xffff, u2 = u >> 16;
            if (u1 == 0 || u2 == 0) continue;             // early-out pairs have no tile
            if (A.fast_R > 0 && fast_fits(u1, u2, A.fast_R, A.fast_T)) continue;
            pq = q; pu1 = u1; pu2 = u2;
            int64_t aa; int l;
            doc_span(A.s1, A.p0 + q, aa, l); po1 = slot_off(A.s1, tok1, q, aa);
            doc_span(A.s2, A.p0 + q, aa, l); po2 = slot_off(A.s2, tok2, q, aa);
            choose_blocks(u1, u2, BI, BJ, nbi, nbj);
            bi = 0; bj = 0;
            return true;
        }
    };

    // per-lane source row of the unit being issued, and unit geometry (warp-uniform)
    const float *srcrow = nullptr;
    int nsubsU[2] = { 0, 0 }, nrowsU[2] = { 0, 0 }, ntilesU[2] = { 0, 0 };

    auto build_unit = [&](int slot) -> bool {
        int nsubs = 0, rows = 0, tiles = 0;
        int myrow = -1;
        while (nsubs < kUnitSubs) {
            if (!have_pair) { have_pair = next_pair(); if (!have_pair) break; }
            const int i0 = bi * BI, j0 = bj * BJ;
            const int ni = min(BI, pu1 - i0), nj = min(BJ, pu2 - j0);
            const int tr = pick_orientation(ni, nj);
            const int nt = unit_tiles(ni, nj, tr);
            if (nsubs > 0 && (rows + ni + nj > kUnitRows || tiles + nt > kUnitTiles)) break;
            if (lane == 0) {
                CostSub sb;
                sb.q = pq; sb.i0 = i0; sb.ni = ni; sb.j0 = j0; sb.nj = nj; sb.u2 = pu2; sb.tr = tr;
                sb.tilebase = tiles; sb.rowbase = rows; sb._pad = 0; sb.o1 = po1; sb.o2 = po2;
                W.subs[slot][nsubs] = sb;
            }
            const int local = lane - rows;
            if (local >= 0 && local < ni + nj)
                myrow = local < ni ? __ldg(A.rows1 + po1 + i0 + local) : __ldg(A.rows2 + po2 + j0 + (local - ni));
            rows += ni + nj; tiles += nt; ++nsubs;
            if (++bj == nbj) { bj = 0; if (++bi == nbi) have_pair = false; }
        }
        nsubsU[slot] = nsubs; nrowsU[slot] = rows; ntilesU[slot] = tiles;
        if (nsubs == 0) return false;
        srcrow = A.vc.table + (int64_t)(myrow < 0 ? 0 : myrow) * A.vc.ld;
        __syncwarp();
        return true;
    };

    unsigned seq = 0;                         // chunks issued so far (buffer = seq & 1)
    auto issue = [&](int slot, int c) {
        const CostChunk ck = A.chunks[c];
        const int b = seq & 1;
        fence_proxy_async();                  // generic-proxy reads of this buffer precede the async-proxy writes
        if (lane == 0) mbar_arrive_expect_tx(&W.bar[b], (uint32_t)nrowsU[slot] * ck.bytes);
        __syncwarp();
        if (lane < nrowsU[slot])
            tma_row_g2s(ring_u32 + b * bufbytes + lane * A.pitch, srcrow + ck.foff, ck.bytes, &W.bar[b]);
        ++seq;
    };

    int cur = 0;
    if (!build_unit(cur)) return;
    issue(cur, 0);
    unsigned done = 0;                        // chunks consumed so far
    const int pitchf = A.pitch >> 2;

    for (;;) {
        // ---- this lane pair's tile task in the current unit ----
        const int ntiles = ntilesU[cur];
        const bool live = lp < ntiles;
        const int t = live ? lp : 0;
        int g = 0;
        while (g + 1 < nsubsU[cur] && W.subs[cur][g + 1].tilebase <= t) ++g;
        const CostSub sb = W.subs[cur][g];
        const int local = t - sb.tilebase;
        const int na = sb.tr ? sb.nj : sb.ni, nb = sb.tr ? sb.ni : sb.nj;
        const int abase = sb.rowbase + (sb.tr ? sb.ni : 0), bbase = sb.rowbase + (sb.tr ? 0 : sb.ni);
        const int TI = (na + 1) >> 1, TJ = (nb + 3) >> 2;
        const int ti = local / TJ, tj = local - ti * TJ;
        int ia[2], jb[4], ra[2], rb[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) { ia[r] = ti + r * TI; ra[r] = (abase + (ia[r] < na ? ia[r] : ti)) * pitchf; }
#pragma unroll
        for (int c = 0; c < 4; ++c) { jb[c] = tj + c * TJ; rb[c] = (bbase + (jb[c] < nb ? jb[c] : tj)) * pitchf; }

        Q4 acc[8];
        float st[kStackDepth][4];
        bool have_next = false;
        int nxt = cur ^ 1;
        for (int c = 0; c < A.nchunks; ++c) {
            // keep one chunk in flight behind the one being summed
            if (c + 1 < A.nchunks) issue(cur, c + 1);
            else { have_next = build_unit(nxt); if (have_next) issue(nxt, 0); }
            const int b = done & 1;
            mbar_wait(&W.bar[b], (done >> 1) & 1u);
            const float *buf = reinterpret_cast<const float *>(ring + (size_t)b * bufbytes);
            const float *a[2] = { buf + ra[0], buf + ra[1] };
            const float *bb[4] = { buf + rb[0], buf + rb[1], buf + rb[2], buf + rb[3] };
            const CostChunk ck = A.chunks[c];
            float res[4];
            if (chunk_2x4(ck, a, bb, half, A.negzero2, acc, res)) {
                // push the leaf sum, then pop-add `adds` times (static indexing: the stack top is st[0])
#pragma unroll
                for (int k = kStackDepth - 1; k > 0; --k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) st[k][j] = st[k - 1][j];
#pragma unroll
                for (int j = 0; j < 4; ++j) st[0][j] = res[j];
                for (int k = 0; k < ck.adds; ++k) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) st[0][j] = __fadd_rn(st[1][j], st[0][j]);
#pragma unroll
                    for (int m = 1; m < kStackDepth - 1; ++m)
#pragma unroll
                        for (int j = 0; j < 4; ++j) st[m][j] = st[m + 1][j];
                }
            }
            ++done;
            __syncwarp();                     // every lane is done reading buffer b before it is refilled
        }
        // ---- epilogue: distances, tile store, per-pair maximum ----
        float mx = 0.f;
        if (live) {
            float *tile_p = A.tiles + (int64_t)sb.q * A.tile_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cidx = 2 * j + half;
                const int r = cidx >> 2, cc = cidx & 3;
                const int iar = r ? ia[1] : ia[0];
                const int jbc = cc == 0 ? jb[0] : (cc == 1 ? jb[1] : (cc == 2 ? jb[2] : jb[3]));
                if (iar < na && jbc < nb) {
                    const float v = __fsqrt_rn(st[0][j]);
                    const int i = sb.i0 + (sb.tr ? jbc : iar);
                    const int jj = sb.j0 + (sb.tr ? iar : jbc);
                    tile_p[(int64_t)i * sb.u2 + jj] = v;
                    mx = fmaxf(mx, v);
                }
            }
        }
        const unsigned mxb = __float_as_uint(mx);            // distances are >= 0: uint order == float order
        for (int s = 0; s < nsubsU[cur]; ++s) {
            const unsigned m = __reduce_max_sync(kFull, (live && g == s) ? mxb : 0u);
            if (lane == 0 && m) atomicMax(A.maxc + W.subs[cur][s].q, m);
        }
        if (!have_next) break;
        cur = nxt;
    }
}

// init_sims(replace=True): v /= sqrt((v ** 2).sum(-1)) per row, float32, numpy summation order.
// One-off at table load (src/wmd.py:54); one thread per row.
constexpr int kPlanDepth = 8;
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
