// wmd_b200.cu -- host engine + C ABI of libwmd_b200.so (see include/wmd_b200.h).
//
// Default path (word-distance table resident): one persistent wmd_fused_small_kernel per chunk scores every pair of
// documents of <= 32 tokens end to end; what it leaves behind goes through list-mode nbow_pairs_kernel, list_sort_kernel
// (longest first) and the solver classes -- emd_solve_small_kernel and emd_solve_wide_kernel<1 .. 8> side by side on
// their own streams -- which read their costs from the same table.
// Direct path per chunk of pairs (chunks alternate between two streams so that the FP32-bound cost kernel of one chunk
// overlaps the latency-bound solver of the other and the chunk's tiles stay L2-resident between the two):
//     K1 nbow_pairs_kernel -> K2 cost_plan_kernel + cost_tiles_fast_kernel -> K3 emd_solve_small_kernel / emd_solve_wide_kernel
// There is no CPU implementation behind this ABI: without a device every entry fails.
#include "../../include/wmd_b200.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "common.cuh"
#include "nbow.cuh"
#include "cost.cuh"
#include "cost_fast.cuh"
#include "solve.cuh"
#include "solve_wide.cuh"
#include "fused.cuh"
#include "rwmd.cuh"
#include "allpairs.cuh"
#include "emd.cuh"

using namespace wmd;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(WMD_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return WMD_OK;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return fail(WMD_ECUDA, "cudaFree: %s", cudaGetErrorString(e)); }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(WMD_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return WMD_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Workspace {
    DevBuf ids1, ids2, off1, off2;              // staged inputs (host entries)
    DevBuf rows1, cnt1, ip1, rows2, cnt2, ip2;  // per token slot
    DevBuf u12, meta, pqn, extra, maxc;         // per pair
    DevBuf tiles, out, status, scratch, plan, wt1, wt2;
    DevBuf lb, l1, l2, am1, am2;                // rwmd outputs (host entry)
    DevBuf counters;                            // kCounterBytes: work-claim counters, stage / list counts
    DevBuf biglist;                             // table mode: pairs the fused kernel leaves to the general path
    DevBuf sorted;                              // the chunk's pairs, longest first (list_sort_kernel)
    // The wide solver classes of one chunk run side by side, each on its own stream with its own cost / flow scratch:
    // one launch per class in sequence would pay one tail per class, and a pair of the larger classes runs for milliseconds.
    DevBuf wscratch[8];
    cudaStream_t wstream[8] = {};
    cudaEvent_t wfork = nullptr, wjoin[8] = {};
    int ensure_wide_streams()
    {
        if (wfork) return WMD_OK;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);                     // hi = the numerically smallest = most urgent priority
        for (int k = 0; k < 8; ++k) {
            const int prio = std::max(hi, lo - k);                      // the larger the class, the earlier its blocks are placed
            if (cudaStreamCreateWithPriority(&wstream[k], cudaStreamNonBlocking, prio) != cudaSuccess) return fail(WMD_ECUDA, "cudaStreamCreateWithPriority failed");
            if (cudaEventCreateWithFlags(&wjoin[k], cudaEventDisableTiming) != cudaSuccess) return fail(WMD_ECUDA, "cudaEventCreate failed");
        }
        if (cudaEventCreateWithFlags(&wfork, cudaEventDisableTiming) != cudaSuccess) return fail(WMD_ECUDA, "cudaEventCreate failed");
        return WMD_OK;
    }
    void release()
    {
        sorted.release();
        for (int k = 0; k < 8; ++k) {
            wscratch[k].release();
            if (wstream[k]) cudaStreamDestroy(wstream[k]);
            if (wjoin[k]) cudaEventDestroy(wjoin[k]);
            wstream[k] = nullptr; wjoin[k] = nullptr;
        }
        if (wfork) cudaEventDestroy(wfork);
        wfork = nullptr;
        DevBuf *all[] = { &ids1, &ids2, &off1, &off2, &rows1, &cnt1, &ip1, &rows2, &cnt2, &ip2, &u12, &meta, &pqn,
                          &extra, &maxc, &tiles, &out, &status, &scratch, &plan, &wt1, &wt2, &lb, &l1, &l2, &am1, &am2, &counters, &biglist };
        for (DevBuf *b : all) b->release();
    }
};

struct ProfRec { int kind; cudaEvent_t a, b; };

// unsigned slots of Workspace::counters
constexpr size_t kCounterBytes = 128;
constexpr int kCtrFused = 18;                   // fused kernel's work-claim counter (slots 1..9: the solver classes)
constexpr int kCtrNBig = 19;                    // length of Workspace::biglist
constexpr int kCtrCost = 16;                    // general cost kernel's claim counter
constexpr int kCtrStages = 17;                  // planned stages of the fast cost path
constexpr int kCtrDmax = 20;                    // largest entry of the word-distance table (float bits)

}  // namespace

struct wmd_engine {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    float *table = nullptr;
    int64_t V = 0;
    int32_t d = 0, ld = 0;
    int32_t *map = nullptr;
    int64_t nmap = 0;
    int32_t *rank = nullptr;
    SumPlan plan;
    CostChunk cost_chunks[kMaxChunks];           // K2's chunk program (the plan's leaves cut into staged pieces)
    int32_t cost_nchunks = 0, cost_pitch = 0, cost_ctas_per_sm = 1;
    int32_t fast_R = 0, fast_S = 0, fast_ldr = 0, fast_PL = 1;   // planned fast path (cost_fast.cuh); fast_R == 0: disabled
    size_t fast_smem = 0;
    cudaStream_t streams[2] = { nullptr, nullptr };
    cudaEvent_t ev_fork = nullptr, ev_join[2] = { nullptr, nullptr }, ev_slot[2] = { nullptr, nullptr };
    bool slot_used[2] = { false, false };
    Workspace ws[2];
    // all-pairs mode (allpairs.cuh): V x V distance table (lazy) and a grow-only workspace
    float *dtab = nullptr;
    __half *dtab16 = nullptr;                    // all-pairs mode: the table rounded down to half precision (bounds only)
    bool use_dtab = false;                       // pair path takes its costs from dtab (default policy: on when the table fits dtab_budget)
    bool dtab_forced = false;                    // the caller asked for the table (wmd_set_distance_table(1), WMD_DTAB=1): failing to build it is an error
    float dmax = 0.f;
    size_t dtab_budget = 0;                      // bytes the table may take for the default policy to switch it on
    double dtab_build_ms = 0.0;                  // wall time of the one-off build
    bool ids_are_rows = false;                   // this call's ids are table rows although a token map is installed (WMD_IDS_ARE_ROWS)
    void *pin_in = nullptr, *pin_out = nullptr;  // wmd_pairs_submit / wmd_pairs_wait: handle-owned pinned staging
    size_t pin_in_cap = 0, pin_out_cap = 0;
    int64_t pending_pairs = -1;                  // pairs of the job in flight between submit and wait (-1: none)
    int host_chunk_first = 32768, host_chunk_max = 131072;  // host jobs in table mode: pairs of the first chunk, cap of the doubling schedule (WMD_HOST_CHUNK=first,max)
    int ap_r1_mult = 3;                          // all-pairs: round 1 solves ap_r1_mult * k candidates per row (WMD_AP_R1MULT)
    OutFan fan;                                  // wmd_set_fanout: the peers' result arrays (device pointers valid in this process)
    int fused_minb = 9;                          // fused kernel variant: __launch_bounds__(128, 8 / 9 / 10) = 64 / 56 / 48 registers (WMD_FUSED_MINB)
    int fused_blocks_per_sm = 0;                 // fused kernel: resident blocks per SM at the last smem size
    size_t fused_smem_cached = 0;
    bool fused_fan_cached = false;
    cudaStream_t ap_stream = nullptr;
    DevBuf ap[32];
    unsigned long long *stats = nullptr;        // device [6]
    bool profiling = false;
    int solve_blocks_per_sm = 8;                 // K3 grid cap per SM (WMD_SOLVE_BLOCKS): fewer leaves room for a co-resident K2b
    int fast_stage_cap = kFastMaxStages;         // K2b ring depth cap (WMD_FAST_S)
    int slot_mask = 1;                           // 0 (WMD_SERIAL=1): every chunk on one stream, for clean per-kernel timings
    std::vector<ProfRec> prof;
    double prof_ms[WMD_K_COUNT] = { 0 };
    int64_t prof_n[WMD_K_COUNT] = { 0 };
};

namespace {

void build_plan_rec(int start, int n, SumPlan &pl, bool &ok)
{
    if (n <= 128) {
        if (pl.nops >= kMaxPlanOps) { ok = false; return; }
        pl.start[pl.nops] = start; pl.len[pl.nops] = n; pl.adds[pl.nops] = 0; pl.nops++;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    build_plan_rec(start, n2, pl, ok);
    build_plan_rec(start + n2, n - n2, pl, ok);
    if (ok) pl.adds[pl.nops - 1]++;
}

// Cuts the plan's leaves into K2's staged chunks (<= max_iters 8-float iterations per chunk; the
// len % 8 tail rides on a leaf's last chunk) and sizes the per-warp ring.
int build_cost_chunks(wmd_engine *E, int max_iters)
{
    const SumPlan &pl = E->plan;
    int n = 0, depth = 0, maxdepth = 0, maxbytes = 16;
    for (int o = 0; o < pl.nops; ++o) {
        const int start = pl.start[o], len = pl.len[o];
        if (start % 8) return fail(WMD_EINVAL, "internal: leaf start %d is not a multiple of 8", start);
        if (len < 8) {
            if (n >= kMaxChunks) return fail(WMD_EINVAL, "embedding width %d needs too many chunks", E->d);
            CostChunk c{};
            c.foff = (uint16_t)start; c.bytes = (uint16_t)(((len * 4) + 15) & ~15); c.niter = 0; c.tail = (uint8_t)len;
            c.flags = 1 | 2 | 4; c.adds = (uint8_t)pl.adds[o];
            E->cost_chunks[n++] = c;
            maxbytes = std::max(maxbytes, (int)c.bytes);
        } else {
            const int iters = len / 8, tail = len % 8;
            const int pieces = (iters + max_iters - 1) / max_iters;
            int done = 0;
            for (int k = 0; k < pieces; ++k) {
                if (n >= kMaxChunks) return fail(WMD_EINVAL, "embedding width %d needs too many chunks", E->d);
                const int it = (iters - done + (pieces - k) - 1) / (pieces - k);      // balanced pieces
                const bool last = k == pieces - 1;
                CostChunk c{};
                c.foff = (uint16_t)(start + 8 * done);
                const int floats = 8 * it + (last ? tail : 0);
                c.bytes = (uint16_t)(((floats * 4) + 15) & ~15);
                c.niter = (uint16_t)it; c.tail = (uint8_t)(last ? tail : 0);
                c.flags = (uint8_t)((k == 0 ? 1 : 0) | (last ? 2 : 0));
                c.adds = (uint8_t)(last ? pl.adds[o] : 0);
                E->cost_chunks[n++] = c;
                maxbytes = std::max(maxbytes, (int)c.bytes);
                done += it;
            }
        }
        depth += 1; maxdepth = std::max(maxdepth, depth); depth -= pl.adds[o];
    }
    if (maxdepth > kStackDepth) return fail(WMD_EINVAL, "embedding width %d too large (summation tree deeper than %d)", E->d, kStackDepth);
    E->cost_nchunks = n;
    int pitch = (maxbytes + 63) & ~63;                       // == 32 (mod 64): LDS.128 of 4 adjacent rows hit 4 bank groups
    pitch += 32;
    E->cost_pitch = pitch;
    const size_t smem = (size_t)kCostWarps * 2 * kUnitRows * pitch;
    if (smem > E->smem_optin) return fail(WMD_EINVAL, "embedding width %d does not fit the shared-memory ring", E->d);
    cudaError_t e = cudaFuncSetAttribute(cost_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(WMD_ECUDA, "cudaFuncSetAttribute(cost_tiles_kernel): %s", cudaGetErrorString(e));
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cost_tiles_kernel, kCostThreads, smem);
    if (e != cudaSuccess || nb < 1) return fail(WMD_ECUDA, "cost_tiles_kernel cannot be resident (%s)", cudaGetErrorString(e));
    int cap = 4;                                             // 16 warps / SM: leaves room for the co-resident solver
    if (const char *v = getenv("WMD_COST_CTAS_PER_SM")) cap = std::max(1, atoi(v));
    E->cost_ctas_per_sm = std::min(nb, cap);
    return WMD_OK;
}

// Sizes the planned fast path's stage ring for this embedding width (cost_fast.cuh).
int setup_fast_path(wmd_engine *E)
{
    E->fast_R = 0;
    if (const char *v = getenv("WMD_COST_FAST")) { if (atoi(v) == 0) return WMD_OK; }
    int ldr4 = E->ld / 4;
    if ((ldr4 & 1) == 0) ldr4++;
    const int ldr = ldr4 * 4;
    const size_t pitch = (size_t)ldr * 4;
    const size_t budget = std::min<size_t>(E->smem_optin, 227 * 1024) - 1024;    // static barriers / counters
    int R = (int)std::min<size_t>(kStageRowsMax, (48 * 1024) / pitch);
    // wide embeddings: prefer stages that still hold a 20 + 20 token pair (or as close as three stages allow)
    // over a deeper ring -- pairs that do not fit a stage fall to the much slower general kernel
    if (R < 40 && budget / 3 > (size_t)kStageDescBytes + pitch)
        R = std::max(R, (int)std::min<size_t>(40, (budget / 3 - kStageDescBytes) / pitch));
    if (const char *v = getenv("WMD_FAST_R")) R = std::max(8, std::min(R, atoi(v)));
    if (R < 8) return WMD_OK;                                 // very wide embeddings: general kernel only
    const size_t stage_bytes = (size_t)kStageDescBytes + (size_t)R * pitch;
    int S = (int)std::min<size_t>(kFastMaxStages, budget / stage_bytes);
    if (const char *v = getenv("WMD_FAST_S")) S = std::max(2, std::min(S, atoi(v)));
    if (S < 2) return WMD_OK;
    const SumPlan &pl = E->plan;
    int PL = 1;
    bool minlen = true;
    for (int o = 0; o < pl.nops; ++o) minlen = minlen && pl.len[o] >= 8;
    if (minlen && pl.nops == 4 && pl.adds[0] == 0 && pl.adds[1] == 1 && pl.adds[2] == 0 && pl.adds[3] == 2) PL = 4;
    else if (minlen && pl.nops == 2 && pl.adds[0] == 0 && pl.adds[1] == 1) PL = 2;
    // lanes of one tile work on PL leaves at once: if the leaves start a multiple of 32 floats apart (d = 256, 512, ...)
    // all of them hit the same banks on every load -- walk the leaves one after the other instead
    if (PL > 1) {
        bool collide = false;
        for (int o = 1; o < PL; ++o) collide = collide || ((pl.start[o] - pl.start[0]) % 32 == 0);
        if (collide) PL = 1;
    }
    if (const char *v = getenv("WMD_FAST_PL")) { const int f = atoi(v); if (f == 1 || (f == 2 && PL >= 2) || (f == 4 && PL == 4)) PL = f; }
    const size_t smem = (size_t)S * stage_bytes;
    cudaError_t e;
    if (PL == 4) e = cudaFuncSetAttribute(cost_tiles_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else if (PL == 2) e = cudaFuncSetAttribute(cost_tiles_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else e = cudaFuncSetAttribute(cost_tiles_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(WMD_ECUDA, "cudaFuncSetAttribute(cost_tiles_fast_kernel): %s", cudaGetErrorString(e));
    E->fast_R = R; E->fast_S = S; E->fast_ldr = ldr; E->fast_PL = PL; E->fast_smem = smem;
    return WMD_OK;
}

struct Prof {
    wmd_engine *E; int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    Prof(wmd_engine *E_, int kind_, cudaStream_t st_) : E(E_), kind(kind_), st(st_)
    {
        if (E->profiling) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~Prof()
    {
        if (E->profiling) { cudaEventRecord(b, st); E->prof.push_back({ kind, a, b }); }
    }
};

bool takes_fused(const wmd_engine *E, bool solve, bool rwmd, int mode);

Vocab make_vocab(const wmd_engine *E)
{
    Vocab v;
    v.table = E->table; v.V = E->V; v.d = E->d; v.ld = E->ld; v.map = E->ids_are_rows ? nullptr : E->map; v.nmap = E->nmap; v.rank = E->rank;
    return v;
}

struct ChunkOut {
    double *out; int32_t *status;           // indexed by global pair number p
    // optional rwmd outputs
    bool rwmd = false;
    double *lb = nullptr, *l1 = nullptr, *l2 = nullptr;
    int32_t *am1 = nullptr, *am2 = nullptr; // chunk-relative token offsets
    bool solve = true;
    int mode = WMD_MODE_PYEMD;
    bool am_abs = false;                    // device entry: argmins at absolute CSR offsets
    OutFan fan;                             // pair entries in pyemd mode: further copies of out / status (wmd_set_fanout)
};

// K2: cost tiles of pairs [p0, p0 + Bc): the planned fast path for pairs that fit a stage, the
// general kernel for the rest.  rows1 / rows2 / u12 come from K1 (or are synthesised by the
// distance-table build); tiles is [Bc, tile_stride]; maxc receives the per-pair maximum (float bits).
int launch_cost(wmd_engine *E, Workspace &W, cudaStream_t st, const DocSide &s1, const DocSide &s2, int64_t p0, int32_t Bc,
                int32_t ml1, int32_t ml2, int64_t tokcap1, int64_t tokcap2, const int32_t *rows1, const int32_t *rows2,
                const int32_t *u12, float *tiles, int64_t tile_stride, unsigned int *maxc)
{
    int rc;
    if ((rc = W.counters.ensure(kCounterBytes))) return rc;
    if (E->use_dtab && E->dtab) {                        // tiles gathered from the word-distance table
        CK(cudaMemsetAsync(W.counters.p, 0, kCounterBytes, st));    // K3's work-claim counters
        CK(cudaMemsetAsync(maxc, 0, (size_t)Bc * 4, st));
        GatherArgs G;
        G.s1 = s1; G.s2 = s2; G.p0 = p0; G.npairs = Bc; G._pad = 0;
        G.rows1 = rows1; G.rows2 = rows2; G.u12 = u12;
        G.D = E->dtab; G.V = E->V; G.tiles = tiles; G.tile_stride = tile_stride; G.maxc = maxc;
        const int grid = (int)std::min<int64_t>(((int64_t)Bc + 7) / 8, (int64_t)E->sm_count * 8);
        Prof pr(E, WMD_K_COST, st);
        cost_gather_kernel<<<std::max(grid, 1), 256, 0, st>>>(G);
        CK(cudaGetLastError());
        return WMD_OK;
    }
    const Vocab vc = make_vocab(E);
    {
        CK(cudaMemsetAsync(W.counters.p, 0, kCounterBytes, st));
        CK(cudaMemsetAsync(maxc, 0, (size_t)Bc * 4, st));
        const int R = E->fast_R, T = kStageTilesMax;
        bool need_general = true;
        if (R > 0) {
            if ((rc = W.plan.ensure((size_t)plan_stage_bound(Bc, ml1, ml2, tokcap1, tokcap2, R, T) * sizeof(StageRec)))) return rc;
            PlanArgs P;
            P.s1 = s1; P.s2 = s2; P.p0 = p0; P.npairs = Bc; P.R = R; P.T = T; P._pad = 0;
            P.rows1 = rows1; P.rows2 = rows2; P.u12 = u12;
            P.stages = W.plan.as<StageRec>(); P.nstages = W.counters.as<unsigned int>() + kCtrStages;
            const int pblocks = (int)std::min<int64_t>(((int64_t)Bc + 127) / 128, (int64_t)E->sm_count * 16);
            {
                Prof pr(E, WMD_K_COST, st);
                cost_plan_kernel<<<pblocks, 128, 0, st>>>(P);
                CK(cudaGetLastError());
            }
            FastArgs F;
            F.vc = vc; F.plan = E->plan; F.R = R; F.S = E->fast_S; F.ldr = E->fast_ldr; F.rowbytes = E->ld * 4;
            F.negzero2 = 0x8000000080000000ull;
            F.common_iters = 1; F._pad = 0;
            if (E->fast_PL > 1) {
                int cm = 1 << 30;
                for (int o = 0; o < E->fast_PL; ++o) cm = std::min(cm, E->plan.len[o] >> 3);
                F.common_iters = std::max(cm, 1);
            }
            F.stages = P.stages; F.nstages = P.nstages;
            F.tiles = tiles; F.tile_stride = tile_stride; F.maxc = maxc;
            const int grid = (int)std::min<int64_t>(Bc, (int64_t)E->sm_count);
            Prof pr(E, WMD_K_COST, st);
            if (E->fast_PL == 4) cost_tiles_fast_kernel<4><<<grid, kFastThreads, E->fast_smem, st>>>(F);
            else if (E->fast_PL == 2) cost_tiles_fast_kernel<2><<<grid, kFastThreads, E->fast_smem, st>>>(F);
            else cost_tiles_fast_kernel<1><<<grid, kFastThreads, E->fast_smem, st>>>(F);
            CK(cudaGetLastError());
            need_general = false;                         // longer pairs are cut into blocks by the plan
        }
        if (need_general) {
            CostArgs A;
            A.vc = vc; A.s1 = s1; A.s2 = s2; A.p0 = p0; A.npairs = Bc;
            A.nchunks = E->cost_nchunks; A.pitch = E->cost_pitch; A.fast_R = R; A.fast_T = T; A._pad = 0;
            memcpy(A.chunks, E->cost_chunks, sizeof A.chunks);
            A.negzero2 = 0x8000000080000000ull;
            A.rows1 = rows1; A.rows2 = rows2; A.u12 = u12;
            A.tiles = tiles; A.tile_stride = tile_stride;
            A.maxc = maxc;
            A.counter = W.counters.as<unsigned int>() + kCtrCost;
            const size_t smem = (size_t)kCostWarps * 2 * kUnitRows * A.pitch;
            const int grid = (int)std::min<int64_t>(((int64_t)Bc + kCostWarps - 1) / kCostWarps, (int64_t)E->sm_count * E->cost_ctas_per_sm);
            Prof pr(E, WMD_K_COST, st);
            cost_tiles_kernel<<<grid, kCostThreads, smem, st>>>(A);
            CK(cudaGetLastError());
        }
    }
    return WMD_OK;
}

// K3 launches of one chunk: one launch per solver class over all pairs of the chunk, or (list mode) over the pairs the
// fused kernel left behind.  gather: costs come from the word-distance table instead of cost tiles.
// shortest longest document of a chunk that can produce a residual problem of KC column words: the shorter side has more
// than 32 (KC - 1) nodes (KC = 1: the longer side more than 32, one of which may be the dummy)
inline int wide_min_ml(int kc) { return kc == 1 ? 32 : 32 * (kc - 1) + 1; }

template <class K>
int launch_wide_solver(wmd_engine *E, DevBuf &scratch, cudaStream_t st, K kernel, SolveArgs &S, size_t per_warp, int32_t Bc)
{
    int rc;
    const int wpb = 4;
    const size_t smem = per_warp * wpb;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, wpb * 32, smem));
    if (nb < 1) return fail(WMD_ECUDA, "solver class %d cannot be resident", S.cls);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)Bc + wpb - 1) / wpb, (int64_t)E->sm_count * nb));
    if ((rc = scratch.ensure((size_t)grid * wpb * solve_wide_scratch_ints_per_warp(S.mr, S.ldc) * 4))) return rc;      // per warp: quantised costs + flow + supplies
    S.scratch = scratch.as<int32_t>();
    kernel<<<grid, wpb * 32, smem, st>>>(S);
    CK(cudaGetLastError());
    return WMD_OK;
}

int launch_solvers(wmd_engine *E, Workspace &W, cudaStream_t st, const DocSide &s1, const DocSide &s2, int64_t p0, int32_t Bc, int ML,
                   const PairWork &pw, const float *tiles, int64_t tile_stride, const ChunkOut &O, const int32_t *list,
                   const unsigned int *nlist, bool gather)
{
    int rc;
    SolveArgs S;
    S.s1 = s1; S.s2 = s2; S.p0 = p0; S.npairs = Bc;
    S.ip1 = pw.ip1; S.ip2 = pw.ip2; S.u12 = pw.u12; S.meta = pw.meta; S.pqn = pw.pqn; S.extra = pw.extra;
    S.tiles = tiles; S.tile_stride = tile_stride; S.maxc = W.maxc.as<float>();
    S.out = O.out; S.status = O.status; S.fan = O.fan;
    S.list = list; S.nlist = nlist;
    S.D = gather ? E->dtab : nullptr; S.V = E->V; S.rows1 = pw.rows1; S.rows2 = pw.rows2; S.maxc_w = W.maxc.as<float>();
    S.scratch = nullptr;
    // Fork: the wide classes (largest first) each on their own stream, class A on the chunk's stream next to them.  Chunks
    // whose longest document allows two wide classes at most (<= 64 tokens) stay on the chunk's stream: their pairs are
    // short-lived, and the kernels measured slower side by side (64-token pairs 15.9 -> 18.2 ms per 2^18).
    const bool wide = ML >= wide_min_ml(3);
    Prof pr(E, WMD_K_SOLVE, st);
    if (wide) {
        if ((rc = W.ensure_wide_streams()) || (rc = W.sorted.ensure((size_t)Bc * 4))) return rc;
        ListSortArgs L;
        L.list = list; L.nlist = nlist; L.npairs = Bc; L.meta = pw.meta; L.sorted = W.sorted.as<int32_t>();
        list_sort_kernel<<<1, 1024, 0, st>>>(L);
        CK(cudaGetLastError());
        S.list = L.sorted;
        CK(cudaEventRecord(W.wfork, st));
    }
    // Larger problems, largest class first.  The per-launch capacity follows the chunk's longest document; the instance
    // only depends on the pair's own shorter side.
    for (int kc = 8; kc >= 1; --kc) {
        if (ML < wide_min_ml(kc)) continue;           // no document of the chunk is long enough for a problem of this class
        S.cls = kClsW1 + kc - 1;
        S.mr = std::min(kc == 8 ? kMaxDocLen + 1 : kMaxDocLen, ML + 1); S.mc = 32 * kc; S.ldc = 32 * kc;
        S.counter = W.counters.as<unsigned int>() + S.cls;
        cudaStream_t ws = wide ? W.wstream[kc - 1] : st;
        if (wide) CK(cudaStreamWaitEvent(ws, W.wfork, 0));
#define WMD_WIDE(KC, MINB)                                                                                                           \
        (gather ? launch_wide_solver(E, W.wscratch[KC - 1], ws, emd_solve_wide_kernel<KC, true, MINB>, S, solve_wide_smem_per_warp<KC>(S.mr), Bc)      \
                : launch_wide_solver(E, W.wscratch[KC - 1], ws, emd_solve_wide_kernel<KC, false, MINB>, S, solve_wide_smem_per_warp<KC>(S.mr), Bc))
        switch (kc) {
        case 1: rc = WMD_WIDE(1, 8); break;
        case 2: rc = WMD_WIDE(2, 8); break;
        case 3: rc = WMD_WIDE(3, 8); break;
        case 4: rc = WMD_WIDE(4, 8); break;
        case 5: rc = WMD_WIDE(5, 6); break;
        case 6: rc = WMD_WIDE(6, 6); break;
        case 7: rc = WMD_WIDE(7, 5); break;
        default: rc = WMD_WIDE(8, 5); break;
        }
#undef WMD_WIDE
        if (rc) return rc;
        if (wide) CK(cudaEventRecord(W.wjoin[kc - 1], ws));
    }
    {                                                 // class A: both sides of the residual problem fit one word
        S.cls = kClsA;
        S.mr = S.mc = std::min(32, ML + 1);           // class A may turn the problem round: the dummy then is a row
        S.ldc = S.mc | 1;
        S.counter = W.counters.as<unsigned int>() + kClsA;
        const int wpb = 4;
        const size_t smem = solve_small_smem_per_warp(S.mr, S.mc, S.ldc) * wpb;
        const int blocks_per_sm = E->solve_blocks_per_sm * 2;                         // __launch_bounds__(128, 9)
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)Bc + 8 * wpb - 1) / (8 * wpb), (int64_t)E->sm_count * blocks_per_sm));
        if (gather) {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(emd_solve_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            emd_solve_small_kernel<true><<<grid, wpb * 32, smem, st>>>(S);
        } else {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(emd_solve_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            emd_solve_small_kernel<false><<<grid, wpb * 32, smem, st>>>(S);
        }
        CK(cudaGetLastError());
    }
    for (int kc = 8; wide && kc >= 1; --kc)            // join
        if (ML >= wide_min_ml(kc)) CK(cudaStreamWaitEvent(st, W.wjoin[kc - 1], 0));
    return WMD_OK;
}

int launch_nbow_pairs(wmd_engine *E, cudaStream_t st, const DocSide &s1, const DocSide &s2, int64_t p0, int32_t Bc, int ML, const PairWork &pw,
                      const ChunkOut &O, const int32_t *list, const unsigned int *nlist)
{
    const int Lp = ML;
    const int wpb = Lp <= 64 ? 8 : 4;
    const size_t smem = nbow_smem_per_warp(Lp) * wpb;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(nbow_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((Bc + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
    Prof pr(E, WMD_K_NBOW, st);
    nbow_pairs_kernel<<<grid, wpb * 32, smem, st>>>(s1, s2, make_vocab(E), p0, Bc, Lp, pw, O.out, O.status, list, nlist);
    CK(cudaGetLastError());
    return WMD_OK;
}

// Table mode (fused.cuh): one persistent warp-per-pair kernel does K1 -> K2 -> K3 for every pair of <= 32 tokens per
// side; whatever it leaves behind goes through list-mode K1 and the gather-mode solvers.  No cost tiles exist, so the
// launch geometry does not depend on the longest document of the chunk beyond the per-warp matrix capacity.
int run_chunk_fused(wmd_engine *E, Workspace &W, cudaStream_t st, const DocSide &s1, const DocSide &s2,
                    int64_t p0, int32_t Bc, int64_t tokcap1, int64_t tokcap2, int ML, const ChunkOut &O)
{
    int rc;
    const bool big = ML >= 32;                           // a pair can only be left behind when a side reaches 32 tokens
    if ((rc = W.counters.ensure(kCounterBytes)) || (rc = W.biglist.ensure((size_t)Bc * 4))) return rc;
    CK(cudaMemsetAsync(W.counters.p, 0, kCounterBytes, st));
    FusedArgs F;
    F.s1 = s1; F.s2 = s2; F.vc = make_vocab(E); F.D = E->dtab; F.p0 = p0; F.npairs = Bc;
    F.cap = std::min(32, ML + 1); F.ldc = F.cap | 1; F._pad = 0;
    F.counter = W.counters.as<unsigned int>() + kCtrFused;
    F.biglist = W.biglist.as<int32_t>(); F.nbig = W.counters.as<unsigned int>() + kCtrNBig;
    F.stats = E->stats; F.out = O.out; F.status = O.status; F.fan = O.fan;
    {
        const int wpb = 4;
        const size_t smem = fused_smem_per_warp(F.cap, F.ldc) * wpb;
        const bool fan = F.fan.n > 0;
        auto kern = fan ? wmd_fused_small_kernel<9, true>
                        : E->fused_minb >= 10 ? wmd_fused_small_kernel<10, false> : E->fused_minb == 9 ? wmd_fused_small_kernel<9, false> : wmd_fused_small_kernel<8, false>;
        if (smem != E->fused_smem_cached || fan != E->fused_fan_cached) {
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, wpb * 32, smem));
            if (nb < 1) return fail(WMD_ECUDA, "wmd_fused_small_kernel cannot be resident");
            E->fused_blocks_per_sm = nb; E->fused_smem_cached = smem; E->fused_fan_cached = fan;
        }
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)Bc + 8 * wpb - 1) / (8 * wpb), (int64_t)E->sm_count * E->fused_blocks_per_sm));
        Prof pr(E, WMD_K_FUSED, st);
        kern<<<grid, wpb * 32, smem, st>>>(F);
        CK(cudaGetLastError());
    }
    if (!big) return WMD_OK;
    if ((rc = W.rows1.ensure(tokcap1 * 4)) || (rc = W.cnt1.ensure(tokcap1 * 4)) || (rc = W.ip1.ensure(tokcap1 * 4)) ||
        (rc = W.rows2.ensure(tokcap2 * 4)) || (rc = W.cnt2.ensure(tokcap2 * 4)) || (rc = W.ip2.ensure(tokcap2 * 4)) ||
        (rc = W.u12.ensure((size_t)Bc * 4)) || (rc = W.meta.ensure((size_t)Bc * 4)) || (rc = W.pqn.ensure((size_t)Bc * 8)) ||
        (rc = W.extra.ensure((size_t)Bc * 8)) || (rc = W.maxc.ensure((size_t)Bc * 4)))
        return rc;
    PairWork pw;
    pw.rows1 = W.rows1.as<int32_t>(); pw.cnt1 = W.cnt1.as<int32_t>(); pw.ip1 = W.ip1.as<int32_t>();
    pw.rows2 = W.rows2.as<int32_t>(); pw.cnt2 = W.cnt2.as<int32_t>(); pw.ip2 = W.ip2.as<int32_t>();
    pw.u12 = W.u12.as<int32_t>(); pw.meta = W.meta.as<int32_t>(); pw.pqn = W.pqn.as<double>(); pw.extra = W.extra.as<double>();
    pw.stats = E->stats; pw.exact = 0; pw._pad = 0; pw.wt1 = nullptr; pw.wt2 = nullptr; pw.fan = O.fan;
    const int32_t *list = W.biglist.as<int32_t>();
    const unsigned int *nlist = W.counters.as<unsigned int>() + kCtrNBig;
    if ((rc = launch_nbow_pairs(E, st, s1, s2, p0, Bc, ML, pw, O, list, nlist))) return rc;
    return launch_solvers(E, W, st, s1, s2, p0, Bc, ML, pw, nullptr, 0, O, list, nlist, true);
}

// Launch K1..K3 for pairs [p0, p0 + Bc) on stream st. tok caps bound the token slots of the chunk.
int run_chunk(wmd_engine *E, Workspace &W, cudaStream_t st, const DocSide &s1, const DocSide &s2,
              int64_t p0, int32_t Bc, int64_t tokcap1, int64_t tokcap2, int32_t ml1, int32_t ml2,
              const ChunkOut &O)
{
    if (Bc <= 0) return WMD_OK;
    ml1 = std::max(ml1, 1); ml2 = std::max(ml2, 1);
    const int ML = std::max(ml1, ml2);
    if (takes_fused(E, O.solve, O.rwmd, O.mode))
        return run_chunk_fused(E, W, st, s1, s2, p0, Bc, tokcap1, tokcap2, ML, O);
    const int64_t tile_stride = (int64_t)ml1 * ml2;
    int rc;
    if ((rc = W.rows1.ensure(tokcap1 * 4)) || (rc = W.cnt1.ensure(tokcap1 * 4)) || (rc = W.ip1.ensure(tokcap1 * 4)) ||
        (rc = W.rows2.ensure(tokcap2 * 4)) || (rc = W.cnt2.ensure(tokcap2 * 4)) || (rc = W.ip2.ensure(tokcap2 * 4)) ||
        (rc = W.u12.ensure((size_t)Bc * 4)) || (rc = W.meta.ensure((size_t)Bc * 4)) || (rc = W.pqn.ensure((size_t)Bc * 8)) ||
        (rc = W.extra.ensure((size_t)Bc * 8)) || (rc = W.maxc.ensure((size_t)Bc * 4)) ||
        (rc = W.tiles.ensure((size_t)Bc * tile_stride * 4)) || (rc = W.counters.ensure(kCounterBytes)))
        return rc;

    // ---- K1
    PairWork pw;
    pw.rows1 = W.rows1.as<int32_t>(); pw.cnt1 = W.cnt1.as<int32_t>(); pw.ip1 = W.ip1.as<int32_t>();
    pw.rows2 = W.rows2.as<int32_t>(); pw.cnt2 = W.cnt2.as<int32_t>(); pw.ip2 = W.ip2.as<int32_t>();
    pw.u12 = W.u12.as<int32_t>(); pw.meta = W.meta.as<int32_t>(); pw.pqn = W.pqn.as<double>(); pw.extra = W.extra.as<double>();
    pw.stats = E->stats;
    pw.exact = O.mode == WMD_MODE_EXACT; pw._pad = 0; pw.wt1 = nullptr; pw.wt2 = nullptr;
    if (!pw.exact && O.solve && !O.rwmd) pw.fan = O.fan;
    if (pw.exact) {
        if ((rc = W.wt1.ensure((size_t)tokcap1 * 8)) || (rc = W.wt2.ensure((size_t)tokcap2 * 8))) return rc;
        pw.wt1 = W.wt1.as<double>(); pw.wt2 = W.wt2.as<double>();
    }
    if ((rc = launch_nbow_pairs(E, st, s1, s2, p0, Bc, ML, pw, O, nullptr, nullptr))) return rc;
    // ---- K2
    if ((rc = launch_cost(E, W, st, s1, s2, p0, Bc, ml1, ml2, tokcap1, tokcap2, pw.rows1, pw.rows2, pw.u12, W.tiles.as<float>(), tile_stride,
                          W.maxc.as<unsigned int>())))
        return rc;
    // ---- K5 (optional)
    if (O.rwmd) {
        RwmdArgs R;
        R.s1 = s1; R.s2 = s2; R.p0 = p0; R.npairs = Bc; R.Lp = ML;
        R.cnt1 = pw.cnt1; R.cnt2 = pw.cnt2; R.u12 = pw.u12;
        R.tiles = W.tiles.as<float>(); R.tile_stride = tile_stride; R.status = O.status;
        R.lb = O.lb; R.l1 = O.l1; R.l2 = O.l2; R.argmin_rows = O.am1; R.argmin_cols = O.am2; R.am_abs = O.am_abs ? 1 : 0; R._pad = 0;
        const int wpb = 8;
        const size_t smem = (size_t)wpb * ML * 8;
        const int grid = (int)std::min<int64_t>((Bc + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
        Prof pr(E, WMD_K_RWMD, st);
        rwmd_pairs_kernel<<<grid, wpb * 32, smem, st>>>(R);
        CK(cudaGetLastError());
    }
    if (!O.solve) return WMD_OK;
    // ---- K3 (work counters were zeroed before K2)
    if (O.mode == WMD_MODE_EXACT) {
        ExactArgs X;
        X.s1 = s1; X.s2 = s2; X.p0 = p0; X.npairs = Bc;
        X.mr = ml1; X.mc = ml2; X.ldc = ml2 | 1;
        X.u12 = pw.u12; X.wt1 = pw.wt1; X.wt2 = pw.wt2;
        X.tiles = W.tiles.as<float>(); X.tile_stride = tile_stride; X.maxc = W.maxc.as<float>();
        X.counter = W.counters.as<unsigned int>() + kClsA;
        X.out = O.out; X.status = O.status;
        const int wpb = 4;
        const size_t per_warp = (exact_smem_per_warp(X.mr, X.mc) + 15) & ~(size_t)15;
        const size_t smem = per_warp * wpb;
        const int grid = (int)std::min<int64_t>(((int64_t)Bc + 8 * wpb - 1) / (8 * wpb), (int64_t)E->sm_count * (ML > 64 ? 1 : 4));
        if ((rc = W.scratch.ensure((size_t)grid * wpb * 2 * X.mr * X.ldc * 8))) return rc;
        X.scratch = W.scratch.as<double>();
        Prof pr(E, WMD_K_SOLVE, st);
        if (ml2 <= 32) emd_solve_exact_kernel<1><<<grid, wpb * 32, smem, st>>>(X);
        else if (ml2 <= 64) emd_solve_exact_kernel<2><<<grid, wpb * 32, smem, st>>>(X);
        else emd_solve_exact_kernel<8><<<grid, wpb * 32, smem, st>>>(X);
        CK(cudaGetLastError());
        return WMD_OK;
    }
    return launch_solvers(E, W, st, s1, s2, p0, Bc, ML, pw, W.tiles.as<float>(), tile_stride, O, nullptr, nullptr, false);
}

int ensure_dtab(wmd_engine *E, cudaStream_t st);

// a job of this kind runs on the fused table-mode path (run_chunk_fused): no cost tiles, no per-length chunk limit
bool takes_fused(const wmd_engine *E, bool solve, bool rwmd, int mode)
{
    return E->use_dtab && E->dtab && solve && !rwmd && mode == WMD_MODE_PYEMD;
}

// Pairs per chunk.  Direct path: what 768 MB of cost tiles hold (at most 65 536: the tile descriptors carry 16-bit pair
// numbers).  Table mode has no tiles: a host job still goes in chunks of 65 536 so that the copies of one chunk overlap
// the kernels of another; a device job takes up to 2^20 pairs per launch of the persistent fused kernel -- one tail per
// million pairs instead of sixteen.
// First scoring call of a handle whose policy wants the table.  Under the default policy a table that cannot be built
// (memory taken since wmd_create) is not an error: the handle falls back to the direct path for good.
int lazy_dtab(wmd_engine *E)
{
    if (!E->use_dtab || E->dtab) return WMD_OK;
    const int rc = ensure_dtab(E, E->streams[0]);
    if (rc == WMD_ENOMEM && !E->dtab_forced) { E->use_dtab = false; return WMD_OK; }
    return rc;
}

int64_t chunk_pairs(int ml1, int ml2, bool fused, bool host_job = true)
{
    if (fused) return host_job ? 65536 : (1 << 20);
    const int64_t tile = (int64_t)std::max(ml1, 1) * std::max(ml2, 1) * 4;
    int64_t c = (int64_t)(768ll << 20) / tile;
    c = std::min<int64_t>(c, 65536);
    return std::max<int64_t>(c, 256);
}

// max document length + monotonicity of offsets [0, n] (branch-free main loop: the host compiler vectorises it)
int scan_offsets(const int64_t *off, int64_t n, int32_t &maxlen, const char *name)
{
    int64_t mx = 0, mn = 0;
    for (int64_t p = 0; p < n; ++p) {
        const int64_t l = off[p + 1] - off[p];
        mx = l > mx ? l : mx;
        mn = l < mn ? l : mn;
    }
    if (mn < 0) {
        for (int64_t p = 0; p < n; ++p)
            if (off[p + 1] < off[p]) return fail(WMD_EINVAL, "%s offsets are not monotone at %lld", name, (long long)p);
    }
    if (n > 0 && off[0] < 0) return fail(WMD_EINVAL, "%s offsets start below zero", name);
    if (mx > WMD_MAX_DOC_LEN) return fail(WMD_EINVAL, "%s holds a document of %lld tokens; the limit is %d", name, (long long)mx, WMD_MAX_DOC_LEN);
    maxlen = (int32_t)mx;
    return WMD_OK;
}

int set_device(wmd_engine *E)
{
    CK(cudaSetDevice(E->device));
    return WMD_OK;
}

int reset_stats(wmd_engine *E, cudaStream_t st)
{
    CK(cudaMemsetAsync(E->stats, 0, 6 * sizeof(unsigned long long), st));
    return WMD_OK;
}

// shared body of the host entries: chunked H2D -> kernels -> D2H on the two internal streams
struct HostJob {
    const int32_t *ids1; const int64_t *off1; const int32_t *ids2; const int64_t *off2; int64_t npairs;
    double *out; int32_t *status;
    bool rwmd = false, solve = true;
    int mode = WMD_MODE_PYEMD;
    double *lb = nullptr, *l1 = nullptr, *l2 = nullptr; int32_t *am1 = nullptr, *am2 = nullptr;
    bool out_on_device = false;             // out / status are DEVICE arrays (wmd_pairs_host_in_dev_out)
};

int enqueue_host_job(wmd_engine *E, const HostJob &J)
{
    int rc;
    if ((rc = lazy_dtab(E))) return rc;                          // first call: build the word-distance table
    if ((rc = reset_stats(E, E->streams[0]))) return rc;
    CK(cudaEventRecord(E->ev_fork, E->streams[0]));
    CK(cudaStreamWaitEvent(E->streams[1], E->ev_fork, 0));
    // The offsets are validated and measured chunk by chunk, while the GPU works on the chunks already
    // queued: one pass over all of them up front cost 1.5 ms of a 16 ms call (16 MB the caller just wrote).
    int slot = 0;
    int64_t Bnext = 0, nchunk = 0;
    const bool fused = takes_fused(E, J.solve, J.rwmd, J.mode);
    for (int64_t c0 = 0; c0 < J.npairs; c0 += Bnext, slot = (slot ^ 1) & E->slot_mask) {
        // table mode: the first chunks are small so that the first kernel starts early, later ones larger (fewer kernel tails)
        const int64_t want = fused ? std::min<int64_t>(E->host_chunk_max, (int64_t)E->host_chunk_first << std::min<int64_t>(nchunk, 8)) : 65536;
        ++nchunk;
        const int64_t Btry = std::min<int64_t>(want, J.npairs - c0);
        int32_t ml1, ml2;
        if ((rc = scan_offsets(J.off1 + c0, Btry, ml1, "side 1")) || (rc = scan_offsets(J.off2 + c0, Btry, ml2, "side 2"))) return rc;
        const int32_t Bc = fused ? (int32_t)Btry : (int32_t)std::min<int64_t>(Btry, chunk_pairs(ml1, ml2, false));   // a shorter prefix keeps the same bounds
        Bnext = Bc;
        Workspace &W = E->ws[slot];
        cudaStream_t st = E->streams[slot];
        const int64_t t1 = J.off1[c0 + Bc] - J.off1[c0], t2 = J.off2[c0 + Bc] - J.off2[c0];
        if ((rc = W.ids1.ensure((size_t)std::max<int64_t>(t1, 1) * 4)) || (rc = W.ids2.ensure((size_t)std::max<int64_t>(t2, 1) * 4)) ||
            (rc = W.off1.ensure((size_t)(Bc + 1) * 8)) || (rc = W.off2.ensure((size_t)(Bc + 1) * 8)) ||
            (rc = W.out.ensure((size_t)Bc * 8)) || (rc = W.status.ensure((size_t)Bc * 4)))
            return rc;
        if (t1) CK(cudaMemcpyAsync(W.ids1.p, J.ids1 + J.off1[c0], (size_t)t1 * 4, cudaMemcpyHostToDevice, st));
        if (t2) CK(cudaMemcpyAsync(W.ids2.p, J.ids2 + J.off2[c0], (size_t)t2 * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(W.off1.p, J.off1 + c0, (size_t)(Bc + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(W.off2.p, J.off2 + c0, (size_t)(Bc + 1) * 8, cudaMemcpyHostToDevice, st));
        DocSide s1{}, s2{};
        s1.ids = W.ids1.as<int32_t>() - J.off1[c0]; s1.off = W.off1.as<int64_t>();
        s2.ids = W.ids2.as<int32_t>() - J.off2[c0]; s2.off = W.off2.as<int64_t>();
        ChunkOut O;
        O.out = W.out.as<double>(); O.status = W.status.as<int32_t>(); O.solve = J.solve; O.rwmd = J.rwmd; O.mode = J.mode;
        if (J.solve && !J.rwmd && J.mode == WMD_MODE_PYEMD) {        // the chunk's kernels index from 0: the peers' arrays start at the chunk
            O.fan = E->fan;
            for (int k = 0; k < O.fan.n; ++k) { O.fan.out[k] += c0; O.fan.status[k] += c0; }
        }
        if (J.rwmd) {
            if ((rc = W.lb.ensure((size_t)Bc * 8)) || (rc = W.l1.ensure((size_t)Bc * 8)) || (rc = W.l2.ensure((size_t)Bc * 8)) ||
                (rc = W.am1.ensure((size_t)std::max<int64_t>(t1, 1) * 4)) || (rc = W.am2.ensure((size_t)std::max<int64_t>(t2, 1) * 4)))
                return rc;
            O.lb = W.lb.as<double>(); O.l1 = W.l1.as<double>(); O.l2 = W.l2.as<double>();
            O.am1 = J.am1 ? W.am1.as<int32_t>() : nullptr; O.am2 = J.am2 ? W.am2.as<int32_t>() : nullptr;
            if (O.am1) CK(cudaMemsetAsync(W.am1.p, 0xff, (size_t)std::max<int64_t>(t1, 1) * 4, st));
            if (O.am2) CK(cudaMemsetAsync(W.am2.p, 0xff, (size_t)std::max<int64_t>(t2, 1) * 4, st));
        }
        if ((rc = run_chunk(E, W, st, s1, s2, 0, Bc, std::max<int64_t>(t1, 1), std::max<int64_t>(t2, 1), ml1, ml2, O))) return rc;
        const cudaMemcpyKind back = J.out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        if (J.out) CK(cudaMemcpyAsync(J.out + c0, W.out.p, (size_t)Bc * 8, back, st));
        if (J.status) CK(cudaMemcpyAsync(J.status + c0, W.status.p, (size_t)Bc * 4, back, st));
        if (J.rwmd) {
            if (J.lb) CK(cudaMemcpyAsync(J.lb + c0, W.lb.p, (size_t)Bc * 8, cudaMemcpyDeviceToHost, st));
            if (J.l1) CK(cudaMemcpyAsync(J.l1 + c0, W.l1.p, (size_t)Bc * 8, cudaMemcpyDeviceToHost, st));
            if (J.l2) CK(cudaMemcpyAsync(J.l2 + c0, W.l2.p, (size_t)Bc * 8, cudaMemcpyDeviceToHost, st));
            if (J.am1 && t1) CK(cudaMemcpyAsync(J.am1 + J.off1[c0], W.am1.p, (size_t)t1 * 4, cudaMemcpyDeviceToHost, st));
            if (J.am2 && t2) CK(cudaMemcpyAsync(J.am2 + J.off2[c0], W.am2.p, (size_t)t2 * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaEventRecord(E->ev_slot[slot], st));
        E->slot_used[slot] = true;
    }
    return WMD_OK;
}

// Waits for everything the engine's two streams hold.  Also the single exit of a failed host job: copies into the
// caller's buffers that are already queued must have landed before the call reports its error.
int drain_streams(wmd_engine *E)
{
    for (Workspace &W : E->ws)                          // an error exit between fork and join leaves the class streams on their own
        for (cudaStream_t ws : W.wstream) if (ws) cudaStreamSynchronize(ws);
    const cudaError_t e0 = cudaStreamSynchronize(E->streams[0]), e1 = cudaStreamSynchronize(E->streams[1]);
    E->slot_used[0] = E->slot_used[1] = false;
    if (e0 != cudaSuccess || e1 != cudaSuccess)
        return fail(WMD_ECUDA, "stream synchronisation failed: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    return WMD_OK;
}

int check_host_job(wmd_engine *E, const HostJob &J)
{
    int rc;
    if ((rc = set_device(E))) return rc;
    if (E->pending_pairs >= 0) return fail(WMD_EINVAL, "a submitted job is still in flight: call wmd_pairs_wait first");
    if (J.npairs < 0 || (J.npairs > 0 && (!J.off1 || !J.off2))) return fail(WMD_EINVAL, "null offsets");
    if (J.npairs > 0 && ((J.off1[J.npairs] > J.off1[0] && !J.ids1) || (J.off2[J.npairs] > J.off2[0] && !J.ids2))) return fail(WMD_EINVAL, "null ids");
    return WMD_OK;
}

int run_host_job(wmd_engine *E, const HostJob &J)
{
    const auto t_job0 = std::chrono::steady_clock::now();
    int rc;
    if ((rc = check_host_job(E, J))) return rc;
    if (J.npairs == 0) return WMD_OK;
    rc = enqueue_host_job(E, J);
    const auto t_enq = std::chrono::steady_clock::now();
    if (rc) {                                                    // keep the job's own message
        const std::string msg = g_err;
        drain_streams(E);
        g_err = msg;
        return rc;
    }
    if ((rc = drain_streams(E))) return rc;
    if (getenv("WMD_TRACE")) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[wmd] host job %lld pairs: scan+enqueue %.3f ms, wait %.3f ms\n", (long long)J.npairs,
                std::chrono::duration<double, std::milli>(t_enq - t_job0).count(), std::chrono::duration<double, std::milli>(t_end - t_enq).count());
    }
    return WMD_OK;
}

// shared body of the device entries: fork from the caller's stream, chunk, join back; no host sync
int run_dev_job(wmd_engine *E, const DocSide &s1, const DocSide &s2, int64_t total1, int64_t total2,
                int32_t ml1, int32_t ml2, int64_t npairs, const ChunkOut &O0, cudaStream_t us)
{
    int rc;
    if ((rc = set_device(E))) return rc;
    if (E->pending_pairs >= 0) return fail(WMD_EINVAL, "a submitted job is still in flight: call wmd_pairs_wait first");
    if (npairs < 0 || (O0.solve && !O0.out) || (O0.rwmd && !O0.lb)) return fail(WMD_EINVAL, "bad arguments");
    if (npairs == 0) return WMD_OK;
    if (ml1 < 0 || ml2 < 0 || ml1 > WMD_MAX_DOC_LEN || ml2 > WMD_MAX_DOC_LEN)
        return fail(WMD_EINVAL, "max_len must be within [0, %d]", WMD_MAX_DOC_LEN);
    ml1 = std::max(ml1, 1); ml2 = std::max(ml2, 1);
    if ((rc = lazy_dtab(E))) return rc;                          // first call only (host-synchronous once)
    CK(cudaEventRecord(E->ev_fork, us));
    CK(cudaStreamWaitEvent(E->streams[0], E->ev_fork, 0));
    CK(cudaStreamWaitEvent(E->streams[1], E->ev_fork, 0));
    if ((rc = reset_stats(E, E->streams[0]))) return rc;
    CK(cudaEventRecord(E->ev_join[0], E->streams[0]));
    CK(cudaStreamWaitEvent(E->streams[1], E->ev_join[0], 0));        // stats reset precedes both streams' kernels
    int64_t CH = chunk_pairs(ml1, ml2, takes_fused(E, O0.solve, O0.rwmd, O0.mode), false);
    if (const char *v = getenv("WMD_DEV_CHUNK")) CH = std::max<int64_t>(1024, std::min<int64_t>(CH, atoll(v)));
    int slot = 0;
    for (int64_t c0 = 0; c0 < npairs; c0 += CH, slot = (slot ^ 1) & E->slot_mask) {
        const int32_t Bc = (int32_t)std::min<int64_t>(CH, npairs - c0);
        Workspace &W = E->ws[slot];
        ChunkOut O = O0;
        if (!O.status) {
            // scratch status indexed from 0: shift so that status[c0 + q] lands in the scratch buffer
            if ((rc = W.status.ensure((size_t)Bc * 4))) return rc;
            O.status = W.status.as<int32_t>() - c0;
        }
        if (!O.out) {                                                // bounds only: K1 still writes its early-outs somewhere
            if ((rc = W.out.ensure((size_t)Bc * 8))) return rc;
            O.out = W.out.as<double>() - c0;
        }
        const int64_t cap1 = std::max<int64_t>(1, std::min<int64_t>(total1, (int64_t)Bc * ml1));
        const int64_t cap2 = std::max<int64_t>(1, std::min<int64_t>(total2, (int64_t)Bc * ml2));
        if ((rc = run_chunk(E, W, E->streams[slot], s1, s2, c0, Bc, cap1, cap2, ml1, ml2, O))) return rc;
    }
    CK(cudaEventRecord(E->ev_join[0], E->streams[0]));
    CK(cudaEventRecord(E->ev_join[1], E->streams[1]));
    CK(cudaStreamWaitEvent(us, E->ev_join[0], 0));
    CK(cudaStreamWaitEvent(us, E->ev_join[1], 0));
    return WMD_OK;
}

int run_dev_pairs(wmd_engine *E, const DocSide &s1, const DocSide &s2, int64_t total1, int64_t total2, int32_t ml1, int32_t ml2,
                  int64_t npairs, double *out, int32_t *status, cudaStream_t us, int mode = WMD_MODE_PYEMD)
{
    if (!out && npairs > 0) return fail(WMD_EINVAL, "out is null");
    ChunkOut O;
    O.out = out; O.status = status; O.mode = mode;
    if (mode == WMD_MODE_PYEMD && status) O.fan = E->fan;         // indexed like out / status (scratch status has no peers)
    return run_dev_job(E, s1, s2, total1, total2, ml1, ml2, npairs, O, us);
}

// ------------------------------------------------------------------------------------------------
// all-pairs mode
// ------------------------------------------------------------------------------------------------
enum {
    AP_IDSA, AP_OFFA, AP_IDSB, AP_OFFB, AP_ROWSA, AP_CNTA, AP_UNIQA, AP_NVALA, AP_ROWSB, AP_CNTB, AP_UNIQB, AP_NVALB,
    AP_ZB, AP_ZA, AP_LB, AP_KTH, AP_THR, AP_COUNTS, AP_OFFS, AP_CI, AP_CJ, AP_CD, AP_CST, AP_TOPJ, AP_TOPD, AP_KCUR,
    AP_BIJ, AP_LISTA, AP_LISTB, AP_COUNT_
};
static_assert(AP_COUNT_ <= 32, "wmd_engine::ap too small");

// D[a][b] for every two table rows, from the K2 kernels run on blocks of consecutive rows.
int ensure_dtab(wmd_engine *E, cudaStream_t st)
{
    if (E->dtab) return WMD_OK;
    int rc;
    // one-off and host-synchronous: the build borrows the slot-0 workspace, so nothing of an earlier asynchronous
    // device call may still be in flight on it
    CK(cudaDeviceSynchronize());
    const auto t_build0 = std::chrono::steady_clock::now();
    const int64_t V = E->V;
    const size_t bytes = (size_t)V * V * 4;
    size_t freeb = 0, totalb = 0;
    CK(cudaMemGetInfo(&freeb, &totalb));
    if (bytes + (4ull << 30) > freeb)
        return fail(WMD_ENOMEM, "all-pairs mode needs a %lld x %lld float32 distance table (%.1f GB); %.1f GB free",
                    (long long)V, (long long)V, bytes / 1e9, freeb / 1e9);
    float *D = nullptr;
    if (cudaMalloc(&D, bytes) != cudaSuccess) return fail(WMD_ENOMEM, "cudaMalloc(distance table) failed");
    const int BS = E->fast_R > 0 ? std::min(20, E->fast_R / 2) : 16;
    const int nb = (int)((V + BS - 1) / BS);
    const int64_t npairs = (int64_t)nb * (nb + 1) / 2;
    const int64_t CH = 65536;
    Workspace &W = E->ws[0];
    const int64_t tile_stride = (int64_t)BS * BS;
    if ((rc = W.rows1.ensure((size_t)CH * BS * 4)) || (rc = W.rows2.ensure((size_t)CH * BS * 4)) || (rc = W.u12.ensure((size_t)CH * 4)) ||
        (rc = W.maxc.ensure((size_t)CH * 4)) || (rc = W.tiles.ensure((size_t)CH * tile_stride * 4)) || (rc = E->ap[AP_BIJ].ensure((size_t)CH * 8))) {
        cudaFree(D);
        return rc;
    }
    DocSide s1{}, s2{};
    s1.L = BS; s2.L = BS;                                  // padded layout: work slots at q * BS
    for (int64_t q0 = 0; q0 < npairs; q0 += CH) {
        const int32_t Bc = (int32_t)std::min<int64_t>(CH, npairs - q0);
        dtab_make_pairs_kernel<<<(Bc + 255) / 256, 256, 0, st>>>((int32_t)V, BS, nb, q0, Bc, W.rows1.as<int32_t>(), W.rows2.as<int32_t>(),
                                                                 W.u12.as<int32_t>(), E->ap[AP_BIJ].as<int32_t>());
        if (cudaGetLastError() != cudaSuccess) { cudaFree(D); return fail(WMD_ECUDA, "dtab_make_pairs_kernel launch failed"); }
        if ((rc = launch_cost(E, W, st, s1, s2, 0, Bc, BS, BS, (int64_t)Bc * BS, (int64_t)Bc * BS, W.rows1.as<int32_t>(), W.rows2.as<int32_t>(), W.u12.as<int32_t>(),
                              W.tiles.as<float>(), tile_stride, W.maxc.as<unsigned int>()))) { cudaFree(D); return rc; }
        dtab_scatter_kernel<<<Bc, 128, 0, st>>>((int32_t)V, BS, Bc, W.u12.as<int32_t>(), E->ap[AP_BIJ].as<int32_t>(), W.tiles.as<float>(),
                                                tile_stride, D);
        if (cudaGetLastError() != cudaSuccess) { cudaFree(D); return fail(WMD_ECUDA, "dtab_scatter_kernel launch failed"); }
    }
    // largest distance in the table: scales the pruning margin (allpairs.cuh)
    unsigned int *dmaxbits = W.counters.as<unsigned int>() + kCtrDmax;
    if (cudaMemsetAsync(dmaxbits, 0, 4, st) != cudaSuccess) { cudaFree(D); return fail(WMD_ECUDA, "memset failed"); }
    table_max_kernel<<<E->sm_count * 4, 256, 0, st>>>(D, (int64_t)V * V, dmaxbits);
    unsigned int hb = 0;
    if (cudaMemcpyAsync(&hb, dmaxbits, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaFree(D);
        return fail(WMD_ECUDA, "distance table build failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    memcpy(&E->dmax, &hb, 4);
    E->dtab = D;
    E->dtab_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
    return WMD_OK;
}

// all-pairs mode: the half-precision copy of the table the bound kernels read
int ensure_dtab16(wmd_engine *E, cudaStream_t st)
{
    if (E->dtab16) return WMD_OK;
    const int64_t n = (int64_t)E->V * E->V;
    __half *p = nullptr;
    if (cudaMalloc(&p, (size_t)n * 2) != cudaSuccess) return fail(WMD_ENOMEM, "cudaMalloc(half-precision distance table) failed");
    dtab_to_half_kernel<<<E->sm_count * 8, 256, 0, st>>>(E->dtab, n, p);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(p); return fail(WMD_ECUDA, "dtab_to_half_kernel failed"); }
    E->dtab16 = p;
    return WMD_OK;
}

// rows / counts / uniq / nval and the packed (row, weight) lists of ndocs documents already on the device
int ap_nbow(wmd_engine *E, cudaStream_t st, const int32_t *ids_dev, const int64_t *off_dev, int64_t ndocs, int32_t ml, int64_t total,
            DevBuf &rows, DevBuf &cnt, DevBuf &uniq, DevBuf &nval, DevBuf &lists)
{
    int rc;
    const size_t tb = (size_t)std::max<int64_t>(total, 1);
    if ((rc = rows.ensure(tb * 4)) || (rc = cnt.ensure(tb * 4)) || (rc = uniq.ensure((size_t)ndocs * 4)) || (rc = nval.ensure((size_t)ndocs * 4)) ||
        (rc = lists.ensure(tb * 8)))
        return rc;
    DocSide s{};
    s.ids = ids_dev; s.off = off_dev;
    const int Lp = std::max(ml, 1);
    const int wpb = Lp <= 64 ? 8 : 4;
    const size_t smem = nbow_smem_per_warp(Lp) * wpb;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(nbow_docs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((ndocs + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
    Prof pr(E, WMD_K_NBOW, st);
    nbow_docs_kernel<<<grid, wpb * 32, smem, st>>>(s, make_vocab(E), (int32_t)ndocs, Lp, rows.as<int32_t>(), cnt.as<int32_t>(), nullptr,
                                                  uniq.as<int32_t>(), nval.as<int32_t>());
    CK(cudaGetLastError());
    pack_lists_kernel<<<(unsigned)((ndocs + 255) / 256), 256, 0, st>>>(rows.as<int32_t>(), cnt.as<int32_t>(), off_dev, uniq.as<int32_t>(),
                                                                        nval.as<int32_t>(), ndocs, lists.as<int2>());
    CK(cudaGetLastError());
    return WMD_OK;
}

// Phase timer of the all-pairs entry: events are only recorded while the job is queued and read once at the end, so
// timing never stalls the stream (the previous version synchronised at every phase boundary).
struct ApPhases {
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<int> phase;                                    // phase of the interval that ENDS at ev[i] (ev[0]: start)
    explicit ApPhases(cudaStream_t s) : st(s) { mark(-1); }
    ~ApPhases() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    void mark(int ph) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); phase.push_back(ph); }
    void collect(double *ms, int n)                            // call after the stream has been synchronised
    {
        for (size_t i = 1; i < ev.size(); ++i) {
            float t = 0.f;
            if (phase[i] >= 0 && phase[i] < n && cudaEventElapsedTime(&t, ev[i - 1], ev[i]) == cudaSuccess) ms[phase[i]] += t;
        }
    }
};

// exact WMD of the candidate pairs (ci[p] in A, cj[p] in B) -> cd[p]
int ap_exact(wmd_engine *E, cudaStream_t st, const int32_t *idsA, const int64_t *offA, int32_t mlA, const int32_t *idsB,
             const int64_t *offB, int32_t mlB, const int32_t *ci, const int32_t *cj, int64_t n, double *cd, int32_t *cst)
{
    if (n == 0) return WMD_OK;
    DocSide s1{}, s2{};
    s1.ids = idsA; s1.off = offA; s1.sel = ci; s1.slot = std::max(mlA, 1);
    s2.ids = idsB; s2.off = offB; s2.sel = cj; s2.slot = std::max(mlB, 1);
    // all-pairs mode owns the word-distance table (its bounds are built from it), so the candidates go through the
    // table-mode path -- the fused warp-per-pair kernel -- whatever the handle's policy for the pair entries is
    const bool prev = E->use_dtab;
    E->use_dtab = E->dtab != nullptr;
    const int rc = run_dev_pairs(E, s1, s2, n * (int64_t)std::max(mlA, 1), n * (int64_t)std::max(mlB, 1), mlA, mlB, n, cd, cst, st);
    E->use_dtab = prev;
    return rc;
}

int run_allpairs(wmd_engine *E, const int32_t *idsA, const int64_t *offA, int64_t nA, const int32_t *idsB, const int64_t *offB, int64_t nB,
                 int32_t k, int64_t row_begin, int64_t row_end, int32_t *out_idx, double *out_dist, bool out_on_device,
                 int64_t *stats, double *ms)
{
    int rc;
    if ((rc = set_device(E))) return rc;
    if (E->pending_pairs >= 0) return fail(WMD_EINVAL, "a submitted job is still in flight: call wmd_pairs_wait first");
    if (nA < 0 || nB < 0 || k <= 0 || row_begin < 0 || row_end > nA || row_begin > row_end) return fail(WMD_EINVAL, "bad all-pairs arguments");
    if (nA > 0x7fffffff || nB > 0x7fffffff) return fail(WMD_EINVAL, "too many documents");
    if (k > nB) return fail(WMD_EINVAL, "k = %d exceeds the %lld documents of set B", k, (long long)nB);
    if (k > 1024) return fail(WMD_EINVAL, "k is limited to 1024");
    const int64_t nR = row_end - row_begin;
    if (nR == 0) return WMD_OK;
    if (!offA || !offB || !out_idx || !out_dist) return fail(WMD_EINVAL, "null argument");
    E->ids_are_rows = false;
    int32_t mlA = 0, mlB = 0;
    if ((rc = scan_offsets(offA + row_begin, nR, mlA, "set A"))) return rc;
    if ((rc = scan_offsets(offB, nB, mlB, "set B"))) return rc;
    cudaStream_t st = E->ap_stream;
    DevBuf *B = E->ap;
    double ms_local[4] = { 0, 0, 0, 0 };                       // distance table, corpus index (Z_B), bounds + selection, exact solves
    int64_t st_local[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };          // lb pairs, round-1 solves, round-2 solves, query blocks

    {
        const auto t0 = std::chrono::steady_clock::now();
        const bool fresh = !E->dtab || !E->dtab16;
        if ((rc = ensure_dtab(E, st)) || (rc = ensure_dtab16(E, st))) return rc;
        if (fresh) ms_local[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    ApPhases tm(st);

    // ---- documents to the device; nBOW of both sets ------------------------------------------------
    const int64_t baseA = offA[row_begin], totA = offA[row_end] - baseA, baseB = offB[0], totB = offB[nB] - baseB;
    if ((totA > 0 && !idsA) || (totB > 0 && !idsB)) return fail(WMD_EINVAL, "null ids");
    if ((rc = B[AP_IDSA].ensure((size_t)std::max<int64_t>(totA, 1) * 4)) || (rc = B[AP_OFFA].ensure((size_t)(nR + 1) * 8)) ||
        (rc = B[AP_IDSB].ensure((size_t)std::max<int64_t>(totB, 1) * 4)) || (rc = B[AP_OFFB].ensure((size_t)(nB + 1) * 8)))
        return rc;
    if (totA) CK(cudaMemcpyAsync(B[AP_IDSA].p, idsA + baseA, (size_t)totA * 4, cudaMemcpyHostToDevice, st));
    if (totB) CK(cudaMemcpyAsync(B[AP_IDSB].p, idsB + baseB, (size_t)totB * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[AP_OFFA].p, offA + row_begin, (size_t)(nR + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[AP_OFFB].p, offB, (size_t)(nB + 1) * 8, cudaMemcpyHostToDevice, st));
    // offsets stay absolute; rebase the id pointers instead
    const int32_t *dIdsA = B[AP_IDSA].as<int32_t>() - baseA, *dIdsB = B[AP_IDSB].as<int32_t>() - baseB;
    const int64_t *dOffA = B[AP_OFFA].as<int64_t>(), *dOffB = B[AP_OFFB].as<int64_t>();
    if ((rc = ap_nbow(E, st, dIdsA, dOffA, nR, mlA, totA + baseA, B[AP_ROWSA], B[AP_CNTA], B[AP_UNIQA], B[AP_NVALA], B[AP_LISTA]))) return rc;
    if ((rc = ap_nbow(E, st, dIdsB, dOffB, nB, mlB, totB + baseB, B[AP_ROWSB], B[AP_CNTB], B[AP_UNIQB], B[AP_NVALB], B[AP_LISTB]))) return rc;
    // ---- Z_B[w][j] ----------------------------------------------------------------------------------
    const int64_t ldzb = (nB + kLbTile - 1) / kLbTile * kLbTile;
    if ((rc = B[AP_ZB].ensure((size_t)E->V * ldzb * 2))) return rc;
    {
        ZArgs Z;
        Z.D16 = E->dtab16; Z.V = (int32_t)E->V; Z.rows = B[AP_ROWSB].as<int32_t>(); Z.off = dOffB; Z.uniq = B[AP_UNIQB].as<int32_t>();
        Z.doc0 = 0; Z.ndocs = (int32_t)nB; Z.Z = B[AP_ZB].as<__half>(); Z.ldz = ldzb;
        dim3 grid((unsigned)((nB + 31) / 32), (unsigned)((E->V + kZWords - 1) / kZWords));
        z_build16_kernel<<<grid, dim3(32, 32), 0, st>>>(Z);
        CK(cudaGetLastError());
    }
    tm.mark(1);

    // ---- query blocks ---------------------------------------------------------------------------------
    // Query rows per block: as many as a bound matrix of <= 8 GiB (and <= a quarter of the free memory) holds, at most
    // 16 384, and all blocks of a call equally large -- every block costs two host round trips (its candidate counts) and
    // a tail per kernel, so a rank of an 8-GPU job (12 500 rows of 100 000) runs as ONE block.
    int64_t IB;
    {
        size_t freeb = 0, totalb = 0;
        CK(cudaMemGetInfo(&freeb, &totalb));
        const size_t budget = std::min<size_t>((size_t)8 << 30, (freeb + B[AP_LB].cap) / 4);
        int64_t cap = std::max<int64_t>(kLbTile, (int64_t)(budget / ((size_t)ldzb * 4)) / kLbTile * kLbTile);
        cap = std::min<int64_t>(cap, 16384);
        if (const char *v = getenv("WMD_AP_BLOCK")) cap = std::max<int64_t>(kLbTile, std::min<int64_t>(cap, atoi(v) / kLbTile * kLbTile));
        const int64_t nblocks = (nR + cap - 1) / cap;
        IB = ((nR + nblocks - 1) / nblocks + kLbTile - 1) / kLbTile * kLbTile;
    }
    const int64_t ldza = IB, ldlb = ldzb;
    if ((rc = B[AP_ZA].ensure((size_t)E->V * ldza * 2)) || (rc = B[AP_LB].ensure((size_t)IB * ldlb * 4)) || (rc = B[AP_KTH].ensure((size_t)IB * 4)) ||
        (rc = B[AP_THR].ensure((size_t)IB * 4)) || (rc = B[AP_COUNTS].ensure((size_t)IB * kCandWarps * 4)) || (rc = B[AP_OFFS].ensure((size_t)(IB * kCandWarps + 1) * 8)) ||
        (rc = B[AP_TOPJ].ensure((size_t)IB * k * 4)) || (rc = B[AP_TOPD].ensure((size_t)IB * k * 8)) || (rc = B[AP_KCUR].ensure((size_t)IB * 4)))
        return rc;
    // Round 1 solves the k1 = ap_r1_mult * k documents with the smallest bounds, not just k: the k-th exact distance among
    // more candidates is a tighter threshold for round 2 (k1 = k: 273 exact solves per row of the 100k x 100k job; see
    // profiles/README.md for the sweep).
    const int32_t k1 = (int32_t)std::min<int64_t>(nB, std::min<int64_t>(256, (int64_t)E->ap_r1_mult * k));
    const size_t lb_smem = (size_t)kLbTile * kLbPitch * 4;
    CK(cudaFuncSetAttribute(lb_tile16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb_smem));
    const cudaMemcpyKind back = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (int64_t i0 = 0; i0 < nR; i0 += IB) {
        const int32_t ni = (int32_t)std::min<int64_t>(IB, nR - i0);
        st_local[3] += 1;
        {
            ZArgs Z;
            Z.D16 = E->dtab16; Z.V = (int32_t)E->V; Z.rows = B[AP_ROWSA].as<int32_t>(); Z.off = dOffA; Z.uniq = B[AP_UNIQA].as<int32_t>();
            Z.doc0 = i0; Z.ndocs = ni; Z.Z = B[AP_ZA].as<__half>(); Z.ldz = ldza;
            dim3 grid((unsigned)((ni + 31) / 32), (unsigned)((E->V + kZWords - 1) / kZWords));
            z_build16_kernel<<<grid, dim3(32, 32), 0, st>>>(Z);
            CK(cudaGetLastError());
            LbArgs L;
            L.listA = B[AP_LISTA].as<int2>(); L.offA = dOffA; L.uniqA = B[AP_UNIQA].as<int32_t>(); L.nvalA = B[AP_NVALA].as<int32_t>();
            L.i0 = i0; L.ni = ni;
            L.listB = B[AP_LISTB].as<int2>(); L.offB = dOffB; L.uniqB = B[AP_UNIQB].as<int32_t>(); L.nvalB = B[AP_NVALB].as<int32_t>();
            L.nB = (int32_t)nB;
            L.ZB = B[AP_ZB].as<__half>(); L.ldzb = ldzb; L.ZA = B[AP_ZA].as<__half>(); L.ldza = ldza;
            L.LB = B[AP_LB].as<float>(); L.ldlb = ldlb;
            dim3 g2((unsigned)((ni + kLbTile - 1) / kLbTile), (unsigned)(ldzb / kLbTile));
            lb_tile16_kernel<<<g2, 32 * kLbWarps, lb_smem, st>>>(L);
            CK(cudaGetLastError());
            st_local[0] += (int64_t)ni * nB;
        }
        CK(cudaMemsetAsync(B[AP_KCUR].p, 0, (size_t)ni * 4, st));
        row_kth_kernel<<<ni, 256, 0, st>>>(B[AP_LB].as<float>(), ldlb, (int32_t)nB, k1, B[AP_KTH].as<float>());
        CK(cudaGetLastError());
        for (int round = 0; round < 2; ++round) {
            const float *lo = round == 0 ? nullptr : B[AP_KTH].as<float>();
            const float *hi = round == 0 ? B[AP_KTH].as<float>() : B[AP_THR].as<float>();
            cand_rows_kernel<<<ni, 256, 0, st>>>(B[AP_LB].as<float>(), ldlb, (int32_t)nB, lo, hi, 0, B[AP_COUNTS].as<int32_t>(), nullptr,
                                                (int32_t)i0, nullptr, nullptr);
            CK(cudaGetLastError());
            scan_counts_kernel<<<1, 1024, 0, st>>>(B[AP_COUNTS].as<int32_t>(), ni * kCandWarps, B[AP_OFFS].as<int64_t>());
            CK(cudaGetLastError());
            int64_t ncand = 0;
            CK(cudaMemcpyAsync(&ncand, B[AP_OFFS].as<int64_t>() + (int64_t)ni * kCandWarps, 8, cudaMemcpyDeviceToHost, st));
            tm.mark(2);
            CK(cudaStreamSynchronize(st));                     // the one host round trip per round: the candidate count sizes the exact job
            if (ncand > 0) {
                if ((rc = B[AP_CI].ensure((size_t)ncand * 4)) || (rc = B[AP_CJ].ensure((size_t)ncand * 4)) || (rc = B[AP_CD].ensure((size_t)ncand * 8)) ||
                    (rc = B[AP_CST].ensure((size_t)ncand * 4)))
                    return rc;
                cand_rows_kernel<<<ni, 256, 0, st>>>(B[AP_LB].as<float>(), ldlb, (int32_t)nB, lo, hi, 1, nullptr, B[AP_OFFS].as<int64_t>(),
                                                    (int32_t)i0, B[AP_CI].as<int32_t>(), B[AP_CJ].as<int32_t>());
                CK(cudaGetLastError());
                if ((rc = ap_exact(E, st, dIdsA, dOffA, mlA, dIdsB, dOffB, mlB, B[AP_CI].as<int32_t>(), B[AP_CJ].as<int32_t>(), ncand,
                                   B[AP_CD].as<double>(), B[AP_CST].as<int32_t>())))
                    return rc;
            }
            topk_merge_kernel<<<ni, 256, (size_t)k * 12, st>>>(k, B[AP_OFFS].as<int64_t>(), B[AP_CJ].as<int32_t>(), B[AP_CD].as<double>(),
                                                             B[AP_TOPJ].as<int32_t>(), B[AP_TOPD].as<double>(), B[AP_KCUR].as<int32_t>(),
                                                             B[AP_THR].as<float>(), E->dmax);
            CK(cudaGetLastError());
            st_local[1 + round] += ncand;
            tm.mark(3);
        }
        CK(cudaMemcpyAsync(out_idx + i0 * k, B[AP_TOPJ].p, (size_t)ni * k * 4, back, st));
        CK(cudaMemcpyAsync(out_dist + i0 * k, B[AP_TOPD].p, (size_t)ni * k * 8, back, st));
        tm.mark(2);
    }
    CK(cudaStreamSynchronize(st));
    tm.collect(ms_local, 4);
    if (stats) for (int i = 0; i < 8; ++i) stats[i] = st_local[i];
    if (ms) for (int i = 0; i < 4; ++i) ms[i] = ms_local[i];
    return WMD_OK;
}

// mode argument of the pair entries: WMD_MODE_* in the low byte, WMD_IDS_ARE_ROWS as a flag
int parse_mode(wmd_engine *E, int32_t mode, int &base)
{
    base = mode & 0xff;
    if ((mode & ~(0xff | WMD_IDS_ARE_ROWS)) || (base != WMD_MODE_PYEMD && base != WMD_MODE_EXACT)) return fail(WMD_EINVAL, "unknown mode %d", mode);
    E->ids_are_rows = (mode & WMD_IDS_ARE_ROWS) != 0;
    return WMD_OK;
}

int ensure_pinned(void *&p, size_t &cap, size_t bytes)
{
    if (bytes <= cap) return WMD_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    const size_t want = bytes + bytes / 2 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return fail(WMD_ENOMEM, "cudaHostAlloc(%zu) failed", want); }
    cap = want;
    return WMD_OK;
}

size_t workspace_resident(const Workspace &W)
{
    const DevBuf *all[] = { &W.ids1, &W.ids2, &W.off1, &W.off2, &W.rows1, &W.cnt1, &W.ip1, &W.rows2, &W.cnt2, &W.ip2, &W.u12, &W.meta, &W.pqn,
                            &W.extra, &W.maxc, &W.tiles, &W.out, &W.status, &W.scratch, &W.plan, &W.wt1, &W.wt2, &W.lb, &W.l1, &W.l2, &W.am1, &W.am2,
                            &W.counters, &W.biglist };
    size_t t = 0;
    for (const DevBuf *b : all) t += b->cap;
    for (const DevBuf &b : W.wscratch) t += b.cap;
    t += W.sorted.cap;
    return t;
}

}  // namespace

extern "C" {

const char *wmd_last_error(void) { return g_err.c_str(); }
const char *wmd_version(void) { return "wmd_b200 0.1 (sm_100a)"; }

int wmd_create(const float *table_host, int64_t V, int32_t d, int64_t row_stride, int32_t normalize,
               int32_t device, wmd_handle *out)
{
    if (!out) return fail(WMD_EINVAL, "out is null");
    *out = nullptr;
    if (!table_host || V <= 0 || d <= 0 || row_stride < d) return fail(WMD_EINVAL, "bad table arguments");
    if (V > 0x7fffffff) return fail(WMD_EINVAL, "V too large");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(WMD_ENODEV, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(WMD_ENODEV, "device %d out of range (have %d)", device, ndev);
    wmd_engine *E = new wmd_engine();
    E->device = device;
    int rc = WMD_OK;
    auto bail = [&](int code) { wmd_destroy(E); return code; };
    if ((rc = set_device(E))) return bail(rc);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(WMD_ECUDA, "cudaGetDeviceProperties failed"));
    E->sm_count = prop.multiProcessorCount;
    E->smem_optin = prop.sharedMemPerBlockOptin;
    E->V = V; E->d = d; E->ld = (d + 3) & ~3;
    E->plan.nops = 0;
    bool ok = true;
    build_plan_rec(0, d, E->plan, ok);
    if (!ok) return bail(fail(WMD_EINVAL, "embedding width %d too large", d));
    {
        int max_iters = 10;
        if (const char *v = getenv("WMD_COST_CHUNK_ITERS")) max_iters = std::max(1, std::min(16, atoi(v)));
        if ((rc = build_cost_chunks(E, max_iters))) return bail(rc);
        if ((rc = setup_fast_path(E))) return bail(rc);
        if (const char *v = getenv("WMD_SERIAL")) E->slot_mask = atoi(v) ? 0 : 1;
        if (const char *v = getenv("WMD_SOLVE_BLOCKS")) E->solve_blocks_per_sm = std::max(1, atoi(v));
        if (const char *v = getenv("WMD_HOST_CHUNK")) {
            int a = 0, b = 0;
            if (sscanf(v, "%d,%d", &a, &b) == 2 && a >= 1024 && b >= a) { E->host_chunk_first = a; E->host_chunk_max = std::min(b, 1 << 20); }
        }
        if (const char *v = getenv("WMD_AP_R1MULT")) E->ap_r1_mult = std::max(1, atoi(v));
        if (const char *v = getenv("WMD_FUSED_MINB")) E->fused_minb = std::max(8, std::min(10, atoi(v)));
    }
    if (cudaMalloc(&E->table, (size_t)V * E->ld * 4) != cudaSuccess) return bail(fail(WMD_ENOMEM, "cudaMalloc table failed"));
    if (cudaMemset(E->table, 0, (size_t)V * E->ld * 4) != cudaSuccess) return bail(fail(WMD_ECUDA, "memset failed"));
    if (cudaMemcpy2D(E->table, (size_t)E->ld * 4, table_host, (size_t)row_stride * 4, (size_t)d * 4, (size_t)V, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(WMD_ECUDA, "table upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    for (int i = 0; i < 2; ++i) {
        if (cudaStreamCreateWithFlags(&E->streams[i], cudaStreamNonBlocking) != cudaSuccess) return bail(fail(WMD_ECUDA, "stream create failed"));
        if (cudaEventCreateWithFlags(&E->ev_join[i], cudaEventDisableTiming) != cudaSuccess) return bail(fail(WMD_ECUDA, "event create failed"));
        if (cudaEventCreateWithFlags(&E->ev_slot[i], cudaEventDisableTiming) != cudaSuccess) return bail(fail(WMD_ECUDA, "event create failed"));
    }
    if (cudaEventCreateWithFlags(&E->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(fail(WMD_ECUDA, "event create failed"));
    if (cudaStreamCreateWithFlags(&E->ap_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(WMD_ECUDA, "stream create failed"));
    if (cudaMalloc(&E->stats, 6 * sizeof(unsigned long long)) != cudaSuccess) return bail(fail(WMD_ENOMEM, "cudaMalloc failed"));
    cudaMemset(E->stats, 0, 6 * sizeof(unsigned long long));
    if (normalize) {
        normalize_rows_kernel<<<(unsigned)((V + 127) / 128), 128>>>(E->table, V, d, E->ld, E->plan);
        if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(WMD_ECUDA, "normalize failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    // Default policy of the pair entries: keep the V x V word-distance table when it fits the budget (built by the
    // first scoring call, so a handle that never scores never pays for it).  WMD_DTAB=0 / 1 forces the direct path /
    // the table; WMD_DTAB_BUDGET_MB moves the budget (default 4 GiB and at most a quarter of the free device memory).
    {
        size_t freeb = 0, totalb = 0;
        cudaMemGetInfo(&freeb, &totalb);
        E->dtab_budget = std::min<size_t>((size_t)4 << 30, freeb / 4);
        if (const char *v = getenv("WMD_DTAB_BUDGET_MB")) E->dtab_budget = (size_t)std::max(0, atoi(v)) << 20;
        E->use_dtab = (size_t)V * (size_t)V * 4 <= E->dtab_budget;
        if (const char *v = getenv("WMD_DTAB")) { E->use_dtab = atoi(v) != 0; E->dtab_forced = E->use_dtab; }
    }
    *out = E;
    return WMD_OK;
}

int wmd_destroy(wmd_handle E)
{
    if (!E) return WMD_OK;
    cudaSetDevice(E->device);
    cudaDeviceSynchronize();
    for (auto &r : E->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (int i = 0; i < 2; ++i) {
        E->ws[i].release();
        if (E->streams[i]) cudaStreamDestroy(E->streams[i]);
        if (E->ev_join[i]) cudaEventDestroy(E->ev_join[i]);
        if (E->ev_slot[i]) cudaEventDestroy(E->ev_slot[i]);
    }
    if (E->ev_fork) cudaEventDestroy(E->ev_fork);
    if (E->dtab) cudaFree(E->dtab);
    if (E->dtab16) cudaFree(E->dtab16);
    for (auto &b : E->ap) b.release();
    if (E->ap_stream) cudaStreamDestroy(E->ap_stream);
    if (E->table) cudaFree(E->table);
    if (E->map) cudaFree(E->map);
    if (E->rank) cudaFree(E->rank);
    if (E->stats) cudaFree(E->stats);
    if (E->pin_in) cudaFreeHost(E->pin_in);
    if (E->pin_out) cudaFreeHost(E->pin_out);
    delete E;
    return WMD_OK;
}

int wmd_set_token_map(wmd_handle E, const int32_t *id_to_row_host, int64_t n)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int rc;
    if ((rc = set_device(E))) return rc;
    CK(cudaDeviceSynchronize());
    if (E->map) { CK(cudaFree(E->map)); E->map = nullptr; E->nmap = 0; }
    if (!id_to_row_host || n <= 0) return WMD_OK;
    CK(cudaMalloc(&E->map, (size_t)n * 4));
    CK(cudaMemcpy(E->map, id_to_row_host, (size_t)n * 4, cudaMemcpyHostToDevice));
    E->nmap = n;
    return WMD_OK;
}

int wmd_set_rank(wmd_handle E, const int32_t *rank_host, int64_t V)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int rc;
    if ((rc = set_device(E))) return rc;
    CK(cudaDeviceSynchronize());
    if (E->rank) { CK(cudaFree(E->rank)); E->rank = nullptr; }
    if (!rank_host) return WMD_OK;
    if (V != E->V) return fail(WMD_EINVAL, "rank table has %lld entries, the embedding table %lld rows", (long long)V, (long long)E->V);
    std::vector<char> seen((size_t)V, 0);
    for (int64_t i = 0; i < V; ++i) {
        const int32_t r = rank_host[i];
        if (r < 0 || r >= V || seen[(size_t)r]) return fail(WMD_EINVAL, "rank table is not a permutation of 0..V-1 (entry %lld)", (long long)i);
        seen[(size_t)r] = 1;
    }
    CK(cudaMalloc(&E->rank, (size_t)V * 4));
    CK(cudaMemcpy(E->rank, rank_host, (size_t)V * 4, cudaMemcpyHostToDevice));
    return WMD_OK;
}

int wmd_get_table(wmd_handle E, float *out_host)
{
    if (!E || !out_host) return fail(WMD_EINVAL, "null argument");
    int rc;
    if ((rc = set_device(E))) return rc;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy2D(out_host, (size_t)E->d * 4, E->table, (size_t)E->ld * 4, (size_t)E->d * 4, (size_t)E->V, cudaMemcpyDeviceToHost));
    return WMD_OK;
}

int wmd_pairs_host(wmd_handle E, const int32_t *ids1, const int64_t *off1, const int32_t *ids2, const int64_t *off2,
                   int64_t npairs, int32_t mode, double *out, int32_t *status)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int base, rc;
    if ((rc = parse_mode(E, mode, base))) return rc;
    if (!out && npairs > 0) return fail(WMD_EINVAL, "out is null");
    HostJob J{ ids1, off1, ids2, off2, npairs, out, status };
    J.mode = base;
    return run_host_job(E, J);
}

int wmd_pairs_host_in_dev_out(wmd_handle E, const int32_t *ids1, const int64_t *off1, const int32_t *ids2, const int64_t *off2,
                              int64_t npairs, int32_t mode, double *out_dev, int32_t *status_dev)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int base, rc;
    if ((rc = parse_mode(E, mode, base))) return rc;
    if (!out_dev && npairs > 0) return fail(WMD_EINVAL, "out is null");
    HostJob J{ ids1, off1, ids2, off2, npairs, out_dev, status_dev };
    J.mode = base; J.out_on_device = true;
    return run_host_job(E, J);
}

/* Asynchronous form of wmd_pairs_host for batches of the in-loop caller's size (src/loader.py:60): the documents are
 * staged into pinned memory the handle owns, copies and kernels are queued on the handle's streams and the call
 * returns; wmd_pairs_wait blocks until the scores are back and hands them over. */
int wmd_pairs_submit(wmd_handle E, const int32_t *ids1, const int64_t *off1, const int32_t *ids2, const int64_t *off2,
                     int64_t npairs, int32_t mode)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int base, rc;
    if ((rc = parse_mode(E, mode, base))) return rc;
    HostJob J{ ids1, off1, ids2, off2, npairs, nullptr, nullptr };
    J.mode = base;
    if ((rc = check_host_job(E, J))) return rc;
    if (npairs == 0) { E->pending_pairs = 0; return WMD_OK; }
    const int64_t b1 = off1[0], b2 = off2[0], t1 = off1[npairs] - b1, t2 = off2[npairs] - b2;
    if (b1 < 0 || b2 < 0 || t1 < 0 || t2 < 0) return fail(WMD_EINVAL, "bad offsets");
    const size_t n_off = (size_t)(npairs + 1) * 8, n1 = (size_t)t1 * 4, n2 = (size_t)t2 * 4;
    const size_t in_bytes = 2 * n_off + ((n1 + 7) & ~(size_t)7) + n2 + 16;
    if ((rc = ensure_pinned(E->pin_in, E->pin_in_cap, in_bytes)) || (rc = ensure_pinned(E->pin_out, E->pin_out_cap, (size_t)npairs * 12 + 16))) return rc;
    // pinned layout: off1 | off2 | ids1 | ids2  (offsets rebased to zero)
    int64_t *po1 = static_cast<int64_t *>(E->pin_in), *po2 = po1 + npairs + 1;
    int32_t *pi1 = reinterpret_cast<int32_t *>(po2 + npairs + 1);
    int32_t *pi2 = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(pi1) + ((n1 + 7) & ~(size_t)7));
    for (int64_t p = 0; p <= npairs; ++p) { po1[p] = off1[p] - b1; po2[p] = off2[p] - b2; }
    if (t1) memcpy(pi1, ids1 + b1, n1);
    if (t2) memcpy(pi2, ids2 + b2, n2);
    J.ids1 = pi1; J.off1 = po1; J.ids2 = pi2; J.off2 = po2;
    J.out = static_cast<double *>(E->pin_out);
    J.status = reinterpret_cast<int32_t *>(J.out + npairs);
    if ((rc = enqueue_host_job(E, J))) {
        const std::string msg = g_err;
        drain_streams(E);
        g_err = msg;
        return rc;
    }
    E->pending_pairs = npairs;
    return WMD_OK;
}

int wmd_pairs_wait(wmd_handle E, double *out, int32_t *status)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (E->pending_pairs < 0) return fail(WMD_EINVAL, "no submitted job to wait for");
    const int64_t n = E->pending_pairs;
    E->pending_pairs = -1;
    int rc;
    if ((rc = set_device(E))) return rc;
    if (n == 0) return WMD_OK;
    if (!out) { drain_streams(E); return fail(WMD_EINVAL, "out is null"); }
    if ((rc = drain_streams(E))) return rc;
    const double *po = static_cast<const double *>(E->pin_out);
    memcpy(out, po, (size_t)n * 8);
    if (status) memcpy(status, po + n, (size_t)n * 4);
    return WMD_OK;
}

int wmd_set_fanout(wmd_handle E, int32_t n, double *const *out_ptrs, int32_t *const *status_ptrs)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (n < 0 || n > kMaxFan) return fail(WMD_EINVAL, "fan-out of %d arrays; the limit is %d", (int)n, kMaxFan);
    if (n > 0 && (!out_ptrs || !status_ptrs)) return fail(WMD_EINVAL, "null fan-out pointers");
    OutFan F;
    F.n = n;
    for (int k = 0; k < n; ++k) {
        if (!out_ptrs[k] || !status_ptrs[k]) return fail(WMD_EINVAL, "fan-out array %d is null", k);
        F.out[k] = out_ptrs[k]; F.status[k] = status_ptrs[k];
    }
    E->fan = F;
    return WMD_OK;
}

int wmd_peer_alloc(wmd_handle E, int64_t bytes, void **dev_ptr, unsigned char *ipc_handle)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (bytes <= 0 || !dev_ptr || !ipc_handle) return fail(WMD_EINVAL, "bad arguments");
    int rc;
    if ((rc = set_device(E))) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == WMD_IPC_HANDLE_BYTES, "WMD_IPC_HANDLE_BYTES");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return fail(WMD_ENOMEM, "cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    if ((e = cudaMemset(p, 0, (size_t)bytes)) != cudaSuccess || (e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
        cudaFree(p);
        return fail(WMD_ECUDA, "peer buffer: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, sizeof h);
    *dev_ptr = p;
    return WMD_OK;
}

int wmd_peer_open(wmd_handle E, const unsigned char *ipc_handle, void **dev_ptr)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (!ipc_handle || !dev_ptr) return fail(WMD_EINVAL, "bad arguments");
    int rc;
    if ((rc = set_device(E))) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(WMD_ECUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *dev_ptr = p;
    return WMD_OK;
}

int wmd_peer_close(wmd_handle E, void *dev_ptr, int32_t opened)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (!dev_ptr) return WMD_OK;
    int rc;
    if ((rc = set_device(E))) return rc;
    const cudaError_t e = opened ? cudaIpcCloseMemHandle(dev_ptr) : cudaFree(dev_ptr);
    if (e != cudaSuccess) return fail(WMD_ECUDA, "%s failed: %s", opened ? "cudaIpcCloseMemHandle" : "cudaFree", cudaGetErrorString(e));
    return WMD_OK;
}

int wmd_workspace_bytes(wmd_handle E, int64_t npairs, int32_t max_len1, int32_t max_len2, int64_t *estimate, int64_t *resident)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (npairs < 0 || max_len1 < 0 || max_len2 < 0 || max_len1 > WMD_MAX_DOC_LEN || max_len2 > WMD_MAX_DOC_LEN) return fail(WMD_EINVAL, "bad sizes");
    const int ml1 = std::max(max_len1, 1), ml2 = std::max(max_len2, 1), ML = std::max(ml1, ml2);
    const bool table = E->use_dtab;
    const size_t fixed = (size_t)E->V * E->ld * 4 + (size_t)E->nmap * 4 + (E->rank ? (size_t)E->V * 4 : 0) +
                         (table ? (size_t)E->V * E->V * 4 : 0);
    if (estimate) {
        const int64_t Bc = std::min<int64_t>(npairs, table ? std::max<int64_t>(E->host_chunk_max, 65536) : chunk_pairs(ml1, ml2, false));
        const int slots = npairs > Bc ? 2 : 1;
        size_t per = (size_t)Bc * (2 * 8 + 8 + 4 + 4);                                   // offsets, out, status, biglist
        per += (size_t)Bc * (size_t)(ml1 + ml2) * 4;                                     // staged ids (upper bound)
        const bool general = !table || ML >= 32;
        if (general) {
            per += (size_t)Bc * (size_t)(ml1 + ml2) * 12 + (size_t)Bc * 28;              // rows / counts / masses, per-pair records
            if (!table) {
                per += (size_t)Bc * ml1 * ml2 * 4;                                       // cost tiles
                per += (size_t)plan_stage_bound(Bc, ml1, ml2, Bc * (int64_t)ml1, Bc * (int64_t)ml2, std::max(E->fast_R, 8), kStageTilesMax) * sizeof(StageRec);
            }
            for (int kc = 1; kc <= 8; ++kc) {                                            // wide solver scratch: costs + flow per resident warp, per class
                if (ML < wide_min_ml(kc)) continue;
                static const int blocks[8] = { 8, 8, 6, 6, 6, 6, 5, 5 };                 // resident blocks per SM (__launch_bounds__ of the instances)
                const size_t warps = std::min<size_t>((size_t)E->sm_count * blocks[kc - 1] * 4, (size_t)((Bc + 3) / 4) * 4);
                const size_t mr = (size_t)std::min(kc == 8 ? kMaxDocLen + 1 : kMaxDocLen, ML + 1);
                per += warps * solve_wide_scratch_ints_per_warp((int)mr, 32 * kc) * 4;
            }
        }
        *estimate = (int64_t)(fixed + (size_t)slots * per);
    }
    if (resident) {
        size_t r = (size_t)E->V * E->ld * 4 + (size_t)E->nmap * 4 + (E->rank ? (size_t)E->V * 4 : 0) + (E->dtab ? (size_t)E->V * E->V * 4 : 0);
        r += workspace_resident(E->ws[0]) + workspace_resident(E->ws[1]);
        for (const DevBuf &b : E->ap) r += b.cap;
        *resident = (int64_t)r;
    }
    return WMD_OK;
}

int wmd_rwmd_pairs_host(wmd_handle E, const int32_t *ids1, const int64_t *off1, const int32_t *ids2, const int64_t *off2,
                        int64_t npairs, double *lb, double *l1, double *l2, int32_t *argmin_rows, int32_t *argmin_cols,
                        int32_t *status)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (!lb && npairs > 0) return fail(WMD_EINVAL, "lb is null");
    E->ids_are_rows = false;
    HostJob J{ ids1, off1, ids2, off2, npairs, nullptr, status };
    J.rwmd = true; J.solve = false; J.lb = lb; J.l1 = l1; J.l2 = l2; J.am1 = argmin_rows; J.am2 = argmin_cols;
    return run_host_job(E, J);
}

int wmd_pairs_dev(wmd_handle E, const int32_t *ids1, const int64_t *off1, int64_t total1, int32_t max_len1,
                  const int32_t *ids2, const int64_t *off2, int64_t total2, int32_t max_len2,
                  int64_t npairs, int32_t mode, double *out, int32_t *status, void *stream)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int base, rc;
    if ((rc = parse_mode(E, mode, base))) return rc;
    if (npairs > 0 && (!off1 || !off2)) return fail(WMD_EINVAL, "null offsets");
    DocSide s1{}, s2{};
    s1.ids = ids1; s1.off = off1; s2.ids = ids2; s2.off = off2;
    return run_dev_pairs(E, s1, s2, total1, total2, max_len1, max_len2, npairs, out, status, (cudaStream_t)stream, base);
}

int wmd_rwmd_pairs_dev(wmd_handle E, const int32_t *ids1, const int64_t *off1, int64_t total1, int32_t max_len1,
                       const int32_t *ids2, const int64_t *off2, int64_t total2, int32_t max_len2, int64_t npairs,
                       double *lb, double *l1, double *l2, int32_t *argmin_rows, int32_t *argmin_cols, int32_t *status, void *stream)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (npairs > 0 && (!off1 || !off2 || !lb)) return fail(WMD_EINVAL, "null argument");
    E->ids_are_rows = false;
    DocSide s1{}, s2{};
    s1.ids = ids1; s1.off = off1; s2.ids = ids2; s2.off = off2;
    ChunkOut O;
    O.out = nullptr; O.status = status; O.solve = false; O.rwmd = true; O.am_abs = true;
    O.lb = lb; O.l1 = l1; O.l2 = l2; O.am1 = argmin_rows; O.am2 = argmin_cols;
    return run_dev_job(E, s1, s2, total1, total2, max_len1, max_len2, npairs, O, (cudaStream_t)stream);
}

int wmd_nbow_dev(wmd_handle E, const int32_t *ids, const int64_t *off, int64_t ndocs, int32_t max_len,
                 int32_t *rows, int32_t *counts, double *weights, int32_t *uniq, void *stream)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (ndocs < 0 || (ndocs > 0 && (!off || !rows || !counts || !uniq))) return fail(WMD_EINVAL, "null argument");
    if (max_len < 0 || max_len > WMD_MAX_DOC_LEN) return fail(WMD_EINVAL, "max_len must be within [0, %d]", WMD_MAX_DOC_LEN);
    if (ndocs == 0) return WMD_OK;
    if (ndocs > 0x7fffffff) return fail(WMD_EINVAL, "too many documents");
    int rc;
    if ((rc = set_device(E))) return rc;
    E->ids_are_rows = false;
    cudaStream_t st = (cudaStream_t)stream;
    DocSide s{};
    s.ids = ids; s.off = off;
    const int Lp = std::max(max_len, 1);
    const int wpb = Lp <= 64 ? 8 : 4;
    const size_t smem = nbow_smem_per_warp(Lp) * wpb;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(nbow_docs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((ndocs + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
    Prof pr(E, WMD_K_NBOW, st);
    nbow_docs_kernel<<<grid, wpb * 32, smem, st>>>(s, make_vocab(E), (int32_t)ndocs, Lp, rows, counts, weights, uniq, nullptr);
    CK(cudaGetLastError());
    return WMD_OK;
}

int wmd_pairs_padded_dev(wmd_handle E, const int32_t *a, int32_t L1, const int32_t *b, int32_t L2, int64_t npairs,
                         int32_t pad_id, int32_t mode, double *out, int32_t *status, void *stream)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int base, rc;
    if ((rc = parse_mode(E, mode, base))) return rc;
    if (npairs > 0 && (!a || !b)) return fail(WMD_EINVAL, "null ids");
    if (L1 <= 0 || L2 <= 0) return fail(WMD_EINVAL, "padded lengths must be positive");
    DocSide s1{}, s2{};
    s1.ids = a; s1.off = nullptr; s1.L = L1; s1.pad_id = pad_id; s1.has_pad = 1;
    s2.ids = b; s2.off = nullptr; s2.L = L2; s2.pad_id = pad_id; s2.has_pad = 1;
    return run_dev_pairs(E, s1, s2, npairs * L1, npairs * L2, L1, L2, npairs, out, status, (cudaStream_t)stream, base);
}

int wmd_nbow_host(wmd_handle E, const int32_t *ids, const int64_t *off, int64_t ndocs,
                  int32_t *rows, int32_t *counts, double *weights, int32_t *uniq)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (ndocs < 0 || (ndocs > 0 && (!off || !rows || !counts || !weights || !uniq))) return fail(WMD_EINVAL, "null argument");
    if (ndocs == 0) return WMD_OK;
    if (ndocs > 0x7fffffff) return fail(WMD_EINVAL, "too many documents");
    int rc;
    if ((rc = set_device(E))) return rc;
    if (E->pending_pairs >= 0) return fail(WMD_EINVAL, "a submitted job is still in flight: call wmd_pairs_wait first");   // shares the slot-0 workspace
    E->ids_are_rows = false;
    int32_t ml;
    if ((rc = scan_offsets(off, ndocs, ml, "documents"))) return rc;
    const int64_t base = off[0], total = off[ndocs] - base;
    if (total > 0 && !ids) return fail(WMD_EINVAL, "null ids");
    Workspace &W = E->ws[0];
    cudaStream_t st = E->streams[0];
    const size_t tb = (size_t)std::max<int64_t>(total, 1);
    if ((rc = W.ids1.ensure(tb * 4)) || (rc = W.off1.ensure((size_t)(ndocs + 1) * 8)) || (rc = W.rows1.ensure(tb * 4)) ||
        (rc = W.cnt1.ensure(tb * 4)) || (rc = W.pqn.ensure(tb * 8)) || (rc = W.u12.ensure((size_t)ndocs * 4)))
        return rc;
    if (total) CK(cudaMemcpyAsync(W.ids1.p, ids + base, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(W.off1.p, off, (size_t)(ndocs + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(W.rows1.p, 0xff, tb * 4, st));
    CK(cudaMemsetAsync(W.cnt1.p, 0, tb * 4, st));
    CK(cudaMemsetAsync(W.pqn.p, 0, tb * 8, st));
    DocSide s{};
    s.ids = W.ids1.as<int32_t>() - base; s.off = W.off1.as<int64_t>();
    const int Lp = std::max(ml, 1);
    const int wpb = Lp <= 64 ? 8 : 4;
    const size_t smem = nbow_smem_per_warp(Lp) * wpb;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(nbow_docs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((ndocs + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
    {
        Prof pr(E, WMD_K_NBOW, st);
        nbow_docs_kernel<<<grid, wpb * 32, smem, st>>>(s, make_vocab(E), (int32_t)ndocs, Lp,
                                                      W.rows1.as<int32_t>() - base, W.cnt1.as<int32_t>() - base,
                                                      W.pqn.as<double>() - base, W.u12.as<int32_t>(), nullptr);
        CK(cudaGetLastError());
    }
    if (total) {
        CK(cudaMemcpyAsync(rows + base, W.rows1.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(counts + base, W.cnt1.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(weights + base, W.pqn.p, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(uniq, W.u12.p, (size_t)ndocs * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return WMD_OK;
}

int wmd_allpairs_topk_host(wmd_handle E, const int32_t *idsA, const int64_t *offA, int64_t nA,
                           const int32_t *idsB, const int64_t *offB, int64_t nB, int32_t k, int64_t row_begin, int64_t row_end,
                           int32_t *out_idx, double *out_dist, int64_t *stats, double *ms)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    return run_allpairs(E, idsA, offA, nA, idsB, offB, nB, k, row_begin, row_end, out_idx, out_dist, false, stats, ms);
}

int wmd_allpairs_topk_dev(wmd_handle E, const int32_t *idsA, const int64_t *offA, int64_t nA,
                          const int32_t *idsB, const int64_t *offB, int64_t nB, int32_t k, int64_t row_begin, int64_t row_end,
                          int32_t *out_idx_dev, double *out_dist_dev, int64_t *stats, double *ms)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    return run_allpairs(E, idsA, offA, nA, idsB, offB, nB, k, row_begin, row_end, out_idx_dev, out_dist_dev, true, stats, ms);
}

int wmd_emd_batch_host(wmd_handle E, const double *P, const double *Q, const double *D, int64_t nprob, int32_t n,
                       int32_t shared_d, double extra_mass_penalty, double *out)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (nprob < 0 || n <= 0) return fail(WMD_EINVAL, "bad sizes");
    if (n > kEmdMaxBins) return fail(WMD_EINVAL, "histograms of %d bins: the limit is %d", n, kEmdMaxBins);
    if (nprob == 0) return WMD_OK;
    if (!P || !Q || !D || !out) return fail(WMD_EINVAL, "null argument");
    int rc;
    if ((rc = set_device(E))) return rc;
    cudaStream_t st = E->ap_stream;
    DevBuf *B = E->ap;
    const int64_t CH = 1 << 20;                                        // problems per launch
    const size_t dsz = (size_t)n * n * 8;
    for (int64_t p0 = 0; p0 < nprob; p0 += CH) {
        const int64_t nb = std::min<int64_t>(CH, nprob - p0);
        if ((rc = B[AP_IDSA].ensure((size_t)nb * n * 8)) || (rc = B[AP_IDSB].ensure((size_t)nb * n * 8)) ||
            (rc = B[AP_ZA].ensure(shared_d ? dsz : (size_t)nb * dsz)) || (rc = B[AP_CD].ensure((size_t)nb * 8)))
            return rc;
        CK(cudaMemcpyAsync(B[AP_IDSA].p, P + p0 * n, (size_t)nb * n * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B[AP_IDSB].p, Q + p0 * n, (size_t)nb * n * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B[AP_ZA].p, shared_d ? D : D + p0 * (int64_t)n * n, shared_d ? dsz : (size_t)nb * dsz, cudaMemcpyHostToDevice, st));
        EmdArgs A;
        A.P = B[AP_IDSA].as<double>(); A.Q = B[AP_IDSB].as<double>(); A.D = B[AP_ZA].as<double>();
        A.nprob = nb; A.n = n; A.shared_d = shared_d; A.extra_mass_penalty = extra_mass_penalty; A.out = B[AP_CD].as<double>();
        const int wpb = 4;
        const size_t smem = emd_smem_per_warp(n) * wpb;
        const int grid = (int)std::min<int64_t>((nb + wpb - 1) / wpb, (int64_t)E->sm_count * 8);
        {
            Prof pr(E, WMD_K_SOLVE, st);
            emd_hat_batch_kernel<<<grid, wpb * 32, smem, st>>>(A);
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(out + p0, B[AP_CD].p, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return WMD_OK;
}

int wmd_set_profiling(wmd_handle E, int32_t enabled)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    E->profiling = enabled != 0;
    return WMD_OK;
}

int wmd_set_serial(wmd_handle E, int32_t enabled)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    E->slot_mask = enabled ? 0 : 1;
    return WMD_OK;
}

int wmd_set_distance_table(wmd_handle E, int32_t enabled)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (!enabled) { E->use_dtab = false; E->dtab_forced = false; return WMD_OK; }
    int rc;
    if ((rc = set_device(E))) return rc;
    if (E->pending_pairs >= 0) return fail(WMD_EINVAL, "a submitted job is still in flight: call wmd_pairs_wait first");
    if ((rc = ensure_dtab(E, E->streams[0]))) return rc;
    E->use_dtab = true; E->dtab_forced = true;
    return WMD_OK;
}

int wmd_distance_table_info(wmd_handle E, int64_t *bytes, double *build_ms, int32_t *enabled, int32_t *resident)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    if (bytes) *bytes = (int64_t)E->V * E->V * 4;
    if (build_ms) *build_ms = E->dtab_build_ms;
    if (enabled) *enabled = E->use_dtab ? 1 : 0;
    if (resident) *resident = E->dtab ? 1 : 0;
    return WMD_OK;
}

int wmd_get_profile(wmd_handle E, double *ms, int64_t *launches, int32_t reset)
{
    if (!E) return fail(WMD_EINVAL, "null handle");
    int rc;
    if ((rc = set_device(E))) return rc;
    CK(cudaDeviceSynchronize());
    for (auto &r : E->prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { E->prof_ms[r.kind] += t; E->prof_n[r.kind] += 1; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    E->prof.clear();
    for (int k = 0; k < WMD_K_COUNT; ++k) {
        if (ms) ms[k] = E->prof_ms[k];
        if (launches) launches[k] = E->prof_n[k];
        if (reset) { E->prof_ms[k] = 0; E->prof_n[k] = 0; }
    }
    return WMD_OK;
}

int wmd_get_last_stats(wmd_handle E, int64_t *values)
{
    if (!E || !values) return fail(WMD_EINVAL, "null argument");
    int rc;
    if ((rc = set_device(E))) return rc;
    CK(cudaDeviceSynchronize());
    unsigned long long h[6];
    CK(cudaMemcpy(h, E->stats, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) values[i] = (int64_t)h[i];
    return WMD_OK;
}

}  // extern "C"
