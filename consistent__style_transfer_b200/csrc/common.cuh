// common.cuh -- shared definitions of the sm_100a WMD kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

namespace wmd {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxDocLen = 256;          // WMD_MAX_DOC_LEN
constexpr int kIntInf = 0x3fffffff;      // "infinite" tentative distance (sums of two stay < 2^31)

// per-pair solver classes written by the nBOW kernel (meta & kMetaCls)
constexpr int kClsNone = 0;              // nothing to do (early-out already written)
constexpr int kClsA = 1;                 // residual rows <= 32 and columns (incl. dummy) <= 32: the fused / class-A kernels
// Larger problems (solve_wide.cuh): rows = the side with more nodes (<= 257 incl. a dummy row), class = 1 + the number
// KC of 32-column words of the SHORTER side (emd_solve_wide_kernel<KC>, KC = 1 .. 8).
constexpr int kClsW1 = 2;
constexpr int kClsW8 = kClsW1 + 7;
constexpr int kClsLast = kClsW8;
constexpr int kMetaCls = 15;
// class of a residual problem of m supplying rows and nc columns (incl. the dummy column)
__host__ __device__ inline int solver_class(int m, int nc)
{
    const int hi = m > nc ? m : nc, lo = m > nc ? nc : m;
    if (hi <= 32) return kClsA;
    if (hi > kMaxDocLen) return kClsW8;                          // 256 nodes a side plus the dummy: the one launch with 257 rows
    return kClsW1 + ((lo + 31) >> 5) - 1;
}
constexpr int kMetaWorkShift = 8;        // bits 8 .. 15: work estimate of the pair, min(255, rows x columns / 256) (longest-first order)
constexpr int kMetaSwap = 16;            // doc2 is the heavier (supplying) side

// One side of a batch of documents. CSR (off != nullptr) or padded [npairs, L] (off == nullptr).
// With sel != nullptr pair p uses document sel[p] (all-pairs candidate lists: many pairs share a
// document), and the per-pair work slots are laid out at q * slot instead of at the CSR offsets.
struct DocSide {
    const int32_t *ids;
    const int64_t *off;
    int32_t L;            // padded row length when off == nullptr
    int32_t pad_id;
    int32_t has_pad;
    int32_t slot;         // work-slot pitch when sel != nullptr (>= the longest document)
    const int32_t *sel;   // optional document index per pair
};

__device__ __forceinline__ void doc_span(const DocSide &s, int64_t p, int64_t &start, int &len)
{
    if (s.sel) p = s.sel[p];
    if (s.off) { start = s.off[p]; len = (int)(s.off[p + 1] - start); }
    else       { start = p * (int64_t)s.L; len = s.L; }
}
// offset of pair q's work slots (rows / counts / masses) inside the launch's per-token arrays
__device__ __forceinline__ int64_t slot_off(const DocSide &s, int64_t tokbase, int q, int64_t start)
{
    return s.sel ? (int64_t)q * s.slot : start - tokbase;
}

// Multi-GPU gather fused into the kernels (wmd_set_fanout): every score / status a pair entry stores ALSO goes to the same
// pair index of up to kMaxFan further arrays -- the peers' copies of the global result, mapped over NVLink -- so that the
// ranks of a sharded job need no all-gather of the scores, only a barrier.  Stores are fire-and-forget: the transfer
// overlaps the solves pair by pair.
constexpr int kMaxFan = 7;
struct OutFan {
    int32_t n = 0;                    // default: off (argument structs are filled field by field)
    int32_t _pad = 0;
    double *out[kMaxFan] = {};
    int32_t *status[kMaxFan] = {};
};
// (statically indexed: a run-time index into kernel parameters would copy the whole struct to local memory)
__device__ __forceinline__ void fan_score(const OutFan &F, int64_t p, double v)
{
#pragma unroll
    for (int k = 0; k < kMaxFan; ++k) if (k < F.n) F.out[k][p] = v;
}
__device__ __forceinline__ void fan_status(const OutFan &F, int64_t p, int s)
{
#pragma unroll
    for (int k = 0; k < kMaxFan; ++k) if (k < F.n) F.status[k][p] = s;
}

struct Vocab {
    const float *table;       // [V, ld] float32, device
    int64_t V;
    int32_t d;
    int32_t ld;               // floats between rows (multiple of 4)
    const int32_t *map;       // tokenizer id -> row, or nullptr
    int64_t nmap;
    const int32_t *rank;      // row -> canonical order key, or nullptr
};

// numpy FLOAT_pairwise_sum recursion flattened to a postfix program: each op sums one leaf
// block [start, start+len) with eight strided accumulators and pushes it; `adds` pops follow.
constexpr int kMaxPlanOps = 64;
struct SumPlan {
    int32_t nops;
    int32_t start[kMaxPlanOps];
    int32_t len[kMaxPlanOps];
    int32_t adds[kMaxPlanOps];
};

__device__ __forceinline__ int warp_sum(int v)
{
    return __reduce_add_sync(kFull, v);
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

}  // namespace wmd
