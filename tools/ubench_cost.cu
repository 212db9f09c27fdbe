// Micro-benchmark of the K2 consumer loop (run on the B200 via gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -o tools/ubench_cost tools/ubench_cost.cu
// A stage (two 9x9 / 10x9 pairs worth of table rows + tile descriptors) is built once in shared
// memory and the 12 consumer warps run the real run_desc_batch<4> over it again and again with no
// producer, barrier or atomics in the way.  Reports SM cycles per tile task: the ceiling the full
// kernel can approach.  Variants: row pitch (bank mapping) and number of consumer warps.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../consistent__style_transfer_b200/csrc/cost_fast.cuh"
using namespace wmd;

__global__ void __launch_bounds__(512, 1)
bench_kernel(FastArgs A0, const uint4 *descs, int ntiles, int nrows, const float *table, int reps, long long *cyc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FastArgs A = A0;
    A.tiles += (size_t)blockIdx.x * 1024; A.maxc += blockIdx.x * 2;       // no cross-SM contention on the outputs
    uint4 *d = reinterpret_cast<uint4 *>(smem_raw);
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) d[i] = descs[i];
    float *rows = reinterpret_cast<float *>(smem_raw + kStageDescBytes);
    for (int i = threadIdx.x; i < nrows * A.vc.ld; i += blockDim.x) {
        const int r = i / A.vc.ld, e = i - r * A.vc.ld;
        rows[r * A.ldr + e] = table[(size_t)((r * 37 + blockIdx.x) % 1000) * A.vc.ld + e];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
        for (int t = warp * 4; t < ntiles; t += nw * 4) run_desc_batch<4>(A, smem_raw, ntiles, t);
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static void make_descs(std::vector<uint4> &out, int u1, int u2, int rowbase, int q)
{
    const int p0 = ((u1 + 1) / 2) * ((u2 + 3) / 4), p1 = ((u2 + 1) / 2) * ((u1 + 3) / 4);
    const int tr = p1 < p0;
    const int na = tr ? u2 : u1, nb = tr ? u1 : u2;
    const int abase = rowbase + (tr ? u1 : 0), bbase = rowbase + (tr ? 0 : u1);
    const int TI = (na + 1) / 2, TJ = (nb + 3) / 4;
    for (int t = 0; t < TI * TJ; ++t) {
        const int ti = t / TJ, tj = t % TJ;
        const int va = ti + TI < na ? 2 : 1;
        int vb = 1;
        for (int c = 1; c < 4; ++c) vb += tj + c * TJ < nb;
        unsigned ra[2], rb[4];
        for (int r = 0; r < 2; ++r) ra[r] = abase + (r < va ? ti + r * TI : ti);
        for (int c = 0; c < 4; ++c) rb[c] = bbase + (c < vb ? tj + c * TJ : tj);
        const unsigned sr = tr ? TI : TI * u2, sc = tr ? TJ * u2 : TJ, off = tr ? tj * u2 + ti : ti * u2 + tj;
        uint4 w;
        w.x = ra[0] | (ra[1] << 8) | (rb[0] << 16) | (rb[1] << 24);
        w.y = rb[2] | (rb[3] << 8) | (va << 16) | (vb << 24);
        w.z = sr | (sc << 16);
        w.w = q | (off << 16);
        out.push_back(w);
    }
}

int main()
{
    const int d = 300, ld = 300, V = 1000;
    std::vector<float> tab((size_t)V * ld);
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = (float)((i * 2654435761u) % 1000) / 1000.f;
    float *dtab; cudaMalloc(&dtab, tab.size() * 4); cudaMemcpy(dtab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice);
    std::vector<uint4> descs;
    make_descs(descs, 9, 9, 0, 0);
    make_descs(descs, 10, 9, 18, 1);
    make_descs(descs, 8, 8, 37, 2);
    make_descs(descs, 9, 10, 53, 3);
    const int ntiles = (int)descs.size(), nrows = 72;
    uint4 *ddesc; cudaMalloc(&ddesc, descs.size() * 16); cudaMemcpy(ddesc, descs.data(), descs.size() * 16, cudaMemcpyHostToDevice);
    float *tiles; cudaMalloc(&tiles, 1 << 20);
    unsigned *maxc; cudaMalloc(&maxc, 4096); cudaMemset(maxc, 0, 4096);
    long long *cyc; cudaMalloc(&cyc, 148 * 8);
    FastArgs A{};
    A.vc.table = dtab; A.vc.V = V; A.vc.d = d; A.vc.ld = ld;
    A.plan.nops = 4;
    const int st[4] = { 0, 72, 144, 216 }, ln[4] = { 72, 72, 72, 84 }, ad[4] = { 0, 1, 0, 2 };
    for (int i = 0; i < 4; ++i) { A.plan.start[i] = st[i]; A.plan.len[i] = ln[i]; A.plan.adds[i] = ad[i]; }
    A.R = 40; A.S = 1; A.common_iters = 9; A.rowbytes = ld * 4; A.negzero2 = 0x8000000080000000ull;
    A.tiles = tiles; A.tile_stride = 512; A.maxc = maxc;
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    const int reps = 200;
    printf("stage: %d tiles, %d rows\n", ntiles, nrows);
    for (int ldr : { 300 })
        for (int warps : { 1, 2, 4, 8, 12, 16 }) {
            A.ldr = ldr;
            bench_kernel<<<148, warps * 32, 120 * 1024>>>(A, ddesc, ntiles, nrows, dtab, reps, cyc);
            cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
            printf("ldr %3d warps %2d: %.1f SM-cycles per tile task, %.0f cycles per batch per warp (%s)\n", ldr, warps, (double)mx / ((double)reps * ntiles), (double)mx / ((double)reps * ((ntiles + 4 * warps - 1) / (4 * warps))), cudaGetErrorString(e));
        }
    return 0;
}
