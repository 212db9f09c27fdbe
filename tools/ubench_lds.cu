// LDS.128 throughput without arithmetic in the way (ubench_smem's loop was bound by its own FADDs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_lds tools/ubench_lds.cu
// Patterns: 0 linear 512 B; 1 K2b tile pattern (quarter-warp = one tile: 4 leaves x 2 halves of ONE row, rows 0..3);
// 2 same, all four tiles on the same row (broadcast between quarters); 3 rows 0,4,8,12; 4 rows 0,1,2,3 with pitch 304;
// 5 two tiles share a row (rows 0,0,1,1); 6 LDS.64 linear; 7 LDS.32 linear
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int pat, int iters, int pitch, unsigned *out, long long *cyc, int r0 = 0, int r1 = 0, int r2 = 0, int r3 = 0)
{
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, q = lane >> 3, sub = lane & 7, l = sub >> 1, h = sub & 1;
    int off;
    switch (pat) {
    case 0: off = lane * 4; break;
    case 1: off = q * pitch + 72 * l + 4 * h; break;
    case 2: off = 72 * l + 4 * h; break;
    case 3: off = 4 * q * pitch + 72 * l + 4 * h; break;
    case 4: off = q * 304 + 72 * l + 4 * h; break;
    case 5: off = (q >> 1) * pitch + 72 * l + 4 * h; break;
    case 6: off = lane * 2; break;
    case 8: off = (q == 0 ? r0 : q == 1 ? r1 : q == 2 ? r2 : r3) * pitch + 72 * l + 4 * h; break;
    default: off = lane; break;
    }
    unsigned a0 = (unsigned)__cvta_generic_to_shared(sm + off);
    unsigned x = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            unsigned r0, r1, r2, r3;
            if (pat == 6) { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a0 + 32 * kk)); r2 = r3 = 0; }
            else if (pat == 7) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r0) : "r"(a0 + 32 * kk)); r1 = r2 = r3 = 0; }
            else asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a0 + 32 * kk));
            x ^= r0;                      // one LOP per load keeps the result alive
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main()
{
    unsigned *out; long long *cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4000;
    for (int warps : { 4, 12, 16, 32 })
        for (int pat = 0; pat <= 7; ++pat) {
            k<<<1, 32 * warps, 64 * 1024>>>(pat, iters, 300, out, cyc);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("pattern %d warps %2d: %.2f SM-cycles per warp-wide load\n", pat, warps, (double)c / (iters * 8.0 * warps));
        }
    const int combos[][4] = { {0,0,0,1}, {0,1,2,0}, {9,10,11,9}, {0,5,0,5}, {3,3,8,8}, {1,2,3,4}, {0,2,4,6}, {0,8,16,24}, {5,5,5,5}, {0,3,6,9}, {7,12,17,22} };
    for (auto &c : combos) {
        k<<<1, 32 * 12, 64 * 1024>>>(8, iters, 300, out, cyc, c[0], c[1], c[2], c[3]);
        cudaDeviceSynchronize();
        long long cc; cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("rows {%d,%d,%d,%d} pitch 300, 12 warps: %.2f SM-cycles per LDS.128\n", c[0], c[1], c[2], c[3], (double)cc / (iters * 8.0 * 12));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
