// Micro-benchmarks used while designing the cost-tile kernel (run on the B200 via gpurun):
//  (1) LDS.128 wavefront cost of candidate lane->address maps, (2) FADD2/FFMA2 vs FADD/FFMA issue rate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__global__ void lds_pat(int pat, int iters, float *out, long long *cyc)
{
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < 48 * 320; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int off;   // float offset
    const int q = lane >> 3, sub = lane & 7, l = sub >> 1, h = sub & 1;
    switch (pat) {
    case 0: off = lane * 4; break;                                   // linear 512 B
    case 1: off = q * 300 + 72 * l + 4 * h; break;                   // kernel v2 pattern (4 tiles, 4 leaves x 2 halves)
    case 2: off = q * 300 + sub * 4; break;                          // tile reads 128 contiguous bytes of its row
    case 3: off = 72 * l + 4 * h; break;                             // all tiles same row (broadcast across quarters)
    case 4: off = q * 304 + 72 * l + 4 * h; break;                   // row stride 304
    case 5: off = (lane >> 1) * 300 + 4 * h; break;                  // v1 pattern: 16 rows x 2 halves
    case 6: off = q * 312 + 72 * l + 4 * h; break;                   // row stride 312
    case 7: off = 0; break;                                          // full broadcast
    case 8: off = q * 320 + sub * 4; break;                          // 4 rows, same banks: linear per quarter
    case 9: off = q * 300 + 64 * l + 4 * h; break;                   // leaf stride 64 floats (bank-aligned) -> 4-way within tile
    default: off = (lane & 1) * 4 + (lane >> 1) * 8 * 0 + (lane>>1) * 300; break;
    }
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned a0 = (unsigned)__cvta_generic_to_shared(sm + off);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4 v = lds128(a0 + 32 * k);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
template <int MODE>
__global__ void fp_rate(int iters, float *out, long long *cyc, unsigned long long nz, float c)
{
    float a[8]; unsigned long long b[8];
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x + k; b[k] = 0x3f8000003f800000ull + threadIdx.x + k; }
    unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(c));
            else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[k]) : "f"(c));
            else if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[k]) : "l"(c2));
            else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b[k]) : "l"(c2), "l"(nz));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0; for (int k = 0; k < 8; ++k) s += a[k] + (float)(b[k] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *out; long long *cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
    const int iters = 4000;
    cudaFuncSetAttribute(lds_pat, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int warps = 1; warps <= 16; warps *= 16)
        for (int pat = 0; pat <= 10; ++pat) {
            lds_pat<<<1, 32 * warps, 64 * 1024>>>(pat, iters, out, cyc);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("lds pattern %2d warps %2d: %.2f SM-cycles per warp LDS.128\n", pat, warps, (double)c / (iters * 8.0 * warps));
        }
    for (int mode = 0; mode < 4; ++mode) {
        const int nthreads = 512;                                                  // 16 warps = 4 per SMSP
        if (mode == 0) fp_rate<0><<<1, nthreads>>>(iters, out, cyc, 0x8000000080000000ull, 1.0001f);
        if (mode == 1) fp_rate<1><<<1, nthreads>>>(iters, out, cyc, 0x8000000080000000ull, 1.0001f);
        if (mode == 2) fp_rate<2><<<1, nthreads>>>(iters, out, cyc, 0x8000000080000000ull, 1.0001f);
        if (mode == 3) fp_rate<3><<<1, nthreads>>>(iters, out, cyc, 0x8000000080000000ull, 1.0001f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("fp mode %d (0 FADD 1 FFMA 2 FADD2 3 FFMA2): %.3f cycles per warp-instr per SMSP\n", mode, (double)c / (iters * 8.0 * 4.0));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
