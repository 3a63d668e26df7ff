// Co-issue micro-benchmark: pairs of instruction types interleaved 1:1 on independent register chains, 4 warps per sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pairs pairs.cu ; run on the GPU box: ./pairs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { VMAX, PRMT, IADD3, LOP3, IMAD, DP4A, SHFL, VMAXP, NOPS };
static const char* names[NOPS] = {"vmax.s16x2", "prmt", "iadd3", "lop3", "imad", "dp4a", "shfl", "vmax+preds"};

template <int OP> __device__ __forceinline__ void op(uint32_t& r, uint32_t c, uint32_t m, uint32_t& side)
{
    if (OP == VMAX) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r) : "r"(m));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0xfbd9;" : "+r"(r) : "r"(c));
    if (OP == IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(r) : "r"(c), "r"(m));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xd8;" : "+r"(r) : "r"(c), "r"(m));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(c), "r"(m));
    if (OP == DP4A) asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(r) : "r"(c), "r"(m));
    if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r));
    if (OP == VMAXP) asm volatile("{.reg .pred pu, pv; .reg .s16 a, b, c, d; max.s16x2 %0, %0, %2; mov.b32 {a, b}, %0; mov.b32 {c, d}, %2; setp.eq.s16 pv, a, c; setp.eq.s16 pu, b, d; @pv add.u32 %1, %1, 1; @pu add.u32 %1, %1, 2;}" : "+r"(r), "+r"(side) : "r"(m));
}

template <int A, int B, int NA, int NB>
__global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    uint32_t r[8], q[8], side = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = seed * (threadIdx.x + 1) + i * 0x01010101u; q[i] = r[i] ^ 0x5a5a5a5au; }
    uint32_t c = seed | 1u, m = seed * 3u + 1u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int x = 0; x < NA; x++) op<A>(r[(i + x) & 7], c, r[(i + 3) & 7] | 1u, side);
#pragma unroll
                for (int x = 0; x < NB; x++) op<B>(q[(i + x) & 7], c, m, side);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = side;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int A, int B, int NA, int NB> void run(uint32_t* d_out, long long* d_cyc)
{
    const int iters = 2000;
    for (int w = 1; w <= 4; w += 3) {
        k<A, B, NA, NB><<<1, 128 * w>>>(d_out, d_cyc, iters, 12345u);
        k<A, B, NA, NB><<<1, 128 * w>>>(d_out, d_cyc, iters, 12345u);
        long long c = 0;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double groups = (double)iters * 32 * w;
        printf("%d x %-12s + %d x %-6s %d warps/sub-partition: %.2f cycles per group of %d\n", NA, names[A], NB, names[B], w, c / groups, NA + NB);
    }
}

int main()
{
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 1 << 20); cudaMalloc(&d_cyc, 1024);
    run<VMAX, IMAD, 1, 1>(d_out, d_cyc); run<PRMT, IMAD, 1, 1>(d_out, d_cyc); run<IADD3, IMAD, 1, 1>(d_out, d_cyc); run<LOP3, IMAD, 1, 1>(d_out, d_cyc);
    run<DP4A, IMAD, 1, 1>(d_out, d_cyc); run<DP4A, LOP3, 1, 1>(d_out, d_cyc); run<SHFL, IMAD, 1, 2>(d_out, d_cyc); run<SHFL, LOP3, 1, 2>(d_out, d_cyc);
    run<VMAXP, IMAD, 1, 1>(d_out, d_cyc); run<VMAXP, IMAD, 1, 3>(d_out, d_cyc); run<VMAX, IMAD, 1, 2>(d_out, d_cyc); run<VMAX, LOP3, 1, 1>(d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
