// Issue-rate micro-benchmark for the instruction types of k_viterbi on one SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_VMAX, OP_PRMT, OP_IADD3, OP_LOP3, OP_IMAD, OP_DP4A, OP_SHFL, OP_MIX, OP_IADD, N_OPS };
static const char* names[N_OPS] = {"vmax.s16x2", "prmt", "iadd3", "lop3", "imad", "dp4a", "shfl.bfly", "vmax+imad 1:1", "add.s32 (compiler's choice)"};

template <int OP>
__global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    uint32_t c = seed | 1u, m = seed * 3u + 1u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == OP_VMAX) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 3) & 7]));
                if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0xfbd9;" : "+r"(r[i]) : "r"(c));
                if (OP == OP_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(r[i]) : "r"(c), "r"(m));
                if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xd8;" : "+r"(r[i]) : "r"(c), "r"(m));
                if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c), "r"(m));
                if (OP == OP_DP4A) asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c), "r"(m));
                if (OP == OP_SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r[i]));
                if (OP == OP_MIX) { if (i & 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 2) & 7])); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c), "r"(m)); }
                if (OP == OP_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(c));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(uint32_t* d_out, long long* d_cyc)
{
    const int iters = 2000;
    for (int w = 1; w <= 4; w++) {                       // warps per sub-partition (block = 128 * w threads on one SM)
        k<OP><<<1, 128 * w>>>(d_out, d_cyc, iters, 12345u);
        k<OP><<<1, 128 * w>>>(d_out, d_cyc, iters, 12345u);
        long long c = 0;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double per_sp = (double)iters * 64 * w;     // warp instructions issued per sub-partition
        printf("%-28s %d warps/sub-partition: %.3f warp-instr/cycle/sub-partition (%.2f cycles per instruction)\n", names[OP], w, per_sp / c, c / per_sp);
    }
}

int main()
{
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 1 << 20); cudaMalloc(&d_cyc, 1024);
    run<OP_VMAX>(d_out, d_cyc); run<OP_PRMT>(d_out, d_cyc); run<OP_IADD3>(d_out, d_cyc); run<OP_LOP3>(d_out, d_cyc);
    run<OP_IMAD>(d_out, d_cyc); run<OP_DP4A>(d_out, d_cyc); run<OP_SHFL>(d_out, d_cyc); run<OP_MIX>(d_out, d_cyc); run<OP_IADD>(d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
