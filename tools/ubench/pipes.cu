// pipes.cu -- issue-rate microbenchmark for the integer instructions the legacy decode leans on (sm_100a):
// which pipe takes SHF / LOP3 / PRMT / IMAD / IMAD.HI / IMAD.WIDE, and whether an ALU + FMA mix issues at 1 IPC per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipes tools/ubench/pipes.cu && gpurun_out/pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define U 16
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned a, unsigned b, unsigned long long* cyc) {
    unsigned x[U];
#pragma unroll
    for (int i = 0; i < U; i++) x[i] = threadIdx.x * 17u + i + a;
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            unsigned v = x[i];
            if (MODE == 0) asm volatile("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(v) : "r"(v), "r"(b));
            if (MODE == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(v) : "r"(v), "r"(b), "r"(a));
            if (MODE == 2) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(a), "r"(b));
            if (MODE == 3) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(b), "r"(a));
            if (MODE == 4) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(b));
            if (MODE == 5) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(v), "r"(b)); v = (unsigned)w ^ (unsigned)(w >> 32); }
            if (MODE == 6) {   // SHF + IMAD pair (independent of each other)
                unsigned u2;
                asm volatile("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(u2) : "r"(v), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(b), "r"(u2));
            }
            if (MODE == 7) {   // mul.hi + LOP3 (the proposed sample extraction)
                unsigned u2;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(u2) : "r"(v), "r"(b));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(v) : "r"(u2), "r"(a), "r"(v));
            }
            if (MODE == 8) {   // SHF + LOP3 (today's sample extraction)
                unsigned u2;
                asm volatile("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(u2) : "r"(v), "r"(b));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(v) : "r"(u2), "r"(a), "r"(v));
            }
            if (MODE == 9) asm volatile("shl.b32 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(b));
            if (MODE == 10) asm volatile("add.u32 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(b));
            if (MODE == 11) { asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %2; selp.u32 %0, %3, %1, p;}" : "=r"(v) : "r"(v), "r"(b), "r"(a)); }
            if (MODE == 12) {  // mul.wide + LOP3 (lo | hi) & mask: a rotate on the FMA pipe
                unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(v), "r"(b));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(v) : "r"((unsigned)w), "r"((unsigned)(w >> 32)), "r"(a));
            }
            x[i] = v;
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < U; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int per_iter, unsigned* d_out, unsigned long long* d_cyc) {
    for (int warps = 4; warps <= 32; warps *= 2) {      // warps per SM (one CTA per SM)
        k<MODE><<<148, warps * 32>>>(d_out, 3, 5, d_cyc);
        k<MODE><<<148, warps * 32>>>(d_out, 3, 5, d_cyc);
        cudaDeviceSynchronize();
        unsigned long long c = 0;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double inst = (double)ITERS * U * per_iter * (warps / 4.0);   // warp instructions per SMSP
        printf("%-28s warps/SM %2d  cycles %9llu  IPC/SMSP %.3f\n", name, warps, c, inst / (double)c);
    }
}

int main() {
    unsigned* d_out; unsigned long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 8);
    run<0>("shf.r.wrap", 1, d_out, d_cyc);
    run<1>("lop3", 1, d_out, d_cyc);
    run<2>("prmt", 1, d_out, d_cyc);
    run<3>("mad.lo (IMAD)", 1, d_out, d_cyc);
    run<4>("mul.hi (IMAD.HI)", 1, d_out, d_cyc);
    run<5>("mul.wide + xor", 2, d_out, d_cyc);
    run<6>("shf + imad", 2, d_out, d_cyc);
    run<7>("mul.hi + lop3", 2, d_out, d_cyc);
    run<8>("shf + lop3", 2, d_out, d_cyc);
    run<9>("shl", 1, d_out, d_cyc);
    run<10>("add", 1, d_out, d_cyc);
    run<11>("setp+selp", 2, d_out, d_cyc);
    run<12>("mul.wide + lop3", 2, d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
