// hbm_mix.cu -- what HBM gives for the traffic mixes of the pixel kernels (sm_100a): copy 1:1 (the measured-peak pattern),
// read only, write only, and 1 byte read : 2 bytes written (k_units: compressed payload in, 16-bit pixels out), with the
// stores issued as 16-byte st.global by every lane or as 2 KiB bulk stores (TMA) from shared memory; the tiled variant writes
// 512-byte pieces at a row pitch like k_units does.  Buffers are far larger than L2; every byte is touched once per launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hbm_mix tools/ubench/hbm_mix.cu && gpurun_out/hbm_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    uint4 a = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i]; a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
    }
    if (a.x == 0x12345678u && a.y == 1u) out[0] = a;
}
template <int CS>
__global__ void __launch_bounds__(256) k_write(uint4* __restrict__ out, size_t n) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (CS) __stcs(out + i, v); else out[i] = v;
    }
}
// 1 : 2 -- thread i reads in[i] and writes out[2i], out[2i + 1] as two coalesced rows of the warp
template <int CS>
__global__ void __launch_bounds__(256) k_mix12(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t lane = threadIdx.x & 31;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i];
        const size_t w = (i - lane) * 2;
        if (CS) { __stcs(out + w + lane, v); __stcs(out + w + 32 + lane, v); }
        else { out[w + lane] = v; out[w + 32 + lane] = v; }
    }
}
// 1 : 2, k_units-like: a warp reads 4 KiB in a row and writes 8 KiB as 4 rows x 2 KiB (512 B per store instruction) at a pitch
__global__ void __launch_bounds__(128) k_mix12_tiled(const uint4* __restrict__ in, uint4* __restrict__ out, size_t units, size_t pitch16, size_t units_per_row) {
    const size_t lane = threadIdx.x & 31, warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t u = warp; u < units; u += nw) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = in[u * 256 + k * 32 + lane];
        const size_t ty = u / units_per_row, ux = u % units_per_row;
        uint4* o = out + (ty * 4) * pitch16 + ux * 128;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 4; k++) o[r * pitch16 + k * 32 + lane] = v[(r * 4 + k) & 7];
    }
}
// write only through shared memory + 2 KiB bulk stores (one elected lane per warp), 8 KiB per warp per round
__global__ void __launch_bounds__(128) k_write_bulk(uint8_t* __restrict__ out, size_t chunks) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* my = sm + wid * 8192;
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t c = warp; c < chunks; c += nw) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) reinterpret_cast<uint4*>(my)[k * 32 + lane] = make_uint4(lane, k, (uint32_t)c, 7);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(out + c * 8192 + k * 2048), "r"(my_s + k * 2048) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// 1 : 2 through shared memory: 4 KiB bulk load, 8 KiB written by 2 KiB bulk stores (the data is stored twice)
__global__ void __launch_bounds__(128) k_mix12_bulk(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t chunks) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bars[4];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* my = sm + wid * 8192;
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my), bar = (uint32_t)__cvta_generic_to_shared(&bars[wid]);
    if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    uint32_t ph = 0;
    for (size_t c = warp; c < chunks; c += nw) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;" ::"r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 4096, [%2];" ::"r"(my_s), "l"(in + c * 4096), "r"(bar) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(ph & 1u) : "memory");
        ph++;
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(out + c * 8192 + k * 2048), "r"(my_s + (k & 1) * 2048) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t IN = (size_t)512 << 20, OUT = (size_t)1024 << 20;          // bytes
    uint8_t *in, *out;
    CK(cudaMalloc(&in, OUT)); CK(cudaMalloc(&out, OUT + (1 << 20)));
    CK(cudaMemset(in, 1, OUT)); CK(cudaMemset(out, 2, OUT));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    CK(cudaFuncSetAttribute(k_write_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaFuncSetAttribute(k_mix12_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    auto rep = [&](const char* name, float ms, double bytes) { printf("%-44s %8.4f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    for (int cps : {4, 8}) {
        const int g = sms * cps;
        printf("-- %d CTAs per SM (256 threads)\n", cps);
        rep("copy 1:1 (1 GiB read + 1 GiB written)", time_ms([&] { k_copy<<<g, 256>>>((const uint4*)in, (uint4*)out, OUT / 16); }, 8), 2.0 * OUT);
        rep("read only (1 GiB)", time_ms([&] { k_read<<<g, 256>>>((const uint4*)in, (uint4*)out, OUT / 16); }, 8), 1.0 * OUT);
        rep("write only (1 GiB, st.global.v4)", time_ms([&] { k_write<0><<<g, 256>>>((uint4*)out, OUT / 16); }, 8), 1.0 * OUT);
        rep("write only (1 GiB, st.global.cs.v4)", time_ms([&] { k_write<1><<<g, 256>>>((uint4*)out, OUT / 16); }, 8), 1.0 * OUT);
        rep("mix 1:2 (0.5 GiB read + 1 GiB written)", time_ms([&] { k_mix12<0><<<g, 256>>>((const uint4*)in, (uint4*)out, IN / 16); }, 8), 1.0 * IN + OUT);
        rep("mix 1:2, st.global.cs", time_ms([&] { k_mix12<1><<<g, 256>>>((const uint4*)in, (uint4*)out, IN / 16); }, 8), 1.0 * IN + OUT);
    }
    for (int cps : {3, 4, 6}) {
        const int g = sms * cps;
        printf("-- %d CTAs per SM (128 threads)\n", cps);
        // a 4096-pixel image: rows of 8 KiB, four units side by side in a tile row
        rep("mix 1:2 tiled (4 rows x 2 KiB at pitch 8 KiB)", time_ms([&] { k_mix12_tiled<<<g, 128>>>((const uint4*)in, (uint4*)out, IN / 4096, 8192 / 16, 4); }, 8), 1.0 * IN + (double)(IN / 4096) * 8192);
        rep("write only, 2 KiB bulk stores from smem", time_ms([&] { k_write_bulk<<<g, 128, 32768>>>(out, OUT / 8192); }, 8), 1.0 * OUT);
        rep("mix 1:2, bulk load 4 KiB + 4 bulk stores 2 KiB", time_ms([&] { k_mix12_bulk<<<g, 128, 32768>>>(in, out, IN / 4096); }, 8), 1.0 * IN + OUT);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
