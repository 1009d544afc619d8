// Micro-benchmark (GPU box): how fast can epilogue warps drain tensor memory?  Decides whether the conv can afford a
// two-level accumulation (short tcgen05 accumulation chains in TMEM, summed into fp32 registers with round-to-nearest).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tools/probes/tmem_bw.cu && ./tmem_bw
// One CTA per SM, 512 TMEM columns; W warps (4 or 8) each loop over tcgen05.ld.32x32b.x32 of their lane quarter and
// (optionally) add the values into 128 register accumulators.  Reports bytes/clock/SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: ld only (values xor-folded so the loads stay live); MODE 1: ld + 128 running fp32 sums per thread
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_tmem_bw(int iters, int ncols_per_warp, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const int q = warp & 3, half = warp >> 2;                       // lane quarter, column half (8-warp runs)
    const uint32_t lane_addr = base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * ncols_per_warp);
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; i++) acc[i] = 0.f;
    uint32_t fold = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c += 2) {                            // 2 loads (64 columns) in flight per wait
            uint32_t v[32], u[32];
            tmem_ld32_nowait(lane_addr + (uint32_t)((c * 32) % ncols_per_warp), v);
            tmem_ld32_nowait(lane_addr + (uint32_t)(((c + 1) * 32) % ncols_per_warp), u);
            tmem_wait_ld();
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 32; i++) fold ^= v[i] ^ u[i];
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) { acc[c * 32 + i] += __uint_as_float(v[i]); acc[(c + 1) * 32 + i] += __uint_as_float(u[i]); }
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = __uint_as_float(fold);
#pragma unroll
    for (int i = 0; i < 128; i++) s += acc[i];
    if (s == 1234.5f) sink[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main() {
    long long* cyc;
    float* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 1024);
    const int iters = 2000;
    for (int mode = 0; mode < 2; mode++)
        for (int warps = 4; warps <= 8; warps += 4) {
            const int ncols = 128;                                  // every warp streams over 128 columns of its lane quarter
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k_tmem_bw<0><<<148, warps * 32, 0>>>(iters, ncols, cyc, sink);
                else k_tmem_bw<1><<<148, warps * 32, 0>>>(iters, ncols, cyc, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < 148; i++) mean += (double)h[i] / 148;
            const double bytes = (double)iters * 4 * 32 * 32 * 4 * warps;       // per SM
            printf("mode %d (%s) warps %d: %.0f cycles, %.1f bytes/clock/SM, %.1f clocks per 32x32b.x32 load per warp\n", mode,
                   mode ? "ld + 128 register sums" : "ld only", warps, mean, bytes / mean, mean / (iters * 4.0));
        }
    return 0;
}
