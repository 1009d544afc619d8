// CUDA-core fp32 GEMM core with functor operand loaders and a functor epilogue.
//   C[m][n] = sum_k A(m,k) * B(n,k)
// Used for the fp32-exact stages (GRU projections, covariance, Gauss-Jordan updates, stem) and as the
// validation path of the tensor-core conv.  Loaders return 4 consecutive elements:
//   A loader:  float4 operator()(int m, int k)   -- k % 4 == 0, 4 consecutive k, caller guarantees m < M, k < K
//   B loader:  if (BL::n_major) float4 operator()(int n, int k) -- 4 consecutive n (n % 4 == 0) at one k
//              else             float4 operator()(int n, int k) -- 4 consecutive k at one n
//   epilogue:  void operator()(int m, int n, float4 acc)        -- n % 4 == 0, m < M, n < N
// Requirements: K % 4 == 0, N % 4 == 0.
#pragma once
#include <cuda_runtime.h>

template <int T>
struct SgemmCfg {
    static constexpr int BM = 16 * T, BN = 16 * T, BK = 16, LD = 16 * T + 4;
};

template <int T, class AL, class BL, class EP>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, AL al, BL bl, EP ep) {
    using C = SgemmCfg<T>;
    __shared__ __align__(16) float As[C::BK][C::LD];
    __shared__ __align__(16) float Bs[C::BK][C::LD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * C::BM, n0 = blockIdx.x * C::BN;
    constexpr int G = T / 4;                        // float4 groups per thread per dimension
    float acc[T][T];
#pragma unroll
    for (int i = 0; i < T; i++)
#pragma unroll
        for (int j = 0; j < T; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += C::BK) {
        // ---- A tile: BM rows x 16 k  (BM*4 float4, 256 threads)
#pragma unroll
        for (int it = 0; it < C::BM * 4 / 256; it++) {
            int idx = tid + it * 256;
            int row = idx >> 2, kq = (idx & 3) * 4;
            int m = m0 + row, k = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && k < K) v = al(m, k);
            As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
        }
        // ---- B tile
        if constexpr (BL::n_major) {
#pragma unroll
            for (int it = 0; it < C::BN * 4 / 256; it++) {
                int idx = tid + it * 256;
                int kk = idx / (C::BN / 4), nq = (idx % (C::BN / 4)) * 4;
                int n = n0 + nq, k = k0 + kk;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < N && k < K) v = bl(n, k);
                *reinterpret_cast<float4*>(&Bs[kk][nq]) = v;
            }
        } else {
#pragma unroll
            for (int it = 0; it < C::BN * 4 / 256; it++) {
                int idx = tid + it * 256;
                int row = idx >> 2, kq = (idx & 3) * 4;
                int n = n0 + row, k = k0 + kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < N && k < K) v = bl(n, k);
                Bs[kq + 0][row] = v.x; Bs[kq + 1][row] = v.y; Bs[kq + 2][row] = v.z; Bs[kq + 3][row] = v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < C::BK; kk++) {
            float a[T], b[T];
#pragma unroll
            for (int g = 0; g < G; g++) {
                float4 va = *reinterpret_cast<const float4*>(&As[kk][g * 64 + ty * 4]);
                float4 vb = *reinterpret_cast<const float4*>(&Bs[kk][g * 64 + tx * 4]);
                a[g * 4 + 0] = va.x; a[g * 4 + 1] = va.y; a[g * 4 + 2] = va.z; a[g * 4 + 3] = va.w;
                b[g * 4 + 0] = vb.x; b[g * 4 + 1] = vb.y; b[g * 4 + 2] = vb.z; b[g * 4 + 3] = vb.w;
            }
#pragma unroll
            for (int i = 0; i < T; i++)
#pragma unroll
                for (int j = 0; j < T; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int gi = 0; gi < G; gi++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int m = m0 + gi * 64 + ty * 4 + i;
            if (m >= M) continue;
#pragma unroll
            for (int gj = 0; gj < G; gj++) {
                int n = n0 + gj * 64 + tx * 4;
                if (n >= N) continue;
                ep(m, n, make_float4(acc[gi * 4 + i][gj * 4 + 0], acc[gi * 4 + i][gj * 4 + 1],
                                     acc[gi * 4 + i][gj * 4 + 2], acc[gi * 4 + i][gj * 4 + 3]));
            }
        }
}

// ---- common loaders / epilogues -------------------------------------------------------------------
struct LoadRowMajorK {              // element (r,k) at p[r*ld + k], k contiguous; ld % 4 == 0
    static constexpr bool n_major = false;
    const float* p; int64_t ld;
    __device__ float4 operator()(int r, int k) const { return *reinterpret_cast<const float4*>(p + (int64_t)r * ld + k); }
};
struct LoadColMajorN {              // element (n,k) at p[k*ld + n], n contiguous (B operand only)
    static constexpr bool n_major = true;
    const float* p; int64_t ld;
    __device__ float4 operator()(int n, int k) const { return *reinterpret_cast<const float4*>(p + (int64_t)k * ld + n); }
};
struct StoreRowMajor {              // C[m*ld + n] = alpha*acc (+ bias[n])
    float* c; int64_t ld; const float* bias; float alpha;
    __device__ void operator()(int m, int n, float4 v) const {
        float4 o = make_float4(v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha);
        if (bias) { o.x += bias[n]; o.y += bias[n + 1]; o.z += bias[n + 2]; o.w += bias[n + 3]; }
        *reinterpret_cast<float4*>(c + (int64_t)m * ld + n) = o;
    }
};

template <int T, class AL, class BL, class EP>
static inline void sgemm_launch(int M, int N, int K, AL al, BL bl, EP ep, cudaStream_t st) {
    dim3 grid((N + 16 * T - 1) / (16 * T), (M + 16 * T - 1) / (16 * T));
    sgemm_kernel<T, AL, BL, EP><<<grid, 256, 0, st>>>(M, N, K, al, bl, ep);
}
