// Geometry stages: CA distance map (network.py:272 / predict.py:143), the steric/bond minimiser
// (network.py:106-137), the backbone build (network.py:141-177) and the on-device best-of-n selection
// (network.py:302-306).
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_fill(float* __restrict__ p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int run_fill(dmp2_engine* e, float* p, int64_t n, float v, cudaStream_t st) {
    k_fill<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(p, n, v);
    POST_LAUNCH(e, "k_fill");
    return 0;
}

// dmap[i][j] = sqrt(clamp(|ca_j - ca_i|^2, 1e-8))  (recycling, network.py:272) or unclamped (template, predict.py:143)
__global__ void k_dmap(const float* __restrict__ ca, int L, int clamp, float* __restrict__ dmap) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * L) return;
    int i = idx / L, j = idx - i * L;
    float dx = __fsub_rn(ca[j * 3], ca[i * 3]), dy = __fsub_rn(ca[j * 3 + 1], ca[i * 3 + 1]), dz = __fsub_rn(ca[j * 3 + 2], ca[i * 3 + 2]);
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (clamp) s = fmaxf(s, 1e-8f);
    dmap[idx] = sqrtf(s);
}
int run_dmap(dmp2_engine* e, const float* ca, int L, float* dmap, bool clamp, cudaStream_t st) {
    k_dmap<<<cdiv(L * L, 256), 256, 0, st>>>(ca, L, clamp ? 1 : 0, dmap);
    POST_LAUNCH(e, "k_dmap");
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// refine_coords: one 8-CTA cluster, every CTA keeps the full trace double-buffered in shared memory and
// integrates a contiguous eighth of the atoms with S threads per atom; the new positions are broadcast to
// the peers through distributed shared memory, one cluster barrier per step.  No L x L x 3 tensors, no
// autograd.  Pairs at >= 3.0 A contribute exactly zero force in the reference (violate = 0), so they are
// skipped on the squared distance.
// ---------------------------------------------------------------------------------------------------
#define REFINE_CL 8
__global__ void __cluster_dims__(REFINE_CL, 1, 1) __launch_bounds__(1024, 1)
k_refine(float* __restrict__ ca, int L, int steps, int per, int S) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ float sh[];
    float* buf[2] = {sh, sh + 3 * L};
    float* part = sh + 6 * L;                         // [per][S][3] partial accelerations
    for (int i = threadIdx.x; i < 3 * L; i += blockDim.x) { sh[i] = ca[i]; sh[3 * L + i] = ca[i]; }
    cluster.sync();
    const int jl = threadIdx.x / S, sub = threadIdx.x - jl * S;
    const int j = c * per + jl;
    const bool act = jl < per && j < L;
    for (int s = 0; s < steps; s++) {
        const float* cc = buf[s & 1];
        float* o = buf[(s & 1) ^ 1];
        float ax = 0.f, ay = 0.f, az = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
        if (act) {
            cx = cc[3 * j]; cy = cc[3 * j + 1]; cz = cc[3 * j + 2];
            for (int i = sub; i < L; i += S) {
                float dx = cx - cc[3 * i], dy = cy - cc[3 * i + 1], dz = cz - cc[3 * i + 2];
                float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d2 < 9.0f) {
                    float d = fmaxf(sqrtf(d2), 0.01f);
                    float f = 100.0f * (3.0f - d);
                    ax += f * (dx / d); ay += f * (dy / d); az += f * (dz / d);
                }
            }
            float* pp = part + (jl * S + sub) * 3;
            pp[0] = ax; pp[1] = ay; pp[2] = az;
        }
        __syncthreads();
        if (act && sub == 0) {
            ax = 0.f; ay = 0.f; az = 0.f;
            for (int q = 0; q < S; q++) { const float* pp = part + (jl * S + q) * 3; ax += pp[0]; ay += pp[1]; az += pp[2]; }
            if (j < L - 1) {              // bond to j+1: accels[j] += acov_j
                float ux = cc[3 * j + 3] - cx, uy = cc[3 * j + 4] - cy, uz = cc[3 * j + 5] - cz;
                float d = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz))), 0.1f);
                float f = 100.0f * fminf(d - 3.78f, 3.0f);
                ax += f * (ux / d); ay += f * (uy / d); az += f * (uz / d);
            }
            if (j > 0) {                  // bond from j-1: accels[j] -= acov_{j-1}
                float ux = cx - cc[3 * j - 3], uy = cy - cc[3 * j - 2], uz = cz - cc[3 * j - 1];
                float d = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz))), 0.1f);
                float f = 100.0f * fminf(d - 3.78f, 3.0f);
                ax -= f * (ux / d); ay -= f * (uy / d); az -= f * (uz / d);
            }
            const float nx = cx + fminf(fmaxf(ax, -100.f), 100.f) * 0.001f;
            const float ny = cy + fminf(fmaxf(ay, -100.f), 100.f) * 0.001f;
            const float nz = cz + fminf(fmaxf(az, -100.f), 100.f) * 0.001f;
#pragma unroll
            for (int d = 0; d < REFINE_CL; d++) {
                float* r = cluster.map_shared_rank(o, d);
                r[3 * j] = nx; r[3 * j + 1] = ny; r[3 * j + 2] = nz;
            }
        }
        cluster.sync();
    }
    if (c == 0) {
        const float* cc = buf[steps & 1];
        for (int i = threadIdx.x; i < 3 * L; i += blockDim.x) ca[i] = cc[i];
    }
}

int run_refine(dmp2_engine* e, float* ca, int L, int steps, cudaStream_t st) {
    if (steps <= 0) return 0;
    if (!e->attr_refine) {
        CUDA_TRY(e, cudaFuncSetAttribute(k_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        e->attr_refine = true;
    }
    const int per = cdiv(L, REFINE_CL);
    if (per > 1024) return e->fail(DMP2_ERR_UNSUPPORTED, "refine: L too large");
    const int S = 1024 / per;                         // threads per atom
    size_t smem = ((size_t)6 * L + (size_t)per * S * 3) * sizeof(float);
    if (smem > 200 * 1024) return e->fail(DMP2_ERR_UNSUPPORTED, "refine: L too large");
    k_refine<<<REFINE_CL, 1024, smem, st>>>(ca, L, steps, per, S);
    POST_LAUNCH(e, "k_refine");
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// calpha_to_main_chain + sigmoid(conf)
// ---------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float norm3(V3 a) {
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z)));
}
__device__ __forceinline__ V3 unit(V3 a) { return a / fmaxf(norm3(a), 1e-12f); }   // F.normalize eps
__device__ __forceinline__ V3 ld3(const float* p, int i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

// extended trace: index -1 and L are the dummy CA atoms of network.py:143-149
__device__ V3 ext_ca(const float* ca, int L, int k) {
    if (k >= 0 && k < L) return ld3(ca, k);
    if (k < 0) {
        V3 a = ld3(ca, 0) - ld3(ca, 1), b = ld3(ca, 2) - ld3(ca, 1);
        return ld3(ca, 0) + unit(cross(a, b)) * 3.82f;
    }
    V3 a = ld3(ca, L - 1) - ld3(ca, L - 2), b = ld3(ca, L - 3) - ld3(ca, L - 2);
    return ld3(ca, L - 1) + unit(cross(a, b)) * 3.82f;
}

__global__ void k_backbone(const float* __restrict__ ca, const float* __restrict__ conf_logit, int L, float* __restrict__ out,
                           float* __restrict__ conf_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    V3 c = ld3(ca, i), prev = ext_ca(ca, L, i - 1), next = ext_ca(ca, L, i + 1);
    V3 vn = prev - c, vc = next - c;
    V3 mid = (c + prev) / 2.0f;
    V3 x = unit(cross(vn, vc));
    V3 at_n = mid - vn / 8.0f + x / 4.0f;
    V3 at_c, at_o;
    if (i < L - 1) {
        V3 nn = ext_ca(ca, L, i + 2);
        V3 vn1 = c - next, vc1 = nn - next;
        V3 mid1 = (next + c) / 2.0f;
        V3 x1 = unit(cross(vn1, vc1));
        at_c = mid1 + vn1 / 8.0f - x1 / 2.0f;
        at_o = mid1 - x1 * 1.8f;
    } else {
        V3 midl = (next + c) / 2.0f;
        at_c = midl - vc / 8.0f + x / 2.0f;
        at_o = midl + x * 2.0f;
    }
    V3 vnca = c - at_n, vcca = c - at_c;
    V3 cr = cross(vnca, vcca);
    V3 bis = vnca + vcca;
    const float ang = 1.5707963267948966f - asinf(1.0f / sqrtf(3.0f));
    const float kx = (float)(1.5 * 0.5773502691896258);     // 1.5*cos(pi/2 - asin(1/sqrt3)) = 1.5/sqrt(3)
    const float ky = (float)(1.5 * 0.816496580927726);      // 1.5*sin(...) = 1.5*sqrt(2/3)
    (void)ang;
    float sx = __fmul_rn(__frcp_rn(norm3(bis)), kx);
    float sy = __fmul_rn(__frcp_rn(norm3(cr)), ky);
    V3 at_cb = c + bis * sx + cr * sy;
    float* o = out + (int64_t)i * 15;
    o[0] = at_n.x; o[1] = at_n.y; o[2] = at_n.z;
    o[3] = c.x; o[4] = c.y; o[5] = c.z;
    o[6] = at_c.x; o[7] = at_c.y; o[8] = at_c.z;
    o[9] = at_o.x; o[10] = at_o.y; o[11] = at_o.z;
    o[12] = at_cb.x; o[13] = at_cb.y; o[14] = at_cb.z;
    if (conf_out) conf_out[i] = 1.0f / (1.0f + expf(-conf_logit[i]));
}

int run_backbone(dmp2_engine* e, const float* ca, const float* conf_logit, int L, float* out, float* conf_out, cudaStream_t st) {
    k_backbone<<<cdiv(L, 128), 128, 0, st>>>(ca, conf_logit, L, out, conf_logit ? conf_out : nullptr);
    POST_LAUNCH(e, "k_backbone");
    return 0;
}

// best-of-n selection on the device (strict '>' on the mean confidence logit, network.py:302-306)
__global__ void __launch_bounds__(256) k_select(const float* __restrict__ ca, const float* __restrict__ conf, int L, int first,
                                                float* __restrict__ best_ca, float* __restrict__ best_conf,
                                                float* __restrict__ best_mean) {
    __shared__ double red[256];
    __shared__ int take;
    double s = 0;
    for (int i = threadIdx.x; i < L; i += 256) s += (double)conf[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float mean = (float)(red[0] / (double)L);
        take = first || (mean > best_mean[0]);
        if (take) best_mean[0] = mean;
    }
    __syncthreads();
    if (!take) return;
    for (int i = threadIdx.x; i < 3 * L; i += 256) best_ca[i] = ca[i];
    for (int i = threadIdx.x; i < L; i += 256) best_conf[i] = conf[i];
}

int run_select(dmp2_engine* e, const float* ca, const float* conf, int L, int first, cudaStream_t st) {
    k_select<<<1, 256, 0, st>>>(ca, conf, L, first, e->ws.best_ca, e->ws.best_conf, e->ws.best_mean);
    POST_LAUNCH(e, "k_select");
    return 0;
}
