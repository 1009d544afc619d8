// 2-D track: stem (network.py:194 Maxout2d 955->128, pool 3, k=1), the per-block InstanceNorm + scSE gate +
// residual pass (network.py:32, :37-81, :100-101), the CUDA-core validation conv, and the head (:207, :237-246).
// Activations are NHWC: pixel p = i*L + j, 128 channels contiguous.
#include "common.cuh"
#include "sgemm.cuh"
#include <cuda_fp8.h>

// x -> fp16 hi, fp16 lo, and the two pre-scaled e4m3 copies used by the FP8 correction terms of the conv
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}

// ---------------------------------------------------------------------------------------------------
// Stem.  The 955-channel input (512 outer-product + 441 DCA + 1 APC + 1 distance map) is never built:
//   base[p][o] = b[o] + sum_c W[o][c] m[c][i] m[c][j] + sum_k W[o][512+k] feat[p][k]          (once per target)
//   pre[p][o]  = base[p][o] + W[o][954] * dmap[p]                                             (per recycle)
// ---------------------------------------------------------------------------------------------------
struct StemLoad {
    static constexpr bool n_major = false;
    const float* m1t;            // [L][512] hgru output (time-major)
    const float* feat;           // [L*L][444]
    int L;
    int m0;                      // first pixel of this engine's rows (0 unless halo-sharded)
    __device__ float4 operator()(int m, int k) const {
        m += m0;
        if (k < 512) {
            int i = m / L, j = m - i * L;
            float4 a = *reinterpret_cast<const float4*>(m1t + (int64_t)i * 512 + k);
            float4 b = *reinterpret_cast<const float4*>(m1t + (int64_t)j * 512 + k);
            return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        }
        return *reinterpret_cast<const float4*>(feat + (int64_t)m * DMP2_FEAT_LD + (k - 512));
    }
};

// The A operand of the stem GEMM for `rows` pixels starting at pixel m0, as the fp16 hi/lo pair the tensor-core GEMM
// reads: [rows][DMP2_STEM_KP], columns 0..511 = outer product, 512..955 = DCA features (+ APC, + 2 zero pads), rest 0.
// Values are multiplied by STEM_SA (a power of two) so that the lo parts stay inside the fp16 range.
constexpr float STEM_SA = 16.0f;
__global__ void __launch_bounds__(256) k_stem_operand(const float* __restrict__ m1t, const float* __restrict__ feat, int L, int64_t m0,
                                                      int rows, __half* __restrict__ hi, __half* __restrict__ lo) {
    constexpr int KQ = DMP2_STEM_KP / 4;
    const int64_t n = (int64_t)rows * KQ;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / KQ), k = (int)(idx % KQ) * 4;
        const int64_t pix = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < 512) {
            const int i = (int)(pix / L), j = (int)(pix - (int64_t)i * L);
            const float4 a = *reinterpret_cast<const float4*>(m1t + (int64_t)i * 512 + k);
            const float4 b = *reinterpret_cast<const float4*>(m1t + (int64_t)j * 512 + k);
            v = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        } else if (k < 512 + DMP2_FEAT_LD) {
            v = *reinterpret_cast<const float4*>(feat + pix * DMP2_FEAT_LD + (k - 512));
        }
        v.x *= STEM_SA; v.y *= STEM_SA; v.z *= STEM_SA; v.w *= STEM_SA;
        __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(hi + (int64_t)r * DMP2_STEM_KP + k) = hv;
        *reinterpret_cast<uint2*>(lo + (int64_t)r * DMP2_STEM_KP + k) = lv;
    }
}

// base384 WITHOUT the bias when the tensor-core GEMM produced it (k_stem_update adds it), with it on the CUDA-core path
int run_stem_base(dmp2_engine* e, const float* mat1d_t, const float* feat444, int L, cudaStream_t st) {
    const Rows rw = rows_of(e, L);
    if (!e->gemm_tc) {
        sgemm_launch<8>(rw.R * L, 384, DMP2_STEM_K, StemLoad{mat1d_t, feat444, L, rw.r0 * L}, LoadRowMajorK{e->w.stem_w, DMP2_STEM_K},
                        StoreRowMajor{e->ws.base384, 384, e->w.stem_b, 1.0f}, st);
        POST_LAUNCH(e, "sgemm<stem>");
        return 0;
    }
    // 66 GFLOP at L=300: the pixels go through the tcgen05 GEMM (fp16 hi/lo split, K=128 accumulation chains) in slabs
    // that fit the operand scratch
    const int64_t npix = (int64_t)rw.R * L, first = (int64_t)rw.r0 * L;
    __half* a_hi = e->ws.tc_scratch;
    __half* a_lo = a_hi + (int64_t)DMP2_TC_SLAB * DMP2_STEM_KP;
    for (int64_t m0 = 0; m0 < npix; m0 += DMP2_TC_SLAB) {
        const int rows = (int)std::min<int64_t>(DMP2_TC_SLAB, npix - m0);
        const int grid = (int)std::min<int64_t>(cdiv64((int64_t)rows * (DMP2_STEM_KP / 4), 256), (int64_t)e->num_sms * 8);
        k_stem_operand<<<grid, 256, 0, st>>>(mat1d_t, feat444, L, first + m0, rows, a_hi, a_lo);
        POST_LAUNCH(e, "k_stem_operand");
        TRY(run_gemm_tc(e, a_hi, a_lo, e->w.stem_w_hi, e->w.stem_w_lo, rows, 384, DMP2_STEM_KP, 1.0f / (STEM_SA * DMP2_STEM_SW),
                        e->ws.base384 + m0 * 384, 384, 128, st));
    }
    return 0;
}

// raw[p][g] = max_{q<3} (base[p][3g+q] + wd[3g+q] * dmap[p])                    (network.py:30-31, pool 3)
__global__ void __launch_bounds__(256) k_stem_update(const float* __restrict__ base, const float* __restrict__ wd,
                                                     const float* __restrict__ bias /* nullptr: already in base */,
                                                     const float* __restrict__ dmap, int64_t npix, float* __restrict__ raw) {
    int g = threadIdx.x & 127;
    float w0 = wd[3 * g], w1 = wd[3 * g + 1], w2 = wd[3 * g + 2];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (bias) { c0 = bias[3 * g]; c1 = bias[3 * g + 1]; c2 = bias[3 * g + 2]; }
    for (int64_t p = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); p < npix; p += (int64_t)gridDim.x * 2) {
        float d = dmap[p];
        const float* b = base + p * 384 + 3 * g;
        float v = fmaxf(fmaxf(fmaf(w0, d, b[0] + c0), fmaf(w1, d, b[1] + c1)), fmaf(w2, d, b[2] + c2));
        raw[p * 128 + g] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// InstanceNorm statistics (biased variance over all L*L pixels per channel, eps 1e-5), deterministic:
// per-CTA fp64 partials, the last CTA to finish folds them in a fixed order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_in_stats(const float* __restrict__ raw, int64_t npix, const float* __restrict__ gamma,
                                                   double* __restrict__ part, unsigned int* __restrict__ ticket,
                                                   float* __restrict__ norm /* [mean 128 | gamma*rstd 128] */,
                                                   double* __restrict__ totals /* halo-sharded: [sum 128 | sumsq 128] of these rows, norm unused */) {
    // 32 lanes x float4 cover one pixel's 128 channels; 16 pixel groups per CTA; 4 loads in flight per thread
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * 16;
    int64_t p = (int64_t)blockIdx.x * 16 + grp;
    for (; p + 7 * stride < npix; p += 8 * stride) {                // 8 independent 16-byte loads in flight per thread
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = *reinterpret_cast<const float4*>(raw + (p + u * stride) * 128 + lane * 4);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
            ss[0] += (double)v[u].x * v[u].x; ss[1] += (double)v[u].y * v[u].y;
            ss[2] += (double)v[u].z * v[u].z; ss[3] += (double)v[u].w * v[u].w;
        }
    }
    for (; p < npix; p += stride) {
        float4 v = *reinterpret_cast<const float4*>(raw + p * 128 + lane * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        ss[0] += (double)v.x * v.x; ss[1] += (double)v.y * v.y; ss[2] += (double)v.z * v.z; ss[3] += (double)v.w * v.w;
    }
    __shared__ double sh[16][256];                    // [group][sum 128 | sumsq 128]
#pragma unroll
    for (int u = 0; u < 4; u++) { sh[grp][lane * 4 + u] = s[u]; sh[grp][128 + lane * 4 + u] = ss[u]; }
    __syncthreads();
    if (threadIdx.x < 256) {
        double t = 0;
#pragma unroll 8
        for (int g = 0; g < 16; g++) t += sh[g][threadIdx.x];
        part[(int64_t)blockIdx.x * 256 + threadIdx.x] = t;
    }
    // Deterministic two-level fold of the per-CTA partials: the last CTA of every group of 16 folds its group (fixed
    // order), the last group to finish folds the group sums.  The serial tail is 16 + #groups loads instead of #CTAs.
    const unsigned G = 16, grp_id = blockIdx.x / G, ngrp = (gridDim.x + G - 1) / G;
    const unsigned gfirst = grp_id * G, gsize = min(G, gridDim.x - gfirst);
    double* part2 = part + (int64_t)gridDim.x * 256;
    __threadfence();
    __shared__ int stage;                              // 0 = not last in group, 1 = last in group, 2 = last overall
    __syncthreads();
    if (threadIdx.x == 0) stage = (atomicAdd(ticket + 1 + grp_id, 1u) == gsize - 1) ? 1 : 0;
    __syncthreads();
    if (stage == 0) return;
    __threadfence();
    if (threadIdx.x < 256) {
        double t = 0;
#pragma unroll 16
        for (unsigned b = 0; b < gsize; b++) t += part[(int64_t)(gfirst + b) * 256 + threadIdx.x];
        part2[(int64_t)grp_id * 256 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        ticket[1 + grp_id] = 0;
        stage = (atomicAdd(ticket, 1u) == ngrp - 1) ? 2 : 1;
    }
    __syncthreads();
    if (stage != 2) return;
    __threadfence();
    if (threadIdx.x < 256) {
        double t = 0;
#pragma unroll 8
        for (unsigned b = 0; b < ngrp; b++) t += part2[(int64_t)b * 256 + threadIdx.x];
        sh[0][threadIdx.x] = t;
    }
    __syncthreads();
    if (totals) {                                      // the cross-rank fold finishes the statistics (strip.cu)
        if (threadIdx.x < 256) totals[threadIdx.x] = sh[0][threadIdx.x];
        if (threadIdx.x == 0) *ticket = 0;
        return;
    }
    if (threadIdx.x < 128) {
        const int c = threadIdx.x;
        double mean = sh[0][c] / (double)npix;
        double var = sh[0][128 + c] / (double)npix - mean * mean;
        if (var < 0) var = 0;
        norm[c] = (float)mean;
        norm[128 + c] = (float)((double)gamma[c] / sqrt(var + 1e-5));
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// y = (raw - mean) * (gamma*rstd) + beta;  stem: x = y;  block: x = y * (gate_c + sigmoid(w_s . y + b_s)) + x
// One warp per pixel, 4 channels per lane.  Also emits the fp16 hi/lo split consumed by the tensor-core conv.
__global__ void __launch_bounds__(256) k_norm_gate(const float* __restrict__ raw, const float* __restrict__ norm,
                                                   const float* __restrict__ beta, const float* __restrict__ gate_c,
                                                   const float* __restrict__ sse_w, float sse_b, int stem, int64_t npix,
                                                   float* __restrict__ x, __half* __restrict__ xh, __half* __restrict__ xl,
                                                   uint8_t* __restrict__ x8lo, uint8_t* __restrict__ x8hi, int write_lo16,
                                                   int write_fp8) {
    const int lane = threadIdx.x & 31;
    const int c = lane * 4;
    const float4 mean = *reinterpret_cast<const float4*>(norm + c);
    const float4 sg = *reinterpret_cast<const float4*>(norm + 128 + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float4 gc = make_float4(0, 0, 0, 0), sw = make_float4(0, 0, 0, 0);
    if (!stem) {
        gc = *reinterpret_cast<const float4*>(gate_c + c);
        sw = *reinterpret_cast<const float4*>(sse_w + c);
    }
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < npix; p += nwarps) {
        float4 r = *reinterpret_cast<const float4*>(raw + p * 128 + c);
        float4 y = make_float4(fmaf(r.x - mean.x, sg.x, be.x), fmaf(r.y - mean.y, sg.y, be.y),
                               fmaf(r.z - mean.z, sg.z, be.z), fmaf(r.w - mean.w, sg.w, be.w));
        float4 o = y;
        if (!stem) {
            float d = y.x * sw.x + y.y * sw.y + y.z * sw.z + y.w * sw.w;
            for (int s = 16; s; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
            float sg2 = 1.0f / (1.0f + expf(-(d + sse_b)));
            float4 xo = *reinterpret_cast<const float4*>(x + p * 128 + c);
            o = make_float4(fmaf(y.x, gc.x + sg2, xo.x), fmaf(y.y, gc.y + sg2, xo.y), fmaf(y.z, gc.z + sg2, xo.z),
                            fmaf(y.w, gc.w + sg2, xo.w));
        }
        *reinterpret_cast<float4*>(x + p * 128 + c) = o;
        __half2 h01 = __floats2half2_rn(o.x, o.y), h23 = __floats2half2_rn(o.z, o.w);
        float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        __half2 l01 = __floats2half2_rn(o.x - f01.x, o.y - f01.y), l23 = __floats2half2_rn(o.z - f23.x, o.w - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(xh + p * 128 + c) = hv;
        if (write_lo16) *reinterpret_cast<uint2*>(xl + p * 128 + c) = lv;            // only the f16x3 conv reads it
        if (write_fp8) {                                                              // only the f16f8 conv reads these
            *reinterpret_cast<uint32_t*>(x8lo + p * 128 + c) =
                pack_e4m3x4((o.x - f01.x) * 256.f, (o.y - f01.y) * 256.f, (o.z - f23.x) * 256.f, (o.w - f23.y) * 256.f);
            *reinterpret_cast<uint32_t*>(x8hi + p * 128 + c) = pack_e4m3x4(f01.x * 0.0625f, f01.y * 0.0625f, f23.x * 0.0625f, f23.y * 0.0625f);
        }
    }
}

__global__ void k_split_half(const float* __restrict__ x, int64_t n, __half* __restrict__ hi, __half* __restrict__ lo,
                             uint8_t* __restrict__ x8lo, uint8_t* __restrict__ x8hi) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    __half h = __float2half_rn(v);
    float hf = __half2float(h);
    hi[i] = h;
    lo[i] = __float2half_rn(v - hf);
    if (x8lo) x8lo[i] = (uint8_t)__nv_cvt_float_to_fp8((v - hf) * 256.f, __NV_SATFINITE, __NV_E4M3);
    if (x8hi) x8hi[i] = (uint8_t)__nv_cvt_float_to_fp8(hf * 0.0625f, __NV_SATFINITE, __NV_E4M3);
}

int run_split_half(dmp2_engine* e, const float* x, int64_t n, __half* hi, __half* lo, uint8_t* x8lo, uint8_t* x8hi, cudaStream_t st) {
    k_split_half<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(x, n, hi, lo, x8lo, x8hi);
    POST_LAUNCH(e, "k_split_half");
    return 0;
}

int run_stem_update(dmp2_engine* e, const float* dmap, int L, cudaStream_t st) {
    Workspace& ws = e->ws;
    const Rows rw = rows_of(e, L);
    const int64_t npix = (int64_t)rw.R * L;
    int grid = (int)std::min<int64_t>(cdiv64(npix, 2), (int64_t)e->num_sms * 8);
    k_stem_update<<<grid, 256, 0, st>>>(ws.base384, e->w.stem_wd, e->gemm_tc ? e->w.stem_b : nullptr, dmap + (int64_t)rw.r0 * L, npix, ws.raw);
    POST_LAUNCH(e, "k_stem_update");
    return run_norm_gate(e, -1, ws.raw, ws.x, L, true, st);
}

int run_norm_gate(dmp2_engine* e, int blk, const float* raw, float* x, int L, bool stem, cudaStream_t st, bool have_stats) {
    Workspace& ws = e->ws;
    const Rows rw = rows_of(e, L);
    const int64_t npix = (int64_t)rw.R * L;
    const float* gamma = stem ? e->w.stem_gamma : e->w.blk[blk].gamma;
    const float* beta = stem ? e->w.stem_beta : e->w.blk[blk].beta;
    if (!have_stats) {                                   // (after a tensor-core conv the sums come from its epilogue)
        int sgrid = (int)std::min<int64_t>(cdiv64(npix, 64), (int64_t)e->num_sms * 2);
        k_in_stats<<<sgrid, 512, 0, st>>>(raw, npix, gamma, ws.stat_part, ws.ticket, ws.norm_ss, e->strip_on ? e->sp.totals : nullptr);
        POST_LAUNCH(e, "k_in_stats");
    }
    if (e->strip_on) TRY(strip_stats_exchange(e, gamma, ws.norm_ss, st));      // statistics are over ALL rows of the map
    const ActPtrs act = act_of(e, L);
    int agrid = (int)std::min<int64_t>(cdiv64(npix, 8), (int64_t)e->num_sms * 8);
    k_norm_gate<<<agrid, 256, 0, st>>>(raw, ws.norm_ss, beta, stem ? nullptr : e->w.blk[blk].gate_c,
                                       stem ? nullptr : e->w.blk[blk].sse_w, stem ? 0.f : e->w.blk[blk].sse_b, stem ? 1 : 0,
                                       npix, x, act.xh, act.xl, act.x8lo, act.x8hi, e->conv_mode == DMP2_CONV_TC_F16X3 ? 1 : 0,
                                       e->conv_mode == DMP2_CONV_TC_F16F8 ? 1 : 0);
    POST_LAUNCH(e, "k_norm_gate");
    if (e->strip_on && blk != DMP2_NBLOCKS - 1) TRY(strip_halo_push(e, st));   // the next conv reads 2 rows of each neighbour
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// CUDA-core 5x5 conv (implicit GEMM M=L*L, N=512, K=25*128) with bias + max over 4 consecutive channels.
// Validation path for conv_tc.cu; same math as network.py:26 + :30-31.
// ---------------------------------------------------------------------------------------------------
struct ConvLoad {
    static constexpr bool n_major = false;
    const float* x; int L;
    __device__ float4 operator()(int m, int k) const {
        int tap = k >> 7, c = k & 127;
        int ky = tap / 5, kx = tap - ky * 5;
        int y = m / L, xq = m - y * L;
        int yy = y + ky - 2, xx = xq + kx - 2;
        if (yy < 0 || yy >= L || xx < 0 || xx >= L) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(x + ((int64_t)yy * L + xx) * 128 + c);
    }
};
struct ConvMaxEpilogue {
    float* raw; const float* bias;
    __device__ void operator()(int m, int n, float4 v) const {
        float4 b = *reinterpret_cast<const float4*>(bias + n);
        raw[(int64_t)m * 128 + (n >> 2)] = fmaxf(fmaxf(v.x + b.x, v.y + b.y), fmaxf(v.z + b.z, v.w + b.w));
    }
};

int run_conv_ffma(dmp2_engine* e, int blk, const float* x, int L, float* raw, cudaStream_t st) {
    sgemm_launch<8>(L * L, 512, 3200, ConvLoad{x, L}, LoadRowMajorK{e->w.blk[blk].w_f32, 3200},
                    ConvMaxEpilogue{raw, e->w.blk[blk].bias}, st);
    POST_LAUNCH(e, "sgemm<conv5>");
    return 0;
}

// One ResNet block in place on ws.x (and its fp16 split ws.xh/xl).
int run_resblock(dmp2_engine* e, int blk, int L, cudaStream_t st) {
    Workspace& ws = e->ws;
    // captured into the recycling graph: the graph owns one event pair per block (re-recorded by every replay)
    const bool gprof = e->capturing && e->profile && e->gprof_ev.size() >= (size_t)2 * DMP2_NBLOCKS;
    const bool prof = !e->capturing && e->profile && e->prof_used + 2 <= e->prof_ev.size();
    const bool fused = conv_tc_fuses_stats(e);
    if (prof) cudaEventRecord(e->prof_ev[e->prof_used], st);
    if (gprof) cudaEventRecord(e->gprof_ev[2 * blk], st);
    if (e->strip_on) {
        // strip of R rows; the activation copies carry 2 halo rows on each side, filled by the neighbours
        const Rows rw = rows_of(e, L);
        const StripCtx& sp = e->sp;
        TRY(strip_halo_wait(e, st));
        TRY(run_conv_tc(e, blk, reinterpret_cast<const __half*>(sp.win + sp.off_act[0]), reinterpret_cast<const __half*>(sp.win + sp.off_act[1]),
                        sp.win + sp.off_act[2], sp.win + sp.off_act[3], L, rw.R, 2, rw.R + 4, ws.raw, e->conv_mode, st, fused));
    } else if (e->conv_mode == DMP2_CONV_FFMA) TRY(run_conv_ffma(e, blk, ws.x, L, ws.raw, st));
    else TRY(run_conv_tc(e, blk, ws.xh, ws.xl, ws.x8lo, ws.x8hi, L, L, 0, L, ws.raw, e->conv_mode, st, fused));
    if (prof) { cudaEventRecord(e->prof_ev[e->prof_used + 1], st); e->prof_used += 2; }
    if (gprof) cudaEventRecord(e->gprof_ev[2 * blk + 1], st);
    return run_norm_gate(e, blk, ws.raw, ws.x, L, false, st, fused && e->conv_mode != DMP2_CONV_FFMA);
}

// ---------------------------------------------------------------------------------------------------
// Head: 1x1 conv 128 -> 2 (network.py:207), then conf = row mean of channel 1, dm = |sym(channel 0)|,
// M[i][j] = 0.5 * (dm[0][j]^2 + dm[i][0]^2 - dm[i][j]^2)                               (network.py:237-246)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_head(const float* __restrict__ x, const float* __restrict__ w, float b0, float b1,
                                              int64_t npix, float* __restrict__ head0, float* __restrict__ head1) {
    const int lane = threadIdx.x & 31;
    const float4 w0 = *reinterpret_cast<const float4*>(w + lane * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(w + 128 + lane * 4);
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < npix; p += nwarps) {
        float4 v = *reinterpret_cast<const float4*>(x + p * 128 + lane * 4);
        float d0 = v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
        float d1 = v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
        for (int s = 16; s; s >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, s);
            d1 += __shfl_xor_sync(0xffffffffu, d1, s);
        }
        if (lane == 0) { head0[p] = d0 + b0; head1[p] = d1 + b1; }
    }
}

__device__ __forceinline__ float sym_abs(const float* dm, int L, int i, int j) {
    return fabsf(__fmul_rn(__fadd_rn(dm[(int64_t)i * L + j], dm[(int64_t)j * L + i]), 0.5f));
}

__global__ void __launch_bounds__(256) k_head_post(const float* __restrict__ head, int L, float* __restrict__ conf,
                                                   float* __restrict__ mmat) {
    const int i = blockIdx.x;
    const float* dm = head;
    const float* cf = head + (int64_t)L * L;
    __shared__ double red[256];
    double s = 0;
    const float di0 = sym_abs(dm, L, i, 0);
    const float di0sq = __fmul_rn(di0, di0);
    for (int j = threadIdx.x; j < L; j += 256) {
        s += (double)cf[(int64_t)i * L + j];
        float d0j = sym_abs(dm, L, 0, j);
        float dij = sym_abs(dm, L, i, j);
        float t = __fadd_rn(__fmul_rn(d0j, d0j), di0sq);
        t = __fsub_rn(t, __fmul_rn(dij, dij));
        mmat[(int64_t)i * L + j] = __fmul_rn(0.5f, t);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) conf[i] = (float)(red[0] / (double)L);
}

int run_head(dmp2_engine* e, const float* x, int L, float* head2, cudaStream_t st) {
    // x: this engine's rows of the residual stream; head2: the whole (2, L, L) map (other rows come from the peers)
    const Rows rw = rows_of(e, L);
    const int64_t npix = (int64_t)rw.R * L, first = (int64_t)rw.r0 * L;
    int grid = (int)std::min<int64_t>(cdiv64(npix, 8), (int64_t)e->num_sms * 8);
    k_head<<<grid, 256, 0, st>>>(x, e->w.head_w, e->w.head_b[0], e->w.head_b[1], npix, head2 + first, head2 + (int64_t)L * L + first);
    POST_LAUNCH(e, "k_head");
    if (e->strip_on) TRY(strip_head_gather(e, st));
    return 0;
}
int run_head_post(dmp2_engine* e, const float* head2, int L, float* conf, float* mmat, cudaStream_t st) {
    k_head_post<<<L, 256, 0, st>>>(head2, L, conf, mmat);
    POST_LAUNCH(e, "k_head_post");
    return 0;
}
