// MSA -> pair features: sequence re-weighting (predict.py:32-37) and the shrunk-covariance / inverse /
// APC feature stack of fast_dca (predict.py:41-61).
#include "common.cuh"
#include "sgemm.cuh"

// ---------------------------------------------------------------------------------------------------
// msa [N][L] -> msa_t [L][Npad] with codes clamped to 20 (predict.py:136: gap and unknown share a class);
// pad columns get 255, which never equals a residue code.
// ---------------------------------------------------------------------------------------------------
__global__ void k_msa_transpose(const uint8_t* __restrict__ msa, int N, int L, int Npad, uint8_t* __restrict__ msa_t) {
    __shared__ uint8_t tile[32][33];
    int n0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) {
        int n = n0 + r, l = l0 + tx;
        uint8_t v = 255;
        if (n < N && l < L) { v = msa[(int64_t)n * L + l]; v = v > 20 ? 20 : v; }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int l = l0 + r, n = n0 + tx;
        if (l < L && n < Npad) msa_t[(int64_t)l * Npad + n] = tile[tx][r];
    }
}

// One CTA per sequence n: count identical columns against every m with byte-SIMD compares, then
// w_n = 1 / #{m : count > thr}  (strict '>' against float32(0.8*L), predict.py:33-36).
__global__ void __launch_bounds__(256) k_identity_weights(const uint8_t* __restrict__ msa_t, int N, int L, int Npad,
                                                         float thr, float* __restrict__ w) {
    extern __shared__ uint32_t row_bcast[];           // [L] code of row n replicated into 4 bytes
    __shared__ int red[8];
    const int n = blockIdx.x;
    for (int l = threadIdx.x; l < L; l += blockDim.x) row_bcast[l] = 0x01010101u * msa_t[(int64_t)l * Npad + n];
    __syncthreads();
    const int nwords = Npad >> 2;
    const uint32_t* mt = reinterpret_cast<const uint32_t*>(msa_t);
    int hits = 0;
    for (int wd = threadIdx.x; wd < nwords; wd += blockDim.x) {
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        for (int lb = 0; lb < L; lb += 255) {
            uint32_t packed = 0;
            int le = min(L, lb + 255);
            for (int l = lb; l < le; l++) {
                uint32_t v = mt[(int64_t)l * nwords + wd];
                packed += __vcmpeq4(v, row_bcast[l]) & 0x01010101u;
            }
            c0 += packed & 255; c1 += (packed >> 8) & 255; c2 += (packed >> 16) & 255; c3 += packed >> 24;
        }
        int m = wd * 4;
        hits += (m < N && (float)c0 > thr) + (m + 1 < N && (float)c1 > thr) + (m + 2 < N && (float)c2 > thr) +
                (m + 3 < N && (float)c3 > thr);
    }
    for (int o = 16; o; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hits;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        w[n] = 1.0f / (float)t;
    }
}

// scal[0] = sum w, scal[1] = n_eff = sum w - sqrt(mean w), scal[2] = ridge = 4.5 / sqrt(sum w)   (predict.py:45,51)
__global__ void __launch_bounds__(1024) k_weight_scalars(const float* __restrict__ w, int N, float* __restrict__ scal) {
    __shared__ double red[1024];
    double s = 0;
    for (int i = threadIdx.x; i < N; i += 1024) s += (double)w[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float sum = (float)red[0];
        float mean = sum / (float)N;
        scal[0] = sum;
        scal[1] = sum - sqrtf(mean);
        scal[2] = 4.5f / sqrtf(sum);
    }
}

// Weighted column frequencies: mean[l*21+a] = sum_{n: s_nl = a} w_n / n_eff       (predict.py:47)
__global__ void k_col_mean(const uint8_t* __restrict__ msa_t, const float* __restrict__ w, int N, int Npad,
                           const float* __restrict__ scal, float* __restrict__ mean) {
    int l = blockIdx.x, a = threadIdx.x;
    if (a >= 21) return;
    const uint8_t* col = msa_t + (int64_t)l * Npad;
    float acc = 0.f;
    for (int n = 0; n < N; n++) acc += (col[n] == a) ? w[n] : 0.f;
    mean[l * 21 + a] = acc / scal[1];
}

// xc[(l*21+a)][n] = (onehot - mean) * sqrt(w_n), zero in the pad columns              (predict.py:48)
// optionally also the transposed copy xct[n][l*21+a] (row stride n4) for the Woodbury path
__global__ void k_center(const uint8_t* __restrict__ msa_t, const float* __restrict__ w, const float* __restrict__ mean,
                         int N, int Npad, float* __restrict__ xc, float* __restrict__ xct, int n4) {
    int l = blockIdx.y;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    uint8_t c = msa_t[(int64_t)l * Npad + n];
    float sw = n < N ? sqrtf(w[n]) : 0.f;
#pragma unroll
    for (int a = 0; a < 21; a++) {
        float v = ((c == a ? 1.f : 0.f) - mean[l * 21 + a]) * sw;
        v = n < N ? v : 0.f;
        xc[((int64_t)(l * 21 + a)) * Npad + n] = v;
        if (xct) xct[(int64_t)n * n4 + l * 21 + a] = v;
    }
}

// Woodbury epilogues:  K = X^T X + ridge*n_eff*I ;  inv = (I - X Y) / ridge
struct GramEpilogue {
    float* c; int64_t ld; const float* scal;
    __device__ void operator()(int m, int n, float4 v) const {
        float d = scal[2] * scal[1];
        if (m == n) v.x += d;
        if (m == n + 1) v.y += d;
        if (m == n + 2) v.z += d;
        if (m == n + 3) v.w += d;
        *reinterpret_cast<float4*>(c + (int64_t)m * ld + n) = v;
    }
};
struct WoodburyEpilogue {
    float* c; int64_t ld; const float* scal;
    int m0;                              // first row of the inverse this launch computes (0 unless halo-sharded)
    __device__ void operator()(int m, int n, float4 v) const {
        m += m0;
        float r = 1.0f / scal[2];
        float4 o = make_float4(-v.x * r, -v.y * r, -v.z * r, -v.w * r);
        if (m == n) o.x += r;
        if (m == n + 1) o.y += r;
        if (m == n + 2) o.z += r;
        if (m == n + 3) o.w += r;
        *reinterpret_cast<float4*>(c + (int64_t)m * ld + n) = o;
    }
};

// cov = xc xc^T / n_eff + ridge * I, written into the npad x npad Gauss-Jordan buffer (predict.py:50-51)
struct CovEpilogue {
    float* c; int64_t ld; const float* scal;
    __device__ void operator()(int m, int n, float4 v) const {
        float inv = 1.0f / scal[1], ridge = scal[2];
        float4 o = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
        if (m == n) o.x += ridge;
        if (m == n + 1) o.y += ridge;
        if (m == n + 2) o.z += ridge;
        if (m == n + 3) o.w += ridge;
        *reinterpret_cast<float4*>(c + (int64_t)m * ld + n) = o;
    }
};

// identity in the padding rows/cols so that the padded matrix stays invertible
__global__ void k_pad_identity(float* __restrict__ a, int n, int npad) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)npad * npad;
    if (idx >= total) return;
    int r = (int)(idx / npad), c = (int)(idx % npad);
    if (r >= n || c >= n) a[idx] = (r == c) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// In-place blocked Gauss-Jordan inversion (SPD input, no pivoting; cond(cov_reg) ~ 20-40).
// Per 64-wide pivot block k:  P = A_kk^-1;  R = P A_k,:;  A_ij -= A_ik R_j (i,j != k);
//                             A_ik = -A_ik P;  A_kj = R_j;  A_kk = P.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gj_pivot(const float* __restrict__ a, int64_t ld, int k0, float* __restrict__ p) {
    __shared__ double s[64][65];
    for (int i = threadIdx.x; i < 4096; i += 256) s[i >> 6][i & 63] = (double)a[(int64_t)(k0 + (i >> 6)) * ld + k0 + (i & 63)];
    __syncthreads();
    for (int c = 0; c < 64; c++) {
        __shared__ double colc[64];
        __shared__ double pinv;
        if (threadIdx.x < 64) colc[threadIdx.x] = s[threadIdx.x][c];
        if (threadIdx.x == 0) pinv = 1.0 / s[c][c];
        __syncthreads();
        // row c scaled (with the implicit identity column): a[c][c] := 1 first
        if (threadIdx.x < 64) {
            double v = (threadIdx.x == (unsigned)c) ? 1.0 : s[c][threadIdx.x];
            s[c][threadIdx.x] = v * pinv;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 4096; i += 256) {
            int r = i >> 6, cc = i & 63;
            if (r == c) continue;
            double f = colc[r];
            double base = (cc == c) ? 0.0 : s[r][cc];
            s[r][cc] = base - f * s[c][cc];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 4096; i += 256) p[i] = (float)s[i >> 6][i & 63];
}

struct GjUpdateEpilogue {
    float* a; int64_t ld; int k0;
    __device__ void operator()(int m, int n, float4 v) const {
        if ((m >= k0 && m < k0 + 64) || (n >= k0 && n < k0 + 64)) return;   // n%4==0 and k0%64==0: whole quad in/out
        float4* d = reinterpret_cast<float4*>(a + (int64_t)m * ld + n);
        float4 o = *d;
        o.x -= v.x; o.y -= v.y; o.z -= v.z; o.w -= v.w;
        *d = o;
    }
};

// column panel A_ik = -A_ik P (one warp per row i), row panel A_kj = R_j, A_kk = P
__global__ void __launch_bounds__(256) k_gj_panels(float* __restrict__ a, int64_t ld, int npad, int k0,
                                                   const float* __restrict__ p, const float* __restrict__ r) {
    __shared__ float ps[64][64];
    for (int i = threadIdx.x; i < 4096; i += 256) ps[i >> 6][i & 63] = p[i];
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * 8 + warp;
    if (row >= npad) return;
    if (row >= k0 && row < k0 + 64) {
        // pivot row block: copy R (and P on the diagonal block)
        int kk = row - k0;
        for (int j = lane; j < npad; j += 32) {
            float v = (j >= k0 && j < k0 + 64) ? ps[kk][j - k0] : r[(int64_t)kk * npad + j];
            a[(int64_t)row * ld + j] = v;
        }
    } else {
        float* seg = a + (int64_t)row * ld + k0;
        float s0 = seg[lane], s1 = seg[lane + 32];
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
        for (int m = 0; m < 64; m++) {
            float sm = __shfl_sync(0xffffffffu, m < 32 ? s0 : s1, m & 31);
            o0 = fmaf(sm, ps[m][lane], o0);
            o1 = fmaf(sm, ps[m][lane + 32], o1);
        }
        seg[lane] = -o0;
        seg[lane + 32] = -o1;
    }
}

// ---------------------------------------------------------------------------------------------------
// inverse covariance -> features                                                        (predict.py:54-61)
// ---------------------------------------------------------------------------------------------------
// feat[(i*L+j)][a*21+b] = inv[i*21+a][j*21+b];  x3[i][j] = ||inv[i,:20,j,:20]||_F (0 on the diagonal)
__global__ void __launch_bounds__(128) k_feat_gather(const float* __restrict__ inv, int64_t ld, int L, int i0, float* __restrict__ feat,
                                                     float* __restrict__ x3) {
    int i = blockIdx.y + i0, j = blockIdx.x;
    __shared__ float red[4];
    float ss = 0.f;
    float* out = feat + ((int64_t)i * L + j) * DMP2_FEAT_LD;
    for (int t = threadIdx.x; t < DMP2_FEAT_LD; t += 128) {
        float v = 0.f;
        if (t < 441) {
            int a = t / 21, b = t - a * 21;
            v = inv[(int64_t)(i * 21 + a) * ld + j * 21 + b];
            if (a < 20 && b < 20) ss += v * v;
        }
        out[t] = v;
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) x3[i * L + j] = (i == j) ? 0.f : sqrtf(red[0] + red[1] + red[2] + red[3]);
}

// apc[0..L) = column sums (sum over i), apc[L..2L) = row sums (sum over j), apc[2L] = total
__global__ void __launch_bounds__(256) k_apc_sums(const float* __restrict__ x3, int L, float* __restrict__ apc) {
    int t = blockIdx.x;          // 0..L-1: column t; L..2L-1: row t-L; 2L: total (from the row sums, after them)
    __shared__ double red[256];
    double s = 0;
    if (t < L) for (int i = threadIdx.x; i < L; i += 256) s += (double)x3[(int64_t)i * L + t];
    else for (int j = threadIdx.x; j < L; j += 256) s += (double)x3[(int64_t)(t - L) * L + j];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) apc[t] = (float)red[0];
}
__global__ void __launch_bounds__(1024) k_apc_total(int L, float* __restrict__ apc) {
    __shared__ double red[1024];
    double s = 0;
    for (int i = threadIdx.x; i < L; i += 1024) s += (double)apc[L + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) apc[2 * L] = (float)red[0];
}
__global__ void k_apc_apply(const float* __restrict__ x3, const float* __restrict__ apc, int L, int first, int count,
                            float* __restrict__ feat) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    idx += first;
    int i = idx / L, j = idx - i * L;
    float v = (i == j) ? 0.f : x3[idx] - apc[j] * apc[L + i] / apc[2 * L];
    feat[(int64_t)idx * DMP2_FEAT_LD + 441] = v;
}

// (L,L,442) reference layout <-> padded 444 layout (stage entry points only)
__global__ void k_feat_repack(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld, int64_t npix) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * dst_ld) return;
    int64_t p = idx / dst_ld;
    int c = (int)(idx - p * dst_ld);
    dst[idx] = c < src_ld && c < 442 ? src[p * src_ld + c] : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// in-place inverse of the SPD matrix a (npad x npad, npad % 64 == 0)
static int gj_invert(dmp2_engine* e, float* a, int npad, cudaStream_t st) {
    Workspace& ws = e->ws;
    for (int k0 = 0; k0 < npad; k0 += 64) {
        k_gj_pivot<<<1, 256, 0, st>>>(a, npad, k0, ws.gj_p);
        POST_LAUNCH(e, "k_gj_pivot");
        sgemm_launch<4>(64, npad, 64, LoadRowMajorK{ws.gj_p, 64}, LoadColMajorN{a + (int64_t)k0 * npad, npad},
                        StoreRowMajor{ws.gj_r, npad, nullptr, 1.0f}, st);
        POST_LAUNCH(e, "sgemm<gj_row>");
        if (npad >= 2048)
            sgemm_launch<8>(npad, npad, 64, LoadRowMajorK{a + k0, npad}, LoadColMajorN{ws.gj_r, npad}, GjUpdateEpilogue{a, npad, k0}, st);
        else
            sgemm_launch<4>(npad, npad, 64, LoadRowMajorK{a + k0, npad}, LoadColMajorN{ws.gj_r, npad}, GjUpdateEpilogue{a, npad, k0}, st);
        POST_LAUNCH(e, "sgemm<gj_update>");
        k_gj_panels<<<cdiv(npad, 8), 256, 0, st>>>(a, npad, npad, k0, ws.gj_p, ws.gj_r);
        POST_LAUNCH(e, "k_gj_panels");
    }
    return 0;
}

static int prep_msa(dmp2_engine* e, const uint8_t* msa, int N, int L, cudaStream_t st) {
    int Npad = (N + 3) & ~3;
    dim3 g(cdiv(Npad, 32), cdiv(L, 32));
    k_msa_transpose<<<g, dim3(32, 8), 0, st>>>(msa, N, L, Npad, e->ws.msa_t);
    POST_LAUNCH(e, "k_msa_transpose");
    return 0;
}

int run_reweight(dmp2_engine* e, const uint8_t* msa, int N, int L, float* w_out, cudaStream_t st) {
    TRY(prep_msa(e, msa, N, L, st));
    int Npad = (N + 3) & ~3;
    float thr = (float)((double)L * 0.8);
    k_identity_weights<<<N, 256, L * sizeof(uint32_t), st>>>(e->ws.msa_t, N, L, Npad, thr, w_out);
    POST_LAUNCH(e, "k_identity_weights");
    return 0;
}

// Requires ws.msa_t from run_reweight(e, msa, ...) on the same stream.
int run_dca(dmp2_engine* e, const uint8_t* msa, int N, int L, const float* w, float* feat444, cudaStream_t st) {
    (void)msa;
    Workspace& ws = e->ws;
    const int64_t npix = (int64_t)L * L;
    if (N <= 1) {                       // predict.py:139 -- a single sequence gets all-zero features
        CUDA_TRY(e, cudaMemsetAsync(feat444, 0, npix * DMP2_FEAT_LD * sizeof(float), st));
        return 0;
    }
    const int Npad = (N + 3) & ~3;
    const int n = 21 * L, npad = (n + 63) & ~63;
    float* mean = ws.gj_r;              // [64][npad] >= 21L floats, free until the Gauss-Jordan loop
    k_weight_scalars<<<1, 1024, 0, st>>>(w, N, ws.scal);
    POST_LAUNCH(e, "k_weight_scalars");
    k_col_mean<<<L, 32, 0, st>>>(ws.msa_t, w, N, Npad, ws.scal, mean);
    POST_LAUNCH(e, "k_col_mean");
    const int Npad64 = (N + 63) & ~63, n4 = (n + 3) & ~3;
    const bool woodbury = Npad64 < npad;       // N < 21 L: invert the N x N Gram system instead (Woodbury identity)
    // halo-sharded fold: the last (and largest) product of the Woodbury path, the feature gather and the APC term are
    // restricted to this rank's rows; the contact-norm map is completed through the window (strip.cu)
    const bool shard = e->strip_on && e->sp.world > 1 && woodbury;
    const Rows rw = shard ? rows_of(e, L) : Rows{0, L};
    float* x3 = shard ? reinterpret_cast<float*>(e->sp.win + e->sp.off_x3) : ws.x3;
    if (woodbury && n4 != n) CUDA_TRY(e, cudaMemsetAsync(ws.xct, 0, (size_t)Npad * n4 * sizeof(float), st));   // pad columns
    k_center<<<dim3(cdiv(Npad, 256), L), 256, 0, st>>>(ws.msa_t, w, mean, N, Npad, ws.xc, woodbury ? ws.xct : nullptr, n4);
    POST_LAUNCH(e, "k_center");
    // ---- tensor-core form of the dense contractions (default; DMP2_GEMM=ffma keeps them on the CUDA-core GEMM) -------
    // Every fp32 operand is split into an fp16 hi/lo pair after a power-of-two scaling that centres it in the fp16 range
    // (scale found on the device: run_operand_scale); products are exact, 3 MMAs per MAC, accumulation chains of 128
    // summed in fp32 registers (conv_tc.cu) -- at least as accurate as the sequential fp32 sums of the CUDA-core path.
    const int Kp2 = (Npad + 127) & ~127;                   // contraction over the sequence axis
    const int Kp1 = (n4 + 127) & ~127;                     // contraction over the 21 L axis (Gram matrix)
    __half* t = ws.dca_tc;
    __half *x_hi = t, *x_lo = x_hi + (int64_t)n * Kp2;                                            // xc   [n][Kp2]
    __half *w_hi = x_lo + (int64_t)n * Kp2, *w_lo = w_hi + (int64_t)n * Kp2;                      // wy^T [n][Kp2]
    __half *k_hi = w_lo + (int64_t)n * Kp2, *k_lo = k_hi + (int64_t)Npad64 * Kp2;                 // K^-1 [Npad64][Kp2]
    __half *xt_hi = k_lo + (int64_t)Npad64 * Kp2, *xt_lo = xt_hi + (int64_t)Npad * Kp1;           // xct  [Npad][Kp1]
    float *ds_x = ws.tc_scal, *ds_k = ws.tc_scal + 4, *ds_w = ws.tc_scal + 8;
    if (e->gemm_tc) {
        TRY(run_operand_scale(e, ws.xc, n, Npad, Npad, ds_x, st));
        TRY(run_split_scaled(e, ws.xc, n, Npad, Npad, 1.0f, ds_x, x_hi, x_lo, Kp2, st));
    }
    if (!woodbury && e->gemm_tc) {
        const GemmTcEpilogue ep{3, 0, ds_x, ds_x, ws.scal};
        TRY(run_gemm_tc(e, x_hi, x_lo, x_hi, x_lo, n, n4, Kp2, 1.0f, ws.cov, npad, 128, st, &ep, n));      // covariance (predict.py:50-51)
        if (npad != n) {
            k_pad_identity<<<(unsigned)cdiv64((int64_t)npad * npad, 256), 256, 0, st>>>(ws.cov, n, npad);
            POST_LAUNCH(e, "k_pad_identity");
        }
        TRY(gj_invert(e, ws.cov, npad, st));
    } else if (woodbury && e->gemm_tc) {
        TRY(run_split_scaled(e, ws.xct, Npad, n4, n4, 1.0f, ds_x, xt_hi, xt_lo, Kp1, st));
        const GemmTcEpilogue epg{1, 0, ds_x, ds_x, ws.scal};
        TRY(run_gemm_tc(e, xt_hi, xt_lo, xt_hi, xt_lo, Npad, Npad, Kp1, 1.0f, ws.kmat, Npad64, 128, st, &epg));   // Gram system
        k_pad_identity<<<(unsigned)cdiv64((int64_t)Npad64 * Npad64, 256), 256, 0, st>>>(ws.kmat, N, Npad64);
        POST_LAUNCH(e, "k_pad_identity");
        TRY(gj_invert(e, ws.kmat, Npad64, st));
        // wy^T = X K^-1 (K^-1 is symmetric, so its rows serve as the K-major B operand), then inv = (I - X wy) / ridge
        TRY(run_operand_scale(e, ws.kmat, Npad, Npad, Npad64, ds_k, st));
        TRY(run_split_scaled(e, ws.kmat, Npad, Npad, Npad64, 1.0f, ds_k, k_hi, k_lo, Kp2, st));
        const GemmTcEpilogue epy{0, 0, ds_x, ds_k, ws.scal};
        float* wyt = ws.wy;                                // [n][Npad] fp32
        TRY(run_gemm_tc(e, x_hi, x_lo, k_hi, k_lo, n, Npad, Kp2, 1.0f, wyt, Npad, 128, st, &epy));
        TRY(run_operand_scale(e, wyt, n, Npad, Npad, ds_w, st));
        TRY(run_split_scaled(e, wyt, n, Npad, Npad, 1.0f, ds_w, w_hi, w_lo, Kp2, st));
        const GemmTcEpilogue epw{2, 21 * rw.r0, ds_x, ds_w, ws.scal};
        // halo-sharded: only the rows of the inverse that feed this rank's strip of the pair features
        TRY(run_gemm_tc(e, x_hi + (int64_t)21 * rw.r0 * Kp2, x_lo + (int64_t)21 * rw.r0 * Kp2, w_hi, w_lo, 21 * rw.R, n4, Kp2, 1.0f,
                        ws.cov + (int64_t)21 * rw.r0 * npad, npad, 128, st, &epw, n));
    } else if (!woodbury) {
        sgemm_launch<8>(n, n, Npad, LoadRowMajorK{ws.xc, Npad}, LoadRowMajorK{ws.xc, Npad}, CovEpilogue{ws.cov, npad, ws.scal}, st);
        POST_LAUNCH(e, "sgemm<cov>");
        if (npad != n) {
            k_pad_identity<<<(unsigned)cdiv64((int64_t)npad * npad, 256), 256, 0, st>>>(ws.cov, n, npad);
            POST_LAUNCH(e, "k_pad_identity");
        }
        TRY(gj_invert(e, ws.cov, npad, st));
    } else {
        // cov_reg = ridge*I + X X^T / n_eff  with X = xc (21L x N)
        // inv     = (I - X (ridge*n_eff*I_N + X^T X)^-1 X^T) / ridge
        sgemm_launch<4>(Npad, Npad, n4, LoadRowMajorK{ws.xct, n4}, LoadRowMajorK{ws.xct, n4}, GramEpilogue{ws.kmat, Npad64, ws.scal}, st);
        POST_LAUNCH(e, "sgemm<gram>");
        k_pad_identity<<<(unsigned)cdiv64((int64_t)Npad64 * Npad64, 256), 256, 0, st>>>(ws.kmat, N, Npad64);
        POST_LAUNCH(e, "k_pad_identity");
        TRY(gj_invert(e, ws.kmat, Npad64, st));
        sgemm_launch<8>(Npad, n4, Npad, LoadRowMajorK{ws.kmat, Npad64}, LoadColMajorN{ws.xct, n4}, StoreRowMajor{ws.wy, n4, nullptr, 1.0f}, st);
        POST_LAUNCH(e, "sgemm<woodbury_y>");
        // halo-sharded: only the rows of the inverse that feed this rank's strip of the pair features
        sgemm_launch<8>(21 * rw.R, n4, Npad, LoadRowMajorK{ws.xc + (int64_t)21 * rw.r0 * Npad, Npad}, LoadColMajorN{ws.wy, n4},
                        WoodburyEpilogue{ws.cov, npad, ws.scal, 21 * rw.r0}, st);
        POST_LAUNCH(e, "sgemm<woodbury_inv>");
    }
    k_feat_gather<<<dim3(L, rw.R), 128, 0, st>>>(ws.cov, npad, L, rw.r0, feat444, x3);
    POST_LAUNCH(e, "k_feat_gather");
    if (shard) TRY(strip_x3_gather(e, st));             // APC needs the sums over ALL rows
    k_apc_sums<<<2 * L, 256, 0, st>>>(x3, L, ws.apc);
    POST_LAUNCH(e, "k_apc_sums");
    k_apc_total<<<1, 1024, 0, st>>>(L, ws.apc);
    POST_LAUNCH(e, "k_apc_total");
    k_apc_apply<<<cdiv(rw.R * L, 256), 256, 0, st>>>(x3, ws.apc, L, rw.r0 * L, rw.R * L, feat444);
    POST_LAUNCH(e, "k_apc_apply");
    return 0;
}

int run_feat_export(dmp2_engine* e, const float* feat444, int L, float* feat442, cudaStream_t st) {
    int64_t npix = (int64_t)L * L;
    k_feat_repack<<<(unsigned)cdiv64(npix * 442, 256), 256, 0, st>>>(feat444, DMP2_FEAT_LD, feat442, 442, npix);
    POST_LAUNCH(e, "k_feat_repack");
    return 0;
}
int run_feat_import(dmp2_engine* e, const float* feat442, int L, float* feat444, cudaStream_t st) {
    int64_t npix = (int64_t)L * L;
    k_feat_repack<<<(unsigned)cdiv64(npix * DMP2_FEAT_LD, 256), 256, 0, st>>>(feat442, 442, feat444, DMP2_FEAT_LD, npix);
    POST_LAUNCH(e, "k_feat_repack");
    return 0;
}
