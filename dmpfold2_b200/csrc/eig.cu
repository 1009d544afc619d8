// Top-8 eigenpairs of the symmetric L x L Gram-like matrix M (network.py:247-250), replacing torch.symeig.
// Direct method, fp64 internally (the spectrum is hostile to iterative schemes: strongly indefinite,
// lambda_9/lambda_8 up to 0.996 -- SURVEY.md section 7.2):
//   1. Householder tridiagonalisation          2. Sturm multi-section for the 8 largest eigenvalues
//   3. inverse iteration + Gram-Schmidt        4. back-transformation
//   5. canonical sign (largest-|component| positive, lowest index wins ties), sqrt(clamp(relu(l),1e-8)) scaling.
// v0: one CTA, matrix resident in L2.
#include "common.cuh"

#define EIG_THREADS 1024

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// all threads get the block-wide sum; `red` is a 33-double shared scratch
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (EIG_THREADS >> 5) ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// number of eigenvalues of the tridiagonal (d, e2 = e^2) that are < x
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int n, double x, double tiny) {
    double q = d[0] - x;
    int cnt = q < 0.0;
    for (int i = 1; i < n; i++) {
        if (fabs(q) < tiny) q = q < 0.0 ? -tiny : tiny;
        q = d[i] - x - e2[i - 1] / q;
        cnt += q < 0.0;
    }
    return cnt;
}

__global__ void __launch_bounds__(EIG_THREADS, 1)
k_eig_top8(const float* __restrict__ M, int n, double* __restrict__ A, double* __restrict__ V, double* __restrict__ wk,
           float* __restrict__ vals_out, float* __restrict__ mds_out, float* __restrict__ vec_out) {
    extern __shared__ double sm[];
    double* sv = sm;                 // [n] householder vector
    double* sp = sm + n;             // [n] p, then w
    double* sd = sm + 2 * n;         // [n] diagonal of T
    double* se = sm + 3 * n;         // [n] off-diagonal of T
    double* se2 = sm + 4 * n;        // [n] e^2
    __shared__ double red[33];
    __shared__ double lam[8];
    __shared__ double dots[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EIG_THREADS / 32;
    // workspace in global: beta[n], then per-vector arrays
    double* beta = wk;               // [n]
    double* zs = wk + n;             // [8][n] eigenvectors of T, then of M
    double* u0 = zs + 8 * n;         // [8][n] LU of T - lambda I (three diagonals + multipliers)
    double* u1 = u0 + 8 * n;
    double* u2 = u1 + 8 * n;
    double* mu = u2 + 8 * n;
    // swap flags share mu's sign-free storage: keep a separate array
    double* sw = mu + 8 * n;         // [8][n] 0/1

    for (int64_t i = tid; i < (int64_t)n * n; i += EIG_THREADS) A[i] = (double)M[i];
    __syncthreads();

    // ---------------- 1. Householder tridiagonalisation (full symmetric storage kept up to date) --------
    for (int k = 0; k < n - 2; k++) {
        const int m = n - k - 1;                     // length of the column below the diagonal
        double part = 0.0;
        for (int i = tid; i < m; i += EIG_THREADS) {
            double x = A[(int64_t)(k + 1 + i) * n + k];
            sv[i] = x;
            part += x * x;
        }
        double sigma = block_sum(part, red);
        double x0 = sv[0];
        double tail = sigma - x0 * x0;
        if (tid == 0) sd[k] = A[(int64_t)k * n + k];
        if (!(tail > 0.0)) {                        // column already tridiagonal
            if (tid == 0) { se[k] = x0; beta[k] = 0.0; }
            for (int i = tid; i < m; i += EIG_THREADS) V[(int64_t)k * n + i] = 0.0;
            __syncthreads();
            continue;
        }
        double alpha = (x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma);
        double v0 = x0 - alpha;
        double bt = 2.0 / (tail + v0 * v0);
        __syncthreads();
        if (tid == 0) { sv[0] = v0; se[k] = alpha; beta[k] = bt; }
        __syncthreads();
        for (int i = tid; i < m; i += EIG_THREADS) V[(int64_t)k * n + i] = sv[i];
        // p = beta * A22 v   (warp per row, lanes along the row)
        for (int i = warp; i < m; i += NW) {
            const double* row = A + (int64_t)(k + 1 + i) * n + (k + 1);
            double acc = 0.0;
            for (int j = lane; j < m; j += 32) acc += row[j] * sv[j];
            acc = warp_sum(acc);
            if (lane == 0) sp[i] = bt * acc;
        }
        __syncthreads();
        double pv = 0.0;
        for (int i = tid; i < m; i += EIG_THREADS) pv += sp[i] * sv[i];
        double kk = 0.5 * bt * block_sum(pv, red);
        for (int i = tid; i < m; i += EIG_THREADS) sp[i] -= kk * sv[i];      // w
        __syncthreads();
        for (int i = warp; i < m; i += NW) {
            double* row = A + (int64_t)(k + 1 + i) * n + (k + 1);
            const double vi = sv[i], wi = sp[i];
            for (int j = lane; j < m; j += 32) row[j] -= vi * sp[j] + wi * sv[j];
        }
        __syncthreads();
    }
    if (tid == 0) {
        sd[n - 2] = A[(int64_t)(n - 2) * n + (n - 2)];
        sd[n - 1] = A[(int64_t)(n - 1) * n + (n - 1)];
        se[n - 2] = A[(int64_t)(n - 1) * n + (n - 2)];
        se[n - 1] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += EIG_THREADS) se2[i] = se[i] * se[i];
    // Gershgorin bounds
    double lo_p = 1e300, hi_p = -1e300;
    for (int i = tid; i < n; i += EIG_THREADS) {
        double r = (i > 0 ? fabs(se[i - 1]) : 0.0) + (i < n - 1 ? fabs(se[i]) : 0.0);
        lo_p = fmin(lo_p, sd[i] - r);
        hi_p = fmax(hi_p, sd[i] + r);
    }
    __shared__ double glo[EIG_THREADS / 32], ghi[EIG_THREADS / 32];
    for (int o = 16; o; o >>= 1) {
        lo_p = fmin(lo_p, __shfl_xor_sync(0xffffffffu, lo_p, o));
        hi_p = fmax(hi_p, __shfl_xor_sync(0xffffffffu, hi_p, o));
    }
    if (lane == 0) { glo[warp] = lo_p; ghi[warp] = hi_p; }
    __syncthreads();
    double gl = glo[0], gh = ghi[0];
    for (int i = 1; i < NW; i++) { gl = fmin(gl, glo[i]); gh = fmax(gh, ghi[i]); }
    const double tnorm = fmax(fabs(gl), fabs(gh));
    const double tiny = fmax(tnorm * 1e-20, 1e-300);
    gl -= tnorm * 1e-12 + 1e-300;
    gh += tnorm * 1e-12 + 1e-300;

    // ---------------- 2. Sturm multi-section: warp w finds eigenvalue index n-8+w (ascending) -----------
    if (warp < 8) {
        const int kidx = n - 8 + warp;
        double lo = gl, hi = gh;
        for (int round = 0; round < 16; round++) {
            double step = (hi - lo) / 33.0;
            double x = lo + step * (double)(lane + 1);
            int c = sturm_count(sd, se2, n, x, tiny);
            unsigned bal = __ballot_sync(0xffffffffu, c <= kidx);
            int npre = __popc(bal);
            double x_lo = __shfl_sync(0xffffffffu, x, npre > 0 ? npre - 1 : 0);
            double x_hi = __shfl_sync(0xffffffffu, x, npre < 32 ? npre : 31);
            double nlo = npre > 0 ? x_lo : lo;
            double nhi = npre < 32 ? x_hi : hi;
            bool done = !(nhi - nlo < hi - lo) || (nhi - nlo) <= 4.0 * 2.3e-16 * fmax(fabs(nlo), fabs(nhi));
            lo = nlo; hi = nhi;
            if (done) break;
        }
        if (lane == 0) lam[warp] = 0.5 * (lo + hi);
    }
    __syncthreads();

    // ---------------- 3. inverse iteration on T ---------------------------------------------------------
    for (int i = tid; i < 8 * n; i += EIG_THREADS) {
        int w = i / n, r = i - w * n;
        unsigned h = (unsigned)(r * 2654435761u) ^ (unsigned)((w + 1) * 40503u);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        zs[i] = 0.5 + (double)(h & 0xffff) / 65536.0;
    }
    __syncthreads();
    const double pivmin = fmax(tnorm * 2.3e-16, 1e-290);
    for (int iter = 0; iter < 3; iter++) {
        if (warp < 8 && lane == 0) {
            const int w = warp;
            const double l = lam[w];
            double* z = zs + w * n;
            double *U0 = u0 + w * n, *U1 = u1 + w * n, *U2 = u2 + w * n, *MU = mu + w * n, *SW = sw + w * n;
            // factor + forward substitution in one sweep (factor recomputed every iteration: cheap)
            double c0 = sd[0] - l, c1 = n > 1 ? se[0] : 0.0;          // current row i: (diag, super)
            double rhs = z[0];
            for (int i = 0; i < n - 1; i++) {
                double sub = se[i];                                     // T[i+1][i]
                double an = sd[i + 1] - l;                              // T[i+1][i+1]
                double bn = (i + 2 < n) ? se[i + 1] : 0.0;              // T[i+1][i+2]
                double rn = z[i + 1];
                if (fabs(sub) <= fabs(c0)) {
                    if (fabs(c0) < pivmin) c0 = c0 < 0 ? -pivmin : pivmin;
                    double mlt = sub / c0;
                    U0[i] = c0; U1[i] = c1; U2[i] = 0.0; MU[i] = mlt; SW[i] = 0.0;
                    z[i] = rhs;
                    c0 = an - mlt * c1; c1 = bn; rhs = rn - mlt * rhs;
                } else {
                    double mlt = c0 / sub;
                    U0[i] = sub; U1[i] = an; U2[i] = bn; MU[i] = mlt; SW[i] = 1.0;
                    z[i] = rn;
                    c0 = c1 - mlt * an; c1 = -mlt * bn; rhs = rhs - mlt * rn;
                }
            }
            if (fabs(c0) < pivmin) c0 = c0 < 0 ? -pivmin : pivmin;
            U0[n - 1] = c0; U1[n - 1] = 0.0; U2[n - 1] = 0.0;
            z[n - 1] = rhs;
            // back substitution (the two previous solutions stay in registers)
            double zmax = 0.0, z1 = 0.0, z2 = 0.0;
#pragma unroll 4
            for (int i = n - 1; i >= 0; i--) {
                double t = (z[i] - U1[i] * z1 - U2[i] * z2) / U0[i];
                z[i] = t;
                z2 = z1; z1 = t;
                zmax = fmax(zmax, fabs(t));
            }
            // scale to avoid overflow in the dot products
            double sc = zmax > 0 ? 1.0 / zmax : 1.0;
            for (int i = 0; i < n; i++) z[i] *= sc;
        }
        __syncthreads();
        // modified Gram-Schmidt in ascending order + normalisation
        for (int w = 0; w < 8; w++) {
            double* z = zs + w * n;
            for (int p = 0; p < w; p++) {
                const double* zp = zs + p * n;
                double a = 0.0;
                for (int i = tid; i < n; i += EIG_THREADS) a += zp[i] * z[i];
                double dt = block_sum(a, red);
                for (int i = tid; i < n; i += EIG_THREADS) z[i] -= dt * zp[i];
                __syncthreads();
            }
            double a = 0.0;
            for (int i = tid; i < n; i += EIG_THREADS) a += z[i] * z[i];
            double nr = block_sum(a, red);
            double inv = nr > 0 ? 1.0 / sqrt(nr) : 0.0;
            for (int i = tid; i < n; i += EIG_THREADS) z[i] *= inv;
            __syncthreads();
        }
    }
    (void)dots;

    // ---------------- 4. back-transform: z <- H_0 H_1 ... H_{n-3} z  (warp w owns vector w) --------------
    if (warp < 8) {
        double* z = zs + warp * n;
        for (int k = n - 3; k >= 0; k--) {
            const double bt = beta[k];
            if (bt == 0.0) continue;
            const int m = n - k - 1;
            const double* v = V + (int64_t)k * n;
            double* zz = z + k + 1;
            double a = 0.0;
            for (int i = lane; i < m; i += 32) a += v[i] * zz[i];
            a = warp_sum(a) * bt;
            for (int i = lane; i < m; i += 32) zz[i] -= a * v[i];
            __syncwarp();
        }
        // ---------------- 5. canonical sign + MDS scaling -------------------------------------------------
        double best = -1.0; int bi = 0;
        for (int i = lane; i < n; i += 32) {
            double a = fabs((double)(float)z[i]);
            if (a > best) { best = a; bi = i; }
        }
        for (int o = 16; o; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const double sgn = z[bi] < 0.0 ? -1.0 : 1.0;
        const float lf = (float)lam[warp];
        const float sc = sqrtf(fmaxf(fmaxf(lf, 0.0f), 1e-8f));
        for (int i = lane; i < n; i += 32) {
            float vf = (float)(z[i] * sgn);
            if (vec_out) vec_out[(int64_t)i * 8 + warp] = vf;
            if (mds_out) mds_out[(int64_t)i * 8 + warp] = vf * sc;
        }
        if (lane == 0 && vals_out) vals_out[warp] = lf;
    }
}

int run_eig_top8(dmp2_engine* e, const float* m, int L, float* vals, float* mds_scaled, float* vecs_raw, cudaStream_t st) {
    static bool attr_set = false;
    size_t smem = (size_t)5 * L * sizeof(double);
    if (!attr_set) {
        CUDA_TRY(e, cudaFuncSetAttribute(k_eig_top8, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) return e->fail(DMP2_ERR_UNSUPPORTED, "eig_top8: L too large for the single-CTA solver");
    double* V = e->ws.eig_a + (int64_t)L * L;
    k_eig_top8<<<1, EIG_THREADS, smem, st>>>(m, L, e->ws.eig_a, V, e->ws.eig_w, vals, mds_scaled, vecs_raw);
    POST_LAUNCH(e, "k_eig_top8");
    return 0;
}
