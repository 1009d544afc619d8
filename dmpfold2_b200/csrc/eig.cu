// Top-8 eigenpairs of the symmetric L x L Gram-like matrix M (network.py:247-250), replacing torch.symeig.
// Direct method, fp64 internally (the spectrum is hostile to iterative schemes: strongly indefinite,
// lambda_9/lambda_8 up to 0.996 -- SURVEY.md section 7.2):
//   1. Householder tridiagonalisation, rows distributed cyclically over an 8-CTA cluster (resident in shared
//      memory when they fit); rank-2 update and next matrix-vector product fused into one sweep, ONE all-gather
//      per column through distributed shared memory (st.async + mbarrier, no cluster barrier in the loop);
//      beyond the cluster's capacity the same fused scheme runs on the whole GPU (k_tridiag_grid);
//   2. Sturm multi-section (32 shifts per round) for the 8 largest eigenvalues      } CTA 0
//   3. inverse iteration + Gram-Schmidt                                              }
//   4. back-transformation z <- H_0 ... H_{n-3} z: ONE CTA PER EIGENVECTOR (CTA w of the cluster takes vector w)
//   5. canonical sign (largest-|component| positive, lowest index wins ties), sqrt(clamp(relu(l),1e-8)) scaling.
#include "common.cuh"
#include <stdlib.h>
#include "cluster_comm.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define EIG_THREADS 512

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

// Number of eigenvalues of the tridiagonal (d, e2 = e^2) that are < x = sign changes in the sequence of leading
// principal minors p_-1 = 1, p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (division-free Sturm sequence).  One DFMA on
// the dependent chain per row; signs are read from the high word with integer ops; a zero minor takes the sign
// opposite to its predecessor; both running minors are rescaled by a power of two every 8 rows.
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int n, double x) {
    double pm = 1.0, p = d[0] - x;
    if (p == 0.0) p = -1e-300;
    int sp = (unsigned)__double2hiint(p) >> 31;
    int cnt = sp;
    auto step = [&](double dx, double ee) {
        double pn = fma(dx, p, -(ee * pm));
        const int hi = __double2hiint(pn), lo = __double2loint(pn);
        if (((hi << 1) | lo) == 0) pn = sp ? 1e-300 : -1e-300;           // exact zero: opposite sign of its predecessor
        const int sn = (unsigned)__double2hiint(pn) >> 31;
        cnt += sn ^ sp;
        sp = sn;
        pm = p; p = pn;
    };
    auto rescale = [&]() {
        const int ex = (__double2hiint(p) >> 20) & 0x7ff;
        if (ex > 1023 + 400) { p *= 3.8725919148493183e-121; pm *= 3.8725919148493183e-121; }        // 2^-400
        else if (ex < 1023 - 400) { p *= 2.5822498780869086e120; pm *= 2.5822498780869086e120; }      // 2^400
    };
    int i = 1;
    // blocks of 8 rows: all operands of the block are fetched (and d - x formed) BEFORE the dependent chain starts, so a
    // row costs one DFMA + the zero test instead of a shared-memory round trip per row (which tripled the chain)
    for (; i + 8 <= n; i += 8) {
        double dx[8], ee[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { dx[u] = d[i + u] - x; ee[u] = e2[i + u - 1]; }
#pragma unroll
        for (int u = 0; u < 8; u++) step(dx[u], ee[u]);
        rescale();
    }
    if (i < n) {
        for (; i < n; i++) step(d[i] - x, e2[i - 1]);
        rescale();
    }
    return cnt;
}

// one-barrier block reduction: every thread returns the same total (fixed order); `scr` holds EIG_THREADS/32 doubles
// and must not be reused before another block-wide barrier has been passed
__device__ __forceinline__ double block_sum1(double v, double* scr) {
    static_assert(EIG_THREADS == 512, "16 warp partials: one per lane of a half warp");
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = scr[threadIdx.x & 15];              // lane-parallel pairwise tree: 4 adds per warp instead of a chain of 16
#pragma unroll
    for (int o = 8; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// fixed-order pairwise sum of 16 doubles in shared memory (depth 4 instead of a chain of 16 dependent adds)
__device__ __forceinline__ double tree16(const double* p) {
    const double2* pq = reinterpret_cast<const double2*>(p);
    const double2 q0 = pq[0], q1 = pq[1], q2 = pq[2], q3 = pq[3], q4 = pq[4], q5 = pq[5], q6 = pq[6], q7 = pq[7];
    return (((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y))) +
           (((q4.x + q4.y) + (q5.x + q5.y)) + ((q6.x + q6.y) + (q7.x + q7.y)));
}

// two warp reductions for the price of one and a bit: lanes < 16 return the warp's sum of p, lanes >= 16 that of q
__device__ __forceinline__ double warp_sum_pair(double p, double q, int lane) {
    const bool up = (lane & 16) != 0;
    double keep = up ? q : p;
    keep += __shfl_xor_sync(0xffffffffu, up ? p : q, 16);
#pragma unroll
    for (int o = 8; o; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}

// rows_smem: doubles of shared memory reserved for the matrix rows (0 = rows stay in global memory A)
// vec_smem:  1 = the inverse-iteration work vectors (32 n doubles) live in shared memory
// inverse-iteration sweeps per eigenvector (DMP2_EIG_INVIT overrides; the shifts come out of the Sturm multi-section
// accurate to ~1e-15 relative, so the second sweep already reproduces fp64 LAPACK eigenvectors to 1e-8: a third changes
// nothing in max|dvec| (7.2e-9 at L=300, profiles/round2_eig.txt) and costs 73 us)
__device__ int g_invit_iters = 2;

template <int EIG_CL>
__global__ void __launch_bounds__(EIG_THREADS, 1)
k_eig_top8(const float* __restrict__ M, int n, double* __restrict__ A, double* __restrict__ V, double* __restrict__ wk,
           int rows_in_smem, int vec_in_smem, int pre_tridiag, int stage_rows, float* __restrict__ vals_out,
           float* __restrict__ mds_out, float* __restrict__ vec_out) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    extern __shared__ double sm[];
    // [0, 7n): Householder work vectors during phase 1 (see there); afterwards
    double* sd = sm + 4 * n;         // [n] diagonal of T
    double* se = sm + 5 * n;         // [n] off-diagonal of T
    double* se2 = sm + 6 * n;        // [n] e^2
    double* big = sm + 7 * n;        // matrix rows during phase 1, inverse-iteration vectors afterwards
    if (pre_tridiag) { sd = sm; se = sm + n; se2 = sm + 2 * n; big = sm + 3 * n; }    // no Householder buffers needed
    __shared__ double red[33];
    __shared__ double lam[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EIG_THREADS / 32;
    // global workspace: d[n], e[n], beta[n], then the per-vector arrays when they do not fit in shared memory
    double* gd = wk;
    double* ge = wk + n;
    double* beta = wk + 2 * n;
    double* gvec = wk + 3 * n;       // [32][n]

    // this CTA owns rows i with i % EIG_CL == c; local row li = i / EIG_CL
    // phase timestamps (ns, %globaltimer) for dmp2_debug_eig_phases: start, tridiag, bisect, invit, done
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(wk + 35 * n);
    auto stamp = [&](int i) { if (c == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); stamps[i] = t; } };
    // pre_tridiag: d, e, beta and V were produced by k_tridiag_grid; only CTA 0 has work left (no cluster barrier is
    // executed by anybody on this path)
    __shared__ double scr_a[EIG_THREADS / 32], scr_b[EIG_THREADS / 32];
    if (!pre_tridiag) {
    stamp(0);
    const int nloc = (n - c + EIG_CL - 1) / EIG_CL;
    double* rbase;
    int64_t rstride;
    if (rows_in_smem) { rbase = big; rstride = n; }
    else { rbase = A + (int64_t)c * n; rstride = (int64_t)EIG_CL * n; }
    for (int li = warp; li < nloc; li += NW) {
        const float* src = M + (int64_t)(li * EIG_CL + c) * n;
        double* dst = rbase + li * rstride;
        for (int j = lane; j < n; j += 32) dst[j] = (double)src[j];
    }

    // ---------------- 1. Householder tridiagonalisation ----------------------------------------------------
    // Fused form: iteration k applies the rank-2 update of column k (v, w) to the owned rows and, in the same sweep,
    // multiplies them with the NEXT Householder vector v' -- which every CTA can form on its own, because the raw
    // entries of the next column were all-gathered one iteration earlier and the pending update of that column only
    // needs v and w.  So there is ONE all-gather per column, carrying two doubles per row: p'_i = (A v')_i and the
    // row's freshly updated entry of the column after next.  Every value travels as an st.async store that
    // complete_tx's on the destination CTA's mbarrier: no cluster-wide barrier in the loop, a CTA waits only until its
    // own copy is complete.  Buffers and barriers are ping-ponged by column parity; a CTA can run at most one gather
    // ahead of its slowest peer.  Shared memory (7 n doubles, the d/e/e^2 arrays alias the tail):
    double* sv = sm;                 // [n] v of the column being applied (index = row - c0), v[0] = v0
    double* sw = sm + n;             // [n] w = p - kk v
    double* sv2 = sm + 2 * n;        // [n] next Householder vector (entry 0 is kept in a register: v0n)
    double* pgb = sm + 3 * n;        // [2][n] gathered p'
    double* xgb = sm + 5 * n;        // [2][n] gathered raw next column
    __shared__ __align__(8) unsigned long long gbar[2];
    const uint32_t gbar0 = cc::smem_u32(&gbar[0]);
    if (tid == 0) {
        cc::bar_init(gbar0, 1);
        cc::bar_init(gbar0 + 8, 1);
        cc::bar_init_fence();
    }
    for (int j = tid; j < n; j += EIG_THREADS) {
        sv[j] = 0.0; sw[j] = 0.0;
        xgb[j] = (double)M[(int64_t)j * n];                       // raw column 0 (parity 0 buffer)
    }
    const uint32_t pg_a = cc::smem_u32(pgb), xg_a = cc::smem_u32(xgb);
    cluster.sync();
    uint32_t gph = 0;                                             // phase parity bits of the two barriers
    for (int k = -1; k <= n - 3; k++) {
        const int c0 = k + 1;                          // first row / column of the trailing matrix
        const int m = n - c0;
        const int par = c0 & 1;                        // buffers read in this iteration; the gather fills par ^ 1
        const double* xcol = xgb + par * n;            // raw column c0, rows c0.. (index = row - c0)
        const bool last = (k == n - 3);
        if (tid == 0 && !last) cc::bar_expect_tx(gbar0 + 8 * (par ^ 1), (uint32_t)(m - 1) * 16u);
        // ---- A. finish column c0 locally: r = raw - (v w[0] + w v[0]);  d = r[0];  x = r[1..) -> next Householder vector
        const double v_0 = sv[0], w_0 = sw[0];
        double part = 0.0;
        for (int j = tid; j < m; j += EIG_THREADS) {
            const double r = xcol[j] - (sv[j] * w_0 + sw[j] * v_0);
            if (j == 0) { if (c == 0) gd[c0] = r; }
            else { sv2[j - 1] = r; part += r * r; }
        }
        const double sigma = block_sum1(part, scr_a);
        double bt2 = 0.0, v0n = 0.0;
        if (m >= 3) {
            const double x0 = sv2[0];
            const double tail = sigma - x0 * x0;
            const bool reflect = tail > 0.0;               // false: column already tridiagonal, H = I
            const double alpha = reflect ? ((x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma)) : x0;
            v0n = x0 - alpha;
            bt2 = reflect ? 2.0 / (tail + v0n * v0n) : 0.0;
            if (c == 0) {
                if (tid == 0) { ge[c0] = alpha; beta[c0] = bt2; }
                for (int j = tid; j < m - 1; j += EIG_THREADS) V[(int64_t)c0 * n + j] = j == 0 ? v0n : sv2[j];
            }
        } else if (c == 0 && tid == 0) {                    // last 2 x 2 block: no reflector left
            ge[c0] = sv2[0];
            ge[n - 1] = 0.0;
        }
        // ---- B. one sweep over the owned rows i > c0: update with (v, w), dot with v', all-gather (p'_i, A[i][c0+1])
        const int li0 = (c0 + 1 - c + EIG_CL - 1) / EIG_CL;           // first local row with global index > c0
        for (int li = li0 + warp; li < nloc; li += NW) {
            const int i = li * EIG_CL + c, ti = i - c0;
            double* row = rbase + li * rstride + c0;
            const double vi = sv[ti], wi = sw[ti];
            double acc = 0.0, acc2 = 0.0, xn = 0.0;
            int j = 1 + lane;
            if (j < m) {                                   // first chunk: column c0+1 is the one gathered for the next step
                const double a = row[j] - (vi * sw[j] + wi * sv[j]);
                row[j] = a;
                acc = a * (j == 1 ? v0n : sv2[j - 1]);
                xn = a;
                j += 32;
            }
#pragma unroll 2
            for (; j + 32 < m; j += 64) {
                const double a = row[j] - (vi * sw[j] + wi * sv[j]);
                const double b = row[j + 32] - (vi * sw[j + 32] + wi * sv[j + 32]);
                row[j] = a; row[j + 32] = b;
                acc += a * sv2[j - 1];
                acc2 += b * sv2[j + 31];
            }
            if (j < m) {
                const double a = row[j] - (vi * sw[j] + wi * sv[j]);
                row[j] = a;
                acc += a * sv2[j - 1];
            }
            acc = warp_sum(acc + acc2) * bt2;
            xn = __shfl_sync(0xffffffffu, xn, 0);
            if (last) {
                if (lane == 0 && i == n - 1) gd[n - 1] = xn;
            } else if (lane < EIG_CL) {
                const uint32_t off = (uint32_t)((par ^ 1) * n + ti - 1) * 8u;
                const uint32_t pbar = cc::mapa(gbar0, lane) + 8 * (par ^ 1);
                cc::st_async_f64(cc::mapa(pg_a, lane) + off, acc, pbar);
                cc::st_async_f64(cc::mapa(xg_a, lane) + off, xn, pbar);
            }
        }
        if (last) break;
        // ---- C. gathered p' -> w' = p' - kk' v';  v <- v'
        cc::bar_wait(gbar0 + 8 * (par ^ 1), (gph >> (par ^ 1)) & 1u);
        gph ^= 1u << (par ^ 1);
        const double* pg = pgb + (par ^ 1) * n;
        double pv = 0.0;
        for (int j = tid; j < m - 1; j += EIG_THREADS) pv += pg[j] * (j == 0 ? v0n : sv2[j]);
        const double kk = 0.5 * bt2 * block_sum1(pv, scr_b);
        for (int j = tid; j < m - 1; j += EIG_THREADS) {
            const double vj = j == 0 ? v0n : sv2[j];
            sv[j] = vj;
            sw[j] = pg[j] - kk * vj;
        }
        __syncthreads();
    }
    __threadfence();
    cluster.sync();
    stamp(1);
    }
    // phase 2 and phases 4-5: CTA w < 8 takes eigenvalue / eigenvector w; phase 3: CTA 0 (the others wait at a cluster barrier)
    double* lam_g = wk + 35 * n + 8;                   // [8] eigenvalues, exchanged through global memory (stamps sit at 35 n)
    double tnorm = 0.0;
    if (c < 8) {
    for (int i = tid; i < n; i += EIG_THREADS) { sd[i] = gd[i]; se[i] = ge[i]; se2[i] = ge[i] * ge[i]; }
    __syncthreads();
    // Gershgorin bounds
    double lo_p = 1e300, hi_p = -1e300;
    for (int i = tid; i < n; i += EIG_THREADS) {
        double r = (i > 0 ? fabs(se[i - 1]) : 0.0) + (i < n - 1 ? fabs(se[i]) : 0.0);
        lo_p = fmin(lo_p, sd[i] - r);
        hi_p = fmax(hi_p, sd[i] + r);
    }
    __shared__ double glo[NW], ghi[NW];
    for (int o = 16; o; o >>= 1) {
        lo_p = fmin(lo_p, __shfl_xor_sync(0xffffffffu, lo_p, o));
        hi_p = fmax(hi_p, __shfl_xor_sync(0xffffffffu, hi_p, o));
    }
    if (lane == 0) { glo[warp] = lo_p; ghi[warp] = hi_p; }
    __syncthreads();
    double gl = glo[0], gh = ghi[0];
    for (int i = 1; i < NW; i++) { gl = fmin(gl, glo[i]); gh = fmax(gh, ghi[i]); }
    tnorm = fmax(fabs(gl), fabs(gh));
    gl -= tnorm * 1e-12 + 1e-300;
    gh += tnorm * 1e-12 + 1e-300;

    // ---------------- 2. Sturm multi-section: CTA w finds eigenvalue index n-8+w (ascending) --------------
    // 128 shifts per round, one per thread of the first four warps (one warp per SM sub-partition; each shift is a serial
    // n-step Sturm recurrence and the fp64 pipe is what bounds it: measured 7 us per round and warp-per-partition at
    // n=300).  The bracket shrinks 129x per round: 8 rounds from the Gershgorin interval to 4 ulp, where 33 shifts in
    // one warp needed 11 and 513 shifts in sixteen warps 7 rounds of four times the work.
    {
        constexpr int NSH = 128;
        const int kidx = n - 8 + c;
        __shared__ int wcnt[2][NSH / 32];
        double lo = gl, hi = gh;
        for (int round = 0; round < 16; round++) {
            const double step = (hi - lo) / (double)(NSH + 1);
            if (tid < NSH) {
                const double x = lo + step * (double)(tid + 1);
                const int cnt = sturm_count(sd, se2, n, x);
                const unsigned bal = __ballot_sync(0xffffffffu, cnt <= kidx);
                if (lane == 0) wcnt[round & 1][warp] = __popc(bal);
            }
            __syncthreads();
            int npre = 0;                                       // shifts with at most kidx eigenvalues below them (a prefix)
#pragma unroll
            for (int i = 0; i < NSH / 32; i++) npre += wcnt[round & 1][i];
            const double nlo = npre > 0 ? lo + step * (double)npre : lo;
            const double nhi = npre < NSH ? lo + step * (double)(npre + 1) : hi;
            const bool done = !(nhi - nlo < hi - lo) || (nhi - nlo) <= 4.0 * 2.3e-16 * fmax(fabs(nlo), fabs(nhi));
            lo = nlo; hi = nhi;
            if (done) break;                                    // (uniform: every thread holds the same bracket)
        }
        if (tid == 0) lam_g[c] = 0.5 * (lo + hi);
    }
    __threadfence();
    }
    cluster.sync();
    if (c == 0) {
    if (tid < 8) lam[tid] = __ldcg(lam_g + tid);
    __syncthreads();

    stamp(2);
    // ---------------- 3. inverse iteration on T ---------------------------------------------------------
    // vec_in_smem: 0 = all work vectors in global memory, 1 = all 32 n doubles in shared memory,
    //              2 = the 8 eigenvectors in shared memory, the LU factors in global memory (large n)
    double* zs = vec_in_smem ? big : gvec;                       // [8][n] eigenvectors of T, then of M
    double* u0 = (vec_in_smem == 1 ? big : gvec) + 8 * n;        // [8][n] LU of T - lambda I: three diagonals of U
    double* u1 = u0 + 8 * n;
    double* u2 = u1 + 8 * n;
    for (int i = tid; i < 8 * n; i += EIG_THREADS) {
        int w = i / n, r = i - w * n;
        unsigned h = (unsigned)(r * 2654435761u) ^ (unsigned)((w + 1) * 40503u);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        zs[i] = 0.5 + (double)(h & 0xffff) / 65536.0;
    }
    __syncthreads();
    const double pivmin = fmax(tnorm * 2.3e-16, 1e-290);
    const int invit_iters = g_invit_iters;
    for (int iter = 0; iter < invit_iters; iter++) {
        double sc = 1.0;
        if (warp < 8 && lane == 0) {
            const int w = warp;
            const double l = lam[w];
            double* z = zs + w * n;
            double *U0 = u0 + w * n, *U1 = u1 + w * n, *U2 = u2 + w * n;
            // LU with partial pivoting of the tridiagonal, forward substitution in the same sweep
            double c0 = sd[0] - l, c1 = n > 1 ? se[0] : 0.0;          // current row i: (diag, super)
            double rhs = z[0];
            double nx_sub = se[0], nx_d = n > 1 ? sd[1] : 0.0, nx_b = n > 2 ? se[1] : 0.0, nx_r = n > 1 ? z[1] : 0.0;
            for (int i = 0; i < n - 1; i++) {
                const double sub = nx_sub;                              // T[i+1][i]
                const double an = nx_d - l;                             // T[i+1][i+1]
                const double bn = nx_b;                                 // T[i+1][i+2]
                const double rn = nx_r;
                if (i + 2 < n) {                                        // operands of the next row, off the dependent chain
                    nx_sub = se[i + 1]; nx_d = sd[i + 2]; nx_b = (i + 3 < n) ? se[i + 2] : 0.0; nx_r = z[i + 2];
                }
                if (fabs(sub) <= fabs(c0)) {
                    if (fabs(c0) < pivmin) c0 = c0 < 0 ? -pivmin : pivmin;
                    double rc = 1.0 / c0;
                    double mlt = sub * rc;
                    U0[i] = rc; U1[i] = c1; U2[i] = 0.0;
                    z[i] = rhs;
                    c0 = an - mlt * c1; c1 = bn; rhs = rn - mlt * rhs;
                } else {
                    double rs = 1.0 / sub;
                    double mlt = c0 * rs;
                    U0[i] = rs; U1[i] = an; U2[i] = bn;
                    z[i] = rn;
                    c0 = c1 - mlt * an; c1 = -mlt * bn; rhs = rhs - mlt * rn;
                }
            }
            if (fabs(c0) < pivmin) c0 = c0 < 0 ? -pivmin : pivmin;
            U0[n - 1] = 1.0 / c0; U1[n - 1] = 0.0; U2[n - 1] = 0.0;
            z[n - 1] = rhs;
            // back substitution (U0 holds reciprocal pivots; the two previous solutions stay in registers)
            double zmax = 0.0, z1 = 0.0, z2 = 0.0;
#pragma unroll 4
            for (int i = n - 1; i >= 0; i--) {
                double t = (z[i] - U1[i] * z1 - U2[i] * z2) * U0[i];
                z[i] = t;
                z2 = z1; z1 = t;
                zmax = fmax(zmax, fabs(t));
            }
            sc = zmax > 0 ? 1.0 / zmax : 1.0;                         // avoid overflow in the dot products
        }
        if (warp < 8) {                                               // the whole warp applies lane 0's scale
            sc = __shfl_sync(0xffffffffu, sc, 0);
            double* z = zs + warp * n;
            for (int i = lane; i < n; i += 32) z[i] *= sc;
        }
        __syncthreads();
        // Gram-Schmidt in ascending order + normalisation: warp p < w forms <z_p, z_w>, then one fused update
        __shared__ double dots[8];
        for (int w = 0; w < 8; w++) {
            double* z = zs + w * n;
            if (warp < w) {
                const double* zp = zs + warp * n;
                double a = 0.0;
                for (int i = lane; i < n; i += 32) a += zp[i] * z[i];
                a = warp_sum(a);
                if (lane == 0) dots[warp] = a;
            }
            __syncthreads();
            double nrm = 0.0;
            for (int i = tid; i < n; i += EIG_THREADS) {
                double v = z[i];
                for (int p = 0; p < w; p++) v -= dots[p] * zs[p * n + i];
                z[i] = v;
                nrm += v * v;
            }
            const double nr = block_sum1(nrm, (w & 1) ? scr_a : scr_b);
            const double inv = nr > 0 ? 1.0 / sqrt(nr) : 0.0;
            for (int i = tid; i < n; i += EIG_THREADS) z[i] *= inv;
            __syncthreads();
        }
    }

    stamp(3);
    // hand the eigenvectors of T (and the eigenvalues) to the other CTAs through global memory
    if (vec_in_smem) for (int i = tid; i < 8 * n; i += EIG_THREADS) gvec[i] = zs[i];
    __threadfence();
    }
    cluster.sync();
    if (c >= 8) return;
    // ---------------- 4. back-transform: z <- H_0 H_1 ... H_{n-3} z, ONE CTA PER EIGENVECTOR ---------------
    // The eight vectors are independent, and so far seven (or fifteen) CTAs of the cluster had nothing left to do.  CTA w
    // keeps vector w in shared memory, thread t owning the rows j = t (mod EIG_THREADS) for the whole phase (no
    // cross-thread hazard on z); the reflectors are staged `srows` at a time into shared memory by all threads (one L2
    // round trip per block); per reflector: partial dot, warp shuffle, ONE block barrier, total, update.
    {
        const int w = c;
        double* z = sm;                                         // [n]
        double* stage = sm + n;                                 // [srows][n]
        __shared__ __align__(16) double part[2][3][EIG_THREADS / 32];
        static_assert(EIG_THREADS == 512, "the partial-sum tree below is written for 16 warps");
        __shared__ double sbeta[64];
        const int srows = stage_rows;
        for (int i = tid; i < n; i += EIG_THREADS) z[i] = __ldcg(gvec + (int64_t)w * n + i);
        const double lamw = __ldcg(lam_g + w);
        int pp = 0;
        const bool act = warp * 32 < n;                         // this warp owns at least one row
        if (!act && lane < 6) part[lane & 1][lane >> 1][warp] = 0.0;      // its partial sums stay zero
        for (int khi = n - 3; khi >= 0; khi -= srows) {
            const int klo = khi - (srows - 1) > 0 ? khi - (srows - 1) : 0;
            __syncthreads();                                    // the previous block of reflectors is no longer read
            const int tot = (khi - klo + 1) * n;
#pragma unroll 4
            for (int idx = tid; idx < tot; idx += EIG_THREADS) {
                const int r = idx / n, i = idx - r * n;
                if (i < n - (klo + r) - 1) stage[idx] = __ldcg(V + (int64_t)(klo + r) * n + i);
            }
            if (tid <= khi - klo) sbeta[tid] = beta[klo + tid];
            __syncthreads();
            int k = khi;
            for (; k - 1 >= klo; k -= 2) {
                // reflectors k (applied first) and k-1 in ONE barrier round: with p = v_k.z, q = v_{k-1}.z, r = v_{k-1}.v_k
                //   a_k = beta_k p,   a_{k-1} = beta_{k-1} (q - a_k r),   z -= a_k v_k + a_{k-1} v_{k-1}
                // (the fp64 pipe is what this loop waits for: the reductions are arranged for few fp64 instructions, and
                //  warps that own no row only keep the barriers)
                const double* vk = stage + (k - klo) * n - (k + 1);         // vk[j]: component of row j (rows j > k)
                const double* vm = stage + (k - 1 - klo) * n - k;           // vm[j]: rows j > k-1
                if (act) {
                    double pa = 0.0, qa = 0.0, ra = 0.0;
                    for (int j = tid; j < n; j += EIG_THREADS) {
                        if (j > k) { const double a = vk[j], b = vm[j], zz = z[j]; pa += a * zz; qa += b * zz; ra += a * b; }
                        else if (j == k) qa += vm[j] * z[j];
                    }
                    const double pq = warp_sum_pair(pa, qa, lane);          // lanes < 16: sum of pa, lanes >= 16: sum of qa
                    ra = warp_sum(ra);
                    if ((lane & 15) == 0) part[pp][lane >> 4][warp] = pq;
                    if (lane == 0) part[pp][2][warp] = ra;
                }
                __syncthreads();
                if (act) {
                    double pq = part[pp][lane >> 4][lane & 15], rr = part[pp][2][lane & 15];
#pragma unroll
                    for (int o = 8; o; o >>= 1) { pq += __shfl_xor_sync(0xffffffffu, pq, o); rr += __shfl_xor_sync(0xffffffffu, rr, o); }
                    const double ak = sbeta[k - klo] * __shfl_sync(0xffffffffu, pq, 0);
                    const double am = sbeta[k - 1 - klo] * (__shfl_sync(0xffffffffu, pq, 16) - ak * rr);
                    for (int j = tid; j < n; j += EIG_THREADS) {
                        if (j > k) z[j] -= ak * vk[j] + am * vm[j];
                        else if (j == k) z[j] -= am * vm[j];
                    }
                }
                pp ^= 1;
            }
            if (k >= klo) {                                     // odd one left in this block
                const double* v = stage + (k - klo) * n - (k + 1);
                if (act) {
                    double a = 0.0;
                    for (int j = tid; j < n; j += EIG_THREADS) if (j > k) a += v[j] * z[j];
                    a = warp_sum(a);
                    if (lane == 0) part[pp][0][warp] = a;
                }
                __syncthreads();
                if (act) {
                    double t = part[pp][0][lane & 15];
#pragma unroll
                    for (int o = 8; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    t *= sbeta[k - klo];
                    for (int j = tid; j < n; j += EIG_THREADS) if (j > k) z[j] -= t * v[j];
                }
                pp ^= 1;
            }
        }
        __syncthreads();
        // ---------------- 5. canonical sign + MDS scaling -------------------------------------------------
        double best = -1.0; int bi = 0;
        for (int i = tid; i < n; i += EIG_THREADS) {
            double a = fabs((double)(float)z[i]);
            if (a > best) { best = a; bi = i; }
        }
        for (int o = 16; o; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        __shared__ double wbest[EIG_THREADS / 32];
        __shared__ int wbi[EIG_THREADS / 32];
        if (lane == 0) { wbest[warp] = best; wbi[warp] = bi; }
        __syncthreads();
        for (int i = 0; i < EIG_THREADS / 32; i++) {
            const double ob = wbest[i]; const int oi = wbi[i];
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const double sgn = z[bi] < 0.0 ? -1.0 : 1.0;
        const float lf = (float)lamw;
        const float sc = sqrtf(fmaxf(fmaxf(lf, 0.0f), 1e-8f));
        for (int i = tid; i < n; i += EIG_THREADS) {
            float vf = (float)(z[i] * sgn);
            if (vec_out) vec_out[(int64_t)i * 8 + w] = vf;
            if (mds_out) mds_out[(int64_t)i * 8 + w] = vf * sc;
        }
        if (tid == 0 && vals_out) vals_out[w] = lf;
    }
    __syncthreads();
    stamp(4);
}

// ---------------------------------------------------------------------------------------------------
// Large L (the rows no longer fit a cluster's shared memory): Householder tridiagonalisation on the WHOLE GPU.
// One CTA per SM (cooperative launch), matrix in global memory (L2-resident: 33 MB of fp64 at L = 2048), rows owned
// cyclically by CTA.  The rank-2 update of step k and the matrix-vector product of step k+1 are fused into ONE pass
// over the trailing matrix: every CTA first recomputes the updated row k+1 by itself (it only needs v, w and the old
// row), which gives the next Householder vector v' without waiting for anybody; each owned row is then read once,
// updated, written, and dotted with v' on the fly.  One grid-wide barrier per column (the all-gather of p' = A v'),
// half the memory traffic of update-then-multiply.  Writes d, e, beta and the reflectors V exactly like phase 1 of
// k_eig_top8, which then runs phases 2-5 (pre_tridiag = 1).
// ---------------------------------------------------------------------------------------------------
#define TG_THREADS 512
constexpr int TG_NW = TG_THREADS / 32;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all threads of all CTAs; `target` = CTAs x barriers passed so far (monotonic counter, zeroed before the launch)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        if (ld_acquire_gpu_u32(bar) < target) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(bar) < target) {
                if (clock64() - t0 > 8000000000LL) { printf("k_tridiag_grid: grid barrier timed out\n"); __trap(); }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TG_THREADS, 1)
k_tridiag_grid(const float* __restrict__ M, int n, double* __restrict__ A, double* __restrict__ V, double* __restrict__ wk,
               unsigned int* __restrict__ bar) {
    const int G = gridDim.x, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ double sm[];
    double* sv = sm;                 // [n] Householder vector of the column being eliminated (v[0] = v0)
    double* sw = sm + n;             // [n] w = p - kk v
    double* sv2 = sm + 2 * n;        // [n] next Householder vector
    __shared__ double scr_a[TG_NW], scr_b[TG_NW], spart[TG_NW];
    double* gd = wk;
    double* ge = wk + n;
    double* beta = wk + 2 * n;
    double* pbuf = wk + 40 * (int64_t)n;                 // [2][n] gathered p' (ping-pong by column parity)
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(wk + 35 * (int64_t)n);
    if (b == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); stamps[0] = t; }
    unsigned int nbar = 0;

    for (int i = b; i < n; i += G) {                     // fp32 -> fp64, owned rows
        const float* src = M + (int64_t)i * n;
        double* dst = A + (int64_t)i * n;
        for (int j = tid; j < n; j += TG_THREADS) dst[j] = (double)src[j];
    }
    for (int j = tid; j < n; j += TG_THREADS) { sv[j] = 0.0; sw[j] = 0.0; }
    grid_barrier(bar, (unsigned)G * ++nbar);

    // iteration k eliminates nothing itself: it applies the update of column k (v, w; none for k = -1) and prepares
    // column k+1 (d, e, v', p').  m = rows/columns of the trailing matrix A[k+1.., k+1..].
    for (int k = -1; k <= n - 3; k++) {
        const int m = n - k - 1;
        const int c0 = k + 1;                                                 // first trailing row / column
        // ---- updated row c0, recomputed by every CTA: r[j] = A[c0][c0+j] - (v[0] w[j] + w[0] v[j])
        const double* r0 = A + (int64_t)c0 * n + c0;
        const double v_0 = sv[0], w_0 = sw[0];
        double part = 0.0;
        for (int j = tid; j < m; j += TG_THREADS) {
            const double r = __ldcg(r0 + j) - (v_0 * sw[j] + w_0 * sv[j]);
            if (j == 0) { if (b == 0) gd[c0] = r; }
            else { sv2[j - 1] = r; part += r * r; }
        }
        const double sigma = block_sum1(part, scr_a);                         // includes the barrier that publishes sv2
        double bt2 = 0.0;
        const int m2 = m - 1;                                                 // length of the next Householder vector
        if (m >= 3) {
            const double x0 = sv2[0];
            const double tail = sigma - x0 * x0;
            const bool reflect = tail > 0.0;
            const double alpha = reflect ? ((x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma)) : x0;
            const double v0 = x0 - alpha;
            bt2 = reflect ? 2.0 / (tail + v0 * v0) : 0.0;
            __syncthreads();                                                  // everyone has read sv2[0]
            if (tid == 0) sv2[0] = v0;
            if (b == 0 && tid == 0) { ge[c0] = alpha; beta[c0] = bt2; }
            __syncthreads();
            if (b == 0) for (int j = tid; j < m2; j += TG_THREADS) V[(int64_t)c0 * n + j] = sv2[j];
        } else {                                                              // last 2 x 2 block: no reflector left
            if (b == 0 && tid == 0) { ge[c0] = sv2[0]; ge[n - 1] = 0.0; }
            __syncthreads();
            if (tid == 0) sv2[0] = 0.0;
            __syncthreads();
        }
        // ---- one pass over the owned rows i > c0: update with (v, w), then dot with v'
        double* pout = pbuf + (int64_t)((k + 1) & 1) * n;
        const int first = c0 + 1 + ((b - (c0 + 1)) % G + G) % G;             // smallest owned row index > c0
        const int nl = first < n ? (n - 1 - first) / G + 1 : 0;               // owned trailing rows
        int W = TG_NW;                                                        // warps cooperating on one row
        while (W > 1 && nl * W > TG_NW) W >>= 1;
        const int per_round = TG_NW / W, grp = warp / W, sub = warp % W;
        for (int base = 0; base < nl; base += per_round) {
            const int lr = base + grp;
            double acc = 0.0;
            int i = 0;
            if (lr < nl) {
                i = first + lr * G;
                const int ti = i - c0;
                double* row = A + (int64_t)i * n + c0;
                const double vi = sv[ti], wi = sw[ti];
                const int S = 32 * W;
                int j = 1 + sub * 32 + lane;
                for (; j + 3 * S < m; j += 4 * S) {
                    double a0 = __ldcg(row + j), a1 = __ldcg(row + j + S), a2 = __ldcg(row + j + 2 * S), a3 = __ldcg(row + j + 3 * S);
                    a0 -= vi * sw[j] + wi * sv[j];
                    a1 -= vi * sw[j + S] + wi * sv[j + S];
                    a2 -= vi * sw[j + 2 * S] + wi * sv[j + 2 * S];
                    a3 -= vi * sw[j + 3 * S] + wi * sv[j + 3 * S];
                    __stcg(row + j, a0); __stcg(row + j + S, a1); __stcg(row + j + 2 * S, a2); __stcg(row + j + 3 * S, a3);
                    acc += a0 * sv2[j - 1] + a1 * sv2[j + S - 1] + a2 * sv2[j + 2 * S - 1] + a3 * sv2[j + 3 * S - 1];
                }
                for (; j < m; j += S) {
                    double a0 = __ldcg(row + j) - (vi * sw[j] + wi * sv[j]);
                    __stcg(row + j, a0);
                    acc += a0 * sv2[j - 1];
                }
                acc = warp_sum(acc);
            }
            if (lane == 0) spart[warp] = acc;
            __syncthreads();
            if (lr < nl && sub == 0 && lane == 0) {
                double t = 0.0;
                for (int q = 0; q < W; q++) t += spart[warp + q];
                pout[i - c0 - 1] = t * bt2;
            }
            __syncthreads();
        }
        grid_barrier(bar, (unsigned)G * ++nbar);
        if (k == n - 3) break;
        // ---- gathered p' -> w' = p' - kk' v'; v <- v'
        double pv = 0.0;
        for (int j = tid; j < m2; j += TG_THREADS) {
            const double pj = __ldcg(pout + j);
            sw[j] = pj;
            pv += pj * sv2[j];
        }
        const double kk = 0.5 * bt2 * block_sum1(pv, scr_b);
        for (int j = tid; j < m2; j += TG_THREADS) { const double vj = sv2[j]; sw[j] -= kk * vj; sv[j] = vj; }
        __syncthreads();
    }
    if (b == 0 && tid == 0) {
        gd[n - 1] = __ldcg(A + (int64_t)(n - 1) * n + (n - 1));
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); stamps[1] = t;
    }
}

template <int CL>
static int launch_eig(dmp2_engine* e, const float* m, int L, int rows_in_smem, int vec_in_smem, int pre_tridiag, size_t smem,
                      float* vals, float* mds_scaled, float* vecs_raw, cudaStream_t st, int stage_rows = 16) {
    double* V = e->ws.eig_a + (int64_t)L * L;
    // phase 4 keeps one eigenvector (L doubles) + a block of reflectors in every CTA's shared memory
    stage_rows = (int)std::min<size_t>(64, (smem / 8 - (size_t)L) / (size_t)L);
    if (stage_rows < 1) return e->fail(DMP2_ERR_UNSUPPORTED, "eig_top8: L too large");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL);
    cfg.blockDim = dim3(EIG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(e, cudaLaunchKernelEx(&cfg, k_eig_top8<CL>, m, L, e->ws.eig_a, V, e->ws.eig_w, rows_in_smem, vec_in_smem, pre_tridiag, stage_rows, vals,
                                   mds_scaled, vecs_raw));
    POST_LAUNCH(e, "k_eig_top8");
    return 0;
}

int run_eig_top8(dmp2_engine* e, const float* m, int L, float* vals, float* mds_scaled, float* vecs_raw, cudaStream_t st) {
    constexpr size_t LIMIT = 216 * 1024;
    if (!e->attr_eig) {
        if (getenv("DMP2_EIG_INVIT") && atoi(getenv("DMP2_EIG_INVIT")) >= 1) {
            const int it = atoi(getenv("DMP2_EIG_INVIT"));
            CUDA_TRY(e, cudaMemcpyToSymbol(g_invit_iters, &it, sizeof(int)));
        }
        CUDA_TRY(e, cudaFuncSetAttribute(k_eig_top8<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT));
        CUDA_TRY(e, cudaFuncSetAttribute(k_eig_top8<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT));
        CUDA_TRY(e, cudaFuncSetAttribute(k_eig_top8<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        e->attr_eig = true;
    }
    const size_t n = (size_t)L;
    const size_t vec = 32 * n;
    // cluster of 8 while the rows fit its shared memory; beyond that a (non-portable) 16-CTA cluster: rows stay
    // shared-memory resident up to L ~ 650 and, past that, twice as many SMs stream them from L2
    auto plan = [&](int cl, int& rows_in_smem, int& vec_in_smem, size_t& smem) {
        const size_t rows = ((n + cl - 1) / cl) * n;
        size_t big = 0;
        rows_in_smem = 0; vec_in_smem = 0;
        if ((7 * n + std::max(rows, vec)) * 8 <= LIMIT) { rows_in_smem = 1; vec_in_smem = 1; big = std::max(rows, vec); }
        else if ((7 * n + rows) * 8 <= LIMIT) { rows_in_smem = 1; big = rows; }
        else if ((7 * n + vec) * 8 <= LIMIT) { vec_in_smem = 1; big = vec; }
        smem = (7 * n + big) * 8;
    };
    int r8, v8, r16, v16;
    size_t s8, s16;
    plan(8, r8, v8, s8);
    // With the fused sweep the column time is compute-bound, so from L ~ 200 on the 16-CTA cluster is faster even though
    // the rows would fit 8 CTAs (L=300: 1.34 vs 1.46 ms, profiles/round1_eig_cl16_L300.txt).  DMP2_EIG_CL=8|16 forces one.
    static const int force_cl = getenv("DMP2_EIG_CL") ? atoi(getenv("DMP2_EIG_CL")) : 0;      // tuning knob
    const bool prefer16 = force_cl == 16 || (force_cl != 8 && L >= 200 && !e->eig_no_cl16);
    if (r8 && !prefer16) return launch_eig<8>(e, m, L, r8, v8, 0, s8, vals, mds_scaled, vecs_raw, st);
    plan(16, r16, v16, s16);
    // rows do not even fit a 16-CTA cluster: tridiagonalise on the whole GPU, then phases 2-5 in one CTA
    static const bool no_grid = getenv("DMP2_EIG_GRID") && atoi(getenv("DMP2_EIG_GRID")) == 0;       // tuning knob
    if (!r16 && !no_grid && 3 * n * 8 <= LIMIT) {
        if (!e->attr_eig_grid) {
            CUDA_TRY(e, cudaFuncSetAttribute(k_tridiag_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT));
            e->attr_eig_grid = true;
        }
        double* Vg = e->ws.eig_a + (int64_t)L * L;
        unsigned int* bar = reinterpret_cast<unsigned int*>(e->ws.eig_w + 42 * (int64_t)L);
        CUDA_TRY(e, cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
        int n_arg = L;
        const float* m_arg = m;
        double* a_arg = e->ws.eig_a;
        double* wk_arg = e->ws.eig_w;
        void* args[6] = {(void*)&m_arg, (void*)&n_arg, (void*)&a_arg, (void*)&Vg, (void*)&wk_arg, (void*)&bar};
        CUDA_TRY(e, cudaLaunchCooperativeKernel((const void*)k_tridiag_grid, dim3(e->num_sms), dim3(TG_THREADS), args, 3 * n * 8, st));
        POST_LAUNCH(e, "k_tridiag_grid");
        // phases 2-5 in CTA 0: d, e, e^2 (3 n) + as much of the work vectors as fits
        const size_t cap = LIMIT / 8;
        int vmode = 0;
        size_t words = 3 * n;
        if (35 * n <= cap) { vmode = 1; words = 35 * n; }
        else if (11 * n <= cap) { vmode = 2; words = 11 * n; }
        if (words > cap) return e->fail(DMP2_ERR_UNSUPPORTED, "eig_top8: L too large");
        return launch_eig<8>(e, m, L, 0, vmode, 1, LIMIT, vals, mds_scaled, vecs_raw, st);     // all of it: phase 4 stages reflectors
    }
    if (s16 > LIMIT) return e->fail(DMP2_ERR_UNSUPPORTED, "eig_top8: L too large");
    if (!e->eig_no_cl16) {
        if (launch_eig<16>(e, m, L, r16, v16, 0, s16, vals, mds_scaled, vecs_raw, st) == 0) return 0;
        cudaGetLastError();                              // 16-CTA clusters not schedulable here: use 8 from now on
        e->eig_no_cl16 = true;
        e->status = 0;
        e->err.clear();
    }
    if (s8 > LIMIT) return e->fail(DMP2_ERR_UNSUPPORTED, "eig_top8: L too large");
    return launch_eig<8>(e, m, L, r8, v8, 0, s8, vals, mds_scaled, vecs_raw, st);
}
