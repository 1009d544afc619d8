// The three recurrent stages: vgru down the MSA (network.py:223-224), hgru along the sequence (:225) and
// the coordinate GRU + linear head (:251-255).  All arithmetic fp32 (GRU weights do not tolerate tf32/fp16:
// SURVEY.md section 7.3).
#include "common.cuh"
#include "sgemm.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// vgru: time axis = MSA rows (N steps), batch = L columns.  One GEMM per layer per step with the GRU cell
// fused into the epilogue: weight rows are packed 4 per hidden unit so a float4 of accumulators carries all
// gate pre-activations of one (column, unit) pair.
// ---------------------------------------------------------------------------------------------------
struct VgCell0 {                 // layer 0: acc = (W_hr h, W_hz h, W_hn h, 0)
    const float* gi0;            // [22][512][4]
    const float* b;              // [2048]
    const uint8_t* codes;        // msa row t: [L]
    const float* h_old;          // [L][512]
    float* h_new;
    __device__ void operator()(int m, int n, float4 a) const {
        int j = n >> 2;
        float4 gi = *reinterpret_cast<const float4*>(gi0 + ((int64_t)codes[m] * 512 + j) * 4);
        float4 bb = *reinterpret_cast<const float4*>(b + n);
        float r = sigmoid_acc(gi.x + (a.x + bb.x));
        float z = sigmoid_acc(gi.y + (a.y + bb.y));
        float nn = tanhf(gi.z + r * (a.z + bb.z));
        float ho = h_old[(int64_t)m * 512 + j];
        h_new[(int64_t)m * 512 + j] = (1.0f - z) * nn + z * ho;
    }
};
struct VgCell1 {                 // layer 1: acc = (gi_r+gh_r, gi_z+gh_z, gi_n, gh_n) without biases
    const float* b;              // [2048]
    const float* h_old;
    float* h_new;
    __device__ void operator()(int m, int n, float4 a) const {
        int j = n >> 2;
        float4 bb = *reinterpret_cast<const float4*>(b + n);
        float r = sigmoid_acc(a.x + bb.x);
        float z = sigmoid_acc(a.y + bb.y);
        float nn = tanhf((a.z + bb.z) + r * (a.w + bb.w));
        float ho = h_old[(int64_t)m * 512 + j];
        h_new[(int64_t)m * 512 + j] = (1.0f - z) * nn + z * ho;
    }
};
struct LoadConcat2 {             // A(m,k) = k < k1 ? p1[m*ld1+k] : p2[m*ld2 + k-k1]
    static constexpr bool n_major = false;
    const float* p1; int ld1; int k1; const float* p2; int ld2;
    __device__ float4 operator()(int m, int k) const {
        return k < k1 ? *reinterpret_cast<const float4*>(p1 + (int64_t)m * ld1 + k)
                      : *reinterpret_cast<const float4*>(p2 + (int64_t)m * ld2 + (k - k1));
    }
};

int run_vgru(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st) {
    return e->vgru_mode == 0 ? run_vgru_tc(e, msa, N, L, out, st) : run_vgru_ffma(e, msa, N, L, out, st);
}

int run_vgru_ffma(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st) {
    const Weights& w = e->w;
    const int64_t hsz = (int64_t)L * 512;
    float* h0[2] = {e->ws.vg_h, e->ws.vg_h + hsz};
    float* h1[2] = {e->ws.vg_h + 2 * hsz, e->ws.vg_h + 3 * hsz};
    CUDA_TRY(e, cudaMemsetAsync(e->ws.vg_h, 0, 4 * hsz * sizeof(float), st));
    for (int t = 0; t < N; t++) {
        int cur = t & 1, nxt = cur ^ 1;
        sgemm_launch<4>(L, 2048, 512, LoadRowMajorK{h0[cur], 512}, LoadRowMajorK{w.vg_w0, 512},
                        VgCell0{w.vg_gi0, w.vg_b0, msa + (int64_t)t * L, h0[cur], h0[nxt]}, st);
        POST_LAUNCH(e, "sgemm<vgru0>");
        sgemm_launch<4>(L, 2048, 1024, LoadConcat2{h0[nxt], 512, 512, h1[cur], 512}, LoadRowMajorK{w.vg_w1, 1024},
                        VgCell1{w.vg_b1, h1[cur], h1[nxt]}, st);
        POST_LAUNCH(e, "sgemm<vgru1>");
    }
    CUDA_TRY(e, cudaMemcpyAsync(out, h1[N & 1], hsz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Bidirectional GRU, hidden 256, batch 1, time L (hgru and coord_gru).  The input projections of a layer
// are one GEMM; the recurrence runs in one launch: an 8-CTA cluster per direction, each CTA keeps the
// 96 x 256 slice of W_hh for its 32 hidden units in REGISTERS (64 per thread), reads h from shared
// memory and broadcasts its 32 new values to the 7 peers through distributed shared memory each step.
// ---------------------------------------------------------------------------------------------------
// Per step the 32 new hidden values of a CTA travel to every peer as st.async stores that complete_tx on the
// destination CTA's mbarrier: data and "ready" signal in one DSMEM operation, no cluster-wide barrier in the loop.
__device__ __forceinline__ uint32_t gru_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gru_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void gru_st_async(uint32_t dst, float v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(dst), "r"(__float_as_uint(v)), "r"(bar) : "memory");
}
__device__ __forceinline__ void gru_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) { printf("k_bigru_rec: barrier timeout\n"); __trap(); }
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(384, 1)
k_bigru_rec(const float* __restrict__ gi, const float* __restrict__ whh_f, const float* __restrict__ bhh_f,
            const float* __restrict__ whh_b, const float* __restrict__ bhh_b, float* __restrict__ out, int L) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int dir = blockIdx.y;
    const float* whh = dir ? whh_b : whh_f;
    const float* bhh = dir ? bhh_b : bhh_f;
    __shared__ __align__(16) float hbuf[2][256];
    __shared__ float part[4][96];
    __shared__ __align__(8) unsigned long long hbar[2];
    const int tid = threadIdx.x;
    const int r = tid % 96, q = tid / 96;
    const int g = r >> 5, jj = r & 31;
    float wreg[64];
    {
        const float* src = whh + (int64_t)(g * 256 + 32 * c + jj) * 256 + 64 * q;
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            float4 v = *reinterpret_cast<const float4*>(src + k);
            wreg[k] = v.x; wreg[k + 1] = v.y; wreg[k + 2] = v.z; wreg[k + 3] = v.w;
        }
    }
    float b_r = 0.f, b_z = 0.f, b_n = 0.f;
    const int j = 32 * c + tid;                       // hidden unit of the gate threads (tid < 32)
    if (tid < 32) { b_r = bhh[j]; b_z = bhh[256 + j]; b_n = bhh[512 + j]; }
    for (int i = tid; i < 512; i += 384) (&hbuf[0][0])[i] = 0.f;
    const uint32_t bar0 = gru_smem_u32(&hbar[0]);     // barrier b lives at bar0 + 8*b
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // peer addresses of my slot in the h buffers and of the peers' barriers
    uint32_t peer_h[8], peer_bar[8];                 // buffer/barrier 0; buffer 1 is +1024 bytes, barrier 1 is +8 bytes
    if (tid < 32) {
#pragma unroll
        for (int d = 0; d < 8; d++) {
            peer_h[d] = gru_mapa(gru_smem_u32(&hbuf[0][j]), d);
            peer_bar[d] = gru_mapa(bar0, d);
        }
    }
    cluster.sync();
    uint32_t ph = 0;                                  // phase parity bits of the two barriers
    for (int t = 0; t < L; t++) {
        const int te = dir ? (L - 1 - t) : t;
        const int cur = t & 1, nxt = cur ^ 1;
        float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f;
        if (tid < 32) {
            const float* gp = gi + (int64_t)te * 1536 + dir * 768 + j;
            gi_r = gp[0]; gi_z = gp[256]; gi_n = gp[512];
        }
        if (tid == 0 && t + 1 < L)                    // arm the barrier that collects the 256 values of step t+1's input
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * nxt), "r"(1024u) : "memory");
        if (t > 0) { gru_bar_wait(bar0 + 8 * cur, (ph >> cur) & 1u); ph ^= 1u << cur; }
        const float4* h4 = reinterpret_cast<const float4*>(&hbuf[cur][64 * q]);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float4 hv = h4[k];
            a0 = fmaf(wreg[4 * k + 0], hv.x, a0);
            a1 = fmaf(wreg[4 * k + 1], hv.y, a1);
            a2 = fmaf(wreg[4 * k + 2], hv.z, a2);
            a3 = fmaf(wreg[4 * k + 3], hv.w, a3);
        }
        part[q][r] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (tid < 32) {
            float gh_r = ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid])) + b_r;
            float gh_z = ((part[0][32 + tid] + part[1][32 + tid]) + (part[2][32 + tid] + part[3][32 + tid])) + b_z;
            float gh_n = ((part[0][64 + tid] + part[1][64 + tid]) + (part[2][64 + tid] + part[3][64 + tid])) + b_n;
            float rr = sigmoid_acc(gi_r + gh_r);
            float zz = sigmoid_acc(gi_z + gh_z);
            float nn = tanhf(gi_n + rr * gh_n);
            float hn = (1.0f - zz) * nn + zz * hbuf[cur][j];
            out[(int64_t)te * 512 + dir * 256 + j] = hn;
            if (t + 1 < L) {
#pragma unroll
                for (int d = 0; d < 8; d++) gru_st_async(peer_h[d] + nxt * 1024, hn, peer_bar[d] + nxt * 8);
            }
        }
        // No second block barrier: `part` is rewritten only after the next step's mbarrier wait, which completes
        // only once this CTA's own gate threads have issued their stores, i.e. finished reading `part`.
    }
    cluster.sync();                                   // nobody exits while a peer could still be storing into it
}

// layer input [L][k1 (+ k2)] * DMP2_GRU_SA -> fp16 hi/lo [L][Kp] (zero-padded), the A operand of the tensor-core projection
__global__ void __launch_bounds__(256) k_gru_operand(const float* __restrict__ p1, int ld1, int k1, const float* __restrict__ p2, int ld2,
                                                     int k2, int L, int Kp, __half* __restrict__ hi, __half* __restrict__ lo) {
    const int n = L * Kp;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int r = idx / Kp, k = idx - r * Kp;
        float v = 0.f;
        if (k < k1) v = p1[(int64_t)r * ld1 + k];
        else if (k < k1 + k2) v = p2[(int64_t)r * ld2 + (k - k1)];
        v *= DMP2_GRU_SA;
        const __half h = __float2half_rn(v);
        hi[idx] = h;
        lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

// gi = x W_ih^T + b_ih on the tcgen05 GEMM (3 MMAs per MAC, chains of 128)
static int gru_project_tc(dmp2_engine* e, const BiGruLayer& ly, const float* p1, int ld1, int k1, const float* p2, int ld2, int k2,
                          int L, cudaStream_t st) {
    const int Kp = (ly.K + 63) & ~63;
    __half* a_hi = e->ws.tc_scratch;
    __half* a_lo = a_hi + (int64_t)L * Kp;
    k_gru_operand<<<std::min(cdiv(L * Kp, 256), e->num_sms * 4), 256, 0, st>>>(p1, ld1, k1, p2, ld2, k2, L, Kp, a_hi, a_lo);
    POST_LAUNCH(e, "k_gru_operand");
    GemmTcEpilogue ep{0, 0, nullptr, nullptr, nullptr};
    ep.bias = ly.b_ih;
    return run_gemm_tc(e, a_hi, a_lo, ly.w_ih_hi, ly.w_ih_lo, L, 1536, Kp, 1.0f / (DMP2_GRU_SA * DMP2_GRU_SW), e->ws.gi, 1536, 128, st, &ep);
}

template <class AL>
static int bigru_stack(dmp2_engine* e, const BiGruLayer* layers, int nlayers, AL first, int L, float* out, cudaStream_t st,
                       const float* in1 = nullptr, int ld1 = 0, int k1 = 0, const float* in2 = nullptr, int ld2 = 0, int k2 = 0) {
    Workspace& ws = e->ws;
    float* bufs[2] = {ws.seq_a, ws.seq_b};
    const float* prev = nullptr;
    for (int k = 0; k < nlayers; k++) {
        const BiGruLayer& ly = layers[k];
        if (e->gemm_tc && in1 && L <= DMP2_TC_SLAB) {
            if (k == 0) TRY(gru_project_tc(e, ly, in1, ld1, k1, in2, ld2, k2, L, st));
            else TRY(gru_project_tc(e, ly, prev, 512, 512, nullptr, 0, 0, L, st));
        } else if (k == 0)
            sgemm_launch<4>(L, 1536, ly.K, first, LoadRowMajorK{ly.w_ih, ly.K}, StoreRowMajor{ws.gi, 1536, ly.b_ih, 1.0f}, st);
        else
            sgemm_launch<4>(L, 1536, ly.K, LoadRowMajorK{prev, 512}, LoadRowMajorK{ly.w_ih, ly.K},
                            StoreRowMajor{ws.gi, 1536, ly.b_ih, 1.0f}, st);
        POST_LAUNCH(e, "sgemm<gru_ih>");
        float* dst = (k == nlayers - 1) ? out : bufs[k & 1];
        k_bigru_rec<<<dim3(8, 2), 384, 0, st>>>(ws.gi, ly.dir[0].w_hh, ly.dir[0].b_hh, ly.dir[1].w_hh, ly.dir[1].b_hh, dst, L);
        POST_LAUNCH(e, "k_bigru_rec");
        prev = dst;
    }
    return 0;
}

int run_bigru(dmp2_engine* e, const BiGruLayer* layers, int nlayers, const float* in, int L, float* out, cudaStream_t st) {
    return bigru_stack(e, layers, nlayers, LoadRowMajorK{in, layers[0].K}, L, out, st, in, layers[0].K, layers[0].K);
}

// ca[t][d] = sum_k h[t][k] * W[d][k]                                                  (network.py:255)
__global__ void __launch_bounds__(128) k_coord_fc(const float* __restrict__ h, const float* __restrict__ w, int L,
                                                  float* __restrict__ ca) {
    int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= L) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane; k < 512; k += 32) {
        float v = h[(int64_t)t * 512 + k];
        a0 = fmaf(v, w[k], a0); a1 = fmaf(v, w[512 + k], a1); a2 = fmaf(v, w[1024 + k], a2);
    }
    for (int o = 16; o; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) { ca[t * 3] = a0; ca[t * 3 + 1] = a1; ca[t * 3 + 2] = a2; }
}

int run_coord_head(dmp2_engine* e, const float* mat1d_t, const float* mds, int L, float* ca, cudaStream_t st) {
    float* gout = e->ws.seq_a;     // last layer output; layers 0,1 ping-pong seq_a/seq_b, layer 2 (k&1==0) would alias
    // bigru_stack writes layer k into bufs[k&1] except the last which goes to `out`; use seq_b-safe target:
    gout = e->ws.v_last;           // [L][512] scratch, free after hgru
    TRY(bigru_stack(e, e->w.cgru, 3, LoadConcat2{mat1d_t, 512, 512, mds, 8}, L, gout, st, mat1d_t, 512, 512, mds, 8, 8));
    k_coord_fc<<<cdiv(L, 4), 128, 0, st>>>(gout, e->w.coord_fc, L, ca);
    POST_LAUNCH(e, "k_coord_fc");
    return 0;
}
