// vgru as ONE persistent kernel (network.py:223-224): the same three K=512 GEMM roles as vgru_tc.cu
//   role 0  gh0 = h0[t-1] W_hh0^T -> layer-0 cell -> h0[t]
//   role 1  gi1 = h0[t]   W_ih1^T + b_ih1
//   role 2  gh1 = h1[t-1] W_hh1^T -> layer-1 cell with gi1[t] -> h1[t]
// but every CTA (128 rows x 32 hidden units of one role) loops over all N MSA rows itself instead of being
// relaunched per row.  The recurrence is independent across alignment columns, so a step only has to wait for the 16
// slice-CTAs of the same (role, row tile) and for the producing role of the same row tile: the kernel is a dataflow
// pipeline driven by release/acquire counters in global memory (one per role and row tile), with small rings between
// the roles (h0: 4 slots, gi1: 4 slots) for slack and back-pressure.  No kernel boundary, TMEM allocation or barrier
// initialisation per step.  Launched cooperatively (at most 3 row tiles = 144 CTAs per launch, one per SM) so that
// all CTAs of a launch are co-resident, which the spin-waits require.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int VP_N = 96, VP_NKB = 8;
constexpr int VP_A_BYTES = 128 * 64 * 2, VP_B_BYTES = VP_N * 64 * 2;
constexpr int VP_STAGE_BYTES = 2 * VP_A_BYTES + 2 * VP_B_BYTES;
constexpr int VP_STAGES = 4;
constexpr int VP_SMEM = VP_STAGES * VP_STAGE_BYTES + 1024 + 256;
constexpr int VP_THREADS = 320;                // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int VP_D0 = 4, VP_D1 = 4;            // ring depths: h0 (role 0 -> 1), gi1 (role 1 -> 2)
constexpr int VP_MAX_RT = 3;                   // row tiles per launch

struct VpMaps {
    CUtensorMap h0_hi, h0_lo, h1_hi, h1_lo;    // 3-D: {512, L, slots}
    CUtensorMap b_hi[3], b_lo[3];
};
struct VpParams {
    int L, N, rt0;
    const uint8_t* msa;
    const float* gi0;
    const float* bias[3];
    float* h0_f32; __half* h0_hi; __half* h0_lo;          // [VP_D0][L][512]
    float* h1_f32; __half* h1_hi; __half* h1_lo;          // [2][L][512]
    float* gi1;                                            // [VP_D1][L][1536]
    unsigned int* cnt;                                     // [3][gridDim.y], zeroed before the launch
    long long* stamps;                                     // debug (DMP2_VGRU_STAMPS=1): [3 roles][N][4] globaltimer values, else null
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded spin until *p >= target (a protocol bug must trap, never hang the box)
__device__ __forceinline__ void wait_ge(const unsigned int* p, unsigned int target) {
    if (ld_acquire(p) >= target) return;
    long long t0 = clock64();
    while (ld_acquire(p) < target) {
        if (clock64() - t0 > 8000000000LL) {
            printf("k_vgru_persist: counter wait timed out (block %d,%d,%d want %u have %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   target, ld_acquire(p));
            __trap();
        }
    }
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(VP_THREADS, 1) k_vgru_persist(const __grid_constant__ VpMaps maps, const VpParams p) {
    const int role = blockIdx.z, rt = blockIdx.y, slice = blockIdx.x;
    const unsigned int G = gridDim.x;                       // CTAs per (role, row tile) group
    const int m0 = (p.rt0 + rt) * 128;
    const unsigned int* cnt0 = p.cnt + 0 * gridDim.y + rt;
    const unsigned int* cnt1 = p.cnt + 1 * gridDim.y + rt;
    const unsigned int* cnt2 = p.cnt + 2 * gridDim.y + rt;
    unsigned int* my_cnt = p.cnt + role * gridDim.y + rt;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + VP_STAGES * VP_STAGE_BYTES;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (VP_STAGES + s); };
    const uint32_t acc_full = bar_base + 8u * (2 * VP_STAGES);
    const uint32_t acc_empty = acc_full + 8;
    const uint32_t tmem_slot = acc_empty + 8;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < VP_STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);                            // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // when may step t of this role start?  (a group "finished step s" <=> its counter >= G*(s+1))
    auto wait_ready = [&](int t) {
        if (role == 0) {
            wait_ge(cnt0, G * (unsigned)t);                                        // own group wrote h0[t-1]
            if (t >= VP_D0) wait_ge(cnt1, G * (unsigned)(t - VP_D0 + 1));          // role 1 consumed the slot we overwrite
        } else if (role == 1) {
            wait_ge(cnt0, G * (unsigned)(t + 1));                                  // h0[t] is complete
            if (t >= VP_D1) wait_ge(cnt2, G * (unsigned)(t - VP_D1 + 1));          // role 2 consumed the gi1 slot we overwrite
        } else {
            wait_ge(cnt2, G * (unsigned)t);                                        // own group wrote h1[t-1]
            wait_ge(cnt1, G * (unsigned)(t + 1));                                  // gi1[t] is complete
        }
    };

    if (warp == 0) {
        if (lane == 0) {
            const CUtensorMap* ah = role < 2 ? &maps.h0_hi : &maps.h1_hi;
            const CUtensorMap* al = role < 2 ? &maps.h0_lo : &maps.h1_lo;
            const CUtensorMap* bh = &maps.b_hi[role];
            const CUtensorMap* bl = &maps.b_lo[role];
            int s = 0, ph = 0;
            for (int t = 0; t < p.N; t++) {
                wait_ready(t);
                asm volatile("fence.proxy.async;" ::: "memory");           // peers' generic-proxy writes -> our TMA reads
                const int slot = role == 0 ? t % VP_D0 : role == 1 ? (t + 1) % VP_D0 : (t & 1);
                for (int kb = 0; kb < VP_NKB; kb++) {
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_expect_tx(full(s), VP_STAGE_BYTES);
                    const uint32_t st = base + s * VP_STAGE_BYTES;
                    tma_load_3d(st, ah, full(s), kb * 64, m0, slot);
                    tma_load_3d(st + VP_A_BYTES, al, full(s), kb * 64, m0, slot);
                    tma_load_2d(st + 2 * VP_A_BYTES, bh, full(s), kb * 64, slice * VP_N);
                    tma_load_2d(st + 2 * VP_A_BYTES + VP_B_BYTES, bl, full(s), kb * 64, slice * VP_N);
                    if (++s == VP_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, VP_N);
            int s = 0, ph = 0;
            for (int t = 0; t < p.N; t++) {
                if (t > 0) { mbar_wait(acc_empty, (t - 1) & 1); tc_fence_after(); }     // epilogue drained the accumulator
                for (int kb = 0; kb < VP_NKB; kb++) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * VP_STAGE_BYTES, a_lo = a_hi + VP_A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * VP_A_BYTES, b_lo = b_hi + VP_B_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        tc_mma_f16(tmem_base, make_smem_desc(a_hi + k * 32), make_smem_desc(b_hi + k * 32), idesc, !(kb == 0 && k == 0));
                        tc_mma_f16(tmem_base, make_smem_desc(a_lo + k * 32), make_smem_desc(b_hi + k * 32), idesc, 1u);
                        tc_mma_f16(tmem_base, make_smem_desc(a_hi + k * 32), make_smem_desc(b_lo + k * 32), idesc, 1u);
                    }
                    tc_commit(empty(s));
                    if (++s == VP_STAGES) { s = 0; ph ^= 1; }
                }
                tc_commit(acc_full);
            }
        }
    } else {
        const int q = warp & 3, uh = (warp - 2) >> 2;
        const int row = m0 + q * 32 + lane;
        const bool valid = row < p.L;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + uh * 16;
        const float* bias = p.bias[role] + slice * VP_N + uh * 16;
        const int64_t hsz = (int64_t)p.L * 512, gsz = (int64_t)p.L * 1536;
        for (int t = 0; t < p.N; t++) {
            wait_ready(t);
            const bool stamp = p.stamps && slice == 0 && rt == 0 && warp == 2 && lane == 0;
            long long* sp = p.stamps + ((int64_t)role * p.N + t) * 4;
            if (stamp) sp[0] = gtime();
            float gi[3][16], ho[16];
            if (valid && role != 1) {
                const float* gp = (role == 0 ? p.gi0 + (int64_t)p.msa[(int64_t)t * p.L + row] * 1536
                                             : p.gi1 + (t % VP_D1) * gsz + (int64_t)row * 1536) + slice * VP_N + uh * 16;
                const float* hp = (role == 0 ? p.h0_f32 + (t % VP_D0) * hsz : p.h1_f32 + (t & 1) * hsz) + (int64_t)row * 512 + slice * 32 + uh * 16;
#pragma unroll
                for (int g = 0; g < 3; g++)
#pragma unroll
                    for (int v = 0; v < 4; v++) *reinterpret_cast<float4*>(&gi[g][4 * v]) = __ldcg(reinterpret_cast<const float4*>(gp + g * 32 + 4 * v));
#pragma unroll
                for (int v = 0; v < 4; v++) *reinterpret_cast<float4*>(&ho[4 * v]) = __ldcg(reinterpret_cast<const float4*>(hp + 4 * v));   // ring data: L2 only
            }
            mbar_wait(acc_full, t & 1);
            tc_fence_after();
            if (stamp) sp[1] = gtime();
            const int oslot = role == 0 ? (t + 1) % VP_D0 : ((t + 1) & 1);
            float* of32 = role == 0 ? p.h0_f32 + oslot * hsz : p.h1_f32 + oslot * hsz;
            __half* ohi = role == 0 ? p.h0_hi + oslot * hsz : p.h1_hi + oslot * hsz;
            __half* olo = role == 0 ? p.h0_lo + oslot * hsz : p.h1_lo + oslot * hsz;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float ar[8], az[8], an[8];
                tmem_ld8(lane_addr + 8 * i, ar);
                tmem_ld8(lane_addr + 32 + 8 * i, az);
                tmem_ld8(lane_addr + 64 + 8 * i, an);
                tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    ar[j] += bias[8 * i + j]; az[j] += bias[32 + 8 * i + j]; an[j] += bias[64 + 8 * i + j];
                }
                if (role == 1) {
                    float* o = p.gi1 + (t % VP_D1) * gsz + (int64_t)row * 1536 + slice * VP_N + uh * 16 + 8 * i;
                    *reinterpret_cast<float4*>(o) = make_float4(ar[0], ar[1], ar[2], ar[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(ar[4], ar[5], ar[6], ar[7]);
                    *reinterpret_cast<float4*>(o + 32) = make_float4(az[0], az[1], az[2], az[3]);
                    *reinterpret_cast<float4*>(o + 36) = make_float4(az[4], az[5], az[6], az[7]);
                    *reinterpret_cast<float4*>(o + 64) = make_float4(an[0], an[1], an[2], an[3]);
                    *reinterpret_cast<float4*>(o + 68) = make_float4(an[4], an[5], an[6], an[7]);
                    continue;
                }
                const int64_t hofs = (int64_t)row * 512 + slice * 32 + uh * 16 + 8 * i;
                float hn[8];
                __align__(16) __half hh[8], hl[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float rr = sigmoid_acc(gi[0][8 * i + j] + ar[j]);
                    float zz = sigmoid_acc(gi[1][8 * i + j] + az[j]);
                    float nn = tanhf(gi[2][8 * i + j] + rr * an[j]);
                    hn[j] = (1.0f - zz) * nn + zz * ho[8 * i + j];
                    hh[j] = __float2half_rn(hn[j]);
                    hl[j] = __float2half_rn(hn[j] - __half2float(hh[j]));
                }
                *reinterpret_cast<float4*>(of32 + hofs) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(of32 + hofs + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                *reinterpret_cast<uint4*>(ohi + hofs) = *reinterpret_cast<const uint4*>(hh);
                *reinterpret_cast<uint4*>(olo + hofs) = *reinterpret_cast<const uint4*>(hl);
            }
            if (stamp) sp[2] = gtime();
            // the accumulator may be overwritten by the next step's MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            // publish this CTA's part of step t: every writer fences, the epilogue warps meet, one thread releases
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) {
                __threadfence();
                atomicAdd(my_cnt, 1u);
                if (stamp) sp[3] = gtime();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

struct VpState {
    VpMaps maps;
    bool b_ok = false, attr_set = false;
    const void* a_ptr = nullptr;
    int a_L = 0;
};

int map3d(dmp2_engine* e, CUtensorMap* m, const __half* ptr, int L, int slots) {
    uint64_t dims[3] = {512, (uint64_t)L, (uint64_t)slots};
    uint64_t str[2] = {1024, (uint64_t)L * 1024};
    uint32_t box[3] = {64, 128, 1};
    int r = tc::encode_f16_map(m, ptr, 3, dims, str, box);
    if (r != 0) return e->fail(DMP2_ERR_CUDA, "vgru: cuTensorMapEncodeTiled failed with code " + std::to_string(r));
    return 0;
}
int map2d_w(dmp2_engine* e, CUtensorMap* m, const __half* ptr) {
    uint64_t dims[2] = {512, 1536};
    uint64_t str[1] = {1024};
    uint32_t box[2] = {64, VP_N};
    int r = tc::encode_f16_map(m, ptr, 2, dims, str, box);
    if (r != 0) return e->fail(DMP2_ERR_CUDA, "vgru: cuTensorMapEncodeTiled failed with code " + std::to_string(r));
    return 0;
}

}  // namespace

int run_vgru_persist(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st) {
    if (!e->vp_state) e->vp_state = new VpState();
    VpState* S = (VpState*)e->vp_state;
    const Weights& w = e->w;
    Workspace& ws = e->ws;
    if (!S->attr_set) {
        CUDA_TRY(e, cudaFuncSetAttribute(k_vgru_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, VP_SMEM));
        S->attr_set = true;
    }
    if (!S->b_ok) {
        for (int r = 0; r < 3; r++) {
            TRY(map2d_w(e, &S->maps.b_hi[r], w.vt_w_hi[r]));
            TRY(map2d_w(e, &S->maps.b_lo[r], w.vt_w_lo[r]));
        }
        S->b_ok = true;
    }
    const int64_t hsz = (int64_t)L * 512;
    // workspace layout: vp_h16 = [h0_hi D0][h0_lo D0][h1_hi 2][h1_lo 2] (x hsz halves); vp_f32 = [h0 D0][h1 2] (x hsz floats)
    __half* h0_hi = ws.vp_h16;
    __half* h0_lo = h0_hi + VP_D0 * hsz;
    __half* h1_hi = h0_lo + VP_D0 * hsz;
    __half* h1_lo = h1_hi + 2 * hsz;
    float* h0_f32 = ws.vp_f32;
    float* h1_f32 = h0_f32 + VP_D0 * hsz;
    if (S->a_ptr != ws.vp_h16 || S->a_L != L) {
        TRY(map3d(e, &S->maps.h0_hi, h0_hi, L, VP_D0));
        TRY(map3d(e, &S->maps.h0_lo, h0_lo, L, VP_D0));
        TRY(map3d(e, &S->maps.h1_hi, h1_hi, L, 2));
        TRY(map3d(e, &S->maps.h1_lo, h1_lo, L, 2));
        S->a_ptr = ws.vp_h16;
        S->a_L = L;
    }
    CUDA_TRY(e, cudaMemsetAsync(ws.vp_h16, 0, (2 * VP_D0 + 4) * hsz * sizeof(__half), st));
    CUDA_TRY(e, cudaMemsetAsync(ws.vp_f32, 0, (VP_D0 + 2) * hsz * sizeof(float), st));
    const int nrt = cdiv(L, 128);
    CUDA_TRY(e, cudaMemsetAsync(ws.vp_cnt, 0, 3 * (size_t)nrt * sizeof(unsigned int), st));
    for (int rt0 = 0; rt0 < nrt; rt0 += VP_MAX_RT) {
        const int g = std::min(VP_MAX_RT, nrt - rt0);
        VpParams p;
        p.L = L; p.N = N; p.rt0 = rt0; p.msa = msa; p.gi0 = w.vt_gi0;
        for (int r = 0; r < 3; r++) p.bias[r] = w.vt_bias[r];
        p.h0_f32 = h0_f32; p.h0_hi = h0_hi; p.h0_lo = h0_lo;
        p.h1_f32 = h1_f32; p.h1_hi = h1_hi; p.h1_lo = h1_lo;
        p.gi1 = ws.vp_gi1;
        p.cnt = ws.vp_cnt + 3 * rt0;
        p.stamps = nullptr;
        long long* d_stamps = nullptr;
        if (rt0 == 0 && getenv("DMP2_VGRU_STAMPS")) {
            CUDA_TRY(e, cudaMalloc(&d_stamps, (size_t)3 * N * 4 * sizeof(long long)));
            CUDA_TRY(e, cudaMemsetAsync(d_stamps, 0, (size_t)3 * N * 4 * sizeof(long long), st));
            p.stamps = d_stamps;
        }
        void* args[2] = {(void*)&S->maps, (void*)&p};
        CUDA_TRY(e, cudaLaunchCooperativeKernel((const void*)k_vgru_persist, dim3(16, g, 3), dim3(VP_THREADS), args, VP_SMEM, st));
        POST_LAUNCH(e, "k_vgru_persist");
        if (d_stamps) {      // debug: where does a step's time go?  ready -> accumulator full -> cell done -> published -> next ready
            CUDA_TRY(e, cudaStreamSynchronize(st));
            std::vector<long long> h((size_t)3 * N * 4);
            CUDA_TRY(e, cudaMemcpy(h.data(), d_stamps, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            cudaFree(d_stamps);
            for (int r = 0; r < 3; r++) {
                double a = 0, b = 0, c = 0, d = 0;
                const int t0 = N / 4, t1 = N - 1;
                for (int t = t0; t < t1; t++) {
                    const long long* s = &h[((size_t)r * N + t) * 4];
                    a += s[1] - s[0]; b += s[2] - s[1]; c += s[3] - s[2]; d += s[4] - s[3];
                }
                const double n = t1 - t0;
                fprintf(stderr, "vgru persist role %d (L=%d): ready->acc_full %.2f us, cell %.2f us, publish %.2f us, published->next ready %.2f us\n", r, L,
                        a / n * 1e-3, b / n * 1e-3, c / n * 1e-3, d / n * 1e-3);
            }
        }
    }
    CUDA_TRY(e, cudaMemcpyAsync(out, h1_f32 + (N & 1) * hsz, hsz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

void vgru_persist_destroy(dmp2_engine* e) {
    if (e->vp_state) {
        delete (VpState*)e->vp_state;
        e->vp_state = nullptr;
    }
}
