// vgru on the tensor cores (network.py:223-224): 2-layer GRU(22 -> 512 -> 512) scanned down the N rows of the
// MSA, batch = the L alignment columns; only the state after the last row is kept.
//
// The recurrence is split into three K=512 GEMM roles that run as a 3-stage wavefront, ONE launch per step:
//   role 0  (time t = s)    gh0 = h0[t-1] W_hh0^T  -> layer-0 cell (input term = column gather of W_ih0) -> h0[t]
//   role 1  (time t = s-1)  gi1 = h0[t]   W_ih1^T + b_ih1                                   -> gi1[t] (fp32, global)
//   role 2  (time t = s-2)  gh1 = h1[t-1] W_hh1^T  -> layer-1 cell with gi1[t]                       -> h1[t]
// Each CTA owns a 128-row x 32-hidden-unit tile: M=128, N=96 (gates r|z|n of its 32 units), K=512, computed as
// TMA -> smem ring -> tcgen05.mma (fp16 hi/lo split of both operands, 3 MMAs, fp32 accumulate in TMEM) ->
// epilogue warps apply the GRU cell and write h as fp32 + fp16 hi/lo (the next step's A operand).
// GRU weights do not tolerate single-pass fp16/tf32, hence the 3-term split.
// The tcgen05 accumulator truncates on every MMA (see conv_tc.cu), so K = 512 is accumulated as FOUR chains of 24 MMAs
// in four TMEM accumulators which the epilogue adds in fp32 with round-to-nearest.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int VT_N = 96;                       // accumulator columns per CTA: 3 gates x 32 units
constexpr int VT_NKB = 8;                      // K = 512 in chunks of 64
constexpr int VT_NACC = 4;                     // accumulation chains (2 k-blocks each), 128 TMEM columns apart
constexpr int VT_A_BYTES = 128 * 64 * 2;       // 16 KB
constexpr int VT_B_BYTES = VT_N * 64 * 2;      // 12 KB
constexpr int VT_STAGE_BYTES = 2 * VT_A_BYTES + 2 * VT_B_BYTES;     // 56 KB
constexpr int VT_STAGES = 4;
constexpr int VT_SMEM = VT_STAGES * VT_STAGE_BYTES + 1024 + 256;
constexpr int VT_THREADS = 576;             // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (4 per TMEM lane quarter, 8 units each)

struct VtMaps {
    CUtensorMap a_hi[4], a_lo[4];              // h buffers: 0,1 = h0 ping/pong, 2,3 = h1 ping/pong
    CUtensorMap b_hi[3], b_lo[3];              // packed weights of the three roles
};

struct VtRole {
    int active;
    int a_sel;                                 // which h buffer is the A operand
    const float* bias;                         // [1536] packed
    const float* gi;                           // role 0: gi0 table [22][1536]; role 2: gi1 [L][1536]; role 1: unused
    const float* h_old;                        // fp32 previous state [L][512] (roles 0, 2)
    float* out_f32;                            // roles 0,2: new state [L][512]; role 1: gi1 [L][1536]
    __half* out_hi;
    __half* out_lo;
};

struct VtParams {
    int L;
    const uint8_t* codes;                      // MSA row consumed by role 0 this step: [L]
    VtRole role[3];
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

__global__ void __launch_bounds__(VT_THREADS, 1) k_vgru_step(const __grid_constant__ VtMaps maps, const VtParams p) {
    const int role_id = blockIdx.z;
    const VtRole& R = p.role[role_id];
    // Programmatic dependent launch: let the next step's grid be scheduled as soon as SMs free up; its prologue
    // (barrier init, TMEM allocation) then overlaps our tail.  Nothing below touches global memory before
    // griddepcontrol.wait, which returns only when the previous step has completed and its writes are visible.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!R.active) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + VT_STAGES * VT_STAGE_BYTES;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (VT_STAGES + s); };
    auto acc_full = [&](int c) { return bar_base + 8u * (2 * VT_STAGES + c); };          // one per accumulation chain
    const uint32_t tmem_slot = acc_full(VT_NACC);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x;              // 32 hidden units
    const int m0 = blockIdx.y * 128;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < VT_STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int c = 0; c < VT_NACC; c++) mbar_init(acc_full(c), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            const CUtensorMap* ah = &maps.a_hi[R.a_sel];
            const CUtensorMap* al = &maps.a_lo[R.a_sel];
            const CUtensorMap* bh = &maps.b_hi[role_id];
            const CUtensorMap* bl = &maps.b_lo[role_id];
            // The weights do not depend on the previous step: fill the B half of every ring stage before the
            // dependency wait (byte count raised without arriving, so the stage stays open until its A half is issued).
            for (int kb = 0; kb < VT_STAGES; kb++) {
                const uint32_t st = base + kb * VT_STAGE_BYTES;
                mbar_add_tx(full(kb), 2 * VT_B_BYTES);
                tma_load_2d(st + 2 * VT_A_BYTES, bh, full(kb), kb * 64, slice * VT_N);
                tma_load_2d(st + 2 * VT_A_BYTES + VT_B_BYTES, bl, full(kb), kb * 64, slice * VT_N);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            int s = 0, ph = 0;
            for (int kb = 0; kb < VT_NKB; kb++) {
                mbar_wait(empty(s), ph ^ 1);
                const uint32_t st = base + s * VT_STAGE_BYTES;
                if (kb < VT_STAGES) {
                    mbar_expect_tx(full(s), 2 * VT_A_BYTES);
                } else {
                    mbar_expect_tx(full(s), VT_STAGE_BYTES);
                    tma_load_2d(st + 2 * VT_A_BYTES, bh, full(s), kb * 64, slice * VT_N);
                    tma_load_2d(st + 2 * VT_A_BYTES + VT_B_BYTES, bl, full(s), kb * 64, slice * VT_N);
                }
                tma_load_2d(st, ah, full(s), kb * 64, m0);
                tma_load_2d(st + VT_A_BYTES, al, full(s), kb * 64, m0);
                if (++s == VT_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, VT_N);
            int s = 0, ph = 0;
            for (int kb = 0; kb < VT_NKB; kb++) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t a_hi = base + s * VT_STAGE_BYTES, a_lo = a_hi + VT_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * VT_A_BYTES, b_lo = b_hi + VT_B_BYTES;
                const uint32_t d = tmem_base + (uint32_t)(kb >> 1) * 128u;        // chain kb/2
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    tc_mma_f16(d, make_smem_desc(a_hi + k * 32), make_smem_desc(b_hi + k * 32), idesc, !((kb & 1) == 0 && k == 0));
                    tc_mma_f16(d, make_smem_desc(a_lo + k * 32), make_smem_desc(b_hi + k * 32), idesc, 1u);
                    tc_mma_f16(d, make_smem_desc(a_hi + k * 32), make_smem_desc(b_lo + k * 32), idesc, 1u);
                }
                tc_commit(empty(s));
                if (kb & 1) tc_commit(acc_full(kb >> 1));          // chain finished: the epilogue may read it while the next one runs
                if (++s == VT_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // 16 epilogue warps: TMEM lane quarter q = warp % 4 (hardware rule), unit quarter uq = 0..3 -> 8 units per thread.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int q = warp & 3;
        const int uq = (warp - 2) >> 2;
        const int row = m0 + q * 32 + lane;
        const bool valid = row < p.L;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + uq * 8;
        const float* bias = R.bias + slice * VT_N + uq * 8;
        // Everything that does not depend on the accumulator is fetched BEFORE waiting on the MMA, so the
        // global-load latency hides behind the TMA/MMA phase.
        float gi[3][8], ho[8];
        if (valid && role_id != 1) {
            const float* gp = (role_id == 0 ? R.gi + (int64_t)p.codes[row] * 1536 : R.gi + (int64_t)row * 1536) + slice * VT_N + uq * 8;
            const float* hp = R.h_old + (int64_t)row * 512 + slice * 32 + uq * 8;
#pragma unroll
            for (int g = 0; g < 3; g++)
#pragma unroll
                for (int v = 0; v < 2; v++) *reinterpret_cast<float4*>(&gi[g][4 * v]) = *reinterpret_cast<const float4*>(gp + g * 32 + 4 * v);
#pragma unroll
            for (int v = 0; v < 2; v++) *reinterpret_cast<float4*>(&ho[4 * v]) = *reinterpret_cast<const float4*>(hp + 4 * v);
        }
        float ar[8], az[8], an[8];
#pragma unroll
        for (int c = 0; c < VT_NACC; c++) {                    // chains are read as they complete, behind the MMAs of the next ones
            mbar_wait(acc_full(c), 0);
            tc_fence_after();
            float pr[8], pz[8], pn[8];
            tmem_ld8(lane_addr + c * 128, pr);
            tmem_ld8(lane_addr + c * 128 + 32, pz);
            tmem_ld8(lane_addr + c * 128 + 64, pn);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                ar[j] = c == 0 ? pr[j] : ar[j] + pr[j];
                az[j] = c == 0 ? pz[j] : az[j] + pz[j];
                an[j] = c == 0 ? pn[j] : an[j] + pn[j];
            }
        }
        if (valid) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                ar[j] += bias[j]; az[j] += bias[32 + j]; an[j] += bias[64 + j];
            }
            if (role_id == 1) {
                float* o = R.out_f32 + (int64_t)row * 1536 + slice * VT_N + uq * 8;
                *reinterpret_cast<float4*>(o) = make_float4(ar[0], ar[1], ar[2], ar[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(ar[4], ar[5], ar[6], ar[7]);
                *reinterpret_cast<float4*>(o + 32) = make_float4(az[0], az[1], az[2], az[3]);
                *reinterpret_cast<float4*>(o + 36) = make_float4(az[4], az[5], az[6], az[7]);
                *reinterpret_cast<float4*>(o + 64) = make_float4(an[0], an[1], an[2], an[3]);
                *reinterpret_cast<float4*>(o + 68) = make_float4(an[4], an[5], an[6], an[7]);
            } else {
                const int64_t hofs = (int64_t)row * 512 + slice * 32 + uq * 8;
                float hn[8];
                __align__(16) __half hh[8], hl[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float rr = sigmoid_acc(gi[0][j] + ar[j]);
                    float zz = sigmoid_acc(gi[1][j] + az[j]);
                    float nn = tanhf(gi[2][j] + rr * an[j]);
                    hn[j] = (1.0f - zz) * nn + zz * ho[j];
                    hh[j] = __float2half_rn(hn[j]);
                    hl[j] = __float2half_rn(hn[j] - __half2float(hh[j]));
                }
                *reinterpret_cast<float4*>(R.out_f32 + hofs) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(R.out_f32 + hofs + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                *reinterpret_cast<uint4*>(R.out_hi + hofs) = *reinterpret_cast<const uint4*>(hh);
                *reinterpret_cast<uint4*>(R.out_lo + hofs) = *reinterpret_cast<const uint4*>(hl);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

struct VtState {
    VtMaps maps;
    bool b_ok = false;
    const void* a_ptr = nullptr;
    int a_L = 0;
    bool attr_set = false;
};

int make_map2d(dmp2_engine* e, CUtensorMap* m, const __half* ptr, int rows, int rows_box) {
    uint64_t dims[2] = {512, (uint64_t)rows};
    uint64_t str[1] = {1024};
    uint32_t box[2] = {64, (uint32_t)rows_box};
    int r = tc::encode_f16_map(m, ptr, 2, dims, str, box);
    if (r != 0) return e->fail(DMP2_ERR_CUDA, "vgru: cuTensorMapEncodeTiled failed with code " + std::to_string(r));
    return 0;
}

}  // namespace

// L alignment columns starting at msa[0] (row stride ld): a column range of a wider alignment when ld > L
int run_vgru_tc(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st, int ld) {
    if (ld <= 0) ld = L;
    if (!e->vt_state) e->vt_state = new VtState();
    VtState* S = (VtState*)e->vt_state;
    const Weights& w = e->w;
    Workspace& ws = e->ws;
    if (!S->attr_set) {
        CUDA_TRY(e, cudaFuncSetAttribute(k_vgru_step, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM));
        S->attr_set = true;
    }
    if (!S->b_ok) {
        for (int r = 0; r < 3; r++) {
            TRY(make_map2d(e, &S->maps.b_hi[r], w.vt_w_hi[r], 1536, VT_N));
            TRY(make_map2d(e, &S->maps.b_lo[r], w.vt_w_lo[r], 1536, VT_N));
        }
        S->b_ok = true;
    }
    const int64_t hsz = (int64_t)L * 512;
    if (S->a_ptr != ws.vt_h16 || S->a_L != L) {
        for (int b = 0; b < 4; b++) {                                   // [buffer][hi|lo][L*512]
            TRY(make_map2d(e, &S->maps.a_hi[b], ws.vt_h16 + (2 * b) * hsz, L, 128));
            TRY(make_map2d(e, &S->maps.a_lo[b], ws.vt_h16 + (2 * b + 1) * hsz, L, 128));
        }
        S->a_ptr = ws.vt_h16;
        S->a_L = L;
    }
    CUDA_TRY(e, cudaMemsetAsync(ws.vg_h, 0, 4 * hsz * sizeof(float), st));
    CUDA_TRY(e, cudaMemsetAsync(ws.vt_h16, 0, 8 * hsz * sizeof(__half), st));
    float* hf[4] = {ws.vg_h, ws.vg_h + hsz, ws.vg_h + 2 * hsz, ws.vg_h + 3 * hsz};     // h0 ping/pong, h1 ping/pong
    auto hi = [&](int b) { return ws.vt_h16 + (2 * b) * hsz; };
    auto lo = [&](int b) { return ws.vt_h16 + (2 * b + 1) * hsz; };
    float* gi1[2] = {ws.vt_gi1, ws.vt_gi1 + (int64_t)L * 1536};
    dim3 grid(16, cdiv(L, 128), 3);
    for (int s = 0; s < N + 2; s++) {
        VtParams p;
        p.L = L;
        p.codes = msa + (int64_t)std::min(s, N - 1) * ld;
        // role 0: time t = s, h0 state before t lives in buffer (s & 1)
        p.role[0] = {s < N, s & 1, w.vt_bias[0], w.vt_gi0, hf[s & 1], hf[(s + 1) & 1], hi((s + 1) & 1), lo((s + 1) & 1)};
        // role 1: time t = s-1, input h0[t] = buffer (s & 1)
        p.role[1] = {s >= 1 && s <= N, s & 1, w.vt_bias[1], nullptr, nullptr, gi1[s & 1], nullptr, nullptr};
        // role 2: time t = s-2, h1 state before t in buffer 2 + (t & 1), gi1[t] written by launch s-1
        const int t2 = s - 2;
        p.role[2] = {s >= 2, 2 + (t2 & 1), w.vt_bias[2], gi1[(s - 1) & 1], hf[2 + (t2 & 1)], hf[2 + ((t2 + 1) & 1)],
                     hi(2 + ((t2 + 1) & 1)), lo(2 + ((t2 + 1) & 1))};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(VT_THREADS);
        cfg.dynamicSmemBytes = VT_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(e, cudaLaunchKernelEx(&cfg, k_vgru_step, S->maps, p));
        POST_LAUNCH(e, "k_vgru_step");
    }
    CUDA_TRY(e, cudaMemcpyAsync(out, hf[2 + (N & 1)], hsz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

void vgru_tc_invalidate(dmp2_engine* e) {             // the state buffers the cached tensor maps describe are being freed
    if (e->vt_state) { ((VtState*)e->vt_state)->a_ptr = nullptr; ((VtState*)e->vt_state)->a_L = 0; }
}

void vgru_tc_destroy(dmp2_engine* e) {
    if (e->vt_state) {
        delete (VtState*)e->vt_state;
        e->vt_state = nullptr;
    }
}
