// Tensor-core 5x5 convolution of the ResNet blocks (network.py:26 inside ResNet_Block, + maxout :30-31) as an
// implicit GEMM on the 5th-generation tensor cores:  M = L*L pixels, N = 512 output channels, K = 25 taps x 128.
//
//   TMA (cp.async.bulk.tensor, 128B swizzle, out-of-bounds zero fill = the pad-2 border)
//     -> shared-memory rings (A: shifted 8x16-pixel patch; B: 256-cout weight pieces, multicast across a cluster)
//     -> tcgen05.mma, M=128 N=256, fp32 accumulators in TMEM (128 lanes x 512 columns = all of it)
//     -> epilogue warps: tcgen05.ld -> + bias -> max over 4 consecutive couts -> NHWC fp32 store.
//
// Precision modes (x = x_hi + x_lo, w = w_hi + w_lo are fp16 hi/lo splits of the fp32 operands):
//   F16    x_hi*w_hi                                   1 MMA  per algorithmic MAC   (fast, ~1e-3 A drift)
//   F16X3  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (all fp16) 3 MMAs                       (fp32-equivalent)
//   F16F8  x_hi*w_hi in fp16 + the two correction terms in FP8 (e4m3 activations x e5m2 weights, kind::f8f6f4,
//          twice the fp16 rate): the corrections are ~2^-11 of the main term, so 2-3 mantissa bits suffice.
//          Fixed power-of-two pre-scales keep the fp8 operands in range and cancel in the product:
//          (x_lo*2^8)(w*2^-8) and (x_hi*2^-4)(w_lo*2^4).  2 MMA-equivalents per MAC.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (a warp may only touch TMEM lanes 32*(warp%4)..+31).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;               // pixels per CTA (8 rows x 16 columns)
constexpr int TILE_H = 8, TILE_W = 16;
constexpr int KCHUNK = 64;                // fp16 channels per 128-byte swizzle row
constexpr int A_BYTES = TILE_M * 128;     // 16 KB: 128 pixels x 128 bytes (64 fp16 or 128 fp8 channels)
constexpr int B_BYTES = 256 * 128;        // 32 KB: one B piece = 256 couts x 128 bytes
constexpr int BQ_ROWS = 64;               // TMA granule of a B piece: a quarter (multicast unit for clusters up to 4)
constexpr int BQ_BYTES = BQ_ROWS * 128;
constexpr int NUM_B_SLOTS = 5;
constexpr int NUM_THREADS = 192;
enum { M_F16 = 0, M_F16X3 = 1, M_F16F8 = 2 };

template <int MODE>
struct Cfg {
    static constexpr int A_STAGE_BYTES = MODE == M_F16 ? A_BYTES : 2 * A_BYTES;
    static constexpr int NUM_A_STAGES = MODE == M_F16 ? 4 : 2;
    static constexpr int PIECES = MODE == M_F16 ? 2 : 4;      // B pieces per k-block
    static constexpr int SMEM_BYTES = NUM_A_STAGES * A_STAGE_BYTES + NUM_B_SLOTS * B_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct ConvMaps {
    CUtensorMap a_hi, a_lo;              // fp16 activations  [L][L][128]
    CUtensorMap a8_lo, a8_hi;            // e4m3: x_lo * 2^8, x_hi * 2^-4
    CUtensorMap b_hi, b_lo;              // fp16 weights [512][3200]
    CUtensorMap b8_w, b8_lo;             // e5m2: w * 2^-8, w_lo * 2^4
};

struct TcParams {
    int gemm;            // 0 = conv (3-D activation map, taps), 1 = plain GEMM test (rows x K)
    int L;               // conv: image width (pixels per row)
    int H;               // conv: output rows (== L for a whole image, the strip height for a halo-sharded fold)
    int y_off;           // conv: row of the activation map that holds output row 0 (0, or 2 when the map starts with halo rows)
    int tiles_x;         // conv: tiles per image row
    int num_kb;          // k-blocks: conv 50, gemm K/64
    int M;               // gemm: rows
    float* out;          // conv: raw [L*L][128]; gemm: C [M][512]
    const float* bias;   // conv: [512]
    // fused InstanceNorm statistics of the output (network.py:32), nullptr = off: every CTA leaves the fp64 sums of
    // its tile in stat_part, the last CTA of a group folds the group, the last group finishes (same deterministic
    // two-level fold as k_in_stats, which this replaces after a conv)
    double* stat_part;           // [grid + groups][256]
    unsigned int* ticket;        // [1 + groups], zero between launches
    float* norm;                 // [mean 128 | gamma * rstd 128]
    const float* gamma;
    double* totals;              // halo-sharded: [sum 128 | sumsq 128] of this launch instead of norm
    double npix;                 // pixels the statistics are over (H * L)
};
constexpr int STAT_GROUP = 32;   // CTAs per first-level fold
constexpr int STAT_LD = 132;     // floats per pixel row of the staged tile (conflict-free 16-byte stores)

using namespace tc;

__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f8f6f4 instruction descriptor: A = e4m3 (0), B = e5m2 (1), fp32 accumulate, K-major, M x N
__device__ __forceinline__ uint32_t make_idesc_f8(int m, int n) {
    return (1u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- the kernel --------------------------------------------------------------------------------------
// CL = CTAs per cluster sharing the weight (B) stream: each CTA loads 1/CL of every B piece and multicasts it
// to all CTAs of the cluster, so the L2 -> SM weight traffic (the dominant operand stream) drops by CL.
template <int MODE, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv5_tc(const __grid_constant__ ConvMaps maps, const TcParams p) {
    using C = Cfg<MODE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 128B swizzle needs 1024-byte alignment
    const uint32_t a_base = base;
    const uint32_t b_base = base + C::NUM_A_STAGES * C::A_STAGE_BYTES;
    const uint32_t bar_base = b_base + NUM_B_SLOTS * B_BYTES;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (C::NUM_A_STAGES + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * C::NUM_A_STAGES + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * C::NUM_A_STAGES + NUM_B_SLOTS + s); };
    const uint32_t acc_full = bar_base + 8u * (2 * C::NUM_A_STAGES + 2 * NUM_B_SLOTS);
    const uint32_t tmem_slot = acc_full + 8;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

    // tile coordinates (tiles past the end of the image are all out-of-bounds: zero loads, no stores)
    int x0 = 0, y0 = 0, m0 = 0;
    if (p.gemm) m0 = blockIdx.x * TILE_M;
    else { y0 = (blockIdx.x / p.tiles_x) * TILE_H; x0 = (blockIdx.x % p.tiles_x) * TILE_W; }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::NUM_A_STAGES; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < NUM_B_SLOTS; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), CL); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                  // peers' barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int sa = 0, pa = 0, sb = 0, pb = 0;
            for (int kb = 0; kb < p.num_kb; kb++) {
                // ---- A stage
                const int tap = kb >> 1, sub = kb & 1;
                int c1, c2;
                if (p.gemm) { c1 = m0; c2 = 0; }
                else {
                    const int dy = tap / 5, dx = tap - dy * 5;
                    c1 = x0 + dx - 2; c2 = y0 + dy - 2 + p.y_off;
                }
                mbar_wait(a_empty(sa), pa ^ 1);
                mbar_expect_tx(a_full(sa), C::A_STAGE_BYTES);
                const uint32_t ast = a_base + sa * C::A_STAGE_BYTES;
                if (MODE == M_F16) {
                    tma_load_3d(ast, &maps.a_hi, a_full(sa), p.gemm ? kb * KCHUNK : sub * KCHUNK, c1, c2);
                } else if (MODE == M_F16X3) {
                    const int c0 = p.gemm ? kb * KCHUNK : sub * KCHUNK;
                    tma_load_3d(ast, &maps.a_hi, a_full(sa), c0, c1, c2);
                    tma_load_3d(ast + A_BYTES, &maps.a_lo, a_full(sa), c0, c1, c2);
                } else if (sub == 0) {                         // F16F8: both fp16 channel chunks of x_hi
                    tma_load_3d(ast, &maps.a_hi, a_full(sa), 0, c1, c2);
                    tma_load_3d(ast + A_BYTES, &maps.a_hi, a_full(sa), KCHUNK, c1, c2);
                } else {                                       // F16F8: the two fp8 correction operands (128 channels each)
                    tma_load_3d(ast, &maps.a8_lo, a_full(sa), 0, c1, c2);
                    tma_load_3d(ast + A_BYTES, &maps.a8_hi, a_full(sa), 0, c1, c2);
                }
                if (++sa == C::NUM_A_STAGES) { sa = 0; pa ^= 1; }
                // ---- B pieces
                for (int piece = 0; piece < C::PIECES; piece++) {
                    const CUtensorMap* bm;
                    int k0;                                    // element offset along K of the weight row
                    if (MODE == M_F16F8) {
                        if (sub == 0) { bm = &maps.b_hi; k0 = tap * 128 + (piece >> 1) * KCHUNK; }
                        else { bm = (piece >> 1) ? &maps.b8_lo : &maps.b8_w; k0 = tap * 128; }
                    } else {
                        bm = piece < 2 ? &maps.b_hi : &maps.b_lo;
                        k0 = kb * KCHUNK;                      // weights are [512][K] with k = tap*128 + c = kb*64 + ...
                    }
                    const int n0 = (piece & 1) * 256;
                    mbar_wait(b_empty(sb), pb ^ 1);
                    mbar_expect_tx(b_full(sb), B_BYTES);
                    const uint32_t bdst = b_base + sb * B_BYTES;
#pragma unroll
                    for (int i = 0; i < 4 / CL; i++) {         // my quarters of the piece, delivered to every CTA of the cluster
                        const int qd = crank * (4 / CL) + i;
                        if (CL == 1) tma_load_2d(bdst + qd * BQ_BYTES, bm, b_full(sb), k0, n0 + qd * BQ_ROWS);
                        else tma_load_2d_mc(bdst + qd * BQ_BYTES, bm, b_full(sb), k0, n0 + qd * BQ_ROWS, MC_MASK);
                    }
                    if (++sb == NUM_B_SLOTS) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, 256);
            const uint32_t idesc8 = make_idesc_f8(128, 256);
            int sa = 0, pa = 0, sb = 0, pb = 0;
            for (int kb = 0; kb < p.num_kb; kb++) {
                const int sub = kb & 1;
                mbar_wait(a_full(sa), pa);
                const uint32_t a0 = a_base + sa * C::A_STAGE_BYTES;
                const uint32_t a1 = a0 + A_BYTES;
                for (int piece = 0; piece < C::PIECES; piece++) {
                    mbar_wait(b_full(sb), pb);
                    tc_fence_after();
                    const uint32_t b_addr = b_base + sb * B_BYTES;
                    const uint32_t d = tmem_base + (uint32_t)(piece & 1) * 256u;
                    if (MODE == M_F16F8) {
                        const uint32_t a = (piece >> 1) ? a1 : a0;
                        if (sub == 0) {
                            const bool first = (kb == 0) && (piece < 2);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc_mma_f16(d, make_smem_desc(a + k * 32), make_smem_desc(b_addr + k * 32), idesc, !(first && k == 0));
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; k++)        // K = 32 fp8 elements = 32 bytes per MMA
                                tc_mma_f8(d, make_smem_desc(a + k * 32), make_smem_desc(b_addr + k * 32), idesc8, 1u);
                        }
                    } else {
                        const bool first = (kb == 0) && (piece < 2);          // first touch of this accumulator half
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            tc_mma_f16(d, make_smem_desc(a0 + k * 32), make_smem_desc(b_addr + k * 32), idesc, !(first && k == 0));
                        if (MODE == M_F16X3 && piece < 2) {                    // lo(A) x hi(B)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc_mma_f16(d, make_smem_desc(a1 + k * 32), make_smem_desc(b_addr + k * 32), idesc, 1u);
                        }
                    }
                    if (CL == 1) tc_commit(b_empty(sb));
                    else tc_commit_mc(b_empty(sb), MC_MASK);   // the slot is free only when every CTA of the cluster is done with it
                    if (++sb == NUM_B_SLOTS) { sb = 0; pb ^= 1; }
                }
                tc_commit(a_empty(sa));
                if (++sa == C::NUM_A_STAGES) { sa = 0; pa ^= 1; }
            }
            tc_commit(acc_full);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                 // accumulator row = pixel within the tile
        mbar_wait(acc_full, 0);
        tc_fence_after();
        bool valid;
        int64_t row;
        if (p.gemm) { valid = (m0 + r) < p.M; row = m0 + r; }
        else {
            int y = y0 + (r >> 4), x = x0 + (r & 15);
            valid = (y < p.H) && (x < p.L);
            row = (int64_t)y * p.L + x;
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int ch = 0; ch < 16; ch++) {
            uint32_t v[32];
            tmem_ld32(lane_addr + ch * 32, v);
            if (!valid) continue;
            if (p.gemm) {
                float4* dst = reinterpret_cast<float4*>(p.out + row * 512 + ch * 32);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                         __uint_as_float(v[4 * i + 3]));
            } else {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ch * 32 + 4 * i));
                    o[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]) + b.x, __uint_as_float(v[4 * i + 1]) + b.y),
                                 fmaxf(__uint_as_float(v[4 * i + 2]) + b.z, __uint_as_float(v[4 * i + 3]) + b.w));
                }
                float4* dst = reinterpret_cast<float4*>(p.out + row * 128 + ch * 8);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                if (p.stat_part) {                   // stage the tile for the per-channel sums (operand rings are idle now)
                    float4* sd4 = reinterpret_cast<float4*>(smem_gen + (size_t)(r * STAT_LD + ch * 8) * 4);
                    sd4[0] = make_float4(o[0], o[1], o[2], o[3]); sd4[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        if (!p.gemm && p.stat_part) {
            float* tile = reinterpret_cast<float*>(smem_gen);
            if (!valid) {                            // pixels outside the image count as zeros
                for (int c = 0; c < 128; c += 4) *reinterpret_cast<float4*>(tile + r * STAT_LD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = r;                         // this thread now owns channel t
            double s4[4] = {0, 0, 0, 0}, q4[4] = {0, 0, 0, 0};        // four independent chains, combined in a fixed order
#pragma unroll 4
            for (int px = 0; px < TILE_M; px += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double v = (double)tile[(px + u) * STAT_LD + t];
                    s4[u] += v;
                    q4[u] += v * v;
                }
            }
            const double s = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            double* part = p.stat_part;
            part[(int64_t)blockIdx.x * 256 + t] = s;
            part[(int64_t)blockIdx.x * 256 + 128 + t] = ss;
            // deterministic two-level fold: last CTA of a group folds the group, last group folds the group sums
            const unsigned G = STAT_GROUP, grp_id = blockIdx.x / G, ngrp = (gridDim.x + G - 1) / G;
            const unsigned gfirst = grp_id * G, gsize = min(G, gridDim.x - gfirst);
            double* part2 = part + (int64_t)gridDim.x * 256;
            int* stage = reinterpret_cast<int*>(tile + TILE_M * STAT_LD);      // one word past the staged tile
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == 0) *stage = (atomicAdd(p.ticket + 1 + grp_id, 1u) == gsize - 1) ? 1 : 0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (*stage != 0) {
                __threadfence();
                double a = 0.0, b = 0.0;
#pragma unroll 8
                for (unsigned q2 = 0; q2 < gsize; q2++) {
                    a += __ldcg(part + (int64_t)(gfirst + q2) * 256 + t);
                    b += __ldcg(part + (int64_t)(gfirst + q2) * 256 + 128 + t);
                }
                part2[(int64_t)grp_id * 256 + t] = a;
                part2[(int64_t)grp_id * 256 + 128 + t] = b;
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (t == 0) {
                    p.ticket[1 + grp_id] = 0;
                    *stage = (atomicAdd(p.ticket, 1u) == ngrp - 1) ? 2 : 1;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (*stage == 2) {
                    __threadfence();
                    a = 0.0; b = 0.0;
#pragma unroll 8
                    for (unsigned q2 = 0; q2 < ngrp; q2++) {
                        a += __ldcg(part2 + (int64_t)q2 * 256 + t);
                        b += __ldcg(part2 + (int64_t)q2 * 256 + 128 + t);
                    }
                    if (p.totals) { p.totals[t] = a; p.totals[128 + t] = b; }
                    else {
                        const double mean = a / p.npix;
                        double var = b / p.npix - mean * mean;
                        if (var < 0) var = 0;
                        p.norm[t] = (float)mean;
                        p.norm[128 + t] = (float)((double)p.gamma[t] / sqrt(var + 1e-5));
                    }
                    if (t == 0) *p.ticket = 0;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                  // no CTA may exit while a peer can still arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- CTA-pair variant (cta_group::2) ---------------------------------------------------------------------
// The two CTAs of a cluster own two adjacent 128-pixel tiles and behave as one 256 x 512 tile: every MMA is
// M=256 N=256, issued by the leader CTA, reading the A rows of both CTAs and HALF of the B piece from each CTA's
// shared memory.  Per CTA the weight stream through shared memory is halved (16 KB per piece instead of 32 KB),
// which both halves the SM ingest bandwidth and doubles the number of B slots (pipeline depth) in the same 160 KB.
constexpr int B2_BYTES = 128 * 128;       // per-CTA half of a B piece: 128 couts x 128 bytes
constexpr int NUM_B2_SLOTS = 10;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
k_conv5_tc2(const __grid_constant__ ConvMaps maps, const TcParams p) {
    using C = Cfg<MODE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base;
    const uint32_t b_base = base + C::NUM_A_STAGES * C::A_STAGE_BYTES;
    const uint32_t bar_base = b_base + NUM_B2_SLOTS * B2_BYTES;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (C::NUM_A_STAGES + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * C::NUM_A_STAGES + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * C::NUM_A_STAGES + NUM_B2_SLOTS + s); };
    const uint32_t acc_full = bar_base + 8u * (2 * C::NUM_A_STAGES + 2 * NUM_B2_SLOTS);
    const uint32_t tmem_slot = acc_full + 8;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    const int y0 = (blockIdx.x / p.tiles_x) * TILE_H, x0 = (blockIdx.x % p.tiles_x) * TILE_W;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::NUM_A_STAGES; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < NUM_B2_SLOTS; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {                                 // both CTAs, same logical warp, same destination offset
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; full barriers live in the leader) =====================
        if (lane == 0) {
            int sa = 0, pa = 0, sb = 0, pb = 0;
            for (int kb = 0; kb < p.num_kb; kb++) {
                const int tap = kb >> 1, sub = kb & 1;
                const int dy = tap / 5, dx = tap - dy * 5;
                const int c1 = x0 + dx - 2, c2 = y0 + dy - 2 + p.y_off;
                mbar_wait(a_empty(sa), pa ^ 1);
                const uint32_t afl = mapa_u32(a_full(sa), 0);
                if (leader) mbar_expect_tx(a_full(sa), 2 * C::A_STAGE_BYTES);
                const uint32_t ast = a_base + sa * C::A_STAGE_BYTES;
                if (MODE == M_F16) {
                    tma2_load_3d(ast, &maps.a_hi, afl, sub * KCHUNK, c1, c2);
                } else if (MODE == M_F16X3) {
                    tma2_load_3d(ast, &maps.a_hi, afl, sub * KCHUNK, c1, c2);
                    tma2_load_3d(ast + A_BYTES, &maps.a_lo, afl, sub * KCHUNK, c1, c2);
                } else if (sub == 0) {
                    tma2_load_3d(ast, &maps.a_hi, afl, 0, c1, c2);
                    tma2_load_3d(ast + A_BYTES, &maps.a_hi, afl, KCHUNK, c1, c2);
                } else {
                    tma2_load_3d(ast, &maps.a8_lo, afl, 0, c1, c2);
                    tma2_load_3d(ast + A_BYTES, &maps.a8_hi, afl, 0, c1, c2);
                }
                if (++sa == C::NUM_A_STAGES) { sa = 0; pa ^= 1; }
                for (int piece = 0; piece < C::PIECES; piece++) {
                    const CUtensorMap* bm;
                    int k0;
                    if (MODE == M_F16F8) {
                        if (sub == 0) { bm = &maps.b_hi; k0 = tap * 128 + (piece >> 1) * KCHUNK; }
                        else { bm = (piece >> 1) ? &maps.b8_lo : &maps.b8_w; k0 = tap * 128; }
                    } else {
                        bm = piece < 2 ? &maps.b_hi : &maps.b_lo;
                        k0 = kb * KCHUNK;
                    }
                    const int n0 = (piece & 1) * 256 + (int)crank * 128;      // my half of the 256 couts of this piece
                    mbar_wait(b_empty(sb), pb ^ 1);
                    const uint32_t bfl = mapa_u32(b_full(sb), 0);
                    if (leader) mbar_expect_tx(b_full(sb), 2 * B2_BYTES);
                    const uint32_t bdst = b_base + sb * B2_BYTES;
                    tma2_load_2d(bdst, bm, bfl, k0, n0);
                    tma2_load_2d(bdst + BQ_BYTES, bm, bfl, k0, n0 + BQ_ROWS);
                    if (++sb == NUM_B2_SLOTS) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc(256, 256);
            const uint32_t idesc8 = make_idesc_f8(256, 256);
            int sa = 0, pa = 0, sb = 0, pb = 0;
            for (int kb = 0; kb < p.num_kb; kb++) {
                const int sub = kb & 1;
                mbar_wait(a_full(sa), pa);
                const uint32_t a0 = a_base + sa * C::A_STAGE_BYTES;
                const uint32_t a1 = a0 + A_BYTES;
                for (int piece = 0; piece < C::PIECES; piece++) {
                    mbar_wait(b_full(sb), pb);
                    tc_fence_after();
                    const uint32_t b_addr = b_base + sb * B2_BYTES;
                    const uint32_t d = tmem_base + (uint32_t)(piece & 1) * 256u;
                    if (MODE == M_F16F8) {
                        const uint32_t a = (piece >> 1) ? a1 : a0;
                        if (sub == 0) {
                            const bool first = (kb == 0) && (piece < 2);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc2_mma_f16(d, make_smem_desc(a + k * 32), make_smem_desc(b_addr + k * 32), idesc, !(first && k == 0));
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc2_mma_f8(d, make_smem_desc(a + k * 32), make_smem_desc(b_addr + k * 32), idesc8, 1u);
                        }
                    } else {
                        const bool first = (kb == 0) && (piece < 2);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            tc2_mma_f16(d, make_smem_desc(a0 + k * 32), make_smem_desc(b_addr + k * 32), idesc, !(first && k == 0));
                        if (MODE == M_F16X3 && piece < 2) {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc2_mma_f16(d, make_smem_desc(a1 + k * 32), make_smem_desc(b_addr + k * 32), idesc, 1u);
                        }
                    }
                    tc2_commit_mc(b_empty(sb), 3);         // frees the slot in both CTAs
                    if (++sb == NUM_B2_SLOTS) { sb = 0; pb ^= 1; }
                }
                tc2_commit_mc(a_empty(sa), 3);
                if (++sa == C::NUM_A_STAGES) { sa = 0; pa ^= 1; }
            }
            tc2_commit_mc(acc_full, 3);
        }
    } else {
        // ===================== epilogue (warps 2..5, each CTA drains its own 128 accumulator rows) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int y = y0 + (r >> 4), x = x0 + (r & 15);
        const bool valid = (y < p.H) && (x < p.L);
        const int64_t row = (int64_t)y * p.L + x;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int ch = 0; ch < 16; ch++) {
            uint32_t v[32];
            tmem_ld32(lane_addr + ch * 32, v);
            if (!valid) continue;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ch * 32 + 4 * i));
                o[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]) + b.x, __uint_as_float(v[4 * i + 1]) + b.y),
                             fmaxf(__uint_as_float(v[4 * i + 2]) + b.z, __uint_as_float(v[4 * i + 3]) + b.w));
            }
            float4* dst = reinterpret_cast<float4*>(p.out + row * 128 + ch * 8);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // the leader's MMAs read the peer's shared memory until the very end
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- host side: tensor maps -----------------------------------------------------------------------------
struct TcState {
    CUtensorMap wmap[DMP2_NBLOCKS][4];               // b_hi, b_lo, b8_w, b8_lo
    bool wmap_ok[DMP2_NBLOCKS] = {false};
    CUtensorMap amap[4];                             // a_hi, a_lo, a8_lo, a8_hi
    const void* amap_ptr = nullptr;
    int amap_L = 0, amap_rows = 0;
    bool attr_set = false;
};

template <int MODE, int CL>
int set_attr(dmp2_engine* e) {
    CUDA_TRY(e, cudaFuncSetAttribute(k_conv5_tc<MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE>::SMEM_BYTES));
    return 0;
}

int get_state(dmp2_engine* e, TcState** out) {
    if (!e->tc_state) {
        if (!tc::get_encode_fn()) return e->fail(DMP2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
        e->tc_state = new TcState();
    }
    TcState* s = (TcState*)e->tc_state;
    if (!s->attr_set) {
        TRY((set_attr<M_F16, 1>(e))); TRY((set_attr<M_F16, 2>(e))); TRY((set_attr<M_F16, 4>(e)));
        TRY((set_attr<M_F16X3, 1>(e))); TRY((set_attr<M_F16X3, 2>(e))); TRY((set_attr<M_F16X3, 4>(e)));
        TRY((set_attr<M_F16F8, 1>(e))); TRY((set_attr<M_F16F8, 2>(e))); TRY((set_attr<M_F16F8, 4>(e)));
        CUDA_TRY(e, cudaFuncSetAttribute(k_conv5_tc2<M_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<M_F16>::SMEM_BYTES));
        CUDA_TRY(e, cudaFuncSetAttribute(k_conv5_tc2<M_F16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<M_F16X3>::SMEM_BYTES));
        CUDA_TRY(e, cudaFuncSetAttribute(k_conv5_tc2<M_F16F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<M_F16F8>::SMEM_BYTES));
        s->attr_set = true;
    }
    *out = s;
    return 0;
}

int encode(dmp2_engine* e, CUtensorMap* map, const void* ptr, int elem_bytes, int rank, const uint64_t* dims,
           const uint64_t* strides_bytes, const uint32_t* box) {
    int r = tc::encode_map(map, ptr, elem_bytes, rank, dims, strides_bytes, box);
    if (r != 0) return e->fail(DMP2_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
    return 0;
}

// weights [512][K] (K contiguous), box = one 128-byte swizzle row x 64 couts
int weight_map(dmp2_engine* e, CUtensorMap* map, const void* w, int K, int elem_bytes) {
    uint64_t dims[2] = {(uint64_t)K, 512};
    uint64_t str[1] = {(uint64_t)K * elem_bytes};
    uint32_t box[2] = {(uint32_t)(128 / elem_bytes), BQ_ROWS};
    return encode(e, map, w, elem_bytes, 2, dims, str, box);
}

template <int MODE, int CL>
int launch(dmp2_engine* e, const ConvMaps& maps, const TcParams& p, int grid, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((grid + CL - 1) / CL * CL);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<MODE>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(e, cudaLaunchKernelEx(&cfg, k_conv5_tc<MODE, CL>, maps, p));
    POST_LAUNCH(e, "k_conv5_tc");
    return 0;
}

template <int MODE>
int launch_pair(dmp2_engine* e, const ConvMaps& maps, const TcParams& p, int grid, cudaStream_t st) {
    k_conv5_tc2<MODE><<<(grid + 1) / 2 * 2, NUM_THREADS, Cfg<MODE>::SMEM_BYTES, st>>>(maps, p);
    POST_LAUNCH(e, "k_conv5_tc2");
    return 0;
}

template <int MODE>
int launch_cl(dmp2_engine* e, int cl, const ConvMaps& maps, const TcParams& p, int grid, cudaStream_t st) {
    if (cl == 0) return launch_pair<MODE>(e, maps, p, grid, st);
    if (cl == 4) return launch<MODE, 4>(e, maps, p, grid, st);
    if (cl == 2) return launch<MODE, 2>(e, maps, p, grid, st);
    return launch<MODE, 1>(e, maps, p, grid, st);
}

}  // namespace

// xh/xl/x8lo/x8hi: activation maps of map_rows x L pixels; output row y reads map rows y + y_off - 2 .. y + y_off + 2
// (rows outside the map read as zero), H output rows are written to raw.  Whole image: map_rows = H = L, y_off = 0.
int run_conv_tc(dmp2_engine* e, int blk, const __half* xh, const __half* xl, const uint8_t* x8lo, const uint8_t* x8hi, int L,
                int H, int y_off, int map_rows, float* raw, int mode, cudaStream_t st, bool fuse_stats) {
    TcState* s;
    TRY(get_state(e, &s));
    const ResBlockW& bw = e->w.blk[blk];
    if (!s->wmap_ok[blk]) {
        TRY(weight_map(e, &s->wmap[blk][0], bw.w_hi, 3200, 2));
        TRY(weight_map(e, &s->wmap[blk][1], bw.w_lo, 3200, 2));
        TRY(weight_map(e, &s->wmap[blk][2], bw.w8_w, 3200, 1));
        TRY(weight_map(e, &s->wmap[blk][3], bw.w8_lo, 3200, 1));
        s->wmap_ok[blk] = true;
    }
    if (s->amap_ptr != xh || s->amap_L != L || s->amap_rows != map_rows) {
        uint64_t dims[3] = {128, (uint64_t)L, (uint64_t)map_rows};
        uint64_t str16[2] = {256, (uint64_t)L * 256}, str8[2] = {128, (uint64_t)L * 128};
        uint32_t box16[3] = {KCHUNK, TILE_W, TILE_H}, box8[3] = {128, TILE_W, TILE_H};
        TRY(encode(e, &s->amap[0], xh, 2, 3, dims, str16, box16));
        TRY(encode(e, &s->amap[1], xl, 2, 3, dims, str16, box16));
        TRY(encode(e, &s->amap[2], x8lo, 1, 3, dims, str8, box8));
        TRY(encode(e, &s->amap[3], x8hi, 1, 3, dims, str8, box8));
        s->amap_ptr = xh; s->amap_L = L; s->amap_rows = map_rows;
    }
    ConvMaps maps;
    maps.a_hi = s->amap[0]; maps.a_lo = s->amap[1]; maps.a8_lo = s->amap[2]; maps.a8_hi = s->amap[3];
    maps.b_hi = s->wmap[blk][0]; maps.b_lo = s->wmap[blk][1]; maps.b8_w = s->wmap[blk][2]; maps.b8_lo = s->wmap[blk][3];
    TcParams p;
    p.gemm = 0; p.L = L; p.H = H; p.y_off = y_off; p.tiles_x = cdiv(L, TILE_W); p.num_kb = 50; p.M = H * L; p.out = raw; p.bias = bw.bias;
    const int grid = p.tiles_x * cdiv(H, TILE_H);
    const int cl = e->conv_cluster;
    p.stat_part = nullptr; p.ticket = nullptr; p.norm = nullptr; p.gamma = nullptr; p.totals = nullptr; p.npix = 0;
    if (fuse_stats && cl != 0) {                     // (the CTA-pair kernel keeps the separate statistics pass)
        p.stat_part = e->ws.stat_part; p.ticket = e->ws.ticket; p.norm = e->ws.norm_ss; p.gamma = bw.gamma;
        p.totals = e->strip_on ? e->sp.totals : nullptr;
        p.npix = (double)H * (double)L;
    }
    if (mode == DMP2_CONV_TC_F16X3) return launch_cl<M_F16X3>(e, cl, maps, p, grid, st);
    if (mode == DMP2_CONV_TC_F16F8) return launch_cl<M_F16F8>(e, cl, maps, p, grid, st);
    return launch_cl<M_F16>(e, cl, maps, p, grid, st);
}

bool conv_tc_fuses_stats(const dmp2_engine* e) {
    return e->fuse_stats && e->conv_mode != DMP2_CONV_FFMA && e->conv_cluster != 0;
}

// C[M,512] = A[M,K] * B[512,K]^T through the same TMA / tcgen05 / TMEM pipeline (descriptor + pipeline self-test)
int run_gemm_tn_test(dmp2_engine* e, const float* a, const float* b, int M, int N, int K, int mode, float* c, cudaStream_t st) {
    if (N != 512 || K % KCHUNK != 0 || M < 1) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tn_test: need N == 512 and K % 64 == 0");
    if (mode != DMP2_CONV_TC_F16X3 && mode != DMP2_CONV_TC_F16) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tn_test: f16 / f16x3 modes only");
    TcState* s;
    TRY(get_state(e, &s));
    __half *ah, *al, *bh, *bl;
    const int64_t na = (int64_t)M * K, nb = (int64_t)N * K;
    CUDA_TRY(e, cudaMalloc(&ah, na * 2)); CUDA_TRY(e, cudaMalloc(&al, na * 2));
    CUDA_TRY(e, cudaMalloc(&bh, nb * 2)); CUDA_TRY(e, cudaMalloc(&bl, nb * 2));
    int rc = 0;
    do {
        if ((rc = run_split_half(e, a, na, ah, al, nullptr, nullptr, st))) break;
        if ((rc = run_split_half(e, b, nb, bh, bl, nullptr, nullptr, st))) break;
        ConvMaps maps;
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 1};
        uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * (uint64_t)M};
        uint32_t box[3] = {KCHUNK, TILE_M, 1};
        if ((rc = encode(e, &maps.a_hi, ah, 2, 3, dims, str, box))) break;
        if ((rc = encode(e, &maps.a_lo, al, 2, 3, dims, str, box))) break;
        if ((rc = weight_map(e, &maps.b_hi, bh, K, 2))) break;
        if ((rc = weight_map(e, &maps.b_lo, bl, K, 2))) break;
        maps.a8_lo = maps.a_hi; maps.a8_hi = maps.a_hi; maps.b8_w = maps.b_hi; maps.b8_lo = maps.b_hi;    // unused in these modes
        TcParams p;
        p.gemm = 1; p.L = 0; p.H = 0; p.y_off = 0; p.stat_part = nullptr; p.ticket = nullptr; p.norm = nullptr; p.gamma = nullptr; p.totals = nullptr; p.npix = 0; p.tiles_x = 1; p.num_kb = K / KCHUNK; p.M = M; p.out = c; p.bias = nullptr;
        int grid = cdiv(M, TILE_M);
        rc = (mode == DMP2_CONV_TC_F16X3) ? launch<M_F16X3, 1>(e, maps, p, grid, st) : launch<M_F16, 1>(e, maps, p, grid, st);
    } while (0);
    cudaStreamSynchronize(st);
    cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl);
    return rc;
}

void conv_tc_destroy(dmp2_engine* e) {
    if (e->tc_state) {
        delete (TcState*)e->tc_state;
        e->tc_state = nullptr;
    }
}
