// Tensor-core 5x5 convolution of the ResNet blocks (network.py:26 inside ResNet_Block, + maxout :30-31, + the
// InstanceNorm sums of :32) as an implicit GEMM on the 5th-generation tensor cores:
//   M = L*L pixels, N = 512 output channels, K = 25 taps x 128 input channels.
//
// PERSISTENT kernel, one CTA per SM.  A work unit is a 128-pixel tile (8 rows x 16 columns) x 256 output channels.
// Default form: CTA PAIRS (cta_group::2).  The two CTAs of a cluster own two pixel tiles and behave as one 256 x 256
// tile: every MMA is M=256 N=256, issued by the leader CTA, reading the A rows of both CTAs and HALF of the B piece
// from each CTA's shared memory.  What bounds this kernel once the epilogue is overlapped is the operand stream into
// the SM (measured ceiling ~45-55 bytes/clock/SM whatever the precision mode, profiles/round2_conv_recut.txt); the
// pair form needs 128 KB per SM per tap instead of 192 KB, which puts the fp16x3 mode under that ceiling.
// (cta_group::1 forms with or without TMA multicast of the weights are kept as validation variants.)
//
//   warp 0       TMA producer: cp.async.bulk.tensor (128B swizzle; out-of-bounds zero fill = the pad-2 border) into an
//                ring of k-block stages (shifted pixel patch + the 256-cout weight pieces of that k-block)
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer, M=128 N=256
//   warps 2..3   idle (they complete the first warpgroup, which gives its registers away: setmaxnreg 56)
//   warps 4..11  epilogue (two warpgroups, setmaxnreg 224): drain tensor memory into registers
//
// TWO-LEVEL ACCUMULATION.  The tcgen05 fp32 accumulator truncates (round-toward-zero) on every MMA: one chain over all
// K = 3200 ends 1e-5 (relative) away from the exact sum, 50x the error of an fp32 CPU conv, and that error is what
// kept the fold 5-10x further from the reference than the reference is from exact arithmetic
// (profiles/round2_accumulation_probe.txt).  So a tcgen05 chain here never spans more than ONE TAP (8-24 MMAs): the
// 512 TMEM columns hold two 128 x 256 chunk accumulators; while the issuer fills one, the eight epilogue warps read
// the other with tcgen05.ld and add it, with round-to-nearest, to a running sum held in registers (128 fp32 per
// thread).  Measured on real activations: residual error 1.8e-7 relative for the fp16x3 operand split, the same as
// oneDNN's fp32 conv (2.0e-7).  The drain (128 KB per tap at ~400 B/clock) hides completely behind the next tap's MMAs,
// and so does the tile epilogue (bias, max over 4 consecutive couts, NHWC store, InstanceNorm partial sums), because
// the issuer is already working on the next unit.
//
// Precision modes (x = x_hi + x_lo, w = w_hi + w_lo are fp16 hi/lo splits of the fp32 operands):
//   F16X3  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, all fp16 MMAs     3 MMAs per algorithmic MAC  (fp32-equivalent; parity mode)
//   F16F8  x_hi*w_hi in fp16 + the two correction terms in FP8 (e4m3 activations x e5m2 weights, kind::f8f6f4, twice
//          the fp16 rate), pre-scaled by powers of two that cancel in the product: (x_lo*2^8)(w*2^-8), (x_hi*2^-4)(w_lo*2^4)
//          2 MMA-equivalents per MAC; operand error ~6e-6 relative (fast mode)
//   F16    x_hi*w_hi only                                        1 MMA (informational; 1e-4 relative)
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;               // pixels per unit (8 rows x 16 columns)
constexpr int TILE_H = 8, TILE_W = 16;
constexpr int TILE_N = 256;               // output channels per unit = columns of one chunk accumulator
constexpr int KCHUNK = 64;                // fp16 channels per 128-byte swizzle row
constexpr int A_BYTES = TILE_M * 128;     // 16 KB: 128 pixels x 128 bytes (64 fp16 or 128 fp8 channels)
constexpr int B_BYTES = TILE_N * 128;     // 32 KB: one B piece = 256 couts x 128 bytes
constexpr int BQ_ROWS = 64;               // TMA granule of a B piece (multicast unit)
constexpr int BQ_BYTES = BQ_ROWS * 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + NUM_EPI_WARPS) * 32;
enum { M_F16 = 0, M_F16X3 = 1, M_F16F8 = 2 };

// One ring of k-block stages: [A operand tiles | B pieces], one full / empty mbarrier pair per stage.
template <int MODE, bool PAIR>
struct Cfg {
    static constexpr int A_STAGE_BYTES = MODE == M_F16 ? A_BYTES : 2 * A_BYTES;
    static constexpr int PIECES = MODE == M_F16 ? 1 : 2;      // B pieces per k-block
    static constexpr int B_SLOT_BYTES = PAIR ? B_BYTES / 2 : B_BYTES;     // a CTA of a pair holds 128 of the piece's 256 couts
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + PIECES * B_SLOT_BYTES;
    static constexpr int NUM_STAGES = (224 * 1024) / STAGE_BYTES;        // pair: 3 (f16x3, f16f8) / 7 (f16); else 2 / 4
    static constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct ConvMaps {
    CUtensorMap a_hi, a_lo;              // fp16 activations  [rows][L][128]   (gemm: [M][K])
    CUtensorMap a8_lo, a8_hi;            // e4m3: x_lo * 2^8, x_hi * 2^-4
    CUtensorMap b_hi, b_lo;              // fp16 weights [512][3200]          (gemm: [N][K])
    CUtensorMap b8_w, b8_lo;             // e5m2: w * 2^-8, w_lo * 2^4
};

struct TcParams {
    int gemm;            // 0 = conv (3-D activation map, taps), 1 = plain GEMM C = A B^T (rows x K)
    int L;               // conv: image width (pixels per row)
    int H;               // conv: output rows (== L for a whole image, the strip height for a halo-sharded fold)
    int y_off;           // conv: row of the activation map that holds output row 0 (0, or 2 when the map starts with halo rows)
    int tiles_x;         // conv: tiles per image row
    int num_kb;          // k-blocks per unit: conv 50, gemm K/64
    int chunk_kb;        // k-blocks per tcgen05 accumulation chain (conv: 2 = one tap)
    int M, N;            // gemm: rows, columns of C
    int ldc;             // gemm: row stride of C (floats)
    float alpha;         // gemm: C = alpha * A B^T
    // gemm epilogue variants of the MSA-feature GEMMs (msa.cu): the operand scales live on the device (dsa/dsb point to
    // {scale, 1/scale}; nullptr = 1) and so do the scalars of predict.py:45-51 (scal[1] = n_eff, scal[2] = ridge)
    int ep;              // 0: alpha*acc   1 (Gram): acc + [m==n] ridge n_eff   2 (Woodbury): ([m==n] - acc) / ridge   3 (cov): acc / n_eff + [m==n] ridge
    int m_off;           // row index of C's first row in the full matrix (diagonal test; halo-sharded Woodbury)
    const float* dsa;
    const float* dsb;
    const float* scal;
    int n_tiles_n;       // column tiles of 256 (conv: 2)
    int units;           // cluster-level work units: ceil(row tiles / CL) * n_tiles_n
    float* out;          // conv: raw [pixels][128]; gemm: C [M][N]
    const float* bias;   // conv: [512]
    // fused InstanceNorm statistics of the conv output (network.py:32), nullptr = off.  Every CTA keeps fp64 sums over
    // all its units, leaves them in stat_part at the end; the last CTA of a group folds the group, the last group
    // finishes (deterministic two-level fold, fixed order).
    double* stat_part;           // [grid + groups][256]
    unsigned int* ticket;        // [1 + groups], zero between launches
    float* norm;                 // [mean 128 | gamma * rstd 128]
    const float* gamma;
    double* totals;              // halo-sharded: [sum 128 | sumsq 128] of this launch instead of norm
    double npix;                 // pixels the statistics are over (H * L)
    // dynamic unit scheduler (throughput mode, several folds in flight on one GPU): nullptr = the static round-robin
    // schedule; otherwise the clusters claim units from a global counter, so a launch that only gets part of the SMs
    // (the others are busy with another stream's kernels) is finished by the CTAs that did start, and CTAs that
    // become resident late find nothing left and exit.  The counter is never reset: launch k owns the values
    // [sched_base, sched_base + units + clusters) (every cluster makes exactly one failing claim).
    unsigned long long* sched;
    unsigned long long sched_base;
    // launches captured into a CUDA graph are replayed with frozen arguments: they use a second counter that the launch
    // itself puts back to sched_base (the cluster that draws the last of the units + clusters values knows every other
    // claim has been made)
    int sched_reset;
};
constexpr int STAT_GROUP = 16;   // CTAs per first-level fold
constexpr int UQ = 4;            // depth of the claimed-unit queue (the roles of a cluster are never 4 units apart, see unit_at)

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One lane of the (converged) warp.  The producer and issuer warps run their loops with ALL lanes (warp-uniform control
// flow and operands) and only the TMA / MMA / commit instructions are guarded by this: the single-thread instructions then
// take their operands from uniform registers instead of a per-instruction R2UR + ELECT + retry loop (which made the
// issuer the bottleneck: ~21 SASS instructions per tcgen05.mma, profiles/round2_conv_recut.txt).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Sum v[i] over the 32 lanes of the warp; lane l ends up with the total of v[l] in v[0] (31 shuffles).
__device__ __forceinline__ double lane_transpose_sum(double (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const double send = up ? v[i] : v[i + s];
            const double keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

struct Unit {            // decoded work unit of this CTA
    int x0, y0, m0;      // conv: first pixel of the tile; gemm: first row
    int n0;              // first output column (cout) of the unit
    int nt;              // column tile index
};

template <int CL>
__device__ __forceinline__ Unit decode_unit(const TcParams& p, int cu, uint32_t crank) {
    Unit u;
    u.nt = cu % p.n_tiles_n;
    const int rt = (cu / p.n_tiles_n) * CL + (int)crank;          // row tile (tiles past the end are all out-of-bounds)
    u.n0 = u.nt * TILE_N;
    u.x0 = 0; u.y0 = 0; u.m0 = 0;
    if (p.gemm) u.m0 = rt * TILE_M;
    else { u.y0 = (rt / p.tiles_x) * TILE_H; u.x0 = (rt % p.tiles_x) * TILE_W; }
    return u;
}

// ---- the kernel --------------------------------------------------------------------------------------
// CL = CTAs per cluster sharing the weight (B) stream: each CTA loads 1/CL of every B piece and multicasts it to all
// CTAs of the cluster, which work on CL different pixel tiles of the SAME 256 output channels in lock step.
// PAIR: the cluster is one cta_group::2 CTA pair (see the header comment); CL must be 2.
template <int MODE, int CL, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv5_tc(const __grid_constant__ ConvMaps maps, const TcParams p) {
    using C = Cfg<MODE, PAIR>;
    constexpr int NS = C::NUM_STAGES;
    constexpr int B_SLOT = C::B_SLOT_BYTES;
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 128B swizzle needs 1024-byte alignment
    const uint32_t bar_base = base + NS * C::STAGE_BYTES;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (NS + s); };
    auto c_full = [&](int s) { return bar_base + 8u * (2 * NS + s); };
    auto c_empty = [&](int s) { return bar_base + 8u * (2 * NS + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NS + 4);
    auto uq_full = [&](int s) { return bar_base + 8u * (2 * NS + 5 + s); };
    auto uq_val = [&](int s) { return bar_base + 8u * (2 * NS + 5 + UQ) + 4u * s; };
    static_assert(8 * (2 * NS + 5 + UQ) + 4 * UQ <= 256, "barrier area");
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    const bool leader = crank == 0;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
    const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
    const int chunks_per_unit = p.num_kb / p.chunk_kb;

    if (warp == 0 && lane == 0) {
        // a stage is free when every CTA that delivers into it (weight multicast) has finished reading it
        for (int s = 0; s < NS; s++) { mbar_init(full(s), 1); mbar_init(empty(s), PAIR ? 1 : CL); }
        // pair: the leader's issuer waits for the epilogue warps of BOTH CTAs before it reuses a chunk accumulator
        for (int s = 0; s < 2; s++) { mbar_init(c_full(s), 1); mbar_init(c_empty(s), PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS); }
        for (int s = 0; s < UQ; s++) mbar_init(uq_full(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (PAIR) {
        __syncthreads();
        cluster_sync_all();
        if (warp == 1) {                             // both CTAs, same logical warp, same destination offset
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                  // peers' barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // The i-th unit of this cluster (>= p.units: no more work).  Static schedule: round-robin over the clusters.  Dynamic:
    // rank 0's producer warp claims units from the global counter and posts each one into the queue slot i % UQ of
    // EVERY CTA of the cluster (st.async + expect_tx on that CTA's mbarrier: value and signal in one DSMEM operation);
    // all roles of all CTAs read the same sequence.  No "slot free" handshake is needed: the producer is at most a ring
    // of stages ahead of the issuer and the issuer at most two chunks ahead of the epilogue, so unit i + UQ is claimed
    // long after every role has read unit i.
    const bool dyn = p.sched != nullptr;
    auto unit_at = [&](int i) -> int {
        if (!dyn) return cluster_id + i * num_clusters;
        mbar_wait(uq_full(i & (UQ - 1)), (uint32_t)(i / UQ) & 1u);
        int v;
        asm volatile("ld.volatile.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(uq_val(i & (UQ - 1))) : "memory");
        return v;
    };
    auto claim = [&]() -> unsigned long long { return atomicAdd(p.sched, 1ULL); };     // one lane; consumed much later (post_unit)
    auto post_unit = [&](int i, unsigned long long raw) {                                // one lane of rank 0's producer warp
        const unsigned long long rel = raw - p.sched_base;
        const int v = rel < (unsigned long long)p.units ? (int)rel : p.units;
        if (p.sched_reset && rel == (unsigned long long)(p.units + num_clusters - 1)) atomicExch(p.sched, p.sched_base);
#pragma unroll
        for (int rk = 0; rk < CL; rk++) {
            const uint32_t rb = CL > 1 ? mapa_u32(uq_full(i & (UQ - 1)), rk) : uq_full(i & (UQ - 1));
            const uint32_t rv = CL > 1 ? mapa_u32(uq_val(i & (UQ - 1)), rk) : uq_val(i & (UQ - 1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], 4;" ::"r"(rb) : "memory");
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rv), "r"(v), "r"(rb) : "memory");
        }
    };
    // register re-distribution between the warpgroups (the epilogue holds 128 running sums per thread).  The CTA is
    // launched with 168 registers x 384 threads = 64512; 128 x 56 + 256 x 224 uses exactly that pool (asking for more
    // would block setmaxnreg.inc forever).
    if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===================== TMA producer (whole warp, one elected lane issues) =====================
        {
            int st = 0, ph = 0;
            const bool claimer = dyn && leader && lane == 0;
            if (claimer) post_unit(0, claim());
            __syncwarp();
            for (int ui = 0;; ui++) {
                const int cu = unit_at(ui);
                if (cu >= p.units) break;
                unsigned long long nxt = 0;
                if (claimer) nxt = claim();                    // next unit: the atomic's latency hides behind this unit's loads
                const Unit u = decode_unit<CL>(p, cu, crank);
                for (int kb = 0; kb < p.num_kb; kb++) {
                    const int tap = kb >> 1, sub = kb & 1;
                    int c1, c2;
                    if (p.gemm) { c1 = u.m0; c2 = 0; }
                    else {
                        const int dy = tap / 5, dx = tap - dy * 5;
                        c1 = u.x0 + dx - 2; c2 = u.y0 + dy - 2 + p.y_off;
                    }
                    mbar_wait(empty(st), ph ^ 1);
                    // pair: the loads of both CTAs signal the LEADER's barrier, which expects the bytes of both
                    const uint32_t fl = PAIR ? mapa_u32(full(st), 0) : full(st);
                    const uint32_t ast = base + st * C::STAGE_BYTES;
                    const uint32_t bst = ast + C::A_STAGE_BYTES;
                    if (elect_one()) {
                        if (!PAIR) mbar_expect_tx(full(st), C::A_STAGE_BYTES + C::PIECES * B_BYTES);
                        else if (leader) mbar_expect_tx(full(st), 2 * C::STAGE_BYTES);
                        auto load_a = [&](uint32_t dst, const CUtensorMap* m, int c0) {
                            if (PAIR) tma2_load_3d(dst, m, fl, c0, c1, c2);
                            else tma_load_3d(dst, m, fl, c0, c1, c2);
                        };
                        if (MODE == M_F16) {
                            load_a(ast, &maps.a_hi, p.gemm ? kb * KCHUNK : sub * KCHUNK);
                        } else if (MODE == M_F16X3) {
                            const int c0 = p.gemm ? kb * KCHUNK : sub * KCHUNK;
                            load_a(ast, &maps.a_hi, c0);
                            load_a(ast + A_BYTES, &maps.a_lo, c0);
                        } else if (sub == 0) {                     // F16F8: the two fp8 correction operands (128 channels each)
                            load_a(ast, &maps.a8_lo, 0);
                            load_a(ast + A_BYTES, &maps.a8_hi, 0);
                        } else {                                   // F16F8: both fp16 channel chunks of x_hi
                            load_a(ast, &maps.a_hi, 0);
                            load_a(ast + A_BYTES, &maps.a_hi, KCHUNK);
                        }
#pragma unroll
                        for (int piece = 0; piece < C::PIECES; piece++) {
                            const CUtensorMap* bm;
                            int k0;                                // element offset along K of the weight row
                            if (MODE == M_F16F8) {
                                if (sub == 0) { bm = piece ? &maps.b8_lo : &maps.b8_w; k0 = tap * 128; }
                                else { bm = &maps.b_hi; k0 = tap * 128 + piece * KCHUNK; }
                            } else {
                                bm = piece == 0 ? &maps.b_hi : &maps.b_lo;
                                k0 = kb * KCHUNK;                  // weights are [512][K] with k = tap*128 + c = kb*64 + ...
                            }
                            const uint32_t bdst = bst + piece * B_SLOT;
                            if (PAIR) {                            // my 128 couts of the piece, as two 64-row boxes
                                const int nr = u.n0 + (int)crank * 128;
                                tma2_load_2d(bdst, bm, fl, k0, nr);
                                tma2_load_2d(bdst + BQ_BYTES, bm, fl, k0, nr + BQ_ROWS);
                            } else {
#pragma unroll
                                for (int i = 0; i < 4 / CL; i++) { // my quarters of the piece, delivered to every CTA of the cluster
                                    const int qd = crank * (4 / CL) + i;
                                    if (CL == 1) tma_load_2d(bdst + qd * BQ_BYTES, bm, fl, k0, u.n0 + qd * BQ_ROWS);
                                    else tma_load_2d_mc(bdst + qd * BQ_BYTES, bm, fl, k0, u.n0 + qd * BQ_ROWS, MC_MASK);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (++st == NS) { st = 0; ph ^= 1; }
                }
                if (claimer) post_unit(ui + 1, nxt);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane issues; pair: the leader CTA only) ==========
        if (!PAIR || leader) {
            const uint32_t idesc = make_idesc(PAIR ? 2 * TILE_M : TILE_M, TILE_N);
            const uint32_t idesc8 = make_idesc_f8(PAIR ? 2 * TILE_M : TILE_M, TILE_N);
            // 4 MMAs over one 128-byte swizzle row of K: the descriptors advance by 32 bytes (2 units of 16) per MMA
            auto mma16x4 = [&](uint32_t d, uint64_t da, uint64_t db, bool first) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (PAIR) tc2_mma_f16(d, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                    else tc_mma_f16(d, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                }
            };
            auto mma8x4 = [&](uint32_t d, uint64_t da, uint64_t db, bool first) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (PAIR) tc2_mma_f8(d, da + 2 * k, db + 2 * k, idesc8, !(first && k == 0));
                    else tc_mma_f8(d, da + 2 * k, db + 2 * k, idesc8, !(first && k == 0));
                }
            };
            auto commit = [&](uint32_t bar, bool shared_slot) {            // shared_slot: a B slot every CTA of the cluster waits for
                if (PAIR) tc2_commit_mc(bar, 3);
                else if (shared_slot && CL > 1) tc_commit_mc(bar, MC_MASK);
                else tc_commit(bar);
            };
            int st = 0, ph = 0;
            uint32_t cc = 0;                                      // chunk counter over the whole kernel: buffer cc & 1
            for (int ui = 0; unit_at(ui) < p.units; ui++) {
                for (int kb = 0; kb < p.num_kb; kb++) {
                    const int sub = kb & 1;
                    const int in_chunk = kb % p.chunk_kb;
                    const uint32_t cs = cc & 1u;
                    if (in_chunk == 0) mbar_wait(c_empty(cs), ((cc >> 1) & 1u) ^ 1u);   // the epilogue has drained this chunk accumulator
                    const uint32_t d = tmem_base + cs * (uint32_t)TILE_N;
                    mbar_wait(full(st), ph);
                    tc_fence_after();
                    const uint32_t ast = base + st * C::STAGE_BYTES;
                    const uint64_t da0 = make_smem_desc(ast), da1 = make_smem_desc(ast + A_BYTES);
                    const uint64_t db0 = make_smem_desc(ast + C::A_STAGE_BYTES), db1 = make_smem_desc(ast + C::A_STAGE_BYTES + B_SLOT);
                    const bool first = in_chunk == 0;                                   // first MMA of the chain overwrites
                    if (elect_one()) {
                        if (MODE == M_F16F8) {
                            if (sub == 0) { mma8x4(d, da0, db0, first); mma8x4(d, da1, db1, false); }   // corrections first (K = 32 fp8 per MMA)
                            else { mma16x4(d, da0, db0, first); mma16x4(d, da1, db1, false); }
                        } else {
                            mma16x4(d, da0, db0, first);                                // hi(A) x hi(B)
                            if (MODE == M_F16X3) { mma16x4(d, da1, db0, false); mma16x4(d, da0, db1, false); }   // lo x hi, hi x lo
                        }
                        commit(empty(st), true);                   // the stage is free when every CTA of the cluster is done with it
                        if (in_chunk == p.chunk_kb - 1) commit(c_full(cs), false);      // chain finished: hand it to the epilogue
                    }
                    __syncwarp();
                    if (++st == NS) { st = 0; ph ^= 1; }
                    if (in_chunk == p.chunk_kb - 1) cc++;
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===================== epilogue (warps 4..11) =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int ew = warp - FIRST_EPI_WARP;
        const int hcol = ew >> 2;                    // which 128 of the chunk's 256 columns
        const int r = q * 32 + lane;                 // accumulator row = pixel within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hcol * 128);
        double st_s[2] = {0.0, 0.0}, st_q[2] = {0.0, 0.0};        // per column tile: this lane's channel, summed over all units
        uint32_t cc = 0;
        for (int ui = 0;; ui++) {
            const int cu = unit_at(ui);
            if (cu >= p.units) break;
            const Unit u = decode_unit<CL>(p, cu, crank);
            float R[128];
            for (int ch = 0; ch < chunks_per_unit; ch++, cc++) {
                const uint32_t cs = cc & 1u;
                mbar_wait(c_full(cs), (cc >> 1) & 1u);
                tc_fence_after();
                const uint32_t ta = lane_addr + cs * (uint32_t)TILE_N;
#pragma unroll
                for (int g = 0; g < 4; g += 2) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(ta + g * 32, v0);
                    tmem_ld32_nowait(ta + (g + 1) * 32, v1);
                    tmem_wait_ld();
                    if (ch == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i++) { R[g * 32 + i] = __uint_as_float(v0[i]); R[(g + 1) * 32 + i] = __uint_as_float(v1[i]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) { R[g * 32 + i] += __uint_as_float(v0[i]); R[(g + 1) * 32 + i] += __uint_as_float(v1[i]); }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && !leader)
                        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(c_empty(cs), 0)) : "memory");
                    else mbar_arrive(c_empty(cs));
                }
            }
            // ---- unit epilogue (the issuer is already filling the next unit's chunks)
            if (p.gemm) {
                const int64_t row = u.m0 + r;
                if (row < p.M) {
                    float* dst = p.out + row * p.ldc + u.n0 + hcol * 128;
                    const int ncol = p.N - (u.n0 + hcol * 128);
                    float al = p.alpha;
                    if (p.dsa) al *= __ldg(p.dsa + 1);
                    if (p.dsb) al *= __ldg(p.dsb + 1);
                    float mul = al, dg = 0.f;                      // C = mul * acc + [m == n] * dg
                    if (p.ep == 1) dg = __ldg(p.scal + 2) * __ldg(p.scal + 1);
                    else if (p.ep == 2) { const float rr = 1.0f / __ldg(p.scal + 2); mul = -al * rr; dg = rr; }
                    else if (p.ep == 3) { mul = al / __ldg(p.scal + 1); dg = __ldg(p.scal + 2); }
                    const int dcol = (int)(row + p.m_off) - (u.n0 + hcol * 128);      // column of this thread's slice on the diagonal
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        if (4 * i < ncol) {
                            float4 o = make_float4(mul * R[4 * i], mul * R[4 * i + 1], mul * R[4 * i + 2], mul * R[4 * i + 3]);
                            if (p.bias) {                          // plain GEMM + per-column bias (GRU input projections)
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + u.n0 + hcol * 128 + 4 * i));
                                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                            }
                            if (p.ep != 0 && (dcol >> 2) == i) {
                                if ((dcol & 3) == 0) o.x += dg; else if ((dcol & 3) == 1) o.y += dg; else if ((dcol & 3) == 2) o.z += dg; else o.w += dg;
                            }
                            *reinterpret_cast<float4*>(dst + 4 * i) = o;
                        }
                }
            } else {
                const int y = u.y0 + (r >> 4), x = u.x0 + (r & 15);
                const bool valid = (y < p.H) && (x < p.L);
                const float* bias = p.bias + u.n0 + hcol * 128;
                float o[32];
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + 4 * i));
                    o[i] = fmaxf(fmaxf(R[4 * i] + b.x, R[4 * i + 1] + b.y), fmaxf(R[4 * i + 2] + b.z, R[4 * i + 3] + b.w));
                }
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(p.out + ((int64_t)y * p.L + x) * 128 + u.nt * 64 + hcol * 32);
#pragma unroll
                    for (int i = 0; i < 8; i++) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                }
                if (p.stat_part) {                   // InstanceNorm partial sums, fp64 throughout (the same sums k_in_stats forms)
                    double sv[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) sv[i] = valid ? (double)o[i] : 0.0;
                    const double s1 = lane_transpose_sum(sv, lane);
#pragma unroll
                    for (int i = 0; i < 32; i++) sv[i] = valid ? (double)o[i] * (double)o[i] : 0.0;
                    const double s2 = lane_transpose_sum(sv, lane);
                    if (u.nt & 1) { st_s[1] += s1; st_q[1] += s2; }
                    else { st_s[0] += s1; st_q[0] += s2; }
                }
            }
        }
        if (!p.gemm && p.stat_part) {
            // every MMA has completed and every TMA load has been consumed: the operand rings are free
            asm volatile("bar.sync 1, 256;" ::: "memory");
            double* stage = reinterpret_cast<double*>(smem_gen);              // [8 warps][2 column tiles][sum|sumsq][32 lanes]
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                stage[((ew * 2 + nt) * 2 + 0) * 32 + lane] = st_s[nt];
                stage[((ew * 2 + nt) * 2 + 1) * 32 + lane] = st_q[nt];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int t = threadIdx.x - FIRST_EPI_WARP * 32;          // 0..255: (sum|sumsq, channel)
            const int sidx = t >> 7, c = t & 127, nt = c >> 6, hc = (c >> 5) & 1, ln = c & 31;
            double a = 0.0;
#pragma unroll
            for (int qq = 0; qq < 4; qq++)           // lane quarters in a fixed order (epilogue warp e serves quarter e & 3)
                a += stage[(((hc * 4 + qq) * 2 + nt) * 2 + sidx) * 32 + ln];
            double* part = p.stat_part;
            part[(int64_t)blockIdx.x * 256 + t] = a;
            // deterministic two-level fold: last CTA of a group folds the group, last group folds the group sums
            const unsigned G = STAT_GROUP, grp_id = blockIdx.x / G, ngrp = (gridDim.x + G - 1) / G;
            const unsigned gfirst = grp_id * G, gsize = min(G, gridDim.x - gfirst);
            double* part2 = part + (int64_t)gridDim.x * 256;
            int* flag = reinterpret_cast<int*>(smem_gen + 16384);
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (t == 0) *flag = (atomicAdd(p.ticket + 1 + grp_id, 1u) == gsize - 1) ? 1 : 0;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (*flag != 0) {
                __threadfence();
                a = 0.0;
#pragma unroll 8
                for (unsigned q2 = 0; q2 < gsize; q2++) a += __ldcg(part + (int64_t)(gfirst + q2) * 256 + t);
                part2[(int64_t)grp_id * 256 + t] = a;
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (t == 0) {
                    p.ticket[1 + grp_id] = 0;
                    *flag = (atomicAdd(p.ticket, 1u) == ngrp - 1) ? 2 : 1;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (*flag == 2) {
                    __threadfence();
                    a = 0.0;
#pragma unroll 8
                    for (unsigned q2 = 0; q2 < ngrp; q2++) a += __ldcg(part2 + (int64_t)q2 * 256 + t);
                    if (p.totals) p.totals[t] = a;
                    else {
                        stage[t] = a;
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        if (t < 128) {
                            const double mean = stage[t] / p.npix;
                            double var = stage[128 + t] / p.npix - mean * mean;
                            if (var < 0) var = 0;
                            p.norm[t] = (float)mean;
                            p.norm[128 + t] = (float)((double)p.gamma[t] / sqrt(var + 1e-5));
                        }
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
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- host side: tensor maps -----------------------------------------------------------------------------
struct TcState {
    CUtensorMap wmap[DMP2_NBLOCKS][4];               // b_hi, b_lo, b8_w, b8_lo
    bool wmap_ok[DMP2_NBLOCKS] = {false};
    CUtensorMap amap[4];                             // a_hi, a_lo, a8_lo, a8_hi
    const void* amap_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    int amap_L = 0, amap_rows = 0;
    bool attr_set = false;
};

template <int MODE, int CL, bool PAIR>
int set_attr(dmp2_engine* e) {
    CUDA_TRY(e, cudaFuncSetAttribute(k_conv5_tc<MODE, CL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE, PAIR>::SMEM_BYTES));
    return 0;
}

int get_state(dmp2_engine* e, TcState** out) {
    if (!e->tc_state) {
        if (!tc::get_encode_fn()) return e->fail(DMP2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
        e->tc_state = new TcState();
    }
    TcState* s = (TcState*)e->tc_state;
    if (!s->attr_set) {
        TRY((set_attr<M_F16, 1, false>(e))); TRY((set_attr<M_F16, 2, false>(e))); TRY((set_attr<M_F16, 2, true>(e)));
        TRY((set_attr<M_F16X3, 1, false>(e))); TRY((set_attr<M_F16X3, 2, false>(e))); TRY((set_attr<M_F16X3, 2, true>(e)));
        TRY((set_attr<M_F16F8, 1, false>(e))); TRY((set_attr<M_F16F8, 2, false>(e))); TRY((set_attr<M_F16F8, 2, true>(e)));
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

// weights [rows][K] (K contiguous), box = one 128-byte swizzle row x 64 rows
int weight_map(dmp2_engine* e, CUtensorMap* map, const void* w, int rows, int K, int elem_bytes) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)K * elem_bytes};
    uint32_t box[2] = {(uint32_t)(128 / elem_bytes), BQ_ROWS};
    return encode(e, map, w, elem_bytes, 2, dims, str, box);
}

template <int MODE, int CL, bool PAIR>
int launch(dmp2_engine* e, const ConvMaps& maps, const TcParams& p, int grid, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<MODE, PAIR>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(e, cudaLaunchKernelEx(&cfg, k_conv5_tc<MODE, CL, PAIR>, maps, p));
    POST_LAUNCH(e, "k_conv5_tc");
    return 0;
}

template <int MODE>
int launch_form(dmp2_engine* e, int form, const ConvMaps& maps, const TcParams& p, int grid, cudaStream_t st) {
    if (form == 0) return launch<MODE, 2, true>(e, maps, p, grid, st);
    if (form == 2) return launch<MODE, 2, false>(e, maps, p, grid, st);
    return launch<MODE, 1, false>(e, maps, p, grid, st);
}

// persistent grid: one CTA per SM (or fewer when there is less work), a whole number of clusters.
// form: 0 = cta_group::2 CTA pairs (default), 1 = independent CTAs, 2 = cta_group::1 with 2-CTA weight multicast
int launch_mode(dmp2_engine* e, int mode, int form, const ConvMaps& maps, TcParams& p, int row_tiles, cudaStream_t st) {
    const int cl = form == 1 ? 1 : 2;
    p.units = cdiv(row_tiles, cl) * p.n_tiles_n;
    int sms = e->conv_sms > 0 ? std::min(e->conv_sms, e->num_sms) : e->num_sms;
    int grid = std::min(sms / cl, p.units) * cl;
    if (grid < cl) grid = cl;
    p.sched_reset = 0;
    if (p.sched && e->capturing) {                   // graph node: frozen arguments, so a counter the launch itself rewinds
        p.sched = e->ws.sched + 1;
        p.sched_base = 0;
        p.sched_reset = 1;
    } else if (p.sched) {                            // this launch owns the next units + clusters counter values
        p.sched_base = e->conv_sched_base;
        e->conv_sched_base += (unsigned long long)p.units + (unsigned long long)(grid / cl);
    }
    if (mode == DMP2_CONV_TC_F16X3) return launch_form<M_F16X3>(e, form, maps, p, grid, st);
    if (mode == DMP2_CONV_TC_F16F8) return launch_form<M_F16F8>(e, form, maps, p, grid, st);
    return launch_form<M_F16>(e, form, maps, p, grid, st);
}

}  // namespace

// xh/xl/x8lo/x8hi: activation maps of map_rows x L pixels; output row y reads map rows y + y_off - 2 .. y + y_off + 2
// (rows outside the map read as zero), H output rows are written to raw.  Whole image: map_rows = H = L, y_off = 0.
// fuse_stats: also leave the InstanceNorm statistics of the output in ws.norm_ss (halo-sharded: sp.totals).
int run_conv_tc(dmp2_engine* e, int blk, const __half* xh, const __half* xl, const uint8_t* x8lo, const uint8_t* x8hi, int L,
                int H, int y_off, int map_rows, float* raw, int mode, cudaStream_t st, bool fuse_stats) {
    TcState* s;
    TRY(get_state(e, &s));
    const ResBlockW& bw = e->w.blk[blk];
    if (!s->wmap_ok[blk]) {
        TRY(weight_map(e, &s->wmap[blk][0], bw.w_hi, 512, 3200, 2));
        TRY(weight_map(e, &s->wmap[blk][1], bw.w_lo, 512, 3200, 2));
        TRY(weight_map(e, &s->wmap[blk][2], bw.w8_w, 512, 3200, 1));
        TRY(weight_map(e, &s->wmap[blk][3], bw.w8_lo, 512, 3200, 1));
        s->wmap_ok[blk] = true;
    }
    const void* ptrs[4] = {xh, xl, x8lo, x8hi};
    if (memcmp(s->amap_ptr, ptrs, sizeof(ptrs)) != 0 || s->amap_L != L || s->amap_rows != map_rows) {
        uint64_t dims[3] = {128, (uint64_t)L, (uint64_t)map_rows};
        uint64_t str16[2] = {256, (uint64_t)L * 256}, str8[2] = {128, (uint64_t)L * 128};
        uint32_t box16[3] = {KCHUNK, TILE_W, TILE_H}, box8[3] = {128, TILE_W, TILE_H};
        TRY(encode(e, &s->amap[0], xh, 2, 3, dims, str16, box16));
        TRY(encode(e, &s->amap[1], xl, 2, 3, dims, str16, box16));
        TRY(encode(e, &s->amap[2], x8lo, 1, 3, dims, str8, box8));
        TRY(encode(e, &s->amap[3], x8hi, 1, 3, dims, str8, box8));
        memcpy(s->amap_ptr, ptrs, sizeof(ptrs)); s->amap_L = L; s->amap_rows = map_rows;
    }
    ConvMaps maps;
    maps.a_hi = s->amap[0]; maps.a_lo = s->amap[1]; maps.a8_lo = s->amap[2]; maps.a8_hi = s->amap[3];
    maps.b_hi = s->wmap[blk][0]; maps.b_lo = s->wmap[blk][1]; maps.b8_w = s->wmap[blk][2]; maps.b8_lo = s->wmap[blk][3];
    TcParams p;
    p.gemm = 0; p.L = L; p.H = H; p.y_off = y_off; p.tiles_x = cdiv(L, TILE_W); p.num_kb = 50;
    p.chunk_kb = 2 * (e->conv_chunk_taps > 0 ? e->conv_chunk_taps : (mode == DMP2_CONV_TC_F16X3 ? 1 : 5));
    p.M = H * L; p.N = 512; p.ldc = 512; p.alpha = 1.0f; p.ep = 0; p.m_off = 0; p.dsa = nullptr; p.dsb = nullptr; p.scal = nullptr;
    p.n_tiles_n = 2; p.out = raw; p.bias = bw.bias;
    p.stat_part = nullptr; p.ticket = nullptr; p.norm = nullptr; p.gamma = nullptr; p.totals = nullptr; p.npix = 0;
    p.sched = e->conv_dynamic ? e->ws.sched : nullptr; p.sched_base = 0;
    if (fuse_stats) {
        p.stat_part = e->ws.stat_part; p.ticket = e->ws.ticket; p.norm = e->ws.norm_ss; p.gamma = bw.gamma;
        p.totals = e->strip_on ? e->sp.totals : nullptr;
        p.npix = (double)H * (double)L;
    }
    return launch_mode(e, mode, e->conv_cluster, maps, p, p.tiles_x * cdiv(H, TILE_H), st);
}

bool conv_tc_fuses_stats(const dmp2_engine* e) { return e->fuse_stats && e->conv_mode != DMP2_CONV_FFMA; }

// the activation tensor maps are cached by pointer: forget them whenever the buffers they describe are released
void conv_tc_invalidate(dmp2_engine* e) {
    if (!e->tc_state) return;
    TcState* s = (TcState*)e->tc_state;
    for (auto& ptr : s->amap_ptr) ptr = nullptr;
    s->amap_L = 0; s->amap_rows = 0;
}

// C[M,N] = A[M,K] * B[N,K]^T through the same TMA / tcgen05 / TMEM pipeline, accumulation chains of chunk_k elements
// (descriptor + pipeline self-test, and the accumulation-error probes of tools/)
int run_gemm_tn_test(dmp2_engine* e, const float* a, const float* b, int M, int N, int K, int mode, int chunk_k, float* c, cudaStream_t st) {
    if (N < 4 || N % 4 != 0 || K % KCHUNK != 0 || M < 1) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tn_test: need N % 4 == 0 and K % 64 == 0");
    if (mode != DMP2_CONV_TC_F16X3 && mode != DMP2_CONV_TC_F16) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tn_test: f16 / f16x3 modes only");
    if (chunk_k <= 0) chunk_k = K;
    if (chunk_k % KCHUNK != 0 || K % chunk_k != 0) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tn_test: chunk_k must be a multiple of 64 that divides K");
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
        if ((rc = weight_map(e, &maps.b_hi, bh, N, K, 2))) break;
        if ((rc = weight_map(e, &maps.b_lo, bl, N, K, 2))) break;
        maps.a8_lo = maps.a_hi; maps.a8_hi = maps.a_hi; maps.b8_w = maps.b_hi; maps.b8_lo = maps.b_hi;    // unused in these modes
        TcParams p;
        p.gemm = 1; p.L = 0; p.H = 0; p.y_off = 0; p.tiles_x = 1; p.num_kb = K / KCHUNK; p.chunk_kb = chunk_k / KCHUNK;
        p.M = M; p.N = N; p.ldc = N; p.alpha = 1.0f; p.ep = 0; p.m_off = 0; p.dsa = nullptr; p.dsb = nullptr; p.scal = nullptr;
        p.n_tiles_n = cdiv(N, TILE_N); p.out = c; p.bias = nullptr;
        p.sched = (e->conv_dynamic && e->ws.sched) ? e->ws.sched : nullptr; p.sched_base = 0;      // (after dmp2_reserve: exercises the dynamic schedule)
        p.stat_part = nullptr; p.ticket = nullptr; p.norm = nullptr; p.gamma = nullptr; p.totals = nullptr; p.npix = 0;
        rc = launch_mode(e, mode, e->conv_cluster, maps, p, cdiv(M, TILE_M), st);
    } while (0);
    cudaStreamSynchronize(st);
    cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl);
    return rc;
}

// ---- fp32 GEMMs of the other stages on the same tensor-core pipeline ---------------------------------------------
// x (rows x cols, row stride ld) * scale -> fp16 hi / lo copies [rows][Kp], zero-padded to Kp columns.  `scale` is a
// power of two chosen by the caller so that the values sit well inside the fp16 range (the lo parts are 2^-12 of the
// values; without it small entries would lose their lo part to fp16 underflow).
__global__ void __launch_bounds__(256) k_split_scaled(const float* __restrict__ x, int rows, int cols, int64_t ld, float scale,
                                                      const float* __restrict__ dscale, __half* __restrict__ hi,
                                                      __half* __restrict__ lo, int Kp) {
    if (dscale) scale *= __ldg(dscale);
    const int64_t n = (int64_t)rows * (Kp / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / (Kp / 4)), c = (int)(i % (Kp / 4)) * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (c + j < cols) ? x[(int64_t)r * ld + c + j] * scale : 0.f;
        __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
        float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(hi + (int64_t)r * Kp + c) = hv;
        *reinterpret_cast<uint2*>(lo + (int64_t)r * Kp + c) = lv;
    }
}

// largest |x| of a rows x cols matrix -> ds = {2^k, 2^-k} with max * 2^k in [2^12, 2^13): the operand then sits in the
// middle of the fp16 range and its lo part (2^-12 of it) stays normal
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, int rows, int cols, int64_t ld, unsigned int* __restrict__ mx) {
    float m = 0.f;
    const int64_t n = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[(i / cols) * ld + (i % cols)]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(mx, __float_as_uint(m));             // non-negative floats order like their bit patterns
}
__global__ void k_scale_from_max(unsigned int* mx, float* ds) {
    const float m = __uint_as_float(*mx);
    int ex = 0;
    if (m > 0.f && isfinite(m)) frexpf(m, &ex);                                   // m = f * 2^ex, f in [0.5, 1)
    const int k = 13 - ex;                                                         // m * 2^k in [2^12, 2^13)
    ds[0] = ldexpf(1.0f, k);
    ds[1] = ldexpf(1.0f, -k);
    *mx = 0u;
}

int run_operand_scale(dmp2_engine* e, const float* x, int rows, int cols, int64_t ld, float* ds, cudaStream_t st) {
    unsigned int* mx = reinterpret_cast<unsigned int*>(ds + 2);                  // scratch word next to the pair, zero between uses
    const int64_t n = (int64_t)rows * cols;
    const int grid = (int)std::min<int64_t>(cdiv64(n, 1024), (int64_t)e->num_sms * 4);
    k_absmax<<<std::max(grid, 1), 256, 0, st>>>(x, rows, cols, ld, mx);
    POST_LAUNCH(e, "k_absmax");
    k_scale_from_max<<<1, 1, 0, st>>>(mx, ds);
    POST_LAUNCH(e, "k_scale_from_max");
    return 0;
}

int run_split_scaled(dmp2_engine* e, const float* x, int rows, int cols, int64_t ld, float scale, const float* dscale, __half* hi,
                     __half* lo, int Kp, cudaStream_t st) {
    const int64_t n = (int64_t)rows * (Kp / 4);
    const int grid = (int)std::min<int64_t>(cdiv64(n, 256), (int64_t)e->num_sms * 8);
    k_split_scaled<<<grid, 256, 0, st>>>(x, rows, cols, ld, scale, dscale, hi, lo, Kp);
    POST_LAUNCH(e, "k_split_scaled");
    return 0;
}

// C[M][ldc] = alpha * A[M][Kp] * B[N][Kp]^T, operands given as fp16 hi/lo pairs (Kp % 64 == 0), 3 MMAs per MAC,
// accumulation chains of chunk_k elements summed in fp32 registers.  Asynchronous on st, no allocation.
int run_gemm_tc(dmp2_engine* e, const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo, int M, int N, int Kp,
                float alpha, float* c, int ldc, int chunk_k, cudaStream_t st, const GemmTcEpilogue* ep, int b_rows) {
    if (b_rows <= 0) b_rows = N;                                       // rows of B that exist (beyond them the TMA zero-fills)
    if (M < 1 || N < 4 || N % 4 != 0 || Kp % KCHUNK != 0 || ldc % 4 != 0) return e->fail(DMP2_ERR_BAD_ARG, "gemm_tc: bad shape");
    int ck = chunk_k > 0 ? chunk_k : 128;
    while (ck > KCHUNK && Kp % ck != 0) ck -= KCHUNK;               // longest chain <= chunk_k that divides Kp
    TcState* s;
    TRY(get_state(e, &s));
    ConvMaps maps;
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)M, 1};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Kp * 2 * (uint64_t)M};
    uint32_t box[3] = {KCHUNK, TILE_M, 1};
    TRY(encode(e, &maps.a_hi, a_hi, 2, 3, dims, str, box));
    TRY(encode(e, &maps.a_lo, a_lo, 2, 3, dims, str, box));
    TRY(weight_map(e, &maps.b_hi, b_hi, b_rows, Kp, 2));
    TRY(weight_map(e, &maps.b_lo, b_lo, b_rows, Kp, 2));
    maps.a8_lo = maps.a_hi; maps.a8_hi = maps.a_hi; maps.b8_w = maps.b_hi; maps.b8_lo = maps.b_hi;        // unused
    TcParams p;
    p.gemm = 1; p.L = 0; p.H = 0; p.y_off = 0; p.tiles_x = 1; p.num_kb = Kp / KCHUNK; p.chunk_kb = ck / KCHUNK;
    p.M = M; p.N = N; p.ldc = ldc; p.alpha = alpha; p.n_tiles_n = cdiv(N, TILE_N); p.out = c; p.bias = nullptr;
    p.sched = nullptr; p.sched_base = 0;
    p.ep = ep ? ep->kind : 0; p.m_off = ep ? ep->m_off : 0; p.dsa = ep ? ep->dsa : nullptr; p.dsb = ep ? ep->dsb : nullptr; p.scal = ep ? ep->scal : nullptr;
    p.bias = ep ? ep->bias : nullptr;
    p.stat_part = nullptr; p.ticket = nullptr; p.norm = nullptr; p.gamma = nullptr; p.totals = nullptr; p.npix = 0;
    return launch_mode(e, DMP2_CONV_TC_F16X3, e->conv_cluster, maps, p, cdiv(M, TILE_M), st);
}

void conv_tc_destroy(dmp2_engine* e) {
    if (e->tc_state) {
        delete (TcState*)e->tc_state;
        e->tc_state = nullptr;
    }
}
