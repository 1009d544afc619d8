// Shared declarations for libdmp2.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/dmp2.h"

#define DMP2_CH 128          // ResNet width (network.py: GRUResNet(512,128))
#define DMP2_W 512           // GRU width
#define DMP2_NBLOCKS 16
#define DMP2_FEAT_LD 444     // 441 DCA + 1 APC + 2 zero pad (float4-aligned rows)
#define DMP2_STEM_K (512 + DMP2_FEAT_LD)
#define DMP2_STEM_KP 1024     // K of the stem GEMM padded to whole 128-element accumulation chains (tensor-core path)
#define DMP2_STEM_SW 256.0f   // power-of-two scale of the stem weights in their fp16 hi/lo copies
#define DMP2_GRU_SW 1024.0f   // power-of-two scale of the GRU input weights in their fp16 hi/lo copies
#define DMP2_GRU_SA 64.0f     // ... and of the GRU layer inputs (hidden states in [-1,1], MDS coordinates up to a few 10)
#define DMP2_TC_SLAB 32768    // rows of a GEMM A operand staged at a time in ws.tc_scratch

struct Launch {              // per-engine launch bookkeeping (gpu_launches) + sticky error
    int64_t count = 0;
    std::string err;
    int status = 0;
};

#define CUDA_TRY(e, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t _c = (call);                                                                            \
        if (_c != cudaSuccess) {                                                                            \
            return (e)->fail(_c == cudaErrorMemoryAllocation ? DMP2_ERR_OOM : DMP2_ERR_CUDA,                \
                             std::string(#call) + ": " + cudaGetErrorString(_c));                           \
        }                                                                                                   \
    } while (0)

// After every kernel launch: count it and catch launch-configuration errors immediately.
#define POST_LAUNCH(e, name)                                                                                \
    do {                                                                                                    \
        (e)->launches++;                                                                                    \
        cudaError_t _c = cudaGetLastError();                                                                \
        if (_c != cudaSuccess) return (e)->fail(DMP2_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(_c)); \
    } while (0)

#define TRY(x)                                                                                              \
    do {                                                                                                    \
        int _s = (x);                                                                                       \
        if (_s != 0) return _s;                                                                             \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------
// Repacked weights (device).  Layouts are documented in DESIGN.md "Data layout in HBM".
// ---------------------------------------------------------------------------------------------------
struct GruDir {              // one direction of one layer of a hidden-256 bidirectional GRU
    float* w_hh;             // [768][256] as in the state_dict (gate order r,z,n)
    float* b_hh;             // [768]
};
struct BiGruLayer {
    float* w_ih;             // [1536][K] : rows 0..767 forward, 768..1535 reverse  (K = input width, mult of 4)
    float* b_ih;             // [1536]
    __half* w_ih_hi;         // [1536][Kp] fp16 hi/lo split of w_ih * DMP2_GRU_SW (tensor-core input projection), Kp = K rounded up to 64
    __half* w_ih_lo;
    int K;
    GruDir dir[2];
};

struct ResBlockW {
    float* w_f32;            // [512][25][128]  (cout, tap=ky*5+kx, cin)  fp32, for the FFMA path
    __half* w_hi;            // same layout, fp16 high part
    __half* w_lo;            // fp16 residual  (w - float(w_hi))
    uint8_t* w8_w;           // e5m2 of w * 2^-8   (FP8 correction term x_lo * w)
    uint8_t* w8_lo;          // e5m2 of w_lo * 2^4 (FP8 correction term x_hi * w_lo)
    float* bias;             // [512]
    float* gamma;            // [128]
    float* beta;             // [128]
    float* gate_c;           // [128]  cSE gate, a weights-only constant: sigmoid(W2 relu(W1 beta))
    float* sse_w;            // [128]
    float sse_b;
};

struct Weights {
    // vgru (network.py:189): packed so that one GEMM row-quad yields all gates of a hidden unit
    float* vg_gi0;           // [22][512][4]   W_ih_l0 column gather + b_ih (r, z, n, 0)
    float* vg_w0;            // [2048][512]    rows 4j+{0,1,2}: W_hh_l0 rows {j,512+j,1024+j}; 4j+3: 0
    float* vg_b0;            // [2048]         b_hh_l0 in the same order
    float* vg_w1;            // [2048][1024]   4j+0/1: [W_ih_l1 | W_hh_l1] (r,z); 4j+2: [W_ih_n | 0]; 4j+3: [0 | W_hh_n]
    float* vg_b1;            // [2048]         (b_ih_r+b_hh_r, b_ih_z+b_hh_z, b_ih_n, b_hh_n)
    // vgru, tensor-core path: rows packed [slice of 32 units][gate r|z|n][32] for the three K=512 GEMM roles
    // (0: W_hh_l0, 1: W_ih_l1, 2: W_hh_l1), fp16 hi/lo split
    __half* vt_w_hi[3];
    __half* vt_w_lo[3];
    float* vt_bias[3];       // [1536] each, same packed order (b_hh_l0, b_ih_l1, b_hh_l1)
    float* vt_gi0;           // [22][1536] W_ih_l0 column gather + b_ih_l0, packed order
    BiGruLayer hgru[2];
    BiGruLayer cgru[3];
    float* coord_fc;         // [3][512]
    float* stem_w;           // [384][DMP2_STEM_K]  cols 0..511 outer, 512..953 DCA+APC, pad 0
    __half* stem_w_hi;       // [384][DMP2_STEM_KP] fp16 hi/lo split of stem_w * DMP2_STEM_SW (tensor-core stem GEMM)
    __half* stem_w_lo;
    float* stem_b;           // [384]
    float* stem_wd;          // [384]  weight of input channel 954 (the recycled distance map)
    float* stem_gamma;       // [128]
    float* stem_beta;        // [128]
    ResBlockW blk[DMP2_NBLOCKS];
    float* head_w;           // [2][128]
    float head_b[2];
};

// ---------------------------------------------------------------------------------------------------
// Workspace (device), sized for (L, N); grown on demand.
// ---------------------------------------------------------------------------------------------------
struct Workspace {
    int L = 0, N = 0;
    // MSA features
    uint8_t* msa = nullptr;        // [N][L] codes (engine-owned copy for the host entry point)
    uint8_t* msa_t = nullptr;      // [L][Npad4] clamped codes, column-major for coalesced identity counts
    float* seqw = nullptr;         // [N]
    float* scal = nullptr;         // [8] scalars: sum w, n_eff, ridge, ...
    float* xc = nullptr;           // [21L][Npad4] centred, sqrt(w)-scaled one-hot, transposed (K = sequence axis)
    float* xct = nullptr;          // [Npad4][n4] transposed copy (Woodbury path)
    float* kmat = nullptr;         // [Npad64][Npad64] Gram system of the Woodbury path
    float* wy = nullptr;           // [Npad4][n4] K^-1 X^T
    float* cov = nullptr;          // [npad][npad] covariance -> inverse in place
    float* gj_p = nullptr;         // [64][64] pivot inverse
    float* gj_r = nullptr;         // [64][npad] row panel
    float* x3 = nullptr;           // [L][L] contact norms
    float* apc = nullptr;          // [2L+1] row sums, col sums, total
    float* feat = nullptr;         // [L*L][444]
    // 1-D track
    float* vg_h = nullptr;         // [2 layers][2 buffers][L][512]
    __half* vt_h16 = nullptr;      // [4 buffers][hi|lo][L][512] fp16 split of the vgru states
    float* vt_gi1 = nullptr;       // [2][L][1536] layer-1 input projections in flight
    float* v_last = nullptr;       // [L][512]
    float* gi = nullptr;           // [L][1536] input projections of the current bi-GRU layer
    float* seq_a = nullptr;        // [L][520] layer input / output ping
    float* seq_b = nullptr;        // [L][512] pong
    float* mat1d_t = nullptr;      // [L][512] hgru output (time-major == mat1d transposed)
    // 2-D track
    float* dmap = nullptr;         // [L*L]
    float* base384 = nullptr;      // [L*L][384] cached stem pre-activation without the dmap term
    float* raw = nullptr;          // [L*L][128] maxout output before InstanceNorm
    __half* tc_scratch = nullptr;  // operand staging of the tensor-core GEMMs: 4 x [DMP2_TC_SLAB][1024] fp16
    __half* dca_tc = nullptr;      // fp16 hi/lo operands of the MSA-feature GEMMs (sized for (L, N))
    float* tc_scal = nullptr;      // [4][4] device operand scales {2^k, 2^-k, scratch, -}
    float* x = nullptr;            // [L*L][128] residual stream fp32 (NHWC)
    __half* xh = nullptr;          // [L*L][128] fp16 high part of x
    __half* xl = nullptr;          // [L*L][128] fp16 low part
    uint8_t* x8lo = nullptr;       // [L*L][128] e4m3 of x_lo * 2^8
    uint8_t* x8hi = nullptr;       // [L*L][128] e4m3 of x_hi * 2^-4
    double* stat_part = nullptr;   // [nparts][256] partial sums
    float* norm_ss = nullptr;      // [256] per-channel scale, shift
    unsigned int* ticket = nullptr;
    unsigned long long* sched = nullptr;       // unit counter of the dynamic conv schedule
    float* head = nullptr;         // [2][L*L]
    float* conf = nullptr;         // [L]
    float* mmat = nullptr;         // [L*L]
    double* eig_a = nullptr;       // [L*L] working matrix (fp64)
    double* eig_w = nullptr;       // small vectors: d, e, beta, z[8][L], ...
    float* eig_val = nullptr;      // [8]
    float* mds = nullptr;          // [L][8]
    float* ca = nullptr;           // [L][3]
    float* best_ca = nullptr;      // [L][3]
    float* best_conf = nullptr;    // [L]
    float* best_mean = nullptr;    // [1]
    float* coords_out = nullptr;   // [L][5][3]
    float* conf_out = nullptr;     // [L]
    int rows2d = 0;                // rows of the 2-D track held here (L, or a strip height in halo-sharded mode)
    bool full_act = false;         // xh/xl/x8lo/x8hi hold a whole L x L image (false: halo-sharded, they live in the window)
    std::vector<void*> allocs;
};

// ---------------------------------------------------------------------------------------------------
// Halo-sharded fold of ONE target over `world` GPUs (strip.cu): every L x L map is split into row strips, rank g
// owns image rows [r0, r1).  The tensors other ranks write into live in one IPC-shared window per rank with the
// same layout everywhere; exchanges are peer stores over NVLink followed by an epoch flag, consumers spin on
// their own flags -- no host synchronisation and no collective library on the data path.
// ---------------------------------------------------------------------------------------------------
enum { STRIP_FLAG_STATS = 0, STRIP_FLAG_HALO = 1, STRIP_FLAG_HEAD = 2, STRIP_FLAG_VGRU = 3, STRIP_FLAG_X3 = 4, STRIP_NFLAGS = 5 };
struct StripCtx {
    int rank = 0, world = 1;
    int L = 0, rows_per = 0, r0 = 0, r1 = 0;
    bool attached = false, ipc = false;
    uint8_t* win = nullptr;                    // this rank's window
    uint8_t* peer[DMP2_MAX_RANKS] = {};        // every rank's window as mapped here (peer[rank] == win)
    size_t win_bytes = 0;
    size_t off_act[4] = {0, 0, 0, 0};          // xh, xl, x8lo, x8hi: [2 halo rows | rows_per | 2 halo rows][L][128]
    size_t off_stats = 0;                      // [2 parity][world][256] double: per-rank InstanceNorm partial sums
    size_t off_head = 0;                       // [2][L][L] float: every rank's head strip lands here
    size_t off_x3 = 0;                         // [L][L] float: DCA contact norms (APC needs every row's sums)
    size_t off_vlast = 0;                      // [L][512] float: vgru output, every rank computes a range of columns
    size_t off_flags = 0;                      // [STRIP_NFLAGS][DMP2_MAX_RANKS] uint32 epochs, indexed by source rank
    uint32_t epoch[STRIP_NFLAGS] = {0, 0, 0, 0, 0};
    unsigned int* ticket = nullptr;            // local: last-CTA detection of the push kernel ([0] main stream, [1] side stream)
    double* totals = nullptr;                  // local: [256] this rank's partial sums of the current map
};

struct dmp2_engine {
    int device = 0;
    int num_sms = 148;
    int conv_mode = DMP2_CONV_TC_F16F8;   // default; DMP2_CONV_MODE=f16x3 is the reference-accuracy conv (3 MMAs per MAC)
    int64_t launches = 0;
    int status = 0;
    std::string err;
    Weights w;
    std::vector<void*> weight_allocs;
    Workspace ws;
    cudaStream_t side = nullptr;     // MSA features run here, concurrently with the vgru on the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[16];
    bool ev_ok = false;
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void* tc_state = nullptr;        // tensor-core conv state (tensor maps), owned by conv_tc.cu
    void* vt_state = nullptr;        // tensor-core vgru state, owned by vgru_tc.cu
    int conv_cluster = 0;            // conv kernel form: 0 = cta_group::2 CTA pairs, 1 = independent CTAs, 2 = 2-CTA weight multicast (DMP2_CONV_CLUSTER=pair|1|2)
    int conv_sms = 0;                // SMs the persistent conv kernel occupies (0 = all)
    const float* vgru_pre = nullptr; // one-shot: the next dmp2_fold takes the vgru state [L][512] from here instead of scanning the MSA
                                     // (dmp2_set_vgru_input; a throughput scheduler scans the columns of several targets in ONE vgru call)
    bool conv_dynamic = false;       // conv units claimed from a global counter instead of the static round-robin (throughput mode:
                                     // several folds in flight on one GPU; dmp2_set_conv_dynamic / DMP2_CONV_DYNAMIC=1)
    unsigned long long conv_sched_base = 0;      // first counter value of the next dynamic conv launch (ws.sched is never reset)
    int conv_chunk_taps = 0;         // taps per tcgen05 accumulation chain (1, 5 or 25; DMP2_CONV_CHUNK); 0 = per mode: 1 for f16x3 (whose
                                     // operands are exact to 2^-22, so the chain length IS its error), 5 for f16f8 / f16 (operand error 6e-6 / 1e-4
                                     // dwarfs the 6e-7 of a 5-tap chain; 2-3 % faster)
    int vgru_mode = 0;               // 0 = tensor cores, one launch per MSA row; 1 = CUDA-core fp32 validation path
    bool eig_no_cl16 = false;        // set when a 16-CTA cluster launch was refused
    bool attr_eig = false, attr_refine = false, attr_eig_grid = false;   // per-engine (= per-device) cudaFuncSetAttribute done
    StripCtx sp;                     // halo-sharded mode: window + peers (strip.cu)
    bool gemm_tc = true;             // stem GEMM on the tcgen05 pipeline (DMP2_GEMM=ffma: CUDA-core validation path)
    bool fuse_stats = true;          // InstanceNorm sums come out of the conv epilogue (DMP2_FUSE_STATS=0: separate k_in_stats pass)
    bool strip_on = false;           // true while dmp2_fold_strip runs: the 2-D track works on rows [sp.r0, sp.r1)
    bool profile = false;            // record a CUDA-event pair around every conv launch (bench.py roofline)
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    // CUDA-graph replay of the recycling iterations (network.py:264-306; dmp2_set_graph / DMP2_GRAPH=1): iteration = distance
    // map -> ResNet pass -> head -> eigen step -> coordinate GRU -> best-of-n select is the same ~50 launches on the same
    // buffers every time, so it is captured once per (L, workspace, kernel configuration) and replayed `iterations` times.
    bool graph_on = false;
    bool capturing = false;          // true while the iteration is being captured (conv launcher: self-resetting unit counter)
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t pass_exec = nullptr;
    int64_t pass_nodes = 0;          // kernel launches one replay stands for (gpu_launches bookkeeping)
    uint64_t ws_gen = 0;             // bumped whenever the workspace is (re)allocated: every captured pointer is stale then
    struct PassKey {
        int L = 0, conv_mode = -1, conv_cluster = -1, conv_chunk_taps = -1, conv_sms = -1;
        bool conv_dynamic = false, gemm_tc = false, fuse_stats = false, eig_no_cl16 = false, profile = false;
        uint64_t ws_gen = 0;
        bool operator==(const PassKey& o) const {
            return L == o.L && conv_mode == o.conv_mode && conv_cluster == o.conv_cluster && conv_chunk_taps == o.conv_chunk_taps &&
                   conv_sms == o.conv_sms && conv_dynamic == o.conv_dynamic && gemm_tc == o.gemm_tc && fuse_stats == o.fuse_stats &&
                   eig_no_cl16 == o.eig_no_cl16 && profile == o.profile && ws_gen == o.ws_gen;
        }
    } pass_key;
    std::vector<cudaEvent_t> gprof_ev;   // event pairs recorded by the graph's conv nodes when profiling (2 per block)
    bool gprof_valid = false;            // the graph has been replayed since dmp2_conv_profile last looked

    int fail(int code, const std::string& msg) {
        status = code;
        err = msg;
        return code;
    }
};

// ---------------------------------------------------------------------------------------------------
// Stage launchers (one per .cu file).  All asynchronous on `st`.
// ---------------------------------------------------------------------------------------------------
// msa.cu
int run_reweight(dmp2_engine* e, const uint8_t* msa, int N, int L, float* w_out, cudaStream_t st);
int run_dca(dmp2_engine* e, const uint8_t* msa, int N, int L, const float* w, float* feat444, cudaStream_t st);
int run_feat_export(dmp2_engine* e, const float* feat444, int L, float* feat442, cudaStream_t st);
int run_feat_import(dmp2_engine* e, const float* feat442, int L, float* feat444, cudaStream_t st);
// gru.cu
int run_vgru(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st);
int run_vgru_ffma(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st);
int run_vgru_tc(dmp2_engine* e, const uint8_t* msa, int N, int L, float* out, cudaStream_t st, int ld = 0);   // ld: row stride of msa (0 = L)
void vgru_tc_destroy(dmp2_engine* e);
void vgru_tc_invalidate(dmp2_engine* e);
int run_bigru(dmp2_engine* e, const BiGruLayer* layers, int nlayers, const float* in, int L, float* out, cudaStream_t st);
int run_coord_head(dmp2_engine* e, const float* mat1d_t, const float* mds, int L, float* ca, cudaStream_t st);
// resnet.cu
int run_stem_base(dmp2_engine* e, const float* mat1d_t, const float* feat444, int L, cudaStream_t st);
int run_stem_update(dmp2_engine* e, const float* dmap, int L, cudaStream_t st);   // -> ws.x (+ xh, xl)
int run_conv_ffma(dmp2_engine* e, int blk, const float* x, int L, float* raw, cudaStream_t st);
int run_norm_gate(dmp2_engine* e, int blk, const float* raw, float* x, int L, bool stem, cudaStream_t st, bool have_stats = false);
int run_split_half(dmp2_engine* e, const float* x, int64_t n, __half* hi, __half* lo, uint8_t* x8lo, uint8_t* x8hi, cudaStream_t st);
int run_resblock(dmp2_engine* e, int blk, int L, cudaStream_t st);                // ws.x -> ws.x
int run_head(dmp2_engine* e, const float* x, int L, float* head2, cudaStream_t st);
int run_head_post(dmp2_engine* e, const float* head2, int L, float* conf, float* mmat, cudaStream_t st);
// conv_tc.cu
int run_conv_tc(dmp2_engine* e, int blk, const __half* xh, const __half* xl, const uint8_t* x8lo, const uint8_t* x8hi, int L,
                int H, int y_off, int map_rows, float* raw, int mode, cudaStream_t st, bool fuse_stats = false);
bool conv_tc_fuses_stats(const dmp2_engine* e);      // true: run_conv_tc(..., fuse_stats = true) leaves ws.norm_ss / sp.totals ready
int run_gemm_tn_test(dmp2_engine* e, const float* a, const float* b, int M, int N, int K, int mode, int chunk_k, float* c, cudaStream_t st);
void conv_tc_invalidate(dmp2_engine* e);
// fp32 GEMMs on the tensor-core pipeline: scaled fp16 hi/lo split of an operand, and C = alpha * A B^T from split operands
struct GemmTcEpilogue {               // epilogue variants of the MSA-feature GEMMs (see TcParams in conv_tc.cu)
    int kind;                         // 0 plain, 1 Gram, 2 Woodbury, 3 covariance
    int m_off;
    const float* dsa;                 // device {scale, 1/scale} of the A / B operand split (nullptr = 1)
    const float* dsb;
    const float* scal;                // device scalars of predict.py:45-51: [1] = n_eff, [2] = ridge
    const float* bias = nullptr;      // kind 0 only: per-column bias [N] added after the scaling
};
int run_operand_scale(dmp2_engine* e, const float* x, int rows, int cols, int64_t ld, float* ds /* 3 floats */, cudaStream_t st);
int run_split_scaled(dmp2_engine* e, const float* x, int rows, int cols, int64_t ld, float scale, const float* dscale, __half* hi,
                     __half* lo, int Kp, cudaStream_t st);
int run_gemm_tc(dmp2_engine* e, const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo, int M, int N, int Kp,
                float alpha, float* c, int ldc, int chunk_k, cudaStream_t st, const GemmTcEpilogue* ep = nullptr, int b_rows = 0);             // forget cached activation tensor maps (their buffers are being freed)
void conv_tc_destroy(dmp2_engine* e);
// eig.cu
int run_eig_top8(dmp2_engine* e, const float* m, int L, float* vals, float* mds_scaled, float* vecs_raw, cudaStream_t st);
// geom.cu
int run_dmap(dmp2_engine* e, const float* ca, int L, float* dmap, bool clamp, cudaStream_t st);
int run_fill(dmp2_engine* e, float* p, int64_t n, float v, cudaStream_t st);
int run_refine(dmp2_engine* e, float* ca, int L, int steps, cudaStream_t st);
int run_backbone(dmp2_engine* e, const float* ca, const float* conf_logit, int L, float* out, float* conf_out, cudaStream_t st);
int run_select(dmp2_engine* e, const float* ca, const float* conf, int L, int first, cudaStream_t st);

// strip.cu (halo-sharded mode)
int strip_rows(int L, int world, int rank, int* r0, int* r1, int* rows_per);
int strip_setup(dmp2_engine* e, int rank, int world, int L, unsigned char* handle_out);
int strip_attach(dmp2_engine* e, const unsigned char* handles, void* const* ptrs);
int strip_detach(dmp2_engine* e);
int strip_stats_exchange(dmp2_engine* e, const float* gamma, float* norm_ss, cudaStream_t st);   // sp.totals -> all ranks -> norm_ss
int strip_halo_push(dmp2_engine* e, cudaStream_t st);
int strip_halo_wait(dmp2_engine* e, cudaStream_t st);
int strip_head_gather(dmp2_engine* e, cudaStream_t st);
int strip_x3_gather(dmp2_engine* e, cudaStream_t st);                       // my rows of x3 -> every rank (side stream)
int strip_vgru_gather(dmp2_engine* e, int c0, int c1, cudaStream_t st);      // columns [c0, c1) of v_last -> every rank
// rows of the 2-D track this engine works on, and where its conv activations live
struct Rows { int r0, R; };
static inline Rows rows_of(const dmp2_engine* e, int L) {
    return e->strip_on ? Rows{e->sp.r0, e->sp.r1 - e->sp.r0} : Rows{0, L};
}
struct ActPtrs { __half* xh; __half* xl; uint8_t* x8lo; uint8_t* x8hi; };   // first INTERIOR pixel of every copy
static inline ActPtrs act_of(dmp2_engine* e, int L) {
    if (!e->strip_on) return ActPtrs{e->ws.xh, e->ws.xl, e->ws.x8lo, e->ws.x8hi};
    const int64_t skip = 2 * (int64_t)L * 128;
    uint8_t* w = e->sp.win;
    return ActPtrs{reinterpret_cast<__half*>(w + e->sp.off_act[0]) + skip, reinterpret_cast<__half*>(w + e->sp.off_act[1]) + skip,
                   w + e->sp.off_act[2] + skip, w + e->sp.off_act[3] + skip};
}
static inline float* head_of(dmp2_engine* e) {
    return e->strip_on ? reinterpret_cast<float*>(e->sp.win + e->sp.off_head) : e->ws.head;
}

// workspace: rows2d = rows of the 2-D track held locally (L, or the strip height in halo-sharded mode)
int ensure_workspace(dmp2_engine* e, int L, int N, int rows2d = -1);
