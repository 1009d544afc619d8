/*
 * dmp2.h -- C ABI of libdmp2.so, the B200-native (sm_100a) DMPfold2 inference engine.
 *
 * The reference (psipred/DMPfold2) has no FFI/plugin layer: its boundary is the Python call
 *     network(inputs, inputs2, nloops, refine_steps)            dmpfold/predict.py:151, network.py:218
 * made by aln_to_coords()                                       dmpfold/predict.py:74-158
 * plus the state_dict key/shape schema it loads                 dmpfold/predict.py:83-98.
 * This header is what a ctypes binding on the reference side would bind instead of that call; see
 * INTEGRATION.md for the stub.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative dmp2_status; nothing throws across the ABI;
 *     dmp2_last_error() gives the message for the last failure on that engine (or the global one for
 *     dmp2_create failures when engine == NULL).
 *   - "dev" pointers are device memory on the engine's CUDA device, "host" pointers are CPU memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Entry points that
 *     take device pointers are asynchronous on that stream and never synchronise the host; the caller
 *     keeps the buffers alive until the stream is synchronised.
 *   - an engine is bound to one device and is not re-entrant: one engine per (device, stream).
 */
#ifndef DMP2_H
#define DMP2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmp2_engine dmp2_engine;

#define DMP2_MAX_RANKS 8          /* GPUs one halo-sharded fold can span (one NVSwitch box) */
#define DMP2_IPC_HANDLE_BYTES 64  /* sizeof(cudaIpcMemHandle_t) */

typedef enum {
    DMP2_OK = 0,
    DMP2_ERR_BAD_ARG = -1,        /* NULL pointer, L < 8, N < 1, negative counts ...            */
    DMP2_ERR_MISSING_WEIGHT = -2, /* a state_dict key of predict.py:98 is absent or mis-sized   */
    DMP2_ERR_CUDA = -3,           /* a CUDA runtime/driver call or kernel launch failed         */
    DMP2_ERR_OOM = -4,            /* device allocation failed                                   */
    DMP2_ERR_NO_DEVICE = -5,      /* no usable sm_100 device: there is NO CPU fallback          */
    DMP2_ERR_UNSUPPORTED = -6
} dmp2_status;

/* Convolution arithmetic of the sixteen 5x5 ResNet blocks (network.py:26 inside ResNet_Block). */
typedef enum {
    DMP2_CONV_TC_F16X3 = 0, /* tcgen05, fp16 hi+lo split of both operands, 3 MMAs per MAC, per-tap accumulation chains
                               summed in fp32 registers: per-block error equal to an fp32 CPU conv's            */
    DMP2_CONV_TC_F16 = 1,   /* tcgen05, single fp16 MMA (informational; ~1e-4 relative operand error, fails parity) */
    DMP2_CONV_FFMA = 2,     /* CUDA-core fp32 implicit GEMM (validation path for the tensor-core kernels)        */
    DMP2_CONV_TC_F16F8 = 3  /* tcgen05, fp16 main term + the two hi/lo correction terms in FP8 (e4m3 x e5m2): 2 MMA-
                               equivalents per MAC, same per-tap chains; the default -- whole folds end closer to the
                               fp64 evaluation than the fp32 reference does (tests/test_gpu_parity_r2.py)        */
} dmp2_conv_mode;

/* ---- lifetime ------------------------------------------------------------------------------------ */

/* Build an engine on CUDA device `device` from the reference's state_dict (predict.py:89-98): `names[i]`
 * is the state_dict key, `host_ptrs[i]` a contiguous fp32 host array of `numels[i]` elements.  All 184
 * keys of the published model are required (strict, like load_state_dict).  The engine uploads and
 * repacks the weights into its own layouts and owns the copies. */
int dmp2_create(dmp2_engine** out, int device, int n_tensors, const char* const* names,
                const float* const* host_ptrs, const int64_t* numels);
void dmp2_destroy(dmp2_engine* e);
const char* dmp2_last_error(const dmp2_engine* e);
int dmp2_set_conv_mode(dmp2_engine* e, int mode /* dmp2_conv_mode */);
/* Number of this library's kernels launched by the engine since creation (bench.py's gpu_launches). */
int64_t dmp2_launch_count(const dmp2_engine* e);
/* Per-stage device time of the last dmp2_fold_host call, in ms (CUDA events): out[0..n) in the order
 * vgru, hgru, MSA features not hidden behind the vgru (they run on a side stream), stem base, all recycling
 * passes, final refine + backbone.  Returns n (6). */
int dmp2_stage_times(const dmp2_engine* e, float* out_ms, int cap);

/* Debug: device time (us) of the four phases of the last top-8 eigensolve at size L (tridiagonalisation,
 * bisection, inverse iteration, back-transform); synchronises the device. */
int dmp2_debug_eig_phases(dmp2_engine* e, int L, double* out_us, int cap);

/* CUDA-graph replay of the recycling iterations (network.py:264-306): when on, one iteration (distance map -> ResNet
 * pass -> head -> eigen step -> coordinate GRU -> best-of-n select, ~50 launches on fixed workspace buffers) is captured
 * once per (L, workspace, kernel configuration) and replayed `iterations` times; results are bit-identical to the eager
 * loop.  Used for L <= 600 and iterations >= 2; falls back to eager launches where capture is refused.
 * Off by default (environment: DMP2_GRAPH=1). */
int dmp2_set_graph(dmp2_engine* e, int on);

/* Roofline support: when on, a CUDA-event pair is recorded (on the launching stream) around every 5x5-conv
 * kernel launch; dmp2_conv_profile synchronises the device, returns the number of launches seen and their
 * summed device time since the last call, and resets the counters. */
int dmp2_set_profile(dmp2_engine* e, int on);
int dmp2_conv_profile(dmp2_engine* e, int* n_launches, float* total_ms);

/* ---- the hot path: replaces network(inputs, inputs2, nloops, refine_steps) + predict.py:136-147 ---- */

/* msa_dev: uint8 N x L row-major residue codes 0..21 as produced by predict.py:124-128 (row 0 = query).
 * tmpl_ca_dev: L x 3 template CA coordinates (predict.py:106-119) or NULL (dmap channel = -1).
 * coords_out_dev: L x 5 x 3 fp32 (N, CA, C, O, CB per residue; predict.py:152); conf_out_dev: L fp32. */
int dmp2_fold(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, const float* tmpl_ca_dev, int iterations,
              int minsteps, float* coords_out_dev, float* conf_out_dev, void* stream);

/* Size the engine's workspace for alignments of up to N rows x L columns now, so that later dmp2_fold calls within
 * those bounds allocate nothing.  (A fold that needs a LARGER workspace than any seen before re-allocates, which
 * synchronises the device once; reserve the maximum up front to keep dmp2_fold strictly asynchronous.) */
int dmp2_reserve(dmp2_engine* e, int L, int N);

/* Same with HOST buffers: copies the alignment up, runs the fold, copies coords/conf back and
 * synchronises.  This is the call timed as the end-to-end number. */
int dmp2_fold_host(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                   int minsteps, float* coords_out_host, float* conf_out_host);

/* ---- halo-sharded fold of ONE target over several GPUs (BASELINE.json configs[4]) -------------------
 * The reference has no multi-GPU path (predict.py:74-158 is single-device); this is the form its
 * network(...) call (predict.py:151) takes when the L x L pair maps are split into row strips over `world`
 * GPUs, one process (or one engine) per GPU.  Rank g owns rows [r0, r1) of every map of the 2-D track
 * (network.py:229-246); conv halo rows, InstanceNorm sums and the head rows move between the ranks' windows
 * by peer stores over NVLink + epoch flags, with no host synchronisation.  The 1-D track, the MSA features,
 * the eigen step, the coordinate GRU and the minimiser are replicated, so every rank returns the full result.
 *
 *   dmp2_strip_rows    pure host arithmetic: the rows of `rank` (multiples of 8; fails if the last rank would
 *                      own fewer than 2 rows)
 *   dmp2_strip_setup   allocates this rank's exchange window for targets of length L (and, if reserve_N > 0, the
 *                      workspace for alignments of up to reserve_N rows, so that the folds allocate nothing);
 *                      writes the window's CUDA IPC handle (64 bytes) and/or its device address
 *   dmp2_strip_attach  maps every rank's window: `handles` = world x 64 bytes in rank order, gathered by the
 *                      caller (e.g. torch.distributed.all_gather_object).  Ranks in separate processes.
 *   dmp2_strip_attach_local   same for ranks living in ONE process: `windows` = world device addresses
 *   dmp2_strip_detach  unmaps and frees; call on every rank (then barrier) before a new setup
 *   dmp2_fold_strip / dmp2_fold_strip_host   like dmp2_fold / dmp2_fold_host, but COLLECTIVE: every rank
 *                      calls with identical arguments.  L must equal the L given to dmp2_strip_setup. */
int dmp2_strip_rows(int L, int world, int rank, int* r0, int* r1);
int dmp2_strip_setup(dmp2_engine* e, int rank, int world, int L, int reserve_N, unsigned char* ipc_handle_out, void** window_out);
int dmp2_strip_attach(dmp2_engine* e, const unsigned char* handles);
int dmp2_strip_attach_local(dmp2_engine* e, void* const* windows);
int dmp2_strip_detach(dmp2_engine* e);
int dmp2_fold_strip(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, const float* tmpl_ca_dev, int iterations,
                    int minsteps, float* coords_out_dev, float* conf_out_dev, void* stream);
int dmp2_fold_strip_host(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                         int minsteps, float* coords_out_host, float* conf_out_host);

/* ---- stage entry points (teacher-forced parity tests; all device pointers, async on `stream`) ------ */

/* predict.py:32-37  reweight: w (N) */
int dmp2_reweight(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* w_out_dev, void* stream);
/* predict.py:41-61  fast_dca: feat (L, L, 442) row-major exactly like the reference tensor */
int dmp2_dca(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* feat_out_dev, void* stream);
/* network.py:223-225  embed + vgru: final hidden state per column, (L, 512) */
int dmp2_vgru(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* out_dev, void* stream);
/* network.py:225-226  hgru: in (L, 512) -> out (L, 512) [fwd | bwd] */
int dmp2_hgru(dmp2_engine* e, const float* in_dev, int L, float* out_dev, void* stream);
/* network.py:26 + :30-31 of block `block` (1..16): x (L*L, 128) NHWC fp32 -> max over 4 consecutive conv
 * channels of (conv5x5 + bias), (L*L, 128) NHWC fp32, before the InstanceNorm. */
int dmp2_conv5_maxout(dmp2_engine* e, int block, const float* x_nhwc_dev, int L, float* out_nhwc_dev, void* stream);
/* network.py:94-103  one full ResNet_Block: x (L*L,128) NHWC -> (L*L,128) NHWC */
int dmp2_resblock(dmp2_engine* e, int block, const float* x_nhwc_dev, int L, float* out_nhwc_dev, void* stream);
/* network.py:194 (resnet.0 = Maxout2d(955 -> 128, pool 3, k=1), network.py:25-34) on the concatenated input of
 * network.py:227-229, which is never materialised: mat1d_t (L,512) time-major hgru output, feat (L,L,442), dmap (L,L)
 * -> the stem output (L*L, 128) NHWC fp32 = the input of ResNet block 1. */
int dmp2_stem(dmp2_engine* e, const float* mat1d_t_dev, const float* feat_dev, const float* dmap_dev, int L,
              float* out_nhwc_dev, void* stream);
/* network.py:207 (resnet.17, the final 1x1 conv 128 -> 2): x (L*L, 128) NHWC fp32 -> head (2, L, L) */
int dmp2_head(dmp2_engine* e, const float* x_nhwc_dev, int L, float* head_out_dev, void* stream);
/* network.py:229-235  one ResNet pass from its inputs: mat1d_t (L,512) time-major hgru output, feat (L,L,442),
 * dmap (L,L) -> head (2, L, L) like resnet.17's output. */
int dmp2_resnet_pass(dmp2_engine* e, const float* mat1d_t_dev, const float* feat_dev, const float* dmap_dev, int L,
                     float* head_out_dev, void* stream);
/* network.py:237-250  head (2,L,L) -> conf (L), M (L,L), mds (L,8) (ascending eigenvalue order, canonical sign) */
int dmp2_head_mds(dmp2_engine* e, const float* head_dev, int L, float* conf_out_dev, float* m_out_dev,
                  float* mds_out_dev, void* stream);
/* top-8 eigenpairs of a symmetric L x L fp32 matrix: vals (8) ascending, vecs (L,8), canonical sign */
int dmp2_eig_top8(dmp2_engine* e, const float* m_dev, int L, float* vals_out_dev, float* vecs_out_dev, void* stream);
/* network.py:251-255  coord_gru + coord_fc: mat1d_t (L,512), mds (L,8) -> CA (L,3) */
int dmp2_coord_gru(dmp2_engine* e, const float* mat1d_t_dev, const float* mds_dev, int L, float* ca_out_dev, void* stream);
/* network.py:106-137  refine_coords, in place on ca (L,3) */
int dmp2_refine(dmp2_engine* e, float* ca_dev, int L, int steps, void* stream);
/* network.py:141-177  calpha_to_main_chain: ca (L,3) -> (L,5,3) */
int dmp2_backbone(dmp2_engine* e, const float* ca_dev, int L, float* out_dev, void* stream);

/* Generic fp32 GEMM self-test hook for the tensor-core GEMM core: C[M,N] = A[M,K] * B[N,K]^T through the
 * same tcgen05 pipeline the conv uses (mode as dmp2_conv_mode: f16 or f16x3).  K % 64 == 0, N % 4 == 0.
 * chunk_k = length of one tcgen05 accumulation chain (a multiple of 64 dividing K; 0 = the whole K): the chains are
 * summed in fp32 registers with round-to-nearest, exactly as the conv does per tap. */
int dmp2_gemm_tn_test(dmp2_engine* e, const float* a_dev, const float* b_dev, int M, int N, int K, int mode, int chunk_k,
                      float* c_dev, void* stream);

/* Tuning: number of SMs the persistent conv kernel occupies (0 = all).  A throughput scheduler that folds several
 * targets on several streams leaves a few SMs to the latency-bound kernels of the other targets. */
int dmp2_set_conv_sms(dmp2_engine* e, int sms);

/* Throughput mode: hand the NEXT dmp2_fold / dmp2_fold_host call of this engine the vgru state of its alignment
 * ([L][512] fp32 on the device, what dmp2_vgru returns) instead of letting it scan the MSA.  The scan (network.py:223-224)
 * is independent per alignment column and latency-bound (N dependent steps that each occupy ~all SMs mostly waiting), so a
 * scheduler folding several alignments with the same N calls dmp2_vgru ONCE on their columns side by side
 * ([N][L1 + L2 + ...]) and gives every fold its slice -- bit-identical to the fold's own scan.  One-shot: cleared by the
 * fold that consumes it; the buffer must stay valid until that fold has finished on its stream.  NULL cancels. */
int dmp2_set_vgru_input(dmp2_engine* e, const float* vgru_dev);

/* Tuning: dynamic unit schedule of the persistent conv kernel (on != 0).  The CTA clusters claim their work units from a
 * device counter instead of the static round-robin, so a conv launch that finds part of the SMs occupied by the kernels
 * of another stream is finished by the CTAs that did start.  Meant for several folds in flight on one GPU
 * (parallel.StreamPool turns it on); the conv output is bit-identical, the fused InstanceNorm sums are fp64 sums formed
 * in a run-dependent order (differences at the 1e-16 level before rounding to fp32). */
int dmp2_set_conv_dynamic(dmp2_engine* e, int on);

#ifdef __cplusplus
}
#endif
#endif /* DMP2_H */
