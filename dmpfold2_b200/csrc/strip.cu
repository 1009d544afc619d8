// Halo-sharded fold of one target over several GPUs (BASELINE.json configs[4]).
//
// Every L x L map of the 2-D track (network.py:229-246) is split into row strips, rank g owns image rows
// [g*rows_per, min(L, (g+1)*rows_per)), rows_per a multiple of the conv tile height.  What crosses GPUs:
//   * 2 halo rows per side of every 5x5 conv input (network.py:26, padding 2), in the operand formats the conv reads;
//   * the per-channel InstanceNorm sums (network.py:32 normalises over the WHOLE map): 256 doubles per rank;
//   * the two head channels (network.py:237-246 needs dm^T, the eigen step needs all of M): each strip to everyone.
// Everything else is either local to a strip or cheap and replicated.  The exchanged tensors live in one window per
// rank (cudaMalloc, exported through CUDA IPC, same layout everywhere).  A transfer is a kernel that stores straight
// into the peers' windows over NVLink, fences at system scope, and then raises an epoch flag in each destination
// window; consumers wait on flags in their OWN memory.  No host synchronisation, no collective library.
#include "common.cuh"

namespace {

constexpr int PUSH_MAX_SEG = 16;
struct PushSeg { const uint8_t* src; uint8_t* dst; unsigned long long bytes; };   // 16-byte aligned, bytes % 16 == 0
struct PushArgs {
    PushSeg seg[PUSH_MAX_SEG];
    int nseg;
    uint32_t* flag[DMP2_MAX_RANKS];      // flag words (in the destination windows) to raise when all data has landed
    int nflag;
    uint32_t epoch;
    int words;                           // 1: segments are only 4-byte aligned, copy word by word
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait until the flag reaches `epoch`.  Bounded: a rank that never arrives must not hang the box.
__device__ __forceinline__ void wait_epoch(const uint32_t* flag, uint32_t epoch) {
    if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const unsigned long long t0 = gtimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(200);
        if (gtimer_ns() - t0 > 30ull * 1000000000ull) {
            printf("strip: wait for a peer timed out (flag word %d of this window: want epoch %u, have %u)\n",
                   (int)(((unsigned long long)flag >> 2) & 31), epoch, ld_acquire_sys(flag));
            __trap();
        }
    }
}

// grid (x, nseg): copy every segment into the peer windows, then the last CTA to finish raises the flags
__global__ void __launch_bounds__(256) k_push(const PushArgs a, unsigned int* __restrict__ ticket) {
    const PushSeg sg = a.seg[blockIdx.y];
    const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, dt = (unsigned long long)gridDim.x * blockDim.x;
    if (a.words) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sg.src);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sg.dst);
        for (unsigned long long i = t0; i < (sg.bytes >> 2); i += dt) dst[i] = src[i];
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(sg.src);
        uint4* dst = reinterpret_cast<uint4*>(sg.dst);
        for (unsigned long long i = t0; i < (sg.bytes >> 4); i += dt) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(ticket, 1u) == total - 1) {
            *ticket = 0;
            __threadfence_system();
            for (int f = 0; f < a.nflag; f++) st_release_sys(a.flag[f], a.epoch);
        }
    }
}

__global__ void k_wait(const uint32_t* __restrict__ flags, uint32_t mask, uint32_t epoch) {
    if (threadIdx.x < DMP2_MAX_RANKS && ((mask >> threadIdx.x) & 1u)) wait_epoch(flags + threadIdx.x, epoch);
}

// fold the per-rank sums in rank order (identical on every rank) and finish the InstanceNorm statistics
__global__ void __launch_bounds__(128) k_stats_finalize(const uint32_t* __restrict__ flags, int world, uint32_t epoch,
                                                        const double* __restrict__ slots, double npix_total,
                                                        const float* __restrict__ gamma, float* __restrict__ norm) {
    if ((int)threadIdx.x < world) wait_epoch(flags + threadIdx.x, epoch);
    __syncthreads();
    const int c = threadIdx.x;
    double s = 0, ss = 0;
    for (int g = 0; g < world; g++) {
        s += __ldcg(slots + g * 256 + c);
        ss += __ldcg(slots + g * 256 + 128 + c);
    }
    double mean = s / npix_total;
    double var = ss / npix_total - mean * mean;
    if (var < 0) var = 0;
    norm[c] = (float)mean;
    norm[128 + c] = (float)((double)gamma[c] / sqrt(var + 1e-5));
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

uint32_t* flag_ptr(const StripCtx& sp, int dst_rank, int kind, int src_rank) {
    return reinterpret_cast<uint32_t*>(sp.peer[dst_rank] + sp.off_flags) + kind * DMP2_MAX_RANKS + src_rank;
}

int launch_push(dmp2_engine* e, PushArgs& a, cudaStream_t st, int ticket = 0) {
    unsigned long long maxb = 0;
    for (int i = 0; i < a.nseg; i++) maxb = std::max(maxb, a.seg[i].bytes);
    int gx = (int)std::min<unsigned long long>(std::max<unsigned long long>(maxb / (16 * 256 * 4), 1), 32);
    k_push<<<dim3(gx, a.nseg), 256, 0, st>>>(a, e->sp.ticket + ticket);
    POST_LAUNCH(e, "k_push");
    return 0;
}

}  // namespace

// rows [r0, r1) of rank `rank`; fails unless every rank gets at least the 2 rows its neighbours need as halo
int strip_rows(int L, int world, int rank, int* r0, int* r1, int* rows_per) {
    if (L < 8 || world < 1 || world > DMP2_MAX_RANKS || rank < 0 || rank >= world) return DMP2_ERR_BAD_ARG;
    const int per = cdiv(cdiv(L, 8), world) * 8;
    if ((world - 1) * per + (world > 1 ? 2 : 0) > L) return DMP2_ERR_BAD_ARG;      // the last strip would be empty or thinner than a halo
    if (r0) *r0 = rank * per;
    if (r1) *r1 = std::min(L, (rank + 1) * per);
    if (rows_per) *rows_per = per;
    return 0;
}

int strip_setup(dmp2_engine* e, int rank, int world, int L, unsigned char* handle_out) {
    StripCtx& sp = e->sp;
    if (sp.win) return e->fail(DMP2_ERR_BAD_ARG, "strip_setup: a window exists already (call dmp2_strip_detach on every rank first)");
    int r0, r1, per;
    if (strip_rows(L, world, rank, &r0, &r1, &per) != 0)
        return e->fail(DMP2_ERR_BAD_ARG, "strip_setup: need 1 <= world <= 8, 0 <= rank < world and at least 2 rows of the map on the last rank");
    CUDA_TRY(e, cudaSetDevice(e->device));
    sp = StripCtx();
    sp.rank = rank; sp.world = world; sp.L = L; sp.rows_per = per; sp.r0 = r0; sp.r1 = r1;
    const size_t px = (size_t)(per + 4) * L * 128;
    size_t off = 0;
    const size_t esz[4] = {2, 2, 1, 1};
    for (int i = 0; i < 4; i++) { sp.off_act[i] = off; off = align_up(off + px * esz[i], 1024); }
    sp.off_stats = off; off = align_up(off + (size_t)2 * DMP2_MAX_RANKS * 256 * sizeof(double), 1024);
    sp.off_head = off; off = align_up(off + (size_t)2 * L * L * sizeof(float), 1024);
    sp.off_x3 = off; off = align_up(off + (size_t)L * L * sizeof(float), 1024);
    sp.off_vlast = off; off = align_up(off + (size_t)L * 512 * sizeof(float), 1024);
    sp.off_flags = off; off = align_up(off + (size_t)STRIP_NFLAGS * DMP2_MAX_RANKS * sizeof(uint32_t), 1024);
    sp.win_bytes = off;
    CUDA_TRY(e, cudaMalloc(&sp.win, sp.win_bytes));
    CUDA_TRY(e, cudaMemset(sp.win, 0, sp.win_bytes));          // halo rows beyond the image edge stay zero for good
    CUDA_TRY(e, cudaMalloc(&sp.ticket, 256));
    CUDA_TRY(e, cudaMemset(sp.ticket, 0, 256));
    CUDA_TRY(e, cudaMalloc(&sp.totals, 256 * sizeof(double)));
    CUDA_TRY(e, cudaDeviceSynchronize());
    sp.peer[rank] = sp.win;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == DMP2_IPC_HANDLE_BYTES, "IPC handle size");
        CUDA_TRY(e, cudaIpcGetMemHandle(&h, sp.win));
        memcpy(handle_out, &h, sizeof(h));
    }
    return 0;
}

// handles: world x 64 bytes from dmp2_strip_setup on every rank (separate processes), or ptrs: the windows' device
// addresses when all ranks live in this process (tests; several engines on one or more devices).
int strip_attach(dmp2_engine* e, const unsigned char* handles, void* const* ptrs) {
    StripCtx& sp = e->sp;
    if (!sp.win) return e->fail(DMP2_ERR_BAD_ARG, "strip_attach: call dmp2_strip_setup first");
    if (sp.attached) return e->fail(DMP2_ERR_BAD_ARG, "strip_attach: already attached");
    if (!handles && !ptrs) return e->fail(DMP2_ERR_BAD_ARG, "strip_attach: no peer handles");
    CUDA_TRY(e, cudaSetDevice(e->device));
    for (int r = 0; r < sp.world; r++) {
        if (r == sp.rank) continue;
        if (handles) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)r * DMP2_IPC_HANDLE_BYTES, sizeof(h));
            void* p = nullptr;
            CUDA_TRY(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            sp.peer[r] = (uint8_t*)p;
        } else {
            cudaPointerAttributes at;
            CUDA_TRY(e, cudaPointerGetAttributes(&at, ptrs[r]));
            if (at.device != e->device) {
                int can = 0;
                CUDA_TRY(e, cudaDeviceCanAccessPeer(&can, e->device, at.device));
                if (!can) return e->fail(DMP2_ERR_UNSUPPORTED, "strip_attach: no peer access between the devices");
                cudaError_t c = cudaDeviceEnablePeerAccess(at.device, 0);
                if (c != cudaSuccess && c != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e, c);
                cudaGetLastError();
            }
            sp.peer[r] = (uint8_t*)ptrs[r];
        }
    }
    sp.ipc = handles != nullptr;
    sp.attached = true;
    return 0;
}

int strip_detach(dmp2_engine* e) {
    StripCtx& sp = e->sp;
    if (!sp.win) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    conv_tc_invalidate(e);               // the conv's cached tensor maps point into the window
    if (sp.attached && sp.ipc)
        for (int r = 0; r < sp.world; r++)
            if (r != sp.rank && sp.peer[r]) cudaIpcCloseMemHandle(sp.peer[r]);
    cudaFree(sp.win);
    cudaFree(sp.ticket);
    cudaFree(sp.totals);
    sp = StripCtx();
    return 0;
}

// sp.totals (this rank's sums) -> slot [parity][rank] of every window; then wait for all ranks and finish the statistics
int strip_stats_exchange(dmp2_engine* e, const float* gamma, float* norm_ss, cudaStream_t st) {
    StripCtx& sp = e->sp;
    const uint32_t ep = ++sp.epoch[STRIP_FLAG_STATS];
    const size_t slot = sp.off_stats + ((size_t)(ep & 1) * DMP2_MAX_RANKS) * 256 * sizeof(double);
    PushArgs a;
    a.nseg = 0; a.nflag = 0; a.epoch = ep; a.words = 0;
    for (int r = 0; r < sp.world; r++) {
        a.seg[a.nseg++] = {reinterpret_cast<const uint8_t*>(sp.totals), sp.peer[r] + slot + (size_t)sp.rank * 256 * sizeof(double), 256 * sizeof(double)};
        a.flag[a.nflag++] = flag_ptr(sp, r, STRIP_FLAG_STATS, sp.rank);
    }
    TRY(launch_push(e, a, st));
    k_stats_finalize<<<1, 128, 0, st>>>(reinterpret_cast<const uint32_t*>(sp.win + sp.off_flags) + STRIP_FLAG_STATS * DMP2_MAX_RANKS, sp.world, ep,
                                        reinterpret_cast<const double*>(sp.win + slot), (double)sp.L * (double)sp.L, gamma, norm_ss);
    POST_LAUNCH(e, "k_stats_finalize");
    return 0;
}

// my first / last two interior rows -> the bottom / top halo rows of the strips above / below, in every format the conv reads
int strip_halo_push(dmp2_engine* e, cudaStream_t st) {
    StripCtx& sp = e->sp;
    const uint32_t ep = ++sp.epoch[STRIP_FLAG_HALO];
    if (sp.world == 1) return 0;
    const int R = sp.r1 - sp.r0;
    const size_t esz[4] = {2, 2, 1, 1};
    const bool use[4] = {true, e->conv_mode == DMP2_CONV_TC_F16X3, e->conv_mode == DMP2_CONV_TC_F16F8, e->conv_mode == DMP2_CONV_TC_F16F8};
    PushArgs a;
    a.nseg = 0; a.nflag = 0; a.epoch = ep; a.words = 0;
    for (int i = 0; i < 4; i++) {
        if (!use[i]) continue;
        const size_t rowb = (size_t)sp.L * 128 * esz[i];
        const uint8_t* mine = sp.win + sp.off_act[i];
        if (sp.rank > 0)                 // rows 2..3 of my copy -> rows rows_per+2 .. rows_per+3 of the strip above (always a full strip)
            a.seg[a.nseg++] = {mine + 2 * rowb, sp.peer[sp.rank - 1] + sp.off_act[i] + (size_t)(sp.rows_per + 2) * rowb, 2 * rowb};
        if (sp.rank < sp.world - 1)      // my last two interior rows -> rows 0..1 of the strip below
            a.seg[a.nseg++] = {mine + (size_t)R * rowb, sp.peer[sp.rank + 1] + sp.off_act[i], 2 * rowb};
    }
    if (sp.rank > 0) a.flag[a.nflag++] = flag_ptr(sp, sp.rank - 1, STRIP_FLAG_HALO, sp.rank);
    if (sp.rank < sp.world - 1) a.flag[a.nflag++] = flag_ptr(sp, sp.rank + 1, STRIP_FLAG_HALO, sp.rank);
    return launch_push(e, a, st);
}

int strip_halo_wait(dmp2_engine* e, cudaStream_t st) {
    StripCtx& sp = e->sp;
    if (sp.world == 1) return 0;
    uint32_t mask = 0;
    if (sp.rank > 0) mask |= 1u << (sp.rank - 1);
    if (sp.rank < sp.world - 1) mask |= 1u << (sp.rank + 1);
    k_wait<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(sp.win + sp.off_flags) + STRIP_FLAG_HALO * DMP2_MAX_RANKS, mask, sp.epoch[STRIP_FLAG_HALO]);
    POST_LAUNCH(e, "k_wait");
    return 0;
}

// my rows of both head channels -> the same place in every other window; then wait for everybody's rows
int strip_head_gather(dmp2_engine* e, cudaStream_t st) {
    StripCtx& sp = e->sp;
    const uint32_t ep = ++sp.epoch[STRIP_FLAG_HEAD];
    if (sp.world == 1) return 0;
    const size_t first = (size_t)sp.r0 * sp.L * sizeof(float), bytes = (size_t)(sp.r1 - sp.r0) * sp.L * sizeof(float);
    const size_t chan = (size_t)sp.L * sp.L * sizeof(float);
    PushArgs a;
    a.nseg = 0; a.nflag = 0; a.epoch = ep;
    a.words = ((first | bytes | chan | sp.off_head) & 15) ? 1 : 0;      // odd L: rows are not 16-byte aligned
    uint32_t mask = 0;
    for (int r = 0; r < sp.world; r++) {
        if (r == sp.rank) continue;
        for (int c = 0; c < 2; c++)
            a.seg[a.nseg++] = {sp.win + sp.off_head + c * chan + first, sp.peer[r] + sp.off_head + c * chan + first, bytes};
        a.flag[a.nflag++] = flag_ptr(sp, r, STRIP_FLAG_HEAD, sp.rank);
        mask |= 1u << r;
    }
    TRY(launch_push(e, a, st));
    k_wait<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(sp.win + sp.off_flags) + STRIP_FLAG_HEAD * DMP2_MAX_RANKS, mask, ep);
    POST_LAUNCH(e, "k_wait");
    return 0;
}

// vgru is independent per alignment column (network.py:223-224: the MSA columns are its batch): every rank scans a
// range of columns and stores its rows of the (L, 512) result into every window
int strip_vgru_gather(dmp2_engine* e, int c0, int c1, cudaStream_t st) {
    StripCtx& sp = e->sp;
    const uint32_t ep = ++sp.epoch[STRIP_FLAG_VGRU];
    if (sp.world == 1) return 0;
    PushArgs a;
    a.nseg = 0; a.nflag = 0; a.epoch = ep; a.words = 0;
    uint32_t mask = 0;
    const size_t first = sp.off_vlast + (size_t)c0 * 512 * sizeof(float), bytes = (size_t)std::max(c1 - c0, 0) * 512 * sizeof(float);
    for (int r = 0; r < sp.world; r++) {
        if (r == sp.rank) continue;
        if (bytes) a.seg[a.nseg++] = {sp.win + first, sp.peer[r] + first, bytes};
        a.flag[a.nflag++] = flag_ptr(sp, r, STRIP_FLAG_VGRU, sp.rank);
        mask |= 1u << r;
    }
    if (a.nseg == 0) a.seg[a.nseg++] = {sp.win + sp.off_vlast, sp.win + sp.off_vlast, 0};      // nothing to send, flags only
    TRY(launch_push(e, a, st));
    k_wait<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(sp.win + sp.off_flags) + STRIP_FLAG_VGRU * DMP2_MAX_RANKS, mask, ep);
    POST_LAUNCH(e, "k_wait");
    return 0;
}

// DCA tail: every rank forms only its rows of the inverse covariance, hence only its rows of the contact-norm map
// x3; the APC term (predict.py:57-60) needs the row and column sums of the whole map.  Runs on the engine's side
// stream (second ticket word), concurrently with the vgru exchange on the main stream.
int strip_x3_gather(dmp2_engine* e, cudaStream_t st) {
    StripCtx& sp = e->sp;
    const uint32_t ep = ++sp.epoch[STRIP_FLAG_X3];
    if (sp.world == 1) return 0;
    const size_t first = sp.off_x3 + (size_t)sp.r0 * sp.L * sizeof(float), bytes = (size_t)(sp.r1 - sp.r0) * sp.L * sizeof(float);
    PushArgs a;
    a.nseg = 0; a.nflag = 0; a.epoch = ep;
    a.words = ((first | bytes) & 15) ? 1 : 0;
    uint32_t mask = 0;
    for (int r = 0; r < sp.world; r++) {
        if (r == sp.rank) continue;
        a.seg[a.nseg++] = {sp.win + first, sp.peer[r] + first, bytes};
        a.flag[a.nflag++] = flag_ptr(sp, r, STRIP_FLAG_X3, sp.rank);
        mask |= 1u << r;
    }
    TRY(launch_push(e, a, st, 1));
    k_wait<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(sp.win + sp.off_flags) + STRIP_FLAG_X3 * DMP2_MAX_RANKS, mask, ep);
    POST_LAUNCH(e, "k_wait");
    return 0;
}
