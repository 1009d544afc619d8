// Engine lifetime, weight repacking, workspace, the fold orchestration and the C ABI (include/dmp2.h).
#include "common.cuh"
#include <cuda_fp8.h>
#include <math.h>
#include <string.h>
#include <algorithm>

static std::string g_create_error;

// ---------------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------------
struct HostSD {
    std::map<std::string, std::pair<const float*, int64_t>> t;
    std::string missing;
    const float* get(const std::string& k, int64_t numel) {
        auto it = t.find(k);
        if (it == t.end() || it->second.second != numel) {
            if (missing.empty()) missing = k + (it == t.end() ? " (absent)" : " (wrong size)");
            return nullptr;
        }
        return it->second.first;
    }
};

template <class T>
static int upload(dmp2_engine* e, const std::vector<T>& h, T** out) {
    void* d = nullptr;
    CUDA_TRY(e, cudaMalloc(&d, h.size() * sizeof(T)));
    e->weight_allocs.push_back(d);
    CUDA_TRY(e, cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T*>(d);
    return 0;
}
static int upload_raw(dmp2_engine* e, const float* src, int64_t n, float** out) {
    std::vector<float> h(src, src + n);
    return upload(e, h, out);
}

static int pack_bigru(dmp2_engine* e, HostSD& sd, const std::string& prefix, int layer, int K, BiGruLayer* out) {
    const char* suf[2] = {"", "_reverse"};
    std::vector<float> wih((size_t)1536 * K), bih(1536);
    for (int d = 0; d < 2; d++) {
        std::string l = "_l" + std::to_string(layer) + suf[d];
        const float* w_ih = sd.get(prefix + ".weight_ih" + l, (int64_t)768 * K);
        const float* w_hh = sd.get(prefix + ".weight_hh" + l, 768 * 256);
        const float* b_ih = sd.get(prefix + ".bias_ih" + l, 768);
        const float* b_hh = sd.get(prefix + ".bias_hh" + l, 768);
        if (!w_ih || !w_hh || !b_ih || !b_hh) return DMP2_ERR_MISSING_WEIGHT;
        memcpy(wih.data() + (size_t)d * 768 * K, w_ih, sizeof(float) * 768 * K);
        memcpy(bih.data() + d * 768, b_ih, sizeof(float) * 768);
        TRY(upload_raw(e, w_hh, 768 * 256, &out->dir[d].w_hh));
        TRY(upload_raw(e, b_hh, 768, &out->dir[d].b_hh));
    }
    out->K = K;
    TRY(upload(e, wih, &out->w_ih));
    TRY(upload(e, bih, &out->b_ih));
    const int Kp = (K + 63) & ~63;
    std::vector<__half> wh((size_t)1536 * Kp, __float2half_rn(0.f)), wl(wh);
    for (int r = 0; r < 1536; r++)
        for (int k = 0; k < K; k++) {
            const float v = wih[(size_t)r * K + k] * DMP2_GRU_SW;
            const __half h = __float2half_rn(v);
            wh[(size_t)r * Kp + k] = h;
            wl[(size_t)r * Kp + k] = __float2half_rn(v - __half2float(h));
        }
    TRY(upload(e, wh, &out->w_ih_hi));
    TRY(upload(e, wl, &out->w_ih_lo));
    return 0;
}

static int load_weights(dmp2_engine* e, HostSD& sd) {
    Weights& w = e->w;
    // ---- vgru
    {
        const float* wih0 = sd.get("vgru.weight_ih_l0", 1536 * 22);
        const float* whh0 = sd.get("vgru.weight_hh_l0", 1536 * 512);
        const float* bih0 = sd.get("vgru.bias_ih_l0", 1536);
        const float* bhh0 = sd.get("vgru.bias_hh_l0", 1536);
        const float* wih1 = sd.get("vgru.weight_ih_l1", 1536 * 512);
        const float* whh1 = sd.get("vgru.weight_hh_l1", 1536 * 512);
        const float* bih1 = sd.get("vgru.bias_ih_l1", 1536);
        const float* bhh1 = sd.get("vgru.bias_hh_l1", 1536);
        if (!wih0 || !whh0 || !bih0 || !bhh0 || !wih1 || !whh1 || !bih1 || !bhh1) return DMP2_ERR_MISSING_WEIGHT;
        std::vector<float> gi0((size_t)22 * 512 * 4, 0.f), w0((size_t)2048 * 512, 0.f), b0(2048, 0.f),
            w1((size_t)2048 * 1024, 0.f), b1(2048, 0.f);
        for (int code = 0; code < 22; code++)
            for (int j = 0; j < 512; j++)
                for (int g = 0; g < 3; g++)
                    gi0[((size_t)code * 512 + j) * 4 + g] = wih0[(size_t)(g * 512 + j) * 22 + code] + bih0[g * 512 + j];
        for (int j = 0; j < 512; j++) {
            for (int g = 0; g < 3; g++) {
                memcpy(&w0[(size_t)(4 * j + g) * 512], &whh0[(size_t)(g * 512 + j) * 512], 512 * sizeof(float));
                b0[4 * j + g] = bhh0[g * 512 + j];
            }
            for (int g = 0; g < 2; g++) {
                memcpy(&w1[(size_t)(4 * j + g) * 1024], &wih1[(size_t)(g * 512 + j) * 512], 512 * sizeof(float));
                memcpy(&w1[(size_t)(4 * j + g) * 1024 + 512], &whh1[(size_t)(g * 512 + j) * 512], 512 * sizeof(float));
                b1[4 * j + g] = bih1[g * 512 + j] + bhh1[g * 512 + j];
            }
            memcpy(&w1[(size_t)(4 * j + 2) * 1024], &wih1[(size_t)(1024 + j) * 512], 512 * sizeof(float));
            memcpy(&w1[(size_t)(4 * j + 3) * 1024 + 512], &whh1[(size_t)(1024 + j) * 512], 512 * sizeof(float));
            b1[4 * j + 2] = bih1[1024 + j];
            b1[4 * j + 3] = bhh1[1024 + j];
        }
        // tensor-core packing: row (slice*96 + g*32 + ul) <- gate g of hidden unit slice*32 + ul
        const float* mats[3] = {whh0, wih1, whh1};
        const float* biases[3] = {bhh0, bih1, bhh1};
        for (int r = 0; r < 3; r++) {
            std::vector<__half> wh((size_t)1536 * 512), wl((size_t)1536 * 512);
            std::vector<float> bp(1536);
            for (int sl = 0; sl < 16; sl++)
                for (int g = 0; g < 3; g++)
                    for (int ul = 0; ul < 32; ul++) {
                        const int pr = sl * 96 + g * 32 + ul, src = g * 512 + sl * 32 + ul;
                        bp[pr] = biases[r][src];
                        for (int k = 0; k < 512; k++) {
                            float v = mats[r][(size_t)src * 512 + k];
                            __half h = __float2half_rn(v);
                            wh[(size_t)pr * 512 + k] = h;
                            wl[(size_t)pr * 512 + k] = __float2half_rn(v - __half2float(h));
                        }
                    }
            TRY(upload(e, wh, &w.vt_w_hi[r])); TRY(upload(e, wl, &w.vt_w_lo[r])); TRY(upload(e, bp, &w.vt_bias[r]));
        }
        {
            std::vector<float> g0((size_t)22 * 1536);
            for (int code = 0; code < 22; code++)
                for (int sl = 0; sl < 16; sl++)
                    for (int g = 0; g < 3; g++)
                        for (int ul = 0; ul < 32; ul++) {
                            const int src = g * 512 + sl * 32 + ul;
                            g0[(size_t)code * 1536 + sl * 96 + g * 32 + ul] = wih0[(size_t)src * 22 + code] + bih0[src];
                        }
            TRY(upload(e, g0, &w.vt_gi0));
        }
        TRY(upload(e, gi0, &w.vg_gi0)); TRY(upload(e, w0, &w.vg_w0)); TRY(upload(e, b0, &w.vg_b0));
        TRY(upload(e, w1, &w.vg_w1)); TRY(upload(e, b1, &w.vg_b1));
    }
    // ---- hgru / coord_gru
    for (int k = 0; k < 2; k++) TRY(pack_bigru(e, sd, "hgru", k, 512, &w.hgru[k]));
    for (int k = 0; k < 3; k++) TRY(pack_bigru(e, sd, "coord_gru", k, k == 0 ? 520 : 512, &w.cgru[k]));
    {
        const float* fc = sd.get("coord_fc.weight", 3 * 512);
        if (!fc) return DMP2_ERR_MISSING_WEIGHT;
        TRY(upload_raw(e, fc, 3 * 512, &w.coord_fc));
    }
    // ---- stem
    {
        const float* lw = sd.get("resnet.0.lin.weight", 384 * 955);
        const float* lb = sd.get("resnet.0.lin.bias", 384);
        const float* g = sd.get("resnet.0.norm.weight", 128);
        const float* b = sd.get("resnet.0.norm.bias", 128);
        if (!lw || !lb || !g || !b) return DMP2_ERR_MISSING_WEIGHT;
        std::vector<float> sw((size_t)384 * DMP2_STEM_K, 0.f), wd(384);
        for (int o = 0; o < 384; o++) {
            memcpy(&sw[(size_t)o * DMP2_STEM_K], &lw[(size_t)o * 955], 954 * sizeof(float));
            wd[o] = lw[(size_t)o * 955 + 954];
        }
        TRY(upload(e, sw, &w.stem_w)); TRY(upload(e, wd, &w.stem_wd));
        std::vector<__half> swh((size_t)384 * DMP2_STEM_KP, __float2half_rn(0.f)), swl(swh);
        for (int o = 0; o < 384; o++)
            for (int k = 0; k < DMP2_STEM_K; k++) {
                const float v = sw[(size_t)o * DMP2_STEM_K + k] * DMP2_STEM_SW;
                const __half h = __float2half_rn(v);
                swh[(size_t)o * DMP2_STEM_KP + k] = h;
                swl[(size_t)o * DMP2_STEM_KP + k] = __float2half_rn(v - __half2float(h));
            }
        TRY(upload(e, swh, &w.stem_w_hi)); TRY(upload(e, swl, &w.stem_w_lo));
        TRY(upload_raw(e, lb, 384, &w.stem_b)); TRY(upload_raw(e, g, 128, &w.stem_gamma)); TRY(upload_raw(e, b, 128, &w.stem_beta));
    }
    // ---- ResNet blocks
    for (int k = 0; k < DMP2_NBLOCKS; k++) {
        std::string p = "resnet." + std::to_string(k + 1);
        const float* cw = sd.get(p + ".layer1.lin.weight", (int64_t)512 * 128 * 25);
        const float* cb = sd.get(p + ".layer1.lin.bias", 512);
        const float* g = sd.get(p + ".layer1.norm.weight", 128);
        const float* b = sd.get(p + ".layer1.norm.bias", 128);
        const float* f0 = sd.get(p + ".scSE.cSE.fc.0.weight", 8 * 128);
        const float* f2 = sd.get(p + ".scSE.cSE.fc.2.weight", 128 * 8);
        const float* sw = sd.get(p + ".scSE.sSE.conv.weight", 128);
        const float* sb = sd.get(p + ".scSE.sSE.conv.bias", 1);
        if (!cw || !cb || !g || !b || !f0 || !f2 || !sw || !sb) return DMP2_ERR_MISSING_WEIGHT;
        ResBlockW& bw = w.blk[k];
        std::vector<float> wf((size_t)512 * 3200);
        std::vector<__half> wh(wf.size()), wl(wf.size());
        std::vector<uint8_t> w8w(wf.size()), w8l(wf.size());
        for (int o = 0; o < 512; o++)
            for (int c = 0; c < 128; c++)
                for (int t = 0; t < 25; t++) {
                    float v = cw[((size_t)o * 128 + c) * 25 + t];
                    size_t d = (size_t)o * 3200 + (size_t)t * 128 + c;
                    wf[d] = v;
                    __half h = __float2half_rn(v);
                    wh[d] = h;
                    float lo = v - __half2float(h);
                    wl[d] = __float2half_rn(lo);
                    w8w[d] = (uint8_t)__nv_cvt_float_to_fp8(v * (1.0f / 256.0f), __NV_SATFINITE, __NV_E5M2);
                    w8l[d] = (uint8_t)__nv_cvt_float_to_fp8(lo * 16.0f, __NV_SATFINITE, __NV_E5M2);
                }
        // cSE gate: global-average-pool of an affine InstanceNorm output is exactly beta (network.py:32,:50)
        std::vector<float> gate(128);
        float hid[8];
        for (int r = 0; r < 8; r++) {
            float a = 0.f;
            for (int c = 0; c < 128; c++) a += f0[r * 128 + c] * b[c];
            hid[r] = a > 0.f ? a : 0.f;
        }
        for (int c = 0; c < 128; c++) {
            float a = 0.f;
            for (int r = 0; r < 8; r++) a += f2[c * 8 + r] * hid[r];
            gate[c] = 1.0f / (1.0f + expf(-a));
        }
        TRY(upload(e, wf, &bw.w_f32)); TRY(upload(e, wh, &bw.w_hi)); TRY(upload(e, wl, &bw.w_lo));
        TRY(upload(e, w8w, &bw.w8_w)); TRY(upload(e, w8l, &bw.w8_lo));
        TRY(upload_raw(e, cb, 512, &bw.bias)); TRY(upload_raw(e, g, 128, &bw.gamma)); TRY(upload_raw(e, b, 128, &bw.beta));
        TRY(upload(e, gate, &bw.gate_c)); TRY(upload_raw(e, sw, 128, &bw.sse_w));
        bw.sse_b = sb[0];
    }
    // ---- head
    {
        const float* hw = sd.get("resnet.17.weight", 2 * 128);
        const float* hb = sd.get("resnet.17.bias", 2);
        if (!hw || !hb) return DMP2_ERR_MISSING_WEIGHT;
        TRY(upload_raw(e, hw, 256, &w.head_w));
        w.head_b[0] = hb[0]; w.head_b[1] = hb[1];
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------
static void drop_pass_graph(dmp2_engine* e) {
    if (e->pass_exec) cudaGraphExecDestroy(e->pass_exec);
    e->pass_exec = nullptr;
    e->pass_nodes = 0;
    e->gprof_valid = false;
}

static void free_workspace(dmp2_engine* e) {
    conv_tc_invalidate(e);               // cached tensor maps point into the buffers released here
    vgru_tc_invalidate(e);
    drop_pass_graph(e);                  // ... and so does every node of the captured recycling iteration
    e->ws_gen++;
    for (void* p : e->ws.allocs) cudaFree(p);
    e->ws = Workspace();
}

template <class T>
static int wsalloc(dmp2_engine* e, T** p, int64_t count) {
    void* d = nullptr;
    size_t bytes = (size_t)std::max<int64_t>(count, 1) * sizeof(T);
    bytes = (bytes + 255) & ~(size_t)255;
    cudaError_t c = cudaMalloc(&d, bytes);
    if (c != cudaSuccess) return e->fail(DMP2_ERR_OOM, std::string("workspace cudaMalloc: ") + cudaGetErrorString(c));
    e->ws.allocs.push_back(d);
    *p = reinterpret_cast<T*>(d);
    return 0;
}

int ensure_workspace(dmp2_engine* e, int L, int N, int rows2d) {
    Workspace& ws = e->ws;
    if (rows2d < 0) rows2d = L;
    // halo-sharded folds keep only their strip of the 2-D track (and the conv operand copies in the window)
    // (a halo-sharded workspace has no full-size conv operand copies -- they live in the IPC window -- so a whole-image
    //  call must never reuse it, whatever the sizes say)
    const bool want_full = rows2d == L;
    if (ws.L >= L && ws.N >= N && ((int64_t)ws.rows2d * ws.L >= (int64_t)rows2d * L) && (!want_full || ws.full_act)) return 0;
    if (want_full) { L = std::max(L, ws.L); rows2d = L; }
    N = std::max(N, ws.N);
    CUDA_TRY(e, cudaDeviceSynchronize());
    free_workspace(e);
    const int64_t P = (int64_t)L * L, P2 = (int64_t)rows2d * L, PA = want_full ? P : 1, Npad = (N + 3) & ~3, n = 21 * (int64_t)L, npad = (n + 63) & ~(int64_t)63;
    TRY(wsalloc(e, &ws.msa, (int64_t)N * L));
    TRY(wsalloc(e, &ws.msa_t, (int64_t)L * Npad));
    TRY(wsalloc(e, &ws.seqw, N));
    TRY(wsalloc(e, &ws.scal, 8));
    TRY(wsalloc(e, &ws.xc, n * Npad));
    TRY(wsalloc(e, &ws.cov, npad * npad));
    TRY(wsalloc(e, &ws.gj_p, 4096));
    const int64_t Npad64 = (N + 63) & ~(int64_t)63, n4 = (n + 3) & ~(int64_t)3;
    // Woodbury buffers: always present.  The workspace is reused for every later (L', N') <= (L, N), and whether the
    // Woodbury or the direct path runs depends on THAT call's N' vs 21 L', not on the sizes seen here.
    TRY(wsalloc(e, &ws.xct, Npad * n4));
    TRY(wsalloc(e, &ws.kmat, Npad64 * Npad64));
    TRY(wsalloc(e, &ws.wy, Npad * n4));
    TRY(wsalloc(e, &ws.gj_r, 64 * std::max(npad, Npad64)));
    {   // fp16 hi/lo operands of the MSA-feature GEMMs: xc, wy^T [n][Kp2], K^-1 [Npad64][Kp2], xct [Npad][Kp1] (msa.cu)
        const int64_t Kp2 = (Npad + 127) & ~(int64_t)127, Kp1 = (n4 + 127) & ~(int64_t)127;
        TRY(wsalloc(e, &ws.dca_tc, 2 * (2 * n * Kp2 + Npad64 * Kp2 + Npad * Kp1)));
        TRY(wsalloc(e, &ws.tc_scal, 16));
        CUDA_TRY(e, cudaMemset(ws.tc_scal, 0, 16 * sizeof(float)));
    }
    TRY(wsalloc(e, &ws.x3, P));
    TRY(wsalloc(e, &ws.apc, 2 * L + 1));
    TRY(wsalloc(e, &ws.feat, P * DMP2_FEAT_LD));
    TRY(wsalloc(e, &ws.vg_h, 4 * (int64_t)L * 512));
    TRY(wsalloc(e, &ws.vt_h16, 8 * (int64_t)L * 512));
    TRY(wsalloc(e, &ws.vt_gi1, 2 * (int64_t)L * 1536));
    TRY(wsalloc(e, &ws.v_last, (int64_t)L * 512));
    TRY(wsalloc(e, &ws.gi, (int64_t)L * 1536));
    TRY(wsalloc(e, &ws.seq_a, (int64_t)L * 520));
    TRY(wsalloc(e, &ws.seq_b, (int64_t)L * 512));
    TRY(wsalloc(e, &ws.mat1d_t, (int64_t)L * 512));
    TRY(wsalloc(e, &ws.dmap, P));
    TRY(wsalloc(e, &ws.base384, P2 * 384));
    TRY(wsalloc(e, &ws.raw, P2 * 128));
    TRY(wsalloc(e, &ws.tc_scratch, (int64_t)4 * DMP2_TC_SLAB * 1024));
    TRY(wsalloc(e, &ws.x, P2 * 128));
    TRY(wsalloc(e, &ws.xh, PA * 128));
    TRY(wsalloc(e, &ws.xl, PA * 128));
    TRY(wsalloc(e, &ws.x8lo, PA * 128));
    TRY(wsalloc(e, &ws.x8hi, PA * 128));
    // InstanceNorm partial sums: one row per k_in_stats CTA, or per conv tile (+ its cluster padding) when the conv
    // epilogue produces them, plus the first-level fold results
    const int64_t stat_ctas = std::max<int64_t>((int64_t)e->num_sms * 2, (int64_t)cdiv(L, 16) * cdiv(rows2d, 8) + 4);
    const int64_t stat_grps = stat_ctas / 16 + 2;
    TRY(wsalloc(e, &ws.stat_part, (stat_ctas + stat_grps) * 256));
    TRY(wsalloc(e, &ws.norm_ss, 256));
    TRY(wsalloc(e, &ws.ticket, 64 + stat_grps));
    CUDA_TRY(e, cudaMemset(ws.ticket, 0, (64 + stat_grps) * sizeof(unsigned int)));
    TRY(wsalloc(e, &ws.sched, 2));
    CUDA_TRY(e, cudaMemset(ws.sched, 0, 2 * sizeof(unsigned long long)));
    e->conv_sched_base = 0;
    TRY(wsalloc(e, &ws.head, 2 * P));
    TRY(wsalloc(e, &ws.conf, L));
    TRY(wsalloc(e, &ws.mmat, P));
    TRY(wsalloc(e, &ws.eig_a, 2 * P));
    TRY(wsalloc(e, &ws.eig_w, 64 * (int64_t)L));
    TRY(wsalloc(e, &ws.eig_val, 8));
    TRY(wsalloc(e, &ws.mds, (int64_t)L * 8));
    TRY(wsalloc(e, &ws.ca, 3 * L));
    TRY(wsalloc(e, &ws.best_ca, 3 * L));
    TRY(wsalloc(e, &ws.best_conf, L));
    TRY(wsalloc(e, &ws.best_mean, 4));
    TRY(wsalloc(e, &ws.coords_out, 15 * L));
    TRY(wsalloc(e, &ws.conf_out, L));
    ws.L = L;
    ws.N = N;
    ws.rows2d = rows2d;
    ws.full_act = want_full;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------------
static int check_device(dmp2_engine* e) {
    CUDA_TRY(e, cudaSetDevice(e->device));
    return 0;
}

// stem update -> 16 blocks -> head -> eig/MDS -> coordinate GRU; results in ws.ca / ws.conf
static int one_pass(dmp2_engine* e, int L, cudaStream_t st) {
    Workspace& ws = e->ws;
    TRY(run_stem_update(e, ws.dmap, L, st));
    for (int k = 0; k < DMP2_NBLOCKS; k++) TRY(run_resblock(e, k, L, st));
    float* head = head_of(e);                  // halo-sharded: the window, where the peers' rows land too
    TRY(run_head(e, ws.x, L, head, st));
    TRY(run_head_post(e, head, L, ws.conf, ws.mmat, st));
    TRY(run_eig_top8(e, ws.mmat, L, ws.eig_val, ws.mds, nullptr, st));
    TRY(run_coord_head(e, ws.mat1d_t, ws.mds, L, ws.ca, st));
    return 0;
}

// one recycling iteration (network.py:265-306): distance map of the current coordinates -> pass -> keep the best
static int recycle_once(dmp2_engine* e, int L, cudaStream_t st) {
    Workspace& ws = e->ws;
    TRY(run_dmap(e, ws.ca, L, ws.dmap, true, st));
    TRY(one_pass(e, L, st));
    return run_select(e, ws.ca, ws.conf, L, 0, st);
}

// Largest L whose eigen step is a plain cluster launch; beyond it run_eig_top8 uses a cooperative whole-GPU launch,
// which is left out of graphs.
#define DMP2_GRAPH_MAX_L 600

// Capture one recycling iteration on the engine's capture stream and instantiate it.  Every launch of the iteration
// reads and writes workspace buffers only, and no host-side value changes between iterations, so the captured nodes
// are valid until the workspace or the kernel configuration changes (PassKey).  Returns 0 with e->pass_exec == nullptr
// when capture is not possible here (the caller then runs the iterations eagerly).
static int capture_pass(dmp2_engine* e, int L) {
    drop_pass_graph(e);
    if (!e->cap_stream) CUDA_TRY(e, cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    if (e->profile && e->gprof_ev.empty()) {
        e->gprof_ev.resize(2 * DMP2_NBLOCKS);
        for (auto& ev : e->gprof_ev) CUDA_TRY(e, cudaEventCreate(&ev));
    }
    const int64_t l0 = e->launches;
    if (cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        e->graph_on = false;
        return 0;
    }
    e->capturing = true;
    const int rc = recycle_once(e, L, e->cap_stream);
    e->capturing = false;
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
    const int64_t nodes = e->launches - l0;
    e->launches = l0;                                   // nothing has run yet
    cudaGraphExec_t exec = nullptr;
    if (rc == 0 && ce == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) {
        e->pass_exec = exec;
        e->pass_nodes = nodes;
    } else {                                            // not capturable on this driver / configuration: stay eager from now on
        cudaGetLastError();
        e->status = 0;
        e->err.clear();
        e->graph_on = false;
    }
    if (g) cudaGraphDestroy(g);
    return 0;
}

static int recycle(dmp2_engine* e, int L, int iterations, cudaStream_t st) {
    if (e->graph_on && iterations >= 2 && !e->strip_on && L <= DMP2_GRAPH_MAX_L) {
        dmp2_engine::PassKey k;
        k.L = L; k.conv_mode = e->conv_mode; k.conv_cluster = e->conv_cluster; k.conv_chunk_taps = e->conv_chunk_taps;
        k.conv_sms = e->conv_sms; k.conv_dynamic = e->conv_dynamic; k.gemm_tc = e->gemm_tc; k.fuse_stats = e->fuse_stats;
        k.eig_no_cl16 = e->eig_no_cl16; k.profile = e->profile; k.ws_gen = e->ws_gen;
        if (!e->pass_exec || !(e->pass_key == k)) {
            TRY(capture_pass(e, L));
            e->pass_key = k;
        }
        if (e->pass_exec) {
            for (int it = 0; it < iterations; it++) CUDA_TRY(e, cudaGraphLaunch(e->pass_exec, st));
            e->launches += e->pass_nodes * iterations;
            if (e->profile) e->gprof_valid = true;
            return 0;
        }
    }
    for (int it = 0; it < iterations; it++) TRY(recycle_once(e, L, st));
    return 0;
}

static int fold_impl(dmp2_engine* e, const uint8_t* msa, int N, int L, const float* tmpl, int iterations, int minsteps,
                     float* coords_out, float* conf_out, cudaStream_t st, bool timed) {
    if (!msa || !coords_out || !conf_out) return e->fail(DMP2_ERR_BAD_ARG, "fold: null pointer");
    if (L < 8) return e->fail(DMP2_ERR_BAD_ARG, "fold: L must be >= 8 (top-8 MDS embedding, network.py:250)");
    if (N < 1) return e->fail(DMP2_ERR_BAD_ARG, "fold: N must be >= 1");
    iterations = std::max(iterations, 0);       // predict.py:121-122
    minsteps = std::max(minsteps, 0);
    Workspace& ws = e->ws;
    int evi = 0;
    auto mark = [&]() { if (timed) cudaEventRecord(e->ev[evi++], st); };
    mark();
    // The MSA features (reweight + DCA) and the 1-D track (vgru + hgru) are independent until the stem: the
    // features run on the engine's side stream, forked from and joined back into the caller's stream.
    CUDA_TRY(e, cudaEventRecord(e->ev_fork, st));
    CUDA_TRY(e, cudaStreamWaitEvent(e->side, e->ev_fork, 0));
    TRY(run_reweight(e, msa, N, L, ws.seqw, e->side));
    TRY(run_dca(e, msa, N, L, ws.seqw, ws.feat, e->side));
    CUDA_TRY(e, cudaEventRecord(e->ev_join, e->side));
    float* v_last = ws.v_last;
    if (e->strip_on && e->sp.world > 1 && e->vgru_mode == 0) {
        // halo-sharded: the scan is independent per alignment column, so every rank takes a range of columns
        StripCtx& sp = e->sp;
        v_last = reinterpret_cast<float*>(sp.win + sp.off_vlast);
        const int cper = cdiv(L, sp.world), c0 = std::min(L, sp.rank * cper), c1 = std::min(L, c0 + cper);
        if (c1 > c0) TRY(run_vgru_tc(e, msa + c0, N, c1 - c0, v_last + (int64_t)c0 * 512, st, L));
        TRY(strip_vgru_gather(e, c0, c1, st));
    } else if (e->vgru_pre && !e->strip_on) {
        // the scan was done elsewhere (one call over the columns of several targets): take its result
        CUDA_TRY(e, cudaMemcpyAsync(ws.v_last, e->vgru_pre, (size_t)L * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        TRY(run_vgru(e, msa, N, L, ws.v_last, st));
    }
    e->vgru_pre = nullptr;
    mark();
    TRY(run_bigru(e, e->w.hgru, 2, v_last, L, ws.mat1d_t, st));
    mark();
    CUDA_TRY(e, cudaStreamWaitEvent(st, e->ev_join, 0));
    mark();
    TRY(run_stem_base(e, ws.mat1d_t, ws.feat, L, st));
    if (tmpl) TRY(run_dmap(e, tmpl, L, ws.dmap, false, st));          // predict.py:142-143
    else TRY(run_fill(e, ws.dmap, (int64_t)L * L, -1.0f, st));        // predict.py:145
    mark();
    TRY(one_pass(e, L, st));                                          // network.py:235-255
    if (minsteps > 0) TRY(run_refine(e, ws.ca, L, minsteps, st));     // network.py:257-258
    TRY(run_select(e, ws.ca, ws.conf, L, 1, st));
    TRY(recycle(e, L, iterations, st));                               // network.py:264-306
    mark();
    if (minsteps > 0) TRY(run_refine(e, ws.best_ca, L, minsteps, st)); // network.py:308-309
    TRY(run_backbone(e, ws.best_ca, ws.best_conf, L, coords_out, conf_out, st));
    mark();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int dmp2_create(dmp2_engine** out, int device, int n_tensors, const char* const* names, const float* const* host_ptrs,
                const int64_t* numels) {
    if (!out || !names || !host_ptrs || !numels || n_tensors <= 0) {
        g_create_error = "dmp2_create: bad arguments";
        return DMP2_ERR_BAD_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = "dmp2_create: no such CUDA device (this engine has no CPU fallback)";
        return DMP2_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        g_create_error = std::string("dmp2_create: device ") + prop.name + " is not sm_100 (Blackwell B200 required)";
        return DMP2_ERR_NO_DEVICE;
    }
    dmp2_engine* e = new dmp2_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    cudaSetDevice(device);
    HostSD sd;
    for (int i = 0; i < n_tensors; i++) sd.t[names[i]] = {host_ptrs[i], numels[i]};
    int s = load_weights(e, sd);
    if (s == DMP2_ERR_MISSING_WEIGHT) e->fail(s, "state_dict key missing or mis-sized: " + sd.missing);
    if (s != 0) {
        g_create_error = e->err;
        dmp2_destroy(e);
        return s;
    }
    for (int i = 0; i < 16; i++) cudaEventCreate(&e->ev[i]);
    {   // The MSA features run beside the vgru; the vgru is the critical path (1000 dependent steps that each want the
        // whole GPU), so the side stream gets the LOWEST priority: its kernels fill SMs only when no vgru CTA is pending.
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const char* sp = getenv("DMP2_SIDE_PRIORITY");
        if (sp && !strcmp(sp, "default")) cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking);
        else if (sp && !strcmp(sp, "high")) cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, hi);   // tuning knob: features first
        else cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, lo);
    }
    cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
    e->ev_ok = true;
    const char* mode = getenv("DMP2_CONV_MODE");
    if (mode) {
        if (!strcmp(mode, "f16x3")) e->conv_mode = DMP2_CONV_TC_F16X3;
        else if (!strcmp(mode, "f16")) e->conv_mode = DMP2_CONV_TC_F16;
        else if (!strcmp(mode, "ffma")) e->conv_mode = DMP2_CONV_FFMA;
        else if (!strcmp(mode, "f16f8")) e->conv_mode = DMP2_CONV_TC_F16F8;
    }
    const char* cc = getenv("DMP2_CONV_CLUSTER");
    if (cc && (!strcmp(cc, "1") || !strcmp(cc, "2"))) e->conv_cluster = atoi(cc);
    if (cc && !strcmp(cc, "pair")) e->conv_cluster = 0;
    const char* ck = getenv("DMP2_CONV_CHUNK");
    if (ck && (atoi(ck) == 1 || atoi(ck) == 5 || atoi(ck) == 25)) e->conv_chunk_taps = atoi(ck);
    const char* cs = getenv("DMP2_CONV_SMS");
    if (cs && atoi(cs) > 0) e->conv_sms = atoi(cs);
    const char* cd = getenv("DMP2_CONV_DYNAMIC");
    if (cd) e->conv_dynamic = strcmp(cd, "0") != 0;
    const char* vm = getenv("DMP2_VGRU");
    const char* gm = getenv("DMP2_GEMM");
    if (gm && !strcmp(gm, "ffma")) e->gemm_tc = false;
    const char* gr = getenv("DMP2_GRAPH");
    if (gr) e->graph_on = strcmp(gr, "0") != 0;
    const char* fs = getenv("DMP2_FUSE_STATS");
    if (fs) e->fuse_stats = strcmp(fs, "0") != 0;
    if (vm && !strcmp(vm, "ffma")) e->vgru_mode = 1;
    if (vm && !strcmp(vm, "steps")) e->vgru_mode = 0;
    *out = e;
    return 0;
}

void dmp2_destroy(dmp2_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    conv_tc_destroy(e);
    vgru_tc_destroy(e);
    strip_detach(e);
    free_workspace(e);
    for (void* p : e->weight_allocs) cudaFree(p);
    if (e->ev_ok) for (int i = 0; i < 16; i++) cudaEventDestroy(e->ev[i]);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (auto& ev : e->prof_ev) cudaEventDestroy(ev);
    for (auto& ev : e->gprof_ev) cudaEventDestroy(ev);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    delete e;
}

const char* dmp2_last_error(const dmp2_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int dmp2_set_conv_mode(dmp2_engine* e, int mode) {
    if (!e || mode < 0 || mode > 3) return DMP2_ERR_BAD_ARG;
    e->conv_mode = mode;
    return 0;
}
int64_t dmp2_launch_count(const dmp2_engine* e) { return e ? e->launches : 0; }

int dmp2_stage_times(const dmp2_engine* e, float* out_ms, int cap) {
    if (!e || !out_ms) return 0;
    int n = std::min(cap, 6);
    for (int i = 0; i < n; i++) out_ms[i] = e->stage_ms[i];
    return n;
}

int dmp2_debug_eig_phases(dmp2_engine* e, int L, double* out_us, int cap) {
    if (!e || !out_us || cap < 4 || !e->ws.eig_w) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    CUDA_TRY(e, cudaDeviceSynchronize());
    unsigned long long st[5];
    CUDA_TRY(e, cudaMemcpy(st, e->ws.eig_w + 35 * (int64_t)L, sizeof(st), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) out_us[i] = (double)(st[i + 1] - st[i]) * 1e-3;
    return 0;
}

int dmp2_set_profile(dmp2_engine* e, int on) {
    if (!e) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    if (on && e->prof_ev.empty()) {
        e->prof_ev.resize(2 * DMP2_NBLOCKS * 11 * 64);       // 64 folds of 10 recycles; beyond that launches go untimed
        for (auto& ev : e->prof_ev) CUDA_TRY(e, cudaEventCreate(&ev));
    }
    e->profile = on != 0;
    e->prof_used = 0;
    return 0;
}

int dmp2_conv_profile(dmp2_engine* e, int* n_launches, float* total_ms) {
    if (!e || !n_launches || !total_ms) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    CUDA_TRY(e, cudaDeviceSynchronize());
    float tot = 0.f;
    for (size_t i = 0; i + 1 < e->prof_used; i += 2) {
        float ms = 0.f;
        CUDA_TRY(e, cudaEventElapsedTime(&ms, e->prof_ev[i], e->prof_ev[i + 1]));
        tot += ms;
    }
    int n = (int)(e->prof_used / 2);
    if (e->gprof_valid) {                              // the conv nodes of the last graph replay (a sample: replays overwrite them)
        for (size_t i = 0; i + 1 < e->gprof_ev.size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e->gprof_ev[i], e->gprof_ev[i + 1]) == cudaSuccess) { tot += ms; n++; }
            else cudaGetLastError();
        }
        e->gprof_valid = false;
    }
    *n_launches = n;
    *total_ms = tot;
    e->prof_used = 0;
    return 0;
}

int dmp2_reserve(dmp2_engine* e, int L, int N) {
    if (!e) return DMP2_ERR_BAD_ARG;
    if (L < 8 || N < 1) return e->fail(DMP2_ERR_BAD_ARG, "reserve: need L >= 8 and N >= 1");
    TRY(check_device(e));
    return ensure_workspace(e, L, N);
}

int dmp2_fold(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, const float* tmpl_ca_dev, int iterations, int minsteps,
              float* coords_out_dev, float* conf_out_dev, void* stream) {
    if (!e) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    if (L >= 8 && N >= 1) TRY(ensure_workspace(e, L, N));
    return fold_impl(e, msa_dev, N, L, tmpl_ca_dev, iterations, minsteps, coords_out_dev, conf_out_dev, (cudaStream_t)stream, false);
}

// halo-sharded mode: same contract as dmp2_fold, but COLLECTIVE -- every rank of the strip group calls it with the
// same arguments; the 2-D track runs on this rank's rows, everything else is replicated, all ranks get the result.
static int strip_ready(dmp2_engine* e, int L) {
    if (!e->sp.attached || e->sp.L != L)
        return e->fail(DMP2_ERR_BAD_ARG, "fold_strip: call dmp2_strip_setup + dmp2_strip_attach for this L on every rank first");
    if (e->conv_mode == DMP2_CONV_FFMA) return e->fail(DMP2_ERR_UNSUPPORTED, "fold_strip: the CUDA-core validation conv has no halo-sharded form");
    return 0;
}

int dmp2_fold_strip(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, const float* tmpl_ca_dev, int iterations, int minsteps,
                    float* coords_out_dev, float* conf_out_dev, void* stream) {
    if (!e) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    if (L < 8 || N < 1) return e->fail(DMP2_ERR_BAD_ARG, "fold_strip: need L >= 8 and N >= 1");
    TRY(strip_ready(e, L));
    TRY(ensure_workspace(e, L, N, e->sp.r1 - e->sp.r0));
    e->strip_on = true;
    int rc = fold_impl(e, msa_dev, N, L, tmpl_ca_dev, iterations, minsteps, coords_out_dev, conf_out_dev, (cudaStream_t)stream, false);
    e->strip_on = false;
    return rc;
}

static int fold_host_impl(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                          int minsteps, float* coords_out_host, float* conf_out_host, bool strip);

int dmp2_fold_strip_host(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                         int minsteps, float* coords_out_host, float* conf_out_host) {
    return fold_host_impl(e, msa_host, N, L, tmpl_ca_host, iterations, minsteps, coords_out_host, conf_out_host, true);
}

int dmp2_strip_rows(int L, int world, int rank, int* r0, int* r1) { return strip_rows(L, world, rank, r0, r1, nullptr); }

int dmp2_strip_setup(dmp2_engine* e, int rank, int world, int L, int reserve_N, unsigned char* handle_out, void** window_out) {
    if (!e) return DMP2_ERR_BAD_ARG;
    TRY(strip_setup(e, rank, world, L, handle_out));
    if (reserve_N > 0) TRY(ensure_workspace(e, L, reserve_N, e->sp.r1 - e->sp.r0));   // no allocation (= device sync) inside the folds
    if (window_out) *window_out = e->sp.win;
    return 0;
}
int dmp2_strip_attach(dmp2_engine* e, const unsigned char* handles) {
    if (!e) return DMP2_ERR_BAD_ARG;
    return strip_attach(e, handles, nullptr);
}
int dmp2_strip_attach_local(dmp2_engine* e, void* const* windows) {
    if (!e) return DMP2_ERR_BAD_ARG;
    return strip_attach(e, nullptr, windows);
}
int dmp2_strip_detach(dmp2_engine* e) {
    if (!e) return DMP2_ERR_BAD_ARG;
    return strip_detach(e);
}

int dmp2_fold_host(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                   int minsteps, float* coords_out_host, float* conf_out_host) {
    return fold_host_impl(e, msa_host, N, L, tmpl_ca_host, iterations, minsteps, coords_out_host, conf_out_host, false);
}

static int fold_host_impl(dmp2_engine* e, const uint8_t* msa_host, int N, int L, const float* tmpl_ca_host, int iterations,
                          int minsteps, float* coords_out_host, float* conf_out_host, bool strip) {
    if (!e) return DMP2_ERR_BAD_ARG;
    if (!msa_host || !coords_out_host || !conf_out_host) return e->fail(DMP2_ERR_BAD_ARG, "fold_host: null pointer");
    if (L < 8 || N < 1) return e->fail(DMP2_ERR_BAD_ARG, "fold_host: need L >= 8 and N >= 1");
    TRY(check_device(e));
    if (strip) TRY(strip_ready(e, L));
    TRY(ensure_workspace(e, L, N, strip ? e->sp.r1 - e->sp.r0 : -1));
    Workspace& ws = e->ws;
    cudaStream_t st = 0;
    CUDA_TRY(e, cudaMemcpyAsync(ws.msa, msa_host, (size_t)N * L, cudaMemcpyHostToDevice, st));
    float* tmpl = nullptr;
    if (tmpl_ca_host) {
        tmpl = ws.best_ca;       // staged here; consumed by run_dmap before best_ca is first written
        CUDA_TRY(e, cudaMemcpyAsync(tmpl, tmpl_ca_host, (size_t)3 * L * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    e->strip_on = strip;
    int rc = fold_impl(e, ws.msa, N, L, tmpl, iterations, minsteps, ws.coords_out, ws.conf_out, st, true);
    e->strip_on = false;
    if (rc != 0) return rc;
    CUDA_TRY(e, cudaMemcpyAsync(coords_out_host, ws.coords_out, (size_t)15 * L * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaMemcpyAsync(conf_out_host, ws.conf_out, (size_t)L * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&e->stage_ms[i], e->ev[i], e->ev[i + 1]);
    return 0;
}

// ---- stage entry points ------------------------------------------------------------------------------
#define STAGE_PROLOGUE(Lv, Nv)                                                  \
    if (!e) return DMP2_ERR_BAD_ARG;                                            \
    TRY(check_device(e));                                                       \
    TRY(ensure_workspace(e, (Lv), (Nv)));                                       \
    cudaStream_t st = (cudaStream_t)stream;

int dmp2_reweight(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* w_out_dev, void* stream) {
    STAGE_PROLOGUE(L, N);
    return run_reweight(e, msa_dev, N, L, w_out_dev, st);
}

int dmp2_dca(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* feat_out_dev, void* stream) {
    STAGE_PROLOGUE(L, N);
    TRY(run_reweight(e, msa_dev, N, L, e->ws.seqw, st));
    TRY(run_dca(e, msa_dev, N, L, e->ws.seqw, e->ws.feat, st));
    return run_feat_export(e, e->ws.feat, L, feat_out_dev, st);
}

int dmp2_vgru(dmp2_engine* e, const uint8_t* msa_dev, int N, int L, float* out_dev, void* stream) {
    STAGE_PROLOGUE(L, N);
    return run_vgru(e, msa_dev, N, L, out_dev, st);
}

int dmp2_hgru(dmp2_engine* e, const float* in_dev, int L, float* out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    return run_bigru(e, e->w.hgru, 2, in_dev, L, out_dev, st);
}

int dmp2_conv5_maxout(dmp2_engine* e, int block, const float* x_nhwc_dev, int L, float* out_nhwc_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (block < 1 || block > DMP2_NBLOCKS) return e->fail(DMP2_ERR_BAD_ARG, "conv5_maxout: block must be 1..16");
    if (e->conv_mode == DMP2_CONV_FFMA) return run_conv_ffma(e, block - 1, x_nhwc_dev, L, out_nhwc_dev, st);
    TRY(run_split_half(e, x_nhwc_dev, (int64_t)L * L * 128, e->ws.xh, e->ws.xl, e->ws.x8lo, e->ws.x8hi, st));
    return run_conv_tc(e, block - 1, e->ws.xh, e->ws.xl, e->ws.x8lo, e->ws.x8hi, L, L, 0, L, out_nhwc_dev, e->conv_mode, st);
}

int dmp2_resblock(dmp2_engine* e, int block, const float* x_nhwc_dev, int L, float* out_nhwc_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (block < 1 || block > DMP2_NBLOCKS) return e->fail(DMP2_ERR_BAD_ARG, "resblock: block must be 1..16");
    const int64_t n = (int64_t)L * L * 128;
    CUDA_TRY(e, cudaMemcpyAsync(e->ws.x, x_nhwc_dev, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY(run_split_half(e, e->ws.x, n, e->ws.xh, e->ws.xl, e->ws.x8lo, e->ws.x8hi, st));
    TRY(run_resblock(e, block - 1, L, st));
    CUDA_TRY(e, cudaMemcpyAsync(out_nhwc_dev, e->ws.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dmp2_stem(dmp2_engine* e, const float* mat1d_t_dev, const float* feat_dev, const float* dmap_dev, int L, float* out_nhwc_dev,
              void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (!mat1d_t_dev || !feat_dev || !dmap_dev || !out_nhwc_dev) return e->fail(DMP2_ERR_BAD_ARG, "stem: null pointer");
    TRY(run_feat_import(e, feat_dev, L, e->ws.feat, st));
    TRY(run_stem_base(e, mat1d_t_dev, e->ws.feat, L, st));
    TRY(run_stem_update(e, dmap_dev, L, st));
    CUDA_TRY(e, cudaMemcpyAsync(out_nhwc_dev, e->ws.x, (size_t)L * L * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dmp2_head(dmp2_engine* e, const float* x_nhwc_dev, int L, float* head_out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (!x_nhwc_dev || !head_out_dev) return e->fail(DMP2_ERR_BAD_ARG, "head: null pointer");
    return run_head(e, x_nhwc_dev, L, head_out_dev, st);
}

int dmp2_resnet_pass(dmp2_engine* e, const float* mat1d_t_dev, const float* feat_dev, const float* dmap_dev, int L,
                     float* head_out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    TRY(run_feat_import(e, feat_dev, L, e->ws.feat, st));
    TRY(run_stem_base(e, mat1d_t_dev, e->ws.feat, L, st));
    TRY(run_stem_update(e, dmap_dev, L, st));
    for (int k = 0; k < DMP2_NBLOCKS; k++) TRY(run_resblock(e, k, L, st));
    return run_head(e, e->ws.x, L, head_out_dev, st);
}

int dmp2_head_mds(dmp2_engine* e, const float* head_dev, int L, float* conf_out_dev, float* m_out_dev, float* mds_out_dev,
                  void* stream) {
    STAGE_PROLOGUE(L, 1);
    TRY(run_head_post(e, head_dev, L, conf_out_dev, m_out_dev, st));
    return run_eig_top8(e, m_out_dev, L, e->ws.eig_val, mds_out_dev, nullptr, st);
}

int dmp2_eig_top8(dmp2_engine* e, const float* m_dev, int L, float* vals_out_dev, float* vecs_out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (L < 8) return e->fail(DMP2_ERR_BAD_ARG, "eig_top8: L must be >= 8");
    return run_eig_top8(e, m_dev, L, vals_out_dev, nullptr, vecs_out_dev, st);
}

int dmp2_coord_gru(dmp2_engine* e, const float* mat1d_t_dev, const float* mds_dev, int L, float* ca_out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    return run_coord_head(e, mat1d_t_dev, mds_dev, L, ca_out_dev, st);
}

int dmp2_refine(dmp2_engine* e, float* ca_dev, int L, int steps, void* stream) {
    STAGE_PROLOGUE(L, 1);
    return run_refine(e, ca_dev, L, steps, st);
}

int dmp2_backbone(dmp2_engine* e, const float* ca_dev, int L, float* out_dev, void* stream) {
    STAGE_PROLOGUE(L, 1);
    if (L < 3) return e->fail(DMP2_ERR_BAD_ARG, "backbone: L must be >= 3");
    return run_backbone(e, ca_dev, nullptr, L, out_dev, nullptr, st);
}

int dmp2_gemm_tn_test(dmp2_engine* e, const float* a_dev, const float* b_dev, int M, int N, int K, int mode, int chunk_k,
                      float* c_dev, void* stream) {
    if (!e) return DMP2_ERR_BAD_ARG;
    TRY(check_device(e));
    return run_gemm_tn_test(e, a_dev, b_dev, M, N, K, mode, chunk_k, c_dev, (cudaStream_t)stream);
}

int dmp2_set_vgru_input(dmp2_engine* e, const float* vgru_dev) {
    if (!e) return DMP2_ERR_BAD_ARG;
    e->vgru_pre = vgru_dev;
    return 0;
}

int dmp2_set_conv_dynamic(dmp2_engine* e, int on) {
    if (!e) return DMP2_ERR_BAD_ARG;
    e->conv_dynamic = on != 0;
    return 0;
}

int dmp2_set_graph(dmp2_engine* e, int on) {
    if (!e) return DMP2_ERR_BAD_ARG;
    e->graph_on = on != 0;
    return 0;
}

int dmp2_set_conv_sms(dmp2_engine* e, int sms) {
    if (!e || sms < 0) return DMP2_ERR_BAD_ARG;
    e->conv_sms = sms;
    return 0;
}

}  // extern "C"
