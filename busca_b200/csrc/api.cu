// C ABI of libbusca_b200.so (include/busca_b200.h): context, weights, patch bank, workspace and the orchestration of
// the per-frame hot path.  Kernels live in geometry.cu / crop.cu / reid.cu / conv_tc.cu / transformer.cu.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <map>
#include <string>
#include <vector>

#include "../../include/busca_b200.h"
#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_OK(expr)                                                                                           \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return set_err(BUSCA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostTensor {
    std::vector<float> f;
    std::vector<int64_t> shape;
};

struct TLayer {
    float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b;
    void *in_w16, *out_w16, *l1_w16, *l2_w16;      // bf16 copies for the tensor-core linears (bf16 mode)
};

struct ProfEntry {
    std::string name;
    cudaEvent_t a, b;
    double flops;                                 // ALGORITHMIC FLOPs of the launch: every convolution counted once (statistics-only passes: 0)
    double xflops;                                // FLOPs the launch executed (statistics-only passes included)
    std::string kernel;                           // the __global__ instantiation, as the ncu launch list names it
};

struct busca_ctx {
    busca_config cfg;
    cudaStream_t stream = nullptr;
    bool finalized = false;
    std::map<std::string, HostTensor> host;      // staged until finalize
    std::vector<void *> owned;                   // device allocations freed on destroy
    // ReID
    std::vector<ConvLayer> convs;
    double *stats_pool = nullptr;
    size_t stats_bytes = 0;
    float *red_w = nullptr, *red_b = nullptr, *lut = nullptr;
    // Transformer
    float *enc_w = nullptr, *enc_b = nullptr, *sep = nullptr, *non = nullptr, *bad = nullptr;
    void *enc_w16 = nullptr;
    void *red_w16 = nullptr;                      // bf16 copy of red.weight: the 2048->512 reduction on the tensor cores (bf16 mode)
    std::vector<TLayer> layers;
    float *dec_g = nullptr, *dec_b = nullptr, *dec_w = nullptr, *dec_bias = nullptr;
    __half *pe_xy = nullptr, *pe_size = nullptr, *pe_t = nullptr;
    // frame + bank
    DevBuf frame;
    uint8_t *mirror = nullptr;                    // page-locked host copy of the frame in HBM (busca_sync_frame)
    size_t mirror_cap = 0;
    bool mirror_valid = false;
    int fH = 0, fW = 0;
    int64_t fstride = 0;
    uint8_t *bank = nullptr;
    int64_t bank_slots = 0;
    // scratch
    DevBuf ws_reid, ws_tr, ws_io, ws_small, ws_gram;
    uint8_t *d2h_ring = nullptr;                  // page-locked staging ring of d2h_pageable
    cudaEvent_t d2h_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf ws_ecc;                                // camera-motion compensation: 5 fp32 planes + partial sums + state
    int ecc_H = 0, ecc_W = 0, ecc_cur = 0;        // size of the cached planes; which of the two smoothed planes holds the LAST current frame
    bool ecc_have_prev = false;
    bool gram = true;                             // Gram-matrix statistics for the 1x1 convolutions with Cin <= 256 (BUSCA_GRAM=0 / option "gram")
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    int64_t launches = 0;
    bool use_tc = false;                          // bf16 mode: tcgen05 convolutions (BUSCA_CONV=simt forces the SIMT bf16 path)
    // duplicate elimination inside a BatchNorm batch (tensor-core path; BUSCA_DEDUP=0 / busca_set_option("dedup", 0) disables)
    bool dedup = true;
    bool defer_crop_copies = false;               // option "defer_crop_copies": busca_crop returns before its device->host copy has landed
    bool tr_tc = true;                            // bf16 mode: Decision-Transformer GEMMs on the tensor cores (option "tr_tc" 0: the fp32 SIMT linears)
    int *dedup_table = nullptr;                   // [bank_slots + 1], all 0x7f7f7f7f between kernels
    DevBuf ws_dedup[2];                           // per planned batch: uniq | map | weight | count
    int *h_nuniq = nullptr;                       // pinned: distinct-image counts of the (up to two) planned batches
    cudaEvent_t ev_plan = nullptr;
    int64_t reid_images_run = 0, reid_images_total = 0;
    // profiling
    bool profiling = false;
    std::vector<ProfEntry> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::string prof_json;
    double next_flops = 0.0, next_xflops = 0.0;   // consumed by the next prof_begin
    std::string next_kernel;
};

static cudaEvent_t get_event(busca_ctx *c) {
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
static void prof_begin(busca_ctx *c, const char *name) {
    if (!c->profiling) return;
    ProfEntry pe{name, get_event(c), get_event(c), c->next_flops, c->next_xflops > 0.0 ? c->next_xflops : c->next_flops, c->next_kernel};
    c->next_flops = c->next_xflops = 0.0;
    c->next_kernel.clear();
    cudaEventRecord(pe.a, c->stream);
    c->prof.push_back(pe);
}
static void prof_end(busca_ctx *c) {
    if (!c->profiling) return;
    cudaEventRecord(c->prof.back().b, c->stream);
}
static void prof_reset(busca_ctx *c) {
    c->prof.clear();
    c->ev_used = 0;
}
static void prof_collect(busca_ctx *c) {
    if (!c->profiling) return;
    struct Acc { double ms = 0, flops = 0, xflops = 0; int n = 0; std::string kernel; };
    std::map<std::string, Acc> acc;
    std::vector<std::string> order;
    for (auto &pe : c->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pe.a, pe.b);
        if (!acc.count(pe.name)) order.push_back(pe.name);
        Acc &a = acc[pe.name];
        a.ms += ms; a.flops += pe.flops; a.xflops += pe.xflops; a.n += 1; a.kernel = pe.kernel;
    }
    std::string js = "{";
    for (size_t i = 0; i < order.size(); ++i) {
        char buf[512];
        const Acc &a = acc[order[i]];
        snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"launches\": %d, \"flops\": %.6e, \"xflops\": %.6e, \"kernel\": \"%s\"}", i ? ", " : "",
                 order[i].c_str(), a.ms, a.n, a.flops, a.xflops, a.kernel.c_str());
        js += buf;
    }
    js += "}";
    c->prof_json = js;
}

// Wait for the stream after a SHORT operation (a single-box crop, a T x D matrix, one association round): the adapters make hundreds
// of such calls per frame (byte_tracker.py:468-479) and a blocking cudaStreamSynchronize costs a futex sleep / wake-up (tens of
// microseconds) each time.  Poll the stream for up to ~300 us first (BUSCA_SPIN=0 disables), then fall back to the blocking wait.
static int g_spin = -1;
static cudaError_t stream_wait_short(busca_ctx *c) {
    if (g_spin < 0) {
        const char *e = getenv("BUSCA_SPIN");
        g_spin = !(e && e[0] == '0');
    }
    if (g_spin) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 0;; ++it) {
            cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) return cudaSuccess;
            if (q != cudaErrorNotReady) return q;
            if ((it & 15) == 15 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
        }
    }
    return cudaStreamSynchronize(c->stream);
}

// programmatic dependent launch of the ReID kernel chain (common.cuh); BUSCA_PDL=1 enables
static int g_pdl = -1;
int pdl_enabled() {
    if (g_pdl < 0) {
        const char *e = getenv("BUSCA_PDL");
        g_pdl = (e && e[0] == '1');          // off by default: measured 47.1 ms/frame without, 48.0 with (profiles/r02u)
    }
    return g_pdl;
}
void pdl_set(int on) { g_pdl = on ? 1 : 0; }

// BUSCA_TRACE=1: print every kernel name and synchronise after it (attributing a hang or a fault to one launch)
static const bool g_trace = getenv("BUSCA_TRACE") != nullptr;
#define LAUNCH(ctx, name, call)                                                                            \
    do {                                                                                                   \
        if (g_trace) { fprintf(stderr, "[busca] %s ...", name); fflush(stderr); }                          \
        prof_begin(ctx, name);                                                                             \
        cudaError_t e_ = (call);                                                                           \
        prof_end(ctx);                                                                                     \
        (ctx)->launches++;                                                                                 \
        if (g_trace && e_ == cudaSuccess) { e_ = cudaStreamSynchronize((ctx)->stream); fprintf(stderr, " %s\n", cudaGetErrorString(e_)); } \
        if (e_ != cudaSuccess) return set_err(BUSCA_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
struct ConvSpec {
    std::string conv, bn;
    int cin, cout, k, stride;
};
static std::vector<ConvSpec> reid_specs() {
    std::vector<ConvSpec> v;
    v.push_back({"conv1", "bn1", 3, 64, 7, 2});
    const int planes[4] = {64, 128, 256, 512}, blocks[4] = {3, 4, 6, 3}, strides[4] = {1, 2, 2, 2};
    int inpl = 64;
    for (int li = 0; li < 4; ++li)
        for (int b = 0; b < blocks[li]; ++b) {
            std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
            int s = b == 0 ? strides[li] : 1;
            v.push_back({p + ".conv1", p + ".bn1", inpl, planes[li], 1, 1});
            v.push_back({p + ".conv2", p + ".bn2", planes[li], planes[li], 3, s});
            v.push_back({p + ".conv3", p + ".bn3", planes[li], planes[li] * 4, 1, 1});
            if (b == 0) v.push_back({p + ".downsample.0", p + ".downsample.1", inpl, planes[li] * 4, 1, s});
            inpl = planes[li] * 4;
        }
    return v;
}

extern "C" const char *busca_version(void) { return "busca_b200 0.1 (sm_100a)"; }
extern "C" const char *busca_last_error(void) { return g_err; }

extern "C" int busca_create(const busca_config *cfg, busca_ctx **out) {
    if (!cfg || !out) return set_err(BUSCA_ERR_ARG, "null argument");
    if (cfg->d_model != EMB_DIM || cfg->d_model % cfg->nhead != 0 || cfg->d_model / cfg->nhead != 128)
        return set_err(BUSCA_ERR_ARG, "unsupported transformer shape d_model=%d nhead=%d (need 512 / head dim 128)", cfg->d_model, cfg->nhead);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(BUSCA_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    CUDA_OK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return set_err(BUSCA_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    busca_ctx *c = new busca_ctx();
    c->cfg = *cfg;
    const char *cm = getenv("BUSCA_CONV");
    c->use_tc = cfg->precision == BUSCA_PREC_BF16 && !(cm && strcmp(cm, "simt") == 0);
    const char *dd = getenv("BUSCA_DEDUP");
    c->dedup = !(dd && strcmp(dd, "0") == 0);
    const char *gg = getenv("BUSCA_GRAM");
    c->gram = !(gg && strcmp(gg, "0") == 0);
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaHostAlloc((void **)&c->h_nuniq, 4 * sizeof(int), cudaHostAllocPortable));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming));
    *out = c;
    int64_t slots = cfg->bank_slots > 0 ? cfg->bank_slots : 1024;
    return busca_bank_reserve(c, slots);
}

extern "C" void busca_destroy(busca_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    for (void *p : c->owned) cudaFree(p);
    if (c->bank) cudaFree(c->bank);
    if (c->dedup_table) cudaFree(c->dedup_table);
    if (c->h_nuniq) cudaFreeHost(c->h_nuniq);
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    c->ws_dedup[0].release();
    c->ws_dedup[1].release();
    c->frame.release();
    if (c->mirror) cudaFreeHost(c->mirror);
    c->ws_reid.release();
    c->ws_tr.release();
    c->ws_io.release();
    c->ws_small.release();
    c->ws_gram.release();
    c->ws_ecc.release();
    if (c->d2h_ring) cudaFreeHost(c->d2h_ring);
    for (auto e : c->d2h_ev) if (e) cudaEventDestroy(e);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int busca_load_tensor(busca_ctx *c, const char *name, const void *data, int32_t dtype, int32_t ndim, const int64_t *shape) {
    if (!c || !name || (!data && ndim > 0)) return set_err(BUSCA_ERR_ARG, "null argument");
    int64_t n = 1;
    HostTensor t;
    for (int i = 0; i < ndim; ++i) { n *= shape[i]; t.shape.push_back(shape[i]); }
    t.f.resize((size_t)n);
    if (dtype == BUSCA_F32) memcpy(t.f.data(), data, (size_t)n * 4);
    else if (dtype == BUSCA_F16) { const __half *h = (const __half *)data; for (int64_t i = 0; i < n; ++i) t.f[i] = __half2float(h[i]); }
    else if (dtype == BUSCA_I64) { const int64_t *h = (const int64_t *)data; for (int64_t i = 0; i < n; ++i) t.f[i] = (float)h[i]; }
    else return set_err(BUSCA_ERR_ARG, "bad dtype %d for %s", dtype, name);
    c->host[name] = std::move(t);
    c->finalized = false;
    return BUSCA_OK;
}

static const HostTensor *find(busca_ctx *c, const std::string &name, std::initializer_list<int64_t> shape) {
    auto it = c->host.find(name);
    if (it == c->host.end()) { set_err(BUSCA_ERR_STATE, "missing tensor '%s'", name.c_str()); return nullptr; }
    std::vector<int64_t> want(shape);
    if (it->second.shape != want) { set_err(BUSCA_ERR_STATE, "tensor '%s' has the wrong shape", name.c_str()); return nullptr; }
    return &it->second;
}
template <typename T>
static T *upload(busca_ctx *c, const T *src, size_t n) {
    void *p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T) + 16) != cudaSuccess) return nullptr;
    cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice);
    c->owned.push_back(p);
    return (T *)p;
}
static float *upload_named(busca_ctx *c, const std::string &name, std::initializer_list<int64_t> shape) {
    const HostTensor *t = find(c, name, shape);
    return t ? upload(c, t->f.data(), t->f.size()) : nullptr;
}
static void *upload_named_bf16(busca_ctx *c, const std::string &name, std::initializer_list<int64_t> shape) {
    const HostTensor *t = find(c, name, shape);
    if (!t) return nullptr;
    std::vector<__nv_bfloat16> h(t->f.size());
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16(t->f[i]);
    return upload(c, h.data(), h.size());
}
#define NEED(ptr) do { if (!(ptr)) return g_err[0] ? BUSCA_ERR_STATE : set_err(BUSCA_ERR_NOMEM, "upload failed"); } while (0)

extern "C" int busca_finalize(busca_ctx *c) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    g_err[0] = 0;
    for (void *p : c->owned) cudaFree(p);
    c->owned.clear();
    c->convs.clear();
    c->layers.clear();
    const std::string r = "reid_encoder.model.";
    auto specs = reid_specs();
    size_t stats_doubles = 0;
    for (auto &sp : specs) stats_doubles += 2 * (size_t)sp.cout;
    CUDA_OK(cudaMalloc((void **)&c->stats_pool, stats_doubles * sizeof(double)));
    c->owned.push_back(c->stats_pool);
    c->stats_bytes = stats_doubles * sizeof(double);
    size_t soff = 0;
    for (auto &sp : specs) {
        ConvLayer L{};
        L.cin = sp.cin; L.cout = sp.cout; L.k = sp.k; L.stride = sp.stride;
        const HostTensor *w = find(c, r + sp.conv + ".weight", {sp.cout, sp.cin, sp.k, sp.k});
        NEED(w);
        std::vector<float> re((size_t)sp.cout * sp.k * sp.k * sp.cin);
        if (sp.cin == 3) {
            // stem: [tap = (ky*7+kx)*3 + c_bgr][cout]; the reference feeds RGB (network.py:397), the bank holds BGR
            for (int o = 0; o < sp.cout; ++o)
                for (int ci = 0; ci < 3; ++ci)
                    for (int ky = 0; ky < 7; ++ky)
                        for (int kx = 0; kx < 7; ++kx)
                            re[(size_t)((ky * 7 + kx) * 3 + (2 - ci)) * 64 + o] = w->f[(((size_t)o * 3 + ci) * 7 + ky) * 7 + kx];
        } else {
            // OIHW -> O,kh,kw,I  (K-major rows for the implicit GEMM)
            for (int o = 0; o < sp.cout; ++o)
                for (int ci = 0; ci < sp.cin; ++ci)
                    for (int ky = 0; ky < sp.k; ++ky)
                        for (int kx = 0; kx < sp.k; ++kx)
                            re[(((size_t)o * sp.k + ky) * sp.k + kx) * sp.cin + ci] = w->f[(((size_t)o * sp.cin + ci) * sp.k + ky) * sp.k + kx];
        }
        if (sp.cin == 3) {
            // tensor-core stem weights: bf16 [cout][row pair p][kx*8 + r*4 + c_bgr] = W[o][c][2p+r][kx], zeros for kx = 7, ky = 7, c = 3
            std::vector<__nv_bfloat16> h((size_t)64 * 4 * 64, __float2bfloat16(0.f));
            for (int o = 0; o < 64; ++o)
                for (int ci = 0; ci < 3; ++ci)
                    for (int ky = 0; ky < 7; ++ky)
                        for (int kx = 0; kx < 7; ++kx)
                            h[((size_t)o * 4 + ky / 2) * 64 + kx * 8 + (ky & 1) * 4 + (2 - ci)] = __float2bfloat16(w->f[(((size_t)o * 3 + ci) * 7 + ky) * 7 + kx]);
            L.w16 = upload(c, h.data(), h.size());
            NEED(L.w16);
        }
        if (sp.cin != 3) {
            std::vector<__nv_bfloat16> h(re.size());
            for (size_t i = 0; i < re.size(); ++i) h[i] = __float2bfloat16(re[i]);
            L.w16 = upload(c, h.data(), h.size());
            NEED(L.w16);
            if (c->cfg.precision == BUSCA_PREC_BF16 && sp.cin <= 512) {
                // tensor-core path: unrounded master + scratch for the per-call folded weights and transform parameters
                L.w32m = upload(c, re.data(), re.size());
                NEED(L.w32m);
                L.w16s = upload(c, h.data(), h.size());
                NEED(L.w16s);
                std::vector<uint16_t> z(2 * (size_t)sp.cin, 0);
                L.xf = upload(c, z.data(), z.size());
                NEED(L.xf);
            }
            if (c->cfg.precision == BUSCA_PREC_BF16 && sp.k == 1 && sp.cout >= 256 && sp.cout > sp.cin) {
                // last conv of a bottleneck / downsample conv: weights of the FINAL pass (times the scale of the conv's own BatchNorm)
                if (!L.w32m) { L.w32m = upload(c, re.data(), re.size()); NEED(L.w32m); }
                L.w16f = upload(c, h.data(), h.size());
                NEED(L.w16f);
            }
            // bf16 mode: the SIMT kernel multiplies by the same bf16-rounded weights as the tensor-core kernel
            if (c->cfg.precision == BUSCA_PREC_BF16)
                for (size_t i = 0; i < re.size(); ++i) re[i] = __bfloat162float(h[i]);
        }
        L.w32 = upload(c, re.data(), re.size());
        NEED(L.w32);
        L.gamma = upload_named(c, r + sp.bn + ".weight", {sp.cout});
        NEED(L.gamma);
        L.beta = upload_named(c, r + sp.bn + ".bias", {sp.cout});
        NEED(L.beta);
        L.stats = c->stats_pool + soff;
        soff += 2 * (size_t)sp.cout;
        std::vector<float> z(2 * (size_t)sp.cout, 0.f);
        L.scale = upload(c, z.data(), z.size());
        NEED(L.scale);
        L.shift = L.scale + sp.cout;
        c->convs.push_back(L);
    }
    NEED(c->red_w = upload_named(c, r + "red.weight", {512, 2048}));
    NEED(c->red_b = upload_named(c, r + "red.bias", {512}));
    NEED(c->red_w16 = upload_named_bf16(c, r + "red.weight", {512, 2048}));
    NEED(c->lut = upload_named(c, "norm.lut", {256, 3}));
    const int d = c->cfg.d_model, ff = c->cfg.ff_size;
    NEED(c->enc_w = upload_named(c, "encoder.weight", {d, d}));
    NEED(c->enc_b = upload_named(c, "encoder.bias", {d}));
    NEED(c->enc_w16 = upload_named_bf16(c, "encoder.weight", {d, d}));
    NEED(c->sep = upload_named(c, "sep_token", {d}));
    NEED(c->non = upload_named(c, "non_token", {d}));
    NEED(c->bad = upload_named(c, "bad_token", {d}));
    for (int l = 0; l < c->cfg.num_layers; ++l) {
        std::string p = "transformer_encoder.layers." + std::to_string(l) + ".";
        TLayer t{};
        NEED(t.in_w = upload_named(c, p + "self_attn.in_proj_weight", {3 * d, d}));
        NEED(t.in_b = upload_named(c, p + "self_attn.in_proj_bias", {3 * d}));
        NEED(t.out_w = upload_named(c, p + "self_attn.out_proj.weight", {d, d}));
        NEED(t.out_b = upload_named(c, p + "self_attn.out_proj.bias", {d}));
        NEED(t.l1_w = upload_named(c, p + "linear1.weight", {ff, d}));
        NEED(t.l1_b = upload_named(c, p + "linear1.bias", {ff}));
        NEED(t.l2_w = upload_named(c, p + "linear2.weight", {d, ff}));
        NEED(t.l2_b = upload_named(c, p + "linear2.bias", {d}));
        NEED(t.n1_g = upload_named(c, p + "norm1.weight", {d}));
        NEED(t.n1_b = upload_named(c, p + "norm1.bias", {d}));
        NEED(t.n2_g = upload_named(c, p + "norm2.weight", {d}));
        NEED(t.n2_b = upload_named(c, p + "norm2.bias", {d}));
        NEED(t.in_w16 = upload_named_bf16(c, p + "self_attn.in_proj_weight", {3 * d, d}));
        NEED(t.out_w16 = upload_named_bf16(c, p + "self_attn.out_proj.weight", {d, d}));
        NEED(t.l1_w16 = upload_named_bf16(c, p + "linear1.weight", {ff, d}));
        NEED(t.l2_w16 = upload_named_bf16(c, p + "linear2.weight", {d, ff}));
        c->layers.push_back(t);
    }
    NEED(c->dec_g = upload_named(c, "decoder.0.weight", {d}));
    NEED(c->dec_b = upload_named(c, "decoder.0.bias", {d}));
    NEED(c->dec_w = upload_named(c, "decoder.1.weight", {1, d}));
    NEED(c->dec_bias = upload_named(c, "decoder.1.bias", {1}));
    auto upload_half = [&](const char *name, int64_t rows, int64_t cols) -> __half * {
        const HostTensor *t = find(c, name, {rows, cols});
        if (!t) return nullptr;
        std::vector<__half> h(t->f.size());
        for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half(t->f[i]);   // exact: values were fp16
        return upload(c, h.data(), h.size());
    };
    NEED(c->pe_xy = upload_half("pe.tab_xy", 2 * PE_MAX_XY + 1, PE_CH));
    NEED(c->pe_size = upload_half("pe.tab_size", 2 * PE_MAX_SIZE + 1, PE_CH));
    NEED(c->pe_t = upload_half("pe.tab_t", 2 * PE_MAX_T + 1, PE_CH_T));
    c->host.clear();
    c->finalized = true;
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// frame + bank
// ------------------------------------------------------------------------------------------------
extern "C" int busca_upload_frame(busca_ctx *c, const uint8_t *bgr, int32_t H, int32_t W, int64_t row_stride) {
    if (!c || !bgr || H <= 0 || W <= 0 || row_stride < (int64_t)W * 3) return set_err(BUSCA_ERR_ARG, "bad frame");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(c->frame.ensure((size_t)H * W * 3 + 64));
    CUDA_OK(cudaMemcpy2DAsync(c->frame.p, (size_t)W * 3, bgr, (size_t)row_stride, (size_t)W * 3, H, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->fH = H; c->fW = W; c->fstride = (int64_t)W * 3;
    c->mirror_valid = false;
    return BUSCA_OK;
}

// Make the frame in HBM show, inside the given boxes (or everywhere), exactly the pixels of `bgr` - uploading only when they differ
// from the page-locked host mirror of what is already there.  The adapters call get_image_crops 3 + T times per frame with the same
// image (byte_tracker.py:278-282, 468-479); this keeps those calls at a memcmp of the rows they read.
extern "C" int busca_sync_frame(busca_ctx *c, const uint8_t *bgr, int32_t H, int32_t W, int64_t row_stride, const double *boxes, int32_t n_boxes,
                                int32_t *uploaded) {
    if (!c || !bgr || H <= 0 || W <= 0 || row_stride < (int64_t)W * 3) return set_err(BUSCA_ERR_ARG, "bad frame");
    if (uploaded) *uploaded = 0;
    const size_t rowb = (size_t)W * 3;
    if (c->mirror_valid && c->fH == H && c->fW == W) {
        int y0 = 0, y1 = H, x0 = 0, x1 = W;
        if (boxes && n_boxes > 0 && n_boxes <= 8) {
            double bx0 = 1e300, by0 = 1e300, bx1 = -1e300, by1 = -1e300;
            for (int i = 0; i < n_boxes; ++i) {
                bx0 = fmin(bx0, boxes[4 * i]); by0 = fmin(by0, boxes[4 * i + 1]);
                bx1 = fmax(bx1, boxes[4 * i + 2]); by1 = fmax(by1, boxes[4 * i + 3]);
            }
            if (bx0 == bx0 && bx1 == bx1 && by0 == by0 && by1 == by1) {       // no NaN
                x0 = (int)fmax(0.0, fmin((double)W, floor(bx0) - 1)); x1 = (int)fmax(0.0, fmin((double)W, ceil(bx1) + 1));
                y0 = (int)fmax(0.0, fmin((double)H, floor(by0) - 1)); y1 = (int)fmax(0.0, fmin((double)H, ceil(by1) + 1));
            }
        }
        bool same = true;
        if (x1 > x0)
            for (int y = y0; y < y1 && same; ++y)
                same = memcmp(bgr + (size_t)y * row_stride + (size_t)x0 * 3, c->mirror + (size_t)y * rowb + (size_t)x0 * 3, (size_t)(x1 - x0) * 3) == 0;
        if (same) return BUSCA_OK;
    }
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t bytes = (size_t)H * rowb;
    if (bytes > c->mirror_cap) {
        if (c->mirror) cudaFreeHost(c->mirror);
        c->mirror = nullptr; c->mirror_cap = 0;
        CUDA_OK(cudaHostAlloc((void **)&c->mirror, bytes, cudaHostAllocPortable));
        c->mirror_cap = bytes;
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));                 // an earlier upload may still be reading the mirror
    for (int y = 0; y < H; ++y) memcpy(c->mirror + (size_t)y * rowb, bgr + (size_t)y * row_stride, rowb);
    CUDA_OK(c->frame.ensure(bytes + 64));
    CUDA_OK(cudaMemcpyAsync(c->frame.p, c->mirror, bytes, cudaMemcpyHostToDevice, c->stream));     // page-locked source: a plain DMA, stream ordered
    c->fH = H; c->fW = W; c->fstride = (int64_t)rowb;
    c->mirror_valid = true;
    if (uploaded) *uploaded = 1;
    return BUSCA_OK;
}

// Frame ingest on the device (mot_evaluator.py:198-204).  The result becomes the current frame of the context (what
// busca_crop reads) AND the page-locked mirror, so that a later busca_sync_frame with the returned host copy finds it in place.
extern "C" int busca_ingest_frame(busca_ctx *c, const float *chw, int32_t chw_on_device, int32_t H, int32_t W, const float *rgb_mean,
                                  const float *rgb_std, uint8_t *frame_out) {
    if (!c || !chw || H <= 0 || W <= 0 || !rgb_mean || !rgb_std) return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t npix = (size_t)H * W, bytes = npix * 3;
    const float *src = chw;
    if (!chw_on_device) {
        CUDA_OK(c->ws_io.ensure(npix * 12 + 64));
        CUDA_OK(cudaMemcpyAsync(c->ws_io.p, chw, npix * 12, cudaMemcpyHostToDevice, c->stream));
        src = (const float *)c->ws_io.p;
    }
    CUDA_OK(c->frame.ensure(bytes + 64));
    prof_reset(c);
    LAUNCH(c, "frame_ingest", launch_frame_ingest(src, H, W, rgb_mean, rgb_std, (uint8_t *)c->frame.p, c->stream));
    c->fH = H; c->fW = W; c->fstride = (int64_t)W * 3;
    c->mirror_valid = false;
    if (frame_out) {
        if (bytes > c->mirror_cap) {
            CUDA_OK(cudaStreamSynchronize(c->stream));
            if (c->mirror) cudaFreeHost(c->mirror);
            c->mirror = nullptr; c->mirror_cap = 0;
            CUDA_OK(cudaHostAlloc((void **)&c->mirror, bytes, cudaHostAllocPortable));
            c->mirror_cap = bytes;
        }
        CUDA_OK(cudaMemcpyAsync(c->mirror, c->frame.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        memcpy(frame_out, c->mirror, bytes);
        c->mirror_valid = true;
    } else {
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    prof_collect(c);
    return BUSCA_OK;
}

// Camera-motion compensation (byte_tracker.py:626-657): ecc.cu
extern "C" int busca_camera_motion(busca_ctx *c, const uint8_t *prev_bgr, const uint8_t *cur_bgr, int32_t H, int32_t W, int64_t row_stride,
                                   int32_t iterations, double eps, float *warp_out, double *rho_out, int32_t *iterations_out) {
    if (!c || H < 5 || W < 5 || !warp_out) return set_err(BUSCA_ERR_ARG, "bad argument");
    if ((prev_bgr || cur_bgr) && row_stride < (int64_t)W * 3) return set_err(BUSCA_ERR_ARG, "bad row stride");
    if (!prev_bgr && !(c->ecc_have_prev && c->ecc_H == H && c->ecc_W == W))
        return set_err(BUSCA_ERR_STATE, "camera motion: no previous frame of this size is cached (pass prev_bgr)");
    if (!cur_bgr && !(c->frame.p && c->fH == H && c->fW == W)) return set_err(BUSCA_ERR_STATE, "camera motion: no current frame of this size in HBM");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t plane = ((size_t)H * W * 4 + 255) / 256 * 256;
    int bx, by, rpb;
    ecc_grid(H, W, &bx, &by, &rpb);
    const size_t part = ((size_t)bx * by * 21 * 8 + 255) / 256 * 256;
    if (c->ecc_H != H || c->ecc_W != W) { c->ecc_have_prev = false; c->ecc_H = H; c->ecc_W = W; c->ecc_cur = 0; }
    const size_t need = 5 * plane + part + 512;
    if (need > c->ws_ecc.cap) {
        if (!prev_bgr) return set_err(BUSCA_ERR_STATE, "camera motion: cache lost");     // cannot happen: the size check above
        CUDA_OK(c->ws_ecc.ensure(need));
    }
    char *base = (char *)c->ws_ecc.p;
    float *smooth[2] = {(float *)base, (float *)(base + plane)};
    float *gx = (float *)(base + 2 * plane), *gy = (float *)(base + 3 * plane), *rows = (float *)(base + 4 * plane);
    double *partials = (double *)(base + 5 * plane);
    EccState *dst = (EccState *)(base + 5 * plane + part);
    unsigned int *ticket = (unsigned int *)(base + 5 * plane + part + 256);
    prof_reset(c);
    const size_t fbytes = (size_t)H * row_stride;
    // the previous call's current frame sits in smooth[ecc_cur]; it becomes the template, the new frame goes into the other plane
    int t_idx = c->ecc_cur, i_idx = c->ecc_cur ^ 1;
    if (prev_bgr) {
        CUDA_OK(c->ws_io.ensure(fbytes + 64));
        CUDA_OK(cudaMemcpyAsync(c->ws_io.p, prev_bgr, fbytes, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, "ecc_prepare", launch_ecc_prepare((const uint8_t *)c->ws_io.p, row_stride, H, W, rows, smooth[t_idx], nullptr, nullptr, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));           // ws_io is reused for the current frame below
    }
    const uint8_t *cur_dev = (const uint8_t *)c->frame.p;
    long long cur_stride = c->fstride;
    if (cur_bgr) {
        CUDA_OK(c->ws_io.ensure(fbytes + 64));
        CUDA_OK(cudaMemcpyAsync(c->ws_io.p, cur_bgr, fbytes, cudaMemcpyHostToDevice, c->stream));
        cur_dev = (const uint8_t *)c->ws_io.p;
        cur_stride = row_stride;
    }
    LAUNCH(c, "ecc_prepare", launch_ecc_prepare(cur_dev, cur_stride, H, W, rows, smooth[i_idx], gx, gy, c->stream));
    EccState st = {};
    st.map[0] = st.map[4] = 1.0f;
    st.rho = -1.0; st.last_rho = -eps;
    st.max_iterations = iterations;
    st.done = iterations <= 0 ? 1 : 0;
    CUDA_OK(cudaMemcpyAsync(dst, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(ticket, 0, 4, c->stream));
    while (!st.done) {
        for (int k = 0; k < 4; ++k)
            LAUNCH(c, "ecc_iteration", launch_ecc_iteration(smooth[t_idx], smooth[i_idx], gx, gy, H, W, dst, partials, ticket, eps, c->stream));
        CUDA_OK(cudaMemcpyAsync(&st, dst, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    prof_collect(c);
    c->ecc_cur = i_idx;
    c->ecc_have_prev = true;
    for (int k = 0; k < 6; ++k) warp_out[k] = st.map[k];
    if (rho_out) *rho_out = st.rho;
    if (iterations_out) *iterations_out = st.iterations;
    if (st.status == 1) return set_err(BUSCA_ERR_STATE, "ECC: the correlation is going to be minimized - images may be uncorrelated or non-overlapped (cv2 raises StsNoConv)");
    if (st.status == 2) return set_err(BUSCA_ERR_STATE, "ECC: NaN encountered (cv2 raises StsNoConv)");
    return BUSCA_OK;
}

extern "C" int busca_bank_reserve(busca_ctx *c, int64_t n_slots) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    if (n_slots <= c->bank_slots) return BUSCA_OK;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    uint8_t *nb = nullptr;
    cudaError_t e = cudaMalloc((void **)&nb, (size_t)n_slots * PATCH_BYTES);
    if (e != cudaSuccess) return set_err(BUSCA_ERR_NOMEM, "patch bank of %lld slots: %s", (long long)n_slots, cudaGetErrorString(e));
    if (c->bank) {
        CUDA_OK(cudaStreamSynchronize(c->stream));
        CUDA_OK(cudaMemcpy(nb, c->bank, (size_t)c->bank_slots * PATCH_BYTES, cudaMemcpyDeviceToDevice));
        cudaFree(c->bank);
    }
    c->bank = nb;
    c->bank_slots = n_slots;
    if (c->dedup_table) cudaFree(c->dedup_table);
    c->dedup_table = nullptr;
    CUDA_OK(cudaMalloc((void **)&c->dedup_table, ((size_t)n_slots + 1) * sizeof(int)));
    CUDA_OK(cudaMemsetAsync(c->dedup_table, 0x7f, ((size_t)n_slots + 1) * sizeof(int), c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}
extern "C" int64_t busca_bank_capacity(busca_ctx *c) { return c ? c->bank_slots : 0; }

static int check_slots(busca_ctx *c, const int32_t *slots, int n, bool allow_neg) {
    for (int i = 0; i < n; ++i)
        if (slots[i] >= c->bank_slots || (slots[i] < 0 && !allow_neg)) return set_err(BUSCA_ERR_ARG, "slot %d out of range (bank has %lld)", slots[i], (long long)c->bank_slots);
    return BUSCA_OK;
}

// Device -> PAGEABLE host for the big crop arrays the adapters keep (byte_tracker.py:278-282 returns ~45 MB per call at MOT20 scale and
// every track keeps a row of it alive, so those arrays cannot come from the page-locked pool).  A plain cudaMemcpyAsync into pageable
// memory is staged by the driver on one thread (~6.5 GB/s measured, page faults of the fresh array included); here the DMA lands in a
// ring of page-locked chunks at PCIe speed and worker threads copy finished chunks into the destination (first-touch faults in
// parallel).  The stream is idle when this returns.
constexpr size_t D2H_CHUNK = 2u << 20;
constexpr int D2H_RING = 8, D2H_WORKERS_DEFAULT = 6;       // BUSCA_D2H_WORKERS overrides (1..16)
static int d2h_pageable(busca_ctx *c, uint8_t *dst, const uint8_t *src, size_t bytes) {
    if (!c->d2h_ring) {
        CUDA_OK(cudaHostAlloc((void **)&c->d2h_ring, D2H_CHUNK * D2H_RING, cudaHostAllocPortable));
        for (int i = 0; i < D2H_RING; ++i) CUDA_OK(cudaEventCreateWithFlags(&c->d2h_ev[i], cudaEventDisableTiming));
    }
    const int K = (int)((bytes + D2H_CHUNK - 1) / D2H_CHUNK);
    std::vector<std::atomic<int>> state(K);                 // 0 = not enqueued, 1 = DMA enqueued (event recorded), 2 = copied out
    for (auto &x : state) x.store(0);
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    auto worker = [&]() {
        cudaSetDevice(c->cfg.device);
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= K) return;
            while (state[k].load(std::memory_order_acquire) == 0 && !failed.load()) std::this_thread::yield();
            if (failed.load()) return;
            if (cudaEventSynchronize(c->d2h_ev[k % D2H_RING]) != cudaSuccess) { failed.store(1); return; }
            const size_t off = (size_t)k * D2H_CHUNK, nb = std::min(D2H_CHUNK, bytes - off);
            memcpy(dst + off, c->d2h_ring + (size_t)(k % D2H_RING) * D2H_CHUNK, nb);
            state[k].store(2, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    static int n_workers = 0;
    if (!n_workers) {
        const char *e = getenv("BUSCA_D2H_WORKERS");
        n_workers = e ? std::max(1, std::min(16, atoi(e))) : D2H_WORKERS_DEFAULT;
    }
    for (int w = 0; w < std::min(n_workers, K); ++w) pool.emplace_back(worker);
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < K && err == cudaSuccess && !failed.load(); ++k) {
        if (k >= D2H_RING)
            while (state[k - D2H_RING].load(std::memory_order_acquire) != 2 && !failed.load()) std::this_thread::yield();   // ring slot free again
        const size_t off = (size_t)k * D2H_CHUNK, nb = std::min(D2H_CHUNK, bytes - off);
        err = cudaMemcpyAsync(c->d2h_ring + (size_t)(k % D2H_RING) * D2H_CHUNK, src + off, nb, cudaMemcpyDeviceToHost, c->stream);
        if (err == cudaSuccess) err = cudaEventRecord(c->d2h_ev[k % D2H_RING], c->stream);
        if (err != cudaSuccess) { failed.store(1); break; }
        state[k].store(1, std::memory_order_release);
    }
    for (auto &t : pool) t.join();
    if (err != cudaSuccess) return set_err(BUSCA_ERR_CUDA, "device->host copy: %s", cudaGetErrorString(err));
    if (failed.load()) return set_err(BUSCA_ERR_CUDA, "device->host copy failed");
    return BUSCA_OK;
}
static bool host_is_pageable(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// bank <-> host, one copy per run of consecutive slots (alloc_slots hands out ascending runs, so a frame's crops are
// usually a single DMA; with page-locked host memory it runs at PCIe speed)
static int bank_copy_runs(busca_ctx *c, const int32_t *slots, int n, uint8_t *host, bool to_bank) {
    static int pipelined = -1;
    if (pipelined < 0) { const char *e = getenv("BUSCA_D2H_PIPELINE"); pipelined = !(e && e[0] == '0'); }
    const bool big_pageable = !to_bank && pipelined && (size_t)n * PATCH_BYTES >= 2 * D2H_CHUNK && host_is_pageable(host);
    int i = 0;
    while (i < n) {
        int j = i + 1;
        while (j < n && slots[j] == slots[j - 1] + 1) ++j;
        uint8_t *dev = c->bank + (size_t)slots[i] * PATCH_BYTES, *h = host + (size_t)i * PATCH_BYTES;
        const size_t bytes = (size_t)(j - i) * PATCH_BYTES;
        if (to_bank) CUDA_OK(cudaMemcpyAsync(dev, h, bytes, cudaMemcpyHostToDevice, c->stream));
        else if (big_pageable && bytes >= D2H_CHUNK) { int rc = d2h_pageable(c, h, dev, bytes); if (rc) return rc; }
        else CUDA_OK(cudaMemcpyAsync(h, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        i = j;
    }
    return BUSCA_OK;
}

extern "C" int busca_crop(busca_ctx *c, const double *boxes, int32_t n, const int32_t *slots, uint8_t *host_out) {
    if (!c || n < 0 || (n > 0 && (!boxes || !slots))) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (n == 0) return BUSCA_OK;
    if (!c->frame.p || c->fH == 0) return set_err(BUSCA_ERR_STATE, "busca_crop before busca_upload_frame");
    int rc = check_slots(c, slots, n, false);
    if (rc) return rc;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    prof_reset(c);
    if (n <= 4) {
        CropSmall sm{};
        sm.n = n;
        memcpy(sm.boxes, boxes, (size_t)n * 4 * sizeof(double));
        memcpy(sm.slots, slots, (size_t)n * sizeof(int32_t));
        LAUNCH(c, "crop_resize", launch_crop_resize_small((const uint8_t *)c->frame.p, c->fH, c->fW, c->fstride, sm, c->bank, c->stream));
    } else {
        size_t bb = (size_t)n * 4 * sizeof(double), sb = (size_t)n * sizeof(int32_t);
        CUDA_OK(c->ws_small.ensure(bb + sb + 64));
        double *dbox = (double *)c->ws_small.p;
        int32_t *dslots = (int32_t *)((char *)c->ws_small.p + bb);
        CUDA_OK(cudaMemcpyAsync(dbox, boxes, bb, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(dslots, slots, sb, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, "crop_resize", launch_crop_resize((const uint8_t *)c->frame.p, c->fH, c->fW, c->fstride, dbox, n, dslots, c->bank, c->stream));
    }
    if (host_out) {
        int rc2 = bank_copy_runs(c, slots, n, host_out, false);
        if (rc2) return rc2;
    }
    // Opt-in (busca_set_option "defer_crop_copies"): return once the gather and its device->host copy are ENQUEUED.  The bank slot is
    // valid for every later call on this context (same stream); the HOST bytes are valid after the next call that waits for the stream
    // (busca_associate, busca_sync, any matrix call ...).  The adapters only store the crops between the two (byte_tracker.py:468-479).
    if (c->defer_crop_copies && !c->profiling) return BUSCA_OK;
    CUDA_OK((n <= 8) ? stream_wait_short(c) : cudaStreamSynchronize(c->stream));
    prof_collect(c);
    return BUSCA_OK;
}

extern "C" int busca_bank_upload(busca_ctx *c, const uint8_t *patches, int32_t n, const int32_t *slots) {
    if (!c || n < 0 || (n > 0 && (!patches || !slots))) return set_err(BUSCA_ERR_ARG, "bad argument");
    int rc = check_slots(c, slots, n, false);
    if (rc) return rc;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    rc = bank_copy_runs(c, slots, n, const_cast<uint8_t *>(patches), true);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}

extern "C" int busca_bank_download(busca_ctx *c, const int32_t *slots, int32_t n, uint8_t *host_out) {
    if (!c || n < 0 || (n > 0 && (!host_out || !slots))) return set_err(BUSCA_ERR_ARG, "bad argument");
    int rc = check_slots(c, slots, n, false);
    if (rc) return rc;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    rc = bank_copy_runs(c, slots, n, host_out, false);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
static int pair_matrix(busca_ctx *c, const double *a, int na, const double *b, int nb, double *out, int want_iou) {
    if (!c || na < 0 || nb < 0) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (na == 0 || nb == 0) return BUSCA_OK;
    if (!a || !b || !out) return set_err(BUSCA_ERR_ARG, "null pointer");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    size_t ab = (size_t)na * 32, bb = (size_t)nb * 32, ob = (size_t)na * nb * 8;
    CUDA_OK(c->ws_small.ensure(ab + bb + ob + 64));
    double *da = (double *)c->ws_small.p, *db = da + (size_t)na * 4, *dout = db + (size_t)nb * 4;
    CUDA_OK(cudaMemcpyAsync(da, a, ab, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(db, b, bb, cudaMemcpyHostToDevice, c->stream));
    prof_reset(c);
    LAUNCH(c, want_iou ? "iou_matrix" : "center_distance", launch_pair_matrix(da, na, db, nb, dout, want_iou, c->stream));
    CUDA_OK(cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" int busca_center_distance(busca_ctx *c, const double *a, int32_t na, const double *b, int32_t nb, double *out) {
    return pair_matrix(c, a, na, b, nb, out, 0);
}
extern "C" int busca_iou(busca_ctx *c, const double *a, int32_t na, const double *b, int32_t nb, double *out) {
    return pair_matrix(c, a, na, b, nb, out, 1);
}

extern "C" int busca_detection_coverage(busca_ctx *c, const double *tlbr, int32_t n, int32_t H, int32_t W, int64_t *nonzero_out, double *bbox_areas_out) {
    if (!c || n < 0 || H <= 0 || W <= 0 || !nonzero_out || (n > 0 && !tlbr)) return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t bb = (size_t)(n > 0 ? n : 1) * 32, ab = (size_t)(n > 0 ? n : 1) * 8;
    CUDA_OK(c->ws_small.ensure(bb + ab + 64));
    double *dbox = (double *)c->ws_small.p, *dar = dbox + (size_t)(n > 0 ? n : 1) * 4;
    unsigned long long *dcnt = (unsigned long long *)((char *)c->ws_small.p + bb + ab);
    if (n > 0) CUDA_OK(cudaMemcpyAsync(dbox, tlbr, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(dcnt, 0, 8, c->stream));
    prof_reset(c);
    LAUNCH(c, "detection_coverage", launch_coverage(dbox, n, H, W, dcnt, dar, c->stream));
    unsigned long long cnt = 0;
    CUDA_OK(cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, c->stream));
    if (bbox_areas_out && n > 0) CUDA_OK(cudaMemcpyAsync(bbox_areas_out, dar, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    *nonzero_out = (int64_t)cnt;
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// host-tracker rounds: rounds.cu
// ------------------------------------------------------------------------------------------------
extern "C" int busca_kalman_predict(busca_ctx *c, const double *mean, const double *cov, const uint8_t *tracked, int32_t n,
                                    double *mean_out, double *cov_out) {
    if (!c || n < 0) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (n == 0) return BUSCA_OK;
    if (!mean || !cov || !mean_out || !cov_out) return set_err(BUSCA_ERR_ARG, "null pointer");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t mb = (size_t)n * 64, cb = (size_t)n * 512, tb = ((size_t)n + 63) / 64 * 64;
    CUDA_OK(c->ws_small.ensure(2 * mb + 2 * cb + tb + 64));
    double *dm = (double *)c->ws_small.p, *dc = dm + (size_t)n * 8, *dmo = dc + (size_t)n * 64, *dco = dmo + (size_t)n * 8;
    uint8_t *dt = (uint8_t *)(dco + (size_t)n * 64);
    CUDA_OK(cudaMemcpyAsync(dm, mean, mb, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dc, cov, cb, cudaMemcpyHostToDevice, c->stream));
    if (tracked) CUDA_OK(cudaMemcpyAsync(dt, tracked, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    prof_reset(c);
    LAUNCH(c, "kalman_predict", launch_kalman_predict(dm, dc, tracked ? dt : nullptr, n, dmo, dco, c->stream));
    CUDA_OK(cudaMemcpyAsync(mean_out, dmo, mb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(cov_out, dco, cb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}

extern "C" int busca_kalman_update(busca_ctx *c, const double *mean, const double *cov, const double *xyah, int32_t n, double *mean_out,
                                   double *cov_out) {
    if (!c || n < 0) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (n == 0) return BUSCA_OK;
    if (!mean || !cov || !xyah || !mean_out || !cov_out) return set_err(BUSCA_ERR_ARG, "null pointer");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t mb = (size_t)n * 64, cb = (size_t)n * 512, zb = (size_t)n * 32;
    CUDA_OK(c->ws_small.ensure(2 * mb + 2 * cb + zb + 64));
    double *dm = (double *)c->ws_small.p, *dc = dm + (size_t)n * 8, *dmo = dc + (size_t)n * 64, *dco = dmo + (size_t)n * 8,
           *dz = dco + (size_t)n * 64;
    CUDA_OK(cudaMemcpyAsync(dm, mean, mb, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dc, cov, cb, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dz, xyah, zb, cudaMemcpyHostToDevice, c->stream));
    prof_reset(c);
    LAUNCH(c, "kalman_update", launch_kalman_update(dm, dc, dz, n, dmo, dco, c->stream));
    CUDA_OK(cudaMemcpyAsync(mean_out, dmo, mb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(cov_out, dco, cb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}

// One association round without a host visit between its stages: cost matrix (IoU distance, optionally fused with the detection
// scores) -> assignment with a cost limit.  cost_in != NULL solves a caller-supplied matrix instead (busca_linear_assignment).
static int match_round(busca_ctx *c, const double *a, int32_t na, const double *b, int32_t nb, const double *score, const double *cost_in,
                       double limit, int32_t *x, int32_t *y, double *cost_out) {
    if (!c || na < 0 || nb < 0) return set_err(BUSCA_ERR_ARG, "bad argument");
    if ((na > 0 && !x) || (nb > 0 && !y)) return set_err(BUSCA_ERR_ARG, "null pointer");
    if (!(limit == limit) || limit >= 1e290 || limit <= -1e290)           // every row needs its finite 'unassigned' option: the solver terminates on it
        return set_err(BUSCA_ERR_ARG, "assignment: cost_limit must be finite (lap.lapjv without a limit needs a square matrix; the trackers always pass one)");
    if (na == 0 || nb == 0) {                                    // matching.linear_assignment: empty matrix -> everything unmatched
        for (int i = 0; i < na; ++i) x[i] = -1;
        for (int j = 0; j < nb; ++j) y[j] = -1;
        return BUSCA_OK;
    }
    if (!cost_in && (!a || !b)) return set_err(BUSCA_ERR_ARG, "null pointer");
    if (assignment_smem_bytes(na, nb) > 200 * 1024) return set_err(BUSCA_ERR_ARG, "assignment: rows + columns exceed the shared-memory solver (about 5000)");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t ab = (size_t)na * 32, bb = (size_t)nb * 32, sb = (size_t)nb * 8, cb = (size_t)na * nb * 8;
    CUDA_OK(c->ws_small.ensure(ab + bb + sb + cb + (size_t)(na + nb) * 4 + 64));
    double *da = (double *)c->ws_small.p, *db = da + (size_t)na * 4, *ds = db + (size_t)nb * 4, *dcost = ds + nb;
    int *dx = (int *)(dcost + (size_t)na * nb), *dy = dx + na;
    prof_reset(c);
    if (cost_in) {
        CUDA_OK(cudaMemcpyAsync(dcost, cost_in, cb, cudaMemcpyHostToDevice, c->stream));
    } else {
        CUDA_OK(cudaMemcpyAsync(da, a, ab, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(db, b, bb, cudaMemcpyHostToDevice, c->stream));
        if (score) CUDA_OK(cudaMemcpyAsync(ds, score, sb, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, "match_cost", launch_match_cost(da, na, db, nb, score ? ds : nullptr, dcost, c->stream));
    }
    LAUNCH(c, "assignment", launch_assignment(dcost, na, nb, limit, dx, dy, c->stream));
    CUDA_OK(cudaMemcpyAsync(x, dx, (size_t)na * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(y, dy, (size_t)nb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cost_out) CUDA_OK(cudaMemcpyAsync(cost_out, dcost, cb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" int busca_match_round(busca_ctx *c, const double *a_tlbr, int32_t na, const double *b_tlbr, int32_t nb, const double *b_score,
                                 double cost_limit, int32_t *x, int32_t *y, double *cost_out) {
    return match_round(c, a_tlbr, na, b_tlbr, nb, b_score, nullptr, cost_limit, x, y, cost_out);
}
extern "C" int busca_linear_assignment(busca_ctx *c, const double *cost, int32_t n, int32_t m, double cost_limit, int32_t *x, int32_t *y) {
    if (n > 0 && m > 0 && !cost) return set_err(BUSCA_ERR_ARG, "null pointer");
    return match_round(c, nullptr, n, nullptr, m, nullptr, cost, cost_limit, x, y, nullptr);
}

extern "C" int busca_duplicate_tracks(busca_ctx *c, const double *a_tlbr, const int32_t *a_age, int32_t na, const double *b_tlbr,
                                      const int32_t *b_age, int32_t nb, double thresh, uint8_t *drop_a, uint8_t *drop_b) {
    if (!c || na < 0 || nb < 0) return set_err(BUSCA_ERR_ARG, "bad argument");
    if ((na > 0 && !drop_a) || (nb > 0 && !drop_b)) return set_err(BUSCA_ERR_ARG, "null pointer");
    for (int i = 0; i < na; ++i) drop_a[i] = 0;
    for (int j = 0; j < nb; ++j) drop_b[j] = 0;
    if (na == 0 || nb == 0) return BUSCA_OK;
    if (!a_tlbr || !b_tlbr || !a_age || !b_age) return set_err(BUSCA_ERR_ARG, "null pointer");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t ab = (size_t)na * 32, bb = (size_t)nb * 32, fa = ((size_t)na + 7) / 8 * 8, fb = ((size_t)nb + 7) / 8 * 8;
    CUDA_OK(c->ws_small.ensure(ab + bb + (size_t)(na + nb) * 4 + fa + fb + 64));
    double *da = (double *)c->ws_small.p, *db = da + (size_t)na * 4;
    int *dga = (int *)(db + (size_t)nb * 4), *dgb = dga + na;
    uint8_t *dfa = (uint8_t *)(dgb + nb), *dfb = dfa + fa;
    CUDA_OK(cudaMemcpyAsync(da, a_tlbr, ab, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(db, b_tlbr, bb, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dga, a_age, (size_t)na * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dgb, b_age, (size_t)nb * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(dfa, 0, fa + fb, c->stream));
    prof_reset(c);
    LAUNCH(c, "duplicate_tracks", launch_duplicates(da, dga, na, db, dgb, nb, thresh, dfa, dfb, c->stream));
    CUDA_OK(cudaMemcpyAsync(drop_a, dfa, (size_t)na, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(drop_b, dfb, (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}

// B frames (of B independent sequences: one tracker process usually owns several) in ONE launch: grid (T, B)
static int frame_geometry_batch(busca_ctx *c, int32_t B, const double *mean, const uint8_t *tracked, int32_t T0, const double *det_tlbr,
                                int32_t D, int32_t C, int32_t use_kalman, double *tlwh, double *tlbr, double *dist, double *iou,
                                int32_t *cand) {
    if (!c || B < 1 || T0 < 0 || D < 0 || C < 1 || !mean) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (T0 == 0) return BUSCA_OK;
    if ((long long)B * T0 > 0x7fffffffLL / 64 || B > 65535) return set_err(BUSCA_ERR_ARG, "batch too large");
    const int32_t T = B * T0;                                   // rows of every per-track array below
    CUDA_OK(cudaSetDevice(c->cfg.device));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 63) & ~(size_t)63; return o; };
    size_t o_mean = take((size_t)T * 64), o_trk = take(T), o_det = take((size_t)B * (D ? D : 1) * 32), o_tlwh = take((size_t)T * 32),
           o_tlbr = take((size_t)T * 32), o_dist = take((size_t)T * (D ? D : 1) * 8), o_iou = take((size_t)T * (D ? D : 1) * 8),
           o_cand = take((size_t)T * C * 4);
    CUDA_OK(c->ws_small.ensure(off));
    char *base = (char *)c->ws_small.p;
    CUDA_OK(cudaMemcpyAsync(base + o_mean, mean, (size_t)T * 64, cudaMemcpyHostToDevice, c->stream));
    if (tracked) CUDA_OK(cudaMemcpyAsync(base + o_trk, tracked, T, cudaMemcpyHostToDevice, c->stream));
    if (D) CUDA_OK(cudaMemcpyAsync(base + o_det, det_tlbr, (size_t)B * D * 32, cudaMemcpyHostToDevice, c->stream));
    GeomParams p{};
    p.T = T0; p.D = D; p.C = C; p.use_kalman = use_kalman; p.nbatch = B;
    p.mean = (const double *)(base + o_mean);
    p.tracked = tracked ? (const uint8_t *)(base + o_trk) : nullptr;
    p.det_tlbr = (const double *)(base + o_det);
    p.tlwh_out = (double *)(base + o_tlwh);
    p.tlbr_out = (double *)(base + o_tlbr);
    p.dist_out = dist ? (double *)(base + o_dist) : nullptr;
    p.iou_out = iou ? (double *)(base + o_iou) : nullptr;
    p.cand_out = cand ? (int *)(base + o_cand) : nullptr;
    prof_reset(c);
    LAUNCH(c, "frame_geometry", launch_frame_geometry(p, c->stream));
    if (tlwh) CUDA_OK(cudaMemcpyAsync(tlwh, base + o_tlwh, (size_t)T * 32, cudaMemcpyDeviceToHost, c->stream));
    if (tlbr) CUDA_OK(cudaMemcpyAsync(tlbr, base + o_tlbr, (size_t)T * 32, cudaMemcpyDeviceToHost, c->stream));
    if (dist && D) CUDA_OK(cudaMemcpyAsync(dist, base + o_dist, (size_t)T * D * 8, cudaMemcpyDeviceToHost, c->stream));
    if (iou && D) CUDA_OK(cudaMemcpyAsync(iou, base + o_iou, (size_t)T * D * 8, cudaMemcpyDeviceToHost, c->stream));
    if (cand) CUDA_OK(cudaMemcpyAsync(cand, base + o_cand, (size_t)T * C * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" int busca_frame_geometry(busca_ctx *c, const double *mean, const uint8_t *tracked, int32_t T, const double *det_tlbr,
                                    int32_t D, int32_t C, int32_t use_kalman, double *tlwh, double *tlbr, double *dist, double *iou,
                                    int32_t *cand) {
    return frame_geometry_batch(c, 1, mean, tracked, T, det_tlbr, D, C, use_kalman, tlwh, tlbr, dist, iou, cand);
}
extern "C" int busca_frame_geometry_batch(busca_ctx *c, int32_t B, const double *mean, const uint8_t *tracked, int32_t T, const double *det_tlbr,
                                          int32_t D, int32_t C, int32_t use_kalman, double *tlwh, double *tlbr, double *dist, double *iou,
                                          int32_t *cand) {
    return frame_geometry_batch(c, B, mean, tracked, T, det_tlbr, D, C, use_kalman, tlwh, tlbr, dist, iou, cand);
}

extern "C" int busca_motion_proposals(busca_ctx *c, const double *mean, const uint8_t *tracked, int32_t n, double *mean_out,
                                      double *tlwh, double *tlbr) {
    if (!c || n < 0 || (n > 0 && !mean)) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (n == 0) return BUSCA_OK;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    size_t need = (size_t)n * (64 + 64 + 64 + 32 + 32) + 256;
    CUDA_OK(c->ws_small.ensure(need));
    char *base = (char *)c->ws_small.p;
    double *dmean = (double *)base, *dmo = dmean + (size_t)n * 8, *dtlwh = dmo + (size_t)n * 8, *dtlbr = dtlwh + (size_t)n * 4;
    uint8_t *dtrk = (uint8_t *)(dtlbr + (size_t)n * 4);
    CUDA_OK(cudaMemcpyAsync(dmean, mean, (size_t)n * 64, cudaMemcpyHostToDevice, c->stream));
    if (tracked) CUDA_OK(cudaMemcpyAsync(dtrk, tracked, n, cudaMemcpyHostToDevice, c->stream));
    GeomParams p{};
    p.T = n; p.D = 0; p.C = 1; p.nbatch = 1;
    p.mean = dmean; p.tracked = tracked ? dtrk : nullptr; p.det_tlbr = dmean;
    p.mean_out = dmo; p.tlwh_out = dtlwh; p.tlbr_out = dtlbr;
    prof_reset(c);
    LAUNCH(c, "frame_geometry", launch_frame_geometry(p, c->stream));
    if (mean_out) CUDA_OK(cudaMemcpyAsync(mean_out, dmo, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
    if (tlwh) CUDA_OK(cudaMemcpyAsync(tlwh, dtlwh, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
    if (tlbr) CUDA_OK(cudaMemcpyAsync(tlbr, dtlbr, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(stream_wait_short(c));
    prof_collect(c);
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// ReID forward on one BatchNorm batch (device pointers)
// ------------------------------------------------------------------------------------------------
// tcgen05 path (bf16): per bottleneck  conv1 -> conv2 -> conv3 (statistics only) [-> downsample (statistics only)] -> conv3 again
// with BN3 + identity/downsample + ReLU in its epilogue.  BN + ReLU of conv1 / conv2 is applied to the consumer's A tile in
// shared memory, so no elementwise pass and no raw conv3 / downsample tensor ever touches HBM.
// One BatchNorm batch as the encoder runs it: the distinct images, their multiplicities, and the size of the batch the
// reference stacks (the divisor of every batch statistic).
struct ReidBatch {
    const int32_t *slots = nullptr;   // device [n]
    int n = 0;                        // images the encoder runs on
    int n_total = 0;                  // images of the stacked batch
    const float *weight = nullptr;    // device [n] multiplicities, null = all 1 (n == n_total)
    const int32_t *map = nullptr;     // device [n_total]: row of slots[] / of the embeddings for every stacked image
    int n_single = 0;                 // the first n_single images have multiplicity 1 (dedup_partition_kernel); == n without weights
};

static int reid_forward_tc(busca_ctx *c, const ReidBatch &rb, float *d_emb) {
    const int32_t *d_slots = rb.slots;
    const int N = rb.n;
    const long long NT = rb.n_total;
    const float *img_w = rb.weight;
    const size_t big = (size_t)786432 * N * 2;
    // two ping-pong block tensors (96x32x256 per patch) + the two bottleneck intermediates: 4.3 MB per patch, so the 30,000 distinct
    // patches of BASELINE.json configs[4] (1000 tracks x 30 history frames) need 130 GB of the 180 GB part.  The stem's im2col entries
    // live in X0 and its raw output in X1 (both dead before the first bottleneck writes them).
    const size_t total = 2 * big + (size_t)393216 * N * 2 + (size_t)196608 * N * 2 + (size_t)N * 2048 * 4 + (size_t)N * 512 * 4 + 2048;
    cudaError_t e = c->ws_reid.ensure(total);
    if (e != cudaSuccess) return set_err(BUSCA_ERR_NOMEM, "ReID workspace for %d patches (%.1f GB): %s", N, total / 1e9, cudaGetErrorString(e));
    char *base = (char *)c->ws_reid.p;
    void *X0 = base, *X1 = base + big;
    void *R1 = base + 2 * big, *R2 = (char *)R1 + (size_t)393216 * N * 2;
    float *pooled = (float *)((char *)R2 + (size_t)196608 * N * 2);
    float *emb_u = rb.map ? pooled + (((size_t)N * 2048 + 63) & ~(size_t)63) : d_emb;      // embeddings of the distinct images
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemsetAsync(c->stats_pool, 0, c->stats_bytes, s));
    // Gram-matrix partials of the (at most 20) statistics passes of one call: one arena, zeroed once
    const size_t gram_arena = c->gram ? 20 * ((size_t)256 * 256 * 4 + 1024) : 0;
    size_t gram_cursor = 0;
    if (gram_arena) {
        CUDA_OK(c->ws_gram.ensure(gram_arena));
        CUDA_OK(cudaMemsetAsync(c->ws_gram.p, 0, gram_arena, s));
    }
    ConvLayer &stem = c->convs[0];
    if (stem_tc_scratch_bytes(N) > big) return set_err(BUSCA_ERR_STATE, "stem scratch does not fit");
    if (c->profiling) { c->next_flops = 2.0 * N * 192 * 64 * 64.0 * 147; c->next_kernel = "conv_tc_kernel<64, 128, 0>"; }
    LAUNCH(c, "stem_conv7x7", launch_stem_tc(c->bank, d_slots, N, c->lut, stem.w16, X0, X1, stem.stats, img_w, s));
    if (c->profiling) c->prof.back().kernel = conv_tc_last_kernel();
    LAUNCH(c, "bn_finalize", launch_bn_finalize(stem, NT * 192 * 64, s));
    LAUNCH(c, "bn_relu_maxpool", launch_bn_relu_maxpool(X1, X0, N, 192, 64, 64, stem.scale, stem.shift, 1, s));
    void *x = X0, *other = X1;
    int H = 96, W = 32;
    size_t ci = 1;
    const int blocks[4] = {3, 4, 6, 3};
    auto conv = [&](ConvLayer &L, const ConvArgs &a, const ConvTcOpts &o, const char *name) -> int {
        char nm[96];
        if (c->profiling) {
            snprintf(nm, sizeof(nm), "%s_tc[%d>%d s%d %dx%d%s]", name, L.cin, L.cout, L.stride, a.H, a.W,
                     o.mode == TC_MODE_STATS ? " stats" : (o.mode == TC_MODE_FINAL ? (o.ds ? " final+ds" : " final") : ""));
            const bool dual = o.mode == TC_MODE_FINAL && o.ds;
            c->next_xflops = 2.0 * a.N * a.Ho * a.Wo * (double)L.cout * ((double)L.cin * L.k * L.k + (dual ? (double)o.ds->cin : 0.0));
            c->next_flops = o.mode == TC_MODE_STATS ? 0.0 : c->next_xflops;     // a statistics-only pass recomputes a GEMM the FINAL pass is credited for
            char kn[64];
            snprintf(kn, sizeof(kn), "conv_tc_kernel<%d, 128, %d>", L.cout >= 256 ? 256 : (L.cout >= 128 ? 128 : 64), dual ? 1 : 0);
            c->next_kernel = kn;
        }
        LAUNCH(c, c->profiling ? nm : name, launch_conv_tc(L, a, o, s));
        if (c->profiling) c->prof.back().kernel = conv_tc_last_kernel();   // the instantiation actually launched (resident weights or not)
        return BUSCA_OK;
    };
    int rc;
    // Batch statistics of a 1x1 convolution whose output is never stored (conv3, downsample): the Gram matrix of its input for the images
    // of multiplicity 1 (Cin <= 256), the weighted statistics-only GEMM pass for the repeated ones (and for Cin > 256).
    const int n_gram = c->gram ? (rb.n_single / 8) * 8 : 0;                // whole pixel tiles (a tile holds up to 8 images)
    auto stats_pass = [&](ConvLayer &Lc, const ConvArgs &a, const float **gram_G, const float **gram_m, const char *name) -> int {
        int ng = (Lc.k == 1 && (Lc.cin == 64 || Lc.cin == 128 || Lc.cin == 256)) ? n_gram : 0;
        if (ng > 0) {
            // partial G [C*C] and m [C] of this pass: a fresh, zeroed slice of the Gram arena (one memset per ReID call, above)
            const size_t C2 = (size_t)Lc.cin * Lc.cin;
            const size_t o_sp = (C2 * 4 + 255) & ~(size_t)255, slice = o_sp + (((size_t)Lc.cin * 4 + 255) & ~(size_t)255);
            if (gram_cursor + slice > gram_arena) return set_err(BUSCA_ERR_STATE, "Gram arena exhausted");
            char *gb = (char *)c->ws_gram.p + gram_cursor;
            gram_cursor += slice;
            ConvArgs ag = a;
            ag.N = ng; ag.img_w = nullptr;
            int grid = 0;
            if (c->profiling) {
                c->next_xflops = 2.0 * ng * a.Ho * a.Wo * (double)Lc.cin * (Lc.cin + 16.0) + 2.0 * (double)Lc.cout * Lc.cin * Lc.cin;
                c->next_flops = 0.0;
            }
            char nm[96];
            snprintf(nm, sizeof(nm), "gram_stats[%d>%d s%d %dx%d]", Lc.cin, Lc.cout, Lc.stride, a.H, a.W);
            LAUNCH(c, c->profiling ? nm : "gram_stats", launch_gram_stats(Lc, ag, (float *)gb, (float *)(gb + o_sp), &grid, s));
            if (c->profiling) c->prof.back().kernel = conv_tc_last_kernel();
            *gram_G = (const float *)gb;                             // the quadratic forms are taken by bn_fold_final
            *gram_m = (const float *)(gb + o_sp);
        }
        if (a.N - ng > 0) {
            ConvArgs as = a;
            as.N = a.N - ng;
            as.in = (const char *)a.in + (size_t)ng * a.H * a.W * Lc.cin * 2;
            as.img_w = a.img_w ? a.img_w + ng : nullptr;
            ConvTcOpts st{};
            st.mode = TC_MODE_STATS;
            int rc2 = conv(Lc, as, st, name);
            if (rc2) return rc2;
        }
        return BUSCA_OK;
    };
    for (int li = 0; li < 4; ++li)
        for (int b = 0; b < blocks[li]; ++b) {
            ConvLayer &c1 = c->convs[ci], &c2 = c->convs[ci + 1], &c3 = c->convs[ci + 2];
            const int st = c2.stride, Ho = H / st, Wo = W / st;
            ConvArgs a{};
            ConvTcOpts raw{}, stats{}, fin{};
            stats.mode = TC_MODE_STATS;
            fin.mode = TC_MODE_FINAL;
            a.N = N; a.img_w = img_w;
            a.in = x; a.out = R1; a.H = H; a.W = W; a.Ho = H; a.Wo = W; a.in_scale = nullptr; a.in_shift = nullptr;
            if ((rc = conv(c1, a, raw, "conv1x1"))) return rc;
            LAUNCH(c, "bn_finalize_fold", launch_bn_finalize_fold(c1, NT * H * W, c2, s));
            a.in = R1; a.out = R2; a.Ho = Ho; a.Wo = Wo; a.in_xf = c2.xf;
            if ((rc = conv(c2, a, raw, "conv3x3"))) return rc;
            LAUNCH(c, "bn_finalize_fold", launch_bn_finalize_fold(c2, NT * Ho * Wo, c3, s));
            ConvArgs a3{};
            a3.N = N; a3.img_w = img_w; a3.in = R2; a3.out = other; a3.H = Ho; a3.W = Wo; a3.Ho = Ho; a3.Wo = Wo; a3.in_xf = c3.xf;
            FoldFinalArgs ff{};                                  // BN3 (and the downsample BN): scales into the FINAL pass's weights
            ff.L = &c3; ff.in_scale = c2.scale; ff.count = NT * Ho * Wo; ff.shift_out = c3.shift;
            if ((rc = stats_pass(c3, a3, &ff.gram_G, &ff.gram_m, "conv1x1"))) return rc;
            fin.e_shift = c3.shift;
            if (b == 0) {
                ConvLayer &ds = c->convs[ci + 3];
                ConvArgs ad{};
                ad.N = N; ad.img_w = img_w; ad.in = x; ad.out = other; ad.H = H; ad.W = W; ad.Ho = Ho; ad.Wo = Wo;
                ff.ds = &ds;
                if ((rc = stats_pass(ds, ad, &ff.ds_gram_G, &ff.ds_gram_m, "conv1x1"))) return rc;
                fin.ds = &ds; fin.ds_in = x; fin.ds_H = H; fin.ds_W = W;
                ci += 4;
            } else {
                fin.idt = x;
                ci += 3;
            }
            LAUNCH(c, "bn_fold_final", launch_bn_fold_final(ff, s));
            if ((rc = conv(c3, a3, fin, "conv1x1"))) return rc;
            void *t = x; x = other; other = t;
            H = Ho; W = Wo;
        }
    LAUNCH(c, "global_maxpool", launch_global_maxpool(x, pooled, N, H * W, 2048, 1, s));
    LinearArgs la{};
    la.A = pooled; la.W = c->red_w; la.bias = c->red_b; la.residual = nullptr; la.out = emb_u; la.M = N; la.N = 512; la.K = 2048; la.alpha = 1.f; la.act = 0;
    static const bool red_tc = !(getenv("BUSCA_RED_TC") && getenv("BUSCA_RED_TC")[0] == '0');
    if (red_tc) {
        // the pooled features are maxima of bf16 activations, so the cast is exact; `other` is free after the last block
        LAUNCH(c, "cast_bf16", launch_cast_bf16(pooled, other, (long long)N * 2048, s));
        LAUNCH(c, "linear", launch_linear_tc(other, c->red_w16, la, s));
    } else {
        LAUNCH(c, "linear", launch_linear_f32(la, s));
    }
    LAUNCH(c, "l2norm", launch_l2norm_rows(emb_u, N, 512, s));
    if (rb.map) LAUNCH(c, "gather_rows", launch_gather_rows(emb_u, rb.map, d_emb, rb.n_total, 512, s));
    c->reid_images_run += N;
    c->reid_images_total += NT;
    return BUSCA_OK;
}

// Plan a batch: enqueue the duplicate elimination of `d_slots` (plan index 0 or 1) and the copy of the distinct count.
// Without dedup (fp32 parity mode keeps the reference's summation over the stacked batch) the plan is the identity.
static int reid_plan(busca_ctx *c, const int32_t *d_slots, int N, int which, ReidBatch *rb) {
    *rb = ReidBatch{};
    rb->slots = d_slots; rb->n = N; rb->n_total = N; rb->n_single = N;
    if (!(c->use_tc && c->dedup) || N <= 1) return BUSCA_OK;
    const size_t per = (((size_t)N * 4 + 255) & ~(size_t)255);
    CUDA_OK(c->ws_dedup[which].ensure(6 * per + 256));
    char *b = (char *)c->ws_dedup[which].p;
    int32_t *uniq = (int32_t *)b, *map = (int32_t *)(b + per);
    float *w = (float *)(b + 2 * per);
    int32_t *tmp_u = (int32_t *)(b + 3 * per), *newpos = (int32_t *)(b + 5 * per);
    float *tmp_w = (float *)(b + 4 * per);
    int *nu = (int *)(b + 6 * per);
    LAUNCH(c, "dedup_slots", launch_dedup_slots(d_slots, N, c->dedup_table, uniq, map, w, nu, c->stream));
    // multiplicity-1 images first: they take the unweighted Gram-matrix statistics path, the repeated ones the weighted pass
    LAUNCH(c, "dedup_partition", launch_dedup_partition(uniq, w, map, N, nu, tmp_u, tmp_w, newpos, nu + 1, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->h_nuniq + 2 * which, nu, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    rb->slots = uniq; rb->map = map; rb->weight = w; rb->n = -1 - which;      // resolved by reid_plan_wait
    return BUSCA_OK;
}
// Wait (host) for the distinct counts of the planned batches; everything enqueued before the plans has then run.
static int reid_plan_wait(busca_ctx *c, ReidBatch *a, ReidBatch *b) {
    if ((a && a->n < 0) || (b && b->n < 0)) {
        CUDA_OK(cudaEventRecord(c->ev_plan, c->stream));
        CUDA_OK(cudaEventSynchronize(c->ev_plan));
        for (ReidBatch *r : {a, b})
            if (r && r->n < 0) {
                const int which = -1 - r->n;
                r->n = c->h_nuniq[2 * which];
                r->n_single = c->h_nuniq[2 * which + 1];
                if (r->n <= 0 || r->n > r->n_total || r->n_single < 0 || r->n_single > r->n)
                    return set_err(BUSCA_ERR_STATE, "dedup returned %d distinct images (%d single) of %d", r->n, r->n_single, r->n_total);
                if (r->n == r->n_total) { r->weight = nullptr; r->map = nullptr; r->n_single = r->n; }   // nothing repeated: uniq == the stacked batch, in order
            }
    }
    return BUSCA_OK;
}

static int reid_forward_dev(busca_ctx *c, const ReidBatch &rb, float *d_emb) {
    if (rb.n_total <= 0) return BUSCA_OK;
    if (c->use_tc) return reid_forward_tc(c, rb, d_emb);
    const int32_t *d_slots = rb.slots;
    const int N = rb.n;
    const int bf16 = c->cfg.precision == BUSCA_PREC_BF16;
    const size_t es = bf16 ? 2 : 4;
    const size_t big = (size_t)786432 * N * es;
    const size_t total = 4 * big + (size_t)393216 * N * es + (size_t)196608 * N * es + (size_t)N * 2048 * 4 + 1024;
    cudaError_t e = c->ws_reid.ensure(total);
    if (e != cudaSuccess) return set_err(BUSCA_ERR_NOMEM, "ReID workspace for %d patches (%.1f GB): %s", N, total / 1e9, cudaGetErrorString(e));
    char *base = (char *)c->ws_reid.p;
    void *A = base, *B = base + big, *R3 = base + 2 * big, *RDS = base + 3 * big;
    void *R1 = base + 4 * big, *R2 = (char *)R1 + (size_t)393216 * N * es;
    float *pooled = (float *)((char *)R2 + (size_t)196608 * N * es);
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemsetAsync(c->stats_pool, 0, c->stats_bytes, s));

    ConvLayer &stem = c->convs[0];
    LAUNCH(c, "stem_conv7x7", launch_stem(c->bank, d_slots, N, c->lut, stem, R3, bf16, s));
    LAUNCH(c, "bn_finalize", launch_bn_finalize(stem, (long long)N * 192 * 64, s));
    LAUNCH(c, "bn_relu_maxpool", launch_bn_relu_maxpool(R3, A, N, 192, 64, 64, stem.scale, stem.shift, bf16, s));
    void *x = A, *other = B;
    int H = 96, W = 32;
    size_t ci = 1;
    const int blocks[4] = {3, 4, 6, 3};
    // SIMT path (fp32 parity mode, or bf16 storage with BUSCA_CONV=simt): the producer's BN+ReLU is applied while loading A
    auto conv = [&](ConvLayer &L, ConvArgs a, const char *name) -> int {
        LAUNCH(c, name, launch_conv_simt(L, a, bf16, s));
        return BUSCA_OK;
    };
    int rc;
    for (int li = 0; li < 4; ++li)
        for (int b = 0; b < blocks[li]; ++b) {
            ConvLayer &c1 = c->convs[ci], &c2 = c->convs[ci + 1], &c3 = c->convs[ci + 2];
            const int st = c2.stride, Ho = H / st, Wo = W / st;
            ConvArgs a{};
            a.N = N;
            a.in = x; a.out = R1; a.H = H; a.W = W; a.Ho = H; a.Wo = W; a.in_scale = nullptr; a.in_shift = nullptr;
            if ((rc = conv(c1, a, "conv1x1"))) return rc;
            LAUNCH(c, "bn_finalize", launch_bn_finalize(c1, (long long)N * H * W, s));
            a.in = R1; a.out = R2; a.Ho = Ho; a.Wo = Wo; a.in_scale = c1.scale; a.in_shift = c1.shift;
            if ((rc = conv(c2, a, "conv3x3"))) return rc;
            LAUNCH(c, "bn_finalize", launch_bn_finalize(c2, (long long)N * Ho * Wo, s));
            a.in = R2; a.out = R3; a.H = Ho; a.W = Wo; a.in_scale = c2.scale; a.in_shift = c2.shift;
            if ((rc = conv(c3, a, "conv1x1"))) return rc;
            LAUNCH(c, "bn_finalize", launch_bn_finalize(c3, (long long)N * Ho * Wo, s));
            const long long rows = (long long)N * Ho * Wo;
            if (b == 0) {
                ConvLayer &ds = c->convs[ci + 3];
                a.in = x; a.out = RDS; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.in_scale = nullptr; a.in_shift = nullptr;
                if ((rc = conv(ds, a, "conv1x1"))) return rc;
                LAUNCH(c, "bn_finalize", launch_bn_finalize(ds, rows, s));
                LAUNCH(c, "bn_add_relu", launch_bn_add_relu(R3, c3.scale, c3.shift, RDS, ds.scale, ds.shift, other, rows, c3.cout, bf16, s));
                void *t = x; x = other; other = t;
                ci += 4;
            } else {
                LAUNCH(c, "bn_add_relu", launch_bn_add_relu(R3, c3.scale, c3.shift, x, nullptr, nullptr, x, rows, c3.cout, bf16, s));
                ci += 3;
            }
            H = Ho; W = Wo;
        }
    LAUNCH(c, "global_maxpool", launch_global_maxpool(x, pooled, N, H * W, 2048, bf16, s));
    LinearArgs la{};
    la.A = pooled; la.W = c->red_w; la.bias = c->red_b; la.residual = nullptr; la.out = d_emb; la.M = N; la.N = 512; la.K = 2048; la.alpha = 1.f; la.act = 0;
    LAUNCH(c, "linear", launch_linear_f32(la, s));
    LAUNCH(c, "l2norm", launch_l2norm_rows(d_emb, N, 512, s));
    c->reid_images_run += N;
    c->reid_images_total += N;
    return BUSCA_OK;
}

extern "C" int busca_reid_embed(busca_ctx *c, const int32_t *slots, int32_t n, float *out) {
    if (!c || n < 0 || (n > 0 && (!slots || !out))) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (!c->finalized) return set_err(BUSCA_ERR_STATE, "weights not finalized");
    if (n == 0) return BUSCA_OK;
    int rc = check_slots(c, slots, n, true);
    if (rc) return rc;
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(c->ws_io.ensure((size_t)n * 4 + (size_t)n * 512 * 4 + 256));
    int32_t *dsl = (int32_t *)c->ws_io.p;
    float *demb = (float *)((char *)c->ws_io.p + (((size_t)n * 4 + 255) & ~(size_t)255));
    CUDA_OK(cudaMemcpyAsync(dsl, slots, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    prof_reset(c);
    ReidBatch rb;
    if ((rc = reid_plan(c, dsl, n, 0, &rb))) return rc;
    if ((rc = reid_plan_wait(c, &rb, nullptr))) return rc;
    rc = reid_forward_dev(c, rb, demb);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(out, demb, (size_t)n * 512 * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// Decision Transformer (device pointers).  Workspace carved from ws_tr.
// ------------------------------------------------------------------------------------------------
struct TrOut {
    float *logits, *probs, *cand_rows, *mem_logits, *input_seq;   // device, may be null except logits/probs
};
static int transformer_dev(busca_ctx *c, int T, int L, int C, const float *mem_emb, const float *can_emb, const int32_t *idx, TrOut o) {
    const int S = L + 2 * (C + 2), d = c->cfg.d_model, ff = c->cfg.ff_size;
    const size_t rows = (size_t)T * S;
    size_t off = 0;
    auto take = [&](size_t n_floats) { size_t r = off; off += (n_floats * 4 + 255) & ~(size_t)255; return r; };
    size_t o_me = take((size_t)T * L * d), o_ce = take((size_t)T * C * d), o_x = take(rows * d), o_y = take(rows * d), o_qkv = take(rows * 3 * d),
           o_att = take(rows * d), o_h = take(rows * ff), o_a16 = take(rows * (size_t)(ff > d ? ff : d) / 2 + 64);
    CUDA_OK(c->ws_tr.ensure(off));
    char *b = (char *)c->ws_tr.p;
    float *me = (float *)(b + o_me), *ce = (float *)(b + o_ce), *X = (float *)(b + o_x), *Y = (float *)(b + o_y), *QKV = (float *)(b + o_qkv),
          *ATT = (float *)(b + o_att), *Hd = (float *)(b + o_h);
    cudaStream_t s = c->stream;
    const float alpha = (float)sqrt((double)d);                    // * np.sqrt(self.dim_model), network.py:203-204
    // bf16 mode: the GEMMs run on the tensor cores (A cast to bf16 per call, fp32 accumulate / bias / residual / output)
    const bool tc = c->use_tc && c->tr_tc;
    void *A16 = b + o_a16;
    auto linear = [&](const LinearArgs &g, const void *w16) -> int {
        if (tc) {
            LAUNCH(c, "cast_bf16", launch_cast_bf16(g.A, A16, (long long)g.M * g.K, s));
            LAUNCH(c, "linear", launch_linear_tc(A16, w16, g, s));
        } else {
            LAUNCH(c, "linear", launch_linear_f32(g, s));
        }
        return BUSCA_OK;
    };
    int rc;
    LinearArgs la{};
    la.alpha = alpha; la.act = 0; la.residual = nullptr; la.W = c->enc_w; la.bias = c->enc_b; la.N = d; la.K = d;
    la.A = mem_emb; la.out = me; la.M = T * L;
    if ((rc = linear(la, c->enc_w16))) return rc;
    la.A = can_emb; la.out = ce; la.M = T * C;
    if ((rc = linear(la, c->enc_w16))) return rc;
    PeTables pe{c->pe_xy, c->pe_size, c->pe_t};
    LAUNCH(c, "build_tokens", launch_build_tokens(me, ce, c->sep, c->non, c->bad, idx, pe, T, L, C, X, s));
    if (o.input_seq) CUDA_OK(cudaMemcpyAsync(o.input_seq, X, rows * d * 4, cudaMemcpyDeviceToDevice, s));
    const int act = c->cfg.activation == BUSCA_ACT_GELU ? 2 : 1;
    if (tc) {
        // tensor-core path: five launches per layer, every cast / residual / LayerNorm in a GEMM epilogue (linear_tc.cu).  X16 = bf16 copy of
        // the token rows (A operand), ATT and the FFN hidden exist in bf16 only - the same rounding points as separate cast launches.
        void *X16 = A16, *ATT16 = (void *)ATT, *H16 = (void *)Hd;
        LAUNCH(c, "cast_bf16", launch_cast_bf16(X, X16, (long long)rows * d, s));
        for (auto &ly : c->layers) {
            LinearFusedArgs g{};
            g.M = (int)rows;
            g.epilogue = LE_F32; g.N = 3 * d; g.K = d; g.bias = ly.in_b; g.out_f32 = QKV;
            LAUNCH(c, "linear_qkv", launch_linear_fused(X16, ly.in_w16, g, s));
            LAUNCH(c, "attention", launch_attention(QKV, ATT16, 1, T, S, c->cfg.nhead, d / c->cfg.nhead, s));
            g = LinearFusedArgs{};
            g.M = (int)rows; g.epilogue = LE_LN; g.N = d; g.K = d; g.bias = ly.out_b; g.residual = X; g.gamma = ly.n1_g; g.beta = ly.n1_b; g.out_f32 = X; g.out_bf16 = X16;
            LAUNCH(c, "linear_out_ln", launch_linear_fused(ATT16, ly.out_w16, g, s));
            g = LinearFusedArgs{};
            g.M = (int)rows; g.epilogue = LE_BF16; g.N = ff; g.K = d; g.bias = ly.l1_b; g.act = act; g.out_bf16 = H16;
            LAUNCH(c, "linear_ffn1", launch_linear_fused(X16, ly.l1_w16, g, s));
            g = LinearFusedArgs{};
            g.M = (int)rows; g.epilogue = LE_LN; g.N = d; g.K = ff; g.bias = ly.l2_b; g.residual = X; g.gamma = ly.n2_g; g.beta = ly.n2_b; g.out_f32 = X; g.out_bf16 = X16;
            LAUNCH(c, "linear_ffn2_ln", launch_linear_fused(H16, ly.l2_w16, g, s));
        }
    } else
    for (auto &ly : c->layers) {
        LinearArgs g{};
        g.alpha = 1.f; g.M = (int)rows;
        g.A = X; g.W = ly.in_w; g.bias = ly.in_b; g.residual = nullptr; g.out = QKV; g.N = 3 * d; g.K = d; g.act = 0;
        if ((rc = linear(g, ly.in_w16))) return rc;
        LAUNCH(c, "attention", launch_attention(QKV, ATT, 0, T, S, c->cfg.nhead, d / c->cfg.nhead, s));
        g.A = ATT; g.W = ly.out_w; g.bias = ly.out_b; g.residual = X; g.out = Y; g.N = d; g.K = d;
        if ((rc = linear(g, ly.out_w16))) return rc;
        LAUNCH(c, "layernorm", launch_layernorm(Y, ly.n1_g, ly.n1_b, X, (int)rows, d, s));
        g.A = X; g.W = ly.l1_w; g.bias = ly.l1_b; g.residual = nullptr; g.out = Hd; g.N = ff; g.K = d; g.act = act;
        if ((rc = linear(g, ly.l1_w16))) return rc;
        g.A = Hd; g.W = ly.l2_w; g.bias = ly.l2_b; g.residual = X; g.out = Y; g.N = d; g.K = ff; g.act = 0;
        if ((rc = linear(g, ly.l2_w16))) return rc;
        LAUNCH(c, "layernorm", launch_layernorm(Y, ly.n2_g, ly.n2_b, X, (int)rows, d, s));
    }
    LAUNCH(c, "decoder", launch_decoder(X, T, S, L, C, c->dec_g, c->dec_b, c->dec_w, c->dec_bias, o.logits, o.probs, o.cand_rows, o.mem_logits, s));
    return BUSCA_OK;
}

// carve helper for io scratch
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; }
};

extern "C" int busca_transformer(busca_ctx *c, int32_t T, int32_t L, int32_t C, const float *mem_emb, const float *can_emb,
                                 const double *mem_ltwh, const double *can_ltwh, float *logits, float *probs, int32_t *pe_index,
                                 float *cand_rows, float *input_seq) {
    if (!c || T <= 0 || L < 1 || C < 1 || !mem_emb || !can_emb || !mem_ltwh || !can_ltwh) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (!c->finalized) return set_err(BUSCA_ERR_STATE, "weights not finalized");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const int S = L + 2 * (C + 2), nc = C + 2;
    Carver cv;
    size_t o_me = cv.take((size_t)T * L * 512 * 4), o_ce = cv.take((size_t)T * C * 512 * 4), o_mb = cv.take((size_t)T * L * 32), o_cb = cv.take((size_t)T * C * 32),
           o_idx = cv.take((size_t)T * S * 12), o_lg = cv.take((size_t)T * nc * 4), o_pr = cv.take((size_t)T * nc * 4), o_cr = cv.take((size_t)T * nc * 512 * 4),
           o_is = cv.take((size_t)T * S * 512 * 4);
    CUDA_OK(c->ws_io.ensure(cv.off));
    char *b = (char *)c->ws_io.p;
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemcpyAsync(b + o_me, mem_emb, (size_t)T * L * 512 * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_ce, can_emb, (size_t)T * C * 512 * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_mb, mem_ltwh, (size_t)T * L * 32, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_cb, can_ltwh, (size_t)T * C * 32, cudaMemcpyHostToDevice, s));
    prof_reset(c);
    LAUNCH(c, "pe_index", launch_pe_index((const double *)(b + o_mb), (const double *)(b + o_cb), T, L, C, c->cfg.sentinel_fp64, (int32_t *)(b + o_idx), s));
    TrOut o{(float *)(b + o_lg), (float *)(b + o_pr), cand_rows ? (float *)(b + o_cr) : nullptr, nullptr, input_seq ? (float *)(b + o_is) : nullptr};
    int rc = transformer_dev(c, T, L, C, (const float *)(b + o_me), (const float *)(b + o_ce), (const int32_t *)(b + o_idx), o);
    if (rc) return rc;
    if (logits) CUDA_OK(cudaMemcpyAsync(logits, b + o_lg, (size_t)T * nc * 4, cudaMemcpyDeviceToHost, s));
    if (probs) CUDA_OK(cudaMemcpyAsync(probs, b + o_pr, (size_t)T * nc * 4, cudaMemcpyDeviceToHost, s));
    if (pe_index) CUDA_OK(cudaMemcpyAsync(pe_index, b + o_idx, (size_t)T * S * 12, cudaMemcpyDeviceToHost, s));
    if (cand_rows) CUDA_OK(cudaMemcpyAsync(cand_rows, b + o_cr, (size_t)T * nc * 512 * 4, cudaMemcpyDeviceToHost, s));
    if (input_seq) CUDA_OK(cudaMemcpyAsync(input_seq, b + o_is, (size_t)T * S * 512 * 4, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    prof_collect(c);
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// association (host pointers in, host pointers out)
// ------------------------------------------------------------------------------------------------
extern "C" int busca_associate(busca_ctx *c, const busca_assoc_args *a) {
    if (!c || !a) return set_err(BUSCA_ERR_ARG, "null argument");
    if (!c->finalized) return set_err(BUSCA_ERR_STATE, "weights not finalized");
    const int T = a->T, D = a->D, L = a->L, C = a->C;
    if (T <= 0 || D < 0 || L < 1 || C < 1 || C + 2 > 30 || L + 2 * (C + 2) > 64) return set_err(BUSCA_ERR_ARG, "unsupported sizes T=%d D=%d L=%d C=%d", T, D, L, C);
    if (!a->mem_slots || !a->mem_ltwh || (D > 0 && (!a->det_slots || !a->det_ltwh || !a->dists))) return set_err(BUSCA_ERR_ARG, "null input");
    if (a->use_kalman && (!a->kal_slots || !a->kal_ltwh)) return set_err(BUSCA_ERR_ARG, "use_kalman without kalman inputs");
    if (D == 0 && !a->use_kalman) return set_err(BUSCA_ERR_ARG, "no detections and no kalman candidates (the reference returns None here)");
    int rc = check_slots(c, a->mem_slots, T * L, true);
    if (rc) return rc;
    if (D) { rc = check_slots(c, a->det_slots, D, true); if (rc) return rc; }
    if (a->use_kalman) { rc = check_slots(c, a->kal_slots, T, true); if (rc) return rc; }
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const int S = L + 2 * (C + 2), nc = C + 2, Dn = D ? D : 1;
    Carver cv;
    size_t o_ms = cv.take((size_t)T * L * 4), o_mb = cv.take((size_t)T * L * 32), o_ds = cv.take((size_t)Dn * 4), o_db = cv.take((size_t)Dn * 32),
           o_dist = cv.take((size_t)T * Dn * 8), o_ks = cv.take((size_t)T * 4), o_kb = cv.take((size_t)T * 32), o_cand = cv.take((size_t)T * C * 4),
           o_cb = cv.take((size_t)T * C * 32), o_cs = cv.take((size_t)T * C * 4), o_idx = cv.take((size_t)T * S * 12), o_me = cv.take((size_t)T * L * 512 * 4),
           o_ce = cv.take((size_t)T * C * 512 * 4), o_lg = cv.take((size_t)T * nc * 4), o_pr = cv.take((size_t)T * nc * 4),
           o_cr = cv.take((size_t)T * nc * 512 * 4), o_ml = cv.take((size_t)T * 512 * 4), o_is = cv.take((size_t)T * S * 512 * 4);
    CUDA_OK(c->ws_io.ensure(cv.off));
    char *b = (char *)c->ws_io.p;
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemcpyAsync(b + o_ms, a->mem_slots, (size_t)T * L * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_mb, a->mem_ltwh, (size_t)T * L * 32, cudaMemcpyHostToDevice, s));
    if (D) {
        CUDA_OK(cudaMemcpyAsync(b + o_ds, a->det_slots, (size_t)D * 4, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(b + o_db, a->det_ltwh, (size_t)D * 32, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(b + o_dist, a->dists, (size_t)T * D * 8, cudaMemcpyHostToDevice, s));
    }
    if (a->use_kalman) {
        CUDA_OK(cudaMemcpyAsync(b + o_ks, a->kal_slots, (size_t)T * 4, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(b + o_kb, a->kal_ltwh, (size_t)T * 32, cudaMemcpyHostToDevice, s));
    }
    prof_reset(c);
    GeomParams p{};
    p.T = T; p.D = D; p.C = C; p.use_kalman = a->use_kalman ? 1 : 0; p.nbatch = 1;
    p.mean = nullptr; p.trk_tlbr = (const double *)(b + o_mb);      // unused values: distances are given
    p.det_tlbr = (const double *)(b + o_db);
    p.dists_in = (const double *)(b + o_dist);
    p.cand_out = (int *)(b + o_cand);
    LAUNCH(c, "frame_geometry", launch_frame_geometry(p, s));
    LAUNCH(c, "assemble_candidates", launch_assemble_candidates((const int *)(b + o_cand), T, D, C, (const double *)(b + o_db), (const int32_t *)(b + o_ds),
                                                                 a->use_kalman ? (const double *)(b + o_kb) : nullptr,
                                                                 a->use_kalman ? (const int32_t *)(b + o_ks) : nullptr, (double *)(b + o_cb),
                                                                 (int32_t *)(b + o_cs), c->cfg.sentinel_fp64, s));
    LAUNCH(c, "pe_index", launch_pe_index((const double *)(b + o_mb), (const double *)(b + o_cb), T, L, C, c->cfg.sentinel_fp64, (int32_t *)(b + o_idx), s));
    ReidBatch rb_mem, rb_can;
    if ((rc = reid_plan(c, (const int32_t *)(b + o_ms), T * L, 0, &rb_mem))) return rc;
    if ((rc = reid_plan(c, (const int32_t *)(b + o_cs), T * C, 1, &rb_can))) return rc;
    if ((rc = reid_plan_wait(c, &rb_mem, &rb_can))) return rc;
    rc = reid_forward_dev(c, rb_mem, (float *)(b + o_me));
    if (rc) return rc;
    rc = reid_forward_dev(c, rb_can, (float *)(b + o_ce));
    if (rc) return rc;
    TrOut o{(float *)(b + o_lg), (float *)(b + o_pr), a->cand_rows ? (float *)(b + o_cr) : nullptr, a->mem_logits ? (float *)(b + o_ml) : nullptr,
            a->input_seq ? (float *)(b + o_is) : nullptr};
    rc = transformer_dev(c, T, L, C, (const float *)(b + o_me), (const float *)(b + o_ce), (const int32_t *)(b + o_idx), o);
    if (rc) return rc;
    if (a->probs) CUDA_OK(cudaMemcpyAsync(a->probs, b + o_pr, (size_t)T * nc * 4, cudaMemcpyDeviceToHost, s));
    if (a->logits) CUDA_OK(cudaMemcpyAsync(a->logits, b + o_lg, (size_t)T * nc * 4, cudaMemcpyDeviceToHost, s));
    if (a->cand) CUDA_OK(cudaMemcpyAsync(a->cand, b + o_cand, (size_t)T * C * 4, cudaMemcpyDeviceToHost, s));
    if (a->pe_index) CUDA_OK(cudaMemcpyAsync(a->pe_index, b + o_idx, (size_t)T * S * 12, cudaMemcpyDeviceToHost, s));
    if (a->mem_emb) CUDA_OK(cudaMemcpyAsync(a->mem_emb, b + o_me, (size_t)T * L * 512 * 4, cudaMemcpyDeviceToHost, s));
    if (a->can_emb) CUDA_OK(cudaMemcpyAsync(a->can_emb, b + o_ce, (size_t)T * C * 512 * 4, cudaMemcpyDeviceToHost, s));
    if (a->cand_rows) CUDA_OK(cudaMemcpyAsync(a->cand_rows, b + o_cr, (size_t)T * nc * 512 * 4, cudaMemcpyDeviceToHost, s));
    if (a->mem_logits) CUDA_OK(cudaMemcpyAsync(a->mem_logits, b + o_ml, (size_t)T * 512 * 4, cudaMemcpyDeviceToHost, s));
    if (a->input_seq) CUDA_OK(cudaMemcpyAsync(a->input_seq, b + o_is, (size_t)T * S * 512 * 4, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    prof_collect(c);
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident frame step
// ------------------------------------------------------------------------------------------------
__global__ void tlbr_to_ltwh_kernel(const double *__restrict__ tlbr, double *__restrict__ ltwh, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x1 = tlbr[4 * i], y1 = tlbr[4 * i + 1];
    ltwh[4 * i] = x1; ltwh[4 * i + 1] = y1;
    ltwh[4 * i + 2] = __dsub_rn(tlbr[4 * i + 2], x1);
    ltwh[4 * i + 3] = __dsub_rn(tlbr[4 * i + 3], y1);
}

extern "C" int busca_frame_step_dev(busca_ctx *c, const busca_step_args *a) {
    if (!c || !a) return set_err(BUSCA_ERR_ARG, "null argument");
    if (!c->finalized) return set_err(BUSCA_ERR_STATE, "weights not finalized");
    const int T = a->T, D = a->D, L = a->L, C = a->C;
    if (T <= 0 || D <= 0 || L < 1 || C < 1 || C + 2 > 30 || L + 2 * (C + 2) > 64) return set_err(BUSCA_ERR_ARG, "unsupported sizes");
    const uint8_t *frame = a->frame_dev ? a->frame_dev : (const uint8_t *)c->frame.p;
    const int fH = a->frame_dev ? a->frame_H : c->fH, fW = a->frame_dev ? a->frame_W : c->fW;
    const int64_t fstride = a->frame_dev ? (int64_t)a->frame_W * 3 : c->fstride;
    if (!frame || fH <= 0 || fW <= 0) return set_err(BUSCA_ERR_STATE, "no frame uploaded");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const int S = L + 2 * (C + 2), nc = C + 2;
    Carver cv;
    size_t o_tlwh = cv.take((size_t)T * 32), o_tlbr = cv.take((size_t)T * 32), o_dist = cv.take((size_t)T * D * 8), o_iou = cv.take((size_t)T * D * 8),
           o_cand = cv.take((size_t)T * C * 4), o_dl = cv.take((size_t)D * 32), o_cb = cv.take((size_t)T * C * 32), o_cs = cv.take((size_t)T * C * 4),
           o_idx = cv.take((size_t)T * S * 12), o_me = cv.take((size_t)T * L * 512 * 4), o_ce = cv.take((size_t)T * C * 512 * 4), o_lg = cv.take((size_t)T * nc * 4);
    CUDA_OK(c->ws_io.ensure(cv.off));
    char *b = (char *)c->ws_io.p;
    cudaStream_t s = c->stream;
    prof_reset(c);
    GeomParams p{};
    p.T = T; p.D = D; p.C = C; p.use_kalman = 1; p.nbatch = 1;
    p.mean = a->track_mean_dev; p.tracked = a->tracked_dev; p.det_tlbr = a->det_tlbr_dev;
    p.tlwh_out = (double *)(b + o_tlwh); p.tlbr_out = (double *)(b + o_tlbr);
    p.dist_out = (double *)(b + o_dist); p.iou_out = (double *)(b + o_iou); p.cand_out = (int *)(b + o_cand);
    LAUNCH(c, "frame_geometry", launch_frame_geometry(p, s));
    // crops of the D detections and the T motion proposals, straight from the resident frame
    LAUNCH(c, "crop_resize", launch_crop_resize(frame, fH, fW, fstride, a->det_tlbr_dev, D, a->det_slots_dev, c->bank, s));
    LAUNCH(c, "crop_resize", launch_crop_resize(frame, fH, fW, fstride, (const double *)(b + o_tlbr), T, a->kal_slots_dev, c->bank, s));
    prof_begin(c, "tlbr_to_ltwh");
    tlbr_to_ltwh_kernel<<<ceil_div(D, 128), 128, 0, s>>>(a->det_tlbr_dev, (double *)(b + o_dl), D);
    prof_end(c);
    c->launches++;
    LAUNCH(c, "assemble_candidates", launch_assemble_candidates((const int *)(b + o_cand), T, D, C, (const double *)(b + o_dl), a->det_slots_dev,
                                                                 (const double *)(b + o_tlwh), a->kal_slots_dev, (double *)(b + o_cb), (int32_t *)(b + o_cs),
                                                                 c->cfg.sentinel_fp64, s));
    LAUNCH(c, "pe_index", launch_pe_index(a->mem_ltwh_dev, (const double *)(b + o_cb), T, L, C, c->cfg.sentinel_fp64, (int32_t *)(b + o_idx), s));
    int rc;
    ReidBatch rb_mem, rb_can;
    if ((rc = reid_plan(c, a->mem_slots_dev, T * L, 0, &rb_mem))) return rc;
    if ((rc = reid_plan(c, (const int32_t *)(b + o_cs), T * C, 1, &rb_can))) return rc;
    if ((rc = reid_plan_wait(c, &rb_mem, &rb_can))) return rc;
    rc = reid_forward_dev(c, rb_mem, (float *)(b + o_me));
    if (rc) return rc;
    rc = reid_forward_dev(c, rb_can, (float *)(b + o_ce));
    if (rc) return rc;
    TrOut o{(float *)(b + o_lg), a->probs_dev, nullptr, nullptr, nullptr};
    rc = transformer_dev(c, T, L, C, (const float *)(b + o_me), (const float *)(b + o_ce), (const int32_t *)(b + o_idx), o);
    if (rc) return rc;
    if (a->keep_dev)
        LAUNCH(c, "decide", launch_decide(a->probs_dev, (const int *)(b + o_cand), a->reliable_dev, T, D, C, a->busca_thresh, a->select_highest,
                                          a->highest_min_thresh, a->keep_highest_value, a->keep_dev, s));
    if (a->cand_dev) CUDA_OK(cudaMemcpyAsync(a->cand_dev, b + o_cand, (size_t)T * C * 4, cudaMemcpyDeviceToDevice, s));
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// test hook: one convolution of the ReID network on caller-provided bf16 NHWC input
// ------------------------------------------------------------------------------------------------
extern "C" int busca_debug_conv_ex(busca_ctx *c, const busca_debug_conv_args *d) {
    if (!c || !d || !c->finalized || d->conv_index < 1 || d->conv_index >= (int)c->convs.size() || !d->in_bf16) return set_err(BUSCA_ERR_ARG, "bad argument");
    if (!d->use_tc && d->mode != 0) return set_err(BUSCA_ERR_ARG, "the SIMT kernel only has the raw mode");
    if (d->mode == 3 && !d->use_tc) return set_err(BUSCA_ERR_ARG, "mode 3 (Gram-matrix statistics) is a tensor-core mode");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    ConvLayer &L = c->convs[d->conv_index];
    const int N = d->N, H = d->H, W = d->W, Ho = H / L.stride, Wo = W / L.stride;
    const ConvLayer *DS = d->ds_index >= 0 ? &c->convs[d->ds_index] : nullptr;
    Carver cv;
    const size_t in_b = (size_t)N * H * W * L.cin * 2, out_b = (size_t)N * Ho * Wo * L.cout * 2;
    const size_t ds_b = DS ? (size_t)N * d->ds_H * d->ds_W * DS->cin * 2 : 0;
    size_t o_in = cv.take(in_b), o_out = cv.take(out_b), o_idt = cv.take(out_b), o_ds = cv.take(ds_b + 16), o_par = cv.take((size_t)(2 * L.cin + 4 * L.cout) * 4),
           o_w = cv.take((size_t)N * 4 + 16);
    CUDA_OK(c->ws_reid.ensure(cv.off + 512));
    char *b = (char *)c->ws_reid.p;
    float *par = (float *)(b + o_par);
    float *in_sc = par, *in_sh = par + L.cin, *e_sc = in_sh + L.cin, *e_sh = e_sc + L.cout, *d_sc = e_sh + L.cout, *d_sh = d_sc + L.cout;
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemcpyAsync(b + o_in, d->in_bf16, in_b, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(b + o_out, 0xff, out_b, s));
    CUDA_OK(cudaMemsetAsync(L.stats, 0, 2 * (size_t)L.cout * sizeof(double), s));
    ConvArgs a{};
    a.in = b + o_in; a.out = b + o_out; a.N = N; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo;
    if (d->img_w && d->use_tc) {
        CUDA_OK(cudaMemcpyAsync(b + o_w, d->img_w, (size_t)N * 4, cudaMemcpyHostToDevice, s));
        a.img_w = (const float *)(b + o_w);
    }
    if (d->in_scale) {
        CUDA_OK(cudaMemcpyAsync(in_sc, d->in_scale, (size_t)L.cin * 4, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(in_sh, d->in_shift, (size_t)L.cin * 4, cudaMemcpyHostToDevice, s));
        a.in_scale = in_sc; a.in_shift = in_sh;
        if (d->use_tc) {
            LAUNCH(c, "bn_fold", launch_bn_fold(in_sc, in_sh, L, s));
            a.in_xf = L.xf;
        }
    }
    ConvTcOpts o{};
    o.mode = d->mode;
    if (d->mode == TC_MODE_FINAL) {
        if (!d->e_scale || !d->e_shift) return set_err(BUSCA_ERR_ARG, "final mode needs e_scale / e_shift");
        CUDA_OK(cudaMemcpyAsync(e_sc, d->e_scale, (size_t)L.cout * 4, cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(e_sh, d->e_shift, (size_t)L.cout * 4, cudaMemcpyHostToDevice, s));
        FoldFinalArgs ff{};
        ff.L = &L; ff.in_scale = d->in_scale ? in_sc : nullptr; ff.scale = e_sc; ff.shift = e_sh; ff.shift_out = L.shift;
        o.e_shift = L.shift;
        if (DS) {
            if (!d->ds_in_bf16 || !d->ds_scale || !d->ds_shift) return set_err(BUSCA_ERR_ARG, "downsample inputs missing");
            CUDA_OK(cudaMemcpyAsync(b + o_ds, d->ds_in_bf16, ds_b, cudaMemcpyHostToDevice, s));
            CUDA_OK(cudaMemcpyAsync(d_sc, d->ds_scale, (size_t)L.cout * 4, cudaMemcpyHostToDevice, s));
            CUDA_OK(cudaMemcpyAsync(d_sh, d->ds_shift, (size_t)L.cout * 4, cudaMemcpyHostToDevice, s));
            o.ds = DS; o.ds_in = b + o_ds; o.ds_H = d->ds_H; o.ds_W = d->ds_W;
            ff.ds = DS; ff.ds_scale = d_sc; ff.ds_shift = d_sh;
        } else {
            if (!d->idt_bf16) return set_err(BUSCA_ERR_ARG, "final mode needs an identity tensor or a downsample conv");
            CUDA_OK(cudaMemcpyAsync(b + o_idt, d->idt_bf16, out_b, cudaMemcpyHostToDevice, s));
            o.idt = b + o_idt;
        }
        LAUNCH(c, "bn_fold_final", launch_bn_fold_final(ff, s));
    }
    prof_reset(c);
    if (d->mode == 3) {
        // Gram-matrix statistics of this (1x1) convolution on the given input: stats_out must equal mode 1's.  The quadratic forms are
        // part of bn_fold_final (which also folds the FINAL-pass weights of L from those statistics).
        if (!L.w16f) return set_err(BUSCA_ERR_ARG, "mode 3 is for the last conv of a bottleneck / a downsample conv");
        const size_t C2 = (size_t)L.cin * L.cin;
        const size_t o_sp = (C2 * 4 + 255) & ~(size_t)255;
        CUDA_OK(c->ws_gram.ensure(o_sp + (size_t)L.cin * 4 + 256));
        char *gb = (char *)c->ws_gram.p;
        int grid = 0;
        CUDA_OK(cudaMemsetAsync(gb, 0, o_sp + (size_t)L.cin * 4, s));
        LAUNCH(c, "gram_stats", launch_gram_stats(L, a, (float *)gb, (float *)(gb + o_sp), &grid, s));
        FoldFinalArgs ff{};
        ff.L = &L; ff.in_scale = d->in_scale ? in_sc : nullptr; ff.count = (long long)N * Ho * Wo; ff.shift_out = L.shift;
        ff.gram_G = (const float *)gb; ff.gram_m = (const float *)(gb + o_sp);
        LAUNCH(c, "bn_fold_final", launch_bn_fold_final(ff, s));
    } else if (d->use_tc) {
        LAUNCH(c, "conv_tc", launch_conv_tc(L, a, o, s));
        if (c->profiling) c->prof.back().kernel = conv_tc_last_kernel();      // which instantiation ran (the tests check the variant they asked for)
    } else LAUNCH(c, "conv_simt", launch_conv_simt(L, a, 1, s));
    if (d->out_bf16) CUDA_OK(cudaMemcpyAsync(d->out_bf16, b + o_out, out_b, cudaMemcpyDeviceToHost, s));
    if (d->stats_out) CUDA_OK(cudaMemcpyAsync(d->stats_out, L.stats, 2 * (size_t)L.cout * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" int busca_debug_conv(busca_ctx *c, int32_t conv_index, const uint16_t *in_bf16, int32_t N, int32_t H, int32_t W, int32_t use_tc,
                                uint16_t *out_bf16, double *stats_out) {
    busca_debug_conv_args d{};
    d.conv_index = conv_index; d.N = N; d.H = H; d.W = W; d.use_tc = use_tc; d.mode = 0; d.in_bf16 = in_bf16; d.ds_index = -1;
    d.out_bf16 = out_bf16; d.stats_out = stats_out;
    return busca_debug_conv_ex(c, &d);
}
// test hook: relu(BN(x)) followed by the 3x3/2 max-pool of the ReID stem on a caller-provided bf16 NHWC tensor
extern "C" int busca_debug_maxpool(busca_ctx *c, const uint16_t *in_bf16, int32_t N, int32_t H, int32_t W, int32_t Cc, const float *scale,
                                   const float *shift, uint16_t *out_bf16) {
    if (!c || !in_bf16 || !scale || !shift || !out_bf16 || N <= 0 || H <= 0 || W <= 0 || Cc <= 0 || Cc % 8 != 0 || ((H | W) & 1))
        return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    Carver cv;
    const size_t in_b = (size_t)N * H * W * Cc * 2, out_b = (size_t)N * (H / 2) * (W / 2) * Cc * 2;
    size_t o_in = cv.take(in_b), o_out = cv.take(out_b), o_par = cv.take((size_t)2 * Cc * 4);
    CUDA_OK(c->ws_io.ensure(cv.off));
    char *b = (char *)c->ws_io.p;
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemcpyAsync(b + o_in, in_bf16, in_b, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_par, scale, (size_t)Cc * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(b + o_par + (size_t)Cc * 4, shift, (size_t)Cc * 4, cudaMemcpyHostToDevice, s));
    prof_reset(c);
    LAUNCH(c, "bn_relu_maxpool", launch_bn_relu_maxpool(b + o_in, b + o_out, N, H, W, Cc, (const float *)(b + o_par), (const float *)(b + o_par) + Cc, 1, s));
    CUDA_OK(cudaMemcpyAsync(out_bf16, b + o_out, out_b, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    prof_collect(c);
    return BUSCA_OK;
}

// hardware probe: D = A(shifted by `shift_rows` rows inside a swizzled tile) * I, see umma_rowshift_probe_kernel (conv_tc.cu)
extern "C" int busca_debug_umma_rowshift(busca_ctx *c, int32_t shift_rows, int32_t fill, int32_t use_base_offset, float *out) {
    if (!c || !out) return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(c->ws_small.ensure(128 * 64 * 4));
    float *d = (float *)c->ws_small.p;
    CUDA_OK(cudaMemsetAsync(d, 0xff, 128 * 64 * 4, c->stream));
    LAUNCH(c, "umma_rowshift_probe", launch_umma_rowshift_probe(shift_rows, fill, use_base_offset, d, c->stream));
    CUDA_OK(cudaMemcpyAsync(out, d, 128 * 64 * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}

// hardware probe: Gram matrix A^T A of nb pixel tiles through MN-major operand descriptors (umma_gram_probe_kernel, conv_tc.cu)
extern "C" int busca_debug_gram(busca_ctx *c, const uint16_t *a_bf16, int32_t nb, float *out) {
    if (!c || !a_bf16 || !out || (nb != 1 && nb != 2 && nb != 4)) return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t C = 64 * (size_t)nb, in_b = 128 * C * 2, out_b = C * C * 4;
    CUDA_OK(c->ws_small.ensure(in_b + out_b + 512));
    uint16_t *da = (uint16_t *)c->ws_small.p;
    float *dout = (float *)((char *)c->ws_small.p + ((in_b + 255) & ~(size_t)255));
    CUDA_OK(cudaMemcpyAsync(da, a_bf16, in_b, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(dout, 0xff, out_b, c->stream));
    LAUNCH(c, "umma_gram_probe", launch_umma_gram_probe(da, nb, dout, c->stream));
    CUDA_OK(cudaMemcpyAsync(out, dout, out_b, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}

extern "C" int busca_debug_stem(busca_ctx *c, const int32_t *slots, int32_t N, int32_t use_tc, uint16_t *out_bf16, double *stats_out) {
    if (!c || !c->finalized || !slots || N <= 0 || !out_bf16) return set_err(BUSCA_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    ConvLayer &L = c->convs[0];
    const size_t out_b = (size_t)N * 192 * 64 * 64 * 2, scr = stem_tc_scratch_bytes(N);
    CUDA_OK(c->ws_reid.ensure(out_b + scr + (size_t)N * 4 + 1024));
    char *dout = (char *)c->ws_reid.p, *dscr = dout + ((out_b + 255) & ~(size_t)255);
    int32_t *dsl = (int32_t *)(dscr + ((scr + 255) & ~(size_t)255));
    cudaStream_t s = c->stream;
    CUDA_OK(cudaMemcpyAsync(dsl, slots, (size_t)N * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(L.stats, 0, 2 * 64 * sizeof(double), s));
    prof_reset(c);
    if (use_tc) LAUNCH(c, "stem_tc", launch_stem_tc(c->bank, dsl, N, c->lut, L.w16, dscr, dout, L.stats, nullptr, s));
    else LAUNCH(c, "stem_simt", launch_stem(c->bank, dsl, N, c->lut, L, dout, 1, s));
    CUDA_OK(cudaMemcpyAsync(out_bf16, dout, out_b, cudaMemcpyDeviceToHost, s));
    if (stats_out) CUDA_OK(cudaMemcpyAsync(stats_out, L.stats, 2 * 64 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" int busca_conv_info(busca_ctx *c, int32_t conv_index, int32_t *out4) {
    if (!c || !c->finalized || conv_index < 0 || conv_index >= (int)c->convs.size()) return set_err(BUSCA_ERR_ARG, "bad argument");
    ConvLayer &L = c->convs[conv_index];
    out4[0] = L.cin; out4[1] = L.cout; out4[2] = L.k; out4[3] = L.stride;
    return BUSCA_OK;
}

// ------------------------------------------------------------------------------------------------
// plumbing
// ------------------------------------------------------------------------------------------------
extern "C" void *busca_dev_alloc(busca_ctx *c, int64_t bytes) {
    if (!c || bytes <= 0) return nullptr;
    cudaSetDevice(c->cfg.device);
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { set_err(BUSCA_ERR_NOMEM, "cudaMalloc(%lld) failed", (long long)bytes); return nullptr; }
    return p;
}
extern "C" void busca_dev_free(busca_ctx *c, void *p) { if (c && p) { cudaSetDevice(c->cfg.device); cudaFree(p); } }
extern "C" void *busca_host_alloc(busca_ctx *c, int64_t bytes) {
    if (!c || bytes <= 0) return nullptr;
    cudaSetDevice(c->cfg.device);
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { set_err(BUSCA_ERR_NOMEM, "cudaHostAlloc(%lld) failed", (long long)bytes); return nullptr; }
    return p;
}
extern "C" void busca_host_free(busca_ctx *c, void *p) {      // ctx may be NULL (a block outliving its context)
    if (!p) return;
    if (c) cudaSetDevice(c->cfg.device);
    cudaFreeHost(p);
}
extern "C" int busca_memcpy_h2d(busca_ctx *c, void *dst, const void *src, int64_t bytes) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}
extern "C" int busca_memcpy_d2h(busca_ctx *c, void *dst, const void *src, int64_t bytes) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return BUSCA_OK;
}
extern "C" int busca_sync(busca_ctx *c) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    CUDA_OK(cudaSetDevice(c->cfg.device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    return BUSCA_OK;
}
extern "C" void *busca_stream(busca_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t busca_kernel_launches(busca_ctx *c) { return c ? c->launches : 0; }
extern "C" int busca_set_profiling(busca_ctx *c, int32_t on) {
    if (!c) return set_err(BUSCA_ERR_ARG, "null ctx");
    c->profiling = on != 0;
    prof_reset(c);
    return BUSCA_OK;
}
extern "C" const char *busca_last_profile(busca_ctx *c) { return c ? c->prof_json.c_str() : "{}"; }
extern "C" int busca_set_option(busca_ctx *c, const char *name, int64_t value) {
    if (!c || !name) return set_err(BUSCA_ERR_ARG, "null argument");
    if (strcmp(name, "dedup") == 0) { c->dedup = value != 0; return BUSCA_OK; }
    if (strcmp(name, "gram") == 0) { c->gram = value != 0; return BUSCA_OK; }
    if (strcmp(name, "tr_tc") == 0) { c->tr_tc = value != 0; return BUSCA_OK; }
    if (strcmp(name, "defer_crop_copies") == 0) { c->defer_crop_copies = value != 0; return BUSCA_OK; }
    if (strcmp(name, "pool_mono") == 0) { reid_set_pool_mono(value); return BUSCA_OK; } // process-wide (experimental max-pool kernel)
    if (strcmp(name, "halo") == 0) { conv_tc_set_halo(value); return BUSCA_OK; }     // process-wide (experimental 3x3 kernel)
    if (strcmp(name, "mc_min_tiles") == 0) { conv_tc_set_mc_min_tiles((int)value); return BUSCA_OK; }   // process-wide
    if (strcmp(name, "cg2_min_tiles") == 0) { conv_tc_set_cg2_min_tiles((int)value); return BUSCA_OK; } // process-wide (cta_group::2 convolutions)
    if (strcmp(name, "pdl") == 0) { pdl_set(value != 0); return BUSCA_OK; }          // process-wide (programmatic dependent launch)
    return set_err(BUSCA_ERR_ARG, "unknown option '%s'", name);
}
extern "C" int64_t busca_counter(busca_ctx *c, const char *name) {
    if (!c || !name) return -1;
    if (strcmp(name, "reid_images_run") == 0) return c->reid_images_run;
    if (strcmp(name, "reid_images_total") == 0) return c->reid_images_total;
    if (strcmp(name, "kernel_launches") == 0) return c->launches;
    return -1;
}
