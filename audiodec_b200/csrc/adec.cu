// audiodec_b200 host side: model plans, weight ingest/packing, the C ABI (include/audiodec_b200.h).
//
// The reference builds torch modules from config.yml and runs ~60 nn.Conv1d calls per encode/decode,
// each preceded by a torch.cat of its pad_buffer (layers/conv_layer.py:153-156).  Here a model is a
// flat list of `Op`s over three ping-pong activation buffers in HBM (channels-last), each Op one
// kernel launch that reads its causal history from a per-stream state buffer and writes the next one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../include/audiodec_b200.h"
#include "kernels.cuh"
#include "tc_kernels.cuh"
#include "tc_persist.cuh"
#include "tc_f16.cuh"
#include "probe.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

using namespace adec;

namespace {
#ifdef ADEC_TIMELINE
unsigned int* g_tl_last = nullptr;
#endif

thread_local std::string g_create_error;

std::string fmt(const char* f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

int round_up(int v, int m) { return (v + m - 1) / m * m; }
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// ------------------------------------------------------------------------------------------------
// kernel dispatch table
// ------------------------------------------------------------------------------------------------
typedef cudaError_t (*ConvLaunchFn)(const ConvArgs&, dim3, int, cudaStream_t);

template <int CW, int CO, int TT, int KC, bool F>
cudaError_t launch_conv(const ConvArgs& a, dim3 grid, int window_rows, cudaStream_t s) {
    using Cfg = ConvCfg<CW, CO, TT, KC>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kern = conv_gemm_kernel<CW, CO, TT, KC, F>;
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const size_t smem = Cfg::smem_bytes(window_rows);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    kern<<<grid, Cfg::NTHREADS, smem, s>>>(a);
    return cudaGetLastError();
}

struct ConvKernelCfg { int CW, CO, TT, KC; bool fuse; ConvLaunchFn fn; };

#define ADEC_PLAIN(CW) \
    {CW, 256, 64, 8, false, launch_conv<CW, 256, 64, 8, false>}, \
    {CW, 128, 64, 16, false, launch_conv<CW, 128, 64, 16, false>}, \
    {CW, 64, 128, 32, false, launch_conv<CW, 64, 128, 32, false>}, \
    {CW, 32, 256, 32, false, launch_conv<CW, 32, 256, 32, false>}

const ConvKernelCfg kConvKernels[] = {
    ADEC_PLAIN(32), ADEC_PLAIN(64), ADEC_PLAIN(96), ADEC_PLAIN(128), ADEC_PLAIN(256),
    {32, 32, 256, 32, true, launch_conv<32, 32, 256, 32, true>},
    {64, 64, 128, 32, true, launch_conv<64, 64, 128, 32, true>},
    {128, 128, 64, 16, true, launch_conv<128, 128, 64, 16, true>},
    {256, 256, 64, 8, true, launch_conv<256, 256, 64, 8, true>},
};

const ConvKernelCfg* find_conv_kernel(int CW, int CO, bool fuse) {
    for (const auto& k : kConvKernels)
        if (k.CW == CW && k.CO == CO && k.fuse == fuse) return &k;
    return nullptr;
}

// tensor-core (tcgen05, 3xTF32; ADEC_CONV_PATH=tf32) instantiations: NT = output-channel tile (UMMA N)
inline int TcpCfgStages(int NT) { return NT == 128 ? TcpCfg<128>::STAGES : NT == 64 ? TcpCfg<64>::STAGES : TcpCfg<32>::STAGES; }
// persistent variant: one CTA per SM loops over (time tile, channel tile, stream) tiles
typedef cudaError_t (*TcPersistFn)(const ConvArgs&, int, int, int, int, int, cudaStream_t);
template <int NT, bool F, int PRE>
cudaError_t launch_tcp(const ConvArgs& a, int n_xtiles, int n_ytiles, int n_tiles, int n_ctas, int smem_bytes, cudaStream_t s) {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kern = tc_conv_persist_kernel<NT, F, PRE>;
#ifdef ADEC_TIMELINE
    constexpr int kMaxDyn = 227 * 1024 - 2048;
#else
    constexpr int kMaxDyn = 227 * 1024;
#endif
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    if (smem_bytes > kMaxDyn) return cudaErrorInvalidConfiguration;
    kern<<<n_ctas, TcpCfg<NT>::THREADS, smem_bytes, s>>>(a, n_xtiles, n_ytiles, n_tiles);
    return cudaGetLastError();
}

struct TcKernelCfg { int NT, KS, stages; bool fuse; int pre; TcPersistFn pfn; };
#define ADEC_TC1(NT, F, PRE) {NT, TC_CP, TcCfg<NT>::STAGES, F, PRE, launch_tcp<NT, F, PRE>}
#define ADEC_TC(NT) \
    ADEC_TC1(NT, true, ACT_ELU), ADEC_TC1(NT, false, ACT_NONE), ADEC_TC1(NT, false, ACT_ELU), ADEC_TC1(NT, false, ACT_LRELU), \
    ADEC_TC1(NT, false, ACT_NORM)
const TcKernelCfg kTcKernels[] = {ADEC_TC(128), ADEC_TC(64), ADEC_TC(32)};
int kTcMaxFuse = 128;    // residual units wider than this run as two launches on the tensor-core path (ADEC_TC_MAXFUSE)

const TcKernelCfg* find_tc_kernel(int NT, bool fuse, int pre) {
    for (const auto& k : kTcKernels)
        if (k.NT == NT && k.fuse == fuse && k.pre == pre) return &k;
    return nullptr;
}

// tcgen05 kind::f16 engine (tc_f16.cuh; default): PREC 3 = two fp16 pieces per operand, three products (fp32-grade);
// PREC 1 = bf16 operands, one product (the vocoder's bf16 mode)
template <int NT, bool F, int PRE, int PREC>
cudaError_t launch_tcf(const ConvArgs& a, int n_xtiles, int n_ytiles, int n_tiles, int n_ctas, int smem_bytes, cudaStream_t s) {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kern = tc_conv_f16_kernel<NT, F, PRE, PREC>;
    constexpr int kMaxDyn = 227 * 1024;
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    if (smem_bytes > kMaxDyn) return cudaErrorInvalidConfiguration;
    // Programmatic dependent launch: the next launch's CTAs may start (barrier init, TMEM allocation, first weight stages - weights are
    // constants) while the tail of the previous kernel of the stream is still running; its producer and drain warps execute
    // griddepcontrol.wait before they touch any activation or state (tc_f16.cuh), which blocks until the previous grid has completed.
    static const bool pdl = [] { const char* e = getenv("ADEC_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3((unsigned)TcfCfg<NT, PREC>::threads(F));
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, n_xtiles, n_ytiles, n_tiles);
}
typedef size_t (*TcfSmemFn)(int, bool);
typedef int (*TcfWbufFn)(int, bool);
struct TcfKernelCfg { int NT; bool fuse; int pre, prec, tap_bytes; TcPersistFn pfn; TcfSmemFn smem; TcfWbufFn n_wbuf; };
#define ADEC_TCF1(NT, F, PRE, PREC) \
    {NT, F, PRE, PREC, TcfCfg<NT, PREC>::TAP_BYTES, launch_tcf<NT, F, PRE, PREC>, TcfCfg<NT, PREC>::smem_bytes, TcfCfg<NT, PREC>::n_wbuf}
#define ADEC_TCF(NT) \
    ADEC_TCF1(NT, true, ACT_ELU, 3), ADEC_TCF1(NT, false, ACT_NONE, 3), ADEC_TCF1(NT, false, ACT_ELU, 3), ADEC_TCF1(NT, false, ACT_LRELU, 3), \
    ADEC_TCF1(NT, false, ACT_NORM, 3), ADEC_TCF1(NT, false, ACT_NONE, 1), ADEC_TCF1(NT, false, ACT_LRELU, 1), ADEC_TCF1(NT, false, ACT_NORM, 1)
const TcfKernelCfg kTcfKernels[] = {ADEC_TCF(128), ADEC_TCF(64), ADEC_TCF(32)};

const TcfKernelCfg* find_tcf_kernel(int NT, bool fuse, int pre, int prec) {
    for (const auto& k : kTcfKernels)
        if (k.NT == NT && k.fuse == fuse && k.pre == pre && k.prec == prec) return &k;
    return nullptr;
}

uint16_t half_bits(float x) { const __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
float half_value(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
uint16_t bf16_bits(float x) { const __nv_bfloat16 h = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }

float tf32_round_host(float x) {   // cvt.rna.tf32.f32: round to nearest (ties away), 10-bit mantissa
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// ------------------------------------------------------------------------------------------------
// Op: one kernel launch of the plan
// ------------------------------------------------------------------------------------------------
enum OpKind { OP_STEM, OP_CONV, OP_HEAD };
enum { BUF_EXT_IN = -10, BUF_EXT_OUT = -11, BUF_NONE = -1 };

struct Op {
    OpKind kind = OP_CONV;
    std::string name;
    // logical description (per group, after channel padding)
    int G = 1, Cin = 0, Cin_eff = 0, Cout = 0, Ktaps = 1, dil = 1, RG = 1;
    int P = 0;        // history rows (of x~)
    int down = 1;     // Tout = (T-1)/down + 1
    int up = 1;       // rows after reinterpretation = Tout*up (transposed conv)
    bool fuse = false, shared_in = false, out_nct = false, post_tanh = false;
    int pre_act = ACT_NONE, mid_act = ACT_NONE;
    float slope = 0.f;
    // host weights until finalize
    std::vector<float> weff;    // [G][Ktaps][Cin_eff][Cout]
    std::vector<float> weff2;   // fuse: [Cout][Cout] as [ci][co]
    std::vector<float> hbias;   // [G*Cout] or empty
    std::vector<float> hstate;  // initial state (P, st_C) or empty (zeros)
    // kernel config
    const ConvKernelCfg* kc = nullptr;
    const TcKernelCfg* tc = nullptr;   // 3xTF32 tensor-core path; kc = FFMA path
    const TcfKernelCfg* tcf = nullptr; // kind::f16 tensor-core path (default)
    float w_scale = 1.f, w2_scale = 1.f;
    int n_pieces = 1, n_co_tiles = 1;
    // device
    float *w = nullptr, *w2 = nullptr, *bias = nullptr;
    const float *mean = nullptr, *scale = nullptr;
    float head_bias = 0.f;
    long long w_tile_floats = 0;
    int st_C = 0, st_groups = 1;
    float* st[2] = {nullptr, nullptr};
    int cur = 0;
    // wiring
    int in_buf = BUF_NONE, out_buf = BUF_NONE, res_buf = BUF_NONE;
    int ldx = 0, x_goff = 0, ldy = 0, y_goff = 0, ldr = 0, r_goff = 0;
};

struct DevBuf {
    float* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct adec_handle {
    adec_config cfg{};
    int device = 0;
    bool finalized = false;
    std::string err;
    std::map<std::string, HostTensor> tensors;
    std::set<std::string> consumed;
    std::vector<Op> enc_ops, dec_ops;
    int n_streams = 1;
    int st_cap = 1;        // streams the state buffers were allocated for
    int engine = 2;               // ADEC_CONV_PATH: 2 = f16 (tcgen05 kind::f16, default), 1 = tf32 (round-1 3xTF32), 0 = ffma (CUDA cores)
    bool use_tc = true;           // any tensor-core engine
    bool bf16 = false;            // cfg.compute_dtype == 1: bf16 operands (HiFi-GAN vocoder, f16 engine only)
    int plain_teams = 2;          // ADEC_PLAIN_TEAMS=1: one producer team on the un-fused launches (A/B)
    int gspan = 0;                // ADEC_GSPAN: accumulation span of the f16 engine (ConvArgs::gspan)
    int dbg_flags = 0;
    int dbg_wdiv = 0;             // ADEC_DBG_WDIV (timing experiments, wrong results)
    bool stack_rows = true;       // ADEC_STACK_ROWS=0: never stack several streams' rows into one tile (A/B)
    unsigned long long* d_ktrace = nullptr;   // ADEC_KTRACE=1: per-launch {start ns, end ns, SM cycles} records (diagnostics)
    int ktrace_n = 0;
    int n_sms = 148;
    DevBuf ws[3];
    std::vector<void*> owned;     // device allocations freed in destroy
    // rvq
    float *d_embed = nullptr, *d_e2 = nullptr, *d_codebook = nullptr;
    float *d_mean = nullptr, *d_scale = nullptr;
    int* d_err = nullptr;
    int64_t launches = 0;
    bool profiling = false;       // per-op CUDA-event timing (adec_profile)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<std::string> prof_names;
    std::vector<double> prof_bytes;   // algorithmic bytes of each recorded launch (SURVEY 8(d) per-layer model)
    // host-path scratch
    DevBuf hx, hz, hzq, hy;
    long long* hidx = nullptr;
    size_t hidx_cap = 0;

    int fail(const std::string& m) { err = m; return 1; }
};

namespace {

#define CK(h, call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return (h)->fail(fmt("%s failed: %s", #call, cudaGetErrorString(e_))); \
    } while (0)

int dev_alloc(adec_handle* h, float** p, size_t n_floats) {
    CK(h, cudaMalloc((void**)p, std::max<size_t>(n_floats, 4) * sizeof(float)));
    h->owned.push_back(*p);
    return 0;
}

int dev_upload(adec_handle* h, float** p, const std::vector<float>& v) {
    if (dev_alloc(h, p, v.size())) return 1;
    if (!v.empty()) CK(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

int ensure(adec_handle* h, DevBuf& b, size_t n_floats) {
    if (b.cap >= n_floats) return 0;
    if (b.p) CK(h, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CK(h, cudaMalloc((void**)&b.p, n_floats * sizeof(float)));
    b.cap = n_floats;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight access (state-dict keys), weight-norm folding
// ------------------------------------------------------------------------------------------------
const HostTensor* find(adec_handle* h, const std::string& key) {
    auto it = h->tensors.find(key);
    if (it == h->tensors.end()) return nullptr;
    h->consumed.insert(key);
    return &it->second;
}

// returns the effective weight for "<prefix>.weight" or folds "<prefix>.weight_g/_v"
// (torch.nn.utils.weight_norm, dim=0: w = v * (g / ||v||), norm over all dims but 0; HiFiGAN.py:193-203)
int get_weight(adec_handle* h, const std::string& prefix, HostTensor* out) {
    if (const HostTensor* w = find(h, prefix + ".weight")) { *out = *w; return 0; }
    const HostTensor* g = find(h, prefix + ".weight_g");
    const HostTensor* v = find(h, prefix + ".weight_v");
    if (!g || !v) return h->fail("missing key " + prefix + ".weight (or .weight_g/.weight_v)");
    const int64_t n0 = v->shape[0], inner = v->numel() / n0;
    if (g->numel() != n0) return h->fail("bad weight_g shape for " + prefix);
    *out = *v;
    for (int64_t i = 0; i < n0; ++i) {
        double ss = 0;
        for (int64_t j = 0; j < inner; ++j) ss += (double)v->data[i * inner + j] * v->data[i * inner + j];
        const float scale = g->data[i] / (float)std::sqrt(ss);
        for (int64_t j = 0; j < inner; ++j) out->data[i * inner + j] = v->data[i * inner + j] * scale;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Op builders
// ------------------------------------------------------------------------------------------------
// causal conv, weight (Cout_total, Cin_g, K)  (layers/conv_layer.py:118-156)
int make_conv_op(adec_handle* h, Op* op, const std::string& name, const HostTensor& W, const HostTensor* bias,
                 int stride, int dil, int groups, int pre_act, float slope, bool shared_in) {
    if (W.shape.size() != 3) return h->fail("conv weight must be 3-D: " + name);
    const int cout_t = (int)W.shape[0], cin_g = (int)W.shape[1], K = (int)W.shape[2];
    if (cout_t % groups) return h->fail("Cout not divisible by groups: " + name);
    const int cout_g = cout_t / groups;
    op->kind = OP_CONV;
    op->name = name;
    op->G = groups;
    op->pre_act = pre_act;
    op->slope = slope;
    op->shared_in = shared_in;
    op->Cout = round_up(cout_g, 32);
    if (stride == 1) {
        op->RG = 1; op->Ktaps = K; op->dil = dil; op->P = (K - 1) * dil; op->down = 1;
        op->Cin = round_up(cin_g, 32);
        op->Cin_eff = op->Cin;
    } else {
        // k = 2s strided conv (encoder.py:62-68) -> 2-tap conv over s-row groups
        if (K != 2 * stride || dil != 1) return h->fail(fmt("%s: strided conv needs kernel=2*stride, dilation 1", name.c_str()));
        if (!is_pow2(cin_g) || cin_g < 4) return h->fail(fmt("%s: strided conv needs power-of-two Cin>=4", name.c_str()));
        op->RG = stride; op->Ktaps = 2; op->dil = 1; op->P = K - 1; op->down = stride;
        op->Cin = cin_g;
        op->Cin_eff = round_up(stride * cin_g, 32);
    }
    op->weff.assign((size_t)groups * op->Ktaps * op->Cin_eff * op->Cout, 0.f);
    for (int g = 0; g < groups; ++g)
        for (int co = 0; co < cout_g; ++co)
            for (int ci = 0; ci < cin_g; ++ci)
                for (int k = 0; k < K; ++k) {
                    const float v = W.data[((size_t)(g * cout_g + co) * cin_g + ci) * K + k];
                    int tap, q;
                    if (stride == 1) { tap = k; q = ci; }
                    else { tap = k / stride; q = (k % stride) * cin_g + ci; }
                    op->weff[(((size_t)g * op->Ktaps + tap) * op->Cin_eff + q) * op->Cout + co] = v;
                }
    if (bias) {
        op->hbias.assign((size_t)groups * op->Cout, 0.f);
        for (int g = 0; g < groups; ++g)
            for (int co = 0; co < cout_g; ++co) op->hbias[(size_t)g * op->Cout + co] = bias->data[g * cout_g + co];
    }
    op->st_groups = shared_in ? 1 : groups;
    op->st_C = op->st_groups * op->Cin;
    return 0;
}

// causal transposed conv, weight (Cin, Cout, 2s)  (layers/conv_layer.py:162-197)
//   y[j*s+r] = b + W[:,:,r]^T x[j] + W[:,:,s+r]^T x[j-1]      (x[j-1] = tap 0, x[j] = tap 1)
int make_convtr_op(adec_handle* h, Op* op, const std::string& name, const HostTensor& W, const HostTensor* bias,
                   int stride, int pre_act, float slope) {
    if (W.shape.size() != 3 || W.shape[2] != 2 * stride) return h->fail(name + ": transposed conv needs kernel = 2*stride");
    const int cin = (int)W.shape[0], cout = (int)W.shape[1], K = 2 * stride;
    op->kind = OP_CONV;
    op->name = name;
    op->G = 1; op->RG = 1; op->Ktaps = 2; op->dil = 1; op->P = 1; op->down = 1; op->up = stride;
    op->pre_act = pre_act; op->slope = slope;
    op->Cin = round_up(cin, 32);
    op->Cin_eff = op->Cin;
    op->Cout = round_up(stride * cout, 32);
    op->weff.assign((size_t)2 * op->Cin_eff * op->Cout, 0.f);
    for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
            for (int r = 0; r < stride; ++r) {
                op->weff[((size_t)0 * op->Cin_eff + ci) * op->Cout + r * cout + co] = W.data[((size_t)ci * cout + co) * K + stride + r];
                op->weff[((size_t)1 * op->Cin_eff + ci) * op->Cout + r * cout + co] = W.data[((size_t)ci * cout + co) * K + r];
            }
    if (bias) {
        op->hbias.assign(op->Cout, 0.f);
        for (int r = 0; r < stride; ++r)
            for (int co = 0; co < cout; ++co) op->hbias[r * cout + co] = bias->data[co];
    }
    op->st_groups = 1;
    op->st_C = op->Cin;
    return 0;
}

// residual unit (models/autoencoder/modules/residual_unit.py:49-81): x + W2 * act(conv_k7_dil(act(x)))
int make_ru_op(adec_handle* h, Op* op, const std::string& name, const HostTensor& W1, const HostTensor& W2, int dil, int act) {
    if (make_conv_op(h, op, name, W1, nullptr, 1, dil, 1, act, 0.f, false)) return 1;
    const int c = (int)W1.shape[0];
    if (W1.shape[1] != c || W2.shape[0] != c || W2.shape[1] != c || W2.shape[2] != 1 || c != op->Cout)
        return h->fail(name + ": residual unit needs square weights with C % 32 == 0");
    op->fuse = true;
    op->mid_act = act;
    op->weff2.assign((size_t)c * c, 0.f);
    for (int co = 0; co < c; ++co)
        for (int ci = 0; ci < c; ++ci) op->weff2[(size_t)ci * c + co] = W2.data[(size_t)co * c + ci];
    return 0;
}

// Tensor-core path, C > 128: the fused unit would need 2*C > 256 accumulator registers per drain thread, so it runs
// as two launches: k7 dilated conv (ELU on load) -> mid, then 1x1 conv (ELU on load) + skip.
void split_ru_op(const Op& ru, Op* conv, Op* pw) {
    *conv = ru;
    conv->fuse = false;
    conv->weff2.clear();
    conv->mid_act = ACT_NONE;
    *pw = Op();
    pw->kind = OP_CONV;
    pw->name = ru.name + ".conv2";
    pw->G = 1; pw->Cin = ru.Cout; pw->Cin_eff = ru.Cout; pw->Cout = ru.Cout; pw->Ktaps = 1; pw->dil = 1; pw->RG = 1; pw->P = 0;
    pw->pre_act = ru.mid_act; pw->slope = ru.slope;
    pw->weff = ru.weff2;            // [ci][co] == [1 tap][Cin_eff][Cout]
    pw->st_groups = 1; pw->st_C = ru.Cout;
}

int pick_piece_width(const Op& op) {
    static const int cands[] = {256, 128, 96, 64, 32};
    for (int cw : cands) {
        if (op.Cin_eff % cw) continue;
        if (op.RG > 1 && !(cw % op.Cin == 0 || op.Cin % cw == 0)) continue;
        return cw;
    }
    return 0;
}

// choose the kernel instantiation, pack + upload weights, allocate state
int finalize_op_tc(adec_handle* h, Op* op) {
    int NT = op->Cout % 128 == 0 ? 128 : op->Cout % 64 == 0 ? 64 : 32;
    // 96 outputs (transposed conv 64 -> 3*32): one zero-padded 128-wide tile beats three 32-wide tiles that each rebuild the
    // same activation window and are smem-operand bound (persistent kernel only: its epilogue masks per 32-column piece)
    const bool pad_tile = !op->fuse && NT == 32 && op->Cout > 64 && op->Cout < 128;
    if (pad_tile) NT = 128;
    op->tc = find_tc_kernel(NT, op->fuse, op->pre_act);
    if (!op->tc || (op->fuse && (NT != op->Cout || op->mid_act != op->pre_act))) return h->fail(op->name + ": no tensor-core kernel");
    const int KS = op->tc->KS, CP = TC_CP;
    op->n_pieces = op->Cin_eff / CP;
    op->n_co_tiles = pad_tile ? 1 : op->Cout / NT;
    op->w_tile_floats = (long long)op->Ktaps * op->Cin_eff * NT * 2;
    // stage c = ((piece*Ktaps + tap)*(CP/KS) + ks): [hi: (KS/4)][NT][4] | [lo: same]   (UMMA K-major, no swizzle)
    auto pack = [&](const float* weff, int G, int ntiles, int pieces, int taps, int cin_eff, int cout, std::vector<float>* out) {
        out->assign((size_t)G * ntiles * taps * cin_eff * NT * 2, 0.f);
        size_t o = 0;
        for (int g = 0; g < G; ++g)
            for (int nt = 0; nt < ntiles; ++nt)
                for (int pc = 0; pc < pieces; ++pc)
                    for (int tap = 0; tap < taps; ++tap)
                        for (int ks = 0; ks < CP / KS; ++ks) {
                            float* hi = out->data() + o;
                            float* lo = hi + (size_t)KS * NT;
                            for (int c4 = 0; c4 < KS / 4; ++c4)
                                for (int n = 0; n < NT; ++n)
                                    for (int e = 0; e < 4; ++e) {
                                        const int k = pc * CP + ks * KS + c4 * 4 + e;
                                        const float w = nt * NT + n < cout ? weff[(((size_t)g * taps + tap) * cin_eff + k) * cout + nt * NT + n] : 0.f;
                                        const float wh = tf32_round_host(w);
                                        hi[((size_t)c4 * NT + n) * 4 + e] = wh;
                                        lo[((size_t)c4 * NT + n) * 4 + e] = tf32_round_host(w - wh);
                                    }
                            o += (size_t)2 * KS * NT;
                        }
    };
    std::vector<float> packed;
    pack(op->weff.data(), op->G, op->n_co_tiles, op->n_pieces, op->Ktaps, op->Cin_eff, op->Cout, &packed);
    if (dev_upload(h, &op->w, packed)) return 1;
    if (op->fuse) {
        std::vector<float> p2;
        pack(op->weff2.data(), 1, 1, op->Cout / CP, 1, op->Cout, op->Cout, &p2);
        if (dev_upload(h, &op->w2, p2)) return 1;
    }
    if (!op->hbias.empty() && dev_upload(h, &op->bias, op->hbias)) return 1;
    std::vector<float>().swap(op->weff);
    std::vector<float>().swap(op->weff2);
    return 0;
}

// kind::f16 engine: weights as fp16 (hi | lo | hi * 2^-11) of w * 2^p, or one bf16 plane, in UMMA K-major no-swizzle blocks:
//   [group g][co tile][piece][tap][plane][kb = 8-channel block][NT rows][8 x 16 bit]
// one (piece, tap) = TAP_BYTES; the kernel's weight producer copies one or two consecutive taps per stage (tc_f16.cuh).
int finalize_op_f16(adec_handle* h, Op* op) {
    int NT = op->Cout % 128 == 0 ? 128 : op->Cout % 64 == 0 ? 64 : 32;
    const bool pad_tile = !op->fuse && NT == 32 && op->Cout > 64 && op->Cout < 128;     // e.g. transposed conv 64 -> 3*32: one padded 128-wide tile
    if (pad_tile) NT = 128;
    const int prec = (h->bf16 && !op->fuse) ? 1 : 3;
    op->tcf = find_tcf_kernel(NT, op->fuse, op->pre_act, prec);
    if (!op->tcf || (op->fuse && (NT != op->Cout || op->mid_act != op->pre_act))) return h->fail(op->name + ": no tensor-core kernel");
    const int CP = TC_CP, KB = F16_KB, npl = prec == 3 ? 3 : 1;
    op->n_pieces = op->Cin_eff / CP;
    op->n_co_tiles = pad_tile ? 1 : op->Cout / NT;
    const size_t tap_bytes = (size_t)op->tcf->tap_bytes;
    op->w_tile_floats = (long long)((size_t)op->n_pieces * op->Ktaps * tap_bytes / 4);
    auto pow2_scale = [](const std::vector<float>& w, int* p_out) {
        float wmax = 0.f;
        for (float v : w) wmax = std::max(wmax, std::fabs(v));
        int p = 0;
        if (wmax > 0.f && std::isfinite(wmax)) {
            while (std::ldexp(wmax, p) < 4096.f) ++p;
            while (std::ldexp(wmax, p) >= 8192.f) --p;
        }
        *p_out = p;
    };
    auto pack = [&](const float* weff, int G, int ntiles, int pieces, int taps, int cin_eff, int cout, int p2, std::vector<float>* out) {
        std::vector<uint16_t> img((size_t)G * ntiles * pieces * taps * tap_bytes / 2, 0);
        size_t o = 0;     // in 16-bit units
        for (int g = 0; g < G; ++g)
            for (int nt = 0; nt < ntiles; ++nt)
                for (int pc = 0; pc < pieces; ++pc)
                    for (int tap = 0; tap < taps; ++tap) {
                        for (int kb = 0; kb < KB; ++kb)
                            for (int n = 0; n < NT; ++n)
                                for (int e = 0; e < 8; ++e) {
                                    const int k = pc * CP + kb * 8 + e;
                                    const float w = nt * NT + n < cout ? weff[(((size_t)g * taps + tap) * cin_eff + k) * cout + nt * NT + n] : 0.f;
                                    const size_t at = o + ((size_t)kb * NT + n) * 8 + e, plane = (size_t)KB * NT * 8;
                                    if (prec == 3) {
                                        const float ws = std::ldexp(w, p2);
                                        const uint16_t hi = half_bits(ws);
                                        img[at] = hi;
                                        img[at + plane] = half_bits(ws - half_value(hi));
                                        img[at + 2 * plane] = half_bits(half_value(hi) * (1.0f / 2048.0f));
                                    } else {
                                        img[at] = bf16_bits(w);
                                    }
                                }
                        o += (size_t)npl * KB * NT * 8;
                    }
        out->assign((img.size() + 1) / 2, 0.f);
        memcpy(out->data(), img.data(), img.size() * 2);
    };
    int p1 = 0, p2 = 0;
    if (prec == 3) pow2_scale(op->weff, &p1);
    op->w_scale = std::ldexp(1.0f, -p1);
    std::vector<float> packed;
    pack(op->weff.data(), op->G, op->n_co_tiles, op->n_pieces, op->Ktaps, op->Cin_eff, op->Cout, p1, &packed);
    if (dev_upload(h, &op->w, packed)) return 1;
    if (op->fuse) {
        pow2_scale(op->weff2, &p2);
        op->w2_scale = std::ldexp(1.0f, -p2);
        std::vector<float> pk2;
        pack(op->weff2.data(), 1, 1, op->Cout / CP, 1, op->Cout, op->Cout, p2, &pk2);
        if (dev_upload(h, &op->w2, pk2)) return 1;
    }
    if (!op->hbias.empty() && dev_upload(h, &op->bias, op->hbias)) return 1;
    std::vector<float>().swap(op->weff);
    std::vector<float>().swap(op->weff2);
    return 0;
}

int finalize_op(adec_handle* h, Op* op) {
    if (op->kind != OP_CONV) return 0;
    if (h->engine == 2) return finalize_op_f16(h, op);
    if (h->use_tc) return finalize_op_tc(h, op);
    int CW, CO;
    if (op->fuse) {
        CW = CO = op->Cout;
    } else {
        CW = pick_piece_width(*op);
        CO = op->Cout % 256 == 0 ? 256 : op->Cout % 128 == 0 ? 128 : op->Cout % 64 == 0 ? 64 : 32;
    }
    op->kc = CW ? find_conv_kernel(CW, CO, op->fuse) : nullptr;
    if (!op->kc) return h->fail(fmt("%s: no kernel for Cin_eff=%d Cout=%d fuse=%d", op->name.c_str(), op->Cin_eff, op->Cout, (int)op->fuse));
    op->n_pieces = op->Cin_eff / CW;
    op->n_co_tiles = op->Cout / CO;
    const int KC = op->kc->KC, nkc = CW / KC;
    op->w_tile_floats = (long long)op->Ktaps * op->Cin_eff * CO;
    std::vector<float> packed((size_t)op->G * op->n_co_tiles * op->w_tile_floats);
    size_t o = 0;
    for (int g = 0; g < op->G; ++g)
        for (int ct = 0; ct < op->n_co_tiles; ++ct)
            for (int pc = 0; pc < op->n_pieces; ++pc)
                for (int tap = 0; tap < op->Ktaps; ++tap)
                    for (int kcc = 0; kcc < nkc; ++kcc)
                        for (int k = 0; k < KC; ++k) {
                            const int q = pc * CW + kcc * KC + k;
                            const float* src = &op->weff[(((size_t)g * op->Ktaps + tap) * op->Cin_eff + q) * op->Cout + ct * CO];
                            for (int co = 0; co < CO; ++co) packed[o++] = src[co];
                        }
    if (dev_upload(h, &op->w, packed)) return 1;
    if (op->fuse) {
        if (dev_upload(h, &op->w2, op->weff2)) return 1;   // [ci][co] == [chunk][kc][co] for CO == C
    }
    if (!op->hbias.empty() && dev_upload(h, &op->bias, op->hbias)) return 1;
    std::vector<float>().swap(op->weff);
    std::vector<float>().swap(op->weff2);
    return 0;
}

int alloc_state(adec_handle* h, Op* op, int n_streams) {
    const size_t per = (size_t)op->P * op->st_C;
    if (per == 0) return 0;
    for (int i = 0; i < 2; ++i) {
        CK(h, cudaMalloc((void**)&op->st[i], per * n_streams * sizeof(float)));     // owned by the op: freed on resize / destroy
        CK(h, cudaMemset(op->st[i], 0, per * n_streams * sizeof(float)));
    }
    op->cur = 0;
    if (!op->hstate.empty()) {
        for (int s = 0; s < n_streams; ++s)
            CK(h, cudaMemcpy(op->st[0] + s * per, op->hstate.data(), per * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

// pad_buffer (1, C, P) from the state dict -> (P, st_C) channels-last initial state.  The reference stores in pad_buffer the tail of
// what it feeds to conv.inference, i.e. values AFTER the pre-activation / normalisation (residual_unit.py:79 `conv1.inference(
// self.activation(x))`, HiFiGAN.py:276-284,288), and the kernels keep exactly that: state rows are never activated again (only chunk
// rows are, while the window is written), so checkpoint values are copied as they are.
void load_pad_buffer(adec_handle* h, Op* op, const std::string& key, int c_real) {
    const HostTensor* pb = find(h, key);
    if (!pb || pb->shape.size() != 3 || op->P == 0) return;
    const int C = (int)pb->shape[1], P = (int)pb->shape[2];
    if (P != op->P) return;
    bool any = false;
    for (float v : pb->data) any |= (v != 0.f);
    if (!any) return;
    op->hstate.assign((size_t)op->P * op->st_C, 0.f);
    const int groups = op->st_groups, cg = c_real;   // real channels per group
    for (int g = 0; g < groups; ++g)
        for (int c = 0; c < cg && g * cg + c < C; ++c)
            for (int p = 0; p < P; ++p) op->hstate[(size_t)p * op->st_C + g * op->Cin + c] = pb->data[((size_t)(g * cg + c)) * P + p];
}

// ------------------------------------------------------------------------------------------------
// plan execution
// ------------------------------------------------------------------------------------------------
struct RunCtx {
    int B;
    const float* ext_in;
    float* ext_out;
    cudaStream_t stream;
    bool offline = false;   // non-streaming forward: transposed convs replicate their first input row instead of reading state
};

int run_ops(adec_handle* h, std::vector<Op>& ops, const RunCtx& rc, int T_in, int* T_out_final) {
    // pass 1: workspace sizes
    size_t need[3] = {0, 0, 0};
    {
        int T = T_in;
        for (const Op& op : ops) {
            const int Tout = (T - 1) / op.down + 1;
            if (op.out_buf >= 0) need[op.out_buf] = std::max(need[op.out_buf], (size_t)rc.B * Tout * op.ldy);
            T = Tout * op.up;
        }
    }
    for (int i = 0; i < 3; ++i)
        if (need[i] && ensure(h, h->ws[i], need[i])) return 1;
    if (rc.B != h->n_streams) return h->fail(fmt("batch %d != n_streams %d (call adec_set_streams)", rc.B, h->n_streams));

    int T = T_in;
    for (Op& op : ops) {
        const int Tout = (T - 1) / op.down + 1;
        const float* xin = op.in_buf == BUF_EXT_IN ? rc.ext_in : h->ws[op.in_buf].p;
        float* yout = op.out_buf == BUF_EXT_OUT ? rc.ext_out : h->ws[op.out_buf].p;
        const float* st_in = op.st[op.cur];
        float* st_out = op.st[op.cur ^ 1];
        cudaError_t e = cudaSuccess;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (h->profiling) {
            cudaEventCreate(&ev0); cudaEventCreate(&ev1);
            cudaEventRecord(ev0, rc.stream);
        }
        if (op.kind == OP_STEM) {
            StemArgs a{};
            a.x = xin; a.x_bs = T; a.st_in = st_in; a.st_out = st_out; a.T = T;
            a.w = op.w; a.bias = op.bias; a.y = yout; a.y_bs = (long long)T * op.ldy;
            dim3 grid((T + 1023) / 1024, rc.B);
            stem_kernel<32, 7><<<grid, 256, 0, rc.stream>>>(a);
            e = cudaGetLastError();
        } else if (op.kind == OP_HEAD) {
            HeadArgs a{};
            a.x = xin; a.x_bs = (long long)T * op.ldx; a.ldx = op.ldx; a.st_in = st_in; a.st_out = st_out; a.T = T;
            a.w = op.w; a.bias = op.head_bias; a.pre_act = op.pre_act; a.slope = op.slope; a.post_tanh = op.post_tanh;
            a.y = yout; a.y_bs = T;
            dim3 grid((T + 255) / 256, rc.B);
            head_kernel<32, 7><<<grid, 256, 0, rc.stream>>>(a);
            e = cudaGetLastError();
        } else {
            ConvArgs a{};
            a.x = xin; a.x_bs = (long long)T * op.ldx; a.ldx = op.ldx; a.x_goff = op.x_goff;
            a.st_in = st_in; a.st_out = st_out; a.st_ld = op.st_C; a.st_goff = op.shared_in ? 0 : op.Cin; a.st_groups = op.st_groups;
            a.P = op.P; a.T = T; a.Tout = Tout;
            a.Ktaps = op.Ktaps; a.dil = op.dil; a.RG = op.RG; a.lgCin = ilog2(op.Cin); a.Cin = op.Cin;
            a.n_pieces = op.n_pieces; a.pre_act = op.pre_act; a.slope = op.slope; a.mean = op.mean; a.scale = op.scale;
            a.w = op.w; a.w2 = op.w2; a.bias = op.bias; a.n_co_tiles = op.n_co_tiles; a.Cout_g = op.Cout;
            a.w_tile_floats = op.w_tile_floats;
            if (op.fuse) { a.res = xin; a.res_bs = a.x_bs; a.ldr = op.ldx; a.r_goff = 0; }
            else if (op.res_buf >= 0) { a.res = h->ws[op.res_buf].p; a.res_bs = (long long)Tout * op.ldr; a.ldr = op.ldr; a.r_goff = op.r_goff; }
            a.y = yout; a.ldy = op.ldy; a.y_goff = op.y_goff; a.out_nct = op.out_nct;
            a.y_bs = op.out_nct ? (long long)op.G * op.Cout * Tout : (long long)Tout * op.ldy;
            a.mid_act = op.mid_act;
            a.hist_rep = (rc.offline && op.up > 1) ? 1 : 0;
            a.w_scale = op.w_scale; a.w2_scale = op.w2_scale; a.err = h->d_err; a.dbg_wdiv = h->dbg_wdiv; a.dbg_flags = h->dbg_flags;
            a.teams = h->plain_teams;
            // bf16-operand launches (the vocoder's reduced-precision mode) accumulate a whole 32-channel piece per TMEM partial: their groups
            // are 4 MMAs, so the TMEM -> register round trip per group is what they wait for, and a 22-step accumulation chain's truncation
            // (3e-7) is far below bf16 operand rounding (measured at batch 128: 43.9 -> 41.5 ms per step, blocks.3.convs2 2.06 -> 1.72 ms)
            const bool bf16_span = op.tcf && op.tcf->prec == 1 && h->gspan == 0;
            a.gspan = (bf16_span || h->gspan == 1 || (h->gspan == 2 && op.tcf && op.tcf->NT >= 128) || (h->gspan == 3 && op.tcf && op.tcf->NT >= 64)) ? 1 : 0;
#ifdef ADEC_TIMELINE
            // debug build: record CTA 1's event timeline of the op named by ADEC_TIMELINE_OP on its 4th launch, dump it to ADEC_TIMELINE_OUT
            static unsigned int* tl_buf = nullptr;
            static int tl_hits = 0;
            if (const char* tlop = getenv("ADEC_TIMELINE_OP")) {
                if (op.name == tlop && ++tl_hits == 4) {
                    if (!tl_buf) cudaMalloc((void**)&tl_buf, (2 + 2 * 8192 * 6) * sizeof(unsigned int));
                    cudaMemsetAsync(tl_buf, 0, (2 + 2 * 8192 * 6) * sizeof(unsigned int), rc.stream);
                    a.tl = tl_buf;
                    g_tl_last = tl_buf;
                }
            }
#endif
            if (h->d_ktrace && h->ktrace_n < 4096) a.dbg = h->d_ktrace + 3 * (size_t)(h->ktrace_n++);
            if (op.tcf || op.tc) {
                // persistent tensor-core kernels: one CTA per SM loops over (time tile, channel tile, stream) tiles
                const int wrows = TC_TT + (op.Ktaps - 1) * op.dil;
                const int NT = op.tcf ? op.tcf->NT : op.tc->NT;
                dim3 grid((Tout + TC_TT - 1) / TC_TT, op.G * op.n_co_tiles, rc.B);
                a.n_streams = rc.B;
                if (op.tcf && h->stack_rows) {
                    // fill the 128-row tiles across streams when that needs fewer tiles (short chunks: 256 streams x 5..25 rows per layer)
                    const long long L = (long long)Tout + (long long)(op.Ktaps - 1) * op.dil;
                    const long long stacked = (rc.B * L + TC_TT - 1) / TC_TT;
                    // (stacked tiles take the per-row edge path in the producers: only worth it when at least a quarter of the tiles go away)
                    if (stacked * 4 <= (long long)grid.x * rc.B * 3 && rc.B * L < (1ll << 30)) {
                        a.stack_L = (int)L;
                        grid = dim3((unsigned)stacked, grid.y, 1);
                    }
                }
                const long long n_tiles = (long long)grid.x * grid.y * grid.z;
                const int n_ctas = (int)std::min<long long>(n_tiles, h->n_sms);
                size_t psmem;
                if (op.tcf) {
                    psmem = op.tcf->smem(wrows, op.fuse);
                    a.n_wbuf = op.tcf->n_wbuf(wrows, op.fuse);
                    if (a.n_wbuf < 2) return h->fail(fmt("%s: window of %d rows does not fit in shared memory", op.name.c_str(), wrows));
                } else {
                    const int wrp = std::max(wrows, 129) | 1;
                    const int pst = TcpCfgStages(NT), pmb = NT == 128 ? TcpCfg<128>::MB : NT == 64 ? TcpCfg<64>::MB : TcpCfg<32>::MB;
                    psmem = 512 + sizeof(float) * ((size_t)pst * 2 * op.tc->KS * NT + (size_t)4 * TC_CP * wrp + (op.fuse ? (size_t)pmb * 2 * TC_CP * TC_MIDP : 0));
                }
                e = (op.tcf ? op.tcf->pfn : op.tc->pfn)(a, (int)grid.x, (int)grid.y, (int)n_tiles, n_ctas, (int)psmem, rc.stream);
            } else {
                const int TT = op.kc->TT;
                dim3 grid((Tout + TT - 1) / TT, op.G * op.n_co_tiles, rc.B);
                e = op.kc->fn(a, grid, TT + (op.Ktaps - 1) * op.dil, rc.stream);
            }
        }
        if (e != cudaSuccess) return h->fail(fmt("launch of %s failed: %s", op.name.c_str(), cudaGetErrorString(e)));
#ifdef ADEC_TIMELINE
        if (op.kind == OP_CONV && getenv("ADEC_TIMELINE_OP") && op.name == getenv("ADEC_TIMELINE_OP")) {
            static bool dumped = false;
            if (g_tl_last && !dumped) {
                dumped = true;
                cudaStreamSynchronize(rc.stream);
                std::vector<unsigned int> hb(2 + 2 * 8192 * 6);
                cudaMemcpy(hb.data(), g_tl_last, hb.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
                const char* outp = getenv("ADEC_TIMELINE_OUT");
                if (FILE* f = fopen(outp ? outp : "timeline.txt", "w")) {
                    fprintf(f, "# %s\n", op.name.c_str());
                    for (unsigned r = 0; r < 6; ++r)
                        for (unsigned i = 0; i < 8192; ++i) {
                            const unsigned w0 = hb[2 + 2 * (8192 * r + i)], w1 = hb[3 + 2 * (8192 * r + i)];
                            if (!w0 && !w1) break;
                            fprintf(f, "%u %u %u\n", w0 >> 24, w0 & 0xffffffu, w1);
                        }
                    fclose(f);
                }
            }
        }
#endif
        ++h->launches;
        if (h->profiling) {
            cudaEventRecord(ev1, rc.stream);
            h->prof_events.emplace_back(ev0, ev1);
            h->prof_names.push_back(op.name);
            // algorithmic bytes: 4*(Cin*Tin + Cout*Tout [+ Cout*Tout residual]) per stream (fused unit: its two convs + skip)
            const double cin = op.kind == OP_STEM ? 1 : (double)op.G * op.Cin_eff / std::max(1, op.RG) * (op.shared_in ? 1.0 / op.G : 1.0);
            const double cout = op.kind == OP_HEAD ? 1 : (double)op.G * op.Cout;
            double bytes = 4.0 * rc.B * (cin * T + cout * Tout);
            if (op.fuse) bytes += 4.0 * rc.B * (3.0 * cout * Tout);      // mid write+read (1x1 conv in/out) and the skip read
            else if (op.res_buf >= 0) bytes += 4.0 * rc.B * cout * Tout;
            h->prof_bytes.push_back(bytes);
        }
        if (op.P > 0) op.cur ^= 1;
        T = Tout * op.up;
    }
    if (T_out_final) *T_out_final = T;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// model builders
// ------------------------------------------------------------------------------------------------
int need_tensor(adec_handle* h, const std::string& key, const HostTensor** out) {
    *out = find(h, key);
    if (!*out) return h->fail("missing key " + key);
    return 0;
}

int build_stem(adec_handle* h, Op* op, const std::string& prefix, int cout) {
    HostTensor Wt;
    if (get_weight(h, prefix + ".conv", &Wt)) return 1;
    const HostTensor* W = &Wt;
    if (W->shape[0] != 32 || W->shape[1] != 1 || W->shape[2] != 7 || cout != 32)
        return h->fail("stem conv: only input_channels=1, encode_channels=32, kernel 7 is built");
    op->kind = OP_STEM; op->name = prefix; op->P = 6; op->st_C = 1; op->Cin = 1; op->Cout = 32;
    std::vector<float> w(7 * 32);
    for (int co = 0; co < 32; ++co)
        for (int k = 0; k < 7; ++k) w[k * 32 + co] = W->data[co * 7 + k];
    if (dev_upload(h, &op->w, w)) return 1;
    if (const HostTensor* b = find(h, prefix + ".conv.bias")) { if (dev_upload(h, &op->bias, b->data)) return 1; }
    if (const HostTensor* pb = find(h, prefix + ".pad_buffer")) {
        bool any = false; for (float v : pb->data) any |= v != 0.f;
        if (any && pb->numel() == 6) op->hstate = pb->data;
    }
    return 0;
}

int build_head(adec_handle* h, Op* op, const std::string& prefix, int pre_act, float slope, bool tanh_out) {
    HostTensor W;
    if (get_weight(h, prefix + ".conv", &W)) return 1;
    if (W.shape[0] != 1 || W.shape[1] != 32 || W.shape[2] != 7)
        return h->fail("head conv: only 32 -> 1 channels, kernel 7 is built");
    op->kind = OP_HEAD; op->name = prefix; op->P = 6; op->st_C = 32; op->Cin = 32; op->Cout = 1;
    op->pre_act = pre_act; op->slope = slope; op->post_tanh = tanh_out;
    std::vector<float> w(7 * 32);
    for (int ci = 0; ci < 32; ++ci)
        for (int k = 0; k < 7; ++k) w[k * 32 + ci] = W.data[ci * 7 + k];
    if (dev_upload(h, &op->w, w)) return 1;
    if (const HostTensor* b = find(h, prefix + ".conv.bias")) op->head_bias = b->data[0];
    op->st_groups = 1;
    load_pad_buffer(h, op, prefix + ".pad_buffer", 32);
    return 0;
}

// wiring helper: ops read `cur` and write another workspace buffer
struct Wire {
    int cur = BUF_EXT_IN;
    int cur_ld = 0;
    int pick(int avoid1 = -99, int avoid2 = -99) const {
        for (int i = 0; i < 3; ++i)
            if (i != cur && i != avoid1 && i != avoid2) return i;
        return 0;
    }
};

void chain(Op* op, Wire* w, int out_ld) {
    op->in_buf = w->cur; op->ldx = w->cur_ld; op->x_goff = op->shared_in ? 0 : op->Cin;
    op->out_buf = w->pick(); op->ldy = out_ld; op->y_goff = op->Cout;
    w->cur = op->out_buf; w->cur_ld = out_ld;
}

// append a residual unit to `ops`: one fused launch, or (tensor-core path, C > 128) conv + 1x1 launches
int push_ru(adec_handle* h, std::vector<Op>* ops, Wire* w, const std::string& ru, int dil, int ch) {
    HostTensor W1, W2;
    if (get_weight(h, ru + ".conv1.conv", &W1) || get_weight(h, ru + ".conv2", &W2)) return 1;
    Op op;
    if (make_ru_op(h, &op, ru, W1, W2, dil, ACT_ELU)) return 1;
    load_pad_buffer(h, &op, ru + ".conv1.pad_buffer", ch);
    if (h->use_tc && op.Cout > kTcMaxFuse) {
        Op conv, pw;
        split_ru_op(op, &conv, &pw);
        const int x_buf = w->cur;
        chain(&conv, w, ch);
        pw.in_buf = w->cur; pw.ldx = ch; pw.x_goff = pw.Cin;
        pw.res_buf = x_buf; pw.ldr = ch; pw.r_goff = 0;
        pw.out_buf = w->pick(x_buf); pw.ldy = ch; pw.y_goff = pw.Cout;
        w->cur = pw.out_buf; w->cur_ld = ch;
        ops->push_back(std::move(conv));
        ops->push_back(std::move(pw));
    } else {
        chain(&op, w, ch);
        ops->push_back(std::move(op));
    }
    return 0;
}

int build_symad(adec_handle* h) {
    const adec_config& c = h->cfg;
    if (c.input_channels != 1 || c.output_channels != 1) return h->fail("symAD: only mono (input/output_channels=1) is built");
    if (c.code_dim != 64) return h->fail("symAD: only code_dim=64 is built");
    // ---- encoder (models/autoencoder/modules/encoder.py:84-142)
    Wire w; w.cur = BUF_EXT_IN; w.cur_ld = 1;
    {
        Op op;
        if (build_stem(h, &op, "encoder.conv", c.encode_channels)) return 1;
        op.in_buf = BUF_EXT_IN; op.ldx = 1; op.out_buf = 0; op.ldy = 32;
        w.cur = 0; w.cur_ld = 32;
        h->enc_ops.push_back(std::move(op));
    }
    int ch = c.encode_channels;
    static const int dils[3] = {1, 3, 9};   // encoder.py:33
    for (int i = 0; i < c.n_enc; ++i) {
        const std::string pre = fmt("encoder.conv_blocks.%d", i);
        for (int j = 0; j < 3; ++j)
            if (push_ru(h, &h->enc_ops, &w, pre + fmt(".res_units.%d", j), dils[j], ch)) return 1;
        HostTensor W;
        if (get_weight(h, pre + ".conv.conv", &W)) return 1;
        const HostTensor* b = find(h, pre + ".conv.conv.bias");
        Op op;
        if (make_conv_op(h, &op, pre + ".conv", W, b, c.enc_strides[i], 1, 1, ACT_NONE, 0.f, false)) return 1;
        load_pad_buffer(h, &op, pre + ".conv.pad_buffer", ch);
        ch = c.encode_channels * c.enc_ratios[i];
        chain(&op, &w, ch);
        h->enc_ops.push_back(std::move(op));
    }
    {   // projector (projector.py:40,52-54): k=3, no bias; writes z channels-first (B,64,F)
        HostTensor W;
        if (get_weight(h, "projector.project.conv", &W)) return 1;
        Op op;   // symAAD: the encoder's trailing ELU (encoder.py:174-175) is this conv's pre-activation
        if (make_conv_op(h, &op, "projector.project", W, find(h, "projector.project.conv.bias"), 1, 1, 1,
                         c.codec_activate ? ACT_ELU : ACT_NONE, 0.f, false)) return 1;
        load_pad_buffer(h, &op, "projector.project.pad_buffer", ch);
        op.in_buf = w.cur; op.ldx = w.cur_ld; op.x_goff = op.Cin;
        op.out_buf = BUF_EXT_OUT; op.out_nct = true; op.ldy = op.Cout; op.y_goff = op.Cout;
        if (op.Cout != c.code_dim) return h->fail("projector: code_dim must be a multiple of 32");
        h->enc_ops.push_back(std::move(op));
    }
    // ---- decoder (models/autoencoder/modules/decoder.py:84-148)
    Wire d; d.cur = BUF_EXT_IN; d.cur_ld = c.code_dim;
    {
        HostTensor W;
        if (get_weight(h, "decoder.conv1.conv", &W)) return 1;
        Op op;
        if (make_conv_op(h, &op, "decoder.conv1", W, find(h, "decoder.conv1.conv.bias"), 1, 1, 1, ACT_NONE, 0.f, false)) return 1;
        load_pad_buffer(h, &op, "decoder.conv1.pad_buffer", c.code_dim);
        chain(&op, &d, op.Cout);
        h->dec_ops.push_back(std::move(op));
    }
    for (int i = 0; i < c.n_dec; ++i) {
        // symAAD wraps each block as Sequential(ELU, DecoderBlock) -> keys "...conv_blocks.i.1.*" (decoder.py:183-195)
        const std::string pre = fmt(c.codec_activate ? "decoder.conv_blocks.%d.1" : "decoder.conv_blocks.%d", i);
        const int cin = c.decode_channels * c.dec_ratios[i];
        const int cout = i < c.n_dec - 1 ? c.decode_channels * c.dec_ratios[i + 1] : c.decode_channels;
        HostTensor W;
        if (get_weight(h, pre + ".conv.deconv", &W)) return 1;
        Op op;
        if (make_convtr_op(h, &op, pre + ".conv", W, find(h, pre + ".conv.deconv.bias"), c.dec_strides[i],
                           c.codec_activate ? ACT_ELU : ACT_NONE, 0.f)) return 1;
        if (op.Cout != c.dec_strides[i] * cout) return h->fail(pre + ": stride*Cout must be a multiple of 32");
        load_pad_buffer(h, &op, pre + ".conv.pad_buffer", cin);
        chain(&op, &d, op.Cout);
        d.cur_ld = cout;   // (T, s*Cout) is (T*s, Cout)
        h->dec_ops.push_back(std::move(op));
        for (int j = 0; j < 3; ++j)
            if (push_ru(h, &h->dec_ops, &d, pre + fmt(".res_units.%d", j), dils[j], cout)) return 1;
    }
    {
        Op op;
        if (build_head(h, &op, "decoder.conv2", c.codec_activate ? ACT_ELU : ACT_NONE, 0.f, c.codec_activate != 0)) return 1;   // decoder.py:209-211
        op.in_buf = d.cur; op.ldx = d.cur_ld; op.out_buf = BUF_EXT_OUT; op.ldy = 1;
        h->dec_ops.push_back(std::move(op));
    }
    // ---- residual VQ (layers/vq_module.py)
    const int nq = c.codebook_num, D = c.code_dim, N = c.codebook_size;
    if (N != 1024) return h->fail("RVQ: only codebook_size=1024 is built");
    std::vector<float> embed((size_t)nq * D * N), e2((size_t)nq * N), cb((size_t)nq * N * D);
    for (int i = 0; i < nq; ++i) {
        const HostTensor* E;
        if (need_tensor(h, fmt("quantizer.codebook.layers.%d.embed", i), &E)) return 1;
        if (E->shape.size() != 2 || E->shape[0] != D || E->shape[1] != N) return h->fail("bad embed shape");
        std::copy(E->data.begin(), E->data.end(), embed.begin() + (size_t)i * D * N);
        for (int cdx = 0; cdx < N; ++cdx) {
            // embed.pow(2).sum(0) in torch's order: cascade sum over blocks of 16 rows (see oracle/rvq_oracle.c)
            float total = 0.f;
            for (int bk = 0; bk < D; bk += 16) {
                float part = 0.f;
                for (int k = bk; k < bk + 16 && k < D; ++k) {
                    volatile float sq = E->data[(size_t)k * N + cdx] * E->data[(size_t)k * N + cdx];
                    part = part + sq;
                }
                total = bk == 0 ? part : total + part;
            }
            e2[(size_t)i * N + cdx] = total;
            for (int k = 0; k < D; ++k) cb[((size_t)i * N + cdx) * D + k] = E->data[(size_t)k * N + cdx];   // vq_module.py:151-157
        }
        find(h, fmt("quantizer.codebook.layers.%d.cluster_size", i));
        find(h, fmt("quantizer.codebook.layers.%d.embed_avg", i));
    }
    if (dev_upload(h, &h->d_embed, embed) || dev_upload(h, &h->d_e2, e2) || dev_upload(h, &h->d_codebook, cb)) return 1;
    return 0;
}

// AD v0 (MultiReceptiveField, multi_fusion.py:23-79): the mean of one residual block per kernel size.  A causal conv with
// kernel k equals one with kernel K >= k whose first K-k taps are zero, so the three blocks are expressed as ONE grouped
// conv stack with the largest kernel (zero-padded taps) followed by a 1x1 "conv_out" holding [I/3 I/3 I/3] - exactly
// the MultiGroupConv1d graph the kernels already run.
int synthesize_mrf_as_groups(adec_handle* h, int stage, int C, HostTensor W1[], HostTensor B1[], HostTensor W2[], HostTensor B2[],
                             HostTensor* Wout) {
    const adec_config& c = h->cfg;
    const int nb = c.n_resblocks;
    int K = 0;
    for (int b = 0; b < nb; ++b) K = std::max(K, c.resblock_kernel_sizes[b]);
    for (int j = 0; j < c.n_dil; ++j)
        for (int which = 0; which < 2; ++which) {
            HostTensor& W = which ? W2[j] : W1[j];
            HostTensor& Bt = which ? B2[j] : B1[j];
            W.shape = {nb * C, C, K};
            W.data.assign((size_t)nb * C * C * K, 0.f);
            Bt.shape = {nb * C};
            Bt.data.assign((size_t)nb * C, 0.f);
            for (int b = 0; b < nb; ++b) {
                const int kb = c.resblock_kernel_sizes[b];
                const std::string pre = fmt("blocks.%d.blocks.%d.convs%d.%d", stage, b, which + 1, j);
                HostTensor w;
                if (get_weight(h, pre + ".conv", &w)) return 1;
                if (w.shape.size() != 3 || w.shape[0] != C || w.shape[1] != C || w.shape[2] != kb) return h->fail("bad shape: " + pre);
                for (int co = 0; co < C; ++co)
                    for (int ci = 0; ci < C; ++ci)
                        for (int k = 0; k < kb; ++k)
                            W.data[((size_t)(b * C + co) * C + ci) * K + (K - kb) + k] = w.data[((size_t)co * C + ci) * kb + k];
                if (const HostTensor* bb = find(h, pre + ".conv.bias"))
                    for (int co = 0; co < C; ++co) Bt.data[b * C + co] = bb->data[co];
                find(h, pre + ".pad_buffer");   // zeros in every released checkpoint; the longer zero-tap history starts at zero too
            }
        }
    Wout->shape = {C, nb * C, 1};
    Wout->data.assign((size_t)C * nb * C, 0.f);
    for (int co = 0; co < C; ++co)
        for (int b = 0; b < nb; ++b) Wout->data[(size_t)co * nb * C + b * C + co] = 1.0f / nb;
    return 0;
}

int build_hifigan(adec_handle* h) {
    const adec_config& c = h->cfg;
    if (c.out_channels != 1) return h->fail("HiFi-GAN: only out_channels=1 is built");
    const bool mrf = c.n_resblocks > 0;       // AD v0
    if (!mrf && c.groups < 2) return h->fail("HiFi-GAN: groups must be > 1 for the MultiGroupConv1d variant");
    if (mrf && c.groups != 1) return h->fail("HiFi-GAN: MultiReceptiveField needs groups = 1");
    const float slope = c.negative_slope;
    if (c.has_stats) {
        const HostTensor *m, *s;
        if (need_tensor(h, "mean", &m) || need_tensor(h, "scale", &s)) return 1;
        std::vector<float> mp(round_up(c.in_channels, 32), 0.f), sp(round_up(c.in_channels, 32), 1.f);
        std::copy(m->data.begin(), m->data.end(), mp.begin());
        std::copy(s->data.begin(), s->data.end(), sp.begin());
        if (dev_upload(h, &h->d_mean, mp) || dev_upload(h, &h->d_scale, sp)) return 1;
    }
    Wire w; w.cur = BUF_EXT_IN; w.cur_ld = c.in_channels;
    {   // input_conv (HiFiGAN.py:84-89, :282-284); decode_norm folded into its window load (:276-279)
        HostTensor W;
        if (get_weight(h, "input_conv.conv", &W)) return 1;
        Op op;
        if (make_conv_op(h, &op, "input_conv", W, find(h, "input_conv.conv.bias"), 1, 1, 1, c.has_stats ? ACT_NORM : ACT_NONE, 0.f, false)) return 1;
        op.mean = h->d_mean; op.scale = h->d_scale;
        load_pad_buffer(h, &op, "input_conv.pad_buffer", c.in_channels);
        chain(&op, &w, op.Cout);
        h->dec_ops.push_back(std::move(op));
    }
    for (int i = 0; i < c.n_up; ++i) {
        const int cin = c.channels >> i, cout = c.channels >> (i + 1), s = c.upsample_scales[i];
        if (c.upsample_kernel_sizes[i] != 2 * s) return h->fail("HiFi-GAN: upsample kernel must be 2*scale (HiFiGAN.py:95)");
        {
            HostTensor W;
            const std::string pre = fmt("upsamples.%d", i);
            if (get_weight(h, pre + ".deconv", &W)) return 1;
            Op op;
            if (make_convtr_op(h, &op, pre, W, find(h, pre + ".deconv.bias"), s, ACT_LRELU, slope)) return 1;
            if (op.Cout != s * cout) return h->fail(pre + ": scale*Cout must be a multiple of 32");
            load_pad_buffer(h, &op, pre + ".pad_buffer", cin);
            chain(&op, &w, op.Cout);
            w.cur_ld = cout;
            h->dec_ops.push_back(std::move(op));
        }
        // MultiGroupConv1d (multi_fusion.py:82-141): x.repeat folded away (shared_in on the first conv)
        const int G = mrf ? c.n_resblocks : c.groups, C3 = G * cout;
        HostTensor mW1[ADEC_MAX_STAGES], mB1[ADEC_MAX_STAGES], mW2[ADEC_MAX_STAGES], mB2[ADEC_MAX_STAGES], mWout;
        if (mrf && synthesize_mrf_as_groups(h, i, cout, mW1, mB1, mW2, mB2, &mWout)) return 1;
        const int A = w.cur;                 // c (T, cout)
        const int Bb = w.pick();             // xt
        const int Cc = w.pick(Bb);           // x (T, 3C)
        for (int j = 0; j < c.n_dil; ++j) {
            const std::string p1 = fmt("blocks.%d.convs1.%d", i, j), p2 = fmt("blocks.%d.convs2.%d", i, j);
            HostTensor W1, W2;
            const HostTensor *b1 = nullptr, *b2 = nullptr;
            if (mrf) {
                W1 = mW1[j]; W2 = mW2[j]; b1 = &mB1[j]; b2 = &mB2[j];
            } else {
                if (get_weight(h, p1 + ".conv", &W1) || get_weight(h, p2 + ".conv", &W2)) return 1;
                b1 = find(h, p1 + ".conv.bias"); b2 = find(h, p2 + ".conv.bias");
            }
            Op o1, o2;
            if (make_conv_op(h, &o1, p1, W1, b1, 1, c.resblock_dilations[j], G, ACT_LRELU, slope, j == 0)) return 1;
            if (make_conv_op(h, &o2, p2, W2, b2, 1, 1, G, ACT_LRELU, slope, false)) return 1;
            if (o1.Cout != cout || o1.Cin != cout) return h->fail(p1 + ": channels must be a multiple of 32");
            if (!mrf) {
                load_pad_buffer(h, &o1, p1 + ".pad_buffer", cout);
                load_pad_buffer(h, &o2, p2 + ".pad_buffer", cout);
            }
            o1.in_buf = j == 0 ? A : Cc; o1.ldx = j == 0 ? cout : C3; o1.x_goff = j == 0 ? 0 : cout;
            o1.out_buf = Bb; o1.ldy = C3; o1.y_goff = cout;
            o2.in_buf = Bb; o2.ldx = C3; o2.x_goff = cout;
            o2.res_buf = j == 0 ? A : Cc; o2.ldr = j == 0 ? cout : C3; o2.r_goff = j == 0 ? 0 : cout;
            o2.out_buf = Cc; o2.ldy = C3; o2.y_goff = cout;     // j>0: in place over the residual (element-wise safe)
            h->dec_ops.push_back(std::move(o1));
            h->dec_ops.push_back(std::move(o2));
        }
        {
            HostTensor W;
            const std::string po = fmt("blocks.%d.conv_out", i);
            if (mrf) W = mWout;
            else if (get_weight(h, po, &W)) return 1;
            Op op;
            if (make_conv_op(h, &op, po, W, find(h, po + ".bias"), 1, 1, 1, ACT_NONE, 0.f, false)) return 1;
            op.in_buf = Cc; op.ldx = C3; op.x_goff = op.Cin;
            op.out_buf = A; op.ldy = cout; op.y_goff = op.Cout;
            w.cur = A; w.cur_ld = cout;
            h->dec_ops.push_back(std::move(op));
        }
    }
    {   // output: LeakyReLU(0.01) -> conv -> tanh (HiFiGAN.py:116-123, :294-296)
        Op op;
        if (build_head(h, &op, "output_conv", ACT_LRELU, 0.01f, true)) return 1;
        op.in_buf = w.cur; op.ldx = w.cur_ld; op.out_buf = BUF_EXT_OUT; op.ldy = 1;
        h->dec_ops.push_back(std::move(op));
    }
    return 0;
}

int hop_of(const adec_handle* h) {
    int hop = 1;
    if (h->cfg.model_type == ADEC_MODEL_SYMAD) for (int i = 0; i < h->cfg.n_enc; ++i) hop *= h->cfg.enc_strides[i];
    else for (int i = 0; i < h->cfg.n_up; ++i) hop *= h->cfg.upsample_scales[i];
    return hop;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

int adec_create(const adec_config* cfg, int device, adec_handle** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return 1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = fmt("no usable CUDA device %d (%s); audiodec_b200 has no CPU fallback", device,
                             e == cudaSuccess ? "out of range" : cudaGetErrorString(e));
        return 1;
    }
    if (cfg->n_enc > ADEC_MAX_STAGES || cfg->n_dec > ADEC_MAX_STAGES || cfg->n_up > ADEC_MAX_STAGES || cfg->n_dil > ADEC_MAX_STAGES) {
        g_create_error = "too many stages";
        return 1;
    }
    auto* h = new adec_handle();
    h->cfg = *cfg;
    h->device = device;
    if (const char* pth = getenv("ADEC_CONV_PATH")) {
        if (!strcmp(pth, "ffma")) h->engine = 0;
        else if (!strcmp(pth, "tf32")) h->engine = 1;
        else if (!strcmp(pth, "f16") || !strcmp(pth, "tc") || !*pth) h->engine = 2;
        else { g_create_error = std::string("ADEC_CONV_PATH must be f16, tf32 or ffma, not ") + pth; delete h; return 1; }
    }
    h->use_tc = h->engine != 0;
    if (const char* sr = getenv("ADEC_STACK_ROWS")) h->stack_rows = atoi(sr) != 0;
    if (const char* gs = getenv("ADEC_GSPAN")) h->gspan = atoi(gs);
    if (const char* pt = getenv("ADEC_PLAIN_TEAMS")) h->plain_teams = atoi(pt) == 1 ? 1 : 2;
    if (const char* wd = getenv("ADEC_DBG_WDIV")) h->dbg_wdiv = atoi(wd);
    if (const char* df = getenv("ADEC_DBG_FLAGS")) h->dbg_flags = atoi(df);
    if (const char* kt = getenv("ADEC_KTRACE")) {
        if (atoi(kt)) { DeviceGuard dgk(device); cudaMalloc((void**)&h->d_ktrace, 4096 * 3 * sizeof(unsigned long long)); }
    }
    h->bf16 = cfg->compute_dtype == 1;
    if (cfg->compute_dtype != 0 && cfg->compute_dtype != 1) { g_create_error = "compute_dtype must be 0 (fp32) or 1 (bf16)"; delete h; return 1; }
    if (h->bf16 && (cfg->model_type != ADEC_MODEL_HIFIGAN || h->engine != 2)) {
        g_create_error = "compute_dtype = bf16 is built for the HiFi-GAN vocoder on the f16 tensor-core engine only (the encoder, projector and RVQ "
                         "must stay fp32-grade for bit-identical indices)";
        delete h;
        return 1;
    }
    if (const char* mf = getenv("ADEC_TC_MAXFUSE")) kTcMaxFuse = atoi(mf);
    cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device);
    DeviceGuard dg(device);
    if (cudaMalloc((void**)&h->d_err, sizeof(int)) != cudaSuccess || cudaMemset(h->d_err, 0, sizeof(int)) != cudaSuccess) {
        g_create_error = "cudaMalloc failed";
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

void adec_destroy(adec_handle* h) {
    if (!h) return;
    DeviceGuard dg(h->device);
    for (void* p : h->owned) cudaFree(p);
    for (auto* ops : {&h->enc_ops, &h->dec_ops})
        for (Op& op : *ops)
            for (int i = 0; i < 2; ++i) if (op.st[i]) cudaFree(op.st[i]);
    for (auto& b : h->ws) if (b.p) cudaFree(b.p);
    for (DevBuf* b : {&h->hx, &h->hz, &h->hzq, &h->hy}) if (b->p) cudaFree(b->p);
    if (h->hidx) cudaFree(h->hidx);
    if (h->d_err) cudaFree(h->d_err);
    if (h->d_ktrace) cudaFree(h->d_ktrace);
    delete h;
}

const char* adec_last_error(const adec_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int adec_set_tensor(adec_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
    if (!h || !key || !data || ndim < 0 || ndim > 4) return h ? h->fail("bad argument to adec_set_tensor") : 1;
    if (h->finalized) return h->fail("adec_set_tensor after adec_finalize");
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    t.data.assign(data, data + t.numel());
    h->tensors[key] = std::move(t);
    return 0;
}

int adec_finalize(adec_handle* h) {
    if (!h) return 1;
    if (h->finalized) return h->fail("already finalized");
    DeviceGuard dg(h->device);
    int rc = h->cfg.model_type == ADEC_MODEL_SYMAD ? build_symad(h)
           : h->cfg.model_type == ADEC_MODEL_HIFIGAN ? build_hifigan(h)
           : h->fail("unknown model_type");
    if (rc) return rc;
    for (const auto& kv : h->tensors)   // load_state_dict(strict=True): unexpected keys are an error
        if (!h->consumed.count(kv.first)) return h->fail("unexpected key in state dict: " + kv.first);
    for (auto* ops : {&h->enc_ops, &h->dec_ops})
        for (Op& op : *ops) {
            if (finalize_op(h, &op)) return 1;
            if (alloc_state(h, &op, 1)) return 1;
        }
    h->n_streams = 1;
    h->tensors.clear();
    h->finalized = true;
    return 0;
}

int adec_n_streams(const adec_handle* h) { return h ? h->n_streams : 0; }

// Resize the per-stream causal state to n streams.  Streams that exist keep their state; new streams either copy stream 0
// (replicate: the "warm one stream, then fan out" case) or start from zero history (= reset_buffer() on them).  Buffers grow on
// demand and the replaced ones are freed.
static int resize_state(adec_handle* h, int n, bool replicate) {
    const int keep = std::min(n, h->n_streams);
    CK(h, cudaDeviceSynchronize());
    for (auto* ops : {&h->enc_ops, &h->dec_ops})
        for (Op& op : *ops) {
            const size_t per = (size_t)op.P * op.st_C;
            if (!per) continue;
            if (n > h->st_cap) {
                float* nb[2] = {nullptr, nullptr};
                for (int i = 0; i < 2; ++i) CK(h, cudaMalloc((void**)&nb[i], per * n * sizeof(float)));
                CK(h, cudaMemcpy(nb[0], op.st[op.cur], per * keep * sizeof(float), cudaMemcpyDeviceToDevice));
                for (int i = 0; i < 2; ++i) cudaFree(op.st[i]);
                op.st[0] = nb[0]; op.st[1] = nb[1]; op.cur = 0;
            }
            if (n > keep) {
                float* tail = op.st[op.cur] + per * keep;
                const long long tot = (long long)per * (n - keep);
                if (replicate) {
                    replicate_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(tail, op.st[op.cur], (long long)per, n - keep);
                    CK(h, cudaGetLastError());
                } else {
                    CK(h, cudaMemset(tail, 0, tot * sizeof(float)));
                }
            }
        }
    CK(h, cudaDeviceSynchronize());
    h->st_cap = std::max(h->st_cap, n);
    h->n_streams = n;
    return 0;
}

int adec_set_streams(adec_handle* h, int n) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (n == h->n_streams) return 0;
    if (n < 1) return h->fail("n_streams must be >= 1");
    DeviceGuard dg(h->device);
    return resize_state(h, n, /*replicate=*/h->n_streams == 1);
}

// Non-streaming forward (codecTest.py:78-95): any batch size, every causal conv starts from a zero left-pad
// (conv_layer.py:148-151) = zeroed state for exactly B streams.  Discards the handle's streaming state.
static int offline_state(adec_handle* h, int B, cudaStream_t s) {
    if (B != h->n_streams && resize_state(h, B, false)) return 1;
    return adec_reset(h, (void*)s);
}

int adec_reset(adec_handle* h, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    DeviceGuard dg(h->device);
    for (auto* ops : {&h->enc_ops, &h->dec_ops})
        for (Op& op : *ops) {
            const size_t per = (size_t)op.P * op.st_C;
            if (!per) continue;
            for (int i = 0; i < 2; ++i) CK(h, cudaMemsetAsync(op.st[i], 0, per * h->n_streams * sizeof(float), (cudaStream_t)stream));
        }
    return 0;
}

int adec_frames_for(const adec_handle* h, int T) {
    if (!h || h->cfg.model_type != ADEC_MODEL_SYMAD) return -1;
    int t = T;
    for (int i = 0; i < h->cfg.n_enc; ++i) t = (t - 1) / h->cfg.enc_strides[i] + 1;
    return t;
}

int adec_hop_length(const adec_handle* h) { return h ? hop_of(h) : -1; }

int adec_encode(adec_handle* h, const float* x, int B, int T, float* z, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (h->cfg.model_type != ADEC_MODEL_SYMAD) return h->fail("encode: not a symAD handle");
    if (B < 1 || T < 1) return h->fail("encode: empty input");
    DeviceGuard dg(h->device);
    RunCtx rc{B, x, z, (cudaStream_t)stream};
    return run_ops(h, h->enc_ops, rc, T, nullptr);
}

int adec_decode(adec_handle* h, const float* zq, int B, int F, float* y, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (B < 1 || F < 1) return h->fail("decode: empty input");
    DeviceGuard dg(h->device);
    RunCtx rc{B, zq, y, (cudaStream_t)stream};
    return run_ops(h, h->dec_ops, rc, F, nullptr);
}

int adec_encode_offline(adec_handle* h, const float* x, int B, int T, float* z, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (h->cfg.model_type != ADEC_MODEL_SYMAD) return h->fail("encode_offline: not a symAD handle");
    if (B < 1 || T < 1) return h->fail("encode_offline: empty input");
    DeviceGuard dg(h->device);
    if (offline_state(h, B, (cudaStream_t)stream)) return 1;
    RunCtx rc{B, x, z, (cudaStream_t)stream, true};
    return run_ops(h, h->enc_ops, rc, T, nullptr);
}

int adec_decode_offline(adec_handle* h, const float* zq, int B, int F, float* y, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (B < 1 || F < 1) return h->fail("decode_offline: empty input");
    DeviceGuard dg(h->device);
    if (offline_state(h, B, (cudaStream_t)stream)) return 1;
    RunCtx rc{B, zq, y, (cudaStream_t)stream, true};
    return run_ops(h, h->dec_ops, rc, F, nullptr);
}

static int index_bits(int n) { int b = 1; while ((1 << b) < n) ++b; return b; }

int adec_packed_frame_bytes(const adec_handle* h) {
    if (!h || h->cfg.model_type != ADEC_MODEL_SYMAD) return -1;
    return (h->cfg.codebook_num * index_bits(h->cfg.codebook_size) + 7) / 8;
}

// residual VQ launch: pick frames-per-pass FP and passes per block so that the grid is ONE wave with the smallest idle tail
extern "C++" template <int FP>
cudaError_t launch_rvq(const adec_handle* h, RvqArgs a, cudaStream_t s, int* cost) {
    static int per_sm[64] = {0};
    auto kern = rvq_kernel<64, 4, FP>;
    const int dev = h->device;
    const auto smem_of = [&](int n_pass) { const int FR = n_pass * FP; return (size_t)FR * (3 * 64 + 1 + 2 * (RVQ_THREADS / 32) + a.nq) * 4; };
    if (dev < 64 && !per_sm[dev]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, RVQ_THREADS, smem_of(2));
        per_sm[dev] = std::max(1, nb);
    }
    const long long nfr = (long long)a.B * a.F, slots = (long long)h->n_sms * per_sm[dev < 64 ? dev : 0];
    int n_pass = (int)((nfr + slots * FP - 1) / (slots * FP));
    n_pass = std::max(1, std::min(n_pass, 96 / FP));                 // <= 96 frames of residuals (x1, x2) + zq in shared memory (79 KB)
    if (cost) { *cost = n_pass * (FP + 2) * (int)((nfr + slots * n_pass * FP - 1) / (slots * n_pass * FP)); return cudaSuccess; }
    a.n_pass = n_pass;
    const long long FR = (long long)n_pass * FP;
    kern<<<(unsigned)((nfr + FR - 1) / FR), RVQ_THREADS, smem_of(n_pass), s>>>(a);
    return cudaGetLastError();
}

int adec_quantize_ex(adec_handle* h, const float* z, int B, int F, int64_t* idx, uint8_t* packed, float* zq, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (h->cfg.model_type != ADEC_MODEL_SYMAD) return h->fail("quantize: not a symAD handle");
    if (B < 1 || F < 1) return h->fail("quantize: empty input");
    if (!idx && !packed && !zq) return h->fail("quantize: no output requested");
    DeviceGuard dg(h->device);
    RvqArgs a{};
    a.z = z; a.B = B; a.F = F; a.nq = h->cfg.codebook_num; a.embed = h->d_embed; a.e2 = h->d_e2; a.idx = (long long*)idx;
    a.packed = packed; a.zq = zq; a.bits = index_bits(h->cfg.codebook_size); a.bpf = adec_packed_frame_bytes(h);
    int c16 = 0, c12 = 0, c8 = 0;
    launch_rvq<16>(h, a, nullptr, &c16); launch_rvq<12>(h, a, nullptr, &c12); launch_rvq<8>(h, a, nullptr, &c8);
    cudaError_t e = (c12 <= c16 && c12 <= c8) ? launch_rvq<12>(h, a, (cudaStream_t)stream, nullptr)
                  : (c16 <= c8)               ? launch_rvq<16>(h, a, (cudaStream_t)stream, nullptr)
                                              : launch_rvq<8>(h, a, (cudaStream_t)stream, nullptr);
    if (e != cudaSuccess) return h->fail(fmt("launch of rvq_kernel failed: %s", cudaGetErrorString(e)));
    ++h->launches;
    return 0;
}

int adec_quantize(adec_handle* h, const float* z, int B, int F, int64_t* idx, void* stream) {
    return adec_quantize_ex(h, z, B, F, idx, nullptr, nullptr, stream);
}

static int lookup_common(adec_handle* h, const int64_t* idx, const uint8_t* packed, int B, int F, float* zq, void* stream) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (h->cfg.model_type != ADEC_MODEL_SYMAD) return h->fail("lookup: not a symAD handle");
    if (B < 1 || F < 1) return h->fail("lookup: empty input");
    DeviceGuard dg(h->device);
    LookupArgs a{};
    a.idx = (const long long*)idx; a.packed = packed; a.nfr = (long long)B * F; a.nq = h->cfg.codebook_num; a.D = h->cfg.code_dim;
    a.N = h->cfg.codebook_size; a.bits = index_bits(a.N); a.bpf = adec_packed_frame_bytes(h);
    a.codebook = h->d_codebook; a.n_rows = (long long)h->cfg.codebook_num * h->cfg.codebook_size; a.zq = zq; a.err = h->d_err;
    const long long nth = a.nfr * (a.D / 4);
    lookup_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    CK(h, cudaGetLastError());
    ++h->launches;
    return 0;
}

int adec_lookup(adec_handle* h, const int64_t* idx, int B, int F, float* zq, void* stream) {
    return lookup_common(h, idx, nullptr, B, F, zq, stream);
}

int adec_lookup_packed(adec_handle* h, const uint8_t* packed, int B, int F, float* zq, void* stream) {
    return lookup_common(h, nullptr, packed, B, F, zq, stream);
}

int adec_codec_host(adec_handle* enc, adec_handle* dec, const float* x_host, int B, int T, int64_t* idx_host,
                    float* y_host, void* stream) {
    if (!enc || !dec) return 1;
    if (enc->device != dec->device) return enc->fail("codec_host: handles on different devices");
    DeviceGuard dg(enc->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int F = adec_frames_for(enc, T);
    if (F < 1) return enc->fail("codec_host: bad T or not a symAD encoder");
    const int D = enc->cfg.code_dim, nq = enc->cfg.codebook_num, hop = hop_of(dec);
    if (ensure(enc, enc->hx, (size_t)B * T) || ensure(enc, enc->hz, (size_t)B * D * F) || ensure(enc, enc->hzq, (size_t)B * F * D) ||
        ensure(enc, enc->hy, (size_t)B * F * hop))
        return 1;
    const size_t nidx = (size_t)nq * B * F;
    if (enc->hidx_cap < nidx) {
        if (enc->hidx) CK(enc, cudaFree(enc->hidx));
        CK(enc, cudaMalloc((void**)&enc->hidx, nidx * sizeof(long long)));
        enc->hidx_cap = nidx;
    }
    CK(enc, cudaMemcpyAsync(enc->hx.p, x_host, (size_t)B * T * sizeof(float), cudaMemcpyHostToDevice, s));
    if (adec_encode(enc, enc->hx.p, B, T, enc->hz.p, stream)) return 1;
    // quantize + lookup in ONE launch: the RVQ kernel also emits zq (bin/stream.py:224 hand-off without the int64 round trip)
    if (adec_quantize_ex(enc, enc->hz.p, B, F, (int64_t*)enc->hidx, nullptr, enc->hzq.p, stream)) return 1;
    if (adec_decode(dec, enc->hzq.p, B, F, enc->hy.p, stream)) { enc->err = dec->err; return 1; }
    if (idx_host) CK(enc, cudaMemcpyAsync(idx_host, enc->hidx, nidx * sizeof(long long), cudaMemcpyDeviceToHost, s));
    CK(enc, cudaMemcpyAsync(y_host, enc->hy.p, (size_t)B * F * hop * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(enc, cudaStreamSynchronize(s));
    for (adec_handle* hh : {enc, dec}) {
        int herr = 0;
        CK(enc, cudaMemcpy(&herr, hh->d_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) CK(enc, cudaMemset(hh->d_err, 0, sizeof(int)));
        if (hh->dbg_flags || hh->dbg_wdiv) continue;      // timing experiments: results are wrong on purpose
        if (herr & 1) return enc->fail("lookup: index out of range");
        if (herr & 2) return enc->fail("an activation left the range of the fp16-split tensor-core engine (|a| >= 6e4); use ADEC_CONV_PATH=tf32");
        if (hh == dec) break;     // enc == dec
    }
    return 0;
}

static int pack_common(adec_handle* h, const char* what, int64_t* idx, int B, int F, uint8_t* packed, void* stream, bool pack) {
    if (!h || !h->finalized) return h ? h->fail("not finalized") : 1;
    if (h->cfg.model_type != ADEC_MODEL_SYMAD) return h->fail(std::string(what) + ": not a symAD handle");
    if (B < 1 || F < 1) return h->fail(std::string(what) + ": empty input");
    if (h->cfg.codebook_size < 2 || h->cfg.codebook_size > 65536) return h->fail(std::string(what) + ": codebook_size must be in [2, 65536]");
    DeviceGuard dg(h->device);
    PackArgs a{};
    a.idx = (long long*)idx; a.packed = packed; a.nfr = (long long)B * F; a.nq = h->cfg.codebook_num; a.N = h->cfg.codebook_size;
    a.bits = index_bits(a.N); a.bpf = adec_packed_frame_bytes(h); a.err = h->d_err;
    const unsigned grid = (unsigned)((a.nfr + 255) / 256);
    if (pack) pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    CK(h, cudaGetLastError());
    ++h->launches;
    return 0;
}

int adec_pack_indices(adec_handle* h, const int64_t* idx, int B, int F, uint8_t* packed, void* stream) {
    return pack_common(h, "pack_indices", const_cast<int64_t*>(idx), B, F, packed, stream, true);
}

int adec_unpack_indices(adec_handle* h, const uint8_t* packed, int B, int F, int64_t* idx, void* stream) {
    return pack_common(h, "unpack_indices", idx, B, F, const_cast<uint8_t*>(packed), stream, false);
}

// device flag word: bit 0 = out-of-range index (lookup / pack / unpack), bit 1 = activation outside the fp16-split range
static int read_flag(adec_handle* h, void* stream, int bit) {
    if (!h || !h->d_err) return -1;
    DeviceGuard dg(h->device);
    int herr = 0;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess ||
        cudaMemcpy(&herr, h->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        h->fail("CUDA error while reading the device flag word");
        return -1;
    }
    if (herr & bit) {
        const int rest = herr & ~bit;
        if (cudaMemcpy(h->d_err, &rest, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) { h->fail("CUDA error while clearing the device flag word"); return -1; }
    }
    return (herr & bit) ? 1 : 0;
}

int adec_index_error(adec_handle* h, void* stream) { return read_flag(h, stream, 1); }
int adec_range_error(adec_handle* h, void* stream) { return read_flag(h, stream, 2); }

int64_t adec_launch_count(const adec_handle* h) { return h ? h->launches : 0; }

int adec_ktrace(adec_handle* h, unsigned long long* out, int max_records) {
    if (!h || !h->d_ktrace || !out) return -1;
    DeviceGuard dg(h->device);
    const int n = std::min(h->ktrace_n, max_records);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(out, h->d_ktrace, (size_t)n * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    h->ktrace_n = 0;
    return n;
}

int adec_probe_mma_ex(int device, int kind, int NT, int n_groups, int a_off_rows, int a_pitch_rows, int tap_step_rows, int n_issuers, double* tflops, double* ms) {
    if (!tflops || (kind != 0 && kind != 1) || NT < 16 || NT > 256 || NT % 16 || n_groups < 1 || a_off_rows < 0 || a_pitch_rows < 128 || a_pitch_rows > 256 || tap_step_rows < 0 || tap_step_rows > 16 || n_issuers < 1 || n_issuers > 4) { g_create_error = "probe_mma: bad argument"; return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_create_error = "probe_mma: no usable CUDA device"; return 1; }
    DeviceGuard dg(device);
    int n_sms = 0;
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
    const int smem = ((8 * 2 * a_pitch_rows + a_off_rows + 8 * tap_step_rows + 8) * 16 & ~127) + 8 * 2 * NT * 16;
    auto launch = [&](cudaStream_t s) {
        if (kind == 0) { cudaFuncSetAttribute(mma_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); mma_probe_kernel<0><<<n_sms, 128, smem, s>>>(NT, n_groups, a_off_rows, a_pitch_rows, tap_step_rows, n_issuers); }
        else { cudaFuncSetAttribute(mma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); mma_probe_kernel<1><<<n_sms, 128, smem, s>>>(NT, n_groups, a_off_rows, a_pitch_rows, tap_step_rows, n_issuers); }
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(0);                                   // warm-up: clocks, instruction cache
    cudaEventRecord(e0, 0);
    launch(0);
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaEventSynchronize(e1);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { g_create_error = fmt("probe_mma: %s", cudaGetErrorString(e)); return 1; }
    const double flops = (double)n_sms * n_groups * 12.0 * 2.0 * 128.0 * NT * (kind == 0 ? 8.0 : 16.0);
    *tflops = flops / (t * 1e-3) / 1e12;
    if (ms) *ms = t;
    return 0;
}

int adec_probe_mma(int device, int kind, int NT, int n_groups, double* tflops, double* ms) {
    return adec_probe_mma_ex(device, kind, NT, n_groups, 0, 128, 0, 1, tflops, ms);
}

int adec_profile(adec_handle* h, int enable) {
    if (!h) return 1;
    for (auto& ev : h->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    h->prof_events.clear(); h->prof_names.clear(); h->prof_bytes.clear();
    h->profiling = enable != 0;
    return 0;
}

int adec_profile_report(adec_handle* h, char* buf, int buf_len) {
    if (!h || !buf || buf_len < 2) return 1;
    DeviceGuard dg(h->device);
    std::string out;
    for (size_t i = 0; i < h->prof_events.size(); ++i) {
        if (cudaEventSynchronize(h->prof_events[i].second) != cudaSuccess) return h->fail("profile: event sync failed");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second);
        out += fmt("%s\t%.6f\t%.0f\n", h->prof_names[i].c_str(), ms, h->prof_bytes[i]);
    }
    if ((int)out.size() + 1 > buf_len) return h->fail("profile: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

// -------------------------------------------------------------------------------------------------
// single-layer entry points for the unit tests (HOST pointers, reference layouts)
// -------------------------------------------------------------------------------------------------
static int run_single(adec_handle* h, Op& op, const float* x, int B, int Cin_real, int T, int groups, int Cout_real_total,
                      float* state, int P_real, float* y, bool convtr, int stride) {
    // x (B, Cin_total, T) channels-first -> channels-last with per-group channel padding
    const int G = op.shared_in ? 1 : op.G;
    const int cin_g = Cin_real / G, ldx = G * op.Cin;
    std::vector<float> xl((size_t)B * T * ldx, 0.f);
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int c = 0; c < cin_g; ++c)
                for (int t = 0; t < T; ++t) xl[((size_t)b * T + t) * ldx + g * op.Cin + c] = x[((size_t)b * Cin_real + g * cin_g + c) * T + t];
    op.hstate.assign((size_t)B * op.P * op.st_C, 0.f);
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int c = 0; c < cin_g; ++c)
                for (int p = 0; p < P_real; ++p)
                    op.hstate[((size_t)b * op.P + p) * op.st_C + g * op.Cin + c] = state[((size_t)b * Cin_real + g * cin_g + c) * P_real + p];
    op.in_buf = BUF_EXT_IN; op.ldx = ldx; op.x_goff = op.Cin;
    op.out_buf = BUF_EXT_OUT; op.ldy = op.G * op.Cout; op.y_goff = op.Cout;
    if (finalize_op(h, &op)) return 1;
    const size_t per = (size_t)op.P * op.st_C;
    for (int i = 0; i < 2; ++i) if (dev_alloc(h, &op.st[i], per * B)) return 1;
    CK(h, cudaMemcpy(op.st[0], op.hstate.data(), per * B * sizeof(float), cudaMemcpyHostToDevice));
    h->n_streams = B;
    const int Tout = (T - 1) / op.down + 1;
    float *dx, *dy;
    if (dev_alloc(h, &dx, xl.size()) || dev_alloc(h, &dy, (size_t)B * Tout * op.ldy)) return 1;
    CK(h, cudaMemcpy(dx, xl.data(), xl.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<Op> ops;
    ops.push_back(op);
    RunCtx rc{B, dx, dy, 0};
    if (run_ops(h, ops, rc, T, nullptr)) return 1;
    CK(h, cudaDeviceSynchronize());
    std::vector<float> yl((size_t)B * Tout * op.ldy), sl(per * B);
    CK(h, cudaMemcpy(yl.data(), dy, yl.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(sl.data(), ops[0].st[ops[0].cur], sl.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (!convtr) {
        const int cout_g = Cout_real_total / op.G;
        for (int b = 0; b < B; ++b)
            for (int g = 0; g < op.G; ++g)
                for (int c = 0; c < cout_g; ++c)
                    for (int t = 0; t < Tout; ++t)
                        y[((size_t)b * Cout_real_total + g * cout_g + c) * Tout + t] = yl[((size_t)b * Tout + t) * op.ldy + g * op.Cout + c];
    } else {
        for (int b = 0; b < B; ++b)
            for (int c = 0; c < Cout_real_total; ++c)
                for (int j = 0; j < Tout; ++j)
                    for (int r = 0; r < stride; ++r)
                        y[((size_t)b * Cout_real_total + c) * Tout * stride + j * stride + r] = yl[((size_t)b * Tout + j) * op.ldy + r * Cout_real_total + c];
    }
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int c = 0; c < cin_g; ++c)
                for (int p = 0; p < P_real; ++p)
                    state[((size_t)b * Cin_real + g * cin_g + c) * P_real + p] = sl[((size_t)b * op.P + p) * op.st_C + g * op.Cin + c];
    return 0;
}

int adec_test_causal_conv(int device, const float* x, int B, int Cin, int T, const float* w, const float* bias, int Cout,
                          int K, int stride, int dil, int groups, int pre_act, float slope, float* state, float* y) {
    adec_config cfg{};
    adec_handle* h = nullptr;
    if (adec_create(&cfg, device, &h)) return 1;
    DeviceGuard dg(device);
    HostTensor W, Bt;
    W.shape = {Cout, Cin / groups, K};
    W.data.assign(w, w + (size_t)Cout * (Cin / groups) * K);
    if (bias) { Bt.shape = {Cout}; Bt.data.assign(bias, bias + Cout); }
    Op op;
    int rc = make_conv_op(h, &op, "test_conv", W, bias ? &Bt : nullptr, stride, dil, groups, pre_act, slope, false);
    if (!rc) rc = run_single(h, op, x, B, Cin, T, groups, Cout, state, (K - 1) * dil, y, false, 1);
    if (rc) g_create_error = h->err;
    adec_destroy(h);
    return rc;
}

int adec_test_residual_unit(int device, const float* x, int B, int C, int T, const float* w1, const float* w2, int K, int dil,
                            float* state, float* y) {
    adec_config cfg{};
    adec_handle* h = nullptr;
    if (adec_create(&cfg, device, &h)) return 1;
    DeviceGuard dg(device);
    HostTensor W1, W2;
    W1.shape = {C, C, K};
    W1.data.assign(w1, w1 + (size_t)C * C * K);
    W2.shape = {C, C, 1};
    W2.data.assign(w2, w2 + (size_t)C * C);
    Op op;
    int rc = make_ru_op(h, &op, "test_ru", W1, W2, dil, ACT_ELU);
    if (!rc && h->use_tc && op.Cout > kTcMaxFuse) rc = h->fail("test_ru: C > 128 runs as two ops on the tensor-core path; test those separately");
    if (!rc) rc = run_single(h, op, x, B, C, T, 1, C, state, (K - 1) * dil, y, false, 1);
    if (rc) g_create_error = h->err;
    adec_destroy(h);
    return rc;
}

int adec_test_causal_convtr(int device, const float* x, int B, int Cin, int T, const float* w, const float* bias, int Cout,
                            int stride, float* state, float* y) {
    adec_config cfg{};
    adec_handle* h = nullptr;
    if (adec_create(&cfg, device, &h)) return 1;
    DeviceGuard dg(device);
    HostTensor W, Bt;
    W.shape = {Cin, Cout, 2 * stride};
    W.data.assign(w, w + (size_t)Cin * Cout * 2 * stride);
    if (bias) { Bt.shape = {Cout}; Bt.data.assign(bias, bias + Cout); }
    Op op;
    int rc = make_convtr_op(h, &op, "test_convtr", W, bias ? &Bt : nullptr, stride, ACT_NONE, 0.f);
    if (!rc) rc = run_single(h, op, x, B, Cin, T, 1, Cout, state, 1, y, true, stride);
    if (rc) g_create_error = h->err;
    adec_destroy(h);
    return rc;
}

}  // extern "C"
