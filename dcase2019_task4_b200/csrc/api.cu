// C-ABI entry points: context, parameter / workspace layout, CRNN forward / backward orchestration.
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/dcase_b200.h"
#include "cnn.cuh"
#include "common.cuh"
#include "ctx.h"
#include "gru.cuh"
#include "gemm_tc.cuh"
#include "head_loss.cuh"

int dcase_logmel_tables_create(dcase_ctx* ctx);
void dcase_logmel_tables_destroy(dcase_ctx* ctx);

static thread_local char g_err[512] = "";
unsigned long long g_dcase_launches = 0;

namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; int stream_id; };
std::vector<cudaStream_t> g_prof_streams;
bool g_prof_on = false;       // per-kernel events; the orchestration then keeps everything on ONE stream (isolated durations)
bool g_prof_timeline = false; // events with the streams left as they are: a timeline of the overlapped step (dcase_profile_timeline)
std::vector<ProfRec> g_prof;
}  // namespace

DcaseProfScope::DcaseProfScope(const char* name, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on && !g_prof_timeline) return;
    ProfRec r{name, nullptr, nullptr, 0};
    for (size_t i = 0; i <= g_prof_streams.size(); ++i) {
        if (i == g_prof_streams.size()) { g_prof_streams.push_back(s); }
        if (g_prof_streams[i] == s) { r.stream_id = (int)i; break; }
    }
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
DcaseProfScope::~DcaseProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

void dcase_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

constexpr int kH = 64;   // GRU hidden size of cfg.crnn_kwargs (config.py:53-58)
constexpr int kC = 64;

struct ParamEntry { std::string name; long long off; long long n; };

std::vector<ParamEntry> param_table(int NC) {
    std::vector<ParamEntry> t;
    long long off = 0;
    auto add = [&](const std::string& n, long long cnt) { t.push_back({n, off, cnt}); off += cnt; };
    for (int i = 0; i < 3; ++i) {
        const std::string p = "cnn.cnn.";
        const std::string s = std::to_string(i);
        add(p + "conv" + s + ".weight", (i == 0 ? 1 : kC) * kC * 9);
        add(p + "conv" + s + ".bias", kC);
        add(p + "batchnorm" + s + ".weight", kC);
        add(p + "batchnorm" + s + ".bias", kC);
        add(p + "glu" + s + ".linear.weight", kC * kC);
        add(p + "glu" + s + ".linear.bias", kC);
    }
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
            const std::string suf = "_l" + std::to_string(l) + (d ? "_reverse" : "");
            const int nin = l == 0 ? kC : 2 * kH;
            add("rnn.rnn.weight_ih" + suf, 3 * kH * nin);
            add("rnn.rnn.weight_hh" + suf, 3 * kH * kH);
            add("rnn.rnn.bias_ih" + suf, 3 * kH);
            add("rnn.rnn.bias_hh" + suf, 3 * kH);
        }
    add("dense.weight", NC * 2 * kH);
    add("dense.bias", NC);
    add("dense_softmax.weight", NC * 2 * kH);
    add("dense_softmax.bias", NC);
    return t;
}

struct POff {   // element offsets into the flat slab
    long long conv_w[3], conv_b[3], bn_w[3], bn_b[3], glu_w[3], glu_b[3];
    long long w_ih[2][2], w_hh[2][2], b_ih[2][2], b_hh[2][2];
    long long dense_w, dense_b, soft_w, soft_b, total;
};

POff param_offsets(int NC) {
    POff o{};
    const auto t = param_table(NC);
    size_t i = 0;
    for (int l = 0; l < 3; ++l) {
        o.conv_w[l] = t[i++].off; o.conv_b[l] = t[i++].off; o.bn_w[l] = t[i++].off; o.bn_b[l] = t[i++].off;
        o.glu_w[l] = t[i++].off; o.glu_b[l] = t[i++].off;
    }
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
            o.w_ih[l][d] = t[i++].off; o.w_hh[l][d] = t[i++].off; o.b_ih[l][d] = t[i++].off; o.b_hh[l][d] = t[i++].off;
        }
    o.dense_w = t[i++].off; o.dense_b = t[i++].off; o.soft_w = t[i++].off; o.soft_b = t[i++].off;
    o.total = t.back().off + t.back().n;
    return o;
}

struct WsEntry { const char* name; size_t off; size_t n_elems; };

struct WsLayout {
    std::vector<WsEntry> e;
    size_t total = 0;
    size_t acc_off = 0, acc_bytes = 0;   // region zeroed at the start of every backward
    size_t off(const char* name) const {
        for (const auto& x : e) if (!strcmp(x.name, name)) return x.off;
        return (size_t)-1;
    }
};

WsLayout ws_layout(int B, int T, int NC) {
    (void)NC;
    WsLayout L;
    auto add = [&](const char* name, size_t n_elems, size_t elem = sizeof(float)) {
        L.e.push_back({name, L.total, n_elems});
        L.total += (n_elems * elem + 255) / 256 * 256;
    };
    const size_t BT = (size_t)B * (T / 8);
    const size_t n0 = (size_t)B * (T / 2) * 16 * 64;   // out0 / ypre1 elements
    const size_t n1 = (size_t)B * (T / 4) * 4 * 64;    // out1 / ypre2
    add("mom0", 56, sizeof(double));        // 54 tap moments + completion ticket of the moments kernel
    add("tile_ctr", 16);                    // tile counters of the dynamically scheduled block-0 kernels
    add("stats1", 128, sizeof(double));
    add("stats2", 128, sizeof(double));
    add("fold0", kFold0Size);
    add("bn1", kBnSize);
    add("bn2", kBnSize);
    add("gluimg1", kGluImgBytes / 4); add("gluimg2", kGluImgBytes / 4);
    add("wprep1_d", 36864); add("wprep2_d", 36864);                 // tf32 images of the data-gradient pass
    add("wprep1_h", 36864 / 2); add("wprep2_h", 36864 / 2);         // fp16 images of the forward pass
    add("out0", n0); add("ypre1", n0); add("out1", n1); add("ypre2", n1); add("out2", BT * 64);
    add("out0_h", n0 / 2); add("out1_h", n1 / 2);                   // fp16 copies: operands of the forward convs
    add("gi", 2 * BT * 192);
    const char* sv[2][5] = {{"s0_r", "s0_z", "s0_n", "s0_hn", "s0_hp"}, {"s1_r", "s1_z", "s1_n", "s1_hn", "s1_hp"}};
    for (int l = 0; l < 2; ++l) for (int k = 0; k < 5; ++k) add(sv[l][k], 2 * BT * 64);
    add("rnn0", BT * 128); add("rnn1", BT * 128);
    add("den", (size_t)B * 16);
    L.acc_off = L.total;
    add("acc0", kCnn0AccFloats); add("s12_1", 128); add("s12_2", 128);
    L.acc_bytes = L.total - L.acc_off;
    add("d_rnn1", BT * 128); add("d_rnn0", BT * 128);
    add("dgi0", 2 * BT * 192); add("dgh0", 2 * BT * 192); add("dgi1", 2 * BT * 192); add("dgh1", 2 * BT * 192);
    add("d_out2", BT * 64);
    add("dy2", n1); add("d_out1", n1); add("dy1", n0); add("d_out0", n0);
    return L;
}

template <typename T>
T* wsp(void* ws, const WsLayout& L, const char* name) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + L.off(name));
}

int check_shape(int B, int T, int NC) {
    DCASE_REQUIRE(B >= 1, "batch must be >= 1");
    DCASE_REQUIRE(T >= 8 && T % 8 == 0, "frames must be a positive multiple of 8 (pooling_time_ratio)");
    DCASE_REQUIRE(NC >= 1 && NC <= 16, "n_class must be in [1,16]");
    return DCASE_OK;
}

#define DCASE_TRY(expr) do { int rc__ = (expr); if (rc__ != DCASE_OK) return rc__; } while (0)

}  // namespace

extern "C" {

int dcase_version(void) { return DCASE_B200_VERSION; }
size_t dcase_sizeof_mt_args(void) { return sizeof(dcase_mt_args); }
size_t dcase_sizeof_step_scalars(void) {
    static_assert(sizeof(dcase_step_scalars) == sizeof(DcaseStepScalars), "scalar struct mirrors diverged");
    return sizeof(dcase_step_scalars);
}

unsigned long long dcase_launch_count(void) { return g_dcase_launches; }

int dcase_profile_begin(void) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
    return DCASE_OK;
}

int dcase_profile_end(char* buf, size_t cap) {
    g_prof_on = false;
    DCASE_REQUIRE(buf && cap > 0, "null buffer");
    DCASE_CUDA_CHECK(cudaDeviceSynchronize());
    struct Agg { std::string name; int count; double ms; };
    std::vector<Agg> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        bool found = false;
        for (auto& a : agg) if (a.name == r.name) { a.count++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({r.name, 1, ms});
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    std::string out;
    for (auto& a : agg) {
        char line[160];
        snprintf(line, sizeof(line), "%s,%d,%.6f\n", a.name.c_str(), a.count, a.ms);
        out += line;
    }
    if (out.size() + 1 > cap) { dcase_set_error("profile buffer too small"); return DCASE_ERR_ARG; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return DCASE_OK;
}
// Timeline mode: dcase_profile_timeline_begin(), run a step EAGERLY, dcase_profile_timeline_end(buf): one line per launch,
// "name,stream,start_us,end_us" relative to the first launch; start = when the stream reached the launch, end = when the
// kernel finished (CUDA events on the launching stream; the streams overlap as in production).
int dcase_profile_timeline_begin(void) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_timeline = true;
    return DCASE_OK;
}
int dcase_profile_timeline_end(char* buf, size_t cap) {
    g_prof_timeline = false;
    DCASE_REQUIRE(buf && cap > 0, "null buffer");
    DCASE_CUDA_CHECK(cudaDeviceSynchronize());
    std::string out;
    for (size_t i = 0; i < g_prof.size(); ++i) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, g_prof[0].a, g_prof[i].a);
        cudaEventElapsedTime(&b, g_prof[0].a, g_prof[i].b);
        char line[200];
        snprintf(line, sizeof(line), "%s,%d,%.2f,%.2f\n", g_prof[i].name, g_prof[i].stream_id, a * 1e3, b * 1e3);
        out += line;
    }
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    if (out.size() + 1 > cap) { dcase_set_error("timeline buffer too small"); return DCASE_ERR_ARG; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return DCASE_OK;
}
const char* dcase_last_error(void) { return g_err; }

int dcase_ctx_create(dcase_ctx** out, int device) {
    DCASE_REQUIRE(out, "null out pointer");
    int n_dev = 0;
    DCASE_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
    DCASE_REQUIRE(device >= 0 && device < n_dev, "no such CUDA device");
    DCASE_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DCASE_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        dcase_set_error("dcase_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return DCASE_ERR_STATE;
    }
    dcase_ctx* ctx = new dcase_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    DCASE_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        DCASE_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->prep_stream[i], cudaStreamNonBlocking));
        DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_prep_fork[i], cudaEventDisableTiming));
        DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_prep_join[i], cudaEventDisableTiming));
    }
    DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) {
        DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_bwd_fork[i], cudaEventDisableTiming));
        DCASE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_bwd_join[i], cudaEventDisableTiming));
    }
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_loss_scratch, kLossScratchBytes));
    DCASE_CUDA_CHECK(cudaMemset(ctx->d_loss_scratch, 0, kLossScratchBytes));
    int rc = dcase_logmel_tables_create(ctx);
    if (rc == DCASE_OK) rc = cnn_kernels_init();
    if (rc == DCASE_OK) rc = cnn0_kernels_init();
    if (rc == DCASE_OK) rc = conv_tc_kernels_init();
    if (rc == DCASE_OK) rc = glu_tma_kernels_init();
    if (rc == DCASE_OK) rc = head_kernels_init();
    if (rc == DCASE_OK) rc = gru_kernels_init();
    if (rc == DCASE_OK) rc = gemm_tc_init();
    if (rc != DCASE_OK) { delete ctx; return rc; }
    *out = ctx;
    return DCASE_OK;
}

int dcase_ctx_destroy(dcase_ctx* ctx) {
    if (!ctx) return DCASE_OK;
    dcase_logmel_tables_destroy(ctx);
    cudaFree(ctx->d_loss_scratch);
    for (int i = 0; i < 2; ++i) {
        cudaStreamDestroy(ctx->prep_stream[i]);
        cudaEventDestroy(ctx->ev_prep_fork[i]);
        cudaEventDestroy(ctx->ev_prep_join[i]);
    }
    cudaStreamDestroy(ctx->aux_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    for (int i = 0; i < 4; ++i) { cudaEventDestroy(ctx->ev_bwd_fork[i]); cudaEventDestroy(ctx->ev_bwd_join[i]); }
    delete ctx;
    return DCASE_OK;
}

size_t dcase_crnn_param_count(int n_class) { return (size_t)param_offsets(n_class).total; }

long long dcase_crnn_param_offset(int n_class, const char* name) {
    if (!name) return -1;
    for (const auto& e : param_table(n_class)) if (e.name == name) return e.off;
    return -1;
}

size_t dcase_crnn_workspace_bytes(int B, int T, int n_class) { return ws_layout(B, T, n_class).total; }

int dcase_crnn_ws_tensor(int B, int T, int n_class, const char* name, size_t* offset_bytes, size_t* n_elems) {
    DCASE_REQUIRE(name && offset_bytes && n_elems, "null argument");
    const WsLayout L = ws_layout(B, T, n_class);
    for (const auto& x : L.e)
        if (!strcmp(x.name, name)) { *offset_bytes = x.off; *n_elems = x.n_elems; return DCASE_OK; }
    dcase_set_error("unknown workspace tensor '%s'", name);
    return DCASE_ERR_ARG;
}

// mom_in: the 54 tap moments of x computed ahead of the step (dcase_cnn0_input_moments) or NULL
static int crnn_forward_impl(dcase_ctx* ctx, const float* x, int B, int T, int NC, const float* params, float* bn_running,
                             int flags, uint64_t seed, uint32_t step, int model_id, const void* scalars, float* strong,
                             float* weak, void* ws, const double* mom_in, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && x && params && bn_running && strong && weak && ws, "null argument");
    DCASE_TRY(check_shape(B, T, NC));
    const POff o = param_offsets(NC);
    const WsLayout L = ws_layout(B, T, NC);
    const int training = (flags & DCASE_FLAG_BN_BATCH_STATS) ? 1 : 0;
    const int dropout = (flags & DCASE_FLAG_DROPOUT) ? 1 : 0;
    const DcaseStepScalars* sc = (const DcaseStepScalars*)scalars;
    const int sms = ctx->num_sms;
    auto drop = [&](int layer) { return DropoutCfg{dropout, seed, step, (uint32_t)(8 * model_id + layer), sc}; };
    const int To = T / 8;
    const int BT = B * To;

    // ---- conv weight images of blocks 1, 2: parameters only, so they are built beside block 0 on a side stream ----
    const int mid = model_id ? 1 : 0;
    cudaStream_t prep = g_prof_on ? s : ctx->prep_stream[mid];
    DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_prep_fork[mid], s));
    DCASE_CUDA_CHECK(cudaStreamWaitEvent(prep, ctx->ev_prep_fork[mid], 0));
    DCASE_TRY(launch_conv_w_prep(params + o.conv_w[1], nullptr, wsp<float>(ws, L, "wprep1_d"), wsp<void>(ws, L, "wprep1_h"),
                                 params + o.conv_w[2], nullptr, wsp<float>(ws, L, "wprep2_d"), wsp<void>(ws, L, "wprep2_h"), prep));
    DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_prep_join[mid], prep));

    // ---- CNN block 0 (fused, nothing materialised at [B,64,T,64]) ----
    double* mom0 = wsp<double>(ws, L, "mom0");
    float* fold0 = wsp<float>(ws, L, "fold0");
    const long long n_pix0 = (long long)B * T * 64;
    // exact-global-batch BatchNorm (dcase_ctx_set_syncbn): the sums behind every batch statistic cross the ranks between
    // their producer and their consumer; slots 3 * model + layer (forward), 6 + layer (backward)
    const dcase_syncbn* sb = training ? ctx->syncbn : nullptr;
    const long long world = syncbn_world(sb);
    if (training && !sb && !mom_in)
        DCASE_TRY(launch_cnn0_moments(x, B, T, mom0, params + o.conv_w[0], params + o.conv_b[0], params + o.bn_w[0],
                                      params + o.bn_b[0], bn_running, fold0, sms, s));
    else if (training && !sb)       // the moments only depend on the input: a pipelined caller computed them beside the previous step
        DCASE_TRY(launch_bn0_finalize(mom_in, n_pix0, params + o.conv_w[0], params + o.conv_b[0], params + o.bn_w[0],
                                      params + o.bn_b[0], bn_running, 1, fold0, mom0, s));
    else if (training) {
        if (mom_in) DCASE_CUDA_CHECK(cudaMemcpyAsync(mom0, mom_in, 54 * sizeof(double), cudaMemcpyDeviceToDevice, s));
        else DCASE_TRY(launch_cnn0_moments(x, B, T, mom0, params + o.conv_w[0], params + o.conv_b[0], params + o.bn_w[0],
                                           params + o.bn_b[0], bn_running, nullptr, sms, s));
        DCASE_TRY(syncbn_allreduce_f64(sb, mom0, 54, 3 * mid + 0, s));
        DCASE_TRY(launch_bn0_finalize(mom0, n_pix0 * world, params + o.conv_w[0], params + o.conv_b[0], params + o.bn_w[0],
                                      params + o.bn_b[0], bn_running, 1, fold0, nullptr, s));
    } else
        DCASE_TRY(launch_bn0_finalize(mom0, n_pix0, params + o.conv_w[0], params + o.conv_b[0], params + o.bn_w[0],
                                      params + o.bn_b[0], bn_running, training, fold0, nullptr, s));
    // The forward convs multiply fp16 copies of the block outputs (same 10 explicit mantissa bits as the tf32-rounded
    // fp32 values, half the operand bytes); the fp32 copies only feed the backward's weight-gradient MMAs, and the teacher
    // (model_id 1: no backward, main.py:87-89) does not write them.
    const bool keep_f32 = model_id == 0;
    float* out0 = wsp<float>(ws, L, "out0");
    DCASE_TRY(launch_cnn0_fwd(x, B, T, fold0, params + o.glu_w[0], params + o.glu_b[0], drop(0), keep_f32 ? out0 : nullptr,
                              wsp<void>(ws, L, "out0_h"), wsp<unsigned int>(ws, L, "tile_ctr"), sms, s));

    // ---- CNN blocks 1, 2 ----
    const char* names[3][7] = {{}, {"wprep1_h", "wprep1_d", "ypre1", "stats1", "bn1", "out1", "gluimg1"},
                               {"wprep2_h", "wprep2_d", "ypre2", "stats2", "bn2", "out2", "gluimg2"}};
    const void* in_h = wsp<void>(ws, L, "out0_h");
    for (int l = 1; l <= 2; ++l) {
        const int T_l = l == 1 ? T / 2 : T / 4;
        const int F_l = l == 1 ? 16 : 4;
        const int n_rows = B * T_l;
        const long long n_pix = (long long)n_rows * F_l;
        float* wf = wsp<float>(ws, L, names[l][0]);
        float* wd = wsp<float>(ws, L, names[l][1]);
        float* ypre = wsp<float>(ws, L, names[l][2]);
        double* stats = wsp<double>(ws, L, names[l][3]);
        float* bn = wsp<float>(ws, L, names[l][4]);
        float* out = wsp<float>(ws, L, names[l][5]);
        (void)wd;
        if (l == 1) DCASE_CUDA_CHECK(cudaStreamWaitEvent(s, ctx->ev_prep_join[mid], 0));
        // (folding this one-block pass into the conv's last CTA was measured: +10 us on the conv against 7.9 us for the
        // launch -- a single CTA pays the fp64 and L2 latencies that eight blocks share here)
        DCASE_TRY(launch_conv3x3_h(in_h, B, T_l, F_l, wf, params + o.conv_b[l], ypre, training ? stats : nullptr, sms, s));
        float* glu_img = wsp<float>(ws, L, names[l][6]);
        if (sb) DCASE_TRY(syncbn_allreduce_f64(sb, stats, 128, 3 * mid + l, s));
        DCASE_TRY(launch_bn_finalize(stats, n_pix * world, params + o.bn_w[l], params + o.bn_b[l], bn_running + l * 128,
                                     training, bn, params + o.glu_w[l], params + o.glu_b[l], F_l, glu_img, s));
        DCASE_TRY(launch_glu_pool_fwd(ypre, n_pix, F_l, glu_img, drop(l), out, l == 1 ? wsp<void>(ws, L, "out1_h") : nullptr, sms, s));
        in_h = wsp<void>(ws, L, "out1_h");
    }

    // ---- BiGRU, 2 layers (RNN.py:12-16) ----
    float* gi = wsp<float>(ws, L, "gi");
    const char* sv[2][5] = {{"s0_r", "s0_z", "s0_n", "s0_hn", "s0_hp"}, {"s1_r", "s1_z", "s1_n", "s1_hn", "s1_hp"}};
    const float* rin = wsp<float>(ws, L, "out2");
    for (int l = 0; l < 2; ++l) {
        const int nin = l == 0 ? kC : 2 * kH;
        float* rout = wsp<float>(ws, L, l == 0 ? "rnn0" : "rnn1");
        {   // input projections of both directions in one launch on the tensor cores: gi[d] = X W_ih[d]^T + b_ih[d]
            GemmTcBatch gb{};
            gb.problems = 2; gb.parts = 1; gb.M = BT; gb.N = 3 * kH; gb.K = nin;
            gb.lda = nin; gb.ldb = nin; gb.b_mn_major = 0; gb.ldc = 3 * kH;
            for (int d = 0; d < 2; ++d) {
                gb.A[d] = rin; gb.B[d] = params + o.w_ih[l][d];
                gb.C[d] = gi + (size_t)d * BT * 3 * kH; gb.bias[d] = params + o.b_ih[l][d];
            }
            DCASE_TRY(launch_gemm_tc(gb, s));
        }
        GruFwdArgs g{};
        g.gi = gi;
        for (int d = 0; d < 2; ++d) { g.w_hh[d] = params + o.w_hh[l][d]; g.b_hh[d] = params + o.b_hh[l][d]; }
        g.out = rout;
        if (training) {
            g.save_r = wsp<float>(ws, L, sv[l][0]); g.save_z = wsp<float>(ws, L, sv[l][1]);
            g.save_n = wsp<float>(ws, L, sv[l][2]); g.save_hn = wsp<float>(ws, L, sv[l][3]);
            g.save_hp = wsp<float>(ws, L, sv[l][4]);
        }
        g.B = B; g.T = To;
        DCASE_TRY(launch_gru_fwd(g, s));
        rin = rout;
    }

    // ---- head (CRNN.py:74-81) ----
    HeadArgs h{};
    h.x = rin; h.w_dense = params + o.dense_w; h.b_dense = params + o.dense_b;
    h.w_soft = params + o.soft_w; h.b_soft = params + o.soft_b;
    h.B = B; h.To = To; h.NC = NC; h.drop = dropout; h.seed = seed; h.step = step;
    h.stream = (uint32_t)(8 * model_id) + DCASE_STREAM_HEAD; h.sc = sc;
    h.strong = strong; h.weak = weak; h.den = wsp<float>(ws, L, "den");
    DCASE_TRY(launch_head_fwd(h, s));
    return DCASE_OK;
}

int dcase_crnn_forward(dcase_ctx* ctx, const float* x, int B, int T, int NC, const float* params, float* bn_running,
                       int flags, uint64_t seed, uint32_t step, int model_id, const void* scalars, float* strong,
                       float* weak, void* ws, void* stream_) {
    return crnn_forward_impl(ctx, x, B, T, NC, params, bn_running, flags, seed, step, model_id, scalars, strong, weak, ws,
                             nullptr, stream_);
}

int dcase_cnn0_input_moments(dcase_ctx* ctx, const float* x, int B, int T, double* mom, void* stream) {
    DCASE_REQUIRE(ctx && x && mom, "null argument");
    DCASE_TRY(check_shape(B, T, 10));
    return launch_cnn0_moments(x, B, T, mom, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->num_sms, (cudaStream_t)stream);
}

size_t dcase_bigru_workspace_bytes(int B, int To) {
    if (B < 1 || To < 1) return 0;
    const size_t BT = (size_t)B * To;
    return (BT * 2 * 3 * kH + BT * 2 * kH) * sizeof(float);          // gi of both directions + layer-0 output
}

int dcase_bigru_forward(dcase_ctx* ctx, const float* x, int B, int To, const float* rnn_params, float* out, void* ws,
                        void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && x && rnn_params && out && ws, "null argument");
    DCASE_REQUIRE(B >= 1 && To >= 1, "bad shape");
    const POff o = param_offsets(10);
    const long long base = o.w_ih[0][0];                             // rnn.rnn.weight_ih_l0 opens the GRU block
    const int BT = B * To;
    float* gi = (float*)ws;
    float* mid = gi + (size_t)BT * 2 * 3 * kH;
    const float* rin = x;
    for (int l = 0; l < 2; ++l) {
        const int nin = l == 0 ? kC : 2 * kH;
        float* rout = l == 0 ? mid : out;
        GemmTcBatch gb{};        // the same tensor-core projection the CRNN forward uses
        gb.problems = 2; gb.parts = 1; gb.M = BT; gb.N = 3 * kH; gb.K = nin;
        gb.lda = nin; gb.ldb = nin; gb.b_mn_major = 0; gb.ldc = 3 * kH;
        for (int d = 0; d < 2; ++d) {
            gb.A[d] = rin; gb.B[d] = rnn_params + (o.w_ih[l][d] - base);
            gb.C[d] = gi + (size_t)d * BT * 3 * kH; gb.bias[d] = rnn_params + (o.b_ih[l][d] - base);
        }
        DCASE_TRY(launch_gemm_tc(gb, s));
        GruFwdArgs g{};
        g.gi = gi;
        for (int d = 0; d < 2; ++d) {
            g.w_hh[d] = rnn_params + (o.w_hh[l][d] - base);
            g.b_hh[d] = rnn_params + (o.b_hh[l][d] - base);
        }
        g.out = rout;
        g.B = B; g.T = To;
        DCASE_TRY(launch_gru_fwd(g, s));
        rin = rout;
    }
    return DCASE_OK;
}

// fused_loss: the mean-teacher losses and their gradients are computed by the first phase of head_bwd (dcase_mt_fwd_bwd)
static int crnn_backward_impl(dcase_ctx* ctx, const float* x, int B, int T, int NC, const float* params, int flags,
                              uint64_t seed, uint32_t step, int model_id, const void* scalars, const float* d_strong,
                              const float* d_weak, const float* weak, void* ws, float* grads, const LossArgs* fused_loss,
                              void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && x && params && d_strong && d_weak && weak && ws && grads, "null argument");
    DCASE_TRY(check_shape(B, T, NC));
    DCASE_REQUIRE(flags & DCASE_FLAG_BN_BATCH_STATS, "backward needs the train-mode forward (batch statistics)");
    const POff o = param_offsets(NC);
    const WsLayout L = ws_layout(B, T, NC);
    const int dropout = (flags & DCASE_FLAG_DROPOUT) ? 1 : 0;
    const DcaseStepScalars* sc = (const DcaseStepScalars*)scalars;
    const int sms = ctx->num_sms;
    auto drop = [&](int layer) { return DropoutCfg{dropout, seed, step, (uint32_t)(8 * model_id + layer), sc}; };
    const int To = T / 8;
    const int BT = B * To;
    const dcase_syncbn* sb = ctx->syncbn;                 // see dcase_crnn_forward
    const long long world = syncbn_world(sb);
    const float pgs = 1.f / (float)world;                 // BatchNorm-derived parameter gradients come out as GLOBAL sums

    DCASE_CUDA_CHECK(cudaMemsetAsync(grads, 0, (size_t)o.total * sizeof(float), s));
    DCASE_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(ws) + L.acc_off, 0, L.acc_bytes, s));

    // ---- head ----
    HeadArgs h{};
    h.x = wsp<float>(ws, L, "rnn1"); h.w_dense = params + o.dense_w; h.b_dense = params + o.dense_b;
    h.w_soft = params + o.soft_w; h.b_soft = params + o.soft_b;
    h.B = B; h.To = To; h.NC = NC; h.drop = dropout; h.seed = seed; h.step = step;
    h.stream = (uint32_t)(8 * model_id) + DCASE_STREAM_HEAD; h.sc = sc;
    h.weak = const_cast<float*>(weak); h.den = wsp<float>(ws, L, "den");
    h.d_strong = d_strong; h.d_weak = d_weak; h.d_x = wsp<float>(ws, L, "d_rnn1");
    h.g_w_dense = grads + o.dense_w; h.g_b_dense = grads + o.dense_b;
    h.g_w_soft = grads + o.soft_w; h.g_b_soft = grads + o.soft_b;
    if (fused_loss) { h.fused_loss = 1; h.loss = *fused_loss; }
    DCASE_TRY(launch_head_bwd(h, s));

    // ---- BiGRU BPTT, layer 1 then layer 0 ----
    // Only the input gradient d_in feeds the next kernel of the chain; the weight / bias gradients of both directions
    // (4 split-K GEMMs + 4 column sums per layer) run on the context's second stream beside the next layer's recurrence
    // (latency bound, 48 CTAs) and are joined at the end of the backward.
    const char* sv[2][5] = {{"s0_r", "s0_z", "s0_n", "s0_hn", "s0_hp"}, {"s1_r", "s1_z", "s1_n", "s1_hn", "s1_hp"}};
    // while per-kernel profiling is on, everything runs on ONE stream so that the event-bracketed durations are the
    // kernels' own (two overlapping kernels would each be charged the other's time)
    cudaStream_t aux = g_prof_on ? s : ctx->aux_stream;
    for (int l = 1; l >= 0; --l) {
        const int nin = l == 0 ? kC : 2 * kH;
        float* dgi = wsp<float>(ws, L, l == 1 ? "dgi1" : "dgi0");
        float* dgh = wsp<float>(ws, L, l == 1 ? "dgh1" : "dgh0");
        const float* d_out = wsp<float>(ws, L, l == 1 ? "d_rnn1" : "d_rnn0");
        const float* xin = wsp<float>(ws, L, l == 1 ? "rnn0" : "out2");
        float* d_in = wsp<float>(ws, L, l == 1 ? "d_rnn0" : "d_out2");
        GruBwdArgs g{};
        g.d_out = d_out;
        for (int d = 0; d < 2; ++d) g.w_hh[d] = params + o.w_hh[l][d];
        g.save_r = wsp<float>(ws, L, sv[l][0]); g.save_z = wsp<float>(ws, L, sv[l][1]);
        g.save_n = wsp<float>(ws, L, sv[l][2]); g.save_hn = wsp<float>(ws, L, sv[l][3]);
        g.save_hp = wsp<float>(ws, L, sv[l][4]);
        g.dgi = dgi; g.dgh = dgh; g.B = B; g.T = To;
        DCASE_TRY(launch_gru_bwd(g, s));
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_bwd_fork[l], s));
        DCASE_CUDA_CHECK(cudaStreamWaitEvent(aux, ctx->ev_bwd_fork[l], 0));
        GemmTcBatch gin{};       // d_in [BT][nin] = dgi_fwd W_ih_fwd + dgi_bwd W_ih_bwd: two parts accumulate in tensor memory
        gin.problems = 1; gin.parts = 2; gin.M = BT; gin.N = nin; gin.K = 3 * kH;
        gin.lda = 3 * kH; gin.ldb = nin; gin.b_mn_major = 1; gin.ldc = nin;
        gin.C[0] = d_in; gin.bias[0] = nullptr;
        GemmBatch gw{};
        ColsumBatch cb{};
        for (int d = 0; d < 2; ++d) {
            const float* dgi_d = dgi + (size_t)d * BT * 3 * kH;
            const float* dgh_d = dgh + (size_t)d * BT * 3 * kH;
            const float* hp_d = g.save_hp + (size_t)d * BT * kH;
            // dW_ih [3H][nin] = dgi^T X ;  dW_hh [3H][H] = dgh^T Hprev
            gin.A[d] = dgi_d; gin.B[d] = params + o.w_ih[l][d];
            gw.p[2 * d + 0] = GemmProblem{3 * kH, nin, BT, dgi_d, 1, 3 * kH, xin, nin, 1, grads + o.w_ih[l][d], nin, nullptr};
            gw.p[2 * d + 1] = GemmProblem{3 * kH, kH, BT, dgh_d, 1, 3 * kH, hp_d, kH, 1, grads + o.w_hh[l][d], kH, nullptr};
            cb.A[2 * d] = dgi_d; cb.out[2 * d] = grads + o.b_ih[l][d];
            cb.A[2 * d + 1] = dgh_d; cb.out[2 * d + 1] = grads + o.b_hh[l][d];
        }
        gw.n = 4; gw.split = BT >= 512 ? 16 : 1; gw.mode = 1;
        cb.n = 4; cb.M = BT; cb.N = 3 * kH;
        DCASE_TRY(launch_gemm_tc(gin, s));
        DCASE_TRY(launch_sgemm_batch(gw, aux));
        DCASE_TRY(launch_colsum_batch(cb, aux));
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_bwd_join[l], aux));
    }

    // ---- CNN blocks 2, 1 ----
    const char* names[3][8] = {{}, {"ypre1", "bn1", "s12_1", "dy1", "d_out1", "out0", "wprep1_d", "d_out0"},
                               {"ypre2", "bn2", "s12_2", "dy2", "d_out2", "out1", "wprep2_d", "d_out1"}};
    for (int l = 2; l >= 1; --l) {
        const int T_l = l == 1 ? T / 2 : T / 4;
        const int F_l = l == 1 ? 16 : 4;
        const int n_rows = B * T_l;
        const long long n_pix = (long long)n_rows * F_l;
        const float* ypre = wsp<float>(ws, L, names[l][0]);
        const float* bn = wsp<float>(ws, L, names[l][1]);
        float* s12 = wsp<float>(ws, L, names[l][2]);
        float* dy = wsp<float>(ws, L, names[l][3]);
        const float* d_out = wsp<float>(ws, L, names[l][4]);
        const float* lin = wsp<float>(ws, L, names[l][5]);
        const float* wd = wsp<float>(ws, L, names[l][6]);
        float* d_in = wsp<float>(ws, L, names[l][7]);
        DCASE_TRY(launch_glu_pool_bwd(ypre, n_pix, F_l, bn, wsp<float>(ws, L, l == 1 ? "gluimg1" : "gluimg2"), drop(l), d_out, dy,
                                      s12, grads + o.glu_w[l], grads + o.glu_b[l], sms, s));
        if (sb) DCASE_TRY(syncbn_allreduce_f32(sb, s12, 128, 6 + l, s));
        DCASE_TRY(launch_bn_bwd_apply(dy, ypre, n_pix, n_pix * world, bn, params + o.bn_w[l], s12, pgs, grads + o.bn_w[l],
                                      grads + o.bn_b[l], grads + o.conv_b[l], sms, s));
        // the weight gradient only feeds the optimizer: second stream, beside the data gradient and the next block
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_bwd_fork[1 + l], s));
        DCASE_CUDA_CHECK(cudaStreamWaitEvent(aux, ctx->ev_bwd_fork[1 + l], 0));
        DCASE_TRY(launch_conv_wgrad(dy, lin, B, T_l, F_l, grads + o.conv_w[l], sms, aux));
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_bwd_join[1 + l], aux));
        DCASE_TRY(launch_conv3x3(dy, B, T_l, F_l, wd, nullptr, d_in, nullptr, sms, s));
    }

    // ---- CNN block 0 ----
    float* acc0 = wsp<float>(ws, L, "acc0");
    const float* fold0 = wsp<float>(ws, L, "fold0");
    const Cnn0BwdFinalize fin{wsp<double>(ws, L, "mom0"), (long long)B * T * 64 * world, params + o.conv_w[0], params + o.conv_b[0],
                              pgs, grads + o.conv_w[0], grads + o.conv_b[0], grads + o.bn_w[0], grads + o.bn_b[0],
                              grads + o.glu_w[0], grads + o.glu_b[0]};
    // the finalize pass runs in the last CTA of cnn0_bwd; SyncBN: the accumulator crosses the ranks first (mom0 did in the forward)
    DCASE_TRY(launch_cnn0_bwd(x, B, T, fold0, params + o.glu_w[0], params + o.glu_b[0], drop(0),
                              wsp<float>(ws, L, "d_out0"), acc0, sb ? nullptr : &fin, sms, s));
    if (sb) {
        DCASE_TRY(syncbn_allreduce_f32(sb, acc0, kCnn0AccFloats, 6, s));
        DCASE_TRY(launch_cnn0_bwd_finalize(fin, fold0, params + o.glu_w[0], acc0, s));
    }
    for (int l = 0; l < 4; ++l) DCASE_CUDA_CHECK(cudaStreamWaitEvent(s, ctx->ev_bwd_join[l], 0));
    return DCASE_OK;
}

int dcase_crnn_backward(dcase_ctx* ctx, const float* x, int B, int T, int NC, const float* params, int flags,
                        uint64_t seed, uint32_t step, int model_id, const void* scalars, const float* d_strong,
                        const float* d_weak, const float* weak, void* ws, float* grads, void* stream_) {
    return crnn_backward_impl(ctx, x, B, T, NC, params, flags, seed, step, model_id, scalars, d_strong, d_weak, weak, ws, grads,
                              nullptr, stream_);
}

static int fill_loss_args(dcase_ctx* ctx, const float* strong_s, const float* weak_s, const float* strong_t, const float* weak_t,
                          const float* target, int B, int To, int NC, int weak_lo, int weak_hi, int strong_lo, int strong_hi,
                          float cons_weight, const void* scalars, float* meters, float* d_strong, float* d_weak, LossArgs* out) {
    DCASE_REQUIRE(ctx && strong_s && weak_s && target && meters && d_strong && d_weak, "null argument");
    DCASE_REQUIRE((strong_t == nullptr) == (weak_t == nullptr), "teacher outputs must both be given or both be NULL");
    DCASE_REQUIRE(B >= 1 && To >= 1 && NC >= 1 && NC <= 16, "bad shape");
    DCASE_REQUIRE(weak_lo >= 0 && weak_hi <= B && strong_lo >= 0 && strong_hi <= B, "mask slice out of range");
    LossArgs a{};
    a.strong_s = strong_s; a.weak_s = weak_s; a.strong_t = strong_t; a.weak_t = weak_t; a.target = target;
    a.B = B; a.To = To; a.NC = NC; a.weak_lo = weak_lo; a.weak_hi = weak_hi; a.strong_lo = strong_lo; a.strong_hi = strong_hi;
    a.cons_weight = cons_weight; a.sc = (const DcaseStepScalars*)scalars; a.meters = meters;
    a.d_strong = d_strong; a.d_weak = d_weak;
    a.partials = ctx->d_loss_scratch;
    a.ticket = reinterpret_cast<unsigned int*>(ctx->d_loss_scratch + kLossMaxCtas * 8);
    *out = a;
    return DCASE_OK;
}

int dcase_mt_loss(dcase_ctx* ctx, const float* strong_s, const float* weak_s, const float* strong_t,
                  const float* weak_t, const float* target, int B, int To, int NC, int weak_lo, int weak_hi,
                  int strong_lo, int strong_hi, float cons_weight, const void* scalars, float* meters,
                  float* d_strong, float* d_weak, void* stream_) {
    LossArgs a{};
    DCASE_TRY(fill_loss_args(ctx, strong_s, weak_s, strong_t, weak_t, target, B, To, NC, weak_lo, weak_hi, strong_lo, strong_hi,
                             cons_weight, scalars, meters, d_strong, d_weak, &a));
    return launch_mt_loss(a, (cudaStream_t)stream_);
}

int dcase_adam_ema_step(dcase_ctx* ctx, float* p, const float* g, float* m, float* v, float* p_ema, size_t n, float lr,
                        float beta1, float beta2, float eps, int step_t, float ema_alpha, float grad_scale,
                        const void* scalars, void* stream_) {
    DCASE_REQUIRE(ctx && p && g && m && v, "null argument");
    DCASE_REQUIRE(step_t >= 1 || scalars, "Adam step count starts at 1");
    float bc1 = 1.f, bc2 = 1.f;
    if (step_t >= 1) {
        bc1 = (float)(1.0 - pow((double)beta1, (double)step_t));
        bc2 = (float)(1.0 - pow((double)beta2, (double)step_t));
    }
    return launch_adam_ema(p, g, m, v, p_ema, (long long)n, lr, beta1, beta2, eps, bc1, bc2, ema_alpha, grad_scale,
                           (const DcaseStepScalars*)scalars, ctx->num_sms, (cudaStream_t)stream_);
}

int dcase_mt_fwd_bwd(dcase_ctx* ctx, const dcase_mt_args* a, void* stream) {
    DCASE_REQUIRE(ctx && a, "null argument");
    const int To = a->T / 8;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->x_teacher) {
        // the teacher forward (no grad, main.py:87-89) is independent of the student forward until the losses:
        // fork it onto the context's second stream so the two overlap (latency-bound GRU / single-CTA kernels)
        DCASE_REQUIRE(a->params_t && a->bn_t && a->strong_t && a->weak_t && a->ws_t, "teacher buffers missing");
        // profiling: one stream, isolated kernel durations; SyncBN: both models' statistics exchanges spin on the peers,
        // so they stay in ONE stream order that is the same on every rank
        cudaStream_t ts = (g_prof_on || ctx->syncbn) ? s : ctx->aux_stream;
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, s));
        DCASE_CUDA_CHECK(cudaStreamWaitEvent(ts, ctx->ev_fork, 0));
        DCASE_TRY(crnn_forward_impl(ctx, a->x_teacher, a->B, a->T, a->n_class, a->params_t, a->bn_t, a->flags, a->seed,
                                    a->step, 1, a->scalars, a->strong_t, a->weak_t, a->ws_t, a->mom_t, ts));
        DCASE_CUDA_CHECK(cudaEventRecord(ctx->ev_join, ts));
    }
    DCASE_TRY(crnn_forward_impl(ctx, a->x_student, a->B, a->T, a->n_class, a->params_s, a->bn_s, a->flags, a->seed,
                                a->step, 0, a->scalars, a->strong_s, a->weak_s, a->ws_s, a->mom_s, stream));
    if (a->after_forward_event) DCASE_CUDA_CHECK(cudaEventRecord((cudaEvent_t)a->after_forward_event, s));
    if (a->x_teacher) DCASE_CUDA_CHECK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    // the losses run as the first phase of the backward's head kernel (one launch less between the forwards and the backward)
    LossArgs la{};
    DCASE_TRY(fill_loss_args(ctx, a->strong_s, a->weak_s, a->x_teacher ? a->strong_t : nullptr, a->x_teacher ? a->weak_t : nullptr,
                             a->target, a->B, To, a->n_class, a->weak_lo, a->weak_hi, a->strong_lo, a->strong_hi, a->cons_weight,
                             a->scalars, a->meters, a->d_strong, a->d_weak, &la));
    const bool fuse = a->B <= kLossMaxCtas;
    if (!fuse) DCASE_TRY(launch_mt_loss(la, s));
    DCASE_TRY(crnn_backward_impl(ctx, a->x_student, a->B, a->T, a->n_class, a->params_s, a->flags, a->seed, a->step, 0,
                                 a->scalars, a->d_strong, a->d_weak, a->weak_s, a->ws_s, a->grads, fuse ? &la : nullptr, stream));
    return DCASE_OK;
}

}  // extern "C"
