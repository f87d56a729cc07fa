// C-ABI of libdiffphar_b200.so (include/diffphar_b200.h): handle, weight packing, batch plan,
// and the orchestration of one denoiser evaluation / one reverse-diffusion run.
#include "common.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

// --------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void dp_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* dp_last_error(void) { return g_err; }
extern "C" int dp_abi_version(void) { return DP_ABI_VERSION; }

extern "C" int dp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

// --------------------------------------------------------------------------------------
// profiling spans (eager mode only)
// --------------------------------------------------------------------------------------
void prof_begin(dp_handle* h, int which, cudaStream_t st)
{
    if (!h->profile || h->spans.size() > 60000) return;
    if (h->profile == 2 && !h->capturing) return;              // in-graph mode: only the captured step is instrumented
    dp_handle::Span s; s.which = which;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    if (h->profile == 2) cudaEventRecordWithFlags(s.a, st, cudaEventRecordExternal);   // an event-record NODE of the graph
    else cudaEventRecord(s.a, st);
    h->spans.push_back(s);
}

void prof_end(dp_handle* h, cudaStream_t st)
{
    if (!h->profile || h->spans.empty()) return;
    if (h->profile == 2 && !h->capturing) return;
    if (h->profile == 2) cudaEventRecordWithFlags(h->spans.back().b, st, cudaEventRecordExternal);
    else cudaEventRecord(h->spans.back().b, st);
}

// in-graph mode: the spans' events were re-recorded by the replay that just finished on `st`
static int prof_collect_replay(dp_handle* h, cudaStream_t st)
{
    DP_CUDA(cudaStreamSynchronize(st));
    for (auto& s : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { h->prof_ms[s.which] += ms; h->prof_n[s.which] += 1; }
    }
    cudaGetLastError();
    return DP_OK;
}

// --------------------------------------------------------------------------------------
// allocation helpers
// --------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(std::vector<void*>& bag, T** out, size_t count)
{
    void* p = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    DP_CUDA(cudaMalloc(&p, bytes));
    bag.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return DP_OK;
}

static int upload(std::vector<void*>& bag, float** out, const std::vector<float>& v)
{
    int rc = dev_alloc(bag, out, v.size());
    if (rc) return rc;
    if (!v.empty()) DP_CUDA(cudaMemcpy(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return DP_OK;
}

static void free_bag(std::vector<void*>& bag)
{
    for (void* p : bag) cudaFree(p);
    bag.clear();
}

int GrowBuf::reserve(size_t bytes, bool* moved)
{
    if (moved) *moved = false;
    if (bytes <= cap && p) return DP_OK;
    DP_CUDA(cudaDeviceSynchronize());                      // earlier work may still read the old allocation
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const size_t want = bytes + bytes / 4 + 256;
    DP_CUDA(cudaMalloc(&p, want));
    cap = want;
    if (moved) *moved = true;
    return DP_OK;
}

// Carves the per-plan buffers out of the handle's arena: first pass sizes, second pass assigns.
namespace {
struct Carver {
    struct Slot { void** out; size_t off; };
    std::vector<Slot> slots;
    size_t total = 0;
    template <typename T> void take(T** out, size_t count)
    {
        total = (total + 255) & ~(size_t)255;
        slots.push_back({reinterpret_cast<void**>(out), total});
        total += (count ? count : 1) * sizeof(T);
    }
    void assign(void* base) { for (auto& sl : slots) *sl.out = static_cast<unsigned char*>(base) + sl.off; }
};
}  // namespace

// --------------------------------------------------------------------------------------
// create / destroy
// --------------------------------------------------------------------------------------
extern "C" int dp_create(const dp_config* cfg, int device, dp_handle** out)
{
    DP_CHECK(cfg && out, DP_ERR_INVALID, "dp_create: null argument");
    DP_CHECK(cfg->hidden_nf == H, DP_ERR_INVALID, "hidden_nf must be %d (got %d): tiles are compile-time", H, cfg->hidden_nf);
    DP_CHECK(cfg->n_dims == 3, DP_ERR_INVALID, "n_dims must be 3");
    DP_CHECK(cfg->phar_nf > 0 && cfg->phar_nf <= 64 && cfg->residue_nf > 0 && cfg->residue_nf <= 64,
             DP_ERR_INVALID, "phar_nf/residue_nf must be in [1,64]");
    DP_CHECK(cfg->joint_nf > 0 && cfg->joint_nf <= 64, DP_ERR_INVALID, "joint_nf must be in [1,64]");
    DP_CHECK(cfg->n_layers > 0 && cfg->inv_sublayers > 0, DP_ERR_INVALID, "n_layers/inv_sublayers must be positive");
    DP_CHECK(cfg->precision >= DP_FP32 && cfg->precision <= DP_F16_FAST32, DP_ERR_INVALID, "unknown precision %d", cfg->precision);
    int n = 0;
    DP_CUDA(cudaGetDeviceCount(&n));
    DP_CHECK(device >= 0 && device < n, DP_ERR_INVALID, "device %d out of range (%d visible)", device, n);
    int major = 0;
    DP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    DP_CHECK(major == 10, DP_ERR_INVALID, "device %d is sm_%dx; this library is built for sm_100a only", device, major);
    DP_CUDA(cudaSetDevice(device));
    dp_handle* h = new dp_handle();
    h->cfg = *cfg;
    h->device = device;
    h->precision = cfg->precision;
    DP_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (const char* m = getenv("DIFFPHAR_TC_MASK")) {
        // debug switch: 3 = tcgen05 edge + node kernels (default), 0 = FFMA kernels even in the 16-bit modes.
        // Mixed settings are not meaningful: the tensor-core path keeps pq pre-scaled by 1/2.
        h->tc_mask = atoi(m) ? 3 : 0;
    }
    if (const char* m = getenv("DIFFPHAR_PDL")) h->pdl = atoi(m) != 0;
    if (const char* m = getenv("DIFFPHAR_SEG")) h->seg_mode = !strcmp(m, "units") ? 1 : !strcmp(m, "lanes") ? 2 : 0;   // tests force a scheme
    if (const char* m = getenv("DIFFPHAR_TMA_FILL")) h->tma_fill = atoi(m);
    if (const char* m = getenv("DIFFPHAR_NODE_PAIR")) h->node_pair = atoi(m);
    if (const char* m = getenv("DIFFPHAR_NODE_MC")) h->node_mc = atoi(m);
    if (const char* m = getenv("DIFFPHAR_NODE_SPLIT")) h->node_split = atoi(m);
    if (const char* m = getenv("DIFFPHAR_NODE_H16")) h->node_h16 = atoi(m);
    if (const char* m = getenv("DIFFPHAR_COORD_FUSED")) h->coord_fused = atoi(m);
    if (const char* m = getenv("DIFFPHAR_TRACE_CTA")) h->trace_cta = atoi(m);
    if (const char* m = getenv("DIFFPHAR_TRACE_V")) h->trace_v = atoi(m);
    if (const char* m = getenv("DIFFPHAR_DBG")) h->dbg = atoi(m);
    if (const char* m = getenv("DIFFPHAR_SKIP")) h->skip_mask = atoi(m);
    if (const char* m = getenv("DIFFPHAR_GRAPH")) h->graph_mode = !strcmp(m, "scan") ? 1 : !strcmp(m, "cells") ? 2 : !strcmp(m, "fused") ? 4 : 0;
    if (const char* m = getenv("DIFFPHAR_TRACE")) {
        if (atoi(m)) {
            h->trace_kernel = atoi(m);
            DP_CUDA(cudaMalloc(&h->trace, DP_TRACE_WORDS * sizeof(long long)));
            DP_CUDA(cudaMemset(h->trace, 0, DP_TRACE_WORDS * sizeof(long long)));
        }
    }
    DP_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    DP_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    DP_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    int rc = egnn_f32_init();
    if (!rc) rc = tc_init();
    if (rc) { delete h; return rc; }
    *out = h;
    return DP_OK;
}

static void drop_graph(dp_handle* h)
{
    if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
}

static void free_plan(dp_handle* h)
{
    drop_graph(h);
    h->plan = Plan();
    h->has_plan = false;
}

extern "C" int dp_destroy(dp_handle* h)
{
    if (!h) return DP_OK;
    cudaSetDevice(h->device);
    free_plan(h);
    h->arena.release(); h->steps.release(); h->noise.release(); h->frames.release();
    free_bag(h->w.allocations);
    tc_free_weights(h);
    for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->trace) cudaFree(h->trace);
    delete h;
    return DP_OK;
}

// --------------------------------------------------------------------------------------
// weights
// --------------------------------------------------------------------------------------
static int64_t weight_count_of(const dp_config& c)
{
    const int64_t P = c.phar_nf, R = c.residue_nf, J = c.joint_nf, D = J + (c.condition_time ? 1 : 0);
    auto lin = [](int64_t o, int64_t i, bool b = true) { return o * i + (b ? o : 0); };
    int64_t n = lin(2 * P, P) + lin(J, 2 * P) + lin(2 * P, J) + lin(P, 2 * P) + lin(2 * R, R) + lin(J, 2 * R) +
                lin(2 * R, J) + lin(R, 2 * R) + lin(H, D) + lin(D, H);
    const int64_t gcl = lin(H, 2 * H + 2) + lin(H, H) + lin(H, 2 * H) + lin(H, H) + (c.attention ? lin(1, H) : 0);
    const int64_t crd = lin(H, 2 * H + 2) + lin(H, H) + lin(1, H, false);
    n += (int64_t)c.n_layers * (c.inv_sublayers * gcl + crd);
    return n;
}

extern "C" int64_t dp_weight_count(const dp_handle* h) { return h ? weight_count_of(h->cfg) : 0; }

namespace {
struct BlobReader {
    const float* p;
    const float* take(int64_t n) { const float* r = p; p += n; return r; }
};

// [out][in] row-major + bias  ->  transposed device linear
int make_linear(std::vector<void*>& bag, DevLinear& L, const float* w, const float* b, int out, int in)
{
    std::vector<float> t((size_t)in * out);
    for (int o = 0; o < out; ++o)
        for (int k = 0; k < in; ++k) t[(size_t)k * out + o] = w[(size_t)o * in + k];
    L.in = in; L.out = out;
    int rc = upload(bag, &L.wt, t);
    if (rc) return rc;
    if (b) {
        std::vector<float> bv(b, b + out);
        rc = upload(bag, &L.b, bv);
    }
    return rc;
}

struct HostFirstLayer { const float* w; const float* b; };   // an [H][2H+2] first edge/coord layer
}  // namespace

extern "C" int dp_set_weights(dp_handle* h, const float* blob, int64_t n_floats)
{
    DP_CHECK(h && blob, DP_ERR_INVALID, "dp_set_weights: null argument");
    const dp_config& c = h->cfg;
    DP_CHECK(n_floats == weight_count_of(c), DP_ERR_INVALID, "weight blob has %lld floats, expected %lld",
             (long long)n_floats, (long long)weight_count_of(c));
    DP_CUDA(cudaSetDevice(h->device));
    free_bag(h->w.allocations);
    h->w = DeviceWeights();
    DeviceWeights& W = h->w;
    auto& bag = W.allocations;
    const int P = c.phar_nf, R = c.residue_nf, J = c.joint_nf, D = J + (c.condition_time ? 1 : 0);
    BlobReader rd{blob};
    int rc = 0;
    auto lin = [&](DevLinear& L, int out, int in) {
        const float* w = rd.take((int64_t)out * in);
        const float* b = rd.take(out);
        return make_linear(bag, L, w, b, out, in);
    };
    if ((rc = lin(W.phar_enc0, 2 * P, P))) return rc;
    if ((rc = lin(W.phar_enc2, J, 2 * P))) return rc;
    if ((rc = lin(W.phar_dec0, 2 * P, J))) return rc;
    if ((rc = lin(W.phar_dec2, P, 2 * P))) return rc;
    if ((rc = lin(W.res_enc0, 2 * R, R))) return rc;
    if ((rc = lin(W.res_enc2, J, 2 * R))) return rc;
    if ((rc = lin(W.res_dec0, 2 * R, J))) return rc;
    if ((rc = lin(W.res_dec2, R, 2 * R))) return rc;
    if ((rc = lin(W.emb, H, D))) return rc;
    if ((rc = lin(W.emb_out, D, H))) return rc;

    const int S = c.inv_sublayers, G = c.n_layers * S;
    // lin_id numbering shared with run_denoiser: GCL i: 4i+0 edge_mlp.2, 4i+1 node_mlp.0, 4i+2 node_mlp.2;
    // block b: 4G+b coord_mlp.2; projection set v: 4G+L+v
    h->tc_host.assign((size_t)4 * G + c.n_layers + G + 1, HostLinear());
    auto reg = [&](int id, const float* w, int out, int in) {       // [out][in] -> k-major copy
        HostLinear& L = h->tc_host[id];
        L.K = in; L.n_out = out; L.wt.resize((size_t)in * out);
        for (int o = 0; o < out; ++o)
            for (int k = 0; k < in; ++k) L.wt[(size_t)k * out + o] = w[(size_t)o * in + k];
    };
    W.gcl.resize(G);
    W.coord.resize(c.n_layers);
    std::vector<HostFirstLayer> gcl_first(G), coord_first(c.n_layers);
    const int K1 = 2 * H + 2;
    auto scal_cols = [&](const float* w, float** wr, float** wd) {
        std::vector<float> r(H), d(H);
        for (int o = 0; o < H; ++o) { r[o] = w[(size_t)o * K1 + 2 * H]; d[o] = w[(size_t)o * K1 + 2 * H + 1]; }
        int e = upload(bag, wr, r);
        return e ? e : upload(bag, wd, d);
    };
    for (int b = 0; b < c.n_layers; ++b) {
        for (int g = 0; g < S; ++g) {
            GclWeights& L = W.gcl[b * S + g];
            gcl_first[b * S + g].w = rd.take((int64_t)H * K1);
            gcl_first[b * S + g].b = rd.take(H);
            if ((rc = scal_cols(gcl_first[b * S + g].w, &L.wr, &L.wd))) return rc;
            reg(4 * (b * S + g) + 0, rd.p, H, H);
            if ((rc = lin(L.e2, H, H))) return rc;
            reg(4 * (b * S + g) + 1, rd.p, H, 2 * H);
            if ((rc = lin(L.n0, H, 2 * H))) return rc;
            reg(4 * (b * S + g) + 2, rd.p, H, H);
            if ((rc = lin(L.n2, H, H))) return rc;
            if (c.attention) {
                const float* wa = rd.take(H);
                const float* ba = rd.take(1);
                std::vector<float> v(wa, wa + H);
                if ((rc = upload(bag, &L.wa, v))) return rc;
                L.ba = ba[0];
            }
        }
        CoordWeights& Cw = W.coord[b];
        coord_first[b].w = rd.take((int64_t)H * K1);
        coord_first[b].b = rd.take(H);
        if ((rc = scal_cols(coord_first[b].w, &Cw.wr, &Cw.wd))) return rc;
        reg(4 * G + b, rd.p, H, H);
        if ((rc = lin(Cw.c2, H, H))) return rc;
        const float* w4 = rd.take(H);
        std::vector<float> v(w4, w4 + H);
        if ((rc = upload(bag, &Cw.w4, v))) return rc;
    }
    DP_CHECK(rd.p - blob == n_floats, DP_ERR_INVALID, "internal: blob walk consumed %lld of %lld floats",
             (long long)(rd.p - blob), (long long)n_floats);

    // Projection sets: h version v (v = 0 after the embedding, v = i+1 after GCL i) feeds the first
    // layers of its consumers as ONE per-node GEMM with (row-part | col-part) outputs per consumer.
    W.proj.resize(G + 1);
    for (int v = 0; v <= G; ++v) {
        std::vector<HostFirstLayer> cons;
        ProjSet& ps = W.proj[v];
        if (v > 0 && v % S == 0) { ps.off_coord = (int)cons.size() * 2 * H; cons.push_back(coord_first[v / S - 1]); }
        if (v < G) { ps.off_gcl = (int)cons.size() * 2 * H; cons.push_back(gcl_first[v]); }
        const int n_out = (int)cons.size() * 2 * H;
        ps.lin.in = H; ps.lin.out = n_out;
        if (n_out == 0) continue;
        std::vector<float> wt((size_t)H * n_out), bias(n_out, 0.f);
        for (size_t ci = 0; ci < cons.size(); ++ci) {
            const int off = (int)ci * 2 * H;
            for (int o = 0; o < H; ++o) {
                bias[off + o] = cons[ci].b[o];                       // bias rides on the row part
                for (int k = 0; k < H; ++k) {
                    wt[(size_t)k * n_out + off + o] = cons[ci].w[(size_t)o * K1 + k];           // acts on h[row]
                    wt[(size_t)k * n_out + off + H + o] = cons[ci].w[(size_t)o * K1 + H + k];   // acts on h[col]
                }
            }
        }
        if ((rc = upload(bag, &ps.lin.wt, wt))) return rc;
        if ((rc = upload(bag, &ps.lin.b, bias))) return rc;
        // tensor-core image: weights and bias scaled by 1/2 (exact in every format), so the edge kernels get
        // hv = x / 2 of the factored first layer without a multiply (tc_edge.cu)
        std::vector<float> bias_half(bias);
        for (float& b : bias_half) b *= 0.5f;
        if ((rc = upload(bag, &ps.b_half, bias_half))) return rc;
        HostLinear& TL = h->tc_host[4 * G + c.n_layers + v];
        TL.K = H; TL.n_out = n_out; TL.wt = wt; TL.tf32_scale = 2.0f;
        for (float& w : TL.wt) w *= 0.5f;
    }
    if ((rc = tc_prepare_weights(h))) return rc;
    h->tc_host.clear();
    h->tc_host.shrink_to_fit();
    h->has_weights = true;
    drop_graph(h);
    return DP_OK;
}

extern "C" int dp_set_precision(dp_handle* h, int precision)
{
    DP_CHECK(h, DP_ERR_INVALID, "null handle");
    DP_CHECK(precision >= DP_FP32 && precision <= DP_F16_FAST32, DP_ERR_INVALID, "unknown precision %d", precision);
    h->precision = precision;
    return DP_OK;
}

extern "C" int dp_set_update_pocket_coords(dp_handle* h, int32_t on)
{
    DP_CHECK(h, DP_ERR_INVALID, "null handle");
    if (h->joint != (on != 0)) drop_graph(h);                 // the captured step bakes the coordinate mask
    h->joint = on != 0;
    return DP_OK;
}

// --------------------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------------------
static int bind_step_table(dp_handle* h);

extern "C" int dp_plan(dp_handle* h, int32_t B, const int32_t* phar_counts, const int32_t* res_counts, int64_t edge_capacity)
{
    DP_CHECK(h && phar_counts && res_counts, DP_ERR_INVALID, "dp_plan: null argument");
    DP_CHECK(B > 0, DP_ERR_INVALID, "dp_plan: n_samples must be positive");
    DP_CUDA(cudaSetDevice(h->device));
    // the same layout again (a pocket list with repeated shapes, the mirror's per-call plan): nothing to do, the
    // captured step graph stays valid
    if (h->has_plan && h->plan.B == B && h->plan.ecap_request == edge_capacity &&
        std::equal(phar_counts, phar_counts + B, h->plan.phar_counts_host.begin()) &&
        std::equal(res_counts, res_counts + B, h->plan.res_counts_host.begin()))
        return DP_OK;
    DP_CUDA(cudaDeviceSynchronize());                          // the previous layout's work is done before its buffers are re-carved
    free_plan(h);
    Plan& p = h->plan;
    const dp_config& c = h->cfg;
    std::vector<int> poff(B + 1, 0), roff(B + 1, 0);
    double pairs = 0;
    for (int b = 0; b < B; ++b) {
        DP_CHECK(phar_counts[b] >= 0 && res_counts[b] >= 0, DP_ERR_INVALID, "negative node count in sample %d", b);
        poff[b + 1] = poff[b] + phar_counts[b];
        roff[b + 1] = roff[b] + res_counts[b];
        const double nb = (double)phar_counts[b] + res_counts[b];
        pairs += nb * nb;
        if (phar_counts[b] > p.max_phar) p.max_phar = phar_counts[b];
        if (phar_counts[b] + res_counts[b] > p.max_nodes) p.max_nodes = phar_counts[b] + res_counts[b];
    }
    p.B = B; p.Np = poff[B]; p.Nr = roff[B]; p.N = p.Np + p.Nr;
    DP_CHECK(p.N > 0, DP_ERR_INVALID, "dp_plan: empty batch");
    p.phar_counts_host.assign(phar_counts, phar_counts + B);
    p.res_counts_host.assign(res_counts, res_counts + B);
    p.ecap_request = edge_capacity;
    int64_t ecap = edge_capacity;
    if (ecap <= 0) {
        const double guess = (double)p.N * 128.0 > 65536.0 ? (double)p.N * 128.0 : 65536.0;
        ecap = (int64_t)((c.edge_cutoff < 0.f || pairs < guess) ? pairs : guess);
    }
    DP_CHECK(ecap < (int64_t)2147483000, DP_ERR_INVALID, "edge capacity %lld exceeds int32 indexing", (long long)ecap);
    // segmented-sum scheme of the tcgen05 edge kernel (common.cuh): contiguous lanes for full-atom pockets (the same
    // size criterion as the cell-list graph builder), round-robin tiles with per-unit partial rows for Calpha pockets
    // (and for Calpha batches too large for the per-unit bookkeeping: agg_src encodes at most 2^20 units)
    p.seg_lanes = h->seg_mode == 2 || (h->seg_mode == 0 && (p.max_nodes >= 512 || ecap / UNIT_TC + 2 >= (int64_t)(1 << 20)));
    p.n_lanes = p.seg_lanes ? 4 * h->sm_count : 0;
    DP_CHECK(p.n_lanes <= MAX_LANES, DP_ERR_INVALID, "%d SMs: more segmented-sum lanes than agg_src can encode", h->sm_count);
    DP_CHECK(p.seg_lanes || ecap / UNIT_TC + 2 < (int64_t)(1 << 20), DP_ERR_INVALID,
             "edge capacity %lld too large for the per-unit segmented-sum scheme (set DIFFPHAR_SEG=lanes)", (long long)ecap);
    DP_CHECK((int64_t)p.N + 2 * (ecap / UNIT_TC + 2 + MAX_LANES) < (int64_t)2147483000, DP_ERR_INVALID, "batch too large for int32 row indexing");
    p.Ecap = ecap;
    std::vector<int> sample_of(p.N);
    std::vector<int64_t> ids(B);
    for (int b = 0; b < B; ++b) {
        ids[b] = b;
        for (int i = poff[b]; i < poff[b + 1]; ++i) sample_of[i] = b;
        for (int i = roff[b]; i < roff[b + 1]; ++i) sample_of[p.Np + i] = b;
    }
    const int PW = 3 + c.phar_nf, RW = 3 + c.residue_nf;
    // partial rows: per 64-edge tile on the FFMA path, per lane on the tcgen05 path (a handful of KB instead of E / 8 KB)
    size_t units = (size_t)(ecap / UNIT_F32 + 2);
    if (p.seg_lanes && (size_t)p.n_lanes > units) units = (size_t)p.n_lanes;
    if (!p.seg_lanes) units = (size_t)(ecap / UNIT_TC + 2);
    // bucketed cell list: pays off once a sample has more nodes than a handful of warp sweeps (full-atom pockets)
    p.use_cells = c.edge_cutoff > 0.f && p.max_nodes <= CELL_SAMPLE_MAX_NODES &&
                  (h->graph_mode == 2 || (h->graph_mode == 0 && p.max_nodes >= 512));
    // one-launch scan builder: a sample's candidates must fit 32 ballot chunks (two ranges, each rounded up to 32)
    // (DIFFPHAR_GRAPH=fused; OFF by default: measured 2 % slower per config-2 step than the three launches, graph.cu)
    p.fused_graph = h->graph_mode == 4 && !p.use_cells && !p.seg_lanes && p.max_nodes <= 960;
    // One arena for every per-plan buffer: a layout that fits the arena re-carves it (no cudaMalloc / cudaFree, each of
    // which synchronises the device), so walking a pocket list re-plans in microseconds.
    Carver cv;
    cv.take(&p.phar_off, B + 1); cv.take(&p.res_off, B + 1); cv.take(&p.sample_of, p.N); cv.take(&p.sample_ids, B);
    cv.take(&p.deg, p.N); cv.take(&p.rowptr, p.N + 1); cv.take(&p.agg_src, p.N);
    cv.take(&p.col, ecap); cv.take(&p.erow, ecap); cv.take(&p.edst, ecap); cv.take(&p.d0, ecap); cv.take(&p.escal, ecap);
    cv.take(&p.counts, 4);
    cv.take(&p.cpart, ((size_t)ecap / UNIT_TC + 2) * 8); cv.take(&p.cticket, p.N);
    cv.take(&p.scan_status, (size_t)p.N / 8 + 2); cv.take(&p.scan_ticket, 4);
    if (p.use_cells) {
        cv.take(&p.cell_start, (size_t)B * (CELLS_MAX + 1)); cv.take(&p.cell_nodes, p.N); cv.take(&p.cell_grid, (size_t)B * 8);
        p.bitmap_words = (p.max_nodes + 31) / 32;
        cv.take(&p.row_bitmap, (size_t)p.N * p.bitmap_words);  // 8 MB at config 3: the fill pass skips the bucket search
    }
    cv.take(&p.h, (size_t)p.N * H); cv.take(&p.tbuf, (size_t)p.N * H); cv.take(&p.h_base, (size_t)p.Nr * H);
    cv.take(&p.agg, ((size_t)p.N + units * 2) * H);        // [agg rows | partial rows] contiguous (graph.cu edge_dst)
    cv.take(&p.pq, (size_t)p.N * 4 * H);
    {
        NodeTiling nt, nu;                                   // with and without the phar-row split (joint mode / switches may change later)
        node_tiling(h, p.N, p.Np, &nt);
        node_tiling(h, p.N, 0, &nu);
        cv.take(&p.h16, (size_t)(std::max(nt.grid, nu.grid) + 1) * node_tile_image_bytes());
    }
    cv.take(&p.x_in, (size_t)p.N * 3); cv.take(&p.x_a, (size_t)p.N * 3); cv.take(&p.x_b, (size_t)p.N * 3);
    cv.take(&p.z, (size_t)p.Np * PW); cv.take(&p.eps_hat, (size_t)p.Np * PW); cv.take(&p.pocket, (size_t)p.Nr * RW);
    cv.take(&p.out_buf, (size_t)p.Np * PW);
    cv.take(&p.t_const, 4); cv.take(&p.step_idx, 1); cv.take(&p.nan_flag, 4);
    int rc = h->arena.reserve(cv.total);
    if (rc) return rc;
    cv.assign(h->arena.p);
    p.partials = p.agg + (size_t)p.N * H;
    DP_CUDA(cudaMemcpy(p.phar_off, poff.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(p.res_off, roff.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(p.sample_of, sample_of.data(), (size_t)p.N * sizeof(int), cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(p.sample_ids, ids.data(), (size_t)B * sizeof(int64_t), cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemset(p.counts, 0, 4 * sizeof(int)));
    DP_CUDA(cudaMemset(p.cticket, 0, (size_t)p.N * sizeof(int)));             // self-resetting arrival tickets (tc_edge.cu, coordinate mode)
    DP_CUDA(cudaMemset(p.scan_ticket, 0, 4 * sizeof(int)));                   // the count pass's arrival ticket (graph.cu)
    DP_CUDA(cudaMemset(p.nan_flag, 0, 4 * sizeof(int)));
    DP_CUDA(cudaMemset(p.step_idx, 0, sizeof(int)));
    DP_CUDA(cudaMemset(p.t_const, 0, 4 * sizeof(float)));
    h->has_plan = true;
    return bind_step_table(h);
}

// The step table lives in a handle-owned buffer (it does not depend on the batch layout): a new plan only re-binds it.
static int bind_step_table(dp_handle* h)
{
    Plan& p = h->plan;
    if (h->n_steps <= 0 || !h->steps.p) { p.step_rows = nullptr; p.stats = nullptr; p.stats_cap = 0; return DP_OK; }
    p.step_rows = reinterpret_cast<float*>(h->steps.p);
    p.stats = p.step_rows + (((size_t)h->n_steps * 4 + 63) & ~(size_t)63);
    p.stats_cap = h->n_steps + 2;
    return DP_OK;
}

extern "C" int dp_set_step_table(dp_handle* h, const float* rows, int32_t n_steps, const float* fin)
{
    DP_CHECK(h && rows && fin && n_steps > 0, DP_ERR_INVALID, "dp_set_step_table: bad argument");
    DP_CUDA(cudaSetDevice(h->device));
    // unchanged table (the mirror sets it on every sample_given_pocket call): keep buffers and the captured graph
    if (h->n_steps == n_steps && h->steps.p && h->step_rows_host.size() == (size_t)n_steps * 4 &&
        !memcmp(h->step_rows_host.data(), rows, (size_t)n_steps * 4 * sizeof(float)) && !memcmp(h->final_host, fin, 4 * sizeof(float)))
        return bind_step_table(h);
    if (rows != h->step_rows_host.data()) h->step_rows_host.assign(rows, rows + (size_t)n_steps * 4);
    memcpy(h->final_host, fin, 4 * sizeof(float));
    const bool new_count = n_steps != h->n_steps;
    h->n_steps = n_steps;
    const size_t rows_f = ((size_t)n_steps * 4 + 63) & ~(size_t)63, stats_f = (size_t)(n_steps + 2) * 2;
    bool moved = false;
    DP_CUDA(cudaDeviceSynchronize());
    int rc = h->steps.reserve((rows_f + stats_f) * sizeof(float), &moved);
    if (rc) return rc;
    DP_CUDA(cudaMemcpy(h->steps.p, h->step_rows_host.data(), (size_t)n_steps * 4 * sizeof(float), cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemset(reinterpret_cast<float*>(h->steps.p) + rows_f, 0, stats_f * sizeof(float)));
    if (moved || new_count) drop_graph(h);                    // the graph bakes the table pointer and the step count
    return bind_step_table(h);
}

// --------------------------------------------------------------------------------------
// graph
// --------------------------------------------------------------------------------------
static int require(dp_handle* h, bool weights, bool plan)
{
    DP_CHECK(h, DP_ERR_INVALID, "null handle");
    DP_CHECK(!weights || h->has_weights, DP_ERR_STATE, "dp_set_weights has not been called");
    DP_CHECK(!plan || h->has_plan, DP_ERR_STATE, "dp_plan has not been called");
    DP_CUDA(cudaSetDevice(h->device));
    return DP_OK;
}

extern "C" int dp_build_edges(dp_handle* h, const float* x_dev, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    DP_CHECK(x_dev, DP_ERR_INVALID, "dp_build_edges: null x");
    return launch_build_edges(h, x_dev, (cudaStream_t)stream);
}

extern "C" int dp_get_graph(dp_handle* h, const int32_t** rowptr, const int32_t** col, int64_t* n_edges, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    int counts[4];
    DP_CUDA(cudaMemcpyAsync(counts, h->plan.counts, sizeof(counts), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    DP_CHECK(counts[2] == 0, DP_ERR_CAPACITY, "graph has %d edges but edge_capacity is %lld", counts[2], (long long)h->plan.Ecap);
    if (rowptr) *rowptr = h->plan.rowptr;
    if (col) *col = h->plan.col;
    if (n_edges) *n_edges = counts[0];
    return DP_OK;
}

// --------------------------------------------------------------------------------------
// one denoiser evaluation (EGNNDynamics.forward, dynamics.py:75-139)
// --------------------------------------------------------------------------------------
static int run_linear(dp_handle* h, const LinearArgs& a, int lin_id, cudaStream_t st)
{
    prof_begin(h, PROF_NODE, st);
    // DP_TF32: the same per-node GEMMs on kind::tf32 tiles; the 16-bit modes run the fused node kernel instead (tc_node.cu)
    int rc = h->precision == DP_TF32 ? launch_linear_tf32(h, a, lin_id, st) : launch_linear_f32(h, a, st);
    prof_end(h, st);
    return rc;
}

static int run_edge(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st)
{
    if (h->skip_mask & (a.coord ? 4 : 1)) return DP_OK;
    prof_begin(h, a.coord ? PROF_EDGE_COORD : PROF_EDGE_MSG, st);
    // profile mode 3: the message kernel (a pure function of its inputs) runs DP_PROFILE_REPEAT times back to back inside
    // one event pair, so the pair's own cost (~5 us: a span around ONE 20 us kernel overstates it by a quarter) is amortised
    const int reps = (h->profile == 3 && !a.coord) ? DP_PROFILE_REPEAT : 1;
    int rc = DP_OK;
    for (int r = 0; r < reps && !rc; ++r)
        rc = h->precision == DP_TF32 ? launch_edge_tf32(h, a, lin_id, st)
             : (h->precision == DP_FP32 || !(h->tc_mask & 1)) ? launch_edge_f32(h, a, st) : launch_edge_tc(h, a, lin_id, st);
    prof_end(h, st);
    return rc;
}

// lin_id numbering for the tensor-core weight images: per GCL i: 4i+0 = edge_mlp.2, 4i+1 = node_mlp.0,
// 4i+2 = node_mlp.2; per block b: 4G + b = coord_mlp.2; projections: 4G + L + v.
static int run_denoiser(dp_handle* h, const float* xh_phar, const float* xh_res, const float* t_base,
                        const int* step_idx, int row_stride, int t_stride, float* out_phar, float* out_res,
                        cudaStream_t st, bool pocket_base = false)
{
    Plan& p = h->plan; const dp_config& c = h->cfg; DeviceWeights& W = h->w;
    const int S = c.inv_sublayers, G = c.n_layers * S;
    const bool fp32_layout = h->precision == DP_FP32 || h->precision == DP_TF32 || !(h->tc_mask & 1);   // fp32 pq, 64-edge units
    const int unit = fp32_layout ? UNIT_F32 : UNIT_TC;
    int rc = 0;
    // Fork: the radius graph (needs x only) runs on a side branch while the main branch encodes and projects the node
    // features (need h only); they join before the first edge kernel.  Works eagerly and under stream capture
    // (the side stream joins the capture through the event).  Profiling spans need one stream: no fork then.
    // In the sampler (pocket_base) the coordinates were already written by the previous DDPM update, so the branch
    // starts BEFORE the encoder; a stand-alone evaluation gets them from the encoder and forks after it.
    const bool fork = (h->profile == 0 || h->profile == 2) && !(h->skip_mask & 16);
    const bool early = fork && pocket_base && !(h->dbg & 32);             // dbg bit 5: fork after the encoder (A/B)
    if (early) {
        DP_CUDA(cudaEventRecord(h->ev_fork, st));
        DP_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        if ((rc = launch_build_edges(h, p.x_in, h->side_stream))) return rc;
        DP_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
    }
    if (!(h->skip_mask & 32) && (rc = launch_encode_nodes(h, xh_phar, xh_res, t_base, step_idx, row_stride, t_stride, pocket_base ? 2 : 0, st))) return rc;
    if (early) {
    } else if (fork) {
        DP_CUDA(cudaEventRecord(h->ev_fork, st));
        DP_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        if ((rc = launch_build_edges(h, p.x_in, h->side_stream))) return rc;
        DP_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
    } else if (!(h->skip_mask & 16) && (rc = launch_build_edges(h, p.x_in, st))) return rc;
    float* x_cur = p.x_a; float* x_next = p.x_b;
    // update_coords_mask (dynamics.py:104-107): the phar rows [0, Np), or every row in joint mode.  The CSR puts the moving
    // rows' edges first, so the coordinate MLP runs on counts[1] = E_p edges (conditional) or on all counts[0] = E (joint).
    const int n_moving = h->joint ? p.N : p.Np;
    const int* n_coord_edges = h->joint ? p.counts : p.counts + 1;

    auto project = [&](int v) -> int {
        const ProjSet& ps = W.proj[v];
        if (ps.lin.out == 0) return DP_OK;
        LinearArgs a{};
        a.x = p.h; a.ldx = H; a.two_source = 0; a.n_rows = p.N; a.K = H;
        a.wt = ps.lin.wt; a.bias = ps.lin.b; a.n_out = ps.lin.out; a.y = p.pq; a.ldy = ps.lin.out; a.epi = 0;
        return run_linear(h, a, 4 * G + c.n_layers + v, st);
    };
    AggView av{};
    av.agg = p.agg; av.partials = p.partials; av.rowptr = p.rowptr; av.unit = unit;
    av.src = unit == UNIT_TC ? p.agg_src : nullptr;
    av.norm = c.normalization_factor; av.inv_norm = 1.0f / c.normalization_factor; av.mean = c.aggregation_mean;

    const bool fused_node = !fp32_layout && (h->tc_mask & 2);
    auto node_phase = [&](int v) -> int {      // tcgen05: node MLP of GCL v-1 + projection of the new h, one launch
        if (h->skip_mask & 2) return DP_OK;
        prof_begin(h, PROF_NODE, st);
        int e = launch_node_tc(h, v, av, st);
        prof_end(h, st);
        return e;
    };
    if ((rc = fused_node ? node_phase(0) : project(0))) return rc;
    if (fork) DP_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    for (int i = 0; i < G; ++i) {
        const GclWeights& L = W.gcl[i];
        const ProjSet& pin = W.proj[i];
        EdgeArgs e{};
        e.p = p.pq; e.ldp = pin.lin.out; e.off_a = pin.off_gcl; e.off_b = pin.off_gcl + H;
        e.wr = L.wr; e.wd = L.wd; e.w2t = L.e2.wt; e.b2 = L.e2.b; e.wv = L.wa; e.bv = L.ba;
        e.x = x_cur; e.d0 = p.d0; e.erow = p.erow; e.ecol = p.col; e.rowptr = p.rowptr;
        e.edst = p.edst; e.n_moving = n_moving; e.ecap = (int)p.Ecap; e.tma_fill = h->tma_fill; e.dbg = h->dbg; e.contig = p.seg_lanes;
        e.n_edges = p.counts; e.agg = p.agg; e.partials = p.partials; e.escal = nullptr;
        e.coord = 0; e.attention = c.attention; e.use_tanh = c.use_tanh; e.trace = h->trace_kernel == 2 ? h->trace : nullptr;
        e.range_flag = p.nan_flag + 2;
        if ((rc = run_edge(h, e, 4 * i + 0, st))) return rc;
        // node model: h <- h + W4 silu(W3 [h | agg] + b3) + b4  (egnn_new.py:54-57)
        if (fused_node) {
            if ((rc = node_phase(i + 1))) return rc;
        } else {
        LinearArgs n0{};
        n0.x = p.h; n0.ldx = H; n0.two_source = 1; n0.aggv = av; n0.n_rows = p.N; n0.K = 2 * H;
        n0.wt = L.n0.wt; n0.bias = L.n0.b; n0.n_out = H; n0.y = p.tbuf; n0.ldy = H; n0.epi = 1;
        if ((rc = run_linear(h, n0, 4 * i + 1, st))) return rc;
        LinearArgs n2{};
        n2.x = p.tbuf; n2.ldx = H; n2.two_source = 0; n2.n_rows = p.N; n2.K = H;
        n2.wt = L.n2.wt; n2.bias = L.n2.b; n2.n_out = H; n2.y = p.h; n2.ldy = H; n2.resid = p.h; n2.ldr = H; n2.epi = 2;
        if ((rc = run_linear(h, n2, 4 * i + 2, st))) return rc;
        if ((rc = project(i + 1))) return rc;
        }
        if ((i + 1) % S == 0) {
            const int b = i / S;
            const CoordWeights& Cw = W.coord[b];
            const ProjSet& pc = W.proj[i + 1];
            EdgeArgs q{};
            q.p = p.pq; q.ldp = pc.lin.out; q.off_a = pc.off_coord; q.off_b = pc.off_coord + H;
            q.wr = Cw.wr; q.wd = Cw.wd; q.w2t = Cw.c2.wt; q.b2 = Cw.c2.b; q.wv = Cw.w4; q.bv = 0.f;
            q.x = x_cur; q.d0 = p.d0; q.erow = p.erow; q.ecol = p.col; q.rowptr = p.rowptr;
            q.edst = p.edst; q.n_moving = n_moving; q.ecap = (int)p.Ecap; q.tma_fill = h->tma_fill; q.dbg = h->dbg; q.contig = p.seg_lanes;
            q.n_edges = n_coord_edges; q.agg = nullptr; q.partials = nullptr; q.escal = p.escal;
            q.coord = 1; q.attention = 0; q.use_tanh = c.use_tanh; q.trace = nullptr; q.range_flag = p.nan_flag + 2;
            // tcgen05 modes: the kernel finishes the rows itself (x_next); the FFMA / TF32 kernels leave escal for coord_finish
            const bool fused_finish = !fp32_layout && h->coord_fused && !(h->skip_mask & (4 | 8));
            q.x_next = fused_finish ? x_next : nullptr; q.cpart = p.cpart; q.cticket = p.cticket;
            q.norm_constant = c.norm_constant; q.coords_range = c.coords_range; q.norm_factor = c.normalization_factor; q.mean = c.aggregation_mean;
            if ((rc = run_edge(h, q, 4 * G + b, st))) return rc;
            // (a coordinate-mode kernel with row-owned tiles that finishes its phar rows itself — no second launch — was built,
            //  parity-green, and measured 5 % SLOWER per step; commit 5e2a79c, profiles/r05e_ab_summary.txt)
            if (!fused_finish && !(h->skip_mask & 8)) {
                prof_begin(h, PROF_EDGE_COORD, st);
                rc = launch_coord_finish(h, x_cur, x_next, n_moving, st);
                prof_end(h, st);
                if (rc) return rc;
            }
            float* t = x_cur; x_cur = x_next; x_next = t;
        }
    }
    if (h->skip_mask & 32) return DP_OK;
    return launch_decode(h, x_cur, out_phar, out_res, st);
}

extern "C" int dp_dynamics_forward(dp_handle* h, const float* xh_phar, const float* xh_res, const float* t_dev,
                                   int32_t t_stride, float* out_phar, float* out_res, void* stream)
{
    int rc = require(h, true, true);
    if (rc) return rc;
    DP_CHECK(xh_phar && xh_res && out_phar, DP_ERR_INVALID, "dp_dynamics_forward: null tensor");
    DP_CHECK(t_dev || !h->cfg.condition_time, DP_ERR_INVALID, "dp_dynamics_forward: t is required");
    DP_CHECK(t_stride == 0 || t_stride == 1, DP_ERR_INVALID, "t_stride must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    DP_CHECK(!h->joint || out_res, DP_ERR_INVALID, "dp_dynamics_forward: joint mode returns the pocket velocities, out_res_dev is required");
    if ((rc = run_denoiser(h, xh_phar, xh_res, t_dev ? t_dev : h->plan.t_const, nullptr, 0, t_stride, out_phar, out_res, st))) return rc;
    if ((rc = launch_nan_fixup(h, out_phar, out_res, st))) return rc;
    return h->joint ? launch_velocity_center(h, out_phar, out_res, st) : DP_OK;     // dynamics.py:133-136
}

// --------------------------------------------------------------------------------------
// DDPM update + sampler loop
// --------------------------------------------------------------------------------------
extern "C" int dp_ddpm_update(dp_handle* h, int32_t kind, float a, float c, float sigma, float* z, float* pocket,
                              const float* eps_hat, const float* noise, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    DP_CHECK(kind >= 0 && kind <= 2, DP_ERR_INVALID, "dp_ddpm_update: kind %d", kind);
    DP_CHECK(z && pocket && noise && (eps_hat || kind == 2), DP_ERR_INVALID, "dp_ddpm_update: null tensor");
    DdpmArgs d{};
    d.kind = kind; d.a = a; d.c = c; d.sigma = sigma; d.table = nullptr; d.step_idx = nullptr;
    d.z = z; d.pocket = pocket; d.eps_hat = eps_hat; d.noise = noise; d.stat_index = -1; d.advance = 0;
    // the standalone call never saw the denoiser's NaN flag of another stream; clear semantics: caller's eps_hat
    // already went through dp_dynamics_forward's fix-up, so the flag must not zero it again
    DP_CUDA(cudaMemsetAsync(h->plan.nan_flag, 0, sizeof(int), (cudaStream_t)stream));
    return launch_ddpm(h, d, (cudaStream_t)stream);
}

struct FrameSpec {                      // return_frames > 1 (conditional_model.py:439-442): where the DDPM update drops its frames
    int return_frames = 0;              // 0: none
    float* phar = nullptr; float* pocket = nullptr;
    float norm_x = 1.f, norm_h = 1.f, bias_h = 0.f;
};

static int sampler_step_launches(dp_handle* h, float* pocket, const float* noise, const FrameSpec& fs, cudaStream_t st)
{
    Plan& p = h->plan; const dp_config& c = h->cfg;
    const int PW = 3 + c.phar_nf;
    int rc = run_denoiser(h, p.z, pocket, p.step_rows, p.step_idx, 4, 0, p.eps_hat, nullptr, st, true);
    if (rc) return rc;
    DdpmArgs d{};
    d.kind = 0; d.table = p.step_rows; d.step_idx = p.step_idx;
    d.z = p.z; d.pocket = pocket; d.eps_hat = p.eps_hat; d.noise = noise;
    d.noise_step_stride = (int64_t)p.Np * PW; d.noise_step_base = 1;
    d.stat_base = 0; d.stat_index = -1; d.advance = 1;
    d.x_in = p.x_in; d.x_a = p.x_a; d.x_b = p.x_b;                        // the next evaluation's coordinates
    if (fs.return_frames > 1) {
        d.frames_phar = fs.phar; d.frames_pocket = fs.pocket; d.return_frames = fs.return_frames; d.n_steps = h->n_steps;
        d.norm_x = fs.norm_x; d.norm_h = fs.norm_h; d.bias_h = fs.bias_h;
    }
    return launch_ddpm(h, d, st);
}

// The loop of sample_given_pocket on handle-owned buffers only (plan.pocket, handle noise / frame buffers): the
// captured step graph therefore never bakes a caller pointer and survives any number of calls with fresh tensors.
static int sample_core(dp_handle* h, const FrameSpec& fs, cudaStream_t st)
{
    Plan& p = h->plan; const dp_config& c = h->cfg;
    const int PW = 3 + c.phar_nf;
    float* pocket = p.pocket;
    const float* noise = reinterpret_cast<const float*>(h->noise.p);
    int rc = 0;
    DP_CUDA(cudaMemsetAsync(p.step_idx, 0, sizeof(int), st));
    DP_CUDA(cudaMemsetAsync(p.stats, 0, (size_t)p.stats_cap * 2 * sizeof(float), st));
    DP_CUDA(cudaMemsetAsync(p.nan_flag, 0, sizeof(int), st));
    // the pocket's type features never change while sampling: embed them once, without the time term
    if ((rc = launch_encode_nodes(h, p.z, pocket, p.t_const, nullptr, 0, 0, 1, st))) return rc;
    // z_T ~ N(pocket COM, I), projected (conditional_model.py:412-418)
    if ((rc = launch_pocket_com_init(h, p.z, pocket, st))) return rc;
    DdpmArgs d0{};
    d0.kind = 2; d0.sigma = 1.0f; d0.z = p.z; d0.pocket = pocket; d0.eps_hat = nullptr; d0.noise = noise;
    d0.stat_index = -1; d0.advance = 0;
    d0.x_in = p.x_in; d0.x_a = p.x_a; d0.x_b = p.x_b;
    if ((rc = launch_ddpm(h, d0, st))) return rc;

    if (h->profile == 1 || h->profile == 3) {
        for (int k = 0; k < h->n_steps; ++k)
            if ((rc = sampler_step_launches(h, pocket, noise, fs, st))) return rc;
    } else {
        const int want_frames = fs.return_frames > 1 ? fs.return_frames : 0;
        if (h->profile == 2) {                                    // the instrumented graph is captured afresh (and dropped afterwards)
            drop_graph(h);
            for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
            h->spans.clear();
        }
        if (!h->step_graph || h->graph_precision != h->precision || h->graph_frames != want_frames) {
            drop_graph(h);
            const int64_t before = h->launches;
            cudaGraph_t g = nullptr;
            // the legacy default stream cannot be captured: record on a handle-owned stream and
            // replay the instantiated graph on the caller's stream
            if (!h->capture_stream) DP_CUDA(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
            DP_CUDA(cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
            h->capturing = true;
            rc = sampler_step_launches(h, pocket, noise, fs, h->capture_stream);
            h->capturing = false;
            cudaError_t ce = cudaStreamEndCapture(h->capture_stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            DP_CUDA(ce);
            h->graph_launches = h->launches - before;
            h->launches = before;
            DP_CUDA(cudaGraphInstantiate(&h->step_graph, g, 0));
            cudaGraphDestroy(g);
            h->graph_precision = h->precision; h->graph_frames = want_frames;
            h->graph_captures += 1;
        }
        for (int k = 0; k < h->n_steps; ++k) {
            DP_CUDA(cudaGraphLaunch(h->step_graph, st));
            if (h->profile == 2 && (rc = prof_collect_replay(h, st))) return rc;
        }
        h->launches += h->graph_launches * h->n_steps;
        if (h->profile == 2) drop_graph(h);
    }
    // p(x | z0): conditional_model.py:108-131
    DP_CUDA(cudaMemcpyAsync(p.t_const, &h->final_host[0], sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = run_denoiser(h, p.z, pocket, p.t_const, nullptr, 0, 0, p.eps_hat, nullptr, st, true))) return rc;
    DP_CUDA(cudaMemcpyAsync(p.out_buf, p.z, (size_t)p.Np * PW * sizeof(float), cudaMemcpyDeviceToDevice, st));   // keeps z0's feature columns
    DdpmArgs df{};
    df.kind = 1; df.a = h->final_host[1]; df.c = h->final_host[2]; df.sigma = h->final_host[3];
    df.z = p.z; df.pocket = pocket; df.eps_hat = p.eps_hat;
    df.noise = noise + (size_t)(h->n_steps + 1) * p.Np * PW; df.stat_index = h->n_steps; df.advance = 0;
    if ((rc = launch_ddpm(h, df, st))) return rc;
    DP_CUDA(cudaMemcpy2DAsync(p.out_buf, PW * sizeof(float), p.z, PW * sizeof(float), 3 * sizeof(float), p.Np,
                              cudaMemcpyDeviceToDevice, st));
    return DP_OK;
}

// noise / frame buffers of the current (plan, step table); a moved buffer invalidates the captured graph
static int reserve_run_buffers(dp_handle* h, int return_frames)
{
    Plan& p = h->plan; const dp_config& c = h->cfg;
    const size_t PW = 3 + c.phar_nf, RW = 3 + c.residue_nf;
    bool moved = false;
    int rc = h->noise.reserve((size_t)(h->n_steps + 2) * p.Np * PW * sizeof(float), &moved);
    if (rc) return rc;
    if (moved) drop_graph(h);
    if (return_frames > 1) {
        rc = h->frames.reserve((size_t)return_frames * ((size_t)p.Np * PW + (size_t)p.Nr * RW) * sizeof(float), &moved);
        if (rc) return rc;
        if (moved) drop_graph(h);
    }
    return DP_OK;
}

extern "C" int dp_fill_noise(dp_handle* h, uint64_t seed, const int64_t* sample_ids_host, int32_t n_draws, float* noise_dev, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    DP_CHECK(n_draws > 0 && noise_dev, DP_ERR_INVALID, "dp_fill_noise: bad argument");
    Plan& p = h->plan;
    cudaStream_t st = (cudaStream_t)stream;
    if (sample_ids_host) {
        DP_CUDA(cudaMemcpyAsync(p.sample_ids, sample_ids_host, (size_t)p.B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        DP_CUDA(cudaStreamSynchronize(st));                // the host array may be a temporary of the caller
    }
    return launch_fill_noise(h, seed, n_draws, noise_dev, st);
}

extern "C" int dp_sample_ex(dp_handle* h, float* pocket, const dp_sample_opts* o, float* out_phar, void* stream)
{
    int rc = require(h, true, true);
    if (rc) return rc;
    DP_CHECK(h->n_steps > 0 && h->plan.step_rows, DP_ERR_STATE, "dp_set_step_table has not been called");
    DP_CHECK(!h->joint, DP_ERR_STATE, "the pocket-conditioned sampler needs update_pocket_coords = False (conditional_model.py:18)");
    DP_CHECK(pocket && o && out_phar, DP_ERR_INVALID, "dp_sample_ex: null argument");
    const int F = o->return_frames > 1 ? o->return_frames : 1;
    DP_CHECK(F <= h->n_steps && h->n_steps % F == 0, DP_ERR_INVALID,
             "return_frames %d must divide the %d sampling steps", F, h->n_steps);   // conditional_model.py:395-396
    DP_CHECK(F == 1 || (o->frames_phar_dev && o->frames_pocket_dev), DP_ERR_INVALID, "dp_sample_ex: frame buffers missing");
    Plan& p = h->plan; const dp_config& c = h->cfg;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t PW = 3 + c.phar_nf, RW = 3 + c.residue_nf;
    if ((rc = reserve_run_buffers(h, F))) return rc;
    const size_t noise_count = (size_t)(h->n_steps + 2) * p.Np * PW;
    if (o->noise_dev) {
        if (o->noise_dev != h->noise.p)
            DP_CUDA(cudaMemcpyAsync(h->noise.p, o->noise_dev, noise_count * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if ((rc = dp_fill_noise(h, o->seed, o->sample_ids_host, h->n_steps + 2, reinterpret_cast<float*>(h->noise.p), stream))) {
        return rc;
    }
    if (pocket != p.pocket)
        DP_CUDA(cudaMemcpyAsync(p.pocket, pocket, (size_t)p.Nr * RW * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FrameSpec fs;
    if (F > 1) {
        fs.return_frames = F; fs.norm_x = o->norm_x; fs.norm_h = o->norm_h; fs.bias_h = o->bias_h;
        fs.phar = reinterpret_cast<float*>(h->frames.p);
        fs.pocket = fs.phar + (size_t)F * p.Np * PW;
    }
    if ((rc = sample_core(h, fs, st))) return rc;
    if (pocket != p.pocket)
        DP_CUDA(cudaMemcpyAsync(pocket, p.pocket, (size_t)p.Nr * RW * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out_phar != p.out_buf)
        DP_CUDA(cudaMemcpyAsync(out_phar, p.out_buf, (size_t)p.Np * PW * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (F > 1) {
        DP_CUDA(cudaMemcpyAsync(o->frames_phar_dev, fs.phar, (size_t)F * p.Np * PW * sizeof(float), cudaMemcpyDeviceToDevice, st));
        DP_CUDA(cudaMemcpyAsync(o->frames_pocket_dev, fs.pocket, (size_t)F * p.Nr * RW * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return DP_OK;
}

extern "C" int dp_sample(dp_handle* h, float* pocket, const float* noise, float* out_phar, void* stream)
{
    DP_CHECK(noise, DP_ERR_INVALID, "dp_sample: null noise (dp_sample_ex draws counter-based noise on the device)");
    dp_sample_opts o{};
    o.noise_dev = noise; o.return_frames = 1;
    return dp_sample_ex(h, pocket, &o, out_phar, stream);
}

static int sample_host_common(dp_handle* h, const float* pocket_host, const float* noise_host, uint64_t seed,
                              const int64_t* sample_ids_host, float* out_phar_host, float* pocket_out_host)
{
    int rc = require(h, true, true);
    if (rc) return rc;
    DP_CHECK(pocket_host && out_phar_host, DP_ERR_INVALID, "dp_sample_host: null buffer");
    DP_CHECK(!h->joint, DP_ERR_STATE, "the pocket-conditioned sampler needs update_pocket_coords = False (conditional_model.py:18)");
    DP_CHECK(h->n_steps > 0 && h->plan.step_rows, DP_ERR_STATE, "dp_set_step_table has not been called");
    Plan& p = h->plan; const dp_config& c = h->cfg;
    const int PW = 3 + c.phar_nf, RW = 3 + c.residue_nf;
    if ((rc = reserve_run_buffers(h, 1))) return rc;
    const size_t noise_count = (size_t)(h->n_steps + 2) * p.Np * PW;
    cudaStream_t st = 0;
    DP_CUDA(cudaMemcpyAsync(p.pocket, pocket_host, (size_t)p.Nr * RW * sizeof(float), cudaMemcpyHostToDevice, st));
    dp_sample_opts o{};
    o.return_frames = 1;
    if (noise_host) {
        DP_CUDA(cudaMemcpyAsync(h->noise.p, noise_host, noise_count * sizeof(float), cudaMemcpyHostToDevice, st));
        o.noise_dev = reinterpret_cast<const float*>(h->noise.p);
    } else {
        o.seed = seed; o.sample_ids_host = sample_ids_host;
    }
    if ((rc = dp_sample_ex(h, p.pocket, &o, p.out_buf, st))) return rc;
    DP_CUDA(cudaMemcpyAsync(out_phar_host, p.out_buf, (size_t)p.Np * PW * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (pocket_out_host)
        DP_CUDA(cudaMemcpyAsync(pocket_out_host, p.pocket, (size_t)p.Nr * RW * sizeof(float), cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaStreamSynchronize(st));
    return DP_OK;
}

extern "C" int dp_sample_host(dp_handle* h, const float* pocket_host, const float* noise_host, float* out_phar_host,
                              float* pocket_out_host)
{
    DP_CHECK(noise_host, DP_ERR_INVALID, "dp_sample_host: null noise (dp_sample_host_seeded draws it on the device)");
    return sample_host_common(h, pocket_host, noise_host, 0, nullptr, out_phar_host, pocket_out_host);
}

extern "C" int dp_sample_host_seeded(dp_handle* h, const float* pocket_host, uint64_t seed, const int64_t* sample_ids_host,
                                     float* out_phar_host, float* pocket_out_host)
{
    return sample_host_common(h, pocket_host, nullptr, seed, sample_ids_host, out_phar_host, pocket_out_host);
}

extern "C" int64_t dp_graph_captures(const dp_handle* h) { return h ? h->graph_captures : 0; }

// --------------------------------------------------------------------------------------
// flags / counters / profiling
// --------------------------------------------------------------------------------------
extern "C" int dp_get_flags(dp_handle* h, dp_flags* out, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    DP_CHECK(out, DP_ERR_INVALID, "dp_get_flags: null out");
    Plan& p = h->plan;
    cudaStream_t st = (cudaStream_t)stream;
    int counts[4], nan[4];
    DP_CUDA(cudaMemcpyAsync(counts, p.counts, sizeof(counts), cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaMemcpyAsync(nan, p.nan_flag, sizeof(nan), cudaMemcpyDeviceToHost, st));
    std::vector<float> stats((size_t)p.stats_cap * 2, 0.f);
    if (p.stats) DP_CUDA(cudaMemcpyAsync(stats.data(), p.stats, stats.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    out->nan_resets = nan[1];
    out->f16_range = nan[2];
    out->edge_overflow = counts[2];
    out->last_n_edges = counts[0];
    out->last_n_edges_phar = counts[1];
    float worst = 0.f;
    for (int k = 0; k < p.stats_cap; ++k) {
        const float rel = stats[2 * k] / (stats[2 * k + 1] + 1e-10f);
        if (rel > worst) worst = rel;
    }
    out->max_mean_rel_err = worst;
    out->last_max_cog = p.stats_cap ? stats[2 * (size_t)(p.stats_cap - 2)] : 0.f;
    return DP_OK;
}

extern "C" int dp_reset_flags(dp_handle* h, void* stream)
{
    int rc = require(h, false, true);
    if (rc) return rc;
    Plan& p = h->plan;
    cudaStream_t st = (cudaStream_t)stream;
    DP_CUDA(cudaMemsetAsync(p.counts + 2, 0, sizeof(int), st));
    DP_CUDA(cudaMemsetAsync(p.nan_flag, 0, 4 * sizeof(int), st));
    if (p.stats) DP_CUDA(cudaMemsetAsync(p.stats, 0, (size_t)p.stats_cap * 2 * sizeof(float), st));
    return DP_OK;
}

// debug: copies the clock64 timeline the last traced kernel wrote (DIFFPHAR_TRACE=1); not part of the stable ABI
extern "C" int dp_debug_trace(dp_handle* h, long long* out_host, int32_t n_words)
{
    DP_CHECK(h && out_host && n_words > 0 && n_words <= DP_TRACE_WORDS, DP_ERR_INVALID, "dp_debug_trace: bad argument");
    DP_CHECK(h->trace, DP_ERR_STATE, "tracing is off (set DIFFPHAR_TRACE=1 before dp_create)");
    DP_CUDA(cudaSetDevice(h->device));
    DP_CUDA(cudaDeviceSynchronize());
    DP_CUDA(cudaMemcpy(out_host, h->trace, (size_t)n_words * sizeof(long long), cudaMemcpyDeviceToHost));
    return DP_OK;
}

extern "C" int64_t dp_launch_count(const dp_handle* h) { return h ? h->launches : 0; }

extern "C" int dp_profile_enable(dp_handle* h, int32_t on)
{
    DP_CHECK(h, DP_ERR_INVALID, "null handle");
    h->profile = (on == 2 || on == 3) ? on : (on != 0 ? 1 : 0);
    if (!on) { for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); } h->spans.clear(); }
    if (on) {
        for (auto& s : h->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
        h->spans.clear();
        for (int i = 0; i < 8; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
    }
    return DP_OK;
}

extern "C" int dp_profile_read(dp_handle* h, int32_t which, double* total_ms, int64_t* launches)
{
    DP_CHECK(h && which >= 0 && which < 8, DP_ERR_INVALID, "dp_profile_read: bad argument");
    DP_CUDA(cudaSetDevice(h->device));
    if (!h->spans.empty() && h->profile != 2) {
        DP_CUDA(cudaDeviceSynchronize());
        for (auto& s : h->spans) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { h->prof_ms[s.which] += ms; h->prof_n[s.which] += 1; }
            cudaEventDestroy(s.a); cudaEventDestroy(s.b);
        }
        cudaGetLastError();
        h->spans.clear();
    }
    if (total_ms) *total_ms = h->prof_ms[which];
    if (launches) *launches = h->prof_n[which];
    return DP_OK;
}
