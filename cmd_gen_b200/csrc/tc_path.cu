// tcgen05 tensor-core path (DP_BF16 / DP_F16): the dense contractions of the EGNN layers on the
// 5th-generation tensor cores with fp32 accumulators in TMEM.
//
// Orientation ("channels on lanes"): every contraction is computed TRANSPOSED,
//      D[out channel, item] = sum_k W[out channel, k] * In[item, k]          (item = edge or node)
// so A = the nn.Linear weight exactly as stored ([out,in] row-major == K-major) and B = the tile of
// activations (K-major).  The accumulator then has the OUTPUT CHANNEL on the TMEM lane and the
// items along TMEM columns, which makes the epilogue cheap where it matters:
//   * the CSR segmented sum over a row's edges is a run of register adds inside one thread
//     (thread = channel, registers = consecutive edges) — no atomics, no shuffles;
//   * agg / h stores are 128 B coalesced (32 lanes = 32 consecutive channels of one node);
//   * bias and gate weights are per-thread scalars.
// The only cross-lane step is the 256-channel dot product of the attention gate / coordinate
// scalar: a 31-shuffle transposing butterfly per 32x32 block + one 8-warp named barrier.
//
// Operands live in shared memory in the canonical K-major SWIZZLE_128B layout (8 rows x 128 B
// atoms, 16-byte chunk index XOR row%8).  Weights are pre-swizzled on the host into exactly that
// image and arrive with cp.async.bulk (TMA engine, UBLKCP) + mbarrier complete_tx; activations are
// produced by the CUDA cores (gather + first-layer SiLU) straight into the swizzled tile, then
// fence.proxy.async hands them to the tensor core.  One elected thread issues tcgen05.mma;
// tcgen05.commit arrives on the mbarrier the epilogue warps wait on.
#include "tc_common.cuh"

#include <cstring>

namespace {
using namespace tc;

constexpr int TILE = 64;                 // nodes per MMA tile  (UMMA N)
constexpr int X_PANEL_BYTES = TILE * 128;
constexpr int THREADS = 512;

__device__ __forceinline__ float silu_fast(float v) { return v * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v)); }

// write 8 consecutive K elements (one 16-byte chunk) of item `i` into a swizzled tile
__device__ __forceinline__ void store_chunk(unsigned char* tile, int i, int chunk, uint4 v)
{
    *reinterpret_cast<uint4*>(tile + chunk_offset(i, chunk, X_PANEL_BYTES)) = v;
}

// issue the MMAs of one tile: D[256 x 64] (+)= W[256 x 64*n_panels] * X[64 x 64*n_panels]^T
// w_base / x_base: shared addresses of panel 0; D columns: [d_col, d_col+64) rows 0..127, [d_col+64, +128) rows 128..255
__device__ __forceinline__ void issue_tile_mma(uint32_t tmem_d, uint32_t w_base, uint32_t x_base, int n_panels,
                                               uint32_t idesc, bool accumulate_first)
{
#pragma unroll 1
    for (int kp = 0; kp < n_panels; ++kp) {
#pragma unroll
        for (int ks = 0; ks < PANEL_K / 16; ++ks) {
            const uint64_t bdesc = make_desc(x_base + kp * X_PANEL_BYTES + ks * 32);
            const uint32_t acc = (accumulate_first || kp > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint64_t adesc = make_desc(w_base + kp * W_PANEL_BYTES + hh * (128 * 128) + ks * 32);
                umma_f16(tmem_d + hh * TILE, adesc, bdesc, idesc, acc);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-node linear: y[n, col0 + c] = epi( sum_k x[n, k] W[col0 + c, k] + b )
// grid = (node tiles of 64, out blocks of 256); weights streamed in K=128 chunks (2 buffers)
// ------------------------------------------------------------------------------------------
struct LinTcSmem {
    unsigned char w[2][2 * W_PANEL_BYTES];            // 2 x 64 KB
    unsigned char x[8 * X_PANEL_BYTES];               // up to K = 512: 64 KB
    unsigned long long bar_full[2];
    unsigned long long bar_empty[2];
    unsigned long long bar_mma;
    uint32_t tmem_holder;
};

template <int FMT>
__global__ void __launch_bounds__(THREADS, 1) linear_tc_kernel(LinearArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    LinTcSmem& s = *reinterpret_cast<LinTcSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = blockIdx.x * TILE, ob = blockIdx.y;
    const int n_panels = a.K / PANEL_K, n_chunks = n_panels / 2;
    const unsigned char* img = w_img + (size_t)ob * n_panels * W_PANEL_BYTES;
    const uint32_t bf0 = smem_u32(&s.bar_full[0]), bf1 = smem_u32(&s.bar_full[1]);
    const uint32_t be0 = smem_u32(&s.bar_empty[0]), be1 = smem_u32(&s.bar_empty[1]);
    const uint32_t bm = smem_u32(&s.bar_mma);
    if (tid == 0) {
        mbar_init(bf0, 1); mbar_init(bf1, 1); mbar_init(be0, 1); mbar_init(be1, 1); mbar_init(bm, 1);
        fence_barrier_init();
    }
    if (wid == 1) tmem_alloc(smem_u32(&s.tmem_holder), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_holder;
    if (tid == 0) {
        for (int c = 0; c < 2 && c < n_chunks; ++c) {
            mbar_expect_tx(c ? bf1 : bf0, 2 * W_PANEL_BYTES);
            bulk_g2s(smem_u32(s.w[c]), img + (size_t)c * 2 * W_PANEL_BYTES, 2 * W_PANEL_BYTES, c ? bf1 : bf0);
        }
    }
    // ---- activations -> swizzled tile (fp32 rows, converted on the fly; [h | agg] for the node model)
    {
        const int chunks = a.K / 8;
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * wid + u, row = n0 + i;
            for (int c = lane; c < chunks; c += 32) {
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (row < a.n_rows) {
                    const int k = 8 * c;
                    float4 f0, f1;
                    if (a.two_source && k >= H) { f0 = agg_load4(a.aggv, row, k - H); f1 = agg_load4(a.aggv, row, k - H + 4); }
                    else {
                        const float* src = a.x + (size_t)row * a.ldx + k;
                        f0 = *reinterpret_cast<const float4*>(src); f1 = *reinterpret_cast<const float4*>(src + 4);
                    }
                    o = make_uint4(pack2<FMT>(f0.x, f0.y), pack2<FMT>(f0.z, f0.w), pack2<FMT>(f1.x, f1.y), pack2<FMT>(f1.z, f1.w));
                }
                store_chunk(s.x, i, c, o);
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc(FMT, 128, TILE);
        for (int c = 0; c < n_chunks; ++c) {
            const int b = c & 1;
            mbar_wait(b ? bf1 : bf0, (c >> 1) & 1);
            tc_fence_after();
            issue_tile_mma(tmem_base, smem_u32(s.w[b]), smem_u32(s.x) + c * 2 * X_PANEL_BYTES, 2, idesc, c > 0);
            umma_commit(b ? be1 : be0);
            if (c >= 1 && c + 1 < n_chunks) {            // refill the other buffer once its MMAs retired
                const int pb = (c - 1) & 1;
                mbar_wait(pb ? be1 : be0, ((c - 1) >> 1) & 1);
                mbar_expect_tx(pb ? bf1 : bf0, 2 * W_PANEL_BYTES);
                bulk_g2s(smem_u32(s.w[pb]), img + (size_t)(c + 1) * 2 * W_PANEL_BYTES, 2 * W_PANEL_BYTES, pb ? bf1 : bf0);
            }
        }
        umma_commit(bm);
    }
    // ---- epilogue: thread = out channel, registers = 32 nodes
    const int q = wid & 3, g = wid >> 2, half = g & 1, ng = g >> 1;
    const int ch = 128 * half + 32 * q + lane;
    const int col = ob * 256 + ch;
    const float bias = a.bias ? a.bias[col] : 0.f;
    mbar_wait(bm, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + half * TILE + ng * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int row = n0 + 32 * ng + j;
        if (row < a.n_rows) {
            float o = v[j] + bias;
            if (a.epi == 1) o = silu_fast(o);
            else if (a.epi == 2) o += a.resid[(size_t)row * a.ldr + col];
            a.y[(size_t)row * a.ldy + col] = o;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------
// host: weight images
// ------------------------------------------------------------------------------------------
uint16_t f32_to_bf16(float f)
{
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
uint16_t f32_to_f16(float f)
{
    __half h = __float2half_rn(f);
    uint16_t r; memcpy(&r, &h, 2);
    return r;
}

}  // namespace

int tc_init()
{
    int rc = tc_edge_init();
    if (!rc) rc = tc_node_init();
    if (rc) return rc;
    DP_CUDA(cudaFuncSetAttribute(linear_tc_kernel<FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinTcSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(linear_tc_kernel<FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinTcSmem) + 1024));
    return DP_OK;
}

void tc_free_weights(dp_handle* h)
{
    if (!h->tc) return;
    for (void* p : h->tc->allocations) cudaFree(p);
    delete h->tc;
    h->tc = nullptr;
}

// Packs every registered linear (k-major fp32 host copy, out channels padded to blocks of 256) into
// the swizzled smem image, once per 16-bit format.
int tc_prepare_weights(dp_handle* h)
{
    tc_free_weights(h);
    h->tc = new TcWeights();
    TcWeights& T = *h->tc;
    T.lin.resize(h->tc_host.size());
    for (size_t id = 0; id < h->tc_host.size(); ++id) {
        const HostLinear& L = h->tc_host[id];
        if (L.n_out == 0) continue;
        DP_CHECK(L.K % 128 == 0 && L.n_out % 256 == 0, DP_ERR_INVALID, "tc linear %zu: K=%d n_out=%d not tileable", id, L.K, L.n_out);
        const int n_panels = L.K / PANEL_K, n_blocks = L.n_out / 256;
        const size_t bytes = (size_t)n_blocks * n_panels * W_PANEL_BYTES;
        std::vector<uint16_t> img(bytes / 2);
        for (int fmt = 0; fmt < 2; ++fmt) {
            for (int o = 0; o < L.n_out; ++o) {
                const int ob = o / 256, r = o % 256;
                for (int k = 0; k < L.K; ++k) {
                    const int kp = k / PANEL_K, kb = (k % PANEL_K) * 2;
                    const size_t off = ((size_t)ob * n_panels + kp) * W_PANEL_BYTES + (size_t)r * 128 +
                                       ((((kb >> 4) ^ (r & 7))) << 4) + (kb & 15);
                    const float w = L.wt[(size_t)k * L.n_out + o];
                    img[off / 2] = fmt == FMT_BF16 ? f32_to_bf16(w) : f32_to_f16(w);
                }
            }
            void* d = nullptr;
            DP_CUDA(cudaMalloc(&d, bytes));
            T.allocations.push_back(d);
            DP_CUDA(cudaMemcpy(d, img.data(), bytes, cudaMemcpyHostToDevice));
            T.lin[id].img[fmt] = reinterpret_cast<unsigned char*>(d);
        }
        T.lin[id].K = L.K; T.lin[id].n_out = L.n_out;
    }
    // Fused node-phase launches (tc_node.cu) stream ONE contiguous panel sequence: for h version v > 0 the
    // node MLP of GCL v-1 (node_mlp.0: 8 panels, node_mlp.2: 4 panels) followed by the projection blocks of
    // the new h; v = 0 is the projection of the embedded features alone.
    const dp_config& c = h->cfg;
    const int G = c.n_layers * c.inv_sublayers;
    T.node.resize(G + 1);
    for (int v = 0; v <= G; ++v) {
        std::vector<int> parts;
        if (v > 0) { parts.push_back(4 * (v - 1) + 1); parts.push_back(4 * (v - 1) + 2); }
        parts.push_back(4 * G + c.n_layers + v);
        size_t bytes = 0;
        for (int id : parts) bytes += (size_t)(T.lin[id].K / PANEL_K) * (T.lin[id].n_out / 256) * W_PANEL_BYTES;
        T.node[v].n_panels = (int)(bytes / W_PANEL_BYTES);
        if (bytes == 0) continue;
        for (int fmt = 0; fmt < 2; ++fmt) {
            void* d = nullptr;
            DP_CUDA(cudaMalloc(&d, bytes));
            T.allocations.push_back(d);
            size_t off = 0;
            for (int id : parts) {
                const size_t b = (size_t)(T.lin[id].K / PANEL_K) * (T.lin[id].n_out / 256) * W_PANEL_BYTES;
                if (b) DP_CUDA(cudaMemcpy(reinterpret_cast<unsigned char*>(d) + off, T.lin[id].img[fmt], b, cudaMemcpyDeviceToDevice));
                off += b;
            }
            T.node[v].img[fmt] = reinterpret_cast<unsigned char*>(d);
        }
    }
    return DP_OK;
}

int tc_fmt_of(dp_handle* h, int* fmt)
{
    if (h->precision == DP_BF16) { *fmt = FMT_BF16; return DP_OK; }
    if (h->precision == DP_F16) { *fmt = FMT_F16; return DP_OK; }
    dp_set_error("precision mode %d: the tcgen05 kind::tf32 path (streamed fp32 weight tiles) is not built yet; "
                 "use DP_FP32, DP_F16 (same 10-bit mantissa as TF32) or DP_BF16", h->precision);
    return DP_ERR_INVALID;
}

int launch_linear_tc(dp_handle* h, const LinearArgs& a, int lin_id, cudaStream_t st)
{
    int fmt = 0, rc = tc_fmt_of(h, &fmt);
    if (rc) return rc;
    DP_CHECK(h->tc && lin_id >= 0 && lin_id < (int)h->tc->lin.size() && h->tc->lin[lin_id].img[fmt], DP_ERR_STATE,
             "tc linear %d has no weight image", lin_id);
    const TcLinearImg& L = h->tc->lin[lin_id];
    DP_CHECK(L.K == a.K && L.n_out == a.n_out, DP_ERR_INVALID, "tc linear %d: shape mismatch", lin_id);
    if (a.n_rows <= 0) return DP_OK;
    dim3 grid((a.n_rows + TILE - 1) / TILE, a.n_out / 256);
    const int smem = (int)sizeof(LinTcSmem) + 1024;
    if (fmt == FMT_BF16) linear_tc_kernel<FMT_BF16><<<grid, THREADS, smem, st>>>(a, L.img[fmt]);
    else linear_tc_kernel<FMT_F16><<<grid, THREADS, smem, st>>>(a, L.img[fmt]);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
