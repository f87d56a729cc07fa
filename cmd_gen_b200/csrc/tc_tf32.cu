// DP_TF32 — the EGNN contractions on tcgen05 kind::tf32 tiles (fp32 storage everywhere, 10-bit-mantissa operands rounded
// to nearest, fp32 accumulation in tensor memory).  north_star's "TF32 path": the reference's arithmetic is fp32
// (constants.py:8-9), TF32 is the tensor-core format closest to it and the yardstick BASELINE config 5 asks the bf16
// tiles to be compared with.  Same data layout and segmented-sum scheme as the FFMA mode (egnn_f32.cu: fp32 `pq`,
// 64-edge units) — only the dense products move to the tensor cores:
//
//   linear_tf32_kernel : y = epi(x W^T + b) per 64-node tile and 256-channel block (node MLP, factored first layers)
//   edge_tf32_kernel   : GCL.edge_model + attention + segmented sum / coordinate scalar per 64-edge tile
//
// Orientation as in tc_edge.cu ("channels on lanes"): A = the nn.Linear weight block [256 out, K] exactly as stored,
// B = the activation tile [64 items, K], D[256 ch, 64 items] in 128 TMEM columns.  Operands are K-major SWIZZLE_128B
// with 32 tf32 elements per 128-byte row; one MMA covers K = 8 (32 bytes).  Weight panels (256 rows x 128 B = 32 KB,
// pre-rounded and pre-swizzled by tc_weights.cu) stream through a two-slot ring filled by cp.async.bulk.
// This is the accuracy mode, not the throughput mode: one tile in flight per CTA, no warp specialisation.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TN = 64;                       // items per tile (UMMA N)
constexpr int KP32 = 32;                     // tf32 elements per swizzle row
constexpr int XP_BYTES = TN * 128;           // one K panel of the activation tile: 8 KB
constexpr int N_WS = 2;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 128;               // two 128-channel halves x 64 items
constexpr int FMT_TF32 = 2;                  // cute::UMMA::F16F32Format::TF32

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float to_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

// byte offset of elements k .. k+3 (k % 4 == 0) of item `i` inside the swizzled activation tile
__device__ __forceinline__ uint32_t x_offset(int i, int k)
{
    return (uint32_t)((k >> 5) * XP_BYTES + i * 128 + ((((k & 31) >> 2) ^ (i & 7)) << 4));
}

struct Pipe {                                // weight ring + accumulator hand-off, shared by both kernels
    unsigned long long bar_wfull[N_WS], bar_wempty[N_WS], bar_acc;
    uint32_t tmem_holder;
};

__device__ __forceinline__ void pipe_init(Pipe& p, int tid, int wid)
{
    if (tid == 0) {
        for (int i = 0; i < N_WS; ++i) { mbar_init(smem_u32(&p.bar_wfull[i]), 1); mbar_init(smem_u32(&p.bar_wempty[i]), 1); }
        mbar_init(smem_u32(&p.bar_acc), 1);
        fence_barrier_init();
    }
    if (wid == 0) tmem_alloc(smem_u32(&p.tmem_holder), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// Warp 0 (whole warp, uniform control flow; one elected lane issues): D[256 x 64] = W[256 x 32 n_panels] X^T.
// `g0` = panels this CTA has consumed so far (slot and mbarrier parity follow from it).
__device__ __forceinline__ void gemm_tile(Pipe& p, unsigned char* w_ring, const unsigned char* __restrict__ w_img, int n_panels,
                                          uint32_t x_base, uint32_t tmem_d, int g0)
{
    constexpr uint32_t idesc = make_idesc(FMT_TF32, 128, TN);
    const uint32_t w0 = warp_uniform(smem_u32(w_ring));
    auto fill = [&](int g, int panel) {
        const int slot = g % N_WS, use = g / N_WS;
        mbar_wait(smem_u32(&p.bar_wempty[slot]), (use & 1) ^ 1);                  // the MMAs that read this slot have retired
        if (elect_one()) {
            mbar_expect_tx(smem_u32(&p.bar_wfull[slot]), W_PANEL_BYTES);
            bulk_g2s(w0 + slot * W_PANEL_BYTES, w_img + (size_t)panel * W_PANEL_BYTES, W_PANEL_BYTES, smem_u32(&p.bar_wfull[slot]));
        }
        __syncwarp();
    };
    fill(g0, 0);
    if (n_panels > 1) fill(g0 + 1, 1);
    for (int kp = 0; kp < n_panels; ++kp) {
        const int g = g0 + kp, slot = g % N_WS;
        mbar_wait(smem_u32(&p.bar_wfull[slot]), (g / N_WS) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KP32 / 8; ++ks) {
                const uint64_t bdesc = make_desc(x_base + kp * XP_BYTES + ks * 32);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    umma_tf32(tmem_d + hh * TN, make_desc(w0 + slot * W_PANEL_BYTES + hh * (128 * 128) + ks * 32), bdesc, idesc,
                              (kp > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(&p.bar_wempty[slot]));
            if (kp == n_panels - 1) umma_commit(smem_u32(&p.bar_acc));
        }
        __syncwarp();
        if (kp + N_WS < n_panels) fill(g + N_WS, kp + N_WS);
    }
}

// ------------------------------------------------------------------------------------------------------------
struct LinSmem {
    unsigned char x[16 * XP_BYTES];          // activation tile, up to K = 512: 128 KB
    unsigned char w[N_WS][W_PANEL_BYTES];    // 64 KB
    Pipe pipe;
};

__global__ void __launch_bounds__(THREADS, 1) linear_tf32_kernel(LinearArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    LinSmem& s = *reinterpret_cast<LinSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row0 = blockIdx.x * TN, col0 = blockIdx.y * 256;
    const int n_panels = a.K / KP32;
    pipe_init(s.pipe, tid, wid);
    const uint32_t tmem_d = s.pipe.tmem_holder;
    // ---- stage the 64 input rows (rounded to tf32): [x | agg] assembled on the fly like linear_f32_kernel
    const int k4 = a.K / 4;
    for (int idx = tid; idx < TN * k4; idx += THREADS) {
        const int i = idx / k4, k = 4 * (idx - i * k4);
        const int row = row0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.n_rows) {
            if (a.two_source && k >= H) v = agg_load4(a.aggv, row, k - H);
            else v = *reinterpret_cast<const float4*>(a.x + (size_t)row * a.ldx + k);
        }
        *reinterpret_cast<float4*>(s.x + x_offset(i, k)) = to_tf32(v);
    }
    fence_proxy_async();
    __syncthreads();
    if (wid == 0) gemm_tile(s.pipe, s.w[0], w_img + (size_t)blockIdx.y * n_panels * W_PANEL_BYTES, n_panels, smem_u32(s.x), tmem_d, 0);
    mbar_wait_relaxed(smem_u32(&s.pipe.bar_acc), 0);
    tc_fence_after();
    // ---- epilogue: thread = output channel (TMEM lane), registers = 16 rows per load
    const int q = wid & 3, half = wid >> 2;
    const int ch = col0 + 128 * half + 32 * q + lane;
    const float bias = a.bias ? a.bias[ch] : 0.f;
#pragma unroll 1
    for (int c = 0; c < TN / 16; ++c) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + half * TN + 16 * c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int r = row0 + 16 * c + j;
            if (r >= a.n_rows) continue;
            float o = v[j] + bias;
            if (a.epi == 1) o = silu_f(o);
            else if (a.epi == 2) o += a.resid[(size_t)r * a.ldr + ch];
            a.y[(size_t)r * a.ldy + ch] = o;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
constexpr int ET = UNIT_F32;                 // edges per tile == segmented-sum unit of the FFMA scheme
static_assert(ET == TN, "one MMA tile per segmented-sum unit");
constexpr int MS = H + 4;

struct EdgeSmem {
    unsigned char x[8 * XP_BYTES];           // first-layer activations, swizzled tf32: 64 KB
    unsigned char w[N_WS][W_PANEL_BYTES];    // 64 KB
    float m[ET][MS];                         // second-layer activations, row-major: 65 KB
    int row[ET]; int col[ET]; int rs[ET]; int re[ET];
    float r2[ET]; float d0[ET]; float gate[ET];
    Pipe pipe;
};

__global__ void __launch_bounds__(THREADS, 1) edge_tf32_kernel(EdgeArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int E = *a.n_edges;
    const int n_tiles = (E + ET - 1) / ET;
    pipe_init(s.pipe, tid, wid);
    const uint32_t tmem_d = s.pipe.tmem_holder;
    int done = 0;                                                  // tiles this CTA has finished (mbarrier parities)

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++done) {
        const int e0 = tile * ET;
        const int cnt = min(ET, E - e0);
        // ---- phase 0: per-edge metadata + current squared distance (coord2diff, egnn_new.py:265-268)
        if (tid < ET) {
            int r = 0, c = 0; float r2 = 0.f, d0 = 0.f; int rs = 0, re = 0;
            if (tid < cnt) {
                r = a.erow[e0 + tid]; c = a.ecol[e0 + tid]; d0 = a.d0[e0 + tid];
                const float dx = a.x[3 * r] - a.x[3 * c], dy = a.x[3 * r + 1] - a.x[3 * c + 1], dz = a.x[3 * r + 2] - a.x[3 * c + 2];
                r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                rs = a.rowptr[r]; re = a.rowptr[r + 1];
            }
            s.row[tid] = r; s.col[tid] = c; s.r2[tid] = r2; s.d0[tid] = d0; s.rs[tid] = rs; s.re[tid] = re;
        }
        __syncthreads();
        // ---- phase 1: first layer in fp32 from the pre-projected rows, rounded to tf32 into the MMA's B tile
        {
            const int c0 = lane * 4, c1 = 128 + lane * 4;
            const float4 wr0 = *reinterpret_cast<const float4*>(a.wr + c0), wr1 = *reinterpret_cast<const float4*>(a.wr + c1);
            const float4 wd0 = *reinterpret_cast<const float4*>(a.wd + c0), wd1 = *reinterpret_cast<const float4*>(a.wd + c1);
            for (int i = wid; i < ET; i += THREADS / 32) {
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                if (i < cnt) {
                    const float* pa = a.p + (size_t)s.row[i] * a.ldp + a.off_a;
                    const float* pb = a.p + (size_t)s.col[i] * a.ldp + a.off_b;
                    const float4 a0 = *reinterpret_cast<const float4*>(pa + c0), a1 = *reinterpret_cast<const float4*>(pa + c1);
                    const float4 b0 = *reinterpret_cast<const float4*>(pb + c0), b1 = *reinterpret_cast<const float4*>(pb + c1);
                    const float r2 = s.r2[i], d0 = s.d0[i];
                    o0.x = silu_f(a0.x + b0.x + r2 * wr0.x + d0 * wd0.x); o0.y = silu_f(a0.y + b0.y + r2 * wr0.y + d0 * wd0.y);
                    o0.z = silu_f(a0.z + b0.z + r2 * wr0.z + d0 * wd0.z); o0.w = silu_f(a0.w + b0.w + r2 * wr0.w + d0 * wd0.w);
                    o1.x = silu_f(a1.x + b1.x + r2 * wr1.x + d0 * wd1.x); o1.y = silu_f(a1.y + b1.y + r2 * wr1.y + d0 * wd1.y);
                    o1.z = silu_f(a1.z + b1.z + r2 * wr1.z + d0 * wd1.z); o1.w = silu_f(a1.w + b1.w + r2 * wr1.w + d0 * wd1.w);
                }
                *reinterpret_cast<float4*>(s.x + x_offset(i, c0)) = to_tf32(o0);
                *reinterpret_cast<float4*>(s.x + x_offset(i, c1)) = to_tf32(o1);
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- phase 2: second layer on the tensor cores
        if (wid == 0) gemm_tile(s.pipe, s.w[0], w_img, H / KP32, smem_u32(s.x), tmem_d, done * (H / KP32));
        mbar_wait_relaxed(smem_u32(&s.pipe.bar_acc), done & 1);
        tc_fence_after();
        // ---- phase 3: bias + SiLU into the row-major tile (thread = channel, 16 edges per load)
        {
            const int q = wid & 3, half = wid >> 2;
            const int ch = 128 * half + 32 * q + lane;
            const float b = a.b2[ch];
#pragma unroll 1
            for (int c = 0; c < ET / 16; ++c) {
                float v[16];
                tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + half * TN + 16 * c, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) s.m[16 * c + j][ch] = silu_f(v[j] + b);
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- phase 4: per-edge scalar = wv . m (+ bv): attention gate or coordinate scalar
        if (a.coord || a.attention) {
            float wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = a.wv[lane + 32 * j];
            for (int i = wid * 8; i < wid * 8 + 8; ++i) {
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) part = fmaf(wv[j], s.m[i][lane + 32 * j], part);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane == 0) {
                    float v = part + a.bv;
                    if (a.coord) v = a.use_tanh ? tanhf(v) : v;     // egnn_new.py:90-93
                    else v = sigmoid_f(v);                          // egnn_new.py:26-29
                    s.gate[i] = v;
                }
            }
        } else if (tid < ET) {
            s.gate[tid] = 1.f;
        }
        __syncthreads();
        // ---- phase 5: coordinate scalar out, or the segmented sum of the tile (egnn_new.py:50-52), FFMA-path scheme
        if (a.coord) {
            if (tid < cnt) a.escal[e0 + tid] = s.gate[tid];
        } else {
            const int c = tid;
            float sum = 0.f;
            for (int i = 0; i < cnt; ++i) {
                sum = fmaf(s.gate[i], s.m[i][c], sum);
                const bool last = (i == cnt - 1) || (s.row[i + 1] != s.row[i]);
                if (last) {
                    const int rs = s.rs[i], re = s.re[i];
                    if (rs >= e0 && re <= e0 + ET) a.agg[(size_t)s.row[i] * H + c] = sum;
                    else a.partials[((size_t)tile * 2 + (rs <= e0 ? 0 : 1)) * H + c] = sum;
                    sum = 0.f;
                }
            }
        }
        tc_fence_after();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

}  // namespace

int tc_tf32_init()
{
    static_assert(sizeof(LinSmem) + 1024 <= 232448 && sizeof(EdgeSmem) + 1024 <= 232448, "tf32 kernels exceed 227 KB of shared memory");
    DP_CUDA(cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(edge_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeSmem) + 1024));
    return DP_OK;
}

static const unsigned char* tf32_image(dp_handle* h, int lin_id, int K, int n_out)
{
    if (!h->tc || lin_id < 0 || lin_id >= (int)h->tc->lin.size()) return nullptr;
    const TcLinearImg& L = h->tc->lin[lin_id];
    if (!L.img_tf32 || L.K != K || L.n_out != n_out) return nullptr;
    return L.img_tf32;
}

int launch_linear_tf32(dp_handle* h, const LinearArgs& a, int lin_id, cudaStream_t st)
{
    DP_CHECK(a.n_out % 256 == 0 && (a.K == H || a.K == 2 * H), DP_ERR_INVALID, "linear_tf32: n_out %d / K %d not tileable", a.n_out, a.K);
    const unsigned char* img = tf32_image(h, lin_id, a.K, a.n_out);
    DP_CHECK(img, DP_ERR_STATE, "tf32 linear %d has no weight image", lin_id);
    if (a.n_rows <= 0) return DP_OK;
    dim3 grid((a.n_rows + TN - 1) / TN, a.n_out / 256);
    linear_tf32_kernel<<<grid, THREADS, (int)sizeof(LinSmem) + 1024, st>>>(a, img);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_edge_tf32(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st)
{
    const unsigned char* img = tf32_image(h, lin_id, H, H);
    DP_CHECK(img, DP_ERR_STATE, "tf32 edge layer %d has no weight image", lin_id);
    edge_tf32_kernel<<<h->sm_count, THREADS, (int)sizeof(EdgeSmem) + 1024, st>>>(a, img);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
