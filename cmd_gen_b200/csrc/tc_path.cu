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
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>

namespace {

constexpr int FMT_F16 = 0, FMT_BF16 = 1;
constexpr int TILE = 64;                 // items (edges / nodes) per MMA tile  (UMMA N)
constexpr int PANEL_K = 64;              // 16-bit elements per 128-byte swizzle row
constexpr int W_PANEL_BYTES = 256 * 128; // 256 out channels x 128 B
constexpr int X_PANEL_BYTES = TILE * 128;
constexpr int THREADS = 512;

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 or bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major; 1) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: [4,6) D fmt (1 = f32) | [7,10) A fmt | [10,13) B fmt | bit 15/16 A/B major
// (0 = K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N)
{
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    if (FMT == FMT_BF16) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    } else {
        __half2 t = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    }
}
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

// write 8 consecutive K elements (one 16-byte chunk) of item `i` into a swizzled tile
__device__ __forceinline__ void store_chunk(unsigned char* tile, int i, int chunk, uint4 v)
{
    unsigned char* p = tile + (chunk >> 3) * X_PANEL_BYTES + i * 128 + (((chunk & 7) ^ (i & 7)) << 4);
    *reinterpret_cast<uint4*>(p) = v;
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
// fused edge kernel (GCL edge model + gate + segmented sum, or coordinate MLP scalar)
// ------------------------------------------------------------------------------------------
struct EdgeMeta {
    int row[TILE]; int col[TILE]; int rs[TILE]; int re[TILE];
    float r2[TILE]; float d0[TILE];
};

struct EdgeTcSmem {                                   // offsets from a 1024-aligned base
    unsigned char w[4 * W_PANEL_BYTES];               // 128 KB: resident second-layer weights
    unsigned char x[2][4 * X_PANEL_BYTES];            // 2 x 32 KB: double-buffered first-layer activations
    EdgeMeta meta[2];
    float red[2][8][32];
    unsigned long long bar_w;
    unsigned long long bar_mma[2];
    uint32_t tmem_holder;
};

template <int FMT>
__global__ void __launch_bounds__(THREADS, 1) edge_tc_kernel(EdgeArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    EdgeTcSmem& s = *reinterpret_cast<EdgeTcSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int E = *a.n_edges;
    const int n_tiles = (E + TILE - 1) / TILE;
    if ((int)blockIdx.x >= n_tiles) return;           // uniform per CTA, before any barrier / allocation

    const uint32_t bar_w = smem_u32(&s.bar_w), bar_m0 = smem_u32(&s.bar_mma[0]), bar_m1 = smem_u32(&s.bar_mma[1]);
    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_m0, 1); mbar_init(bar_m1, 1);
        fence_barrier_init();
    }
    if (wid == 1) tmem_alloc(smem_u32(&s.tmem_holder), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_holder;
    if (tid == 0) {
        mbar_expect_tx(bar_w, 4 * W_PANEL_BYTES);
        for (int p = 0; p < 4; ++p) bulk_g2s(smem_u32(s.w + p * W_PANEL_BYTES), w_img + (size_t)p * W_PANEL_BYTES, W_PANEL_BYTES, bar_w);
    }
    constexpr uint32_t idesc = make_idesc(FMT, 128, TILE);

    // per-thread constants.  Prologue: lane owns channels [8 lane, 8 lane + 8).  Epilogue: one channel.
    float wr[8], wd[8];
    {
        const float4 r0 = *reinterpret_cast<const float4*>(a.wr + 8 * lane), r1 = *reinterpret_cast<const float4*>(a.wr + 8 * lane + 4);
        const float4 d0 = *reinterpret_cast<const float4*>(a.wd + 8 * lane), d1 = *reinterpret_cast<const float4*>(a.wd + 8 * lane + 4);
        wr[0] = r0.x; wr[1] = r0.y; wr[2] = r0.z; wr[3] = r0.w; wr[4] = r1.x; wr[5] = r1.y; wr[6] = r1.z; wr[7] = r1.w;
        wd[0] = d0.x; wd[1] = d0.y; wd[2] = d0.z; wd[3] = d0.w; wd[4] = d1.x; wd[5] = d1.y; wd[6] = d1.z; wd[7] = d1.w;
    }
    const int q = wid & 3, g = wid >> 2, half = g & 1, eg = g >> 1;
    const int w8 = q + 4 * half;
    const int ch = 128 * half + 32 * q + lane;
    const float b2c = a.b2[ch];
    const float wvc = (a.coord || a.attention) ? a.wv[ch] : 0.f;

    int it = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        const bool have = tile < n_tiles;
        const int buf = it & 1;
        const int e0 = tile * TILE;
        if (have) {
            EdgeMeta& m = s.meta[buf];
            if (tid < TILE) {
                int r = 0, c = 0, rs = 0, re = 0; float r2 = 0.f, d0 = 0.f;
                if (e0 + tid < E) {
                    r = a.erow[e0 + tid]; c = a.ecol[e0 + tid]; d0 = a.d0[e0 + tid];
                    const float dx = a.x[3 * r] - a.x[3 * c], dy = a.x[3 * r + 1] - a.x[3 * c + 1], dz = a.x[3 * r + 2] - a.x[3 * c + 2];
                    r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    rs = a.rowptr[r]; re = a.rowptr[r + 1];
                }
                m.row[tid] = r; m.col[tid] = c; m.rs[tid] = rs; m.re[tid] = re; m.r2[tid] = r2; m.d0[tid] = d0;
            }
        }
        __syncthreads();                                                            // (A) metadata visible
        if (have) {
            // ---- prologue: first layer from the pre-projected rows, into the swizzled B tile
            const EdgeMeta& m = s.meta[buf];
            unsigned char* xt = s.x[buf];
            float4 pa[4][2], pb[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 4 * wid + u;
                if (e0 + i < E) {
                    const float* ra = a.p + (size_t)m.row[i] * a.ldp + a.off_a + 8 * lane;
                    const float* rb = a.p + (size_t)m.col[i] * a.ldp + a.off_b + 8 * lane;
                    pa[u][0] = *reinterpret_cast<const float4*>(ra); pa[u][1] = *reinterpret_cast<const float4*>(ra + 4);
                    pb[u][0] = *reinterpret_cast<const float4*>(rb); pb[u][1] = *reinterpret_cast<const float4*>(rb + 4);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 4 * wid + u;
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (e0 + i < E) {
                    const float r2 = m.r2[i], d0 = m.d0[i];
                    const float va[8] = {pa[u][0].x, pa[u][0].y, pa[u][0].z, pa[u][0].w, pa[u][1].x, pa[u][1].y, pa[u][1].z, pa[u][1].w};
                    const float vb[8] = {pb[u][0].x, pb[u][0].y, pb[u][0].z, pb[u][0].w, pb[u][1].x, pb[u][1].y, pb[u][1].z, pb[u][1].w};
                    float y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) y[k] = silu_fast(va[k] + vb[k] + fmaf(r2, wr[k], d0 * wd[k]));
                    o = make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
                }
                store_chunk(xt, i, lane, o);
            }
            fence_proxy_async();                                                    // generic-proxy writes -> async proxy
        }
        tc_fence_before();
        __syncthreads();                                                            // (B) tile complete; TMEM[buf] drained
        if (have && tid == 0) {
            tc_fence_after();
            if (it == 0) mbar_wait(bar_w, 0);
            issue_tile_mma(tmem_base + buf * 2 * TILE, smem_u32(s.w), smem_u32(s.x[buf]), 4, idesc, false);
            umma_commit(buf ? bar_m1 : bar_m0);
        }
        if (it > 0) {
            // ---- epilogue of the previous tile (overlaps the MMAs just issued)
            const int pit = it - 1, pbuf = pit & 1;
            const int ptile = tile - gridDim.x;
            const int pe0 = ptile * TILE;
            const EdgeMeta& m = s.meta[pbuf];
            mbar_wait(pbuf ? bar_m1 : bar_m0, (pit >> 1) & 1);
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + pbuf * 2 * TILE + half * TILE + eg * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j] + b2c);
            const int u0 = pe0 + 32 * eg;                                            // first edge of this 32-edge unit
            const int n_g = min(32, max(0, E - u0));
            float gate = 1.f;
            if (a.coord || a.attention) {
                // dot over the 256 channels: transposing butterfly (lane j ends with edge j), then 8 warps via smem
                float p[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) p[j] = wvc * v[j];
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                    const bool up = (lane & sft) != 0;
#pragma unroll
                    for (int j = 0; j < sft; ++j) {
                        const float keep = up ? p[j + sft] : p[j];
                        const float send = up ? p[j] : p[j + sft];
                        p[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                    }
                }
                s.red[eg][w8][lane] = p[0];
                named_bar_sync(1 + eg, 256);
                float tot = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) tot += s.red[eg][k][lane];
                tot += a.bv;
                if (a.coord) gate = a.use_tanh ? tanhf(tot) : tot;
                else gate = sigmoid_fast(tot);
            }
            if (a.coord) {
                if (w8 == 0 && lane < n_g) a.escal[u0 + lane] = gate;
            } else {
                // segmented sum over this unit's edges: thread = channel, registers = edges
                const int il = 32 * eg + lane;
                const bool is_last = (lane < n_g) && (lane == n_g - 1 || m.row[il + 1] != m.row[il]);
                const unsigned last_mask = __ballot_sync(0xffffffffu, is_last);
                const int unit = ptile * 2 + eg;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < n_g) {
                        const float gj = __shfl_sync(0xffffffffu, gate, j);
                        sum = fmaf(gj, v[j], sum);
                        if ((last_mask >> j) & 1u) {
                            const int ii = 32 * eg + j;
                            const int rs = m.rs[ii], re = m.re[ii];
                            if (rs >= u0 && re <= u0 + 32) a.agg[(size_t)m.row[ii] * H + ch] = sum;
                            else a.partials[((size_t)unit * 2 + (rs <= u0 ? 0 : 1)) * H + ch] = sum;
                            sum = 0.f;
                        }
                    }
                }
            }
        }
        if (!have) break;
        __syncthreads();                                                            // (C) metadata / red reuse
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) tmem_dealloc(tmem_base, 256);
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

struct TcLinearImg { unsigned char* img[2] = {nullptr, nullptr}; int K = 0, n_out = 0; };
struct TcWeights { std::vector<TcLinearImg> lin; std::vector<void*> allocations; };

int tc_init()
{
    DP_CUDA(cudaFuncSetAttribute(edge_tc_kernel<FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeTcSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(edge_tc_kernel<FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeTcSmem) + 1024));
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
    return DP_OK;
}

static int fmt_of(dp_handle* h, int* fmt)
{
    if (h->precision == DP_BF16) { *fmt = FMT_BF16; return DP_OK; }
    if (h->precision == DP_F16) { *fmt = FMT_F16; return DP_OK; }
    dp_set_error("precision mode %d: the tcgen05 kind::tf32 path (streamed fp32 weight tiles) is not built yet; "
                 "use DP_FP32, DP_F16 (same 10-bit mantissa as TF32) or DP_BF16", h->precision);
    return DP_ERR_INVALID;
}

int launch_linear_tc(dp_handle* h, const LinearArgs& a, int lin_id, cudaStream_t st)
{
    int fmt = 0, rc = fmt_of(h, &fmt);
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

int launch_edge_tc(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st)
{
    int fmt = 0, rc = fmt_of(h, &fmt);
    if (rc) return rc;
    DP_CHECK(h->tc && lin_id >= 0 && lin_id < (int)h->tc->lin.size() && h->tc->lin[lin_id].img[fmt], DP_ERR_STATE,
             "tc edge layer %d has no weight image", lin_id);
    const TcLinearImg& L = h->tc->lin[lin_id];
    DP_CHECK(L.K == H && L.n_out == H, DP_ERR_INVALID, "tc edge layer %d: shape mismatch", lin_id);
    const int smem = (int)sizeof(EdgeTcSmem) + 1024;
    const int grid = h->sm_count;
    if (fmt == FMT_BF16) edge_tc_kernel<FMT_BF16><<<grid, THREADS, smem, st>>>(a, L.img[fmt]);
    else edge_tc_kernel<FMT_F16><<<grid, THREADS, smem, st>>>(a, L.img[fmt]);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
