// K2b — the fused per-node phase of one GCL on the 5th-generation tensor cores (DP_BF16 / DP_F16):
//
//   t      = SiLU(W3 [h | agg] + b3)                 GCL.node_model, first layer   (egnn_new.py:21-24, 54-56)
//   h     <- h + W4 t + b4                            second layer + residual        (egnn_new.py:57)
//   P      = Wp h + bp                                the factored FIRST layers of the consumers of the new h:
//                                                     (Pa | Pb) of the next GCL's edge MLP and / or (Qa | Qb) of
//                                                     this block's coordinate MLP — W1 [h_i; h_j; e] = W1a h_i +
//                                                     W1b h_j + W1c e, so the H x H products are per node
//
// One CTA per node tile (capacity 96, sized to fill whole waves of SMs) runs the three GEMMs back to back without leaving the SM: the activations
// of each stage are written by the epilogue straight into the next stage's swizzled K-major B tile in
// shared memory (never to HBM), accumulators live in TMEM (2 x 192 columns), and the weight panels
// (32 KB each: 256 out channels x 64 K, pre-swizzled bf16/f16) stream through a 4-slot ring filled by
// cp.async.bulk from one concatenated image per launch.  Channels-on-lanes orientation as in
// tc_edge.cu: D[out channel, node] = W . X^T, so bias / residual / stores are per-lane scalars and
// 128-byte coalesced rows.
//
//   warps 0-15  compute : stage [h | agg] -> bf16 tiles; the three epilogues
//   warp  16    TMA     : one thread streams the weight panels through the ring
//   warp  17    MMA     : one thread issues tcgen05.mma (M=128, N=tile nodes rounded up to 16, K=16), commits to mbarriers
// node_pair_kernel (below) is the CTA-pair variant: cluster of two, tcgen05 cta_group::2, half of every weight panel per CTA.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int NT = 96;                           // tile CAPACITY in nodes (smem / TMEM layout).  A launch uses `stride` <= NT nodes per CTA so that
                                                 // the tiles fill whole waves of SMs (N = 10 112: 148 CTAs x 69 nodes, not 106 x 96) and an MMA N
                                                 // of stride rounded up to 16
constexpr int NX_PANEL = NT * 128;               // 12 KB: 96 nodes x 64 K x 2 B
constexpr int X_BYTES = 4 * NX_PANEL;            // 48 KB: K = 256
constexpr int N_WS = 4;                          // weight ring slots (the stream is latency-bound: depth = throughput)
constexpr int ROWS_PER_WARP = NT / 16;           // staging: 6 rows per compute warp
constexpr int COLS_PER_WARP = NT / 2;            // epilogues: 48 node columns per warp, as 3 chunks of 16
constexpr int COMPUTE_WARPS = 16;
constexpr int TMA_WARP = 16, MMA_WARP = 17;
constexpr int THREADS = 18 * 32;
constexpr int ACC_COLS = 2 * NT;                 // TMEM columns per accumulator (two 128-channel halves)

struct NodeSmem {
    unsigned char xa[X_BYTES];                   // h tile, later t = SiLU(n0) tile
    unsigned char xb[X_BYTES];                   // agg tile, later the new-h tile
    unsigned char w[N_WS][W_PANEL_BYTES];        // 128 KB ring
    unsigned long long bar_wfull[N_WS], bar_wempty[N_WS];
    unsigned long long bar_x[4];                 // B tile ready: agg, t, new h, h (GEMM 1 starts on the h half while agg is staged)
    unsigned long long bar_accfull[2], bar_accempty[2];
    uint32_t tmem_holder;
};

struct NodeArgs {
    float* h;                                    // [N][H] in / out (in place)
    AggView aggv;
    int n_rows;
    int do_mlp;                                  // 0: projection only (h version 0, straight from the embedding)
    const float* b3; const float* b4;            // node_mlp biases
    const float* bp;                             // projection bias [n_blocks * 256], pre-scaled by 1/2 like the weight image
    __half* pq; int ldp; int n_blocks;           // projection output [N][ldp] f16 (pre-scaled by 1/2), n_blocks x 256 channels
    int stride; int n_mma;                       // nodes per CTA (<= NT) and the UMMA N that covers them (multiple of 16)
    int tp; int sp;                              // the first `tp` CTAs take `sp` nodes each (sp <= stride): the tiles that hold the moving (phar)
                                                 // rows run one projection block more than the others (Qa), so they get fewer nodes — the
                                                 // launch ends with its slowest CTA.  tp = 0: uniform tiles.  Single-CTA kernel only.
    int trace_cta;                               // which CTA writes the debug timeline
    // 16-bit images of the h tiles (one [4 K panels][NT rows][128 B] swizzled B tile per node tile, the layout the MMAs read):
    // a launch that produces an h version stores its new-h tile with one bulk copy per K panel (h16_out), the next launch —
    // same tiling — brings it straight into shared memory through the TMA (h16_in) instead of loading fp32 rows through
    // registers, converting and storing them: the compute warps start on the aggregated messages at once.
    const unsigned char* h16_in; unsigned char* h16_out;
    int row_block; int n_moving;                 // block `row_block` (row part Qa of the coordinate MLP, -1: none) is only read for rows < n_moving
                                                 // (update_coords_mask keeps phar rows, dynamics.py:105-107): tiles past them skip it
    long long* trace;                            // debug timeline (dp_debug_trace), normally null
    int dbg;                                     // timing experiments (DIFFPHAR_DBG), 0 in production
    int fast_silu;                               // SiLU in the one-MUFU tanh form (DP_F16_FAST / DP_F16_FAST32; bf16 always uses it)
    int* range_flag;                             // sticky bit 0: a projected feature beyond the range where the edge kernels' f16 adds stay finite
};

// debug timeline of CTA 0: role 0 = compute warp 0, 1 = MMA thread, 2 = TMA thread; 16 slots per (role, row)
__device__ __forceinline__ void trace_mark(long long* trace, int role, int row, int slot)
{
    if (trace) trace[(role * 64 + row) * 16 + slot] = clock64();
}

// 8 MMAs of one K panel: D[256 x 128] (+)= Wpanel[256 x 64] * Xpanel[128 x 64]^T
__device__ __forceinline__ void issue_panel(uint32_t tmem_d, uint32_t w_slot, uint32_t x_panel, uint32_t idesc, bool first)
{
#pragma unroll
    for (int ks = 0; ks < PANEL_K / 16; ++ks) {
        const uint64_t bdesc = make_desc(x_panel + ks * 32);
        const uint32_t acc = (!first || ks > 0) ? 1u : 0u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const uint64_t adesc = make_desc(w_slot + hh * (128 * 128) + ks * 32);
            umma_f16(tmem_d + hh * NT, adesc, bdesc, idesc, acc);
        }
    }
}

__device__ __forceinline__ uint4 zero4() { return make_uint4(0u, 0u, 0u, 0u); }

template <int FMT>
__device__ __forceinline__ uint4 pack8(const float4& f0, const float4& f1)
{
    return make_uint4(pack2<FMT>(f0.x, f0.y), pack2<FMT>(f0.z, f0.w), pack2<FMT>(f1.x, f1.y), pack2<FMT>(f1.z, f1.w));
}

// h rows (fp32) -> swizzled 16-bit K-major tile.  Warp w owns rows 6w .. 6w+5 of the tile; lane = 16-byte
// chunk; all 12 loads of a warp are in flight together.
template <int FMT>
__device__ __forceinline__ void stage_h(unsigned char* tile, const NodeArgs& a, int n0, int row_end, int wid, int lane)
{
    float4 f[ROWS_PER_WARP][2];
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
        const int row = n0 + ROWS_PER_WARP * wid + u;
        f[u][0] = f[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_end) {
            const float* src = a.h + (size_t)row * H + 8 * lane;
            f[u][0] = *reinterpret_cast<const float4*>(src); f[u][1] = *reinterpret_cast<const float4*>(src + 4);
        }
    }
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u)
        *reinterpret_cast<uint4*>(tile + chunk_offset(ROWS_PER_WARP * wid + u, lane, NX_PANEL)) = pack8<FMT>(f[u][0], f[u][1]);
}

// Aggregated messages -> tile.  agg_src[row] (graph builder, common.cuh) says where the row's sum is: agg[row] for
// all but the <= L - 1 rows that cross a segmented-sum lane boundary; those are the sum of one partial row per lane
// crossed, added in lane order.  unsorted_segment_sum's normalisation (egnn_new.py:283-291) is applied as a
// reciprocal multiply.  `rp` = rowptr[first row of the warp + lane] for lanes 0..6 and `code` = agg_src[first row +
// lane] for lanes 0..5, both loaded by the caller ahead of time.
template <int FMT>
__device__ __forceinline__ void stage_agg(unsigned char* tile, const NodeArgs& a, int n0, int row_end, int wid, int lane, int rp, int code)
{
    const AggView& g = a.aggv;
    const int r0 = n0 + ROWS_PER_WARP * wid;
    // Two batches of three rows.  A row that crosses a segmented-sum boundary is the sum of its pieces in lane / unit order:
    // with 16-edge units and ~7 edges per row (Calpha pockets) four rows in ten have a second piece, so it is requested
    // TOGETHER with the first (one L2 round trip per batch; a dependent trip per split row before: 2.4 per warp).  Third
    // and later pieces (a row longer than a unit) stay a rare serial loop.
    constexpr int RB = ROWS_PER_WARP / 2;
    static_assert(ROWS_PER_WARP % 2 == 0, "two batches");
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        float4 f[RB][2], f2[RB][2];
        int cd[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int ru = RB * hb + u;
            cd[u] = __shfl_sync(0xffffffffu, code, ru);
            if (r0 + ru >= row_end) cd[u] = AGG_EMPTY;
            f[u][0] = f[u][1] = f2[u][0] = f2[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cd[u] != AGG_EMPTY) {
                const unsigned k = (unsigned)(-(cd[u] + 1));                           // split rows: first piece = partial row 2 lf + slot
                const float* src = cd[u] >= 0 ? g.agg + (size_t)cd[u] * H : g.partials + ((size_t)(k >> 11) * 2 + (k & 1u)) * H;
                f[u][0] = *reinterpret_cast<const float4*>(src + 8 * lane);
                f[u][1] = *reinterpret_cast<const float4*>(src + 8 * lane + 4);
                if (cd[u] < 0 && ((k >> 1) & 1023u) >= 1u) {                           // warp-uniform: the second piece
                    const float* s2 = g.partials + ((size_t)((k >> 11) + 1u) * 2) * H + 8 * lane;
                    f2[u][0] = *reinterpret_cast<const float4*>(s2);
                    f2[u][1] = *reinterpret_cast<const float4*>(s2 + 4);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int ru = RB * hb + u;
            float v[8] = {f[u][0].x + f2[u][0].x, f[u][0].y + f2[u][0].y, f[u][0].z + f2[u][0].z, f[u][0].w + f2[u][0].w,
                          f[u][1].x + f2[u][1].x, f[u][1].y + f2[u][1].y, f[u][1].z + f2[u][1].z, f[u][1].w + f2[u][1].w};
            if (cd[u] < 0 && cd[u] != AGG_EMPTY) {                                     // warp-uniform, rare: pieces three and up
                const unsigned k = (unsigned)(-(cd[u] + 1)), lf = k >> 11, extra = (k >> 1) & 1023u;
                for (unsigned i = 2; i <= extra; ++i) {
                    const float* src = g.partials + ((size_t)(lf + i) * 2) * H + 8 * lane;
                    const float4 p0 = *reinterpret_cast<const float4*>(src), p1 = *reinterpret_cast<const float4*>(src + 4);
                    v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
                }
            }
            const int deg = __shfl_sync(0xffffffffu, rp, ru + 1) - __shfl_sync(0xffffffffu, rp, ru);
            const float sc = g.mean ? __fdividef(1.0f, (float)max(deg, 1)) : g.inv_norm;
            *reinterpret_cast<uint4*>(tile + chunk_offset(ROWS_PER_WARP * wid + ru, lane, NX_PANEL)) =
                make_uint4(pack2<FMT>(v[0] * sc, v[1] * sc), pack2<FMT>(v[2] * sc, v[3] * sc),
                           pack2<FMT>(v[4] * sc, v[5] * sc), pack2<FMT>(v[6] * sc, v[7] * sc));
        }
    }
}

template <int FMT>
__device__ __forceinline__ void store_k16(unsigned char* tile, int i, int k, float v)
{
    unsigned char* p = tile + chunk_offset(i, k >> 3, NX_PANEL) + ((k & 7) << 1);
    if (FMT == FMT_BF16) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}

// MC = the launch runs as clusters of two CTAs that SHARE the weight stream: both walk the same panel sequence on their
// own node tiles; CTA r fetches half r of every 32 KB panel and the TMA multicast delivers it to the same ring slot of
// both CTAs (one L2 read instead of two).  Why: all 148 CTAs pull the same 640-896 KB per launch and the L2's output —
// ~6 300 B/clk for the whole chip, 43 B/clk per SM — is what paces the GEMM phases (a panel lands every ~730 cycles,
// its 8 MMAs need ~400).  A ring slot is free when BOTH CTAs' MMAs have read it (multicast tcgen05.commit, count 2);
// a slot is full when both halves have landed (32 KB of complete_tx on the CTA's own mbarrier).
template <int FMT, bool MC>
__global__ void __launch_bounds__(THREADS, 1) node_tc_kernel(NodeArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // the same offset in every CTA (multicast targets it)
    NodeSmem& s = *reinterpret_cast<NodeSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    const int cta = (int)blockIdx.x;
    const int my_stride = cta < a.tp ? a.sp : a.stride;
    const int n0 = cta < a.tp ? cta * a.sp : a.tp * a.sp + (cta - a.tp) * a.stride;
    const int n_mma = cta < a.tp ? (a.sp + 15) / 16 * 16 : a.n_mma;
    const int n0_first = MC ? (int)(blockIdx.x & ~1u) * a.stride : n0;              // first node of the cluster (tp = 0 there)
    long long* const trace_p = cta == a.trace_cta ? a.trace : nullptr;
    const int skip_b = (a.row_block >= 0 && n0_first >= a.n_moving) ? a.row_block : -1;   // cluster-uniform: every role of both CTAs agrees
    const int mlp_panels = a.do_mlp ? 12 : 0;

    if (tid == 0) {
        for (int i = 0; i < N_WS; ++i) { mbar_init(smem_u32(&s.bar_wfull[i]), 1); mbar_init(smem_u32(&s.bar_wempty[i]), MC ? 2 : 1); }
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s.bar_x[i]), (i == 3 && a.h16_in) ? 1 : COMPUTE_WARPS);
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s.bar_accfull[i]), 1); mbar_init(smem_u32(&s.bar_accempty[i]), COMPUTE_WARPS); }
        fence_barrier_init();
    }
    if (wid == MMA_WARP) tmem_alloc(smem_u32(&s.tmem_holder), 512);          // 2 x 192 accumulator columns; allocations are powers of two
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();                                              // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_holder;
    pdl_launch_dependents();
    // Weights are launch-invariant: the TMA thread starts streaming them while the previous kernel drains
    // (programmatic dependent launch); every warp that touches h / agg / pq waits for it first.

    if (wid == TMA_WARP) {
        // ================================ weight stream ================================
        {
            const uint32_t w0 = warp_uniform(smem_u32(s.w[0]));
            if (a.h16_in && a.do_mlp) {
                // the h tile of this CTA, written by the previous node launch in the layout GEMM 1 reads: rows [0, n_mma) of each K panel
                pdl_wait();
                if (elect_one()) {
                    const uint32_t bytes = (uint32_t)n_mma * 128u;
                    mbar_expect_tx(smem_u32(&s.bar_x[3]), 4 * bytes);
#pragma unroll
                    for (int kp = 0; kp < 4; ++kp)
                        bulk_g2s(smem_u32(s.xa) + kp * NX_PANEL, a.h16_in + ((size_t)cta * 4 + kp) * NX_PANEL, bytes, smem_u32(&s.bar_x[3]));
                }
                __syncwarp();
            }
            int p = 0;                                                                // ring position
            for (int g = 0; g < mlp_panels + 4 * a.n_blocks; ++g) {                    // g = panel of the image
                if (g >= mlp_panels && (g - mlp_panels) / 4 == skip_b) continue;
                const int slot = p % N_WS;
                if (lane == 0) trace_mark(trace_p, 2, p, 0);
                mbar_wait(smem_u32(&s.bar_wempty[slot]), ((p / N_WS) & 1) ^ 1);
                if (lane == 0) trace_mark(trace_p, 2, p, 1);
                if (elect_one()) {
                    mbar_expect_tx(smem_u32(&s.bar_wfull[slot]), W_PANEL_BYTES);
                    if (MC)                                                            // my half, into both CTAs' slot; the peer sends the other half
                        bulk_g2s_multicast(w0 + slot * W_PANEL_BYTES + rank * (W_PANEL_BYTES / 2),
                                           w_img + (size_t)g * W_PANEL_BYTES + rank * (W_PANEL_BYTES / 2), W_PANEL_BYTES / 2,
                                           smem_u32(&s.bar_wfull[slot]), 3);
                    else
                        bulk_g2s(w0 + slot * W_PANEL_BYTES, w_img + (size_t)g * W_PANEL_BYTES, W_PANEL_BYTES, smem_u32(&s.bar_wfull[slot]));
                }
                __syncwarp();
                ++p;
            }
        }
    } else if (wid == MMA_WARP) {
        // ================================ MMA issuer ================================
        // the whole warp runs the control flow (uniform: descriptors stay in uniform registers), one elected lane issues
        {
            const uint32_t idesc = make_idesc(FMT, 128, n_mma);
            const uint32_t td = warp_uniform(tmem_base);
            const uint32_t w0 = warp_uniform(smem_u32(s.w[0]));
            const uint32_t xa = warp_uniform(smem_u32(s.xa)), xb = warp_uniform(smem_u32(s.xb));
            const bool tr = lane == 0;
            int p = 0;
            int acc_uses[2] = {0, 0};
            auto run_gemm = [&](int acc, uint32_t x0, uint32_t x1, int k_panels) {
                // x0: first 4 panels' tile, x1: panels 4-7 (K = 512 only: the agg half, staged after the h half)
                if (acc_uses[acc] > 0) mbar_wait(smem_u32(&s.bar_accempty[acc]), (acc_uses[acc] - 1) & 1);
                tc_fence_after();
                for (int kp = 0; kp < k_panels; ++kp, ++p) {
                    const int slot = p % N_WS;
                    if (kp == 4) mbar_wait(smem_u32(&s.bar_x[0]), 0);
                    if (tr) trace_mark(trace_p, 1, p, 0);
                    mbar_wait(smem_u32(&s.bar_wfull[slot]), (p / N_WS) & 1);
                    if (tr) trace_mark(trace_p, 1, p, 1);
                    tc_fence_after();
                    const uint32_t xp = (kp < 4 ? x0 + kp * NX_PANEL : x1 + (kp - 4) * NX_PANEL);
                    if (elect_one()) {
                        issue_panel(td + acc * ACC_COLS, w0 + slot * W_PANEL_BYTES, xp, idesc, kp == 0);
                        if (MC) umma_commit_multicast(smem_u32(&s.bar_wempty[slot]), 3);
                        else umma_commit(smem_u32(&s.bar_wempty[slot]));
                        if (kp == k_panels - 1) umma_commit(smem_u32(&s.bar_accfull[acc]));
                    }
                    __syncwarp();
                    if (tr) trace_mark(trace_p, 1, p, 2);
                }
                acc_uses[acc] += 1;
            };
            if (a.do_mlp) {
                mbar_wait(smem_u32(&s.bar_x[3]), 0);
                run_gemm(0, xa, xb, 8);
                mbar_wait(smem_u32(&s.bar_x[1]), 0);
                run_gemm(1, xa, 0, 4);
            }
            mbar_wait(smem_u32(&s.bar_x[2]), 0);
            if (a.h16_out && lane == 0) {                                            // the new h tile is complete (and stays: the projections only read it)
                const uint32_t bytes = (uint32_t)n_mma * 128u;
#pragma unroll
                for (int kp = 0; kp < 4; ++kp) bulk_s2g(a.h16_out + ((size_t)cta * 4 + kp) * NX_PANEL, xb + kp * NX_PANEL, bytes);
                bulk_commit_group();
            }
            __syncwarp();
            for (int b = 0, cnt = 0; b < a.n_blocks; ++b) {
                if (b == skip_b) continue;
                run_gemm(cnt & 1, xb, 0, 4);
                ++cnt;
            }
            if (a.h16_out && lane == 0) bulk_wait_group0();                          // the image is in global memory before the CTA ends
            __syncwarp();
        }
    } else {
        // ================================ compute warps ================================
        // epilogue mapping: warp = (TMEM lane quarter q, channel half, 48-column half of the tile);
        // thread = one output channel, registers = 16 node columns per tcgen05.ld
        const int q = wid & 3, g = wid >> 2, half = g >> 1, cpar = g & 1;
        const int ch = 128 * half + 32 * q + lane;
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        int acc_uses[2] = {0, 0};
        auto wait_acc = [&](int acc) {
            mbar_wait(smem_u32(&s.bar_accfull[acc]), acc_uses[acc] & 1);
            acc_uses[acc] += 1;
            tc_fence_after();
        };
        auto release_acc = [&](int acc) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.bar_accempty[acc]));
        };
        auto publish = [&](int which) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.bar_x[which]));
        };
        const bool tr = wid == 0 && lane == 0;
        pdl_wait();
        if (tr) trace_mark(trace_p, 0, 0, 0);
        const int n_valid = min(my_stride, a.n_rows - n0);                           // rows of this tile that exist
        if (a.do_mlp) {
            int rp = 0;                                                              // CSR bounds of the warp's rows: issued first
            if (lane <= ROWS_PER_WARP) rp = a.aggv.rowptr[min(n0 + ROWS_PER_WARP * wid + lane, a.n_rows)];
            int code = AGG_EMPTY;                                                    // where each row's aggregate lives
            if (lane < ROWS_PER_WARP && n0 + ROWS_PER_WARP * wid + lane < a.n_rows) code = a.aggv.src[n0 + ROWS_PER_WARP * wid + lane];
            if (!a.h16_in) {                                                         // else: the TMA warp brings the 16-bit tile
                stage_h<FMT>(s.xa, a, n0, n0 + n_valid, wid, lane);
                publish(3);
            }
            if (tr) trace_mark(trace_p, 0, 0, 1);
            stage_agg<FMT>(s.xb, a, n0, n0 + n_valid, wid, lane, rp, code);
            publish(0);
            if (tr) trace_mark(trace_p, 0, 0, 2);
            // ---- epilogue 1: t = SiLU(D1 + b3) -> xa (the n0 MMAs have all retired: acc_full follows them)
            // A warp takes every other 16-column chunk (i0 = 16 (2 cc + cpar)), so both warps of a channel half stay
            // busy whatever the tile's node count; chunks past the tile's last node are skipped.
            const float b3c = a.b3[ch];
            wait_acc(0);
            // xa is re-used for t: the accumulator-full mbarrier (completed by tcgen05.commit after the last MMA that read the
            // tile) already orders these stores after every warp's staging stores AND the tensor core's reads; the CTA
            // barrier among the compute warps adds nothing to that but lets racecheck, which does not model the commit, see it
            named_bar_sync(1, COMPUTE_WARPS * 32);
            if (tr) trace_mark(trace_p, 0, 0, 3);
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {
                const int i0 = 16 * (2 * cc + cpar);
                if (i0 >= n_valid) break;
                float v[16];
                tmem_ld16(tmem_base + t_lane + half * NT + i0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += b3c;
                if (a.fast_silu) {                                                   // one-MUFU tanh form (DP_F16_FAST): epilogue 1 is MUFU-bound
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = silu_tc<FMT_BF16>(v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = silu_tc<FMT>(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) store_k16<FMT>(s.xa, i0 + j, ch, v[j]);
            }
            release_acc(0);
            publish(1);
            if (tr) trace_mark(trace_p, 0, 0, 4);
            // ---- epilogue 2: h <- h + D2 + b4 (fp32, in place) and its 16-bit copy -> xb
            const float b4c = a.b4[ch];
            float* hrow = a.h + (size_t)n0 * H + ch;
            float r[16];                                                             // residual rows of chunk 0: fetched under the MMAs
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = (16 * cpar + j < n_valid) ? hrow[(size_t)(16 * cpar + j) * H] : 0.f;
            wait_acc(1);
            named_bar_sync(1, COMPUTE_WARPS * 32);                                   // xb is re-used for the new h: see above
            if (tr) trace_mark(trace_p, 0, 0, 5);
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {
                const int i0 = 16 * (2 * cc + cpar);
                if (i0 >= n_valid) break;
                float v[16], rn[16];
                tmem_ld16(tmem_base + ACC_COLS + t_lane + half * NT + i0, v);
                if (cc < 2) {                                                        // next chunk's residual rows
#pragma unroll
                    for (int j = 0; j < 16; ++j) rn[j] = (i0 + 32 + j < n_valid) ? hrow[(size_t)(i0 + 32 + j) * H] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float o = r[j] + (v[j] + b4c);
                    if (i0 + j < n_valid) hrow[(size_t)(i0 + j) * H] = o;
                    store_k16<FMT>(s.xb, i0 + j, ch, (i0 + j < n_valid) ? o : 0.f);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = rn[j];
            }
            release_acc(1);
            publish(2);
            if (tr) trace_mark(trace_p, 0, 0, 6);
        } else {
            stage_h<FMT>(s.xb, a, n0, n0 + n_valid, wid, lane);
            publish(2);
        }
        // ---- epilogue 3: projection blocks -> P (f16: halves the edge kernels' gather bytes and staging registers)
        float big = 0.f;                                                             // largest |P'| this thread stored (f16 range guard)
        for (int b = 0, cnt = 0; b < a.n_blocks; ++b) {
            if (b == skip_b) continue;
            const int acc = cnt & 1;
            ++cnt;
            const float bias = a.bp[b * 256 + ch];
            __half* dst = a.pq + (size_t)n0 * a.ldp + (size_t)b * 256 + ch;
            wait_acc(acc);
            if (tr && b < 4) trace_mark(trace_p, 0, 0, 7 + 2 * b);
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {
                const int i0 = 16 * (2 * cc + cpar);
                if (i0 >= n_valid) break;
                float v[16];
                tmem_ld16(tmem_base + acc * ACC_COLS + t_lane + half * NT + i0, v);
                __half* d = dst + (size_t)i0 * a.ldp;
#pragma unroll
                for (int j = 0; j < 16; ++j, d += a.ldp)
                    if (i0 + j < n_valid) { const float o = v[j] + bias; big = fmaxf(big, fabsf(o)); *d = __float2half_rn(o); }
            }
            release_acc(acc);
            if (tr && b < 4) trace_mark(trace_p, 0, 0, 8 + 2 * b);
        }
        if (big > 32000.f) atomicOr(a.range_flag, 1);                                // pq is f16 (pre-halved): Pa' + Pb' must stay finite
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();                                              // nobody leaves while the peer's commits / copies may still target it
    if (wid == MMA_WARP) tmem_dealloc(tmem_base, 512);
}


// ====================================================================================================
// CTA-pair version: a cluster of two CTAs (one TPC) runs each GEMM as tcgen05.mma.cta_group::2 —
// M = 256 output channels split 128 / 128 across the two SMs, N = the nodes of BOTH CTAs.  Each CTA
// streams only ITS half of every weight panel (16 KB instead of 32 KB): the per-SM weight traffic — the
// single-CTA kernel's bound (every CTA pulls all 640-896 KB through L2 -> SM delivery) — is halved.
// CTA r's tensor memory holds channels 128 r .. 128 r + 127 for all nodes of the pair, so its epilogue
// produces K panels 2 r and 2 r + 1 of the next GEMM's B tile of BOTH CTAs: the own half goes straight into
// the own tile, the peer's half into a staging buffer that ONE bulk copy (cp.async.bulk shared::cta ->
// shared::cluster, async proxy on both ends, complete_tx on the peer's mbarrier) ships across — remote
// 16-bit generic stores measured 2 k cycles slower per epilogue and need a GPU-scope membar.
//   * leader (cluster rank 0) MMA warp issues every MMA and commits with .multicast::cluster to the
//     mbarriers of both CTAs; the peer's MMA warp forwards "my half of panel p / my received half tile has
//     landed" to the leader in the order the leader consumes them;
//   * tile-ready and accumulator-drained barriers live in the leader and count the compute warps of both
//     CTAs (relaxed remote arrive after a CTA-scope fence.proxy.async: every CTA only writes its OWN memory).
// ====================================================================================================
constexpr int N_WS2 = 6;                          // ring slots of half panels
constexpr int W_HALF_BYTES = W_PANEL_BYTES / 2;   // 16 KB: 128 channels x 64 K

struct NodeSmem2 {
    unsigned char xa[X_BYTES];
    unsigned char xb[X_BYTES];
    unsigned char stg[2 * NX_PANEL];              // this CTA's two K panels of the PEER's next B tile, before the bulk copy
    unsigned char w[N_WS2][W_HALF_BYTES];         // 96 KB ring
    unsigned long long bar_wfull[N_WS2], bar_wempty[N_WS2], bar_wpeer[N_WS2];
    unsigned long long bar_x[4];                  // used in the leader: agg, t, new h, h tiles (own parts) of BOTH CTAs written
    unsigned long long bar_stage[2];              // per CTA: staging buffer complete (t, new h) -> warp 0 ships it
    unsigned long long bar_rx[2];                 // per CTA: the peer's half of my t / new-h tile has landed (complete_tx)
    unsigned long long bar_rxpeer[2];             // used in the leader: the peer received ITS half
    unsigned long long bar_accfull[2];            // per CTA (multicast commit)
    unsigned long long bar_accempty[2];           // used in the leader: both CTAs drained the accumulator
    uint32_t tmem_holder;
};

// 4 MMAs of one K panel: D[256 x N] (+)= [W half of rank 0 ; W half of rank 1][256 x 64] * [X rank 0 ; X rank 1][N x 64]^T
__device__ __forceinline__ void issue_panel_pair(uint32_t tmem_d, uint32_t w_slot, uint32_t x_panel, uint32_t idesc, bool first)
{
#pragma unroll
    for (int ks = 0; ks < PANEL_K / 16; ++ks)
        umma_f16_pair(tmem_d, make_desc(w_slot + ks * 32), make_desc(x_panel + ks * 32), idesc, (!first || ks > 0) ? 1u : 0u);
}

template <int FMT>
__device__ __forceinline__ unsigned short bits16(float v)
{
    if (FMT == FMT_BF16) { const __nv_bfloat16 t = __float2bfloat16_rn(v); return *reinterpret_cast<const unsigned short*>(&t); }
    const __half t = __float2half_rn(v);
    return *reinterpret_cast<const unsigned short*>(&t);
}

template <int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) node_pair_kernel(NodeArgs a, const unsigned char* __restrict__ w_img)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // same offset in both CTAs
    NodeSmem2& s = *reinterpret_cast<NodeSmem2*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int n0_pair = (int)(blockIdx.x & ~1u) * a.stride;                  // first node of rank 0
    const int n0 = n0_pair + (int)rank * a.stride;                           // first node of this CTA
    const int skip_b = (a.row_block >= 0 && n0_pair >= a.n_moving) ? a.row_block : -1;   // pair-uniform
    const int mlp_panels = a.do_mlp ? 12 : 0;
    const int nm = a.n_mma;                                                  // B rows per CTA (multiple of 16); the MMA's N = 2 nm
    long long* const trace_p = blockIdx.x == 0 ? a.trace : nullptr;
    const uint32_t ship_bytes = (uint32_t)(NX_PANEL + nm * 128);             // panel 0 whole + rows < nm of panel 1 (rows >= nm are never read)

    if (tid == 0) {
        for (int i = 0; i < N_WS2; ++i) {
            mbar_init(smem_u32(&s.bar_wfull[i]), 1); mbar_init(smem_u32(&s.bar_wempty[i]), 1); mbar_init(smem_u32(&s.bar_wpeer[i]), 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s.bar_x[i]), 2 * COMPUTE_WARPS);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s.bar_stage[i]), COMPUTE_WARPS); mbar_init(smem_u32(&s.bar_rx[i]), 1); mbar_init(smem_u32(&s.bar_rxpeer[i]), 1);
            mbar_init(smem_u32(&s.bar_accfull[i]), 1); mbar_init(smem_u32(&s.bar_accempty[i]), 2 * COMPUTE_WARPS);
        }
        fence_barrier_init();
        if (a.do_mlp) { mbar_expect_tx(smem_u32(&s.bar_rx[0]), ship_bytes); mbar_expect_tx(smem_u32(&s.bar_rx[1]), ship_bytes); }
    }
    if (wid == MMA_WARP) tmem_alloc_pair(smem_u32(&s.tmem_holder), 512);     // the same warp of both CTAs
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                      // the peer's barriers exist (and are armed) before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_holder;
    pdl_launch_dependents();

    if (wid == TMA_WARP) {
        // ================================ weight stream: this CTA's 128 channels of every panel ================================
        const uint32_t w0 = warp_uniform(smem_u32(s.w[0]));
        int p = 0;
        for (int g = 0; g < mlp_panels + 4 * a.n_blocks; ++g) {
            if (g >= mlp_panels && (g - mlp_panels) / 4 == skip_b) continue;
            const int slot = p % N_WS2;
            if (lane == 0) trace_mark(trace_p, 2, p, 0);
            mbar_wait(smem_u32(&s.bar_wempty[slot]), ((p / N_WS2) & 1) ^ 1);
            if (lane == 0) trace_mark(trace_p, 2, p, 1);
            if (elect_one()) {
                mbar_expect_tx(smem_u32(&s.bar_wfull[slot]), W_HALF_BYTES);
                bulk_g2s(w0 + slot * W_HALF_BYTES, w_img + (size_t)g * W_PANEL_BYTES + (size_t)rank * W_HALF_BYTES, W_HALF_BYTES,
                         smem_u32(&s.bar_wfull[slot]));
            }
            __syncwarp();
            ++p;
        }
    } else if (wid == MMA_WARP) {
        const int n_proj = a.n_blocks - (skip_b >= 0 ? 1 : 0);
        if (rank == 0) {
            // ================================ MMA issuer (leader) ================================
            const uint32_t idesc = make_idesc(FMT, 256, 2 * nm);
            const uint32_t td = warp_uniform(tmem_base);
            const uint32_t w0 = warp_uniform(smem_u32(s.w[0]));
            const uint32_t xa = warp_uniform(smem_u32(s.xa)), xb = warp_uniform(smem_u32(s.xb));
            int p = 0;
            int acc_uses[2] = {0, 0};
            // tile `which` of both CTAs is complete: own parts written (32 warps) and, for the exchanged tiles
            // (t = 1, new h = 2 after the MLP), both shipped halves landed
            auto wait_tile = [&](int which, int rx) {
                mbar_wait(smem_u32(&s.bar_x[which]), 0);
                if (rx >= 0) { mbar_wait(smem_u32(&s.bar_rx[rx]), 0); mbar_wait(smem_u32(&s.bar_rxpeer[rx]), 0); }
            };
            auto run_gemm = [&](int acc, uint32_t x0, uint32_t x1, int k_panels) {
                if (acc_uses[acc] > 0) mbar_wait(smem_u32(&s.bar_accempty[acc]), (acc_uses[acc] - 1) & 1);   // a signal: nothing to acquire
                tc_fence_after();
                for (int kp = 0; kp < k_panels; ++kp, ++p) {
                    const int slot = p % N_WS2;
                    if (kp == 4) wait_tile(0, -1);
                    if (lane == 0) trace_mark(trace_p, 1, p, 0);
                    mbar_wait(smem_u32(&s.bar_wfull[slot]), (p / N_WS2) & 1);
                    mbar_wait(smem_u32(&s.bar_wpeer[slot]), (p / N_WS2) & 1);      // the peer's half arrived through ITS async proxy
                    tc_fence_after();
                    if (lane == 0) trace_mark(trace_p, 1, p, 1);
                    const uint32_t xp = (kp < 4 ? x0 + kp * NX_PANEL : x1 + (kp - 4) * NX_PANEL);
                    if (elect_one()) {
                        issue_panel_pair(td + acc * ACC_COLS, w0 + slot * W_HALF_BYTES, xp, idesc, kp == 0);
                        umma_commit_pair(smem_u32(&s.bar_wempty[slot]), 3);
                        if (kp == k_panels - 1) umma_commit_pair(smem_u32(&s.bar_accfull[acc]), 3);
                    }
                    __syncwarp();
                    if (lane == 0) trace_mark(trace_p, 1, p, 2);
                }
                acc_uses[acc] += 1;
            };
            if (a.do_mlp) {
                wait_tile(3, -1);
                run_gemm(0, xa, xb, 8);
                wait_tile(1, 0);
                run_gemm(1, xa, 0, 4);
                wait_tile(2, 1);
            } else {
                wait_tile(2, -1);
            }
            for (int cnt = 0; cnt < n_proj; ++cnt) run_gemm(cnt & 1, xb, 0, 4);
        } else {
            // ================================ peer: forwards its arrivals to the leader, in the leader's order ================================
            const int total = mlp_panels + 4 * n_proj;
            for (int p = 0; p < total; ++p) {
                if (a.do_mlp && (p == 8 || p == 12)) {                               // the leader's half of my t / new-h tile
                    const int rx = p == 8 ? 0 : 1;
                    mbar_wait(smem_u32(&s.bar_rx[rx]), 0);
                    if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&s.bar_rxpeer[rx]), 0));
                    __syncwarp();
                }
                const int slot = p % N_WS2;
                mbar_wait(smem_u32(&s.bar_wfull[slot]), (p / N_WS2) & 1);
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&s.bar_wpeer[slot]), 0));
                __syncwarp();
            }
        }
    } else {
        // ================================ compute warps ================================
        // epilogue mapping: warp = (TMEM lane quarter q, chunk phase g); thread = one output channel of this CTA's
        // half, registers = 16 node columns per tcgen05.ld; the warp takes the 16-column chunks g, g + 4, g + 8 of
        // the pair's 2 nm columns (columns [0, nm) are rank 0's nodes, [nm, 2 nm) rank 1's)
        const int q = wid & 3, g = wid >> 2;
        const int ch = 128 * (int)rank + 32 * q + lane;
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        const int nv[2] = {max(0, min(a.stride, a.n_rows - n0_pair)), max(0, min(a.stride, a.n_rows - n0_pair - a.stride))};
        const int n_valid = nv[rank];
        const int n_chunks = 2 * nm / 16;
        const uint32_t peer = rank ^ 1u;
        uint32_t bx[4], bempty[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) bx[i] = mapa_u32(smem_u32(&s.bar_x[i]), 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) bempty[i] = mapa_u32(smem_u32(&s.bar_accempty[i]), 0);
        int acc_uses[2] = {0, 0};
        auto wait_acc = [&](int acc) {
            mbar_wait(smem_u32(&s.bar_accfull[acc]), acc_uses[acc] & 1);
            acc_uses[acc] += 1;
            tc_fence_after();
        };
        auto release_acc = [&](int acc) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(bempty[acc]);         // tcgen05.ld results are in registers (wait::ld): a pure signal
        };
        // this warp's share of tile `which` (and of the staging buffer `stage`, -1: none) is written: CTA-scope proxy
        // fence (every CTA writes only its own shared memory), then plain signals; warp 0 ships the staging buffer
        auto publish = [&](int which, int stage, unsigned char* peer_tile) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster_relaxed(bx[which]);
                if (stage >= 0) mbar_arrive(smem_u32(&s.bar_stage[stage]));
            }
            if (stage >= 0 && wid == 0) {
                mbar_wait(smem_u32(&s.bar_stage[stage]), 0);
                if (elect_one())
                    bulk_s2peer(mapa_u32(smem_u32(peer_tile) + 2 * rank * NX_PANEL, peer), smem_u32(s.stg), ship_bytes,
                                mapa_u32(smem_u32(&s.bar_rx[stage]), peer));
                __syncwarp();
            }
        };
        // 16-bit element (node row i of CTA `owner`, K = ch) of the next B tile: own tile, or the staging copy of the
        // peer's K panels 2 rank, 2 rank + 1 (same swizzle: 16 rank chunks = 0 mod 8)
        auto store_tile = [&](unsigned char* local_tile, int owner, int i, float v) {
            unsigned char* dst = owner == (int)rank ? local_tile + chunk_offset(i, ch >> 3, NX_PANEL)
                                                    : s.stg + chunk_offset(i, (ch >> 3) - 16 * (int)rank, NX_PANEL);
            *reinterpret_cast<unsigned short*>(dst + ((ch & 7) << 1)) = bits16<FMT>(v);
        };
        const bool tr = wid == 0 && lane == 0;
        pdl_wait();
        if (tr) trace_mark(trace_p, 0, 0, 0);
        if (a.do_mlp) {
            int rp = 0;
            if (lane <= ROWS_PER_WARP) rp = a.aggv.rowptr[min(n0 + ROWS_PER_WARP * wid + lane, a.n_rows)];
            int code = AGG_EMPTY;                                                    // where each row's aggregate lives
            if (lane < ROWS_PER_WARP && n0 + ROWS_PER_WARP * wid + lane < a.n_rows) code = a.aggv.src[n0 + ROWS_PER_WARP * wid + lane];
            stage_h<FMT>(s.xa, a, n0, n0 + n_valid, wid, lane);
            publish(3, -1, nullptr);
            if (tr) trace_mark(trace_p, 0, 0, 1);
            stage_agg<FMT>(s.xb, a, n0, n0 + n_valid, wid, lane, rp, code);
            publish(0, -1, nullptr);
            if (tr) trace_mark(trace_p, 0, 0, 2);
            // ---- epilogue 1: t = SiLU(D1 + b3) -> K panels 2 rank, 2 rank + 1 of both xa tiles
            const float b3c = a.b3[ch];
            wait_acc(0);
            if (tr) trace_mark(trace_p, 0, 0, 3);
#pragma unroll 1
            for (int c = g; c < n_chunks; c += 4) {
                const int i0 = 16 * c, owner = i0 >= nm ? 1 : 0, l0 = i0 - owner * nm;
                if (l0 >= nv[owner]) continue;
                float v[16];
                tmem_ld16(tmem_base + t_lane + i0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += b3c;
                if (a.fast_silu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = silu_tc<FMT_BF16>(v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = silu_tc<FMT>(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) store_tile(s.xa, owner, l0 + j, v[j]);
            }
            release_acc(0);
            publish(1, 0, s.xa);
            if (tr) trace_mark(trace_p, 0, 0, 4);
            // ---- epilogue 2: h <- h + D2 + b4 (fp32, in place) and its 16-bit copy -> both xb tiles.  The staging
            // buffer is free again: acc 1 is full only after GEMM 2 consumed the t tiles, i.e. after the t shipment landed.
            const float b4c = a.b4[ch];
            wait_acc(1);
            if (tr) trace_mark(trace_p, 0, 0, 5);
#pragma unroll 1
            for (int c = g; c < n_chunks; c += 4) {
                const int i0 = 16 * c, owner = i0 >= nm ? 1 : 0, l0 = i0 - owner * nm;
                const int nvo = nv[owner];
                if (l0 >= nvo) continue;
                float* hrow = a.h + (size_t)(n0_pair + owner * a.stride + l0) * H + ch;
                float r[16], v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = (l0 + j < nvo) ? hrow[(size_t)j * H] : 0.f;
                tmem_ld16(tmem_base + ACC_COLS + t_lane + i0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float o = r[j] + (v[j] + b4c);
                    if (l0 + j < nvo) hrow[(size_t)j * H] = o;
                    store_tile(s.xb, owner, l0 + j, (l0 + j < nvo) ? o : 0.f);
                }
            }
            release_acc(1);
            publish(2, 1, s.xb);
            if (tr) trace_mark(trace_p, 0, 0, 6);
        } else {
            stage_h<FMT>(s.xb, a, n0, n0 + n_valid, wid, lane);
            publish(2, -1, nullptr);
        }
        // ---- epilogue 3: projection blocks -> P (f16)
        float big = 0.f;
        for (int b = 0, cnt = 0; b < a.n_blocks; ++b) {
            if (b == skip_b) continue;
            const int acc = cnt & 1;
            ++cnt;
            const float bias = a.bp[b * 256 + ch];
            wait_acc(acc);
            if (tr && b < 4) trace_mark(trace_p, 0, 0, 7 + 2 * b);
#pragma unroll 1
            for (int c = g; c < n_chunks; c += 4) {
                const int i0 = 16 * c, owner = i0 >= nm ? 1 : 0, l0 = i0 - owner * nm;
                const int nvo = nv[owner];
                if (l0 >= nvo) continue;
                float v[16];
                tmem_ld16(tmem_base + acc * ACC_COLS + t_lane + i0, v);
                __half* d = a.pq + (size_t)(n0_pair + owner * a.stride + l0) * a.ldp + (size_t)b * 256 + ch;
#pragma unroll
                for (int j = 0; j < 16; ++j, d += a.ldp)
                    if (l0 + j < nvo) { const float o = v[j] + bias; big = fmaxf(big, fabsf(o)); *d = __float2half_rn(o); }
            }
            release_acc(acc);
            if (tr && b < 4) trace_mark(trace_p, 0, 0, 8 + 2 * b);
        }
        if (big > 32000.f) atomicOr(a.range_flag, 1);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                      // nobody leaves while the peer may still touch its shared / tensor memory
    if (wid == MMA_WARP) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace

int tc_node_init()
{
    static_assert(sizeof(NodeSmem) + 1024 <= 232448, "node kernel shared memory exceeds 227 KB");
    DP_CUDA(cudaFuncSetAttribute(node_tc_kernel<tc::FMT_BF16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(node_tc_kernel<tc::FMT_F16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(node_tc_kernel<tc::FMT_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem) + 1024));
    DP_CUDA(cudaFuncSetAttribute(node_tc_kernel<tc::FMT_F16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem) + 1024));
    static_assert(sizeof(NodeSmem2) + 1024 <= 232448, "pair node kernel shared memory exceeds 227 KB");
    DP_CUDA(cudaFuncSetAttribute(node_pair_kernel<tc::FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem2) + 1024));
    DP_CUDA(cudaFuncSetAttribute(node_pair_kernel<tc::FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NodeSmem2) + 1024));
    return DP_OK;
}

size_t node_tile_image_bytes() { return (size_t)X_BYTES; }

// Whole waves of SMs: the fewest waves that fit NT-node tiles, then the smallest stride that keeps that count.  The tiles
// that hold phar rows (one projection block more than the rest in most launches) get h->node_split nodes each; every
// launch of a denoiser evaluation uses the same tiling (the h images are exchanged tile by tile).
void node_tiling(const dp_handle* h, int N, int Np, NodeTiling* t)
{
    const int sm = h->sm_count > 0 ? h->sm_count : 148;
    const int waves = std::max(1, (N + NT * sm - 1) / (NT * sm));
    int stride = (N + waves * sm - 1) / (waves * sm);
    if (stride < 16) stride = 16;
    t->tp = 0; t->sp = stride; t->stride = stride; t->grid = (N + stride - 1) / stride;
    if (h->node_split && !h->joint && !h->node_pair && !h->node_mc && Np > 0 && Np < N) {
        const int sp = std::min(stride, h->node_split), tp = (Np + sp - 1) / sp, rest = N - tp * sp;
        const int ctas = waves * sm - tp;
        if (rest > 0 && ctas > 0) {
            const int sr = std::max(16, (rest + ctas - 1) / ctas);
            if (sr <= NT) { t->tp = tp; t->sp = sp; t->stride = sr; t->grid = tp + (rest + sr - 1) / sr; }
        }
    }
}

// One launch per h version v: v = 0 is the projection of the embedded features; v = i + 1 runs GCL i's node
// model and then projects the new h for its consumers.
int launch_node_tc(dp_handle* h, int v, const AggView& av, cudaStream_t st)
{
    int fmt = 0, rc = tc_fmt_of(h, &fmt);
    if (rc) return rc;
    Plan& p = h->plan; const DeviceWeights& W = h->w;
    DP_CHECK(h->tc && v >= 0 && v < (int)h->tc->node.size() && h->tc->node[v].img[fmt], DP_ERR_STATE,
             "tc node phase %d has no weight image", v);
    const ProjSet& ps = W.proj[v];
    NodeArgs a{};
    a.h = p.h; a.aggv = av; a.n_rows = p.N; a.do_mlp = v > 0;
    if (v > 0) { a.b3 = W.gcl[v - 1].n0.b; a.b4 = W.gcl[v - 1].n2.b; }
    a.bp = ps.b_half; a.pq = reinterpret_cast<__half*>(p.pq); a.ldp = ps.lin.out; a.n_blocks = ps.lin.out / 256;
    a.row_block = ps.off_coord >= 0 ? ps.off_coord / 256 : -1; a.n_moving = h->joint ? p.N : p.Np;
    a.trace = (h->trace && h->trace_kernel == 1 && (h->trace_v < 0 || h->trace_v == v)) ? h->trace : nullptr;
    a.dbg = h->dbg;
    a.fast_silu = (h->precision == DP_F16_FAST || h->precision == DP_F16_FAST32) ? 1 : 0;
    a.range_flag = p.nan_flag + 2;
    DP_CHECK(h->tc->node[v].n_panels == (a.do_mlp ? 12 : 0) + 4 * a.n_blocks, DP_ERR_STATE, "tc node phase %d: image / shape mismatch", v);
    if (p.N <= 0) return DP_OK;
    NodeTiling nt;
    node_tiling(h, p.N, p.Np, &nt);
    a.tp = nt.tp; a.sp = nt.sp; a.stride = nt.stride; a.n_mma = (nt.stride + 15) / 16 * 16; a.trace_cta = h->trace_cta;
    const int grid = nt.grid;
    // h as 16-bit tile images between the launches of one evaluation (single-CTA kernel; DIFFPHAR_NODE_H16=0: fp32 staging)
    const bool images = h->node_h16 && !h->node_pair && !h->node_mc && p.h16;
    a.h16_in = (images && v > 0) ? p.h16 : nullptr;
    a.h16_out = (images && v + 1 < (int)h->tc->node.size()) ? p.h16 : nullptr;
    const unsigned char* img = h->tc->node[v].img[fmt];
    if (h->node_pair) {
        // cluster of two CTAs per pair of tiles (cta_group::2): an odd tile count gets one empty tile
        const int grid2 = (grid + 1) & ~1, smem2 = (int)sizeof(NodeSmem2) + 1024;
        if (fmt == tc::FMT_BF16) DP_CUDA(launch_kernel(h->pdl, node_pair_kernel<tc::FMT_BF16>, dim3(grid2), dim3(THREADS), smem2, st, a, img));
        else DP_CUDA(launch_kernel(h->pdl, node_pair_kernel<tc::FMT_F16>, dim3(grid2), dim3(THREADS), smem2, st, a, img));
        h->launches += 1;
        DP_CUDA(cudaGetLastError());
        return DP_OK;
    }
    const int smem = (int)sizeof(NodeSmem) + 1024;
    if (h->node_mc && grid > 1) {
        // clusters of two CTAs sharing one multicast weight stream; an odd tile count gets one empty tile
        const int grid2 = (grid + 1) & ~1;
        if (fmt == tc::FMT_BF16) DP_CUDA(launch_kernel_cluster(h->pdl, 2, node_tc_kernel<tc::FMT_BF16, true>, dim3(grid2), dim3(THREADS), smem, st, a, img));
        else DP_CUDA(launch_kernel_cluster(h->pdl, 2, node_tc_kernel<tc::FMT_F16, true>, dim3(grid2), dim3(THREADS), smem, st, a, img));
        h->launches += 1;
        DP_CUDA(cudaGetLastError());
        return DP_OK;
    }
    if (fmt == tc::FMT_BF16) DP_CUDA(launch_kernel(h->pdl, node_tc_kernel<tc::FMT_BF16, false>, dim3(grid), dim3(THREADS), smem, st, a, img));
    else DP_CUDA(launch_kernel(h->pdl, node_tc_kernel<tc::FMT_F16, false>, dim3(grid), dim3(THREADS), smem, st, a, img));
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
