// K2 / K3 — the fused edge kernel on the 5th-generation tensor cores (DP_BF16 / DP_F16 / DP_F16_FAST; arithmetic
// modes: tc_common.cuh MODE_*; work split over CTAs and the atomic-free segmented sum: common.cuh "segmented sum").
//
//   message mode : GCL.edge_model + attention gate + CSR segmented sum   (egnn_new.py:31-52, 276-285)
//   coord mode   : EquivariantUpdate.coord_mlp up to its per-edge scalar  (egnn_new.py:87-91)
//
// Per 64-edge tile:   X = SiLU(Pa[row] + Pb[col] + r2 wr + d0 wd)            first layer (factored: the two
//                                                                            H x H products are per NODE, tc_node.cu)
//                     D[256 ch, 64 edges] = W2[256, 256] . X^T              tcgen05.mma, fp32 accumulators in TMEM
//                     m = SiLU(D + b2);  g = sigmoid(wa . m + ba);  agg[row] += g m
//
// Orientation ("channels on lanes"): A = the nn.Linear weight exactly as stored ([out,in] row-major ==
// K-major), B = the activation tile.  The accumulator has the OUTPUT CHANNEL on the TMEM lane and the
// edges along TMEM columns, so the CSR segmented sum is a run of register FMAs inside one thread —
// no atomics, no shuffles — and agg stores are 128 B coalesced.
//
// Warp-specialised, persistent (one CTA per SM, tiles round-robin), 28 warps = 7 per SM sub-partition:
//   warps 0-15  epilogue : four independent groups of 4 warps; group g owns edges 16 g .. 16 g + 15 of every tile
//                          (= one segmented-sum unit, UNIT_TC).  Warp (g, q) reads TMEM lanes 32 q .. +32 of BOTH
//                          accumulator halves, so a thread holds channels c and c + 128 for 16 edges: SiLU, gate
//                          (in-thread pre-add of the two channels, one smem-transposed reduce per warp, 4-warp named
//                          barrier per group), segmented sum as two independent FMA chains, stores.
//   warps 16-23 producer : each warp owns 8 edges of the tile: gathers the pre-projected f16 rows (Pa once per
//                          CSR row run, Pb per edge, one 128-bit load per lane and row) through an 8-slot register
//                          pipeline, first layer, swizzled K-major B tile, fence.proxy.async, arrive on full[stage].
//   warp  24    MMA      : (warps 25-27 only donate registers, setmaxnreg) waits full[stage] / tmem_empty[acc]; one
//                          elected lane issues 32 tcgen05.mma (M=128, N=64, K=16) per tile and commits to
//                          x_empty[stage] + tmem_full[acc].
// Second-layer weights are the A operand and stay resident in TENSOR MEMORY for the whole kernel (256
// columns: 2 M-halves x 128 columns of packed 16-bit pairs, written once per CTA by the epilogue warps with
// tcgen05.st), so an MMA reads only its 2 KB B tile from shared memory (the SS form read 6 KB: the
// shared-memory port was the MMA's limit).  4 activation stages x 32 KB, 2 accumulator stages x 128 columns.
#include "tc_common.cuh"

namespace {
using namespace tc;

#ifndef EDGE_PRODUCER_SLEEP_NS
#define EDGE_PRODUCER_SLEEP_NS 64
#endif
#ifndef EDGE_EPILOGUE_SLEEP_NS
#define EDGE_EPILOGUE_SLEEP_NS 64
#endif
constexpr int TILE = 64;                          // edges per tile (UMMA N)
constexpr int X_PANEL_BYTES = TILE * 128;         // 8 KB
constexpr int X_TILE_BYTES = 4 * X_PANEL_BYTES;   // 32 KB: 64 edges x 256 K x 2 B
constexpr int N_XS = 4;                           // activation stages
constexpr int N_TS = 2;                           // accumulator stages
constexpr int N_MS = 4;                           // metadata slots (EDGE_PREP_WARP)
constexpr int TS_COLS = 2 * TILE;                 // TMEM columns per accumulator stage
constexpr int W_COLS = 256;                       // TMEM columns of the resident weights: half hh at hh * 128, K pair j at column j
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 16, PRO_WARPS = 8, EPI_GROUPS = EPI_WARPS / 4;
constexpr int GROUP_EDGES = TILE / EPI_GROUPS;    // 16 = UNIT_TC
static_assert(GROUP_EDGES == UNIT_TC, "an epilogue group owns exactly one segmented-sum unit");
constexpr int MMA_WARP = EPI_WARPS + PRO_WARPS;
constexpr int THREADS = (EPI_WARPS + PRO_WARPS + 4) * 32;   // 7 warpgroups: 4 epilogue, 2 producer, 1 MMA (+3 idle warps)
// Registers are allocated per SM sub-partition (16384 each, 7 warps per sub-partition here), so the launch
// gets 72 per thread; setmaxnreg then rebalances.  4 Re + 2 Rp + Rm <= 7 x 72: the CTA pool only holds what
// its own warps released.
// The packed-f16 producer would fit in 88 registers (72 for the epilogue warps then): measured 1 % slower than 104 / 64
// (scripts/gpu_env_ab.sh with -DEDGE_REGS_PACKED_* builds), so every mode uses the same split.
#ifndef EDGE_PREP_WARP
#define EDGE_PREP_WARP 0                           // 1: a warp of the MMA warpgroup prepares the tiles' edge metadata (row, col, packed
                                                  // (r2, d0)) in shared memory, a few tiles ahead; 0: every producer warp fetches its own
#endif
#ifndef EDGE_PA_AHEAD
#define EDGE_PA_AHEAD 1                            // how many edges ahead the producer requests the Pa row of a new CSR row run (1 or 2)
#endif
#ifndef EDGE_REGS_MMA
#define EDGE_REGS_MMA (EDGE_PREP_WARP ? 40 : 32)
#endif
constexpr int REGS_MMA = EDGE_REGS_MMA;                      // spills ~20 registers around the once-per-launch weight fill (tcgen05.cp descriptors): harmless;
                                                  // 40 (no spill) left the producers' setmaxnreg.inc no slack and measured no faster
#ifndef EDGE_REGS_PACKED_PRODUCER
#define EDGE_REGS_PACKED_PRODUCER 104
#define EDGE_REGS_PACKED_EPILOGUE 64
#endif
__host__ __device__ constexpr int regs_producer(int mode) { return (mode == MODE_F16P || mode == MODE_F16Q) ? EDGE_REGS_PACKED_PRODUCER : 104; }
__host__ __device__ constexpr int regs_epilogue(int mode) { return (mode == MODE_F16P || mode == MODE_F16Q) ? EDGE_REGS_PACKED_EPILOGUE : 64; }
static_assert(4 * regs_epilogue(MODE_F16P) + 2 * regs_producer(MODE_F16P) + REGS_MMA <= 7 * 72, "register pool");
static_assert(4 * regs_epilogue(MODE_BF16) + 2 * regs_producer(MODE_BF16) + REGS_MMA <= 7 * 72, "register pool");
constexpr int RED_STRIDE = 20;                    // floats per channel row of the transposed-reduce buffer (16 edges + pad)

struct EdgeSmem {                                 // offsets from a 1024-aligned base
    unsigned char x[N_XS][X_TILE_BYTES];          // 128 KB
    float red[EPI_WARPS][32 * RED_STRIDE];        // 40 KB: per-warp [channel pair][16 edges (+4 pad)]
    float part[2][EPI_WARPS][GROUP_EDGES];        // per-warp partial gate sums, double-buffered over tiles
    float gate[EPI_WARPS][GROUP_EDGES];
    float cstash[EPI_GROUPS][GROUP_EDGES][12];
    // EDGE_PREP_WARP: metadata of N_MS tiles, written by the prep warp, read by the producers (broadcast LDS)
    struct Meta { int row[TILE]; int col[TILE]; uint32_t rd[TILE]; float r2[TILE]; float d0[TILE]; } meta[4];
    unsigned long long bar_mfull[4], bar_mempty[4];        // coordinate mode with the in-kernel finish: per edge (coord_diff[3], x_row[3], row, rowptr[row], rowptr[row + 1])
    unsigned long long bar_w;
    unsigned long long bar_wload, bar_wdone;          // a.tma_fill: weight panels landed in shared memory / copied to tensor memory
    unsigned long long bar_full[N_XS], bar_xempty[N_XS];
    unsigned long long bar_tfull[N_TS], bar_tempty[N_TS];
    uint32_t tmem_holder;
};

// D[256 x 64] = W[256 x 256] * X[64 x 256]^T  as 2 (M halves) x 4 (panels) x 4 (K steps) MMAs, A from TMEM
__device__ __forceinline__ void issue_tile_mma(uint32_t tmem_d, uint32_t tmem_w, uint32_t x_base, uint32_t idesc)
{
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
#pragma unroll
        for (int ks = 0; ks < PANEL_K / 16; ++ks) {
            const uint64_t bdesc = make_desc(x_base + kp * X_PANEL_BYTES + ks * 32);
            const uint32_t acc = (kp > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
                umma_f16_ts(tmem_d + hh * TILE, tmem_w + hh * 128 + kp * 32 + ks * 8, bdesc, idesc, acc);
        }
    }
}

__device__ __forceinline__ void unpack8(const float4& a, const float4& b, float (&v)[8])
{
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// debug timeline: role 0 = producer warp 0, 1 = MMA thread, 2 = epilogue warp 0; one CTA (phar-row lanes: CTA 0; pocket rows: 100)
#ifndef TRACE_MARKS
#define TRACE_MARKS 1                             // 0: the traced build keeps only the per-CTA global-timer stamps (production code otherwise)
#endif
#ifndef TRACE_EDGES
#define TRACE_EDGES 0                             // 1: per-edge marks of the first two edges of a tile (they serialise the edges: coarse marks then lie)
#endif
#ifndef TRACE_CTA
#define TRACE_CTA 100
#endif
template <bool TRACE>
__device__ __forceinline__ void trace_mark_t(long long* trace, int role, int it, int slot)
{
    if (TRACE && TRACE_MARKS) { if (blockIdx.x == TRACE_CTA && it < 64) trace[(role * 64 + it) * 16 + slot] = clock64(); }
}
#define trace_mark(tr, role, it, slot) trace_mark_t<TRACE>(tr, role, it, slot)

// COORD: the coordinate mode is its own instantiation (a.coord == COORD), so the message kernel — the hot one — carries
// neither its branches nor the registers of the in-kernel coordinate finish.
template <int MODE, bool TRACE, bool COORD>
__global__ void __launch_bounds__(THREADS, 1) edge_tc_kernel(EdgeArgs a, const unsigned char* __restrict__ w_img)
{
    constexpr int FMT = fmt_of_mode(MODE);            // operand format of the MMA
    constexpr int SFMT = silu_of_mode(MODE);          // SiLU flavour of the fp32 paths
    constexpr bool PACKED = MODE == MODE_F16P || MODE == MODE_F16Q;   // first layer in f16x2
    extern __shared__ unsigned char smem_raw[];
    // offset arithmetic on the shared array itself (not through an integer cast): the compiler keeps the shared
    // state space and emits STS / LDS with 32-bit addresses instead of generic ST.E / LD.E
    unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(base);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == MMA_WARP * 32) trace_mark(a.trace, 1, 62, 0);                              // kernel entry
    if (TRACE && tid == MMA_WARP * 32 && blockIdx.x < 376) {                              // every CTA: global-timer stamp at entry (ns)
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[(64 + 16 + (blockIdx.x >> 3)) * 16 + 2 * (blockIdx.x & 7)] = (long long)gt;
    }

    // The edge count (and, for the round-robin split, the producers' first-tile metadata: loaded speculatively — the
    // arrays hold ecap entries — and masked once E has arrived) is requested before anything else, so that its L2
    // round trip overlaps the prologue (barrier init, tensor-memory allocation): 1 k cycles of every launch.  The
    // grid dependency therefore sits at the very top (a no-op without programmatic dependent launch).
    pdl_launch_dependents();
    pdl_wait();                                       // from here on: data written by earlier kernels of the step
    int s_row = 0, s_col = 0; float s_d0 = 0.f;
#if EDGE_PREP_WARP
    int s_row2 = 0, s_col2 = 0; float s_d02 = 0.f;                      // the prep warp: edges lane and lane + 32 of the CTA's first tile
    if (!a.contig && wid == MMA_WARP + 1) {
        const int e = (int)blockIdx.x * TILE + lane;
        if (e < a.ecap) { s_row = a.erow[e]; s_col = a.ecol[e]; s_d0 = a.d0[e]; }
        if (e + 32 < a.ecap) { s_row2 = a.erow[e + 32]; s_col2 = a.ecol[e + 32]; s_d02 = a.d0[e + 32]; }
    }
#else
    if (!a.contig && wid >= EPI_WARPS && wid < MMA_WARP && lane < 8) {
        const int e = (int)blockIdx.x * TILE + 8 * (wid - EPI_WARPS) + lane;
        if (e < a.ecap) { s_row = a.erow[e]; s_col = a.ecol[e]; s_d0 = a.d0[e]; }
    }
#endif
    const int E = *a.n_edges;

    // ---- prologue
    if (tid == 0) {
        mbar_init(smem_u32(&s.bar_w), EPI_WARPS);              // every epilogue warp fills its share of tensor memory
        mbar_init(smem_u32(&s.bar_wload), 1); mbar_init(smem_u32(&s.bar_wdone), 1);
        for (int i = 0; i < N_XS; ++i) { mbar_init(smem_u32(&s.bar_full[i]), PRO_WARPS); mbar_init(smem_u32(&s.bar_xempty[i]), 1); }
        for (int i = 0; i < N_TS; ++i) { mbar_init(smem_u32(&s.bar_tfull[i]), 1); mbar_init(smem_u32(&s.bar_tempty[i]), EPI_WARPS); }
        for (int i = 0; i < N_MS; ++i) { mbar_init(smem_u32(&s.bar_mfull[i]), 1); mbar_init(smem_u32(&s.bar_mempty[i]), PRO_WARPS); }
        fence_barrier_init();
    }
    if (wid == MMA_WARP) tmem_alloc(smem_u32(&s.tmem_holder), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_w = s.tmem_holder;                   // resident weights
    const uint32_t tmem_base = tmem_w + W_COLS;              // accumulator stages
    static_assert(W_COLS + N_TS * TS_COLS <= TMEM_COLS, "tensor memory budget");
    // Resident weights, load / store path (a.tma_fill = 0; the default goes through TMA + tcgen05.cp in the MMA warp):
    // thread (q, lane) of epilogue group gi owns row c = 128 (gi & 1) + 32 q + lane of the [out][in] matrix = TMEM
    // lane 32 q + lane of M-half gi & 1, K panels 2 (gi >> 1) and 2 (gi >> 1) + 1.  The 128 K elements are read from
    // the swizzled panel image (chunk c of a row sits at chunk c ^ (row % 8)) and stored as 64 packed 32-bit columns.
    // All 148 CTAs pull the same 128 KB at once (~19 MB through L2) and 16 warps of weight loads sit in the SM's
    // load queues ahead of the producers' first gathers: 9 k cycles of pipeline fill.
    auto fill_weights = [&]() {
        const int q = wid & 3, gi = wid >> 2;
        const int mh = gi & 1, r = 128 * mh + 32 * q + lane;
#pragma unroll 1
        for (int kp = 2 * (gi >> 1); kp < 2 * (gi >> 1) + 2; ++kp) {
            const unsigned char* src = w_img + (size_t)kp * W_PANEL_BYTES + (size_t)r * 128;
            uint32_t wa[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 t = *reinterpret_cast<const uint4*>(src + ((c ^ (r & 7)) << 4));
                wa[4 * c] = t.x; wa[4 * c + 1] = t.y; wa[4 * c + 2] = t.z; wa[4 * c + 3] = t.w;
            }
            tmem_st32(tmem_w + ((uint32_t)(32 * q) << 16) + mh * 128 + kp * 32, wa);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.bar_w));
    };
    // Work split (common.cuh, "segmented sum of the tcgen05 edge kernel"):
    //   a.contig = 0 (Calpha pockets): tile t = 64 consecutive edges, CTA t mod gridDim (its first tile's metadata was
    //     requested speculatively at the top of the kernel);
    //   a.contig = 1 (full-atom pockets): the U 16-edge units are cut into 4 x gridDim contiguous balanced lanes;
    //     epilogue group g (and the two producer warps that feed it) walks lane 4 blockIdx + g in order, one unit per
    //     tile, and the running sum of a CSR row stays in the group's registers from tile to tile.
    const unsigned U = (unsigned)((E + UNIT_TC - 1) / UNIT_TC), L = 4u * gridDim.x;
    // tiles of this CTA.  Lanes: its longest lane; lane lengths are floor(U / L) or ceil(U / L), so ceil(U / L) is exact
    // for every CTA that has a long lane and one (empty, skipped by has_unit) tile too many for the others — which
    // finish no later than the CTAs that need it.  Idle CTAs fall through with 0.
    int my_tiles;
    if (a.contig) my_tiles = lane_first_unit(4u * blockIdx.x + 4u, U, L) > lane_first_unit(4u * blockIdx.x, U, L) ? (int)((U + L - 1u) / L) : 0;
    else my_tiles = max(0, ((E + TILE - 1) / TILE - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
    // group g's unit of tile `it` is unit_base(g) + it * unit_step, and exists while it < unit_count(g)
    const int unit_step = a.contig ? 1 : 4 * (int)gridDim.x;
    auto unit_base = [&](int g) { return a.contig ? (int)lane_first_unit(4u * blockIdx.x + g, U, L) : 4 * (int)blockIdx.x + g; };
    auto unit_count = [&](int g) { return a.contig ? (int)lane_first_unit(4u * blockIdx.x + g + 1u, U, L) - unit_base(g) : my_tiles; };

    if (wid >= MMA_WARP) {
        // ================================ MMA issuer ================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA));
        if (wid == MMA_WARP) {
            // the whole warp runs the loop (uniform control flow); one elected lane issues
            const uint32_t bar_w = smem_u32(&s.bar_w);
            constexpr uint32_t idesc = make_idesc(FMT, 128, TILE);
            const uint32_t tw = warp_uniform(tmem_w), td = warp_uniform(tmem_base);
            const uint32_t x0 = warp_uniform(smem_u32(s.x[0]));
            if (lane == 0) trace_mark(a.trace, 1, 63, 0);
            if (a.tma_fill && my_tiles > 0) {
                // Resident weights without the load / store units: four bulk copies bring the 128 KB image into the
                // activation stages 1-3 and the gate-reduce buffer (all idle until the first tile is through), 32
                // tcgen05.cp move it on to tensor memory (128 rows x 16 K per copy: the A-operand layout of the .ts
                // MMAs), and the commit frees the buffers.  The producers' first loads no longer queue behind 16 warps
                // of weight loads.
                const uint32_t wl = smem_u32(&s.bar_wload);
                const uint32_t dst[4] = {x0 + 1 * X_TILE_BYTES, x0 + 2 * X_TILE_BYTES, x0 + 3 * X_TILE_BYTES, warp_uniform(smem_u32(s.red))};
                if (elect_one()) {
                    mbar_expect_tx(wl, 4 * W_PANEL_BYTES);
#pragma unroll
                    for (int kp = 0; kp < 4; ++kp) bulk_g2s(dst[kp], w_img + (size_t)kp * W_PANEL_BYTES, W_PANEL_BYTES, wl);
                }
                __syncwarp();
                mbar_wait(wl, 0);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kp = 0; kp < 4; ++kp)
#pragma unroll
                        for (int ks = 0; ks < PANEL_K / 16; ++ks)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh)
                                tmem_cp_128x256b(tw + hh * 128 + kp * 32 + ks * 8, make_desc(dst[kp] + hh * (128 * 128) + ks * 32));
                    umma_commit(smem_u32(&s.bar_wdone));
                }
                __syncwarp();
            } else if (my_tiles > 0) {
                mbar_wait(bar_w, 0);                                   // weights are in tensor memory
            }
            if (lane == 0) trace_mark(a.trace, 1, 63, 1);
            for (int it = 0; it < my_tiles; ++it) {
                const int xs = it % N_XS, ts = it % N_TS;
                if (lane == 0) trace_mark(a.trace, 1, it, 0);
                mbar_wait(smem_u32(&s.bar_full[xs]), (it / N_XS) & 1);
                if (lane == 0) trace_mark(a.trace, 1, it, 1);
                mbar_wait(smem_u32(&s.bar_tempty[ts]), ((it / N_TS) & 1) ^ 1);
                if (lane == 0) trace_mark(a.trace, 1, it, 2);
                tc_fence_after();
                if (elect_one()) {
                    issue_tile_mma(td + ts * TS_COLS, tw, x0 + xs * X_TILE_BYTES, idesc);
                    umma_commit(smem_u32(&s.bar_xempty[xs]));
                    umma_commit(smem_u32(&s.bar_tfull[ts]));
                }
                __syncwarp();
                if (lane == 0) trace_mark(a.trace, 1, it, 3);
            }
        }
#if EDGE_PREP_WARP
        else if (wid == MMA_WARP + 1) {
            // ================================ metadata prep ================================
            // Per tile and edge slot j (group j / 16, edge j % 16 of the group's unit): row, col and the two scalar edge
            // features — d0 from the graph build, r2 recomputed when an endpoint moved (coord2diff, egnn_new.py:265-268) —
            // as one packed f16x2 word in the packed modes.  One warp does it for all 64 edges (two per lane), up to N_MS
            // tiles ahead of the producers, instead of each of the 8 producer warps for its own 8 edges on 8 lanes: ~150
            // index / load instructions per producer warp and tile leave the warps that set the tile period at Calpha size.
            // Tile my_tiles (no edges: all zero) is produced too, so the producers read "the next tile" unconditionally.
            float rd_max = 0.f;                                                  // largest r2 / d0 packed to f16 (range guard)
            // Software pipeline, one L2 round trip per tile in steady state: (row, col, d0) of tile it + 1 are requested
            // before tile it's coordinates are consumed; the coordinates of moved endpoints go through unconditional loads
            // of a valid address (no branch), both halves of the tile together.
            int nr[2], nc[2]; float nd[2];                                       // tile `it`: requested one iteration earlier
            auto request = [&](int it, int (&r)[2], int (&c)[2], float (&d)[2]) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int j = lane + 32 * hh, g = j >> 4;
                    const int e = (unit_base(g) + it * unit_step) * UNIT_TC + (j & 15);
                    r[hh] = 0; c[hh] = 0; d[hh] = 0.f;
                    if (it < unit_count(g) && e < E) { r[hh] = a.erow[e]; c[hh] = a.ecol[e]; d[hh] = a.d0[e]; }
                }
            };
            if (a.contig) {
                request(0, nr, nc, nd);
            } else {                                                             // requested at kernel entry, masked now that E is known
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int j = lane + 32 * hh, g = j >> 4;
                    const bool valid = 0 < unit_count(g) && unit_base(g) * UNIT_TC + (j & 15) < E;
                    nr[hh] = valid ? (hh ? s_row2 : s_row) : 0; nc[hh] = valid ? (hh ? s_col2 : s_col) : 0; nd[hh] = valid ? (hh ? s_d02 : s_d0) : 0.f;
                }
            }
            for (int it = 0; it <= my_tiles; ++it) {
                const int ms = it % N_MS;
                int r[2] = {nr[0], nr[1]}, c[2] = {nc[0], nc[1]};
                float d0[2] = {nd[0], nd[1]};
                float xr[2][3], xc[2][3];
                bool moving[2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    moving[hh] = r[hh] < a.n_moving || c[hh] < a.n_moving;      // an endpoint moved since the graph build
                    const int ri = moving[hh] ? 3 * r[hh] : 0, ci = moving[hh] ? 3 * c[hh] : 0;
                    xr[hh][0] = a.x[ri]; xr[hh][1] = a.x[ri + 1]; xr[hh][2] = a.x[ri + 2];
                    xc[hh][0] = a.x[ci]; xc[hh][1] = a.x[ci + 1]; xc[hh][2] = a.x[ci + 2];
                }
                if (it < my_tiles) request(it + 1, nr, nc, nd);
                else { nr[0] = nr[1] = nc[0] = nc[1] = 0; nd[0] = nd[1] = 0.f; }
                if (it >= N_MS) mbar_wait_relaxed(smem_u32(&s.bar_mempty[ms]), ((it / N_MS) & 1) ^ 1);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int j = lane + 32 * hh;
                    float r2 = d0[hh];
                    if (moving[hh]) {
                        const float dx = xr[hh][0] - xc[hh][0], dy = xr[hh][1] - xc[hh][1], dz = xr[hh][2] - xc[hh][2];
                        r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    }
                    s.meta[ms].row[j] = r[hh]; s.meta[ms].col[j] = c[hh];
                    if (PACKED) {
                        const __half2 t = __floats2half2_rn(fminf(r2, 60000.f), fminf(d0[hh], 60000.f));
                        s.meta[ms].rd[j] = *reinterpret_cast<const uint32_t*>(&t);
                        rd_max = fmaxf(rd_max, fmaxf(r2, d0[hh]));
                    } else {
                        s.meta[ms].r2[j] = r2; s.meta[ms].d0[j] = d0[hh];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s.bar_mfull[ms]));
            }
            // beyond f16's range the clamp makes an edge differ from the reference (only reachable without a cutoff):
            // flagged once per launch, the caller re-runs in a mode with fp32 edge features (dp_flags.f16_range)
            if (rd_max > 60000.f) atomicOr(a.range_flag, 2);
        }
#endif
    } else if (wid >= EPI_WARPS) {
        // ================================ producer ================================
        // Warp pw owns edges 8 pw .. 8 pw + 7 of every tile; a lane owns 8 channels (128-bit loads / stores).
        //   * pq is pre-scaled by 1/2 (api.cu), wr / wd are halved here: hv = Pa' + Pb' + r2 wr' + d0 wd' feeds
        //     SiLU(2 hv) = hv + hv tanh(hv) directly (FADD, 2 FFMA, MUFU, FFMA per element).
        //   * Pb rows (512 B of f16) stream through an 8-slot register pipeline (no L1 allocation): the slot of
        //     edge i is refilled with edge i of the NEXT tile right after edge i is computed, so a gather has a
        //     whole tile to land.  Pa + Pb is added in f16x2, widened once, the rest of the layer runs in fp32.
        //   * Pa changes once per CSR row run: a predicated 128-bit reload (an L1 hit when a neighbouring warp's chunk
        //     of the same row came first, else L2).  An explicit L1 prefetch of the next tile's rows (cp.async touch, one
        //     loop iteration per distinct row) cost more serial issue time than the misses it removed: -2.4 % step time
        //     without it (profiles/r04d_producer_dbg.txt).
        //   * (row, col, d0, r2) of the 8 edges live on lanes 0-7, fetched one tile ahead.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(regs_producer(MODE)));
#if EDGE_PREP_WARP
        const int pw = wid - EPI_WARPS;
        float wr[8], wd[8];
        unpack8(*reinterpret_cast<const float4*>(a.wr + 8 * lane), *reinterpret_cast<const float4*>(a.wr + 8 * lane + 4), wr);
        unpack8(*reinterpret_cast<const float4*>(a.wd + 8 * lane), *reinterpret_cast<const float4*>(a.wd + 8 * lane + 4), wd);
#pragma unroll
        for (int k = 0; k < 8; ++k) { wr[k] *= 0.5f; wd[k] *= 0.5f; }
        // packed modes: the same eight halved weights as four f16x2 pairs (the fp32 copies are dead then)
        __half2 wr2[4], wd2[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            wr2[k2] = __floats2half2_rn(wr[2 * k2], wr[2 * k2 + 1]);
            wd2[k2] = __floats2half2_rn(wd[2 * k2], wd[2 * k2 + 1]);
        }
        const uint32_t ldp_b = 2u * (uint32_t)a.ldp;                                     // row stride in bytes (f16 rows)
        const __half* pq = reinterpret_cast<const __half*>(a.p);                         // f16 rows, pre-scaled by 1/2 (tc_node.cu)
        const __half* pa_base = pq + a.off_a + 8 * lane;
        const __half* pb_base = pq + a.off_b + 8 * lane;
        // (row, col, packed (r2, d0)) of every tile come from the prep warp through shared memory (broadcast LDS); slots
        // without an edge hold (0, 0, 0): finite garbage in columns nobody reads.
        const int e8 = 8 * pw;                                                           // this warp's first edge slot of a tile
        mbar_wait_relaxed(smem_u32(&s.bar_mfull[0]), 0);
        uint4 pb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)                                                      // fill the pipeline: the first tile's gathers go out first
            pb[u] = ldg_na_u4(row_ptr(pb_base, (uint32_t)s.meta[0].col[e8 + u], ldp_b));
        // Pa of the current CSR row run (8 halves).  It changes once per run; the reload for the NEXT edge's row is issued
        // before this edge's arithmetic, so an L2 round trip hides behind one edge of work instead of sitting in the chain
        int cur_row = s.meta[0].row[e8];
        uint4 cur = ldg_u4(row_ptr(pa_base, (uint32_t)cur_row, ldp_b));
        unsigned char* const x_gen = s.x[0] + (((lane >> 3) << 13) | (pw << 10));     // K panel of the lane's chunk, row 8 pw
        const int l74 = (lane & 7) << 4;

        for (int it = 0; it < my_tiles; ++it) {
            const int xs = it % N_XS, cs = it % N_MS, ns = (it + 1) % N_MS;
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 0);
            mbar_wait_relaxed(smem_u32(&s.bar_mfull[ns]), ((it + 1) / N_MS) & 1);        // the next tile's rows / columns (tile my_tiles: zeros)
            mbar_wait_relaxed<EDGE_PRODUCER_SLEEP_NS>(smem_u32(&s.bar_xempty[xs]), ((it / N_XS) & 1) ^ 1);   // up to 3 tiles ahead: a late wake-up costs nothing
            if (a.tma_fill && it == 1) mbar_wait_relaxed(smem_u32(&s.bar_wdone), 0);    // stages 1-3 carried the weight panels
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 1);
            unsigned char* const xt = x_gen + xs * X_TILE_BYTES;
            const int* const rows_c = s.meta[cs].row + e8;
            const int* const rows_n = s.meta[ns].row + e8;
            const int* const cols_n = s.meta[ns].col + e8;
            // one basic block for the 8 edges: no branches, so the scheduler overlaps neighbouring edges
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int next_row = i < 7 ? rows_c[i + 1] : rows_n[0];
                // loaded into fresh registers and selected afterwards: the reloads of a tile do not depend on each other
                const bool new_run = next_row != cur_row && !(a.dbg & 2);                 // next edge starts a new row run
                uint4 fresh;                                                             // only read under new_run
                ldg4_if_noinit(fresh, row_ptr(pa_base, (uint32_t)next_row, ldp_b), new_run);
                const uint4 nxt = make_uint4(new_run ? fresh.x : cur.x, new_run ? fresh.y : cur.y, new_run ? fresh.z : cur.z, new_run ? fresh.w : cur.w);
                const uint32_t ca[4] = {cur.x, cur.y, cur.z, cur.w}, cb[4] = {pb[i].x, pb[i].y, pb[i].z, pb[i].w};
                uint32_t o[4];
                if (PACKED) {
                    // two channels per instruction: (Pa' + Pb') + r2 wr' + d0 wd' and SiLU(2 hv) = hv + hv tanh(hv) in
                    // f16x2; (r2, d0) travel as one packed word, the halves are broadcast by the operand selectors
                    const uint32_t rd = s.meta[cs].rd[e8 + i];
                    const __half2 r2h = __low2half2(*reinterpret_cast<const __half2*>(&rd));
                    const __half2 d0h = __high2half2(*reinterpret_cast<const __half2*>(&rd));
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&ca[k2]), *reinterpret_cast<const __half2*>(&cb[k2]));
                        const __half2 hv = __hfma2(d0h, wd2[k2], __hfma2(r2h, wr2[k2], sum));
                        __half2 y2;
                        if (MODE == MODE_F16P) {
                            y2 = __hfma2(hv, tanh_approx_h2(hv), hv);
                        } else {
                            const float2 f = __half22float2(hv);
                            y2 = __floats2half2_rn(fmaf(f.x, tanh_approx(f.x), f.x), fmaf(f.y, tanh_approx(f.y), f.y));
                        }
                        o[k2] = *reinterpret_cast<const uint32_t*>(&y2);
                    }
                } else {
                    const float r2 = s.meta[cs].r2[e8 + i];
                    const float d0 = s.meta[cs].d0[e8 + i];
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&ca[k2]), *reinterpret_cast<const __half2*>(&cb[k2]));
                        const float2 f = __half22float2(sum);
                        o[k2] = pack2<FMT>(silu_half<SFMT>(fmaf(d0, wd[2 * k2], fmaf(r2, wr[2 * k2], f.x))),
                                           silu_half<SFMT>(fmaf(d0, wd[2 * k2 + 1], fmaf(r2, wr[2 * k2 + 1], f.y))));
                    }
                }
                *reinterpret_cast<uint4*>(xt + ((i << 7) | (l74 ^ (i << 4)))) = make_uint4(o[0], o[1], o[2], o[3]);   // row 8 pw + i, chunk (lane % 8) ^ (row % 8)
                // refill the slot with the same edge of the next tile
                pb[i] = ldg_na_u4(row_ptr(pb_base, (uint32_t)cols_n[i], ldp_b));
                cur = nxt; cur_row = next_row;
            }
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 2);
            fence_proxy_async();                                                        // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) { mbar_arrive(smem_u32(&s.bar_full[xs])); mbar_arrive(smem_u32(&s.bar_mempty[cs])); }
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 6);
        }
#else
        const int pw = wid - EPI_WARPS;
        const int unit0 = unit_base(pw >> 1), n_units = unit_count(pw >> 1);             // the units of this warp's epilogue group
        const int e_off = 8 * (pw & 1) + lane;                                           // its half of the unit's 16 edges
        float wr[8], wd[8];
        unpack8(*reinterpret_cast<const float4*>(a.wr + 8 * lane), *reinterpret_cast<const float4*>(a.wr + 8 * lane + 4), wr);
        unpack8(*reinterpret_cast<const float4*>(a.wd + 8 * lane), *reinterpret_cast<const float4*>(a.wd + 8 * lane + 4), wd);
#pragma unroll
        for (int k = 0; k < 8; ++k) { wr[k] *= 0.5f; wd[k] *= 0.5f; }
        // packed modes: the same eight halved weights as four f16x2 pairs (the fp32 copies are dead then)
        __half2 wr2[4], wd2[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            wr2[k2] = __floats2half2_rn(wr[2 * k2], wr[2 * k2 + 1]);
            wd2[k2] = __floats2half2_rn(wd[2 * k2], wd[2 * k2 + 1]);
        }
        const uint32_t ldp_b = 2u * (uint32_t)a.ldp;                                     // row stride in bytes (f16 rows)
        const __half* pq = reinterpret_cast<const __half*>(a.p);                         // f16 rows, pre-scaled by 1/2 (tc_node.cu)
        const __half* pa_base = pq + a.off_a + 8 * lane;
        const __half* pb_base = pq + a.off_b + 8 * lane;
        // Slots without an edge (past E, or past the end of a shorter lane) are processed as edge (0, 0): finite garbage in columns nobody reads.
        int m_row = 0, m_col = 0; float m_r2 = 0.f, m_d0 = 0.f;
        auto load_rc = [&](int it, int& r, int& c, float& d0) {
            const int e = (unit0 + it * unit_step) * UNIT_TC + e_off;
            r = 0; c = 0; d0 = 0.f;
            if (it < n_units && lane < 8 && e < E) { r = a.erow[e]; c = a.ecol[e]; d0 = a.d0[e]; }
        };
        // metadata runs two tiles ahead: m_ = this tile (complete), n_ = next tile (r2 pending), f_ = loading
        auto dist2 = [&](float xr0, float xr1, float xr2, float xc0, float xc1, float xc2) {
            const float dx = xr0 - xc0, dy = xr1 - xc1, dz = xr2 - xc2;
            return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   // coord2diff, egnn_new.py:265-268
        };
        int n_row, n_col; float n_d0;
        if (a.contig) {
            load_rc(0, m_row, m_col, m_d0);
        } else {
            const bool ok = my_tiles > 0 && lane < 8 && unit0 * UNIT_TC + e_off < E;       // the speculative loads were real edges
            m_row = ok ? s_row : 0; m_col = ok ? s_col : 0; m_d0 = ok ? s_d0 : 0.f;
        }
        uint4 pb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)                                                      // fill the pipeline: the first tile's gathers go out first
            pb[u] = ldg_na_u4(row_ptr(pb_base, (uint32_t)__shfl_sync(0xffffffffu, m_col, u), ldp_b));
        load_rc(1, n_row, n_col, n_d0);
        m_r2 = m_d0;
        if (m_row < a.n_moving || m_col < a.n_moving)                                    // an endpoint moved since the graph build
            m_r2 = dist2(a.x[3 * m_row], a.x[3 * m_row + 1], a.x[3 * m_row + 2], a.x[3 * m_col], a.x[3 * m_col + 1], a.x[3 * m_col + 2]);
        // Pa of the current CSR row run (8 halves).  It changes once per run; the reload for the NEXT edge's row is issued
        // before this edge's arithmetic (its row is known from the metadata lanes), so an L2 round trip hides behind one
        // edge of work instead of sitting in the chain
        int cur_row = __shfl_sync(0xffffffffu, m_row, 0);
        uint4 cur = ldg_u4(row_ptr(pa_base, (uint32_t)cur_row, ldp_b));
#if EDGE_PA_AHEAD == 2
        // two edges ahead: `pend` holds the Pa of the NEXT edge's row when that edge starts a new row run (pend_new), requested
        // one edge ago; the reload for the edge after that is issued now.  At Calpha degrees (6.7 edges per row) a warp's 8
        // edges change rows 1.2 times per tile and one edge of work (~450 cycles) does not cover an L2 round trip.
        int row1 = __shfl_sync(0xffffffffu, m_row, 1);
        bool pend_new = row1 != cur_row;
        uint4 pend;
        ldg4_if_noinit(pend, row_ptr(pa_base, (uint32_t)row1, ldp_b), pend_new);
#endif
        unsigned char* const x_gen = s.x[0] + (((lane >> 3) << 13) | (pw << 10));     // K panel of the lane's chunk, row 8 pw
        const int l74 = (lane & 7) << 4;

        float rd_max = 0.f;                                                              // largest r2 / d0 packed to f16 (range guard)
        for (int it = 0; it < my_tiles; ++it) {
            const int xs = it % N_XS;
            int f_row, f_col; float f_d0;
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 0);
            load_rc(it + 2, f_row, f_col, f_d0);
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 3);
            // next tile: coordinates of moved endpoints (unconditional loads of a valid address, no branch) and L1 touch of its Pa rows
            const bool n_moving = n_row < a.n_moving || n_col < a.n_moving;
            const int xr_i = n_moving ? 3 * n_row : 0, xc_i = n_moving ? 3 * n_col : 0;
            const float xr0 = a.x[xr_i], xr1 = a.x[xr_i + 1], xr2 = a.x[xr_i + 2];
            const float xc0 = a.x[xc_i], xc1 = a.x[xc_i + 1], xc2 = a.x[xc_i + 2];
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 4);
            mbar_wait_relaxed<EDGE_PRODUCER_SLEEP_NS>(smem_u32(&s.bar_xempty[xs]), ((it / N_XS) & 1) ^ 1);   // up to 3 tiles ahead: a late wake-up costs nothing
            if (a.tma_fill && it == 1) mbar_wait_relaxed(smem_u32(&s.bar_wdone), 0);    // stages 1-3 carried the weight panels
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 1);
            unsigned char* const xt = x_gen + xs * X_TILE_BYTES;
            uint32_t m_rd = 0u;
            if (PACKED) {                                                               // (r2, d0) of the lane's edge as one f16x2 word
                const __half2 t = __floats2half2_rn(fminf(m_r2, 60000.f), fminf(m_d0, 60000.f));
                m_rd = *reinterpret_cast<const uint32_t*>(&t);
                rd_max = fmaxf(rd_max, fmaxf(m_r2, m_d0));
            }
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 5);
            // one basic block for the 8 edges: no branches, so the scheduler overlaps neighbouring edges
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#if EDGE_PA_AHEAD == 2
                const int row2 = i < 6 ? __shfl_sync(0xffffffffu, m_row, i + 2) : __shfl_sync(0xffffffffu, n_row, i - 6);
                if (TRACE && TRACE_EDGES && i < 2 && pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 7 + 4 * i);
                const bool new2 = row2 != row1 && !(a.dbg & 2);
                uint4 fresh2;                                                            // only read under new2 (as `pend`, two edges on)
                ldg4_if_noinit(fresh2, row_ptr(pa_base, (uint32_t)row2, ldp_b), new2);
                const int next_row = row1;
                const uint4 nxt = make_uint4(pend_new ? pend.x : cur.x, pend_new ? pend.y : cur.y, pend_new ? pend.z : cur.z, pend_new ? pend.w : cur.w);
                pend = fresh2; pend_new = new2; row1 = row2;
#else
                const int next_row = i < 7 ? __shfl_sync(0xffffffffu, m_row, i + 1) : __shfl_sync(0xffffffffu, n_row, 0);
                if (TRACE && TRACE_EDGES && i < 2 && pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 7 + 4 * i);
                // loaded into fresh registers and selected afterwards: the reloads of a tile do not depend on each other
                // (a predicated load straight into a copy of `cur` chains every reload behind the previous one's arrival)
                const bool new_run = next_row != cur_row && !(a.dbg & 2);                 // next edge starts a new row run
#ifndef EDGE_FRESH_INIT
                // the destination registers are NOT initialised (only read under new_run by the selects below): saves the two
                // CS2R per edge that zero a uint4 — message launch 27.1 -> 26.6 us at config 2 (profiles/r06b_ab_summary.txt)
                uint4 fresh;
                ldg4_if_noinit(fresh, row_ptr(pa_base, (uint32_t)next_row, ldp_b), new_run);
#else
                uint4 fresh = make_uint4(0u, 0u, 0u, 0u);
                ldg4_if(fresh, row_ptr(pa_base, (uint32_t)next_row, ldp_b), new_run);
#endif
                const uint4 nxt = make_uint4(new_run ? fresh.x : cur.x, new_run ? fresh.y : cur.y, new_run ? fresh.z : cur.z, new_run ? fresh.w : cur.w);
#endif
                if (TRACE && TRACE_EDGES && i < 2) {                                     // timeline only: when Pa / Pb of this edge have landed
                    uint32_t t0, t1;
                    asm volatile("mov.b32 %0, %1;" : "=r"(t0) : "r"(cur.x));
                    if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 8 + 4 * i);
                    asm volatile("mov.b32 %0, %1;" : "=r"(t1) : "r"(pb[i].x));
                    if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 9 + 4 * i);
                }
                const uint32_t ca[4] = {cur.x, cur.y, cur.z, cur.w}, cb[4] = {pb[i].x, pb[i].y, pb[i].z, pb[i].w};
                uint32_t o[4];
                if (PACKED) {
                    // two channels per instruction: (Pa' + Pb') + r2 wr' + d0 wd' and SiLU(2 hv) = hv + hv tanh(hv) in
                    // f16x2; (r2, d0) travel as one packed word, the halves are broadcast by the operand selectors
                    const uint32_t rd = __shfl_sync(0xffffffffu, m_rd, i);
                    const __half2 r2h = __low2half2(*reinterpret_cast<const __half2*>(&rd));
                    const __half2 d0h = __high2half2(*reinterpret_cast<const __half2*>(&rd));
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&ca[k2]), *reinterpret_cast<const __half2*>(&cb[k2]));
                        const __half2 hv = __hfma2(d0h, wd2[k2], __hfma2(r2h, wr2[k2], sum));
                        __half2 y2;
                        if (MODE == MODE_F16P) {
                            y2 = __hfma2(hv, tanh_approx_h2(hv), hv);
                        } else {
                            const float2 f = __half22float2(hv);
                            y2 = __floats2half2_rn(fmaf(f.x, tanh_approx(f.x), f.x), fmaf(f.y, tanh_approx(f.y), f.y));
                        }
                        o[k2] = *reinterpret_cast<const uint32_t*>(&y2);
                    }
                } else {
                    const float r2 = __shfl_sync(0xffffffffu, m_r2, i);
                    const float d0 = __shfl_sync(0xffffffffu, m_d0, i);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&ca[k2]), *reinterpret_cast<const __half2*>(&cb[k2]));
                        const float2 f = __half22float2(sum);
                        o[k2] = pack2<FMT>(silu_half<SFMT>(fmaf(d0, wd[2 * k2], fmaf(r2, wr[2 * k2], f.x))),
                                           silu_half<SFMT>(fmaf(d0, wd[2 * k2 + 1], fmaf(r2, wr[2 * k2 + 1], f.y))));
                    }
                }
                *reinterpret_cast<uint4*>(xt + ((i << 7) | (l74 ^ (i << 4)))) = make_uint4(o[0], o[1], o[2], o[3]);   // row 8 pw + i, chunk (lane % 8) ^ (row % 8)
                // refill the slot with the same edge of the next tile
                pb[i] = ldg_na_u4(row_ptr(pb_base, (uint32_t)__shfl_sync(0xffffffffu, n_col, i), ldp_b));
                if (TRACE && TRACE_EDGES && i < 2) { asm volatile("" ::: "memory"); if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 10 + 4 * i); }
                cur = nxt; cur_row = next_row;
            }
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 2);
            fence_proxy_async();                                                        // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.bar_full[xs]));
            if (pw == 0 && lane == 0) trace_mark(a.trace, 0, it, 6);
            m_row = n_row; m_col = n_col; m_d0 = n_d0;
            m_r2 = n_moving ? dist2(xr0, xr1, xr2, xc0, xc1, xc2) : n_d0;
            n_row = f_row; n_col = f_col; n_d0 = f_d0;
        }
        // beyond f16's range the clamp above makes an edge differ from the reference (only reachable without a cutoff):
        // flagged once per launch, the caller re-runs in a mode with fp32 edge features (dp_flags.f16_range)
        if (rd_max > 60000.f) atomicOr(a.range_flag, 2);
#endif
    } else {
        // ================================ epilogue ================================
        if (regs_epilogue(MODE) > 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(regs_epilogue(MODE)));
        else if (regs_epilogue(MODE) < 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(regs_epilogue(MODE)));
        const int ew = wid, q = ew & 3, gi = ew >> 2;
        const int c0 = 32 * q + lane, c1 = c0 + 128;                                     // this thread's two channels
        const float hb0 = 0.5f * a.b2[c0], hb1 = 0.5f * a.b2[c1];                        // SiLU(v + b) from hv = v / 2 + b / 2
        const bool gated = COORD || a.attention;
        const float wv0 = gated ? a.wv[c0] : 0.f, wv1 = gated ? a.wv[c1] : 0.f;
        static_assert(H == 256, "segment stores shift by 8");
        float* out0 = a.agg + c0;                                                        // [agg rows | partial rows], see graph.cu edge_dst
        float* out1 = a.agg + c1;
        float* redw = s.red[ew];
        float* gatew = s.gate[ew];
        const int l16 = lane & 15, g16 = lane >> 4;
        if (!a.tma_fill && my_tiles > 0) fill_weights();
        const int unit0 = unit_base(gi), n_units = unit_count(gi);                      // this group's units
        float s0 = 0.f, s1 = 0.f;                                                        // running sum of the current CSR row: carried across tiles
        for (int it = 0; it < my_tiles; ++it) {
            const int ts = it % N_TS;
            const int u0 = (unit0 + it * unit_step) * UNIT_TC;                           // first edge of this group's unit
            const bool has_unit = it < n_units;                                          // shorter lanes idle through the CTA's last tile
            const int my_dst = (!COORD && has_unit && lane < GROUP_EDGES && u0 + lane < E) ? a.edst[u0 + lane] : -1;
            const bool finish = COORD && a.x_next != nullptr;
            if (finish && q == 0) {
                // geometry of the unit's edges, fetched while the tile is still in the pipeline (two dependent L2 trips that
                // nobody waits for) and parked in shared memory so that no register stays live across the epilogue's hot part
                float* st = s.cstash[gi][lane & (GROUP_EDGES - 1)];
                int r = -1 - lane, rs = 0, re = 0;
                float cx = 0.f, cy = 0.f, cz = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (has_unit && lane < GROUP_EDGES && u0 + lane < E) {
                    r = a.erow[u0 + lane];
                    const int c = a.ecol[u0 + lane];
                    x0 = a.x[3 * r]; x1 = a.x[3 * r + 1]; x2 = a.x[3 * r + 2];
                    rs = a.rowptr[r]; re = a.rowptr[r + 1];
                    const float dx = x0 - a.x[3 * c], dy = x1 - a.x[3 * c + 1], dz = x2 - a.x[3 * c + 2];
                    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    const float nrm = __fadd_rn(__fsqrt_rn(__fadd_rn(r2, 1e-8f)), a.norm_constant);      // coord2diff, egnn_new.py:265-270
                    cx = __fdiv_rn(dx, nrm); cy = __fdiv_rn(dy, nrm); cz = __fdiv_rn(dz, nrm);
                }
                if (lane < GROUP_EDGES) {
                    st[0] = cx; st[1] = cy; st[2] = cz; st[3] = x0; st[4] = x1; st[5] = x2;
                    st[6] = __int_as_float(r); st[7] = __int_as_float(rs); st[8] = __int_as_float(re);
                }
                __syncwarp();
            }
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 0);
            mbar_wait_relaxed<EDGE_EPILOGUE_SLEEP_NS>(smem_u32(&s.bar_tfull[ts]), (it / N_TS) & 1);
            tc_fence_after();
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 1);
            float v0[16], v1[16];
            {
                uint32_t r0[16], r1[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + ts * TS_COLS + GROUP_EDGES * gi;
                tmem_ld16_issue(taddr, r0);
                tmem_ld16_issue(taddr + TILE, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) { v0[j] = __uint_as_float(r0[j]); v1[j] = __uint_as_float(r1[j]); }
            }
            tc_fence_before();                                                           // this warp's share of the stage is drained
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.bar_tempty[ts]));
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 2);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v0[j] = silu_half<SFMT>(fmaf(0.5f, v0[j], hb0)); v1[j] = silu_half<SFMT>(fmaf(0.5f, v1[j], hb1)); }
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 3);
            float gate = 1.f;
            if (gated) {
                // sum over the 256 channels of wv[c] * m[c, edge]: the thread's two channels are added in
                // registers, the warp's 32 lanes through a transposed pass over shared memory (thread = channel
                // pair writes a row of 16 edges, thread = (edge, row half) sums a column), the group's 4 warps
                // through s.part
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int j = 4 * j4;
                    *reinterpret_cast<float4*>(redw + lane * RED_STRIDE + j) =
                        make_float4(fmaf(wv1, v1[j], wv0 * v0[j]), fmaf(wv1, v1[j + 1], wv0 * v0[j + 1]),
                                    fmaf(wv1, v1[j + 2], wv0 * v0[j + 2]), fmaf(wv1, v1[j + 3], wv0 * v0[j + 3]));
                }
                __syncwarp();
                float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                                         // bank-conflict-free split of the 32 rows
                    const int k = 8 * kk + 4 * g16;
                    t0 += redw[k * RED_STRIDE + l16];
                    t1 += redw[(k + 1) * RED_STRIDE + l16];
                    t2 += redw[(k + 2) * RED_STRIDE + l16];
                    t3 += redw[(k + 3) * RED_STRIDE + l16];
                }
                float t = (t0 + t1) + (t2 + t3);
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                const int pbuf = it & 1;
                if (lane < GROUP_EDGES) s.part[pbuf][ew][lane] = t;                       // lane = edge inside the unit
                if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 4);
                named_bar_sync(1 + gi, 4 * 32);
                if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 5);
                const float tot = a.bv + ((s.part[pbuf][4 * gi][l16] + s.part[pbuf][4 * gi + 1][l16]) +
                                          (s.part[pbuf][4 * gi + 2][l16] + s.part[pbuf][4 * gi + 3][l16]));
                if (COORD) gate = a.use_tanh ? tanhf(tot) : tot;                         // egnn_new.py:90-93
                else gate = sigmoid_fast(tot);                                           // egnn_new.py:26-29
            }
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 6);
            if (finish) {
                if (q == 0) {
                    // trans = coord_diff * scalar (* range), summed over each CSR row run of the unit in edge order (egnn_new.py:91-103)
                    const float* st = s.cstash[gi][lane & (GROUP_EDGES - 1)];
                    const bool live = lane < GROUP_EDGES;
                    const int row = live ? __float_as_int(st[6]) : -1 - lane;
                    float tx = 0.f, ty = 0.f, tz = 0.f;
                    if (live && row >= 0) {
                        tx = __fmul_rn(st[0], gate); ty = __fmul_rn(st[1], gate); tz = __fmul_rn(st[2], gate);
                        if (a.use_tanh) { tx = __fmul_rn(tx, a.coords_range); ty = __fmul_rn(ty, a.coords_range); tz = __fmul_rn(tz, a.coords_range); }
                    }
                    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                    for (int j = 0; j < GROUP_EDGES; ++j) {
                        const int rj = __shfl_sync(0xffffffffu, row, j);
                        const float vx = __shfl_sync(0xffffffffu, tx, j), vy = __shfl_sync(0xffffffffu, ty, j), vz = __shfl_sync(0xffffffffu, tz, j);
                        if (j <= lane && rj == row) { sx = __fadd_rn(sx, vx); sy = __fadd_rn(sy, vy); sz = __fadd_rn(sz, vz); }
                    }
                    const int row_after = __shfl_sync(0xffffffffu, row, (lane + 1) & 31);
                    if (live && row >= 0 && (lane == GROUP_EDGES - 1 || row_after != row)) {                 // last edge of a row run: its sum is complete
                        const int rs = __float_as_int(st[7]), re = __float_as_int(st[8]);
                        const float d = a.mean ? (float)max(re - rs, 1) : a.norm_factor;
                        float* xo = a.x_next + 3 * (size_t)row;
                        if (rs >= u0 && re <= u0 + GROUP_EDGES) {                                            // the whole row lies in this unit
                            xo[0] = __fadd_rn(st[3], __fdiv_rn(sx, d)); xo[1] = __fadd_rn(st[4], __fdiv_rn(sy, d)); xo[2] = __fadd_rn(st[5], __fdiv_rn(sz, d));
                        } else {
                            const int un = u0 / GROUP_EDGES, fu = rs / GROUP_EDGES, lu = (re - 1) / GROUP_EDGES;
                            float* cp = a.cpart + ((size_t)un * 2 + (un == fu ? 1 : 0)) * 4;
                            __stcg(cp, sx); __stcg(cp + 1, sy); __stcg(cp + 2, sz);
                            __threadfence();
                            if (atomicAdd(a.cticket + row, 1) == lu - fu) {                                  // every other unit of the row has arrived
                                __threadfence();
                                float ax = 0.f, ay = 0.f, az = 0.f;
                                for (int t = fu; t <= lu; ++t) {
                                    const float* pp = a.cpart + ((size_t)t * 2 + (t == fu ? 1 : 0)) * 4;
                                    ax = __fadd_rn(ax, __ldcg(pp)); ay = __fadd_rn(ay, __ldcg(pp + 1)); az = __fadd_rn(az, __ldcg(pp + 2));
                                }
                                xo[0] = __fadd_rn(st[3], __fdiv_rn(ax, d)); xo[1] = __fadd_rn(st[4], __fdiv_rn(ay, d)); xo[2] = __fadd_rn(st[5], __fdiv_rn(az, d));
                                a.cticket[row] = 0;
                            }
                        }
                    }
                    __syncwarp();                                                                            // cstash is rewritten next tile
                }
            } else if (COORD) {
                if (q == 0 && has_unit && lane < GROUP_EDGES && u0 + lane < E) a.escal[u0 + lane] = gate;
            } else {
                // segmented sum over the unit's edges: thread = channel pair, registers = edges (two FMA chains)
                if (gated) {
                    if (lane < GROUP_EDGES) gatew[lane] = gate;
                    __syncwarp();
                }
                const unsigned last_mask = __ballot_sync(0xffffffffu, my_dst >= 0);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float gq[4] = {1.f, 1.f, 1.f, 1.f};
                    if (gated) {
                        const float4 t = *reinterpret_cast<const float4*>(gatew + 4 * j4);
                        gq[0] = t.x; gq[1] = t.y; gq[2] = t.z; gq[3] = t.w;
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * j4 + jj;
                        s0 = fmaf(gq[jj], v0[j], s0);
                        s1 = fmaf(gq[jj], v1[j], s1);
                        if (last_mask & (1u << j)) {                                     // warp-uniform
                            const size_t d = (size_t)(unsigned)__shfl_sync(0xffffffffu, my_dst, j) << 8;   // row d of [agg | partials], H = 256
                            out0[d] = s0; out1[d] = s1;
                            s0 = 0.f; s1 = 0.f;
                        }
                    }
                }
                __syncwarp();                                                            // gatew / redw are rewritten next tile
            }
            if (q == 0 && lane == 0) trace_mark(a.trace, 2 + gi, it, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == MMA_WARP) tmem_dealloc(tmem_w, TMEM_COLS);
    if (tid == MMA_WARP * 32) trace_mark(a.trace, 1, 62, 1);                              // kernel exit
    if (TRACE && tid == MMA_WARP * 32 && blockIdx.x < 376) {                              // ... and at exit
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[(64 + 16 + (blockIdx.x >> 3)) * 16 + 2 * (blockIdx.x & 7) + 1] = (long long)gt;
    }
}

}  // namespace

template <int MODE>
static int edge_set_smem(int smem)
{
    DP_CUDA(cudaFuncSetAttribute(edge_tc_kernel<MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DP_CUDA(cudaFuncSetAttribute(edge_tc_kernel<MODE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DP_CUDA(cudaFuncSetAttribute(edge_tc_kernel<MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return DP_OK;
}

int tc_edge_init()
{
    static_assert(sizeof(EdgeSmem) + 1024 <= 232448, "edge kernel shared memory exceeds 227 KB");
    const int smem = (int)sizeof(EdgeSmem) + 1024;
    int rc = edge_set_smem<tc::MODE_F16>(smem);
    if (!rc) rc = edge_set_smem<tc::MODE_BF16>(smem);
    if (!rc) rc = edge_set_smem<tc::MODE_F16P>(smem);
    if (!rc) rc = edge_set_smem<tc::MODE_F16Q>(smem);
    return rc;
}

template <int MODE>
static cudaError_t edge_launch(dp_handle* h, const EdgeArgs& a, const unsigned char* img, int grid, int smem, cudaStream_t st)
{
    if (a.coord) return launch_kernel(h->pdl, edge_tc_kernel<MODE, false, true>, dim3(grid), dim3(THREADS), smem, st, a, img);
    if (a.trace) return launch_kernel(h->pdl, edge_tc_kernel<MODE, true, false>, dim3(grid), dim3(THREADS), smem, st, a, img);
    return launch_kernel(h->pdl, edge_tc_kernel<MODE, false, false>, dim3(grid), dim3(THREADS), smem, st, a, img);
}

int launch_edge_tc(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st)
{
    int fmt = 0, rc = tc_fmt_of(h, &fmt);
    if (rc) return rc;
    DP_CHECK(h->tc && lin_id >= 0 && lin_id < (int)h->tc->lin.size() && h->tc->lin[lin_id].img[fmt], DP_ERR_STATE,
             "tc edge layer %d has no weight image", lin_id);
    const TcLinearImg& L = h->tc->lin[lin_id];
    DP_CHECK(L.K == H && L.n_out == H, DP_ERR_INVALID, "tc edge layer %d: shape mismatch", lin_id);
    DP_CHECK((unsigned long long)h->plan.N * (unsigned long long)a.ldp < (1ull << 32), DP_ERR_INVALID,
             "batch of %d nodes exceeds the 32-bit element indexing of the projected features", h->plan.N);
    const int smem = (int)sizeof(EdgeSmem) + 1024;
    const int grid = h->sm_count;
    const unsigned char* img = L.img[fmt];
    switch (h->precision) {
    case DP_BF16: DP_CUDA(edge_launch<tc::MODE_BF16>(h, a, img, grid, smem, st)); break;
    case DP_F16: DP_CUDA(edge_launch<tc::MODE_F16>(h, a, img, grid, smem, st)); break;
    case DP_F16_FAST: DP_CUDA(edge_launch<tc::MODE_F16P>(h, a, img, grid, smem, st)); break;
    case DP_F16_FAST32: DP_CUDA(edge_launch<tc::MODE_F16Q>(h, a, img, grid, smem, st)); break;
    default: DP_CHECK(false, DP_ERR_STATE, "precision %d has no tcgen05 edge kernel", h->precision);
    }
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
