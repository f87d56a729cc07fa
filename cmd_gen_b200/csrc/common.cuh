// Internal declarations shared by the translation units of libdiffphar_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/diffphar_b200.h"

constexpr int H = 256;            // hidden_nf (compile-time tile width)
constexpr int UNIT_F32 = 64;      // edges per segmented-sum unit, FFMA path
constexpr int UNIT_TC = 16;       // edges per segmented-sum unit, tcgen05 path (one epilogue group's share of a 64-edge tile)
constexpr int CELLS_DIM_MAX = 16;                 // cells per axis of one sample's grid
constexpr int CELLS_MAX = CELLS_DIM_MAX * CELLS_DIM_MAX * CELLS_DIM_MAX;
constexpr int CELL_SAMPLE_MAX_NODES = 8192;       // per-row bitmap of the cell-list builder: 256 words per warp
constexpr int DP_TRACE_WORDS = 6 * 64 * 16;   // debug timeline: [role][tile iteration][slot]

void dp_set_error(const char* fmt, ...);

#define DP_CUDA(expr)                                                                 \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            dp_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DP_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

#define DP_CHECK(cond, code, ...)                                                     \
    do {                                                                              \
        if (!(cond)) {                                                                \
            dp_set_error(__VA_ARGS__);                                                \
            return (code);                                                            \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with the attribute may begin while its predecessor
// in the stream drains; everything before pdl_wait() must touch only launch-invariant data (weights, the
// plan's layout arrays), everything after sees the predecessor's memory.  Without the attribute both
// instructions are no-ops.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// the same for a kernel that runs as thread-block clusters of `cluster` CTAs along x (grid.x a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(bool pdl, int cluster, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------
// device-side math shared by kernels
// ---------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// ---------------------------------------------------------------------------
// Segmented sum of the tcgen05 edge kernel: two work splits, one bookkeeping.  An epilogue group (4 warps) reduces one
// 16-edge unit (UNIT_TC) per tile and keeps the running sum of the current CSR row in registers.
//   * lanes (Plan::seg_lanes, full-atom pockets, degree ~ 40): the U = ceil(E / 16) units are split into L contiguous,
//     balanced ranges ("lanes": one per epilogue group of one CTA, L = 4 x CTAs); lane l owns units
//     [l U / L, (l + 1) U / L) and walks them in order, so the sum carries from tile to tile and a row is stored
//     ONCE: whole to agg[row] when its edges lie in one lane (all but <= L - 1 rows), else one partial row per lane.
//   * units (Calpha pockets, degree ~ 7): tile t = 64 consecutive edges goes to CTA t mod CTAs (consecutive edges
//     share their Pa rows in L1: measured 17 % faster than lanes at this degree); a "lane" is then a single unit,
//     l = u, and a row crossing 16-edge boundaries is stored as one partial row per unit.
// Partial row 2 l + slot: slot 0 = the row that contains the lane's first edge, slot 1 = the row that starts inside
// the lane and runs past its end.
// ---------------------------------------------------------------------------
// 64-bit products: U L reaches 2^32 for a config-3 batch of 512 samples (41 M edges, 592 lanes); the divisions run a
// few times per CTA / per row, never per edge (edge_dst only asks for rows that cross a lane boundary)
__host__ __device__ __forceinline__ unsigned lane_first_unit(unsigned l, unsigned U, unsigned L) { return (unsigned)((unsigned long long)l * U / L); }
__host__ __device__ __forceinline__ unsigned lane_of_unit(unsigned u, unsigned U, unsigned L) { return (unsigned)((((unsigned long long)u + 1ull) * L - 1ull) / U); }
// agg_src[row] (written by the graph builder's fill pass, read by every consumer of the aggregate):
//   >= 0      : the whole sum is agg[row]
//   AGG_EMPTY : the row has no edges
//   otherwise : -(1 + (((first lane << 10) | extra lanes) << 1 | slot of the first piece)): pieces are
//               partial[2 lf + slot], then partial[2 l] for l = lf + 1 .. lf + extra, added in that order
constexpr int AGG_EMPTY = -2147483647 - 1;
constexpr int MAX_LANES = 1024;              // lanes scheme; the units scheme allows 2^20 units (first-lane field of agg_src)
__host__ __device__ __forceinline__ int agg_src_split(unsigned lf, unsigned extra, unsigned slot) { return -(int)(1u + ((((lf << 10) | extra) << 1) | slot)); }

// Per-row aggregate assembled from the edge kernel's outputs.  A row whose edge range
// [s,e) lies inside one segmented-sum unit was stored to agg[]; a row that crosses unit
// boundaries was stored as per-unit partial sums (slot 0: segment containing the unit's
// first edge, slot 1: the other boundary segment).  Summed here in unit order — no atomics.
struct AggView {
    const float* agg;        // [N][H] raw sums of complete rows
    const float* partials;   // [units][2][H] (FFMA path) or [lanes][2][H] (tcgen05 path)
    const int* src;          // tcgen05 path: agg_src[N] (see above); null: the FFMA path's per-tile scheme below
    const int* rowptr;       // [N+1]
    int unit;                // FFMA path: edges per unit (= tile)
    float inv_norm;          // 1/normalization_factor ('sum'); unused for 'mean'
    float norm;              // normalization_factor
    int mean;                // aggregation_method == 'mean'
};

__device__ __forceinline__ float4 agg_load4(const AggView& a, int row, int c)
{
    const int s = a.rowptr[row], e = a.rowptr[row + 1];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e > s) {
        const int uf = a.src ? 0 : s / a.unit, ul = a.src ? 0 : (e - 1) / a.unit;
        if (a.src) {
            const int code = a.src[row];
            if (code >= 0) {
                v = *reinterpret_cast<const float4*>(a.agg + (size_t)code * H + c);
            } else if (code != AGG_EMPTY) {
                const unsigned k = (unsigned)(-(code + 1));
                const unsigned lf = k >> 11, extra = (k >> 1) & 1023u, slot = k & 1u;
                for (unsigned i = 0; i <= extra; ++i) {
                    const float4 p = *reinterpret_cast<const float4*>(a.partials + ((size_t)(lf + i) * 2 + (i == 0 ? slot : 0u)) * H + c);
                    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
                }
            }
        } else if (uf == ul) {
            v = *reinterpret_cast<const float4*>(a.agg + (size_t)row * H + c);
        } else {
            for (int u = uf; u <= ul; ++u) {
                const int slot = (s <= u * a.unit) ? 0 : 1;
                const float4 p = *reinterpret_cast<const float4*>(a.partials + ((size_t)u * 2 + slot) * H + c);
                v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
            }
        }
        // reference: result / normalization_factor (egnn_new.py:283-285) or / count (:287-291)
        const float d = a.mean ? (float)(e - s) : a.norm;
        v.x = __fdiv_rn(v.x, d); v.y = __fdiv_rn(v.y, d); v.z = __fdiv_rn(v.z, d); v.w = __fdiv_rn(v.w, d);
    }
    return v;
}

// ---------------------------------------------------------------------------
// weights as laid out on the device
// ---------------------------------------------------------------------------
struct DevLinear {          // y = x W^T + b, stored transposed: wt[k * out + o]
    float* wt = nullptr;
    float* b = nullptr;     // may be null
    int in = 0, out = 0;
};

struct GclWeights {
    // edge_mlp.0 split: columns [0,H) act on h[row], [H,2H) on h[col], 2H on r2, 2H+1 on d0
    float* wr = nullptr;    // [H]
    float* wd = nullptr;    // [H]
    DevLinear e2;           // edge_mlp.2
    float* wa = nullptr;    // att_mlp.0 weight [H]
    float ba = 0.f;
    DevLinear n0, n2;       // node_mlp
};

struct CoordWeights {
    float* wr = nullptr;
    float* wd = nullptr;
    DevLinear c2;
    float* w4 = nullptr;    // coord_mlp.4 weight [H]
};

// A projection set: one GEMM h[N,H] -> PQ[N, n_out] feeding the next edge kernels.
struct ProjSet {
    DevLinear lin;          // in = H, out = 512 or 1024
    float* b_half = nullptr; // 0.5 * bias: the tcgen05 path keeps pq pre-scaled by 1/2 (SiLU(x) = hv + hv tanh(hv), hv = x / 2)
    int off_gcl = -1;       // column offset of (Pa|Pb) for the next GCL, -1 if none
    int off_coord = -1;     // column offset of (Qa|Qb) for the coordinate update, -1 if none
};

struct DeviceWeights {
    DevLinear phar_enc0, phar_enc2, phar_dec0, phar_dec2;
    DevLinear res_enc0, res_enc2, res_dec0, res_dec2;
    DevLinear emb, emb_out;
    std::vector<GclWeights> gcl;       // n_layers * inv_sublayers
    std::vector<CoordWeights> coord;   // n_layers
    std::vector<ProjSet> proj;         // 1 + number of GCLs
    std::vector<void*> allocations;
};

struct TcWeights;   // tcgen05 operand images (tc_weights.cu)
struct HostLinear {  // k-major fp32 host copy of a linear the tensor-core path packs: wt[k * n_out + o]
    std::vector<float> wt;
    int K = 0, n_out = 0;
    float tf32_scale = 1.f;   // the projections are registered pre-scaled by 1/2 for the 16-bit path; the tf32 image undoes it
};

// ---------------------------------------------------------------------------
// the plan: batch layout + workspace
// ---------------------------------------------------------------------------
struct Plan {
    int B = 0, Np = 0, Nr = 0, N = 0;
    int64_t Ecap = 0;
    int max_phar = 0;
    int max_nodes = 0;           // largest sample (phar + pocket nodes)
    // layout
    int* phar_off = nullptr;     // [B+1]
    int* res_off = nullptr;      // [B+1]
    int* sample_of = nullptr;    // [N]
    // graph
    int* deg = nullptr;          // [N]
    int* rowptr = nullptr;       // [N+1]
    int* col = nullptr;          // [Ecap]
    int* erow = nullptr;         // [Ecap]
    int* edst = nullptr;         // [Ecap] tcgen05 path: where the running segmented sum is stored after this edge (-1: keep going)
    int* agg_src = nullptr;      // [N] tcgen05 path: where a consumer finds the row's aggregate (AggView::src)
    int n_lanes = 0;             // segmented-sum lanes of the tcgen05 edge kernel = 4 x its CTAs; 0 = per-unit scheme (see seg_lanes)
    int seg_lanes = 0;           // 1: contiguous lane ranges (high-degree graphs: full-atom pockets); 0: round-robin tiles, per-unit
                                 // partial rows (Calpha pockets: consecutive edges of a tile share their Pa rows in L1)
    float* d0 = nullptr;         // [Ecap] squared input-frame distances (edge_attr, egnn_new.py:195)
    int* counts = nullptr;       // [4]: E, E_p, overflow, spare
    // cell list of the bucketed radius-graph builder (graph.cu), rebuilt by every denoiser evaluation
    int use_cells = 0;
    int* cell_start = nullptr;   // [B][CELLS_MAX + 1] offsets into cell_nodes (absolute)
    int* cell_nodes = nullptr;   // [N] node ids bucketed by cell, samples back to back
    float* cell_grid = nullptr;  // [B][8]: origin xyz, inverse cell size xyz, dims packed (nx | ny << 8 | nz << 16) as int bits
    unsigned* row_bitmap = nullptr;  // [N][bitmap_words]: per-row hit bitmap over the sample's nodes, count pass -> fill pass
    int bitmap_words = 0;        // ceil(max nodes per sample / 32)
    int fused_graph = 0;         // one-launch scan builder (graph.cu radius_rows_fused_kernel): small samples, units scheme
    unsigned long long* scan_status = nullptr;   // [ceil(N / 8)] look-back status words of that kernel
    int* scan_ticket = nullptr;  // [4] arrival counter of the count pass: its last CTA runs the rowptr scan
    // node state
    float* h = nullptr;          // [N][H]
    float* h_base = nullptr;     // [Nr][H] sampler only: embedding of the (static) pocket features without the time term
    float* tbuf = nullptr;       // [N][H] node-MLP hidden
    float* agg = nullptr;        // [N][H]
    float* partials = nullptr;   // [max(FFMA units, lanes)][2][H]
    float* pq = nullptr;         // [N][1024] fp32 (FFMA mode) or [N][1024] f16 pre-scaled by 1/2 (tcgen05 modes)
    float* cpart = nullptr; int* cticket = nullptr;   // in-kernel coordinate finish of the tcgen05 edge kernel (EdgeArgs::x_next)
    unsigned char* h16 = nullptr; // 16-bit copy of h as the node kernel's own swizzled B-tile images, one per node tile (tc_node.cu):
                                 // written by the launch that produces an h version, bulk-loaded by the launch that consumes it
    float* x_in = nullptr;       // [N][3]
    float* x_a = nullptr;        // [N][3]
    float* x_b = nullptr;        // [N][3]
    float* escal = nullptr;      // [Ecap] coordinate-MLP scalar per edge
    // sampler state
    float* z = nullptr;          // [Np][3+P]
    float* eps_hat = nullptr;    // [Np][3+P]
    float* pocket = nullptr;     // [Nr][3+R]
    float* t_const = nullptr;    // [1]
    int* step_idx = nullptr;     // [1]
    float* step_rows = nullptr;  // [n_steps][4]   (handle-owned: dp_handle::steps)
    float* stats = nullptr;      // [n_steps+2][2] (max |sum x|, max |x|) as float bits (handle-owned)
    int stats_cap = 0;
    int* nan_flag = nullptr;     // [4]: current-call NaN flag, sticky NaN count, sticky f16-range bits (1: |pq| beyond what f16 adds
                                 // safely, 2: r2 / d0 clamped to 60 000 by the packed-f16 first layer), spare
    int64_t* sample_ids = nullptr;   // [B] global id of each sample: selects its counter-based noise stream (small.cu)
    float* out_buf = nullptr;    // [Np][3+P] dp_sample_host staging
    // layout as planned (host copies: an identical dp_plan call keeps the captured graph)
    std::vector<int> phar_counts_host, res_counts_host;
    int64_t ecap_request = 0;
};

// A device buffer that only ever grows: re-planning a smaller or equal batch costs no cudaMalloc / cudaFree
// (each is a device-wide synchronisation) — what a pocket-list workload (BASELINE config 4) does per pocket.
struct GrowBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes, bool* moved = nullptr);    // api.cu; 25 % headroom when it has to grow
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct dp_handle {
    dp_config cfg;
    int device = 0;
    int sm_count = 148;
    int precision = 0;
    bool joint = false;                // dp_set_update_pocket_coords: every node's coordinates move (update_pocket_coords=True, dynamics.py:104-107, 133-136)
    int dbg = 0;                       // DIFFPHAR_DBG: timing-experiment bits (results may be wrong), 0 in production
    int seg_mode = 0;                  // DIFFPHAR_SEG: 0 automatic, 1 units, 2 lanes (Plan::seg_lanes)
    int node_pair = 0;                 // DIFFPHAR_NODE_PAIR: node kernel as CTA pairs (cluster of 2, tcgen05 cta_group::2)
    int node_mc = 0;                   // DIFFPHAR_NODE_MC: node kernel in clusters of 2 whose weight panels arrive by TMA multicast (each CTA
                                       // fetches half of every panel for both).  Measured 3 % SLOWER (profiles/r06a_ab_summary.txt): the GEMM
                                       // phases are paced by the MMAs' own operand fetch, not by the L2 -> SM weight stream
    int coord_fused = 1;               // DIFFPHAR_COORD_FUSED: the tcgen05 coordinate-mode edge kernel finishes its rows itself (no coord_finish launch)
    int node_h16 = 1;                  // DIFFPHAR_NODE_H16: h travels between the node launches as 16-bit tile images (TMA in / out)
    int node_split = 64;               // DIFFPHAR_NODE_SPLIT: nodes per tile for the tiles that hold phar rows (one projection block more
                                       // than the rest); 0 = uniform tiles
    int trace_cta = 0;                 // DIFFPHAR_TRACE_CTA: which CTA of the traced kernel writes the timeline
    int trace_v = -1;                  // DIFFPHAR_TRACE_V: only the node launch of h version v writes it (-1: every launch, the last one stays)
    int tma_fill = 1;                  // DIFFPHAR_TMA_FILL=0: resident weights through LDG + tcgen05.st (A/B; EdgeArgs::tma_fill)
    bool pdl = false;                  // programmatic dependent launch between the kernels of a step (DIFFPHAR_PDL=1 enables; measured neutral inside graph replay)
    int skip_mask = 0;                 // DIFFPHAR_SKIP (timing experiments only, results are garbage): 1 edge msg, 2 node, 4 coord edge, 8 coord finish, 16 graph, 32 encode/decode, 64 ddpm
    int graph_mode = 0;                // DIFFPHAR_GRAPH: 0 = auto (cell list for samples of >= 512 nodes), 1 = always scan, 2 = always cells,
                                       // 4 = scan as ONE launch where it applies (count + look-back scan + fill; measured slower)
    int tc_mask = 3;                   // debug: bit 0 = edge kernels on tcgen05, bit 1 = node linears (DIFFPHAR_TC_MASK)
    bool has_weights = false;
    DeviceWeights w;
    TcWeights* tc = nullptr;
    std::vector<HostLinear> tc_host;   // indexed by lin_id (see run_denoiser)
    Plan plan;
    bool has_plan = false;
    GrowBuf arena;                     // every per-plan buffer is carved out of this one allocation (dp_plan)
    GrowBuf steps;                     // step table rows + guard statistics (dp_set_step_table)
    GrowBuf noise;                     // [n_steps+2][Np][3+P]: the noise the captured step graph reads (injected noise is copied
                                       // in, seeded noise is generated in place), so the graph never bakes a caller pointer
    GrowBuf frames;                    // return_frames > 1: [frames][Np][3+P] then [frames][Nr][3+R], un-normalised
    // the captured denoising step (one per handle; re-captured when something baked into it changes)
    cudaGraphExec_t step_graph = nullptr;
    int graph_precision = -1;
    int graph_frames = 0;              // return_frames the graph was captured for (0: no frame output)
    int64_t graph_launches = 0;        // kernels per replay of step_graph
    int64_t graph_captures = 0;        // how many times the step graph was (re)captured (tests: re-planning the same
                                       // layout or sampling with fresh caller tensors must not re-capture)
    std::vector<float> step_rows_host;
    float final_host[4] = {0, 0, 0, 0};
    int n_steps = 0;
    int64_t launches = 0;
    cudaStream_t capture_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // second branch of a denoiser evaluation: the radius graph runs beside the first projection
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int trace_kernel = 0;              // DIFFPHAR_TRACE: 1 = node kernel, 2 = edge message kernel
    long long* trace = nullptr;        // debug: per-role clock64 timeline of CTA 0 of the tcgen05 kernels (DIFFPHAR_TRACE=1)
    // profiling
    int profile = 0;                   // dp_profile_enable: 1 = eager launches, CUDA events around every launch; 2 = the events are recorded
                                       // INSIDE the captured step graph (external event-record nodes) and read after every replay:
                                       // launch durations in the production context, without the idle front end of an eager launch
    // 3 = as 1, with every message-kernel span holding DP_PROFILE_REPEAT back-to-back launches of the same kernel
    bool capturing = false;            // between cudaStreamBeginCapture / EndCapture of the step graph
    struct Span { int which; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double prof_ms[8] = {0};
    int64_t prof_n[8] = {0};
};

constexpr int DP_PROFILE_REPEAT = 8;
enum ProfWhich { PROF_EDGE_MSG = 0, PROF_NODE = 1, PROF_EDGE_COORD = 2, PROF_GRAPH = 3, PROF_DDPM = 4, PROF_OTHER = 5 };

// ---------------------------------------------------------------------------
// kernel launchers (one per translation unit)
// ---------------------------------------------------------------------------
// graph.cu
int launch_build_edges(dp_handle* h, const float* x_dev, cudaStream_t st);

// egnn_f32.cu
struct LinearArgs {
    const float* x; int ldx;          // primary input rows
    AggView aggv;                     // used when two_source
    int two_source;                   // input = [x(0:H) | agg(0:H)], K = 2H
    int n_rows; int K;
    const float* wt; const float* bias; int n_out;
    float* y; int ldy;
    const float* resid; int ldr;      // epi 2: y = resid + acc + bias
    int epi;                          // 0: +bias, 1: silu(+bias), 2: residual
};
int launch_linear_f32(dp_handle* h, const LinearArgs& a, cudaStream_t st);

struct EdgeArgs {
    const float* p; int ldp; int off_a; int off_b;   // pre-projected node features
    const float* wr; const float* wd;                // [H] weights of the two edge scalars
    const float* w2t; const float* b2;               // second layer, k-major [H][H]
    const float* wv; float bv;                       // final vector: attention / coord_mlp.4
    const float* x;                                  // current coordinates [N][3]
    const float* d0; const int* erow; const int* ecol; const int* rowptr;
    const int* edst;                                 // per-edge segmented-sum destination (graph.cu), tcgen05 path
    int n_moving;                                    // rows [0, n_moving) changed coordinates since the graph build
    const int* n_edges;                              // device scalar: edges to process
    int contig;                                      // tcgen05 path: 1 = contiguous lane ranges, 0 = round-robin 64-edge tiles (Plan::seg_lanes)
    int dbg;                                         // timing experiments (DIFFPHAR_DBG bits; results are wrong), 0 in production
    int tma_fill;                                    // tcgen05 path: resident weights through TMA + tcgen05.cp instead of LDG + tcgen05.st
    int ecap;                                        // allocated length of the per-edge arrays (speculative first-tile loads)
    float* agg; float* partials;                     // message outputs (coord == 0)
    float* escal;                                    // per-edge scalar output (coord == 1)
    // coord == 1, tcgen05 path: the coordinate update is finished inside the kernel (x_next != null) — no second launch.
    // Per 16-edge unit the epilogue sums coord_diff * scalar over each CSR row run; a row that lies inside one unit is
    // written at once, a row that spans several units leaves one partial per unit in cpart[2 * unit + slot] (slot 1: the row
    // runs on into the next unit, slot 0: it came from the previous one) and bumps cticket[row]: the unit that arrives
    // last adds the pieces in unit order (deterministic whatever the arrival order), writes the row and clears the ticket.
    float* x_next; float* cpart; int* cticket;
    float norm_constant, coords_range, norm_factor; int mean;
    int coord; int attention; int use_tanh;
    long long* trace;                                // debug timeline (dp_debug_trace), normally null
    int* range_flag;                                 // tcgen05 path: sticky f16-range bits (Plan::nan_flag + 2)
};
int launch_edge_f32(dp_handle* h, const EdgeArgs& a, cudaStream_t st);
int egnn_f32_init();

// small.cu
// base_mode 0: full encoders for every node; 1: write plan.h_base for the pocket nodes only (no time term);
// 2: pocket nodes = h_base + t * w_time (their type features are constant during sampling), phar nodes in full
int launch_encode_nodes(dp_handle* h, const float* xh_phar, const float* xh_res, const float* t_base,
                        const int* step_idx, int row_stride, int t_stride, int base_mode, cudaStream_t st);
// node tiles of the fused tcgen05 node kernel (tc_node.cu): the first `tp` tiles hold `sp` nodes each, the rest `stride`
struct NodeTiling { int tp, sp, stride, grid; };
void node_tiling(const dp_handle* h, int N, int Np, NodeTiling* t);
size_t node_tile_image_bytes();
int launch_coord_finish(dp_handle* h, const float* x_cur, float* x_next, int n_moving, cudaStream_t st);
int launch_velocity_center(dp_handle* h, float* out_phar, float* out_res, cudaStream_t st);
int launch_decode(dp_handle* h, const float* x_final, float* out_phar, float* out_res, cudaStream_t st);
int launch_nan_fixup(dp_handle* h, float* out_phar, float* out_res, cudaStream_t st);
struct DdpmArgs {
    int kind; float a, c, sigma;              // immediate constants (table == null)
    const float* table; const int* step_idx;  // or: row = table[*step_idx] = (t, a, c, sigma), kind 0
    float* z; float* pocket; const float* eps_hat; const float* noise;
    int64_t noise_step_stride;                // noise + (*step_idx + noise_step_base) * stride when table != null
    int noise_step_base;
    int stat_index;                           // stats row (table == null), else *step_idx + stat_base
    int stat_base;
    int advance;                              // 1: increment *step_idx afterwards (separate tiny kernel)
    // return_frames > 1 (conditional_model.py:439-442): the step with s = n_steps - 1 - *step_idx writes the
    // un-normalised state (en_diffusion.py:891-906) to frame s * return_frames / n_steps when that division is exact
    float* frames_phar; float* frames_pocket; int return_frames; int n_steps;
    float norm_x, norm_h, bias_h;
    // sampler: the update also writes the coordinates of the NEXT denoiser evaluation ([N][3] each: input frame + the two
    // ping-pong copies), so that its radius graph can start before the encoder (null: the encoder writes them)
    float* x_in; float* x_a; float* x_b;
};
int launch_ddpm(dp_handle* h, const DdpmArgs& a, cudaStream_t st);
int launch_fill_noise(dp_handle* h, uint64_t seed, int n_draws, float* noise_dev, cudaStream_t st);
int launch_pocket_com_init(dp_handle* h, float* z, const float* pocket, cudaStream_t st);

// tc_weights.cu (tcgen05)
int tc_init();
int tc_prepare_weights(dp_handle* h);
void tc_free_weights(dp_handle* h);
int launch_edge_tc(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st);   // tc_edge.cu
int tc_edge_init();
int launch_node_tc(dp_handle* h, int v, const AggView& av, cudaStream_t st);                 // tc_node.cu
int tc_node_init();
int launch_linear_tf32(dp_handle* h, const LinearArgs& a, int lin_id, cudaStream_t st);     // tc_tf32.cu (DP_TF32)
int launch_edge_tf32(dp_handle* h, const EdgeArgs& a, int lin_id, cudaStream_t st);
int tc_tf32_init();

// api.cu helpers
void prof_begin(dp_handle* h, int which, cudaStream_t st);
void prof_end(dp_handle* h, cudaStream_t st);
