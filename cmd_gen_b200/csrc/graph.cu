// K1 — radius graph as CSR.  Replaces EGNNDynamics.get_edges
// (DiffPhar/equivariant_diffusion/dynamics.py:141-147): the reference materialises a dense
// N x N adjacency (cdist + mask compare + nonzero).  Two builders, identical output:
//   * scan  (small samples, C-alpha pockets): one warp per row sweeps the nodes of the row's own sample
//     (the only possible neighbours) in ascending node index; ballot compaction emits sorted columns.
//   * cells (samples of >= 512 nodes, full-atom pockets): a bucketed cell list per sample (cells >= cutoff,
//     <= 16 per axis, rebuilt per call because the phar nodes move and the pocket is re-centred), one warp
//     per row tests the 27 neighbouring buckets and records hits in a per-warp BITMAP over the sample's
//     nodes; reading the bitmap in word order emits the columns sorted, whatever order the buckets hold.
// Either way the output is the (row, col) lexicographic order torch.where yields, bit-exact by construction.
//
// Predicate (contract, SURVEY.md §7 hard part 1):
//   fp32  sqrt((dx*dx + dy*dy) + dz*dz) <= cutoff, every op rounded separately (no FMA).
// The squared distance doubles as the edge attribute d0 of EGNN.forward (egnn_new.py:195).
#include "common.cuh"

namespace {

struct GraphArgs {
    const float* x;        // [N][3]
    const int* sample_of;  // [N]
    const int* phar_off;   // [B+1]
    const int* res_off;    // [B+1]
    int N, Np;
    float cutoff;          // < 0: none
    int* deg;
    const int* rowptr;
    int* col; int* erow; float* d0; int* edst; int* agg_src; int n_lanes;
    int* counts;           // [0]=E [1]=E_p [2]=overflow
    long long ecap;
    int* cell_start; int* cell_nodes; float* cell_grid;
    unsigned* row_bitmap; int bitmap_words;          // cells builder: the count pass keeps each row's hit bitmap for the fill pass
    unsigned long long* status;                      // fused builder: per-CTA scan status words (decoupled look-back), zeroed per build
    int* ticket;                                     // count pass: arrival counter of the last-block scan (null: stand-alone scan kernel)
};

__device__ void scan_rowptr_block(const int* __restrict__ deg, int* __restrict__ rowptr, int N, int Np, int* counts, long long ecap);
__device__ bool last_block_done(int* ticket);

__device__ __forceinline__ float dist2_exact(float xi, float yi, float zi, float xj, float yj, float zj)
{
    const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Segmented-sum bookkeeping of one CSR row [rs, re) for the tcgen05 edge kernel (lanes: common.cuh).  Computed once
// per row by the fill pass: the lanes of its first and last edge.  A row inside ONE lane (all but <= L - 1 rows)
// is stored whole after its last edge; a row that crosses lane boundaries is stored as one partial row per lane.
struct RowSeg { unsigned lf, ll, U, L; int N; };   // L == 0: units scheme (lane = unit)
__device__ __forceinline__ unsigned seg_lane_of(const RowSeg& r, unsigned u) { return r.L ? lane_of_unit(u, r.U, r.L) : u; }
__device__ __forceinline__ unsigned seg_lane_first(const RowSeg& r, unsigned l) { return r.L ? lane_first_unit(l, r.U, r.L) : l; }
__device__ __forceinline__ RowSeg row_seg(int n_nodes, long long rs, long long re, int E, int L)
{
    RowSeg r;
    r.U = (unsigned)((E + UNIT_TC - 1) / UNIT_TC); r.L = (unsigned)L; r.N = n_nodes;
    r.lf = r.ll = 0;
    if (re > rs) { r.lf = seg_lane_of(r, (unsigned)(rs / UNIT_TC)); r.ll = seg_lane_of(r, (unsigned)((re - 1) / UNIT_TC)); }
    return r;
}
// destination of the running sum after edge `pos` of the row: -1 = keep accumulating, else a row of the
// contiguous [agg (N rows) | partials (2 per lane)] buffer
__device__ __forceinline__ int edge_dst(const RowSeg& r, int row, long long rs, long long re, long long pos)
{
    if (r.lf == r.ll) return pos == re - 1 ? row : -1;
    const unsigned l = seg_lane_of(r, (unsigned)(pos / UNIT_TC));
    const long long lane_start = (long long)seg_lane_first(r, l) * UNIT_TC, lane_end = (long long)seg_lane_first(r, l + 1) * UNIT_TC;
    if (pos != re - 1 && pos != lane_end - 1) return -1;
    return r.N + (int)(2u * l + (rs <= lane_start ? 0u : 1u));
}
__device__ __forceinline__ int row_agg_src(const RowSeg& r, int row, long long rs, long long re)
{
    if (re <= rs) return AGG_EMPTY;
    if (r.lf == r.ll) return row;
    const long long first_start = (long long)seg_lane_first(r, r.lf) * UNIT_TC;
    return agg_src_split(r.lf, r.ll - r.lf, rs <= first_start ? 0u : 1u);
}

// FILL = false: count neighbours into deg[]; FILL = true: write col/erow/d0 at rowptr[row].
template <bool FILL>
__global__ void __launch_bounds__(256) radius_rows_kernel(GraphArgs a)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < a.N; row += gridDim.x * warps_per_block) {
        const int b = a.sample_of[row];
        const float xi = a.x[3 * row], yi = a.x[3 * row + 1], zi = a.x[3 * row + 2];
        // candidate ranges in ascending node index: the sample's phar nodes, then its pocket nodes
        const int lo[2] = {a.phar_off[b], a.Np + a.res_off[b]};
        const int hi[2] = {a.phar_off[b + 1], a.Np + a.res_off[b + 1]};
        int found = 0;
        long long base = FILL ? (long long)a.rowptr[row] : 0;
        const long long row_end = FILL ? (long long)a.rowptr[row + 1] : 0;
        RowSeg seg{};
        if (FILL) {
            seg = row_seg(a.N, base, row_end, a.rowptr[a.N], a.n_lanes);
            if (lane == 0) a.agg_src[row] = row_agg_src(seg, row, base, row_end);
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            for (int j0 = lo[part]; j0 < hi[part]; j0 += 32) {
                const int j = j0 + lane;
                bool hit = false;
                float d2 = 0.f;
                if (j < hi[part]) {
                    d2 = dist2_exact(xi, yi, zi, a.x[3 * j], a.x[3 * j + 1], a.x[3 * j + 2]);
                    hit = (a.cutoff < 0.f) || (__fsqrt_rn(d2) <= a.cutoff);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (FILL && hit) {
                    const long long pos = base + found + __popc(m & ((1u << lane) - 1u));
                    if (pos < a.ecap) {
                        a.col[pos] = j;
                        a.erow[pos] = row;
                        a.d0[pos] = d2;
                        a.edst[pos] = edge_dst(seg, row, base, row_end, pos);
                    }
                }
                found += __popc(m);
            }
        }
        if (!FILL && lane == 0) a.deg[row] = found;
    }
    if (!FILL && a.ticket && last_block_done(a.ticket))
        scan_rowptr_block(a.deg, const_cast<int*>(a.rowptr), a.N, a.Np, a.counts, a.ecap);
}

// ---- one-launch builder for small samples (Calpha pockets), DIFFPHAR_GRAPH=fused, OFF by default --------------
// Parity-green (bit-identical CSR, multi-wave grids included) but NOT faster: 337.6 vs 331.0 us per config-2 step against
// the three launches on one box (profiles/r05e_ab_summary.txt).  All CTAs run at the same time here, so the prefix
// has to propagate through the look-back window by window while hundreds of CTAs spin on acquire loads; the three
// short launches the stream serialises for free are cheaper than that, and most of either hides beside the encoder and
// the first projection anyway.  Kept as the measured answer to "fuse K1 into one launch".
// count -> scan -> fill in ONE kernel: the three-launch version costs ~29 us at config-2 size (N = 10 k, 158 nodes per
// sample), almost all of it launch latency and the single-CTA scan, and only ~10 us of it hide beside the first
// projection.  Here a CTA owns 32 consecutive rows, one per warp (a first version with 8 rows per warp serialised
// eight latency-bound sweeps and was slower than three launches):
//   A  every warp sweeps its row's sample once and keeps the 32-candidate ballots in a register (lane c = chunk c:
//      a sample of < ~960 nodes has at most 32 chunks), degrees go to shared memory;
//   B  warp 0 scans the 32 degrees and publishes the CTA's total; the whole CTA then sums its predecessors' status
//      words (1024 per round trip; inclusive prefixes short-cut larger grids) — CTAs are dispatched in index order and
//      never wait for a later one;
//   C  the fill replays the kept ballots: no second distance sweep, only d0 of the hits is recomputed.
// Units scheme of the segmented sum only (edge_dst / agg_src need no global edge count there); identical output.
constexpr int FUSED_ROWS = 8;                    // rows per CTA = warps per CTA: ONE row per warp (rows in flight hide the L2 round trips of a sweep).
                                                 // 256-thread CTAs: five fit beside a resident node-kernel CTA (the builder runs forked beside the first
                                                 // projection) — 1024-thread CTAs got one slot per SM there and ran in three waves
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INCL = 2ull << 62, ST_VAL = (1ull << 40) - 1ull;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(32 * FUSED_ROWS) radius_rows_fused_kernel(GraphArgs a)
{
    __shared__ int degs[FUSED_ROWS];
    __shared__ int excl[FUSED_ROWS];
    __shared__ long long base_s;
    __shared__ unsigned long long total_s, sum_s;
    __shared__ int first_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * FUSED_ROWS + wid;
    pdl_launch_dependents();
    pdl_wait();
    // ---- A: one sweep of the row's sample, ballots kept (lane c holds chunk c)
    unsigned mask = 0u;
    int found = 0, b = 0;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    int lo[2] = {0, 0}, hi[2] = {0, 0};
    if (row < a.N) {
        b = a.sample_of[row];
        xi = a.x[3 * row]; yi = a.x[3 * row + 1]; zi = a.x[3 * row + 2];
        lo[0] = a.phar_off[b]; lo[1] = a.Np + a.res_off[b];
        hi[0] = a.phar_off[b + 1]; hi[1] = a.Np + a.res_off[b + 1];
        int chunk = 0;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            for (int j0 = lo[part]; j0 < hi[part]; j0 += 32, ++chunk) {
                const int j = j0 + lane;
                bool hit = false;
                if (j < hi[part]) {
                    const float d2 = dist2_exact(xi, yi, zi, a.x[3 * j], a.x[3 * j + 1], a.x[3 * j + 2]);
                    hit = (a.cutoff < 0.f) || (__fsqrt_rn(d2) <= a.cutoff);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == chunk) mask = m;
                found += __popc(m);
            }
        }
    }
    if (lane == 0) degs[wid] = found;
    __syncthreads();
    // ---- B: CTA scan, then the CTA's base offset = sum of the predecessors' totals.  The whole CTA looks back, 256
    // predecessors per round trip at first (thread t reads CTA c - 1 - t): at config-2 size (1264 CTAs) a handful of round trips after the
    // slowest predecessor published; a 32-wide look-back would propagate the prefix 32 CTAs per round trip — all CTAs
    // run at the same time here, so there are no long-finished predecessors to short-cut to.  Larger grids stop at the
    // nearest predecessor whose INCLUSIVE prefix is known.
    if (wid == 0) {
        const int v = lane < FUSED_ROWS ? degs[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane < FUSED_ROWS) excl[lane] = incl - v;
        if (lane == 31) {
            total_s = (unsigned long long)incl;
            if (blockIdx.x > 0) st_release_u64(a.status + blockIdx.x, ST_AGG | (unsigned long long)incl);
            sum_s = 0ull; first_s = 0x7fffffff;
        }
    }
    __syncthreads();
    {
        const int t = threadIdx.x;
        unsigned long long prefix = 0ull;
        for (int jb = (int)blockIdx.x - 1; blockIdx.x > 0; jb -= 32 * FUSED_ROWS) {
            const int idx = jb - t;
            unsigned long long sv = ST_INCL;                                     // before CTA 0: inclusive prefix 0
            if (idx >= 0) { do { sv = ld_acquire_u64(a.status + idx); } while (sv == 0ull); }
            const bool is_incl = (sv >> 62) == 2ull;
            const unsigned im = __ballot_sync(0xffffffffu, is_incl);
            if (im && lane == 0) atomicMin(&first_s, t + __ffs(im) - 1);         // nearest predecessor with a known inclusive prefix
            __syncthreads();
            const int first = first_s;                                            // 0x7fffffff: none in this window
            unsigned long long c = t <= first ? (sv & ST_VAL) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && c) atomicAdd(&sum_s, c);
            __syncthreads();
            prefix = sum_s;
            if (first != 0x7fffffff) break;
        }
        if (t == 0) {
            st_release_u64(a.status + blockIdx.x, ST_INCL | (prefix + total_s));
            base_s = (long long)prefix;
            if (blockIdx.x == gridDim.x - 1) {                                   // the edge count: published clamped, see scan_rowptr_kernel
                const long long E = (long long)(prefix + total_s);
                const int Ec = (int)(E < a.ecap ? E : a.ecap);
                const_cast<int*>(a.rowptr)[a.N] = Ec;
                a.counts[0] = Ec;
                if (a.Np >= a.N) a.counts[1] = Ec;
                if (E > a.ecap && (int)E > a.counts[2]) a.counts[2] = (int)E;
            }
        }
    }
    __syncthreads();
    // ---- C: fill from the kept ballots
    if (row >= a.N) return;
    const long long base = base_s + excl[wid];
    const long long row_end = base + found;
    const RowSeg seg = row_seg(a.N, base, row_end, 0, 0);
    if (lane == 0) {
        a.agg_src[row] = row_agg_src(seg, row, base, row_end);
        const int bc = (int)(base < a.ecap ? base : a.ecap);
        const_cast<int*>(a.rowptr)[row] = bc;
        if (row == a.Np) a.counts[1] = bc;                                       // E_p: rows [0, Np) are the phar nodes
    }
    int chunk = 0, done = 0;
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        for (int j0 = lo[part]; j0 < hi[part]; j0 += 32, ++chunk) {
            const unsigned m = __shfl_sync(0xffffffffu, mask, chunk);
            if ((m >> lane) & 1u) {
                const int j = j0 + lane;
                const long long pos = base + done + __popc(m & ((1u << lane) - 1u));
                if (pos < a.ecap) {
                    a.col[pos] = j;
                    a.erow[pos] = row;
                    a.d0[pos] = dist2_exact(xi, yi, zi, a.x[3 * j], a.x[3 * j + 1], a.x[3 * j + 2]);
                    a.edst[pos] = edge_dst(seg, row, base, row_end, pos);
                }
            }
            done += __popc(m);
        }
    }
}

// ---- bucketed cell list --------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(float x, float origin, float inv, int n)
{
    const int c = (int)floorf(__fmul_rn(__fsub_rn(x, origin), inv));   // monotone in x: neighbours within the cutoff differ by <= 1
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// One CTA per sample: bounding box -> grid (cell >= 1.001 cutoff per axis, <= 16 cells per axis) -> histogram ->
// exclusive scan -> bucket fill.  Bucket order is arbitrary (atomics); the row kernel's bitmap restores order.
__global__ void __launch_bounds__(256) cell_build_kernel(GraphArgs a)
{
    __shared__ float red[6][8];
    __shared__ float grid[8];
    __shared__ int hist[CELLS_MAX + 1];
    __shared__ int wsum[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int p0 = a.phar_off[b], np = a.phar_off[b + 1] - p0;
    const int r0 = a.Np + a.res_off[b], nr = a.res_off[b + 1] - a.res_off[b];
    const int n = np + nr;
    auto node_of = [&](int i) { return i < np ? p0 + i : r0 + (i - np); };
    // The grid spans the POCKET nodes only (all nodes when the sample has none): pharmacophore points may sit far outside
    // it — with random-init weights they drift to |x| ~ 1000 A — and a bounding box over them would blow the 16 x 16 x 16
    // cell budget up to cells of > 100 A, i.e. every row tests the whole sample (measured: 5 000 warp instructions per row,
    // 230 us per count pass at config 3).  Points outside the box are clamped into the boundary cells by cell_coord,
    // which stays monotone, so neighbours within the cutoff still differ by at most one cell per axis.
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = tid + (nr > 0 ? np : 0); i < n; i += 256) {
        const int j = node_of(i);
#pragma unroll
        for (int d = 0; d < 3; ++d) { const float v = a.x[3 * j + d]; mn[d] = fminf(mn[d], v); mx[d] = fmaxf(mx[d], v); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
        if (lane == 0) { red[d][wid] = mn[d]; red[3 + d][wid] = mx[d]; }
    }
    __syncthreads();
    if (tid < 3) {
        float lo = red[tid][0], hi = red[3 + tid][0];
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, red[tid][w]); hi = fmaxf(hi, red[3 + tid][w]); }
        if (n == 0) { lo = 0.f; hi = 0.f; }
        const float ext = hi - lo;
        float cs = 1.001f * a.cutoff;
        int nc = (int)(ext / cs) + 1;
        if (nc > CELLS_DIM_MAX) { nc = CELLS_DIM_MAX; cs = 1.001f * ext / CELLS_DIM_MAX; }
        grid[tid] = lo; grid[3 + tid] = 1.0f / cs;
        red[tid][0] = __int_as_float(nc);
    }
    __syncthreads();
    const int nx = __float_as_int(red[0][0]), ny = __float_as_int(red[1][0]), nz = __float_as_int(red[2][0]);
    const int n_cells = nx * ny * nz;
    if (tid == 0) grid[6] = __int_as_float(nx | (ny << 8) | (nz << 16));
    for (int c = tid; c <= n_cells; c += 256) hist[c] = 0;
    __syncthreads();
    if (tid < 7) a.cell_grid[8 * b + tid] = grid[tid];
    auto cell_of = [&](int j) {
        const int cx = cell_coord(a.x[3 * j], grid[0], grid[3], nx);
        const int cy = cell_coord(a.x[3 * j + 1], grid[1], grid[4], ny);
        const int cz = cell_coord(a.x[3 * j + 2], grid[2], grid[5], nz);
        return (cz * ny + cy) * nx + cx;
    };
    for (int i = tid; i < n; i += 256) atomicAdd(&hist[cell_of(node_of(i))], 1);
    __syncthreads();
    // exclusive scan of hist[0 .. n_cells) in chunks of 256
    int carry = 0;
    for (int c0 = 0; c0 < n_cells; c0 += 256) {
        const int c = c0 + tid;
        const int v = c < n_cells ? hist[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < wid; ++w) before += wsum[w];
        int total = 0;
        for (int w = 0; w < 8; ++w) total += wsum[w];
        if (c < n_cells) hist[c] = carry + before + incl - v;
        carry += total;
        __syncthreads();
    }
    const int base = p0 + a.res_off[b];                    // nodes of earlier samples
    int* cs_out = a.cell_start + (size_t)b * (CELLS_MAX + 1);
    for (int c = tid; c < n_cells; c += 256) cs_out[c] = base + hist[c];
    if (tid == 0) cs_out[n_cells] = base + n;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int j = node_of(i);
        const int pos = atomicAdd(&hist[cell_of(j)], 1);
        a.cell_nodes[base + pos] = j;
    }
}

// One warp per row: hits of the 27 neighbouring buckets go into the warp's bitmap over the sample's nodes
// (local index = position in ascending node order: the sample's phar nodes, then its pocket nodes).
template <bool FILL>
__global__ void __launch_bounds__(256) radius_cells_kernel(GraphArgs a)
{
    __shared__ unsigned bitmap_s[8][CELL_SAMPLE_MAX_NODES / 32];
    __shared__ int rng_s[8][2][12];                  // per warp: prefix sums / first entries of the row's 9 candidate ranges
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned* bm = bitmap_s[wid];
    int (*rng)[12] = rng_s[wid];
    pdl_launch_dependents();
    pdl_wait();
    for (int row = blockIdx.x * 8 + wid; row < a.N; row += gridDim.x * 8) {
        const int b = a.sample_of[row];
        const int p0 = a.phar_off[b], np = a.phar_off[b + 1] - p0;
        const int r0 = a.Np + a.res_off[b], nr = a.res_off[b + 1] - a.res_off[b];
        const int n_words = (np + nr + 31) >> 5;
        const float xi = a.x[3 * row], yi = a.x[3 * row + 1], zi = a.x[3 * row + 2];
        unsigned* keep = a.row_bitmap ? a.row_bitmap + (size_t)row * a.bitmap_words : nullptr;
        if (FILL && keep) {
            // the count pass already searched the 27 buckets: its bitmap comes back from L2 (63 words per row at 2 k nodes)
            for (int w = lane; w < n_words; w += 32) bm[w] = keep[w];
            __syncwarp();
        } else {
        for (int w = lane; w < n_words; w += 32) bm[w] = 0u;
        __syncwarp();
        const float* g = a.cell_grid + 8 * b;
        const int dims = __float_as_int(g[6]);
        const int nx = dims & 255, ny = (dims >> 8) & 255, nz = dims >> 16;
        const int cx = cell_coord(xi, g[0], g[3], nx), cy = cell_coord(yi, g[1], g[4], ny), cz = cell_coord(zi, g[2], g[5], nz);
        const int* cs = a.cell_start + (size_t)b * (CELLS_MAX + 1);
        // The 27 neighbouring buckets are 9 contiguous candidate ranges (the x-neighbours are consecutive buckets: one
        // range per (y, z)).  Walking them one after the other costs three DEPENDENT L2 round trips per range (bounds ->
        // bucket entry -> coordinates), 27 per row — the pass was latency-bound at ~220 us for config 3.  Instead: lanes
        // 0..8 fetch the 9 range bounds together, a warp scan flattens the ranges into one candidate list (~300 entries at
        // full-atom density), and every lane takes 8 candidates per batch with all their bucket loads, then all their
        // coordinate loads, in flight at once: ~3 round trips per batch, one or two batches per row.
        int r_lo = 0, r_len = 0;
        if (lane < 9) {
            const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
            if (z >= 0 && z < nz && y >= 0 && y < ny) {
                const int c_lo = (z * ny + y) * nx + (cx > 0 ? cx - 1 : 0);
                const int c_hi = (z * ny + y) * nx + (cx + 1 < nx ? cx + 1 : nx - 1);
                r_lo = cs[c_lo];
                r_len = cs[c_hi + 1] - r_lo;
            }
        }
        int incl = r_len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 8);
        if (lane < 9) { rng[0][lane + 1] = incl; rng[1][lane] = r_lo; }
        if (lane == 0) rng[0][0] = 0;
        __syncwarp();
        constexpr int CB = 8;                                                    // candidates per lane and batch
        for (int c0 = 0; c0 < total; c0 += 32 * CB) {
            int jj[CB];
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                const int c = c0 + 32 * q + lane;
                jj[q] = -1;
                if (c < total) {
                    int r = 0;
                    while (c >= rng[0][r + 1]) ++r;                              // at most 8 steps over the 9 prefix sums
                    jj[q] = a.cell_nodes[rng[1][r] + (c - rng[0][r])];
                }
            }
            float d2[CB];
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                const int j = jj[q] >= 0 ? jj[q] : row;
                d2[q] = dist2_exact(xi, yi, zi, a.x[3 * j], a.x[3 * j + 1], a.x[3 * j + 2]);
            }
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                if (jj[q] >= 0 && __fsqrt_rn(d2[q]) <= a.cutoff) {
                    const int j = jj[q];
                    const int loc = j < a.Np ? j - p0 : np + (j - r0);
                    atomicOr(&bm[loc >> 5], 1u << (loc & 31));
                }
            }
        }
        __syncwarp();
        if (!FILL && keep) for (int w = lane; w < n_words; w += 32) keep[w] = bm[w];
        }
        long long base = FILL ? (long long)a.rowptr[row] : 0;
        const long long row_start = base;
        const long long row_end = FILL ? (long long)a.rowptr[row + 1] : 0;
        RowSeg seg{};
        if (FILL) {
            seg = row_seg(a.N, row_start, row_end, a.rowptr[a.N], a.n_lanes);
            if (lane == 0) a.agg_src[row] = row_agg_src(seg, row, row_start, row_end);
        }
        int found = 0;
        for (int w0 = 0; w0 < n_words; w0 += 32) {
            const int w = w0 + lane;
            unsigned bits = w < n_words ? bm[w] : 0u;
            const int cnt = __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (FILL) {
                long long pos = base + found + incl - cnt;
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int loc = (w << 5) + bit;
                    const int j = loc < np ? p0 + loc : r0 + (loc - np);
                    if (pos < a.ecap) {
                        a.col[pos] = j;
                        a.erow[pos] = row;
                        a.d0[pos] = dist2_exact(xi, yi, zi, a.x[3 * j], a.x[3 * j + 1], a.x[3 * j + 2]);
                        a.edst[pos] = edge_dst(seg, row, row_start, row_end, pos);
                    }
                    ++pos;
                }
            }
            found += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (!FILL && lane == 0) a.deg[row] = found;
        __syncwarp();
    }
    if (!FILL && a.ticket && last_block_done(a.ticket))
        scan_rowptr_block(a.deg, const_cast<int*>(a.rowptr), a.N, a.Np, a.counts, a.ecap);
}

// Exclusive scan of deg[N] -> rowptr[N+1] by ONE CTA (any block size that is a multiple of 32, at most 1024): every
// thread owns 16 consecutive rows (four 128-bit loads in flight).  Runs as the tail of the count pass: the CTA that
// finishes counting LAST (atomic ticket, threadfence on both sides) scans — one launch (and its ~2.5 us of launch
// latency on the critical side branch of every denoising step) less than a stand-alone scan kernel.
__device__ void scan_rowptr_block(const int* __restrict__ deg, int* __restrict__ rowptr, int N, int Np, int* counts, long long ecap)
{
    constexpr int PER = 16;
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
    if (tid == 0) carry_s = 0;
    const int cap = ecap < 2147483647LL ? (int)ecap : 2147483647;
    __syncthreads();
    for (int base = 0; base < N; base += nthr * PER) {
        const int i0 = base + tid * PER;
        int v[PER];
        if (i0 + PER <= N && (N & 3) == 0) {
#pragma unroll
            for (int q = 0; q < PER / 4; ++q) {
                const int4 t = *reinterpret_cast<const int4*>(deg + i0 + 4 * q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < PER; ++q) v[q] = (i0 + q < N) ? deg[i0 + q] : 0;
        }
        int sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) sum += v[q];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = lane < nwarp ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += n;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const int carry = carry_s;
        int run = carry + (wid ? warp_sums[wid - 1] : 0) + incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            if (i0 + q < N) rowptr[i0 + q] = run < cap ? run : cap;      // clamped: see the overflow note below
            run += v[q];
        }
        __syncthreads();
        if (tid == nthr - 1) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    // Overflow (more edges than the plan's capacity): every consumer indexes the per-edge arrays by rowptr / counts,
    // so both are published CLAMPED to the capacity — the truncated graph is wrong but memory-safe — and counts[2]
    // carries the true edge count (non-zero = overflow; the caller re-plans with at least that capacity).
    if (tid == 0) {
        const int E = carry_s;
        rowptr[N] = E < cap ? E : cap;
        counts[0] = E < cap ? E : cap;
        if ((long long)E > ecap && E > counts[2]) counts[2] = E;
    }
    __syncthreads();
    if (tid == 0) counts[1] = rowptr[Np];   // E_p: rows [0, Np) are the phar nodes (clamped like every rowptr entry)
}

// the count pass's tail: true in exactly one CTA, after every CTA's deg[] stores are visible to it
__device__ bool last_block_done(int* ticket)
{
    __shared__ int is_last_s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        is_last_s = t == (int)gridDim.x - 1;
        if (is_last_s) *ticket = 0;                                      // ready for the next build
    }
    __syncthreads();
    const bool last = is_last_s != 0;
    if (last) __threadfence();
    return last;
}

__global__ void __launch_bounds__(1024) scan_rowptr_kernel(const int* __restrict__ deg, int* __restrict__ rowptr,
                                                          int N, int Np, int* counts, long long ecap)
{
    pdl_launch_dependents();
    pdl_wait();
    scan_rowptr_block(deg, rowptr, N, Np, counts, ecap);
}

}  // namespace

int launch_build_edges(dp_handle* h, const float* x_dev, cudaStream_t st)
{
    Plan& p = h->plan;
    GraphArgs a;
    a.x = x_dev; a.sample_of = p.sample_of; a.phar_off = p.phar_off; a.res_off = p.res_off;
    a.N = p.N; a.Np = p.Np; a.cutoff = h->cfg.edge_cutoff;
    a.deg = p.deg; a.rowptr = p.rowptr; a.col = p.col; a.erow = p.erow; a.d0 = p.d0; a.edst = p.edst;
    a.agg_src = p.agg_src; a.n_lanes = p.n_lanes;
    a.counts = p.counts; a.ecap = p.Ecap;
    a.cell_start = p.cell_start; a.cell_nodes = p.cell_nodes; a.cell_grid = p.cell_grid;
    a.row_bitmap = p.row_bitmap; a.bitmap_words = p.bitmap_words;
    a.status = p.scan_status;
    const bool tail_scan = !(h->dbg & 128);                              // dbg bit 7: stand-alone scan kernel (A/B)
    a.ticket = tail_scan ? p.scan_ticket : nullptr;                      // zeroed at plan time, reset by the block that scans
    const int wpb = 8;
    int grid = (p.N + wpb - 1) / wpb;
    const int max_grid = h->sm_count * 16;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    prof_begin(h, PROF_GRAPH, st);
    // Lane-range segmented sum with fewer units than lanes (E < 16 x lanes: small graphs): some lanes own no unit and
    // never store their partial rows, but a row that crosses such a lane still adds them (agg_src names a lane RANGE).
    // The lane boundaries move with E from one denoiser call to the next, so those rows must not keep an older call's
    // sums: cleared here, once per graph build (1.2 MB; the lanes that do own units rewrite theirs in every launch).
    if (p.seg_lanes && p.n_lanes > 0)
        DP_CUDA(cudaMemsetAsync(p.partials, 0, (size_t)2 * p.n_lanes * H * sizeof(float), st));
    if (p.use_cells) {
        DP_CUDA(launch_kernel(h->pdl, cell_build_kernel, dim3(p.B), dim3(256), 0, st, a));
        DP_CUDA(launch_kernel(h->pdl, radius_cells_kernel<false>, dim3(grid), dim3(256), 0, st, a));
        if (!tail_scan) DP_CUDA(launch_kernel(h->pdl, scan_rowptr_kernel, dim3(1), dim3(1024), 0, st, (const int*)p.deg, p.rowptr, p.N, p.Np, p.counts, (long long)p.Ecap));
        DP_CUDA(launch_kernel(h->pdl, radius_cells_kernel<true>, dim3(grid), dim3(256), 0, st, a));
        h->launches += tail_scan ? 0 : 1;
    } else if (p.fused_graph) {
        // one launch (+ the status clear): see radius_rows_fused_kernel
        const int n_cta = (p.N + FUSED_ROWS - 1) / FUSED_ROWS;
        DP_CUDA(cudaMemsetAsync(p.scan_status, 0, (size_t)n_cta * sizeof(unsigned long long), st));
        DP_CUDA(launch_kernel(h->pdl, radius_rows_fused_kernel, dim3(n_cta), dim3(32 * FUSED_ROWS), 0, st, a));
        h->launches -= 2;
    } else {
        DP_CUDA(launch_kernel(h->pdl, radius_rows_kernel<false>, dim3(grid), dim3(256), 0, st, a));
        if (!tail_scan) DP_CUDA(launch_kernel(h->pdl, scan_rowptr_kernel, dim3(1), dim3(1024), 0, st, (const int*)p.deg, p.rowptr, p.N, p.Np, p.counts, (long long)p.Ecap));
        DP_CUDA(launch_kernel(h->pdl, radius_rows_kernel<true>, dim3(grid), dim3(256), 0, st, a));
        if (tail_scan) h->launches -= 1;
    }
    prof_end(h, st);
    h->launches += 3;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
