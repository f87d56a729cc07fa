// K1 — radius graph as CSR.  Replaces EGNNDynamics.get_edges
// (DiffPhar/equivariant_diffusion/dynamics.py:141-147): the reference materialises a dense
// N x N adjacency (cdist + mask compare + nonzero).  Here each row scans only the nodes of
// its own sample (the only possible neighbours), one warp per row, candidates visited in
// ascending node index so ballot compaction emits columns already sorted — the (row, col)
// lexicographic order torch.where yields.  Integer outputs are bit-exact by construction.
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
    int* col; int* erow; float* d0; int* edst;
    int* counts;           // [0]=E [1]=E_p [2]=overflow
    long long ecap;
};

__device__ __forceinline__ float dist2_exact(float xi, float yi, float zi, float xj, float yj, float zj)
{
    const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Where the tcgen05 edge kernel's epilogue stores the running sum after edge `pos` (32-edge units,
// UNIT_TC), as a row of the contiguous [agg (N rows) | partials (2 per unit)] buffer: -1 = keep
// accumulating; the row lies inside one unit -> agg row; the row crosses a unit boundary -> partial row
// N + 2 unit + slot (slot 0: the segment containing the unit's first edge).
__device__ __forceinline__ int edge_dst(int n_nodes, int row, long long rs, long long re, long long pos)
{
    const bool is_end = (pos == re - 1) || ((pos & (UNIT_TC - 1)) == UNIT_TC - 1);
    if (!is_end) return -1;
    const long long u0 = pos & ~(long long)(UNIT_TC - 1);
    if (rs >= u0 && re <= u0 + UNIT_TC) return row;
    return n_nodes + (int)((pos / UNIT_TC) * 2 + (rs <= u0 ? 0 : 1));
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
                        a.edst[pos] = edge_dst(a.N, row, base, row_end, pos);
                    }
                }
                found += __popc(m);
            }
        }
        if (!FILL && lane == 0) a.deg[row] = found;
    }
}

// Exclusive scan of deg[N] -> rowptr[N+1] by one CTA (N <= a few million: microseconds).
__global__ void __launch_bounds__(1024) scan_rowptr_kernel(const int* __restrict__ deg, int* __restrict__ rowptr,
                                                          int N, int Np, int* counts, long long ecap)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int i = base + tid;
        const int v = (i < N) ? deg[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += n;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
        if (i < N) rowptr[i] = excl;
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (tid == 0) {
        const int E = carry_s;
        rowptr[N] = E;
        counts[0] = E;
        if ((long long)E > ecap) counts[2] = 1;
    }
    __syncthreads();
    if (tid == 0) counts[1] = rowptr[Np];   // E_p: rows [0, Np) are the phar nodes
}

}  // namespace

int launch_build_edges(dp_handle* h, const float* x_dev, cudaStream_t st)
{
    Plan& p = h->plan;
    GraphArgs a;
    a.x = x_dev; a.sample_of = p.sample_of; a.phar_off = p.phar_off; a.res_off = p.res_off;
    a.N = p.N; a.Np = p.Np; a.cutoff = h->cfg.edge_cutoff;
    a.deg = p.deg; a.rowptr = p.rowptr; a.col = p.col; a.erow = p.erow; a.d0 = p.d0; a.edst = p.edst;
    a.counts = p.counts; a.ecap = p.Ecap;
    const int wpb = 8;
    int grid = (p.N + wpb - 1) / wpb;
    const int max_grid = h->sm_count * 16;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    prof_begin(h, PROF_GRAPH, st);
    DP_CUDA(launch_kernel(h->pdl, radius_rows_kernel<false>, dim3(grid), dim3(256), 0, st, a));
    DP_CUDA(launch_kernel(h->pdl, scan_rowptr_kernel, dim3(1), dim3(1024), 0, st, (const int*)p.deg, p.rowptr, p.N, p.Np, p.counts, (long long)p.Ecap));
    DP_CUDA(launch_kernel(h->pdl, radius_rows_kernel<true>, dim3(grid), dim3(256), 0, st, a));
    prof_end(h, st);
    h->launches += 3;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
