// FP32 (CUDA-core FFMA) contraction path of the EGNN layers — DP_FP32 precision mode.
// Reference-grade numerics: it is the on-device yardstick the tensor-core path is checked
// against, and the mode the tight parity tests run in.
//
//   linear_f32_kernel : y = epi(x W^T + b) for the per-node GEMMs
//       - GCL.node_model (egnn_new.py:48-58) with input [h | agg] assembled on the fly
//       - the factored first edge layer: W1 [h_i; h_j; e] = W1a h_i + W1b h_j + W1c e, so the
//         two H x H products are per NODE ("projection"), not per edge (SURVEY.md hard part 3)
//   edge_f32_kernel   : GCL.edge_model + attention + segment sum (egnn_new.py:31-52) or the
//         coordinate MLP up to its per-edge scalar (egnn_new.py:87-91), one 64-edge tile at a
//         time, never materialising the [E, 2H+2] concat the reference builds.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------
// 64 x 64 output tile, 256 threads, 4 x 4 micro-tile, K step 16.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_f32_kernel(LinearArgs a)
{
    __shared__ __align__(16) float Xs[16][68];
    __shared__ __align__(16) float Ws[16][64];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * 64, col0 = blockIdx.y * 64;
    const int tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 2, lk = (tid & 3) * 4;      // X loader: row, k offset
    const int wk = tid >> 4, wc = (tid & 15) * 4;     // W loader
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int grow = row0 + lr;
    for (int k0 = 0; k0 < a.K; k0 += 16) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < a.n_rows) {
            const int k = k0 + lk;
            if (a.two_source && k >= H) xv = agg_load4(a.aggv, grow, k - H);
            else xv = *reinterpret_cast<const float4*>(a.x + (size_t)grow * a.ldx + k);
        }
        Xs[lk + 0][lr] = xv.x; Xs[lk + 1][lr] = xv.y; Xs[lk + 2][lr] = xv.z; Xs[lk + 3][lr] = xv.w;
        *reinterpret_cast<float4*>(&Ws[wk][wc]) =
            *reinterpret_cast<const float4*>(a.wt + (size_t)(k0 + wk) * a.n_out + col0 + wc);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
            const float4 wb = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float xr[4] = {xa.x, xa.y, xa.z, xa.w};
            const float wr[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = *reinterpret_cast<const float4*>(a.bias + col0 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty * 4 + i;
        if (r >= a.n_rows) continue;
        float4 o = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
        if (a.epi == 1) {
            o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w);
        } else if (a.epi == 2) {
            const float4 rv = *reinterpret_cast<const float4*>(a.resid + (size_t)r * a.ldr + col0 + tx * 4);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        *reinterpret_cast<float4*>(a.y + (size_t)r * a.ldy + col0 + tx * 4) = o;
    }
}

// ------------------------------------------------------------------------------------
// Edge tile kernel: 64 consecutive CSR edges per tile, persistent grid-stride over tiles.
// ------------------------------------------------------------------------------------
constexpr int ET = UNIT_F32;        // edges per tile == segmented-sum unit
constexpr int MS = H + 4;           // padded row stride of the tile buffer (floats)

struct EdgeSmem {
    float m[ET][MS];                // layer-1 activations, later layer-2 activations
    float ws[16][H];                // K-chunk of the second-layer weights
    int row[ET]; int col[ET]; int rs[ET]; int re[ET];
    float r2[ET]; float d0[ET]; float gate[ET];
};

__global__ void __launch_bounds__(256) edge_f32_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EdgeSmem& s = *reinterpret_cast<EdgeSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int E = *a.n_edges;
    const int n_tiles = (E + ET - 1) / ET;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int e0 = tile * ET;
        const int cnt = min(ET, E - e0);
        // ---- phase 0: per-edge metadata + current squared distance (coord2diff, egnn_new.py:265-268)
        if (tid < ET) {
            int r = 0, c = 0; float r2 = 0.f, d0 = 0.f; int rs = 0, re = 0;
            if (tid < cnt) {
                r = a.erow[e0 + tid]; c = a.ecol[e0 + tid]; d0 = a.d0[e0 + tid];
                const float dx = a.x[3 * r] - a.x[3 * c], dy = a.x[3 * r + 1] - a.x[3 * c + 1],
                            dz = a.x[3 * r + 2] - a.x[3 * c + 2];
                r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                rs = a.rowptr[r]; re = a.rowptr[r + 1];
            }
            s.row[tid] = r; s.col[tid] = c; s.r2[tid] = r2; s.d0[tid] = d0; s.rs[tid] = rs; s.re[tid] = re;
        }
        __syncthreads();
        // ---- phase 1: first layer from the pre-projected rows:  silu(Pa[row] + Pb[col] + r2*wr + d0*wd)
        {
            const int c0 = lane * 4, c1 = 128 + lane * 4;
            const float4 wr0 = *reinterpret_cast<const float4*>(a.wr + c0), wr1 = *reinterpret_cast<const float4*>(a.wr + c1);
            const float4 wd0 = *reinterpret_cast<const float4*>(a.wd + c0), wd1 = *reinterpret_cast<const float4*>(a.wd + c1);
            for (int i = wid; i < ET; i += 8) {
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                if (i < cnt) {
                    const float* pa = a.p + (size_t)s.row[i] * a.ldp + a.off_a;
                    const float* pb = a.p + (size_t)s.col[i] * a.ldp + a.off_b;
                    const float4 a0 = *reinterpret_cast<const float4*>(pa + c0), a1 = *reinterpret_cast<const float4*>(pa + c1);
                    const float4 b0 = *reinterpret_cast<const float4*>(pb + c0), b1 = *reinterpret_cast<const float4*>(pb + c1);
                    const float r2 = s.r2[i], d0 = s.d0[i];
                    o0.x = silu_f(a0.x + b0.x + r2 * wr0.x + d0 * wd0.x);
                    o0.y = silu_f(a0.y + b0.y + r2 * wr0.y + d0 * wd0.y);
                    o0.z = silu_f(a0.z + b0.z + r2 * wr0.z + d0 * wd0.z);
                    o0.w = silu_f(a0.w + b0.w + r2 * wr0.w + d0 * wd0.w);
                    o1.x = silu_f(a1.x + b1.x + r2 * wr1.x + d0 * wd1.x);
                    o1.y = silu_f(a1.y + b1.y + r2 * wr1.y + d0 * wd1.y);
                    o1.z = silu_f(a1.z + b1.z + r2 * wr1.z + d0 * wd1.z);
                    o1.w = silu_f(a1.w + b1.w + r2 * wr1.w + d0 * wd1.w);
                }
                *reinterpret_cast<float4*>(&s.m[i][c0]) = o0;
                *reinterpret_cast<float4*>(&s.m[i][c1]) = o1;
            }
        }
        __syncthreads();
        // ---- phase 2: second layer, [64 x 256] x [256 x 256]; thread = 8 edges x 8 channels (lane + 32 j)
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < H; k0 += 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * 256;              // float4 index inside the 16 x 256 chunk
                *reinterpret_cast<float4*>(&s.ws[0][0] + f * 4) =
                    *reinterpret_cast<const float4*>(a.w2t + (size_t)k0 * H + f * 4);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk += 4) {
                float w[4][8];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[q][j] = s.ws[kk + q][lane + 32 * j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 mv = *reinterpret_cast<const float4*>(&s.m[wid * 8 + i][k0 + kk]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j] = fmaf(mv.x, w[0][j], acc[i][j]);
                        acc[i][j] = fmaf(mv.y, w[1][j], acc[i][j]);
                        acc[i][j] = fmaf(mv.z, w[2][j], acc[i][j]);
                        acc[i][j] = fmaf(mv.w, w[3][j], acc[i][j]);
                    }
                }
            }
            __syncthreads();
        }
        // ---- phase 3: bias + SiLU back into the tile buffer
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b = a.b2[lane + 32 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i) s.m[wid * 8 + i][lane + 32 * j] = silu_f(acc[i][j] + b);
        }
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
        // ---- phase 5
        if (a.coord) {
            if (tid < cnt) a.escal[e0 + tid] = s.gate[tid];
        } else {
            // thread = channel; walk the tile's edges in CSR order, flush at row ends (egnn_new.py:50-52)
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
        __syncthreads();
    }
}

}  // namespace

int launch_linear_f32(dp_handle* h, const LinearArgs& a, cudaStream_t st)
{
    DP_CHECK(a.n_out % 64 == 0 && a.K % 16 == 0, DP_ERR_INVALID, "linear_f32: n_out %d / K %d not tile multiples", a.n_out, a.K);
    if (a.n_rows <= 0) return DP_OK;
    dim3 grid((a.n_rows + 63) / 64, a.n_out / 64);
    linear_f32_kernel<<<grid, 256, 0, st>>>(a);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int egnn_f32_init()
{
    DP_CUDA(cudaFuncSetAttribute(edge_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeSmem)));
    return DP_OK;
}

int launch_edge_f32(dp_handle* h, const EdgeArgs& a, cudaStream_t st)
{
    const int smem = (int)sizeof(EdgeSmem);
    const int grid = h->sm_count * 2;
    edge_f32_kernel<<<grid, 256, smem, st>>>(a);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
