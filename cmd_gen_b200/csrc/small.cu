// Prologue / epilogue / per-timestep kernels around the EGNN layers.
//   encode_nodes   : type encoders + time column + EGNN.embedding   (dynamics.py:84-99, egnn_new.py:198)
//   coord_finish   : coord_diff * scalar, row sum, masked x update    (egnn_new.py:91-103, 269-270)
//   decode         : embedding_out, drop time column, type decoders, velocity, NaN flag
//                                                                     (egnn_new.py:205, dynamics.py:110-131)
//   ddpm           : K4 — mu arithmetic + noise + per-sample COM projection
//                                                                     (conditional_model.py:136-156, 361-369, 467-475)
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------
struct EncodeArgs {
    const float* xh_phar; const float* xh_res;       // [Np][3+P], [Nr][3+R]
    const float* t_base; const int* step_idx; int row_stride; int t_stride;
    const int* sample_of;
    int N, Np, P, R, J, D;                            // D = node_nf (J or J+1)
    const float *pe0w, *pe0b, *pe2w, *pe2b;           // transposed [in][out]
    const float *re0w, *re0b, *re2w, *re2b;
    const float *embw, *embb;                         // [D][H], [H]
    float* h; float* x_in; float* x_a; float* x_b;
    int* nan_flag;                                    // [0] cleared here: first kernel of every denoiser evaluation
    float* h_base; int base_mode;                     // see launch_encode_nodes
    int skip_x;                                       // sampler (base_mode 2): x_in / x_a / x_b were written by the previous DDPM update
};

// One warp per node.  Every layer is out[o] = b[o] + sum_k in[k] w[k][o] with the outputs spread over
// lanes (o = lane + 32 m: coalesced weight reads, weights stay L1-resident) and the inputs broadcast by
// shuffle from the lanes that hold them.  Accumulation order = ascending k, like the reference's addmm.
template <int MAX_OUT32>
__device__ __forceinline__ void warp_linear(const float (&in)[4], int K, const float* __restrict__ w, const float* __restrict__ b,
                                            int n_out, int lane, float (&out)[MAX_OUT32])
{
#pragma unroll
    for (int m = 0; m < MAX_OUT32; ++m) out[m] = (lane + 32 * m < n_out) ? b[lane + 32 * m] : 0.f;
    // weights of KC consecutive inputs are fetched together (independent loads: one L2 round trip per chunk,
    // not per input); the FMA chain itself stays in ascending k
    constexpr int KC = MAX_OUT32 >= 8 ? 4 : 8;
    for (int k0 = 0; k0 < K; k0 += KC) {
        float wv[KC][MAX_OUT32];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float* wr = w + (size_t)(k0 + kk) * n_out;
#pragma unroll
            for (int m = 0; m < MAX_OUT32; ++m)
                wv[kk][m] = (k0 + kk < K && lane + 32 * m < n_out) ? wr[lane + 32 * m] : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int k = k0 + kk;
            const int src = k & 31, reg = k >> 5;
            const float v = __shfl_sync(0xffffffffu, reg == 0 ? in[0] : reg == 1 ? in[1] : reg == 2 ? in[2] : in[3], src);
            if (k < K) {
#pragma unroll
                for (int m = 0; m < MAX_OUT32; ++m)
                    if (lane + 32 * m < n_out) out[m] = fmaf(v, wv[kk][m], out[m]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) encode_nodes_kernel(EncodeArgs a)
{
    const int lane = threadIdx.x & 31;
    const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.base_mode != 1) a.nan_flag[0] = 0;
    if (node >= a.N) return;
    const bool phar = node < a.Np;
    if (a.base_mode == 1 && phar) return;
    if (a.base_mode == 2 && !phar) {
        // pocket node during sampling: features are constant, only the time column moves.  The time weight is the
        // LAST term of the reference's accumulation chain, so base + t * w_time is bit-identical to the full path.
        const float* src = a.xh_res + (size_t)(node - a.Np) * (3 + a.R);
        if (lane < 3 && !a.skip_x) {
            const float v = src[lane];
            a.x_in[3 * node + lane] = v; a.x_a[3 * node + lane] = v; a.x_b[3 * node + lane] = v;
        }
        float t = 0.f;
        if (a.D > a.J) {
            const int step = a.step_idx ? *a.step_idx : 0;
            t = a.t_base[(size_t)step * a.row_stride + (size_t)a.sample_of[node] * a.t_stride];
        }
        const float* base = a.h_base + (size_t)(node - a.Np) * H;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int c = lane + 32 * m;
            a.h[(size_t)node * H + c] = (a.D > a.J) ? fmaf(t, a.embw[(size_t)a.J * H + c], base[c]) : base[c];
        }
        return;
    }
    const int nf = phar ? a.P : a.R;
    const float* src = phar ? a.xh_phar + (size_t)node * (3 + a.P) : a.xh_res + (size_t)(node - a.Np) * (3 + a.R);
    const float *w0 = phar ? a.pe0w : a.re0w, *b0 = phar ? a.pe0b : a.re0b;
    const float *w2 = phar ? a.pe2w : a.re2w, *b2 = phar ? a.pe2b : a.re2b;
    if (lane < 3 && !a.skip_x) {
        const float v = src[lane];
        a.x_in[3 * node + lane] = v; a.x_a[3 * node + lane] = v; a.x_b[3 * node + lane] = v;
    }
    float feat[4] = {lane < nf ? src[3 + lane] : 0.f, lane + 32 < nf ? src[3 + lane + 32] : 0.f, 0.f, 0.f};
    float hid[4];
    warp_linear<4>(feat, nf, w0, b0, 2 * nf, lane, hid);
#pragma unroll
    for (int m = 0; m < 4; ++m) hid[m] = silu_f(hid[m]);
    float jo[2];
    warp_linear<2>(hid, 2 * nf, w2, b2, a.J, lane, jo);
    float joint[4] = {jo[0], jo[1], 0.f, 0.f};
    if (a.D > a.J) {                                      // time feature = last input column (dynamics.py:92-99)
        const int step = a.step_idx ? *a.step_idx : 0;
        const float t = a.t_base[(size_t)step * a.row_stride + (size_t)a.sample_of[node] * a.t_stride];
        if (lane == (a.J & 31)) joint[a.J >> 5] = t;
    }
    float o[8];
    if (a.base_mode == 1) {                               // everything but the time term, for mode 2 to finish
        warp_linear<8>(joint, a.J, a.embw, a.embb, H, lane, o);
#pragma unroll
        for (int m = 0; m < 8; ++m) a.h_base[(size_t)(node - a.Np) * H + lane + 32 * m] = o[m];
        return;
    }
    warp_linear<8>(joint, a.D, a.embw, a.embb, H, lane, o);
#pragma unroll
    for (int m = 0; m < 8; ++m) a.h[(size_t)node * H + lane + 32 * m] = o[m];
}

// ------------------------------------------------------------------------------------
struct CoordFinishArgs {
    const float* x_cur; float* x_next; const float* escal;
    const int* rowptr; const int* col;
    int Np; float norm_constant, coords_range, norm_factor; int use_tanh, mean;
};

// Eight lanes per phar row (the only rows update_coords_mask keeps, dynamics.py:105-107): lane l takes the
// row's edges l, l+8, ... in CSR order, then a fixed-order 3-step shuffle tree combines the lanes —
// deterministic, no atomics.  All loads of a lane are independent, so a row costs two L2 round trips.
__global__ void __launch_bounds__(256) coord_finish_kernel(CoordFinishArgs a)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, l = threadIdx.x & 7;
    const bool live = r < a.Np;
    pdl_launch_dependents();
    pdl_wait();
    const int s = live ? a.rowptr[r] : 0, e = live ? a.rowptr[r + 1] : 0;
    const float xi = live ? a.x_cur[3 * r] : 0.f, yi = live ? a.x_cur[3 * r + 1] : 0.f, zi = live ? a.x_cur[3 * r + 2] : 0.f;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int k = s + l; k < e; k += 8) {
        const int j = a.col[k];
        const float dx = xi - a.x_cur[3 * j], dy = yi - a.x_cur[3 * j + 1], dz = zi - a.x_cur[3 * j + 2];
        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float nrm = __fadd_rn(__fsqrt_rn(__fadd_rn(r2, 1e-8f)), a.norm_constant);
        const float v = a.escal[k];
        float tx = __fmul_rn(__fdiv_rn(dx, nrm), v), ty = __fmul_rn(__fdiv_rn(dy, nrm), v), tz = __fmul_rn(__fdiv_rn(dz, nrm), v);
        if (a.use_tanh) { tx = __fmul_rn(tx, a.coords_range); ty = __fmul_rn(ty, a.coords_range); tz = __fmul_rn(tz, a.coords_range); }
        sx = __fadd_rn(sx, tx); sy = __fadd_rn(sy, ty); sz = __fadd_rn(sz, tz);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        sx = __fadd_rn(sx, __shfl_down_sync(0xffffffffu, sx, o, 8));
        sy = __fadd_rn(sy, __shfl_down_sync(0xffffffffu, sy, o, 8));
        sz = __fadd_rn(sz, __shfl_down_sync(0xffffffffu, sz, o, 8));
    }
    if (live && l == 0) {
        const float d = a.mean ? (float)max(e - s, 1) : a.norm_factor;
        a.x_next[3 * r] = __fadd_rn(xi, __fdiv_rn(sx, d));
        a.x_next[3 * r + 1] = __fadd_rn(yi, __fdiv_rn(sy, d));
        a.x_next[3 * r + 2] = __fadd_rn(zi, __fdiv_rn(sz, d));
    }
}

// ------------------------------------------------------------------------------------
struct DecodeArgs {
    const float* h; const float* x_final; const float* x_in;
    int N, Np, P, R, J, D;
    const float *eow, *eob;                           // embedding_out transposed [H][D], [D]
    const float *pd0w, *pd0b, *pd2w, *pd2b;
    const float *rd0w, *rd0b, *rd2w, *rd2b;
    float* out_phar; float* out_res; int* nan_flag;
    int n_nodes;                                      // Np, or N when out_res != null
};

__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a)
{
    __shared__ float hrow[H];
    __shared__ float joint[64];
    __shared__ float hid[128];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    for (int node = blockIdx.x; node < a.n_nodes; node += gridDim.x) {
        const bool phar = node < a.Np;
        const int nf = phar ? a.P : a.R;
        const float *w0 = phar ? a.pd0w : a.rd0w, *b0 = phar ? a.pd0b : a.rd0b;
        const float *w2 = phar ? a.pd2w : a.rd2w, *b2 = phar ? a.pd2b : a.rd2b;
        float* out = phar ? a.out_phar + (size_t)node * (3 + a.P) : a.out_res + (size_t)(node - a.Np) * (3 + a.R);
        hrow[tid] = a.h[(size_t)node * H + tid];
        __syncthreads();
        // embedding_out rows 0..J-1 (the time column, row J, is dropped: dynamics.py:121-123)
        for (int j = wid; j < a.J; j += 8) {
            float part = 0.f;
            for (int k = lane; k < H; k += 32) part = fmaf(hrow[k], a.eow[k * a.D + j], part);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) joint[j] = part + a.eob[j];
        }
        __syncthreads();
        if (tid < 2 * nf) {
            float acc = b0[tid];
            for (int k = 0; k < a.J; ++k) acc = fmaf(joint[k], w0[k * 2 * nf + tid], acc);
            hid[tid] = silu_f(acc);
        }
        __syncthreads();
        if (tid < nf) {
            float acc = b2[tid];
            for (int k = 0; k < 2 * nf; ++k) acc = fmaf(hid[k], w2[k * nf + tid], acc);
            out[3 + tid] = acc;
        }
        if (tid < 3) {
            const float v = __fsub_rn(a.x_final[3 * node + tid], a.x_in[3 * node + tid]);   // dynamics.py:110
            out[tid] = v;
            if (v != v) a.nan_flag[0] = 1;
        }
        __syncthreads();
    }
}

__global__ void nan_fixup_kernel(float* out_phar, float* out_res, int Np, int Nr, int P, int R, int* nan_flag)
{
    // dynamics.py:129-131 — a NaN anywhere in the velocity zeroes it for the whole batch
    const bool bad = nan_flag[0] != 0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (bad) {
        if (i < Np) { out_phar[(size_t)i * (3 + P)] = 0.f; out_phar[(size_t)i * (3 + P) + 1] = 0.f; out_phar[(size_t)i * (3 + P) + 2] = 0.f; }
        if (out_res && i < Nr) { out_res[(size_t)i * (3 + R)] = 0.f; out_res[(size_t)i * (3 + R) + 1] = 0.f; out_res[(size_t)i * (3 + R) + 2] = 0.f; }
    }
    if (i == 0 && bad) nan_flag[1] += 1;
}


// ------------------------------------------------------------------------------------
struct DdpmKArgs {
    DdpmArgs d;
    const int* phar_off; const int* res_off;
    int P, R; int* nan_flag; float* stats;
    int* ticket;                                      // advance: the block that finishes last bumps *step_idx
    int Np, Nr;                                       // totals: frame strides
};

__device__ __forceinline__ void atomic_max_pos(float* addr, float v)
{   // v >= 0: float order == int order; a guard statistic, not a data-path reduction
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__global__ void __launch_bounds__(128) ddpm_kernel(DdpmKArgs k)
{
    extern __shared__ float buf[];             // [n_p][D] new values, then mean[3]
    const DdpmArgs& d = k.d;
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int D = 3 + k.P, RW = 3 + k.R;
    const int p0 = k.phar_off[b], np = k.phar_off[b + 1] - p0;
    const int r0 = k.res_off[b], nr = k.res_off[b + 1] - r0;
    int kind = d.kind; float A = d.a, C = d.c, S = d.sigma;
    const float* noise = d.noise;
    int stat_row = d.stat_index;
    int frame = -1;                                   // frame index this step writes, -1: none
    if (d.table) {
        const int step = *d.step_idx;
        if (d.frames_phar) {
            const int s = d.n_steps - 1 - step;       // conditional_model.py:428, 440-441
            if ((s * d.return_frames) % d.n_steps == 0) frame = (s * d.return_frames) / d.n_steps;
        }
        const float* row = d.table + 4 * (size_t)step;
        kind = 0; A = row[1]; C = row[2]; S = row[3];
        noise = d.noise + (size_t)(step + d.noise_step_base) * d.noise_step_stride;
        stat_row = step + d.stat_base;
    }
    const bool nan = k.nan_flag[0] != 0;
    // the sticky count behind the reference's 'Warning: detected nan, resetting EGNN output to zero.' (dynamics.py:129-131):
    // the fused sampler never runs nan_fixup_kernel, so the update that consumes the flag counts it
    if (nan && kind != 2 && b == 0 && tid == 0) k.nan_flag[1] += 1;
    float* mean = buf + (size_t)np * D;
    for (int idx = tid; idx < np * D; idx += blockDim.x) {
        const int c = idx % D;
        const size_t g = (size_t)p0 * D + idx;
        const float zt = d.z[g];
        float mu;
        if (kind == 2) mu = zt;
        else {
            float e = d.eps_hat[g];
            if (nan && c < 3) e = 0.f;
            if (kind == 0) mu = __fsub_rn(__fdiv_rn(zt, A), __fmul_rn(C, e));       // conditional_model.py:361-363
            else mu = __fmul_rn(A, __fsub_rn(zt, __fmul_rn(C, e)));                  // en_diffusion.py:161
        }
        buf[idx] = __fadd_rn(mu, __fmul_rn(S, noise[g]));                            // conditional_model.py:147
    }
    __syncthreads();
    if (tid < 3) {
        // guard statistics on the INPUT state (assert_mean_zero_with_mask, conditional_model.py:372)
        float sin_ = 0.f, amax = 0.f, tot = 0.f;
        for (int i = 0; i < np; ++i) {
            const float zi = d.z[(size_t)(p0 + i) * D + tid];
            sin_ = __fadd_rn(sin_, zi); amax = fmaxf(amax, fabsf(zi));
            tot = __fadd_rn(tot, buf[i * D + tid]);
        }
        mean[tid] = __fdiv_rn(tot, (float)max(np, 1));                               // scatter_mean
        if (k.stats && stat_row >= 0 && kind != 2) {
            atomic_max_pos(k.stats + 2 * stat_row, fabsf(sin_));
            atomic_max_pos(k.stats + 2 * stat_row + 1, amax);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < np * D; idx += blockDim.x) {
        const int c = idx % D;
        float v = buf[idx];
        if (c < 3) v = __fsub_rn(v, mean[c]);
        d.z[(size_t)p0 * D + idx] = v;
        if (d.x_in && c < 3) {                        // the next denoiser evaluation's coordinates (dynamics.py:77-81 clones them)
            const size_t o = 3 * (size_t)(p0 + idx / D) + c;
            d.x_in[o] = v; d.x_a[o] = v; d.x_b[o] = v;
        }
        if (frame >= 0)                               // unnormalize_z, en_diffusion.py:891-906
            d.frames_phar[((size_t)frame * k.Np + p0) * D + idx] =
                c < 3 ? __fmul_rn(v, d.norm_x) : __fadd_rn(__fmul_rn(v, d.norm_h), d.bias_h);
    }
    for (int idx = tid; idx < nr * 3; idx += blockDim.x) {
        const int i = idx / 3, c = idx - 3 * i;
        float* px = d.pocket + (size_t)(r0 + i) * RW + c;
        const float v = __fsub_rn(*px, mean[c]);
        *px = v;
        if (d.x_in) {
            const size_t o = 3 * (size_t)(k.Np + r0 + i) + c;
            d.x_in[o] = v; d.x_a[o] = v; d.x_b[o] = v;
        }
    }
    if (frame >= 0) {
        __syncthreads();                              // the translated coordinates of this sample's pocket rows
        for (int idx = tid; idx < nr * RW; idx += blockDim.x) {
            const int c = idx % RW;
            const float v = d.pocket[(size_t)r0 * RW + idx];
            d.frames_pocket[((size_t)frame * k.Nr + r0) * RW + idx] =
                c < 3 ? __fmul_rn(v, d.norm_x) : __fadd_rn(__fmul_rn(v, d.norm_h), d.bias_h);
        }
    }
    if (d.advance && tid == 0) {
        // every block read *step_idx before arriving here; the last one to arrive moves the sampler on
        // (a control ticket, not a data-path reduction)
        __threadfence();
        if (atomicAdd(k.ticket, 1) == (int)gridDim.x - 1) { *k.ticket = 0; *const_cast<int*>(d.step_idx) = *d.step_idx + 1; }
    }
}


// ------------------------------------------------------------------------------------
// Counter-based gaussian noise (replaces torch.randn in sample_gaussian, en_diffusion.py:946-949, when the caller
// injects none): Philox4x32-10 keyed by the seed, counter = (quad of the sample's flat [n_p][3+P] block, draw index,
// GLOBAL sample id), Box-Muller on the four words.  A sample's noise depends on its global id only — not on the batch
// it is packed into or the GPU that runs it — so a sharded pocket list reproduces the single-GPU result bit for bit.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}

__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b)
{
    const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-8f;      // (0, 1): 24 bits, never 0 or 1
    const float u2 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-8f;
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

__global__ void __launch_bounds__(128) fill_noise_kernel(float* noise, const int* phar_off, const long long* sample_ids,
                                                         int Np, int D, unsigned seed_lo, unsigned seed_hi)
{
    const int b = blockIdx.x, k = blockIdx.y;
    const int p0 = phar_off[b], n = (phar_off[b + 1] - p0) * D;
    const unsigned long long gid = (unsigned long long)sample_ids[b];
    float* dst = noise + ((size_t)k * Np + p0) * D;
    for (int q = threadIdx.x; 4 * q < n; q += blockDim.x) {
        const uint4 w = philox4x32_10(make_uint4((unsigned)q, (unsigned)k, (unsigned)gid, (unsigned)(gid >> 32)), make_uint2(seed_lo, seed_hi));
        const float2 g0 = box_muller(w.x, w.y), g1 = box_muller(w.z, w.w);
        const float v[4] = {g0.x, g0.y, g1.x, g1.y};
        for (int j = 0; j < 4; ++j)
            if (4 * q + j < n) dst[4 * q + j] = v[j];
    }
}

// mu of the initial draw: pocket COM per sample, zero features (conditional_model.py:412-414)
__global__ void __launch_bounds__(128) pocket_com_init_kernel(float* z, const float* pocket, const int* phar_off,
                                                              const int* res_off, int P, int R)
{
    __shared__ float com[3];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int D = 3 + P, RW = 3 + R;
    const int p0 = phar_off[b], np = phar_off[b + 1] - p0;
    const int r0 = res_off[b], nr = res_off[b + 1] - r0;
    if (tid < 3) {
        float tot = 0.f;
        for (int i = 0; i < nr; ++i) tot = __fadd_rn(tot, pocket[(size_t)(r0 + i) * RW + tid]);
        com[tid] = __fdiv_rn(tot, (float)max(nr, 1));
    }
    __syncthreads();
    for (int idx = tid; idx < np * D; idx += blockDim.x) {
        const int c = idx % D;
        z[(size_t)p0 * D + idx] = c < 3 ? com[c] : 0.f;
    }
}

}  // namespace

// ======================================================================================
int launch_encode_nodes(dp_handle* h, const float* xh_phar, const float* xh_res, const float* t_base,
                        const int* step_idx, int row_stride, int t_stride, int base_mode, cudaStream_t st)
{
    const Plan& p = h->plan; const DeviceWeights& w = h->w; const dp_config& c = h->cfg;
    EncodeArgs a;
    a.xh_phar = xh_phar; a.xh_res = xh_res; a.t_base = t_base; a.step_idx = step_idx;
    a.row_stride = row_stride; a.t_stride = t_stride; a.sample_of = p.sample_of;
    a.N = p.N; a.Np = p.Np; a.P = c.phar_nf; a.R = c.residue_nf; a.J = c.joint_nf;
    a.D = c.joint_nf + (c.condition_time ? 1 : 0);
    a.pe0w = w.phar_enc0.wt; a.pe0b = w.phar_enc0.b; a.pe2w = w.phar_enc2.wt; a.pe2b = w.phar_enc2.b;
    a.re0w = w.res_enc0.wt; a.re0b = w.res_enc0.b; a.re2w = w.res_enc2.wt; a.re2b = w.res_enc2.b;
    a.embw = w.emb.wt; a.embb = w.emb.b;
    a.h = p.h; a.x_in = p.x_in; a.x_a = p.x_a; a.x_b = p.x_b;
    const int grid = (p.N + 7) / 8;
    a.nan_flag = p.nan_flag; a.h_base = p.h_base; a.base_mode = base_mode; a.skip_x = base_mode == 2 ? 1 : 0;
    prof_begin(h, PROF_OTHER, st);
    DP_CUDA(launch_kernel(h->pdl, encode_nodes_kernel, dim3(grid), dim3(256), 0, st, a));
    prof_end(h, st);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_coord_finish(dp_handle* h, const float* x_cur, float* x_next, int n_moving, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    if (n_moving == 0) return DP_OK;
    CoordFinishArgs a;
    a.x_cur = x_cur; a.x_next = x_next; a.escal = p.escal; a.rowptr = p.rowptr; a.col = p.col;
    a.Np = n_moving; a.norm_constant = c.norm_constant; a.coords_range = c.coords_range;
    a.norm_factor = c.normalization_factor; a.use_tanh = c.use_tanh; a.mean = c.aggregation_mean;
    DP_CUDA(launch_kernel(h->pdl, coord_finish_kernel, dim3((n_moving * 8 + 255) / 256), dim3(256), 0, st, a));
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_decode(dp_handle* h, const float* x_final, float* out_phar, float* out_res, cudaStream_t st)
{
    const Plan& p = h->plan; const DeviceWeights& w = h->w; const dp_config& c = h->cfg;
    DecodeArgs a;
    a.h = p.h; a.x_final = x_final; a.x_in = p.x_in;
    a.N = p.N; a.Np = p.Np; a.P = c.phar_nf; a.R = c.residue_nf; a.J = c.joint_nf;
    a.D = c.joint_nf + (c.condition_time ? 1 : 0);
    a.eow = w.emb_out.wt; a.eob = w.emb_out.b;
    a.pd0w = w.phar_dec0.wt; a.pd0b = w.phar_dec0.b; a.pd2w = w.phar_dec2.wt; a.pd2b = w.phar_dec2.b;
    a.rd0w = w.res_dec0.wt; a.rd0b = w.res_dec0.b; a.rd2w = w.res_dec2.wt; a.rd2b = w.res_dec2.b;
    a.out_phar = out_phar; a.out_res = out_res; a.nan_flag = p.nan_flag;
    a.n_nodes = out_res ? p.N : p.Np;
    if (a.n_nodes == 0) return DP_OK;
    int grid = a.n_nodes < h->sm_count * 8 ? a.n_nodes : h->sm_count * 8;
    prof_begin(h, PROF_OTHER, st);
    DP_CUDA(launch_kernel(h->pdl, decode_kernel, dim3(grid), dim3(256), 0, st, a));
    prof_end(h, st);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_nan_fixup(dp_handle* h, float* out_phar, float* out_res, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    const int n = p.Np > p.Nr ? p.Np : p.Nr;
    nan_fixup_kernel<<<(n + 255) / 256 + 1, 256, 0, st>>>(out_phar, out_res, p.Np, p.Nr, c.phar_nf, c.residue_nf, p.nan_flag);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

// Joint mode only (update_pocket_coords=True): vel = remove_mean_batch(vel, mask), dynamics.py:133-136 — the mean over ALL
// nodes of a sample (phar and pocket) leaves the velocity columns.  One CTA per sample; fixed-order tree reduction.
__global__ void __launch_bounds__(256) velocity_center_kernel(float* out_phar, float* out_res, const int* phar_off, const int* res_off, int P, int R)
{
    __shared__ float red[3][256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int p0 = phar_off[b], np = phar_off[b + 1] - p0, r0 = res_off[b], nr = res_off[b + 1] - r0;
    const int PW = 3 + P, RW = 3 + R, n = np + nr;
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = tid; i < n; i += 256) {
        const float* v = i < np ? out_phar + (size_t)(p0 + i) * PW : out_res + (size_t)(r0 + i - np) * RW;
        s[0] += v[0]; s[1] += v[1]; s[2] += v[2];
    }
    for (int c = 0; c < 3; ++c) red[c][tid] = s[c];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) for (int c = 0; c < 3; ++c) red[c][tid] += red[c][tid + o];
        __syncthreads();
    }
    const float inv = 1.0f / (float)max(n, 1);
    const float m0 = red[0][0] * inv, m1 = red[1][0] * inv, m2 = red[2][0] * inv;
    for (int i = tid; i < n; i += 256) {
        float* v = i < np ? out_phar + (size_t)(p0 + i) * PW : out_res + (size_t)(r0 + i - np) * RW;
        v[0] -= m0; v[1] -= m1; v[2] -= m2;
    }
}

int launch_velocity_center(dp_handle* h, float* out_phar, float* out_res, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    if (p.B == 0) return DP_OK;
    velocity_center_kernel<<<p.B, 256, 0, st>>>(out_phar, out_res, p.phar_off, p.res_off, c.phar_nf, c.residue_nf);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_ddpm(dp_handle* h, const DdpmArgs& d, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    DdpmKArgs k;
    k.d = d; k.phar_off = p.phar_off; k.res_off = p.res_off; k.P = c.phar_nf; k.R = c.residue_nf;
    k.nan_flag = p.nan_flag; k.stats = p.stats; k.ticket = p.counts + 3; k.Np = p.Np; k.Nr = p.Nr;
    const size_t smem = ((size_t)p.max_phar * (3 + c.phar_nf) + 4) * sizeof(float);
    DP_CHECK(smem <= 48 * 1024, DP_ERR_INVALID, "ddpm: %d phar nodes in one sample exceed the shared-memory tile", p.max_phar);
    prof_begin(h, PROF_DDPM, st);
    DP_CUDA(launch_kernel(h->pdl, ddpm_kernel, dim3(p.B), dim3(128), smem, st, k));
    h->launches += 1;
    prof_end(h, st);
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_fill_noise(dp_handle* h, uint64_t seed, int n_draws, float* noise_dev, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    if (p.Np == 0) return DP_OK;
    DP_CHECK(n_draws <= 65535, DP_ERR_INVALID, "dp_fill_noise: %d draws exceed the grid's y extent", n_draws);
    fill_noise_kernel<<<dim3(p.B, n_draws), 128, 0, st>>>(noise_dev, p.phar_off, reinterpret_cast<const long long*>(p.sample_ids),
                                                          p.Np, 3 + c.phar_nf, (unsigned)seed, (unsigned)(seed >> 32));
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int launch_pocket_com_init(dp_handle* h, float* z, const float* pocket, cudaStream_t st)
{
    const Plan& p = h->plan; const dp_config& c = h->cfg;
    pocket_com_init_kernel<<<p.B, 128, 0, st>>>(z, pocket, p.phar_off, p.res_off, c.phar_nf, c.residue_nf);
    h->launches += 1;
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
