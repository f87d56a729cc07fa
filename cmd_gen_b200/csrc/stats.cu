// Per-"molecule" statistics of generated pharmacophore point clouds — the GPU reduction behind test.py's
// evaluation loop (reference DiffPhar/test.py:165-197): point count, distance of the cloud's centroid to a
// reference centroid, largest pairwise distance.  Coordinates arrive as doubles (the reference works on the
// float64 arrays numpy builds from the JSON lists).  One CTA per group of points; fixed-order reductions, no
// atomics: the result does not depend on the launch configuration.
#include "common.cuh"

namespace {

constexpr int STAT_THREADS = 128;

__global__ void __launch_bounds__(STAT_THREADS) pointcloud_stats_kernel(const double* __restrict__ xyz, const int* __restrict__ group_off,
                                                                         int n_groups, double rx, double ry, double rz,
                                                                         double* __restrict__ out)
{
    __shared__ double red[4][STAT_THREADS / 32];
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (g >= n_groups) return;
    const int s = group_off[g], e = group_off[g + 1], n = e - s;
    // centroid: np.mean(all_phar_coords, axis=0)  (test.py:183)
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = s + tid; i < e; i += STAT_THREADS) { sx += xyz[3 * i]; sy += xyz[3 * i + 1]; sz += xyz[3 * i + 2]; }
    // largest pairwise distance (test.py:187-192): pairs (i, j > i) dealt round-robin over the threads
    double mx = 0.0;
    const long long n_pairs = (long long)n * (n - 1) / 2;
    for (long long p = tid; p < n_pairs; p += STAT_THREADS) {
        // row i of the strictly upper triangle: the largest i with i (2n - i - 1) / 2 <= p
        long long i = (long long)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * (double)p)) * 0.5);
        while (i * (2LL * n - i - 1) / 2 > p) --i;
        while ((i + 1) * (2LL * n - i - 2) / 2 <= p) ++i;
        const long long j = p - i * (2LL * n - i - 1) / 2 + i + 1;
        const double dx = xyz[3 * (s + i)] - xyz[3 * (s + j)], dy = xyz[3 * (s + i) + 1] - xyz[3 * (s + j) + 1],
                     dz = xyz[3 * (s + i) + 2] - xyz[3 * (s + j) + 2];
        mx = fmax(mx, sqrt(dx * dx + dy * dy + dz * dz));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_down_sync(0xffffffffu, sx, o); sy += __shfl_down_sync(0xffffffffu, sy, o); sz += __shfl_down_sync(0xffffffffu, sz, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red[0][wid] = sx; red[1][wid] = sy; red[2][wid] = sz; red[3][wid] = mx; }
    __syncthreads();
    if (tid == 0) {
        double tx = 0.0, ty = 0.0, tz = 0.0, tm = 0.0;
        for (int w = 0; w < STAT_THREADS / 32; ++w) { tx += red[0][w]; ty += red[1][w]; tz += red[2][w]; tm = fmax(tm, red[3][w]); }
        double com = 0.0;
        if (n > 0) {
            const double cx = tx / n - rx, cy = ty / n - ry, cz = tz / n - rz;      // np.linalg.norm(centroid - molecule_centroid), test.py:185
            com = sqrt(cx * cx + cy * cy + cz * cz);
        }
        out[3 * g] = (double)n; out[3 * g + 1] = com; out[3 * g + 2] = tm;
    }
}

}  // namespace

extern "C" int dp_pointcloud_stats(const double* xyz_dev, const int32_t* group_off_dev, int32_t n_groups,
                                   const double* ref_centroid_host, double* out_dev, void* stream)
{
    DP_CHECK(n_groups >= 0 && (n_groups == 0 || (xyz_dev && group_off_dev && out_dev)) && ref_centroid_host, DP_ERR_INVALID,
             "dp_pointcloud_stats: null argument");
    if (n_groups == 0) return DP_OK;
    pointcloud_stats_kernel<<<n_groups, STAT_THREADS, 0, (cudaStream_t)stream>>>(xyz_dev, group_off_dev, n_groups, ref_centroid_host[0],
                                                                                ref_centroid_host[1], ref_centroid_host[2], out_dev);
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}
