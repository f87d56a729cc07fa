/*
 * diffphar_b200.h — C-ABI of the B200-native DiffPhar sampler hot path.
 *
 * The reference (zyrlia1018/CMD-GEN, DiffPhar/) has NO native/FFI layer: its boundary
 * is a Python nn.Module interface.  These entry points are what a reference-side
 * binding (ctypes) calls underneath the unchanged Python API.
 * Each function names the reference code it replaces (paths relative to
 * /root/reference/DiffPhar/).
 *
 * Conventions: plain pointers + sizes, no torch types.  Return 0 on success, a
 * negative dp_status otherwise; dp_last_error() returns a thread-local message.
 * Pointers suffixed _dev are device pointers on the handle's device, _host are host
 * pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * One handle per (device, stream); a handle is not thread-safe, distinct handles are.
 * There is no CPU fallback: every compute entry point fails if no sm_100 device.
 */
#ifndef DIFFPHAR_B200_H
#define DIFFPHAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_ABI_VERSION 3

typedef enum dp_status {
    DP_OK = 0,
    DP_ERR_INVALID = -1,    /* bad argument / unsupported configuration */
    DP_ERR_CUDA = -2,       /* a CUDA runtime call failed */
    DP_ERR_STATE = -3,      /* call order violated (no weights / no plan / no table) */
    DP_ERR_CAPACITY = -4    /* edge buffer overflow (see dp_plan edge_capacity) */
} dp_status;

typedef enum dp_precision {
    DP_FP32 = 0,            /* CUDA-core FFMA contractions (reference-grade numerics) */
    DP_TF32 = 1,            /* tcgen05 kind::tf32 tiles, fp32 accumulate in TMEM */
    DP_BF16 = 2,            /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate */
    DP_F16 = 3,             /* tcgen05 kind::f16 (fp16 operands), fp32 accumulate; accurate SiLU (ex2 / rcp) */
    DP_F16_FAST = 4,        /* fp16 operands; the edge kernels' first layer runs in packed f16x2 (two channels per
                               instruction) and SiLU takes the one-MUFU tanh form: the throughput mode */
    DP_F16_FAST32 = 5       /* as DP_F16_FAST with the first layer's tanh in fp32 (A/B of the f16 MUFU rate) */
} dp_precision;

/* Architecture of EGNNDynamics (ctor: equivariant_diffusion/dynamics.py:10-73,
 * egnn_new.py:159-191).  hidden_nf must be 256 and n_dims 3 (compile-time tiles). */
typedef struct dp_config {
    int32_t phar_nf;              /* 8 */
    int32_t residue_nf;           /* 20 (CA) or 11 (full-atom) */
    int32_t n_dims;               /* 3 */
    int32_t joint_nf;             /* 32 */
    int32_t hidden_nf;            /* 256 */
    int32_t n_layers;             /* EquivariantBlocks */
    int32_t inv_sublayers;        /* GCLs per block */
    int32_t attention;            /* egnn_new.py:26-29 */
    int32_t use_tanh;             /* egnn_new.py:90-91 */
    int32_t condition_time;       /* dynamics.py:45-49 */
    int32_t aggregation_mean;     /* 0: 'sum' (/normalization_factor), 1: 'mean' */
    float norm_constant;          /* coord2diff, egnn_new.py:265-271 */
    float coords_range;           /* 15: EGNN passes the raw range, egnn_new.py:187 */
    float normalization_factor;   /* 100 */
    float edge_cutoff;            /* < 0: no cutoff (fully connected inside a sample) */
    int32_t precision;            /* dp_precision */
} dp_config;

typedef struct dp_handle dp_handle;

/* Device-side guards the reference evaluates with host syncs every step
 * (dynamics.py:129-131, en_diffusion.py:919-924, conditional_model.py:450-457). */
typedef struct dp_flags {
    int32_t nan_resets;           /* denoiser calls whose velocity held a NaN (then zeroed) */
    int32_t edge_overflow;        /* 0, or the edge count of a graph build that exceeded edge_capacity (the graph was
                                     truncated to the capacity: results are invalid, re-plan with at least this many) */
    float max_mean_rel_err;       /* max over checked steps of |sum x| / (max|x| + 1e-10) */
    float last_max_cog;           /* max |sum_sample x_phar| after the final step */
    int64_t last_n_edges;         /* E of the most recent graph build */
    int64_t last_n_edges_phar;    /* E_p: edges whose row is a phar node */
    int32_t f16_range;            /* 16-bit modes only, sticky: bit 0 = a pre-projected feature left the range in which f16
                                     sums stay finite (|P| > 32 000), bit 1 = a squared distance was clamped to 60 000 by the
                                     packed-f16 first layer (edge_cutoff = None with far-apart points): the result differs
                                     from the reference — re-run in DP_TF32 / DP_FP32 (the Python mirror does) */
    int32_t reserved;
} dp_flags;

const char* dp_last_error(void);
int dp_abi_version(void);
/* number of CUDA devices with compute capability 10.x visible; 0 => nothing can run */
int dp_device_count(void);

int dp_create(const dp_config* cfg, int device, dp_handle** out);
int dp_destroy(dp_handle* h);

/* Weight ABI: flat fp32 blob in the canonical order of cmd_gen_b200/config.py:weight_spec
 * (== the reference state-dict keys under `ddpm.dynamics.`, each [out,in] row-major). */
int64_t dp_weight_count(const dp_handle* h);
int dp_set_weights(dp_handle* h, const float* blob_host, int64_t n_floats);
int dp_set_precision(dp_handle* h, int precision);
/* EGNNDynamics(update_pocket_coords=...) (dynamics.py:16, 104-107, 133-136).  0 (default): pocket-conditioning mode — only
 * the phar rows' coordinates are updated (update_coords_mask) and the pocket velocities are exactly zero.  1: the joint
 * mode — every node moves, dp_dynamics_forward returns the pocket velocities too (out_res_dev is then required) and
 * removes the per-sample mean over ALL nodes from the velocity (remove_mean_batch).  The sampler entry points
 * (dp_sample*) are ConditionalDDPM's and refuse a handle in joint mode, as the reference asserts (conditional_model.py:18). */
int dp_set_update_pocket_coords(dp_handle* h, int32_t on);

/* Batch layout.  Sample b owns phar_counts[b] pharmacophore nodes and res_counts[b]
 * pocket nodes; node index space is the reference's: all phar nodes (samples in
 * order) then all pocket nodes (dynamics.py:88-90; masks from utils.py:137-145 and
 * lightning_modules.py:447-450 are always sorted).  edge_capacity 0 = automatic. */
int dp_plan(dp_handle* h, int32_t n_samples, const int32_t* phar_counts_host,
            const int32_t* res_counts_host, int64_t edge_capacity);

/* K1 — replaces EGNNDynamics.get_edges (dynamics.py:141-147): same-sample pairs with
 * fp32 sqrt((dx*dx+dy*dy)+dz*dz) <= cutoff, self loops included, as CSR sorted by
 * (row, col).  x_dev is [N,3].  Results stay in handle-owned buffers. */
int dp_build_edges(dp_handle* h, const float* x_dev, void* stream);
/* Synchronises `stream`; returns device pointers to int32 rowptr[N+1], col[E]. */
int dp_get_graph(dp_handle* h, const int32_t** rowptr_dev, const int32_t** col_dev,
                 int64_t* n_edges, void* stream);

/* Replaces EGNNDynamics.forward (dynamics.py:75-139) in pocket-conditioning mode.
 * xh_phar_dev [N_p, 3+phar_nf], xh_res_dev [N_r, 3+residue_nf], t_dev [n_samples]
 * (t_stride 1) or one value (t_stride 0).  out_phar_dev [N_p, 3+phar_nf];
 * out_res_dev [N_r, 3+residue_nf] or NULL to skip the residue decoder. */
int dp_dynamics_forward(dp_handle* h, const float* xh_phar_dev, const float* xh_res_dev,
                        const float* t_dev, int32_t t_stride,
                        float* out_phar_dev, float* out_res_dev, void* stream);

/* K4 — replaces sample_normal_zero_com + remove_mean_batch and the mu arithmetic of
 * sample_p_zs_given_zt / sample_p_xh_given_z0 (conditional_model.py:136-156,
 * :361-369, :467-475; en_diffusion.py:153-165).  In place on z_phar_dev / xh_pocket_dev.
 *   kind 0: mu = z / a - c * eps_hat      (a = alpha_{t|s}, c = sigma2_{t|s}/alpha_{t|s}/sigma_t)
 *   kind 1: mu = a * (z - c * eps_hat)    (a = 1/alpha_0, c = sigma_0)
 *   kind 2: mu = z (eps_hat ignored)
 * then z = mu + sigma * noise; per-sample mean of the x columns removed from z and
 * from the pocket coordinates. */
int dp_ddpm_update(dp_handle* h, int32_t kind, float a, float c, float sigma,
                   float* z_phar_dev, float* xh_pocket_dev, const float* eps_hat_dev,
                   const float* noise_dev, void* stream);

/* Per-step constants computed on the host with the reference's torch op sequence
 * (cmd_gen_b200/schedule.py): rows [n_steps,4] = (t, alpha_ts, c_eps, sigma) in executed
 * order (s = n_steps-1 .. 0); final [4] = (0, 1/alpha_0, sigma_0, sigma_x). */
int dp_set_step_table(dp_handle* h, const float* rows_host, int32_t n_steps, const float* final_host);

/* Replaces the loop of ConditionalDDPM.sample_given_pocket (conditional_model.py:
 * 406-445): initial draw at the pocket COM, n_steps CUDA-graph-replayed denoising
 * steps, final p(x|z0).  xh_pocket_dev [N_r,3+residue_nf] is the NORMALISED pocket
 * (in: original frame; out: translated).  noise_dev [n_steps+2, N_p, 3+phar_nf].
 * out_phar_dev [N_p, 3+phar_nf] = (x after the final draw | z0 feature columns). */
int dp_sample(dp_handle* h, float* xh_pocket_dev, const float* noise_dev,
              float* out_phar_dev, void* stream);
/* Same through HOST buffers (H2D + D2H inside; what bench.py's e2e times). */
int dp_sample_host(dp_handle* h, const float* xh_pocket_host, const float* noise_host,
                   float* out_phar_host, float* xh_pocket_out_host);

/* Options of dp_sample_ex.  Replaces the rest of sample_given_pocket's loop body:
 *   - noise_dev == NULL: the gaussian draws of sample_gaussian (en_diffusion.py:946-949) come from a counter-based
 *     generator on the device (Philox4x32-10 + Box-Muller) keyed by `seed` and each sample's GLOBAL id, so the result
 *     of a sample does not depend on how a pocket list is batched or sharded over GPUs;
 *   - return_frames > 1: the intermediate states the reference saves every n_steps / return_frames steps
 *     (conditional_model.py:439-442), un-normalised like unnormalize_z (en_diffusion.py:891-906), written from inside
 *     the captured step; frame 0 is left for the caller to overwrite with the final result (conditional_model.py:459-460). */
typedef struct dp_sample_opts {
    const float* noise_dev;            /* [n_steps+2, N_p, 3+phar_nf] injected noise, or NULL */
    uint64_t seed;                     /* used when noise_dev is NULL */
    const int64_t* sample_ids_host;    /* [n_samples] global sample ids, or NULL: keep the current ones (0..n-1 after dp_plan) */
    int32_t return_frames;             /* <= 1: final state only */
    float norm_x, norm_h, bias_h;      /* norm_values[0], norm_values[1], norm_biases[1] */
    float* frames_phar_dev;            /* [return_frames, N_p, 3+phar_nf] (return_frames > 1) */
    float* frames_pocket_dev;          /* [return_frames, N_r, 3+residue_nf] */
} dp_sample_opts;
int dp_sample_ex(dp_handle* h, float* xh_pocket_dev, const dp_sample_opts* opts, float* out_phar_dev, void* stream);
/* The same generator into a caller buffer [n_draws, N_p, 3+phar_nf] (parity tests feed it to the CPU oracle). */
int dp_fill_noise(dp_handle* h, uint64_t seed, const int64_t* sample_ids_host, int32_t n_draws, float* noise_dev, void* stream);
/* dp_sample_host without a noise upload: the draws are generated on the device. */
int dp_sample_host_seeded(dp_handle* h, const float* xh_pocket_host, uint64_t seed, const int64_t* sample_ids_host,
                          float* out_phar_host, float* xh_pocket_out_host);
/* how many times the denoising-step CUDA graph was captured by this handle (it is re-captured only when the batch
 * layout, the precision, the step count or the frame count changes) */
int64_t dp_graph_captures(const dp_handle* h);

int dp_get_flags(dp_handle* h, dp_flags* out, void* stream);
int dp_reset_flags(dp_handle* h, void* stream);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t dp_launch_count(const dp_handle* h);
/* time of the edge-message kernel accumulated by CUDA events when enabled */
int dp_profile_enable(dp_handle* h, int32_t on);
int dp_profile_read(dp_handle* h, int32_t which, double* total_ms, int64_t* launches);

/* Evaluation statistics of generated point clouds (reference DiffPhar/test.py:165-197: per "molecule" the number of
 * points, the distance of their centroid to the reference ligand's centroid, the largest pairwise distance).
 * xyz_dev [n][3] doubles on the device, group g = rows [group_off_dev[g], group_off_dev[g + 1]);
 * ref_centroid_host [3] on the host; out_dev [n_groups][3] = {count, centroid distance, max pair distance}.
 * Needs no handle; `stream` is a cudaStream_t. */
int dp_pointcloud_stats(const double* xyz_dev, const int32_t* group_off_dev, int32_t n_groups,
                        const double* ref_centroid_host, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFPHAR_B200_H */
