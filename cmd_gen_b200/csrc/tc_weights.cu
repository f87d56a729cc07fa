// Weight images for the tcgen05 kernels (tc_edge.cu, tc_node.cu): every dense layer the tensor cores run
// is packed ONCE per dp_set_weights, per 16-bit format (bf16 and f16), into the exact shared-memory image
// the MMA reads — K-major SWIZZLE_128B panels of 256 output channels x 64 K (32 KB: row r at r * 128 B,
// 16-byte chunk index XOR (r % 8)) — so a panel arrives with one cp.async.bulk and needs no reformatting.
#include "tc_common.cuh"

#include <cstring>

namespace {
using namespace tc;

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

int tc_init()
{
    int rc = tc_edge_init();
    if (!rc) rc = tc_node_init();
    if (!rc) rc = tc_tf32_init();
    if (rc) return rc;
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
        // tf32 image: 32 K elements per 128-byte row, values rounded to nearest (ties away, like cvt.rna.tf32.f32)
        {
            const int np32 = L.K / 32;
            std::vector<uint32_t> img32((size_t)n_blocks * np32 * W_PANEL_BYTES / 4);
            for (int o = 0; o < L.n_out; ++o) {
                const int ob = o / 256, r = o % 256;
                for (int k = 0; k < L.K; ++k) {
                    const int kp = k / 32, kb = (k % 32) * 4;
                    const size_t off = ((size_t)ob * np32 + kp) * W_PANEL_BYTES + (size_t)r * 128 + ((((kb >> 4) ^ (r & 7))) << 4) + (kb & 15);
                    const float w = L.wt[(size_t)k * L.n_out + o] * L.tf32_scale;
                    uint32_t u; memcpy(&u, &w, 4);
                    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
                    img32[off / 4] = u;
                }
            }
            void* d = nullptr;
            DP_CUDA(cudaMalloc(&d, img32.size() * 4));
            T.allocations.push_back(d);
            DP_CUDA(cudaMemcpy(d, img32.data(), img32.size() * 4, cudaMemcpyHostToDevice));
            T.lin[id].img_tf32 = reinterpret_cast<unsigned char*>(d);
        }
    }
    // Fused node-phase launches (tc_node.cu) stream ONE contiguous panel sequence: for h version v > 0 the
    // node MLP of GCL v-1 (node_mlp.0: 8 panels, node_mlp.2: 4 panels) followed by the projection blocks of
    // the new h; v = 0 is the projection of the embedded features alone.
    const dp_config& c = h->cfg;
    const int G = c.n_layers * c.inv_sublayers;
    T.node.resize(G + 1);
    for (int v = 0; v <= G; ++v) {
        std::vector<int> parts;
        if (v > 0) { parts.push_back(4 * (v - 1) + 1); parts.push_back(4 * (v - 1) + 2); }
        parts.push_back(4 * G + c.n_layers + v);
        size_t bytes = 0;
        for (int id : parts) bytes += (size_t)(T.lin[id].K / PANEL_K) * (T.lin[id].n_out / 256) * W_PANEL_BYTES;
        T.node[v].n_panels = (int)(bytes / W_PANEL_BYTES);
        if (bytes == 0) continue;
        for (int fmt = 0; fmt < 2; ++fmt) {
            void* d = nullptr;
            DP_CUDA(cudaMalloc(&d, bytes));
            T.allocations.push_back(d);
            size_t off = 0;
            for (int id : parts) {
                const size_t b = (size_t)(T.lin[id].K / PANEL_K) * (T.lin[id].n_out / 256) * W_PANEL_BYTES;
                if (b) DP_CUDA(cudaMemcpy(reinterpret_cast<unsigned char*>(d) + off, T.lin[id].img[fmt], b, cudaMemcpyDeviceToDevice));
                off += b;
            }
            T.node[v].img[fmt] = reinterpret_cast<unsigned char*>(d);
        }
    }
    return DP_OK;
}

int tc_fmt_of(dp_handle* h, int* fmt)
{
    if (h->precision == DP_BF16) { *fmt = FMT_BF16; return DP_OK; }
    if (h->precision == DP_F16 || h->precision == DP_F16_FAST || h->precision == DP_F16_FAST32) { *fmt = FMT_F16; return DP_OK; }
    dp_set_error("precision mode %d has no 16-bit operand format (DP_TF32 runs tc_tf32.cu, DP_FP32 egnn_f32.cu)", h->precision);
    return DP_ERR_INVALID;
}
