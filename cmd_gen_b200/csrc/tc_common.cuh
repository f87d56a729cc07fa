// PTX wrappers and operand-layout helpers shared by the tcgen05 kernels (tc_edge.cu, tc_node.cu).
//
// Operand layout: canonical K-major SWIZZLE_128B — rows of 128 bytes (64 16-bit K elements = one
// "panel"), 8-row atoms of 1024 B, the 16-byte chunk index XORed with (row % 8).  Weights are
// pre-swizzled on the host into exactly that image (tc_weights.cu) and arrive with cp.async.bulk
// (TMA engine, SASS UBLKCP) + mbarrier complete_tx; activation tiles are produced by CUDA cores
// straight into the same layout and handed to the async proxy with fence.proxy.async.
#pragma once
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace tc {

constexpr int FMT_F16 = 0, FMT_BF16 = 1;
// Arithmetic mode of the edge kernel (operand format + how the CUDA cores compute around the MMA):
//   MODE_F16  : f16 operands, fp32 first layer, SiLU with ex2 / rcp (~1e-7)           DP_F16   "TF32-class" accuracy
//   MODE_BF16 : bf16 operands, fp32 first layer, SiLU with one tanh.approx.f32         DP_BF16
//   MODE_F16P : f16 operands, first layer in PACKED f16x2 (HADD2 / HFMA2, two channels per instruction, SiLU with
//               tanh.approx.f16x2), fp32 tanh SiLU in the epilogue                     DP_F16_FAST (bench default)
//   MODE_F16Q : as MODE_F16P but the producer's tanh runs in fp32 (A/B of the MUFU.TANH.F16 rate)   DP_F16_FAST32
constexpr int MODE_F16 = 0, MODE_BF16 = 1, MODE_F16P = 2, MODE_F16Q = 3;
__host__ __device__ constexpr int fmt_of_mode(int mode) { return mode == MODE_BF16 ? FMT_BF16 : FMT_F16; }
// which SiLU flavour silu_half<> / silu_tc<> use: every mode but the accurate f16 one takes the one-MUFU tanh form
__host__ __device__ constexpr int silu_of_mode(int mode) { return mode == MODE_F16 ? FMT_F16 : FMT_BF16; }
constexpr int PANEL_K = 64;                 // 16-bit elements per 128-byte swizzle row
constexpr int W_PANEL_BYTES = 256 * 128;    // 256 out channels x 128 B (one K panel of a 256-channel block)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// Wait of a role that is usually early (epilogue / producer warps): the poll loop of 16-24 waiting warps would
// otherwise eat a seventh of the SM's issue slots; try_wait may suspend up to the hint, then the warp backs off.
template <int SLEEP_NS = 64>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
        if (!ok) __nanosleep(SLEEP_NS);
    } while (!ok);
}
// base + index * stride_bytes in ONE instruction (IMAD.WIDE.U32): gather addresses without 64-bit add chains
__device__ __forceinline__ const void* row_ptr(const void* base, uint32_t index, uint32_t stride_bytes)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(index), "r"(stride_bytes), "l"(reinterpret_cast<unsigned long long>(base)));
    return reinterpret_cast<const void*>(r);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a converged warp (the tcgen05 / TMA issue idiom: the WHOLE warp runs the role's control flow so
// addresses and descriptors stay in uniform registers; a lane-0 branch makes the compiler wrap every
// tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall, ~90 cycles per issue)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- bulk copy global -> shared (TMA engine, 1-D) ---------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global (TMA engine); completion is tracked by the issuing THREAD's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 or bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows on the TMEM lanes, two 16-bit K elements per
// 32-bit column) stays resident in tensor memory, so one MMA only reads its B tile from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared memory -> tensor memory, asynchronous (tensor-core proxy, ordered with the same thread's later tcgen05.mma):
// 128 rows x 256 bits of the K-major operand tile the descriptor names land on lanes 0..127, 8 consecutive columns —
// the layout a .ts MMA reads its A operand in (K = 16 16-bit elements per row)
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// registers -> 32 lanes x 32 consecutive columns (thread = TMEM lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    tmem_ld32_issue(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    tmem_ld16_issue(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// 128-bit global load that does not allocate in L1 (rows used once per warp; keeps L1 for the reused Pa rows).
// Not volatile: the source is read-only for the whole launch, the scheduler may move the load freely.
__device__ __forceinline__ float4 ldg_na(const float* p)
{
    float4 v;
    asm("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// 16 bytes (8 halves) reloaded only when `take` is set: a predicated load straight into the live registers
__device__ __forceinline__ void ldg4_if(uint4& v, const void* p, bool take)
{
    asm("{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %5, 0;\n"
        "@q ld.global.v4.b32 {%0, %1, %2, %3}, [%4];\n"
        "}" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"((int)take));
}
// the same without initialised destinations: the registers hold garbage when `take` is clear — the caller must only
// use them under the same predicate (saves the two CS2R per call that zero a uint4)
__device__ __forceinline__ void ldg4_if_noinit(uint4& v, const void* p, bool take)
{
    asm("{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %5, 0;\n"
        "@q ld.global.v4.b32 {%0, %1, %2, %3}, [%4];\n"
        "}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "r"((int)take));
}
__device__ __forceinline__ uint4 ldg_u4(const void* p)
{
    uint4 v;
    asm("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_na_u4(const void* p)
{
    uint4 v;
    asm("ld.global.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 32 bytes (8 floats) reloaded only when `take` is set: predicated loads straight into the live registers,
// no branch (keeps the caller's loop one basic block, so independent edges interleave)
__device__ __forceinline__ void ldg8_if(float (&v)[8], const float* p, bool take)
{
    asm("{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %9, 0;\n"
        "@q ld.global.v4.f32 {%0, %1, %2, %3}, [%8];\n"
        "@q ld.global.v4.f32 {%4, %5, %6, %7}, [%8+16];\n"
        "}" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])
            : "l"(p), "r"((int)take));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- CTA pair (cluster of 2, tcgen05 cta_group::2) ------------------------------------------------
// One tcgen05.mma.cta_group::2 issued by the leader CTA (cluster rank 0) multiplies A = [rank 0's 128 rows ; rank 1's
// 128 rows] (M = 256) by B = [rank 0's N/2 rows ; rank 1's N/2 rows], every operand read from the SAME shared-memory
// offsets in both CTAs; CTA r's tensor memory receives rows 128 r .. 128 r + 127 of D for all N columns.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_b16(uint32_t cluster_addr, unsigned short v)
{
    asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(cluster_addr), "h"(v) : "memory");
}
// arrive on an mbarrier of any CTA of the cluster; release at cluster scope orders this thread's earlier writes
// (local and remote shared memory) before the arrival
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pure signal (no data of this thread is published with it): no fence
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// bulk copy from this CTA's shared memory into a peer CTA's (both ends through the async proxy); the bytes complete
// on an mbarrier of the DESTINATION CTA
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// bulk copy global -> the SAME shared-memory offset of every CTA in `cta_mask` (one L2 read, replicated on the way to
// the SMs); the bytes complete on the mbarrier at the same offset in each destination CTA
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, unsigned short cta_mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
// completion of THIS CTA's earlier MMAs (cta_group::1) arrives on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, unsigned short cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t holder_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of the pair's earlier MMAs arrives on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, unsigned short cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major; 1) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: [4,6) D fmt (1 = f32) | [7,10) A fmt | [10,13) B fmt | bit 15/16 A/B major
// (0 = K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N)
{
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- numerics -----------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    if (FMT == FMT_BF16) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    } else {
        __half2 t = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    }
}
__device__ __forceinline__ float tanh_approx(float v)
{
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// two f16 lanes at once: MUFU.TANH.F16 on each half + one PRMT (max abs error 2^-10.987)
__device__ __forceinline__ __half2 tanh_approx_h2(__half2 v)
{
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&v)));
    return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ float ex2_approx(float v)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rcp_approx(float v)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// SiLU(v) = v * sigmoid(v).
//   FMT_BF16: 0.5 v (1 + tanh(0.5 v)) — ONE MUFU op (tanh.approx.f32, rel. error 2^-11, far inside
//             bf16's 2^-9 operand rounding) + 2 FMA-pipe ops.  The MUFU pipe (16 lanes/clk/SM) is the
//             scarcest resource of the edge kernel: two SiLUs per edge-channel.
//   FMT_F16 : v * rcp(1 + 2^(-v log2 e)) — two MUFU ops, ~1e-7 relative: the accurate mode keeps
//             its error budget for the fp16 operand rounding.
template <int FMT>
__device__ __forceinline__ float silu_tc(float v)
{
    if (FMT == FMT_BF16) {
        const float hv = 0.5f * v;
        return fmaf(hv, tanh_approx(hv), hv);
    } else {
        return v * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v));
    }
}
// SiLU(2 hv) from the HALVED pre-activation hv (the halving is folded into the preceding FMA's constants):
//   FMT_BF16: hv + hv tanh(hv)                              -> MUFU + FFMA
//   FMT_F16 : hv / (0.5 + 0.5 * 2^(-2 log2e hv))            -> FMUL, MUFU, FFMA, MUFU, FMUL
template <int FMT>
__device__ __forceinline__ float silu_half(float hv)
{
    if (FMT == FMT_BF16) return fmaf(hv, tanh_approx(hv), hv);
    return hv * rcp_approx(fmaf(0.5f, ex2_approx(-2.8853900817779268f * hv), 0.5f));
}
__device__ __forceinline__ float sigmoid_fast(float v) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v)); }

// byte offset of the 16-byte chunk `chunk` (8 consecutive K elements) of item `i` inside a swizzled
// activation tile whose K panels are `panel_bytes` apart (panel_bytes = items * 128)
__device__ __forceinline__ uint32_t chunk_offset(int i, int chunk, int panel_bytes)
{
    return (uint32_t)((chunk >> 3) * panel_bytes + i * 128 + (((chunk & 7) ^ (i & 7)) << 4));
}

}  // namespace tc

// ---- weight images (tc_weights.cu) ----------------------------------------------------------------
struct TcLinearImg {
    unsigned char* img[2] = {nullptr, nullptr};   // swizzled 16-bit panels (64 K per 128-byte row), per format
    unsigned char* img_tf32 = nullptr;            // swizzled tf32 panels (32 K per 128-byte row), UN-scaled weights (tc_tf32.cu)
    int K = 0, n_out = 0;
};
struct TcNodeImg {                     // concatenated panel stream of one fused node-phase launch
    unsigned char* img[2] = {nullptr, nullptr};
    int n_panels = 0;
};
struct TcWeights {
    std::vector<TcLinearImg> lin;      // indexed by lin_id (see run_denoiser)
    std::vector<TcNodeImg> node;       // indexed by h version v = 0..G
    std::vector<void*> allocations;
};
int tc_fmt_of(dp_handle* h, int* fmt);
