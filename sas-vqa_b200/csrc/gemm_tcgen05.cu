// Encoder GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate in TMEM)
// with the encoder's epilogues fused (bias / quick_gelu / fp32 residual add / patch-embed
// scatter + position embedding).
//
// Replaces the cuBLAS/cuDNN calls behind the HF GitVisionTransformer the reference runs
// (reference: src/preprocessing/datautils/utils.py:40 -> transformers modeling_git.py:596-666,
// :461-467): q/k/v/out projections, fc1, fc2 and the patch-embedding convolution.
//
// Structure: persistent, warp-specialised, one CTA per SM, CTA PAIRS (cluster 2x1x1) driving
// tcgen05.mma.cta_group::2 -- UMMA tile 256 (M, 128 rows per CTA) x 256 (N) x 16 (K).  Each CTA
// stages its own 128 A rows and HALF of the B tile (128 of the 256 weight rows), so per 64-wide
// K block a CTA pulls 32 KiB from L2 instead of 48 KiB (the v1 single-CTA kernel was L2->SMEM
// bandwidth bound, see profiles/r01).
//   warp 0     : TMA producer (both CTAs) -- cp.async.bulk.tensor.2d.cta_group::2 loads, 128B
//                swizzle, 6-stage ring; all transaction bytes land on the LEADER CTA's mbarrier
//   warp 1     : TMEM allocator (both CTAs); in the leader CTA one lane issues the MMAs and
//                tcgen05.commit-multicasts "stage free" / "accumulator full" to both CTAs
//   warps 2..9 : epilogue (both CTAs, 128 accumulator rows each; two warps per TMEM lane quarter): tcgen05.ld -> registers ->
//                fused elementwise -> 128B-swizzled smem staging -> TMA store (bf16 outputs) or
//                TMA reduce-add (fp32 residual stream, performed in L2: x is never read by the
//                SM); accumulators are double buffered (2 x 256 TMEM columns) so the epilogue of
//                tile i overlaps the MMAs of tile i+1
#include <algorithm>

#include "common.cuh"

namespace sasvqa {

namespace {

constexpr int BLOCK_M = 128;          // accumulator rows per CTA
constexpr int PAIR_M = 256;           // rows per CTA pair = UMMA M
constexpr int BLOCK_N = 256;          // UMMA N
constexpr int HALF_N = 128;           // weight rows staged per CTA
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
#ifndef SASVQA_GEMM_STAGES
#define SASVQA_GEMM_STAGES 6
#endif
#ifndef SASVQA_GEMM_EPI_WARPS
#define SASVQA_GEMM_EPI_WARPS 8
#endif
constexpr int STAGES = SASVQA_GEMM_STAGES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;          // 512
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;     // 16 KiB
constexpr int B_STAGE_BYTES = HALF_N * BLOCK_K * 2;      // 16 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_EPI_WARPS = SASVQA_GEMM_EPI_WARPS;   // 8: two warps per TMEM lane quarter, 128 accumulator columns each
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int EPI_BUF_BYTES = 32 * 128;                  // 32 rows x 128 B, one TMA store box
#ifndef SASVQA_GEMM_EPI_BUFS
#define SASVQA_GEMM_EPI_BUFS 1
#endif
constexpr int EPI_BUFS_PER_WARP = SASVQA_GEMM_EPI_BUFS;
constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_BUFS_PER_WARP * EPI_BUF_BYTES;   // 64 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of dynamic shared memory");

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Arrive on a barrier of the peer CTA.  Plain arrive (release at CTA scope, what CUTLASS' ClusterBarrier::arrive
// emits): the only thing the waiting MMA issuer needs ordered is this warp's TMEM reads, which
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync already give.  `.release.cluster` compiled to
// MEMBAR.ALL.GPU + ERRBAR per warp per tile -- 30 % of the kernel's stall samples (profiles/r01/ncu_gemm_v4).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#ifdef SASVQA_HEAVY_ARRIVE
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (~seconds) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("sasvqa gemm: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
                   parity);
            __trap();
        }
    }
}

// 2-CTA TMA load: destination is this CTA's smem, completion bytes go to `bar_cluster` (the leader's barrier)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                                uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tcgen05_alloc_cg2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_relinquish_cg2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive (once) on the barrier at this smem offset in every CTA of `cta_mask` when all prior MMAs retire
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tcgen05_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tcgen05_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128B swizzle (PTX ISA "tcgen05 matrix
// descriptor"): start>>4 @[0,14), LBO>>4 @[16,30) (unused for one swizzle atom along K),
// SBO>>4 @[32,46) = 1024 B between 8-row groups, version=1 @[46,48), layout SWIZZLE_128B=2 @[61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor for kind::f16: D=f32 (1@[4,6)), A=B=bf16 (1@[7,10), 1@[10,13)), both K-major,
// N>>3 @[17,23), M>>4 @[24,29).  M is the PAIR's 256 rows.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                            ((uint32_t)(PAIR_M >> 4) << 24);

// packed fp32x2 (FADD2 / FMUL2 / FFMA2, same IEEE results as the scalar forms): two columns per instruction
// (fc1 epilogue 1335 -> 1450 TFLOP/s stand-alone against scalar math)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// (acc + bias) [-> quick_gelu] for two adjacent columns -> packed bf16x2.
// quick_gelu(x) = x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x)); one MUFU op (tanh.approx, abs err
// ~2^-11, far below the bf16 rounding of the output)
template <int MODE>
__device__ __forceinline__ uint32_t epi_pair(uint32_t a0, uint32_t a1, float b0, float b1) {
    uint64_t v = add2(pk2(__uint_as_float(a0), __uint_as_float(a1)), pk2(b0, b1));
    float f0, f1;
    if (MODE == EPI_BIAS_GELU_BF16) {
        float t0, t1;
        upk2(mul2(v, pk2(0.851f, 0.851f)), t0, t1);
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(t0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(t1));
        const uint64_t hx = mul2(v, pk2(0.5f, 0.5f));
        upk2(fma2(hx, pk2(t0, t1), hx), f0, f1);
    } else if (MODE == EPI_BIAS_ERF_GELU_BF16) {
        // exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) = max(x, 0) - |x / 2| erfc(|x| / sqrt 2), branch-free in packed
        // fp32x2 math: erfc(t) = 2^(t q(t)) with a degree-7 fit of q on [0, 4] (|erf error| <= 1.3e-6 over all t >= 0,
        // relative in erfc so the negative tail keeps its precision; libdevice erff in this epilogue cost fc1 25 %)
        const uint64_t ax = v & 0x7fffffff7fffffffull;
        const uint64_t t = mul2(ax, pk2(0.70710678f, 0.70710678f));            // no clamp: t q(t) keeps falling past 4
        uint64_t q = fma2(pk2(-5.904118097532773e-06f, -5.904118097532773e-06f), t, pk2(6.987361120991409e-05f, 6.987361120991409e-05f));
        q = fma2(q, t, pk2(-6.779028626624495e-05f, -6.779028626624495e-05f));
        q = fma2(q, t, pk2(-0.003477875841781497f, -0.003477875841781497f));
        q = fma2(q, t, pk2(0.030925802886486053f, 0.030925802886486053f));
        q = fma2(q, t, pk2(-0.14975078403949738f, -0.14975078403949738f));
        q = fma2(q, t, pk2(-0.9181910753250122f, -0.9181910753250122f));
        q = fma2(q, t, pk2(-1.627914547920227f, -1.627914547920227f));
        float e0, e1;
        upk2(mul2(q, t), e0, e1);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(e0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(e1));
        upk2(v, f0, f1);
        const uint64_t nh = mul2(ax, pk2(-0.5f, -0.5f));                       // -|x| / 2
        upk2(fma2(nh, pk2(e0, e1), pk2(fmaxf(f0, 0.f), fmaxf(f1, 0.f))), f0, f1);
    } else {
        upk2(v, f0, f1);
    }
    return pack_bf16x2(f0, f1);
}

// ---------------------------------------------------------------- epilogue
struct EpiParams {
    int M, N;
    const float* bias;
    const float* pos;
    float* out_f32;       // EPI_PATCH_EMBED_F32 only (direct stores with the frame/token row remap)
};

// staging layout == what a SWIZZLE_128B TMA box of 32 rows x 128 B expects (buffer 1024B-aligned)
__device__ __forceinline__ uint32_t stage_addr(uint32_t buf, int row, int chunk16) {
    return buf + (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

// bf16 outputs: 64 accumulator columns [col, col+64) of this warp's 32 rows -> one TMA store
template <int MODE>
__device__ __forceinline__ void epilogue_bf16_chunk(const EpiParams& p, const CUtensorMap* map_out, uint32_t taddr,
                                                    uint32_t buf, int row0, int col, int lane) {
    uint32_t v0[32], v1[32];
    tcgen05_ld32(taddr, v0);
    tcgen05_ld32(taddr + 32u, v1);
    tcgen05_wait_ld();
    uint32_t packed[32];
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 b = __ldg(b4 + q);
        const uint32_t* v = q < 8 ? v0 : v1;
        const int o = (q & 7) * 4;
        packed[2 * q] = epi_pair<MODE>(v[o + 0], v[o + 1], b.x, b.y);
        packed[2 * q + 1] = epi_pair<MODE>(v[o + 2], v[o + 3], b.z, b.w);
    }
    // the TMA store issued two chunks ago read this buffer; make sure it is done with it
    if (lane == 0) bulk_wait_read<EPI_BUFS_PER_WARP - 1>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st_shared_v4(stage_addr(buf, lane, j), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map_out, buf, col, row0);
        bulk_commit();
    }
}

// fp32 residual stream: 32 accumulator columns -> x[row, col..col+32) += acc + bias, added in L2 by TMA
__device__ __forceinline__ void epilogue_resid_chunk(const EpiParams& p, const CUtensorMap* map_out, uint32_t taddr,
                                                     uint32_t buf, int row0, int col, int lane) {
    uint32_t v[32];
    tcgen05_ld32(taddr, v);
    tcgen05_wait_ld();
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
    if (lane == 0) bulk_wait_read<EPI_BUFS_PER_WARP - 1>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        st_shared_v4(stage_addr(buf, lane, j), __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x),
                     __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y),
                     __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z),
                     __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_reduce_add_2d(map_out, buf, col, row0);
        bulk_commit();
    }
}

// patch embedding: x[frame*197 + 1 + patch, col..col+32) = acc + pos[1 + patch]  (rows are remapped per
// frame, so this one writes straight from registers; it is 0.7 % of the encoder's FLOPs)
__device__ __forceinline__ void epilogue_patch_chunk(const EpiParams& p, uint32_t taddr, int row, int col) {
    uint32_t v[32];
    tcgen05_ld32(taddr, v);
    tcgen05_wait_ld();
    if (row >= p.M) return;
    const int frame = row / kPatches, patch = row - frame * kPatches;
    const float4* p4 = reinterpret_cast<const float4*>(p.pos + (size_t)(1 + patch) * p.N + col);
    float* x = p.out_f32 + ((size_t)frame * kTokens + 1 + patch) * p.N + col;
#pragma unroll
    for (int q = 0; q < 4; ++q) {                               // 256-bit stores: every store fills a whole 32-byte sector
        const float4 e0 = __ldg(p4 + 2 * q), e1 = __ldg(p4 + 2 * q + 1);
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(x + 8 * q),
                     "f"(__uint_as_float(v[8 * q + 0]) + e0.x), "f"(__uint_as_float(v[8 * q + 1]) + e0.y),
                     "f"(__uint_as_float(v[8 * q + 2]) + e0.z), "f"(__uint_as_float(v[8 * q + 3]) + e0.w),
                     "f"(__uint_as_float(v[8 * q + 4]) + e1.x), "f"(__uint_as_float(v[8 * q + 5]) + e1.y),
                     "f"(__uint_as_float(v[8 * q + 6]) + e1.z), "f"(__uint_as_float(v[8 * q + 7]) + e1.w)
                     : "memory");
    }
}

// ---------------------------------------------------------------- kernel
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, EpiParams epi, int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B needs 1024B alignment
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + EPI_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m_tiles = (epi.M + PAIR_M - 1) / PAIR_M;
    const int n_tiles = epi.N / BLOCK_N;
    const int total_tiles = m_tiles * n_tiles;
    const int k_blocks = K / BLOCK_K;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
        if (MODE != EPI_PATCH_EMBED_F32) prefetch_tensormap(&map_out);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);                  // leader's producer arrive (+ bytes of both CTAs)
            mbar_init(empty_bar(s), 1);                 // one multicast tcgen05.commit
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(tmem_full_bar(a), 1);             // one multicast tcgen05.commit
            mbar_init(tmem_empty_bar(a), 2 * NUM_EPI_WARPS);   // epilogue warps of BOTH CTAs (leader's copy is used)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tcgen05_alloc_cg2(tmem_slot, TMEM_COLS);
        tcgen05_relinquish_cg2();
    }
    tcgen05_fence_before();
    cluster_sync_all();                                 // peer's barriers are initialised, TMEM is allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                const int a_row = m_blk * PAIR_M + (int)cta_rank * BLOCK_M;
                const int b_row = n_blk * BLOCK_N + (int)cta_rank * HALF_N;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * STAGE_BYTES);
                    const uint32_t leader_full = mapa_cluster(full_bar(s), 0);
                    const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                    tma_load_2d_cg2(a_dst, &map_a, kb * BLOCK_K, a_row, leader_full);
                    tma_load_2d_cg2(a_dst + A_STAGE_BYTES, &map_b, kb * BLOCK_K, b_row, leader_full);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (cta_rank == 0 && lane == 0) {
            uint32_t it = 0, t = 0;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs, ++t) {
                const int a = t & 1;
                const uint32_t aph = (t >> 1) & 1u;
                mbar_wait(tmem_empty_bar(a), aph ^ 1u);     // both CTAs' epilogues have drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);             // both CTAs' TMA bytes have landed
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_base + s * STAGE_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
                        tcgen05_mma_bf16_cg2(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                             (kb | k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit_mc(empty_bar(s), 0b11);  // stage reusable in both CTAs once these MMAs retire
                }
                tcgen05_commit_mc(tmem_full_bar(a), 0b11);  // accumulator complete -> both epilogues
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs) =====================
        const int lane_grp = warp & 3;                      // TMEM lane quarter this warp may access
        constexpr int COLS_PER_WARP = BLOCK_N / (NUM_EPI_WARPS / 4);
        const int col_lo = ((warp - 2) >> 2) * COLS_PER_WARP;   // 8 warps: warps 2-5 columns [0,128), warps 6-9 [128,256)
        const int col_hi = col_lo + COLS_PER_WARP;
        const uint32_t buf0 = epi_base + (uint32_t)((warp - 2) * EPI_BUFS_PER_WARP * EPI_BUF_BYTES);
        uint32_t t = 0, chunk_ctr = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs, ++t) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const int a = t & 1;
            const uint32_t aph = (t >> 1) & 1u;
            mbar_wait(tmem_full_bar(a), aph);
            tcgen05_fence_after();
            const int row0 = m_blk * PAIR_M + (int)cta_rank * BLOCK_M + lane_grp * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * BLOCK_N);
            if (MODE == EPI_BIAS_BF16 || MODE == EPI_BIAS_GELU_BF16 || MODE == EPI_BIAS_ERF_GELU_BF16) {
#pragma unroll 1
                for (int c = col_lo; c < col_hi; c += 64, ++chunk_ctr)
                    epilogue_bf16_chunk<MODE>(epi, &map_out, taddr + (uint32_t)c,
                                              buf0 + (chunk_ctr % EPI_BUFS_PER_WARP) * EPI_BUF_BYTES, row0, n_blk * BLOCK_N + c, lane);
            } else if (MODE == EPI_BIAS_RESID_F32) {
#pragma unroll 1
                for (int c = col_lo; c < col_hi; c += 32, ++chunk_ctr)
                    epilogue_resid_chunk(epi, &map_out, taddr + (uint32_t)c, buf0 + (chunk_ctr % EPI_BUFS_PER_WARP) * EPI_BUF_BYTES,
                                         row0, n_blk * BLOCK_N + c, lane);
            } else {
#pragma unroll 1
                for (int c = col_lo; c < col_hi; c += 32)
                    epilogue_patch_chunk(epi, taddr + (uint32_t)c, row0 + lane, n_blk * BLOCK_N + c);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_cluster(tmem_empty_bar(a), 0));
        }
        if (lane == 0) bulk_wait_read<0>();                 // staging smem must outlive the last TMA stores
    }

    tcgen05_fence_before();
    cluster_sync_all();                                     // nobody exits while its peer can still signal it
    if (warp == 1) {
        tcgen05_fence_after();
        tcgen05_dealloc_cg2(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

int encode_2d(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t rows, uint64_t cols,
              uint32_t box_cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    SASVQA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    SASVQA_REQUIRE(((uintptr_t)base & 15) == 0, "tensor base must be 16-byte aligned");
    SASVQA_REQUIRE(box_cols * elt_bytes == 128, "box must span exactly one 128-byte swizzle row");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * (uint64_t)elt_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

template <int MODE>
int launch_mode(const GemmArgs& g, const CUtensorMap* ma, const CUtensorMap* mb, const CUtensorMap* mo, int num_sms,
                cudaStream_t stream) {
    static SmemAttrCache smem_attr;
    if (int rc = smem_attr.ensure(gemm_tcgen05_kernel<MODE>, SMEM_BYTES)) return rc;
    EpiParams epi{g.M, g.N, g.bias, g.pos, g.out_f32};
    const int total = ((g.M + PAIR_M - 1) / PAIR_M) * (g.N / BLOCK_N);
    const int pairs = std::max(1, std::min(num_sms / 2, total));
    gemm_tcgen05_kernel<MODE><<<2 * pairs, NUM_THREADS, SMEM_BYTES, stream>>>(*ma, *mb, *mo, epi, g.K);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace

// operand maps: bf16 [rows, cols] row-major, box = 64 columns (128 B) x 128 rows, 128B swizzle, OOB rows read 0
int make_tensor_map_bf16_kmajor(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    SASVQA_REQUIRE(cols % BLOCK_K == 0, "GEMM K must be a multiple of 64");
    return encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, BLOCK_K, box_rows);
}

// output maps for the TMA-store epilogues: box = 32 rows x 128 B (64 bf16 or 32 fp32 columns); rows past
// `rows` are clipped by the hardware
int make_tensor_map_out(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, int is_f32) {
    if (is_f32) return encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rows, cols, 32, 32);
    return encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, 64, 32);
}

int launch_gemm_tcgen05(const GemmArgs& g, const CUtensorMap* map_a, const CUtensorMap* map_b,
                        const CUtensorMap* map_out, int num_sms, cudaStream_t stream) {
    SASVQA_REQUIRE(g.N % BLOCK_N == 0, "GEMM N must be a multiple of 256");
    SASVQA_REQUIRE(g.K % BLOCK_K == 0 && g.K >= BLOCK_K, "GEMM K must be a positive multiple of 64");
    SASVQA_REQUIRE(g.M > 0, "GEMM M must be positive");
    switch (g.epilogue) {
        case EPI_BIAS_BF16: return launch_mode<EPI_BIAS_BF16>(g, map_a, map_b, map_out, num_sms, stream);
        case EPI_BIAS_GELU_BF16: return launch_mode<EPI_BIAS_GELU_BF16>(g, map_a, map_b, map_out, num_sms, stream);
        case EPI_BIAS_RESID_F32: return launch_mode<EPI_BIAS_RESID_F32>(g, map_a, map_b, map_out, num_sms, stream);
        case EPI_PATCH_EMBED_F32: return launch_mode<EPI_PATCH_EMBED_F32>(g, map_a, map_b, map_a, num_sms, stream);
        case EPI_BIAS_ERF_GELU_BF16: return launch_mode<EPI_BIAS_ERF_GELU_BF16>(g, map_a, map_b, map_out, num_sms, stream);
        default: break;
    }
    SASVQA_REQUIRE(false, "unknown GEMM epilogue");
    return 1;
}

}  // namespace sasvqa
