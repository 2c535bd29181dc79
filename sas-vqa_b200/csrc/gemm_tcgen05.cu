// Encoder GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate in TMEM)
// with the encoder's epilogues fused (bias / quick_gelu / in-place fp32 residual / patch-embed
// scatter + position embedding).
//
// Replaces the cuBLAS/cuDNN calls behind the HF GitVisionTransformer the reference runs
// (reference: src/preprocessing/datautils/utils.py:40 -> transformers modeling_git.py:596-666,
// :461-467): q/k/v/out projections, fc1, fc2 and the patch-embedding convolution.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0     : TMA producer  -- cp.async.bulk.tensor 2D loads of A (128x64) and B (256x64) bf16
//                tiles, 128B-swizzled, into a 4-stage shared-memory ring (mbarrier full/empty)
//   warp 1     : TMEM allocator + MMA issuer -- one lane issues tcgen05.mma.cta_group::1.kind::f16
//                (M=128, N=256, K=16) x4 per stage; tcgen05.commit frees the stage / publishes
//                the accumulator
//   warps 2..5 : epilogue -- tcgen05.ld 32x32b.x32 of the fp32 accumulator (2 x 256 TMEM columns,
//                double buffered so the epilogue of tile i overlaps the MMAs of tile i+1),
//                fused elementwise work, vectorised global stores
#include "common.cuh"

namespace sasvqa {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;   // 512
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_WARPS = 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tcgen05_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
// N>>3 @[17,23), M>>4 @[24,29).
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                            ((uint32_t)(BLOCK_M >> 4) << 24);

// ---------------------------------------------------------------- epilogue
struct EpiParams {
    int M, N;
    int mode;
    const float* bias;
    const float* pos;
    __nv_bfloat16* out_bf16;
    float* out_f32;
};

// One thread owns one accumulator row; `v` holds 32 consecutive columns starting at `col`.
__device__ __forceinline__ void epilogue_store(const EpiParams& p, int row, int col, uint32_t (&v)[32]) {
    if (row >= p.M) return;
    if (p.mode == EPI_BIAS_BF16 || p.mode == EPI_BIAS_GELU_BF16) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
        uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.N + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
            float4 b0 = __ldg(b4 + 2 * q), b1 = __ldg(b4 + 2 * q + 1);
            f[0] = __uint_as_float(v[8 * q + 0]) + b0.x;
            f[1] = __uint_as_float(v[8 * q + 1]) + b0.y;
            f[2] = __uint_as_float(v[8 * q + 2]) + b0.z;
            f[3] = __uint_as_float(v[8 * q + 3]) + b0.w;
            f[4] = __uint_as_float(v[8 * q + 4]) + b1.x;
            f[5] = __uint_as_float(v[8 * q + 5]) + b1.y;
            f[6] = __uint_as_float(v[8 * q + 6]) + b1.z;
            f[7] = __uint_as_float(v[8 * q + 7]) + b1.w;
            if (p.mode == EPI_BIAS_GELU_BF16) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = quick_gelu(f[e]);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            dst[q] = o;
        }
    } else if (p.mode == EPI_BIAS_RESID_F32) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
        float4* x4 = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.N + col);
        float4 r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = x4[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 b = __ldg(b4 + q);
            r[q].x += __uint_as_float(v[4 * q + 0]) + b.x;
            r[q].y += __uint_as_float(v[4 * q + 1]) + b.y;
            r[q].z += __uint_as_float(v[4 * q + 2]) + b.z;
            r[q].w += __uint_as_float(v[4 * q + 3]) + b.w;
            x4[q] = r[q];
        }
    } else {  // EPI_PATCH_EMBED_F32
        const int frame = row / kPatches, patch = row - frame * kPatches;
        const float4* p4 = reinterpret_cast<const float4*>(p.pos + (size_t)(1 + patch) * p.N + col);
        float4* x4 = reinterpret_cast<float4*>(p.out_f32 + ((size_t)frame * kTokens + 1 + patch) * p.N + col);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 e = __ldg(p4 + q), o;
            o.x = __uint_as_float(v[4 * q + 0]) + e.x;
            o.y = __uint_as_float(v[4 * q + 1]) + e.y;
            o.z = __uint_as_float(v[4 * q + 2]) + e.z;
            o.w = __uint_as_float(v[4 * q + 3]) + e.w;
            x4[q] = o;
        }
    }
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, EpiParams epi,
                    int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B needs 1024B alignment
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem base slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (epi.M + BLOCK_M - 1) / BLOCK_M;
    const int n_tiles = epi.N / BLOCK_N;
    const int total_tiles = m_tiles * n_tiles;
    const int k_blocks = K / BLOCK_K;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tcgen05_alloc(tmem_slot, TMEM_COLS);
        tcgen05_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), STAGE_BYTES);
                    const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                    tma_load_2d(a_dst, &map_a, kb * BLOCK_K, m_blk * BLOCK_M, full_bar(s));
                    tma_load_2d(a_dst + A_STAGE_BYTES, &map_b, kb * BLOCK_K, n_blk * BLOCK_N, full_bar(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
                const int a = t & 1;
                const uint32_t aph = (t >> 1) & 1u;
                mbar_wait(tmem_empty_bar(a), aph ^ 1u);     // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full_bar(s), ph);             // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_base + s * STAGE_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
                        tcgen05_mma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                         (kb | k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(empty_bar(s));           // stage reusable once these MMAs retire
                }
                tcgen05_commit(tmem_full_bar(a));           // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int lane_grp = warp & 3;                      // TMEM lane quarter this warp may access
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const int a = t & 1;
            const uint32_t aph = (t >> 1) & 1u;
            mbar_wait(tmem_full_bar(a), aph);
            tcgen05_fence_after();
            const int row = m_blk * BLOCK_M + lane_grp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t v[32];
                tcgen05_ld32(taddr + (uint32_t)c, v);
                tcgen05_wait_ld();
                epilogue_store(epi, row, n_blk * BLOCK_N + c, v);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(a));
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tcgen05_dealloc(tmem_base, TMEM_COLS);
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

}  // namespace

// 2D bf16 tensor [rows, cols] row-major; box = 64 columns (128 B) x box_rows rows; 128B swizzle;
// out-of-bounds rows read as zero.
int make_tensor_map_bf16_kmajor(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    SASVQA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    SASVQA_REQUIRE(cols % BLOCK_K == 0, "GEMM K must be a multiple of 64");
    SASVQA_REQUIRE(((uintptr_t)base & 15) == 0, "tensor base must be 16-byte aligned");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

int launch_gemm_tcgen05(const GemmArgs& g, const CUtensorMap* map_a, const CUtensorMap* map_b, int num_sms,
                        cudaStream_t stream) {
    SASVQA_REQUIRE(g.N % BLOCK_N == 0, "GEMM N must be a multiple of 256");
    SASVQA_REQUIRE(g.K % BLOCK_K == 0 && g.K >= BLOCK_K, "GEMM K must be a positive multiple of 64");
    SASVQA_REQUIRE(g.M > 0, "GEMM M must be positive");
    static bool attr_set = false;
    if (!attr_set) {
        SASVQA_CUDA_CHECK(
            cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    EpiParams epi{g.M, g.N, g.epilogue, g.bias, g.pos, g.out_bf16, g.out_f32};
    const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    const int total = m_tiles * (g.N / BLOCK_N);
    const int grid = total < num_sms ? total : num_sms;
    gemm_tcgen05_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(*map_a, *map_b, epi, g.K);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
