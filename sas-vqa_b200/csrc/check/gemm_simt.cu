// CUDA-core check GEMM with the same fused epilogues as gemm_tcgen05.cu.
// NOT on the product path: it exists so tests (and SASVQA_DEBUG_SIMT_GEMM=1 when bisecting a
// failure on the GPU box) can tell a tcgen05/TMA descriptor bug from a bug in the other kernels.
#include "../common.cuh"

namespace sasvqa {

namespace {

constexpr int TS = 32;

__global__ void gemm_simt_kernel(GemmArgs g) {
    __shared__ float As[TS][TS + 1];
    __shared__ float Bs[TS][TS + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * TS + ty;
    const int col = blockIdx.x * TS + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < g.K; k0 += TS) {
        const int ar = blockIdx.y * TS + ty, br = blockIdx.x * TS + ty;
        As[ty][tx] = (ar < g.M && k0 + tx < g.K) ? __bfloat162float(g.A[(size_t)ar * g.K + k0 + tx]) : 0.f;
        Bs[ty][tx] = (br < g.N && k0 + tx < g.K) ? __bfloat162float(g.B[(size_t)br * g.K + k0 + tx]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TS; ++k) acc = fmaf(As[ty][k], Bs[tx][k], acc);
        __syncthreads();
    }
    if (row >= g.M || col >= g.N) return;
    switch (g.epilogue) {
        case EPI_BIAS_BF16:
            g.out_bf16[(size_t)row * g.N + col] = __float2bfloat16_rn(acc + g.bias[col]);
            break;
        case EPI_BIAS_GELU_BF16:
            g.out_bf16[(size_t)row * g.N + col] = __float2bfloat16_rn(quick_gelu(acc + g.bias[col]));
            break;
        case EPI_BIAS_ERF_GELU_BF16:
            g.out_bf16[(size_t)row * g.N + col] = __float2bfloat16_rn(erf_gelu(acc + g.bias[col]));
            break;
        case EPI_BIAS_RESID_F32:
            g.out_f32[(size_t)row * g.N + col] += acc + g.bias[col];
            break;
        default: {
            const int frame = row / kPatches, patch = row - frame * kPatches;
            g.out_f32[((size_t)frame * kTokens + 1 + patch) * g.N + col] = acc + g.pos[(size_t)(1 + patch) * g.N + col];
        }
    }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
    SASVQA_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "bad GEMM shape");
    dim3 block(TS, TS), grid((g.N + TS - 1) / TS, (g.M + TS - 1) / TS);
    SASVQA_REQUIRE(grid.y <= 65535, "check GEMM: M too large (test-only kernel)");
    gemm_simt_kernel<<<grid, block, 0, stream>>>(g);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
