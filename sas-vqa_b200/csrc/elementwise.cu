// HBM-bound stages around the encoder GEMMs: fused uint8 -> normalised bf16 patch matrix (K1),
// LayerNorms, post-LN + token-mean + L2-normalise pooling (K3a), and the frame gathers (K5).
#include "common.cuh"

namespace sasvqa {

namespace {

// ---------------------------------------------------------------------------------------------
// K1: uint8 HWC frames -> (x/255 - mean)/std -> bf16 patch matrix [n*196, 768]
// Replaces the host image processor (reference: src/preprocessing/prefetch_loader.py:74-75) for
// 224x224 input plus the im2col of the patch-embedding conv (HF modeling_git.py:461-467, :527).
// Column order of a patch row = conv weight order: c*256 + iy*16 + ix.
// One thread: 8 pixels x RGB = 24 contiguous input bytes -> three 16-byte bf16 stores.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ frames, int n_frames,
                                                             __nv_bfloat16* __restrict__ patches) {
    const long long total = (long long)n_frames * kImg * (kImg / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int xc = (int)(i % (kImg / 8));
        const long long fy = i / (kImg / 8);
        const int y = (int)(fy % kImg);
        const long long f = fy / kImg;
        const uint2* src = reinterpret_cast<const uint2*>(frames + ((fy * kImg) + xc * 8) * 3);
        uint2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
        uint32_t w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
        float v[3][8];
#pragma unroll
        for (int b = 0; b < 24; ++b) {
            const uint32_t u = (w[b >> 2] >> ((b & 3) * 8)) & 0xffu;
            v[b % 3][b / 3] = normalize_px(u, px_mean(b % 3), px_std(b % 3));
        }
        const long long prow = f * kPatches + (y / kPatch) * kGrid + (xc >> 1);
        __nv_bfloat16* dst = patches + prow * kHidden + (y % kPatch) * kPatch + (xc & 1) * 8;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint4 o;
            o.x = pack_bf16x2(v[c][0], v[c][1]);
            o.y = pack_bf16x2(v[c][2], v[c][3]);
            o.z = pack_bf16x2(v[c][4], v[c][5]);
            o.w = pack_bf16x2(v[c][6], v[c][7]);
            *reinterpret_cast<uint4*>(dst + c * (kPatch * kPatch)) = o;
        }
    }
}

// Same patch layout from already-normalised fp32 CHW frames (the reference sampler's input type,
// src/preprocessing/datautils/utils.py:31).
__global__ void __launch_bounds__(256) patchify_f32_kernel(const float* __restrict__ frames, int n_frames,
                                                            __nv_bfloat16* __restrict__ patches) {
    const long long total = (long long)n_frames * 3 * kImg * (kImg / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int xc = (int)(i % (kImg / 8));
        long long r = i / (kImg / 8);
        const int y = (int)(r % kImg);
        r /= kImg;
        const int c = (int)(r % 3);
        const long long f = r / 3;
        const float4* src = reinterpret_cast<const float4*>(frames + ((f * 3 + c) * kImg + y) * kImg + xc * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y);
        o.w = pack_bf16x2(b.z, b.w);
        const long long prow = f * kPatches + (y / kPatch) * kGrid + (xc >> 1);
        *reinterpret_cast<uint4*>(patches + prow * kHidden + c * (kPatch * kPatch) + (y % kPatch) * kPatch +
                                  (xc & 1) * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 768 columns, one warp per row: lane holds 6 float4 (columns 4*(lane + 32 j) ..).
// Two-pass (mean, then centred variance) like torch's CPU kernel; eps 1e-5 (GitVisionConfig).
// ---------------------------------------------------------------------------------------------
struct RowRegs {
    float4 v[6];
};

__device__ __forceinline__ void row_load(const float* __restrict__ row, int lane, RowRegs& r) {
    const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int j = 0; j < 6; ++j) r.v[j] = p[lane + 32 * j];
}

__device__ __forceinline__ void row_normalize(RowRegs& r, int lane, const float* __restrict__ gamma,
                                              const float* __restrict__ beta) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (r.v[j].x + r.v[j].y) + (r.v[j].z + r.v[j].w);
    const float mean = warp_sum(s) * (1.0f / kHidden);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        r.v[j].x -= mean;
        r.v[j].y -= mean;
        r.v[j].z -= mean;
        r.v[j].w -= mean;
        q += (r.v[j].x * r.v[j].x + r.v[j].y * r.v[j].y) + (r.v[j].z * r.v[j].z + r.v[j].w * r.v[j].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kHidden) + kLnEps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j), b = __ldg(b4 + lane + 32 * j);
        r.v[j].x = r.v[j].x * rstd * g.x + b.x;
        r.v[j].y = r.v[j].y * rstd * g.y + b.y;
        r.v[j].z = r.v[j].z * rstd * g.z + b.z;
        r.v[j].w = r.v[j].w * rstd * g.w + b.w;
    }
}

// LN1 / LN2 of every block: fp32 residual stream -> bf16 GEMM operand.
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ h,
                                                              int rows, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        RowRegs r;
        row_load(x + row * kHidden, lane, r);
        row_normalize(r, lane, gamma, beta);
        uint2* dst = reinterpret_cast<uint2*>(h + row * kHidden);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            uint2 o;
            o.x = pack_bf16x2(r.v[j].x, r.v[j].y);
            o.y = pack_bf16x2(r.v[j].z, r.v[j].w);
            dst[lane + 32 * j] = o;
        }
    }
}

// fp32 -> fp32 LayerNorm (the LayerNorm of GIT's visual projection, modeling_git.py GitProjection); out may alias x.
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* x, float* out, long long rows,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        RowRegs r;
        row_load(x + row * kHidden, lane, r);
        row_normalize(r, lane, gamma, beta);
        float4* dst = reinterpret_cast<float4*>(out + row * kHidden);
#pragma unroll
        for (int j = 0; j < 6; ++j) dst[lane + 32 * j] = r.v[j];
    }
}

// pre_layrnorm (HF modeling_git.py:742), in place on the fp32 stream.  Token 0 of every frame is
// the class token: its input is class_embedding + position_embedding[0] (precomputed), the other
// 196 rows were written by the patch-embedding GEMM epilogue.
__global__ void __launch_bounds__(256) pre_layernorm_kernel(float* __restrict__ x, long long rows,
                                                             const float* __restrict__ cls_pos0,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        RowRegs r;
        row_load((row % kTokens == 0) ? cls_pos0 : x + row * kHidden, lane, r);
        row_normalize(r, lane, gamma, beta);
        float4* dst = reinterpret_cast<float4*>(x + row * kHidden);
#pragma unroll
        for (int j = 0; j < 6; ++j) dst[lane + 32 * j] = r.v[j];
    }
}

// ---------------------------------------------------------------------------------------------
// K3a: post_layernorm on all 197 tokens (HF modeling_git.py:751) + token mean (reference
// utils.py:44) + L2 normalise (utils.py:47, eps 1e-12).  One CTA per frame, 8 warps.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ feats) {
    __shared__ float part[8][kHidden];
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long frame = blockIdx.x;
    RowRegs acc;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = warp; t < kTokens; t += 8) {
        RowRegs r;
        row_load(x + (frame * kTokens + t) * kHidden, lane, r);
        row_normalize(r, lane, gamma, beta);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            acc.v[j].x += r.v[j].x;
            acc.v[j].y += r.v[j].y;
            acc.v[j].z += r.v[j].z;
            acc.v[j].w += r.v[j].w;
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) reinterpret_cast<float4*>(part[warp])[lane + 32 * j] = acc.v[j];
    __syncthreads();
    float m[3], sq = 0.f;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int c = threadIdx.x + 256 * e;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][c];
        m[e] = s * (1.0f / kTokens);
        sq += m[e] * m[e];
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll
    for (int e = 0; e < 3; ++e) feats[frame * kHidden + threadIdx.x + 256 * e] = m[e] * inv;
}

// ---------------------------------------------------------------------------------------------
// K5: gather the K selected frames of every clip as normalised fp32 CHW rows -- what the
// reference returns (`frames[res]`, utils.py:94) and stores (extract_features.py:96-97).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_u8_kernel(const uint8_t* __restrict__ clips,
                                                         const int32_t* __restrict__ idx, int B, int T, int K,
                                                         float* __restrict__ out) {
    const long long total = (long long)B * K * kImg * (kImg / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int xc = (int)(i % (kImg / 8));
        long long r = i / (kImg / 8);
        const int y = (int)(r % kImg);
        const long long bk = r / kImg;
        const long long b = bk / K;
        const int t = idx[bk];
        float v[3][8];
        if (t >= 0 && t < T) {
            const uint2* src = reinterpret_cast<const uint2*>(clips + (((b * T + t) * kImg + y) * kImg + xc * 8) * 3);
            uint2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
            uint32_t w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
#pragma unroll
            for (int q = 0; q < 24; ++q) {
                const uint32_t u = (w[q >> 2] >> ((q & 3) * 8)) & 0xffu;
                v[q % 3][q / 3] = normalize_px(u, px_mean(q % 3), px_std(q % 3));
            }
        } else {
#pragma unroll
            for (int q = 0; q < 24; ++q) v[q % 3][q / 3] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4* dst = reinterpret_cast<float4*>(out + ((bk * 3 + c) * kImg + y) * kImg + xc * 8);
            dst[0] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
            dst[1] = make_float4(v[c][4], v[c][5], v[c][6], v[c][7]);
        }
    }
}

__global__ void __launch_bounds__(256) gather_f32_kernel(const float4* __restrict__ frames,
                                                          const int32_t* __restrict__ idx, int B, int T, int K,
                                                          long long row_vec4, float4* __restrict__ out) {
    const long long total = (long long)B * K * row_vec4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long bk = i / row_vec4, e = i - bk * row_vec4;
        const long long b = bk / K;
        const int t = idx[bk];
        out[i] = (t >= 0 && t < T) ? __ldg(frames + (b * T + t) * row_vec4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

inline int grid_for(long long work_items, int block, int cap = 148 * 16) {
    long long g = (work_items + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

int launch_preprocess_u8(const uint8_t* frames_hwc, int n_frames, __nv_bfloat16* patches, cudaStream_t s) {
    if (n_frames == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)frames_hwc & 7) == 0 && ((uintptr_t)patches & 15) == 0, "unaligned buffers");
    const long long total = (long long)n_frames * kImg * (kImg / 8);
    preprocess_u8_kernel<<<grid_for(total, 256), 256, 0, s>>>(frames_hwc, n_frames, patches);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_patchify_f32(const float* frames_chw, int n_frames, __nv_bfloat16* patches, cudaStream_t s) {
    if (n_frames == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)frames_chw & 15) == 0 && ((uintptr_t)patches & 15) == 0, "unaligned buffers");
    const long long total = (long long)n_frames * 3 * kImg * (kImg / 8);
    patchify_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(frames_chw, n_frames, patches);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_pre_layernorm(float* x, int n_frames, const float* cls_pos0, const float* gamma, const float* beta,
                         cudaStream_t s) {
    if (n_frames == 0) return 0;
    const long long rows = (long long)n_frames * kTokens;
    pre_layernorm_kernel<<<grid_for(rows, 8), 256, 0, s>>>(x, rows, cls_pos0, gamma, beta);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_layernorm_bf16(const float* x, __nv_bfloat16* h, int rows, const float* gamma, const float* beta,
                          cudaStream_t s) {
    if (rows == 0) return 0;
    layernorm_bf16_kernel<<<grid_for(rows, 8), 256, 0, s>>>(x, h, rows, gamma, beta);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_layernorm_f32(const float* x, float* out, long long rows, const float* gamma, const float* beta,
                         cudaStream_t s) {
    if (rows == 0) return 0;
    layernorm_f32_kernel<<<grid_for(rows, 8), 256, 0, s>>>(x, out, rows, gamma, beta);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_pool_norm(const float* x, int n_frames, const float* gamma, const float* beta, float* feats,
                     cudaStream_t s) {
    if (n_frames == 0) return 0;
    pool_norm_kernel<<<n_frames, 256, 0, s>>>(x, gamma, beta, feats);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_u8(const uint8_t* clips, const int32_t* idx, int B, int T, int K, float* out, cudaStream_t s) {
    if (B == 0 || K == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)clips & 7) == 0 && ((uintptr_t)out & 15) == 0, "unaligned buffers");
    const long long total = (long long)B * K * kImg * (kImg / 8);
    gather_u8_kernel<<<grid_for(total, 256), 256, 0, s>>>(clips, idx, B, T, K, out);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_f32(const float* frames, const int32_t* idx, int B, int T, int K, int64_t row_elems, float* out,
                      cudaStream_t s) {
    if (B == 0 || K == 0 || row_elems == 0) return 0;
    SASVQA_REQUIRE(row_elems % 4 == 0, "row_elems must be a multiple of 4");
    SASVQA_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)out & 15) == 0, "unaligned buffers");
    const long long total = (long long)B * K * (row_elems / 4);
    gather_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(frames), idx, B, T, K,
                                                           row_elems / 4, reinterpret_cast<float4*>(out));
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
