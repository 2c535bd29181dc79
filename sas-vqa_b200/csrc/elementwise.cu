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
// The output is bf16, and for every one of the 3 x 256 possible inputs bf16(fma(u, 1/(255 std), -mean/std)) equals
// bf16 of the processor's exact (u * (1/255) - mean) / std (checked exhaustively by the parity test), so the
// streaming loop is loads, one FMA per value and stores.  One thread: the 16 pixels of one patch row = 48
// contiguous input bytes (three 16-byte loads) -> for each channel 32 contiguous output bytes (two 16-byte stores).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ frames, int n_frames,
                                                             __nv_bfloat16* __restrict__ patches) {
    const float sc[3] = {(float)(1.0 / (255.0 * (double)0.26862954f)), (float)(1.0 / (255.0 * (double)0.26130258f)),
                         (float)(1.0 / (255.0 * (double)0.27577711f))};
    const float of[3] = {(float)(-(double)0.48145466f / (double)0.26862954f), (float)(-(double)0.4578275f / (double)0.26130258f),
                         (float)(-(double)0.40821073f / (double)0.27577711f)};
    uint64_t sc2[3], of2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        asm("mov.b64 %0, {%1, %1};" : "=l"(sc2[c]) : "f"(sc[c]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(of2[c]) : "f"(of[c]));
    }
    uint64_t kNegMagic2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(kNegMagic2) : "f"(-8388608.0f));
    const long long total = (long long)n_frames * kImg * kGrid;
    auto load = [&](long long i, uint32_t (&w)[12]) {
        const int px = (int)(i % kGrid);                         // patch column
        const long long fy = i / kGrid;                          // frame * 224 + y
        const uint4* src = reinterpret_cast<const uint4*>(frames + ((fy * kImg) + px * kPatch) * 3);
        const uint4 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
        w[8] = w2.x; w[9] = w2.y; w[10] = w2.z; w[11] = w2.w;
    };
    auto emit = [&](long long i, const uint32_t (&w)[12]) {
        const int px = (int)(i % kGrid);
        const long long fy = i / kGrid;
        const int y = (int)(fy % kImg);
        const long long f = fy / kImg;
        uint32_t o[3][8];                                        // [channel][pixel pair] packed bf16x2
        // uint8 -> float without the quarter-rate I2F: PRMT drops the byte into the mantissa of 2^23 (0x4B0000uu =
        // 8388608 + u exactly), one packed FADD2 removes the 2^23, one packed FFMA2 normalises both pixels of the pair
        // (same values as (float)u and fmaf: the exhaustive parity check covers this path)
#pragma unroll
        for (int p = 0; p < 16; p += 2) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int b0 = 3 * p + c, b1 = 3 * (p + 1) + c;
                const uint32_t m0 = __byte_perm(w[b0 >> 2], 0x4B000000u, 0x7440u | (uint32_t)(b0 & 3));
                const uint32_t m1 = __byte_perm(w[b1 >> 2], 0x4B000000u, 0x7440u | (uint32_t)(b1 & 3));
                uint64_t v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(m0), "r"(m1));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(kNegMagic2));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(sc2[c]), "l"(of2[c]));
                float f0, f1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(v));
                o[c][p >> 1] = pack_bf16x2(f0, f1);
            }
        }
        const long long prow = f * kPatches + (y / kPatch) * kGrid + px;
        __nv_bfloat16* dst = patches + prow * kHidden + (y % kPatch) * kPatch;
#pragma unroll
        for (int c = 0; c < 3; ++c)                              // one full 32-byte sector per store (sm_100 256-bit st)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c * (kPatch * kPatch)),
                         "r"(o[c][0]), "r"(o[c][1]), "r"(o[c][2]), "r"(o[c][3]), "r"(o[c][4]), "r"(o[c][5]), "r"(o[c][6]),
                         "r"(o[c][7])
                         : "memory");
    };
    // two patch rows per thread per trip, both loads issued before either is consumed: the stage is latency-bound
    // under the step's power-capped SM clock, so bytes in flight per SM are what buys bandwidth
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < total; i += 2 * stride) {
        uint32_t wa[12], wb[12];
        load(i, wa);
        load(i + stride, wb);
        emit(i, wa);
        emit(i + stride, wb);
    }
    if (i < total) {
        uint32_t wa[12];
        load(i, wa);
        emit(i, wa);
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
                                              const float* __restrict__ beta, float eps = kLnEps) {
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
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kHidden) + eps);
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
// utils.py:44) + L2 normalise (utils.py:47, eps 1e-12).  One CTA of 4 warps per frame: with 16 such CTAs
// resident per SM the 2048 frames of a chunk are one wave (8-warp CTAs needed two, the second 73 % full).
// ---------------------------------------------------------------------------------------------
constexpr int POOL_WARPS = 4;
__global__ void __launch_bounds__(POOL_WARPS * 32) pool_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float* __restrict__ feats) {
    __shared__ float part[POOL_WARPS][kHidden];
    __shared__ float red[POOL_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long frame = blockIdx.x;
    RowRegs acc;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = warp; t < kTokens; t += POOL_WARPS) {
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
    constexpr int PER_THREAD = kHidden / (POOL_WARPS * 32);     // 6
    float m[PER_THREAD], sq = 0.f;
#pragma unroll
    for (int e = 0; e < PER_THREAD; ++e) {
        const int c = threadIdx.x + POOL_WARPS * 32 * e;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < POOL_WARPS; ++w) s += part[w][c];
        m[e] = s * (1.0f / kTokens);
        sq += m[e] * m[e];
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < POOL_WARPS; ++w) tot += red[w];
    const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll
    for (int e = 0; e < PER_THREAD; ++e) feats[frame * kHidden + threadIdx.x + POOL_WARPS * 32 * e] = m[e] * inv;
}

// ---------------------------------------------------------------------------------------------
// K5: gather the K selected frames of every clip as normalised fp32 CHW rows -- what the
// reference returns (`frames[res]`, utils.py:94) and stores (extract_features.py:96-97).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_u8_kernel(const uint8_t* __restrict__ clips,
                                                         const int32_t* __restrict__ idx, int B, int T, int K,
                                                         float* __restrict__ out, const int32_t* __restrict__ clip_off) {
    __shared__ float lut[3][256];                                // exact normalised value of every possible input
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8;
        lut[c][i & 255] = normalize_px((uint32_t)(i & 255), px_mean(c), px_std(c));
    }
    __syncthreads();
    const long long total = (long long)B * K * kImg * (kImg / 16);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int xc = (int)(i % (kImg / 16));                   // 16 pixels = 48 input bytes per thread
        long long r = i / (kImg / 16);
        const int y = (int)(r % kImg);
        const long long bk = r / kImg;
        const long long b = bk / K;
        const int t = idx[bk];
        const long long f0 = clip_off ? (long long)clip_off[b] : b * T;            // ragged batch: the clip's own span
        const int Tb = clip_off ? clip_off[b + 1] - clip_off[b] : T;
        const bool ok = t >= 0 && t < Tb;
        uint32_t w[12] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (ok) {
            const uint4* src = reinterpret_cast<const uint4*>(clips + (((f0 + t) * kImg + y) * kImg + xc * 16) * 3);
            const uint4 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
            w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
            w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            w[8] = w2.x; w[9] = w2.y; w[10] = w2.z; w[11] = w2.w;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int q = 3 * p + c;
                v[p] = ok ? lut[c][(w[q >> 2] >> ((q & 3) * 8)) & 0xffu] : 0.f;
            }
            float* dst = out + ((bk * 3 + c) * kImg + y) * kImg + xc * 16;
#pragma unroll
            for (int q = 0; q < 2; ++q)                          // full 32-byte sectors (sm_100 256-bit stores)
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * q), "f"(v[8 * q]),
                             "f"(v[8 * q + 1]), "f"(v[8 * q + 2]), "f"(v[8 * q + 3]), "f"(v[8 * q + 4]), "f"(v[8 * q + 5]),
                             "f"(v[8 * q + 6]), "f"(v[8 * q + 7])
                             : "memory");
        }
    }
}

__global__ void __launch_bounds__(256) gather_f32_kernel(const float4* __restrict__ frames,
                                                          const int32_t* __restrict__ idx, int B, int T, int K,
                                                          long long row_vec4, float4* __restrict__ out,
                                                          const int32_t* __restrict__ clip_off) {
    const long long total = (long long)B * K * row_vec4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long bk = i / row_vec4, e = i - bk * row_vec4;
        const long long b = bk / K;
        const int t = idx[bk];
        const long long f0 = clip_off ? (long long)clip_off[b] : b * T;
        const int Tb = clip_off ? clip_off[b + 1] - clip_off[b] : T;
        out[i] = (t >= 0 && t < Tb) ? __ldg(frames + (f0 + t) * row_vec4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}


// ---------------------------------------------------------------------------------------------
// MIF cross-encoder (BERT, post-LN; eps 1e-12): the two elementwise stages around its GEMMs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void row_store_both(const RowRegs& r, int lane, float* __restrict__ x,
                                               __nv_bfloat16* __restrict__ h) {
    float4* dx = reinterpret_cast<float4*>(x);
    uint2* dh = reinterpret_cast<uint2*>(h);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        dx[lane + 32 * j] = r.v[j];
        uint2 o;
        o.x = pack_bf16x2(r.v[j].x, r.v[j].y);
        o.y = pack_bf16x2(r.v[j].z, r.v[j].w);
        dh[lane + 32 * j] = o;
    }
}

// BertSelfOutput / BertOutput LayerNorm (modeling_bert.py:287-298, 345-356): the GEMM epilogue has already added
// dense(.) + bias into the fp32 stream, so x <- LN(x) in place, and h <- bf16(x) is the next GEMM's A operand.
__global__ void __launch_bounds__(256) layernorm_post_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ h,
                                                              long long rows, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        RowRegs r;
        row_load(x + row * kHidden, lane, r);
        row_normalize(r, lane, gamma, beta, eps);
        row_store_both(r, lane, x + row * kHidden, h + row * kHidden);
    }
}

// BertEmbeddings (modeling_bert.py:53-140): word[id] + position[p] + token_type[tt] -> LayerNorm, written to the
// PACKED row cu[s] - row_base + p of x (fp32) and h (bf16).  One warp per (sequence, position) of the padded
// [n_seqs, L] id matrix; positions at or past the sequence's length (the tokenizer's padding) are skipped.
__global__ void __launch_bounds__(256)
embed_layernorm_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ type_ids,
                       const int32_t* __restrict__ cu_seqlens, int row_base, int n_seqs, int L, int vocab, int n_types,
                       const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type_emb,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       float* __restrict__ x, __nv_bfloat16* __restrict__ h) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long total = (long long)n_seqs * L;
    for (long long i = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < total;
         i += (long long)gridDim.x * warps_per_block) {
        const int s = (int)(i / L), p = (int)(i - (long long)s * L);
        const int begin = cu_seqlens[s], len = cu_seqlens[s + 1] - begin;
        if (p >= len) continue;
        int id = ids[i], tt = type_ids ? type_ids[i] : 0;
        id = min(max(id, 0), vocab - 1);                     // host side validates; never read out of the table
        tt = min(max(tt, 0), n_types - 1);
        RowRegs r, a, b;
        row_load(word + (long long)id * kHidden, lane, r);
        row_load(pos + (long long)p * kHidden, lane, a);
        row_load(type_emb + (long long)tt * kHidden, lane, b);
#pragma unroll
        for (int j = 0; j < 6; ++j) {                        // (word + type) + position, the order HF adds them
            r.v[j].x = (r.v[j].x + b.v[j].x) + a.v[j].x;
            r.v[j].y = (r.v[j].y + b.v[j].y) + a.v[j].y;
            r.v[j].z = (r.v[j].z + b.v[j].z) + a.v[j].z;
            r.v[j].w = (r.v[j].w + b.v[j].w) + a.v[j].w;
        }
        row_normalize(r, lane, gamma, beta, eps);
        const long long row = (long long)(begin - row_base) + p;
        row_store_both(r, lane, x + row * kHidden, h + row * kHidden);
    }
}

// BertPooler + classifier (modeling_bert.py:456-468, 1077-1155): logits[s] = Wc tanh(Wp x[cls(s)] + bp) + bc in
// fp32.  One CTA per 16 sequences so that the 2.4 MB pooler matrix is streamed from L2 once per 16 rows (this
// stage is bound by that L2 -> SM traffic: 8 rows per CTA took 4.5 ms per c3x step): each warp owns 96 of the 768
// pooled outputs (lanes split the 768-long dot products, coalesced float4 weight loads), then the warps share out
// the sequences for the `labels` logits.  Dynamic shared memory: cls | pooled, 2 x 16 x 768 fp32 = 96 KiB.
constexpr int POOL_SEQS = 16;
constexpr int POOLER_SMEM = 2 * POOL_SEQS * kHidden * (int)sizeof(float);
__global__ void __launch_bounds__(256)
pooler_classifier_kernel(const float* __restrict__ x, const int32_t* __restrict__ cu_seqlens, int row_base, int n_seqs,
                         const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ wc,
                         const float* __restrict__ bc, int labels, float* __restrict__ logits) {
    extern __shared__ __align__(16) float pool_smem[];
    float (*cls)[kHidden] = reinterpret_cast<float (*)[kHidden]>(pool_smem);
    float (*pooled)[kHidden] = reinterpret_cast<float (*)[kHidden]>(pool_smem + POOL_SEQS * kHidden);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x * POOL_SEQS;
    for (int i = threadIdx.x; i < POOL_SEQS * (kHidden / 4); i += blockDim.x) {
        const int q = i / (kHidden / 4), c = i - q * (kHidden / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s0 + q < n_seqs && cu_seqlens[s0 + q + 1] > cu_seqlens[s0 + q])
            v = reinterpret_cast<const float4*>(x + (long long)(cu_seqlens[s0 + q] - row_base) * kHidden)[c];
        reinterpret_cast<float4*>(cls[q])[c] = v;
    }
    __syncthreads();
    for (int j = warp * (kHidden / 8); j < (warp + 1) * (kHidden / 8); ++j) {
        const float4* w4 = reinterpret_cast<const float4*>(wp + (long long)j * kHidden);
        float acc[POOL_SEQS];
#pragma unroll
        for (int q = 0; q < POOL_SEQS; ++q) acc[q] = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float4 w = __ldg(w4 + lane + 32 * c);
#pragma unroll
            for (int q = 0; q < POOL_SEQS; ++q) {
                const float4 v = reinterpret_cast<const float4*>(cls[q])[lane + 32 * c];
                acc[q] = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, acc[q]))));
            }
        }
#pragma unroll
        for (int q = 0; q < POOL_SEQS; ++q) acc[q] = warp_sum(acc[q]);
        if (lane < POOL_SEQS) {
            float a = acc[0];
#pragma unroll
            for (int q = 1; q < POOL_SEQS; ++q) a = lane == q ? acc[q] : a;
            pooled[lane][j] = tanhf(a + __ldg(bp + j));
        }
    }
    __syncthreads();
    for (int q = warp; q < POOL_SEQS && s0 + q < n_seqs; q += 8) {
        for (int c = 0; c < labels; ++c) {
            float a = 0.f;
            for (int k = lane; k < kHidden; k += 32) a = fmaf(__ldg(wc + (long long)c * kHidden + k), pooled[q][k], a);
            a = warp_sum(a);
            if (lane == 0) logits[(long long)(s0 + q) * labels + c] = a + __ldg(bc + c);
        }
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
    SASVQA_REQUIRE(((uintptr_t)frames_hwc & 15) == 0 && ((uintptr_t)patches & 31) == 0, "unaligned buffers");
    const long long total = (long long)n_frames * kImg * kGrid;
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
    pool_norm_kernel<<<n_frames, POOL_WARPS * 32, 0, s>>>(x, gamma, beta, feats);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_u8(const uint8_t* clips, const int32_t* idx, int B, int T, int K, float* out, cudaStream_t s,
                     const int32_t* clip_off) {
    if (B == 0 || K == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)clips & 15) == 0 && ((uintptr_t)out & 31) == 0, "unaligned buffers");
    const long long total = (long long)B * K * kImg * (kImg / 16);
    gather_u8_kernel<<<grid_for(total, 256), 256, 0, s>>>(clips, idx, B, T, K, out, clip_off);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gather_f32(const float* frames, const int32_t* idx, int B, int T, int K, int64_t row_elems, float* out,
                      cudaStream_t s, const int32_t* clip_off) {
    if (B == 0 || K == 0 || row_elems == 0) return 0;
    SASVQA_REQUIRE(row_elems % 4 == 0, "row_elems must be a multiple of 4");
    SASVQA_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)out & 15) == 0, "unaligned buffers");
    const long long total = (long long)B * K * (row_elems / 4);
    gather_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(frames), idx, B, T, K,
                                                           row_elems / 4, reinterpret_cast<float4*>(out), clip_off);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa

namespace sasvqa {

int launch_layernorm_post(float* x, __nv_bfloat16* h, long long rows, const float* gamma, const float* beta, float eps,
                          cudaStream_t s) {
    if (rows == 0) return 0;
    layernorm_post_kernel<<<grid_for(rows, 8), 256, 0, s>>>(x, h, rows, gamma, beta, eps);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_embed_layernorm(const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens, int row_base,
                           int n_seqs, int L, int vocab, int n_types, const float* word, const float* pos,
                           const float* type_emb, const float* gamma, const float* beta, float eps, float* x,
                           __nv_bfloat16* h, cudaStream_t s) {
    if (n_seqs == 0 || L == 0) return 0;
    embed_layernorm_kernel<<<grid_for((long long)n_seqs * L, 8), 256, 0, s>>>(ids, type_ids, cu_seqlens, row_base, n_seqs,
                                                                              L, vocab, n_types, word, pos, type_emb,
                                                                              gamma, beta, eps, x, h);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_pooler_classifier(const float* x, const int32_t* cu_seqlens, int row_base, int n_seqs, const float* wp,
                             const float* bp, const float* wc, const float* bc, int labels, float* logits,
                             cudaStream_t s) {
    if (n_seqs == 0) return 0;
    static SmemAttrCache smem_attr;
    if (int rc = smem_attr.ensure(pooler_classifier_kernel, POOLER_SMEM)) return rc;
    pooler_classifier_kernel<<<(n_seqs + POOL_SEQS - 1) / POOL_SEQS, 256, POOLER_SMEM, s>>>(x, cu_seqlens, row_base, n_seqs,
                                                                                             wp, bp, wc, bc, labels, logits);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
