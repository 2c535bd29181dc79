// Shared device/host helpers for the sasvqa_b200 kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace sasvqa {

// ---- model constants: ViT-B/16 @224 (HF GitVisionConfig defaults; the reference's encoder,
// src/preprocessing/extract_features.py:145)
constexpr int kImg = 224;
constexpr int kPatch = 16;
constexpr int kGrid = kImg / kPatch;          // 14
constexpr int kPatches = kGrid * kGrid;       // 196
constexpr int kTokens = kPatches + 1;         // 197
constexpr int kHidden = 768;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kFfn = 3072;
constexpr int kLayers = 12;
constexpr int kQkv = 3 * kHidden;             // 2304
constexpr float kLnEps = 1e-5f;
constexpr int kFrameElems = 3 * kImg * kImg;  // 150528

// ---- error plumbing (no exceptions across the C ABI)
void set_last_error(const std::string& msg);
void count_launch(int n = 1);   // every kernel launch of this library is counted (bench.py "gpu_launches")

#define SASVQA_CUDA_CHECK(expr)                                                              \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::sasvqa::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) +    \
                                     " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
            return 2; /* SASVQA_ERR_CUDA */                                                  \
        }                                                                                    \
    } while (0)

#define SASVQA_REQUIRE(cond, msg)                                                            \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::sasvqa::set_last_error(std::string(msg) + " [" #cond "]");                     \
            return 1; /* SASVQA_ERR_INVALID */                                               \
        }                                                                                    \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember the largest size set on each one
// (one process per GPU is the normal deployment, but nothing here may silently depend on it)
struct SmemAttrCache {
    size_t set[64] = {};
    template <class Kernel>
    int ensure(Kernel kernel, size_t bytes) {
        int dev = 0;
        SASVQA_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || bytes > set[dev]) {
            SASVQA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            if (dev >= 0 && dev < 64) set[dev] = bytes;
        }
        return 0;
    }
};

// A handle owns ONE workspace, but its entry points may be called on different streams (the caller's stream for the
// device-pointer entries, the handle's own copy / compute streams for the host-buffer pipelines).  Every entry opens a
// StreamOrder scope: its stream first waits for the last work any earlier entry queued on the workspace, and the scope's
// end marks the new tail -- so "a handle serialises its own work" (include/sasvqa.h) holds across streams too.
struct WorkspaceOrder {
    cudaEvent_t tail = nullptr;     // created with the handle (cudaEventDisableTiming)
    bool armed = false;
};
struct StreamOrder {
    WorkspaceOrder* o;
    cudaStream_t s;
    StreamOrder(WorkspaceOrder* o_, cudaStream_t s_) : o(o_), s(s_) {
        if (o && o->tail && o->armed) cudaStreamWaitEvent(s, o->tail, 0);
    }
    ~StreamOrder() {
        if (o && o->tail && cudaEventRecord(o->tail, s) == cudaSuccess) o->armed = true;
    }
    StreamOrder(const StreamOrder&) = delete;
    StreamOrder& operator=(const StreamOrder&) = delete;
};

// ---- epilogue modes of the encoder GEMM
enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,        // out_bf16 = acc + bias                         (fused q|k|v projection)
    EPI_BIAS_GELU_BF16 = 1,   // out_bf16 = quick_gelu(acc + bias)             (fc1)
    EPI_BIAS_RESID_F32 = 2,   // x_f32   += acc + bias  (in place)             (out_proj, fc2)
    EPI_PATCH_EMBED_F32 = 3,  // x_f32[frame*197 + 1 + p] = acc + pos[1 + p]   (patch embedding)
    EPI_BIAS_ERF_GELU_BF16 = 4,  // out_bf16 = gelu(acc + bias), exact erf form (BERT intermediate layer of the MIF scorer)
};

struct GemmArgs {
    const __nv_bfloat16* A;   // [M, K] row-major (K contiguous)
    const __nv_bfloat16* B;   // [N, K] row-major (nn.Linear weight layout)
    int M, N, K;
    int epilogue;
    const float* bias;        // [N]            (modes 0,1,2)
    const float* pos;         // [197, 768]     (mode 3)
    __nv_bfloat16* out_bf16;  // [M, N]         (modes 0,1)
    float* out_f32;           // [M, N] or [frames*197, N] (modes 2,3)
};

// launchers (each returns 0 or an error code, async on `stream`)
int launch_gemm_tcgen05(const GemmArgs& g, const CUtensorMap* map_a, const CUtensorMap* map_b,
                        const CUtensorMap* map_out, int num_sms, cudaStream_t stream);
int make_tensor_map_out(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, int is_f32);
int make_tensor_map_bf16_kmajor(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

int launch_resize_crop_u8(const uint8_t* frames_hwc, long long n_src, int H, int W, const int32_t* frame_map,
                          long long src_frame0, int n_out, uint8_t* out_224, cudaStream_t s);
long long resize_source_bytes_per_frame(int H, int W);
int launch_preprocess_u8(const uint8_t* frames_hwc, int n_frames, __nv_bfloat16* patches, cudaStream_t s);
int launch_patchify_f32(const float* frames_chw, int n_frames, __nv_bfloat16* patches, cudaStream_t s);
int launch_pre_layernorm(float* x, int n_frames, const float* cls_pos0, const float* gamma, const float* beta,
                         cudaStream_t s);
int launch_layernorm_bf16(const float* x, __nv_bfloat16* h, int rows, const float* gamma, const float* beta,
                          cudaStream_t s);
int launch_layernorm_f32(const float* x, float* out, long long rows, const float* gamma, const float* beta,
                         cudaStream_t s);
int launch_pool_norm(const float* x, int n_frames, const float* gamma, const float* beta, float* feats,
                     cudaStream_t s);
int make_attention_maps(CUtensorMap* map_q, CUtensorMap* map_kv, CUtensorMap* map_out, const void* qkv, const void* out,
                        uint64_t rows);
int launch_attention_tcgen05(const CUtensorMap* map_q, const CUtensorMap* map_kv, const CUtensorMap* map_out,
                             __nv_bfloat16* out, int n_frames, int num_sms, cudaStream_t s, int variant = 0);
// clip_off (device, [B + 1]) switches the selection stages to RAGGED batches: clip b = rows [clip_off[b], clip_off[b+1]) of the
// packed per-frame arrays, T = the longest clip, W may be -1 (adaptive per clip)
int launch_mdf_scores(const float* feats, int B, int T, int W, float* lcl_avg, float* gram, cudaStream_t s,
                      const int32_t* clip_off = nullptr);
int launch_mdf_select(const float* lcl_avg, int B, int T, int K, int W, int32_t* idx, int32_t* status,
                      cudaStream_t s, const int32_t* clip_off = nullptr);
int launch_mif_scores(const float* feats, const float* q, int B, int T, float* scores, cudaStream_t s);
int launch_topk_strided(const float* scores, int B, int T, int ds_rate, int K, int32_t* idx,
                        const int32_t* only_if_status, cudaStream_t s, const int32_t* clip_off = nullptr);
int launch_gather_u8(const uint8_t* clips, const int32_t* idx, int B, int T, int K, float* out, cudaStream_t s,
                     const int32_t* clip_off = nullptr);
int launch_gather_f32(const float* frames, const int32_t* idx, int B, int T, int K, int64_t row_elems, float* out,
                      cudaStream_t s, const int32_t* clip_off = nullptr);

// MIF cross-encoder (BERT sequence classifier, src/preprocessing/gen_sample.py:79-83): packed variable-length rows
int launch_attention_varlen(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int32_t* cu_seqlens_dev, int row_base,
                            int n_seqs, int max_len, cudaStream_t s);
int launch_attention_git(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_samples, int n_vis, int L, int text_only,
                         cudaStream_t s, const __nv_bfloat16* vis_kv = nullptr);
// the same attention on tcgen05 (attention_git_tcgen05.cu): visual query tiles (include_visual) + one text tile per 128 positions
int launch_attention_git_tcgen05(const __nv_bfloat16* qkv, __nv_bfloat16* out, long long rows_total, int n_samples, int n_vis,
                                 int L, int include_visual, int num_sms, cudaStream_t s);
int launch_layernorm_post(float* x, __nv_bfloat16* h, long long rows, const float* gamma, const float* beta, float eps,
                          cudaStream_t s);
int launch_embed_layernorm(const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens, int row_base,
                           int n_seqs, int L, int vocab, int n_types, const float* word, const float* pos,
                           const float* type_emb, const float* gamma, const float* beta, float eps, float* x,
                           __nv_bfloat16* h, cudaStream_t s);
int launch_pooler_classifier(const float* x, const int32_t* cu_seqlens, int row_base, int n_seqs, const float* wp,
                             const float* bp, const float* wc, const float* bc, int labels, float* logits,
                             cudaStream_t s);

// ---- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// x * sigmoid(1.702 x)  (HF "quick_gelu", modeling_git.py GitVisionMLP)
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
// 0.5 x (1 + erf(x / sqrt 2))  (HF "gelu", the BERT activation; modeling_bert.py BertIntermediate)
__device__ __forceinline__ float erf_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// CLIPImageProcessor arithmetic for one channel value, rounded like torch on the CPU
// (separate multiply, subtract, divide -- no FMA contraction): (u * (1/255) - mean) / std
__device__ __forceinline__ float normalize_px(uint32_t u, float mean, float stdv) {
    return __fdiv_rn(__fsub_rn(__fmul_rn((float)u, 1.0f / 255.0f), mean), stdv);
}

__device__ __forceinline__ float px_mean(int c) { return c == 0 ? 0.48145466f : (c == 1 ? 0.4578275f : 0.40821073f); }
__device__ __forceinline__ float px_std(int c) { return c == 0 ? 0.26862954f : (c == 1 ? 0.26130258f : 0.27577711f); }

#endif  // __CUDACC__

}  // namespace sasvqa
