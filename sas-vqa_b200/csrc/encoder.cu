// Encoder handle (bf16 weights + workspace + TMA descriptors) and the per-clip-list pipelines.
// Replaces the model side of the reference's extraction loop:
//   model.eval().cuda(); DataParallel(...)            src/preprocessing/extract_features.py:45-48
//   for chunk: model(frames[chunk]) -> mean -> norm   src/preprocessing/datautils/utils.py:39-48
// One process per GPU; the handle is bound to the device current at creation.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/sasvqa.h"
#include "common.cuh"

namespace sasvqa {

namespace {

constexpr int kDefaultChunkFrames = 2048;     // measured best: smaller chunks are slower at every stage

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
__global__ void fill_i32_kernel(int32_t* p, int32_t v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// selected (clip, index) pairs -> flat source frame numbers for the resize-on-gather path; picks < 0 stay < 0
__global__ void frame_map_kernel(const int32_t* idx, int B, int T, int K, int32_t* map, int32_t* unit_idx,
                                 const int32_t* clip_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int t = idx[i], b = i / K;
    const int f0 = clip_off ? clip_off[b] : b * T, Tb = clip_off ? clip_off[b + 1] - clip_off[b] : T;
    const bool ok = t >= 0 && t < Tb;
    map[i] = ok ? f0 + t : -1;
    unit_idx[i] = ok ? 0 : -1;
}

// profiling scopes (bench.py roofline): CUDA-event pairs around each stage's launches
enum ProfKind {
    PK_PREPROCESS = 0, PK_GEMM_PATCH, PK_PRE_LN, PK_LN, PK_GEMM_QKV, PK_ATTENTION, PK_GEMM_OUT, PK_GEMM_FC1,
    PK_GEMM_FC2, PK_POOL, PK_SCORES, PK_SELECT, PK_GATHER, PK_RESIZE, PK_PROJECTION, PK_COUNT
};
struct ProfRec {
    int kind;
    cudaEvent_t a, b;
};

struct Layer {
    __nv_bfloat16 *w_qkv, *w_out, *w_fc1, *w_fc2;
    float *b_qkv, *b_out, *b_fc1, *b_fc2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    CUtensorMap m_qkv, m_out, m_fc1, m_fc2;
};

}  // namespace

}  // namespace sasvqa

using namespace sasvqa;

struct SasvqaEncoder {
    int device = 0;
    int num_sms = 148;
    int chunk_frames = 0;
    // weights (one arena each for bf16 matrices and fp32 vectors)
    __nv_bfloat16* arena_bf16 = nullptr;
    float* arena_f32 = nullptr;
    __nv_bfloat16* w_patch = nullptr;
    float *pos = nullptr, *cls_pos0 = nullptr, *pre_g = nullptr, *pre_b = nullptr, *post_g = nullptr, *post_b = nullptr;
    Layer L[kLayers];
    CUtensorMap m_patch_w;
    // optional visual projection of the downstream GIT model (Linear 768->768 + LayerNorm): row f2
    __nv_bfloat16* w_proj = nullptr;
    float* proj_vec = nullptr;        // bias | ln gamma | ln beta
    CUtensorMap m_proj;
    // workspace for one chunk
    float* x = nullptr;               // [chunk*197, 768]  fp32 residual stream
    __nv_bfloat16* h = nullptr;       // [chunk*197, 768]  LN output / attention output
    __nv_bfloat16* big = nullptr;     // [chunk*197, 3072] qkv (as [.,2304]) / fc1 output / patch matrix (as [chunk*196,768])
    CUtensorMap m_h, m_big_fc, m_big_patch;       // A-operand views
    CUtensorMap m_att_q, m_att_kv, m_att_out;     // attention operand views of big as [., 2304]; output view of h
    CUtensorMap m_out_qkv, m_out_fc1, m_out_x;    // TMA-store views of big ([.,2304] / [.,3072]) and x
    // scratch for the whole-path entry points (grown on demand)
    float* feats = nullptr;
    size_t feats_cap = 0;
    float* lcl = nullptr;
    size_t lcl_cap = 0;
    // non-224 input: resized+cropped uint8 frames of one chunk / of the K picks, and the pick -> frame map
    uint8_t* resized = nullptr;
    size_t resized_cap = 0;
    uint8_t* picked = nullptr;
    size_t picked_cap = 0;
    int32_t* pick_map = nullptr;
    size_t pick_map_cap = 0;
    int32_t* clip_off = nullptr;          // ragged batches: [B + 1] frame offsets of the clips
    size_t clip_off_cap = 0;
    // host-buffer pipeline
    cudaStream_t h2d_stream = nullptr, compute_stream = nullptr, d2h_stream = nullptr;
    uint8_t* stage[2] = {nullptr, nullptr};
    size_t stage_cap[2] = {0, 0};
    float* out_stage[2] = {nullptr, nullptr};
    size_t out_stage_cap[2] = {0, 0};
    int32_t* idx_stage[2] = {nullptr, nullptr};
    size_t idx_stage_cap[2] = {0, 0};
    float* q_stage[2] = {nullptr, nullptr};
    size_t q_stage_cap[2] = {0, 0};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    WorkspaceOrder order;             // serialises the entry points across the streams they are called on
    // optional per-stage timing
    bool profile = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
};

namespace sasvqa {

namespace {

struct Scope {
    SasvqaEncoder* e;
    int kind;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(SasvqaEncoder* e) {
        if (!e->ev_pool.empty()) {
            cudaEvent_t ev = e->ev_pool.back();
            e->ev_pool.pop_back();
            return ev;
        }
        cudaEvent_t ev = nullptr;
        cudaEventCreate(&ev);
        return ev;
    }
    Scope(SasvqaEncoder* e_, int kind_, cudaStream_t s_) : e(e_), kind(kind_), s(s_) {
        if (e->profile) {
            a = take(e);
            b = take(e);
            cudaEventRecord(a, s);
        }
    }
    ~Scope() {
        if (e->profile) {
            cudaEventRecord(b, s);
            e->prof.push_back({kind, a, b});
        }
    }
};

int gemm(SasvqaEncoder* e, int kind, const GemmArgs& g, const CUtensorMap* ma, const CUtensorMap* mb,
         const CUtensorMap* mo, cudaStream_t s) {
    Scope sc(e, kind, s);
    return launch_gemm_tcgen05(g, ma, mb, mo, e->num_sms, s);
}

// encode `n` frames whose bf16 patch matrix is `patches` (map_patches describes it); leaves the
// hidden state after `n_layers` blocks in e->x
int encode_chunk(SasvqaEncoder* e, const __nv_bfloat16* patches, const CUtensorMap* map_patches, int n, int n_layers,
                 cudaStream_t s) {
    const int M = n * kTokens;
    int rc;
    GemmArgs g{};
    g.A = patches; g.B = e->w_patch; g.M = n * kPatches; g.N = kHidden; g.K = kHidden;
    g.epilogue = EPI_PATCH_EMBED_F32; g.pos = e->pos; g.out_f32 = e->x;
    if ((rc = gemm(e, PK_GEMM_PATCH, g, map_patches, &e->m_patch_w, nullptr, s))) return rc;
    {
        Scope sc(e, PK_PRE_LN, s);
        if ((rc = launch_pre_layernorm(e->x, n, e->cls_pos0, e->pre_g, e->pre_b, s))) return rc;
    }
    for (int l = 0; l < n_layers; ++l) {
        Layer& L = e->L[l];
        {
            Scope sc(e, PK_LN, s);
            if ((rc = launch_layernorm_bf16(e->x, e->h, M, L.ln1_g, L.ln1_b, s))) return rc;
        }
        g = GemmArgs{};
        g.A = e->h; g.B = L.w_qkv; g.M = M; g.N = kQkv; g.K = kHidden;
        g.epilogue = EPI_BIAS_BF16; g.bias = L.b_qkv; g.out_bf16 = e->big;
        if ((rc = gemm(e, PK_GEMM_QKV, g, &e->m_h, &L.m_qkv, &e->m_out_qkv, s))) return rc;
        {
            Scope sc(e, PK_ATTENTION, s);
            rc = launch_attention_tcgen05(&e->m_att_q, &e->m_att_kv, &e->m_att_out, e->h, n, e->num_sms, s);
            if (rc) return rc;
        }
        g = GemmArgs{};
        g.A = e->h; g.B = L.w_out; g.M = M; g.N = kHidden; g.K = kHidden;
        g.epilogue = EPI_BIAS_RESID_F32; g.bias = L.b_out; g.out_f32 = e->x;
        if ((rc = gemm(e, PK_GEMM_OUT, g, &e->m_h, &L.m_out, &e->m_out_x, s))) return rc;
        {
            Scope sc(e, PK_LN, s);
            if ((rc = launch_layernorm_bf16(e->x, e->h, M, L.ln2_g, L.ln2_b, s))) return rc;
        }
        g = GemmArgs{};
        g.A = e->h; g.B = L.w_fc1; g.M = M; g.N = kFfn; g.K = kHidden;
        g.epilogue = EPI_BIAS_GELU_BF16; g.bias = L.b_fc1; g.out_bf16 = e->big;
        if ((rc = gemm(e, PK_GEMM_FC1, g, &e->m_h, &L.m_fc1, &e->m_out_fc1, s))) return rc;
        g = GemmArgs{};
        g.A = e->big; g.B = L.w_fc2; g.M = M; g.N = kHidden; g.K = kFfn;
        g.epilogue = EPI_BIAS_RESID_F32; g.bias = L.b_fc2; g.out_f32 = e->x;
        if ((rc = gemm(e, PK_GEMM_FC2, g, &e->m_big_fc, &L.m_fc2, &e->m_out_x, s))) return rc;
    }
    return 0;
}

int grow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) SASVQA_CUDA_CHECK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SASVQA_CUDA_CHECK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

}  // namespace

int encoder_create(const float* params_host, uint64_t n_params, int chunk_frames, SasvqaEncoder** out) {
    SASVQA_REQUIRE(out != nullptr && params_host != nullptr, "null argument");
    SASVQA_REQUIRE(n_params == SASVQA_NUM_ENCODER_PARAMS, "state dict must hold 85 799 424 fp32 values (ViT-B/16)");
    if (chunk_frames <= 0) {
        // auto-size: the measured-best chunk, or as many frames as half of the free HBM holds (2.1 MB of workspace per frame:
        // fp32 stream + bf16 operand + the 3072-wide bf16 buffer), in steps of 64 frames
        chunk_frames = kDefaultChunkFrames;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t per_frame = (size_t)kTokens * (kHidden * sizeof(float) + kHidden * 2 + kFfn * 2);
            const size_t fit = free_b / 2 / per_frame;
            if (fit < (size_t)chunk_frames) chunk_frames = (int)std::max<size_t>(64, fit / 64 * 64);
        }
    }
    SasvqaEncoder* e = new SasvqaEncoder();
    auto fail = [&](int rc) { sasvqa_encoder_destroy(e); return rc; };
#define TRY(expr) do { int _rc = (expr); if (_rc) return fail(_rc); } while (0)
#define TRYCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return fail(SASVQA_ERR_CUDA); } } while (0)
    TRYCUDA(cudaGetDevice(&e->device));
    cudaDeviceProp prop;
    TRYCUDA(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major != 10) {
        set_last_error("sasvqa_b200 needs an sm_100a GPU (B200); found compute capability " +
                       std::to_string(prop.major) + "." + std::to_string(prop.minor));
        return fail(SASVQA_ERR_INVALID);
    }
    e->num_sms = prop.multiProcessorCount;
    e->chunk_frames = chunk_frames;

    // ---- upload the fp32 state dict once, carve bf16 matrices / fp32 vectors out of it on device
    float* raw = nullptr;
    TRYCUDA(cudaMalloc(&raw, n_params * sizeof(float)));
    if (cudaMemcpy(raw, params_host, n_params * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("uploading encoder parameters failed");
        return fail(SASVQA_ERR_CUDA);
    }
    const size_t n_mat = (size_t)kHidden * kHidden + (size_t)kLayers * ((size_t)kQkv * kHidden + (size_t)kHidden * kHidden +
                                                                          2 * (size_t)kFfn * kHidden);
    const size_t n_vec = (size_t)kTokens * kHidden + kHidden + 4 * kHidden +
                         (size_t)kLayers * (kQkv + kHidden + kFfn + kHidden + 4 * kHidden);
    if (cudaMalloc(&e->arena_bf16, n_mat * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&e->arena_f32, n_vec * sizeof(float)) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("allocating encoder weights failed");
        return fail(SASVQA_ERR_NOMEM);
    }
    __nv_bfloat16* mp = e->arena_bf16;
    float* vp = e->arena_f32;
    const float* rp = raw;
    auto take_mat = [&](size_t n) {      // fp32 -> bf16 matrix
        __nv_bfloat16* dst = mp;
        f32_to_bf16_kernel<<<592, 256>>>(rp, dst, (long long)n);
        mp += n; rp += n;
        return dst;
    };
    auto take_vec = [&](size_t n) {      // fp32 vector copy
        float* dst = vp;
        cudaMemcpyAsync(dst, rp, n * sizeof(float), cudaMemcpyDeviceToDevice, 0);
        vp += n; rp += n;
        return dst;
    };
    const float* cls_raw = rp; rp += kHidden;                       // class_embedding
    e->w_patch = take_mat((size_t)kHidden * kHidden);               // patch_embedding.weight [768, 3*16*16]
    e->pos = take_vec((size_t)kTokens * kHidden);                   // position_embedding.weight
    e->cls_pos0 = vp; vp += kHidden;
    add_vec_kernel<<<3, 256>>>(cls_raw, e->pos, e->cls_pos0, kHidden);
    e->pre_g = take_vec(kHidden);
    e->pre_b = take_vec(kHidden);
    for (int l = 0; l < kLayers; ++l) {
        Layer& L = e->L[l];
        // HF order: k, v, q, out.  Fused projection rows are ordered q | k | v.
        L.w_qkv = mp; mp += (size_t)kQkv * kHidden;
        L.b_qkv = vp; vp += kQkv;
        const size_t ww = (size_t)kHidden * kHidden;
        const int slot_of[3] = {1, 2, 0};   // k -> slot 1, v -> slot 2, q -> slot 0
        for (int j = 0; j < 3; ++j) {
            f32_to_bf16_kernel<<<592, 256>>>(rp, L.w_qkv + slot_of[j] * ww, (long long)ww);
            rp += ww;
            cudaMemcpyAsync(L.b_qkv + slot_of[j] * kHidden, rp, kHidden * sizeof(float), cudaMemcpyDeviceToDevice, 0);
            rp += kHidden;
        }
        L.w_out = take_mat(ww);
        L.b_out = take_vec(kHidden);
        L.ln1_g = take_vec(kHidden);
        L.ln1_b = take_vec(kHidden);
        L.w_fc1 = take_mat((size_t)kFfn * kHidden);
        L.b_fc1 = take_vec(kFfn);
        L.w_fc2 = take_mat((size_t)kHidden * kFfn);
        L.b_fc2 = take_vec(kHidden);
        L.ln2_g = take_vec(kHidden);
        L.ln2_b = take_vec(kHidden);
    }
    e->post_g = take_vec(kHidden);
    e->post_b = take_vec(kHidden);
    cudaError_t ce = cudaDeviceSynchronize();
    cudaFree(raw);
    if (ce != cudaSuccess || (size_t)(rp - raw) != n_params || (size_t)(mp - e->arena_bf16) != n_mat ||
        (size_t)(vp - e->arena_f32) != n_vec) {
        set_last_error(std::string("encoder weight conversion failed: ") + cudaGetErrorString(ce));
        return fail(SASVQA_ERR_CUDA);
    }

    // ---- workspace + TMA descriptors
    const size_t rows = (size_t)chunk_frames * kTokens;
    if (cudaMalloc(&e->x, rows * kHidden * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&e->h, rows * kHidden * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&e->big, rows * kFfn * sizeof(__nv_bfloat16)) != cudaSuccess) {
        set_last_error("allocating encoder workspace failed (lower chunk_frames)");
        return fail(SASVQA_ERR_NOMEM);
    }
    TRYCUDA(cudaMemset(e->h, 0, rows * kHidden * sizeof(__nv_bfloat16)));
    TRYCUDA(cudaMemset(e->big, 0, rows * kFfn * sizeof(__nv_bfloat16)));
    TRY(make_tensor_map_bf16_kmajor(&e->m_patch_w, e->w_patch, kHidden, kHidden, 128));
    TRY(make_attention_maps(&e->m_att_q, &e->m_att_kv, &e->m_att_out, e->big, e->h, rows));
    TRY(make_tensor_map_out(&e->m_out_qkv, e->big, rows, kQkv, 0));
    TRY(make_tensor_map_out(&e->m_out_fc1, e->big, rows, kFfn, 0));
    TRY(make_tensor_map_out(&e->m_out_x, e->x, rows, kHidden, 1));
    TRY(make_tensor_map_bf16_kmajor(&e->m_h, e->h, rows, kHidden, 128));
    TRY(make_tensor_map_bf16_kmajor(&e->m_big_fc, e->big, rows, kFfn, 128));
    TRY(make_tensor_map_bf16_kmajor(&e->m_big_patch, e->big, (uint64_t)chunk_frames * kPatches, kHidden, 128));
    for (int l = 0; l < kLayers; ++l) {
        Layer& L = e->L[l];
        TRY(make_tensor_map_bf16_kmajor(&L.m_qkv, L.w_qkv, kQkv, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&L.m_out, L.w_out, kHidden, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&L.m_fc1, L.w_fc1, kFfn, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&L.m_fc2, L.w_fc2, kHidden, kFfn, 128));
    }
    TRYCUDA(cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&e->compute_stream, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    TRYCUDA(cudaEventCreateWithFlags(&e->order.tail, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        TRYCUDA(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
        TRYCUDA(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming));
        TRYCUDA(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
    }
#undef TRY
#undef TRYCUDA
    *out = e;
    return 0;
}

int encoder_chunk_frames(const SasvqaEncoder* e) { return e ? e->chunk_frames : 0; }

int profile_enable(SasvqaEncoder* e, int on) {
    SASVQA_REQUIRE(e != nullptr, "null encoder");
    e->profile = on != 0;
    return 0;
}

// Sums the recorded stage timings (ms) and launch-scope counts per ProfKind, then resets.
int profile_read(SasvqaEncoder* e, double* ms, int64_t* scopes, int n_kinds) {
    SASVQA_REQUIRE(e != nullptr && ms != nullptr && scopes != nullptr && n_kinds >= PK_COUNT, "bad arguments");
    SASVQA_CUDA_CHECK(cudaDeviceSynchronize());
    for (int k = 0; k < n_kinds; ++k) {
        ms[k] = 0.0;
        scopes[k] = 0;
    }
    for (const ProfRec& r : e->prof) {
        float t = 0.f;
        SASVQA_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.kind] += t;
        scopes[r.kind] += 1;
        e->ev_pool.push_back(r.a);
        e->ev_pool.push_back(r.b);
    }
    e->prof.clear();
    return 0;
}

void encoder_destroy(SasvqaEncoder* e) {
    if (!e) return;
    cudaFree(e->arena_bf16); cudaFree(e->arena_f32);
    cudaFree(e->x); cudaFree(e->h); cudaFree(e->big);
    cudaFree(e->w_proj); cudaFree(e->proj_vec);
    cudaFree(e->feats); cudaFree(e->lcl);
    cudaFree(e->resized); cudaFree(e->picked); cudaFree(e->pick_map); cudaFree(e->clip_off);
    for (int i = 0; i < 2; ++i) {
        cudaFree(e->stage[i]); cudaFree(e->out_stage[i]); cudaFree(e->idx_stage[i]); cudaFree(e->q_stage[i]);
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    if (e->order.tail) cudaEventDestroy(e->order.tail);
    for (const ProfRec& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->h2d_stream) cudaStreamDestroy(e->h2d_stream);
    if (e->compute_stream) cudaStreamDestroy(e->compute_stream);
    if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
    delete e;
}

// patches (caller memory) -> feats, any n
int encoder_fwd(SasvqaEncoder* e, const __nv_bfloat16* patches, int n_frames, float* feats, cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && n_frames >= 0, "bad arguments");
    StreamOrder order(&e->order, s);
    for (int f0 = 0; f0 < n_frames; f0 += e->chunk_frames) {
        const int n = std::min(e->chunk_frames, n_frames - f0);
        const __nv_bfloat16* p = patches + (size_t)f0 * kPatches * kHidden;
        CUtensorMap mp;
        int rc = make_tensor_map_bf16_kmajor(&mp, p, (uint64_t)n * kPatches, kHidden, 128);
        if (rc) return rc;
        if ((rc = encode_chunk(e, p, &mp, n, kLayers, s))) return rc;
        {
            Scope sc(e, PK_POOL, s);
            if ((rc = launch_pool_norm(e->x, n, e->post_g, e->post_b, feats + (size_t)f0 * kHidden, s))) return rc;
        }
    }
    return 0;
}

int encoder_fwd_hidden(SasvqaEncoder* e, const __nv_bfloat16* patches, int n_frames, int n_layers, float* hidden,
                       cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && n_frames > 0 && n_frames <= e->chunk_frames, "n_frames must be in (0, chunk_frames]");
    SASVQA_REQUIRE(n_layers >= 0 && n_layers <= kLayers, "n_layers out of range");
    StreamOrder order(&e->order, s);
    CUtensorMap mp;
    int rc = make_tensor_map_bf16_kmajor(&mp, patches, (uint64_t)n_frames * kPatches, kHidden, 128);
    if (rc) return rc;
    if ((rc = encode_chunk(e, patches, &mp, n_frames, n_layers, s))) return rc;
    SASVQA_CUDA_CHECK(cudaMemcpyAsync(hidden, e->x, (size_t)n_frames * kTokens * kHidden * sizeof(float),
                                      cudaMemcpyDeviceToDevice, s));
    return 0;
}

// frames (uint8 HWC or fp32 CHW, device) -> feats [n_frames, 768], chunked through the workspace
static int encode_frames(SasvqaEncoder* e, const uint8_t* u8, const float* f32, long long n_frames, int H, int W,
                         float* feats, cudaStream_t s) {
    const bool resize = u8 != nullptr && (H != kImg || W != kImg);
    int rc;
    if (resize && (rc = grow((void**)&e->resized, &e->resized_cap, (size_t)e->chunk_frames * kFrameElems))) return rc;
    for (long long f0 = 0; f0 < n_frames; f0 += e->chunk_frames) {
        const int n = (int)std::min<long long>(e->chunk_frames, n_frames - f0);
        if (resize) {                                            // K0: shortest-edge bicubic resize + centre crop
            Scope sc(e, PK_RESIZE, s);
            if ((rc = launch_resize_crop_u8(u8, n_frames, H, W, nullptr, f0, n, e->resized, s))) return rc;
        }
        {
            Scope sc(e, PK_PREPROCESS, s);
            if (resize) rc = launch_preprocess_u8(e->resized, n, e->big, s);
            else if (u8) rc = launch_preprocess_u8(u8 + (size_t)f0 * kFrameElems, n, e->big, s);
            else rc = launch_patchify_f32(f32 + (size_t)f0 * kFrameElems, n, e->big, s);
            if (rc) return rc;
        }
        if ((rc = encode_chunk(e, e->big, &e->m_big_patch, n, kLayers, s))) return rc;
        {
            Scope sc(e, PK_POOL, s);
            if ((rc = launch_pool_norm(e->x, n, e->post_g, e->post_b, feats + (size_t)f0 * kHidden, s))) return rc;
        }
    }
    return 0;
}

// ---- row f2: the visual side of the downstream video-QA forward on the sampled frames -------------------
// src/modeling/modeling.py:76-95 (MyGitModel.forward): per frame `image_encoder(frame).last_hidden_state`
// (ALL 197 tokens after post_layernorm), concatenated along the sequence, then `visual_projection`
// (HF GitProjection: Linear(768, 768) + LayerNorm, eps 1e-5).  The image encoder IS the sampler's encoder, so
// the K sampled frames of every clip run through the same kernels in one batch instead of a Python loop.
int encoder_set_projection(SasvqaEncoder* e, const float* w_host, const float* b_host, const float* g_host,
                           const float* beta_host) {
    SASVQA_REQUIRE(e && w_host && b_host && g_host && beta_host, "null argument");
    const size_t nw = (size_t)kHidden * kHidden;
    float* raw = nullptr;
    SASVQA_CUDA_CHECK(cudaMalloc(&raw, nw * sizeof(float)));
    if (!e->w_proj && cudaMalloc(&e->w_proj, nw * sizeof(__nv_bfloat16)) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("allocating projection weights failed");
        return SASVQA_ERR_NOMEM;
    }
    if (!e->proj_vec && cudaMalloc(&e->proj_vec, 3 * kHidden * sizeof(float)) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("allocating projection vectors failed");
        return SASVQA_ERR_NOMEM;
    }
    cudaError_t ce = cudaMemcpy(raw, w_host, nw * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        f32_to_bf16_kernel<<<592, 256>>>(raw, e->w_proj, (long long)nw);
        ce = cudaMemcpy(e->proj_vec, b_host, kHidden * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(e->proj_vec + kHidden, g_host, kHidden * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->proj_vec + 2 * kHidden, beta_host, kHidden * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    cudaFree(raw);
    if (ce != cudaSuccess) {
        set_last_error(std::string("uploading projection weights failed: ") + cudaGetErrorString(ce));
        return SASVQA_ERR_CUDA;
    }
    return make_tensor_map_bf16_kmajor(&e->m_proj, e->w_proj, kHidden, kHidden, 128);
}

// frames (uint8 HWC 224x224 or normalised fp32 CHW -- the rows of the "sampled_frames" dataset) ->
// tokens [n_frames * 197, 768] fp32: last_hidden_state (project == 0) or visual_projection(last_hidden_state)
int visual_tokens(SasvqaEncoder* e, const uint8_t* u8, const float* f32, int n_frames, int project, float* tokens,
                  cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && n_frames >= 0, "bad arguments");
    SASVQA_REQUIRE(n_frames == 0 || ((u8 != nullptr || f32 != nullptr) && tokens != nullptr), "null argument");
    SASVQA_REQUIRE(!project || e->w_proj != nullptr, "no visual projection loaded (sasvqa_encoder_set_projection)");
    StreamOrder order(&e->order, s);
    for (int f0 = 0; f0 < n_frames; f0 += e->chunk_frames) {
        const int n = std::min(e->chunk_frames, n_frames - f0);
        const long long rows = (long long)n * kTokens;
        float* out = tokens + (size_t)f0 * kTokens * kHidden;
        int rc;
        {
            Scope sc(e, PK_PREPROCESS, s);
            if (u8) rc = launch_preprocess_u8(u8 + (size_t)f0 * kFrameElems, n, e->big, s);
            else rc = launch_patchify_f32(f32 + (size_t)f0 * kFrameElems, n, e->big, s);
            if (rc) return rc;
        }
        if ((rc = encode_chunk(e, e->big, &e->m_big_patch, n, kLayers, s))) return rc;
        if (!project) {                                          // post_layernorm of every token, fp32
            Scope sc(e, PK_PROJECTION, s);
            if ((rc = launch_layernorm_f32(e->x, out, rows, e->post_g, e->post_b, s))) return rc;
            continue;
        }
        {
            Scope sc(e, PK_LN, s);
            if ((rc = launch_layernorm_bf16(e->x, e->h, (int)rows, e->post_g, e->post_b, s))) return rc;
        }
        {   // Linear: the residual-mode epilogue accumulates acc + bias into a zeroed fp32 buffer
            Scope sc(e, PK_PROJECTION, s);
            SASVQA_CUDA_CHECK(cudaMemsetAsync(e->x, 0, (size_t)rows * kHidden * sizeof(float), s));
            GemmArgs g{};
            g.A = e->h; g.B = e->w_proj; g.M = (int)rows; g.N = kHidden; g.K = kHidden;
            g.epilogue = EPI_BIAS_RESID_F32; g.bias = e->proj_vec; g.out_f32 = e->x;
            rc = launch_gemm_tcgen05(g, &e->m_h, &e->m_proj, &e->m_out_x, e->num_sms, s);
            if (rc) return rc;
            if ((rc = launch_layernorm_f32(e->x, out, rows, e->proj_vec + kHidden, e->proj_vec + 2 * kHidden, s))) return rc;
        }
    }
    return 0;
}

// K5 for any frame size: the sampled rows are the image processor's output for the K picks of every clip
static int gather_picks(SasvqaEncoder* e, const uint8_t* u8, const float* f32, int B, int T, int H, int Wd, int K,
                        const int32_t* idx, float* sampled, cudaStream_t s, const int32_t* clip_off = nullptr,
                        long long n_frames_total = -1) {
    int rc = 0;
    const size_t nf = n_frames_total >= 0 ? (size_t)n_frames_total : (size_t)B * T;
    if (u8 && (H != kImg || Wd != kImg)) {                       // resize only the picks, then gather
        const size_t np = (size_t)B * K;
        if ((rc = grow((void**)&e->picked, &e->picked_cap, np * kFrameElems))) return rc;
        if ((rc = grow((void**)&e->pick_map, &e->pick_map_cap, 2 * np * sizeof(int32_t)))) return rc;
        int32_t* unit_idx = e->pick_map + np;
        {
            Scope sc(e, PK_RESIZE, s);
            frame_map_kernel<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(idx, B, T, K, e->pick_map, unit_idx, clip_off);
            SASVQA_CUDA_CHECK(cudaGetLastError());
            count_launch();
            if ((rc = launch_resize_crop_u8(u8, (long long)nf, H, Wd, e->pick_map, 0, (int)np, e->picked, s))) return rc;
        }
        Scope sc(e, PK_GATHER, s);
        return launch_gather_u8(e->picked, unit_idx, (int)np, 1, 1, sampled, s);
    }
    Scope sc(e, PK_GATHER, s);
    if (u8) rc = launch_gather_u8(u8, idx, B, T, K, sampled, s, clip_off);
    else rc = launch_gather_f32(f32, idx, B, T, K, kFrameElems, sampled, s, clip_off);
    return rc;
}

// Ragged batch: B clips of different lengths packed back to back, clip b = frames [off[b], off[b+1]) (the reference
// handles one video of any length per call, extract_features.py:80-97; real datasets are ragged).  The encoder is
// frame-batched already; the selection stages take the offsets.  Per clip exactly what the uniform path does for a
// clip of that length: T_b == 0 -> status EMPTY, zero frames; W == -1 -> T_b / 20; T_b < K on the fallback -> TOO_FEW.
// lcl_out / feats_out are packed per frame ([sum T], [sum T, 768]).
int mdf_sample_ragged_device(SasvqaEncoder* e, const uint8_t* u8, const float* f32, int B, const int32_t* off_host, int H,
                             int Wd, int K, int W, int32_t* idx, int32_t* status, float* lcl_out, float* feats_out,
                             float* sampled, cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && idx != nullptr && status != nullptr && off_host != nullptr, "null argument");
    SASVQA_REQUIRE(B >= 0 && K >= 1 && K <= 2048, "bad B/K");
    SASVQA_REQUIRE(H > 0 && Wd > 0, "bad frame size");
    SASVQA_REQUIRE(u8 != nullptr || f32 != nullptr || off_host[B] == 0, "null frames");
    SASVQA_REQUIRE(u8 != nullptr || (H == kImg && Wd == kImg), "fp32 frames are already processed: they must be 224x224");
    SASVQA_REQUIRE(W >= -1, "W must be >= 0, or -1 for the adaptive width T / 20");
    if (B == 0) return 0;
    SASVQA_REQUIRE(off_host[0] == 0, "clip offsets must start at 0");
    StreamOrder order(&e->order, s);
    int t_max = 0;
    for (int b = 0; b < B; ++b) {
        SASVQA_REQUIRE(off_host[b + 1] >= off_host[b], "clip offsets must not decrease");
        t_max = std::max(t_max, off_host[b + 1] - off_host[b]);
    }
    const long long nf = off_host[B];
    int rc;
    if ((rc = grow((void**)&e->clip_off, &e->clip_off_cap, ((size_t)B + 1) * sizeof(int32_t)))) return rc;
    SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->clip_off, off_host, ((size_t)B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (t_max == 0) {                                           // every clip empty
        fill_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(status, SASVQA_STATUS_EMPTY, B);
        fill_i32_kernel<<<(B * K + 255) / 256, 256, 0, s>>>(idx, -1, B * K);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        if (sampled) SASVQA_CUDA_CHECK(cudaMemsetAsync(sampled, 0, (size_t)B * K * kFrameElems * sizeof(float), s));
        return 0;
    }
    float* feats = feats_out;
    if (!feats) {
        if ((rc = grow((void**)&e->feats, &e->feats_cap, (size_t)nf * kHidden * sizeof(float)))) return rc;
        feats = e->feats;
    }
    float* lcl = lcl_out;
    if (!lcl) {
        if ((rc = grow((void**)&e->lcl, &e->lcl_cap, (size_t)nf * sizeof(float)))) return rc;
        lcl = e->lcl;
    }
    if ((rc = encode_frames(e, u8, f32, nf, H, Wd, feats, s))) return rc;
    {
        Scope sc(e, PK_SCORES, s);
        if ((rc = launch_mdf_scores(feats, B, t_max, W, lcl, nullptr, s, e->clip_off))) return rc;
    }
    {
        Scope sc(e, PK_SELECT, s);
        if ((rc = launch_mdf_select(lcl, B, t_max, K, W, idx, status, s, e->clip_off))) return rc;
    }
    if (sampled && (rc = gather_picks(e, u8, f32, B, t_max, H, Wd, K, idx, sampled, s, e->clip_off, nf))) return rc;
    return 0;
}

int mdf_sample_device(SasvqaEncoder* e, const uint8_t* u8, const float* f32, int B, int T, int H, int Wd, int K, int W,
                      int32_t* idx, int32_t* status, float* lcl_out, float* feats_out, float* sampled, cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && idx != nullptr && status != nullptr, "null argument");
    SASVQA_REQUIRE(B >= 0 && T >= 0 && K >= 1, "bad B/T/K");
    SASVQA_REQUIRE(H > 0 && Wd > 0, "bad frame size");
    SASVQA_REQUIRE(u8 != nullptr || (H == kImg && Wd == kImg), "fp32 frames are already processed: they must be 224x224");
    SASVQA_REQUIRE(W >= -1, "W must be >= 0, or -1 for the adaptive width T / 20");
    if (B == 0) return 0;
    StreamOrder order(&e->order, s);
    if (W == -1) W = T / 20;                                    // utils.py:32-33
    if (T == 0) {                                               // utils.py:50-52: zero frames, 'Zeros'
        fill_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(status, SASVQA_STATUS_EMPTY, B);
        fill_i32_kernel<<<(B * K + 255) / 256, 256, 0, s>>>(idx, -1, B * K);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        if (sampled) SASVQA_CUDA_CHECK(cudaMemsetAsync(sampled, 0, (size_t)B * K * kFrameElems * sizeof(float), s));
        return 0;
    }
    int rc;
    const size_t nf = (size_t)B * T;
    float* feats = feats_out;
    if (!feats) {
        if ((rc = grow((void**)&e->feats, &e->feats_cap, nf * kHidden * sizeof(float)))) return rc;
        feats = e->feats;
    }
    float* lcl = lcl_out;
    if (!lcl) {
        if ((rc = grow((void**)&e->lcl, &e->lcl_cap, nf * sizeof(float)))) return rc;
        lcl = e->lcl;
    }
    if ((rc = encode_frames(e, u8, f32, (long long)nf, H, Wd, feats, s))) return rc;
    {
        Scope sc(e, PK_SCORES, s);
        if ((rc = launch_mdf_scores(feats, B, T, W, lcl, nullptr, s))) return rc;
    }
    {
        Scope sc(e, PK_SELECT, s);
        if ((rc = launch_mdf_select(lcl, B, T, K, W, idx, status, s))) return rc;
    }
    if (sampled && (rc = gather_picks(e, u8, f32, B, T, H, Wd, K, idx, sampled, s))) return rc;
    return 0;
}

// MIF with embedding-space relevance (BASELINE config 3): encode -> <feat, q> -> strided top-K -> gather.
int mif_sample_device(SasvqaEncoder* e, const uint8_t* u8, const float* f32, int B, int T, int H, int Wd, const float* q,
                      int K, int ds_rate, int32_t* idx, float* scores_out, float* feats_out, float* sampled,
                      cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && idx != nullptr && q != nullptr, "null argument");
    SASVQA_REQUIRE(B >= 0 && T >= 1 && K >= 1 && ds_rate >= 1, "bad B/T/K/ds_rate");
    SASVQA_REQUIRE(H > 0 && Wd > 0, "bad frame size");
    SASVQA_REQUIRE(u8 != nullptr || (H == kImg && Wd == kImg), "fp32 frames are already processed: they must be 224x224");
    SASVQA_REQUIRE(K <= (T + ds_rate - 1) / ds_rate, "selected index k out of range");
    if (B == 0) return 0;
    StreamOrder order(&e->order, s);
    int rc;
    const size_t nf = (size_t)B * T;
    float* feats = feats_out;
    if (!feats) {
        if ((rc = grow((void**)&e->feats, &e->feats_cap, nf * kHidden * sizeof(float)))) return rc;
        feats = e->feats;
    }
    float* scores = scores_out;
    if (!scores) {
        if ((rc = grow((void**)&e->lcl, &e->lcl_cap, nf * sizeof(float)))) return rc;
        scores = e->lcl;
    }
    if ((rc = encode_frames(e, u8, f32, (long long)nf, H, Wd, feats, s))) return rc;
    {
        Scope sc(e, PK_SCORES, s);
        if ((rc = launch_mif_scores(feats, q, B, T, scores, s))) return rc;
    }
    {
        Scope sc(e, PK_SELECT, s);
        if ((rc = launch_topk_strided(scores, B, T, ds_rate, K, idx, nullptr, s))) return rc;
    }
    return sampled ? gather_picks(e, u8, f32, B, T, H, Wd, K, idx, sampled, s) : 0;
}

// waits for everything the host-buffer pipelines queued; returns the first CUDA error (reported, not thrown)
static int drain_pipeline(SasvqaEncoder* e) {
    const cudaError_t a = cudaStreamSynchronize(e->d2h_stream);
    const cudaError_t b = cudaStreamSynchronize(e->compute_stream);
    const cudaError_t c = cudaStreamSynchronize(e->h2d_stream);
    const cudaError_t bad = a != cudaSuccess ? a : (b != cudaSuccess ? b : c);
    if (bad != cudaSuccess) {
        set_last_error(std::string("host-buffer pipeline: ") + cudaGetErrorString(bad));
        return SASVQA_ERR_CUDA;
    }
    return 0;
}

// Host-buffer pipeline: groups of whole clips; H2D (h2d_stream), compute (compute_stream) and
// D2H (d2h_stream) of consecutive groups overlap through two staging slots.
// q_host != nullptr switches the groups to the MIF path (question embeddings [B, 768] fp32 travel with their clips,
// ds_rate is the stride of the top-K; W is unused and every status is 0).
int sample_host_pipeline(SasvqaEncoder* e, const uint8_t* clips, int B, int T, int H, int Wd, int K, int W,
                         const float* q_host, int ds_rate, int32_t* idx_host, int32_t* status_host, float* sampled_host) {
    const bool mif = q_host != nullptr;
    SASVQA_REQUIRE(e != nullptr && idx_host != nullptr && (mif || status_host != nullptr), "null argument");
    SASVQA_REQUIRE(B >= 0 && T >= 0 && K >= 1 && W >= -1, "bad B/T/K/W");
    SASVQA_REQUIRE(H > 0 && Wd > 0, "bad frame size");
    SASVQA_REQUIRE(!mif || (ds_rate >= 1 && T >= 1 && K <= (T + ds_rate - 1) / ds_rate), "selected index k out of range");
    const size_t frame_bytes = (size_t)H * Wd * 3;
    if (B == 0) return 0;
    if (T == 0) {
        for (int b = 0; b < B && status_host; ++b) status_host[b] = SASVQA_STATUS_EMPTY;
        for (long long i = 0; i < (long long)B * K; ++i) idx_host[i] = -1;
        if (sampled_host) memset(sampled_host, 0, (size_t)B * K * kFrameElems * sizeof(float));
        return 0;
    }
    SASVQA_REQUIRE(clips != nullptr, "null clips");
    const int group = std::max(1, e->chunk_frames / T);          // whole clips per group
    const size_t out_need = sampled_host ? (size_t)group * K * kFrameElems * sizeof(float) : 0;
    int rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = grow((void**)&e->stage[i], &e->stage_cap[i], (size_t)group * T * frame_bytes))) return rc;
        if (out_need && (rc = grow((void**)&e->out_stage[i], &e->out_stage_cap[i], out_need))) return rc;
        if ((rc = grow((void**)&e->idx_stage[i], &e->idx_stage_cap[i], (size_t)group * (K + 1) * sizeof(int32_t))))
            return rc;
        if (mif && (rc = grow((void**)&e->q_stage[i], &e->q_stage_cap[i], (size_t)group * kHidden * sizeof(float)))) return rc;
    }
    const int n_groups = (B + group - 1) / group;
    auto run_group = [&](int gi) -> int {
        const int slot = gi & 1;
        const int b0 = gi * group, nb = std::min(group, B - b0);
        // stage[slot] is free once group gi-2 has finished computing (its gather reads the frames)
        if (gi >= 2) SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->h2d_stream, e->ev_comp[slot], 0));
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->stage[slot], clips + (size_t)b0 * T * frame_bytes,
                                          (size_t)nb * T * frame_bytes, cudaMemcpyHostToDevice, e->h2d_stream));
        if (mif)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->q_stage[slot], q_host + (size_t)b0 * kHidden, (size_t)nb * kHidden * sizeof(float),
                                              cudaMemcpyHostToDevice, e->h2d_stream));
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_in[slot], e->h2d_stream));
        SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->compute_stream, e->ev_in[slot], 0));
        // idx/out staging of this slot is free once group gi-2's results are on the host
        if (gi >= 2) SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->compute_stream, e->ev_out[slot], 0));
        int32_t* d_idx = e->idx_stage[slot];
        int32_t* d_status = d_idx + (size_t)group * K;
        float* d_out = sampled_host ? e->out_stage[slot] : nullptr;
        const int rc2 = mif ? mif_sample_device(e, e->stage[slot], nullptr, nb, T, H, Wd, e->q_stage[slot], K, ds_rate, d_idx,
                                                nullptr, nullptr, d_out, e->compute_stream)
                            : mdf_sample_device(e, e->stage[slot], nullptr, nb, T, H, Wd, K, W, d_idx, d_status, nullptr, nullptr,
                                                d_out, e->compute_stream);
        if (rc2) return rc2;
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_comp[slot], e->compute_stream));
        SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->d2h_stream, e->ev_comp[slot], 0));
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(idx_host + (size_t)b0 * K, d_idx, (size_t)nb * K * sizeof(int32_t),
                                          cudaMemcpyDeviceToHost, e->d2h_stream));
        if (!mif)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(status_host + b0, d_status, (size_t)nb * sizeof(int32_t),
                                              cudaMemcpyDeviceToHost, e->d2h_stream));
        if (sampled_host)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(sampled_host + (size_t)b0 * K * kFrameElems, d_out,
                                              (size_t)nb * K * kFrameElems * sizeof(float), cudaMemcpyDeviceToHost,
                                              e->d2h_stream));
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_out[slot], e->d2h_stream));
        return 0;
    };
    rc = 0;
    for (int gi = 0; gi < n_groups && rc == 0; ++gi) rc = run_group(gi);
    const int rc_drain = drain_pipeline(e);     // also on failure: no copy to / from the caller's buffers stays in flight
    return rc ? rc : rc_drain;
}

int mdf_sample_host(SasvqaEncoder* e, const uint8_t* clips, int B, int T, int H, int Wd, int K, int W, int32_t* idx_host,
                    int32_t* status_host, float* sampled_host) {
    return sample_host_pipeline(e, clips, B, T, H, Wd, K, W, nullptr, 1, idx_host, status_host, sampled_host);
}

// MIF (embedding-space relevance, BASELINE config 3) from host buffers: clips [B, T, H, W, 3] uint8, q [B, 768] fp32
int mif_sample_host(SasvqaEncoder* e, const uint8_t* clips, int B, int T, int H, int Wd, const float* q_host, int K, int ds_rate,
                    int32_t* idx_host, float* sampled_host) {
    SASVQA_REQUIRE(q_host != nullptr, "null question embeddings");
    return sample_host_pipeline(e, clips, B, T, H, Wd, K, 0, q_host, ds_rate, idx_host, nullptr, sampled_host);
}

// Host-buffer pipeline for ragged batches: consecutive whole clips are grouped up to chunk_frames frames (a longer clip
// is a group of its own); H2D, compute and D2H of consecutive groups overlap through the same two staging slots as
// mdf_sample_host.  frames_host [sum T, H, W, 3] uint8, off_host [B + 1].
int mdf_sample_ragged_host(SasvqaEncoder* e, const uint8_t* frames, int B, const int32_t* off_host, int H, int Wd, int K, int W,
                           int32_t* idx_host, int32_t* status_host, float* sampled_host) {
    SASVQA_REQUIRE(e != nullptr && idx_host != nullptr && status_host != nullptr && off_host != nullptr, "null argument");
    SASVQA_REQUIRE(B >= 0 && K >= 1 && W >= -1, "bad B/K/W");
    SASVQA_REQUIRE(H > 0 && Wd > 0, "bad frame size");
    if (B == 0) return 0;
    SASVQA_REQUIRE(off_host[0] == 0, "clip offsets must start at 0");
    for (int b = 0; b < B; ++b) SASVQA_REQUIRE(off_host[b + 1] >= off_host[b], "clip offsets must not decrease");
    SASVQA_REQUIRE(frames != nullptr || off_host[B] == 0, "null frames");
    const size_t frame_bytes = (size_t)H * Wd * 3;
    // groups of whole clips: [g_begin[i], g_begin[i+1])
    std::vector<int> g_begin{0};
    long long max_frames = 1;
    int max_clips = 1;
    for (int b = 0; b < B;) {
        int j = b + 1;
        while (j < B && (long long)off_host[j + 1] - off_host[b] <= e->chunk_frames) ++j;
        max_frames = std::max<long long>(max_frames, off_host[j] - off_host[b]);
        max_clips = std::max(max_clips, j - b);
        g_begin.push_back(j);
        b = j;
    }
    const size_t out_need = sampled_host ? (size_t)max_clips * K * kFrameElems * sizeof(float) : 0;
    int rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = grow((void**)&e->stage[i], &e->stage_cap[i], (size_t)max_frames * frame_bytes))) return rc;
        if (out_need && (rc = grow((void**)&e->out_stage[i], &e->out_stage_cap[i], out_need))) return rc;
        if ((rc = grow((void**)&e->idx_stage[i], &e->idx_stage_cap[i], (size_t)max_clips * (K + 1) * sizeof(int32_t)))) return rc;
    }
    std::vector<int32_t> off_g;
    const int n_groups = (int)g_begin.size() - 1;
    auto run_group = [&](int gi) -> int {
        const int slot = gi & 1;
        const int b0 = g_begin[gi], nb = g_begin[gi + 1] - b0;
        const long long f0 = off_host[b0], nf = off_host[b0 + nb] - f0;
        off_g.assign((size_t)nb + 1, 0);
        for (int i = 0; i <= nb; ++i) off_g[i] = off_host[b0 + i] - (int32_t)f0;
        if (gi >= 2) SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->h2d_stream, e->ev_comp[slot], 0));
        if (nf > 0)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->stage[slot], frames + (size_t)f0 * frame_bytes, (size_t)nf * frame_bytes,
                                              cudaMemcpyHostToDevice, e->h2d_stream));
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_in[slot], e->h2d_stream));
        SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->compute_stream, e->ev_in[slot], 0));
        if (gi >= 2) SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->compute_stream, e->ev_out[slot], 0));
        int32_t* d_idx = e->idx_stage[slot];
        int32_t* d_status = d_idx + (size_t)max_clips * K;
        float* d_out = sampled_host ? e->out_stage[slot] : nullptr;
        if (int rc2 = mdf_sample_ragged_device(e, e->stage[slot], nullptr, nb, off_g.data(), H, Wd, K, W, d_idx, d_status,
                                               nullptr, nullptr, d_out, e->compute_stream))
            return rc2;
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_comp[slot], e->compute_stream));
        SASVQA_CUDA_CHECK(cudaStreamWaitEvent(e->d2h_stream, e->ev_comp[slot], 0));
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(idx_host + (size_t)b0 * K, d_idx, (size_t)nb * K * sizeof(int32_t),
                                          cudaMemcpyDeviceToHost, e->d2h_stream));
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(status_host + b0, d_status, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                          e->d2h_stream));
        if (sampled_host)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(sampled_host + (size_t)b0 * K * kFrameElems, d_out,
                                              (size_t)nb * K * kFrameElems * sizeof(float), cudaMemcpyDeviceToHost,
                                              e->d2h_stream));
        SASVQA_CUDA_CHECK(cudaEventRecord(e->ev_out[slot], e->d2h_stream));
        return 0;
    };
    rc = 0;
    for (int gi = 0; gi < n_groups && rc == 0; ++gi) rc = run_group(gi);
    const int rc_drain = drain_pipeline(e);
    return rc ? rc : rc_drain;
}

}  // namespace sasvqa
