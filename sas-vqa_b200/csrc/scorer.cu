// MIF relevance model on the GPU: the BERT sequence classifier the reference scores (question, caption) pairs with.
// Replaces   model = AutoModelForSequenceClassification.from_pretrained(args.sim_model)   gen_sample.py:160
//            model.eval().cuda(); output = model(**inputs); scores = output[0][:,0]       gen_sample.py:49,82-83
// (HF transformers BertForSequenceClassification, modeling_bert.py:53-468,1077-1155; bert-base geometry = the ViT's:
// 768 hidden, 12 x 64 heads, FFN 3072, so every dense layer runs on the encoder's tcgen05 GEMM kernel.)
//
// B200-first differences from the reference's execution, none of which change a logit:
//  * sequences are PACKED -- the tokenizer's right padding is dropped on entry (lengths = attention_mask.sum(1)),
//    the GEMMs run over real tokens only and the key mask becomes a per-sequence length in the attention kernel;
//  * many QA samples are scored per call (the reference runs one tokenizer + model call per sample);
//  * post-LN residuals: the GEMM epilogue adds dense(.) + bias into the fp32 stream in L2 (TMA reduce-add), one
//    kernel then normalises in place and emits the bf16 operand of the next GEMM.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/sasvqa.h"
#include "common.cuh"

namespace sasvqa {

namespace {

constexpr float kBertLnEps = 1e-12f;
constexpr int kBertMaxPos = 512;
constexpr int kBertTypes = 2;
constexpr int kDefaultMaxTokens = 65536;

__global__ void scorer_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}
// scores[g, t] = logits[(g*T + t), label]
__global__ void take_label_kernel(const float* __restrict__ logits, int n, int labels, int label, float* __restrict__ scores) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scores[i] = logits[(long long)i * labels + label];
}

struct BertLayer {
    __nv_bfloat16 *w_qkv, *w_out, *w_fc1, *w_fc2;
    float *b_qkv, *b_out, *ln1_g, *ln1_b, *b_fc1, *b_fc2, *ln2_g, *ln2_b;
    CUtensorMap m_qkv, m_out, m_fc1, m_fc2;
};

enum ScorerProfKind { SP_EMBED = 0, SP_LN, SP_GEMM_QKV, SP_ATTENTION, SP_GEMM_OUT, SP_GEMM_FC1, SP_GEMM_FC2, SP_POOLER, SP_COUNT };
struct ScorerProfRec {
    int kind;
    cudaEvent_t a, b;
};

}  // namespace

}  // namespace sasvqa

using namespace sasvqa;

struct SasvqaScorer {
    int device = 0;
    int num_sms = 148;
    int vocab = 0, labels = 0, max_tokens = 0;
    __nv_bfloat16* arena_bf16 = nullptr;
    float* arena_f32 = nullptr;
    float *word = nullptr, *pos = nullptr, *type_emb = nullptr, *emb_g = nullptr, *emb_b = nullptr;
    float *w_pool = nullptr, *b_pool = nullptr, *w_cls = nullptr, *b_cls = nullptr;
    BertLayer L[kLayers];
    float* x = nullptr;             // [max_tokens, 768] fp32 stream (post-LN values)
    __nv_bfloat16* h = nullptr;     // [max_tokens, 768] bf16 GEMM operand (LN output / attention output)
    __nv_bfloat16* big = nullptr;   // [max_tokens, 3072] q|k|v (as [., 2304]) / intermediate
    CUtensorMap m_h, m_big_fc, m_out_qkv, m_out_fc1, m_out_x;
    int32_t* cu_dev = nullptr;      // [N + 1] packed row offsets
    size_t cu_cap = 0;
    // host-entry staging
    int32_t *ids_dev = nullptr, *type_dev = nullptr;
    size_t ids_cap = 0, type_cap = 0;
    float* logits_dev = nullptr;
    size_t logits_cap = 0;
    float* scores_dev = nullptr;
    size_t scores_cap = 0;
    int32_t* idx_dev = nullptr;
    size_t idx_cap = 0;
    cudaStream_t stream = nullptr;
    WorkspaceOrder order;           // serialises the entry points across the streams they are called on
    bool profile = false;
    std::vector<ScorerProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
};

namespace sasvqa {

namespace {

struct SScope {
    SasvqaScorer* e;
    int kind;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(SasvqaScorer* e) {
        if (!e->ev_pool.empty()) {
            cudaEvent_t ev = e->ev_pool.back();
            e->ev_pool.pop_back();
            return ev;
        }
        cudaEvent_t ev = nullptr;
        cudaEventCreate(&ev);
        return ev;
    }
    SScope(SasvqaScorer* e_, int kind_, cudaStream_t s_) : e(e_), kind(kind_), s(s_) {
        if (e->profile) {
            a = take(e);
            b = take(e);
            cudaEventRecord(a, s);
        }
    }
    ~SScope() {
        if (e->profile) {
            cudaEventRecord(b, s);
            e->prof.push_back({kind, a, b});
        }
    }
};

int sgemm(SasvqaScorer* e, int kind, const GemmArgs& g, const CUtensorMap* ma, const CUtensorMap* mb, const CUtensorMap* mo,
          cudaStream_t s) {
    SScope sc(e, kind, s);
    return launch_gemm_tcgen05(g, ma, mb, mo, e->num_sms, s);
}

int sgrow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) SASVQA_CUDA_CHECK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SASVQA_CUDA_CHECK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

// sequences [s0, s1) of the padded id matrix (rows of L ids) -> hidden state after n_layers blocks in e->x, packed
int forward_chunk(SasvqaScorer* e, const int32_t* ids, const int32_t* type_ids, int s0, int s1, int L, int row_base, int M,
                  int max_len, int n_layers, cudaStream_t s) {
    const int n = s1 - s0;
    const int32_t* cu = e->cu_dev + s0;
    int rc;
    {
        SScope sc(e, SP_EMBED, s);
        if ((rc = launch_embed_layernorm(ids + (size_t)s0 * L, type_ids ? type_ids + (size_t)s0 * L : nullptr, cu, row_base, n,
                                         L, e->vocab, kBertTypes, e->word, e->pos, e->type_emb, e->emb_g, e->emb_b,
                                         kBertLnEps, e->x, e->h, s)))
            return rc;
    }
    for (int l = 0; l < n_layers; ++l) {
        BertLayer& Ly = e->L[l];
        GemmArgs g{};
        g.A = e->h; g.B = Ly.w_qkv; g.M = M; g.N = kQkv; g.K = kHidden;
        g.epilogue = EPI_BIAS_BF16; g.bias = Ly.b_qkv; g.out_bf16 = e->big;
        if ((rc = sgemm(e, SP_GEMM_QKV, g, &e->m_h, &Ly.m_qkv, &e->m_out_qkv, s))) return rc;
        {
            SScope sc(e, SP_ATTENTION, s);
            if ((rc = launch_attention_varlen(e->big, e->h, cu, row_base, n, max_len, s))) return rc;
        }
        g = GemmArgs{};
        g.A = e->h; g.B = Ly.w_out; g.M = M; g.N = kHidden; g.K = kHidden;
        g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_out; g.out_f32 = e->x;
        if ((rc = sgemm(e, SP_GEMM_OUT, g, &e->m_h, &Ly.m_out, &e->m_out_x, s))) return rc;
        {
            SScope sc(e, SP_LN, s);
            if ((rc = launch_layernorm_post(e->x, e->h, M, Ly.ln1_g, Ly.ln1_b, kBertLnEps, s))) return rc;
        }
        g = GemmArgs{};
        g.A = e->h; g.B = Ly.w_fc1; g.M = M; g.N = kFfn; g.K = kHidden;
        g.epilogue = EPI_BIAS_ERF_GELU_BF16; g.bias = Ly.b_fc1; g.out_bf16 = e->big;
        if ((rc = sgemm(e, SP_GEMM_FC1, g, &e->m_h, &Ly.m_fc1, &e->m_out_fc1, s))) return rc;
        g = GemmArgs{};
        g.A = e->big; g.B = Ly.w_fc2; g.M = M; g.N = kHidden; g.K = kFfn;
        g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_fc2; g.out_f32 = e->x;
        if ((rc = sgemm(e, SP_GEMM_FC2, g, &e->m_big_fc, &Ly.m_fc2, &e->m_out_x, s))) return rc;
        {
            SScope sc(e, SP_LN, s);
            if ((rc = launch_layernorm_post(e->x, e->h, M, Ly.ln2_g, Ly.ln2_b, kBertLnEps, s))) return rc;
        }
    }
    return 0;
}

// lengths -> packed offsets (host) + upload; validates every length
int upload_offsets(SasvqaScorer* e, const int32_t* lengths_host, int N, int L, std::vector<int32_t>& cu, cudaStream_t s) {
    cu.resize((size_t)N + 1);
    cu[0] = 0;
    for (int i = 0; i < N; ++i) {
        SASVQA_REQUIRE(lengths_host[i] >= 1 && lengths_host[i] <= L, "sequence length outside [1, L] (a tokenized pair holds at least [CLS] and [SEP])");
        SASVQA_REQUIRE((long long)cu[i] + lengths_host[i] < 2147483647ll, "more than 2^31 tokens in one call");
        cu[i + 1] = cu[i] + lengths_host[i];
    }
    if (int rc = sgrow((void**)&e->cu_dev, &e->cu_cap, ((size_t)N + 1) * sizeof(int32_t))) return rc;
    SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->cu_dev, cu.data(), ((size_t)N + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    return 0;
}

}  // namespace

uint64_t scorer_num_params(int vocab, int labels) {
    const uint64_t H = kHidden, F = kFfn;
    return (uint64_t)vocab * H + (uint64_t)kBertMaxPos * H + (uint64_t)kBertTypes * H + 2 * H +
           (uint64_t)kLayers * (4 * (H * H + H) + 2 * H + (F * H + F) + (H * F + H) + 2 * H) + (H * H + H) +
           ((uint64_t)labels * H + labels);
}

void scorer_destroy(SasvqaScorer* e) {
    if (!e) return;
    cudaFree(e->arena_bf16); cudaFree(e->arena_f32);
    cudaFree(e->x); cudaFree(e->h); cudaFree(e->big);
    cudaFree(e->cu_dev); cudaFree(e->ids_dev); cudaFree(e->type_dev);
    cudaFree(e->logits_dev); cudaFree(e->scores_dev); cudaFree(e->idx_dev);
    for (const ScorerProfRec& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->order.tail) cudaEventDestroy(e->order.tail);
    delete e;
}

int scorer_create(const float* params_host, uint64_t n_params, int vocab, int labels, int max_tokens, SasvqaScorer** out) {
    SASVQA_REQUIRE(out != nullptr && params_host != nullptr, "null argument");
    SASVQA_REQUIRE(vocab >= 1 && labels >= 1 && labels <= 64, "bad vocabulary size / label count");
    SASVQA_REQUIRE(n_params == scorer_num_params(vocab, labels),
                   "state dict size does not match a bert-base sequence classifier with this vocabulary and label count");
    if (max_tokens <= 0) max_tokens = kDefaultMaxTokens;
    SASVQA_REQUIRE(max_tokens >= kBertMaxPos, "max_tokens must hold at least one 512-token sequence");
    SasvqaScorer* e = new SasvqaScorer();
    auto fail = [&](int rc) { scorer_destroy(e); return rc; };
#define TRY(expr) do { int _rc = (expr); if (_rc) return fail(_rc); } while (0)
#define TRYCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return fail(SASVQA_ERR_CUDA); } } while (0)
    TRYCUDA(cudaGetDevice(&e->device));
    cudaDeviceProp prop;
    TRYCUDA(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major != 10) {
        set_last_error("sasvqa_b200 needs an sm_100a GPU (B200); found compute capability " + std::to_string(prop.major) +
                       "." + std::to_string(prop.minor));
        return fail(SASVQA_ERR_INVALID);
    }
    e->num_sms = prop.multiProcessorCount;
    e->vocab = vocab;
    e->labels = labels;
    e->max_tokens = max_tokens;

    float* raw = nullptr;
    TRYCUDA(cudaMalloc(&raw, n_params * sizeof(float)));
    if (cudaMemcpy(raw, params_host, n_params * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("uploading scorer parameters failed");
        return fail(SASVQA_ERR_CUDA);
    }
    const size_t H = kHidden, F = kFfn;
    const size_t n_mat = (size_t)kLayers * (3 * H * H + H * H + 2 * F * H);
    const size_t n_vec = (size_t)vocab * H + kBertMaxPos * H + kBertTypes * H + 2 * H +
                         (size_t)kLayers * (3 * H + H + 2 * H + F + H + 2 * H) + H * H + H + (size_t)labels * H + labels;
    if (cudaMalloc(&e->arena_bf16, n_mat * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&e->arena_f32, n_vec * sizeof(float)) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("allocating scorer weights failed");
        return fail(SASVQA_ERR_NOMEM);
    }
    __nv_bfloat16* mp = e->arena_bf16;
    float* vp = e->arena_f32;
    const float* rp = raw;
    auto take_mat = [&](size_t n) {
        __nv_bfloat16* dst = mp;
        scorer_f32_to_bf16_kernel<<<592, 256>>>(rp, dst, (long long)n);
        mp += n; rp += n;
        return dst;
    };
    auto take_vec = [&](size_t n) {
        float* dst = vp;
        cudaMemcpyAsync(dst, rp, n * sizeof(float), cudaMemcpyDeviceToDevice, 0);
        vp += n; rp += n;
        return dst;
    };
    e->word = take_vec((size_t)vocab * H);
    e->pos = take_vec((size_t)kBertMaxPos * H);
    e->type_emb = take_vec((size_t)kBertTypes * H);
    e->emb_g = take_vec(H);
    e->emb_b = take_vec(H);
    for (int l = 0; l < kLayers; ++l) {
        BertLayer& Ly = e->L[l];
        // HF order: query, key, value (weight then bias each) = the fused projection's q | k | v row order
        Ly.w_qkv = mp; mp += 3 * H * H;
        Ly.b_qkv = vp; vp += 3 * H;
        for (int j = 0; j < 3; ++j) {
            scorer_f32_to_bf16_kernel<<<592, 256>>>(rp, Ly.w_qkv + j * H * H, (long long)(H * H));
            rp += H * H;
            cudaMemcpyAsync(Ly.b_qkv + j * H, rp, H * sizeof(float), cudaMemcpyDeviceToDevice, 0);
            rp += H;
        }
        Ly.w_out = take_mat(H * H);
        Ly.b_out = take_vec(H);
        Ly.ln1_g = take_vec(H);
        Ly.ln1_b = take_vec(H);
        Ly.w_fc1 = take_mat(F * H);
        Ly.b_fc1 = take_vec(F);
        Ly.w_fc2 = take_mat(H * F);
        Ly.b_fc2 = take_vec(H);
        Ly.ln2_g = take_vec(H);
        Ly.ln2_b = take_vec(H);
    }
    e->w_pool = take_vec(H * H);
    e->b_pool = take_vec(H);
    e->w_cls = take_vec((size_t)labels * H);
    e->b_cls = take_vec(labels);
    cudaError_t ce = cudaDeviceSynchronize();
    cudaFree(raw);
    if (ce != cudaSuccess || (size_t)(rp - raw) != n_params || (size_t)(mp - e->arena_bf16) != n_mat ||
        (size_t)(vp - e->arena_f32) != n_vec) {
        set_last_error(std::string("scorer weight conversion failed: ") + cudaGetErrorString(ce));
        return fail(SASVQA_ERR_CUDA);
    }

    const size_t rows = (size_t)max_tokens;
    if (cudaMalloc(&e->x, rows * H * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&e->h, rows * H * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&e->big, rows * F * sizeof(__nv_bfloat16)) != cudaSuccess) {
        set_last_error("allocating scorer workspace failed (lower max_tokens)");
        return fail(SASVQA_ERR_NOMEM);
    }
    TRYCUDA(cudaMemset(e->x, 0, rows * H * sizeof(float)));
    TRYCUDA(cudaMemset(e->h, 0, rows * H * sizeof(__nv_bfloat16)));
    TRYCUDA(cudaMemset(e->big, 0, rows * F * sizeof(__nv_bfloat16)));
    TRY(make_tensor_map_out(&e->m_out_qkv, e->big, rows, kQkv, 0));
    TRY(make_tensor_map_out(&e->m_out_fc1, e->big, rows, kFfn, 0));
    TRY(make_tensor_map_out(&e->m_out_x, e->x, rows, kHidden, 1));
    TRY(make_tensor_map_bf16_kmajor(&e->m_h, e->h, rows, kHidden, 128));
    TRY(make_tensor_map_bf16_kmajor(&e->m_big_fc, e->big, rows, kFfn, 128));
    for (int l = 0; l < kLayers; ++l) {
        BertLayer& Ly = e->L[l];
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_qkv, Ly.w_qkv, kQkv, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_out, Ly.w_out, kHidden, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_fc1, Ly.w_fc1, kFfn, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_fc2, Ly.w_fc2, kHidden, kFfn, 128));
    }
    TRYCUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    TRYCUDA(cudaEventCreateWithFlags(&e->order.tail, cudaEventDisableTiming));
#undef TRY
#undef TRYCUDA
    *out = e;
    return 0;
}

int scorer_max_tokens(const SasvqaScorer* e) { return e ? e->max_tokens : 0; }

// device ids [N, L] (right-padded), host lengths [N] -> logits [N, labels] on the device.  Sequences are taken in
// groups whose packed token count fits the workspace.  hidden_or_null: packed fp32 hidden state after n_layers
// blocks (inspection; the whole call must then fit one group).
int scorer_logits(SasvqaScorer* e, const int32_t* ids, const int32_t* type_ids, const int32_t* lengths_host, int N, int L,
                  float* logits, int n_layers, float* hidden_or_null, cudaStream_t s) {
    SASVQA_REQUIRE(e != nullptr && N >= 0 && L >= 0, "bad arguments");
    if (N == 0) return 0;
    SASVQA_REQUIRE(L >= 1 && L <= kBertMaxPos, "padded length must be in [1, 512] (BERT position table)");
    SASVQA_REQUIRE(ids != nullptr && lengths_host != nullptr && (logits != nullptr || hidden_or_null != nullptr), "null argument");
    SASVQA_REQUIRE(n_layers >= 0 && n_layers <= kLayers, "bad layer count");
    StreamOrder order(&e->order, s);
    std::vector<int32_t> cu;
    if (int rc = upload_offsets(e, lengths_host, N, L, cu, s)) return rc;
    SASVQA_REQUIRE(hidden_or_null == nullptr || cu[N] <= e->max_tokens, "hidden-state inspection needs one group");
    int s0 = 0;
    while (s0 < N) {
        int s1 = s0, max_len = 0;
        while (s1 < N && cu[s1 + 1] - cu[s0] <= e->max_tokens) {
            max_len = std::max(max_len, cu[s1 + 1] - cu[s1]);
            ++s1;
        }
        const int M = cu[s1] - cu[s0];
        if (M > 0) {
            if (int rc = forward_chunk(e, ids, type_ids, s0, s1, L, cu[s0], M, max_len, n_layers, s)) return rc;
            if (hidden_or_null)
                SASVQA_CUDA_CHECK(cudaMemcpyAsync(hidden_or_null, e->x, (size_t)M * kHidden * sizeof(float),
                                                  cudaMemcpyDeviceToDevice, s));
        }
        if (logits) {
            SScope sc(e, SP_POOLER, s);
            if (int rc = launch_pooler_classifier(e->x, e->cu_dev + s0, cu[s0], s1 - s0, e->w_pool, e->b_pool, e->w_cls,
                                                  e->b_cls, e->labels, logits + (size_t)s0 * e->labels, s))
                return rc;
        }
        s0 = s1;
    }
    return 0;
}

namespace {

// tokenizer output (int64 like return_tensors='pt') -> int32 ids + lengths; the mask must be a prefix of ones
int pack_host_inputs(const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int N, int L, int vocab,
                     std::vector<int32_t>& ids32, std::vector<int32_t>& type32, std::vector<int32_t>& lens) {
    ids32.resize((size_t)N * L);
    if (type_ids) type32.resize((size_t)N * L);
    lens.resize(N);
    for (int i = 0; i < N; ++i) {
        int len = L;
        if (mask) {
            len = 0;
            while (len < L && mask[(size_t)i * L + len] != 0) ++len;
            for (int p = len; p < L; ++p)
                SASVQA_REQUIRE(mask[(size_t)i * L + p] == 0, "attention_mask must be right-padded (ones then zeros)");
        }
        lens[i] = len;
        for (int p = 0; p < L; ++p) {
            const int64_t id = ids[(size_t)i * L + p];
            SASVQA_REQUIRE(p >= len || (id >= 0 && id < vocab), "input id outside the vocabulary");
            ids32[(size_t)i * L + p] = p < len ? (int32_t)id : 0;
            if (type_ids) {
                const int64_t tt = type_ids[(size_t)i * L + p];
                SASVQA_REQUIRE(p >= len || (tt >= 0 && tt < kBertTypes), "token type id outside {0, 1}");
                type32[(size_t)i * L + p] = p < len ? (int32_t)tt : 0;
            }
        }
    }
    return 0;
}

int stage_host_inputs(SasvqaScorer* e, const std::vector<int32_t>& ids32, const std::vector<int32_t>& type32, bool has_type,
                      int N) {
    if (int rc = sgrow((void**)&e->ids_dev, &e->ids_cap, std::max<size_t>(ids32.size(), 1) * sizeof(int32_t))) return rc;
    SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->ids_dev, ids32.data(), ids32.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (has_type) {
        if (int rc = sgrow((void**)&e->type_dev, &e->type_cap, std::max<size_t>(type32.size(), 1) * sizeof(int32_t))) return rc;
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(e->type_dev, type32.data(), type32.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    }
    return sgrow((void**)&e->logits_dev, &e->logits_cap, std::max<size_t>((size_t)N * e->labels, 1) * sizeof(float));
}

}  // namespace

// host arrays exactly as the tokenizer returns them -> logits [N, labels] in host memory
int scorer_logits_host(SasvqaScorer* e, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int N, int L,
                       float* logits_host) {
    SASVQA_REQUIRE(e != nullptr && N >= 0 && L >= 0, "bad arguments");
    if (N == 0) return 0;
    SASVQA_REQUIRE(ids != nullptr && logits_host != nullptr, "null argument");
    SASVQA_REQUIRE(L >= 1 && L <= kBertMaxPos, "padded length must be in [1, 512] (BERT position table)");
    std::vector<int32_t> ids32, type32, lens;
    if (int rc = pack_host_inputs(ids, type_ids, mask, N, L, e->vocab, ids32, type32, lens)) return rc;
    auto run = [&]() -> int {
        if (int rc = stage_host_inputs(e, ids32, type32, type_ids != nullptr, N)) return rc;
        if (int rc = scorer_logits(e, e->ids_dev, type_ids ? e->type_dev : nullptr, lens.data(), N, L, e->logits_dev, kLayers,
                                   nullptr, e->stream))
            return rc;
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(logits_host, e->logits_dev, (size_t)N * e->labels * sizeof(float),
                                          cudaMemcpyDeviceToHost, e->stream));
        return 0;
    };
    const int rc = run();
    const cudaError_t ce = cudaStreamSynchronize(e->stream);    // on failure too: nothing stays in flight on the caller's buffers
    if (rc) return rc;
    SASVQA_CUDA_CHECK(ce);
    return 0;
}

// gen_sample.py:79-88 for G QA samples of T captions each: ids [G*T, L] -> logits -> scores = logits[:, label] ->
// idx[g] = ds_rate * topk(scores[g, ::ds_rate], K), best first
int mif_select_captions_host(SasvqaScorer* e, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int G, int T,
                             int L, int K, int ds_rate, int label, int32_t* idx_host, float* scores_host) {
    SASVQA_REQUIRE(e != nullptr && G >= 0 && T >= 0 && K >= 0 && ds_rate >= 1, "bad arguments");
    SASVQA_REQUIRE(label >= 0 && label < e->labels, "label outside the classifier's outputs");
    if (G == 0) return 0;
    const int n_cand = T == 0 ? 0 : (T + ds_rate - 1) / ds_rate;
    SASVQA_REQUIRE(K <= n_cand, "selected index k out of range: fewer than K candidate captions (torch.topk raises here)");
    if (K == 0 && scores_host == nullptr) return 0;
    SASVQA_REQUIRE(ids != nullptr && (K == 0 || idx_host != nullptr), "null argument");
    const int N = G * T;
    std::vector<int32_t> ids32, type32, lens;
    if (int rc = pack_host_inputs(ids, type_ids, mask, N, L, e->vocab, ids32, type32, lens)) return rc;
    auto run = [&]() -> int {
        if (int rc = stage_host_inputs(e, ids32, type32, type_ids != nullptr, N)) return rc;
        if (int rc = scorer_logits(e, e->ids_dev, type_ids ? e->type_dev : nullptr, lens.data(), N, L, e->logits_dev, kLayers,
                                   nullptr, e->stream))
            return rc;
        if (int rc = sgrow((void**)&e->scores_dev, &e->scores_cap, (size_t)N * sizeof(float))) return rc;
        if (int rc = sgrow((void**)&e->idx_dev, &e->idx_cap, std::max<size_t>((size_t)G * K, 1) * sizeof(int32_t))) return rc;
        StreamOrder order(&e->order, e->stream);
        take_label_kernel<<<(N + 255) / 256, 256, 0, e->stream>>>(e->logits_dev, N, e->labels, label, e->scores_dev);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
        if (K > 0) {
            if (int rc = launch_topk_strided(e->scores_dev, G, T, ds_rate, K, e->idx_dev, nullptr, e->stream)) return rc;
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(idx_host, e->idx_dev, (size_t)G * K * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                              e->stream));
        }
        if (scores_host)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(scores_host, e->scores_dev, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost,
                                              e->stream));
        return 0;
    };
    const int rc = run();
    const cudaError_t ce = cudaStreamSynchronize(e->stream);    // on failure too: nothing stays in flight on the caller's buffers
    if (rc) return rc;
    SASVQA_CUDA_CHECK(ce);
    return 0;
}

int scorer_profile_enable(SasvqaScorer* e, int on) {
    SASVQA_REQUIRE(e != nullptr, "null scorer");
    e->profile = on != 0;
    return 0;
}

int scorer_profile_read(SasvqaScorer* e, double* ms, int64_t* scopes, int n_kinds) {
    SASVQA_REQUIRE(e != nullptr && ms != nullptr && scopes != nullptr && n_kinds >= SP_COUNT, "bad arguments");
    SASVQA_CUDA_CHECK(cudaDeviceSynchronize());
    for (int k = 0; k < n_kinds; ++k) {
        ms[k] = 0.0;
        scopes[k] = 0;
    }
    for (const ScorerProfRec& r : e->prof) {
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

}  // namespace sasvqa
