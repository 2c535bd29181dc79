// Downstream consumer of the sampled frames: the video-QA forward of the reference's GIT model
// (src/modeling/modeling.py:29-232, MyGitModel / MyGitForCausalLM over HF transformers GitModel), inference logits.
//   visual side   per-frame image_encoder(...).last_hidden_state, concatenated, visual_projection   modeling.py:76-95
//                 -> the sampler's own encoder kernels (encoder.cu: visual_tokens), all K frames of all samples in one batch
//   text side     GitEmbeddings (word + position -> LayerNorm)                                       modeling.py:97-102
//   decoder       6 post-LN BERT-style blocks over [visual rows | text rows] with the combined mask   modeling.py:114-152
//   head          logits = output(sequence_output)                                                   modeling.py:206-207
// git-base has the ViT's geometry (768 hidden, 12 x 64 heads, FFN 3072): every dense layer is the tcgen05 GEMM kernel.
//
// B200-first differences, none of which change a text-row logit:
//  * rows of a group of samples are stored visual-first (all visual rows, then all text rows), so the encoder writes
//    the projected visual tokens straight into the decoder's residual stream and the text rows are one contiguous
//    A operand for the output head;
//  * the head runs on the text rows only -- the reference also computes the K*197 visual rows' logits (97 % of that
//    30522-wide GEMM) and slices them away (modeling.py:211-215);
//  * the mask is never materialised: "keys [0, limit(row))" inside the attention kernel (attention.cu).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/sasvqa.h"
#include "common.cuh"

namespace sasvqa {

int visual_tokens(SasvqaEncoder*, const uint8_t*, const float*, int, int, float*, cudaStream_t);

namespace {

constexpr float kGitLnEps = 1e-12f;
constexpr int kGitMaxPos = 1024;
constexpr int kGitMaxLayers = 12;
constexpr int kDefaultMaxRows = 65536;

__global__ void git_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}
// fp32 rows -> bf16 rows, 8 values per thread (the visual rows the encoder wrote into the residual stream)
__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = src[2 * i], b = src[2 * i + 1];
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y);
        o.w = pack_bf16x2(b.z, b.w);
        dst[i] = o;
    }
}

// Next-token cross-entropy of the text rows (MyGitForCausalLM.forward, modeling.py:208-215: logits[:, n_vis:-1] against
// labels[:, 1:], nn.CrossEntropyLoss defaults = mean over labels != -100).  One CTA per (sample, position < L - 1):
// a single pass keeps a running (max, sum of exp) per thread, a block reduction combines them, thread 0 writes
// logsumexp - logit[label].  A second one-CTA kernel sums the rows in a fixed order (no atomics: run-to-run identical).
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, int vocab, int vocab_pad, int L,
                                                       const int32_t* __restrict__ labels, float* __restrict__ row_loss,
                                                       int32_t* __restrict__ row_valid) {
    __shared__ float sm[8], ss[8];
    const int b = blockIdx.x / (L - 1), t = blockIdx.x - b * (L - 1);
    const int label = labels[(long long)b * L + t + 1];
    if (label < 0 || label >= vocab) {                       // ignore_index (-100); out-of-range labels are ignored too
        if (threadIdx.x == 0) {
            row_loss[blockIdx.x] = 0.f;
            row_valid[blockIdx.x] = 0;
        }
        return;
    }
    const float* row = logits + ((long long)b * L + t) * vocab_pad;
    float m = -INFINITY, sum = 0.f;
    for (int i = threadIdx.x; i < vocab; i += 256) {
        const float v = row[i];
        if (v > m) {
            sum = sum * __expf(m - v) + 1.f;
            m = v;
        } else {
            sum += __expf(v - m);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
        const float mn = fmaxf(m, m2);
        sum = (m == -INFINITY ? 0.f : sum * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
        m = mn;
    }
    if (lane == 0) {
        sm[warp] = m;
        ss[warp] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mt = sm[0], st = ss[0];
        for (int w = 1; w < 8; ++w) {
            const float mn = fmaxf(mt, sm[w]);
            st = (mt == -INFINITY ? 0.f : st * __expf(mt - mn)) + (sm[w] == -INFINITY ? 0.f : ss[w] * __expf(sm[w] - mn));
            mt = mn;
        }
        row_loss[blockIdx.x] = (mt + logf(st)) - row[label];
        row_valid[blockIdx.x] = 1;
    }
}
__global__ void __launch_bounds__(256) ce_mean_kernel(const float* __restrict__ row_loss, const int32_t* __restrict__ row_valid,
                                                       int n, float* __restrict__ loss) {
    __shared__ double acc[256];
    __shared__ int cnt[256];
    double a = 0.0;
    int c = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += (double)row_loss[i];
        c += row_valid[i];
    }
    acc[threadIdx.x] = a;
    cnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            acc[threadIdx.x] += acc[threadIdx.x + o];
            cnt[threadIdx.x] += cnt[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(acc[0] / (double)cnt[0]);      // no valid label: 0 / 0 = NaN, as torch
}

// Greedy decoding step (HF generate, greedy search): next = argmax(logits[row]) (lowest index among equals); a sequence that
// has produced eos keeps emitting pad.  One CTA per sample; writes ids[s, pos] and updates finished[s].
__global__ void __launch_bounds__(256) greedy_next_kernel(const float* __restrict__ logits, int vocab, int vocab_pad,
                                                           int32_t* __restrict__ ids, int Lmax, int pos, int eos, int pad,
                                                           int32_t* __restrict__ finished) {
    __shared__ float bv[8];
    __shared__ int bi[8];
    const int smp = blockIdx.x;
    const float* row = logits + (long long)smp * vocab_pad;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += 256) {
        const float v = row[i];
        if (v > best || arg == 0x7fffffff) {        // strided scan: i increases, so '>' keeps the lowest index per thread
            best = v;
            arg = i;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oi < arg)) {
            best = ov;
            arg = oi;
        }
    }
    if (lane == 0) {
        bv[warp] = best;
        bi[warp] = arg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < arg)) {
                best = bv[w];
                arg = bi[w];
            }
        const int done = finished[smp];
        const int tok = done ? pad : arg;
        ids[(long long)smp * Lmax + pos] = tok;
        if (!done && tok == eos) finished[smp] = 1;
    }
}
__global__ void init_generate_kernel(const int32_t* __restrict__ prompt, int L0, int32_t* __restrict__ ids, int Lmax, int pad,
                                     int32_t* __restrict__ finished, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * Lmax) {
        const int smp = i / Lmax, p = i - smp * Lmax;
        ids[i] = p < L0 ? prompt[(long long)smp * L0 + p] : pad;
    }
    if (i < n) finished[i] = 0;
}

struct GitLayer {
    __nv_bfloat16 *w_qkv, *w_out, *w_fc1, *w_fc2;
    float *b_qkv, *b_out, *ln1_g, *ln1_b, *b_fc1, *b_fc2, *ln2_g, *ln2_b;
    CUtensorMap m_qkv, m_out, m_fc1, m_fc2;
};

}  // namespace

}  // namespace sasvqa

using namespace sasvqa;

struct SasvqaGitDecoder {
    int device = 0;
    int num_sms = 148;
    int vocab = 0, vocab_pad = 0, n_layers = 0, max_rows = 0;
    WorkspaceOrder order;              // serialises the entry points across the streams they are called on
    __nv_bfloat16* arena_bf16 = nullptr;
    float* arena_f32 = nullptr;
    float *word = nullptr, *pos = nullptr, *emb_g = nullptr, *emb_b = nullptr, *zero_row = nullptr;
    __nv_bfloat16* w_head = nullptr;   // [vocab_pad, 768], rows >= vocab are zero
    float* b_head = nullptr;           // [vocab_pad]
    CUtensorMap m_head;
    GitLayer L[kGitMaxLayers];
    float* x = nullptr;                // [max_rows, 768] fp32 stream
    __nv_bfloat16* h = nullptr;        // [max_rows, 768]
    __nv_bfloat16* big = nullptr;      // [max_rows, 3072]
    CUtensorMap m_h, m_big_fc, m_out_qkv, m_out_fc1, m_out_x;
    int32_t* cu_dev = nullptr;
    size_t cu_cap = 0;
    // loss path scratch: logits of one group, per-row losses of the whole call
    float* logits_scratch = nullptr;
    size_t logits_cap = 0;
    // incremental decoding: per-layer k | v of the visual rows of one group, ids / flags / last-row scratch
    __nv_bfloat16* kv_cache = nullptr;   // [n_layers][kv_rows, 1536]
    size_t kv_cap = 0;
    int32_t* gen_ids = nullptr;          // [group, Lmax]
    size_t gen_ids_cap = 0;
    int32_t* gen_done = nullptr;         // [group]
    size_t gen_done_cap = 0;
    __nv_bfloat16* last_h = nullptr;     // [group, 768] hidden state of the newest position
    size_t last_h_cap = 0;
    float* row_loss = nullptr;
    size_t row_loss_cap = 0;
    int32_t* row_valid = nullptr;
    size_t row_valid_cap = 0;
};

namespace sasvqa {

namespace {

int dgrow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) SASVQA_CUDA_CHECK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SASVQA_CUDA_CHECK(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

int dgemm(SasvqaGitDecoder* d, const GemmArgs& g, const CUtensorMap* ma, const CUtensorMap* mb, const CUtensorMap* mo,
          cudaStream_t s) {
    return launch_gemm_tcgen05(g, ma, mb, mo, d->num_sms, s);
}

}  // namespace

uint64_t git_decoder_num_params(int vocab, int n_layers) {
    const uint64_t H = kHidden, F = kFfn;
    return (uint64_t)vocab * H + (uint64_t)kGitMaxPos * H + 2 * H +
           (uint64_t)n_layers * (4 * (H * H + H) + 2 * H + (F * H + F) + (H * F + H) + 2 * H) + ((uint64_t)vocab * H + vocab);
}

void git_decoder_destroy(SasvqaGitDecoder* d) {
    if (!d) return;
    cudaFree(d->arena_bf16); cudaFree(d->arena_f32);
    cudaFree(d->x); cudaFree(d->h); cudaFree(d->big); cudaFree(d->cu_dev);
    cudaFree(d->logits_scratch); cudaFree(d->row_loss); cudaFree(d->row_valid);
    cudaFree(d->kv_cache); cudaFree(d->gen_ids); cudaFree(d->gen_done); cudaFree(d->last_h);
    if (d->order.tail) cudaEventDestroy(d->order.tail);
    delete d;
}

int git_decoder_create(const float* params_host, uint64_t n_params, int vocab, int n_layers, int max_rows,
                       SasvqaGitDecoder** out) {
    SASVQA_REQUIRE(out != nullptr && params_host != nullptr, "null argument");
    SASVQA_REQUIRE(vocab >= 1 && n_layers >= 1 && n_layers <= kGitMaxLayers, "bad vocabulary size / layer count");
    SASVQA_REQUIRE(n_params == git_decoder_num_params(vocab, n_layers),
                   "state dict size does not match a git-base text decoder with this vocabulary and layer count");
    if (max_rows <= 0) max_rows = kDefaultMaxRows;
    SasvqaGitDecoder* d = new SasvqaGitDecoder();
    auto fail = [&](int rc) { git_decoder_destroy(d); return rc; };
#define TRY(expr) do { int _rc = (expr); if (_rc) return fail(_rc); } while (0)
#define TRYCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return fail(SASVQA_ERR_CUDA); } } while (0)
    TRYCUDA(cudaGetDevice(&d->device));
    cudaDeviceProp prop;
    TRYCUDA(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10) {
        set_last_error("sasvqa_b200 needs an sm_100a GPU (B200); found compute capability " + std::to_string(prop.major) +
                       "." + std::to_string(prop.minor));
        return fail(SASVQA_ERR_INVALID);
    }
    d->num_sms = prop.multiProcessorCount;
    d->vocab = vocab;
    d->vocab_pad = (vocab + 255) / 256 * 256;               // the GEMM's N tile
    d->n_layers = n_layers;
    d->max_rows = max_rows;

    float* raw = nullptr;
    TRYCUDA(cudaMalloc(&raw, n_params * sizeof(float)));
    if (cudaMemcpy(raw, params_host, n_params * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("uploading decoder parameters failed");
        return fail(SASVQA_ERR_CUDA);
    }
    const size_t H = kHidden, F = kFfn;
    const size_t n_mat = (size_t)n_layers * (4 * H * H + 2 * F * H) + (size_t)d->vocab_pad * H;
    const size_t n_vec = (size_t)vocab * H + kGitMaxPos * H + 2 * H + H /*zero row*/ +
                         (size_t)n_layers * (3 * H + H + 2 * H + F + H + 2 * H) + d->vocab_pad;
    if (cudaMalloc(&d->arena_bf16, n_mat * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&d->arena_f32, n_vec * sizeof(float)) != cudaSuccess) {
        cudaFree(raw);
        set_last_error("allocating decoder weights failed");
        return fail(SASVQA_ERR_NOMEM);
    }
    cudaMemset(d->arena_bf16, 0, n_mat * sizeof(__nv_bfloat16));
    cudaMemset(d->arena_f32, 0, n_vec * sizeof(float));
    __nv_bfloat16* mp = d->arena_bf16;
    float* vp = d->arena_f32;
    const float* rp = raw;
    auto take_mat = [&](size_t n) {
        __nv_bfloat16* dst = mp;
        git_f32_to_bf16_kernel<<<592, 256>>>(rp, dst, (long long)n);
        mp += n; rp += n;
        return dst;
    };
    auto take_vec = [&](size_t n) {
        float* dst = vp;
        cudaMemcpyAsync(dst, rp, n * sizeof(float), cudaMemcpyDeviceToDevice, 0);
        vp += n; rp += n;
        return dst;
    };
    d->word = take_vec((size_t)vocab * H);
    d->pos = take_vec((size_t)kGitMaxPos * H);
    d->emb_g = take_vec(H);
    d->emb_b = take_vec(H);
    d->zero_row = vp; vp += H;                               // GIT has no token types: the embedding kernel adds this row of zeros
    for (int l = 0; l < n_layers; ++l) {
        GitLayer& Ly = d->L[l];
        Ly.w_qkv = mp; mp += 3 * H * H;                      // HF order query, key, value = the fused q | k | v rows
        Ly.b_qkv = vp; vp += 3 * H;
        for (int j = 0; j < 3; ++j) {
            git_f32_to_bf16_kernel<<<592, 256>>>(rp, Ly.w_qkv + j * H * H, (long long)(H * H));
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
    d->w_head = take_mat((size_t)vocab * H);
    mp += (size_t)(d->vocab_pad - vocab) * H;                // zero rows up to the N tile
    d->b_head = take_vec(vocab);
    vp += d->vocab_pad - vocab;
    cudaError_t ce = cudaDeviceSynchronize();
    cudaFree(raw);
    if (ce != cudaSuccess || (size_t)(rp - raw) != n_params || (size_t)(mp - d->arena_bf16) != n_mat ||
        (size_t)(vp - d->arena_f32) != n_vec) {
        set_last_error(std::string("decoder weight conversion failed: ") + cudaGetErrorString(ce));
        return fail(SASVQA_ERR_CUDA);
    }
    const size_t rows = (size_t)max_rows;
    if (cudaMalloc(&d->x, rows * H * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&d->h, rows * H * sizeof(__nv_bfloat16)) != cudaSuccess ||
        cudaMalloc(&d->big, rows * F * sizeof(__nv_bfloat16)) != cudaSuccess) {
        set_last_error("allocating decoder workspace failed (lower max_rows)");
        return fail(SASVQA_ERR_NOMEM);
    }
    TRYCUDA(cudaMemset(d->x, 0, rows * H * sizeof(float)));
    TRYCUDA(cudaMemset(d->h, 0, rows * H * sizeof(__nv_bfloat16)));
    TRYCUDA(cudaMemset(d->big, 0, rows * F * sizeof(__nv_bfloat16)));
    TRY(make_tensor_map_out(&d->m_out_qkv, d->big, rows, kQkv, 0));
    TRY(make_tensor_map_out(&d->m_out_fc1, d->big, rows, kFfn, 0));
    TRY(make_tensor_map_out(&d->m_out_x, d->x, rows, kHidden, 1));
    TRY(make_tensor_map_bf16_kmajor(&d->m_h, d->h, rows, kHidden, 128));
    TRY(make_tensor_map_bf16_kmajor(&d->m_big_fc, d->big, rows, kFfn, 128));
    TRY(make_tensor_map_bf16_kmajor(&d->m_head, d->w_head, (uint64_t)d->vocab_pad, kHidden, 128));
    for (int l = 0; l < n_layers; ++l) {
        GitLayer& Ly = d->L[l];
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_qkv, Ly.w_qkv, kQkv, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_out, Ly.w_out, kHidden, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_fc1, Ly.w_fc1, kFfn, kHidden, 128));
        TRY(make_tensor_map_bf16_kmajor(&Ly.m_fc2, Ly.w_fc2, kHidden, kFfn, 128));
    }
    TRYCUDA(cudaEventCreateWithFlags(&d->order.tail, cudaEventDisableTiming));
#undef TRY
#undef TRYCUDA
    *out = d;
    return 0;
}

int git_decoder_vocab_padded(const SasvqaGitDecoder* d) { return d ? d->vocab_pad : 0; }

// frames [B, K, 3, 224, 224] fp32 (rows of "sampled_frames") + input_ids [B, L] int32 -> logits of the text rows
// [B, L, vocab_pad] fp32 (columns >= vocab are zero).  hidden_or_null: [n*K*197 + n*L, 768] fp32 stream of the
// FIRST group after n_layers blocks (inspection; visual rows first, then text rows).
// labels_or_null [B, L] int32 + loss_or_null [1]: the mean next-token cross-entropy of modeling.py:208-215 (labels of -100
// are ignored); logits may then be NULL (a per-group scratch holds them).
int git_vqa_logits(SasvqaGitDecoder* d, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* ids, int L,
                   float* logits, int n_layers, float* hidden_or_null, const int32_t* labels_or_null, float* loss_or_null,
                   cudaStream_t s) {
    SASVQA_REQUIRE(d != nullptr && enc != nullptr && B >= 0 && K >= 1, "bad arguments");
    SASVQA_REQUIRE(L >= 1 && L <= kGitMaxPos, "text length must be in [1, 1024] (GIT position table)");
    if (n_layers < 0) n_layers = d->n_layers;                   // -1 = all of them (the logits / loss entry points)
    SASVQA_REQUIRE(n_layers >= 0 && n_layers <= d->n_layers, "bad layer count");
    if (B == 0) return 0;
    const bool want_loss = loss_or_null != nullptr;
    SASVQA_REQUIRE(!want_loss || (labels_or_null != nullptr && L >= 2), "the loss needs labels and at least two text positions");
    SASVQA_REQUIRE(frames != nullptr && ids != nullptr && (logits != nullptr || hidden_or_null != nullptr || want_loss),
                   "null argument");
    const int n_vis = K * kTokens;
    const long long S = (long long)n_vis + L;
    SASVQA_REQUIRE(S <= d->max_rows, "one sample's sequence does not fit the decoder workspace (raise max_rows)");
    SASVQA_REQUIRE(hidden_or_null == nullptr || (long long)B * S <= d->max_rows, "hidden-state inspection needs one pass");
    StreamOrder order(&d->order, s);
    const int group = (int)std::min<long long>(B, d->max_rows / S);
    if ((size_t)(group + 1) * sizeof(int32_t) > d->cu_cap) {
        if (d->cu_dev) SASVQA_CUDA_CHECK(cudaFree(d->cu_dev));
        d->cu_dev = nullptr;
        d->cu_cap = 0;
        SASVQA_CUDA_CHECK(cudaMalloc((void**)&d->cu_dev, (size_t)(group + 1) * sizeof(int32_t)));
        d->cu_cap = (size_t)(group + 1) * sizeof(int32_t);
    }
    if (want_loss) {
        int rc;
        if ((rc = dgrow((void**)&d->row_loss, &d->row_loss_cap, (size_t)B * (L - 1) * sizeof(float)))) return rc;
        if ((rc = dgrow((void**)&d->row_valid, &d->row_valid_cap, (size_t)B * (L - 1) * sizeof(int32_t)))) return rc;
        if (!logits && (rc = dgrow((void**)&d->logits_scratch, &d->logits_cap, (size_t)group * L * d->vocab_pad * sizeof(float))))
            return rc;
    }
    std::vector<int32_t> cu((size_t)group + 1);
    for (int b0 = 0; b0 < B; b0 += group) {
        const int n = std::min(group, B - b0);
        const long long rows_vis = (long long)n * n_vis, M = rows_vis + (long long)n * L;
        int rc;
        // visual rows: encoder + visual_projection write fp32 rows straight into the residual stream
        if ((rc = visual_tokens(enc, nullptr, frames + (size_t)b0 * K * kFrameElems, n * K, 1, d->x, s))) return rc;
        rows_to_bf16_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(d->x), reinterpret_cast<uint4*>(d->h),
                                                    rows_vis * kHidden / 8);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
        // text rows: row of (sample i, position p) = rows_vis + i * L + p
        for (int i = 0; i <= n; ++i) cu[i] = (int32_t)(rows_vis + (long long)i * L);
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(d->cu_dev, cu.data(), ((size_t)n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        if ((rc = launch_embed_layernorm(ids + (size_t)b0 * L, nullptr, d->cu_dev, 0, n, L, d->vocab, 1, d->word, d->pos,
                                         d->zero_row, d->emb_g, d->emb_b, kGitLnEps, d->x, d->h, s)))
            return rc;
        for (int l = 0; l < n_layers; ++l) {
            GitLayer& Ly = d->L[l];
            // Nobody reads the visual rows after the last block when only text logits are wanted: that block still
            // projects q|k|v for every row (text queries need the visual keys and values), then runs attention,
            // out_proj, both LayerNorms and the MLP on the text rows alone -- 1/6 of the visual-row attention and
            // 3/4 of one block's GEMM FLOPs never happen.  The text rows are contiguous (visual-first layout).
            const bool text_only = hidden_or_null == nullptr && l == n_layers - 1 && l == d->n_layers - 1;
            const long long r0 = text_only ? rows_vis : 0, Mr = M - r0;
            CUtensorMap ma_h = d->m_h, ma_big = d->m_big_fc, mo_x = d->m_out_x, mo_fc1 = d->m_out_fc1;
            if (text_only) {
                if ((rc = make_tensor_map_bf16_kmajor(&ma_h, d->h + (size_t)r0 * kHidden, (uint64_t)Mr, kHidden, 128))) return rc;
                if ((rc = make_tensor_map_bf16_kmajor(&ma_big, d->big + (size_t)r0 * kFfn, (uint64_t)Mr, kFfn, 128))) return rc;
                if ((rc = make_tensor_map_out(&mo_x, d->x + (size_t)r0 * kHidden, (uint64_t)Mr, kHidden, 1))) return rc;
                if ((rc = make_tensor_map_out(&mo_fc1, d->big + (size_t)r0 * kFfn, (uint64_t)Mr, kFfn, 0))) return rc;
            }
            float* xr = d->x + (size_t)r0 * kHidden;
            __nv_bfloat16* hr = d->h + (size_t)r0 * kHidden;
            __nv_bfloat16* bigr = d->big + (size_t)r0 * kFfn;
            GemmArgs g{};
            g.A = d->h; g.B = Ly.w_qkv; g.M = (int)M; g.N = kQkv; g.K = kHidden;
            g.epilogue = EPI_BIAS_BF16; g.bias = Ly.b_qkv; g.out_bf16 = d->big;
            if ((rc = dgemm(d, g, &d->m_h, &Ly.m_qkv, &d->m_out_qkv, s))) return rc;
            // flash attention on tcgen05: the visual query tiles (skipped in the text-only block) and one text tile per sample
            if ((rc = launch_attention_git_tcgen05(d->big, d->h, M, n, n_vis, L, text_only ? 0 : 1, d->num_sms, s))) return rc;
            g = GemmArgs{};
            g.A = hr; g.B = Ly.w_out; g.M = (int)Mr; g.N = kHidden; g.K = kHidden;
            g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_out; g.out_f32 = xr;
            if ((rc = dgemm(d, g, &ma_h, &Ly.m_out, &mo_x, s))) return rc;
            if ((rc = launch_layernorm_post(xr, hr, Mr, Ly.ln1_g, Ly.ln1_b, kGitLnEps, s))) return rc;
            g = GemmArgs{};
            g.A = hr; g.B = Ly.w_fc1; g.M = (int)Mr; g.N = kFfn; g.K = kHidden;
            g.epilogue = EPI_BIAS_ERF_GELU_BF16; g.bias = Ly.b_fc1; g.out_bf16 = bigr;
            if ((rc = dgemm(d, g, &ma_h, &Ly.m_fc1, &mo_fc1, s))) return rc;
            g = GemmArgs{};
            g.A = bigr; g.B = Ly.w_fc2; g.M = (int)Mr; g.N = kHidden; g.K = kFfn;
            g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_fc2; g.out_f32 = xr;
            if ((rc = dgemm(d, g, &ma_big, &Ly.m_fc2, &mo_x, s))) return rc;
            if ((rc = launch_layernorm_post(xr, hr, Mr, Ly.ln2_g, Ly.ln2_b, kGitLnEps, s))) return rc;
        }
        if (hidden_or_null && b0 == 0)
            SASVQA_CUDA_CHECK(cudaMemcpyAsync(hidden_or_null, d->x, (size_t)M * kHidden * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (logits || want_loss) {
            // output head on the text rows: they are one contiguous [n*L, 768] bf16 block of h
            const int Mt = n * L;
            float* out = logits ? logits + (size_t)b0 * L * d->vocab_pad : d->logits_scratch;
            SASVQA_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)Mt * d->vocab_pad * sizeof(float), s));
            const __nv_bfloat16* a = d->h + (size_t)rows_vis * kHidden;
            GemmArgs g{};
            g.A = a; g.B = d->w_head; g.M = Mt; g.N = d->vocab_pad; g.K = kHidden;
            g.epilogue = EPI_BIAS_RESID_F32; g.bias = d->b_head; g.out_f32 = out;
            CUtensorMap ma, mo;
            if ((rc = make_tensor_map_bf16_kmajor(&ma, a, (uint64_t)Mt, kHidden, 128))) return rc;
            if ((rc = make_tensor_map_out(&mo, out, (uint64_t)Mt, (uint64_t)d->vocab_pad, 1))) return rc;
            if ((rc = dgemm(d, g, &ma, &d->m_head, &mo, s))) return rc;
            if (want_loss) {
                ce_rows_kernel<<<n * (L - 1), 256, 0, s>>>(out, d->vocab, d->vocab_pad, L, labels_or_null + (size_t)b0 * L,
                                                           d->row_loss + (size_t)b0 * (L - 1), d->row_valid + (size_t)b0 * (L - 1));
                SASVQA_CUDA_CHECK(cudaGetLastError());
                count_launch();
            }
        }
    }
    if (want_loss) {
        ce_mean_kernel<<<1, 256, 0, s>>>(d->row_loss, d->row_valid, B * (L - 1), loss_or_null);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    return 0;
}

namespace {

// one decoder block over rows [0, M) of the workspace; the attention is supplied by the caller
template <class Attention>
int run_block(SasvqaGitDecoder* d, GitLayer& Ly, long long M, Attention&& attention, cudaStream_t s) {
    int rc;
    GemmArgs g{};
    g.A = d->h; g.B = Ly.w_qkv; g.M = (int)M; g.N = kQkv; g.K = kHidden;
    g.epilogue = EPI_BIAS_BF16; g.bias = Ly.b_qkv; g.out_bf16 = d->big;
    if ((rc = dgemm(d, g, &d->m_h, &Ly.m_qkv, &d->m_out_qkv, s))) return rc;
    rc = attention();                                           // reads big (q|k|v), writes h
    if (rc < 0) return 0;                                       // < 0: the caller only wanted the projection
    if (rc) return rc;
    g = GemmArgs{};
    g.A = d->h; g.B = Ly.w_out; g.M = (int)M; g.N = kHidden; g.K = kHidden;
    g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_out; g.out_f32 = d->x;
    if ((rc = dgemm(d, g, &d->m_h, &Ly.m_out, &d->m_out_x, s))) return rc;
    if ((rc = launch_layernorm_post(d->x, d->h, M, Ly.ln1_g, Ly.ln1_b, kGitLnEps, s))) return rc;
    g = GemmArgs{};
    g.A = d->h; g.B = Ly.w_fc1; g.M = (int)M; g.N = kFfn; g.K = kHidden;
    g.epilogue = EPI_BIAS_ERF_GELU_BF16; g.bias = Ly.b_fc1; g.out_bf16 = d->big;
    if ((rc = dgemm(d, g, &d->m_h, &Ly.m_fc1, &d->m_out_fc1, s))) return rc;
    g = GemmArgs{};
    g.A = d->big; g.B = Ly.w_fc2; g.M = (int)M; g.N = kHidden; g.K = kFfn;
    g.epilogue = EPI_BIAS_RESID_F32; g.bias = Ly.b_fc2; g.out_f32 = d->x;
    if ((rc = dgemm(d, g, &d->m_big_fc, &Ly.m_fc2, &d->m_out_x, s))) return rc;
    return launch_layernorm_post(d->x, d->h, M, Ly.ln2_g, Ly.ln2_b, kGitLnEps, s);
}

}  // namespace

// Greedy answer decoding, what the reference's evaluation runs: `self.model.generate(**inputs, max_length=50)`
// (src/modeling/modeling.py:330-333, HF greedy search on MyGitForCausalLM).  prompt [B, L0] int32 (equal lengths, no
// padding), out_ids [B, max_length] int32: the prompt, then one argmax token per step; after eos a sequence emits pad.
// The visual rows never depend on the text (they only see each other), so they are run ONCE per group -- five full
// blocks plus the sixth block's k|v projection -- and every block's visual keys and values are cached; a step then runs the
// text rows alone (all max_length positions at a fixed stride: causal masking makes the not-yet-written ones harmless)
// against the cache.  Every fourth step the finished flags are read back (one stream sync) and the group stops once
// every sequence has emitted eos, as HF does; the tail is pad.
int git_vqa_generate(SasvqaGitDecoder* d, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* prompt, int L0,
                     int max_length, int eos, int pad, int32_t* out_ids, cudaStream_t s) {
    SASVQA_REQUIRE(d != nullptr && enc != nullptr && B >= 0 && K >= 1, "bad arguments");
    SASVQA_REQUIRE(L0 >= 1 && max_length >= L0 && max_length <= kGitMaxPos, "need 1 <= prompt length <= max_length <= 1024");
    if (B == 0) return 0;
    SASVQA_REQUIRE(frames != nullptr && prompt != nullptr && out_ids != nullptr, "null argument");
    const int n_vis = K * kTokens, Lmax = max_length;
    SASVQA_REQUIRE((long long)n_vis <= d->max_rows && (long long)Lmax <= d->max_rows, "one sample does not fit the workspace");
    StreamOrder order(&d->order, s);
    const int group = (int)std::min<long long>(B, std::min<long long>(d->max_rows / n_vis, d->max_rows / Lmax));
    int rc;
    const size_t kv_rows = (size_t)group * n_vis;
    if ((rc = dgrow((void**)&d->kv_cache, &d->kv_cap, (size_t)d->n_layers * kv_rows * 2 * kHidden * sizeof(__nv_bfloat16)))) return rc;
    if ((rc = dgrow((void**)&d->gen_ids, &d->gen_ids_cap, (size_t)group * Lmax * sizeof(int32_t)))) return rc;
    if ((rc = dgrow((void**)&d->gen_done, &d->gen_done_cap, (size_t)group * sizeof(int32_t)))) return rc;
    if ((rc = dgrow((void**)&d->last_h, &d->last_h_cap, (size_t)std::max(group, 1) * kHidden * sizeof(__nv_bfloat16)))) return rc;
    if ((rc = dgrow((void**)&d->logits_scratch, &d->logits_cap, (size_t)group * d->vocab_pad * sizeof(float)))) return rc;
    if ((size_t)(group + 1) * sizeof(int32_t) > d->cu_cap) {
        if (d->cu_dev) SASVQA_CUDA_CHECK(cudaFree(d->cu_dev));
        d->cu_dev = nullptr;
        d->cu_cap = 0;
        SASVQA_CUDA_CHECK(cudaMalloc((void**)&d->cu_dev, (size_t)(group + 1) * sizeof(int32_t)));
        d->cu_cap = (size_t)(group + 1) * sizeof(int32_t);
    }
    std::vector<int32_t> cu((size_t)group + 1), done_host;
    for (int b0 = 0; b0 < B; b0 += group) {
        const int n = std::min(group, B - b0);
        const long long rows_vis = (long long)n * n_vis, rows_txt = (long long)n * Lmax;
        // ---- prefill: the visual rows through the blocks, k|v of every block into the cache
        if ((rc = visual_tokens(enc, nullptr, frames + (size_t)b0 * K * kFrameElems, n * K, 1, d->x, s))) return rc;
        rows_to_bf16_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(d->x), reinterpret_cast<uint4*>(d->h),
                                                    rows_vis * kHidden / 8);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
        for (int l = 0; l < d->n_layers; ++l) {
            __nv_bfloat16* cache = d->kv_cache + (size_t)l * kv_rows * 2 * kHidden;
            const bool last = l == d->n_layers - 1;
            auto att = [&]() -> int {
                SASVQA_CUDA_CHECK(cudaMemcpy2DAsync(cache, 2 * kHidden * sizeof(__nv_bfloat16), d->big + kHidden,
                                                    kQkv * sizeof(__nv_bfloat16), 2 * kHidden * sizeof(__nv_bfloat16),
                                                    (size_t)rows_vis, cudaMemcpyDeviceToDevice, s));
                if (last) return -1;                            // nobody reads the visual rows after the last block
                return launch_attention_git_tcgen05(d->big, d->h, rows_vis, n, n_vis, 0, 1, d->num_sms, s);
            };
            if ((rc = run_block(d, d->L[l], rows_vis, att, s))) return rc;
        }
        // ---- decode: text rows only, sample i at rows [i * Lmax, (i + 1) * Lmax)
        init_generate_kernel<<<(unsigned)((rows_txt + 255) / 256), 256, 0, s>>>(prompt + (size_t)b0 * L0, L0, d->gen_ids, Lmax, pad,
                                                                                 d->gen_done, n);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
        for (int i = 0; i <= n; ++i) cu[i] = (int32_t)((long long)i * Lmax);
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(d->cu_dev, cu.data(), ((size_t)n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        for (int pos = L0; pos < Lmax; ++pos) {                 // the token at `pos` comes from the row at pos - 1
            if ((rc = launch_embed_layernorm(d->gen_ids, nullptr, d->cu_dev, 0, n, Lmax, d->vocab, 1, d->word, d->pos, d->zero_row,
                                             d->emb_g, d->emb_b, kGitLnEps, d->x, d->h, s)))
                return rc;
            for (int l = 0; l < d->n_layers; ++l) {
                const __nv_bfloat16* cache = d->kv_cache + (size_t)l * kv_rows * 2 * kHidden;
                auto att = [&]() -> int { return launch_attention_git(d->big, d->h, n, n_vis, Lmax, 1, s, cache); };
                if ((rc = run_block(d, d->L[l], rows_txt, att, s))) return rc;
            }
            SASVQA_CUDA_CHECK(cudaMemcpy2DAsync(d->last_h, kHidden * sizeof(__nv_bfloat16), d->h + (size_t)(pos - 1) * kHidden,
                                                (size_t)Lmax * kHidden * sizeof(__nv_bfloat16), kHidden * sizeof(__nv_bfloat16),
                                                (size_t)n, cudaMemcpyDeviceToDevice, s));
            SASVQA_CUDA_CHECK(cudaMemsetAsync(d->logits_scratch, 0, (size_t)n * d->vocab_pad * sizeof(float), s));
            GemmArgs g{};
            g.A = d->last_h; g.B = d->w_head; g.M = n; g.N = d->vocab_pad; g.K = kHidden;
            g.epilogue = EPI_BIAS_RESID_F32; g.bias = d->b_head; g.out_f32 = d->logits_scratch;
            CUtensorMap ma, mo;
            if ((rc = make_tensor_map_bf16_kmajor(&ma, d->last_h, (uint64_t)n, kHidden, 128))) return rc;
            if ((rc = make_tensor_map_out(&mo, d->logits_scratch, (uint64_t)n, (uint64_t)d->vocab_pad, 1))) return rc;
            if ((rc = dgemm(d, g, &ma, &d->m_head, &mo, s))) return rc;
            greedy_next_kernel<<<n, 256, 0, s>>>(d->logits_scratch, d->vocab, d->vocab_pad, d->gen_ids, Lmax, pos, eos, pad,
                                                 d->gen_done);
            SASVQA_CUDA_CHECK(cudaGetLastError());
            count_launch();
            // HF stops as soon as every sequence has produced eos; VQA answers are a few tokens, so look every 4th step
            // (one small D2H + stream sync) instead of always running to max_length -- the tail is pad already
            if ((pos - L0) % 4 == 3 && pos + 1 < Lmax) {
                done_host.resize((size_t)n);
                SASVQA_CUDA_CHECK(cudaMemcpyAsync(done_host.data(), d->gen_done, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
                SASVQA_CUDA_CHECK(cudaStreamSynchronize(s));
                if (std::all_of(done_host.begin(), done_host.end(), [](int32_t v) { return v != 0; })) break;
            }
        }
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(out_ids + (size_t)b0 * Lmax, d->gen_ids, (size_t)rows_txt * sizeof(int32_t),
                                          cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

}  // namespace sasvqa
