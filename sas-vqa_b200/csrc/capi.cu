// extern "C" surface of libsasvqa_b200.so (declared in include/sasvqa.h).
#include <atomic>
#include <exception>
#include <new>

#include "../../include/sasvqa.h"
#include "common.cuh"

namespace sasvqa {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int profile_enable(SasvqaEncoder*, int);
int profile_read(SasvqaEncoder*, double*, int64_t*, int);

int encoder_create(const float*, uint64_t, int, SasvqaEncoder**);
void encoder_destroy(SasvqaEncoder*);
int encoder_chunk_frames(const SasvqaEncoder*);
int encoder_fwd(SasvqaEncoder*, const __nv_bfloat16*, int, float*, cudaStream_t);
int encoder_fwd_hidden(SasvqaEncoder*, const __nv_bfloat16*, int, int, float*, cudaStream_t);
int mdf_sample_device(SasvqaEncoder*, const uint8_t*, const float*, int, int, int, int, int, int, int32_t*, int32_t*,
                      float*, float*, float*, cudaStream_t);
int mdf_sample_host(SasvqaEncoder*, const uint8_t*, int, int, int, int, int, int, int32_t*, int32_t*, float*);
int mif_sample_host(SasvqaEncoder*, const uint8_t*, int, int, int, int, const float*, int, int, int32_t*, float*);
int mdf_sample_ragged_host(SasvqaEncoder*, const uint8_t*, int, const int32_t*, int, int, int, int, int32_t*, int32_t*, float*);
int mdf_sample_ragged_device(SasvqaEncoder*, const uint8_t*, const float*, int, const int32_t*, int, int, int, int, int32_t*,
                             int32_t*, float*, float*, float*, cudaStream_t);
int encoder_set_projection(SasvqaEncoder*, const float*, const float*, const float*, const float*);
int visual_tokens(SasvqaEncoder*, const uint8_t*, const float*, int, int, float*, cudaStream_t);
int mif_sample_device(SasvqaEncoder*, const uint8_t*, const float*, int, int, int, int, const float*, int, int, int32_t*,
                      float*, float*, float*, cudaStream_t);

int video_probe(const uint8_t*, uint64_t, int, int, int32_t*);
int video_decode(const uint8_t*, uint64_t, int, int, uint8_t*, int, int, int, int, int32_t*, cudaStream_t);
uint64_t git_decoder_num_params(int, int);
int git_decoder_create(const float*, uint64_t, int, int, int, SasvqaGitDecoder**);
void git_decoder_destroy(SasvqaGitDecoder*);
int git_decoder_vocab_padded(const SasvqaGitDecoder*);
int git_vqa_logits(SasvqaGitDecoder*, SasvqaEncoder*, const float*, int, int, const int32_t*, int, float*, int, float*,
                   const int32_t*, float*, cudaStream_t);
int git_vqa_generate(SasvqaGitDecoder*, SasvqaEncoder*, const float*, int, int, const int32_t*, int, int, int, int, int32_t*,
                     cudaStream_t);
uint64_t scorer_num_params(int, int);
int scorer_create(const float*, uint64_t, int, int, int, SasvqaScorer**);
void scorer_destroy(SasvqaScorer*);
int scorer_max_tokens(const SasvqaScorer*);
int scorer_logits(SasvqaScorer*, const int32_t*, const int32_t*, const int32_t*, int, int, float*, int, float*, cudaStream_t);
int scorer_logits_host(SasvqaScorer*, const int64_t*, const int64_t*, const int64_t*, int, int, float*);
int mif_select_captions_host(SasvqaScorer*, const int64_t*, const int64_t*, const int64_t*, int, int, int, int, int, int,
                             int32_t*, float*);
int scorer_profile_enable(SasvqaScorer*, int);
int scorer_profile_read(SasvqaScorer*, double*, int64_t*, int);

}  // namespace sasvqa

using namespace sasvqa;

// No C++ exception crosses the C ABI: host-side allocations (std::vector / std::string in the pipelines, `new` of a handle)
// that fail come back as SASVQA_ERR_NOMEM, anything else unexpected as SASVQA_ERR_INVALID.
template <class F>
static int guarded(F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        try { set_last_error("out of host memory"); } catch (...) {}
        return SASVQA_ERR_NOMEM;
    } catch (const std::exception& ex) {
        try { set_last_error(std::string("unexpected C++ exception: ") + ex.what()); } catch (...) {}
        return SASVQA_ERR_INVALID;
    } catch (...) {
        return SASVQA_ERR_INVALID;
    }
}

#define S(stream) (reinterpret_cast<cudaStream_t>(stream))
#define BF(p) (reinterpret_cast<__nv_bfloat16*>(p))
#define CBF(p) (reinterpret_cast<const __nv_bfloat16*>(p))

extern "C" {

int sasvqa_abi_version(void) { return 1; }
const char* sasvqa_last_error(void) { return g_last_error.c_str(); }
int64_t sasvqa_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
int sasvqa_profile_enable(SasvqaEncoder* enc, int on) { return guarded([&]() -> int { return profile_enable(enc, on); }); }
int sasvqa_profile_read(SasvqaEncoder* enc, double* ms, int64_t* launches, int n_kinds) {
    return guarded([&]() -> int {
        return profile_read(enc, ms, launches, n_kinds);
    });
}

int sasvqa_encoder_create(const float* params_host, uint64_t n_params, int chunk_frames, SasvqaEncoder** out) {
    return guarded([&]() -> int {
        return encoder_create(params_host, n_params, chunk_frames, out);
    });
}
void sasvqa_encoder_destroy(SasvqaEncoder* enc) { encoder_destroy(enc); }
int sasvqa_encoder_chunk_frames(const SasvqaEncoder* enc) { return guarded([&]() -> int { return encoder_chunk_frames(enc); }); }

int sasvqa_preprocess_u8(const uint8_t* frames, int n_frames, uint16_t* patches, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(n_frames >= 0 && (n_frames == 0 || (frames && patches)), "bad arguments");
        return launch_preprocess_u8(frames, n_frames, BF(patches), S(stream));
    });
}
int sasvqa_patchify_f32(const float* frames, int n_frames, uint16_t* patches, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(n_frames >= 0 && (n_frames == 0 || (frames && patches)), "bad arguments");
        return launch_patchify_f32(frames, n_frames, BF(patches), S(stream));
    });
}

int sasvqa_encoder_fwd(SasvqaEncoder* enc, const uint16_t* patches, int n_frames, float* feats, void* stream) {
    return guarded([&]() -> int {
        return encoder_fwd(enc, CBF(patches), n_frames, feats, S(stream));
    });
}
int sasvqa_encoder_fwd_hidden(SasvqaEncoder* enc, const uint16_t* patches, int n_frames, int n_layers, float* hidden,
                              void* stream) {
    return guarded([&]() -> int {
        return encoder_fwd_hidden(enc, CBF(patches), n_frames, n_layers, hidden, S(stream));
    });
}

int sasvqa_mdf_scores(const float* feats, int B, int T, int W, float* lcl_avg, float* gram_or_null, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0 && T >= 0, "bad B/T");
        SASVQA_REQUIRE(B == 0 || T == 0 || (feats && lcl_avg), "null argument");
        return launch_mdf_scores(feats, B, T, W, lcl_avg, gram_or_null, S(stream));
    });
}
int sasvqa_mdf_select(const float* lcl_avg, int B, int T, int K, int W, int32_t* idx, int32_t* status, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0, "bad B");
        SASVQA_REQUIRE(B == 0 || (lcl_avg && idx && status), "null argument");
        return launch_mdf_select(lcl_avg, B, T, K, W, idx, status, S(stream));
    });
}
int sasvqa_topk_strided(const float* scores, int B, int T, int ds_rate, int K, int32_t* idx, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0 && T >= 0 && K >= 0, "bad B/T/K");
        SASVQA_REQUIRE(B == 0 || K == 0 || (scores && idx), "null argument");
        return launch_topk_strided(scores, B, T, ds_rate, K, idx, nullptr, S(stream));
    });
}

int sasvqa_encoder_set_projection(SasvqaEncoder* enc, const float* w, const float* b, const float* ln_g, const float* ln_b) {
    return guarded([&]() -> int {
        return encoder_set_projection(enc, w, b, ln_g, ln_b);
    });
}
int sasvqa_visual_tokens_f32(SasvqaEncoder* enc, const float* frames, int n_frames, int project, float* tokens, void* stream) {
    return guarded([&]() -> int {
        return visual_tokens(enc, nullptr, frames, n_frames, project, tokens, S(stream));
    });
}
int sasvqa_visual_tokens_u8(SasvqaEncoder* enc, const uint8_t* frames, int n_frames, int project, float* tokens, void* stream) {
    return guarded([&]() -> int {
        return visual_tokens(enc, frames, nullptr, n_frames, project, tokens, S(stream));
    });
}

int sasvqa_mif_scores(const float* feats, const float* q, int B, int T, float* scores, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0 && T >= 0, "bad B/T");
        SASVQA_REQUIRE(B == 0 || T == 0 || (feats && q && scores), "null argument");
        return launch_mif_scores(feats, q, B, T, scores, S(stream));
    });
}
int sasvqa_mif_sample_u8_hw(SasvqaEncoder* enc, const uint8_t* clips, int B, int T, int H, int Wd, const float* q, int K,
                            int ds_rate, int32_t* idx, float* scores, float* feats, float* sampled, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || clips != nullptr, "null clips");
        return mif_sample_device(enc, clips, nullptr, B, T, H, Wd, q, K, ds_rate, idx, scores, feats, sampled, S(stream));
    });
}

int sasvqa_mif_sample_host_hw(SasvqaEncoder* enc, const uint8_t* clips_host, int B, int T, int H, int Wd, const float* q_host,
                              int K, int ds_rate, int32_t* idx_host, float* sampled_host) {
    return guarded([&]() -> int {
        return mif_sample_host(enc, clips_host, B, T, H, Wd, q_host, K, ds_rate, idx_host, sampled_host);
    });
}

int sasvqa_gather_frames_u8(const uint8_t* clips, const int32_t* idx, int B, int T, int K, float* out, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0 && T >= 0 && K >= 0, "bad B/T/K");
        return launch_gather_u8(clips, idx, B, T, K, out, S(stream));
    });
}
int sasvqa_gather_frames_f32(const float* frames, const int32_t* idx, int B, int T, int K, int64_t row_elems,
                             float* out, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B >= 0 && T >= 0 && K >= 0 && row_elems >= 0, "bad B/T/K/row_elems");
        return launch_gather_f32(frames, idx, B, T, K, row_elems, out, S(stream));
    });
}

int sasvqa_mdf_sample_u8(SasvqaEncoder* enc, const uint8_t* clips, int B, int T, int K, int W, int32_t* idx,
                         int32_t* status, float* lcl_avg, float* feats, float* sampled, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || T == 0 || clips != nullptr, "null clips");
        return mdf_sample_device(enc, clips, nullptr, B, T, kImg, kImg, K, W, idx, status, lcl_avg, feats, sampled, S(stream));
    });
}
int sasvqa_mdf_sample_u8_hw(SasvqaEncoder* enc, const uint8_t* clips, int B, int T, int H, int Wd, int K, int W, int32_t* idx,
                            int32_t* status, float* lcl_avg, float* feats, float* sampled, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || T == 0 || clips != nullptr, "null clips");
        return mdf_sample_device(enc, clips, nullptr, B, T, H, Wd, K, W, idx, status, lcl_avg, feats, sampled, S(stream));
    });
}
int sasvqa_mdf_sample_ragged_u8(SasvqaEncoder* enc, const uint8_t* frames, int B, const int32_t* clip_offsets_host, int H,
                                int Wd, int K, int W, int32_t* idx, int32_t* status, float* lcl_avg, float* feats,
                                float* sampled, void* stream) {
    return guarded([&]() -> int {
        return mdf_sample_ragged_device(enc, frames, nullptr, B, clip_offsets_host, H, Wd, K, W, idx, status, lcl_avg, feats, sampled,
                                        S(stream));
    });
}
int sasvqa_mdf_sample_ragged_host(SasvqaEncoder* enc, const uint8_t* frames_host, int B, const int32_t* clip_offsets_host, int H,
                                  int Wd, int K, int W, int32_t* idx_host, int32_t* status_host, float* sampled_host) {
    return guarded([&]() -> int {
        return mdf_sample_ragged_host(enc, frames_host, B, clip_offsets_host, H, Wd, K, W, idx_host, status_host, sampled_host);
    });
}
int sasvqa_resize_crop_u8(const uint8_t* frames, int n_frames, int H, int Wd, uint8_t* out, void* stream) {
    return guarded([&]() -> int {
        return launch_resize_crop_u8(frames, n_frames, H, Wd, nullptr, 0, n_frames, out, S(stream));
    });
}
int sasvqa_mdf_sample_f32(SasvqaEncoder* enc, const float* clips, int B, int T, int K, int W, int32_t* idx,
                          int32_t* status, float* lcl_avg, float* feats, float* sampled, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || T == 0 || clips != nullptr, "null clips");
        return mdf_sample_device(enc, nullptr, clips, B, T, kImg, kImg, K, W, idx, status, lcl_avg, feats, sampled, S(stream));
    });
}
int sasvqa_mdf_sample_host(SasvqaEncoder* enc, const uint8_t* clips_host, int B, int T, int K, int W, int32_t* idx_host,
                           int32_t* status_host, float* sampled_host) {
    return guarded([&]() -> int {
        return mdf_sample_host(enc, clips_host, B, T, kImg, kImg, K, W, idx_host, status_host, sampled_host);
    });
}
int sasvqa_mdf_sample_host_hw(SasvqaEncoder* enc, const uint8_t* clips_host, int B, int T, int H, int Wd, int K, int W,
                              int32_t* idx_host, int32_t* status_host, float* sampled_host) {
    return guarded([&]() -> int {
        return mdf_sample_host(enc, clips_host, B, T, H, Wd, K, W, idx_host, status_host, sampled_host);
    });
}

uint64_t sasvqa_scorer_num_params(int vocab_size, int num_labels) {
    return vocab_size >= 1 && num_labels >= 1 ? scorer_num_params(vocab_size, num_labels) : 0;
}
int sasvqa_scorer_create(const float* params_host, uint64_t n_params, int vocab_size, int num_labels, int max_tokens,
                         SasvqaScorer** out) {
    return guarded([&]() -> int {
        return scorer_create(params_host, n_params, vocab_size, num_labels, max_tokens, out);
    });
}
void sasvqa_scorer_destroy(SasvqaScorer* scorer) { scorer_destroy(scorer); }
int sasvqa_scorer_max_tokens(const SasvqaScorer* scorer) { return guarded([&]() -> int { return scorer_max_tokens(scorer); }); }
int sasvqa_scorer_logits(SasvqaScorer* scorer, const int32_t* ids, const int32_t* type_ids, const int32_t* lengths_host, int N,
                         int L, float* logits, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(N == 0 || logits != nullptr, "null logits");
        return scorer_logits(scorer, ids, type_ids, lengths_host, N, L, logits, kLayers, nullptr, S(stream));
    });
}
int sasvqa_scorer_hidden(SasvqaScorer* scorer, const int32_t* ids, const int32_t* type_ids, const int32_t* lengths_host, int N,
                         int L, int n_layers, float* hidden, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(N == 0 || hidden != nullptr, "null hidden");
        return scorer_logits(scorer, ids, type_ids, lengths_host, N, L, nullptr, n_layers, hidden, S(stream));
    });
}
int sasvqa_scorer_logits_host(SasvqaScorer* scorer, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int N, int L,
                              float* logits_host) {
    return guarded([&]() -> int {
        return scorer_logits_host(scorer, ids, type_ids, mask, N, L, logits_host);
    });
}
int sasvqa_mif_select_captions_host(SasvqaScorer* scorer, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int G,
                                    int T, int L, int K, int ds_rate, int label, int32_t* idx_host, float* scores_host) {
    return guarded([&]() -> int {
        return mif_select_captions_host(scorer, ids, type_ids, mask, G, T, L, K, ds_rate, label, idx_host, scores_host);
    });
}
int sasvqa_scorer_profile_enable(SasvqaScorer* scorer, int on) { return guarded([&]() -> int { return scorer_profile_enable(scorer, on); }); }
int sasvqa_scorer_profile_read(SasvqaScorer* scorer, double* ms, int64_t* scopes, int n_kinds) {
    return guarded([&]() -> int {
        return scorer_profile_read(scorer, ms, scopes, n_kinds);
    });
}

uint64_t sasvqa_git_decoder_num_params(int vocab_size, int n_layers) {
    return vocab_size >= 1 && n_layers >= 1 ? git_decoder_num_params(vocab_size, n_layers) : 0;
}
int sasvqa_git_decoder_create(const float* params_host, uint64_t n_params, int vocab_size, int n_layers, int max_rows,
                              SasvqaGitDecoder** out) {
    return guarded([&]() -> int {
        return git_decoder_create(params_host, n_params, vocab_size, n_layers, max_rows, out);
    });
}
void sasvqa_git_decoder_destroy(SasvqaGitDecoder* dec) { git_decoder_destroy(dec); }
int sasvqa_git_decoder_vocab_padded(const SasvqaGitDecoder* dec) { return guarded([&]() -> int { return git_decoder_vocab_padded(dec); }); }
int sasvqa_git_vqa_logits_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* ids,
                              int L, float* logits, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || logits != nullptr, "null logits");
        return git_vqa_logits(dec, enc, frames, B, K, ids, L, logits, -1, nullptr, nullptr, nullptr, S(stream));
    });
}
int sasvqa_git_vqa_hidden_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* ids,
                              int L, int n_layers, float* hidden, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(B == 0 || hidden != nullptr, "null hidden");
        SASVQA_REQUIRE(n_layers >= 0, "bad layer count");
        return git_vqa_logits(dec, enc, frames, B, K, ids, L, nullptr, n_layers, hidden, nullptr, nullptr, S(stream));
    });
}
int sasvqa_git_vqa_loss_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* ids,
                            const int32_t* labels, int L, float* loss, float* logits_or_null, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(loss != nullptr && labels != nullptr, "null loss / labels");
        SASVQA_REQUIRE(B >= 1, "the loss of an empty batch is undefined");
        return git_vqa_logits(dec, enc, frames, B, K, ids, L, logits_or_null, -1, nullptr, labels, loss, S(stream));
    });
}
int sasvqa_git_vqa_generate_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames, int B, int K, const int32_t* prompt,
                                int L0, int max_length, int eos_token_id, int pad_token_id, int32_t* out_ids, void* stream) {
    return guarded([&]() -> int {
        return git_vqa_generate(dec, enc, frames, B, K, prompt, L0, max_length, eos_token_id, pad_token_id, out_ids, S(stream));
    });
}
int sasvqa_test_attention_git(const uint16_t* qkv, int n_samples, int n_vis, int L, uint16_t* out, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(n_samples == 0 || (qkv && out), "null argument");
        SASVQA_REQUIRE(n_vis >= 1 && L >= 0, "the visual prefix must hold at least one token");
        if (n_samples == 0) return 0;
        int dev = 0, sms = 148;                                     // the decoder's own sequence: visual rows, then text rows
        SASVQA_CUDA_CHECK(cudaGetDevice(&dev));
        SASVQA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        return launch_attention_git_tcgen05(CBF(qkv), BF(out), (long long)n_samples * (n_vis + L), n_samples, n_vis, L, 1, sms,
                                            S(stream));
    });
}
int sasvqa_test_attention_varlen(const uint16_t* qkv, const int32_t* cu_seqlens, int n_seqs, int max_len, uint16_t* out,
                                 void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(n_seqs == 0 || (qkv && cu_seqlens && out), "null argument");
        return launch_attention_varlen(CBF(qkv), BF(out), cu_seqlens, 0, n_seqs, max_len, S(stream));
    });
}
int sasvqa_video_probe(const uint8_t* bitstream_host, uint64_t n_bytes, int codec, int intv, int32_t* info_out) {
    return guarded([&]() -> int {
        return video_probe(bitstream_host, n_bytes, codec, intv, info_out);
    });
}
int sasvqa_video_decode(const uint8_t* bitstream_host, uint64_t n_bytes, int codec, int intv, uint8_t* frames_dev, int capacity_frames,
                        int H, int W, int format, int32_t* n_frames_out, void* stream) {
    return guarded([&]() -> int {
        return video_decode(bitstream_host, n_bytes, codec, intv, frames_dev, capacity_frames, H, W, format, n_frames_out, S(stream));
    });
}
int sasvqa_test_gemm(const uint16_t* a, const uint16_t* b, int M, int N, int K, int mode, const float* bias_or_pos,
                     uint16_t* out_bf16, float* out_f32, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(a && b && mode >= 0 && mode <= 4, "bad arguments");
        GemmArgs g{};
        g.A = CBF(a); g.B = CBF(b); g.M = M; g.N = N; g.K = K; g.epilogue = mode;
        g.bias = bias_or_pos; g.pos = bias_or_pos; g.out_bf16 = BF(out_bf16); g.out_f32 = out_f32;
        CUtensorMap ma, mb;
        int rc = make_tensor_map_bf16_kmajor(&ma, a, (uint64_t)M, (uint64_t)K, 128);
        if (rc) return rc;
        if ((rc = make_tensor_map_bf16_kmajor(&mb, b, (uint64_t)N, (uint64_t)K, 128))) return rc;
        CUtensorMap mo = ma;
        if (mode == EPI_BIAS_BF16 || mode == EPI_BIAS_GELU_BF16 || mode == EPI_BIAS_ERF_GELU_BF16) rc = make_tensor_map_out(&mo, out_bf16, (uint64_t)M, (uint64_t)N, 0);
        else if (mode == EPI_BIAS_RESID_F32) rc = make_tensor_map_out(&mo, out_f32, (uint64_t)M, (uint64_t)N, 1);
        if (rc) return rc;
        int dev = 0, sms = 148;
        SASVQA_CUDA_CHECK(cudaGetDevice(&dev));
        SASVQA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        return launch_gemm_tcgen05(g, &ma, &mb, &mo, sms, S(stream));
    });
}
int sasvqa_test_attention(const uint16_t* qkv, int n_frames, uint16_t* out, int trace_variant, void* stream) {
    return guarded([&]() -> int {
        SASVQA_REQUIRE(trace_variant >= 0, "bad trace variant");
        CUtensorMap mq, mkv, mo;
        int rc = make_attention_maps(&mq, &mkv, &mo, qkv, out, (uint64_t)n_frames * kTokens);
        if (rc) return rc;
        int dev = 0, sms = 148;
        SASVQA_CUDA_CHECK(cudaGetDevice(&dev));
        SASVQA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        return launch_attention_tcgen05(&mq, &mkv, &mo, BF(out), n_frames, sms, S(stream), trace_variant);
    });
}
int sasvqa_test_layernorm(const float* x, int rows, const float* gamma, const float* beta, uint16_t* out, void* stream) {
    return guarded([&]() -> int {
        return launch_layernorm_bf16(x, BF(out), rows, gamma, beta, S(stream));
    });
}

}  // extern "C"
