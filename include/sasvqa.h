/* sasvqa.h -- C ABI of libsasvqa_b200.so: the B200-native frame-sampling hot path of SAS-VQA.
 *
 * The reference (Clement25/SAS-VQA) has no FFI: its seam is the Python call
 *     exted_frms = sample_representative_frames(video_frms, model, args.K, args.W, debug_counter)
 * (src/preprocessing/extract_features.py:90 -> src/preprocessing/datautils/utils.py:31-94) and
 * its MIF sibling  inds = scores[::ds_rate].topk(args.K)[1]  (src/preprocessing/gen_sample.py:87-88).
 * These entry points are what a ctypes binding of that seam calls (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 (SASVQA_OK) or an error code and never throws;
 * sasvqa_last_error() returns the message of the calling thread's last failure.  Pointers named
 * *_dev are CUDA device pointers on the current device, *_host are host pointers (pinned memory
 * makes the copies asynchronous).  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream); device-pointer entry points are asynchronous on it, host-pointer entry points return
 * after the results are in host memory.  Outputs are caller-allocated.  bf16 buffers are passed
 * as uint16_t*.  Threading: a handle (SasvqaEncoder / SasvqaScorer) owns one workspace and serialises its own work --
 * call it from one thread at a time (the reference calls its sampler from the main thread only,
 * extract_features.py:80-97); different handles, and the handle-free entry points, may be used concurrently.
 * Frames are 224x224 (ViT-B/16 input of the reference's GitVisionModel) unless an entry
 * point takes H and W (decoded frames of any size: K0 resizes and crops them as the image processor does).
 */
#ifndef SASVQA_H
#define SASVQA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SASVQA_OK 0
#define SASVQA_ERR_INVALID 1     /* bad argument / unsupported shape */
#define SASVQA_ERR_CUDA 2        /* a CUDA call failed */
#define SASVQA_ERR_NOMEM 3

/* per-clip status written by the selection stage */
#define SASVQA_STATUS_OK 0         /* K picks from the greedy spacing rule            (utils.py:63-88) */
#define SASVQA_STATUS_FALLBACK 1   /* fewer than K: plain top-K taken, 'Failure' += 1 (utils.py:91-93) */
#define SASVQA_STATUS_EMPTY 2      /* T == 0: zero frames returned, 'Zeros' += 1      (utils.py:50-52) */
#define SASVQA_STATUS_TOO_FEW 3    /* fallback with T < K: the reference raises RuntimeError here      */

#define SASVQA_NUM_ENCODER_PARAMS 85799424ull /* fp32 values in a GitVisionModel ViT-B/16 state dict */

typedef struct SasvqaEncoder SasvqaEncoder; /* frozen frame encoder: bf16 weights + workspace on one GPU */

int sasvqa_abi_version(void);
const char* sasvqa_last_error(void);

/* ---- encoder handle ------------------------------------------------------------------------
 * Replaces `GitVisionModel.from_pretrained(...)` + `.eval().cuda()` (extract_features.py:145,45-47).
 * params_host: the model's state_dict flattened to fp32 in HF key order: class_embedding,
 * patch_embedding.weight, position_embedding.weight, pre_layrnorm.{weight,bias}, then per layer
 * {k,v,q,out}_proj.{weight,bias}, layer_norm1.{weight,bias}, mlp.fc1.{weight,bias},
 * mlp.fc2.{weight,bias}, layer_norm2.{weight,bias}, then post_layernorm.{weight,bias}.
 * chunk_frames: frames encoded per pass (workspace ~2.1 MB per frame); <= 0 picks the default. */
int sasvqa_encoder_create(const float* params_host, uint64_t n_params, int chunk_frames, SasvqaEncoder** out);
void sasvqa_encoder_destroy(SasvqaEncoder* enc);
int sasvqa_encoder_chunk_frames(const SasvqaEncoder* enc);

/* ---- K0: resize + centre crop (frames that are not 224x224) -----------------------------------
 * uint8 HWC [n, H, W, 3] -> uint8 HWC [n, 224, 224, 3]: shortest edge to 224 with the anti-aliased bicubic
 * filter, then centre crop -- bit-exact with the host image processor's resize/crop
 * (prefetch_loader.py:74-75 -> HF CLIPImageProcessor -> torchvision/ATen uint8 resampler). */
int sasvqa_resize_crop_u8(const uint8_t* frames_hwc_dev, int n_frames, int H, int W, uint8_t* out_224_dev, void* stream);

/* ---- K1: preprocessing ---------------------------------------------------------------------
 * uint8 HWC frames -> (x/255 - mean)/std -> bf16 patch matrix [n*196, 768], column = c*256+iy*16+ix.
 * Replaces the host image processor (prefetch_loader.py:74-75; 224x224 input) + the conv's im2col. */
int sasvqa_preprocess_u8(const uint8_t* frames_hwc_dev, int n_frames, uint16_t* patches_bf16_dev, void* stream);
/* same layout from already normalised fp32 CHW frames (the type utils.py:31 receives) */
int sasvqa_patchify_f32(const float* frames_chw_dev, int n_frames, uint16_t* patches_bf16_dev, void* stream);

/* ---- K2 + K3a: encoder forward + pooling -----------------------------------------------------
 * patches -> ViT-B/16 -> post-LN -> mean over 197 tokens -> L2 normalise: feats [n, 768] fp32 unit rows.
 * Replaces utils.py:39-48 (model(chunk), .last_hidden_state.mean(1), normalize). */
int sasvqa_encoder_fwd(SasvqaEncoder* enc, const uint16_t* patches_bf16_dev, int n_frames, float* feats_dev,
                       void* stream);
/* inspection: fp32 hidden state [n, 197, 768] after `n_layers` blocks (before post_layernorm); n <= chunk */
int sasvqa_encoder_fwd_hidden(SasvqaEncoder* enc, const uint16_t* patches_bf16_dev, int n_frames, int n_layers,
                              float* hidden_dev, void* stream);

/* ---- K3b + K4a: windowed mean cosine similarity (utils.py:55-61) -----------------------------
 * feats [B, T, 768] -> lcl_avg [B, T]; W >= 0 (resolve W == -1 to T / 20 first, utils.py:32-33).
 * gram_or_null: optional [B, T, T] full Gram matrix (the reference materialises it; we only need the band). */
int sasvqa_mdf_scores(const float* feats_dev, int B, int T, int W, float* lcl_avg_dev, float* gram_or_null_dev,
                      void* stream);

/* ---- K4b: greedy selection + top-K fallback (utils.py:63-93) ---------------------------------
 * lcl_avg [B, T] -> idx [B, K] int32 in importance order, status [B] (SASVQA_STATUS_*). */
int sasvqa_mdf_select(const float* lcl_avg_dev, int B, int T, int K, int W, int32_t* idx_dev, int32_t* status_dev,
                      void* stream);

/* ---- K4c: MIF strided top-K (gen_sample.py:87-88) --------------------------------------------
 * idx[b, :] = ds_rate * topk(scores[b, ::ds_rate], K), best first. */
int sasvqa_topk_strided(const float* scores_dev, int B, int T, int ds_rate, int K, int32_t* idx_dev, void* stream);

/* ---- MIF relevance in embedding space (BASELINE config 3) --------------------------------------
 * scores[b, t] = <feats[b, t], q[b]>: feats [B, T, 768], one question embedding q [B, 768] per clip.
 * (The reference scores frames with a caption cross-encoder, gen_sample.py:80-83; only the top-K that
 * follows is pinned by it.)  Whole path: clips [B, T, H, W, 3] uint8 -> encoder -> scores -> strided top-K
 * (-> optional gather of the K picks); optional outputs may be NULL. */
int sasvqa_mif_scores(const float* feats_dev, const float* q_dev, int B, int T, float* scores_dev, void* stream);
int sasvqa_mif_sample_u8_hw(SasvqaEncoder* enc, const uint8_t* clips_hwc_dev, int B, int T, int H, int W,
                            const float* q_dev, int K, int ds_rate, int32_t* idx_dev, float* scores_or_null_dev,
                            float* feats_or_null_dev, float* sampled_or_null_dev, void* stream);

/* the same from HOST buffers (clips_hwc_host [B, T, H, W, 3] uint8, q_host [B, 768] fp32; pinned memory makes the copies
 * asynchronous), streamed through the double-buffered pipeline of sasvqa_mdf_sample_host: idx_host [B, K] and, if not
 * NULL, the sampled fp32 frames [B, K, 3*224*224] land in host memory before the call returns. */
int sasvqa_mif_sample_host_hw(SasvqaEncoder* enc, const uint8_t* clips_hwc_host, int B, int T, int H, int W,
                              const float* q_host, int K, int ds_rate, int32_t* idx_host, float* sampled_or_null_host);

/* ---- K5: gather the selected frames as normalised fp32 rows (utils.py:94, extract_features.py:96)
 * out [B, K, 3*224*224]; out-of-range indices give zero rows. */
int sasvqa_gather_frames_u8(const uint8_t* clips_hwc_dev, const int32_t* idx_dev, int B, int T, int K,
                            float* out_dev, void* stream);
int sasvqa_gather_frames_f32(const float* frames_dev, const int32_t* idx_dev, int B, int T, int K,
                             int64_t row_elems, float* out_dev, void* stream);

/* ---- whole path, device-resident clips ------------------------------------------------------
 * clips [B, T, 224, 224, 3] uint8 -> idx [B, K], status [B]; optional outputs may be NULL:
 * lcl_avg [B, T], feats [B, T, 768], sampled [B, K, 3*224*224] fp32.  W may be -1 (adaptive). */
int sasvqa_mdf_sample_u8(SasvqaEncoder* enc, const uint8_t* clips_hwc_dev, int B, int T, int K, int W,
                         int32_t* idx_dev, int32_t* status_dev, float* lcl_avg_or_null_dev,
                         float* feats_or_null_dev, float* sampled_or_null_dev, void* stream);
/* same for decoded frames of any size [B, T, H, W, 3]: K0 runs per chunk in front of K1, and on the K picks
 * of every clip in front of the gather (the sampled rows are the processor's output for those frames) */
int sasvqa_mdf_sample_u8_hw(SasvqaEncoder* enc, const uint8_t* clips_hwc_dev, int B, int T, int H, int W, int K,
                            int Wwin, int32_t* idx_dev, int32_t* status_dev, float* lcl_avg_or_null_dev,
                            float* feats_or_null_dev, float* sampled_or_null_dev, void* stream);
/* same from normalised fp32 CHW frames [B, T, 3, 224, 224] (the reference sampler's input) */
int sasvqa_mdf_sample_f32(SasvqaEncoder* enc, const float* clips_chw_dev, int B, int T, int K, int W,
                          int32_t* idx_dev, int32_t* status_dev, float* lcl_avg_or_null_dev,
                          float* feats_or_null_dev, float* sampled_or_null_dev, void* stream);

/* ---- whole path, RAGGED batch: B clips of different lengths packed back to back --------------------------------
 * frames [sum T, H, W, 3] uint8, clip b = frames [clip_offsets_host[b], clip_offsets_host[b+1]) (offsets start at 0).
 * Per clip exactly what the calls above do for a clip of that length (the reference takes one video of any length per
 * call, extract_features.py:80-97): an empty clip gives SASVQA_STATUS_EMPTY and zero frames, W == -1 is T_b / 20 per
 * clip, the fallback with T_b < K gives SASVQA_STATUS_TOO_FEW.  lcl_avg / feats are packed per frame ([sum T],
 * [sum T, 768]); idx [B, K], status [B], sampled [B, K, 3*224*224]. */
int sasvqa_mdf_sample_ragged_u8(SasvqaEncoder* enc, const uint8_t* frames_hwc_dev, int B, const int32_t* clip_offsets_host,
                                int H, int W, int K, int Wwin, int32_t* idx_dev, int32_t* status_dev,
                                float* lcl_avg_or_null_dev, float* feats_or_null_dev, float* sampled_or_null_dev,
                                void* stream);

/* the same from HOST buffers (frames_hwc_host [sum T, H, W, 3]; pinned memory makes the copies asynchronous): whole clips
 * are grouped up to chunk_frames frames and streamed through the double-buffered pipeline of sasvqa_mdf_sample_host */
int sasvqa_mdf_sample_ragged_host(SasvqaEncoder* enc, const uint8_t* frames_hwc_host, int B, const int32_t* clip_offsets_host,
                                  int H, int W, int K, int Wwin, int32_t* idx_host, int32_t* status_host,
                                  float* sampled_or_null_host);

/* ---- whole path, HOST buffers (the extraction loop extract_features.py:80-97 for a clip list) --
 * Streams clips host->device in double-buffered groups overlapped with compute, writes idx/status
 * (and, if not NULL, the sampled frames -- the rows of the reference's "sampled_frames" dataset)
 * back to host memory.  Returns after everything is on the host. */
int sasvqa_mdf_sample_host(SasvqaEncoder* enc, const uint8_t* clips_hwc_host, int B, int T, int K, int W,
                           int32_t* idx_host, int32_t* status_host, float* sampled_or_null_host);

int sasvqa_mdf_sample_host_hw(SasvqaEncoder* enc, const uint8_t* clips_hwc_host, int B, int T, int H, int W, int K,
                              int Wwin, int32_t* idx_host, int32_t* status_host, float* sampled_or_null_host);

/* ---- downstream consumer: visual tokens of the sampled frames (src/modeling/modeling.py:76-95) ----
 * MyGitModel.forward runs `image_encoder(frame).last_hidden_state` frame by frame, concatenates along the
 * sequence and applies `visual_projection` (HF GitProjection: Linear(768,768) + LayerNorm).  The image
 * encoder is the sampler's encoder: frames [n, 3, 224, 224] fp32 (rows of "sampled_frames") or uint8 HWC ->
 * tokens [n * 197, 768] fp32; project == 0 stops at last_hidden_state.  Projection weights are host fp32:
 * visual_projection.0.weight [768, 768], .0.bias, .1.weight, .1.bias. */
int sasvqa_encoder_set_projection(SasvqaEncoder* enc, const float* weight_host, const float* bias_host,
                                  const float* ln_weight_host, const float* ln_bias_host);
int sasvqa_visual_tokens_f32(SasvqaEncoder* enc, const float* frames_chw_dev, int n_frames, int project,
                             float* tokens_dev, void* stream);
int sasvqa_visual_tokens_u8(SasvqaEncoder* enc, const uint8_t* frames_hwc_dev, int n_frames, int project,
                            float* tokens_dev, void* stream);

/* ---- MIF relevance model: caption cross-encoder (src/preprocessing/gen_sample.py:48-94) --------------------
 * Replaces `AutoModelForSequenceClassification.from_pretrained(args.sim_model)` (gen_sample.py:160; a bert-base
 * sequence classifier: 768 hidden, 12 layers x 12 heads, FFN 3072, 512 positions, 2 token types, LayerNorm eps
 * 1e-12, erf GELU) and `output = model(**inputs); scores = output[0][:,0]` (gen_sample.py:82-83).
 * params_host: the model's state_dict flattened to fp32 in HF key order: bert.embeddings.{word,position,token_type}
 * _embeddings.weight, embeddings.LayerNorm.{weight,bias}, per layer attention.self.{query,key,value}.{weight,bias},
 * attention.output.dense.{weight,bias}, attention.output.LayerNorm.{weight,bias}, intermediate.dense.{weight,bias},
 * output.dense.{weight,bias}, output.LayerNorm.{weight,bias}, then bert.pooler.dense.{weight,bias},
 * classifier.{weight,bias}.  n_params must equal sasvqa_scorer_num_params(vocab_size, num_labels).
 * max_tokens: packed tokens per pass (workspace ~10.8 KB per token); <= 0 picks the default (65 536). */
typedef struct SasvqaScorer SasvqaScorer;
uint64_t sasvqa_scorer_num_params(int vocab_size, int num_labels);
int sasvqa_scorer_create(const float* params_host, uint64_t n_params, int vocab_size, int num_labels, int max_tokens,
                         SasvqaScorer** out);
void sasvqa_scorer_destroy(SasvqaScorer* scorer);
int sasvqa_scorer_max_tokens(const SasvqaScorer* scorer);
/* Device entry: input_ids / token_type_ids [N, L] int32 as the tokenizer pads them (right padding; type ids may be
 * NULL = all zero), lengths_host [N] = attention_mask.sum(1) (1 <= length <= L <= 512; positions past the length are
 * never read) -> logits [N, num_labels] fp32.  Asynchronous on `stream`. */
int sasvqa_scorer_logits(SasvqaScorer* scorer, const int32_t* input_ids_dev, const int32_t* token_type_ids_or_null_dev,
                         const int32_t* lengths_host, int N, int L, float* logits_dev, void* stream);
/* inspection: packed fp32 hidden state [sum(lengths), 768] after `n_layers` blocks (0 = embeddings); the call must
 * fit one pass (sum(lengths) <= max_tokens) */
int sasvqa_scorer_hidden(SasvqaScorer* scorer, const int32_t* input_ids_dev, const int32_t* token_type_ids_or_null_dev,
                         const int32_t* lengths_host, int N, int L, int n_layers, float* hidden_dev, void* stream);
/* Host entry: the tokenizer's int64 arrays as they are (`return_tensors='pt'`, gen_sample.py:80): input_ids,
 * token_type_ids (or NULL), attention_mask (or NULL = no padding; must be ones then zeros) [N, L] -> logits in host
 * memory.  Returns after the results are on the host. */
int sasvqa_scorer_logits_host(SasvqaScorer* scorer, const int64_t* input_ids_host, const int64_t* token_type_ids_or_null_host,
                              const int64_t* attention_mask_or_null_host, int N, int L, float* logits_host);
/* Whole MIF step for G QA samples of T captions each (rows g*T .. g*T+T-1 of the [G*T, L] arrays are sample g's
 * (question, caption_t) pairs): scores[g, t] = logits[g*T + t, label]; idx[g, :] = ds_rate * topk(scores[g, ::ds_rate], K),
 * best first (gen_sample.py:83-88; label 0 there).  idx_host [G, K]; scores_or_null_host [G, T]. */
int sasvqa_mif_select_captions_host(SasvqaScorer* scorer, const int64_t* input_ids_host,
                                    const int64_t* token_type_ids_or_null_host, const int64_t* attention_mask_or_null_host,
                                    int G, int T, int L, int K, int ds_rate, int label, int32_t* idx_host,
                                    float* scores_or_null_host);
/* per-stage timing of the scorer (same protocol as sasvqa_profile_*).  Kinds: 0 embed, 1 layernorm, 2 gemm_qkv,
 * 3 attention, 4 gemm_out_proj, 5 gemm_fc1, 6 gemm_fc2, 7 pooler. */
#define SASVQA_SCORER_PROFILE_KINDS 8
int sasvqa_scorer_profile_enable(SasvqaScorer* scorer, int on);
int sasvqa_scorer_profile_read(SasvqaScorer* scorer, double* ms_out, int64_t* scopes_out, int n_kinds);

/* ---- downstream consumer, text side: the video-QA forward on the sampled frames (src/modeling/modeling.py:29-232) ----
 * MyGitForCausalLM.forward (inference): [visual_projection(image_encoder(frame_0..K-1)) | GitEmbeddings(input_ids)]
 * -> 6 post-LN decoder blocks under the combined mask (visual rows see visual rows; text row t sees all visual rows and
 * text rows <= t) -> `logits = output(sequence_output)`, returned for the TEXT rows (the reference slices the visual
 * rows' logits away, modeling.py:211-215).  The visual side runs on the SasvqaEncoder (its projection must be loaded).
 * params_host: decoder-side entries of the model's state_dict flattened to fp32 in HF key order:
 * git.embeddings.{word,position}_embeddings.weight, git.embeddings.LayerNorm.{weight,bias}, per layer
 * attention.self.{query,key,value}.{weight,bias}, attention.output.dense.{weight,bias},
 * attention.output.LayerNorm.{weight,bias}, intermediate.dense.{weight,bias}, output.dense.{weight,bias},
 * output.LayerNorm.{weight,bias}, then output.{weight,bias} (the vocabulary head).  git-base geometry only.
 * max_rows: rows (visual + text tokens) per pass, ~10.8 KB each; <= 0 picks the default (65 536). */
typedef struct SasvqaGitDecoder SasvqaGitDecoder;
uint64_t sasvqa_git_decoder_num_params(int vocab_size, int n_layers);
int sasvqa_git_decoder_create(const float* params_host, uint64_t n_params, int vocab_size, int n_layers, int max_rows,
                              SasvqaGitDecoder** out);
void sasvqa_git_decoder_destroy(SasvqaGitDecoder* dec);
int sasvqa_git_decoder_vocab_padded(const SasvqaGitDecoder* dec);   /* vocab_size rounded up to 256: the logits row stride */
/* frames [B, K, 3, 224, 224] fp32 (rows of "sampled_frames", what the collator feeds as pixel_values), input_ids [B, L]
 * int32 (right-padded; padded positions give rows the caller ignores) -> logits [B, L, vocab_padded] fp32, columns
 * >= vocab_size zero.  Asynchronous on `stream`. */
int sasvqa_git_vqa_logits_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames_chw_dev, int B, int K,
                              const int32_t* input_ids_dev, int L, float* logits_dev, void* stream);
/* the same forward with `labels` [B, L] int32 (-100 = ignore): loss[0] = CrossEntropyLoss()(logits[:, :-1], labels[:, 1:]),
 * the next-token objective of MyGitForCausalLM.forward (modeling.py:208-215; mean over the labels that are not ignored,
 * NaN when there is none, as torch).  logits_or_null: also return the logits [B, L, vocab_padded]. */
int sasvqa_git_vqa_loss_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames_chw_dev, int B, int K,
                            const int32_t* input_ids_dev, const int32_t* labels_dev, int L, float* loss_dev,
                            float* logits_or_null_dev, void* stream);
/* greedy answer decoding, the reference's evaluation path `self.model.generate(**inputs, max_length=50)`
 * (modeling.py:330-333; HF greedy search): prompt [B, L0] int32 (equal lengths, no padding) -> out_ids [B, max_length]
 * int32 = the prompt followed by one argmax token per step; a sequence that produced eos_token_id emits pad_token_id from
 * then on (HF stops once every sequence has finished: trim the all-pad tail).  The visual rows are run once per group
 * and their per-block keys / values cached; every step runs the text rows against the cache.  Synchronises `stream`
 * every fourth step to stop a group early once all its sequences have finished. */
int sasvqa_git_vqa_generate_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames_chw_dev, int B, int K,
                                const int32_t* prompt_ids_dev, int L0, int max_length, int eos_token_id, int pad_token_id,
                                int32_t* out_ids_dev, void* stream);
/* inspection: fp32 stream [B*K*197 + B*L, 768] after `n_layers` >= 0 blocks (all visual rows first, then all text rows);
 * the call must fit one pass */
int sasvqa_git_vqa_hidden_f32(SasvqaGitDecoder* dec, SasvqaEncoder* enc, const float* frames_chw_dev, int B, int K,
                              const int32_t* input_ids_dev, int L, int n_layers, float* hidden_dev, void* stream);

/* ---- decode front end: NVDEC in front of K0 (src/preprocessing/prefetch_loader.py:57-67) -----------------------------
 * The reference decodes with cv2 on one host thread: `cap.read()` per frame, BGR -> RGB, keep frame i when i % intv == 0.
 * Here the GPU's video engine decodes an ELEMENTARY stream held in host memory (H.264 / HEVC Annex B, 8-bit 4:2:0; demux
 * the container on the host, e.g. `ffmpeg -c copy -bsf:v h264_mp4toannexb`) into device memory, display order:
 *   format 0: uint8 RGB frames [n, H, W, 3] -- the layout the uint8 entry points above take (K0 resizes any H x W);
 *   format 1: raw NV12 [n, H * 3 / 2, W] (luma plane, then interleaved chroma rows).
 * codec: SASVQA_CODEC_*.  sasvqa_video_probe parses without decoding: info_out[0..3] = width, height, frames a decode
 * with this `intv` returns, frames in the stream.  sasvqa_video_decode needs H, W equal to the probed size and room for
 * capacity_frames frames; *n_frames_out = frames written.  Returns after the frames are in device memory (`stream` is
 * synchronised per frame).  Needs libnvcuvid.so.1 (ships with the driver; loaded on first use); SASVQA_ERR_INVALID with
 * a message when it, or an NVDEC engine, is missing.  Colour: BT.601 limited-range integer matrix, chroma of the co-sited
 * 2x2 block -- decode is bit-exact by the codec standard, the colour step is unpinned against cv2's swscale path. */
#define SASVQA_CODEC_H264 4
#define SASVQA_CODEC_HEVC 8
int sasvqa_video_probe(const uint8_t* bitstream_host, uint64_t n_bytes, int codec, int intv, int32_t* info_out);
int sasvqa_video_decode(const uint8_t* bitstream_host, uint64_t n_bytes, int codec, int intv, uint8_t* frames_dev,
                        int capacity_frames, int H, int W, int format, int32_t* n_frames_out, void* stream);

/* ---- instrumentation ------------------------------------------------------------------------
 * sasvqa_launch_count: kernels launched by this library in this process so far.
 * Profiling: when enabled, CUDA-event pairs bracket every stage launch on its stream;
 * sasvqa_profile_read synchronises, sums milliseconds and scope counts per stage kind and resets.
 * Kinds: 0 preprocess, 1 gemm_patch_embed, 2 pre_layernorm, 3 layernorm, 4 gemm_qkv, 5 attention,
 * 6 gemm_out_proj, 7 gemm_fc1, 8 gemm_fc2, 9 pool_norm, 10 scores, 11 select, 12 gather, 13 resize,
 * 14 projection. */
#define SASVQA_PROFILE_KINDS 15
int64_t sasvqa_launch_count(void);
int sasvqa_profile_enable(SasvqaEncoder* enc, int on);
int sasvqa_profile_read(SasvqaEncoder* enc, double* ms_out, int64_t* scopes_out, int n_kinds);

/* ---- test hooks (not on the product path) ---------------------------------------------------
 * One encoder GEMM with a fused epilogue (mode = 0 bias->bf16, 1 bias+quick_gelu->bf16,
 * 2 bias+residual in place fp32, 3 patch-embed scatter + position embedding, 4 bias+erf-gelu->bf16) on the product's
 * tcgen05 kernel.  The library has ONE backend per operation and no runtime switch: the CUDA-core / mma.sync check
 * kernels the parity tests compare against live in a separate test-only library (libsasvqa_b200_test.so,
 * csrc/check/: sasvqa_check_gemm_simt, sasvqa_check_attention_mma). */
int sasvqa_test_gemm(const uint16_t* a_bf16_dev, const uint16_t* b_bf16_dev, int M, int N, int K, int mode,
                     const float* bias_or_pos_dev, uint16_t* out_bf16_dev, float* out_f32_dev, void* stream);
/* the encoder's per-frame attention (tcgen05 kernel); trace_variant > 0 selects an instrumented timing variant of
 * the same kernel (tools/att_trace.py), 0 = the product launch */
int sasvqa_test_attention(const uint16_t* qkv_bf16_dev, int n_frames, uint16_t* out_bf16_dev, int trace_variant, void* stream);
int sasvqa_test_layernorm(const float* x_dev, int rows, const float* gamma_dev, const float* beta_dev,
                          uint16_t* out_bf16_dev, void* stream);
/* variable-length attention of the scorer: packed qkv [M, 2304] bf16, cu_seqlens_dev [n_seqs + 1] -> out [M, 768] */
/* attention of the GIT decoder: qkv [n*(n_vis+L), 2304] bf16 stored visual-first -> out [., 768] */
int sasvqa_test_attention_git(const uint16_t* qkv_bf16_dev, int n_samples, int n_vis, int L, uint16_t* out_bf16_dev, void* stream);
int sasvqa_test_attention_varlen(const uint16_t* qkv_bf16_dev, const int32_t* cu_seqlens_dev, int n_seqs, int max_len,
                                 uint16_t* out_bf16_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SASVQA_H */
