// Scoring and selection stages of the samplers (all per clip, HBM/latency bound):
//   K3b/K4a  windowed mean cosine similarity  (reference src/preprocessing/datautils/utils.py:55-61)
//   K4b      greedy best-first interval selection with the +-W spacing rule (utils.py:63-88)
//   K4b'     plain top-K fallback                                            (utils.py:91-93)
//   K4c      strided top-K for the MIF sampler          (src/preprocessing/gen_sample.py:87-88)
// Tie rule everywhere: among exactly equal scores the LOWEST index wins (argmax semantics); the
// reference's torch.topk order among equal values is implementation-defined (see oracle/mdf.py).
#include <algorithm>

#include "common.cuh"

namespace sasvqa {

namespace {

// ---------------------------------------------------------------------------------------------
// lcl_avg[b, i] = (sum_{j=i-W}^{i+W-1} <f_i, f_j> - 1) / (2W - 1) for W <= i < T-W, else 0.
// Only the band of the Gram matrix the reference ever reads is computed (it builds all T x T), and by
// linearity as ONE dot product per frame:  sum_j <f_i, f_j> = <f_i, S_i>,  S_i = sum_{j=i-W}^{i+W-1} f_j,
// with the window sum slid along the clip (S_{i+1} = S_i + f_{i+W} - f_{i-W}).  That is 3 row reads and
// ~2.3 kFLOP per frame instead of 2W + 1 reads and 2W dot products: the first version (one warp per frame, 2W
// dots against rows streamed through L1) moved 16x the feature bytes through L1 and sat at 1.0 TB/s of
// algorithmic bytes; two of the three reads here hit L1 (the row was read by this CTA at most 2W steps ago), so
// HBM sees each row once plus a halo of 2W - 1 rows per segment.
// One CTA of 192 threads per (clip, segment of 32 frames): thread t owns columns [4t, 4t + 4) of every row, keeps
// its slice of S in registers, accumulates its partial dot of each of the 32 frames in registers, and one
// transposed reduction through shared memory finishes all 32 at once.  S is rebuilt from its 2W rows at the start
// of every segment, so the rounding drift of the sliding update is bounded by 32 steps (|d lcl| ~ 1e-7).
// ---------------------------------------------------------------------------------------------
constexpr int SC_SEG = 32;
constexpr int SC_THREADS = kHidden / 4;      // 192

// Ragged batches (clip_off != nullptr): clip b owns frames [clip_off[b], clip_off[b+1]) of the packed feature matrix,
// T is then the longest clip (it sizes the grid) and W == -1 selects the adaptive width T_b / 20 per clip (utils.py:32-33).
__global__ void __launch_bounds__(SC_THREADS) mdf_scores_kernel(const float* __restrict__ feats, int T, int W, int n_seg,
                                                                 float* __restrict__ lcl, const int32_t* __restrict__ clip_off) {
    __shared__ float red[SC_SEG][SC_THREADS + 1];
    const long long b = blockIdx.x / n_seg;
    const int seg = blockIdx.x - (int)b * n_seg;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long row0 = b * T;
    if (clip_off != nullptr) {
        row0 = clip_off[b];
        T = clip_off[b + 1] - clip_off[b];
    }
    if (W < 0) W = T / 20;
    const int lo = W, hi = T - W;                       // frames with a full window: [lo, hi)
    float* out = lcl + row0;
    if (seg == 0) {                                      // borders stay exactly 0.0 (utils.py:57)
        for (int i = t; i < T; i += SC_THREADS)
            if (i < lo || i >= hi) out[i] = 0.0f;
    }
    const int i0 = lo + seg * SC_SEG;
    if (i0 >= hi) return;
    const int n = min(SC_SEG, hi - i0);
    const float4* rows = reinterpret_cast<const float4*>(feats + row0 * (long long)kHidden) + t;     // + i * 192 per row
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = i0 - W; j < i0 + W; ++j) {
        const float4 v = __ldg(rows + (long long)j * SC_THREADS);
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
    }
    float part[SC_SEG];
#pragma unroll
    for (int k = 0; k < SC_SEG; ++k) {
        part[k] = 0.f;
        if (k < n) {
            const int i = i0 + k;
            const float4 c = __ldg(rows + (long long)i * SC_THREADS);
            part[k] = fmaf(c.x, S.x, fmaf(c.y, S.y, fmaf(c.z, S.z, c.w * S.w)));
            if (k + 1 < n) {                             // slide: i + W < T holds because i + 1 < hi = T - W
                const float4 e = __ldg(rows + (long long)(i + W) * SC_THREADS);
                const float4 l = __ldg(rows + (long long)(i - W) * SC_THREADS);
                S.x += e.x - l.x; S.y += e.y - l.y; S.z += e.z - l.z; S.w += e.w - l.w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < SC_SEG; ++k) red[k][t] = part[k];
    __syncthreads();
    for (int k = warp; k < n; k += SC_THREADS / 32) {
        float d = 0.f;
#pragma unroll
        for (int m = 0; m < SC_THREADS / 32; ++m) d += red[k][lane + 32 * m];
        d = warp_sum(d);
        if (lane == 0) out[i0 + k] = __fdiv_rn(__fsub_rn(d, 1.0f), (float)(2 * W - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// MIF relevance (BASELINE config 3): scores[b, t] = <feats[b, t], q[b]> with one question embedding per
// clip.  (The reference scores frames with a BERT cross-encoder over generated captions,
// src/preprocessing/gen_sample.py:80-83 -- out of scope; this is the embedding-space surrogate the
// north star names.  What the reference pins is the strided top-K that follows, gen_sample.py:87-88.)
// One warp per (clip, frame) row: 3 KB of features streamed once with 16-byte loads; q stays in L1/L2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mif_scores_kernel(const float* __restrict__ feats, const float* __restrict__ q, int B,
                                                          int T, float* __restrict__ scores) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long long)B * T) return;
    const long long b = wid / T;
    const float4* f = reinterpret_cast<const float4*>(feats + wid * kHidden);
    const float4* qq = reinterpret_cast<const float4*>(q + b * kHidden);
    float d = 0.0f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float4 v = __ldg(f + lane + 32 * j), w = __ldg(qq + lane + 32 * j);
        d = fmaf(v.x, w.x, d);
        d = fmaf(v.y, w.y, d);
        d = fmaf(v.z, w.z, d);
        d = fmaf(v.w, w.w, d);
    }
    d = warp_sum(d);
    if (lane == 0) scores[wid] = d;
}

// Optional full Gram (debug / inspection output of sasvqa_mdf_scores): S[b] = F_b F_b^T.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ feats, int T, float* __restrict__ gram) {
    __shared__ float fa[16][33], fb[16][33];
    const long long b = blockIdx.z;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r = blockIdx.y * 16 + ty, c = blockIdx.x * 16 + tx;
    const float* F = feats + b * T * (long long)kHidden;
    float acc = 0.f;
    for (int k0 = 0; k0 < kHidden; k0 += 32) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int rr = e >> 5, kk = e & 31;
            const int ra = blockIdx.y * 16 + rr, rb = blockIdx.x * 16 + rr;
            fa[rr][kk] = ra < T ? F[(long long)ra * kHidden + k0 + kk] : 0.f;
            fb[rr][kk] = rb < T ? F[(long long)rb * kHidden + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) acc = fmaf(fa[ty][kk], fb[tx][kk], acc);
        __syncthreads();
    }
    if (r < T && c < T) gram[(b * T + r) * (long long)T + c] = acc;
}

// ---------------------------------------------------------------------------------------------
// warp-wide first-argmax of v[lo, hi): highest value, lowest index among equals.  hi > lo.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_first_argmax(const float* __restrict__ v, int lo, int hi, int lane, float* best_v) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lo + lane; i < hi; i += 32) {
        const float x = v[i];
        if (x > bv || bi == 0x7fffffff) {
            bv = x;
            bi = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
            bv = ov;
            bi = oi;
        }
    }
    *best_v = bv;
    return bi;
}

// ---------------------------------------------------------------------------------------------
// K4b: one warp per clip.  Open intervals live in shared memory as (score, l, r, argmax); the
// next pick is the interval with the highest score, smaller left edge on ties -- exactly the
// order in which the reference's heap of (-v, (l, r), idx) tuples pops.  At most K+1 intervals
// are ever open.  status: 0 = K picks found, 1 = fewer than K (caller runs the top-K fallback),
// 3 = T < K (the reference's fallback raises).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) mdf_greedy_kernel(const float* __restrict__ lcl_all, int T, int K, int W,
                                                         int32_t* __restrict__ idx_all, int32_t* __restrict__ status,
                                                         const int32_t* __restrict__ clip_off) {
    extern __shared__ int32_t sel_smem[];
    float* iv_score = reinterpret_cast<float*>(sel_smem);
    int32_t* iv_l = sel_smem + (K + 2);
    int32_t* iv_r = sel_smem + 2 * (K + 2);
    int32_t* iv_p = sel_smem + 3 * (K + 2);
    const int lane = threadIdx.x;
    const long long b = blockIdx.x;
    long long row0 = b * T;
    if (clip_off != nullptr) {                                  // ragged batch: this clip's own length (and window)
        row0 = clip_off[b];
        T = clip_off[b + 1] - clip_off[b];
    }
    if (W < 0) W = T / 20;
    const float* lcl = lcl_all + row0;
    int32_t* out = idx_all + b * K;
    if (T == 0) {                                               // utils.py:50-52: empty clip, 'Zeros'
        for (int k = lane; k < K; k += 32) out[k] = -1;
        if (lane == 0) status[b] = 2;
        return;
    }

    int n_open = 0;
    auto push = [&](int l, int r) {
        float v;
        const int p = warp_first_argmax(lcl, l, r, lane, &v);
        if (lane == 0) {
            iv_score[n_open] = v;
            iv_l[n_open] = l;
            iv_r[n_open] = r;
            iv_p[n_open] = p;
        }
        ++n_open;
        __syncwarp();
    };

    float v0;
    const int top = warp_first_argmax(lcl, 0, T, lane, &v0);
    if (lane == 0) out[0] = top;
    int n_picks = 1;
    if (top - W > 0) push(0, top - W);
    if (top + W < T) push(top + W, T);
    while (n_picks < K && n_open > 0) {
        // best open interval: max score, then min l
        float bv = -INFINITY;
        int bl = 0x7fffffff, bj = -1;
        for (int j = lane; j < n_open; j += 32) {
            const float s = iv_score[j];
            const int l = iv_l[j];
            if (bj < 0 || s > bv || (s == bv && l < bl)) {
                bv = s;
                bl = l;
                bj = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0 && (bj < 0 || ov > bv || (ov == bv && ol < bl))) {
                bv = ov;
                bl = ol;
                bj = oj;
            }
        }
        const int l = iv_l[bj], r = iv_r[bj], p = iv_p[bj];
        __syncwarp();
        if (lane == 0) {
            out[n_picks] = p;
            const int last = n_open - 1;     // remove by moving the last interval into the hole
            iv_score[bj] = iv_score[last];
            iv_l[bj] = iv_l[last];
            iv_r[bj] = iv_r[last];
            iv_p[bj] = iv_p[last];
        }
        --n_open;
        ++n_picks;
        __syncwarp();
        if (p - W > l) push(l, p - W);
        if (p + W < r) push(p + W, r);
    }
    // fewer than K picks: the tail is defined (-1 = "no frame", gathers as a zero row) -- status 1 rows are overwritten
    // by the top-K fallback that follows, status 3 rows (T < K, the reference raises) keep their n_picks real picks
    for (int k = n_picks + lane; k < K; k += 32) out[k] = -1;
    if (lane == 0) status[b] = (n_picks >= K) ? 0 : (T < K ? 3 : 1);
}

// ---------------------------------------------------------------------------------------------
// Top-K (descending score, lowest index first among equals) over v[0 : T : stride].
// Sort key: (~orderable(score) << 32) | position, ascending.  -0.0 == +0.0, NaN sorts first
// (torch.topk treats NaN as the largest value).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t topk_key(float x, uint32_t pos) {
    uint32_t u = __float_as_uint(x == 0.0f ? 0.0f : x);
    if (x != x) u = 0x7fffffffu;                                // canonical +NaN, above +inf
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);            // order-preserving map to unsigned
    return ((uint64_t)(~u) << 32) | pos;
}

// one CTA per row; n = ceil(T / stride) <= n_pad (power of two) keys bitonic-sorted in shared memory
__global__ void __launch_bounds__(256) topk_bitonic_kernel(const float* __restrict__ scores, int T, int stride, int K,
                                                            int n_pad, int32_t* __restrict__ idx_all,
                                                            const int32_t* __restrict__ only_if_status,
                                                            const int32_t* __restrict__ clip_off) {
    extern __shared__ uint64_t keys[];
    const long long b = blockIdx.x;
    if (only_if_status != nullptr && only_if_status[b] != 1) return;
    long long row0 = b * T;
    if (clip_off != nullptr) {
        row0 = clip_off[b];
        T = clip_off[b + 1] - clip_off[b];
    }
    const int n = (T + stride - 1) / stride;
    const float* v = scores + row0;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
        keys[i] = i < n ? topk_key(v[(long long)i * stride], (uint32_t)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int partner = i ^ j;
                if (partner > i) {
                    const uint64_t a = keys[i], c = keys[partner];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) {
                        keys[i] = c;
                        keys[partner] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        idx_all[b * K + i] = (int32_t)((uint32_t)(keys[i] & 0xffffffffu)) * stride;
}

// any T: K rounds of a block-wide "smallest key greater than the previous pick"
__global__ void __launch_bounds__(256) topk_rounds_kernel(const float* __restrict__ scores, int T, int stride, int K,
                                                           int32_t* __restrict__ idx_all,
                                                           const int32_t* __restrict__ only_if_status,
                                                           const int32_t* __restrict__ clip_off) {
    __shared__ uint64_t red[8];
    __shared__ uint64_t prev_s;
    const long long b = blockIdx.x;
    if (only_if_status != nullptr && only_if_status[b] != 1) return;
    long long row0 = b * T;
    if (clip_off != nullptr) {
        row0 = clip_off[b];
        T = clip_off[b + 1] - clip_off[b];
    }
    const int n = (T + stride - 1) / stride;
    const float* v = scores + row0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t prev = 0;
    bool have_prev = false;
    for (int round = 0; round < K; ++round) {
        uint64_t best = ~0ull;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = topk_key(v[(long long)i * stride], (uint32_t)i);
            if ((!have_prev || key > prev) && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t m = red[0];
            for (int w = 1; w < 8; ++w) m = red[w] < m ? red[w] : m;
            prev_s = m;
            idx_all[b * K + round] = (int32_t)((uint32_t)(m & 0xffffffffu)) * stride;
        }
        __syncthreads();
        prev = prev_s;
        have_prev = true;
        __syncthreads();
    }
}

}  // namespace

int launch_mdf_scores(const float* feats, int B, int T, int W, float* lcl_avg, float* gram, cudaStream_t s,
                      const int32_t* clip_off) {
    if (B == 0 || T == 0) return 0;
    SASVQA_REQUIRE(W >= 0 || clip_off != nullptr, "W must be >= 0 here (resolve W = -1 to T / 20 on the host)");
    SASVQA_REQUIRE(clip_off == nullptr || gram == nullptr, "the full Gram output is for uniform batches");
    SASVQA_REQUIRE(((uintptr_t)feats & 15) == 0, "feats must be 16-byte aligned");
    const int n_seg = std::max(1, (T - 2 * std::max(W, 0) + SC_SEG - 1) / SC_SEG);
    const long long blocks = (long long)B * n_seg;
    SASVQA_REQUIRE(blocks < 2147483647LL, "too many frames for one scores launch");
    mdf_scores_kernel<<<(unsigned)blocks, SC_THREADS, 0, s>>>(feats, T, W, n_seg, lcl_avg, clip_off);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (gram != nullptr) {
        SASVQA_REQUIRE(B <= 65535, "gram output supports at most 65535 clips per call");
        dim3 grid((T + 15) / 16, (T + 15) / 16, B);
        gram_kernel<<<grid, 256, 0, s>>>(feats, T, gram);
        SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    }
    return 0;
}

int launch_mif_scores(const float* feats, const float* q, int B, int T, float* scores, cudaStream_t s) {
    if (B == 0 || T == 0) return 0;
    SASVQA_REQUIRE((((uintptr_t)feats | (uintptr_t)q) & 15) == 0, "feats and q must be 16-byte aligned");
    const long long blocks = ((long long)B * T + 7) / 8;
    SASVQA_REQUIRE(blocks < 2147483647LL, "too many frames for one scores launch");
    mif_scores_kernel<<<(unsigned)blocks, 256, 0, s>>>(feats, q, B, T, scores);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_topk_strided(const float* scores, int B, int T, int ds_rate, int K, int32_t* idx,
                        const int32_t* only_if_status, cudaStream_t s, const int32_t* clip_off) {
    if (B == 0 || K == 0) return 0;
    SASVQA_REQUIRE(ds_rate >= 1, "ds_rate must be >= 1");
    const int n = (T + ds_rate - 1) / ds_rate;                  // ragged: T is the longest clip, rows are filtered by status
    SASVQA_REQUIRE(K <= n, "selected index k out of range");
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    if (n_pad <= 4096) {
        topk_bitonic_kernel<<<B, 256, n_pad * sizeof(uint64_t), s>>>(scores, T, ds_rate, K, n_pad, idx, only_if_status, clip_off);
    } else {
        topk_rounds_kernel<<<B, 256, 0, s>>>(scores, T, ds_rate, K, idx, only_if_status, clip_off);
    }
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_mdf_select(const float* lcl_avg, int B, int T, int K, int W, int32_t* idx, int32_t* status,
                      cudaStream_t s, const int32_t* clip_off) {
    if (B == 0) return 0;
    SASVQA_REQUIRE(T >= 1 && K >= 1 && (W >= 0 || clip_off != nullptr), "mdf_select needs T >= 1, K >= 1, W >= 0");
    SASVQA_REQUIRE(K <= 2048, "K too large");
    const size_t smem = 4 * (size_t)(K + 2) * sizeof(int32_t);
    mdf_greedy_kernel<<<B, 32, smem, s>>>(lcl_avg, T, K, W, idx, status, clip_off);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (T >= K) {   // fallback rows (status == 1): discard the greedy picks, plain top-K (utils.py:91-93)
        int rc = launch_topk_strided(lcl_avg, B, T, 1, K, idx, status, s, clip_off);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace sasvqa
