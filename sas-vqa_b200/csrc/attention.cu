// Small attention kernels on bf16 mma.sync.m16n8k16 with fp32 softmax statistics:
//  * attention_git_kernel     -- the GIT decoder's TEXT rows against a per-layer k|v CACHE of the visual rows: the steps of
//                                incremental greedy decoding (git_decoder.cu: git_vqa_generate).  The forward itself -- visual
//                                and text query tiles -- runs on tcgen05 (attention_git_tcgen05.cu).
//  * attention_short_kernel / attention_varlen_kernel -- the MIF caption cross-encoder's (question, caption) pairs
//                                (~20 tokens; a tcgen05 instruction needs 128 query rows, six times a pair's length)
// The encoder's per-frame attention is attention_tcgen05.cu.
#include <algorithm>

#include "attention_mma.cuh"

namespace sasvqa {

namespace {

// ---- attention of the downstream GIT decoder over [visual tokens | text] (HF GitSelfAttention, modeling_git.py:202-280,
// with the mask MyGitModel.forward builds, src/modeling/modeling.py:116-140): a visual row sees the n_vis visual rows,
// text row t sees every visual row and text rows <= t.  Both cases are "keys [0, limit(row))" with
// limit = n_vis for visual rows and row + 1 for text rows, so the kernel is the flash loop above with a per-row limit.
// Rows of a group of n samples are stored visual-first: sample s owns rows [s * n_vis, (s+1) * n_vis) and
// [n * n_vis + s * L, n * n_vis + (s+1) * L).  One CTA of 8 warps per (sample, head, 128-query block); K and V stream
// through a double-buffered shared-memory chunk of 64 keys (cp.async), warps skip chunks past their own limit.
constexpr int GIT_WARPS = 8;
constexpr int GIT_QBLOCK = GIT_WARPS * 16;

__global__ void __launch_bounds__(GIT_WARPS * 32, 2)
attention_git_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n_samples, int n_vis, int L,
                     int q_blocks, int q_row0, const __nv_bfloat16* __restrict__ vis_kv, long long txt_row0) {
    __shared__ __align__(128) uint8_t kv[2][2][64 * 128];              // [buffer][K | V][64 keys x 64 d bf16]
    const int S = n_vis + L;
    const int qb = blockIdx.x % q_blocks;
    const int head = (blockIdx.x / q_blocks) % kHeads;
    const int smp = blockIdx.x / (q_blocks * kHeads);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // vis_kv != nullptr (incremental decoding): qkv / out hold the TEXT rows only (sample s at txt_row0 + s * L) and the
    // visual keys and values come from the per-layer cache vis_kv [n * n_vis, k | v] written by the prefill pass
    const bool cached = vis_kv != nullptr;
    const long long vis_base = (long long)smp * n_vis, txt_base = txt_row0 + (long long)smp * L;
    auto row_of = [&](int j) -> long long { return j < n_vis ? vis_base + j : txt_base + (j - n_vis); };
    const int q0 = q_row0 + qb * GIT_QBLOCK;                     // q_row0 = n_vis: only text rows are queries
    const int q_last = min(q0 + GIT_QBLOCK, S) - 1;                      // last real query row of this block
    const int block_limit = q_last < n_vis ? n_vis : q_last + 1;         // keys any row of the block can see
    const int n_chunks = (block_limit + 63) >> 6;
    const __nv_bfloat16* kbase = qkv + kHidden + head * kHeadDim;
    const __nv_bfloat16* vbase = qkv + 2 * kHidden + head * kHeadDim;

    auto stage = [&](int chunk, int buf) {
        const uint32_t k_smem = (uint32_t)__cvta_generic_to_shared(kv[buf][0]);
        const uint32_t v_smem = (uint32_t)__cvta_generic_to_shared(kv[buf][1]);
        for (int i = threadIdx.x; i < 64 * 8; i += GIT_WARPS * 32) {
            const int r = i >> 3, c = i & 7;
            const int j = min(chunk * 64 + r, S - 1);                    // rows past the sequence: any real row (masked)
            const long long row = row_of(j);
            if (cached && j < n_vis) {
                const __nv_bfloat16* src = vis_kv + row * (2 * kHidden) + head * kHeadDim + c * 8;
                cp_async16(k_smem + tile_off(r, c), src);
                cp_async16(v_smem + tile_off(r, c), src + kHidden);
            } else {
                cp_async16(k_smem + tile_off(r, c), kbase + row * kQkv + c * 8);
                cp_async16(v_smem + tile_off(r, c), vbase + row * kQkv + c * 8);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // this warp's 16 query rows
    const int g = lane >> 2, t = lane & 3;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    int r0c = min(r0, S - 1), r1c = min(r1, S - 1);
    if (cached) {                                                        // visual query rows do not exist in this mode
        r0c = max(r0c, n_vis);
        r1c = max(r1c, n_vis);
    }
    const int lim0 = r0c < n_vis ? n_vis : r0c + 1, lim1 = r1c < n_vis ? n_vis : r1c + 1;
    const int w_last = min(q0 + warp * 16 + 15, S - 1);
    const int warp_limit = w_last < n_vis ? n_vis : w_last + 1;          // warp-uniform
    const bool warp_active = q0 + warp * 16 < S;
    uint32_t qf[4][4];
    {
        const __nv_bfloat16* qa = qkv + row_of(r0c) * kQkv + head * kHeadDim + 2 * t;
        const __nv_bfloat16* qb_ = qkv + row_of(r1c) * kQkv + head * kHeadDim + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qf[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qa + 16 * ks));
            qf[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qb_ + 16 * ks));
            qf[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qa + 16 * ks + 8));
            qf[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qb_ + 16 * ks + 8));
        }
    }
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;

    stage(0, 0);
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < n_chunks) {
            stage(c + 1, buf ^ 1);                                       // buffer buf^1 was released by the barrier below
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                                 // chunk c is in shared memory for every warp
        if (warp_active && c * 64 < warp_limit) {
            const uint32_t k_smem = (uint32_t)__cvta_generic_to_shared(kv[buf][0]);
            const uint32_t v_smem = (uint32_t)__cvta_generic_to_shared(kv[buf][1]);
            // attend_chunk indexes keys as key0 + local column and rows of the tile in shared memory from 0
            attend_chunk<8>(qf, k_smem - (uint32_t)(c * 64) * 128u, v_smem - (uint32_t)(c * 64) * 128u, c * 64, lane, m, l, o, lim0,
                            lim1);
        }
        __syncthreads();                                                 // everyone is done with buffer buf
    }
    if (!warp_active) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
    __nv_bfloat16* o0 = out + row_of(r0c) * kHidden + head * kHeadDim + 2 * t;
    __nv_bfloat16* o1 = out + row_of(r1c) * kHidden + head * kHeadDim + 2 * t;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        if (r0 < S && r0 == r0c) *reinterpret_cast<uint32_t*>(o0 + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
        if (r1 < S && r1 == r1c) *reinterpret_cast<uint32_t*>(o1 + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
}

}  // namespace

// text_only != 0: only the text rows are queries (the last decoder block when nobody reads the visual rows afterwards)
// vis_kv != nullptr: incremental decoding -- qkv / out are the text rows only ([n * L, .], sample-major), the visual keys
// and values are read from the cache [n * n_vis, 1536] (k | v) of this layer; only text query blocks run
int launch_attention_git(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_samples, int n_vis, int L, int text_only,
                         cudaStream_t s, const __nv_bfloat16* vis_kv) {
    if (n_samples == 0) return 0;
    if (vis_kv != nullptr) text_only = 1;
    SASVQA_REQUIRE(n_vis >= 1 && L >= 0, "the visual prefix must hold at least one token");
    SASVQA_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 3) == 0, "unaligned buffers");
    // text_only: query blocks are laid over the text rows alone (first block starts at row n_vis), so no warp spends its
    // chunks on visual rows nobody reads
    const int q_row0 = text_only ? n_vis : 0;
    const int q_blocks = (n_vis + L - q_row0 + GIT_QBLOCK - 1) / GIT_QBLOCK;
    if (q_blocks <= 0) return 0;                                  // text_only with L == 0
    const long long grid = (long long)n_samples * kHeads * q_blocks;
    SASVQA_REQUIRE(grid < 2147483647LL, "too many attention blocks for one launch");
    const long long txt_row0 = vis_kv != nullptr ? 0 : (long long)n_samples * n_vis;
    attention_git_kernel<<<(unsigned)grid, GIT_WARPS * 32, 0, s>>>(qkv, out, n_samples, n_vis, L, q_blocks, q_row0, vis_kv,
                                                                   txt_row0);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

namespace {

// ---- variable-length self-attention of the MIF cross-encoder (HF BertSelfAttention, modeling_bert.py:143-207,
// reached from the reference at src/preprocessing/gen_sample.py:82).  Sequences are PACKED: rows
// [cu[s], cu[s+1]) of qkv [M, 2304] belong to sequence s, so the tokenizer's padding never reaches the GPU and the
// additive key mask of the reference (masked keys -> finfo.min before the softmax) reduces to "keys of my own
// sequence".  One CTA of 4 warps per (sequence, head); K and V of the sequence are staged once (rows padded to a
// multiple of 64 with zeros), each warp owns 16-query tiles and runs the same online-softmax chunk loop as above.
constexpr int VAR_WARPS = 4;

__global__ void __launch_bounds__(VAR_WARPS * 32)
attention_varlen_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                        const int32_t* __restrict__ cu_seqlens, int row_base) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    const int head = blockIdx.x % kHeads;
    const int seq = blockIdx.x / kHeads;
    const long long row_begin = cu_seqlens[seq] - row_base;      // row_base = first packed row of this chunk
    const int len = cu_seqlens[seq + 1] - cu_seqlens[seq];
    if (len <= 0) return;
    const int keys_pad = (len + 63) & ~63;
    uint8_t* k_tile = att_smem;
    uint8_t* v_tile = att_smem + keys_pad * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const __nv_bfloat16* base = qkv + row_begin * (long long)kQkv + head * kHeadDim;
    const uint32_t k_smem = (uint32_t)__cvta_generic_to_shared(k_tile);
    const uint32_t v_smem = (uint32_t)__cvta_generic_to_shared(v_tile);

    for (int i = threadIdx.x; i < keys_pad * 8; i += VAR_WARPS * 32) {
        const int r = i >> 3, c = i & 7;
        if (r < len) {
            const __nv_bfloat16* src = base + (long long)r * kQkv + c * 8;
            cp_async16(k_smem + tile_off(r, c), src + kHidden);
            cp_async16(v_smem + tile_off(r, c), src + 2 * kHidden);
        } else {
            *reinterpret_cast<uint4*>(k_tile + tile_off(r, c)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(v_tile + tile_off(r, c)) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int q_tiles = (len + 15) >> 4;
    for (int qt = warp; qt < q_tiles; qt += VAR_WARPS) {
        const int row0 = qt * 16 + g, row1 = row0 + 8;
        const int r0c = min(row0, len - 1), r1c = min(row1, len - 1);
        uint32_t qf[4][4];
        const __nv_bfloat16* q0 = base + (long long)r0c * kQkv + 2 * t;
        const __nv_bfloat16* q1 = base + (long long)r1c * kQkv + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qf[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks));
            qf[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks));
            qf[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks + 8));
            qf[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks + 8));
        }
        float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
        for (int key0 = 0; key0 < len; key0 += 64)          // every chunk starts below len: >= 1 valid key
            attend_chunk<8>(qf, k_smem, v_smem, key0, lane, m, l, o, len);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
        }
        const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
        __nv_bfloat16* o0 = out + (row_begin + row0) * (long long)kHidden + head * kHeadDim + 2 * t;
        __nv_bfloat16* o1 = out + (row_begin + row1) * (long long)kHidden + head * kHeadDim + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            if (row0 < len) *reinterpret_cast<uint32_t*>(o0 + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
            if (row1 < len) *reinterpret_cast<uint32_t*>(o1 + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
    }
}

// Short sequences (max_len <= 64, the usual (question, caption) pair is ~20 tokens): a CTA per (sequence, head)
// would leave three of its four warps idle and pay a block barrier per 20 tokens.  Here every WARP owns one
// (sequence, head) item at a time: it stages that item's K and V in its private shared-memory slice (cp.async +
// __syncwarp, no block barrier), and since all keys fit one chunk the softmax needs no online rescaling.
// Consecutive warps take consecutive heads of one sequence, so their 128-byte row segments are neighbours in L2.
__global__ void __launch_bounds__(VAR_WARPS * 32)
attention_short_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                       const int32_t* __restrict__ cu_seqlens, int row_base, int n_items, int slice_bytes) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* k_tile = att_smem + warp * slice_bytes;
    uint8_t* v_tile = k_tile + slice_bytes / 2;
    const uint32_t k_smem = (uint32_t)__cvta_generic_to_shared(k_tile);
    const uint32_t v_smem = (uint32_t)__cvta_generic_to_shared(v_tile);
    const int g = lane >> 2, t = lane & 3;
    for (int item = blockIdx.x * VAR_WARPS + warp; item < n_items; item += gridDim.x * VAR_WARPS) {
        const int seq = item / kHeads, head = item - seq * kHeads;
        const long long row_begin = cu_seqlens[seq] - row_base;
        const int len = cu_seqlens[seq + 1] - cu_seqlens[seq];
        if (len <= 0) continue;
        const int keys_pad = (len + 15) & ~15;
        const __nv_bfloat16* base = qkv + row_begin * (long long)kQkv + head * kHeadDim;
        __syncwarp();                                          // previous item's ldmatrix reads are done
        for (int i = lane; i < keys_pad * 8; i += 32) {
            const int r = i >> 3, c = i & 7;
            if (r < len) {
                const __nv_bfloat16* src = base + (long long)r * kQkv + c * 8;
                cp_async16(k_smem + tile_off(r, c), src + kHidden);
                cp_async16(v_smem + tile_off(r, c), src + 2 * kHidden);
            } else {
                *reinterpret_cast<uint4*>(k_tile + tile_off(r, c)) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(v_tile + tile_off(r, c)) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int q_tiles = keys_pad >> 4;
        for (int qt = 0; qt < q_tiles; ++qt) {
            const int row0 = qt * 16 + g, row1 = row0 + 8;
            const int r0c = min(row0, len - 1), r1c = min(row1, len - 1);
            uint32_t qf[4][4];
            const __nv_bfloat16* q0 = base + (long long)r0c * kQkv + 2 * t;
            const __nv_bfloat16* q1 = base + (long long)r1c * kQkv + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qf[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks));
                qf[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks));
                qf[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks + 8));
                qf[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks + 8));
            }
            float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[8][4];
#pragma unroll
            for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
            switch (q_tiles) {                                  // all keys in one chunk (16 keys per step)
                case 1: attend_chunk<2>(qf, k_smem, v_smem, 0, lane, m, l, o, len); break;
                case 2: attend_chunk<4>(qf, k_smem, v_smem, 0, lane, m, l, o, len); break;
                case 3: attend_chunk<6>(qf, k_smem, v_smem, 0, lane, m, l, o, len); break;
                default: attend_chunk<8>(qf, k_smem, v_smem, 0, lane, m, l, o, len); break;
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
                l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
            }
            const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
            __nv_bfloat16* o0 = out + (row_begin + row0) * (long long)kHidden + head * kHeadDim + 2 * t;
            __nv_bfloat16* o1 = out + (row_begin + row1) * (long long)kHidden + head * kHeadDim + 2 * t;
#pragma unroll
            for (int dt = 0; dt < 8; ++dt) {
                if (row0 < len) *reinterpret_cast<uint32_t*>(o0 + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
                if (row1 < len) *reinterpret_cast<uint32_t*>(o1 + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
            }
        }
    }
}

}  // namespace

int launch_attention_varlen(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int32_t* cu_seqlens_dev, int row_base,
                            int n_seqs, int max_len, cudaStream_t s) {
    if (n_seqs == 0) return 0;
    SASVQA_REQUIRE(max_len >= 1 && max_len <= 512, "sequence length must be in [1, 512]");
    SASVQA_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 3) == 0, "unaligned buffers");
    if (max_len <= 64) {
        const int slice = 2 * ((max_len + 15) & ~15) * 128;  // K + V of one item, <= 16 KiB per warp
        const int smem = VAR_WARPS * slice;
        static SmemAttrCache smem_short;
        if (int rc = smem_short.ensure(attention_short_kernel, smem)) return rc;
        const long long n_items = (long long)n_seqs * kHeads;
        const int grid = (int)std::min<long long>((n_items + VAR_WARPS - 1) / VAR_WARPS, 148 * 16);
        attention_short_kernel<<<grid, VAR_WARPS * 32, smem, s>>>(qkv, out, cu_seqlens_dev, row_base, (int)n_items, slice);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
        return 0;
    }
    const int smem = 2 * ((max_len + 63) & ~63) * 128;      // <= 131 072 B
    static SmemAttrCache smem_attr;
    if (int rc = smem_attr.ensure(attention_varlen_kernel, smem)) return rc;
    attention_varlen_kernel<<<n_seqs * kHeads, VAR_WARPS * 32, smem, s>>>(qkv, out, cu_seqlens_dev, row_base);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
