// Attention of the downstream GIT decoder for the VISUAL rows, on tcgen05 tensor cores (sm_100a).
//
// MyGitModel.forward builds a combined mask (src/modeling/modeling.py:116-140): a visual row sees the n_vis = K * 197
// visual rows of its sample and nothing else; text row t sees every visual row and text rows <= t.  Both are "keys
// [0, limit(row))".  Visual query tiles (99.4 % of the queries at K = 16: 3152 of 3172) run a plain non-causal
// softmax(Q K^T / 8) V over the n_vis keys; the text rows of a sample are ONE more query tile per 128 positions whose
// key chunks are the visual chunks followed by the text chunks up to its own, with the causal limit applied per row in
// those last chunks (a 20-row text tile fills 16 % of the M = 128 UMMA -- accepted, it is 0.6 % of the work, and it keeps
// the whole forward on tcgen05).  Incremental decoding against a k|v cache stays on attention.cu's mma.sync kernel.
//
// qkv [rows, 2304] bf16 (q | k | v, 12 heads x 64), rows stored visual-first: sample s owns rows [s*n_vis, (s+1)*n_vis)
// and [n*n_vis + s*L, n*n_vis + (s+1)*L).  out [rows, 768] bf16.  HF GitSelfAttention, transformers modeling_git.py:202-280.
//
// Flash attention with the accumulators in TMEM.  Persistent CTAs of 256 threads, TWO per SM (each allocates 256 of the
// 512 TMEM columns; while one CTA's softmax runs the other's MMAs have the tensor pipe).  One work item = (sample, head,
// tile of 128 query rows); its keys stream through shared memory in chunks of 64:
//   warp 0     TMA: Q tile once per item, K and V chunks through a 4-stage ring (128B-swizzled 64 x 64 bf16 tiles)
//   warp 1     MMA issuer:  S = Q K_c^T        (M=128, N=64, K=64: 4 UMMAs, both operands K-major smem) into one of TWO S
//                                               buffers, issued one chunk AHEAD: S of chunk c+1 runs under the softmax of chunk c
//                           O += P V_c         (A = P from TMEM as packed bf16, B = V chunk as an MN-major smem operand,
//                                               4 UMMAs of 16 keys alternating between two accumulators O_a / O_b:
//                                               back-to-back UMMAs into one N=64 accumulator are latency-chained)
//              MMAs execute in issue order, so S of chunk c+2 may be issued behind P V of chunk c although it overwrites
//              the buffer P of chunk c lives in.
//   warps 4-7  online softmax, one thread per query row: the 64 scores are read from TMEM once into registers, the
//              running reference maximum only moves when a chunk's maximum exceeds it by more than 2^8 (then -- rarely
//              after the first chunks -- the thread waits for P V of the previous chunk (its own barrier: S of this chunk
//              was issued before that P V) and rescales its rows of O_a / O_b in TMEM), P = exp2(S * scale - ref) goes
//              back as packed bf16 over the consumed columns (every 4th pair of exponentials on the FMA pipe).  "P written"
//              has one barrier per S buffer: a fast warp may be a chunk ahead of a slow one.  After the last chunk the same
//              warps run the epilogue: (O_a + O_b) / l -> bf16 -> swizzled smem slab -> one TMA store per 32 rows.
// TMEM columns: S buffers fp32 [0,64) and [64,128), P packed over the first 32 columns of its buffer; O_a [128,192); O_b [192,256).
#include <algorithm>

#include "tcgen05_util.cuh"

namespace sasvqa {

namespace {

constexpr int GQ_TILE = 128;                      // query rows per work item
constexpr int GK_CHUNK = 64;                      // keys per chunk (two S buffers of 64 TMEM columns: S of chunk c+1 runs under the softmax of chunk c)
constexpr int G_TILE_BYTES = 128 * 128;           // Q tile: 128 rows x 64 bf16
constexpr int G_KV_BYTES = GK_CHUNK * 128;        // K or V chunk: 64 rows x 64 bf16
constexpr int G_KV_STAGES = 4;
constexpr int G_THREADS = 256;
constexpr int G_REGS_CTRL = 40, G_REGS_SOFTMAX = 208;     // 128 * (40 + 208) = 31 744 <= 32 768 per CTA (two CTAs per SM)
constexpr int G_SLAB_BYTES = 32 * 128;
constexpr int G_SMEM = G_TILE_BYTES + 2 * G_KV_STAGES * G_KV_BYTES + 4 * G_SLAB_BYTES + 1024 + 256;    // 99 584 B
constexpr uint32_t G_OA_COL = 128, G_OB_COL = 192, G_TMEM_COLS = 256;
constexpr float kGitScaleLog2e = 0.125f * 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;         // log2 units: P stays <= 2^8 between two moves of the reference maximum

// S = Q K^T : M=128, N=128, A and B K-major, bf16 x bf16 -> f32
constexpr uint32_t kGitIdescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GK_CHUNK >> 3) << 17) | ((128u >> 4) << 24);
// O += P V  : M=128, N=64, A (TMEM) K-major, B MN-major (bit 16): V rows are keys with d contiguous
constexpr uint32_t kGitIdescPV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

#ifndef SASVQA_GIT_POLY_EVERY
#define SASVQA_GIT_POLY_EVERY 4                   // every 4th pair of exponentials on the FMA pipe instead of MUFU (A/B: 0.760 -> 0.726 ms)
#endif
constexpr int kPolyEvery = SASVQA_GIT_POLY_EVERY;

// One chunk of one query row.  v: the row's 128 raw scores; keys >= nvalid (the last chunk of a sample) are masked.
// Returns the chunk's contribution to the row sum and writes P (packed bf16) over S columns [0,64).
template <bool TAIL>
__device__ __forceinline__ float git_exp_and_store(const uint32_t (&v)[GK_CHUNK], float m_ref, int nvalid, uint32_t trow) {
    const uint64_t scale2 = pack_f32x2(kGitScaleLog2e, kGitScaleLog2e), neg_m2 = pack_f32x2(-m_ref, -m_ref);
    uint64_t acc[2] = {0ull, 0ull};
#pragma unroll
    for (int g = 0; g < GK_CHUNK / 32; ++g) {                      // 32 scores -> 16 packed registers -> one TMEM store
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int a = g * 32 + 2 * j;
            float p0, p1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[a]), __uint_as_float(v[a + 1])), scale2, neg_m2), p0, p1);
            if (kPolyEvery > 0 && j % (kPolyEvery > 0 ? kPolyEvery : 1) == kPolyEvery - 1) {
                exp2_poly_x2(p0, p1);
            } else {
                p0 = ex2(p0);
                p1 = ex2(p1);
            }
            if (TAIL) {
                if (a >= nvalid) p0 = 0.f;
                if (a + 1 >= nvalid) p1 = 0.f;
            }
            acc[j & 1] = add_f32x2(acc[j & 1], pack_f32x2(p0, p1));
            pk[j] = pack_bf16x2(p0, p1);
        }
        tmem_st16(trow + (uint32_t)(16 * g), pk);
    }
    float l0, l1;
    unpack_f32x2(add_f32x2(acc[0], acc[1]), l0, l1);
    return l0 + l1;
}

__global__ void __launch_bounds__(G_THREADS, 2)
attention_git_tcgen05_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                             const __grid_constant__ CUtensorMap map_out,
                             __nv_bfloat16* __restrict__ out, int n_vis, int L, long long txt_row0, int n_qtiles_vis,
                             int n_qtiles_txt, int n_chunks_vis, int n_items) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;
    auto k_smem = [&](int st) { return smem_base + (uint32_t)(G_TILE_BYTES + st * G_KV_BYTES); };
    auto v_smem = [&](int st) { return smem_base + (uint32_t)(G_TILE_BYTES + (G_KV_STAGES + st) * G_KV_BYTES); };
    const uint32_t slab_base = smem_base + (uint32_t)(G_TILE_BYTES + 2 * G_KV_STAGES * G_KV_BYTES);
    const uint32_t bar_base = slab_base + 4 * G_SLAB_BYTES;
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto kv_full = [&](int st) { return bar_base + 8u * (2 + st); };
    auto kv_empty = [&](int st) { return bar_base + 8u * (2 + G_KV_STAGES + st); };
    auto s_full = [&](int b) { return bar_base + 8u * (2 + 2 * G_KV_STAGES + b); };       // one per S buffer
    // one "P written" barrier per S buffer: S of chunk c+1 is ready early, so a fast softmax warp may arrive for chunk c+1 before a
    // slow one arrived for chunk c -- with a single barrier the two arrivals would complete the same phase
    auto p_full = [&](int b) { return bar_base + 8u * (4 + 2 * G_KV_STAGES + b); };
    const uint32_t pv_done = bar_base + 8u * (6 + 2 * G_KV_STAGES), o_full = pv_done + 8u;
    const uint32_t tmem_slot = pv_done + 16u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int st = 0; st < G_KV_STAGES; ++st) {
            mbar_init(kv_full(st), 1);
            mbar_init(kv_empty(st), 1);
        }
        mbar_init(s_full(0), 1);
        mbar_init(s_full(1), 1);
        mbar_init(p_full(0), 128);
        mbar_init(p_full(1), 128);
        mbar_init(pv_done, 1);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(G_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // work item -> (sample, head, query tile); tiles [0, n_qtiles_vis) are visual, the rest text tiles.  A tile's key chunks
    // are the n_chunks_vis visual chunks, then (text tiles) the text chunks 0..tq.
    struct Item {
        int smp, head, qt;        // qt < n_qtiles_vis: visual tile qt; else text tile qt - n_qtiles_vis
        bool text;
        int tq, n_chunks;
        long long q_row, vis_row0, txt_base;
    };
    const int n_qtiles = n_qtiles_vis + n_qtiles_txt;
    auto item_coords = [&](int item) {
        Item t;
        t.qt = item % n_qtiles;
        const int sh = item / n_qtiles;
        t.head = sh % kHeads;
        t.smp = sh / kHeads;
        t.text = t.qt >= n_qtiles_vis;
        t.tq = t.text ? t.qt - n_qtiles_vis : 0;
        t.vis_row0 = (long long)t.smp * n_vis;
        t.txt_base = txt_row0 + (long long)t.smp * L;
        t.q_row = t.text ? t.txt_base + (long long)t.tq * GQ_TILE : t.vis_row0 + (long long)t.qt * GQ_TILE;
        t.n_chunks = n_chunks_vis + (t.text ? (t.tq + 1) * (GQ_TILE / GK_CHUNK) : 0);
        return t;
    };
    auto chunk_row = [&](const Item& t, int c) -> int {
        return (int)(c < n_chunks_vis ? t.vis_row0 + (long long)c * GK_CHUNK : t.txt_base + (long long)(c - n_chunks_vis) * GK_CHUNK);
    };

    if (warp < 4) {
        setmaxnreg_dec<G_REGS_CTRL>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            uint32_t it = 0, kv_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const Item t = item_coords(item);
                mbar_wait(q_empty, (it & 1u) ^ 1u);
                mbar_arrive_expect_tx(q_full, G_TILE_BYTES);
                tma_load_2d(q_smem, &map_qkv, t.head * kHeadDim, (int)t.q_row, q_full);
                for (int c = 0; c < t.n_chunks; ++c, ++kv_it) {
                    const int st = (int)(kv_it % G_KV_STAGES);
                    mbar_wait(kv_empty(st), ((kv_it / G_KV_STAGES) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(kv_full(st), 2 * G_KV_BYTES);
                    tma_load_2d(k_smem(st), &map_kv, kHidden + t.head * kHeadDim, chunk_row(t, c), kv_full(st));
                    tma_load_2d(v_smem(st), &map_kv, 2 * kHidden + t.head * kHeadDim, chunk_row(t, c), kv_full(st));
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer =====================
            // S of chunk c+1 goes into the OTHER S buffer while the softmax warps work on chunk c (MMAs execute in issue order, so
            // it is safe behind P V of chunk c-1, the last reader of that buffer); P V of chunk c follows once P is written
            uint32_t it = 0, kv_it = 0, ch = 0;
            auto issue_s = [&](uint64_t adesc, uint32_t kv_index, uint32_t chunk_index) {
                const int st = (int)(kv_index % G_KV_STAGES);
                mbar_wait(kv_full(st), (kv_index / G_KV_STAGES) & 1u);
                tcgen05_fence_after();
                const uint64_t bdesc = desc_sw128(k_smem(st), 0);
                const uint32_t buf = chunk_index & 1u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_ss(tmem_base + buf * GK_CHUNK, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kGitIdescS, k != 0);
                tcgen05_commit(s_full((int)buf));
            };
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int n_chunks = item_coords(item).n_chunks;
                mbar_wait(q_full, it & 1u);
                const uint64_t adesc = desc_sw128(q_smem, 0);
                issue_s(adesc, kv_it, ch);
                for (int c = 0; c < n_chunks; ++c, ++kv_it, ++ch) {
                    if (c + 1 < n_chunks) issue_s(adesc, kv_it + 1, ch + 1);
                    else tcgen05_commit(q_empty);                            // the item's last scores are issued: Q smem reusable
                    const int st = (int)(kv_it % G_KV_STAGES);
                    mbar_wait(p_full((int)(ch & 1u)), (ch >> 1) & 1u);       // softmax wrote P (and rescaled O if needed)
                    tcgen05_fence_after();
                    const uint64_t vdesc = desc_sw128(v_smem(st), GK_CHUNK * 128);
                    const uint32_t pcol = tmem_base + (ch & 1u) * GK_CHUNK;  // P packed over the first half of its S buffer
                    // 16 keys per UMMA = 8 packed-bf16 TMEM columns of P = 2048 B of V; even k-steps -> O_a, odd -> O_b
#pragma unroll
                    for (int j = 0; j < GK_CHUNK / 16; ++j)
                        mma_ts(tmem_base + ((j & 1) ? G_OB_COL : G_OA_COL), pcol + (uint32_t)(8 * j), vdesc + (uint64_t)(128 * j),
                               kGitIdescPV, (c != 0 || j >= 2) ? 1u : 0u);
                    tcgen05_commit(kv_empty(st));                            // K and V of this chunk consumed
                    tcgen05_commit(pv_done);                                 // O holds chunks <= c (a later rescale may touch it)
                    if (c == n_chunks - 1) tcgen05_commit(o_full);
                }
            }
        }
    } else {
        // ===================== online softmax + epilogue =====================
        setmaxnreg_inc<G_REGS_SOFTMAX>();
        const int quarter = warp & 3;                                       // TMEM lane quarter this warp may touch
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t slab = slab_base + (uint32_t)(quarter * G_SLAB_BYTES);
        uint32_t it = 0, ch = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const Item t = item_coords(item);
            const int smp = t.smp, head = t.head;
            // position of this thread's query inside the text (text tiles; rows past L compute on garbage and are never stored)
            const int txt_pos = min(t.tq * GQ_TILE + quarter * 32 + lane, L - 1);
            float m_ref = -INFINITY, l = 0.f;
            for (int c = 0; c < t.n_chunks; ++c, ++ch) {
                // keys [0, limit) of this chunk are visible to this row: the tail of the visual keys, or the causal limit
                const int limit = c < n_chunks_vis ? min(GK_CHUNK, n_vis - c * GK_CHUNK)
                                             : min(GK_CHUNK, txt_pos - (c - n_chunks_vis) * GK_CHUNK + 1);
                const bool full = __all_sync(0xffffffffu, limit == GK_CHUNK);        // warp-uniform: the unmasked fast path
                const uint32_t scol = trow + (ch & 1u) * GK_CHUNK;                     // this chunk's S buffer
                mbar_wait(s_full((int)(ch & 1u)), (ch >> 1) & 1u);
                tcgen05_fence_after();
                uint32_t v[GK_CHUNK];
                tmem_ld32(scol, v);
                tmem_ld32(scol + 32u, v + 32);
                tmem_wait_ld();
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (full) {
#pragma unroll
                    for (int j = 0; j < GK_CHUNK; j += 8)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            mx[q] = max3(mx[q], __uint_as_float(v[j + 2 * q]), __uint_as_float(v[j + 2 * q + 1]));
                } else {
#pragma unroll
                    for (int j = 0; j < GK_CHUNK; ++j)
                        if (j < limit) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(v[j]));
                }
                const float cmax = max3(mx[0], mx[1], fmaxf(mx[2], mx[3])) * kGitScaleLog2e;
                // move the reference maximum only when it is more than 2^8 behind (always on the first chunk: ref = -inf)
                const bool move = cmax > m_ref + kRescaleThreshold;
                if (__any_sync(0xffffffffu, move)) {
                    const float alpha = move ? ex2(m_ref - cmax) : 1.0f;            // first chunk: exp2(-inf) = 0
                    if (move) m_ref = cmax;
                    l *= alpha;
                    if (c != 0) {                                                   // O_a | O_b hold P V of chunks < c: rescale this row
                        // S of this chunk was issued BEFORE P V of chunk c-1: wait for that P V explicitly.  The barrier cannot be
                        // more than one phase ahead (P V of chunk c needs this thread's arrival below), so the parity is unambiguous
                        mbar_wait(pv_done, (ch - 1u) & 1u);
                        tcgen05_fence_after();
#pragma unroll
                        for (int piece = 0; piece < 4; ++piece) {
                            uint32_t o[32];
                            tmem_ld32(trow + G_OA_COL + (uint32_t)(32 * piece), o);
                            tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                            tmem_st16(trow + G_OA_COL + (uint32_t)(32 * piece), o);
                            tmem_st16(trow + G_OA_COL + (uint32_t)(32 * piece + 16), o + 16);
                        }
                    }
                }
                l += full ? git_exp_and_store<false>(v, m_ref, limit, scol) : git_exp_and_store<true>(v, m_ref, limit, scol);
                tmem_wait_st();
                tcgen05_fence_before();
                mbar_arrive(p_full((int)(ch & 1u)));
            }
            // ---- epilogue: this warp's 32 query rows of this head
            mbar_wait(o_full, it & 1u);
            tcgen05_fence_after();
            const int row_in_sample = (t.text ? t.tq : t.qt) * GQ_TILE + quarter * 32;
            const int rows_valid = (t.text ? L : n_vis) - row_in_sample;           // <= 0: the whole slab is past the sample
            const float inv_l = 1.0f / l;
            uint32_t w[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {                                   // two rounds of 32 columns
                uint32_t oa[32], ob[32];
                tmem_ld32(trow + G_OA_COL + (uint32_t)(32 * half), oa);
                tmem_ld32(trow + G_OB_COL + (uint32_t)(32 * half), ob);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    w[16 * half + j] = pack_bf16x2((__uint_as_float(oa[2 * j]) + __uint_as_float(ob[2 * j])) * inv_l,
                                                   (__uint_as_float(oa[2 * j + 1]) + __uint_as_float(ob[2 * j + 1])) * inv_l);
            }
            const long long grow = t.q_row + quarter * 32;                           // global row of this warp's first query
            if (rows_valid >= 32) {
                if (lane == 0) bulk_wait_read_all();                                 // the store that last read this slab is done
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q)                                          // 16-byte chunk q of row `lane`, 128B swizzle
                    st_shared_v4(slab + (uint32_t)(lane * 128 + ((q ^ (lane & 7)) << 4)), w[4 * q], w[4 * q + 1], w[4 * q + 2],
                                 w[4 * q + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_out, slab, head * kHeadDim, (int)grow);
                    bulk_commit();
                }
            } else if (lane < rows_valid) {                                          // the last rows of a sample
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(grow + lane) * kHidden + head * kHeadDim);
#pragma unroll
                for (int q = 0; q < 8; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
        }
        if (lane == 0) bulk_wait_all();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_TMEM_COLS) : "memory");
    }
}

}  // namespace

// Attention of one group of n_samples sequences: qkv / out hold `rows_total` rows (n_samples * n_vis visual rows first, then
// n_samples * L text rows).  include_visual = 0: only the text rows are queries (the last block when nobody reads the
// visual rows afterwards); L = 0: visual rows only (the prefill of incremental decoding).
int launch_attention_git_tcgen05(const __nv_bfloat16* qkv, __nv_bfloat16* out, long long rows_total, int n_samples, int n_vis,
                                 int L, int include_visual, int num_sms, cudaStream_t s) {
    if (n_samples == 0) return 0;
    SASVQA_REQUIRE(n_vis >= 1 && L >= 0 && rows_total >= (long long)n_samples * (n_vis + L), "bad row counts");
    SASVQA_REQUIRE(rows_total < 2147483647LL, "too many rows for one launch");
    SASVQA_REQUIRE(((uintptr_t)qkv & 127) == 0 && ((uintptr_t)out & 127) == 0, "unaligned attention buffers");
    const int n_qtiles_vis = include_visual ? (n_vis + GQ_TILE - 1) / GQ_TILE : 0;
    const int n_qtiles_txt = (L + GQ_TILE - 1) / GQ_TILE;
    const int n_chunks_vis = (n_vis + GK_CHUNK - 1) / GK_CHUNK;
    const long long n_items = (long long)n_samples * kHeads * (n_qtiles_vis + n_qtiles_txt);
    if (n_items == 0) return 0;
    SASVQA_REQUIRE(n_items < 2147483647LL, "too many attention work items for one launch");
    CUtensorMap map_qkv, map_kv, map_out;
    int rc = make_tensor_map_bf16_kmajor(&map_qkv, qkv, (uint64_t)rows_total, kQkv, 128);
    if (rc) return rc;
    if ((rc = make_tensor_map_bf16_kmajor(&map_kv, qkv, (uint64_t)rows_total, kQkv, GK_CHUNK))) return rc;
    if ((rc = make_tensor_map_out(&map_out, out, (uint64_t)rows_total, kHidden, 0))) return rc;
    static SmemAttrCache smem_attr;
    if ((rc = smem_attr.ensure(attention_git_tcgen05_kernel, G_SMEM))) return rc;
    const int grid = (int)std::min<long long>(n_items, 2LL * num_sms);
    attention_git_tcgen05_kernel<<<grid, G_THREADS, G_SMEM, s>>>(map_qkv, map_kv, map_out, out, n_vis, L, (long long)n_samples * n_vis,
                                                                 n_qtiles_vis, n_qtiles_txt, n_chunks_vis, (int)n_items);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
