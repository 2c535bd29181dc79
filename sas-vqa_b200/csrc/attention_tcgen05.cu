// Per-frame multi-head self-attention of the ViT encoder on tcgen05 tensor cores (sm_100a):
// 197 tokens, 12 heads x 64, softmax(Q K^T / 8) V with fp32 statistics (HF eager_attention_forward,
// transformers modeling_git.py:556-575, reached from src/preprocessing/datautils/utils.py:40).
//
// Input  qkv [n*197, 2304] bf16 (q | k | v).   Output out [n*197, 768] bf16 (heads concatenated).
//
// Persistent CTAs (512 threads, four warpgroups with setmaxnreg register budgets), one work item =
// (frame, head), queries in two halves h of 128 rows.  Roles:
//   warp 0      TMA: Q (2 x 128 rows) + K (208 rows) and V (208 rows) -> 128B-swizzled smem; the Q|K and the V
//               buffers are double buffered and recycled separately (Q|K die after S, V after P V)
//   warps 1-2   MMA issuers (one per half): S_h = Q_h K^T  (M=128, N=208, K=64 -> 4 UMMAs, both operands K-major smem)
//                           O_h = P_h V    (M=128, N=64, K=208 -> 13 UMMAs; A = P from TMEM, B = V as an
//                                           MN-major smem operand, i.e. V exactly as TMA delivered it).
//               Back-to-back UMMAs into ONE accumulator are latency-chained (~150 cycles each at N=64,
//               measured), so P V runs as two independent chains -- keys [0,112) -> O_a, keys [112,208) ->
//               O_b, issued alternately -- and the epilogue adds the two accumulators.
//               Issue order per half h: PV_h(i), S_h(i+1): the next item's scores are ready before the softmax
//               warps finish the current item.  Each half has its own issuing thread (warps 1 and 2).
//   warps 4-11  softmax only.  All eight warps work on the same query half; the two warps that share a TMEM
//               lane quarter (w, w+4) split a row's keys: part 0 = keys [0,112), part 1 = [112,208), exchanging
//               the row max through shared memory.  Each thread reads its scores from TMEM once and keeps
//               them in registers.  (Moving a share of the ex2 to an FMA-pipe polynomial was measured and is
//               slower: 0.59 -> 0.61 ms at 1/3 -- the stage is latency-, not MUFU-bound.)
//   warps 12-15 epilogue: (O_a + O_b) / l -> bf16 -> 128B-swizzled smem slab of 32 rows -> one TMA store per
//               warp (row stride in global memory is 1536 B: direct stores would be 64-byte fragments).
//               The TMEM slot is released as soon as O is in registers.
// TMEM slot per half (256-column stride): S_h fp32 in columns [0,208); P = exp2(S - max) as packed bf16 goes
// to columns [0,104) (part 0: [0,56), part 1: [56,104) -- written only after both parts hold their scores in
// registers); O_a in [104,168), O_b in [168,232).
#include "tcgen05_util.cuh"

namespace sasvqa {

namespace {

constexpr int KEYS = 208;                         // 197 keys padded to a multiple of 16
constexpr int Q_HALF_BYTES = 128 * 128;           // 128 rows x 64 bf16
constexpr int Q_BYTES = 2 * Q_HALF_BYTES;
constexpr int KV_BYTES = KEYS * 128;              // 26 624
constexpr int QK_BYTES = Q_BYTES + KV_BYTES;         // 59 392 (multiple of 1024)
constexpr int ATT_THREADS = 512;
constexpr int REGS_CTRL = 56, REGS_SOFTMAX = 168, REGS_EPILOGUE = 112;   // 128 * (56 + 2 * 168 + 112) = 64 512
constexpr int XMAX_BYTES = 2 * 2 * 128 * 4;          // row max exchange: [half][part][128 rows] fp32
constexpr int XSUM_BYTES = 2 * 2 * 2 * 128 * 4;      // partial row sums: [item parity][half][part][128 rows] fp32
constexpr int SLAB_BYTES = 32 * 128;                 // output staging: 32 query rows x 64 bf16 of one head
constexpr int STAGE_BYTES = 2 * 4 * SLAB_BYTES;      // [half][quarter]
constexpr int ATT_SMEM = 2 * QK_BYTES + 2 * KV_BYTES + STAGE_BYTES + XMAX_BYTES + XSUM_BYTES + 1024 + 256;
constexpr int P1_COL = 56;                         // packed-bf16 P of part 1 (keys [112,208)) inside a slot
constexpr int OA_COL = 104, OB_COL = 168;          // the two O accumulators inside a slot
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

// One thread's share of a softmax row: the NV = 112 (part 0, S columns [0,112)) or 96 (part 1, S columns
// [112,208), of which 85 are real keys) scores are read from TMEM ONCE into registers; the row max is combined
// with the partner warp through shared memory; P = exp2(S*scale - max) goes back as packed bf16 over the
// consumed columns.  Returns this part's partial row sum (fp32, before the bf16 rounding of P).
template <int PART, class Exchange>
__device__ __forceinline__ float softmax_part(uint32_t trow, Exchange&& exchange_max) {
    constexpr int NV = PART == 0 ? 112 : 96;
    constexpr int VALID = PART == 0 ? 112 : kTokens - 112;
    constexpr uint32_t S_COL = PART == 0 ? 0u : 112u;
    constexpr uint32_t P_COL = PART == 0 ? 0u : (uint32_t)P1_COL;
    uint32_t v[NV];
    tmem_ld32(trow + S_COL, v);
    tmem_ld32(trow + S_COL + 32u, v + 32);
    tmem_ld32(trow + S_COL + 64u, v + 64);
    if (PART == 0) tmem_ld16(trow + 96u, v + 96);
    tmem_wait_ld();
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < VALID; j += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int a = j + 2 * c, b = a + 1;
            if (b < VALID) mx[c] = max3(mx[c], __uint_as_float(v[a]), __uint_as_float(v[b]));
            else if (a < VALID) mx[c] = fmaxf(mx[c], __uint_as_float(v[a]));
        }
    }
    const float m2 = exchange_max(max3(mx[0], mx[1], fmaxf(mx[2], mx[3]))) * kScaleLog2e;
    const uint64_t scale2 = pack_f32x2(kScaleLog2e, kScaleLog2e), neg_m2 = pack_f32x2(-m2, -m2);
    uint64_t acc[2] = {0ull, 0ull};
    uint32_t pk[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
        if (2 * j >= VALID) {                                   // keys >= 197 contribute nothing
            pk[j] = 0u;
            continue;
        }
        float p0, p1;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), scale2, neg_m2), p0, p1);
        if (SASVQA_ATT_POLY_EVERY > 0 && j % SASVQA_ATT_POLY_EVERY == SASVQA_ATT_POLY_EVERY - 1) {
            exp2_poly_x2(p0, p1);
        } else {
            p0 = ex2(p0);
            p1 = ex2(p1);
        }
        if (2 * j + 1 >= VALID) p1 = 0.f;
        acc[j & 1] = add_f32x2(acc[j & 1], pack_f32x2(p0, p1));
        pk[j] = pack_bf16x2(p0, p1);
    }
    tmem_st16(trow + P_COL, pk);
    tmem_st16(trow + P_COL + 16u, pk + 16);
    tmem_st16(trow + P_COL + 32u, pk + 32);
    if (PART == 0) tmem_st8(trow + 48u, pk + 48);
    float l0, l1;
    unpack_f32x2(add_f32x2(acc[0], acc[1]), l0, l1);
    tmem_wait_st();
    return l0 + l1;
}

// S = Q K^T : M=128, N=208, A and B K-major, bf16 x bf16 -> f32
constexpr uint32_t kIdescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KEYS >> 3) << 17) | ((128u >> 4) << 24);
// O = P V   : M=128, N=64, A (TMEM) K-major, B MN-major (bit 16): V rows are keys with d contiguous
constexpr uint32_t kIdescPV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// Development aid (build with -DSASVQA_ATT_TRACE): CTA 0 stamps clock64() at the protocol events of its first
// items (role 0 = MMA issuer, 1 = softmax warp 6, 2 = epilogue warp 14); launch_attention_tcgen05 prints the
// table.  Compiled out of the product library.
#ifdef SASVQA_ATT_TRACE
__device__ long long g_att_trace[16 * 3 * 16];
#define TR(role, ev)                                                                                   \
    do {                                                                                               \
        if (blockIdx.x == 0 && lane == 0 && it < 16) g_att_trace[(it * 3 + (role)) * 16 + (ev)] = clock64(); \
    } while (0)
#else
#define TR(role, ev) do { } while (0)
#endif

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                         const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, int n_items,
                         int variant) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t v_base = smem_base + 2 * QK_BYTES;
    const uint32_t stage_base = v_base + 2 * KV_BYTES;
    const uint32_t xmax_base = stage_base + STAGE_BYTES;
    const uint32_t xsum_base = xmax_base + XMAX_BYTES;
    const uint32_t bar_base = xsum_base + XSUM_BYTES;
    auto qk_full = [&](int b) { return bar_base + 8u * b; };
    auto qk_empty = [&](int b) { return bar_base + 8u * (2 + b); };
    auto v_full = [&](int b) { return bar_base + 8u * (4 + b); };
    auto v_empty = [&](int b) { return bar_base + 8u * (6 + b); };
    auto s_full = [&](int h) { return bar_base + 8u * (8 + h); };
    auto p_full = [&](int h) { return bar_base + 8u * (10 + h); };
    auto o_full = [&](int h) { return bar_base + 8u * (12 + h); };
    auto o_empty = [&](int h) { return bar_base + 8u * (14 + h); };
    const uint32_t tmem_slot = bar_base + 8u * 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(qk_full(b), 1);
            mbar_init(qk_empty(b), 2);               // one commit per MMA issuer
            mbar_init(v_full(b), 1);
            mbar_init(v_empty(b), 2);
            mbar_init(s_full(b), 1);
            mbar_init(p_full(b), 256);
            mbar_init(o_full(b), 1);
            mbar_init(o_empty(b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        setmaxnreg_dec<REGS_CTRL>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = it & 1;
                const uint32_t free_par = ((it >> 1) & 1u) ^ 1u;
                const int frame = item / kHeads, head = item - frame * kHeads;
                const int row = frame * kTokens;
                const uint32_t qk = smem_base + b * QK_BYTES;
                mbar_wait(qk_empty(b), free_par);
                mbar_arrive_expect_tx(qk_full(b), QK_BYTES);
                tma_load_2d(qk, &map_q, head * kHeadDim, row, qk_full(b));
                tma_load_2d(qk + Q_HALF_BYTES, &map_q, head * kHeadDim, row + 128, qk_full(b));
                tma_load_2d(qk + Q_BYTES, &map_kv, kHidden + head * kHeadDim, row, qk_full(b));
                mbar_wait(v_empty(b), free_par);
                mbar_arrive_expect_tx(v_full(b), KV_BYTES);
                tma_load_2d(v_base + b * KV_BYTES, &map_kv, 2 * kHidden + head * kHeadDim, row, v_full(b));
            }
        } else if ((warp == 1 || warp == 2) && lane == 0) {
            // ===================== MMA issuers: warp 1 drives query half 0, warp 2 half 1 =====================
            // Two issuing threads so that neither half's chain  P_h ready -> P_h V -> slot free -> S_h(next)  waits
            // behind the other half's barriers in program order (one thread cost ~650 cycles per item in such waits).
            const int h = warp - 1;
            auto issue_s = [&](uint32_t qk) {                  // S_h = Q_h K^T into slot h
                const uint64_t adesc = desc_sw128(qk + h * Q_HALF_BYTES, 0);
                const uint64_t bdesc = desc_sw128(qk + Q_BYTES, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_ss(tmem_base + (uint32_t)(h * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdescS,
                           k != 0);
                tcgen05_commit(s_full(h));
            };
            auto issue_pv = [&](uint32_t v_smem) {             // O_h = P_h V, two alternating accumulator chains
                const uint64_t vdesc = desc_sw128(v_smem, KEYS * 128);
                const uint32_t slot = tmem_base + (uint32_t)(h * 256);
                // 16 keys per UMMA = 8 packed-bf16 TMEM columns of P = 2048 B of V
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    mma_ts(slot + OA_COL, slot + (uint32_t)(8 * j), vdesc + (uint64_t)(128 * j), kIdescPV, j != 0);
                    if (j < 6)
                        mma_ts(slot + OB_COL, slot + (uint32_t)(8 * (7 + j)), vdesc + (uint64_t)(128 * (7 + j)), kIdescPV,
                               j != 0);
                }
                tcgen05_commit(o_full(h));
            };
            uint32_t it = 0;
            if ((int)blockIdx.x < n_items) {                    // prologue: scores of the first item
                mbar_wait(qk_full(0), 0);
                tcgen05_fence_after();
                issue_s(smem_base);
                tcgen05_commit(qk_empty(0));
            }
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = it & 1, nb = b ^ 1;
                const uint32_t par = it & 1u;
                const bool has_next = item + (int)gridDim.x < n_items;
                mbar_wait(v_full(b), (it >> 1) & 1u);
                mbar_wait(p_full(h), par);                      // softmax wrote P_h into TMEM
                tcgen05_fence_after();
                TR(0, 2 * h);
                issue_pv(v_base + b * KV_BYTES);
                tcgen05_commit(v_empty(b));                     // (both issuers) V smem of this item reusable
                if (has_next) {
                    mbar_wait(qk_full(nb), ((it + 1) >> 1) & 1u);
                    mbar_wait(o_empty(h), par);                 // epilogue holds O_h in registers: slot h reusable
                    tcgen05_fence_after();
                    TR(0, 2 * h + 1);
                    issue_s(smem_base + nb * QK_BYTES);
                    tcgen05_commit(qk_empty(nb));               // (both issuers) Q|K smem of the next item reusable
                }
            }
        }
    } else if (warp < 12) {
        // ===================== softmax =====================
        setmaxnreg_inc<REGS_SOFTMAX>();
        const int quarter = warp & 3;                           // TMEM lane quarter this warp may touch
        const int part = (warp - 4) >> 2;                       // 0: keys [0,112)   1: keys [112,208)
        const int lrow = quarter * 32 + lane;                   // row inside the query half
        const uint32_t pair_bar = 1u + (uint32_t)quarter;       // named barrier shared by warps (w, w+4)
        float* xmax = reinterpret_cast<float*>(smem_raw + (xmax_base - smem_u32(smem_raw)));
        float* xsum = reinterpret_cast<float*>(smem_raw + (xsum_base - smem_u32(smem_raw)));
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            for (int h = 0; h < 2; ++h) {
                const bool active = (h * 128 + quarter * 32) < kTokens;     // warp-uniform, same for both parts
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * 256);
                mbar_wait(s_full(h), par);
                tcgen05_fence_after();
                if (warp == 6) TR(1, 2 * h);
                if (active && !(variant & 8)) {                 // (variant & 8: timing experiment, protocol only)
                    auto exchange_max = [&](float mx) {
                        xmax[(h * 2 + part) * 128 + lrow] = mx;
                        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                        return fmaxf(mx, xmax[(h * 2 + (part ^ 1)) * 128 + lrow]);
                    };
                    const float l = part == 0 ? softmax_part<0>(trow, exchange_max) : softmax_part<1>(trow, exchange_max);
                    xsum[((par * 2 + h) * 2 + part) * 128 + lrow] = l;
                }
                if (warp == 6) TR(1, 2 * h + 1);
                tcgen05_fence_before();
                mbar_arrive(p_full(h));
            }
        }
    } else {
        // ===================== epilogue =====================
        setmaxnreg_dec<REGS_EPILOGUE>();
        const int quarter = warp & 3;
        const int lrow = quarter * 32 + lane;
        const float* xsum = reinterpret_cast<const float*>(smem_raw + (xsum_base - smem_u32(smem_raw)));
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            const int frame = item / kHeads, head = item - frame * kHeads;
            for (int h = 0; h < 2; ++h) {
                const int row0 = h * 128 + quarter * 32;        // first query row of this warp
                const bool active = row0 < kTokens;
                const bool full_slab = row0 + 32 <= kTokens;    // all 32 rows are real tokens -> TMA store
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * 256);
                const uint32_t slab = stage_base + (uint32_t)((h * 4 + quarter) * SLAB_BYTES);
                mbar_wait(p_full(h), par);                      // the softmax warps' row sums are in smem
                mbar_wait(o_full(h), par);
                tcgen05_fence_after();
                if (warp == 14) TR(2, 3 * h);
                float o[64];
                if (active) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {               // two rounds of 32 columns keep the register peak low
                        uint32_t oa[32], ob[32];
                        tmem_ld32(trow + (uint32_t)(OA_COL + 32 * c), oa);
                        tmem_ld32(trow + (uint32_t)(OB_COL + 32 * c), ob);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[32 * c + j] = __uint_as_float(oa[j]) + __uint_as_float(ob[j]);
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(o_empty(h));                        // O_h is in registers: the slot may be overwritten
                if (warp == 14) TR(2, 3 * h + 1);
                if (active) {
                    const float* xs = xsum + (par * 2 + h) * 2 * 128 + lrow;
                    const float inv_l = 1.0f / (xs[0] + xs[128]);
                    uint32_t w[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[j] = pack_bf16x2(o[2 * j] * inv_l, o[2 * j + 1] * inv_l);
                    if (full_slab) {
                        if (lane == 0) bulk_wait_read_all();    // the store that last read this slab is done with it
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 8; ++q)             // 16-byte chunk q of row `lane`, 128B swizzle
                            st_shared_v4(slab + (uint32_t)(lane * 128 + ((q ^ (lane & 7)) << 4)), w[4 * q], w[4 * q + 1],
                                         w[4 * q + 2], w[4 * q + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && !(variant & 16)) {
                            tma_store_2d(&map_out, slab, head * kHeadDim, frame * kTokens + row0);
                            bulk_commit();
                        }
                    } else if (row0 + lane < kTokens && !(variant & 16)) {   // the 5 tail rows of a frame
                        uint4* dst =
                            reinterpret_cast<uint4*>(out + ((size_t)frame * kTokens + row0 + lane) * kHidden + head * kHeadDim);
#pragma unroll
                        for (int q = 0; q < 8; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                    }
                }
                if (warp == 14) TR(2, 3 * h + 2);
            }
        }
        if (lane == 0) bulk_wait_all();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

// qkv viewed as bf16 [rows, 2304]; boxes of 64 columns (one head of q, k or v) x 128 rows (Q) / 208 rows (K, V)
// out viewed as bf16 [rows, 768]: boxes of 64 columns x 32 rows (one head of one warp pair's query rows)
int make_attention_maps(CUtensorMap* map_q, CUtensorMap* map_kv, CUtensorMap* map_out, const void* qkv, const void* out,
                        uint64_t rows) {
    int rc = make_tensor_map_bf16_kmajor(map_q, qkv, rows, kQkv, 128);
    if (rc) return rc;
    if ((rc = make_tensor_map_bf16_kmajor(map_kv, qkv, rows, kQkv, KEYS))) return rc;
    return make_tensor_map_out(map_out, out, rows, kHidden, 0);
}

int launch_attention_tcgen05(const CUtensorMap* map_q, const CUtensorMap* map_kv, const CUtensorMap* map_out,
                             __nv_bfloat16* out, int n_frames, int num_sms, cudaStream_t s, int variant) {
    if (n_frames == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)out & 15) == 0, "unaligned attention output");
    static SmemAttrCache smem_attr;
    if (int rc = smem_attr.ensure(attention_tcgen05_kernel, ATT_SMEM)) return rc;
    const int n_items = n_frames * kHeads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    attention_tcgen05_kernel<<<grid, ATT_THREADS, ATT_SMEM, s>>>(*map_q, *map_kv, *map_out, out, n_items, variant);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
#ifdef SASVQA_ATT_TRACE
    {
        static int shots = 0;
        if (shots++ == 3) {
            long long t[16 * 3 * 16];
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(t, g_att_trace, sizeof(t));
            const long long t0 = t[(2 * 3 + 1) * 16 + 0];
            auto T = [&](int it, int role, int ev) { return t[(it * 3 + role) * 16 + ev] - t0; };
            for (int it = 2; it < 8; ++it) {
                printf("item %d  mma: PV0 %lld S0' %lld PV1 %lld S1' %lld\n", it, T(it, 0, 0), T(it, 0, 1), T(it, 0, 2), T(it, 0, 3));
                printf("        softmax: s0 %lld p0 %lld | s1 %lld p1 %lld\n", T(it, 1, 0), T(it, 1, 1), T(it, 1, 2), T(it, 1, 3));
                printf("        epilogue: o0 %lld free0 %lld end0 %lld | o1 %lld free1 %lld end1 %lld\n", T(it, 2, 0), T(it, 2, 1),
                       T(it, 2, 2), T(it, 2, 3), T(it, 2, 4), T(it, 2, 5));
            }
        }
    }
#endif
    return 0;
}

}  // namespace sasvqa
