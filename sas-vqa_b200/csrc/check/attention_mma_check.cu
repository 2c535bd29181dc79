// mma.sync CHECK kernel for the encoder's per-frame attention (197 tokens, 12 heads x 64): the first, pre-tcgen05
// version of the kernel, kept OUT of the product library (libsasvqa_b200_test.so only) so tests can tell a
// tcgen05 / TMEM / TMA-descriptor bug in attention_tcgen05.cu from a bug elsewhere.
// softmax(Q K^T / 8) V with fp32 softmax statistics (HF eager_attention_forward, transformers modeling_git.py:556-575).
// One CTA per (frame, head): K and V (197x64 bf16 each, XOR-swizzled 16-byte chunks) staged in shared memory with
// cp.async, 7 warps each own 16-query tiles and run the online softmax over key chunks 64/64/64/16.
#include "../attention_mma.cuh"

namespace sasvqa {

namespace {

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    uint8_t* k_tile = att_smem;
    uint8_t* v_tile = att_smem + KEYS_PAD * 128;
    const int head = blockIdx.x % kHeads;
    const long long frame = blockIdx.x / kHeads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const __nv_bfloat16* base = qkv + frame * kTokens * (long long)kQkv + head * kHeadDim;
    const uint32_t k_smem = (uint32_t)__cvta_generic_to_shared(k_tile);
    const uint32_t v_smem = (uint32_t)__cvta_generic_to_shared(v_tile);

    // ---- stage K and V (zero the padded rows: P is 0 there but 0 * garbage could be NaN)
    for (int i = threadIdx.x; i < KEYS_PAD * 8; i += ATT_THREADS) {
        const int r = i >> 3, c = i & 7;
        if (r < kTokens) {
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
    for (int qt = warp; qt < Q_TILES; qt += ATT_WARPS) {
        const int row0 = qt * 16 + g, row1 = row0 + 8;
        const int r0c = min(row0, kTokens - 1), r1c = min(row1, kTokens - 1);
        // ---- Q fragments (A operand, 16 queries x 64 d) straight from global
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

        attend_chunk<8>(qf, k_smem, v_smem, 0, lane, m, l, o);
        attend_chunk<8>(qf, k_smem, v_smem, 64, lane, m, l, o);
        attend_chunk<8>(qf, k_smem, v_smem, 128, lane, m, l, o);
        attend_chunk<2>(qf, k_smem, v_smem, 192, lane, m, l, o);

#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
        }
        const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
        __nv_bfloat16* o0 = out + (frame * kTokens + row0) * (long long)kHidden + head * kHeadDim + 2 * t;
        __nv_bfloat16* o1 = out + (frame * kTokens + row1) * (long long)kHidden + head * kHeadDim + 2 * t;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            if (row0 < kTokens) *reinterpret_cast<uint32_t*>(o0 + dt * 8) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
            if (row1 < kTokens) *reinterpret_cast<uint32_t*>(o1 + dt * 8) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
        }
    }
}

}  // namespace

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_frames, cudaStream_t s) {
    if (n_frames == 0) return 0;
    SASVQA_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 3) == 0, "unaligned buffers");
    constexpr int smem = 2 * KEYS_PAD * 128;   // 53 248 B: above the 48 KiB static limit
    static SmemAttrCache smem_attr;
    if (int rc = smem_attr.ensure(attention_kernel, smem)) return rc;
    attention_kernel<<<n_frames * kHeads, ATT_THREADS, smem, s>>>(qkv, out);
    SASVQA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace sasvqa
