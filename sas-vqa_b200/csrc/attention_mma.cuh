// mma.sync building blocks of the small attention kernels (bf16 m16n8k16, fp32 softmax statistics): fragment loads,
// the XOR-swizzled [rows][64] shared-memory tile, and one flash-style key chunk with a per-row key limit.
// Used by attention.cu (the MIF scorer's short / variable-length attention, the GIT decoder's text-row steps) and by the
// check kernel in check/attention_mma_check.cu.  The encoder's own attention is attention_tcgen05.cu.
#pragma once

#include "common.cuh"

namespace sasvqa {

namespace {

constexpr int ATT_WARPS = 7;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int KEYS_PAD = 208;                       // 197 rounded up to 16
constexpr int Q_TILES = (kTokens + 15) / 16;        // 13
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {          // one MUFU op, no range fix-ups (arguments are <= 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` in a [rows][64] bf16 tile, XOR-swizzled so
// that ldmatrix (8 rows x 16 B) is bank-conflict free
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// one key chunk of NT*8 keys starting at key0: S = Q K^T, online softmax update, O += P V
template <int NT>
__device__ __forceinline__ void attend_chunk(const uint32_t (&qf)[4][4], uint32_t k_smem, uint32_t v_smem, int key0,
                                             int lane, float (&m)[2], float (&l)[2], float (&o)[8][4],
                                             int n_valid = kTokens, int n_valid_hi = -1) {
    // n_valid: keys [0, n_valid) are visible to row g of the tile; n_valid_hi (default: the same) to row g + 8
    if (n_valid_hi < 0) n_valid_hi = n_valid;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    // ---- S = Q K^T : B fragments straight from K rows (key-major, d contiguous)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int key = key0 + nt * 8 + (lane & 7);
#pragma unroll
        for (int half = 0; half < 2; ++half) {          // d 0..31, d 32..63
            uint32_t b0, b1, b2, b3;
            ldsm_x4(k_smem + tile_off(key, half * 4 + (lane >> 3)), b0, b1, b2, b3);
            mma_bf16(s[nt], qf[half * 2 + 0], b0, b1);
            mma_bf16(s[nt], qf[half * 2 + 1], b2, b3);
        }
    }
    // ---- mask invisible keys (only chunks that reach past a row's limit pay for it), chunk row maxima of the RAW
    // scores (rows g and g+8); the 1/8 * log2(e) scale is folded into the exponent's FFMA below
    float cmax[2] = {-INFINITY, -INFINITY};
    if (key0 + NT * 8 > min(n_valid, n_valid_hi)) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int kcol = key0 + nt * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool valid = (kcol + (e & 1)) < ((e >> 1) ? n_valid_hi : n_valid);
                s[nt][e] = valid ? s[nt][e] : -INFINITY;
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        cmax[0] = fmaxf(cmax[0], fmaxf(s[nt][0], s[nt][1]));
        cmax[1] = fmaxf(cmax[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 1));
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 2));
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mnew[r] = fmaxf(m[r], cmax[r] * kScaleLog2);  // finite: a row's FIRST chunk always holds >= 1 visible key
        alpha[r] = ex2_approx(m[r] - mnew[r]);        // first chunk: exp2(-inf) = 0
        m[r] = mnew[r];
        l[r] *= alpha[r];
    }
    // rescale the accumulator only when some row's maximum moved (alpha == 1 exactly otherwise: a bit-identical skip;
    // after the first few chunks of a long key range that is most of the time)
    if (__any_sync(0xffffffffu, alpha[0] != 1.0f || alpha[1] != 1.0f)) {
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            o[dt][0] *= alpha[0];
            o[dt][1] *= alpha[0];
            o[dt][2] *= alpha[1];
            o[dt][3] *= alpha[1];
        }
    }
    // ---- P = exp2(S * scale - m), packed to bf16 A fragments; O += P V
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pa[4];
        float p[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                p[j][e] = ex2_approx(fmaf(s[2 * kk + j][e], kScaleLog2, -mnew[e >> 1]));   // masked: fma(-inf) = -inf -> 0
                l[e >> 1] += p[j][e];
            }
        }
        pa[0] = pack_bf16x2(p[0][0], p[0][1]);
        pa[1] = pack_bf16x2(p[0][2], p[0][3]);
        pa[2] = pack_bf16x2(p[1][0], p[1][1]);
        pa[3] = pack_bf16x2(p[1][2], p[1][3]);
        const int vrow = key0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {                // two 8-wide d tiles per ldmatrix.x4
            uint32_t b0, b1, b2, b3;
            ldsm_x4_trans(v_smem + tile_off(vrow, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
            mma_bf16(o[dp * 2 + 0], pa, b0, b1);
            mma_bf16(o[dp * 2 + 1], pa, b2, b3);
        }
    }
}

}  // namespace

}  // namespace sasvqa
