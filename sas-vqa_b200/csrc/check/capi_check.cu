// extern "C" surface of libsasvqa_b200_test.so: CHECK kernels that are deliberately NOT in the product library
// (libsasvqa_b200.so has one backend per operation and no runtime switch).  The parity tests run the same inputs
// through the product kernel and through these to tell a tcgen05 / TMEM / TMA-descriptor bug from a bug elsewhere.
//   sasvqa_check_gemm_simt      CUDA-core GEMM with the fused epilogues of gemm_tcgen05.cu (modes 0..4)
//   sasvqa_check_attention_mma  mma.sync version of the encoder's per-frame attention
#include <atomic>

#include "../common.cuh"

namespace sasvqa {

static thread_local std::string g_check_error;
void set_last_error(const std::string& msg) { g_check_error = msg; }
void count_launch(int) {}
int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);
int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_frames, cudaStream_t s);

}  // namespace sasvqa

using namespace sasvqa;

extern "C" {

const char* sasvqa_check_last_error(void) { return g_check_error.c_str(); }

int sasvqa_check_gemm_simt(const uint16_t* a, const uint16_t* b, int M, int N, int K, int mode, const float* bias_or_pos,
                           uint16_t* out_bf16, float* out_f32, void* stream) {
    SASVQA_REQUIRE(a && b && mode >= 0 && mode <= 4, "bad arguments");
    GemmArgs g{};
    g.A = reinterpret_cast<const __nv_bfloat16*>(a);
    g.B = reinterpret_cast<const __nv_bfloat16*>(b);
    g.M = M; g.N = N; g.K = K; g.epilogue = mode;
    g.bias = bias_or_pos; g.pos = bias_or_pos;
    g.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    g.out_f32 = out_f32;
    return launch_gemm_simt(g, reinterpret_cast<cudaStream_t>(stream));
}

int sasvqa_check_attention_mma(const uint16_t* qkv, int n_frames, uint16_t* out, void* stream) {
    SASVQA_REQUIRE(n_frames == 0 || (qkv && out), "null argument");
    return launch_attention(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), n_frames,
                            reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
