// K0: shortest-edge bicubic (anti-aliased) resize to 224 + centre crop of decoded uint8 frames, bit-exact
// with the image processor the reference runs on the host for every clip
// (src/preprocessing/prefetch_loader.py:74-75 -> HF CLIPImageProcessor -> torchvision resize(uint8, BICUBIC,
// antialias=True) -> ATen's int16 fixed-point separable resampler, i.e. Pillow's ImagingResample; restated
// in oracle/resize.py).  224x224 input never comes here (resize and crop are identities).
//
// Frames [n, H, W, 3] uint8 -> [n, 224, 224, 3] uint8.  One CTA per (band of 16 output rows, frame):
// phase A streams the source rows the band needs once through a double-buffered shared-memory row segment
// (16-byte coalesced loads of just the columns the crop needs); thread x makes the horizontal pass for its
// output column into a uint8 intermediate in shared memory (the reference also rounds to uint8 between the
// passes); phase B makes the vertical pass over exactly the taps each output row has and the band leaves as
// full 32-byte sectors.  Algorithmic bytes per frame = H * (cropped source width) * 3 in + 150 528 out.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace sasvqa {

namespace {

constexpr int OUT = kImg;            // 224
constexpr int MAX_TAPS = 64;         // filter taps per axis (source/224 up to ~15x)
constexpr int RS_THREADS = 256;

struct AxisPlan {
    int precision = 0;
    int taps = 0;                    // max xsize over the 224 cropped outputs
    std::vector<int32_t> xmin, xsize;       // [224]
    std::vector<int16_t> w;                 // [taps][224]  (tap-major: threads x read consecutive entries)
};

double cubic_aa(double x) {
    const double a = -0.5;
    x = fabs(x);
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
    return 0.0;
}

// weights of one axis for output indices [first, first + 224) of an in_size -> out_size resample
// (ATen _compute_indices_int16_weights_aa / Pillow precompute_coeffs + normalize_coeffs_8bpc, float64)
AxisPlan plan_axis(int in_size, int out_size, int first) {
    AxisPlan p;
    const double scale = (double)in_size / (double)out_size;
    const double support = scale >= 1.0 ? 2.0 * scale : 2.0;
    const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
    std::vector<std::vector<double>> wf(out_size);
    std::vector<int> xmin_all(out_size), xsize_all(out_size);
    double w_max = 0.0;
    for (int i = 0; i < out_size; ++i) {         // the precision depends on the maximum over ALL outputs
        const double center = scale * (i + 0.5);
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        const int xsize = xmax - xmin;
        double total = 0.0;
        wf[i].resize(xsize > 0 ? xsize : 0);
        for (int j = 0; j < xsize; ++j) {
            const double w = cubic_aa((j + xmin - center + 0.5) * invscale);
            wf[i][j] = w;
            total += w;
        }
        for (int j = 0; j < xsize; ++j) {
            if (total != 0.0) wf[i][j] /= total;
            if (wf[i][j] > w_max) w_max = wf[i][j];
        }
        xmin_all[i] = xmin;
        xsize_all[i] = xsize;
    }
    int precision = 0;
    for (; precision < 22; ++precision) {
        const int next_value = (int)(0.5 + w_max * (double)(1 << (precision + 1)));
        if (next_value >= (1 << 15)) break;
    }
    p.precision = precision;
    p.xmin.resize(OUT);
    p.xsize.resize(OUT);
    for (int o = 0; o < OUT; ++o) {
        p.xmin[o] = xmin_all[first + o];
        p.xsize[o] = xsize_all[first + o];
        if (p.xsize[o] > p.taps) p.taps = p.xsize[o];
    }
    p.w.assign((size_t)p.taps * OUT, 0);
    for (int o = 0; o < OUT; ++o)
        for (int j = 0; j < p.xsize[o]; ++j) {
            const double v = wf[first + o][j];
            p.w[(size_t)j * OUT + o] = (int16_t)(int)(v < 0 ? -0.5 + v * (double)(1 << precision) : 0.5 + v * (double)(1 << precision));
        }
    return p;
}

struct DevicePlan {
    int H = 0, W = 0;
    int prec_x = 0, prec_y = 0, taps_x = 0, taps_y = 0;
    int x_lo = 0, seg_cols = 0;      // source columns [x_lo, x_lo + seg_cols) cover every horizontal tap
    int band_rows[3] = {0, 0, 0};    // max source rows touched by a band of 16 / 8 / 4 output rows
    int32_t* xmin = nullptr;         // [224] relative to x_lo
    int32_t* xsize = nullptr;
    int16_t* wx = nullptr;           // [taps_x][224]
    int32_t* ymin = nullptr;         // [224] absolute source rows
    int32_t* ysize = nullptr;
    int16_t* wy = nullptr;           // [224][taps_y]
};

std::mutex g_plan_mutex;
std::map<std::tuple<int, int, int>, DevicePlan> g_plans;      // (device, H, W)

int get_plan(int H, int W, const DevicePlan** out) {
    int dev = 0;
    SASVQA_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto key = std::make_tuple(dev, H, W);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        *out = &it->second;
        return 0;
    }
    // HF get_resize_output_image_size(shortest_edge=224, default_to_square=False) + center_crop
    const int short_side = W <= H ? W : H, long_side = W <= H ? H : W;
    const int new_long = (int)((double)OUT * (double)long_side / (double)short_side);
    const int oh = W <= H ? new_long : OUT, ow = W <= H ? OUT : new_long;
    const int top = (int)((oh - OUT) / 2.0), left = (int)((ow - OUT) / 2.0);
    AxisPlan px = plan_axis(W, ow, left), py = plan_axis(H, oh, top);
    SASVQA_REQUIRE(px.taps <= MAX_TAPS && py.taps <= MAX_TAPS && px.taps > 0 && py.taps > 0,
                   "frame size outside the supported resize range (source / 224 must be below ~15)");
    DevicePlan d;
    d.H = H; d.W = W;
    d.prec_x = px.precision; d.prec_y = py.precision; d.taps_x = px.taps; d.taps_y = py.taps;
    d.x_lo = px.xmin[0];
    int x_hi = 0;
    for (int o = 0; o < OUT; ++o) x_hi = std::max(x_hi, px.xmin[o] + px.xsize[o]);
    d.seg_cols = x_hi - d.x_lo;
    const int bands[3] = {16, 8, 4};
    for (int bi = 0; bi < 3; ++bi)
        for (int r0 = 0; r0 < OUT; r0 += bands[bi]) {
            const int last = r0 + bands[bi] - 1;
            d.band_rows[bi] = std::max(d.band_rows[bi], py.xmin[last] + py.xsize[last] - py.xmin[r0]);
        }
    std::vector<int32_t> xmin_rel(OUT);
    for (int o = 0; o < OUT; ++o) xmin_rel[o] = px.xmin[o] - d.x_lo;
    std::vector<int16_t> wy((size_t)OUT * py.taps, 0);           // row-major [224][taps_y]
    for (int o = 0; o < OUT; ++o)
        for (int j = 0; j < py.xsize[o]; ++j) wy[(size_t)o * py.taps + j] = py.w[(size_t)j * OUT + o];
    auto upload = [](void** dst, const void* src, size_t bytes) -> int {
        SASVQA_CUDA_CHECK(cudaMalloc(dst, bytes));
        SASVQA_CUDA_CHECK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return 0;
    };
    int rc;
    if ((rc = upload((void**)&d.xmin, xmin_rel.data(), OUT * 4))) return rc;
    if ((rc = upload((void**)&d.xsize, px.xsize.data(), OUT * 4))) return rc;
    if ((rc = upload((void**)&d.wx, px.w.data(), px.w.size() * 2))) return rc;
    if ((rc = upload((void**)&d.ymin, py.xmin.data(), OUT * 4))) return rc;
    if ((rc = upload((void**)&d.ysize, py.xsize.data(), OUT * 4))) return rc;
    if ((rc = upload((void**)&d.wy, wy.data(), wy.size() * 2))) return rc;
    *out = &(g_plans[key] = d);
    return 0;
}

__device__ __forceinline__ int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Copies source bytes [g0, g0 + nbytes) of `src` into seg + (g0 & 15): whole 16-byte chunks with vector
// loads, the ragged ends (and anything past the end of the buffer) byte by byte.
__device__ __forceinline__ void load_segment(const uint8_t* __restrict__ src, long long g0, int nbytes, long long total,
                                             uint8_t* seg) {
    const long long a0 = g0 & ~15LL;                             // aligned start (src itself is 16B-aligned)
    const int n_chunks = (int)((g0 + nbytes - a0 + 15) >> 4);
    for (int c = threadIdx.x; c < n_chunks; c += RS_THREADS) {
        const long long a = a0 + 16LL * c;
        if (a + 16 <= total) {
            *reinterpret_cast<uint4*>(seg + 16 * c) = __ldg(reinterpret_cast<const uint4*>(src + a));
        } else {
            for (int b = 0; b < 16; ++b) seg[16 * c + b] = a + b < total ? src[a + b] : (uint8_t)0;
        }
    }
}

// smem layout (bytes): [2 x seg_stride row segments][tmp: max_rows x 672 horizontal results][out: BAND x 672]
//                      [wx: taps_x x 224 int16][wy: BAND x taps_y int16][ymin, ysize: BAND int32 each]
template <int BAND>
__global__ void __launch_bounds__(RS_THREADS)
resize_crop_u8_kernel(const uint8_t* __restrict__ frames, long long total_bytes, int H, int W, int prec_x, int prec_y,
                      int taps_x, int taps_y, int x_lo, int seg_cols, int max_rows, const int32_t* __restrict__ xmin,
                      const int32_t* __restrict__ xsize, const int16_t* __restrict__ wx, const int32_t* __restrict__ ymin,
                      const int32_t* __restrict__ ysize, const int16_t* __restrict__ wy, const int32_t* __restrict__ frame_map,
                      long long src_frame0, int frame0, uint8_t* __restrict__ out) {
    extern __shared__ __align__(32) uint8_t rs_smem[];
    constexpr int ROW_BYTES = OUT * 3;                                             // 672
    const int seg_stride = ((seg_cols * 3 + 15 + 16) + 15) & ~15;                  // bytes per row-segment buffer
    uint8_t* seg0 = rs_smem;
    uint8_t* tmp = rs_smem + 2 * seg_stride;
    uint8_t* obuf = tmp + max_rows * ROW_BYTES;
    int16_t* s_wx = reinterpret_cast<int16_t*>(obuf + BAND * ROW_BYTES);
    int16_t* s_wy = s_wx + taps_x * OUT;
    int32_t* s_ymin = reinterpret_cast<int32_t*>(s_wy + BAND * taps_y + ((BAND * taps_y) & 1));
    int32_t* s_ysize = s_ymin + BAND;

    const int r0 = blockIdx.x * BAND, frame = frame0 + blockIdx.y;
    const int tid = threadIdx.x;
    uint8_t* dst = out + ((size_t)frame * OUT + r0) * ROW_BYTES;                   // the band is contiguous in memory
    const long long src_frame = frame_map ? (long long)frame_map[frame] : src_frame0 + frame;   // < 0: zero rows
    if (src_frame < 0) {
        for (int i = tid; i < BAND * ROW_BYTES / 16; i += RS_THREADS) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    for (int i = tid; i < taps_x * OUT; i += RS_THREADS) s_wx[i] = wx[i];
    for (int i = tid; i < BAND * taps_y; i += RS_THREADS) s_wy[i] = wy[r0 * taps_y + i];
    if (tid < BAND) {
        s_ymin[tid] = ymin[r0 + tid];
        s_ysize[tid] = ysize[r0 + tid];
    }
    const int y_lo = ymin[r0];
    const int y_hi = ymin[r0 + BAND - 1] + ysize[r0 + BAND - 1];                   // ymin / ymin+ysize are monotonic
    const int my_xmin = tid < OUT ? xmin[tid] : 0, my_xsize = tid < OUT ? xsize[tid] : 0;
    const long long frame_base = src_frame * H * W * 3;
    const int seg_bytes = seg_cols * 3;

    // ---- phase A: horizontal pass of every source row the band needs -> uint8 intermediate in shared memory
    long long g = frame_base + ((long long)y_lo * W + x_lo) * 3;
    load_segment(frames, g, seg_bytes, total_bytes, seg0);
    __syncthreads();
    for (int y = y_lo; y < y_hi; ++y) {
        const int buf = (y - y_lo) & 1;
        const uint8_t* seg = seg0 + buf * seg_stride + (int)(g & 15);
        if (y + 1 < y_hi)                                                          // prefetch the next row
            load_segment(frames, g + (long long)W * 3, seg_bytes, total_bytes, seg0 + (buf ^ 1) * seg_stride);
        if (tid < OUT) {
            int h0 = 1 << (prec_x - 1), h1 = h0, h2 = h0;
            const uint8_t* p = seg + my_xmin * 3;
            for (int j = 0; j < my_xsize; ++j) {
                const int w = s_wx[j * OUT + tid];
                h0 += p[3 * j] * w;
                h1 += p[3 * j + 1] * w;
                h2 += p[3 * j + 2] * w;
            }
            uint8_t* t = tmp + (y - y_lo) * ROW_BYTES + tid * 3;
            t[0] = (uint8_t)clamp_u8(h0 >> prec_x);
            t[1] = (uint8_t)clamp_u8(h1 >> prec_x);
            t[2] = (uint8_t)clamp_u8(h2 >> prec_x);
        }
        g += (long long)W * 3;
        __syncthreads();
    }
    // ---- phase B: vertical pass, only the taps each output row really has
    if (tid < OUT) {
#pragma unroll 1
        for (int r = 0; r < BAND; ++r) {
            int v0 = 1 << (prec_y - 1), v1 = v0, v2 = v0;
            const uint8_t* t = tmp + (s_ymin[r] - y_lo) * ROW_BYTES + tid * 3;
            const int16_t* wr = s_wy + r * taps_y;
            const int n = s_ysize[r];
            for (int k = 0; k < n; ++k) {
                const int w = wr[k];
                v0 += t[k * ROW_BYTES] * w;
                v1 += t[k * ROW_BYTES + 1] * w;
                v2 += t[k * ROW_BYTES + 2] * w;
            }
            uint8_t* o = obuf + r * ROW_BYTES + tid * 3;
            o[0] = (uint8_t)clamp_u8(v0 >> prec_y);
            o[1] = (uint8_t)clamp_u8(v1 >> prec_y);
            o[2] = (uint8_t)clamp_u8(v2 >> prec_y);
        }
    }
    __syncthreads();
    for (int i = tid; i < BAND * ROW_BYTES / 32; i += RS_THREADS) {               // full 32-byte sectors
        const uint4 a = reinterpret_cast<const uint4*>(obuf)[2 * i], c = reinterpret_cast<const uint4*>(obuf)[2 * i + 1];
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 32 * i), "r"(a.x), "r"(a.y),
                     "r"(a.z), "r"(a.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w)
                     : "memory");
    }
}

template <int BAND>
int launch_band(const DevicePlan* p, const uint8_t* frames, long long n_src, int H, int W, const int32_t* frame_map,
                long long src_frame0, int n_frames, uint8_t* out, size_t smem, int max_rows, cudaStream_t s) {
    static SmemAttrCache smem_attr;
    if (smem > 48 * 1024)
        if (int rc = smem_attr.ensure(resize_crop_u8_kernel<BAND>, smem)) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {               // gridDim.y limit
        const int n = std::min(65535, n_frames - f0);
        resize_crop_u8_kernel<BAND><<<dim3(OUT / BAND, n), RS_THREADS, smem, s>>>(
            frames, n_src * H * W * 3, H, W, p->prec_x, p->prec_y, p->taps_x, p->taps_y, p->x_lo, p->seg_cols, max_rows,
            p->xmin, p->xsize, p->wx, p->ymin, p->ysize, p->wy, frame_map, src_frame0, f0, out);
        SASVQA_CUDA_CHECK(cudaGetLastError());
        count_launch();
    }
    return 0;
}

}  // namespace

// algorithmic source bytes one frame needs (cropped columns x all rows the vertical taps touch)
long long resize_source_bytes_per_frame(int H, int W) {
    const DevicePlan* p = nullptr;
    if (get_plan(H, W, &p)) return 0;
    return (long long)H * p->seg_cols * 3;
}

// frames [n_src, H, W, 3] -> out [n_out, 224, 224, 3].  frame_map == nullptr: out row i comes from source frame
// src_frame0 + i; else from source frame frame_map[i] (negative: a zero row).
int launch_resize_crop_u8(const uint8_t* frames, long long n_src, int H, int W, const int32_t* frame_map,
                          long long src_frame0, int n_out, uint8_t* out, cudaStream_t s) {
    SASVQA_REQUIRE(n_out >= 0 && n_src >= 0 && H > 0 && W > 0, "bad frame count / size");
    if (n_out == 0) return 0;
    const int n_frames = n_out;
    SASVQA_REQUIRE(frames != nullptr && out != nullptr, "null argument");
    SASVQA_REQUIRE(((uintptr_t)frames & 15) == 0, "frames must be 16-byte aligned");
    SASVQA_REQUIRE(frame_map != nullptr || (src_frame0 >= 0 && src_frame0 + n_out <= n_src), "frame range out of bounds");
    if (H == OUT && W == OUT && frame_map == nullptr) {          // resize and crop are identities
        SASVQA_CUDA_CHECK(cudaMemcpyAsync(out, frames + (size_t)src_frame0 * kFrameElems, (size_t)n_frames * kFrameElems,
                                          cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    const DevicePlan* p = nullptr;
    int rc = get_plan(H, W, &p);
    if (rc) return rc;
    SASVQA_REQUIRE(((uintptr_t)out & 31) == 0, "output must be 32-byte aligned");
    const int seg_stride = ((p->seg_cols * 3 + 15 + 16) + 15) & ~15;
    // rows of source a band of `band` output rows can touch (ymin is monotonic with slope <= H / 224 + 1)
    auto plan_smem = [&](int band, int* max_rows) {
        *max_rows = p->band_rows[band == 16 ? 0 : (band == 8 ? 1 : 2)];
        return 2 * (size_t)seg_stride + (size_t)(*max_rows) * OUT * 3 + (size_t)band * OUT * 3 + (size_t)p->taps_x * OUT * 2 +
               (size_t)(band * p->taps_y + 1) * 2 + 2 * band * 4 + 32;
    };
    int max_rows = 0;
    const size_t kLimit = 200 * 1024;
    size_t smem = plan_smem(16, &max_rows);
    if (smem <= kLimit) return launch_band<16>(p, frames, n_src, H, W, frame_map, src_frame0, n_frames, out, smem, max_rows, s);
    smem = plan_smem(8, &max_rows);
    if (smem <= kLimit) return launch_band<8>(p, frames, n_src, H, W, frame_map, src_frame0, n_frames, out, smem, max_rows, s);
    smem = plan_smem(4, &max_rows);
    SASVQA_REQUIRE(smem <= kLimit, "frame too large for the resize kernel's shared-memory band");
    return launch_band<4>(p, frames, n_src, H, W, frame_map, src_frame0, n_frames, out, smem, max_rows, s);
}

}  // namespace sasvqa
