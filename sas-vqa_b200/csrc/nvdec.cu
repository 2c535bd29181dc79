// Row f3, decode front end: NVDEC in front of K0.
//
// The reference decodes every video on ONE host thread with cv2 (ffmpeg) -- `cap.read()` per frame, BGR -> RGB swap,
// keep every `intv`-th frame (src/preprocessing/prefetch_loader.py:57-67) -- which is its true end-to-end bottleneck
// (SURVEY.md 8(f) f3).  Here the GPU's video engine decodes an elementary stream (H.264 / HEVC Annex B, e.g. what
// `ffmpeg -c copy -bsf:v h264_mp4toannexb` extracts from the dataset's .avi / .mp4 files; container demuxing stays on the
// host) straight into the uint8 [T, H, W, 3] RGB layout K0 / K1 read, so decoded frames never visit host memory.
//
// libnvcuvid.so.1 ships with the driver; the Video Codec SDK headers do not ship with CUDA and are absent from this image,
// so the handful of entry points and structs used are declared by hand below (NVIDIA Video Codec SDK interface,
// nvcuvid.h / cuviddec.h; layouts are ABI-stable since SDK 9) and the library is dlopen()ed on first use -- the product
// library does not link against it and every other entry point works without it.
//
// Colour: NV12 -> RGB with the BT.601 limited-range integer matrix (298/409/-100/-208/516, +128 >> 8), chroma taken from
// the co-sited 2x2 block.  Decode itself is bit-exact by the codec standard (tests: a lossless I_PCM stream comes back
// identical); the colour conversion is "parity unpinned" against cv2's swscale path, which cannot be run offline here.
#include <dlfcn.h>

#include <mutex>
#include <string>

#include "../../include/sasvqa.h"
#include "common.cuh"

namespace sasvqa {

namespace {

// ---- hand-declared nvcuvid interface --------------------------------------------------------------------------------
typedef void* CUvideoparser;
typedef void* CUvideodecoder;
typedef void* CUvideoctxlock;
typedef long long CUvideotimestamp;

struct NvFormat {                       // CUVIDEOFORMAT (leading fields)
    int codec;
    struct { unsigned int numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
    unsigned int coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned int bitrate;
    struct { int x, y; } display_aspect_ratio;
    unsigned char video_signal_description[4];
    unsigned int seqhdr_data_length;
};
struct NvCreateInfo {                   // CUVIDDECODECREATEINFO
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
    struct { short left, top, right, bottom; } display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    CUvideoctxlock vidLock;
    struct { short left, top, right, bottom; } target_rect;
    unsigned long enableHistogram;
    unsigned long Reserved2[4];
};
struct NvDispInfo {                     // CUVIDPARSERDISPINFO
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    CUvideotimestamp timestamp;
};
struct NvProcParams {                   // CUVIDPROCPARAMS (padded generously: trailing zeros are reserved fields)
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned int reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned int raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned int raw_output_pitch, Reserved1;
    CUstream output_stream;
    unsigned int Reserved[46];
    unsigned long long* histogram_dptr;
    void* Reserved2[1];
    unsigned char tail_padding[64];
};
struct NvPacket {                       // CUVIDSOURCEDATAPACKET
    unsigned long flags, payload_size;
    const unsigned char* payload;
    CUvideotimestamp timestamp;
};
typedef int (*SeqCb)(void*, NvFormat*);
typedef int (*DecCb)(void*, void* /* CUVIDPICPARAMS*, passed through untouched */);
typedef int (*DispCb)(void*, NvDispInfo*);
struct NvParserParams {                 // CUVIDPARSERPARAMS
    int CodecType;
    unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned int flags_annexb_reserved;
    unsigned int uReserved1[4];
    void* pUserData;
    SeqCb pfnSequenceCallback;
    DecCb pfnDecodePicture;
    DispCb pfnDisplayPicture;
    void* pvReserved2[7];               // pfnGetOperatingPoint, pfnGetSEIMsg and reserved slots (unused: NULL)
    void* pExtVideoInfo;
};
constexpr unsigned long kPktEndOfStream = 0x01;
constexpr int kSurfaceNV12 = 0, kDeinterlaceWeave = 0, kChroma420 = 1;
constexpr unsigned long kCreatePreferCUVID = 0x04;

struct NvApi {
    void* handle = nullptr;
    int (*CreateVideoParser)(CUvideoparser*, NvParserParams*) = nullptr;
    int (*ParseVideoData)(CUvideoparser, NvPacket*) = nullptr;
    int (*DestroyVideoParser)(CUvideoparser) = nullptr;
    int (*CreateDecoder)(CUvideodecoder*, NvCreateInfo*) = nullptr;
    int (*DestroyDecoder)(CUvideodecoder) = nullptr;
    int (*DecodePicture)(CUvideodecoder, void*) = nullptr;
    int (*MapVideoFrame64)(CUvideodecoder, int, unsigned long long*, unsigned int*, NvProcParams*) = nullptr;
    int (*UnmapVideoFrame64)(CUvideodecoder, unsigned long long) = nullptr;
    std::string error;
};

NvApi* nv_api() {
    static NvApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.handle = dlopen("libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) {
            api.error = std::string("libnvcuvid.so.1 not loadable (NVDEC needs the driver's video library): ") + dlerror();
            return;
        }
        auto sym = [&](const char* name) {
            void* p = dlsym(api.handle, name);
            if (!p && api.error.empty()) api.error = std::string("libnvcuvid.so.1 lacks ") + name;
            return p;
        };
        api.CreateVideoParser = reinterpret_cast<decltype(api.CreateVideoParser)>(sym("cuvidCreateVideoParser"));
        api.ParseVideoData = reinterpret_cast<decltype(api.ParseVideoData)>(sym("cuvidParseVideoData"));
        api.DestroyVideoParser = reinterpret_cast<decltype(api.DestroyVideoParser)>(sym("cuvidDestroyVideoParser"));
        api.CreateDecoder = reinterpret_cast<decltype(api.CreateDecoder)>(sym("cuvidCreateDecoder"));
        api.DestroyDecoder = reinterpret_cast<decltype(api.DestroyDecoder)>(sym("cuvidDestroyDecoder"));
        api.DecodePicture = reinterpret_cast<decltype(api.DecodePicture)>(sym("cuvidDecodePicture"));
        api.MapVideoFrame64 = reinterpret_cast<decltype(api.MapVideoFrame64)>(sym("cuvidMapVideoFrame64"));
        api.UnmapVideoFrame64 = reinterpret_cast<decltype(api.UnmapVideoFrame64)>(sym("cuvidUnmapVideoFrame64"));
    });
    return &api;
}

// ---- NV12 -> RGB HWC (or a plain NV12 copy), one thread per two horizontally adjacent pixels -------------------------
__device__ __forceinline__ uint8_t clip_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

__global__ void nv12_to_rgb_kernel(const uint8_t* __restrict__ src, unsigned int pitch, int chroma_row0, int H, int W,
                                   uint8_t* __restrict__ dst) {
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (2 * x2 >= W) return;
    const uint8_t* yrow = src + (size_t)y * pitch;
    const uint8_t* crow = src + (size_t)(chroma_row0 + (y >> 1)) * pitch;
    const int d = (int)crow[2 * x2] - 128, e = (int)crow[2 * x2 + 1] - 128;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = 2 * x2 + k;
        if (x >= W) break;
        const int c = 298 * ((int)yrow[x] - 16) + 128;
        uint8_t* o = dst + ((size_t)y * W + x) * 3;
        o[0] = clip_u8((c + 409 * e) >> 8);
        o[1] = clip_u8((c - 100 * d - 208 * e) >> 8);
        o[2] = clip_u8((c + 516 * d) >> 8);
    }
}
__global__ void nv12_copy_kernel(const uint8_t* __restrict__ src, unsigned int pitch, int chroma_row0, int H, int W,
                                 uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;           // y in [0, H * 3 / 2)
    if (x >= W) return;
    const int srow = y < H ? y : chroma_row0 + (y - H);
    dst[(size_t)y * W + x] = src[(size_t)srow * pitch + x];
}

struct DecodeCtx {
    NvApi* api;
    CUvideodecoder dec = nullptr;
    bool probe;
    int intv, capacity, H, W, format;
    uint8_t* out;
    cudaStream_t stream;
    int width = 0, height = 0, seen = 0, kept = 0;
    unsigned long surf_height = 0;
    std::string error;
};

int on_sequence(void* user, NvFormat* f) {
    DecodeCtx* c = static_cast<DecodeCtx*>(user);
    const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
    if (c->width != 0 && (w != c->width || h != c->height)) {
        c->error = "the stream changes its frame size";
        return 0;
    }
    c->width = w;
    c->height = h;
    const int n_surf = f->min_num_decode_surfaces > 4 ? f->min_num_decode_surfaces : 4;
    if (c->probe) return n_surf;
    if (f->chroma_format != kChroma420 || f->bit_depth_luma_minus8 != 0) {
        c->error = "only 8-bit 4:2:0 streams are supported";
        return 0;
    }
    if (w != c->W || h != c->H) {
        c->error = "decoded frame size " + std::to_string(w) + "x" + std::to_string(h) + " differs from the output buffer's";
        return 0;
    }
    if (c->dec) return n_surf;                                                   // repeated sequence header, same size
    NvCreateInfo ci{};
    ci.ulWidth = f->coded_width;
    ci.ulHeight = f->coded_height;
    ci.ulNumDecodeSurfaces = (unsigned long)n_surf;
    ci.CodecType = f->codec;
    ci.ChromaFormat = f->chroma_format;
    ci.ulCreationFlags = kCreatePreferCUVID;
    ci.bitDepthMinus8 = 0;
    ci.ulMaxWidth = f->coded_width;
    ci.ulMaxHeight = f->coded_height;
    ci.display_area.left = (short)f->display_area.left;
    ci.display_area.top = (short)f->display_area.top;
    ci.display_area.right = (short)f->display_area.right;
    ci.display_area.bottom = (short)f->display_area.bottom;
    ci.OutputFormat = kSurfaceNV12;
    ci.DeinterlaceMode = kDeinterlaceWeave;
    ci.ulTargetWidth = (unsigned long)w;
    ci.ulTargetHeight = (unsigned long)h;
    ci.ulNumOutputSurfaces = 2;
    c->surf_height = ci.ulTargetHeight;
    const int rc = c->api->CreateDecoder(&c->dec, &ci);
    if (rc != 0) {
        c->error = "cuvidCreateDecoder failed with CUresult " + std::to_string(rc) + " (no NVDEC engine / unsupported stream?)";
        c->dec = nullptr;
        return 0;
    }
    return n_surf;
}

int on_decode(void* user, void* pic_params) {
    DecodeCtx* c = static_cast<DecodeCtx*>(user);
    if (c->probe) return 1;
    if (!c->dec) return 0;
    const int rc = c->api->DecodePicture(c->dec, pic_params);
    if (rc != 0) {
        c->error = "cuvidDecodePicture failed with CUresult " + std::to_string(rc);
        return 0;
    }
    return 1;
}

int on_display(void* user, NvDispInfo* d) {
    DecodeCtx* c = static_cast<DecodeCtx*>(user);
    const int index = c->seen++;
    if (index % c->intv != 0) return 1;                                          // prefetch_loader.py:63
    const int slot = c->kept++;
    if (c->probe) return 1;
    if (slot >= c->capacity) {
        c->error = "more frames than the output buffer holds (probe the stream first)";
        return 0;
    }
    NvProcParams vpp{};
    vpp.progressive_frame = d->progressive_frame;
    vpp.second_field = d->repeat_first_field + 1;
    vpp.top_field_first = d->top_field_first;
    vpp.unpaired_field = d->repeat_first_field < 0;
    vpp.output_stream = c->stream;
    unsigned long long src = 0;
    unsigned int pitch = 0;
    int rc = c->api->MapVideoFrame64(c->dec, d->picture_index, &src, &pitch, &vpp);
    if (rc != 0) {
        c->error = "cuvidMapVideoFrame64 failed with CUresult " + std::to_string(rc);
        return 0;
    }
    const uint8_t* nv12 = reinterpret_cast<const uint8_t*>(src);
    const int chroma_row0 = (int)((c->surf_height + 1) & ~1ul);
    if (c->format == 0) {
        uint8_t* dst = c->out + (size_t)slot * c->H * c->W * 3;
        dim3 grid((unsigned)((c->W / 2 + 1 + 127) / 128), (unsigned)c->H);
        nv12_to_rgb_kernel<<<grid, 128, 0, c->stream>>>(nv12, pitch, chroma_row0, c->H, c->W, dst);
    } else {
        uint8_t* dst = c->out + (size_t)slot * (c->H * 3 / 2) * c->W;
        dim3 grid((unsigned)((c->W + 127) / 128), (unsigned)(c->H * 3 / 2));
        nv12_copy_kernel<<<grid, 128, 0, c->stream>>>(nv12, pitch, chroma_row0, c->H, c->W, dst);
    }
    count_launch();
    const cudaError_t ke = cudaGetLastError();
    const cudaError_t se = cudaStreamSynchronize(c->stream);                     // the surface may be recycled after the unmap
    c->api->UnmapVideoFrame64(c->dec, src);
    if (ke != cudaSuccess || se != cudaSuccess) {
        c->error = std::string("NV12 conversion failed: ") + cudaGetErrorString(ke != cudaSuccess ? ke : se);
        return 0;
    }
    return 1;
}

int run_parser(DecodeCtx& c, const uint8_t* bitstream, uint64_t n_bytes, int codec) {
    NvParserParams pp{};
    pp.CodecType = codec;
    pp.ulMaxNumDecodeSurfaces = 1;                                               // the sequence callback returns the real count
    pp.ulMaxDisplayDelay = 2;                                                    // decode / display pipelining; order is the parser's
    pp.pUserData = &c;
    pp.pfnSequenceCallback = on_sequence;
    pp.pfnDecodePicture = on_decode;
    pp.pfnDisplayPicture = on_display;
    CUvideoparser parser = nullptr;
    int rc = c.api->CreateVideoParser(&parser, &pp);
    if (rc != 0) {
        set_last_error("cuvidCreateVideoParser failed with CUresult " + std::to_string(rc));
        return SASVQA_ERR_CUDA;
    }
    NvPacket pkt{};
    pkt.payload = bitstream;
    pkt.payload_size = (unsigned long)n_bytes;
    pkt.flags = kPktEndOfStream;
    rc = c.api->ParseVideoData(parser, &pkt);
    c.api->DestroyVideoParser(parser);
    if (c.dec) c.api->DestroyDecoder(c.dec);
    c.dec = nullptr;
    if (!c.error.empty()) {
        set_last_error("NVDEC: " + c.error);
        return SASVQA_ERR_INVALID;
    }
    if (rc != 0) {
        set_last_error("cuvidParseVideoData failed with CUresult " + std::to_string(rc));
        return SASVQA_ERR_CUDA;
    }
    return 0;
}

}  // namespace

int video_probe(const uint8_t* bitstream, uint64_t n_bytes, int codec, int intv, int32_t* info) {
    SASVQA_REQUIRE(bitstream != nullptr && n_bytes > 0 && info != nullptr && intv >= 1, "bad arguments");
    NvApi* api = nv_api();
    SASVQA_REQUIRE(api->error.empty(), api->error.c_str());
    SASVQA_CUDA_CHECK(cudaFree(nullptr));                                        // the parser needs a current CUDA context
    DecodeCtx c{};
    c.api = api;
    c.probe = true;
    c.intv = intv;
    if (int rc = run_parser(c, bitstream, n_bytes, codec)) return rc;
    SASVQA_REQUIRE(c.width > 0 && c.height > 0, "no sequence header found in the stream");
    info[0] = c.width;
    info[1] = c.height;
    info[2] = c.kept;                                                            // frames a decode with this intv returns
    info[3] = c.seen;                                                            // frames in the stream
    return 0;
}

int video_decode(const uint8_t* bitstream, uint64_t n_bytes, int codec, int intv, uint8_t* out_dev, int capacity, int H, int W,
                 int format, int32_t* n_frames_out, cudaStream_t stream) {
    SASVQA_REQUIRE(bitstream != nullptr && n_bytes > 0 && intv >= 1 && capacity >= 0, "bad arguments");
    SASVQA_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && (format == 0 || format == 1), "bad frame size / format");
    SASVQA_REQUIRE(capacity == 0 || out_dev != nullptr, "null output buffer");
    NvApi* api = nv_api();
    SASVQA_REQUIRE(api->error.empty(), api->error.c_str());
    SASVQA_CUDA_CHECK(cudaFree(nullptr));
    DecodeCtx c{};
    c.api = api;
    c.probe = false;
    c.intv = intv;
    c.capacity = capacity;
    c.H = H;
    c.W = W;
    c.format = format;
    c.out = out_dev;
    c.stream = stream;
    if (int rc = run_parser(c, bitstream, n_bytes, codec)) return rc;
    if (n_frames_out) *n_frames_out = c.kept;
    return 0;
}

}  // namespace sasvqa
