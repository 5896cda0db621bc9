// cabi.cu -- the reference-facing seams of include/mjpeg423_b200.h:
//   1. library seam   lossless_decode / idct / ycbcr_to_rgb / mjpeg423_decode
//                     (LIB/decoder/mjpeg423_decoder.h:14-17, LIB = .../common/libs/mjpeg423)
//   2. accelerator seam  init_idct_ycbcr_to_rgb_accel, idct_accel_calculate_buffer_{y,cb,cr},
//                     ycbcr_to_rgb_accel_get_results, wait_for_* (C0/idct_ycbcr_to_rgb_accel.h:13-22)
// Every call runs the CUDA kernels (batch of one where the reference API is per block); there is no
// host arithmetic path.  Failures abort like the reference's error_and_exit (LIB/common/util.c:13-16).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "runtime.h"

using namespace mj;

namespace {

std::mutex g_mu;
mjpeg423_b200_ctx* g_ctx = nullptr;
size_t g_read_limit = 0;

[[noreturn]] void die(const char* where) {
    std::fprintf(stderr, "Error: mjpeg423_b200 %s: %s\n", where, mjpeg423_b200_last_error());
    std::abort();
}
#define CUX(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { mj::cuda_fail(e_, #call); return MJPEG423_E_CUDA; }   \
    } while (0)

// Process-wide context for the shims (device 0, or MJPEG423_B200_DEVICE).
mjpeg423_b200_ctx* shim_ctx() {
    if (!g_ctx) {
        const char* env = std::getenv("MJPEG423_B200_DEVICE");
        if (mjpeg423_b200_create(&g_ctx, env ? std::atoi(env) : 0) != MJPEG423_OK) return nullptr;
    }
    cudaSetDevice(g_ctx->device);
    return g_ctx;
}

// Scratch device buffers of the shims.
DevBuf s_in, s_mid, s_out, s_tab, s_seg, s_idx;

int single_stream_decode(mjpeg423_b200_ctx* c, int num_blocks, const void* bitstream, size_t len, int16_t* DCACq,
                         const int16_t* quant, int P) {
    if (num_blocks < 0 || !bitstream || !DCACq || !quant) { set_error("lossless_decode: bad argument"); return MJPEG423_E_ARG; }
    if (num_blocks == 0) return MJPEG423_OK;
    if (len >= MAX_STREAM_BYTES) { set_error("lossless_decode: stream too large"); return MJPEG423_E_ARG; }
    cudaStream_t s = c->s_compute;
    StreamDesc sd{};
    sd.byte_off = 0; sd.byte_len = (uint32_t)len; sd.nb = (uint32_t)num_blocks; sd.seg_base = 0;
    sd.nseg = std::max<uint32_t>(1, ((uint32_t)len + SEG_BYTES - 1) / SEG_BYTES);
    sd.block_base = 0; sd.prev_base = 0; sd.quant_id = 0; sd.ptype = P ? 1 : 0;   // P: accumulate in place
    const size_t coef_bytes = (size_t)num_blocks * 128;
    int rc;
    if ((rc = s_in.reserve(len + 64))) return rc;
    if ((rc = s_tab.reserve(256 + 256))) return rc;
    const uint32_t nseg_pad = (sd.nseg + SUPER - 1) / SUPER * SUPER;      // global segment numbering is SUPER-aligned
    if ((rc = s_seg.reserve((size_t)nseg_pad * 24 + 96))) return rc;
    if ((rc = s_idx.reserve((size_t)num_blocks * 8 + (size_t)nseg_pad * SYM_STRIDE * 4 + 128))) return rc;
    if ((rc = s_mid.reserve(coef_bytes))) return rc;
    uint8_t* tab = s_tab.as<uint8_t>();
    int16_t* d_q = reinterpret_cast<int16_t*>(tab);                  // 128 int16 (table 0 used)
    StreamDesc* d_sd = reinterpret_cast<StreamDesc*>(tab + 256);
    CUX(cudaMemcpyAsync(s_in.p, bitstream, len, cudaMemcpyHostToDevice, s));
    CUX(cudaMemsetAsync(s_in.as<uint8_t>() + len, 0, 64, s));
    CUX(cudaMemcpyAsync(d_q, quant, 128, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(d_sd, &sd, sizeof(sd), cudaMemcpyHostToDevice, s));
    if (P) CUX(cudaMemcpyAsync(s_mid.p, DCACq, coef_bytes, cudaMemcpyHostToDevice, s));   // in/out state
    EntropyJob j;
    j.d_payload = s_in.as<uint8_t>();
    j.d_streams = d_sd;
    j.stream_lo = 0; j.n_streams = 1;
    j.seg_lo = 0; j.seg_hi = nseg_pad;
    j.fold_end = (uint64_t)len * 8 < (uint64_t)num_blocks * 64;
    uint32_t* seg = s_seg.as<uint32_t>();
    CUX(cudaMemsetAsync(seg + 5 * (size_t)nseg_pad, 0, (size_t)nseg_pad * 4, s));   // every segment belongs to stream 0
    j.d_seg_stream = seg + 5 * (size_t)nseg_pad;
    j.d_seg_entry = seg; j.d_seg_exit = seg + nseg_pad; j.d_seg_cnt = seg + 2 * (size_t)nseg_pad;
    j.d_seg_first = seg + 3 * (size_t)nseg_pad;
    j.d_seg_dc = seg + 4 * (size_t)nseg_pad;
    j.d_stream_blocks = seg + 6 * (size_t)nseg_pad;
    j.d_fixups = reinterpret_cast<unsigned long long*>(seg + ((6 * (size_t)nseg_pad + 3) & ~(size_t)1));
    j.d_blk_info = s_idx.as<uint2>();
    j.d_sym = s_idx.as<uint32_t>() + ((2 * (size_t)num_blocks + 7) & ~(size_t)7);
    j.sym_seg0 = 0;
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(tab + 256 + 128);     // one id: stream 0
    CUX(cudaMemsetAsync(d_ids, 0, 4, s));
    CUX(cudaMemsetAsync(j.d_fixups, 0, 16, s));
    CUX(launch_entropy_sync(j, s));
    CUX(launch_entropy_chain(j, s));
    CUX(launch_entropy_index(j, s));
    CUX(launch_decode_coef(j, d_ids, 1, sd.nb, d_q, s_mid.as<int16_t>(), s));
    CUX(cudaMemcpyAsync(DCACq, s_mid.p, coef_bytes, cudaMemcpyDeviceToHost, s));
    CUX(cudaStreamSynchronize(s));
    return MJPEG423_OK;
}

int idct_blocks(mjpeg423_b200_ctx* c, const int16_t* coef, uint8_t* samples, size_t n) {
    if (!coef || !samples) return MJPEG423_E_ARG;
    if (n == 0) return MJPEG423_OK;
    int rc;
    if ((rc = s_mid.reserve(n * 128))) return rc;
    if ((rc = s_out.reserve(n * 64))) return rc;
    cudaStream_t s = c->s_compute;
    CUX(cudaMemcpyAsync(s_mid.p, coef, n * 128, cudaMemcpyHostToDevice, s));
    CUX(launch_idct(s_mid.as<int16_t>(), s_out.as<uint8_t>(), n, s));
    CUX(cudaMemcpyAsync(samples, s_out.p, n * 64, cudaMemcpyDeviceToHost, s));
    CUX(cudaStreamSynchronize(s));
    return MJPEG423_OK;
}

// Y/Cb/Cr block-major planes (nb*64 bytes each) -> W x H BGRA at `rgb` (host).
int colour_frame(mjpeg423_b200_ctx* c, const uint8_t* Y, const uint8_t* Cb, const uint8_t* Cr, uint32_t W, uint32_t H,
                 void* rgb) {
    if (!Y || !Cb || !Cr || !rgb || !W || !H || (W & 7) || (H & 7)) { set_error("ycbcr_to_rgb: bad argument"); return MJPEG423_E_ARG; }
    const size_t nb = (size_t)(W / 8) * (H / 8);
    int rc;
    if ((rc = s_in.reserve(3 * nb * 64))) return rc;
    if ((rc = s_out.reserve(nb * 256))) return rc;
    cudaStream_t s = c->s_compute;
    CUX(cudaMemcpyAsync(s_in.as<uint8_t>(), Y, nb * 64, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(s_in.as<uint8_t>() + nb * 64, Cb, nb * 64, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpyAsync(s_in.as<uint8_t>() + 2 * nb * 64, Cr, nb * 64, cudaMemcpyHostToDevice, s));
    CUX(launch_colour(s_in.as<uint8_t>(), s_out.p, 1, W, H, s));
    CUX(cudaMemcpyAsync(rgb, s_out.p, nb * 256, cudaMemcpyDeviceToHost, s));
    CUX(cudaStreamSynchronize(s));
    return MJPEG423_OK;
}

// Accelerator-seam state (one "device", like the single FPGA block of the reference).
struct Accel {
    bool inited = false;
    uint32_t W = 640, H = 480;                 // COMMON/config.h:23-24
    DevBuf coef, out;
    cudaEvent_t ev_y = nullptr;
} g_accel;

void accel_submit(int plane, void* buf, uint32_t bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c || !g_accel.inited) { set_error("accelerator not initialised"); die("idct_accel_calculate_buffer"); }
    const size_t plane_bytes = (size_t)(g_accel.W / 8) * (g_accel.H / 8) * 128;
    if (!buf || bytes > plane_bytes) { set_error("plane larger than the configured geometry"); die("idct_accel_calculate_buffer"); }
    if (g_accel.coef.reserve(3 * plane_bytes)) die("idct_accel_calculate_buffer");
    cudaError_t e = cudaMemcpyAsync(g_accel.coef.as<uint8_t>() + plane * plane_bytes, buf, bytes, cudaMemcpyHostToDevice,
                                    c->s_compute);
    if (e == cudaSuccess && plane == 0) e = cudaEventRecord(g_accel.ev_y, c->s_compute);
    if (e != cudaSuccess) { cuda_fail(e, "accelerator upload"); die("idct_accel_calculate_buffer"); }
}

void write_bmp32(const char* path, const uint8_t* bgra, uint32_t W, uint32_t H) {
    if (mjpeg423_b200_write_bmp(path, reinterpret_cast<const rgb_pixel_t*>(bgra), W, H) != MJPEG423_OK) die("mjpeg423_decode");
}

}  // namespace

extern "C" {

void mjpeg423_b200_set_read_limit(size_t bytes) { g_read_limit = bytes; }
size_t mjpeg423_b200_get_read_limit(void) { return g_read_limit; }

int mjpeg423_b200_lossless_decode(int num_blocks, const void* bitstream, size_t bitstream_len, int16_t* DCACq,
                                  const int16_t* quant, int P) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c) return MJPEG423_E_CUDA;
    return single_stream_decode(c, num_blocks, bitstream, bitstream_len, DCACq, quant, P);
}

// The reference signature carries no length.  So that the shim touches no byte the reference would not have touched
// itself, the extent of the stream is found first: a walk over the SYMBOL HEADERS of num_blocks blocks with the control
// flow of LIB/decoder/lossless_decode.c:84-133 (4-bit size / run + size nibbles; no amplitude is decoded, no
// coefficient produced -- the decode itself runs on the GPU).  Stops at `limit` bytes.
static size_t stream_extent(int num_blocks, const uint8_t* p, size_t limit) {
    uint64_t bit = 0;
    auto nib = [&](uint64_t b) -> unsigned {             // the 4 bits at bit position b (reads p[i + 1] only when they reach into it)
        const size_t i = (size_t)(b >> 3);
        const unsigned sh = (unsigned)(b & 7);
        const unsigned v = ((unsigned)p[i] << 8) | (sh > 4 ? p[i + 1] : 0u);
        return (v >> (12 - sh)) & 15u;
    };
    for (int blk = 0; blk < num_blocks; blk++) {
        if ((bit >> 3) + 2 > limit) return limit;
        bit += 4 + nib(bit);                             // input_DC, :210-224
        uint8_t index = 1;                               // :100
        for (;;) {
            if ((bit >> 3) + 3 > limit) return limit;
            const unsigned run = nib(bit), size = nib(bit + 4);   // input_AC, :227-246
            bit += 8 + size;
            if (size == 0) {
                if (run == 15) index = (uint8_t)(index + 16);     // ZRL, :106-109
                else break;                                       // END, :110-113
            } else {
                index = (uint8_t)(index + run);
                if (index >= 63) break;                           // :130
                index++;
            }
        }
    }
    const size_t bytes = (size_t)((bit + 7) >> 3);
    return bytes < limit ? bytes : limit;
}

void lossless_decode(int num_blocks, void* bitstream, dct_block_t* DCACq, dct_block_t quant, int P) {
    // Upper bound of the walk: the longest stream the reference decoder accepts (19 + 63 * 23 = 1468 bits = 184 bytes per
    // block) or the configured limit.
    size_t limit = (size_t)(num_blocks > 0 ? num_blocks : 0) * 184 + 8;
    if (g_read_limit && g_read_limit < limit) limit = g_read_limit;
    const size_t len = bitstream ? stream_extent(num_blocks, static_cast<const uint8_t*>(bitstream), limit) : 0;
    if (mjpeg423_b200_lossless_decode(num_blocks, bitstream, len, &DCACq[0][0][0], &quant[0][0], P) != MJPEG423_OK)
        die("lossless_decode");
}

int mjpeg423_b200_idct_blocks(const int16_t* coef, uint8_t* samples, size_t n_blocks) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c) return MJPEG423_E_CUDA;
    return idct_blocks(c, coef, samples, n_blocks);
}

void idct(dct_block_t DCAC, color_block_t block) {
    if (mjpeg423_b200_idct_blocks(&DCAC[0][0], &block[0][0], 1) != MJPEG423_OK) die("idct");
}

int mjpeg423_b200_ycbcr_to_rgb_frame(const uint8_t* Y, const uint8_t* Cb, const uint8_t* Cr, uint32_t w_size,
                                     uint32_t h_size, rgb_pixel_t* rgb) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c) return MJPEG423_E_CUDA;
    return colour_frame(c, Y, Cb, Cr, w_size, h_size, rgb);
}

void ycbcr_to_rgb(int h, int w, uint32_t w_size, pcolor_block_t Y, pcolor_block_t Cb, pcolor_block_t Cr,
                  rgb_pixel_t* rgbblock) {
    rgb_pixel_t tile[64];
    if (mjpeg423_b200_ycbcr_to_rgb_frame(&Y[0][0], &Cb[0][0], &Cr[0][0], 8, 8, tile) != MJPEG423_OK) die("ycbcr_to_rgb");
    for (int y = 0; y < 8; y++)      // place the converted tile at (h, w) of the caller's raster (ycbcr_to_rgb.c:30,46)
        std::memcpy(rgbblock + (size_t)(h + y) * w_size + w, tile + y * 8, 8 * sizeof(rgb_pixel_t));
}

void mjpeg423_decode(const char* filename_in, const char* filenamebase_out) {
    FILE* f = std::fopen(filename_in, "rb");
    if (!f) { set_error("cannot open input file"); die("mjpeg423_decode"); }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> file((size_t)sz);
    if (std::fread(file.data(), 1, (size_t)sz, f) != (size_t)sz) { set_error("cannot read input file"); die("mjpeg423_decode"); }
    std::fclose(f);
    mjpeg423_b200_info info;
    if (mjpeg423_b200_probe(file.data(), file.size(), &info) != MJPEG423_OK) die("mjpeg423_decode");
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c) die("mjpeg423_decode");
    std::string name(filenamebase_out);
    if (name.size() < 8) { set_error("output name must look like name0000.bmp"); die("mjpeg423_decode"); }
    const size_t pos = name.size() - 8;                     // mjpeg423_decoder.c:127
    // Decode whole GOP-aligned batches so P frames always follow their I frame inside a batch.
    MpgIndex idx;
    if (parse_mpg(file.data(), file.size(), idx, false) != MJPEG423_OK) die("mjpeg423_decode");
    const uint32_t max_batch = (uint32_t)std::max<uint64_t>(1, ((uint64_t)256 << 20) / info.frame_bytes);
    uint8_t* host = nullptr;
    size_t host_frames = 0;
    uint32_t f0 = 0;
    while (f0 < info.num_frames) {
        uint32_t f1 = f0 + 1;
        while (f1 < info.num_frames && (idx.frames[f1].type != 0 || f1 - f0 < max_batch)) f1++;
        if (f1 - f0 > host_frames) {
            mjpeg423_b200_host_free(host);
            host_frames = f1 - f0;
            host = (uint8_t*)mjpeg423_b200_host_alloc(host_frames * info.frame_bytes);
            if (!host) { set_error("cannot allocate rgbblock"); die("mjpeg423_decode"); }
        }
        if (mjpeg423_b200_decode_frames(c, file.data(), file.size(), f0, f1 - f0, host, 0) != MJPEG423_OK) die("mjpeg423_decode");
        for (uint32_t k = f0; k < f1; k++) {
            name[pos] = (char)('0' + k / 1000 % 10); name[pos + 1] = (char)('0' + k / 100 % 10);
            name[pos + 2] = (char)('0' + k / 10 % 10); name[pos + 3] = (char)('0' + k % 10);
            write_bmp32(name.c_str(), host + (size_t)(k - f0) * info.frame_bytes, info.w_size, info.h_size);
        }
        f0 = f1;
    }
    mjpeg423_b200_host_free(host);
}

// ---- accelerator seam ----------------------------------------------------------------------------------
int init_idct_ycbcr_to_rgb_accel(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c) return 0;                                       // reference convention: 1 = success
    if (!g_accel.ev_y && cudaEventCreate(&g_accel.ev_y) != cudaSuccess) return 0;
    g_accel.inited = true;
    return 1;
}
int mjpeg423_b200_accel_set_geometry(uint32_t w_size, uint32_t h_size) {
    if (!w_size || !h_size || (w_size & 7) || (h_size & 7)) return MJPEG423_E_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    g_accel.W = w_size; g_accel.H = h_size;
    return MJPEG423_OK;
}
void idct_accel_calculate_buffer_y(void* b, uint32_t n) { accel_submit(0, b, n); }
void idct_accel_calculate_buffer_cb(void* b, uint32_t n) { accel_submit(1, b, n); }
void idct_accel_calculate_buffer_cr(void* b, uint32_t n) { accel_submit(2, b, n); }

void ycbcr_to_rgb_accel_get_results(void* outputBuffer, uint32_t sizeOfOutputBuffer) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c || !g_accel.inited || !g_accel.coef.p) { set_error("accelerator not initialised / no planes submitted"); die("ycbcr_to_rgb_accel_get_results"); }
    const size_t frame = (size_t)g_accel.W * g_accel.H * 4;
    if (!outputBuffer || sizeOfOutputBuffer > frame) { set_error("output larger than the configured geometry"); die("ycbcr_to_rgb_accel_get_results"); }
    if (g_accel.out.reserve(frame)) die("ycbcr_to_rgb_accel_get_results");
    cudaError_t e = launch_idct_colour(g_accel.coef.as<int16_t>(), g_accel.out.p, 1, g_accel.W, g_accel.H, c->s_compute);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(outputBuffer, g_accel.out.p, sizeOfOutputBuffer, cudaMemcpyDeviceToHost, c->s_compute);
    if (e != cudaSuccess) { cuda_fail(e, "accelerator launch"); die("ycbcr_to_rgb_accel_get_results"); }
}
// C0/idct_ycbcr_to_rgb_accel.h:19-20 (declared there without a body): colour conversion of hCb_size x wCb_size sample
// blocks (block-major planes; note the reference's argument order Y, Cr, Cb) into the raster outputBuffer, whose rows
// are w_size pixels long.  Enqueued like the other accelerator calls; wait_for_ycbcr_to_rgb_finsh() completes it.
void ycbcr_to_rgb_accel_calculate_buffer(color_block_t* yBlock, color_block_t* crBlock, color_block_t* cbBlock,
                                         rgb_pixel_t* outputBuffer, int hCb_size, int wCb_size, int w_size) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c || !g_accel.inited) { set_error("accelerator not initialised"); die("ycbcr_to_rgb_accel_calculate_buffer"); }
    if (!yBlock || !crBlock || !cbBlock || !outputBuffer || hCb_size <= 0 || wCb_size <= 0 || w_size < wCb_size * 8) {
        set_error("bad argument"); die("ycbcr_to_rgb_accel_calculate_buffer");
    }
    const size_t nb = (size_t)hCb_size * wCb_size;
    const uint32_t W = (uint32_t)wCb_size * 8, H = (uint32_t)hCb_size * 8;
    if (s_in.reserve(3 * nb * 64) || s_out.reserve(nb * 256)) die("ycbcr_to_rgb_accel_calculate_buffer");
    cudaStream_t s = c->s_compute;
    cudaError_t e = cudaMemcpyAsync(s_in.as<uint8_t>(), yBlock, nb * 64, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s_in.as<uint8_t>() + nb * 64, cbBlock, nb * 64, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s_in.as<uint8_t>() + 2 * nb * 64, crBlock, nb * 64, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_colour(s_in.as<uint8_t>(), s_out.p, 1, W, H, s);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(outputBuffer, (size_t)w_size * 4, s_out.p, (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) { cuda_fail(e, "accelerator colour conversion"); die("ycbcr_to_rgb_accel_calculate_buffer"); }
}
void wait_for_ycbcr_to_rgb_finsh(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    mjpeg423_b200_ctx* c = shim_ctx();
    if (!c || cudaStreamSynchronize(c->s_compute) != cudaSuccess) { set_error("stream sync failed"); die("wait_for_ycbcr_to_rgb_finsh"); }
}
void wait_for_idct_y_finsh(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!shim_ctx() || !g_accel.ev_y || cudaEventSynchronize(g_accel.ev_y) != cudaSuccess) { set_error("event sync failed"); die("wait_for_idct_y_finsh"); }
}

}  // extern "C"
