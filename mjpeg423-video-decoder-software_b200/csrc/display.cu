// display.cu -- the consumers AFTER the hot path: the N-buffer frame ring of the
// reference's video controller and its paced playback loop, host side.
//
//   ring      C0/libs/ece423_vid_ctl/ece423_vid_ctl.{h,c} (C0 = /root/reference/core0/software): `num_buffers`
//             frame buffers, `buffer_being_written` (the producer's slot; free unless the scan-out is on it) and
//             `buffer_being_displayed` (the slot the mSGDMA scans out); register_written_buffer() advances the
//             first (:125-140), switch_frames() the second when a newer frame exists (:175-224),
//             buffer_is_available() tells the producer whether its slot is free (:157-173).  The mSGDMA descriptors
//             and the HDMI chip are HAL; here the buffers are pinned host memory (the D2H target of the decoder)
//             and "scan-out" is whatever the caller does with get_displayed_buffer().
//   player    C0/playback.c: process() decodes one frame into the written buffer and registers it (:80-134), a
//             timer at FRAME_RATE_US = 41666 (COMMON/config.h:29) flips the display (timerFunction, :36-46);
//             `noTimer` mode flips right after every frame (:130-133).  mjpeg423_b200_play() is that loop with
//             the GPU decoder as producer: frames are decoded in pipeline chunks and copied slot by slot into
//             the ring as slots free up, the consumer side runs from the same thread against a monotonic clock.
//   BMP dump  LIB/libbmp/encode_bmp.c:7-25: 32-bpp bottom-up BMP, byte-identical header (3780 px/m).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "runtime.h"

using namespace mj;

struct mjpeg423_b200_display {
    int width = 0, height = 0, bytes_per_pixel = 4, num_frame_buffers = 0, num_frame_buffers_mask = 0;
    size_t bytes_per_frame = 0;
    std::atomic<int> buffer_being_displayed{0}, buffer_being_written{0};
    void* buffers[MJPEG423_DISPLAY_MAX_BUFFERS] = {};
    bool pinned = false;
};

extern "C" mjpeg423_b200_display* mjpeg423_b200_display_init(int width, int height, int num_buffers) {
    if (width <= 0 || height <= 0) { set_error("display_init: bad geometry"); return nullptr; }
    // ece423_vid_ctl.c:55-61 clamps the count; the index arithmetic is a mask (:139,181), so it must be a power of two
    if (num_buffers > MJPEG423_DISPLAY_MAX_BUFFERS) num_buffers = MJPEG423_DISPLAY_MAX_BUFFERS;
    if (num_buffers < 2) num_buffers = 2;
    while (num_buffers & (num_buffers - 1)) num_buffers &= num_buffers - 1;      // round down to a power of two
    mjpeg423_b200_display* d = new mjpeg423_b200_display();
    d->width = width; d->height = height;
    d->bytes_per_frame = (size_t)width * height * 4;
    d->num_frame_buffers = num_buffers;
    d->num_frame_buffers_mask = num_buffers - 1;
    d->buffer_being_displayed = 0;
    d->buffer_being_written = 1;                                                   // :80
    int ndev = 0;
    d->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
    if (!d->pinned) cudaGetLastError();
    for (int i = 0; i < num_buffers; i++) {
        void* p = nullptr;
        if (d->pinned) { if (cudaMallocHost(&p, d->bytes_per_frame) != cudaSuccess) { cudaGetLastError(); p = nullptr; } }
        else p = std::malloc(d->bytes_per_frame);
        if (!p) { set_error("display_init: out of memory"); mjpeg423_b200_display_free(d); return nullptr; }
        std::memset(p, 0, d->bytes_per_frame);                                     // :104-107 clears every buffer to black
        d->buffers[i] = p;
    }
    return d;
}

extern "C" void mjpeg423_b200_display_free(mjpeg423_b200_display* d) {
    if (!d) return;
    for (void* p : d->buffers)
        if (p) { if (d->pinned) cudaFreeHost(p); else std::free(p); }
    delete d;
}

extern "C" void mjpeg423_b200_display_register_written_buffer(mjpeg423_b200_display* d) {       // :125-140
    d->buffer_being_written = (d->buffer_being_written + 1) & d->num_frame_buffers_mask;
}
extern "C" int mjpeg423_b200_display_buffer_is_available(mjpeg423_b200_display* d) {           // :157-173
    return d->buffer_being_displayed == d->buffer_being_written ? -1 : 0;
}
extern "C" int mjpeg423_b200_display_switch_frames(mjpeg423_b200_display* d) {                 // :175-224
    const int next = (d->buffer_being_displayed + 1) & d->num_frame_buffers_mask;
    if (next == d->buffer_being_written) return -1;                                             // no newer frame
    d->buffer_being_displayed = next;
    return 0;
}
extern "C" void* mjpeg423_b200_display_get_buffer(mjpeg423_b200_display* d) { return d->buffers[d->buffer_being_written]; }
extern "C" void* mjpeg423_b200_display_get_displayed_buffer(mjpeg423_b200_display* d) { return d->buffers[d->buffer_being_displayed]; }
extern "C" void mjpeg423_b200_display_clear_screen(mjpeg423_b200_display* d, char color) {     // :232-237
    std::memset(d->buffers[d->buffer_being_written], color, d->bytes_per_frame);
}
extern "C" int mjpeg423_b200_display_num_buffers(const mjpeg423_b200_display* d) { return d->num_frame_buffers; }

// encode_bmp(), LIB/libbmp/encode_bmp.c:7-25 (bmp_create_e(w, h, 32) + bmp_save): 14-byte file header, 40-byte
// BITMAPINFOHEADER with 3780 px/m, rows bottom-up, pixels as stored (B, G, R, A).
extern "C" int mjpeg423_b200_write_bmp(const char* path, const rgb_pixel_t* rgb, uint32_t W, uint32_t H) {
    if (!path || !rgb || !W || !H) return MJPEG423_E_ARG;
    FILE* f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return MJPEG423_E_ARG; }
    const uint32_t img = W * H * 4, off = 54, size = off + img;
    uint8_t h[54] = {'B', 'M'};
    auto p32 = [&](int at, uint32_t v) { h[at] = (uint8_t)v; h[at + 1] = (uint8_t)(v >> 8); h[at + 2] = (uint8_t)(v >> 16); h[at + 3] = (uint8_t)(v >> 24); };
    p32(2, size); p32(10, off); p32(14, 40); p32(18, W); p32(22, H);
    h[26] = 1; h[28] = 32; p32(34, img); p32(38, 3780); p32(42, 3780);
    bool ok = std::fwrite(h, 1, 54, f) == 54;
    const uint8_t* px = reinterpret_cast<const uint8_t*>(rgb);
    for (uint32_t y = 0; ok && y < H; y++) ok = std::fwrite(px + (size_t)(H - 1 - y) * W * 4, 1, (size_t)W * 4, f) == (size_t)W * 4;
    std::fclose(f);
    if (!ok) { set_error(std::string("short write to ") + path); return MJPEG423_E_ARG; }
    return MJPEG423_OK;
}

// The playback loop (C0/playback.c: playVideo -> process + timerFunction).  frame_period_us == 0 is the reference's
// noTimer mode: the display flips right after every frame.  on_display (may be NULL) is called once per displayed
// frame, in display order, with the scanned-out buffer.  Returns the number of frames displayed or a negative
// MJPEG423_E_* code; *dropped (optional) counts timer ticks that found no new frame (the reference's switch_frames
// returning -1: the previous frame stays on screen).
static long play_impl(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n,
                      mjpeg423_b200_display* d, uint32_t frame_period_us,
                      void (*on_display)(void* user, uint32_t frame_index, const rgb_pixel_t* frame), void* user,
                      uint32_t* dropped) {
    if (!c || !d || !mpg) return MJPEG423_E_ARG;
    mjpeg423_b200_info info;
    int rc = mjpeg423_b200_probe(mpg, len, &info);
    if (rc) return rc;
    if ((int)info.w_size != d->width || (int)info.h_size != d->height) { set_error("play: display geometry differs from the stream"); return MJPEG423_E_ARG; }
    if ((uint64_t)first + n > info.num_frames) { set_error("play: frame range exceeds num_frames"); return MJPEG423_E_ARG; }
    if (cudaSetDevice(c->device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    // Producer side: frames are decoded on the device a batch at a time (batches start on I frames, so a batch may
    // be longer than `batch` when the stream holds P frames) and leave it slot by slot.
    const uint32_t batch = 64;
    DevBuf d_frames;
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    uint64_t ticks = 0;
    uint32_t produced = 0, displayed = 0, misses = 0;
    uint32_t have_lo = 0, have_hi = 0;                  // frames [have_lo, have_hi) of the range are resident in d_frames
    auto tick = [&]() {                                 // timerFunction, playback.c:36-46
        if (mjpeg423_b200_display_switch_frames(d) == 0) {
            if (on_display) on_display(user, first + displayed, (const rgb_pixel_t*)mjpeg423_b200_display_get_displayed_buffer(d));
            displayed++;
        } else if (displayed < n) misses++;
    };
    while (displayed < n) {
        if (produced < n && mjpeg423_b200_display_buffer_is_available(d) == 0) {
            if (produced >= have_hi) {                  // decode the next batch into device memory
                uint32_t hi = std::min(n, produced + batch);
                // extend to the next I frame so that the following batch can start there
                // (probe() only counts; walk the frame headers)
                {
                    size_t off = 20; uint32_t f = 0;
                    for (; f < first + hi && off + 16 <= len; f++) { uint32_t sz; std::memcpy(&sz, mpg + off, 4); off += sz; }
                    while (first + hi < info.num_frames && hi < n && off + 16 <= len) {
                        uint32_t sz, type; std::memcpy(&sz, mpg + off, 4); std::memcpy(&type, mpg + off + 4, 4);
                        if (type == 0) break;
                        hi++; off += sz;
                    }
                }
                if ((rc = d_frames.reserve((size_t)(hi - produced) * info.frame_bytes))) return rc;
                if ((rc = mjpeg423_b200_decode_frames(c, mpg, len, first + produced, hi - produced, d_frames.p, 1))) { d_frames.release(); return rc; }
                have_lo = produced; have_hi = hi;
            }
            cudaError_t e = cudaMemcpy(mjpeg423_b200_display_get_buffer(d), d_frames.as<uint8_t>() + (size_t)(produced - have_lo) * info.frame_bytes,
                                       info.frame_bytes, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { d_frames.release(); return cuda_fail(e, "play: frame read-back"); }
            mjpeg423_b200_display_register_written_buffer(d);      // playback.c:125
            produced++;
            if (frame_period_us == 0) { tick(); continue; }        // noTimer, playback.c:130-133
        }
        if (frame_period_us) {
            const auto due = t0 + std::chrono::microseconds((uint64_t)frame_period_us * (ticks + 1));
            if (clk::now() >= due) { ticks++; tick(); }
            else if (produced >= n || mjpeg423_b200_display_buffer_is_available(d) != 0) std::this_thread::sleep_until(due);
        }
    }
    d_frames.release();
    if (dropped) *dropped = misses;
    return (long)displayed;
}
extern "C" long mjpeg423_b200_play(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n,
                                   mjpeg423_b200_display* d, uint32_t frame_period_us,
                                   void (*on_display)(void* user, uint32_t frame_index, const rgb_pixel_t* frame), void* user,
                                   uint32_t* dropped) {
    return mj::guard<long>([&]() -> long { return play_impl(c, mpg, len, first, n, d, frame_period_us, on_display, user, dropped); });
}
