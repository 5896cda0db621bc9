/*
 * mjpeg423_b200.h -- C-ABI of the B200-native MJPEG423 decode hot path (libmjpeg423_b200.so).
 *
 * Reference citations: LIB = /root/reference/core0/software/common/libs/mjpeg423,
 *                      C0  = /root/reference/core0/software.
 *
 * Three groups of entry points:
 *   1. the reference LIBRARY seam   -- the exact symbols of LIB/decoder/mjpeg423_decoder.h:14-17 and the
 *      tables of LIB/common/mjpeg423_types.h:64-66, so a caller of the reference links against this
 *      library unchanged (drop-in shims: every call runs the CUDA path on a batch of one);
 *   2. the reference ACCELERATOR seam -- the submit/collect/poll calls of
 *      C0/idct_ycbcr_to_rgb_accel.h:13-22 (the FPGA IDCT + colour block this library replaces);
 *   3. the THROUGHPUT API (additive) -- batched frame-range decode of an in-memory .mpg, mirroring the
 *      loop body LIB/decoder/mjpeg423_decoder.c:90-124, with device-resident and host-buffer variants.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Group 1/2 functions return void/int exactly
 * as the reference does (1 = success for init, C0/idct_ycbcr_to_rgb_accel.c:39-59).  Group 3 functions
 * return 0 on success or a negative MJPEG423_E_* code and never call exit().  The caller owns every
 * buffer passed in.  There is NO CPU fallback: without a CUDA device every entry point fails (group 3:
 * MJPEG423_E_CUDA; group 1/2: message on stderr + abort(), the library analogue of the reference's
 * error_and_exit, LIB/common/util.c:13-16).
 */
#ifndef MJPEG423_B200_H
#define MJPEG423_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types: LIB/common/mjpeg423_types.h:22-61 ---------------------------------------------------- */
#ifndef MJPEG423_B200_NO_REFERENCE_TYPES
/* mjpeg423_types.h:15-19: `bool` is `int` in the reference ABI; spelled int below. */
typedef struct { uint32_t frame_index; uint32_t frame_position; } iframe_trailer_t; /* :22-25 */
typedef uint8_t color_block_t[8][8];    /* :33 */
typedef uint8_t (*pcolor_block_t)[8];   /* :34 */
#ifndef DCTELEM
#define DCTELEM int16_t                 /* :39 */
#endif
typedef DCTELEM dct_block_t[8][8];      /* :42 */
typedef DCTELEM (*pdct_block_t)[8];     /* :43 */
typedef struct { uint8_t blue, green, red, alpha; } rgb_pixel_t; /* :56-61, BMP byte order */
#endif

/* ---- 1. reference library seam: LIB/decoder/mjpeg423_decoder.h:14-17 ------------------------------ */
/* P is the reference's `bool` (int).  DCACq is in/out: for P != 0 decoded deltas are ADDED to it
 * (LIB/decoder/lossless_decode.c:90-92,121-123).  The bitstream must be readable 4 bytes past its last
 * symbol.  The reference signature carries no length: the shim first walks the symbol headers of
 * num_blocks blocks on the host to find where the stream ends (it reads no byte the reference decoder would not read)
 * and uploads exactly that; mjpeg423_b200_set_read_limit() bounds the walk for non-conforming input (default:
 * num_blocks*184 + 8 bytes, the longest stream the reference decoder accepts). */
void lossless_decode(int num_blocks, void* bitstream, dct_block_t* DCACq, dct_block_t quant, int P);
void idct(dct_block_t DCAC, color_block_t block);
void ycbcr_to_rgb(int h, int w, uint32_t w_size, pcolor_block_t Y, pcolor_block_t Cb, pcolor_block_t Cr,
                  rgb_pixel_t* rgbblock);
/* Decodes file `filename_in` and writes one 32-bpp BMP per frame, names derived from
 * `filenamebase_out` ("name0000.bmp" pattern, LIB/decoder/mjpeg423_decoder.c:127-132). */
void mjpeg423_decode(const char* filename_in, const char* filenamebase_out);
extern dct_block_t Yquant;              /* LIB/common/tables.c:13-21 */
extern dct_block_t Cquant;              /* LIB/common/tables.c:24-32 */
extern int zigzag_table[64];            /* LIB/common/tables.c:35-42 */

void   mjpeg423_b200_set_read_limit(size_t bytes);   /* 0 restores the default */
size_t mjpeg423_b200_get_read_limit(void);

/* Plane-at-a-time variants of the per-block shims (same arithmetic, one launch per call):
 * n blocks of coefficients -> n blocks of samples; block-major planes -> W x H BGRA raster. */
int mjpeg423_b200_idct_blocks(const int16_t* coef, uint8_t* samples, size_t n_blocks);
int mjpeg423_b200_ycbcr_to_rgb_frame(const uint8_t* Y, const uint8_t* Cb, const uint8_t* Cr, uint32_t w_size,
                                     uint32_t h_size, rgb_pixel_t* rgb);
/* lossless_decode with an explicit stream length. */
int mjpeg423_b200_lossless_decode(int num_blocks, const void* bitstream, size_t bitstream_len,
                                  int16_t* DCACq, const int16_t* quant, int P);

/* ---- 2. reference accelerator seam: C0/idct_ycbcr_to_rgb_accel.h:13-22 ----------------------------- */
/* Asynchronous: the calculate_buffer_* calls enqueue the upload of one dequantised coefficient plane
 * (sizeOfInputBuffer bytes = blocks*128, C0/playback.c:71-75,102-103); get_results enqueues IDCT +
 * colour conversion of the three planes and the read-back of the BGRA frame into outputBuffer
 * (C0/playback.c:108-109); the wait_* calls block like the reference's CSR busy-polls
 * (C0/idct_ycbcr_to_rgb_accel.c:84-98).  Frame geometry defaults to the reference's 640x480
 * (COMMON/config.h:23-24) and is changed with mjpeg423_b200_accel_set_geometry. */
int  init_idct_ycbcr_to_rgb_accel(void);
void idct_accel_calculate_buffer_y(void* inputBuffer, uint32_t sizeOfInputBuffer);
void idct_accel_calculate_buffer_cb(void* inputBuffer, uint32_t sizeOfInputBuffer);
void idct_accel_calculate_buffer_cr(void* inputBuffer, uint32_t sizeOfInputBuffer);
void ycbcr_to_rgb_accel_get_results(void* outputBuffer, uint32_t sizeOfOutputBuffer);
/* C0/idct_ycbcr_to_rgb_accel.h:19-20 (declared in the reference, no body there): colour conversion of
 * hCb_size x wCb_size sample blocks (block-major planes; argument order Y, Cr, Cb as in the reference) into the
 * raster outputBuffer with rows of w_size pixels; asynchronous like the calls above. */
void ycbcr_to_rgb_accel_calculate_buffer(color_block_t* yBlock, color_block_t* crBlock, color_block_t* cbBlock,
                                         rgb_pixel_t* outputBuffer, int hCb_size, int wCb_size, int w_size);
void wait_for_ycbcr_to_rgb_finsh(void);   /* sic: reference spelling */
void wait_for_idct_y_finsh(void);
int  mjpeg423_b200_accel_set_geometry(uint32_t w_size, uint32_t h_size);

/* ---- 3. throughput API ------------------------------------------------------------------------------ */
#define MJPEG423_OK            0
#define MJPEG423_E_ARG        -1   /* bad argument / geometry (W, H must be non-zero multiples of 8) */
#define MJPEG423_E_FORMAT     -2   /* container does not parse (truncated, sizes inconsistent) */
#define MJPEG423_E_CUDA       -3   /* CUDA runtime error or no device (see mjpeg423_b200_last_error) */
#define MJPEG423_E_NOMEM      -4
#define MJPEG423_E_STREAM     -5   /* a plane stream did not contain num_blocks blocks */
#define MJPEG423_E_PFRAME     -6   /* frame range starts on a P frame (needs the preceding I frame) */

typedef struct mjpeg423_b200_ctx mjpeg423_b200_ctx;

typedef struct {
    uint32_t num_frames, w_size, h_size, num_iframes, payload_size; /* file header, mjpeg423_decoder.c:33-38 */
    uint32_t num_pframes;
    uint64_t frame_bytes;        /* w_size*h_size*4: one decoded BGRA frame */
    uint64_t max_frame_payload;  /* largest frame record */
} mjpeg423_b200_info;

typedef struct {
    /* CUDA-event times (ms) of the most recent decode on this context, measured on the library's
     * streams.  total_ms is always collected; the per-kernel times only when profiling is on
     * (set_option PROFILE 1), because the extra events serialise the kernels. */
    float total_ms;              /* whole device-side decode (all kernels of the call) */
    float sync_ms;               /* k_entropy_sync: speculative parse + merge */
    float chain_ms;              /* k_entropy_chain: chain fix-up + scans */
    float index_ms;              /* k_entropy_index: per-block bit positions and DC levels */
    float decode_ms;             /* k_decode_fused (default) or k_decode_coef (staged modes) */
    float idct_colour_ms;        /* k_idct_colour (staged mode 1) */
    float idct_ms, colour_ms;    /* k_idct, k_colour (staged mode 2) */
    uint64_t kernel_launches;    /* kernels launched by the call */
    uint64_t payload_bytes;      /* compressed bytes consumed (sum of plane streams) */
    uint64_t segments, fixups;   /* bitstream segments / segments re-parsed by the chain kernel */
    uint64_t frames;
    uint64_t list_entries;       /* coded AC coefficients written to the symbol lists (4 bytes each) */
} mjpeg423_b200_stats;

enum {
    MJPEG423_OPT_PROFILE      = 1,  /* 0/1: collect per-stage event times */
    MJPEG423_OPT_STAGED       = 2,  /* 0: one fused decode kernel, bitstream -> BGRA (default; ranges with P frames
                                          use its GOP-walking variant); 1: coefficient planes in HBM + fused
                                          IDCT/colour kernel; 2: entropy, IDCT and colour kernels separate */
    MJPEG423_OPT_CHUNK_FRAMES = 3,  /* frames per pipeline chunk in the host-buffer path (0 = auto) */
    MJPEG423_OPT_VALIDATE     = 4   /* 0/1: check every stream decoded exactly num_blocks blocks (default 1) */
};

int  mjpeg423_b200_create(mjpeg423_b200_ctx** ctx, int device);
void mjpeg423_b200_destroy(mjpeg423_b200_ctx* ctx);
int  mjpeg423_b200_set_option(mjpeg423_b200_ctx* ctx, int option, int64_t value);
const char* mjpeg423_b200_last_error(void);      /* thread-local, never NULL */

/* Parse the header of an in-memory .mpg (layout: SURVEY.md A.1). */
int mjpeg423_b200_probe(const uint8_t* mpg, size_t len, mjpeg423_b200_info* info);

/* Custom quantisation tables (natural order, 64 x int16 each); NULL restores Yquant / Cquant.  The file
 * format carries no tables: like the reference, they are a parameter of the decode
 * (LIB/decoder/mjpeg423_decoder.c:110-112). */
int mjpeg423_b200_set_quant(mjpeg423_b200_ctx* ctx, const int16_t* yquant, const int16_t* cquant);

/* End-to-end: decode frames [first, first+n) of the .mpg held in HOST memory into `out`.
 * out_on_device == 0: out is host memory (n * frame_bytes); pinned memory (mjpeg423_b200_host_alloc)
 * gives full PCIe speed.  The call uploads the bitstream, decodes and reads back in overlapping
 * chunks on several CUDA streams and returns when `out` is complete.
 * out_on_device != 0: out is device memory on the context's device; no read-back. */
int mjpeg423_b200_decode_frames(mjpeg423_b200_ctx* ctx, const uint8_t* mpg, size_t len, uint32_t first,
                                uint32_t n, void* out, int out_on_device);

/* Device-resident variant: upload once, decode many times (bench `value`, long-lived servers).
 * upload() parses frames [first, first+n), copies their payload to HBM and builds the stream tables;
 * decode_resident() runs only kernels and writes n * frame_bytes to device memory d_out. */
int mjpeg423_b200_upload(mjpeg423_b200_ctx* ctx, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n);
int mjpeg423_b200_decode_resident(mjpeg423_b200_ctx* ctx, void* d_out);
int mjpeg423_b200_get_stats(mjpeg423_b200_ctx* ctx, mjpeg423_b200_stats* stats);

/* Stage-level device entry points on the resident job (per-stage profiling and parity tests):
 * coefficient planes are n frames x 3 planes x nb blocks x 64 int16 (frame-major, Y|Cb|Cr);
 * sample planes the same shape in uint8. */
int mjpeg423_b200_resident_entropy(mjpeg423_b200_ctx* ctx, int16_t* d_coef);
int mjpeg423_b200_resident_idct(mjpeg423_b200_ctx* ctx, const int16_t* d_coef, uint8_t* d_samples);
int mjpeg423_b200_resident_colour(mjpeg423_b200_ctx* ctx, const uint8_t* d_samples, void* d_out);
int mjpeg423_b200_resident_idct_colour(mjpeg423_b200_ctx* ctx, const int16_t* d_coef, void* d_out);

/* Memory helpers (so a C caller needs no CUDA headers). */
void* mjpeg423_b200_host_alloc(size_t bytes);     /* pinned host memory, NULL on failure */
void  mjpeg423_b200_host_free(void* p);
void* mjpeg423_b200_device_alloc(mjpeg423_b200_ctx* ctx, size_t bytes);
void  mjpeg423_b200_device_free(mjpeg423_b200_ctx* ctx, void* p);
int   mjpeg423_b200_memcpy_d2h(mjpeg423_b200_ctx* ctx, void* dst, const void* d_src, size_t bytes);
int   mjpeg423_b200_memcpy_h2d(mjpeg423_b200_ctx* ctx, void* d_dst, const void* src, size_t bytes);
int   mjpeg423_b200_sync(mjpeg423_b200_ctx* ctx);
int   mjpeg423_b200_device_count(void);
/* Position-mixed 64-bit checksum of every frame of a device-resident output (n x frame_bytes), computed
 * on the GPU: sum over 8-byte words w_i of splitmix64(w_i ^ (i+1)*0x9E3779B97F4A7C15).  hashes (n x uint64)
 * is host memory.  Used by the bench for whole-batch bit-exactness checks. */
int   mjpeg423_b200_hash_frames(mjpeg423_b200_ctx* ctx, const void* d_frames, uint64_t frame_bytes, uint32_t n,
                                uint64_t* hashes);

/* ---- 3b. container index and seek ---------------------------------------------------- */
/* The I-frame index of a .mpg: one iframe_trailer_t {frame_index, frame_position = file offset of the frame's
 * frame_size field} per I frame, in file order -- what the reference reads from the trailer at 20 + payload_size
 * (LIB/decoder/mjpeg423_decoder.c:78-86, C1/main.c:77-113).  The entries are rebuilt from a bounds-checked walk over
 * the frame headers and COMPARED with the file's trailer: *trailer_ok (may be NULL) is 1 when the trailer is present and
 * says the same, 0 when it is missing, truncated or wrong (the index returned is valid either way).  out may be NULL to
 * query *n_iframes; cap = entries out can hold. */
int mjpeg423_b200_index(const uint8_t* mpg, size_t len, iframe_trailer_t* out, uint32_t cap, uint32_t* n_iframes,
                        int* trailer_ok);
/* Position in idx[] of the I frame a decode of `frame` must start from (direction <= 0: the last I frame at or before
 * it) or of the first I frame at or after it (direction > 0); -1 if there is none. */
int mjpeg423_b200_seek_iframe(const iframe_trailer_t* idx, uint32_t n, uint32_t frame, int direction);
/* The player's jumps, C0/playback.c:157-227: fast-forward = the first I frame at least 108 frames ahead of `current`
 * (-1 = nothing happens: fewer than 120 frames left); rewind = the last I frame at least 108 frames back (entry 0 when
 * `current` is less than 120 frames from the start).  Both return a position in idx[]. */
int mjpeg423_b200_fast_forward(const iframe_trailer_t* idx, uint32_t n, uint32_t num_frames, uint32_t current);
int mjpeg423_b200_rewind(const iframe_trailer_t* idx, uint32_t n, uint32_t current);

/* ---- 3c. frame-range sharding over the GPUs of one box (SURVEY.md 8 row e) ------------------------------------ */
/* Frames [first, first + n) of the logical stream formed by n_shards .mpg files in order (one file is the common case;
 * long streams come as several files because the container's offsets are 32-bit, LIB/encoder/mjpeg423_encoder.c:67,
 * 209,222-225; all files must share one geometry and start on an I frame) are decoded on n_dev devices: the range is cut
 * into n_dev contiguous pieces on I frames, one host thread per device runs mjpeg423_b200_decode_frames on its piece
 * and writes its slice of `out` (host memory, n x frame_bytes; pinned memory gives full PCIe speed).  No data moves
 * between GPUs.  devices = NULL means 0 .. n_dev-1; a device may be listed more than once (one context per entry).
 * cuts (may be NULL) receives the n_dev + 1 piece boundaries.  The tables set with mjpeg423_b200_set_quant do not
 * apply: the per-device contexts are private to this call and use Yquant / Cquant. */
typedef struct { const uint8_t* mpg; size_t len; } mjpeg423_b200_shard;
int mjpeg423_b200_decode_frames_multi(const int* devices, int n_dev, const mjpeg423_b200_shard* shards, uint32_t n_shards,
                                      uint64_t first, uint64_t n, void* out, uint64_t* cuts);

/* ---- 4. encoder: LIB/encoder/mjpeg423_encoder.h ---------------------------------- */
/* The frame loop of mjpeg423_encode(), LIB/encoder/mjpeg423_encoder.c:97-225, on in-memory frames: n frames of
 * w_size x h_size rgb_pixel_t (BGRA, top-down raster -- what decode_bmp() hands the reference, alpha ignored),
 * in host memory or (frames_on_device != 0) on the context's device, are colour-converted, transformed,
 * quantised with the context's tables (mjpeg423_b200_set_quant), coded as I or P frames by the reference's rule
 * (I when first frame, I not larger than P, or max_I_interval frames after the last I frame) and written as a
 * complete .mpg into the HOST buffer `mpg` of `cap` bytes; *mpg_len receives its length.  The
 * file equals the reference encoder's byte for byte, except its last 512 bytes (uninitialised stack in the
 * reference, zeros here).  The BMP reader of the reference (libnsbmp) is file I/O and stays with the caller.
 * flags: MJPEG423_ENC_FIX_TAIL writes the last partial byte of every plane stream; by default it is 0 like the
 * reference's (output_rest, LIB/encoder/lossless_encode.c:80-83). */
#define MJPEG423_ENC_FIX_TAIL 1u
size_t mjpeg423_b200_encode_bound(uint32_t n, uint32_t w_size, uint32_t h_size);   /* a sufficient `cap` */
int mjpeg423_b200_encode_frames(mjpeg423_b200_ctx* ctx, const void* frames, int frames_on_device, uint32_t n,
                                uint32_t w_size, uint32_t h_size, uint32_t max_I_interval, uint32_t flags,
                                uint8_t* mpg, size_t cap, size_t* mpg_len);

/* ---- 5. display side: C0/libs/ece423_vid_ctl/ece423_vid_ctl.h:67-77, C0/playback.c ---- */
/* The reference's N-buffer frame ring with the mSGDMA/HDMI parts removed: buffers are pinned host memory (plain
 * memory without a CUDA device), "scan-out" is get_displayed_buffer().  Same state machine and return values as
 * ece423_video_display_{init, register_written_buffer, buffer_is_available, switch_frames, get_buffer,
 * clear_screen}: buffer_is_available() returns 0 when the producer's slot is free and -1 while it is on screen;
 * switch_frames() returns 0 after flipping to a newer frame and -1 when there is none.  num_buffers is clamped to
 * [2, 25] like the reference and rounded down to a power of two (the indices are masked, config.h:27). */
#define MJPEG423_DISPLAY_MAX_BUFFERS 25
typedef struct mjpeg423_b200_display mjpeg423_b200_display;
mjpeg423_b200_display* mjpeg423_b200_display_init(int width, int height, int num_buffers);
void  mjpeg423_b200_display_free(mjpeg423_b200_display* display);
void  mjpeg423_b200_display_register_written_buffer(mjpeg423_b200_display* display);
int   mjpeg423_b200_display_buffer_is_available(mjpeg423_b200_display* display);
int   mjpeg423_b200_display_switch_frames(mjpeg423_b200_display* display);
void* mjpeg423_b200_display_get_buffer(mjpeg423_b200_display* display);            /* the producer's slot */
void* mjpeg423_b200_display_get_displayed_buffer(mjpeg423_b200_display* display);  /* the slot on screen */
void  mjpeg423_b200_display_clear_screen(mjpeg423_b200_display* display, char color);
int   mjpeg423_b200_display_num_buffers(const mjpeg423_b200_display* display);
/* The playback loop of C0/playback.c with the GPU decoder as producer: frames [first, first+n) go through the ring
 * and are displayed every frame_period_us (41666 = the reference's 24 fps, COMMON/config.h:29; 0 = the reference's
 * noTimer mode: flip after every frame).  on_display (may be NULL) sees every displayed frame in order.  Returns
 * the number of frames displayed (n) or a negative MJPEG423_E_*; *dropped (may be NULL) = timer ticks without a new
 * frame. */
long mjpeg423_b200_play(mjpeg423_b200_ctx* ctx, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n,
                        mjpeg423_b200_display* display, uint32_t frame_period_us,
                        void (*on_display)(void* user, uint32_t frame_index, const rgb_pixel_t* frame), void* user,
                        uint32_t* dropped);
/* encode_bmp(), LIB/libbmp/encode_bmp.c:7-25: one 32-bpp bottom-up BMP, byte-identical to the reference's file. */
int mjpeg423_b200_write_bmp(const char* path, const rgb_pixel_t* rgb, uint32_t w_size, uint32_t h_size);

#ifdef __cplusplus
}
#endif
#endif /* MJPEG423_B200_H */
