// runtime.h -- host runtime of libmjpeg423_b200: container walk, stream/tile tables, device buffers,
// chunked multi-stream pipeline.  Internal header (the public surface is include/mjpeg423_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "mjpeg423_b200.h"

namespace mj {

// Device view of the entropy stage for one launch (one chunk of frames).  All table pointers are
// already offset so that the indices stored in StreamDesc (which are relative to the whole
// plan) address the right element.
struct EntropyJob {
    const uint8_t* d_payload = nullptr;          // address of plan.payload_off (may lie before the chunk's buffer)
    const StreamDesc* d_streams = nullptr;       // plan-wide table
    const uint32_t* d_seg_stream = nullptr;      // plan-wide: global segment -> stream
    uint32_t seg_lo = 0, seg_hi = 0;             // global segments of the chunk
    bool fold_end = true;                        // Parser::step look-ahead for END symbols: pays for sparse streams (few
                                                 // symbols per block), costs for dense ones; same trajectory either way
    uint32_t stream_lo = 0, n_streams = 0;       // streams of the chunk (chain kernel: one CTA each)
    uint32_t *d_seg_entry = nullptr, *d_seg_exit = nullptr, *d_seg_cnt = nullptr, *d_seg_first = nullptr; // plan-wide
    uint32_t* d_seg_dc = nullptr;                // plan-wide: DC total, then DC predictor, of every segment
    uint32_t* d_stream_blocks = nullptr;         // plan-wide, per stream
    unsigned long long* d_fixups = nullptr;
    uint2* d_blk_info = nullptr;                 // block index, addressed by StreamDesc.block_base + b
    uint32_t* d_sym = nullptr;                   // symbol lists: segment g's region starts at (g - sym_seg0) * SYM_STRIDE
    uint32_t sym_seg0 = 0;
};

cudaError_t launch_seg_stream(const StreamDesc* d_streams, uint32_t n_streams, uint32_t* d_seg_stream, cudaStream_t s);
cudaError_t launch_entropy_sync(const EntropyJob& j, cudaStream_t s);
cudaError_t launch_entropy_chain(const EntropyJob& j, cudaStream_t s);
cudaError_t launch_entropy_index(const EntropyJob& j, cudaStream_t s);
cudaError_t launch_decode_coef(const EntropyJob& j, const uint32_t* d_stream_ids, uint32_t n_ids, uint32_t nb,
                               const int16_t* d_quant, int16_t* d_coef, cudaStream_t s);
// d_gop_first == nullptr: intra-only range.  Otherwise d_gop_first[0 .. n_gops] = chunk-relative first frames of the
// range's GOPs (+ its end) and d_state = FUSED_STATE_BYTES of scratch that no other launch in flight uses.
constexpr int FUSED_MAX_CTAS = 192;                                     // persistent grid: one CTA per SM, at most this many
constexpr size_t FUSED_STATE_BYTES = (size_t)FUSED_MAX_CTAS * 18 * 3 * 8 * 32 * 16;   // 12 KB per warp
cudaError_t launch_decode_fused(const EntropyJob& j, const int16_t* d_quant, void* d_out, uint32_t n_frames,
                                uint32_t W, uint32_t H, const uint32_t* d_gop_first, uint32_t n_gops, void* d_state,
                                cudaStream_t s);
cudaError_t launch_idct(const int16_t* d_coef, uint8_t* d_samples, size_t n_blocks, cudaStream_t s);
cudaError_t launch_colour(const uint8_t* d_samples, void* d_out, uint32_t n_frames, uint32_t W, uint32_t H,
                          cudaStream_t s);
cudaError_t launch_idct_colour(const int16_t* d_coef, void* d_out, uint32_t n_frames, uint32_t W, uint32_t H,
                               cudaStream_t s);
cudaError_t launch_hash_frames(const void* d_frames, uint64_t frame_bytes, uint32_t n, unsigned long long* d_hashes,
                               cudaStream_t s);

// ---- container (reader LIB/decoder/mjpeg423_decoder.c:33-38,94-107) -----------------
struct FrameRec {
    uint64_t off;          // file offset of the 16-byte frame header
    uint32_t size, type, ysize, cbsize, crsize;
};
struct MpgIndex {
    mjpeg423_b200_info info{};
    std::vector<FrameRec> frames;
};
int parse_mpg(const uint8_t* mpg, size_t len, MpgIndex& idx, bool headers_only);

// Host-side tables for frames [first, first+n).
struct Plan {
    uint32_t W = 0, H = 0, nb = 0, first = 0, n = 0;
    uint64_t payload_off = 0, payload_len = 0;   // byte range of the file covered by the frames
    std::vector<FrameRec> frames;
    std::vector<StreamDesc> streams;             // 3 per frame: Y, Cb, Cr
    std::vector<uint32_t> f_seg0;                // n+1: first global segment of every frame
    uint64_t stream_bytes = 0;                   // sum of plane stream lengths
    uint32_t n_pframes = 0;
};
int build_plan(const MpgIndex& idx, uint32_t first, uint32_t n, Plan& plan);

// A pipeline chunk: frames [f0, f1), always starting on an I frame.  P frames accumulate on the
// previous frame's coefficients, so k_decode_coef runs once per GOP depth ("level"): level l holds the
// streams of the frames that are l frames after their I frame.  ids_off indexes the uploaded id list.
struct Chunk {
    uint32_t f0 = 0, f1 = 0;
    std::vector<uint32_t> level_off;             // size levels+1, offsets into the chunk's ids
    uint32_t ids_off = 0;
    uint32_t gop_off = 0, n_gops = 0;            // gops[gop_off .. gop_off + n_gops]: chunk-relative first frame of every
                                                 // GOP of the chunk, then f1 - f0 (what k_decode_fused<true> walks)
};
void make_chunks(const Plan& plan, uint32_t K, std::vector<Chunk>& chunks, std::vector<uint32_t>& ids,
                 std::vector<uint32_t>& gops);

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// No C++ exception may cross the C boundary (a crafted header must not end the host process in std::terminate):
// every extern "C" entry that allocates host memory runs its body through guard().
template <class R = int, class F>
R guard(F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return (R)MJPEG423_E_NOMEM;
    } catch (const std::exception& e) {
        set_error(std::string("internal error: ") + e.what());
        return (R)MJPEG423_E_ARG;
    }
}

}  // namespace mj

// Device buffer with lazy growth.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct mjpeg423_b200_ctx {
    int device = 0;
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr, s_aux = nullptr;
    // options
    bool profile = false, validate = true;
    int staged = 0;             // 0: fused decode, 1: coefficient planes + fused IDCT/colour, 2: all stages separate
    uint32_t chunk_frames = 0;
    // quant tables (2 x 64 int16, natural order) on device
    int16_t h_quant[128];
    int16_t* d_quant = nullptr;
    // resident job
    mj::Plan plan;
    bool have_plan = false;
    DevBuf payload, tables, segs, coef[2], blkidx[2], samples[2], stream_blocks, misc, ids;   // [2]: one per chunk buffer in flight
    DevBuf fstate[2];           // k_decode_fused<true>: parked coefficient slots, one area per chunk buffer in flight
    std::vector<mj::Chunk> chunks;
    uint32_t chunk_K = 0;
    size_t gops_base = 0;       // the GOP tables follow the stream ids in `ids` (in uint32 units)
    // staging for the host-buffer path
    DevBuf in_ring[2], out_ring[2];
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_cap[2] = {0, 0};
    cudaEvent_t ev[16] = {};
    mjpeg423_b200_stats stats{};
};
