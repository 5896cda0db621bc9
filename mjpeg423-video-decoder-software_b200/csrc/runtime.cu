// runtime.cu -- host runtime: container walk, stream/tile tables, device buffers and the chunked
// multi-stream decode pipeline behind the throughput API of include/mjpeg423_b200.h.
//
// It batches the per-frame loop body of the reference decoder,
// LIB/decoder/mjpeg423_decoder.c:90-124 (LIB = /root/reference/core0/software/common/libs/mjpeg423):
// read frame header -> 3 x lossless_decode -> 3*nb x idct -> nb x ycbcr_to_rgb, over a frame range.
#include "runtime.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace mj {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return MJPEG423_E_CUDA;
}
#define CU(call)                                                \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) return mj::cuda_fail(e_, #call); \
    } while (0)

static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// File header: 5 x u32 {num_frames, w_size, h_size, num_iframes, payload_size}; then per frame a 16-byte
// header {frame_size (incl. header, padded to x4), frame_type, Ysize, Cbsize} + Y|Cb|Cr streams.
int parse_mpg(const uint8_t* mpg, size_t len, MpgIndex& idx, bool headers_only) {
    if (!mpg || len < 20) { set_error("mpg: shorter than the 20-byte file header"); return MJPEG423_E_FORMAT; }
    mjpeg423_b200_info& in = idx.info;
    in.num_frames = rd32(mpg); in.w_size = rd32(mpg + 4); in.h_size = rd32(mpg + 8);
    in.num_iframes = rd32(mpg + 12); in.payload_size = rd32(mpg + 16);
    if (!in.w_size || !in.h_size || (in.w_size & 7) || (in.h_size & 7)) {
        set_error("mpg: width and height must be non-zero multiples of 8");   // mjpeg423_decoder.c:45-48: no edge handling
        return MJPEG423_E_ARG;
    }
    // (W/8)*(H/8) must stay well inside 32 bits: block indices and bit positions are 32-bit throughout
    if ((uint64_t)(in.w_size / 8) * (in.h_size / 8) > MAX_PLANE_BLOCKS) {
        set_error("mpg: picture too large (more than 2^24 blocks per plane)");
        return MJPEG423_E_ARG;
    }
    in.frame_bytes = (uint64_t)in.w_size * in.h_size * 4;
    in.num_pframes = 0; in.max_frame_payload = 0;
    idx.frames.clear();
    if (headers_only && in.num_frames == 0) return MJPEG423_OK;
    // num_frames is untrusted: a frame takes at least its 16-byte header, so the file length bounds the reserve
    idx.frames.reserve((size_t)std::min<uint64_t>(in.num_frames, (len - 20) / 16));
    uint64_t off = 20;
    for (uint32_t f = 0; f < in.num_frames; f++) {
        if (off + 16 > len) { set_error("mpg: truncated at frame header " + std::to_string(f)); return MJPEG423_E_FORMAT; }
        FrameRec r;
        r.off = off; r.size = rd32(mpg + off); r.type = rd32(mpg + off + 4);
        r.ysize = rd32(mpg + off + 8); r.cbsize = rd32(mpg + off + 12);
        if (r.size < 16 || off + r.size > len || (uint64_t)r.ysize + r.cbsize > r.size - 16u || r.type > 1) {
            set_error("mpg: inconsistent sizes in frame " + std::to_string(f));
            return MJPEG423_E_FORMAT;
        }
        r.crsize = r.size - 16u - r.ysize - r.cbsize;     // implied; includes the 0-3 pad bytes
        if (r.type) in.num_pframes++;
        in.max_frame_payload = std::max<uint64_t>(in.max_frame_payload, r.size);
        idx.frames.push_back(r);
        off += r.size;
    }
    return MJPEG423_OK;
}

int build_plan(const MpgIndex& idx, uint32_t first, uint32_t n, Plan& plan) {
    const mjpeg423_b200_info& in = idx.info;
    if ((uint64_t)first + n > idx.frames.size()) { set_error("frame range exceeds num_frames"); return MJPEG423_E_ARG; }
    plan = Plan();
    plan.W = in.w_size; plan.H = in.h_size; plan.nb = (in.w_size / 8) * (in.h_size / 8);
    plan.first = first; plan.n = n;
    if (n == 0) return MJPEG423_OK;
    if (idx.frames[first].type != 0) {
        set_error("frame range starts on a P frame; start at an I frame (use the trailer / probe)");
        return MJPEG423_E_PFRAME;
    }
    if ((uint64_t)n * 3 * plan.nb >= 0xFFFFFFFFull) { set_error("too many blocks in one plan"); return MJPEG423_E_ARG; }
    plan.frames.assign(idx.frames.begin() + first, idx.frames.begin() + first + n);
    for (const FrameRec& r : plan.frames) plan.n_pframes += r.type != 0;
    plan.payload_off = plan.frames.front().off;
    plan.payload_len = plan.frames.back().off + plan.frames.back().size - plan.payload_off;
    plan.streams.reserve((size_t)n * 3);
    plan.f_seg0.resize(n + 1);
    uint32_t seg_base = 0;
    for (uint32_t f = 0; f < n; f++) {
        const FrameRec& r = plan.frames[f];
        plan.f_seg0[f] = seg_base;
        const uint32_t lens[3] = {r.ysize, r.cbsize, r.crsize};
        uint64_t so = r.off + 16 - plan.payload_off;
        for (int p = 0; p < 3; p++) {
            if (lens[p] >= MAX_STREAM_BYTES) { set_error("plane stream too large"); return MJPEG423_E_ARG; }
            StreamDesc sd;
            sd.byte_off = so; sd.byte_len = lens[p]; sd.nb = plan.nb;
            sd.seg_base = seg_base;
            sd.nseg = std::max<uint32_t>(1, (lens[p] + SEG_BYTES - 1) / SEG_BYTES);
            sd.block_base = (f * 3 + p) * plan.nb;
            sd.prev_base = r.type ? sd.block_base - 3 * plan.nb : sd.block_base;   // frame 0 is an I frame
            sd.quant_id = p ? 1 : 0; sd.ptype = (uint16_t)r.type;
            plan.streams.push_back(sd);
            seg_base += (sd.nseg + SUPER - 1) / SUPER * SUPER; so += lens[p];     // padded: see SUPER (common.cuh)
            plan.stream_bytes += lens[p];
        }
    }
    plan.f_seg0[n] = seg_base;
    return MJPEG423_OK;
}

// Cut the plan into chunks of about K frames that start on I frames, and order each chunk's streams by
// GOP depth (see Chunk in runtime.h).  For intra-only plans `ids` is simply 0, 1, 2, ...
void make_chunks(const Plan& plan, uint32_t K, std::vector<Chunk>& chunks, std::vector<uint32_t>& ids,
                 std::vector<uint32_t>& gops) {
    chunks.clear(); ids.clear(); gops.clear();
    ids.reserve((size_t)plan.n * 3);
    K = std::max<uint32_t>(1, K);
    uint32_t f = 0;
    while (f < plan.n) {
        Chunk ch;
        ch.f0 = f;
        uint32_t e = (uint32_t)std::min<uint64_t>(plan.n, (uint64_t)f + K);
        while (e < plan.n && plan.frames[e].type != 0) e++;          // never split a GOP
        ch.f1 = e;
        ch.ids_off = (uint32_t)ids.size();
        std::vector<uint32_t> depth(e - f);
        uint32_t maxd = 0;
        for (uint32_t k = f; k < e; k++) {
            depth[k - f] = plan.frames[k].type ? depth[k - f - 1] + 1 : 0;
            maxd = std::max(maxd, depth[k - f]);
        }
        for (uint32_t l = 0; l <= maxd; l++) {
            ch.level_off.push_back((uint32_t)ids.size() - ch.ids_off);
            for (uint32_t k = f; k < e; k++)
                if (depth[k - f] == l) for (uint32_t p = 0; p < 3; p++) ids.push_back(k * 3 + p);
        }
        ch.level_off.push_back((uint32_t)ids.size() - ch.ids_off);
        ch.gop_off = (uint32_t)gops.size();
        for (uint32_t k = f; k < e; k++)
            if (depth[k - f] == 0) gops.push_back(k - f);
        ch.n_gops = (uint32_t)gops.size() - ch.gop_off;
        gops.push_back(e - f);
        chunks.push_back(std::move(ch));
        f = e;
    }
}

}  // namespace mj

using namespace mj;

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return MJPEG423_OK;
    release();
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; cap = 0; cuda_fail(e, "cudaMalloc"); return MJPEG423_E_NOMEM; }
    cap = want;
    return MJPEG423_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

// ---- default tables: LIB/common/tables.c:13-32 -------------------------------------------------------
extern "C" {
dct_block_t Yquant = {{16, 11, 10, 16, 24, 40, 51, 61},     {12, 12, 14, 19, 26, 58, 60, 55},
                      {14, 13, 16, 24, 40, 57, 69, 56},     {14, 17, 22, 29, 51, 87, 80, 62},
                      {18, 22, 37, 56, 68, 109, 103, 77},   {24, 35, 55, 64, 81, 104, 113, 92},
                      {49, 64, 78, 87, 103, 121, 120, 101}, {72, 92, 95, 98, 112, 100, 103, 99}};
dct_block_t Cquant = {{17, 18, 24, 47, 99, 99, 99, 99}, {18, 21, 26, 66, 99, 99, 99, 99},
                      {24, 26, 56, 99, 99, 99, 99, 99}, {47, 66, 99, 99, 99, 99, 99, 99},
                      {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
                      {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99}};
int zigzag_table[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                        41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                        30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
}

// ---- context ----------------------------------------------------------------------------------------
extern "C" int mjpeg423_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* mjpeg423_b200_last_error(void) { return mj::last_error().c_str(); }

extern "C" void mjpeg423_b200_destroy(mjpeg423_b200_ctx* c);
extern "C" int mjpeg423_b200_create(mjpeg423_b200_ctx** out, int device) {
    if (!out) return MJPEG423_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (this library has no CPU fallback)");
        cudaGetLastError();
        return MJPEG423_E_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("bad device ordinal"); return MJPEG423_E_ARG; }
    CU(cudaSetDevice(device));
    mjpeg423_b200_ctx* c = new (std::nothrow) mjpeg423_b200_ctx();
    if (!c) { set_error("out of host memory"); return MJPEG423_E_NOMEM; }
    c->device = device;
    auto init = [&]() -> int {                       // (everything created so far is released on any failure)
        CU(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
        for (auto& ev : c->ev) CU(cudaEventCreate(&ev));
        CU(cudaMalloc(&c->d_quant, sizeof(c->h_quant)));
        return mjpeg423_b200_set_quant(c, nullptr, nullptr);
    };
    const int rc = init();
    if (rc != MJPEG423_OK) { mjpeg423_b200_destroy(c); return rc; }
    *out = c;
    return MJPEG423_OK;
}

extern "C" void mjpeg423_b200_destroy(mjpeg423_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&c->payload, &c->tables, &c->segs, &c->coef[0], &c->coef[1], &c->blkidx[0], &c->blkidx[1],
                      &c->samples[0], &c->samples[1], &c->stream_blocks, &c->misc, &c->ids, &c->in_ring[0], &c->in_ring[1], &c->out_ring[0],
                      &c->out_ring[1], &c->fstate[0], &c->fstate[1]})
        b->release();
    for (int i = 0; i < 2; i++) if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    if (c->d_quant) cudaFree(c->d_quant);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t s : {c->s_compute, c->s_in, c->s_out, c->s_aux}) if (s) cudaStreamDestroy(s);
    delete c;
}

extern "C" int mjpeg423_b200_set_option(mjpeg423_b200_ctx* c, int option, int64_t value) {
    if (!c) return MJPEG423_E_ARG;
    switch (option) {
        case MJPEG423_OPT_PROFILE: c->profile = value != 0; return MJPEG423_OK;
        case MJPEG423_OPT_STAGED:
            if (value < 0 || value > 2) { set_error("STAGED must be 0, 1 or 2"); return MJPEG423_E_ARG; }
            c->staged = (int)value;
            return MJPEG423_OK;
        case MJPEG423_OPT_CHUNK_FRAMES: c->chunk_frames = value < 0 ? 0 : (uint32_t)value; return MJPEG423_OK;
        case MJPEG423_OPT_VALIDATE: c->validate = value != 0; return MJPEG423_OK;
    }
    set_error("unknown option");
    return MJPEG423_E_ARG;
}

extern "C" int mjpeg423_b200_set_quant(mjpeg423_b200_ctx* c, const int16_t* yq, const int16_t* cq) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    std::memcpy(c->h_quant, yq ? yq : &Yquant[0][0], 128);
    std::memcpy(c->h_quant + 64, cq ? cq : &Cquant[0][0], 128);
    CU(cudaMemcpyAsync(c->d_quant, c->h_quant, sizeof(c->h_quant), cudaMemcpyHostToDevice, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    return MJPEG423_OK;
}

static int mjpeg423_b200_probe_impl(const uint8_t* mpg, size_t len, mjpeg423_b200_info* info) {
    if (!info) return MJPEG423_E_ARG;
    MpgIndex idx;
    int rc = parse_mpg(mpg, len, idx, false);
    *info = idx.info;
    return rc;
}
extern "C" int mjpeg423_b200_probe(const uint8_t* mpg, size_t len, mjpeg423_b200_info* info) {
    return mj::guard([&]() -> int { return mjpeg423_b200_probe_impl(mpg, len, info); });
}

// ---- plan upload --------------------------------------------------------------------------------------
namespace {

struct Tables {           // device addresses inside ctx->tables / ctx->segs
    StreamDesc* streams;
    uint32_t *seg_stream, *seg_entry, *seg_exit, *seg_cnt, *seg_first, *seg_dc, *stream_blocks;
    unsigned long long* fixups;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

Tables tables_of(mjpeg423_b200_ctx* c, const Plan& plan) {
    Tables t;
    t.streams = c->tables.as<StreamDesc>();
    const size_t nseg = plan.f_seg0.empty() ? 0 : plan.f_seg0.back();
    const size_t b_seg = align256(nseg * 4), b_sb = align256(plan.streams.size() * 4);
    uint8_t* sb = c->segs.as<uint8_t>();
    t.seg_entry = reinterpret_cast<uint32_t*>(sb);
    t.seg_exit = reinterpret_cast<uint32_t*>(sb + b_seg);
    t.seg_cnt = reinterpret_cast<uint32_t*>(sb + 2 * b_seg);
    t.seg_first = reinterpret_cast<uint32_t*>(sb + 3 * b_seg);
    t.seg_dc = reinterpret_cast<uint32_t*>(sb + 4 * b_seg);
    t.seg_stream = reinterpret_cast<uint32_t*>(sb + 5 * b_seg);
    t.stream_blocks = reinterpret_cast<uint32_t*>(sb + 6 * b_seg);
    t.fixups = reinterpret_cast<unsigned long long*>(sb + 6 * b_seg + b_sb);
    return t;
}

// Lays the plan's tables out in ctx->tables (uploaded) and ctx->segs (device scratch).
int upload_tables(mjpeg423_b200_ctx* c, const Plan& plan, Tables& t, cudaStream_t s) {
    int rc = c->tables.reserve(plan.streams.size() * sizeof(StreamDesc) + 256);
    if (rc) return rc;
    const size_t nseg = plan.f_seg0.empty() ? 0 : plan.f_seg0.back();
    rc = c->segs.reserve(6 * align256(nseg * 4) + align256(plan.streams.size() * 4) + 256);   // ... + 2 x u64 counters
    if (rc) return rc;
    t = tables_of(c, plan);
    CU(cudaMemcpyAsync(t.streams, plan.streams.data(), plan.streams.size() * sizeof(StreamDesc), cudaMemcpyHostToDevice, s));
    CU(launch_seg_stream(t.streams, (uint32_t)plan.streams.size(), t.seg_stream, s));
    c->chunk_K = 0;      // chunk table must be rebuilt for this plan
    return MJPEG423_OK;
}

// (Re)build the chunk table for chunk size K and upload the level-ordered stream id list.
int prepare_chunks(mjpeg423_b200_ctx* c, const Plan& plan, uint32_t K, cudaStream_t s) {
    if (c->chunk_K == K && !c->chunks.empty()) return MJPEG423_OK;
    std::vector<uint32_t> ids, gops;
    make_chunks(plan, K, c->chunks, ids, gops);
    c->gops_base = ids.size();
    ids.insert(ids.end(), gops.begin(), gops.end());                  // one upload: stream ids, then the GOP tables
    int rc = c->ids.reserve(ids.size() * 4 + 4);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->ids.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));          // `ids` is a local
    c->chunk_K = K;
    return MJPEG423_OK;
}

uint32_t max_chunk_frames(const std::vector<Chunk>& chunks) {
    uint32_t m = 0;
    for (const Chunk& ch : chunks) m = std::max(m, ch.f1 - ch.f0);
    return m;
}

uint32_t auto_chunk(const Plan& plan, uint64_t target_bytes, uint64_t bytes_per_frame) {
    uint64_t k = target_bytes / std::max<uint64_t>(1, bytes_per_frame);
    k = std::max<uint64_t>(1, std::min<uint64_t>(k, plan.n));
    return (uint32_t)k;
}

int decode_mode(const mjpeg423_b200_ctx* c, const Plan& plan) {
    (void)plan;                  // (mode 0 decodes P frames too: k_decode_fused<true> keeps the state inside the warp)
    return c->staged;
}

// Per-chunk scratch: block index (8 bytes per block) + symbol lists (SYM_STRIDE entries per segment) and,
// in the staged modes, coefficient planes.
size_t chunk_index_bytes(const Plan& plan, uint32_t f0, uint32_t f1) {
    const size_t blocks = (size_t)(f1 - f0) * 3 * plan.nb, segs = plan.f_seg0[f1] - plan.f_seg0[f0];
    return blocks * 8 + 32 + segs * (size_t)SYM_STRIDE * 4 + 256;
}
int reserve_chunk_buffers(mjpeg423_b200_ctx* c, const Plan& plan, const std::vector<Chunk>& chunks, int nbuf) {
    size_t idx_bytes = 0, frames = 0;
    for (const Chunk& ch : chunks) {
        idx_bytes = std::max(idx_bytes, chunk_index_bytes(plan, ch.f0, ch.f1));
        frames = std::max<size_t>(frames, ch.f1 - ch.f0);
    }
    const size_t blocks = frames * 3 * plan.nb;
    for (int i = 0; i < nbuf; i++) {
        int rc = c->blkidx[i].reserve(idx_bytes);
        if (rc) return rc;
        if (decode_mode(c, plan) && (rc = c->coef[i].reserve(blocks * 128))) return rc;
        // (every buffer in flight owns its sample planes too: two chunks run concurrently on s_compute / s_aux)
        if (decode_mode(c, plan) == 2 && (rc = c->samples[i].reserve(blocks * 64))) return rc;
    }
    return MJPEG423_OK;
}

EntropyJob make_job(const Plan& plan, const Tables& t, const uint8_t* d_payload_origin, uint32_t f0, uint32_t f1,
                    void* d_blkidx) {
    EntropyJob j;
    j.d_payload = d_payload_origin;
    j.d_streams = t.streams;
    j.d_seg_stream = t.seg_stream;
    j.seg_lo = plan.f_seg0[f0]; j.seg_hi = plan.f_seg0[f1];
    {   // fold END symbols into the previous step unless the chunk is dense: break-even is ~7 symbols (~64 bits) per block
        uint64_t bytes = 0;
        for (uint32_t f = f0; f < f1; f++) bytes += plan.frames[f].ysize + plan.frames[f].cbsize + plan.frames[f].crsize;
        j.fold_end = bytes * 8 < (uint64_t)(f1 - f0) * 3 * plan.nb * 64;
    }
    j.stream_lo = f0 * 3; j.n_streams = (f1 - f0) * 3;
    j.d_seg_entry = t.seg_entry; j.d_seg_exit = t.seg_exit; j.d_seg_cnt = t.seg_cnt; j.d_seg_first = t.seg_first;
    j.d_seg_dc = t.seg_dc;
    j.d_stream_blocks = t.stream_blocks; j.d_fixups = t.fixups;
    // StreamDesc.block_base is plan-relative: shift the chunk buffers back by the chunk's first block
    // ... and StreamDesc.seg_base too: same for the symbol lists
    const size_t blocks = (size_t)(f1 - f0) * 3 * plan.nb, first_block = (size_t)f0 * 3 * plan.nb;
    uint32_t* w = static_cast<uint32_t*>(d_blkidx);
    j.d_blk_info = reinterpret_cast<uint2*>(w) - first_block;
    j.d_sym = w + ((2 * blocks + 7) & ~(size_t)7);   // 32-byte aligned; entries are relative to the chunk's first segment
    j.sym_seg0 = plan.f_seg0[f0];
    return j;
}

constexpr int N_PROF = 8;   // events per profiled chunk

// Enqueue the whole decode of chunk `ch` on stream s; d_out receives its frames.  When prof != nullptr,
// events prof[0..7] bracket the kernels (sync | chain | index | decode | idct | colour).
// synced = true: the synchronisation and chain passes of the chunk's segments have been run already (resident path:
// once for the whole plan, see mjpeg423_b200_decode_resident).
int enqueue_chunk(mjpeg423_b200_ctx* c, const Plan& plan, const Tables& t, const uint8_t* d_payload_origin,
                  const Chunk& ch, int buf, void* d_out, cudaStream_t s, cudaEvent_t* prof, bool synced = false) {
    const uint32_t f0 = ch.f0, f1 = ch.f1;
    const int mode = decode_mode(c, plan);
    EntropyJob j = make_job(plan, t, d_payload_origin, f0, f1, c->blkidx[buf].p);
    if (prof) CU(cudaEventRecord(prof[0], s));
    if (!synced) CU(launch_entropy_sync(j, s));
    if (prof) CU(cudaEventRecord(prof[1], s));
    if (!synced) CU(launch_entropy_chain(j, s));
    if (prof) CU(cudaEventRecord(prof[2], s));
    CU(launch_entropy_index(j, s));
    if (prof) CU(cudaEventRecord(prof[3], s));
    c->stats.kernel_launches += synced ? 2 : 4;
    if (mode == 0) {
        const uint32_t* d_gops = nullptr;
        if (plan.n_pframes) {                    // GOP-walking variant; its scratch belongs to this chunk buffer
            int rc = c->fstate[buf].reserve(FUSED_STATE_BYTES);
            if (rc) return rc;
            d_gops = c->ids.as<uint32_t>() + c->gops_base + ch.gop_off;
        }
        CU(launch_decode_fused(j, c->d_quant, d_out, f1 - f0, plan.W, plan.H, d_gops, ch.n_gops, c->fstate[buf].p, s));
        c->stats.kernel_launches += 1;
        if (prof) for (int k = 4; k < N_PROF; k++) CU(cudaEventRecord(prof[k], s));
        return MJPEG423_OK;
    }
    int16_t* d_coef = c->coef[buf].as<int16_t>();
    int16_t* coef_origin = d_coef - (size_t)f0 * 3 * plan.nb * 64;
    const uint32_t* ids = c->ids.as<uint32_t>() + ch.ids_off;
    for (size_t l = 0; l + 1 < ch.level_off.size(); l++) {       // one launch per GOP depth
        CU(launch_decode_coef(j, ids + ch.level_off[l], ch.level_off[l + 1] - ch.level_off[l], plan.nb, c->d_quant,
                              coef_origin, s));
        c->stats.kernel_launches += 1;
    }
    if (prof) CU(cudaEventRecord(prof[4], s));
    if (mode == 2) {
        CU(launch_idct(d_coef, c->samples[buf].as<uint8_t>(), (size_t)(f1 - f0) * 3 * plan.nb, s));
        if (prof) CU(cudaEventRecord(prof[5], s));
        CU(launch_colour(c->samples[buf].as<uint8_t>(), d_out, f1 - f0, plan.W, plan.H, s));
        if (prof) { CU(cudaEventRecord(prof[6], s)); CU(cudaEventRecord(prof[7], s)); }
        c->stats.kernel_launches += 2;
    } else {
        if (prof) { CU(cudaEventRecord(prof[5], s)); CU(cudaEventRecord(prof[6], s)); }
        CU(launch_idct_colour(d_coef, d_out, f1 - f0, plan.W, plan.H, s));
        if (prof) CU(cudaEventRecord(prof[7], s));
        c->stats.kernel_launches += 1;
    }
    return MJPEG423_OK;
}

int accumulate_profile(mjpeg423_b200_ctx* c, cudaEvent_t* prof) {
    CU(cudaEventSynchronize(prof[N_PROF - 1]));
    float d[N_PROF - 1];
    for (int k = 0; k + 1 < N_PROF; k++) CU(cudaEventElapsedTime(&d[k], prof[k], prof[k + 1]));
    c->stats.sync_ms += d[0];
    c->stats.chain_ms += d[1];
    c->stats.index_ms += d[2];
    c->stats.decode_ms += d[3];
    c->stats.idct_ms += d[4];
    c->stats.colour_ms += d[5];
    c->stats.idct_colour_ms += d[6];
    return MJPEG423_OK;
}

// After all chunks: fetch the fix-up counter and (optionally) check that every stream held nb blocks.
int finish_stats(mjpeg423_b200_ctx* c, const Plan& plan, const Tables& t, cudaStream_t s) {
    unsigned long long fix[2] = {0, 0};
    CU(cudaMemcpyAsync(fix, t.fixups, 16, cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> blocks;
    if (c->validate) {
        blocks.resize(plan.streams.size());
        CU(cudaMemcpyAsync(blocks.data(), t.stream_blocks, blocks.size() * 4, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    c->stats.fixups = fix[0];
    c->stats.list_entries = fix[1];
    c->stats.segments = plan.f_seg0.back();
    c->stats.payload_bytes = plan.stream_bytes;
    c->stats.frames = plan.n;
    for (size_t i = 0; i < blocks.size(); i++)
        if (blocks[i] < plan.nb) {
            set_error("stream " + std::to_string(i % 3) + " of frame " + std::to_string(plan.first + i / 3) + " holds " +
                      std::to_string(blocks[i]) + " blocks, expected " + std::to_string(plan.nb));
            return MJPEG423_E_STREAM;
        }
    return MJPEG423_OK;
}

}  // namespace

static int mjpeg423_b200_upload_impl(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    c->have_plan = false;
    MpgIndex idx;
    int rc = parse_mpg(mpg, len, idx, false);
    if (rc) return rc;
    rc = build_plan(idx, first, n, c->plan);
    if (rc) return rc;
    if (n == 0) { c->have_plan = true; return MJPEG423_OK; }
    rc = c->payload.reserve(c->plan.payload_len + 64);       // + look-ahead pad
    if (rc) return rc;
    Tables t;
    rc = upload_tables(c, c->plan, t, c->s_compute);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->payload.p, mpg + c->plan.payload_off, c->plan.payload_len, cudaMemcpyHostToDevice, c->s_compute));
    CU(cudaMemsetAsync(c->payload.as<uint8_t>() + c->plan.payload_len, 0, 64, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    c->have_plan = true;
    return MJPEG423_OK;
}
extern "C" int mjpeg423_b200_upload(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first, uint32_t n) {
    return mj::guard([&]() -> int { return mjpeg423_b200_upload_impl(c, mpg, len, first, n); });
}

static int mjpeg423_b200_decode_resident_impl(mjpeg423_b200_ctx* c, void* d_out) {
    if (!c || !c->have_plan) { set_error("no resident job: call mjpeg423_b200_upload first"); return MJPEG423_E_ARG; }
    CU(cudaSetDevice(c->device));
    const Plan& plan = c->plan;
    c->stats = mjpeg423_b200_stats{};
    if (plan.n == 0) return MJPEG423_OK;
    if (!d_out) return MJPEG423_E_ARG;
    Tables t = tables_of(c, plan);
    // chunk size: bounded scratch (block index 6 B/block, + 128 B/block of coefficients in the staged modes)
    const size_t scratch_frame = (size_t)3 * plan.nb * (decode_mode(c, plan) ? 136 : 8) +
                                 (size_t)(plan.f_seg0.back() / plan.n + 1) * SYM_STRIDE * 4;
    const uint32_t K = c->chunk_frames ? std::min(c->chunk_frames, plan.n) : auto_chunk(plan, (uint64_t)3 << 30, scratch_frame);
    int rc = prepare_chunks(c, plan, K, c->s_compute);
    if (rc) return rc;
    const int nbuf = (c->chunks.size() > 1 && !c->profile) ? 2 : 1;
    if ((rc = reserve_chunk_buffers(c, plan, c->chunks, nbuf))) return rc;
    cudaStream_t st[2] = {c->s_compute, c->s_aux};
    cudaEvent_t ev_start = c->ev[0], ev_stop = c->ev[1], ev_fork = c->ev[2], ev_join = c->ev[3];
    cudaEvent_t* prof = c->profile ? &c->ev[4] : nullptr;
    CU(cudaMemsetAsync(t.fixups, 0, 16, st[0]));
    CU(cudaEventRecord(ev_start, st[0]));
    // The synchronisation and chain passes write only the plan-wide per-segment tables.  With more than two chunks
    // they run ONCE over the whole plan, in front of the chunks: the chain kernel is one CTA per stream and, on streams
    // that do not self-synchronise, serial inside it, so it wants every stream of the plan at once, not a chunk's
    // worth (4K all-ones, 4 chunks: 43.8 -> 30.8 ms); one big synchronisation launch also fills the machine better
    // (4K dense, 5 chunks: 24.6 -> 22.5 ms).  With two chunks the per-chunk order wins (1080p x 2000: 10.7 vs 11.0 ms):
    // the second chunk's parse kernels then have something to run beside.
    const bool global_sync = c->chunks.size() > 2;
    if (global_sync) {
        EntropyJob jall = make_job(plan, t, c->payload.as<uint8_t>(), 0, plan.n, c->blkidx[0].p);
        if (prof) CU(cudaEventRecord(prof[0], st[0]));
        CU(launch_entropy_sync(jall, st[0]));
        if (prof) CU(cudaEventRecord(prof[1], st[0]));
        CU(launch_entropy_chain(jall, st[0]));
        c->stats.kernel_launches += 2;
        if (prof) {
            CU(cudaEventRecord(prof[2], st[0]));
            CU(cudaEventSynchronize(prof[2]));
            float a = 0, b = 0;
            CU(cudaEventElapsedTime(&a, prof[0], prof[1]));
            CU(cudaEventElapsedTime(&b, prof[1], prof[2]));
            c->stats.sync_ms += a;
            c->stats.chain_ms += b;
        }
    }
    if (nbuf == 2) { CU(cudaEventRecord(ev_fork, st[0])); CU(cudaStreamWaitEvent(st[1], ev_fork, 0)); }
    for (size_t k = 0; k < c->chunks.size(); k++) {
        const Chunk& ch = c->chunks[k];
        const int b = (int)(k % nbuf);
        rc = enqueue_chunk(c, plan, t, c->payload.as<uint8_t>(), ch, b, (uint8_t*)d_out + (size_t)ch.f0 * plan.nb * 256,
                           st[b], prof, global_sync);
        if (rc) return rc;
        if (prof) { rc = accumulate_profile(c, prof); if (rc) return rc; }
    }
    if (nbuf == 2) { CU(cudaEventRecord(ev_join, st[1])); CU(cudaStreamWaitEvent(st[0], ev_join, 0)); }
    CU(cudaEventRecord(ev_stop, st[0]));
    rc = finish_stats(c, plan, t, st[0]);
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev_start, ev_stop));
    c->stats.total_ms = ms;
    return rc;
}
extern "C" int mjpeg423_b200_decode_resident(mjpeg423_b200_ctx* c, void* d_out) {
    return mj::guard([&]() -> int { return mjpeg423_b200_decode_resident_impl(c, d_out); });
}

extern "C" int mjpeg423_b200_get_stats(mjpeg423_b200_ctx* c, mjpeg423_b200_stats* s) {
    if (!c || !s) return MJPEG423_E_ARG;
    *s = c->stats;
    return MJPEG423_OK;
}

// ---- stage-level entry points on the resident job ------------------------------------------------------
// lossless_decode() of every plane of every resident frame: d_coef = n x 3 x nb x 64 int16.
static int mjpeg423_b200_resident_entropy_impl(mjpeg423_b200_ctx* c, int16_t* d_coef) {
    if (!c || !c->have_plan || !d_coef) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    const Plan& plan = c->plan;
    c->stats = mjpeg423_b200_stats{};
    if (plan.n == 0) return MJPEG423_OK;
    Tables t = tables_of(c, plan);
    cudaStream_t s = c->s_compute;
    int rc = prepare_chunks(c, plan, plan.n, s);          // one chunk: the caller's buffer holds every frame
    if (rc) return rc;
    if ((rc = c->blkidx[0].reserve(chunk_index_bytes(plan, 0, plan.n)))) return rc;
    const Chunk& ch = c->chunks[0];
    EntropyJob j = make_job(plan, t, c->payload.as<uint8_t>(), 0, plan.n, c->blkidx[0].p);
    cudaEvent_t* e = &c->ev[4];
    CU(cudaMemsetAsync(t.fixups, 0, 16, s));
    CU(cudaEventRecord(e[0], s));
    CU(launch_entropy_sync(j, s));
    CU(cudaEventRecord(e[1], s));
    CU(launch_entropy_chain(j, s));
    CU(cudaEventRecord(e[2], s));
    CU(launch_entropy_index(j, s));
    CU(cudaEventRecord(e[3], s));
    const uint32_t* ids = c->ids.as<uint32_t>() + ch.ids_off;
    for (size_t l = 0; l + 1 < ch.level_off.size(); l++) {
        CU(launch_decode_coef(j, ids + ch.level_off[l], ch.level_off[l + 1] - ch.level_off[l], plan.nb, c->d_quant, d_coef, s));
        c->stats.kernel_launches += 1;
    }
    CU(cudaEventRecord(e[4], s));
    c->stats.kernel_launches += 4;
    rc = finish_stats(c, plan, t, s);
    CU(cudaEventElapsedTime(&c->stats.total_ms, e[0], e[4]));
    CU(cudaEventElapsedTime(&c->stats.sync_ms, e[0], e[1]));
    CU(cudaEventElapsedTime(&c->stats.chain_ms, e[1], e[2]));
    CU(cudaEventElapsedTime(&c->stats.index_ms, e[2], e[3]));
    CU(cudaEventElapsedTime(&c->stats.decode_ms, e[3], e[4]));
    return rc;
}
extern "C" int mjpeg423_b200_resident_entropy(mjpeg423_b200_ctx* c, int16_t* d_coef) {
    return mj::guard([&]() -> int { return mjpeg423_b200_resident_entropy_impl(c, d_coef); });
}

namespace {
template <class F>
int timed_stage(mjpeg423_b200_ctx* c, float mjpeg423_b200_stats::*slot, F&& launch) {
    CU(cudaSetDevice(c->device));
    cudaStream_t s = c->s_compute;
    CU(cudaEventRecord(c->ev[0], s));
    CU(launch(s));
    CU(cudaEventRecord(c->ev[1], s));
    CU(cudaEventSynchronize(c->ev[1]));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    c->stats = mjpeg423_b200_stats{};
    c->stats.total_ms = ms; c->stats.kernel_launches = 1; c->stats.frames = c->plan.n;
    c->stats.*slot = ms;
    return MJPEG423_OK;
}
}  // namespace

extern "C" int mjpeg423_b200_resident_idct(mjpeg423_b200_ctx* c, const int16_t* d_coef, uint8_t* d_samples) {
    if (!c || !c->have_plan || !d_coef || !d_samples) return MJPEG423_E_ARG;
    const size_t nblk = (size_t)c->plan.n * 3 * c->plan.nb;
    return timed_stage(c, &mjpeg423_b200_stats::idct_ms, [&](cudaStream_t s) { return launch_idct(d_coef, d_samples, nblk, s); });
}
extern "C" int mjpeg423_b200_resident_colour(mjpeg423_b200_ctx* c, const uint8_t* d_samples, void* d_out) {
    if (!c || !c->have_plan || !d_samples || !d_out) return MJPEG423_E_ARG;
    return timed_stage(c, &mjpeg423_b200_stats::colour_ms,
                       [&](cudaStream_t s) { return launch_colour(d_samples, d_out, c->plan.n, c->plan.W, c->plan.H, s); });
}
extern "C" int mjpeg423_b200_resident_idct_colour(mjpeg423_b200_ctx* c, const int16_t* d_coef, void* d_out) {
    if (!c || !c->have_plan || !d_coef || !d_out) return MJPEG423_E_ARG;
    return timed_stage(c, &mjpeg423_b200_stats::idct_colour_ms, [&](cudaStream_t s) {
        return launch_idct_colour(d_coef, d_out, c->plan.n, c->plan.W, c->plan.H, s);
    });
}

// ---- end-to-end path: host .mpg -> (host | device) frames, chunked and overlapped ----------------------
static int mjpeg423_b200_decode_frames_impl(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first,
                                           uint32_t n, void* out, int out_on_device) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    MpgIndex idx;
    int rc = parse_mpg(mpg, len, idx, false);
    if (rc) return rc;
    c->have_plan = false;                       // the resident tables are about to be overwritten
    Plan& plan = c->plan;
    rc = build_plan(idx, first, n, plan);
    if (rc) return rc;
    c->stats = mjpeg423_b200_stats{};
    if (n == 0) return MJPEG423_OK;
    if (!out) return MJPEG423_E_ARG;
    Tables t;
    rc = upload_tables(c, plan, t, c->s_in);
    if (rc) return rc;
    const size_t frame_bytes = (size_t)plan.nb * 256;
    const uint32_t K = c->chunk_frames ? std::min(c->chunk_frames, n) : auto_chunk(plan, (uint64_t)512 << 20, frame_bytes);
    if ((rc = prepare_chunks(c, plan, K, c->s_in))) return rc;
    const uint32_t maxf = max_chunk_frames(c->chunks);
    if ((rc = reserve_chunk_buffers(c, plan, c->chunks, 1))) return rc;
    size_t max_in = 0;                                      // largest compressed chunk
    for (const Chunk& ch : c->chunks)
        max_in = std::max<size_t>(max_in, plan.frames[ch.f1 - 1].off + plan.frames[ch.f1 - 1].size - plan.frames[ch.f0].off);
    for (int i = 0; i < 2; i++) {
        if ((rc = c->in_ring[i].reserve(max_in + 64))) return rc;
        if (!out_on_device && (rc = c->out_ring[i].reserve((size_t)maxf * frame_bytes))) return rc;
    }
    cudaEvent_t ev_start = c->ev[0], ev_stop = c->ev[1];
    cudaEvent_t* ev_in = &c->ev[2];     // [2]: payload slot uploaded
    cudaEvent_t* ev_comp = &c->ev[4];   // [2]: chunk decoded (payload slot consumed, out slot produced)
    cudaEvent_t* ev_out = &c->ev[6];    // [2]: out slot read back
    cudaEvent_t* prof = c->profile ? &c->ev[8] : nullptr;
    CU(cudaMemsetAsync(t.fixups, 0, 16, c->s_in));
    CU(cudaEventRecord(ev_start, c->s_in));
    CU(cudaStreamWaitEvent(c->s_compute, ev_start, 0));
    for (size_t k = 0; k < c->chunks.size(); k++) {
        const Chunk& ch = c->chunks[k];
        const uint32_t f0 = ch.f0, f1 = ch.f1;
        const int b = (int)(k & 1);
        const uint64_t in_off = plan.frames[f0].off;
        const size_t in_len = plan.frames[f1 - 1].off + plan.frames[f1 - 1].size - in_off;
        // upload: the slot is free once the chunk that used it two iterations ago has been decoded
        if (k >= 2) CU(cudaStreamWaitEvent(c->s_in, ev_comp[b], 0));
        CU(cudaMemcpyAsync(c->in_ring[b].p, mpg + in_off, in_len, cudaMemcpyHostToDevice, c->s_in));
        CU(cudaMemsetAsync(c->in_ring[b].as<uint8_t>() + in_len, 0, 64, c->s_in));
        CU(cudaEventRecord(ev_in[b], c->s_in));
        // decode
        CU(cudaStreamWaitEvent(c->s_compute, ev_in[b], 0));
        if (!out_on_device && k >= 2) CU(cudaStreamWaitEvent(c->s_compute, ev_out[b], 0));
        uint8_t* d_dst = out_on_device ? (uint8_t*)out + (size_t)f0 * frame_bytes : c->out_ring[b].as<uint8_t>();
        const uint8_t* origin = c->in_ring[b].as<uint8_t>() - (in_off - plan.payload_off);
        rc = enqueue_chunk(c, plan, t, origin, ch, 0, d_dst, c->s_compute, prof);
        if (rc) return rc;
        CU(cudaEventRecord(ev_comp[b], c->s_compute));
        if (prof) { rc = accumulate_profile(c, prof); if (rc) return rc; }
        // read back
        if (!out_on_device) {
            CU(cudaStreamWaitEvent(c->s_out, ev_comp[b], 0));
            CU(cudaMemcpyAsync((uint8_t*)out + (size_t)f0 * frame_bytes, d_dst, (size_t)(f1 - f0) * frame_bytes,
                               cudaMemcpyDeviceToHost, c->s_out));
            CU(cudaEventRecord(ev_out[b], c->s_out));
        }
    }
    cudaStream_t last = out_on_device ? c->s_compute : c->s_out;
    CU(cudaEventRecord(ev_stop, last));
    CU(cudaStreamSynchronize(last));
    CU(cudaStreamSynchronize(c->s_compute));
    rc = finish_stats(c, plan, t, c->s_compute);
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev_start, ev_stop));
    c->stats.total_ms = ms;
    return rc;
}
extern "C" int mjpeg423_b200_decode_frames(mjpeg423_b200_ctx* c, const uint8_t* mpg, size_t len, uint32_t first,
                                           uint32_t n, void* out, int out_on_device) {
    return mj::guard([&]() -> int { return mjpeg423_b200_decode_frames_impl(c, mpg, len, first, n, out, out_on_device); });
}

// ---- memory helpers ------------------------------------------------------------------------------------
extern "C" void* mjpeg423_b200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void mjpeg423_b200_host_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" void* mjpeg423_b200_device_alloc(mjpeg423_b200_ctx* c, size_t bytes) {
    if (!c || cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void mjpeg423_b200_device_free(mjpeg423_b200_ctx* c, void* p) {
    if (c && p) { cudaSetDevice(c->device); cudaFree(p); }
}
extern "C" int mjpeg423_b200_memcpy_d2h(mjpeg423_b200_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    return MJPEG423_OK;
}
extern "C" int mjpeg423_b200_memcpy_h2d(mjpeg423_b200_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    return MJPEG423_OK;
}
extern "C" int mjpeg423_b200_sync(mjpeg423_b200_ctx* c) {
    if (!c) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    return MJPEG423_OK;
}
extern "C" int mjpeg423_b200_hash_frames(mjpeg423_b200_ctx* c, const void* d_frames, uint64_t frame_bytes, uint32_t n,
                                         uint64_t* hashes) {
    if (!c || !hashes || (frame_bytes & 7)) return MJPEG423_E_ARG;
    CU(cudaSetDevice(c->device));
    int rc = c->misc.reserve((size_t)n * 8 + 8);
    if (rc) return rc;
    CU(launch_hash_frames(d_frames, frame_bytes, n, c->misc.as<unsigned long long>(), c->s_compute));
    CU(cudaMemcpyAsync(hashes, c->misc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->s_compute));
    CU(cudaStreamSynchronize(c->s_compute));
    return MJPEG423_OK;
}
