// multi.cu -- the callers' side of the hot path (SURVEY.md 8 rows e and f2):
//   * the I-frame trailer index of a .mpg (reader: LIB/decoder/mjpeg423_decoder.c:78-86, C1/main.c:77-113) and the
//     player's seek rules on it (C0/playback.c:157-227: fast-forward / rewind land on the first I frame at least 108
//     frames away);
//   * frame-range sharding of one decode over the GPUs of a box: one host thread per device, pieces cut on I frames
//     (from the index), every device writing its own slice of the output, no inter-GPU traffic; the input may be a SET
//     of .mpg files (the container's offsets are 32-bit, so long streams are split into files below 4 GiB:
//     LIB/encoder/mjpeg423_encoder.c:67,209,222-225).
// (LIB = /root/reference/core0/software/common/libs/mjpeg423, C0 / C1 = /root/reference/core{0,1}/software.)
#include <algorithm>
#include <mutex>
#include <thread>

#include "runtime.h"

using namespace mj;

namespace {

inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

int index_impl(const uint8_t* mpg, size_t len, iframe_trailer_t* out, uint32_t cap, uint32_t* n_iframes, int* trailer_ok) {
    if (!n_iframes) return MJPEG423_E_ARG;
    *n_iframes = 0;
    if (trailer_ok) *trailer_ok = 0;
    MpgIndex idx;
    int rc = parse_mpg(mpg, len, idx, false);
    if (rc) return rc;
    // The truth is the header walk: every frame of type 0, at the file offset of its frame_size field.
    std::vector<iframe_trailer_t> walk;
    for (size_t f = 0; f < idx.frames.size(); f++)
        if (idx.frames[f].type == 0) {
            if (idx.frames[f].off > 0xFFFFFFFFull) { set_error("mpg: I frame beyond 4 GiB (split the stream into files)"); return MJPEG423_E_FORMAT; }
            walk.push_back(iframe_trailer_t{(uint32_t)f, (uint32_t)idx.frames[f].off});
        }
    // The trailer (num_iframes x {frame_index, frame_position} at 20 + payload_size) must say the same.
    const uint64_t toff = 20ull + idx.info.payload_size;
    bool ok = idx.info.num_iframes == walk.size() && toff + 8ull * walk.size() <= len;
    for (size_t i = 0; ok && i < walk.size(); i++)
        ok = rd32(mpg + toff + 8 * i) == walk[i].frame_index && rd32(mpg + toff + 8 * i + 4) == walk[i].frame_position;
    if (trailer_ok) *trailer_ok = ok ? 1 : 0;
    *n_iframes = (uint32_t)walk.size();
    if (out) {
        if (cap < walk.size()) { set_error("index: output holds fewer than num_iframes entries"); return MJPEG423_E_ARG; }
        std::copy(walk.begin(), walk.end(), out);
    }
    return MJPEG423_OK;
}

// One logical stream over a set of files.
struct ShardSet {
    std::vector<MpgIndex> idx;
    std::vector<uint64_t> first;                 // logical index of every file's first frame, then the total
    mjpeg423_b200_info info{};
};
int open_shards(const mjpeg423_b200_shard* shards, uint32_t n_shards, ShardSet& set) {
    if (!shards || n_shards == 0) { set_error("multi: no input file"); return MJPEG423_E_ARG; }
    set.idx.resize(n_shards);
    set.first.assign(1, 0);
    for (uint32_t s = 0; s < n_shards; s++) {
        int rc = parse_mpg(shards[s].mpg, shards[s].len, set.idx[s], false);
        if (rc) return rc;
        const mjpeg423_b200_info& in = set.idx[s].info;
        if (s == 0) set.info = in;
        else if (in.w_size != set.info.w_size || in.h_size != set.info.h_size) { set_error("multi: the files differ in geometry"); return MJPEG423_E_ARG; }
        if (!set.idx[s].frames.empty() && set.idx[s].frames[0].type != 0) { set_error("multi: a file starts on a P frame"); return MJPEG423_E_PFRAME; }
        set.first.push_back(set.first.back() + set.idx[s].frames.size());
    }
    return MJPEG423_OK;
}
bool is_iframe(const ShardSet& set, uint64_t f) {
    const size_t s = std::upper_bound(set.first.begin(), set.first.end(), f) - set.first.begin() - 1;
    return set.idx[s].frames[f - set.first[s]].type == 0;
}

// Contexts of the sharded decode: one per slot of the caller's device list, kept for the life of the process.
std::mutex g_multi_mu;
std::vector<std::pair<int, mjpeg423_b200_ctx*>> g_multi_ctx;
mjpeg423_b200_ctx* slot_ctx(size_t slot, int device) {
    if (g_multi_ctx.size() <= slot) g_multi_ctx.resize(slot + 1, {-1, nullptr});
    auto& e = g_multi_ctx[slot];
    if (e.second && e.first != device) { mjpeg423_b200_destroy(e.second); e.second = nullptr; }
    if (!e.second) {
        if (mjpeg423_b200_create(&e.second, device) != MJPEG423_OK) return nullptr;
        e.first = device;
    }
    return e.second;
}

int decode_multi_impl(const int* devices, int n_dev, const mjpeg423_b200_shard* shards, uint32_t n_shards, uint64_t first,
                      uint64_t n, void* out, uint64_t* cuts) {
    if (n_dev <= 0 || n_dev > 64) { set_error("multi: bad device count"); return MJPEG423_E_ARG; }
    ShardSet set;
    int rc = open_shards(shards, n_shards, set);
    if (rc) return rc;
    const uint64_t total = set.first.back();
    if (first > total || n > total - first) { set_error("multi: frame range exceeds the stream"); return MJPEG423_E_ARG; }
    if (n == 0) return MJPEG423_OK;
    if (!out) return MJPEG423_E_ARG;
    if (!is_iframe(set, first)) { set_error("frame range starts on a P frame; start at an I frame (use mjpeg423_b200_index)"); return MJPEG423_E_PFRAME; }
    // Piece k = [cut[k], cut[k+1]): the even split, every cut moved forward to the next I frame (GOPs are never split).
    std::vector<uint64_t> cut(n_dev + 1);
    cut[0] = first; cut[n_dev] = first + n;
    for (int k = 1; k < n_dev; k++) {
        uint64_t c = std::max(cut[k - 1], first + n * k / n_dev);
        while (c < first + n && !is_iframe(set, c)) c++;
        cut[k] = c;
    }
    if (cuts) std::copy(cut.begin(), cut.end(), cuts);
    const uint64_t frame_bytes = set.info.frame_bytes;
    std::lock_guard<std::mutex> lk(g_multi_mu);               // the slot contexts serve one sharded decode at a time
    std::vector<mjpeg423_b200_ctx*> ctx(n_dev, nullptr);
    for (int k = 0; k < n_dev; k++)
        if (cut[k + 1] > cut[k] && !(ctx[k] = slot_ctx((size_t)k, devices ? devices[k] : k))) return MJPEG423_E_CUDA;
    std::vector<int> rcs(n_dev, MJPEG423_OK);
    std::vector<std::string> errs(n_dev);
    std::vector<std::thread> th;
    for (int k = 0; k < n_dev; k++) {
        if (cut[k + 1] <= cut[k]) continue;
        th.emplace_back([&, k]() {
            rcs[k] = guard([&]() -> int {
                for (size_t s = 0; s + 1 < set.first.size(); s++) {        // the files the piece touches
                    const uint64_t lo = std::max(cut[k], set.first[s]), hi = std::min(cut[k + 1], set.first[s + 1]);
                    if (lo >= hi) continue;
                    int r = mjpeg423_b200_decode_frames(ctx[k], shards[s].mpg, shards[s].len, (uint32_t)(lo - set.first[s]),
                                                        (uint32_t)(hi - lo), (uint8_t*)out + (lo - first) * frame_bytes, 0);
                    if (r) return r;
                }
                return MJPEG423_OK;
            });
            if (rcs[k]) errs[k] = mjpeg423_b200_last_error();      // (thread-local: carried back to the caller's thread)
        });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < n_dev; k++)
        if (rcs[k]) { set_error("device slot " + std::to_string(k) + ": " + errs[k]); return rcs[k]; }
    return MJPEG423_OK;
}

}  // namespace

extern "C" int mjpeg423_b200_index(const uint8_t* mpg, size_t len, iframe_trailer_t* out, uint32_t cap, uint32_t* n_iframes,
                                   int* trailer_ok) {
    return guard([&]() -> int { return index_impl(mpg, len, out, cap, n_iframes, trailer_ok); });
}

// The I frame a decode of `frame` has to start from (direction <= 0: the last one at or before it), or the next place a
// player can jump to (direction > 0: the first one at or after it).  Returns the position in idx[], -1 if there is none.
extern "C" int mjpeg423_b200_seek_iframe(const iframe_trailer_t* idx, uint32_t n, uint32_t frame, int direction) {
    if (!idx || n == 0) return -1;
    const iframe_trailer_t* e = idx + n;
    if (direction > 0) {
        const iframe_trailer_t* p = std::lower_bound(idx, e, frame, [](const iframe_trailer_t& a, uint32_t f) { return a.frame_index < f; });
        return p == e ? -1 : (int)(p - idx);
    }
    const iframe_trailer_t* p = std::upper_bound(idx, e, frame, [](uint32_t f, const iframe_trailer_t& a) { return f < a.frame_index; });
    return p == idx ? -1 : (int)(p - idx) - 1;
}

// fastForwardVideo(), C0/playback.c:157-194: nothing happens with fewer than 120 frames left (-1); otherwise the first
// I frame at least 108 frames ahead (the reference starts its search at entry current / 24, COMMON/config.h:54 -- here
// the search starts where the index says, so streams with denser I frames land on the same frame).
extern "C" int mjpeg423_b200_fast_forward(const iframe_trailer_t* idx, uint32_t n, uint32_t num_frames, uint32_t current) {
    if (!idx || n == 0 || current >= num_frames || num_frames - current < 120) return -1;
    for (uint32_t i = 0; i < n; i++)
        if (idx[i].frame_index >= current && idx[i].frame_index - current >= 108) return (int)i;
    return (int)n - 1;                                   // (the reference would run off its table here)
}
// rewindVideo(), C0/playback.c:196-227: less than 120 frames from the start -> the first I frame; otherwise the last I
// frame at least 108 frames back.
extern "C" int mjpeg423_b200_rewind(const iframe_trailer_t* idx, uint32_t n, uint32_t current) {
    if (!idx || n == 0) return -1;
    if (current < 120) return 0;
    for (uint32_t i = n; i-- > 0;)
        if (idx[i].frame_index <= current && current - idx[i].frame_index >= 108) return (int)i;
    return 0;
}

extern "C" int mjpeg423_b200_decode_frames_multi(const int* devices, int n_dev, const mjpeg423_b200_shard* shards,
                                                 uint32_t n_shards, uint64_t first, uint64_t n, void* out, uint64_t* cuts) {
    return guard([&]() -> int { return decode_multi_impl(devices, n_dev, shards, n_shards, first, n, out, cuts); });
}
