// synth_encoder.cpp -- from-spec MJPEG423 stream producer for tests and the bench (host only, no CUDA).
//
// The reference's encoder (LIB/encoder/*, LIB = /root/reference/core0/software/common/libs/mjpeg423)
// is out of scope as a product feature but it IS the format specification.
// This file restates that specification so the bench can make
// .mpg streams of any size in memory:
//   colour   Y/Cb/Cr from BGRA in double, truncated to uint8     LIB/encoder/rgb_to_ycbcr.c:58-70
//   FDCT     LL&M forward DCT on unsigned samples, output x8      LIB/encoder/fdct.c:17-161
//   quantise round(coef / q) in double; I: DC differential vs the previous block of the plane,
//            P: every level differential vs the same level of the previous frame
//                                                                   LIB/encoder/quantize.c:16-42
//   entropy  4-bit size (DC) / 4-bit run + 4-bit size (AC) + JPEG VLI amplitude, ZRL = F0,
//            END = 00 only when the last non-zero zig-zag position is < 63
//                                                                   LIB/encoder/lossless_encode.c:30-138
//   container  20-byte header, 16-byte frame headers, frames padded to x4, I-frame trailer,
//            512 pad bytes                                          LIB/encoder/mjpeg423_encoder.c:82-88,188-225
// Deliberate difference: the final partial byte of a plane stream is flushed correctly (the
// reference writes the wrong byte of its bit buffer, lossless_encode.c:80-83).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const int16_t kYquant[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                             14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                             18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                             49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int16_t kCquant[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                             24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                             99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                             99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

struct ZigZag {
    uint8_t nat[64];
    ZigZag() {
        int r = 0, c = 0;
        for (int k = 0; k < 64; k++) {
            nat[k] = uint8_t(r * 8 + c);
            if ((r + c) & 1) { if (r == 7) c++; else if (c == 0) r++; else { r++; c--; } }
            else             { if (c == 7) r++; else if (r == 0) c++; else { r--; c++; } }
        }
    }
};
const ZigZag kZZ;

inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }

// 8-point forward LL&M butterfly shared by both passes; `even_shift_up` selects pass 1 (rows)
// where the two plain sums are scaled UP by PASS1_BITS and the rotated terms descaled by 11,
// vs pass 2 (columns) where sums are descaled by 5 and rotated terms by 18.
inline void fdct8(const int32_t in[8], int32_t out[8], bool pass1) {
    int32_t s07 = in[0] + in[7], d07 = in[0] - in[7];
    int32_t s16 = in[1] + in[6], d16 = in[1] - in[6];
    int32_t s25 = in[2] + in[5], d25 = in[2] - in[5];
    int32_t s34 = in[3] + in[4], d34 = in[3] - in[4];
    int32_t e0 = s07 + s34, e3 = s07 - s34, e1 = s16 + s25, e2 = s16 - s25;
    const int rot = pass1 ? 11 : 18;
    out[0] = pass1 ? (e0 + e1) << 2 : descale(e0 + e1, 5);
    out[4] = pass1 ? (e0 - e1) << 2 : descale(e0 - e1, 5);
    int32_t z1 = (e2 + e3) * 4433;
    out[2] = descale(z1 + e3 * 6270, rot);
    out[6] = descale(z1 - e2 * 15137, rot);
    int32_t y1 = d34 + d07, y2 = d25 + d16, y3 = d34 + d16, y4 = d25 + d07;
    int32_t y5 = (y3 + y4) * 9633;
    int32_t t4 = d34 * 2446, t5 = d25 * 16819, t6 = d16 * 25172, t7 = d07 * 12299;
    y1 *= -7373; y2 *= -20995;
    y3 = y3 * -16069 + y5; y4 = y4 * -3196 + y5;
    out[7] = descale(t4 + y1 + y3, rot);
    out[5] = descale(t5 + y2 + y4, rot);
    out[3] = descale(t6 + y2 + y3, rot);
    out[1] = descale(t7 + y1 + y4, rot);
}

void fdct_block(const uint8_t* px, int16_t* coef) {
    int32_t in[8], out[8];
    int16_t tmp[64];  // pass-1 results are stored as DCTELEM (int16) by the reference
    for (int r = 0; r < 8; r++) {
        for (int c = 0; c < 8; c++) in[c] = px[r * 8 + c];
        fdct8(in, out, true);
        for (int c = 0; c < 8; c++) tmp[r * 8 + c] = int16_t(out[c]);
    }
    for (int c = 0; c < 8; c++) {
        for (int r = 0; r < 8; r++) in[r] = tmp[r * 8 + c];
        fdct8(in, out, false);
        for (int r = 0; r < 8; r++) coef[r * 8 + c] = int16_t(out[r]);
    }
}

struct BitWriter {
    std::vector<uint8_t>& out;
    uint64_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(uint32_t bits, int len) {
        if (!len) return;
        acc = (acc << len) | (bits & ((1u << len) - 1u));
        n += len;
        while (n >= 8) { out.push_back(uint8_t(acc >> (n - 8))); n -= 8; }
    }
    void flush() { if (n) { out.push_back(uint8_t(acc << (8 - n))); n = 0; } }
};

inline int vli_size(int v) { int a = v < 0 ? -v : v, s = 0; while (a) { s++; a >>= 1; } return s; }
inline uint32_t vli_bits(int v, int s) { return uint32_t(v > 0 ? v : v - 1) & ((1u << s) - 1u); }

// levels: nb x 64 natural-order quantised values exactly as they go on the wire
// (DC already differential for I frames; everything differential for P frames).
void entropy_encode(const int16_t* levels, size_t nb, std::vector<uint8_t>& out) {
    BitWriter bw(out);
    for (size_t b = 0; b < nb; b++) {
        const int16_t* v = levels + b * 64;
        int s = vli_size(v[0]);
        bw.put(uint32_t(s), 4);
        bw.put(vli_bits(v[0], s), s);
        int last = 63;
        while (last > 0 && v[kZZ.nat[last]] == 0) last--;
        int run = 0;
        for (int k = 1; k <= last; k++) {
            int x = v[kZZ.nat[k]];
            if (x == 0) {
                if (++run == 16) { bw.put(0xF0, 8); run = 0; }
                continue;
            }
            s = vli_size(x);
            bw.put(uint32_t(run), 4);
            bw.put(uint32_t(s), 4);
            bw.put(vli_bits(x, s), s);
            run = 0;
        }
        if (last < 63) bw.put(0, 8);
    }
    bw.flush();
}

inline int16_t quantise(int16_t c, int16_t q) { return int16_t(std::round(double(c) / double(q))); }

struct PlaneState { std::vector<int16_t> prev; };  // previous frame's absolute levels (for P frames)

// Encode one plane of block-major samples; appends the stream to `out`.
void encode_plane(const uint8_t* samp, size_t nb, const int16_t* q, bool P, PlaneState& st,
                  std::vector<uint8_t>& out) {
    std::vector<int16_t> wire(nb * 64);
    if (st.prev.size() != nb * 64) st.prev.assign(nb * 64, 0);
    int16_t coef[64];
    int16_t dc_prev = 0;
    for (size_t b = 0; b < nb; b++) {
        fdct_block(samp + b * 64, coef);
        int16_t* w = &wire[b * 64];
        int16_t* pv = &st.prev[b * 64];
        for (int k = 0; k < 64; k++) {
            int16_t lv = quantise(coef[k], q[k]);
            if (P) w[k] = int16_t(lv - pv[k]);
            else if (k == 0) { w[0] = int16_t(lv - dc_prev); dc_prev = lv; }
            else w[k] = lv;
            pv[k] = lv;
        }
    }
    entropy_encode(wire.data(), nb, out);
}

void bgra_to_planes(const uint8_t* bgra, uint32_t W, uint32_t H, uint8_t* samp) {
    size_t nb = size_t(W / 8) * (H / 8), wb = W / 8;
    for (size_t b = 0; b < nb; b++) {
        size_t by = b / wb * 8, bx = b % wb * 8;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                const uint8_t* p = bgra + ((by + y) * W + bx + x) * 4;
                double B = p[0], G = p[1], R = p[2];
                samp[b * 64 + y * 8 + x] = uint8_t(0.299 * R + 0.587 * G + 0.114 * B);
                samp[(nb + b) * 64 + y * 8 + x] = uint8_t(-0.168736 * R - 0.331264 * G + 0.5 * B + 128);
                samp[(2 * nb + b) * 64 + y * 8 + x] = uint8_t(0.5 * R - 0.418688 * G - 0.081312 * B + 128);
            }
    }
}

inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 24));
}

struct FrameState { PlaneState y, cb, cr; };

// Appends one frame record (16-byte header + Y|Cb|Cr streams + pad to x4) to `rec`.
void encode_frame(const uint8_t* bgra, uint32_t W, uint32_t H, const int16_t* yq, const int16_t* cq,
                  bool P, FrameState& st, std::vector<uint8_t>& rec) {
    size_t nb = size_t(W / 8) * (H / 8);
    std::vector<uint8_t> samp(3 * nb * 64), ys, cbs, crs;
    bgra_to_planes(bgra, W, H, samp.data());
    encode_plane(samp.data(), nb, yq, P, st.y, ys);
    encode_plane(samp.data() + nb * 64, nb, cq, P, st.cb, cbs);
    encode_plane(samp.data() + 2 * nb * 64, nb, cq, P, st.cr, crs);
    uint32_t fsz = uint32_t(16 + ys.size() + cbs.size() + crs.size());
    fsz = (fsz + 3u) & ~3u;
    size_t at = rec.size();
    put32(rec, fsz); put32(rec, P ? 1u : 0u); put32(rec, uint32_t(ys.size())); put32(rec, uint32_t(cbs.size()));
    rec.insert(rec.end(), ys.begin(), ys.end());
    rec.insert(rec.end(), cbs.begin(), cbs.end());
    rec.insert(rec.end(), crs.begin(), crs.end());
    rec.resize(at + fsz, 0);
}

}  // namespace

extern "C" {

// Procedural test picture: colour ramps + per-pixel LCG noise of amplitude `amp`
// (0 = none, 256 = full range), shifted by a per-frame phase so frames differ.  `flat_rows` > 0
// blanks (sets to mid-grey) the top `flat_rows` pixel rows: the all-zero-run adversarial case of
// SURVEY.md 7.3 H1 (flat_rows >= H gives a completely flat frame).
void mjpeg423_synth_frame(uint32_t W, uint32_t H, uint32_t frame_index, uint32_t amp, uint32_t flat_rows,
                          uint8_t* bgra) {
    uint32_t s = 0x423u + frame_index;
    uint32_t px = (7u * frame_index) % W, py = (3u * frame_index) % H;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            s = s * 1664525u + 1013904223u;
            uint32_t n = amp ? ((s >> 16) * amp) >> 16 : 0;
            uint32_t xx = (x + px) % W, yy = (y + py) % H;
            uint8_t* p = bgra + (size_t(y) * W + x) * 4;
            if (y < flat_rows) { p[0] = p[1] = p[2] = 128; p[3] = 0; continue; }
            p[2] = uint8_t((255u * xx / W + n) & 255u);
            p[1] = uint8_t((255u * yy / H + n) & 255u);
            p[0] = uint8_t((255u * (xx + yy) / (W + H) + n) & 255u);
            p[3] = 0;
        }
}

// Encode `n` BGRA frames (contiguous, W*H*4 bytes each) into a complete .mpg.  Frame f is an I frame
// when f % gop == 0 (gop <= 1: all intra), else a P frame.  Returns the file size, or 0 if `cap` is
// too small (call with out == NULL to get the size).  yq/cq NULL = default tables.
size_t mjpeg423_encode_mpg(uint32_t W, uint32_t H, uint32_t n, const uint8_t* frames, const int16_t* yq,
                           const int16_t* cq, uint32_t gop, uint8_t* out, size_t cap) {
    if (!W || !H || (W & 7) || (H & 7)) return 0;
    if (!yq) yq = kYquant;
    if (!cq) cq = kCquant;
    std::vector<uint8_t> file;
    put32(file, n); put32(file, W); put32(file, H); put32(file, 0); put32(file, 0);
    std::vector<uint32_t> trailer;
    FrameState st;
    for (uint32_t f = 0; f < n; f++) {
        bool P = gop > 1 && (f % gop) != 0;
        if (!P) { trailer.push_back(f); trailer.push_back(uint32_t(file.size())); }
        encode_frame(frames + size_t(f) * W * H * 4, W, H, yq, cq, P, st, file);
    }
    uint32_t payload = uint32_t(file.size() - 20);
    for (uint32_t t : trailer) put32(file, t);
    file.resize(file.size() + 512, 0);
    uint32_t ni = uint32_t(trailer.size() / 2);
    std::memcpy(&file[12], &ni, 4);
    std::memcpy(&file[16], &payload, 4);
    if (out && file.size() <= cap) std::memcpy(out, file.data(), file.size());
    else if (out) return 0;
    return file.size();
}

// Bench-scale generator: an intra-only .mpg of `n` frames made from `n_unique` distinct procedural
// frames (frame f carries picture f % n_unique), encoded on `nthreads` host threads.  Every frame
// record is a separate copy in the file, so the decoder reads distinct bytes for every frame.
// Returns the file size; 0 if cap is too small or the file would exceed the format's 32-bit offsets
// (SURVEY.md H5).  Call with out == NULL to size the buffer.
size_t mjpeg423_synth_mpg(uint32_t W, uint32_t H, uint32_t n, uint32_t n_unique, uint32_t amp,
                          uint32_t flat_rows, const int16_t* yq, const int16_t* cq, uint8_t* out, size_t cap,
                          int nthreads) {
    if (!W || !H || (W & 7) || (H & 7) || !n) return 0;
    if (!yq) yq = kYquant;
    if (!cq) cq = kCquant;
    if (n_unique == 0 || n_unique > n) n_unique = n;
    if (nthreads < 1) nthreads = 1;
    // The sizing call (out == NULL) and the filling call share one encode through this cache.
    struct Cache { std::vector<uint32_t> key; std::vector<std::vector<uint8_t>> recs; };
    static Cache cache;
    std::vector<uint32_t> key = {W, H, n_unique, amp, flat_rows};
    for (int k = 0; k < 64; k++) { key.push_back(uint32_t(yq[k])); key.push_back(uint32_t(cq[k])); }
    if (cache.key != key) { cache.recs.assign(n_unique, {}); cache.key.clear(); }
    std::vector<std::vector<uint8_t>>& recs = cache.recs;
    std::atomic<uint32_t> next{cache.key == key ? n_unique : 0u};
    auto work = [&]() {
        std::vector<uint8_t> pic(size_t(W) * H * 4);
        for (;;) {
            uint32_t u = next.fetch_add(1);
            if (u >= n_unique) break;
            mjpeg423_synth_frame(W, H, u, amp, flat_rows, pic.data());
            FrameState st;
            encode_frame(pic.data(), W, H, yq, cq, false, st, recs[u]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    cache.key = key;
    uint64_t payload = 0;
    for (uint32_t f = 0; f < n; f++) payload += recs[f % n_unique].size();
    uint64_t total = 20 + payload + uint64_t(n) * 8 + 512;
    if (total > 0xFFFFFFFFull) return 0;
    if (!out) return size_t(total);
    if (total > cap) return 0;
    uint32_t hdr[5] = {n, W, H, n, uint32_t(payload)};
    std::memcpy(out, hdr, 20);
    size_t off = 20;
    std::vector<uint32_t> trailer(size_t(n) * 2);
    for (uint32_t f = 0; f < n; f++) {
        const auto& r = recs[f % n_unique];
        trailer[2 * f] = f; trailer[2 * f + 1] = uint32_t(off);
        std::memcpy(out + off, r.data(), r.size());
        off += r.size();
    }
    std::memcpy(out + off, trailer.data(), trailer.size() * 4);
    off += trailer.size() * 4;
    std::memset(out + off, 0, 512);
    return size_t(total);
}

}  // extern "C"
