// common.cuh -- shared device helpers for the MJPEG423 kernels (sm_100a).
//
// LIB = /root/reference/core0/software/common/libs/mjpeg423.  Every arithmetic helper here restates
// one reference macro or statement and cites it; results must be bit-identical to the C reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mj {

// ------------------------------------------------------------------------------------------------
// Layout constants of the entropy stage.
// A plane bitstream is cut into fixed SEG_BYTES segments; a segment owns the blocks whose FIRST bit
// lies inside it.  CP_BITS is the spacing of the merge checkpoints inside a segment.
// ------------------------------------------------------------------------------------------------
constexpr int SEG_BYTES = 512;
constexpr int SEG_BITS = SEG_BYTES * 8;
constexpr int CP_BITS = 512;
constexpr int NCP = SEG_BITS / CP_BITS;        // checkpoints per segment (the last one is the segment end)
constexpr int SUPER = 4;                       // segments parsed by one lane of the synchronisation pass; every stream's
                                               // range of global segment numbers is padded to a multiple of it
constexpr int ENT_TPB = 128;                   // threads per CTA in the entropy kernels
constexpr int MIN_BLOCK_BITS = 12;             // DC size 0 + END
constexpr uint32_t RUNAWAY_BITS = 8192;        // parse guard for non-conforming / speculative garbage
constexpr uint32_t MAX_STREAM_BYTES = 1u << 28; // bit positions are 32-bit
constexpr uint32_t MAX_PLANE_BLOCKS = 1u << 24; // blocks per plane (W/8 * H/8): 32768 x 32768 pixels
// Symbol list: the index pass writes every coded AC coefficient of a segment's blocks as one 32-bit entry
// (zig-zag index | amplitude << 16) into a fixed-stride region.  A segment owns at most SEG_BITS/9 coded
// symbols (>= 9 bits each) that start inside it plus the rest of its last block (<= 63 AC coefficients).
constexpr uint32_t SYM_STRIDE = (SEG_BYTES * 8 / 9 + 63 + 7) / 8 * 8;   // entries per segment, a multiple of 8
// Block index entry (uint2): .x = position of the block's first list entry -- always inside its segment's
// region, so .x / SYM_STRIDE names the segment (whose DC predictor the decode kernels add) -- or BLK_NO_SEG
// for a block the stream does not hold; .y = segment-relative DC level | list entries << 16.
constexpr uint32_t BLK_NO_SEG = 0xFFFFFFFFu;

// One plane bitstream of one frame (built by the host from the 16-byte frame headers,
// LIB/decoder/mjpeg423_decoder.c:94-107).
struct StreamDesc {
    uint64_t byte_off;     // offset of the stream's first byte in the device payload buffer
    uint32_t byte_len;     // Ysize / Cbsize / implied Crsize
    uint32_t nb;           // blocks per plane
    uint32_t seg_base;     // index of this stream's first segment in the per-segment arrays
    uint32_t nseg;         // ceil(byte_len / SEG_BYTES), at least 1
    uint32_t block_base;   // index (in 64-coefficient blocks) of this plane in the coefficient buffer
    uint32_t prev_base;    // P frames: block index of the same plane of the previous frame (the state
                           // the deltas are added to, LIB/decoder/lossless_decode.c:90-92,121-123)
    uint16_t quant_id;     // 0 = luminance table, 1 = chrominance table
    uint16_t ptype;        // 0 = I frame (zero-fill, DC differential), 1 = P frame (accumulate)
};

// ------------------------------------------------------------------------------------------------
// Symbol stepper shared by every segment-parallel pass (speculative parse, merge, chain re-parse,
// block index) so that they all follow ONE trajectory function.  The passes run a single flat loop,
// one symbol per iteration for every lane; step() is branch-free (DC/AC and block-end handling are
// selects), so the lanes of a warp stay converged however their blocks are laid out.
//   DC symbol  input_DC  LIB/decoder/lossless_decode.c:210-224  (4-bit size + amplitude)
//   AC symbol  input_AC  :227-246 (4-bit run, 4-bit size, amplitude); size 0: run 15 = ZRL else END
//   block loop :101-133; `index` is uint8_t there and wraps, so it does here.
// A block is also ended when it reaches `flim` = min(end of the pass's job + RUNAWAY_BITS, end of stream):
// memory safety and termination on non-conforming input and on speculative garbage (a block made of ZRL symbols
// never ends by itself; conforming blocks are <= 1212 bits, SURVEY.md A.6, and never get near the limit).  The
// limit is a constant of the job, not of the block: keeping it per block cost two ALU-pipe instructions per symbol
// in passes that are bound by that pipe.
//
// Bit window (replaces update_buffer / INPUT_BITS, lossless_decode.c:139-162,207; only the number of
// consumed bits is observable, so the mechanics are free): two consecutive big-endian 32-bit words
// (w0 = current, w1 = next).  Positions are counted from the ALIGNED word that holds the stream's first
// byte ("f" positions = stream bit position + bias, bias = 8 * (address & 3)), so the offset into w0 is
// simply fpos & 31: the next 32 stream bits are ONE funnel shift (which takes its amount mod 32) and a
// word crossing is bit 5 of fpos flipping.  A third word w2 is in flight behind them: the funnel shift reads
// w1 in EVERY step, so a word loaded straight into w1 would be waited for one step later; loaded into w2 it
// is not touched until the next crossing (3-4 symbols), which hides an L2 round trip.  The payload buffer
// is padded: reads up to 16 bytes past any stream end are in bounds.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stream_bias(const uint8_t* base) {
    return (uint32_t)(reinterpret_cast<uintptr_t>(base) & 3u) * 8u;
}

struct Parser {
    const uint32_t* wp;    // next aligned word to fetch
    uint32_t w0, w1, w2;   // current word, look-ahead word, word in flight; MSB first
    uint32_t fpos;         // f position of the next symbol
    uint32_t flim;         // a block is ended at or after this f position (see above)
    uint32_t idx;          // zig-zag index of the next AC coefficient << 24 (the add wraps at 8 bits like the reference's
                           // uint8_t index, and "coefficient 63" is one unsigned compare)
    uint32_t nh;           // minus the header length of the next symbol: -4 = DC (a block start), -8 = AC
    uint32_t rmask;        // 32 while live; 0 = PARKED: the window is never refilled again (see park())

    // fbits = f position to start at (a block start), job_end = f position where the caller's job ends (it stops at
    // the first block start at or after it), ftotal = f position of the end of the stream.
    __device__ __forceinline__ void start(const uint8_t* base, uint32_t fbits, uint32_t job_end, uint32_t ftotal) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(base) & ~(uintptr_t)3) + (fbits >> 5);
        w0 = __byte_perm(__ldg(w), 0, 0x0123);
        w1 = __byte_perm(__ldg(w + 1), 0, 0x0123);
        w2 = __ldg(w + 2);                               // kept raw: byte-swapped when it moves into w1
        wp = w + 3;
        fpos = fbits;
        flim = min(job_end + RUNAWAY_BITS, ftotal);
        idx = 1u << 24;
        nh = (uint32_t)-4;
        rmask = 32u;
    }
    // The passes run UNIFORM loops: every lane of the warp steps every iteration, and a lane with nothing
    // (more) to parse is parked instead of branching around the step -- it keeps stepping through an all-zero
    // window (DC size 0 / END symbols: no coefficient, 4 or 8 bits each) without ever touching memory again,
    // and the pass ignores what it returns.
    __device__ __forceinline__ void park() { rmask = 0u; w0 = 0u; w1 = 0u; }
    __device__ __forceinline__ void init_parked() {
        wp = nullptr; w0 = w1 = w2 = 0u; fpos = 0u; flim = 0u; idx = 1u << 24; nh = (uint32_t)-4; rmask = 0u;
    }

    // What the last step() consumed.
    struct Sym {
        bool dc;           // it was the block's DC symbol
        bool coded;        // it was a non-zero AC coefficient ...
        uint32_t at;       // ... at this zig-zag index << 24 (may be >= 64 << 24 on non-conforming input: ignore then)
        int e;             // WANT_E only: amplitude of the DC / coded AC coefficient (HUFF_EXTEND, :204); 0 for a
                           // size-0 DC symbol, unspecified for END / ZRL
    };
    // Consume one symbol -- and the END symbol (eight zero bits) when one follows it directly: END carries no
    // information of its own, and folding it into the step of the symbol before it saves one of the ~4.7 steps of
    // an average block for four more instructions per step.  (The block boundaries, i.e. the trajectory, are the
    // same as with END as a step of its own, so passes with and without FOLD_END agree: the launchers turn it off
    // for dense streams, where a block has tens of symbols and the look-ahead costs more than the saved step.)
    // Returns true when the block ended (the parser is then positioned on the next block's DC symbol).
    template <bool WANT_E, bool FOLD_END = true>
    __device__ __forceinline__ bool step(Sym& sym) {
        const uint32_t t = __funnelshift_l(w1, w0, fpos);               // next 32 stream bits
        const bool dc = nh == (uint32_t)-4;
        const uint32_t rs = __funnelshift_r(t, 0u, nh);                 // t >> (32 - header bits): the 4 / 8 header bits
        const uint32_t size = rs & 15u, run = rs >> 4;                  // run == 0 for a DC symbol (rs < 16)
        const uint32_t len = size - nh;                                 // header + amplitude bits, <= 23
        if (WANT_E) {
            // amplitude: the `size` bits after the header, JPEG VLI sign extension (HUFF_EXTEND): a field whose
            // top bit is clear stands for field - 2^size + 1.  size 0 gives 0.
            const uint32_t v = t << (0u - nh);
            const uint32_t amp = (v >> 1) >> (31u ^ size);
            sym.e = (int)amp + (((int)v >= 0) ? (int)(0xFFFFFFFFu << size) + 1 : 0);
        } else {
            sym.e = 0;
        }
        const bool szd = size != 0u || dc;
        const bool coded = size != 0u && !dc;                           // a non-zero AC coefficient
        const uint32_t at = idx + ((szd ? run : 16u) << 24);            // DC: idx (1); ZRL: idx + 16; coefficient: its index
        const bool end0 = (!szd && run != 15u) || (coded && at >= (63u << 24));   // END / coefficient 63
        const bool end_next = FOLD_END && !end0 && ((t << len) >> 24) == 0u;   // ... or an END right behind this symbol
        const uint32_t fnew = fpos + len + (end_next ? 8u : 0u);        // <= 31 bits: at most one word crossing
        if ((fpos ^ fnew) & rmask) {                                    // crossed into w1: fetch the word after w2
            w0 = w1;
            w1 = __byte_perm(w2, 0, 0x0123);
            w2 = __ldg(wp);
            wp++;
        }
        const bool end = end0 || end_next || fnew >= flim;              // ... or the guard
        idx = end ? (1u << 24) : at + (coded ? (1u << 24) : 0u);
        nh = end ? (uint32_t)-4 : (uint32_t)-8;
        fpos = fnew;
        sym.dc = dc;
        sym.coded = coded;
        sym.at = at;
        return end;
    }
};

constexpr uint32_t FULL_MASK = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// 8-point LL&M inverse DCT pass: LIB/decoder/idct.c:46-97 (pass 1) == :123-168 (pass 2).
// Constants LIB/common/dct_math.h:53-64.  All arithmetic is int32 with wrap-around, exactly the
// reference's operation order (MULTIPLY is a plain 32-bit multiply, dct_math.h:76).
// SHIFT = 11 for pass 1 (CONST_BITS - PASS1_BITS), 18 for pass 2 (CONST_BITS + PASS1_BITS + 3).
// ------------------------------------------------------------------------------------------------
template <int SHIFT>
__device__ __forceinline__ void idct8(int i0, int i1, int i2, int i3, int i4, int i5, int i6, int i7,
                                      int (&o)[8]) {
    constexpr int RND = 1 << (SHIFT - 1);   // DESCALE, dct_math.h:48 -- added once to the even part (wrap-around
                                            // int32 sums are associative), not to each of the 8 outputs
    int z1 = (i2 + i6) * 4433;
    int tmp2 = z1 + i6 * -15137;
    int tmp3 = z1 + i2 * 6270;
    int tmp0 = (int)(((unsigned)(i0 + i4) << 13) + (unsigned)RND);
    int tmp1 = (int)(((unsigned)(i0 - i4) << 13) + (unsigned)RND);
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int a0 = i7, a1 = i5, a2 = i3, a3 = i1;
    int y1 = a0 + a3, y2 = a1 + a2, y3 = a0 + a2, y4 = a1 + a3;
    int y5 = (y3 + y4) * 9633;
    a0 *= 2446; a1 *= 16819; a2 *= 25172; a3 *= 12299;
    y1 *= -7373; y2 *= -20995;
    y3 = y3 * -16069 + y5;
    y4 = y4 * -3196 + y5;
    a0 += y1 + y3; a1 += y2 + y4; a2 += y2 + y3; a3 += y1 + y4;
    o[0] = (tmp10 + a3) >> SHIFT;
    o[7] = (tmp10 - a3) >> SHIFT;
    o[1] = (tmp11 + a2) >> SHIFT;
    o[6] = (tmp11 - a2) >> SHIFT;
    o[2] = (tmp12 + a1) >> SHIFT;
    o[5] = (tmp12 - a1) >> SHIFT;
    o[3] = (tmp13 + a0) >> SHIFT;
    o[4] = (tmp13 - a0) >> SHIFT;
}

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)w >> 16; }
// NORMALIZE, idct.c:20: max(min(v, 255), 0) -- one VIMNMX.RELU.
__device__ __forceinline__ uint32_t clamp255(int v) { return (uint32_t)__vimin_s32_relu(v, 255); }

// Four int32 -> four bytes, each clamped to 0..255 (NORMALIZE, idct.c:20), v0 in the low byte: two I2IP.U8.S32.SAT.
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(hi));
    return d;
}

// Column occupancy of a block held as 8 rows of packed int16: bit c of `ac` is set when column c has a
// non-zero coefficient in rows 1..7, bit c of `any` when it has one in any row.
__device__ __forceinline__ void block_masks(const uint4 (&rows)[8], uint32_t& ac, uint32_t& any) {
    uint32_t a[4] = {rows[1].x, rows[1].y, rows[1].z, rows[1].w};
#pragma unroll
    for (int r = 2; r < 8; r++) { a[0] |= rows[r].x; a[1] |= rows[r].y; a[2] |= rows[r].z; a[3] |= rows[r].w; }
    const uint32_t n[4] = {a[0] | rows[0].x, a[1] | rows[0].y, a[2] | rows[0].z, a[3] | rows[0].w};
    ac = any = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ac |= ((a[j] & 0xFFFFu) ? 1u : 0u) << (2 * j) | ((a[j] >> 16) ? 1u : 0u) << (2 * j + 1);
        any |= ((n[j] & 0xFFFFu) ? 1u : 0u) << (2 * j) | ((n[j] >> 16) ? 1u : 0u) << (2 * j + 1);
    }
}

// Full 8x8 IDCT of one block held as 8 rows of packed int16 (uint4 = one 16-byte row).
// Output: 16 words, word 2r / 2r+1 = pixels 0-3 / 4-7 of row r (little-endian bytes).
//
// acmask / anymask are WARP-UNIFORM supersets of block_masks() over the warp's 32 blocks, so every
// branch below is divergence-free.  They only select bit-exact shortcuts of the reference arithmetic:
//   * a column whose rows 1..7 are zero: every pass-1 output is DESCALE(in0 << 13, 11) == in0 << 2
//     (the low 11 bits of in0 << 13 are zero, so the rounding term never carries) -- idct.c:55-63,99-106;
//   * an all-zero column gives an all-zero workspace column;
//   * workspace columns 4..7 all zero: pass 2 with those inputs as literal zeros (same expression tree);
//   * only column 0 non-zero and no AC in it (DC-only blocks): pass 2 is DESCALE(ws << 13, 18) ==
//     (ws + 16) >> 5 with ws = dc << 2 -- idct.c:123-180.
// Pass 0xFF / 0xFF to disable every shortcut.
__device__ __forceinline__ void idct_block(const uint4 (&rows)[8], uint32_t acmask, uint32_t anymask,
                                           uint32_t (&out)[16]) {
    if (((anymask & 0xFEu) | (acmask & 1u)) == 0) {
        const uint32_t v = clamp255(((lo16(rows[0].x) << 2) + 16) >> 5) * 0x01010101u;
#pragma unroll
        for (int k = 0; k < 16; k++) out[k] = v;
        return;
    }
    int ws[8][8];
    // Pass 1: columns (idct.c:41-109).  Column c takes element c of every row.
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int in[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint32_t w = (c >> 1) == 0 ? rows[r].x : (c >> 1) == 1 ? rows[r].y : (c >> 1) == 2 ? rows[r].z : rows[r].w;
            in[r] = (c & 1) ? hi16(w) : lo16(w);
        }
        if (acmask & (1u << c)) {
            int o[8];
            idct8<11>(in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], o);
#pragma unroll
            for (int r = 0; r < 8; r++) ws[r][c] = o[r];
        } else {
            const int v = (anymask & (1u << c)) ? (int)((unsigned)in[0] << 2) : 0;
#pragma unroll
            for (int r = 0; r < 8; r++) ws[r][c] = v;
        }
    }
    // Pass 2: rows (idct.c:116-180), clamp to 0..255, no level shift.
    const bool low_half_only = (anymask & 0xF0u) == 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int o[8];
        if (low_half_only) idct8<18>(ws[r][0], ws[r][1], ws[r][2], ws[r][3], 0, 0, 0, 0, o);
        else idct8<18>(ws[r][0], ws[r][1], ws[r][2], ws[r][3], ws[r][4], ws[r][5], ws[r][6], ws[r][7], o);
        out[2 * r] = pack4_sat_u8(o[0], o[1], o[2], o[3]);
        out[2 * r + 1] = pack4_sat_u8(o[4], o[5], o[6], o[7]);
    }
}

__device__ __forceinline__ uint32_t warp_or(uint32_t v) { return __reduce_or_sync(0xFFFFFFFFu, v); }

// 256-bit global store (sm_100+): one full 32-byte sector per lane.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// 256-bit global load (coherent: the data may have been written by this thread earlier in the kernel).
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p)
                 : "memory");
}

// ---- packed colour conversion: LIB/decoder/ycbcr_to_rgb.c:26-49 ---------------------------------------------
// R = (Y << 14) + 22970 (Cr - 128), G = (Y << 14) - 5638 (Cb - 128) - 11700 (Cr - 128), B = (Y << 14) + 29032 (Cb - 128)
// (:33-37), each through NORMALIZE_RGB (:19): negative -> 0, else >> 14, capped at 255.  Pixel word = B | G << 8 |
// R << 16 | A(0) << 24 (rgb_pixel_t, mjpeg423_types.h:56-61).
// Because Y << 14 has no low bits, NORMALIZE_RGB((Y << 14) + k) == clamp(Y + (k >> 14), 0, 255) exactly
// (arithmetic shift = floor; k = the chroma terms of ycbcr_to_rgb.c:33-37, |k >> 14| < 256).  Two pixels are
// processed per register in signed 16-bit lanes: one VIADDMNMX.S16x2.RELU does the add and both clamps of a
// channel for two pixels, and PRMT assembles the BGRA words (alpha = the zero high byte of a lane).
// Chroma terms are computed x4 in 32 bits, so that floor(k / 2^14) is the HIGH half of the product and one
// PRMT packs two of them.
__device__ __forceinline__ uint32_t addclamp2(uint32_t y2, uint32_t k2) { return __viaddmin_s16x2_relu(y2, k2, 0x00FF00FFu); }
__device__ __forceinline__ void bgra2(uint32_t b, uint32_t g, uint32_t r, uint32_t& p0, uint32_t& p1) {
    const uint32_t bg = __byte_perm(b, g, 0x6240);       // B0 G0 B1 G1
    p0 = __byte_perm(bg, r, 0x5410);                     // B0 G0 R0 0
    p1 = __byte_perm(bg, r, 0x7632);                     // B1 G1 R1 0
}
__device__ __forceinline__ int kr4(uint32_t cr) { return 4 * 22970 * (int)cr - 4 * 128 * 22970; }
__device__ __forceinline__ int kb4(uint32_t cb) { return 4 * 29032 * (int)cb - 4 * 128 * 29032; }
__device__ __forceinline__ int kg4(uint32_t cb, uint32_t cr) {
    return -4 * 5638 * (int)cb - 4 * 11700 * (int)cr + 4 * 128 * (5638 + 11700);
}
__device__ __forceinline__ uint32_t hi2(int a0, int a1) { return __byte_perm((uint32_t)a0, (uint32_t)a1, 0x7632); }   // a0 >> 16 | a1 >> 16 << 16

// Pixels that share one (Cb, Cr) pair -- a block whose chroma blocks are flat: the chroma terms are per-block constants.
struct FlatChroma {
    uint32_t kr2, kg2, kb2;
    __device__ __forceinline__ FlatChroma(uint32_t cb, uint32_t cr) {
        const int r = kr4(cr), g = kg4(cb, cr), b = kb4(cb);
        kr2 = hi2(r, r); kg2 = hi2(g, g); kb2 = hi2(b, b);
    }
    // y2 = two Y samples in 16-bit lanes -> two BGRA words
    __device__ __forceinline__ void px2(uint32_t y2, uint32_t& p0, uint32_t& p1) const {
        bgra2(addclamp2(y2, kb2), addclamp2(y2, kg2), addclamp2(y2, kr2), p0, p1);
    }
    __device__ __forceinline__ void row_store(uint32_t y0, uint32_t y1, uint8_t* dst) const {
        uint32_t v[8];
        px2(__byte_perm(y0, 0, 0x4140), v[0], v[1]);
        px2(__byte_perm(y0, 0, 0x4342), v[2], v[3]);
        px2(__byte_perm(y1, 0, 0x4140), v[4], v[5]);
        px2(__byte_perm(y1, 0, 0x4342), v[6], v[7]);
        st_global_v8(dst, v);
    }
};

// Four pixels (one packed word of each plane) -> four BGRA words.
__device__ __forceinline__ void colour4(uint32_t y, uint32_t cb, uint32_t cr, uint32_t* v) {
    uint32_t b[4], r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { b[k] = __byte_perm(cb, 0, 0x4440 + k); r[k] = __byte_perm(cr, 0, 0x4440 + k); }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t y2 = __byte_perm(y, 0, h ? 0x4342 : 0x4140);
        const uint32_t kr2 = hi2(kr4(r[2 * h]), kr4(r[2 * h + 1]));
        const uint32_t kg2 = hi2(kg4(b[2 * h], r[2 * h]), kg4(b[2 * h + 1], r[2 * h + 1]));
        const uint32_t kb2 = hi2(kb4(b[2 * h]), kb4(b[2 * h + 1]));
        bgra2(addclamp2(y2, kb2), addclamp2(y2, kg2), addclamp2(y2, kr2), v[2 * h], v[2 * h + 1]);
    }
}

// One block row (8 pixels) of Y/Cb/Cr packed samples -> 8 BGRA words, stored as one 32-byte sector.
__device__ __forceinline__ void colour_row_store(uint32_t y0, uint32_t y1, uint32_t cb0, uint32_t cb1, uint32_t cr0,
                                                 uint32_t cr1, uint8_t* dst) {
    uint32_t v[8];
    colour4(y0, cb0, cr0, v);
    colour4(y1, cb1, cr1, v + 4);
    st_global_v8(dst, v);
}

}  // namespace mj
