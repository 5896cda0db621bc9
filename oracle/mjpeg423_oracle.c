/*
 * mjpeg423_oracle.c -- CPU restatement of the MJPEG423 decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mjpeg423-video-decoder-software_b200/ may link, import or call
 * this file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * have an independent checker for the CUDA path.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so this
 * restatement is pinned two ways: (1) against the known-answer vectors in BASELINE.md
 * section 5 / tests/golden/kat.json, which were produced by running the unmodified
 * reference functions; (2) against oracle/_ref/libmjpeg423_ref.so, i.e. the reference's
 * own C files compiled in place by oracle/Makefile, on randomised inputs
 * (tests/test_oracle.py).
 *
 * LIB/ below abbreviates /root/reference/core0/software/common/libs/mjpeg423/.
 * All arithmetic is done on uint32_t and reinterpreted, so wrap-around is defined
 * behaviour here even where the reference relies on -fwrapv.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- tables: LIB/common/tables.c:13-42 (JPEG Annex K, natural order; zig-zag scan) ---- */
const int16_t orc_Yquant[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,   12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,   14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,   24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,  72, 92, 95, 98, 112, 100, 103, 99};
const int16_t orc_Cquant[64] = {
    17, 18, 24, 47, 99, 99, 99, 99,  18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,  47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,  99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,  99, 99, 99, 99, 99, 99, 99, 99};

/* zig-zag position -> natural (row*8+col) index, generated rather than typed in. */
static uint8_t zz_nat[64];
static pthread_once_t zz_once = PTHREAD_ONCE_INIT;
static void zz_build(void) {
    int r = 0, c = 0;
    for (int k = 0; k < 64; k++) {
        zz_nat[k] = (uint8_t)(r * 8 + c);
        if ((r + c) & 1) {            /* moving down-left */
            if (r == 7) c++; else if (c == 0) r++; else { r++; c--; }
        } else {                      /* moving up-right */
            if (c == 7) r++; else if (r == 0) c++; else { r--; c++; }
        }
    }
}
const uint8_t* orc_zigzag(void) { pthread_once(&zz_once, zz_build); return zz_nat; }

/* ---- entropy decode: LIB/decoder/lossless_decode.c:60-135 ---- */
/* MSB-first reader.  The reference keeps a 32-bit window that it tops up by whole bytes
 * (update_buffer, :139-162); only the consumed-bit count matters for the result, so this
 * reader simply indexes bits.  `peek` returns the next n (<=24) bits. */
typedef struct { const uint8_t* p; uint64_t bit; } bitrd_t;
static inline uint32_t peek(const bitrd_t* r, int n) {
    const uint8_t* q = r->p + (r->bit >> 3);
    uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    w <<= (r->bit & 7);           /* <=7 bits dropped, 25 left: enough for header+15 via two peeks */
    return n ? (w >> (32 - n)) : 0;
}
/* JPEG VLI sign extension, HUFF_EXTEND at :204, result truncated to DCTELEM as in huff_input_t. */
static inline int16_t vli(uint32_t x, int s) {
    if (x < (1u << (s - 1))) x += (0xFFFFFFFFu << s) + 1u;
    return (int16_t)x;
}
/* Returns the number of bits consumed (test hook; the reference returns void). */
uint64_t orc_lossless_decode(int num_blocks, const void* bitstream, int16_t* DCACq,
                             const int16_t* quant, int P) {
    const uint8_t* zz = orc_zigzag();
    bitrd_t rd = {(const uint8_t*)bitstream, 0};
    int16_t cur = 0;                                   /* :73 DC predictor, DCTELEM wide */
    if (!P) memset(DCACq, 0, (size_t)num_blocks * 128); /* :77-78 */
    for (int b = 0; b < num_blocks; b++) {
        int16_t* pe = DCACq + (size_t)b * 64;
        /* DC symbol: input_DC :210-224 */
        int s = (int)peek(&rd, 4); rd.bit += 4;
        int16_t e = 0;
        if (s) { e = vli(peek(&rd, s), s); rd.bit += s; }
        if (P) pe[0] = (int16_t)(pe[0] + e * quant[0]);            /* :90-92 */
        else { cur = (int16_t)(cur + e); pe[0] = (int16_t)(cur * quant[0]); } /* :93-96 */
        /* AC symbols: input_AC :227-246, loop :101-133.  index is uint8_t in the reference. */
        uint8_t idx = 1;
        for (;;) {
            int run = (int)peek(&rd, 4);
            rd.bit += 4;
            s = (int)peek(&rd, 4);
            rd.bit += 4;
            if (s == 0) {                       /* e == 0 <=> size == 0 */
                if (run == 15) { idx = (uint8_t)(idx + 16); continue; } /* ZRL :107-110 */
                break;                          /* END :111-114 (any run != 15) */
            }
            e = vli(peek(&rd, s), s); rd.bit += s;
            idx = (uint8_t)(idx + run);
            if (idx < 64) {                     /* reference indexes unchecked; conforming streams stay < 64 */
                int n = zz[idx];
                if (P) pe[n] = (int16_t)(pe[n] + e * quant[n]);     /* :121-123 */
                else   pe[n] = (int16_t)(e * quant[n]);             /* :124-126 */
            }
            if (idx >= 63) break;               /* :130 */
            idx++;
        }
    }
    return rd.bit;
}

/* ---- 8x8 IDCT: LIB/decoder/idct.c:22-181, constants LIB/common/dct_math.h:48-76 ---- */
#define C0_298 2446u
#define C0_390 3196u
#define C0_541 4433u
#define C0_765 6270u
#define C0_899 7373u
#define C1_175 9633u
#define C1_501 12299u
#define C1_847 15137u
#define C1_961 16069u
#define C2_053 16819u
#define C2_562 20995u
#define C3_072 25172u
static inline int32_t asr_round(uint32_t x, int n) {   /* DESCALE, dct_math.h:48 */
    return (int32_t)(x + (1u << (n - 1))) >> n;
}
/* One LL&M 8-point inverse pass on in[0..7] (idct.c:46-97 == :123-168); out[k] are the
 * un-descaled sums in output order 0..7. */
static inline void llm8(const int32_t in[8], uint32_t out[8]) {
    uint32_t a2 = (uint32_t)in[2], a6 = (uint32_t)in[6];
    uint32_t z1 = (a2 + a6) * C0_541;
    uint32_t t2 = z1 - a6 * C1_847;
    uint32_t t3 = z1 + a2 * C0_765;
    uint32_t t0 = ((uint32_t)in[0] + (uint32_t)in[4]) << 13;
    uint32_t t1 = ((uint32_t)in[0] - (uint32_t)in[4]) << 13;
    uint32_t e0 = t0 + t3, e3 = t0 - t3, e1 = t1 + t2, e2 = t1 - t2;
    uint32_t o0 = (uint32_t)in[7], o1 = (uint32_t)in[5], o2 = (uint32_t)in[3], o3 = (uint32_t)in[1];
    uint32_t y1 = o0 + o3, y2 = o1 + o2, y3 = o0 + o2, y4 = o1 + o3;
    uint32_t y5 = (y3 + y4) * C1_175;
    o0 *= C0_298; o1 *= C2_053; o2 *= C3_072; o3 *= C1_501;
    y1 = 0u - y1 * C0_899; y2 = 0u - y2 * C2_562;
    y3 = y5 - y3 * C1_961; y4 = y5 - y4 * C0_390;
    o0 += y1 + y3; o1 += y2 + y4; o2 += y2 + y3; o3 += y1 + y4;
    out[0] = e0 + o3; out[7] = e0 - o3;
    out[1] = e1 + o2; out[6] = e1 - o2;
    out[2] = e2 + o1; out[5] = e2 - o1;
    out[3] = e3 + o0; out[4] = e3 - o0;
}
void orc_idct(const int16_t* DCAC, uint8_t* block) {
    int32_t ws[64], v[8];
    uint32_t o[8];
    for (int c = 0; c < 8; c++) {              /* pass 1, columns: descale by CONST_BITS-PASS1_BITS = 11 */
        for (int r = 0; r < 8; r++) v[r] = DCAC[r * 8 + c];
        llm8(v, o);
        for (int r = 0; r < 8; r++) ws[r * 8 + c] = asr_round(o[r], 11);
    }
    for (int r = 0; r < 8; r++) {              /* pass 2, rows: descale by 13+2+3 = 18, clamp (:20) */
        llm8(ws + r * 8, o);
        for (int c = 0; c < 8; c++) {
            int32_t t = asr_round(o[c], 18);
            block[r * 8 + c] = (uint8_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
        }
    }
}

/* ---- colour: LIB/decoder/ycbcr_to_rgb.c:26-49; pixel layout LIB/common/mjpeg423_types.h:56-61 ---- */
static inline uint32_t sat14(int32_t t) {          /* NORMALIZE_RGB :19 */
    if (t < 0) return 0;
    t >>= 14;
    return t > 255 ? 255u : (uint32_t)t;
}
void orc_ycbcr_to_rgb(int h, int w, uint32_t w_size, const uint8_t* Y, const uint8_t* Cb,
                      const uint8_t* Cr, uint8_t* rgb /* BGRA bytes */) {
    for (int y = 0; y < 8; y++) {
        uint8_t* dst = rgb + ((size_t)(h + y) * w_size + (size_t)w) * 4;
        for (int x = 0; x < 8; x++) {
            int32_t cb = (int32_t)Cb[y * 8 + x] - 128, cr = (int32_t)Cr[y * 8 + x] - 128;
            int32_t yy = (int32_t)Y[y * 8 + x] << 14;
            dst[4 * x + 0] = (uint8_t)sat14(yy + 29032 * cb);
            dst[4 * x + 1] = (uint8_t)sat14(yy - 5638 * cb - 11700 * cr);
            dst[4 * x + 2] = (uint8_t)sat14(yy + 22970 * cr);
            dst[4 * x + 3] = 0;
        }
    }
}

/* ==== encoder restatement ==================================================
 * LIB/encoder/{rgb_to_ycbcr,fdct,quantize,lossless_encode}.c.  Pinned against the reference's own
 * functions in oracle/_ref (tests/test_oracle.py::test_encoder_*).  Floating point: the reference
 * evaluates the colour equations in double, left to right, and truncates to uint8_t; this file must be
 * compiled without FMA contraction (oracle/Makefile passes -ffp-contract=off). */

/* LIB/encoder/rgb_to_ycbcr.c:58-70.  rgb = BGRA raster (rgb_pixel_t, mjpeg423_types.h:56-61). */
void orc_rgb_to_ycbcr(int h, int w, uint32_t w_size, const uint8_t* rgb, uint8_t* Y, uint8_t* Cb, uint8_t* Cr) {
    for (int y = 0; y < 8; y++) {
        const uint8_t* px = rgb + ((size_t)(y + h) * w_size + (size_t)w) * 4;
        for (int x = 0; x < 8; x++, px += 4) {
            const int B = px[0], G = px[1], R = px[2];
            const double yy = 0.299 * R + 0.587 * G + 0.114 * B;
            const double cb = -0.168736 * R - 0.331264 * G + 0.5 * B + 128;
            const double cr = 0.5 * R - 0.418688 * G - 0.081312 * B + 128;
            Y[y * 8 + x] = (uint8_t)yy;
            Cb[y * 8 + x] = (uint8_t)cb;
            Cr[y * 8 + x] = (uint8_t)cr;
        }
    }
}

/* 8-point forward LL&M stage shared by both passes of LIB/encoder/fdct.c:33-161.  `in` are the 8 inputs;
 * out[0]/out[4] receive the two plain sums (scaled by the caller), out[1..3,5..7] the rotated terms
 * BEFORE descaling. */
static inline void fllm8(const int32_t in[8], int32_t sum[2], int32_t rot[8]) {
    const int32_t t0 = in[0] + in[7], t7 = in[0] - in[7], t1 = in[1] + in[6], t6 = in[1] - in[6];
    const int32_t t2 = in[2] + in[5], t5 = in[2] - in[5], t3 = in[3] + in[4], t4 = in[3] - in[4];
    const int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    sum[0] = t10 + t11;
    sum[1] = t10 - t11;
    int32_t z1 = (t12 + t13) * 4433;
    rot[2] = z1 + t13 * 6270;
    rot[6] = z1 + t12 * -15137;
    z1 = t4 + t7;
    int32_t z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int32_t z5 = (z3 + z4) * 9633;
    const int32_t a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    rot[7] = a4 + z1 + z3;
    rot[5] = a5 + z2 + z4;
    rot[3] = a6 + z2 + z3;
    rot[1] = a7 + z1 + z4;
}
static inline int32_t fdescale(int32_t x, int n) { return (int32_t)(((uint32_t)x + (1u << (n - 1)))) >> n; }

/* LIB/encoder/fdct.c:17-161: rows (results scaled up by 2^PASS1_BITS, stored as int16), then columns. */
void orc_fdct(const uint8_t* block, int16_t* DCAC) {
    for (int r = 0; r < 8; r++) {
        int32_t in[8], sum[2], rot[8];
        for (int k = 0; k < 8; k++) in[k] = block[r * 8 + k];
        fllm8(in, sum, rot);
        DCAC[r * 8 + 0] = (int16_t)(sum[0] << 2);
        DCAC[r * 8 + 4] = (int16_t)(sum[1] << 2);
        for (int k = 1; k < 8; k++) if (k != 4) DCAC[r * 8 + k] = (int16_t)fdescale(rot[k], 11);
    }
    for (int c = 0; c < 8; c++) {
        int32_t in[8], sum[2], rot[8];
        for (int k = 0; k < 8; k++) in[k] = DCAC[k * 8 + c];
        fllm8(in, sum, rot);
        DCAC[0 * 8 + c] = (int16_t)fdescale(sum[0], 5);
        DCAC[4 * 8 + c] = (int16_t)fdescale(sum[1], 5);
        for (int k = 1; k < 8; k++) if (k != 4) DCAC[k * 8 + c] = (int16_t)fdescale(rot[k], 18);
    }
}

/* DOUBLE_QUANTIZE, LIB/encoder/quantize.c:16: (DCTELEM) round((double)x / (double)q), C round() =
 * half away from zero.  Restated in integers (exact: a quotient that is not a half-integer is at least
 * 1/(2q) away from one, far more than a double rounding error). */
static inline int16_t quant1(int16_t x, int16_t q) {
    const int32_t a = x < 0 ? -(int32_t)x : (int32_t)x, d = q < 0 ? -(int32_t)q : (int32_t)q;
    const int32_t m = (2 * a + d) / (2 * d);
    return (int16_t)(((x < 0) != (q < 0)) ? -m : m);
}
/* quantize.c:18-31.  prev = DC level of the previous block of the plane. */
void orc_quantize_I(int16_t* prev, const int16_t* quant, const int16_t* DCAC, int16_t* DCACq, int16_t* DCACq_next) {
    const int16_t dc = quant1(DCAC[0], quant[0]);
    DCACq[0] = (int16_t)(dc - *prev);
    *prev = dc;
    DCACq_next[0] = dc;
    for (int k = 1; k < 64; k++) DCACq_next[k] = DCACq[k] = quant1(DCAC[k], quant[k]);
}
/* quantize.c:33-42. */
void orc_quantize_P(const int16_t* quant, int16_t* DCACq_prev, const int16_t* DCAC, int16_t* DCACq) {
    for (int k = 0; k < 64; k++) {
        const int16_t v = quant1(DCAC[k], quant[k]);
        DCACq[k] = (int16_t)(v - DCACq_prev[k]);
        DCACq_prev[k] = v;
    }
}

/* LIB/encoder/lossless_encode.c:30-138.  MSB-first bit writer; returns the stream length in bytes.
 * fix_tail == 0 reproduces output_rest() (:80-83), which stores the LOW byte of the bit buffer (always
 * zero) instead of the pending top byte: the last partial byte of the stream is 0. */
typedef struct { uint8_t* p; uint64_t bits; uint32_t acc; int nacc; } bitwr_t;
static inline void put_bits(bitwr_t* w, int n, uint32_t v) {
    for (int i = n - 1; i >= 0; i--) {
        w->acc = (w->acc << 1) | ((v >> i) & 1u);
        if (++w->nacc == 8) { w->p[w->bits >> 3] = (uint8_t)w->acc; w->acc = 0; w->nacc = 0; }
        w->bits++;
    }
}
static inline uint32_t vli_enc(int32_t x, uint32_t* size) {          /* encode_VLI :121-138 */
    uint32_t a = (uint32_t)(x < 0 ? -x : x), s = 0;
    while (s < 11 && a >> s) s++;                                    /* sizes are capped at 11 (:135) */
    *size = s;
    return x > 0 ? (uint32_t)x : (((uint32_t)(x - 1)) & (s ? (0xFFFFFFFFu >> (32 - s)) : 0u));
}
uint32_t orc_lossless_encode(int num_blocks, const int16_t* DCACq, uint8_t* bitstream, int fix_tail) {
    const uint8_t* zz = orc_zigzag();
    bitwr_t w = {bitstream, 0, 0, 0};
    for (int b = 0; b < num_blocks; b++) {
        const int16_t* dct = DCACq + (size_t)b * 64;
        uint32_t size, amp = vli_enc(dct[0], &size);
        put_bits(&w, 4, size);                                        /* output_DC :86-96 */
        put_bits(&w, (int)size, amp);
        int last = 63;
        while (last > 0 && dct[zz[last]] == 0) last--;
        int idx = 1;
        while (idx <= last) {
            int run = 0;
            while (run < 16 && dct[zz[idx]] == 0) { run++; idx++; }
            if (run == 16) { put_bits(&w, 8, 0xF0); }                 /* output_ZRL */
            else {
                amp = vli_enc(dct[zz[idx]], &size);
                put_bits(&w, 4, (uint32_t)run);
                put_bits(&w, 4, size);
                put_bits(&w, (int)size, amp);
                idx++;
            }
        }
        if (last < 63) put_bits(&w, 8, 0);                            /* output_END */
    }
    if (w.nacc) w.p[w.bits >> 3] = fix_tail ? (uint8_t)(w.acc << (8 - w.nacc)) : 0;
    return (uint32_t)((w.bits + 7) >> 3);
}
