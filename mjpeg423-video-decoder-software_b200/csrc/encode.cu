// encode.cu -- batched MJPEG423 encoder (sm_100a), SURVEY.md section 8 row f3: the step on the other side of
// the wire format.  Restates the frame loop of mjpeg423_encode(), LIB/encoder/mjpeg423_encoder.c:97-225
// (LIB = /root/reference/core0/software/common/libs/mjpeg423), for MANY frames at once; output is
// byte-identical to the reference encoder's file (up to its 512 uninitialised trailing bytes).
//
//   k_enc_transform  colour (LIB/encoder/rgb_to_ycbcr.c:58-70, in double like the reference: the products are
//                    not exactly representable, so the truncation to uint8 depends on IEEE rounding) ->
//                    8x8 forward LL&M DCT (LIB/encoder/fdct.c:17-161, rows then columns, int16 stores) ->
//                    quantise (LIB/encoder/quantize.c:16: round(x / q), half away from zero) to ABSOLUTE
//                    levels.  8 lanes per block position: lane = pixel row, then = coefficient column.
//   k_enc_size       one THREAD per block: code length in bits of the block as I-frame block (DC differential
//                    against the previous block of the plane, quantize.c:18-25) and as P-frame block (every
//                    level differential against the previous frame, :33-42): 64-bit occupancy mask in zig-zag
//                    order + sum of VLI sizes, the ZRLs from the gaps between set bits.
//   k_enc_scan       exclusive scan of the block lengths of every (frame, plane, variant): bit offset of every
//                    block inside its plane stream + the stream length.
//   (host)           the I/P decision of mjpeg423_encoder.c:155-185 needs only the six stream lengths of
//                    every frame: I when first frame, I not larger than P, or max_I_interval reached.
//   k_enc_emit       one thread per block of the chosen variant: walks the set bits of the mask and ORs the symbols
//                    (LIB/encoder/lossless_encode.c:30-138: 4-bit DC size / 4-bit run + 4-bit size, JPEG VLI
//                    amplitude, ZRL = F0, END = 00 unless position 63 is coded) into the zero-initialised output
//                    (MSB first).
// The quirk of output_rest() (lossless_encode.c:80-83: the last partial byte of every plane stream is written
// as 0) is reproduced unless MJPEG423_ENC_FIX_TAIL is set.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "runtime.h"

namespace mj {

__constant__ uint8_t c_enc_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                         12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                         58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ---- colour + FDCT + quantise ---------------------------------------------------------------------------------
// 8-point forward LL&M stage (fdct.c:33-96 == :104-161): sum[0..1] = the two plain sums (outputs 0 and 4 before
// scaling), rot[k] = outputs 1,2,3,5,6,7 before DESCALE.  Constants LIB/common/dct_math.h:53-64.
__device__ __forceinline__ void fllm8(const int (&in)[8], int (&sum)[2], int (&rot)[8]) {
    const int t0 = in[0] + in[7], t7 = in[0] - in[7], t1 = in[1] + in[6], t6 = in[1] - in[6];
    const int t2 = in[2] + in[5], t5 = in[2] - in[5], t3 = in[3] + in[4], t4 = in[3] - in[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    sum[0] = t10 + t11;
    sum[1] = t10 - t11;
    int z1 = (t12 + t13) * 4433;
    rot[2] = z1 + t13 * 6270;
    rot[6] = z1 + t12 * -15137;
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    rot[7] = a4 + z1 + z3;
    rot[5] = a5 + z2 + z4;
    rot[3] = a6 + z2 + z3;
    rot[1] = a7 + z1 + z4;
    rot[0] = rot[4] = 0;
}
__device__ __forceinline__ int fdescale(int x, int n) { return (int)((unsigned)x + (1u << (n - 1))) >> n; }   // DESCALE
// (DCTELEM) round((double)x / (double)q), quantize.c:16, in integers: exact, because a quotient that is not a
// half-integer is at least 1/(2q) away from one -- far more than a double rounding error.
__device__ __forceinline__ int quant1(int x, int q) {
    const int a = abs(x), d = abs(q);
    const int m = (2 * a + d) / (2 * d);
    return (int)(int16_t)(((x < 0) != (q < 0)) ? -m : m);
}
// rgb_to_ycbcr.c:64-66: double arithmetic, evaluated left to right, no contraction, truncated to uint8.
__device__ __forceinline__ void rgb2ycc(int R, int G, int B, uint32_t& y, uint32_t& cb, uint32_t& cr) {
    const double r = (double)R, g = (double)G, b = (double)B;
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.299, r), __dmul_rn(0.587, g)), __dmul_rn(0.114, b));
    const double cbb = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(-0.168736, r), __dmul_rn(0.331264, g)), __dmul_rn(0.5, b)), 128.0);
    const double crr = __dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(0.5, r), __dmul_rn(0.418688, g)), __dmul_rn(0.081312, b)), 128.0);
    y = (uint32_t)__double2int_rz(yy) & 255u;
    cb = (uint32_t)__double2int_rz(cbb) & 255u;
    cr = (uint32_t)__double2int_rz(crr) & 255u;
}

constexpr int ENC_TPB = 256;
// levels: [frame][plane][block][64] int16, natural order.  Warp = 4 block positions x 8 lanes.
__global__ void __launch_bounds__(ENC_TPB)
k_enc_transform(const uint8_t* __restrict__ frames, const int16_t* __restrict__ quant, int16_t* __restrict__ levels,
                uint32_t nb, uint32_t wb, uint32_t W, uint32_t H, uint32_t n_frames) {
    __shared__ __align__(16) int16_t s_tile[ENC_TPB / 8][8][8 + 8];   // one padded 8x8 tile per block position
    const int t = threadIdx.x, r = t & 7, grp = t >> 3;
    const uint64_t gp = (uint64_t)blockIdx.x * (ENC_TPB / 8) + grp;   // global block position
    const uint32_t f = (uint32_t)(gp / nb), b = (uint32_t)(gp % nb);
    const bool valid = f < n_frames;
    uint32_t ys[8], cbs[8], crs[8];
    if (valid) {
        const uint8_t* src = frames + (((size_t)f * H + (size_t)(b / wb) * 8 + r) * W + (size_t)(b % wb) * 8) * 4;
        const uint4 p0 = __ldg(reinterpret_cast<const uint4*>(src)), p1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
        const uint32_t px[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) rgb2ycc((px[k] >> 16) & 255, (px[k] >> 8) & 255, px[k] & 255, ys[k], cbs[k], crs[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) ys[k] = cbs[k] = crs[k] = 0;
    }
    int16_t (*tile)[16] = s_tile[grp];
#pragma unroll 1
    for (int p = 0; p < 3; p++) {
        int in[8], sum[2], rot[8];
#pragma unroll
        for (int k = 0; k < 8; k++) in[k] = (int)(p == 0 ? ys[k] : p == 1 ? cbs[k] : crs[k]);
        fllm8(in, sum, rot);                                             // pass 1: row r (fdct.c:33-96)
        tile[r][0] = (int16_t)(sum[0] << 2);
        tile[r][4] = (int16_t)(sum[1] << 2);
#pragma unroll
        for (int k = 1; k < 8; k++) if (k != 4) tile[r][k] = (int16_t)fdescale(rot[k], 11);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) in[k] = tile[k][r];                   // pass 2: column r (fdct.c:104-161)
        __syncwarp();
        fllm8(in, sum, rot);
        const int16_t* q = quant + (p ? 64 : 0);
        int out[8];
        out[0] = (int16_t)fdescale(sum[0], 5);
        out[4] = (int16_t)fdescale(sum[1], 5);
#pragma unroll
        for (int k = 1; k < 8; k++) if (k != 4) out[k] = (int16_t)fdescale(rot[k], 18);
#pragma unroll
        for (int k = 0; k < 8; k++) tile[k][r] = (int16_t)quant1(out[k], q[k * 8 + r]);
        __syncwarp();
        if (valid)
            *reinterpret_cast<uint4*>(levels + (((size_t)f * 3 + p) * nb + b) * 64 + r * 8) = *reinterpret_cast<const uint4*>(&tile[r][0]);
        __syncwarp();
    }
}

// ---- entropy coder: symbols of one block, two zig-zag positions per lane ----------------------------------------
// encode_VLI, lossless_encode.c:121-138: size = bit length of |x| capped at 11, amplitude = x, or the low `size`
// bits of x - 1 for negative x.  (The amplitude is masked to `size` bits; the reference relies on |x| < 2048.)
__device__ __forceinline__ void vli(int x, uint32_t& size, uint32_t& amp) {
    const uint32_t a = (uint32_t)abs(x);
    size = min(32u - (uint32_t)__clz(a), 11u);
    amp = (uint32_t)(x > 0 ? x : x - 1) & ((1u << size) - 1u);
}
// Inverse zig-zag: natural index -> zig-zag position (the unrolled loops below call it with constants, so it folds).
__host__ __device__ constexpr int inv_zigzag(int n) {
    constexpr uint8_t inv[64] = {0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
    return inv[n];
}

// Occupancy of one block: `mask` bit k = the value coded at zig-zag position k (1..63) is non-zero (bit 0 is always set:
// the DC symbol anchors the first run), `sizes` = sum of the VLI sizes of those values, `dc` = the value of the DC symbol.
struct BlockOcc { uint64_t mask; uint32_t sizes; int dc; };
// variant 0 = I frame (levels as they are, DC differential against the previous block), 1 = P frame (every level
// differential against the previous frame, int16 wrap like the reference's DCTELEM stores).
template <int VARIANT>
__device__ __forceinline__ BlockOcc block_occ(const int16_t* __restrict__ cur, const int16_t* __restrict__ prev, bool first_block) {
    uint32_t lo = 1u, hi = 0u, sizes = 0;
    int dc = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(cur) + r);
        uint4 q = make_uint4(0, 0, 0, 0);
        if (VARIANT == 1) q = __ldg(reinterpret_cast<const uint4*>(prev) + r);
        const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int v = (k & 1) ? hi16(cw[k >> 1]) : lo16(cw[k >> 1]);
            if (VARIANT == 1) v = (int)(int16_t)(v - ((k & 1) ? hi16(qw[k >> 1]) : lo16(qw[k >> 1])));       // quantize.c:38-39
            const int n = r * 8 + k;
            if (n == 0) {
                if (VARIANT == 0 && !first_block) v = (int)(int16_t)(v - (int)cur[-64]);                     // quantize.c:22-24
                dc = v;
            } else {
                const int zp = inv_zigzag(n);
                const uint32_t bit = v != 0 ? 1u : 0u;
                if (zp < 32) lo |= bit << zp; else hi |= bit << (zp - 32);
                sizes += min(32u - (uint32_t)__clz((uint32_t)abs(v)), 11u);
            }
        }
    }
    return BlockOcc{((uint64_t)hi << 32) | lo, sizes, dc};
}
// Code length of the block: DC symbol + (8 + size) per coded AC + 8 per ZRL + END unless position 63 is coded.
__device__ __forceinline__ uint32_t block_length(const BlockOcc& o) {
    uint32_t bits = 4u + min(32u - (uint32_t)__clz((uint32_t)abs(o.dc)), 11u) + o.sizes;
    uint64_t m = o.mask & ~1ull;
    bits += 8u * (uint32_t)__popcll(m) + ((o.mask >> 63) ? 0u : 8u);                  // lossless_encode.c:43,54
    int prev = 0;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        bits += 8u * (uint32_t)((k - prev - 1) >> 4);                                  // output_ZRL per 16 zeros
        prev = k;
        m &= m - 1;
    }
    return bits;
}

// block_bits: [frame][variant][plane][block] u32.  `levels` points at the chunk's first frame; frame -1 (the
// previous chunk's last frame) precedes it in memory.  first_has_prev = 0 when frame 0 is the very first frame.
__global__ void __launch_bounds__(128)
k_enc_size(const int16_t* __restrict__ levels, uint32_t* __restrict__ block_bits, uint32_t nb, uint32_t n_frames,
           int first_has_prev) {
    const uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t total = (uint64_t)n_frames * 3 * nb;
    if (w >= total) return;
    const uint32_t f = (uint32_t)(w / ((uint64_t)3 * nb)), rem = (uint32_t)(w % ((uint64_t)3 * nb));
    const uint32_t p = rem / nb, b = rem % nb;
    const int16_t* cur = levels + w * 64;                                  // [frame][plane][block][64]
    const int16_t* prev = cur - (size_t)3 * nb * 64;                       // same plane of the previous frame
    block_bits[(((size_t)f * 2 + 0) * 3 + p) * nb + b] = block_length(block_occ<0>(cur, prev, b == 0));
    block_bits[(((size_t)f * 2 + 1) * 3 + p) * nb + b] = (f != 0 || first_has_prev) ? block_length(block_occ<1>(cur, prev, false)) : 0u;
}

// In-place exclusive scan of every (frame, variant, plane) row of block_bits; totals[row] = stream length in bits.
__global__ void __launch_bounds__(256)
k_enc_scan(uint32_t* __restrict__ block_bits, uint32_t* __restrict__ totals, uint32_t nb) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_carry;
    uint32_t* row = block_bits + (size_t)blockIdx.x * nb;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nb; i0 += 256) {
        const uint32_t i = i0 + t;
        const uint32_t x = i < nb ? row[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += a;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint32_t base = s_carry;
        for (int k = 0; k < warp; k++) base += s_w[k];
        if (i < nb) row[i] = base + inc - x;
        __syncthreads();
        if (t == 255) s_carry = base + inc;
        __syncthreads();
    }
    if (t == 0) totals[blockIdx.x] = s_carry;
}

// Per-frame emission table (host-built after the I/P decision).
struct EncFrame {
    uint64_t bit_off[3];   // bit position of each plane stream in the chunk's output buffer
    uint32_t bits[3];      // stream lengths in bits
    uint32_t variant;      // 0 = I, 1 = P (= frame_type)
};

// OR `len` (<= 43) code bits into the big-endian bit string `words` at bit position pos; bits at or beyond
// `limit` are dropped (the reference's zeroed last partial byte).
__device__ __forceinline__ void or_bits(uint32_t* __restrict__ words, uint64_t pos, uint32_t len, uint64_t bits, uint64_t limit) {
    if (len == 0 || pos >= limit) return;
    if (pos + len > limit) { const uint32_t cut = (uint32_t)(pos + len - limit); bits >>= cut; len -= cut; }
    const uint64_t w = pos >> 5;
    const uint32_t sh = (uint32_t)pos & 31u;                               // bits already used in word w
    // the code in a 96-bit window (w0:w1:w2) that starts at word w: it spans at most 3 words (sh + len <= 31 + 43)
    const uint32_t room = 96u - sh - len;                                  // zero bits below the code in the window
    const uint32_t w2 = room < 32u ? (uint32_t)(bits << room) : 0u;
    const uint64_t hi = room >= 32u ? bits << (room - 32u) : bits >> (32u - room);
    const uint32_t w0 = (uint32_t)(hi >> 32), w1 = (uint32_t)hi;
    if (w0) atomicOr(words + w, __byte_perm(w0, 0, 0x0123));
    if (w1) atomicOr(words + w + 1, __byte_perm(w1, 0, 0x0123));
    if (w2) atomicOr(words + w + 2, __byte_perm(w2, 0, 0x0123));
}

// The value coded at zig-zag position k (>= 1) of the block: re-read from the levels (L1-resident: the thread has just
// read the whole block).
__device__ __forceinline__ int coded_value(const int16_t* __restrict__ cur, const int16_t* __restrict__ prev, int k, int variant) {
    const int n = c_enc_zigzag[k];
    const int v = cur[n];
    return variant ? (int)(int16_t)(v - prev[n]) : v;
}

__global__ void __launch_bounds__(128)
k_enc_emit(const int16_t* __restrict__ levels, const uint32_t* __restrict__ block_off, const EncFrame* __restrict__ table,
           uint32_t* __restrict__ out_words, uint32_t nb, uint32_t n_frames, int fix_tail) {
    const uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t total = (uint64_t)n_frames * 3 * nb;
    if (w >= total) return;
    const uint32_t f = (uint32_t)(w / ((uint64_t)3 * nb)), rem = (uint32_t)(w % ((uint64_t)3 * nb));
    const uint32_t p = rem / nb, b = rem % nb;
    const EncFrame ef = table[f];
    const int16_t* cur = levels + w * 64;
    const int16_t* prev = cur - (size_t)3 * nb * 64;
    const int variant = (int)ef.variant;
    const BlockOcc o = variant ? block_occ<1>(cur, prev, false) : block_occ<0>(cur, prev, b == 0);
    uint64_t pos = ef.bit_off[p] + block_off[(((size_t)f * 2 + ef.variant) * 3 + p) * nb + b];
    const uint64_t limit = ef.bit_off[p] + (fix_tail ? (uint64_t)ef.bits[p] : (uint64_t)(ef.bits[p] & ~7u));
    uint32_t size, amp;
    vli(o.dc, size, amp);                                                          // output_DC :86-96
    or_bits(out_words, pos, 4u + size, ((uint64_t)size << size) | amp, limit);
    pos += 4u + size;
    uint64_t m = o.mask & ~1ull;
    int last = 0;
    while (m) {                                                                    // output_ZRL / output_AC :98-112
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        const uint32_t run = (uint32_t)(k - last - 1), nzrl = run >> 4;
        last = k;
        vli(coded_value(cur, prev, k, variant), size, amp);
        const uint64_t zrl = nzrl == 0 ? 0ull : nzrl == 1 ? 0xF0ull : nzrl == 2 ? 0xF0F0ull : 0xF0F0F0ull;
        const uint32_t len = 8u * nzrl + 8u + size;
        or_bits(out_words, pos, len, (zrl << (8u + size)) | ((uint64_t)(((run & 15u) << 4) | size) << size) | amp, limit);
        pos += len;
    }
    // END is eight zero bits: the buffer is zero-initialised, nothing to write.
}

// ---- launchers ---------------------------------------------------------------------------------------------------
cudaError_t launch_enc_transform(const void* d_frames, const int16_t* d_quant, int16_t* d_levels, uint32_t n_frames,
                                 uint32_t W, uint32_t H, cudaStream_t s) {
    const uint32_t wb = W / 8, nb = wb * (H / 8);
    const uint64_t positions = (uint64_t)n_frames * nb;
    if (!positions) return cudaSuccess;
    k_enc_transform<<<(unsigned)((positions + ENC_TPB / 8 - 1) / (ENC_TPB / 8)), ENC_TPB, 0, s>>>(
        (const uint8_t*)d_frames, d_quant, d_levels, nb, wb, W, H, n_frames);
    return cudaGetLastError();
}
cudaError_t launch_enc_size(const int16_t* d_levels, uint32_t* d_block_bits, uint32_t* d_totals, uint32_t nb,
                            uint32_t n_frames, int first_has_prev, cudaStream_t s) {
    const uint64_t warps = (uint64_t)n_frames * 3 * nb;
    if (!warps) return cudaSuccess;
    k_enc_size<<<(unsigned)((warps + 127) / 128), 128, 0, s>>>(d_levels, d_block_bits, nb, n_frames, first_has_prev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_enc_scan<<<n_frames * 6, 256, 0, s>>>(d_block_bits, d_totals, nb);
    return cudaGetLastError();
}
cudaError_t launch_enc_emit(const int16_t* d_levels, const uint32_t* d_block_off, const void* d_table, void* d_out,
                            uint32_t nb, uint32_t n_frames, int fix_tail, cudaStream_t s) {
    const uint64_t warps = (uint64_t)n_frames * 3 * nb;
    if (!warps) return cudaSuccess;
    k_enc_emit<<<(unsigned)((warps + 127) / 128), 128, 0, s>>>(d_levels, d_block_off, (const EncFrame*)d_table,
                                                            (uint32_t*)d_out, nb, n_frames, fix_tail);
    return cudaGetLastError();
}

}  // namespace mj

// ---- host API ---------------------------------------------------------------------------------------------------
using namespace mj;

#define CUE(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

extern "C" size_t mjpeg423_b200_encode_bound(uint32_t n, uint32_t w_size, uint32_t h_size) {
    const size_t nb = (size_t)(w_size / 8) * (h_size / 8);
    return 20 + (size_t)n * (16 + 3 * nb * 160 + 4 + 8) + 512;       // <= 1212 bits per block, SURVEY.md A.6
}

static int encode_frames_impl(mjpeg423_b200_ctx* c, const void* frames, int frames_on_device, uint32_t n,
                              uint32_t W, uint32_t H, uint32_t max_I_interval, uint32_t flags,
                              uint8_t* mpg, size_t cap, size_t* mpg_len) {
    if (!c || !mpg || !mpg_len || (n && !frames)) return MJPEG423_E_ARG;
    if (!W || !H || (W & 7) || (H & 7)) { set_error("encode: W and H must be non-zero multiples of 8"); return MJPEG423_E_ARG; }
    CUE(cudaSetDevice(c->device));
    const uint32_t wb = W / 8, nb = wb * (H / 8);
    const size_t frame_bytes = (size_t)W * H * 4, level_frame = (size_t)3 * nb * 64;       // int16 elements
    const int fix_tail = (flags & MJPEG423_ENC_FIX_TAIL) ? 1 : 0;
    if (cap < 20 + 512) { set_error("encode: output buffer too small"); return MJPEG423_E_NOMEM; }
    cudaStream_t s = c->s_compute;
    // chunk: bounded scratch (levels 384 B + lengths 24 B per block, + the frames when they come from the host)
    const size_t per_frame = level_frame * 2 + (size_t)nb * 24 + (frames_on_device ? 0 : frame_bytes);
    uint32_t K = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n ? n : 1, ((uint64_t)2 << 30) / per_frame));
    if (c->chunk_frames) K = std::min(K, c->chunk_frames);
    int rc;
    if ((rc = c->coef[0].reserve((size_t)(K + 1) * level_frame * 2))) return rc;          // slot 0 = previous chunk's last frame
    if ((rc = c->blkidx[0].reserve((size_t)K * 6 * nb * 4 + (size_t)K * 6 * 4 + 256))) return rc;
    if ((rc = c->ids.reserve((size_t)K * sizeof(EncFrame) + 64))) return rc;
    if (!frames_on_device && (rc = c->in_ring[0].reserve((size_t)K * frame_bytes))) return rc;
    int16_t* d_levels = c->coef[0].as<int16_t>() + level_frame;                            // frame 0 of the chunk
    uint32_t* d_bits = c->blkidx[0].as<uint32_t>();
    uint32_t* d_totals = d_bits + (size_t)K * 6 * nb;
    std::vector<uint32_t> totals((size_t)K * 6);
    std::vector<EncFrame> table(K);
    std::vector<uint32_t> trailer, headers((size_t)K * 4);
    size_t pos = 20;
    c->chunk_K = 0;                                    // c->ids is about to be overwritten: a later decode rebuilds its chunk table
    uint32_t last_i = 0;
    c->stats = mjpeg423_b200_stats{};
    CUE(cudaEventRecord(c->ev[0], s));
    for (uint32_t f0 = 0; f0 < n; f0 += K) {
        const uint32_t k = std::min(K, n - f0);
        const void* d_frames = frames;
        if (frames_on_device) d_frames = (const uint8_t*)frames + (size_t)f0 * frame_bytes;
        else {
            CUE(cudaMemcpyAsync(c->in_ring[0].p, (const uint8_t*)frames + (size_t)f0 * frame_bytes, (size_t)k * frame_bytes,
                                cudaMemcpyHostToDevice, s));
            d_frames = c->in_ring[0].p;
        }
        CUE(launch_enc_transform(d_frames, c->d_quant, d_levels, k, W, H, s));
        CUE(launch_enc_size(d_levels, d_bits, d_totals, nb, k, f0 != 0, s));
        CUE(cudaMemcpyAsync(totals.data(), d_totals, (size_t)k * 6 * 4, cudaMemcpyDeviceToHost, s));
        CUE(cudaStreamSynchronize(s));
        c->stats.kernel_launches += 3;
        // I/P decision and layout of the chunk's frame records (mjpeg423_encoder.c:155-199)
        const size_t chunk_pos = pos;
        for (uint32_t i = 0; i < k; i++) {
            const uint32_t f = f0 + i;
            uint32_t szI[3], szP[3];
            for (int p = 0; p < 3; p++) { szI[p] = (totals[(size_t)i * 6 + p] + 7) / 8; szP[p] = (totals[(size_t)i * 6 + 3 + p] + 7) / 8; }
            const uint32_t totI = szI[0] + szI[1] + szI[2], totP = szP[0] + szP[1] + szP[2];
            const bool isI = f == 0 || totI <= totP || f - last_i >= max_I_interval;
            if (isI) { last_i = f; trailer.push_back(f); trailer.push_back((uint32_t)pos); }
            const uint32_t* sz = isI ? szI : szP;
            const uint32_t body = sz[0] + sz[1] + sz[2];
            const uint32_t fsz = (body + 16 + 3u) & ~3u;
            if (pos + fsz + 512 > cap || pos + fsz > 0xFFFFFFFFull) { set_error("encode: output buffer too small"); return MJPEG423_E_NOMEM; }
            const uint32_t hdr[4] = {fsz, isI ? 0u : 1u, sz[0], sz[1]};   // frame header (:188-191), little-endian
            std::memcpy(&headers[(size_t)i * 4], hdr, 16);
            EncFrame& ef = table[i];
            ef.variant = isI ? 0 : 1;
            size_t o = pos - chunk_pos + 16;
            for (int p = 0; p < 3; p++) { ef.bit_off[p] = (uint64_t)o * 8; ef.bits[p] = totals[(size_t)i * 6 + ef.variant * 3 + p]; o += sz[p]; }
            pos += fsz;
        }
        const size_t chunk_bytes = pos - chunk_pos;
        if ((rc = c->out_ring[0].reserve(chunk_bytes + 16))) return rc;
        CUE(cudaMemsetAsync(c->out_ring[0].p, 0, chunk_bytes + 16, s));
        CUE(cudaMemcpyAsync(c->ids.p, table.data(), (size_t)k * sizeof(EncFrame), cudaMemcpyHostToDevice, s));
        CUE(launch_enc_emit(d_levels, d_bits, c->ids.p, c->out_ring[0].p, nb, k, fix_tail, s));
        c->stats.kernel_launches += 1;
        CUE(cudaMemcpyAsync(mpg + chunk_pos, c->out_ring[0].p, chunk_bytes, cudaMemcpyDeviceToHost, s));
        // the last frame's levels are the P-frame reference of the next chunk
        CUE(cudaMemcpyAsync(c->coef[0].p, d_levels + (size_t)(k - 1) * level_frame, level_frame * 2, cudaMemcpyDeviceToDevice, s));
        CUE(cudaStreamSynchronize(s));
        for (size_t i = 0, o = chunk_pos; i < k; i++) {                    // the 16-byte frame headers go in on the host
            std::memcpy(mpg + o, &headers[i * 4], 16);
            o += headers[i * 4];
        }
    }
    CUE(cudaEventRecord(c->ev[1], s));
    CUE(cudaEventSynchronize(c->ev[1]));
    CUE(cudaEventElapsedTime(&c->stats.total_ms, c->ev[0], c->ev[1]));
    c->stats.frames = n;
    // trailer, 512 pad bytes (zeros here), header (mjpeg423_encoder.c:82-88,209-225)
    const uint32_t n_i = (uint32_t)(trailer.size() / 2);
    if (pos + trailer.size() * 4 + 512 > cap) { set_error("encode: output buffer too small"); return MJPEG423_E_NOMEM; }
    const uint32_t hdr[5] = {n, W, H, n_i, (uint32_t)(pos - 20)};
    std::memcpy(mpg, hdr, 20);
    if (!trailer.empty()) std::memcpy(mpg + pos, trailer.data(), trailer.size() * 4);
    pos += trailer.size() * 4;
    std::memset(mpg + pos, 0, 512);
    pos += 512;
    c->stats.payload_bytes = pos;
    *mpg_len = pos;
    return MJPEG423_OK;
}
extern "C" int mjpeg423_b200_encode_frames(mjpeg423_b200_ctx* c, const void* frames, int frames_on_device, uint32_t n,
                                           uint32_t W, uint32_t H, uint32_t max_I_interval, uint32_t flags,
                                           uint8_t* mpg, size_t cap, size_t* mpg_len) {
    return mj::guard([&]() -> int { return encode_frames_impl(c, frames, frames_on_device, n, W, H, max_I_interval, flags, mpg, cap, mpg_len); });
}
