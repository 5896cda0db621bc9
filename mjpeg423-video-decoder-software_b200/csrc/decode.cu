// decode.cu -- block-parallel final decode (sm_100a): one thread per 8x8 block.
//
// After entropy.cu has produced the block index (bit position of every block's DC symbol + absolute DC
// level), every block of every plane can be decoded independently:
//
//   k_decode_coef   lossless_decode() output, LIB/decoder/lossless_decode.c:60-135: the thread parses its
//                   block into a 128-byte shared-memory slot (zeroed = the memset of :77-78, or preloaded
//                   with the previous frame's coefficients for P frames, :90-92,121-123), dequantising
//                   as it scatters zig-zag -> natural order (:122-126); the CTA then stores its 128
//                   consecutive blocks as one contiguous, fully coalesced 16 KB run.
//   k_decode_fused  the whole reference loop body, LIB/decoder/mjpeg423_decoder.c:110-124, for intra
//                   frames: the thread parses the Y, Cb and Cr block of one block position through the
//                   same slot, runs the three IDCTs in registers (idct.c:22-181) and writes the 8x8
//                   BGRA pixels (ycbcr_to_rgb.c:26-49).  Coefficients and samples never touch HBM:
//                   the kernel reads C + 6 bytes of index per block and writes 4 bytes per pixel.
// (LIB = /root/reference/core0/software/common/libs/mjpeg423.)
//
// Slots use the XOR swizzle of idct_colour.cu (16-byte chunk r of slot t at chunk r ^ (t & 7)), so both
// the row reads of the IDCT and the cooperative copy-out are bank-conflict free.
#include "common.cuh"
#include "runtime.h"

namespace mj {

constexpr int DEC_TPB = 128;

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Byte offset of natural-order coefficient n inside thread t's swizzled slot.
__device__ __forceinline__ uint32_t slot_off(int t, uint32_t n) {
    return (uint32_t)t * 128u + ((((n >> 3) ^ (uint32_t)t) & 7u) << 4) + ((n & 7u) << 1);
}

// Sink of parse_block(): dequantise and scatter into the slot, tracking column occupancy for the IDCT.
template <bool PFRAME>
struct SlotSink {
    uint8_t* smem;              // slot array
    const uint32_t* zq;         // smem: natural index | quant << 16, by zig-zag position
    int t;
    int cur;                    // I frames: absolute DC level from the block index
    uint32_t m_ac, m_any;       // column masks (see block_masks() in common.cuh)
    __device__ __forceinline__ int16_t& at(uint32_t n) { return *reinterpret_cast<int16_t*>(smem + slot_off(t, n)); }
    __device__ __forceinline__ void dc(int e) {
        const int q0 = (int)(zq[0] >> 16);
        int16_t& d = at(0);
        if (PFRAME) d = (int16_t)(d + e * q0);                     // lossless_decode.c:91
        else d = (int16_t)((int)(int16_t)cur * q0);                // :94-95 (cur already includes e)
        m_any |= 1u;
    }
    __device__ __forceinline__ void ac_(uint32_t n, int v) {
        int16_t& d = at(n);
        if (PFRAME) d = (int16_t)(d + v);                          // :122
        else d = (int16_t)v;                                       // :125
        m_any |= 1u << (n & 7u);
        if (n >= 8u) m_ac |= 1u << (n & 7u);
    }
    __device__ __forceinline__ void ac(uint32_t idx, int e) {     // never called with idx >= 64
        const uint32_t z = zq[idx];
        ac_(z & 0xFFFFu, e * (int)(z >> 16));
    }
};

__device__ __forceinline__ void load_zq(uint32_t* s_zq, const int16_t* quant, int t) {
    if (t < 64) {
        const uint32_t n = c_zigzag[t];
        s_zq[t] = n | ((uint32_t)(uint16_t)quant[n] << 16);
    }
}
__device__ __forceinline__ void zero_slot(uint8_t* smem, int t) {
#pragma unroll
    for (int r = 0; r < 8; r++) *reinterpret_cast<uint4*>(smem + t * 128 + ((r ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void load_slot_rows(const uint8_t* smem, int t, uint4 (&rows)[8]) {
#pragma unroll
    for (int r = 0; r < 8; r++) rows[r] = *reinterpret_cast<const uint4*>(smem + t * 128 + ((r ^ (t & 7)) << 4));
}

// ---- coefficient planes ------------------------------------------------------------------------------
// grid = (ceil(nb / 128), number of streams in `stream_ids`).
__global__ void __launch_bounds__(DEC_TPB)
k_decode_coef(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
              const uint32_t* __restrict__ stream_ids, const uint32_t* __restrict__ blk_pos,
              const int16_t* __restrict__ blk_dc, const int16_t* __restrict__ quant, int16_t* coef) {
    __shared__ __align__(128) uint8_t slots[DEC_TPB * 128];
    __shared__ uint32_t s_zq[64];
    const int t = threadIdx.x;
    const StreamDesc sd = streams[stream_ids[blockIdx.y]];
    const uint32_t b0 = blockIdx.x * DEC_TPB;
    if (b0 >= sd.nb) return;
    const uint32_t nblk = min((uint32_t)DEC_TPB, sd.nb - b0);
    load_zq(s_zq, quant + sd.quant_id * 64, t);
    uint4* dst = reinterpret_cast<uint4*>(coef + ((size_t)sd.block_base + b0) * 64);
    if (sd.ptype) {
        // Preload the previous frame's coefficients (coalesced, swizzled) -- the P-frame state.
        const uint4* src = reinterpret_cast<const uint4*>(coef + ((size_t)sd.prev_base + b0) * 64);
        for (uint32_t i = t; i < nblk * 8u; i += DEC_TPB) {
            const uint32_t blk = i >> 3, row = i & 7u;
            *reinterpret_cast<uint4*>(slots + blk * 128u + ((row ^ (blk & 7u)) << 4)) = src[i];
        }
    } else {
        zero_slot(slots, t);
    }
    __syncthreads();
    if ((uint32_t)t < nblk) {
        const uint32_t gb = sd.block_base + b0 + (uint32_t)t;
        const uint32_t pos = blk_pos[gb];
        if (pos != NO_BLOCK) {
            const uint8_t* base = payload + sd.byte_off;
            if (sd.ptype) {
                SlotSink<true> sink{slots, s_zq, t, 0, 0, 0};
                parse_block(base, pos, sd.byte_len * 8u, sink);
            } else {
                SlotSink<false> sink{slots, s_zq, t, (int)blk_dc[gb], 0, 0};
                parse_block(base, pos, sd.byte_len * 8u, sink);
            }
        }
    }
    __syncthreads();
    for (uint32_t i = t; i < nblk * 8u; i += DEC_TPB) {      // 16 KB contiguous, 512 B per warp instruction
        const uint32_t blk = i >> 3, row = i & 7u;
        dst[i] = *reinterpret_cast<const uint4*>(slots + blk * 128u + ((row ^ (blk & 7u)) << 4));
    }
}

// ---- fully fused: bitstream + block index -> BGRA ---------------------------------------------------------
// grid = (ceil(nb / 128), frames); streams of frame f are streams[stream_lo + 3f + {0,1,2}].
__global__ void __launch_bounds__(DEC_TPB, 4)
k_decode_fused(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const uint32_t* __restrict__ blk_pos, const int16_t* __restrict__ blk_dc,
               const int16_t* __restrict__ quant, uint8_t* __restrict__ out, uint32_t nb, uint32_t wb, uint32_t W) {
    __shared__ __align__(128) uint8_t slots[DEC_TPB * 128];
    __shared__ uint32_t s_zq[2][64];
    const int t = threadIdx.x;
    const uint32_t f = blockIdx.y;
    const uint32_t b = blockIdx.x * DEC_TPB + (uint32_t)t;
    const bool live = b < nb;
    load_zq(s_zq[0], quant, t);
    load_zq(s_zq[1], quant + 64, t);
    __syncthreads();
    uint32_t px[3][16];
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const StreamDesc* sd = streams + (size_t)f * 3 + p;
        zero_slot(slots, t);                                    // thread-private slot: no barrier needed
        SlotSink<false> sink{slots, s_zq[p ? 1 : 0], t, 0, 0, 0};
        if (live) {
            const uint32_t gb = sd->block_base + b;
            const uint32_t pos = blk_pos[gb];
            if (pos != NO_BLOCK) {
                sink.cur = (int)blk_dc[gb];
                parse_block(payload + sd->byte_off, pos, sd->byte_len * 8u, sink);
            }
        }
        __syncwarp();
        const uint32_t acm = warp_or(sink.m_ac), anym = warp_or(sink.m_any);
        uint4 rows[8];
        load_slot_rows(slots, t, rows);
        idct_block(rows, acm, anym, px[p]);
    }
    if (!live) return;
    uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;
#pragma unroll
    for (int r = 0; r < 8; r++)
        colour_row_store(px[0][2 * r], px[0][2 * r + 1], px[1][2 * r], px[1][2 * r + 1], px[2][2 * r], px[2][2 * r + 1],
                         dst + (size_t)r * W * 4);
}

// ---- launchers -----------------------------------------------------------------------------------------------
cudaError_t launch_decode_coef(const EntropyJob& j, const uint32_t* d_stream_ids, uint32_t n_ids, uint32_t nb,
                               const int16_t* d_quant, int16_t* d_coef, cudaStream_t s) {
    if (n_ids == 0 || nb == 0) return cudaSuccess;
    dim3 grid((nb + DEC_TPB - 1) / DEC_TPB, n_ids);
    k_decode_coef<<<grid, DEC_TPB, 0, s>>>(j.d_payload, j.d_streams, d_stream_ids, j.d_blk_pos, j.d_blk_dc, d_quant, d_coef);
    return cudaGetLastError();
}
cudaError_t launch_decode_fused(const EntropyJob& j, const int16_t* d_quant, void* d_out, uint32_t n_frames,
                                uint32_t W, uint32_t H, cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    const uint32_t wb = W / 8, nb = wb * (H / 8);
    dim3 grid((nb + DEC_TPB - 1) / DEC_TPB, n_frames);
    k_decode_fused<<<grid, DEC_TPB, 0, s>>>(j.d_payload, j.d_streams + j.stream_lo, j.d_blk_pos, j.d_blk_dc, d_quant,
                                            (uint8_t*)d_out, nb, wb, W);
    return cudaGetLastError();
}

}  // namespace mj
