// entropy.cu -- batched, segment-parallel "lossless" decode + dequantise (sm_100a).
//
// Replaces lossless_decode(), LIB/decoder/lossless_decode.c:60-135 (LIB =
// /root/reference/core0/software/common/libs/mjpeg423), for MANY plane streams at once and, inside a
// stream, for many fixed-size bitstream segments in parallel.  The code has no markers or restart
// intervals, so segment entry points are found by self-synchronisation:
//
//   k_entropy_sync   one thread per segment parses speculatively from the segment's first bit as if a
//                    block started there, leaving NCP checkpoints (first block start at or after every
//                    CP_BITS boundary, blocks and DC sum so far).  The predecessor's speculative exit
//                    is then taken as the segment's entry and parsed only until it MERGES with the
//                    recorded trajectory (same bit position at a block start => identical future).
//   k_entropy_chain  one CTA per stream: re-parses the few segments whose entry differs from the
//                    predecessor's resolved exit until the chain entry[i] == exit[i-1] holds from
//                    entry[0] = 0 (correctness never rests on self-synchronisation, only speed does),
//                    then exclusive-scans block counts and DC sums (mod 2^16) to
//                    give every segment its first block index and DC predictor.
//   k_entropy_write  one thread per segment decodes its blocks from the now exact state, dequantises
//                    and scatters int16 coefficients in zig-zag -> natural order into a plane the CTA
//                    has just zero-filled with coalesced 128-bit stores (the memset at :77-78).
//
// All three parse with the same parse_block() (common.cuh), so they follow one trajectory function.
#include "common.cuh"
#include "runtime.h"

namespace mj {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct ParseSink {          // parse-only: accumulate the DC delta (mod 2^16 is taken by the caller)
    int dcsum = 0;
    __device__ __forceinline__ void dc(int e) { dcsum += e; }
    __device__ __forceinline__ void ac(uint32_t, int) {}
};

__device__ __forceinline__ uint32_t block_budget(uint32_t pos, uint32_t total_bits) {
    return min(RUNAWAY_BITS, total_bits - pos);
}

// Parse segment `seg` from block start `entry` to the first block start at or after the segment end
// (or the end of the stream).  Returns exit; cnt / dc receive the blocks started and their DC sum.
__device__ __forceinline__ uint32_t parse_segment(const uint8_t* base, uint32_t entry, uint32_t seg_end,
                                                  uint32_t total_bits, uint32_t& cnt, uint32_t& dc) {
    uint32_t pos = entry;
    ParseSink sink;
    cnt = 0;
    if (pos < seg_end && pos + MIN_BLOCK_BITS <= total_bits) {
        BitReader r;
        r.init(base, pos);
        do {
            pos += parse_block(r, block_budget(pos, total_bits), sink);
            cnt++;
        } while (pos < seg_end && pos + MIN_BLOCK_BITS <= total_bits);
    }
    dc = (uint32_t)sink.dcsum & 0xFFFFu;
    return pos;
}

// ------------------------------------------------------------------------------------------------
// Speculative parse + merge.  A tile owns ENT_TPB-1 segments; thread 0 parses the segment BEFORE the
// tile (halo) so that thread 1 has a predecessor exit without any inter-CTA dependency.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENT_TPB)
k_entropy_sync(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const TileDesc* __restrict__ tiles, uint32_t* __restrict__ seg_entry,
               uint32_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_cd) {
    __shared__ uint32_t s_pos[NCP][ENT_TPB];   // [j] = first block start >= seg_start + (j+1)*CP_BITS
    __shared__ uint32_t s_cd[NCP][ENT_TPB];    // blocks started before it | DC sum of them << 16
    const int t = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const StreamDesc sd = streams[td.stream];
    const int seg = (int)td.seg0 - 1 + t;
    const bool valid = seg >= 0 && seg < (int)sd.nseg;
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t total_bits = sd.byte_len * 8u;
    const uint32_t seg_start = (uint32_t)seg * SEG_BITS, seg_end = seg_start + SEG_BITS;

    if (valid) {
        uint32_t pos = seg_start, cnt = 0;
        ParseSink sink;
        BitReader r;
        r.init(base, pos);
        int j = 0;
        for (;;) {
            while (j < NCP && pos >= seg_start + (uint32_t)(j + 1) * CP_BITS) {
                s_pos[j][t] = pos;
                s_cd[j][t] = cnt | ((uint32_t)sink.dcsum << 16);
                j++;
            }
            if (j == NCP) break;
            if (pos + MIN_BLOCK_BITS > total_bits) {       // end of stream: no further block can start
                for (; j < NCP; j++) { s_pos[j][t] = pos; s_cd[j][t] = cnt | ((uint32_t)sink.dcsum << 16); }
                break;
            }
            pos += parse_block(r, block_budget(pos, total_bits), sink);
            cnt++;
        }
    }
    __syncthreads();
    if (!valid || t == 0) return;

    const uint32_t spec_exit = s_pos[NCP - 1][t], spec_cd = s_cd[NCP - 1][t];
    const uint32_t E = seg == 0 ? 0u : s_pos[NCP - 1][t - 1];
    uint32_t exit_pos, cnt = 0, dc = 0;
    if (E >= seg_end) {                 // a block spans the whole segment: it owns nothing
        exit_pos = E;
    } else if (E == seg_start) {        // speculation started on the true entry
        exit_pos = spec_exit; cnt = spec_cd & 0xFFFFu; dc = spec_cd >> 16;
    } else {
        uint32_t pos = E;
        ParseSink sink;
        BitReader r;
        r.init(base, pos);
        for (;;) {
            if (pos >= seg_start + CP_BITS) {
                int j = (int)min((uint32_t)NCP, (pos - seg_start) / CP_BITS) - 1;
                if (s_pos[j][t] == pos) {                  // merged with the speculative trajectory
                    uint32_t at = s_cd[j][t];
                    cnt += (spec_cd & 0xFFFFu) - (at & 0xFFFFu);
                    sink.dcsum += (int)(spec_cd >> 16) - (int)(at >> 16);
                    pos = spec_exit;
                    break;
                }
                if (pos >= seg_end) break;
            }
            if (pos + MIN_BLOCK_BITS > total_bits) break;
            pos += parse_block(r, block_budget(pos, total_bits), sink);
            cnt++;
        }
        exit_pos = pos;
        dc = (uint32_t)sink.dcsum & 0xFFFFu;
    }
    const uint32_t g = sd.seg_base + (uint32_t)seg;
    seg_entry[g] = E;
    seg_exit[g] = exit_pos;
    seg_cd[g] = (cnt & 0xFFFFu) | (dc << 16);
}

// ------------------------------------------------------------------------------------------------
// Chain fix-up + scans: one CTA per stream.
// ------------------------------------------------------------------------------------------------
constexpr int CHAIN_TPB = 128;

__global__ void __launch_bounds__(CHAIN_TPB)
k_entropy_chain(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                uint32_t* seg_entry, uint32_t* seg_exit, uint32_t* seg_cd, uint32_t* __restrict__ seg_first,
                uint32_t* __restrict__ stream_blocks, unsigned long long* __restrict__ fixups) {
    const StreamDesc sd = streams[blockIdx.x];
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t total_bits = sd.byte_len * 8u;
    volatile uint32_t* v_entry = seg_entry + sd.seg_base;
    volatile uint32_t* v_exit = seg_exit + sd.seg_base;
    volatile uint32_t* v_cd = seg_cd + sd.seg_base;
    const int t = threadIdx.x;

    // 1. make the chain exact: entry[0] = 0, entry[i] = exit[i-1].
    uint32_t nfix = 0;
    for (uint32_t sweep = 0; sweep <= sd.nseg + 1; sweep++) {   // converges in <= nseg sweeps; bound it anyway
        int changed = 0;
        for (uint32_t i = t; i < sd.nseg; i += CHAIN_TPB) {
            uint32_t E = i ? v_exit[i - 1] : 0u;
            if (v_entry[i] != E) {
                uint32_t cnt, dc;
                uint32_t x = parse_segment(base, E, (i + 1) * SEG_BITS, total_bits, cnt, dc);
                v_entry[i] = E;
                v_exit[i] = x;
                v_cd[i] = (cnt & 0xFFFFu) | (dc << 16);
                changed = 1;
                nfix++;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (nfix) atomicAdd(fixups, (unsigned long long)nfix);

    // 2. exclusive scans over segments: first block index, DC predictor (mod 2^16).
    __shared__ uint32_t s_wcnt[CHAIN_TPB / 32], s_wdc[CHAIN_TPB / 32];
    __shared__ uint32_t s_carry[2];
    if (t == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    for (uint32_t i0 = 0; i0 < sd.nseg; i0 += CHAIN_TPB) {
        uint32_t i = i0 + t;
        uint32_t cd = i < sd.nseg ? v_cd[i] : 0u;
        uint32_t cnt = cd & 0xFFFFu, dc = cd >> 16;
        uint32_t icnt = cnt, idc = dc;                    // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t a = __shfl_up_sync(0xFFFFFFFFu, icnt, d), b = __shfl_up_sync(0xFFFFFFFFu, idc, d);
            if (lane >= d) { icnt += a; idc += b; }
        }
        if (lane == 31) { s_wcnt[warp] = icnt; s_wdc[warp] = idc; }
        __syncthreads();
        uint32_t bcnt = s_carry[0], bdc = s_carry[1];
        for (int w = 0; w < warp; w++) { bcnt += s_wcnt[w]; bdc += s_wdc[w]; }
        if (i < sd.nseg) {
            seg_first[sd.seg_base + i] = bcnt + icnt - cnt;
            v_cd[i] = cnt | (((bdc + idc - dc) & 0xFFFFu) << 16);   // DC predictor entering the segment
        }
        __syncthreads();
        if (t == CHAIN_TPB - 1) { s_carry[0] = bcnt + icnt; s_carry[1] = bdc + idc; }
        __syncthreads();
    }
    if (t == 0) stream_blocks[blockIdx.x] = s_carry[0];
}

// ------------------------------------------------------------------------------------------------
// Coefficient write.  quant: 2 tables x 64 int16, natural order (LIB/common/tables.c:13-32 by default).
// ------------------------------------------------------------------------------------------------
template <bool PFRAME>
struct WriteSink {
    int16_t* out;               // current block, natural order
    const uint32_t* zq;         // smem: natural index | quant << 16, by zig-zag position
    int cur;                    // I frames: running DC (LIB/decoder/lossless_decode.c:73,94)
    __device__ __forceinline__ void dc(int e) {
        int q0 = (int)(zq[0] >> 16);
        if (PFRAME) out[0] = (int16_t)(out[0] + e * q0);          // :91
        else { cur += e; out[0] = (int16_t)((int)(int16_t)cur * q0); }   // :94-95
    }
    __device__ __forceinline__ void ac(uint32_t idx, int e) {
        uint32_t z = zq[idx];
        int n = (int)(z & 0xFFFFu), q = (int)(z >> 16);
        if (PFRAME) out[n] = (int16_t)(out[n] + e * q);           // :122
        else out[n] = (int16_t)(e * q);                           // :125
    }
};

__global__ void __launch_bounds__(ENT_TPB)
k_entropy_write(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ seg_entry,
                const uint32_t* __restrict__ seg_cd, const uint32_t* __restrict__ seg_first,
                const int16_t* __restrict__ quant, int16_t* __restrict__ coef) {
    __shared__ uint32_t s_zq[64];
    __shared__ uint32_t s_range[2];
    const int t = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const StreamDesc sd = streams[td.stream];
    const uint32_t seg = td.seg0 + (uint32_t)t;
    const bool valid = seg < sd.nseg;
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t total_bits = sd.byte_len * 8u;
    if (t < 64) {
        uint32_t n = c_zigzag[t];
        s_zq[t] = n | ((uint32_t)(uint16_t)quant[sd.quant_id * 64 + n] << 16);
    }
    uint32_t E = 0, first = 0, cnt = 0, dc0 = 0;
    if (valid) {
        const uint32_t g = sd.seg_base + seg;
        E = seg_entry[g];
        first = seg_first[g];
        uint32_t cd = seg_cd[g];
        cnt = cd & 0xFFFFu;
        dc0 = cd >> 16;
        if (first >= sd.nb) cnt = 0;
        else cnt = min(cnt, sd.nb - first);       // trailing pad bits can look like blocks
    }
    int16_t* plane = coef + (size_t)sd.block_base * 64;
    const bool last_tile = td.seg0 + ENT_TPB >= sd.nseg;
    if (!sd.ptype) {
        // Zero-fill the tile's contiguous block range (the memset of :77-78), coalesced.
        if (t == 0) s_range[0] = min(first, sd.nb);
        if (valid && (seg + 1 == sd.nseg || t == ENT_TPB - 1)) s_range[1] = last_tile ? sd.nb : first + cnt;
        __syncthreads();
        uint4* z = reinterpret_cast<uint4*>(plane + (size_t)s_range[0] * 64);
        const uint32_t nvec = (s_range[1] - s_range[0]) * 8u;     // 8 x 16 B per block
        for (uint32_t i = t; i < nvec; i += ENT_TPB) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (cnt == 0) return;

    BitReader r;
    r.init(base, E);
    uint32_t pos = E;
    if (sd.ptype) {
        WriteSink<true> sink{plane + (size_t)first * 64, s_zq, 0};
        for (uint32_t k = 0; k < cnt; k++, sink.out += 64) pos += parse_block(r, block_budget(pos, total_bits), sink);
    } else {
        WriteSink<false> sink{plane + (size_t)first * 64, s_zq, (int)dc0};
        for (uint32_t k = 0; k < cnt; k++, sink.out += 64) pos += parse_block(r, block_budget(pos, total_bits), sink);
    }
}

// ------------------------------------------------------------------------------------------------
// Host launchers (declared in runtime.h).
// ------------------------------------------------------------------------------------------------
cudaError_t launch_entropy_sync(const EntropyJob& j, cudaStream_t s) {
    if (j.n_sync_tiles == 0) return cudaSuccess;
    k_entropy_sync<<<j.n_sync_tiles, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_sync_tiles, j.d_seg_entry,
                                                      j.d_seg_exit, j.d_seg_cd);
    return cudaGetLastError();
}
cudaError_t launch_entropy_chain(const EntropyJob& j, cudaStream_t s) {
    if (j.n_streams == 0) return cudaSuccess;
    k_entropy_chain<<<j.n_streams, CHAIN_TPB, 0, s>>>(j.d_payload, j.d_streams + j.stream_lo, j.d_seg_entry,
                                                      j.d_seg_exit, j.d_seg_cd, j.d_seg_first,
                                                      j.d_stream_blocks + j.stream_lo, j.d_fixups);
    return cudaGetLastError();
}
cudaError_t launch_entropy_write(const EntropyJob& j, const int16_t* d_quant, int16_t* d_coef, cudaStream_t s) {
    if (j.n_write_tiles == 0) return cudaSuccess;
    k_entropy_write<<<j.n_write_tiles, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_write_tiles, j.d_seg_entry,
                                                        j.d_seg_cd, j.d_seg_first, d_quant, d_coef);
    return cudaGetLastError();
}

}  // namespace mj
