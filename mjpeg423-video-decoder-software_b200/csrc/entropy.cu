// entropy.cu -- batched, segment-parallel "lossless" decode (sm_100a): synchronisation, chain, block index.
//
// Replaces the serial walk of lossless_decode(), LIB/decoder/lossless_decode.c:60-135 (LIB =
// /root/reference/core0/software/common/libs/mjpeg423), for MANY plane streams at once and, inside a
// stream, for many fixed-size bitstream segments in parallel.  The code has no markers or restart
// intervals, so segment entry points are found by self-synchronisation:
//
//   k_entropy_sync   one thread per segment parses speculatively from the segment's first bit as if a
//                    block started there, leaving NCP checkpoints (first block start at or after every
//                    CP_BITS boundary, blocks and DC sum so far).  The predecessor's speculative exit
//                    is then taken as the segment's entry and parsed only until it MERGES with the
//                    recorded trajectory (same bit position at a block start => identical future).
//                    A second in-CTA round re-merges the segments whose predecessor's exit moved.
//   k_entropy_chain  one CTA per stream: re-parses the few segments whose entry still differs from the
//                    predecessor's resolved exit until the chain entry[i] == exit[i-1] holds from
//                    entry[0] = 0 (correctness never rests on self-synchronisation, only speed does),
//                    then exclusive-scans block counts and DC sums (mod 2^16) to
//                    give every segment its first block index and DC predictor.
//   k_entropy_index  one thread per segment walks its blocks from the now exact state and writes every
//                    coded coefficient as a 32-bit entry into the segment's symbol list, plus per block
//                    the list position, entry count and absolute DC level: what the block-parallel
//                    decode kernels (decode.cu) consume without touching the bitstream again.
//
// Every pass advances with Parser::step() (common.cuh): one flat loop, one symbol per iteration per
// lane, DC/AC and block-end handling predicated, so the lanes of a warp stay converged.
#include "common.cuh"
#include "runtime.h"

namespace mj {

// All positions inside the kernels are "f" positions (stream bit position + stream_bias(), see Parser);
// the per-segment arrays in global memory hold plain stream bit positions.
//
// fstop_eos: a block start at or after this f position cannot hold a block any more (fewer than
// MIN_BLOCK_BITS left): the stream's trailing pad bits.
__device__ __forceinline__ uint32_t eos_stop(uint32_t ftotal) {
    return ftotal >= (uint32_t)MIN_BLOCK_BITS ? ftotal - (MIN_BLOCK_BITS - 1) : 0u;
}

// Parse from block start `entry` to the first block start at or after seg_end (or the end of the
// stream).  Returns the exit position; cnt / dc receive the blocks started and their DC sum mod 2^16.
// (f positions.)
__device__ __forceinline__ uint32_t parse_segment(const uint8_t* base, uint32_t entry, uint32_t seg_end,
                                                  uint32_t ftotal, uint32_t& cnt, uint32_t& dc) {
    cnt = 0;
    int dcsum = 0;
    uint32_t pos = entry;
    const uint32_t stop = min(seg_end, eos_stop(ftotal));
    if (pos < stop) {
        Parser ps;
        ps.start(base, entry, ftotal);
        for (;;) {
            Parser::Sym y;
            const bool end = ps.step(ftotal, y);
            dcsum += y.dc ? y.e : 0;
            if (end) {
                cnt++;
                if (ps.fpos >= stop) break;
            }
        }
        pos = ps.fpos;
    }
    dc = (uint32_t)dcsum & 0xFFFFu;
    return pos;
}

// ------------------------------------------------------------------------------------------------
// Speculative parse + merge.  A tile owns ENT_TPB-1 segments; thread 0 parses the segment BEFORE the
// tile (halo) so that thread 1 has a predecessor exit without any inter-CTA dependency.
// ------------------------------------------------------------------------------------------------
struct Resolved { uint32_t exit_pos, cnt, dc; };

// Resolve segment [seg_start, seg_end) for true entry E against the recorded speculative trajectory.
// WARP-COLLECTIVE (lanes with nothing to resolve pass need = false): the symbol loop runs under a warp
// vote so the lanes re-converge every iteration.  The true trajectory is compared with the recorded one
// at its first block start at or after every checkpoint boundary (once merged, that IS the checkpoint).
__device__ __forceinline__ Resolved resolve_by_merge(const uint8_t* base, uint32_t E, uint32_t seg_start,
                                                     uint32_t seg_end, uint32_t ftotal,
                                                     const uint32_t (*s_pos)[ENT_TPB], const uint32_t (*s_cd)[ENT_TPB],
                                                     int t, bool need) {
    const uint32_t spec_exit = s_pos[NCP - 1][t], spec_cd = s_cd[NCP - 1][t];
    const uint32_t fstop_eos = eos_stop(ftotal);
    Resolved rs{E, 0, 0};
    bool active = need;
    if (active && (E >= seg_end || E >= fstop_eos)) active = false;   // owns nothing
    if (active && E == seg_start) {                      // speculation started on the true entry
        rs.exit_pos = spec_exit; rs.cnt = spec_cd & 0xFFFFu; rs.dc = spec_cd >> 16;
        active = false;
    }
    uint32_t next_cp = seg_start + CP_BITS;
    if (active && E >= seg_start + CP_BITS) {            // E itself may be a recorded block start
        const int j = (int)((E - seg_start) / CP_BITS) - 1;
        if (s_pos[j][t] == E) {
            const uint32_t at = s_cd[j][t];
            rs.exit_pos = spec_exit;
            rs.cnt = (spec_cd & 0xFFFFu) - (at & 0xFFFFu);
            rs.dc = ((spec_cd >> 16) - (at >> 16)) & 0xFFFFu;
            active = false;
        }
        next_cp = seg_start + (uint32_t)(j + 2) * CP_BITS;   // first boundary after E
    }
    Parser ps;
    if (active) ps.start(base, E, ftotal);
    int dcsum = 0;
    uint32_t cnt = 0;
    const bool parsed = active;
    uint32_t next_stop = min(next_cp, fstop_eos);
    while (__any_sync(FULL_MASK, active)) {
        if (active) {
            Parser::Sym y;
            const bool end = ps.step(ftotal, y);
            dcsum += y.dc ? y.e : 0;
            cnt += end ? 1u : 0u;
            if (end && ps.fpos >= next_stop) {           // rare: first block start past a boundary / end of stream
                const uint32_t pos = ps.fpos;
                bool done = false;
                if (pos >= next_cp) {
                    const int j = (int)min((uint32_t)NCP, (pos - seg_start) / CP_BITS) - 1;
                    if (s_pos[j][t] == pos) {            // merged with the speculative trajectory
                        const uint32_t at = s_cd[j][t];
                        cnt += (spec_cd & 0xFFFFu) - (at & 0xFFFFu);
                        dcsum += (int)(spec_cd >> 16) - (int)(at >> 16);
                        rs.exit_pos = spec_exit;
                        done = true;
                    }
                    next_cp = seg_start + (uint32_t)(j + 2) * CP_BITS;
                }
                if (!done && (pos >= seg_end || pos >= fstop_eos)) { rs.exit_pos = pos; done = true; }
                if (done) active = false;
                next_stop = min(next_cp, fstop_eos);
            }
        }
    }
    if (parsed) { rs.cnt = cnt; rs.dc = (uint32_t)dcsum & 0xFFFFu; }
    return rs;
}

__global__ void __launch_bounds__(ENT_TPB)
k_entropy_sync(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const TileDesc* __restrict__ tiles, uint32_t* __restrict__ seg_entry,
               uint32_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_cd) {
    __shared__ uint32_t s_pos[NCP][ENT_TPB];   // [j] = first block start >= seg_start + (j+1)*CP_BITS (f position)
    __shared__ uint32_t s_cd[NCP][ENT_TPB];    // blocks started before it | DC sum of them << 16
    __shared__ uint32_t s_rexit[ENT_TPB];      // resolved exits (round 1)
    const int t = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const StreamDesc sd = streams[td.stream];
    const int seg = (int)td.seg0 - 1 + t;
    const bool valid = seg >= 0 && seg < (int)sd.nseg;
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t bias = stream_bias(base);
    const uint32_t ftotal = sd.byte_len * 8u + bias;
    const uint32_t fstop_eos = eos_stop(ftotal);
    const uint32_t seg_start = (uint32_t)seg * SEG_BITS + bias, seg_end = seg_start + SEG_BITS;

    // ---- speculative parse from the segment's first bit --------------------------------------------
    {
        int j = 0;
        uint32_t cnt = 0;
        int dcsum = 0;
        uint32_t pos = seg_start;
        bool active = valid && pos < fstop_eos;
        Parser ps;
        if (active) ps.start(base, seg_start, ftotal);
        uint32_t next_cp = seg_start + CP_BITS;
        uint32_t next_stop = min(next_cp, fstop_eos);
        while (__any_sync(FULL_MASK, active)) {
            if (active) {
                Parser::Sym y;
                const bool end = ps.step(ftotal, y);
                dcsum += y.dc ? y.e : 0;
                cnt += end ? 1u : 0u;
                if (end && ps.fpos >= next_stop) {           // rare: a checkpoint boundary or the end of the stream passed
                    pos = ps.fpos;
                    const uint32_t rec = cnt | ((uint32_t)dcsum << 16);
                    while (j < NCP && pos >= next_cp) { s_pos[j][t] = pos; s_cd[j][t] = rec; j++; next_cp += CP_BITS; }
                    if (j == NCP || pos >= fstop_eos) active = false;   // end of stream: no further block can start
                    next_stop = min(next_cp, fstop_eos);
                }
            }
        }
        if (valid) {
            const uint32_t rec = cnt | ((uint32_t)dcsum << 16);
            for (; j < NCP; j++) { s_pos[j][t] = pos; s_cd[j][t] = rec; }
        }
    }
    __syncthreads();

    // ---- round 1: entry = predecessor's speculative exit -------------------------------------------------
    const bool own = valid && t >= 1;
    uint32_t E = (own && seg != 0) ? s_pos[NCP - 1][t - 1] : bias;
    Resolved rs = resolve_by_merge(base, E, seg_start, seg_end, ftotal, s_pos, s_cd, t, own);
    s_rexit[t] = own ? rs.exit_pos : (valid ? s_pos[NCP - 1][t] : 0u);   // the halo keeps its speculative exit
    __syncthreads();
    // ---- round 2: re-merge where the predecessor's resolved exit differs from its speculative one ----------
    {
        const uint32_t E2 = (own && seg != 0) ? s_rexit[t - 1] : E;
        const bool redo = own && E2 != E;
        const Resolved r2 = resolve_by_merge(base, E2, seg_start, seg_end, ftotal, s_pos, s_cd, t, redo);
        if (redo) { E = E2; rs = r2; }
    }
    if (own) {
        const uint32_t g = sd.seg_base + (uint32_t)seg;
        seg_entry[g] = E - bias;
        seg_exit[g] = rs.exit_pos - bias;
        seg_cd[g] = (rs.cnt & 0xFFFFu) | (rs.dc << 16);
    }
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
    const uint32_t bias = stream_bias(base);
    const uint32_t ftotal = sd.byte_len * 8u + bias;
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
                uint32_t x = parse_segment(base, E + bias, (i + 1) * SEG_BITS + bias, ftotal, cnt, dc) - bias;
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
// Block index + symbol lists.  Walking its blocks from the exact state, a segment's thread writes
//   sym[seg * SYM_STRIDE + ...]  one entry per coded AC coefficient: zig-zag index | amplitude << 16
//   blk_info[block].x            index of the block's first entry in sym[]
//   blk_info[block].y            absolute DC level (`cur` of LIB/decoder/lossless_decode.c:73,94, the int16
//                                running sum of DC deltas; for P frames the DC delta itself, :91) | entries << 16
// After this pass no kernel touches the bitstream again: the block-parallel decode kernels (decode.cu)
// read the lists with independent, look-ahead loads instead of a bit-serial dependent chain.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENT_TPB)
k_entropy_index(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ seg_entry,
                const uint32_t* __restrict__ seg_cd, const uint32_t* __restrict__ seg_first,
                uint2* __restrict__ blk_info, uint32_t* __restrict__ sym, uint32_t sym_seg0,
                unsigned long long* __restrict__ n_entries) {
    const int t = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const StreamDesc sd = streams[td.stream];
    const uint32_t seg = td.seg0 + (uint32_t)t;
    const bool valid = seg < sd.nseg;
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t bias = stream_bias(base);
    const uint32_t ftotal = sd.byte_len * 8u + bias;
    const uint32_t g = sd.seg_base + (valid ? seg : 0u);
    const uint32_t first = valid ? seg_first[g] : 0u;
    const uint32_t cd = valid ? seg_cd[g] : 0u;
    uint32_t cnt = cd & 0xFFFFu;
    cnt = first >= sd.nb ? 0u : min(cnt, sd.nb - first);          // trailing pad bits can look like blocks
    uint2* bi = blk_info + sd.block_base + first;
    const uint32_t o_base = (g - sym_seg0) * SYM_STRIDE, o_end = o_base + SYM_STRIDE;   // chunk-relative entry index
    uint32_t o = o_base, o_blk = o_base;
    bool active = valid && cnt != 0;
    Parser ps;
    if (active) ps.start(base, seg_entry[g] + bias, ftotal);
    int cur = sd.ptype ? 0 : (int)(cd >> 16);
    const bool pframe = sd.ptype != 0;
    uint32_t k = 0;
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // the last (o & 7) entries, newest in q[7]: stored one full 32-byte
                                               // sector at a time (a partial-sector store makes L2 fetch the rest)
    while (__any_sync(FULL_MASK, active)) {
        if (active) {
            Parser::Sym y;
            const bool end = ps.step(ftotal, y);
            if (y.dc) { cur = pframe ? y.e : cur + y.e; o_blk = o; }
            if (y.coded && y.at < 64u) {
#pragma unroll
                for (int i = 0; i < 7; i++) q[i] = q[i + 1];
                q[7] = y.at | ((uint32_t)y.e << 16);
                o++;
                if ((o & 7u) == 0u && o <= o_end) st_global_v8(sym + o - 8, q);   // never overflows on conforming streams
            }
            if (end) {
                bi[k] = make_uint2(o_blk, ((uint32_t)cur & 0xFFFFu) | ((min(o, o_end) - min(o_blk, o_end)) << 16));
                if (++k == cnt) active = false;
            }
        }
    }
    if (valid && (o & 7u) && o < o_end) {      // flush the partial group (entries beyond o are never read)
        const uint32_t r = o & 7u;
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {          // v[i] = q[8 - r + i] for i < r
            uint32_t x = 0;
#pragma unroll
            for (int j = 1; j < 8; j++) if ((uint32_t)j == r && 8 - j + i < 8) x = q[8 - j + i];
            v[i] = x;
        }
        st_global_v8(sym + (o & ~7u), v);
    }
    {   // statistics: list entries written by this launch (one atomic per warp)
        uint32_t mine = valid ? min(o, o_end) - o_base : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, d);
        if ((t & 31) == 0 && mine) atomicAdd(n_entries, (unsigned long long)mine);
    }
    // A stream that ends early leaves the remaining blocks empty (zero coefficients).
    if (valid && seg + 1 == sd.nseg)
        for (uint32_t b = first + cnt; b < sd.nb; b++) blk_info[sd.block_base + b] = make_uint2(0, 0);
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
cudaError_t launch_entropy_index(const EntropyJob& j, cudaStream_t s) {
    if (j.n_write_tiles == 0) return cudaSuccess;
    k_entropy_index<<<j.n_write_tiles, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_write_tiles, j.d_seg_entry,
                                                        j.d_seg_cd, j.d_seg_first, j.d_blk_info, j.d_sym, j.sym_seg0,
                                                        j.d_fixups + 1);
    return cudaGetLastError();
}

}  // namespace mj
