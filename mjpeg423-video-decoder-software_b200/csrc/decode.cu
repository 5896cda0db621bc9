// decode.cu -- block-parallel final decode (sm_100a): one thread per 8x8 block.
//
// After entropy.cu has produced the symbol lists and the block index (list position, entry count and
// absolute DC level of every block), every block of every plane can be decoded independently and
// without any bit-serial work:
//
//   k_decode_coef   lossless_decode() output, LIB/decoder/lossless_decode.c:60-135: the thread scatters its
//                   block's entries into a 128-byte shared-memory slot (zeroed = the memset of :77-78, or preloaded
//                   with the previous frame's coefficients for P frames, :90-92,121-123), dequantising
//                   as it scatters zig-zag -> natural order (:122-126); the CTA then stores its 128
//                   consecutive blocks as one contiguous, fully coalesced 16 KB run.
//   k_decode_fused  the whole reference loop body, LIB/decoder/mjpeg423_decoder.c:110-124 (<false>: intra-only
//                   ranges; <true>: ranges with P frames, a GOP at a time): a warp takes 32 block positions, reads the lists of their Y, Cb and Cr blocks with
//                   coalesced loads and scatters them into the owners' shared-memory slots, every thread then runs
//                   the three IDCTs of its position through shared memory (idct.c:22-181) and writes the 8x8
//                   BGRA pixels (ycbcr_to_rgb.c:26-49).  Coefficients and samples never touch HBM:
//                   the kernel reads the lists (4 bytes per coded coefficient + 8 per block) and writes
//                   4 bytes per pixel.
// (LIB = /root/reference/core0/software/common/libs/mjpeg423.)
//
// Slots use the XOR swizzle of idct_colour.cu (16-byte chunk r of slot t at chunk r ^ (t & 7)), so both
// the row reads of the IDCT and the cooperative copy-out are bank-conflict free.
#include <mutex>

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

// Block index entry -> .y with the DC level made absolute: adds the DC predictor entering the block's segment
// (k_entropy_dcscan; zero for P frames, whose DC symbols are deltas against the previous frame).
__device__ __forceinline__ uint32_t absolute_dc(uint2 info, const uint32_t* __restrict__ seg_dc) {
    const uint32_t pred = info.x == BLK_NO_SEG ? 0u : __ldg(seg_dc + info.x / SYM_STRIDE);
    return (info.y & 0xFFFF0000u) | ((info.y + pred) & 0xFFFFu);
}

// ---- coefficient planes ------------------------------------------------------------------------------
// grid = (ceil(nb / 128), number of streams in `stream_ids`).
__global__ void __launch_bounds__(DEC_TPB)
k_decode_coef(const StreamDesc* __restrict__ streams, const uint32_t* __restrict__ stream_ids, uint32_t groups,
              const uint2* __restrict__ blk_info, const uint32_t* __restrict__ sym, const uint32_t* __restrict__ seg_dc,
              const int16_t* __restrict__ quant, int16_t* coef) {
    __shared__ __align__(128) uint8_t slots[DEC_TPB * 128];
    __shared__ uint32_t s_zq[64];
    const int t = threadIdx.x;
    const uint32_t id = blockIdx.x / groups, grp = blockIdx.x - id * groups;   // (gridDim.y is capped at 65535)
    const StreamDesc sd = streams[stream_ids[id]];
    const uint32_t b0 = grp * DEC_TPB;
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
        const uint2 info = __ldg(blk_info + gb);
        const uint32_t meta = absolute_dc(info, seg_dc);
        const uint32_t* src = sym + info.x;
        const uint32_t n = meta >> 16;
        const int dc = (int)(int16_t)(meta & 0xFFFFu);
        int16_t* d0 = reinterpret_cast<int16_t*>(slots + slot_off(t, 0));
        const int q0 = (int)(s_zq[0] >> 16);
        if (sd.ptype) *d0 = (int16_t)(*d0 + dc * q0);                    // lossless_decode.c:91 (dc = the delta)
        else *d0 = (int16_t)(dc * q0);                                   // :94-95 (dc = running sum `cur`)
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t ent = __ldg(src + i);
            const uint32_t z = s_zq[ent & 63u];
            int16_t* d = reinterpret_cast<int16_t*>(slots + slot_off(t, z & 0xFFFFu));
            const int v = ((int)ent >> 16) * (int)(z >> 16);
            if (sd.ptype) *d = (int16_t)(*d + v);                        // :122
            else *d = (int16_t)v;                                        // :125
        }
    }
    __syncthreads();
    for (uint32_t i = t; i < nblk * 8u; i += DEC_TPB) {      // 16 KB contiguous, 512 B per warp instruction
        const uint32_t blk = i >> 3, row = i & 7u;
        dst[i] = *reinterpret_cast<const uint4*>(slots + blk * 128u + ((row ^ (blk & 7u)) << 4));
    }
}

// ---- fully fused: symbol lists + block index -> BGRA ------------------------------------------------------
// PERSISTENT kernel: one CTA of FUSED_TPB threads per SM (all the shared memory of the SM), every WARP
// loops over warp tiles of 32 consecutive block positions of one frame.  Warps never synchronise with
// each other after the table set-up, so a warp that finishes a cheap tile (flat picture area) starts the
// next one at once and the SM stays at its full 18 resident warps.
//
// The body is written as LOOPS over planes, column pairs and rows with its working set in shared
// memory, not as one unrolled register-resident IDCT: warps drift apart in the data-dependent scatter,
// so the instruction working set has to fit the instruction cache (the unrolled form is ~100 KB of
// SASS and ran instruction-fetch bound, profiles/r01b).  Shared memory, in 16-byte granules interleaved
// by thread (granule g of thread t at (g*(T+1) + t)*16, T = FUSED_TPB; conflict-free for 128-bit access; the extra
// granule per row rotates the banks by 4 from row to row: without it all columns of a block's slot share their banks
// and the data-dependent scatter ran at 21 % conflicted wavefronts, with it at 10-15 %, profiles/r02i):
//   ws    granules 0..15   granule cp*4 + r/2 = pass-1 outputs {ws[r][2cp], ws[r][2cp+1], ws[r+1][2cp], ws[r+1][2cp+1]}
//   coef  granules 8..15   granule 8+c = column c of the block, rows 0..7 as int16.  ALIASES the upper half of
//                          ws: pass 1 consumes columns 2cp, 2cp+1 in iteration cp and only then writes granules
//                          4cp..4cp+3, so every coefficient granule is dead before it is overwritten.
//                          (The parser's zig-zag table scatters straight into this TRANSPOSED layout, so
//                          pass 1 reads a column with one LDS.128.)
//   stash words [32][T]    word p*16 + 2r + h = samples of plane p (Y, Cb), row r, half h
constexpr int FUSED_TPB = 576;                                   // 18 warps x 384 B/thread = 216 KB of the SM's 227 KB
constexpr int FUSED_ROW = FUSED_TPB + 1;                         // granules per workspace row: the odd granule rotates the banks by 4 from
                                                                 // row to row, so a block's entries in different COLUMNS (rows FUSED_ROW * 16
                                                                 // bytes apart) no longer meet in one bank when they are scattered
constexpr int FUSED_OFF_STASH = 16 * FUSED_ROW * 16, FUSED_OFF_ZQ = FUSED_OFF_STASH + FUSED_TPB * 128;
constexpr int FUSED_SMEM = FUSED_OFF_ZQ + 2 * 64 * 8 + 16;       // + the CTA's tile counter

// PF = true is the variant for ranges that hold P frames (LIB/decoder/lossless_decode.c:90-92,121-123: every decoded
// value is ADDED to the previous frame's coefficient).  A work item is then (tile, GOP): the warp walks the GOP's frames
// in order for one tile, so the inter-frame state never leaves the warp: the DC values and the accumulated column
// masks stay in registers, and the coefficient slots of a plane that has held an AC coefficient since the I frame are
// parked between frames in a per-warp scratch area (12 KB per warp: it stays in L2) and brought back with cp.async.
// gop_first[g] .. gop_first[g + 1] are the (chunk-relative) frames of GOP g.
template <bool PF>
__global__ void __launch_bounds__(FUSED_TPB, 1)
k_decode_fused(const uint2* __restrict__ blk_info,
               const uint32_t* __restrict__ sym, const uint32_t* __restrict__ seg_dc,
               const int16_t* __restrict__ quant, uint8_t* __restrict__ out, uint32_t nb, uint32_t wb, uint32_t W,
               uint32_t n_frames, const uint32_t* __restrict__ gop_first, uint32_t n_gops, uint4* __restrict__ state) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* s_ws = reinterpret_cast<uint4*>(smem);                                   // granules 0..15
    uint8_t* s_coef = smem + 8 * FUSED_ROW * 16;                                    // granules 8..15
    uint32_t* s_stash = reinterpret_cast<uint32_t*>(smem + FUSED_OFF_STASH);
    uint2* s_zq = reinterpret_cast<uint2*>(smem + FUSED_OFF_ZQ);                 // 2 x 64 entries
    uint32_t* s_next = reinterpret_cast<uint32_t*>(smem + FUSED_OFF_ZQ + 1024);  // tiles handed out so far
    const int t = threadIdx.x;
    if (t == 0) *s_next = 0u;
    if (t < 128) {     // zig-zag index -> .x = transposed slot offset | quant << 16, .y = column bit | (row >= 1) column bit << 8
        const int tab = t >> 6, k = t & 63;
        const uint32_t n = c_zigzag[k], col = n & 7u, row = n >> 3;
        s_zq[t] = make_uint2((col * (FUSED_ROW * 16u) + row * 2u) | ((uint32_t)(uint16_t)quant[tab * 64 + n] << 16),
                             (1u << col) | ((row ? 1u : 0u) << (8 + col)));
    }
    __syncthreads();
    uint8_t* my_coef = s_coef + t * 16;
    const uint32_t tiles_per_frame = (nb + 31u) / 32u;
    const uint32_t n_items = tiles_per_frame * (PF ? n_gops : n_frames);  // work items: (tile, frame) / (tile, GOP)

    // Block index entries (blk_info is the chunk's table: plane p of frame f starts at (f * 3 + p) * nb) are
    // fetched TWO tiles ahead and the head of every plane's lists ONE tile ahead, both right behind the scatter
    // of the luminance plane: the IDCT of that plane (the one phase every tile has) then covers their latency.
    // (ptxas tracks all global loads of this kernel with one scoreboard, so a wait for any of them waits for all
    // that are in flight: nothing may be issued shortly before a point that consumes an older load.)
    const uint32_t lane = (uint32_t)t & 31u;
    // Tiles are handed out DYNAMICALLY inside a CTA: CTA c owns tiles c, c + grid, c + 2 grid, ... and its warps take
    // the next one from a shared-memory counter.  (18 warps on 4 schedulers: with a fixed share per warp the two
    // schedulers that hold 5 warps set the kernel's duration while the other two idle for the last 15 % of it.)
    // The counter is read at the top of a tile for the tile after the next (index entries are fetched two tiles ahead).
    auto take = [&]() { return lane == 0u ? atomicAdd(s_next, 1u) : 0u; };          // raw: valid in lane 0
    struct Pos { uint32_t f, tb, fend; };         // frame (chunk-relative; n_frames = nothing left), tile of the frame, end of the item
    auto item_pos = [&](uint32_t raw) {
        const unsigned long long k = __shfl_sync(FULL_MASK, raw, 0);
        const uint32_t i = (uint32_t)min(k * gridDim.x + blockIdx.x, (unsigned long long)n_items);
        Pos q;
        q.f = q.fend = n_frames; q.tb = 0u;
        if (i < n_items) {
            const uint32_t g = i / tiles_per_frame;
            q.tb = i - g * tiles_per_frame;
            if (PF) { q.f = __ldg(gop_first + g); q.fend = __ldg(gop_first + g + 1); }
            else { q.f = g; q.fend = g + 1u; }
        }
        return q;
    };
    uint8_t* warp_coef = s_coef + ((uint32_t)t & ~31u) * 16u;            // slot of lane 0 of this warp
    auto load_info = [&](const Pos& q, uint2 (&inf)[3]) {
        const uint32_t f_ = q.f;
        const uint32_t b_ = q.tb * 32u + lane;
        const bool ok = f_ < n_frames && b_ < nb;
#pragma unroll
        for (int p = 0; p < 3; p++) {                                    // (volatile: keeps its place behind the rotation below)
            uint32_t vx = BLK_NO_SEG, vy = 0u;
            if (ok) asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "l"(blk_info + (size_t)(f_ * 3u + p) * nb + b_));
            inf[p] = make_uint2(vx, vy);
        }
    };
    // The lists of consecutive blocks of a segment are consecutive in memory, so the 32 lists of a tile form
    // one contiguous run of entries per bitstream segment the tile touches (at the usual rates one or two).
    // runs(): entry range [r0, r0 + n1) of the run that starts at lane 0, [s0, s0 + n2) of the second run
    // (n2 = 0 without one); `rest` = first lanes of the runs after these.
    auto runs = [&](uint32_t x, uint32_t xe, uint32_t& r0, uint32_t& n1, uint32_t& s0, uint32_t& n2, uint32_t& rest) {
        const uint32_t prev_e = __shfl_up_sync(FULL_MASK, xe, 1);
        rest = __ballot_sync(FULL_MASK, lane != 0u && x != prev_e);
        r0 = __shfl_sync(FULL_MASK, x, 0);
        n1 = __shfl_sync(FULL_MASK, xe, rest ? __ffs(rest) - 2 : 31) - r0;
        s0 = 0u; n2 = 0u;
        if (rest) {
            const int l2 = __ffs(rest) - 1;
            rest &= rest - 1u;
            s0 = __shfl_sync(FULL_MASK, x, l2);
            n2 = __shfl_sync(FULL_MASK, xe, rest ? __ffs(rest) - 2 : 31) - s0;
        }
    };
    // Entries carry their owner (the lane whose block they belong to), so ANY lane may process any entry: the
    // head of a tile's entries -- slot v = v-th entry of the first run, continued in the second run -- is
    // fetched into registers one tile ahead (4 x 32 luminance slots, 32 of each chrominance plane: all of a
    // quiet tile), the remainder is staged through shared memory when the plane is decoded.
    constexpr int PRE_Y = 4;
    uint2 infoA[3], infoB[3];                                            // index entries of the next tile / the one after
    uint32_t npreY[PRE_Y], npreC[2], npred[3];
    auto prefetch_lists = [&]() {
#pragma unroll
        for (int p = 0; p < 3; p++) {
            npred[p] = infoA[p].x == BLK_NO_SEG ? 0u : __ldg(seg_dc + infoA[p].x / SYM_STRIDE);   // DC predictor of the block's segment
            if (p == 0) {
#pragma unroll
                for (int i = 0; i < PRE_Y; i++) npreY[i] = 0u;
            } else {
                npreC[p - 1] = 0u;
            }
            if (!__any_sync(FULL_MASK, (infoA[p].y >> 16) != 0u)) continue;   // a plane without any entry (flat chrominance)
            uint32_t r0, n1, s0, n2, rest;
            runs(infoA[p].x, infoA[p].x + (infoA[p].y >> 16), r0, n1, s0, n2, rest);
            const uint32_t d2 = s0 - n1 - r0;                            // slot v of the second run is entry r0 + d2 + v
            if (p == 0) {
#pragma unroll
                for (int i = 0; i < PRE_Y; i++) {
                    const uint32_t v = lane + 32u * i;
                    if (v < n1 + n2) npreY[i] = __ldg(sym + (r0 + v + (v < n1 ? 0u : d2)));
                }
            } else {
                if (lane < n1 + n2) npreC[p - 1] = __ldg(sym + (r0 + lane + (lane < n1 ? 0u : d2)));
            }
            {   // what does not fit the registers (busy tiles) is pulled into L2, one 128-byte line per lane and run:
                // the staging copy of the next tile then pays an L2 hit instead of a DRAM access
                const uint32_t pre = p == 0 ? 32u * PRE_Y : 32u;
                const uint32_t u1 = min(n1, pre), u2 = min(pre - u1, n2);
                const uint32_t q1 = r0 + u1 + lane * 32u, q2 = s0 + u2 + lane * 32u;
                if (q1 < r0 + n1) asm volatile("prefetch.global.L2 [%0];" ::"l"(sym + q1));
                if (q2 < s0 + n2) asm volatile("prefetch.global.L2 [%0];" ::"l"(sym + q2));
            }
        }
    };
    // Staging area of the warp for list remainders: the lower half of its workspace granules (dead while a plane
    // is scattered: pass 1 writes it, pass 2 reads it), 8 pieces of 512 bytes = 1024 entries.  Entry k = lane + 32 j
    // is copied (cp.async, no register, no scoreboard) and read back by the SAME lane.
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + ((uint32_t)t & ~31u) * 16u + lane * 4u;
    // Positions: `cur` is decoded, `nxt` follows it (its index entries are in flight), `nn` follows that.  Inside an
    // item (PF) the successor is the next frame of the GOP, otherwise the first frame of a newly taken item.
    Pos cur = item_pos(take()), nxt, nn;
    bool cur_first = true, nxt_first, nn_first = true;    // first frame of its item (an I frame: no state before it)
    nxt_first = !(PF && cur.f + 1u < cur.fend);
    if (nxt_first) nxt = item_pos(take());
    else { nxt.f = cur.f + 1u; nxt.tb = cur.tb; nxt.fend = cur.fend; }
    nn = nxt;
    load_info(cur, infoA);
    prefetch_lists();
    load_info(nxt, infoB);
    // PF: inter-frame state of the item -- DC coefficients (per lane) and accumulated masks (warp-uniform) of the planes
    int dcY = 0, dcB = 0, dcR = 0;
    uint32_t mY = 1u, mB = 1u, mR = 1u;
    uint4* my_state = PF ? state + ((size_t)(blockIdx.x * (FUSED_TPB / 32) + (t >> 5)) * 3u * 8u) * 32u + lane : nullptr;

    // The pipeline step: index entries of the next position become current, its list heads are requested, the
    // index entries of the position after it are requested.
    auto advance = [&](uint32_t raw2) {
        // (The rotation is spelled as opaque moves in front of the loads: otherwise the loads land in temporaries
        // that are copied into the loop-carried registers at once, i.e. waited for right here.)
#pragma unroll
        for (int q = 0; q < 3; q++) {
            asm volatile("mov.b32 %0, %1;" : "=r"(infoA[q].x) : "r"(infoB[q].x));
            asm volatile("mov.b32 %0, %1;" : "=r"(infoA[q].y) : "r"(infoB[q].y));
        }
        prefetch_lists();
        if (nn_first) nn = item_pos(raw2);
        else { nn.f = nxt.f + 1u; nn.tb = nxt.tb; nn.fend = nxt.fend; }
        load_info(nn, infoB);
    };

    for (; cur.f < n_frames; cur = nxt, nxt = nn, cur_first = nxt_first, nxt_first = nn_first) {
        nn_first = !(PF && nxt.f + 1u < nxt.fend);
        const uint32_t raw2 = nn_first ? take() : 0u;                     // the item after the next position's
        const uint32_t f = cur.f;
        const uint32_t b = cur.tb * 32u + lane;
        const bool live = b < nb;
        uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;

        uint32_t meta[3], lx[3], lxe[3], preY[PRE_Y], preC[2];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            lx[p] = infoA[p].x; lxe[p] = infoA[p].x + (infoA[p].y >> 16);
            meta[p] = (infoA[p].y & 0xFFFF0000u) | ((infoA[p].y + npred[p]) & 0xFFFFu);          // absolute DC level
        }
#pragma unroll
        for (int i = 0; i < PRE_Y; i++) preY[i] = npreY[i];
        preC[0] = npreC[0]; preC[1] = npreC[1];

        if (PF && !cur_first) {
            // P frame whose 32 positions are all unchanged (no entry, zero DC deltas, in all three planes -- what a
            // static background is coded as): the pixels are the previous frame's, which this lane wrote itself.
            bool same = true;
#pragma unroll
            for (int p = 0; p < 3; p++) same = same && lx[p] == lxe[p] && (meta[p] & 0xFFFFu) == 0u;
            if (__all_sync(FULL_MASK, same)) {
                if (live) {
                    const uint8_t* prev = dst - (size_t)nb * 256;
#pragma unroll 4
                    for (int r = 0; r < 8; r++) {
                        uint32_t v[8];
                        ld_global_v8(prev + (size_t)r * W * 4, v);
                        st_global_v8(dst + (size_t)r * W * 4, v);
                    }
                }
                advance(raw2);
                continue;
            }
        }
        bool cb_flat = false;                                            // warp-uniform: every Cb block of the tile is DC-only;
        uint32_t cb_s8 = 0;                                              // its sample then waits here, not in the stash
#pragma unroll 1
        for (int p = 0; p < 3; p++) {
            // (p is a loop variable: select the pre-fetched registers without dynamic indexing)
            const uint32_t pmeta = p == 0 ? meta[0] : p == 1 ? meta[1] : meta[2];
            const uint32_t x = p == 0 ? lx[0] : p == 1 ? lx[1] : lx[2];
            const uint32_t xe = p == 0 ? lxe[0] : p == 1 ? lxe[1] : lxe[2];
            const uint2* zq = s_zq + (p ? 64 : 0);
            int dc_coef = (int)(int16_t)((int)(int16_t)(pmeta & 0xFFFFu) * (int)(zq[0].x >> 16));   // lossless_decode.c:94-95
            uint32_t macc = 1u;                                           // PF: columns occupied since the item's I frame
            if (PF) {
                if (!cur_first) {                                         // P frame: the level is a delta against the previous frame (:91)
                    dc_coef = (int)(int16_t)(dc_coef + (p == 0 ? dcY : p == 1 ? dcB : dcR));
                    macc = p == 0 ? mY : p == 1 ? mB : mR;
                }
                if (p == 0) dcY = dc_coef; else if (p == 1) dcB = dc_coef; else dcR = dc_coef;
            }
            const bool has_state = PF && (macc & 0xFFFEu) != 0u;          // the plane's slots are parked in the scratch area

            // emit(r, w0, w1): Y and Cb rows go to the stash, a Cr row completes 8 pixels.
            auto emit = [&](int r, uint32_t w0, uint32_t w1) {
                if (p < 2) {
                    s_stash[(p * 16 + 2 * r) * FUSED_TPB + t] = w0;
                    s_stash[(p * 16 + 2 * r + 1) * FUSED_TPB + t] = w1;
                } else if (live) {
                    colour_row_store(s_stash[(2 * r) * FUSED_TPB + t], s_stash[(2 * r + 1) * FUSED_TPB + t],
                                     s_stash[(16 + 2 * r) * FUSED_TPB + t], s_stash[(17 + 2 * r) * FUSED_TPB + t], w0, w1,
                                     dst + (size_t)r * W * 4);
                }
            };
            // A plane whose 32 blocks are all DC-only: both passes collapse to (4*dc + 16) >> 5 (see idct_block()).
            auto dc_only_plane = [&]() {
                const uint32_t s8 = clamp255(((dc_coef << 2) + 16) >> 5);
                const uint32_t v = s8 * 0x01010101u;
                if (p == 1) { cb_flat = true; cb_s8 = s8; return; }       // kept in a register until Cr is known
                if (p == 2 && cb_flat) {
                    // Flat Cb and Cr blocks: the chroma terms are per-block constants.
                    if (live) {
                        const FlatChroma fc(cb_s8, s8);
#pragma unroll 2
                        for (int r = 0; r < 8; r++)
                            fc.row_store(s_stash[(2 * r) * FUSED_TPB + t], s_stash[(2 * r + 1) * FUSED_TPB + t],
                                         dst + (size_t)r * W * 4);
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < 8; r++) emit(r, v, v);
                }
            };
            uint32_t m_all = macc;                                        // column 0 always holds the DC coefficient
            const bool has_ac = __any_sync(FULL_MASK, xe != x);
            if (has_ac || has_state) {                                    // (no AC entry in the whole tile: nothing to scatter)
                // ---- scatter this plane's blocks into the transposed coefficient slots: zeroed (the memset of :77-78)
                // or, for a P frame, holding the previous frame's coefficients --------------------------------------
                if (has_state) {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(my_coef);
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + c * (FUSED_ROW * 16)),
                                     "l"(my_state + (p * 8 + c) * 32) : "memory");
                    asm volatile("cp.async.wait_all;" ::: "memory");
                } else {
#pragma unroll
                    for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(my_coef + c * (FUSED_ROW * 16)) = make_uint4(0, 0, 0, 0);
                }
                *reinterpret_cast<int16_t*>(my_coef) = (int16_t)dc_coef;
                __syncwarp();
                // Whichever lane holds an entry dequantises it and stores it into the slot of the lane that owns
                // the block (entry bits 6..10, written by k_entropy_index): no lane waits for the longest list of
                // the tile, no load depends on another.
                uint32_t m_bits = 1u;
                auto put = [&](uint32_t ent) {                            // dequantise + scatter one entry (:125)
                    const uint2 z = zq[ent & 63u];
                    int16_t* d = reinterpret_cast<int16_t*>(warp_coef + ((ent >> 2) & 0x1F0u) + (z.x & 0xFFFFu));
                    const int v = ((int)ent >> 16) * (int)(z.x >> 16);
                    if (PF) *d = (int16_t)(v + (cur_first ? 0 : (int)*d));   // :122 (P frame: added) / :125 (I frame: stored)
                    else *d = (int16_t)v;
                    m_bits |= z.y;
                };
                uint32_t r0, n1, s0, n2, rest;
                runs(x, xe, r0, n1, s0, n2, rest);
                uint32_t pre;                                             // slots already in registers
                if (p == 0) {
#pragma unroll
                    for (int i = 0; i < PRE_Y; i++) if (lane + 32u * i < n1 + n2) put(preY[i]);
                    pre = 32u * PRE_Y;
                } else {
                    if (lane < n1 + n2) put(p == 1 ? preC[0] : preC[1]);
                    pre = 32u;
                }
                // Remainders (warp-uniform): of the first run, of the second run, then whole further runs.
                const uint32_t u1 = min(n1, pre), u2 = min(pre - u1, n2);
                uint32_t a = r0 + u1, e = r0 + n1, a2 = s0 + u2, e2 = s0 + n2;
                for (;;) {
                    while (a < e) {                                       // stage up to 1024 entries, then scatter them
                        const uint32_t n = min(e - a, 1024u);
                        const uint32_t* src = sym + a + lane;
                        const uint32_t ng = (n + 127u) >> 7;              // pieces of 128 entries
                        for (uint32_t g = 0; g < ng; g++) {
                            const uint32_t sa = stage0 + g * (FUSED_ROW * 16u), k = g * 128u + lane;
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if (k + 32u * i < n)
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + 128u * i), "l"(src + g * 128u + 32u * i) : "memory");
                        }
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        for (uint32_t g = 0; g < ng; g++) {
                            const uint32_t sa = stage0 + g * (FUSED_ROW * 16u), k = g * 128u + lane;
                            uint32_t ent[4];
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if (k + 32u * i < n) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ent[i]) : "r"(sa + 128u * i) : "memory");
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if (k + 32u * i < n) put(ent[i]);
                        }
                        a += n;
                    }
                    if (a2 < e2) { a = a2; e = e2; a2 = e2; continue; }
                    if (!rest) break;
                    const int l0 = __ffs(rest) - 1;
                    rest &= rest - 1u;
                    a = __shfl_sync(FULL_MASK, x, l0);
                    e = __shfl_sync(FULL_MASK, xe, rest ? __ffs(rest) - 2 : 31);
                }
                __syncwarp();
                m_all = warp_or(m_bits);                                  // warp-uniform from here on
                if (PF) {
                    m_all |= macc;
                    if (has_ac && cur.f + 1u < cur.fend && (m_all & 0xFFFEu)) {   // park the slots for the item's next frame
#pragma unroll
                        for (int c = 0; c < 8; c++)
                            my_state[(p * 8 + c) * 32] = *reinterpret_cast<const uint4*>(my_coef + c * (FUSED_ROW * 16));
                    }
                }
            }
            if (PF) { if (p == 0) mY = m_all; else if (p == 1) mB = m_all; else mR = m_all; }
            if (p == 0) advance(raw2);                                    // behind the luminance scatter: the pipeline advances
            const uint32_t acm = m_all >> 8, anym = m_all & 0xFFu;

            if (((anym & 0xFEu) | (acm & 1u)) == 0) { dc_only_plane(); continue; }   // nothing outside the DC position
            if (p == 2 && cb_flat) {                                      // Cr needs the IDCT after all: materialise the flat Cb rows
                const uint32_t v = cb_s8 * 0x01010101u;
#pragma unroll 1
                for (int r = 0; r < 16; r++) s_stash[(16 + r) * FUSED_TPB + t] = v;
                cb_flat = false;
            }
            // ---- pass 1: columns, two at a time (idct.c:41-109) --------------------------------------------------
            const bool high_half = (anym & 0xF0u) != 0;
            const int npair = high_half ? 4 : 2;
#pragma unroll 2                 // (two pairs = four butterflies in flight; within a pair of iterations no store hits a later load)
            for (int cp = 0; cp < npair; cp++) {
                const uint4 c0 = *reinterpret_cast<const uint4*>(my_coef + (2 * cp) * (FUSED_ROW * 16));
                const uint4 c1 = *reinterpret_cast<const uint4*>(my_coef + (2 * cp + 1) * (FUSED_ROW * 16));
                int o0[8], o1[8];
                if ((acm >> (2 * cp)) & 3u) {
                    idct8<11>(lo16(c0.x), hi16(c0.x), lo16(c0.y), hi16(c0.y), lo16(c0.z), hi16(c0.z), lo16(c0.w), hi16(c0.w), o0);
                    idct8<11>(lo16(c1.x), hi16(c1.x), lo16(c1.y), hi16(c1.y), lo16(c1.z), hi16(c1.z), lo16(c1.w), hi16(c1.w), o1);
                } else {                                        // no AC in either column: DESCALE(in0 << 13, 11) == in0 << 2
                    const int v0 = (int)((unsigned)lo16(c0.x) << 2), v1 = (int)((unsigned)lo16(c1.x) << 2);
#pragma unroll
                    for (int r = 0; r < 8; r++) { o0[r] = v0; o1[r] = v1; }
                }
#pragma unroll
                for (int r = 0; r < 8; r += 2)                   // both column loads above precede these stores (aliasing)
                    s_ws[(cp * 4 + r / 2) * FUSED_ROW + t] = make_uint4((uint32_t)o0[r], (uint32_t)o1[r], (uint32_t)o0[r + 1], (uint32_t)o1[r + 1]);
            }
            // ---- pass 2: rows (idct.c:116-180) ------------------------------------------------------------------
#pragma unroll 1
            for (int r = 0; r < 8; r += 2) {                          // two rows per iteration: one LDS.128 per column pair
                const uint4* g = s_ws + (r >> 1) * FUSED_ROW + t;      // and two independent butterflies in flight
                const uint4 a = g[0], bq = g[4 * FUSED_ROW];
                int o0[8], o1[8];
                if (high_half) {
                    const uint4 cq = g[8 * FUSED_ROW], dq = g[12 * FUSED_ROW];
                    idct8<18>((int)a.x, (int)a.y, (int)bq.x, (int)bq.y, (int)cq.x, (int)cq.y, (int)dq.x, (int)dq.y, o0);
                    idct8<18>((int)a.z, (int)a.w, (int)bq.z, (int)bq.w, (int)cq.z, (int)cq.w, (int)dq.z, (int)dq.w, o1);
                } else {
                    idct8<18>((int)a.x, (int)a.y, (int)bq.x, (int)bq.y, 0, 0, 0, 0, o0);
                    idct8<18>((int)a.z, (int)a.w, (int)bq.z, (int)bq.w, 0, 0, 0, 0, o1);
                }
                emit(r, pack4_sat_u8(o0[0], o0[1], o0[2], o0[3]), pack4_sat_u8(o0[4], o0[5], o0[6], o0[7]));
                emit(r + 1, pack4_sat_u8(o1[0], o1[1], o1[2], o1[3]), pack4_sat_u8(o1[4], o1[5], o1[6], o1[7]));
            }
        }
    }
}

// ---- launchers -----------------------------------------------------------------------------------------------
cudaError_t launch_decode_coef(const EntropyJob& j, const uint32_t* d_stream_ids, uint32_t n_ids, uint32_t nb,
                               const int16_t* d_quant, int16_t* d_coef, cudaStream_t s) {
    if (n_ids == 0 || nb == 0) return cudaSuccess;
    const uint32_t groups = (nb + DEC_TPB - 1) / DEC_TPB;
    if ((uint64_t)groups * n_ids > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    k_decode_coef<<<groups * n_ids, DEC_TPB, 0, s>>>(j.d_streams, d_stream_ids, groups, j.d_blk_info, j.d_sym, j.d_seg_dc + j.sym_seg0,
                                           d_quant, d_coef);
    return cudaGetLastError();
}
// gop_first == nullptr: an intra-only range (one work item per tile and frame).  Otherwise the range holds P frames:
// d_gop_first[0 .. n_gops] are the chunk-relative first frames of its GOPs (+ the end), d_state the per-warp scratch of
// FUSED_STATE_BYTES (see k_decode_fused<true>).
cudaError_t launch_decode_fused(const EntropyJob& j, const int16_t* d_quant, void* d_out, uint32_t n_frames,
                                uint32_t W, uint32_t H, const uint32_t* d_gop_first, uint32_t n_gops, void* d_state,
                                cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    // per-device set-up (dynamic shared memory opt-in, SM count): once per device, thread-safe
    static std::once_flag once[64];
    static int n_sm[64];
    static cudaError_t status[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::call_once(once[dev], [dev]() {
        cudaError_t e2 = cudaFuncSetAttribute(k_decode_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_decode_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
        if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e2 == cudaSuccess && n_sm[dev] > FUSED_MAX_CTAS) n_sm[dev] = FUSED_MAX_CTAS;   // the scratch area is sized for this many
        status[dev] = e2;
    });
    if (status[dev] != cudaSuccess) return status[dev];
    const uint32_t wb = W / 8, nb = wb * (H / 8);
    const uint64_t n_items = (uint64_t)((nb + 31) / 32) * (d_gop_first ? n_gops : n_frames);
    const uint64_t want = (n_items + FUSED_TPB / 32 - 1) / (FUSED_TPB / 32);
    const unsigned grid = (unsigned)(want < (uint64_t)n_sm[dev] ? want : (uint64_t)n_sm[dev]);   // persistent: one CTA per SM
    const uint2* bi = j.d_blk_info + (size_t)j.stream_lo * nb;
    if (d_gop_first)
        k_decode_fused<true><<<grid, FUSED_TPB, FUSED_SMEM, s>>>(bi, j.d_sym, j.d_seg_dc + j.sym_seg0, d_quant, (uint8_t*)d_out, nb, wb,
                                                                 W, n_frames, d_gop_first, n_gops, (uint4*)d_state);
    else
        k_decode_fused<false><<<grid, FUSED_TPB, FUSED_SMEM, s>>>(bi, j.d_sym, j.d_seg_dc + j.sym_seg0, d_quant, (uint8_t*)d_out, nb, wb,
                                                                  W, n_frames, nullptr, 0u, nullptr);
    return cudaGetLastError();
}

}  // namespace mj
