// entropy.cu -- batched, segment-parallel "lossless" decode (sm_100a): synchronisation, chain, block index.
//
// Replaces the serial walk of lossless_decode(), LIB/decoder/lossless_decode.c:60-135 (LIB =
// /root/reference/core0/software/common/libs/mjpeg423), for MANY plane streams at once and, inside a
// stream, for many fixed-size bitstream segments in parallel.  The code has no markers or restart
// intervals, so segment entry points are found by self-synchronisation:
//
//   k_entropy_sync   a lane parses a super-segment (SUPER consecutive segments of one stream) speculatively
//                    from its first bit as if a block started there, leaving checkpoints (first block start
//                    at or after every CP_BITS boundary of the first segment and every segment boundary,
//                    blocks so far).  Then the predecessor's exit is taken as the lane's entry and parsed only
//                    until it MERGES with the recorded trajectory (same bit position at a block start =>
//                    identical future); a second round re-merges the lanes whose predecessor's exit moved.
//   k_entropy_chain  one CTA per stream: re-parses the few segments whose entry still differs from the
//                    predecessor's resolved exit until the chain entry[i] == exit[i-1] holds from
//                    entry[0] = 0 (correctness never rests on self-synchronisation, only speed does),
//                    then exclusive-scans the block counts to give every segment its first block index.
//   k_entropy_index  walks every segment's blocks from the now exact state and writes every coded
//                    coefficient as a 32-bit entry into the segment's symbol list, plus per block
//                    the list position, entry count and segment-relative DC level: what the block-parallel
//                    decode kernels (decode.cu) consume without touching the bitstream again.
//   k_entropy_dcscan exclusive scan (mod 2^16) of the segments' DC totals: the DC
//                    predictor entering every segment.
//
// Every pass advances with Parser::step() (common.cuh) in a UNIFORM loop: one step per iteration for
// every lane, DC/AC and block-end handling predicated, the rare events (checkpoint, end of a job) in a
// short divergent branch; a lane without work is parked.  Load balance comes from many small CTAs.
#include "common.cuh"
#include "runtime.h"

namespace mj {

// All positions inside the kernels are "f" positions (stream bit position + stream_bias(), see Parser);
// the per-segment arrays in global memory hold plain stream bit positions.
//
// DC levels are NOT tracked by the synchronisation passes (their symbol loop skips amplitudes altogether):
// the index pass records every block's DC level relative to its segment's first block plus the segment's
// DC total, k_entropy_dcscan turns the totals into the predictor entering each segment, and the decode
// kernels add it (decode.cu: absolute_dc() / the prefetched predictors of k_decode_fused).

// fstop_eos: a block start at or after this f position cannot hold a block any more (fewer than
// MIN_BLOCK_BITS left): the stream's trailing pad bits.
__device__ __forceinline__ uint32_t eos_stop(uint32_t ftotal) {
    return ftotal >= (uint32_t)MIN_BLOCK_BITS ? ftotal - (MIN_BLOCK_BITS - 1) : 0u;
}

// Parse from block start `entry` to the first block start at or after seg_end (or the end of the
// stream).  Returns the exit position; cnt receives the blocks started.  (f positions.)
__device__ __forceinline__ uint32_t parse_segment(const uint8_t* base, uint32_t entry, uint32_t seg_end,
                                                  uint32_t ftotal, uint32_t& cnt) {
    cnt = 0;
    uint32_t pos = entry;
    const uint32_t stop = min(seg_end, eos_stop(ftotal));
    if (pos < stop) {
        Parser ps;
        ps.start(base, entry, seg_end, ftotal);
        for (;;) {
            Parser::Sym y;
            if (ps.step<false>(y)) {
                cnt++;
                if (ps.fpos >= stop) break;
            }
        }
        pos = ps.fpos;
    }
    return pos;
}

// Number of leading all-zero 16-byte units at u (at most maxu are looked at).
__device__ __forceinline__ uint32_t zero_units(const uint4* u, uint32_t maxu) {
    for (uint32_t z = 0; z < maxu; z += 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = z + j < maxu ? __ldg(u + z + j) : make_uint4(1u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((v[j].x | v[j].y | v[j].z | v[j].w) != 0u) return z + j;
    }
    return maxu;
}

// parse_segment() with a shortcut for all-zero stretches: twelve zero bits are a whole block (DC size 0 + END,
// what an unchanged block of a P frame is coded as), and a stream of them never self-synchronises -- every bit
// position looks like a block start -- so on static pictures the chain kernel is what finds the block phase.  Blocks
// that lie completely inside a run of zero 16-byte units are counted arithmetically (same positions and counts as
// stepping through them); the parser takes over where the zeros end.
__device__ __forceinline__ uint32_t parse_segment_z(const uint8_t* base, uint32_t entry, uint32_t seg_end,
                                                    uint32_t ftotal, uint32_t& cnt) {
    const uint32_t stop = min(seg_end, eos_stop(ftotal));
    uint32_t pos = entry, k = 0;
    if (pos < stop) {
        const uintptr_t wbase = reinterpret_cast<uintptr_t>(base) & ~(uintptr_t)3;      // f position 0
        const uintptr_t a0 = (wbase + (pos >> 3)) & ~(uintptr_t)15;                     // unit that holds the entry bit
        const uintptr_t a1 = wbase + ((min(seg_end, ftotal) + 7u) >> 3);                // scan no further than this
        if (a1 > a0) {
            const uint32_t z = zero_units(reinterpret_cast<const uint4*>(a0), (uint32_t)((a1 - a0 + 15) >> 4));
            const uint64_t zend = (uint64_t)(a0 + 16u * (uintptr_t)z - wbase) * 8u;     // f position where the zeros end
            if (zend >= (uint64_t)pos + MIN_BLOCK_BITS) {
                k = min((stop - pos + MIN_BLOCK_BITS - 1u) / MIN_BLOCK_BITS, (uint32_t)((zend - pos) / MIN_BLOCK_BITS));
                pos += k * MIN_BLOCK_BITS;
            }
        }
    }
    uint32_t c1 = 0;
    const uint32_t x = parse_segment(base, pos, seg_end, ftotal, c1);
    cnt = k + c1;
    return x;
}

// ------------------------------------------------------------------------------------------------
// Speculative parse + merge.  A LANE parses a SUPER-segment: SUPER consecutive segments of one stream
// (every stream's global segment range is padded to a multiple of SUPER, so super u = global segments
// [u*SUPER, (u+1)*SUPER)).  Only the first few hundred bits of a lane's parse are ever on the wrong
// trajectory, so the longer the lane's run, the smaller the share of re-parsing: with SUPER = 4 the
// merge phase costs ~8 % of the parse instead of ~30 % with one segment per lane.
//
// CTA c handles SYNC_OWN consecutive supers + the one BEFORE them (slot 0, halo), so that slot 1 has a
// predecessor exit without any inter-CTA dependency.
//
// Phase A: every lane parses its super from its first bit as if a block started there, recording the first
// block start at or after NCK checkpoint boundaries -- every CP_BITS inside the first segment (where
// merges happen), then every segment boundary -- with the number of blocks started before it.
// Phase B: the predecessor's speculative exit is the lane's true entry E; parse from E until the
// trajectory meets the recorded one (equal position at a block start => identical future; tested at the
// first block start past each boundary -- once merged, that IS the checkpoint).  The segment boundaries
// passed before the merge are recorded from the true trajectory, the later ones follow from the
// checkpoints.  Round 2 repeats this for lanes whose predecessor's exit moved in round 1.
// ------------------------------------------------------------------------------------------------
constexpr int NCK = NCP + SUPER - 1;                    // checkpoints of a super (the last NCK - NCP + 1 are segment boundaries)
constexpr int SYNC_OWN = ENT_TPB - 1;
constexpr uint32_t SUPER_BITS = SUPER * SEG_BITS;
constexpr uint32_t NO_WORK = 0xFFFFFFFFu;

// Offset of checkpoint boundary i from the super's first bit, and the checkpoint index of segment boundary j (1..SUPER).
__device__ __forceinline__ uint32_t ck_bnd(uint32_t i) { return i < (uint32_t)NCP ? (i + 1u) * CP_BITS : (i - (NCP - 2u)) * SEG_BITS; }
__device__ __forceinline__ uint32_t ck_of_seg(uint32_t j) { return (uint32_t)NCP - 2u + j; }
// Index of the last checkpoint boundary at or before offset `rel` (rel >= CP_BITS).
__device__ __forceinline__ uint32_t ck_last(uint32_t rel) {
    return rel < (uint32_t)SEG_BITS ? rel / CP_BITS - 1u : min((uint32_t)NCK - 1u, rel / SEG_BITS + (NCP - 2u));
}

struct SyncShared {
    uint32_t cp[NCK][ENT_TPB];         // (first block start >= boundary i, minus the super's first bit) << 16 | blocks before it
    uint32_t tpos[SUPER + 1][ENT_TPB]; // resolved: first block start >= segment boundary j (tpos[0] = the entry E)
    uint32_t tcnt[SUPER + 1][ENT_TPB]; // resolved: blocks started before it, counted from E
};

// Where a global segment lives.
struct SegCtx {
    const uint8_t* base;           // first byte of its stream
    uint32_t bias, ftotal;         // stream_bias(base), f position of the end of the stream
    uint32_t seg;                  // index of the segment inside its stream
    uint32_t seg_start;            // f position of its first bit
    uint32_t sid;                  // stream (index into the StreamDesc table)
};
__device__ __forceinline__ SegCtx seg_ctx(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                                          const uint32_t* __restrict__ seg_stream, uint32_t g) {
    SegCtx c;
    c.sid = __ldg(seg_stream + g);
    const StreamDesc* sd = streams + c.sid;
    c.base = payload + sd->byte_off;
    c.bias = stream_bias(c.base);
    c.ftotal = sd->byte_len * 8u + c.bias;
    c.seg = g - sd->seg_base;
    c.seg_start = c.seg * SEG_BITS + c.bias;
    return c;
}

// Resolve slot t for entry E.  WARP-COLLECTIVE (lanes with nothing to resolve pass need = false): the
// symbol loop is uniform, lanes without a merge to do are parked.  s0 = f position of the super's first bit.
template <bool FOLD>
__device__ __forceinline__ void resolve_super(SyncShared& sh, const uint8_t* base, uint32_t E, uint32_t s0, uint32_t ftotal,
                                              int t, bool need) {
    const uint32_t fstop_eos = eos_stop(ftotal);
    uint32_t jb = 1;                                   // next segment boundary to resolve
    uint32_t cnt = 0, inext = 0, next_stop = NO_WORK;
    Parser ps;
    ps.init_parked();
    // fill boundaries jb.. from checkpoint i (merged there with `cnt` blocks counted from E so far)
    auto finish_from = [&](uint32_t i, uint32_t cnt_here) {
        const uint32_t ci = sh.cp[i][t];
        for (; jb <= (uint32_t)SUPER; jb++) {
            const uint32_t c = sh.cp[ck_of_seg(jb)][t];
            sh.tpos[jb][t] = s0 + (c >> 16);
            sh.tcnt[jb][t] = cnt_here + (c & 0xFFFFu) - (ci & 0xFFFFu);
        }
    };
    auto finish_at = [&](uint32_t pos, uint32_t cnt_here) {      // no merge: every remaining boundary sees `pos`
        for (; jb <= (uint32_t)SUPER; jb++) { sh.tpos[jb][t] = pos; sh.tcnt[jb][t] = cnt_here; }
    };
    if (need) {
        sh.tpos[0][t] = E;
        sh.tcnt[0][t] = 0;
        if (E >= s0 + SUPER_BITS || E >= fstop_eos) finish_at(E, 0);                       // owns nothing
        else if (E == s0) {                                                                 // speculation was right
            for (; jb <= (uint32_t)SUPER; jb++) {
                const uint32_t c = sh.cp[ck_of_seg(jb)][t];
                sh.tpos[jb][t] = s0 + (c >> 16); sh.tcnt[jb][t] = c & 0xFFFFu;
            }
        } else {
            // segment boundaries already behind E (a block that spans them): they see E itself
            for (; jb < (uint32_t)SUPER && E >= s0 + jb * SEG_BITS; jb++) { sh.tpos[jb][t] = E; sh.tcnt[jb][t] = 0; }
            bool merged = false;
            if (E >= s0 + CP_BITS) {                                                        // E itself may be a recorded block start
                const uint32_t i = ck_last(E - s0);
                if (s0 + (sh.cp[i][t] >> 16) == E) { finish_from(i, 0); merged = true; }
                inext = i + 1u;
            }
            if (!merged) {
                ps.start(base, E, s0 + SUPER_BITS, ftotal);
                next_stop = min(s0 + ck_bnd(inext), fstop_eos);
            }
        }
    }
    // (four steps per look at the loop condition: a lane that is done is parked, extra steps change nothing)
    auto one_step = [&]() {
        Parser::Sym y;
        const bool end = ps.step<false, FOLD>(y);
        cnt += end ? 1u : 0u;
        if (end && ps.fpos >= next_stop) {               // rare: first block start past a boundary / end of stream
            const uint32_t pos = ps.fpos;
            bool done = false;
            if (pos >= s0 + ck_bnd(inext)) {
                const uint32_t i = ck_last(pos - s0);
                // segment boundaries passed on the way (all but possibly the last are strictly before pos's checkpoint)
                for (; jb <= (uint32_t)SUPER && ck_of_seg(jb) < i; jb++) { sh.tpos[jb][t] = pos; sh.tcnt[jb][t] = cnt; }
                if (s0 + (sh.cp[i][t] >> 16) == pos) {   // merged with the speculative trajectory
                    finish_from(i, cnt);
                    done = true;
                } else if (jb <= (uint32_t)SUPER && ck_of_seg(jb) == i) { sh.tpos[jb][t] = pos; sh.tcnt[jb][t] = cnt; jb++; }
                inext = i + 1u;
                if (!done && inext == (uint32_t)NCK) done = true;      // past the super's end without a merge
            }
            if (!done && pos >= fstop_eos) { finish_at(pos, cnt); done = true; }
            if (done) { ps.park(); next_stop = NO_WORK; }
            else next_stop = min(s0 + ck_bnd(inext), fstop_eos);
        }
    };
    while (__any_sync(FULL_MASK, next_stop != NO_WORK)) {
        one_step();
        one_step();
        one_step();
        one_step();
    }
}

// Supers [sup_lo, sup_hi) of the plan.
template <bool FOLD>
__global__ void __launch_bounds__(ENT_TPB)
k_entropy_sync(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const uint32_t* __restrict__ seg_stream, uint32_t sup_lo, uint32_t sup_hi,
               uint32_t* __restrict__ seg_entry, uint32_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_cnt) {
    __shared__ SyncShared sh;
    const int t = threadIdx.x;
    const uint32_t u = sup_lo + blockIdx.x * SYNC_OWN + (uint32_t)t - 1u;        // slot t <-> super u; slot 0 = halo
    const bool valid = (blockIdx.x != 0 || t != 0) && u < sup_hi;
    SegCtx c{};
    if (valid) c = seg_ctx(payload, streams, seg_stream, u * SUPER);
    const uint32_t ftotal = c.ftotal, fstop_eos = eos_stop(c.ftotal), s0 = c.seg_start;

    // ---- phase A: speculative parse of the super from its first bit ---------------------------------------
    {
        Parser ps;
        ps.init_parked();
        uint32_t cnt = 0, j = 0, next_stop = NO_WORK;
        if (valid && s0 < fstop_eos) { ps.start(c.base, s0, s0 + SUPER_BITS, ftotal); next_stop = min(s0 + ck_bnd(0), fstop_eos); }
        else if (valid) {                                // super in the stream's trailing pad: nothing to parse
#pragma unroll
            for (int i = 0; i < NCK; i++) sh.cp[i][t] = 0u;
        }
        // (four steps per look at the loop condition: a lane that is done is parked, extra steps change nothing)
        auto one_step = [&]() {
            Parser::Sym y;
            const bool end = ps.step<false, FOLD>(y);
            cnt += end ? 1u : 0u;
            if (end && ps.fpos >= next_stop) {           // rare: a checkpoint boundary or the end of the stream passed
                const uint32_t pos = ps.fpos;
                const uint32_t rec = ((pos - s0) << 16) | cnt;
                while (j < (uint32_t)NCK && pos >= s0 + ck_bnd(j)) { sh.cp[j][t] = rec; j++; }
                if (j == (uint32_t)NCK || pos >= fstop_eos) {   // end of stream: no further block can start
                    for (; j < (uint32_t)NCK; j++) sh.cp[j][t] = rec;
                    ps.park();
                    next_stop = NO_WORK;
                } else {
                    next_stop = min(s0 + ck_bnd(j), fstop_eos);
                }
            }
        };
        while (__any_sync(FULL_MASK, next_stop != NO_WORK)) {
            one_step();
            one_step();
            one_step();
            one_step();
        }
    }
    __syncthreads();

    // ---- phase B, round 1: entry = predecessor's speculative exit -------------------------------------------
    const bool own = valid && t >= 1;
    const bool has_pred = own && c.seg != 0;             // same stream as slot t-1 (streams are SUPER-aligned)
    uint32_t E = has_pred ? s0 - SUPER_BITS + (sh.cp[NCK - 1][t - 1] >> 16) : c.bias;
    resolve_super<FOLD>(sh, c.base, E, s0, ftotal, t, own);
    if (valid && t == 0) sh.tpos[SUPER][0] = s0 + (sh.cp[NCK - 1][0] >> 16);     // the halo keeps its speculative exit
    __syncthreads();
    // ---- round 2: re-merge where the predecessor's resolved exit differs from its speculative one ----------
    const uint32_t E2 = has_pred ? sh.tpos[SUPER][t - 1] : E;
    __syncthreads();                                      // every tpos[SUPER][t-1] read before round 2 overwrites any
    const bool redo = own && E2 != E;
    if (__syncthreads_or(redo)) {
        if (redo) E = E2;
        resolve_super<FOLD>(sh, c.base, E, s0, ftotal, t, redo);
    }
    if (own) {
#pragma unroll
        for (int j = 0; j < SUPER; j++) {
            const uint32_t g = u * SUPER + j;
            seg_entry[g] = sh.tpos[j][t] - c.bias;
            seg_exit[g] = sh.tpos[j + 1][t] - c.bias;
            seg_cnt[g] = sh.tcnt[j + 1][t] - sh.tcnt[j][t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Chain fix-up + scan: one CTA per stream.
// ------------------------------------------------------------------------------------------------
constexpr int CHAIN_TPB = 128;

__global__ void __launch_bounds__(CHAIN_TPB)
k_entropy_chain(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                uint32_t* seg_entry, uint32_t* seg_exit, uint32_t* seg_cnt, uint32_t* __restrict__ seg_first,
                uint32_t* __restrict__ stream_blocks, unsigned long long* __restrict__ fixups) {
    const StreamDesc sd = streams[blockIdx.x];
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t bias = stream_bias(base);
    const uint32_t ftotal = sd.byte_len * 8u + bias;
    volatile uint32_t* v_entry = seg_entry + sd.seg_base;
    volatile uint32_t* v_exit = seg_exit + sd.seg_base;
    volatile uint32_t* v_cnt = seg_cnt + sd.seg_base;
    const int t = threadIdx.x;

    // 1. make the chain exact: entry[0] = 0, entry[i] = exit[i-1].  Every thread owns a CONTIGUOUS range of
    // segments and ripples through it in order: a re-parsed segment's new exit is handed to the next segment
    // at once, so a stream that does not self-synchronise (e.g. blocks that all end on coefficient 63: the
    // zig-zag index then never re-aligns quickly) costs one parse per segment plus a few rounds across range
    // boundaries -- not one round per segment.
    uint32_t nfix = 0;
    const uint32_t per = (sd.nseg + CHAIN_TPB - 1) / CHAIN_TPB;
    const uint32_t i_lo = min(sd.nseg, (uint32_t)t * per), i_hi = min(sd.nseg, i_lo + per);
    for (uint32_t round = 0; round <= CHAIN_TPB + 1; round++) {   // converges in <= CHAIN_TPB rounds; bound it anyway
        int changed = 0;
        for (uint32_t i = i_lo; i < i_hi; i++) {
            const uint32_t E = i ? v_exit[i - 1] : 0u;
            if (v_entry[i] != E) {
                uint32_t cnt;
                const uint32_t x = parse_segment_z(base, E + bias, (i + 1) * SEG_BITS + bias, ftotal, cnt) - bias;
                v_entry[i] = E;
                v_exit[i] = x;
                v_cnt[i] = cnt;
                changed = 1;
                nfix++;
            }
        }
        __threadfence_block();
        if (!__syncthreads_or(changed)) break;
    }
    if (nfix) atomicAdd(fixups, (unsigned long long)nfix);

    // 2. exclusive scan over segments: first block index.
    __shared__ uint32_t s_wcnt[CHAIN_TPB / 32];
    __shared__ uint32_t s_carry;
    if (t == 0) s_carry = 0;
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    for (uint32_t i0 = 0; i0 < sd.nseg; i0 += CHAIN_TPB) {
        uint32_t i = i0 + t;
        uint32_t cnt = i < sd.nseg ? v_cnt[i] : 0u;
        uint32_t icnt = cnt;                              // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t a = __shfl_up_sync(0xFFFFFFFFu, icnt, d);
            if (lane >= d) icnt += a;
        }
        if (lane == 31) s_wcnt[warp] = icnt;
        __syncthreads();
        uint32_t bcnt = s_carry;
        for (int w = 0; w < warp; w++) bcnt += s_wcnt[w];
        if (i < sd.nseg) seg_first[sd.seg_base + i] = bcnt + icnt - cnt;
        __syncthreads();
        if (t == CHAIN_TPB - 1) s_carry = bcnt + icnt;
        __syncthreads();
    }
    if (t == 0) stream_blocks[blockIdx.x] = s_carry;
}

// ------------------------------------------------------------------------------------------------
// Block index + symbol lists.  Walking its blocks from the exact state, a segment's thread writes
//   sym[seg * SYM_STRIDE + ...]  one entry per coded AC coefficient: zig-zag index | (block index & 31) << 6 |
//                                amplitude << 16 (the block bits name the lane that owns the block in
//                                k_decode_fused's warp tiles of 32 consecutive blocks)
//   blk_info[block].x            index of the block's first entry in sym[] (always inside the segment's region,
//                                so x / SYM_STRIDE identifies the segment), or BLK_NO_SEG for a block the stream
//                                does not hold
//   blk_info[block].y            DC level relative to the segment's entry (I frames: the int16 running sum `cur`
//                                of LIB/decoder/lossless_decode.c:73,94 restarted at 0; P frames: the DC delta
//                                itself, :91) | entries << 16
//   seg_dc[segment]              I frames: sum of the segment's DC deltas (mod 2^16); P frames: 0
// After this pass no kernel touches the bitstream again: the block-parallel decode kernels (decode.cu)
// read the lists with independent, look-ahead loads instead of a bit-serial dependent chain.
// ------------------------------------------------------------------------------------------------
// Small CTAs, one segment per lane: measured best (3.75 ms / 2000 frames at 1080p; pools of 2 / 4 / 8 segments per
// lane pulled from the CTA-wide counter: 4.0 / 4.45 / 4.8 ms) -- the hardware CTA scheduler balances many short CTAs
// better than lanes balance inside a long-lived one.  The pull loop below stays general (INDEX_SLOTS >= INDEX_TPB).
constexpr int INDEX_TPB = 64;
constexpr int INDEX_SLOTS = 64;      // segments per CTA

template <bool FOLD>
__global__ void __launch_bounds__(INDEX_TPB, 16)
k_entropy_index(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                const uint32_t* __restrict__ seg_stream, uint32_t seg_lo, uint32_t seg_hi,
                const uint32_t* __restrict__ seg_entry, const uint32_t* __restrict__ seg_cnt,
                const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ seg_dc, uint2* __restrict__ blk_info,
                uint32_t* __restrict__ sym, uint32_t sym_seg0, unsigned long long* __restrict__ n_entries) {
    __shared__ uint32_t s_next;
    const int t = threadIdx.x;
    const uint32_t g0 = seg_lo + blockIdx.x * INDEX_SLOTS;
    const uint32_t k_hi = min((uint32_t)INDEX_SLOTS, seg_hi - g0);
    if (t == 0) s_next = INDEX_TPB;
    __syncthreads();

    Parser ps;
    ps.init_parked();
    uint32_t g = 0, cnt = 0, k = 0, o = 0, o_blk = 0, o_end = 0, written = 0, tag = 0;
    uint2* bi = nullptr;
    int cur = 0;
    bool pframe = false;
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // the last (o & 7) entries, newest in q[7]: stored one full 32-byte
                                               // sector at a time (a partial-sector store makes L2 fetch the rest)
    // Take segments from the CTA's counter until one holds blocks (cnt != 0) or none is left (cnt == 0).
    auto grab = [&](uint32_t slot) {
        for (;; slot = atomicAdd(&s_next, 1u)) {
            k = 0; cnt = 0;
            if (slot >= k_hi) { ps.park(); return; }
            g = g0 + slot;
            const SegCtx c = seg_ctx(payload, streams, seg_stream, g);
            const StreamDesc* sd = streams + c.sid;
            const uint32_t nb = sd->nb, first = seg_first[g];
            cnt = first >= nb ? 0u : min(seg_cnt[g], nb - first);   // trailing pad bits can look like blocks
            // A stream that ends early leaves the remaining blocks empty (zero coefficients).
            if (c.seg + 1u == sd->nseg)
                for (uint32_t b = first + cnt; b < nb; b++) blk_info[sd->block_base + b] = make_uint2(BLK_NO_SEG, 0);
            if (cnt == 0) { seg_dc[g] = 0u; continue; }
            ps.start(c.base, seg_entry[g] + c.bias, c.seg_start + SEG_BITS, c.ftotal);
            bi = blk_info + sd->block_base + first;
            tag = (first & 31u) << 6;                    // (index of the block being parsed & 31) << 6
            o = o_blk = (g - sym_seg0) * SYM_STRIDE;     // chunk-relative entry index
            o_end = o + SYM_STRIDE;
            cur = 0;
            pframe = sd->ptype != 0;
            return;
        }
    };
    grab((uint32_t)t);
    // (four steps per look at the loop condition: a lane that is done is parked, extra steps change nothing)
    auto one_step = [&]() {
        Parser::Sym y;
        const bool end = ps.step<true, FOLD>(y);
        if (y.dc) { cur = pframe ? y.e : cur + y.e; o_blk = o; }
        if (y.coded && y.at < (64u << 24)) {             // (a parked lane never sees a coded symbol)
#pragma unroll
            for (int i = 0; i < 7; i++) q[i] = q[i + 1];
            q[7] = (y.at >> 24) | tag | ((uint32_t)y.e << 16);
            o++;
            if ((o & 7u) == 0u && o <= o_end) st_global_v8(sym + o - 8, q);   // never overflows on conforming streams
        }
        if (end && k < cnt) {
            bi[k] = make_uint2(min(o_blk, o_end - 1u), ((uint32_t)cur & 0xFFFFu) | ((min(o, o_end) - min(o_blk, o_end)) << 16));
            tag = (tag + 64u) & 0x7C0u;
            if (++k == cnt) {                            // segment done
                seg_dc[g] = pframe ? 0u : ((uint32_t)cur & 0xFFFFu);
                if ((o & 7u) && o < o_end) {             // flush the partial group (entries beyond o are never read)
                    const uint32_t r = o & 7u;
                    uint32_t v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {        // v[i] = q[8 - r + i] for i < r
                        uint32_t x = 0;
#pragma unroll
                        for (int j = 1; j < 8; j++) if ((uint32_t)j == r && 8 - j + i < 8) x = q[8 - j + i];
                        v[i] = x;
                    }
                    st_global_v8(sym + (o & ~7u), v);
                }
                written += min(o, o_end) - (o_end - SYM_STRIDE);
                grab(atomicAdd(&s_next, 1u));
            }
        }
    };
    while (__any_sync(FULL_MASK, cnt != 0u)) {
        one_step();
        one_step();
        one_step();
        one_step();
    }
    {   // statistics: list entries written by this launch (one atomic per warp)
#pragma unroll
        for (int d = 16; d; d >>= 1) written += __shfl_xor_sync(FULL_MASK, written, d);
        if ((t & 31) == 0 && written) atomicAdd(n_entries, (unsigned long long)written);
    }
}

// ------------------------------------------------------------------------------------------------
// DC predictors: one warp per stream turns seg_dc (the segments' DC totals) into the exclusive prefix
// sum mod 2^16 = the value of `cur` (lossless_decode.c:73,94) entering each segment.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_entropy_dcscan(const StreamDesc* __restrict__ streams, uint32_t n_streams, uint32_t* __restrict__ seg_dc) {
    const uint32_t s = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (s >= n_streams) return;
    const StreamDesc sd = streams[s];
    const int lane = threadIdx.x & 31;
    uint32_t* v = seg_dc + sd.seg_base;
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < sd.nseg; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t x = i < sd.nseg ? v[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += a;
        }
        if (i < sd.nseg) v[i] = (carry + inc - x) & 0xFFFFu;
        carry += __shfl_sync(FULL_MASK, inc, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// Host launchers (declared in runtime.h).
// ------------------------------------------------------------------------------------------------
// One warp per stream fills seg_stream[] (global segment -> stream) for the whole plan.
__global__ void __launch_bounds__(128)
k_seg_stream(const StreamDesc* __restrict__ streams, uint32_t n_streams, uint32_t* __restrict__ seg_stream) {
    const uint32_t s = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (s >= n_streams) return;
    const uint32_t base = streams[s].seg_base, n = (streams[s].nseg + SUPER - 1) / SUPER * SUPER;   // incl. the padding segments
    for (uint32_t i = threadIdx.x & 31; i < n; i += 32) seg_stream[base + i] = s;
}
cudaError_t launch_seg_stream(const StreamDesc* d_streams, uint32_t n_streams, uint32_t* d_seg_stream, cudaStream_t s) {
    if (n_streams == 0) return cudaSuccess;
    k_seg_stream<<<(n_streams + 3) / 4, 128, 0, s>>>(d_streams, n_streams, d_seg_stream);
    return cudaGetLastError();
}

cudaError_t launch_entropy_sync(const EntropyJob& j, cudaStream_t s) {
    if (j.seg_hi <= j.seg_lo) return cudaSuccess;
    const uint32_t sup_lo = j.seg_lo / SUPER, sup_hi = j.seg_hi / SUPER;       // chunk ranges are SUPER-aligned (build_plan)
    const uint32_t n = sup_hi - sup_lo;
    const unsigned grid = (n + SYNC_OWN - 1) / SYNC_OWN;
    if (j.fold_end)
        k_entropy_sync<true><<<grid, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_seg_stream, sup_lo, sup_hi, j.d_seg_entry,
                                                      j.d_seg_exit, j.d_seg_cnt);
    else
        k_entropy_sync<false><<<grid, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_seg_stream, sup_lo, sup_hi, j.d_seg_entry,
                                                       j.d_seg_exit, j.d_seg_cnt);
    return cudaGetLastError();
}
cudaError_t launch_entropy_chain(const EntropyJob& j, cudaStream_t s) {
    if (j.n_streams == 0) return cudaSuccess;
    k_entropy_chain<<<j.n_streams, CHAIN_TPB, 0, s>>>(j.d_payload, j.d_streams + j.stream_lo, j.d_seg_entry,
                                                      j.d_seg_exit, j.d_seg_cnt, j.d_seg_first,
                                                      j.d_stream_blocks + j.stream_lo, j.d_fixups);
    return cudaGetLastError();
}
cudaError_t launch_entropy_index(const EntropyJob& j, cudaStream_t s) {
    if (j.seg_hi <= j.seg_lo) return cudaSuccess;
    const uint32_t n = j.seg_hi - j.seg_lo;
    const unsigned grid = (n + INDEX_SLOTS - 1) / INDEX_SLOTS;
    if (j.fold_end)
        k_entropy_index<true><<<grid, INDEX_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_seg_stream, j.seg_lo, j.seg_hi, j.d_seg_entry,
                                                         j.d_seg_cnt, j.d_seg_first, j.d_seg_dc, j.d_blk_info, j.d_sym, j.sym_seg0,
                                                         j.d_fixups + 1);
    else
        k_entropy_index<false><<<grid, INDEX_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_seg_stream, j.seg_lo, j.seg_hi, j.d_seg_entry,
                                                          j.d_seg_cnt, j.d_seg_first, j.d_seg_dc, j.d_blk_info, j.d_sym, j.sym_seg0,
                                                          j.d_fixups + 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_entropy_dcscan<<<(j.n_streams + 3) / 4, 128, 0, s>>>(j.d_streams + j.stream_lo, j.n_streams, j.d_seg_dc);
    return cudaGetLastError();
}

}  // namespace mj
