// entropy.cu -- batched, segment-parallel "lossless" decode (sm_100a): synchronisation, chain, block index.
//
// Replaces the serial walk of lossless_decode(), LIB/decoder/lossless_decode.c:60-135 (LIB =
// /root/reference/core0/software/common/libs/mjpeg423), for MANY plane streams at once and, inside a
// stream, for many fixed-size bitstream segments in parallel.  The code has no markers or restart
// intervals, so segment entry points are found by self-synchronisation:
//
//   k_entropy_sync   a CTA owns a run of consecutive segments (of any number of streams).  Phase A parses
//                    every segment speculatively from its first bit as if a block started there, leaving
//                    NCP checkpoints (first block start at or after every CP_BITS boundary, blocks so far).
//                    Phase B takes the predecessor's speculative exit as the segment's entry and parses
//                    only until it MERGES with the recorded trajectory (same bit position at a block
//                    start => identical future); a second round re-merges the segments whose
//                    predecessor's exit moved.
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
// Every pass advances with Parser::step() (common.cuh) in a UNIFORM loop: one symbol per iteration for
// every lane, DC/AC and block-end handling predicated, the rare events (checkpoint, end of a job) in a
// short divergent branch.  Work is handed out per LANE: segments (and merge jobs) are pulled from a
// CTA-wide counter, so a lane whose segment is done starts the next one instead of idling until the
// slowest lane of its warp finishes (symbols per segment vary by 2x with the picture content).
#include "common.cuh"
#include "runtime.h"

namespace mj {

// All positions inside the kernels are "f" positions (stream bit position + stream_bias(), see Parser);
// the per-segment arrays in global memory hold plain stream bit positions.
//
// DC levels are NOT tracked by the synchronisation passes (their symbol loop skips amplitudes altogether):
// the index pass records every block's DC level relative to its segment's first block plus the segment's
// DC total, k_entropy_dcscan turns the totals into the predictor entering each segment, and the decode
// kernels add it (decode.cu: dc_pred()).

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
        ps.start(base, entry, ftotal);
        for (;;) {
            Parser::Sym y;
            if (ps.step<false>(ftotal, y)) {
                cnt++;
                if (ps.fpos >= stop) break;
            }
        }
        pos = ps.fpos;
    }
    return pos;
}

// ------------------------------------------------------------------------------------------------
// Speculative parse + merge.  CTA c handles SYNC_SLOTS consecutive global segments: slot 0 is the
// segment BEFORE its SYNC_OWN own ones (halo), so that slot 1 has a predecessor exit without any
// inter-CTA dependency.  A slot's predecessor is the previous slot unless the segment is the first of
// its stream.
//
// Phase B: a segment whose true entry E (the predecessor's exit) is not where the speculation started is
// a MERGE JOB: parse from E until the trajectory meets the recorded one (tested at the first block start
// at or after every checkpoint boundary -- once merged, that IS the checkpoint).  Jobs are short (a few
// hundred bits) and of very uneven length; pulled per lane from the CTA's pool they keep the lanes busy.
// ------------------------------------------------------------------------------------------------
constexpr int SYNC_SLOTS = 512;
constexpr int SYNC_OWN = SYNC_SLOTS - 1;
constexpr int SYNC_PER_THREAD = SYNC_SLOTS / ENT_TPB;
constexpr uint32_t NO_WORK = 0xFFFFFFFFu;

struct SyncShared {
    uint32_t cp[NCP][SYNC_SLOTS];  // checkpoint j: (first block start >= seg_start + (j+1)*CP_BITS, minus seg_start) << 16
                                   //               | blocks started before it
    uint32_t E[SYNC_SLOTS];        // entry (f position) the result below was resolved for
    uint32_t rexit[SYNC_SLOTS];    // resolved exit (f position)
    uint32_t rcnt[SYNC_SLOTS];     // resolved block count
    uint32_t job[SYNC_SLOTS];      // pool of slots to resolve by parsing
    uint32_t njobs, next;
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

// Resolve slot k for entry E without parsing where possible; returns false when it needs a merge job.
__device__ __forceinline__ bool resolve_trivial(SyncShared& sh, uint32_t k, uint32_t E, uint32_t seg_start, uint32_t fstop_eos) {
    const uint32_t last = sh.cp[NCP - 1][k];
    const uint32_t spec_exit = seg_start + (last >> 16), spec_cnt = last & 0xFFFFu;
    sh.E[k] = E;
    if (E >= seg_start + SEG_BITS || E >= fstop_eos) { sh.rexit[k] = E; sh.rcnt[k] = 0; return true; }   // owns nothing
    if (E == seg_start) { sh.rexit[k] = spec_exit; sh.rcnt[k] = spec_cnt; return true; }                 // speculation was right
    if (E >= seg_start + CP_BITS) {                      // E itself may be a recorded block start
        const uint32_t c = sh.cp[(E - seg_start) / CP_BITS - 1u][k];
        if (seg_start + (c >> 16) == E) { sh.rexit[k] = spec_exit; sh.rcnt[k] = spec_cnt - (c & 0xFFFFu); return true; }
    }
    return false;
}

// CTA-COLLECTIVE (every warp, converged): the lanes drain the job pool.  g0 = global segment of slot 0.
__device__ __forceinline__ void run_merge_jobs(SyncShared& sh, const uint8_t* __restrict__ payload,
                                               const StreamDesc* __restrict__ streams,
                                               const uint32_t* __restrict__ seg_stream, uint32_t g0) {
    const uint32_t njobs = sh.njobs;
    Parser ps;
    ps.init_parked();
    uint32_t jt = 0, seg_start = 0, ftotal = 0, fstop_eos = 0, cnt = 0, next_cp = 0, next_stop = NO_WORK;
    auto grab = [&]() {
        const uint32_t k = atomicAdd(&sh.next, 1u);
        if (k < njobs) {
            jt = sh.job[k];
            const SegCtx c = seg_ctx(payload, streams, seg_stream, g0 + jt);
            const uint32_t E = sh.E[jt];
            seg_start = c.seg_start; ftotal = c.ftotal; fstop_eos = eos_stop(c.ftotal);
            ps.start(c.base, E, ftotal);
            cnt = 0;
            next_cp = seg_start + ((E - seg_start) / CP_BITS + 1u) * CP_BITS;   // first boundary after E (E >= seg_start)
            next_stop = min(next_cp, fstop_eos);
        } else {
            ps.park();
            next_stop = NO_WORK;
        }
    };
    grab();
    while (__any_sync(FULL_MASK, next_stop != NO_WORK)) {
        Parser::Sym y;
        const bool end = ps.step<false>(ftotal, y);
        cnt += end ? 1u : 0u;
        if (end && ps.fpos >= next_stop) {               // rare: first block start past a boundary / end of stream
            const uint32_t pos = ps.fpos;
            bool done = false;
            uint32_t exit_pos = pos;
            if (pos >= next_cp) {
                const uint32_t j = min((uint32_t)NCP, (pos - seg_start) / CP_BITS) - 1u;
                const uint32_t c = sh.cp[j][jt];
                if (seg_start + (c >> 16) == pos) {      // merged with the speculative trajectory
                    const uint32_t last = sh.cp[NCP - 1][jt];
                    cnt += (last & 0xFFFFu) - (c & 0xFFFFu);
                    exit_pos = seg_start + (last >> 16);
                    done = true;
                }
                next_cp = seg_start + (j + 2u) * CP_BITS;
            }
            if (!done && (pos >= seg_start + SEG_BITS || pos >= fstop_eos)) done = true;
            if (done) { sh.rexit[jt] = exit_pos; sh.rcnt[jt] = cnt; grab(); }
            else next_stop = min(next_cp, fstop_eos);
        }
    }
}

// Segments [seg_lo, seg_hi) of the plan.
__global__ void __launch_bounds__(ENT_TPB)
k_entropy_sync(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const uint32_t* __restrict__ seg_stream, uint32_t seg_lo, uint32_t seg_hi,
               uint32_t* __restrict__ seg_entry, uint32_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_cnt) {
    __shared__ SyncShared sh;
    const int t = threadIdx.x;
    // slot k <-> global segment g0 + k; slot 0 (the halo) does not exist for the first CTA
    const uint32_t g0 = seg_lo + blockIdx.x * SYNC_OWN - 1u;
    const uint32_t k_lo = blockIdx.x == 0 ? 1u : 0u;
    const uint32_t k_hi = min((uint32_t)SYNC_SLOTS, seg_hi - g0);       // slots [k_lo, k_hi) exist
    if (t == 0) { sh.njobs = 0; sh.next = ENT_TPB; }
    __syncthreads();

    // ---- phase A: speculative parse of every slot from the segment's first bit ----------------------------
    {
        Parser ps;
        ps.init_parked();
        uint32_t slot = 0, seg_start = 0, ftotal = 0, fstop_eos = 0, cnt = 0, j = 0, next_cp = 0, next_stop = NO_WORK;
        auto grab = [&](uint32_t k) {
            for (;; k = atomicAdd(&sh.next, 1u)) {
                if (k >= k_hi) { ps.park(); next_stop = NO_WORK; return; }
                if (k < k_lo) continue;
                const SegCtx c = seg_ctx(payload, streams, seg_stream, g0 + k);
                slot = k; seg_start = c.seg_start; ftotal = c.ftotal; fstop_eos = eos_stop(c.ftotal);
                if (seg_start >= fstop_eos) {            // segment in the stream's trailing pad: nothing to parse
#pragma unroll
                    for (int i = 0; i < NCP; i++) sh.cp[i][k] = 0u;
                    continue;
                }
                ps.start(c.base, seg_start, ftotal);
                cnt = 0; j = 0;
                next_cp = seg_start + CP_BITS;
                next_stop = min(next_cp, fstop_eos);
                return;
            }
        };
        grab((uint32_t)t);
        while (__any_sync(FULL_MASK, next_stop != NO_WORK)) {
            Parser::Sym y;
            const bool end = ps.step<false>(ftotal, y);
            cnt += end ? 1u : 0u;
            if (end && ps.fpos >= next_stop) {           // rare: a checkpoint boundary or the end of the stream passed
                const uint32_t pos = ps.fpos;
                const uint32_t rec = ((pos - seg_start) << 16) | cnt;
                while (j < (uint32_t)NCP && pos >= next_cp) { sh.cp[j][slot] = rec; j++; next_cp += CP_BITS; }
                if (j == (uint32_t)NCP || pos >= fstop_eos) {   // end of stream: no further block can start
                    for (; j < (uint32_t)NCP; j++) sh.cp[j][slot] = rec;
                    grab(atomicAdd(&sh.next, 1u));
                } else {
                    next_stop = min(next_cp, fstop_eos);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B, round 1: entry = predecessor's speculative exit -------------------------------------------
    if (t == 0) sh.next = 0;
#pragma unroll
    for (int i = 0; i < SYNC_PER_THREAD; i++) {
        const uint32_t k = (uint32_t)t + (uint32_t)i * ENT_TPB;
        if (k < k_lo || k >= k_hi) continue;
        const SegCtx c = seg_ctx(payload, streams, seg_stream, g0 + k);
        if (k == 0) {                                    // the halo keeps its speculative exit
            sh.E[0] = c.seg_start; sh.rexit[0] = c.seg_start + (sh.cp[NCP - 1][0] >> 16); sh.rcnt[0] = 0;
            continue;
        }
        const uint32_t E = c.seg ? c.seg_start - SEG_BITS + (sh.cp[NCP - 1][k - 1] >> 16) : c.bias;
        if (!resolve_trivial(sh, k, E, c.seg_start, eos_stop(c.ftotal))) sh.job[atomicAdd(&sh.njobs, 1u)] = k;
    }
    __syncthreads();
    run_merge_jobs(sh, payload, streams, seg_stream, g0);
    __syncthreads();
    // ---- round 2: re-merge where the predecessor's resolved exit differs from its speculative one ----------
    uint32_t E2[SYNC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < SYNC_PER_THREAD; i++) {
        const uint32_t k = (uint32_t)t + (uint32_t)i * ENT_TPB;
        E2[i] = NO_WORK;
        if (k < max(k_lo, 1u) || k >= k_hi) continue;
        const SegCtx c = seg_ctx(payload, streams, seg_stream, g0 + k);
        if (c.seg && sh.rexit[k - 1] != sh.E[k]) E2[i] = sh.rexit[k - 1];
    }
    __syncthreads();                                      // every rexit[k-1] read before round 2 overwrites any
    if (t == 0) { sh.njobs = 0; sh.next = 0; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SYNC_PER_THREAD; i++) {
        const uint32_t k = (uint32_t)t + (uint32_t)i * ENT_TPB;
        if (E2[i] == NO_WORK) continue;
        const SegCtx c = seg_ctx(payload, streams, seg_stream, g0 + k);
        if (!resolve_trivial(sh, k, E2[i], c.seg_start, eos_stop(c.ftotal))) sh.job[atomicAdd(&sh.njobs, 1u)] = k;
    }
    __syncthreads();
    if (sh.njobs) run_merge_jobs(sh, payload, streams, seg_stream, g0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SYNC_PER_THREAD; i++) {
        const uint32_t k = (uint32_t)t + (uint32_t)i * ENT_TPB;
        if (k < 1u || k >= k_hi) continue;
        const uint32_t g = g0 + k;
        const uint32_t bias = stream_bias(payload + streams[__ldg(seg_stream + g)].byte_off);
        seg_entry[g] = sh.E[k] - bias;
        seg_exit[g] = sh.rexit[k] - bias;
        seg_cnt[g] = sh.rcnt[k];
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
                const uint32_t x = parse_segment(base, E + bias, (i + 1) * SEG_BITS + bias, ftotal, cnt) - bias;
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
constexpr int INDEX_SLOTS = 512;     // segments per CTA

__global__ void __launch_bounds__(ENT_TPB, 8)
k_entropy_index(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                const uint32_t* __restrict__ seg_stream, uint32_t seg_lo, uint32_t seg_hi,
                const uint32_t* __restrict__ seg_entry, const uint32_t* __restrict__ seg_cnt,
                const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ seg_dc, uint2* __restrict__ blk_info,
                uint32_t* __restrict__ sym, uint32_t sym_seg0, unsigned long long* __restrict__ n_entries) {
    __shared__ uint32_t s_next;
    const int t = threadIdx.x;
    const uint32_t g0 = seg_lo + blockIdx.x * INDEX_SLOTS;
    const uint32_t k_hi = min((uint32_t)INDEX_SLOTS, seg_hi - g0);
    if (t == 0) s_next = ENT_TPB;
    __syncthreads();

    Parser ps;
    ps.init_parked();
    uint32_t g = 0, ftotal = 0, cnt = 0, k = 0, o = 0, o_blk = 0, o_end = 0, written = 0, tag = 0;
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
            ftotal = c.ftotal;
            ps.start(c.base, seg_entry[g] + c.bias, ftotal);
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
    while (__any_sync(FULL_MASK, cnt != 0u)) {
        Parser::Sym y;
        const bool end = ps.step<true>(ftotal, y);
        if (y.dc) { cur = pframe ? y.e : cur + y.e; o_blk = o; }
        if (y.coded && y.at < 64u) {                     // (a parked lane never sees a coded symbol)
#pragma unroll
            for (int i = 0; i < 7; i++) q[i] = q[i + 1];
            q[7] = y.at | tag | ((uint32_t)y.e << 16);
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
    const uint32_t base = streams[s].seg_base, n = streams[s].nseg;
    for (uint32_t i = threadIdx.x & 31; i < n; i += 32) seg_stream[base + i] = s;
}
cudaError_t launch_seg_stream(const StreamDesc* d_streams, uint32_t n_streams, uint32_t* d_seg_stream, cudaStream_t s) {
    if (n_streams == 0) return cudaSuccess;
    k_seg_stream<<<(n_streams + 3) / 4, 128, 0, s>>>(d_streams, n_streams, d_seg_stream);
    return cudaGetLastError();
}

cudaError_t launch_entropy_sync(const EntropyJob& j, cudaStream_t s) {
    if (j.seg_hi <= j.seg_lo) return cudaSuccess;
    const uint32_t n = j.seg_hi - j.seg_lo;
    k_entropy_sync<<<(n + SYNC_OWN - 1) / SYNC_OWN, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_seg_stream, j.seg_lo,
                                                                    j.seg_hi, j.d_seg_entry, j.d_seg_exit, j.d_seg_cnt);
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
    k_entropy_index<<<(n + INDEX_SLOTS - 1) / INDEX_SLOTS, ENT_TPB, 0, s>>>(
        j.d_payload, j.d_streams, j.d_seg_stream, j.seg_lo, j.seg_hi, j.d_seg_entry, j.d_seg_cnt, j.d_seg_first, j.d_seg_dc,
        j.d_blk_info, j.d_sym, j.sym_seg0, j.d_fixups + 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_entropy_dcscan<<<(j.n_streams + 3) / 4, 128, 0, s>>>(j.d_streams + j.stream_lo, j.n_streams, j.d_seg_dc);
    return cudaGetLastError();
}

}  // namespace mj
