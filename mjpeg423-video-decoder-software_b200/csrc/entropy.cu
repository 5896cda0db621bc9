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
//                    then exclusive-scans the block counts to give every segment its first block index.
//   k_entropy_index  one thread per segment walks its blocks from the now exact state and writes every
//                    coded coefficient as a 32-bit entry into the segment's symbol list, plus per block
//                    the list position, entry count and segment-relative DC level: what the block-parallel
//                    decode kernels (decode.cu) consume without touching the bitstream again.
//   k_entropy_dcscan exclusive scan (mod 2^16) of the segments' DC totals: the DC
//                    predictor entering every segment.
//
// Every pass advances with Parser::step() (common.cuh): one flat loop, one symbol per iteration per
// lane, DC/AC and block-end handling predicated, so the lanes of a warp stay converged.
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
// Speculative parse + merge.  A tile owns ENT_TPB-1 segments; thread 0 parses the segment BEFORE the
// tile (halo) so that thread 1 has a predecessor exit without any inter-CTA dependency.
//
// Phase A: every thread parses its segment from the segment's first bit, recording NCP checkpoints.
// Phase B: a segment whose true entry E (the predecessor's exit) is not where the speculation started is
// a MERGE JOB: parse from E until the trajectory meets the recorded one (tested at the first block start
// at or after every checkpoint boundary -- once merged, that IS the checkpoint).  Jobs are short (a few
// hundred bits) and of very uneven length, so they are pooled per CTA and pulled by the lanes of ONE
// warp: a lane that finishes a job takes the next one, which keeps the warp's lanes busy instead of
// waiting for the longest of 32 jobs.
// ------------------------------------------------------------------------------------------------
struct SyncShared {
    uint32_t pos[NCP][ENT_TPB];    // [j] = first block start >= seg_start + (j+1)*CP_BITS (f position)
    uint32_t cnt[NCP][ENT_TPB];    // blocks started before it
    uint32_t E[ENT_TPB];           // entry the result below was resolved for
    uint32_t rexit[ENT_TPB];       // resolved exit
    uint32_t rcnt[ENT_TPB];        // resolved block count
    uint32_t job[ENT_TPB];         // pool of slots to resolve by parsing
    uint32_t njobs, next;
};

// Resolve slot t for entry E without parsing where possible; returns false when it needs a merge job.
__device__ __forceinline__ bool resolve_trivial(SyncShared& sh, int t, uint32_t E, uint32_t seg_start, uint32_t fstop_eos) {
    const uint32_t spec_exit = sh.pos[NCP - 1][t], spec_cnt = sh.cnt[NCP - 1][t];
    sh.E[t] = E;
    if (E >= seg_start + SEG_BITS || E >= fstop_eos) { sh.rexit[t] = E; sh.rcnt[t] = 0; return true; }   // owns nothing
    if (E == seg_start) { sh.rexit[t] = spec_exit; sh.rcnt[t] = spec_cnt; return true; }                 // speculation was right
    if (E >= seg_start + CP_BITS) {                      // E itself may be a recorded block start
        const int j = (int)((E - seg_start) / CP_BITS) - 1;
        if (sh.pos[j][t] == E) { sh.rexit[t] = spec_exit; sh.rcnt[t] = spec_cnt - sh.cnt[j][t]; return true; }
    }
    return false;
}

// WARP-COLLECTIVE: the calling warp drains the job pool.  seg0_start = f position of slot 0's segment.
__device__ __forceinline__ void run_merge_jobs(SyncShared& sh, const uint8_t* base, uint32_t seg0_start, uint32_t ftotal) {
    const uint32_t fstop_eos = eos_stop(ftotal);
    const uint32_t njobs = sh.njobs;
    Parser ps;
    ps.init_parked();
    uint32_t jt = 0, seg_start = 0, cnt = 0, next_cp = 0, next_stop = 0xFFFFFFFFu;   // next_stop == ~0: no job
    auto grab = [&]() {
        const uint32_t k = atomicAdd(&sh.next, 1u);
        if (k < njobs) {
            jt = sh.job[k];
            const uint32_t E = sh.E[jt];
            seg_start = seg0_start + jt * SEG_BITS;
            ps.start(base, E, ftotal);
            cnt = 0;
            next_cp = seg_start + ((E - seg_start) / CP_BITS + 1u) * CP_BITS;   // first boundary after E (E >= seg_start)
            next_stop = min(next_cp, fstop_eos);
        } else {
            ps.park();
            next_stop = 0xFFFFFFFFu;
        }
    };
    grab();
    while (__any_sync(FULL_MASK, next_stop != 0xFFFFFFFFu)) {
        Parser::Sym y;
        const bool end = ps.step<false>(ftotal, y);
        cnt += end ? 1u : 0u;
        if (end && ps.fpos >= next_stop) {               // rare: first block start past a boundary / end of stream
            const uint32_t pos = ps.fpos;
            bool done = false;
            uint32_t exit_pos = pos;
            if (pos >= next_cp) {
                const int j = (int)min((uint32_t)NCP, (pos - seg_start) / CP_BITS) - 1;
                if (sh.pos[j][jt] == pos) {              // merged with the speculative trajectory
                    cnt += sh.cnt[NCP - 1][jt] - sh.cnt[j][jt];
                    exit_pos = sh.pos[NCP - 1][jt];
                    done = true;
                }
                next_cp = seg_start + (uint32_t)(j + 2) * CP_BITS;
            }
            if (!done && (pos >= seg_start + SEG_BITS || pos >= fstop_eos)) done = true;
            if (done) { sh.rexit[jt] = exit_pos; sh.rcnt[jt] = cnt; grab(); }
            else next_stop = min(next_cp, fstop_eos);
        }
    }
}

__global__ void __launch_bounds__(ENT_TPB)
k_entropy_sync(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
               const TileDesc* __restrict__ tiles, uint32_t* __restrict__ seg_entry,
               uint32_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_cnt) {
    __shared__ SyncShared sh;
    const int t = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const StreamDesc sd = streams[td.stream];
    const int seg = (int)td.seg0 - 1 + t;
    const bool valid = seg >= 0 && seg < (int)sd.nseg;
    const uint8_t* base = payload + sd.byte_off;
    const uint32_t bias = stream_bias(base);
    const uint32_t ftotal = sd.byte_len * 8u + bias;
    const uint32_t fstop_eos = eos_stop(ftotal);
    const uint32_t seg_start = (uint32_t)seg * SEG_BITS + bias;
    if (t == 0) { sh.njobs = 0; sh.next = 0; }

    // ---- phase A: speculative parse from the segment's first bit ------------------------------------------
    {
        int j = 0;
        uint32_t cnt = 0, pos = seg_start;
        uint32_t next_cp = seg_start + CP_BITS;
        uint32_t next_stop = 0xFFFFFFFFu;                 // ~0: this lane is done (or has nothing to parse)
        Parser ps;
        if (valid && pos < fstop_eos) { ps.start(base, seg_start, ftotal); next_stop = min(next_cp, fstop_eos); }
        else ps.init_parked();
        while (__any_sync(FULL_MASK, next_stop != 0xFFFFFFFFu)) {
            Parser::Sym y;
            const bool end = ps.step<false>(ftotal, y);
            cnt += end ? 1u : 0u;
            if (end && ps.fpos >= next_stop) {           // rare: a checkpoint boundary or the end of the stream passed
                pos = ps.fpos;
                while (j < NCP && pos >= next_cp) { sh.pos[j][t] = pos; sh.cnt[j][t] = cnt; j++; next_cp += CP_BITS; }
                if (j == NCP || pos >= fstop_eos) {      // end of stream: no further block can start
                    for (; j < NCP; j++) { sh.pos[j][t] = pos; sh.cnt[j][t] = cnt; }
                    next_stop = 0xFFFFFFFFu;
                    ps.park();
                } else {
                    next_stop = min(next_cp, fstop_eos);
                }
            }
        }
        if (valid) for (; j < NCP; j++) { sh.pos[j][t] = pos; sh.cnt[j][t] = 0; }   // nothing parsed (segment in the pad)
    }
    __syncthreads();

    // ---- phase B, round 1: entry = predecessor's speculative exit; round 2: its resolved exit ----------------
    const bool own = valid && t >= 1;
    const uint32_t seg0_start = ((uint32_t)td.seg0 - 1u) * SEG_BITS + bias;   // slot 0 (wraps for seg0 == 0: slot 0 is never a job)
    uint32_t E = (own && seg != 0) ? sh.pos[NCP - 1][t - 1] : bias;
    if (!own) { sh.E[t] = E; sh.rexit[t] = valid ? sh.pos[NCP - 1][t] : 0u; sh.rcnt[t] = 0; }   // the halo keeps its speculative exit
    else if (!resolve_trivial(sh, t, E, seg_start, fstop_eos)) sh.job[atomicAdd(&sh.njobs, 1u)] = (uint32_t)t;
    __syncthreads();
    if (t < 32) run_merge_jobs(sh, base, seg0_start, ftotal);
    __syncthreads();
    const uint32_t E2 = (own && seg != 0) ? sh.rexit[t - 1] : E;
    __syncthreads();                                      // every rexit[t-1] read before round 2 overwrites any
    if (t == 0) { sh.njobs = 0; sh.next = 0; }
    __syncthreads();
    if (own && E2 != E) {
        E = E2;
        if (!resolve_trivial(sh, t, E, seg_start, fstop_eos)) sh.job[atomicAdd(&sh.njobs, 1u)] = (uint32_t)t;
    }
    __syncthreads();
    if (t < 32 && sh.njobs) run_merge_jobs(sh, base, seg0_start, ftotal);
    __syncthreads();
    if (own) {
        const uint32_t g = sd.seg_base + (uint32_t)seg;
        seg_entry[g] = E - bias;
        seg_exit[g] = sh.rexit[t] - bias;
        seg_cnt[g] = sh.rcnt[t];
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

    // 1. make the chain exact: entry[0] = 0, entry[i] = exit[i-1].
    uint32_t nfix = 0;
    for (uint32_t sweep = 0; sweep <= sd.nseg + 1; sweep++) {   // converges in <= nseg sweeps; bound it anyway
        int changed = 0;
        for (uint32_t i = t; i < sd.nseg; i += CHAIN_TPB) {
            uint32_t E = i ? v_exit[i - 1] : 0u;
            if (v_entry[i] != E) {
                uint32_t cnt;
                uint32_t x = parse_segment(base, E + bias, (i + 1) * SEG_BITS + bias, ftotal, cnt) - bias;
                v_entry[i] = E;
                v_exit[i] = x;
                v_cnt[i] = cnt;
                changed = 1;
                nfix++;
            }
        }
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
//   sym[seg * SYM_STRIDE + ...]  one entry per coded AC coefficient: zig-zag index | amplitude << 16
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
__global__ void __launch_bounds__(ENT_TPB)
k_entropy_index(const uint8_t* __restrict__ payload, const StreamDesc* __restrict__ streams,
                const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ seg_entry,
                const uint32_t* __restrict__ seg_cnt, const uint32_t* __restrict__ seg_first,
                uint32_t* __restrict__ seg_dc, uint2* __restrict__ blk_info, uint32_t* __restrict__ sym, uint32_t sym_seg0,
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
    uint32_t cnt = valid ? seg_cnt[g] : 0u;
    cnt = first >= sd.nb ? 0u : min(cnt, sd.nb - first);          // trailing pad bits can look like blocks
    uint2* bi = blk_info + sd.block_base + first;
    const uint32_t o_base = (g - sym_seg0) * SYM_STRIDE, o_end = o_base + SYM_STRIDE;   // chunk-relative entry index
    uint32_t o = o_base, o_blk = o_base;
    Parser ps;
    if (cnt) ps.start(base, seg_entry[g] + bias, ftotal);
    else ps.init_parked();
    int cur = 0;
    const bool pframe = sd.ptype != 0;
    uint32_t k = 0;
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // the last (o & 7) entries, newest in q[7]: stored one full 32-byte
                                               // sector at a time (a partial-sector store makes L2 fetch the rest)
    while (__any_sync(FULL_MASK, k < cnt)) {
        Parser::Sym y;
        const bool end = ps.step<true>(ftotal, y);
        if (y.dc) { cur = pframe ? y.e : cur + y.e; o_blk = o; }
        if (y.coded && y.at < 64u) {                     // (a parked lane never sees a coded symbol)
#pragma unroll
            for (int i = 0; i < 7; i++) q[i] = q[i + 1];
            q[7] = y.at | ((uint32_t)y.e << 16);
            o++;
            if ((o & 7u) == 0u && o <= o_end) st_global_v8(sym + o - 8, q);   // never overflows on conforming streams
        }
        if (end && k < cnt) {
            bi[k] = make_uint2(min(o_blk, o_end - 1u), ((uint32_t)cur & 0xFFFFu) | ((min(o, o_end) - min(o_blk, o_end)) << 16));
            if (++k == cnt) {
                ps.park();
                if (valid) seg_dc[g] = pframe ? 0u : ((uint32_t)cur & 0xFFFFu);
            }
        }
    }
    if (valid && cnt == 0) seg_dc[g] = 0u;
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
        for (uint32_t b = first + cnt; b < sd.nb; b++) blk_info[sd.block_base + b] = make_uint2(BLK_NO_SEG, 0);
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
cudaError_t launch_entropy_sync(const EntropyJob& j, cudaStream_t s) {
    if (j.n_sync_tiles == 0) return cudaSuccess;
    k_entropy_sync<<<j.n_sync_tiles, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_sync_tiles, j.d_seg_entry,
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
    if (j.n_write_tiles == 0) return cudaSuccess;
    k_entropy_index<<<j.n_write_tiles, ENT_TPB, 0, s>>>(j.d_payload, j.d_streams, j.d_write_tiles, j.d_seg_entry,
                                                        j.d_seg_cnt, j.d_seg_first, j.d_seg_dc, j.d_blk_info, j.d_sym,
                                                        j.sym_seg0, j.d_fixups + 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_entropy_dcscan<<<(j.n_streams + 3) / 4, 128, 0, s>>>(j.d_streams + j.stream_lo, j.n_streams, j.d_seg_dc);
    return cudaGetLastError();
}

}  // namespace mj
