// decode.cu -- block-parallel final decode (sm_100a): a warp per tile of 32 consecutive block positions.
//
// After entropy.cu has produced the record lists (one 32-bit record per symbol step, every block starting with
// its DC record) and the tile descriptors (where the records of 32 consecutive blocks of a plane lie), every tile
// can be decoded independently and without any bit-serial work:
//
//   k_decode_coef   lossless_decode() output, LIB/decoder/lossless_decode.c:60-135: a warp scatters its tile's
//                   records into 128-byte shared-memory slots (zeroed = the memset of :77-78, or preloaded with the
//                   previous frame's coefficients for P frames, :90-92,121-123), dequantising as it scatters
//                   zig-zag -> natural order (:122-126); the 32 blocks leave as one contiguous, coalesced 4 KB run.
//   k_decode_intra  the whole reference loop body, LIB/decoder/mjpeg423_decoder.c:110-124, for intra-only ranges:
//                   records -> BGRA.  The column pass of the IDCT (idct.c:41-109) is LINEAR before its rounding
//                   shift, so it is never run as a butterfly: whichever lane holds a record multiplies the
//                   dequantised coefficient by its four basis factors and adds them to the owner's workspace with
//                   shared-memory reductions; only the row pass (idct.c:116-180) is a butterfly per block.
//   k_decode_gop    the same for ranges with P frames, a GOP at a time (coefficient slots carried between frames).
// Coefficients and samples never touch HBM: the kernels read the records (4 bytes per symbol step) and write
// 4 bytes per pixel (ycbcr_to_rgb.c:26-49).
// (LIB = /root/reference/core0/software/common/libs/mjpeg423.)
#include <mutex>

#include "common.cuh"
#include "runtime.h"

namespace mj {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ uint32_t lanemask_le() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void red_add(uint32_t saddr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// One run of a tile's records (a contiguous range of one segment's region) and the DC predictor of its segment.
struct Run { uint32_t a, n, dcb; };

// Walks the runs of a tile after the first two (TILE_MORE): run i (2 <= i < nruns) is segment slow.x + i - 1.
__device__ __forceinline__ Run later_run(const uint4& slow, uint32_t i, const uint32_t* __restrict__ seg_nrec,
                                         const uint32_t* __restrict__ seg_dc, uint32_t seg0) {
    const uint32_t s = slow.x + i - 1u;
    Run r;
    r.a = (s - seg0) * REC_STRIDE;
    r.n = i + 1u == slow.y ? slow.z : __ldg(seg_nrec + s);
    r.dcb = __ldg(seg_dc + s);
    return r;
}

// ---- coefficient planes ------------------------------------------------------------------------------
// One warp per tile; a CTA holds 4 tiles.  Slots use the XOR swizzle of idct_colour.cu (16-byte chunk r of slot t at
// chunk r ^ (t & 7)), so the cooperative copy-in / copy-out is bank-conflict free.
constexpr int DEC_TPB = 128;

__device__ __forceinline__ uint32_t slot_off(uint32_t t, uint32_t n) {
    return t * 128u + ((((n >> 3) ^ t) & 7u) << 4) + ((n & 7u) << 1);
}

__global__ void __launch_bounds__(DEC_TPB)
k_decode_coef(const StreamDesc* __restrict__ streams, const uint32_t* __restrict__ stream_ids, uint32_t n_ids,
              uint32_t groups, const TileDesc* __restrict__ tiles, uint32_t stream0, const uint32_t* __restrict__ rec,
              const uint32_t* __restrict__ seg_nrec, const uint32_t* __restrict__ seg_dc, uint32_t seg0,
              const int16_t* __restrict__ quant, int16_t* coef) {
    __shared__ __align__(128) uint8_t slots[DEC_TPB * 128];
    __shared__ uint32_t s_zq[64];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t id = blockIdx.x / groups, grp = blockIdx.x - id * groups;
    const uint32_t sid = stream_ids[id];
    const StreamDesc sd = streams[sid];
    if (t < 64) {
        const uint32_t n = c_zigzag[t];
        s_zq[t] = n | ((uint32_t)(uint16_t)quant[sd.quant_id * 64 + n] << 16);
    }
    __syncthreads();
    const uint32_t tpp = (sd.nb + 31u) / 32u, T = grp * 4u + warp;
    if (T >= tpp) return;
    const uint32_t b0 = T * 32u, nblk = min(32u, sd.nb - b0);
    uint8_t* ws = slots + warp * 32u * 128u;                          // the warp's 32 slots
    uint4* dst = reinterpret_cast<uint4*>(coef + ((size_t)sd.block_base + b0) * 64);
    const bool pf = sd.ptype != 0;
    if (pf) {       // the previous frame's coefficients (coalesced, swizzled) -- the P-frame state
        const uint4* src = reinterpret_cast<const uint4*>(coef + ((size_t)sd.prev_base + b0) * 64);
        for (uint32_t i = lane; i < nblk * 8u; i += 32u) {
            const uint32_t blk = i >> 3, row = i & 7u;
            *reinterpret_cast<uint4*>(ws + blk * 128u + ((row ^ (blk & 7u)) << 4)) = src[i];
        }
    } else {
#pragma unroll
        for (uint32_t r = 0; r < 8; r++) *reinterpret_cast<uint4*>(ws + lane * 128u + ((r ^ (lane & 7u)) << 4)) = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    const TileDesc* td = tiles + (size_t)(sid - stream0) * tpp + T;
    const uint4 fast = __ldg(&td->fast);
    uint4 slow = make_uint4(0u, 2u, 0u, 0u);
    if (fast.y & TILE_MORE) slow = __ldg(&td->slow);
    const uint32_t nruns = (fast.y & TILE_MORE) ? slow.y : 2u;
    uint32_t ord = 0;
    for (uint32_t i = 0; i < nruns; i++) {
        Run r;
        if (i == 0) { r.a = fast.x; r.n = fast.z & 0xFFFFu; r.dcb = fast.w & 0xFFFFu; }
        else if (i == 1) { r.a = fast.y & ~TILE_MORE; r.n = fast.z >> 16; r.dcb = fast.w >> 16; }
        else r = later_run(slow, i, seg_nrec, seg_dc, seg0);
        for (uint32_t v0 = 0; v0 < r.n; v0 += 32u) {
            const uint32_t v = v0 + lane;
            const uint32_t e = v < r.n ? __ldg(rec + r.a + v) : REC_NONE;
            const uint32_t bal = __ballot_sync(FULL_MASK, (e & REC_DC) != 0u);
            const uint32_t owner = (ord + __popc(bal & lanemask_le()) - 1u) & 31u;
            ord += __popc(bal);
            if (rec_valid(e)) {
                const uint32_t z = s_zq[e & 63u];
                int amp = (int)e >> 16;
                if (e & REC_DC) amp = (int)(int16_t)(amp + (int)r.dcb);          // absolute DC level (I) / the delta (P)
                int16_t* d = reinterpret_cast<int16_t*>(ws + slot_off(owner, z & 0xFFFFu));
                const int x = amp * (int)(z >> 16);
                if (pf) *d = (int16_t)(*d + x);                                  // lossless_decode.c:91,122
                else *d = (int16_t)x;                                            // :94-95,125
            }
            __syncwarp();      // (a block's records may be spread over two chunks: keep read-modify-writes ordered)
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < nblk * 8u; i += 32u) {       // 4 KB contiguous per warp
        const uint32_t blk = i >> 3, row = i & 7u;
        dst[i] = *reinterpret_cast<const uint4*>(ws + blk * 128u + ((row ^ (blk & 7u)) << 4));
    }
}

// ---- linear column pass ---------------------------------------------------------------------------------
// idct8<11> (common.cuh) computes out[k] = (E[k] + O[k] + RND) >> 11, out[7-k] = (E[k] - O[k] + RND) >> 11, k = 0..3,
// where E is a Z/2^32-linear function of the even inputs (rows 0, 2, 4, 6 of the column) and O of the odd ones:
// every operation before the shift is a wrap-around int32 multiply or add (MULTIPLY is a plain multiply,
// LIB/common/dct_math.h:76).  idct8_eo() returns E and O for one input vector; run on unit vectors it gives the factors
// M[r][k] with which a coefficient in row r contributes to E[k] (r even) or O[k] (r odd).
__device__ __forceinline__ void idct8_eo(int i0, int i1, int i2, int i3, int i4, int i5, int i6, int i7, int (&E)[4], int (&O)[4]) {
    int z1 = (i2 + i6) * 4433;
    int tmp2 = z1 + i6 * -15137;
    int tmp3 = z1 + i2 * 6270;
    int tmp0 = (int)((unsigned)(i0 + i4) << 13);
    int tmp1 = (int)((unsigned)(i0 - i4) << 13);
    E[0] = tmp0 + tmp3; E[3] = tmp0 - tmp3; E[1] = tmp1 + tmp2; E[2] = tmp1 - tmp2;
    int a0 = i7, a1 = i5, a2 = i3, a3 = i1;
    int y1 = a0 + a3, y2 = a1 + a2, y3 = a0 + a2, y4 = a1 + a3;
    int y5 = (y3 + y4) * 9633;
    a0 *= 2446; a1 *= 16819; a2 *= 25172; a3 *= 12299;
    y1 *= -7373; y2 *= -20995;
    y3 = y3 * -16069 + y5;
    y4 = y4 * -3196 + y5;
    O[3] = a0 + y1 + y3; O[2] = a1 + y2 + y4; O[1] = a2 + y2 + y3; O[0] = a3 + y1 + y4;
}

// ---- fully fused, intra-only: records -> BGRA ---------------------------------------------------------------
// PERSISTENT kernel: one CTA of DI_TPB threads per SM (all the shared memory of the SM), every WARP loops over warp
// tiles of 32 consecutive block positions of one frame, handed out dynamically inside the CTA.  Warps never
// synchronise with each other after the table set-up.
//
// Shared memory:
//   ws     16 granule rows of (T + 1) x 16 bytes (T = DI_TPB; granule row g of thread t at (g*(T+1) + t)*16: conflict-
//          free for 128-bit access).  The column pass's results of the thread's block BEFORE the rounding shift, as even
//          and odd parts: row 4k + h = E[k][4h .. 4h+3] (k = 0..3: workspace rows k and 7-k; h = column half), row
//          4k + 2 + h = O[k][4h .. 4h+3].  The extra granule per row rotates the banks by 4 from one row to the next, so
//          the 16 (parity, column) targets of ONE block's records fall into 16 different banks.  "Clean" = E holds the
//          rounding constant, O zero; the row pass cleans what it reads, so every plane starts on a clean workspace.
//   stash  words [32][T]   word p*16 + 2r + h = samples of plane p (Y, Cb), row r, half h
//   stage  [warps][DI_STAGE] records of the plane the warp decodes next, copied with 16-byte cp.async while the row
//          pass of the plane before runs
constexpr int DI_TPB = 512;                                       // 16 warps: 4 per scheduler
constexpr int DI_WARPS = DI_TPB / 32;
constexpr int DI_STAGE = 512;                                     // records (128 16-byte units) per warp
constexpr uint32_t DI_ROW = (DI_TPB + 1) * 16u;                   // bytes per granule row
constexpr uint32_t DI_KSTRIDE = 4u * DI_ROW;                      // bytes between E[k] and E[k+1] of a thread
constexpr int DI_OFF_STASH = 16 * DI_ROW, DI_OFF_STAGE = DI_OFF_STASH + DI_TPB * 128,
              DI_OFF_ZQ = DI_OFF_STAGE + DI_WARPS * DI_STAGE * 4, DI_OFF_M = DI_OFF_ZQ + 2 * 66 * 8, DI_OFF_NEXT = DI_OFF_M + 128,
              DI_SMEM = DI_OFF_NEXT + 16;

__global__ void __launch_bounds__(DI_TPB, 1)
k_decode_intra(const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ rec, const uint32_t* __restrict__ seg_nrec,
               const uint32_t* __restrict__ seg_dc, uint32_t seg0, const int16_t* __restrict__ quant,
               uint8_t* __restrict__ out, uint32_t nb, uint32_t wb, uint32_t W, uint32_t n_frames) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_stash = reinterpret_cast<uint32_t*>(smem + DI_OFF_STASH);
    uint2* s_zq = reinterpret_cast<uint2*>(smem + DI_OFF_ZQ);                     // 2 x 66 entries (65 used)
    int4* s_M = reinterpret_cast<int4*>(smem + DI_OFF_M);
    uint32_t* s_next = reinterpret_cast<uint32_t*>(smem + DI_OFF_NEXT);
    const int t = threadIdx.x;
    const uint32_t lane = (uint32_t)t & 31u, warp = (uint32_t)t >> 5;
    auto ws_row = [&](int g) { return reinterpret_cast<uint4*>(smem + (size_t)g * DI_ROW) + t; };
    if (t == 0) *s_next = 0u;
    if (t < 132) {     // zig-zag index -> .x = byte offset of E/O[0][col] in a thread's workspace,
                       //                  .y = quant | column bit << 16 | row * 16 << 24;  entry 64: "no coefficient"
        const int tab = t / 66, k = t % 66;
        uint2 z = make_uint2(0u, 0u);
        if (k < 64) {
            const uint32_t n = c_zigzag[k], col = n & 7u, row = n >> 3;
            z.x = ((row & 1u) * 2u + (col >> 2)) * DI_ROW + (col & 3u) * 4u;
            z.y = (uint32_t)(uint16_t)quant[tab * 64 + n] | (1u << (16 + col)) | ((row * 16u) << 24);
        }
        s_zq[t] = z;
    }
    if (t < 8) {
        int in[8], E[4], O[4];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = r == t ? 1 : 0;
        idct8_eo(in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], E, O);
        s_M[t] = (t & 1) ? make_int4(O[0], O[1], O[2], O[3]) : make_int4(E[0], E[1], E[2], E[3]);
    }
    const uint4 CLEAN_E = make_uint4(1u << 10, 1u << 10, 1u << 10, 1u << 10), CLEAN_O = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int g = 0; g < 16; g++) *ws_row(g) = (g & 2) ? CLEAN_O : CLEAN_E;
    __syncthreads();
    const uint32_t tpp = (nb + 31u) / 32u;
    const uint32_t n_items = tpp * n_frames;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t ws0 = smem0 + ((uint32_t)t & ~31u) * 16u;            // workspace (row 0) of the warp's lane 0
    const uint32_t sM0 = smem0 + DI_OFF_M;
    const uint32_t sS = smem0 + DI_OFF_STAGE + warp * (DI_STAGE * 4u);   // the warp's staging buffer

    // Tiles are handed out DYNAMICALLY inside a CTA: CTA c owns tiles c, c + grid, c + 2 grid, ... and its warps take
    // the next one from a shared-memory counter.
    auto take = [&]() { return lane == 0u ? atomicAdd(s_next, 1u) : 0u; };          // raw: valid in lane 0
    struct Pos { uint32_t f, tb; };               // frame (n_frames = nothing left), tile of the frame
    auto item_pos = [&](uint32_t raw) {
        const unsigned long long k = __shfl_sync(FULL_MASK, raw, 0);
        const uint32_t i = (uint32_t)min(k * gridDim.x + blockIdx.x, (unsigned long long)n_items);
        Pos q;
        q.f = n_frames; q.tb = 0u;
        if (i < n_items) { q.f = i / tpp; q.tb = i - q.f * tpp; }
        return q;
    };
    auto desc_of = [&](const Pos& q, int p) { return tiles + ((size_t)(q.f * 3u + p) * tpp + q.tb); };
    auto load_desc = [&](const Pos& q, uint4 (&d)[3]) {
#pragma unroll
        for (int p = 0; p < 3; p++) d[p] = q.f < n_frames ? __ldg(&desc_of(q, p)->fast) : make_uint4(0u, 0u, 0u, 0u);
    };
    // Staging copies whole 16-byte units: a run of n records at record index a occupies units [0, (a % 4 + n + 3) / 4)
    // counted from the aligned index a - a % 4, and its record j lands (a % 4 + j) records behind the unit the copy
    // started with.  stage(): units [u0, u0 + nu) of the run at `a` go to staging position `sp` (in records, x4).
    auto stage = [&](uint32_t a, uint32_t u0, uint32_t nu, uint32_t sp) {
        const uint32_t* src = rec + (a & ~3u) + u0 * 4u;
#pragma unroll 1
        for (uint32_t u = lane; u < nu; u += 32u)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sS + (sp + u * 4u) * 4u), "l"(src + u * 4u) : "memory");
    };
    constexpr uint32_t CAPU = DI_STAGE / 4;                               // units the buffer holds
    // The head of a plane's records, requested one plane ahead: the first run (whole, or its first CAPU units) and, when
    // both fit, the second run behind it.
    auto units_of = [](uint32_t a, uint32_t n) { return ((a & 3u) + n + 3u) >> 2; };
    auto stage_head = [&](const uint4& d) {
        const uint32_t nA = d.z & 0xFFFFu, nB = d.z >> 16, aB = d.y & ~TILE_MORE;
        const uint32_t uA = units_of(d.x, nA), uB = units_of(aB, nB);
        stage(d.x, 0u, min(uA, CAPU), 0u);
        if (uA + uB <= CAPU) stage(aB, 0u, uB, uA * 4u);
    };

    Pos cur = item_pos(take());
    uint4 dA[3], dB[3];
    load_desc(cur, dA);
    stage_head(dA[0]);
    while (cur.f < n_frames) {
        const uint32_t raw = take();                                    // the tile after this one
        Pos nxt;
        const uint32_t f = cur.f, b = cur.tb * 32u + lane;
        const uint32_t nblk = min(32u, nb - cur.tb * 32u);
        const bool live = b < nb;
        uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;
        bool cb_flat = false;                                            // warp-uniform: every Cb block of the tile is DC-only;
        uint32_t cb_s8 = 0;                                              // its sample then waits here, not in the stash
#pragma unroll 1
        for (int p = 0; p < 3; p++) {
            const uint4 d = p == 0 ? dA[0] : p == 1 ? dA[1] : dA[2];
            const uint2* zq = s_zq + (p ? 66 : 0);
            const uint32_t nA = d.z & 0xFFFFu, nB = d.z >> 16, aA = d.x, aB = d.y & ~TILE_MORE;
            const uint32_t dcbA = d.w & 0xFFFFu, dcbB = d.w >> 16;
            const bool more = (d.y & TILE_MORE) != 0u;
            const uint32_t uA = units_of(aA, nA), uB = units_of(aB, nB);
            const bool fitsAB = uA + uB <= CAPU;
            cp_async_wait();                                             // the plane's staged records have landed
            __syncwarp();                                                // ... for every lane (units are copied by any lane)

            // emit(r, w0, w1): Y and Cb rows go to the stash, a Cr row completes 8 pixels.
            auto emit = [&](int r, uint32_t w0, uint32_t w1) {
                if (p < 2) {
                    s_stash[(p * 16 + 2 * r) * DI_TPB + t] = w0;
                    s_stash[(p * 16 + 2 * r + 1) * DI_TPB + t] = w1;
                } else if (live) {
                    colour_row_store(s_stash[(2 * r) * DI_TPB + t], s_stash[(2 * r + 1) * DI_TPB + t],
                                     s_stash[(16 + 2 * r) * DI_TPB + t], s_stash[(17 + 2 * r) * DI_TPB + t], w0, w1,
                                     dst + (size_t)r * W * 4);
                }
            };
            // What follows the record phase of every plane: the staging buffer is free again, so the next plane's
            // records are requested (they land while this plane's row pass runs); behind the luminance records the
            // next tile's descriptors are requested as well.
            auto pipeline = [&]() {
                __syncwarp();                                            // every lane is done with the staged records
                if (p == 0) {
                    nxt = item_pos(raw);
                    load_desc(nxt, dB);
                } else if (p == 1) {
#pragma unroll
                    for (int q = 0; q < 3; q++) {                        // pull the next tile's records into L2
                        const uint32_t nA2 = dB[q].z & 0xFFFFu, nB2 = dB[q].z >> 16;
                        if (lane * 32u < nA2) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + dB[q].x + lane * 32u));
                        if (lane * 32u < nB2) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (dB[q].y & ~TILE_MORE) + lane * 32u));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 3; q++) dA[q] = dB[q];
                }
                stage_head(p == 0 ? dA[1] : p == 1 ? dA[2] : dA[0]);    // (after the rotation: the next tile's luminance)
            };

            // ---- a plane whose blocks are all DC-only: both passes collapse to (4*dc + 16) >> 5 (see idct_block()) ----
            bool dc_only = false;
            uint32_t rec0 = 0u;
            if (nA + nB == nblk && !more && fitsAB) {
                rec0 = lane < nblk ? lds32(sS + ((lane < nA ? (aA & 3u) + lane : uA * 4u + (aB & 3u) + (lane - nA)) << 2)) : REC_DC;
                dc_only = __all_sync(FULL_MASK, (rec0 & REC_DC) != 0u);
            }
            if (dc_only) {
                const int level = (int)(int16_t)(((int)rec0 >> 16) + (int)(lane < nA ? dcbA : dcbB));
                const int dc_coef = (int)(int16_t)(level * (int)(zq[0].y & 0xFFFFu));    // lossless_decode.c:94-95
                pipeline();
                const uint32_t s8 = clamp255(((dc_coef << 2) + 16) >> 5);
                const uint32_t v = s8 * 0x01010101u;
                if (p == 1) { cb_flat = true; cb_s8 = s8; continue; }    // kept in a register until Cr is known
                if (p == 2 && cb_flat) {
                    // Flat Cb and Cr blocks: the chroma terms are per-block constants.
                    if (live) {
                        const FlatChroma fc(cb_s8, s8);
#pragma unroll 2
                        for (int r = 0; r < 8; r++)
                            fc.row_store(s_stash[(2 * r) * DI_TPB + t], s_stash[(2 * r + 1) * DI_TPB + t], dst + (size_t)r * W * 4);
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < 8; r++) emit(r, v, v);
                }
                continue;
            }

            // ---- column pass: every record adds its four basis products to the owner's workspace --------------
            uint32_t ord = 0u, m_bits = 0u;                              // block ordinal inside the tile; columns in use << 16
            auto put = [&](uint32_t e, uint32_t dcb) {
                const uint32_t bal = __ballot_sync(FULL_MASK, (e & REC_DC) != 0u);
                const uint32_t owner = (ord + __popc(bal & lanemask_le()) - 1u) & 31u;
                ord += __popc(bal);
                const uint2 z = zq[min(e & 0xFFu, 64u)];                 // (entry 64: no coefficient -- quant 0, no column)
                int amp = (int)e >> 16;
                if (e & REC_DC) amp += (int)dcb;                         // absolute DC level (mod 2^16)
                const int x = (int)(int16_t)((int)(int16_t)amp * (int)(z.y & 0xFFFFu));   // dequantised, int16 (:94-95,125)
                int4 m;
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(sM0 + (z.y >> 24)));
                const uint32_t addr = ws0 + owner * 16u + z.x;
                red_add(addr, x * m.x);
                red_add(addr + DI_KSTRIDE, x * m.y);
                red_add(addr + 2u * DI_KSTRIDE, x * m.z);
                red_add(addr + 3u * DI_KSTRIDE, x * m.w);
                m_bits |= z.y;
            };
            // The tile's runs, one after the other (ONE copy of the loop body: the kernel has to fit the instruction
            // cache).  The first `have` units of a run are staged at position sp already, the rest is fetched here.
            uint4 slow = make_uint4(0u, nB ? 2u : 1u, 0u, 0u);
            if (more) slow = __ldg(&desc_of(cur, p)->slow);
#pragma unroll 1
            for (uint32_t i = 0; i < slow.y; i++) {
                Run r;
                uint32_t sp = 0u, have = 0u;
                if (i == 0) { r.a = aA; r.n = nA; r.dcb = dcbA; have = min(uA, CAPU); }
                else if (i == 1) { r.a = aB; r.n = nB; r.dcb = dcbB; if (fitsAB) { sp = uA * 4u; have = uB; } }
                else r = later_run(slow, i, seg_nrec, seg_dc, seg0);
                if (r.n == 0u) continue;
                const uint32_t lead = r.a & 3u, nu = units_of(r.a, r.n);
                uint32_t u0 = 0u;
#pragma unroll 1
                for (;;) {
                    if (have) {
                        // records [j0, j1) of the run are staged; record j sits at staging position sp + lead + j - 4 u0
                        const uint32_t j1 = min(r.n, (u0 + have) * 4u - lead), s0 = sp + lead - u0 * 4u;
#pragma unroll 1
                        for (uint32_t j = u0 ? u0 * 4u - lead : 0u; j < j1; j += 32u)
                            put(j + lane < j1 ? lds32(sS + ((s0 + j + lane) << 2)) : REC_NONE, r.dcb);
                        u0 += have;
                        if (u0 >= nu) break;
                    }
                    have = min(nu - u0, CAPU);
                    sp = 0u;
                    __syncwarp();
                    stage(r.a, u0, have, 0u);
                    cp_async_wait();
                    __syncwarp();
                }
            }
            const uint32_t m_all = warp_or(m_bits >> 16) | 1u;           // warp-uniform from here on (column 0: the DC)
            pipeline();                                                   // (its __syncwarp also orders the reductions)
            if (p == 2 && cb_flat) {                                      // Cr needs the IDCT after all: materialise the flat Cb rows
                const uint32_t v = cb_s8 * 0x01010101u;
#pragma unroll 1
                for (int r = 0; r < 16; r++) s_stash[(16 + r) * DI_TPB + t] = v;
                cb_flat = false;
            }
            // ---- row pass (idct.c:116-180): workspace rows k and 7-k are E[k] + O[k] and E[k] - O[k], descaled -------
            const bool high_half = (m_all & 0xF0u) != 0u;
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                uint4* ge = ws_row(4 * k);
                uint4* go = ws_row(4 * k + 2);
                const uint4 el = *ge, ol = *go;
                *ge = CLEAN_E; *go = CLEAN_O;
                int o0[8], o1[8];
                const int a0 = ((int)el.x + (int)ol.x) >> 11, a1 = ((int)el.y + (int)ol.y) >> 11, a2 = ((int)el.z + (int)ol.z) >> 11,
                          a3 = ((int)el.w + (int)ol.w) >> 11;
                const int b0 = ((int)el.x - (int)ol.x) >> 11, b1 = ((int)el.y - (int)ol.y) >> 11, b2 = ((int)el.z - (int)ol.z) >> 11,
                          b3 = ((int)el.w - (int)ol.w) >> 11;
                if (high_half) {
                    uint4* he = ws_row(4 * k + 1);
                    uint4* ho = ws_row(4 * k + 3);
                    const uint4 eh = *he, oh = *ho;
                    *he = CLEAN_E; *ho = CLEAN_O;
                    idct8<18>(a0, a1, a2, a3, ((int)eh.x + (int)oh.x) >> 11, ((int)eh.y + (int)oh.y) >> 11, ((int)eh.z + (int)oh.z) >> 11,
                              ((int)eh.w + (int)oh.w) >> 11, o0);
                    idct8<18>(b0, b1, b2, b3, ((int)eh.x - (int)oh.x) >> 11, ((int)eh.y - (int)oh.y) >> 11, ((int)eh.z - (int)oh.z) >> 11,
                              ((int)eh.w - (int)oh.w) >> 11, o1);
                } else {
                    idct8<18>(a0, a1, a2, a3, 0, 0, 0, 0, o0);
                    idct8<18>(b0, b1, b2, b3, 0, 0, 0, 0, o1);
                }
                emit(k, pack4_sat_u8(o0[0], o0[1], o0[2], o0[3]), pack4_sat_u8(o0[4], o0[5], o0[6], o0[7]));
                emit(7 - k, pack4_sat_u8(o1[0], o1[1], o1[2], o1[3]), pack4_sat_u8(o1[4], o1[5], o1[6], o1[7]));
            }
        }
        cur = nxt;
    }
    cp_async_wait();
}

// ---- ranges with P frames: a GOP at a time ------------------------------------------------------------------
// LIB/decoder/lossless_decode.c:90-92,121-123: every decoded value of a P frame is ADDED to the previous frame's
// coefficient (int16, wrapping), so the coefficients themselves are the inter-frame state and the column pass is a
// butterfly over them.  A work item is (tile, GOP): the warp walks the GOP's frames in order for one tile, so the
// state never leaves the warp: the accumulated column masks stay in registers, and the coefficient slots of a plane
// that has held an AC coefficient since the I frame are parked between frames in a per-warp scratch area (12 KB per
// warp: it stays in L2) and brought back with cp.async.  gop_first[g] .. gop_first[g + 1] are the (chunk-relative)
// frames of GOP g.
//
// Shared memory, in 16-byte granules interleaved by thread (granule g of thread t at (g*T + t)*16, T = FUSED_TPB):
//   ws    granules 0..15   granule cp*4 + r/2 = pass-1 outputs {ws[r][2cp], ws[r][2cp+1], ws[r+1][2cp], ws[r+1][2cp+1]}
//   coef  granules 8..15   granule 8+c = column c of the block, rows 0..7 as int16.  ALIASES the upper half of
//                          ws: pass 1 consumes columns 2cp, 2cp+1 in iteration cp and only then writes granules
//                          4cp..4cp+3, so every coefficient granule is dead before it is overwritten.
//   stage granules 0..7    (lower half of ws, dead while a plane is scattered) the plane's records, 1024 per warp
//   stash words [32][T]    word p*16 + 2r + h = samples of plane p (Y, Cb), row r, half h
constexpr int FUSED_TPB = 576;                                   // 18 warps x 384 B/thread = 216 KB of the SM's 227 KB
constexpr int FUSED_SMEM = FUSED_TPB * (256 + 128) + 2 * 64 * 8 + 16;     // + the CTA's tile counter

__global__ void __launch_bounds__(FUSED_TPB, 1)
k_decode_gop(const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ rec, const uint32_t* __restrict__ seg_nrec,
             const uint32_t* __restrict__ seg_dc, uint32_t seg0, const int16_t* __restrict__ quant, uint8_t* __restrict__ out,
             uint32_t nb, uint32_t wb, uint32_t W, uint32_t n_frames, const uint32_t* __restrict__ gop_first, uint32_t n_gops,
             uint4* __restrict__ state) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* s_ws = reinterpret_cast<uint4*>(smem);                                   // granules 0..15
    uint8_t* s_coef = smem + 8 * FUSED_TPB * 16;                                    // granules 8..15
    uint32_t* s_stash = reinterpret_cast<uint32_t*>(smem + FUSED_TPB * 256);
    uint2* s_zq = reinterpret_cast<uint2*>(smem + FUSED_TPB * 384);                 // 2 x 64 entries
    uint32_t* s_next = reinterpret_cast<uint32_t*>(smem + FUSED_TPB * 384 + 1024);  // items handed out so far
    const int t = threadIdx.x;
    if (t == 0) *s_next = 0u;
    if (t < 128) {     // zig-zag index -> .x = transposed slot offset | quant << 16, .y = column bit | (row >= 1) column bit << 8
        const int tab = t >> 6, k = t & 63;
        const uint32_t n = c_zigzag[k], col = n & 7u, row = n >> 3;
        s_zq[t] = make_uint2((col * (FUSED_TPB * 16u) + row * 2u) | ((uint32_t)(uint16_t)quant[tab * 64 + n] << 16),
                             (1u << col) | ((row ? 1u : 0u) << (8 + col)));
    }
    __syncthreads();
    uint8_t* my_coef = s_coef + t * 16;
    const uint32_t tpp = (nb + 31u) / 32u;
    const uint32_t n_items = tpp * n_gops;                                // work items: (tile, GOP)
    const uint32_t lane = (uint32_t)t & 31u;
    auto take = [&]() { return lane == 0u ? atomicAdd(s_next, 1u) : 0u; };
    struct Pos { uint32_t f, tb, fend; };         // frame (chunk-relative; n_frames = nothing left), tile of the frame, end of the item
    auto item_pos = [&](uint32_t raw) {
        const unsigned long long k = __shfl_sync(FULL_MASK, raw, 0);
        const uint32_t i = (uint32_t)min(k * gridDim.x + blockIdx.x, (unsigned long long)n_items);
        Pos q;
        q.f = q.fend = n_frames; q.tb = 0u;
        if (i < n_items) {
            const uint32_t g = i / tpp;
            q.tb = i - g * tpp;
            q.f = __ldg(gop_first + g); q.fend = __ldg(gop_first + g + 1);
        }
        return q;
    };
    auto desc_of = [&](const Pos& q, int p) { return tiles + ((size_t)(q.f * 3u + p) * tpp + q.tb); };
    auto load_desc = [&](const Pos& q, uint4 (&d)[3]) {
#pragma unroll
        for (int p = 0; p < 3; p++) d[p] = q.f < n_frames ? __ldg(&desc_of(q, p)->fast) : make_uint4(0u, 0u, 0u, 0u);
    };
    uint8_t* warp_coef = s_coef + ((uint32_t)t & ~31u) * 16u;            // slot of lane 0 of this warp
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + ((uint32_t)t & ~31u) * 16u + lane * 4u;
    // Copy slots [v0, v0 + m) of run A | run B into the warp's staging granules (m <= 1024); slot i sits in piece
    // i / 128 (one granule row of the warp = 512 bytes), copied and read back by lane i & 31.
    auto stage = [&](uint32_t aA, uint32_t nA, uint32_t aB, uint32_t v0, uint32_t m) {
        for (uint32_t i = lane; i < m; i += 32u) {
            const uint32_t v = v0 + i;
            cp_async4(stage0 + (i >> 7) * (FUSED_TPB * 16u) + ((i & 127u) - lane) * 4u, rec + (v < nA ? aA + v : aB + (v - nA)));
        }
        cp_async_wait();
    };
    auto staged = [&](uint32_t i0) { return lds32(stage0 + (i0 >> 7) * (FUSED_TPB * 16u) + (i0 & 127u) * 4u); };   // slot i0 + lane

    Pos cur = item_pos(take()), nxt;
    bool cur_first = true, nxt_first;             // first frame of its item (an I frame: no state before it)
    uint4 dA[3], dB[3];
    load_desc(cur, dA);
    // inter-frame state of the item: accumulated masks (warp-uniform) of the planes
    uint32_t mY = 1u, mB = 1u, mR = 1u;
    uint4* my_state = state + ((size_t)(blockIdx.x * (FUSED_TPB / 32) + (t >> 5)) * 3u * 8u) * 32u + lane;

    for (; cur.f < n_frames; cur = nxt, cur_first = nxt_first) {
        nxt_first = !(cur.f + 1u < cur.fend);
        if (nxt_first) nxt = item_pos(take());
        else { nxt.f = cur.f + 1u; nxt.tb = cur.tb; nxt.fend = cur.fend; }
        load_desc(nxt, dB);                                               // lands while this frame's tile is decoded
        const uint32_t f = cur.f;
        const uint32_t b = cur.tb * 32u + lane;
        const uint32_t nblk = min(32u, nb - cur.tb * 32u);
        const bool live = b < nb;
        uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;

        if (!cur_first) {
            // P frame whose positions are all unchanged (only DC records, all of them zero deltas, in all three planes
            // -- what a static background is coded as): the pixels are the previous frame's, which this lane wrote.
            bool same = true;
#pragma unroll
            for (int p = 0; p < 3; p++) {
                const uint32_t nA = dA[p].z & 0xFFFFu, tot = nA + (dA[p].z >> 16);
                bool ok = tot == nblk && !(dA[p].y & TILE_MORE);
                if (ok && lane < nblk) {
                    const uint32_t e = __ldg(rec + (lane < nA ? dA[p].x + lane : (dA[p].y & ~TILE_MORE) + (lane - nA)));
                    ok = (e & REC_DC) != 0u && (e >> 16) == 0u;
                }
                same = same && ok;
            }
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
#pragma unroll
                for (int q = 0; q < 3; q++) dA[q] = dB[q];
                continue;
            }
        }
        bool cb_flat = false;                                            // warp-uniform: every Cb block of the tile is DC-only;
        uint32_t cb_s8 = 0;                                              // its sample then waits here, not in the stash
#pragma unroll 1
        for (int p = 0; p < 3; p++) {
            const uint4 d = p == 0 ? dA[0] : p == 1 ? dA[1] : dA[2];
            const uint2* zq = s_zq + (p ? 64 : 0);
            const uint32_t nA = d.z & 0xFFFFu, nB = d.z >> 16, tot = nA + nB;
            const uint32_t dcbA = d.w & 0xFFFFu, dcbB = d.w >> 16;
            const bool more = (d.y & TILE_MORE) != 0u;
            const uint32_t macc = cur_first ? 1u : (p == 0 ? mY : p == 1 ? mB : mR);   // columns occupied since the item's I frame
            const bool has_state = (macc & 0xFFFEu) != 0u;               // the plane's slots are parked in the scratch area

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
            // ---- the plane's coefficient slots: zeroed (the memset of :77-78) or, for a P frame, holding the previous
            // frame's coefficients (a plane without parked slots holds only its DC coefficient: kept in the state too) ----
            if (has_state) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(my_coef);
#pragma unroll
                for (int c = 0; c < 8; c++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + c * (FUSED_TPB * 16)),
                                 "l"(my_state + (p * 8 + c) * 32) : "memory");
                cp_async_wait();
            } else {
#pragma unroll
                for (int c = 1; c < 8; c++) *reinterpret_cast<uint4*>(my_coef + c * (FUSED_TPB * 16)) = make_uint4(0, 0, 0, 0);
                uint4 c0 = make_uint4(0, 0, 0, 0);
                if (!cur_first) c0.x = my_state[(p * 8) * 32].x & 0xFFFFu;   // the DC coefficient of the previous frame
                *reinterpret_cast<uint4*>(my_coef) = c0;
            }
            __syncwarp();
            // Whichever lane holds a record dequantises it and adds / stores it into the slot of the lane that owns the
            // block: no lane waits for the longest list of the tile.
            uint32_t ord = 0u, m_bits = 1u;
            auto put = [&](uint32_t e, uint32_t dcb) {
                const uint32_t bal = __ballot_sync(FULL_MASK, (e & REC_DC) != 0u);
                const uint32_t owner = (ord + __popc(bal & lanemask_le()) - 1u) & 31u;
                ord += __popc(bal);
                if (rec_valid(e)) {
                    const uint2 z = zq[e & 63u];
                    int amp = (int)e >> 16;
                    if (e & REC_DC) amp = (int)(int16_t)(amp + (int)dcb);
                    int16_t* dp = reinterpret_cast<int16_t*>(warp_coef + owner * 16u + (z.x & 0xFFFFu));
                    const int v = amp * (int)(z.x >> 16);
                    *dp = (int16_t)(v + (cur_first ? 0 : (int)*dp));     // :91,122 (P frame: added) / :94-95,125 (I frame: stored)
                    m_bits |= z.y;
                }
            };
            auto scatter_runs = [&](uint32_t aA, uint32_t nA_, uint32_t aB, uint32_t tot_, uint32_t dA_, uint32_t dB_) {
                for (uint32_t base = 0; base < tot_; base += 1024u) {
                    const uint32_t m = min(tot_ - base, 1024u);
                    stage(aA, nA_, aB, base, m);
                    for (uint32_t i0 = 0; i0 < m; i0 += 32u) {
                        put(i0 + lane < m ? staged(i0) : REC_NONE, base + i0 + lane < nA_ ? dA_ : dB_);
                        __syncwarp();                                    // a block's records may span two chunks
                    }
                }
            };
            scatter_runs(d.x, nA, d.y & ~TILE_MORE, tot, dcbA, dcbB);
            if (more) {
                const uint4 slow = __ldg(&desc_of(cur, p)->slow);
                for (uint32_t i = 2; i < slow.y; i++) {
                    const Run r = later_run(slow, i, seg_nrec, seg_dc, seg0);
                    scatter_runs(r.a, r.n, 0u, r.n, r.dcb, r.dcb);
                }
            }
            __syncwarp();
            uint32_t m_all = warp_or(m_bits) | macc;                     // warp-uniform from here on (supersets stay valid)
            if (cur.f + 1u < cur.fend) {                                  // park the state for the item's next frame
                if (m_all & 0xFFFEu) {
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        my_state[(p * 8 + c) * 32] = *reinterpret_cast<const uint4*>(my_coef + c * (FUSED_TPB * 16));
                } else {
                    my_state[(p * 8) * 32] = *reinterpret_cast<const uint4*>(my_coef);
                }
            }
            if (p == 0) mY = m_all; else if (p == 1) mB = m_all; else mR = m_all;
            const uint32_t acm = m_all >> 8, anym = m_all & 0xFFu;

            if (((anym & 0xFEu) | (acm & 1u)) == 0) {                     // nothing outside the DC position in the whole tile
                const int dc_coef = (int)*reinterpret_cast<const int16_t*>(my_coef);
                const uint32_t s8 = clamp255(((dc_coef << 2) + 16) >> 5);
                const uint32_t v = s8 * 0x01010101u;
                if (p == 1) { cb_flat = true; cb_s8 = s8; continue; }    // kept in a register until Cr is known
                if (p == 2 && cb_flat) {
                    if (live) {
                        const FlatChroma fc(cb_s8, s8);
#pragma unroll 2
                        for (int r = 0; r < 8; r++)
                            fc.row_store(s_stash[(2 * r) * FUSED_TPB + t], s_stash[(2 * r + 1) * FUSED_TPB + t], dst + (size_t)r * W * 4);
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < 8; r++) emit(r, v, v);
                }
                continue;
            }
            if (p == 2 && cb_flat) {                                      // Cr needs the IDCT after all: materialise the flat Cb rows
                const uint32_t v = cb_s8 * 0x01010101u;
#pragma unroll 1
                for (int r = 0; r < 16; r++) s_stash[(16 + r) * FUSED_TPB + t] = v;
                cb_flat = false;
            }
            // ---- pass 1: columns, two at a time (idct.c:41-109) --------------------------------------------------
            const bool high_half = (anym & 0xF0u) != 0;
            const int npair = high_half ? 4 : 2;
#pragma unroll 2
            for (int cp = 0; cp < npair; cp++) {
                const uint4 c0 = *reinterpret_cast<const uint4*>(my_coef + (2 * cp) * (FUSED_TPB * 16));
                const uint4 c1 = *reinterpret_cast<const uint4*>(my_coef + (2 * cp + 1) * (FUSED_TPB * 16));
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
                    s_ws[(cp * 4 + r / 2) * FUSED_TPB + t] = make_uint4((uint32_t)o0[r], (uint32_t)o1[r], (uint32_t)o0[r + 1], (uint32_t)o1[r + 1]);
            }
            // ---- pass 2: rows (idct.c:116-180) ------------------------------------------------------------------
#pragma unroll 1
            for (int r = 0; r < 8; r += 2) {                          // two rows per iteration: one LDS.128 per column pair
                const uint4* g = s_ws + (r >> 1) * FUSED_TPB + t;
                const uint4 a = g[0], bq = g[4 * FUSED_TPB];
                int o0[8], o1[8];
                if (high_half) {
                    const uint4 cq = g[8 * FUSED_TPB], dq = g[12 * FUSED_TPB];
                    idct8<18>((int)a.x, (int)a.y, (int)bq.x, (int)bq.y, (int)cq.x, (int)cq.y, (int)dq.x, (int)dq.y, o0);
                    idct8<18>((int)a.z, (int)a.w, (int)bq.z, (int)bq.w, (int)cq.z, (int)cq.w, (int)dq.z, (int)dq.w, o1);
                } else {
                    idct8<18>((int)a.x, (int)a.y, (int)bq.x, (int)bq.y, 0, 0, 0, 0, o0);
                    idct8<18>((int)a.z, (int)a.w, (int)bq.z, (int)bq.w, 0, 0, 0, 0, o1);
                }
                emit(r, pack4_sat_u8(o0[0], o0[1], o0[2], o0[3]), pack4_sat_u8(o0[4], o0[5], o0[6], o0[7]));
                emit(r + 1, pack4_sat_u8(o1[0], o1[1], o1[2], o1[3]), pack4_sat_u8(o1[4], o1[5], o1[6], o1[7]));
            }
            __syncwarp();                                                // the staging granules are this plane's workspace until here
        }
#pragma unroll
        for (int q = 0; q < 3; q++) dA[q] = dB[q];
    }
}

// ---- launchers -----------------------------------------------------------------------------------------------
cudaError_t launch_decode_coef(const EntropyJob& j, const uint32_t* d_stream_ids, uint32_t n_ids, uint32_t nb,
                               const int16_t* d_quant, int16_t* d_coef, cudaStream_t s) {
    if (n_ids == 0 || nb == 0) return cudaSuccess;
    const uint32_t groups = ((nb + 31u) / 32u + 3u) / 4u;
    if ((uint64_t)groups * n_ids > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    k_decode_coef<<<groups * n_ids, DEC_TPB, 0, s>>>(j.d_streams, d_stream_ids, n_ids, groups, j.d_tiles, j.stream_lo, j.d_rec,
                                                     j.d_seg_nrec, j.d_seg_dc, j.seg0, d_quant, d_coef);
    return cudaGetLastError();
}

// Per-device set-up of the persistent kernels (dynamic shared memory opt-in, SM count): once per device, thread-safe.
static int fused_grid_limit(cudaError_t& err) {
    static std::once_flag once[64];
    static int n_sm[64];
    static cudaError_t status[64];
    int dev = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return 0;
    if (dev < 0 || dev >= 64) { err = cudaErrorInvalidDevice; return 0; }
    std::call_once(once[dev], [dev]() {
        cudaError_t e = cudaFuncSetAttribute(k_decode_intra, cudaFuncAttributeMaxDynamicSharedMemorySize, DI_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_decode_gop, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess && n_sm[dev] > FUSED_MAX_CTAS) n_sm[dev] = FUSED_MAX_CTAS;   // the scratch area is sized for this many
        status[dev] = e;
    });
    err = status[dev];
    return n_sm[dev];
}

// gop_first == nullptr: an intra-only range (one work item per tile and frame).  Otherwise the range holds P frames:
// d_gop_first[0 .. n_gops] are the chunk-relative first frames of its GOPs (+ the end), d_state the per-warp scratch of
// FUSED_STATE_BYTES (see k_decode_gop).
cudaError_t launch_decode_fused(const EntropyJob& j, const int16_t* d_quant, void* d_out, uint32_t n_frames,
                                uint32_t W, uint32_t H, const uint32_t* d_gop_first, uint32_t n_gops, void* d_state,
                                cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    cudaError_t e;
    const int n_sm = fused_grid_limit(e);
    if (e != cudaSuccess) return e;
    const uint32_t wb = W / 8, nb = wb * (H / 8);
    const uint64_t n_items = (uint64_t)((nb + 31) / 32) * (d_gop_first ? n_gops : n_frames);
    const int wpc = (d_gop_first ? FUSED_TPB : DI_TPB) / 32;
    const uint64_t want = (n_items + wpc - 1) / wpc;
    const unsigned grid = (unsigned)(want < (uint64_t)n_sm ? want : (uint64_t)n_sm);   // persistent: one CTA per SM
    if (d_gop_first)
        k_decode_gop<<<grid, FUSED_TPB, FUSED_SMEM, s>>>(j.d_tiles, j.d_rec, j.d_seg_nrec, j.d_seg_dc, j.seg0, d_quant, (uint8_t*)d_out,
                                                         nb, wb, W, n_frames, d_gop_first, n_gops, (uint4*)d_state);
    else
        k_decode_intra<<<grid, DI_TPB, DI_SMEM, s>>>(j.d_tiles, j.d_rec, j.d_seg_nrec, j.d_seg_dc, j.seg0, d_quant, (uint8_t*)d_out, nb,
                                                     wb, W, n_frames);
    return cudaGetLastError();
}

}  // namespace mj
