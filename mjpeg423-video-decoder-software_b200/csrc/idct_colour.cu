// idct_colour.cu -- 8x8 integer IDCT and YCbCr->BGRA raster pack kernels (sm_100a).
//
//   k_idct         idct()          LIB/decoder/idct.c:22-181      coefficient blocks -> sample blocks
//   k_colour       ycbcr_to_rgb()  LIB/decoder/ycbcr_to_rgb.c:26-49  sample blocks -> BGRA raster
//   k_idct_colour  both, fused: the three coefficient tiles of a run of blocks go through shared memory
//                  once and leave as finished BGRA pixels (what the reference's FPGA block does behind
//                  C0/idct_ycbcr_to_rgb_accel.c:61-82), so the sample planes never touch HBM.
// (LIB = /root/reference/core0/software/common/libs/mjpeg423, C0 = /root/reference/core0/software.)
//
// Mapping: one thread per 8x8 block.  Coefficient tiles (128 blocks x 128 B, contiguous in HBM) are
// staged with 16-byte cp.async copies into shared memory, XOR-swizzled by block so that the per-thread
// row reads (LDS.128 at a 128-byte stride) are bank-conflict free.  The IDCT runs entirely in registers
// (int32, reference operation order); pixels leave as one 256-bit store per block row, i.e. one full
// 32-byte sector per lane and 1 KB contiguous per warp and row.
#include <mutex>

#include "common.cuh"
#include "runtime.h"

namespace mj {

constexpr int IDCT_TPB = 128;                       // blocks (threads) per tile
constexpr int TILE_BYTES = IDCT_TPB * 128;          // one plane's coefficient tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Stage `nblk` (<= IDCT_TPB) consecutive coefficient blocks starting at `src` into `tile`.
__device__ __forceinline__ void stage_tile(uint8_t* tile, const int16_t* src, int nblk, int t) {
    const uint8_t* g = reinterpret_cast<const uint8_t*>(src);
    const int nchunk = nblk * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int i = t + k * IDCT_TPB;                   // 16-byte chunk index: consecutive lanes, consecutive chunks
        if (i < nchunk) {
            int blk = i >> 3, row = i & 7;
            cp_async16(tile + blk * 128 + ((row ^ (blk & 7)) << 4), g + (size_t)i * 16);
        }
    }
}
__device__ __forceinline__ void load_block_rows(const uint8_t* tile, int t, uint4 (&rows)[8]) {
#pragma unroll
    for (int r = 0; r < 8; r++)
        rows[r] = *reinterpret_cast<const uint4*>(tile + t * 128 + ((r ^ (t & 7)) << 4));
}

// ---- idct(): n_blocks coefficient blocks -> n_blocks sample blocks (both block-major) ---------------
__global__ void __launch_bounds__(IDCT_TPB)
k_idct(const int16_t* __restrict__ coef, uint8_t* __restrict__ samples, size_t n_blocks) {
    __shared__ __align__(128) uint8_t tile[TILE_BYTES];
    const int t = threadIdx.x;
    const size_t b0 = (size_t)blockIdx.x * IDCT_TPB;
    const int nblk = (int)min((size_t)IDCT_TPB, n_blocks - b0);
    stage_tile(tile, coef + b0 * 64, nblk, t);
    cp_async_wait_all();
    __syncthreads();
    const bool live = t < nblk;
    uint4 rows[8];
    uint32_t ac = 0, any = 0;
    if (live) { load_block_rows(tile, t, rows); block_masks(rows, ac, any); }
    else {
#pragma unroll
        for (int r = 0; r < 8; r++) rows[r] = make_uint4(0, 0, 0, 0);
    }
    uint32_t px[16];
    idct_block(rows, warp_or(ac), warp_or(any), px);
    if (!live) return;
    uint4* dst = reinterpret_cast<uint4*>(samples + (b0 + t) * 64);
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = make_uint4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
}

// ---- ycbcr_to_rgb(): sample planes (frame-major, Y|Cb|Cr, block-major) -> BGRA raster ---------------
__global__ void __launch_bounds__(IDCT_TPB)
k_colour(const uint8_t* __restrict__ samples, uint8_t* __restrict__ out, uint32_t nb, uint32_t wb, uint32_t W, uint32_t groups) {
    const uint32_t f = blockIdx.x / groups;                  // (the frame index is folded into blockIdx.x: gridDim.y <= 65535)
    const uint32_t b = (blockIdx.x - f * groups) * IDCT_TPB + threadIdx.x;
    if (b >= nb) return;
    const uint8_t* fs = samples + (size_t)f * 3 * nb * 64;
    const uint4* yp = reinterpret_cast<const uint4*>(fs + (size_t)b * 64);
    const uint4* cbp = reinterpret_cast<const uint4*>(fs + ((size_t)nb + b) * 64);
    const uint4* crp = reinterpret_cast<const uint4*>(fs + ((size_t)2 * nb + b) * 64);
    uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {                    // 16 B = two block rows per plane
        uint4 y = __ldg(yp + k), cb = __ldg(cbp + k), cr = __ldg(crp + k);
        colour_row_store(y.x, y.y, cb.x, cb.y, cr.x, cr.y, dst + (size_t)(2 * k) * W * 4);
        colour_row_store(y.z, y.w, cb.z, cb.w, cr.z, cr.w, dst + (size_t)(2 * k + 1) * W * 4);
    }
}

// ---- fused IDCT + colour: coefficient planes (frame-major, Y|Cb|Cr) -> BGRA raster ------------------
// One plane at a time: the thread's block goes through the register IDCT and its 64 samples are written back over the
// first half of the block's own (already consumed) coefficient slot; the colour conversion then reads the three sample
// blocks back row by row.  (Keeping all three planes' samples in registers took 128 registers and ran at 20 % of the
// warp slots and 41 % of the HBM peak, profiles/r02a; this form needs about as many as k_idct.)
__global__ void __launch_bounds__(IDCT_TPB, 4)
k_idct_colour(const int16_t* __restrict__ coef, uint8_t* __restrict__ out, uint32_t nb, uint32_t wb, uint32_t W, uint32_t groups) {
    extern __shared__ __align__(128) uint8_t smem[];          // 3 x TILE_BYTES
    const int t = threadIdx.x;
    const uint32_t f = blockIdx.x / groups;
    const uint32_t b0 = (blockIdx.x - f * groups) * IDCT_TPB;
    const int nblk = (int)min((uint32_t)IDCT_TPB, nb - b0);
    const int16_t* fc = coef + (size_t)f * 3 * nb * 64;
#pragma unroll
    for (int p = 0; p < 3; p++) stage_tile(smem + p * TILE_BYTES, fc + ((size_t)p * nb + b0) * 64, nblk, t);
    cp_async_wait_all();
    __syncthreads();
    const bool live = t < nblk;
#pragma unroll 1
    for (int p = 0; p < 3; p++) {
        uint8_t* tile = smem + p * TILE_BYTES;
        uint4 rows[8];
        uint32_t ac = 0, any = 0;
        if (live) { load_block_rows(tile, t, rows); block_masks(rows, ac, any); }
        else {
#pragma unroll
            for (int r = 0; r < 8; r++) rows[r] = make_uint4(0, 0, 0, 0);
        }
        uint32_t px[16];
        idct_block(rows, warp_or(ac), warp_or(any), px);
        // samples of row pair k (rows 2k, 2k+1) -> chunk k of the thread's slot (same swizzle as the coefficients: the
        // slot is the thread's own, and all eight of its coefficient chunks are in registers by now)
#pragma unroll
        for (int k = 0; k < 4; k++)
            *reinterpret_cast<uint4*>(tile + t * 128 + ((k ^ (t & 7)) << 4)) = make_uint4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
    }
    if (!live) return;
    const uint32_t b = b0 + (uint32_t)t;
    uint8_t* dst = out + ((size_t)f * nb * 64 + ((size_t)(b / wb) * 8 * W + (size_t)(b % wb) * 8)) * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t off = t * 128 + ((k ^ (t & 7)) << 4);
        const uint4 y = *reinterpret_cast<const uint4*>(smem + off), cb = *reinterpret_cast<const uint4*>(smem + TILE_BYTES + off),
                    cr = *reinterpret_cast<const uint4*>(smem + 2 * TILE_BYTES + off);
        colour_row_store(y.x, y.y, cb.x, cb.y, cr.x, cr.y, dst + (size_t)(2 * k) * W * 4);
        colour_row_store(y.z, y.w, cb.z, cb.w, cr.z, cr.w, dst + (size_t)(2 * k + 1) * W * 4);
    }
}

// ---- position-mixed 64-bit checksum of each frame (bench: whole-batch bit-exactness) ----------------
// h(frame) = sum over 8-byte words w_i of mix(w_i ^ (i+1)*GOLDEN), mix = splitmix64 finaliser; the sum is
// order-independent, so it parallelises, yet every word is bound to its position.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
k_hash_frames(const uint64_t* __restrict__ frames, uint64_t words_per_frame, unsigned long long* __restrict__ hashes) {
    const uint64_t* fr = frames + (size_t)blockIdx.y * words_per_frame;
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_frame;
         i += (uint64_t)gridDim.x * blockDim.x)
        acc += mix64(fr[i] ^ ((i + 1) * 0x9E3779B97F4A7C15ull));
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&hashes[blockIdx.y], (unsigned long long)acc);
}

// ---- launchers ------------------------------------------------------------------------------------
cudaError_t launch_idct(const int16_t* d_coef, uint8_t* d_samples, size_t n_blocks, cudaStream_t s) {
    if (n_blocks == 0) return cudaSuccess;
    size_t grid = (n_blocks + IDCT_TPB - 1) / IDCT_TPB;
    k_idct<<<(unsigned)grid, IDCT_TPB, 0, s>>>(d_coef, d_samples, n_blocks);
    return cudaGetLastError();
}
cudaError_t launch_colour(const uint8_t* d_samples, void* d_out, uint32_t n_frames, uint32_t W, uint32_t H,
                          cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    uint32_t wb = W / 8, nb = wb * (H / 8);
    const uint32_t groups = (nb + IDCT_TPB - 1) / IDCT_TPB;
    if ((uint64_t)groups * n_frames > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    k_colour<<<groups * n_frames, IDCT_TPB, 0, s>>>(d_samples, (uint8_t*)d_out, nb, wb, W, groups);
    return cudaGetLastError();
}
cudaError_t launch_idct_colour(const int16_t* d_coef, void* d_out, uint32_t n_frames, uint32_t W, uint32_t H,
                               cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    static std::once_flag once[64];                      // per device, thread-safe
    static cudaError_t status[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::call_once(once[dev], [dev]() {
        status[dev] = cudaFuncSetAttribute(k_idct_colour, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_BYTES);
    });
    if (status[dev] != cudaSuccess) return status[dev];
    uint32_t wb = W / 8, nb = wb * (H / 8);
    const uint32_t groups = (nb + IDCT_TPB - 1) / IDCT_TPB;
    if ((uint64_t)groups * n_frames > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    k_idct_colour<<<groups * n_frames, IDCT_TPB, 3 * TILE_BYTES, s>>>(d_coef, (uint8_t*)d_out, nb, wb, W, groups);
    return cudaGetLastError();
}
cudaError_t launch_hash_frames(const void* d_frames, uint64_t frame_bytes, uint32_t n, unsigned long long* d_hashes,
                               cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_hashes, 0, (size_t)n * 8, s);
    if (e != cudaSuccess) return e;
    uint64_t words = frame_bytes / 8;
    uint64_t gx64 = (words + 255) / 256;
    unsigned gx = (unsigned)(gx64 < 64 ? gx64 : 64);
    for (uint32_t f0 = 0; f0 < n; f0 += 65535u) {            // gridDim.y <= 65535
        const uint32_t m = n - f0 < 65535u ? n - f0 : 65535u;
        k_hash_frames<<<dim3(gx ? gx : 1, m), 256, 0, s>>>((const uint64_t*)d_frames + (size_t)f0 * words, words, d_hashes + f0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace mj
