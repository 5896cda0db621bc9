/* seek_and_shard.c -- a C99 caller of the round-2 entry points of include/mjpeg423_b200.h: the I-frame index with the
 * player's jumps (C0/playback.c:157-227) and the frame-range decode sharded over the GPUs of a box.
 *
 *   gcc -std=c99 -Iinclude examples/seek_and_shard.c -Lmjpeg423-video-decoder-software_b200 -lmjpeg423_b200 \
 *       -Wl,-rpath,$PWD/mjpeg423-video-decoder-software_b200 -o seek_and_shard
 *   ./seek_and_shard part0.mpg [part1.mpg ...]        # one logical stream; decoded on every GPU of the box
 */
#include <stdio.h>
#include <stdlib.h>

#include "mjpeg423_b200.h"

static uint8_t* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return NULL; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* p = (uint8_t*)malloc((size_t)n + 64);
    if (!p || fread(p, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "cannot read %s\n", path); fclose(f); free(p); return NULL; }
    fclose(f);
    *len = (size_t)n;
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s part0.mpg [part1.mpg ...]\n", argv[0]);
        return 2;
    }
    uint32_t n_shards = (uint32_t)(argc - 1);
    mjpeg423_b200_shard* shards = (mjpeg423_b200_shard*)calloc(n_shards, sizeof *shards);
    uint64_t total = 0, frame_bytes = 0;
    for (uint32_t s = 0; s < n_shards; s++) {
        size_t len = 0;
        uint8_t* p = slurp(argv[s + 1], &len);
        if (!p) return 1;
        shards[s].mpg = p;
        shards[s].len = len;
        mjpeg423_b200_info info;
        if (mjpeg423_b200_probe(p, len, &info) != MJPEG423_OK) { fprintf(stderr, "Error: %s\n", mjpeg423_b200_last_error()); return 1; }
        total += info.num_frames;
        frame_bytes = info.frame_bytes;
        /* the trailer of every file, checked against its frame headers */
        uint32_t n_i = 0;
        int ok = 0;
        if (mjpeg423_b200_index(p, len, NULL, 0, &n_i, &ok) != MJPEG423_OK) return 1;
        iframe_trailer_t* idx = (iframe_trailer_t*)malloc((n_i ? n_i : 1) * sizeof *idx);
        mjpeg423_b200_index(p, len, idx, n_i, &n_i, &ok);
        int ff = mjpeg423_b200_fast_forward(idx, n_i, info.num_frames, 0);
        printf("%s: %u frames, %u I frames, trailer %s; fast-forward from frame 0 lands on frame %d\n", argv[s + 1],
               info.num_frames, n_i, ok ? "ok" : "missing or wrong (rebuilt from the headers)", ff < 0 ? -1 : (int)idx[ff].frame_index);
        free(idx);
    }
    int n_dev = mjpeg423_b200_device_count();
    if (n_dev <= 0) { fprintf(stderr, "Error: no CUDA device (this library has no CPU fallback)\n"); return 1; }
    void* out = mjpeg423_b200_host_alloc((size_t)(total * frame_bytes));      /* pinned: full PCIe speed */
    uint64_t cuts[65];
    if (!out || mjpeg423_b200_decode_frames_multi(NULL, n_dev > 64 ? 64 : n_dev, shards, n_shards, 0, total, out, cuts) != MJPEG423_OK) {
        fprintf(stderr, "Error: %s\n", mjpeg423_b200_last_error());
        return 1;
    }
    printf("decoded %llu frames on %d GPU(s); first piece ends at frame %llu\n", (unsigned long long)total, n_dev, (unsigned long long)cuts[1]);
    mjpeg423_b200_host_free(out);
    return 0;
}
