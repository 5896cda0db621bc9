/* decode_file.c -- a C99 caller of libmjpeg423_b200: the reference's LIB/sample_main.c, with the decoder behind the
 * C-ABI of include/mjpeg423_b200.h.
 *
 *   gcc -std=c99 -Iinclude examples/decode_file.c -Lmjpeg423-video-decoder-software_b200 -lmjpeg423_b200 \
 *       -Wl,-rpath,$PWD/mjpeg423-video-decoder-software_b200 -o decode_file
 *   ./decode_file in.mpg out0000.bmp          # one 32-bpp BMP per frame, like mjpeg423_decode()
 *   ./decode_file in.mpg --play               # through the frame ring at 24 fps (C0/playback.c)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mjpeg423_b200.h"

static void shown(void* user, uint32_t frame_index, const rgb_pixel_t* frame) {
    (void)frame;
    *(uint32_t*)user = frame_index;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.mpg (out0000.bmp | --play)\n", argv[0]);
        return 2;
    }
    if (strcmp(argv[2], "--play") != 0) {          /* the reference entry point, unchanged */
        mjpeg423_decode(argv[1], argv[2]);
        return 0;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* mpg = (uint8_t*)malloc((size_t)len + 64);
    if (!mpg || fread(mpg, 1, (size_t)len, f) != (size_t)len) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);

    mjpeg423_b200_ctx* ctx = NULL;
    mjpeg423_b200_info info;
    if (mjpeg423_b200_create(&ctx, 0) != MJPEG423_OK || mjpeg423_b200_probe(mpg, (size_t)len, &info) != MJPEG423_OK) {
        fprintf(stderr, "Error: %s\n", mjpeg423_b200_last_error());
        return 1;
    }
    mjpeg423_b200_display* disp = mjpeg423_b200_display_init((int)info.w_size, (int)info.h_size, 4);   /* NUM_OUTPUT_BUFFERS */
    uint32_t last = 0, dropped = 0;
    long n = mjpeg423_b200_play(ctx, mpg, (size_t)len, 0, info.num_frames, disp, 41666 /* FRAME_RATE_US */, shown, &last, &dropped);
    if (n < 0) fprintf(stderr, "Error: %s\n", mjpeg423_b200_last_error());
    else printf("displayed %ld frames of %ux%u (last #%u), %u late ticks\n", n, info.w_size, info.h_size, last, dropped);
    mjpeg423_b200_display_free(disp);
    mjpeg423_b200_destroy(ctx);
    free(mpg);
    return n < 0;
}
