/*
 * harness.c -- frame/container driver used by BOTH checkers (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled twice by oracle/Makefile:
 *   default        -> libmjpeg423_oracle.so, symbols orc_decode_frame / orc_decode_mpg, calling the
 *                     restatement in mjpeg423_oracle.c;
 *   -DHARNESS_REF  -> oracle/_ref/libmjpeg423_ref.so, symbols ref_decode_frame / ref_decode_mpg,
 *                     calling the reference's own lossless_decode / idct / ycbcr_to_rgb compiled
 *                     from /root/reference (this is the "reference" CPU baseline).
 *
 * It drives the three stage functions in the order of the reference loop body,
 * LIB/decoder/mjpeg423_decoder.c:106-124, on an in-memory .mpg (container layout: SURVEY.md A.1,
 * reader LIB/decoder/mjpeg423_decoder.c:33-38,94-107) -- no printf, no BMP.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef HARNESS_REF
#define FN(x) ref_##x
typedef int16_t dct_block_t[8][8];
typedef uint8_t color_block_t[8][8];
typedef struct { uint8_t blue, green, red, alpha; } rgb_pixel_t;
void lossless_decode(int num_blocks, void* bitstream, dct_block_t* DCACq, dct_block_t quant, int P);
void idct(dct_block_t DCAC, color_block_t block);
void ycbcr_to_rgb(int h, int w, uint32_t w_size, uint8_t (*Y)[8], uint8_t (*Cb)[8], uint8_t (*Cr)[8],
                  rgb_pixel_t* rgbblock);
void rgb_to_ycbcr(int h, int w, uint32_t w_size, rgb_pixel_t* rgbblock, uint8_t (*Y)[8], uint8_t (*Cb)[8], uint8_t (*Cr)[8]);
void fdct(uint8_t (*block)[8], int16_t (*DCAC)[8]);
void quantize_I(int16_t* prev, int16_t (*quant)[8], int16_t (*DCAC)[8], int16_t (*DCACq)[8], int16_t (*DCACq_next)[8]);
void quantize_P(int16_t (*quant)[8], int16_t (*DCACq_prev)[8], int16_t (*DCAC)[8], int16_t (*DCACq)[8]);
uint32_t lossless_encode(int num_blocks, dct_block_t* DCACq, void* bitstream);
void mjpeg423_encode(uint32_t num_frames, int first, double stride, uint32_t max_I_interval, uint32_t w_size,
                     uint32_t h_size, const char* filenamebase_in, const char* filename_out);
void encode_bmp(rgb_pixel_t* rgbbblock, uint32_t w_size, uint32_t h_size, const char* filename);
extern dct_block_t Yquant, Cquant;
#define DEFAULT_YQ ((const int16_t*)Yquant)
#define DEFAULT_CQ ((const int16_t*)Cquant)
static void st_entropy(int nb, const uint8_t* bs, int16_t* coef, const int16_t* q, int P) {
    lossless_decode(nb, (void*)bs, (dct_block_t*)coef, (int16_t(*)[8])q, P);
}
static void st_idct(int16_t* coef, uint8_t* samp) { idct((int16_t(*)[8])coef, (uint8_t(*)[8])samp); }
static void st_colour(int h, int w, uint32_t ws, uint8_t* y, uint8_t* cb, uint8_t* cr, uint8_t* rgb) {
    ycbcr_to_rgb(h, w, ws, (uint8_t(*)[8])y, (uint8_t(*)[8])cb, (uint8_t(*)[8])cr, (rgb_pixel_t*)rgb);
}
static void se_colour(int h, int w, uint32_t ws, const uint8_t* rgb, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    rgb_to_ycbcr(h, w, ws, (rgb_pixel_t*)rgb, (uint8_t(*)[8])y, (uint8_t(*)[8])cb, (uint8_t(*)[8])cr);
}
static void se_fdct(const uint8_t* blk, int16_t* coef) { fdct((uint8_t(*)[8])blk, (int16_t(*)[8])coef); }
static void se_quant_I(int16_t* prev, const int16_t* q, const int16_t* c, int16_t* lv, int16_t* next) {
    quantize_I(prev, (int16_t(*)[8])q, (int16_t(*)[8])c, (int16_t(*)[8])lv, (int16_t(*)[8])next);
}
static void se_quant_P(const int16_t* q, int16_t* prev, const int16_t* c, int16_t* lv) {
    quantize_P((int16_t(*)[8])q, (int16_t(*)[8])prev, (int16_t(*)[8])c, (int16_t(*)[8])lv);
}
static uint32_t se_entropy(int nb, const int16_t* lv, uint8_t* out, int fix_tail) {
    (void)fix_tail;                          /* the reference always zeroes the last partial byte */
    return lossless_encode(nb, (dct_block_t*)lv, out);
}
#define SE_CAN_FIX_TAIL 0
#else
#define FN(x) orc_##x
uint64_t orc_lossless_decode(int, const void*, int16_t*, const int16_t*, int);
void orc_idct(const int16_t*, uint8_t*);
void orc_ycbcr_to_rgb(int, int, uint32_t, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*);
extern const int16_t orc_Yquant[64], orc_Cquant[64];
#define DEFAULT_YQ orc_Yquant
#define DEFAULT_CQ orc_Cquant
static void st_entropy(int nb, const uint8_t* bs, int16_t* coef, const int16_t* q, int P) {
    orc_lossless_decode(nb, bs, coef, q, P);
}
static void st_idct(int16_t* coef, uint8_t* samp) { orc_idct(coef, samp); }
static void st_colour(int h, int w, uint32_t ws, uint8_t* y, uint8_t* cb, uint8_t* cr, uint8_t* rgb) {
    orc_ycbcr_to_rgb(h, w, ws, y, cb, cr, rgb);
}
void orc_rgb_to_ycbcr(int, int, uint32_t, const uint8_t*, uint8_t*, uint8_t*, uint8_t*);
void orc_fdct(const uint8_t*, int16_t*);
void orc_quantize_I(int16_t*, const int16_t*, const int16_t*, int16_t*, int16_t*);
void orc_quantize_P(const int16_t*, int16_t*, const int16_t*, int16_t*);
uint32_t orc_lossless_encode(int, const int16_t*, uint8_t*, int);
#define se_colour orc_rgb_to_ycbcr
#define se_fdct orc_fdct
#define se_quant_I orc_quantize_I
#define se_quant_P orc_quantize_P
#define se_entropy orc_lossless_encode
#define SE_CAN_FIX_TAIL 1
#endif

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* coef = 3*nb*64 int16 (Y|Cb|Cr planes; the inter-frame state for P frames), samp = 3*nb*64 bytes
 * scratch, rgb = W*H*4 BGRA.  secs (optional) accumulates {entropy, idct, colour} seconds. */
void FN(decode_frame)(uint32_t W, uint32_t H, const uint8_t* payload, uint32_t Ysize, uint32_t Cbsize,
                      int P, const int16_t* yq, const int16_t* cq, int16_t* coef, uint8_t* samp,
                      uint8_t* rgb, double* secs) {
    int nb = (int)((W / 8) * (H / 8)), wb = (int)(W / 8);
    double t0 = now_s();
    st_entropy(nb, payload, coef, yq, P);
    st_entropy(nb, payload + Ysize, coef + (size_t)nb * 64, cq, P);
    st_entropy(nb, payload + Ysize + Cbsize, coef + (size_t)nb * 128, cq, P);
    double t1 = now_s();
    for (int b = 0; b < 3 * nb; b++) st_idct(coef + (size_t)b * 64, samp + (size_t)b * 64);
    double t2 = now_s();
    for (int b = 0; b < nb; b++)
        st_colour((b / wb) * 8, (b % wb) * 8, W, samp + (size_t)b * 64, samp + (size_t)(nb + b) * 64,
                  samp + (size_t)(2 * nb + b) * 64, rgb);
    double t3 = now_s();
    if (secs) { secs[0] += t1 - t0; secs[1] += t2 - t1; secs[2] += t3 - t2; }
}

static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
typedef struct {
    const uint8_t* mpg; size_t len; uint32_t W, H, first, n;
    const int16_t *yq, *cq; uint8_t* out; int rc; double secs[3];
} job_t;

static void* job_run(void* arg) {
    job_t* j = (job_t*)arg;
    uint32_t nframes = rd32(j->mpg);
    size_t nb = (size_t)(j->W / 8) * (j->H / 8), fbytes = (size_t)j->W * j->H * 4;
    int16_t* coef = (int16_t*)calloc(3 * nb * 64, 2);
    uint8_t* samp = (uint8_t*)malloc(3 * nb * 64);
    uint8_t* scratch = (uint8_t*)malloc(fbytes);
    j->rc = -1;
    if (!coef || !samp || !scratch) goto done;
    {
        /* P frames accumulate on the previous frame's coefficients: start at the last I frame
         * at or before `first` and decode the run-in frames into scratch. */
        size_t off = 20, start_off = 20;
        uint32_t start = 0;
        for (uint32_t f = 0; f <= j->first && f < nframes; f++) {
            if (off + 16 > j->len) goto done;
            if (rd32(j->mpg + off + 4) == 0) { start = f; start_off = off; }
            if (f < j->first) off += rd32(j->mpg + off);
        }
        off = start_off;
        for (uint32_t f = start; f < j->first + j->n; f++) {
            if (f >= nframes || off + 16 > j->len) goto done;
            uint32_t fsz = rd32(j->mpg + off), type = rd32(j->mpg + off + 4);
            uint32_t ys = rd32(j->mpg + off + 8), cbs = rd32(j->mpg + off + 12);
            if (fsz < 16 || off + fsz > j->len) goto done;
            uint8_t* dst = f >= j->first ? j->out + (size_t)(f - j->first) * fbytes : scratch;
            FN(decode_frame)(j->W, j->H, j->mpg + off + 16, ys, cbs, (int)type, j->yq, j->cq, coef, samp,
                             dst, j->secs);
            off += fsz;
        }
    }
    j->rc = 0;
done:
    free(coef); free(samp); free(scratch);
    return NULL;
}

/* Decode frames [first, first+n) of an in-memory .mpg into out (n * W*H*4 BGRA bytes) with nthreads
 * host threads, each taking a contiguous frame range (the stage functions are re-entrant,
 * SURVEY.md 8b "Threading").  stage_secs (optional, 3 doubles) receives the summed per-thread
 * {entropy, idct, colour} seconds.  Returns 0 on success.  The buffer must stay readable 4 bytes
 * past the last payload byte (reference look-ahead); a well-formed file is. */
int FN(decode_mpg)(const uint8_t* mpg, size_t len, uint32_t first, uint32_t n, const int16_t* yq,
                   const int16_t* cq, uint8_t* out, int nthreads, double* stage_secs) {
    if (len < 20) return -1;
    uint32_t W = rd32(mpg + 4), H = rd32(mpg + 8);
    if (!W || !H || (W & 7) || (H & 7)) return -2;
    if (n == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n) nthreads = (int)n;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    size_t fbytes = (size_t)W * H * 4;
    for (int t = 0; t < nthreads; t++) {
        uint32_t a = (uint32_t)((uint64_t)n * (uint64_t)t / (uint64_t)nthreads);
        uint32_t b = (uint32_t)((uint64_t)n * (uint64_t)(t + 1) / (uint64_t)nthreads);
        job_t jb = {mpg, len, W, H, first + a, b - a, yq ? yq : DEFAULT_YQ, cq ? cq : DEFAULT_CQ,
                    out + (size_t)a * fbytes, 0, {0, 0, 0}};
        jobs[t] = jb;
        if (nthreads == 1) job_run(&jobs[t]);
        else pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        if (stage_secs) for (int k = 0; k < 3; k++) stage_secs[k] += jobs[t].secs[k];
    }
    free(jobs); free(th);
    return rc;
}


/* ==== encoder driver ==========================================================
 * The frame loop of mjpeg423_encode(), LIB/encoder/mjpeg423_encoder.c:97-225, on in-memory BGRA frames
 * (the BMP reader is HAL-side I/O): colour -> FDCT -> quantise I (and P from the second frame on) ->
 * entropy code both -> keep the I frame when it is the first frame, not larger than the P frame, or
 * max_I_interval frames after the last I frame -> frame record -> trailer.  The 512 trailing pad bytes
 * (uninitialised stack in the reference, :219-220) are written as zeros.
 * Returns 0 and *len, -1 on bad arguments, -2 when `cap` is too small, -3 when fix_tail is requested from
 * the compiled reference (its lossless_encode cannot do it). */
static void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

int FN(encode_mpg)(const uint8_t* bgra, uint32_t n, uint32_t W, uint32_t H, uint32_t max_I_interval,
                   const int16_t* yq, const int16_t* cq, int fix_tail, uint8_t* out, size_t cap, size_t* len) {
    if (!W || !H || (W & 7) || (H & 7) || !out || !len || (n && !bgra)) return -1;
    if (fix_tail && !SE_CAN_FIX_TAIL) return -3;
    if (!yq) yq = DEFAULT_YQ;
    if (!cq) cq = DEFAULT_CQ;
    const size_t nb = (size_t)(W / 8) * (H / 8), wb = W / 8, plane = nb * 64;
    uint8_t* samp = (uint8_t*)malloc(3 * plane);
    int16_t* coef = (int16_t*)malloc(3 * plane * 2);
    int16_t* lvI = (int16_t*)malloc(3 * plane * 2);
    int16_t* lvP = (int16_t*)malloc(3 * plane * 2);
    int16_t* next = (int16_t*)malloc(3 * plane * 2);
    int16_t* prev = (int16_t*)calloc(3 * plane, 2);
    uint8_t* bsI = (uint8_t*)malloc(3 * plane * 2 + 64);
    uint8_t* bsP = (uint8_t*)malloc(3 * plane * 2 + 64);
    uint32_t* trailer = (uint32_t*)malloc((size_t)(n ? n : 1) * 8);
    int rc = -2;
    size_t pos = 20;
    uint32_t n_i = 0, last_i = 0;
    if (cap < 20) goto done;
    for (uint32_t f = 0; f < n; f++) {
        const uint8_t* frame = bgra + (size_t)f * W * H * 4;
        for (size_t b = 0; b < nb; b++)
            se_colour((int)(b / wb) * 8, (int)(b % wb) * 8, W, frame, samp + b * 64, samp + plane + b * 64,
                      samp + 2 * plane + b * 64);
        for (size_t b = 0; b < 3 * nb; b++) se_fdct(samp + b * 64, coef + b * 64);
        uint32_t szI[3], szP[3] = {0, 0, 0};
        for (int p = 0; p < 3; p++) {
            int16_t dcprev = 0;
            const int16_t* q = p ? cq : yq;
            for (size_t b = 0; b < nb; b++)
                se_quant_I(&dcprev, q, coef + p * plane + b * 64, lvI + p * plane + b * 64, next + p * plane + b * 64);
        }
        {
            uint8_t* o = bsI;
            for (int p = 0; p < 3; p++) { szI[p] = se_entropy((int)nb, lvI + p * plane, o, fix_tail); o += szI[p]; }
        }
        if (f > 0) {
            for (int p = 0; p < 3; p++) {
                const int16_t* q = p ? cq : yq;
                for (size_t b = 0; b < nb; b++)
                    se_quant_P(q, prev + p * plane + b * 64, coef + p * plane + b * 64, lvP + p * plane + b * 64);
            }
            uint8_t* o = bsP;
            for (int p = 0; p < 3; p++) { szP[p] = se_entropy((int)nb, lvP + p * plane, o, fix_tail); o += szP[p]; }
        }
        const uint32_t totI = szI[0] + szI[1] + szI[2], totP = szP[0] + szP[1] + szP[2];
        const int isI = f == 0 || totI <= totP || f - last_i >= max_I_interval;
        const uint32_t* sz = isI ? szI : szP;
        const uint8_t* bs = isI ? bsI : bsP;
        if (isI) { last_i = f; int16_t* t = prev; prev = next; next = t; }
        const uint32_t body = sz[0] + sz[1] + sz[2];
        uint32_t fsz = body + 16;
        fsz = (fsz + 3u) & ~3u;
        if (pos + fsz > cap) goto done;
        wr32(out + pos, fsz); wr32(out + pos + 4, isI ? 0u : 1u); wr32(out + pos + 8, sz[0]); wr32(out + pos + 12, sz[1]);
        memcpy(out + pos + 16, bs, body);
        memset(out + pos + 16 + body, 0, fsz - 16 - body);
        if (isI) { trailer[2 * n_i] = f; trailer[2 * n_i + 1] = (uint32_t)pos; n_i++; }
        pos += fsz;
    }
    if (pos + (size_t)n_i * 8 + 512 > cap) goto done;
    wr32(out, n); wr32(out + 4, W); wr32(out + 8, H); wr32(out + 12, n_i); wr32(out + 16, (uint32_t)(pos - 20));
    for (uint32_t k = 0; k < n_i; k++) { wr32(out + pos, trailer[2 * k]); wr32(out + pos + 4, trailer[2 * k + 1]); pos += 8; }
    memset(out + pos, 0, 512);
    pos += 512;
    *len = pos;
    rc = 0;
done:
    free(samp); free(coef); free(lvI); free(lvP); free(next); free(prev); free(bsI); free(bsP); free(trailer);
    return rc;
}

#ifdef HARNESS_REF
#include <stdio.h>
/* The reference's own mjpeg423_encode() end to end: frames -> 32-bpp BMP files (the reference's encode_bmp,
 * LIB/libbmp/encode_bmp.c) named <dir>/f0000.bmp ... -> mjpeg423_encode -> <dir>/out.mpg, read back into out.
 * Pins ref_encode_mpg's driver loop against the real thing (tests/test_oracle.py).  n <= 10000. */
int ref_encode_mpg_files(const uint8_t* bgra, uint32_t n, uint32_t W, uint32_t H, uint32_t max_I_interval,
                         const char* dir, uint8_t* out, size_t cap, size_t* len) {
    char name[1024], outname[1024];
    if (n > 10000 || strlen(dir) > 900) return -1;
    for (uint32_t f = 0; f < n; f++) {
        snprintf(name, sizeof name, "%s/f%04u.bmp", dir, f);
        encode_bmp((rgb_pixel_t*)(bgra + (size_t)f * W * H * 4), W, H, name);
    }
    snprintf(name, sizeof name, "%s/f0000.bmp", dir);
    snprintf(outname, sizeof outname, "%s/out.mpg", dir);
    mjpeg423_encode(n, 0, 1.0, max_I_interval, W, H, name, outname);
    FILE* fp = fopen(outname, "rb");
    if (!fp) return -2;
    *len = fread(out, 1, cap, fp);
    int more = fgetc(fp) != EOF;
    fclose(fp);
    return more ? -2 : 0;
}
#endif
