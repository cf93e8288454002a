/*
 * mbe_batch_shim.c - the binding a maintainer of arancormonk/mbelib-neo (or of an SDR application that
 * uses it) would add to route MANY voice streams through libmbe_b200.so instead of calling
 * mbe_process<Codec>Frame once per stream per 20 ms.
 *
 * The reference API (include/mbelib-neo/mbelib.h) is one call per frame with caller-owned state:
 *
 *     int mbe_processAmbe3600x2450Frame(short* aout_buf, mbe_process_result* result,
 *                                       const char ambe_fr[4][24], char ambe_d[49],
 *                                       mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);
 *
 * The batched replacement keeps the same per-frame semantics; state lives on the GPU between calls and
 * can be moved in and out in the reference's own mbe_parms layout (mbe_b200_import_state/export_state).
 *
 * Build (host C only, no CUDA headers needed):
 *     gcc -std=c99 -O2 -Iinclude examples/mbe_batch_shim.c -Lmbelib-neo_b200 -lmbe_b200 \
 *         -Wl,-rpath,'$ORIGIN/../mbelib-neo_b200' -o examples/mbe_batch_shim
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mbe_b200.h"

/* A pool of voice channels that decode in lockstep, one 20 ms frame (or several) per call. */
typedef struct {
    mbe_b200_ctx* ctx;
    int codec;
    int n_streams;
} mbe_batch;

/* == per stream: mbe_setThreadRngSeed(seed[s]); mbe_initMbeParms(&cur, &prev, &prev_enh); */
int mbe_batch_open(mbe_batch* b, int codec, int n_streams, const uint32_t* seeds, int gpu) {
    memset(b, 0, sizeof(*b));
    b->codec = codec;
    b->n_streams = n_streams;
    int rc = mbe_b200_create(&b->ctx, gpu, n_streams);
    if (rc != 0) {
        fprintf(stderr, "mbe_b200_create: %s\n", mbe_b200_last_error(NULL));
        return rc; /* MBE_B200_E_NOGPU: there is deliberately no CPU fallback */
    }
    return mbe_b200_init_streams(b->ctx, 0, n_streams, seeds);
}

/* == for every stream s, for f in 0..n_frames-1:
 *        ret[s][f] = mbe_process<Codec>Frame(pcm[s][f], &res[s][f], frames[s][f], bits[s][f], &cur[s], &prev[s], &enh[s]);
 * frames: one byte per bit, [stream][frame][rows*cols] exactly like `char fr[R][C]`.
 * results[s][f].status is the reference's return value (>= 0 total errors, -1 / -2 status codes). */
int mbe_batch_process(mbe_batch* b, int n_frames, const uint8_t* frames, int16_t* pcm, mbe_b200_result* results,
                      uint8_t* bits) {
    return mbe_b200_process_frames(b->ctx, b->codec, /*soft=*/0, 0, b->n_streams, n_frames, frames, pcm, NULL, results,
                                   bits);
}

/* hand one stream back to the CPU library: buf receives {cur_mp, prev_mp, prev_mp_enhanced} as 3 x mbe_parms */
int mbe_batch_export_stream(mbe_batch* b, int stream, void* three_mbe_parms) {
    return mbe_b200_export_state(b->ctx, stream, 1, three_mbe_parms);
}

void mbe_batch_close(mbe_batch* b) {
    mbe_b200_destroy(b->ctx);
    b->ctx = NULL;
}

int main(void) {
    enum { STREAMS = 256, FRAMES = 10 };
    int fbits = 0, pbits = 0;
    mbe_b200_geometry(MBE_B200_AMBE3600X2450, &fbits, &pbits);
    mbe_batch b;
    if (mbe_batch_open(&b, MBE_B200_AMBE3600X2450, STREAMS, NULL, 0) != 0) {
        return 2;
    }
    uint8_t* frames = (uint8_t*)malloc((size_t)STREAMS * FRAMES * fbits);
    int16_t* pcm = (int16_t*)malloc((size_t)STREAMS * FRAMES * MBE_B200_SAMPLES_PER_FRAME * sizeof(int16_t));
    mbe_b200_result* res = (mbe_b200_result*)malloc((size_t)STREAMS * FRAMES * sizeof(mbe_b200_result));
    unsigned x = 12345u;
    for (size_t i = 0; i < (size_t)STREAMS * FRAMES * fbits; ++i) {
        x = x * 1664525u + 1013904223u;
        frames[i] = (uint8_t)((x >> 24) & 1u);
    }
    int rc = mbe_batch_process(&b, FRAMES, frames, pcm, res, NULL);
    if (rc != 0) {
        fprintf(stderr, "process: %s\n", mbe_b200_last_error(b.ctx));
        return 1;
    }
    long long errs = 0;
    int peak = 0;
    for (int i = 0; i < STREAMS * FRAMES; ++i) {
        errs += res[i].total_errors;
    }
    for (int i = 0; i < STREAMS * FRAMES * MBE_B200_SAMPLES_PER_FRAME; ++i) {
        int a = pcm[i] < 0 ? -pcm[i] : pcm[i];
        peak = a > peak ? a : peak;
    }
    printf("%s: decoded %d frames, %lld corrected bit errors, peak |pcm| %d, %lld kernel launches\n", mbe_b200_version(),
           STREAMS * FRAMES, errs, peak, mbe_b200_launch_count(b.ctx));
    free(frames);
    free(pcm);
    free(res);
    mbe_batch_close(&b);
    return 0;
}
