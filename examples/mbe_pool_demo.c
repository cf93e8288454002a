/*
 * mbe_pool_demo.c - one host process, every visible GPU: decode S voice streams x F frames of AMBE+2 through
 * mbe_b200_pool_process_frames_packed and print the throughput and a checksum of the PCM.
 *
 * The checksum does not depend on the number of GPUs (streams are independent and land bit-identical wherever they
 * are sharded), which is the property to look at when this is run with 1, 2, 4 and 8 devices.
 *
 * Build (host C only):
 *     gcc -std=c99 -O2 -Iinclude examples/mbe_pool_demo.c -Lmbelib-neo_b200 -lmbe_b200 \
 *         -Wl,-rpath,'$ORIGIN/../mbelib-neo_b200' -o examples/mbe_pool_demo
 * Run:  examples/mbe_pool_demo [streams=262144] [frames=50] [devices=0 (all)] [pageable]
 *
 * The frame and PCM arrays come from mbe_b200_host_alloc (page-locked, portable across the pool's devices): the caller does
 * not link the CUDA runtime.  A fourth argument "pageable" uses malloc instead, to show what that costs.
 */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mbe_b200.h"

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char** argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 262144;
    const int F = argc > 2 ? atoi(argv[2]) : 50;
    const int ndev = argc > 3 ? atoi(argv[3]) : 0;
    const int pageable = argc > 4 && strcmp(argv[4], "pageable") == 0;
    const int codec = MBE_B200_AMBE3600X2450;
    const size_t fbytes = (size_t)mbe_b200_packed_frame_bytes(codec);

    mbe_b200_pool* pool = NULL;
    int rc = mbe_b200_pool_create(&pool, ndev, NULL, S);
    if (rc != 0) {
        fprintf(stderr, "mbe_b200_pool_create: %s\n", mbe_b200_pool_last_error(NULL));
        return 1; /* no GPU, no decode: there is no CPU fallback */
    }
    const size_t frame_bytes = (size_t)S * F * fbytes, pcm_bytes = (size_t)S * F * MBE_B200_SAMPLES_PER_FRAME * sizeof(int16_t);
    uint8_t* frames = NULL;
    int16_t* pcm = NULL;
    if (pageable) {
        frames = (uint8_t*)malloc(frame_bytes);
        pcm = (int16_t*)malloc(pcm_bytes);
    } else if (mbe_b200_host_alloc((void**)&frames, frame_bytes) != 0 || mbe_b200_host_alloc((void**)&pcm, pcm_bytes) != 0) {
        fprintf(stderr, "mbe_b200_host_alloc: %s\n", mbe_b200_last_error(NULL));
        return 1;
    }
    uint32_t* seeds = (uint32_t*)malloc((size_t)S * sizeof(uint32_t));
    if (!frames || !pcm || !seeds) {
        fprintf(stderr, "out of host memory\n");
        return 1;
    }
    uint64_t st = 0x2450;
    for (size_t i = 0; i < (size_t)S * F * fbytes; i += 8) {
        const uint64_t r = splitmix64(&st);
        memcpy(frames + i, &r, ((size_t)S * F * fbytes - i) < 8 ? ((size_t)S * F * fbytes - i) : 8);
    }
    for (int s = 0; s < S; ++s) {
        seeds[s] = 0xC0FFEEu + (uint32_t)s;
    }
    double best = 1e30;
    for (int it = 0; it < 3; ++it) {
        if ((rc = mbe_b200_pool_init_streams(pool, 0, S, seeds)) != 0) {
            break;
        }
        const double t0 = now();
        rc = mbe_b200_pool_process_frames_packed(pool, codec, 0, S, F, frames, pcm, NULL, NULL, NULL);
        const double dt = now() - t0;
        if (rc != 0) {
            break;
        }
        if (dt < best) {
            best = dt;
        }
    }
    if (rc != 0) {
        fprintf(stderr, "pool call failed: %s\n", mbe_b200_pool_last_error(pool));
        return 1;
    }
    uint32_t h = 2166136261u;
    const uint8_t* b = (const uint8_t*)pcm;
    for (size_t i = 0; i < (size_t)S * F * MBE_B200_SAMPLES_PER_FRAME * sizeof(int16_t); ++i) {
        h = (h ^ b[i]) * 16777619u;
    }
    printf("{\"shards\": %d, \"streams\": %d, \"frames\": %d, \"host_memory\": \"%s\", \"seconds\": %.4f, \"frames_per_s\": %.4g, "
           "\"pcm_fnv1a32\": \"%08x\"}\n",
           mbe_b200_pool_shards(pool), S, F, pageable ? "pageable (malloc)" : "pinned (mbe_b200_host_alloc)", best,
           (double)S * F / best, h);
    mbe_b200_pool_destroy(pool);
    if (pageable) {
        free(frames);
        free(pcm);
    } else {
        mbe_b200_host_free(frames);
        mbe_b200_host_free(pcm);
    }
    free(seeds);
    return 0;
}
