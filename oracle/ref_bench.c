/*
 * TEST INFRASTRUCTURE - not product code.
 *
 * pthread driver around the UNMODIFIED reference library (linked from oracle/_ref/libmberef*.so).
 * It calls only the reference's public API (include/mbelib-neo/mbelib.h) the way an SDR application
 * would: one mbe_parms triplet per voice stream, frames fed in order.  Used for two things:
 *   - differential parity: same seeded inputs through the reference and through the oracle / GPU;
 *   - the CPU throughput arm of bench.py (`--impl reference`, `cpu_baseline`), one stream per task,
 *     tasks pulled by N worker threads.
 *
 * Per-stream RNG convention: mbe_setThreadRngSeed(seed_s) then
 * mbe_initMbeParms() before the first frame of every stream.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mbelib-neo/mbelib.h"

enum { CODEC_IMBE7200 = 0, CODEC_IMBE7100 = 1, CODEC_AMBE2400 = 2, CODEC_AMBE2450 = 3 };

static int
codec_frame_bits(int codec) {
    switch (codec) {
        case CODEC_IMBE7200: return 8 * 23;
        case CODEC_IMBE7100: return 7 * 24;
        default: return 4 * 24;
    }
}

static int
codec_param_bits(int codec) {
    return (codec == CODEC_IMBE7200 || codec == CODEC_IMBE7100) ? 88 : 49;
}

typedef struct {
    int codec, soft, n_streams, n_frames;
    const uint8_t* frames;
    const uint32_t* seeds;
    int16_t* pcm;
    float* pcmf;
    int32_t* results; /* [stream][frame][6]: ret, c0, protected, c4, total, flags */
    uint8_t* bits;    /* [stream][frame][param_bits] */
    mbe_parms* state; /* [stream][3] cur, prev, prev_enh (final) */
    volatile int next;
    pthread_mutex_t lock;
} job_t;

static void
run_stream(job_t* j, int s) {
    mbe_parms cur, prev, enh;
    char d[88];
    short pcm[160];
    float pcmf[160];
    const int fb = codec_frame_bits(j->codec);
    const int pb = codec_param_bits(j->codec);
    const size_t fstride = (size_t)fb * (j->soft ? 2u : 1u);

    mbe_setThreadRngSeed(j->seeds ? j->seeds[s] : 0u);
    mbe_initMbeParms(&cur, &prev, &enh);

    for (int f = 0; f < j->n_frames; ++f) {
        const uint8_t* fr = j->frames + ((size_t)s * j->n_frames + f) * fstride;
        mbe_process_result res;
        int ret;
        memset(&res, 0, sizeof(res));
        memset(d, 0, sizeof(d));
        if (!j->soft) {
            switch (j->codec) {
                case CODEC_IMBE7200:
                    ret = mbe_processImbe7200x4400Framef(pcmf, &res, (const char(*)[23])fr, d, &cur, &prev, &enh);
                    break;
                case CODEC_IMBE7100:
                    ret = mbe_processImbe7100x4400Framef(pcmf, &res, (const char(*)[24])fr, d, &cur, &prev, &enh);
                    break;
                case CODEC_AMBE2400:
                    ret = mbe_processAmbe3600x2400Framef(pcmf, &res, (const char(*)[24])fr, d, &cur, &prev, &enh);
                    break;
                default:
                    ret = mbe_processAmbe3600x2450Framef(pcmf, &res, (const char(*)[24])fr, d, &cur, &prev, &enh);
                    break;
            }
        } else {
            switch (j->codec) {
                case CODEC_IMBE7200:
                    ret = mbe_processImbe7200x4400SoftFramef(pcmf, &res, (const mbe_soft_bit(*)[23])fr, d, &cur, &prev,
                                                             &enh);
                    break;
                case CODEC_IMBE7100:
                    ret = mbe_processImbe7100x4400SoftFramef(pcmf, &res, (const mbe_soft_bit(*)[24])fr, d, &cur, &prev,
                                                             &enh);
                    break;
                case CODEC_AMBE2400:
                    ret = mbe_processAmbe3600x2400SoftFramef(pcmf, &res, (const mbe_soft_bit(*)[24])fr, d, &cur, &prev,
                                                             &enh);
                    break;
                default:
                    ret = mbe_processAmbe3600x2450SoftFramef(pcmf, &res, (const mbe_soft_bit(*)[24])fr, d, &cur, &prev,
                                                             &enh);
                    break;
            }
        }
        if (ret < 0) {
            memset(pcmf, 0, sizeof(pcmf));
            memset(pcm, 0, sizeof(pcm));
        } else {
            mbe_floattoshort(pcmf, pcm);
        }
        const size_t idx = (size_t)s * j->n_frames + f;
        if (j->pcm) {
            memcpy(j->pcm + idx * 160, pcm, sizeof(pcm));
        }
        if (j->pcmf) {
            memcpy(j->pcmf + idx * 160, pcmf, sizeof(pcmf));
        }
        if (j->results) {
            int32_t* r = j->results + idx * 6;
            r[0] = ret;
            r[1] = res.c0_errors;
            r[2] = res.protected_errors;
            r[3] = res.c4_errors;
            r[4] = res.total_errors;
            r[5] = (int32_t)res.flags;
        }
        if (j->bits) {
            memcpy(j->bits + idx * pb, d, (size_t)pb);
        }
    }
    if (j->state) {
        j->state[(size_t)s * 3 + 0] = cur;
        j->state[(size_t)s * 3 + 1] = prev;
        j->state[(size_t)s * 3 + 2] = enh;
    }
}

static void*
worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&j->lock);
        int s = j->next;
        j->next = s + 16;
        pthread_mutex_unlock(&j->lock);
        if (s >= j->n_streams) {
            break;
        }
        int e = s + 16;
        if (e > j->n_streams) {
            e = j->n_streams;
        }
        for (; s < e; ++s) {
            run_stream(j, s);
        }
    }
    return NULL;
}

/* Returns wall-clock seconds spent decoding (threads started -> joined). */
double
ref_bench_run(int codec, int soft, int n_streams, int n_frames, const uint8_t* frames, const uint32_t* seeds,
              int16_t* pcm, float* pcmf, int32_t* results, uint8_t* bits, void* state, int n_threads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.codec = codec;
    j.soft = soft;
    j.n_streams = n_streams;
    j.n_frames = n_frames;
    j.frames = frames;
    j.seeds = seeds;
    j.pcm = pcm;
    j.pcmf = pcmf;
    j.results = results;
    j.bits = bits;
    j.state = (mbe_parms*)state;
    j.next = 0;
    pthread_mutex_init(&j.lock, NULL);
    if (n_threads < 1) {
        n_threads = 1;
    }
    if (n_threads > 1024) {
        n_threads = 1024;
    }

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (n_threads == 1) {
        worker(&j);
    } else {
        /* worker i is pinned to the i-th CPU of the caller's affinity mask, 1:1 while there are CPUs left
           ("nproc worker threads pinned 1:1 to cores"); best effort, unpinned on failure */
        cpu_set_t allowed;
        int cpus[1024], n_cpus = 0;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            for (int c = 0; c < CPU_SETSIZE && n_cpus < 1024; ++c) {
                if (CPU_ISSET(c, &allowed)) {
                    cpus[n_cpus++] = c;
                }
            }
        }
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
        for (int i = 0; i < n_threads; ++i) {
            pthread_attr_t at;
            pthread_attr_init(&at);
            if (n_cpus > 0 && n_threads <= n_cpus) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[i], &one);
                pthread_attr_setaffinity_np(&at, sizeof(one), &one);
            }
            if (pthread_create(&th[i], &at, worker, &j) != 0) {
                pthread_create(&th[i], NULL, worker, &j);
            }
            pthread_attr_destroy(&at);
        }
        for (int i = 0; i < n_threads; ++i) {
            pthread_join(th[i], NULL);
        }
        free(th);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_mutex_destroy(&j.lock);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int
ref_bench_sizeof_parms(void) {
    return (int)sizeof(mbe_parms);
}
