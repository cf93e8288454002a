/*
 * mbe_single_shim.c - single-stream drop-in shim (include/mbe_b200_compat.h): the reference's per-frame
 * entry points on top of a batch of ONE through libmbe_b200.so.  Plain C host code; all decode and synthesis work
 * happens in the CUDA kernels behind the C-ABI, this file only moves the caller-owned state in and out and mirrors
 * the reference's argument checks (NULL pointers -> MBE_STATUS_INVALID_ARGUMENT before anything is touched).
 *
 * State mapping per call: {cur_mp, prev_mp, prev_mp_enhanced} -> stream slot 0 (mbe_b200_import_state), the calling
 * thread's RNG words (the reference's thread-locals, src/core/mbe_adaptive.c:29-30, src/core/mbe_unvoiced_fft.c:29-30)
 * -> mbe_b200_import_rng; one frame; both exported back.
 *
 * Re-entrancy (mbelib.h:28-30: "re-entrant per stream, the caller owns every buffer"): every calling thread gets its OWN
 * context (one stream slot, own CUDA streams), created on first use and destroyed when the thread exits - no lock is held
 * around a call, threads decode side by side.
 * Failure (no CUDA device, a CUDA error): never abort().  The reference's int functions report MBE_STATUS_INVALID_ARGUMENT,
 * its void helpers leave silence / their arguments untouched (mbelib.c:1048-1055); the shim does the same and says why on
 * stderr once per thread.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mbe_b200.h"
#include "../../include/mbe_b200_compat.h"

#define NSAMP 160

static __thread mbe_b200_ctx* t_ctx;
static __thread int t_ctx_tried;   /* creation failed once: do not retry (and do not repeat the message) on every frame */
static __thread int t_failed;      /* a C-ABI call of the current shim call failed */
static __thread int t_warned;
static __thread uint32_t t_rng[4];
static __thread int t_rng_ready;
static pthread_key_t g_key;
static pthread_once_t g_key_once = PTHREAD_ONCE_INIT;

static void ctx_destructor(void* p) { mbe_b200_destroy((mbe_b200_ctx*)p); }
static void key_make(void) { pthread_key_create(&g_key, ctx_destructor); }

/* the calling thread's context, created on first use; NULL (and t_failed set) when there is no usable GPU: the product has
 * no CPU path, so the call then fails the way the reference's argument checks fail */
static mbe_b200_ctx* shim_ctx(void) {
    t_failed = 0;
    if (!t_ctx && !t_ctx_tried) {
        t_ctx_tried = 1;
        const char* dev = getenv("MBE_B200_DEVICE");
        const int rc = mbe_b200_create(&t_ctx, dev ? atoi(dev) : 0, 1);
        if (rc != 0) {
            fprintf(stderr, "libmbe-neo-b200shim: mbe_b200_create failed (%d): %s - there is no CPU fallback; calls on this "
                            "thread return MBE_STATUS_INVALID_ARGUMENT / silence\n", rc, mbe_b200_last_error(NULL));
            t_ctx = NULL;
        } else {
            pthread_once(&g_key_once, key_make);
            pthread_setspecific(g_key, t_ctx);
        }
    }
    if (!t_ctx) {
        t_failed = 1;
    }
    return t_ctx;
}

static void note_failure(mbe_b200_ctx* c, const char* what, int rc) {
    t_failed = 1;
    if (!t_warned) {
        t_warned = 1;
        fprintf(stderr, "libmbe-neo-b200shim: %s failed (%d): %s\n", what, rc, mbe_b200_last_error(c));
    }
}
/* one C-ABI call; skipped once an earlier call of the same shim call has failed */
#define CK(call)                                                                                                      \
    do {                                                                                                              \
        if (!t_failed) {                                                                                              \
            const int rc_ = (call);                                                                                   \
            if (rc_ != 0) {                                                                                           \
                note_failure(c, #call, rc_);                                                                          \
            }                                                                                                         \
        }                                                                                                             \
    } while (0)

/* this thread's RNG words; a thread that never called mbe_setThreadRngSeed starts from the fresh-thread defaults */
static void rng_ready_locked(mbe_b200_ctx* c) {
    if (!t_rng_ready) {
        CK(mbe_b200_init_streams(c, 0, 1, NULL));
        CK(mbe_b200_export_rng(c, 0, 1, t_rng));
        t_rng_ready = !t_failed;
    }
}

/* ---- host-only helpers ----------------------------------------------------------------------------------- */
void mbe_initProcessResult(mbe_process_result* result) { /* mbelib.c:61-67 */
    if (result) {
        memset(result, 0, sizeof(*result));
    }
}

void mbe_formatProcessResult(char* str, size_t size, const mbe_process_result* result) { /* mbelib.c:69-105 */
    static const unsigned flag[4] = {MBE_PROCESS_FLAG_ERASURE, MBE_PROCESS_FLAG_TONE, MBE_PROCESS_FLAG_REPEAT,
                                     MBE_PROCESS_FLAG_MUTE};
    static const char mark[4] = {'E', 'T', 'R', 'M'};
    if (!str || size == 0u) {
        return;
    }
    size_t n = 0;
    const int errs = (result && result->total_errors > 0) ? result->total_errors : 0;
    while (n < (size_t)errs && n + 1u < size) {
        str[n++] = '=';
    }
    for (int i = 0; result && i < 4 && n + 1u < size; ++i) {
        if (result->flags & flag[i]) {
            str[n++] = mark[i];
        }
    }
    str[n] = '\0';
}

mbe_soft_bit mbe_softBitFromHard(int bit, uint8_t reliability) { /* mbelib.c:117-123 */
    mbe_soft_bit s = {(uint8_t)(bit ? 1 : 0), reliability};
    return s;
}

mbe_soft_bit mbe_softBitFromLlr(int16_t llr) { /* mbelib.c:125-132: sign = bit, |llr| saturated to 255 = reliability */
    const int mag = llr < 0 ? -(int)llr : (int)llr;
    mbe_soft_bit s = {(uint8_t)(llr > 0 ? 1 : 0), (uint8_t)(mag > 255 ? 255 : mag)};
    return s;
}

static int bits_status(const char* bits, size_t count) { /* src/internal/mbe_result.h:18-30 */
    if (!bits) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    for (size_t i = 0; i < count; ++i) {
        if (bits[i] != 0 && bits[i] != 1) {
            return MBE_STATUS_INVALID_BITS;
        }
    }
    return 0;
}

int mbe_softBitsFromHard(const char* bits, mbe_soft_bit* soft, size_t count, uint8_t reliability) { /* mbelib.c:134-147 */
    if (!soft) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    const int st = bits_status(bits, count);
    if (st < 0) {
        return st;
    }
    for (size_t i = 0; i < count; ++i) {
        soft[i] = mbe_softBitFromHard(bits[i], reliability);
    }
    return 0;
}

int mbe_softBitsFromLlr(const int16_t* llr, mbe_soft_bit* soft, size_t count) { /* mbelib.c:149-158 */
    if (!llr || !soft) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    for (size_t i = 0; i < count; ++i) {
        soft[i] = mbe_softBitFromLlr(llr[i]);
    }
    return 0;
}

const char* mbe_versionString(void) { return "2.0.0"; } /* the reference version this shim mirrors (mbelib.c:323-326) */

void mbe_moveMbeParms(const mbe_parms* source_mp, mbe_parms* destination_mp) { /* mbelib.c:338-344 */
    if (source_mp && destination_mp) {
        memmove(destination_mp, source_mp, sizeof(*destination_mp));
    }
}

void mbe_useLastMbeParms(mbe_parms* cur_mp, const mbe_parms* prev_mp) { /* mbelib.c:353-359 */
    if (cur_mp && prev_mp) {
        memmove(cur_mp, prev_mp, sizeof(*cur_mp));
    }
}

void mbe_synthesizeSilencef(float* aout_buf) { /* mbelib.c:862-868 */
    if (aout_buf) {
        memset(aout_buf, 0, NSAMP * sizeof(float));
    }
}

void mbe_synthesizeSilence(short* aout_buf) { /* mbelib.c:874-880 */
    if (aout_buf) {
        memset(aout_buf, 0, NSAMP * sizeof(short));
    }
}

/* ---- state ------------------------------------------------------------------------------------------------- */
void mbe_setThreadRngSeed(uint32_t seed) { /* mbelib.c:173-181: the device derives both generators from the seed */
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_init_streams(c, 0, 1, &seed));
    CK(mbe_b200_export_rng(c, 0, 1, t_rng));
    if (!t_failed) {
        t_rng_ready = 1;
    }
}

void mbe_initMbeParms(mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced) { /* mbelib.c:367-410 */
    if (!cur_mp || !prev_mp || !prev_mp_enhanced) {
        return;
    }
    mbe_parms t[3];
    mbe_b200_ctx* c = shim_ctx();
    rng_ready_locked(c);  /* (before the slot is reset, so a fresh thread's defaults are captured once) */
    CK(mbe_b200_init_streams(c, 0, 1, NULL));
    CK(mbe_b200_export_state(c, 0, 1, t));
    if (t_failed) {
        return;   /* the caller's structs stay as they were */
    }
    *cur_mp = t[0];
    *prev_mp = t[1];
    *prev_mp_enhanced = t[2];
}

/* ---- the hot path, one frame at a time --------------------------------------------------------------------- */
static void result_out(mbe_process_result* result, const mbe_b200_result* r) {
    if (result) {
        result->c0_errors = r->c0_errors;
        result->protected_errors = r->protected_errors;
        result->c4_errors = r->c4_errors;
        result->total_errors = r->total_errors;
        result->flags = r->flags;
    }
}

/* mbe_decode<Codec>[Soft]Frame: the result is reset first, then the arguments are checked (imbe7200x4400.c:709-744) */
static int shim_decode(int codec, int soft, const void* fr, char* d, mbe_process_result* result) {
    int fbits = 0, pbits = 0;
    mbe_b200_geometry(codec, &fbits, &pbits);
    mbe_initProcessResult(result);
    if (!d || !fr) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    uint8_t bits[88];
    mbe_b200_result r;
    mbe_b200_ctx* c = shim_ctx();
    memset(&r, 0, sizeof(r));
    CK(mbe_b200_decode_frames(c, codec, soft, 1, (const uint8_t*)fr, bits, &r));
    if (t_failed) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    if (r.status < 0) {
        return r.status;
    }
    memcpy(d, bits, (size_t)pbits);
    result_out(result, &r);
    return r.status;
}

/* mbe_process<Codec>[Soft]Frame[f] = decode stage, then the Dataf stage on its output (imbe7200x4400.c:933-1007): the
 * decode stage has already written `d` and `result` when the Dataf stage rejects a NULL state pointer */
static int shim_frame(int codec, int soft, float* outf, short* outs, int want_short, mbe_process_result* result,
                      const void* fr, char* d, mbe_parms* cur, mbe_parms* prev, mbe_parms* enh) {
    int fbits = 0, pbits = 0;
    mbe_b200_geometry(codec, &fbits, &pbits);
    if (want_short && !outs) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    if ((!want_short && !outf) || !cur || !prev || !enh) {
        const int st = shim_decode(codec, soft, fr, d, result);
        return st < 0 ? st : MBE_STATUS_INVALID_ARGUMENT;
    }
    mbe_initProcessResult(result);
    if (!d || !fr) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    uint8_t bits[88];
    float pf[NSAMP];
    int16_t ps[NSAMP];
    mbe_b200_result r;
    memset(&r, 0, sizeof(r));
    mbe_parms t[3];
    t[0] = *cur;
    t[1] = *prev;
    t[2] = *enh;
    mbe_b200_ctx* c = shim_ctx();
    rng_ready_locked(c);
    /* state in, one frame, state out: one upload, one download, one synchronisation (mbe_b200_single_frame) */
    CK(mbe_b200_single_frame(c, codec, soft ? 1 : 0, 0, fr, t, t_rng, want_short ? ps : NULL, want_short ? NULL : pf, &r, bits));
    if (t_failed) {
        return MBE_STATUS_INVALID_ARGUMENT;  /* device failure: outputs and state untouched */
    }
    if (r.status >= 0) {
        *cur = t[0];
        *prev = t[1];
        *enh = t[2];
    }
    if (r.status < 0) {
        return r.status;  /* nothing but the (reset) result has been touched, like the reference's early return */
    }
    memcpy(d, bits, (size_t)pbits);
    if (want_short) {
        memcpy(outs, ps, sizeof(ps));
    } else {
        memcpy(outf, pf, sizeof(pf));
    }
    result_out(result, &r);
    return r.status;
}

/* mbe_process<Codec>Data[f] (imbe7200x4400.c:863-924): result is an optional in/out decode context */
static int shim_data(int codec, float* outf, short* outs, int want_short, mbe_process_result* result, const char* d,
                     mbe_parms* cur, mbe_parms* prev, mbe_parms* enh) {
    int fbits = 0, pbits = 0;
    mbe_b200_geometry(codec, &fbits, &pbits);
    if ((want_short ? (void*)outs : (void*)outf) == NULL || !cur || !prev || !enh) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    mbe_b200_result r;
    memset(&r, 0, sizeof(r));
    if (result) {
        r.c0_errors = result->c0_errors;
        r.protected_errors = result->protected_errors;
        r.c4_errors = result->c4_errors;
        r.total_errors = result->total_errors;
        r.flags = result->flags;
    }
    if (!d) {
        return MBE_STATUS_INVALID_ARGUMENT;  /* (an inconsistent result context gives the same code, mbe_result.h:44-97) */
    }
    float pf[NSAMP];
    int16_t ps[NSAMP];
    mbe_parms t[3];
    t[0] = *cur;
    t[1] = *prev;
    t[2] = *enh;
    mbe_b200_ctx* c = shim_ctx();
    rng_ready_locked(c);
    CK(mbe_b200_single_frame(c, codec, 2, 0, d, t, t_rng, want_short ? ps : NULL, want_short ? NULL : pf, &r, NULL));
    if (t_failed) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    if (r.status >= 0) {
        *cur = t[0];
        *prev = t[1];
        *enh = t[2];
    }
    if (r.status < 0) {
        return r.status;
    }
    if (want_short) {
        memcpy(outs, ps, sizeof(ps));
    } else {
        memcpy(outf, pf, sizeof(pf));
    }
    result_out(result, &r);
    return r.status;
}

#define DECODE_FN(name, codec, R, C, N)                                                                               \
    int name##Frame(const char fr[R][C], char d[N], mbe_process_result* result) {                                     \
        return shim_decode(codec, 0, fr, d, result);                                                                  \
    }                                                                                                                 \
    int name##SoftFrame(const mbe_soft_bit fr[R][C], char d[N], mbe_process_result* result) {                         \
        return shim_decode(codec, 1, fr, d, result);                                                                  \
    }
DECODE_FN(mbe_decodeImbe7200x4400, MBE_B200_IMBE7200X4400, 8, 23, 88)
DECODE_FN(mbe_decodeImbe7100x4400, MBE_B200_IMBE7100X4400, 7, 24, 88)
DECODE_FN(mbe_decodeAmbe3600x2400, MBE_B200_AMBE3600X2400, 4, 24, 49)
DECODE_FN(mbe_decodeAmbe3600x2450, MBE_B200_AMBE3600X2450, 4, 24, 49)

#define FRAME_FN(name, codec, R, C, N)                                                                                \
    int name##Framef(float* o, mbe_process_result* res, const char fr[R][C], char d[N], mbe_parms* c, mbe_parms* p,   \
                     mbe_parms* e) {                                                                                  \
        return shim_frame(codec, 0, o, NULL, 0, res, fr, d, c, p, e);                                                 \
    }                                                                                                                 \
    int name##Frame(short* o, mbe_process_result* res, const char fr[R][C], char d[N], mbe_parms* c, mbe_parms* p,    \
                    mbe_parms* e) {                                                                                   \
        return shim_frame(codec, 0, NULL, o, 1, res, fr, d, c, p, e);                                                 \
    }                                                                                                                 \
    int name##SoftFramef(float* o, mbe_process_result* res, const mbe_soft_bit fr[R][C], char d[N], mbe_parms* c,     \
                         mbe_parms* p, mbe_parms* e) {                                                                \
        return shim_frame(codec, 1, o, NULL, 0, res, fr, d, c, p, e);                                                 \
    }                                                                                                                 \
    int name##SoftFrame(short* o, mbe_process_result* res, const mbe_soft_bit fr[R][C], char d[N], mbe_parms* c,      \
                        mbe_parms* p, mbe_parms* e) {                                                                 \
        return shim_frame(codec, 1, NULL, o, 1, res, fr, d, c, p, e);                                                 \
    }
FRAME_FN(mbe_processImbe7200x4400, MBE_B200_IMBE7200X4400, 8, 23, 88)
FRAME_FN(mbe_processImbe7100x4400, MBE_B200_IMBE7100X4400, 7, 24, 88)
FRAME_FN(mbe_processAmbe3600x2400, MBE_B200_AMBE3600X2400, 4, 24, 49)
FRAME_FN(mbe_processAmbe3600x2450, MBE_B200_AMBE3600X2450, 4, 24, 49)

#define DATA_FN(name, codec, N)                                                                                       \
    int name##f(float* o, mbe_process_result* res, const char d[N], mbe_parms* c, mbe_parms* p, mbe_parms* e) {       \
        return shim_data(codec, o, NULL, 0, res, d, c, p, e);                                                         \
    }                                                                                                                 \
    int name(short* o, mbe_process_result* res, const char d[N], mbe_parms* c, mbe_parms* p, mbe_parms* e) {          \
        return shim_data(codec, NULL, o, 1, res, d, c, p, e);                                                         \
    }
DATA_FN(mbe_processImbe4400Data, MBE_B200_IMBE7200X4400, 88)
DATA_FN(mbe_processAmbe2400Data, MBE_B200_AMBE3600X2400, 49)
DATA_FN(mbe_processAmbe2450Data, MBE_B200_AMBE3600X2450, 49)

/* ---- single stages on the caller's parameter sets (mbelib.h:301,385,461,623,725) ---------------------------- */
static int shim_decode_parms(int codec, const char* d, mbe_parms* cur, mbe_parms* prev) {
    if (!cur || !prev || !d) {
        return MBE_STATUS_INVALID_ARGUMENT;  /* imbe7200x4400.c:596-602: state pointers, then the bit array */
    }
    int32_t st = 0;
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_decode_parms(c, codec, 1, (const uint8_t*)d, cur, prev, &st));
    return t_failed ? MBE_STATUS_INVALID_ARGUMENT : st;
}
int mbe_decodeImbe4400Parms(const char* imbe_d, mbe_parms* cur_mp, mbe_parms* prev_mp) {
    return shim_decode_parms(MBE_B200_IMBE7200X4400, imbe_d, cur_mp, prev_mp);
}
int mbe_decodeAmbe2400Parms(const char* ambe_d, mbe_parms* cur_mp, mbe_parms* prev_mp) {
    return shim_decode_parms(MBE_B200_AMBE3600X2400, ambe_d, cur_mp, prev_mp);
}
int mbe_decodeAmbe2450Parms(const char* ambe_d, mbe_parms* cur_mp, mbe_parms* prev_mp) {
    return shim_decode_parms(MBE_B200_AMBE3600X2450, ambe_d, cur_mp, prev_mp);
}

void mbe_spectralAmpEnhance(mbe_parms* cur_mp) { /* mbelib.c:663-666 */
    if (!cur_mp) {
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_spectral_amp_enhance(c, 1, cur_mp, NULL));
}

void mbe_applyAdaptiveSmoothing(mbe_parms* cur_mp, const mbe_parms* prev_mp) { /* mbe_adaptive.c:266-276 */
    if (!cur_mp || !prev_mp) {
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_adaptive_smoothing(c, 1, cur_mp, prev_mp));
}

/* host-only predicates on a parameter set (mbe_adaptive.c:70-107) */
int mbe_requiresAdaptiveSmoothing(const mbe_parms* mp) { return mp ? (mp->errorRate > 0.0125f || mp->errorCountTotal > 4) : 0; }
int mbe_requiresMuting(const mbe_parms* mp) { return mp ? (mp->errorRate > mp->mutingThreshold) : 0; }
int mbe_isMaxFrameRepeat(const mbe_parms* mp) { return mp ? (mp->repeatCount >= 4) : 0; }

/* ---- the channel front-end one step at a time (mbelib.h:286-307,381-387,457-463,531-537) --------------------- */
static int shim_step(int codec, int step, char* fr, char* d) {
    if ((step <= 2 && !fr) || (step >= 2 && !d)) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    int32_t st = 0;
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_channel_step(c, codec, step, 1, (uint8_t*)fr, (uint8_t*)d, &st));
    return t_failed ? MBE_STATUS_INVALID_ARGUMENT : st;
}
#define STEP_FN(name, codec, R, C)                                                                                    \
    int mbe_ecc##name##C0(char fr[R][C]) { return shim_step(codec, 0, (char*)fr, NULL); }                             \
    int mbe_demodulate##name##Data(char fr[R][C]) { return shim_step(codec, 1, (char*)fr, NULL); }                    \
    int mbe_ecc##name##Data(char fr[R][C], char* d) {                                                                 \
        return d ? shim_step(codec, 2, (char*)fr, d) : MBE_STATUS_INVALID_ARGUMENT;                                   \
    }
STEP_FN(Imbe7200x4400, MBE_B200_IMBE7200X4400, 8, 23)
STEP_FN(Imbe7100x4400, MBE_B200_IMBE7100X4400, 7, 24)
STEP_FN(Ambe3600x2400, MBE_B200_AMBE3600X2400, 4, 24)
STEP_FN(Ambe3600x2450, MBE_B200_AMBE3600X2450, 4, 24)
int mbe_convertImbe7100to7200(char* imbe_d) { return shim_step(MBE_B200_IMBE7100X4400, 3, NULL, imbe_d); }

/* ---- tone and comfort-noise generators (mbelib.h:630,638,706,712) -------------------------------------------- */
static void shim_tone(float* aout_buf, const char* ambe_d, mbe_parms* cur_mp, const int32_t* dstar_id) {
    if (!aout_buf) {
        return;
    }
    if (!cur_mp || (!dstar_id && !ambe_d)) {
        mbe_synthesizeSilencef(aout_buf);  /* mbelib.c:771-774,826-829 */
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_synthesize_tone(c, 1, dstar_id ? NULL : (const uint8_t*)ambe_d, dstar_id, cur_mp, aout_buf));
    if (t_failed) {
        mbe_synthesizeSilencef(aout_buf);
    }
}
void mbe_synthesizeTonef(float* aout_buf, const char* ambe_d, mbe_parms* cur_mp) { shim_tone(aout_buf, ambe_d, cur_mp, NULL); }
void mbe_synthesizeTonefdstar(float* aout_buf, const char* ambe_d, mbe_parms* cur_mp, int ID1) {
    const int32_t id = ID1;
    (void)ambe_d;
    shim_tone(aout_buf, NULL, cur_mp, &id);
}

void mbe_synthesizeComfortNoisef(float* aout_buf) { /* mbe_adaptive.c:116-131: advances the calling thread's generator */
    if (!aout_buf) {
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    rng_ready_locked(c);
    CK(mbe_b200_comfort_noise(c, 1, t_rng, aout_buf));
    if (t_failed) {
        mbe_synthesizeSilencef(aout_buf);
    }
}
void mbe_synthesizeComfortNoise(short* aout_buf) { /* mbe_adaptive.c:133-149 */
    float f[NSAMP];
    if (!aout_buf) {
        return;
    }
    mbe_synthesizeComfortNoisef(f);
    mbe_floattoshort(f, aout_buf);
}

/* ---- debug printers: host-only text output on stderr in the reference's formats -------------------------------- */
static void dump_run(const char* row, int hi, int lo, int gap_before) { /* bits hi..lo, a blank before bit gap_before */
    for (int j = hi; j >= lo; --j) {
        if (j == gap_before) {
            fputc(' ', stderr);
        }
        fprintf(stderr, "%i", row[j]);
    }
}
static void dump_flat(const char* d, int n, const int* gaps, int n_gaps) {
    for (int i = 0; i < n; ++i) {
        for (int g = 0; g < n_gaps; ++g) {
            if (gaps[g] == i) {
                fputc(' ', stderr);
            }
        }
        fprintf(stderr, "%i", d[i]);
    }
}
static void dump_ambe_frame(const char fr[4][24]) { /* ambe3600x2450.c:113-142, ambe3600x2400.c:101-130 */
    static const int hi[4] = {23, 22, 10, 13};
    for (int r = 0; r < 4; ++r) {
        fprintf(stderr, "ambe_fr c%d: ", r);
        dump_run(fr[r], hi[r], 0, -1);
        fputc(' ', stderr);
    }
}
void mbe_dumpAmbe2400Data(const char* ambe_d) { dump_flat(ambe_d, 49, NULL, 0); fputc(' ', stderr); }
void mbe_dumpAmbe2450Data(const char* ambe_d) { dump_flat(ambe_d, 49, NULL, 0); fputc(' ', stderr); }
void mbe_dumpAmbe3600x2400Frame(const char ambe_fr[4][24]) { dump_ambe_frame(ambe_fr); }
void mbe_dumpAmbe3600x2450Frame(const char ambe_fr[4][24]) { dump_ambe_frame(ambe_fr); }
void mbe_dumpImbe4400Data(const char* imbe_d) { dump_flat(imbe_d, 88, NULL, 0); }
void mbe_dumpImbe7200x4400Data(const char* imbe_d) { /* imbe7200x4400.c:377-391 */
    static const int gaps[7] = {12, 24, 36, 48, 59, 70, 81};
    dump_flat(imbe_d, 88, gaps, 7);
}
void mbe_dumpImbe7100x4400Data(const char* imbe_d) { /* imbe7100x4400.c:30-44 */
    static const int gaps[6] = {7, 19, 31, 43, 54, 65};
    dump_flat(imbe_d, 88, gaps, 6);
}
void mbe_dumpImbe7200x4400Frame(const char imbe_fr[8][23]) { /* imbe7200x4400.c:397-417 */
    for (int r = 0; r < 7; ++r) {
        dump_run(imbe_fr[r], r < 4 ? 22 : 14, 0, -1);
        fputc(' ', stderr);
    }
    dump_run(imbe_fr[7], 6, 0, -1);
}
void mbe_dumpImbe7100x4400Frame(const char imbe_fr[7][24]) { /* imbe7100x4400.c:50-92 */
    dump_run(imbe_fr[0], 18, 0, 11);
    fputc(' ', stderr);
    dump_run(imbe_fr[1], 23, 0, 11);
    fputc(' ', stderr);
    for (int r = 2; r < 6; ++r) {
        dump_run(imbe_fr[r], r < 4 ? 22 : 14, 0, r < 4 ? 10 : 3);
        fputc(' ', stderr);
    }
    dump_run(imbe_fr[6], 22, 0, -1);
}

/* ---- block decoders (mbelib.h:231-274; src/ecc/ecc.c:221-469) -------------------------------------------- */
static int shim_ecc(int code, int soft, const void* in, char* out, int len) {
    if (!out || !in) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    uint8_t o[23];
    int32_t st = 0;
    memcpy(o, out, (size_t)len);
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_ecc_blocks(c, code, soft, 1, (const uint8_t*)in, o, &st));
    if (t_failed) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    if (st >= 0) {
        memcpy(out, o, (size_t)len);
    }
    return st;
}

int mbe_golay2312(const char* in, char* out) { return shim_ecc(MBE_B200_ECC_GOLAY2312, 0, in, out, 23); }
int mbe_golay2312Soft(const mbe_soft_bit* in, char* out) { return shim_ecc(MBE_B200_ECC_GOLAY2312, 1, in, out, 23); }
int mbe_hamming1511(const char* in, char* out) { return shim_ecc(MBE_B200_ECC_HAMMING1511, 0, in, out, 15); }
int mbe_hamming1511Soft(const mbe_soft_bit* in, char* out) { return shim_ecc(MBE_B200_ECC_HAMMING1511, 1, in, out, 15); }
int mbe_7100x4400hamming1511(const char* in, char* out) { return shim_ecc(MBE_B200_ECC_HAMMING1511_7100, 0, in, out, 15); }
int mbe_7100x4400hamming1511Soft(const mbe_soft_bit* in, char* out) {
    return shim_ecc(MBE_B200_ECC_HAMMING1511_7100, 1, in, out, 15);
}

/* packed 23-bit code word in, corrected 12 data bits out; bits above the code word pass through the shift like in the
 * reference (ecc.c:221-251: databits = block >> 11) */
int mbe_checkGolayBlock(long int* block) {
    if (!block) {
        return MBE_STATUS_INVALID_ARGUMENT;
    }
    const uint32_t b = (uint32_t)(*block);
    char in[23], out[23];
    for (int i = 0; i < 23; ++i) {
        in[i] = (char)((b >> i) & 1u);
    }
    memset(out, 0, sizeof(out));
    const int st = shim_ecc(MBE_B200_ECC_GOLAY2312, 0, in, out, 23);
    if (st < 0) {
        return st;
    }
    uint32_t data = 0;
    for (int i = 22; i >= 11; --i) {
        data = (data << 1) | (uint32_t)(out[i] & 1);
    }
    *block = (long)(int)(((b >> 23) << 12) | data);
    return 0;
}

/* ---- synthesis only ---------------------------------------------------------------------------------------- */
static void shim_synth(float* outf, short* outs, mbe_parms* cur, mbe_parms* prev) { /* mbelib.c:1042-1146 */
    if ((!outf && !outs) || !cur || !prev) {
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    rng_ready_locked(c);
    CK(mbe_b200_synthesize_speech_rng(c, 1, cur, prev, t_rng, outf, (int16_t*)outs));
    if (t_failed) {   /* like mbe_synthesizeSpeechCore on unusable arguments: silence (mbelib.c:1048-1055) */
        if (outf) {
            mbe_synthesizeSilencef(outf);
        }
        if (outs) {
            mbe_synthesizeSilence(outs);
        }
    }
}

void mbe_synthesizeSpeechf(float* aout_buf, mbe_parms* cur_mp, mbe_parms* prev_mp) {
    shim_synth(aout_buf, NULL, cur_mp, prev_mp);
}

void mbe_synthesizeSpeech(short* aout_buf, mbe_parms* cur_mp, mbe_parms* prev_mp) {
    shim_synth(NULL, aout_buf, cur_mp, prev_mp);
}

void mbe_floattoshort(const float* float_buf, short* aout_buf) { /* mbelib.c:1148-1177,1312-1320 */
    if (!float_buf || !aout_buf) {
        return;
    }
    mbe_b200_ctx* c = shim_ctx();
    CK(mbe_b200_floattoshort(c, 1, float_buf, (int16_t*)aout_buf));
    if (t_failed) {
        mbe_synthesizeSilence(aout_buf);
    }
}
