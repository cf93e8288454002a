/*
 * mbe_b200_compat.h - single-stream drop-in shim: the reference's per-frame C API
 * (/root/reference/include/mbelib-neo/mbelib.h) served by a batch of one on the GPU.
 *
 * libmbe-neo-b200shim.so exports, under the reference's own names and signatures, the entry points of the decode-and-
 * synthesis hot path plus the host-only helpers their callers need.  A program built against the reference's header can
 * be linked against the shim instead of libmbe-neo (the structs below are layout-compatible with mbelib.h:88-191); the
 * reference's own test binaries test_api, test_ecc, test_floattoshort_parity, test_frame_paths,
 * test_golden_pcm, test_input_validation, test_noise_determinism and test_params run unmodified on top of it (tests/test_gpu_shim.py).
 *
 * Every call moves the caller-owned mbe_parms triplet and the calling thread's RNG words to the device, runs ONE frame
 * through the same kernels as the batched API and moves the state back: it is latency-bound (tens of microseconds of
 * copies and launch per 20 ms frame) and exists for API completeness and for testing - use mbe_b200.h for throughput.
 * There is no CPU fallback: without a CUDA device the first call prints the error and aborts.
 *
 * All 87 functions of the reference's header are provided.  The steps inside the path (block decoders, C0 / de-scramble /
 * data ECC, parameter decoders, enhancement, smoothing, tone and comfort-noise generators) run the same device functions
 * the fused frame kernels use, one call = one small launch.
 */
#ifndef MBE_B200_COMPAT_H
#define MBE_B200_COMPAT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#ifndef MBE_B200_COMPAT_NO_TYPES /* define when the reference's own mbelib.h is included as well */
/* mbelib.h:88-139, 2604 bytes; arrays are indexed 1..56, element 0 is scratch */
struct mbe_parameters {
    float w0;
    int L;
    int K;
    int Vl[57];
    float Ml[57];
    float log2Ml[57];
    float PHIl[57];
    float PSIl[57];
    float gamma;
    uint32_t tonePhase;
    int swn;
    float localEnergy;
    int amplitudeThreshold;
    float errorRate;
    int errorCountTotal;
    int errorCount4;
    int repeatCount;
    float mutingThreshold;
    float previousUw[256];
    float noiseSeed;
    float noiseOverlap[96];
};
typedef struct mbe_parameters mbe_parms;

typedef struct mbe_soft_bit { /* mbelib.h:148-151 */
    uint8_t bit;
    uint8_t reliability;
} mbe_soft_bit;

#define MBE_PROCESS_FLAG_SOFT_INPUT 0x0001u /* mbelib.h:153-166 */
#define MBE_PROCESS_FLAG_C0_VALID   0x0002u
#define MBE_PROCESS_FLAG_C4_VALID   0x0004u
#define MBE_PROCESS_FLAG_TONE       0x0010u
#define MBE_PROCESS_FLAG_ERASURE    0x0020u
#define MBE_PROCESS_FLAG_REPEAT     0x0040u
#define MBE_PROCESS_FLAG_MUTE       0x0080u
#define MBE_STATUS_INVALID_ARGUMENT (-1)
#define MBE_STATUS_INVALID_BITS     (-2)

typedef struct mbe_process_result { /* mbelib.h:180-191 */
    int c0_errors;
    int protected_errors;
    int c4_errors;
    int total_errors;
    unsigned flags;
} mbe_process_result;
#endif

#define MBE_COMPAT_API __attribute__((visibility("default")))

/* host-only helpers (mbelib.h:194-224,588,602,608,640,642) */
MBE_COMPAT_API void mbe_initProcessResult(mbe_process_result* result);
MBE_COMPAT_API void mbe_formatProcessResult(char* str, size_t size, const mbe_process_result* result);
MBE_COMPAT_API mbe_soft_bit mbe_softBitFromHard(int bit, uint8_t reliability);
MBE_COMPAT_API mbe_soft_bit mbe_softBitFromLlr(int16_t llr);
MBE_COMPAT_API int mbe_softBitsFromHard(const char* bits, mbe_soft_bit* soft, size_t count, uint8_t reliability);
MBE_COMPAT_API int mbe_softBitsFromLlr(const int16_t* llr, mbe_soft_bit* soft, size_t count);
MBE_COMPAT_API const char* mbe_versionString(void);
MBE_COMPAT_API void mbe_moveMbeParms(const mbe_parms* source_mp, mbe_parms* destination_mp);
MBE_COMPAT_API void mbe_useLastMbeParms(mbe_parms* cur_mp, const mbe_parms* prev_mp);
MBE_COMPAT_API void mbe_synthesizeSilencef(float* aout_buf);
MBE_COMPAT_API void mbe_synthesizeSilence(short* aout_buf);

/* state (mbelib.h:596,615) */
MBE_COMPAT_API void mbe_setThreadRngSeed(uint32_t seed);
MBE_COMPAT_API void mbe_initMbeParms(mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);

/* single stages on the caller's parameter sets (mbelib.h:301,385,461,623,693,700,725,732) */
MBE_COMPAT_API int mbe_decodeImbe4400Parms(const char* imbe_d, mbe_parms* cur_mp, mbe_parms* prev_mp);
MBE_COMPAT_API int mbe_decodeAmbe2400Parms(const char* ambe_d, mbe_parms* cur_mp, mbe_parms* prev_mp);
MBE_COMPAT_API int mbe_decodeAmbe2450Parms(const char* ambe_d, mbe_parms* cur_mp, mbe_parms* prev_mp);
MBE_COMPAT_API void mbe_spectralAmpEnhance(mbe_parms* cur_mp);
MBE_COMPAT_API void mbe_applyAdaptiveSmoothing(mbe_parms* cur_mp, const mbe_parms* prev_mp);
MBE_COMPAT_API int mbe_requiresAdaptiveSmoothing(const mbe_parms* mp);
MBE_COMPAT_API int mbe_requiresMuting(const mbe_parms* mp);
MBE_COMPAT_API int mbe_isMaxFrameRepeat(const mbe_parms* mp);

/* the channel front-end one step at a time (mbelib.h:286-307,381-387,457-463,531-537) */
#define MBE_COMPAT_STEPS(name, R, C)                                                                                  \
    MBE_COMPAT_API int mbe_ecc##name##C0(char fr[R][C]);                                                              \
    MBE_COMPAT_API int mbe_demodulate##name##Data(char fr[R][C]);                                                     \
    MBE_COMPAT_API int mbe_ecc##name##Data(char fr[R][C], char* d);
MBE_COMPAT_STEPS(Imbe7200x4400, 8, 23)
MBE_COMPAT_STEPS(Imbe7100x4400, 7, 24)
MBE_COMPAT_STEPS(Ambe3600x2400, 4, 24)
MBE_COMPAT_STEPS(Ambe3600x2450, 4, 24)
MBE_COMPAT_API int mbe_convertImbe7100to7200(char* imbe_d);

/* tone and comfort-noise generators (mbelib.h:630,638,706,712) */
MBE_COMPAT_API void mbe_synthesizeTonef(float* aout_buf, const char* ambe_d, mbe_parms* cur_mp);
MBE_COMPAT_API void mbe_synthesizeTonefdstar(float* aout_buf, const char* ambe_d, mbe_parms* cur_mp, int ID1);
MBE_COMPAT_API void mbe_synthesizeComfortNoisef(float* aout_buf);
MBE_COMPAT_API void mbe_synthesizeComfortNoise(short* aout_buf);

/* debug printers (stderr; mbelib.h:278-280,377-379,451-455,527-529) */
MBE_COMPAT_API void mbe_dumpAmbe2400Data(const char* ambe_d);
MBE_COMPAT_API void mbe_dumpAmbe3600x2400Frame(const char ambe_fr[4][24]);
MBE_COMPAT_API void mbe_dumpAmbe2450Data(const char* ambe_d);
MBE_COMPAT_API void mbe_dumpAmbe3600x2450Frame(const char ambe_fr[4][24]);
MBE_COMPAT_API void mbe_dumpImbe4400Data(const char* imbe_d);
MBE_COMPAT_API void mbe_dumpImbe7200x4400Data(const char* imbe_d);
MBE_COMPAT_API void mbe_dumpImbe7200x4400Frame(const char imbe_fr[8][23]);
MBE_COMPAT_API void mbe_dumpImbe7100x4400Data(const char* imbe_d);
MBE_COMPAT_API void mbe_dumpImbe7100x4400Frame(const char imbe_fr[7][24]);

/* block decoders (mbelib.h:231-274) */
MBE_COMPAT_API int mbe_checkGolayBlock(long int* block);
MBE_COMPAT_API int mbe_golay2312(const char* in, char* out);
MBE_COMPAT_API int mbe_golay2312Soft(const mbe_soft_bit* in, char* out);
MBE_COMPAT_API int mbe_hamming1511(const char* in, char* out);
MBE_COMPAT_API int mbe_hamming1511Soft(const mbe_soft_bit* in, char* out);
MBE_COMPAT_API int mbe_7100x4400hamming1511(const char* in, char* out);
MBE_COMPAT_API int mbe_7100x4400hamming1511Soft(const mbe_soft_bit* in, char* out);

/* ECC stage (mbelib.h:315,323,395,403,471,479,545,553) */
MBE_COMPAT_API int mbe_decodeImbe7200x4400Frame(const char imbe_fr[8][23], char imbe_d[88], mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeImbe7200x4400SoftFrame(const mbe_soft_bit imbe_fr[8][23], char imbe_d[88],
                                                    mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeImbe7100x4400Frame(const char imbe_fr[7][24], char imbe_d[88], mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeImbe7100x4400SoftFrame(const mbe_soft_bit imbe_fr[7][24], char imbe_d[88],
                                                    mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeAmbe3600x2400Frame(const char ambe_fr[4][24], char ambe_d[49], mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeAmbe3600x2400SoftFrame(const mbe_soft_bit ambe_fr[4][24], char ambe_d[49],
                                                    mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeAmbe3600x2450Frame(const char ambe_fr[4][24], char ambe_d[49], mbe_process_result* result);
MBE_COMPAT_API int mbe_decodeAmbe3600x2450SoftFrame(const mbe_soft_bit ambe_fr[4][24], char ambe_d[49],
                                                    mbe_process_result* result);

/* parameter bits -> PCM (mbelib.h:335-342,415-419,491-495) */
#define MBE_COMPAT_DATA(name, n)                                                                                     \
    MBE_COMPAT_API int name##f(float* aout_buf, mbe_process_result* result, const char d[n], mbe_parms* cur_mp,     \
                               mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);                                     \
    MBE_COMPAT_API int name(short* aout_buf, mbe_process_result* result, const char d[n], mbe_parms* cur_mp,        \
                            mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);
MBE_COMPAT_DATA(mbe_processImbe4400Data, 88)
MBE_COMPAT_DATA(mbe_processAmbe2400Data, 49)
MBE_COMPAT_DATA(mbe_processAmbe2450Data, 49)

/* frame -> PCM (mbelib.h:352-373,429-447,505-523,564-582) */
#define MBE_COMPAT_FRAME(name, R, C, N)                                                                              \
    MBE_COMPAT_API int name##Framef(float* aout_buf, mbe_process_result* result, const char fr[R][C], char d[N],    \
                                    mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);             \
    MBE_COMPAT_API int name##Frame(short* aout_buf, mbe_process_result* result, const char fr[R][C], char d[N],     \
                                   mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);              \
    MBE_COMPAT_API int name##SoftFramef(float* aout_buf, mbe_process_result* result, const mbe_soft_bit fr[R][C],   \
                                        char d[N], mbe_parms* cur_mp, mbe_parms* prev_mp,                            \
                                        mbe_parms* prev_mp_enhanced);                                                \
    MBE_COMPAT_API int name##SoftFrame(short* aout_buf, mbe_process_result* result, const mbe_soft_bit fr[R][C],    \
                                       char d[N], mbe_parms* cur_mp, mbe_parms* prev_mp,                             \
                                       mbe_parms* prev_mp_enhanced);
MBE_COMPAT_FRAME(mbe_processImbe7200x4400, 8, 23, 88)
MBE_COMPAT_FRAME(mbe_processImbe7100x4400, 7, 24, 88)
MBE_COMPAT_FRAME(mbe_processAmbe3600x2400, 4, 24, 49)
MBE_COMPAT_FRAME(mbe_processAmbe3600x2450, 4, 24, 49)

/* synthesis only (mbelib.h:652,662,675) */
MBE_COMPAT_API void mbe_synthesizeSpeechf(float* aout_buf, mbe_parms* cur_mp, mbe_parms* prev_mp);
MBE_COMPAT_API void mbe_synthesizeSpeech(short* aout_buf, mbe_parms* cur_mp, mbe_parms* prev_mp);
MBE_COMPAT_API void mbe_floattoshort(const float* float_buf, short* aout_buf);

#ifdef __cplusplus
}
#endif
#endif /* MBE_B200_COMPAT_H */
