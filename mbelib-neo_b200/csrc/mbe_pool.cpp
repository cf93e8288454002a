// Many GPUs from one process: a pool of per-GPU contexts that shards the global stream range into
// contiguous blocks, one block per device, and runs a batched call on every device that owns part of it, concurrently,
// from one host thread per device.  Streams are independent, so nothing is exchanged between GPUs: the host only
// scatters the input bit buffers and gathers PCM / results (plain pointer offsets into the caller's arrays).
// Host-only C++ on top of the C-ABI of include/mbe_b200.h; no CUDA headers.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/mbe_b200.h"

struct mbe_b200_pool {
    std::vector<mbe_b200_ctx*> ctx;
    std::vector<int> first;   // global id of the first stream of each shard (+ one past the end)
    int max_streams;
    char err[320];
};

static char g_pool_err[320] = "";

extern "C" {

int mbe_b200_pool_create(mbe_b200_pool** out, int n_devices, const int* device_ordinals, int max_streams) {
    if (!out || n_devices < 0 || max_streams <= 0) {
        snprintf(g_pool_err, sizeof(g_pool_err), "mbe_b200_pool_create: bad argument");
        return MBE_B200_E_ARG;
    }
    *out = nullptr;
    if (n_devices == 0) {
        n_devices = mbe_b200_device_count();
        device_ordinals = nullptr;
        if (n_devices <= 0) {
            snprintf(g_pool_err, sizeof(g_pool_err), "mbe_b200_pool_create: no CUDA device (there is no CPU fallback)");
            return MBE_B200_E_NOGPU;
        }
    }
    n_devices = std::min(n_devices, max_streams);
    mbe_b200_pool* p = new mbe_b200_pool();
    p->max_streams = max_streams;
    p->err[0] = 0;
    // base + remainder, like sharding.shard_range: n_devices <= max_streams, so no shard is empty and every entry of
    // p->ctx is a live context (9 streams on 8 GPUs: 2 1 1 1 1 1 1 1, not 2 2 2 2 1 0 0 0)
    const int base = max_streams / n_devices, rem = max_streams % n_devices;
    for (int i = 0; i < n_devices; ++i) {
        const int lo = i * base + std::min(i, rem), hi = lo + base + (i < rem ? 1 : 0);
        p->first.push_back(lo);
        mbe_b200_ctx* c = nullptr;
        const int rc = mbe_b200_create(&c, device_ordinals ? device_ordinals[i] : i, hi - lo);
        if (rc != 0) {
            snprintf(g_pool_err, sizeof(g_pool_err), "mbe_b200_pool_create: shard %d: %s", i, mbe_b200_last_error(nullptr));
            for (mbe_b200_ctx* q : p->ctx) {
                mbe_b200_destroy(q);
            }
            delete p;
            return rc;
        }
        p->ctx.push_back(c);
    }
    p->first.push_back(max_streams);
    *out = p;
    return 0;
}

void mbe_b200_pool_destroy(mbe_b200_pool* p) {
    if (!p) {
        return;
    }
    for (mbe_b200_ctx* c : p->ctx) {
        mbe_b200_destroy(c);
    }
    delete p;
}

const char* mbe_b200_pool_last_error(const mbe_b200_pool* p) { return p ? p->err : g_pool_err; }

int mbe_b200_pool_shards(const mbe_b200_pool* p) { return p ? (int)p->ctx.size() : 0; }

int mbe_b200_pool_shard(const mbe_b200_pool* p, int shard, int* first_stream, int* n_streams, mbe_b200_ctx** ctx) {
    if (!p || shard < 0 || shard >= (int)p->ctx.size()) {
        return MBE_B200_E_ARG;
    }
    if (first_stream) {
        *first_stream = p->first[shard];
    }
    if (n_streams) {
        *n_streams = p->first[shard + 1] - p->first[shard];
    }
    if (ctx) {
        *ctx = p->ctx[shard];
    }
    return 0;
}

}  // extern "C"

// run fn(shard, ctx, local first stream, count, offset of that block inside the caller's range) on every shard that
// overlaps [first, first + count), one host thread per shard; the first failure wins
template <class F>
static int for_each_shard(mbe_b200_pool* p, const char* what, int first, int count, F fn) {
    if (!p) {
        return MBE_B200_E_ARG;
    }
    if (first < 0 || count < 0 || first > p->max_streams - count) {
        snprintf(p->err, sizeof(p->err), "%s: stream range out of bounds", what);
        return MBE_B200_E_ARG;
    }
    const int n = (int)p->ctx.size();
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) {
        const int lo = std::max(first, p->first[i]), hi = std::min(first + count, p->first[i + 1]);
        if (hi <= lo || !p->ctx[i]) {
            continue;
        }
        th.emplace_back([=, &rc]() { rc[i] = fn(i, p->ctx[i], lo - p->first[i], hi - lo, (size_t)(lo - first)); });
    }
    for (std::thread& t : th) {
        t.join();
    }
    for (int i = 0; i < n; ++i) {
        if (rc[i] != 0) {
            snprintf(p->err, sizeof(p->err), "%s: shard %d: %s", what, i, mbe_b200_last_error(p->ctx[i]));
            return rc[i];
        }
    }
    return 0;
}

extern "C" {

int mbe_b200_pool_init_streams(mbe_b200_pool* p, int first_stream, int count, const uint32_t* seeds) {
    return for_each_shard(p, "pool_init_streams", first_stream, count,
                          [=](int, mbe_b200_ctx* c, int lo, int n, size_t off) {
                              return mbe_b200_init_streams(c, lo, n, seeds ? seeds + off : nullptr);
                          });
}

int mbe_b200_pool_export_state(mbe_b200_pool* p, int first_stream, int count, void* parms_triplets) {
    return for_each_shard(p, "pool_export_state", first_stream, count,
                          [=](int, mbe_b200_ctx* c, int lo, int n, size_t off) {
                              return mbe_b200_export_state(c, lo, n, (uint8_t*)parms_triplets + off * 3 * MBE_B200_PARMS_BYTES);
                          });
}

int mbe_b200_pool_import_state(mbe_b200_pool* p, int first_stream, int count, const void* parms_triplets) {
    return for_each_shard(p, "pool_import_state", first_stream, count,
                          [=](int, mbe_b200_ctx* c, int lo, int n, size_t off) {
                              return mbe_b200_import_state(c, lo, n,
                                                           (const uint8_t*)parms_triplets + off * 3 * MBE_B200_PARMS_BYTES);
                          });
}

// packed < 0: one byte per bit (soft = 0) or mbe_soft_bit pairs (soft = 1); packed = 1: eight hard bits per byte
static int pool_process(mbe_b200_pool* p, const char* what, int codec, int soft, int packed, int first_stream, int n_streams,
                        int n_frames, const uint8_t* frames, int16_t* pcm, float* pcmf, mbe_b200_result* results,
                        uint8_t* bits) {
    int fbits = 0, pbits = 0;
    if (p && (mbe_b200_geometry(codec, &fbits, &pbits) != 0 || n_frames < 0 || !frames)) {
        snprintf(p->err, sizeof(p->err), "%s: bad argument", what);
        return MBE_B200_E_ARG;
    }
    const size_t fbytes = packed ? (size_t)mbe_b200_channel_frame_bytes(p ? p->ctx[0] : nullptr, codec)
                                 : (size_t)fbits * (soft ? 2u : 1u);
    const size_t F = (size_t)n_frames;
    return for_each_shard(p, what, first_stream, n_streams, [=](int, mbe_b200_ctx* c, int lo, int n, size_t off) {
        const uint8_t* fr = frames + off * F * fbytes;
        int16_t* o16 = pcm ? pcm + off * F * MBE_B200_SAMPLES_PER_FRAME : nullptr;
        float* of = pcmf ? pcmf + off * F * MBE_B200_SAMPLES_PER_FRAME : nullptr;
        mbe_b200_result* r = results ? results + off * F : nullptr;
        uint8_t* b = bits ? bits + off * F * (size_t)pbits : nullptr;
        return packed ? mbe_b200_process_frames_packed(c, codec, lo, n, n_frames, fr, o16, of, r, b)
                      : mbe_b200_process_frames(c, codec, soft, lo, n, n_frames, fr, o16, of, r, b);
    });
}

int mbe_b200_pool_set_channel_map(mbe_b200_pool* p, int codec, const uint16_t* map, int n_bits) {
    if (!p) {
        return MBE_B200_E_ARG;
    }
    for (size_t i = 0; i < p->ctx.size(); ++i) {
        if (!p->ctx[i]) {
            continue;
        }
        const int rc = mbe_b200_set_channel_map(p->ctx[i], codec, map, n_bits);
        if (rc != 0) {
            snprintf(p->err, sizeof(p->err), "pool_set_channel_map: shard %d: %s", (int)i, mbe_b200_last_error(p->ctx[i]));
            return rc;
        }
    }
    return 0;
}

int mbe_b200_pool_process_frames(mbe_b200_pool* p, int codec, int soft, int first_stream, int n_streams, int n_frames,
                                 const uint8_t* frames, int16_t* pcm, float* pcmf, mbe_b200_result* results, uint8_t* bits) {
    return pool_process(p, "pool_process_frames", codec, soft ? 1 : 0, 0, first_stream, n_streams, n_frames, frames, pcm, pcmf,
                        results, bits);
}

int mbe_b200_pool_process_frames_packed(mbe_b200_pool* p, int codec, int first_stream, int n_streams, int n_frames,
                                        const uint8_t* packed, int16_t* pcm, float* pcmf, mbe_b200_result* results,
                                        uint8_t* bits) {
    return pool_process(p, "pool_process_frames_packed", codec, 0, 1, first_stream, n_streams, n_frames, packed, pcm, pcmf,
                        results, bits);
}

}  // extern "C"
