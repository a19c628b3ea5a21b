// tests/emu/hash_service_emu.cpp -- CPU harness of the frame submission service (TEST INFRASTRUCTURE).
// Instantiates hydrus_video_deduplicator_b200/csrc/hash_service.h -- the very template the CUDA library builds --
// with a mock device (host memory for the "device" ring, a launch that completes a few polls later and "hashes" a
// frame by copying its first 32 bytes) so that push order, ring wrap, back-pressure, the consumed watermark, reuse
// after finish and concurrent hashers are exercised without a GPU.
// Build: g++ -O1 -std=c++17 -pthread -shared -fPIC -o libhash_service_emu.so hash_service_emu.cpp
#include <stdlib.h>

#include <random>

#include "../../hydrus_video_deduplicator_b200/csrc/hash_service.h"

using namespace vpdq_service;

namespace {
struct MockDev {
    struct Event {
        int polls_left = 0;
    };
    size_t fb = 0, n = 0;
    uint8_t *h_frames = nullptr, *d_frames = nullptr, *h_hash = nullptr;
    int32_t* h_quality = nullptr;
    int latency = 3;
    std::atomic<long long> launches{0}, launched_frames{0}, max_launch{0};
    int fail_after = -1;  // launch number that reports a device error (-1: never)

    int alloc(size_t n_, size_t fb_, uint8_t** hf, uint8_t** df, uint8_t** hh, int32_t** hq) {
        n = n_;
        fb = fb_;
        h_frames = (uint8_t*)malloc(n * fb);
        d_frames = (uint8_t*)malloc(n * fb);
        h_hash = (uint8_t*)malloc(n * 32);
        h_quality = (int32_t*)malloc(n * 4);
        memset(d_frames, 0xEE, n * fb);
        *hf = h_frames; *df = d_frames; *hh = h_hash; *hq = h_quality;
        return 0;
    }
    void thread_init() {}
    int upload(uint8_t* d, const uint8_t* h, size_t bytes) {
        memcpy(d, h, bytes);
        return 0;
    }
    int launch(size_t first, size_t cnt, Event* ev) {
        const long long k = launches.fetch_add(1);
        if (fail_after >= 0 && k == fail_after) return -2;
        launched_frames += (long long)cnt;
        if ((long long)cnt > max_launch.load()) max_launch = (long long)cnt;
        for (size_t i = first; i < first + cnt; ++i) {
            memcpy(h_hash + i * 32, d_frames + i * fb, 32);     // "hash" = the frame's first 32 bytes
            int32_t q;
            memcpy(&q, d_frames + i * fb + 32, 4);               // "quality" = the next 4
            h_quality[i] = q;
        }
        ev->polls_left = latency;
        return 0;
    }
    bool is_done(Event& e, int*) { return --e.polls_left <= 0; }
    int launch_status() { return 0; }
    void release(Event&) {}
    void idle_pause(bool) { std::this_thread::yield(); }
};

// frame content: bytes 0..7 = hasher id, 8..15 = sequence number, 32..35 = quality
void make_frame(uint8_t* f, size_t fb, long long id, long long seq, int quality) {
    memset(f, (int)(seq & 0xFF), fb);
    memcpy(f, &id, 8);
    memcpy(f + 8, &seq, 8);
    memcpy(f + 32, &quality, 4);
}
}  // namespace

// Runs `threads` caller threads; each hashes `videos` videos with frame counts drawn from `counts` (cycled), pushing
// with a mix of copy / nocopy pushes, and checks every result.  Returns the number of errors found (0 = pass);
// stats[0] = launches, stats[1] = frames launched, stats[2] = largest launch.
extern "C" __attribute__((visibility("default"))) int emu_service_run(int arena_frames, int workers, int threads,
                                                                       int videos, const int* counts, int n_counts,
                                                                       int frame_bytes, long long* stats) {
    MockDev dev;
    Config cfg;
    cfg.frame_bytes = (size_t)frame_bytes;
    cfg.arena_frames = arena_frames;
    cfg.copy_workers = workers;
    cfg.copy_parts = 4;
    cfg.max_launch = arena_frames;
    cfg.max_inflight = 2;
    HashService<MockDev> svc(&dev, cfg);
    if (svc.start()) return 1000000;
    std::atomic<int> errors{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            HasherState* h = new HasherState;  // one handle, reused for every video of this thread (as after finish())
            for (int v = 0; v < videos; ++v) {
                const int n = counts[(t * videos + v) % n_counts];
                const long long id = (long long)t * 100000 + v;
                std::vector<uint8_t> buf((size_t)n * frame_bytes);
                for (int k = 0; k < n; ++k) make_frame(buf.data() + (size_t)k * frame_bytes, frame_bytes, id, k, k % 101);
                // push: alternate single-frame nocopy, single-frame copy, and a batch copy of the rest
                int k = 0;
                for (; k < n && k < 5; ++k)
                    if (svc.push(h, buf.data() + (size_t)k * frame_bytes, 1, (k & 1) != 0)) ++errors;
                if (k < n && svc.push(h, buf.data() + (size_t)k * frame_bytes, n - k, (v & 1) != 0)) ++errors;
                if (svc.wait_all(h)) ++errors;
                std::unique_lock<std::mutex> lk(h->mu);
                if (h->pushed != n || h->done != n) ++errors;
                if (h->consumed.load() != n) ++errors;  // every source byte has been copied out when finish returns
                for (int j = 0; j < n; ++j) {
                    long long got_id, got_seq;
                    memcpy(&got_id, h->hashes.data() + (size_t)j * 32, 8);
                    memcpy(&got_seq, h->hashes.data() + (size_t)j * 32 + 8, 8);
                    if (got_id != id || got_seq != j || h->quality[j] != j % 101) ++errors;
                }
                lk.unlock();
                h->reset();
            }
            delete h;
        });
    for (auto& th : pool) th.join();
    svc.stop();
    if (stats) {
        stats[0] = dev.launches.load();
        stats[1] = dev.launched_frames.load();
        stats[2] = dev.max_launch.load();
    }
    return errors.load();
}

// a device error in launch number `fail_after` must surface in wait_all / later pushes, never hang
extern "C" __attribute__((visibility("default"))) int emu_service_failure(int fail_after) {
    MockDev dev;
    dev.fail_after = fail_after;
    Config cfg;
    cfg.frame_bytes = 4096;
    cfg.arena_frames = 8;
    cfg.copy_workers = 2;
    cfg.max_launch = 8;
    HashService<MockDev> svc(&dev, cfg);
    if (svc.start()) return -100;
    HasherState h;
    std::vector<uint8_t> buf(40 * 4096);
    for (int k = 0; k < 40; ++k) make_frame(buf.data() + (size_t)k * 4096, 4096, 7, k, 50);
    int push_rc = 0;
    for (int k = 0; k < 40 && !push_rc; ++k) push_rc = svc.push(&h, buf.data() + (size_t)k * 4096, 1, true);
    const int wait_rc = svc.wait_all(&h);
    svc.stop();
    return (wait_rc != 0 ? 1 : 0) | (push_rc != 0 ? 2 : 0) | (svc.broken() != 0 ? 4 : 0);
}
