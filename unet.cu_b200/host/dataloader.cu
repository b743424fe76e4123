// Dataset reader for the prepare_data.py format (/root/reference/prepare_data.py:20-38: int32[256] header
// {20240620, N, C, H, W} followed by N*C*H*W float32 in [-1, 1]) with the batching semantics of the reference's
// DataLoader (train_unet.cu:3035-3099): sequential batches of B images, wrap to the start when fewer than B remain.
// Unlike the reference (a blocking fread into one pinned buffer inside the training loop, train_unet.cu:5021-5023) the
// next batch is read by a background thread into another of THREE page-locked buffers while the GPU trains on the
// current one.  Three, not two: the batch returned by a call stays intact until the call AFTER the next one, so the
// caller can hold the current batch and the next one at the same time -- what ub_trainer_set_next_batch needs (the next
// batch's H2D copy runs under the current step).  Data parallel: rank r reads global batches r, r + world, r + 2 world, ...
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "../../include/unet_b200.h"
#include "host_common.h"

struct UbDataLoader {
    FILE* f = nullptr;
    int B = 0, rank = 0, world = 1;
    int n_imgs = 0, C = 0, H = 0, W = 0;
    size_t img_floats = 0;
    long long batches_per_epoch = 0;  // floor(N / B), as the reference's num_batches
    long long next_global = 0;        // index of the next global batch this rank will read (before wrapping)
    static constexpr int kSlots = 3;
    float* buf[kSlots] = {nullptr, nullptr, nullptr};
    bool pinned = false;
    int cur = 0;  // buffer handed to the caller last
    // prefetch thread
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    bool want = false, ready = false, stop = false, io_error = false;
    int fill = 0;
    long long fill_batch = 0;
};

static const int kDataMagic = 20240620;

static bool read_batch(UbDataLoader* d, long long global_batch, float* dst) {
    const long long b = global_batch % d->batches_per_epoch;  // wrap exactly where dataloader_next_batch resets
    const long long off = 1024 + b * (long long)d->B * (long long)d->img_floats * 4;
    if (fseek(d->f, long(off), SEEK_SET) != 0) return false;
    return fread(dst, sizeof(float), size_t(d->B) * d->img_floats, d->f) == size_t(d->B) * d->img_floats;
}

static void worker_main(UbDataLoader* d) {
    std::unique_lock<std::mutex> lk(d->mu);
    for (;;) {
        d->cv.wait(lk, [&] { return d->want || d->stop; });
        if (d->stop) return;
        const int slot = d->fill;
        const long long gb = d->fill_batch;
        d->want = false;
        lk.unlock();
        const bool ok = read_batch(d, gb, d->buf[slot]);
        lk.lock();
        d->io_error = d->io_error || !ok;
        d->ready = true;
        d->cv.notify_all();
    }
}

static void request(UbDataLoader* d, int slot, long long gb) {  // caller holds the lock
    d->fill = slot, d->fill_batch = gb, d->want = true, d->ready = false;
    d->cv.notify_all();
}

extern "C" int ub_dataloader_open(UbDataLoader** out, const char* path, int B, int rank, int world) {
    *out = nullptr;
    if (B < 1 || world < 1 || rank < 0 || rank >= world) {
        ub_host_set_error("dataloader: bad B / rank / world");
        return UB_ERR_SHAPE;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        ub_host_set_error("dataloader: cannot open file");
        return UB_ERR_IO;
    }
    int header[256];
    if (fread(header, sizeof(int), 256, f) != 256 || header[0] != kDataMagic) {
        fclose(f);
        ub_host_set_error("dataloader: bad header / magic (expected 20240620)");
        return UB_ERR_IO;
    }
    UbDataLoader* d = new UbDataLoader();
    d->f = f, d->B = B, d->rank = rank, d->world = world;
    d->n_imgs = header[1], d->C = header[2], d->H = header[3], d->W = header[4];
    d->img_floats = size_t(d->C) * d->H * d->W;
    fseek(f, 0, SEEK_END);
    const long long fsz = ftell(f);
    if (d->n_imgs < B || d->img_floats == 0 || (fsz - 1024) / (long long)(d->img_floats * 4) < d->n_imgs) {
        fclose(f);
        delete d;
        ub_host_set_error("dataloader: file shorter than its header says, or fewer images than one batch");
        return UB_ERR_IO;
    }
    d->batches_per_epoch = d->n_imgs / B;
    const size_t bytes = size_t(B) * d->img_floats * sizeof(float);
    // page-locked when a CUDA device is there (async H2D), plain aligned memory otherwise (CPU-only tests)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
        d->pinned = true;
        for (int k = 0; k < UbDataLoader::kSlots; ++k)
            if (cudaMallocHost(&d->buf[k], bytes) != cudaSuccess) d->pinned = false, d->buf[k] = nullptr;
        if (!d->pinned)
            for (int k = 0; k < UbDataLoader::kSlots; ++k)
                if (d->buf[k]) cudaFreeHost(d->buf[k]), d->buf[k] = nullptr;
    }
    if (!d->pinned) {
        cudaGetLastError();
        for (int k = 0; k < UbDataLoader::kSlots; ++k)
            d->buf[k] = static_cast<float*>(aligned_alloc(4096, (bytes + 4095) & ~size_t(4095)));
    }
    if (!d->buf[0] || !d->buf[1] || !d->buf[2]) {
        ub_dataloader_close(d);
        ub_host_set_error("dataloader: out of host memory");
        return UB_ERR_IO;
    }
    d->next_global = rank;
    d->worker = std::thread(worker_main, d);
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->cur = UbDataLoader::kSlots - 1;  // so that the first batch lands in buffer 0
        request(d, 0, d->next_global);
    }
    *out = d;
    return UB_OK;
}

extern "C" int ub_dataloader_info(UbDataLoader* d, int* n_imgs, int* C, int* H, int* W, long long* batches_per_epoch) {
    if (n_imgs) *n_imgs = d->n_imgs;
    if (C) *C = d->C;
    if (H) *H = d->H;
    if (W) *W = d->W;
    if (batches_per_epoch) *batches_per_epoch = d->batches_per_epoch;
    return UB_OK;
}

extern "C" const float* ub_dataloader_next(UbDataLoader* d) {
    std::unique_lock<std::mutex> lk(d->mu);
    d->cv.wait(lk, [&] { return d->ready; });
    if (d->io_error) {
        ub_host_set_error("dataloader: read failed");
        return nullptr;
    }
    d->cur = d->fill;
    d->next_global += d->world;
    // prefetch the following batch into the buffer handed out two calls ago (the previous one stays intact)
    request(d, (d->cur + 1) % UbDataLoader::kSlots, d->next_global);
    return d->buf[d->cur];
}

extern "C" void ub_dataloader_reset(UbDataLoader* d) {
    std::unique_lock<std::mutex> lk(d->mu);
    d->cv.wait(lk, [&] { return d->ready; });  // let the in-flight read finish
    d->next_global = d->rank;
    request(d, (d->cur + 1) % UbDataLoader::kSlots, d->next_global);
}

extern "C" void ub_dataloader_close(UbDataLoader* d) {
    if (!d) return;
    if (d->worker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(d->mu);
            d->stop = true;
            d->cv.notify_all();
        }
        d->worker.join();
    }
    for (float*& b : d->buf) {
        if (!b) continue;
        if (d->pinned)
            cudaFreeHost(b);
        else
            free(b);
        b = nullptr;
    }
    if (d->f) fclose(d->f);
    delete d;
}
