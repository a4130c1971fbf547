// files.cu -- conversion of inputs of any size over every device of the process:
// the two headline directions as a pipeline of slabs, on top of the public C ABI of this library.
//
//   .bin -> .binpack   the slabs are the "ranks" of the sharded compressor (include/nnuepack.h): every slab
//                      is read with one halo record in front and an overlap window behind, chains belong
//                      to the slab that holds their head. The heavy part of a slab (K1, payload scan and
//                      write: nnp_shard_compress_begin_dev) runs on whichever device the slab was dealt to,
//                      concurrently with the other devices; the chunk-flush rule (compress_file.cpp:1076-1080)
//                      is then replayed slab after slab in file order (a few hundred microseconds each): the
//                      carry, the payload base and the number of chunks so far travel from slab to slab in
//                      host memory -- the devices share an address space, no collective is needed. The size
//                      field of a slab's last chunk is only known once a later slab opens the next chunk: it
//                      is written as a placeholder and patched in the output. The result is the file ONE
//                      reference run writes (compressBin :1338-1374), whatever the slab size and the number
//                      of devices.
//   .binpack -> .bin   chunks are independent (decompressBin :1376-1412; hasNextChunk / readNextChunk
//                      :468-480): the chunk headers are walked on the host, whole chunks are grouped into
//                      slabs, the slabs are dealt to the devices, and a slab's records land at 40 x the
//                      positions of the slabs in front of it.
//
// Per device three host threads work on double buffers, so that reading + H2D of slab k + 1, the kernels
// of slab k and D2H + writing of slab k - 1 overlap:
//
//   loader   source -> pinned host buffer -> cudaMemcpyAsync -> device buffer
//   compute  the *_dev entry points of this library on the device the thread is bound to
//   drainer  device buffer -> cudaMemcpyAsync -> pinned host buffer -> sink
//
// Sources and sinks are files (pread / pwrite, large reads split over several threads) or host memory
// (the *_multi entry points). Host code only; every conversion runs in the CUDA kernels behind the
// *_dev entry points.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/nnuepack.h"

namespace {

constexpr uint64_t NONE = ~(uint64_t)0;
constexpr uint64_t NO_CARRY = ~(uint64_t)0;

// ------------------------------------------------------------------------------------------------ I/O

bool pread_all(int fd, void* dst, size_t n, uint64_t off)
{
    char* p = static_cast<char*>(dst);
    while (n > 0) {
        const ssize_t r = ::pread(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r;
        off += (uint64_t)r;
        n -= (size_t)r;
    }
    return true;
}
bool pwrite_all(int fd, const void* src, size_t n, uint64_t off)
{
    const char* p = static_cast<const char*>(src);
    while (n > 0) {
        const ssize_t r = ::pwrite(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r;
        off += (uint64_t)r;
        n -= (size_t)r;
    }
    return true;
}

// runs fn(part_offset, part_bytes) over n bytes on up to `ways` threads (page-cache and tmpfs copies are
// bound by one core's memcpy rate)
template <typename Fn>
bool split_parallel(size_t n, int ways, Fn fn)
{
    constexpr size_t MIN_PART = (size_t)16 << 20;
    if (n < 2 * MIN_PART || ways < 2) return fn((size_t)0, n);
    const int parts = (int)std::min<size_t>((size_t)ways, n / MIN_PART);
    const size_t per = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    std::atomic<bool> ok{true};
    for (int i = 0; i < parts; ++i) {
        const size_t lo = (size_t)i * per;
        if (lo >= n) break;
        const size_t len = std::min(per, n - lo);
        th.emplace_back([&, lo, len] { if (!fn(lo, len)) ok = false; });
    }
    for (auto& t : th) t.join();
    return ok;
}

constexpr int IO_WAYS = 4;

struct Source {
    virtual ~Source() {}
    virtual bool read(uint64_t off, void* dst, size_t n) = 0;
};
struct FdSource : Source {
    int fd;
    explicit FdSource(int f) : fd(f) {}
    bool read(uint64_t off, void* dst, size_t n) override
    {
        return split_parallel(n, IO_WAYS, [&](size_t lo, size_t len) { return pread_all(fd, (char*)dst + lo, len, off + lo); });
    }
};
struct MemSource : Source {
    const unsigned char* p;
    explicit MemSource(const void* q) : p(static_cast<const unsigned char*>(q)) {}
    bool read(uint64_t off, void* dst, size_t n) override
    {
        return split_parallel(n, IO_WAYS, [&](size_t lo, size_t len) { std::memcpy((char*)dst + lo, p + off + lo, len); return true; });
    }
};
struct Sink {
    virtual ~Sink() {}
    virtual bool write(uint64_t off, const void* src, size_t n) = 0;
};
struct FdSink : Sink {
    int fd;
    uint64_t base;
    FdSink(int f, uint64_t b) : fd(f), base(b) {}
    bool write(uint64_t off, const void* src, size_t n) override
    {
        return split_parallel(n, IO_WAYS, [&](size_t lo, size_t len) { return pwrite_all(fd, (const char*)src + lo, len, base + off + lo); });
    }
};
struct MemSink : Sink {  // p == nullptr: nothing is kept (size queries)
    unsigned char* p;
    size_t cap;
    std::atomic<bool> overflow{false};
    MemSink(void* q, size_t c) : p(static_cast<unsigned char*>(q)), cap(c) {}
    bool write(uint64_t off, const void* src, size_t n) override
    {
        if (!p) return true;
        if (off + n > cap) { overflow = true; return true; }
        return split_parallel(n, IO_WAYS, [&](size_t lo, size_t len) { std::memcpy(p + off + lo, (const char*)src + lo, len); return true; });
    }
};

// ------------------------------------------------------------------------------------------------ buffers

// Pinned host buffers are expensive to make (page locking runs at a few GB/s), so they are kept between
// calls and handed back at nnp_shutdown.
struct PinnedPool {
    std::mutex m;
    std::vector<std::pair<void*, size_t>> free_list;
    std::atomic<uint64_t> allocs{0}, alloc_bytes{0};
    double alloc_seconds = 0;  // diagnostics (nnp_internal_pool_stats)
    void* get(size_t bytes, size_t* cap)
    {
        {
            std::lock_guard<std::mutex> lock(m);
            size_t best = free_list.size();
            // best fit, but a small request leaves the large buffers to the large requests
            const size_t limit = 2 * bytes + ((size_t)64 << 20);
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].second >= bytes && free_list[i].second <= limit &&
                    (best == free_list.size() || free_list[i].second < free_list[best].second))
                    best = i;
            if (best != free_list.size()) {
                void* p = free_list[best].first;
                *cap = free_list[best].second;
                free_list.erase(free_list.begin() + (long)best);
                return p;
            }
        }
        const auto t0 = std::chrono::steady_clock::now();
        void* p = nnp_host_alloc(bytes + 64);
        *cap = p ? bytes : 0;
        allocs += 1;
        alloc_bytes += bytes;
        alloc_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return p;
    }
    void put(void* p, size_t cap)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lock(m);
        free_list.emplace_back(p, cap);
    }
    void release()
    {
        std::lock_guard<std::mutex> lock(m);
        for (auto& e : free_list) nnp_host_free(e.first);
        free_list.clear();
    }
};
PinnedPool g_pool;

// slabs of one run differ a little in size: buffers grow in generous steps so that they are made once
size_t roomy(size_t bytes)
{
    const size_t step = (size_t)32 << 20;
    return bytes < step ? bytes : ((bytes + bytes / 8 + step - 1) / step) * step;
}

struct PinnedBuffer {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuffer() { g_pool.put(p, cap); }
    bool reserve(size_t bytes)
    {
        if (bytes <= cap && p) return true;
        g_pool.put(p, cap);
        p = g_pool.get(roomy(bytes), &cap);
        return p != nullptr;
    }
};
struct DeviceBuffer {
    void* p = nullptr;
    size_t cap = 0;
    ~DeviceBuffer() { if (p) cudaFree(p); }
    bool reserve(size_t bytes)
    {
        if (bytes <= cap && p) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        bytes = roomy(bytes);
        if (cudaMalloc(&p, bytes + 64) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        cap = bytes;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ pipeline

// What the threads of all devices share: one lock, one condition variable (a few events per slab).
struct Shared {
    std::mutex m;
    std::condition_variable cv;
    int error = NNP_OK;  // first failure; every thread gives up when it is set
    void fail(int rc)
    {
        std::lock_guard<std::mutex> lock(m);
        if (error == NNP_OK) error = rc;
        cv.notify_all();
    }
    template <typename Pred>
    bool wait(Pred pred)  // false: another thread failed
    {
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [&] { return error != NNP_OK || pred(); });
        return error == NNP_OK;
    }
    template <typename Fn>
    void update(Fn fn)
    {
        {
            std::lock_guard<std::mutex> lock(m);
            fn();
        }
        cv.notify_all();
    }
};

// The double buffers of one device. State of a buffer: -1 = free, otherwise the slab it holds (set by the
// producer once the data is complete / the copy is enqueued, cleared by the consumer).
struct Lane {
    int device = 0;
    cudaStream_t up = nullptr, down = nullptr;
    cudaEvent_t up_done[2] = {nullptr, nullptr};
    PinnedBuffer h_in[2], h_out[2];
    DeviceBuffer d_in[2], d_out[2];
    long long in_slab[2] = {-1, -1};   // loader -> compute
    long long out_slab[2] = {-1, -1};  // compute -> drainer
    uint64_t out_bytes[2] = {0, 0}, out_off[2] = {0, 0};
    bool open()
    {
        if (cudaSetDevice(device) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (auto& e : up_done)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
        return true;
    }
    void close()
    {
        cudaSetDevice(device);
        if (up) cudaStreamDestroy(up);
        if (down) cudaStreamDestroy(down);
        for (auto& e : up_done)
            if (e) cudaEventDestroy(e);
        up = down = nullptr;
        up_done[0] = up_done[1] = nullptr;
    }
};

struct Slab {
    uint64_t read_off = 0, read_bytes = 0;  // what the loader brings to the device
    uint64_t lo = 0, hi = 0;                // compress: owned records [lo, hi); decompress: unused
    uint64_t g0 = 0;                        // compress: first record of the buffer (halo)
    // results
    uint64_t positions = 0;
    bool counted = false;  // decompress: `positions` is final
};

std::vector<int> bound_devices()
{
    std::vector<int> out;
    for (int i = 0;; ++i) {
        const int d = nnp_device_at(i);
        if (d < 0) break;
        out.push_back(d);
    }
    return out;
}

// loader of one lane: slabs lane, lane + G, lane + 2G, ... in order
void run_loader(Shared& S, Lane& L, const std::vector<Slab>& slabs, size_t first, size_t stride, Source& src)
{
    if (cudaSetDevice(L.device) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
    int j = 0;
    for (size_t k = first; k < slabs.size(); k += stride, ++j) {
        const int b = j & 1;
        if (!S.wait([&] { return L.in_slab[b] < 0; })) return;
        const Slab& s = slabs[k];
        if (s.read_bytes > 0) {
            if (cudaEventSynchronize(L.up_done[b]) != cudaSuccess) return S.fail(NNP_ERR_CUDA);  // h_in[b] is free again
            if (!L.h_in[b].reserve(s.read_bytes) || !L.d_in[b].reserve(s.read_bytes)) return S.fail(NNP_ERR_NOMEM);
            if (!src.read(s.read_off, L.h_in[b].p, s.read_bytes)) return S.fail(NNP_ERR_BAD_ARG);
            if (cudaMemcpyAsync(L.d_in[b].p, L.h_in[b].p, s.read_bytes, cudaMemcpyHostToDevice, L.up) != cudaSuccess)
                return S.fail(NNP_ERR_CUDA);
            if (cudaEventRecord(L.up_done[b], L.up) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
        }
        S.update([&] { L.in_slab[b] = (long long)k; });
    }
}

// drainer of one lane. `offset_of(k, &off)` blocks until the sink offset of slab k is known.
template <typename OffsetFn>
void run_drainer(Shared& S, Lane& L, size_t n_slabs, size_t first, size_t stride, Sink& sink, OffsetFn offset_of)
{
    if (cudaSetDevice(L.device) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
    int j = 0;
    for (size_t k = first; k < n_slabs; k += stride, ++j) {
        const int b = j & 1;
        if (!S.wait([&] { return L.out_slab[b] == (long long)k; })) return;
        const uint64_t bytes = L.out_bytes[b];
        if (bytes > 0) {
            uint64_t off = 0;
            if (!offset_of(k, b, &off)) return;
            if (!L.h_out[b].reserve(bytes)) return S.fail(NNP_ERR_NOMEM);
            if (cudaMemcpyAsync(L.h_out[b].p, L.d_out[b].p, bytes, cudaMemcpyDeviceToHost, L.down) != cudaSuccess ||
                cudaStreamSynchronize(L.down) != cudaSuccess)
                return S.fail(NNP_ERR_CUDA);
            if (!sink.write(off, L.h_out[b].p, bytes)) return S.fail(NNP_ERR_BAD_ARG);
        }
        S.update([&] { L.out_slab[b] = -1; });
    }
}

// ------------------------------------------------------------------------------------------------ .bin -> .binpack

struct CompressResult {
    uint64_t positions = 0, out_bytes = 0;
    int status = NNP_OK;
};

int compress_pipeline(Source& src, uint64_t n_total, Sink& sink, size_t slab_bytes, CompressResult* res)
{
    *res = CompressResult();
    if (n_total == 0) return NNP_OK;
    const std::vector<int> devices = bound_devices();
    if (devices.empty()) return NNP_ERR_NOT_INITIALISED;
    const size_t G = devices.size();

    uint64_t slab_records = (slab_bytes ? slab_bytes : ((size_t)256 << 20)) / 40;
    if (slab_records < 16) slab_records = 16;
    const uint64_t overlap0 = std::min<uint64_t>(65536, slab_records);
    std::vector<Slab> slabs;
    for (uint64_t lo = 0; lo < n_total; lo += slab_records) {
        Slab s;
        s.lo = lo;
        s.hi = std::min(lo + slab_records, n_total);
        s.g0 = lo > 0 ? lo - 1 : 0;
        const uint64_t g1 = std::min(s.hi + overlap0, n_total);
        s.read_off = s.g0 * 40;
        s.read_bytes = (g1 - s.g0) * 40;
        slabs.push_back(s);
    }

    Shared S;
    std::vector<Lane> lanes(G);
    for (size_t g = 0; g < G; ++g) {
        lanes[g].device = devices[g];
        if (!lanes[g].open()) return NNP_ERR_CUDA;
    }
    // the sequential state of the writer (compress_file.cpp:1061-1092), handed from slab to slab
    struct {
        size_t turn = 0;  // the slab whose orbit comes next
        uint64_t n_limit;
        uint64_t carry = NO_CARRY, chunks = 0, payload_base = 0;
        uint64_t pending_header = NONE, pending_start = 0;  // last chunk header written so far: sink offset, payload offset
        uint64_t positions = 0;
        int status = NNP_OK;
    } Q;
    Q.n_limit = n_total;

    auto patch_header = [&](uint64_t header_off, uint64_t size) {
        const uint32_t v = (uint32_t)size;
        const unsigned char le[4] = {(unsigned char)v, (unsigned char)(v >> 8), (unsigned char)(v >> 16), (unsigned char)(v >> 24)};
        return sink.write(header_off + 4, le, 4);
    };

    auto compute = [&](size_t g) {
        Lane& L = lanes[g];
        if (nnp_bind_device(L.device) != NNP_OK) return S.fail(NNP_ERR_CUDA);
        PinnedBuffer retry_h;
        DeviceBuffer retry_d;
        int j = 0;
        for (size_t k = g; k < slabs.size(); k += G, ++j) {
            const int b = j & 1, ob = j & 1;
            if (!S.wait([&] { return L.in_slab[b] == (long long)k; })) return;
            Slab& s = slabs[k];
            if (cudaEventSynchronize(L.up_done[b]) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
            // ---- the parallel part: chain walk, payload scan, payload write of this slab's chains
            nnp_shard_info info;
            std::memset(&info, 0, sizeof(info));
            const void* d_records = L.d_in[b].p;
            uint64_t overlap = overlap0;
            bool have_shard = false;
            for (;;) {
                uint64_t limit;
                {
                    std::lock_guard<std::mutex> lock(S.m);
                    limit = Q.n_limit;
                }
                if (limit <= s.lo) break;  // a malformed record in front of this slab: nothing of it is written
                const uint64_t hi = std::min(s.hi, limit);
                const uint64_t g1 = std::min(hi + overlap, limit);
                if (overlap != overlap0) {  // a chain longer than the window: read a wider one (rare)
                    const size_t bytes = (g1 - s.g0) * 40;
                    if (!retry_h.reserve(bytes) || !retry_d.reserve(bytes)) return S.fail(NNP_ERR_NOMEM);
                    if (!src.read(s.g0 * 40, retry_h.p, bytes)) return S.fail(NNP_ERR_BAD_ARG);
                    if (cudaMemcpy(retry_d.p, retry_h.p, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
                    d_records = retry_d.p;
                }
                const int rc = nnp_shard_compress_begin_dev(d_records, g1 - s.g0, s.lo - s.g0, hi - s.g0, g1 == limit, &info);
                if (rc == NNP_ERR_WINDOW) {
                    overlap *= 4;
                    continue;
                }
                if (rc == NNP_ERR_BAD_SFEN) {
                    // the reference stops at the first malformed record and its writer flushes what it has
                    // gathered (:407-408, :1094-1106): the output ends with the records in front of it
                    S.update([&] {
                        Q.n_limit = std::min(Q.n_limit, s.g0 + info.first_bad_record);
                        Q.status = NNP_ERR_BAD_SFEN;
                    });
                    continue;
                }
                if (rc != NNP_OK) return S.fail(rc);
                have_shard = true;
                break;
            }
            S.update([&] { L.in_slab[b] = -1; });  // the loader may refill this buffer
            // ---- the sequential part, in file order
            if (!S.wait([&] { return Q.turn == k; })) return;
            uint64_t limit, carry, chunks, payload_base;
            {
                std::lock_guard<std::mutex> lock(S.m);
                limit = Q.n_limit;
                carry = Q.carry;
                chunks = Q.chunks;
                payload_base = Q.payload_base;
            }
            uint64_t produced = 0, sink_off = 0;
            if (!S.wait([&] { return L.out_slab[ob] < 0; })) return;  // the drainer is done with this output buffer
            if (have_shard && limit > s.lo) {
                // (a malformed record a slab in front of this one found after this slab's begin leaves
                // limit <= s.lo: nothing of this slab is written; one inside this slab or its window was seen
                // by this slab's own begin; one further behind does not concern it)
                uint64_t n_starts = 0, first_start = NO_CARRY, carry_out = carry;
                int rc = nnp_shard_compress_orbit(payload_base, carry, &n_starts, &first_start, &carry_out);
                if (rc != NNP_OK) return S.fail(rc);
                // the chunk that was open when this slab began ends at the slab's first chunk start
                if (n_starts > 0 && Q.pending_header != NONE && !patch_header(Q.pending_header, first_start - Q.pending_start))
                    return S.fail(NNP_ERR_BAD_ARG);
                const uint64_t slab_end = payload_base + info.payload_bytes;
                size_t need = 0, got = 0;
                rc = nnp_shard_compress_emit_dev(slab_end, nullptr, 0, &need);  // placeholder: the chunk ends with the slab
                if (rc != NNP_OK) return S.fail(rc);
                if (need > 0) {
                    if (!L.d_out[ob].reserve(need)) return S.fail(NNP_ERR_NOMEM);
                    rc = nnp_shard_compress_emit_dev(slab_end, L.d_out[ob].p, need, &got);
                    if (rc != NNP_OK) return S.fail(rc);
                }
                produced = got;
                sink_off = payload_base + 8 * chunks;
                S.update([&] {
                    if (n_starts > 0) {
                        Q.pending_header = carry_out + 8 * (chunks + n_starts - 1);
                        Q.pending_start = carry_out;
                    }
                    Q.carry = carry_out;
                    Q.chunks = chunks + n_starts;
                    Q.payload_base = slab_end;
                    Q.positions += info.end_owned_record - info.first_owned_record;
                });
            }
            S.update([&] {
                Q.turn = k + 1;
                L.out_bytes[ob] = produced;
                L.out_off[ob] = sink_off;
                L.out_slab[ob] = (long long)k;
            });
        }
    };

    std::vector<std::thread> threads;
    for (size_t g = 0; g < G; ++g) {
        threads.emplace_back([&, g] { run_loader(S, lanes[g], slabs, g, G, src); });
        threads.emplace_back([&, g] { compute(g); });
        threads.emplace_back([&, g] {
            run_drainer(S, lanes[g], slabs.size(), g, G, sink, [&, g](size_t, int b, uint64_t* off) {
                *off = lanes[g].out_off[b];
                return true;
            });
        });
    }
    for (auto& t : threads) t.join();
    for (auto& L : lanes) L.close();
    if (S.error != NNP_OK) return S.error;
    // the last chunk ends with the payload
    if (Q.pending_header != NONE && !patch_header(Q.pending_header, Q.payload_base - Q.pending_start)) return NNP_ERR_BAD_ARG;
    res->positions = Q.positions;
    res->out_bytes = Q.payload_base + 8 * Q.chunks;
    res->status = Q.status;
    return NNP_OK;
}

// ------------------------------------------------------------------------------------------------ .binpack -> .bin

struct DecompressResult {
    uint64_t positions = 0;  // positions written
    int status = NNP_OK;     // header error that ended the walk
};

// the reference hands its output buffer to the file only once it exceeds 1 MiB (:1395-1402) and an
// exception thrown while fetching a chunk leaves Reader::next() before the last entry of the chunk in
// front is returned: this many records reach the file in that case
uint64_t committed_records(uint64_t written)
{
    if (written == 0) return 0;
    const uint64_t per_flush = (1048576 / 40) + 1;
    return (written - 1) / per_flush * per_flush;
}

int decompress_pipeline(Source& src, uint64_t total, Sink& sink, size_t slab_bytes, DecompressResult* res)
{
    *res = DecompressResult();
    if (total == 0) return NNP_OK;
    const std::vector<int> devices = bound_devices();
    if (devices.empty()) return NNP_ERR_NOT_INITIALISED;
    const size_t G = devices.size();
    const uint64_t slab = slab_bytes ? slab_bytes : ((size_t)32 << 20);

    // groups of whole chunks (:500-521 for the header checks)
    std::vector<Slab> slabs;
    int walk_status = NNP_OK;
    {
        uint64_t pos = 0, start = 0;
        while (pos < total) {
            unsigned char hdr[8];
            if (total - pos < 8 || !src.read(pos, hdr, 8) || std::memcmp(hdr, "BINP", 4) != 0) { walk_status = NNP_ERR_BAD_MAGIC; break; }
            const uint64_t size = (uint64_t)hdr[4] | ((uint64_t)hdr[5] << 8) | ((uint64_t)hdr[6] << 16) | ((uint64_t)hdr[7] << 24);
            if (size > 100u * (1u << 20)) { walk_status = NNP_ERR_CHUNK_TOO_LARGE; break; }
            if (total - pos - 8 < size) { walk_status = NNP_ERR_TRUNCATED; break; }
            if (pos > start && pos + 8 + size - start > slab) {
                Slab s;
                s.read_off = start;
                s.read_bytes = pos - start;
                slabs.push_back(s);
                start = pos;
            }
            pos += 8 + size;
        }
        if (pos > start) {
            Slab s;
            s.read_off = start;
            s.read_bytes = pos - start;
            slabs.push_back(s);
        }
    }

    Shared S;
    std::vector<Lane> lanes(G);
    for (size_t g = 0; g < G; ++g) {
        lanes[g].device = devices[g];
        if (!lanes[g].open()) return NNP_ERR_CUDA;
    }

    auto compute = [&](size_t g) {
        Lane& L = lanes[g];
        if (nnp_bind_device(L.device) != NNP_OK) return S.fail(NNP_ERR_CUDA);
        int j = 0;
        for (size_t k = g; k < slabs.size(); k += G, ++j) {
            const int b = j & 1;
            if (!S.wait([&] { return L.in_slab[b] == (long long)k; })) return;
            if (cudaEventSynchronize(L.up_done[b]) != cudaSuccess) return S.fail(NNP_ERR_CUDA);
            if (!S.wait([&] { return L.out_slab[b] < 0; })) return;
            Slab& s = slabs[k];
            size_t cap = L.d_out[b].cap ? L.d_out[b].cap : s.read_bytes * 24 + 4096, got = 0;
            int rc;
            for (;;) {
                if (!L.d_out[b].reserve(cap)) return S.fail(NNP_ERR_NOMEM);
                rc = nnp_binpack_to_bin_dev(L.d_in[b].p, s.read_bytes, L.d_out[b].p, L.d_out[b].cap, &got);
                if (rc != NNP_ERR_CAPACITY) break;
                cap = got + 4096;
            }
            if (rc != NNP_OK) return S.fail(rc);
            S.update([&] {
                s.positions = got / 40;
                s.counted = true;
                L.in_slab[b] = -1;
                L.out_bytes[b] = got;
                L.out_slab[b] = (long long)k;
            });
        }
    };
    // a slab's records land behind those of all slabs in front of it
    auto offset_of = [&](size_t k, int, uint64_t* off) {
        uint64_t before = 0;
        const bool ok = S.wait([&] {
            before = 0;
            for (size_t i = 0; i < k; ++i) {
                if (!slabs[i].counted) return false;
                before += slabs[i].positions;
            }
            return true;
        });
        *off = before * 40;
        return ok;
    };

    std::vector<std::thread> threads;
    for (size_t g = 0; g < G; ++g) {
        threads.emplace_back([&, g] { run_loader(S, lanes[g], slabs, g, G, src); });
        threads.emplace_back([&, g] { compute(g); });
        threads.emplace_back([&, g] { run_drainer(S, lanes[g], slabs.size(), g, G, sink, offset_of); });
    }
    for (auto& t : threads) t.join();
    for (auto& L : lanes) L.close();
    if (S.error != NNP_OK) return S.error;
    for (const Slab& s : slabs) res->positions += s.positions;
    res->status = walk_status;
    return NNP_OK;
}

// there is no CPU path: refuse before touching any file unless the library is bound to a device
bool library_ready() { return nnp_device_at(0) >= 0; }

struct Fd {
    int fd = -1;
    ~Fd() { if (fd >= 0) ::close(fd); }
};

}  // namespace

extern "C" {

void nnp_internal_release_buffers(void) { g_pool.release(); }
void nnp_internal_pool_stats(uint64_t* allocs, uint64_t* bytes, double* seconds)
{
    *allocs = g_pool.allocs;
    *bytes = g_pool.alloc_bytes;
    *seconds = g_pool.alloc_seconds;
}

int nnp_bin_to_binpack_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions)
{
    if (!in_path || !out_path) return NNP_ERR_BAD_ARG;
    if (positions) *positions = 0;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    Fd in, out;
    in.fd = ::open(in_path, O_RDONLY);
    if (in.fd < 0) return NNP_ERR_BAD_ARG;
    struct stat st;
    if (::fstat(in.fd, &st) != 0) return NNP_ERR_BAD_ARG;
    const uint64_t n_total = (uint64_t)st.st_size / 40;  // a short trailing record is dropped (:1360)
    out.fd = ::open(out_path, O_RDWR | O_CREAT | (append ? 0 : O_TRUNC), 0644);
    if (out.fd < 0) return NNP_ERR_BAD_ARG;
    uint64_t out_base = 0;
    if (append) {
        if (::fstat(out.fd, &st) != 0) return NNP_ERR_BAD_ARG;
        out_base = (uint64_t)st.st_size;
    }
    FdSource src(in.fd);
    FdSink sink(out.fd, out_base);
    CompressResult r;
    const int rc = compress_pipeline(src, n_total, sink, slab_bytes, &r);
    if (rc != NNP_OK) return rc;
    if (positions) *positions = r.positions;
    return r.status;
}

int nnp_binpack_to_bin_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions)
{
    if (!in_path || !out_path) return NNP_ERR_BAD_ARG;
    if (positions) *positions = 0;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    Fd in, out;
    in.fd = ::open(in_path, O_RDONLY);
    if (in.fd < 0) return NNP_ERR_BAD_ARG;
    struct stat st;
    if (::fstat(in.fd, &st) != 0) return NNP_ERR_BAD_ARG;
    const uint64_t total = (uint64_t)st.st_size;
    out.fd = ::open(out_path, O_RDWR | O_CREAT | (append ? 0 : O_TRUNC), 0644);
    if (out.fd < 0) return NNP_ERR_BAD_ARG;
    uint64_t out_base = 0;
    if (append) {
        if (::fstat(out.fd, &st) != 0) return NNP_ERR_BAD_ARG;
        out_base = (uint64_t)st.st_size;
    }
    FdSource src(in.fd);
    FdSink sink(out.fd, out_base);
    DecompressResult r;
    const int rc = decompress_pipeline(src, total, sink, slab_bytes, &r);
    if (rc != NNP_OK) return rc;
    uint64_t written = r.positions;
    if (r.status == NNP_ERR_BAD_MAGIC || r.status == NNP_ERR_CHUNK_TOO_LARGE) {
        const uint64_t committed = committed_records(written);
        if (committed < written && ::ftruncate(out.fd, (off_t)(out_base + committed * 40)) != 0) return NNP_ERR_BAD_ARG;
        written = committed;
    }
    if (positions) *positions = written;
    return r.status;
}

int nnp_bin_to_binpack_multi(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    if (!out_bytes || (bin_bytes >= 40 && !bin)) return NNP_ERR_BAD_ARG;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    const uint64_t n = bin_bytes / 40;
    if (!out) {  // every record costs at most a stem + numPlies (34 bytes) and a chunk header every MiB
        const size_t payload = n * 34 + 16;
        *out_bytes = payload + 8 * (payload / (1u << 20) + 2);
        return NNP_OK;
    }
    *out_bytes = 0;
    MemSource src(bin);
    MemSink sink(out, out_cap);
    CompressResult r;
    const int rc = compress_pipeline(src, n, sink, 0, &r);
    if (rc != NNP_OK) return rc;
    *out_bytes = r.out_bytes;
    if (sink.overflow) return NNP_ERR_CAPACITY;
    return r.status;
}

int nnp_binpack_to_bin_multi(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    if (!out_bytes || (binpack_bytes && !binpack)) return NNP_ERR_BAD_ARG;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    *out_bytes = 0;
    MemSource src(binpack);
    MemSink sink(out, out_cap);  // out == NULL: count only
    DecompressResult r;
    const int rc = decompress_pipeline(src, binpack_bytes, sink, 0, &r);
    if (rc != NNP_OK) return rc;
    *out_bytes = r.positions * 40;
    if (!out) return NNP_OK;  // capacity for everything decodable; the status is reported by the real call
    if (sink.overflow) return NNP_ERR_CAPACITY;
    if (r.status == NNP_ERR_BAD_MAGIC || r.status == NNP_ERR_CHUNK_TOO_LARGE) *out_bytes = committed_records(r.positions) * 40;
    return r.status;
}

}  // extern "C"
