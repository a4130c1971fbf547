// files.cu -- file-to-file conversion of inputs of any size: the two headline
// directions in slabs that fit the device, on top of the public C ABI of this library.
//
//   .bin -> .binpack   the slabs are the "ranks" of the sharded compressor (include/nnuepack.h), visited
//                      in file order by one GPU: every slab is read with one halo record in front and an
//                      overlap window behind, chains belong to the slab that holds their head, and the
//                      chunk-flush carry simply travels from slab to slab. The size field of a slab's
//                      last chunk is only known once a later slab opens the next chunk: it is written as
//                      a placeholder and patched in the file. The result is the file ONE reference run
//                      writes (compressBin, compress_file.cpp:1338-1374), whatever the slab size.
//   .binpack -> .bin   chunks are independent (decompressBin :1376-1412): the chunk headers are walked in
//                      the file, whole chunks are grouped into slabs and decoded one slab at a time.
//
// Host code only; every conversion runs in the CUDA kernels behind the *_dev entry points.
#include <cstdio>
#include <cstring>
#include <unistd.h>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/nnuepack.h"

namespace {

struct File {
    FILE* f = nullptr;
    ~File() { if (f) std::fclose(f); }
};

struct DeviceBuffer {
    void* p = nullptr;
    size_t cap = 0;
    ~DeviceBuffer() { if (p) cudaFree(p); }
    bool reserve(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, bytes + 64) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        cap = bytes;
        return true;
    }
};

struct PinnedBuffer {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuffer() { if (p) nnp_host_free(p); }
    bool reserve(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (p) nnp_host_free(p);
        p = nnp_host_alloc(bytes + 64);
        cap = p ? bytes : 0;
        return p != nullptr;
    }
};

bool read_at(FILE* f, uint64_t off, void* dst, size_t n)
{
    if (fseeko(f, (off_t)off, SEEK_SET) != 0) return false;
    return std::fread(dst, 1, n, f) == n;
}
bool write_at(FILE* f, uint64_t off, const void* src, size_t n)
{
    if (fseeko(f, (off_t)off, SEEK_SET) != 0) return false;
    return std::fwrite(src, 1, n, f) == n;
}

uint64_t file_size(FILE* f)
{
    fseeko(f, 0, SEEK_END);
    return (uint64_t)ftello(f);
}

constexpr uint64_t NO_CARRY = ~(uint64_t)0;

// there is no CPU path: refuse before touching any file unless the library is bound to a device
bool library_ready()
{
    size_t bound = 0;
    return nnp_bin_to_binpack_dev(nullptr, 0, nullptr, 0, &bound) == NNP_OK;
}

}  // namespace

extern "C" {

int nnp_bin_to_binpack_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions)
{
    if (!in_path || !out_path) return NNP_ERR_BAD_ARG;
    if (positions) *positions = 0;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    File in, out;
    in.f = std::fopen(in_path, "rb");
    if (!in.f) return NNP_ERR_BAD_ARG;
    const uint64_t n_total = file_size(in.f) / 40;  // a short trailing record is dropped (:1360)
    out.f = std::fopen(out_path, append ? "r+b" : "w+b");
    if (!out.f && append) out.f = std::fopen(out_path, "w+b");
    if (!out.f) return NNP_ERR_BAD_ARG;
    const uint64_t out_base = append ? file_size(out.f) : 0;
    if (n_total == 0) return NNP_OK;

    uint64_t slab_records = (slab_bytes ? slab_bytes : ((size_t)2 << 30)) / 40;
    if (slab_records < 16) slab_records = 16;
    DeviceBuffer d_in, d_out;
    PinnedBuffer h_in, h_out;

    uint64_t carry = NO_CARRY, chunks = 0, payload_base = 0;
    uint64_t pending_header = ~(uint64_t)0, pending_start = 0;  // last chunk header written so far: file offset, payload offset
    int status = NNP_OK;
    uint64_t n_limit = n_total;  // shrinks to the first malformed record
    for (uint64_t lo = 0; lo < n_limit;) {
        uint64_t hi = lo + slab_records < n_limit ? lo + slab_records : n_limit;
        uint64_t overlap = 65536 < slab_records ? 65536 : slab_records;
        nnp_shard_info info;
        uint64_t g0 = 0;
        for (;;) {  // widen the overlap window until the chain crossing `hi` ends inside it
            g0 = lo > 0 ? lo - 1 : 0;
            const uint64_t g1 = hi + overlap < n_limit ? hi + overlap : n_limit;
            const uint64_t n = g1 - g0;
            if (!h_in.reserve(n * 40) || !d_in.reserve(n * 40)) return NNP_ERR_NOMEM;
            if (!read_at(in.f, g0 * 40, h_in.p, n * 40)) return NNP_ERR_BAD_ARG;
            if (cudaMemcpy(d_in.p, h_in.p, n * 40, cudaMemcpyHostToDevice) != cudaSuccess) return NNP_ERR_CUDA;
            const int rc = nnp_shard_compress_begin_dev(d_in.p, n, lo - g0, hi - g0, g1 == n_limit, &info);
            if (rc == NNP_ERR_WINDOW) {
                overlap *= 4;
                continue;
            }
            if (rc == NNP_ERR_BAD_SFEN) {
                // the reference stops at the first malformed record and its writer flushes what it has
                // gathered (:407-408, :1094-1106): the file ends with the records in front of it
                status = NNP_ERR_BAD_SFEN;
                n_limit = g0 + info.first_bad_record;
                if (n_limit <= lo) break;
                if (hi > n_limit) hi = n_limit;
                continue;
            }
            if (rc != NNP_OK) return rc;
            break;
        }
        if (n_limit <= lo) break;
        uint64_t n_starts = 0, first_start = NO_CARRY, carry_out = carry;
        int rc = nnp_shard_compress_orbit(payload_base, carry, &n_starts, &first_start, &carry_out);
        if (rc != NNP_OK) return rc;
        // the chunk that was open when this slab began ends at the slab's first chunk start
        if (n_starts > 0 && pending_header != ~(uint64_t)0) {
            const uint32_t size = (uint32_t)(first_start - pending_start);
            unsigned char le[4] = {(unsigned char)size, (unsigned char)(size >> 8), (unsigned char)(size >> 16),
                                   (unsigned char)(size >> 24)};
            if (!write_at(out.f, pending_header + 4, le, 4)) return NNP_ERR_BAD_ARG;
        }
        const uint64_t slab_end = payload_base + info.payload_bytes;
        size_t need = 0, got = 0;
        rc = nnp_shard_compress_emit_dev(slab_end, nullptr, 0, &need);  // placeholder: the chunk ends with the slab
        if (rc != NNP_OK) return rc;
        if (need > 0) {
            if (!d_out.reserve(need) || !h_out.reserve(need)) return NNP_ERR_NOMEM;
            rc = nnp_shard_compress_emit_dev(slab_end, d_out.p, need, &got);
            if (rc != NNP_OK) return rc;
            if (cudaMemcpy(h_out.p, d_out.p, got, cudaMemcpyDeviceToHost) != cudaSuccess) return NNP_ERR_CUDA;
            if (!write_at(out.f, out_base + payload_base + 8 * chunks, h_out.p, got)) return NNP_ERR_BAD_ARG;
        }
        if (n_starts > 0) {
            pending_header = out_base + carry_out + 8 * (chunks + n_starts - 1);
            pending_start = carry_out;
        }
        carry = carry_out;
        chunks += n_starts;
        payload_base = slab_end;
        if (positions) *positions += info.end_owned_record - info.first_owned_record;
        lo = hi;
    }
    // the last chunk ends with the payload
    if (pending_header != ~(uint64_t)0) {
        const uint32_t size = (uint32_t)(payload_base - pending_start);
        unsigned char le[4] = {(unsigned char)size, (unsigned char)(size >> 8), (unsigned char)(size >> 16),
                               (unsigned char)(size >> 24)};
        if (!write_at(out.f, pending_header + 4, le, 4)) return NNP_ERR_BAD_ARG;
    }
    return status;
}

int nnp_binpack_to_bin_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions)
{
    if (!in_path || !out_path) return NNP_ERR_BAD_ARG;
    if (positions) *positions = 0;
    if (!library_ready()) return NNP_ERR_NOT_INITIALISED;
    File in, out;
    in.f = std::fopen(in_path, "rb");
    if (!in.f) return NNP_ERR_BAD_ARG;
    const uint64_t total = file_size(in.f);
    out.f = std::fopen(out_path, append ? "ab" : "wb");
    if (!out.f) return NNP_ERR_BAD_ARG;
    const uint64_t slab = slab_bytes ? slab_bytes : ((size_t)256 << 20);
    DeviceBuffer d_in, d_out;
    PinnedBuffer h_in, h_out;
    uint64_t pos = 0, written_positions = 0;
    int walk_status = NNP_OK;
    while (pos < total && walk_status == NNP_OK) {
        // a group of whole chunks (:500-521 for the header checks)
        uint64_t end = pos;
        while (end < total) {
            unsigned char hdr[8];
            if (total - end < 8 || !read_at(in.f, end, hdr, 8) || std::memcmp(hdr, "BINP", 4) != 0) {
                walk_status = NNP_ERR_BAD_MAGIC;
                break;
            }
            const uint64_t size = (uint64_t)hdr[4] | ((uint64_t)hdr[5] << 8) | ((uint64_t)hdr[6] << 16) | ((uint64_t)hdr[7] << 24);
            if (size > 100u * (1u << 20)) { walk_status = NNP_ERR_CHUNK_TOO_LARGE; break; }
            if (total - end - 8 < size) { walk_status = NNP_ERR_TRUNCATED; break; }
            if (end > pos && end + 8 + size - pos > slab) break;
            end += 8 + size;
        }
        const uint64_t n = end - pos;
        if (n == 0) break;
        if (!h_in.reserve(n) || !d_in.reserve(n)) return NNP_ERR_NOMEM;
        if (!read_at(in.f, pos, h_in.p, n)) return NNP_ERR_BAD_ARG;
        if (cudaMemcpy(d_in.p, h_in.p, n, cudaMemcpyHostToDevice) != cudaSuccess) return NNP_ERR_CUDA;
        size_t cap = d_out.cap ? d_out.cap : n * 24 + 4096, got = 0;
        int rc;
        for (;;) {
            if (!d_out.reserve(cap)) return NNP_ERR_NOMEM;
            rc = nnp_binpack_to_bin_dev(d_in.p, n, d_out.p, d_out.cap, &got);
            if (rc != NNP_ERR_CAPACITY) break;
            cap = got + 4096;
        }
        if (rc != NNP_OK) return rc;
        if (!h_out.reserve(got)) return NNP_ERR_NOMEM;
        if (cudaMemcpy(h_out.p, d_out.p, got, cudaMemcpyDeviceToHost) != cudaSuccess) return NNP_ERR_CUDA;
        if (std::fwrite(h_out.p, 1, got, out.f) != got) return NNP_ERR_BAD_ARG;
        written_positions += got / 40;
        pos = end;
    }
    if (walk_status == NNP_ERR_BAD_MAGIC || walk_status == NNP_ERR_CHUNK_TOO_LARGE) {
        // the reference hands its output buffer to the file only once it exceeds 1 MiB (:1395-1402) and
        // the exception leaves Reader::next() before the last entry of the chunk in front is returned
        uint64_t committed = 0;
        if (written_positions > 0) {
            const uint64_t per_flush = (1048576 / 40) + 1;
            committed = (written_positions - 1) / per_flush * per_flush;
        }
        std::fflush(out.f);
        if (committed < written_positions) {
            const uint64_t base = append ? (uint64_t)ftello(out.f) - written_positions * 40 : 0;
            if (ftruncate(fileno(out.f), (off_t)(base + committed * 40)) != 0) return NNP_ERR_BAD_ARG;
        }
        written_positions = committed;
    }
    if (positions) *positions = written_positions;
    return walk_status;
}

}  // extern "C"
