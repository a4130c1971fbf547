// api.cu -- the C ABI of libnnuepack.so (include/nnuepack.h): context, workspaces and the
// host-side sequencing of the kernels. No conversion arithmetic happens on the host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nnuepack.h"
#include "common.cuh"
#include "kernels.h"

namespace nnp {

namespace {

struct Workspace {
    void* p = nullptr;
    size_t cap = 0;
};

enum WsSlot {
    WS_CODES = 0, WS_STEMS, WS_TILE_AGG, WS_PAYLOAD, WS_HEAD_OFF, WS_CHUNK_OFF, WS_TOTALS,
    WS_CHUNK_START, WS_CHUNK_LEN, WS_CHUNK_TILE_BASE, WS_CHUNK_INFO, WS_TILE_COUNT, WS_TILE_PREFIX,
    WS_CAND_CHUNK, WS_CAND_OFF, WS_CAND_NEXT, WS_CAND_BASE, WS_CAND_CNT, WS_CHUNK_COUNT, WS_CHUNK_SLOW, WS_CHUNK_BASE,
    WS_DTOTALS, WS_AGG_TOP, WS_HEAD_NEXT, WS_PARK_A, WS_PARK_B, WS_TILE_FLAGS, WS_CAND_REC, WS_LSUM_A, WS_LSUM_B, WS_GAME_LEN, WS_GAME_BASE, WS_STAGE_IN, WS_STAGE_OUT, WS_TEXT_A, WS_TEXT_B, WS_TEXT_C, WS_TEXT_D, WS_BLEED, WS_WALK, WS_SCAN_TILES, WS_SCAN_PREFIX, WS_COUNT
};

// codes/stems of n records -> payload scan and payload write; leaves the headerless payload stream
// and the head offsets in the workspaces
struct PayloadPlan {
    u64 payload_bytes = 0, heads = 0, max_chunks = 0;
    bool dense = false;  // more than half of the records start a chain (picks the writer / orbit forms made for that)
    u32* payload = nullptr;
    u64* head_off = nullptr;
    u64* seg_off = nullptr;
    u32* head_next = nullptr;
    CompressTotals* d_tot = nullptr;
};

struct ShardState {
    bool active = false;
    PayloadPlan plan;
    u64 base = 0, chunks = 0;
    bool orbit_done = false;
};

struct Context {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;      // stream all kernels and copies are issued on
    cudaStream_t own_stream = nullptr;  // the library's default stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [4]: start of a sharded decode
    Workspace ws[WS_COUNT];
    void* pinned = nullptr;  // small pinned scratch for read-backs
    uint64_t launches = 0;
    std::string last_cuda_error;
    float last_total_ms = 0.f, last_dominant_ms = 0.f, last_stage_ms = 0.f;
    const char* last_kernel = "";  // the kernel last_dominant_ms belongs to
    uint64_t last_positions = 0;   // positions converted by the last driver call
    // host-buffer entry points: the copies are pipelined with the dominant kernel of the direction
    const void* pipe_src = nullptr;  // compress: host source of the records (d_bin is the staging buffer)
    void* pipe_dst = nullptr;        // decompress: host destination of the records
    size_t pipe_dst_cap = 0;
    bool pipe_dst_done = false;      // the pipelined D2H delivered the output
    cudaEvent_t pipe_ev[16] = {};
    uint32_t debug_reject_mod = 0;
    bool debug_k1_per_record = false;  // "k1_per_record": always the record-parallel K1
    bool debug_k1_heads = false;       // "k1_heads": always the chain-head transcoder (k_heads_transcode)
    int debug_dec_direct = 0;          // "dec_direct": 2 = never take the candidate-free route for files of single positions
    int debug_k1_direct = 0;           // "k1_direct": 1 = always try the one-kernel route for files of chain heads, 2 = never
    const void* sampled_bin = nullptr;  // the last head-density sample (sample_head_density)
    u64 sampled_n = 0, sampled_heads = 0, sampled_count = 0;
    bool sampled_valid = false;
    bool debug_k1_walk = false;        // "k1_walk": always the chain-owning walk (no density sample)
    bool debug_k1_runs = false;        // "k1_runs": always the run-based walk with parked heads
    bool debug_exhaustive = false;     // NNP_DEBUG_EXHAUSTIVE: skip the optimistic decode strategy
    uint64_t optimistic_misses = 0;    // optimistic decodes that had to be redone exhaustively
    uint64_t optimistic_hits = 0;
    uint64_t last_candidates = 0, last_tentative_positions = 0, last_violations = 0, last_false = 0;
    uint64_t last_false_sample[8] = {0};
    ShardState shard;  // the shard the sharded compressor currently holds (begin .. emit)
    std::mutex mutex;  // one call at a time per device
};

// One context per device. A host thread works on the device it bound itself to (nnp_init /
// nnp_bind_device); threads that never did use the first device the process initialised, which is the
// whole story for the one-process-per-GPU model.
constexpr int MAX_DEVICES = 64;
Context g_ctxs[MAX_DEVICES];
Context* g_default = nullptr;
thread_local Context* t_ctx = nullptr;
std::mutex g_init_mutex;
inline Context& current_context() { return *(t_ctx ? t_ctx : g_default ? g_default : &g_ctxs[0]); }
#define g_ctx (current_context())

int cuda_fail(cudaError_t e, const char* what)
{
    g_ctx.last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? NNP_ERR_NOMEM : NNP_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)

int ws_get(WsSlot slot, size_t bytes, void** out)
{
    Workspace& w = g_ctx.ws[slot];
    if (bytes == 0) bytes = 16;
    if (w.cap < bytes) {
        if (w.p) {
            CK(cudaStreamSynchronize(g_ctx.stream));
            CK(cudaFree(w.p));
            w.p = nullptr;
            w.cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&w.p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&w.p, want);
        }
        if (e != cudaSuccess) {
            w.p = nullptr;
            return cuda_fail(e, "cudaMalloc(workspace)");
        }
        w.cap = want;
    }
    *out = w.p;
    return NNP_OK;
}
#define WS(slot, bytes, type, var)                                            \
    type* var = nullptr;                                                      \
    do {                                                                      \
        void* p_ = nullptr;                                                   \
        int rc_ = ws_get(slot, (size_t)(bytes), &p_);                         \
        if (rc_ != NNP_OK) return rc_;                                        \
        var = reinterpret_cast<type*>(p_);                                    \
    } while (0)

int check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return NNP_OK;
}
#define LAUNCHED(n, what)                                    \
    do {                                                     \
        g_ctx.launches += (n);                               \
        int rc_ = check_launch(what);                        \
        if (rc_ != NNP_OK) return rc_;                       \
    } while (0)

size_t binpack_capacity_for_records(size_t n)
{
    // every record costs at most a stem + numPlies (34 bytes); a continuation ply at most 4
    const size_t payload = n * 34 + 16;
    return payload + 8 * (payload / CHUNK_THRESHOLD + 2);
}

// ---------------------------------------------------------------- shared tail of both compressors

// `rec_base`: index of codes[0] in the record numbering of the K1 pass (the bleed list uses it)
int build_payload(const u32* codes, const u32* stems, u64 n, PayloadPlan& P, u64 rec_base = 0)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    WS(WS_TILE_AGG, scan_tiles(n) * sizeof(Agg), Agg, tile_agg);
    WS(WS_TOTALS, sizeof(CompressTotals), CompressTotals, d_tot);
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    WS(WS_AGG_TOP, (scan_blocks(scan_tiles(n)) + 1) * sizeof(Agg), Agg, agg_top);
    launch_tile_aggregate(codes, n, tile_agg, s);
    launch_scan_aggregates(tile_agg, scan_tiles(n), agg_top, d_tot, s);
    LAUNCHED(4, "payload scan");
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    P.payload_bytes = h_tot->payload_bytes;
    P.heads = h_tot->heads;
    P.max_chunks = P.payload_bytes / CHUNK_THRESHOLD + 2;
    P.d_tot = d_tot;
    WS(WS_PAYLOAD, P.payload_bytes + 64, u32, payload);
    WS(WS_HEAD_OFF, (P.heads + 1) * 8, u64, head_off);
    WS(WS_CHUNK_OFF, (P.max_chunks + 3) * 8, u64, seg_off);
    WS(WS_HEAD_NEXT, (P.heads + 1) * 4, u32, head_next);
    P.head_next = head_next;
    P.payload = payload;
    P.head_off = head_off;
    P.seg_off = seg_off;
    CK(cudaMemsetAsync(payload, 0, P.payload_bytes + 64, s));
    P.dense = P.heads * 2 > n;
    launch_write_payload(codes, stems, n, tile_agg, payload, head_off, P.dense, s);
    launch_head_next(head_off, P.heads, head_next, d_tot, s);
    LAUNCHED(2, "k_write_payload");
    if (h_tot->bleeds > 0) {
        // stored moves that are not pseudo-legal: their ids overflow their fields and the reference ORs the
        // excess into the byte the field starts in (addBitsLE8 :840-862). Sort the few plies by record
        // (host) and let a second pass of the writer's scan add exactly those bits.
        const u64 count = h_tot->bleeds;
        if (count > BLEED_LIST_CAP) {
            C.last_cuda_error = "more than 2^20 stored moves are not pseudo-legal in their positions";
            return NNP_ERR_BAD_ARG;
        }
        WS(WS_BLEED, BLEED_LIST_CAP * 8, u64, bleed_list);
        std::vector<u64> host(count);
        CK(cudaMemcpyAsync(host.data(), bleed_list, count * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::sort(host.begin(), host.end(), [](u64 a, u64 b) { return (a & 0xFFFFFFFFull) < (b & 0xFFFFFFFFull); });
        CK(cudaMemcpyAsync(bleed_list, host.data(), count * 8, cudaMemcpyHostToDevice, s));
        launch_write_bleed(codes, n, tile_agg, payload, bleed_list, count, rec_base, s);
        LAUNCHED(1, "k_write_payload<bleed>");
        CK(cudaStreamSynchronize(s));  // `host` is pageable and goes out of scope
    }
    return NNP_OK;
}

// chunk orbit over the payload's heads; returns the number of chunks opened in it
int run_orbit(const PayloadPlan& P, u64 base, u64 carry, u64* chunks)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    launch_chunk_orbit(P.head_off, P.head_next, P.d_tot, P.seg_off, P.max_chunks, base, carry, P.dense, s);
    LAUNCHED(1, "k_chunk_orbit");
    CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *chunks = h_tot->chunks;
    return NNP_OK;
}

// codes/stems of n records -> payload scan, payload write, chunk orbit, chunk emission
int compress_tail(const u32* codes, const u32* stems, u64 n, int status, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    // (this call reuses the workspace slots an open nnp_shard_compress_* sequence points into: that sequence is
    // over, its later calls answer NNP_ERR_BAD_ARG instead of emitting from overwritten memory)
    C.shard.active = false;
    PayloadPlan P;
    int rc = build_payload(codes, stems, n, P);
    if (rc != NNP_OK) return rc;
    u64 chunks = 0;
    rc = run_orbit(P, 0, NO_CARRY, &chunks);
    if (rc != NNP_OK) return rc;
    const u64 total = P.payload_bytes + 8 * chunks;
    *out_bytes = total;
    if (total > out_cap) return NNP_ERR_CAPACITY;
    launch_emit_chunks(P.payload, P.seg_off, chunks, d_out, NATURAL_SIZE, C.stream);
    LAUNCHED(1, "k_emit_chunks");
    return status;
}

int reset_compress_totals(CompressTotals** d_tot_out)
{
    Context& C = g_ctx;
    WS(WS_TOTALS, sizeof(CompressTotals), CompressTotals, d_tot);
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    h_tot->payload_bytes = 0;
    h_tot->heads = 0;
    h_tot->chunks = 0;
    h_tot->error_index = NO_ERROR_IDX;
    h_tot->parked[0] = h_tot->parked[1] = 0;
    h_tot->bleeds = 0;
    CK(cudaMemcpyAsync(d_tot, h_tot, sizeof(CompressTotals), cudaMemcpyHostToDevice, C.stream));
    *d_tot_out = d_tot;
    return NNP_OK;
}

// ---------------------------------------------------------------- .bin -> .binpack

// The share of records whose ply / result fields rule out a continuation (isContinuation :589-590), on a sample
// of up to 65 536 record pairs of a device-resident file; one launch and one readback. `keep` holds the answer
// for the next question about the same file (compress_dev asks for the direct route, then K1 asks again).
int sample_head_density(const void* d_bin, u64 n_all, u64* heads, u64* samples, bool keep = false)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    if (C.sampled_bin == d_bin && C.sampled_n == n_all && C.sampled_valid) {
        *heads = C.sampled_heads;
        *samples = C.sampled_count;
        C.sampled_valid = false;  // one reuse: the buffer may hold another file next time
        return NNP_OK;
    }
    WS(WS_TOTALS, sizeof(CompressTotals), CompressTotals, d_tot);
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    *samples = n_all / 2 < 65536 ? n_all / 2 : 65536;
    const u64 stride = (n_all - 1) / *samples;
    CK(cudaMemsetAsync(&d_tot->parked[1], 0, 8, s));
    launch_sample_heads(d_bin, n_all, stride, *samples, &d_tot->parked[1], s);
    LAUNCHED(1, "k_sample_heads");
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *heads = h_tot->parked[1];
    CK(cudaMemsetAsync(&d_tot->parked[1], 0, 8, s));
    C.sampled_bin = d_bin;
    C.sampled_n = n_all;
    C.sampled_heads = *heads;
    C.sampled_count = *samples;
    C.sampled_valid = keep;
    return NNP_OK;
}

// A .bin of nothing but chain heads -> .binpack in one kernel (k_heads_direct, compress.cu). *done = false
// when the file turns out not to be of that kind, or does not fit: the caller then takes the general route,
// which overwrites whatever the attempt wrote.
int heads_direct_dev(const void* d_bin, u64 n, void* d_out, size_t out_cap, size_t* out_bytes, bool* done)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *done = false;
    C.shard.active = false;  // (WS_TOTALS is shared with an open nnp_shard_compress_* sequence, see compress_tail)
    const u64 total = heads_direct_bytes(n);
    if (total > out_cap) return NNP_OK;
    CompressTotals* d_tot = nullptr;
    int rc = reset_compress_totals(&d_tot);
    if (rc != NNP_OK) return rc;
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    CK(cudaEventRecord(C.ev[0], s));
    launch_heads_direct(d_bin, n, d_out, reinterpret_cast<u32*>(&d_tot->parked[0]), s);
    LAUNCHED(1, "k_heads_direct");
    CK(cudaEventRecord(C.ev[3], s));
    CK(cudaEventRecord(C.ev[1], s));
    CK(cudaEventRecord(C.ev[2], s));
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h_tot->parked[0] != 0) return NNP_OK;  // a possible continuation, a malformed or irregular stream
    *out_bytes = total;
    *done = true;
    C.last_kernel = "k_heads_direct";
    C.last_positions = n;
    CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
    C.last_dominant_ms = C.last_stage_ms = C.last_total_ms;
    return NNP_OK;
}

// K1: codes and stems of all n records (the chain walk, or the record-parallel kernel under the
// "k1_per_record" switch); *error_index = first malformed record or NO_ERROR_IDX
int link_encode_records(const void* d_bin, u64 n_all, u32** codes_out, u32** stems_out, u64* error_index)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    WS(WS_CODES, n_all * 4, u32, codes);
    WS(WS_STEMS, n_all * 32, u32, stems);
    WS(WS_BLEED, BLEED_LIST_CAP * 8, u64, bleed_list);
    *codes_out = codes;
    *stems_out = stems;
    CompressTotals* d_tot = nullptr;
    int rc = reset_compress_totals(&d_tot);
    if (rc != NNP_OK) return rc;
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    CK(cudaEventRecord(C.ev[0], s));
    // device-resident input: pick K1 from a sample of the chain-head density (host input is walked
    // piece by piece while it arrives, so there is nothing to sample yet)
    // three forms of K1: record-parallel for files of (nearly) single positions, chain-owning threads for
    // ordinary files, runs with parked heads when chains are enormous (or the input is still arriving)
    bool per_record = C.debug_k1_per_record;
    bool heads_only = C.debug_k1_heads;
    bool by_chains = !C.pipe_src && !C.debug_k1_runs && n_all < 0xFFFFFFFFull;
    if (!per_record && !heads_only && !C.debug_k1_walk && !C.debug_k1_runs && !C.pipe_src && n_all >= 4096) {
        u64 heads = 0, samples = 0;
        rc = sample_head_density(d_bin, n_all, &heads, &samples);
        if (rc != NNP_OK) return rc;
        per_record = heads * 3 > samples;       // more than one record in three starts a chain
        heads_only = heads * 10 > samples * 9;  // nearly all of them do: stems are transcoded (heads.cuh)
        // chains of hundreds of records: a thread owns too few of them to keep its warp busy (measured at
        // 400 plies per chain: 9.7 ms against 5.1 ms for the run-based walk), see DESIGN.md 4.1
        if (heads * 200 < samples) by_chains = false;
    }
    if (heads_only) per_record = true;
    if (!per_record && by_chains) {
        launch_walk_chains(d_bin, n_all, codes, stems, d_tot, bleed_list, s);
        LAUNCHED(1, "k_walk_chains");
        C.last_kernel = "k_walk_chains";
        CK(cudaEventRecord(C.ev[3], s));
        CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (h_tot->parked[1] == 0) {
            CK(cudaEventRecord(C.ev[1], s));
            *error_index = h_tot->error_index;
            return NNP_OK;
        }
        // a thread met a chain too long to walk alone: redo the file with the run-based walk
        CK(cudaMemsetAsync(&d_tot->parked[1], 0, 8, s));
        CK(cudaMemsetAsync(&d_tot->bleeds, 0, 8, s));
        CK(cudaEventRecord(C.ev[0], s));
    }
    if (per_record) {
        if (C.pipe_src) CK(cudaMemcpyAsync(const_cast<void*>(d_bin), C.pipe_src, n_all * 40, cudaMemcpyHostToDevice, s));
        if (heads_only) launch_heads_transcode(d_bin, n_all, codes, stems, d_tot, bleed_list, s);
        else launch_decode_link_encode(d_bin, n_all, codes, stems, d_tot, bleed_list, s);
        LAUNCHED(1, heads_only ? "k_heads_transcode" : "k_decode_link_encode");
        C.last_kernel = heads_only ? "k_heads_transcode" : "k_decode_link_encode";
        CK(cudaEventRecord(C.ev[3], s));
        CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    } else {
        if (n_all >= 0xFFFFFFFFull) return NNP_ERR_BAD_ARG;  // record indices travel as 32 bits
        C.last_kernel = "k_walk_runs";
        WS(WS_PARK_A, (walk_runs(n_all) + 1) * 4, u32, park_a);
        WS(WS_PARK_B, (walk_runs(n_all) + 1) * 4, u32, park_b);
        u32* lists[2] = {park_a, park_b};
        const u64 runs = walk_runs(n_all);
        if (C.pipe_src) {
            // host input: copy it in PIPE_PIECES pieces on the copy stream and walk every piece as soon
            // as it has landed (a piece is a whole number of runs; its halo record is in the piece before)
            constexpr u64 PIPE_PIECES = 16;
            const u64 runs_per = (runs + PIPE_PIECES - 1) / PIPE_PIECES;
            const u64 run_bytes = (u64)walk_run_records() * 40;
            CK(cudaEventRecord(C.pipe_ev[0], s));
            CK(cudaStreamWaitEvent(C.copy_stream, C.pipe_ev[0], 0));  // the staging buffer is free
            for (u64 k = 0; k < PIPE_PIECES; ++k) {
                const u64 lo = k * runs_per, hi = lo + runs_per < runs ? lo + runs_per : runs;
                if (lo >= hi) break;
                const u64 b0 = lo * run_bytes, b1 = hi * run_bytes < n_all * 40 ? hi * run_bytes : n_all * 40;
                CK(cudaMemcpyAsync((char*)const_cast<void*>(d_bin) + b0, (const char*)C.pipe_src + b0, b1 - b0,
                                   cudaMemcpyHostToDevice, C.copy_stream));
                CK(cudaEventRecord(C.pipe_ev[k % 16], C.copy_stream));
                CK(cudaStreamWaitEvent(s, C.pipe_ev[k % 16], 0));
                launch_walk_runs(d_bin, n_all, lo, hi, codes, stems, d_tot, lists[0], &d_tot->parked[0], bleed_list, s);
                LAUNCHED(1, "k_walk_runs");
            }
        } else {
            launch_walk_runs(d_bin, n_all, 0, runs, codes, stems, d_tot, lists[0], &d_tot->parked[0], bleed_list, s);
            LAUNCHED(1, "k_walk_runs");
        }
        CK(cudaEventRecord(C.ev[3], s));
        for (int cur = 0;; cur ^= 1) {
            CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            const u64 n_items = h_tot->parked[cur];
            if (n_items == 0) break;
            CK(cudaMemsetAsync(&d_tot->parked[cur ^ 1], 0, 8, s));
            launch_walk_items(d_bin, n_all, codes, stems, d_tot, lists[cur], n_items, lists[cur ^ 1], &d_tot->parked[cur ^ 1],
                              bleed_list, s);
            LAUNCHED(1, "k_walk_items");
        }
    }
    CK(cudaEventRecord(C.ev[1], s));
    *error_index = h_tot->error_index;
    return NNP_OK;
}

int compress_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    const u64 n_all = bin_bytes / 40;  // a short trailing record is dropped (compress_file.cpp:1360)
    if (!d_out) {
        *out_bytes = binpack_capacity_for_records(n_all);
        return NNP_OK;
    }
    *out_bytes = 0;
    if (n_all == 0) return NNP_OK;  // empty input -> empty file
    if (((uintptr_t)d_bin & 7) || ((uintptr_t)d_out & 7)) return NNP_ERR_BAD_ARG;

    cudaStream_t s = C.stream;
    int rc = NNP_OK;
    // a file of single positions (every sampled record fails the field test of isContinuation) is first tried
    // in one kernel; the attempt itself notices if a record might continue a chain after all
    bool direct = C.debug_k1_direct == 1;
    const bool pinned_k1 = C.debug_k1_per_record || C.debug_k1_heads || C.debug_k1_walk || C.debug_k1_runs;
    if (C.debug_k1_direct == 0 && !pinned_k1 && !C.pipe_src && n_all >= 4096) {
        u64 heads = 0, samples = 0;
        rc = sample_head_density(d_bin, n_all, &heads, &samples, true);
        if (rc != NNP_OK) return rc;
        direct = heads == samples;
    }
    if (direct && !C.pipe_src) {
        bool done = false;
        rc = heads_direct_dev(d_bin, n_all, d_out, out_cap, out_bytes, &done);
        if (rc != NNP_OK || done) {
            C.sampled_valid = false;
            return rc;
        }
    }
    u32 *codes = nullptr, *stems = nullptr;
    u64 error_index = NO_ERROR_IDX;
    rc = link_encode_records(d_bin, n_all, &codes, &stems, &error_index);
    if (rc != NNP_OK) return rc;
    int status = NNP_OK;
    u64 n = n_all;
    if (error_index != NO_ERROR_IDX) {
        // "Improperly encoded bin sfen": the reference stops at the first malformed record and its
        // writer still flushes everything gathered before it (compress_file.cpp:407-408, :1094-1106)
        status = NNP_ERR_BAD_SFEN;
        n = error_index;
        if (n == 0) return status;
    }
    rc = compress_tail(codes, stems, n, status, d_out, out_cap, out_bytes);
    if (rc != NNP_OK && rc != NNP_ERR_BAD_SFEN) return rc;
    C.last_positions = n;
    CK(cudaEventRecord(C.ev[2], s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
    CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[0], C.ev[3]));  // k_walk_runs alone
    CK(cudaEventElapsedTime(&C.last_stage_ms, C.ev[0], C.ev[1]));     // + the rounds of k_walk_items
    return rc;
}

// ---------------------------------------------------------------- sharded .bin -> .binpack
//
// One shard = the records a rank owns plus one halo record before them and an overlap window
// behind them. A chain belongs to the shard its head lies in, so the shard's payload is made of the
// chains whose heads fall into [own_lo, own_hi); the records in front of the first such head end a
// chain of the previous shard, which is why that shard looks at them through its overlap window.

int shard_begin_dev(const void* d_bin, u64 n_records, u64 own_lo, u64 own_hi, int reaches_eof, nnp_shard_info* info)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    g_ctx.shard = ShardState();
    std::memset(info, 0, sizeof(*info));
    info->first_bad_record = NO_ERROR_IDX;
    if (own_lo > own_hi || own_hi > n_records) return NNP_ERR_BAD_ARG;
    if ((uintptr_t)d_bin & 7) return NNP_ERR_BAD_ARG;
    if (n_records == 0) {
        g_ctx.shard.active = true;
        return NNP_OK;
    }
    u32 *codes = nullptr, *stems = nullptr;
    u64 error_index = NO_ERROR_IDX;
    int rc = link_encode_records(d_bin, n_records, &codes, &stems, &error_index);
    if (rc != NNP_OK) return rc;
    CK(cudaEventSynchronize(C.ev[1]));
    CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[0], C.ev[3]));  // k_walk_runs alone
    CK(cudaEventElapsedTime(&C.last_stage_ms, C.ev[0], C.ev[1]));
    C.last_total_ms = C.last_stage_ms;
    if (error_index != NO_ERROR_IDX) {
        info->first_bad_record = error_index;
        return NNP_ERR_BAD_SFEN;
    }
    // first owned head and first head of the next shard
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
    WS(WS_TEXT_D, 64, u64, d_found);
    launch_find_head(codes, n_records, own_lo, d_found, s);
    launch_find_head(codes, n_records, own_hi, d_found + 1, s);
    LAUNCHED(2, "k_find_head");
    CK(cudaMemcpyAsync(h_u64, d_found, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    u64 first = h_u64[0];
    const u64 end = h_u64[1];
    // no head behind own_hi: fine when the buffer ends with the file (the chain runs to the end),
    // otherwise the chain crossing own_hi does not end inside the window
    if (end >= n_records && !reaches_eof) return NNP_ERR_WINDOW;
    if (first > end) first = end;  // no head of its own: the whole range continues an earlier shard's chain
    info->first_owned_record = first;
    info->end_owned_record = end;
    g_ctx.shard.active = true;
    if (first == end) return NNP_OK;
    rc = build_payload(codes + first, stems + first * 8, end - first, g_ctx.shard.plan, first);
    if (rc != NNP_OK) return rc;
    info->payload_bytes = g_ctx.shard.plan.payload_bytes;
    info->chains = g_ctx.shard.plan.heads;
    return NNP_OK;
}

int shard_orbit(u64 payload_base, u64 carry_in, u64* n_starts, u64* first_start, u64* carry_out)
{
    if (!g_ctx.shard.active) return NNP_ERR_BAD_ARG;
    g_ctx.shard.base = payload_base;
    g_ctx.shard.chunks = 0;
    g_ctx.shard.orbit_done = true;
    *n_starts = 0;
    *first_start = NO_CARRY;
    *carry_out = carry_in;
    if (g_ctx.shard.plan.payload_bytes == 0) return NNP_OK;
    Context& C = g_ctx;
    int rc = run_orbit(g_ctx.shard.plan, payload_base, carry_in, &g_ctx.shard.chunks);
    if (rc != NNP_OK) return rc;
    *n_starts = g_ctx.shard.chunks;
    if (g_ctx.shard.chunks > 0) {
        u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
        CK(cudaMemcpyAsync(h_u64, g_ctx.shard.plan.seg_off + 1, 8, cudaMemcpyDeviceToHost, C.stream));
        CK(cudaMemcpyAsync(h_u64 + 1, g_ctx.shard.plan.seg_off + g_ctx.shard.chunks, 8, cudaMemcpyDeviceToHost, C.stream));
        CK(cudaStreamSynchronize(C.stream));
        *first_start = payload_base + h_u64[0];
        *carry_out = payload_base + h_u64[1];
    }
    return NNP_OK;
}

int shard_table_dev(void* d_table)
{
    if (!g_ctx.shard.active) return NNP_ERR_BAD_ARG;
    Context& C = g_ctx;
    if ((uintptr_t)d_table & 7) return NNP_ERR_BAD_ARG;
    if (g_ctx.shard.plan.payload_bytes == 0) {
        CK(cudaMemsetAsync(d_table, 0xFF, (size_t)NNP_ORBIT_TABLE_ENTRIES * 24, C.stream));
    } else {
        launch_orbit_table(g_ctx.shard.plan.head_off, g_ctx.shard.plan.head_next, g_ctx.shard.plan.d_tot, (u64*)d_table,
                           NNP_ORBIT_TABLE_ENTRIES, C.stream);
        LAUNCHED(1, "k_orbit_table");
    }
    CK(cudaStreamSynchronize(C.stream));
    return NNP_OK;
}

int shard_resolve_dev(const void* d_tables, const uint64_t* payload_bytes, int world, int rank, uint64_t* out5)
{
    Context& C = g_ctx;
    if (world <= 0 || rank < 0 || rank >= world) return NNP_ERR_BAD_ARG;
    WS(WS_TEXT_C, (size_t)world * 8 + 128, u64, d_sizes);
    if (world > 128) return NNP_ERR_BAD_ARG;  // the pinned scratch holds 128 sizes
    u64* h_sizes = reinterpret_cast<u64*>((char*)C.pinned + 2560);
    for (int r = 0; r < world; ++r) h_sizes[r] = payload_bytes[r];
    CK(cudaMemcpyAsync(d_sizes, h_sizes, (size_t)world * 8, cudaMemcpyHostToDevice, C.stream));
    launch_orbit_resolve((const u64*)d_tables, NNP_ORBIT_TABLE_ENTRIES, d_sizes, world, rank, d_sizes + world, C.stream);
    LAUNCHED(1, "k_orbit_resolve");
    u64* h_out = reinterpret_cast<u64*>((char*)C.pinned + 768);
    CK(cudaMemcpyAsync(h_out, d_sizes + world, 40, cudaMemcpyDeviceToHost, C.stream));
    CK(cudaStreamSynchronize(C.stream));
    for (int i = 0; i < 5; ++i) out5[i] = h_out[i];
    return NNP_OK;
}

int shard_emit_dev(u64 next_start, void* d_out, size_t out_cap, size_t* out_bytes)
{
    if (!g_ctx.shard.active || !g_ctx.shard.orbit_done) return NNP_ERR_BAD_ARG;
    Context& C = g_ctx;
    const PayloadPlan& P = g_ctx.shard.plan;
    const u64 total = P.payload_bytes + 8 * g_ctx.shard.chunks;
    *out_bytes = total;
    if (!d_out) return NNP_OK;
    if (total > out_cap) return NNP_ERR_CAPACITY;
    if ((uintptr_t)d_out & 7) return NNP_ERR_BAD_ARG;
    if (P.payload_bytes == 0) return NNP_OK;
    u64 last_size = NATURAL_SIZE;
    if (g_ctx.shard.chunks > 0) {
        u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
        CK(cudaMemcpyAsync(h_u64, P.seg_off + g_ctx.shard.chunks, 8, cudaMemcpyDeviceToHost, C.stream));
        CK(cudaStreamSynchronize(C.stream));
        const u64 last_start = g_ctx.shard.base + h_u64[0];
        if (next_start <= last_start) return NNP_ERR_BAD_ARG;
        last_size = next_start - last_start;
    }
    launch_emit_chunks(P.payload, P.seg_off, g_ctx.shard.chunks, d_out, last_size, C.stream);
    LAUNCHED(1, "k_emit_chunks");
    CK(cudaStreamSynchronize(C.stream));
    return NNP_OK;
}

// ---------------------------------------------------------------- .plain -> entries

struct ParsedText {
    Entry* entries = nullptr;
    u64 nrec = 0;
};

// record terminators -> per-record parse; NNP_ERR_BAD_TEXT for layouts the reference would
// read differently from their line structure (or crash on)
int parse_plain_dev(const void* d_text, size_t text_bytes, bool count_only, ParsedText& P)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    P.nrec = 0;
    if (text_bytes == 0) return NNP_OK;
    const u64 tiles = find_tiles(text_bytes);
    WS(WS_TILE_COUNT, (tiles + 1) * 4, u32, tile_count);
    WS(WS_TILE_PREFIX, (tiles + 2) * 8, u64, tile_prefix);
    WS(WS_DTOTALS, sizeof(PlainTotals) + 64, PlainTotals, d_tot);
    PlainTotals* h_tot = reinterpret_cast<PlainTotals*>((char*)C.pinned + 1280);
    h_tot->error_pos = NO_ERROR_IDX;
    h_tot->committed = 0;
    h_tot->inherit = 0;
    CK(cudaMemcpyAsync(d_tot, h_tot, sizeof(PlainTotals), cudaMemcpyHostToDevice, s));
    launch_find_records(false, d_text, text_bytes, tile_count, tile_prefix, nullptr, d_tot, s);
    launch_exclusive_sum(tile_count, tiles, tile_prefix, s);
    LAUNCHED(2, "k_find_records<count>");
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
    CK(cudaMemcpyAsync(h_u64, tile_prefix + tiles, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(PlainTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    P.nrec = h_u64[0];
    if (h_tot->error_pos != NO_ERROR_IDX) return NNP_ERR_BAD_TEXT;
    if (count_only || P.nrec == 0) return NNP_OK;
    WS(WS_TEXT_A, (P.nrec + 1) * 8, u64, rec_pos);
    WS(WS_TEXT_B, P.nrec * sizeof(Entry), Entry, entries);
    launch_find_records(true, d_text, text_bytes, tile_count, tile_prefix, rec_pos, d_tot, s);
    launch_parse_records(d_text, text_bytes, rec_pos, P.nrec, entries, d_tot, s);
    LAUNCHED(2, "k_parse_records");
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(PlainTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h_tot->error_pos != NO_ERROR_IDX) return NNP_ERR_BAD_TEXT;
    if (h_tot->inherit) {  // some record inherits a field from an earlier one (:1254): resolved in linear time
        WS(WS_TEXT_C, P.nrec * 40 + 64, u64, defs);
        WS(WS_TEXT_D, defs_tiles(P.nrec) * 40 + 64, u64, tile_max);
        launch_parse_inherited(d_text, text_bytes, rec_pos, P.nrec, defs, tile_max, entries, d_tot, s);
        LAUNCHED(5, "k_parse_inherited");
        CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(PlainTotals), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (h_tot->error_pos != NO_ERROR_IDX) return NNP_ERR_BAD_TEXT;
    }
    P.entries = entries;
    return NNP_OK;
}

int plain_to_binpack_dev(const void* d_text, size_t text_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *out_bytes = 0;
    ParsedText P;
    int rc = parse_plain_dev(d_text, text_bytes, d_out == nullptr, P);
    if (rc != NNP_OK) return rc;
    if (!d_out) {
        *out_bytes = binpack_capacity_for_records(P.nrec);
        return NNP_OK;
    }
    if (P.nrec == 0) return NNP_OK;
    if ((uintptr_t)d_out & 7) return NNP_ERR_BAD_ARG;
    WS(WS_CODES, P.nrec * 4, u32, codes);
    WS(WS_STEMS, P.nrec * 32, u32, stems);
    CompressTotals* d_tot = nullptr;
    rc = reset_compress_totals(&d_tot);
    if (rc != NNP_OK) return rc;
    WS(WS_BLEED, BLEED_LIST_CAP * 8, u64, bleed_list);
    launch_entries_link_encode(P.entries, P.nrec, codes, stems, d_tot, bleed_list, s);
    LAUNCHED(1, "k_entries_link_encode");
    rc = compress_tail(codes, stems, P.nrec, NNP_OK, d_out, out_cap, out_bytes);
    if (rc != NNP_OK) return rc;
    C.last_positions = P.nrec;
    CK(cudaStreamSynchronize(s));
    return NNP_OK;
}

int plain_to_bin_dev(const void* d_text, size_t text_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *out_bytes = 0;
    ParsedText P;
    int rc = parse_plain_dev(d_text, text_bytes, d_out == nullptr, P);
    if (rc != NNP_OK) return rc;
    *out_bytes = P.nrec * 40;
    if (!d_out || P.nrec == 0) return NNP_OK;
    if ((uintptr_t)d_out & 7) return NNP_ERR_BAD_ARG;
    if (P.nrec * 40 > out_cap) return NNP_ERR_CAPACITY;
    launch_entries_to_bin(P.entries, P.nrec, d_out, s);
    LAUNCHED(1, "k_entries_to_bin");
    C.last_positions = P.nrec;
    CK(cudaStreamSynchronize(s));
    return NNP_OK;
}

// ---------------------------------------------------------------- .bin -> .plain

int bin_to_plain_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *out_bytes = 0;
    u64 n = bin_bytes / 40;
    if (n == 0) return NNP_OK;
    if ((uintptr_t)d_bin & 7) return NNP_ERR_BAD_ARG;
    WS(WS_CODES, n * 4, u32, lens);
    WS(WS_TEXT_A, (n + 2) * 8, u64, offs);
    WS(WS_TILE_COUNT, (large_sum_tiles(n) + 1) * 4, u32, tile_sum);
    WS(WS_TILE_PREFIX, (large_sum_tiles(n) + 2) * 8, u64, tile_prefix);
    CompressTotals* d_tot = nullptr;
    int rc = reset_compress_totals(&d_tot);
    if (rc != NNP_OK) return rc;
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    launch_bin_text(false, d_bin, n, lens, nullptr, nullptr, d_tot, s);
    LAUNCHED(1, "k_bin_text<size>");
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    int status = NNP_OK;
    if (h_tot->error_index != NO_ERROR_IDX) {
        status = NNP_ERR_BAD_SFEN;  // convertBinToPlain stops at the first malformed record
        n = h_tot->error_index;
        if (n == 0) return status;
    }
    launch_exclusive_sum_large(lens, n, offs, tile_sum, tile_prefix, s);
    LAUNCHED(3, "text offsets");
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
    CK(cudaMemcpyAsync(h_u64, offs + n, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const u64 total = h_u64[0];
    *out_bytes = total;
    C.last_positions = n;
    if (!d_out) return status == NNP_OK ? NNP_OK : status;
    if (total > out_cap) return NNP_ERR_CAPACITY;
    launch_bin_text(true, d_bin, n, lens, offs, d_out, d_tot, s);
    LAUNCHED(1, "k_bin_text<write>");
    if (status != NNP_OK) {
        // only what the reference had already flushed survives its exception (:1448-1455)
        WS(WS_DTOTALS, sizeof(PlainTotals) + 64, PlainTotals, d_pt);
        launch_text_flush_orbit(d_out, total, 0, &d_pt->committed, s);
        LAUNCHED(1, "k_text_flush_orbit");
        CK(cudaMemcpyAsync(h_u64, &d_pt->committed, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        *out_bytes = h_u64[0];
        return status;
    }
    CK(cudaStreamSynchronize(s));
    return NNP_OK;
}

// ---------------------------------------------------------------- .binpack -> .bin / .plain

struct DecodePlan {
    ChunkTable tab;
    u64 chunks = 0, tiles = 0, ncand = 0, positions = 0, text_bytes = 0;
    int walk_status = 0;
    u32 *cand_chunk = nullptr, *cand_off = nullptr, *cand_next = nullptr, *cand_base = nullptr, *cand_cnt = nullptr;
    u64* cand_tbase = nullptr;
    u32 *chunk_count = nullptr, *chunk_slow = nullptr;
    u64* chunk_base = nullptr;
    u64* chunk_tbase = nullptr;
    u64* tile_prefix = nullptr;
    u32* tile_flags = nullptr;
    DecompressTotals* d_tot = nullptr;
    bool collapse = false;        // chunks of single positions are ONE entry of the candidate list (decompress_dev only)
    bool walked = false;          // decode_walk has run
    u32* chunk_flag = nullptr;    // decode_heads_only has run: 1 = the chunk holds nothing but single positions
    u64 heads_only_chunks = 0;
};

// chunk table: the header walk (:500-521) and the zeroed totals
int decode_walk(const void* d_in, size_t in_bytes, DecodePlan& P)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    WS(WS_CHUNK_INFO, sizeof(ChunkInfo), ChunkInfo, d_info);
    ChunkInfo* h_info = reinterpret_cast<ChunkInfo*>((char*)C.pinned + 256);
    u64 table_cap = in_bytes / 65536 + 1024;
    for (int attempt = 0; attempt < 2; ++attempt) {
        WS(WS_CHUNK_START, (table_cap + 1) * 8, u64, start);
        WS(WS_CHUNK_LEN, (table_cap + 1) * 4, u32, len);
        WS(WS_CHUNK_TILE_BASE, (table_cap + 2) * 8, u64, tile_base);
        P.tab.start = start;
        P.tab.len = len;
        P.tab.tile_base = tile_base;
        P.tab.info = d_info;
        WS(WS_WALK, walk_scratch_bytes(), unsigned char, walk_scratch);
        launch_walk_chunks(d_in, in_bytes, P.tab, table_cap, 1, 0, walk_scratch, s);
        LAUNCHED(1, "k_walk_chunks");
        CK(cudaMemcpyAsync(h_info, d_info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (h_info->chunks <= table_cap) break;
        table_cap = h_info->chunks + 16;
    }
    P.chunks = h_info->chunks;
    P.tiles = h_info->tiles;
    P.walk_status = h_info->status;
    WS(WS_DTOTALS, sizeof(DecompressTotals) + 64, DecompressTotals, d_tot);
    P.d_tot = d_tot;
    DecompressTotals* h_tot = reinterpret_cast<DecompressTotals*>((char*)C.pinned + 512);
    std::memset(h_tot, 0, sizeof(DecompressTotals));
    h_tot->error_chunk = NO_ERROR_IDX;
    CK(cudaMemcpyAsync(d_tot, h_tot, sizeof(DecompressTotals), cudaMemcpyHostToDevice, s));
    P.walked = true;
    return NNP_OK;
}

// which chunks hold nothing but single positions (their stem counts go to WS_CHUNK_COUNT)
int decode_heads_only(const void* d_in, DecodePlan& P)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    WS(WS_CHUNK_SLOW, (P.chunks + 1) * 4, u32, chunk_flag);  // (the exhaustive strategy reuses the slot afterwards)
    WS(WS_CHUNK_COUNT, (P.chunks + 1) * 4, u32, chunk_stems);
    launch_chunk_heads_only(d_in, P.tab, P.chunks, chunk_flag, chunk_stems, &P.d_tot->heads_only_chunks, s);
    LAUNCHED(1, "k_chunk_heads_only");
    P.chunk_flag = chunk_flag;
    P.chunk_count = chunk_stems;
    return NNP_OK;
}

// chunk table + candidate list: the part both decode strategies share
int decode_front(const void* d_in, size_t in_bytes, DecodePlan& P)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    int rc = P.walked ? NNP_OK : decode_walk(d_in, in_bytes, P);
    if (rc != NNP_OK) return rc;
    if (P.chunks == 0) return NNP_OK;
    if (!P.chunk_flag) {
        rc = decode_heads_only(d_in, P);
        if (rc != NNP_OK) return rc;
    }
    u32* chunk_flag = P.chunk_flag;

    WS(WS_TILE_COUNT, (P.tiles + 1) * 4, u32, tile_count);
    WS(WS_TILE_PREFIX, (P.tiles + 2) * 8, u64, tile_prefix);
    WS(WS_TILE_FLAGS, P.tiles * (CAND_TILE / 32) * 4, u32, tile_flags);
    P.tile_prefix = tile_prefix;
    P.tile_flags = tile_flags;
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
    const u64* scan_prefix = nullptr;
    u64 scan_total = 0;
    if (P.collapse) {
        WS(WS_SCAN_TILES, (P.chunks + 1) * 4, u32, scan_tiles);
        WS(WS_SCAN_PREFIX, (P.chunks + 2) * 8, u64, prefix);
        launch_collapsed_tiles(P.tab, P.chunks, chunk_flag, tile_count, tile_flags, scan_tiles, s);
        launch_exclusive_sum(scan_tiles, P.chunks, prefix, s);
        LAUNCHED(2, "k_collapsed_tiles");
        CK(cudaMemcpyAsync(h_u64, prefix + P.chunks, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        scan_prefix = prefix;
        scan_total = h_u64[0];
    }
    launch_candidates_scan(d_in, in_bytes, P.tab, P.chunks, P.tiles, tile_count, tile_flags, C.debug_reject_mod, chunk_flag,
                           scan_prefix, scan_total, s);
    {
        // (one count per 4 KiB of file: a multi-block sum -- a file of single positions has 400 000 tiles per 100 M positions)
        WS(WS_LSUM_A, (large_sum_tiles(P.tiles) + 1) * 4, u32, lsum_a);
        WS(WS_LSUM_B, (large_sum_tiles(P.tiles) + 2) * 8, u64, lsum_b);
        launch_exclusive_sum_large(tile_count, P.tiles, tile_prefix, lsum_a, lsum_b, s);
    }
    LAUNCHED(5, "k_candidates_scan");
    CK(cudaMemcpyAsync(h_u64, tile_prefix + P.tiles, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    P.ncand = h_u64[0];

    WS(WS_CAND_CHUNK, (P.ncand + 1) * 4, u32, cand_chunk);
    WS(WS_CAND_OFF, (P.ncand + 1) * 4, u32, cand_off);
    WS(WS_CAND_CNT, (P.ncand + 1) * 4, u32, cand_cnt);
    P.cand_chunk = cand_chunk; P.cand_off = cand_off; P.cand_cnt = cand_cnt;
    launch_candidates_list(d_in, P.tab, P.tiles, tile_flags, tile_prefix, cand_chunk, cand_off, cand_cnt,
                           P.collapse ? chunk_flag : nullptr, s);
    LAUNCHED(1, "k_candidates_list");
    return NNP_OK;
}

// Exhaustive strategy after decode_front: probe every candidate, resolve the reader's walk per
// chunk (false candidates are skipped, unresolvable chunks go to the sequential kernels) and
// produce the per-chunk position counts (and text sizes when `text`).
// `have_next`: WS_CAND_NEXT already holds the chain ends (the optimistic walk wrote them): no probe pass
int decode_plan(const void* d_in, size_t in_bytes, bool text, DecodePlan& P, bool have_front = false, bool have_next = false)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    if (!have_front) {
        int rc = decode_front(d_in, in_bytes, P);
        if (rc != NNP_OK) return rc;
    }
    if (P.chunks == 0) return NNP_OK;
    DecompressTotals* d_tot = P.d_tot;
    DecompressTotals* h_tot = reinterpret_cast<DecompressTotals*>((char*)C.pinned + 512);
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
    u32 *cand_chunk = P.cand_chunk, *cand_off = P.cand_off, *cand_cnt = P.cand_cnt;
    u64* tile_prefix = P.tile_prefix;

    WS(WS_CAND_NEXT, (P.ncand + 1) * 4, u32, cand_next);
    WS(WS_CAND_BASE, (P.ncand + 1) * 4, u32, cand_base);
    WS(WS_CHUNK_COUNT, (P.chunks + 1) * 4, u32, chunk_count);
    WS(WS_CHUNK_SLOW, (P.chunks + 1) * 4, u32, chunk_slow);
    WS(WS_CHUNK_BASE, (P.chunks + 2) * 8, u64, chunk_base);
    u32* cand_tlen = nullptr;
    u64 *cand_tbase = nullptr, *chunk_tbytes = nullptr, *chunk_tbase = nullptr;
    if (text) {
        WS(WS_TEXT_A, (P.ncand + 1) * 4, u32, tl);
        WS(WS_TEXT_B, (P.ncand + 1) * 8, u64, tb);
        WS(WS_TEXT_C, (P.chunks + 1) * 8, u64, ctb);
        WS(WS_TEXT_D, (P.chunks + 2) * 8, u64, ctbase);
        cand_tlen = tl; cand_tbase = tb; chunk_tbytes = ctb; chunk_tbase = ctbase;
    }
    P.cand_next = cand_next; P.cand_base = cand_base;
    P.cand_tbase = cand_tbase; P.chunk_tbase = chunk_tbase;
    P.chunk_count = chunk_count; P.chunk_slow = chunk_slow; P.chunk_base = chunk_base;

    if (!have_next) launch_probe_chains(d_in, P.tab, cand_chunk, cand_off, P.ncand, cand_next, cand_cnt, cand_tlen, s);
    launch_resolve_chunks(P.tab, P.chunks, tile_prefix, cand_off, cand_next, cand_cnt, cand_base, chunk_count, chunk_slow,
                          cand_tlen, cand_tbase, chunk_tbytes, s);
    launch_slow_count(d_in, P.tab, P.chunks, chunk_slow, chunk_count, chunk_tbytes, d_tot, s);
    launch_exclusive_sum(chunk_count, P.chunks, chunk_base, s);
    LAUNCHED(4, "decode plan");
    if (text) {
        launch_exclusive_sum64(chunk_tbytes, P.chunks, chunk_tbase, s);
        LAUNCHED(1, "k_exclusive_sum64");
        CK(cudaMemcpyAsync(h_u64 + 1, chunk_tbase + P.chunks, 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(h_u64, chunk_base + P.chunks, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    P.positions = h_u64[0];
    P.text_bytes = text ? h_u64[1] : 0;
    if (h_tot->error_chunk != NO_ERROR_IDX) return NNP_ERR_TRUNCATED;
    return NNP_OK;
}

// The reference hands its output buffer to the file only once it exceeds 1 MiB
// (compress_file.cpp:1395-1402), and an exception thrown while fetching chunk k leaves
// Reader::next() before the last entry of chunk k-1 is returned; this is how many records
// reach the file in that case.
u64 committed_bin_records(u64 positions_before_error)
{
    if (positions_before_error == 0) return 0;
    const u64 emitted = positions_before_error - 1;
    const u64 per_flush = (1048576 / 40) + 1;  // first count whose byte size exceeds 1 MiB
    return emitted / per_flush * per_flush;
}

int decompress_dev(const void* d_in, size_t in_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *out_bytes = 0;
    if (in_bytes == 0) return NNP_OK;
    if (d_out && ((uintptr_t)d_out & 7)) return NNP_ERR_BAD_ARG;
    CK(cudaEventRecord(C.ev[0], s));
    DecodePlan P;
    int rc = decode_walk(d_in, in_bytes, P);
    if (rc != NNP_OK) return rc;
    DecompressTotals* h_tot = reinterpret_cast<DecompressTotals*>((char*)C.pinned + 512);
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);

    // A file of single positions (shuffled training data): when every chunk holds nothing but 34-byte chains
    // the reader's walk is known without looking for it -- no candidates, no verification (k_emit_heads_only).
    if (d_out && P.chunks > 0 && !C.debug_exhaustive && C.debug_dec_direct != 2) {
        rc = decode_heads_only(d_in, P);
        if (rc != NNP_OK) return rc;
        CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        P.heads_only_chunks = h_tot->heads_only_chunks;
        if (h_tot->heads_only_chunks == P.chunks) {
            WS(WS_CHUNK_BASE, (P.chunks + 2) * 8, u64, chunk_base);
            launch_exclusive_sum(P.chunk_count, P.chunks, chunk_base, s);
            LAUNCHED(1, "k_exclusive_sum");
            CK(cudaMemcpyAsync(h_u64, chunk_base + P.chunks, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            const u64 positions = h_u64[0];
            *out_bytes = positions * 40;
            C.last_positions = positions;
            if (positions * 40 > out_cap) return NNP_ERR_CAPACITY;
            CK(cudaEventRecord(C.ev[1], s));
            launch_emit_heads_only(d_in, P.tab, P.chunks, chunk_base, positions, d_out, out_cap / 40, s);
            LAUNCHED(1, "k_emit_heads_only");
            CK(cudaEventRecord(C.ev[2], s));
            CK(cudaStreamSynchronize(s));
            C.last_kernel = "k_emit_heads_only";
            C.last_candidates = positions;
            C.last_tentative_positions = positions;
            C.last_violations = 0;
            CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
            CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[1], C.ev[2]));
            if (P.walk_status != 0) {
                *out_bytes = committed_bin_records(positions) * 40;
                C.last_positions = *out_bytes / 40;
                return P.walk_status;
            }
            return NNP_OK;
        }
    }
    // chunks of single positions among ordinary ones (shuffled data in which a few records happen to continue
    // their predecessor): each is one entry of the candidate list and k_emit_heads_chunks writes its records
    P.collapse = d_out && P.chunk_flag && P.heads_only_chunks > 0 && C.debug_reject_mod == 0 && C.debug_dec_direct != 2;
    rc = decode_front(d_in, in_bytes, P);
    if (rc != NNP_OK) return rc;
    const u64* placed_rec = nullptr;
    const u32* placed_next = nullptr;
    bool repaired = false;

    // Optimistic strategy: on files the reference wrote, the candidates are exactly the chains, so
    // every chain can be emitted at the record index its header count implies while the kernel
    // verifies the reader's walk link by link. Any violation reruns the exhaustive strategy.
    if (d_out && P.chunks > 0 && P.ncand > 0 && !C.debug_exhaustive) {
        WS(WS_CAND_REC, (P.ncand + 2) * 8, u64, cand_rec);
        WS(WS_LSUM_A, (large_sum_tiles(P.ncand) + 1) * 4, u32, lsum_a);
        WS(WS_LSUM_B, (large_sum_tiles(P.ncand) + 2) * 8, u64, lsum_b);
        WS(WS_CAND_NEXT, (P.ncand + 1) * 4, u32, cand_next);
        placed_rec = cand_rec;
        placed_next = cand_next;
        launch_mark_conflicts(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_cnt, P.ncand, s);
        launch_exclusive_sum_large(P.cand_cnt, P.ncand, cand_rec, lsum_a, lsum_b, s);
        LAUNCHED(4, "candidate record offsets");
        CK(cudaMemcpyAsync(h_u64, cand_rec + P.ncand, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const u64 positions = h_u64[0];
        C.last_candidates = P.ncand;
        C.last_tentative_positions = positions;
        C.last_violations = ~0ull;
        // (a false candidate's count may push the tentative total past the buffer: the walk is attempted
        // anyway, records behind the buffer's end are not written, and the repair below finds the real total)
        const u64 rec_limit = out_cap / 40;
        {
            CK(cudaEventRecord(C.ev[1], s));
            launch_check_chunks(P.tab, P.chunks, P.tile_prefix, &P.d_tot->violations, s);
            LAUNCHED(1, "k_check_chunks");
            const u32* collapsed = P.collapse ? P.chunk_flag : nullptr;
            if (collapsed) {
                launch_emit_heads_chunks(d_in, P.tab, P.chunks, collapsed, P.tile_prefix, cand_rec, d_out, rec_limit, s);
                LAUNCHED(1, "k_emit_heads_chunks");
            }
            if (C.pipe_dst && positions * 40 <= C.pipe_dst_cap && positions * 40 <= out_cap) {
                // host output: emit in groups of chains and copy every group's records out while the
                // next group is decoded (the record index of a chain is the prefix sum cand_rec)
                constexpr u64 PIPE_GROUPS = 16;
                const u64 per = (P.ncand + PIPE_GROUPS - 1) / PIPE_GROUPS;
                u64* h_edges = reinterpret_cast<u64*>((char*)C.pinned + 2048);
                u64 n_groups = 0;
                for (u64 g = 0; g < PIPE_GROUPS && g * per < P.ncand; ++g, ++n_groups)
                    CK(cudaMemcpyAsync(h_edges + g, cand_rec + g * per, 8, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                h_edges[n_groups] = positions;
                for (u64 g = 0; g < n_groups; ++g) {
                    const u64 lo = g * per, hi = lo + per < P.ncand ? lo + per : P.ncand;
                    launch_emit_chains_verify(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_cnt, cand_rec, P.ncand, lo, hi, d_out,
                                              rec_limit, cand_next, &P.d_tot->violations, collapsed, s);
                    LAUNCHED(1, "k_emit_chains_verify");
                    CK(cudaEventRecord(C.pipe_ev[g % 16], s));
                    CK(cudaStreamWaitEvent(C.copy_stream, C.pipe_ev[g % 16], 0));
                    const u64 b0 = h_edges[g] * 40, b1 = h_edges[g + 1] * 40;
                    if (b1 > b0)
                        CK(cudaMemcpyAsync((char*)C.pipe_dst + b0, (const char*)d_out + b0, b1 - b0, cudaMemcpyDeviceToHost,
                                           C.copy_stream));
                }
                CK(cudaEventRecord(C.pipe_ev[0], C.copy_stream));
                CK(cudaStreamWaitEvent(s, C.pipe_ev[0], 0));  // the call's final sync on `s` covers the copies
                C.pipe_dst_done = true;
            } else {
                launch_emit_chains_verify(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_cnt, cand_rec, P.ncand, 0, P.ncand, d_out,
                                          rec_limit, cand_next, &P.d_tot->violations, collapsed, s);
                LAUNCHED(1, "k_emit_chains_verify");
            }
            CK(cudaEventRecord(C.ev[2], s));
            CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            C.last_violations = h_tot->violations;
            if (h_tot->violations == 0) {  // no violation: the output is the reader's
                ++C.optimistic_hits;
                C.last_kernel = "k_emit_chains_verify";
                CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
                CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[1], C.ev[2]));
                *out_bytes = positions * 40;
                C.last_positions = positions;
                if (P.walk_status != 0) {
                    *out_bytes = committed_bin_records(positions) * 40;
                    C.last_positions = *out_bytes / 40;
                    return P.walk_status;
                }
                return NNP_OK;
            }
        }
        ++C.optimistic_misses;
        C.pipe_dst_done = false;  // whatever was copied out is overwritten by the exhaustive result
        // Repair instead of starting over: every candidate was decoded by the walk above, so its chain end
        // is known (cand_next) and the reader's walk can be resolved per chunk without a probe pass; the
        // header counts that k_mark_conflicts zeroed are listed again; and chains that already lie where the
        // resolved walk wants them (everything in front of the first false candidate) are not written twice.
        if (P.collapse) {
            // (rare twice over: the strategies below know nothing of collapsed chunks -- list every chain start again)
            P.collapse = false;
            rc = decode_front(d_in, in_bytes, P);
            if (rc != NNP_OK) return rc;
        } else {
            launch_candidates_list(d_in, P.tab, P.tiles, P.tile_flags, P.tile_prefix, P.cand_chunk, P.cand_off, P.cand_cnt, nullptr,
                                   s);
            LAUNCHED(1, "k_candidates_list");
            repaired = true;
        }
    }

    rc = decode_plan(d_in, in_bytes, false, P, true, repaired);
    if (rc != NNP_OK) return rc;
    CK(cudaEventRecord(C.ev[1], s));
    if (!d_out) {
        *out_bytes = P.positions * 40;
        return NNP_OK;
    }
    u64 positions = P.positions;
    *out_bytes = positions * 40;
    C.last_positions = positions;
    if (positions * 40 > out_cap) return NNP_ERR_CAPACITY;
    if (P.chunks > 0) {
        launch_emit_chains(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_base, P.ncand, P.chunk_base, d_out, P.d_tot,
                           repaired ? placed_rec : nullptr, repaired ? placed_next : nullptr, s);
        launch_slow_emit(d_in, P.tab, P.chunks, P.chunk_slow, P.chunk_base, d_out, s);
        LAUNCHED(2, "k_emit_chains");
    }
    CK(cudaEventRecord(C.ev[2], s));
    CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
    CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[1], C.ev[2]));
    C.last_kernel = "k_emit_chains";
    C.last_candidates = P.ncand;
    C.last_false = h_tot->false_candidates;
    for (int i = 0; i < 8; ++i) C.last_false_sample[i] = h_tot->false_sample[i];
    if (h_tot->error_chunk != NO_ERROR_IDX) return NNP_ERR_TRUNCATED;
    if (P.walk_status != 0) {
        *out_bytes = committed_bin_records(positions) * 40;
        return P.walk_status;
    }
    return NNP_OK;
}

// ---------------------------------------------------------------- sharded .binpack -> .bin
//
// Chunks are independent (hasNextChunk / readNextChunk, compress_file.cpp:468-480; the reader starts
// every chunk with a stem, :1128-1214): rank r of `world` decodes the contiguous chunk range
// [chunks * r / world, chunks * (r + 1) / world) of ONE file, and the concatenation of the ranks'
// records in rank order is the file one decompressBin run writes (:1376-1412).

// header walk of the whole file (device memory): the rank's chunk range and its bytes
int chunk_range_dev(const void* d_in, size_t in_bytes, int world, int rank, nnp_chunk_range* range, int* walk_status)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    std::memset(range, 0, sizeof(*range));
    *walk_status = 0;
    if (in_bytes == 0) return NNP_OK;
    WS(WS_CHUNK_INFO, sizeof(ChunkInfo), ChunkInfo, d_info);
    ChunkInfo* h_info = reinterpret_cast<ChunkInfo*>((char*)C.pinned + 256);
    u64 table_cap = in_bytes / 65536 + 1024;
    ChunkTable tab;
    for (int attempt = 0; attempt < 2; ++attempt) {
        WS(WS_CHUNK_START, (table_cap + 1) * 8, u64, start);
        WS(WS_CHUNK_LEN, (table_cap + 1) * 4, u32, len);
        WS(WS_CHUNK_TILE_BASE, (table_cap + 2) * 8, u64, tile_base);
        tab.start = start;
        tab.len = len;
        tab.tile_base = tile_base;
        tab.info = d_info;
        WS(WS_WALK, walk_scratch_bytes(), unsigned char, walk_scratch);
        launch_walk_chunks(d_in, in_bytes, tab, table_cap, (u32)world, (u32)rank, walk_scratch, s);
        LAUNCHED(1, "k_walk_chunks");
        CK(cudaMemcpyAsync(h_info, d_info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (h_info->chunks <= table_cap) break;
        table_cap = h_info->chunks + 16;
    }
    range->chunks_total = h_info->chunks;
    range->chunk_lo = h_info->range_lo;
    range->chunk_hi = h_info->range_hi;
    range->byte_lo = h_info->byte_lo;
    range->byte_hi = h_info->byte_hi;
    *walk_status = h_info->status;
    return NNP_OK;
}

int shard_decompress_dev(const void* d_in, size_t in_bytes, int world, int rank, void* d_out, size_t out_cap, size_t* out_bytes,
                         nnp_chunk_range* range)
{
    Context& C = g_ctx;
    *out_bytes = 0;
    CK(cudaEventRecord(C.ev[4], C.stream));
    int walk_status = 0;
    int rc = chunk_range_dev(d_in, in_bytes, world, rank, range, &walk_status);
    if (rc != NNP_OK) return rc;
    if (range->byte_hi > range->byte_lo) {
        rc = decompress_dev((const unsigned char*)d_in + range->byte_lo, range->byte_hi - range->byte_lo, d_out, out_cap, out_bytes);
        if (rc != NNP_OK) return rc;
        // the whole call, header walk of the file included
        float front = 0.f;
        CK(cudaEventElapsedTime(&front, C.ev[4], C.ev[0]));
        C.last_total_ms += front;
    }
    range->positions = *out_bytes / 40;
    // a broken header further on: every rank reports it; the chunks in front of it are decoded (the
    // reference's flush rule for the records before an error, :1395-1402, is a property of the whole
    // run and is not replayed per rank)
    return walk_status;
}

int binpack_to_plain_dev(const void* d_in, size_t in_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *out_bytes = 0;
    if (in_bytes == 0) return NNP_OK;
    DecodePlan P;
    int rc = decode_plan(d_in, in_bytes, true, P);
    if (rc != NNP_OK) return rc;
    *out_bytes = P.text_bytes;
    C.last_positions = P.positions;
    if (!d_out) return NNP_OK;
    if (P.text_bytes > out_cap) return NNP_ERR_CAPACITY;
    if (P.chunks > 0) {
        launch_emit_chains_text(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_base, P.cand_tbase, P.ncand, P.chunk_tbase,
                                d_out, P.d_tot, s);
        launch_slow_emit_text(d_in, P.tab, P.chunks, P.chunk_slow, P.chunk_tbase, d_out, s);
        LAUNCHED(2, "k_emit_chains_text");
    }
    DecompressTotals* h_tot = reinterpret_cast<DecompressTotals*>((char*)C.pinned + 512);
    CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h_tot->error_chunk != NO_ERROR_IDX) return NNP_ERR_TRUNCATED;
    if (P.walk_status != 0) {
        u64 committed = 0;
        if (P.text_bytes > 0) {
            WS(WS_TOTALS, sizeof(CompressTotals), CompressTotals, d_scratch);
            launch_text_flush_orbit(d_out, P.text_bytes, 1, &d_scratch->chunks, s);
            LAUNCHED(1, "k_text_flush_orbit");
            u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);
            CK(cudaMemcpyAsync(h_u64, &d_scratch->chunks, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            committed = h_u64[0];
        }
        *out_bytes = committed;
        return P.walk_status;
    }
    return NNP_OK;
}

// ---------------------------------------------------------------- .binpack / .bin -> HalfKP rows

// decompress_dev with the record writer replaced by the feature-row writer (halfkp.cu)
int binpack_to_halfkp_dev(const void* d_in, size_t in_bytes, int* white, int* black, void* meta, size_t cap, size_t* positions)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    *positions = 0;
    if (white && (!black || !meta)) return NNP_ERR_BAD_ARG;
    if (white && (((uintptr_t)white | (uintptr_t)black | (uintptr_t)meta) & 15)) return NNP_ERR_BAD_ARG;
    if (in_bytes == 0) return NNP_OK;
    CK(cudaEventRecord(C.ev[0], s));
    DecodePlan P;
    int rc = decode_front(d_in, in_bytes, P);
    if (rc != NNP_OK) return rc;
    DecompressTotals* h_tot = reinterpret_cast<DecompressTotals*>((char*)C.pinned + 512);
    u64* h_u64 = reinterpret_cast<u64*>((char*)C.pinned + 768);

    if (white && P.chunks > 0 && P.ncand > 0 && !C.debug_exhaustive) {  // optimistic strategy, see decompress_dev
        WS(WS_CAND_REC, (P.ncand + 2) * 8, u64, cand_rec);
        WS(WS_LSUM_A, (large_sum_tiles(P.ncand) + 1) * 4, u32, lsum_a);
        WS(WS_LSUM_B, (large_sum_tiles(P.ncand) + 2) * 8, u64, lsum_b);
        launch_mark_conflicts(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_cnt, P.ncand, s);
        launch_exclusive_sum_large(P.cand_cnt, P.ncand, cand_rec, lsum_a, lsum_b, s);
        LAUNCHED(4, "candidate record offsets");
        CK(cudaMemcpyAsync(h_u64, cand_rec + P.ncand, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const u64 n = h_u64[0];
        C.last_candidates = P.ncand;
        C.last_tentative_positions = n;
        C.last_violations = ~0ull;
        if (n <= cap) {
            CK(cudaEventRecord(C.ev[1], s));
            launch_check_chunks(P.tab, P.chunks, P.tile_prefix, &P.d_tot->violations, s);
            launch_emit_chains_halfkp_verify(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_cnt, cand_rec, P.ncand, white, black,
                                             meta, &P.d_tot->violations, s);
            LAUNCHED(2, "k_emit_chains_halfkp_verify");
            CK(cudaEventRecord(C.ev[2], s));
            CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            C.last_violations = h_tot->violations;
            if (h_tot->violations == 0) {
                ++C.optimistic_hits;
                C.last_kernel = "k_emit_chains_halfkp_verify";
                CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
                CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[1], C.ev[2]));
                *positions = n;
                return P.walk_status;
            }
        }
        ++C.optimistic_misses;
    }

    rc = decode_plan(d_in, in_bytes, false, P, true);
    if (rc != NNP_OK) return rc;
    *positions = P.positions;
    if (!white) return P.walk_status;
    if (P.positions > cap) return NNP_ERR_CAPACITY;
    CK(cudaEventRecord(C.ev[1], s));
    if (P.chunks > 0) {
        launch_emit_chains_halfkp(d_in, P.tab, P.cand_chunk, P.cand_off, P.cand_base, P.ncand, P.chunk_base, white, black, meta,
                                  P.d_tot, s);
        launch_slow_emit_halfkp(d_in, P.tab, P.chunks, P.chunk_slow, P.chunk_base, white, black, meta, s);
        LAUNCHED(2, "k_emit_chains_halfkp");
    }
    CK(cudaEventRecord(C.ev[2], s));
    CK(cudaMemcpyAsync(h_tot, P.d_tot, sizeof(DecompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
    CK(cudaEventElapsedTime(&C.last_dominant_ms, C.ev[1], C.ev[2]));
    if (h_tot->error_chunk != NO_ERROR_IDX) {
        *positions = 0;
        return NNP_ERR_TRUNCATED;
    }
    return P.walk_status;
}

int bin_to_halfkp_dev(const void* d_bin, size_t bin_bytes, int* white, int* black, void* meta, size_t cap, size_t* positions)
{
    Context& C = g_ctx;
    cudaStream_t s = C.stream;
    const u64 n = bin_bytes / 40;  // a short trailing record is dropped (:1360)
    *positions = n;
    if (!white || n == 0) return NNP_OK;
    if (!black || !meta) return NNP_ERR_BAD_ARG;
    if ((((uintptr_t)white | (uintptr_t)black | (uintptr_t)meta) & 15) || ((uintptr_t)d_bin & 7)) return NNP_ERR_BAD_ARG;
    if (n > cap) return NNP_ERR_CAPACITY;
    CompressTotals* d_tot = nullptr;
    int rc = reset_compress_totals(&d_tot);
    if (rc != NNP_OK) return rc;
    CompressTotals* h_tot = reinterpret_cast<CompressTotals*>(C.pinned);
    CK(cudaEventRecord(C.ev[0], s));
    launch_bin_halfkp(d_bin, n, white, black, meta, d_tot, s);
    LAUNCHED(1, "k_bin_halfkp");
    CK(cudaEventRecord(C.ev[2], s));
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(CompressTotals), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&C.last_total_ms, C.ev[0], C.ev[2]));
    C.last_dominant_ms = C.last_total_ms;
    if (h_tot->error_index != NO_ERROR_IDX) {
        *positions = h_tot->error_index;
        return NNP_ERR_BAD_SFEN;
    }
    return NNP_OK;
}

// ---------------------------------------------------------------- host-buffer wrappers

typedef int (*dev_fn)(const void*, size_t, void*, size_t, size_t*);

bool is_reference_error(int rc) { return rc == NNP_ERR_BAD_MAGIC || rc == NNP_ERR_CHUNK_TOO_LARGE || rc == NNP_ERR_BAD_SFEN; }

// H2D, device driver, D2H. `dev_out_cap` bounds the device-side output buffer; 0 means "the
// caller's capacity" (callers size their buffer with the out == NULL query).
int run_host(dev_fn fn, const void* in, size_t in_bytes, void* out, size_t out_cap, size_t* out_bytes, size_t dev_out_cap)
{
    Context& C = g_ctx;
    WS(WS_STAGE_IN, in_bytes + 64, unsigned char, d_in);
    // .bin -> .binpack overlaps its H2D with K1, .binpack -> .bin its D2H with the chain decoder
    const bool pipe_in = out && fn == compress_dev && in_bytes >= (64u << 20);
    const bool pipe_out = out && fn == decompress_dev;
    if (in_bytes && !pipe_in) CK(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, C.stream));
    if (!out) {
        int rc = fn(d_in, in_bytes, nullptr, 0, out_bytes);
        return is_reference_error(rc) ? NNP_OK : rc;  // the status is reported by the real call
    }
    const size_t need = dev_out_cap ? dev_out_cap : out_cap;
    WS(WS_STAGE_OUT, need + 64, unsigned char, d_out);
    size_t produced = 0;
    C.pipe_src = pipe_in ? in : nullptr;
    C.pipe_dst = pipe_out ? out : nullptr;
    C.pipe_dst_cap = out_cap;
    C.pipe_dst_done = false;
    int rc = fn(d_in, in_bytes, d_out, need, &produced);
    const bool delivered = C.pipe_dst_done;
    C.pipe_src = nullptr;
    C.pipe_dst = nullptr;
    C.pipe_dst_done = false;
    *out_bytes = produced;
    if (rc != NNP_OK && !is_reference_error(rc)) return rc;
    if (produced > out_cap) return NNP_ERR_CAPACITY;
    if (produced && !delivered) {
        CK(cudaMemcpyAsync(out, d_out, produced, cudaMemcpyDeviceToHost, C.stream));
        CK(cudaStreamSynchronize(C.stream));
    }
    return rc;
}

}  // namespace

}  // namespace nnp

using namespace nnp;

// every entry point: one call at a time per device, on the device the calling thread is bound to
#define LOCK_DEVICE()                                                \
    Context& ctx_ = g_ctx;                                           \
    std::lock_guard<std::mutex> lock_(ctx_.mutex);                   \
    if (!ctx_.ready) return NNP_ERR_NOT_INITIALISED;                 \
    if (cudaSetDevice(ctx_.device) != cudaSuccess) return NNP_ERR_CUDA;
#define REQUIRE_READY()                                              \
    LOCK_DEVICE();                                                   \
    if (!out_bytes) return NNP_ERR_BAD_ARG;

extern "C" {

namespace {

// creates the streams, events and scratch of `device` (caller holds g_init_mutex)
int init_device(int device)
{
    Context& C = g_ctxs[device];
    if (C.ready) return NNP_OK;
    if (cudaSetDevice(device) != cudaSuccess) return NNP_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NNP_ERR_NO_DEVICE;
    if (prop.major < 10) {
        C.last_cuda_error = "libnnuepack is built for sm_100a only";
        return NNP_ERR_NO_DEVICE;
    }
    if (cudaStreamCreateWithFlags(&C.own_stream, cudaStreamNonBlocking) != cudaSuccess) return NNP_ERR_CUDA;
    C.stream = C.own_stream;
    if (cudaStreamCreateWithFlags(&C.copy_stream, cudaStreamNonBlocking) != cudaSuccess) return NNP_ERR_CUDA;
    for (auto& ev : C.ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return NNP_ERR_CUDA;
    for (auto& ev : C.pipe_ev)
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return NNP_ERR_CUDA;
    if (cudaMallocHost(&C.pinned, 4096) != cudaSuccess) return NNP_ERR_NOMEM;
    // the device-side lookup tables: the counterpart of the reference's static initialisation
    // (src/chess/Bitboard.cpp:460-464)
    init_tables_compress(C.own_stream);
    init_tables_decompress(C.own_stream);
    init_tables_halfkp(C.own_stream);
    if (cudaStreamSynchronize(C.own_stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return NNP_ERR_CUDA;
    C.launches += 3;
    const char* dbg = std::getenv("NNP_DEBUG_REJECT_MOD");
    C.debug_reject_mod = dbg ? (uint32_t)std::strtoul(dbg, nullptr, 10) : 0u;
    const char* dbg2 = std::getenv("NNP_DEBUG_EXHAUSTIVE");
    C.debug_exhaustive = dbg2 && dbg2[0] == '1';
    // NNP_DEBUG_SINGLES=0: none of the routes made for files of single positions (the general pipeline for everything)
    const char* dbg3 = std::getenv("NNP_DEBUG_SINGLES");
    if (dbg3 && dbg3[0] == '0') {
        C.debug_k1_direct = 2;
        C.debug_dec_direct = 2;
    }
    C.device = device;
    C.ready = true;
    if (!g_default) g_default = &C;
    return NNP_OK;
}

int device_count()
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_ctxs[0].last_cuda_error = e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device";
        (void)cudaGetLastError();
        return 0;
    }
    return count < MAX_DEVICES ? count : MAX_DEVICES;
}

}  // namespace

int nnp_init(int device)
{
    std::lock_guard<std::mutex> lock(g_init_mutex);
    const int count = device_count();
    if (count == 0) return NNP_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return NNP_ERR_BAD_ARG;
    const int rc = init_device(device);
    if (rc != NNP_OK) return rc;
    t_ctx = &g_ctxs[device];
    return NNP_OK;
}

int nnp_init_all(int n_devices)
{
    std::lock_guard<std::mutex> lock(g_init_mutex);
    const int count = device_count();
    if (count == 0) return NNP_ERR_NO_DEVICE;
    if (n_devices <= 0 || n_devices > count) n_devices = count;
    for (int d = 0; d < n_devices; ++d) {
        const int rc = init_device(d);
        if (rc != NNP_OK) return rc;
    }
    t_ctx = &g_ctxs[0];
    if (cudaSetDevice(0) != cudaSuccess) return NNP_ERR_CUDA;
    return n_devices;
}

int nnp_bind_device(int device)
{
    if (device < 0 || device >= MAX_DEVICES || !g_ctxs[device].ready) return NNP_ERR_NOT_INITIALISED;
    if (cudaSetDevice(device) != cudaSuccess) return NNP_ERR_CUDA;
    t_ctx = &g_ctxs[device];
    return NNP_OK;
}

int nnp_device_at(int index)
{
    for (int d = 0; d < MAX_DEVICES; ++d)
        if (g_ctxs[d].ready && index-- == 0) return d;
    return -1;
}

int nnp_device_count(void)
{
    int n = 0;
    for (const Context& C : g_ctxs) n += C.ready ? 1 : 0;
    return n;
}

void nnp_internal_release_buffers(void);  // files.cu: the pinned staging buffers of the file drivers

void nnp_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_init_mutex);
    nnp_internal_release_buffers();
    for (Context& C : g_ctxs) {
        if (!C.ready) continue;
        std::lock_guard<std::mutex> dev_lock(C.mutex);
        cudaSetDevice(C.device);
        cudaStreamSynchronize(C.stream);
        for (auto& w : C.ws) {
            if (w.p) cudaFree(w.p);
            w.p = nullptr;
            w.cap = 0;
        }
        if (C.pinned) cudaFreeHost(C.pinned);
        C.pinned = nullptr;
        for (auto& ev : C.ev) {
            if (ev) cudaEventDestroy(ev);
            ev = nullptr;
        }
        for (auto& ev : C.pipe_ev) {
            if (ev) cudaEventDestroy(ev);
            ev = nullptr;
        }
        cudaStreamDestroy(C.own_stream);
        cudaStreamDestroy(C.copy_stream);
        C.stream = C.own_stream = C.copy_stream = nullptr;
        C.shard = ShardState();
        C.ready = false;
        C.device = -1;
    }
    g_default = nullptr;
    t_ctx = nullptr;  // other threads must bind again after the next nnp_init
}

const char* nnp_strerror(int status)
{
    switch (status) {
    case NNP_OK: return "ok";
    case NNP_ERR_BAD_MAGIC: return "Invalid binpack file or chunk.";
    case NNP_ERR_CHUNK_TOO_LARGE: return "Chunks size larger than supported. Malformed file?";
    case NNP_ERR_BAD_SFEN: return "Improperly encoded bin sfen";
    case NNP_ERR_TRUNCATED: return "binpack chunk or movetext is truncated";
    case NNP_ERR_NOMEM: return "out of memory";
    case NNP_ERR_BAD_ARG: return "bad argument";
    case NNP_ERR_BAD_TEXT: return "malformed .plain input";
    case NNP_ERR_CAPACITY: return "output buffer too small";
    case NNP_ERR_NO_DEVICE: return "no usable sm_100a CUDA device (libnnuepack has no CPU path)";
    case NNP_ERR_NOT_INITIALISED: return "nnp_init() has not been called";
    case NNP_ERR_CUDA: return "CUDA runtime error";
    case NNP_ERR_WINDOW: return "the overlap window behind the shard does not reach the next chain head";
    default: return "unknown status";
    }
}

int nnp_set_stream(void* cuda_stream)
{
    LOCK_DEVICE();
    cudaStreamSynchronize(g_ctx.stream);
    g_ctx.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : g_ctx.own_stream;
    return NNP_OK;
}

const char* nnp_last_cuda_error(void) { return g_ctx.last_cuda_error.c_str(); }
uint64_t nnp_kernel_launches(void)
{
    uint64_t n = 0;
    for (const Context& C : g_ctxs) n += C.launches;  // all devices of the process
    return n;
}

void* nnp_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void nnp_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int nnp_bin_to_binpack_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return compress_dev(d_bin, bin_bytes, d_out, out_cap, out_bytes);
}
int nnp_binpack_to_bin_dev(const void* d_binpack, size_t binpack_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return decompress_dev(d_binpack, binpack_bytes, d_out, out_cap, out_bytes);
}
int nnp_bin_to_binpack(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    if (!out) {
        *out_bytes = binpack_capacity_for_records(bin_bytes / 40);
        return NNP_OK;
    }
    return run_host(compress_dev, bin, bin_bytes, out, out_cap, out_bytes, binpack_capacity_for_records(bin_bytes / 40));
}
int nnp_binpack_to_bin(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    // a continuation ply costs at least 5 bits, a chain head 34 bytes: <= 64 output bytes per input byte
    size_t bound = binpack_bytes * 64 + 40;
    if (out && out_cap < bound) bound = out_cap;
    return run_host(decompress_dev, binpack, binpack_bytes, out, out_cap, out_bytes, out ? bound : 0);
}

int nnp_plain_to_binpack_dev(const void* d_plain, size_t plain_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return plain_to_binpack_dev(d_plain, plain_bytes, d_out, out_cap, out_bytes);
}
int nnp_plain_to_bin_dev(const void* d_plain, size_t plain_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return plain_to_bin_dev(d_plain, plain_bytes, d_out, out_cap, out_bytes);
}
int nnp_bin_to_plain_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return bin_to_plain_dev(d_bin, bin_bytes, d_out, out_cap, out_bytes);
}
int nnp_binpack_to_plain_dev(const void* d_binpack, size_t binpack_bytes, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return binpack_to_plain_dev(d_binpack, binpack_bytes, d_out, out_cap, out_bytes);
}
int nnp_plain_to_binpack(const void* plain, size_t plain_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return run_host(plain_to_binpack_dev, plain, plain_bytes, out, out_cap, out_bytes, 0);
}
int nnp_plain_to_bin(const void* plain, size_t plain_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return run_host(plain_to_bin_dev, plain, plain_bytes, out, out_cap, out_bytes, 0);
}
int nnp_bin_to_plain(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return run_host(bin_to_plain_dev, bin, bin_bytes, out, out_cap, out_bytes, 0);
}
int nnp_binpack_to_plain(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return run_host(binpack_to_plain_dev, binpack, binpack_bytes, out, out_cap, out_bytes, 0);
}

int nnp_binpack_to_halfkp_dev(const void* d_binpack, size_t binpack_bytes, int32_t* d_white, int32_t* d_black,
                              nnp_halfkp_meta* d_meta, size_t cap_positions, size_t* positions)
{
    LOCK_DEVICE();
    if (!positions || (binpack_bytes && !d_binpack)) return NNP_ERR_BAD_ARG;
    return binpack_to_halfkp_dev(d_binpack, binpack_bytes, d_white, d_black, d_meta, cap_positions, positions);
}
int nnp_bin_to_halfkp_dev(const void* d_bin, size_t bin_bytes, int32_t* d_white, int32_t* d_black, nnp_halfkp_meta* d_meta,
                          size_t cap_positions, size_t* positions)
{
    LOCK_DEVICE();
    if (!positions || (bin_bytes >= 40 && !d_bin)) return NNP_ERR_BAD_ARG;
    return bin_to_halfkp_dev(d_bin, bin_bytes, d_white, d_black, d_meta, cap_positions, positions);
}

int nnp_binpack_count_dev(const void* d_binpack, size_t binpack_bytes, uint64_t* n_positions)
{
    LOCK_DEVICE();
    if (!n_positions) return NNP_ERR_BAD_ARG;
    size_t bytes = 0;
    int rc = decompress_dev(d_binpack, binpack_bytes, nullptr, 0, &bytes);
    *n_positions = bytes / 40;
    return rc;
}

int nnp_generate_bin_dev(void* d_out, size_t n_positions, uint32_t max_plies, uint64_t seed)
{
    LOCK_DEVICE();
    if (n_positions == 0) return NNP_OK;
    if (!d_out || max_plies == 0 || ((uintptr_t)d_out & 7)) return NNP_ERR_BAD_ARG;
    cudaStream_t s = g_ctx.stream;
    u64 n_games = n_positions / max_plies + n_positions / ((u64)max_plies * 8) + 64;
    u64* h_total = reinterpret_cast<u64*>((char*)g_ctx.pinned + 1024);
    for (int attempt = 0; attempt < 8; ++attempt) {
        WS(WS_GAME_LEN, n_games * 4, u32, game_len);
        WS(WS_GAME_BASE, (n_games + 1) * 8, u64, game_base);
        launch_play_games(false, n_games, max_plies, seed, game_len, game_base, nullptr, n_positions, s);
        launch_exclusive_sum(game_len, n_games, game_base, s);
        LAUNCHED(2, "k_play_games<count>");
        CK(cudaMemcpyAsync(h_total, game_base + n_games, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (*h_total >= n_positions) {
            launch_play_games(true, n_games, max_plies, seed, game_len, game_base, d_out, n_positions, s);
            LAUNCHED(1, "k_play_games<write>");
            CK(cudaStreamSynchronize(s));
            return NNP_OK;
        }
        n_games = n_games * 2 + 64;
    }
    return NNP_ERR_BAD_ARG;
}

int nnp_shard_compress_begin_dev(const void* d_bin, size_t n_records, size_t own_lo, size_t own_hi, int reaches_eof,
                                 nnp_shard_info* info)
{
    LOCK_DEVICE();
    if (!info) return NNP_ERR_BAD_ARG;
    return shard_begin_dev(d_bin, n_records, own_lo, own_hi, reaches_eof, info);
}
int nnp_shard_compress_orbit(uint64_t payload_base, uint64_t carry_in, uint64_t* n_chunk_starts, uint64_t* first_start,
                             uint64_t* carry_out)
{
    LOCK_DEVICE();
    if (!n_chunk_starts || !first_start || !carry_out) return NNP_ERR_BAD_ARG;
    u64 a = 0, b = 0, c = 0;
    const int rc = shard_orbit(payload_base, carry_in, &a, &b, &c);
    *n_chunk_starts = a;
    *first_start = b;
    *carry_out = c;
    return rc;
}
int nnp_shard_compress_emit_dev(uint64_t next_start, void* d_out, size_t out_cap, size_t* out_bytes)
{
    REQUIRE_READY();
    return shard_emit_dev(next_start, d_out, out_cap, out_bytes);
}
int nnp_shard_compress_table_dev(void* d_table)
{
    LOCK_DEVICE();
    if (!d_table) return NNP_ERR_BAD_ARG;
    return shard_table_dev(d_table);
}
int nnp_shard_compress_resolve_dev(const void* d_tables, const uint64_t* payload_bytes, int world, int rank, uint64_t* carry_in,
                                   uint64_t* chunks_before, uint64_t* next_start, uint64_t* total_chunks)
{
    LOCK_DEVICE();
    if (!d_tables || !payload_bytes || !carry_in || !chunks_before || !next_start || !total_chunks) return NNP_ERR_BAD_ARG;
    uint64_t out5[5] = {0, 0, 0, 0, 0};
    const int rc = shard_resolve_dev(d_tables, payload_bytes, world, rank, out5);
    *carry_in = out5[0];
    *chunks_before = out5[1];
    *next_start = out5[2];
    *total_chunks = out5[3];
    return rc;
}

int nnp_binpack_chunk_range(const void* binpack, size_t binpack_bytes, int world, int rank, nnp_chunk_range* range)
{
    // host memory (or an mmapped file): the header walk of compress_file.cpp:500-521, 8 bytes per chunk
    if (!range || world <= 0 || rank < 0 || rank >= world || (binpack_bytes && !binpack)) return NNP_ERR_BAD_ARG;
    std::memset(range, 0, sizeof(*range));
    const unsigned char* in = static_cast<const unsigned char*>(binpack);
    std::vector<uint64_t> starts;
    uint64_t pos = 0;
    int status = NNP_OK;
    while (pos < binpack_bytes) {
        if (binpack_bytes - pos < 8 || std::memcmp(in + pos, "BINP", 4) != 0) { status = NNP_ERR_BAD_MAGIC; break; }
        const uint64_t size = (uint64_t)in[pos + 4] | ((uint64_t)in[pos + 5] << 8) | ((uint64_t)in[pos + 6] << 16) |
                              ((uint64_t)in[pos + 7] << 24);
        if (size > MAX_CHUNK_SIZE) { status = NNP_ERR_CHUNK_TOO_LARGE; break; }
        if (binpack_bytes - pos - 8 < size || size < 34) { status = NNP_ERR_TRUNCATED; break; }
        starts.push_back(pos);
        pos += 8 + size;
    }
    starts.push_back(pos);  // end of the last good chunk
    const uint64_t k = starts.size() - 1, w = (uint64_t)world, r = (uint64_t)rank;
    const uint64_t base = k / w, extra = k % w;
    const uint64_t lo = r * base + (r < extra ? r : extra), hi = lo + base + (r < extra ? 1 : 0);
    range->chunks_total = k;
    range->chunk_lo = lo;
    range->chunk_hi = hi;
    range->byte_lo = starts[lo];
    range->byte_hi = starts[hi];
    return status;
}
int nnp_binpack_chunk_range_dev(const void* d_binpack, size_t binpack_bytes, int world, int rank, nnp_chunk_range* range)
{
    LOCK_DEVICE();
    if (!range || world <= 0 || rank < 0 || rank >= world || (binpack_bytes && !d_binpack)) return NNP_ERR_BAD_ARG;
    int walk_status = 0;
    const int rc = chunk_range_dev(d_binpack, binpack_bytes, world, rank, range, &walk_status);
    return rc != NNP_OK ? rc : walk_status;
}
int nnp_shard_decompress_dev(const void* d_binpack, size_t binpack_bytes, int world, int rank, void* d_out, size_t out_cap,
                             size_t* out_bytes, nnp_chunk_range* range)
{
    REQUIRE_READY();
    if (!range || world <= 0 || rank < 0 || rank >= world || (binpack_bytes && !d_binpack)) return NNP_ERR_BAD_ARG;
    return shard_decompress_dev(d_binpack, binpack_bytes, world, rank, d_out, out_cap, out_bytes, range);
}

int nnp_debug_config(const char* key, uint64_t value)
{
    std::lock_guard<std::mutex> lock_(g_ctx.mutex);
    if (!key) return NNP_ERR_BAD_ARG;
    if (!std::strcmp(key, "exhaustive")) g_ctx.debug_exhaustive = value != 0;
    else if (!std::strcmp(key, "k1_per_record")) g_ctx.debug_k1_per_record = value != 0;
    else if (!std::strcmp(key, "k1_heads")) g_ctx.debug_k1_heads = value != 0;
    else if (!std::strcmp(key, "k1_direct")) g_ctx.debug_k1_direct = (int)value;
    else if (!std::strcmp(key, "dec_direct")) g_ctx.debug_dec_direct = (int)value;
    else if (!std::strcmp(key, "k1_walk")) g_ctx.debug_k1_walk = value != 0;
    else if (!std::strcmp(key, "k1_runs")) g_ctx.debug_k1_runs = value != 0;
    else if (!std::strcmp(key, "reject_mod")) g_ctx.debug_reject_mod = (uint32_t)value;
    else if (!std::strcmp(key, "walk_seg_bytes")) set_walk_segment_bytes(value);
    else return NNP_ERR_BAD_ARG;
    return NNP_OK;
}

int nnp_decode_stats(uint64_t* out14)
{
    if (!out14) return NNP_ERR_BAD_ARG;
    out14[0] = g_ctx.optimistic_hits;
    out14[1] = g_ctx.optimistic_misses;
    out14[2] = g_ctx.last_candidates;
    out14[3] = g_ctx.last_tentative_positions;
    out14[4] = g_ctx.last_violations;
    out14[5] = g_ctx.last_false;
    for (int i = 0; i < 8; ++i) out14[6 + i] = g_ctx.last_false_sample[i];
    return NNP_OK;
}

const char* nnp_last_dominant_kernel(void) { return g_ctx.last_kernel; }
uint64_t nnp_last_positions(void) { return g_ctx.last_positions; }

int nnp_last_timing(float* total_ms, float* dominant_kernel_ms)
{
    if (total_ms) *total_ms = g_ctx.last_total_ms;
    if (dominant_kernel_ms) *dominant_kernel_ms = g_ctx.last_dominant_ms;
    return NNP_OK;
}

}  // extern "C"

// ---- entry points implemented in plain.cu / generate.cu are declared there ----
