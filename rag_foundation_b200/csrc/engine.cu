// librf_b200.so -- engine and C-ABI (include/rf_b200.h).
//
// One engine owns one GPU's share of the chunk index: a feature arena F[capacity, dim] int8 in HBM (dim = 256, 512 or 1024)
// with three sidecar arrays (store-segment word, sum of squares, nothing else is per-row), rows in
// append order (row index + id_base = global chunk id).  Host-side it keeps, per store, the list of
// row extents that hold the store's rows, so a store-scoped query scans only those rows; the
// per-row segment word is still checked in the kernel, so extents are a performance hint only.
//
// The engine stands where the reference's remote File Search service stands behind
// backend/app/services/gemini_rag.py (GeminiRag :242-599 / MockGeminiRag :602-718); the adapter
// object itself is re-created per request (get_rag_client, :721-725), so all state lives here.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <unistd.h>

#include "rf_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define RF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(RF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct Extent {
    uint32_t lo, hi;
};

void extents_append(std::vector<Extent> &v, uint32_t lo, uint32_t hi) {
    if (lo >= hi) return;
    if (!v.empty() && v.back().hi == lo) v.back().hi = hi;
    else v.push_back({lo, hi});
}

// Remove rows [lo, hi) from an extent list (any order, disjoint).
void extents_subtract(std::vector<Extent> &v, uint32_t lo, uint32_t hi) {
    std::vector<Extent> out;
    out.reserve(v.size() + 1);
    for (const Extent &x : v) {
        if (x.hi <= lo || x.lo >= hi) { out.push_back(x); continue; }
        if (x.lo < lo) out.push_back({x.lo, lo});
        if (x.hi > hi) out.push_back({hi, x.hi});
    }
    v.swap(out);
}

struct Store {
    std::string name;
    std::vector<Extent> ext;
    bool dropped = false;
};

struct Doc {
    uint32_t store;
    std::vector<Extent> ext;
};

constexpr uint32_t kMaxExtPerQuery = 64;   // more are coalesced (the row mask keeps it exact)
constexpr uint32_t kTableMinQueries = 4;   // batches of this many differently-scoped queries read their plans from the store table
constexpr uint32_t kTilesPerBlockTarget = 160;   // ~1.3 MB streamed per block: start-up and merge stay under 10 %

struct DeviceBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(n, 4096);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(n, 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct MappedBuf {   // pinned host memory the GPU writes results into directly (zero-copy)
    void *h = nullptr;
    void *d = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(n, 4096);
        cudaError_t e = cudaHostAlloc(&h, want, cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        e = cudaHostGetDevicePointer(&d, h, 0);
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        cap = 0;
    }
};

struct CopyPool;
void copy_pool_destroy(CopyPool *p);

struct SearchCtx {
    cudaStream_t stream = nullptr;
    MappedBuf m_out;
    PinnedBuf h_in, h_out;
    DeviceBuf d_in, d_out, d_partial, d_tickets;
    uint32_t launches = 0;   // parity selects the sync set
    uint32_t seq = 0;        // completion sequence number written by the kernel into m_out
    // the search in flight on this context, between rf_search_begin and rf_search_end
    struct Pending {
        bool active = false;
        bool mapped = false;       // results land in m_out (spin on the completion word) / are copied to h_out (stream sync)
        uint32_t nq = 0, k = 0;
        size_t flag_off = 0;
        uint32_t done_seq = 0;
        bool has_q = false;        // text searches: the query vector follows the results in h_out
        std::chrono::steady_clock::time_point t_launched;
    } pend;
};

// Scratch of the device-resident searches on one caller stream (rf_search_keys_device*): per STREAM, not
// per scope -- a single-scope plan rides in the kernel parameters (rf::kInlineExt extents), so nothing
// scope-keyed lives on the device and a service with 10 k tenants holds no more than kMaxStreamStates
// of these.  Least recently used states are released when the table is full.
struct StreamState {
    std::mutex mu;                    // one search at a time prepares a launch on this stream
    uint64_t last_use = 0;
    bool overlap = false;             // rf_stream_set_overlap: scan launches may overlap their predecessor's tail
    DeviceBuf blob, partial, tickets;
    PinnedBuf h_blob;                 // per-query plans (rf_search_keys_device_scoped): staging + "copy done" event
    cudaEvent_t h_blob_free = nullptr;
    DeviceBuf gemm_lists, gemm_keys_a, gemm_floors;   // batched tensor-core path scratch
    uint32_t launches = 0;
    void release() {
        blob.release(); partial.release(); tickets.release(); h_blob.release();
        gemm_lists.release(); gemm_keys_a.release(); gemm_floors.release();
        if (h_blob_free) cudaEventDestroy(h_blob_free);
        h_blob_free = nullptr;
    }
};
constexpr size_t kMaxStreamStates = 64;

// Device-resident copy of every store's extents (rf::StoreEntry + flat lo / hi arrays), rebuilt when the
// extents change (engine epoch).  Batches of store-scoped queries read their plans from it in the kernel, so
// the host neither builds nor uploads a plan per query.  Readers hold `mu` shared from the moment they take the
// pointers until their launch is enqueued; a rebuild holds it exclusively (and frees the old buffers with
// cudaFree, which waits for every kernel already enqueued).
struct StoreTable {
    std::shared_mutex mu;
    uint64_t epoch = ~0ull;
    DeviceBuf entries, lo, hi;
    std::vector<uint32_t> h_next, h_tiles;     // host mirror: extents / tiles per store (grid sizing, the 64-extent check)
    uint32_t n_stores = 0;
};

struct PlanBlob {  // host staging of everything one launch needs besides F/seg/ff
    std::vector<uint8_t> bytes;
    size_t off_q = 0, off_plans = 0, off_lo = 0, off_hi = 0, off_tile0 = 0;
    uint32_t max_tiles = 0;
};

// ---- staging of documents that arrive in pageable host memory ------------------------------------------
// A copy from pageable memory makes the driver stage through its own small pinned buffers (~10-15 GB/s).
// The engine stages itself: the document is cut into chunks, helper threads (and the ingesting thread,
// whenever it has nothing to enqueue) memcpy chunks into a pinned buffer, and the ingesting thread enqueues
// each chunk's DMA -- and the tokeniser launch of the chunk before it -- as soon as the chunk is there, in
// order.  Host memcpy, PCIe transfer and the tokenise kernels overlap.  The helpers make no CUDA call and
// wait for nothing: the buffer holds a whole window of the document (32 MB; the upload cap is 25 MB), so no
// slot is ever reused while its copy is in flight.
constexpr size_t kStageWindowBytes = 32u << 20;   // pinned staging buffer per engine: one window of a document (the upload cap is 25 MB)
constexpr size_t kStageMinBytes = 1u << 20;       // smaller documents take one copy, no helpers

struct CopyPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;
    uint64_t job_gen = 0;
    // the job in flight (valid while job_live): stage src[0, n) into dst[0, n), chunk by chunk, in order
    const uint8_t *src = nullptr;
    uint8_t *dst = nullptr;
    size_t n = 0, chunk = 0;
    uint32_t n_chunks = 0;
    std::atomic<uint32_t> next{0};        // next chunk to claim
    std::atomic<uint32_t> helpers_in{0};  // helpers currently inside the job
    std::atomic<bool> job_live{false};
    std::vector<std::atomic<uint8_t>> staged;

    // Claim and stage one chunk (pure host work: no CUDA call, nothing to wait for); false when none is left.
    bool stage_one() {
        const uint32_t c = next.fetch_add(1, std::memory_order_relaxed);
        if (c >= n_chunks) return false;
        const size_t off = static_cast<size_t>(c) * chunk;
        rf::stage_copy(dst + off, src + off, std::min(chunk, n - off));
        staged[c].store(1, std::memory_order_release);
        return true;
    }
    void begin(const uint8_t *s, uint8_t *d, size_t bytes, size_t chunk_bytes) {
        src = s; dst = d; n = bytes; chunk = chunk_bytes;
        n_chunks = static_cast<uint32_t>((bytes + chunk_bytes - 1) / chunk_bytes);
        next.store(0);
        if (staged.size() < n_chunks) staged = std::vector<std::atomic<uint8_t>>(n_chunks);
        for (uint32_t c = 0; c < n_chunks; ++c) staged[c].store(0, std::memory_order_relaxed);
        job_live.store(true, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(mu);
            ++job_gen;
        }
        cv.notify_all();
    }
    void end() {
        job_live.store(false, std::memory_order_release);
        while (helpers_in.load(std::memory_order_acquire)) std::this_thread::yield();
    }
    void helper() {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || job_gen != seen; });
                if (stop) return;
                seen = job_gen;
                if (!job_live.load()) continue;
                helpers_in.fetch_add(1);
            }
            while (job_live.load(std::memory_order_acquire) && stage_one()) {}
            helpers_in.fetch_sub(1);
            // a worker that uploads document after document finds the helpers awake: spin briefly for the next job
            const auto t0 = std::chrono::steady_clock::now();
            while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
                bool again = false;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    again = stop || job_gen != seen;
                }
                if (again) break;
            }
        }
    }
};

CopyPool *copy_pool_create(unsigned n_threads) {
    CopyPool *p = new (std::nothrow) CopyPool();
    if (!p) return nullptr;
    for (unsigned i = 0; i < n_threads; ++i) p->threads.emplace_back([p] { p->helper(); });
    return p;
}
void copy_pool_destroy(CopyPool *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv.notify_all();
    for (std::thread &t : p->threads) t.join();
    delete p;
}

}  // namespace

struct rf_engine {
    rf_config cfg{};
    bool reader = false;             // attached to another process's arena over CUDA IPC: searches only (rf_engine_attach)
    uint32_t dim = RF_DIM;           // features (= bytes) per row: 256, 512 or 1024
    uint32_t tile_rows = 32;         // rows per scan tile: rf::scan_tile_rows(dim)
    int sm_count = 148;
    int8_t *F = nullptr;
    uint32_t *seg = nullptr;
    int32_t *ff = nullptr;
    uint16_t *zipf_bucket = nullptr; // device, 65536 entries, built on first synthetic ingest
    uint64_t hbm_bytes = 0;

    std::shared_mutex meta_mu;       // stores / docs / n_rows
    std::vector<Store> stores;
    std::unordered_map<std::string, uint32_t> store_by_name;
    std::unordered_map<uint64_t, Doc> docs;
    uint64_t n_rows = 0;             // published rows
    std::atomic<uint64_t> epoch{0};  // bumps whenever extents change
    std::atomic<uint64_t> tomb_gen{0};   // bumps whenever rows are tombstoned
    std::mutex df_mu;                // RF-1w statistics cache: scope -> (generation, df[dim] + n)
    std::map<std::vector<uint32_t>, std::pair<uint64_t, std::vector<uint64_t>>> df_cache;

    std::mutex ingest_mu;            // one ingest at a time (shared scratch + append cursor + free list)
    std::vector<Extent> free_ext;    // rows of deleted documents / dropped stores, sorted, coalesced: reused first-fit
    uint64_t free_rows = 0;
    cudaStream_t ingest_stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // the document's DMA runs here, back to back, while the tokeniser works on ingest_stream
    cudaEvent_t copy_ev[16] = {};            // "chunk c has arrived" (reused round-robin; waited on by ingest_stream in order)
    cudaEvent_t ingest_idle = nullptr;       // the previous document's kernels have finished with the text scratch
    DeviceBuf sc_text, sc_state, sc_bucket, sc_cstart, sc_end, sc_ctl, sc_spans, sc_deferred;
    PinnedBuf sc_stage;              // staging ring for documents in pageable host memory
    size_t stage_chunk = 2u << 20;   // bytes per copy chunk (RF_STAGE_CHUNK_KB) and helper threads (RF_STAGE_THREADS)
    uint32_t stage_threads = 1;      // (on the measured hosts aggregate memcpy bandwidth stops scaling at two threads)
    PinnedBuf sc_ctl_host;           // control words read back per document
    struct CopyPool *copy_pool = nullptr;   // helper threads that fill the staging ring (created on the first large document)
    std::atomic<uint64_t> ingest_bytes{0}, ingest_ns{0}, ingest_kernel_ns{0};
    cudaEvent_t ingest_ev[4] = {};   // ingest stream: before the first / after the last tokeniser launch, before / after the rows kernel

    std::mutex ctx_mu;
    std::condition_variable ctx_cv;
    std::vector<SearchCtx *> free_ctx;
    std::vector<SearchCtx *> all_ctx;

    StoreTable tbl;
    std::mutex plan_mu;              // guards the table only; a launch holds its StreamState's own mutex
    std::unordered_map<void *, std::shared_ptr<StreamState>> stream_states;
    uint64_t stream_tick = 0;

    std::atomic<uint64_t> searches{0};
    std::atomic<uint64_t> launches{0};
    uint32_t blocks_override = 0;
    bool table_enabled = true;       // RF_STORE_TABLE=0: host-built plans for every batch (A/B comparison)
    bool gemm_enabled = true;        // RF_GEMM=0 forces the scan kernel for batched device searches
    bool gemm_pair = true;           // RF_GEMM_PAIR=0 keeps batches of more than 256 queries on the single-CTA kernel
    uint32_t gemm_min_queries = 0;   // 0: cost model (gemm_pays); RF_GEMM_MIN_QUERIES=n forces "n queries or more"
    uint32_t gemm_sample = 65536;    // rows of the first (floor-finding) pass
    uint32_t gemm_slices_a = 0;      // 0 = as many as fit
    int scan_variant = rf::kScanVariantTma6x12;
    unsigned long long *debug_ts = nullptr;  // RF_SCAN_DEBUG=1 (diagnostics)
    uint32_t dbg_flags = 0;                  // RF_SCAN_DBG (diagnostics)
    bool profile = false;                    // RF_PROFILE=1: host-side phase times of rf_search on stderr at destroy
    std::atomic<uint64_t> prof_ns[4]{}, prof_n{0};
    size_t debug_cap = 0;
};

namespace {

using rf::ScanArgs;
using rf::ScanPlan;

// Collect the extents to scan for one scope (under a shared lock on meta_mu).
void gather_extents(rf_engine *e, const uint32_t *segs, uint32_t n_segs, std::vector<Extent> &out) {
    out.clear();
    for (uint32_t i = 0; i < n_segs; ++i) {
        const uint32_t s = segs[i];
        if (s >= e->stores.size() || e->stores[s].dropped) continue;
        bool seen = false;
        for (uint32_t j = 0; j < i; ++j) seen |= (segs[j] == s);
        if (seen) continue;
        out.insert(out.end(), e->stores[s].ext.begin(), e->stores[s].ext.end());
    }
    if (out.size() > 1) {
        std::sort(out.begin(), out.end(), [](const Extent &a, const Extent &b) { return a.lo < b.lo; });
        std::vector<Extent> merged;
        for (const Extent &x : out) {
            if (!merged.empty() && merged.back().hi >= x.lo) merged.back().hi = std::max(merged.back().hi, x.hi);
            else merged.push_back(x);
        }
        out.swap(merged);
    }
    if (out.size() > kMaxExtPerQuery) {  // keep the kMaxExtPerQuery-1 widest gaps as split points
        std::vector<std::pair<uint32_t, size_t>> gaps;
        for (size_t i = 1; i < out.size(); ++i) gaps.push_back({out[i].lo - out[i - 1].hi, i});
        std::nth_element(gaps.begin(), gaps.begin() + (kMaxExtPerQuery - 1), gaps.end(),
                         [](const auto &a, const auto &b) { return a.first > b.first; });
        std::vector<char> split(out.size(), 0);
        for (size_t i = 0; i < kMaxExtPerQuery - 1; ++i) split[gaps[i].second] = 1;
        std::vector<Extent> merged;
        for (size_t i = 0; i < out.size(); ++i) {
            if (i == 0 || split[i]) merged.push_back(out[i]);
            else merged.back().hi = out[i].hi;
        }
        out.swap(merged);
    }
}

// Build the launch blob for nq queries with CSR scopes.  q may be null (device-resident queries).
// Keep only the parts of `ext` (sorted, disjoint rows) that fall inside `lim` (sorted, disjoint rows).
void intersect_extents(std::vector<Extent> &ext, const std::vector<Extent> &lim) {
    std::vector<Extent> out;
    size_t i = 0, j = 0;
    while (i < ext.size() && j < lim.size()) {
        const uint32_t lo = std::max(ext[i].lo, lim[j].lo), hi = std::min(ext[i].hi, lim[j].hi);
        if (lo < hi) out.push_back({lo, hi});
        if (ext[i].hi < lim[j].hi) ++i; else ++j;
    }
    ext.swap(out);
}

int build_blob(rf_engine *e, const int8_t *q, uint32_t nq, const uint32_t *store_segs, const uint32_t *seg_off,
               bool shared_scope, PlanBlob &b, const std::vector<Extent> *restrict_rows = nullptr) {
    std::vector<ScanPlan> plans(shared_scope ? 1 : nq);
    std::vector<uint32_t> lo, hi, tile0;
    std::vector<Extent> ext;
    b.max_tiles = 0;
    {
        std::shared_lock<std::shared_mutex> lk(e->meta_mu);
        for (size_t i = 0; i < plans.size(); ++i) {
            const uint32_t s0 = seg_off[i], s1 = seg_off[i + 1];
            if (s1 < s0 || s1 - s0 > RF_SCOPE_MAX) return fail(RF_EINVAL, "scope of query %zu has %u segments (max %u)", i, s1 - s0, RF_SCOPE_MAX);
            ScanPlan &p = plans[i];
            memset(&p, 0, sizeof p);
            p.n_scope = s1 - s0;
            for (uint32_t j = 0; j < RF_SCOPE_MAX; ++j) p.scope[j] = j < p.n_scope ? store_segs[s0 + j] : RF_TOMBSTONE;
            gather_extents(e, store_segs + s0, p.n_scope, ext);
            if (restrict_rows) {
                intersect_extents(ext, *restrict_rows);
                // Unlike store extents these cannot be widened (the row mask knows stores, not
                // documents): callers split long range lists over several calls (engine.py does).
                if (ext.size() > kMaxExtPerQuery)
                    return fail(RF_EINVAL, "row restriction leaves %zu extents (max %u per call)", ext.size(), kMaxExtPerQuery);
            }
            p.ext_off = static_cast<uint32_t>(lo.size());
            p.n_ext = static_cast<uint32_t>(ext.size());
            uint32_t tiles = 0;
            for (const Extent &x : ext) {
                lo.push_back(x.lo);
                hi.push_back(x.hi);
                tile0.push_back(tiles);
                tiles += (x.hi - x.lo + e->tile_rows - 1) / e->tile_rows;
            }
            tile0.push_back(tiles);
            p.total_tiles = tiles;
            b.max_tiles = std::max(b.max_tiles, tiles);
        }
    }
    if (lo.empty()) { lo.push_back(0); hi.push_back(0); }
    auto align = [](size_t x) { return (x + 15) & ~static_cast<size_t>(15); };
    b.off_q = 0;
    b.off_plans = align(q ? static_cast<size_t>(nq) * e->dim : 0);
    b.off_lo = align(b.off_plans + plans.size() * sizeof(ScanPlan));
    b.off_hi = align(b.off_lo + lo.size() * 4);
    b.off_tile0 = align(b.off_hi + hi.size() * 4);
    const size_t total = align(b.off_tile0 + tile0.size() * 4);
    b.bytes.assign(total, 0);
    if (q) memcpy(b.bytes.data() + b.off_q, q, static_cast<size_t>(nq) * e->dim);
    memcpy(b.bytes.data() + b.off_plans, plans.data(), plans.size() * sizeof(ScanPlan));
    memcpy(b.bytes.data() + b.off_lo, lo.data(), lo.size() * 4);
    memcpy(b.bytes.data() + b.off_hi, hi.data(), hi.size() * 4);
    memcpy(b.bytes.data() + b.off_tile0, tile0.data(), tile0.size() * 4);
    return RF_OK;
}

// Take the store table for reading (rebuilding it first if the extents changed since it was built).
int table_acquire(rf_engine *e, std::shared_lock<std::shared_mutex> &lk) {
    StoreTable &t = e->tbl;
    for (int attempt = 0; attempt < 4; ++attempt) {
        lk = std::shared_lock<std::shared_mutex>(t.mu);
        if (t.epoch == e->epoch.load()) return RF_OK;
        lk.unlock();
        std::unique_lock<std::shared_mutex> ul(t.mu);
        if (t.epoch == e->epoch.load()) continue;
        std::vector<rf::StoreEntry> ent;
        std::vector<uint32_t> lo, hi;
        std::vector<Extent> ext;
        uint64_t now;
        {
            std::shared_lock<std::shared_mutex> ml(e->meta_mu);
            now = e->epoch.load();       // extents cannot change while meta_mu is held
            ent.resize(e->stores.size());
            t.h_next.assign(e->stores.size(), 0);
            t.h_tiles.assign(e->stores.size(), 0);
            for (uint32_t sg = 0; sg < e->stores.size(); ++sg) {
                const Store &st = e->stores[sg];
                rf::StoreEntry &en = ent[sg];
                en = rf::StoreEntry{static_cast<uint32_t>(lo.size()), 0, 0, 0};
                if (st.dropped || st.ext.empty()) continue;
                const std::vector<Extent> *src = &st.ext;
                if (st.ext.size() > kMaxExtPerQuery) {   // coalesce across the smallest gaps (the row mask keeps it exact)
                    gather_extents(e, &sg, 1, ext);
                    src = &ext;
                }
                for (const Extent &x : *src) {
                    lo.push_back(x.lo);
                    hi.push_back(x.hi);
                    en.total_tiles += (x.hi - x.lo + e->tile_rows - 1) / e->tile_rows;
                }
                en.n_ext = static_cast<uint32_t>(src->size());
                t.h_next[sg] = en.n_ext;
                t.h_tiles[sg] = en.total_tiles;
            }
        }
        if (lo.empty()) { lo.push_back(0); hi.push_back(0); }
        if (ent.empty()) ent.push_back(rf::StoreEntry{0, 0, 0, 0});
        RF_CUDA(cudaSetDevice(e->cfg.device));
        RF_CUDA(t.entries.reserve(ent.size() * sizeof(rf::StoreEntry) * 2));     // (headroom: fewer re-allocations as stores are added)
        RF_CUDA(t.lo.reserve(lo.size() * 4 * 2));
        RF_CUDA(t.hi.reserve(hi.size() * 4 * 2));
        RF_CUDA(cudaDeviceSynchronize());        // kernels that read the previous contents have finished
        RF_CUDA(cudaMemcpy(t.entries.p, ent.data(), ent.size() * sizeof(rf::StoreEntry), cudaMemcpyHostToDevice));
        RF_CUDA(cudaMemcpy(t.lo.p, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice));
        RF_CUDA(cudaMemcpy(t.hi.p, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
        // (a copy from pageable memory may return once the bytes are staged: the searches that read the table run on
        // non-blocking streams, which do not order with the legacy default stream the DMA is on)
        RF_CUDA(cudaDeviceSynchronize());
        t.n_stores = static_cast<uint32_t>(t.h_next.size());
        t.epoch = now;
    }
    return fail(RF_EBUSY, "store table kept changing under a search");
}

// Can this batch take its plans from the store table?  (Every scope within RF_SCOPE_MAX stores and
// rf::kInlineExt extents in all.)  Also the largest tile count of any query, for the grid.  Caller holds the table.
bool table_batch_ok(const rf_engine *e, uint32_t nq, const uint32_t *store_segs, const uint32_t *seg_off, uint32_t *max_tiles) {
    const StoreTable &t = e->tbl;
    uint32_t mt = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        const uint32_t s0 = seg_off[i], s1 = seg_off[i + 1];
        if (s1 < s0 || s1 - s0 > RF_SCOPE_MAX) return false;
        uint32_t ext = 0, tiles = 0;
        for (uint32_t j = s0; j < s1; ++j) {
            const uint32_t sg = store_segs[j];
            if (sg >= t.n_stores) continue;
            bool dup = false;
            for (uint32_t x = s0; x < j; ++x) dup |= (store_segs[x] == sg);
            if (dup) continue;
            ext += t.h_next[sg];
            tiles += t.h_tiles[sg];
        }
        if (ext > rf::kInlineExt) return false;
        mt = std::max(mt, tiles);
    }
    *max_tiles = mt;
    return true;
}

void fill_table_args(const rf_engine *e, ScanArgs &a, const uint8_t *d_blob, size_t off_segoff, size_t off_segs) {
    a.st_tbl = static_cast<const rf::StoreEntry *>(e->tbl.entries.p);
    a.st_lo = static_cast<const uint32_t *>(e->tbl.lo.p);
    a.st_hi = static_cast<const uint32_t *>(e->tbl.hi.p);
    a.st_n_stores = e->tbl.n_stores;
    a.q_seg_off = reinterpret_cast<const uint32_t *>(d_blob + off_segoff);
    a.q_segs = reinterpret_cast<const uint32_t *>(d_blob + off_segs);
}

uint32_t pick_blocks(rf_engine *e, uint32_t nq, uint32_t max_tiles) {
    if (e->blocks_override) return std::max(1u, std::min(std::min(e->blocks_override, 1024u), std::max(max_tiles, 1u)));
    const uint32_t wave = rf::scan_default_blocks_per_query(e->sm_count, e->scan_variant);
    uint32_t x = (max_tiles + kTilesPerBlockTarget - 1) / kTilesPerBlockTarget;
    const uint32_t fill = (wave + nq - 1) / nq;  // at least one full wave over all queries
    x = std::max(x, fill);
    if (nq == 1) x = wave;                       // single query: exactly one resident wave
    else if (nq < wave && max_tiles >= 8 && x <= 16) {
        // A batch smaller than a wave (store-sharded batches leave a rank 1/G of the queries): blocks per query such
        // that the launch is whole waves -- a last wave that is mostly empty cannot keep HBM busy on its own (128 queries x 3
        // blocks = 1.3 waves ran 12 % slower than x 2).  Cost of a choice ~ waves x (tiles per block + ~12 tiles of
        // block start-up and merge).
        double best = 1e30;
        uint32_t best_x = x;
        for (uint32_t c = 1; c <= 32 && c <= max_tiles; ++c) {
            const double waves = static_cast<double>((static_cast<uint64_t>(nq) * c + wave - 1) / wave);
            const double cost = waves * (static_cast<double>(max_tiles) / c + 12.0);
            if (cost < best - 1e-9) { best = cost; best_x = c; }
        }
        x = best_x;
    }
    x = std::min(x, std::max(max_tiles, 1u));
    x = std::min(x, 1024u);                      // the final merge is two tournament levels of 32
    return std::max(x, 1u);
}

SearchCtx *ctx_acquire(rf_engine *e) {
    std::unique_lock<std::mutex> lk(e->ctx_mu);
    if (!e->ctx_cv.wait_for(lk, std::chrono::seconds(5), [&] { return !e->free_ctx.empty(); })) return nullptr;
    SearchCtx *c = e->free_ctx.back();
    e->free_ctx.pop_back();
    return c;
}
void ctx_release(rf_engine *e, SearchCtx *c) {
    {
        std::lock_guard<std::mutex> lk(e->ctx_mu);
        e->free_ctx.push_back(c);
    }
    e->ctx_cv.notify_one();
}
struct CtxGuard {
    rf_engine *e;
    SearchCtx *c;
    ~CtxGuard() { if (c) ctx_release(e, c); }
};

void fill_args(rf_engine *e, ScanArgs &a, const uint8_t *d_blob, const PlanBlob &b, const int8_t *q_dev, uint32_t k,
               bool shared) {
    a.F = e->F;
    a.seg = e->seg;
    a.ff = e->ff;
    a.q = q_dev;
    if (d_blob) {   // null when the plan (and query) ride in the kernel parameters
        if (!q_dev) a.q = reinterpret_cast<const int8_t *>(d_blob + b.off_q);
        a.plans = reinterpret_cast<const ScanPlan *>(d_blob + b.off_plans);
        a.ext_lo = reinterpret_cast<const uint32_t *>(d_blob + b.off_lo);
        a.ext_hi = reinterpret_cast<const uint32_t *>(d_blob + b.off_hi);
        a.ext_tile0 = reinterpret_cast<const uint32_t *>(d_blob + b.off_tile0);
    }
    a.id_base = static_cast<uint32_t>(e->cfg.id_base);
    a.k = k;
    a.shared_plan = shared ? 1u : 0u;
    a.inline_plan = 0;
    a.debug_ts = e->debug_ts;
    a.dbg_flags = e->dbg_flags;
}

// Single plan with few extents: copy it into the kernel parameters (host copy of the blob).
void maybe_inline_plan(ScanArgs &a, const PlanBlob &b, uint32_t nq, bool shared) {
    if (!(shared || nq == 1)) return;
    const ScanPlan *p = reinterpret_cast<const ScanPlan *>(b.bytes.data() + b.off_plans);
    if (p->n_ext > rf::kInlineExt) return;
    a.plan0 = *p;
    const uint32_t *lo = reinterpret_cast<const uint32_t *>(b.bytes.data() + b.off_lo) + p->ext_off;
    const uint32_t *hi = reinterpret_cast<const uint32_t *>(b.bytes.data() + b.off_hi) + p->ext_off;
    const uint32_t *t0 = reinterpret_cast<const uint32_t *>(b.bytes.data() + b.off_tile0) + p->ext_off;
    for (uint32_t i = 0; i < p->n_ext; ++i) { a.inl_lo[i] = lo[i]; a.inl_hi[i] = hi[i]; }
    for (uint32_t i = 0; i <= p->n_ext; ++i) a.inl_tile0[i] = t0[i];
    a.inline_plan = 1;
}

// The zero-initialised sync buffer holds TWO sets (consecutive launches alternate, so a query
// that overlaps the previous one's tail under programmatic dependent launch never shares its
// floor / tile counter); per set, for a capacity of n queries: n u64 floors, n u32 tickets,
// n u32 tile counters.
constexpr size_t kSyncBytesPerQuery = 2 * 16;
void set_sync_bufs(ScanArgs &a, const DeviceBuf &buf, uint32_t parity) {
    const size_t cap_q = (buf.cap - 8) / kSyncBytesPerQuery;
    uint8_t *base = static_cast<uint8_t *>(buf.p) + (parity & 1u) * cap_q * 16;
    a.floors = reinterpret_cast<uint64_t *>(base);
    a.tickets = reinterpret_cast<uint32_t *>(base + cap_q * 8);
    a.tile_ctr = reinterpret_cast<uint32_t *>(base + cap_q * 12);
    // one finished-query counter per set sits in the last 8 bytes of the buffer (reserve() adds them)
    a.done_count = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(buf.p) + buf.cap - 8) + (parity & 1u);
}

struct OutLayout {
    size_t off_keys, off_ids, off_scores, off_cos, off_counts, total;
    OutLayout(uint32_t nq, uint32_t k) {
        const size_t n = static_cast<size_t>(nq) * k;
        off_keys = 0;
        off_ids = n * 8;
        off_scores = off_ids + n * 8;
        off_cos = off_scores + n * 4;
        off_counts = off_cos + n * 4;
        total = off_counts + static_cast<size_t>(nq) * 4;
    }
};

// One search on a context, first half: (blob upload unless everything rides in the kernel parameters) +
// scan launch; the last block writes ids / scores / cosines straight into mapped pinned host memory when
// the batch is small, so there is no device-to-host copy to enqueue.  search_wait is the second half.
// `tq` non-null: the plans come from the store table (caller holds it): only the queries and their scope lists
// travel to the device, `b` is unused.
struct TableQuery {
    const uint32_t *segs, *off;
    uint32_t max_tiles;
};
int search_launch(rf_engine *e, SearchCtx *c, const PlanBlob &b, const int8_t *q_host, uint32_t nq, uint32_t k, bool shared,
                  const TableQuery *tq = nullptr) {
    const auto t0 = std::chrono::steady_clock::now();
    const OutLayout L(nq, k);
    const uint32_t X = pick_blocks(e, nq, tq ? tq->max_tiles : b.max_tiles);
    const size_t flag_off = (L.total + 15) & ~static_cast<size_t>(15);   // [seq word, finished-query counter]
    if (flag_off + 16 > c->m_out.cap) {
        RF_CUDA(c->m_out.reserve(flag_off + 16));
        memset(c->m_out.h, 0, c->m_out.cap);
        c->seq = 0;
    }
    RF_CUDA(c->d_partial.reserve(static_cast<size_t>(nq) * X * k * 8));
    if (static_cast<size_t>(nq) * kSyncBytesPerQuery + 8 > c->d_tickets.cap) {
        RF_CUDA(cudaStreamSynchronize(c->stream));
        RF_CUDA(c->d_tickets.reserve(static_cast<size_t>(nq) * kSyncBytesPerQuery + 8));
        RF_CUDA(cudaMemsetAsync(c->d_tickets.p, 0, c->d_tickets.cap, c->stream));
    }
    ScanArgs a{};
    const bool inline_q = nq == 1 && !tq;
    if (tq) {
        // [queries | scope offsets | scope stores] in one H2D copy
        auto align = [](size_t x) { return (x + 15) & ~static_cast<size_t>(15); };
        const size_t n_segs = tq->off[nq];
        const size_t off_off = align(static_cast<size_t>(nq) * e->dim), off_segs = align(off_off + (static_cast<size_t>(nq) + 1) * 4);
        const size_t total = align(off_segs + std::max<size_t>(n_segs, 1) * 4);
        RF_CUDA(c->h_in.reserve(total));
        RF_CUDA(c->d_in.reserve(total));
        uint8_t *h = static_cast<uint8_t *>(c->h_in.p);
        memcpy(h, q_host, static_cast<size_t>(nq) * e->dim);
        memcpy(h + off_off, tq->off, (static_cast<size_t>(nq) + 1) * 4);
        if (n_segs) memcpy(h + off_segs, tq->segs, n_segs * 4);
        RF_CUDA(cudaMemcpyAsync(c->d_in.p, h, total, cudaMemcpyHostToDevice, c->stream));
        a.F = e->F;
        a.seg = e->seg;
        a.ff = e->ff;
        a.q = static_cast<const int8_t *>(c->d_in.p);
        a.id_base = static_cast<uint32_t>(e->cfg.id_base);
        a.k = k;
        a.debug_ts = e->debug_ts;
        a.dbg_flags = e->dbg_flags;
        fill_table_args(e, a, static_cast<const uint8_t *>(c->d_in.p), off_off, off_segs);
    } else {
        fill_args(e, a, nullptr, b, nullptr, k, shared);
        maybe_inline_plan(a, b, nq, shared);
    }
    if (inline_q) {
        memcpy(a.q_inline, q_host, e->dim);
        a.q = nullptr;
    }
    if (!tq && !(inline_q && a.inline_plan)) {   // queries and/or plans travel as one H2D copy
        RF_CUDA(c->h_in.reserve(b.bytes.size()));
        RF_CUDA(c->d_in.reserve(b.bytes.size()));
        memcpy(c->h_in.p, b.bytes.data(), b.bytes.size());
        RF_CUDA(cudaMemcpyAsync(c->d_in.p, c->h_in.p, b.bytes.size(), cudaMemcpyHostToDevice, c->stream));
        const uint8_t keep_inline = a.inline_plan;
        fill_args(e, a, static_cast<const uint8_t *>(c->d_in.p), b, nullptr, k, shared);
        a.inline_plan = keep_inline;
        if (inline_q) a.q = nullptr;
    }
    // Few queries: results land straight in mapped pinned host memory and the host spins on a
    // completion word (no copy to enqueue, ~1 us after the last store).  Many queries: one PCIe
    // write per result word would dominate, so results stay in HBM and come back as one copy.
    const bool mapped = nq <= 8;
    if (!mapped) {
        RF_CUDA(c->d_out.reserve(L.total));
        RF_CUDA(c->h_out.reserve(L.total));
    }
    uint8_t *d_out = static_cast<uint8_t *>(mapped ? c->m_out.d : c->d_out.p);
    a.partial = static_cast<uint64_t *>(c->d_partial.p);
    set_sync_bufs(a, c->d_tickets, c->launches++);
    a.out_keys = reinterpret_cast<uint64_t *>(d_out + L.off_keys);
    a.out_ids = reinterpret_cast<uint64_t *>(d_out + L.off_ids);
    a.out_scores = reinterpret_cast<int32_t *>(d_out + L.off_scores);
    a.out_cos = reinterpret_cast<float *>(d_out + L.off_cos);
    a.out_counts = reinterpret_cast<uint32_t *>(d_out + L.off_counts);
    SearchCtx::Pending &pd = c->pend;
    pd = SearchCtx::Pending{};
    pd.mapped = mapped;
    pd.nq = nq;
    pd.k = k;
    // The query and the plan come from the kernel parameters or from the copy just enqueued, never from a
    // preceding kernel: the launch may overlap its predecessor's tail.
    if (mapped) {
        // The flag word sits right after the results in the same mapped allocation.
        volatile uint32_t *h_flag = reinterpret_cast<volatile uint32_t *>(static_cast<uint8_t *>(c->m_out.h) + flag_off);
        h_flag[0] = 0;
        a.done_flag = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(c->m_out.d) + flag_off);
        a.done_seq = ++c->seq ? c->seq : ++c->seq;
        pd.flag_off = flag_off;
        pd.done_seq = a.done_seq;
        RF_CUDA(rf::launch_score_topk_scan(a, nq, X, e->scan_variant, e->dim, c->stream, true));
    } else {
        a.done_flag = nullptr;
        RF_CUDA(rf::launch_score_topk_scan(a, nq, X, e->scan_variant, e->dim, c->stream, true));
        RF_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(c->h_out.p) + L.off_ids, d_out + L.off_ids, L.total - L.off_ids,
                                cudaMemcpyDeviceToHost, c->stream));
    }
    e->launches.fetch_add(1, std::memory_order_relaxed);
    pd.active = true;
    pd.t_launched = std::chrono::steady_clock::now();
    if (e->profile) e->prof_ns[1] += std::chrono::duration_cast<std::chrono::nanoseconds>(pd.t_launched - t0).count();
    return RF_OK;
}

// Second half: wait for the search in flight on `c` and hand its results to the caller.
int search_wait(rf_engine *e, SearchCtx *c, uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts,
                int8_t *out_q) {
    SearchCtx::Pending &pd = c->pend;
    if (!pd.active) return fail(RF_EINVAL, "no search in flight on this context");
    pd.active = false;
    const OutLayout L(pd.nq, pd.k);
    const uint8_t *h = nullptr;
    if (pd.mapped) {
        // Spin on the completion word; fall back to the stream's status so a failed launch can
        // never hang the caller.
        volatile uint32_t *h_flag = reinterpret_cast<volatile uint32_t *>(static_cast<uint8_t *>(c->m_out.h) + pd.flag_off);
        for (uint32_t spins = 0; h_flag[0] != pd.done_seq; ++spins) {
            if ((spins & 0x3FFu) == 0x3FFu) {
                const cudaError_t qe = cudaStreamQuery(c->stream);
                if (qe == cudaSuccess) break;
                if (qe != cudaErrorNotReady) return fail(RF_ECUDA, "scan kernel failed: %s", cudaGetErrorString(qe));
            }
        }
        if (h_flag[0] != pd.done_seq) RF_CUDA(cudaStreamSynchronize(c->stream));
        h = static_cast<const uint8_t *>(c->m_out.h);
    } else {
        RF_CUDA(cudaStreamSynchronize(c->stream));
        h = static_cast<const uint8_t *>(c->h_out.p);
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const auto t2 = std::chrono::steady_clock::now();
    const size_t n = static_cast<size_t>(pd.nq) * pd.k;
    memcpy(out_ids, h + L.off_ids, n * 8);
    memcpy(out_scores, h + L.off_scores, n * 4);
    if (out_cos) memcpy(out_cos, h + L.off_cos, n * 4);
    if (out_counts) memcpy(out_counts, h + L.off_counts, static_cast<size_t>(pd.nq) * 4);
    if (out_q && pd.has_q) memcpy(out_q, h + L.total, e->dim);
    if (e->profile) {
        e->prof_ns[2] += std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - pd.t_launched).count();
        e->prof_ns[3] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t2).count();
        e->prof_n += 1;
    }
    return RF_OK;
}

// Reserve `n` contiguous rows (caller holds ingest_mu): first fit among the rows that deletes gave back,
// else the tail of the arena.  *reused tells the caller the rows lie inside the published range, where a
// concurrent scan of a bounding range may read them: it must then write features first and the segment
// words (which un-mask the rows) in a second, later pass.  Published by publish_rows.
int reserve_rows(rf_engine *e, uint64_t n, uint64_t *first, bool *reused) {
    *reused = false;
    if (n) {
        for (size_t i = 0; i < e->free_ext.size(); ++i) {
            Extent &x = e->free_ext[i];
            if (static_cast<uint64_t>(x.hi - x.lo) < n) continue;
            *first = x.lo;
            x.lo += static_cast<uint32_t>(n);
            if (x.lo == x.hi) e->free_ext.erase(e->free_ext.begin() + static_cast<ptrdiff_t>(i));
            e->free_rows -= n;
            *reused = true;
            return RF_OK;
        }
    }
    if (e->n_rows + n > e->cfg.capacity_rows)
        return fail(RF_ECAPACITY, "arena full: %llu + %llu rows > capacity %llu (%llu freed rows, none in a run of %llu)",
                    (unsigned long long)e->n_rows, (unsigned long long)n, (unsigned long long)e->cfg.capacity_rows,
                    (unsigned long long)e->free_rows, (unsigned long long)n);
    *first = e->n_rows;
    return RF_OK;
}

// Give reserved-but-unpublished rows back (an ingest failed after reserve_rows).
void unreserve_rows(rf_engine *e, uint64_t first, uint64_t n, bool reused);

// Rows [lo, hi) are masked (their segment words are tombstones) and belong to no store or document any
// more: add them to the free list (caller holds ingest_mu).
void free_list_add(rf_engine *e, uint32_t lo, uint32_t hi) {
    if (lo >= hi) return;
    auto it = std::lower_bound(e->free_ext.begin(), e->free_ext.end(), lo, [](const Extent &x, uint32_t v) { return x.lo < v; });
    it = e->free_ext.insert(it, Extent{lo, hi});
    e->free_rows += hi - lo;
    if (it + 1 != e->free_ext.end() && (it + 1)->lo <= it->hi) {
        it->hi = std::max(it->hi, (it + 1)->hi);
        it = e->free_ext.erase(it + 1) - 1;
    }
    if (it != e->free_ext.begin() && (it - 1)->hi >= it->lo) {
        (it - 1)->hi = std::max((it - 1)->hi, it->hi);
        e->free_ext.erase(it);
    }
}

void unreserve_rows(rf_engine *e, uint64_t first, uint64_t n, bool reused) {
    if (reused) free_list_add(e, static_cast<uint32_t>(first), static_cast<uint32_t>(first + n));
}

// Returns RF_ENOTFOUND when the store was dropped while the ingest ran (its rows are then given back).
int publish_rows(rf_engine *e, uint32_t store_seg, uint64_t doc_id, bool track_doc, uint64_t first, uint64_t n) {
    std::unique_lock<std::shared_mutex> lk(e->meta_mu);
    if (store_seg >= e->stores.size() || e->stores[store_seg].dropped) return fail(RF_ENOTFOUND, "unknown store segment %u", store_seg);
    extents_append(e->stores[store_seg].ext, static_cast<uint32_t>(first), static_cast<uint32_t>(first + n));
    if (track_doc) {
        Doc &d = e->docs[doc_id];
        d.store = store_seg;
        extents_append(d.ext, static_cast<uint32_t>(first), static_cast<uint32_t>(first + n));
    }
    e->n_rows = std::max<uint64_t>(e->n_rows, first + n);
    e->epoch.fetch_add(1);
    return RF_OK;
}

int check_store(rf_engine *e, uint32_t seg) {
    std::shared_lock<std::shared_mutex> lk(e->meta_mu);
    if (seg >= e->stores.size() || e->stores[seg].dropped) return fail(RF_ENOTFOUND, "unknown store segment %u", seg);
    return RF_OK;
}

std::shared_ptr<StreamState> stream_state_touch(rf_engine *e, const std::shared_ptr<StreamState> &st) {
    st->last_use = ++e->stream_tick;
    return st;
}

uint32_t fnv1a32(const char *s, size_t n) {
    uint32_t h = 0x811C9DC5u;
    for (size_t i = 0; i < n; ++i) { h ^= static_cast<uint8_t>(s[i]); h *= 0x01000193u; }
    return h;
}


// Batched search on the tensor cores (score_topk_gemm.cu): a first pass over a sample of the
// extent gives every query a floor (the k-th largest per-group maximum of the sample is a valid
// lower bound of its final k-th best score), the second pass scores every row with the
// candidate path rare, and a tournament merge combines the per-slice lists.  All on stream `s`.
// Does the batched tensor-core path beat nq passes of the scan kernel over `rows` rows?  Measured on a
// B200 (tools/small_batch_probe.py): scan ~ 6 + 3.65e-5 * nq * rows us, tensor-core path ~ 33 + 4.6e-5 * rows us
// (floor pass + two merges are its fixed cost) -- e.g. 8 queries x 1 M rows: 291 vs 79 us; 3 x 200 k: 24 vs 45 us.
// `rows`: rows of the scope (what the scan kernel reads per query); `span`: rows of the contiguous range
// the tensor-core kernels would score (a scope with several extents is scored over its bounding range,
// the per-row tenant mask drops the foreign rows in between); large batches are compute-bound:
// 1024 queries x 1 M rows take 0.248 ms.
bool gemm_pays(const rf_engine *e, uint32_t nq, uint64_t rows, uint64_t span, uint32_t k, double extra_us = 0.0) {
    if (!e->gemm_enabled || k > static_cast<uint32_t>(rf::kGemmListK) || span < 32768 || nq < 2) return false;
    if (e->gemm_min_queries) return nq >= e->gemm_min_queries && rows * 2 >= span;
    if (e->dim != RF_DIM) {
        // K = 512 / 1024 pair kernel: 256 queries per pair keep it HBM-bound -- one streaming pass over the span per 256
        // queries (~1.6e-4 us per 1024-byte row and pass), against one (dim + 4)-byte-per-row scan per query
        const double w = static_cast<double>(e->dim) / 1024.0;
        const double scan_us = 6.0 + 1.46e-4 * w * nq * static_cast<double>(rows);
        const double gemm_us = 45.0 + extra_us + 1.6e-4 * w * static_cast<double>(span) * ((nq + 255) / 256);
        return scan_us > gemm_us;
    }
    const double scan_us = 6.0 + 3.65e-5 * nq * static_cast<double>(rows);
    const double gemm_us = 33.0 + extra_us + std::max(4.6e-5, 2.1e-7 * nq) * static_cast<double>(span);
    return scan_us > gemm_us;
}
// rows of a sorted, disjoint extent list and the length of its bounding range
void extent_rows(const uint32_t *lo, const uint32_t *hi, uint32_t n, uint64_t &rows, uint32_t &span_lo, uint32_t &span_hi) {
    rows = 0;
    span_lo = n ? lo[0] : 0;
    span_hi = n ? hi[n - 1] : 0;
    for (uint32_t i = 0; i < n; ++i) rows += hi[i] - lo[i];
}

int search_gemm(rf_engine *e, StreamState *dp, const int8_t *q_dev, uint32_t nq, const ScanPlan *plan, uint32_t lo, uint32_t hi,
                uint32_t k, uint64_t *out_keys_dev, cudaStream_t s) {
    const uint32_t rows = hi - lo;
    // More than 256 queries: CTA pairs (tcgen05 cta_group::2, 256-row tiles, full tensor-pipe rate);
    // smaller batches stay on the single-CTA kernel, whose spare accumulators take alternate tiles.
    const bool wide = e->dim != RF_DIM;             // K = 512 / 1024: the pair kernel with the K dimension streamed, 256 queries per pair
    const bool pair = wide || (e->gemm_pair && nq > 256 && e->sm_count >= 2);
    const uint32_t tile_rows = pair ? rf::kGemmPairTileRows : rf::kGemmTileRows;
    const uint32_t q_groups = wide ? (nq + 255) / 256 : (nq + rf::kGemmMT * 128 - 1) / (rf::kGemmMT * 128);   // 512 queries per block or pair
    const uint32_t units = pair ? static_cast<uint32_t>(e->sm_count) / 2 : static_cast<uint32_t>(e->sm_count);
    const uint32_t slices_full = std::max(1u, units / q_groups);
    // rows of the floor pass: the pair kernel's is branch-free (running maxima, no lists) and takes a sample twice the
    // single-CTA kernel's -- the main pass then meets half the candidates (RF_GEMM_SAMPLE sets the latter; measured at
    // 1024 x 1 M: 64 Ki rows 0.219 ms, 128 Ki 0.208, 256 Ki 0.213 -- tools/cfg2_sample_sweep.sh)
    // (the K = 1024 kernel's epilogue idles most of the time: candidates are cheap there, a 32 Ki-row sample is enough)
    uint32_t sample = std::min(rows / 4, wide ? e->gemm_sample / 2 : pair ? e->gemm_sample * 2 : e->gemm_sample) / tile_rows * tile_rows;
    uint32_t n_a = std::max(1u, std::min(slices_full, sample / tile_rows));
    if (e->gemm_slices_a) n_a = std::max(1u, std::min(n_a, e->gemm_slices_a));
    const uint32_t n_b = std::max(1u, std::min(slices_full, (rows + tile_rows - 1) / tile_rows));
    const uint32_t kl = rf::kGemmListK;
    const uint32_t lps = pair ? static_cast<uint32_t>(rf::kGemmPairLists) : rf::gemm_lists_per_slice(nq);
    const size_t keys_bytes = static_cast<size_t>(nq) * kl * 8;
    const size_t lists_bytes = wide ? rf::gemm_wide_lists_bytes(std::max(n_a, n_b), nq)
                               : pair ? rf::gemm_pair_lists_bytes(std::max(n_a, n_b), nq) : rf::gemm_lists_bytes(std::max(n_a, n_b), nq);
    if (dp->gemm_lists.cap < lists_bytes + keys_bytes || dp->gemm_keys_a.cap < keys_bytes ||
        dp->gemm_floors.cap < static_cast<size_t>(nq) * 8) {
        RF_CUDA(cudaStreamSynchronize(s));
        RF_CUDA(dp->gemm_lists.reserve(lists_bytes + keys_bytes));
        RF_CUDA(dp->gemm_keys_a.reserve(keys_bytes));
        RF_CUDA(dp->gemm_floors.reserve(static_cast<size_t>(nq) * 8));
    }
    auto launch = [&](const rf::GemmArgs &g, uint32_t n_slices) {
        return wide ? rf::launch_score_topk_gemm_wide(g, q_dev, e->F, e->cfg.capacity_rows, e->dim, n_slices, s)
               : pair ? rf::launch_score_topk_gemm_pair(g, q_dev, e->F, e->cfg.capacity_rows, n_slices, s)
                      : rf::launch_score_topk_gemm(g, q_dev, e->F, e->cfg.capacity_rows, n_slices, s);
    };
    uint64_t *lists = static_cast<uint64_t *>(dp->gemm_lists.p);
    uint64_t *keys_a = static_cast<uint64_t *>(dp->gemm_keys_a.p);
    uint64_t *floors = static_cast<uint64_t *>(dp->gemm_floors.p);
    rf::GemmArgs g{};
    g.seg = e->seg;
    g.q = q_dev;
    g.n_scope = plan->n_scope;
    for (uint32_t i = 0; i < RF_SCOPE_MAX; ++i) g.scope[i] = plan->scope[i];
    g.nq = nq;
    g.id_base = static_cast<uint32_t>(e->cfg.id_base);
    g.out_lists = lists;
    g.lists_per_slice = pair ? 0u : rf::gemm_lists_per_slice(nq);
    g.debug = nullptr;
    if (const char *dm = getenv("RF_GEMM_DBG")) g.dbg_mode = static_cast<uint32_t>(atoi(dm));
    // pass A: group maxima over a sample -> per-query floors (a lower bound of the k-th best score)
    g.floors = nullptr;
    g.group_max_mode = 1;
    g.row_lo = lo;
    g.row_hi = lo + sample;
    RF_CUDA(launch(g, n_a));
    if (pair) RF_CUDA(rf::launch_kth_largest(reinterpret_cast<const uint32_t *>(lists), rf::gemm_pair_floor_groups(n_a), nq, k, floors, s));
    else RF_CUDA(rf::launch_merge_lists(lists, n_a * lps, nq, kl, kl, keys_a, s, floors, k));   // floors[q] = k-th best group maximum
    // pass B: every row (the sample included: pass A kept maxima, not chunks), floors from pass A
    g.floors = floors;
    g.group_max_mode = 0;
    g.row_lo = lo;
    g.row_hi = hi;
    g.debug = e->debug_ts;   // RF_SCAN_DEBUG=1: cycle counters of the second pass
    RF_CUDA(launch(g, n_b));
    RF_CUDA(rf::launch_merge_lists(lists, n_b * lps, nq, kl, k, out_keys_dev, s));
    e->launches.fetch_add(4, std::memory_order_relaxed);
    return RF_OK;
}

}  // namespace

extern "C" {

static int search_keys_device_impl(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs, uint32_t n_segs,
                                   uint32_t k, uint64_t *out_keys_dev, void *stream, const rf_peer_exchange *px);

const char *rf_strerror(int code) {
    switch (code) {
        case RF_OK: return "ok";
        case RF_EINVAL: return "invalid argument";
        case RF_ENOMEM: return "out of memory";
        case RF_ECUDA: return "CUDA error";
        case RF_ENODEVICE: return "no sm_100 CUDA device";
        case RF_ENOTFOUND: return "not found";
        case RF_EBUSY: return "busy: no search context became free";
        case RF_ECAPACITY: return "arena capacity exceeded";
        default: return "unknown error";
    }
}

const char *rf_last_error(void) { return g_err; }

int rf_build_info(char *buf, size_t n) {
    if (!buf || !n) return RF_EINVAL;
    snprintf(buf, n, "librf_b200 sm_100a cuda-runtime %d dim %u (also 512, 1024) topk_max %u scope_max %u", CUDART_VERSION, RF_DIM,
             RF_TOPK_MAX, RF_SCOPE_MAX);
    return RF_OK;
}

int rf_engine_create(const rf_config *cfg, rf_engine **out) {
    if (!cfg || !out) return fail(RF_EINVAL, "null argument");
    if (cfg->struct_size != sizeof(rf_config)) return fail(RF_EINVAL, "rf_config.struct_size %u != %zu", cfg->struct_size, sizeof(rf_config));
    if (cfg->dim != 256 && cfg->dim != 512 && cfg->dim != 1024) return fail(RF_EINVAL, "dim must be 256, 512 or 1024");
    if (cfg->capacity_rows == 0 || cfg->id_base + cfg->capacity_rows > 0xFFFFFFFEull)
        return fail(RF_EINVAL, "id_base + capacity_rows must be in (0, 2^32 - 2]");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(RF_ENODEVICE, "no CUDA device visible");
    }
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(RF_EINVAL, "device %d out of range (%d visible)", cfg->device, n_dev);
    cudaDeviceProp prop{};
    RF_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(RF_ENODEVICE, "device %d is sm_%d%d; this library holds sm_100a code only", cfg->device, prop.major, prop.minor);
    RF_CUDA(cudaSetDevice(cfg->device));

    rf_engine *e = new (std::nothrow) rf_engine();
    if (!e) return fail(RF_ENOMEM, "host allocation failed");
    e->cfg = *cfg;
    e->dim = cfg->dim;
    e->tile_rows = rf::scan_tile_rows(cfg->dim);
    if (e->cfg.n_contexts == 0) e->cfg.n_contexts = 8;
    e->sm_count = prop.multiProcessorCount;
    if (const char *s = getenv("RF_SCAN_BLOCKS")) e->blocks_override = static_cast<uint32_t>(atoi(s));
    if (const char *s = getenv("RF_STAGE_THREADS")) e->stage_threads = static_cast<uint32_t>(std::max(0, std::min(32, atoi(s))));
    if (const char *s = getenv("RF_STAGE_CHUNK_KB")) {
        const int kb = atoi(s);
        if (kb >= 64 && kb <= 8192 && (kb & (kb - 1)) == 0) e->stage_chunk = static_cast<size_t>(kb) << 10;
    }
    if (const char *s = getenv("RF_GEMM")) e->gemm_enabled = atoi(s) != 0;
    if (const char *s = getenv("RF_STORE_TABLE")) e->table_enabled = atoi(s) != 0;
    if (const char *s = getenv("RF_GEMM_PAIR")) e->gemm_pair = atoi(s) != 0;
    if (const char *s = getenv("RF_GEMM_SAMPLE")) e->gemm_sample = static_cast<uint32_t>(atoi(s));
    if (const char *s = getenv("RF_GEMM_SLICES_A")) e->gemm_slices_a = static_cast<uint32_t>(atoi(s));
    if (const char *s = getenv("RF_GEMM_MIN_QUERIES")) e->gemm_min_queries = static_cast<uint32_t>(atoi(s));
    if (const char *s = getenv("RF_SCAN_VARIANT")) {
        const int v = atoi(s);
        if (v < 0 || v >= rf::kScanVariantCount) {
            delete e;
            return fail(RF_EINVAL, "RF_SCAN_VARIANT=%d out of range", v);
        }
        e->scan_variant = v;
    }
    if (e->dim != RF_DIM) e->scan_variant = rf::kScanVariantTma6x12;   // wider rows: the default ring only

    const uint64_t cap = cfg->capacity_rows;
    cudaError_t ce;
    e->profile = getenv("RF_PROFILE") != nullptr;
    if (const char *s = getenv("RF_SCAN_DBG")) e->dbg_flags = static_cast<uint32_t>(atoi(s));
    if (getenv("RF_SCAN_DEBUG")) {
        e->debug_cap = 4096 * 8;
        if (cudaMalloc(&e->debug_ts, e->debug_cap * 8) != cudaSuccess) e->debug_ts = nullptr;
        else cudaMemset(e->debug_ts, 0, e->debug_cap * 8);
    }
    if ((ce = cudaMalloc(&e->F, cap * e->dim)) != cudaSuccess || (ce = cudaMalloc(&e->seg, cap * 4)) != cudaSuccess ||
        (ce = cudaMalloc(&e->ff, cap * 4)) != cudaSuccess) {
        const int rc = fail(RF_ENOMEM, "cudaMalloc of %llu rows failed: %s", (unsigned long long)cap, cudaGetErrorString(ce));
        cudaGetLastError();
        rf_engine_destroy(e);
        return rc;
    }
    e->hbm_bytes = cap * (e->dim + 8);
    // unwritten rows read as tombstones
    if ((ce = cudaMemset(e->seg, 0xFF, cap * 4)) != cudaSuccess || (ce = cudaStreamCreateWithFlags(&e->ingest_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ingest_idle, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaDeviceSynchronize()) != cudaSuccess) {   // the fill ran on the legacy default stream, which the engine's
                                                            // (non-blocking) streams and the callers' do not order with
        const int rc = fail(RF_ECUDA, "engine init failed: %s", cudaGetErrorString(ce));
        rf_engine_destroy(e);
        return rc;
    }
    for (uint32_t i = 0; i < e->cfg.n_contexts; ++i) {
        SearchCtx *c = new (std::nothrow) SearchCtx();
        if (!c || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            rf_engine_destroy(e);
            return fail(RF_ECUDA, "search context creation failed");
        }
        e->all_ctx.push_back(c);
        e->free_ctx.push_back(c);
    }
    *out = e;
    return RF_OK;
}

int rf_engine_destroy(rf_engine *e) {
    if (!e) return RF_OK;
    if (e->profile && e->prof_n.load()) {
        const double n = static_cast<double>(e->prof_n.load());
        fprintf(stderr, "[rf profile] rf_search x%.0f: plan %.2f us, prepare+launch %.2f us, wait %.2f us, copy-out %.2f us\n", n,
                e->prof_ns[0].load() / n / 1e3, e->prof_ns[1].load() / n / 1e3, e->prof_ns[2].load() / n / 1e3,
                e->prof_ns[3].load() / n / 1e3);
    }
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (SearchCtx *c : e->all_ctx) {
        if (c->stream) cudaStreamDestroy(c->stream);
        c->h_in.release(); c->h_out.release(); c->m_out.release();
        c->d_in.release(); c->d_out.release(); c->d_partial.release(); c->d_tickets.release();
        delete c;
    }
    for (auto &kv : e->stream_states) kv.second->release();
    e->stream_states.clear();
    e->tbl.entries.release(); e->tbl.lo.release(); e->tbl.hi.release();
    if (e->ingest_stream) cudaStreamDestroy(e->ingest_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->ingest_idle) cudaEventDestroy(e->ingest_idle);
    for (cudaEvent_t ev : e->ingest_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->copy_ev)
        if (ev) cudaEventDestroy(ev);
    copy_pool_destroy(e->copy_pool);
    e->sc_text.release(); e->sc_state.release(); e->sc_bucket.release(); e->sc_cstart.release();
    e->sc_end.release(); e->sc_ctl.release(); e->sc_spans.release(); e->sc_deferred.release();
    e->sc_stage.release(); e->sc_ctl_host.release();
    if (e->debug_ts) cudaFree(e->debug_ts);
    if (e->zipf_bucket) cudaFree(e->zipf_bucket);
    if (e->reader) {   // the arena belongs to another process: unmap it
        if (e->F) cudaIpcCloseMemHandle(e->F);
        if (e->seg) cudaIpcCloseMemHandle(e->seg);
        if (e->ff) cudaIpcCloseMemHandle(e->ff);
    } else {
        if (e->F) cudaFree(e->F);
        if (e->seg) cudaFree(e->seg);
        if (e->ff) cudaFree(e->ff);
    }
    delete e;
    return RF_OK;
}

int rf_debug_timestamps(rf_engine *e, uint64_t *out, uint64_t n_words, int clear) {
    if (!e || !out) return fail(RF_EINVAL, "null argument");
    if (!e->debug_ts) return fail(RF_EINVAL, "engine was not created with RF_SCAN_DEBUG=1");
    RF_CUDA(cudaSetDevice(e->cfg.device));
    RF_CUDA(cudaDeviceSynchronize());
    const size_t n = std::min<size_t>(n_words, e->debug_cap);
    RF_CUDA(cudaMemcpy(out, e->debug_ts, n * 8, cudaMemcpyDeviceToHost));
    if (clear) RF_CUDA(cudaMemset(e->debug_ts, 0, e->debug_cap * 8));
    return RF_OK;
}

int rf_engine_stats(rf_engine *e, rf_stats *out) {
    if (!e || !out) return fail(RF_EINVAL, "null argument");
    std::shared_lock<std::shared_mutex> lk(e->meta_mu);
    out->n_rows = e->n_rows;
    out->capacity_rows = e->cfg.capacity_rows;
    uint64_t live = 0;
    for (const Store &s : e->stores) live += s.dropped ? 0 : 1;
    out->n_stores = live;
    out->n_docs = e->docs.size();
    out->hbm_bytes = e->hbm_bytes;
    out->searches = e->searches.load();
    out->kernel_launches = e->launches.load();
    out->free_rows = e->free_rows;   // (read without ingest_mu: a statistic)
    out->ingest_bytes = e->ingest_bytes.load();
    out->ingest_kernel_ns = e->ingest_kernel_ns.load();
    return RF_OK;
}

int rf_store_open(rf_engine *e, const char *fs_name, uint32_t *store_seg) {
    if (!e || !fs_name || !store_seg) return fail(RF_EINVAL, "null argument");
    if (e->reader) return rf_store_lookup(e, fs_name, store_seg);   // a reader cannot create stores; an existing one is found
    std::unique_lock<std::shared_mutex> lk(e->meta_mu);
    auto it = e->store_by_name.find(fs_name);
    if (it != e->store_by_name.end()) { *store_seg = it->second; return RF_OK; }
    if (e->stores.size() >= 0xFFFFFFFEull) return fail(RF_ENOMEM, "store segment space exhausted");
    const uint32_t seg = static_cast<uint32_t>(e->stores.size());
    e->stores.emplace_back();
    e->stores.back().name = fs_name;
    e->store_by_name.emplace(fs_name, seg);
    *store_seg = seg;
    return RF_OK;
}

int rf_store_lookup(rf_engine *e, const char *fs_name, uint32_t *store_seg) {
    if (!e || !fs_name || !store_seg) return fail(RF_EINVAL, "null argument");
    std::shared_lock<std::shared_mutex> lk(e->meta_mu);
    auto it = e->store_by_name.find(fs_name);
    if (it == e->store_by_name.end()) return fail(RF_ENOTFOUND, "unknown store '%s'", fs_name);
    *store_seg = it->second;
    return RF_OK;
}

static int tombstone_extents(rf_engine *e, const std::vector<Extent> &ext) {
    RF_CUDA(cudaSetDevice(e->cfg.device));
    for (const Extent &x : ext)
        RF_CUDA(cudaMemsetAsync(e->seg + x.lo, 0xFF, static_cast<size_t>(x.hi - x.lo) * 4, e->ingest_stream));
    RF_CUDA(cudaStreamSynchronize(e->ingest_stream));
    e->tomb_gen.fetch_add(1);
    for (const Extent &x : ext) free_list_add(e, x.lo, x.hi);   // masked now: the rows can be written again
    return RF_OK;
}

int rf_store_drop(rf_engine *e, uint32_t store_seg) {
    if (!e) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process drops stores");
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    std::vector<Extent> ext;
    {
        std::unique_lock<std::shared_mutex> lk(e->meta_mu);
        if (store_seg >= e->stores.size() || e->stores[store_seg].dropped) return fail(RF_ENOTFOUND, "unknown store segment %u", store_seg);
        Store &s = e->stores[store_seg];
        ext.swap(s.ext);
        s.dropped = true;
        e->store_by_name.erase(s.name);
        for (auto it = e->docs.begin(); it != e->docs.end();) it = (it->second.store == store_seg) ? e->docs.erase(it) : std::next(it);
        e->epoch.fetch_add(1);
    }
    return tombstone_extents(e, ext);
}

int rf_doc_tombstone(rf_engine *e, uint64_t doc_id) {
    if (!e) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process deletes documents");
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    std::vector<Extent> ext;
    {
        std::unique_lock<std::shared_mutex> lk(e->meta_mu);
        auto it = e->docs.find(doc_id);
        if (it == e->docs.end()) return fail(RF_ENOTFOUND, "unknown document %llu", (unsigned long long)doc_id);
        ext.swap(it->second.ext);
        const uint32_t st = it->second.store;
        e->docs.erase(it);
        // the rows leave the store's extents too: once masked they may be handed to another store's document
        if (st < e->stores.size())
            for (const Extent &x : ext) extents_subtract(e->stores[st].ext, x.lo, x.hi);
        e->epoch.fetch_add(1);
    }
    return tombstone_extents(e, ext);
}

// Text of one document -> device (e->sc_text) with the tokeniser launched behind it, chunk by chunk.
// Caller holds ingest_mu.  On return every tokenise launch is enqueued on the ingest stream.
static int copy_and_tokenize(rf_engine *e, const uint8_t *utf8, size_t n, rf::TokenizeArgs &t) {
    cudaStream_t s = e->ingest_stream;
    const uint32_t n_blocks = t.n_blocks;
    int launches = 0;
    // where does the text live?  (a plain malloc'ed buffer reports "unregistered")
    cudaPointerAttributes attr{};
    cudaMemoryType where = cudaMemoryTypeUnregistered;
    if (n && cudaPointerGetAttributes(&attr, utf8) == cudaSuccess) where = attr.type;
    else cudaGetLastError();
    uint8_t *d_text = static_cast<uint8_t *>(e->sc_text.p);
    auto blocks_upto = [&](size_t bytes) { return static_cast<uint32_t>((bytes + rf::kFeatBlockBytes - 1) / rf::kFeatBlockBytes); };
    if (n == 0) return RF_OK;
    if (where == cudaMemoryTypeDevice || where == cudaMemoryTypeManaged) {
        // already in HBM: one device copy into the padded, aligned scratch, one launch
        RF_CUDA(cudaMemcpyAsync(d_text, utf8, n, cudaMemcpyDeviceToDevice, s));
        t.avail_end = n;
        RF_CUDA(cudaEventRecord(e->ingest_ev[0], s));
        RF_CUDA(rf::launch_tokenize(t, 0, n_blocks, s));
        e->launches.fetch_add(1);
        return RF_OK;
    }
    const bool pinned_src = where == cudaMemoryTypeHost;
    if (n < kStageMinBytes) {
        const uint8_t *src = utf8;
        if (!pinned_src) {
            RF_CUDA(e->sc_stage.reserve(kStageWindowBytes));
            rf::stage_copy(e->sc_stage.p, utf8, n);
            src = static_cast<const uint8_t *>(e->sc_stage.p);
        }
        RF_CUDA(cudaMemcpyAsync(d_text, src, n, cudaMemcpyHostToDevice, s));
        t.avail_end = n;
        RF_CUDA(cudaEventRecord(e->ingest_ev[0], s));
        RF_CUDA(rf::launch_tokenize(t, 0, n_blocks, s));
        e->launches.fetch_add(1);
        return RF_OK;
    }
    // ---- chunked: DMA of chunk c, then the tokeniser over the blocks of chunk c - 1 (whose look-ahead
    // bytes -- a block's 4-byte halo, a token running on -- are in chunk c)
    const size_t chunk = e->stage_chunk;
    CopyPool *pool = nullptr;
    if (!pinned_src) {
        RF_CUDA(e->sc_stage.reserve(kStageWindowBytes));
        if (!e->copy_pool) {
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            e->copy_pool = copy_pool_create(std::min(e->stage_threads, hw > 1 ? hw - 1 : 0u));
            if (!e->copy_pool) return fail(RF_ENOMEM, "host allocation failed");
        }
        pool = e->copy_pool;
    }
    // Copies go to their own stream so the DMA engine never waits for a tokenise kernel: chunk c's arrival is
    // an event the compute stream waits on before it tokenises chunk c - 1.  The memsets of the control words
    // (compute stream, already enqueued) and the copies touch different buffers; the copy stream only has to
    // wait until the PREVIOUS document's kernels are done with the text scratch (they are: every ingest ends
    // with a synchronise of the compute stream).
    cudaStream_t cs = e->copy_stream;
    for (cudaEvent_t &ev : e->copy_ev)
        if (!ev) RF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    int rc = RF_OK;
    uint32_t blocks_done = 0, g = 0;      // g: chunks enqueued so far (over all windows)
    for (size_t w_off = 0; w_off < n && rc == RF_OK; w_off += kStageWindowBytes) {
        const size_t w_len = std::min(kStageWindowBytes, n - w_off);
        const uint32_t n_chunks = static_cast<uint32_t>((w_len + chunk - 1) / chunk);
        if (pool) {
            if (w_off) RF_CUDA(cudaStreamSynchronize(cs));    // the previous window's copies have left the staging buffer
            pool->begin(utf8 + w_off, static_cast<uint8_t *>(e->sc_stage.p), w_len, chunk);
        }
        for (uint32_t c = 0; c < n_chunks && rc == RF_OK; ++c) {
            const size_t off = static_cast<size_t>(c) * chunk, len = std::min(chunk, w_len - off);
            const uint8_t *src = utf8 + w_off + off;
            if (pool) {
                // stage chunks ourselves while the one we need is not there yet (the helpers may still be waking up)
                while (!pool->staged[c].load(std::memory_order_acquire)) pool->stage_one();
                src = static_cast<const uint8_t *>(e->sc_stage.p) + off;
            }
            // (an event may be re-recorded freely: a stream wait binds to the record that preceded it)
            cudaError_t ce = cudaMemcpyAsync(d_text + w_off + off, src, len, cudaMemcpyHostToDevice, cs);
            if (ce == cudaSuccess) ce = cudaEventRecord(e->copy_ev[g % 16], cs);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, e->copy_ev[g % 16], 0);
            ++g;
            if (ce == cudaSuccess && w_off + off > 0) {
                if (blocks_done == 0) ce = cudaEventRecord(e->ingest_ev[0], s);   // (pipelined: the span to ingest_ev[1] includes waiting for copies)
                t.avail_end = w_off + off + len;
                const uint32_t upto = blocks_upto(w_off + off);        // blocks that START before this chunk
                if (ce == cudaSuccess) ce = rf::launch_tokenize(t, blocks_done, upto - blocks_done, s);
                blocks_done = upto;
                ++launches;
            }
            if (ce != cudaSuccess) rc = fail(RF_ECUDA, "ingest pipeline failed: %s", cudaGetErrorString(ce));
        }
        if (pool) pool->end();
    }
    if (rc) return rc;
    t.avail_end = n;
    if (blocks_done == 0) RF_CUDA(cudaEventRecord(e->ingest_ev[0], s));
    RF_CUDA(rf::launch_tokenize(t, blocks_done, n_blocks - blocks_done, s));
    e->launches.fetch_add(launches + 1);
    return RF_OK;
}

int rf_ingest_text(rf_engine *e, uint32_t store_seg, uint64_t doc_id, const uint8_t *utf8, size_t n, uint64_t *first_chunk,
                   uint32_t *n_chunks, int64_t *spans, uint32_t max_spans) {
    if (!e || (!utf8 && n)) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process ingests");
    if (n > 0xFFFFFFF0ull) return fail(RF_EINVAL, "document larger than 4 GiB");
    int rc = check_store(e, store_seg);
    if (rc) return rc;
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    const auto t_start = std::chrono::steady_clock::now();
    RF_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->ingest_stream;
    const uint32_t n_blocks = static_cast<uint32_t>((n + rf::kFeatBlockBytes - 1) / rf::kFeatBlockBytes);
    const size_t max_tokens = n / 2 + 1;
    RF_CUDA(e->sc_text.reserve(n + 64));
    RF_CUDA(e->sc_state.reserve((static_cast<size_t>(n_blocks) + 1) * 8));
    RF_CUDA(e->sc_bucket.reserve(max_tokens * 2));
    RF_CUDA(e->sc_end.reserve(max_tokens * 4));
    RF_CUDA(e->sc_cstart.reserve((max_tokens / 112 + 2) * 4));
    RF_CUDA(e->sc_ctl.reserve(rf::kCtlWords * 4));
    RF_CUDA(e->sc_deferred.reserve(2 * rf::kMaxDeferred * 4));
    RF_CUDA(e->sc_ctl_host.reserve(rf::kCtlWords * 4));
    rf::TokenizeArgs t{};
    t.text = static_cast<const uint8_t *>(e->sc_text.p);
    t.n = n;
    t.avail_end = n;
    t.n_blocks = n_blocks;
    t.state = static_cast<uint64_t *>(e->sc_state.p);
    t.ctl = static_cast<uint32_t *>(e->sc_ctl.p);
    t.tok_bucket = static_cast<uint16_t *>(e->sc_bucket.p);
    t.dim_mask = e->dim - 1;
    t.tok_end = static_cast<uint32_t *>(e->sc_end.p);
    t.chunk_start = static_cast<uint32_t *>(e->sc_cstart.p);
    t.deferred = static_cast<uint32_t *>(e->sc_deferred.p);
    RF_CUDA(cudaMemsetAsync(t.ctl, 0, rf::kCtlWords * 4, s));
    if (n_blocks) RF_CUDA(cudaMemsetAsync(t.state, 0, static_cast<size_t>(n_blocks) * 8, s));
    for (cudaEvent_t &ev : e->ingest_ev)
        if (!ev) RF_CUDA(cudaEventCreate(&ev));
    if ((rc = copy_and_tokenize(e, utf8, n, t))) return rc;
    if (n) RF_CUDA(cudaEventRecord(e->ingest_ev[1], s));
    volatile uint32_t *h_ctl = static_cast<volatile uint32_t *>(e->sc_ctl_host.p);
    RF_CUDA(cudaMemcpyAsync(e->sc_ctl_host.p, t.ctl, rf::kCtlWords * 4, cudaMemcpyDeviceToHost, s));
    RF_CUDA(cudaStreamSynchronize(s));
    const uint32_t n_tokens = h_ctl[rf::kCtlTokens], n_deferred = h_ctl[rf::kCtlDeferred];
    if (h_ctl[rf::kCtlStalled]) return fail(RF_ECUDA, "internal: the tokeniser's token-count exchange timed out");
    int launches = 0;
    if (n_deferred) {   // tokens longer than a whole copy chunk: hash them now that every byte is here
        if (n_deferred > rf::kMaxDeferred) return fail(RF_EINVAL, "internal: %u parked tokens (max %u)", n_deferred, rf::kMaxDeferred);
        t.avail_end = n;
        RF_CUDA(rf::launch_hash_deferred(t, n_deferred, s));
        ++launches;
    }
    const uint32_t nc = n_tokens == 0 ? 0 : 1 + ((n_tokens > 128 ? n_tokens - 128 : 0) + 111) / 112;
    uint64_t first = 0;
    bool reused = false;
    if ((rc = reserve_rows(e, nc, &first, &reused))) return rc;
    if (nc) {
        RF_CUDA(e->sc_spans.reserve(static_cast<size_t>(nc) * 16));
        int64_t *d_spans = static_cast<int64_t *>(e->sc_spans.p);
        // rows taken from the free list lie inside ranges a concurrent search may be scanning (masked):
        // features and norms first, the segment words that un-mask them only after those are complete
        RF_CUDA(cudaEventRecord(e->ingest_ev[2], s));
        RF_CUDA(rf::launch_rows_from_tokens(t, n_tokens, nc, e->F + first * e->dim, e->ff + first, reused ? nullptr : e->seg + first,
                                            store_seg, d_spans, s));
        RF_CUDA(cudaEventRecord(e->ingest_ev[3], s));
        ++launches;
        const uint32_t ns = spans ? std::min(nc, max_spans) : 0;
        if (ns) RF_CUDA(cudaMemcpyAsync(spans, d_spans, static_cast<size_t>(ns) * 16, cudaMemcpyDeviceToHost, s));
        RF_CUDA(cudaStreamSynchronize(s));
        float ms_tok = 0.0f, ms_rows = 0.0f;
        if (cudaEventElapsedTime(&ms_tok, e->ingest_ev[0], e->ingest_ev[1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms_rows, e->ingest_ev[2], e->ingest_ev[3]) == cudaSuccess)
            e->ingest_kernel_ns.fetch_add(static_cast<uint64_t>((static_cast<double>(ms_tok) + ms_rows) * 1e6), std::memory_order_relaxed);
        else cudaGetLastError();
        if (reused) {
            RF_CUDA(rf::launch_fill_u32(e->seg + first, nc, store_seg, s));
            ++launches;
            RF_CUDA(cudaStreamSynchronize(s));
        }
    }
    e->launches.fetch_add(launches);
    if ((rc = publish_rows(e, store_seg, doc_id, true, first, nc))) {   // the store was dropped meanwhile
        if (nc) {
            cudaMemsetAsync(e->seg + first, 0xFF, static_cast<size_t>(nc) * 4, s);
            cudaStreamSynchronize(s);
            unreserve_rows(e, first, nc, reused);
        }
        return rc;
    }
    if (first_chunk) *first_chunk = e->cfg.id_base + first;
    if (n_chunks) *n_chunks = nc;
    e->ingest_bytes.fetch_add(n, std::memory_order_relaxed);
    e->ingest_ns.fetch_add(static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_start).count()),
                           std::memory_order_relaxed);
    return RF_OK;
}

int rf_host_alloc(size_t n, void **out) {
    if (!out || n == 0) return fail(RF_EINVAL, "null argument");
    const cudaError_t ce = cudaHostAlloc(out, n, cudaHostAllocPortable);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return fail(RF_ENOMEM, "cudaHostAlloc of %zu bytes failed: %s", n, cudaGetErrorString(ce));
    }
    return RF_OK;
}

int rf_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return fail(RF_EINVAL, "not a buffer from rf_host_alloc");
    }
    return RF_OK;
}

int rf_ingest_features(rf_engine *e, uint32_t store_seg, uint64_t doc_id, const int8_t *rows, uint64_t n_rows,
                       int rows_on_device, uint64_t *first_chunk) {
    if (!e || (!rows && n_rows)) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process ingests");
    int rc = check_store(e, store_seg);
    if (rc) return rc;
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    RF_CUDA(cudaSetDevice(e->cfg.device));
    uint64_t first = 0;
    bool reused = false;
    if ((rc = reserve_rows(e, n_rows, &first, &reused))) return rc;
    cudaStream_t s = e->ingest_stream;
    if (n_rows) {
        RF_CUDA(cudaMemcpyAsync(e->F + first * e->dim, rows, n_rows * e->dim,
                                rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        RF_CUDA(rf::launch_row_meta(e->F + first * e->dim, n_rows, e->dim, e->ff + first, reused ? nullptr : e->seg + first, store_seg, s));
        e->launches.fetch_add(1);
        RF_CUDA(cudaStreamSynchronize(s));
        if (reused) {   // see rf_ingest_text
            RF_CUDA(rf::launch_fill_u32(e->seg + first, n_rows, store_seg, s));
            e->launches.fetch_add(1);
            RF_CUDA(cudaStreamSynchronize(s));
        }
    }
    if ((rc = publish_rows(e, store_seg, doc_id, true, first, n_rows))) {
        if (n_rows) {
            cudaMemsetAsync(e->seg + first, 0xFF, static_cast<size_t>(n_rows) * 4, s);
            cudaStreamSynchronize(s);
            unreserve_rows(e, first, n_rows, reused);
        }
        return rc;
    }
    if (first_chunk) *first_chunk = e->cfg.id_base + first;
    return RF_OK;
}

int rf_ingest_synthetic(rf_engine *e, uint32_t first_seg, uint64_t rows_per_store, uint64_t seed, uint64_t start_counter,
                        uint64_t n_rows, const uint16_t *zipf_vocab, uint64_t *first_chunk) {
    if (!e || !zipf_vocab) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process ingests");
    const uint64_t n_stores = rows_per_store ? (n_rows + rows_per_store - 1) / rows_per_store : 1;
    {
        std::shared_lock<std::shared_mutex> lk(e->meta_mu);
        if (first_seg + n_stores > e->stores.size()) return fail(RF_ENOTFOUND, "store segments %u..%llu are not all open", first_seg, (unsigned long long)(first_seg + n_stores - 1));
    }
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    RF_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->ingest_stream;
    if (!e->zipf_bucket) {  // bucket of the decimal-ASCII token of each table entry (table prep, 64 KB)
        std::vector<uint16_t> zb(65536);
        char tmp[8];
        for (int r = 0; r < 65536; ++r) {
            const int len = snprintf(tmp, sizeof tmp, "%u", static_cast<unsigned>(zipf_vocab[r]));
            zb[r] = static_cast<uint16_t>(fnv1a32(tmp, static_cast<size_t>(len)) & (e->dim - 1));
        }
        RF_CUDA(cudaMalloc(&e->zipf_bucket, 65536 * 2));
        RF_CUDA(cudaMemcpy(e->zipf_bucket, zb.data(), 65536 * 2, cudaMemcpyHostToDevice));
        RF_CUDA(cudaDeviceSynchronize());       // (staged copy: on the device before the generator reads it on the ingest stream)
    }
    // the generator appends at the tail only (row content is tied to its counter; a measurement input)
    if (e->n_rows + n_rows > e->cfg.capacity_rows)
        return fail(RF_ECAPACITY, "arena full: %llu + %llu rows > capacity %llu", (unsigned long long)e->n_rows,
                    (unsigned long long)n_rows, (unsigned long long)e->cfg.capacity_rows);
    const uint64_t first = e->n_rows;
    RF_CUDA(rf::launch_synth_rows(seed, start_counter, n_rows, e->zipf_bucket, e->dim, e->F + first * e->dim, e->ff + first,
                                  e->seg + first, first_seg, rows_per_store, s));
    e->launches.fetch_add(1);
    RF_CUDA(cudaStreamSynchronize(s));
    {
        std::unique_lock<std::shared_mutex> lk(e->meta_mu);
        for (uint64_t i = 0; i < n_stores; ++i) {
            const uint64_t lo = first + i * (rows_per_store ? rows_per_store : n_rows);
            const uint64_t hi = std::min(first + n_rows, lo + (rows_per_store ? rows_per_store : n_rows));
            extents_append(e->stores[first_seg + i].ext, static_cast<uint32_t>(lo), static_cast<uint32_t>(hi));
        }
        e->n_rows = first + n_rows;
        e->epoch.fetch_add(1);
    }
    if (first_chunk) *first_chunk = e->cfg.id_base + first;
    return RF_OK;
}

namespace {
// Snapshot file, version 2: header, store table, document table, free list, the three row arrays, and a
// trailer {magic "RFB2END2", checksum} -- a 64-bit multiply-xorshift hash of every byte before it, so a
// truncated or corrupted file is refused at load.  Written to "<path>.tmp", flushed, fsync'ed and renamed
// over <path>: a crash or a full disk mid-save leaves the previous snapshot intact.
struct SnapHeader {
    char magic[8];
    uint32_t dim, version;
    uint64_t n_rows, id_base, n_stores, n_docs, n_free;
};
struct SnapTrailer {
    char magic[8];
    uint64_t checksum;
};
struct Hasher {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    void feed(const void *p, size_t n) {
        const uint8_t *b = static_cast<const uint8_t *>(p);
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            memcpy(&w, b + i, 8);
            h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
            h ^= h >> 29;
        }
        if (i < n) {
            uint64_t w = 0;
            memcpy(&w, b + i, n - i);
            h = (h ^ w ^ (static_cast<uint64_t>(n - i) << 56)) * 0x94D049BB133111EBull;
            h ^= h >> 31;
        }
    }
};
struct SnapWriter {
    FILE *f = nullptr;
    Hasher hs;
    bool ok = true;
    void put(const void *p, size_t n) {
        if (!ok || n == 0) return;
        ok = fwrite(p, 1, n, f) == n;
        hs.feed(p, n);
    }
};
struct SnapReader {
    FILE *f = nullptr;
    Hasher hs;
    bool get(void *p, size_t n) {
        if (n == 0) return true;
        if (fread(p, 1, n, f) != n) return false;
        hs.feed(p, n);
        return true;
    }
};
struct FileCloser {
    FILE *f;
    ~FileCloser() { if (f) fclose(f); }
};
constexpr size_t kSnapChunk = 64u << 20;
bool extents_valid(const std::vector<Extent> &v, uint64_t n_rows) {
    for (const Extent &x : v)
        if (x.lo > x.hi || x.hi > n_rows) return false;
    return true;
}
}  // namespace

int rf_snapshot_save(rf_engine *e, const char *path) {
    if (!e || !path) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process writes snapshots");
    std::lock_guard<std::mutex> ing(e->ingest_mu);          // no appends / tombstones while we copy
    std::shared_lock<std::shared_mutex> lk(e->meta_mu);
    RF_CUDA(cudaSetDevice(e->cfg.device));
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(RF_EINVAL, "cannot open %s for writing", tmp.c_str());
    SnapWriter w;
    w.f = f;
    auto abandon = [&](const char *what) {
        fclose(f);
        unlink(tmp.c_str());
        return fail(RF_EINVAL, "%s %s failed; the previous snapshot (if any) is untouched", what, tmp.c_str());
    };
    SnapHeader h{};
    memcpy(h.magic, "RFB2SNP2", 8);
    h.dim = e->dim;
    h.version = 2;
    h.n_rows = e->n_rows;
    h.id_base = e->cfg.id_base;
    h.n_stores = e->stores.size();
    h.n_docs = e->docs.size();
    h.n_free = e->free_ext.size();
    w.put(&h, sizeof h);
    for (const Store &s : e->stores) {
        const uint32_t len = static_cast<uint32_t>(s.name.size()), dropped = s.dropped ? 1u : 0u, n_ext = static_cast<uint32_t>(s.ext.size());
        w.put(&len, 4); w.put(s.name.data(), len); w.put(&dropped, 4); w.put(&n_ext, 4); w.put(s.ext.data(), n_ext * sizeof(Extent));
    }
    for (const auto &kv : e->docs) {
        const uint32_t n_ext = static_cast<uint32_t>(kv.second.ext.size());
        w.put(&kv.first, 8); w.put(&kv.second.store, 4); w.put(&n_ext, 4); w.put(kv.second.ext.data(), n_ext * sizeof(Extent));
    }
    w.put(e->free_ext.data(), e->free_ext.size() * sizeof(Extent));
    if (!w.ok) return abandon("write to");
    std::vector<uint8_t> buf(std::min<size_t>(kSnapChunk, std::max<size_t>(e->n_rows * e->dim, 1)));
    const struct { const void *base; size_t elt; } arrays[3] = {{e->F, e->dim}, {e->seg, 4}, {e->ff, 4}};
    for (const auto &arr : arrays) {
        const size_t total = e->n_rows * arr.elt;
        for (size_t off = 0; off < total; off += buf.size()) {
            const size_t n = std::min(buf.size(), total - off);
            if (cudaMemcpy(buf.data(), static_cast<const uint8_t *>(arr.base) + off, n, cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                return abandon("device read for");
            }
            w.put(buf.data(), n);
            if (!w.ok) return abandon("write to");
        }
    }
    SnapTrailer t{};
    memcpy(t.magic, "RFB2END2", 8);
    t.checksum = w.hs.h;
    if (fwrite(&t, 1, sizeof t, f) != sizeof t || fflush(f) != 0 || fsync(fileno(f)) != 0) return abandon("flush of");
    if (fclose(f) != 0) { unlink(tmp.c_str()); return fail(RF_EINVAL, "close of %s failed", tmp.c_str()); }
    if (rename(tmp.c_str(), path) != 0) { unlink(tmp.c_str()); return fail(RF_EINVAL, "rename of %s over %s failed", tmp.c_str(), path); }
    return RF_OK;
}

int rf_snapshot_load(rf_engine *e, const char *path) {
    if (!e || !path) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "read-only attachment: the owning process loads snapshots");
    std::lock_guard<std::mutex> ing(e->ingest_mu);
    std::unique_lock<std::shared_mutex> lk(e->meta_mu);
    if (e->n_rows || !e->stores.empty()) return fail(RF_EINVAL, "snapshots load into an empty engine");
    RF_CUDA(cudaSetDevice(e->cfg.device));
    FILE *f = fopen(path, "rb");
    if (!f) return fail(RF_ENOTFOUND, "cannot open %s", path);
    FileCloser fc{f};
    SnapReader r;
    r.f = f;
    SnapHeader h{};
    if (!r.get(&h, sizeof h) || memcmp(h.magic, "RFB2SNP2", 8) != 0 || h.version != 2)
        return fail(RF_EINVAL, "%s is not an RF-1 snapshot (version 2)", path);
    if (h.dim != e->dim) return fail(RF_EINVAL, "snapshot rows have %u features, the engine was created with dim %u", h.dim, e->dim);
    if (h.n_rows > e->cfg.capacity_rows) return fail(RF_ECAPACITY, "snapshot holds %llu rows, engine capacity is %llu", (unsigned long long)h.n_rows, (unsigned long long)e->cfg.capacity_rows);
    if (h.id_base != e->cfg.id_base) return fail(RF_EINVAL, "snapshot id_base %llu != engine id_base %llu", (unsigned long long)h.id_base, (unsigned long long)e->cfg.id_base);
    if (h.n_stores > (1ull << 32) || h.n_docs > (1ull << 40) || h.n_free > h.n_rows + 1) return fail(RF_EINVAL, "%s: implausible table sizes", path);
    std::vector<Store> stores(h.n_stores);
    bool ok = true;
    for (Store &s : stores) {
        uint32_t len = 0, dropped = 0, n_ext = 0;
        ok = ok && r.get(&len, 4) && len < (1u << 20);
        if (!ok) break;
        s.name.resize(len);
        ok = r.get(&s.name[0], len) && r.get(&dropped, 4) && r.get(&n_ext, 4) && n_ext <= h.n_rows + 1;
        if (!ok) break;
        s.dropped = dropped != 0;
        s.ext.resize(n_ext);
        ok = r.get(s.ext.data(), n_ext * sizeof(Extent)) && extents_valid(s.ext, h.n_rows);
    }
    std::unordered_map<uint64_t, Doc> docs;
    for (uint64_t i = 0; ok && i < h.n_docs; ++i) {
        uint64_t id = 0;
        uint32_t store = 0, n_ext = 0;
        ok = r.get(&id, 8) && r.get(&store, 4) && r.get(&n_ext, 4) && n_ext <= h.n_rows + 1 && store < h.n_stores;
        if (!ok) break;
        Doc d;
        d.store = store;
        d.ext.resize(n_ext);
        ok = r.get(d.ext.data(), n_ext * sizeof(Extent)) && extents_valid(d.ext, h.n_rows);
        docs.emplace(id, std::move(d));
    }
    std::vector<Extent> free_ext(ok ? h.n_free : 0);
    ok = ok && r.get(free_ext.data(), free_ext.size() * sizeof(Extent)) && extents_valid(free_ext, h.n_rows);
    if (!ok) return fail(RF_EINVAL, "%s is truncated or corrupt (tables)", path);
    // The arrays go straight into the (empty) arena; nothing is published until the checksum has matched,
    // and a failed load masks whatever it wrote.
    auto unwind = [&](const char *what) {
        cudaMemset(e->seg, 0xFF, static_cast<size_t>(h.n_rows) * 4);
        cudaDeviceSynchronize();
        return fail(RF_EINVAL, "%s is %s; nothing was loaded", path, what);
    };
    std::vector<uint8_t> buf(std::min<size_t>(kSnapChunk, std::max<size_t>(h.n_rows * e->dim, 1)));
    const struct { void *base; size_t elt; } arrays[3] = {{e->F, e->dim}, {e->seg, 4}, {e->ff, 4}};
    for (const auto &arr : arrays) {
        const size_t total = h.n_rows * arr.elt;
        for (size_t off = 0; off < total; off += buf.size()) {
            const size_t n = std::min(buf.size(), total - off);
            if (!r.get(buf.data(), n)) return unwind("truncated");
            RF_CUDA(cudaMemcpy(static_cast<uint8_t *>(arr.base) + off, buf.data(), n, cudaMemcpyHostToDevice));
        }
    }
    SnapTrailer t{};
    if (fread(&t, 1, sizeof t, f) != sizeof t || memcmp(t.magic, "RFB2END2", 8) != 0) return unwind("truncated (no trailer)");
    if (t.checksum != r.hs.h) return unwind("corrupt (checksum mismatch)");
    RF_CUDA(cudaDeviceSynchronize());           // the staged copies are on the device before any search can see the rows
    e->stores.swap(stores);
    e->store_by_name.clear();
    for (uint32_t i = 0; i < e->stores.size(); ++i)
        if (!e->stores[i].dropped) e->store_by_name.emplace(e->stores[i].name, i);
    e->docs.swap(docs);
    e->free_ext.swap(free_ext);
    e->free_rows = 0;
    for (const Extent &x : e->free_ext) e->free_rows += x.hi - x.lo;
    e->n_rows = h.n_rows;
    e->epoch.fetch_add(1);
    return RF_OK;
}

// ---- one HBM index, several processes -----------------------------------------------------
// The reference runs four API worker processes and one ingest worker (backend/Dockerfile:42, worker.py:122-126).
// The daemon process owns the arena; rf_engine_export describes it -- CUDA IPC handles of the three row arrays
// plus the store table -- and another process maps it read-only with rf_engine_attach and then launches its own
// searches on its own streams.  Deletes are visible to readers at once (they mask rows in the shared arena);
// new rows and stores after rf_engine_refresh with a newer export.
namespace {
struct IpcHeader {
    char magic[8];                       // "RFB2IPC1"
    uint32_t dim, device;
    uint64_t capacity_rows, id_base, n_rows, epoch, n_stores;
    cudaIpcMemHandle_t h_F, h_seg, h_ff;
};
struct BlobWriter {
    uint8_t *p;
    size_t cap, len = 0;
    void put(const void *src, size_t n) {
        if (p && len + n <= cap) memcpy(p + len, src, n);
        len += n;
    }
};
struct BlobReader {
    const uint8_t *p;
    size_t len, at = 0;
    bool get(void *dst, size_t n) {
        if (at + n > len) return false;
        memcpy(dst, p + at, n);
        at += n;
        return true;
    }
};
// store table of an export -> stores / names (validated against n_rows)
int read_ipc_stores(BlobReader &r, const IpcHeader &h, std::vector<Store> &stores) {
    if (h.n_stores > (1ull << 32)) return fail(RF_EINVAL, "export: implausible store count");
    stores.resize(h.n_stores);
    for (Store &st : stores) {
        uint32_t len = 0, dropped = 0, n_ext = 0;
        if (!r.get(&len, 4) || len > (1u << 20)) return fail(RF_EINVAL, "export is truncated or corrupt");
        st.name.resize(len);
        if (!r.get(&st.name[0], len) || !r.get(&dropped, 4) || !r.get(&n_ext, 4) || n_ext > h.n_rows + 1)
            return fail(RF_EINVAL, "export is truncated or corrupt");
        st.dropped = dropped != 0;
        st.ext.resize(n_ext);
        if (!r.get(st.ext.data(), n_ext * sizeof(Extent)) || !extents_valid(st.ext, h.n_rows)) return fail(RF_EINVAL, "export is truncated or corrupt");
    }
    return RF_OK;
}
}  // namespace

int rf_engine_export(rf_engine *e, void *buf, size_t cap, size_t *len) {
    if (!e || !len) return fail(RF_EINVAL, "null argument");
    if (e->reader) return fail(RF_EINVAL, "a read-only attachment cannot be exported again: ask the owning process");
    RF_CUDA(cudaSetDevice(e->cfg.device));
    IpcHeader h{};
    memcpy(h.magic, "RFB2IPC1", 8);
    h.dim = e->dim;
    h.device = static_cast<uint32_t>(e->cfg.device);
    h.capacity_rows = e->cfg.capacity_rows;
    h.id_base = e->cfg.id_base;
    RF_CUDA(cudaIpcGetMemHandle(&h.h_F, e->F));
    RF_CUDA(cudaIpcGetMemHandle(&h.h_seg, e->seg));
    RF_CUDA(cudaIpcGetMemHandle(&h.h_ff, e->ff));
    BlobWriter w{static_cast<uint8_t *>(buf), buf ? cap : 0};
    std::shared_lock<std::shared_mutex> lk(e->meta_mu);     // rows below n_rows are complete: every ingest synchronises before it publishes
    h.n_rows = e->n_rows;
    h.epoch = e->epoch.load();
    h.n_stores = e->stores.size();
    w.put(&h, sizeof h);
    for (const Store &st : e->stores) {
        const uint32_t nlen = static_cast<uint32_t>(st.name.size()), dropped = st.dropped ? 1u : 0u, n_ext = static_cast<uint32_t>(st.ext.size());
        w.put(&nlen, 4); w.put(st.name.data(), nlen); w.put(&dropped, 4); w.put(&n_ext, 4); w.put(st.ext.data(), n_ext * sizeof(Extent));
    }
    *len = w.len;
    if (buf && w.len > cap) return fail(RF_ENOMEM, "export needs %zu bytes, the buffer holds %zu", w.len, cap);
    return RF_OK;
}

int rf_engine_refresh(rf_engine *e, const void *blob, size_t len) {
    if (!e || !blob) return fail(RF_EINVAL, "null argument");
    if (!e->reader) return fail(RF_EINVAL, "only a read-only attachment is refreshed");
    BlobReader r{static_cast<const uint8_t *>(blob), len};
    IpcHeader h{};
    if (!r.get(&h, sizeof h) || memcmp(h.magic, "RFB2IPC1", 8) != 0) return fail(RF_EINVAL, "not an engine export");
    if (h.dim != e->dim || h.capacity_rows != e->cfg.capacity_rows || h.id_base != e->cfg.id_base || static_cast<int>(h.device) != e->cfg.device)
        return fail(RF_EINVAL, "this export describes a different engine");
    if (h.n_rows > h.capacity_rows) return fail(RF_EINVAL, "export is corrupt (rows beyond capacity)");
    std::vector<Store> stores;
    if (int rc = read_ipc_stores(r, h, stores)) return rc;
    std::unique_lock<std::shared_mutex> lk(e->meta_mu);
    e->stores.swap(stores);
    e->store_by_name.clear();
    for (uint32_t i = 0; i < e->stores.size(); ++i)
        if (!e->stores[i].dropped) e->store_by_name.emplace(e->stores[i].name, i);
    e->n_rows = h.n_rows;
    e->epoch.fetch_add(1);          // the reader's own epoch: its store table and cached statistics follow
    return RF_OK;
}

int rf_engine_attach(const void *blob, size_t len, uint32_t n_contexts, rf_engine **out) {
    if (!blob || !out) return fail(RF_EINVAL, "null argument");
    *out = nullptr;
    BlobReader r{static_cast<const uint8_t *>(blob), len};
    IpcHeader h{};
    if (!r.get(&h, sizeof h) || memcmp(h.magic, "RFB2IPC1", 8) != 0) return fail(RF_EINVAL, "not an engine export");
    if (h.dim != 256 && h.dim != 512 && h.dim != 1024) return fail(RF_EINVAL, "export is corrupt (dim)");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(RF_ENODEVICE, "no CUDA device visible");
    }
    if (static_cast<int>(h.device) >= n_dev) return fail(RF_EINVAL, "the exporting process's device %u is not visible here (%d devices)", h.device, n_dev);
    cudaDeviceProp prop{};
    RF_CUDA(cudaGetDeviceProperties(&prop, static_cast<int>(h.device)));
    if (prop.major != 10) return fail(RF_ENODEVICE, "device %u is sm_%d%d; this library holds sm_100a code only", h.device, prop.major, prop.minor);
    RF_CUDA(cudaSetDevice(static_cast<int>(h.device)));
    rf_engine *e = new (std::nothrow) rf_engine();
    if (!e) return fail(RF_ENOMEM, "host allocation failed");
    e->reader = true;
    e->cfg.struct_size = sizeof(rf_config);
    e->cfg.device = static_cast<int>(h.device);
    e->cfg.dim = h.dim;
    e->cfg.n_contexts = n_contexts ? n_contexts : 8;
    e->cfg.capacity_rows = h.capacity_rows;
    e->cfg.id_base = h.id_base;
    e->dim = h.dim;
    e->tile_rows = rf::scan_tile_rows(h.dim);
    e->sm_count = prop.multiProcessorCount;
    if (const char *s = getenv("RF_GEMM")) e->gemm_enabled = atoi(s) != 0;
    if (const char *s = getenv("RF_STORE_TABLE")) e->table_enabled = atoi(s) != 0;
    cudaError_t ce;
    void *pF = nullptr, *pS = nullptr, *pN = nullptr;
    if ((ce = cudaIpcOpenMemHandle(&pF, h.h_F, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess ||
        (ce = cudaIpcOpenMemHandle(&pS, h.h_seg, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess ||
        (ce = cudaIpcOpenMemHandle(&pN, h.h_ff, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
        e->F = static_cast<int8_t *>(pF); e->seg = static_cast<uint32_t *>(pS); e->ff = static_cast<int32_t *>(pN);
        const int rc = fail(RF_ECUDA, "cudaIpcOpenMemHandle failed: %s (the arena must be opened from a process other than its owner)", cudaGetErrorString(ce));
        cudaGetLastError();
        rf_engine_destroy(e);
        return rc;
    }
    e->F = static_cast<int8_t *>(pF); e->seg = static_cast<uint32_t *>(pS); e->ff = static_cast<int32_t *>(pN);
    for (uint32_t i = 0; i < e->cfg.n_contexts; ++i) {
        SearchCtx *c = new (std::nothrow) SearchCtx();
        if (!c || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            rf_engine_destroy(e);
            return fail(RF_ECUDA, "search context creation failed");
        }
        e->all_ctx.push_back(c);
        e->free_ctx.push_back(c);
    }
    if (int rc = rf_engine_refresh(e, blob, len)) {
        rf_engine_destroy(e);
        return rc;
    }
    *out = e;
    return RF_OK;
}

int rf_rows_read(rf_engine *e, uint64_t first_row, uint64_t n, int8_t *rows, uint32_t *store_seg, int32_t *ff) {
    if (!e) return fail(RF_EINVAL, "null argument");
    {
        std::shared_lock<std::shared_mutex> lk(e->meta_mu);
        if (first_row + n > e->n_rows) return fail(RF_EINVAL, "rows %llu..%llu beyond n_rows %llu", (unsigned long long)first_row, (unsigned long long)(first_row + n), (unsigned long long)e->n_rows);
    }
    RF_CUDA(cudaSetDevice(e->cfg.device));
    if (rows) RF_CUDA(cudaMemcpy(rows, e->F + first_row * e->dim, n * e->dim, cudaMemcpyDeviceToHost));
    if (store_seg) RF_CUDA(cudaMemcpy(store_seg, e->seg + first_row, n * 4, cudaMemcpyDeviceToHost));
    if (ff) RF_CUDA(cudaMemcpy(ff, e->ff + first_row, n * 4, cudaMemcpyDeviceToHost));
    return RF_OK;
}

int rf_search_begin(rf_engine *e, const int8_t *q, uint32_t nq, const uint32_t *store_segs, const uint32_t *seg_off, uint32_t k,
                    rf_pending **out) {
    if (!e || !q || !seg_off || !out) return fail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (nq == 0 || nq > 65535) return fail(RF_EINVAL, "between 1 and 65535 queries per call");
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    // A batch whose queries all have the same scope takes the device-resident route when the
    // tensor-core path pays for it (or the batch is large: one shared plan instead of nq): one H2D of
    // the queries, the batched search, an unpack kernel, one D2H of the results.
    if (nq >= 2 && e->gemm_enabled) {
        const uint32_t n0 = seg_off[1] - seg_off[0];
        bool same = n0 <= RF_SCOPE_MAX;
        for (uint32_t i = 1; same && i < nq; ++i) {
            same = seg_off[i + 1] - seg_off[i] == n0;
            for (uint32_t j = 0; same && j < n0; ++j) same = store_segs[seg_off[i] + j] == store_segs[seg_off[0] + j];
        }
        if (same) {
            std::vector<Extent> ext;
            {
                std::shared_lock<std::shared_mutex> lk(e->meta_mu);
                gather_extents(e, store_segs + seg_off[0], n0, ext);
            }
            // 20 us: the route's own copies, unpack kernel and stream synchronisation
            uint64_t rows = 0;
            for (const Extent &x : ext) rows += x.hi - x.lo;
            same = nq >= 64 || (!ext.empty() && gemm_pays(e, nq, rows, ext.back().hi - ext.front().lo, k, 20.0));
        }
        if (same) {
            SearchCtx *c = ctx_acquire(e);
            if (!c) return fail(RF_EBUSY, "no search context free after 5 s");
            CtxGuard g{e, c};
            RF_CUDA(cudaSetDevice(e->cfg.device));
            const OutLayout L(nq, k);
            const size_t q_bytes = static_cast<size_t>(nq) * e->dim;
            RF_CUDA(c->h_in.reserve(q_bytes));
            RF_CUDA(c->d_in.reserve(q_bytes));
            RF_CUDA(c->d_out.reserve(L.total));
            RF_CUDA(c->h_out.reserve(L.total));
            memcpy(c->h_in.p, q, q_bytes);
            RF_CUDA(cudaMemcpyAsync(c->d_in.p, c->h_in.p, q_bytes, cudaMemcpyHostToDevice, c->stream));
            uint8_t *d_out = static_cast<uint8_t *>(c->d_out.p);
            uint64_t *d_keys = reinterpret_cast<uint64_t *>(d_out + L.off_keys);
            int rc = search_keys_device_impl(e, static_cast<const int8_t *>(c->d_in.p), nq, store_segs + seg_off[0], n0, k, d_keys, c->stream, nullptr);
            if (rc) return rc;
            RF_CUDA(rf::launch_unpack_keys(d_keys, static_cast<const int8_t *>(c->d_in.p), e->dim, e->ff, static_cast<uint32_t>(e->cfg.id_base), nq, k,
                                           reinterpret_cast<uint64_t *>(d_out + L.off_ids), reinterpret_cast<int32_t *>(d_out + L.off_scores),
                                           reinterpret_cast<float *>(d_out + L.off_cos), reinterpret_cast<uint32_t *>(d_out + L.off_counts), c->stream));
            e->launches.fetch_add(1, std::memory_order_relaxed);
            RF_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(c->h_out.p) + L.off_ids, d_out + L.off_ids, L.total - L.off_ids, cudaMemcpyDeviceToHost, c->stream));
            c->pend = SearchCtx::Pending{};
            c->pend.nq = nq;
            c->pend.k = k;
            c->pend.active = true;
            c->pend.t_launched = std::chrono::steady_clock::now();
            g.c = nullptr;     // the caller's rf_search_end releases the context
            *out = reinterpret_cast<rf_pending *>(c);
            return RF_OK;
        }
    }
    PlanBlob b;
    int rc = RF_OK;
    // Several queries with their own scopes: the kernel reads each query's extents from the device-resident
    // store table, so the host builds no plan per query (unless a scope has more extents than a plan holds).
    std::shared_lock<std::shared_mutex> tbl_lock;
    TableQuery tq{store_segs, seg_off, 0};
    bool use_table = false;
    if (nq >= kTableMinQueries && e->table_enabled) {
        if ((rc = table_acquire(e, tbl_lock))) return rc;
        use_table = table_batch_ok(e, nq, store_segs, seg_off, &tq.max_tiles);
        if (!use_table) tbl_lock.unlock();
    }
    if (!use_table && (rc = build_blob(e, q, nq, store_segs, seg_off, false, b))) return rc;
    SearchCtx *c = ctx_acquire(e);
    if (!c) return fail(RF_EBUSY, "no search context free after 5 s");
    CtxGuard g{e, c};
    RF_CUDA(cudaSetDevice(e->cfg.device));
    if (e->profile) e->prof_ns[0] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    if ((rc = search_launch(e, c, b, q, nq, k, false, use_table ? &tq : nullptr))) return rc;
    e->searches.fetch_add(nq, std::memory_order_relaxed);
    g.c = nullptr;
    *out = reinterpret_cast<rf_pending *>(c);
    return RF_OK;
}

int rf_search_end(rf_engine *e, rf_pending *p, uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts,
                  int8_t *out_q) {
    if (!e || !p) return fail(RF_EINVAL, "null argument");
    SearchCtx *c = reinterpret_cast<SearchCtx *>(p);
    CtxGuard g{e, c};              // the context goes back to the pool whatever happens below
    if (!out_ids || !out_scores) {
        cudaSetDevice(e->cfg.device);
        cudaStreamSynchronize(c->stream);
        c->pend.active = false;
        return fail(RF_EINVAL, "null argument");
    }
    RF_CUDA(cudaSetDevice(e->cfg.device));
    return search_wait(e, c, out_ids, out_scores, out_cos, out_counts, out_q);
}

int rf_search(rf_engine *e, const int8_t *q, uint32_t nq, const uint32_t *store_segs, const uint32_t *seg_off, uint32_t k,
              uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts) {
    if (!e || !q || !seg_off || !out_ids || !out_scores) return fail(RF_EINVAL, "null argument");
    if (nq == 0) return (k == 0 || k > RF_TOPK_MAX) ? fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX) : RF_OK;
    rf_pending *p = nullptr;
    const int rc = rf_search_begin(e, q, nq, store_segs, seg_off, k, &p);
    if (rc) return rc;
    return rf_search_end(e, p, out_ids, out_scores, out_cos, out_counts, nullptr);
}

int rf_featurize_query(rf_engine *e, const uint8_t *utf8, size_t n, int8_t *out_q) {
    if (!e || (!utf8 && n) || !out_q) return fail(RF_EINVAL, "null argument");
    if (n > (1u << 26)) return fail(RF_EINVAL, "query text too long");
    SearchCtx *c = ctx_acquire(e);
    if (!c) return fail(RF_EBUSY, "no search context free after 5 s");
    CtxGuard g{e, c};
    RF_CUDA(cudaSetDevice(e->cfg.device));
    const size_t text_pad = (n + 255) & ~static_cast<size_t>(255);
    RF_CUDA(c->h_in.reserve(n + e->dim));
    RF_CUDA(c->d_in.reserve(text_pad + e->dim));
    if (n) memcpy(c->h_in.p, utf8, n);
    uint8_t *d_text = static_cast<uint8_t *>(c->d_in.p);
    int8_t *d_q = reinterpret_cast<int8_t *>(d_text + text_pad);
    if (n) RF_CUDA(cudaMemcpyAsync(d_text, c->h_in.p, n, cudaMemcpyHostToDevice, c->stream));
    RF_CUDA(rf::launch_featurize_query(d_text, static_cast<uint32_t>(n), nullptr, d_q, e->dim, c->stream));
    e->launches.fetch_add(1);
    RF_CUDA(cudaMemcpyAsync(c->h_in.p, d_q, e->dim, cudaMemcpyDeviceToHost, c->stream));
    RF_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(out_q, c->h_in.p, e->dim);
    return RF_OK;
}

int rf_search_text(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs, uint32_t n_segs, uint32_t k,
                   uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_count, int8_t *out_q) {
    return rf_search_text_in(e, utf8, n, store_segs, n_segs, nullptr, 0, k, out_ids, out_scores, out_cos, out_count, out_q);
}

int rf_search_text_in(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs, uint32_t n_segs,
                      const uint64_t *ranges, uint32_t n_ranges, uint32_t k, uint64_t *out_ids, int32_t *out_scores,
                      float *out_cos, uint32_t *out_count, int8_t *out_q) {
    return rf_search_text_w(e, utf8, n, store_segs, n_segs, ranges, n_ranges, nullptr, k, out_ids, out_scores, out_cos, out_count, out_q);
}

int rf_search_text_begin(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs, uint32_t n_segs,
                         const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights, uint32_t k, rf_pending **out) {
    if (!e || (!utf8 && n) || !out || (!ranges && n_ranges)) return fail(RF_EINVAL, "null argument");
    *out = nullptr;
    if (k == 0 || k > RF_TOPK_MAX) return fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (n > (1u << 26)) return fail(RF_EINVAL, "query text too long");
    std::vector<Extent> lim;
    for (uint32_t i = 0; i < n_ranges; ++i) {
        const uint64_t lo = ranges[2 * i], hi = ranges[2 * i + 1];
        if (hi < lo || (i && lo < ranges[2 * i - 1])) return fail(RF_EINVAL, "ranges must be sorted and disjoint");
        const uint64_t base = e->cfg.id_base, cap = e->cfg.capacity_rows;
        const uint64_t rlo = lo > base ? lo - base : 0, rhi = hi > base ? hi - base : 0;
        if (rlo < rhi && rlo < cap) {
            const uint32_t l32 = static_cast<uint32_t>(rlo), h32 = static_cast<uint32_t>(std::min(rhi, cap));
            if (!lim.empty() && lim.back().hi == l32) lim.back().hi = h32;   // adjacent documents merge
            else lim.push_back({l32, h32});
        }
    }
    const uint32_t seg_off[2] = {0, n_segs};
    PlanBlob b;
    int rc = build_blob(e, nullptr, 1, store_segs, seg_off, false, b, n_ranges ? &lim : nullptr);
    if (rc) return rc;
    SearchCtx *c = ctx_acquire(e);
    if (!c) return fail(RF_EBUSY, "no search context free after 5 s");
    CtxGuard g{e, c};
    RF_CUDA(cudaSetDevice(e->cfg.device));
    // weights, query text and the query vector share the context's input buffer: one H2D copy (the plan
    // of a single query rides in the kernel parameters)
    const size_t text_pad = (n + 255) & ~static_cast<size_t>(255);
    const size_t w_pad = weights ? e->dim : 0;   // [weights | text | query vector]
    const size_t need = w_pad + text_pad + e->dim;
    RF_CUDA(c->h_in.reserve(need));
    RF_CUDA(c->d_in.reserve(need));
    uint8_t *h = static_cast<uint8_t *>(c->h_in.p);
    uint8_t *d = static_cast<uint8_t *>(c->d_in.p);
    if (weights) memcpy(h, weights, e->dim);
    if (n) memcpy(h + w_pad, utf8, n);
    if (w_pad + n) RF_CUDA(cudaMemcpyAsync(d, h, w_pad + n, cudaMemcpyHostToDevice, c->stream));
    int8_t *d_q = reinterpret_cast<int8_t *>(d + w_pad + text_pad);
    RF_CUDA(rf::launch_featurize_query(d + w_pad, static_cast<uint32_t>(n), weights ? d : nullptr, d_q, e->dim, c->stream));
    e->launches.fetch_add(1);

    const OutLayout L(1, k);
    const uint32_t X = pick_blocks(e, 1, b.max_tiles);
    RF_CUDA(c->h_out.reserve(L.total + e->dim));
    RF_CUDA(c->d_out.reserve(L.total));
    RF_CUDA(c->d_partial.reserve(static_cast<size_t>(X) * k * 8));
    if (c->d_tickets.cap < kSyncBytesPerQuery + 8) {
        RF_CUDA(c->d_tickets.reserve(kSyncBytesPerQuery + 8));
        RF_CUDA(cudaMemsetAsync(c->d_tickets.p, 0, c->d_tickets.cap, c->stream));
    }
    ScanArgs a{};
    fill_args(e, a, nullptr, b, d_q, k, false);
    maybe_inline_plan(a, b, 1, false);
    if (!a.inline_plan) return fail(RF_EINVAL, "internal: a single-query plan exceeds %u extents", rf::kInlineExt);
    uint8_t *d_out = static_cast<uint8_t *>(c->d_out.p);
    a.partial = static_cast<uint64_t *>(c->d_partial.p);
    set_sync_bufs(a, c->d_tickets, c->launches++);
    a.out_keys = reinterpret_cast<uint64_t *>(d_out + L.off_keys);
    a.out_ids = reinterpret_cast<uint64_t *>(d_out + L.off_ids);
    a.out_scores = reinterpret_cast<int32_t *>(d_out + L.off_scores);
    a.out_cos = reinterpret_cast<float *>(d_out + L.off_cos);
    a.out_counts = reinterpret_cast<uint32_t *>(d_out + L.off_counts);
    // the query vector comes from the kernel just before this one on the stream: a fully serialised launch
    // (an overlapped one may read d_q before featurize_query's writes are visible)
    RF_CUDA(rf::launch_score_topk_scan(a, 1, X, e->scan_variant, e->dim, c->stream, false));
    e->launches.fetch_add(1);
    uint8_t *ho = static_cast<uint8_t *>(c->h_out.p);
    RF_CUDA(cudaMemcpyAsync(ho + L.off_ids, d_out + L.off_ids, L.total - L.off_ids, cudaMemcpyDeviceToHost, c->stream));
    RF_CUDA(cudaMemcpyAsync(ho + L.total, d_q, e->dim, cudaMemcpyDeviceToHost, c->stream));
    c->pend = SearchCtx::Pending{};
    c->pend.nq = 1;
    c->pend.k = k;
    c->pend.has_q = true;
    c->pend.active = true;
    c->pend.t_launched = std::chrono::steady_clock::now();
    e->searches.fetch_add(1);
    g.c = nullptr;
    *out = reinterpret_cast<rf_pending *>(c);
    return RF_OK;
}

int rf_search_text_w(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs, uint32_t n_segs,
                     const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights, uint32_t k, uint64_t *out_ids,
                     int32_t *out_scores, float *out_cos, uint32_t *out_count, int8_t *out_q) {
    if (!out_ids || !out_scores) return fail(RF_EINVAL, "null argument");
    rf_pending *p = nullptr;
    const int rc = rf_search_text_begin(e, utf8, n, store_segs, n_segs, ranges, n_ranges, weights, k, &p);
    if (rc) return rc;
    return rf_search_end(e, p, out_ids, out_scores, out_cos, out_count, out_q);
}

// ---- RF-1w statistics ------------------------------------------------------------------------------
static int fill_df_args(rf_engine *e, const uint32_t *store_segs, uint32_t n_segs, rf::DfArgs &a) {
    if (n_segs > RF_SCOPE_MAX) return fail(RF_EINVAL, "scope has %u segments (max %u)", n_segs, RF_SCOPE_MAX);
    memset(&a, 0, sizeof a);
    a.F = e->F;
    a.seg = e->seg;
    a.n_scope = n_segs;
    a.dim = e->dim;
    for (uint32_t j = 0; j < RF_SCOPE_MAX; ++j) a.scope[j] = j < n_segs ? store_segs[j] : RF_TOMBSTONE;
    std::vector<Extent> ext;
    {
        std::shared_lock<std::shared_mutex> lk(e->meta_mu);
        gather_extents(e, store_segs, n_segs, ext);
    }
    static_assert(rf::kDfMaxExtents >= kMaxExtPerQuery, "extent list must fit the kernel parameters");
    a.n_ext = static_cast<uint32_t>(ext.size());
    uint32_t rows = 0;
    for (uint32_t i = 0; i < a.n_ext; ++i) {
        a.lo[i] = ext[i].lo;
        a.prefix[i] = rows;
        rows += ext[i].hi - ext[i].lo;
    }
    a.prefix[a.n_ext] = rows;
    return RF_OK;
}

int rf_scope_df_device(rf_engine *e, const uint32_t *store_segs, uint32_t n_segs, uint64_t *df_dev, void *stream) {
    if (!e || (!store_segs && n_segs) || !df_dev) return fail(RF_EINVAL, "null argument");
    rf::DfArgs a;
    int rc = fill_df_args(e, store_segs, n_segs, a);
    if (rc) return rc;
    a.out = reinterpret_cast<unsigned long long *>(df_dev);
    RF_CUDA(cudaSetDevice(e->cfg.device));
    RF_CUDA(rf::launch_bucket_df(a, e->sm_count, static_cast<cudaStream_t>(stream)));
    if (a.prefix[a.n_ext]) e->launches.fetch_add(1, std::memory_order_relaxed);
    return RF_OK;
}

int rf_scope_df(rf_engine *e, const uint32_t *store_segs, uint32_t n_segs, uint64_t *out_df, uint64_t *out_n) {
    if (!e || (!store_segs && n_segs) || !out_df || !out_n) return fail(RF_EINVAL, "null argument");
    std::vector<uint32_t> key(store_segs, store_segs + n_segs);
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    const uint64_t gen = e->epoch.load() + e->tomb_gen.load();   // both only grow
    {
        std::lock_guard<std::mutex> lk(e->df_mu);
        auto it = e->df_cache.find(key);
        if (it != e->df_cache.end() && it->second.first == gen) {
            memcpy(out_df, it->second.second.data(), e->dim * 8);
            *out_n = it->second.second[e->dim];
            return RF_OK;
        }
    }
    SearchCtx *c = ctx_acquire(e);
    if (!c) return fail(RF_EBUSY, "no search context free after 5 s");
    CtxGuard g{e, c};
    RF_CUDA(cudaSetDevice(e->cfg.device));
    const size_t bytes = (e->dim + 1) * 8;
    RF_CUDA(c->d_out.reserve(bytes));
    RF_CUDA(c->h_out.reserve(bytes));
    RF_CUDA(cudaMemsetAsync(c->d_out.p, 0, bytes, c->stream));
    int rc = rf_scope_df_device(e, store_segs, n_segs, static_cast<uint64_t *>(c->d_out.p), c->stream);
    if (rc) return rc;
    RF_CUDA(cudaMemcpyAsync(c->h_out.p, c->d_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    RF_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t *h = static_cast<const uint64_t *>(c->h_out.p);
    memcpy(out_df, h, e->dim * 8);
    *out_n = h[e->dim];
    {
        std::lock_guard<std::mutex> lk(e->df_mu);
        if (e->df_cache.size() > 4096) e->df_cache.clear();
        e->df_cache[key] = {gen, std::vector<uint64_t>(h, h + e->dim + 1)};
    }
    return RF_OK;
}

int rf_idf_weights(const uint64_t *df, uint64_t n, uint32_t dim, uint8_t *out_w) {
    if (!df || !out_w) return fail(RF_EINVAL, "null argument");
    if (n >= (1ull << 55)) return fail(RF_EINVAL, "row count out of range");
    for (uint32_t d = 0; d < dim; ++d) {
        if (df[d] > n) return fail(RF_EINVAL, "df[%u] exceeds the row count", d);
        const uint64_t r = ((n + 1) << 8) / (df[d] + 1);          // >= 256
        const int lg = 63 - __builtin_clzll(r);
        out_w[d] = static_cast<uint8_t>(std::min<uint64_t>(4 + 4 * static_cast<uint64_t>(lg - 8) + ((r >> (lg - 2)) & 3), 31));
    }
    return RF_OK;
}

int rf_weight_query(const int8_t *q, const uint8_t *w, uint32_t dim, int8_t *out_qw) {
    if (!q || !w || !out_qw) return fail(RF_EINVAL, "null argument");
    for (uint32_t d = 0; d < dim; ++d) {
        if (q[d] < 0) return fail(RF_EINVAL, "query features are counts (>= 0)");
        out_qw[d] = static_cast<int8_t>(std::min<int>(static_cast<int>(q[d]) * w[d], 127));
    }
    return RF_OK;
}

int rf_search_keys_device(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs, uint32_t n_segs,
                          uint32_t k, uint64_t *out_keys_dev, void *stream) {
    return search_keys_device_impl(e, q_dev, nq, store_segs, n_segs, k, out_keys_dev, stream, nullptr);
}

int rf_search_keys_device_fused(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs, uint32_t n_segs,
                                uint32_t k, const rf_peer_exchange *px, uint64_t *out_keys_dev, void *stream) {
    if (!px || px->struct_size != sizeof(rf_peer_exchange)) return fail(RF_EINVAL, "bad rf_peer_exchange");
    if (px->world == 0 || px->world > 8 || px->rank >= px->world) return fail(RF_EINVAL, "world must be in [1, 8] and rank < world");
    if (!px->keys_ptrs || !px->flag_ptrs || !px->timeout_flag_dev) return fail(RF_EINVAL, "null exchange buffer");
    if (nq > px->nq_cap || k != px->k || px->seq == 0) return fail(RF_EINVAL, "exchange buffers are sized for nq <= %u, k == %u, seq > 0", px->nq_cap, px->k);
    return search_keys_device_impl(e, q_dev, nq, store_segs, n_segs, k, out_keys_dev, stream, px);
}

// The per-stream scratch of the device-resident searches (created on first use; when the table is full the
// least recently used state nobody holds is released -- cudaFree synchronises the device, so kernels that
// still read its buffers have finished).
static std::shared_ptr<StreamState> stream_state(rf_engine *e, void *stream) {
    std::lock_guard<std::mutex> lk(e->plan_mu);
    std::shared_ptr<StreamState> &slot = e->stream_states[stream];
    if (!slot) {
        if (e->stream_states.size() > kMaxStreamStates) {
            auto victim = e->stream_states.end();
            for (auto it = e->stream_states.begin(); it != e->stream_states.end(); ++it)
                if (it->second && it->second.use_count() == 1 && (victim == e->stream_states.end() || it->second->last_use < victim->second->last_use))
                    victim = it;
            if (victim != e->stream_states.end()) {
                victim->second->release();
                e->stream_states.erase(victim);
            }
        }
        e->stream_states[stream] = std::make_shared<StreamState>();
        return stream_state_touch(e, e->stream_states[stream]);
    }
    return stream_state_touch(e, slot);
}

int rf_stream_set_overlap(rf_engine *e, void *stream, int allow) {
    if (!e) return fail(RF_EINVAL, "null argument");
    std::shared_ptr<StreamState> st = stream_state(e, stream);
    std::lock_guard<std::mutex> lk(st->mu);
    st->overlap = allow != 0;
    return RF_OK;
}

static int search_keys_device_impl(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs, uint32_t n_segs,
                                   uint32_t k, uint64_t *out_keys_dev, void *stream, const rf_peer_exchange *px) {
    if (!e || !q_dev || !out_keys_dev) return fail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (nq == 0) return RF_OK;
    if (nq > 65535) return fail(RF_EINVAL, "at most 65535 queries per call");
    if (n_segs > RF_SCOPE_MAX) return fail(RF_EINVAL, "scope has %u segments (max %u)", n_segs, RF_SCOPE_MAX);
    RF_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // one shared plan, built on the host from the scope's current extents and passed in the kernel
    // parameters: nothing about a scope is resident on the device between calls
    PlanBlob b;
    const uint32_t seg_off[2] = {0, n_segs};
    int rc = build_blob(e, nullptr, 1, store_segs, seg_off, true, b);
    if (rc) return rc;
    const ScanPlan *hp = reinterpret_cast<const ScanPlan *>(b.bytes.data() + b.off_plans);
    std::shared_ptr<StreamState> st = stream_state(e, stream);
    std::lock_guard<std::mutex> lk(st->mu);
    // ---- batched tensor-core path: k <= 10, when it beats nq scans (gemm_pays) ----
    if (!px && hp->n_ext >= 1) {
        uint64_t rows;
        uint32_t lo, hi;
        extent_rows(reinterpret_cast<const uint32_t *>(b.bytes.data() + b.off_lo) + hp->ext_off,
                    reinterpret_cast<const uint32_t *>(b.bytes.data() + b.off_hi) + hp->ext_off, hp->n_ext, rows, lo, hi);
        if (gemm_pays(e, nq, rows, hi - lo, k)) {
            rc = search_gemm(e, st.get(), q_dev, nq, hp, lo, hi, k, out_keys_dev, s);
            if (rc == RF_OK) e->searches.fetch_add(nq, std::memory_order_relaxed);
            return rc;
        }
    }
    const uint32_t X = pick_blocks(e, nq, b.max_tiles);
    const size_t need_partial = static_cast<size_t>(nq) * X * k * 8;
    const size_t need_sync = static_cast<size_t>(nq) * kSyncBytesPerQuery + 8;
    if (need_partial > st->partial.cap || need_sync > st->tickets.cap) {
        RF_CUDA(cudaStreamSynchronize(s));
        RF_CUDA(st->partial.reserve(need_partial));
        if (need_sync > st->tickets.cap) {
            RF_CUDA(st->tickets.reserve(need_sync));
            RF_CUDA(cudaMemsetAsync(st->tickets.p, 0, st->tickets.cap, s));   // on the caller's stream: the legacy default stream does not order with a non-blocking one
        }
    }
    ScanArgs a{};
    fill_args(e, a, nullptr, b, q_dev, k, true);
    maybe_inline_plan(a, b, nq, true);
    if (!a.inline_plan) return fail(RF_EINVAL, "internal: a single-scope plan exceeds %u extents", rf::kInlineExt);
    a.partial = static_cast<uint64_t *>(st->partial.p);
    set_sync_bufs(a, st->tickets, st->launches++);
    a.out_keys = out_keys_dev;
    if (px) {
        // fused exchange: the scan publishes into every rank's gather buffer, the merge waits on flags
        const size_t slot = px->seq & 3u;   // four buffers: a rank is never more than three calls ahead of a peer's merge
        const size_t keys_off = slot * px->world * px->nq_cap * static_cast<size_t>(k);
        const size_t flag_off = slot * px->world * static_cast<size_t>(px->nq_cap);
        for (uint32_t r = 0; r < px->world; ++r) {
            a.px_keys[r] = reinterpret_cast<uint64_t *>(px->keys_ptrs[r]) + keys_off;
            a.px_flags[r] = reinterpret_cast<uint32_t *>(px->flag_ptrs[r]) + flag_off;
        }
        a.px_rank = px->rank;
        a.px_world = px->world;
        a.px_seq = px->seq;
        a.px_nq_cap = px->nq_cap;
        a.px_out = out_keys_dev;
        a.px_timeout = px->timeout_flag_dev;
        if (static_cast<size_t>(nq) * k * 8 > st->gemm_keys_a.cap) {          // the local (pre-merge) list
            RF_CUDA(cudaStreamSynchronize(s));
            RF_CUDA(st->gemm_keys_a.reserve(static_cast<size_t>(nq) * k * 8));
        }
        a.out_keys = static_cast<uint64_t *>(st->gemm_keys_a.p);
    }
    RF_CUDA(rf::launch_score_topk_scan(a, nq, X, e->scan_variant, e->dim, s, st->overlap));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    e->searches.fetch_add(nq, std::memory_order_relaxed);
    return RF_OK;
}

static int search_keys_device_scoped_impl(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                          const uint32_t *seg_off, uint32_t k, uint64_t *out_keys_dev, void *stream,
                                          const rf_peer_exchange *px) {
    if (!e || !q_dev || !seg_off || !out_keys_dev) return fail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (nq > 65535) return fail(RF_EINVAL, "at most 65535 queries per call");
    const uint32_t nq_total = (px && px->nq_total) ? px->nq_total : nq;     // queries of the whole batch (rows of q_dev, rows of the result)
    if (nq_total == 0) return RF_OK;
    RF_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // plans: from the device-resident store table when every scope fits a plan, else built here per query
    std::shared_lock<std::shared_mutex> tbl_lock;
    bool use_table = false;
    uint32_t max_tiles = 0;
    int rc = RF_OK;
    if (e->table_enabled && nq) {
        if ((rc = table_acquire(e, tbl_lock))) return rc;
        use_table = table_batch_ok(e, nq, store_segs, seg_off, &max_tiles);
        if (!use_table) tbl_lock.unlock();
    }
    PlanBlob b;
    auto align = [](size_t x) { return (x + 15) & ~static_cast<size_t>(15); };
    size_t off_off = 0, off_segs = 0;
    if (use_table) {
        const size_t n_segs = seg_off[nq];
        off_segs = align((static_cast<size_t>(nq) + 1) * 4);
        b.bytes.resize(align(off_segs + std::max<size_t>(n_segs, 1) * 4));
        memcpy(b.bytes.data() + off_off, seg_off, (static_cast<size_t>(nq) + 1) * 4);
        if (n_segs) memcpy(b.bytes.data() + off_segs, store_segs, n_segs * 4);
        b.max_tiles = max_tiles;
    } else if (nq && (rc = build_blob(e, nullptr, nq, store_segs, seg_off, false, b))) {
        return rc;
    }
    // the exchange's own tables ride behind the plans in the same upload
    size_t off_qindex = 0, off_masks = 0;
    if (px && px->q_index && nq) {
        off_qindex = align(b.bytes.size());
        b.bytes.resize(off_qindex + static_cast<size_t>(nq) * 4);
        memcpy(b.bytes.data() + off_qindex, px->q_index, static_cast<size_t>(nq) * 4);
    }
    if (px && px->owner_masks) {
        off_masks = align(std::max<size_t>(b.bytes.size(), 16));
        b.bytes.resize(off_masks + nq_total);
        memcpy(b.bytes.data() + off_masks, px->owner_masks, nq_total);
    }
    std::shared_ptr<StreamState> st = stream_state(e, stream);
    std::lock_guard<std::mutex> lk(st->mu);
    const uint32_t X = pick_blocks(e, std::max(nq, 1u), b.max_tiles);
    // sized for the whole batch and a few blocks per query whatever this call's share is: a rank's share of a batch varies
    // from call to call, and growing the scratch frees the old one (a device-wide synchronisation, mid-exchange)
    const size_t need_partial = static_cast<size_t>(std::max(nq, nq_total)) * std::max(X, 4u) * k * 8;
    const size_t need_sync = static_cast<size_t>(nq) * kSyncBytesPerQuery + 8;
    const size_t need_local = px ? static_cast<size_t>(nq) * k * 8 : 0;
    if (b.bytes.size() > st->blob.cap || need_partial > st->partial.cap || need_sync > st->tickets.cap || need_local > st->gemm_keys_a.cap) {
        RF_CUDA(cudaStreamSynchronize(s));
        RF_CUDA(st->blob.reserve(b.bytes.size()));
        RF_CUDA(st->partial.reserve(need_partial));
        RF_CUDA(st->gemm_keys_a.reserve(need_local));
        if (need_sync > st->tickets.cap) {
            RF_CUDA(st->tickets.reserve(need_sync));
            RF_CUDA(cudaMemsetAsync(st->tickets.p, 0, st->tickets.cap, s));   // on the caller's stream: the legacy default stream does not order with a non-blocking one
        }
    }
    if (!b.bytes.empty()) {
        if (!st->h_blob_free) RF_CUDA(cudaEventCreateWithFlags(&st->h_blob_free, cudaEventDisableTiming));
        else RF_CUDA(cudaEventSynchronize(st->h_blob_free));      // the previous call's upload has left the staging buffer
        RF_CUDA(st->h_blob.reserve(b.bytes.size()));
        memcpy(st->h_blob.p, b.bytes.data(), b.bytes.size());
        RF_CUDA(cudaMemcpyAsync(st->blob.p, st->h_blob.p, b.bytes.size(), cudaMemcpyHostToDevice, s));
        RF_CUDA(cudaEventRecord(st->h_blob_free, s));
    }
    const uint8_t *d_blob = static_cast<const uint8_t *>(st->blob.p);
    ScanArgs a{};
    if (use_table) {
        a.F = e->F;
        a.seg = e->seg;
        a.ff = e->ff;
        a.q = q_dev;
        a.id_base = static_cast<uint32_t>(e->cfg.id_base);
        a.k = k;
        a.debug_ts = e->debug_ts;
        a.dbg_flags = e->dbg_flags;
        fill_table_args(e, a, d_blob, off_off, off_segs);
    } else if (nq) {
        fill_args(e, a, d_blob, b, q_dev, k, false);
    }
    if (off_qindex) a.q_index = reinterpret_cast<const uint32_t *>(d_blob + off_qindex);
    a.partial = static_cast<uint64_t *>(st->partial.p);
    if (nq) set_sync_bufs(a, st->tickets, st->launches++);
    a.out_keys = out_keys_dev;
    uint64_t *gather_local = nullptr;
    uint32_t *flags_local = nullptr;
    if (px) {
        // Store-sharded exchange: a query's rows live on one rank (or a few).  A rank launches only the queries
        // it has rows for.  A block that WAITED for its peers here would hold an SM slot the peers' own scans may
        // need (hundreds of such blocks per launch), so the scan only publishes -- keys into every rank's
        // gather buffer, then a release flag -- and a second, tiny kernel behind it acquires the flags of each
        // query's owners and merges (one warp per query of the whole batch).
        const size_t slot = px->seq & 3u;
        const size_t keys_off = slot * px->world * px->nq_cap * static_cast<size_t>(k);
        const size_t flag_off = slot * px->world * static_cast<size_t>(px->nq_cap);
        for (uint32_t r = 0; r < px->world; ++r) {
            a.px_keys[r] = reinterpret_cast<uint64_t *>(px->keys_ptrs[r]) + keys_off;
            a.px_flags[r] = reinterpret_cast<uint32_t *>(px->flag_ptrs[r]) + flag_off;
        }
        a.px_rank = px->rank;
        a.px_world = px->world;
        a.px_seq = px->seq;
        a.px_nq_cap = px->nq_cap;
        a.px_out = out_keys_dev;
        a.px_timeout = px->timeout_flag_dev;
        a.px_publish_only = 1;
        a.out_keys = static_cast<uint64_t *>(st->gemm_keys_a.p);    // the local (pre-merge) lists
        gather_local = a.px_keys[px->rank];
        flags_local = a.px_flags[px->rank];
    }
    uint32_t launched = 0;
    if (nq) {
        // the plans / scope lists were copied in just above (a copy, not a kernel): the overlap rule only concerns q_dev
        RF_CUDA(rf::launch_score_topk_scan(a, nq, X, e->scan_variant, e->dim, s, st->overlap));
        ++launched;
    }
    if (px && px->world > 1) {
        RF_CUDA(rf::launch_merge_wait(gather_local, flags_local, off_masks ? d_blob + off_masks : nullptr, px->world, px->nq_cap, nq_total, k,
                                      px->seq, out_keys_dev, px->timeout_flag_dev, s));
        ++launched;
    } else if (px && nq) {
        RF_CUDA(cudaMemcpyAsync(out_keys_dev, a.out_keys, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToDevice, s));
    }
    e->launches.fetch_add(launched, std::memory_order_relaxed);
    e->searches.fetch_add(nq, std::memory_order_relaxed);
    return RF_OK;
}

int rf_search_keys_device_scoped(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                 const uint32_t *seg_off, uint32_t k, uint64_t *out_keys_dev, void *stream) {
    return search_keys_device_scoped_impl(e, q_dev, nq, store_segs, seg_off, k, out_keys_dev, stream, nullptr);
}

int rf_search_keys_device_scoped_fused(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                       const uint32_t *seg_off, uint32_t k, const rf_peer_exchange *px, uint64_t *out_keys_dev,
                                       void *stream) {
    if (!px || px->struct_size != sizeof(rf_peer_exchange)) return fail(RF_EINVAL, "bad rf_peer_exchange");
    if (px->world == 0 || px->world > 8 || px->rank >= px->world) return fail(RF_EINVAL, "world must be in [1, 8] and rank < world");
    if (!px->keys_ptrs || !px->flag_ptrs || !px->timeout_flag_dev) return fail(RF_EINVAL, "null exchange buffer");
    const uint32_t total = px->nq_total ? px->nq_total : nq;
    if (nq > total || total > px->nq_cap || k != px->k || px->seq == 0) return fail(RF_EINVAL, "exchange buffers are sized for nq_total <= %u, k == %u, seq > 0", px->nq_cap, px->k);
    if (px->q_index)
        for (uint32_t j = 0; j < nq; ++j)
            if (px->q_index[j] >= total) return fail(RF_EINVAL, "q_index[%u] = %u is not a query of the batch (%u)", j, px->q_index[j], total);
    return search_keys_device_scoped_impl(e, q_dev, nq, store_segs, seg_off, k, out_keys_dev, stream, px);
}

int rf_merge_topk_device(rf_engine *e, const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k,
                         uint64_t *out_keys_dev, void *stream) {
    if (!e || !keys_dev || !out_keys_dev) return fail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return fail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (nq == 0 || n_lists == 0) return RF_OK;
    RF_CUDA(cudaSetDevice(e->cfg.device));
    RF_CUDA(rf::launch_merge_topk(keys_dev, n_lists, nq, k, out_keys_dev, static_cast<cudaStream_t>(stream)));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    return RF_OK;
}

}  // extern "C"
