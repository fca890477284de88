// Engine group: several GPUs behind ONE host-facing index (include/rf_b200.h, rf_group_*).
//
// The reference's adapter object is one per request over one process-global backend
// (backend/app/services/gemini_rag.py:721-725; callers routes/chat.py:499-505 and
// services/ingestion.py:45-52; 4 API workers + 1 ARQ worker, backend/Dockerfile:42, worker.py:122-126):
// whatever serves it must look like a single index.  A group owns one rf_engine per device in THIS process
// and presents the single-engine surface:
//   * stores are opened on every engine in the same order, so a store's segment number is the same
//     everywhere and equals its group number (no translation table);
//   * placement decides where a document's rows go: RF_PLACE_STORE keeps whole stores on one device
//     (store g -> device g % G: a store-scoped query touches exactly one GPU, the multi-tenant layout),
//     RF_PLACE_SPREAD sends each document to the device holding the fewest rows (one huge store ends up
//     sharded by chunk over all GPUs, the 100 M-chunk layout);
//   * chunk ids are global: engine d numbers its rows from id_bases[d];
//   * a search is launched on every device that holds rows of the scope (rf_search_begin: the launches are
//     enqueued back to back from the calling thread and run concurrently on their GPUs), then collected
//     (rf_search_end: each device's top-k lands in mapped pinned host memory) and merged on the host --
//     G sorted lists of k keys under the RF-1 order (score desc, chunk id asc), which is a total order, so
//     the merged list equals the single-engine answer bit for bit.  The results have to reach the host
//     anyway; merging <= 8 x k keys there costs less than one NVLink round trip, needs no peer access and
//     no collective.  (The SPMD device-resident path -- one process per GPU, results staying in HBM --
//     exchanges inside the scan kernel instead: rf_search_keys_device_fused.)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
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

thread_local char g_gerr[512] = "";

int gfail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int gfail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_gerr, sizeof g_gerr, fmt, ap);
    va_end(ap);
    return code;
}

// One launcher thread per device: the calling thread hands every participating device its "begin" (query
// upload + kernel launch, a few microseconds of driver calls each) and they run side by side instead of one
// after the other -- with 8 GPUs and a batch that keeps each of them busy for only ~50 us, launching in turn
// would cost more than the scan.  A launcher spins for a short while after its last task (a serving loop keeps
// it hot) and sleeps otherwise; a caller that finds the launchers asleep launches in turn itself and wakes
// them for the next call.
struct Launcher {
    std::thread th;
    std::mutex owner;                    // one caller at a time, from post to collect
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint32_t> state{0};      // 0 idle, 1 task posted, 2 task done
    std::atomic<bool> asleep{true};
    std::atomic<bool> kick{false};       // "wake up and stay hot for a while": a caller expects to come back soon
    bool stop = false;
    void (*fn)(void *) = nullptr;
    void *arg = nullptr;

    void run() {
        using clock = std::chrono::steady_clock;
        auto last = clock::now();
        while (true) {
            if (state.load(std::memory_order_acquire) == 1) {
                fn(arg);
                state.store(2, std::memory_order_release);
                last = clock::now();
                continue;
            }
            if (clock::now() - last < std::chrono::microseconds(300)) continue;     // stay hot between back-to-back calls
            std::unique_lock<std::mutex> lk(mu);
            asleep.store(true, std::memory_order_release);
            cv.wait(lk, [&] { return stop || kick.load(std::memory_order_acquire) || state.load(std::memory_order_acquire) == 1; });
            kick.store(false, std::memory_order_release);
            asleep.store(false, std::memory_order_release);
            if (stop) return;
            last = clock::now();
        }
    }
    void wake() {
        std::lock_guard<std::mutex> lk(mu);
        kick.store(true, std::memory_order_release);
        cv.notify_one();
    }
    void post(void (*f)(void *), void *a) {
        fn = f;
        arg = a;
        state.store(1, std::memory_order_release);
        if (asleep.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> lk(mu);
            cv.notify_one();
        }
    }
    void collect() {
        while (state.load(std::memory_order_acquire) != 2) {}
        state.store(0, std::memory_order_release);
    }
};

struct Hit {
    uint64_t id;
    int32_t score;
    float cos;
};
// RF-1 order: score descending, chunk id ascending
inline bool before(const Hit &a, const Hit &b) { return a.score != b.score ? a.score > b.score : a.id < b.id; }

}  // namespace

struct rf_group {
    std::vector<rf_engine *> eng;
    std::vector<int> devices;
    std::vector<uint64_t> id_base;
    uint32_t placement = RF_PLACE_STORE;
    uint32_t dim = RF_DIM;
    uint64_t capacity_rows = 0;

    std::shared_mutex mu;                              // store table, row counts, document map
    std::vector<std::string> names;                    // group store number -> name ("" once dropped)
    std::vector<uint8_t> dropped;
    std::unordered_map<std::string, uint32_t> by_name;
    std::vector<std::vector<uint64_t>> rows_on;        // [device][store]: rows ever placed there (never decremented: conservative)
    std::vector<uint64_t> dev_rows;                    // rows placed per device (placement balance)
    std::unordered_map<uint64_t, uint32_t> doc_dev;    // document -> device
    std::vector<Launcher *> launchers;                 // one per device (groups of two or more)
};

namespace {

// Devices that may hold rows of a scope (bit d set).
uint32_t devices_of(const rf_group *g, const uint32_t *stores, uint32_t n) {
    uint32_t mask = 0;
    const uint32_t G = static_cast<uint32_t>(g->eng.size());
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = stores[i];
        if (s >= g->names.size() || g->dropped[s]) continue;
        for (uint32_t d = 0; d < G; ++d)
            if (g->rows_on[d][s]) mask |= 1u << d;
    }
    return mask;
}

// Merge `src` (n sorted hits) into `dst` (m sorted hits), keeping the best k.  Equal (score, id) pairs --
// impossible across devices, chunk ids are disjoint -- would be kept once.
uint32_t merge_into(Hit *dst, uint32_t m, const Hit *src, uint32_t n, uint32_t k, Hit *tmp) {
    uint32_t i = 0, j = 0, o = 0;
    while (o < k && (i < m || j < n)) {
        if (j >= n || (i < m && !before(src[j], dst[i]))) {
            if (i < m && j < n && dst[i].id == src[j].id && dst[i].score == src[j].score) ++j;
            tmp[o++] = dst[i++];
        } else {
            tmp[o++] = src[j++];
        }
    }
    memcpy(dst, tmp, o * sizeof(Hit));
    return o;
}

}  // namespace

extern "C" {

const char *rf_group_last_error(void) { return g_gerr[0] ? g_gerr : rf_last_error(); }

int rf_group_create(const rf_group_config *cfg, rf_group **out) {
    g_gerr[0] = 0;
    if (!cfg || !out) return gfail(RF_EINVAL, "null argument");
    if (cfg->struct_size != sizeof(rf_group_config)) return gfail(RF_EINVAL, "rf_group_config.struct_size %u != %zu", cfg->struct_size, sizeof(rf_group_config));
    if (cfg->n_devices == 0 || cfg->n_devices > RF_GROUP_MAX || !cfg->devices) return gfail(RF_EINVAL, "a group has 1 to %u devices", RF_GROUP_MAX);
    if (cfg->placement != RF_PLACE_STORE && cfg->placement != RF_PLACE_SPREAD) return gfail(RF_EINVAL, "unknown placement %u", cfg->placement);
    const uint32_t G = cfg->n_devices;
    const uint64_t stride = 0xFFFFFFFEull / G;
    if (!cfg->id_bases && cfg->capacity_rows > stride) return gfail(RF_EINVAL, "capacity_rows %llu exceeds the %llu chunk ids a device of a %u-device group can number", (unsigned long long)cfg->capacity_rows, (unsigned long long)stride, G);
    rf_group *g = new (std::nothrow) rf_group();
    if (!g) return gfail(RF_ENOMEM, "host allocation failed");
    g->placement = cfg->placement;
    g->dim = cfg->dim ? cfg->dim : RF_DIM;
    g->capacity_rows = cfg->capacity_rows;
    for (uint32_t d = 0; d < G; ++d) {
        rf_config ec{};
        ec.struct_size = sizeof(rf_config);
        ec.device = cfg->devices[d];
        ec.dim = cfg->dim ? cfg->dim : RF_DIM;
        ec.n_contexts = cfg->n_contexts;
        ec.capacity_rows = cfg->capacity_rows;
        ec.id_base = cfg->id_bases ? cfg->id_bases[d] : d * stride;
        rf_engine *e = nullptr;
        const int rc = rf_engine_create(&ec, &e);
        if (rc) {
            snprintf(g_gerr, sizeof g_gerr, "engine %u (device %d): %s", d, cfg->devices[d], rf_last_error());
            rf_group_destroy(g);
            return rc;
        }
        g->eng.push_back(e);
        g->devices.push_back(cfg->devices[d]);
        g->id_base.push_back(ec.id_base);
    }
    // id ranges must not overlap (a chunk id names one row of one device)
    for (uint32_t a = 0; a < G; ++a)
        for (uint32_t b = a + 1; b < G; ++b) {
            const uint64_t a0 = g->id_base[a], a1 = a0 + cfg->capacity_rows, b0 = g->id_base[b], b1 = b0 + cfg->capacity_rows;
            if (a0 < b1 && b0 < a1) {
                rf_group_destroy(g);
                return gfail(RF_EINVAL, "chunk id ranges of engines %u and %u overlap", a, b);
            }
        }
    g->rows_on.assign(G, {});
    g->dev_rows.assign(G, 0);
    if (G >= 2 && !getenv("RF_GROUP_SERIAL"))
        for (uint32_t d = 0; d < G; ++d) {
            Launcher *l = new (std::nothrow) Launcher();
            if (!l) break;
            l->th = std::thread([l] { l->run(); });
            g->launchers.push_back(l);
        }
    *out = g;
    return RF_OK;
}

int rf_group_destroy(rf_group *g) {
    if (!g) return RF_OK;
    for (Launcher *l : g->launchers) {
        {
            std::lock_guard<std::mutex> lk(l->mu);
            l->stop = true;
        }
        l->cv.notify_one();
        l->th.join();
        delete l;
    }
    for (rf_engine *e : g->eng) rf_engine_destroy(e);
    delete g;
    return RF_OK;
}

int rf_group_size(rf_group *g, uint32_t *n_devices) {
    if (!g || !n_devices) return gfail(RF_EINVAL, "null argument");
    *n_devices = static_cast<uint32_t>(g->eng.size());
    return RF_OK;
}

int rf_group_engine(rf_group *g, uint32_t index, rf_engine **out) {
    if (!g || !out || index >= g->eng.size()) return gfail(RF_EINVAL, "engine index out of range");
    *out = g->eng[index];
    return RF_OK;
}

int rf_group_stats(rf_group *g, rf_stats *total, rf_stats *per_device) {
    g_gerr[0] = 0;
    if (!g || !total) return gfail(RF_EINVAL, "null argument");
    memset(total, 0, sizeof *total);
    for (size_t d = 0; d < g->eng.size(); ++d) {
        rf_stats st{};
        const int rc = rf_engine_stats(g->eng[d], &st);
        if (rc) return rc;
        if (per_device) per_device[d] = st;
        total->n_rows += st.n_rows;
        total->capacity_rows += st.capacity_rows;
        total->n_docs += st.n_docs;
        total->hbm_bytes += st.hbm_bytes;
        total->searches += st.searches;
        total->kernel_launches += st.kernel_launches;
        total->free_rows += st.free_rows;
        total->ingest_bytes += st.ingest_bytes;
        total->ingest_kernel_ns += st.ingest_kernel_ns;
    }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    for (uint8_t dr : g->dropped) total->n_stores += dr ? 0 : 1;
    return RF_OK;
}

int rf_group_store_open(rf_group *g, const char *fs_name, uint32_t *store) {
    g_gerr[0] = 0;
    if (!g || !fs_name || !store) return gfail(RF_EINVAL, "null argument");
    std::unique_lock<std::shared_mutex> lk(g->mu);
    auto it = g->by_name.find(fs_name);
    if (it != g->by_name.end()) { *store = it->second; return RF_OK; }
    const uint32_t s = static_cast<uint32_t>(g->names.size());
    for (size_t d = 0; d < g->eng.size(); ++d) {
        uint32_t seg = 0;
        const int rc = rf_store_open(g->eng[d], fs_name, &seg);
        if (rc) return rc;
        if (seg != s) return gfail(RF_EINVAL, "internal: engine %zu numbered store '%s' %u, the group %u", d, fs_name, seg, s);
    }
    g->names.emplace_back(fs_name);
    g->dropped.push_back(0);
    g->by_name.emplace(fs_name, s);
    for (auto &v : g->rows_on) v.push_back(0);
    *store = s;
    return RF_OK;
}

int rf_group_store_lookup(rf_group *g, const char *fs_name, uint32_t *store) {
    g_gerr[0] = 0;
    if (!g || !fs_name || !store) return gfail(RF_EINVAL, "null argument");
    std::shared_lock<std::shared_mutex> lk(g->mu);
    auto it = g->by_name.find(fs_name);
    if (it == g->by_name.end()) return gfail(RF_ENOTFOUND, "unknown store '%s'", fs_name);
    *store = it->second;
    return RF_OK;
}

int rf_group_store_drop(rf_group *g, uint32_t store) {
    g_gerr[0] = 0;
    if (!g) return gfail(RF_EINVAL, "null argument");
    {
        std::unique_lock<std::shared_mutex> lk(g->mu);
        if (store >= g->names.size() || g->dropped[store]) return gfail(RF_ENOTFOUND, "unknown store %u", store);
        g->dropped[store] = 1;
        g->by_name.erase(g->names[store]);
        for (auto &v : g->rows_on) v[store] = 0;
    }
    int first_err = RF_OK;
    for (rf_engine *e : g->eng) {
        const int rc = rf_store_drop(e, store);
        if (rc && first_err == RF_OK) first_err = rc;
    }
    return first_err;
}

// Device of the next document of `store` (caller holds the unique lock).
static uint32_t place(rf_group *g, uint32_t store) {
    const uint32_t G = static_cast<uint32_t>(g->eng.size());
    if (g->placement == RF_PLACE_STORE) return store % G;
    uint32_t best = 0;
    for (uint32_t d = 1; d < G; ++d)
        if (g->dev_rows[d] < g->dev_rows[best]) best = d;
    return best;
}

int rf_group_ingest_text(rf_group *g, uint32_t store, uint64_t doc_id, const uint8_t *utf8, size_t n, uint64_t *first_chunk,
                         uint32_t *n_chunks, int64_t *spans, uint32_t max_spans) {
    g_gerr[0] = 0;
    if (!g) return gfail(RF_EINVAL, "null argument");
    uint32_t d;
    {
        std::unique_lock<std::shared_mutex> lk(g->mu);
        if (store >= g->names.size() || g->dropped[store]) return gfail(RF_ENOTFOUND, "unknown store %u", store);
        d = place(g, store);
        g->dev_rows[d] += n / 640 + 1;      // provisional (~ bytes per chunk of running text), corrected below: concurrent uploads spread out
    }
    uint32_t nc = 0;
    const int rc = rf_ingest_text(g->eng[d], store, doc_id, utf8, n, first_chunk, &nc, spans, max_spans);
    {
        std::unique_lock<std::shared_mutex> lk(g->mu);
        g->dev_rows[d] -= n / 640 + 1;
        if (rc == RF_OK) {
            g->dev_rows[d] += nc;
            if (store < g->names.size() && !g->dropped[store]) g->rows_on[d][store] += nc;
            g->doc_dev[doc_id] = d;
        }
    }
    if (n_chunks) *n_chunks = nc;
    return rc;
}

int rf_group_ingest_features(rf_group *g, uint32_t store, uint64_t doc_id, const int8_t *rows, uint64_t n_rows, uint64_t *first_chunk) {
    g_gerr[0] = 0;
    if (!g) return gfail(RF_EINVAL, "null argument");
    uint32_t d;
    {
        std::unique_lock<std::shared_mutex> lk(g->mu);
        if (store >= g->names.size() || g->dropped[store]) return gfail(RF_ENOTFOUND, "unknown store %u", store);
        d = place(g, store);
        g->dev_rows[d] += n_rows;
    }
    const int rc = rf_ingest_features(g->eng[d], store, doc_id, rows, n_rows, 0, first_chunk);
    std::unique_lock<std::shared_mutex> lk(g->mu);
    if (rc != RF_OK) {
        g->dev_rows[d] -= n_rows;
        return rc;
    }
    if (store < g->names.size() && !g->dropped[store]) g->rows_on[d][store] += n_rows;
    g->doc_dev[doc_id] = d;
    return RF_OK;
}

int rf_group_ingest_synthetic(rf_group *g, uint32_t first_store, uint64_t rows_per_store, uint64_t seed, uint64_t start_counter,
                              uint64_t n_rows, const uint16_t *zipf_vocab) {
    g_gerr[0] = 0;
    if (!g || !zipf_vocab) return gfail(RF_EINVAL, "null argument");
    const uint32_t G = static_cast<uint32_t>(g->eng.size());
    const uint64_t n_stores = rows_per_store ? (n_rows + rows_per_store - 1) / rows_per_store : 1;
    {
        std::shared_lock<std::shared_mutex> lk(g->mu);
        if (first_store + n_stores > g->names.size()) return gfail(RF_ENOTFOUND, "stores %u..%llu are not all open", first_store, (unsigned long long)(first_store + n_stores - 1));
    }
    if (g->placement == RF_PLACE_SPREAD) {
        // one corpus cut into G contiguous runs of counters (the chunk-sharded layout): device d generates
        // rows [lo_d, hi_d); with id_bases[d] = start of its run the chunk ids equal a single engine's
        if (rows_per_store) return gfail(RF_EINVAL, "spread placement generates one store at a time (rows_per_store = 0)");
        for (uint32_t d = 0; d < G; ++d) {
            const uint64_t base = n_rows / G, rem = n_rows % G;
            const uint64_t lo = d * base + std::min<uint64_t>(d, rem), cnt = base + (d < rem ? 1 : 0);
            if (!cnt) continue;
            const int rc = rf_ingest_synthetic(g->eng[d], first_store, 0, seed, start_counter + lo, cnt, zipf_vocab, nullptr);
            if (rc) return rc;
            std::unique_lock<std::shared_mutex> lk(g->mu);
            g->rows_on[d][first_store] += cnt;
            g->dev_rows[d] += cnt;
        }
        return RF_OK;
    }
    // whole stores per device: store first_store + i (rows [i * rows_per_store, ...) of the corpus) on device (first_store + i) % G;
    // consecutive stores of one device are generated with one launch each (their counters are not contiguous)
    for (uint64_t i = 0; i < n_stores; ++i) {
        const uint32_t s = first_store + static_cast<uint32_t>(i);
        const uint32_t d = s % G;
        const uint64_t lo = i * (rows_per_store ? rows_per_store : n_rows);
        const uint64_t cnt = std::min(n_rows - lo, rows_per_store ? rows_per_store : n_rows);
        const int rc = rf_ingest_synthetic(g->eng[d], s, 0, seed, start_counter + lo, cnt, zipf_vocab, nullptr);
        if (rc) return rc;
        std::unique_lock<std::shared_mutex> lk(g->mu);
        g->rows_on[d][s] += cnt;
        g->dev_rows[d] += cnt;
    }
    return RF_OK;
}

int rf_group_doc_tombstone(rf_group *g, uint64_t doc_id) {
    g_gerr[0] = 0;
    if (!g) return gfail(RF_EINVAL, "null argument");
    uint32_t d;
    {
        std::unique_lock<std::shared_mutex> lk(g->mu);
        auto it = g->doc_dev.find(doc_id);
        if (it == g->doc_dev.end()) return gfail(RF_ENOTFOUND, "unknown document %llu", (unsigned long long)doc_id);
        d = it->second;
        g->doc_dev.erase(it);
    }
    return rf_doc_tombstone(g->eng[d], doc_id);
}

int rf_group_search(rf_group *g, const int8_t *q, uint32_t nq, const uint32_t *stores, const uint32_t *store_off, uint32_t k,
                    uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts) {
    g_gerr[0] = 0;
    if (!g || !q || !store_off || !out_ids || !out_scores) return gfail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return gfail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (nq == 0) return RF_OK;
    if (nq > 65535) return gfail(RF_EINVAL, "at most 65535 queries per call");
    const uint32_t G = static_cast<uint32_t>(g->eng.size());
    // which queries go to which device
    std::vector<uint32_t> qmask(nq);
    uint32_t any = 0;
    {
        std::shared_lock<std::shared_mutex> lk(g->mu);
        for (uint32_t i = 0; i < nq; ++i) {
            const uint32_t s0 = store_off[i], s1 = store_off[i + 1];
            if (s1 < s0 || s1 - s0 > RF_SCOPE_MAX) return gfail(RF_EINVAL, "scope of query %u has %u stores (max %u)", i, s1 - s0, RF_SCOPE_MAX);
            qmask[i] = devices_of(g, stores + s0, s1 - s0);
            any |= qmask[i];
        }
    }
    struct PerDev {
        rf_pending *p = nullptr;
        std::vector<uint32_t> qidx, segs, off;
        std::vector<int8_t> qrows;
        std::vector<uint64_t> ids;
        std::vector<int32_t> sc;
        std::vector<float> cs;
        std::vector<uint32_t> cnt;
    };
    std::vector<PerDev> pd(G);
    int rc = RF_OK;
    // ---- launch on every device that holds rows of some query's scope (concurrent on the GPUs).  Each device
    // gets only the queries that have rows there: its share of the query rows, compacted, and their scopes.
    struct Begin {
        rf_group *g; PerDev *w; uint32_t d, nq, k;
        const int8_t *q; const uint32_t *stores, *store_off; const uint32_t *qmask;
        int rc;
        char err[256];
    };
    auto begin_fn = [](void *p) {
        Begin &b = *static_cast<Begin *>(p);
        PerDev &w = *b.w;
        w.off.push_back(0);
        for (uint32_t i = 0; i < b.nq; ++i) {
            if (!(b.qmask[i] & (1u << b.d))) continue;
            w.qidx.push_back(i);
            w.segs.insert(w.segs.end(), b.stores + b.store_off[i], b.stores + b.store_off[i + 1]);
            w.off.push_back(static_cast<uint32_t>(w.segs.size()));
        }
        const uint32_t n_here = static_cast<uint32_t>(w.qidx.size());
        const int8_t *qd = b.q;
        if (n_here != b.nq) {                       // compact this device's queries
            const size_t dim = b.g->dim;
            w.qrows.resize(static_cast<size_t>(n_here) * dim);
            for (uint32_t j = 0; j < n_here; ++j) memcpy(w.qrows.data() + static_cast<size_t>(j) * dim, b.q + static_cast<size_t>(w.qidx[j]) * dim, dim);
            qd = w.qrows.data();
        }
        if (w.segs.empty()) w.segs.push_back(0);
        b.rc = rf_search_begin(b.g->eng[b.d], qd, n_here, w.segs.data(), w.off.data(), b.k, &w.p);
        if (b.rc) snprintf(b.err, sizeof b.err, "%s", rf_last_error());      // (the error text is thread-local to the launcher)
    };
    std::vector<Begin> begins(G);
    uint32_t n_part = 0;
    for (uint32_t d = 0; d < G; ++d) {
        begins[d] = Begin{g, &pd[d], d, nq, k, q, stores, store_off, qmask.data(), RF_OK, {0}};
        n_part += (any >> d) & 1u;
    }
    // side by side on the launcher threads when they are awake (or the batch is big enough to be worth waking
    // them); in turn on this thread otherwise
    bool parallel = n_part >= 2 && !g->launchers.empty();
    if (parallel && nq < 8)
        for (uint32_t d = 0; d < G && parallel; ++d)
            if ((any >> d) & 1u) parallel = !g->launchers[d]->asleep.load(std::memory_order_acquire);
    if (parallel) {
        for (uint32_t d = 0; d < G; ++d)
            if ((any >> d) & 1u) {
                g->launchers[d]->owner.lock();
                g->launchers[d]->post(begin_fn, &begins[d]);
            }
        for (uint32_t d = 0; d < G; ++d)
            if ((any >> d) & 1u) {
                g->launchers[d]->collect();
                g->launchers[d]->owner.unlock();
            }
    } else {
        for (uint32_t d = 0; d < G; ++d)
            if ((any >> d) & 1u) {
                begin_fn(&begins[d]);
                if (begins[d].rc) break;
            }
        if (n_part >= 2)      // a serving loop's next call finds the launchers awake
            for (uint32_t d = 0; d < G && !g->launchers.empty(); ++d)
                if (((any >> d) & 1u) && g->launchers[d]->asleep.load(std::memory_order_acquire)) g->launchers[d]->wake();
    }
    for (uint32_t d = 0; d < G; ++d)
        if (begins[d].rc && rc == RF_OK) {
            rc = begins[d].rc;
            snprintf(g_gerr, sizeof g_gerr, "device %u: %s", d, begins[d].err[0] ? begins[d].err : rf_last_error());
        }
    // ---- collect (every begun search must be ended, also after a failure) and merge on the host
    const size_t nk = static_cast<size_t>(nq) * k;
    std::vector<Hit> best(nk), tmp(k), src(k);
    std::vector<uint32_t> have(nq, 0);
    for (uint32_t d = 0; d < G; ++d) {
        PerDev &w = pd[d];
        if (!w.p) continue;
        const uint32_t n_here = static_cast<uint32_t>(w.qidx.size());
        w.ids.resize(static_cast<size_t>(n_here) * k);
        w.sc.resize(static_cast<size_t>(n_here) * k);
        w.cs.resize(static_cast<size_t>(n_here) * k);
        w.cnt.resize(n_here);
        const int rc2 = rf_search_end(g->eng[d], w.p, w.ids.data(), w.sc.data(), w.cs.data(), w.cnt.data(), nullptr);
        if (rc2) { if (rc == RF_OK) rc = rc2; continue; }
        for (uint32_t j = 0; j < n_here; ++j) {
            const uint32_t i = w.qidx[j], c = std::min(w.cnt[j], k);
            for (uint32_t t = 0; t < c; ++t) src[t] = Hit{w.ids[static_cast<size_t>(j) * k + t], w.sc[static_cast<size_t>(j) * k + t], w.cs[static_cast<size_t>(j) * k + t]};
            have[i] = merge_into(best.data() + static_cast<size_t>(i) * k, have[i], src.data(), c, k, tmp.data());
        }
    }
    if (rc) return rc;
    for (uint32_t i = 0; i < nq; ++i) {
        for (uint32_t t = 0; t < k; ++t) {
            const size_t o = static_cast<size_t>(i) * k + t;
            const bool ok = t < have[i];
            out_ids[o] = ok ? best[o].id : ~0ull;
            out_scores[o] = ok ? best[o].score : 0;
            if (out_cos) out_cos[o] = ok ? best[o].cos : 0.0f;
        }
        if (out_counts) out_counts[i] = have[i];
    }
    return RF_OK;
}

int rf_group_search_text(rf_group *g, const uint8_t *utf8, size_t n, const uint32_t *stores, uint32_t n_stores,
                         const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights, uint32_t k, uint64_t *out_ids,
                         int32_t *out_scores, float *out_cos, uint32_t *out_count, int8_t *out_q) {
    g_gerr[0] = 0;
    if (!g || (!utf8 && n) || !out_ids || !out_scores || (!stores && n_stores)) return gfail(RF_EINVAL, "null argument");
    if (k == 0 || k > RF_TOPK_MAX) return gfail(RF_EINVAL, "k must be in [1, %u]", RF_TOPK_MAX);
    if (n_stores > RF_SCOPE_MAX) return gfail(RF_EINVAL, "scope has %u stores (max %u)", n_stores, RF_SCOPE_MAX);
    const uint32_t G = static_cast<uint32_t>(g->eng.size());
    uint32_t mask;
    {
        std::shared_lock<std::shared_mutex> lk(g->mu);
        mask = devices_of(g, stores, n_stores);
    }
    if (!mask) mask = 1u;                      // nothing in scope anywhere: device 0 still turns the text into the query vector
    std::vector<rf_pending *> pend(G, nullptr);
    int rc = RF_OK;
    for (uint32_t d = 0; d < G && rc == RF_OK; ++d)
        if (mask & (1u << d))
            rc = rf_search_text_begin(g->eng[d], utf8, n, stores, n_stores, ranges, n_ranges, weights, k, &pend[d]);   // every device hashes the text itself
    std::vector<Hit> best(k), tmp(k), src(k);
    uint32_t have = 0;
    bool q_done = false;
    std::vector<uint64_t> ids(k);
    std::vector<int32_t> sc(k);
    std::vector<float> cs(k);
    for (uint32_t d = 0; d < G; ++d) {
        if (!pend[d]) continue;
        uint32_t cnt = 0;
        const int rc2 = rf_search_end(g->eng[d], pend[d], ids.data(), sc.data(), cs.data(), &cnt, (out_q && !q_done) ? out_q : nullptr);
        if (rc2) { if (rc == RF_OK) rc = rc2; continue; }
        q_done = true;
        cnt = std::min(cnt, k);
        for (uint32_t t = 0; t < cnt; ++t) src[t] = Hit{ids[t], sc[t], cs[t]};
        have = merge_into(best.data(), have, src.data(), cnt, k, tmp.data());
    }
    if (rc) return rc;
    for (uint32_t t = 0; t < k; ++t) {
        const bool ok = t < have;
        out_ids[t] = ok ? best[t].id : ~0ull;
        out_scores[t] = ok ? best[t].score : 0;
        if (out_cos) out_cos[t] = ok ? best[t].cos : 0.0f;
    }
    if (out_count) *out_count = have;
    return RF_OK;
}

int rf_group_scope_df(rf_group *g, const uint32_t *stores, uint32_t n_stores, uint64_t *out_df, uint64_t *out_n) {
    g_gerr[0] = 0;
    if (!g || (!stores && n_stores) || !out_df || !out_n) return gfail(RF_EINVAL, "null argument");
    // the corpus statistic of RF-1w is a sum over rows: add the devices' counts (each cached per scope)
    memset(out_df, 0, static_cast<size_t>(g->dim) * 8);
    *out_n = 0;
    uint64_t df[RF_DIM_MAX];
    for (rf_engine *e : g->eng) {
        uint64_t nn = 0;
        const int rc = rf_scope_df(e, stores, n_stores, df, &nn);
        if (rc) return rc;
        for (uint32_t i = 0; i < g->dim; ++i) out_df[i] += df[i];
        *out_n += nn;
    }
    return RF_OK;
}

// ---- durability: one engine snapshot per device ("<path>.dev<d>") and the group's own table ("<path>.group")
namespace {
struct GroupHeader {
    char magic[8];
    uint32_t n_devices, placement;
    uint64_t n_stores, n_docs;
};
}  // namespace

int rf_group_snapshot_save(rf_group *g, const char *path) {
    g_gerr[0] = 0;
    if (!g || !path) return gfail(RF_EINVAL, "null argument");
    std::shared_lock<std::shared_mutex> lk(g->mu);
    for (size_t d = 0; d < g->eng.size(); ++d) {
        const std::string p = std::string(path) + ".dev" + std::to_string(d);
        const int rc = rf_snapshot_save(g->eng[d], p.c_str());
        if (rc) return rc;
    }
    const std::string tmp = std::string(path) + ".group.tmp", fin = std::string(path) + ".group";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return gfail(RF_EINVAL, "cannot open %s for writing", tmp.c_str());
    GroupHeader h{};
    memcpy(h.magic, "RFB2GRP1", 8);
    h.n_devices = static_cast<uint32_t>(g->eng.size());
    h.placement = g->placement;
    h.n_stores = g->names.size();
    h.n_docs = g->doc_dev.size();
    bool ok = fwrite(&h, 1, sizeof h, f) == sizeof h;
    for (size_t s = 0; ok && s < g->names.size(); ++s) {
        const uint32_t len = static_cast<uint32_t>(g->names[s].size()), dr = g->dropped[s];
        ok = fwrite(&len, 1, 4, f) == 4 && (len == 0 || fwrite(g->names[s].data(), 1, len, f) == len) && fwrite(&dr, 1, 4, f) == 4;
        for (size_t d = 0; ok && d < g->eng.size(); ++d) ok = fwrite(&g->rows_on[d][s], 1, 8, f) == 8;
    }
    for (const auto &kv : g->doc_dev) {
        if (!ok) break;
        ok = fwrite(&kv.first, 1, 8, f) == 8 && fwrite(&kv.second, 1, 4, f) == 4;
    }
    for (size_t d = 0; ok && d < g->eng.size(); ++d) ok = fwrite(&g->dev_rows[d], 1, 8, f) == 8;
    ok = ok && fflush(f) == 0 && fsync(fileno(f)) == 0;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), fin.c_str()) != 0) {
        unlink(tmp.c_str());
        return gfail(RF_EINVAL, "write of %s failed", fin.c_str());
    }
    return RF_OK;
}

int rf_group_snapshot_load(rf_group *g, const char *path) {
    g_gerr[0] = 0;
    if (!g || !path) return gfail(RF_EINVAL, "null argument");
    std::unique_lock<std::shared_mutex> lk(g->mu);
    if (!g->names.empty()) return gfail(RF_EINVAL, "snapshots load into an empty group");
    const std::string fin = std::string(path) + ".group";
    FILE *f = fopen(fin.c_str(), "rb");
    if (!f) return gfail(RF_ENOTFOUND, "cannot open %s", fin.c_str());
    GroupHeader h{};
    bool ok = fread(&h, 1, sizeof h, f) == sizeof h && memcmp(h.magic, "RFB2GRP1", 8) == 0;
    if (ok && (h.n_devices != g->eng.size() || h.placement != g->placement)) {
        fclose(f);
        return gfail(RF_EINVAL, "%s was written by a group of %u devices with placement %u", fin.c_str(), h.n_devices, h.placement);
    }
    ok = ok && h.n_stores < (1ull << 32) && h.n_docs < (1ull << 40);
    std::vector<std::string> names;
    std::vector<uint8_t> dropped;
    std::vector<std::vector<uint64_t>> rows_on(g->eng.size());
    for (uint64_t s = 0; ok && s < h.n_stores; ++s) {
        uint32_t len = 0, dr = 0;
        ok = fread(&len, 1, 4, f) == 4 && len < (1u << 20);
        std::string nm(ok ? len : 0, '\0');
        ok = ok && (len == 0 || fread(&nm[0], 1, len, f) == len) && fread(&dr, 1, 4, f) == 4;
        names.push_back(nm);
        dropped.push_back(dr ? 1 : 0);
        for (size_t d = 0; ok && d < g->eng.size(); ++d) {
            uint64_t r = 0;
            ok = fread(&r, 1, 8, f) == 8;
            rows_on[d].push_back(r);
        }
    }
    std::unordered_map<uint64_t, uint32_t> doc_dev;
    for (uint64_t i = 0; ok && i < h.n_docs; ++i) {
        uint64_t id = 0;
        uint32_t d = 0;
        ok = fread(&id, 1, 8, f) == 8 && fread(&d, 1, 4, f) == 4 && d < g->eng.size();
        doc_dev[id] = d;
    }
    std::vector<uint64_t> dev_rows(g->eng.size(), 0);
    for (size_t d = 0; ok && d < g->eng.size(); ++d) ok = fread(&dev_rows[d], 1, 8, f) == 8;
    fclose(f);
    if (!ok) return gfail(RF_EINVAL, "%s is truncated or corrupt", fin.c_str());
    for (size_t d = 0; d < g->eng.size(); ++d) {
        const std::string p = std::string(path) + ".dev" + std::to_string(d);
        const int rc = rf_snapshot_load(g->eng[d], p.c_str());
        if (rc) return rc;
    }
    g->names.swap(names);
    g->dropped.swap(dropped);
    g->rows_on.swap(rows_on);
    g->doc_dev.swap(doc_dev);
    g->dev_rows.swap(dev_rows);
    g->by_name.clear();
    for (uint32_t s = 0; s < g->names.size(); ++s)
        if (!g->dropped[s]) g->by_name.emplace(g->names[s], s);
    return RF_OK;
}

}  // extern "C"
