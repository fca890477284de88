/* rf_b200.h -- C-ABI of the B200-native store-scoped chunk retriever (librf_b200.so).
 *
 * This is the drop-in boundary UNDER the Python adapter (rag_foundation_b200/adapter.py:B200Rag),
 * which itself duck-types the reference's adapter object:
 *   backend/app/services/gemini_rag.py:242-599 (GeminiRag), :602-718 (MockGeminiRag),
 *   :721-725 (get_rag_client).
 * The reference has no FFI for this path (it is pure Python calling a remote service), so each
 * entry point below cites the reference METHOD whose work it carries out on the GPU.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; 0 = RF_OK, negative = error
 * (rf_strerror); no exceptions cross the ABI; the caller owns every output buffer; all entry
 * points are thread-safe except rf_engine_create/destroy.  There is NO CPU fallback: every compute
 * entry point launches sm_100a kernels and fails with RF_ECUDA/RF_ENODEVICE when it cannot.
 *
 * Arithmetic is the frozen RF-1 spec (oracle/SPEC.md): D int8 hashed term counts per chunk (D = 256, or the
 * wider 512 / 1024 of SURVEY.md 8f-4: fewer bucket collisions, D + 4 bytes streamed per chunk), int32
 * dot-product scores, rank by (score desc, global chunk id asc), float32 cosine reported.  D is fixed per
 * engine (rf_config.dim); wherever a comment below says RF_DIM for a buffer size, read "the engine's dim".
 */
#ifndef RF_B200_H
#define RF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_DIM 256u           /* default feature dimension (int8 per chunk row)       */
#define RF_DIM_MAX 1024u      /* widest row an engine can be created with             */
#define RF_TOPK_MAX 32u       /* largest k a search may ask for (reference default 10) */
#define RF_SCOPE_MAX 16u      /* store segments per query scope (reference: <= 10 stores/user, config.py:127) */
#define RF_TOMBSTONE 0xFFFFFFFFu

enum {
    RF_OK = 0,
    RF_EINVAL = -1,     /* bad argument                                   -> RuntimeError   */
    RF_ENOMEM = -2,     /* HBM arena / host allocation exhausted          -> RuntimeError   */
    RF_ECUDA = -3,      /* a CUDA call or kernel failed                   -> RuntimeError   */
    RF_ENODEVICE = -4,  /* no sm_100 device visible                       -> RuntimeError   */
    RF_ENOTFOUND = -5,  /* unknown store / document                       -> RuntimeError   */
    RF_EBUSY = -6,      /* no free search context within the wait budget  -> TimeoutError (retryable, gemini_rag.py:22-27) */
    RF_ECAPACITY = -7   /* more rows than capacity_rows                   -> RuntimeError   */
};

typedef struct rf_engine rf_engine;

typedef struct rf_config {
    uint32_t struct_size;    /* sizeof(rf_config), for forward compatibility            */
    int32_t device;          /* CUDA ordinal                                            */
    uint32_t dim;            /* 256 (RF_DIM), 512 or 1024; batches take the tensor cores at every width */
    uint32_t n_contexts;     /* concurrent searches (reference: 50 streams/process, routes/chat.py:40); 0 -> 8 */
    uint64_t capacity_rows;  /* chunk rows reserved in HBM (dim + 8 B each)             */
    uint64_t id_base;        /* global chunk id of row 0 (shard offset in sharded mode) */
} rf_config;

typedef struct rf_stats {
    uint64_t n_rows;          /* rows appended so far (including tombstoned)  */
    uint64_t capacity_rows;
    uint64_t n_stores;
    uint64_t n_docs;
    uint64_t hbm_bytes;       /* device bytes held by the engine              */
    uint64_t searches;        /* queries answered                             */
    uint64_t kernel_launches; /* kernels launched by this engine since create */
    uint64_t free_rows;       /* rows below n_rows given back by deletes, waiting to be reused */
    uint64_t ingest_bytes;    /* text bytes featurised by rf_ingest_text                       */
    uint64_t ingest_kernel_ns;/* device time from the first tokeniser launch to the end of the rows kernel
                                 (CUDA events on the ingest stream), summed over those documents */
} rf_stats;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int rf_engine_create(const rf_config *cfg, rf_engine **out);
int rf_engine_destroy(rf_engine *e);
int rf_engine_stats(rf_engine *e, rf_stats *out);
const char *rf_strerror(int code);
const char *rf_last_error(void);           /* thread-local detail of the last failure */
int rf_build_info(char *buf, size_t n);    /* "sm_100a nvcc 12.9 ..." */

/* Diagnostics: with RF_SCAN_DEBUG=1 in the environment at rf_engine_create, the scan kernel stamps
 * %globaltimer at 6 points per block ([block][8] words); this copies them out (and clears them). */
int rf_debug_timestamps(rf_engine *e, uint64_t *out, uint64_t n_words, int clear);

/* ---- stores: GeminiRag.create_store / delete_store (gemini_rag.py:271-304, 610-612, 696-697) -- */
int rf_store_open(rf_engine *e, const char *fs_name, uint32_t *store_seg);   /* idempotent */
int rf_store_lookup(rf_engine *e, const char *fs_name, uint32_t *store_seg); /* RF_ENOTFOUND if absent */
int rf_store_drop(rf_engine *e, uint32_t store_seg);                         /* tombstones its rows */

/* ---- ingest: GeminiRag.upload_file (gemini_rag.py:307-352, 614-629), called by the ARQ worker at
 * services/ingestion.py:45-52.  Featurises on the GPU (tokenise, chunk, FNV-1a, int8 histogram)
 * and appends the rows to the store.  spans[2*i], spans[2*i+1] = byte span of chunk i in utf8
 * (first max_spans chunks).  Synchronous: rows are searchable when it returns (op_status done,
 * gemini_rag.py:631-638). */
int rf_ingest_text(rf_engine *e, uint32_t store_seg, uint64_t doc_id, const uint8_t *utf8, size_t n,
                   uint64_t *first_chunk, uint32_t *n_chunks, int64_t *spans, uint32_t max_spans);
/* `utf8` may live anywhere; the engine looks (cudaPointerGetAttributes) and takes the matching route:
 *   pageable host memory -- staged through the engine's own ring of pinned slots by helper threads, so host
 *                           memcpy, PCIe transfer and the tokenise kernels overlap chunk by chunk;
 *   pinned host memory   -- (rf_host_alloc, or registered by the caller) DMA straight from it, same pipeline;
 *   device memory        -- no transfer at all.
 * rf_host_alloc / rf_host_free hand out pinned buffers for hosts that want to read an upload straight into
 * DMA-able memory (the reference's worker reads a temp file, services/ingestion.py:45-52, uploads.py:245-261). */
int rf_host_alloc(size_t n, void **out);
int rf_host_free(void *p);

/* Append pre-computed int8 rows (host or device pointer, n_rows x RF_DIM). */
int rf_ingest_features(rf_engine *e, uint32_t store_seg, uint64_t doc_id, const int8_t *rows,
                       uint64_t n_rows, int rows_on_device, uint64_t *first_chunk);

/* Append n_rows rows of the RF-1 synthetic corpus `seed`, counters start_counter.. (oracle/SPEC.md),
 * generated on the GPU.  Row i goes to store segment first_seg + i / rows_per_store (all of which
 * must be open); rows_per_store = 0 puts every row in first_seg.  zipf_vocab: uint16[65536] host. */
int rf_ingest_synthetic(rf_engine *e, uint32_t first_seg, uint64_t rows_per_store, uint64_t seed,
                        uint64_t start_counter, uint64_t n_rows, const uint16_t *zipf_vocab,
                        uint64_t *first_chunk);

/* GeminiRag.delete_document_from_store (gemini_rag.py:354-424, 699-702).  The rows are masked at once
 * and given back to the arena: a later ingest that fits a freed run reuses those rows (and their chunk
 * ids), so a delete-heavy tenant does not exhaust capacity_rows.  rf_store_drop does the same for every
 * row of the store. */
int rf_doc_tombstone(rf_engine *e, uint64_t doc_id);

/* ---- durability: the HBM index is volatile and the reference deletes uploads
 * after ingest (services/ingestion.py:341), so the engine can write / read a snapshot file:
 * header {magic "RFB2SNP2", dim, version, n_rows, id_base, n_stores, n_docs, n_free}, store table (name,
 * dropped, extents), document table (id, store, extents), free list, then int8 rows, uint32 segment
 * words, int32 norms, and a trailer {magic, 64-bit checksum of everything before it}.  Save writes
 * "<path>.tmp", fsyncs and renames it over <path> (a failed save never damages the previous file);
 * load verifies the checksum and every extent before it publishes anything.
 * rf_snapshot_load needs a freshly created engine (no rows, no stores) with capacity >= n_rows and
 * the same id_base. */
int rf_snapshot_save(rf_engine *e, const char *path);
int rf_snapshot_load(rf_engine *e, const char *path);

/* ---- one HBM index, several processes: the reference runs four API worker processes and one
 * ingest worker (backend/Dockerfile:42, backend/app/worker.py:122-126), each calling get_rag_client()
 * (gemini_rag.py:721-725).  ONE process (the engine daemon, rag_foundation_b200/server.py) owns the arena.
 * rf_engine_export writes a description of it -- CUDA IPC handles of the row arrays, geometry, the store table
 * with its extents -- into buf (call with buf = NULL to learn *len).  ANOTHER process on the same GPU maps the
 * arena read-only with rf_engine_attach and runs every search entry point of this header on its own streams;
 * the entry points that change the index fail there with RF_EINVAL (rf_store_open finds existing stores).
 * Deletes by the owner are visible to readers at once (they mask rows of the shared arena); rows and stores
 * added later become visible after rf_engine_refresh with a newer export of the same engine.
 * rf_engine_destroy unmaps.  The owner must outlive its readers' searches. */
int rf_engine_export(rf_engine *e, void *buf, size_t cap, size_t *len);
int rf_engine_attach(const void *export_blob, size_t len, uint32_t n_contexts /* 0 -> 8 */, rf_engine **out);
int rf_engine_refresh(rf_engine *e, const void *export_blob, size_t len);

/* Copy rows [first, first+n) back to the host (tests / snapshots): any of the outputs may be NULL. */
int rf_rows_read(rf_engine *e, uint64_t first_row, uint64_t n, int8_t *rows, uint32_t *store_seg,
                 int32_t *ff);

/* ---- query: GeminiRag.ask_stream / ask retrieval step (gemini_rag.py:472-551, 656-694) --------
 * q: nq x RF_DIM int8 query vectors (HOST).  Scope of query i = store_segs[seg_off[i] .. seg_off[i+1])
 * (CSR, at most RF_SCOPE_MAX each).  Outputs are HOST buffers, nq x k, rows padded with id
 * UINT64_MAX / score 0 / cos 0 beyond out_counts[i].  out_cos / out_counts may be NULL.
 * Blocks until the results are in the caller's buffers.  One launch of the scan kernel (grid.y =
 * queries); a batch whose queries all carry the same scope is scored by the tensor-core kernels
 * instead when that is cheaper than one scan per query (same results). */
int rf_search(rf_engine *e, const int8_t *q, uint32_t nq, const uint32_t *store_segs,
              const uint32_t *seg_off, uint32_t k, uint64_t *out_ids, int32_t *out_scores,
              float *out_cos, uint32_t *out_counts);

/* The same search in two halves, for a host that wants several searches in flight from one thread (the
 * engine group launches one per GPU, then collects them): rf_search_begin validates, takes a search
 * context and enqueues the work; rf_search_end blocks until the results are in the caller's buffers and
 * ALWAYS returns the context to the pool (also when it fails).  Every successful begin must be followed by
 * exactly one end.  out_q (RF_DIM, may be NULL) receives the query vector of a text search. */
typedef struct rf_pending rf_pending;
int rf_search_begin(rf_engine *e, const int8_t *q, uint32_t nq, const uint32_t *store_segs,
                    const uint32_t *seg_off, uint32_t k, rf_pending **out);
int rf_search_text_begin(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs, uint32_t n_segs,
                         const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights, uint32_t k, rf_pending **out);
int rf_search_end(rf_engine *e, rf_pending *p, uint64_t *out_ids, int32_t *out_scores, float *out_cos,
                  uint32_t *out_counts, int8_t *out_q);

/* Same, from query text: tokenise + hash on the GPU, then search, no host round trip between.
 * (MockGeminiRag._contents_to_text picks the text, gemini_rag.py:640-654.) */
int rf_search_text(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs,
                   uint32_t n_segs, uint32_t k, uint64_t *out_ids, int32_t *out_scores,
                   float *out_cos, uint32_t *out_count, int8_t *out_q /* RF_DIM, may be NULL */);

/* rf_search_text restricted further to the global chunk id ranges [ranges[2i], ranges[2i+1])
 * (sorted, disjoint; n_ranges = 0 means no restriction).  Carries doc-level metadata filters
 * (ChatRequest.metadataFilter, routes/chat.py:295-335, forwarded as `metadata_filter` to
 * ask_stream): the host resolves the filter to the matching documents' chunk ranges, the kernel
 * scans the intersection of those ranges with the scoped stores' extents (row mask still applied). */
int rf_search_text_in(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs,
                      uint32_t n_segs, const uint64_t *ranges, uint32_t n_ranges, uint32_t k,
                      uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_count,
                      int8_t *out_q /* RF_DIM, may be NULL */);

/* Device-resident variant for pipelines and the sharded (multi-GPU) path: q_dev and out_keys_dev
 * are DEVICE pointers, the work is enqueued on `stream` (a cudaStream_t; NULL = legacy default
 * stream) and the call returns without synchronising.  out_keys_dev: nq x k packed RF-1 keys
 * ((uint64(score) << 32) | (0xFFFFFFFF - global_id)), 0 = no result.  All queries share one scope. */
int rf_search_keys_device(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                          uint32_t n_segs, uint32_t k, uint64_t *out_keys_dev, void *stream);

/* Overlap promise for the device-resident searches on `stream` (default: off).  With allow != 0 their scan
 * kernels are launched with programmatic stream serialisation: a search's scan phase may start while the
 * previous kernel on the stream is still in its merge tail (back-to-back queries then stream at the full
 * HBM rate).  The kernel reads its query vectors BEFORE it waits for that predecessor, so the caller
 * promises that no KERNEL enqueued on `stream` between two searches writes q_dev (copies are fine; a
 * resident query batch is the intended use).  Without the promise every search is a fully serialised
 * launch. */
int rf_stream_set_overlap(rf_engine *e, void *stream, int allow);

/* rf_search_keys_device with one scope PER QUERY (CSR, as rf_search): the store-sharded multi-GPU
 * path (whole stores per rank, SURVEY.md 8e / configs[4]) gives every rank the full query batch and
 * the part of each query's scope it owns -- possibly empty, then that query's keys are all 0. */
int rf_search_keys_device_scoped(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                 const uint32_t *seg_off, uint32_t k, uint64_t *out_keys_dev, void *stream);

/* Fused variant of the sharded search: local scan + top-k EXCHANGE in one pass over NVLink peer
 * memory, no collective launch.  Every rank passes the same exchange description: keys_ptrs[r] /
 * flag_ptrs[r] are this process's mappings of rank r's gather buffer ([4][world][nq_cap][k] u64)
 * and flag array ([4][world][nq_cap] u32, zero-initialised), e.g. from a symmetric-memory
 * rendezvous; `seq` must be the same on all ranks and increase by one per call (four buffers,
 * slot = seq % 4).  ONE kernel per call does both halves: the block that finishes a query stores its
 * k keys into every rank's buffer and releases a flag there, then acquires all ranks' flags in its
 * own buffer (the peers' kernels run concurrently on their GPUs; bounded wait: out_keys_dev is
 * zeroed and *timeout_flag_dev set to 1 if a peer does not arrive within ~2 s) and merges into
 * out_keys_dev [nq, k].  Collective semantics: all ranks must call it, on one stream each. */
typedef struct rf_peer_exchange {
    uint32_t struct_size;
    uint32_t rank, world;      /* world <= 8 */
    uint32_t nq_cap;           /* queries the buffers are sized for */
    uint32_t k;                /* keys per list the buffers are sized for (== k of the call) */
    uint32_t seq;              /* > 0, +1 per call, identical on every rank */
    const uint64_t *keys_ptrs; /* HOST array [world] of device pointers */
    const uint64_t *flag_ptrs; /* HOST array [world] of device pointers */
    uint32_t *timeout_flag_dev;
    /* rf_search_keys_device_scoped_fused only (zero / NULL elsewhere): the call carries just the queries this
     * rank has rows for.  q_dev still holds the WHOLE batch of nq_total rows; local query j is batch query
     * q_index[j]; owner_masks[i] has bit r set when rank r holds rows of batch query i's scope (it is what the
     * merge waits for, so it must be the same on every rank).  NULL q_index = every query, in order;
     * NULL owner_masks = every rank publishes every query. */
    uint32_t nq_total;
    const uint32_t *q_index;    /* HOST [nq] */
    const uint8_t *owner_masks; /* HOST [nq_total] */
} rf_peer_exchange;
int rf_search_keys_device_fused(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                uint32_t n_segs, uint32_t k, const rf_peer_exchange *px,
                                uint64_t *out_keys_dev, void *stream);

/* The store-sharded form of the fused exchange (whole stores per rank, configs[4]): one scope per query as in
 * rf_search_keys_device_scoped, `px` as above with nq_cap >= px->nq_total.  Each rank scans only the queries it
 * has rows for (px->q_index); its scan kernel stores each such query's k keys into every rank's gather buffer
 * and releases a flag there, and a second small kernel behind it on the stream acquires, per batch query, the
 * flags of the ranks in its owner mask and merges their lists into out_keys_dev [nq_total, k] -- two launches,
 * no collective, no host round trip, and no work at all for a query on a rank that holds none of its stores. */
int rf_search_keys_device_scoped_fused(rf_engine *e, const int8_t *q_dev, uint32_t nq, const uint32_t *store_segs,
                                       const uint32_t *seg_off, uint32_t k, const rf_peer_exchange *px,
                                       uint64_t *out_keys_dev, void *stream);

/* k-way merge after the all-gather of the sharded path: keys_dev is n_lists x nq x k packed keys
 * (device), out_keys_dev nq x k.  Enqueued on `stream`, no synchronisation. */
int rf_merge_topk_device(rf_engine *e, const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq,
                         uint32_t k, uint64_t *out_keys_dev, void *stream);

/* Query featurisation alone (text -> int8[RF_DIM] on the GPU), host buffers. */
int rf_featurize_query(rf_engine *e, const uint8_t *utf8, size_t n, int8_t *out_q);

/* ---- RF-1w: the IDF-weighted scoring variant (oracle/SPEC.md).  The reference has
 * no counterpart: it is the relevance variant its benchmark harness (scripts/benchmark/metrics.py:
 * 73-92, citation_hit) would grade.  Only the query vector changes, so the search kernels and the
 * packed keys are the plain RF-1 ones.
 *
 * rf_scope_df: per-bucket document frequencies of the scope's live rows, one streaming GPU pass
 * (cached per scope until the next ingest / delete).  out_df: RF_DIM u64 (HOST), *out_n: live rows.
 * rf_scope_df_device: the same sums ADDED into df_dev[0 .. RF_DIM) and the row count into
 * df_dev[RF_DIM] (DEVICE, caller zeroes), enqueued on `stream` without synchronising -- the form
 * the sharded path all-reduces over the ranks before taking the weights. */
int rf_scope_df(rf_engine *e, const uint32_t *store_segs, uint32_t n_segs, uint64_t *out_df, uint64_t *out_n);
int rf_scope_df_device(rf_engine *e, const uint32_t *store_segs, uint32_t n_segs, uint64_t *df_dev, void *stream);
/* Pure integer host functions: weights w[d] in [4, 31] from (df, n); qw[d] = min(q[d] * w[d], 127). */
int rf_idf_weights(const uint64_t *df, uint64_t n, uint32_t dim, uint8_t *out_w);
int rf_weight_query(const int8_t *q, const uint8_t *w, uint32_t dim, int8_t *out_qw);
/* rf_search_text_in with the query vector weighted on the GPU by `weights` (RF_DIM u8, HOST; NULL =
 * plain RF-1).  out_q receives the weighted vector. */
int rf_search_text_w(rf_engine *e, const uint8_t *utf8, size_t n, const uint32_t *store_segs,
                     uint32_t n_segs, const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights,
                     uint32_t k, uint64_t *out_ids, int32_t *out_scores, float *out_cos,
                     uint32_t *out_count, int8_t *out_q /* RF_DIM, may be NULL */);

/* ---- measurement aid: the achievable dense int8 rate of the tensor pipe on `device` (a kernel that only
 * issues tcgen05.mma.cta_group::2.kind::i8 256 x 256 x 32, the scoring GEMM's shape), timed with CUDA events:
 * the roofline denominator of the batched scoring path (configs[2]).  n_batches of 8 MMAs per
 * CTA pair (>= 8; 4000 takes a few ms).  *ops_per_s counts 2 * M * N * K per MMA. */
int rf_probe_int8_peak(int device, uint32_t n_batches, double *ops_per_s, double *ms /* may be NULL */);

/* ---- engine group: several GPUs behind ONE index in one process ----------------------------------------
 * The reference's backend is one object per process behind get_rag_client() (gemini_rag.py:721-725), called
 * by the chat route (routes/chat.py:499-505) and the ingest worker (services/ingestion.py:45-52); the north
 * star shards large corpora over the 8 GPUs of a box.  A group owns one engine per listed device and
 * presents the single-engine calls: store numbers are the same on every device, chunk ids are global
 * (engine d numbers its rows from id_bases[d]; NULL = d * floor((2^32 - 2) / n_devices)), placement decides
 * where a document's rows go, and a search runs concurrently on every device that holds rows of the scope,
 * the per-device top-k lists (k keys each, landing in mapped host memory) being merged on the host under the
 * RF-1 order -- a total order, so the result equals the single-engine answer bit for bit.  A device may be
 * listed twice (two engines on one GPU): that is how the single-GPU test-suite covers this code. */
#define RF_GROUP_MAX 8u
enum {
    RF_PLACE_STORE = 0,   /* whole stores per device: store g lives on device g % n (a store-scoped query touches one GPU) */
    RF_PLACE_SPREAD = 1   /* each document goes to the device holding the fewest rows (one huge store is sharded by chunk) */
};
typedef struct rf_group rf_group;
typedef struct rf_group_config {
    uint32_t struct_size;     /* sizeof(rf_group_config) */
    uint32_t n_devices;       /* 1 .. RF_GROUP_MAX */
    const int32_t *devices;   /* CUDA ordinals [n_devices] */
    const uint64_t *id_bases; /* [n_devices] or NULL */
    uint32_t n_contexts;      /* concurrent searches per device; 0 -> 8 */
    uint32_t placement;       /* RF_PLACE_* */
    uint64_t capacity_rows;   /* rows reserved PER DEVICE */
    uint32_t dim;             /* feature dimension of every engine; 0 -> RF_DIM */
    uint32_t reserved;
} rf_group_config;
int rf_group_create(const rf_group_config *cfg, rf_group **out);
int rf_group_destroy(rf_group *g);
int rf_group_size(rf_group *g, uint32_t *n_devices);
int rf_group_engine(rf_group *g, uint32_t index, rf_engine **out);  /* the engine of one device (borrowed) */
int rf_group_stats(rf_group *g, rf_stats *total, rf_stats *per_device /* [n_devices] or NULL */);
const char *rf_group_last_error(void);
/* stores / ingest / deletes: as rf_store_*, rf_ingest_*, rf_doc_tombstone, with group store numbers */
int rf_group_store_open(rf_group *g, const char *fs_name, uint32_t *store);
int rf_group_store_lookup(rf_group *g, const char *fs_name, uint32_t *store);
int rf_group_store_drop(rf_group *g, uint32_t store);
int rf_group_ingest_text(rf_group *g, uint32_t store, uint64_t doc_id, const uint8_t *utf8, size_t n,
                         uint64_t *first_chunk, uint32_t *n_chunks, int64_t *spans, uint32_t max_spans);
int rf_group_ingest_features(rf_group *g, uint32_t store, uint64_t doc_id, const int8_t *rows, uint64_t n_rows,
                             uint64_t *first_chunk);
/* RF-1 synthetic corpus.  RF_PLACE_SPREAD: rows_per_store must be 0; the n_rows counters are cut into
 * n_devices contiguous runs, device d generating run d (with id_bases[d] = start_counter + start of run d
 * the chunk ids equal those of one engine holding the whole corpus).  RF_PLACE_STORE: store first_store + i
 * takes rows [i * rows_per_store, ...) on device (first_store + i) % n_devices. */
int rf_group_ingest_synthetic(rf_group *g, uint32_t first_store, uint64_t rows_per_store, uint64_t seed,
                              uint64_t start_counter, uint64_t n_rows, const uint16_t *zipf_vocab);
int rf_group_doc_tombstone(rf_group *g, uint64_t doc_id);
/* queries: as rf_search / rf_search_text_w (HOST buffers; blocks until the merged results are in them) */
int rf_group_search(rf_group *g, const int8_t *q, uint32_t nq, const uint32_t *stores, const uint32_t *store_off,
                    uint32_t k, uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts);
int rf_group_search_text(rf_group *g, const uint8_t *utf8, size_t n, const uint32_t *stores, uint32_t n_stores,
                         const uint64_t *ranges, uint32_t n_ranges, const uint8_t *weights, uint32_t k,
                         uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_count,
                         int8_t *out_q /* RF_DIM, may be NULL */);
/* RF-1w: document frequencies summed over the devices (the statistic is a sum over rows) */
int rf_group_scope_df(rf_group *g, const uint32_t *stores, uint32_t n_stores, uint64_t *out_df, uint64_t *out_n);
/* durability: "<path>.dev<d>" per device (rf_snapshot_save) + "<path>.group"; load into an empty group of the
 * same size and placement */
int rf_group_snapshot_save(rf_group *g, const char *path);
int rf_group_snapshot_load(rf_group *g, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* RF_B200_H */
