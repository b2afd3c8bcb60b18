/*
 * biorseo_b200.h — C ABI of the B200 insertion-site search.
 *
 * This is the drop-in boundary for ONE path of persalteas/biorseo: the motif
 * insertion-site search that runs inside the MOIP constructor
 * (reference cppsrc/MOIP.cpp:146-296 dispatch, :912-1213 per-module jobs,
 * cppsrc/Motif.cpp:432-494 placement enumerator). The reference has no FFI for this
 * path; the seam is "sequence + allowed_basepair + module library in, insertion sites
 * out". Everything below is plain C: pointers, sizes, status codes. No C++ or torch
 * types cross this boundary, nothing throws, nothing calls exit().
 *
 * There is no CPU implementation behind these entry points: every bss_search* call runs
 * hand-written sm_100a CUDA kernels and fails with BSS_ERR_NO_DEVICE / BSS_ERR_CUDA when
 * no usable GPU is present.
 *
 * Vocabulary
 *   module   one entry of a motif library (a JSON entry, a CaRNAval RIN file, a .desc file)
 *   unit     one searchable pattern set of a module. JSON/RIN: exactly one per module.
 *            DESC: one per "variant" produced by widening short components
 *            (cppsrc/MOIP.cpp:966-1058).
 *   component one strand of a unit: a byte pattern, '.' matching any byte
 *            (cppsrc/Motif.cpp:438: the pattern is compiled as a regex of literals and dots).
 *   check    one base-pair compatibility test between two positions derived from a
 *            placement (JSON/RIN links, DESC closing pairs; cppsrc/MOIP.cpp:1072-1079,
 *            1131-1137, 1194-1205).
 *   site     one surviving placement: (module, start of every component).
 */
#ifndef BIORSEO_B200_H_
#define BIORSEO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSS_ABI_VERSION 2

/* ---- status codes (0 = success, negative = failure) ---- */
#define BSS_OK                 0
#define BSS_ERR_INVALID      (-1)  /* NULL / malformed argument or table                              */
#define BSS_ERR_NO_DEVICE    (-2)  /* no CUDA device: this library has no CPU path                    */
#define BSS_ERR_CUDA         (-3)  /* CUDA runtime failure, see bss_last_error()                      */
#define BSS_ERR_LIMIT        (-4)  /* table or query exceeds a documented limit (BSS_MAX_*)            */
#define BSS_ERR_BAD_SEQUENCE (-5)  /* query contains '\n' or '\r' (regex '.' would not match them)     */
#define BSS_ERR_OVERFLOW     (-6)  /* caller-provided output too small; required size is reported      */
#define BSS_ERR_COUNT        (-7)  /* one unit has >= 2^32 sites on one sequence                       */
#define BSS_ERR_NOMEM        (-8)  /* host allocation failed                                           */
#define BSS_ERR_INTERNAL     (-9)  /* an unexpected C++ exception was caught at the boundary           */

/* ---- documented limits ---- */
#define BSS_MAX_COMPONENTS   16    /* components per unit                                              */
#define BSS_MAX_PATTERN     255    /* bytes per component pattern                                      */
#define BSS_MAX_CHECKS      255    /* checks per unit                                                  */
#define BSS_MAX_ALPHABET     32    /* distinct literal pattern bytes per library                       */
#define BSS_MAX_SEQ_LEN   16000    /* query length in bytes                                            */
#define BSS_ANCHOR_ABSOLUTE (-1)   /* check endpoint that is a fixed sequence position                 */

/*
 * Flat component table: what the host loaders (JSON / RIN / DESC parsers + validators,
 * reference cppsrc/Motif.cpp:218-385, cppsrc/MOIP.cpp:932-1066,1108-1122,1158-1172) hand
 * to the device. All arrays are caller-owned host memory, read once by bss_library_create.
 *
 * Units [0, n_units) are searched and emit sites. Units [n_units, n_units + n_gates) are
 * "gates": pattern-only units that emit nothing; a searched unit whose unit_gate names one
 * is skipped on a sequence where the gate has no placement at all (the regex pre-filter
 * is_desc_insertible, cppsrc/Motif.cpp:387-430).
 *
 * Placement semantics (cppsrc/Motif.cpp:432-494): component c, of length k_c, is placed at
 * p_c; p_0 runs over the leftmost non-overlapping matches of pattern 0 from position 0, and
 * p_{c+1} over the leftmost non-overlapping matches of pattern c+1 in the suffix that starts
 * at p_c + k_c - 1 + delay, which must start inside the sequence. A zero-length pattern
 * matches at every position 0..n.
 *
 * A check (a_anchor, a_off, b_anchor, b_off) passes when allowed(u, v) with
 * u = p[a_anchor] + a_off (or just a_off when a_anchor == BSS_ANCHOR_ABSOLUTE), v likewise;
 * allowed(u, v) = both inside [0, n) and mask bit (min, max) set. A site is emitted when all
 * checks of its unit pass.
 */
typedef struct bss_table {
    uint32_t        n_modules;        /* number of library modules that own units                   */
    uint32_t        n_units;          /* searched units                                             */
    uint32_t        n_gates;          /* gate units, stored after the searched ones                 */
    const uint32_t* unit_module;      /* [n_units + n_gates] module index written into each site    */
    const uint32_t* unit_delay;       /* [n_units + n_gates] Motif::delay for this unit             */
    const uint32_t* unit_gate;        /* [n_units] index of the gate unit, or UINT32_MAX            */
    const uint32_t* unit_comp_begin;  /* [n_units + n_gates + 1] first component of each unit       */
    const uint32_t* comp_pat_begin;   /* [n_components + 1] first pattern byte of each component    */
    const uint8_t*  pattern;          /* pattern bytes; '.' is the wildcard                         */
    const uint32_t* unit_check_begin; /* [n_units + n_gates + 1] first check of each unit           */
    const int16_t*  checks;           /* [n_checks][4] = a_anchor, a_off, b_anchor, b_off           */
} bss_table;

/* Opaque handles. A library owns device memory, one CUDA stream and growable work
 * buffers; it is immutable after creation and is used by one host thread at a time
 * (the reference calls this path once, from the main thread: cppsrc/biorseo.cpp:175). */
/* Lifetime: queries, result handles (bss_sites) and indexes (bss_site_index) may be released before or
 * after the library they came from; bss_library_destroy detaches the ones still alive (they keep their own
 * buffers and can still be read and freed), bss_query_destroy waits for a search of that query that is
 * still queued. A query must outlive every search that uses it. */
typedef struct bss_library bss_library;
typedef struct bss_query   bss_query;   /* sequences + pair masks resident in device memory */
typedef struct bss_sites   bss_sites;   /* one search result                                */

typedef struct bss_stats {
    uint64_t placements_tested;   /* sum over (sequence, unit) of sequence length              */
    uint64_t n_sites;             /* sites emitted                                             */
    uint64_t n_words;             /* u32 words emitted                                         */
    uint64_t h2d_bytes;           /* bytes copied host->device by the last call                */
    uint64_t d2h_bytes;           /* bytes copied device->host by the last call                */
    uint32_t kernel_launches;     /* kernels launched by the last search                       */
    float    count_ms;            /* device time of the match+count kernel (CUDA events)       */
    float    scan_ms;             /* device time of the offset scan kernel                     */
    float    emit_ms;             /* device time of the emit kernel                            */
    float    total_ms;            /* first launch to last kernel end                           */
    uint64_t window_items;        /* items that took the window walk (route_kernel's list)     */
    uint64_t ordinary_items;      /* items left to the ordinary (rank-space / depth-first) passes */
    uint64_t heavy_items;         /* ordinary items parked to be walked block by block; above 1024 (the
                                     parking list's capacity) the rest were walked by the warp that found them */
    uint64_t placements_enumerated; /* set by bss_placements_enumerated (0 until it is called)  */
} bss_stats;

/* ---- library ---- */
/* device < 0 selects the current CUDA device. Replaces: the per-run library walk +
 * per-module job construction of cppsrc/MOIP.cpp:159-188, 211-234, 259-288. */
int  bss_library_create(const bss_table* table, int device, bss_library** out);
void bss_library_destroy(bss_library* lib);
/* Use an externally owned stream (e.g. the caller's current stream) instead of the
 * library's own; pass NULL to return to the internal stream. */
int  bss_library_set_stream(bss_library* lib, void* cuda_stream);
/* Library shards (one per GPU): `base` is added to every module index written into a record,
 * so that records of different shards can be concatenated in canonical global order. */
int  bss_library_set_module_base(bss_library* lib, uint32_t base);
/* Developer knobs: force rarely taken paths on small inputs (tests), compare kernel variants (benchmarks).
 * value < 0 returns a knob to automatic. Values are validated (BSS_ERR_INVALID). Not a tuning interface
 * for production callers: the defaults are the measured optimum. */
#define BSS_TUNE_GROUP            0   /* lanes per (sequence, unit) item: 4, 8, 16 or 32              */
#define BSS_TUNE_UNITS_PER_CHUNK  1   /* units per count-pass CTA (>= 1)                               */
#define BSS_TUNE_HEAVY_RANKS      2   /* placements from which an item is walked block by block (>= 1) */
#define BSS_TUNE_LIST_CAP         3   /* match-list entries per item of the rank-space walk (0: off)   */
#define BSS_TUNE_WINDOW           4   /* 0: never take the window walk                                 */
#define BSS_TUNE_WINDOW_STAGE     5   /* 1: window kernel stages occurrence table + banded mask in shared memory (cp.async.bulk) */
#define BSS_TUNE_WINDOW_DRAW      6   /* list entries a warp of the window count pass draws per atomic (1..1024)        */
#define BSS_TUNE_PATTERN_TABLE    7   /* 0: no pattern table (every unit matches its own patterns)                         */
#define BSS_TUNE_PHASE_EVENTS     8   /* 0: no CUDA events between the phases of a search (count_ms / scan_ms / emit_ms stay 0)   */
int  bss_library_set_tuning(bss_library* lib, int knob, int64_t value);
uint32_t bss_library_units(const bss_library* lib);
uint32_t bss_library_window_units(const bss_library* lib);   /* units every component of which is tied by a pair to an earlier one */
uint32_t bss_library_modules(const bss_library* lib);
/* Number of components of the units of `module` (identical for all its units). */
uint32_t bss_library_module_components(const bss_library* lib, uint32_t module);

/* ---- queries ---- */
/* n_seqs sequences, concatenated in `seqs`; sequence q occupies
 * [seq_begin[q], seq_begin[q+1]). Its pair mask is n_q rows of ceil(n_q/32) little-endian
 * u32 words at masks + mask_begin[q]; bit v of row u (u < v) set means the reference's
 * allowed_basepair(u, v) holds (cppsrc/MOIP.cpp:896-910; theta, the v-u >= 4 and u < n-6
 * rules are already applied by the caller's loop cppsrc/MOIP.cpp:77-90). Bits on or below
 * the diagonal are ignored. Copies everything to the device. */
int  bss_query_create(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin,
                      const uint32_t* masks, const uint64_t* mask_begin, bss_query** out);
void bss_query_destroy(bss_query* q);

/* ---- pair mask from the base-pair probability matrix (row "next" of the path: the caller's loop
 * cppsrc/MOIP.cpp:77-90 that defines MOIP::allowed_basepair, cppsrc/MOIP.cpp:896-910) ---- */
/* bit v of row u (u < v) is set iff v >= u + 4, u + 6 < n and pij[u * row_stride + v * col_stride] > theta
 * (float comparison, NaN is "not allowed"). Strides are in elements: (n, 1) for a row-major matrix,
 * (1, n) for Eigen's column-major MatrixXf. For n <= 6 the mask is empty (the reference's unsigned
 * n - 6 makes its own behaviour undefined there). */
/* Device pointers in and out, asynchronous on `cuda_stream` (NULL: the library's stream). */
int  bss_pair_mask_device(bss_library* lib, const float* d_pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta,
                          uint32_t* d_mask, void* cuda_stream);
/* Host matrix in, host mask out (n * ceil(n/32) words): upload, kernel, download. */
int  bss_pair_mask(bss_library* lib, const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask);
/* One-sequence query whose pair mask is built on the device from the host probability matrix:
 * the packed mask never exists on the host. */
int  bss_query_create_pij(bss_library* lib, const char* seq, uint32_t n, const float* pij, uint64_t row_stride, uint64_t col_stride,
                          float theta, bss_query** out);

/* ---- search ---- */
/* Replaces Pool + MOIP::allowed_motifs_from_{desc,rin,json} + find_next_ones_in for every
 * (sequence, unit). Sites come back in canonical order (sequence, unit, depth-first
 * placement order) as u32 records [module, p_0 .. p_{C-1}]. */

/* Search a resident query. fetch != 0 also copies the records to pinned host memory
 * (bss_sites_records); fetch == 0 leaves them on the device (bss_sites_device_records). */
int  bss_search_query(bss_library* lib, const bss_query* q, int fetch, bss_sites** out);

/* Search a resident query into caller-owned DEVICE memory (e.g. a tensor that will feed an
 * NCCL all-gather). On BSS_ERR_OVERFLOW no record is written and *n_words holds the size needed.
 * d_seq_word_begin may be NULL, else receives n_seqs + 1 u64 word offsets (device memory) — always, also on
 * overflow: its last entry is the number of record words the search produced, which is what an exchange
 * queued behind the search reads (a count above the capacity raises the overflow flag on every rank). */
int  bss_search_query_into(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words,
                           uint64_t* d_seq_word_begin, uint64_t* n_words, uint64_t* n_sites);

/* The same in two halves, for callers that keep the device busy (a stream of queries, a sharded
 * exchange queued right behind the search): _async queues everything on the library's stream and
 * returns at once; bss_search_wait blocks until the last queued search is complete and reports its
 * sizes / errors exactly like bss_search_query_into. Queuing several searches before one wait is
 * allowed (each overwrites the previous one's totals: only the last one is reported). */
int  bss_search_query_into_async(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words,
                                 uint64_t* d_seq_word_begin);
int  bss_search_wait(bss_library* lib, uint64_t* n_words, uint64_t* n_sites);

/* One-shot, host buffers in, host records out: upload, search, download. */
int  bss_search(bss_library* lib, const char* seq, uint32_t n, const uint32_t* mask, bss_sites** out);
int  bss_search_batch(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin,
                      const uint32_t* masks, const uint64_t* mask_begin, bss_sites** out);

/* The same, records left in device memory (bss_sites_device_records; bss_sites_records is NULL): for callers that
 * consume them on the device, or fetch ranges of them with bss_sites_copy_out. */
int  bss_search_batch_resident(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin,
                               const uint32_t* masks, const uint64_t* mask_begin, bss_sites** out);

/* ---- multi-GPU in ONE process: the library cut over the GPUs of the box (SURVEY 8b "X_search_sharded") ---- */
/* The reference's unit of parallelism is one job per module of one library (cppsrc/MOIP.cpp:246-293, and the RIN /
 * DESC twins :159-243): the library is cut into n_gpus CONTIGUOUS module ranges balanced by a search-cost estimate
 * (match steps + expected placements per module on a query of typical_len nucleotides; 0: 1000), every GPU owns one
 * range (its records carry global module indices) and searches it on the replicated query; the ranges' records come
 * back as ONE stream in canonical (sequence, module, placement) order in pinned host memory — the very stream
 * bss_search_batch gives on one GPU, bit for bit. One host thread per GPU inside; no torch, no NCCL, no IPC.
 * devices == NULL selects devices 0 .. n_gpus - 1. Units must come module by module (unit_module non-decreasing).
 * Like a library, a sharded handle is used by one host thread at a time; the per-range handles of bss_sharded_library
 * are for statistics and tuning between searches, not for searches of their own. */
typedef struct bss_sharded bss_sharded;
int          bss_sharded_create(const bss_table* table, const int* devices, uint32_t n_gpus, uint32_t typical_len, bss_sharded** out);
void         bss_sharded_destroy(bss_sharded* s);
uint32_t     bss_sharded_gpus(const bss_sharded* s);
int          bss_sharded_range(const bss_sharded* s, uint32_t shard, int* device, uint32_t* module_begin, uint32_t* module_end);
bss_library* bss_sharded_library(bss_sharded* s, uint32_t shard);   /* the range's own library handle (statistics, tuning) */
int          bss_search_sharded(bss_sharded* s, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin,
                                const uint32_t* masks, const uint64_t* mask_begin, bss_sites** out);
/* The cut alone (host only, no GPU needed): module_begin[0 .. n_shards] receives the range boundaries. */
int          bss_shard_ranges(const bss_table* table, uint32_t n_shards, uint32_t typical_len, uint32_t* module_begin);

/* ---- multi-GPU: cut of a padded all-gather (biorseo_b200/sharded.py) ---- */
/* d_blocks holds `world` blocks of block_words u32 each: [header_words of header | records | padding];
 * the last u64 of a header is the block's number of record words (what bss_search_query_into writes
 * when the header is its d_seq_word_begin and the records follow it). Writes the records of all
 * blocks back to back, in rank order, to d_out (room for world * (block_words - header_words) words),
 * the per-rank counts to d_counts[world], and sets *d_overflow to 1 when a count exceeds a block
 * (then nothing of that exchange may be used). Asynchronous on `cuda_stream`; no host round trip. */
int  bss_gather_cut(bss_library* lib, const uint32_t* d_blocks, uint32_t world, uint64_t block_words, uint32_t header_words,
                    uint32_t* d_out, uint64_t* d_counts, uint32_t* d_overflow, void* cuda_stream);

/* ---- multi-GPU: all-gather-v of the shards' record streams in one kernel over NVLink peer memory ---- */
/* One process per GPU. Every rank creates an exchange of the same geometry, publishes its 64-byte CUDA IPC handle
 * to all ranks (any transport: the callers here use a torch.distributed all-gather), and opens the peers' handles.
 * The search then writes straight into the rank's block — bss_search_query_into(_async) with
 * d_seq_word_begin = bss_exchange_block() and d_records = bss_exchange_block() + header_words (header_words =
 * 2 * (n_seqs + 1): the u64 word offsets, whose last entry is the number of record words) — and
 * bss_exchange_run_async, called by EVERY rank on the stream of its search, exchanges the counts, copies the
 * rank's records into every peer's output at the rank's displacement (peer-to-peer stores) and returns when all
 * peers' records have landed: *d_out is then the rank-order concatenation (= canonical (module, placement) order
 * with bss_library_set_module_base), d_result[0..world) the per-rank word counts, d_result[world] their sum and
 * d_result[world + 1] flags (1: a rank's stream exceeds capacity_words — nothing was copied; 2: a peer did not
 * show up within about two seconds). No host round trip; the output is double-buffered (valid until the second
 * next run). At most 16 ranks. */
typedef struct bss_exchange bss_exchange;
int       bss_exchange_create(int device, uint32_t world, uint32_t rank, uint64_t capacity_words, uint32_t header_words, bss_exchange** out);
int       bss_exchange_handle(const bss_exchange* x, void* handle64);
int       bss_exchange_open(bss_exchange* x, const void* handles /* world x 64 bytes, in rank order */);
/* The same for ONE process that drives every GPU itself (bss_sharded_* below): `all` holds the world exchanges in
 * rank order, each created on its own device; plain peer access (cudaDeviceEnablePeerAccess), no IPC. */
int       bss_exchange_open_local(bss_exchange* const* all, uint32_t world);
uint32_t* bss_exchange_block(const bss_exchange* x);      /* device memory: [header_words | capacity_words] */
uint64_t  bss_exchange_capacity(const bss_exchange* x);   /* record words per rank (rounded up to a multiple of 4) */
int       bss_exchange_run_async(bss_exchange* x, void* cuda_stream, const uint32_t** d_out, const uint64_t** d_result);
void      bss_exchange_destroy(bss_exchange* x);

/* ---- results ---- */
uint64_t        bss_sites_count(const bss_sites* s);
uint64_t        bss_sites_words(const bss_sites* s);
const uint32_t* bss_sites_records(const bss_sites* s);         /* pinned host memory or NULL   */
const uint32_t* bss_sites_device_records(const bss_sites* s);  /* device memory                */
const uint64_t* bss_sites_seq_word_begin(const bss_sites* s);  /* [n_seqs + 1] host            */
/* Copies the words [word_begin, word_begin + n_words) of the record stream to host memory (pinned or pageable),
 * wherever the records live. */
int             bss_sites_copy_out(const bss_sites* s, uint64_t word_begin, uint64_t n_words, uint32_t* host_dst);
void            bss_sites_free(bss_sites* s);

/* ---- insertion-variable index (row "next" of the path: the consumer of the site list) ---- */
/* What the MOIP constructor derives from insertion_sites_ with host loops, built on the device from the
 * records of the LAST search on `lib` (it reads the per-unit site counts that search left behind, so it
 * must run before the next search; a pending asynchronous search must have been waited for):
 *   - the C_{x,i,p} variables, numbered like index_of_first_components / index_of_Cxip_
 *     (cppsrc/MOIP.cpp:302-328): site x owns the variables [VAR_BEGIN[x], VAR_BEGIN[x+1]), one per
 *     component, with the component's first and last nucleotide (the "[first,second]" of the name);
 *   - per variable, the number of y_{u,v} terms of its "c3" constraint (cppsrc/MOIP.cpp:507-566; the
 *     constraint is only added when that number is > 0): c3_rule BSS_C3_COMPONENT_LESS_LINKS for
 *     jsonfolder / rinfolder (every allowed pair with an end on the component, the site's own links
 *     left out), BSS_C3_INTERIOR for descfolder (allowed pairs with an end strictly inside it);
 *   - per nucleotide u, the variables whose component covers u, in variable order: the terms of the
 *     "c4" overlap constraint (cppsrc/MOIP.cpp:569-586), OVERLAP_VARS[OVERLAP_BEGIN[u] .. OVERLAP_BEGIN[u+1]);
 *   - the pair side: ROW_FIRST_PAIR[u] = number of allowed pairs (u', v) with u' < u, i.e. the y index of
 *     the first pair of row u (index_of_yuv_, cppsrc/MOIP.cpp:74-90), and per nucleotide its allowed
 *     partners in increasing order, PAIR_PARTNER[PAIR_BEGIN[u] .. PAIR_BEGIN[u+1]) — the terms of "c1"
 *     (cppsrc/MOIP.cpp:472-487) and, restricted to a component, of c3.
 * q == NULL names the query of the last host-buffer search (bss_search, bss_search_batch).
 * `seq` selects the sequence of a batch query; d_records is the start of the whole record stream of that
 * search (bss_sites_device_records, or the caller's buffer of bss_search_query_into). fetch != 0 also
 * copies every section to pinned host memory. All sections are u32 arrays. */
#define BSS_C3_COMPONENT_LESS_LINKS 0
#define BSS_C3_INTERIOR             1
#define BSS_IDX_VAR_BEGIN       0   /* [sites + 1]                    */
#define BSS_IDX_VAR_FIRST       1   /* [variables]                    */
#define BSS_IDX_VAR_SECOND      2   /* [variables]                    */
#define BSS_IDX_VAR_C3_TERMS    3   /* [variables]                    */
#define BSS_IDX_OVERLAP_BEGIN   4   /* [n + 1]                        */
#define BSS_IDX_OVERLAP_VARS    5   /* [sum of component lengths]     */
#define BSS_IDX_PAIR_BEGIN      6   /* [n + 1]                        */
#define BSS_IDX_PAIR_PARTNER    7   /* [2 x allowed pairs]            */
#define BSS_IDX_ROW_FIRST_PAIR  8   /* [n + 1]                        */
#define BSS_IDX_SECTIONS        9
typedef struct bss_site_index bss_site_index;
int  bss_site_index_build(bss_library* lib, const bss_query* q, uint32_t seq, const uint32_t* d_records, int c3_rule, int fetch,
                          bss_site_index** out);
uint64_t        bss_site_index_count(const bss_site_index* x, int section);    /* elements of a section            */
const uint32_t* bss_site_index_host(const bss_site_index* x, int section);     /* pinned host memory, or NULL      */
const uint32_t* bss_site_index_device(const bss_site_index* x, int section);   /* device memory                    */
float           bss_site_index_ms(const bss_site_index* x);                    /* device time of the build kernels */
uint32_t        bss_site_index_launches(const bss_site_index* x);              /* kernels launched by the build    */
void            bss_site_index_free(bss_site_index* x);

/* ---- diagnostics ---- */
int         bss_last_stats(const bss_library* lib, bss_stats* out);
const char* bss_last_error(void);   /* thread-local, never NULL */
int         bss_abi_version(void);
int         bss_device_count(void);
/* The reference's real work term ("placements enumerated"): the exact number of placements
 * find_next_ones_in materialises BEFORE the pair filters (cppsrc/Motif.cpp:466-474), summed over every
 * (sequence, unit) of the query whose gate is open. A diagnostic kernel of its own (a backward counting
 * programme per item), not part of a search: the search kernels never visit most of these placements.
 * q == NULL names the query of the last host-buffer search. Also recorded in bss_stats. */
int         bss_placements_enumerated(bss_library* lib, const bss_query* q, uint64_t* placements);
/* Measured integer-issue rate of the device (thread-level 32-bit shift / logic operations per second, all SMs
 * busy with nothing else, at the clocks the GPU runs at): the denominator of the search's integer roofline. */
int         bss_measure_int_issue(bss_library* lib, double* thread_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* BIORSEO_B200_H_ */
