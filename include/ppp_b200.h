/* ppp_b200.h -- C ABI of the B200 (sm_100a) CUDA layer that replaces PingPongPro's data-parallel core.
 *
 * The reference (suhrig/pingpongpro, one translation unit, /root/reference/pingpongpro.cpp) has no
 * plugin / FFI interface; the drop-in boundary is the set of internal seams main() calls at
 * pingpongpro.cpp:1490, :1580, :1583, :1584, :1588 and :1606.  Every entry point below names the
 * reference function (file:line) it replaces.  INTEGRATION.md shows the reference-side call sites.
 *
 * Conventions: C linkage; plain pointers and sizes only (no C++/STL/torch types); every function returns
 * 0 (PPP_OK) or a PPP_ERR_* code and ppp_last_error() describes the last failure of the calling thread;
 * the caller owns host buffers, the library owns device memory and the pinned buffers it hands out;
 * ONE host thread drives a context; a context is bound to one GPU; calls on a context are ordered, and
 * device work is asynchronous to the host until a call that returns data.
 * There is no CPU fallback: without a usable CUDA device ppp_init fails with PPP_ERR_CUDA.
 */
#ifndef PPP_B200_H
#define PPP_B200_H
#include "ppp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PPP_OK = 0,
    PPP_ERR_ARG = 1,        /* invalid argument */
    PPP_ERR_CUDA = 2,       /* CUDA runtime failure (message has the CUDA error string) */
    PPP_ERR_STATE = 3,      /* call out of order (e.g. ppp_scan before ppp_height_hist) */
    PPP_ERR_RANGE = 4,      /* a read's key lies outside its contig (beyond @SQ LN + 1 + max overlap) */
    PPP_ERR_DEGENERATE = 5, /* every rounded stack height is unique: the reference divides 0/0 at
                               pingpongpro.cpp:551/:593 and indexes out of bounds */
    PPP_ERR_NOMEM = 6
};

#define PPP_HEIGHT_SCORE_BINS 1000 /* HEIGHT_SCORE_BINS, pingpongpro.cpp:164 */
#define PPP_MAX_OVERLAPS 32        /* upper bound on max_overlap - min_overlap + 1 */

typedef struct ppp_ctx ppp_ctx;

typedef struct ppp_config {
    int32_t device;                 /* CUDA device ordinal */
    uint32_t n_contigs;             /* number of @SQ lines */
    const uint64_t* contig_lengths; /* @SQ LN per contig (rID order) */
    int32_t min_overlap;            /* MIN_ARBITRARY_OVERLAP, pingpongpro.cpp:112 (3)  */
    int32_t max_overlap;            /* MAX_ARBITRARY_OVERLAP, pingpongpro.cpp:113 (23) */
    int32_t pingpong_overlap;       /* PING_PONG_OVERLAP,     pingpongpro.cpp:108 (10) */
    int32_t multi_hits;             /* PPP_MULTIHITS_*, pingpongpro.cpp:59 */
    /* Range sharding: this context owns the 5'-end keys whose global coordinate
     * (sum of the lengths of the preceding contigs + key) lies in [range_begin, range_end); keys at or
     * beyond a contig's length belong to the shard that holds the contig's last base.  Minus-strand keys
     * up to max_overlap beyond range_end are additionally accepted as a read-only halo.
     * range_begin = range_end = 0 means the whole genome. */
    uint64_t range_begin;
    uint64_t range_end;
    uint64_t pair_capacity;         /* initial capacity (records) of the signature store; 0 = automatic */
} ppp_config;

/* One reported ping-pong locus = one row of ping-pong_signatures.tsv (pingpongpro.cpp:903-910). */
typedef struct ppp_locus {
    uint32_t contig;
    uint32_t position;
    float fdr;
    float reads_plus;
    float reads_minus;
} ppp_locus;

/* One putative signature of any overlap = TPingPongSignature (pingpongpro.cpp:121-148) + its overlap. */
typedef struct ppp_pair {
    uint32_t contig;
    uint32_t position;
    uint32_t overlap;          /* in nt (min_overlap..max_overlap) */
    uint32_t height_score_bin; /* before collapseBins */
    uint32_t collapsed_bin;    /* after collapseBins (valid after ppp_score) */
    uint32_t local_bin;        /* IS_ABOVE_COVERAGE 0 / IS_BELOW_COVERAGE 1 (:162-163) */
    uint32_t bias_bin;         /* HAS_BASE_BIAS 0 / HAS_NO_BASE_BIAS 1 (:160-161) */
    float reads_plus;
    float reads_minus;
    float fdr;                 /* valid after ppp_score */
} ppp_pair;

/* Transposon interval (TTransposon, pingpongpro.cpp:173-213). start/end 0-based, end inclusive (:1224). */
typedef struct ppp_interval {
    uint32_t contig;
    uint32_t strand;
    uint32_t start;
    uint32_t end;
} ppp_interval;

typedef struct ppp_tp_result {
    float histogram[PPP_MAX_OVERLAPS]; /* TTransposon::histogram (:183), one entry per overlap */
    float reads_plus;                  /* :181 */
    float reads_minus;                 /* :182 */
    double z;                          /* zValue, :1266 (NaN when the p = 1 shortcut of :1250 applies) */
    double p;                          /* pValue before narrowing, :1267-1269 */
    float p_value;                     /* TTransposon::pValue, :1271 */
    float q_value;                     /* TTransposon::qValue, :1283-1312 */
} ppp_tp_result;

/* Device time of the kernels the last calls launched, from CUDA events on the context's stream. */
typedef struct ppp_timings {
    float stack_build_ms;   /* all ppp_reads_submit kernels since the last reset */
    float scan_ms;          /* fused height-histogram + overlap scan (ppp_height_hist) */
    float lut_ms;           /* frequency LUT kernel (ppp_height_hist) */
    float bin_ms;           /* height-score binning + grouped counts + ordering (ppp_scan) */
    float score_ms;         /* collapse + FDR table + per-locus FDR + compaction (ppp_score) */
    float transposon_ms;    /* ppp_transposons */
    uint32_t launches;      /* kernels launched by the library since the last ppp_timings_reset */
    float exchange_ms;      /* sharded runs: the two table exchanges of the last pass, waiting for the slowest rank included */
    float pass_ms;          /* first to last device operation of the last pass (ppp_height_hist .. ppp_score) */
    uint32_t reserved;
} ppp_timings;

/* Collective hook: in-place all-reduce of a DEVICE buffer across the contexts that share one
 * genome.  dtype: 0 = uint32, 1 = uint64.  op: 0 = sum, 1 = max.  `stream` is the context's cudaStream_t.
 * The host driver binds it to ncclAllReduce (or torch.distributed); a single-GPU run leaves it unset. */
typedef int (*ppp_allreduce_fn)(void* user, void* device_buffer, size_t count, int dtype, int op, void* stream);

PPP_API const char* ppp_last_error(void);
PPP_API int ppp_device_count(int* n);

/* Creates the CUDA context of `device` (the one-off driver initialisation, ~1 s).  Optional: lets a host driver
 * overlap that latency with input decoding by calling it from a helper thread before ppp_init. */
PPP_API int ppp_warmup(int32_t device);

/* Context: allocates and zeroes the dense per-strand stack arrays (TReadStacksPerGenome, :103-105). */
PPP_API int ppp_init(const ppp_config* cfg, ppp_ctx** ctx);
PPP_API void ppp_destroy(ppp_ctx* ctx);
/* Empties the stack store and all results (a new run over the same genome). */
PPP_API int ppp_reset(ppp_ctx* ctx);
PPP_API int ppp_set_allreduce(ppp_ctx* ctx, ppp_allreduce_fn fn, void* user);
/* Optional with a custom hook (ppp_nccl_attach does it itself): tells the context which of `world` shards it is.
 * The rare tall stacks (rounded height >= 512) are then exchanged as one compact list per rank inside a single
 * sum all-reduce, instead of a reduction over height tables as long as the tallest stack. */
PPP_API int ppp_set_rank(ppp_ctx* ctx, int rank, int world);

/* Built-in collective for the hook above: NCCL all-reduces on the context's stream.  libnccl.so.2 is loaded at run
 * time (PPP_NCCL_LIB overrides the path; a copy already loaded into the process is preferred), so single-GPU runs
 * do not need it.  Rank 0 creates the 128-byte id and hands it to the other ranks by any means (MPI, a socket,
 * torch.distributed); every rank then attaches, which creates its communicator (collective call) and installs it
 * with ppp_set_allreduce.  ppp_destroy destroys the communicator. */
#define PPP_NCCL_UNIQUE_ID_BYTES 128
PPP_API int ppp_nccl_unique_id(void* id);
PPP_API int ppp_nccl_attach(ppp_ctx* ctx, const void* id, int rank, int world);

/* Built-in collective for the contexts of ONE process (e.g. one host driver feeding the GPUs of a box, one thread per
 * context): the all-reduces go through host memory (they move kilobytes, twice per run).  Create the group, attach
 * every context with its rank (ascending range order; implies ppp_set_rank), then call ppp_height_hist / ppp_scan of
 * all contexts CONCURRENTLY from one thread each -- they meet inside the collective.  ppp_group_abort releases ranks
 * that wait for one that failed elsewhere (their calls return an error).  Destroy the group after its contexts. */
typedef struct ppp_group ppp_group;
PPP_API int ppp_group_create(int n_ranks, ppp_group** group);
PPP_API int ppp_group_attach(ppp_ctx* ctx, ppp_group* group, int rank);
PPP_API void ppp_group_abort(ppp_group* group);
PPP_API void ppp_group_destroy(ppp_group* group);

/* Pinned host memory for the double-buffered uploads. */
PPP_API int ppp_pinned_alloc(size_t bytes, void** ptr);
PPP_API int ppp_pinned_free(void* ptr);

/* Replaces the accumulation of countReadsInBamFile, pingpongpro.cpp:458-488:
 *   stack[strand][contig][key].reads += weight; .AAtPosition10 |= a10.
 * `reads` are `n` records of ONE contig in file order; calls must be made in file order (the float
 * accumulation of -m weighted is order dependent, :487).  The call enqueues the host-to-device copy and
 * the stack-builder kernels and returns; it blocks only until the copy of the PREVIOUS call has finished,
 * so a caller that alternates between two pinned buffers may refill one as soon as the call that
 * submitted the other returns.  Pageable buffers are staged through internal pinned memory.
 * Reads outside this context's range are counted (ppp_reads_foreign) and ignored.
 * `first_record_index` (the file-order index of reads[0]) is informational. */
PPP_API int ppp_reads_submit(ppp_ctx* ctx, int32_t contig, const ppp_read* reads, size_t n, uint64_t first_record_index);
/* The same with the reads in a 5-byte form (37 % less PCIe traffic than the 8-byte ppp_read): keys[i] as ppp_read.key;
 * meta[i] bit 0 = strand, bit 1 = A10, bits 2..7 = index into nh_table[64], the multi-hit counts (NH) that occur in THIS
 * call -- the caller fills the table while it packs; a batch with more than 64 distinct NH values is split, or goes
 * through ppp_reads_submit.  Same ordering and double-buffer contract as ppp_reads_submit (both arrays belong to the
 * host buffer of the call; nh_table is copied before the call returns). */
#define PPP_PACKED_NH_CODES 64
PPP_API int ppp_reads_submit_packed(ppp_ctx* ctx, int32_t contig, const uint32_t* keys, const uint8_t* meta, const uint32_t* nh_table, size_t n);
/* Same with records that are already in device memory (n records of one contig). */
PPP_API int ppp_reads_submit_device(ppp_ctx* ctx, int32_t contig, const ppp_read* device_reads, size_t n);
/* End of input (end of the loop at :1477-1523): waits for all uploads and kernels.
 * Fails with PPP_ERR_RANGE when a key was beyond its contig. */
PPP_API int ppp_stacks_finish(ppp_ctx* ctx);
PPP_API int ppp_reads_foreign(ppp_ctx* ctx, uint64_t* n);
/* Test hook: stack heights (TReadStack::reads, :93) and A10 flags (:94) of `count` keys of one strand of
 * one contig starting at key `begin` (absolute contig coordinate inside this context's range). */
PPP_API int ppp_stacks_download(ppp_ctx* ctx, int32_t contig, int32_t strand, uint64_t begin, uint64_t count, float* reads, uint8_t* a10);

/* Replaces mapHeightsToScores, pingpongpro.cpp:502-509: freq[(unsigned)(0.5 + reads)] += 1 over both
 * strands.  Runs the fused streaming pass over the stack arrays (which also collects the overlap
 * candidates for ppp_scan).  *freq (optional) receives a library-owned host array of max_height + 1
 * exact counts; the reference's float counter is min(count, 2^24) (:116, :508).  All-reduced when a
 * collective hook is set.  With both outputs NULL the call only enqueues its kernels (so do ppp_scan without
 * outputs): ppp_height_hist(NULL, NULL), ppp_scan(NULL, NULL), ppp_score(...) run as one stream of kernels with a
 * single host synchronisation at the end. */
PPP_API int ppp_height_hist(ppp_ctx* ctx, const uint64_t** freq, uint32_t* max_height);

/* Replaces countStacksByGroup, pingpongpro.cpp:523-624: height-score bin, local-height bin and base-bias
 * bin of every (+ stack, - stack, overlap) pair; grouped[overlap][bin][bias][local] counts (:605) and the
 * signature records (:608).  `grouped` (optional) receives n_overlaps * 1000 * 2 * 2 exact counts
 * (all-reduced); the reference's float counter is min(count, 2^24).  *n_pairs (optional): signatures of
 * this context. */
PPP_API int ppp_scan(ppp_ctx* ctx, uint64_t* grouped, uint64_t* n_pairs);

/* Replaces collapseBins (:631-686) + calculateFDRs (:695-754) + the -s filter and ordering of
 * writePingPongSignaturesToFile (:901-903).  Outputs (each optional):
 *   n_bins                 number of collapsed height-score bins C
 *   bin_map[1000]          oldBinCollapsedBinMap (:637, :666)
 *   cells[W*C*2*2]         collapsed grouped counts (:685), layout [overlap][bin][bias][local]
 *   fdr_table[W*C*2*2]     FDRs (:697, :745)   (pass buffers of W*1000*4 floats)
 *   loci / n_loci          library-owned pinned host array of the overlap-10 signatures with both
 *                          heights >= min_stack_height, in (contig, position) order. */
PPP_API int ppp_score(ppp_ctx* ctx, uint32_t min_stack_height, uint32_t* n_bins, uint16_t* bin_map, float* cells, float* fdr_table,
                      const ppp_locus** loci, size_t* n_loci);

/* Test / writer hook: all signatures of this context in (contig, position, overlap) order. */
PPP_API int ppp_pairs_download(ppp_ctx* ctx, ppp_pair* out, size_t capacity, size_t* n);

/* Replaces findSuppressedTransposons, pingpongpro.cpp:1195-1313.  `intervals` in any order; results are
 * returned in the reference's order (contig, start, end; ties in input order, :1174) together with the
 * permutation `order[i]` = input index of result i (optional). */
PPP_API int ppp_transposons(ppp_ctx* ctx, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order);

/* The same for a range-sharded genome driven by ONE host thread: `ctxs` are the n_ctx contexts of the shards in
 * ascending range order (on one or several devices), each after its ppp_score.  Transposons that straddle a range
 * boundary are scored as in an unsharded run: the sequential single-precision sums of :1227-1234 continue from context
 * to context in position order, and the iterator history of :1215-1222 is replayed on the whole lists.
 * (ppp_transposons on a single shard only sees that shard's signatures.) */
PPP_API int ppp_transposons_sharded(ppp_ctx* const* ctxs, int n_ctx, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order);

/* Test hook, host only (no device needed): the height-score bin thresholds the binning kernel works with.
 * thresholds[b-1] (b = 1..999) = smallest float product hs = freq[h+] * freq[h-] whose bin
 * (int)(0.5 + log10f(hs) / log10f(max_freq * max_freq) * 999) (pingpongpro.cpp:551, :591-594) is >= b, evaluated
 * with the host libm; +inf when max_freq^2 does not reach bin b.  max_freq must be > 1.  Returns PPP_ERR_STATE when
 * the host's log10f is not monotone around a threshold (the device-side bin = number of thresholds <= hs needs that). */
PPP_API int ppp_host_bin_thresholds(float max_freq, float* thresholds /* 999 */);

/* Test hooks, host only: the same table from the restatement of glibc's log10f that the DEVICE evaluates
 * (csrc/log10f_emul.h; the device's table is compared with ppp_host_bin_thresholds' on every pass), and that log10f itself
 * for n floats >= 1. */
PPP_API int ppp_emul_bin_thresholds(float max_freq, float* thresholds /* 999 */);
PPP_API int ppp_emul_log10f(const float* x, size_t n, float* out);
/* Test hook, host only: h0 + w[0] + w[1] + ... in single precision, left to right (pingpongpro.cpp:487), evaluated the
 * way the device folds very long stacks: `piece` additions at a time as one composed integer map per binade
 * (csrc/fadd_chain.h), real additions only where the sum changes binade.  *real_additions counts the latter. */
PPP_API int ppp_emul_fadd_chain(float h0, const float* w, size_t n, uint32_t piece, float* out, uint64_t* real_additions);

/* What the last pass found, once it is known to the host (waits for the pass): the tallest stack (rounded height), the
 * signatures and overlap-`pingpong` signatures of this context, the stacks in its compact store.  Each optional. */
PPP_API int ppp_pass_info(ppp_ctx* ctx, uint32_t* max_height, uint64_t* n_pairs, uint64_t* n_pingpong, uint64_t* n_stacks);

PPP_API int ppp_timings_get(ppp_ctx* ctx, ppp_timings* out);
PPP_API int ppp_timings_reset(ppp_ctx* ctx);
/* Padded length (32-bit words per strand) of the dense stack arrays of this context. */
PPP_API int ppp_layout(ppp_ctx* ctx, uint64_t* words_per_strand, uint64_t* owned_positions);

#ifdef __cplusplus
}
#endif
#endif
