/* ppp_host.h -- C entry points of the host front end (libppp_host.so, no CUDA):
 * SAM/BAM decode and per-record feature extraction, i.e. the host half of
 * countReadsInBamFile (pingpongpro.cpp:402-484) that north_star keeps on the CPU.
 * Its output (ppp_read records per contig, in file order) is the upload format of
 * ppp_reads_submit() in ppp_b200.h. */
#ifndef PPP_HOST_H
#define PPP_HOST_H
#include "ppp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppp_fe ppp_fe; /* a decoded input file */

/* Decodes a whole SAM/BAM file.  Replaces BamStream + the record loop of pingpongpro.cpp:407-484.
 * min_len/max_len: -l/-L (:426).  mode: PPP_MULTIHITS_* (reads whose weight is <= 0 are dropped, :458).
 * Returns 0, or 1 = cannot open ("Failed to open input file", :1485), 2 = malformed record
 * ("Failed to read record", :411). */
PPP_API int ppp_fe_decode(const char* path, uint32_t min_len, uint32_t max_len, int mode, ppp_fe** out);
/* Same with `threads` BGZF inflate workers for BAM input (<= 1: sequential). */
PPP_API int ppp_fe_decode_mt(const char* path, uint32_t min_len, uint32_t max_len, int mode, int threads, ppp_fe** out);
PPP_API void ppp_fe_free(ppp_fe* fe);

/* The raw-DEFLATE decoder the BGZF reader tries before zlib (test hook): 1 = `in` decoded into exactly out_len bytes,
 * 0 = declined (invalid stream, wrong size); the reader then hands the block to zlib. */
PPP_API int ppp_fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
PPP_API int ppp_fe_n_contigs(const ppp_fe* fe);
PPP_API const char* ppp_fe_contig_name(const ppp_fe* fe, int contig);
PPP_API uint64_t ppp_fe_contig_length(const ppp_fe* fe, int contig);
/* Extracted reads of one contig, in file order. */
PPP_API size_t ppp_fe_n_reads(const ppp_fe* fe, int contig);
PPP_API const ppp_read* ppp_fe_reads(const ppp_fe* fe, int contig);
/* totalReadCount (:488): sum of the read weights in file order, as double. */
PPP_API double ppp_fe_total_weight(const ppp_fe* fe);
/* Records seen in the file / records kept after the filters of :415, :426, :458. */
PPP_API uint64_t ppp_fe_n_records(const ppp_fe* fe);
PPP_API uint64_t ppp_fe_n_kept(const ppp_fe* fe);
/* Test hook (threads > 1, BAM): runs of BGZF blocks whose guessed first record boundary was rejected by the in-order
 * check and which were therefore decoded again from the true boundary (host/aln_reader.h). */
PPP_API uint64_t ppp_fe_rejected_guesses(const ppp_fe* fe);

/* Transposon annotation input, replaces readTransposonsFromFile (pingpongpro.cpp:993-1175) incl. the format
 * detection of main() (:1543-1554).  Files are read in order; contig names that `contig_names` lacks are
 * appended to the name store (:1101-1106).  Intervals come back in the reference's order: contig id, then
 * (start, end), ties in file order.  Returns 0, or 1 = a file cannot be opened (:1539). */
typedef struct ppp_annot ppp_annot;
typedef struct ppp_annot_interval {
    uint32_t contig;
    uint32_t strand;
    uint32_t start;
    uint32_t end;
} ppp_annot_interval;
PPP_API int ppp_annot_read(const char* const* paths, int n_paths, const char* const* contig_names, int n_names, ppp_annot** out);
/* -T: clusters overlap-10 loci (given in (contig, position) order) that lie within `range` of each other into putative
 * transposons, predictSuppressedTransposons pingpongpro.cpp:1326-1354 -- the last cluster of a contig is never
 * emitted and clusters must be longer than 30 nt, as in the reference.  Results through the accessors below. */
PPP_API int ppp_annot_predict(const uint32_t* contigs, const uint32_t* positions, size_t n, const char* const* contig_names, int n_names, uint32_t range,
                              ppp_annot** out);
PPP_API void ppp_annot_free(ppp_annot* a);
PPP_API size_t ppp_annot_count(const ppp_annot* a);
PPP_API void ppp_annot_intervals(const ppp_annot* a, ppp_annot_interval* out);
PPP_API const char* ppp_annot_identifier(const ppp_annot* a, size_t i);
PPP_API int ppp_annot_n_names(const ppp_annot* a);
PPP_API const char* ppp_annot_name(const ppp_annot* a, int i);

#ifdef __cplusplus
}
#endif
#endif
