/* ppp_types.h -- plain C data types shared by the C-ABI CUDA layer (ppp_b200.h) and the
 * host front end (ppp_host.h).  No C++ / torch / STL types cross either boundary. */
#ifndef PPP_TYPES_H
#define PPP_TYPES_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PPP_API __attribute__((visibility("default")))
#else
#define PPP_API
#endif

/* One aligned read after feature extraction: what pingpongpro.cpp:415-484 derives from a
 * record before it touches the stack store (:464/:476/:471/:483/:487).
 *   key  : 5'-end key on its contig: beginPos for a + read (:476); beginPos + alignmentLength
 *          for a - read (:464), i.e. one past the 3'-most aligned base.
 *   meta : bit 0 = strand (1 = minus, hasFlagRC :461)
 *          bit 1 = A at read position 10 (:470 / :482)
 *          bits 2..31 = NH, the multi-hit count (:442-444; 1 when the tag is absent). */
typedef struct ppp_read {
    uint32_t key;
    uint32_t meta;
} ppp_read;

#define PPP_READ_MINUS 1u
#define PPP_READ_A10 2u
#define PPP_READ_NH_SHIFT 2
#define PPP_READ_NH_MAX 0x3fffffffu

/* -m/--multi-hits (pingpongpro.cpp:59, :268-270). */
enum { PPP_MULTIHITS_WEIGHTED = 0, PPP_MULTIHITS_DISCARD = 1, PPP_MULTIHITS_UNIQUE = 2 };

#ifdef __cplusplus
}
#endif
#endif
