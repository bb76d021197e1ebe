// common.cuh -- shared declarations of the sm_100a CUDA layer (libppp_b200.so).
//
// Data layout in HBM: one dense 32-bit word per strand per genome position,
//   H[0 .. G)      plus strand,  H[G .. 2G)  minus strand,   G = padded words per strand.
// Contigs occupy consecutive "slots" of G; a slot holds the keys this context owns plus a right
// halo of max_overlap minus-strand keys, and is padded with zero words so that the overlap window of a
// plus position never reaches the next slot.  A word is 0 when no stack exists, else
//   integer representation (-m unique / discard): bits 0..30 = number of reads, bit 31 = A10 flag
//   float   representation (-m weighted)        : bits 0..30 = IEEE-754 bits of the height, bit 31 = A10
// (heights are > 0, so the sign bit is free).  TReadStack of pingpongpro.cpp:91-99.
// An occupancy bitmap (bit i set <=> word i of H is non-zero) is maintained beside H by the stack builder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ppp_b200.h"
#include "ppp_types.h"

namespace pppk {

constexpr uint32_t kA10Bit = 0x80000000u;
constexpr uint32_t kValueMask = 0x7fffffffu;
constexpr uint32_t kSaturate = 1u << 24;        // float counters stop at 2^24
constexpr int kBins = 1000;                     // HEIGHT_SCORE_BINS, pingpongpro.cpp:164
constexpr uint32_t kScanLowBins = 512;          // smem-privatised part of the height histogram
constexpr int kScanTile = 32768;                // positions per tile of the scan (G is a multiple): 256 threads x 4 occupancy words
constexpr uint32_t kFreqCapacity = kSaturate + 1;  // rounded heights never exceed 2^24
constexpr uint32_t kInvalidPos = 0xffffffffu;

enum Repr { kReprInt = 0, kReprFloat = 1 };

// Per-contig slot of this context (device copy).
struct Slot {
    uint32_t off;       // first word of the slot inside a strand array
    uint32_t first;     // first owned key (contig coordinate)
    uint32_t lenPlus;   // owned keys
    uint32_t lenMinus;  // owned keys + halo
    uint32_t keyLimit;  // keys >= keyLimit are outside the contig altogether (range error)
    uint32_t contig;    // rID of the contig this slot belongs to
};

// A signature record while it lives on the device (16 bytes).
// misc: bits 0..4 overlap index, bit 5 local bin, bit 6 bias bin, bits 7..16 height-score bin.
struct __align__(16) PairRec {
    uint32_t lpos;
    float rp;
    float rm;
    uint32_t misc;
};

// Scan-pass state that is zeroed (one memset) at the start of every pass; the look-back states of the tiles follow it
// in the same allocation.
struct PassState {
    unsigned long long freqLow[kScanLowBins];  // height histogram, rounded heights < kScanLowBins
    unsigned long long totals[2];              // signature records / overlap-`pingpong` records found (may exceed the capacities)
    unsigned long long maxCount;               // largest height frequency (k_bin_thresholds)
    uint32_t maxHeight;
    uint32_t ticket;                           // next unclaimed tile of k_scan
    uint32_t overflow;                         // multi-GPU: some rank listed more tall stacks than fit
    uint32_t maxTail;                          // largest frequency among the heights >= kScanLowBins
};
// What the binning kernels need besides the threshold table (written by k_bin_thresholds, or by the host)
struct BinParams {
    float binScale;  // 999 / maxHeightScore: only seeds the device's bin proposal
    float maxFreq;   // the maximum frequency the thresholds were built for
};
// State of one k_score launch (zeroed together with the look-back states that follow it).
struct ScoreState {
    unsigned long long nLoci;                  // reported loci
    uint32_t ticket;                           // next unclaimed tile
    uint32_t pad;
};

// Look-back state of one tile (decoupled look-back, Merrill & Garland): flag in bits 62..63, the number of
// overlap-`pingpong` records in bits 34..61, the number of signature records in bits 0..33.
constexpr int kLbRecBits = 34;
constexpr unsigned long long kLbRecMask = (1ull << kLbRecBits) - 1;
constexpr unsigned long long kLbValueMask = (1ull << 62) - 1;
constexpr unsigned long long kLbAggregate = 1ull << 62, kLbPrefix = 2ull << 62;
constexpr uint64_t kMaxPairCapacity = 0xfffffff0ull;       // signature records per context
constexpr uint64_t kMaxLociCapacity = (1ull << 28) - 16;   // overlap-`pingpong` records per context

struct ScanParams {
    const uint32_t* occ;    // occupancy bitmap: bit i <-> word i of H (plus strand words [0, G/32), minus [G/32, 2G/32))
    const uint32_t* cmpP;   // compact plus-strand stack words in position order (launch_compact)
    const uint32_t* cmpM;   // compact minus-strand stack words
    const uint32_t* tileOffP;  // [nTiles + 1] index in cmpP of the first stack of every tile
    const uint32_t* tileOffM;
    uint64_t G;             // words per strand
    uint32_t nTiles;
    int minOv, W, ppIdx;
    uint32_t haloLo, haloHi;  // minus-strand words in [haloLo, haloHi) are halo copies (not counted)
    PassState* state;
    unsigned long long* tileState;  // [nTiles]
    unsigned long long* freqLow;  // [kScanLowBins] (state->freqLow, or this rank's slot of the gather buffer)
    uint32_t* freqTail;           // [kFreqCapacity] (entries >= kScanLowBins used)
    uint32_t* maxHeight;
    // multi-GPU: rounded heights >= kScanLowBins are appended here instead of counted in freqTail (see launch_tall_apply)
    uint32_t* tallList;           // null: count in freqTail
    unsigned int* tallCount;
    uint32_t tallCap;
    uint32_t* maxTail;            // largest count reached in freqTail
    // signature records in (position, overlap) order, structure of arrays
    uint32_t* recPos;             // padded position of the plus stack
    float* recRp;
    float* recRm;
    uint8_t* recMisc;             // bits 0..4 overlap index, bit 5 local bin, bit 6 bias bin
    uint64_t capacity;
    PairRec* ppRecs;           // the overlap-`pingpong` records once more, in position order (K4 reads only these)
    uint64_t ppCapacity;
};

// The height -> frequency histogram as the kernels read it: exact counts, 64-bit for rounded heights < kScanLowBins,
// 32-bit beyond; the reference's float counter (:508) holds min(count, 2^24).
struct FreqTables {
    const unsigned long long* low;
    const uint32_t* tail;
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float freq_of(const FreqTables& f, uint32_t h) {  // heightScoreMap[h], :582 / :589
    const unsigned long long c = h < kScanLowBins ? f.low[h] : (unsigned long long)f.tail[h];
    return (float)(c < kSaturate ? (uint32_t)c : kSaturate);
}
template <int REPR>
__device__ __forceinline__ float word_height(uint32_t w) {
    uint32_t v = w & kValueMask;
    if (REPR == kReprInt) return (float)(v < kSaturate ? v : kSaturate);
    return __uint_as_float(v);
}
// (unsigned)(0.5 + reads): the std::map key conversion of pingpongpro.cpp:508 / :582 / :589.
__device__ __forceinline__ uint32_t round_height(float h) { return __double2uint_rz(__dadd_rn(0.5, (double)h)); }
template <int REPR>
__device__ __forceinline__ uint32_t word_round_height(uint32_t w) {
    uint32_t v = w & kValueMask;
    if (REPR == kReprInt) return v < kSaturate ? v : kSaturate;
    return round_height(__uint_as_float(v));
}

// ---- host-callable launchers (defined in the .cu files) --------------------------------------
cudaError_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, size_t scratchWords, cudaStream_t s, int* launches);
size_t exclusive_scan_scratch_words(size_t n);

cudaError_t launch_stack_build_int(const ppp_read* reads, size_t n, uint32_t* H, uint64_t G, uint32_t* occ, Slot slot, int mode,
                                   unsigned long long* counters, cudaStream_t s, int* launches);
cudaError_t launch_unpack_reads(const uint32_t* keys, const uint8_t* meta, const uint32_t* nhTable, size_t n, ppp_read* out, cudaStream_t s, int* launches);
struct SortScratch {
    uint32_t *keysA, *keysB, *valsA, *valsB, *blockHist, *scanScratch;
    size_t capacity, blockHistWords, scanScratchWords;
};
size_t sort_block_hist_words(size_t n);
// -m weighted, in two halves so that the sort of one batch can overlap the (latency-bound) fold of the previous one:
// the sort leaves (position, meta) records ordered by position, file order within a position; the fold adds them to H.
struct SortedRun {
    uint32_t *keys, *vals, *spareA, *spareB;
    size_t n;
};
cudaError_t launch_stack_sort(const ppp_read* reads, size_t n, uint64_t G, Slot slot, unsigned long long* counters, const SortScratch& scr,
                              cudaStream_t s, int* launches, SortedRun* out);
// -m weighted, a batch of at most kSmallBatch records: sort and fold in one launch of one block.
constexpr int kSmallBatch = 4096;
cudaError_t launch_small_batch(const ppp_read* reads, size_t n, uint32_t* H, uint64_t G, uint32_t* occ, Slot slot, unsigned long long* counters,
                               cudaStream_t s, int* launches);
cudaError_t launch_stack_fold(const SortedRun& r, uint32_t* H, uint64_t G, uint32_t* occ, size_t capacity, cudaStream_t s, int* launches);

// stacks_finish: dense accumulation arrays -> compact stack store (compact.cu)
cudaError_t launch_tile_popc(const uint32_t* occ, uint64_t G, uint32_t nTiles, uint32_t* cntP, uint32_t* cntM, cudaStream_t s, int* launches);
// zeroes the words of H whose occupancy bits are set, and the bitmaps (occWords: 32-bit words of both strands' bitmaps)
cudaError_t launch_sparse_clear(uint32_t* H, uint32_t* occ, size_t occWords, cudaStream_t s, int* launches);
cudaError_t launch_compact_write(const uint32_t* H, const uint32_t* occ, uint64_t G, uint32_t nTiles, const uint32_t* offP, const uint32_t* offM,
                                 uint32_t* cmpP, uint32_t* cmpM, cudaStream_t s, int* launches);

cudaError_t launch_scan(const ScanParams& p, int repr, cudaStream_t s, int* launches);
// Multi-GPU exchange of the height histogram in ONE collective.  `gather` holds one slot of kTallHeader + cap words
// per rank: [count of tall stacks, local maximum height, the rank's low bins (kScanLowBins x u64), tall heights...];
// every rank fills its own slot of a zeroed buffer and a sum all-reduce turns that into an all-gather.  The kernel
// sums the low bins, adds every listed tall height to freqTail and leaves the global maximum height; *overflow
// becomes non-zero when some rank had more than `cap` tall stacks (the caller then falls back to reducing the tables).
constexpr uint32_t kTallLow = 2;                            // word offset of the low bins inside a slot
constexpr uint32_t kTallHeader = kTallLow + 2 * kScanLowBins;
cudaError_t launch_tall_apply(const uint32_t* gather, int world, uint32_t cap, unsigned long long* freqLow, uint32_t* freqTail, uint32_t* maxHeight,
                              uint32_t* overflow, uint32_t* maxTail, cudaStream_t s, int* launches);
// (only where the tail counts were all-reduced as dense tables) the largest of freqTail[kScanLowBins .. maxHeight]
cudaError_t launch_tail_max(const uint32_t* freqTail, const uint32_t* maxHeight, uint32_t* maxTail, cudaStream_t s, int* launches);
// `expected`: the host's idea of the number of records (sizes the grid; the kernel reads the exact count from `totals`)
// the largest height frequency (-> state->maxCount) and the bin thresholds / parameters for it
cudaError_t launch_bin_thresholds(PassState* state, float* thresholds, BinParams* params, cudaStream_t s, int* launches);
cudaError_t launch_bin(const float* recRp, const float* recRm, const uint8_t* recMisc, uint16_t* recBin, const unsigned long long* totals, uint64_t capacity,
                       uint64_t expected, FreqTables freq, const float* thresholds, const BinParams* params, int W, uint32_t* grouped, cudaStream_t s, int* launches);
cudaError_t launch_overflow_flag(const unsigned long long* totals, uint64_t capacity, uint64_t ppCapacity, uint32_t* flag, cudaStream_t s, int* launches);
cudaError_t launch_pairs_view(const uint32_t* recPos, const float* recRp, const float* recRm, const uint8_t* recMisc, const uint16_t* recBin, size_t n,
                              const float* fdrTable, const uint16_t* binMap, int W, PairRec* pairs, float* pairFdr, cudaStream_t s, int* launches);
cudaError_t launch_collapse(const uint32_t* grouped, int W, float* cells, uint16_t* binMap, uint32_t* nBins, uint16_t* groupStart, uint32_t* binOcc,
                            cudaStream_t s, int* launches);
size_t collapse_occ_words();
cudaError_t launch_fdr_table(const float* cells, const uint32_t* nBins, uint32_t maxBins, int W, int ppIdx, const double* xs, const double* ws, int nSteps,
                             float* fdr, cudaStream_t s, int* launches);
struct LocusOut {
    uint32_t contig, position;
    float fdr, rp, rm;
};
// K4, per-locus half (single pass, ordered): see score.cu.  `tileState` holds score_tiles(ppCapacity) words and is zeroed,
// together with `state`, before every launch.
cudaError_t launch_score(const PairRec* pp, const unsigned long long* totals, uint64_t ppCapacity, uint64_t expected, ScoreState* state,
                         unsigned long long* tileState, FreqTables freq, const float* thresholds, const BinParams* params, const float* fdrTable,
                         const uint16_t* binMap, int W, int ppIdx, uint32_t minHeight, const Slot* slots, uint32_t nSlots, LocusOut* loci, cudaStream_t s,
                         int* launches);
size_t score_tiles(uint64_t records);

}  // namespace pppk
