// transposons.cuh -- K5 interface (transposons.cu).
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "ppp_b200.h"

namespace pppk {

struct TransposonJob {
    const PairRec* pairs;    // signatures in (position, overlap) order
    const float* pairFdr;
    uint64_t nPairs;
    int W, ppIdx;
    const std::vector<Slot>* slots;  // per contig (lenMinus == 0: contig not held)
    cudaStream_t stream;
    int* launches;
};

// Fills results[] in the reference's output order (contig, start, end; ties in input order) and order[] with the
// input index of each result.  `err` is set (and cudaSuccess returned) for argument errors.
cudaError_t run_transposons(const TransposonJob& job, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order, std::string& err);

// The same over the contexts of a range-sharded genome: jobs[k] / devices[k] in ascending range order.  Transposons that
// straddle a range boundary get the sums (and the iterator history) of the unsharded run.
cudaError_t run_transposons_sharded(const TransposonJob* jobs, const int* devices, int nShards, const ppp_interval* intervals, size_t n, ppp_tp_result* results,
                                    uint32_t* order, std::string& err);

}  // namespace pppk
