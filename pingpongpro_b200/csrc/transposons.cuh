// transposons.cuh -- K5 interface (transposons.cu).
#pragma once
#include <functional>
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

// The same when every shard lives in its OWN process (one process per GPU): called collectively by all ranks with the same
// intervals; `reduce(buffer, words)` sums a device buffer of 32-bit words over the ranks in place (the context's all-reduce
// hook).  The whole-list facts behind the iterator history are summed over the ranks; the sequential single-precision sums
// travel rank by rank in position order (round k: rank k continues the running sums and shares them -- the others
// contribute zeros to the sum).  Every rank ends up with the complete results.
cudaError_t run_transposons_collective(const TransposonJob& job, int rank, int world, const std::function<int(void*, size_t)>& reduce, const ppp_interval* intervals,
                                       size_t n, ppp_tp_result* results, uint32_t* order, std::string& err);

}  // namespace pppk
