// local_group.h -- the collective of SURVEY §8(e) for contexts that live in ONE process (one host thread per context,
// any mix of devices): an all-reduce through host memory behind the ppp_allreduce_fn hook.  The exchanged buffers are
// small (the height histogram and the 336 KB of grouped counts, twice per run), so staging them through the host costs
// microseconds; processes on different GPUs use the NCCL binding instead (ppp_nccl_attach).
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <vector>

namespace pppk {

class LocalGroup {
  public:
    explicit LocalGroup(int nRanks) : n_(nRanks), slots_(nRanks) {}
    int size() const { return n_; }
    // In-place all-reduce of `count` elements (dtype 0 = uint32, 1 = uint64; op 0 = sum, 1 = max) of a device buffer;
    // called by every rank's thread, blocks until all have contributed.  0 = ok, 1 = CUDA error, 2 = aborted / misuse.
    int allReduce(int rank, void* deviceBuffer, size_t count, int dtype, int op, cudaStream_t stream);
    // Wakes every waiting rank with an error (a rank that failed elsewhere will never arrive).
    void abort();

  private:
    bool arrive(std::unique_lock<std::mutex>& lock);  // barrier; false when aborted
    const int n_;
    std::mutex mu_;
    std::condition_variable cv_;
    int waiting_ = 0;
    uint64_t generation_ = 0;
    bool aborted_ = false;
    std::vector<std::vector<uint8_t> > slots_;  // one staging buffer per rank
};

}  // namespace pppk
