// local_group.cpp -- see local_group.h.
#include "local_group.h"

#include <algorithm>
#include <cstring>

namespace pppk {

bool LocalGroup::arrive(std::unique_lock<std::mutex>& lock) {
    if (aborted_) return false;
    const uint64_t gen = generation_;
    if (++waiting_ == n_) {
        waiting_ = 0;
        ++generation_;
        cv_.notify_all();
        return true;
    }
    cv_.wait(lock, [&] { return aborted_ || generation_ != gen; });
    return !aborted_;
}

void LocalGroup::abort() {
    std::lock_guard<std::mutex> l(mu_);
    aborted_ = true;
    cv_.notify_all();
}

template <typename T>
static void reduceInto(std::vector<uint8_t>& acc, const std::vector<uint8_t>& other, size_t count, int op) {
    T* a = reinterpret_cast<T*>(acc.data());
    const T* b = reinterpret_cast<const T*>(other.data());
    if (op == 1)
        for (size_t i = 0; i < count; ++i) a[i] = std::max(a[i], b[i]);
    else
        for (size_t i = 0; i < count; ++i) a[i] += b[i];
}

int LocalGroup::allReduce(int rank, void* deviceBuffer, size_t count, int dtype, int op, cudaStream_t stream) {
    if (rank < 0 || rank >= n_ || (dtype != 0 && dtype != 1) || (op != 0 && op != 1)) return 2;
    const size_t bytes = count * (dtype == 1 ? 8 : 4);
    std::vector<uint8_t>& mine = slots_[rank];
    mine.resize(bytes);
    if (cudaMemcpyAsync(mine.data(), deviceBuffer, bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess) { abort(); return 1; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { abort(); return 1; }
    {
        std::unique_lock<std::mutex> l(mu_);
        if (!arrive(l)) return 2;  // every contribution is in host memory
    }
    for (int r = 0; r < n_; ++r)
        if (slots_[r].size() != bytes) { abort(); return 2; }  // the ranks disagree about the collective
    std::vector<uint8_t> acc(slots_[0]);  // integers: the order of the reduction does not matter, every rank gets the same bits
    for (int r = 1; r < n_; ++r) {
        if (dtype == 1) reduceInto<uint64_t>(acc, slots_[r], count, op);
        else reduceInto<uint32_t>(acc, slots_[r], count, op);
    }
    {
        std::unique_lock<std::mutex> l(mu_);
        if (!arrive(l)) return 2;  // nobody refills its slot before every rank has read all of them
    }
    if (cudaMemcpyAsync(deviceBuffer, acc.data(), bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) { abort(); return 1; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { abort(); return 1; }  // acc lives on this stack frame
    return 0;
}

}  // namespace pppk
