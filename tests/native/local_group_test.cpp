// local_group_test.cpp -- CPU test of the in-process collective (csrc/local_group.cpp): the two CUDA calls it makes are
// replaced by host copies here, so that the barrier / reduction logic can run under ThreadSanitizer without a GPU.
// Built and run by tests/test_local_group.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "local_group.h"

extern "C" {
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t count, cudaMemcpyKind, cudaStream_t) {
    memcpy(dst, src, count);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
}

using pppk::LocalGroup;

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

int main() {
    // 1. many rounds of sums and maxima of both widths, every rank must see the same reduced buffer
    for (int n : {1, 2, 5, 8}) {
        LocalGroup g(n);
        const int rounds = 150;
        std::vector<std::thread> th;
        std::vector<int> bad(n, 0);
        for (int r = 0; r < n; ++r)
            th.emplace_back([&, r] {
                for (int it = 0; it < rounds; ++it) {
                    std::mt19937_64 shape(1000 + it);  // same on every rank: the ranks agree about the collective
                    const size_t count = 1 + shape() % 5000;
                    const int dtype = (int)(shape() % 2), op = (int)(shape() % 2);
                    std::mt19937_64 mine(77 * it + r);
                    std::vector<uint64_t> expect(count, 0);
                    for (int q = 0; q < n; ++q) {
                        std::mt19937_64 other(77 * it + q);
                        for (size_t i = 0; i < count; ++i) {
                            uint64_t v = other() % 100000;
                            expect[i] = op == 1 ? std::max(expect[i], v) : expect[i] + v;
                        }
                    }
                    if (dtype == 1) {
                        std::vector<uint64_t> buf(count);
                        for (auto& v : buf) v = mine() % 100000;
                        if (g.allReduce(r, buf.data(), count, 1, op, nullptr) != 0 || buf != expect) bad[r]++;
                    } else {
                        std::vector<uint32_t> buf(count);
                        for (auto& v : buf) v = (uint32_t)(mine() % 100000);
                        bool ok = g.allReduce(r, buf.data(), count, 0, op, nullptr) == 0;
                        for (size_t i = 0; ok && i < count; ++i) ok = buf[i] == (uint32_t)expect[i];
                        if (!ok) bad[r]++;
                    }
                }
            });
        for (auto& t : th) t.join();
        for (int r = 0; r < n; ++r) CHECK(bad[r] == 0);
    }
    // 2. a rank that never arrives: abort() releases the others with an error, and the group stays unusable
    {
        LocalGroup g(3);
        std::vector<int> rc(2, -1);
        std::vector<std::thread> th;
        for (int r = 0; r < 2; ++r)
            th.emplace_back([&, r] {
                std::vector<uint32_t> buf(10, 1);
                rc[r] = g.allReduce(r, buf.data(), buf.size(), 0, 0, nullptr);
            });
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        g.abort();
        for (auto& t : th) t.join();
        CHECK(rc[0] == 2 && rc[1] == 2);
        std::vector<uint32_t> buf(4, 1);
        CHECK(g.allReduce(2, buf.data(), buf.size(), 0, 0, nullptr) == 2);
    }
    // 3. ranks that disagree about the size of the collective get an error instead of a wrong sum
    {
        LocalGroup g(2);
        int rc[2] = {-1, -1};
        std::thread a([&] { std::vector<uint32_t> b(8, 1); rc[0] = g.allReduce(0, b.data(), 8, 0, 0, nullptr); });
        std::thread b([&] { std::vector<uint32_t> b2(9, 1); rc[1] = g.allReduce(1, b2.data(), 9, 0, 0, nullptr); });
        a.join();
        b.join();
        CHECK(rc[0] == 2 && rc[1] == 2);
    }
    // 4. argument checks
    {
        LocalGroup g(1);
        uint32_t x = 5;
        CHECK(g.allReduce(1, &x, 1, 0, 0, nullptr) == 2 && g.allReduce(0, &x, 1, 2, 0, nullptr) == 2 && g.allReduce(0, &x, 1, 0, 3, nullptr) == 2);
        CHECK(g.allReduce(0, &x, 1, 0, 0, nullptr) == 0 && x == 5);
    }
    if (failures) return 1;
    printf("local_group_test: ok\n");
    return 0;
}
