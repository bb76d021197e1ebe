// host_tables.cpp -- see host_tables.h.
#include "host_tables.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "ppp_b200.h"

namespace pppk {

// ---- host libm pieces ---------------------------------------------------------------------------
// heightScoreBin of pingpongpro.cpp:591-594 for one float product, evaluated exactly as the reference does
// (float log10 -> float divide -> float multiply by 999 -> double add 0.5 -> truncate), with the host libm.
static inline int heightScoreBin(float hs, float maxHeightScore) {
    float t = log10f(hs) / maxHeightScore * (float)(size_t)(PPP_HEIGHT_SCORE_BINS - 1);
    return (int)(0.5 + t);
}

// thresholds[b-1] = smallest float hs in [1, maxFreq^2] whose bin is >= b (b = 1..999); +inf when no product
// reaches bin b.  bin() is a monotone step function of hs, so the device computes it as the
// number of thresholds <= hs.  Bisection runs on float bit patterns (order-isomorphic for positive floats).
// The search assumes that bin() is a monotone step function; validateThresholds() checks that assumption around
// every step afterwards (off the critical path: the caller runs it while the binning kernel is executing).
void computeThresholds(float maxFreq, float* thr) {
    float maxHs = log10f(maxFreq * maxFreq);  // :551
    float top = maxFreq * maxFreq;
    uint32_t loBits, topBits;
    float one = 1.0f;
    memcpy(&loBits, &one, 4);
    memcpy(&topBits, &top, 4);
    auto binAt = [&](uint32_t bits) {
        float f;
        memcpy(&f, &bits, 4);
        return heightScoreBin(f, maxHs);
    };
    uint32_t from = loBits;
    const int topBin = binAt(topBits);
    const double ratio = pow(10.0, (double)maxHs / 999.0);
    double lastGuess = 1.0;
    for (int b = 1; b < PPP_HEIGHT_SCORE_BINS; ++b) {
        if (topBin < b) { thr[b - 1] = std::numeric_limits<float>::infinity(); continue; }
        // thresholds form a geometric sequence (bin >= b  <=>  log10(hs) >= (b - 0.5) / 999 * maxHs): the previous exact
        // threshold times the common ratio lands within a few float steps of the next one
        double guessD = b == 1 ? pow(10.0, 0.5 / 999.0 * (double)maxHs) : lastGuess * ratio;
        float guessF = (float)guessD;
        uint32_t g;
        memcpy(&g, &guessF, 4);
        if (!(guessF >= 1.0f) || g < from) g = from;
        if (g > topBits) g = topBits;
        uint32_t lo, hi;  // invariant: bin(lo - 1) < b (or lo == from), bin(hi) >= b
        if (binAt(g) >= b) {
            hi = g;
            uint32_t step = 1;
            lo = from;
            while (hi - from >= step) {
                if (binAt(hi - step) >= b) { hi -= step; step <<= 1; }
                else { lo = hi - step + 1; break; }
            }
        } else {
            lo = g + 1;
            uint32_t step = 1;
            hi = topBits;
            while (topBits - (lo - 1) > step) {
                uint32_t probe = lo - 1 + step;
                if (binAt(probe) >= b) { hi = probe; break; }
                lo = probe + 1;
                step <<= 1;
            }
        }
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            if (binAt(mid) >= b) hi = mid; else lo = mid + 1;
        }
        memcpy(&thr[b - 1], &hi, 4);
        lastGuess = (double)thr[b - 1];
        from = hi;
    }
}

// False when bin() is found non-monotone around a threshold (never observed with glibc): the patterns just below a
// threshold must fall below bin b, the threshold itself and the pattern above it must reach it.
bool validateThresholds(float maxFreq, const float* thr) {
    const float maxHs = log10f(maxFreq * maxFreq);
    const float top = maxFreq * maxFreq;
    uint32_t loBits, topBits;
    const float one = 1.0f;
    memcpy(&loBits, &one, 4);
    memcpy(&topBits, &top, 4);
    auto binAt = [&](uint32_t bits) {
        float f;
        memcpy(&f, &bits, 4);
        return heightScoreBin(f, maxHs);
    };
    for (int b = 1; b < PPP_HEIGHT_SCORE_BINS; ++b) {
        if (std::isinf(thr[b - 1])) continue;
        uint32_t hi;
        memcpy(&hi, &thr[b - 1], 4);
        if (binAt(hi) < b) return false;
        if (hi >= loBits + 1 && (b == 1 || thr[b - 2] < thr[b - 1]) && binAt(hi - 1) >= b) return false;
        if (hi >= loBits + 2 && (b == 1 || thr[b - 2] < thr[b - 1]) && binAt(hi - 2) >= b) return false;
        if (hi + 1 <= topBits && binAt(hi + 1) < b) return false;
        if (b > 1 && thr[b - 2] > thr[b - 1]) return false;
    }
    return true;
}

// x and w(x) of the integral at pingpongpro.cpp:724-734: for (double x = -5; x <= 5; x += 0.01)
//   w = 0.01 * 1 / sqrt(2 * M_PI) * exp(-0.5 * x * x)
void gaussianTable(std::vector<double>& xs, std::vector<double>& ws) {
    const double pi = 3.14159265358979323846;  // :50
    xs.clear();
    ws.clear();
    for (double x = -5.0; x <= +5.0; x += 0.01) {
        xs.push_back(x);
        ws.push_back(0.01 * 1 / sqrt(2 * pi) * exp(-0.5 * x * x));
    }
}

}  // namespace pppk
