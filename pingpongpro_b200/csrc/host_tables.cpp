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

static inline float asFloat(uint32_t bits) {
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static inline uint32_t asBits(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
}

// Smallest positive float x with op(x) >= target, for a monotone non-decreasing op; `guess` is a nearby float.
// Positive floats are ordered like their bit patterns, so the search steps over patterns.
template <typename Op>
static inline float smallestAtLeast(float guess, float target, Op op) {
    uint32_t bits = asBits(guess);
    if (bits < 1u) bits = 1u;
    while (bits > 1u && op(asFloat(bits - 1u)) >= target) --bits;
    while (bits < 0x7f7fffffu && op(asFloat(bits)) < target) ++bits;
    return asFloat(bits);
}

// thresholds[b-1] = smallest float hs in [1, maxFreq^2] whose bin is >= b (b = 1..999); +inf when no product
// reaches bin b.  bin() is a monotone step function of hs, so the device computes it as the
// number of thresholds <= hs.  The search assumes that monotonicity; validateThresholds() checks it around every step
// afterwards (off the critical path: the caller runs it while the binning kernel is executing).
//
// This function sits between the scan and the binning kernel on every pass, and a chain of dependent log10f probes
// (gallop + bisection, ~60 cycles each) used to cost 50-100 us.  Now:
//   1. everything behind the log10f of :593 is plain float arithmetic, so the smallest logarithm L_b that reaches bin b
//      (fl(fl(L / maxHs) * 999) >= b - 0.5) is found WITHOUT libm, independently for every b;
//   2. a correctly rounded log10f returns L_b or more from 10^(midpoint of L_b and its predecessor) on: the smallest
//      float at or above that power is the candidate c_b (the powers follow from each other by one multiplication);
//   3. bin(c_b) and bin(c_b - 1 ulp) are evaluated for all b in one loop of independent calls (throughput, not latency);
//   4. c_b IS the threshold when bin(c_b) >= b > bin(c_b - 1 ulp) -- by definition, no assumption about the libm
//      involved; whatever fails that test (log10f not rounding to nearest there) goes through the exact search.
// The candidates only decide how much work the exact search has left, never the result.
void computeThresholds(float maxFreq, float* thr) {
    constexpr int kB = PPP_HEIGHT_SCORE_BINS - 1;  // thresholds
    const float maxHs = log10f(maxFreq * maxFreq);  // :551
    const float top = maxFreq * maxFreq;
    const uint32_t loBits = asBits(1.0f), topBits = asBits(top);
    auto binAt = [&](uint32_t bits) { return heightScoreBin(asFloat(bits), maxHs); };
    const int topBin = binAt(topBits);
    const float scale = (float)(size_t)kB;
    auto timesScale = [&](float q) { return q * scale; };
    auto overMaxHs = [&](float L) { return L / maxHs; };

    double mid[kB];
    for (int b = 1; b <= kB; ++b) {  // 1.
        const float target = (float)b - 0.5f;
        const float q = smallestAtLeast(target / scale, target, timesScale);
        const float L = smallestAtLeast(q * maxHs, q, overMaxHs);
        const uint32_t Lbits = asBits(L);
        mid[b - 1] = Lbits > 1u ? 0.5 * ((double)L + (double)asFloat(Lbits - 1u)) : (double)L;
    }
    uint32_t cand[kB];
    {  // 2.
        const double ln10 = 2.302585092994045684;
        const double step = (double)maxHs / (double)kB;  // nominal distance of consecutive thresholds in log10
        const double ratio = pow(10.0, step);
        double p = pow(10.0, mid[0]);
        for (int b = 1; b <= kB; ++b) {
            if (b > 1) {  // the bulk of the step through the common ratio, the (tiny) rest to second order
                const double d = ((mid[b - 1] - mid[b - 2]) - step) * ln10;
                p *= ratio * (1.0 + d + 0.5 * d * d);
            }
            float f = (float)p;
            uint32_t c = asBits(f);
            if ((double)f < p) ++c;  // smallest float >= 10^mid
            if (!(f >= 1.0f) || c < loBits + 1u) c = loBits + 1u;
            if (c > topBits) c = topBits;
            cand[b - 1] = c;
        }
    }
    int binHere[kB], binBelow[kB];
    for (int b = 1; b <= kB; ++b) {  // 3.
        binHere[b - 1] = binAt(cand[b - 1]);
        binBelow[b - 1] = binAt(cand[b - 1] - 1u);
    }
    uint32_t from = loBits;
    for (int b = 1; b <= kB; ++b) {  // 4.
        if (topBin < b) { thr[b - 1] = std::numeric_limits<float>::infinity(); continue; }
        uint32_t g = cand[b - 1];
        if (g >= from && binHere[b - 1] >= b && (g == from || binBelow[b - 1] < b)) {
            thr[b - 1] = asFloat(g);
            from = g;
            continue;
        }
        // the usual misses are one float off (log10f is faithfully, not correctly, rounded): one more probe settles them
        if (g > from + 1u && binHere[b - 1] >= b && binBelow[b - 1] >= b && binAt(g - 2u) < b) {
            thr[b - 1] = asFloat(g - 1u);
            from = g - 1u;
            continue;
        }
        if (g >= from && g < topBits && binHere[b - 1] < b && binAt(g + 1u) >= b) {
            thr[b - 1] = asFloat(g + 1u);
            from = g + 1u;
            continue;
        }
        // exact search: gallop from the candidate, then bisection on float bit patterns
        if (g < from) g = from;
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
            uint32_t m = lo + (hi - lo) / 2;
            if (binAt(m) >= b) hi = m; else lo = m + 1;
        }
        thr[b - 1] = asFloat(hi);
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
