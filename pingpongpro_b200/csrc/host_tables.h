// host_tables.h -- the pieces of the path that must be evaluated with the HOST libm (the reference's results depend on
// glibc's log10f and exp): the height-score bin thresholds and the Gaussian table of the FDR integral.
#pragma once
#include <vector>

namespace pppk {

// thresholds[b-1] = smallest float hs in [1, maxFreq^2] whose height-score bin (pingpongpro.cpp:591-594) is >= b,
// b = 1..999; +inf when no product reaches bin b.  The device then computes the bin as the number of thresholds <= hs.
void computeThresholds(float maxFreq, float* thr);
// False when the bin is found non-monotone around a threshold (never observed with glibc).
bool validateThresholds(float maxFreq, const float* thr);
// x and w(x) of the integral at pingpongpro.cpp:724-734.
void gaussianTable(std::vector<double>& xs, std::vector<double>& ws);

}  // namespace pppk
