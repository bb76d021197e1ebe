// fadd_chain.h -- a chain of single-precision additions `h = h + w_k` (round to nearest even, positive operands) as a
// composition of integer maps, for host and device.
//
// `stack.reads += readWeight` of pingpongpro.cpp:487 is a float accumulation in file order; a stack of 10^6 reads is a
// chain of 10^6 dependent additions (4.5 cycles each on the device: k_fold_long).  The chain becomes parallel once it is
// looked at one BINADE at a time.  Inside the binade [2^e, 2^(e+1)) every float is M * u with u = 2^(e-23) and an integer
// mantissa M in [2^23, 2^24), and the bit pattern of the float grows by one when M does.  Adding w = q * u + f * u
// (q integer, 0 <= f < 1) rounds to
//      M + q        if f < 1/2            M + q + 1   if f > 1/2
//      the even one of the two             if f = 1/2  (the tie: it depends on the parity of M + q)
// as long as the result stays below 2^(e+1).  So within a binade one addition is the integer map
//      B -> B + (B even ? a0 : a1)        on the bit pattern B of h,
// with a0 = a1 except at a tie.  Maps of this form are closed under composition (the parity after the first map is known
// from the parity before it), composition is associative, hence a run of additions can be reduced by a parallel scan and
// applied to h in one step.  All increments are >= 0, so if the final pattern is still inside the binade every
// intermediate one was, and the composed map was valid; if not (h crosses into the next binade: at most ~40 times in the
// life of a stack), the caller redoes a short piece with real additions and continues in the new binade.
// Checked against real float additions in tests/test_host_tables.py (host) and by the stack-builder parity tests (device).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PPP_FC_HD __host__ __device__ __forceinline__
#else
#define PPP_FC_HD inline
#endif

namespace pppk {

struct AddMap {
    uint32_t a0, a1;  // increment of the bit pattern for an even / an odd mantissa
};

constexpr uint32_t kAddMapPoison = 0x40000000u;  // "not a map of this binade": larger than any mantissa, survives compose()

// The map of `h + w` for every h whose biased exponent field is hExp (1..254).  wBits: a float >= 0.  Anything that
// cannot be expressed (w not below the binade, w infinite/NaN/subnormal, h zero/subnormal) gives the poison map,
// which makes addmap_crosses() true for every h -- the caller then adds for real.
PPP_FC_HD AddMap addmap_of(uint32_t hExp, uint32_t wBits) {
    AddMap m;
    if (wBits == 0u || hExp == 255u) { m.a0 = m.a1 = 0u; return m; }  // + 0.0f is exact; +inf stays +inf whatever is added
    const uint32_t wExp = wBits >> 23;                // (sign bit clear: weights are positive)
    const int s = (int)hExp - (int)wExp;              // u = 2^(hExp - 150), w = Mw * 2^(wExp - 150)  ->  w / u = Mw / 2^s
    if (wExp == 0u || wExp >= 255u || hExp == 0u || s < 1) { m.a0 = m.a1 = kAddMapPoison; return m; }
    if (s >= 25) { m.a0 = m.a1 = 0u; return m; }      // w < u / 2: h does not move
    const uint32_t Mw = (wBits & 0x7fffffu) | 0x800000u;
    const uint32_t q = Mw >> s, rem = Mw & ((1u << s) - 1u), half = 1u << (s - 1);
    if (rem != half) { m.a0 = m.a1 = q + (rem > half ? 1u : 0u); return m; }
    m.a0 = q + (q & 1u);          // even M: M + q has the parity of q; round to the even neighbour
    m.a1 = q + ((q & 1u) ^ 1u);   // odd M
    return m;
}

// f first, then g.
PPP_FC_HD AddMap addmap_compose(AddMap f, AddMap g) {
    AddMap r;
    const uint32_t r0 = f.a0 + ((f.a0 & 1u) ? g.a1 : g.a0);         // even start: parity after f is that of f.a0
    const uint32_t r1 = f.a1 + ((f.a1 & 1u) ? g.a0 : g.a1);         // odd start: parity after f is 1 ^ (f.a1 & 1)
    r.a0 = r0 < kAddMapPoison ? r0 : kAddMapPoison;                 // (both operands <= poison: no wrap)
    r.a1 = r1 < kAddMapPoison ? r1 : kAddMapPoison;
    return r;
}

PPP_FC_HD uint32_t addmap_increment(uint32_t hBits, AddMap m) { return (hBits & 1u) ? m.a1 : m.a0; }
// true: applying m to h would leave h's binade (or m is poison) -- the composed map says nothing about the result.
PPP_FC_HD bool addmap_crosses(uint32_t hBits, AddMap m) { return (hBits & 0x7fffffu) + addmap_increment(hBits, m) >= 0x800000u; }
PPP_FC_HD uint32_t addmap_apply(uint32_t hBits, AddMap m) { return hBits + addmap_increment(hBits, m); }

}  // namespace pppk
