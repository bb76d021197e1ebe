// fast_inflate.cpp -- see fast_inflate.h.  A raw-DEFLATE (RFC 1951) decoder for BGZF blocks, written for this project:
// whole-buffer decoding (input and output sizes are known from the BGZF header and trailer), a 64-bit bit buffer
// refilled once per symbol pair, table-driven Huffman decoding with an 11-bit (literal/length) and an 8-bit (distance)
// first-level table plus second-level tables for longer codes, and word-wise match copies.  Anything unexpected --
// invalid code lengths, a symbol that does not exist, a match that reaches before the start or past the end of the
// output, input that runs out -- makes it return false; the caller then decodes the block with zlib, which also owns
// the error reporting.
#include "fast_inflate.h"

#include <cstring>

namespace ppp {

namespace {

constexpr int kLitBits = 11, kDistBits = 8, kMaxCodeLen = 15;
constexpr int kLitSyms = 288, kDistSyms = 32;
constexpr uint32_t kLitTableSize = (1u << kLitBits) + kLitSyms * (1u << (kMaxCodeLen - kLitBits));
constexpr uint32_t kDistTableSize = (1u << kDistBits) + kDistSyms * (1u << (kMaxCodeLen - kDistBits));

// table entry: value (16 bits) | extra bits (4) << 16 | kind (3) << 20 | code length or first-level bits (8) << 24
enum Kind : uint32_t { kInvalid = 0, kLiteral = 1, kLength = 2, kEndOfBlock = 3, kSubTable = 4, kDistance = 5 };
inline uint32_t entry(uint32_t value, uint32_t extra, uint32_t kind, uint32_t len) { return value | (extra << 16) | (kind << 20) | (len << 24); }
inline uint32_t eValue(uint32_t e) { return e & 0xffffu; }
inline uint32_t eExtra(uint32_t e) { return (e >> 16) & 15u; }
inline uint32_t eKind(uint32_t e) { return (e >> 20) & 7u; }
inline uint32_t eLen(uint32_t e) { return e >> 24; }

const uint16_t kLengthBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLengthExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kCodeLengthOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t reverseBits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// What a symbol of the literal/length (isDist = false) or distance alphabet decodes to.
inline uint32_t symbolEntry(bool isDist, int sym, uint32_t len) {
    if (isDist) return sym < 30 ? entry(kDistBase[sym], kDistExtra[sym], kDistance, len) : entry(0, 0, kInvalid, len);
    if (sym < 256) return entry((uint32_t)sym, 0, kLiteral, len);
    if (sym == 256) return entry(0, 0, kEndOfBlock, len);
    if (sym < 286) return entry(kLengthBase[sym - 257], kLengthExtra[sym - 257], kLength, len);
    return entry(0, 0, kInvalid, len);
}

// Canonical Huffman decoding table from code lengths.  False for an over-subscribed code; an incomplete code leaves
// kInvalid entries (decoding one of them fails, which sends the block to zlib).
bool buildTable(const uint8_t* lens, int nSyms, bool isDist, int tableBits, uint32_t* table, uint32_t tableSize) {
    int count[kMaxCodeLen + 1] = {0};
    for (int s = 0; s < nSyms; ++s) ++count[lens[s]];
    count[0] = 0;
    uint32_t nextCode[kMaxCodeLen + 2];
    uint32_t code = 0;
    int64_t left = 1;
    int maxLen = 0;
    for (int l = 1; l <= kMaxCodeLen; ++l) {
        left = left * 2 - count[l];
        if (left < 0) return false;
        code = (code + (uint32_t)count[l - 1]) << 1;
        nextCode[l] = code;
        if (count[l]) maxLen = l;
    }
    const uint32_t primary = 1u << tableBits;
    for (uint32_t i = 0; i < primary; ++i) table[i] = entry(0, 0, kInvalid, 1);
    const int subBits = maxLen > tableBits ? maxLen - tableBits : 0;
    uint32_t used = primary;
    for (int s = 0; s < nSyms; ++s) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t rev = reverseBits(nextCode[len]++, len);
        if (len <= tableBits) {
            const uint32_t e = symbolEntry(isDist, s, (uint32_t)len);
            for (uint32_t i = rev; i < primary; i += 1u << len) table[i] = e;
        } else {
            const uint32_t low = rev & (primary - 1u);
            if (eKind(table[low]) != kSubTable) {
                if (used + (1u << subBits) > tableSize) return false;
                table[low] = entry(used, (uint32_t)subBits, kSubTable, (uint32_t)tableBits);
                for (uint32_t i = 0; i < (1u << subBits); ++i) table[used + i] = entry(0, 0, kInvalid, 1);
                used += 1u << subBits;
            }
            const uint32_t base = eValue(table[low]);
            const uint32_t e = symbolEntry(isDist, s, (uint32_t)(len - tableBits));
            for (uint32_t i = rev >> tableBits; i < (1u << subBits); i += 1u << (len - tableBits)) table[base + i] = e;
        }
    }
    return true;
}

struct Bits {
    const uint8_t* p;
    const uint8_t* end;  // one past the last input byte
    uint64_t buf = 0;    // bits above `cnt` may already hold the next input bytes (re-ORed identically by the next refill)
    int cnt = 0;         // valid bits in buf
    int padded = 0;      // zero bytes appended after the end of the input
    // at least 56 valid bits afterwards
    inline void refill() {
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (p < end) buf |= (uint64_t)*p++ << cnt;
                else ++padded;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n) { uint32_t v = peek(n); drop(n); return v; }
    // True when more bits were consumed than the input held.
    inline bool exhausted() const { return padded * 8 > cnt; }
};

}  // namespace

bool fast_inflate(const uint8_t* in, size_t inLen, uint8_t* out, size_t outLen) {
    static thread_local uint32_t litTable[kLitTableSize];
    static thread_local uint32_t distTable[kDistTableSize];
    Bits b;
    b.p = in;
    b.end = in + inLen;
    uint8_t* o = out;
    uint8_t* const oEnd = out + outLen;
    for (;;) {
        b.refill();
        const uint32_t last = b.take(1), type = b.take(2);
        if (type == 0) {  // stored
            b.drop(b.cnt & 7);  // to the byte boundary
            b.refill();
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ nlen) != 0xffffu || b.exhausted()) return false;
            // the bytes still in the bit buffer come first, then the rest straight from the input
            uint32_t n = len;
            for (int inBuffer = (b.cnt - b.padded * 8) / 8; n && inBuffer > 0; --inBuffer, --n) {  // (cnt is a multiple of 8 here)
                if (o >= oEnd) return false;
                *o++ = (uint8_t)b.take(8);
            }
            if (n) {
                if (b.padded || (size_t)(b.end - b.p) < n || (size_t)(oEnd - o) < n) return false;
                b.buf = 0;
                b.cnt = 0;
                memcpy(o, b.p, n);
                o += n;
                b.p += n;
            }
        } else if (type == 1 || type == 2) {
            uint8_t lens[kLitSyms + kDistSyms];
            int nLit, nDist;
            if (type == 1) {  // fixed codes
                nLit = 288;
                nDist = 32;
                for (int s = 0; s < 144; ++s) lens[s] = 8;
                for (int s = 144; s < 256; ++s) lens[s] = 9;
                for (int s = 256; s < 280; ++s) lens[s] = 7;
                for (int s = 280; s < 288; ++s) lens[s] = 8;
                for (int s = 0; s < 32; ++s) lens[288 + s] = 5;
            } else {
                nLit = (int)b.take(5) + 257;
                nDist = (int)b.take(5) + 1;
                const int nCode = (int)b.take(4) + 4;
                if (nLit > 286 || nDist > 30) return false;
                uint8_t clens[19] = {0};
                for (int i = 0; i < nCode; ++i) {
                    if (b.cnt < 3) b.refill();
                    clens[kCodeLengthOrder[i]] = (uint8_t)b.take(3);
                }
                uint32_t clTable[128 + 19 * 1];
                if (!buildTable(clens, 19, false, 7, clTable, 128)) return false;  // (entries decode as literals 0..18)
                int i = 0;
                while (i < nLit + nDist) {
                    b.refill();
                    const uint32_t e = clTable[b.peek(7)];
                    if (eKind(e) != kLiteral) return false;
                    b.drop((int)eLen(e));
                    const uint32_t sym = eValue(e);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) return false;
                        val = lens[i - 1];
                        rep = 3 + b.take(2);
                    } else if (sym == 17) rep = 3 + b.take(3);
                    else rep = 11 + b.take(7);
                    if (i + (int)rep > nLit + nDist) return false;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (b.exhausted() || lens[256] == 0) return false;
                // the distance lengths follow the literal/length ones directly: move them to a fixed offset
                uint8_t tmp[kDistSyms];
                memcpy(tmp, lens + nLit, (size_t)nDist);
                memset(lens + nLit, 0, (size_t)(kLitSyms - nLit));
                memcpy(lens + kLitSyms, tmp, (size_t)nDist);
                memset(lens + kLitSyms + nDist, 0, (size_t)(kDistSyms - nDist));
            }
            if (!buildTable(lens, kLitSyms, false, kLitBits, litTable, kLitTableSize)) return false;
            if (!buildTable(lens + kLitSyms, kDistSyms, true, kDistBits, distTable, kDistTableSize)) return false;
            // Fast loop: far enough from both ends that the refill, the literal stores and the word-wise copies need no
            // bounds checks (a match is at most 258 bytes, a refill reads 8 bytes, a copy may run 7 bytes over).
            // The entry of the NEXT symbol is looked up as soon as its bits are known (before a match is copied), so that
            // the table load overlaps the copy; after an unchecked refill all 64 bits of the buffer are real input bits,
            // which makes that early look-up valid even when fewer than 11 of them are accounted for in `cnt`.
            bool endOfBlock = false;
            if (b.end - b.p >= 16 && oEnd - o >= 258 + 16) {
                constexpr uint64_t kLitMask = (1u << kLitBits) - 1u;
                uint64_t w;
                memcpy(&w, b.p, 8);
                b.buf |= w << b.cnt;
                b.p += (63 - b.cnt) >> 3;
                b.cnt |= 56;
                uint32_t e = litTable[b.buf & kLitMask];
                do {
                    memcpy(&w, b.p, 8);
                    b.buf |= w << b.cnt;
                    b.p += (63 - b.cnt) >> 3;
                    b.cnt |= 56;
                    if (eKind(e) == kLiteral) {  // up to three literals per refill (15 + 11 + 11 bits)
                        b.drop((int)eLen(e));
                        *o++ = (uint8_t)e;
                        e = litTable[b.buf & kLitMask];
                        if (eKind(e) != kLiteral) continue;
                        b.drop((int)eLen(e));
                        *o++ = (uint8_t)e;
                        e = litTable[b.buf & kLitMask];
                        if (eKind(e) != kLiteral) continue;
                        b.drop((int)eLen(e));
                        *o++ = (uint8_t)e;
                        e = litTable[b.buf & kLitMask];
                        continue;
                    }
                    if (eKind(e) == kSubTable) {
                        b.drop(kLitBits);
                        e = litTable[eValue(e) + b.peek((int)eExtra(e))];
                        if (eKind(e) == kLiteral) {
                            b.drop((int)eLen(e));
                            *o++ = (uint8_t)e;
                            e = litTable[b.buf & kLitMask];
                            continue;
                        }
                    }
                    b.drop((int)eLen(e));
                    if (eKind(e) != kLength) {
                        if (eKind(e) != kEndOfBlock) return false;
                        endOfBlock = true;
                        break;
                    }
                    const uint32_t length = eValue(e) + b.take((int)eExtra(e));
                    uint32_t d = distTable[b.buf & ((1u << kDistBits) - 1u)];
                    if (eKind(d) == kSubTable) {
                        b.drop(kDistBits);
                        d = distTable[eValue(d) + b.peek((int)eExtra(d))];
                    }
                    b.drop((int)eLen(d));
                    if (eKind(d) != kDistance) return false;
                    const uint32_t dist = eValue(d) + b.take((int)eExtra(d));
                    e = litTable[b.buf & kLitMask];  // next symbol: its load overlaps the copy
                    if (dist > (size_t)(o - out)) return false;
                    const uint8_t* src = o - dist;
                    uint8_t* dst = o;
                    o += length;
                    if (dist >= 8) {  // 8-byte steps never read a byte before it is written when dist >= 8
                        memcpy(dst, src, 8);
                        memcpy(dst + 8, src + 8, 8);
                        if (length > 16) {
                            dst += 16;
                            src += 16;
                            do {
                                memcpy(dst, src, 8);
                                dst += 8;
                                src += 8;
                            } while (dst < o);
                        }
                    } else if (dist == 1) {
                        const uint64_t v = 0x0101010101010101ull * *src;
                        do {
                            memcpy(dst, &v, 8);
                            dst += 8;
                        } while (dst < o);
                    } else {
                        do *dst++ = *src++; while (dst < o);
                    }
                } while (b.end - b.p >= 16 && oEnd - o >= 258 + 16);
            }
            // Careful loop: the last few hundred bytes of a block, with every bound checked.
            while (!endOfBlock) {
                b.refill();
                uint32_t e = litTable[b.peek(kLitBits)];
                if (eKind(e) == kSubTable) {
                    b.drop(kLitBits);
                    e = litTable[eValue(e) + b.peek((int)eExtra(e))];
                }
                b.drop((int)eLen(e));
                const uint32_t kind = eKind(e);
                if (kind == kLiteral) {
                    if (o >= oEnd) return false;
                    *o++ = (uint8_t)eValue(e);
                    continue;
                }
                if (kind == kEndOfBlock) break;
                if (kind != kLength) return false;
                const uint32_t length = eValue(e) + b.take((int)eExtra(e));
                // at most 15 + 5 bits are gone: 36 left >= 15 + 13 for the distance
                uint32_t d = distTable[b.peek(kDistBits)];
                if (eKind(d) == kSubTable) {
                    b.drop(kDistBits);
                    d = distTable[eValue(d) + b.peek((int)eExtra(d))];
                }
                b.drop((int)eLen(d));
                if (eKind(d) != kDistance) return false;
                const uint32_t dist = eValue(d) + b.take((int)eExtra(d));
                if (dist > (size_t)(o - out) || length > (size_t)(oEnd - o)) return false;
                const uint8_t* src = o - dist;
                for (uint32_t k = 0; k < length; ++k) o[k] = src[k];
                o += length;
            }
            if (b.exhausted()) return false;
        } else {
            return false;
        }
        if (last) break;
    }
    return o == oEnd && !b.exhausted();
}

}  // namespace ppp
