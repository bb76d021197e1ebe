// fast_inflate.h -- raw-DEFLATE decoder for BGZF blocks (whole buffer in, whole buffer out).
#pragma once
#include <cstddef>
#include <cstdint>

namespace ppp {

// Decodes the raw DEFLATE stream in[0, inLen) into exactly outLen bytes.  `in` must be readable up to in + inLen (no
// slack needed), `out` writable up to out + outLen.  Returns false when the stream is not a valid DEFLATE stream of
// exactly that size or uses something this decoder does not handle; `out` is then unspecified and the caller decodes
// the block with zlib instead.
bool fast_inflate(const uint8_t* in, size_t inLen, uint8_t* out, size_t outLen);

}  // namespace ppp
