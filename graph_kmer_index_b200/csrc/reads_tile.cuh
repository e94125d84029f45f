// reads_tile.cuh -- 2-bit packing of ASCII reads staged in shared memory.
//
// Shared by the hashing kernel (K1, hash.cu) and the fused count kernel (K1->K3, count.cu).  Both work on warp-owned tiles of
// consecutive equal-length reads: when the batch is dense (row_stride == read_len) and 16-byte aligned a tile is one contiguous
// span fetched with a single TMA bulk copy (cp.async.bulk -> UBLKCP) that completes on the warp's mbarrier; the next tile's copy is
// issued as soon as the current one is packed.  Otherwise the warp copies its rows itself.
//
// After packing, read r of the tile owns `words` 64-bit words in codes[] and valid[]:
//   base i -> bits 2*(i%32).. of word i/32;  codes: a0 c1 g2 t3 (flat_kmers.py:134-145), other -> 0
//   valid: 0b11 if the byte is one of ACGTacgt else 0b00.  One zero pad word follows the last base.
#pragma once
#include "common.cuh"

namespace gki {

// SWAR encode of 4 ASCII bytes -> 8 bits of codes, 8 bits of validity
__device__ __forceinline__ void encode4(uint32_t w, uint32_t &c8, uint32_t &v8) {
    uint32_t lc = w | 0x20202020u;
    {   // fast path, all four bytes in ACGTacgt: code = (bit1 ^ bit2, bit2 ^ bit3) of the letter gives a0 c1 g2 t3; the letter
        // rebuilt from the code ('a' + {0, 2, 6, 19}[code]) must equal the input; one multiply gathers the four 2-bit codes
        uint32_t c = ((lc >> 1) ^ (lc >> 2)) & 0x03030303u;
        uint32_t c1 = (c >> 1) & 0x01010101u;
        uint32_t expect = 0x61616161u + (c + c1) * 2u + (c & c1) * 11u;
        if (expect == lc) {
            c8 = (c * 0x01041040u) >> 24;
            v8 = 0xFFu;
            return;
        }
    }
    uint32_t v = __vcmpeq4(lc, 0x61616161u) | __vcmpeq4(lc, 0x63636363u) | __vcmpeq4(lc, 0x67676767u) |
                 __vcmpeq4(lc, 0x74747474u);
    uint32_t x = (lc >> 1) & 0x03030303u;
    uint32_t code = (x ^ (x >> 1)) & 0x03030303u & v;
    uint32_t vm = v & 0x03030303u;
    c8 = (code | (code >> 6) | (code >> 12) | (code >> 18)) & 0xFFu;
    v8 = (vm | (vm >> 6) | (vm >> 12) | (vm >> 18)) & 0xFFu;
}

}  // namespace gki
