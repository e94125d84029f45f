// reads_tile.cuh -- staging of a tile of equal-length ASCII reads into shared memory and 2-bit packing.
//
// Shared by the hashing kernel (K1, hash.cu) and the fused count kernel (K1->K3, index.cu).
// A tile is TR consecutive reads.  When the batch is dense (row_stride == read_len) and 16-byte aligned
// the tile is one contiguous span and is fetched with a single TMA bulk copy (cp.async.bulk ->
// UBLKCP) that completes on an mbarrier, double-buffered so the next tile streams in while the current
// one is processed.  Otherwise the CTA falls back to cooperative row-wise loads.
//
// After packing, read r of the tile owns `words` 64-bit words in codes[] and valid[]:
//   base i -> bits 2*(i%32).. of word i/32;  codes: a0 c1 g2 t3 (flat_kmers.py:134-145), other -> 0
//   valid: 0b11 if the byte is one of ACGTacgt else 0b00.  One zero pad word follows the last base.
#pragma once
#include "common.cuh"

namespace gki {

struct ReadBatch {
    const uint8_t *reads;
    int64_t n_reads;
    int64_t row_stride;
    int32_t read_len;
    int32_t k;
    int32_t nk;          // read_len - k + 1
    int32_t words;       // ceil(read_len/32) + 1
    int32_t tile_reads;  // TR
    int32_t bulk_ok;     // dense + aligned: full tiles use TMA bulk copies
    int64_t n_tiles;
};

struct TileSmem {
    uint64_t *bars;      // 2 mbarriers
    uint8_t *ascii0;     // 2 stages of TR*read_len (+16 slack) bytes, ascii_stride apart
    uint32_t ascii_stride;
    __device__ __forceinline__ uint8_t *ascii(int stage) const { return ascii0 + (size_t)stage * ascii_stride; }
    uint64_t *codes;     // TR*words
    uint64_t *valid;     // TR*words
};

__host__ __device__ inline size_t tile_ascii_bytes(int tile_reads, int read_len) {
    return (((size_t)tile_reads * read_len + 15) & ~(size_t)15) + 16;
}
inline size_t tile_smem_bytes(int tile_reads, int read_len, int words) {
    return 16 + 2 * tile_ascii_bytes(tile_reads, read_len) + 2 * (size_t)tile_reads * words * 8;
}

__device__ __forceinline__ TileSmem carve_tile_smem(unsigned char *base, const ReadBatch &b) {
    TileSmem t;
    t.bars = (uint64_t *)base;
    size_t ab = tile_ascii_bytes(b.tile_reads, b.read_len);
    t.ascii0 = base + 16;
    t.ascii_stride = (uint32_t)ab;
    t.codes = (uint64_t *)(base + 16 + 2 * ab);
    t.valid = t.codes + (size_t)b.tile_reads * b.words;
    return t;
}

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

// Is tile `tile` fetched with a bulk copy?  (full tiles of a dense aligned batch)
__device__ __forceinline__ bool tile_uses_bulk(const ReadBatch &b, int64_t tile) {
    return b.bulk_ok && (tile + 1) * (int64_t)b.tile_reads <= b.n_reads;
}

// Issue the fetch of `tile` into stage `stage` (one thread).
__device__ __forceinline__ void tile_issue_bulk(const ReadBatch &b, const TileSmem &t, int64_t tile, int stage) {
    uint32_t bytes = (uint32_t)b.tile_reads * (uint32_t)b.read_len;
    fence_proxy_async();
    mbar_expect_tx(&t.bars[stage], bytes);
    bulk_g2s(t.ascii(stage), b.reads + tile * (int64_t)b.tile_reads * b.row_stride, bytes, &t.bars[stage]);
}

// Cooperative fallback fetch (whole CTA): rows copied warp by warp.
__device__ __forceinline__ void tile_load_fallback(const ReadBatch &b, const TileSmem &t, int64_t tile, int stage) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int64_t r0 = tile * (int64_t)b.tile_reads;
    for (int r = warp; r < b.tile_reads; r += nwarps) {
        if (r0 + r >= b.n_reads) break;
        const uint8_t *src = b.reads + (r0 + r) * b.row_stride;
        uint8_t *dst = t.ascii(stage) + (size_t)r * b.read_len;
        for (int i = lane; i < b.read_len; i += 32) dst[i] = __ldg(src + i);
    }
}

// Pack stage `stage` into codes/valid (whole CTA; caller syncs before and after).
__device__ __forceinline__ void tile_pack(const ReadBatch &b, const TileSmem &t, int64_t tile, int stage) {
    int64_t r0 = tile * (int64_t)b.tile_reads;
    int n_here = (int)min((int64_t)b.tile_reads, b.n_reads - r0);
    int tasks = n_here * b.words;
    const uint8_t *ascii = t.ascii(stage);
    for (int task = threadIdx.x; task < tasks; task += blockDim.x) {
        int r = task / b.words, w = task - r * b.words;
        int first = w * 32;
        int nb = min(32, b.read_len - first);   // bases in this word (<= 0 for the pad word)
        uint64_t cw = 0, vw = 0;
        if (nb > 0) {
            uint32_t addr = (uint32_t)r * (uint32_t)b.read_len + (uint32_t)first;
            const uint32_t *aligned = (const uint32_t *)(ascii + (addr & ~3u));
            uint32_t sh = (addr & 3u) * 8u;
            uint32_t lo = aligned[0];
            int nq = (nb + 3) >> 2;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (q < nq) {
                    uint32_t hi = aligned[q + 1];
                    uint32_t word = __funnelshift_r(lo, hi, sh);
                    lo = hi;
                    uint32_t c8, v8;
                    encode4(word, c8, v8);
                    cw |= (uint64_t)c8 << (8 * q);
                    vw |= (uint64_t)v8 << (8 * q);
                }
            }
            if (nb < 32) {
                uint64_t m = (1ull << (2 * nb)) - 1ull;
                cw &= m;
                vw &= m;
            }
        }
        t.codes[(size_t)r * b.words + w] = cw;
        t.valid[(size_t)r * b.words + w] = vw;
    }
}

// Persistent tile loop (whole CTA): tile = blockIdx.x, blockIdx.x + gridDim.x, ...; the next tile's bulk copy
// is in flight while `process(tile, smem)` runs on the packed current one.
template <typename F>
__device__ __forceinline__ void for_each_tile(const ReadBatch &b, unsigned char *smem_raw, F &&process) {
    TileSmem t = carve_tile_smem(smem_raw, b);
    if (threadIdx.x == 0) {
        mbar_init(&t.bars[0], 1);
        mbar_init(&t.bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    int64_t tile = blockIdx.x;
    if (tile < b.n_tiles && threadIdx.x == 0 && tile_uses_bulk(b, tile)) tile_issue_bulk(b, t, tile, 0);
    for (int it = 0; tile < b.n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int64_t next = tile + gridDim.x;
        if (next < b.n_tiles && threadIdx.x == 0 && tile_uses_bulk(b, next)) tile_issue_bulk(b, t, next, s ^ 1);
        if (tile_uses_bulk(b, tile)) {
            if (s == 0) { mbar_wait(&t.bars[0], phase0); phase0 ^= 1; }
            else        { mbar_wait(&t.bars[1], phase1); phase1 ^= 1; }
        } else {
            tile_load_fallback(b, t, tile, s);
            __syncthreads();
        }
        tile_pack(b, t, tile, s);
        __syncthreads();
        process(tile, t);
        __syncthreads();
    }
}

// Host side: pick the tile shape for a batch.
inline int make_read_batch(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, int32_t k,
                           ReadBatch &b, size_t &smem) {
    b.reads = reads;
    b.n_reads = n_reads;
    b.row_stride = row_stride;
    b.read_len = read_len;
    b.k = k;
    b.nk = read_len - k + 1;
    b.words = (read_len + 31) / 32 + 1;
    int tr = 32;
    while (tr > 8 && tile_smem_bytes(tr, read_len, b.words) > 64 * 1024) tr >>= 1;
    smem = tile_smem_bytes(tr, read_len, b.words);
    b.tile_reads = tr;
    b.bulk_ok = (row_stride == read_len) && (((uintptr_t)reads & 15) == 0) && (((int64_t)tr * read_len) % 16 == 0);
    b.n_tiles = (n_reads + tr - 1) / tr;
    return GKI_OK;
}

}  // namespace gki
