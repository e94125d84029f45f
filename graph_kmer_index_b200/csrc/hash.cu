// hash.cu -- K1: 2-bit encoding + k-mer hashing of read batches (forward and reverse complement).
//
// Reference semantics: flat_kmers.py:134-145 (encoding), read_kmers.py:67-70 (np.convolve == first base in
// the least significant 2 bits), read_kmers.py:21-26 (reverse strand = hash of the reverse-complemented
// string), kmer_hashing.py:24-65 (hash -> bases, complement, reverse complement).
//
// Roofline: HBM-bound on the hash writes -- per read L bytes in, 2*(L-k+1)*8 bytes out
// (L=150,k=31: 2070 B/read = 8.625 B per emitted hash).  A k-mer is a 2k-bit window of the packed read: a lane
// extracts the first of its four consecutive windows from shared memory with one funnel shift and rolls the other
// three; the reverse-complement hash of a window is a bit reversal (reverse_pairs(~x & valid)) and rolls the other way.
#include "reads_tile.cuh"

namespace gki {

constexpr int HASH_THREADS = 256;

__device__ __forceinline__ void st_global_v4_u64(uint64_t *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {   // one 256-bit store (sm_100)
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// All hashes of a read made only of ACGTacgt (the usual case), by one warp: every lane owns four consecutive windows and writes each
// strand's four hashes with one 256-bit store.  cw: the read's code words; of / orc: its rows of the two outputs.
template <bool FWD, bool RC>
__device__ __forceinline__ void emit_clean_read(const uint64_t *cw, int nk, int k, uint64_t mask, uint64_t *of, uint64_t *orc, int lane) {
    const bool wide_f = FWD && (((uintptr_t)of & 31) == 0), wide_r = RC && (((uintptr_t)(orc + nk) & 31) == 0);
    const uint32_t *cw32 = (const uint32_t *)cw;
    for (int i0 = lane * 4; i0 < nk; i0 += 128) {
        // the lane's first window starts on a byte boundary of the packed read (4 bases): three byte permutes bring 96 bits from
        // there into place, and window u is those bits shifted by the constant 2u
        const int wi = i0 >> 4;
        const uint32_t sel = 0x3210u + 0x1111u * (((uint32_t)i0 >> 2) & 3u);
        const uint32_t w0 = cw32[wi], w1 = cw32[wi + 1], w2 = cw32[wi + 2], w3 = cw32[wi + 3];   // (beyond the read: pad word, next read, `valid`)
        const uint32_t a0 = __byte_perm(w0, w1, sel), a1 = __byte_perm(w1, w2, sel), a2 = __byte_perm(w2, w3, sel);
        const int n_valid = min(4, nk - i0);
        uint64_t x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t lo = u ? __funnelshift_r(a0, a1, 2 * u) : a0, hi = u ? __funnelshift_r(a1, a2, 2 * u) : a1;
            x[u] = (((uint64_t)hi << 32) | lo) & mask;
        }
        uint64_t y0 = 0;
        uint32_t tops = 0;      // the bases that enter on top of windows 1..3 (two bits each)
        if (RC) {
            y0 = revcomp_hash(x[0], k);
            const int ts = 2 * (k - 1);
            tops = (uint32_t)((x[1] >> ts) & 3u) | ((uint32_t)((x[2] >> ts) & 3u) << 2) | ((uint32_t)((x[3] >> ts) & 3u) << 4);
        }
        if (FWD) {
            if (n_valid == 4 && wide_f) st_global_v4_u64(of + i0, x[0], x[1], x[2], x[3]);
            else
                for (int u = 0; u < n_valid; u++) of[i0 + u] = x[u];
        }
        if (RC) {   // window i of the read is window nk-1-i of the reverse-complemented read; its hash rolls the other way
            const uint64_t y1 = ((y0 << 2) | (3u - (tops & 3u))) & mask;
            const uint64_t y2 = ((y1 << 2) | (3u - ((tops >> 2) & 3u))) & mask;
            const uint64_t y3 = ((y2 << 2) | (3u - (tops >> 4))) & mask;
            if (n_valid == 4 && wide_r) st_global_v4_u64(orc + (nk - 4 - i0), y3, y2, y1, y0);
            else {
                const uint64_t y[4] = {y0, y1, y2, y3};
                for (int u = 0; u < n_valid; u++) orc[nk - 1 - i0 - u] = y[u];
            }
        }
    }
}

// Warp-autonomous (round 2).  The first version worked on CTA-wide tiles of 32 reads between three barriers, rolled the windows with
// 64-bit variable shifts and spent ~280 warp instructions per read whether it wrote one strand or two: 0.86 of the HBM roofline with
// both strands, 0.49 with the forward strand alone -- the call the reference exposes (read_kmers.py:67-70).  Now 0.92 / 0.98
// (profiles/r2/k1_warp_autonomous.log).  A warp owns tiles of `rpw` reads: its own mbarrier, its own TMA bulk copy of the ASCII
// rows (issued for the next tile as soon as the current one is packed), a pack phase of 16-base tasks that fill the warp evenly,
// then one read at a time.  FWD / RC: the strands written (the other strand's arithmetic is not compiled in).  The grid is NOT
// sized to the resident CTAs: in one wave of persistent CTAs all warps pack at the same time and store at the same time, and the
// launch is slower the fewer CTAs there are (2 M reads, forward strand: 0.419 ms with 6 x 148 CTAs, 0.365 ms with 48 x 148, 0.345 ms
// = 0.98 of the roofline with a tile per warp; both strands 0.808 / 0.722 / 0.684 ms = 0.92); short-lived CTAs arrive staggered and keep
// the store stream even.  The loop over tiles remains for batches larger than the grid limit.
struct HashWarpBatch {
    const uint8_t *reads;
    int64_t n_reads;
    int64_t row_stride;
    int64_t n_wtiles;       // ceil(n_reads / rpw)
    int32_t read_len;
    int32_t k;
    int32_t nk;             // read_len - k + 1
    int32_t words;          // ceil(read_len / 32) + 1
    int32_t rpw;            // reads per warp tile
    int32_t bulk_ok;        // dense + aligned: full tiles are one TMA bulk copy
    uint32_t inv_halves;    // ceil(65536 / (2 * words)): task / (2 * words) == task * inv_halves >> 16 for every task of a tile
    uint32_t stage_bytes;   // bytes of the ASCII stage (16-byte multiple, incl. slack)
    uint32_t warp_bytes;    // shared memory per warp
};
constexpr int HASH_WARPS = HASH_THREADS / 32;

template <bool FWD, bool RC>
__global__ void __launch_bounds__(HASH_THREADS, 6) hash_reads_kernel(HashWarpBatch b, uint64_t *__restrict__ fwd, uint64_t *__restrict__ rc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * b.warp_bytes;     // [mbarrier | ASCII stage | codes | valid | dirty flags]
    uint64_t *bar = (uint64_t *)wbase;
    uint8_t *ascii = wbase + 16;
    uint64_t *codes = (uint64_t *)(wbase + 16 + (size_t)b.stage_bytes);
    uint64_t *valid = codes + (size_t)b.rpw * b.words;
    uint32_t *dirty = (uint32_t *)(valid + (size_t)b.rpw * b.words);
    const uint64_t mask = kmer_mask(b.k);
    const uint32_t tile_bytes = (uint32_t)b.rpw * (uint32_t)b.read_len;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const int64_t total_warps = (int64_t)gridDim.x * HASH_WARPS;
    int64_t wt = (int64_t)blockIdx.x * HASH_WARPS + warp;
    auto uses_bulk = [&](int64_t tile) { return b.bulk_ok && (tile + 1) * (int64_t)b.rpw <= b.n_reads; };
    auto issue = [&](int64_t tile) {
        fence_proxy_async();
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(ascii, b.reads + tile * (int64_t)b.rpw * b.row_stride, tile_bytes, bar);
    };
    if (wt < b.n_wtiles && lane == 0 && uses_bulk(wt)) issue(wt);
    uint32_t phase = 0;
    const int halves = 2 * b.words;
    for (; wt < b.n_wtiles; wt += total_warps) {
        const int64_t next = wt + total_warps;
        const int64_t r0 = wt * (int64_t)b.rpw;
        const int n_here = (int)min((int64_t)b.rpw, b.n_reads - r0);
        if (uses_bulk(wt)) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {   // strided / unaligned / tail tile: the warp copies its rows itself
            for (int r = 0; r < n_here; r++) {
                const uint8_t *src = b.reads + (r0 + r) * b.row_stride;
                for (int i = lane; i < b.read_len; i += 32) ascii[(size_t)r * b.read_len + i] = __ldg(src + i);
            }
            __syncwarp();
        }
        // ---- pack to 2 bits per base (+ validity); one task = 16 bases = one 32-bit half of a code word (the pad word included) ----
        if (lane < n_here) dirty[lane] = 0;
        __syncwarp();
        for (int task = lane; task < n_here * halves; task += 32) {
            const int r = (int)(((uint32_t)task * b.inv_halves) >> 16), h = task - r * halves;   // task / halves (exact, checked on the host)
            const int first = h * 16;
            const int nb = min(16, b.read_len - first);
            uint32_t c32 = 0, v32 = 0;
            if (nb > 0) {
                const uint32_t addr = (uint32_t)r * (uint32_t)b.read_len + (uint32_t)first;
                const uint32_t *aligned = (const uint32_t *)(ascii + (addr & ~3u));
                const uint32_t sh = (addr & 3u) * 8u;
                uint32_t lo = aligned[0];
                const int nq = (nb + 3) >> 2;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (q < nq) {
                        const uint32_t hi = aligned[q + 1];
                        uint32_t c8, v8;
                        encode4(__funnelshift_r(lo, hi, sh), c8, v8);
                        lo = hi;
                        c32 |= c8 << (8 * q);
                        v32 |= v8 << (8 * q);
                    }
                }
                const uint32_t m = nb < 16 ? ((1u << (2 * nb)) - 1u) : ~0u;
                c32 &= m;
                v32 &= m;
                if (v32 != m) dirty[r] = 1;
            }
            ((uint32_t *)codes)[(size_t)r * halves + h] = c32;
            ((uint32_t *)valid)[(size_t)r * halves + h] = v32;
        }
        __syncwarp();
        if (next < b.n_wtiles && lane == 0 && uses_bulk(next)) issue(next);   // the stage is dead once packed: the next tile lands during the walk
        // ---- one read at a time ----
        for (int r = 0; r < n_here; r++) {
            const uint64_t *cw = codes + (size_t)r * b.words;
            uint64_t *of = FWD ? fwd + (r0 + r) * (int64_t)b.nk : nullptr;
            uint64_t *orc = RC ? rc + (r0 + r) * (int64_t)b.nk : nullptr;
            if (!dirty[r]) {
                emit_clean_read<FWD, RC>(cw, b.nk, b.k, mask, of, orc, lane);
                continue;
            }
            const uint64_t *vw = valid + (size_t)r * b.words;
            for (int i = lane; i < b.nk; i += 32) {
                const uint64_t x = extract_window(cw, i, mask);
                if (FWD) of[i] = x;
                if (RC) orc[b.nk - 1 - i] = revcomp_hash_masked(x, extract_window(vw, i, mask), b.k);
            }
        }
        __syncwarp();   // the next round's pack overwrites codes / valid / dirty
    }
}

// ---------------------------------------------------------------------------------------------------
// General path: the batch is seen as one byte stream; position p starts a k-mer iff p..p+k-1 lie in the
// same read.  Handles ragged reads (offsets != NULL) and arbitrarily long fixed-length rows.
constexpr int STREAM_THREADS = 256;
constexpr int STREAM_POS = 4096;  // k-mer start positions per tile

struct StreamBatch {
    const uint8_t *seq;
    const int64_t *offsets;      // n_reads+1 (ragged) or NULL (strided)
    const int64_t *out_offsets;  // n_reads+1 (ragged) or NULL
    int64_t n_reads;
    int64_t total_bytes;         // span of the byte stream
    int64_t row_stride;          // strided mode
    int32_t read_len;            // strided mode
    int32_t k;
};

__device__ __forceinline__ int64_t upper_read(const int64_t *offsets, int64_t lo, int64_t hi, int64_t p) {
    // largest r in [lo, hi] with offsets[r] <= p  (offsets[lo] <= p guaranteed)
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(STREAM_THREADS) hash_stream_kernel(StreamBatch b, uint64_t *__restrict__ fwd,
                                                                     uint64_t *__restrict__ rc) {
    constexpr int WORDS = (STREAM_POS + 31) / 32 + 2;   // covers STREAM_POS + k - 1 <= STREAM_POS + 30 bases + pad
    __shared__ uint64_t codes[WORDS];
    __shared__ uint64_t valid[WORDS];
    __shared__ int64_t rng[2];
    const uint64_t mask = kmer_mask(b.k);
    const int64_t n_tiles = (b.total_bytes + STREAM_POS - 1) / STREAM_POS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * STREAM_POS;
        // pack bases p0 .. p0 + STREAM_POS + 31 (zero beyond the stream)
        for (int w = threadIdx.x; w < WORDS; w += blockDim.x) {
            uint64_t cw = 0, vw = 0;
            int64_t base = p0 + (int64_t)w * 32;
            for (int q = 0; q < 32; q++) {
                int64_t p = base + q;
                if (p < b.total_bytes) {
                    uint32_t e = encode_base(__ldg(b.seq + p));
                    cw |= (uint64_t)(e & 3u) << (2 * q);
                    vw |= (uint64_t)((e & 4u) ? 3u : 0u) << (2 * q);
                }
            }
            codes[w] = cw;
            valid[w] = vw;
        }
        if (b.offsets && threadIdx.x == 0) {
            int64_t last = min(p0 + STREAM_POS - 1, b.total_bytes - 1);
            rng[0] = upper_read(b.offsets, 0, b.n_reads - 1, p0);
            rng[1] = upper_read(b.offsets, rng[0], b.n_reads - 1, last);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < STREAM_POS; i += blockDim.x) {
            int64_t p = p0 + i;
            if (p >= b.total_bytes) break;
            int64_t r, start, end, obase;
            if (b.offsets) {
                r = upper_read(b.offsets, rng[0], rng[1], p);
                start = __ldg(b.offsets + r);
                end = __ldg(b.offsets + r + 1);
                obase = __ldg(b.out_offsets + r);
            } else {
                r = p / b.row_stride;
                start = r * b.row_stride;
                end = start + b.read_len;
                obase = r * (int64_t)(b.read_len - b.k + 1);
            }
            if (p + b.k > end) continue;
            int64_t nk = end - start - b.k + 1;
            int64_t off = p - start;
            uint64_t x = extract_window(codes, i, mask);
            if (fwd) fwd[obase + off] = x;
            if (rc) {
                uint64_t v = extract_window(valid, i, mask);
                rc[obase + (nk - 1 - off)] = revcomp_hash_masked(x, v, b.k);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
__global__ void encode_bases_kernel(const uint8_t *__restrict__ seq, int64_t n, uint64_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = encode_base(__ldg(seq + i)) & 3u;
}

template <int MODE>  // 0 reverse complement, 1 complement
__global__ void transform_hashes_kernel(const uint64_t *__restrict__ in, int64_t n, int k, uint64_t *__restrict__ out) {
    const uint64_t mask = kmer_mask(k);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t x = __ldg(in + i) & mask;
        out[i] = (MODE == 0) ? revcomp_hash(x, k) : ((~x) & mask);
    }
}

__global__ void hashes_to_bases_kernel(const uint64_t *__restrict__ in, int64_t n, int k, uint64_t *__restrict__ out) {
    int64_t total = n * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / k;
        int j = (int)(i - r * k);
        out[i] = (__ldg(in + r) >> (2 * j)) & 3ull;
    }
}

static int check_k(int32_t k) {
    GKI_REQUIRE(k >= 1 && k <= 31, GKI_ERR_INVALID, "k must be in [1, 31] (reference asserts k <= 31, kmer_hashing.py:25), got %d", k);
    return GKI_OK;
}

int launch_hash_stream(const StreamBatch &b, uint64_t *fwd, uint64_t *rc, cudaStream_t s) {
    if (b.total_bytes <= 0) return GKI_OK;
    int64_t n_tiles = (b.total_bytes + STREAM_POS - 1) / STREAM_POS;
    int grid = grid_for(n_tiles, 1, device_info().sms * 8);
    hash_stream_kernel<<<grid, STREAM_THREADS, 0, s>>>(b, fwd, rc);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

}  // namespace gki

using namespace gki;

extern "C" {

int gki_encode_bases(const uint8_t *seq, int64_t n, uint64_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(n >= 0 && (n == 0 || (seq && out)), GKI_ERR_INVALID, "gki_encode_bases: bad arguments");
    if (n == 0) return GKI_OK;
    DevIn in;
    DevOut o;
    GKI_TRY(in.stage(seq, (size_t)n, call.stream));
    GKI_TRY(o.prepare(out, (size_t)n * 8, call.stream));
    encode_bases_kernel<<<grid_for(n, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(in.as<uint8_t>(), n, o.as<uint64_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_hash_reads(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, int32_t k, uint64_t *fwd,
                   uint64_t *rc, gki_stream_t stream) {
    CallScope call(stream);
    GKI_TRY(check_k(k));
    GKI_REQUIRE(n_reads >= 0 && read_len >= 0 && row_stride >= read_len, GKI_ERR_INVALID,
                "gki_hash_reads: need n_reads >= 0 and row_stride >= read_len >= 0");
    if (n_reads == 0 || read_len < k || (!fwd && !rc)) return GKI_OK;
    GKI_REQUIRE(reads, GKI_ERR_INVALID, "gki_hash_reads: reads is NULL");
    const int64_t nk = read_len - k + 1;
    const size_t in_bytes = (size_t)((n_reads - 1) * row_stride + read_len);
    DevIn in;
    DevOut of, orc;
    GKI_TRY(in.stage(reads, in_bytes, call.stream));
    GKI_TRY(of.prepare(fwd, (size_t)n_reads * nk * 8, call.stream));
    GKI_TRY(orc.prepare(rc, (size_t)n_reads * nk * 8, call.stream));
    HashWarpBatch b{};
    b.reads = in.as<uint8_t>();
    b.n_reads = n_reads;
    b.row_stride = row_stride;
    b.read_len = read_len;
    b.k = k;
    b.nk = (int32_t)nk;
    b.words = (read_len + 31) / 32 + 1;
    auto warp_bytes = [&](int rpw) {
        const size_t stage = (((size_t)rpw * read_len + 15) & ~(size_t)15) + 16;
        return (16 + stage + 2 * (size_t)rpw * b.words * 8 + (size_t)rpw * 4 + 15) & ~(size_t)15;
    };
    int rpw = 8;
    if (const char *e = experiment_knob("GKI_HASH_RPW")) rpw = atoi(e) < 1 ? 1 : (atoi(e) > 32 ? 32 : atoi(e));
    while (rpw > 1 && warp_bytes(rpw) * HASH_WARPS > 36 * 1024) rpw >>= 1;
    b.rpw = rpw;
    b.inv_halves = 65536u / (2u * (uint32_t)b.words) + 1u;
    bool warp_ok = warp_bytes(rpw) * HASH_WARPS <= 64 * 1024;
    for (uint32_t task = 0; warp_ok && task < (uint32_t)rpw * 2u * (uint32_t)b.words; task++)
        warp_ok = ((task * b.inv_halves) >> 16) == task / (2u * (uint32_t)b.words);
    if (warp_ok) {
        b.stage_bytes = (uint32_t)((((size_t)rpw * read_len + 15) & ~(size_t)15) + 16);
        b.warp_bytes = (uint32_t)warp_bytes(rpw);
        b.n_wtiles = (n_reads + rpw - 1) / rpw;
        b.bulk_ok = (row_stride == read_len) && (((uintptr_t)b.reads & 15) == 0) && (((int64_t)rpw * read_len) % 16 == 0);
        const size_t smem = (size_t)b.warp_bytes * HASH_WARPS;
        int ctas_per_sm = 4096;   // in effect one tile per warp; 6 CTAs are resident (40 registers), the rest of the grid arrives staggered (see the kernel)
        if (const char *e = experiment_knob("GKI_HASH_CTAS")) ctas_per_sm = atoi(e) > 0 ? atoi(e) : ctas_per_sm;
        const int grid = grid_for(b.n_wtiles, HASH_WARPS, device_info().sms * ctas_per_sm);
        if (fwd && rc) {
            GKI_CUDA(cudaFuncSetAttribute(hash_reads_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            hash_reads_kernel<true, true><<<grid, HASH_THREADS, smem, call.stream>>>(b, of.as<uint64_t>(), orc.as<uint64_t>());
        } else if (fwd) {
            GKI_CUDA(cudaFuncSetAttribute(hash_reads_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            hash_reads_kernel<true, false><<<grid, HASH_THREADS, smem, call.stream>>>(b, of.as<uint64_t>(), nullptr);
        } else {
            GKI_CUDA(cudaFuncSetAttribute(hash_reads_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            hash_reads_kernel<false, true><<<grid, HASH_THREADS, smem, call.stream>>>(b, nullptr, orc.as<uint64_t>());
        }
        GKI_CHECK_LAUNCH();
    } else {  // very long rows: byte-stream path
        StreamBatch sb{in.as<uint8_t>(), nullptr, nullptr, n_reads, (int64_t)in_bytes, row_stride, read_len, k};
        GKI_TRY(launch_hash_stream(sb, of.as<uint64_t>(), orc.as<uint64_t>(), call.stream));
    }
    GKI_TRY(of.finish(call.stream));
    GKI_TRY(orc.finish(call.stream));
    return call.finish();
}

int gki_hash_reads_ragged(const uint8_t *seq, const int64_t *offsets, const int64_t *out_offsets, int64_t n_reads,
                          int32_t k, uint64_t *fwd, uint64_t *rc, gki_stream_t stream) {
    CallScope call(stream);
    GKI_TRY(check_k(k));
    GKI_REQUIRE(n_reads >= 0, GKI_ERR_INVALID, "gki_hash_reads_ragged: n_reads < 0");
    if (n_reads == 0 || (!fwd && !rc)) return GKI_OK;
    GKI_REQUIRE(seq && offsets && out_offsets, GKI_ERR_INVALID, "gki_hash_reads_ragged: NULL argument");
    // the two totals are needed on the host to size the staging
    int64_t ends[2];
    if (is_device_ptr(offsets)) {
        GKI_CUDA(cudaMemcpyAsync(&ends[0], offsets + n_reads, 8, cudaMemcpyDeviceToHost, call.stream));
        GKI_CUDA(cudaMemcpyAsync(&ends[1], out_offsets + n_reads, 8, cudaMemcpyDeviceToHost, call.stream));
        GKI_CUDA(cudaStreamSynchronize(call.stream));
    } else {
        ends[0] = offsets[n_reads];
        ends[1] = out_offsets[n_reads];
    }
    DevIn in, off, ooff;
    DevOut of, orc;
    GKI_TRY(in.stage(seq, (size_t)ends[0], call.stream));
    GKI_TRY(off.stage(offsets, (size_t)(n_reads + 1) * 8, call.stream));
    GKI_TRY(ooff.stage(out_offsets, (size_t)(n_reads + 1) * 8, call.stream));
    GKI_TRY(of.prepare(fwd, (size_t)ends[1] * 8, call.stream));
    GKI_TRY(orc.prepare(rc, (size_t)ends[1] * 8, call.stream));
    StreamBatch sb{in.as<uint8_t>(), off.as<int64_t>(), ooff.as<int64_t>(), n_reads, ends[0], 0, 0, k};
    GKI_TRY(launch_hash_stream(sb, of.as<uint64_t>(), orc.as<uint64_t>(), call.stream));
    GKI_TRY(of.finish(call.stream));
    GKI_TRY(orc.finish(call.stream));
    return call.finish();
}

static int transform_hashes(int mode, const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_TRY(check_k(k));
    GKI_REQUIRE(n >= 0 && (n == 0 || (in && out)), GKI_ERR_INVALID, "hash transform: bad arguments");
    if (n == 0) return GKI_OK;
    DevIn i;
    DevOut o;
    GKI_TRY(i.stage(in, (size_t)n * 8, call.stream));
    GKI_TRY(o.prepare(out, (size_t)n * 8, call.stream));
    int grid = grid_for(n, 256 * 4, device_info().sms * 16);
    if (mode == 0) transform_hashes_kernel<0><<<grid, 256, 0, call.stream>>>(i.as<uint64_t>(), n, k, o.as<uint64_t>());
    else transform_hashes_kernel<1><<<grid, 256, 0, call.stream>>>(i.as<uint64_t>(), n, k, o.as<uint64_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_revcomp_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream) {
    return transform_hashes(0, in, n, k, out, stream);
}
int gki_complement_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream) {
    return transform_hashes(1, in, n, k, out, stream);
}

int gki_hashes_to_bases(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(k >= 1 && k <= 32, GKI_ERR_INVALID, "gki_hashes_to_bases: k must be in [1, 32]");
    GKI_REQUIRE(n >= 0 && (n == 0 || (in && out)), GKI_ERR_INVALID, "gki_hashes_to_bases: bad arguments");
    if (n == 0) return GKI_OK;
    DevIn i;
    DevOut o;
    GKI_TRY(i.stage(in, (size_t)n * 8, call.stream));
    GKI_TRY(o.prepare(out, (size_t)n * k * 8, call.stream));
    hashes_to_bases_kernel<<<grid_for(n * k, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(i.as<uint64_t>(), n, k, o.as<uint64_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

}  // extern "C"
