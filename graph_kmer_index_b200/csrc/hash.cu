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

__global__ void __launch_bounds__(HASH_THREADS) hash_reads_kernel(ReadBatch b, uint64_t *__restrict__ fwd,
                                                                  uint64_t *__restrict__ rc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint64_t mask = kmer_mask(b.k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for_each_tile(b, smem_raw, [&](int64_t tile, const TileSmem &t) {
        int64_t r0 = tile * (int64_t)b.tile_reads;
        for (int r = warp; r < b.tile_reads && r0 + r < b.n_reads; r += nwarps) {
            const uint64_t *cw = t.codes + (size_t)r * b.words;
            const uint64_t *vw = t.valid + (size_t)r * b.words;
            uint64_t *of = fwd ? fwd + (r0 + r) * (int64_t)b.nk : nullptr;
            uint64_t *orc = rc ? rc + (r0 + r) * (int64_t)b.nk : nullptr;
            // a read made only of ACGTacgt (the usual case): every lane owns four consecutive windows, extracted once and
            // rolled three times (x >> 2 | next base on top; the reverse complement rolls the other way), and stores each
            // strand's four hashes with one 256-bit store
            bool ok = true;
            for (int w = lane; w < b.words - 1; w += 32) {
                const int nb = min(32, b.read_len - w * 32);
                ok &= vw[w] == (nb < 32 ? ((1ull << (2 * nb)) - 1ull) : ~0ull);
            }
            if (__all_sync(0xffffffffu, ok)) {
                for (int i0 = lane * 4; i0 < b.nk; i0 += 128) {
                    uint64_t x[4], y[4];
                    x[0] = extract_window(cw, i0, mask);
                    const uint32_t nxt = (uint32_t)extract_window(cw, i0 + b.k, 0x3Full);
                    y[0] = revcomp_hash(x[0], b.k);
#pragma unroll
                    for (int u = 1; u < 4; u++) {
                        const uint64_t nb = (nxt >> (2 * (u - 1))) & 3u;
                        x[u] = (x[u - 1] >> 2) | (nb << (2 * (b.k - 1)));
                        y[u] = ((y[u - 1] << 2) | (3u - nb)) & mask;
                    }
                    const int n_valid = min(4, b.nk - i0);
                    if (of) {
                        uint64_t *p = of + i0;
                        if (n_valid == 4 && ((uintptr_t)p & 31) == 0) st_global_v4_u64(p, x[0], x[1], x[2], x[3]);
                        else
                            for (int u = 0; u < n_valid; u++) p[u] = x[u];
                    }
                    if (orc) {   // window i of the read is window nk-1-i of the reverse-complemented read
                        uint64_t *q = orc + (b.nk - 4 - i0);
                        if (n_valid == 4 && ((uintptr_t)q & 31) == 0) st_global_v4_u64(q, y[3], y[2], y[1], y[0]);
                        else
                            for (int u = 0; u < n_valid; u++) orc[b.nk - 1 - i0 - u] = y[u];
                    }
                }
                continue;
            }
            for (int i = lane; i < b.nk; i += 32) {
                uint64_t x = extract_window(cw, i, mask);
                if (of) of[i] = x;
                if (orc) {
                    uint64_t v = extract_window(vw, i, mask);
                    // window i of the read is window nk-1-i of the reverse-complemented read
                    orc[b.nk - 1 - i] = revcomp_hash_masked(x, v, b.k);
                }
            }
        }
    });
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
    ReadBatch b;
    size_t smem;
    make_read_batch(in.as<uint8_t>(), n_reads, read_len, row_stride, k, b, smem);
    if (smem <= 64 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            GKI_CUDA(cudaFuncSetAttribute(hash_reads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            attr_set = true;
        }
        int ctas_per_sm = 8;   // measured: 4 -> 0.881 ms, 6 -> 0.844 ms, 8 -> 0.822 ms for 2 M x 150 bp reads (registers allow 8)
        if (const char *e = experiment_knob("GKI_HASH_CTAS")) ctas_per_sm = atoi(e) > 0 ? atoi(e) : 8;
        int grid = grid_for(b.n_tiles, 1, device_info().sms * ctas_per_sm);
        hash_reads_kernel<<<grid, HASH_THREADS, smem, call.stream>>>(b, of.as<uint64_t>(), orc.as<uint64_t>());
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
