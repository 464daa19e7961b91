// ht_cog.cu -- GeoTIFF / COG tile encoding on the device.
//
// The reference stores every result with hydrotools.raster.to_raster
// (/root/reference/src/hydrotools/raster.py:794-846): compute -> tiled 512 x 512 LZW GeoTIFF
// -> gdal_edit -stats -> gdal_translate -of COG (utils.py:78-98; config.py:23-45), i.e. three
// full-raster rewrites on the host, and before that the uncompressed result has to cross
// PCIe (13 B/cell for direction + accumulation -- what bounds the end-to-end step).  Here the
// 512 x 512 tiles are compressed where they were computed, and only the compressed bytes
// travel: ht_tiff_encode_tiles turns a device raster into one zlib (RFC 1950 / 1951) stream
// per tile, which is what TIFF compression 8 (Adobe deflate) stores; the host writes the
// BigTIFF / COG container around them (hydrotools/cuda/_cog.py).
//
// Encoder: a tile's byte stream is cut into 4 KiB chunks; ONE WARP deflates one chunk --
// greedy LZ77 (4-byte hash, matches inside the chunk) + the fixed Huffman code -- into a
// private slot, byte-aligned by an empty stored block (the "sync flush" pigz uses to join
// independently compressed pieces).  The lanes look at 32 consecutive positions per round
// (match search in parallel, greedy parse by shuffles, codes placed by a prefix sum of their
// bit lengths); chunk, hash table and output words live in shared memory, 16 warps per SM.  A
// second kernel sizes the tiles, a third lays them out back to back with the zlib header
// and the Adler-32 of the tile (combined from per-chunk sums).  Rasters are HBM-resident
// row-major arrays of 1, 2, 4 or 8-byte samples; cells beyond the raster edge are padding.
#include "ht_common.cuh"

namespace {

constexpr int TILE = 512;
constexpr int CHUNK = 4096;                // input bytes per thread
constexpr int CW = CHUNK / 4;              // words
constexpr int HASH_BITS = 11, HASH_SLOTS = 1 << HASH_BITS;
constexpr int SLOT_WORDS = 1184;           // output slot: 4096 * 9 / 8 + header / flush, in words

struct EncArgs {
    const uint8_t *src;
    int64_t ld_bytes;      // row pitch
    int64_t rows, cols;
    int es;                // sample bytes
    int tiles_x, tiles_y;
    int chunks_per_tile;   // TILE * TILE * es / CHUNK
    int rows_per_chunk;    // tile rows one chunk covers: CHUNK / (TILE * es)
    unsigned long long pad; // padding sample (little endian, es bytes)
    uint32_t *slots;       // [nchunks][SLOT_WORDS]
    uint32_t *chunk_bytes; // [nchunks]
    unsigned long long *chunk_s1, *chunk_s2; // Adler-32 partial sums
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// symbol of one parse position as (bits, nbits), nbits <= 31
__device__ __forceinline__ void literal_code(uint32_t lit, uint32_t &bits, int &nbits)
{
    if (lit < 144) {
        bits = rev_bits(0x30 + lit, 8);
        nbits = 8;
    }
    else {
        bits = rev_bits(0x190 + (lit - 144), 9);
        nbits = 9;
    }
}

__device__ __forceinline__ void match_code(int len, int dist, uint32_t &bits, int &nbits)
{
    int sym, eb = 0, ev = 0;
    if (len == 258)
        sym = 285;
    else {
        const int l = len - 3;
        if (l < 8)
            sym = 257 + l;
        else {
            const int k = 31 - __clz(l);
            eb = k - 2;
            sym = 257 + 4 * eb + 4 + ((l >> eb) & 3);
            ev = l & ((1 << eb) - 1);
        }
    }
    if (sym < 280) {
        bits = rev_bits(sym - 256, 7);
        nbits = 7;
    }
    else {
        bits = rev_bits(0xC0 + (sym - 280), 8);
        nbits = 8;
    }
    bits |= (uint32_t)ev << nbits;
    nbits += eb;
    int dsym, deb = 0, dev = 0;
    if (dist <= 4)
        dsym = dist - 1;
    else {
        const int m = dist - 1, k = 31 - __clz(m);
        deb = k - 1;
        dsym = 2 * k + ((m >> (k - 1)) & 1);
        dev = m & ((1 << deb) - 1);
    }
    bits |= rev_bits(dsym, 5) << nbits;
    nbits += 5;
    bits |= (uint32_t)dev << nbits;
    nbits += deb;
}

// ONE WARP deflates one chunk.  Per round the 32 lanes look at 32 consecutive positions:
// hash lookup + match extension in parallel, a greedy parse of the round by shuffles (the
// lane a match or literal ends on names the next parse position), then the selected lanes
// write their codes at the bit offsets a warp prefix sum gives them (atomicOr into the
// chunk's output words in shared memory) -- no lane ever walks the chunk alone.
constexpr int WARPS = 4;
constexpr int SM_DATA = CHUNK + 32;                    // chunk + zero padding for 4-byte reads
constexpr int SM_TAB = HASH_SLOTS * 2;
constexpr int SM_OUT = SLOT_WORDS * 4;
constexpr int SM_WARP = SM_DATA + SM_TAB + SM_OUT;     // 12 992 B

__global__ void __launch_bounds__(32 * WARPS)
deflate_chunks_kernel(const EncArgs p, int64_t nchunks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + (size_t)warp * SM_WARP;
    uint32_t *dataw = reinterpret_cast<uint32_t *>(base);
    const uint8_t *data = base;
    uint16_t *tab = reinterpret_cast<uint16_t *>(base + SM_DATA);
    uint32_t *outw = reinterpret_cast<uint32_t *>(base + SM_DATA + SM_TAB);
    const int row_bytes = TILE * p.es;
    const bool aligned = !((uintptr_t)p.src & 3) && !(p.ld_bytes & 3);
    const unsigned full = 0xffffffffu;
    auto load32 = [&](int i) -> uint32_t {
        const int k = i >> 2, sh = (i & 3) * 8;
        const uint32_t lo = dataw[k];
        return sh ? __funnelshift_r(lo, dataw[k + 1], sh) : lo;
    };
    for (int64_t c = (int64_t)blockIdx.x * WARPS + warp; c < nchunks; c += (int64_t)gridDim.x * WARPS) {
        // ---- stage the chunk (coalesced), clear table and output words
        const int64_t tile = c / p.chunks_per_tile;
        const int within = (int)(c - tile * p.chunks_per_tile);
        const int64_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        unsigned long long s1 = 0, s2 = 0; // Adler-32 partial sums
        for (int k = lane; k < CW + 8; k += 32) {
            uint32_t word = 0;
            if (k < CW) {
                const int byte0 = 4 * k;
                const int tr = within * p.rows_per_chunk + byte0 / row_bytes; // tile row
                const int off = byte0 - (byte0 / row_bytes) * row_bytes;      // byte inside it
                const int64_t r = ty * TILE + tr;
                const int64_t gbyte = tx * (int64_t)row_bytes + off; // byte column in the raster row
                if (aligned && r < p.rows && gbyte + 4 <= p.cols * p.es)
                    word = *reinterpret_cast<const uint32_t *>(p.src + r * p.ld_bytes + gbyte);
                else {
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int64_t col = (gbyte + b) / p.es;
                        const int sb = (int)((gbyte + b) % p.es);
                        const uint32_t v = (r < p.rows && col < p.cols)
                                               ? p.src[r * p.ld_bytes + col * p.es + sb]
                                               : (uint32_t)((p.pad >> (8 * sb)) & 0xFFu);
                        word |= v << (8 * b);
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const unsigned v = (word >> (8 * b)) & 0xFFu;
                    s1 += v;
                    s2 += (unsigned long long)(CHUNK - (byte0 + b)) * v;
                }
            }
            dataw[k] = word;
        }
        for (int k = lane; k < HASH_SLOTS; k += 32)
            tab[k] = 0xFFFF;
        for (int k = lane; k < SLOT_WORDS; k += 32)
            outw[k] = 0u;
        __syncwarp();
        const bool last = within == p.chunks_per_tile - 1;
        // block header: BFINAL, BTYPE = 01 (fixed Huffman)
        if (lane == 0)
            outw[0] = (last ? 1u : 0u) | (1u << 1);
        int bitpos = 3;
        int next_free = 0; // first position not covered by a match yet
        __syncwarp();
        for (int rbase = 0; rbase < CHUNK; rbase += 32) {
            const int pos = rbase + lane;
            const uint32_t v = load32(pos);
            const bool can = pos + 4 <= CHUNK;
            const uint32_t h = (v * 2654435761u) >> (32 - HASH_BITS);
            if (next_free >= rbase + 32) {
                // the whole round lies inside a match already emitted: only remember the positions
                if (can)
                    tab[h] = (uint16_t)pos;
                __syncwarp();
                continue;
            }
            int len = 0, dist = 0;
            int cand = -1;
            if (can) {
                const int t = tab[h];
                if (t != 0xFFFF && load32(t) == v)
                    cand = t;
                else if (pos >= p.es && load32(pos - p.es) == v) // the previous sample: runs, periods
                    cand = pos - p.es;
                else if (pos >= 1 && load32(pos - 1) == v)
                    cand = pos - 1;
            }
            __syncwarp();
            if (can)
                tab[h] = (uint16_t)pos; // after everybody has read: candidates come from earlier rounds
            const int lim = min(258, CHUNK - pos);
            if (cand >= 0) {
                // enough to know whether the match leaves the round; the one match that does and
                // is selected is extended to its full length below
                len = 4;
                const int cap = min(lim, 36);
                while (len < cap) {
                    const uint32_t x = load32(pos + len) ^ load32(cand + len);
                    if (x) {
                        len += (__ffs(x) - 1) >> 3;
                        break;
                    }
                    len += 4;
                }
                len = min(len, cap);
                dist = pos - cand;
            }
            // ---- greedy parse of the round: where does each position hand over?
            int cur = max(next_free - rbase, 0);
            unsigned sel;
            if (__ballot_sync(full, len >= 4) == 0u) { // literals only
                sel = full << cur;
                cur = 32;
            }
            else {
                const int jump = lane + (len >= 4 ? len : 1); // round-relative; >= 32 leaves the round
                sel = 0u;
                int lastsel = 0;
                while (cur < 32) {
                    sel |= 1u << cur;
                    lastsel = cur;
                    cur = __shfl_sync(full, jump, cur);
                }
                // the last selected position may hold a match cut at the cap: extend it
                if (lane == lastsel && len >= 36 && len < lim) {
                    while (len < lim) {
                        const uint32_t x = load32(pos + len) ^ load32(cand + len);
                        if (x) {
                            len += (__ffs(x) - 1) >> 3;
                            break;
                        }
                        len += 4;
                    }
                    len = min(len, lim);
                }
                cur = __shfl_sync(full, lane + (len >= 4 ? len : 1), lastsel);
            }
            next_free = rbase + cur;
            // ---- codes of the selected positions at their bit offsets
            const bool mine = (sel >> lane) & 1u;
            uint32_t bits = 0;
            int nbits = 0;
            if (mine) {
                if (len >= 4)
                    match_code(len, dist, bits, nbits);
                else
                    literal_code(v & 0xFFu, bits, nbits);
            }
            int inc = nbits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(full, inc, d);
                if (lane >= d)
                    inc += t;
            }
            const int total = __shfl_sync(full, inc, 31);
            if (mine) {
                const int at = bitpos + inc - nbits;
                const int wd = at >> 5, sh = at & 31;
                atomicOr(&outw[wd], bits << sh);
                if (sh + nbits > 32)
                    atomicOr(&outw[wd + 1], bits >> (32 - sh));
            }
            bitpos += total;
            __syncwarp();
        }
        // ---- end of block; a non-final chunk ends on an empty stored block (byte alignment)
        if (lane == 0) {
            bitpos += 7; // symbol 256 = seven zero bits
            if (!last) {
                bitpos += 3;                       // BFINAL = 0, BTYPE = 00
                bitpos = (bitpos + 7) & ~7;        // stored blocks start on a byte
                const int at = bitpos + 16;        // LEN = 0x0000, then NLEN = 0xFFFF
                const int wd = at >> 5, sh = at & 31;
                outw[wd] |= 0xFFFFu << sh;
                if (sh + 16 > 32)
                    outw[wd + 1] |= 0xFFFFu >> (32 - sh);
                bitpos += 32;
            }
            else
                bitpos = (bitpos + 7) & ~7;
            p.chunk_bytes[c] = (uint32_t)(bitpos >> 3);
        }
        bitpos = __shfl_sync(full, bitpos, 0);
        __syncwarp();
        uint32_t *slot = p.slots + c * SLOT_WORDS;
        for (int k = lane; k < (bitpos + 31) >> 5; k += 32)
            slot[k] = outw[k];
        // Adler-32 sums of the chunk
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            s1 += __shfl_down_sync(full, s1, d);
            s2 += __shfl_down_sync(full, s2, d);
        }
        if (lane == 0) {
            p.chunk_s1[c] = s1;
            p.chunk_s2[c] = s2;
        }
        __syncwarp();
    }
}

// bytes of each tile's zlib stream: 2 (header) + chunks + 4 (Adler-32)
__global__ void tile_sizes_kernel(const uint32_t *chunk_bytes, int chunks_per_tile, int64_t ntiles,
                                  unsigned long long *tile_size)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ntiles;
         t += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long s = 6;
        for (int k = 0; k < chunks_per_tile; k++)
            s += chunk_bytes[t * chunks_per_tile + k];
        tile_size[t] = s;
    }
}

// exclusive scan of the tile sizes (one block; a raster has at most a few hundred thousand tiles)
__global__ void tile_offsets_kernel(const unsigned long long *tile_size, int64_t ntiles,
                                    unsigned long long *tile_off, unsigned long long *total)
{
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const unsigned long long own = i < ntiles ? tile_size[i] : 0ull;
        part[threadIdx.x] = own;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const unsigned long long t = threadIdx.x >= d ? part[threadIdx.x - d] : 0ull;
            __syncthreads();
            part[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles)
            tile_off[i] = carry + part[threadIdx.x] - own;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = carry;
}

// one block per tile: zlib header, the chunk streams back to back, Adler-32 (big endian)
__global__ void __launch_bounds__(256)
tile_assemble_kernel(const EncArgs p, const unsigned long long *tile_off, uint8_t *out)
{
    __shared__ unsigned long long off_s;
    const int64_t t = blockIdx.x;
    uint8_t *dst = out + tile_off[t];
    if (threadIdx.x == 0) {
        dst[0] = 0x78; // deflate, 32 KiB window
        dst[1] = 0x01; // no preset dictionary, fastest; (0x7801 % 31 == 0)
        off_s = 2;
    }
    __syncthreads();
    for (int k = 0; k < p.chunks_per_tile; k++) {
        const int64_t c = t * p.chunks_per_tile + k;
        const uint32_t n = p.chunk_bytes[c];
        const uint8_t *srcb = reinterpret_cast<const uint8_t *>(p.slots + c * SLOT_WORDS);
        const unsigned long long at = off_s;
        for (uint32_t i = threadIdx.x; i < n; i += 256)
            dst[at + i] = srcb[i];
        __syncthreads();
        if (threadIdx.x == 0)
            off_s = at + n;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long a = 1, b = 0;
        for (int k = 0; k < p.chunks_per_tile; k++) {
            const int64_t c = t * p.chunks_per_tile + k;
            b = (b + (unsigned long long)CHUNK * a + p.chunk_s2[c]) % 65521ull;
            a = (a + p.chunk_s1[c]) % 65521ull;
        }
        const uint32_t adler = (uint32_t)((b << 16) | a);
        uint8_t *e = dst + off_s;
        e[0] = (uint8_t)(adler >> 24); e[1] = (uint8_t)(adler >> 16);
        e[2] = (uint8_t)(adler >> 8); e[3] = (uint8_t)adler;
    }
}

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

} // namespace

// worst-case bytes of the encoded raster (`out` must hold that much)
extern "C" size_t ht_tiff_encode_bound(int64_t rows, int64_t cols, int elem_bytes)
{
    if (rows <= 0 || cols <= 0 || elem_bytes <= 0)
        return 0;
    const size_t ntiles = (size_t)ht::cdiv(rows, TILE) * ht::cdiv(cols, TILE);
    const size_t cpt = (size_t)TILE * TILE * elem_bytes / CHUNK;
    return ntiles * (cpt * SLOT_WORDS * 4 + 8);
}

extern "C" size_t ht_tiff_encode_workspace_bytes(int64_t rows, int64_t cols, int elem_bytes)
{
    if (rows <= 0 || cols <= 0 || elem_bytes <= 0)
        return 256;
    const size_t ntiles = (size_t)ht::cdiv(rows, TILE) * ht::cdiv(cols, TILE);
    const size_t nchunks = ntiles * ((size_t)TILE * TILE * elem_bytes / CHUNK);
    return align_up(nchunks * SLOT_WORDS * 4) + align_up(nchunks * 4) + 2 * align_up(nchunks * 8) + 1024;
}

// src: device raster, row pitch ld_bytes, samples of elem_bytes (1, 2, 4, 8).  pad_value: the
// sample (little endian in the low bytes) stored beyond the raster edge inside edge tiles.
// out: device buffer of ht_tiff_encode_bound bytes; tile t (row-major over the tile grid)
// occupies out[tile_off[t] .. + tile_size[t]); *total = bytes used.  All counters are device
// memory; nothing synchronises.
extern "C" int ht_tiff_encode_tiles(const void *src, int64_t ld_bytes, int64_t rows, int64_t cols,
                                    int elem_bytes, unsigned long long pad_value, uint8_t *out,
                                    size_t out_cap, unsigned long long *tile_off,
                                    unsigned long long *tile_size, unsigned long long *total, void *ws,
                                    size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!src || !out || !tile_off || !tile_size || !total || rows <= 0 || cols <= 0 ||
        (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8) ||
        ld_bytes < cols * elem_bytes)
        return HT_ERR_ARG;
    if (out_cap < ht_tiff_encode_bound(rows, cols, elem_bytes))
        return HT_ERR_ARG;
    if (!ws || ws_bytes < ht_tiff_encode_workspace_bytes(rows, cols, elem_bytes))
        return HT_ERR_WORKSPACE;
    EncArgs p;
    p.src = (const uint8_t *)src; p.ld_bytes = ld_bytes; p.rows = rows; p.cols = cols; p.es = elem_bytes;
    p.tiles_x = ht::cdiv(cols, TILE); p.tiles_y = ht::cdiv(rows, TILE);
    p.chunks_per_tile = TILE * TILE * elem_bytes / CHUNK;
    p.rows_per_chunk = CHUNK / (TILE * elem_bytes);
    p.pad = pad_value;
    const int64_t ntiles = (int64_t)p.tiles_x * p.tiles_y, nchunks = ntiles * p.chunks_per_tile;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    p.slots = (uint32_t *)b; b += align_up((size_t)nchunks * SLOT_WORDS * 4);
    p.chunk_bytes = (uint32_t *)b; b += align_up((size_t)nchunks * 4);
    p.chunk_s1 = (unsigned long long *)b; b += align_up((size_t)nchunks * 8);
    p.chunk_s2 = (unsigned long long *)b;
    HT_CUDA_TRY(cudaFuncSetAttribute(deflate_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     WARPS * SM_WARP));
    const int grid = (int)std::min<int64_t>((nchunks + WARPS - 1) / WARPS, (int64_t)ht::kSMs * 4 * 8);
    deflate_chunks_kernel<<<grid, 32 * WARPS, WARPS * SM_WARP, st>>>(p, nchunks);
    HT_LAUNCH_CHECK();
    tile_sizes_kernel<<<ht::cdiv(ntiles, 128), 128, 0, st>>>(p.chunk_bytes, p.chunks_per_tile, ntiles, tile_size);
    tile_offsets_kernel<<<1, 1024, 0, st>>>(tile_size, ntiles, tile_off, total);
    tile_assemble_kernel<<<(unsigned)ntiles, 256, 0, st>>>(p, tile_off, out);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
