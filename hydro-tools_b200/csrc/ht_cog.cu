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
// Encoder: a tile's byte stream is cut into 4 KiB chunks; ONE THREAD deflates one chunk --
// greedy LZ77 (4-byte hash, matches inside the chunk) + the fixed Huffman code -- into a
// private slot, byte-aligned by an empty stored block (the "sync flush" pigz uses to join
// independently compressed pieces).  A warp stages its 32 chunks and hash tables in shared
// memory, lane-interleaved so that lanes walking their own chunk hit different banks.  A
// second kernel sizes the tiles, a third lays them out back to back with the zlib header
// and the Adler-32 of the tile (combined from per-chunk sums).  Rasters are HBM-resident
// row-major arrays of 1, 2, 4 or 8-byte samples; cells beyond the raster edge are padding.
#include "ht_common.cuh"

namespace {

constexpr int TILE = 512;
constexpr int CHUNK = 4096;                // input bytes per thread
constexpr int CW = CHUNK / 4;              // words
constexpr int HASH_BITS = 9, HASH_SLOTS = 1 << HASH_BITS;
constexpr int SLOT_WORDS = 1184;           // output slot: 4096 * 9 / 8 + header / flush, in words
constexpr int WARP_SMEM = 32 * (CHUNK + HASH_SLOTS * 2); // 147456 B

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

struct BitWriter {
    uint32_t *out;
    unsigned long long buf;
    int nbits, words;
    __device__ __forceinline__ void put(uint32_t v, int n) // n <= 31
    {
        buf |= (unsigned long long)v << nbits;
        nbits += n;
        if (nbits >= 32) {
            out[words++] = (uint32_t)buf;
            buf >>= 32;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void align_byte()
    {
        const int r = nbits & 7;
        if (r)
            put(0u, 8 - r);
    }
    __device__ __forceinline__ int finish() // bytes written; the last partial word is stored too
    {
        const int bytes = words * 4 + (nbits + 7) / 8;
        if (nbits)
            out[words] = (uint32_t)buf;
        return bytes;
    }
};

// fixed Huffman code of RFC 1951 3.2.6, already bit-reversed for the LSB-first stream
__device__ __forceinline__ void put_literal(BitWriter &w, uint32_t lit)
{
    if (lit < 144)
        w.put(rev_bits(0x30 + lit, 8), 8);
    else
        w.put(rev_bits(0x190 + (lit - 144), 9), 9);
}

__device__ __forceinline__ void put_match(BitWriter &w, int len, int dist)
{
    // length symbol 257..285 + extra bits
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
    if (sym < 280)
        w.put(rev_bits(sym - 256, 7), 7);
    else
        w.put(rev_bits(0xC0 + (sym - 280), 8), 8);
    if (eb)
        w.put((uint32_t)ev, eb);
    // distance symbol 0..29 (5 bits) + extra bits
    int dsym, deb = 0, dev = 0;
    if (dist <= 4)
        dsym = dist - 1;
    else {
        const int m = dist - 1, k = 31 - __clz(m);
        deb = k - 1;
        dsym = 2 * k + ((m >> (k - 1)) & 1);
        dev = m & ((1 << deb) - 1);
    }
    w.put(rev_bits(dsym, 5), 5);
    if (deb)
        w.put((uint32_t)dev, deb);
}

// one warp = 32 chunks; lane L's chunk word k lives at data[k * 32 + L]
__global__ void __launch_bounds__(32)
deflate_chunks_kernel(const EncArgs p, int64_t nchunks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *data = reinterpret_cast<uint32_t *>(smem_raw);                       // [CW][32]
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem_raw + 32 * CHUNK);            // [HASH_SLOTS][32]
    const int lane = threadIdx.x;
    const int row_bytes = TILE * p.es;
    const bool aligned = !((uintptr_t)p.src & 3) && !(p.ld_bytes & 3);
    for (int64_t c0 = (int64_t)blockIdx.x * 32; c0 < nchunks; c0 += (int64_t)gridDim.x * 32) {
        // ---- stage: the warp copies chunk after chunk, 128 B per lane-row, transposing into
        // the lane-interleaved layout
        for (int l = 0; l < 32; l++) {
            const int64_t c = c0 + l;
            if (c >= nchunks)
                break;
            const int64_t tile = c / p.chunks_per_tile;
            const int within = (int)(c - tile * p.chunks_per_tile);
            const int64_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            for (int k = lane; k < CW; k += 32) {
                const int byte0 = 4 * k;
                const int tr = within * p.rows_per_chunk + byte0 / row_bytes; // tile row
                const int off = byte0 - (byte0 / row_bytes) * row_bytes;      // byte inside it
                const int64_t r = ty * TILE + tr;
                const int64_t gbyte = tx * (int64_t)row_bytes + off; // byte column in the raster row
                uint32_t word = 0;
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
                data[k * 32 + l] = word;
            }
        }
        for (int k = 0; k < HASH_SLOTS; k++)
            tab[k * 32 + lane] = 0;
        __syncwarp();
        const int64_t c = c0 + lane;
        if (c < nchunks) {
            auto word_at = [&](int k) -> uint32_t { return data[k * 32 + lane]; };
            auto load32 = [&](int i) -> uint32_t { // 4 bytes at byte offset i (i + 4 <= CHUNK)
                const int k = i >> 2, s = (i & 3) * 8;
                const uint32_t lo = word_at(k), hi = s ? word_at(min(k + 1, CW - 1)) : 0u;
                return s ? __funnelshift_r(lo, hi, s) : lo;
            };
            auto byte_at = [&](int i) -> uint32_t { return (word_at(i >> 2) >> ((i & 3) * 8)) & 0xFFu; };
            // Adler-32 partial sums of the chunk
            unsigned long long s1 = 0, s2 = 0;
            for (int k = 0; k < CW; k++) {
                const uint32_t wv = word_at(k);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const unsigned v = (wv >> (8 * b)) & 0xFFu;
                    s1 += v;
                    s2 += (unsigned long long)(CHUNK - (4 * k + b)) * v;
                }
            }
            p.chunk_s1[c] = s1;
            p.chunk_s2[c] = s2;
            const bool last = (c % p.chunks_per_tile) == p.chunks_per_tile - 1;
            BitWriter w{p.slots + c * SLOT_WORDS, 0ull, 0, 0};
            w.put(last ? 1u : 0u, 1); // BFINAL
            w.put(1u, 2);             // BTYPE = 01: fixed Huffman
            int pos = 0;
            while (pos < CHUNK) {
                int len = 0, dist = 0;
                if (pos + 4 <= CHUNK) {
                    const uint32_t v = load32(pos);
                    const uint32_t h = (v * 2654435761u) >> (32 - HASH_BITS);
                    const int cand = (int)tab[h * 32 + lane] - 1;
                    tab[h * 32 + lane] = (uint16_t)(pos + 1);
                    if (cand >= 0 && load32(cand) == v) {
                        len = 4;
                        const int lim = min(258, CHUNK - pos);
                        while (len + 4 <= lim) {
                            const uint32_t x = load32(pos + len) ^ load32(cand + len);
                            if (x) {
                                len += (__ffs(x) - 1) >> 3;
                                break;
                            }
                            len += 4;
                        }
                        if (len + 4 > lim) // tail, byte by byte
                            while (len < lim && byte_at(pos + len) == byte_at(cand + len))
                                len++;
                        dist = pos - cand;
                    }
                }
                if (len >= 4) {
                    put_match(w, len, dist);
                    pos += len;
                }
                else {
                    put_literal(w, byte_at(pos));
                    pos++;
                }
            }
            w.put(0u, 7); // end of block (symbol 256)
            if (!last) {  // empty stored block: byte-aligns the stream for the next chunk
                w.put(0u, 3);
                w.align_byte();
                w.put(0x0000u, 16); // LEN
                w.put(0xFFFFu, 16); // NLEN
            }
            else
                w.align_byte();
            p.chunk_bytes[c] = (uint32_t)w.finish();
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
                                     WARP_SMEM));
    const int64_t warps = (nchunks + 31) / 32;
    const int grid = (int)std::min<int64_t>(warps, (int64_t)ht::kSMs * 16);
    deflate_chunks_kernel<<<grid, 32, WARP_SMEM, st>>>(p, nchunks);
    HT_LAUNCH_CHECK();
    tile_sizes_kernel<<<ht::cdiv(ntiles, 128), 128, 0, st>>>(p.chunk_bytes, p.chunks_per_tile, ntiles, tile_size);
    tile_offsets_kernel<<<1, 1024, 0, st>>>(tile_size, ntiles, tile_off, total);
    tile_assemble_kernel<<<(unsigned)ntiles, 256, 0, st>>>(p, tile_off, out);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
