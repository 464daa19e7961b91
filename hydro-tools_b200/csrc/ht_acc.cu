// ht_acc.cu -- flow accumulation over a D8 direction grid.
//
// Replaces the accumulation half of GRASS `r.watershed -s [-a]`
// (/root/reference/src/hydrotools/watershed.py:64-75, accumulation="fa"; rule R5
// of oracle/rwatershed.c) and GRASS `r.flowaccumulation [weight=]`
// (/root/reference/src/hydrotools/watershed.py:341-382; oracle/flowacc.c):
//     acc(c) = w(c) + sum of acc(u) over cells u whose receiver is c.
//
// Three stages, all on the device (Barnes-style tile decomposition):
//  A  acc_local_cnt_kernel / acc_f64_kernel<false>: per 64x64 tile held in shared
//     memory, subtree sums over the tile's forest by pointer doubling.  Flow that
//     leaves the tile is added to the boundary graph (`base`), and every entry
//     cell records where its path leaves the tile (`next`).
//  B  coarse_resolve_kernel: the boundary graph (one slot per tile-perimeter
//     cell) is a forest; subtree sums by pointer doubling, iterated inside one
//     cooperative launch until no node is active.
//  C  acc_final_cnt_kernel (counts: tile-local counts + the resolved inflow walked
//     along the entry paths) / acc_f64_kernel<true> (weights: stage A again with
//     the inflow arriving at the entry cells); writes the accumulation (fp64,
//     exact integers for counts) and the final GRASS direction codes.
// Counts are carried as u64 across tiles; weighted sums as fp64.
// Algorithmic traffic: 1 B/cell read + 8 B/cell written (+4 B/cell weights).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "ht_common.cuh"


namespace {

constexpr int T = 64;            // tile edge (cells)
constexpr int RH = T + 2;        // ring tile rows
constexpr int XO = 16;           // ring tile column of tile column 0: TMA needs the box
                                 // to start on a 16-byte boundary of the global row
constexpr int RP = XO + T + 16;  // ring tile pitch = TMA box width in bytes (96)
constexpr int THREADS = 256;
constexpr int CPT = T * T / THREADS; // cells per thread
constexpr int SLOTS = 4 * T;     // boundary-graph slots per tile
constexpr int32_t TERMINAL = -1;

// the ring tile (TMA destination) comes first in every kernel's dynamic shared memory
constexpr int SM_EFF = (RH * RP + 127) / 128 * 128;

enum { W_ONE = 0, W_TAINT = 1, W_F32 = 2 };

struct AccArgs {
    int64_t rows, cols;       // owned rows, columns
    int64_t row0, total_rows;
    int halo_top, halo_bot;   // 0/1: packed row -1 / row `rows` is a neighbour band's
    int tiles_x, tiles_y;
    void *base;               // V[nslots]   boundary-graph node weights (stage A out)
    int32_t *next;            // [nslots]    boundary-graph successor      (stage A out)
    const void *inflow;       // V[nslots]   resolved inflow               (stage C in)
    const float *w;           // weights (W_F32)
    int64_t w_ld;
    int has_wnd;
    float wnd;
    double *acc;
    int64_t acc_ld;
    int8_t *dir_out;
    int64_t dir_ld;
    int32_t *list;            // slots that are boundary-graph nodes (stage A out)
    int *nlist;               // their number (device counter)
    uint16_t *lacc;           // tile-local counts, stage A -> stage C (count modes)
    int64_t lacc_ld;
    const uint8_t *packed;    // owned row 0 of the packed grid (rare global look-ups)
    int64_t packed_ld;
    long long *lacc64;        // weighted mode: tile-local sums in the tile's unit, stage A -> C
    int64_t lacc64_ld;
    int32_t *tile_e;          // weighted mode: log2(tile unit / global unit) per tile
    const unsigned *gmax_bits; // weighted mode: bits of the raster's largest |weight| (float)
    int32_t *tile_first;      // count modes: where the tile's entries start in `list`, and how
    int32_t *tile_cnt;        // many they are (stage A out, for the two-level stage B)
    const uint8_t *tile_phase; // stage C of a sharded run: 1 = the tile's inflow depends on the
    int phase;                 // neighbour bands' deliveries; only tiles of `phase` are processed
};

// row / column step of GRASS code 1..8 (2-bit tables: dr+1 = 0,0,0,1,2,2,2,1; dc+1 = 2,1,0,0,0,1,2,2)
__device__ __forceinline__ int code_dr(int code) { return (int)((0x1A900u >> (2 * code)) & 3u) - 1; }
__device__ __forceinline__ int code_dc(int code) { return (int)((0x29018u >> (2 * code)) & 3u) - 1; }
// GRASS code that points at window cell k (3x3 row-major)
__device__ __forceinline__ int win2code(int k) { return (int)((0x765804123ULL >> (4 * k)) & 0xF); }
// r.watershed neighbour order S,N,W,E,NE,SW,SE,NW as window cells
__device__ __forceinline__ int ct2win(int ct) { return (0x08625317u >> (4 * ct)) & 0xF; }

// perimeter slot k of a th x tw tile -> (ly, lx); false for duplicates / unused
__device__ __forceinline__ bool perimeter_cell(int k, int th, int tw, int &ly, int &lx)
{
    if (k < T) { ly = 0; lx = k; }
    else if (k < 2 * T) { ly = th - 1; lx = k - T; if (th == 1) return false; }
    else if (k < 3 * T) { ly = k - 2 * T; lx = 0; if (ly == 0 || ly >= th - 1) return false; }
    else { ly = k - 3 * T; lx = tw - 1; if (ly == 0 || ly >= th - 1 || tw == 1) return false; }
    return ly < th && lx < tw;
}

static_assert(SLOTS == THREADS, "one perimeter slot per thread");

// 128-bit two's complement integer: weighted sums are carried as exact fixed point
// (order-independent, hence bit-reproducible; see "Weighted mode" below).
struct i128 {
    unsigned long long lo;
    long long hi;
    __host__ __device__ i128() {}
    __host__ __device__ i128(int v) : lo((unsigned long long)(long long)v), hi(v < 0 ? -1ll : 0ll) {}
    __host__ __device__ i128(unsigned long long l, long long h) : lo(l), hi(h) {}
};
__host__ __device__ __forceinline__ i128 operator+(const i128 &a, const i128 &b)
{
    i128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ll : 0ll);
    return r;
}
__host__ __device__ __forceinline__ i128 &operator+=(i128 &a, const i128 &b) { a = a + b; return a; }
__host__ __device__ __forceinline__ bool operator!=(const i128 &a, const i128 &b) { return a.lo != b.lo || a.hi != b.hi; }
// v << e, 0 <= e < 64, v sign-extended to 128 bits
__device__ __forceinline__ i128 shl_i64(long long v, int e)
{
    i128 r;
    r.lo = (unsigned long long)v << e;
    r.hi = e ? (v >> (64 - e)) : (v < 0 ? -1ll : 0ll);
    return r;
}
__device__ __forceinline__ double i128_to_double(const i128 &v)
{
    // hi * 2^64 + lo: both terms exact or correctly rounded, one more rounding in the sum
    return (double)v.hi * 18446744073709551616.0 + (double)v.lo;
}

template <typename V> __device__ __forceinline__ void atomic_add_v(V *p, V v);
template <> __device__ __forceinline__ void atomic_add_v<unsigned long long>(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_v<double>(double *p, double v) { atomicAdd(p, v); }
// two native 64-bit atomics with the carry handed on: the final 128-bit value does not
// depend on the order of the adds (owners read it only after a grid barrier / kernel end)
template <> __device__ __forceinline__ void atomic_add_v<i128>(i128 *p, i128 v)
{
    unsigned long long *w = reinterpret_cast<unsigned long long *>(p);
    unsigned long long carry = 0ull;
    if (v.lo) {
        const unsigned long long old = atomicAdd(w, v.lo);
        carry = (old + v.lo) < old ? 1ull : 0ull;
    }
    const unsigned long long h = (unsigned long long)v.hi + carry;
    if (h)
        atomicAdd(w + 1, h);
}

// ===========================================================================
// Count modes (cell counts, taint counts): lean integer kernels.
//   stage A  acc_local_cnt_kernel : subtree sums over the tile's forest by
//            POINTER DOUBLING.  Per cell a 16-bit pointer entry in shared memory,
//                [12:1] the cell 2^k steps downstream (active) or the root of the
//                cell's path (done; a root points at itself) | [0] active,
//            and, in registers of the owning thread, the sum over the 2^k cells
//            upstream.  Round k: every active cell hands its sum to the cell 2^k
//            steps downstream (one shared atomic into `arr`) and doubles its pointer;
//            ~log2(longest path) rounds, every lane busy in every round -- a
//            river trunk is never walked cell by cell.  The rounds leave, for
//            every cell, its tile-local count (u16 -> stage C) and the cell its
//            path leaves the tile through (-> boundary graph links).
//   stage C  acc_final_cnt_kernel : final = tile-local count + inflow of every
//            entry cell upstream; each entry's inflow is added along its path to
//            the tile exit by an independent walker (no ordering needed).
// Producers of the packed grid guarantee that a flowing cell (code 1..8, no
// NEG / EDGE / NULL flag) points at a valid cell (ht_d8.cu by construction,
// ht_dir_pack_i8 by flagging the others EDGE), so no target is re-checked here.
// ===========================================================================
constexpr unsigned FLAGS_STOP = HT_DIR_NEG | HT_DIR_EDGE | HT_DIR_NULL;
constexpr unsigned J_PTR = 0x1FFEu, J_ACTIVE = 1u;
constexpr int MAX_ROUNDS = 13; // 2^13 > T*T: only a cyclic (invalid) direction grid gets here

constexpr int SMC_PTR = SM_EFF;                  // uint32[T*T] pointer entries
constexpr int SMC_ARR = SMC_PTR + T * T * 4;     // uint32[T*T] own value + sums that arrived at a cell
constexpr int SMC_LUT = SMC_ARR + T * T * 4;     // uint32[2][16] code -> pointer step | active (as stored, transposed)
constexpr int SMC_BAR = SMC_LUT + 128;           // mbarrier, then counters
constexpr int SMC_TOTAL = SMC_BAR + 32;

// offset (dr * T + dc) of GRASS code 1..8 inside the T x T tile arrays
__host__ __device__ constexpr int link_offset_of(int code)
{
    return code == 1 ? -T + 1 : code == 2 ? -T : code == 3 ? -T - 1 : code == 4 ? -1
         : code == 5 ? T - 1 : code == 6 ? T : code == 7 ? T + 1 : code == 8 ? 1 : 0;
}
__host__ __device__ constexpr int link_offset_pitch(int code, int pitch)
{
    return code == 1 ? -pitch + 1 : code == 2 ? -pitch : code == 3 ? -pitch - 1 : code == 4 ? -1
         : code == 5 ? pitch - 1 : code == 6 ? pitch : code == 7 ? pitch + 1 : code == 8 ? 1 : 0;
}
// codes that leave the tile through its west / east / north / south side (bit = code)
constexpr unsigned EX_W = (1u << 3) | (1u << 4) | (1u << 5), EX_E = (1u << 1) | (1u << 7) | (1u << 8);
constexpr unsigned EX_N = (1u << 1) | (1u << 2) | (1u << 3), EX_S = (1u << 5) | (1u << 6) | (1u << 7);

// boundary-graph slot of the cell that (ly, lx) of tile (tx, ty) drains into
// through `code`; the target lies outside the tile.  32-bit throughout
// (fill_args guarantees slots < 2^31).
__device__ __forceinline__ int32_t exit_slot(const AccArgs &p, int tx, int ty, int ly, int lx, int code)
{
    const int lr = ty * T + ly + code_dr(code), gc = tx * T + lx + code_dc(code);
    const int rows = (int)p.rows;
    const int nt = p.tiles_x * p.tiles_y * SLOTS;
    if (lr < 0)
        return nt + gc;
    if (lr >= rows)
        return nt + (int)p.cols + gc;
    const int ty2 = lr >> 6, tx2 = gc >> 6, ly2 = lr & (T - 1), lx2 = gc & (T - 1);
    const int h2 = min(T, rows - ty2 * T);
    const int k2 = ly2 == 0 ? lx2 : ly2 == h2 - 1 ? T + lx2 : lx2 == 0 ? 2 * T + ly2 : 3 * T + ly2;
    return (ty2 * p.tiles_x + tx2) * SLOTS + k2;
}

// does the (valid, flowing) perimeter cell (ly, lx) drain out of a th x tw tile?
__device__ __forceinline__ bool leaves_tile(unsigned b, int ly, int lx, int th, int tw)
{
    const unsigned ex = (lx == 0 ? EX_W : 0u) | (lx == tw - 1 ? EX_E : 0u) |
                        (ly == 0 ? EX_N : 0u) | (ly == th - 1 ? EX_S : 0u);
    return !(b & FLAGS_STOP) && ((ex >> (b & HT_DIR_CODE)) & 1u);
}

// Orientation.  Lanes of a warp own 32 consecutive cells of one row of the tile arrays;
// their hand-offs of round k land 2^k steps downstream.  Flow lines that run ACROSS the
// warp's row coalesce (neighbouring lines merge at every confluence), so after a few rounds
// most lanes of the warp target the same trunk cells and the shared atomics serialise
// (6 wavefronts per ATOMS on the benchmark DEM, whose rivers run north-south; 2 when the
// warp lies along the flow); cells ALONG one flow line hand on to distinct cells in every
// round.  So every tile first votes on its dominant flow axis (two-step displacements of
// sampled cells) and, when it is the row axis, keeps its arrays TRANSPOSED: array index =
// lx * T + ly.  Pointers are array addresses, so the rounds do not know about it.
__device__ __forceinline__ int arr_index(bool tr, int ly, int lx) { return tr ? lx * T + ly : ly * T + lx; }

// The rounds are written branch-free in PTX on 32-bit shared addresses: a pointer entry IS
// the shared address of the target's entry (| active bit), its arrival word sits CNT_DELTA
// bytes behind.  No generic-address arithmetic, no divergent regions: the 16 hand-offs of a
// thread issue back to back and their latencies overlap.
constexpr int CNT_DELTA = T * T * 4;
// round, first half: active -> add `a` to the arrivals of the target, fetch the target's entry
__device__ __forceinline__ void cnt_hand_on(unsigned &low, unsigned a)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, f;\n"
        "and.b32 f, %0, 1;\n"
        "setp.ne.u32 p, f, 0;\n"
        "and.b32 t, %0, 0xFFFFFFFC;\n"
        "@p red.shared.add.u32 [t + %2], %1;\n"
        "@p ld.shared.u32 %0, [t];\n"
        "}\n"
        : "+r"(low)
        : "r"(a), "n"(CNT_DELTA)
        : "memory");
}
// round, second half: publish the new entry, snapshot the own arrivals
__device__ __forceinline__ unsigned cnt_publish(unsigned own_addr, unsigned low)
{
    unsigned a;
    asm volatile(
        "st.shared.u32 [%1], %2;\n"
        "ld.shared.u32 %0, [%1 + %3];\n"
        : "=r"(a)
        : "r"(own_addr), "r"(low), "n"(CNT_DELTA)
        : "memory");
    return a;
}

template <int WMODE>
__global__ void __launch_bounds__(THREADS, 4)
acc_local_cnt_kernel(const __grid_constant__ CUtensorMap map, const AccArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *pk = smem_raw;
    uint32_t *ptr32 = reinterpret_cast<uint32_t *>(smem_raw + SMC_PTR); // entries
    uint32_t *arr = reinterpret_cast<uint32_t *>(smem_raw + SMC_ARR);   // own value + arrivals
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw + SMC_LUT);   // [0..15] as stored, [16..31] transposed
    uint64_t &bar = *reinterpret_cast<uint64_t *>(smem_raw + SMC_BAR);
    int *ctl = reinterpret_cast<int *>(smem_raw + SMC_BAR + 8); // [0] entries of the tile, [1] their place in the list, [2] orientation vote
    static_assert(SMC_ARR - SMC_PTR == CNT_DELTA, "arrival word = entry + CNT_DELTA");

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int64_t x0 = (int64_t)tx * T, y0 = (int64_t)ty * T;
    if (tid == 0) {
        ctl[0] = 0;
        ctl[2] = 0;
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, RH * RP);
        ht::tma_load_2d(pk, &map, (int)x0 - XO, (int)y0 - 1 + p.halo_top, &bar);
    }
    if (tid < 32) { // added to the own entry address; codes 0, 9..15 never link
        const int code = tid & 15;
        const int off = tid < 16 ? link_offset_of(code) : code_dc(code) * T + code_dr(code);
        lut[tid] = link_offset_of(code) ? (unsigned)(off * 4) + J_ACTIVE : 0u;
    }
    const int th = (int)min((int64_t)T, p.rows - y0), tw = (int)min((int64_t)T, p.cols - x0);
    // thread -> array cells tid + 256 k: one array column u, every 4th array row.  A warp
    // covers 32 consecutive array cells in every step (conflict-free own accesses).
    const int u = tid & (T - 1), v0 = tid >> 6;
    const unsigned pbase = ht::smem_u32(ptr32), own0 = pbase + 4u * (unsigned)tid;
    __syncthreads();
    ht::mbar_wait(&bar, 0);

    // ---- orientation vote: 4 sampled cells per thread, displacement over two steps
    {
        int score = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int at = (v0 + 16 * j + 1) * RP + XO + u;
            const unsigned b = pk[at];
            const int c1 = (int)(b & HT_DIR_CODE);
            if (!(b & FLAGS_STOP) && c1) {
                int ddr = code_dr(c1), ddc = code_dc(c1);
                const unsigned b2 = pk[at + ddr * RP + ddc];
                const int c2 = (int)(b2 & HT_DIR_CODE);
                if (!(b2 & FLAGS_STOP) && c2) {
                    ddr += code_dr(c2);
                    ddc += code_dc(c2);
                }
                score += (abs(ddr) > abs(ddc)) - (abs(ddr) < abs(ddc));
            }
        }
        score = __reduce_add_sync(0xffffffffu, score);
        if (lane == 0 && score)
            atomicAdd(&ctl[2], score);
    }
    __syncthreads();
    const bool tr = ctl[2] > 0; // flow runs along the raster's columns: keep the arrays transposed
    const uint32_t *lutm = lut + (tr ? 16 : 0);

    // ---- per cell: low = pointer entry, a = sum to hand on (own value + arrivals so far)
    unsigned low[CPT], a[CPT];
    {
        // u is the raster column (as stored) or the raster row (transposed)
        const int lim_u = tr ? th : tw, lim_v = tr ? tw : th;
        const unsigned lo_u = tr ? EX_N : EX_W, hi_u = tr ? EX_S : EX_E;
        const unsigned lo_v = tr ? EX_W : EX_N, hi_v = tr ? EX_E : EX_S;
        // codes that do not link from this array column; bit 15: outside the raster; bit 0: code 0
        const unsigned cm = (u == 0 ? lo_u : 0u) | (u == lim_u - 1 ? hi_u : 0u) | (u >= lim_u ? 0xFFFFu : 0u) | 1u;
        // packed bytes of the 16 cells.  As stored: a warp reads 32 consecutive bytes of a
        // row.  Transposed: the cells are bytes v0, v0 + 4, ... of ONE raster row -- four
        // 128-bit loads fetch the row, byte v0 of every word is a cell (byte loads down a
        // column of the 96-byte pitch would be 6-way bank conflicts).
        unsigned bw[CPT];
        if (tr) {
            const uint4 *row = reinterpret_cast<const uint4 *>(pk + (u + 1) * RP + XO);
#pragma unroll
            for (int q = 0; q < CPT / 4; q++) {
                const uint4 w4 = row[q];
                bw[4 * q] = w4.x; bw[4 * q + 1] = w4.y; bw[4 * q + 2] = w4.z; bw[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int k = 0; k < CPT; k++)
                bw[k] = (bw[k] >> (8 * v0)) & 0xFFu;
        } else {
#pragma unroll
            for (int k = 0; k < CPT; k++)
                bw[k] = pk[(v0 + 1 + 4 * k) * RP + XO + u];
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int v = v0 + 4 * k;
            const unsigned b = bw[k];
            const unsigned ex = cm | (v == 0 ? lo_v : 0u) | (v == lim_v - 1 ? hi_v : 0u) | (v >= lim_v ? 0xFFFFu : 0u);
            const unsigned code = b & HT_DIR_CODE;
            // no link: a stop flag, or the receiver is outside the tile / the raster
            const bool stop = (b & FLAGS_STOP) || ((ex >> code) & 1u);
            const bool valid = !(b & HT_DIR_NULL) && !(ex >> 15);
            a[k] = WMODE == W_ONE ? (valid ? 1u : 0u) : ((valid && (b & HT_DIR_TAINT)) ? 1u : 0u);
            low[k] = own0 + (unsigned)(k * THREADS * 4) + (stop ? 0u : lutm[code]);
            ptr32[tid + k * THREADS] = low[k];
            arr[tid + k * THREADS] = a[k];
        }
    }
    __syncthreads();

    // ---- pointer doubling
    for (int round = 0; round < MAX_ROUNDS; round++) {
#pragma unroll
        for (int k = 0; k < CPT; k++)
            cnt_hand_on(low[k], a[k]);
        __syncthreads();
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            a[k] = cnt_publish(own0 + (unsigned)(k * THREADS * 4), low[k]);
            any |= low[k];
        }
        if (!__syncthreads_or((int)(any & J_ACTIVE)))
            break;
    }

    // ---- boundary graph.  thread -> one perimeter slot
    {
        int ly = 0, plx = 0;
        const bool per = perimeter_cell(tid, th, tw, ly, plx);
        const unsigned b = per ? pk[(ly + 1) * RP + plx + XO] : (unsigned)HT_DIR_NULL;
        // flow that leaves the tile through this cell
        if (per && leaves_tile(b, ly, plx, th, tw)) {
            const unsigned val = arr[arr_index(tr, ly, plx)];
            if (val)
                atomicAdd(reinterpret_cast<unsigned long long *>(p.base) +
                              exit_slot(p, tx, ty, ly, plx, (int)(b & HT_DIR_CODE)),
                          (unsigned long long)val);
        }
        // entry cell: a cell outside the tile drains into it
        bool entry = false;
        if (per && !(b & HT_DIR_NULL)) {
            const int yy = ly + 1, xx = plx + XO;
#pragma unroll
            for (int kk = 0; kk < 9; kk++) {
                if (kk == 4)
                    continue;
                const int ny = yy + kk / 3 - 1, nx = xx + kk % 3 - 1;
                if (ny >= 1 && ny <= th && nx >= XO && nx < XO + tw)
                    continue; // inside the tile
                const unsigned w = pk[ny * RP + nx]; // outside the raster: TMA filled 0
                entry |= !(w & FLAGS_STOP) && (w & HT_DIR_CODE) == (unsigned)win2code(8 - kk);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, entry);
        int at = 0;
        if (lane == 0 && m)
            at = atomicAdd(&ctl[0], __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (tid == 0) {
            if (ctl[0])
                ctl[1] = atomicAdd(p.nlist, ctl[0]);
            if (p.tile_cnt) {
                p.tile_first[tile] = ctl[0] ? ctl[1] : 0;
                p.tile_cnt[tile] = ctl[0];
            }
        }
        __syncthreads();
        if (entry) {
            // where the path that enters here leaves the tile: the root of the cell
            const int rc = (int)(((ptr32[arr_index(tr, ly, plx)] & ~3u) - pbase) >> 2);
            const int cy = tr ? (rc & (T - 1)) : (rc >> 6), cx = tr ? (rc >> 6) : (rc & (T - 1));
            const unsigned rb = pk[(cy + 1) * RP + cx + XO];
            const int32_t slot = tile * SLOTS + tid;
            p.list[ctl[1] + at] = slot;
            p.next[slot] = leaves_tile(rb, cy, cx, th, tw)
                               ? exit_slot(p, tx, ty, cy, cx, (int)(rb & HT_DIR_CODE))
                               : TERMINAL;
        }
    }

    // ---- tile-local counts for stage C (u16; a warp stores 64 consecutive bytes).
    // a[] holds them: the last snapshot was taken after the last hand-off.
    if (tr) {
        // transposed arrays: turn them through a padded u16 buffer (33-word pitch: both the
        // column-wise writes and the row-wise reads are conflict-free) laid over the entries
        __syncthreads();
        uint16_t *buf = reinterpret_cast<uint16_t *>(ptr32);
#pragma unroll
        for (int k = 0; k < CPT; k++)
            buf[u * (T + 2) + v0 + 4 * k] = (uint16_t)a[k]; // (ly = u, lx = v)
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPT; k++)
            a[k] = buf[(v0 + 4 * k) * (T + 2) + u];
    }
    if (u < tw) {
        uint16_t *dst = p.lacc + (y0 + v0) * p.lacc_ld + x0 + u;
        const int64_t step = 4 * p.lacc_ld;
#pragma unroll
        for (int k = 0; k < CPT; k++, dst += step)
            if (v0 + 4 * k < th)
                *dst = (uint16_t)a[k];
    }
}

// ---------------------------------------------------------------------------
// Weighted mode (fp32 weights, sums reported as fp64): EXACT FIXED POINT.
// Shared-memory fp64 (and 64-bit integer) atomicAdd is a compare-and-swap loop on
// sm_100a (SASS ATOMS.CAST.SPIN.64); 32-bit integer adds are native (ATOMS.ADD).  So
// the weights are quantised once and every sum downstream is integer arithmetic:
//   * global unit 2^-(sg + 24), sg = G_BITS - ilogb(largest |weight| of the raster): the
//     largest weight is ~2^88 units, a sum over 2^34 cells still fits 128 bits; the low
//     24 bits of every value are zero and carry the band-edge marks of a sharded run
//     (as bits 40.. of the u64 counts do);
//   * stage A works in a per-tile unit 2^(et - sg), et >= 0 chosen so that the tile's
//     largest weight is < 2^Q_BITS units: tile-local sums fit 64 bits and live in
//     shared memory as two 32-bit words (native atomic add, carry handed on);
//   * flow that leaves a tile, the boundary graph (stage B) and stage C are 128-bit
//     integers in the global unit (two native 64-bit global atomics with carry).
// Integer addition is associative, so the result does not depend on the order of the
// atomics: weighted accumulation is bit-reproducible from run to run.  Quantisation:
// every nonzero weight keeps at least Q_KEEP = 26 significant bits (relative error
// <= 2^-27, below fp32's own resolution) as long as the weights of its 64 x 64 tile span
// less than 2^(Q_BITS - Q_KEEP) = 4 million to one; beyond that the error is at most
// 2^-(Q_BITS+1) of the tile's largest weight.  The contract is 1e-5 relative.
// Stage A leaves the tile-local sums (int64, 8 B/cell) for stage C, which adds the
// resolved inflow along the entry paths with walkers, like the count mode.
// ---------------------------------------------------------------------------
constexpr int Q_BITS = 48; // at most: the tile's largest weight < 2^48 tile units (sums fit 64 bits)
constexpr int Q_MIN = 30;  // at least
constexpr int Q_KEEP = 26; // significant bits every nonzero weight of a tile keeps (fp32 has 24)
constexpr int G_BITS = 64;
constexpr int W_MARK_BITS = 24; // global-unit values are multiples of 2^24: band-edge marks ride below

__device__ __forceinline__ int scale_exp_of(unsigned gmax_bits)
{
    const float g = __uint_as_float(gmax_bits);
    return g > 0.0f ? G_BITS - ilogbf(g) : 0;
}
__device__ __forceinline__ double pow2d(int e) // -1022 <= e <= 1023
{
    return __longlong_as_double((long long)(1023 + e) << 52);
}
// nodata, NaN and +-inf weigh nothing
__device__ __forceinline__ float clean_weight(float w, int has_wnd, float wnd)
{
    return ((has_wnd && w == wnd) || !(fabsf(w) <= 3.4028234e38f)) ? 0.0f : w;
}

__global__ void __launch_bounds__(256)
weight_max_kernel(const float *w, int64_t w_ld, int64_t rows, int64_t cols, int has_wnd, float wnd,
                  unsigned *out_bits)
{
    float m = 0.0f;
    // a warp reads 128 consecutive floats of a row as float4 when the row allows it
    const bool vec = !((uintptr_t)w & 15) && !(w_ld & 3);
    const int64_t quads = vec ? cols / 4 : 0;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = gtid; i < rows * quads; i += gsize) {
        const int64_t r = i / quads, q = i - r * quads;
        const float4 v = *reinterpret_cast<const float4 *>(w + r * w_ld + 4 * q);
        m = fmaxf(fmaxf(m, fabsf(clean_weight(v.x, has_wnd, wnd))), fabsf(clean_weight(v.y, has_wnd, wnd)));
        m = fmaxf(fmaxf(m, fabsf(clean_weight(v.z, has_wnd, wnd))), fabsf(clean_weight(v.w, has_wnd, wnd)));
    }
    const int64_t tail = cols - 4 * quads; // columns the quads do not cover
    for (int64_t i = gtid; i < rows * tail; i += gsize) {
        const int64_t r = i / tail, c = 4 * quads + (i - r * tail);
        m = fmaxf(m, fabsf(clean_weight(w[r * w_ld + c], has_wnd, wnd)));
    }
    const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m)); // |x| bits order like x
    if ((threadIdx.x & 31) == 0 && wm)
        atomicMax(out_bits, wm);
}

// 64-bit add into two 32-bit shared words
__device__ __forceinline__ void smem_add64(uint32_t *lo, uint32_t *hi, long long v)
{
    const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)((unsigned long long)v >> 32);
    uint32_t carry = 0u;
    if (vlo) {
        const uint32_t old = atomicAdd(lo, vlo);
        carry = (uint32_t)(old + vlo) < old ? 1u : 0u;
    }
    const uint32_t h = vhi + carry;
    if (h)
        atomicAdd(hi, h);
}

constexpr int SMW_PTR = SM_EFF;                  // uint32[T*T] pointer entries (shared addresses | active)
constexpr int SMW_LO = SMW_PTR + T * T * 4;      // uint32[T*T] low words of the sums
constexpr int SMW_HI = SMW_LO + T * T * 4;       // uint32[T*T] high words
constexpr int SMW_LUT = SMW_HI + T * T * 4;      // uint32[2][16] code -> pointer step | active (as stored, transposed)
constexpr int SMW_BAR = SMW_LUT + 128;           // mbarrier, counters, tile maximum
constexpr int SMW_TOTAL = SMW_BAR + 32;
constexpr int W_DLO = SMW_LO - SMW_PTR, W_DHI = SMW_HI - SMW_PTR;

// The rounds of the weighted kernel, branch-free on 32-bit shared addresses like the count
// kernel's.  A 64-bit hand-off is a returning add of the low word (the carry needs the old
// value) and a fire-and-forget add of the high word; the returning adds of a GROUP of cells
// are issued back to back, their carries are resolved afterwards, so that the latencies of
// the atomics overlap instead of adding up.
__device__ __forceinline__ unsigned w_add_lo(unsigned low, unsigned alo)
{
    unsigned old;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, f;\n"
        "and.b32 f, %1, 1;\n"
        "setp.ne.u32 p, f, 0;\n"
        "and.b32 t, %1, 0xFFFFFFFC;\n"
        "mov.b32 %0, 0;\n"
        "@p atom.shared.add.u32 %0, [t + %3], %2;\n"
        "}\n"
        : "=r"(old)
        : "r"(low), "r"(alo), "n"(W_DLO)
        : "memory");
    return old;
}
// carry of the low add into the high word, then the target's entry (active cells only)
__device__ __forceinline__ void w_add_hi(unsigned &low, unsigned old, unsigned alo, unsigned ahi)
{
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 t, f, s, h;\n"
        "and.b32 f, %0, 1;\n"
        "setp.ne.u32 p, f, 0;\n"
        "and.b32 t, %0, 0xFFFFFFFC;\n"
        "add.u32 s, %1, %2;\n"
        "setp.lt.u32 q, s, %1;\n"
        "selp.u32 h, 1, 0, q;\n"
        "add.u32 h, h, %3;\n"
        "setp.ne.and.u32 q, h, 0, p;\n"
        "@q red.shared.add.u32 [t + %4], h;\n"
        "@p ld.shared.u32 %0, [t];\n"
        "}\n"
        : "+r"(low)
        : "r"(old), "r"(alo), "r"(ahi), "n"(W_DHI)
        : "memory");
}
// publish the new entry, snapshot the own sums
__device__ __forceinline__ void w_publish(unsigned own_addr, unsigned low, unsigned &alo, unsigned &ahi)
{
    asm volatile(
        "st.shared.u32 [%2], %3;\n"
        "ld.shared.u32 %0, [%2 + %4];\n"
        "ld.shared.u32 %1, [%2 + %5];\n"
        : "=r"(alo), "=r"(ahi)
        : "r"(own_addr), "r"(low), "n"(W_DLO), "n"(W_DHI)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 4)
acc_local_w_kernel(const __grid_constant__ CUtensorMap map, const AccArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *pk = smem_raw;
    uint32_t *ptr32 = reinterpret_cast<uint32_t *>(smem_raw + SMW_PTR);
    uint32_t *alo = reinterpret_cast<uint32_t *>(smem_raw + SMW_LO);
    uint32_t *ahi = reinterpret_cast<uint32_t *>(smem_raw + SMW_HI);
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw + SMW_LUT);
    uint64_t &bar = *reinterpret_cast<uint64_t *>(smem_raw + SMW_BAR);
    int *ctl = reinterpret_cast<int *>(smem_raw + SMW_BAR + 8); // [0] entries, [1] their place in the list
    unsigned *tmaxb = reinterpret_cast<unsigned *>(smem_raw + SMW_BAR + 16);
    unsigned *tminb = reinterpret_cast<unsigned *>(smem_raw + SMW_BAR + 20);
    int *vote = reinterpret_cast<int *>(smem_raw + SMW_BAR + 24);

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int64_t x0 = (int64_t)tx * T, y0 = (int64_t)ty * T;
    if (tid == 0) {
        ctl[0] = 0;
        *vote = 0;
        *tmaxb = 0u;
        *tminb = 0x7F800000u;
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, RH * RP);
        ht::tma_load_2d(pk, &map, (int)x0 - XO, (int)y0 - 1 + p.halo_top, &bar);
    }
    if (tid < 32) {
        const int code = tid & 15;
        const int off = tid < 16 ? link_offset_of(code) : code_dc(code) * T + code_dr(code);
        lut[tid] = link_offset_of(code) ? (unsigned)(off * 4) + J_ACTIVE : 0u;
    }
    const int th = (int)min((int64_t)T, p.rows - y0), tw = (int)min((int64_t)T, p.cols - x0);
    // array cell (u, v0 + 4k) of the thread: raster (row v, column u) as stored, (row u, column v)
    // when the tile keeps its arrays transposed (see acc_local_cnt_kernel)
    const int u = tid & (T - 1), v0 = tid >> 6;
    const unsigned pbase = ht::smem_u32(ptr32), own0 = pbase + 4u * (unsigned)tid;
    // the thread's weights in RASTER order (coalesced) while the TMA is in flight
    float wv[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int ly = v0 + 4 * k;
        wv[k] = (u < tw && ly < th) ? clean_weight(p.w[(y0 + ly) * p.w_ld + x0 + u], p.has_wnd, p.wnd) : 0.0f;
    }
    __syncthreads();
    ht::mbar_wait(&bar, 0);

    // ---- NULL cells weigh nothing; tile extremes; orientation vote
    {
        float m = 0.0f, mn = __uint_as_float(0x7F800000u);
        int score = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int at = (v0 + 4 * k + 1) * RP + XO + u;
            const unsigned b = pk[at];
            if (b & HT_DIR_NULL)
                wv[k] = 0.0f;
            const float aw = fabsf(wv[k]);
            m = fmaxf(m, aw);
            mn = aw > 0.0f ? fminf(mn, aw) : mn;
            if ((k & 3) == 0) { // 4 sampled cells per thread, displacement over two steps
                const int c1 = (int)(b & HT_DIR_CODE);
                if (!(b & FLAGS_STOP) && c1) {
                    int ddr = code_dr(c1), ddc = code_dc(c1);
                    const unsigned b2 = pk[at + ddr * RP + ddc];
                    const int c2 = (int)(b2 & HT_DIR_CODE);
                    if (!(b2 & FLAGS_STOP) && c2) {
                        ddr += code_dr(c2);
                        ddc += code_dc(c2);
                    }
                    score += (abs(ddr) > abs(ddc)) - (abs(ddr) < abs(ddc));
                }
            }
        }
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        const unsigned wn = __reduce_min_sync(0xffffffffu, __float_as_uint(mn));
        score = __reduce_add_sync(0xffffffffu, score);
        if (lane == 0) {
            if (wm) {
                atomicMax(tmaxb, wm);
                atomicMin(tminb, wn);
            }
            if (score)
                atomicAdd(vote, score);
        }
    }
    __syncthreads();
    const bool tr = *vote > 0; // flow runs along the raster's columns: arrays transposed
    const uint32_t *lutm = lut + (tr ? 16 : 0);
    if (tr) {
        // the weights were read in raster order: turn them through a padded staging block
        // (pitch 65 words: both the row-wise writes and the column-wise reads are
        // conflict-free) laid over the sum planes, which are not initialised yet
        float *stage = reinterpret_cast<float *>(alo);
#pragma unroll
        for (int k = 0; k < CPT; k++)
            stage[u * (T + 1) + v0 + 4 * k] = wv[k]; // raster (row v0 + 4k, column u)
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPT; k++)
            wv[k] = stage[(v0 + 4 * k) * (T + 1) + u]; // raster (row u, column v0 + 4k)
        __syncthreads();
    }
    // per-tile unit: fine enough that the tile's SMALLEST weight keeps Q_KEEP significant bits,
    // never coarser than what lets its largest weight fit Q_BITS bits, never finer than the
    // global unit; and not finer than needed -- sums below 2^32 units cost one shared atomic
    // instead of two (a raster like precipitation, weights within a factor of two: Q_MIN bits)
    const int sg = scale_exp_of(*p.gmax_bits);
    const float tmax = __uint_as_float(*tmaxb), tmin = __uint_as_float(*tminb);
    int et = 0;
    if (tmax > 0.0f) {
        const int span = ilogbf(tmax) - ilogbf(tmin);                   // >= 0
        const int qbits = min(Q_BITS, max(Q_MIN, span + Q_KEEP));       // bits of the largest weight
        et = max(0, ilogbf(tmax) + 1 + sg - qbits);
    }
    const double sc = pow2d(sg - et);

    // ---- per cell: pointer entry and own value
    unsigned low[CPT], slo[CPT], shi[CPT];
    {
        const int lim_u = tr ? th : tw, lim_v = tr ? tw : th;
        const unsigned lo_u = tr ? EX_N : EX_W, hi_u = tr ? EX_S : EX_E;
        const unsigned lo_v = tr ? EX_W : EX_N, hi_v = tr ? EX_E : EX_S;
        const unsigned cm = (u == 0 ? lo_u : 0u) | (u == lim_u - 1 ? hi_u : 0u) | (u >= lim_u ? 0xFFFFu : 0u) | 1u;
        unsigned bw[CPT];
        if (tr) {
            const uint4 *row = reinterpret_cast<const uint4 *>(pk + (u + 1) * RP + XO);
#pragma unroll
            for (int q = 0; q < CPT / 4; q++) {
                const uint4 w4 = row[q];
                bw[4 * q] = w4.x; bw[4 * q + 1] = w4.y; bw[4 * q + 2] = w4.z; bw[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int k = 0; k < CPT; k++)
                bw[k] = (bw[k] >> (8 * v0)) & 0xFFu;
        } else {
#pragma unroll
            for (int k = 0; k < CPT; k++)
                bw[k] = pk[(v0 + 1 + 4 * k) * RP + XO + u];
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int v = v0 + 4 * k;
            const unsigned b = bw[k];
            const unsigned ex = cm | (v == 0 ? lo_v : 0u) | (v == lim_v - 1 ? hi_v : 0u) | (v >= lim_v ? 0xFFFFu : 0u);
            const unsigned code = b & HT_DIR_CODE;
            const bool stop = (b & FLAGS_STOP) || ((ex >> code) & 1u);
            const long long a = __double2ll_rn((double)wv[k] * sc);
            slo[k] = (uint32_t)a;
            shi[k] = (uint32_t)((unsigned long long)a >> 32);
            low[k] = own0 + (unsigned)(k * THREADS * 4) + (stop ? 0u : lutm[code]);
            ptr32[tid + k * THREADS] = low[k];
            alo[tid + k * THREADS] = slo[k];
            ahi[tid + k * THREADS] = shi[k];
        }
    }
    __syncthreads();

    // ---- pointer doubling; the sums include the cell's own weight
    for (int round = 0; round < MAX_ROUNDS; round++) {
#pragma unroll
        for (int g = 0; g < CPT; g += 8) {
            unsigned old[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                old[k] = w_add_lo(low[g + k], slo[g + k]);
#pragma unroll
            for (int k = 0; k < 8; k++)
                w_add_hi(low[g + k], old[k], slo[g + k], shi[g + k]);
        }
        __syncthreads();
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            w_publish(own0 + (unsigned)(k * THREADS * 4), low[k], slo[k], shi[k]);
            any |= low[k];
        }
        if (!__syncthreads_or((int)(any & J_ACTIVE)))
            break;
    }

    // ---- boundary graph.  thread -> one perimeter slot
    long long exit_val = 0;
    int32_t exit_to = 0;
    {
        int ly = 0, plx = 0;
        const bool per = perimeter_cell(tid, th, tw, ly, plx);
        const unsigned b = per ? pk[(ly + 1) * RP + plx + XO] : (unsigned)HT_DIR_NULL;
        // flow that leaves the tile through this cell: read now, added to the boundary graph at
        // the very end -- the 128-bit add is two dependent global atomics (the carry), and the
        // barriers below would make the whole CTA wait for them
        if (per && leaves_tile(b, ly, plx, th, tw)) {
            const int at = arr_index(tr, ly, plx);
            exit_val = (long long)(((unsigned long long)ahi[at] << 32) | alo[at]);
            exit_to = exit_slot(p, tx, ty, ly, plx, (int)(b & HT_DIR_CODE));
        }
        bool entry = false;
        if (per && !(b & HT_DIR_NULL)) {
            const int yy = ly + 1, xx = plx + XO;
#pragma unroll
            for (int kk = 0; kk < 9; kk++) {
                if (kk == 4)
                    continue;
                const int ny = yy + kk / 3 - 1, nx = xx + kk % 3 - 1;
                if (ny >= 1 && ny <= th && nx >= XO && nx < XO + tw)
                    continue; // inside the tile
                const unsigned w = pk[ny * RP + nx]; // outside the raster: TMA filled 0
                entry |= !(w & FLAGS_STOP) && (w & HT_DIR_CODE) == (unsigned)win2code(8 - kk);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, entry);
        int at = 0;
        if (lane == 0 && m)
            at = atomicAdd(&ctl[0], __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (tid == 0 && ctl[0])
            ctl[1] = atomicAdd(p.nlist, ctl[0]);
        __syncthreads();
        if (entry) {
            const int rc = (int)(((ptr32[arr_index(tr, ly, plx)] & ~3u) - pbase) >> 2);
            const int cy = tr ? (rc & (T - 1)) : (rc >> 6), cx = tr ? (rc >> 6) : (rc & (T - 1));
            const unsigned rb = pk[(cy + 1) * RP + cx + XO];
            const int32_t slot = tile * SLOTS + tid;
            p.list[ctl[1] + at] = slot;
            p.next[slot] = leaves_tile(rb, cy, cx, th, tw)
                               ? exit_slot(p, tx, ty, cy, cx, (int)(rb & HT_DIR_CODE))
                               : TERMINAL;
        }
    }

    // ---- tile-local sums (tile unit) and the tile's unit for stage C.  slo / shi hold them:
    // the last snapshot was taken after the last hand-off.
    if (tid == 0)
        p.tile_e[tile] = et;
    if (tr) {
        // transposed arrays: back to raster order through a padded 64-bit staging block laid
        // over the entries and the sum planes (everybody is done with them after the barrier)
        __syncthreads();
        unsigned long long *stage = reinterpret_cast<unsigned long long *>(ptr32);
#pragma unroll
        for (int k = 0; k < CPT; k++)
            stage[u * (T + 1) + v0 + 4 * k] = ((unsigned long long)shi[k] << 32) | slo[k]; // (row u, column v)
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const unsigned long long v = stage[(v0 + 4 * k) * (T + 1) + u];
            slo[k] = (uint32_t)v;
            shi[k] = (uint32_t)(v >> 32);
        }
    }
    if (u < tw) {
        long long *dst = p.lacc64 + (y0 + v0) * p.lacc64_ld + x0 + u;
        const int64_t step = 4 * p.lacc64_ld;
#pragma unroll
        for (int k = 0; k < CPT; k++, dst += step)
            if (v0 + 4 * k < th)
                *dst = (long long)(((unsigned long long)shi[k] << 32) | slo[k]);
    }
    if (exit_val)
        atomic_add_v(reinterpret_cast<i128 *>(p.base) + exit_to, shl_i64(exit_val, et + W_MARK_BITS));
}

// stage C, weighted: final = tile-local sum (shifted to the global unit) + the resolved
// inflow of every entry upstream, added along the entry's path by a walker.  Cells are
// four 32-bit word planes in shared memory (128-bit integers, native 32-bit atomic adds
// with the carry handed on).  Measured and rejected: five planes of 24-bit limbs that never
// carry (fire-and-forget REDs, the walk bound by one byte load per step) -- 95 KB per CTA
// leave two CTAs per SM instead of three, 1.45 ms against 1.09.
constexpr int WP = T + 4;                       // plane row pitch in words (see CP below)
constexpr int SMX_W = T * T;                    // after the 64 x 64 packed tile
constexpr int SMX_BAR = SMX_W + 4 * T * WP * 4;
constexpr int SMX_TOTAL = SMX_BAR + 16;

__global__ void __launch_bounds__(THREADS, 3)
acc_final_w_kernel(const __grid_constant__ CUtensorMap map, const AccArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.tile_phase && p.tile_phase[blockIdx.x] != p.phase)
        return; // the other pass of a sharded run takes this tile
    uint8_t *pk = smem_raw; // 64 x 64, no ring
    uint32_t *wpl = reinterpret_cast<uint32_t *>(smem_raw + SMX_W); // [4][T * WP]
    uint64_t &bar = *reinterpret_cast<uint64_t *>(smem_raw + SMX_BAR);
    constexpr int PLANE = T * WP;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int64_t x0 = (int64_t)tx * T, y0 = (int64_t)ty * T;
    if (tid == 0) {
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, T * T);
        ht::tma_load_2d(pk, &map, (int)x0, (int)y0 + p.halo_top, &bar);
    }
    const int th = (int)min((int64_t)T, p.rows - y0), tw = (int)min((int64_t)T, p.cols - x0);
    const int q4 = 4 * (tid & 15), r0 = tid >> 4;
    const int cbase = r0 * WP + q4;
    const bool wide = q4 + 3 < tw;
    const int et = p.tile_e[tile];
    const int sg = scale_exp_of(*p.gmax_bits);

    // tile-local sums -> 128-bit words in the global unit
    {
        const long long *src = p.lacc64 + (y0 + r0) * p.lacc64_ld + x0 + q4;
        const int64_t step = 16 * p.lacc64_ld;
#pragma unroll
        for (int i = 0; i < 4; i++, src += step) {
            long long v[4] = {0ll, 0ll, 0ll, 0ll};
            if (r0 + 16 * i < th && q4 < tw) {
                if (wide) {
                    const longlong2 u0 = *reinterpret_cast<const longlong2 *>(src);
                    const longlong2 u1 = *reinterpret_cast<const longlong2 *>(src + 2);
                    v[0] = u0.x; v[1] = u0.y; v[2] = u1.x; v[3] = u1.y;
                }
                else
                    for (int j = 0; j < 4 && q4 + j < tw; j++)
                        v[j] = src[j];
            }
            uint32_t wd[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const i128 g = shl_i64(v[j], et + W_MARK_BITS);
                wd[0][j] = (uint32_t)g.lo; wd[1][j] = (uint32_t)(g.lo >> 32);
                wd[2][j] = (uint32_t)g.hi; wd[3][j] = (uint32_t)((unsigned long long)g.hi >> 32);
            }
#pragma unroll
            for (int pl = 0; pl < 4; pl++)
                *reinterpret_cast<uint4 *>(wpl + pl * PLANE + cbase + i * 16 * WP) =
                    make_uint4(wd[pl][0], wd[pl][1], wd[pl][2], wd[pl][3]);
        }
    }
    int ely = 0, elx = 0;
    i128 inflow(0);
    if (perimeter_cell(tid, th, tw, ely, elx)) {
        inflow = reinterpret_cast<const i128 *>(p.inflow)[(int64_t)tile * SLOTS + tid];
        inflow.lo &= ~((1ull << W_MARK_BITS) - 1ull); // band-edge marks of a sharded run
    }
    __syncthreads();
    ht::mbar_wait(&bar, 0);

    // ---- walkers: one per entry cell with inflow
    if (inflow != i128(0)) {
        const uint32_t v[4] = {(uint32_t)inflow.lo, (uint32_t)(inflow.lo >> 32), (uint32_t)inflow.hi,
                               (uint32_t)((unsigned long long)inflow.hi >> 32)};
        int ly = ely, lx = elx;
        for (int step = 0; step < T * T; step++) { // the bound only matters for a cyclic raster
            uint32_t *cell = wpl + ly * WP + lx;
            uint32_t carry = 0u;
#pragma unroll
            for (int pl = 0; pl < 4; pl++) {
                const uint32_t add = v[pl] + carry;
                if (carry && add == 0u)
                    continue; // v[pl] = 0xFFFFFFFF: the carry runs on
                carry = 0u;
                if (add) {
                    const uint32_t old = atomicAdd(cell + pl * PLANE, add);
                    carry = (uint32_t)(old + add) < old ? 1u : 0u;
                }
            }
            const unsigned b = pk[ly * T + lx];
            if (!b || leaves_tile(b, ly, lx, th, tw) || (b & FLAGS_STOP))
                break;
            const int code = (int)(b & HT_DIR_CODE);
            if (!link_offset_of(code))
                break;
            ly += code_dr(code);
            lx += code_dc(code);
        }
    }
    __syncthreads();

    // ---- outputs: 4 cells per thread and pass, NaN where NULL
    if (!p.acc)
        return;
    const double unit = pow2d(-sg - W_MARK_BITS);
    double *accp = p.acc + (y0 + r0) * p.acc_ld + x0 + q4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ly = r0 + 16 * i;
        if (ly >= th || q4 >= tw)
            break;
        uint4 w4[4];
#pragma unroll
        for (int pl = 0; pl < 4; pl++)
            w4[pl] = *reinterpret_cast<const uint4 *>(wpl + pl * PLANE + cbase + i * 16 * WP);
        const unsigned wq = *reinterpret_cast<const unsigned *>(pk + ly * T + q4);
        const uint32_t *w0 = reinterpret_cast<const uint32_t *>(&w4[0]);
        const uint32_t *w1 = reinterpret_cast<const uint32_t *>(&w4[1]);
        const uint32_t *w2 = reinterpret_cast<const uint32_t *>(&w4[2]);
        const uint32_t *w3 = reinterpret_cast<const uint32_t *>(&w4[3]);
        double a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool null = (wq >> (8 * j)) & HT_DIR_NULL;
            const i128 g(((unsigned long long)w1[j] << 32) | w0[j],
                         (long long)(((unsigned long long)w3[j] << 32) | w2[j]));
            a[j] = null ? __longlong_as_double(0x7ff8000000000000LL) : i128_to_double(g) * unit;
        }
        double *dst = accp + (int64_t)(16 * i) * p.acc_ld;
        if (wide && !(((uintptr_t)dst) & 15)) {
            reinterpret_cast<double2 *>(dst)[0] = make_double2(a[0], a[1]);
            reinterpret_cast<double2 *>(dst)[1] = make_double2(a[2], a[3]);
        }
        else
            for (int j = 0; j < 4 && q4 + j < tw; j++)
                dst[j] = a[j];
    }
}

// stage C cell words: lo32 = [31:24] signed receiver offset inside the tile (0 =
// none) | [23:0] low part of the count; hi32 = count >> 16.  A walker adds its
// inflow split 16 / rest, so the low field never carries into the offset
// (at most 252 walkers -- one per perimeter cell -- x 65535 + 4096 < 2^24), and the
// atomic that adds the low part returns the link to follow.
// Rows of the two count arrays are CP = 68 words apart: with a pitch of 64 every cell
// of the tile's first / last column falls into one bank, and the walkers that start
// there (one lane per perimeter cell, moving in step along parallel paths) would
// serialise 32-way on every atomic.
constexpr int CP = T + 4;
constexpr int SMF_LO = T * T;               // after the 64x64 packed tile
constexpr int SMF_HI = SMF_LO + T * CP * 4;
constexpr int SMF_LUT = SMF_HI + T * CP * 4; // uint32[16] code -> offset field
constexpr int SMF_BAR = SMF_LUT + 64;
constexpr int SMF_TOTAL = SMF_BAR + 16;

template <int WMODE>
__global__ void __launch_bounds__(THREADS, 5) // 39 KB of shared memory per CTA: five fit an SM
acc_final_cnt_kernel(const __grid_constant__ CUtensorMap map, const AccArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.tile_phase && p.tile_phase[blockIdx.x] != p.phase)
        return; // the other pass of a sharded run takes this tile
    uint8_t *pk = smem_raw; // 64 x 64, no ring
    uint32_t *lo32 = reinterpret_cast<uint32_t *>(smem_raw + SMF_LO);
    uint32_t *hi32 = reinterpret_cast<uint32_t *>(smem_raw + SMF_HI);
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw + SMF_LUT);
    uint64_t &bar = *reinterpret_cast<uint64_t *>(smem_raw + SMF_BAR);

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int64_t x0 = (int64_t)tx * T, y0 = (int64_t)ty * T;
    if (tid == 0) {
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, T * T);
        ht::tma_load_2d(pk, &map, (int)x0, (int)y0 + p.halo_top, &bar);
    }
    if (tid < 16)
        lut[tid] = ((unsigned)link_offset_pitch(tid, CP) & 0xFFu) << 24;
    const int th = (int)min((int64_t)T, p.rows - y0), tw = (int)min((int64_t)T, p.cols - x0);
    // thread -> 4 consecutive cells of a row, 4 passes of 16 rows (as in stage A)
    const int q4 = 4 * (tid & 15), r0 = tid >> 4;
    const int cbase = r0 * CP + q4; // in the count arrays
    const bool wide = q4 + 3 < tw;

    // tile-local counts and this tile's resolved inflow while the TMA is in flight
    uint2 lc[4];
    {
        const uint16_t *src = p.lacc + (y0 + r0) * p.lacc_ld + x0 + q4;
        const int64_t step = 16 * p.lacc_ld;
#pragma unroll
        for (int i = 0; i < 4; i++, src += step) {
            lc[i] = make_uint2(0u, 0u);
            if (r0 + 16 * i >= th || q4 >= tw)
                continue;
            if (wide)
                lc[i] = *reinterpret_cast<const uint2 *>(src);
            else {
                unsigned h[4] = {0u, 0u, 0u, 0u};
                for (int j = 0; j < 4 && q4 + j < tw; j++)
                    h[j] = src[j];
                lc[i] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
            }
        }
    }
    int ely = 0, elx = 0;
    unsigned long long inflow = 0;
    if (perimeter_cell(tid, th, tw, ely, elx))
        inflow = reinterpret_cast<const unsigned long long *>(p.inflow)[(int64_t)tile * SLOTS + tid];
    __syncthreads();
    ht::mbar_wait(&bar, 0);

    // ---- cell words
    {
        unsigned cm[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            cm[j] = (q4 + j == 0 ? EX_W : 0u) | (q4 + j == tw - 1 ? EX_E : 0u);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int ly = r0 + 16 * i;
            const unsigned wq = *reinterpret_cast<const unsigned *>(pk + ly * T + q4);
            const unsigned rm = (ly == 0 ? EX_N : 0u) | (ly == th - 1 ? EX_S : 0u);
            const unsigned cnt[4] = {lc[i].x & 0xFFFFu, lc[i].x >> 16, lc[i].y & 0xFFFFu, lc[i].y >> 16};
            unsigned wd[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned b = (wq >> (8 * j)) & 0xFFu;
                const unsigned code = b & HT_DIR_CODE;
                const bool link = !(b & FLAGS_STOP) && ly < th && q4 + j < tw &&
                                  !(((rm | cm[j]) >> code) & 1u);
                wd[j] = (link ? lut[code] : 0u) | cnt[j];
            }
            *reinterpret_cast<uint4 *>(lo32 + cbase + i * 16 * CP) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            *reinterpret_cast<uint4 *>(hi32 + cbase + i * 16 * CP) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    __syncthreads();

    // ---- walkers: one per entry cell with inflow
    if (inflow) {
        const unsigned i_lo = (unsigned)(inflow & 0xFFFFu), i_hi = (unsigned)(inflow >> 16);
        // The link (top byte of the low word) never changes: it is read with a plain byte
        // load, and the adds are fire-and-forget REDs -- the walk's critical path is one
        // LDS per step instead of a returning atomic.
        unsigned addr = ht::smem_u32(lo32) + 4u * (unsigned)(ely * CP + elx);
        // a path visits a cell once; the bound only matters for a cyclic (corrupt)
        // direction raster, which must not hang the GPU
        for (int step = 0; step < T * T; step++) {
            int off;
            asm volatile("ld.shared.s8 %0, [%1 + 3];" : "=r"(off) : "r"(addr));
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(i_lo) : "memory");
            if (i_hi)
                asm volatile("red.shared.add.u32 [%0 + %2], %1;" ::"r"(addr), "r"(i_hi), "n"(SMF_HI - SMF_LO) : "memory");
            if (off == 0)
                break;
            addr += 4u * (unsigned)off;
        }
    }
    __syncthreads();

    // ---- outputs: 4 cells per thread and pass
    double *accp = p.acc ? p.acc + (y0 + r0) * p.acc_ld + x0 + q4 : nullptr;
    int8_t *dirp = (WMODE == W_ONE && p.dir_out) ? p.dir_out + (y0 + r0) * p.dir_ld + x0 + q4 : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ly = r0 + 16 * i;
        if (ly >= th || q4 >= tw)
            break;
        const uint4 l4 = *reinterpret_cast<const uint4 *>(lo32 + cbase + i * 16 * CP);
        const uint4 h4 = *reinterpret_cast<const uint4 *>(hi32 + cbase + i * 16 * CP);
        const unsigned wq = *reinterpret_cast<const unsigned *>(pk + ly * T + q4);
        const unsigned lo[4] = {l4.x, l4.y, l4.z, l4.w}, hi[4] = {h4.x, h4.y, h4.z, h4.w};
        double a[4];
        unsigned dout = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned b = (wq >> (8 * j)) & 0xFFu;
            const bool null = b & HT_DIR_NULL;
            const unsigned long long cnt = ((unsigned long long)hi[j] << 16) + (lo[j] & 0x00FFFFFFu);
            a[j] = null ? __longlong_as_double(0x7ff8000000000000LL) : (double)cnt;
            if (dirp) {
                const int code = b & HT_DIR_CODE;
                int out;
                if (null)
                    out = -128;
                else if (b & HT_DIR_NEG)
                    out = -code;
                else if ((b & HT_DIR_EDGE) && cnt >= 60ull) {
                    // rule R5: an edge cell that is a swale (|acc| >= 60) gets its outward
                    // code back: first neighbour outside the raster / NULL in the order
                    // S,N,W,E,NE,SW,SE,NW (rare: read the neighbours from global memory)
                    out = 0;
#pragma unroll
                    for (int ct = 7; ct >= 0; ct--) {
                        const int kk = ct2win(ct);
                        const int64_t lr = y0 + ly + kk / 3 - 1, gc = x0 + q4 + j + kk % 3 - 1;
                        const int64_t gr = p.row0 + lr;
                        bool outside = gr < 0 || gr >= p.total_rows || gc < 0 || gc >= p.cols;
                        if (!outside)
                            outside = p.packed[lr * p.packed_ld + gc] & HT_DIR_NULL;
                        if (outside)
                            out = -win2code(kk);
                    }
                    if (out == 0)
                        out = code; // not an edge after all (foreign packed grid)
                }
                else
                    out = code;
                dout |= ((unsigned)out & 0xFFu) << (8 * j);
            }
        }
        if (wide) {
            if (accp) {
                double2 *d2 = reinterpret_cast<double2 *>(accp + (int64_t)(16 * i) * p.acc_ld);
                d2[0] = make_double2(a[0], a[1]);
                d2[1] = make_double2(a[2], a[3]);
            }
            if (dirp)
                *reinterpret_cast<unsigned *>(dirp + (int64_t)(16 * i) * p.dir_ld) = dout;
        }
        else {
            for (int j = 0; j < 4 && q4 + j < tw; j++) {
                if (accp)
                    accp[(int64_t)(16 * i) * p.acc_ld + j] = a[j];
                if (dirp)
                    dirp[(int64_t)(16 * i) * p.dir_ld + j] = (int8_t)(dout >> (8 * j));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// stage B: boundary-graph resolve (subtree sums over a forest) by pointer doubling
// ---------------------------------------------------------------------------
// Grid-wide barrier for the cooperative launch below (all CTAs co-resident):
// a monotonic arrival counter; barrier number k releases at (k+1) * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = (++epoch) * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    else
        ++epoch;
    __syncthreads();
}

// Subtree sums over a forest, nodes renumbered densely (node id = position in
// `list`; identity when there is no list) so that a round touches, per node,
// two coalesced streams of its own state plus TWO random accesses: the pointer of
// the node 2^k steps downstream and the atomic that hands the sum on.
//   round k: node i hands a[i] (sum over the 2^k nodes upstream of it, itself
//   included) to t = ptr[i] and doubles its pointer, ptr'[i] = ptr[t].
// Sums handed over in round k land in arr[k & 1]; the owner folds them into a[i]
// at the start of round k + 1 (nobody writes that buffer then) -- one grid
// barrier per round.  A node whose pointer has run off the end of its path is
// done; what still arrives for it stays in arr[] and is added at the very end.
// band-edge marks ride in bits 40.. of the u64 sums: counts stay below 2^40 (rasters of up to
// 10^12 cells), and the 24-bit mark field sums up to 16.7 M marked slots upstream of a node
// (a band has 2 * cols of them) without wrapping to zero
constexpr int MARK_SHIFT = 40;
__device__ __forceinline__ unsigned long long masked_plus(unsigned long long old, unsigned long long add)
{
    return (old & ((1ull << MARK_SHIFT) - 1ull)) + add;
}
__device__ __forceinline__ double masked_plus(double old, double add) { return old + add; }
__device__ __forceinline__ i128 masked_plus(const i128 &old, const i128 &add)
{
    return i128(old.lo & ~((1ull << 24) - 1ull), old.hi) + add; // W_MARK_BITS
}
__device__ __forceinline__ bool is_marked(unsigned long long v) { return (v >> MARK_SHIFT) != 0ull; }
__device__ __forceinline__ bool is_marked(const i128 &v) { return (v.lo & ((1ull << 24) - 1ull)) != 0ull; }

// boundary-graph slot of column (tx, lx) of the band's first (side 0) / last (side 1) row
__device__ __forceinline__ int32_t band_edge_slot_of(int side, int tx, int lx, int tiles_x, int tiles_y, int last_h)
{
    if (!side)
        return tx * SLOTS + lx;
    return ((tiles_y - 1) * tiles_x + tx) * SLOTS + (last_h == 1 ? lx : T + lx);
}
__device__ __forceinline__ bool is_band_edge_slot(int32_t u, int tiles_x, int tiles_y, int last_h)
{
    const int tile = u / SLOTS, k = u % SLOTS;
    if (tile >= tiles_x * tiles_y)
        return false; // a virtual slot (a neighbour band's cell)
    const int ty = tile / tiles_x;
    return (ty == 0 && k < T) || (ty == tiles_y - 1 && (last_h == 1 ? k < T : (k >= T && k < 2 * T)));
}

template <typename V>
__global__ void __launch_bounds__(256, 4)
coarse_resolve_kernel(const V *base0, const int32_t *next0, int64_t n, const int32_t *list,
                      const int *nlist, int32_t *idmap, int32_t *ptrA, int32_t *ptrB, V *a,
                      V *arr0, V *arr1, int32_t *rootA, int32_t *rootB, V *agg_out,
                      int32_t *root_out, int *flags, int *rounds_out, int add_to_masked,
                      int edge_tiles_x, int edge_tiles_y, int edge_last_h)
{
    unsigned *gbar = reinterpret_cast<unsigned *>(flags + 96); // zeroed with flags
    unsigned epoch = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // only the slots in `list` are nodes (all n slots if there is no list)
    const int64_t na = list ? (int64_t)*nlist : n;
    if (list) {
        for (int64_t i = t0; i < na; i += stride)
            idmap[list[i]] = (int32_t)i;
        // slots that are not nodes end at themselves.  With an edge filter (row bands) only
        // the slots of the band's first / last row are ever asked for
        if (root_out && edge_tiles_x > 0) {
            for (int64_t c = t0; c < 2ll * edge_tiles_x * T; c += stride) {
                const int side = c >= (int64_t)edge_tiles_x * T, cc = (int)(c - (int64_t)side * edge_tiles_x * T);
                const int32_t s = band_edge_slot_of(side, cc / T, cc % T, edge_tiles_x, edge_tiles_y, edge_last_h);
                root_out[s] = s;
            }
        }
        else if (root_out)
            for (int64_t s = t0; s < n; s += stride)
                root_out[s] = (int32_t)s;
        grid_barrier(gbar, epoch);
    }
    for (int64_t i = t0; i < na; i += stride) {
        const int64_t u = list ? list[i] : i;
        const int32_t nx = next0[u];
        ptrA[i] = nx < 0 ? -1 : (list ? idmap[nx] : nx);
        a[i] = base0[u];
        arr0[i] = (V)0;
        arr1[i] = (V)0;
        if (rootA)
            rootA[i] = (int32_t)i;
    }
    grid_barrier(gbar, epoch);

    int round = 0;
    for (;; round++) {
        V *cur = (round & 1) ? arr1 : arr0, *prev = (round & 1) ? arr0 : arr1;
        bool any = false;
        // independent nodes in flight per thread (8 at 3 CTAs/SM spills and is slower:
        // 0.37 ms against 0.32)
        constexpr int U = 4;
        for (int64_t i0 = t0; i0 < na; i0 += stride * U) {
            int32_t t[U], pt[U], rr[U];
            V v[U], d[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int64_t i = i0 + j * stride;
                t[j] = i < na ? ptrA[i] : -2; // -2: no such node
            }
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int64_t i = i0 + j * stride;
                pt[j] = t[j] >= 0 ? ptrA[t[j]] : -1;
                v[j] = t[j] >= 0 ? a[i] : (V)0;
                d[j] = t[j] >= 0 ? prev[i] : (V)0;
                // the last node of the path, written ONCE: in the round in which the node
                // finishes (its target t has no pointer left, so t finished in an earlier
                // round and root[t] is final), not carried through every round
                if (rootA)
                    rr[j] = (t[j] >= 0 && pt[j] < 0) ? rootA[t[j]] : -1;
            }
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int64_t i = i0 + j * stride;
                if (t[j] == -2)
                    continue;
                if (t[j] >= 0) {
                    if (d[j] != (V)0) { // what arrived in the last round
                        v[j] += d[j];
                        a[i] = v[j];
                        prev[i] = (V)0;
                    }
                    if (v[j] != (V)0)
                        atomic_add_v(&cur[t[j]], v[j]);
                    any = true;
                }
                ptrB[i] = pt[j];
                if (rootA && rr[j] >= 0)
                    rootA[i] = rr[j];
            }
        }
        if (any)
            flags[round] = 1;
        grid_barrier(gbar, epoch);
        int32_t *tp = ptrA; ptrA = ptrB; ptrB = tp;
        if (!flags[round] || round == 62)
            break;
    }
    // totals (and the last node of every path) back to slot order
    for (int64_t i = t0; i < na; i += stride) {
        const int64_t u = list ? list[i] : i;
        const V total = a[i] + arr0[i] + arr1[i];
        // delta resolve: the totals are a correction on top of an earlier result whose
        // upper 24 bits carried band-edge marks (u64 values only)
        agg_out[u] = add_to_masked ? masked_plus(agg_out[u], total) : total;
        if (rootA && root_out &&
            (edge_tiles_x <= 0 || is_band_edge_slot((int32_t)u, edge_tiles_x, edge_tiles_y, edge_last_h)))
            root_out[u] = list ? list[rootA[i]] : rootA[i];
    }
    if (t0 == 0 && rounds_out)
        *rounds_out = round + 1;
}

// ---------------------------------------------------------------------------
// stage B in TWO LEVELS (count modes).  The flat resolve above touches every node of the
// boundary graph in every round, from L2 / DRAM: 2.2 M nodes x 10 rounds for 10 000^2.
// Super-tiles of SG x SG tiles cut the graph: links between tiles of one super-tile are
// INTERNAL, the others EXTERNAL, a node with an external in-link is OUTER.
//   K1  super_tile_kernel<1>: one CTA per super-tile resolves the internal forest in SHARED
//       memory (the rounds of stage A on <= 16 K nodes): S1 = subtree sums over internal
//       links, root1 = last node of the internal path.  A root with an external link hands
//       S1 to its target (base2) and flags it OUTER; the list of outer nodes gets next2[w] =
//       next0[root1(w)].
//   K2  the flat resolve on the OUTER nodes only (~8x fewer): X = what arrives from other
//       super-tiles, in total.
//   K3  super_tile_kernel<3>: every OUTER node adds its X along its internal path (a walk of
//       a few nodes): inflow = S1 + the X of the outer nodes upstream.
// Inside a super-tile a sum of genuine tile-to-tile flows is below 2^19 cells: K1 works on
// 32-bit words (native shared atomics); nodes with a large own value or a band-edge mark
// are promoted to the X side.  K3 adds 64-bit values as two words with the carry handed
// on.  Exact, so the result equals the flat resolve bit for bit.
// ---------------------------------------------------------------------------
constexpr int SGX = 8, SGY = 4;             // tiles per super-tile (columns, rows)
constexpr int SUP_NT = SGX * SGY;
constexpr int SUP_CPT = 16;                 // nodes per thread
constexpr int SUP_MAXN = SUP_NT * SLOTS;    // every slot of every tile a node: always fits
constexpr int SUP_THREADS = SUP_MAXN / SUP_CPT;
constexpr int S1_BIG_BITS = 20;             // own values from 2^20 on (and marked ones) go the X way
static_assert(SUP_NT <= 64 && SUP_THREADS <= 1024, "super-tile size");

struct SupArgs {
    int tiles_x, tiles_y, sup_x;
    int32_t ntile_slots;
    const int32_t *list, *next0, *tile_first, *tile_cnt;
    int32_t *idmap;
    const unsigned long long *base0;
    uint8_t *outer;
    unsigned long long *base2;
    int32_t *rnode;              // per node (list position): slot of the last node of its internal path
    uint32_t *s1w;               // per node (list position): S1
    unsigned long long *agg;     // K3: X of the outer nodes in, inflow of every node out
    int32_t *root1;              // K1, row bands: last node of the internal path, band-edge slots only
    int edge_tiles_x, edge_tiles_y, edge_last_h;
};

constexpr int SUP_SM_PTR = 0;                        // K1 pointer entries / K3 next node
constexpr int SUP_SM_V0 = SUP_SM_PTR + SUP_MAXN * 4; // K1 packed sums / K3 low words
constexpr int SUP_SM_V1 = SUP_SM_V0 + SUP_MAXN * 4;  // K3 high words
constexpr int SUP_SM_TAB = SUP_SM_V1 + SUP_MAXN * 4; // int[SUP_NT + 1] prefix, int[SUP_NT] first
constexpr int SUP_SM_TOTAL = SUP_SM_TAB + (2 * SUP_NT + 2) * 4;

// K1's rounds: cnt_hand_on / cnt_publish with this kernel's delta between entry and sum
__device__ __forceinline__ void sup_hand_on(unsigned &low, unsigned a)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, f;\n"
        "and.b32 f, %0, 1;\n"
        "setp.ne.u32 p, f, 0;\n"
        "and.b32 t, %0, 0xFFFFFFFC;\n"
        "@p red.shared.add.u32 [t + %2], %1;\n"
        "@p ld.shared.u32 %0, [t];\n"
        "}\n"
        : "+r"(low)
        : "r"(a), "n"(SUP_SM_V0 - SUP_SM_PTR)
        : "memory");
}
__device__ __forceinline__ unsigned sup_publish(unsigned own_addr, unsigned low)
{
    unsigned a;
    asm volatile(
        "st.shared.u32 [%1], %2;\n"
        "ld.shared.u32 %0, [%1 + %3];\n"
        : "=r"(a)
        : "r"(own_addr), "r"(low), "n"(SUP_SM_V0 - SUP_SM_PTR)
        : "memory");
    return a;
}

template <int PHASE>
__global__ void __launch_bounds__(SUP_THREADS, 65536 / (SUP_THREADS * 64))
super_tile_kernel(const SupArgs p)
{
    extern __shared__ __align__(16) unsigned char sup_sm[];
    uint32_t *sptr = reinterpret_cast<uint32_t *>(sup_sm + SUP_SM_PTR);
    uint32_t *sv0 = reinterpret_cast<uint32_t *>(sup_sm + SUP_SM_V0);
    uint32_t *sv1 = reinterpret_cast<uint32_t *>(sup_sm + SUP_SM_V1);
    int *tpre = reinterpret_cast<int *>(sup_sm + SUP_SM_TAB); // [SUP_NT + 1]
    int *tfirst = tpre + SUP_NT + 1;                           // [SUP_NT]
    const int tid = threadIdx.x;
    const int gx = blockIdx.x % p.sup_x, gy = blockIdx.x / p.sup_x;

    // ---- the super-tile's tiles and their entries in the node list
    if (tid < SUP_NT) {
        const int ty = gy * SGY + tid / SGX, tx = gx * SGX + tid % SGX;
        int cnt = 0, first = 0;
        if (ty < p.tiles_y && tx < p.tiles_x) {
            const int t = ty * p.tiles_x + tx;
            cnt = p.tile_cnt[t];
            first = p.tile_first[t];
        }
        tfirst[tid] = first;
        tpre[tid + 1] = cnt; // turned into a prefix below
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        tpre[0] = 0;
        for (int j = 1; j <= SUP_NT; j++) {
            run += tpre[j];
            tpre[j] = run;
        }
    }
    __syncthreads();
    const int nG = tpre[SUP_NT];
    if (nG == 0)
        return;
    // local index -> position in the node list (per-node arrays are indexed by it: the nodes
    // of a tile are neighbours there, while their slots are spread over the tile's 256)
    auto node_of_index = [&](int i) -> int {
        int lo = 0, hi = SUP_NT; // largest j with tpre[j] <= i
#pragma unroll
        for (int it = 0; it < 6; it++) { // 2^6 >= SUP_NT
            if (hi - lo <= 1)
                break;
            const int mid = (lo + hi) >> 1;
            if (tpre[mid] <= i)
                lo = mid;
            else
                hi = mid;
        }
        return tfirst[lo] + (i - tpre[lo]);
    };
    auto internal = [&](int32_t nx) -> bool {
        if (nx < 0 || nx >= p.ntile_slots)
            return false;
        const int tile = nx / SLOTS, ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        return ty / SGY == gy && tx / SGX == gx;
    };

    // Global loads are issued in BATCHES of independent, unconditional loads (an index that is
    // out of range is clamped to node 0 and its result ignored): with a branch around every
    // load the 16 cells of a thread paid 3 - 4 dependent round trips to L2 / DRAM EACH.  The
    // loops stop at the super-tile's node count (typically a third of the capacity).
    constexpr int GB = 4; // cells per batch
    int32_t slot[SUP_CPT];
#pragma unroll
    for (int k = 0; k < SUP_CPT; k++) {
        if (k * SUP_THREADS >= nG) { // uniform
            slot[k] = -1;
            continue;
        }
        const int i = tid + k * SUP_THREADS;
        slot[k] = p.list[node_of_index(i < nG ? i : 0)];
    }
#pragma unroll
    for (int k = 0; k < SUP_CPT; k++) {
        const int i = tid + k * SUP_THREADS;
        if (i < nG) {
            if (PHASE == 1) {
                p.idmap[slot[k]] = i;
                sv1[i] = (uint32_t)slot[k]; // K1 keeps the slots in shared memory (the path ends look them up)
            }
        } else
            slot[k] = -1;
    }
    __syncthreads(); // idmap of this super-tile's nodes is read by this CTA only
    const int32_t safe_slot = p.list[node_of_index(0)]; // node 0 exists (nG > 0): a valid slot to clamp to

    const unsigned pbase = ht::smem_u32(sptr), own0 = pbase + 4u * (unsigned)tid;
    if (PHASE == 1) {
        unsigned low[SUP_CPT], a[SUP_CPT];
#pragma unroll
        for (int k = 0; k < SUP_CPT; k++) {
            low[k] = own0 + (unsigned)(k * SUP_THREADS * 4);
            a[k] = 0u;
        }
#pragma unroll
        for (int k0 = 0; k0 < SUP_CPT; k0 += GB) {
            if (k0 * SUP_THREADS >= nG)
                break; // uniform
            int32_t nx[GB], li[GB];
            unsigned long long bv[GB];
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const int32_t sl = slot[k0 + j] >= 0 ? slot[k0 + j] : safe_slot;
                nx[j] = p.next0[sl];
                bv[j] = p.base0[sl];
            }
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const bool in = slot[k0 + j] >= 0 && internal(nx[j]);
                li[j] = p.idmap[in ? nx[j] : safe_slot];
                if (!in)
                    li[j] = -1;
            }
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const int k = k0 + j;
                if (slot[k] < 0)
                    continue;
                if (li[j] >= 0)
                    low[k] = (pbase + 4u * (unsigned)li[j]) | J_ACTIVE;
                // a node whose own value is large (flow imported from a neighbour band) or
                // carries a band-edge mark is PROMOTED: its value joins what arrives from
                // outside (base2 -> X, exact by linearity), so that K1's sums stay plain
                // 32-bit counts: <= 2^19 cells of the super-tile and its ring, + 2^10 band-edge
                // nodes x 2^20
                if (bv[j] >= (1ull << S1_BIG_BITS)) {
                    atomicAdd(p.base2 + slot[k], bv[j]);
                    p.outer[slot[k]] = 1;
                } else
                    a[k] = (unsigned)bv[j];
            }
        }
#pragma unroll
        for (int k = 0; k < SUP_CPT; k++) {
            const int i = tid + k * SUP_THREADS;
            sptr[i] = low[k];
            sv0[i] = a[k];
        }
        __syncthreads();
        for (int round = 0; round < 16; round++) {
#pragma unroll
            for (int k = 0; k < SUP_CPT; k++) {
                if (k * SUP_THREADS >= nG)
                    break; // uniform
                sup_hand_on(low[k], a[k]);
            }
            __syncthreads();
            unsigned any = 0;
#pragma unroll
            for (int k = 0; k < SUP_CPT; k++) {
                if (k * SUP_THREADS >= nG)
                    break; // uniform
                a[k] = sup_publish(own0 + (unsigned)(k * SUP_THREADS * 4), low[k]);
                any |= low[k];
            }
            if (!__syncthreads_or((int)(any & J_ACTIVE)))
                break;
        }
        // a[] = S1, low[] = the root of the internal path
#pragma unroll
        for (int k0 = 0; k0 < SUP_CPT; k0 += GB) {
            if (k0 * SUP_THREADS >= nG)
                break; // uniform
            int32_t nxr[GB];
#pragma unroll
            for (int j = 0; j < GB; j++) // own link: needed by the path ends only, loaded for all (L2 hits)
                nxr[j] = p.next0[slot[k0 + j] >= 0 ? slot[k0 + j] : safe_slot];
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const int k = k0 + j, i = tid + k * SUP_THREADS;
                if (slot[k] < 0)
                    continue;
                const int gi = node_of_index(i);
                const int ri = (int)(((low[k] & ~3u) - pbase) >> 2);
                const int32_t rslot = (int32_t)sv1[ri];
                p.s1w[gi] = a[k];
                p.rnode[gi] = rslot;
                if (ri == i && nxr[j] >= 0) {
                    // the end of an internal path with a link out of the super-tile hands S1 on
                    if (a[k])
                        atomicAdd(p.base2 + nxr[j], (unsigned long long)a[k]);
                    p.outer[nxr[j]] = 1;
                }
                if (p.root1 && is_band_edge_slot(slot[k], p.edge_tiles_x, p.edge_tiles_y, p.edge_last_h))
                    p.root1[slot[k]] = rslot;
            }
        }
    } else {
        // ---- K3: one-hop internal links, S1 as 64-bit words, the outer nodes walk their X down
        int32_t *snext = reinterpret_cast<int32_t *>(sptr);
        unsigned long long x[SUP_CPT];
#pragma unroll
        for (int k = 0; k < SUP_CPT; k++)
            x[k] = 0ull;
#pragma unroll
        for (int k0 = 0; k0 < SUP_CPT; k0 += GB) {
            if (k0 * SUP_THREADS >= nG)
                break; // uniform
            int32_t nx[GB], li[GB];
            unsigned s1[GB];
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const int i = tid + (k0 + j) * SUP_THREADS;
                const int32_t sl = slot[k0 + j] >= 0 ? slot[k0 + j] : safe_slot;
                nx[j] = p.next0[sl];
                x[k0 + j] = p.agg[sl]; // X of an outer node; 0 for every other node
                s1[j] = p.s1w[node_of_index(i < nG ? i : 0)];
            }
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const bool in = slot[k0 + j] >= 0 && internal(nx[j]);
                li[j] = p.idmap[in ? nx[j] : safe_slot];
                if (!in)
                    li[j] = -1;
            }
#pragma unroll
            for (int j = 0; j < GB; j++) {
                const int i = tid + (k0 + j) * SUP_THREADS;
                if (slot[k0 + j] < 0) {
                    x[k0 + j] = 0ull;
                    continue;
                }
                snext[i] = li[j];
                sv0[i] = s1[j];
                sv1[i] = 0u;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SUP_CPT; k++) {
            if (k * SUP_THREADS >= nG)
                break; // uniform
            if (!x[k])
                continue;
            for (int cur = tid + k * SUP_THREADS, step = 0; cur >= 0 && step < SUP_MAXN; cur = snext[cur], step++)
                smem_add64(sv0 + cur, sv1 + cur, (long long)x[k]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SUP_CPT; k++) {
            const int i = tid + k * SUP_THREADS;
            if (i < nG)
                p.agg[slot[k]] = ((unsigned long long)sv1[i] << 32) | sv0[i];
        }
    }
}

// the cells of the neighbour bands (virtual slots) belong to no tile: stage A adds what the
// band's edge rows hand them to their base0 directly; they are level-2 nodes with that value
__global__ void virtual_nodes_kernel(const unsigned long long *base0, unsigned long long *base2,
                                     uint8_t *outer, int64_t first, int64_t count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    const unsigned long long b = base0[first + i];
    if (b) {
        base2[first + i] += b;
        outer[first + i] = 1;
    }
}

// the OUTER nodes of the node list, for the level-2 resolve, and their level-2 links:
// next2[w] = where the internal path that starts at w leaves its super-tile
__global__ void outer_nodes_kernel(const uint8_t *outer, const int32_t *list, const int *nlist,
                                   const int32_t *rnode, const int32_t *next0, int32_t ntile_slots,
                                   int32_t *next2, int32_t *list2, int *nlist2)
{
    const int na = *nlist;
    const int lane = threadIdx.x & 31;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < na; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        const int32_t u = i < na ? list[i] : 0;
        const bool in = i < na && outer[u];
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (!m)
            continue;
        int at = 0;
        if (lane == 0)
            at = atomicAdd(nlist2, __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        if (in) {
            list2[at] = u;
            next2[u] = u < ntile_slots ? next0[rnode[i]] : -1; // a neighbour band's cell ends the path
        }
    }
}

// row bands: the last node of the path of every slot of the band's first / last row --
// root1 (K1) if the internal path ends inside its super-tile, else the level-2 root of the
// outer node that path hands over to
__global__ void edge_roots_kernel(int32_t *root_out, const int32_t *root1, const int32_t *next0,
                                  const int32_t *rootX, int tiles_x, int tiles_y, int last_h)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2ll * tiles_x * T)
        return;
    const int side = c >= (int64_t)tiles_x * T, cc = (int)(c - (int64_t)side * tiles_x * T);
    const int32_t s = band_edge_slot_of(side, cc / T, cc % T, tiles_x, tiles_y, last_h);
    const int32_t r1 = root1[s];                    // -1: not a node
    const int32_t n2 = r1 >= 0 ? next0[r1] : -1;    // where its internal path leaves the super-tile
    root_out[s] = n2 >= 0 ? rootX[n2] : (r1 >= 0 ? r1 : s);
}

__global__ void append_range_kernel(int32_t *list, int *nlist, int32_t first, int32_t count)
{
    __shared__ int base;
    if (threadIdx.x == 0)
        base = atomicAdd(nlist, count);
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += blockDim.x)
        list[base + i] = first + i;
}

// Row-band sharding, delta resolve: the nodes whose resolved sum carries a mark lie
// on or downstream of a band-edge cell, i.e. they are the only ones whose inflow
// changes when the neighbour bands' deliveries come in.  They are closed under
// next[]: a forest of their own, resolved with base = the imported flow.
template <typename V>
__global__ void marked_nodes_kernel(const V *agg, const int32_t *list,
                                    const int *nlist, int32_t *list2, int *nlist2,
                                    V *base2, const V *imp_top,
                                    const V *imp_bot, int64_t rows, int64_t cols,
                                    int tiles_x, int tiles_y)
{
    const int na = *nlist;
    const int lane = threadIdx.x & 31;
    const int last_h = (int)(rows - (int64_t)(tiles_y - 1) * T);
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < na; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        const int32_t u = i < na ? list[i] : 0;
        const bool in = i < na && is_marked(agg[u]);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (!m)
            continue;
        int at = 0;
        if (lane == 0)
            at = atomicAdd(nlist2, __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        if (in) {
            list2[at] = u;
            V b(0);
            const int tile = u / SLOTS, k = u % SLOTS;
            if (tile < tiles_x * tiles_y) { // not a virtual slot
                const int ty = tile / tiles_x, tx = tile % tiles_x;
                const int64_t col = (int64_t)tx * T + (k & (T - 1));
                if (imp_top && ty == 0 && k < T && col < cols)
                    b = imp_top[col];
                else if (imp_bot && ty == tiles_y - 1 && col < cols &&
                         (last_h == 1 ? k < T : (k >= T && k < 2 * T)))
                    b = imp_bot[col];
            }
            base2[u] = b;
        }
    }
}

// Row-band sharding, count modes: the two tables a band publishes after its first
// resolve, in one launch.  out[0 / 1][c] = flow delivered into column c of the edge
// row of the band above / below (marks cleared); out[2 / 3][c] = for the band's own
// first / last row cell c, the foreign cell the path entering there finally
// reaches (side * cols + column, side 0 = band above), or -1.
// Sharded runs: a tile whose entry slots carry no mark receives nothing that depends on the
// neighbour bands, so its stage C can run while the band tables are being exchanged.
template <typename V>
__global__ void tile_phase_kernel(const V *agg, const int32_t *list, const int *nlist, int ntile_slots,
                                  uint8_t *tile_phase)
{
    const int na = *nlist;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) {
        const int32_t u = list[i];
        if (u < ntile_slots && is_marked(agg[u]))
            tile_phase[u / SLOTS] = 1; // every writer stores the same byte
    }
}

__global__ void band_tables_kernel(const unsigned long long *agg, const int32_t *root, int64_t rows,
                                   int64_t cols, int tiles_x, int tiles_y, int64_t vfirst,
                                   long long *out)
{
    const int last_ty = tiles_y - 1;
    const int last_h = (int)(rows - (int64_t)last_ty * T);
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long mask = (1ull << MARK_SHIFT) - 1ull;
        out[c] = (long long)(agg[vfirst + c] & mask);
        out[cols + c] = (long long)(agg[vfirst + cols + c] & mask);
        const int64_t tx = c / T, lx = c % T;
        const int64_t top = tx * SLOTS + lx;
        const int64_t bot = ((int64_t)last_ty * tiles_x + tx) * SLOTS + (last_h == 1 ? lx : T + lx);
        const long long rt = (long long)root[top] - vfirst, rb = (long long)root[bot] - vfirst;
        out[2 * cols + c] = rt >= 0 ? rt : -1;
        out[3 * cols + c] = rb >= 0 ? rb : -1;
    }
}

// Band-level forest from the all-gathered tables (every[b][0..3][c] as written by
// band_tables_kernel): node (b, s, c) = cell c of band b's first (s = 0) / last
// (s = 1) row; base = what the neighbour band delivers into it, next = the node its
// path reaches in a neighbour band, or -1.
__global__ void band_graph_kernel(const long long *every, int P, int64_t cols,
                                  unsigned long long *base, int32_t *next)
{
    const int64_t n = (int64_t)P * 2 * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / (2 * cols));
        const int s = (int)((i / cols) & 1);
        const int64_t c = i % cols;
        unsigned long long v = 0ull;
        if (s == 0 && b > 0)
            v = (unsigned long long)every[((int64_t)(b - 1) * 4 + 1) * cols + c]; // sent down by the band above
        if (s == 1 && b < P - 1)
            v = (unsigned long long)every[((int64_t)(b + 1) * 4 + 0) * cols + c]; // sent up by the band below
        base[i] = v;
        const long long L = every[((int64_t)b * 4 + 2 + s) * cols + c];
        int32_t nx = -1;
        if (L >= 0) {
            const long long side = L / cols, c2 = L - side * cols;
            nx = side == 0 ? (int32_t)(((int64_t)(b - 1) * 2 + 1) * cols + c2)  // last row of the band above
                           : (int32_t)(((int64_t)(b + 1) * 2 + 0) * cols + c2); // first row of the band below
        }
        next[i] = nx;
    }
}

// weighted mode: the same two tables with 128-bit deliveries:
// out[0 / 1][c] = low / high word of the flow delivered into column c of the band above,
// out[2 / 3][c] = the same for the band below, out[4 / 5][c] = links as in band_tables_kernel
__global__ void band_tables_w_kernel(const i128 *agg, const int32_t *root, int64_t rows, int64_t cols,
                                     int tiles_x, int tiles_y, int64_t vfirst, long long *out)
{
    const int last_ty = tiles_y - 1;
    const int last_h = (int)(rows - (int64_t)last_ty * T);
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
        i128 up = agg[vfirst + c], dn = agg[vfirst + cols + c];
        up.lo &= ~((1ull << W_MARK_BITS) - 1ull); // marks cleared
        dn.lo &= ~((1ull << W_MARK_BITS) - 1ull);
        out[c] = (long long)up.lo;
        out[cols + c] = up.hi;
        out[2 * cols + c] = (long long)dn.lo;
        out[3 * cols + c] = dn.hi;
        const int64_t tx = c / T, lx = c % T;
        const int64_t top = tx * SLOTS + lx;
        const int64_t bot = ((int64_t)last_ty * tiles_x + tx) * SLOTS + (last_h == 1 ? lx : T + lx);
        const long long rt = (long long)root[top] - vfirst, rb = (long long)root[bot] - vfirst;
        out[4 * cols + c] = rt >= 0 ? rt : -1;
        out[5 * cols + c] = rb >= 0 ? rb : -1;
    }
}

__global__ void band_graph_w_kernel(const long long *every, int P, int64_t cols, i128 *base,
                                    int32_t *next)
{
    const int64_t n = (int64_t)P * 2 * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / (2 * cols));
        const int s = (int)((i / cols) & 1);
        const int64_t c = i % cols;
        i128 v(0);
        if (s == 0 && b > 0) { // sent down by the band above
            const long long *e = every + ((int64_t)(b - 1) * 6 + 2) * cols;
            v = i128((unsigned long long)e[c], e[cols + c]);
        }
        if (s == 1 && b < P - 1) { // sent up by the band below
            const long long *e = every + ((int64_t)(b + 1) * 6 + 0) * cols;
            v = i128((unsigned long long)e[c], e[cols + c]);
        }
        base[i] = v;
        const long long L = every[((int64_t)b * 6 + 4 + s) * cols + c];
        int32_t nx = -1;
        if (L >= 0) {
            const long long side = L / cols, c2 = L - side * cols;
            nx = side == 0 ? (int32_t)(((int64_t)(b - 1) * 2 + 1) * cols + c2)
                           : (int32_t)(((int64_t)(b + 1) * 2 + 0) * cols + c2);
        }
        next[i] = nx;
    }
}

// boundary-graph slot of column c of the band's first (top) / last row
__device__ __forceinline__ int64_t edge_slot(bool top, int64_t c, int64_t rows, int tiles_x, int tiles_y)
{
    const int64_t tx = c / T, lx = c % T;
    if (top)
        return tx * SLOTS + lx;
    const int last_ty = tiles_y - 1;
    const int last_h = (int)(rows - (int64_t)last_ty * T);
    return ((int64_t)last_ty * tiles_x + tx) * SLOTS + (last_h == 1 ? lx : T + lx);
}

// count modes: tag the slots of the band's first / last row (where flow from a neighbour
// band can come in); the mark rides through the resolve in the upper bits of the sums
// `words` = 64-bit words per value (1: u64 counts, mark 1 << 40; 2: 128-bit weighted sums, mark 1
// in the low word, whose low 24 bits are free)
__global__ void mark_edges_kernel(unsigned long long *base0, int words, unsigned long long mark,
                                  int64_t rows, int64_t cols, int tiles_x, int tiles_y, int top, int bot)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
        // a band of one tile row and height 1 has top == bottom slots; atomics keep both marks
        if (top)
            atomicAdd(base0 + words * edge_slot(true, c, rows, tiles_x, tiles_y), mark);
        if (bot)
            atomicAdd(base0 + words * edge_slot(false, c, rows, tiles_x, tiles_y), mark);
    }
}

// add the flows the neighbour bands deliver into the band's first / last row to the
// boundary graph (before a second full resolve)
template <typename V>
__global__ void add_imports_kernel(V *base0, const V *imp_top, const V *imp_bot, int64_t rows,
                                   int64_t cols, int tiles_x, int tiles_y)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
        if (imp_top)
            atomic_add_v(base0 + edge_slot(true, c, rows, tiles_x, tiles_y), imp_top[c]);
        if (imp_bot)
            atomic_add_v(base0 + edge_slot(false, c, rows, tiles_x, tiles_y), imp_bot[c]);
    }
}

struct Workspace {
    uint16_t *lacc;
    int64_t lacc_ld;
    long long *lacc64;
    int64_t lacc64_ld;
    int32_t *tile_e;
    void *base0, *aggA, *aggB, *arr0, *arr1;
    int32_t *next0, *ptrA, *ptrB, *rootA, *rootB, *list, *idmap, *rootC, *list2;
    void *base2;
    int *nlist2;
    int *flags, *nlist;
    unsigned *gmax_bits;
    // two-level stage B (count modes)
    int32_t *tile_first, *tile_cnt, *next2, *idmapX, *rootX, *root1, *rnode;
    uint32_t *s1w;
    uint8_t *outer;
    int64_t nslots;
    int vsize; // bytes per boundary-graph value: 8 (u64 counts) or 16 (i128 weighted sums)
};

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

int64_t nslots_of(int64_t rows, int64_t cols)
{
    return (int64_t)ht::cdiv(rows, T) * ht::cdiv(cols, T) * SLOTS + 2 * cols;
}

// count modes: u64 values, u16 tile-local counts; weighted: i128 values, int64
// tile-local sums + one int32 per tile
size_t workspace_bytes(int64_t rows, int64_t cols, int mode)
{
    const int64_t n = nslots_of(rows, cols);
    const size_t vs = mode == W_F32 ? 16 : 8;
    size_t local;
    if (mode == W_F32)
        local = align_up((size_t)rows * ((cols + 1) / 2 * 2) * 8) +
                align_up((size_t)ht::cdiv(rows, T) * ht::cdiv(cols, T) * 4);
    else
        local = align_up((size_t)rows * ((cols + 7) / 8 * 8) * 2);
    size_t hier = 0;
    if (mode != W_F32)
        hier = 6 * align_up(n * 4) + align_up(n) +
               2 * align_up((size_t)ht::cdiv(rows, T) * ht::cdiv(cols, T) * 4);
    return 6 * align_up(n * vs) + 9 * align_up(n * 4) + align_up(128 * sizeof(int)) + local + hier + 256;
}

int carve(void *ws, size_t ws_bytes, int64_t rows, int64_t cols, int mode, Workspace &w)
{
    if (!ws || ws_bytes < workspace_bytes(rows, cols, mode))
        return HT_ERR_WORKSPACE;
    const int64_t n = nslots_of(rows, cols);
    const size_t vs = mode == W_F32 ? 16 : 8;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    w.base0 = b; b += align_up(n * vs);
    w.aggA = b; b += align_up(n * vs);
    w.aggB = b; b += align_up(n * vs);
    w.next0 = (int32_t *)b; b += align_up(n * 4);
    w.ptrA = (int32_t *)b; b += align_up(n * 4);
    w.ptrB = (int32_t *)b; b += align_up(n * 4);
    w.rootA = (int32_t *)b; b += align_up(n * 4);
    w.rootB = (int32_t *)b; b += align_up(n * 4);
    w.list = (int32_t *)b; b += align_up(n * 4);
    w.idmap = (int32_t *)b; b += align_up(n * 4);
    w.rootC = (int32_t *)b; b += align_up(n * 4);
    w.list2 = (int32_t *)b; b += align_up(n * 4);
    w.base2 = b; b += align_up(n * vs);
    w.arr0 = b; b += align_up(n * vs);
    w.arr1 = b; b += align_up(n * vs);
    w.flags = (int *)b; b += align_up(128 * sizeof(int));
    w.nlist = w.flags + 64;
    w.nlist2 = w.flags + 65;
    w.gmax_bits = (unsigned *)(w.flags + 100);
    w.lacc = nullptr; w.lacc_ld = 0; w.lacc64 = nullptr; w.lacc64_ld = 0; w.tile_e = nullptr;
    if (mode == W_F32) {
        w.lacc64 = (long long *)b;
        w.lacc64_ld = (cols + 1) / 2 * 2;
        b += align_up((size_t)rows * w.lacc64_ld * 8);
        w.tile_e = (int32_t *)b;
    }
    else {
        w.lacc = (uint16_t *)b;
        w.lacc_ld = (cols + 7) / 8 * 8;
        b += align_up((size_t)rows * w.lacc_ld * 2);
    }
    w.tile_first = w.tile_cnt = w.next2 = w.idmapX = w.rootX = w.root1 = w.rnode = nullptr;
    w.s1w = nullptr; w.outer = nullptr;
    if (mode != W_F32) {
        const size_t nt = (size_t)ht::cdiv(rows, T) * ht::cdiv(cols, T);
        w.next2 = (int32_t *)b; b += align_up(n * 4);
        w.idmapX = (int32_t *)b; b += align_up(n * 4);
        w.rootX = (int32_t *)b; b += align_up(n * 4);
        w.root1 = (int32_t *)b; b += align_up(n * 4);
        w.s1w = (uint32_t *)b; b += align_up(n * 4);
        w.rnode = (int32_t *)b; b += align_up(n * 4);
        w.outer = (uint8_t *)b; b += align_up(n);
        w.tile_first = (int32_t *)b; b += align_up(nt * 4);
        w.tile_cnt = (int32_t *)b; b += align_up(nt * 4);
    }
    w.nslots = n;
    w.vsize = (int)vs;
    return 0;
}

int make_packed_map(CUtensorMap &map, const uint8_t *packed, int64_t packed_ld, int64_t rows,
                    int64_t cols, int halo_top, int halo_bot)
{
    const uint8_t *base = packed - (int64_t)halo_top * packed_ld;
    return ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base,
                           rows + halo_top + halo_bot, cols, packed_ld, RP, RH, false);
}

template <bool FINAL>
int launch_w(const CUtensorMap &map, const AccArgs &p, cudaStream_t st)
{
    if (FINAL) {
        HT_CUDA_TRY(cudaFuncSetAttribute(acc_final_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMX_TOTAL));
        acc_final_w_kernel<<<p.tiles_x * p.tiles_y, THREADS, SMX_TOTAL, st>>>(map, p);
    }
    else {
        HT_CUDA_TRY(cudaFuncSetAttribute(acc_local_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMW_TOTAL));
        acc_local_w_kernel<<<p.tiles_x * p.tiles_y, THREADS, SMW_TOTAL, st>>>(map, p);
    }
    HT_LAUNCH_CHECK();
    return 0;
}

template <int WMODE, bool FINAL>
int launch_cnt(const CUtensorMap &map, const AccArgs &p, cudaStream_t st)
{
    const int ntiles = p.tiles_x * p.tiles_y;
    if (FINAL) {
        auto kern = acc_final_cnt_kernel<WMODE>;
        HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMF_TOTAL));
        kern<<<ntiles, THREADS, SMF_TOTAL, st>>>(map, p);
    }
    else {
        auto kern = acc_local_cnt_kernel<WMODE>;
        HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMC_TOTAL));
        kern<<<ntiles, THREADS, SMC_TOTAL, st>>>(map, p);
    }
    HT_LAUNCH_CHECK();
    return 0;
}

template <bool FINAL>
int dispatch_tile(int mode, const CUtensorMap &map, const AccArgs &p, cudaStream_t st)
{
    switch (mode) {
    case W_ONE: return launch_cnt<W_ONE, FINAL>(map, p, st);
    case W_TAINT: return launch_cnt<W_TAINT, FINAL>(map, p, st);
    case W_F32: return launch_w<FINAL>(map, p, st);
    }
    return HT_ERR_ARG;
}

// agg_out[n] (slot order) must arrive zeroed where a slot is not a node.  Scratch:
// a / arr0 / arr1 (n values each), ptrA / ptrB (n), idmap (n, only with a list),
// rootA / rootB (n each, only with root_out).
template <typename V>
int launch_resolve(const void *base0, const int32_t *next0, int64_t n, const int32_t *list,
                   const int *nlist, int32_t *idmap, int32_t *ptrA, int32_t *ptrB, void *a_,
                   void *arr0_, void *arr1_, int32_t *rootA, int32_t *rootB, void *agg_out_,
                   int32_t *root_out, int *flags, int *rounds_dev, cudaStream_t st,
                   int add_to_masked = 0, int ctas_per_sm = 0, int edge_tiles_x = 0,
                   int edge_tiles_y = 0, int edge_last_h = 0)
{
    auto kern = coarse_resolve_kernel<V>;
    int per_sm = 0;
    HT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1)
        return HT_ERR_DRIVER;
    int dev = 0, sms = 0;
    HT_CUDA_TRY(cudaGetDevice(&dev));
    HT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // the delta resolve of a sharded run visits a few per cent of the nodes and shares the
    // GPU with stage C of the tiles it does not touch: one CTA per SM is plenty
    if (ctas_per_sm > 0 && per_sm > ctas_per_sm)
        per_sm = ctas_per_sm;
    int grid = sms * per_sm;
    const int64_t need = (n + 255) / 256;
    if (grid > need)
        grid = (int)(need > 0 ? need : 1);
    HT_CUDA_TRY(cudaMemsetAsync(flags, 0, 64 * sizeof(int), st));      // per-round flags
    HT_CUDA_TRY(cudaMemsetAsync(flags + 96, 0, sizeof(int), st));      // grid barrier counter
    const V *b0 = (const V *)base0;
    V *a = (V *)a_, *arr0 = (V *)arr0_, *arr1 = (V *)arr1_, *agg_out = (V *)agg_out_;
    void *args[] = {&b0, &next0, &n, &list, &nlist, &idmap, &ptrA, &ptrB, &a, &arr0, &arr1,
                    &rootA, &rootB, &agg_out, &root_out, &flags, &rounds_dev, &add_to_masked,
                    &edge_tiles_x, &edge_tiles_y, &edge_last_h};
    HT_CUDA_TRY(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(256), args, 0, st));
    return 0;
}

// HT_STAGE_B = flat | two forces one variant of stage B (tests); default: two levels from
// one full super-tile on
std::atomic<int> g_stage_b{-1};
int stage_b_choice()
{
    int choice = g_stage_b.load();
    if (choice < 0) {
        const char *e = getenv("HT_STAGE_B");
        choice = !e ? 0 : (e[0] == 'f' ? 1 : (e[0] == 't' ? 2 : 0));
        g_stage_b.store(choice);
    }
    return choice;
}

int resolve_two_level(Workspace &wk, int64_t rows, int64_t cols, int want_roots, int *rounds_dev,
                      cudaStream_t st)
{
    const int tiles_x = ht::cdiv(cols, T), tiles_y = ht::cdiv(rows, T);
    const int sup_x = ht::cdiv(tiles_x, SGX), sup_y = ht::cdiv(tiles_y, SGY);
    const int last_h = (int)(rows - (int64_t)(tiles_y - 1) * T);
    HT_CUDA_TRY(cudaMemsetAsync(wk.outer, 0, wk.nslots, st));
    HT_CUDA_TRY(cudaMemsetAsync(wk.base2, 0, wk.nslots * 8, st));
    // root1 = -1 for the slots of the first / last tile row that are not nodes (row bands)
    const int64_t nts = (int64_t)tiles_x * tiles_y * SLOTS, row_slots = (int64_t)tiles_x * SLOTS;
    if (want_roots) {
        HT_CUDA_TRY(cudaMemsetAsync(wk.root1, 0xFF, row_slots * 4, st));
        HT_CUDA_TRY(cudaMemsetAsync(wk.root1 + nts - row_slots, 0xFF, row_slots * 4, st));
    }
    SupArgs a;
    a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.sup_x = sup_x;
    a.ntile_slots = tiles_x * tiles_y * SLOTS;
    a.list = wk.list; a.next0 = wk.next0; a.tile_first = wk.tile_first; a.tile_cnt = wk.tile_cnt;
    a.idmap = wk.idmap;
    a.base0 = (const unsigned long long *)wk.base0;
    a.outer = wk.outer;
    a.base2 = (unsigned long long *)wk.base2;
    a.rnode = wk.rnode; a.s1w = wk.s1w;
    a.agg = (unsigned long long *)wk.aggB;
    a.root1 = want_roots ? wk.root1 : nullptr;
    a.edge_tiles_x = tiles_x; a.edge_tiles_y = tiles_y; a.edge_last_h = last_h;
    HT_CUDA_TRY(cudaFuncSetAttribute(super_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SUP_SM_TOTAL));
    HT_CUDA_TRY(cudaFuncSetAttribute(super_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SUP_SM_TOTAL));
    super_tile_kernel<1><<<sup_x * sup_y, SUP_THREADS, SUP_SM_TOTAL, st>>>(a);
    HT_LAUNCH_CHECK();
    {
        const int64_t nv = wk.nslots - (int64_t)a.ntile_slots; // 2 * cols
        virtual_nodes_kernel<<<ht::cdiv(nv, 256), 256, 0, st>>>((const unsigned long long *)wk.base0,
                                                               (unsigned long long *)wk.base2, wk.outer,
                                                               (int64_t)a.ntile_slots, nv);
        HT_LAUNCH_CHECK();
    }
    HT_CUDA_TRY(cudaMemsetAsync(wk.nlist2, 0, sizeof(int), st));
    outer_nodes_kernel<<<ht::kSMs * 4, 256, 0, st>>>(wk.outer, wk.list, wk.nlist, wk.rnode, wk.next0,
                                                     a.ntile_slots, wk.next2, wk.list2, wk.nlist2);
    HT_LAUNCH_CHECK();
    int rc = launch_resolve<unsigned long long>(wk.base2, wk.next2, wk.nslots, wk.list2, wk.nlist2, wk.idmapX,
                                                wk.ptrA, wk.ptrB, wk.aggA, wk.arr0, wk.arr1,
                                                want_roots ? wk.rootA : nullptr, want_roots ? wk.rootC : nullptr,
                                                wk.aggB, want_roots ? wk.rootX : nullptr, wk.flags, rounds_dev, st);
    if (rc)
        return rc;
    super_tile_kernel<3><<<sup_x * sup_y, SUP_THREADS, SUP_SM_TOTAL, st>>>(a);
    HT_LAUNCH_CHECK();
    if (want_roots) {
        edge_roots_kernel<<<ht::cdiv(2ll * tiles_x * T, 256), 256, 0, st>>>(wk.rootB, wk.root1, wk.next0, wk.rootX,
                                                                          tiles_x, tiles_y, last_h);
        HT_LAUNCH_CHECK();
    }
    return 0;
}

int fill_args(AccArgs &p, int64_t rows, int64_t cols, int64_t row0, int64_t total_rows,
              int halo_top, int halo_bot)
{
    if (rows <= 0 || cols <= 0 || row0 < 0 || total_rows < row0 + rows)
        return HT_ERR_ARG;
    if (nslots_of(rows, cols) >= INT32_MAX)
        return HT_ERR_ARG; // boundary graph is indexed with int32
    p.rows = rows; p.cols = cols; p.row0 = row0; p.total_rows = total_rows;
    p.halo_top = halo_top ? 1 : 0; p.halo_bot = halo_bot ? 1 : 0;
    p.tiles_x = ht::cdiv(cols, T); p.tiles_y = ht::cdiv(rows, T);
    p.base = nullptr; p.next = nullptr; p.inflow = nullptr;
    p.w = nullptr; p.w_ld = 0; p.has_wnd = 0; p.wnd = 0.f;
    p.acc = nullptr; p.acc_ld = 0; p.dir_out = nullptr; p.dir_ld = 0;
    p.lacc = nullptr; p.lacc_ld = 0; p.packed = nullptr; p.packed_ld = 0;
    p.list = nullptr; p.nlist = nullptr;
    p.lacc64 = nullptr; p.lacc64_ld = 0; p.tile_e = nullptr; p.gmax_bits = nullptr;
    p.tile_phase = nullptr; p.phase = 0;
    p.tile_first = nullptr; p.tile_cnt = nullptr;
    return 0;
}

} // namespace

// 0 = automatic (two levels from 128 tiles on), 1 = always the flat resolve, 2 = always two
// levels; returns the previous setting.  Same as the environment variable HT_STAGE_B
// (flat / two); a test hook -- both variants give identical results.
extern "C" int ht_acc_stage_b_mode(int mode)
{
    const int prev = stage_b_choice();
    if (mode >= 0 && mode <= 2)
        g_stage_b.store(mode);
    return prev;
}

extern "C" size_t ht_acc_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows <= 0 || cols <= 0)
        return 256;
    return workspace_bytes(rows, cols, W_ONE);
}

extern "C" size_t ht_acc_workspace_bytes_mode(int64_t rows, int64_t cols, int mode)
{
    if (rows <= 0 || cols <= 0 || mode < 0 || mode > 2)
        return 256;
    return workspace_bytes(rows, cols, mode);
}

// Weighted mode: bits of the largest finite |weight| (nodata / NaN / inf skipped) of the
// caller's rows, as a float in an unsigned (non-negative floats order like their bits, so
// row bands combine their values with an integer MAX all-reduce).
extern "C" int ht_acc_weight_max(const float *w, int64_t w_ld, int64_t rows, int64_t cols,
                                 int has_wnd, float wnd, unsigned *max_bits, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!max_bits || rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (!w || w_ld < cols)))
        return HT_ERR_ARG;
    HT_CUDA_TRY(cudaMemsetAsync(max_bits, 0, sizeof(unsigned), st));
    if (rows == 0 || cols == 0)
        return HT_OK;
    const int64_t n = rows * cols;
    const int grid = (int)min((int64_t)ht::kSMs * 8, (n / 4 + 255) / 256 + 1);
    weight_max_kernel<<<grid, 256, 0, st>>>(w, w_ld, rows, cols, has_wnd, wnd, max_bits);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Weighted mode: hand the (all-reduced) maximum to the workspace; it fixes the global
// fixed-point unit of stages A, B and C.
extern "C" int ht_acc_set_weight_max(int64_t rows, int64_t cols, void *ws, size_t ws_bytes,
                                     const unsigned *max_bits, void *stream)
{
    if (rows <= 0 || cols <= 0 || !max_bits)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, W_F32, wk);
    if (rc)
        return rc;
    HT_CUDA_TRY(cudaMemcpyAsync(wk.gmax_bits, max_bits, sizeof(unsigned), cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return HT_OK;
}

// Stage A.  After it, the workspace holds the band's boundary graph.
extern "C" int ht_acc_local(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                            int64_t cols, int64_t row0, int64_t total_rows, int halo_top,
                            int halo_bot, int mode, const float *w, int64_t w_ld, int has_wnd,
                            float wnd, void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0 || cols == 0)
        return HT_OK;
    AccArgs p;
    int rc = fill_args(p, rows, cols, row0, total_rows, halo_top, halo_bot);
    if (rc)
        return rc;
    if (!packed || packed_ld < cols || (mode == W_F32 && (!w || w_ld < cols)) || mode < 0 || mode > 2)
        return HT_ERR_ARG;
    Workspace wk;
    if ((rc = carve(ws, ws_bytes, rows, cols, mode, wk)))
        return rc;
    CUtensorMap map;
    if ((rc = make_packed_map(map, packed, packed_ld, rows, cols, p.halo_top, p.halo_bot)))
        return rc;
    HT_CUDA_TRY(cudaMemsetAsync(wk.base0, 0, wk.nslots * wk.vsize, st));
    HT_CUDA_TRY(cudaMemsetAsync(wk.next0, 0xFF, wk.nslots * 4, st));
    p.base = wk.base0; p.next = wk.next0;
    p.w = w; p.w_ld = w_ld; p.has_wnd = has_wnd; p.wnd = wnd;
    p.lacc = wk.lacc; p.lacc_ld = wk.lacc_ld;
    p.lacc64 = wk.lacc64; p.lacc64_ld = wk.lacc64_ld; p.tile_e = wk.tile_e; p.gmax_bits = wk.gmax_bits;
    p.packed = packed; p.packed_ld = packed_ld;
    p.list = wk.list; p.nlist = wk.nlist;
    p.tile_first = wk.tile_first; p.tile_cnt = wk.tile_cnt;
    HT_CUDA_TRY(cudaMemsetAsync(wk.nlist, 0, sizeof(int), st));
    if ((rc = dispatch_tile<false>(mode, map, p, st)))
        return rc;
    if (p.halo_top || p.halo_bot) { // cells of the neighbouring bands are nodes too
        append_range_kernel<<<1, 256, 0, st>>>(wk.list, wk.nlist, (int32_t)(wk.nslots - 2 * cols),
                                               (int32_t)(2 * cols));
        HT_LAUNCH_CHECK();
    }
    return HT_OK;
}

// Stage B.  Resolved inflow per slot ends up in the workspace ("aggB"); with
// `want_roots` the last node of every slot's path ("rootB") is kept as well
// (row-band sharding needs it to build the band-level graph).
extern "C" int ht_acc_resolve(int64_t rows, int64_t cols, int mode, int want_roots, void *ws,
                              size_t ws_bytes, int *rounds_dev, void *stream)
{
    if (rows == 0 || cols == 0)
        return HT_OK;
    if (mode < 0 || mode > 2)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    // rootB is the slot-ordered output ("last node of each slot's path")
    int32_t *rA = want_roots ? wk.rootA : nullptr, *rC = want_roots ? wk.rootC : nullptr;
    int32_t *rOut = want_roots ? wk.rootB : nullptr;
    HT_CUDA_TRY(cudaMemsetAsync(wk.aggB, 0, wk.nslots * wk.vsize, (cudaStream_t)stream));
    // the last node of a path is only ever asked for at the slots of the band's first / last row
    const int etx = want_roots ? ht::cdiv(cols, T) : 0, ety = ht::cdiv(rows, T);
    const int elh = (int)(rows - (int64_t)(ety - 1) * T);
    if (mode != W_F32) {
        const int choice = stage_b_choice();
        const int64_t ntiles = (int64_t)ht::cdiv(cols, T) * ety;
        if (choice == 2 || (choice == 0 && ntiles >= 4 * SUP_NT))
            return resolve_two_level(wk, rows, cols, want_roots, rounds_dev, (cudaStream_t)stream);
    }
    if (mode == W_F32)
        return launch_resolve<i128>(wk.base0, wk.next0, wk.nslots, wk.list, wk.nlist, wk.idmap,
                                    wk.ptrA, wk.ptrB, wk.aggA, wk.arr0, wk.arr1, rA, rC, wk.aggB,
                                    rOut, wk.flags, rounds_dev, (cudaStream_t)stream, 0, 0, etx, ety, elh);
    return launch_resolve<unsigned long long>(wk.base0, wk.next0, wk.nslots, wk.list, wk.nlist,
                                              wk.idmap, wk.ptrA, wk.ptrB, wk.aggA, wk.arr0, wk.arr1,
                                              rA, rC, wk.aggB, rOut, wk.flags, rounds_dev,
                                              (cudaStream_t)stream, 0, 0, etx, ety, elh);
}

// Row-band sharding, count modes, after ht_acc_local: tag the slots of the band's first
// (halo_top) / last (halo_bot) row so that ht_acc_resolve leaves a mark on every node
// downstream of them (see ht_acc_resolve_delta).
extern "C" int ht_acc_mark_edges(int64_t rows, int64_t cols, int mode, int halo_top, int halo_bot,
                                 void *ws, size_t ws_bytes, void *stream)
{
    if (rows <= 0 || cols <= 0 || mode < 0 || mode > 2)
        return HT_ERR_ARG;
    if (mode == W_F32 && 2 * cols >= (1ll << W_MARK_BITS))
        return HT_ERR_ARG; // the marks of one band would overflow their 24 bits
    if (!halo_top && !halo_bot)
        return HT_OK;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    mark_edges_kernel<<<ht::cdiv(cols, 256), 256, 0, (cudaStream_t)stream>>>(
        (unsigned long long *)wk.base0, mode == W_F32 ? 2 : 1, mode == W_F32 ? 1ull : 1ull << MARK_SHIFT,
        rows, cols, ht::cdiv(cols, T), ht::cdiv(rows, T), halo_top ? 1 : 0, halo_bot ? 1 : 0);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Row-band sharding: add the flows the neighbour bands deliver into the band's first /
// last row (imp_top / imp_bot: cols values of the mode's type -- u64 counts or 128-bit
// weighted sums -- or NULL) to the boundary graph; a second ht_acc_resolve follows.
extern "C" int ht_acc_add_imports(int64_t rows, int64_t cols, int mode, const void *imp_top,
                                  const void *imp_bot, void *ws, size_t ws_bytes, void *stream)
{
    if (rows <= 0 || cols <= 0 || mode < 0 || mode > 2)
        return HT_ERR_ARG;
    if (!imp_top && !imp_bot)
        return HT_OK;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    const int grid = ht::cdiv(cols, 256), tiles_x = ht::cdiv(cols, T), tiles_y = ht::cdiv(rows, T);
    if (mode == W_F32)
        add_imports_kernel<i128><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (i128 *)wk.base0, (const i128 *)imp_top, (const i128 *)imp_bot, rows, cols, tiles_x, tiles_y);
    else
        add_imports_kernel<unsigned long long><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (unsigned long long *)wk.base0, (const unsigned long long *)imp_top,
            (const unsigned long long *)imp_bot, rows, cols, tiles_x, tiles_y);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Row-band sharding.  ht_acc_resolve must have run on a boundary graph whose band-edge
// slots carried a mark (ht_acc_mark_edges: 1 << 40 added to the u64 counts, 1 to the low
// word of the 128-bit weighted sums); `imp_top` / `imp_bot` (cols values of the mode's
// type each, or NULL) are the flows the neighbour bands deliver
// into the cells of the band's first / last row.  Adds their effect to the resolved
// inflow of every node downstream and clears the marks -- a resolve over the
// marked nodes only instead of a second full one.
extern "C" int ht_acc_resolve_delta(int64_t rows, int64_t cols, int mode, const void *imp_top,
                                    const void *imp_bot, void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0 || cols == 0)
        return HT_OK;
    if (mode < 0 || mode > 2)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    HT_CUDA_TRY(cudaMemsetAsync(wk.nlist2, 0, sizeof(int), st));
    if (mode == W_F32) {
        marked_nodes_kernel<i128><<<ht::kSMs * 4, 256, 0, st>>>(
            (const i128 *)wk.aggB, wk.list, wk.nlist, wk.list2, wk.nlist2, (i128 *)wk.base2,
            (const i128 *)imp_top, (const i128 *)imp_bot, rows, cols, ht::cdiv(cols, T), ht::cdiv(rows, T));
        HT_LAUNCH_CHECK();
        return launch_resolve<i128>(wk.base2, wk.next0, wk.nslots, wk.list2, wk.nlist2, wk.idmap, wk.ptrA,
                                    wk.ptrB, wk.aggA, wk.arr0, wk.arr1, nullptr, nullptr, wk.aggB, nullptr,
                                    wk.flags, nullptr, st, 1, 1);
    }
    marked_nodes_kernel<unsigned long long><<<ht::kSMs * 4, 256, 0, st>>>(
        (const unsigned long long *)wk.aggB, wk.list, wk.nlist, wk.list2, wk.nlist2,
        (unsigned long long *)wk.base2, (const unsigned long long *)imp_top,
        (const unsigned long long *)imp_bot, rows, cols, ht::cdiv(cols, T), ht::cdiv(rows, T));
    HT_LAUNCH_CHECK();
    return launch_resolve<unsigned long long>(wk.base2, wk.next0, wk.nslots, wk.list2, wk.nlist2,
                                              wk.idmap, wk.ptrA, wk.ptrB, wk.aggA, wk.arr0, wk.arr1,
                                              nullptr, nullptr, wk.aggB, nullptr, wk.flags, nullptr,
                                              st, 1, 1);
}

extern "C" int ht_band_tables(int64_t rows, int64_t cols, int mode, void *ws, size_t ws_bytes,
                              long long *out, void *stream)
{
    if (rows <= 0 || cols <= 0 || !out || mode < 0 || mode > 2)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    if (mode == W_F32)
        band_tables_w_kernel<<<ht::cdiv(cols, 256), 256, 0, (cudaStream_t)stream>>>(
            (const i128 *)wk.aggB, wk.rootB, rows, cols, ht::cdiv(cols, T), ht::cdiv(rows, T),
            wk.nslots - 2 * cols, out);
    else
        band_tables_kernel<<<ht::cdiv(cols, 256), 256, 0, (cudaStream_t)stream>>>(
            (const unsigned long long *)wk.aggB, wk.rootB, rows, cols, ht::cdiv(cols, T),
            ht::cdiv(rows, T), wk.nslots - 2 * cols, out);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_band_graph(const long long *every, int bands, int64_t cols, int mode, void *base,
                             int32_t *next, void *stream)
{
    if (!every || !base || !next || bands <= 0 || cols <= 0 || mode < 0 || mode > 2 ||
        (int64_t)bands * 2 * cols >= INT32_MAX)
        return HT_ERR_ARG;
    const int grid = ht::cdiv((int64_t)bands * 2 * cols, 256);
    if (mode == W_F32)
        band_graph_w_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(every, bands, cols, (i128 *)base, next);
    else
        band_graph_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(every, bands, cols,
                                                                  (unsigned long long *)base, next);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Byte offsets (from the workspace pointer) of the boundary-graph arrays of a COUNT-mode
// workspace, for tests and tools: layout[0..8] = base0, aggA, aggB (= resolved inflow),
// next0, ptrA, ptrB, rootA, rootB (= last node of each path), flags; layout[9] = number
// of slots, layout[10] = first virtual slot (cells of the band above, then of the
// band below, `cols` each).
extern "C" int ht_acc_workspace_layout(int64_t rows, int64_t cols, void *ws, size_t ws_bytes,
                                       int64_t *layout11)
{
    if (!layout11 || rows <= 0 || cols <= 0)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, W_ONE, wk);
    if (rc)
        return rc;
    const uint8_t *b = (const uint8_t *)ws;
    const void *ptrs[9] = {wk.base0, wk.aggA, wk.aggB, wk.next0, wk.ptrA, wk.ptrB, wk.rootA, wk.rootB, wk.flags};
    for (int i = 0; i < 9; i++)
        layout11[i] = (int64_t)((const uint8_t *)ptrs[i] - b);
    layout11[9] = wk.nslots;
    layout11[10] = wk.nslots - 2 * cols;
    return HT_OK;
}

// Generic forest resolve (subtree sums): agg_out[v] = base[v] + sum of base[u] over
// all u whose path along next[] passes through v.  Resolves the band-level boundary
// graph of a row-band sharded raster.  vtype: 0 = u64, 1 = fp64, 2 = 128-bit integers
// (16 B per value).  tmp_bytes >= (3 * value size + 8) * n + 4096.
extern "C" int ht_forest_resolve(const void *base, const int32_t *next, int64_t n, int vtype,
                                 void *agg_out, void *tmp, size_t tmp_bytes, void *stream)
{
    if (n == 0)
        return HT_OK;
    if (!base || !next || !agg_out || !tmp || n < 0 || vtype < 0 || vtype > 2)
        return HT_ERR_ARG;
    const size_t vs = vtype == 2 ? 16 : 8;
    const size_t need = 3 * align_up(n * vs) + 2 * align_up(n * 4) + 512 + 256;
    if (tmp_bytes < need)
        return HT_ERR_WORKSPACE;
    uint8_t *b = (uint8_t *)(((uintptr_t)tmp + 255) / 256 * 256);
    void *a = b; b += align_up(n * vs);
    void *arr0 = b; b += align_up(n * vs);
    void *arr1 = b; b += align_up(n * vs);
    int32_t *ptrA = (int32_t *)b; b += align_up(n * 4);
    int32_t *ptrB = (int32_t *)b; b += align_up(n * 4);
    int *flags = (int *)b;
    if (vtype == 1)
        return launch_resolve<double>(base, next, n, nullptr, nullptr, nullptr, ptrA, ptrB, a, arr0,
                                      arr1, nullptr, nullptr, agg_out, nullptr, flags, nullptr,
                                      (cudaStream_t)stream);
    if (vtype == 2)
        return launch_resolve<i128>(base, next, n, nullptr, nullptr, nullptr, ptrA, ptrB, a, arr0,
                                    arr1, nullptr, nullptr, agg_out, nullptr, flags, nullptr,
                                    (cudaStream_t)stream);
    return launch_resolve<unsigned long long>(base, next, n, nullptr, nullptr, nullptr, ptrA, ptrB, a,
                                              arr0, arr1, nullptr, nullptr, agg_out, nullptr, flags,
                                              nullptr, (cudaStream_t)stream);
}


// Stage C.
// `phase` < 0: every tile.  0 / 1 (after ht_acc_tile_phases): only the tiles whose inflow does
// not / does depend on the neighbour bands' deliveries.
extern "C" int ht_acc_final_phase(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                                  int64_t cols, int64_t row0, int64_t total_rows, int halo_top,
                                  int halo_bot, int mode, const float *w, int64_t w_ld, int has_wnd,
                                  float wnd, double *acc, int64_t acc_ld, int8_t *dir_out,
                                  int64_t dir_ld, void *ws, size_t ws_bytes, int phase, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0 || cols == 0)
        return HT_OK;
    AccArgs p;
    int rc = fill_args(p, rows, cols, row0, total_rows, halo_top, halo_bot);
    if (rc)
        return rc;
    if (!packed || packed_ld < cols || mode < 0 || mode > 2 || (acc && acc_ld < cols) ||
        (dir_out && dir_ld < cols))
        return HT_ERR_ARG;
    // count modes store 2 doubles / 4 direction bytes at a time
    if (mode != W_F32 && ((acc && (((uintptr_t)acc & 15) || (acc_ld & 1))) ||
                          (dir_out && (((uintptr_t)dir_out & 3) || (dir_ld & 3)))))
        return HT_ERR_ALIGN;
    Workspace wk;
    if ((rc = carve(ws, ws_bytes, rows, cols, mode, wk)))
        return rc;
    CUtensorMap map;
    rc = ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1,
                         packed - (int64_t)p.halo_top * packed_ld,
                         rows + p.halo_top + p.halo_bot, cols, packed_ld, T, T, false);
    if (rc)
        return rc;
    p.lacc = wk.lacc; p.lacc_ld = wk.lacc_ld;
    p.lacc64 = wk.lacc64; p.lacc64_ld = wk.lacc64_ld; p.tile_e = wk.tile_e; p.gmax_bits = wk.gmax_bits;
    p.packed = packed; p.packed_ld = packed_ld;
    p.inflow = wk.aggB;
    p.w = w; p.w_ld = w_ld; p.has_wnd = has_wnd; p.wnd = wnd;
    p.acc = acc; p.acc_ld = acc_ld; p.dir_out = dir_out; p.dir_ld = dir_ld;
    if (phase >= 0) {
        p.tile_phase = reinterpret_cast<const uint8_t *>(wk.rootC);
        p.phase = phase ? 1 : 0;
    }
    return dispatch_tile<true>(mode, map, p, st);
}

extern "C" int ht_acc_final(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                            int64_t cols, int64_t row0, int64_t total_rows, int halo_top,
                            int halo_bot, int mode, const float *w, int64_t w_ld, int has_wnd,
                            float wnd, double *acc, int64_t acc_ld, int8_t *dir_out,
                            int64_t dir_ld, void *ws, size_t ws_bytes, void *stream)
{
    return ht_acc_final_phase(packed, packed_ld, rows, cols, row0, total_rows, halo_top, halo_bot,
                              mode, w, w_ld, has_wnd, wnd, acc, acc_ld, dir_out, dir_ld, ws,
                              ws_bytes, -1, stream);
}

// Sharded runs, after ht_acc_resolve(want_roots = 1) on a graph whose band-edge slots were
// marked (ht_acc_mark_edges): one byte per tile, 1 = an entry slot of the tile carries a mark.
extern "C" int ht_acc_tile_phases(int64_t rows, int64_t cols, int mode, void *ws, size_t ws_bytes,
                                  void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (rows <= 0 || cols <= 0)
        return HT_OK;
    if (mode < 0 || mode > 2)
        return HT_ERR_ARG;
    Workspace wk;
    int rc = carve(ws, ws_bytes, rows, cols, mode, wk);
    if (rc)
        return rc;
    const int ntiles = ht::cdiv(rows, T) * ht::cdiv(cols, T);
    uint8_t *ph = reinterpret_cast<uint8_t *>(wk.rootC);
    HT_CUDA_TRY(cudaMemsetAsync(ph, 0, (size_t)ntiles, st));
    const int grid = ht::kSMs * 8;
    if (mode == W_F32)
        tile_phase_kernel<i128><<<grid, 256, 0, st>>>((const i128 *)wk.aggB, wk.list, wk.nlist,
                                                      ntiles * SLOTS, ph);
    else
        tile_phase_kernel<unsigned long long><<<grid, 256, 0, st>>>(
            (const unsigned long long *)wk.aggB, wk.list, wk.nlist, ntiles * SLOTS, ph);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Stages A + B + C on one device.
extern "C" int ht_acc_f64(const uint8_t *packed, int64_t packed_ld, int64_t rows, int64_t cols,
                          int mode, const float *w, int64_t w_ld, int has_wnd, float wnd,
                          double *acc, int64_t acc_ld, int8_t *dir_out, int64_t dir_ld,
                          void *ws, size_t ws_bytes, void *stream)
{
    int rc;
    if (mode == W_F32 && rows > 0 && cols > 0) {
        Workspace wk;
        if ((rc = carve(ws, ws_bytes, rows, cols, mode, wk)))
            return rc;
        if ((rc = ht_acc_weight_max(w, w_ld, rows, cols, has_wnd, wnd, wk.gmax_bits, stream)))
            return rc;
    }
    rc = ht_acc_local(packed, packed_ld, rows, cols, 0, rows, 0, 0, mode, w, w_ld, has_wnd,
                      wnd, ws, ws_bytes, stream);
    if (rc)
        return rc;
    if ((rc = ht_acc_resolve(rows, cols, mode, 0, ws, ws_bytes, nullptr, stream)))
        return rc;
    return ht_acc_final(packed, packed_ld, rows, cols, 0, rows, 0, 0, mode, w, w_ld, has_wnd,
                        wnd, acc, acc_ld, dir_out, dir_ld, ws, ws_bytes, stream);
}
