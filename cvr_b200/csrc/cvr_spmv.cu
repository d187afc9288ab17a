// y = A * x over the CVR arrays on the device (sm_100a), fp64.
//
// Replaces the five-phase streaming loop of spmv_compute_kernel
// (/root/reference/spmv.cpp:1016-1667) and implements its INTENDED semantics
// (paper Alg. 4), including the steal-only chunks (split1 == -1) the
// reference kernel mishandles.
//
// Three generations live in this file; all pass the same parity suite and stay selectable
// (CVR_SPMV_KERNEL = tma | ldg | window) so the choice can be re-measured:
//   1. cvr_spmv_window_kernel     warp covers 4 consecutive steps, segmented reduction per window
//                                 (instruction-bound, profiles/r01_v0_*)
//   2. cvr_spmv_tile_kernel<LDG>  tile walker, operands through the LSU (profiles/r01_v1_*)
//   3. cvr_spmv_tile_kernel<TMA>  tile walker, operands staged by the TMA engine -- the default
//                                 (profiles/r01_v3_*); <..., kPublish> additionally pushes finished
//                                 rows to peer GPUs for the iterated multi-GPU SpMV.
//
// Common semantics.  One WARP owns one chunk (the reference: one OpenMP thread).  A record
// (pos, wb) means "the accumulator of SIMD lane pos%8 is flushed before step pos/8"
// (spmv.cpp:1197-1210):
//   feeding record (pos <= split1)  -> y[wb] = sum          plain store, row owned by chunk
//   stealing record                 -> carry[lane] += sum   (wb == own lane on a first steal)
//   split0 position                 -> y[first_row] += sum  atomic, row shared with the
//                                                           previous chunk (spmv.cpp:1280)
// After the last step the eight pos=-1 records route each lane's remainder into a carry
// slot (spmv.cpp:1633-1638) and the eight carries are added atomically to y[tail[.]]
// (spmv.cpp:1640-1649).  The accumulated / never-written rows of y are cleared by
// cvr_clear_rows_kernel on the same stream, inside the timed region.
//
// The first generation, kept below: thread t of the warp holds element 32k + t of window k,
// i.e. step 4k + (t >> 3), SIMD lane (t & 7); vals / cols are read with coalesced 256 B / 128 B
// warp loads, x is gathered through L1/L2 (ld.global.nc); the warp holds the next 32 records in
// registers, turns the ones inside the window into a flag word with one REDUX.OR and runs a
// segmented reduction along each SIMD lane (stride-8 shuffles); windows without a flag: one FMA.
#include "cvr_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int UNROLL = 4; // windows whose loads are in flight together

__device__ __forceinline__ double ld_stream_f64(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t* p)
{
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

struct WarpState {
    double acc;   // private partial sum of my SIMD lane since its last flush
    double carry; // private share of t_rets[my lane] (spmv.cpp:1124)
    int2 held;    // record rb + t
    int32_t rb, rc; // record batch base / records consumed
    int32_t last_pos;
};

// Segmented flush of one window that contains at least one flag.
__device__ __forceinline__ void flush_window(WarpState& st, const int2* __restrict__ rec,
                                             int32_t n_rec, double prod, unsigned rflags,
                                             unsigned flags, int32_t wstart, int t,
                                             int32_t split0, int32_t split1, int32_t first_row,
                                             const int32_t* __restrict__ tail,
                                             double* __restrict__ y)
{
    const int j = t >> 3, l = t & 7;
    // lane totals of the private partial sums
    double tot = st.acc + __shfl_xor_sync(FULL, st.acc, 8);
    tot += __shfl_xor_sync(FULL, tot, 16);
    // products of the up-to-three earlier steps of my SIMD lane inside this window
    const double q1 = __shfl_up_sync(FULL, prod, 8);
    const double q2 = __shfl_up_sync(FULL, prod, 16);
    const double q3 = __shfl_up_sync(FULL, prod, 24);
    // my record, if my position is flagged by one
    const int rank = __popc(rflags & ((1u << t) - 1u));
    const int32_t wb = __shfl_sync(FULL, st.held.y, (st.rc - st.rb + rank) & 31);

    const unsigned lane_bits = flags & (0x01010101u << l);
    if (lane_bits == 0) {
        st.acc += prod;
    } else {
        if ((flags >> t) & 1u) {
            // sum of the segment that ends right before my step
            double e = 0.0;
            bool open = true;
            if (j >= 1) { e += q1; open = !((flags >> (t - 8)) & 1u); }
            if (j >= 2 && open) { e += q2; open = !((flags >> (t - 16)) & 1u); }
            if (j >= 3 && open) { e += q3; open = !((flags >> (t - 24)) & 1u); }
            if (open) e += tot;
            const int32_t pos = wstart + t;
            if ((rflags >> t) & 1u) {
                if (split1 != -1 && pos <= split1) y[wb] = e; // feeding: exclusive row
                else if (wb == l) st.carry += e;               // stealing: a first steal names its own lane
                else if (tail[wb] != 0) atomicAdd(&y[tail[wb]], e); // (second steal, unreachable: SURVEY 8a-R2 note i)
            } else {
                atomicAdd(&y[first_row], e);                   // split0: shared first row
            }
        }
        // elements at or after the lane's last flag open the next segment
        st.acc = ((lane_bits >> t) >> 1) ? 0.0 : prod;
    }
    st.rc += __popc(rflags);
    (void)rec; (void)n_rec; (void)split0;
}

__global__ void __launch_bounds__(128)
cvr_spmv_window_kernel(const CvrChunk* __restrict__ chunks, int32_t T,
                const double* __restrict__ vals, const int32_t* __restrict__ cols,
                const int32_t* __restrict__ record, const double* __restrict__ x,
                double* __restrict__ y)
{
    const int32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chunk >= T) return;
    const int t = threadIdx.x & 31;

    // chunk descriptor: one 64 B line, every thread reads the same words (broadcast loads)
    const CvrChunk* cp = chunks + chunk;
    const int64_t start = cp->start;
    const int32_t len = cp->len;
    const int32_t first_row = cp->first_row;
    const int32_t split0 = cp->split0;
    const int32_t split1 = cp->split1;
    const int32_t n_rec = cp->n_rec;

    const int2* rec = reinterpret_cast<const int2*>(record + cvr_record_offset(chunk, first_row));
    const double* v = vals + start;
    const int32_t* c = cols + start;

    WarpState st;
    st.acc = 0.0;
    st.carry = 0.0;
    st.rb = st.rc = 0;
    st.held = (t < n_rec) ? rec[t] : make_int2(-1, 0);
    st.last_pos = __shfl_sync(FULL, st.held.x, 31);

    const int32_t n_win = (len + CVR_WIN - 1) / CVR_WIN;
    for (int32_t k0 = 0; k0 < n_win; k0 += UNROLL) {
        double a[UNROLL], xv[UNROLL];
        int32_t ci[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int32_t p = (k0 + u) * CVR_WIN + t;
            const bool in = p < len;
            a[u] = in ? ld_stream_f64(v + p) : 0.0;
            ci[u] = in ? ld_stream_s32(c + p) : 0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) xv[u] = __ldg(x + ci[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int32_t k = k0 + u;
            if (k < n_win) { // warp-uniform
                const int32_t wstart = k * CVR_WIN;
                if (st.rc != st.rb && (uint32_t)st.last_pos < (uint32_t)(wstart + CVR_WIN)) {
                    st.rb = st.rc; // batch may not cover this window: reload from the cursor
                    st.held = (st.rb + t < n_rec) ? rec[st.rb + t] : make_int2(-1, 0);
                    st.last_pos = __shfl_sync(FULL, st.held.x, 31);
                }
                const unsigned rel = (unsigned)(st.held.x - wstart);
                const unsigned rflags = __reduce_or_sync(FULL, rel < 32u ? (1u << rel) : 0u);
                const unsigned s0rel = (unsigned)(split0 - wstart);
                const unsigned flags = rflags | ((split0 != 0 && s0rel < 32u) ? (1u << s0rel) : 0u);
                if (flags == 0) {
                    st.acc = fma(a[u], xv[u], st.acc);
                } else {
                    flush_window(st, rec, n_rec, a[u] * xv[u], rflags, flags, wstart, t, split0,
                                 split1, first_row, cp->tail, y);
                }
            }
        }
    }

    // ---- chunk epilogue: lane remainders through the eight pos=-1 records
    double tot = st.acc + __shfl_xor_sync(FULL, st.acc, 8);
    tot += __shfl_xor_sync(FULL, tot, 16);          // every thread: total of SIMD lane (t & 7)
    double carry = st.carry + __shfl_xor_sync(FULL, st.carry, 8);
    carry += __shfl_xor_sync(FULL, carry, 16);      // every thread: carry slot (t & 7)
    const int32_t term_wb = (t < CVR_W) ? rec[n_rec + t].y : 0;
#pragma unroll
    for (int l = 0; l < CVR_W; l++) {
        const double r = __shfl_sync(FULL, tot, l);
        const int32_t w = __shfl_sync(FULL, term_wb, l);
        if (t == w) carry += r;                     // t_rets[wb] += lane l (spmv.cpp:1637)
    }
    if (t < CVR_W) {
        const int32_t row = cp->tail[t];
        // row 0 is the phantom row unused lanes point at (their carry is 0.0): skipping it
        // avoids n_chunks atomics on one address
        if (row != 0) atomicAdd(&y[row], carry);    // spmv.cpp:1647-1648
    }
}


// ---------------------------------------------------------------------------------------
// Tile walker (the default kernel).
//
// ncu on the window kernel above (profiles/r01_v0_*) showed it instruction-bound: with ~27
// nnz per row some lane switches rows in 3 of 4 windows, so nearly every window paid the
// warp-wide segmented reduction (82 warp-instructions per 32 nnz, 43 % issue utilisation,
// 2.1 TB/s).  Here a warp still owns one chunk, but a pass covers a TILE of 32 steps x 8
// lanes and thread t = (q, l) walks TB CONSECUTIVE steps of SIMD lane l:
// steps TB*q .. TB*q+TB-1 of the tile.  A row switch is then a thread-local event (emit the
// accumulator, clear it); threads of the same SIMD lane only meet once per tile, in a
// 3-shuffle carry chain that hands the open partial sum from walker q to walker q+1.
//   * loads: for each of the 8 steps a warp load touches 4 x 64 B (vals) / 4 x 32 B (cols)
//     fully used sectors; all 16 loads of a tile are issued before the first use.
//   * records are delivered to their owner thread through shared memory: the warp holds 32
//     records in registers (coalesced 256 B load), each holder drops a flag byte and the
//     write-back target into the owner's slot.
//   * kTma = true (default): the vals/cols stream does not go through the LSU at all.  ncu on
//     the LDG variant (profiles/r01_v1_*) shows the L1TEX data pipe as the busiest unit
//     (57-74 % of its wavefront peak): every x gather costs one wavefront per distinct line,
//     so the streamed operands are moved by the TMA engine instead: per tile one lane arms an
//     mbarrier and eight lanes issue cp.async.bulk (global -> shared, L2 evict-first), two
//     stages deep, i.e. the copy of tile k+1/k+2 overlaps the gather + FMA walk of tile k.
//     Each walker's quarter of the tile lands in its own padded slot so that the 64-bit
//     shared loads of the four walkers fall on disjoint banks.
// ---------------------------------------------------------------------------------------
// TB (consecutive steps per walker) is 9 for the TMA variant, ODD on purpose: a walker's
// quarter of the tile is then 9 x 64 B = 576 B of vals and 288 B of cols, so the four walkers
// start 16 (vals) / 8 (cols) shared-memory banks apart and their 64-bit / 32-bit loads of one
// step are conflict-free in the plain linear layout ONE bulk copy per array produces.  (With
// TB = 8 the quarters alias on the same banks; padding each quarter separately needed 8 small
// copies per tile.)  The LDG variant has no such constraint and uses TB = 8.
constexpr int WARPS = 4;              // warps per block
constexpr int32_t WB_SPLIT0 = -2;     // marker: flush into the shared first row
constexpr int32_t PUSH_MAX_ROWS = 1024; // widest row range a warp publishes as one contiguous push
#ifndef CVR_TMA_STAGES
#define CVR_TMA_STAGES 1
#endif
constexpr int STAGES = CVR_TMA_STAGES; // TMA ring depth per warp (1: the registers are the 2nd buffer)

template <int TB>
struct Geo {
    static constexpr int TILE = 4 * TB * CVR_W;      // elements per warp pass (TB=9: 288, 36 steps)
    static constexpr int QUARTER = TB * CVR_W;       // one walker's share of a tile
    static constexpr int FLAG_WORDS = (TB + 3) / 4;  // flag bytes per thread, packed in 32-bit words
    static constexpr int STAGE_BYTES = TILE * 12;    // vals then cols
    static constexpr int WARP_SMEM = STAGES * STAGE_BYTES;
    static constexpr int DYN_SMEM = WARPS * WARP_SMEM;
};

struct TileCtx {
    double* __restrict__ y;
    const int32_t* tail;
    const CvrPublish* pub; // kernel parameter (constant bank); only read when kPublish
    bool scatter;          // kPublish: publish row by row at emit (chunks whose row range is too wide)
    int32_t split1, first_row;
    int l;
};

// A finished row that this chunk owns alone: one plain store (spmv.cpp:1204).
template <bool kPublish>
__device__ __forceinline__ void store_row(const TileCtx& cx, int32_t row, double value)
{
    cx.y[row] = value;
    if (kPublish && cx.scatter) { // scattered 8-byte peer stores, row by row
        const int64_t g = cx.pub->row_offset + row;
        const uint32_t nb = cx.pub->needs ? cx.pub->needs[row] : 0xffu;
#pragma unroll
        for (int p = 0; p < CVR_MAX_PEERS; p++)
            if (p < cx.pub->n_dst && ((nb >> p) & 1u)) cx.pub->dst[p][g] = value;
    }
}

// Iterated multi-GPU SpMV: when a warp has finished a chunk it publishes the chunk's whole row
// range y[first_row .. last_row] -- contiguous, so the peer-mapped stores are coalesced 256 B
// warp stores over NVLink -- into the x vector every GPU reads in the NEXT iteration.  The transfer
// thus runs chunk by chunk inside the SpMV kernel, overlapped with the other warps' sweeps.  Rows of
// the range that are still being accumulated (shared first/last rows, tail rows) are sent as they
// are and overwritten by cvr_publish_rows_kernel once the sweep is complete (same source GPU, same
// address, stream order); empty rows carry the 0.0 the clearing kernel put there.
__device__ __forceinline__ void publish_chunk_rows(const TileCtx& cx, int32_t first_row, int32_t last_row,
                                                   int t)
{
    __threadfence_block(); // this warp's own row stores (made by other lanes) before the re-read
    __syncwarp();
    const CvrPublish& pub = *cx.pub;
    for (int32_t r0 = first_row + t; r0 <= last_row; r0 += 4 * 32) {
        double v[4];
        uint32_t nb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool in = r0 + 32 * u <= last_row;
            v[u] = in ? __ldcg(cx.y + r0 + 32 * u) : 0.0;
            nb[u] = !in ? 0u : (pub.needs ? pub.needs[r0 + 32 * u] : 0xffu);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t g = pub.row_offset + r0 + 32 * u;
#pragma unroll
            for (int p = 0; p < CVR_MAX_PEERS; p++)
                if (p < pub.n_dst && ((nb[u] >> p) & 1u)) pub.dst[p][g] = v[u];
        }
    }
}

template <bool kPublish>
__device__ __forceinline__ void emit(const TileCtx& cx, double value, int32_t pos, int32_t wb,
                                     double& carry_slot)
{
    if (wb >= 0 && pos <= cx.split1) store_row<kPublish>(cx, wb, value);   // feeding, :1204 (split1 = -1: never)
    else if (wb == WB_SPLIT0) atomicAdd(&cx.y[cx.first_row], value);       // spmv.cpp:1280-1282
    else if (wb == cx.l) carry_slot += value;                              // stealing, :1541
    else if (cx.tail[wb] != 0) atomicAdd(&cx.y[cx.tail[wb]], value);       // (unreachable)
}

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// acc = fma(a, x, acc) under a predicate: one ISETP + one predicated DFMA (the plain C++ `if` compiles
// to an unconditional DFMA plus two FSELs per chain step)
__device__ __forceinline__ void fma_if(double& acc, double a, double x, uint32_t cond)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p fma.rn.f64 %0, %1, %2, %0;\n\t}"
        : "+d"(acc) : "d"(a), "d"(x), "r"(cond));
}

// flag bytes (0/1) of a 32-bit word -> 4-bit mask
__device__ __forceinline__ uint32_t bytes_to_bits(uint32_t w) { return ((w * 0x00204081u) >> 21) & 0xfu; }

#ifndef CVR_LDG_BLOCKS
#define CVR_LDG_BLOCKS 8
#endif
#ifndef CVR_TMA_BLOCKS
#define CVR_TMA_BLOCKS 6
#endif

template <bool kTma, int TB, bool kPublish>
__global__ void __launch_bounds__(WARPS * 32, kTma ? CVR_TMA_BLOCKS : CVR_LDG_BLOCKS)
cvr_spmv_tile_kernel(const CvrChunk* __restrict__ chunks, int32_t T,
                     const double* __restrict__ vals, const int32_t* __restrict__ cols,
                     const int32_t* __restrict__ record, const double* __restrict__ x,
                     double* __restrict__ y, const __grid_constant__ CvrPublish pub)
{
    using G = Geo<TB>;
    constexpr int TILE = G::TILE, QUARTER = G::QUARTER, FLAG_WORDS = G::FLAG_WORDS;
    __shared__ uint32_t s_flags[WARPS][FLAG_WORDS][32]; // one flag byte per (thread, step)
    __shared__ int32_t s_wb[WARPS][TB][32];             // write-back target per (step, thread)
    __shared__ __align__(8) unsigned long long s_bar[WARPS][STAGES];
    extern __shared__ __align__(128) unsigned char s_stream[]; // kTma: WARPS x STAGES tiles

    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = t >> 3, l = t & 7;
    const int32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t n_warps = (gridDim.x * blockDim.x) >> 5;

    // ---- per-warp TMA ring, set up once: the warp is persistent and walks chunks
    // warp0, warp0 + n_warps, ... (chunks are nnz-balanced, so a static round robin is even)
    const uint32_t ring = kTma ? smem_u32(s_stream + w * G::WARP_SMEM) : 0u;
    const uint32_t bar0 = kTma ? smem_u32(&s_bar[w][0]) : 0u;
    uint64_t policy = 0;
    uint32_t n_issued = 0, n_waited = 0; // tiles issued to / consumed from the ring, all chunks
    if (kTma) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        if (t == 0) {
#pragma unroll
            for (int sidx = 0; sidx < STAGES; sidx++) mbar_init(bar0 + 8u * sidx, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
#pragma unroll
    for (int k = 0; k < FLAG_WORDS; k++) s_flags[w][k][t] = 0u;
    __syncwarp();

    for (int32_t chunk = warp0; chunk < T; chunk += n_warps) {
        const CvrChunk* cp = chunks + chunk;
        const int64_t start = cp->start;
        const int32_t len = cp->len;
        const int32_t split0 = cp->split0;
        const int32_t n_rec = cp->n_rec;
        TileCtx cx;
        cx.y = y;
        cx.tail = cp->tail;
        cx.pub = &pub;
        // A chunk in a very sparse region can span 10^5 (mostly empty) rows: pushing that range from
        // one warp would serialise; such chunks publish their few finished rows one by one instead
        // (their empty rows get their 0.0 from cvr_publish_rows_kernel either way).
        const int32_t chunk_last_row = cp->last_row;
        cx.scatter = kPublish && ((pub.mode & 1) || chunk_last_row - cp->first_row >= PUSH_MAX_ROWS);
        cx.split1 = cp->split1;
        cx.first_row = cp->first_row;
        cx.l = l;

        const int2* rec = reinterpret_cast<const int2*>(record + cvr_record_offset(chunk, cx.first_row));
        const double* v = vals + start;
        const int32_t* c = cols + start;
        const int32_t n_tiles = (len + TILE - 1) / TILE;

        // one elected lane arms the stage's mbarrier and issues both bulk copies of a tile
        auto issue_tile = [&](int32_t tile) {
            if (t == 0) {
                const int32_t ts = tile * TILE;
                const int32_t n_el = min(TILE, len - ts); // multiple of 16
                const uint32_t sidx = n_issued % STAGES;
                const uint32_t bar = bar0 + 8u * sidx;
                const uint32_t stage = ring + sidx * G::STAGE_BYTES;
                mbar_expect_tx(bar, (uint32_t)n_el * 12u);
                bulk_g2s(stage, v + ts, (uint32_t)n_el * 8u, bar, policy);
                bulk_g2s(stage + TILE * 8, c + ts, (uint32_t)n_el * 4u, bar, policy);
            }
            n_issued++;
        };
        if (kTma) {
#pragma unroll
            for (int sidx = 0; sidx < STAGES; sidx++)
                if (sidx < n_tiles) issue_tile(sidx);
        }

        int32_t rb = 0;
        int2 held = (t < n_rec) ? rec[t] : make_int2(-1, 0);
        double lane_carry = 0.0; // open partial sum of SIMD lane l at the tile boundary
        double carry_slot = 0.0; // private share of t_rets[l] (spmv.cpp:1124)
        // When launched as a programmatic dependent of cvr_clear_rows_kernel everything above (ring
        // set-up, descriptor and record loads, the first bulk copies) overlapped its tail; y may only
        // be written once that grid has completed.  No-op for an ordinary launch.
        asm volatile("griddepcontrol.wait;" ::: "memory");

        for (int32_t tile = 0; tile < n_tiles; tile++) {
            const int32_t ts = tile * TILE;
            const int32_t p0 = ts + q * QUARTER + l; // my element of step TB*q of the tile; next step: +8
            const bool full = ts + TILE <= len;      // warp-uniform: no bounds checks needed

            double a[TB], xv[TB];
            uint32_t ci[TB];
            if (kTma) {
                const uint32_t sidx = n_waited % STAGES;
                mbar_wait(bar0 + 8u * sidx, (n_waited / STAGES) & 1u);
                n_waited++;
                const unsigned char* stage = s_stream + w * G::WARP_SMEM + sidx * G::STAGE_BYTES;
                const double* sv = reinterpret_cast<const double*>(stage) + q * QUARTER + l;
                const uint32_t* sc = reinterpret_cast<const uint32_t*>(stage + TILE * 8) + q * QUARTER + l;
                if (full) {
#pragma unroll
                    for (int b = 0; b < TB; b++) {
                        a[b] = sv[b * CVR_W];
                        ci[b] = sc[b * CVR_W];
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < TB; b++) {
                        const bool in = p0 + b * CVR_W < len;
                        a[b] = in ? sv[b * CVR_W] : 0.0;
                        ci[b] = in ? sc[b * CVR_W] : 0u;
                    }
                }
                // The refill below is an async-proxy write to the stage these generic-proxy loads
                // just read: without the proxy fence the bulk copy can land first (observed on
                // L2-hot inputs).  Fence in every reader, then converge, then re-arm.
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (tile + STAGES < n_tiles) issue_tile(tile + STAGES);
            } else {
                const double* gv = v + p0;
                const uint32_t* gc = reinterpret_cast<const uint32_t*>(c) + p0;
                if (full) {
#pragma unroll
                    for (int b = 0; b < TB; b++) {
                        a[b] = ld_stream_f64(gv + b * CVR_W);
                        ci[b] = (uint32_t)ld_stream_s32(reinterpret_cast<const int32_t*>(gc + b * CVR_W));
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < TB; b++) {
                        const bool in = p0 + b * CVR_W < len;
                        a[b] = in ? ld_stream_f64(gv + b * CVR_W) : 0.0;
                        ci[b] = in ? (uint32_t)ld_stream_s32(reinterpret_cast<const int32_t*>(gc + b * CVR_W)) : 0u;
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < TB; b++) xv[b] = __ldg(x + ci[b]);

            // ---- deliver the records of this tile to their owner threads
            for (;;) {
                const uint32_t rel = (uint32_t)(held.x - ts);
                if (rel < (uint32_t)TILE) {
                    const uint32_t step = rel >> 3, wq = step / TB, b = step - wq * TB;
                    const uint32_t owner = wq * CVR_W + (rel & 7u);
                    reinterpret_cast<unsigned char*>(&s_flags[w][b >> 2][owner])[b & 3u] = 1;
                    s_wb[w][b][owner] = held.y;
                }
                const int32_t last = __shfl_sync(FULL, held.x, 31);
                if ((uint32_t)last >= (uint32_t)(ts + TILE)) break; // batch reaches past the tile (or ended)
                rb += 32;
                held = (rb + t < n_rec) ? rec[rb + t] : make_int2(-1, 0);
            }
            if (t == 0 && split0 != 0) {
                const uint32_t rel = (uint32_t)(split0 - ts);
                if (rel < (uint32_t)TILE) {
                    const uint32_t step = rel >> 3, wq = step / TB, b = step - wq * TB;
                    const uint32_t owner = wq * CVR_W + (rel & 7u);
                    reinterpret_cast<unsigned char*>(&s_flags[w][b >> 2][owner])[b & 3u] = 1;
                    s_wb[w][b][owner] = WB_SPLIT0;
                }
            }
            __syncwarp();
            uint32_t mask = 0; // bit b: my SIMD lane switches rows before step b of my share
#pragma unroll
            for (int k = 0; k < FLAG_WORDS; k++) {
                const uint32_t fw = s_flags[w][k][t];
                if (fw) s_flags[w][k][t] = 0u;
                mask |= bytes_to_bits(fw) << (4 * k);
            }

            // ---- walk my TB steps with two predicated FMA chains: `head` collects the steps before
            // my first flag (everything if I have none), `tail` the steps from my last flag on.
            const int b_first = mask ? __ffs(mask) - 1 : TB;
            const int b_last = mask ? 31 - __clz(mask) : TB;
            const uint32_t below = (1u << b_first) - 1u;        // steps before the first flag
            const uint32_t after = ~((1u << b_last) - 1u);      // steps from the last flag on
            double head = 0.0, tailsum = 0.0;
            const uint32_t after_m = mask ? after : 0u;
#pragma unroll
            for (int b = 0; b < TB; b++) {
                fma_if(head, a[b], xv[b], below & (1u << b));
                fma_if(tailsum, a[b], xv[b], after_m & (1u << b));
            }
            // segments strictly between two flags of the same thread (short rows): emit in place
            if (b_last > b_first) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < TB; b++) {
                    if (b > b_first && ((mask >> b) & 1u)) {
                        emit<kPublish>(cx, acc, p0 + b * CVR_W, s_wb[w][b][t], carry_slot);
                        acc = 0.0;
                    }
                    if (b >= b_first && b < b_last) acc = fma(a[b], xv[b], acc);
                }
            }

            // ---- carry chain over the four walkers of my SIMD lane
            const bool has = mask != 0u;
            double cin = (q == 0) ? lane_carry : 0.0;
            double out = has ? tailsum : cin + head;
#pragma unroll
            for (int r = 1; r < 4; r++) {
                const double prev = __shfl_up_sync(FULL, out, CVR_W);
                if (q == r) {
                    cin = prev;
                    out = has ? tailsum : cin + head;
                }
            }
            lane_carry = __shfl_sync(FULL, out, 24 + l);
            if (has) emit<kPublish>(cx, head + cin, p0 + b_first * CVR_W, s_wb[w][b_first][t], carry_slot);
            __syncwarp(); // slots are reused by the next tile's delivery
        }

        // ---- chunk epilogue: lane remainders through the eight pos=-1 records (spmv.cpp:1633-1649)
        double carry = carry_slot + __shfl_xor_sync(FULL, carry_slot, 8);
        carry += __shfl_xor_sync(FULL, carry, 16);
        const int32_t term_wb = (t < CVR_W) ? rec[n_rec + t].y : 0;
#pragma unroll
        for (int k = 0; k < CVR_W; k++) {
            const double r = __shfl_sync(FULL, lane_carry, k);
            const int32_t wbk = __shfl_sync(FULL, term_wb, k);
            if (t == wbk) carry += r;
        }
        if (t < CVR_W) {
            const int32_t row = cp->tail[t];
            if (row != 0) atomicAdd(&y[row], carry); // row 0 = phantom row of unused lanes (carry 0.0)
        }
        if (kPublish && !cx.scatter && (!pub.chunk_any || pub.chunk_any[chunk]))
            publish_chunk_rows(cx, cx.first_row, chunk_last_row, t);
    }
}

#ifndef CVR_TB_TMA
#define CVR_TB_TMA 9
#endif
constexpr int TB_TMA = CVR_TB_TMA, TB_LDG = 8;
static_assert(TB_TMA % 2 == 1, "the TMA tile needs an odd number of steps per walker (bank layout)");

// ---- small helper kernels around the sweep
// y is cleared only where it is accumulated (boundary rows) or never written (empty rows, row 0):
// every other row is stored exactly once by the sweep.  Replaces an 8*(nRows+1)-byte memset.
__global__ void cvr_clear_rows_kernel(double* __restrict__ y, const int32_t* __restrict__ boundary,
                                      int32_t n_boundary, const int32_t* __restrict__ empty, int32_t n_empty,
                                      bool skip_row0)
{
    // programmatic dependent launch: the sweep kernel behind us may start its prologue right away; it
    // waits (griddepcontrol.wait) for this grid to complete before it touches y
    asm volatile("griddepcontrol.launch_dependents;");
    const int32_t n = n_boundary + n_empty;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t row = i < n_boundary ? boundary[i] : empty[i - n_boundary];
        if (row == 0 && skip_row0) continue; // y aliases a slice of x: y[0] is the neighbour's last row
        y[row] = 0.0;
    }
}

// After the sweep of an iterated multi-GPU SpMV, ONE epilogue kernel does everything that has to
// wait for the sweep to be complete:
//   1. the accumulated rows are final now: publish them; rows nothing writes get an explicit 0.0
//      (only while `publish_empty`: a reused x buffer must not keep a stale value there -- two
//      iterations cover both buffers);
//   2. clear those rows of y again, ready for the next sweep (cvr_launch_spmv then skips its own
//      clearing kernel);
//   3. the last block to finish runs the all-to-all flag barrier over peer memory.
__global__ void cvr_publish_epilogue_kernel(double* __restrict__ y, const int32_t* __restrict__ boundary,
                                            int32_t n_boundary, const int32_t* __restrict__ empty,
                                            int32_t n_empty, const __grid_constant__ CvrPublish pub,
                                            const __grid_constant__ CvrBarrier bar, unsigned int* done_counter)
{
    // the next iteration's sweep may start its prologue (descriptors, records, first bulk copies of the
    // matrix stream) while this kernel publishes and waits at the barrier; it does not touch x or y
    // before its griddepcontrol.wait, i.e. before this grid -- barrier included -- has completed
    asm volatile("griddepcontrol.launch_dependents;");
    const bool publish_empty = (pub.mode & 2) == 0;
    const bool aliased = (pub.mode & 4) != 0;
    double* next_y = pub.clear_next ? pub.clear_next : y;
    const int32_t n = n_boundary + (publish_empty ? n_empty : 0);
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const bool b = i < n_boundary;
        const int32_t row = b ? boundary[i] : empty[i - n_boundary];
        if (row == 0) continue; // the phantom row has no global counterpart
        double v = 0.0;
        if (b) {
            v = y[row];
            next_y[row] = 0.0;
        } else if (aliased) {
            y[row] = 0.0;       // never-written row of my own slice: drop whatever the buffer held
            next_y[row] = 0.0;
        }
        const int64_t g = pub.row_offset + row;
        const uint32_t nb = pub.needs ? pub.needs[row] : 0xffu;
#pragma unroll
        for (int p = 0; p < CVR_MAX_PEERS; p++)
            if (p < pub.n_dst && ((nb >> p) & 1u)) pub.dst[p][g] = v;
    }
    // ---- last block: flag barrier (see cvr_peer_barrier_kernel)
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) *done_counter = 0u;
    const int p = threadIdx.x;
    if (p >= bar.n_ranks) return;
    __threadfence_system();
    volatile uint32_t* remote = bar.flags[p] + bar.rank;
    *remote = bar.epoch;
    volatile uint32_t* mine = bar.flags[bar.rank] + p;
    const long long t0 = clock64();
    while ((int32_t)(*mine - bar.epoch) < 0) {
        if (clock64() - t0 > 6000000000LL) break; // ~3 s: a lost peer must not hang the GPU
    }
    __threadfence_system();
}

// All-to-all flag barrier over peer-mapped memory: every rank writes `epoch` into its slot of every
// rank's flag array (after a system-scope fence, so the rows it published are visible first) and
// waits until all of its own slots carry the epoch.  Bounded spin: a lost peer must not hang the GPU.
__global__ void cvr_peer_barrier_kernel(const __grid_constant__ CvrBarrier b)
{
    const int p = threadIdx.x;
    if (p >= b.n_ranks) return;
    __threadfence_system();
    volatile uint32_t* remote = b.flags[p] + b.rank;
    *remote = b.epoch;
    volatile uint32_t* mine = b.flags[b.rank] + p;
    const long long t0 = clock64();
    while ((int32_t)(*mine - b.epoch) < 0) {
        if (clock64() - t0 > 6000000000LL) break; // ~3 s
    }
    __threadfence_system();
}

enum class SpmvKernel { Tma, Ldg, Window };

SpmvKernel selected_kernel()
{
    // CVR_SPMV_KERNEL = tma (default) | ldg | window: the earlier generations stay selectable so
    // that the choice can be re-measured (profiles/ holds the ncu captures of each)
    const char* e = getenv("CVR_SPMV_KERNEL"); // read per call: tools/kernel_ab.py switches it at run time
    if (e && strcmp(e, "window") == 0) return SpmvKernel::Window;
    if (e && strcmp(e, "ldg") == 0) return SpmvKernel::Ldg;
    return SpmvKernel::Tma;
}

} // namespace

void cvr_preload_spmv_kernels()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<true, TB_TMA, false>);
    cudaFuncGetAttributes(&a, cvr_clear_rows_kernel);
}

// resident warps per SM of the selected kernel (used to size the automatic chunk count)
int cvr_spmv_resident_warps_per_sm()
{
    int blocks = 0;
    const SpmvKernel k = selected_kernel();
    cudaError_t e;
    if (k == SpmvKernel::Window)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_window_kernel, WARPS * 32, 0);
    else if (k == SpmvKernel::Ldg)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_tile_kernel<false, TB_LDG, false>,
                                                          WARPS * 32, 0);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_tile_kernel<true, TB_TMA, false>,
                                                          WARPS * 32, Geo<TB_TMA>::DYN_SMEM);
    if (e != cudaSuccess || blocks <= 0) return 32;
    return blocks * WARPS;
}

int cvr_launch_spmv(const CvrChunk* chunks, int32_t n_chunks, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, const CvrRowLists& rows, const CvrPublish* publish,
                    cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end,
                    const CvrBarrier* barrier, unsigned int* done_counter, bool y_is_clear)
{
    int launched = 0;
    const SpmvKernel k = selected_kernel();
    const int32_t n_clear = rows.n_boundary + rows.n_empty;
    bool after_clear_kernel = false;
    if (y_is_clear) {
        // the previous iteration's epilogue kernel already cleared the accumulated rows
    } else if (rows.boundary && k != SpmvKernel::Window) {
        const int cb = (n_clear + 255) / 256;
        cvr_clear_rows_kernel<<<cb < 1184 ? (cb < 1 ? 1 : cb) : 1184, 256, 0, stream>>>(
            y, rows.boundary, rows.n_boundary, rows.empty, rows.n_empty, publish && (publish->mode & 4));
        launched++;
        after_clear_kernel = true;
    } else if (cudaMemsetAsync(y, 0, sizeof(double) * (size_t)(n_rows + 1), stream) != cudaSuccess) {
        return -1;
    }
    const int threads = WARPS * 32;
    const int blocks = (int)(((int64_t)n_chunks * 32 + threads - 1) / threads);
    // the tile kernels are persistent: one block per resident slot, warps stride over the chunks
    int resident_blocks = 0;
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident_blocks = sms * (cvr_spmv_resident_warps_per_sm() / WARPS);
    }
    const int pblocks = blocks < resident_blocks ? blocks : resident_blocks;
    CvrPublish none{};
    const bool pub = publish && publish->n_dst > 0;
    if (pub && k == SpmvKernel::Window) return -1;
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    if (k == SpmvKernel::Window)
        cvr_spmv_window_kernel<<<blocks, threads, 0, stream>>>(chunks, n_chunks, vals, cols, record, x, y);
    else if (k == SpmvKernel::Ldg) {
        if (pub)
            cvr_spmv_tile_kernel<false, TB_LDG, true><<<pblocks, threads, 0, stream>>>(
                chunks, n_chunks, vals, cols, record, x, y, *publish);
        else
            cvr_spmv_tile_kernel<false, TB_LDG, false><<<pblocks, threads, 0, stream>>>(
                chunks, n_chunks, vals, cols, record, x, y, none);
    } else {
        if (pub) {
            // iterated SpMV: from the second iteration on the kernel in front of us is the previous
            // iteration's epilogue (publish + barrier): launch as its programmatic dependent
            static const bool use_pdl_pub = [] {
                const char* e = getenv("CVR_NO_PDL");
                return !(e && *e == '1');
            }();
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(pblocks);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = Geo<TB_TMA>::DYN_SMEM;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = (use_pdl_pub && (y_is_clear || after_clear_kernel) && !ev_begin) ? 1 : 0;
            cudaLaunchKernelEx(&cfg, cvr_spmv_tile_kernel<true, TB_TMA, true>, chunks, n_chunks, vals, cols, record,
                               x, y, *publish);
        } else {
            // ordinary SpMV: launch the sweep as a programmatic dependent of the clearing kernel
            static const bool use_pdl = [] {
                const char* e = getenv("CVR_NO_PDL");
                return !(e && *e == '1');
            }();
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(pblocks);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = Geo<TB_TMA>::DYN_SMEM;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = (use_pdl && after_clear_kernel && !ev_begin) ? 1 : 0;
            cudaLaunchKernelEx(&cfg, cvr_spmv_tile_kernel<true, TB_TMA, false>, chunks, n_chunks, vals, cols,
                               record, x, y, none);
        }
    }
    launched++;
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (pub) {
        if (!barrier || !done_counter) return -1;
        const int cb = (n_clear + 255) / 256;
        cvr_publish_epilogue_kernel<<<cb < 592 ? (cb < 1 ? 1 : cb) : 592, 256, 0, stream>>>(
            y, rows.boundary, rows.n_boundary, rows.empty, rows.n_empty, *publish, *barrier, done_counter);
        launched++;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launched;
}

namespace {
__global__ void cvr_column_footprint_kernel(const int32_t* __restrict__ cols, int64_t nnz,
                                            uint8_t* __restrict__ used)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        used[cols[i]] = 1; // benign race: every writer stores the same byte
}
} // namespace

namespace {
__global__ void cvr_chunk_needs_kernel(const CvrChunk* __restrict__ chunks, int32_t T,
                                       const uint8_t* __restrict__ needs, uint8_t* __restrict__ chunk_any)
{
    const int32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chunk >= T) return;
    const int t = threadIdx.x & 31;
    unsigned any = 0;
    for (int32_t r = chunks[chunk].first_row + t; r <= chunks[chunk].last_row; r += 32) any |= needs[r];
    any = __reduce_or_sync(0xffffffffu, any);
    if (t == 0) chunk_any[chunk] = (uint8_t)any;
}
} // namespace

int cvr_launch_chunk_needs(const CvrChunk* chunks, int32_t n_chunks, const uint8_t* needs, uint8_t* chunk_any,
                           cudaStream_t stream)
{
    cvr_chunk_needs_kernel<<<(n_chunks * 32 + 127) / 128, 128, 0, stream>>>(chunks, n_chunks, needs, chunk_any);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int cvr_launch_column_footprint(const int32_t* cols, int64_t nnz, uint8_t* used, cudaStream_t stream)
{
    cvr_column_footprint_kernel<<<148 * 16, 256, 0, stream>>>(cols, nnz, used);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int cvr_launch_peer_barrier(const CvrBarrier& b, cudaStream_t stream)
{
    cvr_peer_barrier_kernel<<<1, 32, 0, stream>>>(b);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
