// y = A * x over the CVR arrays on the device (sm_100a), fp64.
//
// Replaces the five-phase streaming loop of spmv_compute_kernel
// (/root/reference/spmv.cpp:1016-1667) and implements its INTENDED semantics
// (paper Alg. 4), including the steal-only chunks (split1 == -1) the
// reference kernel mishandles.
//
// Two sweep kernels live in this file (CVR_SPMV_KERNEL = pipe | tile selects one per launch; both
// pass the same parity suite, tests/test_gpu_parity.py runs it over both):
//   cvr_spmv_tile_kernel   round 1: tile walker, vals/cols staged by the TMA engine, one tile in flight
//                          per warp (profiles/r01_*)
//   cvr_spmv_pipe_kernel   round 2, the default: the same walker, software-pipelined -- the x gathers
//                          of tile k+1 are in flight while tile k is walked, the TMA ring runs two
//                          tiles ahead and crosses chunk boundaries, record batches are prefetched
//                          (profiles/r02_*).  <..., kPublish> variants additionally push finished
//                          rows to peer GPUs for the iterated multi-GPU SpMV.
// The first two generations of round 1 (a window kernel with a warp-wide segmented reduction and the
// walker with LDG-streamed operands) were re-measured against these on B200 (profiles/
// r02_kernel_ab_round1_generations.jsonl: never faster on any workload) and removed.
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
#include "cvr_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// Tile walker.
//
// A warp owns one chunk at a time; a pass covers a TILE of 4*TB steps x 8 lanes and thread
// t = (q, l) walks TB CONSECUTIVE steps of SIMD lane l: steps TB*q .. TB*q+TB-1 of the tile.
// A row switch is then a thread-local event (emit the accumulator, clear it); threads of the same
// SIMD lane only meet once per tile, in a 3-shuffle carry chain that hands the open partial sum
// from walker q to walker q+1.
//   * the vals/cols stream never passes the LSU: per tile ONE lane arms an mbarrier and issues two
//     cp.async.bulk copies (global -> shared, L2 evict-first); the walkers read their operands from
//     the stage with conflict-free 64-bit / 32-bit shared loads (see TB below).
//   * records are delivered to their owner thread through shared memory: the warp holds 32
//     records in registers (coalesced 256 B load), each holder drops a flag byte and the
//     write-back target into the owner's slot.
// ---------------------------------------------------------------------------------------
// TB (consecutive steps per walker) is ODD on purpose: a walker's quarter of the tile is then
// TB x 64 B of vals and TB x 32 B of cols, so the four walkers start 16 (vals) / 8 (cols)
// shared-memory banks apart and their 64-bit / 32-bit loads of one step are conflict-free in the
// plain linear layout ONE bulk copy per array produces.  (With TB = 8 the quarters alias on the
// same banks; padding each quarter separately needed 8 small copies per tile.)
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

// Destinations of one published row: the footprint bits when the exchange is sparse, otherwise
// every destination -- EXCEPT this GPU's own buffer when y already IS its slice of the next x
// (mode bit 2): re-storing a stale copy of y[r] over itself would race with the neighbouring
// chunks' atomicAdds on shared rows.
__device__ __forceinline__ uint32_t publish_mask(const CvrPublish& pub, int32_t row)
{
    if (pub.needs) return pub.needs[row];
    return (pub.mode & 4) ? (0xffu & ~(1u << pub.self)) : 0xffu;
}

// A finished row that this chunk owns alone: one plain store (spmv.cpp:1204).
template <bool kPublish>
__device__ __forceinline__ void store_row(const TileCtx& cx, int32_t row, double value)
{
    cx.y[row] = value;
    if (kPublish && cx.scatter) { // scattered 8-byte peer stores, row by row
        const int64_t g = cx.pub->row_offset + row;
        const uint32_t nb = publish_mask(*cx.pub, row);
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
            nb[u] = !in ? 0u : publish_mask(pub, r0 + 32 * u);
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

#ifndef CVR_TMA_BLOCKS
#define CVR_TMA_BLOCKS 6
#endif

template <int TB, int NB, bool kPublish>
__global__ void __launch_bounds__(WARPS * 32, NB)
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
    extern __shared__ __align__(128) unsigned char s_stream[]; // WARPS x STAGES tiles

    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = t >> 3, l = t & 7;
    const int32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t n_warps = (gridDim.x * blockDim.x) >> 5;

    // ---- per-warp TMA ring, set up once: the warp is persistent and walks chunks
    // warp0, warp0 + n_warps, ... (chunks are nnz-balanced, so a static round robin is even)
    // iterated multi-GPU SpMV: the publish epilogue behind us is a programmatic dependent; let its
    // blocks become resident as ours retire (it waits for this grid to complete before it reads y)
    if (kPublish) asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t ring = smem_u32(s_stream + w * G::WARP_SMEM);
    const uint32_t bar0 = smem_u32(&s_bar[w][0]);
    uint64_t policy = 0;
    uint32_t n_issued = 0, n_waited = 0; // tiles issued to / consumed from the ring, all chunks
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (t == 0) {
#pragma unroll
        for (int sidx = 0; sidx < STAGES; sidx++) mbar_init(bar0 + 8u * sidx, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
#pragma unroll
        for (int sidx = 0; sidx < STAGES; sidx++)
            if (sidx < n_tiles) issue_tile(sidx);

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
            {
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


// ---------------------------------------------------------------------------------------
// cvr_spmv_pipe_kernel -- the walker above, software-pipelined (round 2, the default).
//
// Why: tools/probe (profiles/r02_gather_probe_summary.txt) shows that on matrices whose x gather
// misses L1 (R-MAT, web) one SM retires at most ~1 gathered element per clock however many warps or
// loads are in flight -- the L1 miss path, not HBM, is the ceiling (R-MAT-24: 0.94 ms for "stream +
// gather + FMA" against 1.40 ms for the round-1 sweep).  The round-1 sweep loses the difference in
// DUTY CYCLE: a warp issues the 288 gathers of a tile, waits for them, and only then delivers
// records, walks, emits and synchronises -- all of that with no gather of its own outstanding.
// Here the warp keeps the miss path fed from one tile to the next:
//   * the x gathers of tile g+1 are issued BEFORE tile g is walked (xv / xv_next registers);
//   * the TMA ring is two stages deep and runs two tiles ahead; the fetch side needs no
//     descriptor (chunk start / length are closed-form in (chunk, nnz, T), spmv.cpp:584-627), so it
//     crosses chunk boundaries: the first tiles of the warp's next chunk are already in flight
//     while the current chunk's tail is walked and its epilogue runs;
//   * record batches are prefetched one batch (32 records) ahead, so the delivery loop does not
//     wait on a dependent global load per batch.
// Per tile g (stage s = g & 1):
//   1. wait full[s^1]; read cols of tile g+1 from the stage; issue its gathers -> xv_next
//   2. read vals of tile g (its barrier was waited for one iteration earlier)
//   3. fence.proxy.async + syncwarp; lane 0 re-arms stage s with tile g+2
//   4. deliver records, walk, carry chain, emits of tile g (chunk prologue / epilogue around it)
//   5. xv <- xv_next
// ---------------------------------------------------------------------------------------
template <int TB>
struct PipeGeo {
    static constexpr int TILE = 4 * TB * CVR_W;
    static constexpr int QUARTER = TB * CVR_W;
    static constexpr int FLAG_WORDS = (TB + 3) / 4;
    static constexpr int STAGE_BYTES = TILE * 12;        // vals then cols
    static constexpr int WARP_SMEM = 2 * STAGE_BYTES;    // two stages per warp
    static constexpr int DYN_SMEM = WARPS * WARP_SMEM;
};

template <int TB, int NB, bool kPublish>
__global__ void __launch_bounds__(WARPS * 32, NB)
cvr_spmv_pipe_kernel(const CvrChunk* __restrict__ chunks, int32_t T, int64_t nnz,
                     const double* __restrict__ vals, const int32_t* __restrict__ cols,
                     const int32_t* __restrict__ record, const double* __restrict__ x,
                     double* __restrict__ y, const __grid_constant__ CvrPublish pub)
{
    using G = PipeGeo<TB>;
    constexpr int TILE = G::TILE, QUARTER = G::QUARTER, FLAG_WORDS = G::FLAG_WORDS;
    __shared__ uint32_t s_flags[WARPS][FLAG_WORDS][32]; // one flag byte per (thread, step)
    __shared__ int32_t s_wb[WARPS][TB][32];             // write-back target per (step, thread)
    __shared__ __align__(8) unsigned long long s_bar[WARPS][2];
    extern __shared__ __align__(128) unsigned char s_stream[]; // WARPS x 2 stages

    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = t >> 3, l = t & 7;
    const int32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    // iterated multi-GPU SpMV: the publish epilogue behind us is a programmatic dependent; let its
    // blocks become resident as ours retire (it waits for this grid to complete before it reads y)
    if (kPublish) asm volatile("griddepcontrol.launch_dependents;");
    if (warp0 >= T) return; // no chunk for this warp (never with the launcher's grid)

    const uint32_t ring = smem_u32(s_stream + w * G::WARP_SMEM);
    const uint32_t bar0 = smem_u32(&s_bar[w][0]);
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (t == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll
    for (int k = 0; k < FLAG_WORDS; k++) s_flags[w][k][t] = 0u;
    __syncwarp();

    // ---- fetch side: the warp's tile sequence (chunks warp0, warp0 + n_warps, ...; every tile of each)
    // in closed form -- nnz-balanced slices in multiples of 16 (spmv.cpp:584-586, :615-627)
    const int64_t per = (nnz / T / 16) * 16;
    const int64_t brk = (nnz - per * T) / 16;
    int32_t f_chunk = warp0, f_tile = 0, f_ntiles = 0, f_len = 0;
    int64_t f_start = 0;
    auto fetch_enter = [&](int32_t c) {
        f_chunk = c;
        f_tile = 0;
        if (c < T) {
            f_start = c < brk ? c * (per + 16) : c * per + brk * 16;
            f_len = (int32_t)(c == T - 1 ? nnz - f_start : (c < brk ? per + 16 : per));
            f_ntiles = (f_len + TILE - 1) / TILE;
        }
    };
    // start the bulk copies of the next tile of the sequence into stage `sidx`; returns its element
    // count (0: the sequence is exhausted)
    auto fetch_issue = [&](uint32_t sidx) -> int32_t {
        if (f_chunk >= T) return 0;
        const int32_t ts = f_tile * TILE;
        const int32_t n_el = min(TILE, f_len - ts); // multiple of 16
        if (t == 0) {
            const uint32_t bar = bar0 + 8u * sidx;
            const uint32_t stage = ring + sidx * G::STAGE_BYTES;
            mbar_expect_tx(bar, (uint32_t)n_el * 12u);
            bulk_g2s(stage, vals + f_start + ts, (uint32_t)n_el * 8u, bar, policy);
            bulk_g2s(stage + TILE * 8, cols + f_start + ts, (uint32_t)n_el * 4u, bar, policy);
        }
        if (++f_tile == f_ntiles) fetch_enter(f_chunk + n_warps);
        return n_el;
    };
    fetch_enter(warp0);
    int32_t nel0 = fetch_issue(0); // tile being walked
    int32_t nel1 = fetch_issue(1); // tile whose gathers are in flight
    int32_t nel2 = 0;              // tile just issued to the ring

    // cols of a tile -> gathers of x (columns of elements past the chunk end read the phantom x[0])
    auto gather_tile = [&](uint32_t sidx, int32_t n_el, double (&out)[TB]) {
        const uint32_t* sc =
            reinterpret_cast<const uint32_t*>(s_stream + w * G::WARP_SMEM + sidx * G::STAGE_BYTES + TILE * 8) +
            q * QUARTER + l;
        uint32_t ci[TB];
        if (n_el == TILE) {
#pragma unroll
            for (int b = 0; b < TB; b++) ci[b] = sc[b * CVR_W];
        } else {
#pragma unroll
            for (int b = 0; b < TB; b++) ci[b] = (q * QUARTER + l + b * CVR_W < n_el) ? sc[b * CVR_W] : 0u;
        }
#pragma unroll
        for (int b = 0; b < TB; b++) out[b] = __ldg(x + ci[b]);
    };

    // ---- walk side: per-chunk state
    int32_t chunk = warp0, w_tile = 0, w_ntiles = 1;
    const CvrChunk* cp = chunks + chunk;
    int32_t split0 = 0, n_rec = 0, chunk_last_row = 0;
    const int2* rec = nullptr;
    int32_t rb = 0;
    int2 held = make_int2(-1, 0), nxt = make_int2(-1, 0);
    double lane_carry = 0.0, carry_slot = 0.0;
    TileCtx cx;
    cx.y = y;
    cx.pub = &pub;
    cx.l = l;
    cx.tail = cp->tail;
    cx.scatter = false;
    cx.split1 = 0;
    cx.first_row = 0;

    // When launched as a programmatic dependent (of the clearing kernel, or of the previous iteration's
    // publish epilogue) everything above overlapped the predecessor's tail; x may only be read and y
    // written once that grid has completed.  No-op for an ordinary launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    double xv[TB], xv_next[TB];
    mbar_wait(bar0, 0u);
    gather_tile(0u, nel0, xv);

    for (uint32_t g = 0;; g++) {
        const uint32_t sidx = g & 1u;
        // ---- 0. chunk prologue: descriptor and the first two record batches (their latency overlaps
        // the barrier wait and the gathers below)
        if (w_tile == 0) {
            cp = chunks + chunk;
            const int32_t len = cp->len;
            split0 = cp->split0;
            n_rec = cp->n_rec;
            chunk_last_row = cp->last_row;
            cx.tail = cp->tail;
            cx.split1 = cp->split1;
            cx.first_row = cp->first_row;
            // A chunk in a very sparse region can span 10^5 (mostly empty) rows: pushing that range from
            // one warp would serialise; such chunks publish their few finished rows one by one instead
            cx.scatter = kPublish && ((pub.mode & 1) || chunk_last_row - cx.first_row >= PUSH_MAX_ROWS);
            w_ntiles = (len + TILE - 1) / TILE;
            rec = reinterpret_cast<const int2*>(record + cvr_record_offset(chunk, cx.first_row));
            rb = 0;
            held = (t < n_rec) ? rec[t] : make_int2(-1, 0);
            nxt = (32 + t < n_rec) ? rec[32 + t] : make_int2(-1, 0);
            lane_carry = 0.0;
            carry_slot = 0.0;
        }
        // ---- 1. operands: cols of tile g+1 (its gathers go out below) and vals of tile g, both from the ring
        const int32_t ts = w_tile * TILE;
        const int32_t p0 = ts + q * QUARTER + l; // my element of step TB*q of the tile; next step: +8
        uint32_t ci[TB];
        if (nel1 > 0) {
            mbar_wait(bar0 + 8u * (sidx ^ 1u), ((g + 1u) >> 1) & 1u);
            const uint32_t* sc = reinterpret_cast<const uint32_t*>(s_stream + w * G::WARP_SMEM +
                                                                   (sidx ^ 1u) * G::STAGE_BYTES + TILE * 8) +
                                 q * QUARTER + l;
            if (nel1 == TILE) {
#pragma unroll
                for (int b = 0; b < TB; b++) ci[b] = sc[b * CVR_W];
            } else {
#pragma unroll
                for (int b = 0; b < TB; b++) ci[b] = (q * QUARTER + l + b * CVR_W < nel1) ? sc[b * CVR_W] : 0u;
            }
        }
        double a[TB];
        {
            const double* sv =
                reinterpret_cast<const double*>(s_stream + w * G::WARP_SMEM + sidx * G::STAGE_BYTES) + q * QUARTER + l;
            if (nel0 == TILE) {
#pragma unroll
                for (int b = 0; b < TB; b++) a[b] = sv[b * CVR_W];
            } else {
#pragma unroll
                for (int b = 0; b < TB; b++) a[b] = (q * QUARTER + l + b * CVR_W < nel0) ? sv[b * CVR_W] : 0.0;
            }
        }
        // ---- 2. stage `sidx` is free (cols read one iteration ago, vals just now): refill it with tile
        // g+2.  The refill is an async-proxy write to memory these generic-proxy loads just read: fence in
        // every reader, converge, then re-arm.  The fence (MEMBAR + FENCE.VIEW.ASYNC) waits for every
        // memory operation of the thread that is still in flight, so it has to sit BEFORE the gathers go
        // out -- behind them it would wait for them and serialise the pipeline (ncu: 8 % of the stall
        // samples of the first version, profiles/r02_prof_pipe9x4_v1_rmat24_source_top.txt).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        nel2 = fetch_issue(sidx);
        // ---- 3. gathers of tile g+1: in flight while tile g is walked
        if (nel1 > 0) {
#pragma unroll
            for (int b = 0; b < TB; b++) xv_next[b] = __ldg(x + ci[b]);
        }

        // ---- 4a. deliver the records of this tile to their owner threads
        for (;;) {
            const uint32_t rel = (uint32_t)(held.x - ts);
            if (rel < (uint32_t)TILE) {
                const uint32_t step = rel >> 3, wq = step / TB, b = step - wq * TB;
                const uint32_t owner = wq * CVR_W + (rel & 7u);
                reinterpret_cast<unsigned char*>(&s_flags[w][b >> 2][owner])[b & 3u] = 1;
                s_wb[w][b][owner] = held.y;
            }
            // records are sorted by position: the batch reaches past the tile (or the list ended, pos = -1)
            // as soon as ANY lane holds a position beyond it -- a vote, not a shuffle round trip through
            // the load/store queue the gathers occupy
            if (__any_sync(FULL, (uint32_t)held.x >= (uint32_t)(ts + TILE))) break;
            rb += 32;
            held = nxt;
            nxt = (rb + 32 + t < n_rec) ? rec[rb + 32 + t] : make_int2(-1, 0);
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

        // ---- 4b. walk my TB steps with two predicated FMA chains: `head` collects the steps before my
        // first flag (everything if I have none), `tail` the steps from my last flag on.
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

        // ---- 4c. carry chain over the four walkers of my SIMD lane.  Walker r hands on
        // out_r = has_r ? tail_r : cin_r + head_r, cin_r = out_(r-1), cin_0 = the lane's carry from the
        // previous tile.  With v_r = has_r ? tail_r : head_r that is out_r = v_r + (has_r ? 0 : cin_r): every
        // thread fetches the three other v of its SIMD lane with INDEPENDENT shuffles (they pipeline through
        // the load/store queue; the dependent shuffle chain of round 1 cost three queue round trips) and the
        // has-flags with one vote, then folds the chain locally.
        const bool has = mask != 0u;
        const unsigned has_all = __ballot_sync(FULL, has);
        const double v_mine = has ? tailsum : head;
        double vq[4];
#pragma unroll
        for (int r = 0; r < 4; r++) vq[r] = __shfl_sync(FULL, v_mine, r * CVR_W + l);
        double run = lane_carry, cin = lane_carry; // run = out_(r-1) while folding
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (r == q) cin = run;
            run = ((has_all >> (r * CVR_W + l)) & 1u) ? vq[r] : run + vq[r];
        }
        lane_carry = run;
        if (has) emit<kPublish>(cx, head + cin, p0 + b_first * CVR_W, s_wb[w][b_first][t], carry_slot);
        __syncwarp(); // slots are reused by the next tile's delivery

        // ---- 4d. chunk epilogue: lane remainders through the eight pos=-1 records (spmv.cpp:1633-1649)
        if (++w_tile == w_ntiles) {
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
            chunk += n_warps;
            w_tile = 0;
        }

        // ---- 5. rotate the pipeline
        if (nel1 == 0) break;
#pragma unroll
        for (int b = 0; b < TB; b++) xv[b] = xv_next[b];
        nel0 = nel1;
        nel1 = nel2;
    }
}

constexpr int TB_TILE = 9; // round-1 kernel: tile = 288 elements
static_assert(TB_TILE % 2 == 1, "the TMA tile needs an odd number of steps per walker (bank layout)");

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

// All-to-all flag barrier over peer-mapped memory, run by the first n_ranks threads of one block:
// every rank writes `epoch` into its slot of every rank's flag array (after a system-scope fence,
// so the rows it published are visible first) and waits until all of its own slots carry the epoch.
// The spin is bounded (a lost peer must not hang the GPU) -- and a timeout is NOT silent: the epoch
// is recorded in *b.error, which cvr_check_async_error turns into CVR_ERR_STATE on the host.
__device__ __forceinline__ void peer_flag_barrier(const CvrBarrier& b, int p)
{
    if (p >= b.n_ranks) return;
    __threadfence_system();
    volatile uint32_t* remote = b.flags[p] + b.rank;
    *remote = b.epoch;
    volatile uint32_t* mine = b.flags[b.rank] + p;
    const long long t0 = clock64();
    while ((int32_t)(*mine - b.epoch) < 0) {
        if (clock64() - t0 > b.timeout_cycles) {
            if (b.error) atomicCAS(b.error, 0u, b.epoch); // first failing epoch wins
            break;
        }
    }
    __threadfence_system();
}

// After the sweep of an iterated multi-GPU SpMV, ONE epilogue kernel does everything that has to
// wait for the sweep to be complete:
//   1. the accumulated rows are final now: publish them; rows nothing writes get an explicit 0.0
//      (only while `publish_empty`: a reused x buffer must not keep a stale value there -- two
//      iterations cover both buffers);
//   2. clear those rows of y again, ready for the next sweep (cvr_launch_spmv then skips its own
//      clearing kernel);
//   3. the last block to finish runs the all-to-all flag barrier over peer memory.
// It is launched as a programmatic dependent of the sweep (its blocks are resident and waiting when
// the sweep's last warp retires -- no launch gap), and the next iteration's sweep as a programmatic
// dependent of it.
__global__ void cvr_publish_epilogue_kernel(double* __restrict__ y, const int32_t* __restrict__ boundary,
                                            int32_t n_boundary, const int32_t* __restrict__ empty,
                                            int32_t n_empty, const __grid_constant__ CvrPublish pub,
                                            const __grid_constant__ CvrBarrier bar, unsigned int* done_counter)
{
    // the next iteration's sweep may start its prologue (barrier set-up, first bulk copies of the
    // matrix stream) while this kernel publishes and waits at the barrier; it does not touch x or y
    // before its griddepcontrol.wait, i.e. before this grid -- barrier included -- has completed
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory"); // the sweep in front of us is complete and visible
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
        const uint32_t nb = publish_mask(pub, row);
#pragma unroll
        for (int p = 0; p < CVR_MAX_PEERS; p++)
            if (p < pub.n_dst && ((nb >> p) & 1u)) pub.dst[p][g] = v;
    }
    // ---- last block: flag barrier
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) *done_counter = 0u;
    peer_flag_barrier(bar, threadIdx.x);
}

__global__ void cvr_peer_barrier_kernel(const __grid_constant__ CvrBarrier b)
{
    peer_flag_barrier(b, threadIdx.x);
}

// ---- sweep variants.  `pipeTxB` = cvr_spmv_pipe_kernel<T, B>: T steps per walker (tile = 32 T
// elements), B resident blocks (4 warps each) per SM requested through __launch_bounds__.
struct Variant {
    const char* name;
    bool pipe;
    int tb, nb;
};
constexpr Variant VARIANTS[] = {
    {"tile", false, TB_TILE, CVR_TMA_BLOCKS},
    {"pipe9x4", true, 9, 4},
    {"pipe7x4", true, 7, 4},
    {"pipe7x5", true, 7, 5},
    {"pipe5x5", true, 5, 5},
    {"pipe5x6", true, 5, 6},
    {"pipe9x3", true, 9, 3},
    {"tile7x7", true, 7, 7},
    {"tile7x8", true, 7, 8},
    {"tile11x5", true, 11, 5},
    {"tile13x4", true, 13, 4},
    {"tile5x8", true, 5, 8},
};
constexpr int N_VARIANTS = sizeof(VARIANTS) / sizeof(VARIANTS[0]);
#ifndef CVR_DEFAULT_VARIANT
#define CVR_DEFAULT_VARIANT 1
#endif

int selected_variant()
{
    // CVR_SPMV_KERNEL = tile | pipe (= the default) | pipe<T>x<B>; read per call so that
    // tools/kernel_ab.py and the parity suite can switch it at run time
    const char* e = getenv("CVR_SPMV_KERNEL");
    if (!e || !*e || strcmp(e, "pipe") == 0) return CVR_DEFAULT_VARIANT;
    for (int v = 0; v < N_VARIANTS; v++)
        if (strcmp(e, VARIANTS[v].name) == 0) return v;
    return CVR_DEFAULT_VARIANT;
}

template <typename K, typename... Args>
cudaError_t launch_ex(K kernel, int blocks, int threads, int smem, cudaStream_t stream, bool programmatic,
                      Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = programmatic ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// experimental geometries of the round-1 kernel (plain flavour only): "tile<T>x<B>"
template <int TB, int NB>
struct TileOps {
    static int resident_blocks()
    {
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_tile_kernel<TB, NB, false>, WARPS * 32,
                                                          Geo<TB>::DYN_SMEM) != cudaSuccess)
            return 0;
        return blocks;
    }
    static cudaError_t launch(bool publish, int blocks, cudaStream_t stream, bool programmatic,
                              const CvrChunk* chunks, int32_t T, int64_t, const double* vals,
                              const int32_t* cols, const int32_t* record, const double* x, double* y,
                              const CvrPublish& pub)
    {
        if (publish) return cudaErrorNotSupported;
        return launch_ex(cvr_spmv_tile_kernel<TB, NB, false>, blocks, WARPS * 32, Geo<TB>::DYN_SMEM, stream,
                         programmatic, chunks, T, vals, cols, record, x, y, pub);
    }
    static void preload()
    {
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<TB, NB, false>);
    }
};

// one entry per variant: occupancy query and launch, plain and publishing flavour
template <int TB, int NB>
struct PipeOps {
    static int resident_blocks()
    {
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_pipe_kernel<TB, NB, false>, WARPS * 32,
                                                          PipeGeo<TB>::DYN_SMEM) != cudaSuccess)
            return 0;
        return blocks;
    }
    static cudaError_t launch(bool publish, int blocks, cudaStream_t stream, bool programmatic,
                              const CvrChunk* chunks, int32_t T, int64_t nnz, const double* vals,
                              const int32_t* cols, const int32_t* record, const double* x, double* y,
                              const CvrPublish& pub)
    {
        if (publish)
            return launch_ex(cvr_spmv_pipe_kernel<TB, NB, true>, blocks, WARPS * 32, PipeGeo<TB>::DYN_SMEM, stream,
                             programmatic, chunks, T, nnz, vals, cols, record, x, y, pub);
        return launch_ex(cvr_spmv_pipe_kernel<TB, NB, false>, blocks, WARPS * 32, PipeGeo<TB>::DYN_SMEM, stream,
                         programmatic, chunks, T, nnz, vals, cols, record, x, y, pub);
    }
    static void preload()
    {
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, cvr_spmv_pipe_kernel<TB, NB, false>);
    }
};

#define CVR_FOR_PIPE_VARIANT(v, expr)                   \
    switch (v) {                                        \
    case 1: { using P = PipeOps<9, 4>; expr; } break;   \
    case 2: { using P = PipeOps<7, 4>; expr; } break;   \
    case 3: { using P = PipeOps<7, 5>; expr; } break;   \
    case 4: { using P = PipeOps<5, 5>; expr; } break;   \
    case 5: { using P = PipeOps<5, 6>; expr; } break;   \
    case 6: { using P = PipeOps<9, 3>; expr; } break;   \
    case 7: { using P = TileOps<7, 7>; expr; } break;   \
    case 8: { using P = TileOps<7, 8>; expr; } break;   \
    case 9: { using P = TileOps<11, 5>; expr; } break;  \
    case 10: { using P = TileOps<13, 4>; expr; } break; \
    case 11: { using P = TileOps<5, 8>; expr; } break;  \
    default: break;                                     \
    }

// resident blocks per SM of a variant on the current device (cached per device and variant)
int variant_resident_blocks(int v)
{
    static int cache[64][N_VARIANTS] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int* slot = (dev >= 0 && dev < 64) ? &cache[dev][v] : nullptr;
    if (slot && *slot > 0) return *slot;
    int blocks = 0;
    if (!VARIANTS[v].pipe) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_tile_kernel<TB_TILE, CVR_TMA_BLOCKS, false>, WARPS * 32,
                                                          Geo<TB_TILE>::DYN_SMEM) != cudaSuccess)
            blocks = 0;
    } else {
        CVR_FOR_PIPE_VARIANT(v, blocks = P::resident_blocks())
    }
    if (blocks <= 0) blocks = 4;
    if (slot) *slot = blocks;
    return blocks;
}

int device_sm_count()
{
    static int cache[64] = {};
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (dev >= 0 && dev < 64) cache[dev] = sms;
    return sms;
}

bool pdl_enabled()
{
    const char* e = getenv("CVR_NO_PDL");
    return !(e && *e == '1');
}

} // namespace

void cvr_preload_spmv_kernels()
{
    cudaFuncAttributes a;
    const int v = selected_variant();
    if (!VARIANTS[v].pipe) cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<TB_TILE, CVR_TMA_BLOCKS, false>);
    else {
        CVR_FOR_PIPE_VARIANT(v, P::preload())
    }
    cudaFuncGetAttributes(&a, cvr_clear_rows_kernel);
}

// resident warps per SM of the selected kernel (used to size the automatic chunk count)
int cvr_spmv_resident_warps_per_sm()
{
    return variant_resident_blocks(selected_variant()) * WARPS;
}

const char* cvr_spmv_kernel_name()
{
    return VARIANTS[selected_variant()].name;
}

int cvr_launch_spmv(const CvrChunk* chunks, int32_t n_chunks, int64_t nnz, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, const CvrRowLists& rows, const CvrPublish* publish,
                    cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end,
                    const CvrBarrier* barrier, unsigned int* done_counter, bool y_is_clear)
{
    int launched = 0;
    const int v = selected_variant();
    const int sms = device_sm_count();
    const int32_t n_clear = rows.n_boundary + rows.n_empty;
    bool after_clear_kernel = false;
    if (y_is_clear) {
        // the previous iteration's epilogue kernel already cleared the accumulated rows
    } else if (rows.boundary) {
        const int cb = (n_clear + 255) / 256;
        const int cap = sms * 8;
        cvr_clear_rows_kernel<<<cb < cap ? (cb < 1 ? 1 : cb) : cap, 256, 0, stream>>>(
            y, rows.boundary, rows.n_boundary, rows.empty, rows.n_empty, publish && (publish->mode & 4));
        launched++;
        after_clear_kernel = true;
    } else if (cudaMemsetAsync(y, 0, sizeof(double) * (size_t)(n_rows + 1), stream) != cudaSuccess) {
        return -1;
    }
    const int threads = WARPS * 32;
    const int blocks = (int)(((int64_t)n_chunks * 32 + threads - 1) / threads);
    // the sweeps are persistent: one block per resident slot, warps stride over the chunks
    const int resident = sms * variant_resident_blocks(v);
    const int pblocks = blocks < resident ? blocks : resident;
    CvrPublish none{};
    const bool pub = publish && publish->n_dst > 0;
    // the kernel in front of us on the stream is the clearing kernel or (iterated SpMV, from the second
    // iteration on) the previous iteration's epilogue: launch as its programmatic dependent
    const bool programmatic = pdl_enabled() && (after_clear_kernel || (pub && y_is_clear)) && !ev_begin;
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    cudaError_t e = cudaSuccess;
    if (!VARIANTS[v].pipe) {
        if (pub)
            e = launch_ex(cvr_spmv_tile_kernel<TB_TILE, CVR_TMA_BLOCKS, true>, pblocks, threads, Geo<TB_TILE>::DYN_SMEM, stream,
                          programmatic, chunks, n_chunks, vals, cols, record, x, y, *publish);
        else
            e = launch_ex(cvr_spmv_tile_kernel<TB_TILE, CVR_TMA_BLOCKS, false>, pblocks, threads, Geo<TB_TILE>::DYN_SMEM, stream,
                          programmatic, chunks, n_chunks, vals, cols, record, x, y, none);
    } else {
        CVR_FOR_PIPE_VARIANT(v, e = P::launch(pub, pblocks, stream, programmatic, chunks, n_chunks, nnz, vals, cols,
                                              record, x, y, pub ? *publish : none))
    }
    if (e != cudaSuccess) return -1;
    launched++;
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (pub) {
        if (!barrier || !done_counter) return -1;
        const int cb = (n_clear + 255) / 256;
        const int cap = sms * 4;
        e = launch_ex(cvr_publish_epilogue_kernel, cb < cap ? (cb < 1 ? 1 : cb) : cap, 256, 0, stream,
                      pdl_enabled() && !ev_end, y, (const int32_t*)rows.boundary, rows.n_boundary,
                      (const int32_t*)rows.empty, rows.n_empty, *publish, *barrier, done_counter);
        if (e != cudaSuccess) return -1;
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
    cvr_column_footprint_kernel<<<device_sm_count() * 16, 256, 0, stream>>>(cols, nnz, used);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int cvr_launch_peer_barrier(const CvrBarrier& b, cudaStream_t stream)
{
    cvr_peer_barrier_kernel<<<1, 32, 0, stream>>>(b);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
