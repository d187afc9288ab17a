// y = A * x over the CVR arrays on the device (sm_100a), fp64.
//
// Replaces the five-phase streaming loop of spmv_compute_kernel
// (/root/reference/spmv.cpp:1016-1667) and implements its INTENDED semantics
// (paper Alg. 4), including the steal-only chunks (split1 == -1) the
// reference kernel mishandles.
//
// One sweep kernel, cvr_spmv_tile_kernel<TB, NB, RD, kPublish>, in two geometries picked per matrix (see
// cvr_pick_sweep_variant): TB = 7 steps per walker, 5 resident blocks per SM and a 3-deep cp.async record ring
// for everything but long regular rows, which take TB = 11 at 5 blocks; <..., kPublish> additionally pushes
// finished rows to peer GPUs for the iterated multi-GPU SpMV.  Two things set the geometry besides the walk
// itself (profiles/r02_kernel_ab_l1_capacity.txt, r02_kernel_ab_record_ring.txt): the x gather needs L1 for its
// misses in flight, so the blocks of an SM must fit the 100 KB shared-memory configuration (the kernel asks for
// it: the driver's own pick is 132 KB), and short-row matrices consume several 32-record batches per tile, which
// a ring of cp.async copies delivers without a dependent global load in the walk.
// What was measured against it on B200 in round 2 and removed
// again (never faster on any BASELINE workload; numbers under profiles/, code in the git history):
//   * round 1's window kernel (warp-wide segmented reduction) and LDG-streamed walker
//     (profiles/r02_kernel_ab_round1_generations.jsonl);
//   * a software-pipelined walker -- x gathers of tile k+1 in flight while tile k is walked, TMA ring
//     two tiles ahead and across chunk boundaries, record batches prefetched (commits 0d94e38, d52c6bb;
//     profiles/r02_kernel_ab_pipe_v1.jsonl, r02_kernel_ab_pipe_v2.jsonl, ncu r02_prof_pipe*): 0-20 %
//     SLOWER.  tools/probe shows why nothing of that kind can help: on matrices whose x gather misses L1
//     an SM retires ~1 gathered element per clock at ANY occupancy or depth (the L1 miss path), and every
//     further load/store-unit operation (shared-memory operands, flag delivery, shuffles, stores) queues
//     in front of the same unit -- the sweep is bound by load/store-unit work per element, not by latency
//     (profiles/r02_gather_probe_summary.txt).
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
constexpr int32_t PUSH_MIN_ROWS = 64;   // a warp publishes finished rows to the peers in ranges of at least this many
#ifndef CVR_TMA_STAGES
#define CVR_TMA_STAGES 1
#endif
constexpr int STAGES = CVR_TMA_STAGES; // TMA ring depth per warp (1: the registers are the 2nd buffer)
// Template parameter RD of the sweep: record batches (32 records, one per lane) are prefetched RD - 1 batches ahead
// with cp.async (LDGSTS) into a per-warp ring: lane t copies record 32k + t into its own slot and is the only
// thread that ever reads it, so the ring needs no barrier -- cp.async.wait_group orders a lane's copy before its own
// shared load.  RD = 0: one batch held in registers, the next one requested when the held one is taken.

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
    int32_t last_done;     // kPublish: latest row this thread has stored in the current chunk (rows finish in order)
    bool scatter;          // kPublish: rows of this chunk are mostly EMPTY: publish each finished row on its own
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

// One store to an NVSwitch multicast address: the switch replicates it into the buffer of every GPU bound to
// the multicast object (SASS: an ordinary STG.E.64 -- the address does the work).
__device__ __forceinline__ void multicast_store(double* mc_addr, double value)
{
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(mc_addr), "d"(value) : "memory");
}

// A finished row that this chunk owns alone: one plain store (spmv.cpp:1204).
template <bool kPublish>
__device__ __forceinline__ void store_row(TileCtx& cx, int32_t row, double value)
{
    cx.y[row] = value;
    if (kPublish) {
        cx.last_done = row;
        if (cx.scatter) { // 8-byte peer stores, row by row
            const int64_t g = cx.pub->row_offset + row;
            const uint32_t nb = publish_mask(*cx.pub, row);
#pragma unroll
            for (int p = 0; p < CVR_MAX_PEERS; p++)
                if (p < cx.pub->n_dst && ((nb >> p) & 1u)) cx.pub->dst[p][g] = value;
        }
    }
}

// Iterated multi-GPU SpMV: the rows a warp has FINISHED are published -- stored into the x vector every GPU
// reads in the NEXT iteration -- from inside the sweep, range by range, while the warp is still walking its
// chunk: contiguous rows, so the peer-mapped stores are coalesced 256 B warp stores over NVLink, and the
// transfer overlaps the sweep tile by tile.  Which rows are finished: the conversion hands rows to the eight
// SIMD lanes in increasing order and a lane works through its rows one after the other, so every row up to
// min over the lanes of (the latest row that lane has stored) is complete or empty -- the watermark the
// tile loop maintains with five shuffles per tile.  (Round 1 pushed a chunk's whole range at its end and fell
// back to per-row 8-byte peer stores for chunks spanning more than 1024 rows; on R-MAT-24 over 8 GPUs those
// scattered stores doubled the sweep of the row-heavy ranks, profiles/r02_strong_scaling_v1.txt.)
// Rows of a range that are still being accumulated (shared first/last rows, tail rows) are sent as they are
// and overwritten by the epilogue kernel once the sweep is complete (same source GPU, same address, stream
// order); empty rows are skipped (their `needs` byte is cleared, cvr_chunk_needs) -- the epilogue zeroes them
// everywhere during the first two iterations.
// Exception: a chunk in a very sparse region spans 10^4..10^6 rows of which a few hundred are not empty;
// walking such a range row by row costs more than it saves (measured: the row-heavy shard of R-MAT-24 on
// 2 GPUs went from 883 to 1304 us), so a chunk whose row span exceeds 4x the rows it finishes publishes
// each finished row on its own at the moment it is stored (`scatter`).
__device__ __forceinline__ void publish_rows(const TileCtx& cx, int32_t first_row, int32_t last_row, int t)
{
    __threadfence_block(); // this warp's own row stores (made by other lanes) before the re-read
    __syncwarp();
    const CvrPublish& pub = *cx.pub;
    if (pub.mc) { // every row of the range, one coalesced multicast store per 32 rows
        for (int32_t r0 = first_row + t; r0 <= last_row; r0 += 4 * 32) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (r0 + 32 * u <= last_row) ? __ldcg(cx.y + r0 + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (r0 + 32 * u <= last_row) multicast_store(pub.mc + pub.row_offset + r0 + 32 * u, v[u]);
        }
        return;
    }
    for (int32_t r0 = first_row + t; r0 <= last_row; r0 += 4 * 32) {
        double v[4];
        uint32_t nb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool in = r0 + 32 * u <= last_row;
            nb[u] = !in ? 0u : publish_mask(pub, r0 + 32 * u);
            v[u] = nb[u] ? __ldcg(cx.y + r0 + 32 * u) : 0.0;
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
__device__ __forceinline__ void emit(TileCtx& cx, double value, int32_t pos, int32_t wb,
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


template <int TB, int NB, int RD, bool kPublish>
__global__ void __launch_bounds__(WARPS * 32, NB)
cvr_spmv_tile_kernel(const CvrChunk* __restrict__ chunks, int32_t chunk_begin, int32_t T,
                     const double* __restrict__ vals, const int32_t* __restrict__ cols,
                     const int32_t* __restrict__ record, const double* __restrict__ x,
                     double* __restrict__ y, const __grid_constant__ CvrPublish pub, unsigned int* queue)
{
    using G = Geo<TB>;
    constexpr int TILE = G::TILE, QUARTER = G::QUARTER, FLAG_WORDS = G::FLAG_WORDS;
    __shared__ uint32_t s_flags[WARPS][FLAG_WORDS][32]; // one flag byte per (thread, step)
    __shared__ int32_t s_wb[WARPS][TB][32];             // write-back target per (step, thread)
    __shared__ __align__(8) unsigned long long s_bar[WARPS][STAGES];
    __shared__ __align__(16) int2 s_rec[WARPS][RD > 0 ? RD : 1][RD > 0 ? 32 : 1]; // record ring (cp.async)
    extern __shared__ __align__(128) unsigned char s_stream[]; // WARPS x STAGES tiles

    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = t >> 3, l = t & 7;
    const int32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t n_warps = (gridDim.x * blockDim.x) >> 5;

    // ---- per-warp TMA ring, set up once: the warp is persistent and walks chunks
    // warp0, warp0 + n_warps, ... (chunks are nnz-balanced, so a static round robin is even)
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

    // chunks [chunk_begin, T) of the matrix (the whole matrix unless the host pipelines the sweep in slabs).
    // A warp's first chunk is its own index; further chunks come in index order from an atomic counter when the
    // host passes a queue (requested at the start of a chunk, consumed at its end: the round trip is hidden),
    // otherwise by static round robin.  Chunks are nnz-balanced but not record-balanced: with the static order
    // the warps of R-MAT-24 finish up to 8 % apart (sm__warps_active 40.1 % of a possible 43.75 %).
    for (int32_t chunk = chunk_begin + warp0; chunk < T;) {
        int32_t next_chunk = chunk + n_warps;
        if (queue && t == 0) next_chunk = chunk_begin + n_warps + (int32_t)atomicAdd(queue, 1u);
        const CvrChunk* cp = chunks + chunk;
        const int64_t start = cp->start;
        const int32_t len = cp->len;
        const int32_t split0 = cp->split0;
        const int32_t n_rec = cp->n_rec;
        TileCtx cx;
        cx.y = y;
        cx.tail = cp->tail;
        cx.pub = &pub;
        const int32_t chunk_last_row = cp->last_row;
        const bool reads_elsewhere = kPublish && (!pub.chunk_any || pub.chunk_any[chunk]); // a peer reads some row
        // (with a multicast address a range push costs one store per 32 rows whatever the number of GPUs: always ranges)
        // experiment knobs in the upper mode bits (0 = default): bits 8-15 push_min_rows / 16, bits 16-23 scatter factor
        const int32_t scatter_factor = ((pub.mode >> 16) & 0xff) ? ((pub.mode >> 16) & 0xff) : 4;
        const int32_t push_min_rows = ((pub.mode >> 8) & 0xff) ? 16 * ((pub.mode >> 8) & 0xff) : PUSH_MIN_ROWS;
        cx.scatter = reads_elsewhere && !pub.mc && (chunk_last_row - cp->first_row + 1 > scatter_factor * (n_rec + CVR_W));
        const bool publishing = reads_elsewhere && !cx.scatter; // range pushes behind the watermark
        int32_t pushed_upto = cp->first_row; // kPublish: first row of the chunk not yet sent to the peers
        cx.split1 = cp->split1;
        cx.first_row = cp->first_row;
        cx.l = l;
        cx.last_done = cp->first_row - 1;

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
        // batch k = records [32k, 32k + 32): requested RD - 1 batches before it is taken (ncu on road: 46 % of the
        // sweep's stall samples sat on the first use of a batch requested ONE batch earlier, 3.8 batches per tile).
        // Exactly one cp.async group is committed per batch index, empty or not, so "batch k has landed" is always
        // "at most RD - 1 groups pending".
        const uint32_t my_slot = smem_u32(&s_rec[w][0][RD > 0 ? t : 0]);
        auto fetch_batch = [&](int32_t k) {
            const int32_t i = 32 * k + t;
            if (i < n_rec)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(my_slot + (uint32_t)(k % (RD > 0 ? RD : 1)) * 256u),
                             "l"(rec + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto take_batch = [&](int32_t k) -> int2 {
            asm volatile("cp.async.wait_group %0;" ::"n"(RD > 0 ? RD - 1 : 0) : "memory");
            return (32 * k + t < n_rec) ? s_rec[w][k % (RD > 0 ? RD : 1)][RD > 0 ? t : 0] : make_int2(-1, 0);
        };
        int2 held, nxt = make_int2(-1, 0);
        if constexpr (RD > 0) {
#pragma unroll
            for (int k = 0; k < RD; k++) fetch_batch(k);
            held = take_batch(0);
        } else {
            // the warp holds 32 records (one coalesced 256 B load) and requests the next 32 when it takes them
            held = (t < n_rec) ? rec[t] : make_int2(-1, 0);
            nxt = (32 + t < n_rec) ? rec[32 + t] : make_int2(-1, 0);
        }
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
                // records are sorted by position: the batch reaches past the tile (or the list ended, pos = -1)
                // as soon as ANY lane holds a position beyond it -- a vote, not a shuffle through the LSU queue
                if (__any_sync(FULL, (uint32_t)held.x >= (uint32_t)(ts + TILE))) break;
                rb += 32;
                if constexpr (RD > 0) {
                    fetch_batch(rb / 32 + RD - 1); // into the slot of the batch just used up (my own element of it)
                    held = take_batch(rb / 32);
                } else {
                    held = nxt;
                    nxt = (rb + 32 + t < n_rec) ? rec[rb + 32 + t] : make_int2(-1, 0);
                }
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
            if (kPublish && publishing && ((tile & 3) == 3 || tile + 1 == n_tiles)) {
                // watermark: every row up to min over the SIMD lanes of the lane's latest stored row is final
                // (taken every fourth tile: five shuffles through the unit the sweep is bound by)
                int32_t f = cx.last_done;
                f = max(f, __shfl_xor_sync(FULL, f, 8));
                f = max(f, __shfl_xor_sync(FULL, f, 16)); // latest row of SIMD lane l (any of its four walkers)
                f = min(f, __shfl_xor_sync(FULL, f, 1));
                f = min(f, __shfl_xor_sync(FULL, f, 2));
                f = min(f, __shfl_xor_sync(FULL, f, 4));
                if (f - pushed_upto + 1 >= push_min_rows) {
                    publish_rows(cx, pushed_upto, f, t);
                    pushed_upto = f + 1;
                }
            }
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
        if (kPublish && publishing && pushed_upto <= chunk_last_row)
            publish_rows(cx, pushed_upto, chunk_last_row, t); // what the watermark had not reached
        if constexpr (RD > 0) asm volatile("cp.async.wait_group 0;" ::: "memory"); // the next chunk counts from zero
        chunk = queue ? __shfl_sync(FULL, next_chunk, 0) : next_chunk;
    }
    // the last warp to leave resets the queue for the next launch (every warp has drawn its last ticket by then)
    if (queue && t == 0) {
        if (atomicAdd(queue + 1, 1u) == (unsigned)n_warps - 1u) {
            queue[0] = 0u;
            queue[1] = 0u;
        }
    }
}



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

// When a large share of the rows has to be cleared (R-MAT-24: 6.7 M empty rows of 16.7 M) one streaming pass
// over all of y beats millions of scattered 8-byte stores (60 us -> 25 us per sweep): 16-byte stores,
// grid-stride; same programmatic-launch role as cvr_clear_rows_kernel.
__global__ void cvr_zero_y_kernel(double* __restrict__ y, int64_t n)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    // y comes from cudaMalloc or a tensor: 16-byte aligned in practice, but do not rely on it
    const int64_t head = (reinterpret_cast<uintptr_t>(y) & 8) ? 1 : 0;
    if (tid == 0 && head) y[0] = 0.0;
    double2* y2 = reinterpret_cast<double2*>(y + head);
    const int64_t n2 = (n - head) / 2;
    for (int64_t i = tid; i < n2; i += stride) y2[i] = make_double2(0.0, 0.0);
    if (tid == 0 && ((n - head) & 1)) y[n - 1] = 0.0;
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
// The next iteration's sweep is launched as a programmatic dependent of this kernel.  (Launching THIS
// kernel as a programmatic dependent of the sweep as well was measured on 2 x B200, R-MAT-24: 991 us per
// iteration against 893 us with an ordinary launch -- its early-resident blocks get in the sweep's way --
// profiles/r02_bench_n2_pdl.txt; so it is an ordinary launch.)
__global__ void cvr_publish_epilogue_kernel(double* __restrict__ y, const int32_t* __restrict__ boundary,
                                            int32_t n_boundary, const int32_t* __restrict__ empty,
                                            int32_t n_empty, const __grid_constant__ CvrPublish pub,
                                            const __grid_constant__ CvrBarrier bar, unsigned int* done_counter)
{
    // the next iteration's sweep may start its prologue (barrier set-up, first bulk copies of the
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
        // never-written rows: their `needs` byte is cleared (the sweep's range pushes skip them), so the
        // explicit 0.0 of the first two iterations goes to every destination but the own aliased buffer
        if (pub.mc) {
            multicast_store(pub.mc + g, v);
            continue;
        }
        const uint32_t nb = b ? publish_mask(pub, row) : ((pub.mode & 4) ? (0xffu & ~(1u << pub.self)) : 0xffu);
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

// ---- sweep geometries: cvr_spmv_tile_kernel<TB, NB>, TB steps per walker (tile = 32 TB elements), NB
// resident blocks (4 warps each) per SM requested through __launch_bounds__.  Measured on B200 (kernel us):
//   profiles/r02_kernel_ab_tile_geometries.jsonl   FEM 70.3 / 65.3 / 62.9, R-MAT-24 1358 / 1387 / 1369, web 38.3 / 40.3 /
//   41.3, road 342.8 / 354.6 / 370.9 for 7x7 / 9x6 / 11x5 -- short rows want more warps, long regular rows the
//   larger tile (fewer per-tile flag/carry operations per element);
//   profiles/r02_kernel_ab_l1_capacity.txt   7x6 against 7x7: R-MAT-24 1316 / 1354, web 38.9 / 40.9, road 357 / 359
//   -- 78 registers per thread instead of 72 with spills.  The same file records what the x gather needs from L1:
//   room for its misses in flight.  The driver configures 132 KB of shared memory (L1 124 KB) for these kernels;
//   124 KB are enough (156 KB: no change), but 92 / 60 / 28 KB of L1 cost R-MAT-24 15 / 50 / 180 %, which is why
//   nothing that needs more shared memory per warp (record ring, deeper TMA ring, x cache) ever paid off.
struct Variant {
    const char* name;
    int tb, nb, rd;
};
// The record ring costs shared memory (RD x 1 KB per block) and the x gather wants L1 (above): five blocks with a
// three-deep ring fit the 100 KB shared-memory configuration (L1 156 KB), six blocks with any ring need the 132 KB one.
// Same box, kernel us (profiles/r02_kernel_ab_record_ring.txt):
//                          R-MAT-24    web    road    FEM      records: batches of 32 per 224-element tile
//   tile7x5r (RD 3, L1 156)  1205.9    36.2   321.7   65.3     R-MAT-24 0.2, web 1.2, road 2.9, FEM 0.3
//   tile7x6r (RD 4, L1 124)  1278.6    36.7   319.4   66.8
//   tile7x6  (RD 0, L1 156)  1220.0    38.4   356.5   69.3
//   tile11x5 (RD 0, L1 124)  1332.7    40.7   370.1   63.2
// (tile7x6 with the driver's own 132 KB configuration, the round-2 kernel until then: 1312 / 38.6 / 357 / -.)
constexpr Variant VARIANTS[] = {
    {"tile7x5r", 7, 5, 3},
    {"tile11x5", 11, 5, 0},
    {"tile7x6", 7, 6, 0},
    {"tile7x6r", 7, 6, 4},
};
constexpr int N_VARIANTS = sizeof(VARIANTS) / sizeof(VARIANTS[0]);

int forced_variant()
{
    // CVR_SPMV_KERNEL = tile7x5r | tile11x5 | tile7x6 | tile7x6r overrides the per-matrix choice; read per call so
    // that tools/kernel_ab.py and the parity suite can switch it at run time
    const char* e = getenv("CVR_SPMV_KERNEL");
    if (!e || !*e) return -1;
    for (int v = 0; v < N_VARIANTS; v++)
        if (strcmp(e, VARIANTS[v].name) == 0) return v;
    return -1;
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

// one entry per geometry: occupancy query and launch, plain and publishing flavour
template <int TB, int NB, int RD>
struct TileOps {
    static int resident_blocks()
    {
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, cvr_spmv_tile_kernel<TB, NB, RD, false>, WARPS * 32,
                                                          Geo<TB>::DYN_SMEM) != cudaSuccess)
            return 0;
        return blocks;
    }
    static cudaError_t launch(bool publish, int blocks, cudaStream_t stream, bool programmatic,
                              const CvrChunk* chunks, int32_t chunk_begin, int32_t chunk_end, const double* vals,
                              const int32_t* cols, const int32_t* record, const double* x, double* y,
                              const CvrPublish& pub, unsigned int* queue)
    {
        if (publish)
            return launch_ex(cvr_spmv_tile_kernel<TB, NB, RD, true>, blocks, WARPS * 32, Geo<TB>::DYN_SMEM, stream,
                             programmatic, chunks, chunk_begin, chunk_end, vals, cols, record, x, y, pub, queue);
        return launch_ex(cvr_spmv_tile_kernel<TB, NB, RD, false>, blocks, WARPS * 32, Geo<TB>::DYN_SMEM, stream,
                         programmatic, chunks, chunk_begin, chunk_end, vals, cols, record, x, y, pub, queue);
    }
    // shared-memory carve-out as a percentage of the SM's 228 KB (-1: the driver's choice)
    static void set_carveout(int pct)
    {
        cudaFuncSetAttribute(cvr_spmv_tile_kernel<TB, NB, RD, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(cvr_spmv_tile_kernel<TB, NB, RD, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    static void preload()
    {
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<TB, NB, RD, true>);
        cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<TB, NB, RD, false>);
        // ask for the shared-memory configuration the NB resident blocks need and no more: the driver's own choice is
        // sized for the blocks shared memory ALONE would admit, and every step up (132 -> 164 -> 196 KB) takes L1 away
        // from the x gather (profiles/r02_kernel_ab_l1_capacity.txt)
        set_carveout(carveout_pct());
    }
    static int carveout_pct()
    {
        cudaFuncAttributes a;
        if (cudaFuncGetAttributes(&a, cvr_spmv_tile_kernel<TB, NB, RD, false>) != cudaSuccess) return -1;
        const size_t need = (size_t)NB * (a.sharedSizeBytes + Geo<TB>::DYN_SMEM + 1024);
        const int pct = (int)((need * 100 + 233471) / 233472);
        return pct > 100 ? 100 : pct;
    }
};

#define CVR_FOR_VARIANT(v, expr)                            \
    switch (v) {                                            \
    case 0: { using P = TileOps<7, 5, 3>; expr; } break;    \
    case 1: { using P = TileOps<11, 5, 0>; expr; } break;   \
    case 2: { using P = TileOps<7, 6, 0>; expr; } break;    \
    default: { using P = TileOps<7, 6, 4>; expr; } break;   \
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
    CVR_FOR_VARIANT(v, blocks = P::resident_blocks())
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

// Geometry for a matrix with `nnz` stored elements in `n_rows` rows: the large tile pays off when rows are long and
// regular (few row switches per tile: FEM); everything else takes the small tile with the record ring.
// CVR_SPMV_KERNEL overrides.
int cvr_pick_sweep_variant(int64_t nnz, int64_t n_rows)
{
    const int f = forced_variant();
    if (f >= 0) return f;
    return (n_rows > 0 && nnz / n_rows >= 20) ? 1 : 0;
}

// resident warps per SM of a variant (sizes the automatic chunk count)
int cvr_spmv_resident_warps_per_sm(int variant)
{
    return variant_resident_blocks(variant) * WARPS;
}

const char* cvr_spmv_kernel_name(int variant)
{
    const int f = forced_variant();
    return VARIANTS[f >= 0 ? f : variant].name;
}

int cvr_launch_spmv(int variant, const CvrChunk* chunks, int32_t n_chunks, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, const CvrRowLists& rows, const CvrPublish* publish,
                    cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end,
                    const CvrBarrier* barrier, unsigned int* done_counter, bool y_is_clear,
                    int32_t chunk_begin, int32_t chunk_end, unsigned int* chunk_queue)
{
    if (chunk_end < 0) chunk_end = n_chunks;
    int launched = 0;
    const int f = forced_variant();
    const int v = f >= 0 ? f : variant;
    const int sms = device_sm_count();
    const int32_t n_clear = rows.n_boundary + rows.n_empty;
    bool after_clear_kernel = false;
    if (y_is_clear) {
        // the previous iteration's epilogue kernel already cleared the accumulated rows
    } else if (rows.boundary && !publish && (int64_t)n_clear * 4 > n_rows) {
        // more than a quarter of the rows would be cleared one by one: zero all of y instead
        cvr_zero_y_kernel<<<sms * 8, 256, 0, stream>>>(y, n_rows + 1);
        launched++;
        after_clear_kernel = true;
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
    const int blocks = (int)(((int64_t)(chunk_end - chunk_begin) * 32 + threads - 1) / threads);
    // the sweeps are persistent: one block per resident slot, warps stride over the chunks
    const int resident = sms * variant_resident_blocks(v);
    const int pblocks = blocks < resident ? blocks : resident;
    CvrPublish none{};
    const bool pub = publish && publish->n_dst > 0;
    // the kernel in front of us on the stream is the clearing kernel or (iterated SpMV, from the second
    // iteration on) the previous iteration's epilogue: launch as its programmatic dependent
    const bool pdl = pdl_enabled() && !(pub && (publish->mode & 8)); // mode bit 3: two shards share this device
    const bool programmatic = pdl && (after_clear_kernel || (pub && y_is_clear)) && !ev_begin;
    if (const char* co = getenv("CVR_SMEM_CARVEOUT")) { // experiment: L1 size against shared-memory carve-out
        const int pct = atoi(co);
        CVR_FOR_VARIANT(v, P::set_carveout(pct))
    }
    // the chunk queue pays off when a warp walks many chunks (R-MAT-24, 16 per warp: 1399 -> 1354 us); with one
    // or two chunks per warp the ticket only costs (web 38.9 -> 40.9 us)
    if ((int64_t)(chunk_end - chunk_begin) < 4 * (int64_t)resident * WARPS) chunk_queue = nullptr;
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    cudaError_t e = cudaSuccess;
    CVR_FOR_VARIANT(v, e = P::launch(pub, pblocks, stream, programmatic, chunks, chunk_begin, chunk_end, vals, cols,
                                     record, x, y, pub ? *publish : none, chunk_queue))
    if (e != cudaSuccess) return -1;
    launched++;
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (pub) {
        if (!barrier || !done_counter) return -1;
        const int cb = (n_clear + 255) / 256;
        const int cap = sms * 4;
        e = launch_ex(cvr_publish_epilogue_kernel, cb < cap ? (cb < 1 ? 1 : cb) : cap, 256, 0, stream,
                      false, y, (const int32_t*)rows.boundary, rows.n_boundary,
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
__global__ void cvr_clear_needs_kernel(const int32_t* __restrict__ empty, int32_t n_empty, uint8_t* __restrict__ needs)
{
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_empty; i += gridDim.x * blockDim.x)
        needs[empty[i]] = 0;
}

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

int cvr_launch_chunk_needs(const CvrChunk* chunks, int32_t n_chunks, const CvrRowLists& rows, uint8_t* needs,
                           uint8_t* chunk_any, cudaStream_t stream)
{
    // rows nothing ever writes are published once by the epilogue (an explicit 0.0 during the first two
    // iterations), never by the sweep: drop them from the footprint the range pushes consult
    if (rows.n_empty > 0) {
        const int blocks = (rows.n_empty + 255) / 256;
        cvr_clear_needs_kernel<<<blocks < 4096 ? blocks : 4096, 256, 0, stream>>>(rows.empty, rows.n_empty, needs);
    }
    cvr_chunk_needs_kernel<<<(n_chunks * 32 + 127) / 128, 128, 0, stream>>>(chunks, n_chunks, needs, chunk_any);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
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

void cvr_preload_spmv_kernels()
{
    cudaFuncAttributes a;
    for (int v = 0; v < N_VARIANTS; v++) {
        CVR_FOR_VARIANT(v, P::preload())
    }
    cudaFuncGetAttributes(&a, cvr_clear_rows_kernel);
    cudaFuncGetAttributes(&a, cvr_zero_y_kernel);
    // The multi-GPU kernels wait for each other on the device (flag barrier).  CUDA loads a kernel lazily at
    // its first launch and that load can synchronise with running work: a first launch issued while a peer's
    // epilogue is already spinning at the barrier would never get through.  Load everything up front.
    cudaFuncGetAttributes(&a, cvr_publish_epilogue_kernel);
    cudaFuncGetAttributes(&a, cvr_peer_barrier_kernel);
    cudaFuncGetAttributes(&a, cvr_column_footprint_kernel);
    cudaFuncGetAttributes(&a, cvr_chunk_needs_kernel);
    cudaFuncGetAttributes(&a, cvr_clear_needs_kernel);
}

