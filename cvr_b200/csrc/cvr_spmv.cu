// y = A * x over the CVR arrays on the device (sm_100a), fp64.
//
// Replaces the five-phase streaming loop of spmv_compute_kernel
// (/root/reference/spmv.cpp:1016-1667) and implements its INTENDED semantics
// (paper Alg. 4), including the steal-only chunks (split1 == -1) the
// reference kernel mishandles.
//
// Mapping.  The reference gives one chunk to one OpenMP thread whose AVX-512 register
// holds the 8 SIMD lanes of a step.  Here one WARP owns one chunk and covers 4 steps x 8
// lanes = 32 consecutive CVR elements per pass ("window"): thread t of the warp holds
// element 32k + t, i.e. step 4k + (t >> 3), SIMD lane (t & 7).  vals (8 B) and cols (4 B)
// are therefore read with fully coalesced 256 B / 128 B warp loads, x is gathered through
// L1/L2 (ld.global.nc), and each thread keeps a private partial sum for its SIMD lane.
//
// Row switches.  A record (pos, wb) means "the accumulator of lane pos%8 is flushed
// before step pos/8" (spmv.cpp:1197-1210).  The warp holds the next 32 records in
// registers (one per thread), turns the ones that fall into the current window into a
// 32-bit flag word with one REDUX.OR, and only then runs a segmented reduction along each
// SIMD lane (stride-8 shuffles).  Windows without a flag take the fast path: one FMA.
//   feeding record (pos <= split1)  -> y[wb] = sum          plain store, row owned by chunk
//   stealing record                 -> carry[lane] += sum   (wb == own lane on a first steal)
//   split0 position                 -> y[first_row] += sum  atomic, row shared with the
//                                                           previous chunk (spmv.cpp:1280)
// After the last step the eight pos=-1 records route each lane's remainder into a carry
// slot (spmv.cpp:1633-1638) and the eight carries are added atomically to y[tail[.]]
// (spmv.cpp:1640-1649).  y must be zero on entry (the launcher clears it on the same
// stream, inside the timed region).
#include "cvr_internal.h"

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
// lanes = 256 elements and thread t = (q, l) walks EIGHT CONSECUTIVE steps of SIMD lane l:
// steps 8q .. 8q+7 of the tile.  A row switch is then a thread-local event (emit the
// accumulator, clear it); threads of the same SIMD lane only meet once per tile, in a
// 3-shuffle carry chain that hands the open partial sum from walker q to walker q+1.
//   * loads: for each of the 8 steps a warp load touches 4 x 64 B (vals) / 4 x 32 B (cols)
//     fully used sectors; all 16 loads of a tile are issued before the first use.
//   * records are delivered to their owner thread through shared memory: the warp holds 32
//     records in registers (coalesced 256 B load), each holder drops a flag byte and the
//     write-back target into the owner's slot.
// ---------------------------------------------------------------------------------------
constexpr int TB = 8;                 // consecutive steps per walker
constexpr int TILE = 4 * TB * CVR_W;  // 256 elements per warp pass
constexpr int WARPS = 4;              // warps per block
constexpr int32_t WB_SPLIT0 = -2;     // marker: flush into the shared first row

struct TileCtx {
    double* __restrict__ y;
    const int32_t* tail;
    int32_t split1, first_row;
    int l;
};

__device__ __forceinline__ void emit(const TileCtx& cx, double value, int32_t pos, int32_t wb,
                                     double& carry_slot)
{
    if (wb == WB_SPLIT0) atomicAdd(&cx.y[cx.first_row], value);            // spmv.cpp:1280-1282
    else if (cx.split1 != -1 && pos <= cx.split1) cx.y[wb] = value;        // feeding, :1204
    else if (wb == cx.l) carry_slot += value;                              // stealing, :1541
    else if (cx.tail[wb] != 0) atomicAdd(&cx.y[cx.tail[wb]], value);       // (unreachable)
}

__global__ void __launch_bounds__(WARPS * 32)
cvr_spmv_kernel(const CvrChunk* __restrict__ chunks, int32_t T,
                const double* __restrict__ vals, const int32_t* __restrict__ cols,
                const int32_t* __restrict__ record, const double* __restrict__ x,
                double* __restrict__ y)
{
    __shared__ unsigned long long s_flags[WARPS][32]; // 8 flag bytes per thread (one per step)
    __shared__ int32_t s_wb[WARPS][TB][32];           // write-back target per (step, thread)

    const int32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chunk >= T) return;
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = t >> 3, l = t & 7;

    const CvrChunk* cp = chunks + chunk;
    const int64_t start = cp->start;
    const int32_t len = cp->len;
    const int32_t split0 = cp->split0;
    const int32_t n_rec = cp->n_rec;
    TileCtx cx;
    cx.y = y;
    cx.tail = cp->tail;
    cx.split1 = cp->split1;
    cx.first_row = cp->first_row;
    cx.l = l;

    const int2* rec = reinterpret_cast<const int2*>(record + cvr_record_offset(chunk, cx.first_row));
    const double* v = vals + start;
    const int32_t* c = cols + start;

    s_flags[w][t] = 0ull;
    int32_t rb = 0;
    int2 held = (t < n_rec) ? rec[t] : make_int2(-1, 0);
    double lane_carry = 0.0; // open partial sum of SIMD lane l at the tile boundary
    double carry_slot = 0.0; // private share of t_rets[l] (spmv.cpp:1124)
    __syncwarp();

    const int32_t n_tiles = (len + TILE - 1) / TILE;
    for (int32_t tile = 0; tile < n_tiles; tile++) {
        const int32_t ts = tile * TILE;
        const int32_t p0 = ts + q * (TB * CVR_W) + l; // my element of step 8q; next step: +8

        double a[TB], xv[TB];
        int32_t ci[TB];
#pragma unroll
        for (int b = 0; b < TB; b++) {
            const int32_t p = p0 + b * CVR_W;
            const bool in = p < len;
            a[b] = in ? ld_stream_f64(v + p) : 0.0;
            ci[b] = in ? ld_stream_s32(c + p) : 0;
        }
#pragma unroll
        for (int b = 0; b < TB; b++) xv[b] = __ldg(x + ci[b]);

        // ---- deliver the records of this tile to their owner threads
        for (;;) {
            const uint32_t rel = (uint32_t)(held.x - ts);
            if (rel < (uint32_t)TILE) {
                const uint32_t step = rel >> 3;
                const uint32_t owner = (step >> 3) * CVR_W + (rel & 7u);
                reinterpret_cast<unsigned char*>(&s_flags[w][owner])[step & 7u] = 1;
                s_wb[w][step & 7u][owner] = held.y;
            }
            const int32_t last = __shfl_sync(FULL, held.x, 31);
            if ((uint32_t)last >= (uint32_t)(ts + TILE)) break; // batch reaches past the tile (or ended)
            rb += 32;
            held = (rb + t < n_rec) ? rec[rb + t] : make_int2(-1, 0);
        }
        if (t == 0 && split0 != 0) {
            const uint32_t rel = (uint32_t)(split0 - ts);
            if (rel < (uint32_t)TILE) {
                const uint32_t step = rel >> 3;
                const uint32_t owner = (step >> 3) * CVR_W + (rel & 7u);
                reinterpret_cast<unsigned char*>(&s_flags[w][owner])[step & 7u] = 1;
                s_wb[w][step & 7u][owner] = WB_SPLIT0;
            }
        }
        __syncwarp();
        const unsigned long long f = s_flags[w][t];

        // ---- walk my eight steps; the first flush waits for the carry of earlier walkers
        double acc = 0.0, head = 0.0;
        int first_b = -1;
        if (f == 0ull) {
#pragma unroll
            for (int b = 0; b < TB; b++) acc = fma(a[b], xv[b], acc);
        } else {
            s_flags[w][t] = 0ull;
#pragma unroll
            for (int b = 0; b < TB; b++) {
                if ((f >> (8 * b)) & 1ull) {
                    if (first_b < 0) {
                        head = acc;
                        first_b = b;
                    } else {
                        emit(cx, acc, p0 + b * CVR_W, s_wb[w][b][t], carry_slot);
                    }
                    acc = 0.0;
                }
                acc = fma(a[b], xv[b], acc);
            }
        }
        const bool has = first_b >= 0;
        double cin = (q == 0) ? lane_carry : 0.0;
        double out = has ? acc : cin + acc;
#pragma unroll
        for (int r = 1; r < 4; r++) {
            const double prev = __shfl_up_sync(FULL, out, CVR_W);
            if (q == r) {
                cin = prev;
                out = has ? acc : cin + acc;
            }
        }
        lane_carry = __shfl_sync(FULL, out, 24 + l);
        if (has) emit(cx, head + cin, p0 + first_b * CVR_W, s_wb[w][first_b][t], carry_slot);
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
}

} // namespace

int cvr_launch_spmv(const CvrChunk* chunks, int32_t n_chunks, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end)
{
    if (cudaMemsetAsync(y, 0, sizeof(double) * (size_t)(n_rows + 1), stream) != cudaSuccess)
        return -1;
    const int threads = 128;
    const int blocks = (int)(((int64_t)n_chunks * 32 + threads - 1) / threads);
    // CVR_SPMV_KERNEL=window selects the first-generation kernel (kept for A/B profiling)
    static const bool use_window = [] {
        const char* e = getenv("CVR_SPMV_KERNEL");
        return e && strcmp(e, "window") == 0;
    }();
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    if (use_window)
        cvr_spmv_window_kernel<<<blocks, threads, 0, stream>>>(chunks, n_chunks, vals, cols, record, x, y);
    else
        cvr_spmv_kernel<<<blocks, threads, 0, stream>>>(chunks, n_chunks, vals, cols, record, x, y);
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
}
