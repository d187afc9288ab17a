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
cvr_spmv_kernel(const CvrChunk* __restrict__ chunks, int32_t T,
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

} // namespace

int cvr_launch_spmv(const CvrChunk* chunks, int32_t n_chunks, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end)
{
    if (cudaMemsetAsync(y, 0, sizeof(double) * (size_t)(n_rows + 1), stream) != cudaSuccess)
        return -1;
    const int threads = 128;
    const int blocks = (int)(((int64_t)n_chunks * 32 + threads - 1) / threads);
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    cvr_spmv_kernel<<<blocks, threads, 0, stream>>>(chunks, n_chunks, vals, cols, record, x, y);
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
}
