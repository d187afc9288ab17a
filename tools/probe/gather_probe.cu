// Measurement aid (not part of the product): how fast can one B200 stream a CVR column array and
// gather x through L1/L2, with the SpMV bookkeeping (records, row switches, write-back) removed?
// The SpMV sweep (cvr_b200/csrc/cvr_spmv.cu) cannot beat these kernels on the same cols array:
// they are its ceiling on matrices whose x gather misses L1 (R-MAT, web), and tell which knob
// (warps per SM, gathers in flight per thread, L1 allocation policy) the ceiling responds to.
//
//   probe_run(variant, U, cols, vals, n, x, out, blocks_per_sm, col_mask, reps, &ms)
//     variant 0: cols only                       (4 B/element stream)
//     variant 1: cols + gather x                 (the gather alone)
//     variant 2: cols + vals + gather + FMA      (12 B/element stream + gather: SpMV minus records)
//     variant 3: as 2, gather with L1::no_allocate
//     variant 4: as 2, stream through cp.async.bulk (TMA) into shared memory, as the sweep does
//     variant 5: as 2, gathers of tile k+1 issued before the FMAs of tile k (software pipeline)
//     variant 6: as 2, every warp walks a contiguous chunk (probe_set_chunk) -- the sweep's access pattern
//     variant 7 (probe_hot_run): as 2 plus a shared-memory cache of the hottest x entries
//   U = elements (gathers in flight) per thread and pass: 4, 9 or 18.
//   col_mask != 0: column &= col_mask  (shrinks the x footprint, e.g. to an L2-resident window)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

int g_chunk_el = 4032; // variant 6: elements per contiguous chunk (multiple of every tile size)

__device__ __forceinline__ int32_t ld_stream_s32(const int32_t* p)
{
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_gather_na(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// volatile on purpose: ptxas otherwise interleaves the FMAs with the loads to save registers, and an
// FMA waiting for its gather then blocks the issue of the gathers behind it (U stops being the
// number of gathers in flight)
__device__ __forceinline__ double ld_gather(const double* p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fma_ordered(double& acc, double a, double b)
{
    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b));
}
__device__ __forceinline__ void add_ordered(double& acc, double a)
{
    asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc) : "d"(a));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT, int U>
__global__ void __launch_bounds__(128)
probe_kernel(const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
             const double* __restrict__ x, double* __restrict__ out, uint32_t col_mask)
{
    constexpr int TILE = 32 * U;
    const int t = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = n / TILE;
    double acc = 0.0;
    for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * TILE + t;
        uint32_t c[U];
        double a[U], g[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            c[u] = (uint32_t)ld_stream_s32(cols + base + 32 * u);
            if (VARIANT >= 2) a[u] = ld_stream_f64(vals + base + 32 * u);
        }
        if (VARIANT == 0) {
#pragma unroll
            for (int u = 0; u < U; u++) acc += (double)c[u];
            continue;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t ci = col_mask ? (c[u] & col_mask) : c[u];
            g[u] = VARIANT == 3 ? ld_gather_na(x + ci) : ld_gather(x + ci);
        }
        // basic-block boundary: ptxas schedules inside basic blocks and otherwise interleaves the FMAs with
        // the loads (an FMA waiting for its gather then blocks the gathers behind it)
        if (col_mask == 0xdeadbeefu) out[1 + t] = (double)c[0];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (VARIANT >= 2) fma_ordered(acc, a[u], g[u]);
            else add_ordered(acc, g[u]);
        }
    }
    if (acc == 123.456) out[0] = acc; // keep the loads alive
}

// variant 4: the stream goes through the TMA engine (one elected lane, cp.async.bulk + mbarrier),
// one stage per warp, registers are the second buffer -- the staging scheme of the sweep kernel
template <int U>
__global__ void __launch_bounds__(128)
probe_tma_kernel(const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
                 const double* __restrict__ x, double* __restrict__ out, uint32_t col_mask)
{
    constexpr int TILE = 32 * U;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[4];
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned char* stage = smem + w * TILE * 12;
    const uint32_t bar = smem_u32(&s_bar[w]);
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = n / TILE;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int64_t tile) {
        if (t == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE * 12) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(stage)), "l"(vals + tile * TILE), "r"(TILE * 8), "r"(bar), "l"(policy) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(stage + TILE * 8)), "l"(cols + tile * TILE), "r"(TILE * 4), "r"(bar), "l"(policy) : "memory");
        }
    };
    double acc = 0.0;
    uint32_t phase = 0;
    if (warp0 < n_tiles) issue(warp0);
    for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
            ::"r"(bar), "r"(phase) : "memory");
        phase ^= 1u;
        uint32_t c[U];
        double a[U], g[U];
        const double* sv = reinterpret_cast<const double*>(stage) + t;
        const uint32_t* sc = reinterpret_cast<const uint32_t*>(stage + TILE * 8) + t;
#pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = sv[32 * u];
            c[u] = sc[32 * u];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (tile + n_warps < n_tiles) issue(tile + n_warps);
#pragma unroll
        for (int u = 0; u < U; u++) g[u] = ld_gather(x + (col_mask ? (c[u] & col_mask) : c[u]));
        if (col_mask == 0xdeadbeefu) out[1 + t] = (double)c[0];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; u++) fma_ordered(acc, a[u], g[u]);
    }
    if (acc == 123.456) out[0] = acc;
}


// variant 5: as 2, software-pipelined inside the warp -- the cols/vals loads and the gathers of tile
// k+1 are issued before the FMAs of tile k, so a warp always has U..2U gathers in flight
template <int U>
__global__ void __launch_bounds__(128)
probe_pipe_kernel(const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
                  const double* __restrict__ x, double* __restrict__ out, uint32_t col_mask)
{
    constexpr int TILE = 32 * U;
    const int t = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = n / TILE;
    double acc = 0.0;
    double a[U], g[U];
    if (warp0 < n_tiles) {
        const int64_t base = warp0 * TILE + t;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t c = (uint32_t)ld_stream_s32(cols + base + 32 * u);
            a[u] = ld_stream_f64(vals + base + 32 * u);
            g[u] = ld_gather(x + (col_mask ? (c & col_mask) : c));
        }
    }
    for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        double a2[U], g2[U];
        const bool more = tile + n_warps < n_tiles;
        if (more) {
            const int64_t base = (tile + n_warps) * TILE + t;
            uint32_t c[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                c[u] = (uint32_t)ld_stream_s32(cols + base + 32 * u);
                a2[u] = ld_stream_f64(vals + base + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < U; u++) g2[u] = ld_gather(x + (col_mask ? (c[u] & col_mask) : c[u]));
        }
#pragma unroll
        for (int u = 0; u < U; u++) fma_ordered(acc, a[u], g[u]);
        if (more) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                a[u] = a2[u];
                g[u] = g2[u];
            }
        }
    }
    if (acc == 123.456) out[0] = acc;
}

// variant 6: as 2, but every warp walks a CONTIGUOUS chunk of `chunk_el` elements tile by tile and
// strides over chunks (chunk = warp, warp + n_warps, ...) -- the sweep's access pattern: n_warps
// concurrent streams that are chunk_el * 12 bytes apart instead of one compact front
template <int U>
__global__ void __launch_bounds__(128)
probe_chunked_kernel(const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
                     const double* __restrict__ x, double* __restrict__ out, uint32_t col_mask, int chunk_el)
{
    constexpr int TILE = 32 * U;
    const int t = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_chunks = n / chunk_el;
    const int tiles_per_chunk = chunk_el / TILE;
    double acc = 0.0;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += n_warps) {
        for (int k = 0; k < tiles_per_chunk; k++) {
            const int64_t base = chunk * chunk_el + (int64_t)k * TILE + t;
            uint32_t c[U];
            double a[U], g[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                c[u] = (uint32_t)ld_stream_s32(cols + base + 32 * u);
                a[u] = ld_stream_f64(vals + base + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < U; u++) g[u] = ld_gather(x + (col_mask ? (c[u] & col_mask) : c[u]));
#pragma unroll
            for (int u = 0; u < U; u++) fma_ordered(acc, a[u], g[u]);
        }
    }
    if (acc == 123.456) out[0] = acc;
}

template <int VARIANT, int U>
cudaError_t launch(const int32_t* cols, const double* vals, int64_t n, const double* x, double* out,
                   int grid, uint32_t col_mask, int carveout)
{
    if (VARIANT == 4) {
        auto k = probe_tma_kernel<U>;
        const int smem = 4 * 32 * U * 12;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (carveout >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        k<<<grid, 128, smem>>>(cols, vals, n, x, out, col_mask);
    } else if (VARIANT == 6) {
        auto k = probe_chunked_kernel<U>;
        if (carveout >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        k<<<grid, 128>>>(cols, vals, n, x, out, col_mask, g_chunk_el);
    } else if (VARIANT == 5) {
        auto k = probe_pipe_kernel<U>;
        if (carveout >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        k<<<grid, 128>>>(cols, vals, n, x, out, col_mask);
    } else {
        auto k = probe_kernel<VARIANT, U>;
        if (carveout >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        k<<<grid, 128>>>(cols, vals, n, x, out, col_mask);
    }
    return cudaGetLastError();
}

template <int U>
cudaError_t dispatch(int variant, const int32_t* cols, const double* vals, int64_t n, const double* x,
                     double* out, int grid, uint32_t col_mask, int carveout)
{
    switch (variant) {
    case 0: return launch<0, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 1: return launch<1, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 2: return launch<2, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 3: return launch<3, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 4: return launch<4, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 5: return launch<5, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    case 6: return launch<6, U>(cols, vals, n, x, out, grid, col_mask, carveout);
    }
    return cudaErrorInvalidValue;
}

} // namespace


// variant 7: as 2, plus a shared-memory cache of the K hottest x entries.  `cols` must be relabelled by
// popularity (rank 0 = most frequent column; run_probe.py does it), so "hot" is simply c < K; in the
// product the same effect needs a remapped copy of the column stream.  One block per SM slot, `threads`
// threads each (the cache is per block): measures what the miss path gains when a fraction of the
// gathers never reaches it.
template <int U>
__global__ void __launch_bounds__(1024)
probe_hot_kernel(const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
                 const double* __restrict__ x, double* __restrict__ out, uint32_t hot_k)
{
    constexpr int TILE = 32 * U;
    extern __shared__ __align__(16) double s_hot[];
    for (uint32_t i = threadIdx.x; i < hot_k; i += blockDim.x) s_hot[i] = x[i];
    __syncthreads();
    const int t = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = n / TILE;
    double acc = 0.0;
    for (int64_t tile = warp0; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * TILE + t;
        uint32_t c[U];
        double a[U], g[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            c[u] = (uint32_t)ld_stream_s32(cols + base + 32 * u);
            a[u] = ld_stream_f64(vals + base + 32 * u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (c[u] < hot_k) g[u] = s_hot[c[u]];
            else g[u] = ld_gather(x + c[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) fma_ordered(acc, a[u], g[u]);
    }
    if (acc == 123.456) out[0] = acc;
}

extern "C" int probe_hot_run(int U, const int32_t* cols, const double* vals, int64_t n, const double* x,
                             double* out, int blocks_per_sm, int threads, uint32_t hot_k, int reps, float* ms_out)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = (int)hot_k * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaSuccess;
    float best = 1e30f;
    for (int r = 0; r < reps + 2 && e == cudaSuccess; r++) {
        cudaEventRecord(e0);
        if (U == 4) {
            cudaFuncSetAttribute(probe_hot_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe_hot_kernel<4><<<sms * blocks_per_sm, threads, smem>>>(cols, vals, n, x, out, hot_k);
        } else {
            cudaFuncSetAttribute(probe_hot_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe_hot_kernel<9><<<sms * blocks_per_sm, threads, smem>>>(cols, vals, n, x, out, hot_k);
        }
        e = cudaGetLastError();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) e = cudaGetLastError();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        fprintf(stderr, "probe_hot_run: %s\n", cudaGetErrorString(e));
        return -1;
    }
    *ms_out = best;
    return 0;
}

extern "C" void probe_set_chunk(int chunk_el) { g_chunk_el = chunk_el; }

extern "C" int probe_run(int variant, int U, const int32_t* cols, const double* vals, int64_t n,
                         const double* x, double* out, int blocks_per_sm, uint32_t col_mask,
                         int carveout, int reps, float* ms_out)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaSuccess;
    float best = 1e30f;
    for (int r = 0; r < reps + 2 && e == cudaSuccess; r++) {
        cudaEventRecord(e0);
        if (U == 4) e = dispatch<4>(variant, cols, vals, n, x, out, grid, col_mask, carveout);
        else if (U == 9) e = dispatch<9>(variant, cols, vals, n, x, out, grid, col_mask, carveout);
        else if (U == 18) e = dispatch<18>(variant, cols, vals, n, x, out, grid, col_mask, carveout);
        else e = cudaErrorInvalidValue;
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) e = cudaGetLastError();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms; // two warm-up passes
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        fprintf(stderr, "probe_run: %s\n", cudaGetErrorString(e));
        return -1;
    }
    *ms_out = best;
    return 0;
}
