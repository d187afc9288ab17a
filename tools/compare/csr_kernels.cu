// Comparators for tools/compare.py (NOT part of the product): the GPU stand-ins for the reference's
// solutions_for_comparison/ (CSR5, MKL, ...; run_comparison.sh:17-45) next to which the CVR path is judged.
//
//   csr_balanced_spmv   "CSR5-style": the nnz are cut into equal tiles regardless of rows (CSR5's idea,
//                       csr5/detail/...: tile = fixed number of nonzeros, row boundaries handled inside the
//                       tile), one warp per 512-nnz tile staged through shared memory with coalesced loads;
//                       lane t walks 16 consecutive nonzeros, rows that lie completely inside its range are
//                       stored, the partial first / last row of the range is added atomically.  Not the
//                       CSR5 format (no tile descriptors, no transposed tiles) -- a load-balanced CSR kernel
//                       of the same family, good enough as a yardstick.
//   csr_vector_spmv     the textbook one-warp-per-row CSR kernel (what MKL/CSR-I stand for on the CPU).
// Conventions as everywhere here: 1-based rows/columns, row_delim has n_rows+2 entries, x has n_cols+1.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int TILE = 512, PER_LANE = TILE / 32;

__global__ void __launch_bounds__(128)
csr_balanced_kernel(const int32_t* __restrict__ rd, const double* __restrict__ val, const int32_t* __restrict__ col,
                    int64_t n_rows, int64_t nnz, const double* __restrict__ x, double* __restrict__ y)
{
    __shared__ double s_val[4][TILE];
    __shared__ int32_t s_col[4][TILE];
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t n_tiles = (nnz + TILE - 1) / TILE;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * TILE;
#pragma unroll
        for (int k = 0; k < PER_LANE; k++) {
            const int64_t j = base + 32 * k + t;
            s_val[w][32 * k + t] = j < nnz ? val[j] : 0.0;
            s_col[w][32 * k + t] = j < nnz ? col[j] : 0;
        }
        __syncwarp();
        const int64_t j0 = base + (int64_t)PER_LANE * t, j1 = min(j0 + PER_LANE, nnz);
        if (j0 < j1) {
            // row of element j0: the last row whose start is <= j0
            int64_t lo = 0, hi = n_rows + 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if ((int64_t)rd[mid] <= j0) lo = mid;
                else hi = mid - 1;
            }
            int64_t row = lo;
            int64_t row_end = rd[row + 1];
            bool first = (int64_t)rd[row] < j0; // my first row started before my range: shared with the lane before
            double acc = 0.0;
            for (int64_t j = j0; j < j1; j++) {
                while (j >= row_end) { // row finished (empty rows in between are skipped: y is pre-zeroed)
                    if (first) atomicAdd(&y[row], acc);
                    else y[row] = acc;
                    first = false;
                    acc = 0.0;
                    row++;
                    row_end = rd[row + 1];
                }
                const int k = (int)(j - base);
                acc = fma(s_val[w][k], x[s_col[w][k]], acc);
            }
            // my last row: complete only if it ends exactly at the end of my range
            if (first || j1 < row_end) atomicAdd(&y[row], acc);
            else y[row] = acc;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
csr_vector_kernel(const int32_t* __restrict__ rd, const double* __restrict__ val, const int32_t* __restrict__ col,
                  int64_t n_rows, const double* __restrict__ x, double* __restrict__ y)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row > n_rows) return;
    const int t = threadIdx.x & 31;
    double acc = 0.0;
    for (int64_t j = (int64_t)rd[row] + t; j < (int64_t)rd[row + 1]; j += 32) acc = fma(val[j], x[col[j]], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (t == 0) y[row] = acc;
}

} // namespace

extern "C" int csr_balanced_spmv(const int32_t* rd, const double* val, const int32_t* col, int64_t n_rows, int64_t nnz,
                                 const double* x, double* y, void* stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMemsetAsync(y, 0, sizeof(double) * (size_t)(n_rows + 1), s); // partial rows are accumulated
    csr_balanced_kernel<<<sms * 8, 128, 0, s>>>(rd, val, col, n_rows, nnz, x, y);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int csr_vector_spmv(const int32_t* rd, const double* val, const int32_t* col, int64_t n_rows,
                               const double* x, double* y, void* stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t warps = n_rows + 1;
    csr_vector_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(rd, val, col, n_rows, x, y);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
