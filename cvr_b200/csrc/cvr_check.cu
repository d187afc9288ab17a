// Device-side self-check: y against a plain CSR product of the same matrix and x.
//
// Replaces the reference's verdict (/root/reference/spmv.cpp:1843-1850 computes a scalar CSR SpMV,
// :1916-1938 compares it with the CVR result at 1e-3 absolute and skips the last row) at sizes where
// the host loop is too slow, with the criterion of BASELINE.json: per row
// |y_r - sum_j a_rj x_j| <= rel_tol * sum_j |a_rj x_j|, every row 1..n_rows checked; a row with no
// entries must read exactly 0.  The kernel is deliberately simple -- one warp per row, lanes stride
// over the row, shuffle tree -- and shares nothing with the CVR sweep (different format, different
// order of summation), so agreement is evidence, not tautology.  tests/test_gpu_parity.py pins it
// against the oracle's scalar loop at small sizes.
#include "../../include/cvr_b200.h"
#include "cvr_internal.h"

#include <cstring>

int cvr_set_error(int code, const char* fmt, ...);

namespace {

struct CheckOut {
    unsigned long long rows_failing;
    unsigned long long max_rel_bits; // non-negative doubles order like their bit patterns
    unsigned long long first_bad_row;
};

template <typename RdT>
__global__ void __launch_bounds__(256)
cvr_check_rows_kernel(const RdT* __restrict__ rd, const double* __restrict__ val, const int32_t* __restrict__ col,
                      int64_t n_rows, const double* __restrict__ x, const double* __restrict__ y, double rel_tol,
                      int check_row0, CheckOut* __restrict__ out)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // 0 .. n_rows
    if (row > n_rows) return;
    const int t = threadIdx.x & 31;
    const int64_t a = (int64_t)rd[row], b = (int64_t)rd[row + 1];
    double sum = 0.0, mag = 0.0;
    for (int64_t j = a + t; j < b; j += 32) {
        const double p = val[j] * x[col[j]];
        sum += p;
        mag += fabs(p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mag += __shfl_xor_sync(0xffffffffu, mag, o);
    }
    if (t != 0 || (row == 0 && !check_row0)) return;
    const double got = y[row];
    const double err = fabs(got - sum);
    const bool bad = !(err <= rel_tol * mag); // also catches NaN; an empty row must be exactly 0
    const double rel = mag > 0.0 ? err / mag : (err > 0.0 ? 1.0 : 0.0);
    if (rel > 0.0 && rel == rel) atomicMax(&out->max_rel_bits, (unsigned long long)__double_as_longlong(rel));
    if (bad) {
        atomicAdd(&out->rows_failing, 1ull);
        atomicMin(&out->first_bad_row, (unsigned long long)row);
    }
}

} // namespace

extern "C" int cvr_verify_csr(const cvr_csr_t* csr_dev, int device, const double* x_dev, const double* y_dev,
                              double rel_tol, int check_row0, int64_t* rows_failing, double* max_rel,
                              int64_t* first_bad_row)
{
    if (!csr_dev || !x_dev || !y_dev || !rows_failing)
        return cvr_set_error(CVR_ERR_INVALID, "NULL argument to cvr_verify_csr");
    if ((csr_dev->row_delim32 == nullptr) == (csr_dev->row_delim64 == nullptr))
        return cvr_set_error(CVR_ERR_INVALID, "exactly one of row_delim32 / row_delim64 must be set");
    if (cudaSetDevice(device) != cudaSuccess) return cvr_set_error(CVR_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    CheckOut* d_out = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d_out), sizeof(CheckOut)) != cudaSuccess)
        return cvr_set_error(CVR_ERR_CUDA, "cudaMalloc failed in cvr_verify_csr");
    CheckOut h{0ull, 0ull, ~0ull};
    cudaMemcpy(d_out, &h, sizeof(h), cudaMemcpyHostToDevice);
    const int64_t warps = csr_dev->n_rows + 1;
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    if (csr_dev->row_delim64)
        cvr_check_rows_kernel<int64_t><<<blocks, 256>>>(csr_dev->row_delim64, csr_dev->val, csr_dev->col,
                                                        csr_dev->n_rows, x_dev, y_dev, rel_tol, check_row0, d_out);
    else
        cvr_check_rows_kernel<int32_t><<<blocks, 256>>>(csr_dev->row_delim32, csr_dev->val, csr_dev->col,
                                                        csr_dev->n_rows, x_dev, y_dev, rel_tol, check_row0, d_out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return cvr_set_error(CVR_ERR_CUDA, "cvr_verify_csr failed: %s", cudaGetErrorString(e));
    *rows_failing = (int64_t)h.rows_failing;
    if (max_rel) {
        long long bits = (long long)h.max_rel_bits;
        double v;
        memcpy(&v, &bits, sizeof(v));
        *max_rel = v;
    }
    if (first_bad_row) *first_bad_row = h.rows_failing ? (int64_t)h.first_bad_row : -1;
    return CVR_OK;
}
