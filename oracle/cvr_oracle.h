/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference CVR path.
 *
 * Nothing under oracle/ is linked, imported or executed by the product
 * (cvr_b200/, include/).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, and only as the
 * checker or the reported CPU baseline.
 *
 * Parity status: PINNED.  The port is checked bit-for-bit against the
 * unmodified reference compiled from /root/reference/spmv.cpp
 * (oracle/_ref/libcvr_ref.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py
 * in the build container, and against fixtures that reference build produced
 * (tests/golden/, generator script committed) everywhere else.  The reference
 * itself ships no golden vectors.
 *
 * All citations are /root/reference/spmv.cpp line numbers.
 */
#ifndef CVR_ORACLE_H
#define CVR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define CVR_ORACLE_LANES 8 /* SIMD_LEN for TYPE_DOUBLE, spmv.cpp:43 */

/* ints in the shared record array, as main() allocates it (spmv.cpp:1806) */
long long cvr_oracle_record_ints(int n_rows, int n_chunks);

/* int offset of chunk t's record region (spmv.cpp:709, :1112) */
long long cvr_oracle_record_offset(int chunk, int first_row);

/* CSR -> CVR, restating pre_processing (spmv.cpp:565-1014) for omega=1.
 * Inputs: 1-based CSR as readMatrix builds it: row_delim[n_rows+2], nnz a
 * multiple of 16.  Outputs have the reference's sizes and strides:
 *   cvr_vals[nnz], cvr_cols[nnz], record[cvr_oracle_record_ints()],
 *   nnz_rows[4T], final_2[16T] (8 used per chunk), split[2T] (zeroed here).
 * Bytes the reference never writes are left untouched (pre-fill with a
 * sentinel to see the written extent).  Deviation: when a chunk ends without
 * the reference ever storing final_2 (span <= 8 rows and no steal), the port
 * leaves it untouched too unless fill_missing_tail != 0, in which case it
 * stores the lanes' final row ids (the intended value).
 * Returns 0, or -1 on invalid arguments (T > nnz/16, nnz % 16 != 0). */
int cvr_oracle_convert(int n_chunks, int nnz, int n_rows,
                       const double* csr_val, const int* csr_col, const int* row_delim,
                       double* cvr_vals, int* cvr_cols, int* record, int* nnz_rows,
                       int* final_2, int* split, int fill_missing_tail);

/* y = A*x over the CVR arrays: the INTENDED semantics of spmv_compute_kernel
 * (spmv.cpp:1016-1667; paper Alg. 4) as validated in SURVEY.md section 8a-R3.
 * It differs from what the reference executes only where the reference is
 * wrong (chunks spanning <= 8 rows, split0/8 == split1/8).
 * y has n_rows+1 entries and is zeroed here first.  Per-lane FMA order is the
 * reference's (one fused multiply-add per step, spmv.cpp:1233). */
int cvr_oracle_spmv(int n_chunks, int n_rows,
                    const double* cvr_vals, const int* cvr_cols, const int* record,
                    const int* nnz_rows, const int* final_2, const int* split,
                    const double* x, double* y);

/* Scalar CSR SpMV, the reference's self-check loop (spmv.cpp:1843-1850)
 * extended to rows 0..n_rows inclusive.  abs_sum (optional) receives
 * sum_j |a_ij * x_j|, the normaliser of the 1e-12 parity bound. */
void cvr_oracle_csr_spmv(int n_rows, const double* csr_val, const int* csr_col,
                         const int* row_delim, const double* x, double* y, double* abs_sum);

/* Matrix Market -> CSR restating readMatrix (spmv.cpp:311-535) with its
 * observable quirks: indices stay 1-based (:437-438), values parsed as float
 * (:65, :432), pattern value = running_index % 13 (:417), symmetric mirrored
 * (:443-449), comments skipped only before the size line (:377-383), a last
 * line without '\n' is dropped (:411), nnz padded to x16 with zero-valued
 * copies of the last entry (:457, :474-482), stable (row,col) sort (:485).
 * ref_last_delim != 0 reproduces the reference's off-by-one tail
 * row_delim[k] = nnz-1 (:522-526); 0 stores the correct nnz.
 * Arrays are malloc'd; free with cvr_oracle_free.  Returns 0 or a negative
 * error (-1 open, -2 not a matrix, -3 dense). */
int cvr_oracle_read_mtx(const char* path, int ref_last_delim,
                        double** val, int** col, int** row_delim,
                        int* nnz_padded, int* nnz_file, int* n_rows, int* n_cols);
void cvr_oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
