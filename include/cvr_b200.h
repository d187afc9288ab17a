/* cvr_b200.h -- C ABI of the B200-native CVR SpMV path.
 *
 * This is the drop-in boundary for the one hot path of puckbee/CVR: the
 * CSR -> CVR conversion and the CVR SpMV loop.  The reference has no library
 * API; its main() (spmv.cpp:1675) calls two C++ functions with caller-owned
 * buffers.  Each entry point below names the reference interface it replaces
 * (all citations: /root/reference/spmv.cpp).  INTEGRATION.md shows the patch a
 * maintainer of the reference would apply to call this library instead.
 *
 * Conventions (identical to the reference, SURVEY.md 8b "data conventions"):
 *   - fp64 values, int32 column indices, 1-BASED rows and columns as
 *     readMatrix leaves them (:437-438): row 0 / column 0 are phantoms.
 *   - row_delim has n_rows+2 entries, row r occupies [row_delim[r], row_delim[r+1]).
 *   - nnz is padded to a multiple of 16 (:457).
 *   - x has n_cols+1 entries, y has n_rows+1 entries (index 0 unused / 0.0).
 *   - n_chunks is the reference's numThreads: chunk t of the CVR layout is what
 *     OpenMP thread t owns in the reference, so the structure arrays are
 *     comparable bit for bit at equal n_chunks.
 *
 * Every function returns 0 on success or a negative cvr_status_t; the message
 * is available from cvr_last_error() (thread-local).  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * CVR_ERR_CUDA.
 */
#ifndef CVR_B200_H
#define CVR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVR_B200_ABI_VERSION 3
#define CVR_LANES 8 /* SIMD_LEN for fp64, spmv.cpp:43 -- fixed by the bit-exact contract */

typedef enum cvr_status {
    CVR_OK = 0,
    CVR_ERR_INVALID = -1, /* bad argument (NULL, nnz % 16, n_chunks > nnz/16, ...) */
    CVR_ERR_CUDA = -2,    /* CUDA runtime failure, or no device */
    CVR_ERR_RANGE = -3,   /* a value does not fit the int32 export layout */
    CVR_ERR_STATE = -4    /* call order violated */
} cvr_status_t;

typedef struct cvr_handle cvr_handle_t;

/* The CSR readMatrix produces (out-params of spmv.cpp:311-312).  Exactly one of
 * row_delim32 / row_delim64 is non-NULL; the 64-bit form exists because the
 * reference's `int` delimiters cannot address nnz >= 2^31. */
typedef struct cvr_csr {
    int64_t n_rows;
    int64_t n_cols;
    int64_t nnz;                /* padded, multiple of 16 */
    const double* val;          /* [nnz]   h_val            */
    const int32_t* col;         /* [nnz]   h_cols, 1-based  */
    const int32_t* row_delim32; /* [n_rows+2] h_rowDelimiters, or NULL */
    const int64_t* row_delim64; /* [n_rows+2], or NULL */
} cvr_csr_t;

/* Host buffers of the reference's own sizes (allocation at spmv.cpp:1793-1817)
 * filled by cvr_export for the bit-exact gate.  Any pointer may be NULL to skip
 * that array. */
typedef struct cvr_arrays {
    double* vals;      /* [nnz]                vPack_vec_vals   */
    int32_t* cols;     /* [nnz]                vPack_vec_cols   */
    int32_t* record;   /* [cvr_record_ints()]  vPack_vec_record, (pos,wb) pairs; chunk t's
                          region starts at int offset (2*(32t+first_row))/16*16 (:709) */
    int32_t* nnz_rows; /* [4*n_chunks]         vPack_nnz_rows: s, e, first_row, last_row */
    int32_t* final_2;  /* [16*n_chunks]        vPack_vec_final_2: 8 tail rows used per chunk */
    int32_t* split;    /* [2*n_chunks]         vPack_split */
} cvr_arrays_t;

typedef struct cvr_info {
    int64_t n_rows, n_cols, nnz; /* nnz padded */
    int32_t n_chunks;
    int32_t device;
    int64_t n_records;       /* (pos,wb) pairs incl. the 8 terminators per chunk */
    int64_t record_ints;     /* size of the exported record array in ints */
    int64_t algorithmic_bytes; /* per SpMV: 12*nnz + 8*n_records + 56*T + 8*(n_cols+1) + 8*(n_rows+1) */
    double convert_seconds;  /* device conversion only (CUDA events) */
    double create_seconds;   /* upload + conversion, host wall clock */
    int64_t kernel_launches; /* kernels this handle has launched so far */
    int64_t device_bytes;    /* device memory owned by the handle */
    double convert_kernel_seconds; /* the schedule + permute kernels alone (CUDA events; what B_conv is held against) */
    double row_lists_seconds;      /* building the clear lists after them (host clock: it allocates and synchronises) */
} cvr_info_t;

/* CSR owned by the library, produced by cvr_read_matrix_market. */
typedef struct cvr_host_csr {
    int64_t n_rows, n_cols;
    int64_t nnz;      /* padded to a multiple of 16 */
    int64_t nnz_file; /* entries before padding (mirrored entries included) */
    double* val;
    int32_t* col;
    int32_t* row_delim32; /* set when nnz < 2^31, else NULL */
    int64_t* row_delim64; /* set when nnz >= 2^31, else NULL */
} cvr_host_csr_t;

/* flags for cvr_read_matrix_market */
#define CVR_MM_REF_LAST_DELIM 1 /* reproduce row_delim[k] = nnz-1 after the last row (spmv.cpp:522-526) */
#define CVR_MM_KEEP_LAST_LINE 2 /* keep a final line that lacks '\n' (the reference drops it, :411) */

int cvr_abi_version(void);
const char* cvr_last_error(void);

/* Create the CUDA context on `device` (so that later timings exclude it). */
int cvr_device_init(int device);

/* Replaces readMatrix (spmv.cpp:311-535): Matrix Market coordinate file -> the 1-based,
 * x16-padded CSR described above, with the reference's observable ingest semantics
 * (item 7, the off-by-one last delimiter, only on request).
 * Host only.  Release with cvr_free_host_csr. */
int cvr_read_matrix_market(const char* path, int flags, cvr_host_csr_t* out);
void cvr_free_host_csr(cvr_host_csr_t* csr);

/* ints in the record array for (n_rows, n_chunks): 2*(n_rows+240+32*n_chunks), spmv.cpp:1806 */
int64_t cvr_record_ints(int64_t n_rows, int32_t n_chunks);

/* A chunk count that fills `device` (multiple of the SM count, ~KBs of stream per
 * chunk).  What `numThreads = 0` means on the command line. */
int cvr_auto_chunks(int64_t nnz, int device, int32_t* n_chunks);

/* Replaces pre_processing (spmv.cpp:565, called at :1857): uploads the host CSR
 * to `device`, converts it there and keeps the CVR arrays on the device.
 * n_chunks <= nnz/16 (note (v) of SURVEY 8a-R2); n_chunks == 0 picks cvr_auto_chunks. */
int cvr_create(const cvr_csr_t* csr_host, int32_t n_chunks, int device, cvr_handle_t** out);

/* Same, for a CSR already resident on `device` (device pointers in `csr_dev`):
 * the entry the multi-GPU host and the synthetic generators use.  The CSR is
 * only read during the call. */
int cvr_create_from_device(const cvr_csr_t* csr_dev, int32_t n_chunks, int device,
                           cvr_handle_t** out);

/* Replaces spmv_compute_kernel (spmv.cpp:1016, called at :1882) with HOST
 * vectors: copies x (n_cols+1) in, runs `iters` SpMVs (each zeroes y inside the
 * timed region, unlike spmv.cpp:1026-1033), copies y (n_rows+1) out.
 * seconds_per_iter (optional) is the device time of the iteration loop / iters,
 * the quantity the reference prints at :1662. */
int cvr_spmv(cvr_handle_t* h, const double* x_host, double* y_host, int32_t iters,
             double* seconds_per_iter);

/* One SpMV on device vectors, enqueued on `cuda_stream` (a cudaStream_t, NULL =
 * default stream) without synchronising: y_dev[0..n_rows] = A * x_dev. */
int cvr_spmv_device(cvr_handle_t* h, const double* x_dev, double* y_dev, void* cuda_stream);

/* ---- iterated SpMV on row shards (multi-GPU; no counterpart in the reference) ----
 * The matrix is row-sharded by nnz, every GPU holds a replicated x.  cvr_spmv_publish is
 * cvr_spmv_device plus the exchange: each finished y row is also stored into dst[0..n_dst-1][
 * row_offset + row], the x vectors (own and peer-mapped) that the NEXT iteration reads, from
 * inside the SpMV kernel (accumulated rows follow from a small kernel right after it).  One
 * process per GPU: allocate the x buffers and a flag array with cvr_peer_alloc, exchange the 64-byte
 * handles through any channel (e.g. torch.distributed), map the peers' with cvr_peer_open, and
 * separate iterations with cvr_peer_barrier (double-buffer x: iteration k reads buffer k%2 and
 * publishes into buffer (k+1)%2). */
#define CVR_MAX_PEERS 8
typedef struct cvr_publish {
    int32_t n_dst;       /* 1..CVR_MAX_PEERS destinations, own buffer included */
    int32_t self;        /* index of this GPU's own buffer in dst[] (required with mode bit 2) */
    int32_t mode;        /* bit 0: reserved;
                            bit 1: skip the 0.0 for never-written rows (set from the 3rd iteration on);
                            bit 3: no programmatic dependent launches (set when one device carries two shards);
                            bit 2: y_dev IS this GPU's own slice of the next x (x_next + row_offset), so
                                   own rows need no copy; then leave this GPU's bit out of `needs`, and
                                   set clear_next = x_current + row_offset (the next iteration's y) */
    int64_t row_offset;  /* global row id = row_offset + local row (first cut - 1) */
    const uint8_t* needs; /* device array, needs[local row] bit p set: destination p reads x[that row]
                             (its shard has a nonzero in that column); NULL = every destination */
    const uint8_t* chunk_any; /* device array from cvr_chunk_needs (chunks with nothing to send skip
                                 the push), or NULL */
    double* clear_next;  /* NULL, or the y of the NEXT iteration (see mode bit 2) */
    double* dst[CVR_MAX_PEERS];
    double* multicast;   /* NULL, or ONE NVSwitch multicast address of the next x that maps every GPU's buffer
                            (own included; e.g. torch symmetric memory's multicast_ptr): a finished row is then
                            published with a single multimem.st instead of n_dst - 1 peer stores, always in
                            contiguous ranges.  Requires mode bit 2 clear (y_dev is a buffer of its own) and
                            needs = chunk_any = NULL.  ABI v3. */
} cvr_publish_t;
/* One iteration = at most three kernels on cuda_stream: [clear the accumulated rows of y, unless
 * y_is_clear] + the SpMV sweep that pushes finished rows to every dst + one epilogue kernel
 * (publishes the accumulated rows, clears them in y again for the next sweep -- so pass
 * y_is_clear = 1 from the second iteration on, as long as y_dev is not touched in between --
 * and runs the all-to-all flag barrier: flag_arrays / rank / n_ranks / epoch as cvr_peer_barrier). */
int cvr_spmv_publish(cvr_handle_t* h, const double* x_dev, double* y_dev, const cvr_publish_t* pub,
                     void* const* flag_arrays, int32_t rank, int32_t n_ranks, uint32_t epoch,
                     int32_t y_is_clear, void* cuda_stream);
/* The flag barrier spins for at most CVR_BARRIER_TIMEOUT_MS (default 3000 ms) so that a lost peer
 * cannot hang the GPU; a timeout is recorded on the device.  Call this after synchronising (end of an
 * iteration loop): CVR_ERR_STATE if a barrier of this handle timed out since the last call -- the
 * x vectors are then stale and the results must be discarded. */
int cvr_check_async_error(cvr_handle_t* h);
/* used_dev[c] = 1 for every column id c (0..n_cols) that occurs in the shard: the x entries this
 * GPU actually reads.  Exchanged once, it lets every GPU publish a row only to the peers that
 * read it (a banded matrix then sends halos, not the whole vector). */
int cvr_column_footprint(cvr_handle_t* h, uint8_t* used_dev, void* cuda_stream);
/* Finishes a footprint for cvr_spmv_publish: clears needs_dev[row] for the rows nothing ever writes (the
 * epilogue publishes their 0.0, the sweep's range pushes skip them), then chunk_any_dev[t] (n_chunks bytes)
 * = OR of needs_dev over the row range of chunk t. */
int cvr_chunk_needs(cvr_handle_t* h, uint8_t* needs_dev, uint8_t* chunk_any_dev, void* cuda_stream);
int cvr_peer_alloc(int device, int64_t bytes, void** dev_ptr, unsigned char handle[64]); /* zero-filled */
int cvr_peer_open(int device, const unsigned char handle[64], void** dev_ptr);
int cvr_peer_close(int device, void* dev_ptr);
int cvr_peer_free(int device, void* dev_ptr);
/* flag_arrays[p] = rank p's flag array (n_ranks uint32, from cvr_peer_alloc / cvr_peer_open);
 * epoch must increase by one per call.  Enqueued on cuda_stream. */
int cvr_peer_barrier(int device, void* const* flag_arrays, int32_t rank, int32_t n_ranks, uint32_t epoch,
                     void* cuda_stream);

/* ---- the same path on 1..8 GPUs from ONE process (the host the `spmv.cvr` CLI uses with CVR_DEVICES) ----
 * cvr_create_sharded cuts the host CSR into n_devices contiguous row ranges of equal nnz (snapped to row
 * starts, the bisection of spmv.cpp:631-650), uploads and converts each shard on its device (what
 * pre_processing does per OpenMP thread slice, one level up) and prepares the exchange.  A device may be
 * listed more than once (several shards on one GPU).
 * cvr_sharded_spmv replaces spmv_compute_kernel (spmv.cpp:1016) with HOST vectors:
 *   feed_y_to_x == 0: `iters` SpMVs with the same x (the reference's loop, :1024-1034); no communication;
 *                     y_host[1..n_rows] = A x.
 *   feed_y_to_x != 0: `iters` iterations of x <- A x (square A), one exchange per iteration -- fused into
 *                     the sweep over NVLink peer memory (CVR_SHARD_PEER) or an NCCL all-gather after it
 *                     (CVR_SHARD_NCCL); y_host[1..n_rows] = the last iterate.
 * seconds_per_iter: host wall clock around the iteration loop (all devices synchronised on both sides). */
typedef struct cvr_sharded cvr_sharded_t;
#define CVR_SHARD_PEER 0  /* exchange fused into the sweep kernel over peer memory (default) */
#define CVR_SHARD_NCCL 1  /* NCCL broadcasts after the sweep (libnccl.so.2 is loaded on first use) */
#define CVR_SHARD_DENSE 2 /* CVR_SHARD_PEER: publish every row to every device, not only to its readers */
typedef struct cvr_sharded_info {
    int32_t n_parts;
    int32_t exchange;                 /* 0 = peer, 1 = NCCL */
    int64_t n_rows, n_cols, nnz;
    int32_t device[8];
    int64_t row_begin[8], row_end[8]; /* part g owns global rows [row_begin, row_end), 1-based */
    int64_t part_nnz[8];              /* padded nnz of the shard */
    int32_t part_chunks[8];
    int64_t peer_bytes_per_iter[8];   /* bytes part g stores into other parts' x per iteration (sparse exchange) */
    double create_seconds;            /* partition + upload + conversion + exchange set-up, host wall clock */
    double convert_seconds;           /* sum of the per-device conversion times (CUDA events) */
    int64_t kernel_launches;
} cvr_sharded_info_t;
int cvr_create_sharded(const cvr_csr_t* csr_host, int32_t n_chunks_per_device, const int* devices, int n_devices,
                       int flags, cvr_sharded_t** out);
int cvr_sharded_spmv(cvr_sharded_t* s, const double* x_host, double* y_host, int32_t iters, int32_t feed_y_to_x,
                     double* seconds_per_iter);
int cvr_sharded_get_info(cvr_sharded_t* s, cvr_sharded_info_t* info);
/* the single-GPU handle of one part (owned by `s`), e.g. for cvr_export of a shard */
int cvr_sharded_part(cvr_sharded_t* s, int part, cvr_handle_t** handle);
void cvr_sharded_destroy(cvr_sharded_t* s);

/* Replaces the reference's self-check (spmv.cpp:1843-1850 scalar CSR SpMV, :1916-1938 comparison) on
 * the device: y_dev against the CSR product of `csr_dev` (DEVICE pointers) and x_dev, row by row,
 * |y_r - sum_j a_rj x_j| <= rel_tol * sum_j |a_rj x_j| for rows 1..n_rows (row 0, the phantom, must be
 * 0.0 when check_row0 != 0).  Unlike the reference it checks the last row and uses a relative bound.
 * Outputs: number of rows outside the bound, the largest relative error seen, the first bad row (-1). */
int cvr_verify_csr(const cvr_csr_t* csr_dev, int device, const double* x_dev, const double* y_dev,
                   double rel_tol, int check_row0, int64_t* rows_failing, double* max_rel,
                   int64_t* first_bad_row);

/* Bit-exact gate: copy the CVR structure arrays back in the reference layout. */
int cvr_export(cvr_handle_t* h, cvr_arrays_t* host_out);

/* Save / load a converted matrix (device CVR arrays, chunk descriptors, row lists) so that the
 * conversion is paid once -- the paper's "iterations to amortise" drops to the load time.  The file
 * is a little-endian dump tied to this library version; cvr_load fails on anything else. */
int cvr_save(cvr_handle_t* h, const char* path);
int cvr_load(const char* path, int device, cvr_handle_t** out);

int cvr_get_info(cvr_handle_t* h, cvr_info_t* info);

/* Device pointers of the handle's x / y scratch vectors (n_cols+1 / n_rows+1
 * doubles), for callers that iterate on the device. */
int cvr_device_vectors(cvr_handle_t* h, double** x_dev, double** y_dev);

/* Name of the sweep geometry picked for this matrix ("tile7x7": short or skewed rows, "tile11x5": long
 * regular rows; CVR_SPMV_KERNEL overrides). */
const char* cvr_kernel_variant(cvr_handle_t* h);

/* Device pointers of the converted matrix itself (read-only views, valid until cvr_destroy):
 * vals [nnz] f64, cols [nnz] i32 in CVR order, record [record_ints] i32.  For tools that run their
 * own kernels over the CVR arrays (tools/probe, tools/compare.py); any out-pointer may be NULL. */
int cvr_device_arrays(cvr_handle_t* h, const double** vals_dev, const int32_t** cols_dev,
                      const int32_t** record_dev);

/* Measurement aid: while enabled, every SpMV launch is bracketed by a CUDA event pair on
 * its stream (after y has been cleared), so the SpMV kernel's own device time can be
 * reported next to the whole-step time.  cvr_get_kernel_timing waits for the recorded
 * launches, returns their summed kernel time and count, and resets the counters. */
int cvr_set_kernel_timing(cvr_handle_t* h, int enabled);
int cvr_get_kernel_timing(cvr_handle_t* h, double* total_seconds, int64_t* launches);

void cvr_destroy(cvr_handle_t* h);

#ifdef __cplusplus
}
#endif
#endif /* CVR_B200_H */
