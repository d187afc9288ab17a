// Internal declarations shared by the CUDA translation units of libcvr_b200.
// Not part of the ABI (include/cvr_b200.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define CVR_W 8              // lanes per step, fixed by the bit-exact contract (spmv.cpp:43)
#define CVR_WIN 32           // one warp covers 4 steps x 8 lanes = 32 consecutive CVR elements
#define CVR_SEG_STRIDE 48    // scratch segment entries reserved per chunk on top of its row span
#define CVR_SEG_END 0x7fffffff

// Per-chunk descriptor kept on the device: everything the reference keeps in
// vPack_nnz_rows[4], vPack_split[2] and vPack_vec_final_2[8] for one OpenMP
// thread (spmv.cpp:690-694, :826-891, :853/:893), with a 64-bit start offset.
// 64 bytes, one per chunk.
struct __align__(16) CvrChunk {
    int64_t start;     // first element of the chunk in vals/cols   (nnz_rows[4t])
    int32_t len;       // elements in the chunk, multiple of 16     (nnz_rows[4t+1] - start)
    int32_t first_row; // nnz_rows[4t+2]
    int32_t last_row;  // nnz_rows[4t+3]
    int32_t split0;    // vPack_split[2t]   position where the (shared) first row ended, 0 = never
    int32_t split1;    // vPack_split[2t+1] last feeding position, -1 = steal-only chunk, 0 = no event
    int32_t n_rec;     // record pairs before the eight pos=-1 terminators
    int32_t tail[CVR_W]; // vPack_vec_final_2[16t .. 16t+7]
};
static_assert(sizeof(CvrChunk) == 64, "CvrChunk must stay one 64-byte line");

// int offset of chunk t's record region in the reference layout (spmv.cpp:709)
__host__ __device__ inline int64_t cvr_record_offset(int64_t chunk, int64_t first_row)
{
    return (2 * (32 * chunk + first_row)) / 16 * 16;
}

// entry offset of chunk t's scratch segment list (conversion only)
__host__ __device__ inline int64_t cvr_segment_offset(int64_t chunk, int64_t first_row)
{
    return CVR_SEG_STRIDE * chunk + first_row;
}

#define CVR_MAX_PEERS 8

// Where finished y rows are published for the NEXT iteration of an iterated, row-sharded SpMV:
// the x vectors of this GPU and of its peers (peer-mapped device pointers).  Passed by value.
struct CvrPublish {
    int32_t n_dst;                 // 0: do not publish
    int32_t self;                  // index of THIS GPU's own buffer in dst[] (used with mode bit 2)
    int32_t mode;                  // bit 0: reserved (round 1: per-row stores at emit)
                                   // bit 1: do not re-publish 0.0 for the never-written rows
                                   // bit 2: y IS this GPU's slice of the next x (no local copy; y[0] is foreign)
                                   // bit 3: no programmatic dependent launches (two shards share one device)
    int64_t row_offset;            // global row = row_offset + local row
    const uint8_t* needs;          // needs[local row] bit p: destination p reads that x entry (NULL: all do)
    const uint8_t* chunk_any;      // chunk_any[t] != 0: some row of chunk t's range has a reader elsewhere (NULL: all)
    double* clear_next;            // y of the NEXT sweep: its accumulated rows are cleared by the epilogue
                                   // (NULL: y itself).  Set when y aliases this GPU's slice of the next x.
    double* dst[CVR_MAX_PEERS];
    double* mc;                    // multicast address of the next x (all GPUs, own included) or NULL: one
                                   // multimem.st per published row instead of a store per destination
};

struct CvrBarrier {
    uint32_t* flags[CVR_MAX_PEERS]; // flags[p]: rank p's flag array (n_ranks words), peer-mapped
    int32_t rank, n_ranks;
    uint32_t epoch;
    unsigned int* error;     // device word: set to the epoch of the first barrier that timed out (0 = none)
    long long timeout_cycles; // bound of the spin in SM clocks
};

struct CvrConvertArgs {
    // CSR on the device (1-based, n_rows+2 delimiters); one of rd32 / rd64
    const double* csr_val;
    const int32_t* csr_col;
    const int32_t* rd32;
    const int64_t* rd64;
    int64_t nnz;
    int64_t n_rows;
    int32_t n_chunks;
    // outputs
    double* cvr_vals;
    int32_t* cvr_cols;
    int32_t* record;   // reference layout, pre-filled with 0xff
    CvrChunk* chunks;
    // scratch
    int2* segments;    // (pos, src) entries, CVR_SEG_STRIDE*T + n_rows + slack
    int32_t* seg_count;
    uint32_t* row_bitmap; // (n_rows + 2) / 32 + 2 words: bit r = row r is not empty (NULL: warp scheduler)
};

// Rows that are ACCUMULATED (atomics) rather than stored once: chunk first rows that end while
// feeding (split0) and the eight tail rows of every chunk; and rows nothing writes: empty rows and
// the phantom row 0.  Only these need clearing before a SpMV.
struct CvrRowLists {
    int32_t* boundary = nullptr; // accumulated rows
    int32_t* empty = nullptr;    // never-written rows (incl. row 0)
    int32_t n_boundary = 0, n_empty = 0;
};
// builds the lists on `stream` (synchronises); returns kernels launched or <0
int cvr_build_row_lists(const CvrChunk* chunks, int32_t n_chunks, const int32_t* rd32, const int64_t* rd64,
                        int64_t n_rows, CvrRowLists* out, cudaStream_t stream);

// Launchers (each returns the number of kernels it launched, or <0 on launch failure)
int cvr_launch_convert(const CvrConvertArgs& a, cudaStream_t stream);
// repairs the reference's off-by-one trailing delimiters (nnz-1 -> nnz) in a DEVICE delimiter array we own
int cvr_launch_fix_last_delim(int32_t* rd32, int64_t* rd64, int64_t n_rows, int64_t nnz, cudaStream_t stream);
// ev_begin / ev_end (optional) bracket the SpMV kernel alone, after y has been cleared;
// [chunk_begin, chunk_end) restricts the sweep to a slab of chunks (chunk_end < 0: through the last chunk);
// chunk_queue (optional): two zero-initialised device words -- chunks beyond each warp's first are then handed
// out in index order from an atomic counter (the sweep resets both words itself when its last warp finishes)
int cvr_launch_spmv(int variant, const CvrChunk* chunks, int32_t n_chunks, const double* vals,
                    const int32_t* cols, const int32_t* record, const double* x, double* y,
                    int64_t n_rows, const CvrRowLists& rows, const CvrPublish* publish,
                    cudaStream_t stream, cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr,
                    const CvrBarrier* barrier = nullptr, unsigned int* done_counter = nullptr,
                    bool y_is_clear = false, int32_t chunk_begin = 0, int32_t chunk_end = -1,
                    unsigned int* chunk_queue = nullptr);
int cvr_launch_peer_barrier(const CvrBarrier& b, cudaStream_t stream);
// used[c] = 1 for every column id that occurs in cols[0..nnz)
// chunk_any[t] = OR of needs[first_row..last_row] of chunk t
int cvr_launch_chunk_needs(const CvrChunk* chunks, int32_t n_chunks, const CvrRowLists& rows, uint8_t* needs,
                           uint8_t* chunk_any, cudaStream_t stream);
int cvr_launch_column_footprint(const int32_t* cols, int64_t nnz, uint8_t* used, cudaStream_t stream);

// sweep geometry for a matrix (index into the variant table of cvr_spmv.cu; CVR_SPMV_KERNEL overrides)
int cvr_pick_sweep_variant(int64_t nnz, int64_t n_rows);
// resident warps per SM of that geometry (sizes the automatic chunk count)
int cvr_spmv_resident_warps_per_sm(int variant);
// its name ("tile7x5r", "tile11x5", ...)
const char* cvr_spmv_kernel_name(int variant);

// Device memory of the library: cudaMalloc underneath, freed blocks cached for the next matrix (up to 2 GB,
// CVR_POOL_KEEP_MB; CVR_NO_POOL=1 = plain cudaMalloc / cudaFree).  With cudaMalloc / cudaFree in the creation path the
// 12 allocations of ONE web-Google-sized matrix took 10-75 ms and a single cudaFree of scratch up to 556 ms, against
// 0.26 ms of conversion kernels (profiles/r02_create_trace.txt).  A block is freed only after the work that used it
// has completed (every call site synchronises first); the stream argument is unused and kept for the call sites.
cudaError_t cvr_dev_malloc(void** p, size_t bytes, cudaStream_t stream);
void cvr_dev_free(void* p, cudaStream_t stream);
void cvr_pool_setup(int device); // reads CVR_POOL_KEEP_MB

// force-load the kernels' module so the first timed call does not pay CUDA's lazy loading
void cvr_preload_convert_kernels();
void cvr_preload_spmv_kernels();
