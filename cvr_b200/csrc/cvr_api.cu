// C ABI of libcvr_b200 (include/cvr_b200.h): handle, device memory, timing.
// The compute lives in cvr_convert.cu and cvr_spmv.cu; there is no host fallback.
#include "../../include/cvr_b200.h"
#include "cvr_internal.h"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

thread_local char g_cvr_err[512] = "";

// shared with cvr_mm_reader.cpp
int cvr_set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_cvr_err, sizeof(g_cvr_err), fmt, ap);
    va_end(ap);
    return code;
}
#define fail cvr_set_error

namespace {

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(CVR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// bound of the peer-barrier spin: CVR_BARRIER_TIMEOUT_MS (default 3000) in SM clocks of `device`.
// Computed once per device: cudaDevAttrClockRate is a driver query that takes about a millisecond, and
// this is called on every iteration of the multi-GPU loop (the first N = 2 runs of round 2 spent 1-5 ms
// per step here, profiles/r02_bench_n2_diagnostics.txt).
long long barrier_timeout_cycles(int device)
{
    static long long cache[64] = {};
    if (device >= 0 && device < 64 && cache[device] > 0) return cache[device];
    long long ms = 3000;
    if (const char* e = getenv("CVR_BARRIER_TIMEOUT_MS")) {
        const long long v = atoll(e);
        if (v > 0) ms = v;
    }
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) != cudaSuccess || khz <= 0) khz = 1965000;
    const long long cycles = ms * (long long)khz;
    if (device >= 0 && device < 64) cache[device] = cycles;
    return cycles;
}

double wall_seconds()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// CVR_CREATE_TRACE=1: wall time of every phase of cvr_create* on stderr (where does "Pre-processing time" go?)
struct PhaseTrace {
    bool on;
    double t;
    PhaseTrace() : on(false), t(0.0)
    {
        const char* e = getenv("CVR_CREATE_TRACE");
        on = e && *e == '1';
        if (on) t = wall_seconds();
    }
    void mark(const char* what)
    {
        if (!on) return;
        const double now = wall_seconds();
        fprintf(stderr, "[cvr create] %-34s %9.3f ms\n", what, (now - t) * 1e3);
        t = now;
    }
};

bool pool_enabled()
{
    static const bool on = [] {
        const char* e = getenv("CVR_NO_POOL");
        return !(e && *e == '1');
    }();
    return on;
}

// ---- pinned staging ring for the upload of a HOST CSR: pageable cudaMemcpy runs at 3-10 GB/s on the test boxes
// (one driver thread copies into its own staging buffer); here OpenMP threads fill pinned slots while the DMA of
// the previous slots is in flight.
struct StagingRing {
    static constexpr int SLOTS = 4;
    static constexpr size_t SLOT_BYTES = (size_t)8 << 20;
    char* buf[SLOTS] = {};
    cudaEvent_t ev[SLOTS] = {};
    bool used[SLOTS] = {};
    int next = 0;
    bool ok = false;
    std::mutex busy;
};

StagingRing* staging_ring()
{
    static StagingRing ring;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* e = getenv("CVR_NO_STAGING");
        if (e && *e == '1') return;
        for (int k = 0; k < StagingRing::SLOTS; k++) {
            if (cudaHostAlloc(reinterpret_cast<void**>(&ring.buf[k]), StagingRing::SLOT_BYTES, cudaHostAllocPortable) !=
                    cudaSuccess ||
                cudaEventCreateWithFlags(&ring.ev[k], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return; // no ring: plain cudaMemcpy
            }
        }
        ring.ok = true;
    });
    return ring.ok ? &ring : nullptr;
}

// host (pageable or pinned) -> device on `stream`; returns after the last piece has been QUEUED, the caller
// synchronises the stream before it lets go of `src`
cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t stream)
{
    StagingRing* ring = bytes >= ((size_t)1 << 20) ? staging_ring() : nullptr;
    std::unique_lock<std::mutex> lock;
    if (ring) {
        lock = std::unique_lock<std::mutex>(ring->busy, std::try_to_lock);
        if (!lock.owns_lock()) ring = nullptr; // another thread (sharded host) is uploading: do not queue behind it
    }
    if (!ring) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
        return e;
    }
    const char* s8 = static_cast<const char*>(src);
    char* d8 = static_cast<char*>(dst);
    for (size_t off = 0; off < bytes; off += StagingRing::SLOT_BYTES) {
        const size_t n = bytes - off < StagingRing::SLOT_BYTES ? bytes - off : StagingRing::SLOT_BYTES;
        const int k = ring->next;
        ring->next = (k + 1) % StagingRing::SLOTS;
        cudaError_t e = cudaSuccess;
        if (ring->used[k] && (e = cudaEventSynchronize(ring->ev[k])) != cudaSuccess) return e;
        const long pieces = (long)((n + ((size_t)1 << 20) - 1) >> 20);
#pragma omp parallel for schedule(static) num_threads(8)
        for (long q = 0; q < pieces; q++) {
            const size_t a = (size_t)q << 20, len = n - a < ((size_t)1 << 20) ? n - a : ((size_t)1 << 20);
            memcpy(ring->buf[k] + a, s8 + off + a, len);
        }
        if ((e = cudaMemcpyAsync(d8 + off, ring->buf[k], n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ring->ev[k], stream)) != cudaSuccess) return e;
        ring->used[k] = true;
    }
    return cudaSuccess;
}


} // namespace

struct cvr_handle {
    int device = 0;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int32_t n_chunks = 0;
    int variant = 0; // sweep geometry picked for this matrix (cvr_pick_sweep_variant)
    int64_t record_ints = 0;
    int64_t n_records = 0;
    // device arrays (the CVR structure)
    double* vals = nullptr;
    int32_t* cols = nullptr;
    int32_t* record = nullptr;
    CvrChunk* chunks = nullptr;
    CvrRowLists rows; // accumulated / never-written rows (what needs clearing before a sweep)
    // device vectors for the host-facing call
    double* x = nullptr;
    double* y = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double convert_seconds = 0.0, create_seconds = 0.0;
    double convert_kernel_seconds = 0.0, row_lists_seconds = 0.0;
    cudaEvent_t ev2 = nullptr;
    int64_t launches = 0;
    int64_t device_bytes = 0;
    std::vector<CvrChunk> host_chunks; // copy of the descriptors (export / info)
    // optional per-launch timing of the SpMV kernel alone (cvr_set_kernel_timing)
    unsigned int* done_counter = nullptr; // [2], [3]: next-chunk / warps-finished counters of the sweep's dynamic chunk queue;
                                          // [0] last-block detection of the publish epilogue,
                                          // [1] epoch of the first peer barrier that timed out (0 = none)
    // CUDA graph of GRAPH_UNROLL iterations of the host-facing loop (cvr_spmv with iters >> 1): captured once
    cudaStream_t copy_stream = nullptr;          // D2H of finished row slabs behind the sweep (cvr_spmv, iters = 1)
    std::vector<cudaEvent_t> slab_events;
    cudaGraphExec_t loop_graph = nullptr;
    int loop_graph_variant = -1; // sweep geometry the graph was captured with (CVR_SPMV_KERNEL can change it)
    int64_t loop_graph_kernels = 0;
    bool timing = false;
    std::vector<cudaEvent_t> timing_events; // begin/end pairs, `timing_used` of them recorded
    size_t timing_used = 0;

    ~cvr_handle()
    {
        cudaSetDevice(device);
        if (loop_graph) cudaGraphExecDestroy(loop_graph);
        for (cudaEvent_t e : slab_events) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        // the arrays may still be in use by sweeps the caller queued on its own streams: one device-wide
        // synchronisation, then stream-ordered frees back into the pool (no further blocking)
        cudaDeviceSynchronize();
        cvr_dev_free(vals, stream);
        cvr_dev_free(cols, stream);
        cvr_dev_free(record, stream);
        cvr_dev_free(chunks, stream);
        cvr_dev_free(rows.boundary, stream);
        cvr_dev_free(rows.empty, stream);
        cvr_dev_free(done_counter, stream);
        cvr_dev_free(x, stream);
        cvr_dev_free(y, stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        for (cudaEvent_t e : timing_events) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <typename T>
cudaError_t dev_alloc(cvr_handle* h, T** p, size_t count)
{
    const size_t bytes = sizeof(T) * (count ? count : 1);
    cudaError_t e = cvr_dev_malloc(reinterpret_cast<void**>(p), bytes, h->stream);
    if (e == cudaSuccess) h->device_bytes += (int64_t)bytes;
    return e;
}

// delimiter read that works for both widths (HOST pointers only)
int64_t host_delim(const cvr_csr_t* csr, int64_t k)
{
    return csr->row_delim32 ? (int64_t)csr->row_delim32[k] : csr->row_delim64[k];
}

// Host CSR only: delimiters start at 0, never decrease, and end at nnz -- or at nnz-1, the reference
// reader's off-by-one (spmv.cpp:522-526), which create_common repairs on the device copy.
int check_host_delimiters(const cvr_csr_t* csr, bool* ref_last_delim)
{
    *ref_last_delim = false;
    if (host_delim(csr, 0) != 0) return fail(CVR_ERR_INVALID, "row_delim[0] must be 0");
    int64_t prev = 0;
    for (int64_t k = 1; k <= csr->n_rows + 1; k++) {
        const int64_t v = host_delim(csr, k);
        if (v < prev) return fail(CVR_ERR_INVALID, "row_delim decreases at index %lld", (long long)k);
        prev = v;
    }
    if (prev == csr->nnz) return CVR_OK;
    if (prev == csr->nnz - 1) {
        *ref_last_delim = true;
        return CVR_OK;
    }
    return fail(CVR_ERR_INVALID, "row_delim[n_rows+1] = %lld must be nnz = %lld (or nnz-1, the reference reader's "
                "trailing delimiter)", (long long)prev, (long long)csr->nnz);
}

int check_csr(const cvr_csr_t* csr, int32_t n_chunks)
{
    if (!csr) return fail(CVR_ERR_INVALID, "csr is NULL");
    if (!csr->val || !csr->col) return fail(CVR_ERR_INVALID, "csr val/col is NULL");
    if ((csr->row_delim32 == nullptr) == (csr->row_delim64 == nullptr))
        return fail(CVR_ERR_INVALID, "exactly one of row_delim32 / row_delim64 must be set");
    if (csr->n_rows < 1 || csr->n_cols < 1)
        return fail(CVR_ERR_INVALID, "empty matrix (%lld x %lld)", (long long)csr->n_rows,
                    (long long)csr->n_cols);
    if (csr->n_rows > 0x7ffffff0LL || csr->n_cols > 0x7ffffff0LL)
        return fail(CVR_ERR_RANGE, "row/column ids must fit int32");
    if (csr->nnz < 16 || csr->nnz % 16 != 0)
        return fail(CVR_ERR_INVALID, "nnz = %lld must be a positive multiple of 16 (spmv.cpp:457)",
                    (long long)csr->nnz);
    if (csr->row_delim32 && csr->nnz > 0x7fffffffLL)
        return fail(CVR_ERR_RANGE, "nnz >= 2^31 needs row_delim64");
    if (n_chunks < 0 || (int64_t)n_chunks > csr->nnz / 16)
        return fail(CVR_ERR_INVALID, "n_chunks = %d must be in [0, nnz/16 = %lld]", n_chunks,
                    (long long)(csr->nnz / 16));
    return CVR_OK;
}

// Conversion proper; `csr` holds DEVICE pointers.
int convert_on_device(cvr_handle* h, const cvr_csr_t* csr)
{
    const int32_t T = h->n_chunks;
    PhaseTrace trace;
    CUDA_TRY(dev_alloc(h, &h->vals, (size_t)h->nnz));
    CUDA_TRY(dev_alloc(h, &h->cols, (size_t)h->nnz));
    CUDA_TRY(dev_alloc(h, &h->record, (size_t)h->record_ints));
    CUDA_TRY(dev_alloc(h, &h->chunks, (size_t)T));
    CUDA_TRY(dev_alloc(h, &h->x, (size_t)h->n_cols + 1));
    CUDA_TRY(dev_alloc(h, &h->y, (size_t)h->n_rows + 1));
    // counters of the multi-GPU epilogue: allocated here, not lazily inside the iteration loop (an allocation
    // or memset issued while a peer already spins at the flag barrier may have to wait for it)
    CUDA_TRY(dev_alloc(h, &h->done_counter, 4));
    CUDA_TRY(cudaMemsetAsync(h->done_counter, 0, 4 * sizeof(unsigned int), h->stream));

    int2* segments = nullptr;
    int32_t* seg_count = nullptr;
    const size_t seg_entries = (size_t)CVR_SEG_STRIDE * T + (size_t)h->n_rows + 64;
    CUDA_TRY(cvr_dev_malloc(reinterpret_cast<void**>(&segments), sizeof(int2) * seg_entries, h->stream));
    cudaError_t e = cvr_dev_malloc(reinterpret_cast<void**>(&seg_count), sizeof(int32_t) * ((size_t)T + 2), h->stream);
    if (e != cudaSuccess) {
        cvr_dev_free(segments, h->stream);
        return fail(CVR_ERR_CUDA, "cudaMalloc(seg_count): %s", cudaGetErrorString(e));
    }

    uint32_t* row_bitmap = nullptr;
    const size_t bitmap_words = (size_t)(h->n_rows + 2) / 32 + 2;
    e = cvr_dev_malloc(reinterpret_cast<void**>(&row_bitmap), sizeof(uint32_t) * bitmap_words, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(row_bitmap, 0, sizeof(uint32_t) * bitmap_words, h->stream);
    if (e != cudaSuccess) {
        cvr_dev_free(segments, h->stream);
        cvr_dev_free(seg_count, h->stream);
        cvr_dev_free(row_bitmap, h->stream);
        return fail(CVR_ERR_CUDA, "cudaMalloc(row_bitmap): %s", cudaGetErrorString(e));
    }

    trace.mark("device allocations (12)");
    CvrConvertArgs a{};
    a.row_bitmap = row_bitmap;
    a.csr_val = csr->val;
    a.csr_col = csr->col;
    a.rd32 = csr->row_delim32;
    a.rd64 = csr->row_delim64;
    a.nnz = h->nnz;
    a.n_rows = h->n_rows;
    a.n_chunks = T;
    a.cvr_vals = h->vals;
    a.cvr_cols = h->cols;
    a.record = h->record;
    a.chunks = h->chunks;
    a.segments = segments;
    a.seg_count = seg_count;

    int rc = CVR_OK;
    float ms = 0.f, ms_kernels = 0.f;
    double t_lists = 0.0;
    do {
        // every int the scheduler does not write reads back as -1 (a terminator)
        if ((e = cudaMemsetAsync(h->record, 0xff, sizeof(int32_t) * (size_t)h->record_ints,
                                 h->stream)) != cudaSuccess) break;
        if ((e = cudaEventRecord(h->ev0, h->stream)) != cudaSuccess) break;
        const int launched = cvr_launch_convert(a, h->stream);
        if (launched < 0) {
            e = cudaGetLastError();
            rc = fail(CVR_ERR_CUDA, "conversion kernel launch failed: %s", cudaGetErrorString(e));
            break;
        }
        h->launches += launched;
        // ev0..ev2 brackets the two conversion kernels (schedule + permute) alone; the row lists that follow
        // allocate, synchronise and free around their three small kernels, so they are timed by the host clock
        if ((e = cudaEventRecord(h->ev2, h->stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) break;
        trace.mark("record fill + conversion kernels");
        const double tl0 = wall_seconds();
        const int listed = cvr_build_row_lists(h->chunks, T, csr->row_delim32, csr->row_delim64, h->n_rows,
                                               &h->rows, h->stream);
        if (listed < 0) {
            rc = fail(CVR_ERR_CUDA, "building the row lists failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        h->launches += listed;
        h->device_bytes += 4 * ((int64_t)h->rows.n_boundary + h->rows.n_empty + 2);
        if ((e = cudaEventRecord(h->ev1, h->stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) break;
        t_lists = wall_seconds() - tl0;
        trace.mark("row lists");
        if ((e = cudaEventElapsedTime(&ms, h->ev0, h->ev1)) != cudaSuccess) break;
        if ((e = cudaEventElapsedTime(&ms_kernels, h->ev0, h->ev2)) != cudaSuccess) break;
    } while (0);
    cvr_dev_free(segments, h->stream);
    cvr_dev_free(seg_count, h->stream);
    cvr_dev_free(row_bitmap, h->stream);
    trace.mark("free scratch");
    if (rc != CVR_OK) return rc;
    if (e != cudaSuccess) return fail(CVR_ERR_CUDA, "conversion failed: %s", cudaGetErrorString(e));
    h->convert_kernel_seconds = ms_kernels * 1e-3;
    h->row_lists_seconds = t_lists;
    h->convert_seconds = ms * 1e-3;

    h->host_chunks.resize((size_t)T);
    CUDA_TRY(cudaMemcpy(h->host_chunks.data(), h->chunks, sizeof(CvrChunk) * (size_t)T,
                        cudaMemcpyDeviceToHost));
    h->n_records = 0;
    for (const CvrChunk& c : h->host_chunks) h->n_records += c.n_rec + CVR_W;
    trace.mark("chunk descriptors to host");
    return CVR_OK;
}

int auto_chunks_for(int64_t nnz, int device, int variant, int32_t* n_chunks)
{
    if (!n_chunks || nnz < 16) return fail(CVR_ERR_INVALID, "bad arguments to cvr_auto_chunks");
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
        return fail(CVR_ERR_CUDA, "cannot query device %d: no CPU fallback", device);
    // One warp per chunk, chunks are nnz-balanced: size the count in whole WAVES of resident
    // warps (SMs x warps the SpMV kernel keeps resident per SM) so the last wave is full, and
    // aim at ~4K elements (48 KB of stream) per chunk: the per-chunk prologue (descriptor -> records ->
    // first bulk copy, a chain of dependent loads) costs ~2 us of a warp, measured as FEM 69.7 us at
    // 14208 chunks vs 64.1-64.8 us at 3552-7104; skewed matrices want more, smaller chunks for balance
    // (R-MAT-22: 322 us at 21312 chunks vs 346 us at 7104), 4K is the compromise.  CVR_CHUNK_NNZ overrides.
    int64_t target_nnz = 4096;
    if (const char* s = getenv("CVR_CHUNK_NNZ")) {
        const long long v = atoll(s);
        if (v >= 16) target_nnz = v;
    }
    if (cudaSetDevice(device) != cudaSuccess) return fail(CVR_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    const int64_t wave = (int64_t)sms * cvr_spmv_resident_warps_per_sm(variant);
    int64_t waves = (nnz / target_nnz + wave / 2) / wave;
    if (waves < 1) waves = 1;
    int64_t t = waves * wave;
    if (t > nnz / 16) t = nnz / 16;
    if (t > 0x3fffffff) t = 0x3fffffff;
    if (t < 1) t = 1;
    *n_chunks = (int32_t)t;
    return CVR_OK;
}

int create_common(const cvr_csr_t* csr, int32_t n_chunks, int device, bool csr_on_device,
                  cvr_handle_t** out)
{
    if (!out) return fail(CVR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_csr(csr, n_chunks);
    if (rc != CVR_OK) return rc;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CVR_ERR_CUDA, "no CUDA device: libcvr_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev)
        return fail(CVR_ERR_INVALID, "device %d out of range [0, %d)", device, n_dev);
    CUDA_TRY(cudaSetDevice(device));
    const int variant = cvr_pick_sweep_variant(csr->nnz, csr->n_rows);
    if (n_chunks == 0) {
        rc = auto_chunks_for(csr->nnz, device, variant, &n_chunks);
        if (rc != CVR_OK) return rc;
    }

    {   // CUDA loads a kernel's module lazily at its first launch: do that outside the timed creation
        static thread_local int preloaded_device = -1;
        if (preloaded_device != device) {
            cvr_preload_convert_kernels();
            cvr_preload_spmv_kernels();
            cvr_pool_setup(device);
            cudaGetLastError();
            preloaded_device = device;
        }
        if (!csr_on_device) staging_ring(); // pinned once per process, outside the timed creation
    }
    const double t0 = wall_seconds();
    PhaseTrace trace;
    cvr_handle* h = new (std::nothrow) cvr_handle();
    if (!h) return fail(CVR_ERR_INVALID, "out of host memory");
    h->device = device;
    h->n_rows = csr->n_rows;
    h->n_cols = csr->n_cols;
    h->nnz = csr->nnz;
    h->n_chunks = n_chunks;
    h->variant = variant;
    h->record_ints = cvr_record_ints(csr->n_rows, n_chunks);

    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev1)) != cudaSuccess || (e = cudaEventCreate(&h->ev2)) != cudaSuccess) {
        delete h;
        return fail(CVR_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    }

    trace.mark("handle, stream, events");
    cvr_csr_t dev = *csr;
    double* d_val = nullptr;
    int32_t* d_col = nullptr;
    void* d_rd = nullptr;
    bool ref_last_delim = false;
    if (!csr_on_device) {
        rc = check_host_delimiters(csr, &ref_last_delim);
        if (rc != CVR_OK) {
            delete h;
            return rc;
        }
    } else {
        // device CSR: only the last delimiter is inspected (one 4/8-byte copy)
        int64_t last = 0;
        if (csr->row_delim64) {
            e = cudaMemcpy(&last, csr->row_delim64 + csr->n_rows + 1, 8, cudaMemcpyDeviceToHost);
        } else {
            int32_t l32 = 0;
            e = cudaMemcpy(&l32, csr->row_delim32 + csr->n_rows + 1, 4, cudaMemcpyDeviceToHost);
            last = l32;
        }
        if (e != cudaSuccess) {
            delete h;
            return fail(CVR_ERR_CUDA, "cannot read row_delim[n_rows+1] from the device: %s", cudaGetErrorString(e));
        }
        if (last == csr->nnz - 1) ref_last_delim = true;
        else if (last != csr->nnz) {
            delete h;
            return fail(CVR_ERR_INVALID, "row_delim[n_rows+1] = %lld must be nnz = %lld (or nnz-1)", (long long)last,
                        (long long)csr->nnz);
        }
    }
    if (!csr_on_device) {
        const size_t rd_bytes = (size_t)(csr->n_rows + 2) * (csr->row_delim64 ? 8 : 4);
        if ((e = cvr_dev_malloc(reinterpret_cast<void**>(&d_val), sizeof(double) * (size_t)csr->nnz, h->stream)) == cudaSuccess &&
            (e = cvr_dev_malloc(reinterpret_cast<void**>(&d_col), sizeof(int32_t) * (size_t)csr->nnz, h->stream)) == cudaSuccess &&
            (e = cvr_dev_malloc(&d_rd, rd_bytes, h->stream)) == cudaSuccess &&
            (e = upload_async(d_val, csr->val, sizeof(double) * (size_t)csr->nnz, h->stream)) == cudaSuccess &&
            (e = upload_async(d_col, csr->col, sizeof(int32_t) * (size_t)csr->nnz, h->stream)) == cudaSuccess &&
            (e = upload_async(d_rd, csr->row_delim64 ? (const void*)csr->row_delim64 : (const void*)csr->row_delim32,
                              rd_bytes, h->stream)) == cudaSuccess)
            e = cudaStreamSynchronize(h->stream); // the caller's arrays are free again when cvr_create returns
        if (e != cudaSuccess) {
            cvr_dev_free(d_val, h->stream);
            cvr_dev_free(d_col, h->stream);
            cvr_dev_free(d_rd, h->stream);
            delete h;
            return fail(CVR_ERR_CUDA, "CSR upload failed: %s", cudaGetErrorString(e));
        }
        dev.val = d_val;
        dev.col = d_col;
        dev.row_delim32 = csr->row_delim64 ? nullptr : static_cast<const int32_t*>(d_rd);
        dev.row_delim64 = csr->row_delim64 ? static_cast<const int64_t*>(d_rd) : nullptr;
    } else if (ref_last_delim) {
        // the caller's device CSR is read-only: repair a private copy of the delimiters
        const size_t rd_bytes = (size_t)(csr->n_rows + 2) * (csr->row_delim64 ? 8 : 4);
        if ((e = cvr_dev_malloc(&d_rd, rd_bytes, h->stream)) == cudaSuccess)
            e = cudaMemcpyAsync(d_rd, csr->row_delim64 ? (const void*)csr->row_delim64 : (const void*)csr->row_delim32,
                                rd_bytes, cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) {
            cvr_dev_free(d_rd, h->stream);
            delete h;
            return fail(CVR_ERR_CUDA, "delimiter copy failed: %s", cudaGetErrorString(e));
        }
        dev.row_delim32 = csr->row_delim64 ? nullptr : static_cast<const int32_t*>(d_rd);
        dev.row_delim64 = csr->row_delim64 ? static_cast<const int64_t*>(d_rd) : nullptr;
    }
    if (ref_last_delim) {
        // trailing delimiters nnz-1 -> nnz (see cvr_launch_fix_last_delim): the converter then never looks
        // for a row past the end, and no tail row n_rows+1 can reach the sweep
        if (cvr_launch_fix_last_delim(csr->row_delim64 ? nullptr : static_cast<int32_t*>(d_rd),
                                      csr->row_delim64 ? static_cast<int64_t*>(d_rd) : nullptr, csr->n_rows, csr->nnz,
                                      h->stream) < 0 ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            cvr_dev_free(d_val, h->stream);
            cvr_dev_free(d_col, h->stream);
            cvr_dev_free(d_rd, h->stream);
            delete h;
            return fail(CVR_ERR_CUDA, "delimiter repair failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }

    trace.mark("delimiter check + CSR upload");
    rc = convert_on_device(h, &dev);
    trace.mark("convert_on_device (total)");
    cvr_dev_free(d_val, h->stream);
    cvr_dev_free(d_col, h->stream);
    cvr_dev_free(d_rd, h->stream);
    trace.mark("free device CSR copy");
    if (rc != CVR_OK) {
        delete h;
        return rc;
    }
    h->create_seconds = wall_seconds() - t0;
    *out = h;
    return CVR_OK;
}

} // namespace

// ---- device memory of the library: cudaMalloc underneath, freed blocks kept for the next matrix.
// (The CUDA stream-ordered pool was tried first: warm it is as fast, but GROWING it costs 100-190 ms per GB -- road,
// first matrix of a process: 186 ms against 27 ms with cudaMalloc; R-MAT-24: 60-600 ms against 11 ms --
// profiles/r02_create_trace.txt.)  Contract: a block is freed only after the work that used it has completed; every
// call site synchronises its stream (or the device) first, so a cached block can be handed out again at once.
namespace {
struct CachedBlock {
    void* ptr;
    size_t bytes;
    int device;
};
std::mutex g_mem_mutex;
std::vector<CachedBlock> g_live;   // blocks handed out (ptr -> size, device)
std::vector<CachedBlock> g_cached; // freed blocks kept for reuse
size_t g_cached_bytes = 0;
size_t g_keep_bytes = (size_t)2 << 30;                 // cap of the cache (CVR_POOL_KEEP_MB)
constexpr size_t CACHE_MAX_BLOCK = (size_t)768 << 20;  // larger blocks always go back to the driver
} // namespace

cudaError_t cvr_dev_malloc(void** p, size_t bytes, cudaStream_t)
{
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 255) & ~(size_t)255;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pool_enabled()) {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        size_t best = g_cached.size();
        for (size_t k = 0; k < g_cached.size(); k++) {
            const CachedBlock& c = g_cached[k];
            if (c.device == dev && c.bytes >= bytes && c.bytes <= bytes + bytes / 4 + 4096 &&
                (best == g_cached.size() || c.bytes < g_cached[best].bytes))
                best = k;
        }
        if (best < g_cached.size()) {
            const CachedBlock c = g_cached[best];
            g_cached[best] = g_cached.back();
            g_cached.pop_back();
            g_cached_bytes -= c.bytes;
            g_live.push_back(c);
            *p = c.ptr;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && pool_enabled()) { // out of memory: give the cache back and retry once
        cudaGetLastError();
        std::vector<CachedBlock> drop;
        {
            std::lock_guard<std::mutex> lock(g_mem_mutex);
            for (size_t k = 0; k < g_cached.size();)
                if (g_cached[k].device == dev) {
                    drop.push_back(g_cached[k]);
                    g_cached_bytes -= g_cached[k].bytes;
                    g_cached[k] = g_cached.back();
                    g_cached.pop_back();
                } else k++;
        }
        for (const CachedBlock& c : drop) cudaFree(c.ptr);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess && pool_enabled()) {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        g_live.push_back(CachedBlock{*p, bytes, dev});
    }
    return e;
}

void cvr_dev_free(void* p, cudaStream_t)
{
    if (!p) return;
    if (pool_enabled()) {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        for (size_t k = 0; k < g_live.size(); k++)
            if (g_live[k].ptr == p) {
                const CachedBlock c = g_live[k];
                g_live[k] = g_live.back();
                g_live.pop_back();
                if (c.bytes <= CACHE_MAX_BLOCK && g_cached_bytes + c.bytes <= g_keep_bytes) {
                    g_cached.push_back(c);
                    g_cached_bytes += c.bytes;
                    return;
                }
                break;
            }
    }
    cudaFree(p);
}

void cvr_pool_setup(int)
{
    if (const char* e = getenv("CVR_POOL_KEEP_MB")) {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        g_keep_bytes = (size_t)atoll(e) << 20;
    }
}

extern "C" {

int cvr_abi_version(void) { return CVR_B200_ABI_VERSION; }

const char* cvr_last_error(void) { return g_cvr_err; }

int cvr_device_init(int device)
{
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CVR_ERR_CUDA, "no CUDA device: libcvr_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev)
        return fail(CVR_ERR_INVALID, "device %d out of range [0, %d)", device, n_dev);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(nullptr));
    cvr_preload_convert_kernels();
    cvr_preload_spmv_kernels();
    cvr_pool_setup(device);
    staging_ring();
    cudaGetLastError();
    return CVR_OK;
}

int64_t cvr_record_ints(int64_t n_rows, int32_t n_chunks)
{
    return 2 * (n_rows + 240 + 32 * (int64_t)n_chunks);
}

int cvr_auto_chunks(int64_t nnz, int device, int32_t* n_chunks)
{
    // without the row count the short-row geometry is assumed (cvr_create knows the rows and may differ)
    return auto_chunks_for(nnz, device, cvr_pick_sweep_variant(nnz, nnz), n_chunks);
}

int cvr_create(const cvr_csr_t* csr_host, int32_t n_chunks, int device, cvr_handle_t** out)
{
    return create_common(csr_host, n_chunks, device, false, out);
}

int cvr_create_from_device(const cvr_csr_t* csr_dev, int32_t n_chunks, int device,
                           cvr_handle_t** out)
{
    return create_common(csr_dev, n_chunks, device, true, out);
}

static int spmv_device_impl(cvr_handle_t* h, const double* x_dev, double* y_dev, const CvrPublish* pub,
                            void* cuda_stream, const CvrBarrier* bar = nullptr, bool y_is_clear = false);

// the sweep's dynamic chunk queue (two counters behind done_counter); CVR_DYNAMIC_CHUNKS=0 keeps the static
// round robin.  Read per call so that tools/kernel_ab.py can switch it at run time.
static unsigned int* chunk_queue(cvr_handle_t* h)
{
    const char* e = getenv("CVR_DYNAMIC_CHUNKS");
    if (e && *e == '0') return nullptr;
    return h->done_counter ? h->done_counter + 2 : nullptr;
}

int cvr_spmv_device(cvr_handle_t* h, const double* x_dev, double* y_dev, void* cuda_stream)
{
    return spmv_device_impl(h, x_dev, y_dev, nullptr, cuda_stream);
}

int cvr_spmv_publish(cvr_handle_t* h, const double* x_dev, double* y_dev, const cvr_publish_t* pub,
                     void* const* flag_arrays, int32_t rank, int32_t n_ranks, uint32_t epoch,
                     int32_t y_is_clear, void* cuda_stream)
{
    if (!h || !pub) return fail(CVR_ERR_INVALID, "NULL handle / publish descriptor");
    if (!flag_arrays || n_ranks < 1 || n_ranks > CVR_MAX_PEERS || rank < 0 || rank >= n_ranks)
        return fail(CVR_ERR_INVALID, "bad barrier arguments to cvr_spmv_publish");
    CvrBarrier b{};
    for (int k = 0; k < n_ranks; k++) {
        if (!flag_arrays[k]) return fail(CVR_ERR_INVALID, "flag_arrays[%d] is NULL", k);
        b.flags[k] = static_cast<uint32_t*>(flag_arrays[k]);
    }
    b.rank = rank;
    b.n_ranks = n_ranks;
    b.epoch = epoch;
    b.timeout_cycles = barrier_timeout_cycles(h->device);
    if (!h->done_counter) {
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cvr_dev_malloc(reinterpret_cast<void**>(&h->done_counter), 4 * sizeof(unsigned int), nullptr));
        CUDA_TRY(cudaMemset(h->done_counter, 0, 4 * sizeof(unsigned int)));
    }
    b.error = h->done_counter + 1;
    if (pub->n_dst < 1 || pub->n_dst > CVR_MAX_PEERS)
        return fail(CVR_ERR_INVALID, "n_dst = %d must be in [1, %d]", pub->n_dst, CVR_MAX_PEERS);
    if ((pub->mode & 4) && (pub->self < 0 || pub->self >= pub->n_dst))
        return fail(CVR_ERR_INVALID, "mode bit 2 needs self = %d in [0, n_dst)", pub->self);
    CvrPublish p{};
    p.n_dst = pub->n_dst;
    p.self = pub->self;
    p.mode = pub->mode;
    p.row_offset = pub->row_offset;
    p.needs = pub->needs;
    p.clear_next = pub->clear_next;
    p.chunk_any = pub->chunk_any;
    p.mc = pub->multicast;
    if (p.mc && ((pub->mode & 4) || pub->needs || pub->chunk_any))
        return fail(CVR_ERR_INVALID, "multicast publishing needs mode bit 2 clear and needs = chunk_any = NULL");
    for (int k = 0; k < pub->n_dst; k++) {
        if (!pub->dst[k]) return fail(CVR_ERR_INVALID, "dst[%d] is NULL", k);
        p.dst[k] = pub->dst[k];
    }
    return spmv_device_impl(h, x_dev, y_dev, &p, cuda_stream, &b, y_is_clear != 0);
}

static int spmv_device_impl(cvr_handle_t* h, const double* x_dev, double* y_dev, const CvrPublish* pub,
                            void* cuda_stream, const CvrBarrier* bar, bool y_is_clear)
{
    if (!h || !x_dev || !y_dev) return fail(CVR_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaEvent_t eb = nullptr, ee = nullptr;
    if (h->timing) {
        if (h->timing_used + 2 > h->timing_events.size()) {
            for (int k = 0; k < 2; k++) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e));
                h->timing_events.push_back(e);
            }
        }
        eb = h->timing_events[h->timing_used];
        ee = h->timing_events[h->timing_used + 1];
        h->timing_used += 2;
    }
    const int launched = cvr_launch_spmv(h->variant, h->chunks, h->n_chunks, h->vals, h->cols, h->record,
                                         x_dev, y_dev, h->n_rows, h->rows, pub,
                                         static_cast<cudaStream_t>(cuda_stream), eb, ee, bar,
                                         h->done_counter, y_is_clear, 0, -1, chunk_queue(h));
    if (launched < 0)
        return fail(CVR_ERR_CUDA, "SpMV launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += launched;
    return CVR_OK;
}

int cvr_spmv(cvr_handle_t* h, const double* x_host, double* y_host, int32_t iters,
             double* seconds_per_iter)
{
    if (!h || !x_host || !y_host) return fail(CVR_ERR_INVALID, "NULL argument");
    if (iters < 1) return fail(CVR_ERR_INVALID, "iters = %d must be >= 1", iters);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->x, x_host, sizeof(double) * (size_t)(h->n_cols + 1),
                             cudaMemcpyHostToDevice, h->stream));
    // ---- one SpMV with host vectors: pipeline the sweep and the copy of y.  The chunks are swept in SLABS
    // (one launch each, same kernel); when a slab is done every row before the first row of the next slab is
    // final, so its part of y goes back over PCIe on a second stream while the next slab is swept.  On
    // R-MAT-24 (134 MB of y, 2.4 ms over PCIe against a 1.4 ms sweep) that hides the sweep behind the copy.
    {
        constexpr int32_t SLABS = 8;
        const char* ns = getenv("CVR_NO_SLABS");
        const size_t y_bytes = sizeof(double) * (size_t)(h->n_rows + 1);
        size_t min_bytes = (size_t)8 << 20; // below this the extra launches cost more than the overlap gains
        if (const char* mb = getenv("CVR_SLAB_MIN_BYTES")) min_bytes = (size_t)atoll(mb);
        if (iters == 1 && !h->timing && !(ns && *ns == '1') && y_bytes >= min_bytes &&
            h->n_chunks >= 64 * SLABS) {
            if (!h->copy_stream) {
                CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
                for (int k = 0; k < SLABS; k++) {
                    cudaEvent_t e;
                    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    h->slab_events.push_back(e);
                }
            }
            CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
            int64_t copied_upto = 0; // rows [0, copied_upto) are on their way to the host
            for (int32_t k = 0; k < SLABS; k++) {
                const int32_t c0 = (int32_t)((int64_t)h->n_chunks * k / SLABS);
                const int32_t c1 = (int32_t)((int64_t)h->n_chunks * (k + 1) / SLABS);
                const int launched = cvr_launch_spmv(h->variant, h->chunks, h->n_chunks, h->vals, h->cols, h->record, h->x,
                                                     h->y, h->n_rows, h->rows, nullptr, h->stream, nullptr, nullptr,
                                                     nullptr, nullptr, /*y_is_clear=*/k > 0, c0, c1, chunk_queue(h));
                if (launched < 0)
                    return fail(CVR_ERR_CUDA, "SpMV launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                h->launches += launched;
                CUDA_TRY(cudaEventRecord(h->slab_events[(size_t)k], h->stream));
                // the first row of the next slab may still be accumulated by it: stop one row short
                const int64_t final_upto = k + 1 < SLABS ? (int64_t)h->host_chunks[(size_t)c1].first_row : h->n_rows + 1;
                if (final_upto > copied_upto) {
                    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->slab_events[(size_t)k], 0));
                    CUDA_TRY(cudaMemcpyAsync(y_host + copied_upto, h->y + copied_upto,
                                             sizeof(double) * (size_t)(final_upto - copied_upto), cudaMemcpyDeviceToHost,
                                             h->copy_stream));
                    copied_upto = final_upto;
                }
            }
            CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
            if (seconds_per_iter) {
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
                *seconds_per_iter = (double)ms * 1e-3;
            }
            return CVR_OK;
        }
    }
    // The iteration loop (the reference's spmv.cpp:1024-1034).  From GRAPH_UNROLL iterations on it is replayed
    // from a CUDA graph captured once per handle: GRAPH_UNROLL x (clearing kernel -> sweep, programmatic edge
    // kept) per graph launch instead of two launches per iteration.  CVR_NO_GRAPH=1 keeps the plain loop.
    constexpr int32_t GRAPH_UNROLL = 20;
    int32_t done = 0;
    const char* ng = getenv("CVR_NO_GRAPH");
    if (iters >= GRAPH_UNROLL && !h->timing && !(ng && *ng == '1')) {
        const int variant_now = cvr_pick_sweep_variant(h->nnz, h->n_rows);
        if (h->loop_graph && h->loop_graph_variant != variant_now) {
            cudaGraphExecDestroy(h->loop_graph);
            h->loop_graph = nullptr;
        }
        if (!h->loop_graph) {
            cudaGraph_t graph = nullptr;
            const int64_t l0 = h->launches;
            cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            int rc = CVR_OK;
            if (ce == cudaSuccess) {
                for (int32_t it = 0; it < GRAPH_UNROLL && rc == CVR_OK; it++)
                    rc = cvr_spmv_device(h, h->x, h->y, h->stream);
                ce = cudaStreamEndCapture(h->stream, &graph);
            }
            h->loop_graph_kernels = h->launches - l0;
            h->launches = l0; // nothing ran yet
            if (rc == CVR_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&h->loop_graph, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != CVR_OK || ce != cudaSuccess) {
                cudaGetLastError(); // capture not possible here: fall back to the plain loop
                h->loop_graph = nullptr;
            }
            h->loop_graph_variant = variant_now;
        }
    }
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    if (h->loop_graph) {
        for (; done + GRAPH_UNROLL <= iters; done += GRAPH_UNROLL) {
            CUDA_TRY(cudaGraphLaunch(h->loop_graph, h->stream));
            h->launches += h->loop_graph_kernels;
        }
    }
    for (int32_t it = done; it < iters; it++) {
        const int rc = cvr_spmv_device(h, h->x, h->y, h->stream);
        if (rc != CVR_OK) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaMemcpyAsync(y_host, h->y, sizeof(double) * (size_t)(h->n_rows + 1),
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (seconds_per_iter) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *seconds_per_iter = (double)ms * 1e-3 / iters;
    }
    return CVR_OK;
}

int cvr_export(cvr_handle_t* h, cvr_arrays_t* o)
{
    if (!h || !o) return fail(CVR_ERR_INVALID, "NULL argument");
    if (h->nnz > 0x7fffffffLL && o->nnz_rows)
        return fail(CVR_ERR_RANGE, "nnz_rows cannot hold offsets >= 2^31");
    CUDA_TRY(cudaSetDevice(h->device));
    if (o->vals)
        CUDA_TRY(cudaMemcpy(o->vals, h->vals, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost));
    if (o->cols)
        CUDA_TRY(cudaMemcpy(o->cols, h->cols, sizeof(int32_t) * (size_t)h->nnz, cudaMemcpyDeviceToHost));
    if (o->record)
        CUDA_TRY(cudaMemcpy(o->record, h->record, sizeof(int32_t) * (size_t)h->record_ints,
                            cudaMemcpyDeviceToHost));
    for (int32_t t = 0; t < h->n_chunks; t++) {
        const CvrChunk& c = h->host_chunks[(size_t)t];
        if (o->nnz_rows) {
            o->nnz_rows[4 * t + 0] = (int32_t)c.start;
            o->nnz_rows[4 * t + 1] = (int32_t)(c.start + c.len);
            o->nnz_rows[4 * t + 2] = c.first_row;
            o->nnz_rows[4 * t + 3] = c.last_row;
        }
        if (o->split) {
            o->split[2 * t + 0] = c.split0;
            o->split[2 * t + 1] = c.split1;
        }
        if (o->final_2)
            for (int q = 0; q < CVR_W; q++) o->final_2[16 * t + q] = c.tail[q];
    }
    return CVR_OK;
}

int cvr_get_info(cvr_handle_t* h, cvr_info_t* info)
{
    if (!h || !info) return fail(CVR_ERR_INVALID, "NULL argument");
    info->n_rows = h->n_rows;
    info->n_cols = h->n_cols;
    info->nnz = h->nnz;
    info->n_chunks = h->n_chunks;
    info->device = h->device;
    info->n_records = h->n_records;
    info->record_ints = h->record_ints;
    info->algorithmic_bytes = 12 * h->nnz + 8 * h->n_records + 56 * (int64_t)h->n_chunks +
                              8 * (h->n_cols + 1) + 8 * (h->n_rows + 1);
    info->convert_seconds = h->convert_seconds;
    info->create_seconds = h->create_seconds;
    info->kernel_launches = h->launches;
    info->device_bytes = h->device_bytes;
    info->convert_kernel_seconds = h->convert_kernel_seconds;
    info->row_lists_seconds = h->row_lists_seconds;
    return CVR_OK;
}

int cvr_device_vectors(cvr_handle_t* h, double** x_dev, double** y_dev)
{
    if (!h) return fail(CVR_ERR_INVALID, "NULL handle");
    if (x_dev) *x_dev = h->x;
    if (y_dev) *y_dev = h->y;
    return CVR_OK;
}

const char* cvr_kernel_variant(cvr_handle_t* h) { return cvr_spmv_kernel_name(h ? h->variant : 0); }

int cvr_device_arrays(cvr_handle_t* h, const double** vals_dev, const int32_t** cols_dev,
                      const int32_t** record_dev)
{
    if (!h) return fail(CVR_ERR_INVALID, "NULL handle");
    if (vals_dev) *vals_dev = h->vals;
    if (cols_dev) *cols_dev = h->cols;
    if (record_dev) *record_dev = h->record;
    return CVR_OK;
}

int cvr_set_kernel_timing(cvr_handle_t* h, int enabled)
{
    if (!h) return fail(CVR_ERR_INVALID, "NULL handle");
    h->timing = enabled != 0;
    h->timing_used = 0;
    return CVR_OK;
}

int cvr_get_kernel_timing(cvr_handle_t* h, double* total_seconds, int64_t* launches)
{
    if (!h || !total_seconds || !launches) return fail(CVR_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    double total = 0.0;
    for (size_t k = 0; k + 1 < h->timing_used; k += 2) {
        CUDA_TRY(cudaEventSynchronize(h->timing_events[k + 1]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->timing_events[k], h->timing_events[k + 1]));
        total += (double)ms * 1e-3;
    }
    *total_seconds = total;
    *launches = (int64_t)(h->timing_used / 2);
    h->timing_used = 0;
    return CVR_OK;
}

int cvr_check_async_error(cvr_handle_t* h)
{
    if (!h) return fail(CVR_ERR_INVALID, "NULL handle");
    if (!h->done_counter) return CVR_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned int epoch = 0;
    CUDA_TRY(cudaMemcpy(&epoch, h->done_counter + 1, sizeof(epoch), cudaMemcpyDeviceToHost));
    if (epoch != 0) {
        CUDA_TRY(cudaMemset(h->done_counter + 1, 0, sizeof(unsigned int)));
        return fail(CVR_ERR_STATE,
                    "peer flag barrier timed out at epoch %u (a peer stalled; CVR_BARRIER_TIMEOUT_MS): the x "
                    "vectors of this and later iterations are not trustworthy", epoch);
    }
    return CVR_OK;
}

int cvr_column_footprint(cvr_handle_t* h, uint8_t* used_dev, void* cuda_stream)
{
    if (!h || !used_dev) return fail(CVR_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (cvr_launch_column_footprint(h->cols, h->nnz, used_dev, static_cast<cudaStream_t>(cuda_stream)) < 0)
        return fail(CVR_ERR_CUDA, "footprint launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += 1;
    return CVR_OK;
}

int cvr_chunk_needs(cvr_handle_t* h, uint8_t* needs_dev, uint8_t* chunk_any_dev, void* cuda_stream)
{
    if (!h || !needs_dev || !chunk_any_dev) return fail(CVR_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (cvr_launch_chunk_needs(h->chunks, h->n_chunks, h->rows, needs_dev, chunk_any_dev,
                               static_cast<cudaStream_t>(cuda_stream)) < 0)
        return fail(CVR_ERR_CUDA, "chunk-needs launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += 2;
    return CVR_OK;
}

int cvr_peer_alloc(int device, int64_t bytes, void** dev_ptr, unsigned char handle[64])
{
    if (!dev_ptr || !handle || bytes <= 0) return fail(CVR_ERR_INVALID, "bad arguments to cvr_peer_alloc");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(dev_ptr, (size_t)bytes));
    CUDA_TRY(cudaMemset(*dev_ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t hnd;
    CUDA_TRY(cudaIpcGetMemHandle(&hnd, *dev_ptr));
    memcpy(handle, &hnd, 64);
    return CVR_OK;
}

int cvr_peer_open(int device, const unsigned char handle[64], void** dev_ptr)
{
    if (!dev_ptr || !handle) return fail(CVR_ERR_INVALID, "bad arguments to cvr_peer_open");
    CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    return CVR_OK;
}

int cvr_peer_close(int device, void* dev_ptr)
{
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return CVR_OK;
}

int cvr_peer_free(int device, void* dev_ptr)
{
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(dev_ptr));
    return CVR_OK;
}

int cvr_peer_barrier(int device, void* const* flag_arrays, int32_t rank, int32_t n_ranks, uint32_t epoch,
                     void* cuda_stream)
{
    if (!flag_arrays || n_ranks < 1 || n_ranks > CVR_MAX_PEERS || rank < 0 || rank >= n_ranks)
        return fail(CVR_ERR_INVALID, "bad arguments to cvr_peer_barrier");
    CUDA_TRY(cudaSetDevice(device));
    CvrBarrier b{};
    for (int p = 0; p < n_ranks; p++) {
        if (!flag_arrays[p]) return fail(CVR_ERR_INVALID, "flag_arrays[%d] is NULL", p);
        b.flags[p] = static_cast<uint32_t*>(flag_arrays[p]);
    }
    b.rank = rank;
    b.n_ranks = n_ranks;
    b.epoch = epoch;
    b.error = nullptr;
    b.timeout_cycles = barrier_timeout_cycles(device);
    if (cvr_launch_peer_barrier(b, static_cast<cudaStream_t>(cuda_stream)) < 0)
        return fail(CVR_ERR_CUDA, "barrier launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CVR_OK;
}

// ---- serialisation of a converted matrix: the conversion is paid once
namespace {
const char CVR_FILE_MAGIC[8] = {'C', 'V', 'R', 'B', '2', '0', '0', 1};
struct CvrFileHeader {
    char magic[8];
    int64_t n_rows, n_cols, nnz, record_ints, n_records;
    int32_t n_chunks, n_boundary, n_empty, chunk_bytes;
};

int copy_out(FILE* f, const void* dev, size_t bytes, std::vector<char>& buf)
{
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t n = bytes - off < buf.size() ? bytes - off : buf.size();
        if (cudaMemcpy(buf.data(), static_cast<const char*>(dev) + off, n, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(CVR_ERR_CUDA, "device->host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (fwrite(buf.data(), 1, n, f) != n) return fail(CVR_ERR_INVALID, "short write");
    }
    return CVR_OK;
}

int copy_in(FILE* f, void* dev, size_t bytes, std::vector<char>& buf)
{
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t n = bytes - off < buf.size() ? bytes - off : buf.size();
        if (fread(buf.data(), 1, n, f) != n) return fail(CVR_ERR_INVALID, "truncated CVR file");
        if (cudaMemcpy(static_cast<char*>(dev) + off, buf.data(), n, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(CVR_ERR_CUDA, "host->device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return CVR_OK;
}
} // namespace

int cvr_save(cvr_handle_t* h, const char* path)
{
    if (!h || !path) return fail(CVR_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    // written next to the target and renamed when complete: a failed save never leaves a partial file there
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(CVR_ERR_INVALID, "cannot open %s for writing", tmp.c_str());
    CvrFileHeader hd{};
    memcpy(hd.magic, CVR_FILE_MAGIC, 8);
    hd.n_rows = h->n_rows;
    hd.n_cols = h->n_cols;
    hd.nnz = h->nnz;
    hd.record_ints = h->record_ints;
    hd.n_records = h->n_records;
    hd.n_chunks = h->n_chunks;
    hd.n_boundary = h->rows.n_boundary;
    hd.n_empty = h->rows.n_empty;
    hd.chunk_bytes = (int32_t)sizeof(CvrChunk);
    std::vector<char> buf((size_t)32 << 20);
    int rc = fwrite(&hd, sizeof(hd), 1, f) == 1 ? CVR_OK : fail(CVR_ERR_INVALID, "short write");
    if (rc == CVR_OK) rc = copy_out(f, h->vals, sizeof(double) * (size_t)h->nnz, buf);
    if (rc == CVR_OK) rc = copy_out(f, h->cols, sizeof(int32_t) * (size_t)h->nnz, buf);
    if (rc == CVR_OK) rc = copy_out(f, h->record, sizeof(int32_t) * (size_t)h->record_ints, buf);
    if (rc == CVR_OK) rc = copy_out(f, h->chunks, sizeof(CvrChunk) * (size_t)h->n_chunks, buf);
    if (rc == CVR_OK) rc = copy_out(f, h->rows.boundary, sizeof(int32_t) * (size_t)h->rows.n_boundary, buf);
    if (rc == CVR_OK) rc = copy_out(f, h->rows.empty, sizeof(int32_t) * (size_t)h->rows.n_empty, buf);
    if (fclose(f) != 0 && rc == CVR_OK) rc = fail(CVR_ERR_INVALID, "close failed on %s", tmp.c_str());
    if (rc == CVR_OK && rename(tmp.c_str(), path) != 0) rc = fail(CVR_ERR_INVALID, "cannot rename %s to %s", tmp.c_str(), path);
    if (rc != CVR_OK) remove(tmp.c_str());
    return rc;
}

int cvr_load(const char* path, int device, cvr_handle_t** out)
{
    if (!path || !out) return fail(CVR_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = cvr_device_init(device);
    if (rc != CVR_OK) return rc;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(CVR_ERR_INVALID, "cannot open %s", path);
    const double t0 = wall_seconds();
    CvrFileHeader hd{};
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, CVR_FILE_MAGIC, 8) != 0 ||
        hd.chunk_bytes != (int32_t)sizeof(CvrChunk)) {
        fclose(f);
        return fail(CVR_ERR_INVALID, "%s is not a CVR file of this library version", path);
    }
    // the header is not trusted: every count is range-checked and the file must have exactly the size they imply
    const bool sane = hd.n_rows >= 1 && hd.n_rows <= 0x7ffffff0LL && hd.n_cols >= 1 && hd.n_cols <= 0x7ffffff0LL &&
                      hd.nnz >= 16 && hd.nnz % 16 == 0 && hd.n_chunks >= 1 && (int64_t)hd.n_chunks <= hd.nnz / 16 &&
                      hd.record_ints == cvr_record_ints(hd.n_rows, hd.n_chunks) && hd.n_boundary >= 0 &&
                      hd.n_boundary <= 9 * (int64_t)hd.n_chunks && hd.n_empty >= 0 && hd.n_empty <= hd.n_rows + 1 &&
                      hd.n_records >= 8 * (int64_t)hd.n_chunks && 2 * hd.n_records <= hd.record_ints;
    int64_t expect = (int64_t)sizeof(hd);
    if (sane)
        expect += 12 * hd.nnz + 4 * hd.record_ints + (int64_t)sizeof(CvrChunk) * hd.n_chunks +
                  4 * ((int64_t)hd.n_boundary + hd.n_empty);
    bool size_ok = false;
    if (sane && fseek(f, 0, SEEK_END) == 0) {
        size_ok = (int64_t)ftell(f) == expect;
        fseek(f, (long)sizeof(hd), SEEK_SET);
    }
    if (!sane || !size_ok) {
        fclose(f);
        return fail(CVR_ERR_INVALID, "%s: inconsistent CVR file header (counts out of range or wrong file size)", path);
    }
    cvr_handle* h = new (std::nothrow) cvr_handle();
    if (!h) {
        fclose(f);
        return fail(CVR_ERR_INVALID, "out of host memory");
    }
    h->device = device;
    h->n_rows = hd.n_rows;
    h->n_cols = hd.n_cols;
    h->nnz = hd.nnz;
    h->n_chunks = hd.n_chunks;
    h->variant = cvr_pick_sweep_variant(hd.nnz, hd.n_rows);
    h->record_ints = hd.record_ints;
    h->n_records = hd.n_records;
    h->rows.n_boundary = hd.n_boundary;
    h->rows.n_empty = hd.n_empty;
    std::vector<char> buf((size_t)32 << 20);
    cudaError_t e = cudaSuccess;
    do {
        if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->vals, (size_t)h->nnz)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->cols, (size_t)h->nnz)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->record, (size_t)h->record_ints)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->chunks, (size_t)h->n_chunks)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->rows.boundary, (size_t)hd.n_boundary + 1)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->rows.empty, (size_t)hd.n_empty + 1)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->x, (size_t)h->n_cols + 1)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->y, (size_t)h->n_rows + 1)) != cudaSuccess) break;
        if ((e = dev_alloc(h, &h->done_counter, 4)) != cudaSuccess) break;
        // stream-ordered allocations: the copies below run on the default stream
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) break;
        if ((e = cudaMemset(h->done_counter, 0, 4 * sizeof(unsigned int))) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) {
        fclose(f);
        delete h;
        return fail(CVR_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    }
    rc = copy_in(f, h->vals, sizeof(double) * (size_t)h->nnz, buf);
    if (rc == CVR_OK) rc = copy_in(f, h->cols, sizeof(int32_t) * (size_t)h->nnz, buf);
    if (rc == CVR_OK) rc = copy_in(f, h->record, sizeof(int32_t) * (size_t)h->record_ints, buf);
    if (rc == CVR_OK) rc = copy_in(f, h->chunks, sizeof(CvrChunk) * (size_t)h->n_chunks, buf);
    if (rc == CVR_OK) rc = copy_in(f, h->rows.boundary, sizeof(int32_t) * (size_t)hd.n_boundary, buf);
    if (rc == CVR_OK) rc = copy_in(f, h->rows.empty, sizeof(int32_t) * (size_t)hd.n_empty, buf);
    fclose(f);
    if (rc != CVR_OK) {
        delete h;
        return rc;
    }
    h->host_chunks.resize((size_t)h->n_chunks);
    if (cudaMemcpy(h->host_chunks.data(), h->chunks, sizeof(CvrChunk) * (size_t)h->n_chunks,
                   cudaMemcpyDeviceToHost) != cudaSuccess) {
        delete h;
        return fail(CVR_ERR_CUDA, "descriptor copy failed");
    }
    // chunk descriptors: contiguous nnz slices in multiples of 16, rows in range and non-decreasing, record
    // regions inside the record array -- the kernels index with these without further checks
    {
        int64_t at = 0, n_records = 0;
        int32_t prev_row = 0;
        bool ok = true;
        for (int32_t t = 0; t < h->n_chunks && ok; t++) {
            const CvrChunk& c = h->host_chunks[(size_t)t];
            ok = c.start == at && c.len >= 16 && c.len % 16 == 0 && c.first_row >= prev_row && c.last_row >= c.first_row &&
                 c.last_row <= h->n_rows && c.n_rec >= 0 && c.split0 >= 0 && c.split0 < c.len && c.split1 >= -1 &&
                 c.split1 < c.len &&
                 cvr_record_offset(t, c.first_row) + 2 * ((int64_t)c.n_rec + CVR_W) <= h->record_ints;
            for (int q = 0; q < CVR_W && ok; q++) ok = c.tail[q] >= 0 && c.tail[q] <= h->n_rows;
            at += c.len;
            n_records += c.n_rec + CVR_W;
            prev_row = c.first_row;
        }
        if (!ok || at != h->nnz || n_records != h->n_records) {
            delete h;
            return fail(CVR_ERR_INVALID, "%s: chunk descriptors are inconsistent with the header", path);
        }
    }
    h->create_seconds = wall_seconds() - t0;
    *out = h;
    return CVR_OK;
}

void cvr_destroy(cvr_handle_t* h) { delete h; }

} // extern "C"
