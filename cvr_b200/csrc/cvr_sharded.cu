// Multi-GPU host of the CVR SpMV path, C++ behind the C ABI (include/cvr_b200.h: cvr_create_sharded,
// cvr_sharded_spmv, ...): ONE process drives 1..8 devices.
//
// No counterpart in the reference (a single shared-memory process): this is the north
// star's multi-GPU extension.  The matrix is cut into contiguous row ranges of (nearly)
// equal nnz, snapped to row starts -- the bisection the reference uses for its per-thread slices
// (/root/reference/spmv.cpp:631-650) -- each device converts its re-based shard (own x16 padding,
// spmv.cpp:474-482 applied per shard) to its own CVR and reads a replicated x.  A single SpMV needs no
// communication.  The iterated SpMV x <- A x (square A) needs exactly one exchange per iteration,
// y shards -> everybody's x:
//   * CVR_SHARD_PEER (default): fused into the sweep kernel -- finished rows are stored straight into
//     the peers' next x over NVLink peer memory (cudaDeviceEnablePeerAccess), only to the devices whose
//     shard reads them (column footprints), accumulated rows follow from the epilogue kernel, iterations
//     are separated by a flag barrier over peer memory (cvr_spmv.cu).  x is double-buffered.
//   * CVR_SHARD_NCCL: the sweep, then one grouped NCCL broadcast per shard (all-gather-v) -- the plain
//     collective, kept as the verified fallback (libnccl is dlopen'ed on first use; devices listed twice
//     cannot form a communicator and use cudaMemcpyPeerAsync instead).
// One host thread per device enqueues that device's launches, so the host never serialises the devices.
// A device may be listed more than once (two shards on one GPU): that is how the single-GPU test box
// exercises this code; programmatic dependent launches are switched off then (a waiting dependent grid
// would hold the SM slots the other shard's sweep needs).
#include "../../include/cvr_b200.h"
#include "cvr_internal.h"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

int cvr_set_error(int code, const char* fmt, ...);
#define fail cvr_set_error

namespace {

// ---- the few NCCL entry points used, resolved at run time
typedef struct ncclComm* ncclComm_t;
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        const char* names[] = {getenv("CVR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(lib, "ncclBroadcast"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast;
    }
};
constexpr int NCCL_FLOAT64 = 8; // ncclFloat64 / ncclDouble (nccl.h)

double wall_seconds()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Part {
    int device = 0;
    cvr_handle_t* h = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t lo = 1, hi = 1;   // global rows [lo, hi), 1-based
    int64_t n_local = 0;      // rows of the shard (>= 1 even for an empty range)
    double* X[2] = {nullptr, nullptr}; // replicated x, double-buffered (n_cols + 1 each)
    uint32_t* flags = nullptr;         // this part's barrier flag array (CVR_MAX_PEERS words)
    uint8_t* needs = nullptr;          // needs[local row] bit q: part q reads that x entry
    uint8_t* chunk_any = nullptr;
    int64_t peer_bytes = 0;            // bytes stored into OTHER parts' buffers per iteration
    int rc = CVR_OK;                   // result of the worker thread
    char err[256] = "";
};

} // namespace

struct cvr_sharded {
    int n = 0;
    int flags = 0;
    bool distinct = true; // every device listed once
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    std::vector<int64_t> cuts;
    Part parts[CVR_MAX_PEERS];
    uint32_t epoch = 0;
    double create_seconds = 0.0;
    Nccl nccl;
    ncclComm_t comms[CVR_MAX_PEERS] = {};
    bool comms_ready = false;

    // frees everything the parts hold (the handle can then be rebuilt with other cut points)
    void release()
    {
        for (int g = 0; g < n; g++) {
            Part& p = parts[g];
            cudaSetDevice(p.device);
            if (p.stream) cudaStreamSynchronize(p.stream);
        }
        if (comms_ready)
            for (int g = 0; g < n; g++)
                if (comms[g]) nccl.CommDestroy(comms[g]);
        comms_ready = false;
        for (int g = 0; g < n; g++) {
            Part& p = parts[g];
            cudaSetDevice(p.device);
            cudaFree(p.X[0]);
            cudaFree(p.X[1]);
            cudaFree(p.flags);
            cudaFree(p.needs);
            cudaFree(p.chunk_any);
            if (p.ev0) cudaEventDestroy(p.ev0);
            if (p.ev1) cudaEventDestroy(p.ev1);
            if (p.stream) cudaStreamDestroy(p.stream);
            if (p.h) cvr_destroy(p.h);
            p = Part();
            comms[g] = nullptr;
        }
        epoch = 0;
    }
    ~cvr_sharded() { release(); }
};

namespace {

int64_t delim_at(const cvr_csr_t* csr, int64_t k)
{
    return csr->row_delim32 ? (int64_t)csr->row_delim32[k] : csr->row_delim64[k];
}

// Cut points c[0..G] over the 1-based rows: part g owns rows c[g] .. c[g+1]-1 (c[0] = 1, c[G] = n_rows+1);
// the cut for part g is the first ROW START at or after g * W / G of the cumulative weight W, so no row
// straddles two devices.  The weight of a row is its nnz (row_weight = 0, the north star's rule) or
// nnz + row_weight if it is not empty: a finished row costs the sweep about as much as row_weight nonzeros
// (record, y store, publishing), measured on R-MAT-24 (profiles/r02_strong_scaling_rmat24.txt).
std::vector<int64_t> partition_rows_by_nnz(const cvr_csr_t* csr, int parts, int64_t last_delim, double row_weight)
{
    std::vector<int64_t> cuts((size_t)parts + 1);
    cuts[0] = 1;
    cuts[(size_t)parts] = csr->n_rows + 1;
    std::vector<int64_t> weighted; // weighted[k] = weight of rows < k, only built when row_weight != 0
    if (row_weight > 0.0) {
        weighted.resize((size_t)csr->n_rows + 2);
        double extra = 0.0;
        weighted[0] = 0;
        for (int64_t k = 1; k <= csr->n_rows + 1; k++) {
            if (delim_at(csr, k) > delim_at(csr, k - 1)) extra += row_weight;
            weighted[(size_t)k] = delim_at(csr, k) + (int64_t)extra;
        }
    }
    auto start_weight = [&](int64_t k) { return weighted.empty() ? delim_at(csr, k) : weighted[(size_t)k]; };
    const int64_t total = weighted.empty() ? last_delim : weighted[(size_t)csr->n_rows + 1];
    for (int g = 1; g < parts; g++) {
        const int64_t target = (total * g) / parts;
        int64_t lo = 1, hi = csr->n_rows + 1; // first k in [1, n_rows+1] with start_weight(k) >= target
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (start_weight(mid) >= target) hi = mid;
            else lo = mid + 1;
        }
        cuts[(size_t)g] = std::min(std::max(lo, cuts[(size_t)g - 1]), csr->n_rows + 1);
    }
    return cuts;
}

// New cut points from MEASURED per-part sweep seconds (the C++ twin of cvr_b200/shard.py::rebalance_cuts): every row is
// charged its model weight (nnz, + row_weight if not empty) scaled by (seconds of its part / model weight of its part),
// and the rows are re-cut into parts of equal charged cost.
std::vector<int64_t> rebalance_cuts(const cvr_csr_t* csr, const std::vector<int64_t>& cuts, const double* seconds,
                                    double row_weight)
{
    const int parts = (int)cuts.size() - 1;
    auto weight = [&](int64_t r) { // row r, 1-based
        const int64_t n = delim_at(csr, r + 1) - delim_at(csr, r);
        return (double)n + (n > 0 ? row_weight : 0.0);
    };
    std::vector<double> scale((size_t)parts, 0.0);
    double total = 0.0;
    for (int g = 0; g < parts; g++) {
        double w = 0.0;
        for (int64_t r = cuts[(size_t)g]; r < cuts[(size_t)g + 1]; r++) w += weight(r);
        scale[(size_t)g] = w > 0.0 ? seconds[g] / w : 0.0;
        total += w > 0.0 ? seconds[g] : 0.0;
    }
    std::vector<int64_t> out((size_t)parts + 1);
    out[0] = 1;
    out[(size_t)parts] = csr->n_rows + 1;
    double cum = 0.0;
    int next = 1, g = 0;
    for (int64_t r = 1; r <= csr->n_rows && next < parts; r++) {
        while (g + 1 < parts && r >= cuts[(size_t)g + 1]) g++;
        cum += weight(r) * scale[(size_t)g];
        // rows 1..r reach the target of cut `next`: the part starts at the row after
        while (next < parts && cum >= total * next / parts) out[(size_t)next++] = std::min(r + 1, csr->n_rows + 1);
    }
    for (; next < parts; next++) out[(size_t)next] = csr->n_rows + 1;
    for (int k = 1; k <= parts; k++) out[(size_t)k] = std::max(out[(size_t)k], out[(size_t)k - 1]);
    return out;
}

// Rows [lo, hi) of the host CSR as a CSR of its own: local rows 1..n, global columns, delimiters re-based,
// nnz padded to a multiple of 16 with zero-valued copies of the shard's last entry (they extend its last
// non-empty row).  An empty range becomes one row of 16 explicit zeros (the conversion needs nnz >= 16).
struct HostShard {
    std::vector<double> val;
    std::vector<int32_t> col;
    std::vector<int64_t> rd;
    int64_t n_local = 0;
};

void build_shard(const cvr_csr_t* csr, int64_t lo, int64_t hi, int64_t true_end, HostShard* out)
{
    const int64_t a = delim_at(csr, lo), b = std::min(delim_at(csr, hi), true_end);
    const int64_t n = std::max<int64_t>(b - a, 0);
    out->n_local = std::max<int64_t>(hi - lo, 1);
    out->rd.assign((size_t)out->n_local + 2, 0);
    if (n == 0) {
        out->val.assign(16, 0.0);
        out->col.assign(16, 1);
        for (size_t k = 2; k < out->rd.size(); k++) out->rd[k] = 16;
        return;
    }
    const int64_t npad = (n + 15) / 16 * 16;
    out->val.assign(csr->val + a, csr->val + b);
    out->col.assign(csr->col + a, csr->col + b);
    out->val.resize((size_t)npad, 0.0);
    out->col.resize((size_t)npad, out->col[(size_t)n - 1]);
    int64_t last_nonempty = 1;
    for (int64_t r = 0; r <= hi - lo; r++) {
        const int64_t v = std::min(delim_at(csr, lo + r), true_end) - a;
        out->rd[(size_t)r + 1] = v;
        if (r >= 1 && v > out->rd[(size_t)r]) last_nonempty = r;
    }
    for (int64_t k = last_nonempty + 1; k < (int64_t)out->rd.size(); k++) out->rd[(size_t)k] += npad - n;
}

#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(CVR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                      \
    } while (0)

// who reads what: needs_g[local row] bit q set <=> part q's shard has a nonzero in that (global) column
int build_needs(cvr_sharded* s)
{
    const int n = s->n;
    const int64_t ncol = s->n_cols + 1;
    std::vector<std::vector<uint8_t>> used((size_t)n);
    for (int g = 0; g < n; g++) {
        Part& p = s->parts[g];
        CUDA_OK(cudaSetDevice(p.device));
        uint8_t* d_used = nullptr;
        CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d_used), (size_t)ncol));
        CUDA_OK(cudaMemsetAsync(d_used, 0, (size_t)ncol, p.stream));
        int rc = cvr_column_footprint(p.h, d_used, p.stream);
        if (rc != CVR_OK) {
            cudaFree(d_used);
            return rc;
        }
        used[(size_t)g].resize((size_t)ncol);
        cudaError_t e = cudaMemcpyAsync(used[(size_t)g].data(), d_used, (size_t)ncol, cudaMemcpyDeviceToHost, p.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p.stream);
        cudaFree(d_used);
        if (e != cudaSuccess) return fail(CVR_ERR_CUDA, "footprint copy failed: %s", cudaGetErrorString(e));
    }
    for (int g = 0; g < n; g++) {
        Part& p = s->parts[g];
        std::vector<uint8_t> needs((size_t)p.n_local + 1, 0);
        int64_t sent = 0;
        for (int64_t r = 1; r <= p.hi - p.lo; r++) {
            uint8_t m = 0;
            for (int q = 0; q < n; q++)
                if (q != g && used[(size_t)q][(size_t)(p.lo + r - 1)]) m |= (uint8_t)(1u << q);
            needs[(size_t)r] = m;
            sent += __builtin_popcount(m);
        }
        p.peer_bytes = 8 * sent;
        CUDA_OK(cudaSetDevice(p.device));
        CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p.needs), needs.size()));
        CUDA_OK(cudaMemcpy(p.needs, needs.data(), needs.size(), cudaMemcpyHostToDevice));
        cvr_info_t info;
        int rc = cvr_get_info(p.h, &info);
        if (rc != CVR_OK) return rc;
        CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p.chunk_any), (size_t)info.n_chunks));
        rc = cvr_chunk_needs(p.h, p.needs, p.chunk_any, p.stream);
        if (rc != CVR_OK) return rc;
        CUDA_OK(cudaStreamSynchronize(p.stream));
    }
    return CVR_OK;
}

int init_nccl(cvr_sharded* s)
{
    if (s->comms_ready || !s->distinct) return CVR_OK;
    if (!s->nccl.load())
        return fail(CVR_ERR_STATE, "CVR_SHARD_NCCL: cannot load libnccl.so.2 (%s); set CVR_NCCL_LIB", dlerror());
    int devs[CVR_MAX_PEERS];
    for (int g = 0; g < s->n; g++) devs[g] = s->parts[g].device;
    const int r = s->nccl.CommInitAll(s->comms, s->n, devs);
    if (r != 0)
        return fail(CVR_ERR_CUDA, "ncclCommInitAll failed: %s", s->nccl.GetErrorString ? s->nccl.GetErrorString(r) : "?");
    s->comms_ready = true;
    return CVR_OK;
}

// y shards -> every part's `dst_parity` x buffer, after the sweeps (the plain collective)
int exchange_after_kernel(cvr_sharded* s, int dst_parity)
{
    const int n = s->n;
    if (s->distinct && n > 1) {
        int rc = init_nccl(s);
        if (rc != CVR_OK) return rc;
        if (s->nccl.GroupStart() != 0) return fail(CVR_ERR_CUDA, "ncclGroupStart failed");
        for (int root = 0; root < n; root++) {
            const Part& src = s->parts[root];
            double* y_root = nullptr;
            cvr_device_vectors(src.h, nullptr, &y_root);
            for (int g = 0; g < n; g++) {
                const Part& p = s->parts[g];
                const int r = s->nccl.Broadcast(y_root + 1, p.X[dst_parity] + src.lo, (size_t)(src.hi - src.lo),
                                                NCCL_FLOAT64, root, s->comms[g], p.stream);
                if (r != 0) {
                    s->nccl.GroupEnd();
                    return fail(CVR_ERR_CUDA, "ncclBroadcast failed: %s",
                                s->nccl.GetErrorString ? s->nccl.GetErrorString(r) : "?");
                }
            }
        }
        if (s->nccl.GroupEnd() != 0) return fail(CVR_ERR_CUDA, "ncclGroupEnd failed");
        return CVR_OK;
    }
    // a device listed twice cannot join one communicator twice: plain peer copies, same data movement
    for (int root = 0; root < n; root++) {
        const Part& src = s->parts[root];
        double* y_root = nullptr;
        cvr_device_vectors(src.h, nullptr, &y_root);
        CUDA_OK(cudaSetDevice(src.device));
        CUDA_OK(cudaStreamSynchronize(src.stream)); // y of the root is complete
        for (int g = 0; g < n; g++) {
            const Part& p = s->parts[g];
            CUDA_OK(cudaMemcpyPeerAsync(p.X[dst_parity] + src.lo, p.device, y_root + 1, src.device,
                                        sizeof(double) * (size_t)(src.hi - src.lo), p.stream));
        }
    }
    return CVR_OK;
}

int sync_all(cvr_sharded* s)
{
    for (int g = 0; g < s->n; g++) {
        CUDA_OK(cudaSetDevice(s->parts[g].device));
        CUDA_OK(cudaStreamSynchronize(s->parts[g].stream));
    }
    return CVR_OK;
}

// Builds the parts of `s` for the cut points in s->cuts: shard CSRs, CVR conversion on every device, x buffers,
// flags, peer access, footprints.
int build_parts(cvr_sharded* s, const cvr_csr_t* view_ptr, const cvr_csr_t* csr, int32_t n_chunks_per_device,
                const int* devices, int n_devices, int flags)
{
    int rc = CVR_OK;
    for (int g = 0; g < n_devices && rc == CVR_OK; g++) {
        Part& p = s->parts[g];
        p.device = devices[g];
        p.lo = s->cuts[(size_t)g];
        p.hi = s->cuts[(size_t)g + 1];
        HostShard hs;
        build_shard(view_ptr, p.lo, p.hi, csr->nnz, &hs);
        p.n_local = hs.n_local;
        cvr_csr_t sub{};
        sub.n_rows = hs.n_local;
        sub.n_cols = csr->n_cols;
        sub.nnz = (int64_t)hs.val.size();
        sub.val = hs.val.data();
        sub.col = hs.col.data();
        sub.row_delim64 = hs.rd.data();
        std::vector<int32_t> rd32;
        if (sub.nnz <= 0x7fffffffLL) { // the 32-bit entry where it fits, like the single-GPU path
            rd32.assign(hs.rd.begin(), hs.rd.end());
            sub.row_delim32 = rd32.data();
            sub.row_delim64 = nullptr;
        }
        rc = cvr_create(&sub, n_chunks_per_device, p.device, &p.h);
        if (rc != CVR_OK) break;
        cudaError_t e = cudaSetDevice(p.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&p.ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&p.ev1);
        const size_t xbytes = sizeof(double) * (size_t)(std::max(csr->n_cols, csr->n_rows) + 1);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p.X[0]), xbytes);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p.X[1]), xbytes);
        if (e == cudaSuccess) e = cudaMemset(p.X[0], 0, xbytes);
        if (e == cudaSuccess) e = cudaMemset(p.X[1], 0, xbytes);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p.flags), sizeof(uint32_t) * 64);
        if (e == cudaSuccess) e = cudaMemset(p.flags, 0, sizeof(uint32_t) * 64);
        if (e != cudaSuccess) rc = fail(CVR_ERR_CUDA, "device %d set-up failed: %s", p.device, cudaGetErrorString(e));
    }
    // every device may store into every other device's x buffers and flag array
    if (rc == CVR_OK && n_devices > 1) {
        for (int a = 0; a < n_devices && rc == CVR_OK; a++)
            for (int b = 0; b < n_devices && rc == CVR_OK; b++) {
                const int da = devices[a], db = devices[b];
                if (da == db) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, da, db);
                if (!can) {
                    if (!(flags & CVR_SHARD_NCCL))
                        rc = fail(CVR_ERR_CUDA, "device %d cannot access device %d: use CVR_SHARD_NCCL", da, db);
                    continue;
                }
                cudaSetDevice(da);
                const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    rc = fail(CVR_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", da, db, cudaGetErrorString(e));
                cudaGetLastError();
            }
    }
    if (rc == CVR_OK && n_devices > 1 && !(flags & CVR_SHARD_NCCL) && !(flags & CVR_SHARD_DENSE) &&
        csr->n_rows == csr->n_cols)
        rc = build_needs(s);
    return rc;
}

} // namespace

extern "C" {

static // Seconds of the sweep kernel of every part inside the real iteration (x <- A x, publishing included): a few warm-up
// iterations, then timed ones with per-launch events on every part's handle.
int measure_parts(cvr_sharded* s, double* seconds)
{
    std::vector<double> x((size_t)s->n_cols + 1, 1.0), y((size_t)s->n_rows + 1, 0.0);
    x[0] = 0.0;
    int rc = cvr_sharded_spmv(s, x.data(), y.data(), 3, 1, nullptr);
    for (int g = 0; g < s->n && rc == CVR_OK; g++) rc = cvr_set_kernel_timing(s->parts[g].h, 1);
    if (rc == CVR_OK) rc = cvr_sharded_spmv(s, x.data(), y.data(), 6, 1, nullptr);
    for (int g = 0; g < s->n; g++) {
        double total = 0.0;
        int64_t launches = 0;
        if (rc == CVR_OK) rc = cvr_get_kernel_timing(s->parts[g].h, &total, &launches);
        seconds[g] = launches > 0 ? total / (double)launches : 0.0;
        cvr_set_kernel_timing(s->parts[g].h, 0);
    }
    return rc;
}

int cvr_create_sharded(const cvr_csr_t* csr, int32_t n_chunks_per_device, const int* devices, int n_devices,
                       int flags, cvr_sharded_t** out)
{
    if (!out) return fail(CVR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!csr || !csr->val || !csr->col) return fail(CVR_ERR_INVALID, "csr is NULL / incomplete");
    if ((csr->row_delim32 == nullptr) == (csr->row_delim64 == nullptr))
        return fail(CVR_ERR_INVALID, "exactly one of row_delim32 / row_delim64 must be set");
    if (!devices || n_devices < 1 || n_devices > CVR_MAX_PEERS)
        return fail(CVR_ERR_INVALID, "n_devices = %d must be in [1, %d]", n_devices, CVR_MAX_PEERS);
    if (csr->n_rows < 1 || csr->n_cols < 1 || csr->nnz < 16 || csr->nnz % 16)
        return fail(CVR_ERR_INVALID, "bad matrix dimensions / nnz (must be a positive multiple of 16)");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CVR_ERR_CUDA, "no CUDA device: libcvr_b200 has no CPU fallback");
    for (int g = 0; g < n_devices; g++)
        if (devices[g] < 0 || devices[g] >= n_dev)
            return fail(CVR_ERR_INVALID, "devices[%d] = %d out of range [0, %d)", g, devices[g], n_dev);
    // the reference reader's trailing delimiter nnz-1 (spmv.cpp:522-526) is accepted like in cvr_create
    int64_t true_end = delim_at(csr, csr->n_rows + 1);
    if (true_end == csr->nnz - 1) true_end = csr->nnz;
    else if (true_end != csr->nnz)
        return fail(CVR_ERR_INVALID, "row_delim[n_rows+1] = %lld must be nnz = %lld (or nnz-1)", (long long)true_end,
                    (long long)csr->nnz);

    const double t0 = wall_seconds();
    cvr_sharded* s = new (std::nothrow) cvr_sharded();
    if (!s) return fail(CVR_ERR_INVALID, "out of host memory");
    s->n = n_devices;
    s->flags = flags;
    s->n_rows = csr->n_rows;
    s->n_cols = csr->n_cols;
    s->nnz = csr->nnz;
    for (int a = 0; a < n_devices; a++)
        for (int b = a + 1; b < n_devices; b++)
            if (devices[a] == devices[b]) s->distinct = false;
    // with the reference quirk the last non-empty row owns the last element: cut on corrected delimiters
    std::vector<int64_t> fixed;
    cvr_csr_t view = *csr;
    if (delim_at(csr, csr->n_rows + 1) != csr->nnz) {
        fixed.resize((size_t)csr->n_rows + 2);
        for (int64_t k = 0; k <= csr->n_rows + 1; k++) {
            const int64_t v = delim_at(csr, k);
            fixed[(size_t)k] = v == csr->nnz - 1 ? csr->nnz : v;
        }
        view.row_delim32 = nullptr;
        view.row_delim64 = fixed.data();
    }
    double row_weight = 0.0; // CVR_SHARD_ROW_WEIGHT: balance nnz + w per non-empty row instead of nnz alone
    if (const char* e = getenv("CVR_SHARD_ROW_WEIGHT")) row_weight = atof(e);
    s->cuts = partition_rows_by_nnz(&view, n_devices, csr->nnz, row_weight);

    int rc = build_parts(s, &view, csr, n_chunks_per_device, devices, n_devices, flags);
    // CVR_SHARD_REBALANCE = R: up to R rounds of re-cutting the shards from the MEASURED sweep time of every part
    // (iterated SpMV only, i.e. square matrices; cvr_b200/dist.py does the same for the one-process-per-GPU host)
    int rounds = 0;
    if (const char* e = getenv("CVR_SHARD_REBALANCE")) rounds = atoi(e);
    if (n_devices > 1 && csr->n_rows == csr->n_cols)
        for (int r = 0; r < rounds && rc == CVR_OK; r++) {
            double seconds[CVR_MAX_PEERS] = {};
            rc = measure_parts(s, seconds);
            if (rc != CVR_OK) break;
            double worst = 0.0, mean = 0.0;
            for (int g = 0; g < n_devices; g++) {
                worst = std::max(worst, seconds[g]);
                mean += seconds[g] / n_devices;
            }
            if (worst <= 1.03 * mean) break;
            const std::vector<int64_t> cuts = rebalance_cuts(&view, s->cuts, seconds, row_weight);
            if (cuts == s->cuts) break;
            s->release();
            s->cuts = cuts;
            rc = build_parts(s, &view, csr, n_chunks_per_device, devices, n_devices, flags);
        }
    if (rc != CVR_OK) {
        delete s;
        return rc;
    }
    s->create_seconds = wall_seconds() - t0;
    *out = s;
    return CVR_OK;
}

int cvr_sharded_spmv(cvr_sharded_t* s, const double* x_host, double* y_host, int32_t iters, int32_t feed_y_to_x,
                     double* seconds_per_iter)
{
    if (!s || !x_host || !y_host) return fail(CVR_ERR_INVALID, "NULL argument");
    if (iters < 1) return fail(CVR_ERR_INVALID, "iters = %d must be >= 1", iters);
    if (feed_y_to_x && s->n_rows != s->n_cols)
        return fail(CVR_ERR_INVALID, "feed_y_to_x needs a square matrix (%lld x %lld)", (long long)s->n_rows,
                    (long long)s->n_cols);
    const int n = s->n;
    // ---- replicate x: every device reads all of it
    for (int g = 0; g < n; g++) {
        Part& p = s->parts[g];
        CUDA_OK(cudaSetDevice(p.device));
        CUDA_OK(cudaMemcpyAsync(p.X[0], x_host, sizeof(double) * (size_t)(s->n_cols + 1), cudaMemcpyHostToDevice,
                                p.stream));
    }
    int rc = sync_all(s);
    if (rc != CVR_OK) return rc;

    const bool fused = feed_y_to_x && n > 1 && !(s->flags & CVR_SHARD_NCCL);
    const double t0 = wall_seconds();
    if (!feed_y_to_x || n == 1) {
        // the reference's loop (spmv.cpp:1024-1034): the same x every iteration, no communication; on one
        // device the iterated form just swaps the two x buffers
        std::vector<std::thread> th;
        for (int g = 0; g < n; g++)
            th.emplace_back([s, g, iters, feed_y_to_x]() {
                Part& p = s->parts[g];
                cudaSetDevice(p.device);
                double* y = nullptr;
                cvr_device_vectors(p.h, nullptr, &y);
                for (int32_t it = 0; it < iters && p.rc == CVR_OK; it++) {
                    if (feed_y_to_x) {
                        // single shard: y is written straight over rows 1..n of the other buffer (y[0] = x[0])
                        p.rc = cvr_spmv_device(p.h, p.X[it & 1], p.X[(it + 1) & 1], p.stream);
                    } else {
                        p.rc = cvr_spmv_device(p.h, p.X[0], y, p.stream);
                    }
                    if (p.rc != CVR_OK) snprintf(p.err, sizeof(p.err), "%s", cvr_last_error());
                }
                cudaStreamSynchronize(p.stream);
            });
        for (auto& t : th) t.join();
    } else if (fused) {
        // ---- x <- A x with the exchange fused into the sweep (cvr_spmv_publish), one host thread per device
        const uint32_t epoch0 = s->epoch;
        std::vector<std::thread> th;
        for (int g = 0; g < n; g++)
            th.emplace_back([s, g, n, iters, epoch0]() {
                Part& p = s->parts[g];
                cudaSetDevice(p.device);
                void* flag_arrays[CVR_MAX_PEERS];
                for (int q = 0; q < n; q++) flag_arrays[q] = s->parts[q].flags;
                for (int32_t it = 0; it < iters && p.rc == CVR_OK; it++) {
                    const int cur = it & 1, nxt = cur ^ 1;
                    cvr_publish_t pub{};
                    pub.n_dst = n;
                    pub.self = g;
                    // bit 2: y is my slice of the next x; bit 1 from the third iteration on (both buffers then
                    // hold 0.0 at the never-written rows); bit 3: no programmatic launches when a device
                    // carries two shards
                    pub.mode = 4 | (it >= 2 ? 2 : 0) | (s->distinct ? 0 : 8);
                    pub.row_offset = p.lo - 1;
                    pub.needs = p.needs;
                    pub.chunk_any = p.needs ? p.chunk_any : nullptr;
                    pub.clear_next = p.X[cur] + (p.lo - 1);
                    for (int q = 0; q < n; q++) pub.dst[q] = s->parts[q].X[nxt];
                    p.rc = cvr_spmv_publish(p.h, p.X[cur], p.X[nxt] + (p.lo - 1), &pub, flag_arrays, g, n,
                                            epoch0 + (uint32_t)it + 1u, it > 0, p.stream);
                    if (p.rc != CVR_OK) snprintf(p.err, sizeof(p.err), "%s", cvr_last_error());
                }
                cudaStreamSynchronize(p.stream);
                if (p.rc == CVR_OK) {
                    p.rc = cvr_check_async_error(p.h);
                    if (p.rc != CVR_OK) snprintf(p.err, sizeof(p.err), "%s", cvr_last_error());
                }
            });
        for (auto& t : th) t.join();
        s->epoch = epoch0 + (uint32_t)iters;
    } else {
        // ---- x <- A x with the plain collective after each sweep (NCCL broadcasts, or peer copies)
        for (int32_t it = 0; it < iters && rc == CVR_OK; it++) {
            const int cur = it & 1, nxt = cur ^ 1;
            for (int g = 0; g < n && rc == CVR_OK; g++) {
                Part& p = s->parts[g];
                double* y = nullptr;
                cvr_device_vectors(p.h, nullptr, &y);
                rc = cvr_spmv_device(p.h, p.X[cur], y, p.stream);
            }
            if (rc == CVR_OK) rc = exchange_after_kernel(s, nxt);
            if (rc == CVR_OK && !s->distinct) rc = sync_all(s); // peer copies of one device: order the streams
        }
        if (rc == CVR_OK) rc = sync_all(s);
        if (rc != CVR_OK) return rc;
    }
    for (int g = 0; g < n; g++)
        if (s->parts[g].rc != CVR_OK) {
            const int r = s->parts[g].rc;
            s->parts[g].rc = CVR_OK;
            return fail(r, "device %d: %s", s->parts[g].device, s->parts[g].err);
        }
    const double secs = wall_seconds() - t0;
    if (seconds_per_iter) *seconds_per_iter = secs / iters;

    // ---- collect: every part returns its own rows
    const int fin = iters & 1; // buffer that holds the last iterate when feeding y back
    y_host[0] = 0.0;
    for (int g = 0; g < n; g++) {
        Part& p = s->parts[g];
        if (p.hi <= p.lo) continue;
        CUDA_OK(cudaSetDevice(p.device));
        const double* src;
        if (feed_y_to_x) src = p.X[fin] + p.lo;
        else {
            double* y = nullptr;
            cvr_device_vectors(p.h, nullptr, &y);
            src = y + 1;
        }
        CUDA_OK(cudaMemcpyAsync(y_host + p.lo, src, sizeof(double) * (size_t)(p.hi - p.lo), cudaMemcpyDeviceToHost,
                                p.stream));
    }
    return sync_all(s);
}

int cvr_sharded_get_info(cvr_sharded_t* s, cvr_sharded_info_t* info)
{
    if (!s || !info) return fail(CVR_ERR_INVALID, "NULL argument");
    memset(info, 0, sizeof(*info));
    info->n_parts = s->n;
    info->n_rows = s->n_rows;
    info->n_cols = s->n_cols;
    info->nnz = s->nnz;
    info->create_seconds = s->create_seconds;
    info->exchange = (s->flags & CVR_SHARD_NCCL) ? 1 : 0;
    for (int g = 0; g < s->n; g++) {
        info->device[g] = s->parts[g].device;
        info->row_begin[g] = s->parts[g].lo;
        info->row_end[g] = s->parts[g].hi;
        info->peer_bytes_per_iter[g] = s->parts[g].peer_bytes;
        cvr_info_t pi;
        const int rc = cvr_get_info(s->parts[g].h, &pi);
        if (rc != CVR_OK) return rc;
        info->part_nnz[g] = pi.nnz;
        info->part_chunks[g] = pi.n_chunks;
        info->convert_seconds += pi.convert_seconds;
        info->kernel_launches += pi.kernel_launches;
    }
    return CVR_OK;
}

int cvr_sharded_part(cvr_sharded_t* s, int part, cvr_handle_t** handle)
{
    if (!s || !handle || part < 0 || part >= s->n) return fail(CVR_ERR_INVALID, "bad arguments to cvr_sharded_part");
    *handle = s->parts[part].h;
    return CVR_OK;
}

void cvr_sharded_destroy(cvr_sharded_t* s) { delete s; }

} // extern "C"
