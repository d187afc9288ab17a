// Matrix Market coordinate file -> 1-based, x16-padded CSR (host only).
//
// Replaces readMatrix (/root/reference/spmv.cpp:311-535).  The reference reads the file
// line by line with getline + sscanf into an array of (int x, int y, float val) and
// qsorts it; here the whole file is loaded once, tokenised with hand-written integer
// parsing and strtof, bucketed by row with a stable counting sort and only the rows whose
// columns arrive out of order are sorted.  All counts are 64-bit (the reference's
// `int valSize` overflows at 179 M entries, spmv.cpp:394).
//
// Observable semantics kept from the reference:
//   1. row / column ids stay 1-based; row 0 is an empty phantom row (:437-438, :505)
//   2. values are rounded to float, then widened (:65, :432, :517)
//   3. `pattern` -> value = running entry index % 13, mirrored entries counted (:413-417);
//      `complex` -> real part; `symmetric` (only) is mirrored off the diagonal (:443-449)
//   4. comments are skipped before the size line (:377-383); a final line without '\n' is
//      dropped (:411) unless CVR_MM_KEEP_LAST_LINE
//   5. nnz is padded to a multiple of 16 with zero-valued copies of the last file entry
//      (:457, :474-482)
//   6. entries are ordered by (row, col), duplicates keep file order (:485)
//   7. row_delim after the last row is nnz (correct); the reference's nnz-1 (:522-526)
//      only with CVR_MM_REF_LAST_DELIM
// Deviations: blank lines and '%' lines after the size line are skipped (the reference
// would scan them into garbage entries); ids outside [1, n_rows] x [1, n_cols] are an
// error instead of undefined behaviour.
#include "../../include/cvr_b200.h"

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int cvr_set_error(int code, const char* fmt, ...); // cvr_api.cu

namespace {

struct Entry {
    int32_t row, col;
    float val;
};

inline const char* skip_space(const char* p, const char* end)
{
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    return p;
}

// sscanf("%d") on a token: optional sign, decimal digits
inline const char* parse_int(const char* p, const char* end, long long* out, bool* ok)
{
    p = skip_space(p, end);
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) neg = (*p++ == '-');
    long long v = 0;
    const char* d0 = p;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
    *ok = p > d0;
    *out = neg ? -v : v;
    return p;
}

bool read_file(const char* path, std::vector<char>& buf)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n + 1);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    buf.resize(got + 1);
    buf[got] = '\0'; // strtof needs a terminator
    return true;
}

} // namespace

extern "C" int cvr_read_matrix_market(const char* path, int flags, cvr_host_csr_t* out)
{
    if (!path || !out) return cvr_set_error(CVR_ERR_INVALID, "NULL argument");
    memset(out, 0, sizeof(*out));
    std::vector<char> buf;
    if (!read_file(path, buf))
        return cvr_set_error(CVR_ERR_INVALID, "Error: unable to open matrix file %s", path);
    const char* p = buf.data();
    const char* end = p + buf.size() - 1;

    // ---- banner line (:337-371)
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) return cvr_set_error(CVR_ERR_INVALID, "Error: file %s does not store a matrix", path);
    char id[128] = "", object[128] = "", format[128] = "", field[128] = "", symm[128] = "";
    {
        std::string line(p, eol);
        sscanf(line.c_str(), "%127s %127s %127s %127s %127s", id, object, format, field, symm);
    }
    if (strcmp(object, "matrix") != 0)
        return cvr_set_error(CVR_ERR_INVALID, "Error: file %s does not store a matrix", path);
    if (strcmp(format, "coordinate") != 0)
        return cvr_set_error(CVR_ERR_INVALID, "Error: matrix representation is dense");
    const bool pattern = strcmp(field, "pattern") == 0;
    const bool symmetric = strcmp(symm, "symmetric") == 0;
    p = eol + 1;

    // ---- comments, then the size line (:377-386)
    long long n_rows = 0, n_cols = 0, declared = 0;
    for (;;) {
        eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) return cvr_set_error(CVR_ERR_INVALID, "Error: file %s has no size line", path);
        if (*p != '%') {
            bool ok1, ok2, ok3;
            const char* q = parse_int(p, eol, &n_rows, &ok1);
            q = parse_int(q, eol, &n_cols, &ok2);
            parse_int(q, eol, &declared, &ok3);
            p = eol + 1;
            if (!ok1 || !ok2 || !ok3 || n_rows < 1 || n_cols < 1)
                return cvr_set_error(CVR_ERR_INVALID, "Error: bad size line in %s", path);
            break;
        }
        p = eol + 1;
    }
    if (n_rows > 0x7ffffff0LL || n_cols > 0x7ffffff0LL)
        return cvr_set_error(CVR_ERR_RANGE, "matrix dimensions must fit int32");

    // ---- entries (:411-451).  The region is cut at newline boundaries into one piece per host thread;
    // every piece is tokenised independently (mirrored entries inserted right behind their originals, as
    // the reference does), then the pieces are concatenated in file order.  `pattern` values depend on the
    // running entry index including mirrors (:417) and are filled in after the concatenation.
    bool dropped_last = false;
    const char* region_end = end;
    if (p < end && end[-1] != '\n') { // unterminated last line
        const char* last_nl = end;
        while (last_nl > p && last_nl[-1] != '\n') last_nl--;
        if (!(flags & CVR_MM_KEEP_LAST_LINE)) { // the reference's eof() loop never sees it
            dropped_last = skip_space(last_nl, end) < end;
            region_end = last_nl;
        }
    }
    int n_parts = 1;
#ifdef _OPENMP
    n_parts = omp_get_max_threads();
#endif
    if (region_end - p < (1 << 20)) n_parts = 1; // small files: not worth a team
    std::vector<const char*> cut((size_t)n_parts + 1, region_end);
    cut[0] = p;
    for (int k = 1; k < n_parts; k++) {
        const char* q = p + (region_end - p) / n_parts * k;
        if (q < cut[(size_t)k - 1]) q = cut[(size_t)k - 1];
        while (q < region_end && q[-1] != '\n') q++; // q > p here, so q[-1] is inside the buffer
        cut[(size_t)k] = q;
    }
    std::vector<std::vector<Entry>> parts((size_t)n_parts);
    std::vector<long long> bad_at((size_t)n_parts, -1);
    std::vector<int> bad_kind((size_t)n_parts, 0);
#pragma omp parallel for schedule(static, 1) num_threads(n_parts)
    for (int k = 0; k < n_parts; k++) {
        std::vector<Entry>& out_k = parts[(size_t)k];
        const char* q0 = cut[(size_t)k];
        const char* q1 = cut[(size_t)k + 1];
        out_k.reserve((size_t)((q1 - q0) / 12 + 16) * (symmetric ? 2 : 1));
        const char* lp = q0;
        while (lp < q1) {
            const char* le = (const char*)memchr(lp, '\n', (size_t)(q1 - lp));
            if (!le) le = q1; // only the kept unterminated last line
            const char* q = skip_space(lp, le);
            if (q < le && *q != '%') {
                long long r = 0, c = 0;
                bool okr, okc;
                q = parse_int(q, le, &r, &okr);
                q = parse_int(q, le, &c, &okc);
                if (!okr || !okc || r < 1 || r > n_rows || c < 1 || c > n_cols) {
                    bad_at[(size_t)k] = (long long)(lp - buf.data());
                    bad_kind[(size_t)k] = 1;
                    break;
                }
                Entry e;
                e.row = (int32_t)r;
                e.col = (int32_t)c;
                e.val = 0.0f;
                if (!pattern) {
                    q = skip_space(q, le);
                    e.val = q < le ? strtof(q, nullptr) : 0.0f; // real part for `complex`
                }
                out_k.push_back(e);
                if (symmetric && e.row != e.col) {
                    Entry m;
                    m.row = e.col;
                    m.col = e.row;
                    m.val = pattern ? -1.0f : e.val; // pattern: marker, resolved after the concatenation
                    if (m.row > n_rows || m.col > n_cols) {
                        bad_at[(size_t)k] = (long long)(lp - buf.data());
                        bad_kind[(size_t)k] = 2;
                        break;
                    }
                    out_k.push_back(m);
                }
            }
            lp = (le < q1) ? le + 1 : q1;
        }
    }
    for (int k = 0; k < n_parts; k++) {
        if (bad_kind[(size_t)k] == 1)
            return cvr_set_error(CVR_ERR_INVALID, "bad entry at byte %lld of %s", bad_at[(size_t)k], path);
        if (bad_kind[(size_t)k] == 2)
            return cvr_set_error(CVR_ERR_INVALID, "symmetric entry outside the matrix in %s", path);
    }
    std::vector<Entry> ent;
    {
        size_t total = 0;
        for (const auto& v : parts) total += v.size();
        ent.reserve(total + 16);
        for (auto& v : parts) {
            ent.insert(ent.end(), v.begin(), v.end());
            std::vector<Entry>().swap(v);
        }
    }
    if (pattern) { // value = running index % 13; a mirrored entry (marked -1 above) copies its original (:417, :447)
        for (size_t k = 0; k < ent.size(); k++)
            ent[k].val = (ent[k].val < 0.0f && k > 0) ? ent[k - 1].val : (float)(k % 13);
    }
    if (dropped_last)
        fprintf(stderr, "cvr: %s does not end with a newline; its last line is dropped like the "
                        "reference does (spmv.cpp:411)\n", path);
    if (ent.empty()) return cvr_set_error(CVR_ERR_INVALID, "no entries in %s", path);

    const int64_t n = (int64_t)ent.size();
    const int64_t np = (n % 16 == 0) ? n : (n + 16) / 16 * 16; // :457
    const Entry last = ent.back();
    for (int64_t q = n; q < np; q++) { // :474-482
        Entry z = last;
        z.val = 0.0f;
        ent.push_back(z);
    }

    // ---- stable bucket by row, then order each row by column (:485)
    std::vector<int64_t> rd((size_t)n_rows + 2, 0);
    for (const Entry& e : ent) rd[(size_t)e.row + 1]++;
    // rd[r+1] holds the count of row r; turn into starts: rd[0] = rd[1] = 0 (phantom row 0)
    for (int64_t r = 1; r <= n_rows; r++) rd[(size_t)r + 1] += rd[(size_t)r];
    // now rd[r+1] = end of row r, rd[r] = start of row r
    std::vector<int64_t> cursor(rd.begin(), rd.end() - 1);
    double* val = (double*)malloc(sizeof(double) * (size_t)np);
    int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (size_t)np);
    if (!val || !col) {
        free(val);
        free(col);
        return cvr_set_error(CVR_ERR_INVALID, "out of host memory");
    }
    for (const Entry& e : ent) {
        const int64_t k = cursor[(size_t)e.row]++;
        val[k] = (double)e.val;
        col[k] = e.col;
    }
    std::vector<Entry>().swap(ent);
    std::vector<std::pair<int32_t, double>> tmp;
#pragma omp parallel for schedule(dynamic, 4096) private(tmp)
    for (int64_t r = 1; r <= n_rows; r++) {
        const int64_t a = rd[(size_t)r], b = rd[(size_t)r + 1];
        bool sorted = true;
        for (int64_t k = a + 1; k < b; k++)
            if (col[k] < col[k - 1]) { sorted = false; break; }
        if (sorted) continue;
        tmp.resize((size_t)(b - a));
        for (int64_t k = a; k < b; k++) tmp[(size_t)(k - a)] = {col[k], val[k]};
        std::stable_sort(tmp.begin(), tmp.end(),
                         [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y) {
                             return x.first < y.first;
                         });
        for (int64_t k = a; k < b; k++) {
            col[k] = tmp[(size_t)(k - a)].first;
            val[k] = tmp[(size_t)(k - a)].second;
        }
    }
    if (flags & CVR_MM_REF_LAST_DELIM) { // :522-526
        int64_t last_row = n_rows;
        while (last_row > 1 && rd[(size_t)last_row + 1] == rd[(size_t)last_row]) last_row--;
        for (int64_t k = last_row + 1; k <= n_rows + 1; k++) rd[(size_t)k] = np - 1;
    }

    out->n_rows = n_rows;
    out->n_cols = n_cols;
    out->nnz = np;
    out->nnz_file = n;
    out->val = val;
    out->col = col;
    if (np <= 0x7fffffffLL) {
        out->row_delim32 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_rows + 2));
        for (int64_t k = 0; k < n_rows + 2; k++) out->row_delim32[k] = (int32_t)rd[(size_t)k];
    } else {
        out->row_delim64 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_rows + 2));
        memcpy(out->row_delim64, rd.data(), sizeof(int64_t) * (size_t)(n_rows + 2));
    }
    return CVR_OK;
}

extern "C" void cvr_free_host_csr(cvr_host_csr_t* csr)
{
    if (!csr) return;
    free(csr->val);
    free(csr->col);
    free(csr->row_delim32);
    free(csr->row_delim64);
    memset(csr, 0, sizeof(*csr));
}
