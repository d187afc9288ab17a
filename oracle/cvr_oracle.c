/* TEST INFRASTRUCTURE ONLY -- see cvr_oracle.h for the rules and parity status.
 *
 * Scalar C restatement of the reference CVR path (/root/reference/spmv.cpp).
 * Chunks are independent, so both passes run one OpenMP task per chunk like the
 * reference runs one OpenMP thread per chunk (spmv.cpp:577, :1034), but with
 * any number of chunks on any number of host threads.
 */
#define _GNU_SOURCE
#include "cvr_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define W CVR_ORACLE_LANES

long long cvr_oracle_record_ints(int n_rows, int n_chunks)
{
    return 2LL * ((long long)n_rows + 240 + 32LL * n_chunks); /* spmv.cpp:1806 */
}

long long cvr_oracle_record_offset(int chunk, int first_row)
{
    return (2LL * (32LL * chunk + first_row)) / 16 * 16; /* spmv.cpp:709 */
}

/* largest m in [lo, hi] with delim[m] <= key, by the reference's bisection
 * (spmv.cpp:637-650 / :655-667); returns lo-1 if none. */
static int last_row_not_after(const int* delim, int lo, int hi, int key)
{
    int start = lo, stop = hi;
    while (stop >= start) {
        int mid = (stop + start) / 2;
        if (key >= delim[mid]) start = mid + 1;
        else stop = mid - 1;
    }
    return start - 1;
}

/* nnz range of chunk t (spmv.cpp:584-586, :615-627) */
static void chunk_bounds(int t, int T, int nnz, int* s, int* e)
{
    int per = (nnz / T / 16) * 16;
    int brk = (nnz - per * T) / 16;
    if (t < brk) {
        *s = t * (per + 16);
        *e = (t + 1) * (per + 16);
    } else {
        *s = t * per + brk * 16;
        *e = (t + 1) * per + brk * 16;
    }
    if (t == T - 1) *e = nnz;
}

static void convert_chunk(int t, int T, int nnz, int n_rows,
                          const double* csr_val, const int* csr_col, const int* rd,
                          double* cvr_vals, int* cvr_cols, int* record, int* nnz_rows,
                          int* final_2, int* split, int fill_missing_tail)
{
    int s, e;
    chunk_bounds(t, T, nnz, &s, &e);

    int r0 = last_row_not_after(rd, 0, n_rows, s);      /* :631-650 */
    int r1 = last_row_not_after(rd, r0, n_rows, e - 1); /* :652-667 */
    /* :687-688 -- only moves on a degenerate delimiter tail; the reference
     * would read rd[n_rows+2] here, the port stops at the array end. */
    while (r1 <= n_rows && rd[r1 + 1] == rd[r1]) {
        r1++;
        if (r1 + 1 > n_rows + 1) break;
    }

    nnz_rows[4 * t + 0] = s; /* :690-694 */
    nnz_rows[4 * t + 1] = e;
    nnz_rows[4 * t + 2] = r0;
    nnz_rows[4 * t + 3] = r1;
    const int span = r1 - r0 + 1;

    int* rec = record + cvr_oracle_record_offset(t, r0);
    int* tail = final_2 + 16 * t; /* :705, omega = 1 */
    int n_rec = 0;

    int src[W], row[W], left[W]; /* vPack_valID / rowID / count, :711-714 */
    int steal_from[W], has_stolen[W];
    int next_row = r0;
    for (int l = 0; l < W; l++) { /* :727-759 */
        if (next_row < r1) {
            src[l] = rd[next_row] - s;
            row[l] = next_row;
            left[l] = rd[next_row + 1] - rd[next_row];
        } else if (next_row == r1) {
            src[l] = rd[next_row] - s;
            row[l] = next_row;
            left[l] = e - rd[next_row];
        } else {
            src[l] = row[l] = left[l] = 0;
        }
        if (l == 0) { /* the first row may begin before the chunk does */
            src[0] = 0;
            left[0] = (next_row == r1) ? e - s : rd[next_row + 1] - s;
        }
        steal_from[l] = -1;
        has_stolen[l] = 0;
        next_row++;
    }

    int tail_stored = 0, stealing_started = 0;
    split[2 * t] = 0;
    split[2 * t + 1] = 0;
    const int n_steps = (e - s) / W;

    for (int i = 0; i < n_steps; i++) { /* :808 */
        for (int l = 0; l < W; l++) {
            if (left[l] != 0) continue; /* :810-816 */
            const int pos = i * W + l;
            if (next_row <= r1) {
                /* feeding: the lane takes the next non-empty row, :821-868 */
                if (row[l] == r0) {
                    split[2 * t] = pos; /* :826-829 */
                } else {
                    rec[2 * n_rec] = pos; /* :832-834 */
                    rec[2 * n_rec + 1] = row[l];
                    n_rec++;
                }
                while (rd[next_row + 1] == rd[next_row]) next_row++; /* :837-838 */
                src[l] = rd[next_row] - s;
                row[l] = next_row;
                left[l] = rd[next_row + 1] - rd[next_row];
                if (next_row == r1) { /* :844-857 */
                    if (split[2 * t + 1] == 0) split[2 * t + 1] = pos;
                    left[l] = e - rd[next_row];
                    for (int q = 0; q < W; q++) tail[q] = row[q];
                    tail_stored = 1;
                    for (int q = 0; q < W; q++)
                        if (left[q] == 0) steal_from[q] = 0; /* :855-856, overwritten by the steal */
                }
                next_row++;
            } else {
                /* stealing: split the first above-average lane, :869-943 */
                int total = 0;
                for (int q = 0; q < W; q++) total += left[q];
                const int ave = total / W; /* :871 */
                int victim = 0;
                while (victim < W && !(left[victim] > ave)) victim++; /* :876-879 */
                if (!has_stolen[l]) {
                    if (!stealing_started) { /* :883-896 */
                        if (split[2 * t + 1] == 0) split[2 * t + 1] = (span <= W) ? -1 : pos;
                        for (int q = 0; q < W; q++) tail[q] = row[q];
                        tail_stored = 1;
                        stealing_started = 1;
                    }
                    rec[2 * n_rec] = pos; /* :898-902 */
                    rec[2 * n_rec + 1] = l;
                    has_stolen[l] = 1;
                } else { /* :904-909, unreachable by the invariant in SURVEY 8a-R2 note (i) */
                    rec[2 * n_rec] = pos;
                    rec[2 * n_rec + 1] = steal_from[l];
                }
                n_rec++;
                steal_from[l] = victim;
                src[l] = src[victim]; /* :927-931 */
                row[l] = victim;
                left[l] = ave;
                left[victim] -= ave;
                src[victim] += ave;
            }
        }
        for (int l = 0; l < W; l++) { /* :963-980 */
            cvr_vals[s + i * W + l] = csr_val[s + src[l]];
            cvr_cols[s + i * W + l] = csr_col[s + src[l]];
            src[l]++;
            left[l]--;
        }
    }
    for (int l = 0; l < W; l++) { /* :982-999, emitted inside the last step */
        rec[2 * n_rec] = -1;
        rec[2 * n_rec + 1] = (steal_from[l] == -1) ? l : steal_from[l];
        n_rec++;
    }
    if (!tail_stored && fill_missing_tail)
        for (int q = 0; q < W; q++) tail[q] = row[q];
}

int cvr_oracle_convert(int n_chunks, int nnz, int n_rows,
                       const double* csr_val, const int* csr_col, const int* row_delim,
                       double* cvr_vals, int* cvr_cols, int* record, int* nnz_rows,
                       int* final_2, int* split, int fill_missing_tail)
{
    if (n_chunks < 1 || nnz % 16 != 0 || n_chunks > nnz / 16) return -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_chunks; t++)
        convert_chunk(t, n_chunks, nnz, n_rows, csr_val, csr_col, row_delim, cvr_vals,
                      cvr_cols, record, nnz_rows, final_2, split, fill_missing_tail);
    return 0;
}

static void spmv_chunk(int t, const double* cvr_vals, const int* cvr_cols,
                       const int* record, const int* nnz_rows, const int* final_2,
                       const int* split, const double* x, double* y)
{
    const int s = nnz_rows[4 * t], e = nnz_rows[4 * t + 1], r0 = nnz_rows[4 * t + 2];
    const int first_end = split[2 * t];    /* ncsr_start, :1153 */
    const int feed_end = split[2 * t + 1]; /* ncsr, :1154 */
    const int* rec = record + cvr_oracle_record_offset(t, r0);
    const int* tail = final_2 + 16 * t;
    const double* v = cvr_vals + s;
    const int* c = cvr_cols + s;

    double acc[W], carry[W]; /* r_rets/z_rets and t_rets */
    for (int l = 0; l < W; l++) acc[l] = carry[l] = 0.0;
    int ri = 0;
    const int n_steps = (e - s) / W;
    for (int i = 0; i < n_steps; i++) {
        if (first_end != 0 && first_end / W == i) { /* :1280-1282 */
            const int l = first_end % W;
#pragma omp atomic
            y[r0] += acc[l];
            acc[l] = 0.0;
        }
        while (rec[2 * ri] != -1 && rec[2 * ri] / W == i) {
            const int pos = rec[2 * ri], wb = rec[2 * ri + 1], l = pos % W;
            if (feed_end != -1 && pos <= feed_end)
                y[wb] = acc[l]; /* row owned by this chunk alone, :1204 / :1493 */
            else
                carry[wb] += acc[l]; /* lane slot, :1541 / :1613 */
            acc[l] = 0.0;
            ri++;
        }
        for (int l = 0; l < W; l++) /* :1226-1233 */
            acc[l] = fma(v[i * W + l], x[c[i * W + l]], acc[l]);
    }
    for (int l = 0; l < W; l++) carry[rec[2 * (ri + l) + 1]] += acc[l]; /* :1633-1638 */
    for (int l = 0; l < W; l++) { /* :1640-1649 */
#pragma omp atomic
        y[tail[l]] += carry[l];
    }
}

int cvr_oracle_spmv(int n_chunks, int n_rows,
                    const double* cvr_vals, const int* cvr_cols, const int* record,
                    const int* nnz_rows, const int* final_2, const int* split,
                    const double* x, double* y)
{
    for (int r = 0; r <= n_rows; r++) y[r] = 0.0;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n_chunks; t++)
        spmv_chunk(t, cvr_vals, cvr_cols, record, nnz_rows, final_2, split, x, y);
    return 0;
}

void cvr_oracle_csr_spmv(int n_rows, const double* csr_val, const int* csr_col,
                         const int* row_delim, const double* x, double* y, double* abs_sum)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r <= n_rows; r++) {
        double sum = 0.0, mag = 0.0;
        for (int j = row_delim[r]; j < row_delim[r + 1]; j++) {
            const double p = csr_val[j] * x[csr_col[j]]; /* :1848 */
            sum += p;
            mag += fabs(p);
        }
        y[r] = sum;
        if (abs_sum) abs_sum[r] = mag;
    }
}

/* ------------------------------------------------------------------ ingest */

typedef struct {
    int row, col;
    float val; /* struct Coordinate, :62-66 */
    int seq;   /* file order: makes the qsort below stable like glibc's merge sort */
} entry_t;

static int entry_cmp(const void* a, const void* b)
{
    const entry_t* p = (const entry_t*)a;
    const entry_t* q = (const entry_t*)b;
    if (p->row != q->row) return p->row < q->row ? -1 : 1; /* :136-143 */
    if (p->col != q->col) return p->col < q->col ? -1 : 1;
    return p->seq < q->seq ? -1 : (p->seq > q->seq);
}

int cvr_oracle_read_mtx(const char* path, int ref_last_delim,
                        double** val_out, int** col_out, int** rd_out,
                        int* nnz_padded, int* nnz_file, int* n_rows_out, int* n_cols_out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char* line = NULL;
    size_t cap = 0;
    ssize_t len = getline(&line, &cap, f);
    if (len <= 0 || line[len - 1] != '\n') { /* :337 */
        free(line);
        fclose(f);
        return -2;
    }
    char id[128] = "", object[128] = "", format[128] = "", field[128] = "", symm[128] = "";
    sscanf(line, "%127s %127s %127s %127s %127s", id, object, format, field, symm); /* :344 */
    if (strcmp(object, "matrix") != 0) { free(line); fclose(f); return -2; }  /* :346 */
    if (strcmp(format, "coordinate") != 0) { free(line); fclose(f); return -3; } /* :352 */
    const int pattern = strcmp(field, "pattern") == 0;     /* :358 */
    const int is_complex = strcmp(field, "complex") == 0;  /* :363 */
    const int symmetric = strcmp(symm, "symmetric") == 0;  /* :368 */

    /* comments end at the first line not starting with '%', which is the size line (:377-386).
     * Only newline-terminated lines count: the reference loop tests eof() after getline. */
    int have_size = 0;
    while ((len = getline(&line, &cap, f)) > 0 && line[len - 1] == '\n') {
        if (line[0] != '%') { have_size = 1; break; }
    }
    int n_rows = 0, n_cols = 0, declared = 0;
    if (have_size) sscanf(line, "%d %d %d", &n_rows, &n_cols, &declared);

    long long room = (declared % 16 == 0) ? declared : (declared + 16) / 16 * 16; /* :390 */
    if (symmetric) room *= 2;                                                      /* :396-399 */
    if (room < 16) room = 16;
    entry_t* ent = (entry_t*)malloc(sizeof(entry_t) * (size_t)(room + 16));
    long long n = 0;
    while ((len = getline(&line, &cap, f)) > 0 && line[len - 1] == '\n') { /* :411 */
        if (n + 2 > room) { /* the reference would overrun its buffer; the port grows */
            room *= 2;
            ent = (entry_t*)realloc(ent, sizeof(entry_t) * (size_t)(room + 16));
        }
        entry_t* c = &ent[n];
        c->row = c->col = 0;
        c->val = 0.0f;
        if (pattern) {
            sscanf(line, "%d %d", &c->row, &c->col);
            c->val = (float)(n % 13); /* :417 */
        } else if (is_complex) {
            sscanf(line, "%d %d %f", &c->row, &c->col, &c->val); /* real part, :426 */
        } else {
            sscanf(line, "%d %d %f", &c->row, &c->col, &c->val); /* :432 */
        }
        c->seq = (int)n;
        n++;
        if (symmetric && c->row != c->col) { /* :443-449 */
            ent[n].row = c->col;
            ent[n].col = c->row;
            ent[n].val = c->val;
            ent[n].seq = (int)n;
            n++;
        }
    }
    free(line);
    fclose(f);

    const long long np = (n % 16 == 0) ? n : (n + 16) / 16 * 16; /* :457 */
    if (np + 16 > room + 16) ent = (entry_t*)realloc(ent, sizeof(entry_t) * (size_t)(np + 16));
    for (long long q = n; q < np; q++) { /* :474-482 */
        ent[q].row = n ? ent[n - 1].row : 0;
        ent[q].col = n ? ent[n - 1].col : 0;
        ent[q].val = 0.0f;
        ent[q].seq = (int)q;
    }
    qsort(ent, (size_t)np, sizeof(entry_t), entry_cmp); /* :485 */

    double* val = (double*)malloc(sizeof(double) * (size_t)(np ? np : 1));
    int* col = (int*)malloc(sizeof(int) * (size_t)(np ? np : 1));
    int* rd = (int*)malloc(sizeof(int) * (size_t)(n_rows + 2));
    rd[0] = 0; /* :505 */
    int r = 0;
    long long i;
    for (i = 0; i < np; i++) { /* :511-520 */
        while (ent[i].row != r && r < n_rows + 1) rd[++r] = (int)i;
        val[i] = (double)ent[i].val;
        col[i] = ent[i].col;
    }
    for (int k = r + 1; k <= n_rows + 1; k++) /* :522-526 */
        rd[k] = ref_last_delim ? (int)(i - 1) : (int)i;
    free(ent);

    *val_out = val;
    *col_out = col;
    *rd_out = rd;
    *nnz_padded = (int)np;
    if (nnz_file) *nnz_file = (int)n;
    *n_rows_out = n_rows;
    *n_cols_out = n_cols;
    return 0;
}

void cvr_oracle_free(void* p) { free(p); }
