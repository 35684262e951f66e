/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the Laplace / Poisson fill path of ebiederstadt/satellite-approximation
 * (lib/approx).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / the timed CPU baseline -- the product (libsatfill.so)
 * never links or calls it.
 *
 * PARITY PIN: the reference ships no golden vectors for filled pixel values.  This port
 * is pinned instead against oracle/_ref/libref_eigen.so -- the same assembly executed by the reference's own
 * vendored Eigen ConjugateGradient -- by tests/test_oracle.py (live when _ref is built, and through the
 * fixtures in tests/golden/ that oracle/make_golden.py generated from _ref), and against the two integer
 * known-answer tests the reference holds (tests/approximation.h:9-33 valid_neighbours, :55-75 components).
 *
 * Every function cites the reference lines it follows.  All image pointers take explicit element strides
 * (rs = row stride, cs = column stride) because the reference's MatX is column-major (utils/types.h:31)
 * while every integer output is defined in row-major raster order (laplace.cpp:34-40, poisson.cpp:169-176).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { SO_OK = 0, SO_EMPTY = 1, SO_NOT_CONVERGED = 2, SO_BAD_ARG = 3 };

#define AT(p, r, c) ((p)[(int64_t)(r) * rs + (int64_t)(c) * cs])

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* approx/utils.h:29-50 (within_bounds + valid_neighbours): candidates in the order (-1,0) (+1,0) (0,-1)
 * (0,+1), those outside the image removed.  out_rc receives (row, col) pairs; returns how many (0..4).
 * tests/approximation.h:9-33 pins the counts, including 0 for the out-of-range index {100,100} of a
 * 100x100 image. */
int so_valid_neighbours(int64_t rows, int64_t cols, int64_t r, int64_t c, int64_t* out_rc)
{
    static const int dr[4] = { -1, 1, 0, 0 };
    static const int dc[4] = { 0, 0, -1, 1 };
    int n = 0;
    for (int k = 0; k < 4; ++k) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr >= 0 && rr < rows && cc >= 0 && cc < cols) {
            if (out_rc) {
                out_rc[2 * n] = rr;
                out_rc[2 * n + 1] = cc;
            }
            ++n;
        }
    }
    return n;
}

/* laplace.cpp:33-52: row-major list of invalid pixels and its bounding box.
 * out_pixels (may be NULL) receives (row, col) int64 pairs; bbox = {min_row, max_row, min_col, max_col}.
 * Returns the number of invalid pixels. */
int64_t so_mask_scan(const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int64_t* out_pixels,
    int64_t* bbox)
{
    int64_t n = 0;
    int64_t min_r = rows, max_r = -1, min_c = cols, max_c = -1;
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t c = 0; c < cols; ++c) {
            if (!AT(mask, r, c))
                continue;
            if (out_pixels) {
                out_pixels[2 * n] = r;
                out_pixels[2 * n + 1] = c;
            }
            ++n;
            if (r < min_r) min_r = r;
            if (r > max_r) max_r = r;
            if (c < min_c) min_c = c;
            if (c > max_c) max_c = c;
        }
    }
    if (bbox) {
        bbox[0] = min_r;
        bbox[1] = max_r;
        bbox[2] = min_c;
        bbox[3] = max_c;
    }
    return n;
}

/* poisson.cpp:162-177: numbering[col + row*cols] = number of invalid pixels strictly before (row, col) in
 * row-major order, or -1 for valid pixels (the reference's unordered_map has no entry there).  The output
 * table is always row-major dense regardless of the mask's strides.  Returns n (the reference stores it as
 * a 32-bit int, poisson.cpp:177). */
int64_t so_unknown_numbering(const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int32_t* numbering)
{
    int64_t n = 0;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c)
            numbering[c + r * cols] = AT(mask, r, c) ? (int32_t)(n++) : -1;
    return n;
}

/* approx/laplace.h:11-20 + tests/approximation.h:55-75.  The reference declares find_connected_components
 * but never defines it, so this states the contract the build adopts:
 * 4-connectivity (the stencil's coupling, utils.h:38-44), background 0, labels 1..K numbered by the first
 * pixel of each component in row-major raster order.  labels is a row-major dense rows x cols table.
 * sizes (may be NULL, capacity K) receives per-label pixel counts.  Returns K. */
int32_t so_label_components(const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int32_t* labels)
{
    int64_t total = rows * cols;
    memset(labels, 0, (size_t)total * sizeof(int32_t));
    int64_t* stack = (int64_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof(int64_t));
    int32_t K = 0;
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t c = 0; c < cols; ++c) {
            if (!AT(mask, r, c) || labels[c + r * cols])
                continue;
            ++K;
            int64_t top = 0;
            stack[top++] = c + r * cols;
            labels[c + r * cols] = K;
            while (top) {
                int64_t f = stack[--top];
                int64_t fr = f / cols, fc = f % cols;
                int64_t nb[8];
                int cnt = so_valid_neighbours(rows, cols, fr, fc, nb);
                for (int k = 0; k < cnt; ++k) {
                    int64_t rr = nb[2 * k], cc = nb[2 * k + 1];
                    if (AT(mask, rr, cc) && !labels[cc + rr * cols]) {
                        labels[cc + rr * cols] = K;
                        stack[top++] = cc + rr * cols;
                    }
                }
            }
        }
    }
    free(stack);
    return K;
}

/* ---------------------------------------------------------------------------------------------------------
 * Sparse matrix in CSR built from (row, col, value) triplets listed row by row -- what
 * Eigen::SparseMatrix::setFromTriplets produces for these assemblies (no duplicate entries are ever
 * emitted by laplace.cpp:58-106 / poisson.cpp:179-200, so there is nothing to sum).  Within a row Eigen
 * sorts by column index; the row-times-vector sums below therefore run in ascending column order.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;
    int64_t nnz;
    int64_t* rowptr;
    int32_t* col;
    double* val;
    double* invdiag;
} csr_t;

static void csr_free(csr_t* A)
{
    free(A->rowptr);
    free(A->col);
    free(A->val);
    free(A->invdiag);
    memset(A, 0, sizeof(*A));
}

static void csr_sort_rows_and_diag(csr_t* A)
{
    /* insertion sort per row (rows hold at most 5 entries) + Eigen's DiagonalPreconditioner::factorize
     * (BasicPreconditioners.h:63-75): invdiag = 1/diag, or 1 when the diagonal is missing or zero. */
    for (int64_t i = 0; i < A->n; ++i) {
        int64_t a = A->rowptr[i], b = A->rowptr[i + 1];
        for (int64_t k = a + 1; k < b; ++k) {
            int32_t cj = A->col[k];
            double vj = A->val[k];
            int64_t m = k - 1;
            while (m >= a && A->col[m] > cj) {
                A->col[m + 1] = A->col[m];
                A->val[m + 1] = A->val[m];
                --m;
            }
            A->col[m + 1] = cj;
            A->val[m + 1] = vj;
        }
        double d = 0.0;
        for (int64_t k = a; k < b; ++k)
            if (A->col[k] == i)
                d = A->val[k];
        A->invdiag[i] = d != 0.0 ? 1.0 / d : 1.0;
    }
}

static void csr_matvec(const csr_t* A, const double* x, double* y)
{
    /* SparseDenseProduct.h:34-82, row-major product.  NOTE the reference instantiates
     * ConjugateGradient<ColMajor SparseMatrix, Lower|Upper>, which multiplies by the TRANSPOSE view
     * (ConjugateGradient.h:193-208); for the symmetric matrices this oracle accepts that is the same
     * product (SURVEY.md F5 covers the non-symmetric border case, which the contract excludes). */
    for (int64_t i = 0; i < A->n; ++i) {
        double s = 0.0;
        for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k)
            s += A->val[k] * x[A->col[k]];
        y[i] = s;
    }
}

static double vdot(const double* a, const double* b, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i)
        s += a[i] * b[i];
    return s;
}

/* Eigen conjugate_gradient, ConjugateGradient.h:30-85, statement for statement: zero-RHS shortcut (:43-49),
 * threshold max(tol^2 |b|^2, DBL_MIN) (:50-51), early-out when the guess already meets it (:52-57), Jacobi
 * preconditioner z = invdiag .* r (BasicPreconditioners.h:79-82), strict "<" stop test after the residual
 * update (:72-73), iteration counter incremented at the end of the body (:81).
 * Returns the iteration count; *err receives sqrt(|r|^2/|b|^2). */
static int64_t eigen_style_pcg(const csr_t* A, const double* b, double* x, double tol, int64_t max_iters, double* err)
{
    int64_t n = A->n;
    double* r = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double* p = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double* z = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double* tmp = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    int64_t it = 0;

    csr_matvec(A, x, tmp);
    for (int64_t i = 0; i < n; ++i)
        r[i] = b[i] - tmp[i];
    double rhs2 = vdot(b, b, n);
    if (rhs2 == 0.0) {
        memset(x, 0, (size_t)n * sizeof(double));
        *err = 0.0;
        goto done;
    }
    {
        double threshold = tol * tol * rhs2;
        if (threshold < DBL_MIN)
            threshold = DBL_MIN;
        double r2 = vdot(r, r, n);
        if (r2 < threshold) {
            *err = sqrt(r2 / rhs2);
            goto done;
        }
        for (int64_t i = 0; i < n; ++i)
            p[i] = A->invdiag[i] * r[i];
        double abs_new = vdot(r, p, n);
        while (it < max_iters) {
            csr_matvec(A, p, tmp);
            double alpha = abs_new / vdot(p, tmp, n);
            for (int64_t i = 0; i < n; ++i)
                x[i] += alpha * p[i];
            for (int64_t i = 0; i < n; ++i)
                r[i] -= alpha * tmp[i];
            r2 = vdot(r, r, n);
            if (r2 < threshold)
                break;
            for (int64_t i = 0; i < n; ++i)
                z[i] = A->invdiag[i] * r[i];
            double abs_old = abs_new;
            abs_new = vdot(r, z, n);
            double beta = abs_new / abs_old;
            for (int64_t i = 0; i < n; ++i)
                p[i] = z[i] + beta * p[i];
            ++it;
        }
        *err = sqrt(r2 / rhs2);
    }
done:
    free(r);
    free(p);
    free(z);
    free(tmp);
    return it;
}

typedef struct {
    int64_t unknowns;     /* number of invalid pixels                                  */
    int64_t system_size;  /* rows of the linear system handed to CG                    */
    int64_t iterations;   /* CG iterations (last band for Poisson, like PerfInfo)      */
    double error;         /* Eigen's error estimate sqrt(|r|^2/|b|^2)                  */
    double assemble_s;
    double solve_s;
} so_stats;

/* laplace.cpp:31-120 (solve_matrix) + :122-132.
 * mode 0 = FAITHFUL: the bounding-box system exactly as assembled by the reference -- identity rows for
 *          known and image-border cells, (-4, +1) rows for interior invalid cells (laplace.cpp:58-106).
 *          Indefinite but symmetric when the mask does not touch the image border.
 * mode 1 = REDUCED: the same equations with the identity rows eliminated: unknowns are the invalid pixels
 *          that are not on the image border, 4 x_p - sum_{q invalid, interior} x_q = sum_{q known or border} f_q.
 *          This is the "solve the as-assembled A x = b exactly" semantics the build adopts for masks that
 *          touch the border: border pixels are Dirichlet data and come back unchanged.
 * tol <= 0 -> DBL_EPSILON (Eigen default, IterativeSolverBase.h:367-368); max_it <= 0 -> 2N (:251). */
int so_laplace_fill(double* img, const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int mode,
    double tol, int64_t max_it, so_stats* st)
{
    so_stats local;
    memset(&local, 0, sizeof(local));
    double t0 = now_s();
    int64_t bbox[4];
    int64_t n_inv = so_mask_scan(mask, rows, cols, rs, cs, NULL, bbox);
    local.unknowns = n_inv;
    if (n_inv == 0) { /* laplace.cpp:41-44 */
        if (st) *st = local;
        return SO_EMPTY;
    }
    int64_t min_r = bbox[0], max_r = bbox[1], min_c = bbox[2], max_c = bbox[3];
    int64_t height = max_r - min_r + 1, width = max_c - min_c + 1;

    csr_t A;
    memset(&A, 0, sizeof(A));
    double *b, *x;
    int32_t* number = NULL; /* mode 1: bbox cell -> unknown id */
    int64_t N;
#define BORDER(r, c) ((r) == 0 || (r) == rows - 1 || (c) == 0 || (c) == cols - 1) /* laplace.cpp:23-29 */
#define IDX(r, c) (((c)-min_c) + ((r)-min_r) * width)                              /* laplace.cpp:54-56 */
    if (mode == 0) {
        N = height * width;
        A.n = N;
        A.rowptr = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
        A.col = (int32_t*)malloc((size_t)N * 5 * sizeof(int32_t));
        A.val = (double*)malloc((size_t)N * 5 * sizeof(double));
        A.invdiag = (double*)malloc((size_t)N * sizeof(double));
        b = (double*)calloc((size_t)N, sizeof(double));
        x = (double*)calloc((size_t)N, sizeof(double)); /* solve(): x0 = 0 (IterativeSolverBase.h:357-360) */
        int64_t nnz = 0;
        static const int dr[5] = { -1, 1, 0, 0, 0 };
        static const int dc[5] = { 0, 0, -1, 1, 0 };
        static const double dv[5] = { 1.0, 1.0, 1.0, 1.0, -4.0 }; /* laplace.cpp:87-94 */
        for (int64_t r = min_r; r <= max_r; ++r) {
            for (int64_t c = min_c; c <= max_c; ++c) { /* laplace.cpp:96-106 */
                int64_t i = IDX(r, c);
                A.rowptr[i] = nnz;
                if (BORDER(r, c) || !AT(mask, r, c)) { /* laplace.cpp:63-69 */
                    A.col[nnz] = (int32_t)i;
                    A.val[nnz++] = 1.0;
                    b[i] = AT(img, r, c);
                } else {
                    for (int k = 0; k < 5; ++k) { /* laplace.cpp:71-85 */
                        int64_t r2 = r + dr[k], c2 = c + dc[k];
                        if (!AT(mask, r2, c2)) {
                            b[i] -= dv[k] * AT(img, r2, c2);
                        } else {
                            A.col[nnz] = (int32_t)IDX(r2, c2);
                            A.val[nnz++] = dv[k];
                        }
                    }
                }
            }
        }
        A.rowptr[N] = nnz;
        A.nnz = nnz;
    } else {
        number = (int32_t*)malloc((size_t)(height * width) * sizeof(int32_t));
        N = 0;
        for (int64_t r = min_r; r <= max_r; ++r)
            for (int64_t c = min_c; c <= max_c; ++c)
                number[IDX(r, c)] = (AT(mask, r, c) && !BORDER(r, c)) ? (int32_t)(N++) : -1;
        A.n = N;
        A.rowptr = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
        A.col = (int32_t*)malloc((size_t)(N > 0 ? N : 1) * 5 * sizeof(int32_t));
        A.val = (double*)malloc((size_t)(N > 0 ? N : 1) * 5 * sizeof(double));
        A.invdiag = (double*)malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
        b = (double*)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
        x = (double*)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
        int64_t nnz = 0;
        static const int dr[4] = { -1, 1, 0, 0 };
        static const int dc[4] = { 0, 0, -1, 1 };
        for (int64_t r = min_r; r <= max_r; ++r) {
            for (int64_t c = min_c; c <= max_c; ++c) {
                int32_t i = number[IDX(r, c)];
                if (i < 0)
                    continue;
                A.rowptr[i] = nnz;
                A.col[nnz] = i;
                A.val[nnz++] = 4.0;
                for (int k = 0; k < 4; ++k) {
                    int64_t r2 = r + dr[k], c2 = c + dc[k]; /* interior pixel: always inside the image */
                    int inside_bbox = r2 >= min_r && r2 <= max_r && c2 >= min_c && c2 <= max_c;
                    int32_t j = inside_bbox ? number[IDX(r2, c2)] : -1;
                    if (j >= 0) {
                        A.col[nnz] = j;
                        A.val[nnz++] = -1.0;
                    } else {
                        b[i] += AT(img, r2, c2);
                    }
                }
            }
        }
        A.rowptr[N] = nnz;
        A.nnz = nnz;
    }
    csr_sort_rows_and_diag(&A);
    double t1 = now_s();

    double eff_tol = tol > 0 ? tol : DBL_EPSILON;
    int64_t eff_max = max_it > 0 ? max_it : 2 * N;
    double err = 0.0;
    int64_t it = eigen_style_pcg(&A, b, x, eff_tol, eff_max, &err);
    double t2 = now_s();

    /* laplace.cpp:117-119: write back into invalid pixels only */
    for (int64_t r = min_r; r <= max_r; ++r) {
        for (int64_t c = min_c; c <= max_c; ++c) {
            if (!AT(mask, r, c))
                continue;
            if (mode == 0) {
                AT(img, r, c) = x[IDX(r, c)];
            } else if (number[IDX(r, c)] >= 0) {
                AT(img, r, c) = x[number[IDX(r, c)]];
            }
        }
    }
#undef BORDER
#undef IDX
    local.system_size = N;
    local.iterations = it;
    local.error = err;
    local.assemble_s = t1 - t0;
    local.solve_s = t2 - t1;
    if (st) *st = local;
    free(b);
    free(x);
    free(number);
    csr_free(&A);
    return err <= eff_tol ? SO_OK : SO_NOT_CONVERGED; /* ConjugateGradient.h:209 */
}

/* poisson.cpp:145-290 (mask overload).  inputs[c] / replacements[c] are nbands images sharing rows, cols and
 * strides.  max_it < 0 -> n/2 (poisson.cpp:207).  Like the reference, a band that fails to converge aborts the
 * call and NO band is written (poisson.cpp:263-269).  per_band (may be NULL) receives nbands so_stats. */
int so_poisson_blend(double* const* inputs, const double* const* replacements, int nbands, const uint8_t* mask,
    int64_t rows, int64_t cols, int64_t rs, int64_t cs, double tol, int64_t max_it, so_stats* per_band)
{
    double t0 = now_s();
    int32_t* number = (int32_t*)malloc((size_t)(rows * cols > 0 ? rows * cols : 1) * sizeof(int32_t));
    int64_t n = so_unknown_numbering(mask, rows, cols, rs, cs, number); /* poisson.cpp:162-177 */

    csr_t A;
    memset(&A, 0, sizeof(A));
    A.n = n;
    A.rowptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    A.col = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * 5 * sizeof(int32_t));
    A.val = (double*)malloc((size_t)(n > 0 ? n : 1) * 5 * sizeof(double));
    A.invdiag = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    int64_t nnz = 0;
    int64_t nb[8];
    for (int64_t r = 0; r < rows; ++r) { /* poisson.cpp:179-200 */
        for (int64_t c = 0; c < cols; ++c) {
            int32_t i = number[c + r * cols];
            if (i < 0)
                continue;
            int cnt = so_valid_neighbours(rows, cols, r, c, nb);
            A.rowptr[i] = nnz;
            A.col[nnz] = i;
            A.val[nnz++] = (double)cnt;
            for (int k = 0; k < cnt; ++k) {
                int32_t j = number[nb[2 * k + 1] + nb[2 * k] * cols];
                if (j >= 0) {
                    A.col[nnz] = j;
                    A.val[nnz++] = -1.0;
                }
            }
        }
    }
    A.rowptr[n] = nnz;
    A.nnz = nnz;
    csr_sort_rows_and_diag(&A);
    int64_t eff_max = max_it >= 0 ? max_it : n / 2; /* poisson.cpp:207 */
    double setup_s = now_s() - t0;

    double* b = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double** sol = (double**)calloc((size_t)(nbands > 0 ? nbands : 1), sizeof(double*));
    int status = SO_OK;
    for (int cidx = 0; cidx < nbands; ++cidx) { /* poisson.cpp:226-270 */
        const double* f = inputs[cidx];
        const double* g = replacements[cidx];
        double* x = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
        sol[cidx] = x;
        double tb = now_s();
        for (int64_t r = 0; r < rows; ++r) {
            for (int64_t c = 0; c < cols; ++c) {
                int32_t i = number[c + r * cols];
                if (i < 0)
                    continue;
                x[i] = AT(g, r, c); /* poisson.cpp:239: guess = replacement */
                double acc = 0.0;
                int cnt = so_valid_neighbours(rows, cols, r, c, nb);
                for (int k = 0; k < cnt; ++k) { /* poisson.cpp:241-251 */
                    int64_t r2 = nb[2 * k], c2 = nb[2 * k + 1];
                    acc += (AT(g, r, c) - AT(g, r2, c2));
                    if (!AT(mask, r2, c2))
                        acc += AT(f, r2, c2);
                }
                b[i] = acc;
            }
        }
        double ts = now_s();
        double err = 0.0;
        int64_t it = eigen_style_pcg(&A, b, x, tol, eff_max, &err); /* poisson.cpp:257 */
        double te = now_s();
        if (per_band) {
            per_band[cidx].unknowns = n;
            per_band[cidx].system_size = n;
            per_band[cidx].iterations = it;
            per_band[cidx].error = err;
            per_band[cidx].assemble_s = (cidx == 0 ? setup_s : 0.0) + (ts - tb);
            per_band[cidx].solve_s = te - ts;
        }
        if (!(err <= tol)) { /* poisson.cpp:263-269 */
            status = SO_NOT_CONVERGED;
            break;
        }
    }
    if (status == SO_OK) { /* poisson.cpp:273-283 */
        for (int cidx = 0; cidx < nbands; ++cidx) {
            double* out = inputs[cidx];
            for (int64_t r = 0; r < rows; ++r)
                for (int64_t c = 0; c < cols; ++c)
                    if (number[c + r * cols] >= 0)
                        AT(out, r, c) = sol[cidx][number[c + r * cols]];
        }
    }
    for (int cidx = 0; cidx < nbands; ++cidx)
        free(sol[cidx]);
    free(sol);
    free(b);
    free(number);
    csr_free(&A);
    return status;
}

/* Reduced-system residual of a filled image, used by tests and bench.py to verify a stop rule without
 * trusting the solver under test: for the unknown set U (invalid pixels; for laplace != 0 minus the image
 * border) returns |b_U - A_UU x|_2 / |b_U|_2 with
 *   laplace != 0:  row p:  sum_{q in N4(p)} u_q - 4 u_p          , b_p = sum_{q known} u_q   (Appendix A)
 *   laplace == 0:  Poisson row p with guidance g (poisson.cpp:187-196, 241-251).
 * The residual is evaluated directly from the image u (known neighbours are read from u itself). */
double so_relative_residual(const double* u, const double* g, const uint8_t* mask, int64_t rows, int64_t cols,
    int64_t rs, int64_t cs, int laplace)
{
    double r2 = 0.0, b2 = 0.0;
    int64_t nb[8];
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t c = 0; c < cols; ++c) {
            if (!AT(mask, r, c))
                continue;
            int border = r == 0 || r == rows - 1 || c == 0 || c == cols - 1;
            if (laplace && border)
                continue;
            int cnt = so_valid_neighbours(rows, cols, r, c, nb);
            double bp = 0.0, ax = (double)cnt * AT(u, r, c);
            for (int k = 0; k < cnt; ++k) {
                int64_t r2i = nb[2 * k], c2i = nb[2 * k + 1];
                int nb_border = r2i == 0 || r2i == rows - 1 || c2i == 0 || c2i == cols - 1;
                int unknown = AT(mask, r2i, c2i) && !(laplace && nb_border);
                if (!laplace)
                    bp += AT(g, r, c) - AT(g, r2i, c2i);
                if (unknown)
                    ax -= AT(u, r2i, c2i);
                else
                    bp += AT(u, r2i, c2i);
            }
            r2 += (bp - ax) * (bp - ax);
            b2 += bp * bp;
        }
    }
    return b2 > 0.0 ? sqrt(r2 / b2) : sqrt(r2);
}
