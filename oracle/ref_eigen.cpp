// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// oracle/_ref/libref_eigen.so: the reference's Laplace / Poisson fill arithmetic executed by the
// reference's OWN vendored Eigen (thirdparty/eigen-master, 3.4.90 snapshot), compiled from the headers
// where they lie under /root/reference (see oracle/Makefile).  The two translation units that hold the
// path (lib/approx/source/laplace.cpp, lib/approx/source/poisson.cpp) cannot be compiled verbatim in
// this image: they include range-v3, spdlog, fmt, OpenCV, Boost.date_time and magic_enum, none of which
// is installed (DESIGN.md "Oracle").  What IS reproduced 1:1 is everything that decides the numbers:
//   * scan order, bounding box, index(), border rule, triplet values and order  (laplace.cpp:33-109)
//   * Eigen::SparseMatrix::setFromTriplets + ConjugateGradient<SparseMatrix<f64>, Lower|Upper> with its
//     default DiagonalPreconditioner, default tolerance (epsilon) and max iterations (2N)
//                                                          (approx/utils.h:15-17, laplace.cpp:108-114)
//   * Poisson numbering, diagonal = in-image neighbour count, -1 couplings, RHS, guess, n/2 iterations
//                                                          (poisson.cpp:162-209, 226-257, 273-283)
// The only additions are knobs the reference lacks on Laplace (tolerance, max iterations) and outputs
// (iterations, error estimate, seconds) so that tests and bench.py can read them.
//
// Layout: every image / mask pointer is a COLUMN-MAJOR rows x cols buffer, i.e. exactly the storage of
// the reference's utils::MatX<T> (lib/utils/include/utils/types.h:31).

#include <Eigen/Sparse>

#include <chrono>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

using f64 = double;
using Index = Eigen::Index;
using triplet_t = Eigen::Triplet<f64>;                                         // approx/utils.h:15
using sparse_t = Eigen::SparseMatrix<f64>;                                     // approx/utils.h:16
using SparseSolver = Eigen::ConjugateGradient<sparse_t, Eigen::Lower | Eigen::Upper>;  // approx/utils.h:17
using MapF = Eigen::Map<Eigen::Matrix<f64, Eigen::Dynamic, Eigen::Dynamic>>;
using MapCF = Eigen::Map<const Eigen::Matrix<f64, Eigen::Dynamic, Eigen::Dynamic>>;
using MapB = Eigen::Map<const Eigen::Matrix<uint8_t, Eigen::Dynamic, Eigen::Dynamic>>;

double now_s()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

struct px {
    Index row, col;
};

// approx/utils.h:35-50 -- neighbour order (-1,0) (+1,0) (0,-1) (0,+1), out-of-image entries removed.
int in_image_neighbours(Index rows, Index cols, Index r, Index c, px out[4])
{
    static const int dr[4] = { -1, 1, 0, 0 };
    static const int dc[4] = { 0, 0, -1, 1 };
    int n = 0;
    for (int k = 0; k < 4; ++k) {
        Index rr = r + dr[k], cc = c + dc[k];
        if (rr >= 0 && rr < rows && cc >= 0 && cc < cols)
            out[n++] = { rr, cc };
    }
    return n;
}

}  // namespace

extern "C" {

// Status codes shared with the plain-C oracle.
enum { REF_OK = 0, REF_EMPTY = 1, REF_NOT_CONVERGED = 2, REF_BAD_ARG = 3 };

void ref_set_threads(int n) { Eigen::setNbThreads(n); }  // executables/poisson-main.cpp:35-37

// laplace.cpp:31-120 (solve_matrix).  tol <= 0 -> Eigen default (epsilon); max_it <= 0 -> Eigen default (2N).
int ref_laplace_fill(double* img, const uint8_t* mask, int64_t rows, int64_t cols, double tol, int64_t max_it,
    int64_t* out_iters, double* out_error, double* out_assemble_s, double* out_solve_s, int64_t* out_system_size)
{
    MapF input(img, rows, cols);
    MapB invalid_mask(mask, rows, cols);
    double t0 = now_s();

    std::vector<px> invalid_pixels;  // laplace.cpp:33-40, row outer / col inner
    for (Index row = 0; row < rows; ++row)
        for (Index col = 0; col < cols; ++col)
            if (invalid_mask(row, col))
                invalid_pixels.push_back({ row, col });
    if (invalid_pixels.empty())  // laplace.cpp:41-44
        return REF_EMPTY;

    Index min_row = rows, max_row = -1, min_col = cols, max_col = -1;  // laplace.cpp:46-47
    for (auto const& p : invalid_pixels) {
        min_row = std::min(min_row, p.row);
        max_row = std::max(max_row, p.row);
        min_col = std::min(min_col, p.col);
        max_col = std::max(max_col, p.col);
    }
    Index height = (max_row - min_row) + 1;  // laplace.cpp:49-52
    Index width = (max_col - min_col) + 1;
    Index matrix_size = height * width;
    auto index = [&](Index row, Index col) { return (col - min_col) + (row - min_row) * width; };  // :54-56

    Eigen::VectorXd b(matrix_size);
    b.setZero();
    std::vector<triplet_t> coefficients;

    auto identity_row = [&](Index row, Index col) {  // laplace.cpp:63-69
        Index i = index(row, col);
        coefficients.emplace_back(i, i, 1.0);
        b[i] = input(row, col);
    };
    auto coupling = [&](Index row, Index col, int dr, int dc, f64 v) {  // laplace.cpp:71-85
        Index i = index(row, col);
        Index r2 = row + dr, c2 = col + dc;
        if (!invalid_mask(r2, c2)) {
            b[i] -= v * input(r2, c2);
            return;
        }
        coefficients.emplace_back(i, index(r2, c2), v);
    };
    for (Index row = min_row; row <= max_row; ++row) {      // laplace.cpp:96-106 (cartesian product,
        for (Index col = min_col; col <= max_col; ++col) {  // row-major order)
            bool border = row == 0 || row == rows - 1 || col == 0 || col == cols - 1;  // laplace.cpp:23-29
            if (border || !invalid_mask(row, col)) {
                identity_row(row, col);
            } else {  // laplace.cpp:87-94
                coupling(row, col, -1, 0, 1.0);
                coupling(row, col, +1, 0, 1.0);
                coupling(row, col, 0, -1, 1.0);
                coupling(row, col, 0, +1, 1.0);
                coupling(row, col, 0, 0, -4.0);
            }
        }
    }
    sparse_t A(matrix_size, matrix_size);  // laplace.cpp:108-109
    A.setFromTriplets(coefficients.begin(), coefficients.end());
    double t1 = now_s();

    SparseSolver solver(A);  // laplace.cpp:113
    if (tol > 0)
        solver.setTolerance(tol);
    if (max_it > 0)
        solver.setMaxIterations(max_it);
    Eigen::VectorXd values = solver.solve(b);  // laplace.cpp:114
    double t2 = now_s();

    for (auto const& p : invalid_pixels)  // laplace.cpp:117-119
        input(p.row, p.col) = values[index(p.row, p.col)];

    if (out_iters) *out_iters = solver.iterations();
    if (out_error) *out_error = solver.error();
    if (out_assemble_s) *out_assemble_s = t1 - t0;
    if (out_solve_s) *out_solve_s = t2 - t1;
    if (out_system_size) *out_system_size = matrix_size;
    return solver.info() == Eigen::Success ? REF_OK : REF_NOT_CONVERGED;
}

// poisson.cpp:145-290 (mask overload).  max_it < 0 -> reference default n/2 (poisson.cpp:207).
// inputs[c] are modified in place only when every band converged (poisson.cpp:263-269 returns early).
// out_iters / out_error / out_solve_s are per band (nbands entries each, may be null).
int ref_poisson_blend(double* const* inputs, const double* const* replacements, int nbands, const uint8_t* mask,
    int64_t rows, int64_t cols, double tol, int64_t max_it, int64_t* out_iters, double* out_error, double* out_solve_s,
    double* out_setup_s, int64_t* out_unknowns)
{
    MapB invalid_mask(mask, rows, cols);
    double t0 = now_s();
    auto flatten = [&](Index row, Index col) { return col + row * cols; };  // poisson.cpp:162-164

    // poisson.cpp:167-177: the reference uses unordered_map<flat,int>; a dense table holds the same map.
    std::vector<int> variable_numbers(size_t(rows) * size_t(cols), -1);
    int n = 0;
    for (Index row = 0; row < rows; ++row)
        for (Index col = 0; col < cols; ++col)
            if (invalid_mask(row, col))
                variable_numbers[flatten(row, col)] = n++;
    if (out_unknowns) *out_unknowns = n;

    std::vector<triplet_t> triplets;  // poisson.cpp:179-200
    int irow = 0;
    px nb[4];
    for (Index row = 0; row < rows; ++row) {
        for (Index col = 0; col < cols; ++col) {
            if (!invalid_mask(row, col))
                continue;
            int cnt = in_image_neighbours(rows, cols, row, col, nb);
            triplets.emplace_back(irow, variable_numbers[flatten(row, col)], (f64)cnt);
            for (int k = 0; k < cnt; ++k)
                if (invalid_mask(nb[k].row, nb[k].col))
                    triplets.emplace_back(irow, variable_numbers[flatten(nb[k].row, nb[k].col)], -1);
            irow += 1;
        }
    }
    sparse_t A(n, n);  // poisson.cpp:203-205
    A.setFromTriplets(triplets.begin(), triplets.end());
    SparseSolver solver(A);
    long max_iters = max_it >= 0 ? long(max_it) : long(A.cols() / 2);  // poisson.cpp:207-209
    solver.setMaxIterations(max_iters);
    solver.setTolerance(tol);
    if (out_setup_s) *out_setup_s = now_s() - t0;

    std::vector<Eigen::VectorXd> solutions;
    for (int c = 0; c < nbands; ++c) {  // poisson.cpp:226-270
        MapCF f(inputs[c], rows, cols);
        MapCF g(replacements[c], rows, cols);
        Eigen::VectorXd b(n), guess(n);
        b.setZero();
        irow = 0;
        for (Index row = 0; row < rows; ++row) {
            for (Index col = 0; col < cols; ++col) {
                if (!invalid_mask(row, col))
                    continue;
                guess(irow) = g(row, col);  // poisson.cpp:239
                int cnt = in_image_neighbours(rows, cols, row, col, nb);
                for (int k = 0; k < cnt; ++k) {  // poisson.cpp:241-251
                    b(irow) += (g(row, col) - g(nb[k].row, nb[k].col));
                    if (!invalid_mask(nb[k].row, nb[k].col))
                        b(irow) += f(nb[k].row, nb[k].col);
                }
                irow += 1;
            }
        }
        double s0 = now_s();
        solutions.emplace_back(solver.solveWithGuess(b, guess));  // poisson.cpp:257
        if (out_solve_s) out_solve_s[c] = now_s() - s0;
        if (out_error) out_error[c] = solver.error();
        if (out_iters) out_iters[c] = solver.iterations();
        if (solver.info() != Eigen::Success)  // poisson.cpp:263-269: abort, nothing written
            return REF_NOT_CONVERGED;
    }
    for (int c = 0; c < nbands; ++c) {  // poisson.cpp:273-283
        MapF out(inputs[c], rows, cols);
        for (Index row = 0; row < rows; ++row)
            for (Index col = 0; col < cols; ++col)
                if (invalid_mask(row, col))
                    out(row, col) = solutions[c](variable_numbers[flatten(row, col)]);
    }
    return REF_OK;
}

}  // extern "C"
