/* The C-ABI from plain C: fill a hole in a discrete-harmonic field (3 r - 2 c + 7 is its own harmonic extension, so the
 * filled pixels must reproduce it) with sa_laplace_fill on a column-major image, the reference's MatX layout.
 *   gcc -std=c99 -Iinclude examples/fill_c_abi.c -Lsatellite_approximation_b200/lib -lsatfill -Wl,-rpath,$PWD/satellite_approximation_b200/lib -lm
 * Exit code 0 and "max error ..." on success; non-zero with the library's message otherwise. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "satfill.h"

int main(void)
{
    const int64_t rows = 97, cols = 130;
    double* img = (double*)malloc(sizeof(double) * rows * cols);
    uint8_t* mask = (uint8_t*)calloc((size_t)(rows * cols), 1);
    sa_ctx* ctx = NULL;
    sa_options opts;
    sa_stats st;
    int rc;
    double err = 0.0;
    for (int64_t c = 0; c < cols; ++c)
        for (int64_t r = 0; r < rows; ++r) {
            const int hole = r > 10 && r < 80 && c > 20 && c < 100 && (r - 45) * (r - 45) + (c - 60) * (c - 60) < 30 * 30;
            mask[r + c * rows] = (uint8_t)hole;                               /* column-major: row stride 1 */
            img[r + c * rows] = hole ? -1.0 : 3.0 * (double)r - 2.0 * (double)c + 7.0;
        }
    rc = sa_create(&ctx, 0, NULL);
    if (rc != SA_OK) {
        fprintf(stderr, "sa_create failed (%d): no usable CUDA device, and there is no CPU fallback\n", rc);
        return 2;
    }
    sa_default_options(&opts, SA_LAPLACE);
    opts.precond = SA_PRECOND_MULTIGRID;
    opts.tolerance = 1e-12;
    rc = sa_laplace_fill(ctx, &img, 1, mask, rows, cols, /*row_stride*/ 1, /*col_stride*/ rows, &opts, &st);
    if (rc != SA_OK) {
        fprintf(stderr, "sa_laplace_fill: %d %s\n", rc, sa_last_error(ctx));
        return 3;
    }
    for (int64_t c = 0; c < cols; ++c)
        for (int64_t r = 0; r < rows; ++r) {
            const double e = fabs(img[r + c * rows] - (3.0 * (double)r - 2.0 * (double)c + 7.0));
            if (e > err)
                err = e;
        }
    printf("max error %.3e after %lld iterations on %lld unknowns\n", err, (long long)st.iterations, (long long)st.unknowns);
    sa_destroy(ctx);
    free(img);
    free(mask);
    return err < 1e-7 ? 0 : 1;
}
