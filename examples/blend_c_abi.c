/* The C-ABI from plain C, Poisson side and integer side.  Known answers that follow from the reference's equations
 *  and from its own tests:
 *   - blend_images_poisson with replacement = input + constant returns the input (the guidance field of a constant offset
 *     is the input's own gradient field; poisson.cpp:226-254), two bands, row-major images;
 *   - the unknown numbering is the raster-order rank of each invalid pixel (poisson.cpp:162-177);
 *   - the connected-components case of tests/approximation.h:55-75: a 2x2 block at (1,1) and a 4x2 block at (5,5) in a
 *     10x10 mask give labels 1 (4 pixels) and 2 (8 pixels).
 *   gcc -std=c99 -Iinclude examples/blend_c_abi.c -Lsatellite_approximation_b200/lib -lsatfill -Wl,-rpath,$PWD/satellite_approximation_b200/lib -lm
 * Exit code 0 on success; 2 when there is no CUDA device (there is no CPU fallback); other non-zero on a wrong answer. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "satfill.h"

#define ROWS 64
#define COLS 83

static double field(int band, int64_t r, int64_t c)
{
    return band == 0 ? 100.0 + 40.0 * sin(0.11 * (double)r) * cos(0.07 * (double)c)
                     : 500.0 + 0.5 * (double)r * (double)c / 10.0 - 3.0 * (double)c;
}

int main(void)
{
    static double in0[ROWS * COLS], in1[ROWS * COLS], rp0[ROWS * COLS], rp1[ROWS * COLS];
    static uint8_t mask[ROWS * COLS];
    static int32_t numbering[ROWS * COLS];
    double* inputs[2];
    const double* replacements[2];
    sa_ctx* ctx = NULL;
    sa_options opts;
    sa_stats st[2];
    int64_t n = 0, expect_n = 0, k = 0;
    double err = 0.0;
    int rc;

    for (int64_t r = 0; r < ROWS; ++r)
        for (int64_t c = 0; c < COLS; ++c) {
            const int hole = (r - 30) * (r - 30) + (c - 40) * (c - 40) < 20 * 20 || (r > 50 && r < 60 && c > 5 && c < 30);
            mask[c + r * COLS] = (uint8_t)hole;
            in0[c + r * COLS] = hole ? -999.0 : field(0, r, c); /* what lies under the mask must not matter */
            in1[c + r * COLS] = hole ? 12345.0 : field(1, r, c);
            rp0[c + r * COLS] = field(0, r, c) + 17.0;
            rp1[c + r * COLS] = field(1, r, c) - 250.0;
            expect_n += hole;
        }
    inputs[0] = in0, inputs[1] = in1;
    replacements[0] = rp0, replacements[1] = rp1;

    rc = sa_create(&ctx, 0, NULL);
    if (rc != SA_OK) {
        fprintf(stderr, "sa_create failed (%d): no usable CUDA device, and there is no CPU fallback\n", rc);
        return 2;
    }

    /* integer path: raster-order numbering */
    rc = sa_unknown_numbering(ctx, mask, ROWS, COLS, /*row_stride*/ COLS, /*col_stride*/ 1, numbering, &n);
    if (rc != SA_OK || n != expect_n) {
        fprintf(stderr, "sa_unknown_numbering: rc %d, n %lld (expected %lld) %s\n", rc, (long long)n, (long long)expect_n,
            sa_last_error(ctx));
        return 3;
    }
    for (int64_t i = 0; i < ROWS * COLS; ++i)
        if (numbering[i] != (mask[i] ? (int32_t)k++ : -1)) {
            fprintf(stderr, "numbering[%lld] = %d\n", (long long)i, numbering[i]);
            return 4;
        }

    /* Poisson blend: replacement = input + constant  =>  the input comes back */
    sa_default_options(&opts, SA_POISSON);
    opts.tolerance = 1e-12;
    opts.max_iterations = 100000;
    rc = sa_poisson_blend(ctx, inputs, replacements, 2, mask, ROWS, COLS, COLS, 1, &opts, st);
    if (rc != SA_OK) {
        fprintf(stderr, "sa_poisson_blend: %d %s\n", rc, sa_last_error(ctx));
        return 5;
    }
    for (int64_t r = 0; r < ROWS; ++r)
        for (int64_t c = 0; c < COLS; ++c) {
            const double e0 = fabs(in0[c + r * COLS] - field(0, r, c)) / 140.0;
            const double e1 = fabs(in1[c + r * COLS] - field(1, r, c)) / 800.0;
            if (!mask[c + r * COLS] && (e0 != 0.0 || e1 != 0.0)) {
                fprintf(stderr, "known pixel (%lld, %lld) was modified\n", (long long)r, (long long)c);
                return 6;
            }
            err = e0 > err ? e0 : err;
            err = e1 > err ? e1 : err;
        }
    printf("poisson: max relative error %.3e, %lld + %lld iterations on %lld unknowns per band\n", err,
        (long long)st[0].iterations, (long long)st[1].iterations, (long long)st[0].unknowns);
    if (!(err < 1e-8) || st[0].unknowns != expect_n)
        return 7;

    /* connected components: the reference's own test case (tests/approximation.h:55-75) */
    {
        uint8_t m[100] = { 0 };
        int32_t labels[100], num = -1;
        int count[3] = { 0, 0, 0 };
        for (int r = 1; r <= 2; ++r)
            for (int c = 1; c <= 2; ++c)
                m[c + r * 10] = 1;
        for (int r = 5; r <= 8; ++r)
            for (int c = 5; c <= 6; ++c)
                m[c + r * 10] = 1;
        rc = sa_label_components(ctx, m, 10, 10, 10, 1, labels, &num);
        if (rc != SA_OK || num != 2) {
            fprintf(stderr, "sa_label_components: rc %d, %d labels %s\n", rc, num, sa_last_error(ctx));
            return 8;
        }
        for (int i = 0; i < 100; ++i) {
            if (labels[i] < 0 || labels[i] > 2 || (labels[i] != 0) != (m[i] != 0))
                return 9;
            count[labels[i]]++;
        }
        printf("components: %d labels, sizes %d and %d\n", num, count[1], count[2]);
        if (count[1] != 4 || count[2] != 8)
            return 10;
    }
    sa_destroy(ctx);
    return 0;
}
