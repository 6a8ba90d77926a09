/*
 * Calling librlic_b200.so from plain C: the README example of the reference
 * (256 x 256 noise texture, u = cos(2x), v = sin(x), 65-tap triangle kernel,
 * periodic walls, 5 iterations), through the same entry point the Python
 * binding uses.
 *
 *   gcc -std=c99 -Iinclude examples/convolve_from_c.c -o convolve_from_c \
 *       -Lrlic_b200 -lrlic_b200 -Wl,-rpath,$PWD/rlic_b200 -lm
 *
 * Prints a checksum of the inputs and of the result (sums of the bit patterns,
 * so that a test can compare them exactly) and exits 0; without a CUDA device
 * it prints the library's error and exits with the error code.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rlic_b200.h"

enum { N = 256, TAPS = 65 };

/* any deterministic non-negative texture will do; the test regenerates it in NumPy */
static double noise(uint64_t *state)
{
    *state = *state * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(*state >> 11) / 9007199254740992.0;
}

static uint64_t bit_sum(const double *a, int n)
{
    uint64_t sum = 0;
    for (int i = 0; i < n; ++i) {
        uint64_t bits;
        memcpy(&bits, &a[i], sizeof bits);
        sum += bits;
    }
    return sum;
}

int main(void)
{
    if (rlic_b200_abi_version() != RLIC_B200_ABI_VERSION) {
        fprintf(stderr, "header/library ABI mismatch\n");
        return 100;
    }
    double *texture = malloc(sizeof(double) * N * N), *u = malloc(sizeof(double) * N * N),
           *v = malloc(sizeof(double) * N * N), *out = malloc(sizeof(double) * N * N);
    double taps[TAPS];
    if (!texture || !u || !v || !out) return 101;

    const double pi = 3.14159265358979323846;
    uint64_t state = 42;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double x = pi * j / (N - 1);
            texture[i * N + j] = noise(&state);
            u[i * N + j] = cos(2 * x);
            v[i * N + j] = sin(x);
        }
    for (int k = 0; k < TAPS; ++k) taps[k] = 1.0 - fabs(-1.0 + 2.0 * k / (TAPS - 1));

    printf("inputs %016llx\n", (unsigned long long)(bit_sum(texture, N * N) + bit_sum(u, N * N) +
                                                     bit_sum(v, N * N) + bit_sum(taps, TAPS)));
    fflush(stdout);

    int rc = rlic_b200_convolve_f64(texture, u, v, N, N, taps, TAPS, RLIC_B200_VELOCITY,
                                    RLIC_B200_PERIODIC, RLIC_B200_PERIODIC,
                                    RLIC_B200_PERIODIC, RLIC_B200_PERIODIC, 5, out);
    if (rc != RLIC_B200_OK) {
        fprintf(stderr, "rlic_b200_convolve_f64 failed (%d): %s\n", rc, rlic_b200_last_error());
        return rc;
    }
    printf("devices %d launches %lld checksum %016llx\n", rlic_b200_device_count(),
           (long long)rlic_b200_launch_count(), (unsigned long long)bit_sum(out, N * N));
    free(texture); free(u); free(v); free(out);
    return 0;
}
