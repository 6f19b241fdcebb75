/*
 * oracle_run -- command-line front end of the CPU oracle (TEST INFRASTRUCTURE).
 * Prints the same " Final error:  %.12e" line as the reference (src/main.cpp:572) so that
 * the reference's own test_driver.py string comparison (test/test_driver.py:42-63) applies.
 *
 *   oracle_run <problem> <dim> <nCells> [t_end|-1] [CFL]
 *   oracle_run bench <problem> <dim> <level> <warmup> <steps> <threads>
 */
#include "mmf_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int problem_from_key(const char *k)
{
    /* src/problem.cpp:37-59 */
    static const char *keys[] = { "vortex_xy", "vortex_zx", "vortex_yz", "radsod",
                                  "sod3d_x", "sod3d_y", "sod3d_z", "ffstep" };
    for (int i = 0; i < 8; ++i) if (!strcmp(k, keys[i])) return i;
    fprintf(stderr, "Problem %s is not supported.\n", k);
    exit(1);
}

int main(int argc, char **argv)
{
    if (argc >= 8 && !strcmp(argv[1], "bench")) {
        int problem = problem_from_key(argv[2]);
        int dim = atoi(argv[3]), level = atoi(argv[4]), warm = atoi(argv[5]), steps = atoi(argv[6]);
        int threads = atoi(argv[7]);
        uint64_t h;
        double s = orc_bench_threads(problem, dim, level, warm, steps, threads, 0.45, &h);
        long n = 1L << level; long cells = (dim == 3) ? n * n * n : n * n;
        printf("cells %ld steps %d threads %d seconds %.6f cell-updates/s %.6e hash %016llx\n",
               cells, steps, threads, s, 3.0 * cells * steps / s, (unsigned long long) h);
        return 0;
    }
    if (argc < 4) {
        fprintf(stderr, "usage: %s <problem> <dim> <nCells> [t_end|-1] [CFL]\n", argv[0]);
        return 2;
    }
    int problem = problem_from_key(argv[1]);
    int dim = atoi(argv[2]);
    long n = atol(argv[3]);
    double t_end = (argc > 4) ? atof(argv[4]) : -1.;
    double cfl = (argc > 5) ? atof(argv[5]) : 0.45;
    double err, t;
    int steps = orc_run(problem, dim, n, t_end, cfl, 0, NULL, -1, &err, &t, NULL);
    printf(" steps: %d  t: %.17g\n", steps, t);
    printf(" Final error:  %.12e\n", err);
    return 0;
}
