/* A plain-C consumer of libcodon_b200.so: what a non-Python host (or a cgo / JNI shim) links against.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude examples/c_consumer.c -o /tmp/c_consumer \
 *       -Lcodon_b200 -lcodon_b200 -Wl,-rpath,$PWD/codon_b200 -lm
 *
 *   c_consumer                                   ABI behaviour only (create / set_weight / call-order errors)
 *   c_consumer WEIGHTS.bin CASE.bin MODE TOL     a real forward from C: loads the flat weight file written by
 *                                                codon_b200.checkpoint.export_flat (codon_load_weights_file), runs
 *                                                codon_forward_host on the frames of CASE.bin (tests/golden/
 *                                                c_case_*.bin: inputs and the reference implementation's fp32
 *                                                output, written by oracle/make_golden.py) in arithmetic mode MODE
 *                                                (a codon_mode value) and compares with the reference output.
 *
 * Without a CUDA device codon_create must fail with CODON_ERR_CUDA and a message (there is no CPU path).
 * Exit code 0 = the ABI behaved as include/codon_b200.h says (and, with arguments, max |out - ref| <= TOL).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "codon_b200.h"

static int abi_checks(codon_ctx* ctx) {
  static float w[64 * 1 * 3 * 3];
  const int64_t shape[4] = {64, 1, 3, 3};
  float dummy = 0.0f;
  int rc;
  if (codon_set_weight(ctx, "input.weight", w, shape, 4) != CODON_OK) { printf("set_weight: %s\n", codon_last_error(ctx)); return 1; }
  if (codon_set_weight(ctx, "no_such_layer.weight", w, shape, 4) != CODON_ERR_ARG) { printf("unknown name accepted\n"); return 1; }
  rc = codon_forward(ctx, &dummy, &dummy, &dummy, 1, 8, 8, 0, &dummy, 4, NULL);
  if (rc != CODON_ERR_STATE) { printf("forward before finalize: rc=%d\n", rc); return 1; }
  printf("refused as documented: \"%s\"\n", codon_last_error(ctx));
  if (codon_workspace_bytes(ctx, 1, 480, 640) == 0) { printf("workspace query failed\n"); return 1; }
  if (codon_load_weights_file(ctx, "/nonexistent/weights.bin") != CODON_ERR_ARG) { printf("missing weight file accepted\n"); return 1; }
  return 0;
}

static int forward_case(codon_ctx* ctx, const char* weights, const char* case_path, double tol) {
  FILE* f;
  char magic[8];
  int dims[3];
  size_t n, i;
  float *x, *y, *ref, *out;
  double worst = 0.0;
  int rc;
  if (codon_load_weights_file(ctx, weights) != CODON_OK) { printf("load_weights_file: %s\n", codon_last_error(ctx)); return 1; }
  f = fopen(case_path, "rb");
  if (!f) { printf("cannot open %s\n", case_path); return 1; }
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CODONC1\0", 8) != 0 || fread(dims, sizeof(int), 3, f) != 3) {
    printf("%s is not a CODONC1 case file\n", case_path);
    fclose(f);
    return 1;
  }
  n = (size_t)dims[0] * (size_t)dims[1] * (size_t)dims[2];
  x = (float*)malloc(4 * n * sizeof(float));
  if (!x) { fclose(f); return 1; }
  y = x + n; ref = y + n; out = ref + n;
  if (fread(x, sizeof(float), 3 * n, f) != 3 * n) { printf("truncated case file\n"); fclose(f); free(x); return 1; }
  fclose(f);
  rc = codon_forward_host(ctx, x, y, out, dims[0], dims[1], dims[2]);
  if (rc != CODON_OK) { printf("codon_forward_host: rc=%d %s\n", rc, codon_last_error(ctx)); free(x); return 1; }
  for (i = 0; i < n; ++i) {
    const double d = fabs((double)out[i] - (double)ref[i]);
    if (!(d <= worst)) worst = d;     /* also catches NaN */
  }
  printf("forward from C: %d x %d x %d, %d kernel launches, max |out - reference| = %.3e (tolerance %.1e)\n", dims[0], dims[1],
         dims[2], codon_last_launch_count(ctx), worst, tol);
  free(x);
  return worst <= tol ? 0 : 1;
}

int main(int argc, char** argv) {
  codon_ctx* ctx = NULL;
  const int mode = argc > 3 ? atoi(argv[3]) : 1;
  int rc, bad;
  printf("%s\n", codon_version());
  if (codon_selftest() != 0) { printf("selftest failed\n"); return 1; }
  if (codon_create(NULL, 0, 4, 1) != CODON_ERR_ARG) { printf("NULL out pointer accepted\n"); return 1; }
  rc = codon_create(&ctx, 0, 4, mode);
  if (rc == CODON_ERR_CUDA) {
    const char* msg = codon_last_error(NULL);
    printf("no device: rc=%d \"%s\"\n", rc, msg);
    return (ctx == NULL && msg && strlen(msg) > 0) ? 0 : 1;
  }
  if (rc != CODON_OK) { printf("codon_create: rc=%d %s\n", rc, codon_last_error(NULL)); return 1; }
  bad = abi_checks(ctx);
  if (!bad && argc > 2) bad = forward_case(ctx, argv[1], argv[2], argc > 4 ? atof(argv[4]) : 1e-3);
  codon_destroy(ctx);
  if (!bad) printf("ok\n");
  return bad;
}
