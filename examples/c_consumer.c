/* A plain-C consumer of libcodon_b200.so: what a non-Python host (or a cgo / JNI shim) links against.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -Iinclude examples/c_consumer.c -o /tmp/c_consumer \
 *       -Lcodon_b200 -lcodon_b200 -Wl,-rpath,$PWD/codon_b200
 *
 * Creates a context for CODON x4 in bf16 mode.  Without a CUDA device that must fail with CODON_ERR_CUDA and a
 * message (there is no CPU path); with one it sets a weight, checks that a forward before finalize_weights is
 * refused, and that the workspace size query answers.  Exit code 0 = the ABI behaved as include/codon_b200.h says.
 */
#include <stdio.h>
#include <string.h>

#include "codon_b200.h"

int main(void) {
  codon_ctx* ctx = NULL;
  int rc;
  printf("%s\n", codon_version());
  if (codon_selftest() != 0) { printf("selftest failed\n"); return 1; }
  if (codon_create(NULL, 0, 4, 1) != CODON_ERR_ARG) { printf("NULL out pointer accepted\n"); return 1; }
  rc = codon_create(&ctx, 0, 4, 1);
  if (rc == CODON_ERR_CUDA) {
    const char* msg = codon_last_error(NULL);
    printf("no device: rc=%d \"%s\"\n", rc, msg);
    return (ctx == NULL && msg && strlen(msg) > 0) ? 0 : 1;
  }
  if (rc != CODON_OK) { printf("codon_create: rc=%d %s\n", rc, codon_last_error(NULL)); return 1; }
  {
    static float w[64 * 1 * 3 * 3];
    const int64_t shape[4] = {64, 1, 3, 3};
    float dummy = 0.0f;
    if (codon_set_weight(ctx, "input.weight", w, shape, 4) != CODON_OK) { printf("set_weight: %s\n", codon_last_error(ctx)); return 1; }
    if (codon_set_weight(ctx, "no_such_layer.weight", w, shape, 4) != CODON_ERR_ARG) { printf("unknown name accepted\n"); return 1; }
    rc = codon_forward(ctx, &dummy, &dummy, &dummy, 1, 8, 8, 0, &dummy, 4, NULL);
    if (rc != CODON_ERR_STATE) { printf("forward before finalize: rc=%d\n", rc); return 1; }
    printf("refused as documented: \"%s\"\n", codon_last_error(ctx));
    if (codon_workspace_bytes(ctx, 1, 480, 640) == 0) { printf("workspace query failed\n"); return 1; }
  }
  codon_destroy(ctx);
  printf("ok\n");
  return 0;
}
