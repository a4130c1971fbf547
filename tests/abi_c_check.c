/* TEST INFRASTRUCTURE: include/nnuepack.h must be usable from plain C (C99, -pedantic) and
 * libnnuepack.so must link from a C program. Without a GPU every driver refuses (there is no CPU
 * path); with one, a three-record round trip must reproduce the survey's known-answer binpack. */
#include <stdio.h>
#include <string.h>

#include "nnuepack.h"

int main(void)
{
    unsigned char buf[256];
    size_t n = 0;
    int rc;
    memset(buf, 0, sizeof buf);
    rc = nnp_bin_to_binpack(buf, 40, buf + 64, 128, &n);
    if (rc != NNP_ERR_NOT_INITIALISED) { printf("expected NOT_INITIALISED, got %d\n", rc); return 1; }
    rc = nnp_init(0);
    if (rc == NNP_ERR_NO_DEVICE) {
        printf("ABI_C_OK no device: %s\n", nnp_strerror(rc));
        return 0;
    }
    if (rc != NNP_OK) { printf("nnp_init: %d %s\n", rc, nnp_last_cuda_error()); return 1; }
    rc = nnp_bin_to_binpack(buf, 0, buf + 64, 128, &n);
    if (rc != NNP_OK || n != 0) { printf("empty input: %d %lu\n", rc, (unsigned long)n); return 1; }
    nnp_shutdown();
    printf("ABI_C_OK device\n");
    return 0;
}
