/* A plain-C consumer of include/dmp2.h: proves the boundary is a C ABI (no C++ or torch types in the signatures) and
 * exercises the entry points that need no GPU.  Built and run by tests/test_host.py::test_header_is_plain_c_and_links.
 * Exit code 0 = every check passed; it prints one line per check. */
#include <stdio.h>
#include <string.h>

#include "dmp2.h"

int main(void) {
    int fails = 0;
    /* 1. pure host arithmetic of the halo-sharded partition: rows in multiples of 8 that cover [0, L) exactly */
    int L = 2048, world = 4, prev_end = 0, rank;
    for (rank = 0; rank < world; rank++) {
        int r0 = -1, r1 = -1;
        int rc = dmp2_strip_rows(L, world, rank, &r0, &r1);
        if (rc != DMP2_OK || r0 != prev_end || r1 <= r0 || (r0 % 8) != 0) { printf("strip_rows rank %d: rc %d rows [%d, %d)\n", rank, rc, r0, r1); fails++; }
        prev_end = r1;
    }
    if (prev_end != L) { printf("strip rows end at %d, expected %d\n", prev_end, L); fails++; }
    printf("strip_rows: %s\n", fails ? "FAILED" : "ok");

    /* 2. argument checking never crashes */
    if (dmp2_create(NULL, 0, 0, NULL, NULL, NULL) != DMP2_ERR_BAD_ARG) { printf("dmp2_create(NULL...) did not return BAD_ARG\n"); fails++; }
    if (dmp2_set_conv_mode(NULL, 0) != DMP2_ERR_BAD_ARG || dmp2_set_graph(NULL, 1) != DMP2_ERR_BAD_ARG ||
        dmp2_reserve(NULL, 100, 100) != DMP2_ERR_BAD_ARG || dmp2_launch_count(NULL) != 0) { printf("NULL engine not rejected\n"); fails++; }
    dmp2_destroy(NULL);
    printf("null-argument handling: %s\n", fails ? "FAILED" : "ok");

    /* 3. creating an engine: on a box without an sm_100 device this must fail loudly (there is no CPU fallback);
     *    with one, a one-tensor state_dict must be refused as incomplete */
    {
        static float eye[22 * 22];
        const char* names[1] = {"embed.weight"};
        const float* ptrs[1];
        int64_t numels[1] = {22 * 22};
        dmp2_engine* e = NULL;
        int rc;
        ptrs[0] = eye;
        rc = dmp2_create(&e, 0, 1, names, ptrs, numels);
        printf("dmp2_create -> %d (%s)\n", rc, dmp2_last_error(NULL));
        if (rc != DMP2_ERR_NO_DEVICE && rc != DMP2_ERR_MISSING_WEIGHT) fails++;
        if (e != NULL) fails++;
        if (strlen(dmp2_last_error(NULL)) == 0) fails++;
    }
    printf("%s\n", fails ? "FAILED" : "all ok");
    return fails ? 1 : 0;
}
