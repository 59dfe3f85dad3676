/* The three-library pipeline of INTEGRATION.md ("The same pipeline from C") as a complete PLAIN C program:
 * navigation file -> descriptors (libgpshost) -> I/Q batches (libgpsiq) -> file sink (libgpshost), two pinned
 * buffers in flight.  tests/test_front_end_cpu.py compiles it with gcc -std=c11 (the headers must be C, not only
 * C++) and runs it against the oracle-backed mock of libgpsiq; on a GPU box the same binary runs the real library.
 *   pipeline_example <nav file> <out file> <epochs> <batch> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpshost.h"
#include "gpsiq.h"
#include "gpssink.h"

#define CHECK(call, msg) do { if ((call) != 0) { fprintf(stderr, "%s: %s\n", #call, msg); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const long total = atol(argv[3]);
    const int B = atoi(argv[4]);
    gpshost_config hc;
    memset(&hc, 0, sizeof hc);
    hc.nav_path = argv[1];
    hc.pos_mode = GPSHOST_POS_LLH;
    hc.pos[0] = 30.286502; hc.pos[1] = 120.032669; hc.pos[2] = 100;
    hc.sample_rate = 2600000;
    hc.max_chan = 12;
    gpshost_scenario *sc;
    CHECK(gpshost_open(&sc, &hc), gpshost_last_error());
    gpsiq_config gc;
    memset(&gc, 0, sizeof gc);
    gc.max_chan = 12; gc.samples_per_epoch = GPSSINK_PUSH_PAIRS; gc.max_epochs = B;
    gpsiq_ctx *gq;
    CHECK(gpsiq_create(&gq, &gc), gpsiq_last_error(NULL));
    gpssink *out;
    CHECK(gpssink_open_file(&out, argv[2]), gpssink_last_error());

    gpsiq_chan_desc *desc = gpsiq_host_alloc((size_t) B * 12 * sizeof *desc);
    int16_t *iq[2] = {gpsiq_host_alloc((size_t) B * GPSSINK_PUSH_PAIRS * 4), gpsiq_host_alloc((size_t) B * GPSSINK_PUSH_PAIRS * 4)};
    int64_t ticket[2] = {0, 0};
    long submitted = 0, fetched = 0;
    int pending[2], head = 0, tail = 0;      /* sizes of the (at most two) batches in flight */

    while (fetched < total) {
        while (submitted < total && tail - head < 2) {              /* keep the GPU one batch ahead */
            const int n = (int) (total - submitted < B ? total - submitted : B);
            CHECK(gpshost_next(sc, desc, n), gpshost_last_error());
            CHECK(gpsiq_submit(gq, desc, n), gpsiq_last_error(gq));
            pending[tail++ & 1] = n;
            submitted += n;
        }
        const int k = head & 1, n = pending[head++ & 1];
        if (ticket[k] > 0) CHECK(gpssink_wait(out, ticket[k]), gpssink_last_error());   /* the sink is done with this buffer */
        CHECK(gpsiq_fetch(gq, iq[k]), gpsiq_last_error(gq));
        ticket[k] = gpssink_submit(out, iq[k], (size_t) n * GPSSINK_PUSH_PAIRS);           /* drained while the next batch is fetched */
        if (ticket[k] < 0) { fprintf(stderr, "%s\n", gpssink_last_error()); return 1; }
        fetched += n;
    }
    CHECK(gpssink_close(out), gpssink_last_error());
    gpsiq_host_free(desc); gpsiq_host_free(iq[0]); gpsiq_host_free(iq[1]);
    gpsiq_destroy(gq);
    gpshost_close(sc);
    return 0;
}
