/* TEST INFRASTRUCTURE -- NOT a CPU fallback and never shipped or loaded by the product.
 *
 * A stand-in for libgpsiq.so that lets the CPU test suite (no GPU in the build container) exercise the control
 * flow of the command-line front end gpsiq_sim -- option parsing, host orchestration, batching, the two pinned
 * buffers in flight, the hand-off to the sink's writer thread -- end to end against the reference's stream.
 * It implements only the entry points gpsiq_sim and the literal drop-in build of the reference (oracle/dropin/) call, on top of the parity oracle (oracle/gpsiq_oracle.c,
 * compiled into this mock by tests/test_front_end_cpu.py), and lives in a scratch directory next to a COPY of the
 * gpsiq_sim binary (whose rpath is $ORIGIN); the real libgpsiq.so refuses to work without a GPU
 * (tests/test_capi_load.py::test_no_cpu_fallback_without_gpu) and stays the only library in the package directory. */
#include <stdlib.h>
#include <string.h>

#include "gpsiq.h"
#include "gpsiq_desc.h"   /* the shipped header-only body of gpsiq_make_desc */

int oracle_synth(const gpsiq_chan_desc *desc, int n_epochs, int nslots, int samples_per_epoch, int carrier_mode,
                 double *carr_state, int16_t *iq_out, double *carr_trace);

struct pending { gpsiq_chan_desc *desc; int n; struct pending *next; };
struct gpsiq_ctx {
    gpsiq_config cfg;
    double carr[GPSIQ_MAX_CHAN];
    struct pending *head, *tail;
    int64_t calls;
};

int gpsiq_create(gpsiq_ctx **ctx, const gpsiq_config *cfg) {
    gpsiq_ctx *c = calloc(1, sizeof *c);
    c->cfg = *cfg;
    *ctx = c;
    return GPSIQ_OK;
}
void gpsiq_destroy(gpsiq_ctx *c) { free(c); }
void *gpsiq_host_alloc(size_t bytes) { return malloc(bytes); }
void gpsiq_host_free(void *p) { free(p); }
const char *gpsiq_last_error(const gpsiq_ctx *c) { (void) c; return "mock_gpsiq (test infrastructure)"; }
int64_t gpsiq_launch_count(const gpsiq_ctx *c) { (void) c; return 0; }   /* no kernels: says so in the summary line */

int gpsiq_submit(gpsiq_ctx *c, const gpsiq_chan_desc *desc, int n_epochs) {
    if (n_epochs < 1 || n_epochs > c->cfg.max_epochs) return GPSIQ_ERR_ARG;
    struct pending *p = calloc(1, sizeof *p);
    const size_t bytes = (size_t) n_epochs * c->cfg.max_chan * sizeof *desc;
    p->desc = malloc(bytes);
    memcpy(p->desc, desc, bytes);        /* the real submit copies the descriptors too: the caller reuses its buffer */
    p->n = n_epochs;
    if (c->tail) c->tail->next = p; else c->head = p;
    c->tail = p;
    return GPSIQ_OK;
}

int gpsiq_fetch(gpsiq_ctx *c, int16_t *iq_out) {
    struct pending *p = c->head;
    if (!p) return GPSIQ_ERR_ARG;
    c->head = p->next;
    if (!c->head) c->tail = NULL;
    const int rc = oracle_synth(p->desc, p->n, c->cfg.max_chan, c->cfg.samples_per_epoch, c->cfg.carrier_mode, c->carr,
                                iq_out, NULL);
    free(p->desc);
    free(p);
    return rc == 0 ? GPSIQ_OK : GPSIQ_ERR_ARG;
}

/* the two calls of the literal drop-in (oracle/dropin/dropin_loop.inc) */
int gpsiq_make_desc(gpsiq_chan_desc *out, int carrier_mode, int prn, double f_carr, double f_code, double delt,
                    double carr_phase, double code_phase, const uint64_t *dwrd60, int iword, int ibit, int icode,
                    double gain, int carr_phase_is_new) {
    return gpsiq_make_desc_inline(out, carrier_mode, prn, f_carr, f_code, delt, carr_phase, code_phase, dwrd60, iword, ibit,
                                  icode, gain, carr_phase_is_new);
}

int gpsiq_synth(gpsiq_ctx *c, const gpsiq_chan_desc *desc, int n_epochs, int16_t *iq_out) {
    if (n_epochs < 1 || n_epochs > c->cfg.max_epochs || !iq_out) return GPSIQ_ERR_ARG;
    return oracle_synth(desc, n_epochs, c->cfg.max_chan, c->cfg.samples_per_epoch, c->cfg.carrier_mode, c->carr, iq_out, NULL) == 0
               ? GPSIQ_OK : GPSIQ_ERR_ARG;
}

/* the multi-device entry points the front end drives (gpsiq_sim -g N): here one queue whatever N -- the stream is the
 * same by construction, which is exactly what the real library has to reproduce with N devices */
struct gpsiq_multi { gpsiq_ctx *c; int n; struct pending *begun_head, *begun_tail; int16_t *out[64]; int nb, ne; };
int gpsiq_multi_create(gpsiq_multi **m, const gpsiq_config *cfg, int n_devices) {
    if (n_devices < 1 || n_devices > 16) return GPSIQ_ERR_ARG;
    gpsiq_multi *x = calloc(1, sizeof *x);
    x->n = n_devices;
    gpsiq_create(&x->c, cfg);
    *m = x;
    return GPSIQ_OK;
}
void gpsiq_multi_destroy(gpsiq_multi *m) { if (m) { gpsiq_destroy(m->c); free(m); } }
int gpsiq_multi_submit(gpsiq_multi *m, const gpsiq_chan_desc *desc, int n_epochs) { return gpsiq_submit(m->c, desc, n_epochs); }
int gpsiq_multi_fetch_begin(gpsiq_multi *m, int16_t *iq_out) {
    if (m->nb - m->ne >= 2 * m->n || m->nb - m->ne >= 64) return GPSIQ_ERR_CAPACITY;
    m->out[m->nb++ % 64] = iq_out;
    return GPSIQ_OK;
}
int gpsiq_multi_fetch_end(gpsiq_multi *m) {
    if (m->ne >= m->nb) return GPSIQ_ERR_ARG;
    return gpsiq_fetch(m->c, m->out[m->ne++ % 64]);
}
int gpsiq_multi_fetch(gpsiq_multi *m, int16_t *iq_out) {
    const int rc = gpsiq_multi_fetch_begin(m, iq_out);
    return rc ? rc : gpsiq_multi_fetch_end(m);
}
int gpsiq_multi_devices(const gpsiq_multi *m) { return m->n; }
int64_t gpsiq_multi_launch_count(const gpsiq_multi *m) { (void) m; return 0; }
const char *gpsiq_multi_last_error(const gpsiq_multi *m) { (void) m; return "mock_gpsiq (test infrastructure)"; }
