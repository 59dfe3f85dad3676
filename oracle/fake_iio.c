/* TEST INFRASTRUCTURE — capture backend standing in for libiio/libad9361/libcurl.
 *
 * Linked with the UNMODIFIED reference object (oracle/Makefile target
 * _ref/ref_verbatim).  iio_buffer_push() is where the reference hands a
 * 300 000-sample buffer to the SDR (plutogpssim.c:2146-2158); here it appends
 * the buffer to $FAKE_IIO_OUT instead.  The reference's mutex/condvar
 * handshake is timing based, so a fast consumer sees the calloc'ed all-zero
 * buffer first and may see an epoch twice: a push is kept
 * only if it differs from the previously kept one and is not all zero.
 * After $FAKE_IIO_EPOCHS kept buffers push() returns -1, which sends the
 * reference through its own shutdown path (plutogpssim.c:2153-2156,
 * 2181-2184).  Nothing here is part of the product path.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "stubs/iio.h"
#include "stubs/ad9361.h"
#include "stubs/curl/curl.h"

struct iio_buffer {
    size_t bytes;
    char *data;
    char *prev;
    long kept, pushes, limit;
    FILE *out;
};

/* Named handles + an optional call log ($FAKE_IIO_LOG, one line per set-up / tear-down call; pushes are not
 * logged, their count is timing dependent in the reference).  The same log is produced whether the caller is
 * the reference's TX thread or the product's radio sink (tests/test_sink.py compares the two). */
struct iio_device { char name[48]; };
struct iio_channel { char dev[48]; char name[24]; };
static int dummy_ctx;
static struct iio_device devs[8];
static struct iio_channel chans[32];
static int ndevs, nchans;

static void logf_(const char *fmt, ...) {
    const char *path = getenv("FAKE_IIO_LOG");
    if (!path) return;
    FILE *f = fopen(path, "a");
    if (!f) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(f, fmt, ap);
    va_end(ap);
    fputc('\n', f);
    fclose(f);
}

static struct iio_context *make_ctx(const char *how, const char *arg) {
    const char *deny = getenv("FAKE_IIO_NO_DEFAULT");   /* lets a test reach the -N / -U / pluto.local branches */
    if (strcmp(how, "default") == 0 && deny && *deny == '1') { logf_("create_context default -> none"); return NULL; }
    logf_("create_context %s %s", how, arg);
    return (struct iio_context *) &dummy_ctx;
}
struct iio_context *iio_create_default_context(void) { return make_ctx("default", ""); }
struct iio_context *iio_create_network_context(const char *h) { return make_ctx("network", h); }
struct iio_context *iio_create_context_from_uri(const char *u) { return make_ctx("uri", u); }
void iio_context_destroy(struct iio_context *c) { (void) c; logf_("context_destroy"); }
void iio_strerror(int err, char *dst, size_t len) { snprintf(dst, len, "fake iio error %d", err); }
unsigned int iio_context_get_devices_count(const struct iio_context *c) { (void) c; return 2; }
struct iio_device *iio_context_find_device(const struct iio_context *c, const char *n) {
    (void) c;
    for (int i = 0; i < ndevs; i++) if (strcmp(devs[i].name, n) == 0) return &devs[i];
    if (ndevs == 8) return NULL;
    snprintf(devs[ndevs].name, sizeof devs[ndevs].name, "%s", n);
    return &devs[ndevs++];
}
int iio_device_set_kernel_buffers_count(const struct iio_device *d, unsigned int nb) {
    logf_("kernel_buffers %s %u", d->name, nb);
    return 0;
}
struct iio_channel *iio_device_find_channel(const struct iio_device *d, const char *n, bool o) {
    (void) o;
    for (int i = 0; i < nchans; i++) if (strcmp(chans[i].dev, d->name) == 0 && strcmp(chans[i].name, n) == 0) return &chans[i];
    if (nchans == 32) return NULL;
    snprintf(chans[nchans].dev, sizeof chans[nchans].dev, "%s", d->name);
    snprintf(chans[nchans].name, sizeof chans[nchans].name, "%s", n);
    return &chans[nchans++];
}
ssize_t iio_channel_attr_write(const struct iio_channel *c, const char *a, const char *s) {
    logf_("attr %s/%s %s = %s", c->dev, c->name, a, s);
    return 0;
}
int iio_channel_attr_write_longlong(const struct iio_channel *c, const char *a, long long v) {
    logf_("attr %s/%s %s = %lld", c->dev, c->name, a, v);
    return 0;
}
int iio_channel_attr_write_double(const struct iio_channel *c, const char *a, double v) {
    logf_("attr %s/%s %s = %.17g", c->dev, c->name, a, v);
    return 0;
}
int iio_channel_attr_write_bool(const struct iio_channel *c, const char *a, bool v) {
    logf_("attr %s/%s %s = %s", c->dev, c->name, a, v ? "true" : "false");
    return 0;
}
void iio_channel_enable(struct iio_channel *c) { logf_("enable %s/%s", c->dev, c->name); }
void iio_channel_disable(struct iio_channel *c) { logf_("disable %s/%s", c->dev, c->name); }
int ad9361_set_bb_rate(struct iio_device *d, unsigned long r) { logf_("bb_rate %s %lu", d->name, r); return 0; }

struct iio_buffer *iio_device_create_buffer(const struct iio_device *d, size_t samples, bool cyclic) {
    logf_("create_buffer %s %zu %s", d->name, samples, cyclic ? "cyclic" : "non-cyclic");
    struct iio_buffer *b = calloc(1, sizeof *b);
    const char *path = getenv("FAKE_IIO_OUT");
    const char *lim = getenv("FAKE_IIO_EPOCHS");
    b->bytes = samples * 4;               /* interleaved int16 I,Q */
    b->data = calloc(1, b->bytes);
    b->prev = calloc(1, b->bytes);
    b->limit = lim ? atol(lim) : 10;
    b->out = path ? fopen(path, "wb") : NULL;
    return b;
}

void *iio_buffer_start(const struct iio_buffer *b) { return b->data; }

ssize_t iio_buffer_push(struct iio_buffer *b) {
    b->pushes++;
    {   /* optional real-time pace (a Pluto takes 0.1 s per 300000-sample buffer at 3 MS/s): $FAKE_IIO_PUSH_SLEEP_MS */
        const char *ms = getenv("FAKE_IIO_PUSH_SLEEP_MS");
        if (ms && atoi(ms) > 0) { struct timespec ts = {atoi(ms) / 1000, (long) (atoi(ms) % 1000) * 1000000L}; nanosleep(&ts, NULL); }
    }
    if (memcmp(b->data, b->prev, b->bytes) != 0) {   /* prev starts all-zero: zero pushes are dropped too */
        if (b->out) fwrite(b->data, 1, b->bytes, b->out);
        memcpy(b->prev, b->data, b->bytes);
        b->kept++;
    }
    if (b->kept >= b->limit) return -1;
    return (ssize_t) b->bytes;
}

void iio_buffer_destroy(struct iio_buffer *b) {
    fprintf(stderr, "fake_iio: %ld pushes, %ld kept\n", b->pushes, b->kept);
    logf_("buffer_destroy");
    if (b->out) fclose(b->out);
    free(b->data); free(b->prev); free(b);
}

CURLcode curl_global_init(long f) { (void) f; return CURLE_OK; }
void curl_global_cleanup(void) {}
CURL *curl_easy_init(void) { return NULL; }
CURLcode curl_easy_setopt(CURL *h, CURLoption o, ...) { (void) h; (void) o; return CURLE_OK; }
CURLcode curl_easy_perform(CURL *h) { (void) h; return CURLE_GOT_NOTHING; }
void curl_easy_cleanup(CURL *h) { (void) h; }
