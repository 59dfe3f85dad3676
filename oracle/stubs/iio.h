/* TEST INFRASTRUCTURE — minimal stand-in for <iio.h> (libiio is not installed).
 *
 * Declares just the libiio entry points the reference's TX thread calls
 * (/root/reference/plutogpssim.c:2058-2190) so that the reference translation
 * unit compiles where it lies.  The definitions are in oracle/fake_iio.c: a
 * capture backend that records pushed buffers instead of talking to an SDR.
 * Nothing under oracle/ is part of the shipped product path.
 */
#ifndef ORACLE_STUB_IIO_H
#define ORACLE_STUB_IIO_H

#include <stdbool.h>
#include <stddef.h>
#include <sys/types.h>

struct iio_context;
struct iio_device;
struct iio_channel;
struct iio_buffer;

struct iio_context *iio_create_default_context(void);
struct iio_context *iio_create_network_context(const char *host);
struct iio_context *iio_create_context_from_uri(const char *uri);
void iio_context_destroy(struct iio_context *ctx);
void iio_strerror(int err, char *dst, size_t len);
unsigned int iio_context_get_devices_count(const struct iio_context *ctx);
struct iio_device *iio_context_find_device(const struct iio_context *ctx, const char *name);
int iio_device_set_kernel_buffers_count(const struct iio_device *dev, unsigned int nb);
struct iio_channel *iio_device_find_channel(const struct iio_device *dev, const char *name, bool output);
ssize_t iio_channel_attr_write(const struct iio_channel *chn, const char *attr, const char *src);
int iio_channel_attr_write_longlong(const struct iio_channel *chn, const char *attr, long long val);
int iio_channel_attr_write_double(const struct iio_channel *chn, const char *attr, double val);
int iio_channel_attr_write_bool(const struct iio_channel *chn, const char *attr, bool val);
void iio_channel_enable(struct iio_channel *chn);
void iio_channel_disable(struct iio_channel *chn);
struct iio_buffer *iio_device_create_buffer(const struct iio_device *dev, size_t samples, bool cyclic);
void iio_buffer_destroy(struct iio_buffer *buf);
void *iio_buffer_start(const struct iio_buffer *buf);
ssize_t iio_buffer_push(struct iio_buffer *buf);

#endif
