/* TEST INFRASTRUCTURE — stand-in for <ad9361.h> (libad9361 is not installed).
 * Only ad9361_set_bb_rate is used by the reference (plutogpssim.c:2131). */
#ifndef ORACLE_STUB_AD9361_H
#define ORACLE_STUB_AD9361_H

struct iio_device;
int ad9361_set_bb_rate(struct iio_device *dev, unsigned long rate);

#endif
