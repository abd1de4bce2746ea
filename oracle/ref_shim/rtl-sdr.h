/* rtl-sdr.h -- TEST INFRASTRUCTURE ONLY: stand-in for librtlsdr's header (the library is not
 * installed in this image), declaring exactly what the reference's dab2eti.c uses
 * (dab2eti.c:38,57,76-103,137-249).  The matching implementation, oracle/ref_shim/rtlsdr_file.c,
 * replays an IQ file instead of driving a dongle, so that the UNMODIFIED dab2eti.c can be compiled
 * and run (a) with the reference's own objects and (b) against libdabgpu.so -- the drop-in proof. */
#ifndef ORACLE_RTL_SDR_SHIM_H
#define ORACLE_RTL_SDR_SHIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rtlsdr_dev rtlsdr_dev_t;
typedef void (*rtlsdr_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx);
uint32_t rtlsdr_get_device_count(void);
const char *rtlsdr_get_device_name(uint32_t index);
int rtlsdr_get_device_usb_strings(uint32_t index, char *manufact, char *product, char *serial);
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index);
int rtlsdr_close(rtlsdr_dev_t *dev);
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq);
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *dev, int *gains);
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *dev, int gain);
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual);
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate);
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev);
int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len);
int rtlsdr_cancel_async(rtlsdr_dev_t *dev);
#ifdef __cplusplus
}
#endif
#endif
