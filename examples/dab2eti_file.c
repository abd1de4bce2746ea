/* dab2eti_file.c -- the reference's dab2eti receive loop (dab2eti.c:60-135) on a recorded uint8 I/Q
 * file instead of librtlsdr, written against include/dabgpu_ref_abi.h and linked with libdabgpu.so.
 * It is the reference's own call sequence: sdr_init / sdr_demod / dab_process_frame / eti_callback.
 *
 *   gcc -O2 -Iinclude examples/dab2eti_file.c -Ldabtools_b200 -ldabgpu -Wl,-rpath,$PWD/dabtools_b200 -o dab2eti_file
 *   ./dab2eti_file capture.iq > out.eti
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "dabgpu.h"
#include "dabgpu_ref_abi.h"

static void eti_callback(uint8_t *eti) {
  if (write(1, eti, 6144) != 6144) exit(2);
}

int main(int argc, char **argv) {
  if (argc != 2) {
    fprintf(stderr, "Usage: %s capture.iq  (2.048 Msps uint8 I/Q)  > out.eti\n", argv[0]);
    return 1;
  }
  FILE *f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 1;
  }
  if (dabgpu_set_device(0)) {
    fprintf(stderr, "%s\n", dabgpu_last_error_string());
    return 1;
  }
  struct dab_state_t *dab;
  static struct sdr_state_t sdr;
  memset(&sdr, 0, sizeof sdr);
  init_dab_state(&dab, &sdr, eti_callback);
  dab->device_type = DAB_DEVICE_RTLSDR;
  sdr_init(&sdr);
  for (;;) {
    /* rtlsdr_callback (dab2eti.c:117-130) */
    size_t n = fread(sdr.input_buffer, 1, DEFAULT_BUF_LENGTH, f);
    if (n != DEFAULT_BUF_LENGTH) break;
    sdr.input_buffer_len = (int)n;
    /* demod_thread_fn (dab2eti.c:68-71) */
    if (sdr_demod(&dab->tfs[dab->tfidx], &sdr)) dab_process_frame(dab);
  }
  fclose(f);
  return 0;
}
