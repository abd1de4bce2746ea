/* dab2eti_multi.c -- N receivers in one process: the batched counterpart of dab2eti's RTL-SDR loop.
 *
 *     dab2eti_multi in0.iq out0.eti [in1.iq out1.eti ...]        ("-" = stdin / stdout)
 *
 * Every input is a 2.048 Msps uint8 I/Q source as an RTL-SDR delivers it -- a recording, a FIFO or
 * the pipe of `rtl_sdr -f <freq> -s 2048000 -`; every output receives that ensemble's 6144-byte
 * ETI(NI) frames, like dab2eti writes them to stdout (dab2eti.c:132-135).  All sources advance in
 * lock-step, one 262144-byte callback at a time (dabgpu_engine_pump, include/dabgpu.h).
 *
 *     gcc -O2 -I include examples/dab2eti_multi.c -L dabtools_b200 -ldabgpu -o dab2eti_multi
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "dabgpu.h"

int main(int argc, char **argv) {
  if (argc < 3 || (argc - 1) % 2) {
    fprintf(stderr, "Usage: %s in0.iq out0.eti [in1.iq out1.eti ...]\n", argv[0]);
    return 1;
  }
  const int n = (argc - 1) / 2;
  int *in = malloc(sizeof(int) * n), *out = malloc(sizeof(int) * n);
  for (int s = 0; s < n; s++) {
    const char *a = argv[1 + 2 * s], *b = argv[2 + 2 * s];
    in[s] = strcmp(a, "-") ? open(a, O_RDONLY) : 0;
    out[s] = strcmp(b, "-") ? open(b, O_WRONLY | O_CREAT | O_TRUNC, 0644) : 1;
    if (in[s] < 0 || out[s] < 0) {
      perror(in[s] < 0 ? a : b);
      return 1;
    }
  }
  dabgpu_engine *e = NULL;
  if (dabgpu_set_device(0) || dabgpu_engine_create(&e, n, 200000000u, DABGPU_ENGINE_VERBOSE)) {
    fprintf(stderr, "libdabgpu: %s\n", dabgpu_last_error_string());
    return 1;
  }
  dabgpu_engine_set_msc_batch(e, n >= 64 ? 2 : 1);
  const long long frames = dabgpu_engine_pump(e, n, in, out, -1);
  if (frames < 0) {
    fprintf(stderr, "libdabgpu: %s\n", dabgpu_last_error_string());
    return 1;
  }
  fprintf(stderr, "%lld ETI frames from %d stream(s)\n", frames, n);
  dabgpu_engine_destroy(e);
  return 0;
}
