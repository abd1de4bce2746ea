/* tables_host.c -- host-only build of the Mode I table accessors (dabgpu_tab_*).
 *
 * Same functions, same header (include/dabgpu_tables.h) as the ones libdabgpu.so exports from
 * csrc/api_batch.cu, compiled with plain gcc into dabtools_b200/libdabtables.so so that tools which
 * only need constants -- the synthetic transmitter that manufactures test and benchmark input --
 * never map the CUDA library (bench.py's reference arm must not load the product).
 * tests/test_tables.py pins both builds against the compiled reference tables. */
#include <stdint.h>
#include <string.h>

#include "../../include/dabgpu_tables.h"

#define EXPORT __attribute__((visibility("default")))

EXPORT int dabgpu_tab_shape(int kind, int a, int b, int32_t *out23) {
  dabgpu_cw_shape sh;
  int rc = 0;
  if (kind == 0)
    dabgpu_shape_fic(&sh);
  else if (kind == 1)
    rc = dabgpu_shape_uep(&sh, a);
  else if (kind == 2)
    rc = dabgpu_shape_eep(&sh, a, b, -1);
  else
    rc = -1;
  if (rc) return 1;
  _Static_assert(sizeof(sh) == 23 * sizeof(int32_t), "shape layout");
  memcpy(out23, &sh, sizeof sh);
  return 0;
}
EXPORT void dabgpu_tab_uep(int32_t *o) {
  for (int i = 0; i < 64; i++) {
    const dabgpu_uep_profile *p = &DABGPU_UEP[i];
    int32_t *r = o + 12 * i;
    r[0] = p->bitrate;
    r[1] = p->size_cu;
    r[2] = p->prot_level;
    for (int k = 0; k < 4; k++) {
      r[3 + k] = p->L[k];
      r[7 + k] = p->PI[k];
    }
    r[11] = p->pad_bits;
  }
}
EXPORT uint32_t dabgpu_tab_puncture_mask(int pi) { return dabgpu_puncture_mask(pi); }
EXPORT void dabgpu_tab_freq_deint(uint16_t *rev) { dabgpu_build_freq_deint(rev); }
EXPORT void dabgpu_tab_prs(uint8_t *q) { dabgpu_build_prs(q); }
EXPORT void dabgpu_tab_prbs(uint8_t *out, int nbytes) { dabgpu_build_prbs(out, nbytes); }
EXPORT uint16_t dabgpu_tab_crc16(const uint8_t *p, int n, uint16_t crc) { return dabgpu_crc16(p, n, crc); }
