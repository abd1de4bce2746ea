// hostlogic.cu -- host-side control logic of the receive path (see hostlogic.cuh).
#include "hostlogic.cuh"

namespace dabgpu {

// FIG 0/0, 0/1 and 0/2 extraction, behaviour of fic.c:47-130 fib_parse()
// what the reference's parser does not look at (fic.c:56-70): the C/N flag of FIG type 0 and the
// change flags / occurrence change of FIG 0/0 (EN 300 401 clause 6.4); only filled when asked for
struct ReconfigInfo {
  subchannel_info_t next_sub[64];  // FIG 0/1 entries of the next configuration (C/N = 1)
  int change_flags, occurrence;
};
static void parse_fib(tf_info_t *info, const uint8_t *fib, bool quiet, ReconfigInfo *rc = nullptr) {
  int pos = 0;
  while (fib[pos] != 0xff && pos < 30) {
    const int fig_type = fib[pos] >> 5;
    const int fig_len = fib[pos] & 0x1f;
    pos++;  // pos now at the first data byte of the FIG
    if (fig_type == 0) {
      const int ext = fib[pos] & 0x1f;
      const int pd = (fib[pos] >> 5) & 1;
      const bool next_cfg = rc && (fib[pos] >> 7);
      if (ext == 0) {  // ensemble information
        info->EId = (uint16_t)((fib[pos + 1] << 8) | fib[pos + 2]);
        info->CIFCount_hi = fib[pos + 3] & 0x1f;
        info->CIFCount_lo = fib[pos + 4];
        if (rc) {
          rc->change_flags = fib[pos + 3] >> 6;
          if (rc->change_flags && fig_len >= 6) rc->occurrence = fib[pos + 5];
        }
      } else if (ext == 1) {  // sub-channel organisation
        for (int j = pos + 1; j < pos + fig_len;) {
          const int id = fib[j] >> 2;
          subchannel_info_t *sc = next_cfg ? &rc->next_sub[id] : &info->subchans[id];
          sc->id = id;
          sc->start_cu = ((fib[j] & 3) << 8) | fib[j + 1];
          sc->slForm = fib[j + 2] >> 7;
          sc->eepprot = sc->slForm;
          if (sc->slForm == 0) {
            const dabgpu_uep_profile &p = DABGPU_UEP[fib[j + 2] & 0x3f];
            sc->uep_index = fib[j + 2] & 0x3f;
            sc->size = p.size_cu;
            sc->bitrate = p.bitrate;
            sc->protlev = p.prot_level;
            j += 3;
          } else {
            const int option = (fib[j + 2] >> 4) & 7;
            sc->protlev = ((fib[j + 2] >> 2) & 3) | (option << 2);
            sc->size = ((fib[j + 2] & 3) << 8) | fib[j + 3];
            // options 2..7 are reserved; the reference would index past eeptable[] here
            const int lv = sc->protlev & 7;
            sc->bitrate = (sc->size / dabgpu_eep_size_mul(lv)) * (lv < 4 ? 8 : 32);
            j += 4;
          }
        }
      } else if (ext == 2) {  // service organisation: only ASCTy of stream-audio components
        for (int j = pos + 1; j < pos + fig_len;) {
          j += pd ? 4 : 2;
          const int ncomp = fib[j++] & 0x0f;
          for (int k = 0; k < ncomp; k++, j += 2) {
            const int tmid = fib[j] >> 6;
            if (tmid == 0)
              info->subchans[fib[j + 1] >> 2].ASCTy = fib[j] & 0x3f;
            else if (tmid != 3 && !quiet)  // fic.c:113-119
              fprintf(stderr, "Unhandled TMid %d for subchannel %d\n", tmid, fib[j + 1] >> 2);
          }
        }
      }
    }
    pos += fig_len;
  }
}

// fic.c:132-147
void host_fib_decode(tf_info_t *info, const uint8_t *fibs, const uint8_t *crc_ok, int nfibs, bool quiet) {
  memset(info, 0, sizeof *info);
  for (int i = 0; i < 64; i++) info->subchans[i].id = info->subchans[i].ASCTy = -1;
  for (int i = 0; i < nfibs; i++)
    if (crc_ok[i]) parse_fib(info, fibs + 32 * i, quiet);
}
static void host_fib_decode_follow(tf_info_t *info, ReconfigInfo *rc, const uint8_t *fibs, const uint8_t *crc_ok,
                                   int nfibs, bool quiet) {
  memset(info, 0, sizeof *info);
  for (int i = 0; i < 64; i++) {
    info->subchans[i].id = info->subchans[i].ASCTy = -1;
    rc->next_sub[i].id = rc->next_sub[i].ASCTy = -1;
  }
  rc->change_flags = 0;
  rc->occurrence = -1;
  for (int i = 0; i < nfibs; i++)
    if (crc_ok[i]) parse_fib(info, fibs + 32 * i, quiet, rc);
}

// misc.c:14-27
void host_merge_info(ens_info_t *ei, const tf_info_t *info) {
  for (int i = 0; i < 64; i++)
    if (info->subchans[i].id >= 0) ei->subchans[i] = info->subchans[i];
  ei->EId = info->EId;
  if (ei->CIFCount_hi == 0xff) {
    ei->CIFCount_hi = info->CIFCount_hi;
    ei->CIFCount_lo = info->CIFCount_lo;
  }
}

// dab.c:23-25
void host_init_ens(ens_info_t *ei) {
  memset(ei, 0, sizeof *ei);
  for (int i = 0; i < 64; i++) ei->subchans[i].id = ei->subchans[i].ASCTy = -1;
  ei->CIFCount_hi = 0xff;
  ei->CIFCount_lo = 0xff;
}

int host_subch_shape(const subchannel_info_t *sc, dabgpu_cw_shape *out) {
  if (sc->eepprot) return dabgpu_shape_eep(out, sc->protlev, sc->size, sc->bitrate);
  return dabgpu_shape_uep(out, sc->uep_index);
}

void GlibcRand::seed(unsigned s) {
  if (s == 0) s = 1;
  r[0] = (int32_t)s;
  for (int i = 1; i < 31; i++) {
    const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    r[i] = (int32_t)w;
  }
  idx = 0;
  for (int i = 0; i < 310; i++) next();
}
int GlibcRand::next() {
  const int f = (idx + 3) % 31;
  const uint32_t v = (uint32_t)r[f] + (uint32_t)r[idx];
  r[f] = (int32_t)v;
  idx = (idx + 1) % 31;
  return (int)(v >> 1);
}

void BackendState::reset() {
  memset(&tf_info, 0, sizeof tf_info);
  host_init_ens(&ens_info);
  memset(win, 0, sizeof win);
  ncifs = tfidx = locked = okcount = ens_info_shown = 0;
  phys = 0;
  ens_version = 1;
  for (int i = 0; i < 64; i++) next_sub[i].id = next_sub[i].ASCTy = -1;
  change_flags = 0;
  occurrence = newest_cif_lo = -1;
  for (auto &h : hist) h.version = 0;
  hist_head = 0;
  emit_version = 0;
  emit_sub = nullptr;
}

// misc.c:14-27, additionally reporting whether the sub-channel table changed
static bool merge_and_diff(ens_info_t *ei, const tf_info_t *info) {
  bool changed = false;
  for (int i = 0; i < 64; i++) {
    const subchannel_info_t &y = info->subchans[i];
    if (y.id < 0) continue;
    subchannel_info_t &x = ei->subchans[i];
    if (x.id != y.id || x.eepprot != y.eepprot || x.slForm != y.slForm || x.uep_index != y.uep_index ||
        x.start_cu != y.start_cu || x.size != y.size || x.bitrate != y.bitrate || x.protlev != y.protlev)
      changed = true;
    x = y;
  }
  ei->EId = info->EId;
  if (ei->CIFCount_hi == 0xff) {
    ei->CIFCount_hi = info->CIFCount_hi;
    ei->CIFCount_lo = info->CIFCount_lo;
  }
  return changed;
}

// dab.c:35-99
void host_process_frame(BackendState &st, const uint8_t *fibs, const uint8_t *crc_ok, int slot, FrameWork *out,
                        bool quiet) {
  out->n_eti = 0;
  int ok_count = 0;
  for (int i = 0; i < 12; i++) ok_count += crc_ok[i] ? 1 : 0;
  ReconfigInfo rcf;
  if (ok_count > 0) {
    if (st.follow)
      host_fib_decode_follow(&st.tf_info, &rcf, fibs, crc_ok, 12, quiet);
    else
      host_fib_decode(&st.tf_info, fibs, crc_ok, 12, quiet);
  }

  if (ok_count == 12) {
    st.okcount++;
    if (st.okcount >= 10 && !st.locked) {
      st.locked = 1;
      if (!quiet) fprintf(stderr, "Locked\n");
    }
  } else {
    st.okcount = 0;
    if (st.locked) {
      st.locked = 0;
      if (!quiet) fprintf(stderr, "Lock lost, resetting ringbuffer\n");
      st.ncifs = 0;
      st.tfidx = 0;
      st.hist_head = 0;
      return;
    }
  }
  if (!st.locked) return;

  if (st.follow && ok_count > 0) {
    // The configuration announced for CIF count `occurrence` becomes current with the transmission
    // frame that carries that CIF (in Mode I a reconfiguration starts a transmission frame): the
    // collected next table replaces the current one -- before this frame's own FIG 0/1 entries,
    // which already describe the new configuration as the current one, are merged below.
    const int first_lo = (st.tf_info.CIFCount_lo + 250 - 3) % 250;  // count of this frame's first CIF
    if (st.occurrence >= 0 && ((st.occurrence - first_lo + 250) % 250) < 4) {
      bool any = false;
      for (int i = 0; i < 64; i++) any |= st.next_sub[i].id >= 0;
      if (any) {
        for (int i = 0; i < 64; i++) {
          st.ens_info.subchans[i] = st.next_sub[i];
          st.next_sub[i].id = st.next_sub[i].ASCTy = -1;
        }
        st.ens_version++;
      }
      st.occurrence = -1;
      st.change_flags = 0;
    }
    for (int i = 0; i < 64; i++)
      if (rcf.next_sub[i].id >= 0) st.next_sub[i] = rcf.next_sub[i];
    if (rcf.change_flags && rcf.occurrence >= 0) {
      st.change_flags = rcf.change_flags;
      st.occurrence = rcf.occurrence;
    }
    st.newest_cif_lo = st.tf_info.CIFCount_lo;
  }
  if (merge_and_diff(&st.ens_info, &st.tf_info)) st.ens_version++;
  if (st.follow) {
    // remember which table this transmission frame's CIFs were sent under; frames are emitted with the
    // table of their own (oldest) CIF, i.e. of the oldest transmission frame of the window
    BackendState::TableSnap &h = st.hist[(st.hist_head + (st.ncifs < 16 ? st.ncifs / 4 : 4)) % 5];
    if (h.version != st.ens_version) {
      h.version = st.ens_version;
      memcpy(h.sub, st.ens_info.subchans, sizeof h.sub);
    }
  }

  if (st.ncifs < 16) {
    for (int k = 0; k < 4; k++) st.win[st.ncifs++] = slot * 4 + k;
  } else {
    if (!st.ens_info_shown) {
      if (!quiet) dump_ens_info(&st.ens_info);
      st.ens_info_shown = 1;
    }
    out->n_eti = 4;
    if (st.follow) {  // the table of the oldest transmission frame of the window; its slot is reused in 4 frames
      st.emit_version = st.hist[st.hist_head].version;
      st.emit_sub = st.hist[st.hist_head].sub;
      st.hist_head = (st.hist_head + 1) % 5;
    }
    for (int k = 0; k < 4; k++) {
      memcpy(out->win[k], st.win, sizeof st.win);
      out->cif_hi[k] = st.ens_info.CIFCount_hi;
      out->cif_lo[k] = st.ens_info.CIFCount_lo;
      // misc.c:305-313: free-running CIF counter
      if (++st.ens_info.CIFCount_lo == 250) {
        st.ens_info.CIFCount_lo = 0;
        if (++st.ens_info.CIFCount_hi == 20) st.ens_info.CIFCount_hi = 0;
      }
      memmove(st.win, st.win + 1, 15 * sizeof(int));
      st.win[15] = slot * 4 + k;
    }
  }
  st.tfidx = (st.tfidx + 1) % 5;
}

}  // namespace dabgpu

using namespace dabgpu;

// ---- reference-signature entry points that are pure host logic ----------------------------------
DABGPU_EXPORT void fib_decode(struct tf_info_t *info, struct tf_fibs_t *fibs, int nfibs) {
  host_fib_decode(info, &fibs->FIB[0][0], fibs->FIB_CRC_OK, nfibs, false);
}
DABGPU_EXPORT void merge_info(struct ens_info_t *ei, struct tf_info_t *info) { host_merge_info(ei, info); }

static void dump_subchannels(const subchannel_info_t *sc) {
  for (int i = 0; i < 64; i++)
    if (sc[i].id >= 0)
      fprintf(stderr, "SubChId=%d, slForm=%d, StartAddress=%d, size=%d, bitrate=%d, ASCTy=0x%02x\n", sc[i].id,
              sc[i].slForm, sc[i].start_cu, sc[i].size, sc[i].bitrate, sc[i].ASCTy);
}
DABGPU_EXPORT void dump_tf_info(struct tf_info_t *info) {  // fic.c:33-45
  fprintf(stderr, "EId=0x%04x, CIFCount = %d %d\n", info->EId, info->CIFCount_hi, info->CIFCount_lo);
  dump_subchannels(info->subchans);
}
DABGPU_EXPORT void dump_ens_info(struct ens_info_t *info) {  // misc.c:316-328
  fprintf(stderr, "ENSEMBLE_INFO: EId=0x%04x, CIFCount = %d %d\n", info->EId, info->CIFCount_hi,
          info->CIFCount_lo);
  dump_subchannels(info->subchans);
}

// misc.c:153-213 as a host function (the batched path builds the same header on the GPU)
DABGPU_EXPORT int init_eti(uint8_t *eti, struct ens_info_t *info) {
  int n = 0, nst = 0, fl = 0;
  const bool odd = info->CIFCount_lo & 1;
  eti[n++] = 0xff;
  eti[n++] = odd ? 0xf8 : 0x07;
  eti[n++] = odd ? 0xc5 : 0x3a;
  eti[n++] = odd ? 0x49 : 0xb6;
  eti[n++] = info->CIFCount_lo;
  for (int j = 0; j < 64; j++)
    if (info->subchans[j].id >= 0) {
      nst++;
      fl += info->subchans[j].bitrate * 3 / 4;
    }
  fl += nst + 1 + 24;
  eti[n++] = (uint8_t)(0x80 | nst);
  const int fp = (info->CIFCount_hi * 250 + info->CIFCount_lo) % 8;
  eti[n++] = (uint8_t)((fp << 5) | (1 << 3) | ((fl >> 8) & 7));
  eti[n++] = (uint8_t)fl;
  for (int j = 0; j < 64; j++) {
    const subchannel_info_t &sc = info->subchans[j];
    if (sc.id < 0) continue;
    const int tpl = sc.slForm == 0 ? (0x10 | (sc.protlev - 1)) : (0x20 | sc.protlev);
    const int stl = sc.bitrate * 3 / 8;
    eti[n++] = (uint8_t)((sc.id << 2) | ((sc.start_cu >> 8) & 3));
    eti[n++] = (uint8_t)sc.start_cu;
    eti[n++] = (uint8_t)((tpl << 2) | ((stl >> 8) & 3));
    eti[n++] = (uint8_t)stl;
  }
  eti[n++] = 0xff;
  eti[n++] = 0xff;
  const uint16_t hcrc = (uint16_t)~dabgpu_crc16(eti + 4, n - 4, 0xffff);
  eti[n++] = (uint8_t)(hcrc >> 8);
  eti[n++] = (uint8_t)hcrc;
  return n;
}
