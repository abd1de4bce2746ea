// hostlogic.cuh -- the reference's host-side control logic for the path (no arithmetic on
// sample or bit streams happens here): FIG parsing, ensemble bookkeeping, the lock state
// machine and the 16-CIF window.  The reference keeps this in C on the host too
// (fic.c:47-147, misc.c:14-27, dab.c:35-99); libdabgpu runs one instance per stream.
#pragma once
#include "../../include/dabgpu_ref_abi.h"
#include "common.cuh"

namespace dabgpu {

void host_fib_decode(tf_info_t *info, const uint8_t *fibs384, const uint8_t *crc_ok12, int nfibs, bool quiet);
void host_merge_info(ens_info_t *ei, const tf_info_t *info);
void host_init_ens(ens_info_t *ei);
// puncturing layout of a sub-channel as the reference's create_eti would decode it
int host_subch_shape(const subchannel_info_t *sc, dabgpu_cw_shape *out);
// bytes the sub-channel contributes to the ETI MST (misc.c:259-260)
static inline int host_subch_obytes(int nbits) { return ((nbits / 8) + 7) & 0xfff8; }

// glibc rand() compatible generator (TYPE_3), one per stream, for dab2eti.c:88-96
struct GlibcRand {
  int32_t r[31];
  int idx;
  void seed(unsigned s);
  int next();
};

// The reference keeps 5 TF buffers (4 in the window + the incoming one).  The device store is a
// deeper ring because MSC decoding is batched over up to D = 4 TFs and runs asynchronously to the
// front-end, whose host logic may additionally trail by one frame: a batch references D + 4 slots
// while up to D + 1 newer frames are being written.
enum { PHYS_TF_SLOTS = 14, MAX_MSC_BATCH = 4 };

// per-stream back-end state: dab_state_t without the 1.2 MB of frame buffers, which
// live on the device (dab.h:70-89)
struct BackendState {
  tf_info_t tf_info;
  ens_info_t ens_info;
  int win[16];      // physical CIF store slots (phys_tf*4+cif) of the 16-CIF window, [0] oldest
  int ncifs, tfidx, locked, okcount, ens_info_shown;
  int phys;         // physical TF slot the next incoming frame is written to: a ring deeper than the
                    // reference's 5 (PHYS_TF_SLOTS) that the engine advances with every frame it
                    // stores, so that FIG parsing and MSC decoding can trail the front-end
  uint64_t ens_version;  // bumped whenever ens_info's sub-channel table changes
  // ---- opt-in: follow multiplex reconfigurations the way EN 300 401 signals them (TODO.md:3; the
  // reference merges every FIG 0/1 entry as it arrives, also those of the *next* configuration, and
  // applies the table of the newest CIF to a frame that is 15-16 CIFs old).  With `follow`:
  //   * FIG 0/1 entries with the C/N flag set describe the next configuration and are collected apart;
  //   * FIG 0/0's change flags + occurrence change give the CIF count at which it becomes current: at
  //     that CIF the current table is REPLACED by the collected one (sub-channels can disappear);
  //   * an ETI frame is built with the table that was current for its own (oldest) CIF: the tables of
  //     the 4 transmission frames of the window are kept (hist) and `emit_*` is the oldest one's.
  bool follow = false;
  subchannel_info_t next_sub[64];
  int change_flags = 0, occurrence = -1, newest_cif_lo = -1;
  struct TableSnap {
    uint64_t version;
    subchannel_info_t sub[64];
  } hist[5];                 // ring by window position: hist[(hist_head + ncifs/4) % 5] is the incoming frame's table
  int hist_head = 0;
  uint64_t emit_version = 0; // version / table the frames emitted by the last host_process_frame use
  const subchannel_info_t *emit_sub = nullptr;
  void reset();
};

// one reference-shaped result of dab_process_frame(): how many ETI frames to build and from what
struct FrameWork {
  int n_eti;            // 0 or 4
  int win[4][16];       // window slots for each of the ETI frames
  uint8_t cif_hi[4], cif_lo[4];
};
// dab.c:35-99 with the FIC results (fibs, crc, ok_count) already decoded on the GPU.
// `slot` is the physical TF slot the frame's CIFs and FIBs were written to.
void host_process_frame(BackendState &st, const uint8_t *fibs384, const uint8_t *crc_ok12, int slot, FrameWork *out,
                        bool quiet);

}  // namespace dabgpu
