/* dabgpu.h -- C ABI of libdabgpu: the dabtools receive hot path on NVIDIA B200 (sm_100a).
 *
 * Two layers:
 *
 *  1. include/dabgpu_ref_abi.h -- the reference's own function signatures and struct
 *     layouts for this path (sdr_demod, *_depuncture, viterbi, fic_decode, create_eti,
 *     dab_process_frame, ...), so that an unmodified dab2eti.o links against libdabgpu
 *     instead of the reference's objects.  Those are batch-of-one calls.
 *
 *  2. this header -- the additive batched API the throughput path uses: many independent
 *     codewords / CIFs / ensemble streams per call, device-resident or host buffers.
 *
 * Conventions: every function returns 0 on success or a negative DABGPU_ERR_* code and
 * records a message retrievable with dabgpu_last_error_string() (per thread).  There is
 * no CPU fallback: without a usable sm_100 device every compute entry point fails with
 * DABGPU_ERR_NO_DEVICE.  Pointers are plain host pointers unless `on_device` is non-zero,
 * in which case they are device pointers valid on the current device.  Work is enqueued
 * on the stream set with dabgpu_set_stream() (default: the legacy default stream); calls
 * taking host pointers synchronise that stream before returning.
 */
#ifndef DABGPU_H
#define DABGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DABGPU_OK 0
#define DABGPU_ERR_CUDA (-1)
#define DABGPU_ERR_ARG (-2)
#define DABGPU_ERR_NO_DEVICE (-3)
#define DABGPU_ERR_STATE (-4)

/* ---- status / device ------------------------------------------------------------------ */
int dabgpu_last_error(void);
const char *dabgpu_last_error_string(void);
void dabgpu_clear_error(void);
int dabgpu_device_count(void);
int dabgpu_set_device(int dev);
void dabgpu_set_stream(void *cuda_stream);
int dabgpu_synchronize(void);

/* ---- constant tables (host only, no GPU needed; include/dabgpu_tables.h) ---------------- */
/* kind: 0 = FIC, 1 = UEP (a = table index 0..63), 2 = EEP (a = level 0..7, b = size in CU).
 * out receives the dabgpu_cw_shape as 23 int32 (nbits, in_bits, n_regions, then 5 x
 * {steps, pi, step0, in0}).  Replaces ueptable/eeptable/pvec (dab_tables.c:16-127). */
int dabgpu_tab_shape(int kind, int a, int b, int32_t *out23);
/* 64 rows x {bitrate, size_cu, prot_level, L1..L4, PI1..PI4, pad_bits} */
void dabgpu_tab_uep(int32_t *out64x12);
/* Host-only (no GPU needed): the Viterbi step bytes (see csrc/viterbi.cuh) that the MSC gather
 * derives for one sub-channel (kind/a/b as in dabgpu_tab_shape, first capacity unit start_cu) from
 * a de-interleaved CIF given as 55296 bytes of 0/1 -- uep/eep_depuncture (depuncture.c:84-132) in
 * the library's input format.  Returns the number of step bytes written, or a negative error. */
int dabgpu_tab_depuncture_steps(int kind, int a, int b, int start_cu, const uint8_t *cif_bits55296,
                                uint8_t *steps, int steps_cap);
uint32_t dabgpu_tab_puncture_mask(int pi);
void dabgpu_tab_freq_deint(uint16_t *rev1536);   /* rev_freq_deint_tab, dab_tables.c:164-357 */
void dabgpu_tab_prs(uint8_t *quarter_turns1536); /* prs_static, sdr_prstab.c */
void dabgpu_tab_prbs(uint8_t *out, int nbytes);  /* energy-dispersal sequence, misc.c:41-58 */

/* ---- batched channel decoding ------------------------------------------------------------ */
/* n codewords of `nbits` information bits each.  soft: n rows of 4*(nbits+6) reference
 * soft symbols (127 = 0, 128 = erasure, 129 = 1; anything <128 / >128 is sliced), row pitch
 * soft_pitch bytes.  out: n rows of ceil(nbits/8) bytes, pitch out_pitch (multiple of 4,
 * >= 4*ceil(nbits/32)).  descramble != 0 additionally XORs the energy-dispersal PRBS
 * (needs nbits <= 9216).  Batched form of viterbi(), viterbi.c:352-452. */
int dabgpu_viterbi_batch(const uint8_t *soft, size_t soft_pitch, int n, int nbits, uint8_t *out,
                         size_t out_pitch, int descramble, int on_device);
/* The soft demapper (see DABGPU_ENGINE_SOFT) on one frame, without the synchronisers: 393216 bytes of
 * uint8 I/Q as sdr_read_fifo leaves them -> 230400 symbol bytes in the layout of
 * demapped_transmission_frame_t (fic 9216 + msc 221184). */
int dabgpu_demod_frame_soft(const uint8_t *frame393216, uint8_t *soft230400);
/* Soft-decision variant (opt-in): the symbols are NOT sliced; each is weighted with the
 * reference's own metric table gen_met(amp 1, noise 1.0, bias 0, scale 4) (viterbi.c:126-191,
 * :455-462), which is defined for 121..135 around the erasure value 128 (log-likelihood ratio about
 * 10.7 per LSB) and overflows beyond -- values outside saturate to 121 / 135 here.  Same arguments
 * as dabgpu_viterbi_batch; bit-exact with the reference's viterbi() for symbols inside the table
 * (the hard-decision alphabet 127 / 128 / 129 is a special case). */
int dabgpu_viterbi_soft_batch(const uint8_t *soft, size_t soft_pitch, int n, int nbits, uint8_t *out,
                              size_t out_pitch, int descramble, int on_device);
void dabgpu_tab_soft_metrics(int32_t *out2x256); /* that table; INT32_MIN outside 121..135 */

/* n_groups FIC groups of 2304 hard bits (one byte per bit, values 0/1; 4 groups = the 3
 * FIC symbols of one transmission frame) -> 96 descrambled bytes (3 FIBs) + 3 CRC flags
 * each.  Batched form of the loop body of fic_decode(), fic.c:185-206. */
int dabgpu_fic_decode_batch(const uint8_t *fic_bits, int n_groups, uint8_t *fibs, uint8_t *crc_ok,
                            int on_device);

/* Work accounting of the last dabgpu_*_batch call on this thread: trellis steps decoded. */
uint64_t dabgpu_last_trellis_steps(void);
/* Number of CUDA kernels this library has launched since it was loaded (all threads). */
uint64_t dabgpu_launch_count(void);

/* ---- single-frame front-end (parity / debugging) ------------------------------------------ */
/* frame: 393216 bytes of uint8 I/Q exactly as sdr_demod has them in sdr->buffer after
 * sdr_read_fifo.  Runs the synchronisers of input_sdr.c:65-112; out[0..3] = coarse_timeshift,
 * fine_timeshift (bytes), coarse_freq_shift (carriers), ok; *fine_freq_hz as sdr_state_t. */
int dabgpu_sync_frame(const uint8_t *frame, int force_timesync, int32_t *out4, double *fine_freq_hz);
/* FFT + DQPSK + demap of all 76 symbols regardless of the synchronisers (input_sdr.c:114-162):
 * symbols / symbols_d: 76*2048 complex float each (fftshifted like sdr->symbols; row 0 of
 * symbols_d is not written), bits: 230400 bytes (fic 9216 then msc 221184).  Any may be NULL. */
int dabgpu_demod_frame_debug(const uint8_t *frame, float *symbols, float *symbols_d, uint8_t *bits);

/* ---- the batched receiver ------------------------------------------------------------------- */
/* S independent ensemble streams in lock-step; each stream is one instance of the reference's
 * receive loop (dab2eti.c:60-115: sdr_demod -> dab_process_frame -> tuner feedback). */
typedef struct dabgpu_engine dabgpu_engine;

#define DABGPU_ENGINE_VERBOSE 1        /* print the reference's stderr messages (Locked, ...) */
#define DABGPU_ENGINE_VIRTUAL_TUNER 2  /* apply the tuner feedback as a software NCO on ingest */
/* Soft-decision receiver (opt-in; the default is the reference's hard-decision demapper, bit-exact
 * with it).  The demapper hands symbols instead of bits to the channel decoder: for every channel bit
 * a byte 128 -+ round(8 x), clipped to 121..135, where x is the real / negated imaginary part of the
 * normalised differential product s_l conj(s_l-1) / |s_l-1|^2 that input_sdr.c:132-158 slices at 0;
 * depuncturing inserts 128, and the Viterbi decoder weights each symbol with the reference's own
 * metric table (dabgpu_viterbi_soft_batch).  process_demapped() of a soft engine takes such symbol
 * bytes (230400 per frame) instead of 0/1 bytes.  Worth about 2 dB at the decoder input. */
#define DABGPU_ENGINE_SOFT 4
/* Follow multiplex reconfigurations the way EN 300 401 signals them (opt-in; the reference's open TODO,
 * TODO.md:3).  The reference -- and the default mode, bit-exact with it -- merges every FIG 0/1 entry
 * into the sub-channel table as it arrives, including the entries that announce the NEXT configuration
 * (C/N flag, sent for about six seconds before a change), never removes a sub-channel, and builds an
 * ETI frame with the table of the newest CIF although the frame is 15-16 CIFs older.  With this flag
 * the C/N flag, FIG 0/0's change flags and its occurrence-change field are honoured: the announced
 * table replaces the current one at the signalled CIF, and every ETI frame is decoded and described
 * (STC) with the table that was current for its own CIF. */
#define DABGPU_ENGINE_FOLLOW_RECONFIG 8

typedef struct {
  int32_t locked, okcount, ncifs, tfidx;          /* dab_state_t, dab.h:83-86 */
  int32_t coarse_timeshift, fine_timeshift;       /* sdr_state_t, input_sdr.h:17-19 */
  int32_t coarse_freq_shift, last_ok;             /* last_ok = return value of the last sdr_demod */
  double fine_freq_shift;
  uint32_t frequency;                             /* tuner frequency after feedback (Hz) */
  int32_t n_subchannels;
  uint64_t frames_demodulated, eti_frames, fib_crc_errors;
} dabgpu_stream_status;

int dabgpu_engine_create(dabgpu_engine **out, int n_streams, uint32_t tuner_hz, int flags);
void dabgpu_engine_destroy(dabgpu_engine *e);
/* One rtlsdr callback for every stream: iq + s*pitch holds chunk_len bytes (<= 262144, multiple
 * of 16) for stream s.  Equivalent to one demod_thread_fn iteration per stream. */
int dabgpu_engine_feed_iq(dabgpu_engine *e, const uint8_t *iq, size_t pitch, int chunk_len, int on_device);
/* The same in two halves, so that the host->device copy of the next callback overlaps the
 * processing of the current one: submit_iq() starts the upload of a host chunk (pinned memory
 * makes it asynchronous) and returns; feed_submitted() processes the oldest submitted chunk.  At
 * most three chunks may be in flight.  feed_iq(host pointer) == submit_iq + feed_submitted. */
int dabgpu_engine_submit_iq(dabgpu_engine *e, const uint8_t *iq, size_t pitch, int chunk_len);
/* how many of the chunks given to submit_iq() are still being copied (their host buffers must not
 * be overwritten yet); copies complete in submission order */
int dabgpu_engine_uploads_in_flight(dabgpu_engine *e);
/* Zero-copy source for samples that already are in device memory as one contiguous capture per
 * stream (a recording loaded into HBM, or the output buffer of a device-side producer): stream s
 * at iq_device + s * pitch, `len` bytes each.  attach_capture() must precede the first samples;
 * feed_capture(chunk_len) then consumes the next chunk_len bytes of every stream exactly like
 * feed_iq() would, reading them in place instead of copying them into the engine's FIFO
 * (rtlsdr_callback's cbWrite loop, dab2eti.c:125, has no counterpart).  The capture must stay
 * valid and unchanged while the engine lives.  Not available with DABGPU_ENGINE_VIRTUAL_TUNER. */
int dabgpu_engine_attach_capture(dabgpu_engine *e, const uint8_t *iq_device, size_t pitch, size_t len);
int dabgpu_engine_feed_capture(dabgpu_engine *e, int chunk_len);
/* Treat the attached capture as one period of an endless signal: feed_capture() wraps around at
 * its end instead of refusing (FIFO positions are taken modulo the capture length anyway).  A
 * capture holding a whole number of transmission frames keeps the streams frame-aligned across
 * the seam; the 15 CIFs after it mix two passes in the time de-interleaver, like any receiver fed
 * a looped recording.  Used by bench.py for timed regions longer than HBM could hold as samples. */
int dabgpu_engine_set_capture_cyclic(dabgpu_engine *e, int on);
int dabgpu_engine_feed_submitted(dabgpu_engine *e);
/* Back-end only: one demapped transmission frame (fic 9216 + msc 221184 bytes of 0/1, i.e. the
 * payload of demapped_transmission_frame_t) for every stream with mask[s] != 0 (mask NULL = all);
 * equivalent to dab_process_frame per stream. */
int dabgpu_engine_process_demapped(dabgpu_engine *e, const uint8_t *tfs, size_t pitch, const uint8_t *mask,
                                   int on_device);
/* Second producer of demapped transmission frames: the Psion Wavefinder's USB packets
 * (input_wf.c:23-115).  packets + s*pitch holds n_packets[s] packets of 524 bytes -- the symbols of
 * ONE transmission frame of stream s as the device delivers them between two NULL-symbol packets
 * (byte 2 = symbol number: 2..4 FIC, 5..76 MSC, others ignored; bytes 12..395 = 192 little-endian
 * words of DQPSK decisions in carrier order); n_packets[s] == 0: no frame for that stream.  Each
 * frame is de-interleaved and demapped on the device (wf_demap_symbol) and handed to the same
 * back-end as process_demapped; a frame without all three FIC symbols gets NULL FIBs like
 * fic_decode does for has_fic == 0 (fic.c:167-175), symbols that did not arrive keep the content the
 * frame buffer had five frames earlier (dab->tfs[5]).  The device's own timing/AFC loop
 * (wf_sync.c) needs the hardware and is not part of this path.  Host pointers. */
int dabgpu_engine_process_wavefinder(dabgpu_engine *e, const uint8_t *packets, size_t pitch,
                                     const int32_t *n_packets);
/* Let the MSC decoding (time de-interleave -> Viterbi -> ETI) lag so that one launch covers the
 * frames of `calls` (1..4) frame-producing feed/process calls: more, better balanced work per
 * launch; frames then come out in bursts.  dabgpu_engine_flush() decodes whatever is queued (call
 * it at the end of a capture); its frames are then reported by eti_count()/fetch_eti(). */
int dabgpu_engine_set_msc_batch(dabgpu_engine *e, int calls);
int dabgpu_engine_flush(dabgpu_engine *e);
/* MSC decoding and ETI assembly run on an internal stream so that they overlap the next frames'
 * front-end kernels.  dabgpu_engine_join() makes the caller's stream (dabgpu_set_stream) wait for
 * everything issued so far -- needed before reading dabgpu_engine_eti_device() from that stream or
 * before recording an event that should cover the work; dabgpu_engine_fetch_eti() waits by itself. */
int dabgpu_engine_join(dabgpu_engine *e);
/* ETI frames produced by the last feed/process/flush call, in call order, then stream order. */
int dabgpu_engine_eti_count(dabgpu_engine *e);
const uint8_t *dabgpu_engine_eti_device(dabgpu_engine *e);
/* copies up to max_frames frames (6144 bytes each) and their stream indices to the host;
 * returns the number of frames, or a negative error */
int dabgpu_engine_fetch_eti(dabgpu_engine *e, uint8_t *eti, int32_t *stream_ids, int max_frames);
int dabgpu_engine_status(dabgpu_engine *e, int stream, dabgpu_stream_status *out);
/* Sub-channel filter (the reference's open TODO, TODO.md:3): only the sub-channels whose SubChId
 * bit is set in `mask` are Viterbi-decoded and carried in the ETI frames of `stream` (-1: every
 * stream) -- NST, the STC list, FL and the MST shrink accordingly, FIC and everything else stay as
 * misc.c:153-314 builds them.  Takes effect with the next transmission frame; frames already queued
 * keep the old selection.  Default: all ones (the reference's behaviour). */
int dabgpu_engine_set_subchannel_mask(dabgpu_engine *e, int stream, uint64_t mask);
int dabgpu_engine_set_seed(dabgpu_engine *e, int stream, unsigned seed); /* srand() of dab2eti.c:88-96 */
uint64_t dabgpu_engine_trellis_steps(dabgpu_engine *e);
/* Optional device-side timing of the engine's kernels with CUDA events on the launch stream
 * (adds a stream synchronisation to every feed/process call and runs the engine's streams one
 * after the other, so that every kernel is timed alone).  Order of the entries:
 * ingest, fifo_read, sync, demod, fic_prep, fic_viterbi, msc_gather, msc_viterbi, eti_pack. */
#define DABGPU_ENGINE_KERNELS 9
int dabgpu_engine_enable_timing(dabgpu_engine *e, int on);
int dabgpu_engine_kernel_times(dabgpu_engine *e, double *ms_total, uint64_t *launches, int n);
/* cumulative host wall-clock microseconds: control build, waiting for the GPU, state machines,
 * job construction */
void dabgpu_engine_host_times(dabgpu_engine *e, double *us4);

/* Streaming ingest front: dab2eti's rtlsdr_read_async + demod_thread_fn pair (dab2eti.c:60-135,
 * 237-239) for n_streams sources at once.  Reads 262144-byte callbacks from in_fds[s] (recording,
 * FIFO, `rtl_sdr -` pipe; blocking reads pace a live source), keeps the engine fed through
 * submit_iq / feed_submitted with uploads two callbacks ahead, and writes every 6144-byte ETI frame
 * of stream s to out_fds[s] like eti_callback does (out_fds or single entries may be NULL / -1).
 * Returns at the first source that ends (or after max_callbacks callbacks per stream when >= 0),
 * after flushing the engine: the number of ETI frames written, or a negative DABGPU_ERR_*. */
long long dabgpu_engine_pump(dabgpu_engine *e, int n_streams, const int *in_fds, const int *out_fds,
                             long long max_callbacks);

/* ---- ETI consumers on the device (what follows the path: SURVEY 8f-3) ---------------------------
 * dabgpu_eti_extract_subchannel: eti2mpa.c:32-67 for a batch of frames -- for every 6144-byte frame
 * at eti + f*6144 the bytes of sub-channel `subchid` (STL*8 of them, located through the STC) are
 * copied to out + f*out_pitch and their count to out_len[f] (-1: the frame does not carry it).
 * eti/out/out_len are device pointers when on_device != 0 (asynchronous on the current stream),
 * host pointers otherwise.  out_pitch >= the sub-channel's size (at most 4608 = 384 kbit/s).
 * dabgpu_eti_check: per frame a mask of DABGPU_ETI_BAD_* (0 = consistent frame): ERR/FSYNC and its
 * alternation with FCT, FC (MID, FL against the STC), header CRC, end-of-frame CRC, 0x55 padding
 * (misc.c:153-296; the reference's TODO "check the details of the ETI stream"). */
#define DABGPU_ETI_BAD_SYNC 1u
#define DABGPU_ETI_BAD_FC 2u
#define DABGPU_ETI_BAD_HCRC 4u
#define DABGPU_ETI_BAD_EOF_CRC 8u
#define DABGPU_ETI_BAD_PADDING 16u
int dabgpu_eti_extract_subchannel(const uint8_t *eti, int n_frames, int subchid, uint8_t *out, size_t out_pitch,
                                  int32_t *out_len, int on_device);
int dabgpu_eti_check(const uint8_t *eti, int n_frames, uint32_t *flags, int on_device);
/* The same on the ETI frames of the engine's last feed/process/flush call, where they lie in HBM
 * (dabgpu_engine_eti_device): nothing but the extracted bytes / the flags crosses PCIe. */
int dabgpu_engine_extract_subchannel(dabgpu_engine *e, int subchid, uint8_t *out_host, size_t out_pitch,
                                     int32_t *out_len_host);
int dabgpu_engine_check_eti(dabgpu_engine *e, uint32_t *flags_host);

/* ---- ABI self-description (sizeof / offsetof of include/dabgpu_ref_abi.h's structs) -------- */
int dabgpu_sizeof_dab_state(void);
int dabgpu_sizeof_sdr_state(void);
int dabgpu_sizeof_tf(void);
void dabgpu_abi_offsets(int32_t *out15);

#ifdef __cplusplus
}
#endif
#endif /* DABGPU_H */
