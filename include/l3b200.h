/* l3b200.h -- C-ABI of the B200-native MPEG-1/2/2.5 Layer III granule decode path.
 *
 * This is the drop-in boundary for audio-formats' MP3 hot path.  The reference has no FFI seam for
 * MP3 (stream.d calls D functions of minimp3.d / minimp3_ex.d directly), so the boundary is placed
 * where SURVEY.md 8b puts it:
 *
 *   layer 1  "shim"  (l3b_ctx_*, l3b_decode_batch)      replaces the granule work under
 *                     mp3dec_decode_frame (minimp3.d:1492-1581): L3_decode (:1196) +
 *                     mp3d_synth_granule (:1408) for every granule of every stream of a batch.
 *   layer 2  "host"  (l3b_scan_*, l3b_stream_*)         is what the D host does before/around the
 *                     shim: frame sync (minimp3.d:1436-1485), side-info parsing (:487-611),
 *                     bit-reservoir slicing (:1170-1194) and the stream layer of minimp3_ex.d
 *                     (index :566, seek :662, read :787, open :929) behind the AudioStream surface
 *                     of stream.d (:115, :150, :429, :656, :1095, :1209).  No D compiler exists in
 *                     the build image, so this layer is implemented in C++ in the same library and
 *                     mirrored by the (uncompiled) D sources under audio_formats_b200/dhost/.
 *
 * Conventions: every function returns 0 or a negative MP3D_E_* style code (minimp3_ex.d:30-34),
 * never throws, never calls back, never keeps caller memory after returning.  A context is bound to
 * one GPU and must be used by one host thread at a time.  There is NO CPU fallback: every compute
 * entry point fails with L3B_E_NOGPU when no CUDA device is usable.
 */
#ifndef L3B200_H
#define L3B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define L3B_OK 0
#define L3B_E_PARAM (-1)     /* MP3D_E_PARAM   minimp3_ex.d:30 */
#define L3B_E_MEMORY (-2)    /* MP3D_E_MEMORY  minimp3_ex.d:31 */
#define L3B_E_IOERROR (-3)   /* MP3D_E_IOERROR minimp3_ex.d:32 */
#define L3B_E_USER (-4)      /* MP3D_E_USER    minimp3_ex.d:33 (also: "not an MP3") */
#define L3B_E_DECODE (-5)    /* MP3D_E_DECODE  minimp3_ex.d:34 */
#define L3B_E_NOGPU (-16)    /* no usable CUDA device / kernel launch failed (ours) */
#define L3B_E_UNSUPPORTED (-17) /* reserved (was: Layer I/II stream, which now decodes) */

/* ------------------------------------------------------------------------------------------------
 * Descriptors: what the host prepass hands to the GPU.
 * One l3b_grch_desc_t per granule-channel (16 bytes); the nch descriptors of a granule are
 * adjacent, granules of a stream are consecutive in decode order.  Packed copy of the fields of
 * L3_gr_info_t (minimp3.d:189-196) plus the two things the sequential decoder carries implicitly:
 * where the bits are, and whether decoder state was zeroed before this granule. */
typedef struct {
    uint32_t bit_start; /* first bit of this granule-channel's part2_3 data, relative to the stream's main-data blob */
    uint32_t w1;        /* part_23_length[0:12] big_values[12:21] global_gain[21:29] block_type[29:31] mixed_block_flag[31] */
    uint32_t w2;        /* scalefac_compress[0:9] table_select0[9:14] 1[14:19] 2[19:24] preflag[24] scalefac_scale[25]
                           count1_table[26] scfsi[27:31] second_granule_of_frame[31] */
    uint32_t w3;        /* region1_start/2[0:9] region2_start/2[9:18] subblock_gain0[18:21] 1[21:24] 2[24:27]
                           header byte3 >> 4 (mode, mode_ext)[27:31] state_reset_before[31] */
} l3b_grch_desc_t;

/* One per stream of a batch. */
typedef struct {
    uint64_t maindata_off;   /* byte offset of this stream's main-data blob inside the batch blob (16-byte aligned) */
    uint32_t maindata_bytes; /* valid bytes; the blob must be followed by >= 16 readable zero bytes */
    uint32_t n_granules;     /* decodable granules, decode order (Layer III: 576 frames each; Layer I / II: 384 frames each) */
    uint64_t first_grch;     /* index of the stream's first descriptor in the batch descriptor array */
    uint64_t pcm_off;        /* float offset of the stream's first delivered sample in the PCM output */
    uint64_t pcm_skip;       /* interleaved samples to drop from the front of the decoded signal (encoder delay, minimp3_ex.d:862-867) */
    uint64_t pcm_count;      /* interleaved samples to deliver after the skip (padding trim, minimp3_ex.d:869-873) */
    uint8_t nch;             /* 1 or 2 */
    uint8_t sr_idx;          /* row of the sfb tables, minimp3.d:523 */
    uint8_t mpeg1;           /* 1: MPEG-1 (2 granules/frame), 0: MPEG-2 / 2.5 LSF */
    uint8_t layer;           /* 0 (or 3): Layer III; 1 / 2: Layer I / II -- descriptors then carry, per 12-slot granule and channel:
                                bit_start = first bit of the frame body in the stream's blob, w1 = header bytes 1..3,
                                w2[0:2] = index of the granule inside its frame, w3[31] = state_reset_before */
    uint32_t reserved2;
} l3b_stream_desc_t;

/* Optional test taps (all device-side intermediates copied back; NULL = not wanted).
 * Layout: [granule-channel descriptor index][...]. */
typedef struct {
    int16_t* is;      /* [n_grch][576] signed quantised values (the reference's fused Huffman never materialises these) */
    uint8_t* iscf;    /* [n_grch][40]  integer scalefactors after subblock_gain / preflag (minimp3.d:694-712) */
    uint8_t* ist_pos; /* [n_grch][40]  intensity positions as left by L3_decode_scalefactors */
    /* Float stage snapshots of the bit-exact pipeline, [n_grch][576] each, written for the granules whose PCM is
     * delivered (granules wholly inside the encoder delay stay zero).  Same points as the oracle's taps: */
    float* xr;        /* after Huffman + requantisation, before stereo processing (after minimp3.d:1205) */
    float* st;        /* after MS / intensity stereo (after minimp3.d:1213) */
    float* im;        /* after reorder / alias reduction / IMDCT / frequency inversion (after minimp3.d:1229), [band][18] */
    float* dct;       /* after mp3d_DCT_II (after minimp3.d:1414), the reference's in-place layout [output j][slot] */
} l3b_taps_t;

/* l3b_batch_t.flags */
#define L3B_OUT_S16 1u     /* deliver 16-bit PCM: q = clamp(lrintf(x * 32768), -32768, 32767) of the float sample x (the
                              un-dithered float -> s16 conversion of the reference's WAV writer, wav.d:475-700); `pcm` is int16_t* */
#define L3B_MATH_FUSED 2u  /* tolerance mode: multiply-adds contracted into FMAs (faster, NOT bit-identical to the reference;
                              within 1e-5 of full scale).  Default (0) is the bit-exact pipeline. */

typedef struct {
    const uint8_t* maindata;          /* batch blob: HOST memory (copied in) */
    uint64_t maindata_bytes;
    const l3b_grch_desc_t* grch;      /* HOST */
    uint64_t n_grch;
    const l3b_stream_desc_t* streams; /* HOST */
    uint32_t n_streams;
    void* pcm;                        /* HOST destination, interleaved: float (default) or int16_t (L3B_OUT_S16); laid out by pcm_off */
    uint64_t pcm_floats;              /* samples (elements) in `pcm` */
    int32_t* status;                  /* HOST, optional [n_streams]: 0 or negative code per stream */
    const l3b_taps_t* taps;           /* optional */
    uint32_t flags;                   /* L3B_OUT_S16 | L3B_MATH_FUSED */
    uint32_t reserved;
} l3b_batch_t;

typedef struct l3b_ctx l3b_ctx_t;

/* -------- layer 1: shim ---------------------------------------------------------------------- */
int l3b_device_count(void);
int l3b_ctx_create(int device_id, l3b_ctx_t** out);
void l3b_ctx_destroy(l3b_ctx_t* ctx);
const char* l3b_last_error(const l3b_ctx_t* ctx); /* ctx may be NULL: last error of a failed create */

/* Page-locked host memory for batch inputs / PCM outputs (cudaHostAlloc): copies from and to it run at full PCIe
 * speed and asynchronously.  Returns NULL on failure. */
void* l3b_host_alloc(size_t bytes);
/* The same, with the pages placed on the NUMA node the GPU `device_id` hangs off (read from sysfs through the device's
 * PCI address; falls back to l3b_host_alloc when the topology cannot be read). */
void* l3b_host_alloc_near(int device_id, size_t bytes);
void l3b_host_free(void* p);

/* Decode a whole batch, host buffers in / host buffers out (H2D + kernels + D2H inside). */
int l3b_decode_batch(l3b_ctx_t* ctx, const l3b_batch_t* batch);

/* Device-resident variant used for throughput work: upload once, run many times, read back on demand.
 * l3b_batch_upload keeps descriptors + blob on the GPU and allocates the PCM buffer there. */
typedef struct l3b_resident l3b_resident_t;
int l3b_batch_upload(l3b_ctx_t* ctx, const l3b_batch_t* batch, l3b_resident_t** out);
/* Like l3b_batch_upload, but recycles the device buffers of *inout when they are large enough (a steady-state
 * pipeline decodes wave after wave through one workspace without cudaMalloc/cudaFree).  *inout may be NULL. */
int l3b_batch_upload_reuse(l3b_ctx_t* ctx, const l3b_batch_t* batch, l3b_resident_t** inout);
/* Refresh the inputs of an existing resident batch (same shape) from HOST memory: the per-step H2D of a
 * steady-state pipeline that reuses its device buffers. */
int l3b_batch_reupload(l3b_ctx_t* ctx, l3b_resident_t* r, const l3b_batch_t* batch);
int l3b_batch_run(l3b_ctx_t* ctx, l3b_resident_t* r);                      /* async on the context stream */
int l3b_batch_sync(l3b_ctx_t* ctx);
/* Copies samples [first, first + n) of the batch's PCM (float or int16_t elements, as uploaded) and waits for them.
 * The wait sleeps (cudaEventBlockingSync): a lane blocked on its copy does not occupy a host core. */
int l3b_batch_download(l3b_ctx_t* ctx, l3b_resident_t* r, void* pcm_host, uint64_t first, uint64_t n);
int l3b_batch_download_taps(l3b_ctx_t* ctx, l3b_resident_t* r, const l3b_taps_t* taps);
void* l3b_batch_device_pcm(l3b_resident_t* r);                              /* raw device pointer (for checksums/tests) */
void l3b_batch_free(l3b_ctx_t* ctx, l3b_resident_t* r);
/* CUDA-event timing summed over the most recent `last_runs` calls of l3b_batch_run (at most 64 are kept;
 * synchronises the context).  A run is issued as sub-batches (one by default): the entropy launches (scalefactor,
 * big_values and count1 kernels) on one stream, the granule launches on a second stream as soon as their sub-batch's
 * spectra exist.
 * ms[0] entropy launches (first start to last end), ms[1] sum of the granule launches' durations (measured in situ,
 * i.e. while sharing the GPU with later entropy launches), ms[2] whole run; *launches = kernels launched. */
int l3b_batch_timing(l3b_ctx_t* ctx, int last_runs, float ms[3], int* launches);
void* l3b_ctx_cuda_stream(l3b_ctx_t* ctx);

/* -------- layer 2: host prepass + AudioStream surface ------------------------------------------ */
typedef struct l3b_scan l3b_scan_t;

/* Frame sync + side-info parse + reservoir slicing of a whole in-memory stream, following the
 * reference's open/read loop exactly (which frames yield PCM, resets, delay/padding trim). */
int l3b_scan_memory(const uint8_t* data, size_t size, l3b_scan_t** out);
void l3b_scan_free(l3b_scan_t* s);
int l3b_scan_channels(const l3b_scan_t* s);
int l3b_scan_samplerate(const l3b_scan_t* s);
int l3b_scan_error(const l3b_scan_t* s);                 /* sticky decode error met while scanning (0 = none) */
uint64_t l3b_scan_length_frames(const l3b_scan_t* s);    /* AudioStream.getLengthInFrames (stream.d:1738) */
uint64_t l3b_scan_delivered_samples(const l3b_scan_t* s);/* interleaved samples a read-to-end delivers */
uint32_t l3b_scan_granules(const l3b_scan_t* s);
uint64_t l3b_scan_maindata_bytes(const l3b_scan_t* s);
const uint8_t* l3b_scan_maindata(const l3b_scan_t* s);
const l3b_grch_desc_t* l3b_scan_descs(const l3b_scan_t* s);
void l3b_scan_fill_stream_desc(const l3b_scan_t* s, l3b_stream_desc_t* out); /* offsets left 0 */

/* Assemble the decode program of n scanned streams -- what the D host's batch entry point hands to the shim -- into
 * caller-owned (ideally pinned) memory: `blob` receives the main data of the streams back to back, each 16-byte aligned
 * and followed by >= 16 zero bytes; `descs` the granule-channel descriptors; `streams` the stream table with its offsets
 * filled in (PCM rows 16-byte aligned).  `*batch` is filled to describe them (pcm = NULL).  Sizes: pass NULL buffers to
 * get the required blob bytes / descriptor count / PCM floats in `*batch` without copying anything.
 * Returns L3B_E_PARAM when a capacity is too small. */
int l3b_scans_assemble(l3b_scan_t* const* scans, uint32_t n, uint8_t* blob, uint64_t blob_cap, l3b_grch_desc_t* descs,
                       uint64_t desc_cap, l3b_stream_desc_t* streams, l3b_batch_t* batch);

/* Batch entry point the D host adds next to AudioStream: decode n scanned streams in one go.
 * pcm[i] must have room for l3b_scan_delivered_samples(scans[i]) floats. */
int l3b_decode_scans(l3b_ctx_t* ctx, l3b_scan_t* const* scans, uint32_t n, float* const* pcm, int32_t* status);

/* Batch entry point over RAW streams, pipelined in waves over one or more GPUs (l3_pipeline.cpp): MP3 bytes in host memory
 * in, PCM in host memory out (float, or int16_t with L3B_OUT_S16).  Per GPU `lanes` contexts (CUDA stream + recycled device
 * workspace + pinned staging each) run assemble -> H2D -> kernels -> D2H for their waves of `wave_streams` streams while
 * `scan_threads` host threads run the prepass (l3b_scan_memory) ahead of them.  With several devices the streams are assigned
 * by file, longest first, with no collective on the data path (streams are independent: minimp3.d:38-46).
 * A stream that cannot be decoded gets its status and zero frames; the others are unaffected. */
typedef struct l3b_pipeline l3b_pipeline_t;
typedef struct {
    int32_t lanes;         /* per device; <= 0: 4 */
    int32_t wave_streams;  /* <= 0: 16 */
    int32_t scan_threads;  /* <= 0: host CPUs available to the process minus one per device */
    uint32_t flags;        /* L3B_OUT_S16 | L3B_MATH_FUSED */
} l3b_pipeline_opts_t;
typedef struct {
    uint64_t pcm_off;      /* element offset of the stream's first delivered sample in the output buffer */
    uint64_t frames;       /* frames delivered (samples per channel) */
    int32_t channels, samplerate;
    int32_t status;        /* 0, or the negative code that made the stream undecodable / stopped it early */
    int32_t device;        /* position in device_ids[] of the GPU that decoded it */
} l3b_stream_result_t;
#define L3B_PIPELINE_PHASES 6
int l3b_pipeline_create(const int* device_ids, int n_devices, const l3b_pipeline_opts_t* opts, l3b_pipeline_t** out);
void l3b_pipeline_destroy(l3b_pipeline_t* p);
/* `out` should be page-locked (l3b_host_alloc_near) and hold `out_capacity` elements; every wave's PCM starts 16-byte aligned,
 * so allow 8 elements of slack per wave.  *out_used (optional) receives the elements handed out. */
int l3b_pipeline_decode(l3b_pipeline_t* p, const uint8_t* const* data, const size_t* size, uint32_t n, void* out,
                        uint64_t out_capacity, l3b_stream_result_t* results, uint64_t* out_used);
/* Seconds per phase summed over threads since the last call: [0] host prepass (scan threads), then per lane [1] assemble,
 * [2] upload, [3] launch, [4] PCM download incl. waiting for the kernels, [5] waiting for the prepass. */
int l3b_pipeline_profile(l3b_pipeline_t* p, double seconds[L3B_PIPELINE_PHASES]);
const char* l3b_pipeline_last_error(const l3b_pipeline_t* p);

/* Batch entry point over RAW Layer III files with the prepass ON THE GPU (l3_raw.cu): frame walk, side-info parse, reservoir
 * recurrence and main-data gathering run as kernels for well-formed streams (a clean chain of compatible Layer III frames
 * between the tags, no Xing / Info tag, legal side info, not free-format); every other stream takes the host prepass
 * (l3b_scan_memory) inside the same call, so the PCM is the host route's in all cases.  Two steps because the output sizes
 * are only known after the prepass.  flags: L3B_OUT_S16 | L3B_MATH_FUSED. */
typedef struct l3b_raw l3b_raw_t;
int l3b_raw_open(l3b_ctx_t* ctx, const uint8_t* const* data, const size_t* size, uint32_t n, uint32_t flags, l3b_raw_t** out);
uint32_t l3b_raw_device_streams(const l3b_raw_t* r);            /* streams whose prepass ran on the GPU */
float l3b_raw_prepass_ms(const l3b_raw_t* r);                   /* device time of the three prepass kernels */
int l3b_raw_channels(const l3b_raw_t* r, uint32_t i);
int l3b_raw_samplerate(const l3b_raw_t* r, uint32_t i);
uint64_t l3b_raw_samples(const l3b_raw_t* r, uint32_t i);       /* interleaved samples stream i delivers */
int l3b_raw_status(const l3b_raw_t* r, uint32_t i);             /* 0, or the code that makes stream i undecodable */
int l3b_raw_decode(l3b_raw_t* r, void* const* pcm);             /* pcm[i]: room for l3b_raw_samples(r, i) elements (may be NULL) */
void l3b_raw_free(l3b_raw_t* r);

/* AudioStream mirror (names follow stream.d). */
typedef struct l3b_stream l3b_stream_t;
int l3b_stream_open_memory(l3b_ctx_t* ctx, const uint8_t* data, size_t size, l3b_stream_t** out); /* stream.d:150 (copies input) */
int l3b_stream_open_file(l3b_ctx_t* ctx, const char* path, l3b_stream_t** out);                   /* stream.d:115 */
/* The reference's callback I/O (mp3dec_io_t, minimp3_ex.d:61-71; stream.d:2243-2254 wires IOCallbacks onto it): `read`
 * returns the bytes read (short = end of input), `seek` 0 on success.  The callbacks are drained into memory inside the call. */
typedef size_t (*l3b_read_cb)(void* buf, size_t size, void* user);
typedef int (*l3b_seek_cb)(uint64_t position, void* user);
int l3b_stream_open_callbacks(l3b_ctx_t* ctx, l3b_read_cb read, l3b_seek_cb seek, void* user, l3b_stream_t** out); /* mp3dec_ex_open_cb, minimp3_ex.d:929 */
void l3b_stream_close(l3b_stream_t* s);
int l3b_stream_num_channels(const l3b_stream_t* s);      /* stream.d:396 */
int64_t l3b_stream_length_frames(const l3b_stream_t* s); /* stream.d:402 */
float l3b_stream_samplerate(const l3b_stream_t* s);      /* stream.d:412 */
int l3b_stream_read_float(l3b_stream_t* s, float* out, int frames);   /* stream.d:429: frames read; 0 on error (see is_error) */
int l3b_stream_read_double(l3b_stream_t* s, double* out, int frames); /* stream.d:656 */
int l3b_stream_seek(l3b_stream_t* s, int frame);         /* stream.d:1095: 1 = ok, 0 = refused */
int l3b_stream_tell(const l3b_stream_t* s);              /* stream.d:1209 */
int l3b_stream_is_error(const l3b_stream_t* s);          /* stream.d:295 */
const char* l3b_stream_error_message(const l3b_stream_t* s); /* stream.d:316 */

#ifdef __cplusplus
}
#endif
#endif
