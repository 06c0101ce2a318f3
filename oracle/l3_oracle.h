/* l3_oracle.h -- CPU ORACLE for the MPEG-1/2 Layer III decode path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of /root/reference/source/audioformats/minimp3.d (Layer III branch) and
 * of the stream layer minimp3_ex.d (callback I/O variant, the one stream.d uses).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (audio_formats_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors for MP3 and no D compiler exists in this
 * image, so this oracle could not be checked against a run of the reference itself.  It is pinned
 * by (a) following the D source statement by statement (file:line cited at each function),
 * including D's default initialisation of locals, which the reference relies on,
 * (b) structural checks of the recovered Huffman books (tools/derive_tables.py),
 * (c) encoder->oracle round trips of the synthetic generator (tests/), and
 * (d) an INDEPENDENT decoder: FFmpeg's mp3float (libavcodec inside the image) agrees with it to 2e-6 of
 *     full scale on every format the generator writes (tests/test_oracle_vs_ffmpeg.py; the three places where
 *     the two decoders legitimately differ are listed there and in DESIGN.md section 7).
 */
#ifndef L3_ORACLE_H
#define L3_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define L3O_MAX_SAMPLES_PER_FRAME (1152 * 2)

/* minimp3.d:38-46 */
typedef struct {
    float mdct_overlap[2][9 * 32];
    float qmf_state[15 * 2 * 32];
    int reserv;
    int free_format_bytes;
    uint8_t header[4];
    uint8_t reserv_buf[511];
} l3o_dec_t;

/* minimp3.d:28-36 */
typedef struct {
    int frame_bytes, frame_offset, channels, hz, layer, bitrate_kbps;
} l3o_frame_info_t;

/* One record per decoded granule (both channels).  The reference fuses Huffman decode and
 * requantisation, so the signed quantised integers are recorded at the points named in SURVEY 8c. */
typedef struct {
    int16_t is[2][576];      /* +-(lsb) after linbits add (minimp3.d:805-819,843-848), +-1 for count1 (:874-878) */
    uint8_t iscf[2][40];     /* final integer scalefactors (minimp3.d:694-712) */
    uint8_t ist_pos[2][40];  /* ist_pos after L3_decode_scalefactors (before the intensity fix-up) */
    int32_t gain_exp[2];     /* minimp3.d:714 */
    float scf[2][40];        /* float band gains (minimp3.d:716-719) */
    float xr[2][576];        /* after Huffman+requant, before stereo (after minimp3.d:1205) */
    float st[2][576];        /* after stereo processing (after :1213) */
    float im[2][576];        /* after reorder/antialias/IMDCT/change_sign (after :1229) */
    float dct[2][576];       /* after mp3d_DCT_II (after :1414) */
} l3o_granule_tap_t;

typedef struct {
    l3o_granule_tap_t* rec;  /* caller-provided array */
    int capacity;
    int count;               /* granules seen (may exceed capacity; only the first `capacity` are stored) */
} l3o_tap_t;

/* Install (or clear with NULL) the tap sink used by l3o_decode_frame on THIS thread. */
void l3o_set_tap(l3o_tap_t* tap);

/* Per-stage wall-clock accumulation (seconds), enabled with l3o_enable_timers(1). Order:
 * 0 side-info+reservoir, 1 scalefactors+huffman, 2 stereo, 3 reorder+antialias, 4 imdct, 5 dct32, 6 window */
void l3o_enable_timers(int on);
void l3o_get_timers(double out[7]);

void l3o_init(l3o_dec_t* dec);                                                   /* minimp3.d:1487 */
int l3o_decode_frame(l3o_dec_t* dec, const uint8_t* mp3, int mp3_bytes, float* pcm,
                     l3o_frame_info_t* info);                                    /* minimp3.d:1492 */

/* header helpers (minimp3.d:232-283), exported for tests */
int l3o_hdr_valid(const uint8_t* h);
int l3o_hdr_frame_bytes(const uint8_t* h, int free_format_size);
int l3o_hdr_padding(const uint8_t* h);
unsigned l3o_hdr_sample_rate_hz(const uint8_t* h);
unsigned l3o_hdr_frame_samples(const uint8_t* h);
unsigned l3o_hdr_bitrate_kbps(const uint8_t* h);

/* ---- stream layer (minimp3_ex.d) ---- */
#define L3O_E_PARAM (-1)
#define L3O_E_MEMORY (-2)
#define L3O_E_IOERROR (-3)
#define L3O_E_USER (-4)
#define L3O_E_DECODE (-5)
#define L3O_SEEK_TO_BYTE 0
#define L3O_SEEK_TO_SAMPLE 1

typedef size_t (*l3o_read_cb)(void* buf, size_t size, void* user);
typedef int (*l3o_seek_cb)(uint64_t position, void* user);
typedef struct {
    l3o_read_cb read;
    void* read_data;
    l3o_seek_cb seek;
    void* seek_data;
} l3o_io_t;

typedef struct { uint64_t sample, offset; } l3o_index_frame_t;

/* minimp3_ex.d:73-87 */
typedef struct {
    l3o_dec_t mp3d;
    const uint8_t* file_buffer;
    size_t file_size;
    l3o_io_t* io;
    l3o_index_frame_t* frames;
    size_t num_frames, capacity;
    uint64_t offset, samples, detected_samples, cur_sample, start_offset, end_offset;
    l3o_frame_info_t info;
    float buffer[L3O_MAX_SAMPLES_PER_FRAME];
    size_t input_consumed, input_filled;
    int is_file, seek_method, vbr_tag_found;
    int free_format_bytes;
    int buffer_samples, buffer_consumed, to_skip, start_delay;
    int last_error;
} l3o_ex_t;

int l3o_detect_cb(l3o_io_t* io, uint8_t* buf, size_t buf_size);                 /* minimp3_ex.d:197 */
int l3o_ex_open_cb(l3o_ex_t* dec, l3o_io_t* io, int seek_method);               /* minimp3_ex.d:929 */
size_t l3o_ex_read(l3o_ex_t* dec, float* buf, size_t samples);                  /* minimp3_ex.d:787 */
int l3o_ex_seek(l3o_ex_t* dec, uint64_t position);                              /* minimp3_ex.d:662 */
void l3o_ex_close(l3o_ex_t* dec);                                               /* minimp3_ex.d:953 */

/* ---- convenience for ctypes: an AudioStream-shaped handle over a memory buffer, wired the way
 * stream.d:1706-1749 wires it (detect, then open_cb with SEEK_TO_SAMPLE, MemoryContext I/O) ---- */
typedef struct l3o_stream l3o_stream_t;
l3o_stream_t* l3o_stream_open_memory(const uint8_t* data, size_t size);   /* NULL if not detected as MP3 */
void l3o_stream_close(l3o_stream_t* s);
int l3o_stream_channels(const l3o_stream_t* s);
int l3o_stream_samplerate(const l3o_stream_t* s);
long long l3o_stream_length_frames(const l3o_stream_t* s);
int l3o_stream_read_float(l3o_stream_t* s, float* out, int frames);      /* stream.d:537-551 */
int l3o_stream_seek(l3o_stream_t* s, int frame);                         /* stream.d:1100-1107: 1 = ok, 0 = refused */
int l3o_stream_tell(const l3o_stream_t* s);                              /* stream.d:1214-1218 */
int l3o_stream_last_error(const l3o_stream_t* s);
/* examples/transcode/source/main.d:52-78 loop shape; see l3_oracle_ex.c */
long long l3o_transcode_loop(const uint8_t* data, size_t size, int chunk_frames, float* out, size_t cap_samples,
                             int* channels_out, int* hz_out);

/* the same with the un-dithered float -> s16 conversion per chunk (wav.d:475-700) */
long long l3o_transcode_loop_s16(const uint8_t* data, size_t size, int chunk_frames, int16_t* out16, size_t cap_samples,
                                 int* channels_out, int* hz_out);

#ifdef __cplusplus
}
#endif
#endif
