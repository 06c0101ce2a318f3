/* l3synth.h -- deterministic synthetic Layer III bitstream generator (bench/test infrastructure). */
#ifndef L3SYNTH_H
#define L3SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t seed;
    int hz;            /* 32000/44100/48000 (MPEG-1), 16000/22050/24000 (MPEG-2 LSF), 8000/11025/12000 (MPEG-2.5) */
    int nch;           /* 1 or 2 */
    int bitrate_kbps;  /* a legal Layer III rate for that version */
    int nframes;
    int block_mode;    /* 0: long blocks only; 1: long->start->short(xN, 1/3 of the runs mixed)->stop sequences; 2: same, never mixed */
    int stereo_mode;   /* 0: plain stereo; 1: joint with MS on ~half the frames; 2: joint with MS and intensity */
    int reservoir;     /* 0: main_data_begin always 0; 1: moderate; 2: heavy (sparse/dense alternation) */
    int scfsi;         /* 1: exercise scfsi on granule 1 */
    int crc;           /* 1: protection bit cleared, 16-bit CRC field present (never verified by the reference) */
    int escapes;       /* 1: use linbits tables 16..31 and escape values */
    int gain_base;     /* global_gain centre (+-4 jitter) */
    double level;      /* mean quantised magnitude at DC (geometric, tilted down with frequency) */
    int small_scalefactors; /* 1: restrict scalefac_compress so band attenuation stays small */
    int table_cycle;   /* 1: walk table_select through all books deterministically */
    int table_cycle_pos;
    int no_padding;    /* 1: never set the padding bit */
    int id3v2_bytes;   /* >= 10: prepend an ID3v2 tag of that total size */
    int id3v1;         /* 1: append a 128-byte ID3v1 tag */
    int emphasis_bits; /* low 4 bits of header byte 3 (copyright/original/emphasis) */
    int mixed_only_short; /* 1: mixed_block_flag only on the short blocks of a mixed run, not on its start/stop blocks */
    int free_format;   /* 1: write bitrate index 0 (free format); bitrate_kbps may then be any value (frame <= 2304 bytes) */
    int vbr;           /* 1: the bitrate index and the padding bit change from frame to frame (around bitrate_kbps) */
    int mode_ext_any;  /* 1: random mode_extension bits in frames that are not joint stereo (mono; stereo needs stereo_mode >= 2) */
    int istereo_untied; /* 1: with intensity stereo the two channels still choose their block types independently (the
                         * reference then reads ist_pos entries channel 1 did not transmit: zero in granule 0, granule 0's
                         * leftovers in granule 1 -- defined in D, whose locals start zeroed) */
    int private_bits;  /* 1: random private bits (MPEG-1: the reference takes them for granule 0's scfsi; the stream is written the way the reference reads it) */
} l3s_params_t;

typedef struct {
    int frames, granules, samples_per_frame, mpeg1, sr_idx;
    long long bytes;
} l3s_info_t;

size_t l3s_max_bytes(const l3s_params_t* p);
/* Returns bytes written (<0 on error: -1 bad params, -2 internal overrun, -3 cap too small).
 * is_out (optional): [granule][ch][576] signed quantised values in decode order. */
long long l3s_generate(const l3s_params_t* p, uint8_t* out, size_t cap, int16_t* is_out, l3s_info_t* info_out);

#ifdef __cplusplus
}
#endif
#endif
