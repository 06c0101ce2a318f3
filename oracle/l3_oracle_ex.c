/* l3_oracle_ex.c -- CPU ORACLE, stream layer (test infrastructure only; see l3_oracle.h).
 *
 * Restates the callback-I/O variant of /root/reference/source/audioformats/minimp3_ex.d (the one
 * stream.d:1706-1749 uses): ID3/APE skipping, detection, VBR-tag probe, whole-file index, sample
 * accurate seek with reservoir pre-roll, and the buffered read loop.  PARITY UNPINNED (see header).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "l3_oracle.h"

#define IO_SIZE (128 * 1024) /* minimp3_ex.d:26 */
#define BUF_SIZE (16 * 1024) /* minimp3_ex.d:27 */
#define PREDECODE_FRAMES 2   /* minimp3_ex.d:24 */
#define ID3_DETECT_SIZE 10   /* minimp3_ex.d:113 */
#define HDR_SIZE 4

int l3o__side_info_main_bytes(const uint8_t* hdr, int frame_size, int* main_data_begin_out);
int l3o__side_info_end_byte(const uint8_t* hdr, int frame_size);
int l3o__find_frame(const uint8_t* mp3, int mp3_bytes, int* free_format_bytes, int* ptr_frame_bytes);

static size_t zmin(size_t a, size_t b) { return a > b ? b : a; }

/* minimp3_ex.d:93-111 */
static void skip_id3v1(const uint8_t* buf, size_t* pbuf_size)
{
    size_t buf_size = *pbuf_size;
    if (buf_size >= 128 && !memcmp(buf + buf_size - 128, "TAG", 3)) {
        buf_size -= 128;
        if (buf_size >= 227 && !memcmp(buf + buf_size - 227, "TAG+", 4)) buf_size -= 227;
    }
    if (buf_size > 32 && !memcmp(buf + buf_size - 32, "APETAGEX", 8)) {
        buf_size -= 32;
        const uint8_t* tag = buf + buf_size + 8 + 4;
        uint32_t tag_size = (uint32_t)(tag[3] << 24) | (tag[2] << 16) | (tag[1] << 8) | tag[0];
        if (buf_size >= tag_size) buf_size -= tag_size;
    }
    *pbuf_size = buf_size;
}

/* minimp3_ex.d:115-125 */
static size_t skip_id3v2(const uint8_t* buf, size_t buf_size)
{
    if (buf_size >= ID3_DETECT_SIZE && !memcmp(buf, "ID3", 3) &&
        !((buf[5] & 15) || (buf[6] & 0x80) || (buf[7] & 0x80) || (buf[8] & 0x80) || (buf[9] & 0x80))) {
        size_t id3v2size = (((buf[6] & 0x7f) << 21) | ((buf[7] & 0x7f) << 14) | ((buf[8] & 0x7f) << 7) | (buf[9] & 0x7f)) + 10;
        if ((buf[5] & 16)) id3v2size += 10; /* footer */
        return id3v2size;
    }
    return 0;
}

/* minimp3_ex.d:144-190 */
static int check_vbrtag(const uint8_t* frame, int frame_size, uint32_t* frames, int* delay, int* padding)
{
    enum { FRAMES_FLAG = 1, BYTES_FLAG = 2, TOC_FLAG = 4, VBR_SCALE_FLAG = 8 };
    int side_end = l3o__side_info_end_byte(frame, frame_size);
    if (side_end < 0) return 0; /* side info corrupted */

    const uint8_t* tag = frame + side_end;
    if (memcmp("Xing", tag, 4) && memcmp("Info", tag, 4)) return 0;
    int flags = tag[7];
    if (!(flags & FRAMES_FLAG)) return -1;
    tag += 8;
    *frames = (uint32_t)(tag[0] << 24) | (tag[1] << 16) | (tag[2] << 8) | tag[3];
    tag += 4;
    if (flags & BYTES_FLAG) tag += 4;
    if (flags & TOC_FLAG) tag += 100;
    if (flags & VBR_SCALE_FLAG) tag += 4;
    *delay = *padding = 0;
    if (*tag) { /* extension, LAME, Lavc, etc. */
        tag += 21;
        if (tag - frame + 14 >= frame_size) return 0;
        *delay = ((tag[0] << 4) | (tag[1] >> 4)) + (528 + 1);
        *padding = (((tag[1] & 0xF) << 8) | tag[2]) - (528 + 1);
    }
    return 1;
}

/* minimp3_ex.d:197-233 */
int l3o_detect_cb(l3o_io_t* io, uint8_t* buf, size_t buf_size)
{
    if (!buf || (size_t)-1 == buf_size || (io && buf_size < BUF_SIZE)) return L3O_E_PARAM;
    size_t filled = buf_size;
    if (io) {
        if (io->seek(0, io->seek_data)) return L3O_E_IOERROR;
        filled = io->read(buf, ID3_DETECT_SIZE, io->read_data);
        if (filled > ID3_DETECT_SIZE) return L3O_E_IOERROR;
    }
    if (filled < ID3_DETECT_SIZE) return L3O_E_USER; /* too small, can't be mp3/mpa */
    if (skip_id3v2(buf, filled)) return 0;           /* id3v2 tag is enough evidence */
    if (io) {
        size_t readed = io->read(buf + ID3_DETECT_SIZE, buf_size - ID3_DETECT_SIZE, io->read_data);
        if (readed > (buf_size - ID3_DETECT_SIZE)) return L3O_E_IOERROR;
        filled += readed;
        if (filled < BUF_SIZE) skip_id3v1(buf, &filled);
    } else {
        skip_id3v1(buf, &filled);
        if (filled > BUF_SIZE) filled = BUF_SIZE;
    }
    int free_format_bytes = 0, frame_size = 0; /* D default-initialises locals (int.init == 0): minimp3_ex.d:228 relies on it */
    l3o__find_frame(buf, (int)filled, &free_format_bytes, &frame_size);
    if (frame_size) return 0; /* MAX_FRAME_SYNC_MATCHES consecutive frames found */
    return L3O_E_USER;
}

typedef int (*iterate_cb)(void* user_data, const uint8_t* frame, int frame_size, int free_format_bytes,
                          size_t buf_size, uint64_t offset, l3o_frame_info_t* info);

static void fill_info(l3o_frame_info_t* fi, const uint8_t* hdr, int frame_size)
{
    fi->channels = ((hdr[3] & 0xC0) == 0xC0) ? 1 : 2;
    fi->hz = l3o_hdr_sample_rate_hz(hdr);
    fi->layer = 4 - ((hdr[1] >> 1) & 3);
    fi->bitrate_kbps = l3o_hdr_bitrate_kbps(hdr);
    fi->frame_bytes = frame_size;
}

/* minimp3_ex.d:490-564 */
static int iterate_io(l3o_io_t* io, uint8_t* buf, size_t buf_size, iterate_cb callback, void* user_data)
{
    if (!io || !buf || (size_t)-1 == buf_size || buf_size < BUF_SIZE || !callback) return L3O_E_PARAM;
    size_t filled = io->read(buf, ID3_DETECT_SIZE, io->read_data), consumed = 0;
    uint64_t readed2 = 0;
    l3o_frame_info_t frame_info;
    int eof = 0;
    memset(&frame_info, 0, sizeof frame_info);
    if (filled > ID3_DETECT_SIZE) return L3O_E_IOERROR;
    if (ID3_DETECT_SIZE != filled) return 0;
    size_t id3v2size = skip_id3v2(buf, filled);
    if (id3v2size) {
        if (io->seek(id3v2size, io->seek_data)) return L3O_E_IOERROR;
        filled = io->read(buf, buf_size, io->read_data);
        if (filled > buf_size) return L3O_E_IOERROR;
        readed2 += id3v2size;
    } else {
        size_t readed = io->read(buf + ID3_DETECT_SIZE, buf_size - ID3_DETECT_SIZE, io->read_data);
        if (readed > (buf_size - ID3_DETECT_SIZE)) return L3O_E_IOERROR;
        filled += readed;
    }
    if (filled < BUF_SIZE) skip_id3v1(buf, &filled);
    do {
        int free_format_bytes = 0, frame_size = 0, ret;
        int i = l3o__find_frame(buf + consumed, (int)(filled - consumed), &free_format_bytes, &frame_size);
        if (i && !frame_size) {
            consumed += i;
            continue;
        }
        if (!frame_size) break;
        const uint8_t* hdr = buf + consumed + i;
        fill_info(&frame_info, hdr, frame_size);

        readed2 += i;
        ret = callback(user_data, hdr, frame_size, free_format_bytes, filled - consumed, readed2, &frame_info);
        if (ret) return ret;
        readed2 += frame_size;
        consumed += i + frame_size;
        if (!eof && filled - consumed < BUF_SIZE) { /* keep minimum 10 consecutive mp3 frames (~16KB) worst case */
            memmove(buf, buf + consumed, filled - consumed);
            filled -= consumed;
            consumed = 0;
            size_t readed = io->read(buf + filled, buf_size - filled, io->read_data);
            if (readed > (buf_size - filled)) return L3O_E_IOERROR;
            if (readed != (buf_size - filled)) eof = 1;
            filled += readed;
            if (eof) skip_id3v1(buf, &filled);
        }
    } while (1);
    return 0;
}

/* minimp3_ex.d:566-621 */
static int load_index(void* user_data, const uint8_t* frame, int frame_size, int free_format_bytes, size_t buf_size,
                      uint64_t offset, l3o_frame_info_t* info)
{
    l3o_ex_t* dec = (l3o_ex_t*)user_data;
    if (!dec->frames && !dec->start_offset) { /* detect VBR tag and try to avoid full scan */
        uint32_t frames;
        int delay, padding;
        dec->info = *info;
        dec->start_offset = dec->offset = offset;
        dec->end_offset = offset + buf_size;
        dec->free_format_bytes = free_format_bytes; /* should not change */
        if (3 == dec->info.layer) {
            int ret = check_vbrtag(frame, frame_size, &frames, &delay, &padding);
            if (ret) dec->start_offset = dec->offset = offset + frame_size;
            if (ret > 0) {
                padding *= info->channels;
                dec->start_delay = dec->to_skip = delay * info->channels;
                dec->samples = l3o_hdr_frame_samples(frame) * info->channels * (uint64_t)frames;
                if (dec->samples >= (uint64_t)dec->start_delay) dec->samples -= dec->start_delay;
                if (padding > 0 && dec->samples >= (uint64_t)padding) dec->samples -= padding;
                dec->detected_samples = dec->samples;
                dec->vbr_tag_found = 1;
                return L3O_E_USER;
            } else if (ret < 0)
                return 0;
        }
    }
    if (dec->num_frames + 1 > dec->capacity) {
        if (!dec->capacity)
            dec->capacity = 4096;
        else
            dec->capacity *= 2;
        l3o_index_frame_t* alloc_buf = (l3o_index_frame_t*)realloc(dec->frames, sizeof(l3o_index_frame_t) * dec->capacity);
        if (!alloc_buf) return L3O_E_MEMORY;
        dec->frames = alloc_buf;
    }
    l3o_index_frame_t* idx_frame = &dec->frames[dec->num_frames++];
    idx_frame->offset = offset;
    idx_frame->sample = dec->samples;
    if (!dec->buffer_samples && dec->num_frames < 256) {
        /* try to decode up to 255 first frames till samples start to decode */
        dec->buffer_samples = l3o_decode_frame(&dec->mp3d, frame, (int)zmin(buf_size, (size_t)0x7fffffff), dec->buffer, info);
        dec->samples += dec->buffer_samples * info->channels;
    } else
        dec->samples += l3o_hdr_frame_samples(frame) * info->channels;
    return 0;
}

/* minimp3_ex.d:640-660 */
static size_t idx_binary_search(l3o_ex_t* idx, uint64_t position)
{
    size_t end = idx->num_frames, start = 0, index = 0;
    while (start <= end) {
        size_t mid = (start + end) / 2;
        if (idx->frames[mid].sample >= position) { /* move left side. */
            if (idx->frames[mid].sample == position) return mid;
            end = mid - 1;
        } else { /* move to right side */
            index = mid;
            start = mid + 1;
            if (start == idx->num_frames) break;
        }
    }
    return index;
}

/* minimp3_ex.d:662-785 (callback-I/O arm; dec->io is always set here) */
int l3o_ex_seek(l3o_ex_t* dec, uint64_t position)
{
    size_t i;
    if (!dec) return L3O_E_PARAM;
    if (L3O_SEEK_TO_BYTE == dec->seek_method) {
        dec->offset = position;
        dec->cur_sample = 0;
        goto do_exit;
    }
    dec->cur_sample = position;
    position += dec->start_delay;
    if (0 == position) { /* optimize seek to zero, no index needed */
    seek_zero:
        dec->offset = dec->start_offset;
        dec->to_skip = 0;
        goto do_exit;
    }
    if (!dec->frames && dec->vbr_tag_found) { /* no index created yet (vbr tag used to calculate track length) */
        dec->samples = 0;
        dec->buffer_samples = 0;
        if (dec->io->seek(dec->start_offset, dec->io->seek_data)) return L3O_E_IOERROR;
        int ret = iterate_io(dec->io, (uint8_t*)dec->file_buffer, dec->file_size, &load_index, dec);
        if (ret && L3O_E_USER != ret) return ret;
        for (i = 0; i < dec->num_frames; i++) dec->frames[i].offset += dec->start_offset;
        dec->samples = dec->detected_samples;
    }
    if (!dec->frames) goto seek_zero; /* no frames in file - seek to zero */
    i = idx_binary_search(dec, position);
    if (i) {
        int to_fill_bytes = 511;
        int skip_frames = PREDECODE_FRAMES;
        i -= zmin(i, (size_t)skip_frames);
        if (3 == dec->info.layer) {
            while (i && to_fill_bytes) { /* make sure bit-reservoir is filled when we start decoding */
                int frame_bytes, frame_size;
                uint8_t* hdr = (uint8_t*)dec->file_buffer;
                if (dec->io->seek(dec->frames[i - 1].offset, dec->io->seek_data)) return L3O_E_IOERROR;
                size_t readed = dec->io->read(hdr, HDR_SIZE, dec->io->read_data);
                if (readed != HDR_SIZE) return L3O_E_IOERROR;
                frame_size = l3o_hdr_frame_bytes(hdr, dec->free_format_bytes) + l3o_hdr_padding(hdr);
                readed = dec->io->read(hdr + HDR_SIZE, frame_size - HDR_SIZE, dec->io->read_data);
                if (readed != (size_t)(frame_size - HDR_SIZE)) return L3O_E_IOERROR;
                i--;
                frame_bytes = l3o__side_info_main_bytes(hdr, frame_size, NULL);
                if (frame_bytes < 0) break; /* frame not decodable, we can start from here */
                to_fill_bytes -= (to_fill_bytes < frame_bytes) ? to_fill_bytes : frame_bytes;
            }
        }
    }
    dec->offset = dec->frames[i].offset;
    dec->to_skip = (int)(position - dec->frames[i].sample);
    while ((i + 1) < dec->num_frames && !dec->frames[i].sample && !dec->frames[i + 1].sample) {
        /* skip not decodable first frames */
        uint8_t* hdr = (uint8_t*)dec->file_buffer;
        if (dec->io->seek(dec->frames[i].offset, dec->io->seek_data)) return L3O_E_IOERROR;
        size_t readed = dec->io->read(hdr, HDR_SIZE, dec->io->read_data);
        if (readed != HDR_SIZE) return L3O_E_IOERROR;
        dec->to_skip += l3o_hdr_frame_samples(hdr) * dec->info.channels;
        i++;
    }
do_exit:
    if (dec->io->seek(dec->offset, dec->io->seek_data)) return L3O_E_IOERROR;
    dec->buffer_samples = 0;
    dec->buffer_consumed = 0;
    dec->input_consumed = 0;
    dec->input_filled = 0;
    dec->last_error = 0;
    l3o_init(&dec->mp3d);
    return 0;
}

/* minimp3_ex.d:787-888 (callback-I/O arm) */
size_t l3o_ex_read(l3o_ex_t* dec, float* buf, size_t samples)
{
    if (!dec || !buf) return (size_t)L3O_E_PARAM;
    size_t samples_requested = samples;
    int eof = 0;
    l3o_frame_info_t frame_info;
    memset(&frame_info, 0, sizeof frame_info);
    if (dec->detected_samples && dec->cur_sample >= dec->detected_samples) return 0; /* at end of stream */
    if (dec->last_error) return 0; /* error eof state, seek can reset it */
    if (dec->buffer_consumed < dec->buffer_samples) {
        size_t to_copy = zmin((size_t)(dec->buffer_samples - dec->buffer_consumed), samples);
        if (dec->detected_samples) { /* count decoded samples to properly cut padding */
            if (dec->cur_sample + to_copy >= dec->detected_samples) to_copy = (size_t)(dec->detected_samples - dec->cur_sample);
        }
        dec->cur_sample += to_copy;
        memcpy(buf, dec->buffer + dec->buffer_consumed, to_copy * sizeof(float));
        buf += to_copy;
        dec->buffer_consumed += to_copy;
        samples -= to_copy;
    }
    while (samples) {
        if (dec->detected_samples && dec->cur_sample >= dec->detected_samples) break;
        const uint8_t* dec_buf;
        if (!eof && (dec->input_filled - dec->input_consumed) < BUF_SIZE) {
            /* keep minimum 10 consecutive mp3 frames (~16KB) worst case */
            uint8_t* fb = (uint8_t*)dec->file_buffer;
            memmove(fb, fb + dec->input_consumed, dec->input_filled - dec->input_consumed);
            dec->input_filled -= dec->input_consumed;
            dec->input_consumed = 0;
            size_t readed = dec->io->read(fb + dec->input_filled, dec->file_size - dec->input_filled, dec->io->read_data);
            if (readed > (dec->file_size - dec->input_filled)) {
                dec->last_error = L3O_E_IOERROR;
                readed = 0;
            }
            if (readed != (dec->file_size - dec->input_filled)) eof = 1;
            dec->input_filled += readed;
            if (eof) skip_id3v1(fb, &dec->input_filled);
        }
        dec_buf = dec->file_buffer + dec->input_consumed;
        if (!(dec->input_filled - dec->input_consumed)) break;
        dec->buffer_samples = l3o_decode_frame(&dec->mp3d, dec_buf, (int)(dec->input_filled - dec->input_consumed),
                                               dec->buffer, &frame_info);
        dec->input_consumed += frame_info.frame_bytes;
        dec->buffer_consumed = 0;
        if (dec->info.hz != frame_info.hz || dec->info.layer != frame_info.layer || dec->info.channels != frame_info.channels) {
            dec->last_error = L3O_E_DECODE;
            break;
        }
        if (dec->buffer_samples) {
            dec->buffer_samples *= frame_info.channels;
            if (dec->to_skip) {
                size_t skip = zmin((size_t)dec->buffer_samples, (size_t)dec->to_skip);
                dec->buffer_consumed += skip;
                dec->to_skip -= skip;
            }
            size_t to_copy = zmin((size_t)(dec->buffer_samples - dec->buffer_consumed), samples);
            if (dec->detected_samples) { /* ^ handle padding */
                if (dec->cur_sample + to_copy >= dec->detected_samples) to_copy = (size_t)(dec->detected_samples - dec->cur_sample);
            }
            dec->cur_sample += to_copy;
            memcpy(buf, dec->buffer + dec->buffer_consumed, to_copy * sizeof(float));
            buf += to_copy;
            dec->buffer_consumed += to_copy;
            samples -= to_copy;
        } else if (dec->to_skip) {
            /* frames that cannot decode because of the bit reservoir still count against to_skip */
            int frame_samples = l3o_hdr_frame_samples(dec_buf) * frame_info.channels;
            dec->to_skip -= (frame_samples < dec->to_skip) ? frame_samples : dec->to_skip;
        }
        dec->offset += frame_info.frame_bytes;
    }
    return samples_requested - samples;
}

/* minimp3_ex.d:929-951 */
int l3o_ex_open_cb(l3o_ex_t* dec, l3o_io_t* io, int seek_method)
{
    if (!dec || !io || !(L3O_SEEK_TO_BYTE == seek_method || L3O_SEEK_TO_SAMPLE == seek_method)) return L3O_E_PARAM;
    memset(dec, 0, sizeof *dec);
    dec->file_size = IO_SIZE;
    dec->file_buffer = (const uint8_t*)malloc(dec->file_size);
    if (!dec->file_buffer) return L3O_E_MEMORY;
    dec->seek_method = seek_method;
    dec->io = io;
    l3o_init(&dec->mp3d);
    if (io->seek(0, io->seek_data)) return L3O_E_IOERROR;
    int ret = iterate_io(io, (uint8_t*)dec->file_buffer, dec->file_size, &load_index, dec);
    if (ret && L3O_E_USER != ret) return ret;
    if (dec->io->seek(dec->start_offset, dec->io->seek_data)) return L3O_E_IOERROR;
    l3o_init(&dec->mp3d);
    dec->buffer_samples = 0;
    return 0;
}

/* minimp3_ex.d:953-958.  The reference leaks the 128 KiB I/O buffer here (it only frees the index);
 * the oracle frees it, which is not observable. */
void l3o_ex_close(l3o_ex_t* dec)
{
    if (dec->frames) free(dec->frames);
    if (dec->io && dec->file_buffer) free((void*)dec->file_buffer);
    memset(dec, 0, sizeof *dec);
}

/* ------------------------------------------------------------------------------------------ */
/* AudioStream-shaped handle over memory (stream.d:150, 1706-1749, 2019-2131, 2243-2254)        */
struct l3o_stream {
    uint8_t* data; /* private copy, like openFromMemory (stream.d:2031-2041) */
    size_t size, cursor;
    l3o_io_t io;
    l3o_ex_t ex;
    int channels, samplerate;
    long long length_frames;
};

static size_t mem_read(void* buf, size_t size, void* user)
{
    l3o_stream_t* s = (l3o_stream_t*)user;
    size_t avail = s->size - s->cursor;
    size_t n = (int)size < 0 ? 0 : zmin(avail, size);
    memcpy(buf, s->data + s->cursor, n);
    s->cursor += n;
    return n;
}

static int mem_seek(uint64_t position, void* user)
{
    l3o_stream_t* s = (l3o_stream_t*)user;
    s->cursor = position > s->size ? s->size : (size_t)position;
    return 0; /* stream.d:2253: seek errors are not reported */
}

l3o_stream_t* l3o_stream_open_memory(const uint8_t* data, size_t size)
{
    l3o_stream_t* s = (l3o_stream_t*)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->data = (uint8_t*)malloc(size ? size : 1);
    memcpy(s->data, data, size);
    s->size = size;
    s->io.read = mem_read;
    s->io.read_data = s;
    s->io.seek = mem_seek;
    s->io.seek_data = s;
    uint8_t* scratch = (uint8_t*)malloc(BUF_SIZE * 2);
    int det = l3o_detect_cb(&s->io, scratch, BUF_SIZE * 2);
    free(scratch);
    if (det != 0 || l3o_ex_open_cb(&s->ex, &s->io, L3O_SEEK_TO_SAMPLE) != 0) {
        if (s->ex.file_buffer) l3o_ex_close(&s->ex);
        free(s->data);
        free(s);
        return NULL;
    }
    s->samplerate = s->ex.info.hz;
    s->channels = s->ex.info.channels;
    s->length_frames = s->channels ? (long long)(s->ex.samples / s->channels) : 0;
    return s;
}

void l3o_stream_close(l3o_stream_t* s)
{
    if (!s) return;
    l3o_ex_close(&s->ex);
    free(s->data);
    free(s);
}

int l3o_stream_channels(const l3o_stream_t* s) { return s->channels; }
int l3o_stream_samplerate(const l3o_stream_t* s) { return s->samplerate; }
long long l3o_stream_length_frames(const l3o_stream_t* s) { return s->length_frames; }

int l3o_stream_read_float(l3o_stream_t* s, float* out, int frames)
{
    int needed = frames * s->channels;
    int result = (int)l3o_ex_read(&s->ex, out, needed);
    if (result < 0) return 0;
    return result / s->channels;
}

int l3o_stream_seek(l3o_stream_t* s, int frame)
{
    if (frame < 0 || frame > s->length_frames) return 0;
    return l3o_ex_seek(&s->ex, (uint64_t)frame * s->channels) == 0;
}

int l3o_stream_tell(const l3o_stream_t* s) { return (int)s->ex.cur_sample / s->channels; }
int l3o_stream_last_error(const l3o_stream_t* s) { return s->ex.last_error; }

/* The transcode example's driver loop (examples/transcode/source/main.d:52-78): open, then read
 * `chunk_frames`-frame chunks until a read returns 0.  With out == NULL the chunk buffer is reused,
 * exactly like the example; otherwise the PCM is appended to out (cap_samples interleaved samples).
 * Returns the number of frames decoded, or -1 when the data is not detected as MP3. */
long long l3o_transcode_loop(const uint8_t* data, size_t size, int chunk_frames, float* out, size_t cap_samples,
                             int* channels_out, int* hz_out)
{
    l3o_stream_t* s = l3o_stream_open_memory(data, size);
    if (!s) return -1;
    int nch = s->channels;
    if (channels_out) *channels_out = nch;
    if (hz_out) *hz_out = s->samplerate;
    float* chunk = (float*)malloc(sizeof(float) * (size_t)chunk_frames * (size_t)nch);
    long long total = 0;
    size_t at = 0;
    for (;;) {
        int n = l3o_stream_read_float(s, chunk, chunk_frames);
        if (n <= 0) break;
        if (out) {
            size_t cnt = (size_t)n * (size_t)nch;
            if (at + cnt > cap_samples) cnt = cap_samples - at;
            memcpy(out + at, chunk, cnt * sizeof(float));
            at += cnt;
        }
        total += n;
    }
    free(chunk);
    l3o_stream_close(s);
    return total;
}

/* The same loop with the un-dithered float -> 16-bit conversion the reference's WAV writer applies per chunk
 * (wav.d:475-700 without its rand() dither; q = clamp(lrintf(x * 32768), -32768, 32767)): what a transcode to
 * 16-bit WAV costs on the CPU.  out16 may be NULL (the converted chunk is then discarded, but still computed). */
long long l3o_transcode_loop_s16(const uint8_t* data, size_t size, int chunk_frames, int16_t* out16, size_t cap_samples,
                                 int* channels_out, int* hz_out)
{
    l3o_stream_t* s = l3o_stream_open_memory(data, size);
    if (!s) return -1;
    int nch = s->channels;
    if (channels_out) *channels_out = nch;
    if (hz_out) *hz_out = s->samplerate;
    float* chunk = (float*)malloc(sizeof(float) * (size_t)chunk_frames * (size_t)nch);
    int16_t* chunk16 = (int16_t*)malloc(sizeof(int16_t) * (size_t)chunk_frames * (size_t)nch);
    long long total = 0;
    size_t at = 0;
    volatile int16_t sink = 0;
    for (;;) {
        int n = l3o_stream_read_float(s, chunk, chunk_frames);
        if (n <= 0) break;
        size_t cnt = (size_t)n * (size_t)nch;
        for (size_t i = 0; i < cnt; i++) {
            long q = lrintf(chunk[i] * 32768.0f);
            chunk16[i] = (int16_t)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
        }
        sink = chunk16[cnt - 1];
        if (out16) {
            if (at + cnt > cap_samples) cnt = cap_samples - at;
            memcpy(out16 + at, chunk16, cnt * sizeof(int16_t));
            at += cnt;
        }
        total += n;
    }
    (void)sink;
    free(chunk);
    free(chunk16);
    l3o_stream_close(s);
    return total;
}
