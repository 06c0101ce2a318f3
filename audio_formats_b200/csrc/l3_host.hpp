// l3_host.hpp -- host prepass: turns an MP3 byte stream into the "decode program" the GPU runs.
//
// This is the C++ mirror of the D host named by the north star (no D compiler in the build image).
// It reproduces, without doing any arithmetic on samples, the control flow of
//   mp3dec_decode_frame   minimp3.d:1492-1581   (sync, resync+state reset, side info, reservoir)
//   mp3dec_iterate_cb     minimp3_ex.d:490-564  (whole-file walk through a 128 KiB window)
//   mp3dec_load_index     minimp3_ex.d:566-621  (VBR tag, index, length)
//   mp3dec_ex_read/seek   minimp3_ex.d:662-888  (delay skip, padding trim, pre-roll)
// so that every granule the reference would decode gets a descriptor with an absolute bit offset
// into one linear main-data blob, plus a "state was zeroed before me" flag.
#pragma once
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#include "l3_format.hpp"

namespace l3b {

constexpr size_t kIoSize = 128 * 1024;  // MINIMP3_IO_SIZE   minimp3_ex.d:26
constexpr size_t kBufSize = 16 * 1024;  // MINIMP3_BUF_SIZE  minimp3_ex.d:27
constexpr int kPredecodeFrames = 2;     // minimp3_ex.d:24
constexpr int kId3DetectSize = 10;      // minimp3_ex.d:113

struct FrameInfo {  // mp3dec_frame_info_t, minimp3.d:28-36
    int frame_bytes = 0, frame_offset = 0, channels = 0, hz = 0, layer = 0, bitrate_kbps = 0;
};

// The growing output of the prepass for one decode run.
struct Program {
    std::vector<uint8_t> blob;          // all frame payloads of the run, concatenated
    std::vector<l3b_grch_desc_t> descs; // nch per granule
    uint32_t granules = 0;
    void clear() { blob.clear(); descs.clear(); granules = 0; }
};

// The part of mp3dec_t that steers control flow (minimp3.d:38-46); sample state lives on the GPU.
class FrameWalker {
  public:
    uint8_t header[4] = {0, 0, 0, 0};
    int free_format_bytes = 0;
    int reserv = 0;             // valid reservoir bytes == the last `reserv` bytes of prog->blob
    bool pending_reset = true;  // overlap/qmf/reservoir were zeroed since the last emitted granule

    void init() { header[0] = 0; }  // mp3dec_init, minimp3.d:1487

    // mp3dec_decode_frame without the arithmetic.  Returns samples per channel (0 = nothing decoded).
    // When prog != nullptr and the frame decodes, its payload is appended to prog->blob and one
    // descriptor per granule-channel is appended to prog->descs.  When prog == nullptr only the
    // control state advances (used while indexing, minimp3_ex.d:613-619).
    int step(const uint8_t* mp3, int mp3_bytes, FrameInfo* info, Program* prog);
};

struct IndexEntry { uint64_t sample, offset; };

// What mp3dec_ex_open_cb leaves in mp3dec_ex_t (minimp3_ex.d:73-87, 929-951).
struct OpenInfo {
    FrameInfo info;
    std::vector<IndexEntry> index;
    uint64_t samples = 0, detected_samples = 0, start_offset = 0, end_offset = 0;
    int start_delay = 0, to_skip = 0, vbr_tag_found = 0, free_format_bytes = 0;
    bool index_started = false;  // "dec.index.frames != null"
};

int skip_id3v2(const uint8_t* buf, size_t size, size_t* id3v2size);     // minimp3_ex.d:115
void skip_id3v1(const uint8_t* buf, size_t* size);                      // minimp3_ex.d:93
int detect_mp3(const uint8_t* data, size_t size);                       // minimp3_ex.d:197 via memory I/O
int open_index(const uint8_t* data, size_t size, OpenInfo* out, uint64_t from_offset = 0);  // minimp3_ex.d:490-621

// Sequential reader state == the rest of mp3dec_ex_t, driving FrameWalker through the same
// 128 KiB sliding window the reference reads through (window arithmetic only, no copying).
class Reader {
  public:
    Reader(const uint8_t* data, size_t size) : data_(data), size_(size) {}
    void restart(uint64_t offset);  // after open / seek: mp3dec_ex_seek's do_exit block (minimp3_ex.d:772-784)
    void begin_call() { eof_ = false; }  // `int eof = 0;` is a local of mp3dec_ex_read (minimp3_ex.d:793)

    struct Frame {
        int samples = 0;         // interleaved samples produced (0: skipped / undecodable)
        int hdr_samples = 0;     // hdr_frame_samples*channels of the bytes at the frame position (to_skip accounting)
        uint32_t first_granule = 0;
        bool format_change = false;
        bool end_of_input = false;
    };
    // One iteration of the mp3dec_ex_read loop body up to and including mp3dec_decode_frame
    // (minimp3_ex.d:819-858).  Appends to `prog`.
    Frame next(const OpenInfo& oi, Program* prog);

    FrameWalker walker;
    uint64_t offset = 0;  // dec.offset

  private:
    const uint8_t* data_;
    size_t size_;
    size_t cursor_ = 0;                               // I/O cursor of the memory "file"
    size_t win_start_ = 0, filled_ = 0, consumed_ = 0; // dec.file.buffer window [win_start_, win_start_+filled_)
    bool eof_ = false;
};

// Whole-stream scan: open + read-to-end, for the batch entry point.
struct ScanResult {
    OpenInfo open;
    Program prog;
    int channels = 0, hz = 0, sr_idx = 0, mpeg1 = 0, layer = 0;
    uint64_t length_frames = 0;
    uint64_t pcm_skip = 0, pcm_count = 0;  // delivered window inside the decoded signal (interleaved samples)
    int last_error = 0;
};
int scan_stream(const uint8_t* data, size_t size, ScanResult* out);

}  // namespace l3b

struct l3b_scan { l3b::ScanResult r; };
