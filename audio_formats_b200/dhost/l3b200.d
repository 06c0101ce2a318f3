/**
  extern(C) bindings of include/l3b200.h (the CUDA shim + host layer of libl3b200.so).

  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no D compiler (dmd/ldc2/gdc/dub absent).
  tests/test_dhost_bindings.py checks that every function declared in the C header is declared here
  with the same name, so the two cannot drift silently.
*/
module audioformats.l3b200;

nothrow @nogc extern(C):

enum L3B_OK = 0;
enum L3B_E_PARAM = -1;       // MP3D_E_PARAM   minimp3_ex.d:30
enum L3B_E_MEMORY = -2;      // MP3D_E_MEMORY  minimp3_ex.d:31
enum L3B_E_IOERROR = -3;     // MP3D_E_IOERROR minimp3_ex.d:32
enum L3B_E_USER = -4;        // MP3D_E_USER    minimp3_ex.d:33
enum L3B_E_DECODE = -5;      // MP3D_E_DECODE  minimp3_ex.d:34
enum L3B_E_NOGPU = -16;
enum L3B_E_UNSUPPORTED = -17;

struct l3b_grch_desc_t
{
    uint bit_start;
    uint w1;
    uint w2;
    uint w3;
}

struct l3b_stream_desc_t
{
    ulong maindata_off;
    uint maindata_bytes;
    uint n_granules;
    ulong first_grch;
    ulong pcm_off;
    ulong pcm_skip;
    ulong pcm_count;
    ubyte nch;
    ubyte sr_idx;
    ubyte mpeg1;
    ubyte layer;      // 0 / 3: Layer III; 1 / 2: Layer I / II
    uint reserved2;
}

struct l3b_taps_t
{
    short* is_;
    ubyte* iscf;
    ubyte* ist_pos;
    float* xr;
    float* st;
    float* im;
    float* dct;
}

enum L3B_OUT_S16 = 1;          // 16-bit delivery: clamp(lrintf(x * 32768)), the un-dithered conversion of wav.d:475-700
enum L3B_MATH_FUSED = 2;    // tolerance-mode arithmetic (FMA-contracted), not bit-identical

struct l3b_batch_t
{
    const(ubyte)* maindata;
    ulong maindata_bytes;
    const(l3b_grch_desc_t)* grch;
    ulong n_grch;
    const(l3b_stream_desc_t)* streams;
    uint n_streams;
    void* pcm;               // float* or short* (L3B_OUT_S16)
    ulong pcm_floats;
    int* status;
    const(l3b_taps_t)* taps;
    uint flags;
    uint reserved;
}

struct l3b_pipeline_opts_t
{
    int lanes;
    int wave_streams;
    int scan_threads;
    uint flags;
}

struct l3b_stream_result_t
{
    ulong pcm_off;
    ulong frames;
    int channels;
    int samplerate;
    int status;
    int device;
}

enum L3B_PIPELINE_PHASES = 6;
alias l3b_read_cb = size_t function(void* buf, size_t size, void* user);   // mp3dec_io_t.read, minimp3_ex.d:61-71
alias l3b_seek_cb = int function(ulong position, void* user);              // mp3dec_io_t.seek

struct l3b_ctx;
struct l3b_resident;
struct l3b_scan;
struct l3b_stream;
struct l3b_pipeline;
alias l3b_pipeline_t = l3b_pipeline;
alias l3b_ctx_t = l3b_ctx;
alias l3b_resident_t = l3b_resident;
alias l3b_scan_t = l3b_scan;
alias l3b_stream_t = l3b_stream;

// layer 1: shim
int l3b_device_count();
int l3b_ctx_create(int device_id, l3b_ctx_t** outCtx);
void l3b_ctx_destroy(l3b_ctx_t* ctx);
const(char)* l3b_last_error(const(l3b_ctx_t)* ctx);
void* l3b_host_alloc(size_t bytes);
void* l3b_host_alloc_near(int device_id, size_t bytes);
void l3b_host_free(void* p);
int l3b_decode_batch(l3b_ctx_t* ctx, const(l3b_batch_t)* batch);
int l3b_batch_upload(l3b_ctx_t* ctx, const(l3b_batch_t)* batch, l3b_resident_t** outResident);
int l3b_batch_upload_reuse(l3b_ctx_t* ctx, const(l3b_batch_t)* batch, l3b_resident_t** inoutResident);
int l3b_batch_reupload(l3b_ctx_t* ctx, l3b_resident_t* r, const(l3b_batch_t)* batch);
int l3b_batch_run(l3b_ctx_t* ctx, l3b_resident_t* r);
int l3b_batch_sync(l3b_ctx_t* ctx);
int l3b_batch_download(l3b_ctx_t* ctx, l3b_resident_t* r, void* pcm_host, ulong first, ulong n);
int l3b_batch_download_taps(l3b_ctx_t* ctx, l3b_resident_t* r, const(l3b_taps_t)* taps);
void* l3b_batch_device_pcm(l3b_resident_t* r);
void l3b_batch_free(l3b_ctx_t* ctx, l3b_resident_t* r);
int l3b_batch_timing(l3b_ctx_t* ctx, int last_runs, float* ms3, int* launches);
void* l3b_ctx_cuda_stream(l3b_ctx_t* ctx);

// layer 2: host prepass + AudioStream surface as implemented in C++ (the D host below can use either its own
// prepass, mp3host.d, or these)
int l3b_scan_memory(const(ubyte)* data, size_t size, l3b_scan_t** outScan);
void l3b_scan_free(l3b_scan_t* s);
int l3b_scan_channels(const(l3b_scan_t)* s);
int l3b_scan_samplerate(const(l3b_scan_t)* s);
int l3b_scan_error(const(l3b_scan_t)* s);
ulong l3b_scan_length_frames(const(l3b_scan_t)* s);
ulong l3b_scan_delivered_samples(const(l3b_scan_t)* s);
uint l3b_scan_granules(const(l3b_scan_t)* s);
ulong l3b_scan_maindata_bytes(const(l3b_scan_t)* s);
const(ubyte)* l3b_scan_maindata(const(l3b_scan_t)* s);
const(l3b_grch_desc_t)* l3b_scan_descs(const(l3b_scan_t)* s);
void l3b_scan_fill_stream_desc(const(l3b_scan_t)* s, l3b_stream_desc_t* outDesc);
int l3b_scans_assemble(l3b_scan_t** scans, uint n, ubyte* blob, ulong blobCap, l3b_grch_desc_t* descs, ulong descCap,
                       l3b_stream_desc_t* streams, l3b_batch_t* batch);
int l3b_decode_scans(l3b_ctx_t* ctx, l3b_scan_t** scans, uint n, float** pcm, int* status);

int l3b_pipeline_create(const(int)* device_ids, int n_devices, const(l3b_pipeline_opts_t)* opts, l3b_pipeline_t** outPipeline);
void l3b_pipeline_destroy(l3b_pipeline_t* p);
int l3b_pipeline_decode(l3b_pipeline_t* p, const(ubyte*)* data, const(size_t)* size, uint n, void* outPcm, ulong outCapacity,
                        l3b_stream_result_t* results, ulong* outUsed);
int l3b_pipeline_profile(l3b_pipeline_t* p, double* seconds6);
const(char)* l3b_pipeline_last_error(const(l3b_pipeline_t)* p);

struct l3b_raw;
alias l3b_raw_t = l3b_raw;
int l3b_raw_open(l3b_ctx_t* ctx, const(ubyte*)* data, const(size_t)* size, uint n, uint flags, l3b_raw_t** outRaw);
uint l3b_raw_device_streams(const(l3b_raw_t)* r);
float l3b_raw_prepass_ms(const(l3b_raw_t)* r);
int l3b_raw_channels(const(l3b_raw_t)* r, uint i);
int l3b_raw_samplerate(const(l3b_raw_t)* r, uint i);
ulong l3b_raw_samples(const(l3b_raw_t)* r, uint i);
int l3b_raw_status(const(l3b_raw_t)* r, uint i);
int l3b_raw_decode(l3b_raw_t* r, void** pcm);
void l3b_raw_free(l3b_raw_t* r);

int l3b_stream_open_callbacks(l3b_ctx_t* ctx, l3b_read_cb read, l3b_seek_cb seek, void* user, l3b_stream_t** outStream);
int l3b_stream_open_memory(l3b_ctx_t* ctx, const(ubyte)* data, size_t size, l3b_stream_t** outStream);
int l3b_stream_open_file(l3b_ctx_t* ctx, const(char)* path, l3b_stream_t** outStream);
void l3b_stream_close(l3b_stream_t* s);
int l3b_stream_num_channels(const(l3b_stream_t)* s);
long l3b_stream_length_frames(const(l3b_stream_t)* s);
float l3b_stream_samplerate(const(l3b_stream_t)* s);
int l3b_stream_read_float(l3b_stream_t* s, float* outData, int frames);
int l3b_stream_read_double(l3b_stream_t* s, double* outData, int frames);
int l3b_stream_seek(l3b_stream_t* s, int frame);
int l3b_stream_tell(const(l3b_stream_t)* s);
int l3b_stream_is_error(const(l3b_stream_t)* s);
const(char)* l3b_stream_error_message(const(l3b_stream_t)* s);
