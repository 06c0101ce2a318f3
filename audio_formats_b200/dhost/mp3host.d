/**
  D host prepass for the GPU MP3 path: frame sync, side-info parsing and bit-reservoir main_data slicing,
  producing the descriptors the CUDA shim consumes (l3b200.d).

  Same algorithm as audio_formats_b200/csrc/l3_format.hpp + l3_host.cpp (the C++ mirror that IS compiled and
  tested in this repository); behaviour follows source/audioformats/minimp3.d -- the cited lines.
  NOT COMPILED HERE: no D toolchain in the build image.  Kept deliberately small and free of Phobos.
*/
module audioformats.mp3host;

import core.stdc.stdlib : malloc, realloc, free;
import core.stdc.string : memcpy, memset, memcmp;
import audioformats.l3b200;
import audioformats.minimp3 : hdr_valid, hdr_compare, hdr_frame_bytes, hdr_padding, hdr_frame_samples,
    hdr_sample_rate_hz, hdr_bitrate_kbps, HDR_IS_MONO, HDR_IS_CRC, HDR_TEST_MPEG1, HDR_GET_LAYER,
    HDR_GET_MY_SAMPLE_RATE, HDR_SIZE, bs_t, bs_init, get_bits, L3_gr_info_t, L3_read_side_info;

nothrow @nogc:

enum MAX_RESERVOIR = 511;            // minimp3.d:58

/// Growing output of the prepass for one decode run (the "decode program").
struct Mp3Program
{
nothrow @nogc:
    ubyte* blob;             /// all frame payloads of the run, concatenated
    size_t blobLen, blobCap;
    l3b_grch_desc_t* descs;  /// nch per granule
    size_t nDescs, descCap;
    uint granules;

    void clear() { blobLen = 0; nDescs = 0; granules = 0; }

    void release()
    {
        free(blob); free(descs);
        blob = null; descs = null; blobLen = blobCap = nDescs = descCap = 0; granules = 0;
    }

    bool appendBlob(const(ubyte)* p, size_t n)
    {
        if (blobLen + n + 32 > blobCap)
        {
            size_t cap = blobCap ? blobCap * 2 : 1 << 16;
            while (cap < blobLen + n + 32) cap *= 2;
            auto q = cast(ubyte*) realloc(blob, cap);
            if (q is null) return false;
            blob = q; blobCap = cap;
        }
        memcpy(blob + blobLen, p, n);
        blobLen += n;
        memset(blob + blobLen, 0, 16);   // the Huffman reader may look 16 bytes past the end
        return true;
    }

    bool appendDesc(l3b_grch_desc_t d)
    {
        if (nDescs + 1 > descCap)
        {
            size_t cap = descCap ? descCap * 2 : 4096;
            auto q = cast(l3b_grch_desc_t*) realloc(descs, cap * l3b_grch_desc_t.sizeof);
            if (q is null) return false;
            descs = q; descCap = cap;
        }
        descs[nDescs++] = d;
        return true;
    }
}

/// Pack one granule-channel (layout documented in include/l3b200.h).  Region boundaries are converted from
/// sfb counts to coefficient indices so the entropy kernel never needs the sfb tables (minimp3.d:780-853).
l3b_grch_desc_t packDesc(const(L3_gr_info_t)* g, uint bitStart, ubyte hdr3, bool secondGranule, bool resetBefore)
{
    int acc = 0, i = 0;
    for (; i <= g.region_count[0] && g.sfbtab[i]; i++) acc += g.sfbtab[i];
    int r1 = acc;
    for (int j = 0; j <= g.region_count[1] && g.sfbtab[i]; j++, i++) acc += g.sfbtab[i];
    int r2 = acc;
    l3b_grch_desc_t d;
    d.bit_start = bitStart;
    d.w1 = cast(uint) g.part_23_length | (cast(uint) g.big_values << 12) | (cast(uint) g.global_gain << 21)
         | (cast(uint) g.block_type << 29) | (cast(uint) g.mixed_block_flag << 31);
    d.w2 = cast(uint) g.scalefac_compress | (cast(uint) g.table_select[0] << 9) | (cast(uint) g.table_select[1] << 14)
         | (cast(uint) g.table_select[2] << 19) | (cast(uint) g.preflag << 24) | (cast(uint) g.scalefac_scale << 25)
         | (cast(uint) g.count1_table << 26) | (cast(uint) g.scfsi << 27) | ((secondGranule ? 1u : 0u) << 31);
    d.w3 = cast(uint)(r1 / 2) | (cast(uint)(r2 / 2) << 9) | (cast(uint) g.subblock_gain[0] << 18)
         | (cast(uint) g.subblock_gain[1] << 21) | (cast(uint) g.subblock_gain[2] << 24) | (cast(uint)(hdr3 >> 4) << 27)
         | ((resetBefore ? 1u : 0u) << 31);
    return d;
}

/// The part of mp3dec_t that steers control flow (minimp3.d:38-46); the sample state lives on the GPU.
struct FrameWalker
{
nothrow @nogc:
    ubyte[4] header;
    int freeFormatBytes;
    int reserv;               /// valid reservoir bytes == the last `reserv` bytes of prog.blob
    bool pendingReset = true; /// overlap / qmf / reservoir were zeroed since the last emitted granule

    void init() { header[0] = 0; }   // mp3dec_init, minimp3.d:1487

    /// mp3dec_decode_frame without the arithmetic (minimp3.d:1492-1581).  Returns samples per channel.
    /// *frameBytes receives info.frame_bytes; channels/hz/layer the frame's format.
    int step(const(ubyte)* mp3, int mp3Bytes, Mp3Program* prog, int* frameBytes, int* channels, int* hz, int* layer)
    {
        import audioformats.minimp3 : mp3d_find_frame;
        int i = 0, frameSize = 0;
        if (mp3Bytes > 4 && header[0] == 0xff && hdr_compare(header.ptr, mp3))
        {
            frameSize = hdr_frame_bytes(mp3, freeFormatBytes) + hdr_padding(mp3);
            if (frameSize != mp3Bytes && (frameSize + HDR_SIZE > mp3Bytes || !hdr_compare(mp3, mp3 + frameSize)))
                frameSize = 0;
        }
        if (!frameSize)
        {
            // memset(dec, 0, sizeof(mp3dec_t)): overlap, qmf history and reservoir all go to zero
            header[] = 0; freeFormatBytes = 0; reserv = 0; pendingReset = true;
            i = mp3d_find_frame(mp3, mp3Bytes, &freeFormatBytes, &frameSize);
            if (!frameSize || i + frameSize > mp3Bytes) { *frameBytes = i; return 0; }
        }
        const(ubyte)* hdr = mp3 + i;
        header[0 .. 4] = hdr[0 .. 4];
        *frameBytes = i + frameSize;
        *channels = HDR_IS_MONO(hdr) ? 1 : 2;
        *hz = hdr_sample_rate_hz(hdr);
        *layer = 4 - HDR_GET_LAYER(hdr);

        bs_t bs;
        bs_init(&bs, hdr + HDR_SIZE, frameSize - HDR_SIZE);
        if (HDR_IS_CRC(hdr)) get_bits(&bs, 16);            // skipped, never verified (minimp3.d:1533-1536)
        if (*layer != 3) { init(); return 0; }             // Layer I/II: outside the GPU path
        L3_gr_info_t[4] gr;
        int mdb = L3_read_side_info(&bs, gr.ptr, hdr);
        if (mdb < 0 || bs.pos > bs.limit) { init(); return 0; }

        // L3_restore_reservoir (minimp3.d:1186-1194), as counts
        const int payload = (bs.limit - bs.pos) / 8;
        const(ubyte)* payloadPtr = hdr + HDR_SIZE + bs.pos / 8;
        const bool success = reserv >= mdb;
        const int nch = *channels;
        const int ngr = HDR_TEST_MPEG1(hdr) ? 2 : 1;
        int consumedBits = 0;
        if (success)
        {
            ulong startBit = (cast(ulong) prog.blobLen - cast(ulong) mdb) * 8;
            foreach (g; 0 .. ngr)
            {
                foreach (ch; 0 .. nch)
                {
                    const(L3_gr_info_t)* q = &gr[g * nch + ch];
                    ulong b = startBit + cast(ulong) consumedBits;
                    if (b > uint.max) b = uint.max;
                    if (!prog.appendDesc(packDesc(q, cast(uint) b, hdr[3], g == 1, pendingReset))) return 0;
                    consumedBits += q.part_23_length;
                }
                pendingReset = false;
                prog.granules++;
            }
            // L3_save_reservoir (minimp3.d:1170-1184): keep the unread tail, newest 511 bytes at most
            int remains = (mdb + payload) - (consumedBits + 7) / 8;
            reserv = remains > MAX_RESERVOIR ? MAX_RESERVOIR : (remains < 0 ? 0 : remains);
        }
        else
        {
            int r = reserv + payload;
            reserv = r > MAX_RESERVOIR ? MAX_RESERVOIR : r;
        }
        prog.appendBlob(payloadPtr, payload);
        return success ? cast(int) hdr_frame_samples(header.ptr) : 0;
    }
}
