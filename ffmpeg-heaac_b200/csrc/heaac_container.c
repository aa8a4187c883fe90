/*
 * heaac_container.c -- the two thin container layers a file-driven front end needs, host C:
 *   ADTS framing     write: libavformat/adtsenc.c:84-147 (adts_write_frame_header)
 *                    read:  libavcodec/aac_parser.c:29-67 / aacadtsdec.c (ff_aac_parse_header)
 *   MP4 `esds`       libavformat/isom.c (ff_mp4_read_descr, ff_mp4_read_dec_config_descr): ES_Descriptor (tag 3) ->
 *                    DecoderConfigDescriptor (tag 4) -> DecoderSpecificInfo (tag 5) = the AudioSpecificConfig that
 *                    heaac_fe_open / heaac_dec_open take as extradata (explicit SBR / PS signalling included).
 */
#include <string.h>
#include "../../include/heaac_b200.h"

#define ADTS_HEADER_SIZE 7

/* 7-byte header without CRC for one raw_data_block of `payload_bytes`; profile = object type - 1 (1 = LC) */
int heaac_adts_write_header(uint8_t out[7], int sr_index, int chan_config, int payload_bytes)
{
    const int full = payload_bytes + ADTS_HEADER_SIZE;
    if (!out || sr_index < 0 || sr_index > 12 || chan_config < 1 || chan_config > 7 || payload_bytes < 0 || full >= (1 << 13))
        return HEAAC_E_INVAL;
    out[0] = 0xff;                                             /* syncword 0xfff, ID 0 (MPEG-4), layer 0, protection_absent 1 */
    out[1] = 0xf1;
    out[2] = (uint8_t)((1 << 6) | (sr_index << 2) | (chan_config >> 2));       /* profile LC, sampling index, private 0 */
    out[3] = (uint8_t)(((chan_config & 3) << 6) | (full >> 11));               /* original/copy, home, copyright bits 0 */
    out[4] = (uint8_t)(full >> 3);
    out[5] = (uint8_t)(((full & 7) << 5) | 0x1f);                              /* buffer fullness 0x7ff (VBR) */
    out[6] = 0xfc;                                                             /* ... and one raw_data_block */
    return ADTS_HEADER_SIZE;
}

/* Returns the header size (7 or 9 with CRC) and fills the outputs, or a negative error.  frame_bytes includes the
 * header.  size >= 7 bytes of the candidate frame must be available. */
int heaac_adts_parse_header(const uint8_t *b, int size, int *sr_index, int *chan_config, int *frame_bytes, int *n_blocks)
{
    if (!b || size < ADTS_HEADER_SIZE) return HEAAC_E_INVAL;
    if (b[0] != 0xff || (b[1] & 0xf0) != 0xf0) return HEAAC_E_INVAL;           /* syncword */
    if (b[1] & 0x06) return HEAAC_E_INVAL;                                     /* layer must be 0 */
    const int crc_absent = b[1] & 1;
    const int sr = (b[2] >> 2) & 15;
    const int ch = ((b[2] & 1) << 2) | (b[3] >> 6);
    const int full = ((b[3] & 3) << 11) | (b[4] << 3) | (b[5] >> 5);
    if (sr > 12 || full < ADTS_HEADER_SIZE) return HEAAC_E_INVAL;
    if (sr_index) *sr_index = sr;
    if (chan_config) *chan_config = ch;
    if (frame_bytes) *frame_bytes = full;
    if (n_blocks) *n_blocks = (b[6] & 3) + 1;
    return crc_absent ? 7 : 9;
}

/* Split a buffer of back-to-back ADTS frames: fills offset[] / size[] (whole frames, headers included) for up to
 * max_frames frames and returns how many were found; stops at the first byte that is not a frame start. */
int heaac_adts_split(const uint8_t *buf, int size, int *offset, int *frame_size, int max_frames)
{
    if (!buf || size < 0 || !offset || !frame_size) return HEAAC_E_INVAL;
    int n = 0, pos = 0;
    while (n < max_frames && pos + ADTS_HEADER_SIZE <= size) {
        int full;
        if (heaac_adts_parse_header(buf + pos, size - pos, NULL, NULL, &full, NULL) < 0 || pos + full > size) break;
        offset[n] = pos; frame_size[n] = full;
        pos += full; n++;
    }
    return n;
}

/* MPEG-4 descriptor length: up to four 7-bit groups, MSB = "more" (isom.c: ff_mp4_read_descr_len) */
static int descr_len(const uint8_t *b, int size, int *pos)
{
    int len = 0;
    for (int i = 0; i < 4; i++) {
        if (*pos >= size) return -1;
        const int c = b[(*pos)++];
        len = (len << 7) | (c & 0x7f);
        if (!(c & 0x80)) break;
    }
    return len;
}

/* `esds` box payload (after the 4 version/flags bytes are skipped by the caller, or with them: both accepted)
 * -> pointer to and size of the AudioSpecificConfig inside it.  Returns 0 or a negative error. */
int heaac_esds_find_asc(const uint8_t *b, int size, const uint8_t **asc, int *asc_size)
{
    if (!b || size < 2 || !asc || !asc_size) return HEAAC_E_INVAL;
    int pos = 0;
    if (b[0] == 0 && size > 4 && b[4] == 0x03) pos = 4;                        /* version + flags present */
    if (b[pos++] != 0x03) return HEAAC_E_INVAL;                                /* ES_DescrTag */
    if (descr_len(b, size, &pos) < 0 || pos + 3 > size) return HEAAC_E_INVAL;
    const int flags = b[pos + 2];
    pos += 3;                                                                  /* ES_ID, flags */
    if (flags & 0x80) pos += 2;                                                /* dependsOn_ES_ID */
    if (flags & 0x40) { if (pos >= size) return HEAAC_E_INVAL; pos += 1 + b[pos]; }   /* URL */
    if (flags & 0x20) pos += 2;                                                /* OCR_ES_Id */
    if (pos >= size || b[pos++] != 0x04) return HEAAC_E_INVAL;                 /* DecoderConfigDescrTag */
    if (descr_len(b, size, &pos) < 0 || pos + 13 > size) return HEAAC_E_INVAL;
    const int object_type = b[pos];
    pos += 13;                                                                 /* objectType, streamType, buffer size, bitrates */
    if (object_type != 0x40 && object_type != 0x66 && object_type != 0x67 && object_type != 0x68)
        return HEAAC_E_UNSUPPORTED;                                            /* not MPEG-4 / MPEG-2 AAC */
    if (pos >= size || b[pos++] != 0x05) return HEAAC_E_INVAL;                 /* DecSpecificInfoTag */
    const int len = descr_len(b, size, &pos);
    if (len <= 0 || pos + len > size) return HEAAC_E_INVAL;
    *asc = b + pos; *asc_size = len;
    return 0;
}
