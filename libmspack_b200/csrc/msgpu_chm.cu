/* msgpu_chm.cu - CHM LZX section front end (include/msgpu_chm.h).  Host code only.
 * Restated from the format handling in the reference's chmd.c:1072-1267 and chm.h:75-92; no reference code is used. */
#include <stdint.h>
#include <string.h>
#include "../../include/msgpu_chm.h"

#define MSGPU_ERR_SIGNATURE_ 7

namespace {
inline uint32_t le32(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint64_t le64(const uint8_t *p) { return (uint64_t) le32(p) | ((uint64_t) le32(p + 4) << 32); }
}

extern "C" int msgpu_chm_units(const void *control_data, size_t control_bytes, const void *reset_table, size_t table_bytes,
                               uint64_t content_bytes, msgpu_unit *units, size_t max_units, msgpu_chm_info *info)
{
    const uint8_t *cd = reinterpret_cast<const uint8_t *>(control_data), *rt = reinterpret_cast<const uint8_t *>(reset_table);
    if (!cd || !rt) return MSGPU_ERR_ARGS;
    /* ControlData, chmd.c:1096-1149 (offsets chm.h:75-82) */
    if (control_bytes != 0x1C) return MSGPU_ERR_DATAFORMAT;
    if (le32(cd + 4) != 0x43585A4Cu) return MSGPU_ERR_SIGNATURE_;                       /* "LZXC" */
    uint64_t reset_interval, window_size;
    switch (le32(cd + 8)) {
    case 1: reset_interval = le32(cd + 0x0C); window_size = le32(cd + 0x10); break;
    case 2: reset_interval = (uint64_t) le32(cd + 0x0C) * 32768u; window_size = (uint64_t) le32(cd + 0x10) * 32768u; break;
    default: return MSGPU_ERR_DATAFORMAT;
    }
    /* the reference computes these in a 32-bit int; anything that does not survive that is not a valid section either */
    uint32_t window_bits = 0;
    for (uint32_t b = 15; b <= 21; b++) if (window_size == (1ull << b)) window_bits = b;
    if (!window_bits) return MSGPU_ERR_DATAFORMAT;
    if (reset_interval == 0 || reset_interval % 32768u || reset_interval > 0x7FFF8000ull) return MSGPU_ERR_DATAFORMAT;
    /* ResetTable, chmd.c:1193-1267 (offsets chm.h:84-92).  chmd.c falls back to "decode from the start of the section" when
     * the table is unusable; a batch of independent intervals needs the table, so that is an error here. */
    if (table_bytes < 0x28) return MSGPU_ERR_DATAFORMAT;
    if (le32(rt + 0x20) != 32768u) return MSGPU_ERR_DATAFORMAT;
    const uint64_t uncomp_len = le64(rt + 0x10);
    const uint32_t num_entries = le32(rt + 4), entry_size = le32(rt + 8), table_off = le32(rt + 0x0C);
    if (entry_size != 4 && entry_size != 8) return MSGPU_ERR_DATAFORMAT;
    const uint64_t frames_per_unit = reset_interval / 32768u;
    const uint64_t padded = (uncomp_len + reset_interval - 1) / reset_interval * reset_interval;      /* :1152-1158 */
    const uint64_t n = padded / reset_interval;
    if (info) { info->window_bits = window_bits; info->reset_interval = (uint32_t) reset_interval; info->uncomp_len = uncomp_len;
                info->padded_len = padded; info->num_units = n; }
    if (frames_per_unit > 0xFFFFu) return MSGPU_ERR_DATAFORMAT;                         /* msgpu_unit.reset_interval is 16 bits */
    if (!units) return 0;
    if (n > max_units) return MSGPU_ERR_ARGS;
    uint64_t prev_off = 0;
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t entry = k * frames_per_unit, pos = (uint64_t) table_off + entry * entry_size;
        if (entry >= num_entries || pos + entry_size > table_bytes) return MSGPU_ERR_DATAFORMAT;      /* :1235-1252 */
        const uint64_t off = entry_size == 4 ? le32(rt + pos) : le64(rt + pos);
        if (off > content_bytes || (k && off < prev_off)) return MSGPU_ERR_DATAFORMAT;
        msgpu_unit u; memset(&u, 0, sizeof(u));
        u.codec = MSGPU_CODEC_LZX; u.window_bits = (uint8_t) window_bits; u.reset_interval = (uint16_t) frames_per_unit;
        u.in_off = off; u.out_off = k * reset_interval; u.out_len = (uint32_t) reset_interval;
        /* interval k starts at frame k * frames_per_unit of the section's stream (E8 offsets and the frame-32768 rule count on) */
        if (k * frames_per_unit >= (1ull << 26)) return MSGPU_ERR_DATAFORMAT;
        if (k) u.flags = MSGPU_FLAG_LZX_STREAM_BASE | ((uint32_t) (k * frames_per_unit) << MSGPU_FLAG_REF_SHIFT);
        units[k] = u;
        if (k) {
            /* the zero-sized extra frame pass of lzxd.c:419 reads the NEXT interval's intel header (1 bit, or 1 + 32 bits when E8
             * translation is on: up to 8 bytes in 16-bit words) before the request is found complete: the unit may look that far */
            uint64_t len = off - prev_off, slack = content_bytes - off < 8 ? content_bytes - off : 8;
            if (len + slack >= 0x7FFFFFF0ull) return MSGPU_ERR_DATAFORMAT;
            units[k - 1].in_len = (uint32_t) (len + slack);
        }
        prev_off = off;
    }
    if (n) {
        if (content_bytes - prev_off >= 0x7FFFFFF0ull) return MSGPU_ERR_DATAFORMAT;
        units[n - 1].in_len = (uint32_t) (content_bytes - prev_off);
    }
    return 0;
}
