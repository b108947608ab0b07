/* msgpu_chm.h - CHM LZX section front end (SURVEY.md section 8, row f2 / BASELINE config 4): turns the two small system
 * files that describe a CHM's compressed content section into the unit table of include/msgpu.h, one unit per LZX reset
 * interval, so the whole section decodes as one batch.
 *
 *   reference (chmd.c)                                              here
 *   :1096-1149 ControlData: length 0x1C, "LZXC", version 1 / 2        msgpu_chm_units()
 *              (v2 counts reset interval and window in 32 KiB
 *              frames), window size -> window_bits 15..21, reset
 *              interval a non-zero multiple of 32 KiB
 *   :1193-1267 ResetTable: header 0x28 bytes, frame length must be    msgpu_chm_units()
 *              0x8000, 64-bit uncompressed length, entries of 4 or
 *              8 bytes at TableOffset, one per 32 KiB frame; the
 *              decoder only ever uses entry (k * interval / 32 KiB)
 *   :1152-1158 "the uncompressed length is dishonest": the stream     every unit's out_len is a whole interval; the
 *              is padded out to the next reset interval               caller keeps the first uncomp_len bytes
 *   :1175-1183 lzxd_init(window_bits, interval / 32 KiB, ...,         unit.window_bits / reset_interval / out_len
 *              remaining length) at the interval's compressed offset
 *
 * The reference seeks to one interval per extracted file and decodes forward from there; here all intervals are
 * independent units (fresh LZX state at every reset, lzxd.c:257-270 - what chmd.c itself relies on for random access).
 * Each unit's in_len reaches 8 bytes into the next interval where there is one: the reference's frame loop looks ahead
 * that far at a reset point (lzxd.c:419-453: the next interval's intel header, 1 or 33 bits fetched in 16-bit words), and a
 * unit cut exactly at its last byte would report MSPACK_ERR_READ.  What does NOT start afresh at a reset - the stream's frame
 * count and byte offset, which E8 call translation uses (lzxd.c:706-712) - travels in the unit: interval k > 0 carries
 * MSGPU_FLAG_LZX_STREAM_BASE with its first frame's index, so a section whose intervals have E8 translation on decodes to the
 * bytes the reference produces for the section as one stream.
 * Host-only; no GPU needed.
 */
#ifndef MSGPU_CHM_H
#define MSGPU_CHM_H

#include "msgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msgpu_chm_info {
    uint32_t window_bits;        /* 15..21                                              */
    uint32_t reset_interval;     /* bytes, a multiple of 32768                          */
    uint64_t uncomp_len;         /* the reset table's (honest) uncompressed length      */
    uint64_t padded_len;         /* rounded up to a whole reset interval = sum of the units' out_len */
    uint64_t num_units;
} msgpu_chm_info;

/* control_data / reset_table: the raw bytes of ::DataSpace/Storage/MSCompressed/ControlData and
 * .../Transform/{7FC28940-9D31-11D0-9B27-00A0C91E9C7C}/InstanceData/ResetTable; content_bytes: length of .../Content.
 * units (may be NULL to query info only) receives up to max_units descriptors: unit k has in_off = the interval's offset
 * inside Content, out_off = k * reset_interval.  Returns 0, MSGPU_ERR_SIGNATURE / MSGPU_ERR_DATAFORMAT as chmd.c would, or
 * MSGPU_ERR_ARGS when max_units is too small. */
int msgpu_chm_units(const void *control_data, size_t control_bytes, const void *reset_table, size_t table_bytes,
                    uint64_t content_bytes, msgpu_unit *units, size_t max_units, msgpu_chm_info *info);

#ifdef __cplusplus
}
#endif
#endif
