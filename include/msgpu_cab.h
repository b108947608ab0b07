/* msgpu_cab.h - cabinet front end of the batch decompressor.
 *
 * Builds the unit table straight from a .cab image and does the CFDATA framing the reference does on
 * the host on the device instead:
 *
 *   reference                                                  here
 *   cabd.c:308-480   cabd_read_headers (CFHEADER, CFFOLDER,     msgpu_cab_scan()        host, reads headers only
 *                    CFFILE tables, reserved areas)
 *   cabd.c:1362-1418 cabd_sys_read_block: 8-byte CFDATA header,  msgpu_cab_scan()        (sizes / limits, host)
 *                    block_resv skip, CAB_INPUTMAX / CAB_BLOCKMAX
 *   cabd.c:1412-1419 per-block checksum, cabd_checksum :1456-1479  k_cab_gather            device, one warp per block
 *   cabd.c:1294-1344 cabd_sys_read: payloads concatenated per     k_cab_gather            device (packed codec input,
 *                    folder, 0xFF after every Quantum block,                              Quantum trailer bytes)
 *                    lzxd_set_output_length(sum of block sizes)   msgpu_cab_scan()        unit.out_len
 *   cabd.c:1239-1249 codec per folder (comp_type & 0x000F,        msgpu_cab_scan()        unit.codec / window_bits
 *                    window bits = (comp_type >> 8) & 0x1F)
 *
 * One folder = one unit of include/msgpu.h; all folders of the cabinet decode as one batch.  Folders with
 * compression type 0 ("none", cabd.c noned_*) are copied block by block by the gather kernel.
 *
 * Cabinet SETS (cabd.c:760-1002 append / cabd_merge, :1421-1452): msgpu_cab_scan_set() takes the set's cabinets in order and
 * merges a folder that is continued in the next cabinet with its continuation - CFFILE folder indices 0xFFFD-0xFFFF say which -
 * including a CFDATA block that is split over two cabinets (its two pieces are joined in front of the codec, each piece is
 * checksummed by itself).  A folder whose other half is not among the images is reported as MSGPU_ERR_DATAFORMAT, the rest decodes.
 * MSGPU_CAB_SALVAGE is MSCABD_PARAM_SALVAGE as far as the data path goes: no checksum test, blocks of up to 65535 compressed
 * bytes and any claimed uncompressed size, running out of blocks is not an error of its own (cabd.c:1289-1292, :1312, :1393-1402);
 * the header scan's own salvage behaviour (file tables read from the header's offset, cabd.c:466-489) is not offered.
 *
 * Per-folder status == what the reference's mscab_decompressor::extract() (cabd.c:1004-1140) returns for a file that needs
 * the whole folder: MSGPU_ERR_OK, the codec's MSGPU_ERR_DECRUNCH, MSGPU_ERR_CHECKSUM for a block whose stored checksum is
 * wrong (the codec runs up to that block first, so an earlier codec error wins, as in the reference), MSGPU_ERR_DATAFORMAT
 * for oversized blocks or a folder whose blocks end before the codec is done (cabd.c:1311-1318, :1386-1398).
 */
#ifndef MSGPU_CAB_H
#define MSGPU_CAB_H

#include "msgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MSGPU_ERR_CHECKSUM   9    /* MSPACK_ERR_CHECKSUM, mspack.h:503 */
#define MSGPU_ERR_SIGNATURE  7    /* MSPACK_ERR_SIGNATURE: not a cabinet */

#define MSGPU_CAB_BLOCKMAX   (32768)          /* CAB_BLOCKMAX, cab.h:70 */
#define MSGPU_CAB_INPUTMAX   (32768 + 6144)   /* CAB_INPUTMAX, cab.h:78 */

typedef struct msgpu_cab_folder {
    uint16_t comp_type;        /* raw typeCompress of the CFFOLDER                               */
    uint8_t  codec;            /* MSGPU_CODEC_*, 0 = stored                                      */
    uint8_t  window_bits;
    uint32_t num_blocks;       /* CFDATA blocks found inside the image (<= the header's count)   */
    uint32_t first_block;      /* index into msgpu_cab_blocks()                                  */
    int32_t  scan_status;      /* 0, or the MSGPU_ERR_* the scan already knows (see above)       */
    uint32_t bad_block;        /* when scan_status != 0: folder-relative index of the offending block */
    uint64_t out_off;          /* where the folder's bytes go in the output buffer (multiple of 16) */
    uint64_t out_len;          /* sum of the blocks' uncompressed sizes                          */
    uint64_t in_off, in_len;   /* the folder's packed codec input (payloads + Quantum trailers)  */
} msgpu_cab_folder;

typedef struct msgpu_cab_block {
    uint64_t payload_off;      /* offset of the block's payload in the image (a set: in the images laid end to end, each
                                * starting at a multiple of 16)                                    */
    uint64_t dst_off;          /* where it goes: packed input (codec folders) or output (stored) */
    uint32_t checksum;         /* stored value, 0 = none                                          */
    uint16_t comp_len, uncomp_len;
    uint32_t folder;
    uint32_t flags;            /* bit 0: Quantum (append 0xFF), bit 1: stored (dst is the output buffer), bit 2: first piece of a block
                                * that continues in the next cabinet (uncomp_len 0) */
} msgpu_cab_block;

typedef struct msgpu_cab_file {
    uint32_t folder;           /* index into msgpu_cab_folders() (a file continued from / in another cabinet: the merged folder;
                                * a set lists such a file once per cabinet that names it, as the cabinets do)  */
    uint32_t offset, length;   /* uoffFolderStart / cbFile, cab.h:33-34                          */
    uint32_t name_off;         /* offset of the NUL-terminated name in the image                 */
} msgpu_cab_file;

typedef struct msgpu_cab_plan msgpu_cab_plan;

/* Parse the headers of a single cabinet image held in HOST memory.  No GPU needed.  Returns NULL and sets *err
 * (MSGPU_ERR_SIGNATURE, MSGPU_ERR_DATAFORMAT, MSGPU_ERR_READ for a truncated header area, MSGPU_ERR_NOMEMORY) on failure. */
msgpu_cab_plan *msgpu_cab_scan(const void *image, size_t image_bytes, int *err);
/* The same for a SET of cabinets given in order (see above); flags: MSGPU_CAB_SALVAGE. */
#define MSGPU_CAB_SALVAGE 1u
msgpu_cab_plan *msgpu_cab_scan_set(const void *const *images, const size_t *image_bytes, size_t ncabs, uint32_t flags, int *err);
void msgpu_cab_free(msgpu_cab_plan *plan);

size_t msgpu_cab_num_folders(const msgpu_cab_plan *plan);
size_t msgpu_cab_num_blocks(const msgpu_cab_plan *plan);
size_t msgpu_cab_num_files(const msgpu_cab_plan *plan);
const msgpu_cab_folder *msgpu_cab_folders(const msgpu_cab_plan *plan);
const msgpu_cab_block *msgpu_cab_blocks(const msgpu_cab_plan *plan);
const msgpu_cab_file *msgpu_cab_files(const msgpu_cab_plan *plan);
size_t msgpu_cab_out_bytes(const msgpu_cab_plan *plan);      /* size of the output buffer msgpu_cab_decode_host() fills */
size_t msgpu_cab_packed_bytes(const msgpu_cab_plan *plan);   /* device scratch for the packed codec input              */

/* Decode every folder of the cabinet: image (host) -> h_out (host, msgpu_cab_out_bytes() bytes, folder f at
 * folders[f].out_off) and folder_status[num_folders] (may be NULL).  Synchronous.  Returns 0 if the batch ran. */
int msgpu_cab_decode_host(msgpu_ctx *ctx, const msgpu_cab_plan *plan, const void *image, size_t image_bytes,
                          void *h_out, size_t out_bytes, int32_t *folder_status);
int msgpu_cab_decode_host_set(msgpu_ctx *ctx, const msgpu_cab_plan *plan, const void *const *images, const size_t *image_bytes, size_t ncabs,
                              void *h_out, size_t out_bytes, int32_t *folder_status);

#ifdef __cplusplus
}
#endif
#endif
