/* msgpu.h - C-ABI of the B200 batch decompressor for the three CAB-folder codecs
 * (MSZIP / Quantum / LZX).
 *
 * This is the ONE interface the project adds on top of the reference's own entry
 * points ("New batch surface").  The reference library decodes one
 * stream at a time through
 *     mszipd_init / mszipd_decompress / mszipd_free   (libmspack/mspack/mszip.h:85-120)
 *     qtmd_init   / qtmd_decompress   / qtmd_free     (libmspack/mspack/qtm.h:92-122)
 *     lzxd_init   / lzxd_decompress   / lzxd_free     (libmspack/mspack/lzx.h:146-214)
 * A GPU needs thousands of independent streams ("units") in flight, so the batch
 * call below takes an array of unit descriptors.  One unit == one fresh codec state:
 * a CAB folder (cabd.c:1142-1177 creates a fresh decoder per folder) or a CHM LZX
 * reset interval (chmd.c:1146-1186 starts a fresh lzxd_stream at an interval).
 * The ABI-compatible streaming entry points in mspack_dropin.h are implemented on top of
 * this call with n == 1.
 *
 * Plain C, plain pointers and sizes: no torch / C++ types cross this boundary.
 */
#ifndef MSGPU_H
#define MSGPU_H 1

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* codec ids == the CAB folder compression-type ids (cab.h:50-52 cffoldCOMPTYPE_*) */
#define MSGPU_CODEC_MSZIP   1
#define MSGPU_CODEC_QUANTUM 2
#define MSGPU_CODEC_LZX     3

/* per-unit status == the reference's error codes (mspack.h:485-507) */
#define MSGPU_ERR_OK         0
#define MSGPU_ERR_ARGS       1
#define MSGPU_ERR_READ       3   /* ran past the end of the unit's input (readbits.h:196-208) */
#define MSGPU_ERR_WRITE      4
#define MSGPU_ERR_NOMEMORY   6
#define MSGPU_ERR_DATAFORMAT 8
#define MSGPU_ERR_DECRUNCH   11

/* unit flags */
#define MSGPU_FLAG_MSZIP_REPAIR 0x1u  /* mszipd_init(repair_mode=1), mszipd.c:420-433: a block that does not inflate is zero-filled to 32 KiB
                                       * and decoding goes on behind it.  WHERE the reference goes on depends on the size of its input
                                       * buffer (its stale bit state): flags >> MSGPU_FLAG_REF_SHIFT = that size, 0 = 4096
                                       * (cabd.c's default) */
#define MSGPU_FLAG_LZX_DELTA    0x2u  /* lzxd_init(is_delta=1): LZX DELTA stream (lzxd.c:289-296, :441-444, :589-611), window_bits 17..25 */
#define MSGPU_FLAG_REF_SHIFT    6     /* flags >> 6 = bytes of LZX DELTA reference data (lzxd_set_reference_data, lzxd.c:348-382;
                                       * <= the window size).  The caller stores them in the OUTPUT buffer directly in front
                                       * of the unit, at [out_off - n, out_off): the reference preloads them at the end of its
                                       * window, i.e. logically just before the first output byte */
#define MSGPU_UNIT_REF_BYTES(u) (((u)->flags & MSGPU_FLAG_LZX_STREAM_BASE) ? 0u : ((u)->flags >> MSGPU_FLAG_REF_SHIFT))
/* An LZX unit that is a LATER reset interval of a longer stream (a CHM content section cut at its reset table, msgpu_chm.h): the
 * reference, which decodes the section as one stream, keeps counting frames and bytes across resets - lzx->frame for the "no E8
 * translation from frame 32768 on" rule and lzx->offset for the E8 call offsets (lzxd.c:706-712) - while everything else starts
 * afresh at a reset (lzxd.c:423-438).  flags >> MSGPU_FLAG_REF_SHIFT = the index of the unit's first frame within its stream (not
 * combinable with MSGPU_FLAG_LZX_DELTA / reference data).  intel_started, the third thing that survives a reset, needs no
 * carrying: an interval can only contain a 0xE8 byte after one of ITS blocks coded that literal or was stored, which sets it. */
#define MSGPU_FLAG_LZX_STREAM_BASE 0x20u
#define MSGPU_UNIT_FRAME_BASE(u) (((u)->flags & MSGPU_FLAG_LZX_STREAM_BASE) ? ((u)->flags >> MSGPU_FLAG_REF_SHIFT) : 0u)
/* MSZIP block chains (intra-folder parallelism).  The CK blocks of one MSZIP folder are independent
 * bitstreams - only their match SOURCES reach into the previous block's 32 KiB (mszipd.c:267-268) - so a caller that knows
 * where every block starts (a cabinet's CFDATA table does) can hand them over as consecutive units: CHAIN_FIRST for the first
 * block, CHAIN_NEXT for each following one.  The entropy stage then decodes all blocks in parallel; the resolve stage walks a
 * chain in order.  Rules (msgpu_decode_batch_* returns MSGPU_ERR_ARGS otherwise): codec MSZIP; a NEXT unit directly follows
 * its predecessor in the unit array; every unit but the chain's last produces exactly 32768 bytes; out_off continues where
 * the predecessor's output ends.  Each unit must be exactly one CK block that uses up exactly its in_len bytes and produces
 * exactly out_len bytes; a unit for which that does not hold (or that fails in any other way) reports
 * MSGPU_ERR_CHAIN, and the caller decodes the folder as ONE plain unit to get the reference's result (msgpu_cab.cu does). */
#define MSGPU_FLAG_CHAIN_FIRST  0x4u
#define MSGPU_FLAG_CHAIN_NEXT   0x8u
#define MSGPU_ERR_CHAIN         100   /* not an MSPACK_ERR_*: "decode this chain as one stream instead" */
/* MSZIP inside a KWAJ file (mszipd_decompress_kwaj, mszipd.c:462-495; caller kwajd.c:320-322): every block is preceded by a 16-bit
 * length and the stream ends with a zero length - the amount of output is not known beforehand.  out_len is the CAPACITY of the
 * unit's output area; the unit ends at the zero length with status 0 and msgpu_last_produced() says how much it produced, or with
 * MSGPU_ERR_CAPACITY when the area is too small (decode again with a larger one). */
#define MSGPU_FLAG_MSZIP_KWAJ   0x10u
#define MSGPU_ERR_CAPACITY      101   /* not an MSPACK_ERR_* */

/* One independent compressed unit.  32 bytes, no padding. */
typedef struct msgpu_unit {
    uint8_t  codec;           /* MSGPU_CODEC_*                                        */
    uint8_t  window_bits;     /* LZX 15..21, LZX DELTA 17..25 (lzxd.c:289-296), Quantum 10..21 (qtmd.c:199) */
    uint16_t reset_interval;  /* LZX: frames between resets, 0 = never (lzxd.c:423)   */
    uint32_t flags;           /* MSGPU_FLAG_*                                         */
    uint64_t in_off;          /* byte offset of the unit's compressed bytes           */
    uint32_t in_len;          /* compressed byte count                                */
    uint32_t out_len;         /* bytes to produce == X_decompress(state, out_len)     */
    uint64_t out_off;         /* byte offset of the unit's output; multiple of 16.  The unit owns
                               * [out_off, out_off + out_len); the buffer doubles as its sliding window and is
                               * written in two passes (literals, then matches), so on a unit that FAILS the
                               * bytes after its last completely decoded frame are unspecified */
} msgpu_unit;

typedef struct msgpu_ctx msgpu_ctx;

/* Create / destroy a decoder context on CUDA device `device` (scratch pools, streams).
 * Returns NULL on failure (no CUDA device, out of memory): there is no CPU fallback. */
msgpu_ctx *msgpu_create(int device);
void       msgpu_destroy(msgpu_ctx *ctx);

/* Human-readable text for the last failure on this context ("" if none). */
const char *msgpu_last_error(const msgpu_ctx *ctx);

/* Decode n units whose compressed bytes are ALREADY RESIDENT in device memory.
 *   units     host array of n descriptors (copied to the device by the call)
 *   d_in      device pointer, base for units[i].in_off
 *   d_out     device pointer, base for units[i].out_off (16-byte aligned)
 *   d_status  device pointer to n int32 (MSGPU_ERR_* per unit), may be NULL
 *   stream    a cudaStream_t passed as void* (NULL = the context's own stream)
 * Returns 0 when the work was enqueued, nonzero MSGPU_ERR_* on argument / launch failure.
 * Asynchrony: a batch of LZX / Quantum units that fits one wave (the scratch budget: ~100 000 units on a B200) is queued on
 * `stream` without waiting for the device - the unit table travels through pinned staging memory.  The call does wait
 * (a) for the previous wave's table upload before it reuses the staging memory (batches larger than one wave, back-to-back
 * calls: a wait for a COPY, not for kernels), and (b) per launch round in waves that hold MSZIP units, whose CK blocks may
 * be shorter than 32 KiB, so that the number of rounds a folder needs is only known from a counter read back from the device.
 * Results are complete when `stream` is. */
int msgpu_decode_batch_device(msgpu_ctx *ctx, const msgpu_unit *units, size_t n,
                              const void *d_in, size_t in_bytes,
                              void *d_out, size_t out_bytes,
                              int32_t *d_status, void *stream);

/* Same, but with the unit table already in device memory (no host copy at all). */
int msgpu_decode_batch_device_units(msgpu_ctx *ctx, const msgpu_unit *d_units, size_t n,
                                    const void *d_in, size_t in_bytes,
                                    void *d_out, size_t out_bytes,
                                    int32_t *d_status, void *stream);

/* End-to-end call with HOST buffers: copies units + input to the device, decodes,
 * copies output and status back, synchronises.  `status` may be NULL.
 * Returns 0 if the batch ran (per-unit results are in status[]), nonzero on failure. */
int msgpu_decode_batch_host(msgpu_ctx *ctx, const msgpu_unit *units, size_t n,
                            const void *h_in, size_t in_bytes,
                            void *h_out, size_t out_bytes, int32_t *status);

/* Output sinks on the device: a digest of every unit's decoded bytes instead of the bytes.  The reference's own
 * tests verify extraction this way (test/md5_fh.h:72-77 + cabd_test.c:472-478: MD5 of what cabd writes; oabd.c:98: running CRC-32,
 * mspack/crc32.h); for a caller that only verifies, the device-to-host copy of the output - what bounds msgpu_decode_batch_host -
 * shrinks to 16 / 4 bytes per unit.
 *   MSGPU_DIGEST_MD5    16 bytes per unit (RFC 1321, as printed by md5sum)
 *   MSGPU_DIGEST_CRC32  4 bytes per unit, little-endian: crc32(0xFFFFFFFF, data, len) ^ 0xFFFFFFFF in the terms of mspack/crc32.h (= zlib's)
 * A unit whose status is nonzero gets an all-zero digest (its bytes are unspecified).
 *   msgpu_digest_device             digests of units already decoded in device memory (d_status may be NULL = digest every unit);
 *                                   queued on `stream` like msgpu_decode_batch_device
 *   msgpu_decode_batch_host_digest  msgpu_decode_batch_host without the output copy: compressed bytes in, digests + status out
 *                                   (the output lives in the context's own device buffer; not for LZX DELTA units with reference data) */
#define MSGPU_DIGEST_MD5   1
#define MSGPU_DIGEST_CRC32 2
int msgpu_digest_device(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *d_out, size_t out_bytes, const int32_t *d_status,
                        int kind, void *d_digest, void *stream);
int msgpu_decode_batch_host_digest(msgpu_ctx *ctx, const msgpu_unit *units, size_t n, const void *h_in, size_t in_bytes, size_t out_bytes,
                                   int kind, void *h_digest, int32_t *status);

/* Several GPUs, one call.  Units are independent, so devices share a batch by unit index with no exchange step (the reference's
 * unit of independence: a CAB folder, cabd.c:1142-1177; a CHM reset interval, chmd.c:1146-1186):
 *   msgpu_shard_range            shard `shard` of `nshards` owns units [*lo, *hi) = [floor(shard n / nshards), floor((shard + 1) n /
 *                                nshards)), moved forward where that would cut an MSZIP block chain (the same split bench.py's ranks use)
 *   msgpu_decode_batch_host_multi one thread + one context (ctxs[d], created on ndev different devices) per shard; each device copies
 *                                in / out only its shard's bytes.  status[] and the return value as for msgpu_decode_batch_host. */
int msgpu_shard_range(const msgpu_unit *units, size_t n, int shard, int nshards, size_t *lo, size_t *hi);
int msgpu_decode_batch_host_multi(msgpu_ctx *const *ctxs, int ndev, const msgpu_unit *units, size_t n,
                                  const void *h_in, size_t in_bytes, void *h_out, size_t out_bytes, int32_t *status);

/* Bytes every unit of the most recent batch produced (complete frames only if the unit failed): produced[0..n).  Valid for a
 * batch that ran as one wave (n units; up to several thousand units always do, larger ones as far as the scratch budget
 * reaches); synchronises with the batch.  Returns 0, or MSGPU_ERR_ARGS if the last batch was not a single wave of n units. */
int msgpu_last_produced(msgpu_ctx *ctx, uint32_t *produced, size_t n);

/* Number of kernel launches issued by this context so far (bench.py gpu_launches). */
uint64_t msgpu_launch_count(const msgpu_ctx *ctx);

/* Bytes of device scratch the context currently holds. */
size_t msgpu_scratch_bytes(const msgpu_ctx *ctx);

/* Host only (no device needed): how msgpu_decode_batch_* cuts a batch under a scratch budget.  *fmax = frames per launch round
 * for long units (2 for big batches of short units, up to 64 for batches of few long units - a cabinet's multi-megabyte LZX /
 * Quantum folders decode in order on one lane, DESIGN.md section 7 "Long units"; the MSGPU_FMAX environment variable overrides),
 * *frame_slots = record arrays the batch needs in all, *waves = passes over the scratch memory, *rounds = launch rounds of the
 * longest unit as planned (MSZIP folders of short CK blocks may take more).  Any out pointer may be NULL. */
int msgpu_plan_batch(const msgpu_unit *units, size_t n, size_t scratch_budget_bytes,
                     uint32_t *fmax, uint64_t *frame_slots, uint32_t *waves, uint32_t *rounds);

/* Milliseconds spent in the decode kernels of the most recent batch, measured with CUDA
 * events on the launching stream (valid after the stream is synchronised; < 0 if none). */
float msgpu_last_kernel_ms(msgpu_ctx *ctx);

/* Stage timing (measurement aid): when on, the stages of the next batches run back to back on one stream,
 * each launch bracketed by CUDA events; msgpu_stage_ms(ctx, stage) then returns the summed duration of
 * stage 0 = P1 entropy kernels, 1 = P2 resolve kernels (LZX: including the E8 call translation, their epilogue) for the most
 * recent batch; stage 2 (the separate E8 kernel of round 1) no longer exists and reads 0. */
int   msgpu_set_stage_timing(msgpu_ctx *ctx, int on);
float msgpu_stage_ms(msgpu_ctx *ctx, int stage);

/* Library version string. */
const char *msgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MSGPU_H */
