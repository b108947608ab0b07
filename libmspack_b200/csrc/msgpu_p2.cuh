/* msgpu_p2.cuh - P2 "resolve" stage (one warp per unit) and the LZX E8 post-pass.
 *
 * Input per frame: match records sorted by position + a literal byte stream (msgpu_core.cuh).
 * The byte at frame position q is
 *     literal  lits[q - M_i]                  if q lies before record i's match (M_i = match bytes before i)
 *     match    byte (q - off) of the output   otherwise, where for an overlapping match (off < len) the
 *              source folds back into the off bytes in front of the match (the byte-serial copy of
 *              lzxd.c:636-646 / mszipd.c:271-296 / qtmd.c:391-416 replicates that seed pattern).
 * Each lane resolves 16 consecutive bytes of a 512-byte chunk.  Pass A maps every position of the
 * chunk to its record (one binary search per lane, then a walk); pass B fetches the bytes: sources in
 * earlier chunks are read back from the output buffer (it is the sliding window); a source inside the
 * current chunk is followed through the position map to ITS source until it leaves the chunk or hits
 * a literal (pointer jumping; positions strictly decrease so it terminates).  The chunk is then
 * stored with 16-byte stores.
 */
#pragma once
#include "msgpu_core.cuh"

#define P2_WIN   288         /* records held in shared memory per warp; > 257 so one load always covers a chunk */
#define P2_CHUNK 512u

MS_D uint32_t rec_pos(uint32_t a) { return a & 0xFFFFu; }
MS_D uint32_t rec_M(uint32_t a)   { return a >> 16; }
MS_D uint32_t rec_off(uint32_t b) { return b & 0x3FFFFFu; }
MS_D uint32_t rec_len(uint32_t b) { return b >> 22; }

/* first window index whose match END lies beyond q (ends are non-decreasing; the last entry's end is > q) */
MS_D int p2_search(const uint32_t *wa, const uint32_t *wb, uint32_t q) {
    int lo = 0, hi = P2_WIN - 1;
#pragma unroll 1
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (rec_pos(wa[mid]) + rec_len(wb[mid]) > q) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* Pass A of a chunk: this lane's 16 positions [q0, q0+16) -> owning record (window index) into rid[]. */
MS_D void p2_pass_a(uint32_t q0, uint32_t c, uint32_t size, const uint32_t *wa, const uint32_t *wb, uint16_t *rid)
{
    if (q0 >= size) return;
    int i = p2_search(wa, wb, q0);
#pragma unroll 4
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t q = q0 + k;
        if (q >= size) break;
        while (rec_pos(wa[i]) + rec_len(wb[i]) <= q) i++;
        rid[q - c] = (uint16_t) i;
    }
}

/* One output byte whose source may lie inside the current chunk: follow the position map to ITS source
 * until a literal or an already-written byte is reached (positions strictly decrease). */
MS_D uint32_t p2_chase(uint32_t q, uint32_t c, const uint32_t *wa, const uint32_t *wb, const uint16_t *rid,
                       const uint8_t *lits, const uint8_t *unit_out, uint32_t g0)
{
#pragma unroll 1
    for (;;) {
        int i = rid[q - c];
        uint32_t a = wa[i], b = wb[i], pos = rec_pos(a);
        if (q < pos) return lits[q - rec_M(a)];
        uint32_t off = rec_off(b), len = rec_len(b), kk = q - pos;
        int32_t s;                                            /* frame-relative source position, may be negative */
        if (off < len && kk >= off) s = (int32_t) (pos - off + (kk % off));
        else s = (int32_t) q - (int32_t) off;
        if (s < (int32_t) c) {
            int64_t g = (int64_t) g0 + s;                     /* unit-relative */
            return g >= 0 ? unit_out[g] : 0u;                 /* before the unit's first byte: defined as zero */
        }
        q = (uint32_t) s;
    }
}

/* Pass B: the 16 bytes [q0, q0+16) of the frame (positions >= size give 0), little-endian in 4 words.
 * Every lane runs the same per-byte loop (keeps the warp converged); a source inside the current chunk
 * is chased through the position map. */
MS_D void p2_pass_b(uint32_t q0, uint32_t c, uint32_t size, const uint32_t *wa, const uint32_t *wb, const uint16_t *rid,
                    const uint8_t *lits, const uint8_t *unit_out, uint32_t g0, uint32_t w[4])
{
    w[0] = w[1] = w[2] = w[3] = 0;
    if (q0 >= size) return;
#pragma unroll 4
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t q = q0 + k;
        if (q >= size) break;
        uint32_t v = p2_chase(q, c, wa, wb, rid, lits, unit_out, g0);
        w[k >> 2] |= v << (8 * (k & 3));
    }
}

#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
/* Resolve one frame with one warp.  wa/wb: this warp's P2_WIN-entry windows in shared memory. */
__device__ __forceinline__ void p2_resolve_frame(int lane, const MsRec *recs, uint32_t nrec, const uint8_t *lits,
                                                 uint32_t size, uint8_t *unit_out, uint32_t g0,
                                                 uint32_t *wa, uint32_t *wb, uint16_t *rid)
{
    uint32_t wbase = 0, wcover = 0; bool loaded = false;
    for (uint32_t c = 0; c < size; c += P2_CHUNK) {
        if (!loaded || (c + P2_CHUNK > wcover && wcover < size)) {
            if (loaded) wbase += (uint32_t) p2_search(wa, wb, c);
            __syncwarp();
            for (int j = lane; j < P2_WIN; j += 32) {
                uint32_t r = wbase + (uint32_t) j; if (r > nrec) r = nrec;      /* nrec = the sentinel */
                MsRec x = recs[r]; wa[j] = x.a; wb[j] = x.b;
            }
            __syncwarp();
            wcover = rec_pos(wa[P2_WIN - 1]); loaded = true;
        }
        uint32_t q0 = c + 16u * (uint32_t) lane, w[4];
        p2_pass_a(q0, c, size, wa, wb, rid);
        __syncwarp();
        p2_pass_b(q0, c, size, wa, wb, rid, lits, unit_out, g0, w);
        uint8_t *dst = unit_out + (size_t) g0 + q0;
        if (q0 + 16 <= size && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        else if (q0 < size) {
            uint32_t n = size - q0 < 16 ? size - q0 : 16;
            for (uint32_t k = 0; k < n; k++) dst[k] = (uint8_t) (w[k >> 2] >> (8 * (k & 3)));
        }
        __syncwarp();        /* the chunk is visible to the whole warp before anyone reads it back */
    }
}

/* LZX E8 call translation of one finished frame (lzxd.c:706-737), one warp.  `data` = first byte
 * of the frame, curpos0 = the stream offset of that byte (lzx->offset), filesize = intel_filesize. */
__device__ __forceinline__ void e8_translate_frame(int lane, uint8_t *data, uint32_t frame_size, int32_t curpos0, int32_t filesize)
{
    if (frame_size <= 10) return;
    uint32_t end = frame_size - 10, next_ok = 0;
    for (uint32_t base = 0; base < end; base += 32) {
        uint32_t p = base + (uint32_t) lane;
        bool cand = (p < end) && (data[p] == 0xE8);
        uint32_t m = __ballot_sync(0xFFFFFFFFu, cand), act = 0;
        while (m) {                                            /* an E8 swallows the 4 bytes after it */
            int bit = __ffs((int) m) - 1; m &= m - 1;
            uint32_t pp = base + (uint32_t) bit;
            if (pp >= next_ok) { act |= 1u << bit; next_ok = pp + 5; }
        }
        if ((act >> lane) & 1u) {
            int32_t curpos = curpos0 + (int32_t) p;
            int32_t abs_off = (int32_t) ((uint32_t) data[p + 1] | ((uint32_t) data[p + 2] << 8) | ((uint32_t) data[p + 3] << 16) | ((uint32_t) data[p + 4] << 24));
            if (abs_off >= -curpos && abs_off < filesize) {
                int32_t rel = (abs_off >= 0) ? abs_off - curpos : abs_off + filesize;
                data[p + 1] = (uint8_t) rel; data[p + 2] = (uint8_t) (rel >> 8); data[p + 3] = (uint8_t) (rel >> 16); data[p + 4] = (uint8_t) (rel >> 24);
            }
        }
        __syncwarp();
    }
}
#endif
