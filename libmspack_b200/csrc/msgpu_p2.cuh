/* msgpu_p2.cuh - P2 "resolve" stage (one warp per unit) and the LZX E8 post-pass.
 *
 * Input per frame: match records sorted by position + a literal byte stream (msgpu_core.cuh).
 * The byte at frame position q is
 *     literal  lits[q - M_i]                  if q lies before record i's match (M_i = match bytes before i)
 *     match    byte (q - off) of the output   otherwise, where for an overlapping match (off < len) the
 *              source folds back into the off bytes in front of the match (the byte-serial copy of
 *              lzxd.c:636-646 / mszipd.c:271-296 / qtmd.c:391-416 replicates that seed pattern).
 * Each lane resolves 16 consecutive bytes of a 512-byte chunk.  Pass A turns every position of the chunk
 * into a source descriptor (one binary search per lane, then a walk along the records); pass B fetches the
 * bytes: sources in earlier chunks are read back from the output buffer (it is the sliding window); a
 * source inside the current chunk is followed through the descriptors to ITS source until it leaves the
 * chunk or hits a literal (pointer jumping; positions strictly decrease so it terminates).  The chunk is
 * then stored with 16-byte stores.
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

/* Per-byte source descriptor (pass A -> pass B), one u32 per position of the chunk:
 *     bit 31 set : literal, low bits = index into the frame's literal stream
 *     else       : match,   value   = (frame-relative source position) + P2_SBIAS   (the source may lie in
 *                  earlier frames of the unit, i.e. be negative; overlapping matches are already folded) */
#define P2_SBIAS (1 << 22)
/* descriptor of chunk position p (0..511) lives at P2_SIDX(p) = p + p / 16: a row of 17 words per lane, so the 32
 * lanes, which all touch "their k-th byte" at the same time, hit 32 different banks (17 is odd) */
#define P2_SIDX(p) ((p) + ((p) >> 4))
#define P2_SRC_WORDS (P2_CHUNK + P2_CHUNK / 16)

/* descriptor of frame position p for record (pos, M, off, len): literal before the match, else (folded) match source */
MS_D uint32_t p2_desc(uint32_t p, uint32_t pos, uint32_t M, uint32_t off, uint32_t len) {
    if (p < pos) return (p - M) | 0x80000000u;
    uint32_t kk = p - pos;
    if (off >= len || kk < off) return p - off + P2_SBIAS;
    return pos - off + (kk % off) + P2_SBIAS;              /* overlapping match: fold onto the seed bytes in front of it */
}

#define P2_LONG      48      /* a record covering at least this many positions of the chunk is filled by the whole warp */
#define P2_LONG_MAX  16

/* Pass A of a chunk, RECORD-parallel: the records that intersect the chunk [c, cend) are dealt out to the lanes
 * (r_lo + lane, + 32, ...); a lane writes the descriptors of "its" record - the literal run in front of the match
 * and the match - for the positions inside the chunk.  The record is decoded once, the per-position work is a store.
 * Records with a long span (long literal runs of stored data, 257-byte matches of repetitive data) are queued in
 * shared memory and filled by all 32 lanes together.  longq[0] = count, longq[1..] = window indices. */
MS_D void p2_pass_a_records(int lane, int r_lo, uint32_t c, uint32_t cend, const uint32_t *wa, const uint32_t *wb,
                            uint32_t *src, uint32_t *longq)
{
#pragma unroll 1
    for (int r = r_lo + lane; r < P2_WIN; r += 32) {
        uint32_t lit0 = (r == 0) ? c : rec_pos(wa[r - 1]) + rec_len(wb[r - 1]);     /* window[0]'s predecessors all end at or before c */
        if (lit0 >= cend) break;
        uint32_t a = wa[r], b = wb[r], pos = rec_pos(a), M = rec_M(a), off = rec_off(b), len = rec_len(b);
        uint32_t p0 = lit0 > c ? lit0 : c, p1 = pos + len < cend ? pos + len : cend;
        if (p1 <= p0) continue;
        if (p1 - p0 >= P2_LONG) {
#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
            uint32_t slot = atomicAdd(&longq[0], 1u);
#else
            uint32_t slot = longq[0]++;
#endif
            if (slot < P2_LONG_MAX) { longq[1 + slot] = (uint32_t) r; continue; }
        }
        uint32_t le = pos < p1 ? pos : p1, p = p0;
        uint32_t li = (p - M) | 0x80000000u;
#pragma unroll 1
        for (; p < le; p++, li++) src[P2_SIDX(p - c)] = li;                           /* literal run */
        if (off >= len) {
            uint32_t sv = p - off + P2_SBIAS;
#pragma unroll 1
            for (; p < p1; p++, sv++) src[P2_SIDX(p - c)] = sv;                       /* plain match */
        }
        else {
#pragma unroll 1
            for (; p < p1; p++) src[P2_SIDX(p - c)] = p2_desc(p, pos, M, off, len);   /* overlapping match */
        }
    }
}
/* second half of pass A: the queued long records, all lanes together (call after a warp sync) */
MS_D void p2_pass_a_long(int lane, uint32_t c, uint32_t cend, const uint32_t *wa, const uint32_t *wb, uint32_t *src, const uint32_t *longq)
{
    uint32_t nl = longq[0] < P2_LONG_MAX ? longq[0] : P2_LONG_MAX;
#pragma unroll 1
    for (uint32_t i = 0; i < nl; i++) {
        int r = (int) longq[1 + i];
        uint32_t lit0 = (r == 0) ? c : rec_pos(wa[r - 1]) + rec_len(wb[r - 1]);
        uint32_t a = wa[r], b = wb[r], pos = rec_pos(a), M = rec_M(a), off = rec_off(b), len = rec_len(b);
        uint32_t p0 = lit0 > c ? lit0 : c, p1 = pos + len < cend ? pos + len : cend;
#pragma unroll 1
        for (uint32_t p = p0 + (uint32_t) lane; p < p1; p += 32) src[P2_SIDX(p - c)] = p2_desc(p, pos, M, off, len);
    }
}

/* Pass B: fetch this lane's 16 bytes [q0, q0+16) (little-endian in 4 words; positions >= size give 0).  A source inside
 * the current chunk is followed through the shared descriptors to ITS source (pointer jumping; positions strictly
 * decrease so it terminates).  All chases first, then all byte loads, so the loads overlap. */
MS_D void p2_pass_b(uint32_t q0, uint32_t c, uint32_t size, const uint32_t *src,
                    const uint8_t *lits, const uint8_t *unit_out, uint32_t g0, uint32_t w[4])
{
    w[0] = w[1] = w[2] = w[3] = 0;
    if (q0 >= size) return;
    const uint32_t n = size - q0 < 16 ? size - q0 : 16;
    const uint32_t inchunk = c + P2_SBIAS;
    const uint32_t *row = src + P2_SIDX(q0 - c);
    uint32_t d[16];
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t x = (k < n) ? row[k] : 0x80000000u;
#pragma unroll 1
        while ((int32_t) x >= (int32_t) inchunk) x = src[P2_SIDX(x - inchunk)];    /* literal descriptors are negative as int32 */
        d[k] = x;
    }
    const uint8_t *obase = unit_out + ((int64_t) g0 - P2_SBIAS);
    const bool may_underflow = g0 < (uint32_t) P2_SBIAS;                           /* only a unit's first 4 MiB can reach before the unit */
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t v = 0, x = d[k];
        if (k < n) {
            if (x & 0x80000000u) v = lits[x & 0x7FFFFFFFu];
            else if (may_underflow && (int64_t) g0 + (int64_t) x < (int64_t) P2_SBIAS) v = 0;   /* before the unit's first byte: zero */
            else v = obase[x];
        }
        w[k >> 2] |= v << (8 * (k & 3));
    }
}

#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
/* Resolve one frame with one warp.  wa/wb: this warp's P2_WIN-entry windows in shared memory. */
__device__ __forceinline__ void p2_resolve_frame(int lane, const MsRec *recs, uint32_t nrec, const uint8_t *lits,
                                                 uint32_t size, uint8_t *unit_out, uint32_t g0,
                                                 uint32_t *wa, uint32_t *wb, uint32_t *src, uint32_t *longq)
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
        const uint32_t cend = c + P2_CHUNK < size ? c + P2_CHUNK : size;
        const int r_lo = p2_search(wa, wb, c);                 /* uniform: first record reaching into the chunk */
        if (lane == 0) longq[0] = 0;
        __syncwarp();
        p2_pass_a_records(lane, r_lo, c, cend, wa, wb, src, longq);
        __syncwarp();
        p2_pass_a_long(lane, c, cend, wa, wb, src, longq);
        __syncwarp();
        p2_pass_b(q0, c, size, src, lits, unit_out, g0, w);
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
