/* msgpu_p2.cuh - P2 "resolve" stage (one warp per unit) and the LZX E8 post-pass.
 *
 * Input per frame: the match records sorted by position; the literal bytes already sit at their final positions in
 * the output buffer (P1 stores them in place).  The byte at frame position q is
 *     literal  what P1 stored at q                if no record covers q
 *     match    byte (q - off) of the output       otherwise, where for an overlapping match (off < len) the
 *              source folds back into the off bytes in front of the match (the byte-serial copy of
 *              lzxd.c:636-646 / mszipd.c:271-296 / qtmd.c:391-416 replicates that seed pattern).
 * Each lane resolves 16 consecutive bytes of a 512-byte chunk.  Pass A turns every position of the chunk into a source
 * descriptor in shared memory (lanes first mark their own 16 positions "literal", then the records that reach into the
 * chunk are dealt out to the lanes and overwrite the positions they cover); pass B fetches the bytes: sources in earlier
 * chunks are read back from the output buffer (it is the sliding window); a source inside the current chunk is followed
 * through the descriptors to ITS source until it leaves the chunk or hits a literal (pointer jumping; positions strictly
 * decrease so it terminates).  The chunk is then stored with 16-byte stores.
 */
#pragma once
#include "msgpu_core.cuh"

#ifndef P2_WIN
#define P2_WIN   288         /* records held in shared memory per warp; > 257 so one load always covers a chunk */
#endif
#define P2_CHUNK 512u

MS_D uint32_t rec_pos(uint32_t a) { return a & 0xFFFFu; }
MS_D uint32_t rec_off(uint32_t b) { return b & 0x3FFFFFu; }
/* WIDE (batches with LZX DELTA units, window up to 2^25): offset bits 22.. travel in the upper half of a (emit_match_wide) */
template <bool WIDE> MS_D uint32_t rec_off_w(uint32_t a, uint32_t b) { return WIDE ? ((b & 0x3FFFFFu) | ((a >> 16) << 22)) : (b & 0x3FFFFFu); }
MS_D uint32_t rec_len(uint32_t b) { return b >> 22; }

/* Per-byte source descriptor (pass A -> pass B), one u32 per position of the chunk:
 *     value & 0x7FFFFFFF = (frame-relative position of the byte to copy) + P2_SBIAS; the position may lie in earlier frames
 *                          of the unit, i.e. be negative; overlapping matches are already folded
 *     bit 31 set         : final - the byte is a literal P1 stored at that very position (never followed further) */
#define P2_SBIAS ((WIDE) ? (1 << 26) : (1 << 22))          /* > the largest match offset: 2^21, or 2^25 with LZX DELTA units (WIDE) */
#define P2_LIT   0x80000000u
/* descriptor of chunk position p (0..511) lives at P2_SIDX(p) = p + p / 16: a row of 17 words per lane, so the 32
 * lanes, which all touch "their k-th byte" at the same time, hit 32 different banks (17 is odd) */
#define P2_SIDX(p) ((p) + ((p) >> 4))
#define P2_SRC_WORDS (P2_CHUNK + P2_CHUNK / 16)

/* descriptor of frame position p inside the match (pos, off, len) */
template <bool WIDE>
MS_D uint32_t p2_desc(uint32_t p, uint32_t pos, uint32_t off, uint32_t len) {
    uint32_t kk = p - pos;
    if (off >= len || kk < off) return p - off + P2_SBIAS;
    return pos - off + (kk % off) + P2_SBIAS;              /* overlapping match: fold onto the seed bytes in front of it */
}

/* MSZIP ring history (RING instantiation).  The reference's MSZIP window is a 32 KiB ring that every CK block starts
 * writing at index 0 (mszipd.c:416-417), and a match that reaches in front of its block reads window[32768 + posn - dist]
 * (:267-268): whatever the most recent EARLIER block that was long enough left at that index.  As long as every earlier
 * block is a full 32 KiB that is simply the previous block, i.e. the bytes in front of the frame in the linear output; after
 * a shorter block it is not.  P1 (ZipLaneC::frame_start) gives every frame decoded after a short block a snapshot of the
 * ring's history: n entries {len, g0}, most recent block first, lengths strictly increasing - index i belongs to the first
 * entry with len > i and lives at unit position g0 + i; an index no entry covers was never written (reads as zero). */
#define P2_HIST_K     16        /* more than 16 blocks in a row, each shorter than the one before: the unit fails */
#define P2_HIST_WORDS (1 + 2 * P2_HIST_K)      /* n, then {len, g0} pairs */
#define P2_PLANE_REC  10944                    /* an MSZIP overflow frame (ZipLaneC::qbase) keeps its literal bytes in a 32 KiB plane inside its record
                                                * array, records [10944, 15040): they may not go to their output positions before the block's first
                                                * 32 KiB - whose literals live there - has been resolved */
#define P2_HIST_REC   (MS_MAXREC - 17)         /* the snapshot sits in the spare tail of the frame's record array (an MSZIP frame has at
                                                * most 32768 / 3 records) */
MS_D uint32_t p2_ring_lookup(const uint32_t *hist, uint32_t i, const uint8_t *unit_out) {
    const uint32_t n = hist[0];
    for (uint32_t j = 0; j < n; j++) if (hist[1 + 2 * j] > i) return unit_out[(size_t) hist[2 + 2 * j] + i];
    return 0;
}

#define P2_LONG      48      /* a match covering at least this many positions of the chunk is filled by the whole warp */
#define P2_LONG_MAX  16

/* Pass A, step 1: every lane marks its own 16 positions as literals (call before a warp sync) */
template <bool WIDE>
MS_D void p2_pass_a_literals(uint32_t q0, uint32_t c, uint32_t *src) {
    uint32_t *row = src + P2_SIDX(q0 - c);
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) row[k] = P2_LIT | (q0 + k + P2_SBIAS);
}

/* Pass A, step 2, RECORD-parallel: the records that intersect the chunk [c, cend) are dealt out to the lanes
 * (r_lo + lane, + 32, ...); a lane overwrites the descriptors of the positions "its" match covers inside the chunk.
 * The record is decoded once, the per-position work is a store.  Matches with a long span (257-byte matches of repetitive
 * data) are queued in shared memory and filled by all 32 lanes together.  longq[0] = count, longq[1..] = window indices.
 * Returns the first window index this lane saw whose match ends beyond cend (P2_WIN if none): the minimum over the warp is
 * the next chunk's r_lo. */
template <bool WIDE, int WS = 1>      /* WS: words between two records of the window (2: records as they lie in memory, wb = wa + 1) */
MS_D int p2_pass_a_records(int lane, int r_lo, uint32_t c, uint32_t cend, const uint32_t *wa, const uint32_t *wb,
                           uint32_t *src, uint32_t *longq)
{
    int next_lo = P2_WIN;
#pragma unroll 1
    for (int r = r_lo + lane; r < P2_WIN; r += 32) {
        uint32_t a = wa[r * WS], pos = rec_pos(a), b = wb[r * WS], off = rec_off_w<WIDE>(a, b), len = rec_len(b), end = pos + len;
        if (pos >= cend) { if (next_lo == P2_WIN) next_lo = r; break; }               /* (the sentinel has pos >= size >= cend) */
        if (end > cend && next_lo == P2_WIN) next_lo = r;
        uint32_t p = pos > c ? pos : c, p1 = end < cend ? end : cend;
        if (p1 <= p) continue;
        if (p1 - p >= P2_LONG) {
#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
            uint32_t slot = atomicAdd(&longq[0], 1u);
#else
            uint32_t slot = longq[0]++;
#endif
            if (slot < P2_LONG_MAX) { longq[1 + slot] = (uint32_t) r; continue; }
        }
        if (off >= len) {
            uint32_t sv = p - off + P2_SBIAS;
#pragma unroll 1
            for (; p < p1; p++, sv++) src[P2_SIDX(p - c)] = sv;                       /* plain match */
        }
        else {
#pragma unroll 1
            for (; p < p1; p++) src[P2_SIDX(p - c)] = p2_desc<WIDE>(p, pos, off, len);      /* overlapping match */
        }
    }
    return next_lo;
}
/* Pass A, step 3: the queued long matches, all lanes together (call after a warp sync) */
template <bool WIDE, int WS = 1>
MS_D void p2_pass_a_long(int lane, uint32_t c, uint32_t cend, const uint32_t *wa, const uint32_t *wb, uint32_t *src, const uint32_t *longq)
{
    uint32_t nl = longq[0] < P2_LONG_MAX ? longq[0] : P2_LONG_MAX;
#pragma unroll 1
    for (uint32_t i = 0; i < nl; i++) {
        int r = (int) longq[1 + i];
        uint32_t a = wa[r * WS], pos = rec_pos(a), b = wb[r * WS], off = rec_off_w<WIDE>(a, b), len = rec_len(b);
        uint32_t p0 = pos > c ? pos : c, p1 = pos + len < cend ? pos + len : cend;
#pragma unroll 1
        for (uint32_t p = p0 + (uint32_t) lane; p < p1; p += 32) src[P2_SIDX(p - c)] = p2_desc<WIDE>(p, pos, off, len);
    }
}

template <bool WIDE, bool RING = false, bool PLANE = false>
MS_D void p2_pass_b(uint32_t q0, uint32_t c, uint32_t size, const uint32_t *src, const uint8_t *unit_out, uint32_t g0, uint32_t w[4], uint32_t ref_len = 0,
                    const uint32_t *hist = nullptr, const uint8_t *plane = nullptr)
{
    w[0] = w[1] = w[2] = w[3] = 0;
    if (q0 >= size) return;
    const uint32_t n = size - q0 < 16 ? size - q0 : 16;
    const uint32_t inchunk = c + P2_SBIAS;
    const uint32_t *row = src + P2_SIDX(q0 - c);
    uint32_t d[16];
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t x = (k < n) ? row[k] : P2_LIT;
#pragma unroll 1
        while ((int32_t) x >= (int32_t) inchunk) x = src[P2_SIDX(x - inchunk)];
        d[k] = PLANE ? x : (x & ~P2_LIT);         /* (PLANE keeps the literal flag: an overflow frame's literals sit in its plane) */
    }
    const uint8_t *obase = unit_out + ((int64_t) g0 - P2_SBIAS);
    uint32_t ulim = g0 < (uint32_t) P2_SBIAS ? (uint32_t) P2_SBIAS - g0 : 0u;    /* descriptors below this lie before the unit */
    if (WIDE) ulim = ulim > ref_len ? ulim - ref_len : 0u;
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t v = 0, x = d[k];
        if (RING && PLANE) {
            const bool lit = (int32_t) x < 0;
            x &= ~P2_LIT;
            if (k < n) v = lit ? plane[x - (uint32_t) P2_SBIAS]
                               : (x >= (uint32_t) P2_SBIAS ? obase[x] : p2_ring_lookup(hist, MS_FRAME + x - (uint32_t) P2_SBIAS, unit_out));
        }
        else if (RING) {
            /* a source in front of the frame is the ring index 32768 + (position - frame start), see p2_ring_lookup */
            if (k < n) v = x >= (uint32_t) P2_SBIAS ? obase[x] : p2_ring_lookup(hist, MS_FRAME + x - (uint32_t) P2_SBIAS, unit_out);
        }
        else if (k < n && x >= ulim) v = obase[x];
        w[k >> 2] |= v << (8 * (k & 3));
    }
}

#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
/* Resolve one frame with one warp.  wa/wb: this warp's P2_WIN-entry windows in shared memory, loaded by the lanes (the ring / plane /
 * chain kernels and MSGPU_P2_BULK=0; the main resolve kernel uses p2_resolve_frame_pipe below). */
#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
__device__ __forceinline__ uint32_t p2_smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void p2_mbar_init(uint64_t *mbar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(p2_smem_addr(mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void p2_bulk_load(void *dst, const void *gsrc, uint32_t bytes, uint64_t *mbar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* earlier generic-proxy reads of the window are done before the async proxy overwrites it */
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(p2_smem_addr(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(p2_smem_addr(dst)), "l"(gsrc), "r"(bytes), "r"(p2_smem_addr(mbar)) : "memory");
}
__device__ __forceinline__ void p2_mbar_wait(uint64_t *mbar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(p2_smem_addr(mbar)), "r"(parity) : "memory");
}
#endif
template <bool WIDE, bool RING = false, bool PLANE = false>
__device__ __forceinline__ void p2_resolve_frame(int lane, const MsRec *recs, uint32_t nrec, uint32_t size, uint8_t *unit_out, uint32_t g0,
                                                 uint32_t *wa, uint32_t *wb, uint32_t *src, uint32_t *longq, uint32_t ref_len, const uint32_t *hist = nullptr,
                                                 const uint8_t *plane = nullptr)
{
    constexpr int WS = 1;
    uint32_t wbase = 0, wcover = 0; bool loaded = false;
    int r_lo = 0;                                              /* first window record ending beyond the chunk start */
    for (uint32_t c = 0; c < size; c += P2_CHUNK) {
        const bool reload = !loaded || (c + P2_CHUNK > wcover && wcover < size);
        uint32_t q0 = c + 16u * (uint32_t) lane, w[4];
        const uint32_t cend = c + P2_CHUNK < size ? c + P2_CHUNK : size;
        if (reload) {
            wbase += (uint32_t) r_lo; r_lo = 0;
            __syncwarp();
            for (int j = lane; j < P2_WIN; j += 32) {
                uint32_t r = wbase + (uint32_t) j; if (r > nrec) r = nrec;      /* nrec = the sentinel */
                MsRec x = recs[r]; wa[j] = x.a; wb[j] = x.b;
            }
            __syncwarp();
            wcover = rec_pos(wa[P2_WIN - 1]); loaded = true;
        }
        if (lane == 0) longq[0] = 0;
        p2_pass_a_literals<WIDE>(q0, c, src);
        __syncwarp();
        int nlo = p2_pass_a_records<WIDE, WS>(lane, r_lo, c, cend, wa, wb, src, longq);
        r_lo = __reduce_min_sync(0xFFFFFFFFu, nlo);            /* also orders the descriptor stores (it is a warp sync) */
        __syncwarp();
        p2_pass_a_long<WIDE, WS>(lane, c, cend, wa, wb, src, longq);
        __syncwarp();
        p2_pass_b<WIDE, RING, PLANE>(q0, c, size, src, unit_out, g0, w, ref_len, hist, plane);
        uint8_t *dst = unit_out + (size_t) g0 + q0;
        if (PLANE) __syncwarp();     /* an MSZIP overflow frame reads the bytes it is about to replace (ZipLaneC::qbase): every load before any store */
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

/* The main resolve kernel's frame resolve.
 * (1) The record window is filled by the copy engine (cp.async.bulk into shared memory, completion on an mbarrier) instead of
 * nine 8-byte loads per lane; the chunk's literal descriptors are written while the copy is in flight.  The window holds the
 * records as they lie in memory (wa = window, wb = wa + 1, stride 2) and starts at an even record index (16-byte source alignment).
 * (2) It is SOFTWARE-PIPELINED over the chunks.  The unpipelined loop was latency bound at the 32 warps per SM its registers
 * allow - 3.3 of the 11 cycles per issued instruction went into waiting for the scattered byte loads of the match sources
 * (profiles/r2_p2_f.txt; 0.4 of 9.6 now, profiles/r2_p2_pipe_k.txt) - and nothing in pass A of chunk c + 1 depends on the bytes of chunk c.  So the 16
 * byte loads of chunk c are ISSUED, pass A of chunk c + 1 runs while they are in flight, and only then are the bytes packed and
 * stored.  One descriptor array is enough: pass A of c + 1 may overwrite it as soon as every lane has finished its pointer jumps
 * of chunk c (a warp sync), and the byte loads need registers only.  Sources of chunk c + 1 that lie in chunk c are read after
 * the store of chunk c and the sync behind it, as before. */
template <bool WIDE>
__device__ __forceinline__ void p2_resolve_frame_pipe(int lane, const MsRec *recs, uint32_t nrec, uint32_t size, uint8_t *unit_out, uint32_t g0,
                                                      uint32_t *wa, uint32_t *src, uint32_t *longq, uint32_t ref_len, uint64_t *mbar, uint32_t *mphase)
{
    constexpr int WS = 2;
    uint32_t *wb = wa + 1;
    uint32_t wbase = 0, wcover = 0; bool loaded = false;
    int r_lo = 0;
    if (size == 0) return;
    auto pass_a = [&](uint32_t c) {
        const bool reload = !loaded || (c + P2_CHUNK > wcover && wcover < size);
        const uint32_t q0 = c + 16u * (uint32_t) lane, cend = c + P2_CHUNK < size ? c + P2_CHUNK : size;
        uint32_t cnt = 0;
        if (reload) {
            wbase += (uint32_t) r_lo; r_lo = (int) (wbase & 1u); wbase &= ~1u;
            cnt = MS_MAXREC - wbase < (uint32_t) P2_WIN ? MS_MAXREC - wbase : (uint32_t) P2_WIN;
            __syncwarp();
            if (lane == 0) p2_bulk_load(wa, recs + wbase, cnt * 8u, mbar);
        }
        if (lane == 0) longq[0] = 0;
        p2_pass_a_literals<WIDE>(q0, c, src);
        if (reload) {
            p2_mbar_wait(mbar, *mphase & 1u); *mphase += 1u;
            for (int j = lane; j < P2_WIN; j += 32) if (wbase + (uint32_t) j > nrec || (uint32_t) j >= cnt) { wa[2 * j] = size; wa[2 * j + 1] = 0; }
            __syncwarp();
            wcover = rec_pos(wa[2 * (P2_WIN - 1)]); loaded = true;
        }
        __syncwarp();
        const int nlo = p2_pass_a_records<WIDE, WS>(lane, r_lo, c, cend, wa, wb, src, longq);
        r_lo = __reduce_min_sync(0xFFFFFFFFu, nlo);
        __syncwarp();
        p2_pass_a_long<WIDE, WS>(lane, c, cend, wa, wb, src, longq);
        __syncwarp();
    };
    pass_a(0);
    const uint8_t *obase = unit_out + ((int64_t) g0 - P2_SBIAS);
    uint32_t ulim = g0 < (uint32_t) P2_SBIAS ? (uint32_t) P2_SBIAS - g0 : 0u;
    if (WIDE) ulim = ulim > ref_len ? ulim - ref_len : 0u;
    for (uint32_t c = 0; c < size; c += P2_CHUNK) {
        const uint32_t q0 = c + 16u * (uint32_t) lane;
        const uint32_t n = q0 < size ? (size - q0 < 16 ? size - q0 : 16) : 0u;
        const uint32_t inchunk = c + P2_SBIAS;
        const uint32_t *row = src + P2_SIDX(q0 - c);
        uint32_t v[16];
#pragma unroll
        for (uint32_t k = 0; k < 16; k++) {                      /* pointer jumps (p2_pass_b) */
            uint32_t x = (k < n) ? row[k] : P2_LIT;
#pragma unroll 1
            while ((int32_t) x >= (int32_t) inchunk) x = src[P2_SIDX(x - inchunk)];
            v[k] = x & ~P2_LIT;
        }
        __syncwarp();                                            /* nobody reads this chunk's descriptors any more */
#pragma unroll
        for (uint32_t k = 0; k < 16; k++) v[k] = (k < n && v[k] >= ulim) ? (uint32_t) obase[v[k]] : 0u;      /* issued, not yet needed */
        if (c + P2_CHUNK < size) pass_a(c + P2_CHUNK);
#pragma unroll
        for (uint32_t k = 0; k < 16; k++) asm volatile("" : "+r"(v[k]) :: "memory");      /* (keeps the first use of a loaded byte behind pass A: ptxas would hoist the shifts) */
        uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (uint32_t k = 0; k < 16; k++) w[k >> 2] |= v[k] << (8 * (k & 3));
        uint8_t *dst = unit_out + (size_t) g0 + q0;
        if (n == 16 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        else for (uint32_t k = 0; k < n; k++) dst[k] = (uint8_t) (w[k >> 2] >> (8 * (k & 3)));
        __syncwarp();                                            /* the chunk is visible to the whole warp before anyone reads it back */
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
