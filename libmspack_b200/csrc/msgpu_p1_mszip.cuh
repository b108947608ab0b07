/* msgpu_p1_mszip.cuh - P1 entropy stage for MSZIP units: one lane inflates one unit's "CK" blocks
 * (mszipd.c:154-316 inflate, :91-151 zip_read_lens, :377-460 mszipd_decompress), stores the literal bytes at their
 * output positions and emits one match record per match.  Each CK block (<= 32 KiB of output) is one "frame" of the
 * intermediate form.
 *
 * The lane is a state machine (phases in msgpu_core.cuh): service() does the rare, divergent work (CK signature scan,
 * deflate block headers, code-length tables, stored blocks, frame bookkeeping); step() decodes ONE literal/length
 * symbol (plus its distance) and is what the warp executes in lockstep.  Huffman decoding is table-free: code lengths
 * of the literal/length and distance trees come from 2 x 15 register-resident limits, symbols from a small
 * shared-memory head / global scratch (msgpu_core.cuh "Table-free canonical decoding"); ~0.35 KB of shared memory per lane.
 */
#pragma once
#include "msgpu_core.cuh"
#include "msgpu_p2.cuh"        /* P2_HIST_K: the ring history handed to the resolve stage */

/* per-warp auxiliary block in global scratch (interleaved by lane) */
#define ZIP_AUX_LENS      0                              /* u8  [320][32]  literal/length + distance code lengths */
#define ZIP_AUX_LSORT     (ZIP_AUX_LENS + 320 * 32)      /* u16 [288][32] */
#define ZIP_AUX_DSORT     (ZIP_AUX_LSORT + 288 * 32 * 2) /* u16 [32][32]  */
#define ZIP_AUX_BSORT     (ZIP_AUX_DSORT + 32 * 32 * 2)  /* u16 [32][32]  */
#define ZIP_AUX_LIMIT     (ZIP_AUX_BSORT + 32 * 32 * 2)  /* u32 [3][20][32] */
#define ZIP_AUX_OFFS      (ZIP_AUX_LIMIT + 3 * 20 * 32 * 4) /* u16 [3][20][32] */
#define ZIP_AUX_HIST      (ZIP_AUX_OFFS + 3 * 20 * 32 * 2)  /* u32 [2 * P2_HIST_K + 1][32]  ring history {len, g0}, most recent first; then n | ring << 8 */
#define ZIP_AUX_BYTES     (ZIP_AUX_HIST + (2 * P2_HIST_K + 1) * 32 * 4)

/* HEADN = literal/length symbols (shortest codes first) kept in shared memory */
template <int NT, int HEADN>
struct ZipSharedC {
    uint32_t lbo[17 * NT];                /* literal/length tree: limit[l-1] >> 1 | offs[l] << 16 */
    uint32_t dbo[17 * NT];                /* distance tree; hosts the code-length-code tree while lengths are read */
    uint16_t lhead[HEADN * NT];
    uint16_t dhead[32 * NT];              /* all distance symbols in canonical order */
    uint16_t blim[16 * NT];               /* code-length-code tree limits >> 1 */
    uint16_t cnt[17 * NT];
};

/* KWAJ = the instantiation that also understands MSGPU_FLAG_MSZIP_KWAJ units (mszipd_decompress_kwaj, mszipd.c:462-495) */
template <int NT, int HEADN, bool KWAJ = false>
struct ZipLaneC {
    MsBits b;
    uint32_t *lbo, *dbo; uint16_t *lhead, *dhead, *blim, *cnt;   /* this lane's columns of the shared tables */
    uint8_t *lens;                        /* aux, stride 32 */
    MsHuffAux la, da, ba;                 /* only .sorted is used (global scratch) */
    uint32_t llim[15], dlim[15];          /* limit[1..15] of the two trees, registers */
    /* unit / launch context */
    const msgpu_unit *u; MsRec *recs; uint8_t *uout; MsFrameInfo *finfo;   /* uout = the unit's output buffer */
    MsEmit em;
    uint32_t phase, q, last_block, produced, frame, done; int32_t status;
    int f, max_frames;
    /* The ring history of the reference's 32 KiB window (msgpu_p2.cuh "MSZIP ring history") lives entirely in the lane's aux
     * memory - it is touched at frame boundaries only and must not cost the decode loop a register: entries {len, g0}, most
     * recent block first, then one word n | ring << 8; ring = a block shorter than 32 KiB has been followed by another one, from
     * then on every frame carries a snapshot of the history (in the spare tail of its record array) for k_p2_ring */
    MS_M uint32_t *hist_ptr() const {       /* the aux block is 32-byte aligned and lens = aux + lane */
        const uintptr_t lane = reinterpret_cast<uintptr_t>(lens) & 31u;
        return reinterpret_cast<uint32_t *>(lens - lane + ZIP_AUX_HIST) + lane;
    }

    MS_M void bind(ZipSharedC<NT, HEADN> *sh, int tid, uint8_t *aux_warp, int lane) {
        lbo = sh->lbo + tid; dbo = sh->dbo + tid; lhead = sh->lhead + tid; dhead = sh->dhead + tid; blim = sh->blim + tid; cnt = sh->cnt + tid;
        lens = aux_warp + ZIP_AUX_LENS + lane;
        la.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_LSORT) + lane;
        da.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_DSORT) + lane;
        ba.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_BSORT) + lane;
        uint32_t *lim = reinterpret_cast<uint32_t *>(aux_warp + ZIP_AUX_LIMIT) + lane;
        uint16_t *off = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_OFFS) + lane;
        la.limit = lim; da.limit = lim + 20 * 32; ba.limit = lim + 40 * 32;
        la.offs = off; da.offs = off + 20 * 32; ba.offs = off + 40 * 32;
    }

    /* the block that just ended wrote window[0, len): it shadows every earlier block that was not longer */
    MS_M bool hist_push(uint32_t len, uint32_t g0) {
        if (len == 0) return true;
        uint32_t *hist = hist_ptr();
        const uint32_t stw = hist[2 * P2_HIST_K * 32], hist_n = stw & 0xFFu, ring = stw & ~0xFFu;
        if (len >= MS_FRAME) { hist[0] = len; hist[32] = g0; hist[2 * P2_HIST_K * 32] = ring | 1u; return true; }      /* the usual case: nothing older shows through */
        /* lengths grow with the index, so the entries that stay (longer than the new block) are a suffix [j0, hist_n):
         * new history = {len, g0} followed by that suffix */
        uint32_t j0 = 0;
#pragma unroll 1
        while (j0 < hist_n && hist[2 * j0 * 32] <= len) j0++;
        const uint32_t keep = hist_n - j0;
        if (keep + 1 > P2_HIST_K) { fail(MS_EDECRUNCH); return false; }        /* deeper than the history can describe: refuse rather than guess */
        if (j0 == 0) {
#pragma unroll 1
            for (uint32_t k = keep; k > 0; k--) { hist[2 * k * 32] = hist[2 * (k - 1) * 32]; hist[(2 * k + 1) * 32] = hist[(2 * (k - 1) + 1) * 32]; }
        }
        else if (j0 > 1) {
#pragma unroll 1
            for (uint32_t k = 0; k < keep; k++) { hist[2 * (k + 1) * 32] = hist[2 * (j0 + k) * 32]; hist[(2 * (k + 1) + 1) * 32] = hist[(2 * (j0 + k) + 1) * 32]; }
        }
        hist[0] = len; hist[32] = g0;
        hist[2 * P2_HIST_K * 32] = ring | (keep + 1);
        return true;
    }
    /* frame prologue: does this frame need the ring instantiation of the resolve stage?  If so leave it a snapshot */
    MS_M uint32_t hist_snapshot(MsRec *frame_recs) {
        uint32_t *hist = hist_ptr();
        uint32_t stw = hist[2 * P2_HIST_K * 32]; const uint32_t hist_n = stw & 0xFFu;
        if (hist_n && !(hist_n == 1 && hist[0] >= MS_FRAME)) stw |= 0x100u;       /* an earlier block was short: sticky */
        if (!(stw & 0x100u)) return 0;
        hist[2 * P2_HIST_K * 32] = stw;
        uint32_t *sn = reinterpret_cast<uint32_t *>(frame_recs + P2_HIST_REC);
        sn[0] = hist_n;
#pragma unroll 1
        for (uint32_t j = 0; j < 2 * hist_n; j++) sn[1 + j] = hist[j * 32];
        return 1;
    }

    /* next 16 stream bits, first bit on top (deflate packs Huffman codes starting at the code's MSB) */
    MS_M uint32_t v16() const { return MS_BREV32((uint32_t) b.bb) >> 16; }

    /* mszipd.c:91-151.  Returns 0 or an MSPACK_ERR_* */
    MS_M int read_lens() {
        const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
        lsb_refill(b);
        uint32_t lit_codes = lsb_read(b, 5) + 257, dist_codes = lsb_read(b, 5) + 1, bl_codes = lsb_read(b, 4) + 4;
        if (b.err) return b.err;
        if (lit_codes > 288 || dist_codes > 32) return MS_EDECRUNCH;
        /* 19 code-length-code lengths, 3 bits each, packed into a 64-bit register */
        uint64_t bl = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < bl_codes; i++) { lsb_refill(b); bl |= (uint64_t) lsb_read(b, 3) << (3 * order[i]); }
        if (b.err) return b.err;
        uint32_t lv[16];
        if (ms_canon_build<0, NT>([&](int s) { return (uint32_t) (bl >> (3 * s)) & 7u; }, 19, 7, dbo, cnt, ba.sorted, (uint16_t *) nullptr, 0,
                                  (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
#pragma unroll
        for (int j = 0; j < 15; j++) blim[j * NT] = (uint16_t) (lv[j] >> 1);
        uint32_t total = lit_codes + dist_codes, last_code = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < total;) {
            lsb_refill(b);
            lsb_check(b, 7);                                   /* :117 ENSURE_BITS(7) */
            uint32_t v = v16();
            int cl = ms_canon_len_smem<NT>(blim, v);
            uint32_t code = ba.sorted[ms_canon_index<NT>(dbo, v, cl) * MS_WARP]; lsb_drop(b, cl);
            if (b.err) return b.err;
            if (code < 16) { lens[i * 32] = (uint8_t) code; last_code = code; i++; }
            else {
                uint32_t run, val;
                if (code == 16) { run = lsb_read(b, 2) + 3; val = last_code; }
                else if (code == 17) { run = lsb_read(b, 3) + 3; val = 0; }
                else if (code == 18) { run = lsb_read(b, 7) + 11; val = 0; }
                else return MS_EDECRUNCH;
                if (b.err) return b.err;
                if (i + run > total) return MS_EDECRUNCH;      /* INF_ERR_BITOVERRUN */
                while (run--) { lens[i * 32] = (uint8_t) val; i++; }
            }
        }
        /* :139-146: distance lengths follow the literal lengths; both are zero-extended */
        uint8_t *l = lens;
        if (ms_canon_build<0, NT>([&](int s) { return (uint32_t) (s < (int) lit_codes ? l[s * 32] : 0); }, 288, 9, lbo, cnt, la.sorted, lhead, HEADN,
                                  (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
#pragma unroll
        for (int j = 0; j < 15; j++) llim[j] = lv[j];
        if (ms_canon_build<0, NT>([&](int s) { return (uint32_t) (s < (int) dist_codes ? l[(lit_codes + s) * 32] : 0); }, 32, 6, dbo, cnt, da.sorted, dhead, 32,
                                  (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
#pragma unroll
        for (int j = 0; j < 15; j++) dlim[j] = lv[j];
        return 0;
    }

    MS_M void fail(int err) { status = err; done = 1; phase = PH_IDLE; }

    /* mszipd.c:159-241: one deflate block header.  Stored blocks are copied right here. */
    MS_M void block_header() {
        lsb_refill(b);
        last_block = lsb_read(b, 1);
        uint32_t type = lsb_read(b, 2);
        if (b.err) { fail(b.err); return; }
        if (type == 0) {
            /* stored block :165-207 */
            lsb_align_byte(b);
            lsb_refill(b); uint32_t len = lsb_read(b, 16);
            lsb_refill(b); uint32_t clen = lsb_read(b, 16);
            if (b.err) { fail(b.err); return; }
            if (len != (~clen & 0xFFFFu)) { fail(MS_EDECRUNCH); return; }
            {   /* bulk copy when the block's bytes all lie inside the input and the frame (the common case) */
                int32_t bp = lsb_bytepos(b);
                if (len && q + len <= MS_FRAME && bp + (int32_t) len <= b.in_len) {
                    emit_raw(em, q, b.in, bp, len);
                    q += len; lsb_seek_byte(b, bp + (int32_t) len); len = 0;
                }
            }
#pragma unroll 1
            for (uint32_t k = 0; k < len; k++) {
                lsb_refill(b);
                uint32_t v = lsb_read(b, 8);
                if (b.err) { fail(b.err); return; }
                emit_literal_checked(em, q, v);
                if (++q >= 2 * MS_FRAME) { fail(MS_EDECRUNCH); return; }   /* second FLUSH_IF_NEEDED: bytes_output > 32 KiB (:323-333) */
            }
            phase = last_block ? PH_END : PH_BLOCK;
            return;
        }
        if (type == 3) { fail(MS_EDECRUNCH); return; }
        int e = 0;
        if (type == 1) {
            /* fixed codes :212-220 */
            uint32_t lv[16];
            if (ms_canon_build<0, NT>([](int s) { return (uint32_t) (s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8))); }, 288, 9, lbo, cnt, la.sorted, lhead, HEADN,
                                      (uint16_t *) nullptr, lv)) e = MS_EDECRUNCH;
            else {
#pragma unroll
                for (int j = 0; j < 15; j++) llim[j] = lv[j];
                if (ms_canon_build<0, NT>([](int) { return 5u; }, 32, 6, dbo, cnt, da.sorted, dhead, 32, (uint16_t *) nullptr, lv)) e = MS_EDECRUNCH;
#pragma unroll
                for (int j = 0; j < 15; j++) dlim[j] = lv[j];
            }
        }
        else e = read_lens();
        if (e) { fail(e); return; }
        phase = PH_DECODE;
    }

    /* :405-413 align to a byte, skip to the next 'C','K'; then the frame's emit state */
    MS_M void frame_start() {
        int state = 0;
        lsb_align_byte(b);
        if (KWAJ && (u->flags & MSGPU_FLAG_MSZIP_KWAJ)) {
            /* :471-481: a 16-bit block length (0 ends the stream; otherwise its value is not used), then 'C', 'K' right away */
            lsb_refill(b);
            uint32_t block_len = lsb_read(b, 8); block_len |= lsb_read(b, 8) << 8;
            if (b.err) { fail(b.err); return; }
            if (block_len == 0) { done = 1; phase = PH_IDLE; return; }
            lsb_refill(b);
            uint32_t c = lsb_read(b, 8);
            if (b.err) { fail(b.err); return; }
            if (c != 'C') { fail(MSGPU_ERR_DATAFORMAT); return; }
            c = lsb_read(b, 8);
            if (b.err) { fail(b.err); return; }
            if (c != 'K') { fail(MSGPU_ERR_DATAFORMAT); return; }
            state = 2;
        }
        else do {
            lsb_refill(b);
            uint32_t c = lsb_read(b, 8);
            if (b.err) { fail(b.err); return; }
            if (c == 'C') state = 1; else if (state == 1 && c == 'K') state = 2; else state = 0;
        } while (state != 2);
        emit_begin(em, recs + (size_t) f * MS_MAXREC, uout + produced, ms_min(MS_FRAME, u->out_len - produced));
        q = 0;
        phase = PH_BLOCK;
    }

    MS_M void frame_end() {
        /* a block that grew past 32 KiB keeps being decoded by the reference (so a read error can still win)
         * and only fails at its next window flush (:308-311, :323-333) */
        if (q > MS_FRAME) { fail(MS_EDECRUNCH); return; }
        const bool kwaj = KWAJ && (u->flags & MSGPU_FLAG_MSZIP_KWAJ);
        if (kwaj && q > u->out_len - produced) { fail(MSGPU_ERR_CAPACITY); return; }        /* out_len is the capacity of the output area */
        uint32_t n = ms_min(u->out_len - produced, q);
        emit_end(em, q);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = n; fi.g0 = produced;
        fi.valid = (frame && hist_snapshot(recs + (size_t) f * MS_MAXREC)) ? 2u : 1u;                    /* 2: resolved by k_p2_ring */
        if (u->flags & (MSGPU_FLAG_CHAIN_FIRST | MSGPU_FLAG_CHAIN_NEXT)) {
            /* one block of a chain (include/msgpu.h): it stands for a stretch of ONE stream only if it is exactly one CK block
             * that ends with its input and fills its output; anything else is for the caller to decode as one stream */
            const int64_t used = (ms_bitpos(b) + 7) >> 3;
            if (q != u->out_len || frame != 0 || used != (int64_t) b.in_len) { fail(MSGPU_ERR_CHAIN); return; }
            fi.valid = 3u;                                                                               /* 3: resolved by k_p2_chain, in chain order */
        }
        finfo[f] = fi;
        const uint32_t g0 = produced;
        produced += n; frame++; f++;
        if (produced >= u->out_len && !kwaj) { done = 1; phase = PH_IDLE; }               /* (a KWAJ stream ends at its zero length only) */
        else if (hist_push(q, g0)) phase = (f < max_frames) ? PH_FRAME : PH_IDLE;
    }

    /* the rare, divergent work: run until the lane is decoding symbols or has nothing left to do */
    MS_M void service() {
#pragma unroll 1
        while (phase >= PH_FRAME) {
            if (phase == PH_FRAME) frame_start();
            else if (phase == PH_BLOCK) block_header();
            else frame_end();
        }
    }

    template <bool careful> MS_M uint32_t litlen_sym() {
        if (careful) lsb_check(b, 16);
        uint32_t v = v16();
        int len = ms_canon_len(llim, v);
        uint32_t idx = ms_canon_index<NT>(lbo, v, len);
        lsb_drop(b, len);
        return idx < (uint32_t) HEADN ? (uint32_t) lhead[idx * NT] : (uint32_t) la.sorted[idx * MS_WARP];
    }
    template <bool careful> MS_M uint32_t dist_sym() {
        if (careful) lsb_check(b, 16);
        uint32_t v = v16();
        int len = ms_canon_len(dlim, v);
        uint32_t idx = ms_canon_index<NT>(dbo, v, len);
        lsb_drop(b, len);
        return dhead[(idx & 31u) * NT];
    }
    template <bool careful> MS_M uint32_t extra_bits(int n) {
        if (careful) return lsb_read(b, n);
        uint32_t v = lsb_peek(b, n); lsb_drop(b, n); return v;
    }

    /* the hot step (mszipd.c:243-300): one literal, or one match (length + distance), or the end-of-block code.
     * `careful` = the unit's input ends within the next 24 bytes: only then can one of this step's reads (two 4-byte refills)
     * trip the reference's end-of-input rule, so only then are the exact checks compiled in. */
    MS_M void step() { if (MS_UNLIKELY(b.ipos + 24 > b.in_len)) step_t<true>(); else step_t<false>(); }
    template <bool careful> MS_M void step_t() {
        lsb_refill(b);
        uint32_t sym = litlen_sym<careful>();
        if (sym < 256) { emit_literal_checked(em, q, sym); q++; }
        else if (sym == 256) phase = last_block ? PH_END : PH_BLOCK;
        else {
            uint32_t c = sym - 257, eb, length, dist;
            if (c >= 29) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :255 */
            if (c < 8) { eb = 0; length = c + 3; }                             /* lit_lengths / lit_extrabits, :47-62 */
            else if (c == 28) { eb = 0; length = 258; }
            else { eb = (c >> 2) - 1; length = ((4 + (c & 3)) << eb) + 3; }
            if (eb) length += extra_bits<careful>((int) eb);
            if (careful && b.err) { fail(b.err); return; }
            lsb_refill(b);
            uint32_t d = dist_sym<careful>();
            if (d >= 30) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :260 */
            if (d < 4) { eb = 0; dist = d + 1; }                               /* dist_offsets / dist_extrabits, :53-68 */
            else { eb = (d >> 1) - 1; dist = ((2 + (d & 1)) << eb) + 1; }
            if (eb) dist += extra_bits<careful>((int) eb);
            if (q + length <= MS_FRAME) emit_match(em, q, length, dist);
            q += length;
        }
        if (careful && b.err) { fail(b.err); return; }
        if (MS_UNLIKELY(q >= 2 * MS_FRAME)) fail(MS_EDECRUNCH);
    }

    /* load the unit's state for this launch */
    MS_M void begin(const msgpu_unit *unit, const uint8_t *in_base, const MsUnitState &st, MsRec *r, uint8_t *l, MsFrameInfo *fi, int nframes) {
        u = unit; recs = r; uout = l; finfo = fi; max_frames = nframes; f = 0; q = 0; last_block = 0;
#pragma unroll 1
        for (int k = 0; k < nframes; k++) { MsFrameInfo z; z.nrec = 0; z.size = 0; z.g0 = 0; z.valid = 0; fi[k] = z; }
        if (!st.started) {
            done = 0; status = 0; produced = 0; frame = 0;
            hist_ptr()[2 * P2_HIST_K * 32] = 0;
            ms_bits_init(b, in_base + unit->in_off, unit->in_len);
            if (unit->out_len == 0 && !(KWAJ && (unit->flags & MSGPU_FLAG_MSZIP_KWAJ))) done = 1;
        }
        else {
            done = st.done; status = st.status; produced = st.produced; frame = st.frame;
            ms_bits_restore(b, in_base + unit->in_off, unit->in_len, st.ipos, (int32_t) st.bc, ((uint64_t) st.bb_hi << 32) | st.bb_lo);
        }
        phase = done ? PH_IDLE : PH_FRAME;
    }
    MS_M void end(MsUnitState &st) {
        st.started = 1; st.done = done; st.status = status; st.produced = produced; st.frame = frame;
        st.ipos = b.ipos; st.bc = (uint32_t) b.bc; st.bb_lo = (uint32_t) b.bb; st.bb_hi = (uint32_t) (b.bb >> 32);
    }
};
