/* msgpu_p1_mszip.cuh - P1 entropy stage for MSZIP units: one lane inflates one unit's "CK" blocks
 * (mszipd.c:154-316 inflate, :91-151 zip_read_lens, :377-460 mszipd_decompress) into literal bytes +
 * match records.  Each CK block (<= 32 KiB of output) is one "frame" of the intermediate form.
 *
 * The lane is a state machine (phases in msgpu_core.cuh): service() does the rare, divergent work
 * (CK signature scan, deflate block headers, code-length tables, stored blocks, frame bookkeeping);
 * step() decodes ONE literal/length symbol (plus its distance) and is what the warp executes in
 * lockstep.
 */
#pragma once
#include "msgpu_core.cuh"

/* per-warp auxiliary block in global scratch (interleaved by lane) */
#define ZIP_AUX_LENS      0                              /* u8  [320][32]  literal/length + distance code lengths */
#define ZIP_AUX_LSORT     (ZIP_AUX_LENS + 320 * 32)      /* u16 [288][32] */
#define ZIP_AUX_DSORT     (ZIP_AUX_LSORT + 288 * 32 * 2) /* u16 [32][32]  */
#define ZIP_AUX_BSORT     (ZIP_AUX_DSORT + 32 * 32 * 2)  /* u16 [32][32]  */
#define ZIP_AUX_LIMIT     (ZIP_AUX_BSORT + 32 * 32 * 2)  /* u32 [3][20][32] */
#define ZIP_AUX_OFFS      (ZIP_AUX_LIMIT + 3 * 20 * 32 * 4) /* u16 [3][20][32] */
#define ZIP_AUX_BYTES     (ZIP_AUX_OFFS + 3 * 20 * 32 * 2)

#ifndef ZIP_LITBATCH
#define ZIP_LITBATCH 1
#endif
/* ZIP_LCACHE = literal/length symbols with codes longer than LROOT kept in shared memory */
template <int NT, int LROOT, int DROOT, int ZIP_LCACHE>
struct ZipShared {
    uint16_t llut[(1 << LROOT) * NT];
    uint16_t lsym[ZIP_LCACHE * NT];       /* the first ZIP_LCACHE long-code literal/length symbols in canonical order */
    uint16_t dlut[(1 << DROOT) * NT];     /* also hosts the 7-bit code-length-code LUT while lengths are read (DROOT >= 7) */
    uint16_t cnt[17 * NT];
};

template <int NT, int LROOT, int DROOT, int ZIP_LCACHE>
struct ZipLane {
    MsBits b;
    uint16_t *llut, *lsym, *dlut, *cnt;   /* this lane's column of the shared tables */
    uint8_t *lens;                        /* aux, stride 32 */
    MsHuffAux la, da, ba;
    MsHuffLong<LROOT> ll;
    MsHuffLong<DROOT> dl;
    /* unit / launch context */
    const msgpu_unit *u; MsRec *recs; uint8_t *uout; MsFrameInfo *finfo;   /* uout = the unit's output buffer */
    MsEmit em;
    uint32_t phase, q, last_block, produced, frame, done; int32_t status;
    int f, max_frames;

    MS_M void bind(ZipShared<NT, LROOT, DROOT, ZIP_LCACHE> *sh, int tid, uint8_t *aux_warp, int lane) {
        llut = sh->llut + tid; lsym = sh->lsym + tid; dlut = sh->dlut + tid; cnt = sh->cnt + tid;
        lens = aux_warp + ZIP_AUX_LENS + lane;
        la.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_LSORT) + lane;
        da.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_DSORT) + lane;
        ba.sorted = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_BSORT) + lane;
        uint32_t *lim = reinterpret_cast<uint32_t *>(aux_warp + ZIP_AUX_LIMIT) + lane;
        uint16_t *off = reinterpret_cast<uint16_t *>(aux_warp + ZIP_AUX_OFFS) + lane;
        la.limit = lim; da.limit = lim + 20 * 32; ba.limit = lim + 40 * 32;
        la.offs = off; da.offs = off + 20 * 32; ba.offs = off + 40 * 32;
    }

    /* READ_HUFFSYM (readhuff.h:39-46) on an LSB-first stream; caller refilled (>= 32 bits) */
    template <int ROOT>
    MS_M uint32_t huffsym(const uint16_t *lut, const MsHuffAux &aux, const MsHuffLong<ROOT> &lg) {
        lsb_check(b, 16);
        uint32_t e = lut[lsb_peek(b, ROOT) * NT];
        int len = (int) (e & 15); uint32_t sym = e >> 4;
        if (len == 0) sym = lg.decode(MS_BREV32((uint32_t) b.bb) >> 16, aux, &len);
        lsb_drop(b, len);
        return sym;
    }

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
        int blmax;
        if (ms_huff_build<7, true, NT>([&](int s) { return (uint32_t) (bl >> (3 * s)) & 7u; }, 19, 7, dlut, ba, cnt, NT, &blmax)) return MS_EDECRUNCH;
        uint32_t total = lit_codes + dist_codes, last_code = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < total;) {
            lsb_refill(b);
            lsb_check(b, 7);                                   /* :117 ENSURE_BITS(7) */
            uint32_t e = dlut[lsb_peek(b, 7) * NT];
            uint32_t code = e >> 4; lsb_drop(b, (int) (e & 15));
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
        uint8_t *l = lens; int lmax, dmax;
        if (ms_huff_build<LROOT, true, NT>([&](int s) { return (uint32_t) (s < (int) lit_codes ? l[s * 32] : 0); }, 288, 9, llut, la, cnt, NT, &lmax, lsym, ZIP_LCACHE)) return MS_EDECRUNCH;
        if (ms_huff_build<DROOT, true, NT>([&](int s) { return (uint32_t) (s < (int) dist_codes ? l[(lit_codes + s) * 32] : 0); }, 32, 6, dlut, da, cnt, NT, &dmax)) return MS_EDECRUNCH;
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
            int lmax, dmax;
            if (ms_huff_build<LROOT, true, NT>([](int s) { return (uint32_t) (s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8))); }, 288, 9, llut, la, cnt, NT, &lmax, lsym, ZIP_LCACHE)) e = MS_EDECRUNCH;
            else if (ms_huff_build<DROOT, true, NT>([](int) { return 5u; }, 32, 6, dlut, da, cnt, NT, &dmax)) e = MS_EDECRUNCH;
        }
        else e = read_lens();
        if (e) { fail(e); return; }
        ll.load(la); dl.load(da);
        phase = PH_DECODE;
    }

    /* :405-413 align to a byte, skip to the next 'C','K'; then the frame's emit state */
    MS_M void frame_start() {
        int state = 0;
        lsb_align_byte(b);
        do {
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
        uint32_t n = ms_min(u->out_len - produced, q);
        emit_end(em, q);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = n; fi.g0 = produced; fi.valid = 1;
        finfo[f] = fi;
        produced += n; frame++; f++;
        if (produced >= u->out_len) { done = 1; phase = PH_IDLE; }
        else phase = (f < max_frames) ? PH_FRAME : PH_IDLE;
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

    MS_M uint32_t litlen_sym() {
        lsb_check(b, 16);
        uint32_t e = llut[lsb_peek(b, LROOT) * NT];
        int len = (int) (e & 15); uint32_t sym = e >> 4;
        if (len == 0) sym = ll.template decode_cached<NT>(MS_BREV32((uint32_t) b.bb) >> 16, la, lsym, ZIP_LCACHE, &len);
        lsb_drop(b, len);
        return sym;
    }

    /* the hot step (mszipd.c:243-300): up to ZIP_LITBATCH literals, or one match (length + distance), or the
     * end-of-block code */
    MS_M void step() {
        uint32_t sym;
#pragma unroll 1
        for (int rep = 0;;) {
            lsb_refill(b);
            sym = litlen_sym();
            if (sym >= 256) break;
            emit_literal_checked(em, q, sym);
            q++;
            if (MS_UNLIKELY(b.err)) { fail(b.err); return; }
            if (MS_UNLIKELY(q >= 2 * MS_FRAME)) { fail(MS_EDECRUNCH); return; }
            if (++rep == ZIP_LITBATCH) return;
        }
        if (sym == 256) phase = last_block ? PH_END : PH_BLOCK;

        else {
            uint32_t c = sym - 257, eb, length, dist;
            if (c >= 29) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :255 */
            if (c < 8) { eb = 0; length = c + 3; }                             /* lit_lengths / lit_extrabits, :47-62 */
            else if (c == 28) { eb = 0; length = 258; }
            else { eb = (c >> 2) - 1; length = ((4 + (c & 3)) << eb) + 3; }
            if (eb) length += lsb_read(b, (int) eb);
            if (b.err) { fail(b.err); return; }
            lsb_refill(b);
            uint32_t d = huffsym<DROOT>(dlut, da, dl);
            if (d >= 30) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :260 */
            if (d < 4) { eb = 0; dist = d + 1; }                               /* dist_offsets / dist_extrabits, :53-68 */
            else { eb = (d >> 1) - 1; dist = ((2 + (d & 1)) << eb) + 1; }
            if (eb) dist += lsb_read(b, (int) eb);
            if (q + length <= MS_FRAME) emit_match(em, q, length, dist);
            q += length;
        }
        if (MS_UNLIKELY(b.err)) { fail(b.err); return; }
        if (MS_UNLIKELY(q >= 2 * MS_FRAME)) fail(MS_EDECRUNCH);
    }

    /* load the unit's state for this launch */
    MS_M void begin(const msgpu_unit *unit, const uint8_t *in_base, const MsUnitState &st, MsRec *r, uint8_t *l, MsFrameInfo *fi, int nframes) {
        u = unit; recs = r; uout = l; finfo = fi; max_frames = nframes; f = 0; q = 0; last_block = 0;
#pragma unroll 1
        for (int k = 0; k < nframes; k++) { MsFrameInfo z; z.nrec = 0; z.size = 0; z.g0 = 0; z.valid = 0; fi[k] = z; }
        if (!st.started) {
            done = 0; status = 0; produced = 0; frame = 0;
            ms_bits_init(b, in_base + unit->in_off, unit->in_len);
            if (unit->out_len == 0) done = 1;
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
