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

/* SPECIAL = the instantiation that also understands the two special kinds of MSZIP unit: MSGPU_FLAG_MSZIP_KWAJ
 * (mszipd_decompress_kwaj, mszipd.c:462-495) and MSGPU_FLAG_MSZIP_REPAIR (mszipd_init(repair_mode = 1), mszipd.c:420-433) */
template <int NT, int HEADN, bool SPECIAL = false>
struct ZipLaneC {
    /* BYTEWISE (the plain instantiation): literal bytes go to the output one by one instead of through the word gatherer - ~30
     * instructions less per literal step for the lanes that have one; measured on the B200 (profiles/r2_variants.txt, shape 18):
     * P1 8.65 -> 7.89 ms on the 32 768-unit text batch.  Every literal path of the lane must then do the same, because a gathered
     * word is stored whole; the KWAJ / repair instantiation keeps the gatherer (its overflow frames write to a plane). */
    static constexpr bool BYTEWISE = !SPECIAL;
    /* Repair mode.  A block the reference gives up is zero-filled to 32 KiB and decoding goes on - with the bit state of its last
     * STORE_BITS (mszipd.c:149 / :223 / :419), which is stale in two ways (see oracle/port/mspack_port.c zip_repair_restart, pinned
     * against the reference): the buffered bits are those of the STORE, and the byte pointer is the STORE's only if read_input has
     * not refilled the input buffer since - a refill resets it to the buffer's start.  So the lane tracks what the reference has
     * FETCHED (fx: every ENSURE_BITS / READ_IF_NEEDED of the reference is a rd() / ck() here) and remembers position, fetch extent
     * and buffered bits at every STORE. */
    int32_t fx, store_fx; int64_t store_p; uint32_t store_val, in_block;
    /* A block that grows past 32 KiB is not an error yet for the reference: its window position wraps to 0 and the block goes on
     * OVERWRITING the start of the window until the next flush fails (mszipd.c:38-45, :323-333).  In repair mode that window image
     * is what gets written, so the overflow is kept as a second frame at the SAME output position (qbase = 32768, records at
     * q - qbase): to the resolve stage it is an ordinary ring frame whose one history entry is the block's first 32 KiB. */
    uint32_t qbase;
    MS_M bool repairing() const { return SPECIAL && (u->flags & MSGPU_FLAG_MSZIP_REPAIR); }
    MS_M void trk(int n) { if (SPECIAL) { const int32_t need = (int32_t) ((ms_bitpos(b) + n + 7) >> 3); if (need > fx) fx = need; } }
    MS_M uint32_t rd(int n) { trk(n); return lsb_read(b, n); }
    MS_M void ck(int n) { trk(n); lsb_check(b, n); }
    MS_M void store_bits() {           /* STORE_BITS: position, fetch extent, the bits fetched but not consumed */
        if (SPECIAL) { lsb_refill(b); store_p = ms_bitpos(b); store_fx = fx; const int64_t ns = (int64_t) store_fx * 8 - store_p; store_val = (uint32_t) b.bb & (ns > 0 ? (ns >= 32 ? 0xFFFFFFFFu : ((1u << ns) - 1u)) : 0u); }
    }
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
        uint32_t lit_codes = rd(5) + 257, dist_codes = rd(5) + 1, bl_codes = rd(4) + 4;
        if (b.err) return b.err;
        if (lit_codes > 288 || dist_codes > 32) return MS_EDECRUNCH;
        /* 19 code-length-code lengths, 3 bits each, packed into a 64-bit register */
        uint64_t bl = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < bl_codes; i++) { lsb_refill(b); bl |= (uint64_t) rd(3) << (3 * order[i]); }
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
            ck(7);                                   /* :117 ENSURE_BITS(7) */
            uint32_t v = v16();
            int cl = ms_canon_len_smem<NT>(blim, v);
            uint32_t code = ba.sorted[ms_canon_index<NT>(dbo, v, cl) * MS_WARP]; lsb_drop(b, cl);
            if (b.err) return b.err;
            if (code < 16) { lens[i * 32] = (uint8_t) code; last_code = code; i++; }
            else {
                uint32_t run, val;
                if (code == 16) { run = rd(2) + 3; val = last_code; }
                else if (code == 17) { run = rd(3) + 3; val = 0; }
                else if (code == 18) { run = rd(7) + 11; val = 0; }
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

    MS_M void fail(int err) {
        if (SPECIAL && in_block && repairing()) { repair_block(err); return; }
        status = err; done = 1; phase = PH_IDLE;
    }

    /* the block has filled its 32 KiB: close that frame, open the overflow frame in the next slot (same output position) */
    MS_M void start_overflow() {
        /* the overflow reads the block's first 32 KiB back; if the unit's out_len cuts that frame short (only its last frame can
         * be) those bytes have no place to live: refuse loudly rather than read outside the unit (stated deviation) */
        if (em.limit < MS_FRAME) { in_block = 0; status = MS_EDECRUNCH; done = 1; phase = PH_IDLE; return; }
        emit_end(em, MS_FRAME);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = ms_min(u->out_len - produced, MS_FRAME); fi.g0 = produced;
        fi.valid = (frame && hist_snapshot(recs + (size_t) f * MS_MAXREC)) ? 2u : 1u;
        finfo[f] = fi;
        /* the overflow frame's literals go to its plane, not to the output (msgpu_p2.cuh P2_PLANE_REC) */
        emit_begin(em, recs + (size_t) (f + 1) * MS_MAXREC, reinterpret_cast<uint8_t *>(recs + (size_t) (f + 1) * MS_MAXREC + P2_PLANE_REC),
                   ms_min(MS_FRAME, u->out_len - produced));
        qbase = MS_FRAME;
    }

    /* mszipd.c:420-433 + :404: give the block up - keep what it decoded, zero-fill the rest of its 32 KiB, go on behind it */
    MS_M void repair_block(int err) {
        in_block = 0;
        if (qbase) {
            /* the overflow frame: q - 32768 bytes (at most 32 KiB) on top of the block's first 32 KiB, which is its whole history */
            const uint32_t nov = ms_min(q - MS_FRAME, MS_FRAME);
            emit_end(em, nov);
            uint32_t *sn = reinterpret_cast<uint32_t *>(recs + (size_t) (f + 1) * MS_MAXREC + P2_HIST_REC);
            sn[0] = 1u; sn[1] = MS_FRAME; sn[2] = produced;
            MsFrameInfo fi; fi.nrec = em.nrec; fi.size = ms_min(em.limit, nov); fi.g0 = produced; fi.valid = 4u;      /* 4: ring frame with a literal plane */
            finfo[f + 1] = fi;
            const uint32_t g0 = produced;
            produced += ms_min(u->out_len - produced, MS_FRAME); frame++; f += 2; qbase = 0;
            if (produced >= u->out_len) { done = 1; phase = PH_IDLE; }
            else if (!hist_push(MS_FRAME, g0)) return;
        }
        else {
#pragma unroll 1
            for (uint32_t k = q < MS_FRAME ? q : MS_FRAME; k < em.limit; k++) emit_literal(em, k, 0);
            q = MS_FRAME;
            if (!finish_frame()) return;
        }
        if (err == MS_EREAD) { status = MS_EREAD; done = 1; phase = PH_IDLE; return; }      /* :448 read errors stay fatal (after the frame went out) */
        if (done) return;
        /* where the reference goes on: the whole bytes left in the bit buffer of the last STORE_BITS, then its byte pointer - the
         * STORE's, or the start of the input buffer if that has been refilled since (chunks of `size` bytes of the stream; the two
         * zero bytes invented at the end of the input are a chunk of their own) */
        uint32_t size = MSGPU_UNIT_REF_BYTES(u) ? MSGPU_UNIT_REF_BYTES(u) : 4096u; size = (size + 1u) & ~1u;
        const int32_t cnow = fx - 1 >= b.in_len ? b.in_len : (int32_t) ((uint32_t) (fx - 1) / size * size);
        const int32_t csto = store_fx - 1 >= b.in_len ? b.in_len : (int32_t) ((uint32_t) (store_fx - 1) / size * size);
        const int32_t r = (fx > 0 && (store_fx == 0 || cnow != csto)) ? cnow : store_fx;
        const int64_t ns = (int64_t) store_fx * 8 - store_p;
        const uint32_t nb = ns > 0 ? (uint32_t) (ns >> 3) : 0u;                               /* :405 drops the ns & 7 odd bits */
        const uint64_t stale = (uint64_t) (store_val >> (ns & 7)) & ((1ull << (8 * nb)) - 1ull);
        b.err = 0;
        lsb_seek_byte(b, r);
        b.bb = (b.bb << (8 * nb)) | stale; b.bc += (int32_t) (8 * nb);
        fx = r;
        phase = (f + 2 <= max_frames) ? PH_FRAME : PH_IDLE;           /* (a repair unit's block may need two frame slots) */
    }

    /* mszipd.c:159-241: one deflate block header.  Stored blocks are copied right here. */
    MS_M void block_header() {
        lsb_refill(b);
        last_block = rd(1);
        uint32_t type = rd(2);
        if (b.err) { fail(b.err); return; }
        if (type == 0) {
            /* stored block :165-207 */
            lsb_align_byte(b);
            lsb_refill(b); uint32_t len = rd(16);
            lsb_refill(b); uint32_t clen = rd(16);
            if (b.err) { fail(b.err); return; }
            if (len != (~clen & 0xFFFFu)) { fail(MS_EDECRUNCH); return; }
            {   /* bulk copy when the block's bytes all lie inside the input and the frame (the common case) */
                int32_t bp = lsb_bytepos(b);
                if (len && q + len <= MS_FRAME && bp + (int32_t) len <= b.in_len) {
                    if constexpr (BYTEWISE) { for (uint32_t k = 0; k < len; k++) lit_bytewise(q + k, b.in[bp + (int32_t) k]); }
                    else
                    emit_raw(em, q, b.in, bp, len);
                    q += len; lsb_seek_byte(b, bp + (int32_t) len);
                    if (SPECIAL && bp + (int32_t) len > fx) fx = bp + (int32_t) len;
                    len = 0;
                }
            }
#pragma unroll 1
            for (uint32_t k = 0; k < len; k++) {
                lsb_refill(b);
                uint32_t v = rd(8);
                if (b.err) { fail(b.err); return; }
                if (SPECIAL && q >= MS_FRAME && !qbase && repairing()) { start_overflow(); if (done) return; }
                if constexpr (BYTEWISE) lit_bytewise(q, v); else
                emit_literal_checked(em, q - (SPECIAL ? qbase : 0u), v);
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
        else { store_bits(); e = read_lens(); if (!e) store_bits(); }                     /* mszipd.c:223, :149 */
        if (e) { fail(e); return; }
        phase = PH_DECODE;
    }

    /* :405-413 align to a byte, skip to the next 'C','K'; then the frame's emit state */
    MS_M void frame_start() {
        int state = 0;
        lsb_align_byte(b);
        if (SPECIAL && (u->flags & MSGPU_FLAG_MSZIP_KWAJ)) {
            /* :471-481: a 16-bit block length (0 ends the stream; otherwise its value is not used), then 'C', 'K' right away */
            lsb_refill(b);
            uint32_t block_len = rd(8); block_len |= rd(8) << 8;
            if (b.err) { fail(b.err); return; }
            if (block_len == 0) { done = 1; phase = PH_IDLE; return; }
            lsb_refill(b);
            uint32_t c = rd(8);
            if (b.err) { fail(b.err); return; }
            if (c != 'C') { fail(MSGPU_ERR_DATAFORMAT); return; }
            c = rd(8);
            if (b.err) { fail(b.err); return; }
            if (c != 'K') { fail(MSGPU_ERR_DATAFORMAT); return; }
            state = 2;
        }
        else do {
            lsb_refill(b);
            uint32_t c = rd(8);
            if (b.err) { fail(b.err); return; }
            if (c == 'C') state = 1; else if (state == 1 && c == 'K') state = 2; else state = 0;
        } while (state != 2);
        emit_begin(em, recs + (size_t) f * MS_MAXREC, uout + produced, ms_min(MS_FRAME, u->out_len - produced));
        q = 0;
        if (SPECIAL) { store_bits(); in_block = 1; }                                             /* mszipd.c:419 */
        phase = PH_BLOCK;
    }

    MS_M void frame_end() {
        /* a block that grew past 32 KiB keeps being decoded by the reference (so a read error can still win)
         * and only fails at its next window flush (:308-311, :323-333) */
        if (q > MS_FRAME) { fail(MS_EDECRUNCH); return; }
        if (SPECIAL) in_block = 0;
        (void) finish_frame();
    }
    /* the block's q bytes become a frame of the intermediate form; false if the unit failed on the way */
    MS_M bool finish_frame() {
        const bool kwaj = SPECIAL && (u->flags & MSGPU_FLAG_MSZIP_KWAJ);
        if (kwaj && q > u->out_len - produced) { status = MSGPU_ERR_CAPACITY; done = 1; phase = PH_IDLE; return false; }   /* out_len is the capacity of the output area */
        uint32_t n = ms_min(u->out_len - produced, q);
        emit_end(em, q);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = n; fi.g0 = produced;
        fi.valid = (frame && hist_snapshot(recs + (size_t) f * MS_MAXREC)) ? 2u : 1u;                    /* 2: resolved by k_p2_ring */
        if (u->flags & (MSGPU_FLAG_CHAIN_FIRST | MSGPU_FLAG_CHAIN_NEXT)) {
            /* one block of a chain (include/msgpu.h): it stands for a stretch of ONE stream only if it is exactly one CK block
             * that ends with its input and fills its output; anything else is for the caller to decode as one stream */
            const int64_t used = (ms_bitpos(b) + 7) >> 3;
            if (q != u->out_len || frame != 0 || used != (int64_t) b.in_len) { status = MSGPU_ERR_CHAIN; done = 1; phase = PH_IDLE; return false; }
            fi.valid = 3u;                                                                               /* 3: resolved by k_p2_chain, in chain order */
        }
        finfo[f] = fi;
        const uint32_t g0 = produced;
        produced += n; frame++; f++;
        if (produced >= u->out_len && !kwaj) { done = 1; phase = PH_IDLE; }               /* (a KWAJ stream ends at its zero length only) */
        else if (hist_push(q, g0)) phase = (f + ((SPECIAL && repairing()) ? 2 : 1) <= max_frames) ? PH_FRAME : PH_IDLE;
        else return false;
        return true;
    }

    MS_M void post_step() { }
    /* the rare, divergent work: run until the lane is decoding symbols or has nothing left to do */
    MS_M void service() {
#pragma unroll 1
        while (phase >= PH_FRAME && phase != PH_PARK) {
            if (phase == PH_FRAME) phase = PH_PARK;                   /* wait for the warp (msgpu_core.cuh PH_PARK) */
            else if (phase == (PH_FRAME | 0x100u)) frame_start();
            else if (phase == PH_BLOCK) block_header();
            else frame_end();
        }
    }

    template <bool careful> MS_M uint32_t litlen_sym() {
        if (careful) ck(16);
        uint32_t v = v16();
        int len = ms_canon_len(llim, v);
        uint32_t idx = ms_canon_index<NT>(lbo, v, len);
        lsb_drop(b, len);
        return idx < (uint32_t) HEADN ? (uint32_t) lhead[idx * NT] : (uint32_t) la.sorted[idx * MS_WARP];
    }
    template <bool careful> MS_M uint32_t dist_sym() {
        if (careful) ck(16);
        uint32_t v = v16();
        int len = ms_canon_len(dlim, v);
        uint32_t idx = ms_canon_index<NT>(dbo, v, len);
        lsb_drop(b, len);
        return dhead[(idx & 31u) * NT];
    }
    template <bool careful> MS_M uint32_t extra_bits(int n) {
        if (careful) return rd(n);
        uint32_t v = lsb_peek(b, n); lsb_drop(b, n); return v;
    }

    /* the hot step (mszipd.c:243-300): one literal, or one match (length + distance), or the end-of-block code.
     * `careful` = the unit's input ends within the next 24 bytes: only then can one of this step's reads (two 4-byte refills)
     * trip the reference's end-of-input rule, so only then are the exact checks compiled in. */
    MS_M void step() {
        if (MS_UNLIKELY(b.ipos + 24 > b.in_len) || (SPECIAL && repairing())) step_t<true>(); else step_t<false>();
    }
    template <bool careful> MS_M void refill() { lsb_refill(b); }
    MS_M void lit_bytewise(uint32_t qq, uint32_t v) { if (qq < em.limit) em.out[qq] = (uint8_t) v; }
    template <bool careful> MS_M void step_t() {
        refill<careful>();
        uint32_t sym = litlen_sym<careful>();
        if (sym < 256) {
            if (SPECIAL && q >= MS_FRAME && !qbase && repairing()) { start_overflow(); if (done) return; }
            if constexpr (BYTEWISE) lit_bytewise(q, sym); else
            emit_literal_checked(em, q - (SPECIAL ? qbase : 0u), sym); q++;
        }
        else if (sym == 256) phase = last_block ? PH_END : PH_BLOCK;
        else {
            uint32_t c = sym - 257, eb, length, dist;
            if (c >= 29) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :255 */
            if (c < 8) { eb = 0; length = c + 3; }                             /* lit_lengths / lit_extrabits, :47-62 */
            else if (c == 28) { eb = 0; length = 258; }
            else { eb = (c >> 2) - 1; length = ((4 + (c & 3)) << eb) + 3; }
            if (eb) length += extra_bits<careful>((int) eb);
            if (careful && b.err) { fail(b.err); return; }
            refill<careful>();
            uint32_t d = dist_sym<careful>();
            if (d >= 30) { fail(b.err ? b.err : MS_EDECRUNCH); return; }     /* :260 */
            if (d < 4) { eb = 0; dist = d + 1; }                               /* dist_offsets / dist_extrabits, :53-68 */
            else { eb = (d >> 1) - 1; dist = ((2 + (d & 1)) << eb) + 1; }
            if (eb) dist += extra_bits<careful>((int) eb);
            if (q + length <= MS_FRAME) emit_match(em, q, length, dist);
            else if (SPECIAL && repairing()) {                        /* the match crosses (or lies behind) the 32 KiB mark: see qbase */
                uint32_t first = q < MS_FRAME ? MS_FRAME - q : 0u, rest = length - first, at = q + first - MS_FRAME;
                if (first) emit_match(em, q, first, dist);
                if (!qbase) { start_overflow(); if (done) return; }
                if (at + rest > MS_FRAME) rest = at < MS_FRAME ? MS_FRAME - at : 0u;       /* (the second wrap fails the block) */
                if (rest) emit_match(em, at, rest, dist);
            }
            q += length;
        }
        if (careful && b.err) { fail(b.err); return; }
        if (MS_UNLIKELY(q >= 2 * MS_FRAME)) fail(MS_EDECRUNCH);
    }

    /* load the unit's state for this launch */
    MS_M void begin(const msgpu_unit *unit, const uint8_t *in_base, const MsUnitState &st, MsRec *r, uint8_t *l, MsFrameInfo *fi, int nframes) {
        u = unit; recs = r; uout = l; finfo = fi; max_frames = nframes; f = 0; q = 0; last_block = 0; in_block = 0; store_fx = 0; store_p = 0; store_val = 0; qbase = 0;
#pragma unroll 1
        for (int k = 0; k < nframes; k++) { MsFrameInfo z; z.nrec = 0; z.size = 0; z.g0 = 0; z.valid = 0; fi[k] = z; }
        if (!st.started) {
            done = 0; status = 0; produced = 0; frame = 0; fx = 0;
            hist_ptr()[2 * P2_HIST_K * 32] = 0;
            ms_bits_init(b, in_base + unit->in_off, unit->in_len);
            if (unit->out_len == 0 && !(SPECIAL && (unit->flags & MSGPU_FLAG_MSZIP_KWAJ))) done = 1;
        }
        else {
            done = st.done; status = st.status; produced = st.produced; frame = st.frame; fx = (int32_t) st.R0;      /* (an LZX field, free here) */
            ms_bits_restore(b, in_base + unit->in_off, unit->in_len, st.ipos, (int32_t) st.bc, ((uint64_t) st.bb_hi << 32) | st.bb_lo);
        }
        phase = done ? PH_IDLE : PH_FRAME;
    }
    MS_M void end(MsUnitState &st) {
        st.started = 1; st.done = done; st.status = status; st.produced = produced; st.frame = frame;
        st.ipos = b.ipos; st.bc = (uint32_t) b.bc; st.bb_lo = (uint32_t) b.bb; st.bb_hi = (uint32_t) (b.bb >> 32);
        if (SPECIAL) st.R0 = (uint32_t) fx;
    }
};
