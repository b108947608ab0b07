/* msgpu_p1_lzx.cuh - P1 entropy stage for LZX units: one lane walks one unit's bitstream
 * (lzxd.c:388-771 lzxd_decompress, :138-183 lzxd_read_lens, :257-270 lzxd_reset_state), stores the literal bytes
 * at their output positions and emits one match record per match, per 32 KiB frame.  Window-relative checks are
 * restated for a linear output buffer: the reference's window_posn is (bytes decoded) mod window_size and its
 * lzx->offset is the frame's start (see oracle/port/mspack_port.c for the same restatement on the CPU).
 *
 * Huffman decoding is table-free ("canonical lanes"): no per-lane decode LUT for the main tree, code lengths come
 * from 15 register-resident limits, symbols from a small shared-memory head + global scratch (msgpu_core.cuh
 * "Table-free canonical decoding").  ~0.45 KB of shared memory per lane, so 14 warps per SM; the LUT-based lanes of
 * the first iterations needed ~1 KB per lane (6 warps per SM) and were 2.4x slower.
 */
#pragma once
#include <type_traits>
#include "msgpu_core.cuh"

#define LZX_MAIN_MAX   2576                   /* LZX_MAINTREE_MAXSYMBOLS, lzx.h:38 */
#define LZX_MAIN_ALLOC (LZX_MAIN_MAX + 64)    /* + LZX_LENTABLE_SAFETY, lzx.h:44 */
#define LZX_LEN_SYMS   250                    /* LZX_LENGTH_MAXSYMBOLS */
#define LZX_LEN_ALLOC  320
#define LZX_MSORT_N    2592                   /* coded main symbols <= 2576 (window_bits 25, LZX DELTA) */

#define LZX_AUX_MAINLEN  0                                         /* u8  [2640][32] */
#define LZX_AUX_LENLEN   (LZX_AUX_MAINLEN + LZX_MAIN_ALLOC * 32)   /* u8  [320][32]  */
#define LZX_AUX_MSORT    (LZX_AUX_LENLEN + LZX_LEN_ALLOC * 32)     /* u16 [2592][32] */
#define LZX_AUX_LSORT    (LZX_AUX_MSORT + LZX_MSORT_N * 32 * 2)    /* u16 [256][32]  */
#define LZX_AUX_PSORT    (LZX_AUX_LSORT + 256 * 32 * 2)            /* u16 [32][32]   */
#define LZX_AUX_ASORT    (LZX_AUX_PSORT + 32 * 32 * 2)             /* u16 [16][32]   */
#define LZX_AUX_LIMIT    (LZX_AUX_ASORT + 16 * 32 * 2)             /* u32 [4][20][32] */
#define LZX_AUX_OFFS     (LZX_AUX_LIMIT + 4 * 20 * 32 * 4)         /* u16 [4][20][32] */
#define LZX_AUX_BYTES    (LZX_AUX_OFFS + 4 * 20 * 32 * 2)

/* HEADN = main-tree symbols (shortest codes first) kept in shared memory.
 * DELTA = the instantiation that also understands LZX DELTA units (MSGPU_FLAG_LZX_DELTA: window_bits up to 25, a 16-bit
 * chunk size per frame lzxd.c:441-444, match lengths beyond 257 :589-611, reference data in front of the unit :348-382
 * and :622-628); batches without such units run the plain instantiation, whose code is unchanged by all of this */
template <int NT, int HEADN>
struct LzxSharedC {
    uint32_t mbo[17 * NT];                /* main tree: limit[l-1] >> 1 | offs[l] << 16 */
    uint32_t lbo[17 * NT];                /* LENGTH tree; hosts the pretree while code lengths are being read */
    uint32_t abo[17 * NT];                /* aligned-offset tree */
    uint16_t mhead[HEADN * NT];           /* first HEADN main symbols in canonical order */
    uint16_t llim[16 * NT];               /* LENGTH / pretree limits >> 1 */
    uint16_t alim[16 * NT];               /* aligned tree limits >> 1 */
    uint16_t llut[32 * NT];               /* 5-bit LUT of the LENGTH tree */
    uint16_t cnt[17 * NT];
};

/* The packed layout.  Measured on the headline batch the main tree of a 32 KiB text frame has ~250 coded symbols with codes of 6-8
 * bits: a 72-symbol head misses for 29 % of the symbols, i.e. in every step some lane of the warp goes to L2 for its symbol and the
 * whole warp waits.  A main symbol is < 656 (window_bits <= 21): 10 bits, kept as a byte plus two bits (16 per word), so ~1.25 bytes
 * per head entry instead of 2; the aligned-offset tree (8 symbols, codes <= 7 bits) shrinks from 100 bytes to 4 words, and the
 * builder's counters share the LENGTH limits' array (never live together). */
/* 16-bit per-length bases (MsBoK) instead of one word per length: 64 bytes less for the two big trees,
 * spent on a full 256-entry main head and a 40-entry byte head of the LENGTH tree's symbols (the LENGTH symbols the LUT
 * misses were the other L2 round trip the profile showed: ~10 % of the steps had a lane there).  H8LB = 100 + LUT bits. */
#define LZX_LHEAD 40
template <int NT, int HEADN, int LB>
struct LzxSharedQ {
    uint32_t atree[4 * NT];
    uint32_t mhi[(HEADN / 16) * NT];
    uint16_t mbo[17 * NT];                /* main tree, K form */
    uint16_t lbo[17 * NT];                /* LENGTH tree / pretree, K form */
    uint16_t llim[17 * NT];
    uint16_t llut[(1 << LB) * NT];
    uint8_t lhead[LZX_LHEAD * NT];        /* first LENGTH symbols in canonical order */
    static constexpr int ROW = NT;
    uint8_t mlo[HEADN * ROW];
};
template <int NT, int HEADN, int H8LB> struct LzxSharedSel { static_assert(H8LB >= 100, "H8LB = 100 + LUT bits (LzxSharedQ) or 0 (LzxSharedC)"); typedef LzxSharedQ<NT, HEADN, H8LB - 100> type; };
template <int NT, int HEADN> struct LzxSharedSel<NT, HEADN, 0> { typedef LzxSharedC<NT, HEADN> type; };

template <int NT, int HEADN, bool DELTA = false, int H8LB = 0>
struct LzxLaneC {
    typedef typename LzxSharedSel<NT, HEADN, H8LB>::type Shared;
    static constexpr bool H8 = H8LB != 0;
    static constexpr bool QL = H8LB >= 100;                     /* 16-bit bases + LENGTH head (LzxSharedQ) */
    static constexpr int LUTB = H8 ? (QL ? H8LB - 100 : H8LB) : 5;
    typedef typename std::conditional<QL, MsBoK<NT>, MsBo32<NT>>::type Bo;
    MsBits b;
    uint32_t *atree, *mhi; uint8_t *mlo, *lhead;  /* packed layouts only */
    uint32_t is_delta, ref_len;           /* DELTA only: this unit is an LZX DELTA stream; bytes of reference data in front of it */
    Bo mbo, lbo; uint32_t *abo;
    uint16_t *mhead, *llim, *alim, *llut, *cnt;
    uint8_t *main_len, *len_len;
    MsHuffAux ma, la, pa, aa;             /* only .sorted is used (global scratch) */
    uint32_t mlim[15];                    /* main tree limit[1..15], registers */
    uint32_t R0, R1, R2, block_type, block_length, block_remaining, header_read, intel_started, length_empty, aligned_lens;
    int32_t intel_filesize;
    uint32_t window_size, num_offsets, nsyms_eff, bytemode, base;
    int32_t bytepos;                      /* valid in bytemode: next raw byte (relative to b.in) */
    /* unit / launch context */
    const msgpu_unit *u; MsRec *recs; uint8_t *uout; MsFrameInfo *finfo; int32_t *e8info;   /* uout = the unit's output buffer */
    MsEmit em;
    uint32_t phase, q, produced, frame, done, frame_start_pos, frame_size; int32_t status, bytes_todo, this_run;
    int f, max_frames;

    MS_M void bind(Shared *sh, int tid, uint8_t *aux_warp, int lane) {
        mbo.p = sh->mbo + tid; lbo.p = sh->lbo + tid; llim = sh->llim + tid; llut = sh->llut + tid;
        if constexpr (QL) lhead = sh->lhead + tid; else lhead = nullptr;
        if constexpr (H8) { atree = sh->atree + tid; mhi = sh->mhi + tid; mlo = sh->mlo + tid; cnt = llim; abo = nullptr; mhead = nullptr; alim = nullptr; }
        else { abo = sh->abo + tid; mhead = sh->mhead + tid; alim = sh->alim + tid; cnt = sh->cnt + tid; }
        main_len = aux_warp + LZX_AUX_MAINLEN + lane; len_len = aux_warp + LZX_AUX_LENLEN + lane;
        ma.sorted = reinterpret_cast<uint16_t *>(aux_warp + LZX_AUX_MSORT) + lane;
        la.sorted = reinterpret_cast<uint16_t *>(aux_warp + LZX_AUX_LSORT) + lane;
        pa.sorted = reinterpret_cast<uint16_t *>(aux_warp + LZX_AUX_PSORT) + lane;
        aa.sorted = reinterpret_cast<uint16_t *>(aux_warp + LZX_AUX_ASORT) + lane;
        uint32_t *lim = reinterpret_cast<uint32_t *>(aux_warp + LZX_AUX_LIMIT) + lane;
        uint16_t *off = reinterpret_cast<uint16_t *>(aux_warp + LZX_AUX_OFFS) + lane;
        ma.limit = lim; la.limit = lim + 20 * 32; pa.limit = lim + 40 * 32; aa.limit = lim + 60 * 32;
        ma.offs = off; la.offs = off + 20 * 32; pa.offs = off + 40 * 32; aa.offs = off + 60 * 32;
    }

    MS_M void reset_state() {                 /* lzxd.c:257-270 */
        R0 = R1 = R2 = 1; header_read = 0; block_remaining = 0; block_type = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < nsyms_eff; i++) main_len[i * 32] = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < LZX_LEN_SYMS; i++) len_len[i * 32] = 0;
    }

    /* READ_HUFFSYM for a tree whose limits live in shared memory (pretree, LENGTH slow path, aligned tree) */
    template <class BoT>
    MS_M uint32_t sym_smem(const uint16_t *lim16, BoT bo, const uint16_t *sorted, bool careful = true) {
        if (careful) lzx_check(b, 16);
        uint32_t v16 = msb_peek(b, 16);
        int len = ms_canon_len_smem<NT>(lim16, v16);
        uint32_t idx = bo.index(v16, len);
        msb_drop(b, len);
        return sorted[idx * MS_WARP];
    }
    /* READ_HUFFSYM(ALIGNED) from the four-word tree of the packed layout (build_aligned) */
    MS_M uint32_t aligned_sym(bool careful) {
        if (careful) lzx_check(b, 16);
        const uint32_t v7 = msb_peek(b, 7), w0 = atree[0], w1 = atree[NT], w2 = atree[2 * NT], w3 = atree[3 * NT];
        const uint64_t lims = (uint64_t) w0 | ((uint64_t) w1 << 32);          /* limit[l] at bits 8 (l - 1) */
        uint32_t len = 1;
#pragma unroll
        for (int l = 1; l <= 6; l++) len += (v7 >= ((uint32_t) (lims >> (8 * (l - 1))) & 0xFFu)) ? 1u : 0u;
        const uint32_t below = len > 1 ? (uint32_t) (lims >> (8 * (len - 2))) & 0xFFu : 0u;
        const uint32_t idx = ((w2 >> (4 * (len - 1))) & 15u) + ((v7 - below) >> (7 - len));
        msb_drop(b, (int) len);
        return (w3 >> (3 * idx)) & 7u;
    }
    MS_M uint32_t length_sym(bool careful) {  /* LENGTH tree: 5-bit LUT, then the canonical path */
        if (careful) lzx_check(b, 16);
        uint32_t e = llut[msb_peek(b, LUTB) * NT];
        if (e & 15) { msb_drop(b, (int) (e & 15)); return e >> 4; }
        uint32_t v16 = msb_peek(b, 16);
        int len = ms_canon_len_smem<NT>(llim, v16);
        uint32_t idx = lbo.index(v16, len);
        msb_drop(b, len);
        if constexpr (QL) { if (idx < (uint32_t) LZX_LHEAD) return lhead[idx * NT]; }
        return la.sorted[idx * MS_WARP];
    }

    /* raw byte access for uncompressed blocks; READ_IF_NEEDED semantics (readbits.h:182-214) */
    MS_M uint32_t raw_byte() {
        if (bytepos > b.in_len + 1) { b.err = MS_EREAD; return 0; }
        uint32_t v = bytepos < b.in_len ? b.in[bytepos] : 0u;
        bytepos++;
        return v;
    }

    /* Literal bytes go to the output one by one, not through MsEmit's word gatherer: a byte store per literal instead of the
     * gatherer's compare / merge per literal plus a flush path that a fifth of the warp's lanes walk in every step (4.6 % of the
     * kernel's warp-instructions at 6.7 active threads, profiles/r2_p1lzx_f.txt; the same change was worth 9 % of MSZIP's P1).  The
     * lane's current 32-byte sector stays in L2 between its stores.  Every literal path of the lane does the same - a gathered
     * word is stored whole and would wipe out bytes stored one by one.  q < frame size by construction (lzxd.c:538-651). */
    MS_M void lit_store(uint32_t qq, uint32_t v) { em.out[qq] = (uint8_t) v; }
    /* n raw input bytes to frame positions q.. (uncompressed blocks): whole aligned words directly, the ragged ends bytewise */
    MS_M void raw_store(uint32_t qq, const uint8_t *in, int32_t bp, uint32_t n) {
        if (qq >= em.limit) return;
        if (n > em.limit - qq) n = em.limit - qq;
        const uint8_t *p = in + bp;
#pragma unroll 1
        for (; n && (qq & 3u); n--, qq++, p++) em.out[qq] = *p;
#pragma unroll 1
        for (; n >= 4; n -= 4, p += 4, qq += 4) {
            if ((qq >> 2) < em.wlimit)
                *reinterpret_cast<uint32_t *>(em.out + qq) = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
            else { em.out[qq] = p[0]; em.out[qq + 1] = p[1]; em.out[qq + 2] = p[2]; em.out[qq + 3] = p[3]; }
        }
#pragma unroll 1
        for (; n; n--, qq++, p++) em.out[qq] = *p;
    }

    /* the reference's bit buffer is empty and its byte pointer is at bytepos: go back to bit reading.
     * An odd byte pointer (odd-sized uncompressed block whose pad byte was not skipped because a reset
     * cleared block_type first, lzxd.c:257-270 vs :469-474) moves the 16-bit word grid by one byte. */
    MS_M void enter_bits() {
        if (!bytemode) return;
        if (bytepos & 1) { b.in += 1; b.in_len -= 1; base += 1; bytepos -= 1; ms_bits_rebase(b); }
        ms_bits_seek(b, bytepos & ~3);
        lzx_refill(b);
        if (bytepos & 2) msb_drop(b, 16);
        bytemode = 0;
    }

    /* lzxd.c:138-183: pretree-delta coded lengths; runs are not clamped to `last` */
    MS_M int read_lens(uint8_t *lens, uint32_t first, uint32_t last) {
        uint64_t plo = 0; uint32_t phi = 0; uint32_t lv[16];
#pragma unroll 1
        for (int x = 0; x < 20; x++) {
            lzx_refill(b);
            uint32_t y = lzx_read(b, 4);
            if (x < 16) plo |= (uint64_t) y << (4 * x); else phi |= y << (4 * (x - 16));
        }
        if (b.err) return b.err;
        if (ms_canon_build_h<0, NT>([&](int s) { return (uint32_t) (s < 16 ? (plo >> (4 * s)) : (uint64_t) (phi >> (4 * (s - 16)))) & 15u; },
                                    20, 6, lbo, cnt, pa.sorted, [](uint32_t, uint32_t) { }, (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
#pragma unroll
        for (int j = 0; j < 15; j++) llim[j * NT] = (uint16_t) (lv[j] >> 1);
#pragma unroll 1
        for (uint32_t x = first; x < last;) {
            lzx_refill(b);
            int z = (int) sym_smem(llim, lbo, pa.sorted);
            if (b.err) return b.err;
            if (z == 17) { uint32_t y = lzx_read(b, 4) + 4; if (b.err) return b.err; while (y--) { lens[x * 32] = 0; x++; } }
            else if (z == 18) { uint32_t y = lzx_read(b, 5) + 20; if (b.err) return b.err; while (y--) { lens[x * 32] = 0; x++; } }
            else if (z == 19) {
                uint32_t y = lzx_read(b, 1) + 4;
                lzx_refill(b);
                z = (int) sym_smem(llim, lbo, pa.sorted);
                if (b.err) return b.err;
                z = (int) lens[x * 32] - z; if (z < 0) z += 17;
                while (y--) { lens[x * 32] = (uint8_t) z; x++; }
            }
            else { z = (int) lens[x * 32] - z; if (z < 0) z += 17; lens[x * 32] = (uint8_t) z; x++; }
        }
        return 0;
    }

    MS_M int build_main() {
        uint8_t *l = main_len; uint32_t lv[16];
        if constexpr (H8) {
#pragma unroll 1
            for (int w = 0; w < HEADN / 16; w++) mhi[w * NT] = 0;
            uint8_t *lo = mlo; uint32_t *hi = mhi;
            if (ms_canon_build_h<0, NT>([&](int s) { return (uint32_t) l[s * 32]; }, (int) nsyms_eff, 12, mbo, cnt, ma.sorted,
                                        [=](uint32_t k, uint32_t sym) {
                                            if (k < (uint32_t) HEADN) { lo[k * Shared::ROW] = (uint8_t) sym; hi[(k >> 4) * NT] |= (sym >> 8) << ((k & 15u) * 2u); }
                                        }, (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
        }
        else {
            uint16_t *hd = mhead;
            if (ms_canon_build_h<0, NT>([&](int s) { return (uint32_t) l[s * 32]; }, (int) nsyms_eff, 12, mbo, cnt, ma.sorted,
                                        [=](uint32_t k, uint32_t sym) { if (k < (uint32_t) HEADN) hd[k * NT] = (uint16_t) sym; }, (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
        }
#pragma unroll
        for (int j = 0; j < 15; j++) mlim[j] = lv[j];
        return 0;
    }
    MS_M int build_length() {                 /* BUILD_TABLE_MAYBE_EMPTY, lzxd.c:111-125 */
        uint8_t *l = len_len; uint32_t lv[16];
        length_empty = 0;
        uint8_t *lh = lhead;
        int rc;
        rc = ms_canon_build_h<LUTB, NT>([&](int s) { return (uint32_t) l[s * 32]; }, LZX_LEN_SYMS, 12, lbo, cnt, la.sorted,
                                       [=](uint32_t k, uint32_t sym) { if (QL && k < (uint32_t) LZX_LHEAD) lh[k * NT] = (uint8_t) sym; }, llut, lv);
        if (rc) {
#pragma unroll 1
            for (int i = 0; i < LZX_LEN_SYMS; i++) if (l[i * 32] > 0) return MS_EDECRUNCH;
            length_empty = 1;
        }
        else {
#pragma unroll
            for (int j = 0; j < 15; j++) llim[j * NT] = (uint16_t) (lv[j] >> 1);
        }
        return 0;
    }
    MS_M int build_aligned() {
        uint32_t al = aligned_lens; uint32_t lv[16];
        if constexpr (H8) {
            /* 8 symbols, lengths 0..7: everything fits four words.  make_decode_table (readhuff.h:83-176, 7 table bits)
             * accepts exactly the complete codes */
            uint32_t c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
            for (int k = 0; k < 8; k++) c[(al >> (3 * k)) & 7u]++;
            uint32_t lim = 0, off = 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0, n = 0;
#pragma unroll
            for (int len = 1; len <= 7; len++) {
                w2 |= off << (4 * (len - 1));
                lim += c[len] << (7 - len); off += c[len];
                if (len <= 4) w0 |= lim << (8 * (len - 1)); else w1 |= lim << (8 * (len - 5));
#pragma unroll
                for (int k = 0; k < 8; k++) if (((al >> (3 * k)) & 7u) == (uint32_t) len) { w3 |= (uint32_t) k << (3 * n); n++; }
            }
            if (lim != 128u) return MS_EDECRUNCH;
            atree[0] = w0; atree[NT] = w1; atree[2 * NT] = w2; atree[3 * NT] = w3;
            return 0;
        }
        else {
            if (ms_canon_build<0, NT>([&](int s) { return (al >> (3 * s)) & 7u; }, 8, 7, abo, cnt, aa.sorted, (uint16_t *) nullptr, 0, (uint16_t *) nullptr, lv)) return MS_EDECRUNCH;
#pragma unroll
            for (int j = 0; j < 15; j++) alim[j * NT] = (uint16_t) (lv[j] >> 1);
            return 0;
        }
    }

    /* lzxd.c:465-523: read a block header.  Returns 0 or an MSPACK_ERR_* */
    MS_M int block_header() {
        if (block_type == 3 && (block_length & 1)) { (void) raw_byte(); if (b.err) return b.err; }     /* :469-474 */
        enter_bits();
        lzx_refill(b);
        block_type = lzx_read(b, 3);
        uint32_t i = lzx_read(b, 16);
        lzx_refill(b);
        uint32_t j = lzx_read(b, 8);
        if (b.err) return b.err;
        block_remaining = block_length = (i << 8) | j;
        if (block_type == 2) {
            uint32_t al = 0;
#pragma unroll 1
            for (int k = 0; k < 8; k++) { lzx_refill(b); al |= lzx_read(b, 3) << (3 * k); }
            if (b.err) return b.err;
            aligned_lens = al;
            int e = build_aligned(); if (e) return e;
        }
        if (block_type == 1 || block_type == 2) {
            int e;
            if ((e = read_lens(main_len, 0, 256))) return e;
            if ((e = read_lens(main_len, 256, 256 + num_offsets))) return e;
            if ((e = build_main())) return e;
            if (main_len[0xE8 * 32] != 0) intel_started = 1;                                           /* :495 */
            if ((e = read_lens(len_len, 0, 249))) return e;
            if ((e = build_length())) return e;
            return 0;
        }
        if (block_type == 3) {
            intel_started = 1;                                                                         /* :503 */
            /* :505-507 discard 1..16 bits up to the next 16-bit word */
            int r = b.bc & 15;
            if (r == 0) { lzx_refill(b); lzx_check(b, 16); if (b.err) return b.err; r = 16; }
            msb_drop(b, r);
            bytepos = b.ipos - (b.bc >> 3); bytemode = 1;
            uint32_t v[3];
#pragma unroll 1
            for (int k = 0; k < 3; k++) { uint32_t x = raw_byte(); x |= raw_byte() << 8; x |= raw_byte() << 16; x |= raw_byte() << 24; v[k] = x; }
            if (b.err) return b.err;
            R0 = v[0]; R1 = v[1]; R2 = v[2];
            return 0;
        }
        return MS_EDECRUNCH;                                                                           /* :519-522 */
    }

    MS_M void fail(int err) { status = err; done = 1; phase = PH_IDLE; }

    /* lzxd.c:441-444: LZX DELTA, the 16-bit chunk size in front of every frame: ENSURE_BITS(16), REMOVE_BITS(16).  With an
     * empty bit buffer (after the raw bytes of an uncompressed block) that is one two-byte fetch at the byte pointer, wherever
     * it stands, and the buffer is empty again afterwards */
    MS_M void skip_chunk_size() {
        if (bytemode) { (void) raw_byte(); (void) raw_byte(); }
        else { lzx_refill(b); lzx_check(b, 16); msb_drop(b, 16); }
    }

    /* lzxd.c:419-461: frame prologue (reset interval, intel header, frame size) */
    MS_M void frame_start() {
        frame_start_pos = produced;
        if (u->reset_interval && (frame % u->reset_interval) == 0) reset_state();                     /* :423-438 */
        if (DELTA && is_delta) { skip_chunk_size(); if (b.err) { fail(b.err); return; } }             /* :441-444 */
        if (!header_read) {                                                                          /* :447-453 */
            enter_bits();
            lzx_refill(b);
            uint32_t hi = 0, lo = 0;
            if (lzx_read(b, 1)) { lzx_refill(b); hi = lzx_read(b, 16); lzx_refill(b); lo = lzx_read(b, 16); }
            if (b.err) { fail(b.err); return; }
            intel_filesize = (int32_t) ((hi << 16) | lo); header_read = 1;
        }
        frame_size = ms_min(MS_FRAME, u->out_len - produced);                                        /* :458-461 */
        bytes_todo = (int32_t) frame_size; q = 0;
        emit_begin(em, recs + (size_t) f * MS_MAXREC, uout + produced, frame_size);
        phase = PH_BLOCK;
    }

    /* lzxd.c:463-532 + :654-671: next run of the frame; uncompressed runs are copied right here */
    MS_M void next_run() {
        if (bytes_todo <= 0) { phase = PH_END; return; }
        if (block_remaining == 0) { int e = block_header(); if (e) { fail(e); return; } }
        this_run = (int32_t) block_remaining;
        if (this_run > bytes_todo) this_run = bytes_todo;
        bytes_todo -= this_run; block_remaining -= (uint32_t) this_run;
        if (block_type == 1 || block_type == 2) { if (this_run > 0) phase = PH_DECODE; return; }
        if (block_type == 3) {
            if (this_run > 0 && bytepos + this_run <= b.in_len) {      /* the whole run lies inside the input: bulk copy */
                raw_store(q, b.in, bytepos, (uint32_t) this_run);
                bytepos += this_run; q += (uint32_t) this_run; this_run = 0;
            }
#pragma unroll 1
            while (this_run > 0) { lit_store(q, raw_byte()); q++; this_run--; }
            if (b.err) fail(b.err);
            return;
        }
        fail(MS_EDECRUNCH);
    }

    MS_M void frame_end() {
        /* :696-697 re-align; after raw bytes the reference's bit buffer is empty and nothing happens */
        if (!bytemode && (b.bc & 15)) { lzx_refill(b); lzx_check(b, 16); if (b.err) { fail(b.err); return; } msb_drop(b, b.bc & 15); }
        emit_end(em, frame_size);
        MsFrameInfo fi; fi.nrec = em.nrec; fi.size = frame_size; fi.g0 = frame_start_pos; fi.valid = 1;
        finfo[f] = fi;
        e8info[frame] = (intel_started && intel_filesize && frame + MSGPU_UNIT_FRAME_BASE(u) < 32768 && frame_size > 10) ? intel_filesize : 0;   /* :706-709 (the stream's frame count) */
        produced += frame_size; frame++; f++;
        if (produced >= u->out_len) {
            done = 1; phase = PH_IDLE;
            /* lzxd.c:419: a request ending exactly on a frame boundary runs one more zero-sized frame pass; at a
             * reset point that re-reads the intel header and tops the bit buffer up (see oracle/port/mspack_port.c) -
             * the only effect is MSPACK_ERR_READ on an exactly-cut unit */
            if (DELTA && is_delta && (u->out_len % MS_FRAME) == 0) {          /* the extra pass reads its chunk size too */
                skip_chunk_size();
                if (b.err) { status = b.err; return; }
            }
            if ((u->out_len % MS_FRAME) == 0 && u->reset_interval && (frame % u->reset_interval) == 0) {
                int32_t bp;
                if (bytemode) { if (bytepos & 1) { b.in += 1; b.in_len -= 1; bytepos -= 1; ms_bits_rebase(b); } bp = bytepos; }
                else bp = b.ipos - (b.bc >> 3);
                uint32_t hb = (bp + 1 < b.in_len) ? b.in[bp + 1] : 0u;
                int32_t need = (hb & 0x80) ? bp + 8 : bp + 4;
                if (bp + 2 > b.in_len + 2 || need > b.in_len + 2) status = MS_EREAD;
            }
        }
        else phase = (f < max_frames) ? PH_FRAME : PH_IDLE;
    }

    MS_M void post_step() { }
    MS_M void service() {
#pragma unroll 1
        while (phase >= PH_FRAME && phase != PH_PARK) {
            if (phase == PH_FRAME) phase = PH_PARK;                   /* wait for the warp (msgpu_core.cuh PH_PARK) */
            else if (phase == (PH_FRAME | 0x100u)) frame_start();
            else if (phase == PH_BLOCK) next_run();
            else frame_end();
        }
    }

    /* main-tree symbol: length from the register limits, symbol from the shared-memory head or global scratch */
    MS_M uint32_t main_sym(bool careful) {
        if (careful) lzx_check(b, 16);
        uint32_t v16 = msb_peek(b, 16);
        int len = ms_canon_len(mlim, v16);
        uint32_t idx = mbo.index(v16, len);
        msb_drop(b, len);
        if constexpr (H8) {
            if (idx < (uint32_t) HEADN) return (uint32_t) mlo[idx * Shared::ROW] | (((mhi[(idx >> 4) * NT] >> ((idx & 15u) * 2u)) & 3u) << 8);
            return (uint32_t) ma.sorted[idx * MS_WARP];
        }
        else return idx < (uint32_t) HEADN ? (uint32_t) mhead[idx * NT] : (uint32_t) ma.sorted[idx * MS_WARP];
    }

    /* the hot step (lzxd.c:538-651): one literal, or one match with its length / offset
     * fields.  Batching literals keeps the lanes that are inside a literal run busy while the others
     * handle a match, which is the longer path. */
    MS_M void step() {
        /* `careful` = the unit's input ends within the next 24 bytes: only then can any of this step's reads (at most
         * two 4-byte refills) trip the reference's end-of-input rule, so only then are the exact checks compiled in */
        if (MS_UNLIKELY(b.ipos + (DELTA ? 32 : 24) > b.in_len)) step_plain<true>(); else step_plain<false>();     /* DELTA: one more refill */
    }
    /* the hot step (lzxd.c:538-651): one literal, or one match with its length / offset fields */
    template <bool careful> MS_M void step_plain() {
        lzx_refill(b);
        uint32_t sym = main_sym(careful);
        if (sym < 256) { lit_store(q, sym); q++; this_run--; }
        else {
            sym -= 256;
            uint32_t ml = sym & 7, slot = sym >> 3, off;
            /* A LENGTH symbol beyond the LUT and the shared-memory head comes from L2.  Its value is consumed INSIDE that branch (the
             * empty asm pins the wait for the load there): when the add sat behind the branches' join - `ml += length_sym()`, or the
             * load issued early and `if (lslow) ml += lgv` after the offset decode - the one add every match step executes carried the
             * load's scoreboard in its wait mask, and the same scoreboard guards the input word the bit reader prefetches one refill
             * ahead: 60 % of the kernel's long-scoreboard stall samples sat on that add although the text batch never takes the
             * branch (profiles/r2_p1lzx_z.txt) - every step waited for some lane's prefetch to land. */
            if (ml == 7) {
                if (length_empty) { fail(b.err ? b.err : MS_EDECRUNCH); return; }                    /* :555-558 */
                if (careful) lzx_check(b, 16);
                const uint32_t e = llut[msb_peek(b, LUTB) * NT];
                if (e & 15) { msb_drop(b, (int) (e & 15)); ml += e >> 4; }
                else {
                    const uint32_t v16 = msb_peek(b, 16);
                    const int len = ms_canon_len_smem<NT>(llim, v16);
                    const uint32_t idx = lbo.index(v16, len);
                    msb_drop(b, len);
                    if (QL && idx < (uint32_t) LZX_LHEAD) ml += lhead[idx * NT];
                    else {
                        uint32_t lgv = la.sorted[idx * MS_WARP];
#if defined(__CUDACC__) && !defined(MSGPU_EMULATE)
                        asm volatile("" : "+r"(lgv));
#endif
                        ml += lgv;
                    }
                }
            }
            ml += 2;

            if (slot < 3) {                                         /* repeated offsets, lzxd.c:590-600, as selects */
                const uint32_t r0 = R0;
                off = slot == 0 ? r0 : (slot == 1 ? R1 : R2);
                R1 = slot == 1 ? r0 : R1; R2 = slot == 2 ? r0 : R2; R0 = off;
            }
            else {
                /* extra_bits[] / position_base[] (lzxd.c:199-255) in closed form */
                const uint32_t extra = slot < 4 ? 0 : ((slot >> 1) - 1 < 17 ? (slot >> 1) - 1 : 17);
                const uint32_t pbase = slot < 4 ? slot : (slot < 38 ? (2u + (slot & 1)) << ((slot >> 1) - 1) : (slot - 34) << 17);
                off = pbase - 2;
                /* the refill in front of the offset bits only when the bits at hand do not cover them (extra + 4 <= 21 bits): with 32
                 * lanes per warp an unconditional "below 32 bits" refill body runs in almost every step, this one in ~15 % of them
                 * (measured: P1 9.02 -> 8.52 ms, profiles/r2_variants.txt shape 31) */
                if (b.bc < (int) extra + 4) lzx_refill(b);
                if (block_type == 2 && extra >= 3) {
                    if (extra > 3) { if (careful) lzx_check(b, (int) extra - 3); off += msb_peek(b, (int) extra - 3) << 3; msb_drop(b, (int) extra - 3); }
                    if constexpr (H8) off += aligned_sym(careful); else off += sym_smem(alim, MsBo32<NT>{ abo }, aa.sorted, careful);
                }
                else if (extra) { if (careful) lzx_check(b, (int) extra); off += msb_peek(b, (int) extra); msb_drop(b, (int) extra); }
                R2 = R1; R1 = R0; R0 = off;
            }
            if (DELTA && is_delta && ml == 257) {                    /* lzxd.c:589-611: the longest length announces more */
                lzx_refill(b);
                if (careful) lzx_check(b, 3);
                const uint32_t p3 = msb_peek(b, 3);
                const int pre = p3 < 4 ? 1 : (p3 < 6 ? 2 : 3), nb = p3 < 4 ? 8 : (p3 < 6 ? 10 : (p3 == 6 ? 12 : 15));
                msb_drop(b, pre);
                if (careful) lzx_check(b, nb);
                ml += msb_peek(b, nb) + (p3 < 4 ? 0u : (p3 < 6 ? 0x100u : (p3 == 6 ? 0x500u : 0u)));
                msb_drop(b, nb);
            }
            if (careful && b.err) { fail(b.err); return; }
            if (!resolve_match(ml, off)) return;
        }
        if (careful && b.err) { fail(b.err); return; }
        if (this_run <= 0) phase = PH_BLOCK;
    }
    /* lzxd.c:613-634 restated (window_posn = G mod window_size, lzx->offset = frame start).  Fast path: a source
     * inside the unit during the first lap of the window (every unit up to 2^window_bits bytes never leaves it) */
    MS_M bool resolve_match(uint32_t ml, uint32_t off) {
        uint32_t G = frame_start_pos + q, eff = off;
        if (MS_UNLIKELY(off - 1u >= G || G + ml > window_size)) {
            uint32_t wpr = G & (window_size - 1);
            bool bad = (wpr + ml > window_size);
            if (off > wpr) {
                /* :622-628: beyond the decoded data is fine only inside the reference data (DELTA) */
                bad = bad || (off > frame_start_pos && (!DELTA || off - wpr > ref_len)) || (off - wpr > window_size);
                if (off > window_size) eff = off - window_size;
            }
            if (eff == 0) eff = window_size;          /* source == destination: the bytes one window lap back */
            if (bad) { fail(MS_EDECRUNCH); return false; }
        }
        if (MS_UNLIKELY((int32_t) ml > this_run)) { fail(MS_EDECRUNCH); return false; }   /* :678-693 every overrun ends in an error */
        if (DELTA) emit_match_wide(em, q, ml, eff);
        else emit_match(em, q, ml, eff);
        q += ml; this_run -= (int32_t) ml;
        return true;
    }
    MS_M void begin(const msgpu_unit *unit, const uint8_t *in_base, const MsUnitState &st, MsRec *r, uint8_t *l, MsFrameInfo *fi,
                    int32_t *e8, int nframes) {
        u = unit; recs = r; uout = l; finfo = fi; e8info = e8; max_frames = nframes; f = 0; q = 0; this_run = 0; bytes_todo = 0;
        frame_start_pos = 0; frame_size = 0;
#pragma unroll 1
        for (int k = 0; k < nframes; k++) { MsFrameInfo z; z.nrec = 0; z.size = 0; z.g0 = 0; z.valid = 0; fi[k] = z; }
        const int wb = unit->window_bits;
        is_delta = (DELTA && (unit->flags & MSGPU_FLAG_LZX_DELTA)) ? 1u : 0u; ref_len = DELTA ? MSGPU_UNIT_REF_BYTES(unit) : 0u;
        const uint32_t slots = wb == 15 ? 30u : wb == 16 ? 32u : wb == 17 ? 34u : wb == 18 ? 36u : wb == 19 ? 38u : wb == 20 ? 42u :
                               (DELTA && wb > 21) ? (wb == 22 ? 66u : wb == 23 ? 98u : wb == 24 ? 162u : 290u) : 50u;     /* position_slots[], lzxd.c:209-211 */
        window_size = 1u << (wb & 31); num_offsets = slots << 3;
        nsyms_eff = 256 + num_offsets + 51; if (nsyms_eff > LZX_MAIN_MAX) nsyms_eff = LZX_MAIN_MAX;
        if (!st.started) {
            done = 0; status = 0; produced = 0; frame = 0;
            if (DELTA && is_delta ? (wb < 17 || wb > 25) : (wb < 15 || wb > 21)) { status = MS_ENOMEM; done = 1; }      /* lzxd_init returns NULL -> cabd.c:1255 */
            else if (DELTA && ref_len && (!is_delta || ref_len > (1u << wb))) { status = MS_EARGS; done = 1; }              /* lzxd.c:355-366 */
            ms_bits_init(b, in_base + unit->in_off, unit->in_len);
            base = 0; bytemode = 0; bytepos = 0; intel_filesize = 0; intel_started = 0; length_empty = 0; aligned_lens = 0;
            R0 = R1 = R2 = 1; header_read = 0; block_remaining = 0; block_type = 0; block_length = 0;
            if (!done) reset_state();
            if (unit->out_len == 0) done = 1;
        }
        else {
            done = st.done; status = st.status; produced = st.produced; frame = st.frame;
            base = st.base; bytemode = st.bytemode; bytepos = st.ipos;
            ms_bits_restore(b, in_base + unit->in_off + base, unit->in_len - base, bytemode ? 0 : st.ipos, (int32_t) st.bc, ((uint64_t) st.bb_hi << 32) | st.bb_lo);
            R0 = st.R0; R1 = st.R1; R2 = st.R2; block_type = st.block_type; block_length = st.block_length;
            block_remaining = st.block_remaining; header_read = st.header_read; intel_filesize = (int32_t) st.intel_filesize;
            intel_started = st.intel_started; length_empty = st.length_empty; aligned_lens = st.aligned_lens;
            if (!done && block_remaining > 0 && (block_type == 1 || block_type == 2)) {
                /* the shared-memory tables do not survive a launch: rebuild them from the stored lengths */
                (void) build_main(); (void) build_length();
                if (block_type == 2) (void) build_aligned();
            }
        }
        phase = done ? PH_IDLE : PH_FRAME;
    }
    MS_M void end(MsUnitState &st) {
        st.started = 1; st.done = done; st.status = status; st.produced = produced; st.frame = frame;
        st.base = base; st.bytemode = bytemode;
        st.ipos = bytemode ? bytepos : b.ipos; st.bc = (uint32_t) b.bc; st.bb_lo = (uint32_t) b.bb; st.bb_hi = (uint32_t) (b.bb >> 32);
        st.R0 = R0; st.R1 = R1; st.R2 = R2; st.block_type = block_type; st.block_length = block_length;
        st.block_remaining = block_remaining; st.header_read = header_read; st.intel_filesize = (uint32_t) intel_filesize;
        st.intel_started = intel_started; st.length_empty = length_empty; st.aligned_lens = aligned_lens;
    }
};
