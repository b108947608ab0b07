/* mspack_port.c - TEST INFRASTRUCTURE ONLY (oracle, kind "port").
 *
 * A plain-C restatement, written from scratch for this repository, of the three CAB-folder
 * decoders of kyz/libmspack.  It is NOT the product and is never linked into it; it exists so the
 * parity tests have an independent CPU answer on machines where the reference sources are not
 * available (the GPU box), and it is itself pinned against oracle/_ref/libmspack_ref.so (the
 * unmodified reference) and the reference's golden vectors by tests/test_oracle.py.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/libmspack/mspack/).  The restatement differs from the reference in FORM:
 *   - one-shot buffer -> buffer decode of a whole unit instead of a resumable stream;
 *   - the bit readers work on an absolute bit position p into the unit's input instead of a
 *     32-bit bit buffer with 8/16-bit refills; the reference's EOF rule ("two zero
 *     bytes are supplied once, the next refill fails", readbits.h:192-214) becomes "asking for a
 *     byte at index >= in_len + 2 is MSPACK_ERR_READ";
 *   - Huffman decoding uses canonical first-code arrays, with make_decode_table's acceptance
 *     rule restated (readhuff.h:83-176: accept iff the codes no longer than TABLEBITS fill the
 *     table exactly - longer codes are then unreachable - or else all codes <= 16 bits have a
 *     Kraft sum of exactly 1);
 *   - LZX and Quantum write straight into the unit's output buffer, which doubles as the sliding
 *     window ("the unit's output buffer is the window"); bytes before the
 *     start of the unit read as zero.  LZX E8 translation is a post-pass over finished frames.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#include "../../include/msgpu.h"

#define ERR_OK       MSGPU_ERR_OK
#define ERR_ARGS     MSGPU_ERR_ARGS
#define ERR_READ     MSGPU_ERR_READ
#define ERR_DECRUNCH MSGPU_ERR_DECRUNCH
#define ERR_NOMEM    MSGPU_ERR_NOMEMORY

#define FRAME 32768u

/* ------------------------------------------------------------------------------------------
 * bit input: absolute bit position p over in[0..in_len), bytes past the end read as zero
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *in;
    uint32_t in_len;
    uint64_t p;         /* bits consumed */
    uint64_t loaded;    /* Quantum only: bits fetched by the reference's 2-byte refills */
    uint32_t base;      /* LZX only: byte offset p is measured from (0 or 1, see lzx_enter_bits) */
    int bytemode;       /* LZX only: the reference's bit buffer is empty and it is reading raw bytes */
    int err;
    /* MSZIP repair mode only (zero otherwise): after a repaired block the reference goes on with the bits it had buffered at
     * its last STORE_BITS followed by the bytes at its (possibly refilled) buffer pointer - the stream seen from here on is
     * stale[0..nstale) followed by in[real0..] */
    uint8_t stale[8]; uint32_t nstale; uint64_t real0;        /* (nstale <= 4) */
    uint64_t fetched;   /* bytes of that stream the reference has fetched so far (its ENSURE_BITS / READ_IF_NEEDED calls) */
} bitin;

static inline uint32_t in_byte(const bitin *b, uint64_t i) {
    if (i < b->nstale) return b->stale[i & 7];
    i = i - b->nstale + b->real0 + b->base;
    return i < b->in_len ? b->in[i] : 0u;
}

/* the reference would have to fetch input bytes [0, need_bytes) : legal up to in_len + 2 */
static inline int fetch_ok(bitin *b, uint64_t need_bytes) {
    if (need_bytes > b->nstale && need_bytes - b->nstale + b->real0 + b->base > (uint64_t) b->in_len + 2u) { b->err = ERR_READ; return 0; }
    if (need_bytes > b->fetched) b->fetched = need_bytes;
    return 1;
}

/* ---- LSB-first, byte refills (MSZIP; readbits.h:161-166, mszipd.c:23-26) ---- */
static inline int lsb_ensure(bitin *b, unsigned n) { return fetch_ok(b, (b->p + n + 7) >> 3); }
static inline uint32_t lsb_peek(const bitin *b, unsigned n) {          /* n <= 25 */
    uint64_t i = b->p >> 3; unsigned s = (unsigned) (b->p & 7);
    uint32_t v = in_byte(b, i) | (in_byte(b, i + 1) << 8) | (in_byte(b, i + 2) << 16) | (in_byte(b, i + 3) << 24);
    return (v >> s) & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1u));
}
static inline uint32_t lsb_read(bitin *b, unsigned n) {
    uint32_t v;
    if (n == 0) return 0;
    if (!lsb_ensure(b, n)) return 0;
    v = lsb_peek(b, n); b->p += n; return v;
}

/* ---- MSB-first over 16-bit little-endian words (LZX; readbits.h:155-160, lzxd.c:86-91) ---- */
static inline uint32_t lzx_word(const bitin *b, uint64_t w) { return in_byte(b, 2 * w) | (in_byte(b, 2 * w + 1) << 8); }
static inline int lzx_ensure(bitin *b, unsigned n) { return fetch_ok(b, ((b->p + n + 15) >> 4) << 1); }
static inline uint32_t lzx_peek(const bitin *b, unsigned n) {          /* 1 <= n <= 32 */
    uint64_t w = b->p >> 4; unsigned s = (unsigned) (b->p & 15);
    uint64_t v = ((uint64_t) lzx_word(b, w) << 32) | ((uint64_t) lzx_word(b, w + 1) << 16) | lzx_word(b, w + 2);
    return (uint32_t) ((v << (16 + s)) >> 32) >> (32 - n);
}
static inline uint32_t lzx_read(bitin *b, unsigned n) {                /* READ_BITS, n >= 1 */
    uint32_t v;
    if (!lzx_ensure(b, n)) return 0;
    v = lzx_peek(b, n); b->p += n; return v;
}

/* The reference refills 16-bit words from its BYTE pointer.  After the raw bytes of an uncompressed
 * block that pointer can be odd (an odd-sized block whose pad byte was not skipped because a reset
 * cleared block_type first, lzxd.c:257-270 vs :469-474); words are then fetched from odd offsets.
 * p stays "bits since base", with base moved by one byte so that p is 16-bit aligned again. */
static inline void lzx_enter_bits(bitin *b) {
    if (b->bytemode) { if ((b->p >> 3) & 1) { b->base++; b->p -= 8; } b->bytemode = 0; }
}

/* ---- MSB-first over 16-bit big-endian words == a plain big-endian bit stream (Quantum;
 *      qtmd.c:30-35).  The reference fetches two bytes at a time from a BYTE-granular pointer,
 *      so the fetched extent `loaded` is tracked explicitly for the EOF rule. ---- */
static inline int qtm_fetch2(bitin *b) {               /* READ_BYTES: two READ_IF_NEEDED */
    if (!fetch_ok(b, (b->loaded >> 3) + 2)) return 0;
    b->loaded += 16; return 1;
}
static inline int qtm_ensure(bitin *b, unsigned n) {
    while (b->loaded - b->p < n) if (!qtm_fetch2(b)) return 0;
    return 1;
}
static inline uint32_t qtm_peek(const bitin *b, unsigned n) {          /* 1 <= n <= 24 */
    uint64_t i = b->p >> 3; unsigned s = (unsigned) (b->p & 7);
    uint32_t v = (in_byte(b, i) << 24) | (in_byte(b, i + 1) << 16) | (in_byte(b, i + 2) << 8) | in_byte(b, i + 3);
    return (v << s) >> (32 - n);
}
static inline uint32_t qtm_read(bitin *b, unsigned n) {                /* READ_BITS, n >= 1 */
    uint32_t v;
    if (!qtm_ensure(b, n)) return 0;
    v = qtm_peek(b, n); b->p += n; return v;
}
static inline uint32_t qtm_read_many(bitin *b, unsigned n) {           /* READ_MANY_BITS, readbits.h:143-153 */
    uint32_t v = 0;
    while (n > 0) {
        unsigned left, run;
        if (b->loaded - b->p <= 16) if (!qtm_fetch2(b)) return 0;
        left = (unsigned) (b->loaded - b->p);
        run = left < n ? left : n;
        v = (v << run) | qtm_peek(b, run); b->p += run; n -= run;
    }
    return v;
}

/* ------------------------------------------------------------------------------------------
 * canonical Huffman decoder with make_decode_table's acceptance rule (readhuff.h:83-176)
 * ------------------------------------------------------------------------------------------ */
#define HUFF_MAXSYMS 2576   /* LZX_MAINTREE_MAXSYMBOLS, lzx.h:38 */
typedef struct {
    uint16_t sorted[HUFF_MAXSYMS];  /* symbols in canonical order (length asc, index asc) */
    uint32_t first[18];             /* first[l]: canonical code of the first l-bit symbol, left-aligned to 16 bits */
    uint32_t limit[18];             /* limit[l]: first[l] + count[l] << (16-l)  (exclusive, 16-bit aligned) */
    uint16_t offset[18];            /* index in sorted[] of the first l-bit symbol */
    const uint8_t *lens;
    int maxlen;
} huff;

/* returns 0 if the reference's make_decode_table(nsyms, tablebits, lens) would succeed */
static int huff_build(huff *h, const uint8_t *lens, unsigned nsyms, unsigned tablebits) {
    uint32_t count[17]; uint32_t sum_short = 0, sum_all = 0, code = 0; unsigned l, s, idx = 0;
    int maxlen = 16;
    memset(count, 0, sizeof(count));
    for (s = 0; s < nsyms; s++) if (lens[s] >= 1 && lens[s] <= 16) count[lens[s]]++;
    for (l = 1; l <= 16; l++) {
        sum_all += count[l] << (16 - l);
        if (l <= tablebits) sum_short = sum_all;
    }
    if (sum_short > 65536u) return 1;                    /* readhuff.h:108 table overrun */
    if (sum_short == 65536u) maxlen = (int) tablebits;   /* :122 complete - longer codes never reachable */
    else if (sum_all != 65536u) return 1;                /* :147 / :175 */
    h->lens = lens; h->maxlen = maxlen;
    for (l = 1; l <= 16; l++) {
        uint32_t c = ((int) l <= maxlen) ? count[l] : 0;
        h->first[l] = code; h->offset[l] = (uint16_t) idx;
        code += c << (16 - l); h->limit[l] = code; idx += c;
    }
    for (l = 1, idx = 0; (int) l <= maxlen; l++)
        for (s = 0; s < nsyms; s++) if (lens[s] == l) h->sorted[idx++] = (uint16_t) s;
    return 0;
}

/* v16 = next 16 stream bits, first bit in the MSB.  Returns symbol, *len = code length. */
static inline unsigned huff_decode(const huff *h, uint32_t v16, unsigned *len) {
    int l;
    for (l = 1; l <= h->maxlen; l++) {
        if (v16 < h->limit[l]) {
            *len = (unsigned) l;
            return h->sorted[h->offset[l] + ((v16 - h->first[l]) >> (16 - l))];
        }
    }
    *len = 0; return 0xFFFF;    /* unreachable for an accepted table */
}

static inline uint32_t rev16(uint32_t v) {
    v = ((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u);
    v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u);
    v = ((v & 0x0F0Fu) << 4) | ((v >> 4) & 0x0F0Fu);
    return ((v & 0x00FFu) << 8) | ((v >> 8) & 0x00FFu);
}

/* ==========================================================================================
 * MSZIP  (mszipd.c)
 * ========================================================================================== */
static const uint16_t zip_lit_lengths[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27,
    31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };                 /* mszipd.c:47-50 */
static const uint16_t zip_dist_offsets[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129,
    193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 }; /* :53-56 */
static const uint8_t zip_lit_extrabits[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2,
    2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };                                  /* :59-62 */
static const uint8_t zip_dist_extrabits[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6,
    6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };                          /* :65-68 */
static const uint8_t zip_bitlen_order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 }; /* :71-73 */

typedef struct {
    bitin b;
    uint8_t window[FRAME];          /* mszip.h:70: history persists across CK blocks */
    uint32_t window_posn, bytes_output;
    uint8_t lit_len[288], dist_len[32];
    huff lit, dist, bl;
    uint64_t store_p, store_f;      /* bit position / fetched bytes at the reference's last STORE_BITS (repair mode) */
} zipst;
#define ZIP_STORE(z) do { (z)->store_p = (z)->b.p; (z)->store_f = (z)->b.fetched; } while (0)

/* mszipd.c:38-45 FLUSH_IF_NEEDED + :323-333 mszipd_flush_window */
static inline int zip_flush_if_needed(zipst *z) {
    if (z->window_posn == FRAME) {
        z->bytes_output += FRAME;
        if (z->bytes_output > FRAME) return -3;         /* INF_ERR_FLUSH */
        z->window_posn = 0;
    }
    return 0;
}

static inline unsigned zip_huffsym(zipst *z, const huff *h, int *err) {    /* READ_HUFFSYM, readhuff.h:39-46 */
    unsigned len, sym;
    if (!lsb_ensure(&z->b, 16)) { *err = ERR_READ; return 0; }
    sym = huff_decode(h, rev16(lsb_peek(&z->b, 16)), &len);
    z->b.p += len;
    return sym;
}

/* mszipd.c:91-151 zip_read_lens.  Returns 0, a negative INF_ERR_*, or ERR_READ (>0). */
static int zip_read_lens(zipst *z) {
    uint8_t bl_len[19], lens[288 + 32];
    unsigned lit_codes, dist_codes, bitlen_codes, i, code, last_code = 0, run;
    int err = 0;
    lit_codes = lsb_read(&z->b, 5) + 257; dist_codes = lsb_read(&z->b, 5) + 1; bitlen_codes = lsb_read(&z->b, 4) + 4;
    if (z->b.err) return ERR_READ;
    if (lit_codes > 288) return -5;
    if (dist_codes > 32) return -5;
    for (i = 0; i < bitlen_codes; i++) bl_len[zip_bitlen_order[i]] = (uint8_t) lsb_read(&z->b, 3);
    while (i < 19) bl_len[zip_bitlen_order[i++]] = 0;
    if (z->b.err) return ERR_READ;
    if (huff_build(&z->bl, bl_len, 19, 7)) return -6;
    for (i = 0; i < lit_codes + dist_codes; i++) {
        unsigned len;
        if (!lsb_ensure(&z->b, 7)) return ERR_READ;     /* :117 ENSURE_BITS(7) */
        code = huff_decode(&z->bl, rev16(lsb_peek(&z->b, 7)), &len);
        z->b.p += len;
        if (code < 16) lens[i] = (uint8_t) (last_code = code);
        else {
            switch (code) {
            case 16: run = lsb_read(&z->b, 2) + 3; code = last_code; break;
            case 17: run = lsb_read(&z->b, 3) + 3; code = 0; break;
            case 18: run = lsb_read(&z->b, 7) + 11; code = 0; break;
            default: return -10;
            }
            if (z->b.err) return ERR_READ;
            if (i + run > lit_codes + dist_codes) return -9;
            while (run--) lens[i++] = (uint8_t) code;
            i--;
        }
    }
    (void) err;
    memcpy(z->lit_len, lens, lit_codes); memset(z->lit_len + lit_codes, 0, 288 - lit_codes);
    memcpy(z->dist_len, lens + lit_codes, dist_codes); memset(z->dist_len + dist_codes, 0, 32 - dist_codes);
    ZIP_STORE(z);                                       /* :149 */
    return 0;
}

/* mszipd.c:154-316 inflate.  Returns 0, negative INF_ERR_*, or positive MSPACK_ERR_READ. */
static int zip_inflate(zipst *z) {
    unsigned last_block, block_type;
    int err = 0;
    do {
        last_block = lsb_read(&z->b, 1);
        block_type = lsb_read(&z->b, 2);
        if (z->b.err) return ERR_READ;
        if (block_type == 0) {
            /* stored: go to byte boundary, LEN / NLEN, raw bytes (:165-207) */
            uint8_t lb[4]; unsigned i, length, comp;
            z->b.p = (z->b.p + 7) & ~(uint64_t) 7;
            for (i = 0; i < 4; i++) {
                if (!fetch_ok(&z->b, (z->b.p >> 3) + 1)) return ERR_READ;
                lb[i] = (uint8_t) in_byte(&z->b, z->b.p >> 3); z->b.p += 8;
            }
            length = lb[0] | (lb[1] << 8); comp = lb[2] | (lb[3] << 8);
            if (length != (~comp & 0xFFFFu)) return -2;
            while (length > 0) {
                unsigned this_run = length;
                if (this_run > FRAME - z->window_posn) this_run = FRAME - z->window_posn;
                /* READ_IF_NEEDED per refill: bytes up to index in_len+1 exist (zeros) */
                for (i = 0; i < this_run; i++) {
                    if (!fetch_ok(&z->b, (z->b.p >> 3) + 1)) return ERR_READ;
                    z->window[z->window_posn++] = (uint8_t) in_byte(&z->b, z->b.p >> 3); z->b.p += 8;
                }
                length -= this_run;
                if (zip_flush_if_needed(z)) return -3;
            }
        }
        else if (block_type == 1 || block_type == 2) {
            unsigned code, length, distance, match_posn;
            if (block_type == 1) {
                unsigned i = 0;
                while (i < 144) z->lit_len[i++] = 8;
                while (i < 256) z->lit_len[i++] = 9;
                while (i < 280) z->lit_len[i++] = 7;
                while (i < 288) z->lit_len[i++] = 8;
                for (i = 0; i < 32; i++) z->dist_len[i] = 5;
            }
            else {
                int e;
                ZIP_STORE(z);                           /* :223 */
                e = zip_read_lens(z);
                if (e) return e;
            }
            if (huff_build(&z->lit, z->lit_len, 288, 9)) return -7;     /* :230-235 */
            if (huff_build(&z->dist, z->dist_len, 32, 6)) return -8;    /* :237-241 */
            for (;;) {
                code = zip_huffsym(z, &z->lit, &err);
                if (err) return err;
                if (code < 256) {
                    z->window[z->window_posn++] = (uint8_t) code;
                    if (zip_flush_if_needed(z)) return -3;
                }
                else if (code == 256) break;
                else {
                    code -= 257;
                    if (code >= 29) return -11;
                    length = lsb_read(&z->b, zip_lit_extrabits[code]) + zip_lit_lengths[code];
                    if (z->b.err) return ERR_READ;
                    code = zip_huffsym(z, &z->dist, &err);
                    if (err) return err;
                    if (code >= 30) return -12;
                    distance = lsb_read(&z->b, zip_dist_extrabits[code]) + zip_dist_offsets[code];
                    if (z->b.err) return ERR_READ;
                    /* :267-268: a distance beyond window_posn wraps into the previous block's image */
                    match_posn = ((distance > z->window_posn) ? FRAME : 0) + z->window_posn - distance;
                    while (length--) {
                        z->window[z->window_posn++] = z->window[match_posn++];
                        match_posn &= FRAME - 1;
                        if (zip_flush_if_needed(z)) return -3;
                    }
                }
            }
        }
        else return -1;
    } while (!last_block);
    if (z->window_posn) {                       /* :308-311 flush the remaining data */
        z->bytes_output += z->window_posn;
        if (z->bytes_output > FRAME) return -3;
    }
    return 0;
}

/* Repair mode, where the reference goes on after a block it gave up (mszipd.c:404 RESTORE_BITS of the state its last STORE_BITS
 * left, :149 / :223 / :419).  That state is stale in two ways: bit_buffer / bits_left are the bits buffered at the STORE, and
 * i_ptr is the STORE's pointer only if read_input (readbits.h:192-214) has not refilled the input buffer since - it resets
 * i_ptr to the buffer's start.  The input buffer holds one read() of input_buffer_size bytes (flags >> 6 here, 4096 if 0): the
 * chunks are [k * size, (k + 1) * size) of the stream, and the two zero bytes the reference invents at the end of the input form
 * a last chunk of their own.  The stream from here on: the whole bytes left in the stale bit buffer, then the bytes at i_ptr. */
static uint64_t zip_chunk_start(const bitin *b, uint64_t size, uint64_t x) {      /* start of the chunk that holds stream byte x */
    return x >= b->in_len ? b->in_len : x / size * size;
}
static void zip_repair_restart(zipst *z, const msgpu_unit *u) {
    bitin *b = &z->b;
    uint64_t size = MSGPU_UNIT_REF_BYTES(u) ? MSGPU_UNIT_REF_BYTES(u) : 4096, fr_now, fr_store, r, k;
    uint32_t ns, nb; uint8_t tmp[4];
    size = (size + 1) & ~(uint64_t) 1;
    /* real (stream) fetch pointers now and at the STORE */
    fr_now = b->real0 + (b->fetched > b->nstale ? b->fetched - b->nstale : 0);
    fr_store = b->real0 + (z->store_f > b->nstale ? z->store_f - b->nstale : 0);
    r = fr_store;
    if (fr_now > 0 && (fr_store == 0 || zip_chunk_start(b, size, fr_now - 1) != zip_chunk_start(b, size, fr_store - 1)))
        r = zip_chunk_start(b, size, fr_now - 1);
    ns = (uint32_t) (8 * z->store_f - z->store_p);              /* bits buffered at the STORE; :405 drops ns & 7 of them */
    nb = ns >> 3; if (nb > 4) nb = 4;                            /* (the reference's buffer is 32 bits) */
    for (k = 0; k < nb; k++) tmp[k] = (uint8_t) in_byte(b, z->store_f - nb + k);
    memcpy(b->stale, tmp, nb);
    b->nstale = nb; b->real0 = r; b->p = 0; b->fetched = nb;
}

/* mszipd.c:377-460 mszipd_decompress, for one whole unit */
static int port_mszip(const msgpu_unit *u, const uint8_t *in, uint8_t *out, uint32_t *produced) {
    zipst *z = (zipst *) calloc(1, sizeof(zipst));
    uint32_t done = 0; int ret = ERR_OK;
    if (!z) return ERR_NOMEM;
    z->b.in = in; z->b.in_len = u->in_len;
    if (u->flags & MSGPU_FLAG_MSZIP_KWAJ) {                        /* mszipd.c:462-495 mszipd_decompress_kwaj */
        for (;;) {
            uint32_t block_len, c, n; int error;
            z->b.p = (z->b.p + 7) & ~(uint64_t) 7;
            block_len = lsb_read(&z->b, 8); block_len |= lsb_read(&z->b, 8) << 8;
            if (z->b.err) { ret = ERR_READ; goto out; }
            if (block_len == 0) break;
            c = lsb_read(&z->b, 8); if (z->b.err) { ret = ERR_READ; goto out; } if (c != 'C') { ret = MSGPU_ERR_DATAFORMAT; goto out; }
            c = lsb_read(&z->b, 8); if (z->b.err) { ret = ERR_READ; goto out; } if (c != 'K') { ret = MSGPU_ERR_DATAFORMAT; goto out; }
            z->window_posn = 0; z->bytes_output = 0;
            error = zip_inflate(z);
            if (error) { ret = (error > 0) ? error : ERR_DECRUNCH; goto out; }
            n = z->bytes_output;
            if (n > u->out_len - done) { memcpy(out + done, z->window, u->out_len - done); done = u->out_len; ret = MSGPU_ERR_CAPACITY; goto out; }
            memcpy(out + done, z->window, n); done += n;
        }
        goto out;
    }
    while (done < u->out_len) {
        int state = 0, error; uint32_t n;
        z->b.p = (z->b.p + 7) & ~(uint64_t) 7;                    /* :405 align to bytestream */
        do {                                                      /* :406-413 skip to the next 'CK' */
            uint32_t c = lsb_read(&z->b, 8);
            if (z->b.err) { ret = ERR_READ; goto out; }
            if (c == 'C') state = 1;
            else if (state == 1 && c == 'K') state = 2;
            else state = 0;
        } while (state != 2);
        z->window_posn = 0; z->bytes_output = 0;
        ZIP_STORE(z);                                             /* :419 */
        error = zip_inflate(z);
        if (error) {
            if (u->flags & MSGPU_FLAG_MSZIP_REPAIR) {             /* :420-433 */
                if (z->bytes_output == 0 && z->window_posn > 0) { z->bytes_output += z->window_posn; }
                if (z->bytes_output < FRAME) memset(z->window + z->bytes_output, 0, FRAME - z->bytes_output);
                z->bytes_output = FRAME;
                if (error < 0) zip_repair_restart(z, u);
            }
            else { ret = (error > 0) ? error : ERR_DECRUNCH; goto out; }
        }
        n = u->out_len - done; if (n > z->bytes_output) n = z->bytes_output;
        memcpy(out + done, z->window, n);
        if (error > 0 && (u->flags & MSGPU_FLAG_MSZIP_REPAIR)) { done += n; ret = error; goto out; }  /* :448 */
        done += n;
    }
out:
    if (produced) *produced = done;
    free(z);
    return ret;
}

/* ==========================================================================================
 * LZX  (lzxd.c)
 * ========================================================================================== */
static const uint16_t lzx_position_slots[11] = { 30, 32, 34, 36, 38, 42, 50, 66, 98, 162, 290 };   /* lzxd.c:209-211 (wb 15..25) */
static uint32_t lzx_position_base[290]; static uint8_t lzx_extra_bits[290]; static int lzx_tables_ready;
static void lzx_make_tables(void) {         /* lzxd.c:199-207: the rule the static tables were generated by */
    unsigned i; uint32_t base = 0;
    for (i = 0; i < 290; i++) {
        unsigned e = (i < 4) ? 0 : ((i < 36) ? (i / 2 - 1) : 17);
        lzx_extra_bits[i] = (uint8_t) e; lzx_position_base[i] = base; base += 1u << e;
    }
    lzx_tables_ready = 1;
}

#define LZX_MAIN_SYMS 2576
#define LZX_LEN_SYMS  250
#define LZX_SAFETY    64
typedef struct {
    bitin b;
    uint32_t R0, R1, R2;
    uint32_t block_type, block_length, block_remaining;
    int header_read, intel_started; int32_t intel_filesize;
    uint32_t num_offsets, window_size;
    uint8_t pre_len[20 + LZX_SAFETY], main_len[LZX_MAIN_SYMS + LZX_SAFETY],
            len_len[LZX_LEN_SYMS + LZX_SAFETY], aligned_len[8 + LZX_SAFETY];
    huff pre, main, length, aligned; int length_empty;
} lzxst;

static void lzx_reset_state(lzxst *s) {              /* lzxd.c:257-270 */
    s->R0 = s->R1 = s->R2 = 1; s->header_read = 0; s->block_remaining = 0; s->block_type = 0;
    memset(s->main_len, 0, LZX_MAIN_SYMS); memset(s->len_len, 0, LZX_LEN_SYMS);
}

static inline unsigned lzx_huffsym(lzxst *s, const huff *h) {
    unsigned len, sym;
    if (!lzx_ensure(&s->b, 16)) return 0;
    sym = huff_decode(h, lzx_peek(&s->b, 16), &len);
    s->b.p += len;
    return sym;
}

/* lzxd.c:138-183 lzxd_read_lens: pretree-delta coded lengths; runs are NOT clamped to `last` */
static int lzx_read_lens(lzxst *s, uint8_t *lens, unsigned first, unsigned last) {
    unsigned x, y; int z;
    for (x = 0; x < 20; x++) s->pre_len[x] = (uint8_t) lzx_read(&s->b, 4);
    if (s->b.err) return ERR_READ;
    if (huff_build(&s->pre, s->pre_len, 20, 6)) return ERR_DECRUNCH;
    for (x = first; x < last;) {
        z = (int) lzx_huffsym(s, &s->pre);
        if (s->b.err) return ERR_READ;
        if (z == 17) { y = lzx_read(&s->b, 4) + 4; if (s->b.err) return ERR_READ; while (y--) lens[x++] = 0; }
        else if (z == 18) { y = lzx_read(&s->b, 5) + 20; if (s->b.err) return ERR_READ; while (y--) lens[x++] = 0; }
        else if (z == 19) {
            y = lzx_read(&s->b, 1) + 4; if (s->b.err) return ERR_READ;
            z = (int) lzx_huffsym(s, &s->pre); if (s->b.err) return ERR_READ;
            z = lens[x] - z; if (z < 0) z += 17;
            while (y--) lens[x++] = (uint8_t) z;
        }
        else { z = lens[x] - z; if (z < 0) z += 17; lens[x++] = (uint8_t) z; }
    }
    return ERR_OK;
}

/* per-frame record for the E8 post-pass */
typedef struct { uint32_t start, size; int32_t filesize; int active; } lzxframe;

/* lzxd.c:706-737, applied in place to a finished frame (the window keeps the raw bytes in the
 * reference; here every match of later frames has already been resolved when this runs) */
static void lzx_e8_frame(uint8_t *data, uint32_t frame_size, int32_t curpos, int32_t filesize) {
    uint32_t i = 0, end = frame_size - 10;
    while (i < end) {
        int32_t abs_off, rel_off;
        if (data[i++] != 0xE8) { curpos++; continue; }
        abs_off = (int32_t) ((uint32_t) data[i] | ((uint32_t) data[i + 1] << 8) | ((uint32_t) data[i + 2] << 16) | ((uint32_t) data[i + 3] << 24));
        if (abs_off >= -curpos && abs_off < filesize) {
            rel_off = (abs_off >= 0) ? abs_off - curpos : abs_off + filesize;
            data[i] = (uint8_t) rel_off; data[i + 1] = (uint8_t) (rel_off >> 8);
            data[i + 2] = (uint8_t) (rel_off >> 16); data[i + 3] = (uint8_t) (rel_off >> 24);
        }
        i += 4; curpos += 5;
    }
}

/* lzxd.c:441-444: LZX DELTA has a 16-bit chunk size in front of every frame; ENSURE_BITS(16) + REMOVE_BITS(16).  With an
 * empty bit buffer (after the raw bytes of an uncompressed block) that is one two-byte fetch from the byte pointer,
 * wherever it stands, and the buffer is empty again afterwards */
static int lzx_skip_chunk_size(bitin *b) {
    if (b->bytemode) { if (!fetch_ok(b, (b->p >> 3) + 2)) return 0; b->p += 16; return 1; }
    if (!lzx_ensure(b, 16)) return 0;
    b->p += 16; return 1;
}

/* lzxd.c:388-771 lzxd_decompress for one whole unit: fresh state, output_length == out_len */
static int port_lzx(const msgpu_unit *u, const uint8_t *in, uint8_t *out, uint32_t *produced) {
    lzxst *s; lzxframe *frames = NULL; uint32_t nframes_cap, nframes = 0, G = 0, frame = 0, f;
    int ret = ERR_OK;
    const int is_delta = (u->flags & MSGPU_FLAG_LZX_DELTA) != 0;
    const uint32_t ref_len = MSGPU_UNIT_REF_BYTES(u);                   /* lzxd_set_reference_data, lzxd.c:348-382: sits at out[-ref_len..0) */
    /* lzxd_init returns NULL (cabd.c:1255 turns that into NOMEMORY): lzxd.c:289-296 */
    if (is_delta ? (u->window_bits < 17 || u->window_bits > 25) : (u->window_bits < 15 || u->window_bits > 21)) return ERR_NOMEM;
    if (ref_len && (!is_delta || ref_len > (1u << u->window_bits))) return ERR_ARGS;          /* lzxd.c:355-366 */
    if (!lzx_tables_ready) lzx_make_tables();
    s = (lzxst *) calloc(1, sizeof(lzxst));
    nframes_cap = u->out_len / FRAME + 2;
    frames = (lzxframe *) calloc(nframes_cap, sizeof(lzxframe));
    if (!s || !frames) { free(s); free(frames); return ERR_NOMEM; }
    s->b.in = in; s->b.in_len = u->in_len;
    s->window_size = 1u << u->window_bits;
    s->num_offsets = (uint32_t) lzx_position_slots[u->window_bits - 15] << 3;
    lzx_reset_state(s);

    while (G < u->out_len) {
        uint32_t frame_start = G, frame_size, window_posn_ref;
        int32_t bytes_todo;
        if (u->reset_interval && (frame % u->reset_interval) == 0) lzx_reset_state(s);   /* :423-438 */
        if (is_delta && !lzx_skip_chunk_size(&s->b)) { ret = ERR_READ; goto out; }      /* :441-444 */
        if (!s->header_read) {                                                          /* :447-453 */
            uint32_t i = 0, j = 0;
            lzx_enter_bits(&s->b);
            if (lzx_read(&s->b, 1)) { i = lzx_read(&s->b, 16); j = lzx_read(&s->b, 16); }
            if (s->b.err) { ret = ERR_READ; goto out; }
            s->intel_filesize = (int32_t) ((i << 16) | j); s->header_read = 1;
        }
        frame_size = FRAME; if (u->out_len - G < FRAME) frame_size = u->out_len - G;     /* :458-461 */
        bytes_todo = (int32_t) frame_size;
        while (bytes_todo > 0) {
            int32_t this_run;
            if (s->block_remaining == 0) {
                uint32_t i, j;
                if (s->block_type == 3 && (s->block_length & 1)) {                      /* :469-474 */
                    if (!fetch_ok(&s->b, (s->b.p >> 3) + 1)) { ret = ERR_READ; goto out; }
                    s->b.p += 8;
                }
                lzx_enter_bits(&s->b);
                s->block_type = lzx_read(&s->b, 3); i = lzx_read(&s->b, 16); j = lzx_read(&s->b, 8);   /* :477-479 */
                if (s->b.err) { ret = ERR_READ; goto out; }
                s->block_remaining = s->block_length = (i << 8) | j;
                switch (s->block_type) {
                case 2:
                    for (i = 0; i < 8; i++) s->aligned_len[i] = (uint8_t) lzx_read(&s->b, 3);
                    if (s->b.err) { ret = ERR_READ; goto out; }
                    if (huff_build(&s->aligned, s->aligned_len, 8, 7)) { ret = ERR_DECRUNCH; goto out; }
                    /* fallthrough */
                case 1:
                    if ((ret = lzx_read_lens(s, s->main_len, 0, 256))) goto out;
                    if ((ret = lzx_read_lens(s, s->main_len, 256, 256 + s->num_offsets))) goto out;
                    if (huff_build(&s->main, s->main_len, LZX_MAIN_SYMS, 12)) { ret = ERR_DECRUNCH; goto out; }
                    if (s->main_len[0xE8] != 0) s->intel_started = 1;                   /* :495 */
                    if ((ret = lzx_read_lens(s, s->len_len, 0, 249))) goto out;
                    s->length_empty = 0;
                    if (huff_build(&s->length, s->len_len, LZX_LEN_SYMS, 12)) {         /* :111-125 */
                        for (i = 0; i < LZX_LEN_SYMS; i++) if (s->len_len[i] > 0) { ret = ERR_DECRUNCH; goto out; }
                        s->length_empty = 1;
                    }
                    break;
                case 3: {
                    uint8_t buf[12];
                    s->intel_started = 1;                                               /* :503 */
                    /* :505-507 read 1-16 bits to align to the next 16-bit word */
                    if ((s->b.p & 15) == 0) { if (!lzx_ensure(&s->b, 16)) { ret = ERR_READ; goto out; } s->b.p += 16; }
                    else s->b.p = (s->b.p + 15) & ~(uint64_t) 15;
                    s->b.bytemode = 1;
                    for (i = 0; i < 12; i++) {
                        if (!fetch_ok(&s->b, (s->b.p >> 3) + 1)) { ret = ERR_READ; goto out; }
                        buf[i] = (uint8_t) in_byte(&s->b, s->b.p >> 3); s->b.p += 8;
                    }
                    s->R0 = buf[0] | (buf[1] << 8) | (buf[2] << 16) | ((uint32_t) buf[3] << 24);
                    s->R1 = buf[4] | (buf[5] << 8) | (buf[6] << 16) | ((uint32_t) buf[7] << 24);
                    s->R2 = buf[8] | (buf[9] << 8) | (buf[10] << 16) | ((uint32_t) buf[11] << 24);
                    break;
                }
                default: ret = ERR_DECRUNCH; goto out;                                  /* :519-522 */
                }
            }
            this_run = (int32_t) s->block_remaining;
            if (this_run > bytes_todo) this_run = bytes_todo;
            bytes_todo -= this_run; s->block_remaining -= (uint32_t) this_run;

            if (s->block_type == 1 || s->block_type == 2) {
                while (this_run > 0) {                                                  /* :538-651 */
                    unsigned main_element = lzx_huffsym(s, &s->main);
                    if (s->b.err) { ret = ERR_READ; goto out; }
                    if (main_element < 256) { out[G++] = (uint8_t) main_element; this_run--; }
                    else {
                        uint32_t match_length, match_offset, slot, eff, k;
                        main_element -= 256;
                        match_length = main_element & 7;
                        if (match_length == 7) {
                            if (s->length_empty) { ret = ERR_DECRUNCH; goto out; }
                            match_length += lzx_huffsym(s, &s->length);
                            if (s->b.err) { ret = ERR_READ; goto out; }
                        }
                        match_length += 2;
                        slot = main_element >> 3;
                        if (slot == 0) match_offset = s->R0;
                        else if (slot == 1) { match_offset = s->R1; s->R1 = s->R0; s->R0 = match_offset; }
                        else if (slot == 2) { match_offset = s->R2; s->R2 = s->R0; s->R0 = match_offset; }
                        else {
                            unsigned extra = (slot >= 36) ? 17 : lzx_extra_bits[slot];
                            match_offset = lzx_position_base[slot] - 2;
                            if (extra >= 3 && s->block_type == 2) {
                                if (extra > 3) match_offset += lzx_read(&s->b, extra - 3) << 3;
                                if (s->b.err) { ret = ERR_READ; goto out; }
                                match_offset += lzx_huffsym(s, &s->aligned);
                            }
                            else if (extra) match_offset += lzx_read(&s->b, extra);
                            if (s->b.err) { ret = ERR_READ; goto out; }
                            s->R2 = s->R1; s->R1 = s->R0; s->R0 = match_offset;
                        }
                        if (is_delta && match_length == 257) {                          /* :589-611 longer matches */
                            uint32_t extra_len;
                            if (!lzx_ensure(&s->b, 3)) { ret = ERR_READ; goto out; }
                            if (lzx_peek(&s->b, 1) == 0) { s->b.p += 1; extra_len = lzx_read(&s->b, 8); }
                            else if (lzx_peek(&s->b, 2) == 2) { s->b.p += 2; extra_len = lzx_read(&s->b, 10) + 0x100; }
                            else if (lzx_peek(&s->b, 3) == 6) { s->b.p += 3; extra_len = lzx_read(&s->b, 12) + 0x500; }
                            else { s->b.p += 3; extra_len = lzx_read(&s->b, 15); }
                            if (s->b.err) { ret = ERR_READ; goto out; }
                            match_length += extra_len;
                        }
                        /* :613-634 bounds checks, restated for a linear output buffer: the reference's
                         * window_posn is G modulo the window size, its lzx->offset is frame_start */
                        window_posn_ref = G & (s->window_size - 1);
                        if (window_posn_ref + match_length > s->window_size) { ret = ERR_DECRUNCH; goto out; }
                        eff = match_offset;
                        if (match_offset > window_posn_ref) {
                            if (match_offset > frame_start && match_offset - window_posn_ref > ref_len) { ret = ERR_DECRUNCH; goto out; }   /* :622-628 */
                            if (match_offset - window_posn_ref > s->window_size) { ret = ERR_DECRUNCH; goto out; }
                            if (match_offset > s->window_size) eff = match_offset - s->window_size; /* lands in the current lap */
                        }
                        if ((uint64_t) G + match_length > u->out_len) { ret = ERR_DECRUNCH; goto out; } /* frame overrun, :689-693 */
                        /* bytes in front of the unit: the reference data (the end of the reference's window), zero before that */
                        for (k = 0; k < match_length; k++) { out[G] = (eff <= G || eff - G <= ref_len) ? *(out + G - eff) : 0; G++; }
                        this_run -= (int32_t) match_length;
                    }
                }
            }
            else if (s->block_type == 3) {                                              /* :654-671 */
                while (this_run > 0) {
                    if (!fetch_ok(&s->b, (s->b.p >> 3) + 1)) { ret = ERR_READ; goto out; }
                    out[G++] = (uint8_t) in_byte(&s->b, s->b.p >> 3); s->b.p += 8; this_run--;
                }
            }
            else { ret = ERR_DECRUNCH; goto out; }
            if (this_run < 0) {                                                         /* :678-685 */
                if ((uint32_t) (-this_run) > s->block_remaining) { ret = ERR_DECRUNCH; goto out; }
                s->block_remaining -= (uint32_t) (-this_run);
            }
        }
        if (G - frame_start != frame_size) { ret = ERR_DECRUNCH; goto out; }            /* :689-693 */
        /* :696-697 re-align to the next 16-bit word (bits_left > 0 => the reference tops up first) */
        /* after raw bytes of an uncompressed block the reference's bit buffer is empty (bits_left == 0) and
         * the realign is a no-op even if the byte pointer is odd; the pad byte goes with the next header */
        if (!s->b.bytemode && (s->b.p & 15)) { if (!lzx_ensure(&s->b, 16)) { ret = ERR_READ; goto out; } s->b.p = (s->b.p + 15) & ~(uint64_t) 15; }
        frames[nframes].start = frame_start; frames[nframes].size = frame_size;
        frames[nframes].filesize = s->intel_filesize;
        frames[nframes].active = (s->intel_started && s->intel_filesize && frame < 32768 && frame_size > 10);
        nframes++; frame++;
    }
    /* lzxd.c:419: end_frame = (offset + out_bytes) / 32768 + 1, so a request that ends exactly on a frame
     * boundary runs one more, zero-sized, frame pass.  It decodes nothing, but when that frame index is a
     * reset point it re-reads the intel header (1 or 33 bits, :447-453) and then tops the bit buffer up to
     * re-align (:696-697) - two places where a unit cut exactly at its last byte reports MSPACK_ERR_READ
     * although every output byte has been produced. */
    if ((u->out_len % FRAME) == 0 && u->out_len && is_delta) {
        if (u->reset_interval && (frame % u->reset_interval) == 0) lzx_reset_state(s);
        if (!lzx_skip_chunk_size(&s->b)) { ret = ERR_READ; goto e8; }                   /* the extra pass reads its chunk size too */
    }
    if ((u->out_len % FRAME) == 0 && u->reset_interval && (frame % u->reset_interval) == 0) {
        uint64_t bp;
        lzx_enter_bits(&s->b);
        bp = s->b.p >> 3;                                /* p is 16-bit aligned here */
        if (!fetch_ok(&s->b, bp + 2)) { ret = ERR_READ; goto e8; }
        if (in_byte(&s->b, bp + 1) & 0x80) { if (!fetch_ok(&s->b, bp + 8)) { ret = ERR_READ; goto e8; } }
        else if (!fetch_ok(&s->b, bp + 4)) { ret = ERR_READ; goto e8; }
    }
e8:
    for (f = 0; f < nframes; f++)
        if (frames[f].active) lzx_e8_frame(out + frames[f].start, frames[f].size, (int32_t) frames[f].start, frames[f].filesize);
out:
    if (produced) *produced = (ret == ERR_OK || G == u->out_len) ? u->out_len : 0;
    free(frames); free(s);
    return ret;
}

/* ==========================================================================================
 * Quantum  (qtmd.c)
 * ========================================================================================== */
static uint32_t qtm_position_base[42]; static uint8_t qtm_extra_bits[42];
static uint8_t qtm_length_base[27], qtm_length_extra[27]; static int qtm_tables_ready;
static void qtm_make_tables(void) {           /* qtmd.c:52-64: the generator given in the comment */
    unsigned i; uint32_t off;
    for (i = 0, off = 0; i < 42; i++) { qtm_position_base[i] = off; qtm_extra_bits[i] = (uint8_t) (((i < 2) ? 0 : (i - 2)) >> 1); off += 1u << qtm_extra_bits[i]; }
    for (i = 0, off = 0; i < 26; i++) { qtm_length_base[i] = (uint8_t) off; qtm_length_extra[i] = (uint8_t) ((i < 2 ? 0 : i - 2) >> 2); off += 1u << qtm_length_extra[i]; }
    qtm_length_base[26] = 254; qtm_length_extra[26] = 0;
    qtm_tables_ready = 1;
}

typedef struct { uint16_t sym, cumfreq; } qsym;
typedef struct { int shiftsleft, entries; qsym syms[65]; } qmodel;
typedef struct { bitin b; uint16_t H, L, C; qmodel m0, m1, m2, m3, m4, m5, m6, m6len, m7; } qtmst;

static void qtm_init_model(qmodel *m, int start, int len) {       /* qtmd.c:169-182 */
    int i; m->shiftsleft = 4; m->entries = len;
    for (i = 0; i <= len; i++) { m->syms[i].sym = (uint16_t) (start + i); m->syms[i].cumfreq = (uint16_t) (len - i); }
}

static void qtm_update_model(qmodel *m) {                         /* qtmd.c:125-166 */
    qsym tmp; int i, j;
    if (--m->shiftsleft) {
        for (i = m->entries - 1; i >= 0; i--) {
            m->syms[i].cumfreq >>= 1;
            if (m->syms[i].cumfreq <= m->syms[i + 1].cumfreq) m->syms[i].cumfreq = (uint16_t) (m->syms[i + 1].cumfreq + 1);
        }
    }
    else {
        m->shiftsleft = 50;
        for (i = 0; i < m->entries; i++) {
            m->syms[i].cumfreq = (uint16_t) (m->syms[i].cumfreq - m->syms[i + 1].cumfreq);
            m->syms[i].cumfreq++; m->syms[i].cumfreq >>= 1;
        }
        for (i = 0; i < m->entries - 1; i++)
            for (j = i + 1; j < m->entries; j++)
                if (m->syms[i].cumfreq < m->syms[j].cumfreq) { tmp = m->syms[i]; m->syms[i] = m->syms[j]; m->syms[j] = tmp; }
        for (i = m->entries - 1; i >= 0; i--) m->syms[i].cumfreq = (uint16_t) (m->syms[i].cumfreq + m->syms[i + 1].cumfreq);
    }
}

/* qtmd.c:92-123 GET_SYMBOL; integer widths as in the reference (u16 H/L/C/symf, u32 range) */
static int qtm_get_symbol(qtmst *q, qmodel *m) {
    uint32_t range; uint16_t symf; int i, sym;
    range = ((uint32_t) (q->H - q->L) & 0xFFFFu) + 1u;
    symf = (uint16_t) ((((uint32_t) (((int) q->C - (int) q->L + 1) * (int) m->syms[0].cumfreq - 1)) / range) & 0xFFFFu);
    for (i = 1; i < m->entries; i++) if (m->syms[i].cumfreq <= symf) break;
    sym = m->syms[i - 1].sym;
    range = (uint32_t) ((int) q->H - (int) q->L + 1);
    symf = m->syms[0].cumfreq;
    q->H = (uint16_t) (q->L + (((uint32_t) m->syms[i - 1].cumfreq * range) / symf) - 1);
    q->L = (uint16_t) (q->L + (((uint32_t) m->syms[i].cumfreq * range) / symf));
    do { --i; m->syms[i].cumfreq = (uint16_t) (m->syms[i].cumfreq + 8); } while (i > 0);
    if (m->syms[0].cumfreq > 3800) qtm_update_model(m);
    for (;;) {
        if ((q->L & 0x8000) != (q->H & 0x8000)) {
            if ((q->L & 0x4000) && !(q->H & 0x4000)) { q->C ^= 0x4000; q->L &= 0x3FFF; q->H |= 0x4000; }
            else break;
        }
        q->L = (uint16_t) (q->L << 1); q->H = (uint16_t) ((q->H << 1) | 1);
        q->C = (uint16_t) ((q->C << 1) | qtm_read(&q->b, 1));
        if (q->b.err) return -1;
    }
    return sym;
}

/* qtmd.c:257-479 qtmd_decompress for one whole unit (single request of out_len bytes) */
static int port_qtm(const msgpu_unit *u, const uint8_t *in, uint8_t *out, uint32_t *produced) {
    qtmst *q; uint32_t G = 0, frame_todo = FRAME, window_size; int header_read = 0, ret = ERR_OK, wb2;
    if (u->window_bits < 10 || u->window_bits > 21) return ERR_NOMEM;    /* qtmd_init returns NULL */
    if (!qtm_tables_ready) qtm_make_tables();
    q = (qtmst *) calloc(1, sizeof(qtmst));
    if (!q) return ERR_NOMEM;
    q->b.in = in; q->b.in_len = u->in_len;
    window_size = 1u << u->window_bits; wb2 = u->window_bits * 2;
    qtm_init_model(&q->m0, 0, 64); qtm_init_model(&q->m1, 64, 64); qtm_init_model(&q->m2, 128, 64); qtm_init_model(&q->m3, 192, 64);
    qtm_init_model(&q->m4, 0, wb2 > 24 ? 24 : wb2); qtm_init_model(&q->m5, 0, wb2 > 36 ? 36 : wb2);
    qtm_init_model(&q->m6, 0, wb2); qtm_init_model(&q->m6len, 0, 27); qtm_init_model(&q->m7, 0, 7);

    while (G < u->out_len) {
        uint32_t frame_end;
        if (!header_read) {                                                         /* :290-295 */
            q->H = 0xFFFF; q->L = 0; q->C = (uint16_t) qtm_read(&q->b, 16);
            if (q->b.err) { ret = ERR_READ; goto out; }
            header_read = 1;
        }
        frame_end = u->out_len;                                                     /* :299-305 (window end handled below) */
        if (G + frame_todo < frame_end) frame_end = G + frame_todo;
        while (G < frame_end) {
            int selector = qtm_get_symbol(q, &q->m7), sym;
            if (selector < 0) { ret = ERR_READ; goto out; }
            if (selector < 4) {
                qmodel *mdl = (selector == 0) ? &q->m0 : ((selector == 1) ? &q->m1 : ((selector == 2) ? &q->m2 : &q->m3));
                if ((sym = qtm_get_symbol(q, mdl)) < 0) { ret = ERR_READ; goto out; }
                out[G++] = (uint8_t) sym; frame_todo--;
            }
            else {
                uint32_t match_offset, match_length, extra, window_posn, k;
                switch (selector) {
                case 4:
                    if ((sym = qtm_get_symbol(q, &q->m4)) < 0) { ret = ERR_READ; goto out; }
                    extra = qtm_read_many(&q->b, qtm_extra_bits[sym]);
                    match_offset = qtm_position_base[sym] + extra + 1; match_length = 3; break;
                case 5:
                    if ((sym = qtm_get_symbol(q, &q->m5)) < 0) { ret = ERR_READ; goto out; }
                    extra = qtm_read_many(&q->b, qtm_extra_bits[sym]);
                    match_offset = qtm_position_base[sym] + extra + 1; match_length = 4; break;
                case 6:
                    if ((sym = qtm_get_symbol(q, &q->m6len)) < 0) { ret = ERR_READ; goto out; }
                    extra = qtm_read_many(&q->b, qtm_length_extra[sym]);
                    match_length = qtm_length_base[sym] + extra + 5;
                    if ((sym = qtm_get_symbol(q, &q->m6)) < 0) { ret = ERR_READ; goto out; }
                    extra = qtm_read_many(&q->b, qtm_extra_bits[sym]);
                    match_offset = qtm_position_base[sym] + extra + 1; break;
                default: ret = ERR_DECRUNCH; goto out;
                }
                if (q->b.err) { ret = ERR_READ; goto out; }
                frame_todo -= match_length;                                          /* unsigned wrap checked below */
                window_posn = G & (window_size - 1);
                if (window_posn + match_length > window_size) {
                    /* :358-390 match crossing the end of the window: the reference must flush the whole
                     * window first and bails out if that is more than the caller still wants */
                    uint32_t lap_start = G - window_posn;
                    if ((uint64_t) lap_start + window_size > u->out_len) { ret = ERR_DECRUNCH; goto out; }
                }
                else if (match_offset > window_posn && match_offset - window_posn > window_size) { ret = ERR_DECRUNCH; goto out; } /* :398-401 */
                for (k = 0; k < match_length; k++) {
                    uint8_t v = (match_offset <= G) ? out[G - match_offset] : 0;
                    if (G < u->out_len) out[G] = v;
                    else break;                       /* surplus beyond the request is never observable */
                    G++;
                }
                G += match_length - k;
                if (window_posn + match_length > window_size) break;                /* :389 */
            }
        }
        if (frame_todo > FRAME) { ret = ERR_DECRUNCH; goto out; }                   /* :424-427 */
        if (frame_todo == 0) {                                                      /* :430-442 */
            uint32_t c;
            q->b.p = (q->b.p + 7) & ~(uint64_t) 7;
            do { c = qtm_read(&q->b, 8); if (q->b.err) { ret = ERR_READ; goto out; } } while (c != 0xFF);
            header_read = 0; frame_todo = FRAME;
        }
    }
out:
    if (produced) *produced = (ret == ERR_OK) ? u->out_len : 0;
    free(q);
    return ret;
}

/* ==========================================================================================
 * entry points (same shape as oracle/ref_harness.c)
 * ========================================================================================== */
int oracle_port_decode(const msgpu_unit *u, const unsigned char *in_base, unsigned char *out_base, uint32_t *produced) {
    const uint8_t *in = in_base + u->in_off; uint8_t *out = out_base + u->out_off;
    switch (u->codec) {
    case MSGPU_CODEC_MSZIP:   return port_mszip(u, in, out, produced);
    case MSGPU_CODEC_QUANTUM: return port_qtm(u, in, out, produced);
    case MSGPU_CODEC_LZX:     return port_lzx(u, in, out, produced);
    default: return ERR_ARGS;
    }
}

struct port_job { const msgpu_unit *units; size_t lo, hi; const unsigned char *in_base; unsigned char *out_base; int32_t *status; };
static void *port_worker(void *arg) {
    struct port_job *j = (struct port_job *) arg; size_t i;
    for (i = j->lo; i < j->hi; i++) {
        int e = oracle_port_decode(&j->units[i], j->in_base, j->out_base, NULL);
        if (j->status) j->status[i] = e;
    }
    return NULL;
}

double oracle_port_decode_batch(const msgpu_unit *units, size_t n, const unsigned char *in_base,
                                unsigned char *out_base, int32_t *status, int threads) {
    struct timespec t0, t1; pthread_t *tid; struct port_job *jobs; int t;
    if (!lzx_tables_ready) lzx_make_tables();
    if (!qtm_tables_ready) qtm_make_tables();
    if (threads < 1) threads = 1;
    if ((size_t) threads > n && n > 0) threads = (int) n;
    tid = (pthread_t *) calloc((size_t) threads, sizeof(*tid));
    jobs = (struct port_job *) calloc((size_t) threads, sizeof(*jobs));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < threads; t++) {
        jobs[t].units = units; jobs[t].lo = n * (size_t) t / (size_t) threads; jobs[t].hi = n * (size_t) (t + 1) / (size_t) threads;
        jobs[t].in_base = in_base; jobs[t].out_base = out_base; jobs[t].status = status;
        if (threads == 1) port_worker(&jobs[t]); else pthread_create(&tid[t], NULL, port_worker, &jobs[t]);
    }
    if (threads > 1) for (t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(tid); free(jobs);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

const char *oracle_port_version(void) { return "libmspack_b200 oracle port (plain-C restatement of lzxd.c, qtmd.c, mszipd.c)"; }
